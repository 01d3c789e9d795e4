"""CPU: the oracle restatement against the golden vectors produced by the reference itself."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, golden_input
from cases import CASES, SYNC_THRESHOLDS
from oracle import ref_dsp


@pytest.mark.parametrize("name,gen,n,fs,fo", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(name, gen, n, fs, fo):
    g = load_golden(name)
    x = golden_input(g, **gen)
    r = ref_dsp.process(x.astype(np.complex128), fo, fs)
    assert np.array_equal(r["dibits"], g["dibits"])
    assert r["symbols"].shape == g["symbols"].shape
    if len(g["symbols"]):
        assert np.array_equal(r["symbols"], g["symbols"])      # same SciPy calls: bit-identical
        assert r["best_phase"] == int(g["best_phase"])
    if len(g["dibits"]):
        bits = ref_dsp.symbols_to_bits(r["dibits"])
        assert hashlib.sha256(bits.astype(np.uint8)).hexdigest() == str(g["bits_sha256"])
        for th in SYNC_THRESHOLDS:
            pos, mx = ref_dsp.find_sync(bits, th)
            assert pos == list(g["sync_pos_%03d" % round(th * 100)])
            assert mx == float(g["sync_max_%03d" % round(th * 100)])


def test_find_sync_planted_pattern():
    # the reference's own unit test plants TS1 at bit 20 (tests/unit/test_tetra_decoder.py:56-66)
    bits = np.zeros(600, dtype=np.int64)
    bits[20:42] = ref_dsp.TS1
    pos, mx = ref_dsp.find_sync(bits, 0.85)
    assert pos == [20] and mx == 1.0
    assert ref_dsp.find_sync(np.zeros(10, dtype=np.int64)) == ([], 0.0)


def test_slicer_regions_q7():
    # regions are centred on 0, +pi/2, -pi/2, pi
    ph = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi + 1e-9, 3 * np.pi / 8 - 1e-9, 3 * np.pi / 8 + 1e-9])
    s = np.concatenate([[1.0 + 0j], np.exp(1j * np.cumsum(ph))])
    d, _ = ref_dsp.slice_dqpsk(s)
    assert list(d) == [0, 1, 2, 3, 3, 0, 1]


def test_stft_rows_shape():
    x = (np.random.default_rng(0).standard_normal(10000) + 0j).astype(np.complex64)
    assert ref_dsp.stft_db(x, 4096, 1024).shape == (6, 4096)
    assert ref_dsp.stft_db(x[:100], 4096, 1024).shape == (0, 4096)


def test_burst_parse_matches_reference_golden():
    """oracle parse_burst / check_crc vs TetraProtocolParser.parse_burst run on the same seeded slots."""
    g = load_golden("bursts")
    for f, bt, ok in zip(g["frames"], g["burst_type"], g["crc_ok"]):
        btype, crc_ok, data = ref_dsp.parse_burst(f)
        assert (btype, int(crc_ok)) == (int(bt), int(ok))
        assert len(data) == (510 if btype == ref_dsp.BURST_SYNCHRONIZATION else 216)
    assert g["crc_ok"].sum() > 10 and (g["burst_type"] == 5).sum() > 10


def test_scanner_analysis_matches_reference_golden():
    """oracle analyze_signal vs TetraSignalDetector's own methods on the same seeded captures."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from oracle.make_golden_scanner import captures
    g = load_golden("scanner")
    for name, x in captures():
        r = ref_dsp.analyze_signal(x.astype(np.complex128))
        want = g[name]
        assert abs(r["power_db"] - want[0]) < 1e-9, name
        assert r["modulation_confidence"] == want[1] and float(r["is_tetra_modulation"]) == want[2], name
        assert r["sync_correlation"] == want[3] and float(r["sync_detected"]) == want[4], name
        assert float(r["power_stable"]) == want[5], name


def test_presence_afc_restatement_on_a_centred_carrier():
    """oracle/ref_dsp.presence_afc (ui/modern.py:1945-2012): a carrier at the centre is a strong signal whose AFC offset stays
    inside the channel; noise alone is not; the survey golden holds both kinds."""
    from tetraear_b200 import synth
    x = synth.carrier_iq(8192, 71, snr_db=30.0).astype(np.complex128)
    p = ref_dsp.presence_afc(x, 2.4e6)
    assert p["is_signal_strong"] and abs(p["peak_freq_offset"]) <= 12500 and p["snr"] > 15
    rng = np.random.default_rng(5)
    noise = 0.01 * (rng.standard_normal(4096) + 1j * rng.standard_normal(4096))
    assert not ref_dsp.presence_afc(noise, 2.4e6)["is_signal_strong"]
    assert ref_dsp.presence_afc(noise[:100], 2.4e6)["snr"] == 0.0
    g = load_golden("survey")
    col = {str(n): i for i, n in enumerate(g["fields"])}
    rows, active = g["rows"], g["active"]
    assert rows[active, col["is_signal_strong"]].mean() > 0.9 and rows[~active, col["is_signal_strong"]].mean() < 0.1
    v = ref_dsp.analyze_verdict(*[float(rows[0, col[k]]) for k in ("power_db", "modulation_confidence", "sync_correlation", "power_stable")])
    assert float(v["is_tetra"]) == rows[0, col["is_tetra"]] and abs(v["confidence"] - rows[0, col["confidence"]]) < 1e-12
