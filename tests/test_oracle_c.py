"""CPU: the plain-C restatement (oracle/tetra_oracle.c) against the golden vectors produced by the reference itself
and against the SciPy-calling oracle."""
import numpy as np
import pytest
from scipy import signal

from conftest import load_golden, golden_input
from cases import CASES, SYNC_THRESHOLDS
from oracle import c_oracle, ref_dsp


@pytest.mark.parametrize("name,gen,n,fs,fo", CASES, ids=[c[0] for c in CASES])
def test_c_oracle_matches_reference_golden(name, gen, n, fs, fo):
    g = load_golden(name)
    x = golden_input(g, **gen)
    r = c_oracle.process(x.astype(np.complex128), fo, fs)
    assert np.array_equal(r["dibits"], g["dibits"])
    assert r["symbols"].shape == g["symbols"].shape
    if len(g["symbols"]):
        assert r["best_phase"] == int(g["best_phase"])
        err = np.abs(r["symbols"] - g["symbols"]).max() / np.abs(g["symbols"]).max()
        assert err < 1e-11, err                    # same recursions, independently designed coefficients
    if len(g["dibits"]):
        bits = c_oracle.symbols_to_bits(r["dibits"])
        assert np.array_equal(bits, ref_dsp.symbols_to_bits(g["dibits"]))
        for th in SYNC_THRESHOLDS:
            pos, mx = c_oracle.find_sync(bits, th)
            assert pos == list(g["sync_pos_%03d" % round(th * 100)])
            assert mx == float(g["sync_max_%03d" % round(th * 100)])
        assert c_oracle.sync_cascade(bits) == ref_dsp.sync_cascade(bits)


def test_c_oracle_filter_design_matches_scipy():
    lib = c_oracle.load()
    b, a, s = np.zeros(5), np.zeros(5), np.zeros(24)
    assert lib.oracle_butter4(12500 / 120000, b.ctypes.data, a.ctypes.data) == 0
    bs, as_ = signal.butter(4, 12500 / 120000)
    assert np.allclose(b, bs, rtol=1e-12) and np.allclose(a, as_, rtol=1e-11)
    assert lib.oracle_cheby1_sos8(0.05, 0.08, s.ctypes.data) == 0
    assert np.allclose(s.reshape(4, 6), signal.cheby1(8, 0.05, 0.08, output="sos"), rtol=1e-11)


def test_c_oracle_spectrum_and_sync_edge_cases():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(3000) + 1j * rng.standard_normal(3000)) * 0.1
    assert np.abs(c_oracle.spectrum_db(x, 2048) - ref_dsp.spectrum_db(x, 2048)).max() < 1e-9
    bits = np.zeros(600, dtype=np.uint8)
    bits[20:42] = ref_dsp.TS1
    assert c_oracle.find_sync(bits, 0.85) == ([20], 1.0)
    assert c_oracle.find_sync(np.zeros(10, dtype=np.uint8)) == ([], 0.0)
    noise_bits = rng.integers(0, 2, size=4000).astype(np.uint8)
    assert c_oracle.sync_cascade(noise_bits) == ref_dsp.sync_cascade(noise_bits)
