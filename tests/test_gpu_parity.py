"""GPU: the CUDA path (through the C ABI) against the golden vectors of the reference and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, golden_input
from cases import CASES, SYNC_THRESHOLDS, TOLERANCE
from oracle import ref_dsp
from tetraear_b200 import sync, synth

pytestmark = pytest.mark.gpu
SOFT_TOL = 1e-5      # north star: soft metrics within 1e-5 relative (to max |symbol|)


def _check_case(sp, name, gen, n, fs, fo):
    g = load_golden(name)
    x = golden_input(g, **gen)
    sp.sample_rate = fs
    res = sp.process_batch(x[None, :], [fo], want_symbols=True, want_match=True)
    nd = int(res["n_dibits"][0])
    assert nd == len(g["dibits"])
    assert int(res["n_symbols"][0]) == len(g["symbols"])
    ns = len(g["symbols"])
    if name in TOLERANCE:
        # true-rate input (133.33 samples per symbol): the reference's stride-13 sampling drifts through the symbols, so
        # many of ITS decisions sit on a region border; the timing pick and the soft symbols must still agree, the
        # dibits agree wherever the reference's own decision had a margin above the soft tolerance
        assert int(res["best_phase"][0]) == int(g["best_phase"])
        s = res["symbols"][0, :ns].astype(np.complex128)
        scale = np.abs(g["symbols"]).max()
        assert np.abs(s - g["symbols"]).max() / scale <= SOFT_TOL
        d = g["symbols"][1:] * np.conj(g["symbols"][:-1])
        ph = np.angle(d)
        margin = np.min(np.abs(ph[:, None] - np.array([-5, -3, 3, 5]) * np.pi / 8), axis=1)
        safe = margin * np.abs(d) / scale ** 2 > 4 * SOFT_TOL
        agree = res["dibits"][0, :nd] == g["dibits"]
        print("%s: dibit agreement %.6f (%d of %d differ), %d decisions within the tolerance of a border"
              % (name, agree.mean(), (~agree).sum(), nd, (~safe).sum()))
        assert agree[safe].all()
        assert agree.mean() > 0.995
        return
    assert np.array_equal(res["dibits"][0, :nd], g["dibits"]), "dibits differ from the reference"
    if ns:
        assert int(res["best_phase"][0]) == int(g["best_phase"])
        s = res["symbols"][0, :ns].astype(np.complex128)
        err = np.abs(s - g["symbols"]).max() / np.abs(g["symbols"]).max()
        assert err <= SOFT_TOL, f"soft symbols off by {err:.3e}"
    if nd:
        for th in SYNC_THRESHOLDS:
            pos, mx = sync.find_sync(res["ts_match"][0], nd, th, return_max_corr=True)
            assert pos == list(g["sync_pos_%03d" % round(th * 100)])
            assert mx == float(g["sync_max_%03d" % round(th * 100)])
        bits = ref_dsp.symbols_to_bits(g["dibits"])
        assert np.array_equal(res["ts_match"][0, : 2 * nd - 21], ref_dsp.match_counts(bits))


@pytest.mark.parametrize("name,gen,n,fs,fo", CASES, ids=[c[0] for c in CASES])
def test_process_matches_reference(gpu_processor, name, gen, n, fs, fo):
    _check_case(gpu_processor, name, gen, n, fs, fo)


def test_one_symbol_block_keeps_its_symbol(gpu_processor):
    """A block that yields exactly one symbol: no dibit, but .symbols holds that symbol (processor.py:213-215, 268)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    g = load_golden("one_symbol_200")
    x = golden_input(g, seed=13, alphabet="centred", snr_db=30.0)
    assert len(g["dibits"]) == 0 and len(g["symbols"]) == 1
    d = sp.process(x.astype(np.complex128))
    assert len(d) == 0 and d.dtype == np.uint8
    assert sp.symbols.shape == (1,) and abs(sp.symbols[0] - g["symbols"][0]) <= SOFT_TOL * abs(g["symbols"][0])


def test_process_method_surface(gpu_processor):
    """process() returns what the reference returns and sets .symbols (processor.py:268)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    g = load_golden("gui_131072")
    x = golden_input(g, seed=3, alphabet="centred", snr_db=15.0)
    d = sp.process(x.astype(np.complex128))
    assert d.dtype == np.uint8 and np.array_equal(d, g["dibits"])
    assert sp.symbols.dtype == np.complex128 and len(sp.symbols) == len(d) + 1
    assert len(sp.process(np.array([], dtype=complex))) == 0 and len(sp.symbols) == 0


def test_batch_of_carriers_fast_path(gpu_processor):
    """Several independent carriers in one launch == each one alone through the oracle."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n = 1 << 17
    xs = np.stack([synth.carrier_iq(n, 100 + c, snr_db=15 + 5 * (c % 4), alphabet="centred" if c % 2 else "pi4")
                   for c in range(6)])
    res = sp.process_batch(xs, None, want_symbols=True)
    for c in range(6):
        r = ref_dsp.process(xs[c].astype(np.complex128), 0.0, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"]
        assert np.array_equal(res["dibits"][c, :nd], r["dibits"])
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL


def test_batch_with_mixed_freq_offsets(gpu_processor):
    """freq_offset per carrier (the GUI passes its AFC offset, ui/modern.py:2021-2022): zero and non-zero offsets in
    one batch, up to the edge of the fused path's range; noise-only input included."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n = 1 << 17
    fos = [0.0, 1234.5, -5000.0, 12500.0, -12499.0, 333.25]
    xs = [synth.carrier_iq(n, 200 + c, snr_db=20.0, alphabet="centred" if c % 2 else "pi4") for c in range(5)]
    rng = np.random.default_rng(77)
    xs.append((rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64))   # the reference tests' input
    xs = np.stack(xs)
    res = sp.process_batch(xs, fos, want_symbols=True)
    for c, fo in enumerate(fos):
        r = ref_dsp.process(xs[c].astype(np.complex128), fo, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"], (c, fo)
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL, (c, fo, err)
        if c < 5:
            assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), (c, fo)
        else:
            assert np.mean(res["dibits"][c, :nd] == r["dibits"]) > 0.999


@pytest.mark.parametrize("fo", [0.0, 9000.0])
def test_carriers_of_very_different_level_do_not_leak(gpu_processor, fo):
    """The persistent kernel streams its carriers back to back through one set of buffers (CTA b takes carriers b,
    b + 148, ...): a carrier 80-120 dB below the one before it must still come out as if it were alone (kept outputs
    never see another carrier's tiles)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n, n_car, n_sm = 16384 + 640, 156, 148
    base = [synth.carrier_iq(n, 400 + k, snr_db=25.0) for k in range(4)]
    level = np.where(np.arange(n_car) < n_sm, 1.0e4, 1.0)
    level[n_sm + 1::2] = 1.0e-2
    xs = np.stack([(base[c % 4] * level[c]).astype(np.complex64) for c in range(n_car)])
    res = sp.process_batch(xs, [fo] * n_car, want_symbols=True)
    for c in list(range(4)) + list(range(n_sm, n_car)):
        r = ref_dsp.process(xs[c].astype(np.complex128), fo, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"], c
        assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), c
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL, (c, err)


def test_freq_offset_outside_fused_range_uses_exact_path(gpu_processor):
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    x = synth.carrier_iq(1 << 15, 210, snr_db=20.0)
    r = ref_dsp.process(x.astype(np.complex128), 30000.0, 2.4e6)
    d = sp.process(x, 30000.0)
    assert np.array_equal(d, r["dibits"])
    assert np.abs(sp.symbols - r["symbols"]).max() / np.abs(r["symbols"]).max() <= SOFT_TOL


def test_unaligned_batches_and_small_blocks(gpu_processor):
    """Carriers whose rows are not 16-byte aligned (odd block length: the fused kernel copies tiles by hand instead of
    bulk copies) and the smallest block the fused path takes, several carriers per call, with and without freq_offset."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    for n, fos in ((100003, [0.0, 0.0, 0.0]), (50001, [0.0, 2500.0, -800.0]), (16384, [0.0] * 5), (16385, [1000.0] * 4)):
        xs = np.stack([synth.carrier_iq(n, 700 + c, snr_db=22.0, alphabet="centred" if c % 2 else "pi4") for c in range(len(fos))])
        res = sp.process_batch(xs, fos, want_symbols=True, want_sync=True)
        for c, fo in enumerate(fos):
            r = ref_dsp.process(xs[c].astype(np.complex128), fo, 2.4e6)
            nd = int(res["n_dibits"][c])
            assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"], (n, c)
            assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), (n, c)
            err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
            assert err <= SOFT_TOL, (n, c, err)
            want = ref_dsp.sync_cascade(ref_dsp.symbols_to_bits(r["dibits"]))
            assert [int(p) for p in res["sync_pos"][c, : res["n_sync"][c]]] == want, (n, c)


def test_full_size_batch_of_64_carriers_bit_exact(gpu_processor):
    """BASELINE size: 64 carriers x 2^20 samples (SNR 15..35 dB, both alphabets, as bench.py builds them) in one call,
    every one of them against the oracle: dibits identical, timing phase identical, soft symbols within 1e-5, sync
    positions of the device cascade identical (SURVEY 8d, config 4 subset)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n, n_car, n_base = 1 << 20, 64, 8
    base = [synth.carrier_iq(n, 500 + s, snr_db=40.0, alphabet="centred" if s % 2 else "pi4") for s in range(n_base)]
    rng = np.random.default_rng(64)
    snr = rng.uniform(15.0, 35.0, size=n_car)
    xs = np.empty((n_car, n), dtype=np.complex64)
    for c in range(n_car):
        sigma = np.sqrt(10.0 ** (-snr[c] / 10.0) / 2.0)
        xs[c] = base[c % n_base] + (sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    res = sp.process_batch(xs, None, want_symbols=True, want_sync=True)
    for c in range(n_car):
        r = ref_dsp.process(xs[c].astype(np.complex128), 0.0, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]), c
        assert int(res["best_phase"][c]) == r["best_phase"], c
        assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), c
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL, (c, err)
        want = ref_dsp.sync_cascade(ref_dsp.symbols_to_bits(r["dibits"]))
        assert [int(p) for p in res["sync_pos"][c, : res["n_sync"][c]]] == want, c


def test_full_size_properties(gpu_processor):
    """BASELINE size (2^20): structural checks that do not need the oracle."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n = 1 << 20
    x, inc = synth.dqpsk_baseband(n, 77, "centred", return_increments=True)
    x = x.astype(np.complex64)
    res = sp.process_batch(np.stack([x, x * np.complex64(0.25j)]), None, want_symbols=True)
    nd = int(res["n_dibits"][0])
    assert nd == (104858 - int(res["best_phase"][0])) // 13 - 1
    # scaling / rotating the input changes neither the timing pick nor any decision
    assert np.array_equal(res["dibits"][0], res["dibits"][1]) and res["best_phase"][0] == res["best_phase"][1]
    # the noiseless stream reproduces the transmitted increments (map 0,+pi/2,-pi/2,pi -> 0,1,2,3)
    d = res["dibits"][0, :nd]
    best = max(np.mean(d[200:6000] == inc[200 + lag: 6000 + lag]) for lag in range(0, 20))
    assert best == 1.0


@pytest.mark.parametrize("fs,sps,fo", [(2.048e6, 112, -1500.0), (1.8e6, 98, 0.0), (2.88e6, 156, 3000.0)])
def test_other_sample_rates_full_blocks(gpu_processor, fs, sps, fo):
    """The RTL-SDR rates the fused kernel does not cover (signal/capture.py:83-87; q = 8, 7, 12) at the full block size:
    the chunk-parallel exact kernel (k_exact_block, one CTA per carrier) against the oracle, three carriers per call."""
    sp = gpu_processor
    sp.sample_rate = fs
    n = 1 << 20
    x = np.stack([synth.carrier_iq(n, 700 + c, snr_db=28.0, alphabet="centred" if c & 1 else "pi4", sps=sps) for c in range(3)])
    res = sp.process_batch(x, [fo, 0.0, -fo], want_symbols=True, want_match=True)
    for c, f in enumerate((fo, 0.0, -fo)):
        r = ref_dsp.process(x[c].astype(np.complex128), f, fs)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"], c
        assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), c
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL, (c, err)
    sp.sample_rate = 2.4e6


@pytest.mark.parametrize("kind", ["c64", "c64_fo", "u8", "u8_fo"])
def test_large_batch_with_block_end_kernels_ahead(gpu_processor, kind):
    """From 2048 carriers per call the block-end kernels run ahead of the fused kernel on the same stream (tetra_b200.cu:
    edge_serial_mode) instead of beside it; 2100 short blocks of each input form, a dozen carriers against the oracle."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n, n_car = 16384 + 1280, 2100
    base = [synth.carrier_iq(n, 820 + k, snr_db=22.0 + k, alphabet="centred" if k & 1 else "pi4") for k in range(5)]
    fos = None
    if kind.endswith("_fo"):
        fos = np.random.default_rng(21).uniform(-9000.0, 9000.0, size=n_car)
        fos[::5] = 0.0
    if kind.startswith("u8"):
        b8 = []
        for x in base:
            z = x / np.abs(x).max() * 0.9
            b8.append(np.stack([np.clip(np.round((z.real + 1.0) * 127.5), 0, 255), np.clip(np.round((z.imag + 1.0) * 127.5), 0, 255)],
                               axis=-1).astype(np.uint8))
        raw = np.stack([b8[c % 5] for c in range(n_car)])
        res = sp.process_batch_u8(raw, fos, want_symbols=True)
        ref_in = [(b[:, 0].astype(np.float64) / 127.5 - 1.0) + 1j * (b[:, 1].astype(np.float64) / 127.5 - 1.0) for b in b8]
    else:
        x = np.stack([base[c % 5] for c in range(n_car)])
        res = sp.process_batch(x, fos, want_symbols=True)
        ref_in = [b.astype(np.complex128) for b in base]
    for c in (0, 1, 2, 3, 4, 147, 148, 1000, 2047, 2048, 2099):
        r = ref_dsp.process(ref_in[c % 5], float(fos[c]) if fos is not None else 0.0, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(r["dibits"]) and int(res["best_phase"][c]) == r["best_phase"], (kind, c)
        assert np.array_equal(res["dibits"][c, :nd], r["dibits"]), (kind, c)
        err = np.abs(res["symbols"][c, : nd + 1] - r["symbols"]).max() / np.abs(r["symbols"]).max()
        assert err <= SOFT_TOL, (kind, c, err)
