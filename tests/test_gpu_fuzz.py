"""GPU: seeded sweep over batch shapes -- carriers x block length x input form x freq_offset x host memory kind -- against the
oracle. The fused kernel cuts carriers into segments so that items ~ CTAs for small batches; the block-end kernels run beside
it below 2048 carriers; host batches go through in chunks: the combinations are where the bookkeeping can go wrong (visit
r02t found one: late CTAs of a 144-item launch), so they are swept rather than hand-picked."""
import numpy as np
import pytest

from oracle import ref_dsp
from tetraear_b200 import synth

pytestmark = pytest.mark.gpu
SOFT_TOL = 1e-5
N_MAX = 1 << 20
LENGTHS = [16384, 16390, 20001, 32768, 65536, 70007, 100003, 131072, 150000, 262144, 300001, 409600, 524288, 1000000, 1 << 20]
OFFSETS = [0.0, 1234.5, -7000.0]


@pytest.fixture(scope="module")
def bases():
    x = np.stack([synth.carrier_iq(N_MAX, seed=2000 + c, alphabet="pi4" if c else "centred", snr_db=22.0 + 4 * c).astype(np.complex64)
                  for c in range(3)])
    raw = np.clip(np.round((np.stack([x.real, x.imag], axis=-1) * 0.3 + 1.0) * 127.5), 0, 255).astype(np.uint8)
    return x, raw


def _trial(sp, bases, rng, refs):
    import torch
    x, raw = bases
    n_car = int(rng.choice([1, 2, 3, 5, 7, 8, 9, 12, 13, 16, 17, 24, 31, 37]))
    n = int(rng.choice(LENGTHS))
    u8 = bool(rng.integers(0, 2))
    with_fo = bool(rng.integers(0, 2))
    pinned = bool(rng.integers(0, 2))
    chunk = int(rng.choice([0, 0, 1 << 20, 4 << 20, -1]))
    which = rng.integers(0, 3, size=n_car)
    fo = np.array([OFFSETS[int(k)] for k in rng.integers(0, 3, size=n_car)]) if with_fo else None
    src = raw if u8 else x
    batch = np.ascontiguousarray(src[which, :n])
    keep = None
    if pinned:
        keep = torch.from_numpy(batch).pin_memory()
        batch = keep.numpy()
    tag = dict(n_car=n_car, n=n, u8=u8, with_fo=with_fo, pinned=pinned, chunk=chunk)
    try:
        sp.set_h2d_chunk(chunk)
        call = sp.process_batch_u8 if u8 else sp.process_batch
        res = call(batch, fo, want_symbols=True, want_match=True, want_sync=bool(rng.integers(0, 2)))
    finally:
        sp.set_h2d_chunk(0)
    for c in range(n_car):
        f = float(fo[c]) if with_fo else 0.0
        key = (int(which[c]), n, u8, f)
        if key not in refs:
            if u8:
                xs = raw[which[c], :n].astype(np.float64) / 127.5 - 1.0
                xin = xs[:, 0] + 1j * xs[:, 1]
            else:
                xin = x[which[c], :n].astype(np.complex128)
            r = ref_dsp.process(xin, f, 2.4e6)
            bits = ref_dsp.symbols_to_bits(r["dibits"])
            refs[key] = (r["dibits"], r["symbols"], int(r["best_phase"]), ref_dsp.match_counts(bits), ref_dsp.sync_cascade(bits))
        dib, sym, best, mc, spos = refs[key]
        nd = int(res["n_dibits"][c])
        assert nd == len(dib) and np.array_equal(res["dibits"][c, :nd], dib), (tag, c)
        assert int(res["best_phase"][c]) == best, (tag, c)
        s = res["symbols"][c, : nd + 1].astype(np.complex128)
        assert np.abs(s - sym).max() / np.abs(sym).max() <= SOFT_TOL, (tag, c)
        assert np.array_equal(res["ts_match"][c, : 2 * nd - 21], mc), (tag, c)
        if "sync_pos" in res:
            assert [int(p) for p in res["sync_pos"][c, : res["n_sync"][c]]] == spos, (tag, c)
    return tag


@pytest.mark.parametrize("seed", range(10))
def test_batch_shapes_against_the_oracle(gpu_processor, bases, seed):
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    rng = np.random.default_rng(7700 + seed)
    refs = {}
    for _ in range(6):
        _trial(sp, bases, rng, refs)


# the RTL-SDR rates of signal/capture.py:83-87 (and two low rates): sample_rate, samples per symbol of the generator so that
# the decimated stream has a whole number of samples per symbol (the reference samples every int(rate / 18000)-th)
RATES = [(1.8e6, 98), (1.92e6, 104), (2.048e6, 112), (2.56e6, 140), (2.88e6, 156), (3.2e6, 169), (1.0e6, 52), (240e3, 13)]
RATE_LENGTHS = [4097, 5000, 12345, 16384, 40000, 65536, 100003, 131072, 200000, 262144, 300001]


@pytest.mark.parametrize("seed", range(5))
def test_other_sample_rates_against_the_oracle(gpu_processor, seed):
    """The exact path (k_exact_block / k_exact_chain) over the valid RTL-SDR rates, block lengths, batch sizes, freq_offsets."""
    sp = gpu_processor
    rng = np.random.default_rng(8800 + seed)
    try:
        for _ in range(4):
            fs, sps = RATES[int(rng.integers(0, len(RATES)))]
            n = int(rng.choice(RATE_LENGTHS))
            n_car = int(rng.choice([1, 2, 3, 5]))
            with_fo = bool(rng.integers(0, 2))
            x = np.stack([synth.carrier_iq(n, seed=3000 + 10 * seed + c, alphabet="centred" if c % 2 else "pi4", snr_db=30.0, sps=sps)
                          for c in range(n_car)]).astype(np.complex64)
            fo = np.array([777.7 * (1 + c) for c in range(n_car)]) if with_fo else None
            sp.sample_rate = fs
            res = sp.process_batch(x, fo, want_symbols=True, want_match=True)
            tag = dict(fs=fs, n=n, n_car=n_car, with_fo=with_fo)
            for c in range(n_car):
                r = ref_dsp.process(x[c].astype(np.complex128), float(fo[c]) if with_fo else 0.0, fs)
                nd = int(res["n_dibits"][c])
                assert nd == len(r["dibits"]) and np.array_equal(res["dibits"][c, :nd], r["dibits"]), (tag, c)
                assert int(res["best_phase"][c]) == int(r["best_phase"]), (tag, c)
                if len(r["symbols"]):
                    s = res["symbols"][c, : len(r["symbols"])].astype(np.complex128)
                    assert np.abs(s - r["symbols"]).max() / np.abs(r["symbols"]).max() <= SOFT_TOL, (tag, c)
                if nd >= 11:
                    assert np.array_equal(res["ts_match"][c, : 2 * nd - 21], ref_dsp.match_counts(ref_dsp.symbols_to_bits(r["dibits"]))), (tag, c)
    finally:
        sp.sample_rate = 2.4e6
