"""GPU: host batches that go through in chunks of carriers (tetra_set_h2d_chunk: the H2D copy of the chunks ahead beside the
kernels of the current one) and batches of more carriers than one launch takes -- same results as the one-piece call, and the
oracle's on the carriers checked."""
import numpy as np
import pytest

from oracle import ref_dsp
from tetraear_b200 import synth

pytestmark = pytest.mark.gpu
SOFT_TOL = 1e-5


def _batch(n_car, n, seed0=100):
    return np.stack([synth.carrier_iq(n, seed=seed0 + c, alphabet="pi4" if c % 2 else "centred", snr_db=25.0).astype(np.complex64)
                     for c in range(n_car)])


def _same(a, b):
    """equal results; of every row only what the call defines (n_dibits dibits, n_symbols symbols, 2 n_dibits - 21 windows)"""
    assert a.keys() == b.keys()
    for k in ("n_dibits", "n_symbols", "best_phase"):
        assert np.array_equal(a[k], b[k]), k
    if "n_sync" in a:
        assert np.array_equal(a["n_sync"], b["n_sync"]) and np.array_equal(a["sync_pos"], b["sync_pos"])
    for c in range(len(a["n_dibits"])):
        nd, ns = int(a["n_dibits"][c]), int(a["n_symbols"][c])
        assert np.array_equal(a["dibits"][c, :nd], b["dibits"][c, :nd]), ("dibits", c)
        if "symbols" in a:
            assert np.array_equal(a["symbols"][c, :ns], b["symbols"][c, :ns]), ("symbols", c)
        if "ts_match" in a:
            assert np.array_equal(a["ts_match"][c, : max(0, 2 * nd - 21)], b["ts_match"][c, : max(0, 2 * nd - 21)]), ("ts_match", c)


@pytest.mark.parametrize("with_fo", [False, True], ids=["fo0", "fo"])
def test_chunked_host_batch_equals_one_piece(gpu_processor, with_fo):
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n_car, n = 11, 65536                                   # 512 KiB per carrier: 1 MiB chunks hold two carriers, the last chunk one
    x = _batch(n_car, n)
    fo = np.linspace(-4000.0, 4000.0, n_car) if with_fo else None
    try:
        sp.set_h2d_chunk(-1)
        whole = sp.process_batch(x, fo, want_symbols=True, want_match=True, want_sync=True)
        sp.set_h2d_chunk(1 << 20)
        l0 = sp._lib.tetra_launch_count(sp._ctx)
        chunked = sp.process_batch(x, fo, want_symbols=True, want_match=True, want_sync=True)
        assert sp._lib.tetra_launch_count(sp._ctx) - l0 >= 6 * 3, "the batch did not go through in six chunks"
    finally:
        sp.set_h2d_chunk(0)
    _same(whole, chunked)
    for c in (0, 5, 10):
        ref = ref_dsp.process(x[c].astype(np.complex128), float(fo[c]) if with_fo else 0.0, 2.4e6)
        nd = int(chunked["n_dibits"][c])
        assert nd == len(ref["dibits"]) and np.array_equal(chunked["dibits"][c, :nd], ref["dibits"])
        assert int(chunked["best_phase"][c]) == int(ref["best_phase"])
        s = chunked["symbols"][c, : nd + 1].astype(np.complex128)
        assert np.abs(s - ref["symbols"]).max() / np.abs(ref["symbols"]).max() <= SOFT_TOL


def test_chunked_host_bytes_equal_one_piece(gpu_processor):
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n_car, n = 9, 65536
    x = _batch(n_car, n, seed0=300)
    scale = np.abs(x).max()
    raw = np.clip(np.round((np.stack([x.real, x.imag], axis=-1) / scale * 0.9 + 1.0) * 127.5), 0, 255).astype(np.uint8)
    fo = np.linspace(-3000.0, 3000.0, n_car)
    try:
        sp.set_h2d_chunk(-1)
        whole = sp.process_batch_u8(raw, fo, want_symbols=True, want_match=True)
        sp.set_h2d_chunk(1 << 18)     # two carriers of bytes per chunk
        chunked = sp.process_batch_u8(raw, fo, want_symbols=True, want_match=True)
    finally:
        sp.set_h2d_chunk(0)
    _same(whole, chunked)
    c = 4
    xs = (raw[c].astype(np.float64) / 127.5 - 1.0)
    ref = ref_dsp.process(xs[:, 0] + 1j * xs[:, 1], float(fo[c]), 2.4e6)     # what pyrtlsdr hands the reference
    nd = int(chunked["n_dibits"][c])
    assert nd == len(ref["dibits"]) and np.array_equal(chunked["dibits"][c, :nd], ref["dibits"])


def test_chunked_padded_rows_through_the_c_abi(gpu_processor):
    """Host rows with a pitch larger than the block (cudaMemcpy2DAsync per chunk)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n_car, n, pitch = 5, 40000, 40960
    x = np.zeros((n_car, pitch), dtype=np.complex64)
    x[:, :n] = _batch(n_car, n, seed0=500)
    cap = int(sp._lib.tetra_dibit_capacity(sp._ctx, n))
    out = {}
    try:
        for tag, chunk in (("whole", -1), ("chunked", 2 * pitch * 8)):
            sp.set_h2d_chunk(chunk)
            dib = np.zeros((n_car, cap), dtype=np.uint8)
            nd = np.zeros(n_car, dtype=np.int32)
            ph = np.zeros(n_car, dtype=np.int32)
            sp._check(sp._lib.tetra_process_batch(sp._ctx, x.ctypes.data, n_car, n, pitch, None, dib.ctypes.data, cap, nd.ctypes.data,
                                                  None, ph.ctypes.data, None, 0), "process_batch")
            out[tag] = (dib, nd, ph)
    finally:
        sp.set_h2d_chunk(0)
    for a, b in zip(out["whole"], out["chunked"]):
        assert np.array_equal(a, b)
    ref = ref_dsp.process(x[3, :n].astype(np.complex128), 0.0, 2.4e6)
    assert np.array_equal(out["chunked"][0][3, : out["chunked"][1][3]], ref["dibits"])


@pytest.mark.parametrize("chunk", [0, -1], ids=["copy-ahead", "carrier-cap"])
def test_more_carriers_than_one_launch_takes(gpu_processor, chunk):
    """40 000 short blocks in one call. chunk 0: the default 32 MiB chunks with the copy running ahead (ten chunks);
    chunk -1: host batches are not chunked for the copy's sake, so the call is cut only because the carrier index is a grid
    dimension of several kernels (two chunks of at most 32 768 carriers)."""
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n_car, n = 40000, 1000
    base = _batch(8, n, seed0=700)
    x = np.ascontiguousarray(base[np.arange(n_car) % 8])
    x[33000] = base[3] * np.complex64(0.5)                 # one carrier beyond the first chunk that differs from its neighbours
    try:
        sp.set_h2d_chunk(chunk)
        res = sp.process_batch(x, None, want_symbols=True, want_match=False)
    finally:
        sp.set_h2d_chunk(0)
    for c in (0, 7, 4193, 4194, 32767, 32768, 33000, 39999):
        ref = ref_dsp.process(x[c].astype(np.complex128), 0.0, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(ref["dibits"]) and np.array_equal(res["dibits"][c, :nd], ref["dibits"]), c
        s = res["symbols"][c, : nd + 1].astype(np.complex128)
        assert np.abs(s - ref["symbols"]).max() / np.abs(ref["symbols"]).max() <= SOFT_TOL
    # identical inputs, identical outputs -- across the chunk boundaries too
    assert np.array_equal(res["n_dibits"][:8], res["n_dibits"][32768:32776])
    for c in range(8):
        assert np.array_equal(res["dibits"][c, : res["n_dibits"][c]], res["dibits"][32768 + c, : res["n_dibits"][c]])


def test_launches_with_as_many_items_as_ctas(gpu_processor):
    """48 carriers x 2^20 samples of RTL-SDR bytes from page-locked memory: 32 MiB chunks of 16 carriers, each cut into nine
    segments -> 144 work items on 144 persistent CTAs while the block-end kernels of the side stream hold 16 SMs. The CTAs
    that start late find the item counter run out (visit r02t: they took items beyond the batch) and must simply leave."""
    import torch
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n_car, n = 48, 1 << 20
    base = np.stack([synth.carrier_iq(n, seed=900 + c, alphabet="pi4", snr_db=25.0) for c in range(2)])
    raw2 = np.clip(np.round((np.stack([base.real, base.imag], axis=-1) * 0.3 + 1.0) * 127.5), 0, 255).astype(np.uint8)
    pinned = torch.from_numpy(np.ascontiguousarray(raw2[np.arange(n_car) % 2])).pin_memory()
    refs = []
    for c in range(2):
        xs = raw2[c].astype(np.float64) / 127.5 - 1.0
        refs.append(ref_dsp.process(xs[:, 0] + 1j * xs[:, 1], 0.0, 2.4e6))
    for rep in range(3):
        res = sp.process_batch_u8(pinned.numpy(), None, want_symbols=True, want_match=False)
        for c in range(n_car):
            ref = refs[c % 2]
            nd = int(res["n_dibits"][c])
            assert nd == len(ref["dibits"]) and np.array_equal(res["dibits"][c, :nd], ref["dibits"]), (rep, c)
            assert int(res["best_phase"][c]) == int(ref["best_phase"])
        s = res["symbols"][n_car - 1, : nd + 1].astype(np.complex128)
        assert np.abs(s - refs[(n_car - 1) % 2]["symbols"]).max() / np.abs(refs[(n_car - 1) % 2]["symbols"]).max() <= SOFT_TOL
    # the same shape of launch from complex64: 16 carriers in one piece
    x16 = torch.from_numpy(np.ascontiguousarray(base[np.arange(16) % 2])).pin_memory()
    refs64 = [ref_dsp.process(base[c].astype(np.complex128), 0.0, 2.4e6) for c in range(2)]
    try:
        sp.set_h2d_chunk(-1)
        for rep in range(3):
            res = sp.process_batch(x16.numpy(), None, want_symbols=False, want_match=False)
            for c in range(16):
                r = refs64[c % 2]
                nd = int(res["n_dibits"][c])
                assert nd == len(r["dibits"]) and np.array_equal(res["dibits"][c, :nd], r["dibits"]), (rep, c)
    finally:
        sp.set_h2d_chunk(0)
