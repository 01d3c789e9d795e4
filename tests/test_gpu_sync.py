"""GPU: TetraDecoder.find_sync + decode()'s threshold cascade on the device against the oracle,
whose find_sync is pinned to the reference's own outputs by the golden fixtures (tests/test_oracle.py)."""
import numpy as np
import pytest

from conftest import load_golden, golden_input
from cases import CASES
from oracle import ref_dsp
from tetraear_b200 import sync

pytestmark = pytest.mark.gpu


def _bits_to_dibits(bits):
    bits = np.asarray(bits, dtype=np.uint8)
    assert len(bits) % 2 == 0
    return (bits[0::2] << 1 | bits[1::2]).astype(np.uint8)


def _planted(rng, n_bits, plants):
    """random bits with TS1/TS2 copies (pattern, bit offset, number of flipped bits) planted in"""
    bits = rng.integers(0, 2, size=n_bits).astype(np.int64)
    for pat, off, n_err in plants:
        p = (ref_dsp.TS1 if pat == 1 else ref_dsp.TS2).copy()
        if n_err:
            p[rng.choice(22, size=n_err, replace=False)] ^= 1
        bits[off:off + 22] = p
    return bits


def test_sync_positions_match_oracle_on_crafted_streams(gpu_processor):
    sp = gpu_processor
    rng = np.random.default_rng(5)
    cases = [
        [],                                                        # nothing planted: adaptive path on random bits
        [(1, 20, 0)],                                              # the reference unit test's plant
        [(1, 21, 0), (2, 600, 0), (1, 5001, 1)],                   # odd offsets, both patterns, 0.90 level
        [(2, 100, 3)],                                             # 19/22: found at 0.85
        [(1, 300, 4)],                                             # 18/22: found at 0.80
        [(1, 1000, 5), (2, 4000, 5)],                              # 17/22: only the adaptive pass
        [(1, 50, 0), (1, 200, 0), (2, 299, 0), (1, 300, 0)],       # closer than 250: skipped by the jump
        [(1, 16130 - 22, 0)],                                      # last window
        [(1, 10, 2), (2, 10 + 250, 2), (1, 10 + 499, 2), (2, 9000, 1)],
    ]
    n_bits = 16130
    streams = np.stack([_bits_to_dibits(_planted(rng, n_bits, pl)) for pl in cases])
    got = sp.sync_positions(streams)
    for c, pl in enumerate(cases):
        bits = ref_dsp.symbols_to_bits(streams[c])
        assert got[c] == ref_dsp.sync_cascade(bits), (c, pl)


def test_sync_positions_ragged_and_tiny(gpu_processor):
    sp = gpu_processor
    rng = np.random.default_rng(6)
    cap = 4000
    d = rng.integers(0, 4, size=(5, cap), dtype=np.uint8)
    nd = np.array([4000, 0, 5, 11, 1234], dtype=np.int32)           # 0, 10 and 22 bits: nothing / one window
    d[3, :11] = _bits_to_dibits(ref_dsp.TS2)
    got = sp.sync_positions(d, nd)
    for c in range(5):
        bits = ref_dsp.symbols_to_bits(d[c, :nd[c]])
        assert got[c] == ref_dsp.sync_cascade(bits), c
    assert got[3] == [0]


@pytest.mark.parametrize("name,gen,n,fs,fo", [c for c in CASES if c[2] >= 16384 and c[3] == 2.4e6], ids=lambda v: v if isinstance(v, str) else None)
def test_process_batch_sync_matches_reference_cascade(gpu_processor, name, gen, n, fs, fo):
    sp = gpu_processor
    sp.sample_rate = fs
    g = load_golden(name)
    x = golden_input(g, **gen)
    res = sp.process_batch(x[None, :], [fo], want_symbols=False, want_match=True, want_sync=True)
    nd = int(res["n_dibits"][0])
    bits = ref_dsp.symbols_to_bits(g["dibits"])
    want = ref_dsp.sync_cascade(bits)
    got = [int(p) for p in res["sync_pos"][0, : res["n_sync"][0]]]
    assert got == want
    # the device cascade and the host replay over the device's match counts agree
    assert got == sync.sync_cascade(res["ts_match"][0], nd)
    assert sync.burst_slices(got, nd) == [(p - 216) // 2 and ((p - 216) // 2, p - 216, (p - 216) // 510) for p in got
                                          if p - 216 >= 0 and (p - 216) // 2 + 255 <= nd]


def test_parse_bursts_matches_reference_golden(gpu_processor):
    """Every golden slot (burst type and CRC verdict from the reference's own parse_burst) planted in a stream at the
    position decode() would cut it from."""
    sp = gpu_processor
    g = load_golden("bursts")
    frames = g["frames"]
    n = len(frames)
    rng = np.random.default_rng(8)
    cap = 255 + 300
    dib = rng.integers(0, 4, size=(n, cap), dtype=np.uint8)
    nd = np.full(n, cap, dtype=np.int32)
    spos = np.zeros((n, 6), dtype=np.int32)
    ns = np.zeros(n, dtype=np.int32)
    for c in range(n):
        s0 = int(rng.integers(0, 300))
        dib[c, s0:s0 + 255] = frames[c]
        spos[c, 0] = 2 * s0 + 216 + (c % 2)            # odd positions floor to the same start symbol (decoder.py:869)
        spos[c, 1] = 100                                # start < 0: dropped
        spos[c, 2] = 2 * (cap - 200) + 216              # slot runs past the end: dropped
        ns[c] = 3
    got = sp.parse_bursts(dib, nd, spos, ns)
    for c in range(n):
        want = ref_dsp.decode_bursts(dib[c], [int(p) for p in spos[c, :3]])
        assert got[c] == want, c
        assert len(got[c]) == 1 and got[c][0][2:] == (int(g["burst_type"][c]), int(g["crc_ok"][c])), c


def test_sync_positions_random_streams_property(gpu_processor):
    """128 random dibit streams of ragged length with random plants (0-6 bit errors): device cascade == oracle."""
    sp = gpu_processor
    rng = np.random.default_rng(4242)
    n_car, cap = 128, 3000
    d = rng.integers(0, 4, size=(n_car, cap), dtype=np.uint8)
    nd = rng.integers(11, cap + 1, size=n_car).astype(np.int32)
    for c in range(n_car):
        bits = ref_dsp.symbols_to_bits(d[c, :nd[c]])
        for _ in range(int(rng.integers(0, 7))):
            off = int(rng.integers(0, len(bits) - 22 + 1))
            p = (ref_dsp.TS1 if rng.integers(0, 2) else ref_dsp.TS2).copy()
            n_err = int(rng.integers(0, 7))
            if n_err:
                p[rng.choice(22, size=n_err, replace=False)] ^= 1
            bits[off:off + 22] = p
        d[c, :nd[c]] = _bits_to_dibits(bits)
    got = sp.sync_positions(d, nd)
    for c in range(n_car):
        assert got[c] == ref_dsp.sync_cascade(ref_dsp.symbols_to_bits(d[c, :nd[c]])), c


def test_transport_packing_kernels_match_host_form(gpu_processor):
    """tetra_pack_dibits / tetra_unpack_dibits against the host-tensor form the gloo tests use (tetraear_b200/shard.py)."""
    import torch
    from tetraear_b200 import shard
    sp = gpu_processor
    rng = np.random.default_rng(11)
    world, n_local, cap = 3, 5, 48
    d = torch.from_numpy(rng.integers(0, 4, size=(world, n_local, cap), dtype=np.uint8)).cuda()
    block = n_local * cap // 4 + 4 * n_local
    wire = torch.zeros(world * block, dtype=torch.uint8, device="cuda")
    for r in range(world):                                    # what every rank contributes to the all-gather
        sp.pack_dibits_device(d[r].data_ptr(), n_local * cap, wire[r * block:].data_ptr())
    out = torch.zeros_like(d)
    sp.unpack_dibits_device(wire.data_ptr(), world, n_local * cap // 4, block, out.data_ptr(), n_local * cap)
    sp.synchronize()
    assert torch.equal(out, d)
    assert torch.equal(wire.view(world, block)[0, : n_local * cap // 4].cpu(), shard._pack_cpu(d[0].cpu()))


@pytest.mark.parametrize("world,n_local,cap", [(1, 3, 48), (2, 5, 8080), (4, 2, 64)])
def test_peer_memory_allgather_between_contexts_of_one_gpu(world, n_local, cap):
    """tetra_allgather_dibits with `world` contexts ("ranks") on one device, their receive buffers handed over as plain
    pointers (tetra_p2p_connect_ptrs): every rank ends with every rank's streams and lengths, over several steps (both
    halves of the double buffer, flags reused). The multi-process form (CUDA IPC) runs in bench.py --gpus N."""
    import torch
    from tetraear_b200.processor import SignalProcessor
    rng = np.random.default_rng(12)
    sps = [SignalProcessor(2.4e6) for _ in range(world)]
    try:
        n = n_local * cap
        block = (n // 4 + 4 * n_local + 15) & ~15
        for r, sp in enumerate(sps):
            assert len(sp.p2p_create(r, world, block)) == 64
        bufs = [sp.p2p_buffer() for sp in sps]
        for sp in sps:
            sp.p2p_connect_ptrs(bufs)
        outs = [torch.zeros((world, n_local, cap), dtype=torch.uint8, device="cuda") for _ in range(world)]
        out_ns = [torch.zeros((world, n_local), dtype=torch.int32, device="cuda") for _ in range(world)]
        for step in range(5):
            d = torch.from_numpy(rng.integers(0, 4, size=(world, n_local, cap), dtype=np.uint8)).cuda()
            nd = torch.from_numpy(rng.integers(0, cap, size=(world, n_local)).astype(np.int32)).cuda()
            torch.cuda.synchronize()
            # ranks of one process: every context has its own stream, so a rank's wait overlaps the others' pushes
            for r, sp in enumerate(sps):
                sp.allgather_dibits_device(d[r].data_ptr(), n, nd[r].data_ptr(), n_local, outs[r].data_ptr(), out_ns[r].data_ptr())
            for r, sp in enumerate(sps):
                sp.synchronize()
                assert sp.p2p_status() == 0
                assert torch.equal(outs[r], d) and torch.equal(out_ns[r], nd), (step, r)
    finally:
        for sp in sps:
            sp.close()


def test_peer_memory_allgather_argument_checks():
    """Misuse is an error code with a message, not a launch: no exchange yet, a block too small for the streams, bad sizes."""
    import torch
    from tetraear_b200 import _lib
    from tetraear_b200.processor import SignalProcessor
    sp = SignalProcessor(2.4e6)
    try:
        d = torch.zeros(4 * 64, dtype=torch.uint8, device="cuda")
        nd = torch.zeros(4, dtype=torch.int32, device="cuda")
        out = torch.zeros(4 * 64, dtype=torch.uint8, device="cuda")
        with pytest.raises(_lib.TetraError, match="tetra_p2p_create"):
            sp.allgather_dibits_device(d.data_ptr(), 256, nd.data_ptr(), 4, out.data_ptr())
        with pytest.raises(_lib.TetraError):
            sp.p2p_create(0, 9, 64)                               # world > 8
        with pytest.raises(_lib.TetraError):
            sp.p2p_create(0, 1, 60)                               # block not a multiple of 16
        sp.p2p_create(0, 1, 64)                                   # room for 64 bytes per rank: 256 dibits need 64 + 16
        with pytest.raises(_lib.TetraError, match="exceed the block"):
            sp.allgather_dibits_device(d.data_ptr(), 256, nd.data_ptr(), 4, out.data_ptr())
        sp.p2p_create(0, 1, 80)
        with pytest.raises(_lib.TetraError):
            sp.allgather_dibits_device(d.data_ptr(), 250, nd.data_ptr(), 4, out.data_ptr())     # n not a multiple of 16
        sp.allgather_dibits_device(d.data_ptr(), 256, nd.data_ptr(), 4, out.data_ptr())
        sp.synchronize()
        assert sp.p2p_status() == 0
    finally:
        sp.close()


def test_process_batch_allgather_fused_push(gpu_processor):
    """tetra_process_batch_allgather between two contexts ("ranks") of one GPU: the finalize kernel pushes every carrier's
    packed stream into both receive buffers; each rank ends with both ranks' dibit streams and lengths, equal to what the
    plain call produces, over several steps."""
    import torch
    from tetraear_b200 import synth
    from tetraear_b200.processor import SignalProcessor
    n, n_car, world = 20000, 3, 2
    sps_ = [SignalProcessor(2.4e6) for _ in range(world)]
    try:
        cap = (sps_[0].dibit_capacity(n) + 15) & ~15
        block = (n_car * cap // 4 + 4 * n_car + 15) & ~15
        for r, sp in enumerate(sps_):
            sp.p2p_create(r, world, block)
        bufs = [sp.p2p_buffer() for sp in sps_]
        for sp in sps_:
            sp.p2p_connect_ptrs(bufs)
        for step in range(3):
            xs, want_d, want_n = [], [], []
            for r in range(world):
                x = np.stack([synth.carrier_iq(n, 900 + 10 * step + 3 * r + c, snr_db=25.0, alphabet="centred" if c & 1 else "pi4")
                              for c in range(n_car)])
                ref = gpu_processor.process_batch(x, None, want_symbols=False)
                xs.append(torch.view_as_real(torch.from_numpy(x).cuda()).contiguous())
                want_d.append(ref["dibits"]); want_n.append(ref["n_dibits"])
            outs, out_ns, keep = [], [], []
            for r in range(world):
                keep.append((torch.zeros((n_car, cap), dtype=torch.uint8, device="cuda"), torch.zeros(n_car, dtype=torch.int32, device="cuda"),
                             torch.zeros((n_car, 2 * cap, 2), dtype=torch.uint8, device="cuda")))
                outs.append(torch.zeros((world, n_car, cap), dtype=torch.uint8, device="cuda"))
                out_ns.append(torch.zeros((world, n_car), dtype=torch.int32, device="cuda"))
            torch.cuda.synchronize()
            # both ranks are enqueued back to back (no host synchronisation in between: a rank's wait needs the other's push)
            for r, sp in enumerate(sps_):
                dib, nd, mt = keep[r]
                sp.process_batch_allgather_device(xs[r].data_ptr(), n_car, n, n, dib.data_ptr(), cap, nd.data_ptr(), 0, 0, mt.data_ptr(),
                                                  outs[r].data_ptr(), out_ns[r].data_ptr())
            for r, sp in enumerate(sps_):
                sp.synchronize()
                assert sp.p2p_status() == 0
                for src in range(world):
                    assert np.array_equal(out_ns[r][src].cpu().numpy(), want_n[src]), (step, r, src)
                    for c in range(n_car):
                        k = int(want_n[src][c])
                        got = outs[r][src, c].cpu().numpy()
                        assert np.array_equal(got[:k], want_d[src][c, :k]) and not got[k:].any(), (step, r, src, c)
    finally:
        for sp in sps_:
            sp.close()


def test_sync_positions_on_streams_longer_than_shared_memory(gpu_processor):
    """Streams of 40 000 dibits (a few seconds of signal in one call): the packed bits and the hit mask live in global
    scratch (k_sync_positions_long); same positions as the oracle's cascade."""
    sp = gpu_processor
    rng = np.random.default_rng(15)
    n_bits = 80000
    cases = [
        [],                                                        # nothing planted: adaptive path
        [(1, 20, 0), (2, 30001, 0), (1, 79978, 0)],                # first / odd / last window
        [(2, 45000, 3)],                                           # 19/22: found at 0.85 only
        [(1, 100, 5), (2, 70000, 5)],                              # 17/22: only the adaptive pass
        [(1, 12, 0), (1, 200, 0), (2, 262, 0), (1, 40000, 1), (2, 40249, 0), (2, 40250, 0)],   # inside / at the 250 jump
    ]
    streams = np.stack([_bits_to_dibits(_planted(rng, n_bits, pl)) for pl in cases])
    nd = np.array([40000, 40000, 39999, 40000, 25001], dtype=np.int32)   # ragged: the last stream ends before its late plants
    got = sp.sync_positions(streams, nd)
    for c, pl in enumerate(cases):
        bits = ref_dsp.symbols_to_bits(streams[c, : nd[c]])
        assert got[c] == ref_dsp.sync_cascade(bits), (c, pl)


def test_process_batch_sync_on_a_block_longer_than_shared_memory(gpu_processor):
    """2^21 samples -> 16 130 dibits, more than the fused front end keeps in shared memory: the positions come from the
    separate launch behind the finalize kernel and equal the oracle's cascade on the oracle's dibits."""
    from tetraear_b200 import synth
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    x = np.stack([synth.carrier_iq(1 << 21, seed=40 + c, alphabet="pi4", snr_db=12.0 if c else 30.0).astype(np.complex64) for c in range(2)])
    res = sp.process_batch(x, None, want_symbols=False, want_match=True, want_sync=True)
    for c in range(2):
        ref = ref_dsp.process(x[c].astype(np.complex128), 0.0, 2.4e6)
        nd = int(res["n_dibits"][c])
        assert nd == len(ref["dibits"]) > 12288 and np.array_equal(res["dibits"][c, :nd], ref["dibits"])
        want = ref_dsp.sync_cascade(ref_dsp.symbols_to_bits(ref["dibits"]))
        got = [int(p) for p in res["sync_pos"][c, : res["n_sync"][c]]]
        assert got == want
        assert got == sync.sync_cascade(res["ts_match"][c], nd)
