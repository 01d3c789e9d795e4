"""CPU: host logic and the C ABI surface (no device work)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import signal

from conftest import ROOT, load_golden
from cases import CASES, SYNC_THRESHOLDS
from oracle import ref_dsp
from tetraear_b200 import _lib, sync


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tetra_b200.h")).read()
    declared = set(re.findall(r"\b(tetra_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("tetra_ctx")
    lib = _lib.load()
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _lib.load()
    ctx = _lib.c_ctx_p()
    rc = lib.tetra_create(C.byref(ctx), 0, 2.4e6)
    assert rc < 0 and b"no CUDA device" in lib.tetra_last_error(None)
    from tetraear_b200.processor import SignalProcessor
    with pytest.raises(_lib.TetraError):
        SignalProcessor(2.4e6)


@pytest.mark.parametrize("wn", [12500 / 120000, 0.0104166, 0.02083, 0.01, 0.5, 0.99])
def test_butter_design_matches_scipy(wn):
    lib = _lib.load()
    b, a = np.zeros(5), np.zeros(5)
    assert lib.tetra_design_butter4(wn, b.ctypes.data, a.ctypes.data) == 0
    bs, as_ = signal.butter(4, wn)
    assert np.allclose(b, bs, rtol=1e-12, atol=0) and np.allclose(a, as_, rtol=1e-11, atol=0)


@pytest.mark.parametrize("q", [2, 4, 7, 8, 10, 13])
def test_cheby_design_matches_scipy(q):
    lib = _lib.load()
    s = np.zeros(24)
    assert lib.tetra_design_cheby1_sos8(0.05, 0.8 / q, s.ctypes.data) == 0
    ref = signal.cheby1(8, 0.05, 0.8 / q, output="sos")
    assert np.allclose(s.reshape(4, 6), ref, rtol=1e-11, atol=0)


@pytest.mark.parametrize("name", [c[0] for c in CASES if c[2] >= 300])
def test_find_sync_replay_matches_reference(name):
    """tetra_find_sync (host replay) on oracle match counts == the reference's find_sync output."""
    g = load_golden(name)
    dib = g["dibits"]
    bits = ref_dsp.symbols_to_bits(dib)
    assert np.array_equal(sync.symbols_to_bits(dib), bits)
    mc = ref_dsp.match_counts(bits).astype(np.uint8)
    for th in SYNC_THRESHOLDS:
        pos, mx = sync.find_sync(mc, len(dib), th, return_max_corr=True)
        assert pos == list(g["sync_pos_%03d" % round(th * 100)])
        assert mx == float(g["sync_max_%03d" % round(th * 100)])
    assert sync.sync_cascade(mc, len(dib)) == ref_dsp.sync_cascade(bits)


def test_find_sync_edge_cases():
    assert sync.find_sync(np.zeros((0, 2), np.uint8), 0) == []
    assert sync.find_sync(np.zeros((0, 2), np.uint8), 5, 0.9, True) == ([], 0.0)
    bits = np.zeros(600, dtype=np.int64)
    bits[20:42] = ref_dsp.TS1
    mc = ref_dsp.match_counts(bits).astype(np.uint8)
    assert sync.find_sync(mc, 300, 0.85, True) == ([20], 1.0)
    assert sync.burst_slices([20, 300], 300) == [(42, 84, 0)]


def test_fir_tables_reproduce_reference_interior():
    """The generated FIR cascade (tools/design_filters.py) equals decimate+filtfilt away from the edges."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import design_filters as df
    from tetraear_b200 import synth
    taps = df.design()
    hdr = open(os.path.join(ROOT, "tetraear_b200", "csrc", "taps_generated.h")).read()
    first = float(re.search(r"TB_PROTO_TAPS\[\d+\] = \{\s*([-0-9.e+]+)f", hdr).group(1))
    assert abs(first - taps["proto"][0]) < 1e-9 * abs(first) + 1e-12, "taps_generated.h is stale"
    x = synth.carrier_iq(1 << 17, 21, snr_db=20.0)
    ref = ref_dsp.channel_filter(signal.decimate(x.astype(np.complex128), 10), 25000, 240000.0)
    y = df.simulate(x, taps, np.float32)
    err = np.abs(y - ref) / np.abs(ref).max()
    assert err[160:-160].max() < 2e-6


def test_freq_offset_equaliser_table_and_model():
    """The Chebyshev-series equaliser of the fused freq_offset path (tools/design_filters.py -> TB_REQ_CHEB): the committed
    header matches the generator, and a numpy model of the chain (NCO on the proto output, half-band, 11-tap equaliser,
    fir120, interpolation) equals decimate -> frequency_shift -> filtfilt away from the block ends."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import design_filters as df
    taps = df.design()
    assert taps["req_err"] < 5e-7
    hdr = open(os.path.join(ROOT, "tetraear_b200", "csrc", "taps_generated.h")).read()
    body = re.search(r"TB_REQ_CHEB\[(\d+)\] = \{(.*?)\};", hdr, re.S)
    vals = np.array([float(v) for v in body.group(2).replace("\n", " ").split(",")])
    coef = taps["req_cheb"]
    want = np.stack([coef.real, coef.imag], axis=-1).reshape(-1)
    assert int(body.group(1)) == want.size == (df.REQ_DEG + 1) * (2 * df.REQ_K + 1) * 2
    assert np.abs(vals - want).max() < 1e-12, "taps_generated.h is stale"
    rng = np.random.default_rng(5)
    n = 1 << 16
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)      # white: every frequency is exercised
    for fo in (12500.0, -7777.0, 1234.5):
        r = np.polynomial.chebyshev.chebvander(np.array([fo / df.REQ_FOMAX]), df.REQ_DEG)[0] @ coef
        L = (n + 9) // 10
        m0 = 2 * (df.FIR_H + df.INT_K + df.HB_H + df.REQ_K) + 16
        xe = np.concatenate([np.zeros(10 * m0 + df.PROTO_H, complex), x, np.zeros(10 * m0 + df.PROTO_H + 20, complex)])
        w = signal.fftconvolve(xe, taps["proto"].astype(complex), mode="same")[df.PROTO_H::10][: L + 2 * m0]
        w = w * np.exp(-2j * np.pi * fo * (np.arange(len(w)) - m0) / df.FS1)
        u = np.convolve(np.convolve(w, taps["hb"], mode="same")[0::2], r, mode="same")
        v = np.convolve(u, taps["fir120"], mode="same")
        odd = np.zeros(len(v), complex)
        for k in range(df.INT_K):
            odd += taps["interp_half"][k] * (np.roll(v, k) + np.roll(v, -(k + 1)))
        y = np.empty(2 * len(v), complex)
        y[0::2], y[1::2] = v, odd
        y = y[m0: m0 + L]
        dec = signal.decimate(x, 10)
        ref = ref_dsp.channel_filter(dec * np.exp(-2j * np.pi * fo * np.arange(len(dec)) / df.FS1), 25000, 240000.0)
        err = np.abs(y - ref) / np.abs(ref).max()
        assert err[200:-200].max() < 2.5e-6, (fo, err[200:-200].max())


def test_find_sync_replay_on_crafted_streams():
    """Host replay (tetra_find_sync / tetra_sync_cascade) over oracle match counts on streams that reach every level of
    decode()'s cascade: planted TS1/TS2 with 0..5 bit errors, hits closer than the 250-bit jump, noise only."""
    rng = np.random.default_rng(11)
    plants = [[], [(1, 20, 0)], [(2, 100, 3)], [(1, 300, 4)], [(1, 1000, 5), (2, 4000, 5)],
              [(1, 50, 0), (1, 200, 0), (2, 299, 0), (1, 300, 0)], [(1, 6000 - 22, 1)]]
    for pl in plants:
        bits = rng.integers(0, 2, size=6000).astype(np.int64)
        for pat, off, n_err in pl:
            p = (ref_dsp.TS1 if pat == 1 else ref_dsp.TS2).copy()
            if n_err:
                p[rng.choice(22, size=n_err, replace=False)] ^= 1
            bits[off:off + 22] = p
        mc = ref_dsp.match_counts(bits).astype(np.uint8)
        assert sync.sync_cascade(mc, len(bits) // 2) == ref_dsp.sync_cascade(bits), pl
        for th in (0.9, 0.85, 0.8, 0.76):
            assert sync.find_sync(mc, len(bits) // 2, th, True) == ref_dsp.find_sync(bits, th), (pl, th)


def test_sync_replay_and_c_oracle_agree_on_random_streams():
    """Property check over 60 random bit streams (random length, 0-6 planted training sequences with 0-6 bit errors):
    host replay of device-style match counts == numpy oracle == C oracle, for the cascade and for single thresholds."""
    from oracle import c_oracle
    rng = np.random.default_rng(2025)
    for trial in range(60):
        n_bits = int(rng.integers(22, 5000))
        bits = rng.integers(0, 2, size=n_bits).astype(np.int64)
        for _ in range(int(rng.integers(0, 7))):
            if n_bits <= 22:
                break
            off = int(rng.integers(0, n_bits - 22 + 1))
            p = (ref_dsp.TS1 if rng.integers(0, 2) else ref_dsp.TS2).copy()
            n_err = int(rng.integers(0, 7))
            if n_err:
                p[rng.choice(22, size=n_err, replace=False)] ^= 1
            bits[off:off + 22] = p
        want = ref_dsp.sync_cascade(bits)
        assert c_oracle.sync_cascade(bits.astype(np.uint8)) == want, trial
        if n_bits % 2 == 0:
            mc = ref_dsp.match_counts(bits).astype(np.uint8)
            assert sync.sync_cascade(mc, n_bits // 2) == want, trial
            th = float(rng.choice([0.9, 0.85, 0.8, 0.77]))
            assert sync.find_sync(mc, n_bits // 2, th, True) == ref_dsp.find_sync(bits, th), (trial, th)
            assert c_oracle.find_sync(bits.astype(np.uint8), th) == ref_dsp.find_sync(bits, th), (trial, th)


def test_entry_points_reject_a_null_context_without_a_device():
    """The C ABI never dereferences a NULL context: every compute entry point returns TETRA_E_INVALID (no GPU needed)."""
    import ctypes as C
    lib = _lib.load()
    buf = (C.c_uint8 * 64)()
    st = C.c_int32(0)
    rows = C.c_int64(0)
    assert lib.tetra_p2p_create(None, 0, 1, 64, C.addressof(buf)) < 0
    assert lib.tetra_p2p_connect(None, C.addressof(buf)) < 0
    assert lib.tetra_allgather_dibits(None, None, 16, None, 1, None, None) < 0
    assert lib.tetra_p2p_status(None, C.byref(st)) < 0
    assert lib.tetra_p2p_destroy(None) < 0
    assert lib.tetra_survey_wideband(None, None, 0, None, 0, 2048, None) < 0
    assert lib.tetra_stft_db_f64(None, None, 0, 2048, 2048, None, C.byref(rows)) < 0
    assert lib.tetra_p2p_buffer(None) is None
