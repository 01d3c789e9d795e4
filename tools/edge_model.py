"""Block-end corrections of the fused path: float64 model + table generator (design time only).

The reference chain  decimate(x, 10) -> [NCO] -> filtfilt(butter4)  (tetraear/signal/processor.py:245-264) is
shift-invariant except at the two ends of a block, where SciPy's sosfiltfilt / filtfilt use an odd extension
(27 / 15 samples), steady-state initial conditions (zi * first sample) and -- for the backward pass -- the
forward pass's last output held constant. The fused kernel K1 computes the shift-invariant part: the cascade
applied to the block extended by ZEROS on both sides. Everything below is the difference

        D[m] = reference(x)[m] - K1_ideal(x zero-extended)[m]          (m within 168 outputs of an end)

which is linear in x and is driven by very few quantities: IIR filter states at the block end. They are
obtained as dot products of the block's last (first) samples with fixed weight tables ("functionals"), then
a few hundred literal order-4 recursion steps at 240 kS/s give D. The weight tables come from SciPy at design
time and are written to tetraear_b200/csrc/edge_tables_generated.h; the runtime never imports this module.

Run:  python tools/edge_model.py [--write]
"""
from __future__ import annotations

import argparse
import os

import numpy as np
from scipy import signal

Q = 10
SOS = signal.cheby1(8, 0.05, 0.8 / Q, output="sos")
B_BUT, A_BUT = signal.butter(4, (25000 / 2) / (240000 / 2), btype="low")
ZI1 = signal.sosfilt_zi(SOS)              # [4][2]
ZI2 = signal.lfilter_zi(B_BUT, A_BUT)     # [4]
PAD1, PAD2 = 27, 15
EDGE = 168                                # outputs corrected at each block end (= K1_EDGE)

# Truncation lengths: with these the model reproduces reference - K1_ideal to 4e-10 of full scale (`check` below);
# the fused kernel's own FIR approximation of the interior is ~2e-6.
G_HALF = 1000                             # g1 kept for |u| <= G_HALF (pole radius 0.9821 per sample)
NC = 1000                                 # causal-state functional length
NAC = 1250                                # two-pass (forward + backward) state functional length
TD = 104                                  # support (in 240 kS/s samples) of the stage-1 ringing terms (0.8346 per sample)
NRING = Q * TD + Q                        # ringing tables
T2 = 136                                  # stage-2 settle length (pole radius 0.8837 per sample)
NPTS = 16                                 # pointwise z values needed for the odd extension of stage 2


def _sos_zir(state, n):
    """zero-input response of the biquad cascade from `state` [4][2] (real), n samples"""
    y, _ = signal.sosfilt(SOS, np.zeros(n), zi=np.asarray(state, dtype=float).reshape(4, 2))
    return y


def build_tables():
    t = {}
    # g1[u], |u| <= G_HALF: impulse response of the zero-phase Chebyshev (both passes) at the full rate
    n = 2 * 6000 + 1
    imp = np.zeros(n); imp[n // 2] = 1.0
    f = signal.sosfilt(SOS, imp)
    g = signal.sosfilt(SOS, f[::-1])[::-1]
    t["g1"] = g[n // 2 - G_HALF: n // 2 + G_HALF + 1].copy()
    # WC[k][d]: causal cascade state (flattened [section][2]) after the input ... , x[-d] = 1, 0 (d zeros)
    wc = np.zeros((8, NC))
    _, zf = signal.sosfilt(SOS, np.array([1.0]), zi=np.zeros((4, 2)))
    st = zf
    for d in range(NC):
        wc[:, d] = st.reshape(-1)
        _, st = signal.sosfilt(SOS, np.array([0.0]), zi=st)
    t["wc"] = wc
    # unit states
    ring = np.zeros((8, NRING + 4000))
    for k in range(8):
        e = np.zeros(8); e[k] = 1.0
        ring[k] = _sos_zir(e, NRING + 4000)
    t["ringc"] = ring[:, :NRING].T.copy()            # RINGC[p][k]: zero-input output p steps after state e_k
    # U[:, k]: state of the backward pass after it has run over the whole ringing of e_k (from far right down to
    # the ringing's first sample); RING[p][k]: the backward pass's output at offset p into that ringing
    U = np.zeros((8, 8)); RING = np.zeros((NRING, 8))
    for k in range(8):
        y, zf = signal.sosfilt(SOS, ring[k][::-1], zi=np.zeros((4, 2)))
        U[:, k] = zf.reshape(-1)
        RING[:, k] = y[::-1][:NRING]
    t["u"] = U; t["ring"] = RING
    # WAC[k][i]: state of the backward pass after it has processed samples ... down to position 0, for the input
    # impulse x[i] = 1 (forward pass from a zero state at position 0)
    wac = np.zeros((8, NAC))
    span = NAC + 4000
    for i in range(NAC):
        imp = np.zeros(span); imp[i] = 1.0
        yf = signal.sosfilt(SOS, imp)
        _, zf = signal.sosfilt(SOS, yf[::-1], zi=np.zeros((4, 2)))
        wac[:, i] = zf.reshape(-1)
    t["wac"] = wac
    # U2[:, k]: state of stage 2's backward pass after it has run over the whole zero-input ringing of the forward
    # Butterworth state e_k (lfilter zi layout), from far right down to the ringing's first sample
    U2 = np.zeros((4, 4))
    for k in range(4):
        e = np.zeros(4); e[k] = 1.0
        r2, _ = signal.lfilter(B_BUT, A_BUT, np.zeros(2000), zi=e)
        _, zf = signal.lfilter(B_BUT, A_BUT, r2[::-1], zi=np.zeros(4))
        U2[:, k] = zf
    t["u2"] = U2
    # modal form of lfilter's state recursion s' = A s + Bv x (transposed direct form II): A = V diag(p) V^-1, so that
    # s(L) = V sigma(L), sigma_r(L) = c_r sum_k p_r^k x[L-1-k], c = V^-1 Bv -- scalar recurrences that split into segments
    # W0[k][u + 9], u = -9 .. NU-1: for freq_offset = 0 the weights of s2_functional_right are real and fixed
    A, Bv = s2_weights(0.0, 0)
    n_u = G_HALF + Q * T2 + Q
    lo = -(G_HALF + Q); lo -= lo % Q
    Wf = np.zeros((n_u - lo, 4))
    for u in range(lo, n_u):
        prev = Wf[u - lo - Q] if u - Q >= lo else np.zeros(4)
        gv = t["g1"][u + G_HALF] if abs(u) <= G_HALF else 0.0
        Wf[u - lo] = A @ prev + Bv * gv
    t["w0"] = Wf[-9 - lo:].T.copy()
    p, V = np.linalg.eig(A)
    t["bp"] = p; t["bv"] = V; t["bc"] = np.linalg.solve(V, Bv.astype(complex))
    return t


# ---------------------------------------------------------------------------------------------
# reference and ideal-LTI models
# ---------------------------------------------------------------------------------------------
def reference_y(x, fo=0.0):
    """the reference's filtered 240 kS/s stream (processor.py:254-264), float64"""
    z = signal.decimate(np.asarray(x, dtype=np.complex128), Q)
    if fo != 0.0:
        z = z * np.exp(-1j * 2 * np.pi * fo * (np.arange(len(z)) / 240000.0))
    return signal.filtfilt(B_BUT, A_BUT, z)


def k1_ideal(x, fo=0.0, pad=40000):
    """what K1 approximates: the shift-invariant cascade on the zero-extended block (float64, exact IIRs)"""
    x = np.asarray(x, dtype=np.complex128)
    n = len(x); L = (n + Q - 1) // Q
    xe = np.concatenate([np.zeros(pad, complex), x, np.zeros(pad + 20, complex)])
    f = signal.sosfilt(SOS, xe)
    g = signal.sosfilt(SOS, f[::-1])[::-1]
    z = g[::Q]                                         # pad is a multiple of Q: z index j <-> m = j - pad/Q
    m = np.arange(len(z)) - pad // Q
    if fo != 0.0:
        z = z * np.exp(-1j * 2 * np.pi * fo * (m / 240000.0))
    y = signal.lfilter(B_BUT, A_BUT, z)
    y = signal.lfilter(B_BUT, A_BUT, y[::-1])[::-1]
    return y[pad // Q: pad // Q + L]


# ---------------------------------------------------------------------------------------------
# the staged correction (what the CUDA kernel k_edge_correct does), float64
# ---------------------------------------------------------------------------------------------
def _rot(fo, j):
    return np.exp(-1j * 2 * np.pi * fo * (np.asarray(j, dtype=float) / 240000.0))


def _lf(x, zi):
    return signal.lfilter(B_BUT, A_BUT, x, zi=zi)


def s2_weights(fo, n_w):
    """WB[u + 9][4]: weights of the causal Butterworth state (lfilter zi layout) after z'[.. L-1] w.r.t. the input
    sample at distance u = d - k0 behind position 10 (L-1), for the rotated stream z'[j] = z[j] e^{-j W j}:
    s~(L) = sum_d x[n-1-d] W(d - k0), s(L) = e^{-j W L} s~(L)... built by the recursion
    W(u + 10) = e^{jW} A W(u) + e^{jW} b g1[u + 10]  (10 independent chains)."""
    # state-space of lfilter's transposed direct form II: s' = A s + Bv x ; y = s[0] + b0 x
    a, b = A_BUT, B_BUT
    A = np.zeros((4, 4)); Bv = np.zeros(4)
    for k in range(4):
        A[k, 0] = -a[k + 1]
        if k < 3:
            A[k, k + 1] = 1.0
        Bv[k] = b[k + 1] - a[k + 1] * b[0]
    return A, Bv


def right_edge(x, n, fo, tab, s2_K1=None):
    """D[t] for outputs m = L-1-t, t = 0..EDGE-1"""
    k0 = (n - 1) % Q
    L = (n + Q - 1) // Q
    xr = x[::-1]                                       # xr[d] = x[n-1-d]
    s_c = tab["wc"] @ xr[:NC]                          # causal state after x[n-1]
    ext = 2 * x[n - 1] - x[n - 2 - np.arange(PAD1)]
    yf, _ = signal.sosfilt(SOS, ext, zi=s_c.reshape(4, 2))
    s_b = ZI1 * yf[-1]
    _, s_ac_ex = signal.sosfilt(SOS, yf[::-1], zi=s_b)
    ds = s_ac_ex.reshape(-1) - tab["u"] @ s_c
    nt = EDGE + T2
    tt = np.arange(nt)
    kk = k0 + Q * tt
    d1 = np.where(kk < NRING, 1.0, 0.0) * (tab["ringc"][np.minimum(kk, NRING - 1)] @ ds)     # d1[t] at z index L-1-t
    # pointwise z of the K1 stream
    g1 = tab["g1"]
    zpts = np.zeros(NPTS, complex)
    for t in range(NPTS):
        d = np.arange(0, min(n, G_HALF + k0 + Q * t + 1))
        u = d - k0 - Q * t
        ok = np.abs(u) <= G_HALF
        zpts[t] = np.sum(g1[u[ok] + G_HALF] * xr[d[ok]])
    zex = zpts + d1[:NPTS]                             # exact decimator output at L-1-t
    jz = L - 1 - np.arange(NPTS)
    zpe = zex * _rot(fo, jz)                           # after the NCO
    # stage 2
    if s2_K1 is None:
        s2_K1 = s2_functional_right(xr, n, fo, tab)
    j_d1 = L - 1 - tt[::-1]                            # ascending z indices L-nt .. L-1
    d1p = d1[::-1] * _rot(fo, j_d1)
    ydl, ds2 = _lf(d1p, np.zeros(4, complex))          # causal response to the stage-1 correction
    s2_ex = s2_K1 + ds2
    ext2 = 2 * zpe[0] - zpe[1 + np.arange(PAD2)]
    y2e, _ = _lf(ext2, s2_ex)
    F = y2e[-1]
    tc = np.arange(T2)
    pc = 9 - k0 + Q * tc
    zc = np.where(pc < NRING, 1.0, 0.0) * (tab["ring"][np.minimum(pc, NRING - 1)] @ s_c) * _rot(fo, L + tc)
    y2k, _ = _lf(zc, s2_K1)
    dyr = np.where(tc < PAD2, np.concatenate([y2e, np.zeros(T2 - PAD2)]), F) - y2k
    seq = np.concatenate([dyr[::-1], ydl[::-1]])       # from j = L+T2-1 down to L-nt
    out, _ = _lf(seq, ZI2 * F)
    return out[T2: T2 + EDGE]                          # D at m = L-1-t


def s2_functional_right(xr, n, fo, tab):
    """causal Butterworth state after the K1 stream's rotated decimator output z'[.. L-1] (lfilter zi layout)"""
    k0 = (n - 1) % Q
    L = (n + Q - 1) // Q
    A, Bv = s2_weights(fo, 0)
    w = np.exp(1j * 2 * np.pi * fo / 240000.0)
    g1 = tab["g1"]
    n_u = G_HALF + Q * T2 + Q                          # distances u = d - k0 in [-9, n_u) are used
    lo = -(G_HALF + Q)                                 # the chains start where g1 begins (samples AFTER a z position count too)
    lo -= lo % Q                                       # multiple of Q
    Wf = np.zeros((n_u - lo, 4), complex)              # Wf[u - lo]
    for u in range(lo, n_u):
        prev = Wf[u - lo - Q] if u - Q >= lo else np.zeros(4, complex)
        gv = g1[u + G_HALF] if abs(u) <= G_HALF else 0.0
        Wf[u - lo] = w * (A @ prev) + Bv * gv           # W(u) = Bv g1[u] + w A W(u - 10)
    W = Wf[-9 - lo:]                                    # W[u + 9]
    d = np.arange(0, min(n, n_u + k0))
    u = d - k0
    ok = (u >= -9) & (u < n_u)
    s_t = (W[u[ok] + 9] * xr[d[ok], None]).sum(axis=0)
    # s~ defined with s~(j) = e^{+jWj} s(j) and the input term at step j: z'[j] = e^{-jWj} z[j]
    return s_t * np.exp(-1j * 2 * np.pi * fo * ((L - 1) / 240000.0))


def s2_functional_right_modal(xr, n, fo, tab, seg_len=34):
    """the same state through the modal form, evaluated the way the kernel does: per chain (u mod 10) and segment of
    `seg_len` steps a local pass from zero, then the carries"""
    k0 = (n - 1) % Q
    L = (n + Q - 1) // Q
    lam = tab["bp"] * np.exp(1j * 2 * np.pi * fo / 240000.0)
    g1 = tab["g1"]
    n_u = G_HALF + Q * T2 + Q
    lo = -(G_HALF + Q); lo -= lo % Q
    n_step = (n_u - lo) // Q
    S = np.zeros(4, complex)
    for b in range(Q):
        C = np.zeros(4, complex)
        for s0 in range(0, n_step, seg_len):
            om = np.zeros(4, complex); sl = np.zeros(4, complex); ps = np.zeros(4, complex); pw = lam.copy()
            for j in range(s0, min(s0 + seg_len, n_step)):
                u = lo + b + Q * j
                g = g1[u + G_HALF] if abs(u) <= G_HALF else 0.0
                d = u + k0
                xv = xr[d] if 0 <= d < n else 0.0
                om = lam * om + g
                sl += xv * om
                ps += xv * pw
                pw = pw * lam
            S += sl + C * ps
            C = lam ** min(seg_len, n_step - s0) * C + om
    sigma = tab["bc"] * S * np.exp(-1j * 2 * np.pi * fo * ((L - 1) / 240000.0))
    return tab["bv"] @ sigma


def left_edge(x, n, fo, tab):
    """D[m], m = 0..EDGE-1"""
    s_ac0 = tab["wac"] @ x[:NAC]
    extl = 2 * x[0] - x[PAD1 - np.arange(PAD1)]        # positions -27 .. -1
    _, s_c_ex = signal.sosfilt(SOS, extl, zi=ZI1 * extl[0])
    dsc = s_c_ex.reshape(-1)
    nt = EDGE + T2
    mm = np.arange(nt)
    pp = Q * mm
    d1 = np.where(pp < NRING, 1.0, 0.0) * (tab["ring"][np.minimum(pp, NRING - 1)] @ dsc)     # at z index m
    g1 = tab["g1"]
    zpts = np.zeros(NPTS, complex)
    for t in range(NPTS):
        i = np.arange(0, min(n, Q * t + G_HALF + 1))
        u = Q * t - i
        ok = np.abs(u) <= G_HALF
        zpts[t] = np.sum(g1[u[ok] + G_HALF] * x[i[ok]])
    zpe = (zpts + d1[:NPTS]) * _rot(fo, np.arange(NPTS))
    ext2 = 2 * zpe[0] - zpe[PAD2 - np.arange(PAD2)]    # positions -15 .. -1
    _, s2_ex = _lf(ext2, ZI2 * ext2[0])
    # the K1 stream before the block: the backward pass's zero-input ringing, z[-t] = RINGC[10 t - 1] . s_ac0
    tk = np.arange(T2, 0, -1)                           # positions -T2 .. -1
    pk = Q * tk - 1
    zk = np.where(pk < NRING, 1.0, 0.0) * (tab["ringc"][np.minimum(pk, NRING - 1)] @ s_ac0) * _rot(fo, -tk)
    _, s2_k1 = _lf(zk, np.zeros(4, complex))
    # causal response to the state difference and the stage-1 correction over the EDGE outputs; what follows them is the
    # forward state's zero-input ringing, whose effect on the backward pass is the fixed map U2
    dy, s_end = _lf((d1 * _rot(fo, mm))[:EDGE], s2_ex - s2_k1)
    out, _ = _lf(dy[::-1], tab["u2"] @ s_end)
    return out[::-1]


def check(n=20003, fo=0.0, seed=1, tab=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x[:50] += 3.0; x[-40:] -= 2.0j                     # not zero-mean at the ends
    L = (n + Q - 1) // Q
    ref = reference_y(x, fo)
    lti = k1_ideal(x, fo)
    Dl = (ref - lti)[:EDGE]
    Dr = (ref - lti)[::-1][:EDGE]
    el = left_edge(x, n, fo, tab)
    er = right_edge(x, n, fo, tab)
    sc = np.abs(ref).max()
    mid = np.abs(ref - lti)[EDGE: L - EDGE].max() / sc
    return np.abs(el - Dl).max() / sc, np.abs(er - Dr).max() / sc, mid, np.abs(Dl).max() / sc, np.abs(Dr).max() / sc


def edge_corrections(x, fo=0.0, tab=None):
    """(D_left[m], D_right[t]) for m = 0..EDGE-1 and outputs L-1-t, t = 0..EDGE-1 (complex128)"""
    x = np.asarray(x, dtype=np.complex128)
    tab = tab or build_tables()
    return left_edge(x, len(x), fo, tab), right_edge(x, len(x), fo, tab)


def write_header(tab, path):
    def arr(fh, name, a):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        fh.write("static const double %s[%d] = {\n" % (name, len(a)))
        for i in range(0, len(a), 4):
            fh.write("    " + ", ".join("%.17e" % v for v in a[i:i + 4]) + ",\n")
        fh.write("};\n")
    with open(path, "w") as fh:
        fh.write("// GENERATED by tools/edge_model.py -- do not edit.\n")
        fh.write("// Weight tables of the block-end corrections of the fused path (see DESIGN.md, tools/edge_model.py).\n#pragma once\n")
        fh.write("#define ET_G %d\n#define ET_NC %d\n#define ET_NAC %d\n#define ET_TD %d\n#define ET_NRING %d\n#define ET_T2 %d\n#define ET_NPTS %d\n"
                 % (G_HALF, NC, NAC, TD, NRING, T2, NPTS))
        arr(fh, "ET_G1", tab["g1"])                 # [2 G + 1]      g1[u + G]
        arr(fh, "ET_WC", tab["wc"])                 # [8][NC]        causal cascade state per unit sample at distance d
        arr(fh, "ET_WAC", tab["wac"])               # [8][NAC]       backward-pass state at position 0 per unit sample at i
        arr(fh, "ET_RINGC", tab["ringc"])           # [NRING][8]     zero-input output p steps after unit state k
        arr(fh, "ET_RING", tab["ring"])             # [NRING][8]     backward-pass output at offset p into the ringing of unit state k
        arr(fh, "ET_U", tab["u"])                   # [8][8]         backward-pass state after the whole ringing of unit state k
        arr(fh, "ET_U2", tab["u2"])                 # [4][4]         the same for the Butterworth stage (lfilter zi layout)
        fh.write("#define ET_NW0 %d\n" % tab["w0"].shape[1])
        arr(fh, "ET_W0", tab["w0"])                 # [4][NW0]       freq_offset = 0: weights of the Butterworth state at the right end, W0[k][u + 9]
        cx = lambda a: np.stack([np.asarray(a).real, np.asarray(a).imag], axis=-1)
        arr(fh, "ET_BP", cx(tab["bp"]))             # [4][2]         poles of the Butterworth stage (re, im)
        arr(fh, "ET_BC", cx(tab["bc"]))             # [4][2]         its input vector in modal coordinates
        arr(fh, "ET_BV", cx(tab["bv"]))             # [4][4][2]      modal coordinates -> lfilter state


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--write", action="store_true"); a = ap.parse_args()
    tab = build_tables()
    if a.write:
        here = os.path.dirname(os.path.abspath(__file__))
        write_header(tab, os.path.join(here, "..", "tetraear_b200", "csrc", "edge_tables_generated.h"))
        print("wrote header")
    for n in (20003, 20000, 16384, 131072, 20007):
        for fo in (0.0, 1234.5, -12500.0):
            print(n, fo, ["%.2e" % v for v in check(n, fo, tab=tab)])
