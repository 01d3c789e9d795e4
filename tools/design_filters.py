"""Designs the fixed FIR tables of the fast (fs = 2.4 MS/s) channelize+demod kernel.

The reference chain  decimate(x,10) [cheby1(8,0.05,0.08) zero-phase, keep every 10th]
-> butter(4, 12.5k/120k) zero-phase  (tetraear/signal/processor.py:245-264) has, away from the
block edges, the LTI response  C2(f) * B2(f)  with C2 = |H_cheby|^2 at 2.4 MS/s and
B2 = |H_butter|^2 at 240 kS/s.  B2 is < 6e-7 beyond 60 kHz, so only |f| < 60 kHz matters and the
response can be realised as a multirate FIR cascade (all real, symmetric taps):

    proto  : 2.4 MS/s -> 240 kS/s, short low-pass whose only job is to null the bands that alias
             onto |f| < 60 kHz (least squares, weighted by what survives the later stages)
    hb     : 240 kS/s -> 120 kS/s, same idea for the single alias band around 120 kHz
    fir120 : 120 kS/s, C2*B2 / (proto*hb)  -- equaliser and channel filter in one table
    interp : 120 kS/s -> 240 kS/s, even outputs are the fir120 samples themselves, odd outputs a
             half-sample fractional-delay FIR

freq_offset != 0 (processor.py:259-261 mixes BETWEEN the two filters): the w samples are rotated by the
NCO, so proto and the Chebyshev response act at f + f_off while hb, B2 and fir120 act at f. The chain
above then misses the factor  R(f) = [C2(f+f_off) / P(f+f_off)] / [C2(f) / P(f)]  -- smooth and within
about 1 % of one where B2 is not negligible -- which an 11-tap complex FIR at 120 kS/s supplies
(least squares weighted by the response the signal still meets). Its taps are smooth in f_off and are
stored as Chebyshev series on |f_off| <= 12.5 kHz.

Coefficients of the two IIRs come from SciPy (the reference's own dependency) at design time
only; the result is written to tetraear_b200/csrc/taps_generated.h and committed.
Run:  python tools/design_filters.py [--write]
"""
from __future__ import annotations
import argparse, os, sys
import numpy as np
from scipy import signal

FS = 2.4e6
FS1 = 240e3
FS2 = 120e3
SOS_CHEBY = signal.cheby1(8, 0.05, 0.8 / 10, output="sos")
B_BUT, A_BUT = signal.butter(4, (25000 / 2) / (FS1 / 2), btype="low")

PROTO_H = 20      # proto taps -20..20 (41)
HB_K = 6          # hb: true half-band, centre 0.5 + HB_K odd taps each side (length 4*HB_K-1)
HB_H = 2 * HB_K - 1
FIR_H = 63        # fir120 taps -63..63 at 120k (127; the kernel pads to 128 = 4 quarters of 32)
INT_K = 8         # interp: 2*INT_K taps at half-sample offsets
REQ_K = 5         # freq_offset equaliser: complex taps -5..5 at 120k
REQ_DEG = 12      # ... each a Chebyshev series in f_off / REQ_FOMAX
REQ_FOMAX = 12500.0


def C2(f):
    _, h = signal.sosfreqz(SOS_CHEBY, worN=2 * np.pi * np.asarray(f, dtype=float) / FS)
    return np.abs(h) ** 2


def B2(f):
    _, h = signal.freqz(B_BUT, A_BUT, worN=2 * np.pi * np.asarray(f, dtype=float) / FS1)
    return np.abs(h) ** 2


def cosmat(f, H, fs):
    k = np.arange(0, H + 1)
    c = 2 * np.cos(2 * np.pi * np.outer(f, k) / fs)
    c[:, 0] = 1
    return c


def full(p_half):
    return np.concatenate([p_half[:0:-1], p_half])


def ls_antialias(H, fs, shifts, weight_fn, fmax=60e3, ngrid=241, iters=6):
    """min sum_shifts || W(f) P(f+shift) ||^2  s.t. P(0)=1, W = weight/|P(f)| (iterated)."""
    f = np.linspace(-fmax, fmax, ngrid)
    c0 = cosmat(np.array([0.0]), H, fs)[0]
    Pin = np.ones_like(f)
    wt = weight_fn(f)
    for _ in range(iters):
        W = wt / np.maximum(np.abs(Pin), 1e-3)
        Q = np.zeros((H + 1, H + 1))
        for sh in shifts:
            A = cosmat(f + sh, H, fs) * W[:, None]
            Q += A.T @ A
        Q += 1e-18 * np.trace(Q) / len(Q) * np.eye(H + 1)
        sol = np.linalg.solve(Q, c0)
        p = sol / (c0 @ sol)
        Pin = cosmat(f, H, fs) @ p
    leak = np.zeros_like(f)
    for sh in shifts:
        leak = np.maximum(leak, np.abs(cosmat(f + sh, H, fs) @ p) * wt / np.abs(Pin))
    return p, leak.max()


def design_halfband(K, fmax=60e3, ngrid=241, iters=5):
    """True half-band (centre 0.5, even taps zero, so HB(60k) = 0.5 and the equaliser stays
    bounded); odd taps by least squares on the alias band f+120k weighted by what survives."""
    f = np.linspace(-fmax, fmax, ngrid)
    ks = 2 * np.arange(K) + 1
    wt = C2(f) * B2(f)
    Hin = np.ones_like(f)
    for _ in range(iters):
        W = wt / np.maximum(np.abs(Hin), 1e-3)
        A = 2 * np.cos(2 * np.pi * np.outer(f + FS2, ks) / FS1) * W[:, None]
        hk, *_ = np.linalg.lstsq(A, -0.5 * W, rcond=None)
        Hin = 0.5 + 2 * np.cos(2 * np.pi * np.outer(f, ks) / FS1) @ hk
    leak = np.abs(0.5 + 2 * np.cos(2 * np.pi * np.outer(f + FS2, ks) / FS1) @ hk) * wt / np.abs(Hin)
    half = np.zeros(2 * K)
    half[0] = 0.5
    half[ks] = hk
    return half, leak.max()


def design():
    out = {}
    # --- proto ---
    shifts = [s * m * FS1 for m in range(1, 6) for s in (1, -1) if not (m == 5 and s == -1)]
    p, leak = ls_antialias(PROTO_H, FS, shifts, lambda f: C2(f) * B2(f))
    out["proto"] = full(p); out["proto_leak"] = leak
    # --- hb (input already shaped by proto; alias band = f +- 120k at fs 240k) ---
    h, leak = design_halfband(HB_K)
    out["hb"] = full(h); out["hb_leak"] = leak
    # --- fir120: frequency sampling of C2*B2/(P*HB) on a dense grid, |f| <= 60k ---
    NG = 4096
    fg = np.fft.fftfreq(NG, 1 / FS2)
    P = cosmat(fg, PROTO_H, FS) @ p
    HB = cosmat(fg, HB_H, FS1) @ h
    T = C2(fg) * B2(fg) / (P * HB)
    g = np.real(np.fft.ifft(T))
    g = np.concatenate([g[-FIR_H:], g[:FIR_H + 1]])
    out["fir120"] = g
    out["fir120_trunc"] = np.abs(np.real(np.fft.ifft(T))[FIR_H + 1: NG - FIR_H]).max()
    # --- interp: half-sample fractional delay, weighted by B2 (what the signal still contains) ---
    f = np.linspace(0, 60e3, 481)
    k = np.arange(INT_K)
    A = 2 * np.cos(2 * np.pi * np.outer(f, k + 0.5) / FS2)
    W = np.maximum(B2(f), 1e-9)
    sol, *_ = np.linalg.lstsq(A * W[:, None], W, rcond=None)
    out["interp_half"] = sol                     # taps at +-(k+1/2)
    out["interp_err"] = np.abs((A @ sol - 1) * B2(f)).max()
    # --- req: equaliser taps of the freq_offset path as Chebyshev series in f_off ---
    fq = np.linspace(-60e3, 60e3, 1024, endpoint=False)
    kk = np.arange(-REQ_K, REQ_K + 1)
    E = np.exp(-2j * np.pi * np.outer(fq, kk) / FS2)
    Pq = lambda ff: cosmat(ff, PROTO_H, FS) @ p
    Gq = (np.exp(-2j * np.pi * np.outer(fq, np.arange(-FIR_H, FIR_H + 1)) / FS2) @ g).real
    HBq = cosmat(fq, HB_H, FS1) @ h

    def req_taps(fo):
        R = (C2(fq + fo) / Pq(fq + fo)) / (C2(fq) / Pq(fq))
        W = np.abs(Gq * HBq * Pq(fq + fo))
        sol, *_ = np.linalg.lstsq(E * W[:, None], R * W, rcond=None)
        return sol

    def req_err(fo, taps):
        return np.abs((E @ taps) * Gq * HBq * Pq(fq + fo) - B2(fq) * C2(fq + fo)).max()

    nn = 48
    nodes = REQ_FOMAX * np.cos(np.pi * (np.arange(nn) + 0.5) / nn)
    T = np.array([req_taps(fo) for fo in nodes])
    V = np.polynomial.chebyshev.chebvander(nodes / REQ_FOMAX, REQ_DEG)
    coef, *_ = np.linalg.lstsq(V, T, rcond=None)                    # [DEG+1][2K+1] complex
    out["req_cheb"] = coef
    test = np.linspace(-REQ_FOMAX, REQ_FOMAX, 51)
    out["req_err"] = max(req_err(fo, np.polynomial.chebyshev.chebvander(np.array([fo / REQ_FOMAX]), REQ_DEG)[0] @ coef) for fo in test)
    return out


def simulate(x, taps, dtype=np.float64):
    """Reference model of the fast chain on a zero-extended block. Returns y at 240k, len ceil(N/10)."""
    x = np.asarray(x).astype(np.complex128 if dtype == np.float64 else np.complex64)
    N = len(x)
    L = (N + 9) // 10
    cd = x.dtype
    pr = taps["proto"].astype(dtype); hb = taps["hb"].astype(dtype)
    g = taps["fir120"].astype(dtype); ih = taps["interp_half"].astype(dtype)
    # w[m] = sum_k p[k] x[10m - k]  for m in a range wide enough for the later stages
    M0 = 2 * (FIR_H + INT_K + HB_H) + 16            # pre/post roll in 240k samples (even)
    xe = np.concatenate([np.zeros(10 * M0 + PROTO_H, cd), x, np.zeros(10 * M0 + PROTO_H + 20, cd)])
    conv = signal.fftconvolve(xe, pr.astype(cd), mode="same") if dtype == np.float64 else np.convolve(xe, pr, mode="same").astype(cd)
    w = conv[PROTO_H::10][: L + 2 * M0]             # w index i <-> m = i - M0
    # u[j] = sum_k hb[k] w[2j - k], j index <-> 120k sample j - M0/2
    cu = np.convolve(w, hb, mode="same").astype(cd)
    u = cu[0::2]
    v = np.convolve(u, g, mode="same").astype(cd)
    # y[2n] = v[n]; y[2n+1] = sum_k ih[k] (v[n-k] + v[n+1+k])
    yo = np.zeros(len(v), cd)
    for k in range(INT_K):
        a = np.roll(v, k); b = np.roll(v, -(k + 1))
        yo += ih[k] * (a + b)
    y = np.empty(2 * len(v), cd); y[0::2] = v; y[1::2] = yo
    return y[M0: M0 + L]


def write_header(taps, path):
    def arr(name, a):
        body = ",\n    ".join(", ".join("%.9ef" % v for v in a[i:i + 4]) for i in range(0, len(a), 4))
        return "static const float %s[%d] = {\n    %s\n};\n" % (name, len(a), body)
    with open(path, "w") as fh:
        fh.write("// GENERATED by tools/design_filters.py -- do not edit.\n")
        fh.write("// FIR tables of the fast fs=2.4 MS/s channelize+demod chain (see DESIGN.md).\n#pragma once\n")
        fh.write("#define TB_PROTO_H %d\n#define TB_HB_H %d\n#define TB_FIR_H %d\n#define TB_INT_K %d\n"
                 % (PROTO_H, HB_H, FIR_H, INT_K))
        fh.write(arr("TB_PROTO_TAPS", taps["proto"]))
        fh.write(arr("TB_HB_TAPS", taps["hb"]))
        fh.write(arr("TB_FIR120_TAPS", taps["fir120"]))
        fh.write(arr("TB_INTERP_TAPS", taps["interp_half"]))
        # freq_offset equaliser: tap k (0..2K <-> -K..K) = sum_j T_j(f_off / FOMAX) (COEF[j][k][0] + i COEF[j][k][1])
        c = taps["req_cheb"]
        fh.write("#define TB_REQ_K %d\n#define TB_REQ_DEG %d\n#define TB_REQ_FOMAX %.1f\n" % (REQ_K, REQ_DEG, REQ_FOMAX))
        flat = np.stack([c.real, c.imag], axis=-1).reshape(-1)
        body = ",\n    ".join(", ".join("%.17e" % v for v in flat[i:i + 4]) for i in range(0, len(flat), 4))
        fh.write("static const double TB_REQ_CHEB[%d] = {\n    %s\n};\n" % (len(flat), body))


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--write", action="store_true"); a = ap.parse_args()
    t = design()
    for k in ("proto_leak", "hb_leak", "fir120_trunc", "interp_err", "req_err"):
        print(k, "%.3e" % t[k])
    print("lens", len(t["proto"]), len(t["hb"]), len(t["fir120"]), 2 * len(t["interp_half"]))
    here = os.path.dirname(os.path.abspath(__file__))
    if a.write:
        write_header(t, os.path.join(here, "..", "tetraear_b200", "csrc", "taps_generated.h"))
        print("wrote header")
