/*
 * tetra_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of the reference's IQ -> dibit path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this; the product
 * (tetraear_b200/) never does. It restates, from their published algorithms, the SciPy routines the
 * reference calls (the arithmetic of this path lives in SciPy, an un-vendored dependency:
 * requirements.txt:2 `scipy>=1.10.0`, this image 1.18.1) and the reference's own arithmetic:
 *
 *   scipy.signal.cheby1(8, 0.05, wn, output='sos') / butter(4, wn)   analog prototype, pre-warp, bilinear
 *   scipy.signal.sosfilt_zi / lfilter_zi                              steady-state initial conditions
 *   scipy.signal.sosfiltfilt / filtfilt                               odd extension 27 / 15, forward-backward
 *   scipy.signal.decimate(x, q)            (tetraear/signal/processor.py:254)   cheby1(8, 0.05, 0.8/q) + [::q]
 *   SignalProcessor.frequency_shift        (processor.py:85-100)
 *   SignalProcessor.filter_signal          (processor.py:51-83)
 *   SignalProcessor.extract_symbols        (processor.py:168-219)
 *   SignalProcessor.demodulate_dqpsk       (processor.py:102-166)
 *   SignalProcessor.process                (processor.py:221-273)
 *   TetraDecoder.symbols_to_bits / find_sync / decode() cascade   (core/decoder.py:140-169, 171-295, 845-856)
 *   spectrum block of CaptureThread.run    (ui/modern.py:1921-1934)
 *
 * Pinned by tests/test_oracle_c.py against tests/golden (npz files), which hold outputs of the reference
 * itself (oracle/make_golden.py).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

/* ---------------------------------------------------------------- filter design ---------- */
static void poly_from_roots(const cplx* r, int n, double* out /* n+1 */) {
    cplx c[16];
    c[0] = 1.0;
    for (int k = 0; k < n; ++k) {
        c[k + 1] = 0.0;
        for (int i = k + 1; i >= 1; --i) c[i] = c[i] - c[i - 1] * r[k];
    }
    for (int i = 0; i <= n; ++i) out[i] = creal(c[i]);
}

/* butter(4, wn) -> b[5], a[5] */
int oracle_butter4(double wn, double* b, double* a) {
    if (!(wn > 0.0 && wn < 1.0)) return -1;
    const double fs = 2.0, warped = 2.0 * fs * tan(M_PI * wn / fs);
    cplx pz[4], den = 1.0;
    for (int i = 0; i < 4; ++i) {
        const int m = -4 + 1 + 2 * i;
        cplx p = -cexp(I * (M_PI * m / 8.0)) * warped;
        den *= (2.0 * fs - p);
        pz[i] = (2.0 * fs + p) / (2.0 * fs - p);
    }
    const double k = pow(warped, 4) * creal(1.0 / den);
    poly_from_roots(pz, 4, a);
    const double binom[5] = {1, 4, 6, 4, 1};
    for (int i = 0; i < 5; ++i) b[i] = k * binom[i];
    return 0;
}

/* cheby1(8, rp, wn, output='sos') -> sos[4][6], sections ordered like zpk2sos(pairing='nearest') */
int oracle_cheby1_sos8(double rp, double wn, double* sos) {
    if (!(wn > 0.0 && wn < 1.0) || !(rp > 0.0)) return -1;
    const int N = 8;
    const double fs = 2.0, warped = 2.0 * fs * tan(M_PI * wn / fs);
    const double eps = sqrt(pow(10.0, 0.1 * rp) - 1.0), mu = asinh(1.0 / eps) / N;
    cplx pz[8], kprod = 1.0, den = 1.0;
    for (int i = 0; i < N; ++i) {
        const int m = -N + 1 + 2 * i;
        const cplx p = -csinh(mu + I * (M_PI * m / (2.0 * N)));
        kprod *= -p;
        const cplx pw = p * warped;
        den *= (2.0 * fs - pw);
        pz[i] = (2.0 * fs + pw) / (2.0 * fs - pw);
    }
    double k = creal(kprod) / sqrt(1.0 + eps * eps) * pow(warped, N) * creal(1.0 / den);
    cplx up[4];
    int nu = 0;
    for (int i = 0; i < N; ++i) if (cimag(pz[i]) > 0.0 && nu < 4) up[nu++] = pz[i];
    if (nu != 4) return -1;
    for (int i = 0; i < 4; ++i)          /* farthest from the unit circle first */
        for (int j = i + 1; j < 4; ++j)
            if (fabs(1.0 - cabs(up[j])) > fabs(1.0 - cabs(up[i]))) { cplx t = up[i]; up[i] = up[j]; up[j] = t; }
    for (int s = 0; s < 4; ++s) {
        double* r = sos + 6 * s;
        r[0] = 1.0; r[1] = 2.0; r[2] = 1.0; r[3] = 1.0;
        r[4] = -2.0 * creal(up[s]);
        r[5] = creal(up[s]) * creal(up[s]) + cimag(up[s]) * cimag(up[s]);
    }
    sos[0] *= k; sos[1] *= k; sos[2] *= k;
    return 0;
}

static void lfilter_zi(const double* b, const double* a, int n, double* zi /* n-1 */) {
    double sb = 0, sa = 0, acc = 0;
    for (int i = 0; i < n; ++i) { sb += b[i]; sa += a[i]; }
    const double yinf = sb / sa;
    for (int i = n - 1; i >= 1; --i) { acc += b[i] - yinf * a[i]; zi[i - 1] = acc; }
}
static void sosfilt_zi(const double* sos, int ns, double* zi /* ns*2 */) {
    double scale = 1.0;
    for (int s = 0; s < ns; ++s) {
        const double* r = sos + 6 * s;
        double z[2];
        lfilter_zi(r, r + 3, 3, z);
        zi[2 * s] = scale * z[0];
        zi[2 * s + 1] = scale * z[1];
        scale *= (r[0] + r[1] + r[2]) / (r[3] + r[4] + r[5]);
    }
}

/* ---------------------------------------------------------------- zero-phase filtering --- */
static cplx* odd_ext(const cplx* x, int64_t n, int pad) {
    cplx* e = (cplx*)malloc(sizeof(cplx) * (size_t)(n + 2 * pad));
    for (int i = 0; i < pad; ++i) e[i] = 2.0 * x[0] - x[pad - i];
    memcpy(e + pad, x, sizeof(cplx) * (size_t)n);
    for (int i = 0; i < pad; ++i) e[pad + n + i] = 2.0 * x[n - 1] - x[n - 2 - i];
    return e;
}
static void sosfilt_inplace(const double* sos, int ns, cplx* x, int64_t n, const double* zi, cplx x0, int reverse) {
    cplx z[8][2];
    for (int s = 0; s < ns; ++s) { z[s][0] = zi[2 * s] * x0; z[s][1] = zi[2 * s + 1] * x0; }
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = reverse ? n - 1 - t : t;
        cplx v = x[i];
        for (int s = 0; s < ns; ++s) {
            const double* r = sos + 6 * s;
            const cplx y = r[0] * v + z[s][0];
            z[s][0] = r[1] * v - r[4] * y + z[s][1];
            z[s][1] = r[2] * v - r[5] * y;
            v = y;
        }
        x[i] = v;
    }
}
/* sosfiltfilt; returns 0, or -1 when the input is not longer than the padding (SciPy raises) */
static int sosfiltfilt(const double* sos, int ns, const cplx* x, int64_t n, cplx* out) {
    const int pad = 3 * (2 * ns + 1);
    if (n <= pad) return -1;
    double zi[16];
    sosfilt_zi(sos, ns, zi);
    cplx* e = odd_ext(x, n, pad);
    const int64_t m = n + 2 * pad;
    sosfilt_inplace(sos, ns, e, m, zi, e[0], 0);
    sosfilt_inplace(sos, ns, e, m, zi, e[m - 1], 1);
    memcpy(out, e + pad, sizeof(cplx) * (size_t)n);
    free(e);
    return 0;
}
static void lfilter_inplace(const double* b, const double* a, int nt, cplx* x, int64_t n, const double* zi, cplx x0, int reverse) {
    cplx z[8];
    for (int k = 0; k < nt - 1; ++k) z[k] = zi[k] * x0;
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = reverse ? n - 1 - t : t;
        const cplx v = x[i], y = b[0] * v + z[0];
        for (int k = 0; k < nt - 2; ++k) z[k] = b[k + 1] * v - a[k + 1] * y + z[k + 1];
        z[nt - 2] = b[nt - 1] * v - a[nt - 1] * y;
        x[i] = y;
    }
}
static int filtfilt(const double* b, const double* a, int nt, const cplx* x, int64_t n, cplx* out) {
    const int pad = 3 * nt;
    if (n <= pad) return -1;
    double zi[8];
    lfilter_zi(b, a, nt, zi);
    cplx* e = odd_ext(x, n, pad);
    const int64_t m = n + 2 * pad;
    lfilter_inplace(b, a, nt, e, m, zi, e[0], 0);
    lfilter_inplace(b, a, nt, e, m, zi, e[m - 1], 1);
    memcpy(out, e + pad, sizeof(cplx) * (size_t)n);
    free(e);
    return 0;
}

/* ---------------------------------------------------------------- process() --------------- */
/* slicer of processor.py:152-161 */
static uint8_t slice_phase(double ph) {
    if (ph < -5.0 * M_PI / 8.0) return 3;
    if (ph < -3.0 * M_PI / 8.0) return 2;
    if (ph < 3.0 * M_PI / 8.0) return 0;
    if (ph < 5.0 * M_PI / 8.0) return 1;
    return 3;
}

/*
 * SignalProcessor(sample_rate).process(x, freq_offset). x: n complex128 (interleaved). Outputs: dibits (room for n),
 * symbols (room for n complex128, interleaved), their counts, the timing phase. Returns 0.
 */
int oracle_process(const double* x_ri, int64_t n, double sample_rate, double freq_offset,
                   uint8_t* dibits, int64_t* n_dibits, double* symbols_ri, int64_t* n_symbols, int32_t* best_phase) {
    *n_dibits = 0; *n_symbols = 0; *best_phase = 0;
    if (n <= 0) return 0;
    cplx* cur = (cplx*)malloc(sizeof(cplx) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) cur[i] = x_ri[2 * i] + I * x_ri[2 * i + 1];
    int64_t len = n;
    double rate = sample_rate;
    /* processor.py:245-257: decimate when fs > 2 x 240 kHz */
    if (sample_rate > 240000.0 * 2) {
        const int q = (int)(sample_rate / 240000.0);
        if (q > 1) {
            double sos[24];
            cplx* y = (cplx*)malloc(sizeof(cplx) * (size_t)len);
            if (oracle_cheby1_sos8(0.05, 0.8 / q, sos) == 0 && sosfiltfilt(sos, 4, cur, len, y) == 0) {
                int64_t m = 0;
                for (int64_t i = 0; i < len; i += q) cur[m++] = y[i];
                len = m;
                rate = sample_rate / q;
            }
            free(y);
        }
    }
    /* processor.py:259-261 */
    if (freq_offset != 0.0) {
        for (int64_t i = 0; i < len; ++i) {
            const double t = (double)i / rate;
            const double ph = -(2.0 * M_PI * freq_offset * t);
            cur[i] *= cos(ph) + I * sin(ph);
        }
    }
    /* processor.py:264 -> :66-83 */
    {
        double wn = (25000.0 / 2) / (rate / 2), b[5], a[5];
        if (wn < 0.01) wn = 0.01;
        if (wn > 0.99) wn = 0.99;
        cplx* y = (cplx*)malloc(sizeof(cplx) * (size_t)len);
        if (oracle_butter4(wn, b, a) == 0 && filtfilt(b, a, 5, cur, len, y) == 0) memcpy(cur, y, sizeof(cplx) * (size_t)len);
        free(y);
    }
    /* processor.py:179-219 */
    const int sps = (int)(rate / 18000.0);
    cplx* sym = cur;
    int64_t ns = len;
    cplx* picked = NULL;
    if (sps > 1) {
        const int step = sps / 8 > 1 ? sps / 8 : 1;
        int best = 0;
        double best_pow = -1.0;
        for (int ph = 0; ph < sps; ph += step) {
            const int64_t cnt = (len - ph) / sps;
            if (cnt <= 0) continue;
            double acc = 0.0;
            for (int64_t k = 0; k < cnt; ++k) {
                const cplx v = cur[ph + k * sps];
                acc += creal(v) * creal(v) + cimag(v) * cimag(v);
            }
            const double mean = acc / (double)cnt;
            if (mean > best_pow) { best_pow = mean; best = ph; }
        }
        ns = (len - best) / sps;
        if (ns < 0) ns = 0;
        picked = (cplx*)malloc(sizeof(cplx) * (size_t)(ns > 0 ? ns : 1));
        for (int64_t k = 0; k < ns; ++k) picked[k] = cur[best + k * sps];
        sym = picked;
        *best_phase = best;
    }
    for (int64_t k = 0; k < ns; ++k) { symbols_ri[2 * k] = creal(sym[k]); symbols_ri[2 * k + 1] = cimag(sym[k]); }
    *n_symbols = ns;
    /* processor.py:120-166 */
    if (ns >= 2) {
        double mx = 0.0;
        for (int64_t k = 0; k < ns; ++k) { const double m = cabs(sym[k]); if (m > mx) mx = m; }
        for (int64_t k = 1; k < ns; ++k) {
            cplx s1 = sym[k], s0 = sym[k - 1];
            if (mx > 0) { s1 /= mx; s0 /= mx; }
            const cplx d = s1 * conj(s0);
            dibits[k - 1] = slice_phase(atan2(cimag(d), creal(d)));
        }
        *n_dibits = ns - 1;
    }
    free(picked);
    free(cur);
    return 0;
}

/* ---------------------------------------------------------------- frame sync -------------- */
static const uint8_t TS1[22] = {1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0};   /* decoder.py:196-197 */
static const uint8_t TS2[22] = {0, 1, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 1, 0, 0};   /* decoder.py:198-199 */

/* decoder.py:140-169 for symbols 0..3: (MSB, LSB) */
void oracle_symbols_to_bits(const uint8_t* dibits, int64_t n, uint8_t* bits) {
    for (int64_t i = 0; i < n; ++i) { bits[2 * i] = (dibits[i] >> 1) & 1; bits[2 * i + 1] = dibits[i] & 1; }
}

/* decoder.py:171-295: returns the number of positions, writes up to max_pos of them */
int oracle_find_sync(const uint8_t* bits, int64_t n_bits, double threshold, int32_t* pos, int max_pos, double* max_corr_out) {
    int count = 0;
    double max_corr = 0.0;
    *max_corr_out = 0.0;
    if (n_bits < 22) return 0;
    const int64_t nw = n_bits - 22 + 1;
    double* best = (double*)calloc((size_t)nw, sizeof(double));      /* best_here of visited offsets, 0 = not visited */
    int64_t i = 0;
    while (i < nw) {
        int found = 0;
        double best_here = 0.0;
        for (int k = 0; k < 2; ++k) {
            const uint8_t* ts = k == 0 ? TS1 : TS2;
            int m = 0;
            for (int j = 0; j < 22; ++j) m += bits[i + j] == ts[j];
            const double corr = (double)m / 22.0;
            if (corr > best_here) best_here = corr;
            if (corr > max_corr) max_corr = corr;
            if (corr >= threshold) {
                if (count < max_pos) pos[count] = (int32_t)i;
                ++count; found = 1;
                break;
            }
        }
        best[i] = best_here;
        i = found ? i + 250 : i + 1;
    }
    if (count == 0 && max_corr > 0.75 && max_corr >= (threshold - 0.15)) {
        const double adaptive = max_corr - 0.02 > 0.75 ? max_corr - 0.02 : 0.75;
        if (adaptive < threshold) {
            char* blocked = (char*)calloc((size_t)nw, 1);
            for (int64_t p = 0; p < nw; ++p) {
                if (best[p] > 0 && best[p] >= adaptive && !blocked[p]) {
                    if (count < max_pos) pos[count] = (int32_t)p;
                    ++count;
                    const int64_t lo = p - 250 > 0 ? p - 250 : 0, hi = p + 250 < nw ? p + 250 : nw;
                    for (int64_t t = lo; t < hi; ++t) blocked[t] = 1;
                }
            }
            free(blocked);
        }
    }
    free(best);
    *max_corr_out = max_corr;
    return count;
}

/* decoder.py:845-856 */
int oracle_sync_cascade(const uint8_t* bits, int64_t n_bits, int32_t* pos, int max_pos) {
    double mx = 0.0;
    int n = oracle_find_sync(bits, n_bits, 0.90, pos, max_pos, &mx);
    if (n) return n;
    n = oracle_find_sync(bits, n_bits, 0.85, pos, max_pos, &mx);
    if (n) return n;
    n = oracle_find_sync(bits, n_bits, 0.80, pos, max_pos, &mx);
    if (n) return n;
    if (mx >= 0.75) n = oracle_find_sync(bits, n_bits, mx - 0.02 > 0.75 ? mx - 0.02 : 0.75, pos, max_pos, &mx);
    return n;
}

/* ---------------------------------------------------------------- spectrum ---------------- */
/* ui/modern.py:1924-1934 for one row of nfft samples: hanning window, DFT, fftshift, 20 log10(|X|/N + 1e-20).
 * Plain O(N^2) DFT with exact phase reduction: this is a checker, not a benchmark. */
void oracle_spectrum_db(const double* x_ri, int nfft, double* out) {
    cplx* tw = (cplx*)malloc(sizeof(cplx) * (size_t)nfft);
    cplx* xw = (cplx*)malloc(sizeof(cplx) * (size_t)nfft);
    for (int i = 0; i < nfft; ++i) {
        const double ang = -2.0 * M_PI * (double)i / nfft;
        tw[i] = cos(ang) + I * sin(ang);
        const double w = 0.5 - 0.5 * cos(2.0 * M_PI * i / (nfft - 1));      /* np.hanning (symmetric) */
        xw[i] = (x_ri[2 * i] + I * x_ri[2 * i + 1]) * w;
    }
    for (int k = 0; k < nfft; ++k) {
        cplx acc = 0.0;
        for (int t = 0; t < nfft; ++t) acc += xw[t] * tw[(int)(((int64_t)k * t) % nfft)];
        out[(k + nfft / 2) % nfft] = 20.0 * log10(cabs(acc) / nfft + 1e-20);
    }
    free(tw); free(xw);
}
