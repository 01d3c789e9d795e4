// Host-side IIR design for the generic (exact-recursion) path.
//
// The reference obtains these coefficients from SciPy at run time:
//   scipy.signal.butter(4, wn)                      tetraear/signal/processor.py:78
//   scipy.signal.cheby1(8, 0.05, 0.8/q, 'sos')      inside scipy.signal.decimate, processor.py:254
// This is an independent implementation of the published method (analog prototype ->
// frequency pre-warp -> bilinear transform -> polynomial / second-order sections), so that the
// library needs no Python at run time. tests/test_host.py checks it against SciPy.
#pragma once
#include <cmath>
#include <complex>
#include <algorithm>
#include <vector>

namespace tetra {

typedef std::complex<double> cplx;

// Expand prod (z - r_k) into real polynomial coefficients (roots come in conjugate pairs).
static inline void poly_from_roots(const std::vector<cplx>& roots, std::vector<double>& out) {
    std::vector<cplx> c(1, cplx(1.0, 0.0));
    for (size_t k = 0; k < roots.size(); ++k) {
        std::vector<cplx> n(c.size() + 1, cplx(0.0, 0.0));
        for (size_t i = 0; i < c.size(); ++i) {
            n[i] += c[i];
            n[i + 1] -= c[i] * roots[k];
        }
        c.swap(n);
    }
    out.resize(c.size());
    for (size_t i = 0; i < c.size(); ++i) out[i] = c[i].real();
}

// Low-pass Butterworth, order 4, cutoff wn in (0,1) relative to Nyquist -> b[5], a[5].
static inline bool design_butter4(double wn, double* b, double* a) {
    if (!(wn > 0.0 && wn < 1.0)) return false;
    const int N = 4;
    const double fs = 2.0;
    const double warped = 2.0 * fs * std::tan(M_PI * wn / fs);
    std::vector<cplx> pz;
    cplx den(1.0, 0.0);
    for (int i = 0; i < N; ++i) {
        int m = -N + 1 + 2 * i;
        cplx p = -std::exp(cplx(0.0, M_PI * m / (2.0 * N)));   // analog prototype pole
        p *= warped;                                            // lp2lp
        den *= (2.0 * fs - p);
        pz.push_back((2.0 * fs + p) / (2.0 * fs - p));          // bilinear
    }
    const double k = std::pow(warped, N) * (1.0 / den).real();
    std::vector<double> av;
    poly_from_roots(pz, av);
    const double binom[5] = {1, 4, 6, 4, 1};                    // (z + 1)^4
    for (int i = 0; i < 5; ++i) { b[i] = k * binom[i]; a[i] = av[i]; }
    return true;
}

// Low-pass Chebyshev type I, order 8, ripple rp dB, edge wn -> 4 sections [b0 b1 b2 1 a1 a2],
// ordered like scipy's zpk2sos(pairing='nearest'): poles closest to the unit circle last,
// overall gain folded into the first section.
static inline bool design_cheby1_sos8(double rp, double wn, double* sos /*[4][6]*/) {
    if (!(wn > 0.0 && wn < 1.0) || !(rp > 0.0)) return false;
    const int N = 8;
    const double fs = 2.0;
    const double warped = 2.0 * fs * std::tan(M_PI * wn / fs);
    const double eps = std::sqrt(std::pow(10.0, 0.1 * rp) - 1.0);
    const double mu = std::asinh(1.0 / eps) / N;
    std::vector<cplx> pa;
    cplx kprod(1.0, 0.0);
    for (int i = 0; i < N; ++i) {
        int m = -N + 1 + 2 * i;
        double theta = M_PI * m / (2.0 * N);
        cplx p = -std::sinh(cplx(mu, theta));
        pa.push_back(p);
        kprod *= -p;
    }
    double k = kprod.real() / std::sqrt(1.0 + eps * eps);       // even order
    k *= std::pow(warped, N);                                   // lp2lp (no finite zeros)
    cplx den(1.0, 0.0);
    std::vector<cplx> pz;
    for (int i = 0; i < N; ++i) {
        cplx p = pa[i] * warped;
        den *= (2.0 * fs - p);
        pz.push_back((2.0 * fs + p) / (2.0 * fs - p));
    }
    k *= (1.0 / den).real();
    // keep one pole of each conjugate pair (imag >= 0), sort by distance to the unit circle
    std::vector<cplx> up;
    for (size_t i = 0; i < pz.size(); ++i) if (pz[i].imag() > 0.0) up.push_back(pz[i]);
    if (up.size() != 4) return false;
    std::sort(up.begin(), up.end(), [](const cplx& x, const cplx& y) {
        return std::fabs(1.0 - std::abs(x)) > std::fabs(1.0 - std::abs(y));
    });
    for (int s = 0; s < 4; ++s) {
        double* r = sos + 6 * s;
        r[0] = 1.0; r[1] = 2.0; r[2] = 1.0;
        r[3] = 1.0; r[4] = -2.0 * up[s].real(); r[5] = std::norm(up[s]);
    }
    sos[0] *= k; sos[1] *= k; sos[2] *= k;
    return true;
}

// scipy.signal.lfilter_zi for a[0] == 1: zi[j] = sum_{i>j} (b_i - y_inf a_i).
static inline void lfilter_zi(const double* b, const double* a, int n, double* zi /*[n-1]*/) {
    double sb = 0, sa = 0;
    for (int i = 0; i < n; ++i) { sb += b[i]; sa += a[i]; }
    const double yinf = sb / sa;
    double acc = 0;
    for (int i = n - 1; i >= 1; --i) { acc += b[i] - yinf * a[i]; zi[i - 1] = acc; }
}

// scipy.signal.sosfilt_zi.
static inline void sosfilt_zi(const double* sos, int n_sections, double* zi /*[n][2]*/) {
    double scale = 1.0;
    for (int s = 0; s < n_sections; ++s) {
        const double* r = sos + 6 * s;
        double z[2];
        lfilter_zi(r, r + 3, 3, z);
        zi[2 * s] = scale * z[0];
        zi[2 * s + 1] = scale * z[1];
        scale *= (r[0] + r[1] + r[2]) / (r[3] + r[4] + r[5]);
    }
}

}  // namespace tetra
