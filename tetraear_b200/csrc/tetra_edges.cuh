// KE: the block ends of the fused path -- the reference's zero-phase IIR recursions evaluated literally (fp64) over the
// windows next to each end of a block, where filtfilt is not shift-invariant. Two kernels for two batch regimes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tetra_exact.cuh"

namespace tetra {

// ----------------------------------------------------------------------------------------------
// k_exact_edges: the same recursions for the LEFT / RIGHT edge windows of the fast path, one
// thread per job, with the four biquad sections of the Chebyshev cascade SKEWED in time: at step
// s section k works on sample s - k and takes section k-1's output of the previous step from a
// register. The 4 sections x (re, im) of a step are then independent of each other, so the
// serial dependency per step is one section's own z0 -> y recurrence (2 DFMA) with 40 DFMA of
// independent work to fill the pipe, instead of a chain through the whole cascade.
// Input / scratch are read one block of steps ahead into registers. Scratch is job-major.
// ----------------------------------------------------------------------------------------------
constexpr int EXT_THREADS = 64;
constexpr int EXT_FB = 16;                // forward steps per prefetched block (float2 each)
constexpr int EXT_BB = 8;                 // backward steps per prefetched block (double2 each)

struct EdgeArgs {
    const float2* x;         // [C][pitch] complex64
    int64_t pitch, n;
    int64_t right_shift;     // added to a RIGHT job's row pointer: rows that hold only the two end windows of a block
                             // (uint8 ingest) keep sample i >= n - WR at column i + right_shift; 0 for whole blocks
    int32_t q, L, edge;
    ExactCoef cf;
    float2* y;               // [C][y_pitch], layout y_index(n, y_sps, y_rows)
    int64_t y_pitch;
    int32_t y_sps, y_rows;
    const int2* jobs;        // (carrier, mode), mode in {EX_LEFT, EX_RIGHT}
    int32_t n_jobs;
    double2* scr1;           // [n_jobs][w1] forward stage-1 output
    double2* scrz;           // [n_jobs][wz] stage-1 result
    double2* scr2;           // [n_jobs][wz + 2 PAD2] forward stage-2 output
    int64_t w1, wz;
    const double* fo;        // [C] freq offsets in Hz (device) or null: NCO between the two filters
    double fs_dec;           // sample rate after stage 1
};

// frequency_shift (processor.py:97-100) of stage-1 output sample m
__device__ __forceinline__ double2 edge_nco(double2 v, int m, double w_nco, double fs_dec) {
    if (w_nco == 0.0) return v;
    const double t = (double)m / fs_dec;
    double sn, cs;
    sincos(-(w_nco * t), &sn, &cs);
    return make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
}

struct SkewState {
    double z0[4][2], z1[4][2];            // biquad states [section][re, im]
    double yl[4][2];                      // each section's output of the previous step
};

// section k on input (xr, xi); scipy _sosfilt order of operations
__device__ __forceinline__ void skew_section(const ExactCoef& c, SkewState& st, int k, double xr, double xi) {
    const double b0 = c.sos[k][0], b1 = c.sos[k][1], b2 = c.sos[k][2], a1 = c.sos[k][4], a2 = c.sos[k][5];
    const double yr = b0 * xr + st.z0[k][0], yi = b0 * xi + st.z0[k][1];
    st.z0[k][0] = (b1 * xr + st.z1[k][0]) - a1 * yr;
    st.z0[k][1] = (b1 * xi + st.z1[k][1]) - a1 * yi;
    st.z1[k][0] = b2 * xr - a2 * yr;
    st.z1[k][1] = b2 * xi - a2 * yi;
    st.yl[k][0] = yr; st.yl[k][1] = yi;
}
// all four sections active: descending k so that yl[k-1] still holds the previous step's output
__device__ __forceinline__ void skew_step_all(const ExactCoef& c, SkewState& st, double xr, double xi) {
    skew_section(c, st, 3, st.yl[2][0], st.yl[2][1]);
    skew_section(c, st, 2, st.yl[1][0], st.yl[1][1]);
    skew_section(c, st, 1, st.yl[0][0], st.yl[0][1]);
    skew_section(c, st, 0, xr, xi);
}
// sections k_lo..k_hi only (pipeline fill / drain)
__device__ __forceinline__ void skew_step_some(const ExactCoef& c, SkewState& st, double xr, double xi, int k_lo, int k_hi) {
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        if (k < k_lo || k > k_hi) continue;
        if (k == 0) skew_section(c, st, 0, xr, xi);
        else skew_section(c, st, k, st.yl[k - 1][0], st.yl[k - 1][1]);
    }
}
__device__ __forceinline__ void skew_init(const ExactCoef& c, SkewState& st, double2 x0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        st.z0[k][0] = c.zi1[k][0] * x0.x; st.z0[k][1] = c.zi1[k][0] * x0.y;
        st.z1[k][0] = c.zi1[k][1] * x0.x; st.z1[k][1] = c.zi1[k][1] * x0.y;
        st.yl[k][0] = st.yl[k][1] = 0.0;
    }
}

__global__ void __launch_bounds__(EXT_THREADS) k_exact_edges(const EdgeArgs a) {
    const int j = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (j >= a.n_jobs) return;
    const int car = a.jobs[j].x, mode = a.jobs[j].y;
    const float2* __restrict__ xc = a.x + (int64_t)car * a.pitch + (mode == EX_RIGHT ? a.right_shift : 0);
    const int L = a.L, E = a.edge, q = a.q;
    const int64_t n = a.n;
    int m_lo = 0, m_hi = L, o_lo = 0, o_hi = L;
    if (mode == EX_LEFT) { o_hi = min(L, E); m_hi = min(L, E + EX_T2); }
    else { o_lo = max(0, L - E); m_lo = max(0, L - E - EX_T2); }
    const int64_t tot = n + 2 * EX_PAD1;
    int64_t e_lo = 0, e_hi = tot;
    if (mode == EX_LEFT) e_hi = min(tot, (int64_t)EX_PAD1 + (int64_t)q * (m_hi - 1) + 1 + EX_T1);
    else e_lo = max((int64_t)0, (int64_t)EX_PAD1 + (int64_t)q * m_lo - EX_T1);
    const int nf = (int)(e_hi - e_lo);
    double2* __restrict__ s1 = a.scr1 + (int64_t)j * a.w1;      // [nf]
    double2* __restrict__ sz = a.scrz + (int64_t)j * a.wz;      // [m_hi - m_lo]
    double2* __restrict__ s2 = a.scr2 + (int64_t)j * (a.wz + 2 * EX_PAD2);

    auto xat = [&](int64_t i) { const float2 v = __ldg(xc + i); return make_double2((double)v.x, (double)v.y); };
    SkewState st;

    // ---------------- stage 1, forward: f = sosfilt(ext[e_lo .. e_hi)), sample s <-> e = e_lo + s ----------------
    {
        skew_init(a.cf, st, ex_oddext(xat, n, EX_PAD1, e_lo));
        auto slow = [&](int s) {                          // any step: pads, pipeline fill and drain
            const int k_lo = max(0, s - nf + 1), k_hi = min(3, s);
            double2 X = make_double2(0.0, 0.0);
            if (k_lo == 0) X = ex_oddext(xat, n, EX_PAD1, e_lo + s);
            skew_step_some(a.cf, st, X.x, X.y, k_lo, k_hi);
            if (k_hi == 3) s1[s - 3] = make_double2(st.yl[3][0], st.yl[3][1]);
        };
        int s = 0;
        while (s < nf + 3 && (s < 3 || e_lo + s < EX_PAD1)) { slow(s); ++s; }
        const int fast_end = (int)min((int64_t)nf, (int64_t)EX_PAD1 + n - e_lo);   // samples below lie inside the block
        if (s + EXT_FB <= fast_end) {
            float2 cur[EXT_FB], nxt[EXT_FB];
            const float2* p = xc + (e_lo + s - EX_PAD1);
#pragma unroll
            for (int u = 0; u < EXT_FB; ++u) cur[u] = __ldg(p + u);
            while (s + EXT_FB <= fast_end) {
                const bool more = s + 2 * EXT_FB <= fast_end;
                const float2* pn = p + (more ? EXT_FB : 0);
#pragma unroll
                for (int u = 0; u < EXT_FB; ++u) nxt[u] = __ldg(pn + u);
#pragma unroll
                for (int u = 0; u < EXT_FB; ++u) {
                    skew_step_all(a.cf, st, (double)cur[u].x, (double)cur[u].y);
                    s1[s + u - 3] = make_double2(st.yl[3][0], st.yl[3][1]);
                }
#pragma unroll
                for (int u = 0; u < EXT_FB; ++u) cur[u] = nxt[u];
                s += EXT_FB; p += EXT_FB;
            }
        }
        while (s < nf + 3) { slow(s); ++s; }
    }
    // ---------------- stage 1, backward over the forward output (step s <-> e = e_hi - 1 - s), keep every q-th ----------------
    {
        const int64_t e_stop = max(e_lo, (int64_t)EX_PAD1 + (int64_t)q * m_lo);   // >= PAD1
        const int nb = (int)(e_hi - e_stop);
        skew_init(a.cf, st, s1[nf - 1]);
        // decimation bookkeeping of the emitted samples: input index i = e - PAD1 = q m + r
        const int64_t i0 = e_hi - 1 - EX_PAD1;
        int m = (int)(i0 / q), r = (int)(i0 % q);
        const double w_nco = a.fo ? (2.0 * M_PI) * a.fo[car] : 0.0;
        auto emit = [&]() {
            if (r == 0) {
                if ((int64_t)q * m < n && m >= m_lo && m < m_hi)
                    sz[m - m_lo] = edge_nco(make_double2(st.yl[3][0], st.yl[3][1]), m, w_nco, a.fs_dec);
                r = q; --m;
            }
            --r;
        };
        auto slow = [&](int s) {
            const int k_lo = max(0, s - nb + 1), k_hi = min(3, s);
            double2 X = make_double2(0.0, 0.0);
            if (k_lo == 0) X = s1[nf - 1 - s];
            skew_step_some(a.cf, st, X.x, X.y, k_lo, k_hi);
            if (k_hi == 3) emit();
        };
        int s = 0;
        while (s < nb + 3 && s < 3) { slow(s); ++s; }
        if (s + EXT_BB <= nb) {
            double2 cur[EXT_BB], nxt[EXT_BB];
            const double2* p = s1 + (nf - 1 - s);
#pragma unroll
            for (int u = 0; u < EXT_BB; ++u) cur[u] = p[-u];
            while (s + EXT_BB <= nb) {
                const bool more = s + 2 * EXT_BB <= nb;
                const double2* pn = p - (more ? EXT_BB : 0);
#pragma unroll
                for (int u = 0; u < EXT_BB; ++u) nxt[u] = pn[-u];
#pragma unroll
                for (int u = 0; u < EXT_BB; ++u) {
                    skew_step_all(a.cf, st, cur[u].x, cur[u].y);
                    emit();
                }
#pragma unroll
                for (int u = 0; u < EXT_BB; ++u) cur[u] = nxt[u];
                s += EXT_BB; p -= EXT_BB;
            }
        }
        while (s < nb + 3) { slow(s); ++s; }
    }
    // ---------------- stage 2: filtfilt(b, a) on z ----------------
    {
        auto zat = [&](int64_t mm) { return sz[mm - m_lo]; };
        auto z2 = [&](int64_t e) { return ex_oddext(zat, (int64_t)L, EX_PAD2, e); };   // folds around the true block ends
        const int64_t tot2 = (int64_t)L + 2 * EX_PAD2;
        int64_t f_lo = 0, f_hi = tot2;
        if (mode == EX_LEFT) f_hi = min(tot2, (int64_t)EX_PAD2 + m_hi);
        else f_lo = (int64_t)EX_PAD2 + m_lo;
        BaState bs;
        ba_init(bs, a.cf, z2(f_lo));
        for (int64_t e = f_lo; e < f_hi; e += EX_U) {
            double2 g[EX_U];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) g[u] = z2(min(e + u, f_hi - 1));
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                const double2 v = ba_step(bs, a.cf, g[u]);
                if (e + u < f_hi) s2[e + u - f_lo] = v;
            }
        }
        ba_init(bs, a.cf, s2[f_hi - 1 - f_lo]);
        const int64_t f_stop = (int64_t)EX_PAD2 + o_lo;
        float2* yc = a.y + (int64_t)car * a.y_pitch;
        for (int64_t e = f_hi - 1; e >= f_stop; e -= EX_U) {
            double2 g[EX_U];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) g[u] = s2[max(e - u, f_stop) - f_lo];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                if (e - u >= f_stop) {
                    const double2 v = ba_step(bs, a.cf, g[u]);
                    const int64_t mm = e - u - EX_PAD2;
                    if (mm >= o_lo && mm < o_hi) yc[y_index(mm, a.y_sps, a.y_rows)] = make_float2((float)v.x, (float)v.y);
                }
            }
        }
    }
}

// index ranges of one LEFT / RIGHT edge job (shared by host planning and the kernels)
struct EdgeRange {
    int m_lo, m_hi, o_lo, o_hi;            // stage-1 outputs [m_lo, m_hi), kept outputs [o_lo, o_hi)
    int64_t e_lo, e_hi, e_stop;            // stage-1 extended-input window, backward pass stops at e_stop
    int64_t f_lo, f_hi, f_stop;            // same for stage 2
};
__host__ __device__ inline EdgeRange edge_range(int mode, int64_t n, int L, int q, int E) {
    EdgeRange r;
    r.m_lo = 0; r.m_hi = L; r.o_lo = 0; r.o_hi = L;
    if (mode == EX_LEFT) { r.o_hi = L < E ? L : E; r.m_hi = L < E + EX_T2 ? L : E + EX_T2; }
    else { r.o_lo = L - E > 0 ? L - E : 0; r.m_lo = L - E - EX_T2 > 0 ? L - E - EX_T2 : 0; }
    const int64_t tot = n + 2 * EX_PAD1;
    r.e_lo = 0; r.e_hi = tot;
    if (mode == EX_LEFT) { const int64_t v = (int64_t)EX_PAD1 + (int64_t)q * (r.m_hi - 1) + 1 + EX_T1; r.e_hi = v < tot ? v : tot; }
    else { const int64_t v = (int64_t)EX_PAD1 + (int64_t)q * r.m_lo - EX_T1; r.e_lo = v > 0 ? v : 0; }
    const int64_t es = (int64_t)EX_PAD1 + (int64_t)q * r.m_lo;
    r.e_stop = es > r.e_lo ? es : r.e_lo;
    const int64_t tot2 = (int64_t)L + 2 * EX_PAD2;
    r.f_lo = 0; r.f_hi = tot2;
    if (mode == EX_LEFT) { const int64_t v = (int64_t)EX_PAD2 + r.m_hi; r.f_hi = v < tot2 ? v : tot2; }
    else r.f_lo = (int64_t)EX_PAD2 + r.m_lo;
    r.f_stop = (int64_t)EX_PAD2 + r.o_lo;
    return r;
}

}  // namespace tetra
