// Exact-recursion kernels: the reference's zero-phase IIR chain evaluated literally (fp64,
// SciPy's sosfiltfilt / filtfilt padding and initial-condition rules) on a window of a block.
//
// Used (a) for the block edges of the fast path, where filtfilt is not shift-invariant
// (SURVEY.md 7.3 H3), (b) as the complete path for short blocks, other sample rates and
// non-zero freq_offset, (c) by the public helper entry points (filter_signal, ...).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tetra {

constexpr int EX_PAD1 = 27;   // sosfiltfilt: 3 * (2*n_sections + 1), n_sections = 4
constexpr int EX_PAD2 = 15;   // filtfilt: 3 * max(len(a), len(b)) = 3 * 5
// Warm-up of an approximate start / end. What reaches a kept output has crossed the stage-1 warm-up AND the EX_T2
// stage-2 samples between the window end and the kept range (0.8837^160 = 3e-9), so stage 1 needs little of its own.
constexpr int EX_T1 = 256;    // stage 1 (pole radius 0.9821 per input sample)
constexpr int EX_T2 = 160;    // same for stage 2 (pole radius 0.8837 -> 3e-9)
constexpr int EX_U = 8;       // recursion steps per load group
constexpr int K_EDGE = 160;   // outputs per block end owned by the exact path (= K1_EDGE)
constexpr int K_EDGE_MAX_S2 = K_EDGE + 2 * EX_T2 + 2 * EX_PAD2;   // longest stage-2 window of an edge job

enum { EX_FULL = 0, EX_LEFT = 1, EX_RIGHT = 2 };

// Layout of the filtered 240 kS/s stream y of one carrier in HBM: timing-phase major, y[(n % sps) * rows + n / sps],
// so that the symbols of the phase the timing pick chooses (every sps-th sample, processor.py:213-215) are contiguous.
// rows == 0 means natural order.
__host__ __device__ inline int64_t y_index(int64_t n, int sps, int rows) {
    return rows > 0 ? (n % sps) * (int64_t)rows + n / sps : n;
}

struct ExactCoef {
    double sos[4][6];
    double zi1[4][2];
    double b[5], a[5], zi2[4];
};

struct ExactArgs {
    const float2* x32;       // complex64 input [C][pitch] (batch path) ...
    const double2* x64;      // ... or complex128 input (helpers); exactly one is non-null
    int64_t pitch;
    int64_t n;               // samples per carrier
    int32_t q;               // decimation factor (1: stage 1 absent)
    int32_t L;               // length after stage 1
    int32_t has_s1, has_s2;
    ExactCoef cf;
    const double* fo;        // [C] freq offsets (device) or null
    double fs_dec;           // sample rate after stage 1
    int32_t y_sps, y_rows;   // layout of y32 (y_index); y64 is always in natural order
    float2* y32;             // output [C][y_pitch] (complex64) ...
    double2* y64;            // ... or complex128
    int64_t y_pitch;
    const int2* jobs;        // (carrier, mode)
    int32_t n_jobs;
    int32_t edge;            // E: outputs produced by LEFT / RIGHT jobs
    double2* scr1;           // [W1][n_jobs]  forward stage-1 output
    double2* scrz;           // [WZ][n_jobs]  stage-1 result (after NCO)
    double2* scr2;           // [WZ + 2*PAD2][n_jobs] forward stage-2 output
};

__device__ __forceinline__ double2 ex_load(const ExactArgs& a, int64_t base, int64_t i) {
    if (a.x32) {
        float2 v = __ldg(a.x32 + base + i);
        return make_double2((double)v.x, (double)v.y);
    }
    return a.x64[base + i];
}

// odd extension (scipy.signal._arraytools.odd_ext) of a length-n sequence by `pad`
template <class F>
__device__ __forceinline__ double2 ex_oddext(F&& at, int64_t n, int pad, int64_t e) {
    const int64_t i = e - pad;
    if (i < 0) {
        double2 x0 = at(0), xr = at(-i);
        return make_double2(2.0 * x0.x - xr.x, 2.0 * x0.y - xr.y);
    }
    if (i >= n) {
        double2 x1 = at(n - 1), xr = at(2 * (n - 1) - i);
        return make_double2(2.0 * x1.x - xr.x, 2.0 * x1.y - xr.y);
    }
    return at(i);
}

struct SosState { double z[4][2][2]; };   // [section][state][re/im]

__device__ __forceinline__ void sos_init(SosState& s, const ExactCoef& c, double2 v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.z[k][0][0] = c.zi1[k][0] * v.x; s.z[k][0][1] = c.zi1[k][0] * v.y;
        s.z[k][1][0] = c.zi1[k][1] * v.x; s.z[k][1][1] = c.zi1[k][1] * v.y;
    }
}
// scipy _sosfilt inner loop: x_n = b0 x + z0; z0 = b1 x - a1 x_n + z1; z1 = b2 x - a2 x_n
__device__ __forceinline__ double2 sos_step(SosState& s, const ExactCoef& c, double2 v) {
    double xr = v.x, xi = v.y;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double b0 = c.sos[k][0], b1 = c.sos[k][1], b2 = c.sos[k][2], a1 = c.sos[k][4], a2 = c.sos[k][5];
        const double yr = b0 * xr + s.z[k][0][0];
        const double yi = b0 * xi + s.z[k][0][1];
        s.z[k][0][0] = b1 * xr - a1 * yr + s.z[k][1][0];
        s.z[k][0][1] = b1 * xi - a1 * yi + s.z[k][1][1];
        s.z[k][1][0] = b2 * xr - a2 * yr;
        s.z[k][1][1] = b2 * xi - a2 * yi;
        xr = yr; xi = yi;
    }
    return make_double2(xr, xi);
}

struct BaState { double z[4][2]; };
__device__ __forceinline__ void ba_init(BaState& s, const ExactCoef& c, double2 v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { s.z[k][0] = c.zi2[k] * v.x; s.z[k][1] = c.zi2[k] * v.y; }
}
// scipy lfilter (transposed direct form II), order 4
__device__ __forceinline__ double2 ba_step(BaState& s, const ExactCoef& c, double2 v) {
    const double yr = c.b[0] * v.x + s.z[0][0];
    const double yi = c.b[0] * v.y + s.z[0][1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s.z[k][0] = c.b[k + 1] * v.x - c.a[k + 1] * yr + s.z[k + 1][0];
        s.z[k][1] = c.b[k + 1] * v.y - c.a[k + 1] * yi + s.z[k + 1][1];
    }
    s.z[3][0] = c.b[4] * v.x - c.a[4] * yr;
    s.z[3][1] = c.b[4] * v.y - c.a[4] * yi;
    return make_double2(yr, yi);
}

// One thread = one (carrier, window) job. Scratch is interleaved [step][job] so that the
// threads of a warp touch consecutive addresses.
__global__ void __launch_bounds__(64) k_exact_chain(const ExactArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_jobs) return;
    const int car = a.jobs[j].x, mode = a.jobs[j].y;
    const int64_t xb = (int64_t)car * a.pitch;
    const int64_t nj = a.n_jobs;
    const int L = a.L, E = a.edge, q = a.q;
    const int64_t n = a.n;

    // ranges: z indices [m_lo, m_hi) are produced by stage 1, outputs [o_lo, o_hi) by stage 2
    int m_lo = 0, m_hi = L, o_lo = 0, o_hi = L;
    if (mode == EX_LEFT) { o_hi = min(L, E); m_hi = min(L, E + EX_T2); }
    else if (mode == EX_RIGHT) { o_lo = max(0, L - E); m_lo = max(0, L - E - EX_T2); }

    auto xat = [&](int64_t i) { return ex_load(a, xb, i); };
    const double w_nco = a.fo ? (2.0 * M_PI) * a.fo[car] : 0.0;
    auto nco = [&](double2 v, int m) {
        if (w_nco == 0.0) return v;
        const double t = (double)m / a.fs_dec;
        double sn, cs;
        sincos(-(w_nco * t), &sn, &cs);
        return make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    };

    // ---------------- stage 1: sosfiltfilt + [::q] ----------------
    if (a.has_s1) {
        const int64_t tot = n + 2 * EX_PAD1;
        int64_t e_lo = 0, e_hi = tot;
        if (mode == EX_LEFT) e_hi = min(tot, (int64_t)EX_PAD1 + (int64_t)q * (m_hi - 1) + 1 + EX_T1);
        else if (mode == EX_RIGHT) e_lo = max((int64_t)0, (int64_t)EX_PAD1 + (int64_t)q * m_lo - EX_T1);
        // Both passes run in groups of EX_U steps with the NEXT group's loads in flight while the
        // current group's (serially dependent) recursion steps execute.
        double2* __restrict__ scr1 = a.scr1;
        double2* __restrict__ scrz = a.scrz;
        auto ld_fwd = [&](double2* g, int64_t e) {
            if (e >= EX_PAD1 && e + EX_U <= EX_PAD1 + n) {            // group entirely inside the block
#pragma unroll
                for (int u = 0; u < EX_U; ++u) g[u] = xat(e - EX_PAD1 + u);
            } else {
#pragma unroll
                for (int u = 0; u < EX_U; ++u) g[u] = ex_oddext(xat, n, EX_PAD1, min(e + u, e_hi - 1));
            }
        };
        auto st_fwd = [&](SosState& st, const double2* g, int64_t e) {
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                const double2 v = sos_step(st, a.cf, g[u]);
                if (e + u < e_hi) scr1[(e + u - e_lo) * nj + j] = v;
            }
        };
        SosState st;
        sos_init(st, a.cf, ex_oddext(xat, n, EX_PAD1, e_lo));
        {
            double2 ga[EX_U], gb[EX_U];
            ld_fwd(ga, e_lo);
            for (int64_t e = e_lo; e < e_hi; e += 2 * EX_U) {
                if (e + EX_U < e_hi) ld_fwd(gb, e + EX_U);
                st_fwd(st, ga, e);
                if (e + EX_U < e_hi) {
                    if (e + 2 * EX_U < e_hi) ld_fwd(ga, e + 2 * EX_U);
                    st_fwd(st, gb, e + EX_U);
                }
            }
        }
        __threadfence_block();
        sos_init(st, a.cf, scr1[(e_hi - 1 - e_lo) * nj + j]);
        const int64_t e_stop = max(e_lo, (int64_t)EX_PAD1 + (int64_t)q * m_lo);   // >= EX_PAD1: i below is >= 0
        int64_t i = e_hi - 1 - EX_PAD1;                  // input index of the step being produced
        int m = (int)(i / q), r = (int)(i % q);          // i = q*m + r, kept incrementally
        auto ld_bwd = [&](double2* g, int64_t e) {
#pragma unroll
            for (int u = 0; u < EX_U; ++u) g[u] = scr1[(max(e - u, e_stop) - e_lo) * nj + j];
        };
        auto st_bwd = [&](const double2* g, int64_t e) {
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                if (e - u >= e_stop) {
                    const double2 v = sos_step(st, a.cf, g[u]);
                    if (r == 0) {
                        if (i < n && m >= m_lo && m < m_hi) scrz[(int64_t)(m - m_lo) * nj + j] = nco(v, m);
                        r = q; --m;
                    }
                    --r; --i;
                }
            }
        };
        {
            double2 ga[EX_U], gb[EX_U];
            ld_bwd(ga, e_hi - 1);
            for (int64_t e = e_hi - 1; e >= e_stop; e -= 2 * EX_U) {
                if (e - EX_U >= e_stop) ld_bwd(gb, e - EX_U);
                st_bwd(ga, e);
                if (e - EX_U >= e_stop) {
                    if (e - 2 * EX_U >= e_stop) ld_bwd(ga, e - 2 * EX_U);
                    st_bwd(gb, e - EX_U);
                }
            }
        }
    } else {
        for (int m = m_lo; m < m_hi; ++m) a.scrz[(int64_t)(m - m_lo) * nj + j] = nco(xat(m), m);
    }

    // ---------------- stage 2: filtfilt(b, a) ----------------
    auto zat = [&](int64_t m) { return a.scrz[(m - m_lo) * nj + j]; };
    auto put = [&](int m, double2 v) {
        if (a.y32) a.y32[(int64_t)car * a.y_pitch + y_index(m, a.y_sps, a.y_rows)] = make_float2((float)v.x, (float)v.y);
        else a.y64[(int64_t)car * a.y_pitch + m] = v;
    };
    if (a.has_s2) {
        const int64_t tot = (int64_t)L + 2 * EX_PAD2;
        int64_t e_lo = 0, e_hi = tot;
        if (mode == EX_LEFT) e_hi = min(tot, (int64_t)EX_PAD2 + m_hi);
        else if (mode == EX_RIGHT) e_lo = (int64_t)EX_PAD2 + m_lo;
        // inside a window the odd extension only ever folds around the true block ends
        auto z2 = [&](int64_t e) { return ex_oddext(zat, (int64_t)L, EX_PAD2, e); };
        BaState st;
        ba_init(st, a.cf, z2(e_lo));
        for (int64_t e = e_lo; e < e_hi; e += EX_U) {
            double2 in[EX_U];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) in[u] = z2(min(e + u, e_hi - 1));
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                const double2 v = ba_step(st, a.cf, in[u]);
                if (e + u < e_hi) a.scr2[(e + u - e_lo) * nj + j] = v;
            }
        }
        ba_init(st, a.cf, a.scr2[(e_hi - 1 - e_lo) * nj + j]);
        const int64_t e_stop = (int64_t)EX_PAD2 + o_lo;
        for (int64_t e = e_hi - 1; e >= e_stop; e -= EX_U) {
            double2 in[EX_U];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) in[u] = a.scr2[(max(e - u, e_stop) - e_lo) * nj + j];
#pragma unroll
            for (int u = 0; u < EX_U; ++u) {
                if (e - u >= e_stop) {
                    const double2 v = ba_step(st, a.cf, in[u]);
                    const int64_t m = e - u - EX_PAD2;
                    if (m >= o_lo && m < o_hi) put((int)m, v);
                }
            }
        }
    } else {
        for (int m = o_lo; m < o_hi; ++m) put(m, zat(m));
    }
}

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
    const float2* __restrict__ xc = a.x + (int64_t)car * a.pitch;
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

// ----------------------------------------------------------------------------------------------
// k_exact_edges_warp: the same edge windows with one WARP per job, for batches too small to hide a
// thread's serial recursion behind the fused kernel. A pass of n steps is cut into 32 chunks, one per
// lane. Linearity of the recursion does the rest:
//   1. every lane runs its chunk from a zero state (lane 0 from the true initial state) -> end state E_c
//   2. the true state at the start of chunk c+1 is T_{c+1} = M T_c + E_c with M the zero-input
//      transition over one chunk (host-computed, with M^2, M^4, M^8, M^16): a 5-round warp scan
//   3. every lane re-runs its chunk from its true start state and emits.
// The serial depth drops from n to 2 n / 32 steps plus the scan.
// ----------------------------------------------------------------------------------------------
constexpr int EXW_BLK = 8;                // steps per prefetched block inside a chunk
constexpr int EXW_S2MAX = 16;             // stage-2 chunk length bound (inputs of a chunk stay in registers)
static_assert(K_EDGE_MAX_S2 <= 32 * EXW_S2MAX, "stage-2 window does not fit 32 chunks of EXW_S2MAX");

struct EdgeWarpArgs {
    EdgeArgs e;
    // zero-input chunk transitions, row-major [variant][power r = 0..4][DIM][DIM];
    // variants: 0 LEFT fwd, 1 LEFT bwd, 2 RIGHT fwd, 3 RIGHT bwd
    const double* m1;        // stage 1, DIM = 8: state order (z0_0, z1_0, z0_1, z1_1, ...)
    const double* m2;        // stage 2, DIM = 4
};

// V[c] <- sum_{j <= c} M^{c-j} V[j]  over the lanes of the warp (both components), M^(2^r) at mp + r*DIM*DIM
template <int DIM>
__device__ __forceinline__ void warp_affine_scan(double (&v)[DIM][2], const double* __restrict__ mp, int lane) {
#pragma unroll 1
    for (int r = 0; r < 5; ++r) {
        const int off = 1 << r;
        double w[DIM][2];
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            w[i][0] = __shfl_up_sync(0xffffffffu, v[i][0], off);
            w[i][1] = __shfl_up_sync(0xffffffffu, v[i][1], off);
        }
        if (lane >= off) {
            const double* m = mp + r * DIM * DIM;
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
                double ar = v[i][0], ai = v[i][1];
#pragma unroll
                for (int jj = 0; jj < DIM; ++jj) {
                    const double mij = m[i * DIM + jj];
                    ar += mij * w[jj][0];
                    ai += mij * w[jj][1];
                }
                v[i][0] = ar; v[i][1] = ai;
            }
        }
    }
}

// one chunk of the (time-skewed) biquad cascade: samples [0, len) of `load`/`cook`, states in/out in st
template <class Raw, class Load, class Cook, class Emit>
__device__ __forceinline__ void sos_run_chunk(const ExactCoef& cf, SkewState& st, int len, Load&& load, Cook&& cook, Emit&& emit) {
#pragma unroll
    for (int k = 0; k < 4; ++k) st.yl[k][0] = st.yl[k][1] = 0.0;
    auto slow = [&](int s) {
        const int k_lo = max(0, s - len + 1), k_hi = min(3, s);
        double2 X = make_double2(0.0, 0.0);
        if (k_lo == 0) X = cook(load(s), s);
        skew_step_some(cf, st, X.x, X.y, k_lo, k_hi);
        if (k_hi == 3) emit(s - 3, st.yl[3][0], st.yl[3][1]);
    };
    int s = 0;
    for (; s < min(3, len + 3); ++s) slow(s);
    if (s + EXW_BLK <= len) {
        Raw cur[EXW_BLK], nxt[EXW_BLK];
#pragma unroll
        for (int u = 0; u < EXW_BLK; ++u) cur[u] = load(s + u);
        while (s + EXW_BLK <= len) {
            const int sn = s + 2 * EXW_BLK <= len ? s + EXW_BLK : s;
#pragma unroll
            for (int u = 0; u < EXW_BLK; ++u) nxt[u] = load(sn + u);
#pragma unroll
            for (int u = 0; u < EXW_BLK; ++u) {
                const double2 X = cook(cur[u], s + u);
                skew_step_all(cf, st, X.x, X.y);
                emit(s + u - 3, st.yl[3][0], st.yl[3][1]);
            }
#pragma unroll
            for (int u = 0; u < EXW_BLK; ++u) cur[u] = nxt[u];
            s += EXW_BLK;
        }
    }
    for (; s < len + 3; ++s) slow(s);
}

__device__ __forceinline__ void skew_to_vec(const SkewState& st, double (&v)[8][2]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[2 * k][0] = st.z0[k][0]; v[2 * k][1] = st.z0[k][1];
        v[2 * k + 1][0] = st.z1[k][0]; v[2 * k + 1][1] = st.z1[k][1];
    }
}
__device__ __forceinline__ void vec_to_skew(const double (&v)[8][2], SkewState& st) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        st.z0[k][0] = v[2 * k][0]; st.z0[k][1] = v[2 * k][1];
        st.z1[k][0] = v[2 * k + 1][0]; st.z1[k][1] = v[2 * k + 1][1];
    }
}

// chunk-parallel pass of the biquad cascade over n_steps samples; init = zi * x0 state of sample 0
template <class Raw, class Load, class Cook, class Emit>
__device__ __forceinline__ void sos_pass_warp(const ExactCoef& cf, int n_steps, double2 x0, const double* __restrict__ mp,
                                              int lane, Load&& load, Cook&& cook, Emit&& emit) {
    const int lc = (n_steps + 31) / 32;
    const int start = lane * lc;
    const int len = max(0, min(lc, n_steps - start));
    SkewState st;
    skew_init(cf, st, lane == 0 ? x0 : make_double2(0.0, 0.0));
    auto ld = [&](int sl) { return load(start + sl); };
    auto ck = [&](Raw r, int sl) { return cook(r, start + sl); };
    sos_run_chunk<Raw>(cf, st, len, ld, ck, [](int, double, double) {});
    double v[8][2];
    skew_to_vec(st, v);
    warp_affine_scan<8>(v, mp, lane);
    double t[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        t[i][0] = __shfl_up_sync(0xffffffffu, v[i][0], 1);
        t[i][1] = __shfl_up_sync(0xffffffffu, v[i][1], 1);
    }
    if (lane == 0) skew_init(cf, st, x0); else vec_to_skew(t, st);
    sos_run_chunk<Raw>(cf, st, len, ld, ck, [&](int sl, double yr, double yi) { emit(start + sl, yr, yi); });
}

// chunk-parallel pass of the order-4 (b, a) filter; the chunk's inputs stay in registers
template <class Fetch, class Emit>
__device__ __forceinline__ void ba_pass_warp(const ExactCoef& cf, int n_steps, double2 x0, const double* __restrict__ mp,
                                             int lane, Fetch&& fetch, Emit&& emit) {
    const int lc = (n_steps + 31) / 32;                  // <= EXW_S2MAX
    const int start = lane * lc;
    const int len = max(0, min(lc, n_steps - start));
    double2 in[EXW_S2MAX];
#pragma unroll
    for (int u = 0; u < EXW_S2MAX; ++u) in[u] = u < len ? fetch(start + u) : make_double2(0.0, 0.0);
    BaState bs;
    ba_init(bs, cf, lane == 0 ? x0 : make_double2(0.0, 0.0));
#pragma unroll
    for (int u = 0; u < EXW_S2MAX; ++u) if (u < len) ba_step(bs, cf, in[u]);
    double v[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i][0] = bs.z[i][0]; v[i][1] = bs.z[i][1]; }
    warp_affine_scan<4>(v, mp, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double tr = __shfl_up_sync(0xffffffffu, v[i][0], 1), ti = __shfl_up_sync(0xffffffffu, v[i][1], 1);
        bs.z[i][0] = tr; bs.z[i][1] = ti;
    }
    if (lane == 0) ba_init(bs, cf, x0);
#pragma unroll
    for (int u = 0; u < EXW_S2MAX; ++u) {
        if (u < len) {
            const double2 yv = ba_step(bs, cf, in[u]);
            emit(start + u, yv);
        }
    }
}

__global__ void __launch_bounds__(32) k_exact_edges_warp(const EdgeWarpArgs w) {
    const EdgeArgs& a = w.e;
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= a.n_jobs) return;                           // whole warps only
    const int car = a.jobs[j].x, mode = a.jobs[j].y;
    const float2* __restrict__ xc = a.x + (int64_t)car * a.pitch;
    const int L = a.L, q = a.q;
    const int64_t n = a.n;
    const EdgeRange rg = edge_range(mode, n, L, q, a.edge);
    const int m_lo = rg.m_lo, m_hi = rg.m_hi, o_lo = rg.o_lo, o_hi = rg.o_hi;
    const int64_t e_lo = rg.e_lo, e_hi = rg.e_hi;
    const int nf = (int)(e_hi - e_lo);
    double2* __restrict__ s1 = a.scr1 + (int64_t)j * a.w1;
    double2* __restrict__ sz = a.scrz + (int64_t)j * a.wz;
    double2* __restrict__ s2 = a.scr2 + (int64_t)j * (a.wz + 2 * EX_PAD2);
    const int var = mode == EX_LEFT ? 0 : 2;
    auto xat = [&](int64_t i) { const float2 v = __ldg(xc + i); return make_double2((double)v.x, (double)v.y); };

    // ---- stage 1 forward: sample s <-> e = e_lo + s <-> input index e - PAD1 (reflected in the pads) ----
    {
        const double2 edge_lo = xat(0), edge_hi = xat(n - 1);
        auto refl = [&](int s) {
            const int64_t i = e_lo + s - EX_PAD1;
            return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
        };
        sos_pass_warp<float2>(a.cf, nf, ex_oddext(xat, n, EX_PAD1, e_lo), w.m1 + (var + 0) * 5 * 64, lane,
            [&](int s) { return __ldg(xc + refl(min(s, nf - 1))); },
            [&](float2 r, int s) {
                const int64_t i = e_lo + s - EX_PAD1;
                const double2 v = make_double2((double)r.x, (double)r.y);
                if (i < 0) return make_double2(2.0 * edge_lo.x - v.x, 2.0 * edge_lo.y - v.y);
                if (i >= n) return make_double2(2.0 * edge_hi.x - v.x, 2.0 * edge_hi.y - v.y);
                return v;
            },
            [&](int s, double yr, double yi) { s1[s] = make_double2(yr, yi); });
    }
    __syncwarp();
    // ---- stage 1 backward: step s <-> e = e_hi - 1 - s; keep every q-th ----
    {
        const int nb = (int)(e_hi - rg.e_stop);
        const double w_nco = a.fo ? (2.0 * M_PI) * a.fo[car] : 0.0;
        sos_pass_warp<double2>(a.cf, nb, s1[nf - 1], w.m1 + (var + 1) * 5 * 64, lane,
            [&](int s) { return s1[nf - 1 - min(s, nb - 1)]; },
            [&](double2 r, int) { return r; },
            [&](int s, double yr, double yi) {
                const int64_t i = e_hi - 1 - s - EX_PAD1;  // >= 0
                if (i % q == 0) {
                    const int m = (int)(i / q);
                    if (i < n && m >= m_lo && m < m_hi) sz[m - m_lo] = edge_nco(make_double2(yr, yi), m, w_nco, a.fs_dec);
                }
            });
    }
    __syncwarp();
    // ---- stage 2: filtfilt(b, a) on z ----
    {
        auto zat = [&](int64_t mm) { return sz[mm - m_lo]; };
        auto z2 = [&](int64_t e) { return ex_oddext(zat, (int64_t)L, EX_PAD2, e); };
        const int64_t f_lo = rg.f_lo, f_hi = rg.f_hi;
        const int n2 = (int)(f_hi - f_lo);
        ba_pass_warp(a.cf, n2, z2(f_lo), w.m2 + (var + 0) * 5 * 16, lane,
            [&](int s) { return z2(f_lo + s); },
            [&](int s, double2 v) { s2[s] = v; });
        __syncwarp();
        const int nb2 = (int)(f_hi - rg.f_stop);
        float2* yc = a.y + (int64_t)car * a.y_pitch;
        ba_pass_warp(a.cf, nb2, s2[n2 - 1], w.m2 + (var + 1) * 5 * 16, lane,
            [&](int s) { return s2[n2 - 1 - s]; },
            [&](int s, double2 v) {
                const int64_t mm = f_hi - 1 - s - EX_PAD2;
                if (mm >= o_lo && mm < o_hi) yc[y_index(mm, a.y_sps, a.y_rows)] = make_float2((float)v.x, (float)v.y);
            });
    }
}

// ----------------------------------------------------------------------------------------------
// K_finalize: timing pick (processor.py:189-215) + soft symbols + differential slicer (:129-163)
// ----------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
constexpr int FIN_MAXPH = 32;
constexpr int FIN_B = 8;                  // symbols per thread and batch in k_finalize

struct FinArgs {
    const float2* y;         // [C][y_pitch] filtered samples at the decimated rate, layout y_index(n, sps, y_rows)
    int64_t y_pitch;
    int32_t y_rows;
    int32_t L;
    int32_t sps, step;       // samples per symbol, phase search step
    const double* partial;   // [C][n_seg][16] power sums of the bulk kernel (or null)
    int32_t n_seg;
    int32_t bulk_lo, bulk_hi;  // y range already covered by `partial` (empty if bulk_lo >= bulk_hi)
    uint8_t* dibits;         // [C][cap]
    int64_t cap;
    int32_t* n_dibits;       // [C]
    float2* symbols;         // [C][cap+1] or null
    int32_t* best_phase;     // [C] or null
    int32_t* phase_scratch;  // [C] always written (used by later kernels)
    uint8_t* match;          // [C][2*cap][2] or null: TS1/TS2 agreement counts, fused when cap <= FIN_DIB_SMEM
    int32_t* sync_pos;       // [C][max_pos] or null: sync positions of decode()'s cascade (decoder.py:845-856)
    int32_t max_pos;
    int32_t* n_sync;         // [C]
};

constexpr uint32_t TS1_BITS = 0x343A74u;   // 1101000011101001110100, first bit = MSB of 22 (decoder.py:196-197)
constexpr uint32_t TS2_BITS = 0x1E90DCu;   // 0111101001000011011100                      (decoder.py:198-199)
constexpr int FIN_DIB_SMEM = 12288;         // dibits of one carrier kept in shared memory for the fused correlator

// processor.py:152-161 on the differential product d = s1 * conj(s0) without the arctangent:
//   ph < -5pi/8 -> 3, < -3pi/8 -> 2, < 3pi/8 -> 0, < 5pi/8 -> 1, else 3,   ph = atan2(im, re) in (-pi, pi].
// With k = tan(3pi/8) the four rays are im = +-k re (re > 0: +-3pi/8) and im = -+k re (re < 0: +-5pi/8).
__device__ __forceinline__ uint8_t slice_dqpsk(double re, double im) {
    const double k = 2.414213562373095048801688724209698;   // 1 + sqrt(2)
    const double kr = k * re;
    if (re > 0.0) {
        if (im < -kr) return 2;              // ph < -3pi/8 (and > -pi/2)
        return im < kr ? 0 : 1;              // [-3pi/8, 3pi/8) -> 0, [3pi/8, pi/2) -> 1
    }
    // re <= 0: ph in [pi/2, pi] (im >= 0) or [-pi, -pi/2] (im < 0); -kr >= 0
    if (im > 0.0 || (im == 0.0 && re == 0.0)) {
        if (re == 0.0 && im == 0.0) return 0;   // atan2(0, 0) = 0
        return im > -kr ? 1 : 3;             // ph < 5pi/8  <=>  im > k |re|
    }
    if (im == 0.0) return 3;                 // ph = pi
    return im <= kr ? 2 : 3;                 // ph >= -5pi/8  <=>  -im >= k |re|  <=>  im <= k re
}

// ----------------------------------------------------------------------------------------------
// TetraDecoder.find_sync (core/decoder.py:171-295) and the threshold cascade of decode() (:845-856) for one
// carrier, by all FIN_THREADS threads of a CTA, from the MSB-first packed bit stream in shared memory.
// The reference walks every bit offset serially (jump +250 after a hit, max_corr over the visited offsets only,
// adaptive retry when nothing was found). Here the hits of a pass become a bit mask in parallel, one thread walks
// the mask (a handful of jumps), and the maximum over the visited offsets is a parallel reduction.
// ----------------------------------------------------------------------------------------------
struct SyncScratch {
    uint32_t mask[FIN_DIB_SMEM / 16 + 2];   // one bit per window start
    int n_pos, max_cnt;
};

__device__ __forceinline__ void ts_counts(const uint32_t* __restrict__ bits, int i, int& c1, int& c2) {
    const uint64_t two = ((uint64_t)bits[i >> 5] << 32) | bits[(i >> 5) + 1];
    const uint32_t win = (uint32_t)(two >> (64 - 22 - (i & 31))) & 0x3FFFFFu;
    c1 = 22 - __popc(win ^ TS1_BITS);
    c2 = 22 - __popc(win ^ TS2_BITS);
}
// smallest agreement count c with c / 22 >= threshold, in the reference's own float64 comparison (23: none)
__device__ __forceinline__ int sync_min_count(double thr) {
    int c = 0;
    while (c <= 22 && !((double)c / 22.0 >= thr)) ++c;
    return c;
}
// one thread: walk the hit mask like decoder.py:231-259 (record, jump 250) -> positions
__device__ inline int sync_walk(const uint32_t* mask, int nw, int32_t* pos, int max_pos) {
    int n = 0, i = 0;
    while (i < nw) {
        int wd = i >> 5;
        uint32_t m = mask[wd] & (0xFFFFFFFFu << (i & 31));
        const int n_words = (nw + 31) >> 5;
        while (m == 0 && ++wd < n_words) m = mask[wd];
        if (m == 0) break;
        const int p = (wd << 5) + __ffs(m) - 1;
        if (p >= nw) break;
        if (n < max_pos) pos[n] = p;
        ++n;
        i = p + 250;
    }
    return n;
}

// find_sync(bits, threshold) -> number of positions (written to pos[], global or shared), *max_corr
__device__ int block_find_sync(const uint32_t* __restrict__ bits, int nw, double thr, int32_t* pos, int max_pos,
                               SyncScratch& sc, double* max_corr) {
    const int tid = threadIdx.x;
    const int n_words = (nw + 31) >> 5;
    const int cmin = sync_min_count(thr);
    // pass 1: hits (TS1 is tried first, then TS2: decoder.py:237-259)
    for (int wd = tid; wd < n_words; wd += FIN_THREADS) {
        uint32_t m = 0;
        for (int b = 0; b < 32; ++b) {
            const int i = (wd << 5) + b;
            if (i < nw) {
                int c1, c2;
                ts_counts(bits, i, c1, c2);
                if (c1 >= cmin || c2 >= cmin) m |= 1u << b;
            }
        }
        sc.mask[wd] = m;
    }
    if (tid == 0) sc.max_cnt = 0;
    __syncthreads();
    if (tid == 0) sc.n_pos = sync_walk(sc.mask, nw, pos, max_pos);
    __syncthreads();
    int n = sc.n_pos;
    // max_corr over the VISITED offsets: everything except the 249 offsets skipped after each hit. At a visited
    // offset TS2's correlation only counts when TS1 did not already hit.
    const int n_known = min(n, max_pos);
    int best = 0;
    for (int i = tid; i < nw; i += FIN_THREADS) {
        int lo = 0, hi = n_known;                       // last position <= i
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pos[mid] <= i) lo = mid + 1; else hi = mid; }
        const bool skipped = lo > 0 && pos[lo - 1] < i && i < pos[lo - 1] + 250;
        if (!skipped) {
            int c1, c2;
            ts_counts(bits, i, c1, c2);
            best = max(best, c1 >= cmin ? c1 : max(c1, c2));
        }
    }
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) atomicMax(&sc.max_cnt, best);
    __syncthreads();
    const double mc = (double)sc.max_cnt / 22.0;
    *max_corr = mc;
    // adaptive retry inside find_sync (decoder.py:262-281)
    if (n == 0 && mc > 0.75 && mc >= (thr - 0.15)) {
        const double adaptive = fmax(0.75, mc - 0.02);
        if (adaptive < thr) {
            const int amin = sync_min_count(adaptive);
            __syncthreads();
            for (int wd = tid; wd < n_words; wd += FIN_THREADS) {
                uint32_t m = 0;
                for (int b = 0; b < 32; ++b) {
                    const int i = (wd << 5) + b;
                    if (i < nw) {
                        int c1, c2;
                        ts_counts(bits, i, c1, c2);
                        if (max(c1, c2) >= amin) m |= 1u << b;      // no offset was skipped: best_here = max of both
                    }
                }
                sc.mask[wd] = m;
            }
            __syncthreads();
            // accepted offsets block +-250 around them; scanning upwards that is the same jump-250 walk
            if (tid == 0) sc.n_pos = sync_walk(sc.mask, nw, pos, max_pos);
            __syncthreads();
            n = sc.n_pos;
        }
    }
    __syncthreads();
    return n;
}

// decode()'s cascade 0.90 -> 0.85 -> 0.80 -> adaptive (decoder.py:845-856)
__device__ int block_sync_cascade(const uint32_t* __restrict__ bits, int nd, int32_t* pos, int max_pos, SyncScratch& sc) {
    const int nw = 2 * nd - 22 + 1;
    if (nw <= 0) return 0;                              // decoder.py:226-228: fewer than 22 bits
    double mx = 0.0;
    int n = block_find_sync(bits, nw, 0.90, pos, max_pos, sc, &mx);
    if (n == 0) n = block_find_sync(bits, nw, 0.85, pos, max_pos, sc, &mx);
    if (n == 0) n = block_find_sync(bits, nw, 0.80, pos, max_pos, sc, &mx);
    if (n == 0 && mx >= 0.75) n = block_find_sync(bits, nw, fmax(0.75, mx - 0.02), pos, max_pos, sc, &mx);
    return n;
}

// dibits in shared memory -> MSB-first packed bits (decoder.py:140-169), one zero word behind
__device__ __forceinline__ void pack_dibits(const uint8_t* s_dib, int nd, uint32_t* s_bits) {
    const int n_words = (nd + 15) / 16 + 1;
    for (int j = threadIdx.x; j < n_words; j += blockDim.x) {
        uint32_t w = 0;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int idx = 16 * j + m;
            w = (w << 2) | (idx < nd ? (uint32_t)(s_dib[idx] & 3u) : 0u);
        }
        s_bits[j] = w;
    }
}

__global__ void __launch_bounds__(FIN_THREADS) k_finalize(const FinArgs a) {
    __shared__ double red[FIN_THREADS];
    __shared__ int s_best;
    __shared__ __align__(16) uint8_t s_dib[FIN_DIB_SMEM];
    __shared__ uint32_t s_bits[FIN_DIB_SMEM / 16 + 2];
    __shared__ SyncScratch s_sync;
    const int car = blockIdx.x, tid = threadIdx.x;
    const float2* __restrict__ y = a.y + (int64_t)car * a.y_pitch;
    const int L = a.L, sps = a.sps, step = a.step;
    const int nph = (sps + step - 1) / step;            // phases tried: 0, step, 2 step, ... (<= FIN_MAXPH)
    int best = 0;
    if (sps > 1) {
        // Power sum of phase ph over n = ph + sps*k, k < cnt = (L - ph) / sps. Thread (g, p) = (tid / nph, tid % nph)
        // owns the symbols k = g (mod G) of phase p; samples inside [bulk_lo, bulk_hi) were already summed by the
        // fused kernel and are skipped.
        const int G = FIN_THREADS / nph;
        const int g = tid / nph, p = tid % nph;
        const bool has_bulk = a.bulk_lo < a.bulk_hi;
        double acc = 0.0;
        if (g < G) {
            const int ph = p * step;
            const int cnt = (L - ph) / sps;
            // k < k_lo_end: below the bulk; k >= k_hi_beg: above it
            const int k_lo_end = has_bulk ? min(cnt, max(0, (a.bulk_lo - ph + sps - 1) / sps)) : cnt;
            const int k_hi_beg = has_bulk ? max(k_lo_end, (a.bulk_hi - ph + sps - 1) / sps) : cnt;
            for (int k = g; k < k_lo_end; k += G) {
                const float2 v = y[y_index(ph + sps * k, sps, a.y_rows)];
                acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            }
            for (int k = k_hi_beg + g; k < cnt; k += G) {
                const float2 v = y[y_index(ph + sps * k, sps, a.y_rows)];
                acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            }
        }
        red[tid] = acc;
        __syncthreads();
        if (tid == 0) {
            double best_pow = -1.0;
            for (int pp = 0; pp < nph; ++pp) {
                const int ph = pp * step;
                const int cnt = (L - ph) / sps;
                if (cnt <= 0) continue;
                double sum = 0.0;
                for (int gg = 0; gg < G; ++gg) sum += red[gg * nph + pp];
                if (a.partial && has_bulk)
                    for (int sg = 0; sg < a.n_seg; ++sg) sum += a.partial[((int64_t)car * a.n_seg + sg) * 16 + ph];
                const double mean = sum / (double)cnt;
                if (mean > best_pow) { best_pow = mean; best = ph; }
            }
            s_best = best;
        }
        __syncthreads();
        best = s_best;
    }
    const int stride = sps > 1 ? sps : 1;
    const int n_sym = sps > 1 ? max(0, (L - best) / sps) : L;
    const int nd = n_sym > 1 ? n_sym - 1 : 0;
    if (tid == 0) {
        a.n_dibits[car] = nd;
        if (a.best_phase) a.best_phase[car] = best;
        a.phase_scratch[car] = best;
    }
    uint8_t* dib = a.dibits + (int64_t)car * a.cap;
    float2* sym = a.symbols ? a.symbols + (int64_t)car * (a.cap + 1) : nullptr;
    const bool fuse = (a.match != nullptr || a.sync_pos != nullptr) && nd <= FIN_DIB_SMEM;
    // symbols k = tid + 256 j, FIN_B of them per batch with all loads of a batch issued before any use
    // symbol k is sample best + stride k: in the phase-major layout that is row `best`, contiguous in k
    const float2* ys = a.y_rows > 0 ? y + (int64_t)best * a.y_rows : y + best;
    const int64_t ks = a.y_rows > 0 ? 1 : stride;
    for (int k0 = tid; k0 < n_sym; k0 += FIN_B * FIN_THREADS) {
        float2 s1[FIN_B], s0[FIN_B];
#pragma unroll
        for (int j = 0; j < FIN_B; ++j) {
            const int k = min(k0 + j * FIN_THREADS, n_sym - 1);
            s1[j] = ys[ks * k];
            s0[j] = ys[ks * max(k - 1, 0)];
        }
#pragma unroll
        for (int j = 0; j < FIN_B; ++j) {
            const int k = k0 + j * FIN_THREADS;
            if (k < n_sym) {
                if (sym) sym[k] = s1[j];
                if (k >= 1) {
                    // diff = s1 * conj(s0); the products of two floats are exact in double
                    const double re = (double)s1[j].x * s0[j].x + (double)s1[j].y * s0[j].y;
                    const double im = (double)s1[j].y * s0[j].x - (double)s1[j].x * s0[j].y;
                    const uint8_t d = slice_dqpsk(re, im);
                    dib[k - 1] = d;
                    if (fuse) s_dib[k - 1] = d;
                }
            }
        }
    }
    if (!fuse) return;
    // ---- fused frame-sync front end (decoder.py:140-169 bit expansion, :237-240 agreement counts, :171-295 + :845-856) ----
    __syncthreads();
    pack_dibits(s_dib, nd, s_bits);
    __syncthreads();
    const int nw = 2 * nd - 22 + 1;                     // window starts (decoder.py:232)
    if (a.match) {
        uint8_t* out = a.match + (int64_t)car * a.cap * 4;
        for (int p = tid; 2 * p < nw; p += FIN_THREADS) {   // windows 2p and 2p+1 share their words
            const int i = 2 * p;
            const uint64_t two = ((uint64_t)s_bits[i >> 5] << 32) | s_bits[(i >> 5) + 1];
            const int sh = i & 31;                          // even, <= 30: 23 bits starting at sh fit in 64
            const uint32_t w0 = (uint32_t)(two >> (64 - 22 - sh)) & 0x3FFFFFu;
            const uint32_t w1 = (uint32_t)(two >> (64 - 23 - sh)) & 0x3FFFFFu;
            uchar4 o;
            o.x = (uint8_t)(22 - __popc(w0 ^ TS1_BITS));
            o.y = (uint8_t)(22 - __popc(w0 ^ TS2_BITS));
            o.z = (uint8_t)(22 - __popc(w1 ^ TS1_BITS));
            o.w = (uint8_t)(22 - __popc(w1 ^ TS2_BITS));
            if (i + 1 < nw) *reinterpret_cast<uchar4*>(out + 2 * (int64_t)i) = o;
            else { out[2 * (int64_t)i] = o.x; out[2 * (int64_t)i + 1] = o.y; }
        }
    }
    if (a.sync_pos) {
        const int n = block_sync_cascade(s_bits, nd, a.sync_pos + (int64_t)car * a.max_pos, a.max_pos, s_sync);
        if (tid == 0) a.n_sync[car] = n;
    }
}

// standalone: sync positions from dibit streams (the same device code; used when the streams come from elsewhere)
struct SyncPosArgs {
    const uint8_t* dibits; int64_t cap; const int32_t* n_dibits;
    int32_t* sync_pos; int32_t max_pos; int32_t* n_sync;
};
__global__ void __launch_bounds__(FIN_THREADS) k_sync_positions(const SyncPosArgs a) {
    __shared__ __align__(16) uint8_t s_dib[FIN_DIB_SMEM];
    __shared__ uint32_t s_bits[FIN_DIB_SMEM / 16 + 2];
    __shared__ SyncScratch s_sync;
    const int car = blockIdx.x;
    const int nd = min(a.n_dibits[car], FIN_DIB_SMEM);
    const uint8_t* dib = a.dibits + (int64_t)car * a.cap;
    for (int k = threadIdx.x; k < nd; k += FIN_THREADS) s_dib[k] = dib[k];
    __syncthreads();
    pack_dibits(s_dib, nd, s_bits);
    __syncthreads();
    const int n = block_sync_cascade(s_bits, nd, a.sync_pos + (int64_t)car * a.max_pos, a.max_pos, s_sync);
    if (threadIdx.x == 0) a.n_sync[car] = n;
}

// ----------------------------------------------------------------------------------------------
// k_parse_bursts: what TetraDecoder.decode does with each sync position up to the burst's CRC verdict
// (core/decoder.py:861-888 -> decode_frame :986-992 -> TetraProtocolParser.parse_burst, core/protocol.py:192-347):
// slot start = position - 216 bits, 255 symbols; burst type from the 22 bits at bit 255 (> 0.8 agreement with either
// sync pattern); data bits (normal burst: bits 0-107 + 122-229, sync burst: all 510); the reference's soft CRC-16-CCITT
// check (<= 2 differing CRC bits, forward or reversed payload). One warp per (carrier, position).
// info[car][slot] = (start_symbol or -1 when decode() drops the position, frame_number, burst_type, crc_ok)
// ----------------------------------------------------------------------------------------------
constexpr uint32_t SYNC_CONT_BITS = 0x343A74u;    // 1101000011101001110100 (protocol.py:162), first bit = MSB of 22
constexpr uint32_t SYNC_DISC_BITS = 0x0E90D3u;    // 0011101001000011010011 (protocol.py:163)

struct BurstArgs {
    const uint8_t* dibits; int64_t cap; const int32_t* n_dibits;
    const int32_t* sync_pos; int32_t max_pos; const int32_t* n_sync;
    int4* info;              // [C][max_pos]
};

// bit j of the 510-bit slot (MSB-first expansion of the symbols in shared memory)
__device__ __forceinline__ uint32_t burst_bit(const uint8_t* sym, int j) { return (sym[j >> 1] >> (1 - (j & 1))) & 1u; }

__global__ void __launch_bounds__(128) k_parse_bursts(const BurstArgs a) {
    __shared__ uint8_t s_sym[4][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + warp, car = blockIdx.y;
    if (slot >= a.max_pos) return;
    int4* out = a.info + (int64_t)car * a.max_pos + slot;
    const int nd = a.n_dibits[car];
    const int pos = slot < a.n_sync[car] ? a.sync_pos[(int64_t)car * a.max_pos + slot] : -1;
    const int start = pos - 216;
    const int s0 = start >> 1;                           // start >= 0 below
    if (pos < 0 || start < 0 || s0 + 255 > nd) {
        if (lane == 0) *out = make_int4(-1, 0, 0, 0);
        return;
    }
    uint8_t* sym = s_sym[warp];
    const uint8_t* dib = a.dibits + (int64_t)car * a.cap + s0;
    for (int k = lane; k < 255; k += 32) sym[k] = dib[k] & 3u;
    __syncwarp();
    // burst type (protocol.py:244-266)
    uint32_t win = 0;
    for (int j = 0; j < 22; ++j) win = (win << 1) | burst_bit(sym, 255 + j);
    const int m = max(22 - __popc(win ^ SYNC_CONT_BITS), 22 - __popc(win ^ SYNC_DISC_BITS));
    const bool is_sync = (double)m / 22.0 > 0.8;
    // data bit d of the burst (protocol.py:268-289)
    const int n_data = is_sync ? 510 : 216;
    auto data_bit = [&](int d) { return burst_bit(sym, is_sync ? d : (d < 108 ? d : d + 14)); };
    int ones = 0;
    for (int d = lane; d < n_data; d += 32) ones += data_bit(d);
    for (int o = 16; o; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
    // CRC-16-CCITT of the payload, MSB first, init 0xFFFF: lane 0 forward, lane 1 over the reversed payload
    const int n_pay = n_data - 16;
    uint32_t crc = 0xFFFFu;
    if (lane < 2) {
        for (int d = 0; d < n_pay; ++d) {
            crc ^= data_bit(lane == 0 ? d : n_pay - 1 - d) << 15;
            crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xFFFFu : (crc << 1) & 0xFFFFu;
        }
    }
    uint32_t recv = 0;
    for (int d = 0; d < 16; ++d) recv = (recv << 1) | data_bit(n_pay + d);
    const int err = __popc((crc ^ recv) & 0xFFFFu);
    const int err_fwd = __shfl_sync(0xffffffffu, err, 0), err_rev = __shfl_sync(0xffffffffu, err, 1);
    const bool crc_ok = ones != 0 && ones != n_data && (err_fwd <= 2 || err_rev <= 2);
    if (lane == 0) *out = make_int4(s0, start / 510, is_sync ? 5 : 2, crc_ok ? 1 : 0);
}

// ----------------------------------------------------------------------------------------------
// K_sync: dibits -> bits (decoder.py:140-169) and 22-bit TS1/TS2 agreement at every bit offset
// (decoder.py:237-240). One thread per window start.
// ----------------------------------------------------------------------------------------------
struct SyncArgs {
    const uint8_t* dibits; int64_t cap; const int32_t* n_dibits;
    uint8_t* match;          // [C][2*cap][2]
};

__global__ void __launch_bounds__(256) k_sync_match(const SyncArgs a) {
    const int car = blockIdx.y;
    const int nd = a.n_dibits[car];
    const int nw = 2 * nd - 22 + 1;
    const uint8_t* dib = a.dibits + (int64_t)car * a.cap;
    uint8_t* out = a.match + (int64_t)car * a.cap * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) {
        const int d0 = i >> 1;
        uint32_t bits = 0;                              // 24 bits: dibits d0 .. d0+11, first dibit highest
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int idx = d0 + k;
            const uint32_t v = idx < nd ? (dib[idx] & 3u) : 0u;
            bits = (bits << 2) | v;
        }
        const uint32_t win = (i & 1) ? (bits & 0x7FFFFEu) >> 1 : bits >> 2;   // 22 bits, first bit = MSB
        out[2 * (int64_t)i] = (uint8_t)(22 - __popc((win ^ TS1_BITS) & 0x3FFFFFu));
        out[2 * (int64_t)i + 1] = (uint8_t)(22 - __popc((win ^ TS2_BITS) & 0x3FFFFFu));
    }
}

}  // namespace tetra
