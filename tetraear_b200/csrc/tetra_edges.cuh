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
// chunk length of a pass of n steps over 32 lanes: at least n / 32, stretched so that what follows the three pipeline-fill
// steps is a whole number of blocks (steps outside the blocks take the slow, branching path; the last lanes may run short or empty)
__host__ __device__ inline int exw_chunk_len(int n) {
    int lc = (n + 31) / 32;
    const int r = lc > 3 ? (lc - 3) % EXW_BLK : 0;
    return r ? lc + EXW_BLK - r : lc;
}
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
    const int lc = exw_chunk_len(n_steps);
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
    const float2* __restrict__ xc = a.x + (int64_t)car * a.pitch + (mode == EX_RIGHT ? a.right_shift : 0);
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
        // (a LEFT window never reaches past the block's last sample, a RIGHT one never before its first)
        const double2 zero2 = make_double2(0.0, 0.0);
        const double2 edge_lo = mode == EX_LEFT ? xat(0) : zero2, edge_hi = mode == EX_RIGHT ? xat(n - 1) : zero2;
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
        // decimation bookkeeping of this lane's chunk: the emitted samples come in order, input index i = q m + r counts down
        int m = 0, r = -1;
        sos_pass_warp<double2>(a.cf, nb, s1[nf - 1], w.m1 + (var + 1) * 5 * 64, lane,
            [&](int s) { return s1[nf - 1 - min(s, nb - 1)]; },
            [&](double2 v, int) { return v; },
            [&](int s, double yr, double yi) {
                if (r < 0) {                               // first sample of the chunk
                    const int64_t i = e_hi - 1 - s - EX_PAD1;  // >= 0
                    m = (int)(i / q); r = (int)(i - (int64_t)m * q);
                }
                if (r == 0) {
                    if ((int64_t)q * m < n && m >= m_lo && m < m_hi) sz[m - m_lo] = edge_nco(make_double2(yr, yi), m, w_nco, a.fs_dec);
                    r = q; --m;
                }
                --r;
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

}  // namespace tetra
