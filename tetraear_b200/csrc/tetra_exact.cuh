// Exact-recursion kernels: the reference's zero-phase IIR chain evaluated literally (fp64,
// SciPy's sosfiltfilt / filtfilt padding and initial-condition rules) on a window of a block.
//
// k_exact_chain here is (a) the complete path for short blocks, other sample rates and freq_offset
// beyond the fused kernel's range, (b) the engine of the public helper entry points (filter_signal, ...).
// The block ends of the fused path, where filtfilt is not shift-invariant, have their
// own kernels in tetra_edges.cuh; the shared pieces (padding, initial conditions, y layout) live here.
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
constexpr int K_EDGE = 168;   // outputs per block end owned by the exact path (= K1_EDGE)
constexpr int K_EDGE_MAX_S2 = K_EDGE + EX_T2 + 2 * EX_PAD2;   // bound on the stage-2 window of an edge job (E + T2 + PAD2)

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
    int64_t x_right_shift;   // see EdgeArgs::right_shift (edge kernels only)
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
// k_exact_block: the same literal recursions over a WHOLE block, one CTA per carrier, every pass cut into chunks that the
// CTA's threads run in parallel. A pass is a linear recursion: from a zero state a thread first runs its chunk to learn
// what the chunk contributes to the state (z_c); the true start state of chunk c follows from
//        S_c = z_{c-1} + M z_{c-2} + ... + M^{T-1} z_{c-T} + M^T S_{c-T},        M = zero-input transition over one chunk,
// where the host picks the chunk length so that ||M|| <= 3e-3 and T = 8 terms leave nothing (< 1e-20; chunks c <= T use S_0
// exactly); then the thread runs its chunk again from S_c and emits. Twice the arithmetic of the serial evaluation, a few
// hundred times its speed: this is the path of the sample rates the fused kernel does not cover (signal/capture.py:83-87:
// q = 7, 8, 12, 13 ...), of freq_offsets beyond its range and of the complex128 helper methods on long inputs.
// ----------------------------------------------------------------------------------------------
constexpr int EXB_THREADS = 512;
constexpr int EXB_T = 8;                  // chunks folded into a start state
constexpr int EXB_G = 4;                  // inputs fetched ahead of the (serially dependent) recursion steps

struct ExactBlockArgs {
    ExactArgs e;                          // jobs[].x = carrier (mode EX_FULL); scr1 / scrz / scr2 are [job][...] here (job-major)
    int64_t s1_stride, sz_stride, s2_stride;   // elements per job in scr1 / scrz / scr2
    int32_t lc1, lc2;                     // chunk lengths of the stage-1 / stage-2 passes
    double m1[64];                        // stage 1: zero-input transition of the 8 cascade states over lc1 steps (row-major)
    double m2[16];                        // stage 2: the same for the 4 lfilter states over lc2 steps
};

template <int NS> struct ExbState { double re[NS], im[NS]; };

__device__ __forceinline__ void exb_from(SosState& st, const ExbState<8>& v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { st.z[k][0][0] = v.re[2 * k]; st.z[k][0][1] = v.im[2 * k]; st.z[k][1][0] = v.re[2 * k + 1]; st.z[k][1][1] = v.im[2 * k + 1]; }
}
__device__ __forceinline__ void exb_to(const SosState& st, ExbState<8>& v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { v.re[2 * k] = st.z[k][0][0]; v.im[2 * k] = st.z[k][0][1]; v.re[2 * k + 1] = st.z[k][1][0]; v.im[2 * k + 1] = st.z[k][1][1]; }
}
__device__ __forceinline__ void exb_from(BaState& st, const ExbState<4>& v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { st.z[k][0] = v.re[k]; st.z[k][1] = v.im[k]; }
}
__device__ __forceinline__ void exb_to(const BaState& st, ExbState<4>& v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { v.re[k] = st.z[k][0]; v.im[k] = st.z[k][1]; }
}
__device__ __forceinline__ double2 exb_step(SosState& st, const ExactCoef& c, double2 v) { return sos_step(st, c, v); }
__device__ __forceinline__ double2 exb_step(BaState& st, const ExactCoef& c, double2 v) { return ba_step(st, c, v); }

// steps e0 .. e1-1 of a recursion, the next group's inputs in flight while the current group's dependent steps execute
template <class ST, class IN, class OUT>
__device__ __forceinline__ void exb_run(ST& st, const ExactCoef& cf, int64_t e0, int64_t e1, IN&& in, OUT&& out) {
    if (e0 >= e1) return;
    double2 ga[EXB_G], gb[EXB_G];
    auto fetch = [&](double2* g, int64_t e) {
#pragma unroll
        for (int u = 0; u < EXB_G; ++u) g[u] = in(min(e + u, e1 - 1));
    };
    auto run = [&](const double2* g, int64_t e) {
#pragma unroll
        for (int u = 0; u < EXB_G; ++u)
            if (e + u < e1) out(e + u, exb_step(st, cf, g[u]));
    };
    fetch(ga, e0);
    for (int64_t e = e0; e < e1; e += 2 * EXB_G) {
        if (e + EXB_G < e1) fetch(gb, e + EXB_G);
        run(ga, e);
        if (e + EXB_G < e1) {
            if (e + 2 * EXB_G < e1) fetch(ga, e + 2 * EXB_G);
            run(gb, e + EXB_G);
        }
    }
}

// one pass of length T: in(e), e = 0 .. T-1 in processing order; init = state before e = 0; out(e, value)
template <class ST, int NS, class IN, class OUT>
__device__ __forceinline__ void exb_pass(const ExactCoef& cf, const double* __restrict__ m, int64_t T, int lc, const ExbState<NS>& init,
                                         ExbState<NS>* zst, IN&& in, OUT&& out) {
    const int c = threadIdx.x;
    const int nc = (int)((T + lc - 1) / lc);              // <= EXB_THREADS by the host's choice of lc
    const int64_t e0 = (int64_t)c * lc, e1 = min(T, e0 + lc);
    ST st;
    ExbState<NS> s;
    if (c < nc) {
#pragma unroll
        for (int k = 0; k < NS; ++k) { s.re[k] = 0.0; s.im[k] = 0.0; }
        exb_from(st, s);
        exb_run(st, cf, e0, e1, in, [](int64_t, double2) {});
        exb_to(st, s);
        zst[c] = s;
    }
    __syncthreads();
    if (c < nc) {
        // S_c by Horner from the oldest chunk that still matters
        const int first = max(0, c - EXB_T);
        if (first == 0) s = init;
        else {
#pragma unroll
            for (int k = 0; k < NS; ++k) { s.re[k] = 0.0; s.im[k] = 0.0; }
        }
        for (int kc = first; kc < c; ++kc) {
            ExbState<NS> t;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                double ar = zst[kc].re[i], ai = zst[kc].im[i];
#pragma unroll
                for (int j = 0; j < NS; ++j) { ar += m[i * NS + j] * s.re[j]; ai += m[i * NS + j] * s.im[j]; }
                t.re[i] = ar; t.im[i] = ai;
            }
            s = t;
        }
        exb_from(st, s);
        exb_run(st, cf, e0, e1, in, out);
    }
    __syncthreads();
}

constexpr int EXB_SMEM = EXB_THREADS * (int)sizeof(ExbState<8>);     // 64 KB (dynamic)
__global__ void __launch_bounds__(EXB_THREADS) k_exact_block(const ExactBlockArgs b) {
    extern __shared__ __align__(16) unsigned char exb_smem[];
    ExbState<8>* zst = reinterpret_cast<ExbState<8>*>(exb_smem);
    const ExactArgs& a = b.e;
    const int j = blockIdx.x;
    const int car = a.jobs[j].x;
    const int64_t xb = (int64_t)car * a.pitch;
    const int L = a.L, q = a.q;
    const int64_t n = a.n;
    double2* __restrict__ scr1 = a.scr1 + (int64_t)j * b.s1_stride;
    double2* __restrict__ scrz = a.scrz + (int64_t)j * b.sz_stride;
    double2* __restrict__ scr2 = a.scr2 + (int64_t)j * b.s2_stride;
    auto xat = [&](int64_t i) { return ex_load(a, xb, i); };
    const double w_nco = a.fo ? (2.0 * M_PI) * a.fo[car] : 0.0;
    auto nco = [&](double2 v, int64_t m) {
        if (w_nco == 0.0) return v;
        const double t = (double)m / a.fs_dec;
        double sn, cs;
        sincos(-(w_nco * t), &sn, &cs);
        return make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    };
    // ---------------- stage 1: sosfiltfilt + [::q] ----------------
    if (a.has_s1) {
        const int64_t T = n + 2 * EX_PAD1;
        ExbState<8> init;
        {
            SosState st;
            sos_init(st, a.cf, ex_oddext(xat, n, EX_PAD1, 0));
            exb_to(st, init);
        }
        exb_pass<SosState, 8>(a.cf, b.m1, T, b.lc1, init, zst,
                              [&](int64_t e) { return (e >= EX_PAD1 && e < EX_PAD1 + n) ? xat(e - EX_PAD1) : ex_oddext(xat, n, EX_PAD1, e); },
                              [&](int64_t e, double2 v) { scr1[e] = v; });
        {
            SosState st;
            sos_init(st, a.cf, scr1[T - 1]);
            exb_to(st, init);
        }
        exb_pass<SosState, 8>(a.cf, b.m1, T, b.lc1, init, zst,
                              [&](int64_t e) { return scr1[T - 1 - e]; },
                              [&](int64_t e, double2 v) {
                                  const int64_t i = T - 1 - e - EX_PAD1;             // input index of this output (< 2^31)
                                  if (i >= 0 && i < n) {
                                      const int ii = (int)i, m = ii / q;
                                      if (m * q == ii) scrz[m] = nco(v, m);
                                  }
                              });
    } else {
        for (int64_t m = threadIdx.x; m < L; m += EXB_THREADS) scrz[m] = nco(xat(m), m);
        __syncthreads();
    }
    // ---------------- stage 2: filtfilt(b, a) ----------------
    auto put = [&](int64_t m, double2 v) {
        if (a.y32) a.y32[(int64_t)car * a.y_pitch + y_index(m, a.y_sps, a.y_rows)] = make_float2((float)v.x, (float)v.y);
        else a.y64[(int64_t)car * a.y_pitch + m] = v;
    };
    if (a.has_s2) {
        const int64_t T = (int64_t)L + 2 * EX_PAD2;
        auto zat = [&](int64_t m) { return scrz[m]; };
        ExbState<4>* zst4 = reinterpret_cast<ExbState<4>*>(zst);
        ExbState<4> init;
        {
            BaState st;
            ba_init(st, a.cf, ex_oddext(zat, (int64_t)L, EX_PAD2, 0));
            exb_to(st, init);
        }
        exb_pass<BaState, 4>(a.cf, b.m2, T, b.lc2, init, zst4,
                             [&](int64_t e) { return ex_oddext(zat, (int64_t)L, EX_PAD2, e); },
                             [&](int64_t e, double2 v) { scr2[e] = v; });
        {
            BaState st;
            ba_init(st, a.cf, scr2[T - 1]);
            exb_to(st, init);
        }
        exb_pass<BaState, 4>(a.cf, b.m2, T, b.lc2, init, zst4,
                             [&](int64_t e) { return scr2[T - 1 - e]; },
                             [&](int64_t e, double2 v) {
                                 const int64_t m = T - 1 - e - EX_PAD2;
                                 if (m >= 0 && m < L) put(m, v);
                             });
    } else {
        for (int64_t m = threadIdx.x; m < L; m += EXB_THREADS) put(m, scrz[m]);
    }
}

}  // namespace tetra
