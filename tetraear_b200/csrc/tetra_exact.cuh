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

}  // namespace tetra
