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
constexpr int EX_T1 = 1280;   // warm-up of an approximate start/end, stage 1 (pole radius 0.9821 -> 8e-11)
constexpr int EX_T2 = 160;    // same for stage 2 (pole radius 0.8837 -> 3e-9)
constexpr int EX_U = 8;       // recursion steps per load group

enum { EX_FULL = 0, EX_LEFT = 1, EX_RIGHT = 2 };

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
    else if (mode == EX_RIGHT) { o_lo = max(0, L - E); m_lo = max(0, L - E - 2 * EX_T2); }

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
        if (a.y32) a.y32[(int64_t)car * a.y_pitch + m] = make_float2((float)v.x, (float)v.y);
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
// K_finalize: timing pick (processor.py:189-215) + soft symbols + differential slicer (:129-163)
// ----------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
constexpr int FIN_MAXPH = 32;

struct FinArgs {
    const float2* y;         // [C][y_pitch] filtered samples at the decimated rate
    int64_t y_pitch;
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
};

__global__ void __launch_bounds__(FIN_THREADS) k_finalize(const FinArgs a) {
    __shared__ double red[FIN_MAXPH][FIN_THREADS / 32];
    __shared__ int s_best;
    const int car = blockIdx.x, tid = threadIdx.x;
    const float2* y = a.y + (int64_t)car * a.y_pitch;
    const int L = a.L, sps = a.sps, step = a.step;
    const int nph = (sps + step - 1) / step;            // phases tried: 0, step, 2 step, ...
    int best = 0;
    if (sps > 1) {
        // power sums over n = ph + sps*k, k < (L - ph) / sps  <=>  n + sps <= L
        double acc[FIN_MAXPH];
#pragma unroll
        for (int p = 0; p < FIN_MAXPH; ++p) acc[p] = 0.0;
        const int n_end = L - sps;                       // last admissible n (inclusive)
        for (int n = tid; n <= n_end; n += FIN_THREADS) {
            if (n >= a.bulk_lo && n < a.bulk_hi) continue;
            const int ph = n % sps;
            if (ph % step) continue;
            const float2 v = y[n];
            const double pw = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            const int slot = ph / step;
#pragma unroll
            for (int p = 0; p < FIN_MAXPH; ++p) if (p == slot) acc[p] += pw;
        }
        for (int p = 0; p < nph; ++p) {
            double v = acc[p];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) red[p][tid >> 5] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double best_pow = -1.0;
            for (int p = 0; p < nph; ++p) {
                const int ph = p * step;
                const int cnt = (L - ph) / sps;
                if (cnt <= 0) continue;
                double sum = 0.0;
                for (int w = 0; w < FIN_THREADS / 32; ++w) sum += red[p][w];
                if (a.partial && a.bulk_lo < a.bulk_hi)
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
    if (tid == 0) {
        a.n_dibits[car] = n_sym > 1 ? n_sym - 1 : 0;
        if (a.best_phase) a.best_phase[car] = best;
        a.phase_scratch[car] = best;
    }
    uint8_t* dib = a.dibits + (int64_t)car * a.cap;
    float2* sym = a.symbols ? a.symbols + (int64_t)car * (a.cap + 1) : nullptr;
    const double T3 = 3.0 * M_PI / 8.0, T5 = 5.0 * M_PI / 8.0;
    for (int k = tid; k < n_sym; k += FIN_THREADS) {
        const float2 s1 = y[best + (int64_t)stride * k];
        if (sym) sym[k] = s1;
        if (k >= 1) {
            const float2 s0 = y[best + (int64_t)stride * (k - 1)];
            // diff = s1 * conj(s0)
            const double re = (double)s1.x * s0.x + (double)s1.y * s0.y;
            const double im = (double)s1.y * s0.x - (double)s1.x * s0.y;
            const double ph = atan2(im, re);
            uint8_t d;
            if (ph < -T5) d = 3; else if (ph < -T3) d = 2; else if (ph < T3) d = 0; else if (ph < T5) d = 1; else d = 3;
            dib[k - 1] = d;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// K_sync: dibits -> bits (decoder.py:140-169) and 22-bit TS1/TS2 agreement at every bit offset
// (decoder.py:237-240). One thread per window start.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t TS1_BITS = 0x343A74u;   // 1101000011101001110100, first bit = MSB of 22
constexpr uint32_t TS2_BITS = 0x1E90DCu;   // 0111101001000011011100

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
