// KC: block-end corrections of the fused path (k_edge_states + k_edge_recursions).
//
// The reference chain  decimate(x, 10) -> [NCO] -> filtfilt(butter4)  (tetraear/signal/processor.py:245-264) is
// shift-invariant except next to the two ends of a block, where SciPy's sosfiltfilt / filtfilt use an odd extension
// (27 / 15 samples), steady-state initial conditions (zi * first sample) and hold the forward pass's last output for
// the backward pass. The fused kernel computes the shift-invariant part: its cascade applied to the block extended by
// ZEROS. This kernel computes, in float64, the difference
//        D[m] = reference(x)[m] - cascade(x zero-extended)[m]        for the K_EDGE outputs next to each end,
// which is linear in x and is driven by a handful of IIR filter states at the block end. Those states are dot products
// of the block's first / last ~1200 samples with fixed weight tables (edge_tables_generated.h, from tools/edge_model.py,
// where the derivation and a float64 model of exactly these steps live): k_edge_states, one CTA per carrier, its four warps
// sharing the seven table passes. A few hundred literal order-4 recursion steps at 240 kS/s then give D: k_edge_recursions,
// one thread per (carrier, end), the lanes of a warp walking 32 carriers in lockstep. Without a freq_offset that map from
// the states to D is one fixed real matrix per block end (it depends on (n - 1) mod 10 only): the host builds it once per
// geometry by running k_edge_recursions on unit states, and k_edge_apply evaluates D = M s in parallel -- 168 threads per
// (carrier group, end) instead of a 900-step serial chain. All of it depends on the input only, so it runs beside the fused
// kernel; the finalize kernel adds D to the fused kernel's output where it reads the block ends.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tetra_exact.cuh"
#include "edge_tables_generated.h"

namespace tetra {

constexpr int KC_THREADS = 128;           // k_edge_states: one CTA per carrier, the table passes spread over its four warps
constexpr int KC2_THREADS = 64;           // k_edge_recursions: one thread per (carrier, block end)
// per carrier (complex128): what the right end's correction depends on, then the same for the left end
constexpr int KS_SC = 0;                  // [8]  causal cascade state after x[n-1]
constexpr int KS_PTR = 8;                 // [16] decimator outputs L-1-t of the zero-extended stream
constexpr int KS_S2 = 24;                 // [4]  causal Butterworth state of the zero-extended stream at the right end
constexpr int KS_XR = 28;                 // [28] x[n-1-j], j = 0 .. PAD1 (stage 1's odd extension)
constexpr int KS_NR = 56;
constexpr int KS_SAC = 56;                // [8]  backward-pass state at position 0
constexpr int KS_PTL = 64;                // [16] decimator outputs t of the zero-extended stream
constexpr int KS_XL = 80;                 // [28] x[j], j = 0 .. PAD1
constexpr int KS_NL = 52;
constexpr int KC_NSTATE = KS_NR + KS_NL;
constexpr int KA_THREADS = 192, KA_CPB = 8;   // k_edge_apply: K_EDGE outputs x 8 carriers per CTA
static_assert(ET_NPTS == 16 && EX_PAD1 + 1 == 28 && KA_THREADS >= K_EDGE, "state layout");
static_assert(ET_NPTS == EX_PAD2 + 1, "stage 2's odd extension needs PAD2 + 1 decimator outputs");
static_assert(ET_TD <= K_EDGE && ET_NRING >= 10 * ET_TD + 10, "ringing tables too short");
static_assert(ET_G >= 10 * ET_NPTS, "pointwise windows start inside the block");

struct EdgeTables {          // device copies of edge_tables_generated.h
    const double* g1;        // [2 G + 1]   zero-phase Chebyshev impulse response, g1[u + G]
    const double* wc;        // [8][NC]     causal cascade state after a unit sample d steps back
    const double* wac;       // [8][NAC]    backward-pass state at position 0 for a unit sample at i
    const double* ringc;     // [NRING][8]  zero-input output p steps after unit state k
    const double* ring;      // [NRING][8]  backward-pass output at offset p into the ringing of unit state k
    const double* u;         // [8][8]      backward-pass state after the whole ringing of unit state k
    const double* u2;        // [4][4]      the same for the Butterworth stage (lfilter zi layout)
    const double* w0;        // [4][NW0]    freq_offset = 0: weights of the Butterworth state at the right end, W0[k][u + 9]
    const double* bp;        // [4][2]      poles of the Butterworth stage
    const double* bc;        // [4][2]      its input vector in modal coordinates
    const double* bv;        // [4][4][2]   modal coordinates -> lfilter state
    const double* g1p;       // g1 between two runs of ET_G1_PAD zeros: g1p[i] = g1[i] for 0 <= i <= 2 G, 0 for -PAD <= i < 0 and
                             // 2 G < i <= 2 G + PAD (the pointwise windows index it without a range check)
};
constexpr int ET_G1_PAD = 96;
static_assert(ET_G1_PAD >= 10 * 7 + 1 && ET_G - 9 - 80 * (ET_NPTS / 8 - 1) - 70 >= -ET_G1_PAD && ET_G + 80 * (ET_NPTS / 8 - 1) + 70 <= 2 * ET_G + ET_G1_PAD,
              "pointwise window indices leave the padded table");

struct EdgeCorrArgs {
    const float2* x;         // [C][pitch] complex64 ...
    const uint8_t* x8;       // ... or [C][pitch][2] unsigned bytes (RTL-SDR), converted like pyrtlsdr: b / 127.5 - 1
    int64_t pitch, n;        // pitch 0: every "carrier" is a channel of one shared capture (chan != null)
    int32_t L;               // ceil(n / 10)
    const double* fo;        // [C] NCO between the two filters (Hz at fs_dec), or null
    const double* chan;      // [C] frequency_shift of the shared capture before everything (Hz at fs), or null
    double fs, fs_dec;
    ExactCoef cf;
    EdgeTables t;
    int32_t n_carriers;
    double2* states;         // [C][KC_NSTATE] scratch between the two kernels
    float2* d;               // [C][2][K_EDGE]: D_left[m] (m = 0..), D_right[t] (output L-1-t)
    double* m_out;           // k_edge_recursions on unit states: real parts [C][2][K_EDGE] in float64 instead of d (the matrix M)
    const double* m;         // k_edge_apply: M[k][end][K_EDGE], k = state index
};

struct dcx { double x, y; };
__device__ __forceinline__ dcx cmul(dcx a, dcx b) { return dcx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ dcx cexp_turns(double turns) {          // exp(-j 2 pi turns), phase reduced first
    dcx r;
    sincospi(-2.0 * (turns - rint(turns)), &r.y, &r.x);
    return r;
}

// sample i (0 <= i < n) of this carrier as the reference sees it (complex128)
template <int IN>
__device__ __forceinline__ dcx kc_load(const EdgeCorrArgs& a, int car, double chan_hz, int64_t i) {
    if (IN == 1) {
        const uchar2 p = *reinterpret_cast<const uchar2*>(a.x8 + 2 * ((int64_t)car * a.pitch + i));
        return dcx{(double)p.x / 127.5 - 1.0, (double)p.y / 127.5 - 1.0};
    }
    const float2 v = __ldg(a.x + (int64_t)car * a.pitch + i);
    dcx r{(double)v.x, (double)v.y};
    if (IN == 2) r = cmul(r, cexp_turns(chan_hz * ((double)i / a.fs)));     // processor.py:97-100 at the full rate
    return r;
}

struct KcBa { dcx z[4]; };
__device__ __forceinline__ dcx kc_ba_step(KcBa& s, const ExactCoef& c, dcx v) {     // scipy lfilter, order 4
    const dcx y{c.b[0] * v.x + s.z[0].x, c.b[0] * v.y + s.z[0].y};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s.z[k].x = c.b[k + 1] * v.x - c.a[k + 1] * y.x + s.z[k + 1].x;
        s.z[k].y = c.b[k + 1] * v.y - c.a[k + 1] * y.y + s.z[k + 1].y;
    }
    s.z[3].x = c.b[4] * v.x - c.a[4] * y.x;
    s.z[3].y = c.b[4] * v.y - c.a[4] * y.y;
    return y;
}
__device__ __forceinline__ dcx kc_dot8(const double* __restrict__ row, const dcx* v) {
    const double4 r0 = *reinterpret_cast<const double4*>(row), r1 = *reinterpret_cast<const double4*>(row + 4);
    dcx r;
    r.x = r0.x * v[0].x + r0.y * v[1].x + r0.z * v[2].x + r0.w * v[3].x + r1.x * v[4].x + r1.y * v[5].x + r1.z * v[6].x + r1.w * v[7].x;
    r.y = r0.x * v[0].y + r0.y * v[1].y + r0.z * v[2].y + r0.w * v[3].y + r1.x * v[4].y + r1.y * v[5].y + r1.z * v[6].y + r1.w * v[7].y;
    return r;
}
__device__ __forceinline__ void kc_sos_from(SosState& ss, const dcx* v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ss.z[k][0][0] = v[2 * k].x; ss.z[k][0][1] = v[2 * k].y;
        ss.z[k][1][0] = v[2 * k + 1].x; ss.z[k][1][1] = v[2 * k + 1].y;
    }
}
// sums of NA complex accumulators over the lanes of a warp -> dst[0..NA) (global), written by lane 0
template <int NA>
__device__ __forceinline__ void kc_warp_sum(dcx (&acc)[NA], double2* dst, int lane) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
            acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
        }
        if (lane == 0) dst[k] = make_double2(acc[k].x, acc[k].y);
    }
}

// ----------------------------------------------------------------------------------------------
// k_edge_states: the functionals. One warp per carrier; lane l takes samples l, l + 32, ... of each end.
// ----------------------------------------------------------------------------------------------
// FO: the batch has freq_offsets (a.fo != null). A compile-time switch because the modal recursion of that case keeps 80
// float64 registers alive; without it the common kernel fits more CTAs per SM.
template <int IN, bool FO>   // IN 0: complex64 rows, 1: uint8 rows, 2: channels of one shared complex64 capture
__global__ void __launch_bounds__(KC_THREADS, FO ? 3 : 5) k_edge_states(const EdgeCorrArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Without freq_offsets: one CTA per carrier, its four warps share the seven table passes (role = warp). With them the
    // Butterworth-state pass is a serial chain ~18 table passes long that no split of the others can balance: one warp per
    // carrier does everything (role -1), four carriers per CTA.
    constexpr bool wpc = FO;
    const int car = wpc ? blockIdx.x * (KC_THREADS / 32) + warp : blockIdx.x;
    if (car >= a.n_carriers) return;                          // whole warps
    const int role = wpc ? -1 : warp;
    auto has = [&](int r) { return role < 0 || role == r; };
    const int64_t n = a.n;
    const int L = a.L;
    const int k0 = (int)((n - 1) % 10);
    const double chan_hz = (IN == 2) ? a.chan[car] : 0.0;
    const double fo = FO ? a.fo[car] : 0.0;
    auto xr = [&](int d) { return kc_load<IN>(a, car, chan_hz, n - 1 - d); };      // d samples before the last one
    auto xl = [&](int i) { return kc_load<IN>(a, car, chan_hz, i); };
    double2* out = a.states + (int64_t)car * KC_NSTATE;
    const double* __restrict__ g1 = a.t.g1;
    const double* __restrict__ g1p = a.t.g1p;
    constexpr int UN = 4;                                      // samples in flight per lane

    dcx acc[8];
    auto clear = [&]() {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = dcx{0.0, 0.0};
    };
    // warp 0: cascade state + Butterworth state; warp 1: backward-pass state + the raw end samples; warps 2, 3: pointwise outputs
    if (has(1)) {
        if (lane <= EX_PAD1) {
            const dcx r = xr(lane), l = xl(lane);
            out[KS_XR + lane] = make_double2(r.x, r.y);
            out[KS_XL + lane] = make_double2(l.x, l.y);
        }
    }
    // ---- causal cascade state after x[n-1] ----
    if (has(0)) {
        clear();
        const double* __restrict__ wc = a.t.wc;
        for (int d0 = lane; d0 < ET_NC; d0 += 32 * UN) {
            dcx v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) v[q] = d0 + 32 * q < ET_NC ? xr(d0 + 32 * q) : dcx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const int d = min(d0 + 32 * q, ET_NC - 1);
#pragma unroll
                for (int k = 0; k < 8; ++k) { const double w = wc[k * ET_NC + d]; acc[k].x += w * v[q].x; acc[k].y += w * v[q].y; }
            }
        }
        kc_warp_sum(acc, out + KS_SC, lane);
    }
    // ---- backward-pass state at position 0 of the zero-extended stream ----
    if (has(1)) {
        clear();
        const double* __restrict__ wac = a.t.wac;
        for (int d0 = lane; d0 < ET_NAC; d0 += 32 * UN) {
            dcx v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) v[q] = d0 + 32 * q < ET_NAC ? xl(d0 + 32 * q) : dcx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const int d = min(d0 + 32 * q, ET_NAC - 1);
#pragma unroll
                for (int k = 0; k < 8; ++k) { const double w = wac[k * ET_NAC + d]; acc[k].x += w * v[q].x; acc[k].y += w * v[q].y; }
            }
        }
        kc_warp_sum(acc, out + KS_SAC, lane);
    }
    // ---- pointwise decimator outputs of the zero-extended stream: at output L-1-t, sum_d g1[d - k0 - 10 t] x[n-1-d],
    //      and at output t, sum_i g1[10 t - i] x[i]   (t = 8 h + k) ----
    for (int h = 0; h < ET_NPTS / 8; ++h) {
        if (!has(2 + h)) continue;
        clear();
        const int off_r = ET_G - k0 - 80 * h, cnt_r = k0 + 10 * (8 * h + 7) + ET_G + 1;
        for (int d0 = lane; d0 < cnt_r; d0 += 32 * UN) {
            dcx v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) v[q] = d0 + 32 * q < cnt_r ? xr(d0 + 32 * q) : dcx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                // lanes beyond the window hold v = 0: any finite weight will do, the index only has to stay inside the table;
                // inside the window off_r - 70 <= idx <= 2 G + 70, where the padded table reads g1 or 0
                const double* __restrict__ gp = g1p + (min(d0 + 32 * q, cnt_r - 1) + off_r);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double w = gp[-10 * k];
                    acc[k].x += w * v[q].x; acc[k].y += w * v[q].y;
                }
            }
        }
        kc_warp_sum(acc, out + KS_PTR + 8 * h, lane);
        clear();
        const int off_l = ET_G + 80 * h, cnt_l = 10 * (8 * h + 7) + ET_G + 1;
        for (int d0 = lane; d0 < cnt_l; d0 += 32 * UN) {
            dcx v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) v[q] = d0 + 32 * q < cnt_l ? xl(d0 + 32 * q) : dcx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const double* __restrict__ gp = g1p + (off_l - min(d0 + 32 * q, cnt_l - 1));     // -70 <= idx <= G + 150
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double w = gp[10 * k];
                    acc[k].x += w * v[q].x; acc[k].y += w * v[q].y;
                }
            }
        }
        kc_warp_sum(acc, out + KS_PTL + 8 * h, lane);
    }
    // ---- causal Butterworth state of the zero-extended stream at the right end ----
    // s(L) = e^{-jW(L-1)} sum_d x[n-1-d] Wv(d - k0),  Wv(u) = Bv g1[u] + e^{jW} A Wv(u - 10)  (tools/edge_model.py).
    // Without a freq_offset the weights are real and fixed: one more table pass.
    if (!has(0)) return;
    if (!FO || fo == 0.0) {
        dcx a4[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        const double* __restrict__ w0 = a.t.w0 + 9 - k0;      // w0[k * NW0 + d] = W0[k][d - k0 + 9]
        const int cnt = ET_NW0 - 9 + k0;
        for (int d0 = lane; d0 < cnt; d0 += 32 * UN) {
            dcx v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) v[q] = d0 + 32 * q < cnt ? xr(d0 + 32 * q) : dcx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const int d = min(d0 + 32 * q, cnt - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const double w = w0[k * ET_NW0 + d]; a4[k].x += w * v[q].x; a4[k].y += w * v[q].y; }
            }
        }
        kc_warp_sum(a4, out + KS_S2, lane);
        return;
    }
    if (!FO) return;
    // With one: in modal coordinates (A = V diag(p) V^-1, lambda_r = p_r e^{jW}):
    //   s(L) = V sigma,  sigma_r = c_r e^{-jW(L-1)} sum_d x[n-1-d] Om_r(d - k0),  Om_r(u) = g1[u] + lambda_r Om_r(u - 10).
    // Ten chains (u mod 10) of NSTEP steps, each cut into three segments: a lane runs its segment from zero, then the
    // segments' carries are chained with shuffles.
    {
        constexpr int LO = -(ET_G + 10) - ((-(ET_G + 10)) % 10 + 10) % 10;     // multiple of 10 at or below -(G + 10)
        constexpr int NU = ET_G + 10 * ET_T2 + 10;
        constexpr int NSTEP = (NU - LO) / 10;
        constexpr int NSEG = 3, SL = (NSTEP + NSEG - 1) / NSEG;
        static_assert((NU - LO) % 10 == 0, "whole steps");
        const dcx rstep_c = cexp_turns(-fo / a.fs_dec);       // exp(+j W)
        const int sg = lane / 10, chain = lane % 10;          // lanes 30, 31 idle
        const bool active = lane < 10 * NSEG;
        dcx lam[4], om[4], sl[4], ps[4], pw[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            lam[r] = cmul(dcx{a.t.bp[2 * r], a.t.bp[2 * r + 1]}, rstep_c);
            om[r] = sl[r] = ps[r] = dcx{0.0, 0.0};
            pw[r] = lam[r];
        }
        const int j0 = sg * SL, j1 = min(j0 + SL, NSTEP);
        if (active) {
            constexpr int BLK = 8;
            for (int jb = j0; jb < j1; jb += BLK) {
                double g[BLK];
                dcx v[BLK];
#pragma unroll
                for (int q = 0; q < BLK; ++q) {
                    const int u = LO + chain + 10 * (jb + q);
                    const bool in = jb + q < j1;
                    g[q] = (in && u >= -ET_G && u <= ET_G) ? g1[u + ET_G] : 0.0;
                    const int d = u + k0;
                    v[q] = (in && d >= 0 && (int64_t)d < n) ? xr(d) : dcx{0.0, 0.0};
                }
#pragma unroll
                for (int q = 0; q < BLK; ++q) {
                    if (jb + q < j1) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            om[r] = cmul(lam[r], om[r]);
                            om[r].x += g[q];
                            const dcx t0 = cmul(v[q], om[r]), t1 = cmul(v[q], pw[r]);
                            sl[r].x += t0.x; sl[r].y += t0.y;
                            ps[r].x += t1.x; ps[r].y += t1.y;
                            pw[r] = cmul(pw[r], lam[r]);
                        }
                    }
                }
            }
        }
        // carries: C(segment) = lambda^SL C(previous segment) + Om_end(previous segment); pw / lambda = lambda^(steps taken)
        dcx C[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) C[r] = dcx{0.0, 0.0};
        for (int round = 1; round < NSEG; ++round) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double cx = __shfl_up_sync(0xffffffffu, C[r].x, 10), cy = __shfl_up_sync(0xffffffffu, C[r].y, 10);
                const double ex = __shfl_up_sync(0xffffffffu, om[r].x, 10), ey = __shfl_up_sync(0xffffffffu, om[r].y, 10);
                const double px = __shfl_up_sync(0xffffffffu, pw[r].x, 10), py = __shfl_up_sync(0xffffffffu, pw[r].y, 10);
                if (sg == round) {
                    // the previous segment took SL steps: its pw is lambda^(SL + 1)
                    const double den = lam[r].x * lam[r].x + lam[r].y * lam[r].y;
                    const dcx lsl = cmul(dcx{px, py}, dcx{lam[r].x / den, -lam[r].y / den});
                    const dcx t = cmul(lsl, dcx{cx, cy});
                    C[r] = dcx{t.x + ex, t.y + ey};
                }
            }
        }
        dcx tot[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const dcx t = cmul(C[r], ps[r]);
            tot[r] = active ? dcx{sl[r].x + t.x, sl[r].y + t.y} : dcx{0.0, 0.0};
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                tot[r].x += __shfl_xor_sync(0xffffffffu, tot[r].x, o);
                tot[r].y += __shfl_xor_sync(0xffffffffu, tot[r].y, o);
            }
        }
        if (lane < 4) {
            // sigma_r = c_r e^{-jW(L-1)} S_r ; s = V sigma
            const dcx rl = cexp_turns(fo * (double)(L - 1) / a.fs_dec);
            dcx st{0.0, 0.0};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const dcx sig = cmul(cmul(dcx{a.t.bc[2 * r], a.t.bc[2 * r + 1]}, rl), tot[r]);
                const dcx t = cmul(dcx{a.t.bv[2 * (4 * lane + r)], a.t.bv[2 * (4 * lane + r) + 1]}, sig);
                st.x += t.x; st.y += t.y;
            }
            out[KS_S2 + lane] = make_double2(st.x, st.y);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// k_edge_recursions: from the states to the corrections. Thread = one carrier, blockIdx.y = the end (0 left, 1 right):
// a few hundred serial order-4 steps per thread, the same instruction stream in every lane.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KC2_THREADS) k_edge_recursions(const EdgeCorrArgs a) {
    const int car = blockIdx.x * KC2_THREADS + threadIdx.x;
    if (car >= a.n_carriers) return;
    const int64_t n = a.n;
    const int L = a.L;
    const int k0 = (int)((n - 1) % 10);
    const double fo = a.fo ? a.fo[car] : 0.0;
    const ExactCoef& cf = a.cf;
    const dcx rstep = cexp_turns(fo / a.fs_dec);              // NCO rotation per 240 kS/s sample, exp(-j W)
    const dcx rstep_c{rstep.x, -rstep.y};                     // exp(+j W)
    auto rot_at = [&](int64_t j) { return cexp_turns(fo * (double)j / a.fs_dec); };
    const double2* stt = a.states + (int64_t)car * KC_NSTATE;
    auto ld = [&](int k) { const double2 v = stt[k]; return dcx{v.x, v.y}; };
    auto xr = [&](int d) { return ld(KS_XR + d); };           // x[n-1-d], d <= PAD1
    auto xl = [&](int i) { return ld(KS_XL + i); };           // x[i], i <= PAD1
    float2* dl = a.d + (int64_t)car * 2 * K_EDGE;
    float2* dr = dl + K_EDGE;
    double* ml = a.m_out ? a.m_out + (int64_t)car * 2 * K_EDGE : nullptr;
    double* mr = ml ? ml + K_EDGE : nullptr;
    KcBa st;

    if (blockIdx.y == 1) {
        // ================= right end =================
        dcx scv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) scv[k] = ld(KS_SC + k);
        // stage 1: forward pass over the 27-sample odd extension from the state after x[n-1] (sosfiltfilt's forward pass), then
        // the backward pass back over it from zi * (last forward output); the zero-extended stream has there the backward
        // pass's state after the forward ringing instead (U s_c)
        dcx ds[8];
        {
            SosState ss;
            kc_sos_from(ss, scv);
            const dcx x_last = xr(0);
            double2 yf[EX_PAD1];
#pragma unroll
            for (int j = 0; j < EX_PAD1; ++j) {                   // all loads first
                const dcx v = xr(1 + j);
                yf[j] = make_double2(2.0 * x_last.x - v.x, 2.0 * x_last.y - v.y);
            }
            for (int j = 0; j < EX_PAD1; ++j) yf[j] = sos_step(ss, cf, yf[j]);
            sos_init(ss, cf, yf[EX_PAD1 - 1]);
            for (int j = EX_PAD1 - 1; j >= 0; --j) sos_step(ss, cf, yf[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const dcx ku = kc_dot8(a.t.u + 8 * k, scv);
                ds[k] = dcx{ss.z[k >> 1][k & 1][0] - ku.x, ss.z[k >> 1][k & 1][1] - ku.y};
            }
        }
        // stage-1 correction d1[t] at output L-1-t (zero-input ringing of the backward pass from ds) after the NCO, and its
        // causal Butterworth response over the block's last TD outputs (ascending time = descending t)
        dcx ydl[ET_TD], d1s[ET_NPTS];
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{0.0, 0.0};
        {
#pragma unroll 8
            for (int t = 0; t < ET_TD; ++t) ydl[t] = kc_dot8(a.t.ringc + (int64_t)(k0 + 10 * t) * 8, ds);    // independent: loads in flight
            dcx rot = rot_at(L - ET_TD);
#pragma unroll 4
            for (int t = ET_TD - 1; t >= 0; --t) {
                const dcx d1 = cmul(ydl[t], rot);
                if (t < ET_NPTS) d1s[t] = d1;
                ydl[t] = kc_ba_step(st, cf, d1);
                rot = cmul(rot, rstep);
            }
        }
        // exact stream: the last 16 decimator outputs after the NCO, the 15-sample odd extension, forward pass held at F
        dcx s2k[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s2k[k] = ld(KS_S2 + k);
            st.z[k].x += s2k[k].x; st.z[k].y += s2k[k].y;
        }
        dcx y2e[EX_PAD2];
        {
            dcx zpe[ET_NPTS];
            dcx rot = rot_at(L - 1);
            for (int t = 0; t < ET_NPTS; ++t) {
                const dcx p = cmul(ld(KS_PTR + t), rot);
                zpe[t] = dcx{p.x + d1s[t].x, p.y + d1s[t].y};
                rot = cmul(rot, rstep_c);
            }
            for (int j = 0; j < EX_PAD2; ++j)
                y2e[j] = kc_ba_step(st, cf, dcx{2.0 * zpe[0].x - zpe[1 + j].x, 2.0 * zpe[0].y - zpe[1 + j].y});
        }
        const dcx F = y2e[EX_PAD2 - 1];
        // zero-extended stream beyond the block: the backward pass over the forward ringing, z[L+t] = RING[9-k0+10t] . s_c,
        // through the forward pass from that stream's state at the end
        dcx y2k[ET_T2];
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = s2k[k];
        {
#pragma unroll 8
            for (int t = 0; t < ET_TD; ++t) y2k[t] = kc_dot8(a.t.ring + (int64_t)(9 - k0 + 10 * t) * 8, scv);
            dcx rot = rot_at(L);
#pragma unroll 4
            for (int t = 0; t < ET_T2; ++t) {
                dcx in{0.0, 0.0};
                if (t < ET_TD) in = cmul(y2k[t], rot);
                y2k[t] = kc_ba_step(st, cf, in);
                rot = cmul(rot, rstep);
            }
        }
        // backward pass over the difference of the two forward outputs, from zi * F (filtfilt's backward start)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{cf.zi2[k] * F.x, cf.zi2[k] * F.y};
#pragma unroll 4
        for (int t = ET_T2 - 1; t >= 0; --t) {
            const dcx e = t < EX_PAD2 ? y2e[t] : F;
            kc_ba_step(st, cf, dcx{e.x - y2k[t].x, e.y - y2k[t].y});
        }
#pragma unroll 4
        for (int t = 0; t < K_EDGE; ++t) {
            const dcx v = kc_ba_step(st, cf, t < ET_TD ? ydl[t] : dcx{0.0, 0.0});
            if (mr) mr[t] = v.x; else dr[t] = make_float2((float)v.x, (float)v.y);
        }
    } else {
        // ================= left end =================
        dcx sac[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sac[k] = ld(KS_SAC + k);
        // stage 1: the exact forward state at position 0 -- zi * ext[0], then the 27 odd-extension samples (positions -27 .. -1);
        // the zero-extended stream starts from a zero state
        dcx dsc[8];
        {
            const dcx x0 = xl(0);
            double2 e[EX_PAD1];
#pragma unroll
            for (int j = 0; j < EX_PAD1; ++j) {                   // all loads first
                const dcx v = xl(EX_PAD1 - j);
                e[j] = make_double2(2.0 * x0.x - v.x, 2.0 * x0.y - v.y);
            }
            SosState ss;
            sos_init(ss, cf, e[0]);
            for (int j = 0; j < EX_PAD1; ++j) sos_step(ss, cf, e[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) dsc[k] = dcx{ss.z[k >> 1][k & 1][0], ss.z[k >> 1][k & 1][1]};
        }
        // stage-1 correction d1[m] at output m (backward pass over the forward ringing of dsc) after the NCO
        dcx dy[K_EDGE];
        {
#pragma unroll 8
            for (int m = 0; m < ET_TD; ++m) dy[m] = kc_dot8(a.t.ring + (int64_t)(10 * m) * 8, dsc);
            dcx rot{1.0, 0.0};
            for (int m = 0; m < ET_TD; ++m) {
                dy[m] = cmul(dy[m], rot);
                rot = cmul(rot, rstep);
            }
            for (int m = ET_TD; m < K_EDGE; ++m) dy[m] = dcx{0.0, 0.0};
        }
        // exact decimator outputs 0..15 after the NCO, stage 2's odd extension (positions -15 .. -1) from zi * ext[0]
        {
            dcx zpe[ET_NPTS];
            dcx rot{1.0, 0.0};
            for (int t = 0; t < ET_NPTS; ++t) {
                const dcx p = cmul(ld(KS_PTL + t), rot);
                zpe[t] = dcx{p.x + dy[t].x, p.y + dy[t].y};
                rot = cmul(rot, rstep);
            }
            const dcx e0{2.0 * zpe[0].x - zpe[EX_PAD2].x, 2.0 * zpe[0].y - zpe[EX_PAD2].y};
#pragma unroll
            for (int k = 0; k < 4; ++k) st.z[k] = dcx{cf.zi2[k] * e0.x, cf.zi2[k] * e0.y};
            for (int j = 0; j < EX_PAD2; ++j)
                kc_ba_step(st, cf, dcx{2.0 * zpe[0].x - zpe[EX_PAD2 - j].x, 2.0 * zpe[0].y - zpe[EX_PAD2 - j].y});
        }
        // the zero-extended stream before the block is the backward pass's own ringing, z[-t] = RINGC[10t-1] . s_ac0: its forward
        // state at position 0 (from a zero state)
        {
            KcBa sk;
#pragma unroll
            for (int k = 0; k < 4; ++k) sk.z[k] = dcx{0.0, 0.0};
            dcx zk[ET_TD];
#pragma unroll 8
            for (int t = 1; t <= ET_TD; ++t) zk[t - 1] = kc_dot8(a.t.ringc + (int64_t)(10 * t - 1) * 8, sac);
            dcx rot = rot_at(-ET_TD);
#pragma unroll 4
            for (int t = ET_TD; t >= 1; --t) {
                kc_ba_step(sk, cf, cmul(zk[t - 1], rot));
                rot = cmul(rot, rstep);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { st.z[k].x -= sk.z[k].x; st.z[k].y -= sk.z[k].y; }
        }
        // causal response to the state difference and the stage-1 correction over the K_EDGE outputs; behind them the forward
        // state rings out, which the backward pass sees as the start state U2 . state; then the backward pass
#pragma unroll 4
        for (int m = 0; m < K_EDGE; ++m) dy[m] = kc_ba_step(st, cf, dy[m]);
        KcBa sb;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sb.z[k] = dcx{0.0, 0.0};
#pragma unroll
            for (int j = 0; j < 4; ++j) { sb.z[k].x += a.t.u2[4 * k + j] * st.z[j].x; sb.z[k].y += a.t.u2[4 * k + j] * st.z[j].y; }
        }
#pragma unroll 4
        for (int m = K_EDGE - 1; m >= 0; --m) {
            const dcx v = kc_ba_step(sb, cf, dy[m]);
            if (ml) ml[m] = v.x; else dl[m] = make_float2((float)v.x, (float)v.y);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// k_edge_apply (freq_offset = 0): D = M s. blockIdx.y = the end; a CTA takes KA_CPB carriers, thread t the output t of each:
// a row of M is read once per CTA (coalesced over t) and meets the carriers' states from shared memory.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KA_THREADS) k_edge_apply(const EdgeCorrArgs a) {
    __shared__ double2 s_st[KA_CPB][KS_NR];
    const int e = blockIdx.y;                                  // 0 left, 1 right
    const int base = e ? 0 : KS_NR, nk = e ? KS_NR : KS_NL;
    const int c0 = blockIdx.x * KA_CPB;
    for (int i = threadIdx.x; i < KA_CPB * nk; i += KA_THREADS) {
        const int c = i / nk, k = i - c * nk;
        s_st[c][k] = c0 + c < a.n_carriers ? a.states[(int64_t)(c0 + c) * KC_NSTATE + base + k] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= K_EDGE) return;
    double ax[KA_CPB], ay[KA_CPB];
#pragma unroll
    for (int c = 0; c < KA_CPB; ++c) ax[c] = ay[c] = 0.0;
    const double* __restrict__ mp = a.m + ((int64_t)base * 2 + e) * K_EDGE + t;
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
        const double w = __ldg(mp + (int64_t)k * 2 * K_EDGE);
#pragma unroll
        for (int c = 0; c < KA_CPB; ++c) { ax[c] += w * s_st[c][k].x; ay[c] += w * s_st[c][k].y; }
    }
#pragma unroll
    for (int c = 0; c < KA_CPB; ++c)
        if (c0 + c < a.n_carriers) a.d[((int64_t)(c0 + c) * 2 + e) * K_EDGE + t] = make_float2((float)ax[c], (float)ay[c]);
}

// unit states for the M build: carrier k holds state k = 1
__global__ void k_edge_unit_states(double2* st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < KC_NSTATE * KC_NSTATE) st[i] = make_double2((i / KC_NSTATE) == (i % KC_NSTATE) ? 1.0 : 0.0, 0.0);
}

}  // namespace tetra
