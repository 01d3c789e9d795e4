// KC: block-end corrections of the fused path.
//
// The reference chain  decimate(x, 10) -> [NCO] -> filtfilt(butter4)  (tetraear/signal/processor.py:245-264) is
// shift-invariant except next to the two ends of a block, where SciPy's sosfiltfilt / filtfilt use an odd extension
// (27 / 15 samples), steady-state initial conditions (zi * first sample) and hold the forward pass's last output for
// the backward pass. The fused kernel computes the shift-invariant part: its cascade applied to the block extended by
// ZEROS. This kernel computes, in float64, the difference
//        D[m] = reference(x)[m] - cascade(x zero-extended)[m]        for the K_EDGE outputs next to each end,
// which is linear in x and is driven by a handful of IIR filter states at the block end. Those states are dot products
// of the block's first / last ~1200 samples with fixed weight tables (edge_tables_generated.h, from tools/edge_model.py,
// where the derivation and a float64 model of exactly these steps live); a few hundred literal order-4 recursion
// steps at 240 kS/s then give D. The kernel depends on the input only, so it runs beside the fused kernel; the
// finalize kernel adds D to the fused kernel's output where it reads the block ends.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tetra_exact.cuh"
#include "edge_tables_generated.h"

namespace tetra {

constexpr int KC_THREADS = 128;
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
};

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
    float2* d;               // [C][2][K_EDGE]: D_left[m] (m = 0..), D_right[t] (output L-1-t)
};

struct dcx { double x, y; };
__device__ __forceinline__ dcx cmul(dcx a, dcx b) { return dcx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ dcx cexp_turns(double turns) {          // exp(-j 2 pi turns), phase reduced first
    dcx r;
    sincospi(-2.0 * (turns - rint(turns)), &r.y, &r.x);
    return r;
}

struct KcSmem {
    double red[KC_THREADS / 32][8][2];
    dcx s_c[8], s_ac0[8], pts_r[ET_NPTS], pts_l[ET_NPTS], s2_k1[4], ds[8], dsc[8], ds2[4], s2_l[4], y2e[EX_PAD2];
    dcx buf_r1[ET_TD];       // right end: stage-1 correction d1[t] (output L-1-t), then its causal response in place
    dcx buf_r2[ET_T2];       // right end: zero-extended stream beyond the block z'[L+t], then its causal response in place
    dcx buf_l1[K_EDGE];      // left end: stage-1 correction d1[m], then the causal response of the difference in place
    dcx buf_l2[ET_TD];       // left end: zero-extended stream before the block z'[-t], t = 1..TD at index t-1
};

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

// block-wide sums of 8 complex accumulators -> dst[0..8) (valid after the trailing barrier)
__device__ __forceinline__ void kc_reduce8(KcSmem& s, dcx (&acc)[8], dcx* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
            acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
        }
        if (lane == 0) { s.red[warp][k][0] = acc[k].x; s.red[warp][k][1] = acc[k].y; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        dcx t{0.0, 0.0};
        for (int w = 0; w < KC_THREADS / 32; ++w) { t.x += s.red[w][threadIdx.x][0]; t.y += s.red[w][threadIdx.x][1]; }
        dst[threadIdx.x] = t;
    }
    __syncthreads();
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

// one weighted sum pass: acc[k] += w_k(j) * x(j) over j = tid, tid + 128, ...; two samples in flight per thread
#define KC_PASS(COUNT, LOADX, WEIGHT)                                                                   \
    for (int j0 = tid; j0 < (COUNT); j0 += 2 * KC_THREADS) {                                            \
        const int j1 = j0 + KC_THREADS;                                                                 \
        const bool two = j1 < (COUNT);                                                                  \
        const dcx v0 = LOADX(j0);                                                                       \
        const dcx v1 = two ? LOADX(j1) : dcx{0.0, 0.0};                                                 \
        double w0[8], w1[8];                                                                            \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) { w0[k] = WEIGHT(k, j0); w1[k] = two ? WEIGHT(k, j1) : 0.0; } \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) {                                                 \
            acc[k].x += w0[k] * v0.x; acc[k].y += w0[k] * v0.y;                                         \
            acc[k].x += w1[k] * v1.x; acc[k].y += w1[k] * v1.y;                                         \
        }                                                                                               \
    }

template <int IN>   // 0: complex64 rows, 1: uint8 rows, 2: channels of one shared complex64 capture
__global__ void __launch_bounds__(KC_THREADS) k_edge_correct(const EdgeCorrArgs a) {
    __shared__ KcSmem s;
    const int car = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n = a.n;
    const int L = a.L;
    const int k0 = (int)((n - 1) % 10);
    const double chan_hz = (IN == 2) ? a.chan[car] : 0.0;
    const double fo = a.fo ? a.fo[car] : 0.0;
    auto xr = [&](int d) { return kc_load<IN>(a, car, chan_hz, n - 1 - d); };      // d samples before the last one
    auto xl = [&](int i) { return kc_load<IN>(a, car, chan_hz, i); };
    const ExactCoef& cf = a.cf;
    const dcx rstep = cexp_turns(fo / a.fs_dec);              // NCO rotation per 240 kS/s sample, exp(-j W)
    const dcx rstep_c{rstep.x, -rstep.y};                     // exp(+j W)
    auto rot_at = [&](int64_t j) { return fo == 0.0 ? dcx{1.0, 0.0} : cexp_turns(fo * (double)j / a.fs_dec); };

    // ---------------- phase A: states and pointwise decimator outputs as dot products with the weight tables ----------------
    dcx acc[8];
    auto clear = [&]() {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = dcx{0.0, 0.0};
    };
    {   // causal cascade state after x[n-1]
        clear();
        const double* __restrict__ wc = a.t.wc;
#define KC_W(k, j) wc[(k) * ET_NC + (j)]
        KC_PASS(ET_NC, xr, KC_W)
#undef KC_W
        kc_reduce8(s, acc, s.s_c);
    }
    {   // backward-pass state at position 0 of the zero-extended stream
        clear();
        const double* __restrict__ wac = a.t.wac;
#define KC_W(k, j) wac[(k) * ET_NAC + (j)]
        KC_PASS(ET_NAC, xl, KC_W)
#undef KC_W
        kc_reduce8(s, acc, s.s_ac0);
    }
    const double* __restrict__ g1 = a.t.g1;
    for (int h = 0; h < ET_NPTS / 8; ++h) {
        // z of the zero-extended stream at output L-1-t: sum_d g1[d - k0 - 10 t] x[n-1-d]   (t = 8 h + k)
        clear();
        const int off_r = ET_G - k0 - 80 * h;
#define KC_W(k, j) (((j) + off_r - 10 * (k)) >= 0 && ((j) + off_r - 10 * (k)) <= 2 * ET_G ? g1[(j) + off_r - 10 * (k)] : 0.0)
        KC_PASS(k0 + 10 * (8 * h + 7) + ET_G + 1, xr, KC_W)
#undef KC_W
        kc_reduce8(s, acc, s.pts_r + 8 * h);
        // ... and at output t: sum_i g1[10 t - i] x[i]
        clear();
        const int off_l = ET_G + 80 * h;
#define KC_W(k, j) ((off_l + 10 * (k) - (j)) >= 0 && (off_l + 10 * (k) - (j)) <= 2 * ET_G ? g1[off_l + 10 * (k) - (j)] : 0.0)
        KC_PASS(10 * (8 * h + 7) + ET_G + 1, xl, KC_W)
#undef KC_W
        kc_reduce8(s, acc, s.pts_l + 8 * h);
    }

    // ---------------- phase B: one job per warp ----------------
    if (warp == 0) {
        // right end, stage 1. Forward pass over the 27-sample odd extension from the state after x[n-1] (sosfiltfilt's
        // forward pass), then the backward pass back over it from zi * (last forward output); the zero-extended stream
        // has there the backward pass's state after the forward ringing instead (U s_c)
        if (lane == 0) {
            dcx scv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) scv[k] = s.s_c[k];
            SosState ss;
            kc_sos_from(ss, scv);
            const dcx x_last = xr(0);
            double2 yf[EX_PAD1];
#pragma unroll
            for (int j = 0; j < EX_PAD1; ++j) {
                const dcx v = xr(1 + j);
                yf[j] = make_double2(2.0 * x_last.x - v.x, 2.0 * x_last.y - v.y);
            }
            for (int j = 0; j < EX_PAD1; ++j) yf[j] = sos_step(ss, cf, yf[j]);
            sos_init(ss, cf, yf[EX_PAD1 - 1]);
            for (int j = EX_PAD1 - 1; j >= 0; --j) sos_step(ss, cf, yf[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const dcx ku = kc_dot8(a.t.u + 8 * k, scv);
                s.ds[k] = dcx{ss.z[k >> 1][k & 1][0] - ku.x, ss.z[k >> 1][k & 1][1] - ku.y};
            }
        }
        __syncwarp();
        // stage-1 correction d1[t] at output L-1-t: zero-input ringing of the backward pass from ds, after the NCO
        dcx dsv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dsv[k] = s.ds[k];
        for (int t = lane; t < ET_TD; t += 32)
            s.buf_r1[t] = cmul(kc_dot8(a.t.ringc + (int64_t)(k0 + 10 * t) * 8, dsv), rot_at(L - 1 - t));
    } else if (warp == 1) {
        // left end, stage 1: the exact forward state at position 0 -- zi * ext[0], then the 27 odd-extension samples
        // (positions -27 .. -1); the zero-extended stream starts from a zero state
        if (lane == 0) {
            const dcx x0 = xl(0);
            double2 e[EX_PAD1];
#pragma unroll
            for (int j = 0; j < EX_PAD1; ++j) {
                const dcx v = xl(EX_PAD1 - j);
                e[j] = make_double2(2.0 * x0.x - v.x, 2.0 * x0.y - v.y);
            }
            SosState ss;
            sos_init(ss, cf, e[0]);
            for (int j = 0; j < EX_PAD1; ++j) sos_step(ss, cf, e[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) s.dsc[k] = dcx{ss.z[k >> 1][k & 1][0], ss.z[k >> 1][k & 1][1]};
        }
        __syncwarp();
        dcx dsv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dsv[k] = s.dsc[k];
        for (int m = lane; m < K_EDGE; m += 32)
            s.buf_l1[m] = m < ET_TD ? cmul(kc_dot8(a.t.ring + (int64_t)(10 * m) * 8, dsv), rot_at(m)) : dcx{0.0, 0.0};
    } else if (warp == 2) {
        // the zero-extended stream outside the block: beyond the end the backward pass over the forward ringing,
        // z[L+t] = RING[9-k0+10t] . s_c; before the start the backward pass's own ringing, z[-t] = RINGC[10t-1] . s_ac0
        dcx v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = s.s_c[k];
        for (int t = lane; t < ET_T2; t += 32)
            s.buf_r2[t] = t < ET_TD ? cmul(kc_dot8(a.t.ring + (int64_t)(9 - k0 + 10 * t) * 8, v), rot_at(L + t)) : dcx{0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = s.s_ac0[k];
        for (int t = 1 + lane; t <= ET_TD; t += 32)
            s.buf_l2[t - 1] = cmul(kc_dot8(a.t.ringc + (int64_t)(10 * t - 1) * 8, v), rot_at(-t));
    } else {
        // causal Butterworth state of the zero-extended stream at the right end:
        // s(L) = e^{-jW(L-1)} sum_d x[n-1-d] Wv(d - k0),  Wv(u) = Bv g1[u] + e^{jW} A Wv(u - 10): ten independent chains,
        // inputs fetched a block of steps ahead of the recursion
        dcx S[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        if (lane < 10) {
            double bv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) bv[k] = cf.b[k + 1] - cf.a[k + 1] * cf.b[0];
            constexpr int LO = -(ET_G + 10) - ((-(ET_G + 10)) % 10 + 10) % 10;     // multiple of 10 at or below -(G + 10)
            constexpr int NU = ET_G + 10 * ET_T2 + 10;
            constexpr int NSTEP = (NU - LO) / 10;
            constexpr int BLK = 8;
            static_assert((NU - LO) % 10 == 0, "whole steps");
            dcx W[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            auto fetch = [&](int step, double& g, dcx& v) {
                const int u = LO + lane + 10 * step;
                g = (step < NSTEP && u >= -ET_G && u <= ET_G) ? g1[u + ET_G] : 0.0;
                const int d = u + k0;
                v = (step < NSTEP && d >= 0 && (int64_t)d < n) ? xr(d) : dcx{0.0, 0.0};
            };
            double gc[BLK], gn[BLK];
            dcx vc[BLK], vn[BLK];
#pragma unroll
            for (int q = 0; q < BLK; ++q) fetch(q, gc[q], vc[q]);
            for (int s0 = 0; s0 < NSTEP; s0 += BLK) {
#pragma unroll
                for (int q = 0; q < BLK; ++q) fetch(s0 + BLK + q, gn[q], vn[q]);
#pragma unroll
                for (int q = 0; q < BLK; ++q) {
                    if (s0 + q < NSTEP) {
                        dcx t[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            t[k].x = -cf.a[k + 1] * W[0].x + (k < 3 ? W[k + 1].x : 0.0);
                            t[k].y = -cf.a[k + 1] * W[0].y + (k < 3 ? W[k + 1].y : 0.0);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            W[k] = fo == 0.0 ? t[k] : cmul(rstep_c, t[k]);
                            W[k].x += bv[k] * gc[q];
                            const dcx p = cmul(vc[q], W[k]);
                            S[k].x += p.x; S[k].y += p.y;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < BLK; ++q) { gc[q] = gn[q]; vc[q] = vn[q]; }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 8; o; o >>= 1) {
                S[k].x += __shfl_xor_sync(0xffffffffu, S[k].x, o);
                S[k].y += __shfl_xor_sync(0xffffffffu, S[k].y, o);
            }
        }
        if (lane < 4) s.s2_k1[lane] = cmul(S[lane], rot_at(L - 1));
    }
    __syncthreads();

    // ---------------- phase C: the forward recursions of stage 2, one thread each ----------------
    float2* dl = a.d + (int64_t)car * 2 * K_EDGE;
    float2* dr = dl + K_EDGE;
    KcBa st;
    if (tid == 0) {
        // right end: causal response of the stage-1 correction over the block's last TD outputs (ascending time = descending t)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{0.0, 0.0};
        for (int t = ET_TD - 1; t >= 0; --t) s.buf_r1[t] = kc_ba_step(st, cf, s.buf_r1[t]);
#pragma unroll
        for (int k = 0; k < 4; ++k) s.ds2[k] = st.z[k];
    } else if (tid == 32) {
        // left end: exact decimator outputs 0..15 after the NCO, stage 2's odd extension (positions -15 .. -1) from zi * ext[0]
        dcx zpe[ET_NPTS];
        {
            dcx rot{1.0, 0.0};
#pragma unroll
            for (int t = 0; t < ET_NPTS; ++t) {
                const dcx p = cmul(s.pts_l[t], rot);
                zpe[t] = dcx{p.x + s.buf_l1[t].x, p.y + s.buf_l1[t].y};
                rot = cmul(rot, rstep);
            }
        }
        const dcx e0{2.0 * zpe[0].x - zpe[EX_PAD2].x, 2.0 * zpe[0].y - zpe[EX_PAD2].y};
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{cf.zi2[k] * e0.x, cf.zi2[k] * e0.y};
#pragma unroll
        for (int j = 0; j < EX_PAD2; ++j)
            kc_ba_step(st, cf, dcx{2.0 * zpe[0].x - zpe[EX_PAD2 - j].x, 2.0 * zpe[0].y - zpe[EX_PAD2 - j].y});
#pragma unroll
        for (int k = 0; k < 4; ++k) s.s2_l[k] = st.z[k];
    } else if (tid == 64) {
        // right end: the zero-extended stream's forward pass beyond the block, from its state at the end
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = s.s2_k1[k];
        for (int t = 0; t < ET_T2; ++t) s.buf_r2[t] = kc_ba_step(st, cf, s.buf_r2[t]);
    } else if (tid == 96) {
        // left end: the zero-extended stream's forward state at position 0 (its ringing before the block, from a zero state)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{0.0, 0.0};
        for (int t = ET_TD; t >= 1; --t) kc_ba_step(st, cf, s.buf_l2[t - 1]);
#pragma unroll
        for (int k = 0; k < 4; ++k) s.buf_l2[k] = st.z[k];
    }
    __syncthreads();

    // ---------------- phase D: the rest of the forward passes and the backward passes ----------------
    if (tid == 0) {
        // right end, exact stream: the last 16 decimator outputs after the NCO, the 15-sample odd extension, forward pass held at F
        dcx zpe[ET_NPTS];
        {
            dcx rot = rot_at(L - 1);
#pragma unroll
            for (int t = 0; t < ET_NPTS; ++t) {
                const dcx p = cmul(s.pts_r[t], rot);
                zpe[t] = p;
                rot = cmul(rot, rstep_c);
            }
        }
        // (the stage-1 correction of those outputs: it was overwritten by its causal response; recompute the 16 values)
        {
            dcx dsv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) dsv[k] = s.ds[k];
            dcx rot = rot_at(L - 1);
            for (int t = 0; t < ET_NPTS; ++t) {
                const dcx d1 = cmul(kc_dot8(a.t.ringc + (int64_t)(k0 + 10 * t) * 8, dsv), rot);
                zpe[t].x += d1.x; zpe[t].y += d1.y;
                rot = cmul(rot, rstep_c);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{s.s2_k1[k].x + s.ds2[k].x, s.s2_k1[k].y + s.ds2[k].y};
        dcx y2e[EX_PAD2];
#pragma unroll
        for (int j = 0; j < EX_PAD2; ++j)
            y2e[j] = kc_ba_step(st, cf, dcx{2.0 * zpe[0].x - zpe[1 + j].x, 2.0 * zpe[0].y - zpe[1 + j].y});
        const dcx F = y2e[EX_PAD2 - 1];
        // backward pass over the difference of the two forward outputs, from zi * F (filtfilt's backward start)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{cf.zi2[k] * F.x, cf.zi2[k] * F.y};
        for (int t = ET_T2 - 1; t >= EX_PAD2; --t) kc_ba_step(st, cf, dcx{F.x - s.buf_r2[t].x, F.y - s.buf_r2[t].y});
#pragma unroll
        for (int t = EX_PAD2 - 1; t >= 0; --t) kc_ba_step(st, cf, dcx{y2e[t].x - s.buf_r2[t].x, y2e[t].y - s.buf_r2[t].y});
        for (int t = 0; t < K_EDGE; ++t) {
            const dcx v = kc_ba_step(st, cf, t < ET_TD ? s.buf_r1[t] : dcx{0.0, 0.0});
            dr[t] = make_float2((float)v.x, (float)v.y);
        }
    } else if (tid == 32) {
        // left end: causal response to the state difference and the stage-1 correction over the K_EDGE outputs; behind them
        // the forward state rings out, which the backward pass sees as the start state U2 . state; then the backward pass
#pragma unroll
        for (int k = 0; k < 4; ++k) st.z[k] = dcx{s.s2_l[k].x - s.buf_l2[k].x, s.s2_l[k].y - s.buf_l2[k].y};
        for (int m = 0; m < K_EDGE; ++m) s.buf_l1[m] = kc_ba_step(st, cf, s.buf_l1[m]);
        KcBa sb;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sb.z[k] = dcx{0.0, 0.0};
#pragma unroll
            for (int j = 0; j < 4; ++j) { sb.z[k].x += a.t.u2[4 * k + j] * st.z[j].x; sb.z[k].y += a.t.u2[4 * k + j] * st.z[j].y; }
        }
        for (int m = K_EDGE - 1; m >= 0; --m) {
            const dcx v = kc_ba_step(sb, cf, s.buf_l1[m]);
            dl[m] = make_float2((float)v.x, (float)v.y);
        }
    }
}
#undef KC_PASS

}  // namespace tetra
