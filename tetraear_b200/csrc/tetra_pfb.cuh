// K4: the shared front end of BASELINE config 3 -- a 96-channel polyphase-DFT channelizer.
//
// Config 3 asks, per channel centre f_c of one wideband capture, for process(frequency_shift(x, f_c), 0)
// (signal/processor.py:85-100 + :221-273; the reference has no channelizer, its scanner retunes the hardware in 25 kHz
// steps, signal/scanner.py:383-445). On the 25 kHz grid of a 2.4 MS/s capture f_c / fs = c / 96 with c = -48 .. 47, so the first
// stage of the fused cascade -- the 41-tap proto /10 applied to the shifted stream --
//        w_c[m] = sum_d p[d] x[i] e^{-j 2 pi c i / 96},   i = 10 m - 20 + d,  d = 0 .. 40
// is, for all 96 channels at once, ONE 96-point DFT per output instant m of the tap-weighted samples placed at i mod 96.
// k_pfb96 computes it as 3 x 32 (three-point DFTs in registers, then three radix-2 FFTs of 32 points across the lanes of
// a warp with shuffles) and writes the 96 streams w_c at 240 kS/s; the fused kernel's MODE 5 then runs the per-channel
// stages (half-band, fir120, x2 interpolation, phase sums) on those streams. The block is extended by zeros, as the
// block-end corrections (tetra_edgecorr.cuh) assume.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tetra_kernels.cuh"

namespace tetra {

constexpr int PFB_NCH = 96;
constexpr int PFB_MB = 32;                // output instants per CTA
constexpr int PFB_THREADS = 256;          // 8 warps x 4 instants
constexpr int PFB_M0 = 384;               // w_c[m] is stored at column m + PFB_M0 (the fused kernel's pre-roll reaches m = -322)
static_assert(PFB_M0 >= K1_PREROLL - K1_A0 && (PFB_M0 % 2) == 0, "pre-roll of the fused kernel must fit in front of the block");

__constant__ float2 c_w96[PFB_NCH];       // e^{-j 2 pi q / 96}

struct PfbArgs {
    const float2* x;         // [n] wideband capture
    int64_t n;
    float2* w;               // [96][wp]: channel k = c + 48 at row k
    int64_t wp;              // row pitch (a multiple of PFB_MB)
};

__device__ __forceinline__ float2 pfb_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(PFB_THREADS) k_pfb96(const PfbArgs a) {
    __shared__ float2 xs[10 * PFB_MB + 48];
    __shared__ float2 outs[PFB_NCH][PFB_MB + 1];
    __shared__ float s_proto[2 * TB_PROTO_H + 1];            // lanes index the taps by different d: shared, not constant, memory
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid <= 2 * TB_PROTO_H) s_proto[tid] = c_proto[tid];
    // the twiddles a lane uses never change: read the (lane-indexed, hence serialised) constant bank once, not per instant
    const float2 w_l1 = c_w96[lane], w_l2 = c_w96[2 * lane];
    float2 w_st[5];
#pragma unroll
    for (int st = 0; st < 5; ++st) w_st[st] = c_w96[3 * ((lane & ((16 >> st) - 1)) << st)];   // W32^{(lane mod half) 2^s}
    const int col0 = blockIdx.x * PFB_MB;                    // first column of this CTA; m = col - PFB_M0
    const int64_t i0 = 10 * ((int64_t)col0 - PFB_M0) - TB_PROTO_H;   // input index of xs[0]
    for (int t = tid; t < 10 * PFB_MB + 41; t += PFB_THREADS) {
        const int64_t i = i0 + t;
        xs[t] = (i >= 0 && i < a.n) ? __ldg(a.x + i) : make_float2(0.f, 0.f);      // the block is extended by zeros
    }
    __syncthreads();
    const float2 w3_1 = c_w96[32], w3_2 = c_w96[64];
#pragma unroll
    for (int j = 0; j < PFB_MB / 8; ++j) {
        const int ml = warp * (PFB_MB / 8) + j;
        // first input index of this instant, mod 96 (i0 + 10 ml may be negative)
        const int64_t ifirst = i0 + 10 * ml;
        int r = (int)(ifirst % PFB_NCH);
        if (r < 0) r += PFB_NCH;
        // u[q] = p[d] x[ifirst + d] at q = (r + d) mod 96; this lane holds q = lane, 32 + lane, 64 + lane
        float2 u[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            int d = 32 * g + lane - r;
            if (d < 0) d += PFB_NCH;
            u[g] = make_float2(0.f, 0.f);
            if (d <= 2 * TB_PROTO_H) {
                const float2 v = xs[10 * ml + d];
                const float p = s_proto[d];
                u[g] = make_float2(p * v.x, p * v.y);
            }
        }
        // three-point DFTs over g, then the twiddles W96^{f lane}
        float2 t[3];
        t[0] = make_float2(u[0].x + u[1].x + u[2].x, u[0].y + u[1].y + u[2].y);
        {
            const float2 a1 = pfb_cmul(u[1], w3_1), a2 = pfb_cmul(u[2], w3_2);
            const float2 b1 = pfb_cmul(u[1], w3_2), b2 = pfb_cmul(u[2], w3_1);      // W3^2 and W3^4 = W3
            t[1] = pfb_cmul(make_float2(u[0].x + a1.x + a2.x, u[0].y + a1.y + a2.y), w_l1);
            t[2] = pfb_cmul(make_float2(u[0].x + b1.x + b2.x, u[0].y + b1.y + b2.y), w_l2);
        }
        // three 32-point FFTs across the lanes: decimation in frequency, results in bit-reversed lane order
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int half = 16 >> s;
            const float2 tw = w_st[s];                                        // W32^{(lane mod half) 2^s}
            const bool upper = (lane & half) != 0;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const float ox = __shfl_xor_sync(0xffffffffu, t[f].x, half), oy = __shfl_xor_sync(0xffffffffu, t[f].y, half);
                if (!upper) t[f] = make_float2(t[f].x + ox, t[f].y + oy);
                else t[f] = pfb_cmul(make_float2(ox - t[f].x, oy - t[f].y), tw);
            }
        }
        const int e = (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const int c = 3 * e + f;                          // DFT bin = c mod 96 for channel offset c (c >= 48: c - 96)
            outs[(c + 48) % PFB_NCH][ml] = t[f];
        }
    }
    __syncthreads();
    for (int i = tid; i < PFB_NCH * PFB_MB; i += PFB_THREADS) {
        const int k = i / PFB_MB, ml = i % PFB_MB;
        a.w[(int64_t)k * a.wp + col0 + ml] = outs[k][ml];
    }
}

}  // namespace tetra
