// Device kernels of libtetra_b200 (sm_100a). See DESIGN.md for the data layout and the
// derivation of every stage; tools/design_filters.py for the FIR tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "taps_generated.h"

namespace tetra {

// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float s, float2 c) {
    // packed fp32x2 FMA (Blackwell FFMA2): (a.x, a.y) * (s, s) + c
    float2 b = make_float2(s, s);
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP), completion on mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// K1  channelize_demod_fused  (fused path, fs = 2.4 MS/s)
// ----------------------------------------------------------------------------------------------
// A persistent CTA streams its (carrier, segment) items back to back through the four FIR stages
//   A proto /10 -> B half-band /2 -> C fir120 (127 taps) -> D x2 interp + store + |y|^2 phase sums
// with every stage working, in the same iteration, on data produced in earlier iterations, so a
// single bar.sync per 6400-sample tile is the only block-wide barrier. IQ tiles arrive by
// 1-D TMA bulk copies into a 3-deep ring, two tiles ahead of the compute.
//
// Stage C keeps its accumulators in registers across four iterations: each of its four warps owns
// every fourth group of 320 outputs and adds one quarter of the taps per iteration (oldest
// inputs first), so no partial sums ever go through shared memory.
constexpr int K1_TILE = 6400;             // input samples per iteration
constexpr int K1_HDR = 40;                // samples of the previous tile kept in front of each buffer
constexpr int K1_W = K1_TILE / 10;        // 640 w (240 kS/s) samples per iteration
constexpr int K1_U = K1_TILE / 20;        // 320 u/v (120 kS/s) samples per iteration
constexpr int K1_NBUF = 3;
constexpr int K1_WRING = 2048, K1_URING = 2048, K1_VRING = 1024;
constexpr int K1_RRING = 1024;            // MODE 1: ring of half-band outputs waiting for the equaliser
// every ring repeats its first PAD entries behind its end (writers store those twice), so that a reader's window of up to PAD
// consecutive entries is one base address plus immediate offsets -- no index mask per load
constexpr int K1_WPAD = 32, K1_UPAD = 48, K1_VPAD = 32, K1_RPAD = 16;
constexpr int K1_THREADS = 384;           // warps 0-3: A, 4-7: C, 8-9: B, 10-11: D
constexpr int K1_DLANES = 64;
constexpr int K1_CQ = 4;                  // tap quarters = iterations a C group stays in registers
// local (stream-origin relative) index ranges produced in iteration i
constexpr int K1_A0 = -2;                 // w: [640 i + A0, +640)
constexpr int K1_B0 = -326;               // u: [320 i + B0, +320) (MODE 1: the half-band output before the equaliser)
constexpr int K1_R0 = K1_B0 - 326;        // MODE 1: equalised u [320 i + R0, +320), from earlier iterations' half-band output
constexpr int K1_C0 = -942;               // v group g: [320 g + C0, +320), finished in iteration g + 3
constexpr int K1_D0 = -2230;              // v pairs [320 i + D0, +320) -> y [640 i + 2 D0, +640)
constexpr int K1_PREROLL = 320;           // pre-roll and post-roll of a slot in w samples (> the cascade's reach: 157, with the equaliser 167)
constexpr int K1_EDGE = 168;              // y samples at each block end owned by the exact kernel (> the cascade's reach in every mode)
constexpr int K1_NPH = 13;
constexpr int K1_MIN_T_ITEM = 6;           // a slot outlasts stage C's lag behind stage A (two tap buffers suffice)
constexpr double K1_FO_MAX_HZ = 12500.0;  // freq_offset range of the fused path (the proto's alias nulls cover +-60 kHz +- this)

static_assert(TB_PROTO_H == 20 && TB_HB_H == 11 && TB_FIR_H == 63 && TB_INT_K == 8 && TB_REQ_K == 5, "tables changed: re-derive lags");
// dependency checks (each stage only reads what earlier iterations produced)
static_assert(2 * (K1_B0 + K1_U - 1) + TB_HB_H <= K1_A0 - 1, "B reads w of a later iteration");
// quarter q of group g runs in iteration g + q and reads u up to 320 g + C0 + 319 - 64 + 32 q + 31 (+1 for the 129th tap)
static_assert((K1_C0 + K1_U - 1) - 64 + 31 <= K1_R0 - 1, "C quarter 0 reads u of a later iteration");
static_assert(K1_R0 + K1_U - 1 + TB_REQ_K <= K1_B0 - 1, "the equaliser reads half-band output of this or a later iteration");
static_assert((K1_B0 + K1_U) - (K1_R0 - TB_REQ_K) <= K1_RRING, "equaliser ring too small");
static_assert((K1_C0 + K1_U - 1) - 64 + 32 * 3 + 31 <= 2 * K1_U + K1_B0 + K1_U - 1, "C quarter 3 reads u of a later iteration");
// D in iteration i reads v up to 320 i + D0 + 319 + 8; finished groups then: g <= i - 4
static_assert((K1_D0 + K1_U - 1) + TB_INT_K <= -K1_CQ * K1_U + K1_C0 + K1_U - 1, "D reads v of an unfinished group");
static_assert((K1_A0 + K1_W) - (2 * K1_B0 - TB_HB_H) <= K1_WRING, "w ring too small");
static_assert((3 * K1_U + K1_B0 + K1_U) - (K1_C0 - 64 + 32 * 3) <= K1_URING, "u ring too small");
static_assert((-3 * K1_U + K1_C0 + K1_U) - (K1_D0 - TB_INT_K) <= K1_VRING, "v ring too small");
static_assert((K1_C0 % 2) == 0 && (K1_D0 % 2) == 0 && (K1_PREROLL % 2) == 0, "16-byte aligned ring reads need even offsets");
// reach of the cascade in 240 kS/s samples: interpolator, fir120 and equaliser count double (120 kS/s), + half-band + proto
constexpr int K1_REACH = 2 * (TB_INT_K + TB_FIR_H + TB_REQ_K) + TB_HB_H + (TB_PROTO_H + 9) / 10;
static_assert(K1_REACH < K1_EDGE && K1_REACH < K1_PREROLL, "kept outputs must not see samples outside their block / slot");
static_assert((2 * K1_D0) % 2 == 0, "y pairs must start on even samples");

struct K1Smem {
    float2 in[K1_NBUF][K1_HDR + K1_TILE];
    float2 w[K1_WRING + K1_WPAD];
    float2 u[K1_URING + K1_UPAD];
    float2 v[K1_VRING + K1_VPAD];
    double bins[K1_NPH][K1_DLANES];
    uint64_t full[K1_NBUF];
    int s_item[8];          // work item of slot q at [q & 7] (claimed from the launch's counter, see k1_claim)
    int s_nmy;              // slots of this CTA once the counter has run out (K1_NMY_UNKNOWN before)
};
// freq_offset != 0 variant: the NCO phasors of the w samples of the current / next iteration, the half-band
// output ring and this carrier's equaliser taps
struct K1SmemFo {
    K1Smem base;
    float2 ph[2][K1_W];
    float2 ur[K1_RRING + K1_RPAD];    // MODE 1: half-band output before the equaliser
    float2 rtap[2][16];     // MODE 1: equaliser taps r[-5..5] of the slot (by slot parity)
    float2 ptap[2][48];     // MODE 2: this channel's modulated proto taps
};

// uint8 ingest (MODE 3 / 4): the float tile buffers are not used; their space holds the ring of raw byte tiles (2 bytes per
// sample, each slot behind an 80-byte header with the previous tile's tail) that the bulk copies fill three tiles ahead
constexpr int K1_RAWBUF = 4;
constexpr int K1_RAW_BYTES = 2 * K1_TILE;
constexpr int K1_RAW_HDR = 2 * K1_HDR;                      // bytes of the previous tile kept in front of each raw slot
constexpr int K1_RAW_SLOT = K1_RAW_HDR + K1_RAW_BYTES;      // slot stride: [header | tile]
struct K1SmemU8 {
    K1Smem base;
    uint64_t rawfull[K1_RAWBUF];
};
// uint8 ingest with a freq_offset (MODE 4): both of the above
struct K1SmemFoU8 {
    K1SmemFo fo;
    uint64_t rawfull[K1_RAWBUF];
};
static_assert(K1_RAWBUF * K1_RAW_SLOT <= (int)sizeof(float2) * K1_NBUF * (K1_HDR + K1_TILE), "raw byte ring does not fit the tile buffers");
static_assert(K1_RAW_SLOT % 16 == 0 && K1_RAW_HDR % 16 == 0 && (50 * 2) % 4 == 0, "bulk copies need 16-byte aligned slot bodies, windows start on words");
static_assert(sizeof(K1SmemFo) <= 227 * 1024 && sizeof(K1SmemU8) <= 227 * 1024 && sizeof(K1SmemFoU8) <= 227 * 1024,
              "K1 shared memory exceeds the 227 KB a CTA may use");

__constant__ float c_proto[2 * TB_PROTO_H + 1];
__constant__ float c_proto8[2 * TB_PROTO_H + 1];          // c_proto / 127.5 (byte input)
__constant__ float c_proto8_pre[2 * TB_PROTO_H + 3];      // c_proto8_pre[k] = 127.5 * sum_{d < k} c_proto8[d], k = 0 .. 41
__constant__ float c_hb[2 * TB_HB_H + 1];
__constant__ float c_fir[128];                    // c_fir[0] = 0, c_fir[1 + k] = fir120 tap k (127 taps): v[n] = sum_k c_fir[k] u[n - 64 + k]
__constant__ float c_interp[TB_INT_K];
__constant__ double c_req[(TB_REQ_DEG + 1) * (2 * TB_REQ_K + 1) * 2];   // Chebyshev series of the equaliser taps (taps_generated.h)

struct K1Args {
    const float2* x;        // [C][pitch]
    const uint8_t* x8;      // MODE 3: [C][pitch][2] interleaved unsigned bytes instead of x
    int64_t pitch;
    int64_t n;              // samples per carrier
    int32_t L;              // ceil(n/10)
    int32_t seg_len;        // y samples per segment (multiple of 640)
    int32_t n_seg;          // segments per carrier
    int32_t n_items;        // work items (carrier, segment) of THIS launch
    int32_t item0;          // first work item of this launch (a batch may be cut into several launches, tetra_b200.cu)
    uint32_t* counter;      // zero at launch: the next unclaimed work item of this launch (dynamic slot claims)
    int32_t t_item;         // tiles per item = seg_len / 640 + 1 (pre-roll + post-roll)
    float2* y;              // [C][y_pitch], timing-phase major: y[(n % 13) * y_rows + n / 13]
    int64_t y_pitch;
    int32_t y_rows;
    double* partial;        // [C][n_seg][16]
    int32_t aligned;        // 1: x (x8) base/pitch allow 16-byte bulk copies
    int32_t w_col0;         // MODE 5: x holds channelized 240 kS/s streams [96][pitch] (tetra_pfb.cuh), w index m at column m + w_col0
    int32_t zero_ext;       // 1: the block is extended by zeros and y is written over the whole block (the block-end corrections
                            //    of tetra_edgecorr.cuh are added by the finalize kernel); 0: K1_EDGE outputs at each end are left
                            //    to the exact edge kernels and what lies beyond the block is arbitrary
    // MODE >= 1 only
    const double* fo;       // [C] Hz: freq_offset (MODE 1) / channel offset (MODE 2)
    double fs_dec;          // 240000
    double fs;              // 2.4e6 (MODE 2: the rate the channel offsets refer to)
};

// one quarter (taps 32 q .. 32 q + 31 of the 128-entry table c_fir) of ten consecutive fir120 outputs;
// q is warp-uniform, so the taps arrive through uniform constant loads and there is one copy of this code
__device__ __forceinline__ void k1_fir_quarter(const float2* __restrict__ u, int s0, const float* __restrict__ taps,
                                               float2 (&acc)[10]) {
    static_assert(2 * 21 <= K1_UPAD, "fir120 quarter window exceeds the u ring's pad");
    const float2* __restrict__ up = &u[s0 & (K1_URING - 1)];           // s0 is even: 16-byte aligned
#pragma unroll
    for (int t2 = 0; t2 < 21; ++t2) {
        const float4 v = *reinterpret_cast<const float4*>(&up[2 * t2]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = 2 * t2 + h;
            const float2 xv = h ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                const int k = t - r;               // tap index within the quarter
                if (k >= 0 && k < 32) acc[r] = ffma2(xv, taps[k], acc[r]);
            }
        }
    }
}

__device__ __forceinline__ void k1_bar_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

// ring store: entry idx (unmasked stream index) and, for the first PAD entries, its copy behind the ring's end
template <int RING, int PAD>
__device__ __forceinline__ void k1_ring_store(float2* ring, int idx, float2 v) {
    const int j = idx & (RING - 1);
    ring[j] = v;
    if (j < PAD) ring[j + RING] = v;
}
template <int RING, int PAD>
__device__ __forceinline__ void k1_ring_store2(float2* ring, int idx, float4 v) {     // two entries, idx even
    static_assert((PAD & 1) == 0, "pairs must not straddle the pad");
    const int j = idx & (RING - 1);
    *reinterpret_cast<float4*>(&ring[j]) = v;
    if (j < PAD) *reinterpret_cast<float4*>(&ring[j + RING]) = v;
}

// NCO phasors exp(-j 2 pi f m / fs_dec) (processor.py:97-100 evaluated at the decimated rate). The float64 phase is reduced
// to one turn before the sine/cosine. One out-of-line copy: it runs once per slot and thread, not once per tile.
struct K1Phasor { double c, s; };
struct K1PhasorSet {
    K1Phasor base, step_tile;   // phasor of sample m0; rotation over one tile (640 samples)
    float2 pw[10];              // rotation over g = 0..9 samples, rounded to float32
};
__device__ __forceinline__ K1Phasor k1_rot(K1Phasor p, K1Phasor r) { return K1Phasor{p.c * r.c - p.s * r.s, p.c * r.s + p.s * r.c}; }
__device__ __noinline__ void k1_phasor_setup(double fo, double fs_dec, int64_t m0, K1PhasorSet* out) {
    double turns = fo * (double)m0 / fs_dec;
    sincospi(-2.0 * (turns - rint(turns)), &out->base.s, &out->base.c);
    turns = fo * (double)K1_W / fs_dec;
    sincospi(-2.0 * (turns - rint(turns)), &out->step_tile.s, &out->step_tile.c);
    K1Phasor step1, p = {1.0, 0.0};
    turns = fo / fs_dec;
    sincospi(-2.0 * (turns - rint(turns)), &step1.s, &step1.c);
#pragma unroll
    for (int g = 0; g < 10; ++g) {
        out->pw[g] = make_float2((float)p.c, (float)p.s);
        p = k1_rot(p, step1);
    }
}
// ten consecutive phasors starting at `base`: the float64 base rounded once, times the float32 short rotations
// (|error| < 2e-7, no serial chain)
__device__ __forceinline__ void k1_phasors10(K1Phasor base, const float2 (&pw)[10], float2* dst) {
    const float bc = (float)base.c, bs = (float)base.s;
#pragma unroll
    for (int g = 0; g < 10; ++g)
        dst[g] = make_float2(fmaf(bc, pw[g].x, -bs * pw[g].y), fmaf(bc, pw[g].y, bs * pw[g].x));
}

// Work items and slots. An item is one (carrier, segment); every persistent CTA streams the items it claims back to back
// as ONE continuous sample stream: slot q of the CTA's stream is t_item tiles long (the segment plus 320 w samples of
// pre-roll and post-roll) and the four stages simply keep flowing across slot boundaries -- the pipeline is filled and
// drained once per CTA, not once per item. Only stage A (which tile of which carrier to fetch) and stage D (where an
// output belongs) look at slots.
// Items are claimed from a counter, one slot ahead: thread 0 takes the item of slot q + 1 at the END of iteration
// q t_item + t_item - 6 (the producer first needs it three or four iterations later), and every thread re-reads the slot
// count at the TOP of the next iteration -- one barrier after the claim, t_item - 1 >= 5 barriers before the next one, so all
// threads of the CTA always agree on it. A CTA that shares its SM with other kernels (block-end corrections, finalize)
// or sits on a slower SM simply claims fewer items; with a static assignment the slowest CTA set the kernel's time.
constexpr int K1_NMY_UNKNOWN = 1 << 20;
constexpr int K1_CLAIM_LEAD = 6;
static_assert(K1_MIN_T_ITEM >= K1_CLAIM_LEAD, "a slot must be long enough to claim the next one inside it");
struct K1Slot {
    int car, n_lo, n_hi, O;   // carrier, segment range in y samples, local origin O = n_lo - PREROLL
};
__device__ __forceinline__ K1Slot k1_slot(const K1Args& a, const K1Smem& s, int q) {
    const int it = a.item0 + *reinterpret_cast<const volatile int*>(&s.s_item[q & 7]);
    K1Slot sl;
    sl.car = it / a.n_seg;
    const int seg = it - sl.car * a.n_seg;
    sl.n_lo = seg * a.seg_len;
    sl.n_hi = min(sl.n_lo + a.seg_len, a.L);
    sl.O = sl.n_lo - K1_PREROLL;
    return sl;
}
__device__ __forceinline__ int k1_floordiv(int x, int d) { return x >= 0 ? x / d : -((-x + d - 1) / d); }

// fill buffer (i % NBUF) with tile i of this CTA's stream
__device__ __forceinline__ void k1_issue_stream_tile(K1Smem& s, const K1Args& a, int i, const K1Slot& sl, int t, int lane) {
    const float2* xc = a.x + (int64_t)sl.car * a.pitch;
    const int64_t gx0 = (int64_t)sl.O * 10 + (int64_t)t * K1_TILE;
    float2* dst = &s.in[i % K1_NBUF][K1_HDR];
    uint64_t* bar = &s.full[i % K1_NBUF];
    // in-block part [lo, hi) of the tile. Samples outside the block never reach an output this kernel keeps (those are
    // K1_EDGE away from the block ends, the cascade reaches K1_REACH < K1_EDGE), so the rest of the buffer may hold anything finite.
    const int lo = gx0 < 0 ? (int)min((int64_t)K1_TILE, -gx0) : 0;
    const int hi = (int)max((int64_t)lo, min((int64_t)K1_TILE, a.n - gx0));
    if (a.aligned) {
        const int cnt = (hi - lo) & ~1;                    // bulk copies move multiples of 16 bytes (lo and gx0 are even)
        if (lane == 0) {
            if (cnt > 0) {
                mbar_expect_tx(bar, cnt * 8);
                tma_load_1d(dst + lo, xc + gx0 + lo, cnt * 8, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    } else {
        for (int tt = lo + lane; tt < hi; tt += 32) dst[tt] = __ldg(xc + gx0 + tt);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    }
}

// MODE 5: the 640 w samples of tile i, already channelized (tetra_pfb.cuh), from row (channel offset / 25 kHz + 48)
__device__ __forceinline__ void k1_issue_stream_tile_w(K1Smem& s, const K1Args& a, int i, const K1Slot& sl, int t, int lane) {
    if (lane == 0) {
        const int row = (int)lrint(a.fo[sl.car] * (96.0 / a.fs)) + 48;
        const float2* src = a.x + (int64_t)row * a.pitch + (a.w_col0 + sl.O + K1_W * t + K1_A0);
        uint64_t* bar = &s.full[i % K1_NBUF];
        mbar_expect_tx(bar, K1_W * 8);
        tma_load_1d(&s.in[i % K1_NBUF][K1_HDR], src, K1_W * 8, bar);
    }
}

// MODE 3 / 4: the byte ring lives where the float tile buffers are (stage A filters the bytes as they are)
__device__ __forceinline__ uint8_t* k1_raw_slot(K1Smem& sb, int i) {          // body of slot i % RAWBUF (its header lies in front)
    return reinterpret_cast<uint8_t*>(&sb.in[0][0]) + (i % K1_RAWBUF) * K1_RAW_SLOT + K1_RAW_HDR;
}
// MODE 3: raw byte tile i of this CTA's stream -> slot i % RAWBUF of the byte ring
__device__ __forceinline__ void k1_issue_stream_tile_u8(K1Smem& sb, uint64_t* rawfull, const K1Args& a, int i, const K1Slot& sl, int t, int lane) {
    const uint8_t* xc = a.x8 + 2 * (int64_t)sl.car * a.pitch;
    const int64_t gx0 = (int64_t)sl.O * 10 + (int64_t)t * K1_TILE;        // a multiple of 8 (segments are multiples of 640 outputs)
    uint8_t* dst = k1_raw_slot(sb, i);
    uint64_t* bar = &rawfull[i % K1_RAWBUF];
    const int lo = gx0 < 0 ? (int)min((int64_t)K1_TILE, -gx0) : 0;
    const int hi = (int)max((int64_t)lo, min((int64_t)K1_TILE, a.n - gx0));
    if (a.aligned) {
        const int cnt = (hi - lo) & ~7;                    // 16-byte units; the block's last 30 samples reach no kept output
        if (lane == 0) {
            if (cnt > 0) {
                mbar_expect_tx(bar, cnt * 2);
                tma_load_1d(dst + 2 * lo, xc + 2 * (gx0 + lo), cnt * 2, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    } else {
        for (int tt = 2 * lo + lane; tt < 2 * hi; tt += 32) dst[tt] = __ldg(xc + 2 * gx0 + tt);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    }
}
// MODE 3 / 4: stage A filters the raw bytes. A sample is (byte / 127.5) - 1 per component (pyrtlsdr packed_bytes_to_iq,
// signal/capture.py:143-158), so  sum_d p[d] (b[d] / 127.5 - 1) = sum_d (p[d] / 127.5) b[d] - sum_d p[d]:  the taps are
// scaled once (c_proto8), every output starts from -sum p, and a byte only has to become the float b -- one PRMT puts it
// into the mantissa of 2^23, one packed add per sample takes the 2^23 off again (exact). Nothing is expanded in shared
// memory: 2 bytes per sample are read where the float path reads 8. Samples outside the block are zero in the reference's
// zero-extended stream, i.e. b = 127.5, which no byte holds: their bytes are set to 0 and each output gets back what its
// taps over them took away, 127.5 * sum_{d outside} p8[d] (c_proto8_pre holds the running sums).
// modulated proto taps of a channel at offset f: q[d] = c_proto[d] exp(-j 2 pi f (d - 20) / fs). With them and the w rotation
// exp(-j 2 pi f 10 m / fs) stage A computes proto(x[n] exp(-j 2 pi f n / fs)) without ever forming the shifted stream.
__device__ __forceinline__ float2 k1_modulated_tap(int d, double f, double fs) {
    double sn, cs;
    const double turns = f * (double)(d - TB_PROTO_H) / fs;
    sincospi(-2.0 * (turns - rint(turns)), &sn, &cs);
    return make_float2((float)((double)c_proto[d] * cs), (float)((double)c_proto[d] * sn));
}

// tap k (0..10 <-> lag -5..5) of the freq_offset equaliser: Clenshaw sum of its Chebyshev series in f_off / 12.5 kHz
__device__ __forceinline__ float2 k1_req_tap(int k, double fo) {
    const double x = fo * (1.0 / TB_REQ_FOMAX), x2 = 2.0 * x;
    double br1 = 0.0, br2 = 0.0, bi1 = 0.0, bi2 = 0.0;
#pragma unroll 1
    for (int j = TB_REQ_DEG; j >= 1; --j) {
        const double* c = c_req + (j * (2 * TB_REQ_K + 1) + k) * 2;
        const double tr = x2 * br1 - br2 + c[0], ti = x2 * bi1 - bi2 + c[1];
        br2 = br1; br1 = tr; bi2 = bi1; bi1 = ti;
    }
    return make_float2((float)(x * br1 - br2 + c_req[2 * k]), (float)(x * bi1 - bi2 + c_req[2 * k + 1]));
}

// outputs [R_LO, R_HI) of a thread's five equalised samples ne0 .. ne0 + 4: out[n] = sum_lag r[lag] ur[n - lag], |lag| <= 5,
// from the half-band ring window rp[t] = ur[ne0 - 5 + t]. Stage B's threads take three of the five, stage D's the other two
// (the two roles sit on different SM sub-partitions; the whole equaliser on stage B made its two the busiest of the four).
template <int R_LO, int R_HI>
__device__ __forceinline__ void k1_equalise(const float2* __restrict__ rp, const float2* __restrict__ tp, float2* __restrict__ u_ring, int ne0) {
    float2 e[5];
#pragma unroll
    for (int r = R_LO; r < R_HI; ++r) e[r] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = R_LO; t < R_HI + 2 * TB_REQ_K; ++t) {
        const float2 xv = rp[t];
        const float2 xs = make_float2(-xv.y, xv.x);          // j * x
#pragma unroll
        for (int r = R_LO; r < R_HI; ++r) {
            const int k = r + TB_REQ_K - (t - TB_REQ_K);     // lag = r - (t - 5), tap index lag + 5
            if (k >= 0 && k <= 2 * TB_REQ_K) {
                const float2 c = tp[k];
                e[r] = ffma2(xv, c.x, e[r]);
                e[r] = ffma2(xs, c.y, e[r]);
            }
        }
    }
#pragma unroll
    for (int r = R_LO; r < R_HI; ++r) k1_ring_store<K1_URING, K1_UPAD>(u_ring, ne0 + r, e[r]);
}
#ifndef TETRA_K1_EQ_SPLIT
#define TETRA_K1_EQ_SPLIT 3
#endif
constexpr int K1_EQ_SPLIT = TETRA_K1_EQ_SPLIT;   // stage B equalises outputs 0 .. 2, stage D outputs 3, 4 (5: all on stage B, for A/B builds)

// MODE 0: freq_offset == 0. MODE 1: freq_offset != 0, |f| <= 12.5 kHz: the w samples are rotated by the NCO phasor (stage
// B's warps prepare them one iteration ahead), which leaves proto and the Chebyshev response acting at f + f_off; stage B
// makes up the difference R(f) = [C2(f+f_off)/P(f+f_off)] / [C2(f)/P(f)] with an 11-tap complex equaliser on its half-band
// output (taps: k1_req_tap), and stage C runs the same real fir120 as MODE 0. MODE 2 (config 3): every "carrier" is a channel of ONE
// shared wideband capture (pitch 0) at offset fo[c]: process(frequency_shift(x, f_c), 0) with the shift folded into
// stage A (modulated proto taps + w rotation), so the capture is read as is and never expanded.
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS, 1) k1_channelize_demod(const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    K1Smem& s = *reinterpret_cast<K1Smem*>(smem_raw);
    K1SmemFo& sf = *reinterpret_cast<K1SmemFo*>(smem_raw);
    constexpr bool U8 = MODE == 3 || MODE == 4;           // input rows are RTL-SDR bytes
    constexpr bool FO = MODE == 1 || MODE == 4;           // per-carrier freq_offset: NCO on the w samples + equaliser
    constexpr bool ROT = FO || MODE == 2;                 // the w samples are rotated
    uint64_t* const rawfull = MODE == 4 ? reinterpret_cast<K1SmemFoU8*>(smem_raw)->rawfull : reinterpret_cast<K1SmemU8*>(smem_raw)->rawfull;
    constexpr int NB = K1_NBUF;                           // float tile buffers in rotation (the byte ring takes their place in MODE 3 / 4)
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int S = a.t_item * K1_W;                        // slot length in w / y samples
    // slots of this CTA, tiles of its stream, iterations: every thread keeps its own copy, refreshed by k1_tick at the top of
    // each iteration. The last kept output sits below stream coordinate n_my S - PREROLL; D_i ends at 640 i + 2 D0 + 639.
    constexpr int DRAIN = (-2 * K1_D0 - K1_PREROLL + K1_W - 1) / K1_W;
    int n_my = K1_NMY_UNKNOWN, n_load = K1_NMY_UNKNOWN * a.t_item, n_iter = n_load + DRAIN;
    int rph = a.t_item - (K1_CLAIM_LEAD - 1);             // iterations until the next re-read of the slot count
    auto k1_tick = [&]() {
        if (rph == 0) {
            rph = a.t_item;
            n_my = *reinterpret_cast<const volatile int*>(&s.s_nmy);
            n_load = n_my * a.t_item;
            n_iter = n_load + DRAIN;
        }
        --rph;
    };

    // ---- prologue: zero rings/headers, init barriers, start the first two tiles ----
    for (int i = tid; i < K1_WRING + K1_WPAD; i += K1_THREADS) s.w[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < K1_URING + K1_UPAD; i += K1_THREADS) s.u[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < K1_VRING + K1_VPAD; i += K1_THREADS) s.v[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < K1_NBUF * (K1_HDR + K1_TILE); i += K1_THREADS) (&s.in[0][0])[i] = make_float2(0.f, 0.f);
    if (tid == 0) {
        s.s_item[0] = (int)atomicAdd(a.counter, 1u);      // the grid never exceeds the item count, but see the check below
        s.s_nmy = K1_NMY_UNKNOWN;
        for (int b = 0; b < K1_NBUF; ++b) mbar_init(&s.full[b], 1);
        if (U8) for (int b = 0; b < K1_RAWBUF; ++b) mbar_init(&rawfull[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeroed input buffers are refilled by bulk copies
    __syncthreads();                                      // slot 0's item is known to every thread
    // A CTA that starts late -- its SM was still held by a block-end kernel of the side stream, say -- may find the counter
    // already run out by the other CTAs' look-ahead claims (launches with about as many items as CTAs): nothing is left for it.
    if (*reinterpret_cast<const volatile int*>(&s.s_item[0]) >= a.n_items) return;
    if (ROT) {
        const K1Slot s0 = k1_slot(a, s, 0);
        if (FO && tid < 2 * TB_REQ_K + 1) sf.rtap[0][tid] = k1_req_tap(tid, a.fo[s0.car]);
        if (FO) for (int i = tid; i < K1_RRING + K1_RPAD; i += K1_THREADS) sf.ur[i] = make_float2(0.f, 0.f);
        if (MODE == 2 && tid < 2 * TB_PROTO_H + 1) sf.ptap[0][tid] = k1_modulated_tap(tid, a.fo[s0.car], a.fs);
        if (tid >= 256 && tid < 320) {                    // phasors of iteration 0: w [A0, A0 + 640) of slot 0
            K1PhasorSet ps;
            k1_phasor_setup(a.fo[s0.car], a.fs_dec, (int64_t)s0.O + K1_A0 + 10 * (tid - 256), &ps);
            k1_phasors10(ps.base, ps.pw, &sf.ph[0][10 * (tid - 256)]);
        }
    }
    __syncthreads();
    if (warp == 0) {
        if (U8) {
            for (int j = 0; j < 3; ++j) k1_issue_stream_tile_u8(s, rawfull, a, j, k1_slot(a, s, j / a.t_item), j % a.t_item, lane);
        } else if (MODE == 5) {
            k1_issue_stream_tile_w(s, a, 0, k1_slot(a, s, 0), 0, lane);
            k1_issue_stream_tile_w(s, a, 1, k1_slot(a, s, 1 / a.t_item), 1 % a.t_item, lane);
        } else {
            k1_issue_stream_tile(s, a, 0, k1_slot(a, s, 0), 0, lane);
            k1_issue_stream_tile(s, a, 1, k1_slot(a, s, 1 / a.t_item), 1 % a.t_item, lane);
        }
    }
    __syncthreads();                                      // rows that are not 16-byte aligned are filled by plain stores

    // Each role runs its own loop (own loop-carried registers); all of them meet once per iteration at
    // barrier 0 (bar.sync with the full CTA thread count is well defined from divergent code paths).
    if (warp < 4) {
        // ---------------- role A: proto, 5 outputs per lane ----------------
        // (slot, tile-in-slot) of the tile being filtered and of the tile being fetched (two ahead), kept incrementally
        constexpr int AHEAD = U8 ? 3 : 2;                        // tiles the producer runs ahead of the filter
        int q = 0, t = 0, q2 = AHEAD / a.t_item, t2 = AHEAD % a.t_item;
        K1Slot sl2 = k1_slot(a, s, min(q2, n_my - 1));
        int64_t slot_gx = (int64_t)k1_slot(a, s, 0).O * 10;      // input index of the current slot's first sample
        int q_claim = 1;                                      // thread 0: the next slot to claim an item for
        for (int i = 0; i < n_iter; ++i) {
            k1_tick();
            if (i < n_load) {
                // producer: tile i+2 goes into the buffer tile i-1 just left (its tail is already copied)
                if (U8) {                                      // raw bytes, three tiles ahead into the 4-slot byte ring
                    if (warp == 0 && i + 3 < n_load) k1_issue_stream_tile_u8(s, rawfull, a, i + 3, sl2, t2, lane);
                } else if (MODE == 5) {
                    if (warp == 0 && i + 2 < n_load) k1_issue_stream_tile_w(s, a, i + 2, sl2, t2, lane);
                } else if (warp == 0 && i + 2 < n_load) k1_issue_stream_tile(s, a, i + 2, sl2, t2, lane);
                if (++t2 == a.t_item) { t2 = 0; ++q2; if (q2 < n_my) sl2 = k1_slot(a, s, q2); }
                const int L5 = tid;                            // 0..127
                const int64_t gx0 = slot_gx + (int64_t)t * K1_TILE;
                const int q_now = q;                           // slot of the tile being filtered
                if (++t == a.t_item) { t = 0; ++q; if (q < n_my) slot_gx = (int64_t)k1_slot(a, s, q).O * 10; }
                if (!U8) mbar_wait(&s.full[i % K1_NBUF], (uint32_t)((i / K1_NBUF) & 1));
                if (MODE == 5) {
                    // the proto stage ran in the channelizer: the tile IS this iteration's 640 w samples
                    const float2* wsrc = &s.in[i % NB][K1_HDR + 5 * L5];
                    const int wbase = K1_W * i + K1_A0 + 5 * L5;
#pragma unroll
                    for (int g = 0; g < 5; ++g) k1_ring_store<K1_WRING, K1_WPAD>(s.w, wbase + g, wsrc[g]);
                }
                const bool inside = MODE != 5 && gx0 < a.n && gx0 + K1_TILE > 0;
                if (U8) {
                    mbar_wait(&rawfull[i % K1_RAWBUF], (uint32_t)((i / K1_RAWBUF) & 1));
                    uint8_t* body = k1_raw_slot(s, i);
                    uint32_t* next_hdr = reinterpret_cast<uint32_t*>(k1_raw_slot(s, i + 1) - K1_RAW_HDR);
                    const bool edge_tile = inside && (gx0 < 0 || gx0 + K1_TILE > a.n);
                    if (a.zero_ext && edge_tile) {
                        // a tile that straddles a block end: bytes outside the block become 0 (corrected for below); the block's
                        // last samples a bulk copy of whole 16-byte units left out come by plain loads
                        const int lo = gx0 < 0 ? (int)min((int64_t)K1_TILE, -gx0) : 0;
                        const int hi = (int)min((int64_t)K1_TILE, a.n - gx0);
                        uint16_t* b16 = reinterpret_cast<uint16_t*>(body);
                        if (gx0 <= 0 && L5 < K1_HDR) b16[L5 - K1_HDR] = 0;              // the header precedes the block as well
                        for (int k = L5; k < lo; k += 128) b16[k] = 0;
                        for (int k = hi + L5; k < K1_TILE; k += 128) b16[k] = 0;
                        const int done = a.aligned ? ((hi - lo) & ~7) : (hi - lo);
                        if (L5 < hi - lo - done)
                            b16[lo + done + L5] = __ldg(reinterpret_cast<const uint16_t*>(a.x8 + 2 * ((int64_t)k1_slot(a, s, q_now).car * a.pitch + gx0)) + lo + done + L5);
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
                    if (inside) {
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(body) + 25 * L5 - K1_RAW_HDR / 4;   // sample 50 L5 - 40 of the tile
                        float2 acc[5];
                        const float dc = -c_proto8_pre[2 * TB_PROTO_H + 1];        // -sum p: the "-1" of every sample
#pragma unroll
                        for (int g = 0; g < 5; ++g) acc[g] = make_float2(dc, dc);
                        unsigned long long p_off;
                        asm("mov.b64 %0, {%1, %1};" : "=l"(p_off) : "f"(-8388608.f));
#pragma unroll
                        for (int t2 = 0; t2 < 41; ++t2) {
                            const uint32_t wv = wp[t2];                             // I0 Q0 I1 Q1
                            float2 x0, x1;
                            asm("{\n\t.reg .b32 a0, a1, a2, a3;\n\t.reg .b64 q0, q1;\n\t"
                                "prmt.b32 a0, %4, 0x4B000000, 0x7440;\n\t"
                                "prmt.b32 a1, %4, 0x4B000000, 0x7441;\n\t"
                                "prmt.b32 a2, %4, 0x4B000000, 0x7442;\n\t"
                                "prmt.b32 a3, %4, 0x4B000000, 0x7443;\n\t"
                                "mov.b64 q0, {a0, a1};\n\t"
                                "mov.b64 q1, {a2, a3};\n\t"
                                "add.rn.f32x2 q0, q0, %5;\n\t"
                                "add.rn.f32x2 q1, q1, %5;\n\t"
                                "mov.b64 {%0, %1}, q0;\n\t"
                                "mov.b64 {%2, %3}, q1;\n\t}"
                                : "=f"(x0.x), "=f"(x0.y), "=f"(x1.x), "=f"(x1.y)
                                : "r"(wv), "l"(p_off));
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int tt = 2 * t2 + h;
                                const float2 xv = h ? x1 : x0;
#pragma unroll
                                for (int g = 0; g < 5; ++g) {
                                    const int d = tt - 10 * g;
                                    if (d >= 0 && d <= 40) acc[g] = ffma2(xv, c_proto8[d], acc[g]);
                                }
                            }
                        }
                        if (a.zero_ext && (edge_tile || gx0 == 0)) {
                            // taps over samples outside the block saw byte 0 instead of 127.5
                            const int64_t s0 = gx0 + 50 * L5 - K1_HDR;              // input index of this thread's first window sample
#pragma unroll
                            for (int g = 0; g < 5; ++g) {
                                const int64_t first = s0 + 10 * g;                  // input index under tap 0 of output g
                                const int d0 = (int)min((int64_t)(2 * TB_PROTO_H + 1), max((int64_t)0, -first));          // taps d < d0 lie before the block
                                const int d1 = (int)min((int64_t)(2 * TB_PROTO_H + 1), max((int64_t)0, a.n - first));     // taps d >= d1 lie behind it
                                const float c = c_proto8_pre[d0] + (c_proto8_pre[2 * TB_PROTO_H + 1] - c_proto8_pre[d1]);
                                acc[g].x += c; acc[g].y += c;
                            }
                        }
                        const int wbase = K1_W * i + K1_A0 + 5 * L5;
                        if (ROT) {
#pragma unroll
                            for (int g = 0; g < 5; ++g) {
                                const float2 p = sf.ph[i & 1][5 * L5 + g];
                                acc[g] = make_float2(acc[g].x * p.x - acc[g].y * p.y, acc[g].x * p.y + acc[g].y * p.x);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 5; ++g) k1_ring_store<K1_WRING, K1_WPAD>(s.w, wbase + g, acc[g]);
                        // tail of this tile -> header of the next slot
                        if (L5 < K1_RAW_HDR / 4) next_hdr[L5] = reinterpret_cast<const uint32_t*>(body)[K1_RAW_BYTES / 4 - K1_RAW_HDR / 4 + L5];
                    } else if (a.zero_ext) {                   // a tile entirely outside the block: zeros
                        const int wbase = K1_W * i + K1_A0 + 5 * L5;
#pragma unroll
                        for (int g = 0; g < 5; ++g) k1_ring_store<K1_WRING, K1_WPAD>(s.w, wbase + g, make_float2(0.f, 0.f));
                        if (L5 < K1_RAW_HDR / 4) next_hdr[L5] = 0u;
                    }
                } else {
                if (a.zero_ext && inside && (gx0 < 0 || gx0 + K1_TILE > a.n)) {
                    // a tile that straddles a block end: what lies outside the block becomes zero (the cascade then computes
                    // the shift-invariant response of the zero-extended block, which the block-end corrections refer to)
                    float2* wb = &s.in[i % NB][0];
                    const int lo = gx0 < 0 ? (int)min((int64_t)K1_TILE, -gx0) : 0;
                    const int hi = (int)min((int64_t)K1_TILE, a.n - gx0);
                    const float2 zero = make_float2(0.f, 0.f);
                    if (gx0 < 0 && L5 < K1_HDR) wb[L5] = zero;                  // the header precedes the block as well
                    for (int k = L5; k < lo; k += 128) wb[K1_HDR + k] = zero;
                    for (int k = hi + L5; k < K1_TILE; k += 128) wb[K1_HDR + k] = zero;
                    // a bulk copy moves whole 16-byte units: the block's last sample it left out comes by a plain load
                    if (a.aligned && ((hi - lo) & 1) && L5 == 0) {
                        wb[K1_HDR + hi - 1] = __ldg(a.x + (int64_t)k1_slot(a, s, q_now).car * a.pitch + gx0 + hi - 1);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                if (inside) {
                    const float2* buf = &s.in[i % NB][0];
                    const float4* p4 = reinterpret_cast<const float4*>(buf + 50 * L5);
                    float2 acc[5];
#pragma unroll
                    for (int g = 0; g < 5; ++g) acc[g] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t2 = 0; t2 < 41; ++t2) {
                        const float4 v = p4[t2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int tt = 2 * t2 + h;
                            const float2 xv = h ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                            const float2 xs = make_float2(-xv.y, xv.x);      // j * x (MODE 2)
#pragma unroll
                            for (int g = 0; g < 5; ++g) {
                                const int d = tt - 10 * g;
                                if (d >= 0 && d <= 40) {
                                    if (MODE == 2) {
                                        const float2 tp = sf.ptap[q_now & 1][d];
                                        acc[g] = ffma2(xv, tp.x, acc[g]);
                                        acc[g] = ffma2(xs, tp.y, acc[g]);
                                    } else {
                                        acc[g] = ffma2(xv, c_proto[d], acc[g]);
                                    }
                                }
                            }
                        }
                    }
                    const int wbase = K1_W * i + K1_A0 + 5 * L5;
                    if (ROT) {                              // frequency_shift at the decimated rate (processor.py:259-261)
#pragma unroll
                        for (int g = 0; g < 5; ++g) {
                            const float2 p = sf.ph[i & 1][5 * L5 + g];
                            acc[g] = make_float2(acc[g].x * p.x - acc[g].y * p.y, acc[g].x * p.y + acc[g].y * p.x);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 5; ++g) k1_ring_store<K1_WRING, K1_WPAD>(s.w, wbase + g, acc[g]);
                    // tail of this tile -> header of the next buffer
                    if (L5 < K1_HDR) s.in[(i + 1) % NB][L5] = buf[K1_TILE + L5];
                } else if (MODE != 5 && a.zero_ext) {          // a tile entirely outside the block: zeros
                    const int wbase = K1_W * i + K1_A0 + 5 * L5;
#pragma unroll
                    for (int g = 0; g < 5; ++g) k1_ring_store<K1_WRING, K1_WPAD>(s.w, wbase + g, make_float2(0.f, 0.f));
                    if (L5 < K1_HDR) s.in[(i + 1) % NB][L5] = make_float2(0.f, 0.f);
                }
                }   // float input
            }
            if (tid == 0 && rph == 0 && n_my == K1_NMY_UNKNOWN) {     // the next iteration re-reads the slot count: claim now
                const int it = (int)atomicAdd(a.counter, 1u);
                if (it < a.n_items) s.s_item[q_claim & 7] = it;
                else s.s_nmy = q_claim;
                ++q_claim;
            }
            k1_bar_sync();
        }
    } else if (warp < 8) {
        // ---------------- role C: fir120, 10 outputs per lane, one tap quarter per iteration ----------------
        float2 cacc[10];                                    // outputs of the group in flight
#pragma unroll
        for (int r = 0; r < 10; ++r) cacc[r] = make_float2(0.f, 0.f);
        for (int i = 0; i < n_iter; ++i) {
            k1_tick();
            const int q = (i - (warp - 4)) & 3;             // this warp's group is g = i - q
            const int nu0 = K1_U * (i - q) + K1_C0 + 10 * lane;   // first output (stream v index), even
            const int s0 = nu0 - 64 + 32 * q;               // first input sample of this quarter, even
            if (q == 0) {
#pragma unroll
                for (int r = 0; r < 10; ++r) cacc[r] = make_float2(0.f, 0.f);
            }
            k1_fir_quarter(s.u, s0, c_fir + 32 * q, cacc);
            if (q == 3) {
#pragma unroll
                for (int r = 0; r < 10; r += 2)
                    k1_ring_store2<K1_VRING, K1_VPAD>(s.v, nu0 + r, make_float4(cacc[r].x, cacc[r].y, cacc[r + 1].x, cacc[r + 1].y));
            }
            k1_bar_sync();
        }
    } else if (warp < 10) {
        // ---------------- role B: half-band /2, 5 outputs per lane ----------------
        const int lb = tid - 256;                           // 0..63
        int qn = 1 / a.t_item, tn = 1 % a.t_item;           // (slot, tile) of the tile stage A filters next iteration
        K1Slot sn = k1_slot(a, s, qn);
        // MODE >= 1: phasor of this thread's first w sample of that tile (float64), advanced by one tile per iteration inside
        // a slot (~1e-16 per step over the <= 170 tiles of a slot) and set up afresh when a slot opens
        K1Phasor pbase = {1.0, 0.0}, pstep_tile = {1.0, 0.0};
        float2 pw[10];
#pragma unroll
        for (int g = 0; g < 10; ++g) pw[g] = make_float2(1.f, 0.f);
        // the out-of-line setup writes through memory; the loop keeps plain register copies
        auto open_slot = [&](const K1Slot& sl, int tile) {
            K1PhasorSet ps;
            k1_phasor_setup(a.fo[sl.car], a.fs_dec, (int64_t)sl.O + K1_W * tile + K1_A0 + 10 * lb, &ps);
            pbase = ps.base; pstep_tile = ps.step_tile;
#pragma unroll
            for (int g = 0; g < 10; ++g) pw[g] = ps.pw[g];
        };
        // MODE 1: slot of stream coordinate 640 i + 2 R0 + PREROLL (the kept outputs of the equalised range), kept incrementally
        int qr = k1_floordiv(2 * K1_R0 + K1_PREROLL, S), rr = 2 * K1_R0 + K1_PREROLL - qr * S;
        if (ROT && qn < n_my)
            open_slot(sn, tn);
        for (int i = 0; i < n_iter; ++i) {
            k1_tick();
            const int nu0 = K1_U * i + K1_B0 + 5 * lb;
            float2 acc[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) acc[r] = make_float2(0.f, 0.f);
            const int w0 = 2 * nu0 - TB_HB_H;               // first w sample needed
            static_assert(2 * 4 + 2 * TB_HB_H + 1 <= K1_WPAD, "half-band window exceeds the w ring's pad");
            const float2* __restrict__ wp = &s.w[w0 & (K1_WRING - 1)];
#pragma unroll
            for (int t = 0; t < 2 * 4 + 2 * TB_HB_H + 1; ++t) {
                const float2 xv = wp[t];
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const int d = t - 2 * r;                // tap index 0..22 (centre 11)
                    if (d >= 0 && d <= 2 * TB_HB_H && (d == TB_HB_H || ((d - TB_HB_H) & 1)))
                        acc[r] = ffma2(xv, c_hb[d], acc[r]);
                }
            }
            if (FO) {
                // half-band output -> ring; the carrier's equaliser runs over earlier iterations' entries (no barrier in between)
#pragma unroll
                for (int r = 0; r < 5; ++r) k1_ring_store<K1_RRING, K1_RPAD>(sf.ur, nu0 + r, acc[r]);
                // taps of the slot the kept outputs of this range belong to (kept outputs lie PREROLL inside their slot)
                const float2* tp = sf.rtap[min(max(qr, 0), n_my - 1) & 1];
                const int ne0 = K1_U * i + K1_R0 + 5 * lb;
                static_assert(5 + 2 * TB_REQ_K <= K1_RPAD, "equaliser window exceeds its ring's pad");
                k1_equalise<0, K1_EQ_SPLIT>(&sf.ur[(ne0 - TB_REQ_K) & (K1_RRING - 1)], tp, s.u, ne0);
                rr += K1_W;
                if (rr >= S) { rr -= S; ++qr; }
            } else {
#pragma unroll
                for (int r = 0; r < 5; ++r) k1_ring_store<K1_URING, K1_UPAD>(s.u, nu0 + r, acc[r]);
            }
            if (ROT && i + 1 < n_load) {
                // for the tile stage A filters next iteration: the NCO phasors of its w samples and, when it opens
                // a new slot, that carrier's taps (the buffer's previous owner left stage C a whole slot ago)
                k1_phasors10(pbase, pw, &sf.ph[(i + 1) & 1][10 * lb]);
                if (tn == 0 && FO && lb < 2 * TB_REQ_K + 1) sf.rtap[qn & 1][lb] = k1_req_tap(lb, a.fo[sn.car]);
                if (tn == 0 && MODE == 2 && lb < 2 * TB_PROTO_H + 1) sf.ptap[qn & 1][lb] = k1_modulated_tap(lb, a.fo[sn.car], a.fs);
                if (++tn == a.t_item) {
                    tn = 0; ++qn;
                    if (qn < n_my) {
                        sn = k1_slot(a, s, qn);
                        open_slot(sn, 0);
                    }
                } else {
                    pbase = k1_rot(pbase, pstep_tile);
                }
            }
            k1_bar_sync();
        }
    } else {
        // ---------------- role D: x2 interpolation, store, power sums per timing phase ----------------
        const int ld = tid - 320;                           // 0..63
        double pacc[K1_NPH];                                // |y|^2 sums, pacc[j] <-> phase (n0 + j) % 13
#pragma unroll
        for (int j = 0; j < K1_NPH; ++j) pacc[j] = 0.0;
        int q_cur = -1;                                     // slot the sums belong to (-1: none yet)
        K1Slot sl = {0, 0, 0, 0};
        int y_lo = 0, y_hi = 0;                             // outputs whose power this kernel sums
        int st_lo = 0, st_hi = 0;                           // outputs this kernel stores
        float2* yc = a.y;
        // hand the sums of slot q_cur to `partial`; i = iteration about to run (the rotations so far make slot j of
        // pacc mean phase (n0 + j) % 13 with n0 this thread's first output of iteration i under the OLD slot)
        auto flush = [&](int i) {
            const int n0 = K1_W * i + 2 * K1_D0 + 10 * ld - q_cur * S + sl.O;
            const int base = ((n0 % K1_NPH) + K1_NPH) % K1_NPH;
#pragma unroll
            for (int j = 0; j < K1_NPH; ++j) s.bins[(base + j) % K1_NPH][ld] = pacc[j];
            asm volatile("bar.sync 2, 64;" ::: "memory");
            if (ld < K1_NPH) {
                double t = 0.0;
                for (int j = 0; j < K1_DLANES; ++j) t += s.bins[ld][j];
                a.partial[((int64_t)a.item0 + *reinterpret_cast<const volatile int*>(&s.s_item[q_cur & 7])) * 16 + ld] = t;
            }
            asm volatile("bar.sync 2, 64;" ::: "memory");
#pragma unroll
            for (int j = 0; j < K1_NPH; ++j) pacc[j] = 0.0;
        };
        // every output an iteration keeps lies in the slot of stream coordinate 640 i + 2 D0 + PREROLL (kept outputs
        // are PREROLL away from slot ends); that slot index is kept incrementally
        int q_it = k1_floordiv(2 * K1_D0 + K1_PREROLL, S), r_it = 2 * K1_D0 + K1_PREROLL - q_it * S;
        // MODE 1 / 4: this role's share of the equaliser (the slot of the equalised range is tracked as stage B tracks it)
        int qr = k1_floordiv(2 * K1_R0 + K1_PREROLL, S), rr = 2 * K1_R0 + K1_PREROLL - qr * S;
        for (int i = 0; i < n_iter; ++i) {
            k1_tick();
            if (FO) {
                const float2* tp = sf.rtap[min(max(qr, 0), n_my - 1) & 1];
                const int ne0 = K1_U * i + K1_R0 + 5 * ld;
                k1_equalise<K1_EQ_SPLIT, 5>(&sf.ur[(ne0 - TB_REQ_K) & (K1_RRING - 1)], tp, s.u, ne0);
                rr += K1_W;
                if (rr >= S) { rr -= S; ++qr; }
            }
            if (q_it != q_cur) {
                if (q_cur >= 0 && q_cur < n_my) flush(i);
                q_cur = q_it;
                if (q_cur >= 0 && q_cur < n_my) {
                    sl = k1_slot(a, s, q_cur);
                    y_lo = max(sl.n_lo, K1_EDGE); y_hi = min(sl.n_hi, a.L - K1_EDGE);
                    // with the zero extension the whole block is stored (the K1_EDGE outputs at each end still lack their
                    // correction, so their power is left to the finalize kernel)
                    st_lo = a.zero_ext ? sl.n_lo : y_lo; st_hi = a.zero_ext ? sl.n_hi : y_hi;
                    yc = a.y + (int64_t)sl.car * a.y_pitch;
                } else {
                    y_lo = y_hi = st_lo = st_hi = 0;
                }
            }
            const int nu0 = K1_U * i + K1_D0 + 5 * ld;      // stream v index of the first output pair
            const int n0 = 2 * nu0 - q_cur * S + sl.O;      // y index of the first output within the carrier (even)
            if (n0 + 10 > st_lo && n0 < st_hi) {
                float2 v[5 + 2 * TB_INT_K - 1];             // v[nu0-7 .. nu0+4+8]
                static_assert(5 + 2 * TB_INT_K - 1 <= K1_VPAD, "interpolator window exceeds the v ring's pad");
                const float2* __restrict__ vp = &s.v[(nu0 - (TB_INT_K - 1)) & (K1_VRING - 1)];
#pragma unroll
                for (int t = 0; t < 5 + 2 * TB_INT_K - 1; ++t) v[t] = vp[t];
                float2 yv[10];
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    yv[2 * r] = v[r + TB_INT_K - 1];
                    float2 o = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < TB_INT_K; ++k)
                        o = ffma2(fadd2(v[r + TB_INT_K - 1 - k], v[r + TB_INT_K + k]), c_interp[k], o);
                    yv[2 * r + 1] = o;
                }
                // sample n goes to row n % 13, column n / 13 (kept incrementally over the ten outputs)
                int col = n0 / K1_NPH, row = n0 - col * K1_NPH;
                if (n0 >= y_lo && n0 + 10 <= y_hi) {        // the common case: all ten inside
#pragma unroll
                    for (int r = 0; r < 10; ++r) {
                        yc[(int64_t)row * a.y_rows + col] = yv[r];
                        if (++row == K1_NPH) { row = 0; ++col; }
                        // |y|^2 in fp32 (y itself is fp32), accumulated in fp64; slot r <-> phase (n0 + r) % 13
                        pacc[r] += (double)fmaf(yv[r].x, yv[r].x, yv[r].y * yv[r].y);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 10; ++r) {
                        const int n = n0 + r;
                        if (n >= st_lo && n < st_hi) yc[(int64_t)row * a.y_rows + col] = yv[r];
                        if (n >= y_lo && n < y_hi) pacc[r] += (double)fmaf(yv[r].x, yv[r].x, yv[r].y * yv[r].y);
                        if (++row == K1_NPH) { row = 0; ++col; }
                    }
                }
            }
            // next iteration n0 advances by 640 = 3 (mod 13): rotate so that slot j keeps meaning phase (n0 + j) % 13
            {
                const double t0 = pacc[0], t1 = pacc[1], t2 = pacc[2];
#pragma unroll
                for (int j = 0; j < K1_NPH - 3; ++j) pacc[j] = pacc[j + 3];
                pacc[10] = t0; pacc[11] = t1; pacc[12] = t2;
            }
            r_it += K1_W;
            if (r_it >= S) { r_it -= S; ++q_it; }
            k1_bar_sync();
        }
        if (q_cur >= 0 && q_cur < n_my) flush(n_iter);
    }
}

}  // namespace tetra
