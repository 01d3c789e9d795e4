// KF and what follows the dibits: timing pick + differential slicer (processor.py:168-219, 102-166), bit expansion and the
// TS1/TS2 correlator (decoder.py:140-169, 231-259), find_sync + decode()'s cascade (decoder.py:171-295, 845-856) and the
// burst verdicts of parse_burst (protocol.py:192-347).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tetra_exact.cuh"
#include "tetra_gather.cuh"

namespace tetra {

// ----------------------------------------------------------------------------------------------
// K_finalize: timing pick (processor.py:189-215) + soft symbols + differential slicer (:129-163)
// ----------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
constexpr int FIN_MAXPH = 32;
constexpr int FIN_B = 4;                  // symbols per thread and batch in k_finalize

struct FinArgs {
    const float2* y;         // [C][y_pitch] filtered samples at the decimated rate, layout y_index(n, sps, y_rows)
    int64_t y_pitch;
    int32_t y_rows;
    int32_t L;
    int32_t sps, step;       // samples per symbol, phase search step
    const double* partial;   // [C][n_seg][16] power sums of the bulk kernel (or null)
    int32_t n_seg;
    int32_t bulk_lo, bulk_hi;  // y range already covered by `partial` (empty if bulk_lo >= bulk_hi)
    const float2* edge_corr; // [C][2][K_EDGE] or null: block-end corrections (tetra_edgecorr.cuh) to add to y[m], m < K_EDGE,
                             // and y[L-1-t], t < K_EDGE, wherever they are read
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
    int32_t car0;            // first carrier of this launch (block b works on carrier car0 + b)
    // the exchange fused behind the slicer (tetra_process_batch_allgather): every CTA packs its carrier's dibits four to a byte
    // and stores them, with the stream length, into slot `rank` of every peer's receive buffer; the last CTA of the batch
    // publishes the step (see tetra_gather.cuh). push.world == 0: off.
    struct Push {
        uint8_t* recv[KG_MAX_WORLD];
        int32_t world, rank, n_local;
        int64_t slot_off;    // byte offset of this rank's slot of this step's half of the receive buffers
        int64_t flag_off;
        uint32_t step;
        uint32_t* ticket;    // local: carriers of this batch whose push is complete
    } push;
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

// the same walk by the 32 lanes of one warp, for masks in global memory: empty stretches are scanned 32 words per load
__device__ inline int sync_walk_warp(const uint32_t* mask, int nw, int32_t* pos, int max_pos) {
    const int lane = threadIdx.x & 31;
    const int n_words = (nw + 31) >> 5;
    int n = 0, i = 0;
    while (i < nw) {
        int wd = i >> 5;
        uint32_t m = mask[wd] & (0xFFFFFFFFu << (i & 31));
        if (m == 0) {
            int found = -1;
            for (int base = wd + 1; base < n_words; base += 32) {
                const int w = base + lane;
                const uint32_t v = w < n_words ? mask[w] : 0u;
                const unsigned b = __ballot_sync(0xffffffffu, v != 0u);
                if (b) {
                    const int l = __ffs(b) - 1;
                    found = base + l;
                    m = __shfl_sync(0xffffffffu, v, l);
                    break;
                }
            }
            if (found < 0) break;
            wd = found;
        }
        const int p = (wd << 5) + __ffs(m) - 1;
        if (p >= nw) break;
        if (lane == 0 && n < max_pos) pos[n] = p;
        ++n;
        i = p + 250;
    }
    return n;
}

// find_sync(bits, threshold) -> number of positions (written to pos[], global or shared), *max_corr
__device__ int block_find_sync(const uint32_t* __restrict__ bits, int nw, double thr, int32_t* pos, int max_pos,
                               SyncScratch& sc, double* max_corr, uint32_t* mask) {
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
        mask[wd] = m;
    }
    if (tid == 0) sc.max_cnt = 0;
    __syncthreads();
    const bool global_mask = mask != sc.mask;
    if (global_mask) {
        if (tid < 32) { const int np = sync_walk_warp(mask, nw, pos, max_pos); if (tid == 0) sc.n_pos = np; }
    } else if (tid == 0) sc.n_pos = sync_walk(mask, nw, pos, max_pos);
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
                mask[wd] = m;
            }
            __syncthreads();
            // accepted offsets block +-250 around them; scanning upwards that is the same jump-250 walk
            if (global_mask) {
                if (tid < 32) { const int np = sync_walk_warp(mask, nw, pos, max_pos); if (tid == 0) sc.n_pos = np; }
            } else if (tid == 0) sc.n_pos = sync_walk(mask, nw, pos, max_pos);
            __syncthreads();
            n = sc.n_pos;
        }
    }
    __syncthreads();
    return n;
}

// decode()'s cascade 0.90 -> 0.85 -> 0.80 -> adaptive (decoder.py:845-856)
// `mask`: one bit per window start -- the scratch's own array (blocks up to FIN_DIB_SMEM dibits) or global memory (longer ones)
__device__ int block_sync_cascade(const uint32_t* __restrict__ bits, int nd, int32_t* pos, int max_pos, SyncScratch& sc,
                                  uint32_t* mask = nullptr) {
    if (!mask) mask = sc.mask;
    const int nw = 2 * nd - 22 + 1;
    if (nw <= 0) {                                      // decoder.py:226-228: fewer than 22 bits
        for (int i = threadIdx.x; i < max_pos; i += FIN_THREADS) pos[i] = 0;
        return 0;
    }
    double mx = 0.0;
    int n = block_find_sync(bits, nw, 0.90, pos, max_pos, sc, &mx, mask);
    if (n == 0) n = block_find_sync(bits, nw, 0.85, pos, max_pos, sc, &mx, mask);
    if (n == 0) n = block_find_sync(bits, nw, 0.80, pos, max_pos, sc, &mx, mask);
    if (n == 0 && mx >= 0.75) n = block_find_sync(bits, nw, fmax(0.75, mx - 0.02), pos, max_pos, sc, &mx, mask);
    for (int i = n + (int)threadIdx.x; i < max_pos; i += FIN_THREADS) pos[i] = 0;      // unused entries read 0, whatever the buffer held
    return n;
}

// dibits -> MSB-first packed bits (decoder.py:140-169), one zero word behind. Four dibits in the four bytes of a 32-bit word
// (lowest address first) become one byte, first dibit highest.
__device__ __forceinline__ uint32_t pack4_msb(uint32_t x) {
    const uint32_t t = x & 0x03030303u;
    return ((t << 6) | (t >> 4) | (t >> 14) | (t >> 24)) & 0xFFu;
}
// s_dib: 16-byte aligned shared memory (entries behind nd may hold anything)
__device__ __forceinline__ void pack_dibits(const uint8_t* s_dib, int nd, uint32_t* s_bits) {
    const int n_words = (nd + 15) / 16 + 1;
    for (int j = threadIdx.x; j < n_words; j += blockDim.x) {
        const int valid = nd - 16 * j;                      // dibits of this word that belong to the stream
        uint32_t w = 0;
        if (valid > 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_dib + 16 * j);
            w = (pack4_msb(v.x) << 24) | (pack4_msb(v.y) << 16) | (pack4_msb(v.z) << 8) | pack4_msb(v.w);
            if (valid < 16) w &= ~((1u << (2 * (16 - valid))) - 1u);
        }
        s_bits[j] = w;
    }
}
// the same from global memory with any alignment (k_sync_positions_long)
__device__ __forceinline__ void pack_dibits_bytes(const uint8_t* dib, int nd, uint32_t* bits) {
    const int n_words = (nd + 15) / 16 + 1;
    for (int j = threadIdx.x; j < n_words; j += blockDim.x) {
        uint32_t w = 0;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int idx = 16 * j + m;
            w = (w << 2) | (idx < nd ? (uint32_t)(dib[idx] & 3u) : 0u);
        }
        bits[j] = w;
    }
}

struct FinSmem {
    double red[FIN_THREADS];
    int s_best;
    __align__(16) uint8_t s_dib[FIN_DIB_SMEM];
    uint32_t s_bits[FIN_DIB_SMEM / 16 + 2];
    SyncScratch s_sync;
};

// everything k_finalize does for one carrier, by the FIN_THREADS threads of a CTA. y, partial and the corrections are read
// once: L2 loads (ld.global.cg)
template <bool PREFETCH>
__device__ __forceinline__ void finalize_carrier(const FinArgs& a, const int car, FinSmem& sm) {
    double* red = sm.red;
    int& s_best = sm.s_best;
    uint8_t* s_dib = sm.s_dib;
    uint32_t* s_bits = sm.s_bits;
    SyncScratch& s_sync = sm.s_sync;
    const int tid = threadIdx.x;
    const float2* y = a.y + (int64_t)car * a.y_pitch;
    const int L = a.L, sps = a.sps, step = a.step;
    const int nph = (sps + step - 1) / step;            // phases tried: 0, step, 2 step, ... (<= FIN_MAXPH)
    const float2* ec = a.edge_corr ? a.edge_corr + (int64_t)car * 2 * K_EDGE : nullptr;
    // sample n of the filtered stream: the fused kernel's output plus, next to the block ends, the correction
    auto y_at = [&](int n, float2 v) {
        if (ec) {
            if (n < K_EDGE) { const float2 d = __ldcg(ec + n); v.x += d.x; v.y += d.y; }
            if (n >= L - K_EDGE) { const float2 d = __ldcg(ec + K_EDGE + (L - 1 - n)); v.x += d.x; v.y += d.y; }
        }
        return v;
    };
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
                const float2 v = y_at(ph + sps * k, __ldcg(y + y_index(ph + sps * k, sps, a.y_rows)));
                acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            }
            for (int k = k_hi_beg + g; k < cnt; k += G) {
                const float2 v = y_at(ph + sps * k, __ldcg(y + y_index(ph + sps * k, sps, a.y_rows)));
                acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            }
        }
        red[tid] = acc;
        __syncthreads();
        // one thread per phase adds up its partial sums (same order as before), then thread 0 takes the first maximum
        double mean = -2.0;
        if (tid < nph) {
            const int ph = tid * step;
            const int cnt = (L - ph) / sps;
            if (cnt > 0) {
                double sum = 0.0;
                for (int gg = 0; gg < G; ++gg) sum += red[gg * nph + tid];
                if (a.partial && has_bulk)
                    for (int sg = 0; sg < a.n_seg; ++sg) sum += __ldcg(a.partial + ((int64_t)car * a.n_seg + sg) * 16 + ph);
                mean = sum / (double)cnt;
            }
        }
        __syncthreads();
        if (tid < nph) red[tid] = mean;
        __syncthreads();
        if (tid == 0) {
            double best_pow = -1.0;
            for (int pp = 0; pp < nph; ++pp)
                if (red[pp] > best_pow) { best_pow = red[pp]; best = pp * step; }      // phases without a symbol hold -2
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
    // The block-end corrections touch the first and last few symbols only (sample n = best + stride k with n < K_EDGE or
    // n >= L - K_EDGE): those symbols are prepared in shared memory first, so that the main loop's loads are unconditional
    // and a batch's loads all issue before the first use.
    const int k_head = ec ? min(n_sym, (K_EDGE - 1 - best) / stride + 1) : 0;            // symbols k < k_head need D_left
    const int k_tail = ec ? max(k_head, (L - K_EDGE - best + stride - 1) / stride) : n_sym;   // symbols k >= k_tail need D_right
    float2* s_fix = reinterpret_cast<float2*>(sm.s_sync.mask);                              // free until the sync front end
    const int n_fix_tail = n_sym - min(k_tail, n_sym);
    if (ec) {
        for (int i = tid; i < k_head + n_fix_tail; i += FIN_THREADS) {
            const int k = i < k_head ? i : k_tail + (i - k_head);
            s_fix[i] = y_at(best + stride * k, __ldcg(ys + ks * k));
        }
        __syncthreads();
    }
    // symbol k as the slicer sees it: the loaded sample, or its corrected copy next to a block end
    auto fixed = [&](int k, float2 v) {
        if (k < k_head) v = s_fix[k];
        else if (k >= k_tail) v = s_fix[k_head + (k - k_tail)];
        return v;
    };
    // a batch's loads: symbol k of every slot of this thread and, on lane 0, its predecessor (the other lanes get theirs from
    // the neighbouring lane by shuffle). Every warp runs the same number of batches.
    auto fetch = [&](int base, float2 (&d1)[FIN_B], float2 (&d0)[FIN_B]) {
#pragma unroll
        for (int j = 0; j < FIN_B; ++j) {
            const int k = min(base + tid + j * FIN_THREADS, n_sym - 1);
            d1[j] = __ldcg(ys + ks * k);
            if ((tid & 31) == 0) d0[j] = __ldcg(ys + ks * max(k - 1, 0));
        }
    };
    // PREFETCH: the loads of batch b + 1 are issued before batch b is sliced (a CTA has eight batches of dependent load ->
    // slice -> store rounds per carrier; without the prefetch the load latency of every round is exposed)
    float2 nx1[FIN_B], nx0[FIN_B];
    if (PREFETCH && n_sym > 0) fetch(0, nx1, nx0);
    for (int base = 0; base < n_sym; base += FIN_B * FIN_THREADS) {
        float2 s1[FIN_B], s0[FIN_B];
        if (!PREFETCH) fetch(base, nx1, nx0);
        // only the first and the last batches hold symbols next to a block end (CTA-uniform)
        const bool end_batch = base <= k_head || base + FIN_B * FIN_THREADS >= k_tail;
#pragma unroll
        for (int j = 0; j < FIN_B; ++j) {
            s1[j] = nx1[j];
            if ((tid & 31) == 0) s0[j] = nx0[j];
        }
        if (end_batch) {
#pragma unroll
            for (int j = 0; j < FIN_B; ++j) {
                const int k = min(base + tid + j * FIN_THREADS, n_sym - 1);
                s1[j] = fixed(k, s1[j]);
                if ((tid & 31) == 0) s0[j] = fixed(max(k - 1, 0), s0[j]);
            }
        }
        if (PREFETCH && base + FIN_B * FIN_THREADS < n_sym) fetch(base + FIN_B * FIN_THREADS, nx1, nx0);
#pragma unroll
        for (int j = 0; j < FIN_B; ++j) {
            const float px = __shfl_up_sync(0xffffffffu, s1[j].x, 1), py = __shfl_up_sync(0xffffffffu, s1[j].y, 1);
            if ((tid & 31) != 0) s0[j] = make_float2(px, py);
            const int k = base + tid + j * FIN_THREADS;
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
    if (fuse) {
    // ---- fused frame-sync front end (decoder.py:140-169 bit expansion, :237-240 agreement counts, :171-295 + :845-856) ----
    __syncthreads();
    pack_dibits(s_dib, nd, s_bits);
    __syncthreads();
    const int nw = 2 * nd - 22 + 1;                     // window starts (decoder.py:232)
    if (a.match) {
        uint8_t* out = a.match + (int64_t)car * a.cap * 4;
        for (int p = tid; 4 * p < nw; p += FIN_THREADS) {   // windows 4p .. 4p+3 share their words
            const int i = 4 * p;
            const uint64_t two = ((uint64_t)s_bits[i >> 5] << 32) | s_bits[(i >> 5) + 1];
            const int sh = i & 31;                          // a multiple of 4, <= 28: 25 bits starting at sh fit in 64
            uint32_t c[4];                                  // c[m]: TS1 count | TS2 count << 8 of window i + m
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const uint32_t w = (uint32_t)(two >> (64 - 22 - m - sh)) & 0x3FFFFFu;
                c[m] = (uint32_t)(22 - __popc(w ^ TS1_BITS)) | ((uint32_t)(22 - __popc(w ^ TS2_BITS)) << 8);
            }
            uint8_t* o = out + 2 * (int64_t)i;              // 4-byte aligned: rows are multiples of 4 bytes, i of 4
            if (i + 3 < nw) {
                *reinterpret_cast<uint32_t*>(o) = c[0] | (c[1] << 16);
                *reinterpret_cast<uint32_t*>(o + 4) = c[2] | (c[3] << 16);
            } else {
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    if (i + m < nw) *reinterpret_cast<uint16_t*>(o + 2 * m) = (uint16_t)c[m];
            }
        }
    }
    if (a.sync_pos) {
        const int n = block_sync_cascade(s_bits, nd, a.sync_pos + (int64_t)car * a.max_pos, a.max_pos, s_sync);
        if (tid == 0) a.n_sync[car] = n;
    }
    }   // fuse
    // ---- the exchange, fused: this carrier's packed stream and its length go to every peer ----
    if (a.push.world > 0) {
        __syncthreads();                                    // this CTA's dibits are in global memory (its own stores)
        const int words = (int)(a.cap / 16);               // cap is a multiple of 16 here
        const int64_t row = a.push.slot_off + (int64_t)(car - 0) * (a.cap / 4);
        for (int j = tid; j < words; j += FIN_THREADS) {
            const uint4 v = *reinterpret_cast<const uint4*>(dib + 16 * j);
            uint32_t w = kg_pack4(v.x) | (kg_pack4(v.y) << 8) | (kg_pack4(v.z) << 16) | (kg_pack4(v.w) << 24);
            const int valid = nd - 16 * j;                  // dibits of this word that belong to the stream
            if (valid < 16) w = valid <= 0 ? 0u : (w & ((1u << (2 * valid)) - 1u));
            for (int p = 0; p < a.push.world; ++p) *reinterpret_cast<uint32_t*>(a.push.recv[p] + row + 4 * j) = w;
        }
        if (tid == 0) {
            const int64_t off = a.push.slot_off + (int64_t)a.push.n_local * (a.cap / 4) + 4 * (int64_t)(car - 0);
            for (int p = 0; p < a.push.world; ++p) *reinterpret_cast<int32_t*>(a.push.recv[p] + off) = nd;
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            if (atomicAdd(a.push.ticket, 1u) == (uint32_t)a.push.n_local - 1u) {
                __threadfence_system();
                for (int p = 0; p < a.push.world; ++p)
                    kg_st_release_sys(reinterpret_cast<uint32_t*>(a.push.recv[p] + a.push.flag_off) + (a.push.step & 1) * KG_MAX_WORLD + a.push.rank, a.push.step);
                *a.push.ticket = 0u;
            }
        }
    }
}

// PREFETCH: 64 registers / 4 CTAs per SM with the next batch's loads in flight; without: 48 registers / 5 CTAs per SM
template <bool PREFETCH>
__global__ void __launch_bounds__(FIN_THREADS, PREFETCH ? 4 : 5) k_finalize(const FinArgs a) {
    __shared__ FinSmem sm;
    finalize_carrier<PREFETCH>(a, a.car0 + blockIdx.x, sm);
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

// the same for blocks of more than FIN_DIB_SMEM dibits (a recording processed in one call): the packed bits and the hit
// mask of a carrier live in global scratch ([C][words] each, words = cap / 16 + 2) instead of shared memory
struct SyncPosLongArgs {
    const uint8_t* dibits; int64_t cap; const int32_t* n_dibits;
    int32_t* sync_pos; int32_t max_pos; int32_t* n_sync;
    uint32_t* bits; uint32_t* mask; int64_t words;
};
__global__ void __launch_bounds__(FIN_THREADS) k_sync_positions_long(const SyncPosLongArgs a) {
    __shared__ SyncScratch s_sync;
    const int car = blockIdx.x;
    const int nd = (int)min((int64_t)a.n_dibits[car], a.cap);
    uint32_t* bits = a.bits + (int64_t)car * a.words;
    pack_dibits_bytes(a.dibits + (int64_t)car * a.cap, nd, bits);
    __syncthreads();                                     // the CTA's own global stores, visible to all of its threads
    const int n = block_sync_cascade(bits, nd, a.sync_pos + (int64_t)car * a.max_pos, a.max_pos, s_sync, a.mask + (int64_t)car * a.words);
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

// ----------------------------------------------------------------------------------------------
// Transport packing of dibit streams (values 0..3) for the one collective of the path: four dibits per byte,
//   packed[k] = d[4k] | d[4k+1] << 2 | d[4k+2] << 4 | d[4k+3] << 6.
// A thread turns 16 dibits (one 16-byte load) into one 32-bit word and back.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return (w | (w >> 6) | (w >> 12) | (w >> 18)) & 0xFFu; }
__device__ __forceinline__ uint32_t unpack4(uint32_t b) { return (b & 3u) | ((b & 0xCu) << 6) | ((b & 0x30u) << 12) | ((b & 0xC0u) << 18); }

__global__ void __launch_bounds__(256) k_pack_dibits(const uint4* __restrict__ in, int64_t n16, uint32_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(in + i);
        out[i] = pack4(v.x) | (pack4(v.y) << 8) | (pack4(v.z) << 16) | (pack4(v.w) << 24);
    }
}
// blocks of `words` packed 32-bit words at a stride of in_stride bytes -> 16 * words dibits per block at out_stride bytes
__global__ void __launch_bounds__(256) k_unpack_dibits(const uint8_t* __restrict__ in, int64_t words, int64_t in_stride,
                                                       uint8_t* __restrict__ out, int64_t out_stride) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (int64_t)blockIdx.y * in_stride);
    uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)blockIdx.y * out_stride);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldg(src + i);
        dst[i] = make_uint4(unpack4(w & 0xFFu), unpack4((w >> 8) & 0xFFu), unpack4((w >> 16) & 0xFFu), unpack4(w >> 24));
    }
}

}  // namespace tetra
