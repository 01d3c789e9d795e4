// KG: the one exchange of the path -- every rank ends up with the dibit streams of all carriers (BASELINE north star:
// "a single all-gather of decoded dibit streams over NVLink"; the reference itself is single-process, its hand-off is
// TetraDecoder.decode(symbols), tetraear/core/decoder.py:835) -- written as two kernels over NVLink peer memory instead of
// pack -> ncclAllGather -> unpack:
//   k_gather_push         packs this rank's streams four dibits to a byte (the transport format of k_pack_dibits) and stores
//                         the words straight into slot `rank` of EVERY peer's receive buffer (remote stores through
//                         NVSwitch; the local GPU is one of the peers), appends the stream lengths, and -- last CTA out --
//                         publishes the step number in every peer's flag word with a system-scope release;
//   k_gather_wait_unpack  blockIdx.y = source rank r: waits (bounded) until r's flag shows this step, then expands r's
//                         block into the caller's [world][n_local][cap] dibits. A rank's block is unpacked as soon as
//                         it has landed; there is no barrier over all ranks.
// Receive buffers are double-buffered by step parity: a rank can be at most one step ahead of a peer (its push of step
// s + 2 follows its own wait for every peer's flag s + 1, and a peer raises that flag only after unpacking step s), so
// a slot is never overwritten while it is still being read. The buffers come from cudaMalloc and are opened in the peers
// with CUDA IPC (one process per GPU) or handed over as plain pointers (one process, several contexts).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tetra {

constexpr int KG_MAX_WORLD = 8;
constexpr int KG_THREADS = 256;
constexpr long long KG_SPIN_CLOCKS = 6000000000LL;      // ~3 s at 1.9 GHz: a peer that never arrives is reported, not waited for

struct GatherArgs {
    uint8_t* recv[KG_MAX_WORLD];     // receive buffer of every rank (recv[rank] is the local one): [2][world][block] then flags
    int32_t rank, world;
    int64_t block;                   // bytes per rank: packed dibits (n / 4) followed by n_local int32 lengths, 16-byte multiple
    int64_t flag_off;                // byte offset of the flag words [2][KG_MAX_WORLD] uint32 inside a receive buffer
    uint32_t step;                   // 1, 2, 3, ...
    const uint8_t* dibits;           // [n] local streams, n = n_local * cap, cap a multiple of 16
    int64_t n;
    const int32_t* n_dibits;         // [n_local]
    int32_t n_local;
    uint32_t* ticket;                // local: CTAs of the push kernel that have finished
    uint8_t* out;                    // [world][n] unpacked streams of every rank
    int32_t* out_n;                  // [world][n_local] or null
    int32_t* status;                 // local: set to 1 + r when rank r's block did not arrive in time
};

__device__ __forceinline__ uint32_t kg_pack4(uint32_t w) { return (w | (w >> 6) | (w >> 12) | (w >> 18)) & 0xFFu; }
__device__ __forceinline__ uint32_t kg_unpack4(uint32_t b) { return (b & 3u) | ((b & 0xCu) << 6) | ((b & 0x30u) << 12) | ((b & 0xC0u) << 18); }
__device__ __forceinline__ void kg_st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t kg_ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(KG_THREADS) k_gather_push(const GatherArgs a) {
    const int64_t slot = ((int64_t)(a.step & 1) * a.world + a.rank) * a.block;
    // 64 dibits (four 16-byte loads) -> four packed words -> one 16-byte store per peer
    const int64_t n64 = a.n / 64;
    const uint4* in = reinterpret_cast<const uint4*>(a.dibits);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n64; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = __ldg(in + 4 * i + k);
            w[k] = kg_pack4(v.x) | (kg_pack4(v.y) << 8) | (kg_pack4(v.z) << 16) | (kg_pack4(v.w) << 24);
        }
        const uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
        for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint4*>(a.recv[p] + slot + 16 * i) = o;
    }
    // what is left of n (a multiple of 16, not necessarily of 64): one word per thread
    for (int64_t i = 4 * n64 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n / 16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(in + i);
        const uint32_t o = kg_pack4(v.x) | (kg_pack4(v.y) << 8) | (kg_pack4(v.z) << 16) | (kg_pack4(v.w) << 24);
        for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint32_t*>(a.recv[p] + slot + 4 * i) = o;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n_local; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = a.n_dibits[i];
        for (int p = 0; p < a.world; ++p) *reinterpret_cast<int32_t*>(a.recv[p] + slot + a.n / 4 + 4 * i) = v;
    }
    // every thread's remote stores are ordered before its CTA's ticket; the last CTA publishes the step everywhere
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < a.world) {
        uint32_t* flag = reinterpret_cast<uint32_t*>(a.recv[threadIdx.x] + a.flag_off) + (a.step & 1) * KG_MAX_WORLD + a.rank;
        kg_st_release_sys(flag, a.step);
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

__global__ void __launch_bounds__(KG_THREADS) k_gather_wait_unpack(const GatherArgs a) {
    const int r = blockIdx.y;
    const uint8_t* mine = a.recv[a.rank];
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + a.flag_off) + (a.step & 1) * KG_MAX_WORLD + r;
        const long long t0 = clock64();
        int ok = 1;
        // steps are compared as a signed distance: the counter may wrap
        while ((int32_t)(kg_ld_acquire_sys(flag) - a.step) < 0) {
            if (clock64() - t0 > KG_SPIN_CLOCKS) { ok = 0; break; }
            __nanosleep(64);
        }
        if (!ok) atomicMax(a.status, 1 + r);
        s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
    const uint8_t* blk = mine + ((int64_t)(a.step & 1) * a.world + r) * a.block;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(blk);
    uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)r * a.n);
    const int64_t words = a.n / 16;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldcg(src + i);                    // written by a peer: read through L2
        dst[i] = make_uint4(kg_unpack4(w & 0xFFu), kg_unpack4((w >> 8) & 0xFFu), kg_unpack4((w >> 16) & 0xFFu), kg_unpack4(w >> 24));
    }
    if (a.out_n) {
        const int32_t* ln = reinterpret_cast<const int32_t*>(blk + a.n / 4);
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n_local; i += (int64_t)gridDim.x * blockDim.x)
            a.out_n[(int64_t)r * a.n_local + i] = __ldcg(ln + i);
    }
}

}  // namespace tetra
