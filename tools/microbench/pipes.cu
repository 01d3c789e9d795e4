// Microbenchmarks that size the fused kernel's inner loops on B200 (sm_100a):
//   FFMA2 / FFMA issue rate per SM sub-partition vs warps and independent accumulators,
//   the same with interleaved LDS.128, and DFMA rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__constant__ float c_tap[64];

template <int NACC, int MODE>   // MODE 0: FFMA2 reg taps, 1: FFMA2 const taps, 2: scalar FFMA const taps, 3: DFMA
__global__ void k_fma(float* out, int iters, long long* cyc) {
    float2 acc[NACC];
    double dacc[NACC];
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); dacc[i] = threadIdx.x + i; }
    float2 x = make_float2(1.0001f + threadIdx.x * 1e-6f, 0.9999f);
    float2 t = make_float2(0.999f, 1.001f);
    double dx = 1.0000001, dt = 0.9999999;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (MODE == 0) acc[i] = ffma2(x, t, acc[i]);
                else if (MODE == 1) { float c = c_tap[(r * NACC + i) & 63]; acc[i] = ffma2(x, make_float2(c, c), acc[i]); }
                else if (MODE == 2) { float c = c_tap[(r * NACC + i) & 63]; acc[i].x = fmaf(x.x, c, acc[i].x); acc[i].y = fmaf(x.y, c, acc[i].y); }
                else dacc[i] = fma(dx, dt, dacc[i]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0; double ds = 0;
    for (int i = 0; i < NACC; ++i) { s += acc[i].x + acc[i].y; ds += dacc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)ds;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// FFMA2 fed by LDS.128: each LDS.128 (2 complex samples) feeds 2*NACC FFMA2 (the FIR register tile shape)
template <int NACC>
__global__ void k_fir(float* out, int iters, long long* cyc) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i * 1e-4f, 1.f, 0.5f, 0.25f);
    float2 acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(0.f, 0.f);
    __syncthreads();
    const int base = (threadIdx.x * 5) & 2047;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 v = sm[(base + it + r) & 4095];
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                acc[i] = ffma2(make_float2(v.x, v.y), make_float2(c_tap[(2 * r + i) & 63], c_tap[(2 * r + i) & 63]), acc[i]);
                acc[i] = ffma2(make_float2(v.z, v.w), make_float2(c_tap[(2 * r + i + 1) & 63], c_tap[(2 * r + i + 1) & 63]), acc[i]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class F>
void run(const char* name, F launch, int threads, double inst_per_thread_iter, int iters) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    launch(out, 10, cyc); cudaDeviceSynchronize();
    launch(out, iters, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const double warps_per_smsp = threads / 32.0 / 4.0;
    const double inst = inst_per_thread_iter * iters * warps_per_smsp;      // warp-instr per SMSP
    printf("%-34s thr=%4d  cycles=%9lld  warp-instr/clk/SMSP=%.3f  %s\n", name, threads, c, inst / c, e ? cudaGetErrorString(e) : "");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f / (i + 1);
    cudaMemcpyToSymbol(c_tap, h, sizeof h);
    const int iters = 2000;
    for (int thr : {128, 256, 512, 1024}) {
#define RUN_FMA(NACC, MODE, label) run(label, [&](float* o, int it, long long* c) { k_fma<NACC, MODE><<<148, thr>>>(o, it, c); }, thr, 8.0 * NACC * (MODE == 2 ? 2 : 1), iters)
        RUN_FMA(4, 0, "FFMA2 reg taps  acc=4");
        RUN_FMA(10, 0, "FFMA2 reg taps  acc=10");
        RUN_FMA(4, 1, "FFMA2 const tap acc=4");
        RUN_FMA(10, 1, "FFMA2 const tap acc=10");
        RUN_FMA(10, 2, "FFMA  const tap acc=10x2");
        RUN_FMA(8, 3, "DFMA acc=8");
#define RUN_FIR(NACC, label) run(label, [&](float* o, int it, long long* c) { k_fir<NACC><<<148, thr, 65536>>>(o, it, c); }, thr, 16.0 * (2 * NACC + 1), iters / 4)
        cudaFuncSetAttribute(k_fir<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(k_fir<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        RUN_FIR(5, "LDS.128 + 10 FFMA2 (acc=5)");
        RUN_FIR(10, "LDS.128 + 20 FFMA2 (acc=10)");
    }
    return 0;
}
