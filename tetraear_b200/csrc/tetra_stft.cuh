// K3 waterfall STFT + the small complex128 helper kernels behind the public helper entry points.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tetra {

// ----------------------------------------------------------------------------------------------
// K3  stft_power_db : ui/modern.py:1921-1934 applied at every hop.
//   row r: X = FFT(hann(nfft) * x[r*hop : r*hop+nfft]);  out[r][(k + nfft/2) % nfft] = 20 log10(|X[k]|/nfft + 1e-20)
// One persistent CTA per SM builds the Hann window and the twiddle table once in shared memory
// (fp64 sincospi, rounded to fp32) and then transforms rows with an in-smem radix-2 Stockham FFT.
// ----------------------------------------------------------------------------------------------
constexpr int STFT_THREADS = 512;

template <int NFFT>
__global__ void __launch_bounds__(STFT_THREADS) k_stft_db(const float2* __restrict__ x, int64_t n, int hop, int64_t rows,
                                                            float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    float2* buf0 = reinterpret_cast<float2*>(sm_raw);
    float2* buf1 = buf0 + NFFT;
    float2* tw = buf1 + NFFT;                    // NFFT/2 twiddles exp(-2 pi i t / NFFT)
    float* win = reinterpret_cast<float*>(tw + NFFT / 2);
    const int tid = threadIdx.x;
    for (int t = tid; t < NFFT / 2; t += STFT_THREADS) {
        double s, c;
        sincospi(-2.0 * (double)t / (double)NFFT, &s, &c);
        tw[t] = make_float2((float)c, (float)s);
    }
    for (int t = tid; t < NFFT; t += STFT_THREADS)   // np.hanning: symmetric, 0.5 - 0.5 cos(2 pi n / (N-1))
        win[t] = (float)(0.5 - 0.5 * cospi(2.0 * (double)t / (double)(NFFT - 1)));
    __syncthreads();
    const float inv_n = 1.0f / (float)NFFT;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const float2* xr = x + r * hop;
        for (int t = tid; t < NFFT; t += STFT_THREADS) {
            const float2 v = __ldg(xr + t);
            const float w = win[t];
            buf0[t] = make_float2(v.x * w, v.y * w);
        }
        __syncthreads();
        float2* src = buf0;
        float2* dst = buf1;
#pragma unroll 1
        for (int ns = 1; ns < NFFT; ns <<= 1) {
            const int tw_stride = NFFT / (2 * ns);
            for (int j = tid; j < NFFT / 2; j += STFT_THREADS) {
                const int k = j & (ns - 1);
                const float2 w = tw[k * tw_stride];
                const float2 a = src[j];
                const float2 b = src[j + NFFT / 2];
                const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
                const int j0 = ((j - k) << 1) + k;
                dst[j0] = make_float2(a.x + bw.x, a.y + bw.y);
                dst[j0 + ns] = make_float2(a.x - bw.x, a.y - bw.y);
            }
            __syncthreads();
            float2* t = src; src = dst; dst = t;
        }
        float* orow = out + r * NFFT;
        for (int k = tid; k < NFFT; k += STFT_THREADS) {
            const float2 v = src[(k + NFFT / 2) & (NFFT - 1)];   // fftshift
            const float mag = __fsqrt_rn(v.x * v.x + v.y * v.y) * inv_n + 1e-20f;
            orow[k] = 6.020599913f * __log2f(mag);          // 20 log10(m) through the hardware log2 (~1e-6 dB)
        }
        __syncthreads();
    }
}

// The same in float64 on complex128 input: the reference's own precision (numpy FFT of complex128, ui/modern.py:1924-1934).
// This is what SignalProcessor.spectrum() -- the once-per-chunk spectrum block of the GUI, SURVEY 8 row a11 -- runs: a
// 2048-point row costs microseconds either way, and float64 keeps every bin above -300 dBFS within 1e-9 dB of the reference,
// where a float32 FFT's rounding floor sits ~140 dB below the strongest bin (tests/test_gpu_configs.py). NFFT <= 4096.
template <int NFFT>
__global__ void __launch_bounds__(STFT_THREADS) k_stft_db_f64(const double2* __restrict__ x, int hop, int64_t rows, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double2* buf0 = reinterpret_cast<double2*>(sm_raw);
    double2* buf1 = buf0 + NFFT;
    double2* tw = buf1 + NFFT;
    double* win = reinterpret_cast<double*>(tw + NFFT / 2);
    const int tid = threadIdx.x;
    for (int t = tid; t < NFFT / 2; t += STFT_THREADS) {
        double s, c;
        sincospi(-2.0 * (double)t / (double)NFFT, &s, &c);
        tw[t] = make_double2(c, s);
    }
    for (int t = tid; t < NFFT; t += STFT_THREADS) win[t] = 0.5 - 0.5 * cospi(2.0 * (double)t / (double)(NFFT - 1));
    __syncthreads();
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const double2* xr = x + r * hop;
        for (int t = tid; t < NFFT; t += STFT_THREADS) {
            const double2 v = xr[t];
            buf0[t] = make_double2(v.x * win[t], v.y * win[t]);
        }
        __syncthreads();
        double2* src = buf0;
        double2* dst = buf1;
#pragma unroll 1
        for (int ns = 1; ns < NFFT; ns <<= 1) {
            const int tw_stride = NFFT / (2 * ns);
            for (int j = tid; j < NFFT / 2; j += STFT_THREADS) {
                const int k = j & (ns - 1);
                const double2 w = tw[k * tw_stride];
                const double2 a = src[j];
                const double2 b = src[j + NFFT / 2];
                const double2 bw = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
                const int j0 = ((j - k) << 1) + k;
                dst[j0] = make_double2(a.x + bw.x, a.y + bw.y);
                dst[j0 + ns] = make_double2(a.x - bw.x, a.y - bw.y);
            }
            __syncthreads();
            double2* t = src; src = dst; dst = t;
        }
        double* orow = out + r * NFFT;
        for (int k = tid; k < NFFT; k += STFT_THREADS) {
            const double2 v = src[(k + NFFT / 2) & (NFFT - 1)];
            orow[k] = 20.0 * log10(hypot(v.x, v.y) / (double)NFFT + 1e-20);
        }
        __syncthreads();
    }
}

template <int NFFT>
static int stft_f64_launch_t(cudaStream_t st, const double2* x, int hop, int64_t rows, double* out) {
    const size_t smem = (size_t)NFFT * 16 * 2 + (size_t)NFFT / 2 * 16 + (size_t)NFFT * 8;
    if (cudaFuncSetAttribute(k_stft_db_f64<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    k_stft_db_f64<NFFT><<<(int)std::min<int64_t>(rows, 148), STFT_THREADS, smem, st>>>(x, hop, rows, out);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}
static int stft_f64_launch(cudaStream_t st, const double2* x, int nfft, int hop, int64_t rows, double* out) {
    switch (nfft) {
        case 64: return stft_f64_launch_t<64>(st, x, hop, rows, out);
        case 128: return stft_f64_launch_t<128>(st, x, hop, rows, out);
        case 256: return stft_f64_launch_t<256>(st, x, hop, rows, out);
        case 512: return stft_f64_launch_t<512>(st, x, hop, rows, out);
        case 1024: return stft_f64_launch_t<1024>(st, x, hop, rows, out);
        case 2048: return stft_f64_launch_t<2048>(st, x, hop, rows, out);
        case 4096: return stft_f64_launch_t<4096>(st, x, hop, rows, out);
    }
    return -1;
}

// ----------------------------------------------------------------------------------------------
// K3 for the waterfall size of BASELINE config 5 (4096 = 16^3): three radix-16 Stockham passes with the 16-point
// DFTs in registers (two radix-4 stages), so a row costs 3 shared-memory exchanges instead of 12. 256 threads, one
// (padded) row buffer, Hann window and the W_4096 table built once per persistent CTA; pass 1 reads the row from
// global memory, pass 3 writes the fftshifted dB row. Two CTAs per SM.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }      // a * (-j)
__device__ __forceinline__ float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }      // a * (+j)

// b[k] = sum_n a[n] exp(-2 pi j n k / 16), in place
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    const float2 w16[10] = {{1.f, 0.f}, {c1, -s1}, {h, -h}, {s1, -c1}, {0.f, -1.f}, {-s1, -c1}, {-h, -h}, {-c1, -s1}, {-1.f, 0.f}, {-c1, s1}};
    float2 t[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 s02 = cadd(a[i], a[i + 8]), d02 = csub(a[i], a[i + 8]);
        const float2 s13 = cadd(a[i + 4], a[i + 12]), d13 = csub(a[i + 4], a[i + 12]);
        t[i][0] = cadd(s02, s13);
        t[i][1] = cadd(d02, mul_mj(d13));
        t[i][2] = csub(s02, s13);
        t[i][3] = cadd(d02, mul_pj(d13));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = (i * q == 0) ? t[i][q] : cmulf(t[i][q], w16[i * q]);   // i*q <= 9
        const float2 s02 = cadd(y[0], y[2]), d02 = csub(y[0], y[2]);
        const float2 s13 = cadd(y[1], y[3]), d13 = csub(y[1], y[3]);
        a[q] = cadd(s02, s13);
        a[q + 4] = cadd(d02, mul_mj(d13));
        a[q + 8] = csub(s02, s13);
        a[q + 12] = cadd(d02, mul_pj(d13));
    }
}

constexpr int S4K_N = 4096, S4K_THREADS = 256;
__device__ __forceinline__ int s4k_pad(int i) { return i + (i >> 4); }
// twiddles laid out so that a warp reads consecutive shared-memory words: tw3[q][j] = W_4096^(q j) (pass 3, thread j),
// tw2[q][k] = W_256^(q k) (pass 2, k = j mod 16); every entry is float64 trigonometry rounded once to float32
struct S4kTables { float2 tw3[16][256]; float2 tw2[16][16]; float win[S4K_N]; };
struct S4kSmem {
    float2 buf[S4K_N + S4K_N / 16];
    S4kTables tab;
};

// the tables are built once per context
__global__ void k_stft4096_tables(S4kTables* tab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S4K_N) return;
    double sn, cs;
    const int q = t >> 8, j = t & 255;
    sincospi(-2.0 * (double)((q * j) & (S4K_N - 1)) / (double)S4K_N, &sn, &cs);
    tab->tw3[q][j] = make_float2((float)cs, (float)sn);
    if (t < 256) {
        const int q2 = t >> 4, k = t & 15;
        sincospi(-2.0 * (double)((q2 * k) & 255) / 256.0, &sn, &cs);
        tab->tw2[q2][k] = make_float2((float)cs, (float)sn);
    }
    tab->win[t] = (float)(0.5 - 0.5 * cospi(2.0 * (double)t / (double)(S4K_N - 1)));   // np.hanning (symmetric)
}

__global__ void __launch_bounds__(S4K_THREADS, 2) k_stft4096_db(const float2* __restrict__ x, int hop, int64_t rows, float* __restrict__ out,
                                                                const S4kTables* __restrict__ tab) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    S4kSmem& sm = *reinterpret_cast<S4kSmem*>(sm_raw);
    const int j = threadIdx.x;
    {
        const float4* src = reinterpret_cast<const float4*>(tab);
        float4* dst = reinterpret_cast<float4*>(&sm.tab);
        for (int t = j; t < (int)(sizeof(S4kTables) / 16); t += S4K_THREADS) dst[t] = __ldg(src + t);
    }
    __syncthreads();
    const float inv_n = 1.0f / (float)S4K_N;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const float2* xr = x + r * hop;
        float2 a[16];
        // pass 1 (sub-transform size 1): inputs j + 256 q straight from global memory, windowed
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float2 v = __ldg(xr + j + 256 * q);
            const float w = sm.tab.win[j + 256 * q];
            a[q] = make_float2(v.x * w, v.y * w);
        }
        dft16(a);
#pragma unroll
        for (int q = 0; q < 16; ++q) sm.buf[s4k_pad(16 * j + q)] = a[q];
        __syncthreads();
        // pass 2 (sub-transform size 16): twiddle W_256^(q k), k = j mod 16
        {
            const int k = j & 15;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float2 v = sm.buf[s4k_pad(j + 256 * q)];
                a[q] = q == 0 ? v : cmulf(v, sm.tab.tw2[q][k]);
            }
            __syncthreads();
            dft16(a);
#pragma unroll
            for (int q = 0; q < 16; ++q) sm.buf[s4k_pad((j - k) * 16 + k + 16 * q)] = a[q];
            __syncthreads();
        }
        // pass 3 (sub-transform size 256): twiddle W_4096^(q j); outputs X[j + 256 q]
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float2 v = sm.buf[s4k_pad(j + 256 * q)];
            a[q] = q == 0 ? v : cmulf(v, sm.tab.tw3[q][j]);
        }
        __syncthreads();                                 // the buffer is free for the next row's pass 1
        dft16(a);
        float* orow = out + r * S4K_N;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int kk = j + 256 * q;                  // bin; fftshift puts it at (kk + N/2) mod N
            // 20 log10(m) = 6.0206 log2(m) through the hardware log2 (absolute error ~2e-7 in log2, i.e. ~1e-6 dB)
            const float mag = __fsqrt_rn(a[q].x * a[q].x + a[q].y * a[q].y) * inv_n + 1e-20f;
            orow[(kk + S4K_N / 2) & (S4K_N - 1)] = 6.020599913f * __log2f(mag);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// K3, two rows per thread: the same three radix-16 passes with every value held as a PACKED PAIR (row r, row r + 1) --
// Blackwell's fp32x2 instructions (SASS FADD2 / FMUL2 / FFMA2; twiddles and constants enter as broadcast scalar operands,
// which the packed forms take for free) do each butterfly step of both rows in one issue slot. Real and imaginary parts
// live in separate register pairs, so the +-j rotations of the radix-4 butterflies stay free. The kernel above is
// issue-bound (1099 instructions per thread and row, 59 % of the issue slots, 43 % of the FMA pipes); this one needs about
// half the issue slots per row. Shared memory: one row-pair buffer of float4 {re_A, re_B, im_A, im_B} (padded), the
// twiddle tables, half of the symmetric window: 110 KB, two CTAs per SM.
// ----------------------------------------------------------------------------------------------
typedef unsigned long long pk2;                                      // two floats: (row A, row B)
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) { pk2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b) { pk2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) { pk2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) { pk2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2 pk_dup(float x) { pk2 d; asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(x)); return d; }   // folds into a scalar operand
__device__ __forceinline__ pk2 pk_make(float a, float b) { pk2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ void pk_split(pk2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
struct pkc { pk2 re, im; };                                          // a complex value of both rows
__device__ __forceinline__ pkc pkc_add(pkc a, pkc b) { return pkc{pk_add(a.re, b.re), pk_add(a.im, b.im)}; }
__device__ __forceinline__ pkc pkc_sub(pkc a, pkc b) { return pkc{pk_sub(a.re, b.re), pk_sub(a.im, b.im)}; }
__device__ __forceinline__ pkc pkc_add_mj(pkc a, pkc b) { return pkc{pk_add(a.re, b.im), pk_sub(a.im, b.re)}; }   // a + (-j) b
__device__ __forceinline__ pkc pkc_add_pj(pkc a, pkc b) { return pkc{pk_sub(a.re, b.im), pk_add(a.im, b.re)}; }   // a + (+j) b
__device__ __forceinline__ pkc pkc_mulc(pkc a, float c, float s) {                                                 // a * (c + j s)
    const pk2 cc = pk_dup(c), ss = pk_dup(s);
    return pkc{pk_sub(pk_mul(a.re, cc), pk_mul(a.im, ss)), pk_fma(a.re, ss, pk_mul(a.im, cc))};
}
// b[k] = sum_n a[n] exp(-2 pi j n k / 16), in place (the butterfly network of dft16 above)
__device__ __forceinline__ void dft16_pk(pkc (&a)[16]) {
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    pkc t[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const pkc s02 = pkc_add(a[i], a[i + 8]), d02 = pkc_sub(a[i], a[i + 8]);
        const pkc s13 = pkc_add(a[i + 4], a[i + 12]), d13 = pkc_sub(a[i + 4], a[i + 12]);
        t[i][0] = pkc_add(s02, s13);
        t[i][1] = pkc_add_mj(d02, d13);
        t[i][2] = pkc_sub(s02, s13);
        t[i][3] = pkc_add_pj(d02, d13);
    }
    // y[i] = t[i][q] W16^(i q): q = 0 none; the others by value (i q = 1, 2, 3 | 2, 4, 6 | 3, 6, 9)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pkc y0 = t[0][q], y1, y2, y3;
        if (q == 0) { y1 = t[1][0]; y2 = t[2][0]; y3 = t[3][0]; }
        else if (q == 1) { y1 = pkc_mulc(t[1][1], c1, -s1); y2 = pkc_mulc(t[2][1], h, -h); y3 = pkc_mulc(t[3][1], s1, -c1); }
        else if (q == 2) { y1 = pkc_mulc(t[1][2], h, -h); y2 = t[2][2]; y3 = pkc_mulc(t[3][2], -h, -h); }     // y2 still needs its -j
        else { y1 = pkc_mulc(t[1][3], s1, -c1); y2 = pkc_mulc(t[2][3], -h, -h); y3 = pkc_mulc(t[3][3], -c1, s1); }
        pkc s02, d02;
        if (q == 2) { s02 = pkc_add_mj(y0, y2); d02 = pkc_add_pj(y0, y2); }      // y0 +- (-j) t
        else { s02 = pkc_add(y0, y2); d02 = pkc_sub(y0, y2); }
        const pkc s13 = pkc_add(y1, y3), d13 = pkc_sub(y1, y3);
        a[q] = pkc_add(s02, s13);
        a[q + 4] = pkc_add_mj(d02, d13);
        a[q + 8] = pkc_sub(s02, s13);
        a[q + 12] = pkc_add_pj(d02, d13);
    }
}

struct S4kPairSmem {
    float4 buf[S4K_N + S4K_N / 16];      // {re_A, re_B, im_A, im_B}
    float2 tw3[16][256];
    float2 tw2[16][16];
    float win[S4K_N / 2];                // symmetric: win[t] = win[N - 1 - t]
};
static_assert(2 * (sizeof(S4kPairSmem) + 1024) <= 227 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(S4K_THREADS, 2) k_stft4096_db2(const float2* __restrict__ x, int hop, int64_t rows, float* __restrict__ out,
                                                                 const S4kTables* __restrict__ tab) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    S4kPairSmem& sm = *reinterpret_cast<S4kPairSmem*>(sm_raw);
    const int j = threadIdx.x;
    {
        const float4* src = reinterpret_cast<const float4*>(tab);          // tw3 and tw2 lie first in S4kTables
        float4* dst = reinterpret_cast<float4*>(&sm.tw3[0][0]);
        for (int t = j; t < (int)((sizeof(sm.tw3) + sizeof(sm.tw2)) / 16); t += S4K_THREADS) dst[t] = __ldg(src + t);
        for (int t = j; t < S4K_N / 2; t += S4K_THREADS) sm.win[t] = __ldg(&tab->win[t]);
    }
    __syncthreads();
    const int k = j & 15;
    const int64_t pairs = (rows + 1) / 2;
    for (int64_t p = blockIdx.x; p < pairs; p += gridDim.x) {
        const int64_t ra = 2 * p;
        const bool has_b = ra + 1 < rows;
        const float2* xa = x + ra * hop;
        const float2* xb = has_b ? xa + hop : xa;
        pkc a[16];
        // pass 1 (sub-transform size 1): inputs j + 256 q of both rows straight from global memory, windowed
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int t = j + 256 * q;
            const float2 va = __ldg(xa + t), vb = __ldg(xb + t);
            const float w = sm.win[q < 8 ? t : S4K_N - 1 - t];
            a[q].re = pk_make(va.x * w, vb.x * w);
            a[q].im = pk_make(va.y * w, vb.y * w);
        }
        dft16_pk(a);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float4 v;
            pk_split(a[q].re, v.x, v.y); pk_split(a[q].im, v.z, v.w);
            sm.buf[s4k_pad(16 * j + q)] = v;
        }
        __syncthreads();
        // pass 2 (sub-transform size 16): twiddle W_256^(q k), k = j mod 16
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 v = sm.buf[s4k_pad(j + 256 * q)];
            const pkc c{pk_make(v.x, v.y), pk_make(v.z, v.w)};
            if (q == 0) a[q] = c;
            else { const float2 w = sm.tw2[q][k]; a[q] = pkc_mulc(c, w.x, w.y); }
        }
        __syncthreads();
        dft16_pk(a);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float4 v;
            pk_split(a[q].re, v.x, v.y); pk_split(a[q].im, v.z, v.w);
            sm.buf[s4k_pad((j - k) * 16 + k + 16 * q)] = v;
        }
        __syncthreads();
        // pass 3 (sub-transform size 256): twiddle W_4096^(q j); outputs X[j + 256 q]
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 v = sm.buf[s4k_pad(j + 256 * q)];
            const pkc c{pk_make(v.x, v.y), pk_make(v.z, v.w)};
            if (q == 0) a[q] = c;
            else { const float2 w = sm.tw3[q][j]; a[q] = pkc_mulc(c, w.x, w.y); }
        }
        __syncthreads();                                 // the buffer is free for the next pair's pass 1
        dft16_pk(a);
        float* oa = out + ra * S4K_N;
        float* ob = has_b ? oa + S4K_N : oa;             // an odd last row is simply written twice
        // 20 log10(sqrt(p) / N + 1e-20) = 10 log10(p) - 20 log10(N) while the 1e-20 stays below 4e-9 of the magnitude (3.6e-8 dB);
        // bins below that (an all-zero row, 160 dB under full scale) are redone with the literal expression afterwards
        bool small = false;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int col = (j + 256 * q + S4K_N / 2) & (S4K_N - 1);       // fftshift
            float pa, pb;
            pk_split(pk_fma(a[q].re, a[q].re, pk_mul(a[q].im, a[q].im)), pa, pb);
            small |= fminf(pa, pb) < 1e-16f;
            float la, lb;                                  // bare MUFU.LG2: values below 1e-16 (denormals included) are redone below
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(pa));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(pb));
            oa[col] = fmaf(3.010299956639812f, la, -72.24719895935549f);
            ob[col] = fmaf(3.010299956639812f, lb, -72.24719895935549f);
        }
        if (small) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int col = (j + 256 * q + S4K_N / 2) & (S4K_N - 1);
                float pa, pb;
                pk_split(pk_fma(a[q].re, a[q].re, pk_mul(a[q].im, a[q].im)), pa, pb);
                if (pa < 1e-16f) oa[col] = 6.020599913f * __log2f(sqrtf(pa) * (1.0f / (float)S4K_N) + 1e-20f);
                if (pb < 1e-16f) ob[col] = 6.020599913f * __log2f(sqrtf(pb) * (1.0f / (float)S4K_N) + 1e-20f);
            }
        }
    }
}

static int stft4096_pair_launch(cudaStream_t st, const float2* x, int hop, int64_t rows, float* out, const S4kTables* tab) {
    if (cudaFuncSetAttribute(k_stft4096_db2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S4kPairSmem)) != cudaSuccess) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>((rows + 1) / 2, (int64_t)sms * 2);
    k_stft4096_db2<<<grid, S4K_THREADS, sizeof(S4kPairSmem), st>>>(x, hop, rows, out, tab);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

static int stft4096_launch(cudaStream_t st, const float2* x, int hop, int64_t rows, float* out, const S4kTables* tab) {
    if (cudaFuncSetAttribute(k_stft4096_db, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S4kSmem)) != cudaSuccess) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(rows, (int64_t)sms * 2);
    k_stft4096_db<<<grid, S4K_THREADS, sizeof(S4kSmem), st>>>(x, hop, rows, out, tab);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

template <int NFFT>
static int stft_launch_t(cudaStream_t st, const float2* x, int64_t n, int hop, int64_t rows, float* out) {
    const size_t smem = (size_t)NFFT * 8 * 2 + (size_t)NFFT / 2 * 8 + (size_t)NFFT * 4;
    if (cudaFuncSetAttribute(k_stft_db<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;
    const int grid = (int)std::min<int64_t>(rows, (int64_t)sms * per_sm);
    k_stft_db<NFFT><<<grid, STFT_THREADS, smem, st>>>(x, n, hop, rows, out);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

static int stft_launch(cudaStream_t st, const float2* x, int64_t n, int nfft, int hop, int64_t rows, float* out, const S4kTables* tab4k) {
    switch (nfft) {
        case 64: return stft_launch_t<64>(st, x, n, hop, rows, out);
        case 128: return stft_launch_t<128>(st, x, n, hop, rows, out);
        case 256: return stft_launch_t<256>(st, x, n, hop, rows, out);
        case 512: return stft_launch_t<512>(st, x, n, hop, rows, out);
        case 1024: return stft_launch_t<1024>(st, x, n, hop, rows, out);
        case 2048: return stft_launch_t<2048>(st, x, n, hop, rows, out);
        case 4096: {
            static const int pair = getenv("TETRA_STFT_PAIR") ? atoi(getenv("TETRA_STFT_PAIR")) : 1;   // 0: the one-row kernel (A/B)
            return pair ? stft4096_pair_launch(st, x, hop, rows, out, tab4k) : stft4096_launch(st, x, hop, rows, out, tab4k);
        }
        case 8192: return stft_launch_t<8192>(st, x, n, hop, rows, out);
    }
    return -1;
}

// ----------------------------------------------------------------------------------------------
// complex128 helpers
// ----------------------------------------------------------------------------------------------
// frequency_shift (processor.py:97-100): x[n] * exp(-1j * 2 pi f * (n / fs))
__global__ void k_nco_c128(double2* x, int64_t n, double fo, double fs) {
    const double w = (2.0 * M_PI) * fo;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = (double)i / fs;
        double s, c;
        sincos(-(w * t), &s, &c);
        const double2 v = x[i];
        x[i] = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

// ----------------------------------------------------------------------------------------------
// k_analyze: the per-sample part of TetraSignalDetector.analyze_signal (signal/scanner.py:42-147, 204-231) for C captures
// at once: power, the pi/4 phase-step clustering score, the 31-bit sync pattern correlation on crude bits, and the
// power-stability verdict. float64 throughout (the reference works on complex128). One CTA per capture.
//   out[c] = (power_db, modulation_confidence, sync_correlation, power_stable, modulation_matches, n_phase_diffs)
// ----------------------------------------------------------------------------------------------
constexpr int ANA_THREADS = 512;
constexpr int ANA_MAXBITS = 1 << 20;                   // crude bits kept bit-packed in shared memory (128 KB)
constexpr uint32_t ANA_PATTERN = 0x2CE259C4u;          // 0101100111000100101100111000100 (scanner.py:133-134), first bit = MSB of 31

__device__ __forceinline__ double ana_wrap(double d) {  // (d + pi) % (2 pi) - pi with Python's modulo
    const double t = d + M_PI;
    return (t - (2.0 * M_PI) * floor(t / (2.0 * M_PI))) - M_PI;
}
__device__ __forceinline__ double ana_block_sum(double v, double* red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < ANA_THREADS / 32; ++w) t += red[w];
    return t;
}

// dphi (or null): capture c is analysed as frequency_shift(x, f_c) (processor.py:85-100) -- the scanner retuned to channel c
// (signal/scanner.py:383-445) -- with dphi[c] = 2 pi f_c / fs: the shift leaves every |x| alone and takes dphi off every
// phase step (ds dphi off the steps between every ds-th sample), so the shifted capture is never formed. out_stride >= 6.
__global__ void __launch_bounds__(ANA_THREADS) k_analyze(const float2* __restrict__ x, int64_t pitch, int64_t n, int ds, double* __restrict__ out,
                                                         const double* __restrict__ dphi, int out_stride) {
    extern __shared__ uint32_t s_bits[];                // ceil(nbits / 32) + 2 words
    __shared__ double red[ANA_THREADS / 32];
    __shared__ unsigned long long s_max;
    __shared__ int s_best;
    const int tid = threadIdx.x;
    const float2* xc = x + (int64_t)blockIdx.x * pitch;
    double* o = out + (int64_t)out_stride * blockIdx.x;
    const double dph = dphi ? dphi[blockIdx.x] : 0.0;
    if (tid == 0) { s_max = 0ull; s_best = 0; }
    __syncthreads();
    // ---- power (scanner.py:42-55), per-window powers (:204-231) and max |x| (:72) ----
    const int64_t wlen = n / 5;
    double tot = 0.0, win[5] = {0, 0, 0, 0, 0}, mx = 0.0;
    for (int64_t i = tid; i < n; i += ANA_THREADS) {
        const float2 v = __ldg(xc + i);
        const double p = (double)v.x * v.x + (double)v.y * v.y;
        tot += p;
        mx = fmax(mx, hypot((double)v.x, (double)v.y));
        if (wlen > 0) {
            const int64_t w = i / wlen;
#pragma unroll
            for (int k = 0; k < 5; ++k) if (w == k) win[k] += p;
        }
    }
    tot = ana_block_sum(tot, red);
    double wsum[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) wsum[k] = ana_block_sum(win[k], red);
    atomicMax(&s_max, (unsigned long long)__double_as_longlong(mx));
    __syncthreads();
    const double scale = __longlong_as_double((long long)s_max) + 1e-10;
    // ---- detect_tetra_modulation (:57-96) ----
    double matches = 0.0;
    if (n >= 1000) {
        for (int64_t i = 1 + tid; i < n; i += ANA_THREADS) {
            const float2 a = __ldg(xc + i), b = __ldg(xc + i - 1);
            const double d = ana_wrap(atan2((double)a.y / scale, (double)a.x / scale) - atan2((double)b.y / scale, (double)b.x / scale) - dph);
            double best = 1e9;
#pragma unroll
            for (int k = -4; k < 4; ++k) best = fmin(best, fabs((double)k * (M_PI / 4.0) - d));
            matches += best < M_PI / 8.0 ? 1.0 : 0.0;
        }
    }
    matches = ana_block_sum(matches, red);
    // ---- detect_sync_pattern (:98-147): every ds-th sample, bit = 1 where the phase step rounds to 0 ----
    const int64_t n_sym = (n + ds - 1) / ds;
    const int64_t n_bits = n_sym >= 100 ? n_sym - 1 : 0;
    const int n_words = (int)((n_bits + 31) / 32) + 2;
    for (int w = tid; w < n_words; w += ANA_THREADS) s_bits[w] = 0u;
    __syncthreads();
    for (int64_t j = tid; j < n_bits; j += ANA_THREADS) {
        const float2 a = __ldg(xc + (j + 1) * ds), b = __ldg(xc + j * ds);
        const double d = ana_wrap(atan2((double)a.y, (double)a.x) - atan2((double)b.y, (double)b.x) - dph * (double)ds);
        const double q = rint(d / (M_PI / 4.0)) * (M_PI / 4.0);     // numpy round: half to even
        if (fabs(q) < M_PI / 8.0) atomicOr(&s_bits[j >> 5], 0x80000000u >> (j & 31));
    }
    __syncthreads();
    int best = 0;
    for (int64_t i = tid; i < n_bits - 31; i += ANA_THREADS) {        // range(len(bits) - 31): the last window is not tried
        const uint64_t two = ((uint64_t)s_bits[i >> 5] << 32) | s_bits[(i >> 5) + 1];
        const uint32_t w31 = (uint32_t)(two >> (64 - 31 - (i & 31))) & 0x7FFFFFFFu;
        best = max(best, 31 - __popc(w31 ^ ANA_PATTERN));
    }
    for (int of = 16; of; of >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, of));
    if ((tid & 31) == 0) atomicMax(&s_best, best);
    __syncthreads();
    if (tid == 0) {
        o[0] = n > 0 ? 10.0 * log10(tot / (double)n + 1e-10) : -120.0;
        o[1] = n >= 1000 ? matches / (double)(n - 1) : 0.0;
        o[2] = (double)s_best / 31.0;
        double stable = 0.0;
        if (n >= 5 * 1000) {
            double p[5], mean = 0.0, var = 0.0;
            for (int k = 0; k < 5; ++k) { p[k] = 10.0 * log10(wsum[k] / (double)wlen + 1e-10); mean += p[k]; }
            mean /= 5.0;
            for (int k = 0; k < 5; ++k) var += (p[k] - mean) * (p[k] - mean);
            stable = sqrt(var / 5.0) < 10.0 ? 1.0 : 0.0;
        }
        o[3] = stable;
        o[4] = matches;
        o[5] = n >= 1000 ? (double)(n - 1) : 0.0;
    }
}

// RTL-SDR native samples (interleaved unsigned 8-bit I, Q) -> complex64, the conversion pyrtlsdr's
// packed_bytes_to_iq applies before the reference ever sees the data (signal/capture.py:143-158 ->
// RtlSdr.read_samples):  iq = (byte / 127.5) - 1   per component. 8 samples (16 bytes in, 64 bytes out) per thread step.
__global__ void __launch_bounds__(256) k_u8_to_c64(const uint8_t* __restrict__ in, int64_t n_samples, float2* __restrict__ out) {
    const int64_t n16 = n_samples / 8;
    const uint4* in16 = reinterpret_cast<const uint4*>(in);
    // 16-byte loads AND 16-byte stores: rows of an odd length leave every other output row only 8-byte aligned
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    auto cv = [](uint32_t b) { return (float)((double)b / 127.5 - 1.0); };
    if (aligned) {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
            const uint4 v = __ldg(in16 + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            float4* o = reinterpret_cast<float4*>(out + 8 * i);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = make_float4(cv(w[k] & 255u), cv((w[k] >> 8) & 255u), cv((w[k] >> 16) & 255u), cv(w[k] >> 24));
        }
    }
    const int64_t done = aligned ? n16 * 8 : 0;
    for (int64_t i = done + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_samples; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = make_float2(cv(in[2 * i]), cv(in[2 * i + 1]));
}

// uint8 ingest on the fused path: only the two block-end windows the exact edge kernels read are expanded to complex64
// (the fused kernel converts its own tiles in shared memory). out[c] = [ x[0 .. wl) | x[n - wr .. n) ], same arithmetic.
__global__ void __launch_bounds__(256) k_u8_edge_windows(const uint8_t* __restrict__ in, int64_t pitch, int64_t n, int32_t wl, int32_t wr,
                                                         float2* __restrict__ out) {
    const uint8_t* row = in + 2 * (int64_t)blockIdx.y * pitch;
    float2* o = out + (int64_t)blockIdx.y * (wl + wr);
    auto cv = [](uint32_t b) { return (float)((double)b / 127.5 - 1.0); };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wl + wr; i += gridDim.x * blockDim.x) {
        const int64_t src = i < wl ? i : n - wr + (i - wl);
        o[i] = make_float2(cv(row[2 * src]), cv(row[2 * src + 1]));
    }
}

// Wideband channel selection (BASELINE config 3; the scanner's retune sweep, signal/scanner.py:383-445, done in
// software): out[c][n] = x[n] * exp(-1j * 2 pi f_c * (n / fs)), i.e. frequency_shift (processor.py:97-100) of one
// capture to C channel centres, with the reference's float64 phase and a complex64 result.
constexpr int MIX_RUN = 16;               // samples per thread sharing one float64 sine/cosine (stride = block size)
__global__ void __launch_bounds__(256) k_mix_wide(const float2* __restrict__ x, int64_t n, const double* __restrict__ freqs,
                                                    double fs, float2* __restrict__ out) {
    const int c = blockIdx.y;
    const double w = (2.0 * M_PI) * freqs[c];
    float2* oc = out + (int64_t)c * n;
    // Thread t of a block owns samples base + t + 256 u, u < MIX_RUN (coalesced). Its first phasor comes from the
    // reference's own expression -(w * (i / fs)); the others by rotating with exp(-j w 256 / fs) in float64, so the
    // recurrence error stays at a few ulp before the product is rounded to float32.
    double s1, c1;
    sincos(-(w * (256.0 / fs)), &s1, &c1);
    const int64_t span = 256 * MIX_RUN;
    for (int64_t base = blockIdx.x * span; base < n; base += (int64_t)gridDim.x * span) {
        const int64_t i0 = base + threadIdx.x;
        double sn, cs;
        sincos(-(w * ((double)i0 / fs)), &sn, &cs);
#pragma unroll
        for (int u = 0; u < MIX_RUN; ++u) {
            const int64_t i = i0 + 256 * u;
            if (i < n) {
                const float2 v = __ldg(x + i);
                oc[i] = make_float2((float)((double)v.x * cs - (double)v.y * sn), (float)((double)v.x * sn + (double)v.y * cs));
            }
            const double cn = cs * c1 - sn * s1;
            sn = cs * s1 + sn * c1;
            cs = cn;
        }
    }
}

// the same for two index ranges [lo0, lo0 + len0) and [lo1, lo1 + len1) only: the block-end windows the exact edge
// kernels read when the fused kernel takes the capture as is (MODE 2)
__global__ void __launch_bounds__(256) k_mix_wide_ranges(const float2* __restrict__ x, int64_t n, const double* __restrict__ freqs,
                                                           double fs, float2* __restrict__ out, int64_t lo0, int64_t len0,
                                                           int64_t lo1, int64_t len1) {
    const int c = blockIdx.y;
    const double w = (2.0 * M_PI) * freqs[c];
    float2* oc = out + (int64_t)c * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < len0 + len1; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t < len0 ? lo0 + t : lo1 + (t - len0);
        if (i < 0 || i >= n) continue;
        double sn, cs;
        sincos(-(w * ((double)i / fs)), &sn, &cs);
        const float2 v = __ldg(x + i);
        oc[i] = make_float2((float)((double)v.x * cs - (double)v.y * sn), (float)((double)v.x * sn + (double)v.y * cs));
    }
}

// extract_symbols (processor.py:179-219) on complex128: res[0] = n_symbols, res[1] = best phase
__global__ void __launch_bounds__(256) k_extract_c128(const double2* x, int64_t n, int sps, int step, double2* out, int64_t* res) {
    __shared__ double red[8];
    __shared__ int s_best;
    const int tid = threadIdx.x;
    int best = 0;
    double best_pow = -1.0;
    for (int ph = 0; ph < sps; ph += step) {
        const int64_t cnt = (n - ph) / sps;
        if (cnt <= 0) continue;
        double acc = 0.0;
        for (int64_t k = tid; k < cnt; k += 256) {
            const double2 v = x[ph + k * sps];
            acc += v.x * v.x + v.y * v.y;
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            const double mean = s / (double)cnt;
            if (mean > best_pow) { best_pow = mean; best = ph; }
        }
        __syncthreads();
    }
    if (tid == 0) s_best = best;
    __syncthreads();
    best = s_best;
    const int64_t cnt = (n - best) / sps > 0 ? (n - best) / sps : 0;
    for (int64_t k = tid; k < cnt; k += 256) out[k] = x[best + k * sps];
    if (tid == 0) { res[0] = cnt; res[1] = best; }
}

// max |x| (processor.py:124-125); doubles >= 0 order like their bit patterns
__global__ void k_maxabs_c128(const double2* x, int64_t n, unsigned long long* out) {
    double m = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, hypot(x[i].x, x[i].y));
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// demodulate_dqpsk (processor.py:127-163)
__global__ void k_slice_c128(const double2* x, int64_t n, const double* maxabs, uint8_t* out) {
    const double mx = *maxabs;
    for (int64_t i = 1 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 s1 = x[i], s0 = x[i - 1];
        if (mx > 0) { s1.x /= mx; s1.y /= mx; s0.x /= mx; s0.y /= mx; }
        const double re = s1.x * s0.x + s1.y * s0.y;
        const double im = s1.y * s0.x - s1.x * s0.y;
        const uint8_t d = slice_dqpsk(re, im);
        out[i - 1] = d;
    }
}

// ----------------------------------------------------------------------------------------------
// Wideband survey: per channel, the signal-presence / AFC block of CaptureThread.run
// (tetraear/ui/modern.py:1945-2012) on the spectrum of the first n_fft samples of the shifted capture.
// ----------------------------------------------------------------------------------------------
// out[c][i] = x[i] exp(-1j 2 pi f_c (i / fs)), i < len, complex128 (processor.py:97-100)
__global__ void __launch_bounds__(256) k_mix_head_f64(const float2* __restrict__ x, int len, const double* __restrict__ freqs, double fs,
                                                        double2* __restrict__ out) {
    const int c = blockIdx.y;
    const double w = (2.0 * M_PI) * freqs[c];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        double sn, cs;
        sincos(-(w * ((double)i / fs)), &sn, &cs);
        const float2 v = __ldg(x + i);
        out[(int64_t)c * len + i] = make_double2((double)v.x * cs - (double)v.y * sn, (double)v.x * sn + (double)v.y * cs);
    }
}
// one CTA per channel over its dB row `power` [nfft] (fftshifted): mean and peak of the centre +-bandwidth_bins/2 bins, the
// peak's frequency, the mean of everything more than 10 bins outside, and the verdict snr > 15 and peak > -70 and peak - mean > 3.
//   o[0..5] = signal_power, peak_power, peak_freq_offset_hz, noise_floor, snr, is_signal_strong
__global__ void __launch_bounds__(256) k_presence(const double* __restrict__ power, int nfft, double fs, double* __restrict__ out, int out_stride) {
    __shared__ double r_sum[8], r_noise[8], r_max[8];
    __shared__ int r_idx[8], r_cnt[8];
    const double* p = power + (int64_t)blockIdx.x * nfft;
    double* o = out + (int64_t)blockIdx.x * out_stride;
    const int tid = threadIdx.x;
    const int centre = nfft / 2;
    const int bw_bins = (int)(25000.0 / (fs / (double)nfft));
    const int start = max(0, centre - bw_bins / 2), end = min(nfft, centre + bw_bins / 2);
    const int n_end = max(0, start - 10), n_start2 = min(nfft, end + 10);
    double sum = 0.0, noise = 0.0, mx = -1e300;
    int idx = nfft, cnt = 0;
    for (int k = tid; k < nfft; k += 256) {
        const double v = p[k];
        if (k >= start && k < end) {
            sum += v;
            if (v > mx) { mx = v; idx = k; }                   // ascending k per thread: keeps its first maximum
        }
        if (k < n_end || k >= n_start2) { noise += v; ++cnt; }
    }
    for (int of = 16; of; of >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, of);
        noise += __shfl_xor_sync(0xffffffffu, noise, of);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, of);
        const double om = __shfl_xor_sync(0xffffffffu, mx, of);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, of);
        if (om > mx || (om == mx && oi < idx)) { mx = om; idx = oi; }      // np.argmax: the first maximum
    }
    if ((tid & 31) == 0) { r_sum[tid >> 5] = sum; r_noise[tid >> 5] = noise; r_max[tid >> 5] = mx; r_idx[tid >> 5] = idx; r_cnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            sum += r_sum[w]; noise += r_noise[w]; cnt += r_cnt[w];
            if (r_max[w] > mx || (r_max[w] == mx && r_idx[w] < idx)) { mx = r_max[w]; idx = r_idx[w]; }
        }
        if (end > start) {
            const double sig = sum / (double)(end - start);
            const double floor_db = cnt > 0 ? noise / (double)cnt : -100.0;
            const double snr = sig - floor_db;
            o[0] = sig; o[1] = mx;
            o[2] = (double)(idx - centre) * (fs / (double)nfft);      // fftshift(fftfreq(n, 1/fs))[idx]
            o[3] = floor_db; o[4] = snr;
            o[5] = (snr > 15.0 && mx > -70.0 && (mx - sig) > 3.0) ? 1.0 : 0.0;
        } else {
            for (int k = 0; k < 6; ++k) o[k] = 0.0;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Fourier resampling (scipy.signal.resample, two-sided FFT branch for complex input):
//   Y[:m2] = X[:m2]; Y[m2-m:] = X[m2-m:]; even m: down -> Y[m/2] += X[n-m/2]; up -> split the bin.
//   y = IDFT_num(Y) * (num / n)
// evaluated as two direct DFT sums over the <= m+1 kept bins with exact integer phase reduction.
// ----------------------------------------------------------------------------------------------
__global__ void k_phase_table(double2* tab, int64_t n, double sign) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s, c;
        sincospi(sign * 2.0 * (double)i / (double)n, &s, &c);
        tab[i] = make_double2(c, s);
    }
}
// one block per kept bin b (0 <= b < nb): bin index kx(b) in X; Ykeep[b] = sum_j x[j] tab[(j*kx) % n]
__global__ void __launch_bounds__(256) k_dft_bins(const double2* x, int64_t n, const double2* tab, int64_t m2, int64_t m,
                                                  int64_t nb, double2* ykeep) {
    __shared__ double2 red[8];
    for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        int64_t kx;
        if (b < m2) kx = b; else if (b < m) kx = n - (m - b); else kx = n - m / 2;   // b == m: the extra -m/2 bin
        double ar = 0, ai = 0;
        for (int64_t j = threadIdx.x; j < n; j += 256) {
            const double2 w = tab[(j * kx) % n];
            const double2 v = x[j];
            ar += v.x * w.x - v.y * w.y;
            ai += v.x * w.y + v.y * w.x;
        }
        for (int o = 16; o; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_double2(ar, ai);
        __syncthreads();
        if (threadIdx.x == 0) {
            double sr = 0, si = 0;
            for (int w = 0; w < 8; ++w) { sr += red[w].x; si += red[w].y; }
            ykeep[b] = make_double2(sr, si);
        }
        __syncthreads();
    }
}
// y[t] = (1/n) sum_b Yb * exp(+2 pi i ky(b) t / num)
__global__ void __launch_bounds__(256) k_idft_bins(const double2* ykeep, int64_t n, int64_t num, const double2* tab_out,
                                                   int64_t m2, int64_t m, int mode /*0 same,1 down,2 up*/, double2* y) {
    __shared__ double2 red[8];
    const bool even = (m % 2) == 0;
    for (int64_t t = blockIdx.x; t < num; t += gridDim.x) {
        double ar = 0, ai = 0;
        for (int64_t b = threadIdx.x; b < m; b += 256) {
            int64_t ky = b < m2 ? b : num - (m - b);
            double2 v = ykeep[b];
            if (even && b == m / 2) {
                if (mode == 1) { v.x += ykeep[m].x; v.y += ykeep[m].y; }
                else if (mode == 2) { v.x *= 0.5; v.y *= 0.5; }
            }
            const double2 w = tab_out[(ky * t) % num];
            ar += v.x * w.x - v.y * w.y;
            ai += v.x * w.y + v.y * w.x;
            if (even && b == m / 2 && mode == 2) {           // mirrored half of the split bin
                const double2 w2 = tab_out[((num - m / 2) * t) % num];
                ar += v.x * w2.x - v.y * w2.y;
                ai += v.x * w2.y + v.y * w2.x;
            }
        }
        for (int o = 16; o; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_double2(ar, ai);
        __syncthreads();
        if (threadIdx.x == 0) {
            double sr = 0, si = 0;
            for (int w = 0; w < 8; ++w) { sr += red[w].x; si += red[w].y; }
            y[t] = make_double2(sr / (double)n, si / (double)n);
        }
        __syncthreads();
    }
}

// scratch: tabs holds max(n, num) phase entries (reused for both directions), y receives num outputs
static int resample_c128(int64_t& launches, cudaStream_t st, const double2* x, int64_t n, double2* tabs, int64_t num, double2* y) {
    const int64_t m = std::min(n, num), m2 = m / 2 + 1;
    const bool even = (m % 2) == 0;
    const int mode = num < n ? 1 : (n < num ? 2 : 0);
    const int64_t nb = m + ((even && mode == 1) ? 1 : 0);
    double2* ykeep = nullptr;
    if (cudaMallocAsync((void**)&ykeep, (size_t)(m + 1) * sizeof(double2), st) != cudaSuccess) return -1;
    k_phase_table<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, st>>>(tabs, n, -1.0);
    k_dft_bins<<<(int)std::min<int64_t>(nb, 65535), 256, 0, st>>>(x, n, tabs, m2, m, nb, ykeep);
    k_phase_table<<<(int)std::min<int64_t>((num + 255) / 256, 1024), 256, 0, st>>>(tabs, num, 1.0);
    k_idft_bins<<<(int)std::min<int64_t>(num, 65535), 256, 0, st>>>(ykeep, n, num, tabs, m2, m, mode, y);
    launches += 4;
    cudaFreeAsync(ykeep, st);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace tetra
