// libtetra_b200.so -- host side of the C ABI declared in include/tetra_b200.h.
//
// Mirrors the control flow of the reference's SignalProcessor.process
// (tetraear/signal/processor.py:221-273) for a batch of carriers, choosing between
//   * the fused FIR-cascade kernel (k1_channelize_demod<MODE>, persistent) + exact edge windows on a second stream, when
//     fs = 2.4 MS/s, the block has >= 16384 samples and |freq_offset| <= 12.5 kHz (MODE 1 when any offset is non-zero,
//     MODE 2 for the channels of one wideband capture), and
//   * the exact-recursion kernel over the whole block otherwise,
// then the timing pick / slicer / TS correlator / sync cascade kernel.
// There is no CPU fallback: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../../include/tetra_b200.h"
#include "filter_design.h"
#include "tetra_kernels.cuh"
#include "tetra_exact.cuh"
#include "tetra_edges.cuh"
#include "tetra_edgecorr.cuh"
#include "tetra_pfb.cuh"
#include "tetra_finalize.cuh"
#include "tetra_stft.cuh"
#include "tetra_gather.cuh"

using namespace tetra;
static_assert(K1_EDGE == K_EDGE, "edge width of the fused and the exact kernels must agree");
constexpr int K1_MAX_GROUPS = 4;          // launches a large batch is cut into (finalize of group g beside the fused kernel of g + 1)

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {                      // page-locked host staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct tetra_ctx {
    int device = 0;
    double sample_rate = 2.4e6;
    cudaStream_t own_stream = nullptr, stream = nullptr, side = nullptr;
    cudaStream_t copy = nullptr;       // H2D copies of a chunked host batch (process_chunked)
    cudaEvent_t ev_h2d[3] = {nullptr, nullptr, nullptr};
    int64_t h2d_chunk_bytes = 0;       // tetra_set_h2d_chunk: 0 = default chunk, < 0 = host batches are not chunked
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_grp[4] = {nullptr, nullptr, nullptr, nullptr};   // fused-kernel group g done (K1_MAX_GROUPS)
    // per-launch CUDA-event pairs around the fused kernel (bench.py's roofline leg)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    bool timing = false;
    cudaEvent_t ph_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // call start, edges joined, finalize done (debug timeline)
    cudaEvent_t edge_ev[2] = {nullptr, nullptr};                   // around the edge kernel on the side stream
    bool ph_valid = false;
    int64_t launches = 0;
    std::string err;
    bool tables_uploaded = false;
    DevBuf in, y, partial, dib, ndib, sym, phase, match, fo, jobs, scr1, scrz, scr2, tmp_a, tmp_b, tmp_c, mats, wide, spos, u8, stft_tab;
    DevBuf sync_scr;                   // k_sync_positions_long: packed bits and hit masks of blocks longer than FIN_DIB_SMEM dibits
    DevBuf k1_ctr;                     // work-item counter of the fused kernel (one word, zeroed before every launch)
    DevBuf dout;                       // all results of a small host-side call, back to back (one D2H copy into `hout`)
    PinBuf hout;
    DevBuf pfb;                        // config 3: the 96 channelized 240 kS/s streams of one capture
    DevBuf etab, ecorr, estate;        // block-end correction tables (edge_tables_generated.h), corrections [C][2][K_EDGE], states
    DevBuf emat[10], emat_unit;        // states -> corrections matrices (freq_offset = 0), one per (n - 1) mod 10; unit states
    bool emat_built[10] = {false, false, false, false, false, false, false, false, false, false};
    EdgeTables etab_ptrs{};
    size_t max_scratch_bytes = (size_t)6 << 30;
    // peer-memory all-gather (tetra_gather.cuh)
    DevBuf p2p_buf, p2p_misc;          // receive buffer [2][world][block] + flags; ticket + status words
    uint8_t* p2p_peer[KG_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool p2p_ipc[KG_MAX_WORLD] = {false, false, false, false, false, false, false, false};   // opened with cudaIpcOpenMemHandle
    int p2p_rank = -1, p2p_world = 0;
    int64_t p2p_block = 0, p2p_flag_off = 0;
    uint32_t p2p_step = 0;
    bool p2p_connected = false;
    bool fg_active = false;            // tetra_process_batch_allgather: the finalize kernel pushes the streams to the peers
};

namespace {

int fail(tetra_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ctx, e_ == cudaErrorMemoryAllocation ? TETRA_E_NOMEM : TETRA_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// processor.py:245-250
int decim_factor(double fs) {
    if (fs > 240000.0 * 2) { int q = (int)(fs / 240000.0); if (q > 1) return q; }
    return 1;
}

struct Plan {
    int q; bool has_s1, has_s2; int64_t L; double rate; int sps, step; double wn;
};
Plan make_plan(double fs, int64_t n) {
    Plan p;
    p.q = decim_factor(fs);
    p.has_s1 = p.q > 1 && n > EX_PAD1;             // sosfiltfilt raises on n <= padlen -> decimation skipped
    if (!p.has_s1) p.q = 1;
    p.rate = p.has_s1 ? fs / p.q : fs;
    p.L = p.has_s1 ? (n + p.q - 1) / p.q : n;
    double wn = (25000.0 / 2) / (p.rate / 2);      // processor.py:74-75
    p.wn = std::min(0.99, std::max(0.01, wn));
    p.has_s2 = p.L > EX_PAD2;                      // filtfilt raises on len <= padlen -> unfiltered
    p.sps = (int)(p.rate / 18000.0);               // processor.py:183
    p.step = std::max(1, p.sps / 8);               // processor.py:194
    return p;
}

int upload_tables(tetra_ctx* ctx) {
    if (ctx->tables_uploaded) return 0;
    float fir[128];
    memset(fir, 0, sizeof fir);
    memcpy(fir + 1, TB_FIR120_TAPS, sizeof(float) * (2 * TB_FIR_H + 1));   // 127 taps behind one zero
    CK(cudaMemcpyToSymbol(c_proto, TB_PROTO_TAPS, sizeof(float) * (2 * TB_PROTO_H + 1)));
    {
        // byte input (MODE 3 / 4): taps / 127.5 and 127.5 x their running sums (tetra_kernels.cuh)
        float p8[2 * TB_PROTO_H + 1], pre[2 * TB_PROTO_H + 3];
        double run = 0.0;
        for (int d = 0; d <= 2 * TB_PROTO_H; ++d) p8[d] = (float)((double)TB_PROTO_TAPS[d] / 127.5);
        for (int k = 0; k <= 2 * TB_PROTO_H + 1; ++k) {
            pre[k] = (float)(127.5 * run);
            if (k <= 2 * TB_PROTO_H) run += (double)p8[k];
        }
        pre[2 * TB_PROTO_H + 2] = pre[2 * TB_PROTO_H + 1];
        CK(cudaMemcpyToSymbol(c_proto8, p8, sizeof p8));
        CK(cudaMemcpyToSymbol(c_proto8_pre, pre, sizeof pre));
    }
    CK(cudaMemcpyToSymbol(c_hb, TB_HB_TAPS, sizeof(float) * (2 * TB_HB_H + 1)));
    CK(cudaMemcpyToSymbol(c_fir, fir, sizeof fir));
    CK(cudaMemcpyToSymbol(c_interp, TB_INTERP_TAPS, sizeof(float) * TB_INT_K));
    CK(cudaMemcpyToSymbol(c_req, TB_REQ_CHEB, sizeof TB_REQ_CHEB));
    CK(cudaFuncSetAttribute(k1_channelize_demod<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem)));
    CK(cudaFuncSetAttribute(k1_channelize_demod<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemFo)));
    CK(cudaFuncSetAttribute(k1_channelize_demod<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemFo)));
    CK(cudaFuncSetAttribute(k1_channelize_demod<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemU8)));
    CK(cudaFuncSetAttribute(k1_channelize_demod<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemFoU8)));
    CK(cudaFuncSetAttribute(k1_channelize_demod<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem)));
    {
        float2 w96[PFB_NCH];
        for (int q = 0; q < PFB_NCH; ++q) {
            const double ang = -2.0 * M_PI * q / PFB_NCH;
            w96[q] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        CK(cudaMemcpyToSymbol(c_w96, w96, sizeof w96));
    }
    // block-end correction tables -> one device buffer
    {
        constexpr int NT = 11;
        const double* src[NT] = {ET_G1, ET_WC, ET_WAC, ET_RINGC, ET_RING, ET_U, ET_U2, ET_W0, ET_BP, ET_BC, ET_BV};
        const size_t cnt[NT] = {sizeof ET_G1 / sizeof(double), sizeof ET_WC / sizeof(double), sizeof ET_WAC / sizeof(double),
                                sizeof ET_RINGC / sizeof(double), sizeof ET_RING / sizeof(double), sizeof ET_U / sizeof(double),
                                sizeof ET_U2 / sizeof(double), sizeof ET_W0 / sizeof(double), sizeof ET_BP / sizeof(double),
                                sizeof ET_BC / sizeof(double), sizeof ET_BV / sizeof(double)};
        static_assert(sizeof ET_G1 == sizeof(double) * (2 * ET_G + 1) && sizeof ET_WC == sizeof(double) * 8 * ET_NC &&
                      sizeof ET_WAC == sizeof(double) * 8 * ET_NAC && sizeof ET_RINGC == sizeof(double) * 8 * ET_NRING &&
                      sizeof ET_RING == sizeof(double) * 8 * ET_NRING && sizeof ET_U == sizeof(double) * 64, "edge tables changed shape");
        size_t total = 0;
        for (size_t c : cnt) total += (c + 3) & ~(size_t)3;             // every table starts on a 32-byte boundary
        const size_t g1p_off = total, g1p_cnt = (size_t)(2 * ET_G + 1 + 2 * ET_G1_PAD);
        total += (g1p_cnt + 3) & ~(size_t)3;                            // g1 once more, between two runs of zeros
        CK(ctx->etab.ensure(total * sizeof(double)));
        const double* dev[NT];
        size_t off = 0;
        for (int k = 0; k < NT; ++k) {
            dev[k] = (const double*)ctx->etab.p + off;
            CK(cudaMemcpy((double*)ctx->etab.p + off, src[k], cnt[k] * sizeof(double), cudaMemcpyHostToDevice));
            off += (cnt[k] + 3) & ~(size_t)3;
        }
        CK(cudaMemset((double*)ctx->etab.p + g1p_off, 0, g1p_cnt * sizeof(double)));
        CK(cudaMemcpy((double*)ctx->etab.p + g1p_off + ET_G1_PAD, ET_G1, sizeof ET_G1, cudaMemcpyHostToDevice));
        ctx->etab_ptrs = EdgeTables{dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], dev[8], dev[9], dev[10],
                                    (const double*)ctx->etab.p + g1p_off + ET_G1_PAD};
    }
    // the copies above ran on the legacy default stream and may return with their last DMA in flight; the kernels use
    // non-blocking streams. Only that stream is waited for: another context's kernels may be spinning on this context's push.
    CK(cudaStreamSynchronize(cudaStreamLegacy));
    ctx->tables_uploaded = true;
    return 0;
}

// Where the block-end correction kernels run: beside the fused kernel on the side stream, or ahead of it on the same stream.
// Measured on one B200 box each (profiles/r02_edges_serial_ab.txt): beside it they cost the fused kernel more than they
// take alone once the batch is large (4096 carriers: fused kernel 6.11 -> 5.74 ms and step 6.46 -> 6.31 ms with them
// ahead; 0.26 ms alone), while small batches hide them for less (512 carriers: step 0.855 beside, 0.874 ahead).
// TETRA_EDGE_SERIAL = 0 / 1 forces one or the other.
constexpr int K1_EDGES_AHEAD_FROM = 2048;      // carriers per call from which the corrections run ahead
int edge_serial_mode(int n_carriers) {
    static const int v = getenv("TETRA_EDGE_SERIAL") ? atoi(getenv("TETRA_EDGE_SERIAL")) : -1;
    return v >= 0 ? v : (n_carriers >= K1_EDGES_AHEAD_FROM ? 1 : 0);
}

// block-end corrections of the fused path (k_edge_states + k_edge_recursions) for C carriers -> ctx->ecorr
int launch_edge_correct(tetra_ctx* ctx, cudaStream_t st, const float2* x, const uint8_t* x8, int64_t pitch, int64_t n, int32_t L,
                        const double* d_fo, const double* d_chan, double fs, double fs_dec, const ExactCoef& cf, int32_t C) {
    CK(ctx->ecorr.ensure((size_t)C * 2 * K_EDGE * sizeof(float2)));
    CK(ctx->estate.ensure((size_t)C * KC_NSTATE * sizeof(double2)));
    EdgeCorrArgs ea;
    ea.x = x; ea.x8 = x8; ea.pitch = pitch; ea.n = n; ea.L = L; ea.fo = d_fo; ea.chan = d_chan; ea.fs = fs; ea.fs_dec = fs_dec;
    ea.cf = cf; ea.t = ctx->etab_ptrs; ea.n_carriers = C; ea.states = (double2*)ctx->estate.p; ea.d = (float2*)ctx->ecorr.p;
    ea.m_out = nullptr; ea.m = nullptr;
    if (!d_fo) {
        // no freq_offset: states -> corrections is one real matrix per (n - 1) mod 10, built once by the recursion kernel on unit states
        const int k0 = (int)((n - 1) % 10);
        if (!ctx->emat_built[k0]) {
            CK(ctx->emat[k0].ensure((size_t)KC_NSTATE * 2 * K_EDGE * sizeof(double)));
            CK(ctx->emat_unit.ensure((size_t)KC_NSTATE * KC_NSTATE * sizeof(double2)));
            k_edge_unit_states<<<(KC_NSTATE * KC_NSTATE + 255) / 256, 256, 0, st>>>((double2*)ctx->emat_unit.p);
            EdgeCorrArgs eb = ea;
            eb.fo = nullptr; eb.n_carriers = KC_NSTATE; eb.states = (double2*)ctx->emat_unit.p; eb.m_out = (double*)ctx->emat[k0].p;
            k_edge_recursions<<<dim3((KC_NSTATE + KC2_THREADS - 1) / KC2_THREADS, 2), KC2_THREADS, 0, st>>>(eb);
            ctx->launches += 2;
            CK(cudaGetLastError());
            ctx->emat_built[k0] = true;
        }
        ea.m = (const double*)ctx->emat[k0].p;
    }
    const int g1 = d_fo ? (C + KC_THREADS / 32 - 1) / (KC_THREADS / 32) : C;      // with freq_offsets: one warp per carrier
    if (ea.fo) {
        if (x8) k_edge_states<1, true><<<g1, KC_THREADS, 0, st>>>(ea);
        else if (d_chan) k_edge_states<2, true><<<g1, KC_THREADS, 0, st>>>(ea);
        else k_edge_states<0, true><<<g1, KC_THREADS, 0, st>>>(ea);
    } else {
        if (x8) k_edge_states<1, false><<<g1, KC_THREADS, 0, st>>>(ea);
        else if (d_chan) k_edge_states<2, false><<<g1, KC_THREADS, 0, st>>>(ea);
        else k_edge_states<0, false><<<g1, KC_THREADS, 0, st>>>(ea);
    }
    if (d_fo) k_edge_recursions<<<dim3((C + KC2_THREADS - 1) / KC2_THREADS, 2), KC2_THREADS, 0, st>>>(ea);
    else k_edge_apply<<<dim3((C + KA_CPB - 1) / KA_CPB, 2), KA_THREADS, 0, st>>>(ea);
    ctx->launches++;
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

void fill_coef(ExactCoef& cf, int q, double wn) {
    memset(&cf, 0, sizeof cf);
    if (q > 1) {
        double sos[24], zi[8];
        design_cheby1_sos8(0.05, 0.8 / q, sos);
        sosfilt_zi(sos, 4, zi);
        memcpy(cf.sos, sos, sizeof sos);
        memcpy(cf.zi1, zi, sizeof zi);
    }
    design_butter4(wn, cf.b, cf.a);
    lfilter_zi(cf.b, cf.a, 5, cf.zi2);
}

// window sizes of an exact job (must match k_exact_chain)
void exact_extents(int mode, int64_t n, int64_t L, int q, int E, bool has_s1, int64_t* w1, int64_t* wz) {
    int64_t m_lo = 0, m_hi = L;
    if (mode == EX_LEFT) m_hi = std::min<int64_t>(L, E + EX_T2);
    else if (mode == EX_RIGHT) m_lo = std::max<int64_t>(0, L - E - EX_T2);
    *wz = m_hi - m_lo;
    if (!has_s1) { *w1 = 1; return; }
    int64_t tot = n + 2 * EX_PAD1, e_lo = 0, e_hi = tot;
    if (mode == EX_LEFT) e_hi = std::min<int64_t>(tot, EX_PAD1 + (int64_t)q * (m_hi - 1) + 1 + EX_T1);
    else if (mode == EX_RIGHT) e_lo = std::max<int64_t>(0, EX_PAD1 + (int64_t)q * m_lo - EX_T1);
    *w1 = e_hi - e_lo;
}

// ---- zero-input chunk transitions for k_exact_block (see tetra_exact.cuh) ----
void sos_transition_powers(const ExactCoef& cf, int steps, double* out);
void ba_transition_powers(const ExactCoef& cf, int steps, double* out);
template <int DIM>
double mat_norm_inf(const double* m) {
    double best = 0.0;
    for (int i = 0; i < DIM; ++i) {
        double r = 0.0;
        for (int j = 0; j < DIM; ++j) r += std::fabs(m[i * DIM + j]);
        best = std::max(best, r);
    }
    return best;
}

// whole blocks, one CTA per carrier, chunk-parallel passes (k_exact_block)
int launch_exact_block(tetra_ctx* ctx, cudaStream_t st, ExactArgs a, const std::vector<int2>& jobs) {
    const int64_t w1 = a.has_s1 ? a.n + 2 * EX_PAD1 : 1, wz = a.L, w2 = (int64_t)a.L + 2 * EX_PAD2;
    ExactBlockArgs b;
    memset(&b, 0, sizeof b);
    b.s1_stride = w1; b.sz_stride = wz; b.s2_stride = w2;
    // chunk lengths: as short as the thread count allows, as long as the transition over one chunk needs to die out
    auto pick = [&](int64_t T, bool sos, double* m_out) {
        int64_t lc = std::max<int64_t>(1, (T + EXB_THREADS - 1) / EXB_THREADS);
        double pw[5 * 64];
        for (;;) {
            if (sos) sos_transition_powers(a.cf, (int)lc, pw); else ba_transition_powers(a.cf, (int)lc, pw);
            const double nrm = sos ? mat_norm_inf<8>(pw) : mat_norm_inf<4>(pw);
            if (nrm <= 3e-3 || lc >= T) break;                // EXB_T = 8 terms: what is dropped is below 1e-20
            lc = std::min<int64_t>(T, lc * 2);
        }
        memcpy(m_out, pw, sizeof(double) * (sos ? 64 : 16));
        return (int32_t)lc;
    };
    b.lc1 = a.has_s1 ? pick(w1, true, b.m1) : 1;
    b.lc2 = a.has_s2 ? pick(w2, false, b.m2) : 1;
    const size_t per_job = (size_t)(w1 + wz + w2) * sizeof(double2);
    size_t chunk = std::max<size_t>(1, ctx->max_scratch_bytes / per_job);
    chunk = std::min(chunk, jobs.size());
    CK(ctx->scr1.ensure((size_t)w1 * chunk * sizeof(double2)));
    CK(ctx->scrz.ensure((size_t)wz * chunk * sizeof(double2)));
    CK(ctx->scr2.ensure((size_t)w2 * chunk * sizeof(double2)));
    CK(ctx->jobs.ensure(jobs.size() * sizeof(int2)));
    CK(cudaMemcpyAsync(ctx->jobs.p, jobs.data(), jobs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    CK(cudaFuncSetAttribute(k_exact_block, cudaFuncAttributeMaxDynamicSharedMemorySize, EXB_SMEM));
    a.scr1 = (double2*)ctx->scr1.p; a.scrz = (double2*)ctx->scrz.p; a.scr2 = (double2*)ctx->scr2.p;
    for (size_t off = 0; off < jobs.size(); off += chunk) {
        const int nj = (int)std::min(chunk, jobs.size() - off);
        a.jobs = (const int2*)ctx->jobs.p + off;
        a.n_jobs = nj;
        b.e = a;
        k_exact_block<<<nj, EXB_THREADS, EXB_SMEM, st>>>(b);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

constexpr int64_t EXB_MIN_N = 4096;        // shorter blocks: one thread per carrier is as fast as a CTA

int launch_exact(tetra_ctx* ctx, cudaStream_t st, ExactArgs a, const std::vector<int2>& jobs, int mode_hint) {
    if (jobs.empty()) return 0;
    {
        static const int blk_env = getenv("TETRA_EXACT_BLOCK") ? atoi(getenv("TETRA_EXACT_BLOCK")) : 1;   // 0: the serial kernel (A/B)
        bool all_full = true;
        for (auto& j : jobs) if (j.y != EX_FULL) { all_full = false; break; }
        if (blk_env && all_full && a.n >= EXB_MIN_N) return launch_exact_block(ctx, st, a, jobs);
    }
    // extents: all jobs of one launch share the worst-case window
    int64_t w1 = 1, wz = 1;
    for (int m = 0; m < 3; ++m) {
        bool used = false;
        for (auto& j : jobs) if (j.y == m) { used = true; break; }
        if (!used) continue;
        int64_t a1, az; exact_extents(m, a.n, a.L, a.q, a.edge, a.has_s1, &a1, &az);
        w1 = std::max(w1, a1); wz = std::max(wz, az);
    }
    (void)mode_hint;
    const size_t per_job = (size_t)(w1 + wz + wz + 2 * EX_PAD2) * sizeof(double2);
    size_t chunk = std::max<size_t>(1, ctx->max_scratch_bytes / per_job);
    chunk = std::min(chunk, jobs.size());
    CK(ctx->scr1.ensure((size_t)w1 * chunk * sizeof(double2)));
    CK(ctx->scrz.ensure((size_t)wz * chunk * sizeof(double2)));
    CK(ctx->scr2.ensure((size_t)(wz + 2 * EX_PAD2) * chunk * sizeof(double2)));
    CK(ctx->jobs.ensure(jobs.size() * sizeof(int2)));
    CK(cudaMemcpyAsync(ctx->jobs.p, jobs.data(), jobs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    a.scr1 = (double2*)ctx->scr1.p; a.scrz = (double2*)ctx->scrz.p; a.scr2 = (double2*)ctx->scr2.p;
    for (size_t off = 0; off < jobs.size(); off += chunk) {
        const int nj = (int)std::min(chunk, jobs.size() - off);
        a.jobs = (const int2*)ctx->jobs.p + off;
        a.n_jobs = nj;
        k_exact_chain<<<(nj + 63) / 64, 64, 0, st>>>(a);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    // the job list must outlive the kernels: it lives in ctx->jobs until the next call (same stream order)
    return 0;
}

// LEFT / RIGHT edge windows of the fast path: one thread per job, time-skewed biquad sections (k_exact_edges)
int launch_edges(tetra_ctx* ctx, cudaStream_t st, const ExactArgs& ea, const std::vector<int2>& jobs) {
    if (jobs.empty()) return 0;
    int64_t w1 = 1, wz = 1;
    for (int m = EX_LEFT; m <= EX_RIGHT; ++m) {
        int64_t a1, az; exact_extents(m, ea.n, ea.L, ea.q, ea.edge, true, &a1, &az);
        w1 = std::max(w1, a1); wz = std::max(wz, az);
    }
    const size_t nj = jobs.size();
    CK(ctx->scr1.ensure((size_t)w1 * nj * sizeof(double2)));
    CK(ctx->scrz.ensure((size_t)wz * nj * sizeof(double2)));
    CK(ctx->scr2.ensure((size_t)(wz + 2 * EX_PAD2) * nj * sizeof(double2)));
    CK(ctx->jobs.ensure(nj * sizeof(int2)));
    CK(cudaMemcpyAsync(ctx->jobs.p, jobs.data(), nj * sizeof(int2), cudaMemcpyHostToDevice, st));
    EdgeArgs g;
    g.x = ea.x32; g.pitch = ea.pitch; g.right_shift = ea.x_right_shift; g.n = ea.n; g.q = ea.q; g.L = ea.L; g.edge = ea.edge; g.cf = ea.cf;
    g.y = ea.y32; g.y_pitch = ea.y_pitch; g.y_sps = ea.y_sps; g.y_rows = ea.y_rows; g.jobs = (const int2*)ctx->jobs.p; g.n_jobs = (int32_t)nj;
    g.scr1 = (double2*)ctx->scr1.p; g.scrz = (double2*)ctx->scrz.p; g.scr2 = (double2*)ctx->scr2.p;
    g.w1 = w1; g.wz = wz; g.fo = ea.fo; g.fs_dec = ea.fs_dec;
    k_exact_edges<<<(int)((nj + EXT_THREADS - 1) / EXT_THREADS), EXT_THREADS, 0, st>>>(g);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}


// ---- zero-input chunk transitions over `steps` samples and their powers M^(2^r), r < 5 ----
template <int DIM>
void mat_mul(const double* a, const double* b, double* c) {
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) {
            double t = 0.0;
            for (int k = 0; k < DIM; ++k) t += a[i * DIM + k] * b[k * DIM + j];
            c[i * DIM + j] = t;
        }
}
// out[5][64]: M^(2^r), M = zero-input evolution of the 4-biquad cascade state (z0_0, z1_0, ...) over `steps` samples
void sos_transition_powers(const ExactCoef& cf, int steps, double* out) {
    double m[64];
    for (int col = 0; col < 8; ++col) {
        double z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        z[col] = 1.0;
        for (int t = 0; t < steps; ++t) {
            double x = 0.0;
            for (int k = 0; k < 4; ++k) {
                const double y = cf.sos[k][0] * x + z[2 * k];
                z[2 * k] = cf.sos[k][1] * x - cf.sos[k][4] * y + z[2 * k + 1];
                z[2 * k + 1] = cf.sos[k][2] * x - cf.sos[k][5] * y;
                x = y;
            }
        }
        for (int i = 0; i < 8; ++i) m[i * 8 + col] = z[i];
    }
    memcpy(out, m, sizeof m);
    for (int r = 1; r < 5; ++r) mat_mul<8>(out + (r - 1) * 64, out + (r - 1) * 64, out + r * 64);
}
void ba_transition_powers(const ExactCoef& cf, int steps, double* out) {
    double m[16];
    for (int col = 0; col < 4; ++col) {
        double z[4] = {0, 0, 0, 0};
        z[col] = 1.0;
        for (int t = 0; t < steps; ++t) {
            const double y = z[0];
            for (int k = 0; k < 3; ++k) z[k] = -cf.a[k + 1] * y + z[k + 1];
            z[3] = -cf.a[4] * y;
        }
        for (int i = 0; i < 4; ++i) m[i * 4 + col] = z[i];
    }
    memcpy(out, m, sizeof m);
    for (int r = 1; r < 5; ++r) mat_mul<4>(out + (r - 1) * 16, out + (r - 1) * 16, out + r * 16);
}

// decode()'s sync search over device-resident dibit streams: blocks up to FIN_DIB_SMEM dibits in shared memory, longer ones
// through global scratch
int launch_sync_positions(tetra_ctx* ctx, cudaStream_t st, const uint8_t* dibits, int64_t cap, const int32_t* n_dibits, int32_t C,
                          int32_t* sync_pos, int32_t max_pos, int32_t* n_sync) {
    if (cap <= FIN_DIB_SMEM) {
        SyncPosArgs a;
        a.dibits = dibits; a.cap = cap; a.n_dibits = n_dibits; a.sync_pos = sync_pos; a.max_pos = max_pos; a.n_sync = n_sync;
        k_sync_positions<<<C, FIN_THREADS, 0, st>>>(a);
    } else {
        SyncPosLongArgs a;
        a.dibits = dibits; a.cap = cap; a.n_dibits = n_dibits; a.sync_pos = sync_pos; a.max_pos = max_pos; a.n_sync = n_sync;
        a.words = cap / 16 + 2;
        CK(ctx->sync_scr.ensure((size_t)C * a.words * 2 * sizeof(uint32_t)));
        a.bits = (uint32_t*)ctx->sync_scr.p; a.mask = a.bits + (size_t)C * a.words;
        k_sync_positions_long<<<C, FIN_THREADS, 0, st>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TETRA_OK;
}

}  // namespace

extern "C" {

int tetra_create(tetra_ctx** out, int device, double sample_rate) {
    if (!out) return fail(nullptr, TETRA_E_INVALID, "tetra_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, TETRA_E_CUDA, "tetra_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, TETRA_E_INVALID, "tetra_create: device %d out of range", device);
    if (!(sample_rate > 0)) return fail(nullptr, TETRA_E_INVALID, "tetra_create: sample_rate must be > 0");
    tetra_ctx* ctx = new tetra_ctx();
    ctx->device = device;
    ctx->sample_rate = sample_rate;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, -1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) {
        fail(nullptr, TETRA_E_CUDA, "tetra_create: %s", cudaGetErrorString(e));
        delete ctx;
        return TETRA_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (const char* e = getenv("TETRA_H2D_CHUNK_MB")) {   // measurement switch: chunk size of host batches in MiB (< 0: never chunk)
        const long long mb = atoll(e);
        ctx->h2d_chunk_bytes = mb < 0 ? -1 : (int64_t)mb << 20;
    }
    *out = ctx;
    return TETRA_OK;
}

void tetra_destroy(tetra_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->side);
    DevBuf* bufs[] = {&ctx->in, &ctx->y, &ctx->partial, &ctx->dib, &ctx->ndib, &ctx->sym, &ctx->phase, &ctx->match,
                      &ctx->fo, &ctx->jobs, &ctx->scr1, &ctx->scrz, &ctx->scr2, &ctx->tmp_a, &ctx->tmp_b, &ctx->tmp_c, &ctx->mats, &ctx->wide, &ctx->spos, &ctx->u8, &ctx->stft_tab,
                      &ctx->etab, &ctx->ecorr, &ctx->estate, &ctx->pfb};
    tetra_p2p_destroy(ctx);
    for (DevBuf* b : bufs) b->release();
    ctx->dout.release(); ctx->hout.release(); ctx->k1_ctr.release(); ctx->sync_scr.release();
    for (DevBuf& b : ctx->emat) b.release();
    ctx->emat_unit.release();
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    for (int k = 0; k < 4; ++k) if (ctx->ev_grp[k]) cudaEventDestroy(ctx->ev_grp[k]);
    for (auto& pr : ctx->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (int k = 0; k < 4; ++k) if (ctx->ph_ev[k]) cudaEventDestroy(ctx->ph_ev[k]);
    for (int k = 0; k < 2; ++k) if (ctx->edge_ev[k]) cudaEventDestroy(ctx->edge_ev[k]);
    if (ctx->copy) { cudaStreamSynchronize(ctx->copy); cudaStreamDestroy(ctx->copy); }
    for (int k = 0; k < 3; ++k) if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
    cudaStreamDestroy(ctx->own_stream); cudaStreamDestroy(ctx->side);
    delete ctx;
}

const char* tetra_last_error(const tetra_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tetra_set_sample_rate(tetra_ctx* ctx, double sample_rate) {
    if (!ctx) return TETRA_E_INVALID;
    if (!(sample_rate > 0)) return fail(ctx, TETRA_E_INVALID, "sample_rate must be > 0");
    ctx->sample_rate = sample_rate;
    return TETRA_OK;
}
int tetra_set_stream(tetra_ctx* ctx, void* s) {
    if (!ctx) return TETRA_E_INVALID;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return TETRA_OK;
}
int tetra_synchronize(tetra_ctx* ctx) {
    if (!ctx) return TETRA_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return TETRA_OK;
}
int64_t tetra_dibit_capacity(const tetra_ctx* ctx, int64_t n) {
    if (!ctx || n <= 0) return 0;
    Plan p = make_plan(ctx->sample_rate, n);
    int64_t ns = p.sps > 1 ? p.L / p.sps : p.L;
    return ns > 1 ? ns - 1 : 0;
}
int64_t tetra_symbol_count(const tetra_ctx* ctx, int64_t n, int32_t best_phase) {
    if (!ctx || n <= 0) return 0;
    Plan p = make_plan(ctx->sample_rate, n);
    if (p.sps <= 1) return p.L;                        // processor.py:184-186: no timing pick below 2 samples per symbol
    return std::max<int64_t>(0, (p.L - best_phase) / p.sps);
}
int64_t tetra_launch_count(const tetra_ctx* ctx) { return ctx ? ctx->launches : 0; }
int tetra_enable_kernel_timing(tetra_ctx* ctx, int on) {
    if (!ctx) return TETRA_E_INVALID;
    ctx->timing = on != 0;
    ctx->ev_used = 0;
    return 0;
}
double tetra_kernel_time_ms(tetra_ctx* ctx, int32_t* n_launches) {
    if (n_launches) *n_launches = 0;
    if (!ctx || ctx->ev_used == 0) return -1.0;
    double total = 0.0;
    for (size_t i = 0; i < ctx->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->ev_pool[i].second) != cudaSuccess) return -1.0;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i].first, ctx->ev_pool[i].second) != cudaSuccess) return -1.0;
        total += ms;
    }
    if (n_launches) *n_launches = (int32_t)ctx->ev_used;
    ctx->ev_used = 0;
    return total;
}

int tetra_last_phase_ms(tetra_ctx* ctx, double* out3) {
    if (!ctx || !out3) return TETRA_E_INVALID;
    if (!ctx->ph_valid) return fail(ctx, TETRA_E_INVALID, "no timed fast-path call yet");
    CK(cudaEventSynchronize(ctx->ph_ev[3]));
    for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ph_ev[k], ctx->ph_ev[k + 1]));
        out3[k] = ms;
    }
    if (ctx->edge_ev[0]) {                              // slot 1 (an empty interval on the main stream): the edge kernel's span
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->edge_ev[0], ctx->edge_ev[1]) == cudaSuccess) out3[1] = ms;
        else cudaGetLastError();
    }
    return TETRA_OK;
}

int tetra_process_batch(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                        uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                        uint8_t* ts_match, int32_t async) {
    return tetra_process_batch_sync(ctx, iq, C, N, pitch, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match,
                                    nullptr, 0, nullptr, async);
}

// chan_hz == nullptr: C carriers [C][pitch]. chan_hz != nullptr (config 3): iq is ONE capture of N samples and carrier c is
// its channel at offset chan_hz[c], i.e. process(frequency_shift(iq, chan_hz[c]), 0); needs the fused path's conditions.
static int process_impl(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                        uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                        uint8_t* ts_match, int32_t* sync_pos, int32_t max_pos, int32_t* n_sync, int32_t async,
                        const double* chan_hz, const uint8_t* u8 = nullptr, int64_t u8_pitch = 0);

// A batch in chunks of carriers (carriers are independent: every per-carrier buffer is simply offset).
//  * Large HOST batches: the H2D copy of the chunks ahead runs on the context's copy stream beside the kernels and the D2H
//    copies of the current chunk, so the call costs the PCIe transfer plus ONE chunk's kernels instead of transfer + kernels
//    + result copies back to back (process_impl alone enqueues them in that order on one stream).
//  * More than CHUNK_MAX_CARRIERS carriers (the carrier index is a grid dimension of several kernels): host or device buffers.
// Returns CHUNK_NOT_APPLICABLE when the call should go through process_impl as it is.
constexpr int CHUNK_NOT_APPLICABLE = -1000;
constexpr int CHUNK_MAX_CARRIERS = 32768;
static int process_chunked(tetra_ctx* ctx, const void* in, bool is_u8, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                           uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                           uint8_t* ts_match, int32_t* sync_pos, int32_t max_pos, int32_t* n_sync, int32_t async) {
    if (!ctx || !in || C < 2 || N <= 0 || pitch < N || !n_dibits || cap < 0 || (cap > 0 && !dibits) || ctx->fg_active) return CHUNK_NOT_APPLICABLE;
    if ((sync_pos != nullptr) != (n_sync != nullptr)) return CHUNK_NOT_APPLICABLE;       // process_impl reports it
    const int64_t bps = is_u8 ? 2 : (int64_t)sizeof(float2);
    const bool host_in = !is_device_ptr(in);
    if (host_in && async) return CHUNK_NOT_APPLICABLE;
    int64_t Cc = C;
    bool pipelined = false;
    if (host_in && ctx->h2d_chunk_bytes >= 0) {
        // 32 MiB: measured on 2 GiB of complex64 / 512 MiB of bytes (profiles/r02_h2d_chunk_ab.txt): 16 ... 128 MiB chunks are within
        // 1 % of each other for complex64, bytes gain 8 % from 32 MiB over 128 MiB (the last chunk's kernels are what is left exposed)
        const int64_t chunk_bytes = ctx->h2d_chunk_bytes > 0 ? ctx->h2d_chunk_bytes : ((int64_t)32 << 20);
        const int64_t per = std::max<int64_t>(1, chunk_bytes / (N * bps));
        if ((C + per - 1) / per >= 2) { Cc = per; pipelined = true; }
    }
    Cc = std::min<int64_t>(Cc, CHUNK_MAX_CARRIERS);
    if (Cc >= C) return CHUNK_NOT_APPLICABLE;
    const int n_chunks = (int)((C + Cc - 1) / Cc);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CHUNK_NOT_APPLICABLE;
    auto call = [&](int64_t c0, int32_t cc, const void* chunk_in, int64_t chunk_pitch) {
        return process_impl(ctx, is_u8 ? nullptr : (const float*)chunk_in, cc, N, chunk_pitch, fo_hz ? fo_hz + c0 : nullptr,
                            dibits ? dibits + c0 * cap : nullptr, cap, n_dibits + c0, symbols ? symbols + c0 * (cap + 1) * 2 : nullptr,
                            best_phase ? best_phase + c0 : nullptr, ts_match ? ts_match + c0 * cap * 4 : nullptr,
                            sync_pos ? sync_pos + c0 * max_pos : nullptr, max_pos, n_sync ? n_sync + c0 : nullptr, async, nullptr,
                            is_u8 ? (const uint8_t*)chunk_in : nullptr, chunk_pitch);
    };
    if (!pipelined) {              // device input, or host input with the copy-ahead switched off: chunk by chunk, buffers as they are
        for (int g = 0; g < n_chunks; ++g) {
            const int64_t c0 = (int64_t)g * Cc;
            const int rc = call(c0, (int32_t)std::min<int64_t>(Cc, C - c0), (const uint8_t*)in + c0 * pitch * bps, pitch);
            if (rc) return rc;
        }
        return TETRA_OK;
    }
    DevBuf& buf = is_u8 ? ctx->u8 : ctx->in;
    CK(buf.ensure((size_t)C * N * bps));
    if (!ctx->copy) {
        CK(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
        for (int k = 0; k < 3; ++k) CK(cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
    }
    // the staging buffer may still be read by work enqueued earlier on the context's stream
    CK(cudaEventRecord(ctx->ev_h2d[2], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_h2d[2], 0));
    auto issue = [&](int g) -> cudaError_t {
        const int64_t c0 = (int64_t)g * Cc, cc = std::min<int64_t>(Cc, C - c0);
        uint8_t* dst = (uint8_t*)buf.p + c0 * N * bps;
        const uint8_t* src = (const uint8_t*)in + c0 * pitch * bps;
        cudaError_t e = pitch == N ? cudaMemcpyAsync(dst, src, (size_t)(cc * N * bps), cudaMemcpyHostToDevice, ctx->copy)
                                   : cudaMemcpy2DAsync(dst, (size_t)(N * bps), src, (size_t)(pitch * bps), (size_t)(N * bps), (size_t)cc,
                                                       cudaMemcpyHostToDevice, ctx->copy);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_h2d[g % 3], ctx->copy);
        return e;
    };
    CK(issue(0));
    CK(issue(1));
    int rc = TETRA_OK;
    for (int g = 0; g < n_chunks && rc == TETRA_OK; ++g) {
        const int64_t c0 = (int64_t)g * Cc;
        cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[g % 3], 0);
        if (e == cudaSuccess && g + 2 < n_chunks) e = issue(g + 2);     // two chunks ahead of the kernels
        if (e != cudaSuccess) { rc = fail(ctx, TETRA_E_CUDA, "chunked host batch: %s", cudaGetErrorString(e)); break; }
        rc = call(c0, (int32_t)std::min<int64_t>(Cc, C - c0), (const uint8_t*)buf.p + c0 * N * bps, N);
    }
    if (rc != TETRA_OK) cudaStreamSynchronize(ctx->copy);  // nothing of this call stays in flight behind an error
    return rc;
}

int tetra_set_h2d_chunk(tetra_ctx* ctx, int64_t bytes) {
    if (!ctx) return TETRA_E_INVALID;
    ctx->h2d_chunk_bytes = bytes;
    return TETRA_OK;
}

int tetra_process_batch_sync(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                             uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                             uint8_t* ts_match, int32_t* sync_pos, int32_t max_pos, int32_t* n_sync, int32_t async) {
    const int rc = process_chunked(ctx, iq, false, C, N, pitch, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match, sync_pos,
                                   max_pos, n_sync, async);
    if (rc != CHUNK_NOT_APPLICABLE) return rc;
    return process_impl(ctx, iq, C, N, pitch, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match, sync_pos, max_pos,
                        n_sync, async, nullptr);
}

static int process_impl(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                        uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                        uint8_t* ts_match, int32_t* sync_pos, int32_t max_pos, int32_t* n_sync, int32_t async,
                        const double* chan_hz, const uint8_t* u8, int64_t u8_pitch) {
    // u8 != null: the input is [C][u8_pitch][2] unsigned bytes on the device (RTL-SDR format) and `iq` is unused
    if (!ctx) return TETRA_E_INVALID;
    static const int edge_mode = getenv("TETRA_EDGE_MODE") ? atoi(getenv("TETRA_EDGE_MODE")) : 0;   // debug switch, see DESIGN.md §5
    static const bool fin_prefetch = getenv("TETRA_FIN_PREFETCH") ? atoi(getenv("TETRA_FIN_PREFETCH")) != 0 : true;   // A/B switch
    if ((sync_pos != nullptr) != (n_sync != nullptr) || (sync_pos && max_pos <= 0))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_sync: sync_pos, n_sync and max_positions go together");
    if (C < 0 || N < 0 || (C > 0 && N > 0 && (u8 ? u8_pitch < N : (!iq || pitch < N))) || !n_dibits || (cap > 0 && !dibits) || cap < 0)
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch: bad arguments");
    if (C > 65535) return fail(ctx, TETRA_E_INVALID, "tetra_process_batch: at most 65535 carriers per call");
    if (N > ((int64_t)1 << 30)) return fail(ctx, TETRA_E_INVALID, "tetra_process_batch: block too long");
    CK(cudaSetDevice(ctx->device));
    if (C == 0) return TETRA_OK;
    cudaStream_t st = ctx->stream;
    const bool d_in = u8 ? true : is_device_ptr(iq), d_dib = is_device_ptr(dibits), d_nd = is_device_ptr(n_dibits);
    const bool d_sym = is_device_ptr(symbols), d_ph = is_device_ptr(best_phase), d_match = is_device_ptr(ts_match);
    const bool d_spos = is_device_ptr(sync_pos), d_nsync = is_device_ptr(n_sync);
    if (async && !(d_in && (d_dib || !dibits) && d_nd && (d_sym || !symbols) && (d_ph || !best_phase) && (d_match || !ts_match) &&
                   (!sync_pos || (d_spos && d_nsync))))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch: async needs device buffers");
    if (sync_pos && d_spos != d_nsync) return fail(ctx, TETRA_E_INVALID, "sync_pos and n_sync must both be host or both device");

    const Plan pl = make_plan(ctx->sample_rate, N);
    const int64_t need_cap = tetra_dibit_capacity(ctx, N);
    if (cap < need_cap) return fail(ctx, TETRA_E_INVALID, "tetra_process_batch: cap %lld < required %lld", (long long)cap, (long long)need_cap);
    if (N == 0 || pl.L <= 0) {      // processor.py:239-241: empty in -> empty out
        if (d_nd) CK(cudaMemsetAsync(n_dibits, 0, sizeof(int32_t) * C, st)); else memset(n_dibits, 0, sizeof(int32_t) * C);
        if (best_phase) { if (d_ph) CK(cudaMemsetAsync(best_phase, 0, sizeof(int32_t) * C, st)); else memset(best_phase, 0, sizeof(int32_t) * C); }
        if (n_sync) { if (d_nsync) CK(cudaMemsetAsync(n_sync, 0, sizeof(int32_t) * C, st)); else memset(n_sync, 0, sizeof(int32_t) * C); }
        return TETRA_OK;
    }
    if (sync_pos && (int64_t)max_pos < (2 * cap) / 250 + 2)
        return fail(ctx, TETRA_E_INVALID, "max_positions too small: need at least %lld", (long long)((2 * cap) / 250 + 2));
    if (pl.sps > 1 && (pl.sps + pl.step - 1) / pl.step > FIN_MAXPH)
        return fail(ctx, TETRA_E_UNSUPPORTED, "timing search with more than %d phases", FIN_MAXPH);
    int rc = upload_tables(ctx);
    if (rc) return rc;

    // ---- input on device ----
    const float2* d_x = nullptr;
    int64_t x_pitch = pitch;
    if (u8) {
        // decided below: the fused kernel reads the bytes itself, every other path gets them expanded first
    } else if (chan_hz) {
        if (d_in) d_x = (const float2*)iq;
        else {
            CK(ctx->tmp_a.ensure((size_t)N * sizeof(float2)));
            CK(cudaMemcpyAsync(ctx->tmp_a.p, iq, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, st));
            d_x = (const float2*)ctx->tmp_a.p;
        }
        x_pitch = 0;                                   // every channel reads the same capture
    } else if (d_in) d_x = (const float2*)iq;
    else {
        CK(ctx->in.ensure((size_t)C * N * sizeof(float2)));
        if (pitch == N) CK(cudaMemcpyAsync(ctx->in.p, iq, (size_t)C * N * sizeof(float2), cudaMemcpyHostToDevice, st));
        else CK(cudaMemcpy2DAsync(ctx->in.p, N * sizeof(float2), iq, pitch * sizeof(float2), N * sizeof(float2), C, cudaMemcpyHostToDevice, st));
        d_x = (const float2*)ctx->in.p;
        x_pitch = N;
    }

    // ---- which carriers take the fused path ----
    const bool fast_ok = ctx->sample_rate == 2.4e6 && pl.q == 10 && pl.has_s1 && pl.has_s2 && N >= 16384 && pl.sps == K1_NPH;
    std::vector<int2> edge_jobs, full_jobs;
    bool any_fo = false, fo_in_range = true;
    for (int c = 0; c < C; ++c) {
        const double f = fo_hz ? fo_hz[c] : 0.0;
        if (f != 0.0) any_fo = true;
        if (!(std::fabs(f) <= K1_FO_MAX_HZ)) fo_in_range = false;     // also catches NaN
    }
    // The fused path covers freq_offset = 0 (MODE 0) and, with the NCO + equaliser of MODE 1, |freq_offset| <= 12.5 kHz
    // (MODE 1: the GUI's AFC range, ui/modern.py:1949-1967); anything else runs the exact recursion over the block.
    const bool use_fast = fast_ok && (!any_fo || fo_in_range);
    const bool u8_fused = u8 && use_fast && edge_mode != 2;       // with a freq_offset too (MODE 4): the GUI's live combination
    if (u8 && !u8_fused) {
        CK(ctx->wide.ensure((size_t)C * N * sizeof(float2)));
        if (u8_pitch == N) {
            k_u8_to_c64<<<(unsigned)std::min<int64_t>(((int64_t)C * N / 8 + 255) / 256 + 1, 148 * 16), 256, 0, st>>>(u8, (int64_t)C * N, (float2*)ctx->wide.p);
            ctx->launches++;
        } else {
            for (int c = 0; c < C; ++c) {
                k_u8_to_c64<<<(unsigned)std::min<int64_t>((N / 8 + 255) / 256 + 1, 148 * 4), 256, 0, st>>>(u8 + (size_t)c * u8_pitch * 2, N,
                                                                                                      (float2*)ctx->wide.p + (size_t)c * N);
                ctx->launches++;
            }
        }
        CK(cudaGetLastError());
        d_x = (const float2*)ctx->wide.p;
        x_pitch = N;
    }
    if (chan_hz && (!use_fast || any_fo)) return fail(ctx, TETRA_E_UNSUPPORTED, "wideband channels need the fused path (fs 2.4 MS/s, >= 16384 samples)");
    for (int c = 0; c < C; ++c) {
        if (use_fast) edge_jobs.push_back(make_int2(c, EX_LEFT));
        else full_jobs.push_back(make_int2(c, EX_FULL));
    }
    // LEFT jobs first, then RIGHT: the two window shapes differ in length, keep warps homogeneous
    for (size_t k = 0, n_left = edge_jobs.size(); k < n_left; ++k) edge_jobs.push_back(make_int2(edge_jobs[k].x, EX_RIGHT));
    const double* d_fo = nullptr;
    if (any_fo || chan_hz) {
        CK(ctx->fo.ensure(sizeof(double) * C));
        CK(cudaMemcpyAsync(ctx->fo.p, chan_hz ? chan_hz : fo_hz, sizeof(double) * C, cudaMemcpyHostToDevice, st));
        d_fo = (const double*)ctx->fo.p;
    }

    // ---- buffers ----
    // y is stored timing-phase major (y_index): sps rows of y_rows samples
    const int y_sps = pl.sps > 1 ? pl.sps : 1;
    const int y_rows = pl.sps > 1 ? (int)((((pl.L + y_sps - 1) / y_sps) + 15) & ~(int64_t)15) : 0;
    const int64_t y_pitch = y_rows > 0 ? (int64_t)y_sps * y_rows : ((pl.L + 63) & ~(int64_t)63);
    CK(ctx->y.ensure((size_t)C * y_pitch * sizeof(float2)));
    CK(ctx->phase.ensure(sizeof(int32_t) * C));
    uint8_t* k_dib = d_dib ? dibits : nullptr;
    int32_t* k_nd = d_nd ? n_dibits : nullptr;
    float2* k_sym = d_sym ? (float2*)symbols : nullptr;
    int32_t* k_ph = d_ph ? best_phase : nullptr;
    uint8_t* k_match = d_match ? ts_match : nullptr;
    int32_t* k_spos = d_spos ? sync_pos : nullptr;
    int32_t* k_nsync = d_nsync ? n_sync : nullptr;
    if (sync_pos && (int64_t)max_pos < (2 * cap) / 250 + 2)      // max_corr bookkeeping needs every position: the walk finds at most one per 250 bit offsets
        return fail(ctx, TETRA_E_INVALID, "max_positions too small: need at least %lld", (long long)((2 * cap) / 250 + 2));
    // A small call with every result in host memory (process() on one block): the results sit back to back in one device
    // buffer and come home in ONE copy into page-locked staging -- seven pageable copies cost more than the kernels.
    const size_t sz_dib = (size_t)C * cap, sz_nd = sizeof(int32_t) * C, sz_sym = symbols ? (size_t)C * (cap + 1) * sizeof(float2) : 0;
    const size_t sz_match = ts_match ? (size_t)C * cap * 4 : 0, sz_spos = sync_pos ? (size_t)C * max_pos * sizeof(int32_t) : 0;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_dib = 0, o_nd = o_dib + up16(sz_dib + 16), o_ph = o_nd + up16(sz_nd), o_sym = o_ph + up16(sz_nd);
    const size_t o_match = o_sym + up16(sz_sym), o_spos = o_match + up16(sz_match + 16), o_nsync = o_spos + up16(sz_spos);
    const size_t staged_total = o_nsync + up16(sz_nd);
    const bool staged = !async && !d_dib && !d_nd && !d_sym && !d_ph && !d_match && !d_spos && staged_total <= ((size_t)2 << 20);
    if (staged) {
        CK(ctx->dout.ensure(staged_total));
        CK(ctx->hout.ensure(staged_total));
        uint8_t* b = (uint8_t*)ctx->dout.p;
        k_dib = b + o_dib; k_nd = (int32_t*)(b + o_nd); k_ph = (int32_t*)(b + o_ph);
        if (symbols) k_sym = (float2*)(b + o_sym);
        if (ts_match) k_match = b + o_match;
        if (sync_pos) { k_spos = (int32_t*)(b + o_spos); k_nsync = (int32_t*)(b + o_nsync); }
    } else {
        if (!d_dib) { CK(ctx->dib.ensure((size_t)C * cap + 16)); k_dib = (uint8_t*)ctx->dib.p; }
        if (!d_nd) { CK(ctx->ndib.ensure(sizeof(int32_t) * C)); k_nd = (int32_t*)ctx->ndib.p; }
        if (symbols && !d_sym) { CK(ctx->sym.ensure((size_t)C * (cap + 1) * sizeof(float2))); k_sym = (float2*)ctx->sym.p; }
        if (best_phase && !d_ph) k_ph = (int32_t*)ctx->phase.p;
        if (ts_match && !d_match) { CK(ctx->match.ensure((size_t)C * cap * 4 + 16)); k_match = (uint8_t*)ctx->match.p; }
        if (sync_pos && !d_spos) {
            CK(ctx->spos.ensure((size_t)C * max_pos * sizeof(int32_t) + sizeof(int32_t) * C));
            k_spos = (int32_t*)ctx->spos.p;
            k_nsync = k_spos + (size_t)C * max_pos;
        }
    }

    ExactArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.x32 = d_x; ea.pitch = x_pitch; ea.n = N; ea.q = pl.q; ea.L = (int32_t)pl.L;
    ea.has_s1 = pl.has_s1; ea.has_s2 = pl.has_s2;
    fill_coef(ea.cf, pl.has_s1 ? pl.q : 1, pl.wn);
    ea.fo = chan_hz ? nullptr : d_fo; ea.fs_dec = pl.rate;
    // block ends of the fused path: corrections from the input alone (default), or -- TETRA_EDGE_MODE = 1 / 2 -- the
    // literal recursions of round 1 over windows at each end (kept for A/B measurements)
    const bool edge_corr = edge_mode == 0;
    if (chan_hz && !edge_corr) {
        // the exact edge kernels take each channel's shifted stream; only the block-end windows they read are formed
        CK(ctx->wide.ensure((size_t)C * N * sizeof(float2)));
        int64_t w1 = 0, wz = 0, w1r = 0;
        exact_extents(EX_LEFT, N, pl.L, pl.q, K1_EDGE, true, &w1, &wz);
        exact_extents(EX_RIGHT, N, pl.L, pl.q, K1_EDGE, true, &w1r, &wz);
        const int64_t len0 = std::min<int64_t>(N, w1 + 64), len1 = std::min<int64_t>(N, w1r + 64);
        k_mix_wide_ranges<<<dim3((unsigned)((len0 + len1 + 255) / 256), C), 256, 0, st>>>(d_x, N, d_fo, ctx->sample_rate, (float2*)ctx->wide.p,
                                                                                         0, len0, N - len1, len1);
        ctx->launches++;
        CK(cudaGetLastError());
        ea.x32 = (const float2*)ctx->wide.p; ea.pitch = N;
    }
    if (u8_fused && !edge_corr) {
        // the exact edge kernels read complex64: expand the two end windows of every block, compactly
        const EdgeRange rl = edge_range(EX_LEFT, N, (int)pl.L, pl.q, K1_EDGE), rr = edge_range(EX_RIGHT, N, (int)pl.L, pl.q, K1_EDGE);
        const int64_t wl = std::min<int64_t>(N, ((rl.e_hi - EX_PAD1) + 7) & ~(int64_t)7);
        const int64_t wr = std::min<int64_t>(N, (N - std::max<int64_t>(0, rr.e_lo - EX_PAD1) + 7) & ~(int64_t)7);
        CK(ctx->wide.ensure((size_t)C * (wl + wr) * sizeof(float2)));
        k_u8_edge_windows<<<dim3((unsigned)((wl + wr + 255) / 256), C), 256, 0, st>>>(u8, u8_pitch, N, (int32_t)wl, (int32_t)wr, (float2*)ctx->wide.p);
        ctx->launches++;
        CK(cudaGetLastError());
        ea.x32 = (const float2*)ctx->wide.p; ea.pitch = wl + wr; ea.x_right_shift = wl - (N - wr);
    }
    ea.y32 = (float2*)ctx->y.p; ea.y_pitch = y_pitch; ea.y_sps = y_sps; ea.y_rows = y_rows; ea.edge = K1_EDGE;

    FinArgs fa;
    memset(&fa, 0, sizeof fa);
    int fin_c0 = 0;                                    // first carrier the in-order finalize launch still has to do
    fa.y = (const float2*)ctx->y.p; fa.y_pitch = y_pitch; fa.y_rows = y_rows; fa.L = (int32_t)pl.L; fa.sps = pl.sps; fa.step = pl.step;
    fa.dibits = k_dib; fa.cap = cap; fa.n_dibits = k_nd; fa.symbols = k_sym; fa.best_phase = k_ph;
    fa.phase_scratch = (int32_t*)ctx->phase.p;
    if (ctx->fg_active) {
        for (int r = 0; r < ctx->p2p_world; ++r) fa.push.recv[r] = ctx->p2p_peer[r];
        fa.push.world = ctx->p2p_world; fa.push.rank = ctx->p2p_rank; fa.push.n_local = C;
        fa.push.step = ctx->p2p_step;
        fa.push.slot_off = ((int64_t)(ctx->p2p_step & 1) * ctx->p2p_world + ctx->p2p_rank) * ctx->p2p_block;
        fa.push.flag_off = ctx->p2p_flag_off;
        fa.push.ticket = (uint32_t*)ctx->p2p_misc.p;
    }

    const bool fused_match = ts_match && cap > 0 && cap <= FIN_DIB_SMEM;
    fa.match = fused_match ? k_match : nullptr;
    const bool fused_sync = sync_pos && cap <= FIN_DIB_SMEM;
    fa.sync_pos = fused_sync ? k_spos : nullptr; fa.max_pos = max_pos; fa.n_sync = k_nsync;

    if (use_fast) {
        // segments: enough CTAs to fill the machine when there are few carriers
        // Work items = (carrier, segment), streamed back to back by min(SMs, items) persistent CTAs: the launch takes
        // ceil(items / CTAs) slots of (tiles per segment + 1) iterations; take the cheapest split of a carrier.
        const int tiles = (int)((pl.L + K1_W - 1) / K1_W);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        int n_seg = 1;
        double best_cost = 1e300;
        for (int cand = 1; cand <= std::min(tiles, 64); ++cand) {
            const int seg_tiles = (tiles + cand - 1) / cand;
            if (cand > 1 && seg_tiles < 8) break;              // keep slots much longer than the pipeline
            const int real = (tiles + seg_tiles - 1) / seg_tiles;
            const int64_t items = (int64_t)C * real;
            const int64_t ctas = std::min<int64_t>(sms, items);
            const double cost = (double)((items + ctas - 1) / ctas) * std::max(seg_tiles + 1, K1_MIN_T_ITEM) + 6;
            if (cost < best_cost * 0.999) { best_cost = cost; n_seg = cand; }
        }
        int seg_len = ((tiles + n_seg - 1) / n_seg) * K1_W;
        n_seg = (int)((pl.L + seg_len - 1) / seg_len);
        const int n_items = C * n_seg;
        const int k1_grid = std::min(sms, n_items);
        CK(ctx->partial.ensure((size_t)C * n_seg * 16 * sizeof(double)));
        K1Args ka;
        ka.x = d_x; ka.pitch = x_pitch; ka.n = N; ka.L = (int32_t)pl.L; ka.seg_len = seg_len; ka.n_seg = n_seg; ka.n_items = n_items; ka.t_item = std::max(seg_len / K1_W + 1, K1_MIN_T_ITEM);
        ka.y = (float2*)ctx->y.p; ka.y_pitch = y_pitch; ka.y_rows = y_rows; ka.partial = (double*)ctx->partial.p;
        ka.aligned = ((reinterpret_cast<uintptr_t>(d_x) & 15) == 0) && ((x_pitch & 1) == 0);
        ka.x8 = nullptr;
        if (u8_fused) {
            ka.x8 = u8; ka.pitch = u8_pitch;
            ka.aligned = ((reinterpret_cast<uintptr_t>(u8) & 15) == 0) && ((u8_pitch & 7) == 0);
        }
        ka.fo = d_fo; ka.fs_dec = pl.rate; ka.fs = ctx->sample_rate;
        // config 3 on the 25 kHz grid of a 2.4 MS/s capture: the proto stage of ALL channels is one polyphase DFT (k_pfb96) and
        // the fused kernel runs its remaining stages on the 96 channelized streams (MODE 5); TETRA_PFB=0 keeps the per-channel
        // modulated proto (MODE 2), which is also what off-grid channel lists get
        static const int pfb_env = getenv("TETRA_PFB") ? atoi(getenv("TETRA_PFB")) : 1;
        bool pfb = chan_hz != nullptr && pfb_env != 0 && edge_corr;
        for (int c = 0; c < C && pfb; ++c) {
            const double k = chan_hz[c] * (PFB_NCH / ctx->sample_rate), kr = std::nearbyint(k);   // on the grid up to rounding of the product
            if (std::fabs(k - kr) > 1e-9 || kr < -PFB_NCH / 2 || kr >= PFB_NCH / 2) pfb = false;
        }
        ka.w_col0 = 0;
        bool forked = false;
        if (pfb) {
            const int64_t cols = PFB_M0 + (int64_t)n_seg * seg_len + 2 * K1_W;           // every w index a slot's tiles touch
            const int64_t wp = (cols + PFB_MB - 1) / PFB_MB * PFB_MB;
            CK(ctx->pfb.ensure((size_t)PFB_NCH * wp * sizeof(float2)));
            // the block-end corrections read the capture itself: let the side stream start ahead of the channelizer
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            forked = true;
            PfbArgs pa;
            pa.x = d_x; pa.n = N; pa.w = (float2*)ctx->pfb.p; pa.wp = wp;
            k_pfb96<<<(unsigned)(wp / PFB_MB), PFB_THREADS, 0, st>>>(pa);
            ctx->launches++;
            CK(cudaGetLastError());
            ka.x = (const float2*)ctx->pfb.p; ka.pitch = wp; ka.w_col0 = PFB_M0; ka.aligned = 1;
        }
        ka.zero_ext = edge_corr ? 1 : 0;
        CK(ctx->k1_ctr.ensure(64));
        ka.counter = (uint32_t*)ctx->k1_ctr.p;
        // edge windows go to the side stream: the thread-per-job kernel runs beside the bulk kernel, the warp-per-job one behind it
        if (ctx->timing) {
            if (!ctx->ph_ev[0]) {
                for (int k = 0; k < 4; ++k) CK(cudaEventCreate(&ctx->ph_ev[k]));
                for (int k = 0; k < 2; ++k) CK(cudaEventCreate(&ctx->edge_ev[k]));
            }
            CK(cudaEventRecord(ctx->ph_ev[0], st));
        }
        // large batches: the block-end corrections run ahead of the fused kernel on the same stream (edge_serial_mode)
        const int edge_serial = edge_serial_mode(C);
        if (edge_corr && edge_serial) {
            if (ctx->timing && ctx->edge_ev[0]) CK(cudaEventRecord(ctx->edge_ev[0], st));
            rc = launch_edge_correct(ctx, st, u8_fused ? nullptr : d_x, u8_fused ? u8 : nullptr, u8_fused ? u8_pitch : x_pitch, N,
                                     (int32_t)pl.L, chan_hz ? nullptr : d_fo, chan_hz ? d_fo : nullptr, ctx->sample_rate, pl.rate, ea.cf, C);
            if (rc) return rc;
            if (ctx->timing && ctx->edge_ev[0]) CK(cudaEventRecord(ctx->edge_ev[1], st));
        }
        if (!forked) {
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        }
        // TETRA_K1_GROUPS=g (measurement switch, default 1): cut a large batch into g launches of whole slots (multiples of the
        // CTA count, so the cut costs no extra slot) and run the finalize kernel of group k on the side stream beside the fused
        // kernel of group k + 1 (its 256-thread, 20 KB CTAs fit next to a fused-kernel CTA). Measured at 4096 carriers x 2^20
        // (profiles/r02_groups_ab.txt): the serial finalize shrinks from 0.37 to 0.10 ms, the fused kernel slows by 0.19 ms
        // (both want HBM), the step gains 0.6 % -- not worth a lower roofline fraction of the dominant kernel, so off by default.
        static const int grp_env = getenv("TETRA_K1_GROUPS") ? atoi(getenv("TETRA_K1_GROUPS")) : 1;
        int n_grp = 1;
        if (n_seg == 1 && edge_corr && grp_env > 1) {
            const int slots = (C + sms - 1) / sms;
            n_grp = std::max(1, std::min(std::min(grp_env, K1_MAX_GROUPS), slots));
        }
        const int grp_slots = ((C + sms - 1) / sms + n_grp - 1) / n_grp;
        const int grp_car = n_grp > 1 ? grp_slots * sms : C;            // carriers per group (the last one takes what is left)
        n_grp = n_grp > 1 ? (C + grp_car - 1) / grp_car : 1;
        for (int g = 0; g < K1_MAX_GROUPS && n_grp > 1; ++g)
            if (!ctx->ev_grp[g]) CK(cudaEventCreateWithFlags(&ctx->ev_grp[g], cudaEventDisableTiming));
        fa.partial = (const double*)ctx->partial.p; fa.n_seg = n_seg;
        fa.bulk_lo = K1_EDGE; fa.bulk_hi = (int32_t)pl.L - K1_EDGE;
        if (edge_corr) {                                  // sized here: the finalize launches below take the pointer
            CK(ctx->ecorr.ensure((size_t)C * 2 * K_EDGE * sizeof(float2)));
            fa.edge_corr = (const float2*)ctx->ecorr.p;
        }
        const bool side_edges = !(edge_corr && edge_serial);
        for (int g = 0; g < n_grp; ++g) {
            const int c_lo = g * grp_car, c_hi = std::min(C, c_lo + grp_car);
            ka.item0 = c_lo * n_seg; ka.n_items = (c_hi - c_lo) * n_seg;
            CK(cudaMemsetAsync(ka.counter, 0, sizeof(uint32_t), st));
            const int grid_g = std::min(sms, ka.n_items);
            // the fused kernel goes first: its persistent CTAs (one per SM) must not queue behind the edge blocks
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (ctx->timing && ctx->ev_used < 4096) {
                if (ctx->ev_used == ctx->ev_pool.size()) {
                    cudaEvent_t a = nullptr, b = nullptr;
                    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
                    ctx->ev_pool.emplace_back(a, b);
                }
                t0 = ctx->ev_pool[ctx->ev_used].first; t1 = ctx->ev_pool[ctx->ev_used].second;
                ctx->ev_used++;
                CK(cudaEventRecord(t0, st));
            }
            if (u8_fused && any_fo) k1_channelize_demod<4><<<grid_g, K1_THREADS, sizeof(K1SmemFoU8), st>>>(ka);
            else if (u8_fused) k1_channelize_demod<3><<<grid_g, K1_THREADS, sizeof(K1SmemU8), st>>>(ka);
            else if (pfb) k1_channelize_demod<5><<<grid_g, K1_THREADS, sizeof(K1Smem), st>>>(ka);
            else if (chan_hz) k1_channelize_demod<2><<<grid_g, K1_THREADS, sizeof(K1SmemFo), st>>>(ka);
            else if (any_fo) k1_channelize_demod<1><<<grid_g, K1_THREADS, sizeof(K1SmemFo), st>>>(ka);
            else k1_channelize_demod<0><<<grid_g, K1_THREADS, sizeof(K1Smem), st>>>(ka);
            ctx->launches++;
            CK(cudaGetLastError());
            if (t1) CK(cudaEventRecord(t1, st));
            if (g + 1 < n_grp) CK(cudaEventRecord(ctx->ev_grp[g], st));
            if (g == 0) {
                // block ends on the side stream, beside the fused kernel (TETRA_EDGE_MODE = 1 / 2: the literal recursions of
                // round 1 -- thread per job with time-skewed sections, or the plain sequential kernel)
                if (side_edges && ctx->timing && ctx->edge_ev[0]) CK(cudaEventRecord(ctx->edge_ev[0], ctx->side));
                if (!side_edges) rc = 0;
                else if (edge_corr)
                    rc = launch_edge_correct(ctx, ctx->side, u8_fused ? nullptr : d_x, u8_fused ? u8 : nullptr, u8_fused ? u8_pitch : x_pitch, N,
                                             (int32_t)pl.L, chan_hz ? nullptr : d_fo, chan_hz ? d_fo : nullptr, ctx->sample_rate, pl.rate, ea.cf, C);
                else if (edge_mode == 2) rc = launch_exact(ctx, ctx->side, ea, edge_jobs, 0);
                else rc = launch_edges(ctx, ctx->side, ea, edge_jobs);
                if (rc) return rc;
                if (side_edges && ctx->timing && ctx->edge_ev[0]) CK(cudaEventRecord(ctx->edge_ev[1], ctx->side));
            } else {
                // finalize of the previous group, beside this group's fused kernel
                const int p_lo = (g - 1) * grp_car;
                CK(cudaStreamWaitEvent(ctx->side, ctx->ev_grp[g - 1], 0));
                fa.car0 = p_lo;
                if (fin_prefetch) k_finalize<true><<<c_lo - p_lo, FIN_THREADS, 0, ctx->side>>>(fa);
                else k_finalize<false><<<c_lo - p_lo, FIN_THREADS, 0, ctx->side>>>(fa);
                ctx->launches++;
                CK(cudaGetLastError());
            }
        }
        fin_c0 = (n_grp - 1) * grp_car;
        CK(cudaEventRecord(ctx->ev_join, ctx->side));
        CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        if (ctx->timing && ctx->ph_ev[0]) CK(cudaEventRecord(ctx->ph_ev[1], st));
    } else {
        rc = launch_exact(ctx, st, ea, full_jobs, 0);
        if (rc) return rc;
        fa.partial = nullptr; fa.n_seg = 0; fa.bulk_lo = 0; fa.bulk_hi = 0;
    }
    if (ctx->timing && use_fast && ctx->ph_ev[0]) CK(cudaEventRecord(ctx->ph_ev[2], st));
    fa.car0 = fin_c0;
    if (fin_prefetch) k_finalize<true><<<C - fin_c0, FIN_THREADS, 0, st>>>(fa);
    else k_finalize<false><<<C - fin_c0, FIN_THREADS, 0, st>>>(fa);
    ctx->launches++;
    CK(cudaGetLastError());
    if (ts_match && cap > 0 && !fused_match) {
        SyncArgs sa; sa.dibits = k_dib; sa.cap = cap; sa.n_dibits = k_nd; sa.match = k_match;
        const int gx = (int)std::min<int64_t>(64, (2 * cap + 255) / 256);
        k_sync_match<<<dim3(std::max(gx, 1), C), 256, 0, st>>>(sa);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (sync_pos && !fused_sync) {                      // blocks too long for the fused front end's shared memory
        rc = launch_sync_positions(ctx, st, k_dib, cap, k_nd, C, k_spos, max_pos, k_nsync);
        if (rc) return rc;
    }
    if (ctx->timing && use_fast && ctx->ph_ev[0]) { CK(cudaEventRecord(ctx->ph_ev[3], st)); ctx->ph_valid = true; }
    // ---- results to host buffers ----
    if (staged) {
        CK(cudaMemcpyAsync(ctx->hout.p, ctx->dout.p, staged_total, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint8_t* h = (const uint8_t*)ctx->hout.p;
        if (dibits) memcpy(dibits, h + o_dib, sz_dib);
        memcpy(n_dibits, h + o_nd, sz_nd);
        if (best_phase) memcpy(best_phase, h + o_ph, sz_nd);
        if (symbols) memcpy(symbols, h + o_sym, sz_sym);
        if (ts_match) memcpy(ts_match, h + o_match, sz_match);
        if (sync_pos) { memcpy(sync_pos, h + o_spos, sz_spos); memcpy(n_sync, h + o_nsync, sz_nd); }
        return TETRA_OK;
    }
    if (dibits && !d_dib) CK(cudaMemcpyAsync(dibits, k_dib, (size_t)C * cap, cudaMemcpyDeviceToHost, st));
    if (!d_nd) CK(cudaMemcpyAsync(n_dibits, k_nd, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
    if (symbols && !d_sym) CK(cudaMemcpyAsync(symbols, k_sym, (size_t)C * (cap + 1) * sizeof(float2), cudaMemcpyDeviceToHost, st));
    if (best_phase && !d_ph) CK(cudaMemcpyAsync(best_phase, k_ph, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
    if (ts_match && !d_match) CK(cudaMemcpyAsync(ts_match, k_match, (size_t)C * cap * 4, cudaMemcpyDeviceToHost, st));
    if (sync_pos && !d_spos) {
        CK(cudaMemcpyAsync(sync_pos, k_spos, (size_t)C * max_pos * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(n_sync, k_nsync, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
    }
    if (!async) CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_analyze_signal(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, double* out6) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || N < 0 || (C > 0 && (!out6 || (N > 0 && (!iq || pitch < N)))))
        return fail(ctx, TETRA_E_INVALID, "tetra_analyze_signal: bad arguments");
    if (C == 0) return TETRA_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int ds = std::max(1, (int)(ctx->sample_rate / 18000.0 / 10.0));      // scanner.py:108
    const int64_t n_bits = (N + ds - 1) / ds;
    if (n_bits > ANA_MAXBITS) return fail(ctx, TETRA_E_UNSUPPORTED, "tetra_analyze_signal: capture too long (%lld crude bits > %d)", (long long)n_bits, ANA_MAXBITS);
    const float2* dx = (const float2*)iq;
    int64_t dpitch = pitch;
    if (N > 0 && !is_device_ptr(iq)) {
        CK(ctx->in.ensure((size_t)C * N * sizeof(float2)));
        if (pitch == N) CK(cudaMemcpyAsync(ctx->in.p, iq, (size_t)C * N * sizeof(float2), cudaMemcpyHostToDevice, st));
        else CK(cudaMemcpy2DAsync(ctx->in.p, N * sizeof(float2), iq, pitch * sizeof(float2), N * sizeof(float2), C, cudaMemcpyHostToDevice, st));
        dx = (const float2*)ctx->in.p;
        dpitch = N;
    }
    const bool d_out = is_device_ptr(out6);
    double* dout = out6;
    if (!d_out) { CK(ctx->tmp_c.ensure(sizeof(double) * 6 * C)); dout = (double*)ctx->tmp_c.p; }
    const size_t smem = ((size_t)(n_bits + 31) / 32 + 2) * sizeof(uint32_t);
    CK(cudaFuncSetAttribute(k_analyze, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_analyze<<<C, ANA_THREADS, smem, st>>>(dx, dpitch, N, ds, dout, nullptr, 6);
    ctx->launches++;
    CK(cudaGetLastError());
    if (!d_out) CK(cudaMemcpyAsync(out6, dout, sizeof(double) * 6 * C, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_survey_wideband(tetra_ctx* ctx, const float* iq, int64_t N, const double* channel_hz, int32_t C, int32_t nfft, double* out) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || N < 0 || (C > 0 && (!channel_hz || !out || (N > 0 && !iq))) || nfft < 64 || nfft > 4096 || (nfft & (nfft - 1)))
        return fail(ctx, TETRA_E_INVALID, "tetra_survey_wideband: bad arguments (nfft a power of two in [64, 4096])");
    if (C == 0) return TETRA_OK;
    if (C > 65535) return fail(ctx, TETRA_E_INVALID, "tetra_survey_wideband: at most 65535 channels per call");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const double fs = ctx->sample_rate;
    const int ds = std::max(1, (int)(fs / 18000.0 / 10.0));                   // scanner.py:108
    const int64_t n_bits = (N + ds - 1) / ds;
    if (n_bits > ANA_MAXBITS) return fail(ctx, TETRA_E_UNSUPPORTED, "tetra_survey_wideband: capture too long (%lld crude bits > %d)", (long long)n_bits, ANA_MAXBITS);
    const float2* dx = (const float2*)iq;
    if (N > 0 && !is_device_ptr(iq)) {
        CK(ctx->in.ensure((size_t)N * sizeof(float2)));
        CK(cudaMemcpyAsync(ctx->in.p, iq, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, st));
        dx = (const float2*)ctx->in.p;
    }
    // channel offsets and the phase step each one takes off: [C] Hz, [C] rad
    std::vector<double> hz(2 * (size_t)C);
    for (int c = 0; c < C; ++c) { hz[c] = channel_hz[c]; hz[C + c] = (2.0 * M_PI) * channel_hz[c] / fs; }
    CK(ctx->fo.ensure(sizeof(double) * 2 * C));
    CK(cudaMemcpyAsync(ctx->fo.p, hz.data(), sizeof(double) * 2 * C, cudaMemcpyHostToDevice, st));
    const double* d_hz = (const double*)ctx->fo.p;
    CK(ctx->tmp_c.ensure(sizeof(double) * TETRA_SURVEY_FIELDS * C));
    double* dout = (double*)ctx->tmp_c.p;
    CK(cudaMemsetAsync(dout, 0, sizeof(double) * TETRA_SURVEY_FIELDS * C, st));
    const size_t smem = ((size_t)(n_bits + 31) / 32 + 2) * sizeof(uint32_t);
    CK(cudaFuncSetAttribute(k_analyze, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_analyze<<<C, ANA_THREADS, smem, st>>>(dx, 0, N, ds, dout, d_hz + C, TETRA_SURVEY_FIELDS);
    ctx->launches++;
    CK(cudaGetLastError());
    if (N >= nfft) {                                                         // ui/modern.py:1922: only with a full FFT block
        CK(ctx->tmp_a.ensure((size_t)C * nfft * sizeof(double2)));
        CK(ctx->tmp_b.ensure((size_t)C * nfft * sizeof(double)));
        k_mix_head_f64<<<dim3((nfft + 255) / 256, C), 256, 0, st>>>(dx, nfft, d_hz, fs, (double2*)ctx->tmp_a.p);
        if (stft_f64_launch(st, (const double2*)ctx->tmp_a.p, nfft, nfft, C, (double*)ctx->tmp_b.p))
            return fail(ctx, TETRA_E_CUDA, "tetra_survey_wideband: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        k_presence<<<C, 256, 0, st>>>((const double*)ctx->tmp_b.p, nfft, fs, dout + 6, TETRA_SURVEY_FIELDS);
        ctx->launches += 3;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(out, dout, sizeof(double) * TETRA_SURVEY_FIELDS * C, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    hz.clear();
    return TETRA_OK;
}

int tetra_process_batch_u8(tetra_ctx* ctx, const uint8_t* iq_u8, int32_t C, int64_t N, int64_t pitch,
                           const double* fo_hz, uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols,
                           int32_t* best_phase, uint8_t* ts_match, int32_t* sync_pos, int32_t max_pos, int32_t* n_sync) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || N < 0 || (C > 0 && N > 0 && (!iq_u8 || pitch < N)))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_u8: bad arguments");
    if (C == 0 || N == 0)
        return tetra_process_batch_sync(ctx, nullptr, C, N, N, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match,
                                        sync_pos, max_pos, n_sync, 0);
    {
        const int rc = process_chunked(ctx, iq_u8, true, C, N, pitch, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match, sync_pos,
                                       max_pos, n_sync, 0);
        if (rc != CHUNK_NOT_APPLICABLE) return rc;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t* d_in = iq_u8;
    int64_t d_pitch = pitch;
    if (!is_device_ptr(iq_u8)) {
        CK(ctx->u8.ensure((size_t)C * N * 2));
        if (pitch == N) CK(cudaMemcpyAsync(ctx->u8.p, iq_u8, (size_t)C * N * 2, cudaMemcpyHostToDevice, st));
        else CK(cudaMemcpy2DAsync(ctx->u8.p, N * 2, iq_u8, pitch * 2, N * 2, C, cudaMemcpyHostToDevice, st));
        d_in = (const uint8_t*)ctx->u8.p;
        d_pitch = N;
    }
    // the fused path reads the bytes as they are (2 bytes per sample); other sample rates / offsets expand them first
    return process_impl(ctx, nullptr, C, N, N, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match, sync_pos, max_pos, n_sync, 0,
                        nullptr, d_in, d_pitch);
}

int tetra_process_wideband(tetra_ctx* ctx, const float* iq, int64_t N, const double* channel_hz, int32_t C,
                           uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                           uint8_t* ts_match) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || N < 0 || (C > 0 && (!channel_hz || !n_dibits)) || (C > 0 && N > 0 && !iq))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_wideband: bad arguments");
    if (C == 0) return TETRA_OK;
    if (C > 65535) return fail(ctx, TETRA_E_INVALID, "tetra_process_wideband: at most 65535 channels per call");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (N == 0) return tetra_process_batch(ctx, nullptr, C, 0, 0, nullptr, dibits, cap, n_dibits, symbols, best_phase, ts_match, 0);
    // fused: the capture is read as is, each channel's shift is folded into stage A of the fused kernel (MODE 2)
    {
        const Plan pl = make_plan(ctx->sample_rate, N);
        const bool fast_ok = ctx->sample_rate == 2.4e6 && pl.q == 10 && pl.has_s1 && pl.has_s2 && N >= 16384 && pl.sps == K1_NPH;
        if (fast_ok)
            return process_impl(ctx, iq, C, N, N, nullptr, dibits, cap, n_dibits, symbols, best_phase, ts_match, nullptr, 0, nullptr, 0,
                                channel_hz);
    }
    // otherwise: expand the capture to C baseband streams and run them as ordinary carriers
    const float2* dx = (const float2*)iq;
    if (!is_device_ptr(iq)) {
        CK(ctx->tmp_a.ensure((size_t)N * sizeof(float2)));
        CK(cudaMemcpyAsync(ctx->tmp_a.p, iq, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, st));
        dx = (const float2*)ctx->tmp_a.p;
    }
    CK(ctx->fo.ensure(sizeof(double) * C));
    CK(cudaMemcpyAsync(ctx->fo.p, channel_hz, sizeof(double) * C, cudaMemcpyHostToDevice, st));
    CK(ctx->wide.ensure((size_t)C * N * sizeof(float2)));
    k_mix_wide<<<dim3((unsigned)std::min<int64_t>((N + 256 * MIX_RUN - 1) / (256 * MIX_RUN), 1024), C), 256, 0, st>>>(dx, N, (const double*)ctx->fo.p,
                                                                                            ctx->sample_rate, (float2*)ctx->wide.p);
    ctx->launches++;
    CK(cudaGetLastError());
    // every channel is now an ordinary carrier at baseband: process(shifted, 0)
    return tetra_process_batch(ctx, (const float*)ctx->wide.p, C, N, N, nullptr, dibits, cap, n_dibits, symbols, best_phase, ts_match, 0);
}

int tetra_pack_dibits(tetra_ctx* ctx, const uint8_t* dibits, int64_t n, uint8_t* packed) {
    if (!ctx) return TETRA_E_INVALID;
    if (n < 0 || (n & 15) || (n > 0 && (!dibits || !packed)) || ((reinterpret_cast<uintptr_t>(dibits) & 15) != 0) ||
        ((reinterpret_cast<uintptr_t>(packed) & 3) != 0))
        return fail(ctx, TETRA_E_INVALID, "tetra_pack_dibits: n must be a multiple of 16, dibits 16-byte and packed 4-byte aligned");
    if (n == 0) return TETRA_OK;
    if (!is_device_ptr(dibits) || !is_device_ptr(packed)) return fail(ctx, TETRA_E_INVALID, "tetra_pack_dibits: device buffers only");
    CK(cudaSetDevice(ctx->device));
    const int64_t n16 = n / 16;
    k_pack_dibits<<<(unsigned)std::min<int64_t>((n16 + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>((const uint4*)dibits, n16, (uint32_t*)packed);
    ctx->launches++;
    CK(cudaGetLastError());
    return TETRA_OK;
}

int tetra_unpack_dibits(tetra_ctx* ctx, const uint8_t* packed, int64_t n_blocks, int64_t packed_bytes, int64_t in_stride,
                        uint8_t* dibits, int64_t out_stride) {
    if (!ctx) return TETRA_E_INVALID;
    if (n_blocks < 0 || n_blocks > 65535 || packed_bytes < 0 || (packed_bytes & 3) || (in_stride & 3) || (out_stride & 15) ||
        in_stride < packed_bytes || out_stride < 4 * packed_bytes || (n_blocks > 0 && packed_bytes > 0 && (!packed || !dibits)) ||
        ((reinterpret_cast<uintptr_t>(packed) & 3) != 0) || ((reinterpret_cast<uintptr_t>(dibits) & 15) != 0))
        return fail(ctx, TETRA_E_INVALID, "tetra_unpack_dibits: bad sizes or alignment");
    if (n_blocks == 0 || packed_bytes == 0) return TETRA_OK;
    if (!is_device_ptr(dibits) || !is_device_ptr(packed)) return fail(ctx, TETRA_E_INVALID, "tetra_unpack_dibits: device buffers only");
    CK(cudaSetDevice(ctx->device));
    const int64_t words = packed_bytes / 4;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((words + 255) / 256, (148 * 8 + n_blocks - 1) / n_blocks));
    k_unpack_dibits<<<dim3(gx, (unsigned)n_blocks), 256, 0, ctx->stream>>>(packed, words, in_stride, dibits, out_stride);
    ctx->launches++;
    CK(cudaGetLastError());
    return TETRA_OK;
}

int tetra_sync_positions(tetra_ctx* ctx, const uint8_t* dibits, int64_t cap, const int32_t* n_dibits, int32_t C,
                         int32_t* sync_pos, int32_t max_pos, int32_t* n_sync) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || cap < 0 || (C > 0 && (!dibits || !n_dibits || !sync_pos || !n_sync)) || max_pos <= 0)
        return fail(ctx, TETRA_E_INVALID, "tetra_sync_positions: bad arguments");
    if (C == 0) return TETRA_OK;
    if ((int64_t)max_pos < (2 * cap) / 250 + 2) return fail(ctx, TETRA_E_INVALID, "max_positions too small: need at least %lld", (long long)((2 * cap) / 250 + 2));
    if (cap > ((int64_t)1 << 29)) return fail(ctx, TETRA_E_INVALID, "tetra_sync_positions: streams too long");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool dev_in = is_device_ptr(dibits), dev_nd = is_device_ptr(n_dibits), dev_out = is_device_ptr(sync_pos);
    if (dev_out != is_device_ptr(n_sync)) return fail(ctx, TETRA_E_INVALID, "sync_pos and n_sync must both be host or both device");
    SyncPosArgs a;
    a.dibits = dibits; a.cap = cap; a.n_dibits = n_dibits; a.sync_pos = sync_pos; a.max_pos = max_pos; a.n_sync = n_sync;
    if (!dev_in) {
        CK(ctx->dib.ensure((size_t)C * cap + 16));
        CK(cudaMemcpyAsync(ctx->dib.p, dibits, (size_t)C * cap, cudaMemcpyHostToDevice, st));
        a.dibits = (const uint8_t*)ctx->dib.p;
    }
    if (!dev_nd) {
        CK(ctx->ndib.ensure(sizeof(int32_t) * C));
        CK(cudaMemcpyAsync(ctx->ndib.p, n_dibits, sizeof(int32_t) * C, cudaMemcpyHostToDevice, st));
        a.n_dibits = (const int32_t*)ctx->ndib.p;
    }
    if (!dev_out) {
        CK(ctx->spos.ensure((size_t)C * max_pos * sizeof(int32_t) + sizeof(int32_t) * C));
        a.sync_pos = (int32_t*)ctx->spos.p;
        a.n_sync = a.sync_pos + (size_t)C * max_pos;
    }
    {
        const int rc = launch_sync_positions(ctx, st, a.dibits, cap, a.n_dibits, C, a.sync_pos, max_pos, a.n_sync);
        if (rc) return rc;
    }
    if (!dev_out) {
        CK(cudaMemcpyAsync(sync_pos, a.sync_pos, (size_t)C * max_pos * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(n_sync, a.n_sync, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_parse_bursts(tetra_ctx* ctx, const uint8_t* dibits, int64_t cap, const int32_t* n_dibits, int32_t C,
                       const int32_t* sync_pos, int32_t max_pos, const int32_t* n_sync, int32_t* burst_info) {
    if (!ctx) return TETRA_E_INVALID;
    if (C < 0 || cap < 0 || max_pos <= 0 || (C > 0 && (!dibits || !n_dibits || !sync_pos || !n_sync || !burst_info)))
        return fail(ctx, TETRA_E_INVALID, "tetra_parse_bursts: bad arguments");
    if (C == 0) return TETRA_OK;
    if (C > 65535) return fail(ctx, TETRA_E_INVALID, "tetra_parse_bursts: at most 65535 carriers per call");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool dev = is_device_ptr(dibits);
    if (dev != is_device_ptr(n_dibits) || dev != is_device_ptr(sync_pos) || dev != is_device_ptr(n_sync) || dev != is_device_ptr(burst_info))
        return fail(ctx, TETRA_E_INVALID, "tetra_parse_bursts: all buffers must be host or all device");
    BurstArgs a;
    a.cap = cap; a.max_pos = max_pos;
    if (dev) {
        a.dibits = dibits; a.n_dibits = n_dibits; a.sync_pos = sync_pos; a.n_sync = n_sync; a.info = (int4*)burst_info;
    } else {
        const size_t b_dib = ((size_t)C * cap + 15) & ~(size_t)15, b_pos = (size_t)C * max_pos * sizeof(int32_t);
        CK(ctx->tmp_a.ensure(b_dib + 2 * sizeof(int32_t) * C + b_pos + (size_t)C * max_pos * sizeof(int4) + 64));
        uint8_t* base = (uint8_t*)ctx->tmp_a.p;
        CK(cudaMemcpyAsync(base, dibits, (size_t)C * cap, cudaMemcpyHostToDevice, st));
        int32_t* d_nd = (int32_t*)(base + b_dib);
        int32_t* d_ns = d_nd + C;
        int32_t* d_pos = d_ns + C;
        CK(cudaMemcpyAsync(d_nd, n_dibits, sizeof(int32_t) * C, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_ns, n_sync, sizeof(int32_t) * C, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_pos, sync_pos, b_pos, cudaMemcpyHostToDevice, st));
        size_t off = b_dib + 2 * sizeof(int32_t) * C + b_pos;
        off = (off + 15) & ~(size_t)15;
        a.dibits = base; a.n_dibits = d_nd; a.n_sync = d_ns; a.sync_pos = d_pos; a.info = (int4*)(base + off);
    }
    k_parse_bursts<<<dim3((max_pos + 3) / 4, C), 128, 0, st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    if (!dev) CK(cudaMemcpyAsync(burst_info, a.info, (size_t)C * max_pos * sizeof(int4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

// ------------------------------------------------------------------------------------------------
// find_sync replay (host, integer-exact)
// ------------------------------------------------------------------------------------------------
int tetra_find_sync(const uint8_t* match, int64_t nw, double threshold, int32_t* positions, int32_t max_pos, double* max_corr_out) {
    if (nw < 0 || (nw > 0 && !match) || (max_pos > 0 && !positions)) return TETRA_E_INVALID;
    int count = 0;
    double max_corr = 0.0;
    std::vector<std::pair<int64_t, double>> visited;
    visited.reserve((size_t)nw);
    int64_t i = 0;
    while (i < nw) {
        bool found = false;
        double best_here = 0.0;
        for (int k = 0; k < 2; ++k) {
            const double corr = (double)match[2 * i + k] / 22.0;
            best_here = std::max(best_here, corr);
            max_corr = std::max(max_corr, corr);
            if (corr >= threshold) {
                if (count < max_pos) positions[count] = (int32_t)i;
                ++count; found = true;
                break;
            }
        }
        if (best_here > 0) visited.emplace_back(i, best_here);
        i = found ? i + 250 : i + 1;
    }
    if (count == 0 && max_corr > 0.75 && max_corr >= (threshold - 0.15)) {
        const double adaptive = std::max(0.75, max_corr - 0.02);
        if (adaptive < threshold) {
            std::vector<char> blocked((size_t)nw, 0);
            for (auto& pc : visited) {
                if (pc.second >= adaptive && !blocked[(size_t)pc.first]) {
                    if (count < max_pos) positions[count] = (int32_t)pc.first;
                    ++count;
                    const int64_t lo = std::max<int64_t>(0, pc.first - 250), hi = std::min<int64_t>(nw, pc.first + 250);
                    for (int64_t t = lo; t < hi; ++t) blocked[(size_t)t] = 1;
                }
            }
        }
    }
    if (max_corr_out) *max_corr_out = max_corr;
    return count;
}

int tetra_sync_cascade(const uint8_t* match, int64_t nw, int32_t* positions, int32_t max_pos) {
    double mx = 0.0;
    int n = tetra_find_sync(match, nw, 0.90, positions, max_pos, &mx);
    if (n != 0) return n;
    n = tetra_find_sync(match, nw, 0.85, positions, max_pos, &mx);
    if (n != 0) return n;
    n = tetra_find_sync(match, nw, 0.80, positions, max_pos, &mx);
    if (n != 0) return n;
    if (mx >= 0.75) n = tetra_find_sync(match, nw, std::max(0.75, mx - 0.02), positions, max_pos, &mx);
    return n;
}

// ------------------------------------------------------------------------------------------------
// helper entry points (complex128 host in/out)
// ------------------------------------------------------------------------------------------------
static int run_exact_c128(tetra_ctx* ctx, const double* in, int64_t n, bool filter, double wn, double fo, double fs, double* out) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->tmp_a.ensure((size_t)n * sizeof(double2)));
    CK(ctx->tmp_b.ensure((size_t)n * sizeof(double2)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, in, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st));
    ExactArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.x64 = (const double2*)ctx->tmp_a.p; ea.pitch = n; ea.n = n; ea.q = 1; ea.L = (int32_t)n;
    ea.has_s1 = 0; ea.has_s2 = filter ? 1 : 0;
    fill_coef(ea.cf, 1, filter ? wn : 0.5);
    if (fo != 0.0) {
        CK(ctx->fo.ensure(sizeof(double)));
        CK(cudaMemcpyAsync(ctx->fo.p, &fo, sizeof(double), cudaMemcpyHostToDevice, st));
        ea.fo = (const double*)ctx->fo.p;
    }
    ea.fs_dec = fs;
    ea.y64 = (double2*)ctx->tmp_b.p; ea.y_pitch = n; ea.edge = 0;
    std::vector<int2> jobs(1, make_int2(0, EX_FULL));
    int rc = launch_exact(ctx, st, ea, jobs, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->tmp_b.p, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_filter_signal(tetra_ctx* ctx, const double* in, int64_t n, double bandwidth, double sample_rate, double* out) {
    if (!ctx) return TETRA_E_INVALID;
    if (n < 0 || (n > 0 && (!in || !out)) || !(sample_rate > 0)) return fail(ctx, TETRA_E_INVALID, "tetra_filter_signal: bad arguments");
    if (n == 0) return TETRA_OK;
    if (n > ((int64_t)1 << 30)) return fail(ctx, TETRA_E_INVALID, "too long");
    double wn = (bandwidth / 2) / (sample_rate / 2);
    wn = std::min(0.99, std::max(0.01, wn));
    if (!(wn == wn) || n <= EX_PAD2) {            // reference: exception inside filtfilt -> unfiltered (processor.py:81-83)
        memcpy(out, in, (size_t)n * 2 * sizeof(double));
        return 1;
    }
    return run_exact_c128(ctx, in, n, true, wn, 0.0, sample_rate, out);
}

int tetra_frequency_shift(tetra_ctx* ctx, const double* in, int64_t n, double fo, double fs, double* out) {
    if (!ctx) return TETRA_E_INVALID;
    if (n < 0 || (n > 0 && (!in || !out)) || !(fs > 0)) return fail(ctx, TETRA_E_INVALID, "tetra_frequency_shift: bad arguments");
    if (n == 0) return TETRA_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->tmp_a.ensure((size_t)n * sizeof(double2)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, in, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st));
    k_nco_c128<<<(int)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>((double2*)ctx->tmp_a.p, n, fo, fs);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->tmp_a.p, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_extract_symbols(tetra_ctx* ctx, const double* in, int64_t n, double fs, double* out, int64_t* n_out, int32_t* best_phase) {
    if (!ctx) return TETRA_E_INVALID;
    if (n < 0 || (n > 0 && (!in || !out)) || !n_out || !(fs > 0)) return fail(ctx, TETRA_E_INVALID, "tetra_extract_symbols: bad arguments");
    *n_out = 0;
    if (best_phase) *best_phase = 0;
    if (n == 0) return TETRA_OK;
    const int sps = (int)(fs / 18000.0);
    if (sps <= 1) { memcpy(out, in, (size_t)n * 2 * sizeof(double)); *n_out = n; return TETRA_OK; }
    const int step = std::max(1, sps / 8);
    if ((sps + step - 1) / step > FIN_MAXPH) return fail(ctx, TETRA_E_UNSUPPORTED, "too many timing phases");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->tmp_a.ensure((size_t)n * sizeof(double2)));
    CK(ctx->tmp_b.ensure((size_t)(n / sps + 2) * sizeof(double2)));
    CK(ctx->tmp_c.ensure(2 * sizeof(int64_t)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, in, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st));
    k_extract_c128<<<1, FIN_THREADS, 0, st>>>((const double2*)ctx->tmp_a.p, n, sps, step, (double2*)ctx->tmp_b.p, (int64_t*)ctx->tmp_c.p);
    ctx->launches++;
    CK(cudaGetLastError());
    int64_t res[2];
    CK(cudaMemcpyAsync(res, ctx->tmp_c.p, sizeof res, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = res[0];
    if (best_phase) *best_phase = (int32_t)res[1];
    if (res[0] > 0) CK(cudaMemcpy(out, ctx->tmp_b.p, (size_t)res[0] * sizeof(double2), cudaMemcpyDeviceToHost));
    return TETRA_OK;
}

int tetra_demodulate_dqpsk(tetra_ctx* ctx, const double* in, int64_t n, uint8_t* out, int64_t* n_out) {
    if (!ctx) return TETRA_E_INVALID;
    if (n < 0 || (n > 0 && !in) || !n_out || (n > 1 && !out)) return fail(ctx, TETRA_E_INVALID, "tetra_demodulate_dqpsk: bad arguments");
    *n_out = 0;
    if (n < 2) return TETRA_OK;                    // processor.py:120-121
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->tmp_a.ensure((size_t)n * sizeof(double2)));
    CK(ctx->tmp_b.ensure((size_t)n));
    CK(ctx->tmp_c.ensure(sizeof(double)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, in, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->tmp_c.p, 0, sizeof(double), st));
    const int g = (int)std::min<int64_t>((n + 255) / 256, 2048);
    k_maxabs_c128<<<g, 256, 0, st>>>((const double2*)ctx->tmp_a.p, n, (unsigned long long*)ctx->tmp_c.p);
    k_slice_c128<<<g, 256, 0, st>>>((const double2*)ctx->tmp_a.p, n, (const double*)ctx->tmp_c.p, (uint8_t*)ctx->tmp_b.p);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->tmp_b.p, (size_t)(n - 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = n - 1;
    return TETRA_OK;
}

int tetra_resample(tetra_ctx* ctx, const double* in, int64_t n, int64_t n_out, double* out) {
    if (!ctx) return TETRA_E_INVALID;
    if (n <= 0 || n_out <= 0 || !in || !out) return fail(ctx, TETRA_E_INVALID, "tetra_resample: bad arguments");
    if (n > (1 << 22) || n_out > (1 << 22)) return fail(ctx, TETRA_E_UNSUPPORTED, "tetra_resample: at most 2^22 samples");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // Fourier resampling (scipy.signal.resample, complex input): X = DFT(x); keep the lowest
    // min(n, n_out) bins (split around DC, Nyquist bin handled as SciPy does); y = IDFT * (n_out / n).
    CK(ctx->tmp_a.ensure((size_t)n * sizeof(double2)));
    CK(ctx->tmp_b.ensure((size_t)std::max(n, n_out) * sizeof(double2)));
    CK(ctx->tmp_c.ensure((size_t)n_out * sizeof(double2)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, in, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st));
    int rc = resample_c128(ctx->launches, st, (const double2*)ctx->tmp_a.p, n, (double2*)ctx->tmp_b.p, n_out, (double2*)ctx->tmp_c.p);
    if (rc) return fail(ctx, TETRA_E_CUDA, "tetra_resample: kernel launch failed");
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->tmp_c.p, (size_t)n_out * sizeof(double2), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_stft_db(tetra_ctx* ctx, const float* iq, int64_t n, int32_t nfft, int32_t hop, float* out, int64_t* rows_out) {
    if (!ctx) return TETRA_E_INVALID;
    if (!rows_out || n < 0 || hop <= 0 || nfft < 64 || nfft > 8192 || (nfft & (nfft - 1)))
        return fail(ctx, TETRA_E_INVALID, "tetra_stft_db: nfft must be a power of two in [64, 8192], hop > 0");
    const int64_t rows = n >= nfft ? (n - nfft) / hop + 1 : 0;
    *rows_out = rows;
    if (rows == 0) return TETRA_OK;
    if (!iq || !out) return fail(ctx, TETRA_E_INVALID, "tetra_stft_db: null buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool d_in = is_device_ptr(iq), d_out = is_device_ptr(out);
    const float2* dx = (const float2*)iq;
    float* dout = out;
    if (!d_in) {
        CK(ctx->in.ensure((size_t)n * sizeof(float2)));
        CK(cudaMemcpyAsync(ctx->in.p, iq, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
        dx = (const float2*)ctx->in.p;
    }
    if (!d_out) { CK(ctx->y.ensure((size_t)rows * nfft * sizeof(float))); dout = (float*)ctx->y.p; }
    if (nfft == 4096 && !ctx->stft_tab.p) {
        CK(ctx->stft_tab.ensure(sizeof(S4kTables)));
        k_stft4096_tables<<<S4K_N / 256, 256, 0, st>>>((S4kTables*)ctx->stft_tab.p);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    int rc = stft_launch(st, dx, n, nfft, hop, rows, dout, (const S4kTables*)ctx->stft_tab.p);
    ctx->launches++;
    if (rc) return fail(ctx, TETRA_E_CUDA, "tetra_stft_db: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaGetLastError());
    if (!d_out) CK(cudaMemcpyAsync(out, dout, (size_t)rows * nfft * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_stft_db_f64(tetra_ctx* ctx, const double* iq, int64_t n, int32_t nfft, int32_t hop, double* out, int64_t* rows_out) {
    if (!ctx) return TETRA_E_INVALID;
    if (!rows_out || n < 0 || hop <= 0 || nfft < 64 || nfft > 4096 || (nfft & (nfft - 1)))
        return fail(ctx, TETRA_E_INVALID, "tetra_stft_db_f64: nfft must be a power of two in [64, 4096], hop > 0");
    const int64_t rows = n >= nfft ? (n - nfft) / hop + 1 : 0;
    *rows_out = rows;
    if (rows == 0) return TETRA_OK;
    if (!iq || !out) return fail(ctx, TETRA_E_INVALID, "tetra_stft_db_f64: null buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t used = (rows - 1) * hop + nfft;
    CK(ctx->tmp_a.ensure((size_t)used * sizeof(double2)));
    CK(ctx->tmp_b.ensure((size_t)rows * nfft * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->tmp_a.p, iq, (size_t)used * sizeof(double2), cudaMemcpyHostToDevice, st));
    if (stft_f64_launch(st, (const double2*)ctx->tmp_a.p, nfft, hop, rows, (double*)ctx->tmp_b.p))
        return fail(ctx, TETRA_E_CUDA, "tetra_stft_db_f64: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->tmp_b.p, (size_t)rows * nfft * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

int tetra_edge_corrections(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz, float* out) {
    if (!ctx) return TETRA_E_INVALID;
    if (C <= 0 || !iq || !out || pitch < N) return fail(ctx, TETRA_E_INVALID, "tetra_edge_corrections: bad arguments");
    const Plan pl = make_plan(2.4e6, N);
    if (N < 16384 || N > ((int64_t)1 << 30)) return fail(ctx, TETRA_E_UNSUPPORTED, "tetra_edge_corrections: the fused path needs 16384 <= n <= 2^30");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_tables(ctx);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx->in.ensure((size_t)C * N * sizeof(float2)));
    CK(cudaMemcpy2DAsync(ctx->in.p, N * sizeof(float2), iq, pitch * sizeof(float2), N * sizeof(float2), C, cudaMemcpyHostToDevice, st));
    const double* d_fo = nullptr;
    if (fo_hz) {
        CK(ctx->fo.ensure(sizeof(double) * C));
        CK(cudaMemcpyAsync(ctx->fo.p, fo_hz, sizeof(double) * C, cudaMemcpyHostToDevice, st));
        d_fo = (const double*)ctx->fo.p;
    }
    ExactCoef cf;
    fill_coef(cf, 10, pl.wn);
    rc = launch_edge_correct(ctx, st, (const float2*)ctx->in.p, nullptr, N, N, (int32_t)pl.L, d_fo, nullptr, 2.4e6, 240000.0, cf, C);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->ecorr.p, (size_t)C * 2 * K_EDGE * sizeof(float2), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TETRA_OK;
}

// ------------------------------------------------------------------------------------------------
// peer-memory all-gather of the dibit streams (tetra_gather.cuh)
// ------------------------------------------------------------------------------------------------
int tetra_p2p_create(tetra_ctx* ctx, int32_t rank, int32_t world, int64_t block_bytes, uint8_t* handle_out) {
    if (!ctx) return TETRA_E_INVALID;
    if (world < 1 || world > KG_MAX_WORLD || rank < 0 || rank >= world || block_bytes <= 0 || (block_bytes & 15))
        return fail(ctx, TETRA_E_INVALID, "tetra_p2p_create: 1 <= world <= %d, 0 <= rank < world, block_bytes a positive multiple of 16", KG_MAX_WORLD);
    CK(cudaSetDevice(ctx->device));
    tetra_p2p_destroy(ctx);
    const int64_t flag_off = 2 * (int64_t)world * block_bytes;
    const size_t total = (size_t)flag_off + 2 * KG_MAX_WORLD * sizeof(uint32_t);
    CK(ctx->p2p_buf.ensure(total));
    CK(ctx->p2p_misc.ensure(64));
    CK(cudaMemset(ctx->p2p_buf.p, 0, total));
    CK(cudaMemset(ctx->p2p_misc.p, 0, 64));
    ctx->p2p_rank = rank; ctx->p2p_world = world; ctx->p2p_block = block_bytes; ctx->p2p_flag_off = flag_off; ctx->p2p_step = 0;
    ctx->p2p_connected = false;
    for (int r = 0; r < KG_MAX_WORLD; ++r) { ctx->p2p_peer[r] = nullptr; ctx->p2p_ipc[r] = false; }
    ctx->p2p_peer[rank] = (uint8_t*)ctx->p2p_buf.p;
    if (world == 1) ctx->p2p_connected = true;
    if (handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) <= TETRA_IPC_HANDLE_BYTES, "IPC handle does not fit");
        memset(handle_out, 0, TETRA_IPC_HANDLE_BYTES);
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->p2p_buf.p));
        memcpy(handle_out, &h, sizeof h);
    }
    return TETRA_OK;
}

void* tetra_p2p_buffer(tetra_ctx* ctx) { return ctx && ctx->p2p_rank >= 0 ? ctx->p2p_buf.p : nullptr; }

int tetra_p2p_connect(tetra_ctx* ctx, const uint8_t* handles) {
    if (!ctx) return TETRA_E_INVALID;
    if (ctx->p2p_rank < 0 || !handles) return fail(ctx, TETRA_E_INVALID, "tetra_p2p_connect: call tetra_p2p_create first");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->p2p_world; ++r) {
        if (r == ctx->p2p_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * TETRA_IPC_HANDLE_BYTES, sizeof h);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p_peer[r] = (uint8_t*)ptr;
        ctx->p2p_ipc[r] = true;
    }
    ctx->p2p_connected = true;
    return TETRA_OK;
}

int tetra_p2p_connect_ptrs(tetra_ctx* ctx, void* const* buffers) {
    if (!ctx) return TETRA_E_INVALID;
    if (ctx->p2p_rank < 0 || !buffers) return fail(ctx, TETRA_E_INVALID, "tetra_p2p_connect_ptrs: call tetra_p2p_create first");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->p2p_world; ++r) {
        if (r == ctx->p2p_rank) continue;
        if (!buffers[r]) return fail(ctx, TETRA_E_INVALID, "tetra_p2p_connect_ptrs: buffer of rank %d is NULL", r);
        cudaPointerAttributes at;
        CK(cudaPointerGetAttributes(&at, buffers[r]));
        if (at.type != cudaMemoryTypeDevice) return fail(ctx, TETRA_E_INVALID, "tetra_p2p_connect_ptrs: rank %d's buffer is not device memory", r);
        if (at.device != ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else CK(e);
        }
        ctx->p2p_peer[r] = (uint8_t*)buffers[r];
    }
    ctx->p2p_connected = true;
    return TETRA_OK;
}

int tetra_allgather_dibits(tetra_ctx* ctx, const uint8_t* dibits, int64_t n, const int32_t* n_dibits, int32_t n_local,
                           uint8_t* all_dibits, int32_t* all_n) {
    if (!ctx) return TETRA_E_INVALID;
    if (!ctx->p2p_connected) return fail(ctx, TETRA_E_INVALID, "tetra_allgather_dibits: tetra_p2p_create / tetra_p2p_connect first");
    if (n <= 0 || (n & 15) || n_local <= 0 || !dibits || !n_dibits || !all_dibits ||
        (reinterpret_cast<uintptr_t>(dibits) & 15) || (reinterpret_cast<uintptr_t>(all_dibits) & 15))
        return fail(ctx, TETRA_E_INVALID, "tetra_allgather_dibits: n a positive multiple of 16, 16-byte aligned device buffers");
    if (n / 4 + 4 * (int64_t)n_local > ctx->p2p_block)
        return fail(ctx, TETRA_E_INVALID, "tetra_allgather_dibits: %lld bytes per rank exceed the block of %lld the exchange was created with",
                    (long long)(n / 4 + 4 * (int64_t)n_local), (long long)ctx->p2p_block);
    if (!is_device_ptr(dibits) || !is_device_ptr(n_dibits) || !is_device_ptr(all_dibits) || (all_n && !is_device_ptr(all_n)))
        return fail(ctx, TETRA_E_INVALID, "tetra_allgather_dibits: device buffers only");
    CK(cudaSetDevice(ctx->device));
    GatherArgs ga;
    memset(&ga, 0, sizeof ga);
    for (int r = 0; r < ctx->p2p_world; ++r) ga.recv[r] = ctx->p2p_peer[r];
    ga.rank = ctx->p2p_rank; ga.world = ctx->p2p_world; ga.block = ctx->p2p_block; ga.flag_off = ctx->p2p_flag_off;
    ga.step = ++ctx->p2p_step;
    ga.dibits = dibits; ga.n = n; ga.n_dibits = n_dibits; ga.n_local = n_local;
    ga.ticket = (uint32_t*)ctx->p2p_misc.p; ga.status = (int32_t*)ctx->p2p_misc.p + 4;
    ga.out = all_dibits; ga.out_n = all_n;
    const int64_t thr = std::max<int64_t>(n / 64, 1);
    const unsigned g_push = (unsigned)std::max<int64_t>(1, std::min<int64_t>((thr + KG_THREADS - 1) / KG_THREADS, 148));
    k_gather_push<<<g_push, KG_THREADS, 0, ctx->stream>>>(ga);
    const unsigned g_un = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n / 16 + KG_THREADS - 1) / KG_THREADS, (148 * 4 + ga.world - 1) / ga.world));
    k_gather_wait_unpack<<<dim3(g_un, (unsigned)ga.world), KG_THREADS, 0, ctx->stream>>>(ga);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return TETRA_OK;
}

int tetra_process_batch_allgather(tetra_ctx* ctx, const float* iq, int32_t C, int64_t N, int64_t pitch, const double* fo_hz,
                                  uint8_t* dibits, int64_t cap, int32_t* n_dibits, float* symbols, int32_t* best_phase,
                                  uint8_t* ts_match, uint8_t* all_dibits, int32_t* all_n) {
    if (!ctx) return TETRA_E_INVALID;
    if (!ctx->p2p_connected) return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_allgather: tetra_p2p_create / tetra_p2p_connect first");
    if (C <= 0 || cap <= 0 || (cap & 15) || !dibits || !n_dibits || !all_dibits ||
        (reinterpret_cast<uintptr_t>(dibits) & 15) || (reinterpret_cast<uintptr_t>(all_dibits) & 15))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_allgather: cap a multiple of 16, 16-byte aligned device buffers");
    const int64_t n = (int64_t)C * cap;
    if (n / 4 + 4 * (int64_t)C > ctx->p2p_block)
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_allgather: %lld bytes per rank exceed the block of %lld the exchange was created with",
                    (long long)(n / 4 + 4 * (int64_t)C), (long long)ctx->p2p_block);
    if (!is_device_ptr(dibits) || !is_device_ptr(n_dibits) || !is_device_ptr(all_dibits) || (all_n && !is_device_ptr(all_n)))
        return fail(ctx, TETRA_E_INVALID, "tetra_process_batch_allgather: device buffers only");
    // the finalize kernel pushes (fa.push); then every source rank's block is awaited and unpacked
    ++ctx->p2p_step;
    ctx->fg_active = true;
    const int rc = tetra_process_batch(ctx, iq, C, N, pitch, fo_hz, dibits, cap, n_dibits, symbols, best_phase, ts_match, 1);
    ctx->fg_active = false;
    if (rc != TETRA_OK) { --ctx->p2p_step; return rc; }
    GatherArgs ga;
    memset(&ga, 0, sizeof ga);
    for (int r = 0; r < ctx->p2p_world; ++r) ga.recv[r] = ctx->p2p_peer[r];
    ga.rank = ctx->p2p_rank; ga.world = ctx->p2p_world; ga.block = ctx->p2p_block; ga.flag_off = ctx->p2p_flag_off;
    ga.step = ctx->p2p_step;
    ga.n = n; ga.n_local = C;
    ga.status = (int32_t*)ctx->p2p_misc.p + 4;
    ga.out = all_dibits; ga.out_n = all_n;
    const unsigned g_un = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n / 16 + KG_THREADS - 1) / KG_THREADS, (148 * 4 + ga.world - 1) / ga.world));
    k_gather_wait_unpack<<<dim3(g_un, (unsigned)ga.world), KG_THREADS, 0, ctx->stream>>>(ga);
    ctx->launches++;
    CK(cudaGetLastError());
    return TETRA_OK;
}

int tetra_p2p_status(tetra_ctx* ctx, int32_t* status) {
    if (!ctx || !status) return TETRA_E_INVALID;
    if (ctx->p2p_rank < 0) return fail(ctx, TETRA_E_INVALID, "tetra_p2p_status: no exchange");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(status, (int32_t*)ctx->p2p_misc.p + 4, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TETRA_OK;
}

int tetra_p2p_destroy(tetra_ctx* ctx) {
    if (!ctx) return TETRA_E_INVALID;
    if (ctx->p2p_rank < 0) return TETRA_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < KG_MAX_WORLD; ++r) {
        if (ctx->p2p_ipc[r] && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
        ctx->p2p_peer[r] = nullptr; ctx->p2p_ipc[r] = false;
    }
    ctx->p2p_buf.release(); ctx->p2p_misc.release();
    ctx->p2p_rank = -1; ctx->p2p_world = 0; ctx->p2p_connected = false;
    return TETRA_OK;
}

int tetra_design_butter4(double wn, double* b5, double* a5) {
    if (!b5 || !a5) return TETRA_E_INVALID;
    return design_butter4(wn, b5, a5) ? TETRA_OK : TETRA_E_INVALID;
}
int tetra_design_cheby1_sos8(double rp, double wn, double* sos24) {
    if (!sos24) return TETRA_E_INVALID;
    return design_cheby1_sos8(rp, wn, sos24) ? TETRA_OK : TETRA_E_INVALID;
}

}  // extern "C"
