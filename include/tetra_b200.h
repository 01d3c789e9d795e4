/*
 * tetra_b200.h -- C ABI of libtetra_b200.so: the Blackwell (sm_100a) implementation of the
 * TetraEar IQ -> dibit demodulation hot path.
 *
 * The reference (syrex1013/TetraEar) has no FFI for this path: the boundary is the Python class
 * tetraear.signal.processor.SignalProcessor and the hand-off TetraDecoder.decode(symbols)
 * (SURVEY.md section 8b). Each entry point below states the reference interface it replaces
 * (file:line, relative to the reference tree). tetraear_b200/processor.py binds them with ctypes;
 * INTEGRATION.md shows the stub a TetraEar maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative TETRA_E_* code and never throws
 * or aborts across the ABI; the message is available from tetra_last_error(). Buffers are
 * caller-owned; the library keeps no reference to them after the call returns. Pointers marked
 * "host or device" may be either (detected with cudaPointerGetAttributes). A context is
 * single-caller (the reference uses one SignalProcessor per thread, ui/modern.py:1879); several
 * contexts may coexist.
 */
#ifndef TETRA_B200_H
#define TETRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tetra_ctx tetra_ctx;

#define TETRA_OK            0
#define TETRA_E_INVALID    -1   /* bad argument */
#define TETRA_E_CUDA       -2   /* CUDA runtime failure (message has the cudaError string) */
#define TETRA_E_NOMEM      -3
#define TETRA_E_UNSUPPORTED -4

/* Replaces SignalProcessor.__init__ (signal/processor.py:21-33). device = CUDA ordinal. */
int tetra_create(tetra_ctx** out, int device, double sample_rate);
void tetra_destroy(tetra_ctx* ctx);
/* Last error text of this context (ctx == NULL: last error of tetra_create on this thread). */
const char* tetra_last_error(const tetra_ctx* ctx);
/* SignalProcessor.sample_rate is mutated from outside at run time (ui/modern.py:1851-1852). */
int tetra_set_sample_rate(tetra_ctx* ctx, double sample_rate);
/* Work is enqueued on this CUDA stream (cudaStream_t as void*; NULL = the context's own
 * non-blocking stream; pass cudaStreamLegacy (0x1) for the legacy default stream). */
int tetra_set_stream(tetra_ctx* ctx, void* cuda_stream);
/* Block until everything enqueued by this context has finished. */
int tetra_synchronize(tetra_ctx* ctx);
/* Host batches (tetra_process_batch[_sync|_u8] with the IQ in host memory) larger than two chunks go through in
 * chunks of carriers: the host-to-device copy of the chunks ahead runs on a copy stream beside the kernels and the
 * result copies of the current chunk. bytes = chunk size (0: the default, 32 MiB; < 0: never chunk). The reference
 * has no counterpart: SignalProcessor.process takes one block at a time (processor.py:221). */
int tetra_set_h2d_chunk(tetra_ctx* ctx, int64_t bytes);

/* Upper bound on dibits produced for an N-sample block at the context's sample rate
 * (n_symbols - 1, signal/processor.py:213-215 and :135). */
int64_t tetra_dibit_capacity(const tetra_ctx* ctx, int64_t n_samples);

/* Number of soft symbols extract_symbols keeps for an N-sample block whose timing pick chose `best_phase`
 * ((len - best_phase) // sps, signal/processor.py:213-215; the whole block below 2 samples per symbol, :184-186). */
int64_t tetra_symbol_count(const tetra_ctx* ctx, int64_t n_samples, int32_t best_phase);

/*
 * Replaces SignalProcessor.process (signal/processor.py:221-273) for a batch of C independent
 * carriers, plus -- optionally -- the bit expansion and the training-sequence correlation of
 * TetraDecoder.symbols_to_bits / find_sync (core/decoder.py:140-169, 231-259).
 *
 *   iq            [C][pitch] complex64 as interleaved float (re, im); host or device
 *   n_samples     samples used per carrier (N); pitch >= N is the carrier stride in samples
 *   freq_offset_hz host [C] or NULL (= all zero); the `freq_offset` argument of process()
 *   dibits        [C][cap] uint8 values 0..3; host or device
 *   n_dibits      [C] int32; host or device
 *   symbols       [C][cap+1] complex64 (interleaved float) soft symbols = SignalProcessor.symbols
 *                 (processor.py:268); host or device; may be NULL
 *   best_phase    [C] int32 timing phase picked by extract_symbols (processor.py:189-210); may be NULL
 *   ts_match      [C][2*cap][2] uint8: number of bits (0..22) agreeing with TS1 / TS2 for the 22-bit
 *                 window starting at each bit position (decoder.py:237-240); may be NULL
 * With async != 0 the call only enqueues (device buffers required); use tetra_synchronize().
 */
int tetra_process_batch(tetra_ctx* ctx, const float* iq, int32_t n_carriers, int64_t n_samples,
                        int64_t pitch, const double* freq_offset_hz,
                        uint8_t* dibits, int64_t cap, int32_t* n_dibits,
                        float* symbols, int32_t* best_phase, uint8_t* ts_match, int32_t async);

/*
 * tetra_process_batch plus the frame-sync front end of TetraDecoder.decode (core/decoder.py:835-857) on the
 * device: bit expansion, find_sync at 0.90 / 0.85 / 0.80 / adaptive with the reference's visiting order
 * (jump +250 after a hit, max_corr over visited offsets only, in-function adaptive retry).
 *   sync_pos  [C][max_positions] int32 bit offsets, ascending; host or device
 *   n_sync    [C] int32 number of positions found (same residency as sync_pos)
 * max_positions must be at least 2*cap/250 + 2. Frame start = position - 216 bits (decoder.py:865).
 */
int tetra_process_batch_sync(tetra_ctx* ctx, const float* iq, int32_t n_carriers, int64_t n_samples,
                             int64_t pitch, const double* freq_offset_hz,
                             uint8_t* dibits, int64_t cap, int32_t* n_dibits,
                             float* symbols, int32_t* best_phase, uint8_t* ts_match,
                             int32_t* sync_pos, int32_t max_positions, int32_t* n_sync, int32_t async);
/*
 * Transport packing for the exchange of dibit streams between GPUs (the north star's all-gather before the host-side
 * TetraDecoder.decode, core/decoder.py:835): four dibits (values 0..3) per byte,
 *   packed[k] = d[4k] | d[4k+1] << 2 | d[4k+2] << 4 | d[4k+3] << 6.
 * Device buffers, asynchronous on the context's stream. pack: n dibits, n % 16 == 0, dibits 16-byte aligned.
 * unpack: n_blocks blocks of packed_bytes (% 4 == 0) at a stride of in_stride bytes -> 4 * packed_bytes dibits per
 * block at a stride of out_stride bytes (% 16 == 0): the layout an all-gather of [packed streams | lengths] leaves.
 */
int tetra_pack_dibits(tetra_ctx* ctx, const uint8_t* dibits, int64_t n, uint8_t* packed);
int tetra_unpack_dibits(tetra_ctx* ctx, const uint8_t* packed, int64_t n_blocks, int64_t packed_bytes, int64_t in_stride,
                        uint8_t* dibits, int64_t out_stride);

/*
 * The one exchange of the sharded path (SURVEY.md 8b proposed tetra_allgather_dibits; BASELINE north star: "a single
 * all-gather of decoded dibit streams over NVLink"; downstream of it is the reference's own TetraDecoder.decode(symbols),
 * tetraear/core/decoder.py:835, on every rank) -- without NCCL or torch: two kernels over NVLink peer memory. Every rank
 * (one process per GPU, or one context per GPU inside a process; world <= 8, one node) owns a receive buffer:
 *   tetra_p2p_create      allocates it for blocks of block_bytes per rank (>= n / 4 + 4 * n_local, a multiple of 16) and
 *                         writes its CUDA IPC handle (TETRA_IPC_HANDLE_BYTES) to handle_out (may be NULL);
 *   tetra_p2p_connect     takes the handles of all ranks, [world][TETRA_IPC_HANDLE_BYTES] in rank order (exchanged by the
 *                         caller by any means: torch.distributed, MPI, a file), and opens the peers' buffers;
 *   tetra_p2p_connect_ptrs  the same for contexts of ONE process: buffers[r] = tetra_p2p_buffer(ctx of rank r);
 *   tetra_allgather_dibits  asynchronous on the context's stream: packs dibits[n] (n = n_local * cap, cap % 16 == 0, values
 *                         0..3) four to a byte, stores the words and the n_local lengths into every peer's buffer, raises
 *                         this step's flag there, waits for the peers' flags and unpacks into all_dibits [world][n] and
 *                         all_n [world][n_local] (may be NULL). Device buffers only. Every rank must call it the same
 *                         number of times. A peer that does not show up within ~3 s is reported by tetra_p2p_status
 *                         (status = 1 + its rank; 0 = fine) instead of hanging the GPU.
 */
#define TETRA_IPC_HANDLE_BYTES 64
int tetra_p2p_create(tetra_ctx* ctx, int32_t rank, int32_t world, int64_t block_bytes, uint8_t* handle_out);
void* tetra_p2p_buffer(tetra_ctx* ctx);
int tetra_p2p_connect(tetra_ctx* ctx, const uint8_t* handles);
int tetra_p2p_connect_ptrs(tetra_ctx* ctx, void* const* buffers);
int tetra_allgather_dibits(tetra_ctx* ctx, const uint8_t* dibits, int64_t n, const int32_t* n_dibits, int32_t n_local,
                           uint8_t* all_dibits, int32_t* all_n);
/*
 * tetra_process_batch (asynchronous, device buffers, cap % 16 == 0) with the exchange fused behind the slicer: the CTA that
 * has sliced a carrier packs its dibits and stores them, with the stream length, straight into every peer's receive buffer;
 * the last CTA of the batch raises the step's flag; then the peers' blocks are awaited and unpacked as above. One compute
 * kernel does slicing and push -- no separate pack / all-gather pass over the streams.
 */
int tetra_process_batch_allgather(tetra_ctx* ctx, const float* iq, int32_t n_carriers, int64_t n_samples, int64_t pitch,
                                  const double* freq_offset_hz, uint8_t* dibits, int64_t cap, int32_t* n_dibits,
                                  float* symbols, int32_t* best_phase, uint8_t* ts_match,
                                  uint8_t* all_dibits, int32_t* all_n);
int tetra_p2p_status(tetra_ctx* ctx, int32_t* status);
int tetra_p2p_destroy(tetra_ctx* ctx);

/* The same cascade for dibit streams that are already there ([C][cap] uint8 + lengths; host or device). */
int tetra_sync_positions(tetra_ctx* ctx, const uint8_t* dibits, int64_t cap, const int32_t* n_dibits,
                         int32_t n_carriers, int32_t* sync_pos, int32_t max_positions, int32_t* n_sync);

/*
 * What TetraDecoder.decode does with each sync position up to the burst's CRC verdict (core/decoder.py:861-888,
 * decode_frame :986-992, TetraProtocolParser.parse_burst core/protocol.py:192-347), for every (carrier, position):
 *   burst_info [C][max_positions][4] int32 = (start_symbol or -1 when decode() drops the position,
 *                                             frame_number = start_bit / 510, burst_type (2 normal downlink,
 *                                             5 synchronization), crc_ok)
 * All buffers host or all device.
 */
int tetra_parse_bursts(tetra_ctx* ctx, const uint8_t* dibits, int64_t cap, const int32_t* n_dibits,
                       int32_t n_carriers, const int32_t* sync_pos, int32_t max_positions,
                       const int32_t* n_sync, int32_t* burst_info);

/*
 * The per-sample part of TetraSignalDetector.analyze_signal (signal/scanner.py:42-147, 204-231) for C captures
 * (e.g. the channels of a wideband survey): calculate_power, detect_tetra_modulation, detect_sync_pattern and
 * check_power_stability in float64.  iq [C][pitch] complex64 host or device;
 *   out6 [C][6] float64 = (power_db, modulation_confidence, sync_correlation, power_stable (0/1),
 *                          modulation_matches, number of phase differences); host or device.
 * The decision thresholds stay with the caller (confidence > 0.4, correlation > 0.75, scanner.py:94, 145).
 */
int tetra_analyze_signal(tetra_ctx* ctx, const float* iq, int32_t n_captures, int64_t n_samples, int64_t pitch,
                         double* out6);
/*
 * The scanner's sweep in one capture (signal/scanner.py:383-445 retunes the hardware in 25 kHz steps, 0.3 s each): for every
 * channel offset f_c of ONE wideband capture, what the reference computes after tuning there -- the six numbers above on
 * frequency_shift(iq, f_c) (processor.py:85-100; the shift is applied as a phase step, the shifted capture is never formed)
 * and the signal-presence / AFC block of CaptureThread.run (ui/modern.py:1945-2012) on the float64 spectrum of its first
 * nfft samples (the reference uses 2048):
 *   out[c][6..11] = signal_power (mean dB of the centre 25 kHz), peak_power, peak_freq_offset_hz (the AFC offset handed to
 *                   process()), noise_floor, snr, is_signal_strong (snr > 15 and peak > -70 and peak - mean > 3).
 * out is host memory, [n_channels][TETRA_SURVEY_FIELDS] doubles (fields 12.. are zero); iq complex64 host or device.
 * Per-channel demodulation of the same capture is tetra_process_wideband.
 */
#define TETRA_SURVEY_FIELDS 16
int tetra_survey_wideband(tetra_ctx* ctx, const float* iq, int64_t n_samples, const double* channel_hz, int32_t n_channels,
                          int32_t nfft, double* out);

/*
 * tetra_process_batch_sync for RTL-SDR native samples: iq_u8 [C][pitch][2] interleaved unsigned 8-bit I, Q
 * (host or device), converted on the device exactly like pyrtlsdr's packed_bytes_to_iq does before the
 * reference sees them (RTLCapture.read_samples, signal/capture.py:143-158): (byte / 127.5) - 1.
 * A quarter of the host-to-device bytes of the complex64 entry point. At 2.4 MS/s with no freq_offset the fused
 * kernel reads the bytes itself (2 bytes per sample of HBM traffic; rows that start on 16-byte boundaries, i.e.
 * pitch % 8 == 0, go through bulk copies); otherwise they are expanded to complex64 first. Synchronous.
 */
int tetra_process_batch_u8(tetra_ctx* ctx, const uint8_t* iq_u8, int32_t n_carriers, int64_t n_samples,
                           int64_t pitch, const double* freq_offset_hz,
                           uint8_t* dibits, int64_t cap, int32_t* n_dibits,
                           float* symbols, int32_t* best_phase, uint8_t* ts_match,
                           int32_t* sync_pos, int32_t max_positions, int32_t* n_sync);

/*
 * BASELINE config 3 -- C channels out of ONE wideband capture: for every channel centre f_c (Hz, relative
 * to the capture centre) the composition  process(frequency_shift(iq, f_c, sample_rate), 0)  of the
 * reference's own methods (signal/processor.py:85-100 and :221-273). The reference has no channelizer; its
 * scanner retunes the hardware in 25 kHz steps instead (signal/scanner.py:383-445).
 *   iq [n_samples] complex64, host or device; channel_hz host [C]; outputs as in tetra_process_batch.
 */
int tetra_process_wideband(tetra_ctx* ctx, const float* iq, int64_t n_samples, const double* channel_hz,
                           int32_t n_channels, uint8_t* dibits, int64_t cap, int32_t* n_dibits,
                           float* symbols, int32_t* best_phase, uint8_t* ts_match);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t tetra_launch_count(const tetra_ctx* ctx);
/* With timing enabled every fused channelize+demod kernel launch is bracketed by a CUDA-event pair
 * on the stream it is launched on. tetra_kernel_time_ms waits for them, returns the summed device
 * time in ms of the launches recorded since the last query (< 0 if none), stores their number in
 * *n_launches and resets the record. */
int tetra_enable_kernel_timing(tetra_ctx* ctx, int on);
double tetra_kernel_time_ms(tetra_ctx* ctx, int32_t* n_launches);
/* Timeline of the most recent timed fast-path call, device ms between consecutive marks on the
 * context's stream: out3[0] call start -> fused kernel done and edge windows joined,
 * out3[1] span of the edge-window kernel on the side stream, out3[2] timing pick + slicer + TS correlator. */
int tetra_last_phase_ms(tetra_ctx* ctx, double* out3);

/*
 * Host replay of TetraDecoder.find_sync (core/decoder.py:226-295) over device-computed match
 * counts for ONE carrier: same visiting order (jump +250 after a hit), same max_corr bookkeeping
 * over visited positions only, same in-function adaptive retry.
 *   match [n_windows][2] uint8, threshold as passed by the caller (0.90/0.85/0.80/adaptive)
 *   positions out [max_positions]; returns the count (>= 0) or a negative error; *max_corr out.
 */
int tetra_find_sync(const uint8_t* match, int64_t n_windows, double threshold,
                    int32_t* positions, int32_t max_positions, double* max_corr);
/* decoder.py:845-856: the 0.90 / 0.85 / 0.80 / adaptive cascade of TetraDecoder.decode. */
int tetra_sync_cascade(const uint8_t* match, int64_t n_windows,
                       int32_t* positions, int32_t max_positions);

/*
 * Device versions of the reference's public helper methods, complex128 in / complex128 out as in
 * the reference (host pointers, interleaved double re, im).
 */
/* filter_signal (processor.py:51-83): butter(4) zero-phase low-pass. Returns 1 if the filter was
 * skipped the way the reference skips it (too short / design failure -> input copied). */
int tetra_filter_signal(tetra_ctx* ctx, const double* in, int64_t n, double bandwidth,
                        double sample_rate, double* out);
/* frequency_shift (processor.py:85-100). */
int tetra_frequency_shift(tetra_ctx* ctx, const double* in, int64_t n, double freq_offset,
                          double sample_rate, double* out);
/* extract_symbols (processor.py:168-219): out has room for n elements. */
int tetra_extract_symbols(tetra_ctx* ctx, const double* in, int64_t n, double sample_rate,
                          double* out, int64_t* n_out, int32_t* best_phase);
/* demodulate_dqpsk (processor.py:102-166): out has room for n-1 dibits. */
int tetra_demodulate_dqpsk(tetra_ctx* ctx, const double* in, int64_t n, uint8_t* out, int64_t* n_out);
/* resample (processor.py:35-49, scipy.signal.resample FFT method) -- Fourier resampling on device. */
int tetra_resample(tetra_ctx* ctx, const double* in, int64_t n, int64_t n_out, double* out);

/*
 * Waterfall rows (ui/modern.py:1921-1934 applied at every hop): Hann window, FFT, fftshift,
 * 20*log10(|X|/nfft + 1e-20).  iq complex64 host or device, out [rows][nfft] float host or device.
 * nfft must be a power of two in [64, 8192].
 */
int tetra_stft_db(tetra_ctx* ctx, const float* iq, int64_t n_samples, int32_t nfft, int32_t hop,
                  float* out, int64_t* rows);
/*
 * The same in float64 on complex128 host input, float64 host rows out: the precision of the reference's own numpy FFT.
 * SignalProcessor.spectrum() -- the once-per-chunk spectrum block, ui/modern.py:1919-1943 -- runs this one (a 2048-point
 * row costs microseconds either way); the waterfall at throughput stays float32. nfft a power of two in [64, 4096].
 */
int tetra_stft_db_f64(tetra_ctx* ctx, const double* iq, int64_t n_samples, int32_t nfft, int32_t hop,
                      double* out, int64_t* rows);

/*
 * Block-end corrections of the fused 2.4 MS/s path, exposed for testing: what SciPy's sosfiltfilt / filtfilt edge
 * handling (odd extension, zi * x0, inside scipy.signal.decimate and filter_signal, processor.py:254 and :79) adds to
 * the shift-invariant response of the zero-extended block, for the 168 outputs next to each end of every carrier.
 *   iq host [C][pitch] complex64, freq_offset_hz host [C] or NULL, out host [C][2][168] complex64 (left end m = 0..,
 *   right end output L-1-t, t = 0..). n_samples >= 16384.
 */
int tetra_edge_corrections(tetra_ctx* ctx, const float* iq, int32_t n_carriers, int64_t n_samples, int64_t pitch,
                           const double* freq_offset_hz, float* out);

/* Filter design used by the generic path (what scipy.signal.butter / cheby1 return to the
 * reference at processor.py:78 and inside scipy.signal.decimate). Exposed for testing. */
int tetra_design_butter4(double wn, double* b5, double* a5);
int tetra_design_cheby1_sos8(double rp_db, double wn, double* sos24);

#ifdef __cplusplus
}
#endif
#endif
