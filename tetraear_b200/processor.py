"""Host-side mirror of the reference's ``tetraear.signal.processor.SignalProcessor``.

Same class name, constructor, attributes and method surface as the reference
(tetraear/signal/processor.py:18-273), so that ``TetraDecoder`` and the capture loops
(ui/modern.py:1879, 2022; continuous_capture.py:50-56) run unchanged. Every method calls the
sm_100a kernels through the C ABI of ``libtetra_b200.so`` (include/tetra_b200.h) with ctypes;
there is no NumPy/SciPy implementation behind it and no CPU fallback: without the library or a
CUDA device the constructor raises.

Error behaviour follows the reference: empty input gives empty output (processor.py:239-241,
120-121, 179-180, 66-67); a failure inside the DSP is logged on the reference's logger name and
``process`` returns empty arrays instead of raising (processor.py:81-83, 256-257).
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib

logger = logging.getLogger("tetraear.signal.processor")


def _as_c64(samples) -> np.ndarray:
    a = np.asarray(samples)
    if a.dtype != np.complex64:
        a = a.astype(np.complex64)
    return np.ascontiguousarray(a)


def _as_c128(samples) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(samples), dtype=np.complex128)


class SignalProcessor:
    """Processes raw IQ samples for TETRA demodulation on a B200 (drop-in for the reference class)."""

    def __init__(self, sample_rate=2.4e6, device: int = 0):
        self._lib = _lib.load()
        self._ctx = _lib.c_ctx_p()
        rc = self._lib.tetra_create(C.byref(self._ctx), int(device), float(sample_rate))
        if rc != 0:
            msg = self._lib.tetra_last_error(None)
            raise _lib.TetraError(f"tetra_create failed ({rc}): {msg.decode() if msg else ''}")
        self.sample_rate = sample_rate
        self.symbol_rate = 18000                                  # processor.py:30
        self.samples_per_symbol = int(sample_rate / self.symbol_rate)   # processor.py:31
        self.symbols = None                                       # processor.py:33
        self.best_phase = None
        self.device = int(device)

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.tetra_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        return _lib.check(self._lib, self._ctx, rc, what)

    def _sync_rate(self):
        # sample_rate is a plain attribute mutated from outside (ui/modern.py:1851-1852)
        self._check(self._lib.tetra_set_sample_rate(self._ctx, float(self.sample_rate)), "set_sample_rate")

    # -- reference helper methods -----------------------------------------------------------
    def resample(self, samples, target_rate):
        """processor.py:35-49 (scipy.signal.resample, FFT method) -- Fourier resampling on device."""
        x = _as_c128(samples)
        num = int(len(x) * target_rate / self.sample_rate)
        if len(x) == 0 or num <= 0:
            return np.zeros(max(num, 0), dtype=np.complex128)
        out = np.empty(num, dtype=np.complex128)
        self._check(self._lib.tetra_resample(self._ctx, x.ctypes.data, len(x), num, out.ctypes.data), "resample")
        return out

    def filter_signal(self, samples, bandwidth=25000, sample_rate=None):
        """processor.py:51-83: Butterworth order-4 low-pass, zero phase."""
        if len(samples) == 0:
            return samples
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x = _as_c128(samples)
        out = np.empty_like(x)
        try:
            self._check(self._lib.tetra_filter_signal(self._ctx, x.ctypes.data, len(x), float(bandwidth), float(fs),
                                                      out.ctypes.data), "filter_signal")
        except _lib.TetraError as e:
            logger.warning(f"Filter design failed, using unfiltered samples: {e}")
            return samples
        return out

    def frequency_shift(self, samples, freq_offset, sample_rate=None):
        """processor.py:85-100: multiply by exp(-j 2 pi f n / fs)."""
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x = _as_c128(samples)
        out = np.empty_like(x)
        if len(x):
            self._check(self._lib.tetra_frequency_shift(self._ctx, x.ctypes.data, len(x), float(freq_offset),
                                                        float(fs), out.ctypes.data), "frequency_shift")
        return out

    def demodulate_dqpsk(self, samples):
        """processor.py:102-166: differential phase slicer -> uint8 in 0..3."""
        if len(samples) < 2:
            return np.array([], dtype=np.uint8)
        x = _as_c128(samples)
        out = np.empty(len(x) - 1, dtype=np.uint8)
        n_out = C.c_int64(0)
        self._check(self._lib.tetra_demodulate_dqpsk(self._ctx, x.ctypes.data, len(x), out.ctypes.data,
                                                     C.byref(n_out)), "demodulate_dqpsk")
        return out[: n_out.value]

    def extract_symbols(self, samples, sample_rate=None):
        """processor.py:168-219: block timing pick + symbol-rate gather."""
        if len(samples) == 0:
            return np.array([], dtype=complex)
        fs = sample_rate if sample_rate is not None else self.sample_rate
        x = _as_c128(samples)
        out = np.empty(len(x), dtype=np.complex128)
        n_out, ph = C.c_int64(0), C.c_int32(0)
        self._check(self._lib.tetra_extract_symbols(self._ctx, x.ctypes.data, len(x), float(fs), out.ctypes.data,
                                                    C.byref(n_out), C.byref(ph)), "extract_symbols")
        self.best_phase = int(ph.value)
        return out[: n_out.value].copy()

    # -- the hot path -----------------------------------------------------------------------
    def process(self, samples, freq_offset=0):
        """processor.py:221-273 -> uint8 dibits; side effect ``self.symbols`` (processor.py:268)."""
        if len(samples) == 0:
            self.symbols = np.array([], dtype=complex)
            return np.array([], dtype=np.uint8)
        try:
            res = self.process_batch(_as_c64(samples)[None, :], [float(freq_offset)], want_symbols=True)
        except _lib.TetraError as e:
            logger.warning(f"GPU demodulation failed: {e}")
            self.symbols = np.array([], dtype=complex)
            return np.array([], dtype=np.uint8)
        n = int(res["n_dibits"][0])
        n_sym = int(res["n_symbols"][0])
        self.symbols = res["symbols"][0, :n_sym].astype(np.complex128)
        self.best_phase = int(res["best_phase"][0])
        return res["dibits"][0, :n].copy()

    def _n_symbols(self, n, nd, ph):
        """soft symbols per carrier: n_dibits + 1, and for blocks too short for a dibit what extract_symbols keeps
        (one symbol -> no dibit, but ``.symbols`` still holds it: processor.py:213-215, 268)"""
        out = np.where(nd > 0, nd + 1, 0).astype(np.int32)
        for c in np.nonzero(nd <= 0)[0]:
            out[c] = int(self._lib.tetra_symbol_count(self._ctx, int(n), int(ph[c])))
        return out

    def process_batch(self, iq, freq_offsets=None, want_symbols=True, want_match=False, want_sync=False):
        """Batched ``process``: iq complex64 [C, N] (numpy), freq_offsets [C] or None.

        Returns dict(dibits uint8 [C, cap], n_dibits int32 [C], n_symbols int32 [C], best_phase int32 [C],
        symbols complex64 [C, cap+1] (if want_symbols), ts_match uint8 [C, 2*cap, 2] (if want_match),
        sync_pos int32 [C, max_pos] + n_sync int32 [C] (if want_sync: the positions TetraDecoder.decode's
        threshold cascade finds, core/decoder.py:845-856, computed on the device)).
        """
        self._sync_rate()
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim != 2:
            raise ValueError("iq must be [carriers, samples]")
        n_car, n = iq.shape
        cap = int(self._lib.tetra_dibit_capacity(self._ctx, n))
        dib = np.zeros((n_car, max(cap, 1)), dtype=np.uint8)
        nd = np.zeros(n_car, dtype=np.int32)
        ph = np.zeros(n_car, dtype=np.int32)
        sym = np.zeros((n_car, cap + 1), dtype=np.complex64) if want_symbols else None
        mt = np.zeros((n_car, 2 * max(cap, 1), 2), dtype=np.uint8) if want_match else None
        fo = None
        if freq_offsets is not None:
            fo = np.ascontiguousarray(freq_offsets, dtype=np.float64)
            if fo.shape != (n_car,):
                raise ValueError("freq_offsets must have one entry per carrier")
        max_pos = (2 * cap) // 250 + 2
        spos = np.zeros((n_car, max_pos), dtype=np.int32) if want_sync else None
        nsync = np.zeros(n_car, dtype=np.int32) if want_sync else None
        self._check(self._lib.tetra_process_batch_sync(
            self._ctx, iq.ctypes.data, n_car, n, n, fo.ctypes.data if fo is not None else None,
            dib.ctypes.data, cap, nd.ctypes.data, sym.ctypes.data if sym is not None else None, ph.ctypes.data,
            mt.ctypes.data if mt is not None else None,
            spos.ctypes.data if want_sync else None, max_pos if want_sync else 0, nsync.ctypes.data if want_sync else None,
            0), "process_batch")
        out = dict(dibits=dib[:, :cap], n_dibits=nd, n_symbols=self._n_symbols(n, nd, ph), best_phase=ph)
        if want_sync:
            out["sync_pos"] = spos
            out["n_sync"] = nsync
        if sym is not None:
            out["symbols"] = sym
        if mt is not None:
            out["ts_match"] = mt[:, : 2 * cap]
        return out

    def process_batch_u8(self, iq_u8, freq_offsets=None, want_symbols=True, want_match=False, want_sync=False):
        """``process_batch`` for raw RTL-SDR bytes: iq_u8 uint8 [C, N, 2] (interleaved I, Q as the dongle delivers
        them). The conversion pyrtlsdr applies in ``read_samples`` -- (byte / 127.5) - 1 -- runs on the device, so
        only 2 bytes per sample cross PCIe."""
        self._sync_rate()
        raw = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        if raw.ndim != 3 or raw.shape[2] != 2:
            raise ValueError("iq_u8 must be [carriers, samples, 2]")
        n_car, n = raw.shape[0], raw.shape[1]
        cap = int(self._lib.tetra_dibit_capacity(self._ctx, n))
        dib = np.zeros((n_car, max(cap, 1)), dtype=np.uint8)
        nd = np.zeros(n_car, dtype=np.int32)
        ph = np.zeros(n_car, dtype=np.int32)
        sym = np.zeros((n_car, cap + 1), dtype=np.complex64) if want_symbols else None
        mt = np.zeros((n_car, 2 * max(cap, 1), 2), dtype=np.uint8) if want_match else None
        max_pos = (2 * cap) // 250 + 2
        spos = np.zeros((n_car, max_pos), dtype=np.int32) if want_sync else None
        nsync = np.zeros(n_car, dtype=np.int32) if want_sync else None
        fo = None
        if freq_offsets is not None:
            fo = np.ascontiguousarray(freq_offsets, dtype=np.float64)
        self._check(self._lib.tetra_process_batch_u8(
            self._ctx, raw.ctypes.data, n_car, n, n, fo.ctypes.data if fo is not None else None,
            dib.ctypes.data, cap, nd.ctypes.data, sym.ctypes.data if sym is not None else None, ph.ctypes.data,
            mt.ctypes.data if mt is not None else None,
            spos.ctypes.data if want_sync else None, max_pos if want_sync else 0, nsync.ctypes.data if want_sync else None),
            "process_batch_u8")
        out = dict(dibits=dib[:, :cap], n_dibits=nd, n_symbols=self._n_symbols(n, nd, ph), best_phase=ph)
        if sym is not None:
            out["symbols"] = sym
        if mt is not None:
            out["ts_match"] = mt[:, : 2 * cap]
        if want_sync:
            out["sync_pos"] = spos
            out["n_sync"] = nsync
        return out

    def process_wideband(self, samples, channel_freqs, want_symbols=True, want_match=False):
        """BASELINE config 3: every channel of one wideband capture, i.e. for each centre f_c the composition
        ``process(frequency_shift(samples, f_c), 0)`` of the reference's methods (processor.py:85-100, 221-273).
        samples complex [N]; channel_freqs [C] in Hz relative to the capture centre. Returns the dict of
        ``process_batch`` with one row per channel."""
        self._sync_rate()
        x = _as_c64(samples)
        fr = np.ascontiguousarray(channel_freqs, dtype=np.float64)
        n_ch, n = len(fr), len(x)
        cap = int(self._lib.tetra_dibit_capacity(self._ctx, n))
        dib = np.zeros((n_ch, max(cap, 1)), dtype=np.uint8)
        nd = np.zeros(n_ch, dtype=np.int32)
        ph = np.zeros(n_ch, dtype=np.int32)
        sym = np.zeros((n_ch, cap + 1), dtype=np.complex64) if want_symbols else None
        mt = np.zeros((n_ch, 2 * max(cap, 1), 2), dtype=np.uint8) if want_match else None
        self._check(self._lib.tetra_process_wideband(
            self._ctx, x.ctypes.data, n, fr.ctypes.data, n_ch, dib.ctypes.data, cap, nd.ctypes.data,
            sym.ctypes.data if sym is not None else None, ph.ctypes.data,
            mt.ctypes.data if mt is not None else None), "process_wideband")
        out = dict(dibits=dib[:, :cap], n_dibits=nd, n_symbols=self._n_symbols(n, nd, ph), best_phase=ph)
        if sym is not None:
            out["symbols"] = sym
        if mt is not None:
            out["ts_match"] = mt[:, : 2 * cap]
        return out

    # -- device-resident batch (bench / multi-GPU): pointers are raw CUDA addresses ----------
    def set_stream(self, stream=None):
        """Work of this context is enqueued on `stream` (a cudaStream_t as int; 0 = the legacy default stream,
        None = the context's own non-blocking stream)."""
        if stream is not None and int(stream) == 0:
            stream = 1                                    # cudaStreamLegacy
        self._check(self._lib.tetra_set_stream(self._ctx, stream), "set_stream")

    def set_h2d_chunk(self, n_bytes: int = 0):
        """Host batches larger than two chunks go through in chunks of carriers, the host-to-device copy of the chunks ahead
        beside the kernels of the current one (``tetra_set_h2d_chunk``): chunk size in bytes, 0 = the default (32 MiB),
        negative = host batches go through in one piece."""
        self._check(self._lib.tetra_set_h2d_chunk(self._ctx, int(n_bytes)), "set_h2d_chunk")

    def process_batch_device(self, iq_ptr: int, n_carriers: int, n_samples: int, pitch: int, dibits_ptr: int,
                             cap: int, n_dibits_ptr: int, symbols_ptr: int = 0, best_phase_ptr: int = 0,
                             ts_match_ptr: int = 0, stream=None, freq_offsets=None):
        """Enqueue on `stream` (a cudaStream_t as int; 0 = the legacy default stream, None = the
        context's own stream) with all buffers already in HBM; asynchronous."""
        self._sync_rate()
        self.set_stream(stream)
        fo = None
        if freq_offsets is not None:
            fo = np.ascontiguousarray(freq_offsets, dtype=np.float64)
        self._check(self._lib.tetra_process_batch(
            self._ctx, iq_ptr, n_carriers, n_samples, pitch, fo.ctypes.data if fo is not None else None,
            dibits_ptr, cap, n_dibits_ptr, symbols_ptr or None, best_phase_ptr or None, ts_match_ptr or None, 1),
            "process_batch(device)")

    def sync_positions(self, dibits, n_dibits=None):
        """decode()'s sync search (core/decoder.py:840-856) on the device for dibit streams [C, cap] (or one stream):
        returns a list of position lists."""
        d = np.ascontiguousarray(dibits, dtype=np.uint8)
        one = d.ndim == 1
        if one:
            d = d[None, :]
        n_car, cap = d.shape
        nd = np.full(n_car, cap, dtype=np.int32) if n_dibits is None else np.ascontiguousarray(n_dibits, dtype=np.int32)
        max_pos = (2 * cap) // 250 + 2
        spos = np.zeros((n_car, max_pos), dtype=np.int32)
        nsync = np.zeros(n_car, dtype=np.int32)
        if cap > 0:
            self._check(self._lib.tetra_sync_positions(self._ctx, d.ctypes.data, cap, nd.ctypes.data, n_car,
                                                       spos.ctypes.data, max_pos, nsync.ctypes.data), "sync_positions")
        res = [[int(p) for p in spos[c, : nsync[c]]] for c in range(n_car)]
        return res[0] if one else res

    def parse_bursts(self, dibits, n_dibits, sync_pos, n_sync):
        """For every (carrier, sync position): (start_symbol, frame_number, burst_type, crc_ok) of its 255-symbol
        slot as TetraDecoder.decode / parse_burst determine them (core/decoder.py:861-888, core/protocol.py:192-347);
        positions decode() drops are left out. Returns one list of tuples per carrier."""
        d = np.ascontiguousarray(dibits, dtype=np.uint8)
        nd = np.ascontiguousarray(n_dibits, dtype=np.int32)
        sp_ = np.ascontiguousarray(sync_pos, dtype=np.int32)
        ns = np.ascontiguousarray(n_sync, dtype=np.int32)
        n_car, cap = d.shape
        max_pos = sp_.shape[1]
        info = np.zeros((n_car, max_pos, 4), dtype=np.int32)
        self._check(self._lib.tetra_parse_bursts(self._ctx, d.ctypes.data, cap, nd.ctypes.data, n_car, sp_.ctypes.data, max_pos,
                                                 ns.ctypes.data, info.ctypes.data), "parse_bursts")
        return [[tuple(int(v) for v in info[c, k]) for k in range(ns[c]) if info[c, k, 0] >= 0] for c in range(n_car)]

    def analyze_signal(self, captures):
        """The per-sample part of ``TetraSignalDetector.analyze_signal`` (signal/scanner.py:42-147, 204-231) for one
        capture [N] or several [C, N]: a dict (or list of dicts) with power_db, modulation_confidence,
        is_tetra_modulation (> 0.4), sync_correlation, sync_detected (> 0.75) and power_stable."""
        self._sync_rate()
        x = np.ascontiguousarray(captures, dtype=np.complex64)
        one = x.ndim == 1
        if one:
            x = x[None, :]
        n_cap, n = x.shape
        out = np.zeros((n_cap, 6), dtype=np.float64)
        self._check(self._lib.tetra_analyze_signal(self._ctx, x.ctypes.data, n_cap, n, n, out.ctypes.data), "analyze_signal")
        res = [dict(power_db=float(r[0]), modulation_confidence=float(r[1]), is_tetra_modulation=bool(r[1] > 0.4),
                    sync_correlation=float(r[2]), sync_detected=bool(r[2] > 0.75), power_stable=bool(r[3] > 0.5),
                    modulation_matches=int(r[4])) for r in out]
        return res[0] if one else res

    def survey_wideband(self, capture, channel_freqs, center_frequency=0.0, n_fft=2048, min_power=-70, min_confidence=0.4,
                        bottom_threshold=-85):
        """``FrequencyScanner.scan_range`` (signal/scanner.py:383-445) from ONE wideband capture instead of a hardware retune
        per 25 kHz step: for every channel offset (Hz from the capture's centre) the dict ``TetraSignalDetector.analyze_signal``
        builds (scanner.py:233-289) on the capture shifted to that channel, the presence / AFC numbers of
        ``CaptureThread.run`` (ui/modern.py:1945-2012; ``peak_freq_offset`` is the AFC offset it hands to ``process``), and
        ``frequency`` / ``frequency_mhz`` as ``scan_frequency`` adds them (:367-368). Returns ``(results, found)``: one
        dict per channel and the ones passing ``scan_range``'s filter (:418-425). Frame validation (:149-202) needs the
        reference's ``TetraDecoder`` on the host; feed it ``process_wideband``'s dibits where a channel is worth it."""
        self._sync_rate()
        x = _as_c64(capture)
        f = np.ascontiguousarray(channel_freqs, dtype=np.float64)
        out = np.zeros((len(f), _lib.SURVEY_FIELDS), dtype=np.float64)
        if len(f):
            self._check(self._lib.tetra_survey_wideband(self._ctx, x.ctypes.data, len(x), f.ctypes.data, len(f), int(n_fft),
                                                        out.ctypes.data), "survey_wideband")
        results, found = [], []
        for k, r in enumerate(out):
            mod_conf, sync_corr = float(r[1]), float(r[2])
            is_mod, has_sync, stable = mod_conf > 0.4, sync_corr > 0.75, bool(r[3] > 0.5)
            if has_sync and is_mod:                                  # scanner.py:258-269 (crc_rate = 0: no frame validation here)
                confidence = mod_conf * 0.4 + sync_corr * 0.4
            elif has_sync:
                confidence = sync_corr * 0.6
            elif is_mod:
                confidence = mod_conf * 0.5
            else:
                confidence = 0.0
            d = dict(power_db=float(r[0]), is_tetra=bool(is_mod and has_sync and stable), confidence=confidence,
                     modulation_confidence=mod_conf, sync_detected=bool(has_sync), sync_correlation=sync_corr,
                     frames_validated=False, crc_pass_rate=0.0, power_stable=stable,
                     signal_present=bool(r[0] > bottom_threshold),
                     frequency=float(center_frequency + f[k]), frequency_mhz=float(center_frequency + f[k]) / 1e6,
                     signal_power=float(r[6]), peak_power=float(r[7]), peak_freq_offset=float(r[8]), noise_floor=float(r[9]),
                     snr=float(r[10]), is_signal_strong=bool(r[11] > 0.5))
            results.append(d)
            if d["is_tetra"] and d["power_db"] > min_power and d["confidence"] > min_confidence and d["sync_detected"] and stable:
                found.append(d)
        return results, found

    def dibit_capacity(self, n_samples: int) -> int:
        self._sync_rate()
        return int(self._lib.tetra_dibit_capacity(self._ctx, int(n_samples)))

    def stft_db(self, iq, n_fft=4096, hop=1024):
        """Waterfall rows: ui/modern.py:1921-1934 applied at every hop -> float32 [rows, n_fft]."""
        x = _as_c64(iq)
        rows = C.c_int64(0)
        n_rows = (len(x) - n_fft) // hop + 1 if len(x) >= n_fft else 0
        out = np.empty((max(n_rows, 0), n_fft), dtype=np.float32)
        self._check(self._lib.tetra_stft_db(self._ctx, x.ctypes.data, len(x), int(n_fft), int(hop),
                                            out.ctypes.data, C.byref(rows)), "stft_db")
        return out[: rows.value]

    def stft_db_f64(self, iq, n_fft=2048, hop=None):
        """The same rows in the reference's own precision: complex128 in, float64 FFT on the device, float64 rows out
        (n_fft <= 4096). What ``spectrum`` uses."""
        x = np.ascontiguousarray(iq, dtype=np.complex128)
        hop = int(hop or n_fft)
        rows = C.c_int64(0)
        n_rows = (len(x) - n_fft) // hop + 1 if len(x) >= n_fft else 0
        out = np.empty((max(n_rows, 0), n_fft), dtype=np.float64)
        self._check(self._lib.tetra_stft_db_f64(self._ctx, x.ctypes.data, len(x), int(n_fft), hop,
                                                out.ctypes.data, C.byref(rows)), "stft_db_f64")
        return out[: rows.value]

    def spectrum(self, samples, n_fft=2048, center_frequency=0.0):
        """The spectrum block of CaptureThread.run (ui/modern.py:1921-1937): (freqs, power_dB), float64 like the reference."""
        p = self.stft_db_f64(np.asarray(samples)[:n_fft], n_fft, n_fft)
        freqs = np.fft.fftshift(np.fft.fftfreq(n_fft, 1 / self.sample_rate)) + center_frequency
        return freqs, p[0]

    def enable_kernel_timing(self, on=True):
        self._lib.tetra_enable_kernel_timing(self._ctx, 1 if on else 0)

    def kernel_time_ms(self):
        """(summed device ms, launches) of the fused kernel since the last query (CUDA events on its stream)."""
        n = C.c_int32(0)
        ms = float(self._lib.tetra_kernel_time_ms(self._ctx, C.byref(n)))
        return ms, int(n.value)

    def last_phase_ms(self):
        """Device ms of the last timed fast-path call: [fused kernel + edge join, gap, finalize]."""
        out = np.zeros(3, dtype=np.float64)
        self._check(self._lib.tetra_last_phase_ms(self._ctx, out.ctypes.data), "last_phase_ms")
        return out.tolist()

    def pack_dibits_device(self, dibits_ptr: int, n: int, packed_ptr: int):
        """Four dibits per byte for transport (device pointers, asynchronous on the context's stream)."""
        self._check(self._lib.tetra_pack_dibits(self._ctx, dibits_ptr, n, packed_ptr), "pack_dibits")

    def unpack_dibits_device(self, packed_ptr: int, n_blocks: int, packed_bytes: int, in_stride: int, dibits_ptr: int, out_stride: int):
        """Inverse of ``pack_dibits_device`` over the blocks an all-gather leaves (device pointers, asynchronous)."""
        self._check(self._lib.tetra_unpack_dibits(self._ctx, packed_ptr, n_blocks, packed_bytes, in_stride, dibits_ptr, out_stride),
                    "unpack_dibits")

    # ---- peer-memory all-gather of the dibit streams (include/tetra_b200.h: tetra_p2p_*, tetra_allgather_dibits) ----
    def p2p_create(self, rank: int, world: int, block_bytes: int) -> bytes:
        """Allocate this rank's receive buffer; returns its CUDA IPC handle (to be exchanged with the other ranks)."""
        h = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        self._check(self._lib.tetra_p2p_create(self._ctx, rank, world, block_bytes, C.addressof(h)), "p2p_create")
        return bytes(h)

    def p2p_buffer(self) -> int:
        return int(self._lib.tetra_p2p_buffer(self._ctx) or 0)

    def p2p_connect(self, handles: bytes):
        """`handles`: the IPC handles of all ranks, concatenated in rank order (one process per GPU)."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._check(self._lib.tetra_p2p_connect(self._ctx, C.addressof(buf)), "p2p_connect")

    def p2p_connect_ptrs(self, buffers):
        """`buffers[r]` = ``p2p_buffer()`` of rank r's context (contexts of one process)."""
        arr = (C.c_void_p * len(buffers))(*[C.c_void_p(int(b)) for b in buffers])
        self._check(self._lib.tetra_p2p_connect_ptrs(self._ctx, C.addressof(arr)), "p2p_connect_ptrs")

    def allgather_dibits_device(self, dibits_ptr: int, n: int, n_dibits_ptr: int, n_local: int, all_dibits_ptr: int, all_n_ptr: int = 0):
        """Pack, push to every peer, wait, unpack: asynchronous on the context's stream (device pointers)."""
        self._check(self._lib.tetra_allgather_dibits(self._ctx, dibits_ptr, n, n_dibits_ptr, n_local, all_dibits_ptr,
                                                     all_n_ptr or None), "allgather_dibits")

    def process_batch_allgather_device(self, iq_ptr: int, n_carriers: int, n_samples: int, pitch: int, dibits_ptr: int, cap: int,
                                       n_dibits_ptr: int, symbols_ptr: int, best_phase_ptr: int, ts_match_ptr: int,
                                       all_dibits_ptr: int, all_n_ptr: int = 0, stream=None, freq_offsets=None):
        """``process_batch_device`` with the exchange fused behind the slicer (``tetra_process_batch_allgather``): the
        finalize kernel itself stores every carrier's packed stream into all peers' receive buffers."""
        self._sync_rate()
        self.set_stream(stream)
        fo = None
        if freq_offsets is not None:
            fo = np.ascontiguousarray(freq_offsets, dtype=np.float64)
        self._check(self._lib.tetra_process_batch_allgather(
            self._ctx, iq_ptr, n_carriers, n_samples, pitch, fo.ctypes.data if fo is not None else None,
            dibits_ptr, cap, n_dibits_ptr, symbols_ptr or None, best_phase_ptr or None, ts_match_ptr or None,
            all_dibits_ptr, all_n_ptr or None), "process_batch_allgather")

    def p2p_status(self) -> int:
        st = C.c_int32(0)
        self._check(self._lib.tetra_p2p_status(self._ctx, C.byref(st)), "p2p_status")
        return int(st.value)

    def launch_count(self) -> int:
        return int(self._lib.tetra_launch_count(self._ctx))

    def synchronize(self):
        self._check(self._lib.tetra_synchronize(self._ctx), "synchronize")
