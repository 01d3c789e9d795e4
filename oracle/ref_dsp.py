"""CPU oracle for the TetraEar IQ -> dibit hot path (TEST INFRASTRUCTURE, not product).

A numpy/scipy restatement of the reference's algorithm, written from the behaviour of

  * tetraear/signal/processor.py:221-273   SignalProcessor.process
  * tetraear/signal/processor.py:51-83     filter_signal   (butter(4) + filtfilt)
  * tetraear/signal/processor.py:85-100    frequency_shift (NCO)
  * tetraear/signal/processor.py:168-219   extract_symbols (block timing pick)
  * tetraear/signal/processor.py:102-166   demodulate_dqpsk (differential slicer)
  * tetraear/core/decoder.py:140-169       symbols_to_bits
  * tetraear/core/decoder.py:171-295       find_sync
  * tetraear/core/decoder.py:835-857       decode() threshold cascade
  * tetraear/ui/modern.py:1921-1934        spectrum block

The filter arithmetic of the reference lives in an un-vendored dependency, SciPy
(requirements.txt:2 `scipy>=1.10.0`, this image: scipy 1.18.1): `scipy.signal.decimate`
(cheby1(8, 0.05, 0.8/q) SOS, `sosfiltfilt`, `[::q]`), `scipy.signal.butter` + `filtfilt`.
This module calls the same SciPy entry points (that *is* the reference's algorithm);
oracle/tetra_oracle.c restates those recursions from scratch in C.

Parity pinning: the reference's own tests hold no numeric vectors for this path
(SURVEY.md section 8c), so this oracle is pinned against outputs of the reference itself,
executed in the build container: tests/golden/*.npz, written by oracle/make_golden.py.
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _sig

SYMBOL_RATE = 18000
TARGET_RATE = 240000

# decoder.py:196-199
TS1 = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0], dtype=np.int64)
TS2 = np.array([0, 1, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 1, 0, 0], dtype=np.int64)
SYNC_LEN = 22


def decimation_factor(sample_rate: float) -> int:
    """processor.py:245-250 -- q = int(fs/240k) when fs > 480 kHz, else 1 (no decimation)."""
    if sample_rate > TARGET_RATE * 2:
        q = int(sample_rate / TARGET_RATE)
        if q > 1:
            return q
    return 1


def nco(samples: np.ndarray, freq_offset: float, fs: float) -> np.ndarray:
    """processor.py:97-100."""
    t = np.arange(len(samples)) / fs
    shift = np.exp(-1j * 2 * np.pi * freq_offset * t)   # named temporary: keeps numpy on the
    return samples * shift                              # same (non-aliased) multiply loop


def channel_filter(samples: np.ndarray, bandwidth: float, fs: float) -> np.ndarray:
    """processor.py:66-83 -- butter(4, clamp((bw/2)/(fs/2))) zero-phase."""
    if len(samples) == 0:
        return samples
    wn = min(0.99, max(0.01, (bandwidth / 2) / (fs / 2)))
    try:
        b, a = _sig.butter(4, wn, btype="low")
        return _sig.filtfilt(b, a, samples)
    except Exception:
        return samples


def timing_pick(samples: np.ndarray, fs: float):
    """processor.py:179-219 -> (symbols, best_phase, powers[phases tried])."""
    if len(samples) == 0:
        return np.array([], dtype=complex), 0, np.zeros(0)
    sps = int(fs / SYMBOL_RATE)
    if sps <= 1:
        return samples, 0, np.zeros(0)
    step = max(1, sps // 8)
    best, best_pow, pows = 0, -1.0, []
    for ph in range(0, sps, step):
        n = (len(samples) - ph) // sps
        if n <= 0:
            pows.append(np.nan)
            continue
        p = np.mean(np.abs(samples[ph + sps * np.arange(n)]) ** 2)
        pows.append(p)
        if p > best_pow:
            best_pow, best = p, ph
    n = (len(samples) - best) // sps
    return samples[best + sps * np.arange(n)], best, np.array(pows)


def slice_dqpsk(symbols: np.ndarray):
    """processor.py:120-166, vectorised. Returns (dibits uint8, phase_diff float64)."""
    if len(symbols) < 2:
        return np.array([], dtype=np.uint8), np.zeros(0)
    s = np.asarray(symbols)
    m = np.max(np.abs(s))
    if m > 0:
        s = s / m
    d = s[1:] * np.conj(s[:-1])
    ph = np.arctan2(np.imag(d), np.real(d))
    out = np.full(len(ph), 3, dtype=np.uint8)
    out[ph < 5 * np.pi / 8] = 1
    out[ph < 3 * np.pi / 8] = 0
    out[ph < -3 * np.pi / 8] = 2
    out[ph < -5 * np.pi / 8] = 3
    return out, ph


def process(samples: np.ndarray, freq_offset: float = 0.0, sample_rate: float = 2.4e6):
    """processor.py:221-273. Returns dict(dibits, symbols, best_phase, powers, phase_diff)."""
    samples = np.asarray(samples)
    if len(samples) == 0:
        return dict(dibits=np.array([], dtype=np.uint8), symbols=np.array([], dtype=complex),
                    best_phase=0, powers=np.zeros(0), phase_diff=np.zeros(0))
    rate = float(sample_rate)
    q = decimation_factor(rate)
    if q > 1:
        try:
            samples = _sig.decimate(samples, q)
            rate = rate / q
        except Exception:
            pass
    if freq_offset != 0:
        samples = nco(samples, freq_offset, rate)
    filt = channel_filter(samples, 25000, rate)
    syms, best, pows = timing_pick(filt, rate)
    dib, ph = slice_dqpsk(syms)
    return dict(dibits=dib, symbols=syms, best_phase=best, powers=pows, phase_diff=ph)


def symbols_to_bits(dibits: np.ndarray) -> np.ndarray:
    """decoder.py:140-169 for the 0..3 case (the only one this path produces): MSB first."""
    d = np.asarray(dibits).astype(np.int64) & 3
    bits = np.empty(2 * len(d), dtype=np.int64)
    bits[0::2] = d >> 1
    bits[1::2] = d & 1
    return bits


def match_counts(bits: np.ndarray) -> np.ndarray:
    """Number of agreeing bits vs TS1 and TS2 at every window start -> int[num_windows, 2]
    (decoder.py:237-240 evaluated at every position)."""
    bits = np.asarray(bits).astype(np.int64)
    nw = len(bits) - SYNC_LEN + 1
    if nw <= 0:
        return np.zeros((0, 2), dtype=np.int64)
    win = np.lib.stride_tricks.sliding_window_view(bits, SYNC_LEN)
    return np.stack([(win == TS1).sum(1), (win == TS2).sum(1)], axis=1)


def find_sync(bits: np.ndarray, threshold: float = 0.85):
    """decoder.py:171-295 -> (positions list[int], max_corr float).

    Same visiting order: on a hit at pos, jump to pos+250; max_corr and the adaptive
    re-search only see visited positions; TS1 is tested before TS2 and a TS1 hit hides TS2's
    correlation at that position from max_corr.
    """
    bits = np.asarray(bits)
    if len(bits) < SYNC_LEN:
        return [], 0.0
    mc = match_counts(bits)
    nw = len(mc)
    pos_list, max_corr, visited = [], 0.0, []
    i = 0
    while i < nw:
        found = False
        best_here = 0.0
        for k in range(2):
            c = mc[i, k] / SYNC_LEN
            best_here = max(best_here, c)
            max_corr = max(max_corr, c)
            if c >= threshold:
                pos_list.append(i)
                found = True
                break
        if best_here > 0:
            visited.append((i, best_here))
        i = i + 250 if found else i + 1
    if not pos_list and max_corr > 0.75 and max_corr >= (threshold - 0.15):
        adaptive = max(0.75, max_corr - 0.02)
        if adaptive < threshold:
            seen_until = -1  # positions < pos+250 of the last accepted hit are suppressed
            blocked = np.zeros(nw, dtype=bool)
            for p, c in visited:
                if c >= adaptive and not blocked[p]:
                    pos_list.append(p)
                    blocked[max(0, p - 250): min(nw, p + 250)] = True
    return pos_list, float(max_corr)


def sync_cascade(bits: np.ndarray):
    """decoder.py:845-856 -> positions handed to decode_frame."""
    pos, mx = find_sync(bits, 0.90)
    if not pos:
        pos, mx = find_sync(bits, 0.85)
        if not pos:
            pos, mx = find_sync(bits, 0.80)
            if not pos and mx >= 0.75:
                pos, _ = find_sync(bits, max(0.75, mx - 0.02))
    return pos


# protocol.py:162-163
SYNC_CONTINUOUS_DOWNLINK = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0], dtype=np.int64)
SYNC_DISCONTINUOUS_DOWNLINK = np.array([0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 1, 1], dtype=np.int64)
BURST_NORMAL_DOWNLINK, BURST_SYNCHRONIZATION = 2, 5        # protocol.py BurstType values


def crc16_ccitt_bits(bits) -> int:
    """protocol.py:332-347 -- bitwise CRC-16-CCITT (0x1021, init 0xFFFF), MSB first."""
    crc = 0xFFFF
    for bit in bits:
        crc ^= (int(bit) << 15)
        crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def check_crc(bits) -> bool:
    """protocol.py:291-330 -- the reference's soft CRC check (<= 2 bit errors, forward or reversed payload)."""
    bits = np.asarray(bits).astype(np.int64)
    if len(bits) < 16:
        return False
    ones = int(bits.sum())
    if ones == 0 or ones == len(bits):
        return False
    payload, recv = bits[:-16], bits[-16:]
    recv_word = int("".join(str(int(b)) for b in recv), 2)
    if bin(crc16_ccitt_bits(payload) ^ recv_word).count("1") <= 2:
        return True
    return bin(crc16_ccitt_bits(payload[::-1]) ^ recv_word).count("1") <= 2


def parse_burst(symbols):
    """protocol.py:192-289 for one 255-symbol slot -> (burst_type, crc_ok, data_bits)."""
    sym = np.asarray(symbols[:255]).astype(np.int64)
    bits = np.empty(510, dtype=np.int64)
    bits[0::2] = (sym >> 1) & 1
    bits[1::2] = sym & 1
    w = bits[255:277]
    m = max(int((w == SYNC_CONTINUOUS_DOWNLINK).sum()), int((w == SYNC_DISCONTINUOUS_DOWNLINK).sum())) / 22
    if m > 0.8:
        btype, data = BURST_SYNCHRONIZATION, bits
    else:
        btype, data = BURST_NORMAL_DOWNLINK, np.concatenate([bits[0:108], bits[122:230]])
    return btype, check_crc(data), data


def decode_bursts(dibits, positions):
    """decoder.py:861-888 up to the parse_burst call inside decode_frame (:986-992): for every sync position the
    (start_symbol, frame_number, burst_type, crc_ok) of its 255-symbol slot; positions without a complete slot
    are dropped exactly like decode() drops them."""
    out = []
    d = np.asarray(dibits)
    for pos in positions:
        start = pos - 216
        if start < 0:
            continue
        s0 = start // 2
        if s0 + 255 > len(d):
            continue
        btype, crc_ok, _ = parse_burst(d[s0:s0 + 255])
        out.append((s0, start // 510, btype, int(crc_ok)))
    return out


# scanner.py:133-134
SCANNER_SYNC_PATTERN = np.array([0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0,
                                 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0], dtype=np.int64)


def _wrapped_phase_diffs(x):
    """diff(angle(x)) folded into [-pi, pi) exactly as scanner.py:74-75 / :113-114 do."""
    pd = np.diff(np.angle(x))
    return (pd + np.pi) % (2 * np.pi) - np.pi


def analyze_signal(samples, sample_rate: float = 2.4e6, symbol_rate: float = 18000.0):
    """The per-sample part of TetraSignalDetector.analyze_signal (signal/scanner.py:42-147, 204-231), vectorised:
    calculate_power, detect_tetra_modulation, detect_sync_pattern, check_power_stability."""
    x = np.asarray(samples).astype(np.complex128)
    out = {}
    # calculate_power (:42-55)
    out["power_db"] = float(10 * np.log10(np.mean(np.abs(x) ** 2) + 1e-10)) if x.size else -120.0
    # detect_tetra_modulation (:57-96)
    if len(x) < 1000:
        out["modulation_matches"], out["modulation_confidence"] = 0, 0.0
    else:
        xn = x / (np.abs(x).max() + 1e-10)
        pd = _wrapped_phase_diffs(xn)
        expected = np.array([-np.pi, -3 * np.pi / 4, -np.pi / 2, -np.pi / 4, 0, np.pi / 4, np.pi / 2, 3 * np.pi / 4])
        dist = np.abs(expected[None, :] - pd[:, None]).min(axis=1)
        m = int((dist < np.pi / 8).sum())
        out["modulation_matches"], out["modulation_confidence"] = m, m / len(pd)
    out["is_tetra_modulation"] = out["modulation_confidence"] > 0.4
    # detect_sync_pattern (:98-147)
    ds = max(1, int(sample_rate / symbol_rate / 10))
    sym = x[::ds]
    corr = 0.0
    if len(sym) >= 100:
        pd = _wrapped_phase_diffs(sym)
        q = np.round(pd / (np.pi / 4)) * (np.pi / 4)
        bits = (np.abs(q) < np.pi / 8).astype(np.int64)
        if len(bits) >= 31:
            n_win = len(bits) - 31
            if n_win > 0:
                win = np.lib.stride_tricks.sliding_window_view(bits, 31)[:n_win]
                corr = float((win == SCANNER_SYNC_PATTERN).sum(axis=1).max() / 31)
    out["sync_correlation"], out["sync_detected"] = corr, corr > 0.75
    # check_power_stability (:204-231)
    if len(x) < 5 * 1000:
        out["power_stable"] = False
    else:
        w = len(x) // 5
        pw = [10 * np.log10(np.mean(np.abs(x[i * w:(i + 1) * w]) ** 2) + 1e-10) for i in range(5)]
        out["power_stable"] = bool(np.std(pw) < 10.0)
    return out


def spectrum_db(samples: np.ndarray, n_fft: int = 2048) -> np.ndarray:
    """modern.py:1924-1934 on the first n_fft samples."""
    w = np.hanning(n_fft)
    f = np.fft.fftshift(np.fft.fft(np.asarray(samples[:n_fft]) * w))
    return 20 * np.log10(np.abs(f) / n_fft + 1e-20)


def stft_db(samples: np.ndarray, n_fft: int = 4096, hop: int = 1024) -> np.ndarray:
    """Config-5 generalisation: the modern.py:1924-1934 block applied at every hop."""
    x = np.asarray(samples)
    rows = (len(x) - n_fft) // hop + 1 if len(x) >= n_fft else 0
    out = np.empty((rows, n_fft))
    w = np.hanning(n_fft)
    for r in range(rows):
        f = np.fft.fftshift(np.fft.fft(x[r * hop: r * hop + n_fft] * w))
        out[r] = 20 * np.log10(np.abs(f) / n_fft + 1e-20)
    return out


def presence_afc(samples: np.ndarray, sample_rate: float, n_fft: int = 2048) -> dict:
    """The signal-presence / AFC block of CaptureThread.run, tetraear/ui/modern.py:1919-2003, restated (the module needs
    PyQt6 and cannot be imported here): spectrum of the first n_fft samples (:1921-1934), mean and peak of the centre
    25 kHz (:1948-1957), the peak's frequency = the AFC offset handed to process() (:1959-1966, :2019-2022), noise floor
    from everything more than 10 bins outside (:1968-1984), and the verdict (:1986-1999)."""
    x = np.asarray(samples)
    if len(x) < n_fft:
        return dict(signal_power=0.0, peak_power=0.0, peak_freq_offset=0.0, noise_floor=0.0, snr=0.0, is_signal_strong=False)
    power = spectrum_db(x, n_fft)
    freqs = np.fft.fftshift(np.fft.fftfreq(n_fft, 1 / sample_rate))
    center_idx = len(power) // 2
    bandwidth_bins = int(25000 / (sample_rate / n_fft))
    start_idx = max(0, center_idx - bandwidth_bins // 2)
    end_idx = min(len(power), center_idx + bandwidth_bins // 2)
    signal_power = np.mean(power[start_idx:end_idx])
    peak_power = np.max(power[start_idx:end_idx])
    peak_idx = start_idx + int(np.argmax(power[start_idx:end_idx]))
    noise = []
    if max(0, start_idx - 10) > 0:
        noise.extend(power[0:max(0, start_idx - 10)])
    if len(power) > min(len(power), end_idx + 10):
        noise.extend(power[min(len(power), end_idx + 10):len(power)])
    noise_floor = np.mean(noise) if noise else -100
    snr = signal_power - noise_floor
    strong = bool(snr > 15 and peak_power > -70 and (peak_power - signal_power) > 3)
    return dict(signal_power=float(signal_power), peak_power=float(peak_power), peak_freq_offset=float(freqs[peak_idx]),
                noise_floor=float(noise_floor), snr=float(snr), is_signal_strong=strong)


def analyze_verdict(power_db, mod_conf, sync_corr, power_stable, frames_valid=False, crc_rate=0.0, bottom_threshold=-85):
    """TetraSignalDetector.analyze_signal's combination of its detectors, tetraear/signal/scanner.py:233-289."""
    is_mod, has_sync = mod_conf > 0.4, sync_corr > 0.75
    if has_sync and is_mod:
        confidence = mod_conf * 0.4 + sync_corr * 0.4 + crc_rate * 0.2
    elif has_sync:
        confidence = sync_corr * 0.6
    elif is_mod:
        confidence = mod_conf * 0.5
    else:
        confidence = 0.0
    is_tetra = bool(is_mod and has_sync and power_stable)
    if frames_valid:
        is_tetra = True
        confidence = max(confidence, 0.7)
    return dict(power_db=power_db, is_tetra=is_tetra, confidence=confidence, modulation_confidence=mod_conf,
                sync_detected=bool(has_sync), sync_correlation=sync_corr, frames_validated=bool(frames_valid),
                crc_pass_rate=crc_rate, power_stable=bool(power_stable), signal_present=bool(power_db > bottom_threshold))
