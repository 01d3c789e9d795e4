"""Seeded synthetic TETRA-like IQ generators (host side, numpy only).

These produce the benchmark / parity inputs described in BASELINE.md section 3:
pi/4-DQPSK (or region-centred DQPSK) symbols, root-raised-cosine shaped
(alpha = 0.35, span +-8 symbols) at 130 samples/symbol for a 2.4 MS/s stream
(the reference samples at stride 13 after its /10 decimation, reference
tetraear/signal/processor.py:183), unit mean power, AWGN, rounded through
complex64 -- the dtype that crosses the C ABI.

Nothing here is on the demodulation path; it only makes inputs.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 2.4e6
SPS_FULL = 130            # samples per symbol at 2.4 MS/s used by the generators
ALPHABET_PI4 = np.array([1, 3, -1, -3]) * (np.pi / 4)       # true pi/4-DQPSK increments
ALPHABET_CENTRED = np.array([0.0, 0.5, -0.5, 1.0]) * np.pi  # centres of the reference's slicer regions


def rrc_taps(sps: int, alpha: float = 0.35, span: int = 8) -> np.ndarray:
    """Root-raised-cosine impulse response, `span` symbols each side, unit energy."""
    n = np.arange(-span * sps, span * sps + 1, dtype=np.float64)
    t = n / sps
    h = np.empty_like(t)
    eps = 1e-9
    for i, ti in enumerate(t):
        if abs(ti) < eps:
            h[i] = 1.0 - alpha + 4.0 * alpha / np.pi
        elif abs(abs(ti) - 1.0 / (4.0 * alpha)) < eps:
            h[i] = (alpha / np.sqrt(2.0)) * (
                (1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha))
                + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            num = np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))
            den = np.pi * ti * (1 - (4 * alpha * ti) ** 2)
            h[i] = num / den
    return h / np.sqrt(np.sum(h * h))


_RRC_CACHE: dict = {}


def dqpsk_baseband(n_samples: int, seed: int, alphabet: str = "pi4", sps: int = SPS_FULL,
                   return_increments: bool = False):
    """Noise-free unit-power shaped DQPSK baseband, complex128, length n_samples."""
    rng = np.random.default_rng([seed, 0x7E7A])
    span = 8
    n_sym = n_samples // sps + 2 * span + 2
    alpha_set = ALPHABET_PI4 if alphabet == "pi4" else ALPHABET_CENTRED
    inc_idx = rng.integers(0, 4, size=n_sym)
    phase = np.cumsum(alpha_set[inc_idx])
    syms = np.exp(1j * phase)
    key = (sps, span)
    if key not in _RRC_CACHE:
        _RRC_CACHE[key] = rrc_taps(sps, 0.35, span)
    h = _RRC_CACHE[key]
    # polyphase interpolation: y[sps*k + r] = sum_j syms[k-j] h[sps*j + r]
    hp = np.zeros((2 * span + 1) * sps)
    hp[: len(h)] = h
    hp = hp.reshape(2 * span + 1, sps)            # hp[j, r] = h[sps*j + r]
    full = np.zeros((n_sym + 2 * span, sps), dtype=np.complex128)
    for j in range(2 * span + 1):
        full[j: j + n_sym, :] += syms[:, None] * hp[j][None, :]
    y = full.reshape(-1)[span * sps: span * sps + n_samples]
    y = y / np.sqrt(np.mean(np.abs(y) ** 2))
    if return_increments:
        return y, inc_idx
    return y


def carrier_iq(n_samples: int, seed: int, snr_db: float = 30.0, alphabet: str = "pi4",
               sps: int = SPS_FULL, decim: int = 1) -> np.ndarray:
    """One 25 kHz carrier at baseband + white noise over the full band -> complex64[n_samples].

    ``decim`` > 1 shapes at ``sps`` samples per symbol and keeps every ``decim``-th sample: sps=400, decim=3 is the true
    TETRA rate, 18 000 symbols/s at 2.4 MS/s = 133.33 samples per symbol."""
    y = dqpsk_baseband(n_samples * decim, seed, alphabet, sps)[::decim]
    if decim > 1:
        y = y / np.sqrt(np.mean(np.abs(y) ** 2))
    rng = np.random.default_rng([seed, 0xA17C])
    sigma = np.sqrt(10.0 ** (-snr_db / 10.0) / 2.0)
    noise = sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    return (y + noise).astype(np.complex64)


def wideband_capture(n_samples: int, seed: int = 3, n_channels: int = 96, spacing_hz: float = 25e3,
                     active=None, snr_db: float = 30.0):
    """Sum of carriers on a 25 kHz grid (channel k at (k - n_channels/2)*spacing) -> complex64.

    Returns (iq, active_mask, channel_freqs_hz).
    """
    rng = np.random.default_rng([seed, 0xC4A2])
    if active is None:
        active = rng.random(n_channels) < 0.75
    active = np.asarray(active, dtype=bool)
    freqs = (np.arange(n_channels) - n_channels // 2) * spacing_hz
    n = np.arange(n_samples, dtype=np.float64)
    x = np.zeros(n_samples, dtype=np.complex128)
    for k in range(n_channels):
        if not active[k]:
            continue
        bb = dqpsk_baseband(n_samples, 1000 + k, "centred" if k % 2 else "pi4")
        x += bb * np.exp(2j * np.pi * (freqs[k] / SAMPLE_RATE) * n)
    sigma = np.sqrt(10.0 ** (-snr_db / 10.0) / 2.0)
    x += sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    return x.astype(np.complex64), active, freqs


def stft_test_signal(n_samples: int = 2_400_000, seed: int = 5) -> np.ndarray:
    """Config-5 input: one carrier + 3 tones + noise, complex64."""
    rng = np.random.default_rng([seed, 0x57F7])
    n = np.arange(n_samples, dtype=np.float64)
    x = 0.5 * dqpsk_baseband(n_samples, seed, "pi4")
    for f, a in ((300e3, 0.2), (-712.5e3, 0.05), (1.0e6, 0.01)):
        x = x + a * np.exp(2j * np.pi * (f / SAMPLE_RATE) * n + 1j * rng.uniform(0, 2 * np.pi))
    x = x + 1e-3 * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    return x.astype(np.complex64)
