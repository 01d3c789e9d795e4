"""Writes tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (imported from /root/reference).

Build-container only: /root/reference does not exist on the GPU box. Inputs are regenerated from
seeds by tetraear_b200.synth at test time; each fixture stores a checksum of its input so a
generator drift is detected instead of silently comparing different signals.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden
"""
from __future__ import annotations

import hashlib
import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

# (name, generator kwargs, n_samples, sample_rate, freq_offset)
CASES = [
    ("cfg1_pi4_2p20", dict(seed=0, alphabet="pi4", snr_db=30.0), 1 << 20, 2.4e6, 0.0),
    ("cfg1_pi4_2p20_fo", dict(seed=0, alphabet="pi4", snr_db=30.0), 1 << 20, 2.4e6, 1234.5),
    ("centred_2p18", dict(seed=1, alphabet="centred", snr_db=30.0), 1 << 18, 2.4e6, 0.0),
    ("short_24000", dict(seed=2, alphabet="centred", snr_db=25.0), 24000, 2.4e6, 0.0),
    ("gui_131072", dict(seed=3, alphabet="centred", snr_db=15.0), 131072, 2.4e6, 0.0),
    ("gui_131072_fo", dict(seed=3, alphabet="pi4", snr_db=20.0), 131072, 2.4e6, -5000.0),
    ("odd_100003", dict(seed=4, alphabet="pi4", snr_db=20.0), 100003, 2.4e6, 0.0),
    ("min_fast_16384", dict(seed=5, alphabet="centred", snr_db=30.0), 16384, 2.4e6, 0.0),
    ("rate_1p8M", dict(seed=6, alphabet="centred", snr_db=30.0, sps=98), 65536, 1.8e6, 0.0),
    ("rate_2p048M", dict(seed=7, alphabet="centred", snr_db=30.0, sps=112), 65536, 2.048e6, 250.0),
    ("rate_1M", dict(seed=8, alphabet="centred", snr_db=30.0, sps=52), 50000, 1.0e6, 0.0),
    ("rate_240k", dict(seed=9, alphabet="centred", snr_db=30.0, sps=13), 20000, 240e3, 0.0),
    ("tiny_100", dict(seed=10, alphabet="centred", snr_db=30.0), 100, 2.4e6, 0.0),
    ("tiny_20", dict(seed=11, alphabet="centred", snr_db=30.0), 20, 2.4e6, 0.0),
    ("tiny_300", dict(seed=12, alphabet="centred", snr_db=30.0), 300, 2.4e6, 0.0),
    ("one_symbol_200", dict(seed=13, alphabet="centred", snr_db=30.0), 200, 2.4e6, 0.0),      # 20 samples at 240 kS/s: one symbol, no dibit
    # the true TETRA symbol rate (133.33 samples per symbol at 2.4 MS/s): the reference still samples every 13th of 240 kS/s,
    # so its timing drifts through the block and decisions sit near the region borders
    ("truerate_2p18", dict(seed=14, alphabet="pi4", snr_db=30.0, sps=400, decim=3), 1 << 18, 2.4e6, 0.0),
    ("truerate_2p18_fo", dict(seed=15, alphabet="centred", snr_db=30.0, sps=400, decim=3), 1 << 18, 2.4e6, 2000.0),
]
SYNC_THRESHOLDS = (0.90, 0.85, 0.80, 0.78)


def helper_signal(n: int = 4800) -> np.ndarray:
    """Tone + noise in the style of the reference's `sample_iq_samples` fixture (tests/conftest.py:53-67), seeded."""
    rng = np.random.default_rng(1234)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1 + 0.5 * np.exp(
        2j * np.pi * 1000 * np.arange(n) / 2.4e6)


def input_digest(x: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(x).view(np.uint8)).hexdigest()


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, ".."))
    # tetraear.core.decoder imports bitstring (not installed; only decode_frame needs it)
    bs = types.ModuleType("bitstring")
    bs.BitArray = type("BitArray", (), {})
    sys.modules.setdefault("bitstring", bs)
    logging.disable(logging.CRITICAL)
    import scipy
    from tetraear.signal.processor import SignalProcessor
    from tetraear.core.decoder import TetraDecoder
    from tetraear_b200 import synth

    os.makedirs(OUT, exist_ok=True)
    dec = TetraDecoder()
    versions = np.array([np.__version__, scipy.__version__])
    only_missing = "--missing" in sys.argv
    for name, gen, n, fs, fo in CASES:
        if only_missing and os.path.exists(os.path.join(OUT, name + ".npz")):
            continue
        x = synth.carrier_iq(n, **gen)
        sp = SignalProcessor(fs)
        dib = sp.process(x.astype(np.complex128), fo)
        syms = np.asarray(sp.symbols)
        # best phase: recover it from the reference's own intermediate products
        best = -1
        if len(syms):
            from scipy import signal as _s
            z = x.astype(np.complex128)
            rate = fs
            if fs > 480000 and int(fs / 240000) > 1:
                try:
                    z = _s.decimate(z, int(fs / 240000)); rate = fs / int(fs / 240000)
                except Exception:
                    pass
            if fo != 0:
                z = sp.frequency_shift(z, fo, sample_rate=rate)
            z = sp.filter_signal(z, 25000, sample_rate=rate)
            k = int(rate / 18000)
            for ph in range(max(k, 1)):
                cand = z[ph::k][: len(syms)] if k > 1 else z
                if len(cand) == len(syms) and np.array_equal(cand, syms) and (len(z) - ph) // max(k, 1) == len(syms):
                    best = ph
                    break
        out = dict(dibits=dib, symbols=syms, best_phase=np.int32(best), n_samples=np.int64(n),
                   sample_rate=np.float64(fs), freq_offset=np.float64(fo), input_sha256=np.array(input_digest(x)),
                   versions=versions)
        if len(dib):
            bits, mapped = dec.symbols_to_bits(dib)
            out["bits_sha256"] = np.array(hashlib.sha256(np.asarray(bits, dtype=np.uint8)).hexdigest())
            for th in SYNC_THRESHOLDS:
                pos, mx = dec.find_sync(bits, threshold=th, return_max_corr=True)
                out["sync_pos_%03d" % round(th * 100)] = np.asarray(pos, dtype=np.int32)
                out["sync_max_%03d" % round(th * 100)] = np.float64(mx)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, len(dib), len(syms), best)

    if only_missing and os.path.exists(os.path.join(OUT, "helpers.npz")):
        return
    # ---- helper methods, exercised the way the reference's unit tests do (tests/unit/test_signal_processor.py) ----
    xs = helper_signal()
    sp = SignalProcessor(2.4e6)
    h = dict(input_sha256=np.array(input_digest(xs)), versions=versions)
    h["filter_25k"] = sp.filter_signal(xs, bandwidth=25000)
    h["filter_50k"] = sp.filter_signal(xs, bandwidth=50000)
    h["filter_240k"] = sp.filter_signal(xs, bandwidth=25000, sample_rate=240000.0)
    h["shift_1k"] = sp.frequency_shift(xs, 1000)
    h["shift_m7k_240k"] = sp.frequency_shift(xs, -7777.7, sample_rate=240000.0)
    h["extract_2p4M"] = sp.extract_symbols(xs)
    h["extract_1M"] = sp.extract_symbols(xs, sample_rate=1.0e6)
    h["extract_240k"] = sp.extract_symbols(xs, sample_rate=240000.0)
    h["demod"] = sp.demodulate_dqpsk(xs)
    h["resample_half"] = sp.resample(xs, 1.2e6)
    h["resample_up"] = sp.resample(xs[:3000], 3.6e6)
    h["resample_odd"] = sp.resample(xs[:3001], 1.0e6)
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **h)
    print("helpers written")


if __name__ == "__main__":
    main()
