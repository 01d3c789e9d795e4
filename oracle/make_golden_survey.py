"""Writes tests/golden/survey.npz by RUNNING THE REFERENCE on a seeded wideband capture: for every channel offset of the
25 kHz grid, TetraSignalDetector's methods and analyze_signal (tetraear/signal/scanner.py:42-289) on the reference's own
SignalProcessor.frequency_shift (signal/processor.py:85-100) of the capture -- what FrequencyScanner.scan_range
(scanner.py:383-445) computes after retuning to that channel -- plus the presence / AFC block of ui/modern.py:1945-2012 as
restated in oracle/ref_dsp.py (that module needs PyQt6). Build-container only.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_survey
"""
from __future__ import annotations

import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "survey.npz")
N_SAMPLES = 1 << 16
FIELDS = ("power_db", "modulation_confidence", "is_tetra_modulation", "sync_correlation", "sync_detected", "power_stable",
          "signal_power", "peak_power", "peak_freq_offset", "noise_floor", "snr", "is_signal_strong",
          "is_tetra", "confidence", "signal_present")


def capture():
    """The survey capture: carriers on 40 % of the 25 kHz grid (strong, so that a centre channel shows a spike), complex64."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from tetraear_b200 import synth
    rng = np.random.default_rng(91)
    active = rng.random(96) < 0.4
    active[48] = True
    active[47] = active[49] = False
    return synth.wideband_capture(N_SAMPLES, seed=9, active=active, snr_db=35.0)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    bs = types.ModuleType("bitstring")
    bs.BitArray = type("BitArray", (), {})
    sys.modules.setdefault("bitstring", bs)
    logging.disable(logging.CRITICAL)
    from tetraear.signal.scanner import TetraSignalDetector
    from tetraear.signal.processor import SignalProcessor
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import ref_dsp
    det = TetraSignalDetector(sample_rate=2.4e6)
    proc = SignalProcessor(sample_rate=2.4e6)
    x, active, freqs = capture()
    x128 = x.astype(np.complex128)
    rows = np.zeros((len(freqs), len(FIELDS)))
    for k, f in enumerate(freqs):
        xs = proc.frequency_shift(x128, f)
        a = det.analyze_signal(xs)                       # frame validation: False here (decoder import needs bitstring)
        is_mod, conf = det.detect_tetra_modulation(xs)
        p = ref_dsp.presence_afc(xs, 2.4e6)
        rows[k] = [a["power_db"], a["modulation_confidence"], float(is_mod), a["sync_correlation"], float(a["sync_detected"]),
                   float(a["power_stable"]), p["signal_power"], p["peak_power"], p["peak_freq_offset"], p["noise_floor"], p["snr"],
                   float(p["is_signal_strong"]), float(a["is_tetra"]), a["confidence"], float(a["signal_present"])]
        print(k, int(active[k]), np.round(rows[k], 4))
    import hashlib
    np.savez_compressed(OUT, rows=rows, freqs=freqs, active=active, fields=np.array(FIELDS),
                        input_sha256=hashlib.sha256(np.ascontiguousarray(x).view(np.uint8)).hexdigest())


if __name__ == "__main__":
    main()
