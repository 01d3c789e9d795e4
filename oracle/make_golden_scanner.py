"""Writes tests/golden/scanner.npz by RUNNING THE REFERENCE's TetraSignalDetector (tetraear/signal/scanner.py:42-231)
on seeded captures. Build-container only.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_scanner
"""
from __future__ import annotations

import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "scanner.npz")


def captures():
    """(name, complex64 capture) list: DQPSK carriers, tone + noise (the reference's fixture style), noise, a burst
    that switches off (unstable power), and a short block."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from tetraear_b200 import synth
    rng = np.random.default_rng(77)
    n = 1 << 16
    out = [("carrier_pi4", synth.carrier_iq(n, 400, snr_db=25.0, alphabet="pi4")),
           ("carrier_centred", synth.carrier_iq(n, 401, snr_db=10.0, alphabet="centred"))]
    tone = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1 + 0.5 * np.exp(2j * np.pi * 1000 * np.arange(n) / 2.4e6)
    out.append(("tone_noise", tone.astype(np.complex64)))
    out.append(("noise", ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.05).astype(np.complex64)))
    burst = synth.carrier_iq(n, 402, snr_db=30.0).copy()
    burst[n // 5:] *= 1e-7                                   # four of five windows 140 dB down: unstable
    out.append(("burst_off", burst))
    out.append(("short_3000", synth.carrier_iq(3000, 403, snr_db=20.0)))
    out.append(("tiny_500", synth.carrier_iq(500, 404, snr_db=20.0)))
    return out


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    bs = types.ModuleType("bitstring")
    bs.BitArray = type("BitArray", (), {})
    sys.modules.setdefault("bitstring", bs)
    logging.disable(logging.CRITICAL)
    from tetraear.signal.scanner import TetraSignalDetector
    det = TetraSignalDetector(sample_rate=2.4e6)
    res = {}
    for name, x in captures():
        x128 = x.astype(np.complex128)
        power = det.calculate_power(x128)
        is_mod, conf = det.detect_tetra_modulation(x128)
        has_sync, corr = det.detect_sync_pattern(x128)
        stable = det.check_power_stability(x128)
        res[name] = np.array([power, conf, float(is_mod), corr, float(has_sync), float(stable)], dtype=np.float64)
        print(name, res[name])
    np.savez_compressed(OUT, **res)


if __name__ == "__main__":
    main()
