"""Writes tests/golden/bursts.npz by RUNNING THE REFERENCE's TetraProtocolParser.parse_burst
(tetraear/core/protocol.py:192-347) on seeded 255-symbol slots. Build-container only.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_bursts
"""
from __future__ import annotations

import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "bursts.npz")


def burst_frames(seed=2024, n_random=96):
    """Seeded slots: random symbols, slots with a sync pattern at bit 255, and slots whose payload carries a
    valid CRC-16 (forward or reversed, with 0..3 flipped CRC bits) so that every branch of _check_crc is hit."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import ref_dsp
    rng = np.random.default_rng(seed)
    frames = [rng.integers(0, 4, size=255) for _ in range(n_random)]
    frames.append(np.zeros(255, dtype=np.int64))                     # all zero bits: CRC check refuses
    frames.append(np.full(255, 3, dtype=np.int64))                   # all one bits

    def to_syms(bits):
        return (bits[0::2] << 1) | bits[1::2]

    for k in range(24):
        bits = rng.integers(0, 2, size=510)
        pat = ref_dsp.SYNC_CONTINUOUS_DOWNLINK if k % 2 else ref_dsp.SYNC_DISCONTINUOUS_DOWNLINK
        p = pat.copy()
        p[rng.choice(22, size=k % 6, replace=False)] ^= 1            # 0..5 errors: both sides of the 0.8 limit
        bits[255:277] = p
        frames.append(to_syms(bits))
    for k in range(32):
        bits = rng.integers(0, 2, size=510)
        sync_burst = k % 4 == 3
        if sync_burst:
            bits[255:277] = ref_dsp.SYNC_CONTINUOUS_DOWNLINK
            data_idx = np.arange(510)
        else:
            bits[255:277] = rng.integers(0, 2, size=22) * 0 + np.array([0, 1] * 11)   # far from both patterns
            data_idx = np.concatenate([np.arange(0, 108), np.arange(122, 230)])
        payload = bits[data_idx[:-16]]
        crc = ref_dsp.crc16_ccitt_bits(payload[::-1] if k % 2 else payload)
        crc_bits = np.array([(crc >> i) & 1 for i in range(15, -1, -1)])
        crc_bits[rng.choice(16, size=k % 4, replace=False)] ^= 1      # 0..3 CRC bit errors
        bits[data_idx[-16:]] = crc_bits
        frames.append(to_syms(bits))
    return np.stack(frames).astype(np.uint8)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    bs = types.ModuleType("bitstring")
    bs.BitArray = type("BitArray", (), {})
    sys.modules.setdefault("bitstring", bs)
    logging.disable(logging.CRITICAL)
    from tetraear.core.protocol import TetraProtocolParser
    frames = burst_frames()
    parser = TetraProtocolParser()
    btype, crc = [], []
    for f in frames:
        b = parser.parse_burst(f.astype(np.int64), slot_number=0)
        btype.append(b.burst_type.value)
        crc.append(int(bool(b.crc_ok)))
    np.savez_compressed(OUT, frames=frames, burst_type=np.array(btype, dtype=np.int32), crc_ok=np.array(crc, dtype=np.int32))
    print("bursts:", len(frames), "sync:", int(np.sum(np.array(btype) == 5)), "crc_ok:", int(np.sum(crc)))


if __name__ == "__main__":
    main()
