"""Frame-sync front end: the part of ``TetraDecoder`` that sits on the hot path.

Mirrors tetraear/core/decoder.py:140-169 (symbols_to_bits), :171-295 (find_sync) and the
threshold cascade of decode() (:845-856). The 22-bit TS1/TS2 agreement counts come from the
k_sync_match CUDA kernel (via SignalProcessor.process_batch(want_match=True)); the data-dependent
visiting order (jump +250 after a hit, adaptive retry) is replayed by the library's host routine
tetra_find_sync, which is integer-exact.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def symbols_to_bits(dibits) -> np.ndarray:
    """decoder.py:140-169 for 0..3 input: (MSB, LSB) per dibit. Pure index shuffling (no arithmetic)."""
    d = np.asarray(dibits).astype(np.int64) & 3
    bits = np.empty(2 * len(d), dtype=np.int64)
    bits[0::2] = d >> 1
    bits[1::2] = d & 1
    return bits


def find_sync(match: np.ndarray, n_dibits: int, threshold: float = 0.85, return_max_corr: bool = False):
    """decoder.py:171-295 over device-computed match counts of one carrier ([2*cap, 2] uint8)."""
    lib = _lib.load()
    nw = max(0, 2 * int(n_dibits) - 22 + 1)
    m = np.ascontiguousarray(match[:nw], dtype=np.uint8)
    pos = np.empty(max(nw // 250 + 2, 4), dtype=np.int32)
    mx = C.c_double(0.0)
    n = lib.tetra_find_sync(m.ctypes.data if nw else None, nw, float(threshold), pos.ctypes.data, len(pos), C.byref(mx))
    if n < 0:
        raise _lib.TetraError(f"tetra_find_sync failed ({n})")
    out = [int(p) for p in pos[:n]]
    return (out, float(mx.value)) if return_max_corr else out


def sync_cascade(match: np.ndarray, n_dibits: int):
    """decoder.py:845-856: thresholds 0.90 -> 0.85 -> 0.80 -> adaptive. Returns the sync positions."""
    lib = _lib.load()
    nw = max(0, 2 * int(n_dibits) - 22 + 1)
    m = np.ascontiguousarray(match[:nw], dtype=np.uint8)
    pos = np.empty(max(nw // 250 + 2, 4), dtype=np.int32)
    n = lib.tetra_sync_cascade(m.ctypes.data if nw else None, nw, pos.ctypes.data, len(pos))
    if n < 0:
        raise _lib.TetraError(f"tetra_sync_cascade failed ({n})")
    return [int(p) for p in pos[:n]]


def burst_slices(positions, n_dibits: int):
    """decoder.py:861-877: (start_symbol, start_bit, frame_number) of every 255-symbol burst."""
    out = []
    for pos in positions:
        start = pos - 216
        if start >= 0 and start // 2 + 255 <= n_dibits:
            out.append((start // 2, start, start // 510))
    return out
