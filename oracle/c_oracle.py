"""ctypes loader of oracle/libtetra_oracle.so (TEST INFRASTRUCTURE: the plain-C restatement, tetra_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtetra_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "tetra_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_process.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int64),
                                        C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        _lib.oracle_find_sync.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        _lib.oracle_sync_cascade.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        _lib.oracle_symbols_to_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _lib.oracle_symbols_to_bits.restype = None
        _lib.oracle_spectrum_db.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.oracle_spectrum_db.restype = None
        _lib.oracle_butter4.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        _lib.oracle_cheby1_sos8.argtypes = [C.c_double, C.c_double, C.c_void_p]
    return _lib


def process(samples, freq_offset=0.0, sample_rate=2.4e6):
    """SignalProcessor(sample_rate).process(samples, freq_offset) -> dict(dibits, symbols, best_phase)."""
    lib = load()
    x = np.ascontiguousarray(samples, dtype=np.complex128)
    n = len(x)
    dib = np.empty(max(n, 1), dtype=np.uint8)
    sym = np.empty(max(n, 1), dtype=np.complex128)
    nd, ns, ph = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    lib.oracle_process(x.ctypes.data, n, float(sample_rate), float(freq_offset), dib.ctypes.data, C.byref(nd),
                       sym.ctypes.data, C.byref(ns), C.byref(ph))
    return dict(dibits=dib[: nd.value].copy(), symbols=sym[: ns.value].copy(), best_phase=int(ph.value))


def symbols_to_bits(dibits):
    lib = load()
    d = np.ascontiguousarray(dibits, dtype=np.uint8)
    bits = np.empty(2 * len(d), dtype=np.uint8)
    lib.oracle_symbols_to_bits(d.ctypes.data, len(d), bits.ctypes.data)
    return bits


def find_sync(bits, threshold=0.85):
    lib = load()
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    pos = np.empty(len(b) // 250 + 4, dtype=np.int32)
    mx = C.c_double(0.0)
    n = lib.oracle_find_sync(b.ctypes.data, len(b), float(threshold), pos.ctypes.data, len(pos), C.byref(mx))
    return [int(p) for p in pos[:n]], float(mx.value)


def sync_cascade(bits):
    lib = load()
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    pos = np.empty(len(b) // 250 + 4, dtype=np.int32)
    n = lib.oracle_sync_cascade(b.ctypes.data, len(b), pos.ctypes.data, len(pos))
    return [int(p) for p in pos[:n]]


def spectrum_db(samples, n_fft=2048):
    lib = load()
    x = np.ascontiguousarray(np.asarray(samples[:n_fft]), dtype=np.complex128)
    out = np.empty(n_fft, dtype=np.float64)
    lib.oracle_spectrum_db(x.ctypes.data, n_fft, out.ctypes.data)
    return out
