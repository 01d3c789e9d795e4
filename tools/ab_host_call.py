"""A/B of the one-carrier host call (pinned / pageable complex64 in, host results out) between two builds of the library:
python tools/ab_host_call.py lib_a.so lib_b.so -- plain ctypes, so that builds with a different symbol set can be compared."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tetraear_b200 import synth


def run(path, n, pinned, reps=30):
    lib = C.CDLL(path)
    lib.tetra_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double]
    lib.tetra_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.tetra_dibit_capacity.restype = C.c_int64
    lib.tetra_dibit_capacity.argtypes = [C.c_void_p, C.c_int64]
    lib.tetra_destroy.argtypes = [C.c_void_p]
    ctx = C.c_void_p()
    assert lib.tetra_create(C.byref(ctx), 0, 2.4e6) == 0
    x = synth.carrier_iq(n, 0, snr_db=30.0)
    keep = None
    if pinned:
        keep = torch.from_numpy(x).pin_memory()
        x = keep.numpy()
    cap = lib.tetra_dibit_capacity(ctx, n)
    dib = np.zeros(cap, np.uint8); nd = np.zeros(1, np.int32); sym = np.zeros(cap + 1, np.complex64); ph = np.zeros(1, np.int32)

    def call():
        rc = lib.tetra_process_batch(ctx, x.ctypes.data, 1, n, n, None, dib.ctypes.data, cap, nd.ctypes.data, sym.ctypes.data, ph.ctypes.data, None, 0)
        assert rc == 0
    call(); call(); call()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)
    lib.tetra_destroy(ctx)
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


if __name__ == "__main__":
    for rep in range(2):
        for path in sys.argv[1:]:
            for n in (1 << 20, 131072):
                for pinned in (True, False):
                    med, lo, hi = run(os.path.abspath(path), n, pinned)
                    print("%-45s n=%7d %-8s median %.3f ms (min %.3f max %.3f)" % (os.path.basename(path), n, "pinned" if pinned else "pageable", med, lo, hi), flush=True)
