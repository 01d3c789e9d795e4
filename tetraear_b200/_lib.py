"""ctypes binding of libtetra_b200.so (include/tetra_b200.h). No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_ctx_p = C.c_void_p

_SIGS = {
    "tetra_create": (C.c_int, [C.POINTER(c_ctx_p), C.c_int, C.c_double]),
    "tetra_destroy": (None, [c_ctx_p]),
    "tetra_last_error": (C.c_char_p, [c_ctx_p]),
    "tetra_set_sample_rate": (C.c_int, [c_ctx_p, C.c_double]),
    "tetra_set_stream": (C.c_int, [c_ctx_p, C.c_void_p]),
    "tetra_set_h2d_chunk": (C.c_int, [c_ctx_p, C.c_int64]),
    "tetra_synchronize": (C.c_int, [c_ctx_p]),
    "tetra_dibit_capacity": (C.c_int64, [c_ctx_p, C.c_int64]),
    "tetra_symbol_count": (C.c_int64, [c_ctx_p, C.c_int64, C.c_int32]),
    "tetra_process_batch": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32]),
    "tetra_process_batch_sync": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "tetra_sync_positions": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_void_p]),
    "tetra_analyze_signal": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]),
    "tetra_survey_wideband": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "tetra_process_batch_u8": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p]),
    "tetra_parse_bursts": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "tetra_process_wideband": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tetra_pack_dibits": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tetra_unpack_dibits": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "tetra_p2p_create": (C.c_int, [c_ctx_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "tetra_p2p_buffer": (C.c_void_p, [c_ctx_p]),
    "tetra_p2p_connect": (C.c_int, [c_ctx_p, C.c_void_p]),
    "tetra_p2p_connect_ptrs": (C.c_int, [c_ctx_p, C.c_void_p]),
    "tetra_allgather_dibits": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "tetra_process_batch_allgather": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "tetra_p2p_status": (C.c_int, [c_ctx_p, C.POINTER(C.c_int32)]),
    "tetra_p2p_destroy": (C.c_int, [c_ctx_p]),
    "tetra_launch_count": (C.c_int64, [c_ctx_p]),
    "tetra_enable_kernel_timing": (C.c_int, [c_ctx_p, C.c_int]),
    "tetra_kernel_time_ms": (C.c_double, [c_ctx_p, C.POINTER(C.c_int32)]),
    "tetra_last_phase_ms": (C.c_int, [c_ctx_p, C.c_void_p]),
    "tetra_find_sync": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "tetra_sync_cascade": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "tetra_filter_signal": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p]),
    "tetra_frequency_shift": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p]),
    "tetra_extract_symbols": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "tetra_demodulate_dqpsk": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]),
    "tetra_resample": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "tetra_stft_db": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                C.POINTER(C.c_int64)]),
    "tetra_stft_db_f64": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                    C.POINTER(C.c_int64)]),
    "tetra_edge_corrections": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "tetra_design_butter4": (C.c_int, [C.c_double, C.c_void_p, C.c_void_p]),
    "tetra_design_cheby1_sos8": (C.c_int, [C.c_double, C.c_double, C.c_void_p]),
}

IPC_HANDLE_BYTES = 64            # TETRA_IPC_HANDLE_BYTES
SURVEY_FIELDS = 16               # TETRA_SURVEY_FIELDS
EXPORTS = tuple(_SIGS)


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the sources are newer) and type the shared library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    alt = os.environ.get("TETRA_B200_LIB")               # an explicitly built variant of the library (A/B measurements)
    if alt:
        path = alt
    elif _build.is_stale():
        try:
            path = _build.build()
        except Exception as e:
            if not os.path.exists(path):
                raise
            import logging
            logging.getLogger("tetraear.signal.processor").warning(
                "libtetra_b200.so is older than its sources and could not be rebuilt (%s); loading the stale library", e)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class TetraError(RuntimeError):
    pass


def check(lib, ctx, rc: int, what: str) -> int:
    if rc < 0:
        msg = lib.tetra_last_error(ctx)
        raise TetraError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
    return rc
