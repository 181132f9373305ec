"""ctypes binding of libarkmpc_b200.so — the C ABI declared in include/arkmpc_b200.h.

There is NO CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libarkmpc_b200.so")

FIELD_IDS = {"bn254_fr": 0, "curve25519_fr": 1}
CURVE_IDS = {"bn254_g1": 0, "curve25519_edwards": 1}

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_UNSUPPORTED, ERR_NCCL = -1, -2, -3, -4, -5, -6


class ArkMpcError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: status {status} ({_status_string(status)}){': ' + detail if detail else ''}")


_lib: Optional[C.CDLL] = None
_lock = threading.Lock()

_vp, _sz, _i, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64

# name -> argtypes (restype int unless listed in _RESTYPES)
_PROTOS = {
    "arkmpc_abi_version": [],
    "arkmpc_status_string": [_i],
    "arkmpc_device_count": [C.POINTER(_i)],
    "arkmpc_ctx_create": [_i, C.POINTER(_vp)],
    "arkmpc_ctx_destroy": [_vp],
    "arkmpc_ctx_set_stream": [_vp, _vp],
    "arkmpc_ctx_reset_stream": [_vp],
    "arkmpc_ctx_hint_independent": [_vp],
    "arkmpc_ctx_get_stream": [_vp],
    "arkmpc_ctx_device": [_vp],
    "arkmpc_ctx_sync": [_vp],
    "arkmpc_ctx_sm_count": [_vp],
    "arkmpc_last_error": [_vp],
    "arkmpc_ctx_launch_count": [_vp],
    "arkmpc_malloc": [_vp, _sz, C.POINTER(_vp)],
    "arkmpc_free": [_vp, _vp],
    "arkmpc_mem_trim": [_vp],
    "arkmpc_mem_cached_bytes": [_vp, C.POINTER(_sz)],
    "arkmpc_host_alloc": [_vp, _sz, C.POINTER(_vp)],
    "arkmpc_host_free": [_vp, _vp],
    "arkmpc_memcpy_h2d": [_vp, _vp, _vp, _sz],
    "arkmpc_memcpy_d2h": [_vp, _vp, _vp, _sz],
    "arkmpc_memcpy_d2d": [_vp, _vp, _vp, _sz],
    "arkmpc_share_unzip": [_vp, _sz, _vp, _vp, _vp],
    "arkmpc_share_zip": [_vp, _sz, _vp, _vp, _vp],
    "arkmpc_fr_beaver_mask": [_vp, _i, _sz] + [_vp] * 6,
    "arkmpc_fr_beaver_recombine": [_vp, _i, _i, _vp, _sz] + [_vp] * 14,
    "arkmpc_fr_beaver_recombine_sum": [_vp, _i, _i, _vp, _sz] + [_vp] * 12,
    "arkmpc_fr_add": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_fr_sub": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_fr_mul": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_fr_neg": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_scale": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_fr_share_add": [_vp, _i, _sz] + [_vp] * 6,
    "arkmpc_fr_share_sub": [_vp, _i, _sz] + [_vp] * 6,
    "arkmpc_fr_share_neg": [_vp, _i, _sz] + [_vp] * 4,
    "arkmpc_fr_share_add_public": [_vp, _i, _i, _vp, _sz] + [_vp] * 5,
    "arkmpc_fr_share_sub_public": [_vp, _i, _i, _vp, _sz] + [_vp] * 5,
    "arkmpc_fr_share_mul_public": [_vp, _i, _sz] + [_vp] * 5,
    "arkmpc_fr_mac_check": [_vp, _i, _vp, _sz, _vp, _vp, _vp],
    "arkmpc_fr_sum_is_zero": [_vp, _i, _sz, _vp, _vp, C.POINTER(_i)],
    "arkmpc_fr_validate": [_vp, _i, _sz, _vp, C.POINTER(_i)],
    "arkmpc_fr_to_bytes_be": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_share_sum": [_vp, _i, _sz, _vp, _vp, _vp, _vp],
    "arkmpc_fr_sum": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_batch_inverse": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_fft": [_vp, _i, _i, _i, _vp, _vp],
    "arkmpc_fr_share_fft": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "arkmpc_fr_to_mont": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_from_mont": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_random": [_vp, _i, _u64, _u64, _sz, _vp],
    "arkmpc_ipc_export": [_vp, _vp, _vp],
    "arkmpc_ipc_import": [_vp, _vp, C.POINTER(_vp)],
    "arkmpc_ipc_release": [_vp, _vp],
    "arkmpc_fr_beaver_recombine_gather": [_vp, _i, _i, _vp, _sz] + [_vp] * 12 + [_i, _i, _vp, _vp],
    "arkmpc_mc_supported": [_vp, C.POINTER(_i)],
    "arkmpc_mc_open": [_vp, _sz, _i, _i, _i, _i, C.POINTER(_vp)],
    "arkmpc_mc_export": [_vp, C.POINTER(_i), C.POINTER(_i)],
    "arkmpc_mc_bind": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
    "arkmpc_mc_close": [_vp],
    "arkmpc_fr_beaver_recombine_gather_mc": [_vp, _i, _i, _vp, _sz] + [_vp] * 12 + [_i, _i, _vp, _vp],
    "arkmpc_mc_allgather_rows": [_vp, _sz, _i, _vp, _vp],
    "arkmpc_nccl_unique_id": [_vp],
    "arkmpc_nccl_init": [_vp, _i, _i, _vp],
    "arkmpc_nccl_destroy": [_vp],
    "arkmpc_allgather_open": [_vp, _sz, _vp, _vp, _vp, _vp],
    "arkmpc_point_bytes": [_i],
    "arkmpc_pt_add": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_sub": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_neg": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_pt_share_add_public": [_vp, _i, _i, _vp, _sz, _vp, _vp, _vp],
    "arkmpc_pt_share_sub_public": [_vp, _i, _i, _vp, _sz, _vp, _vp, _vp],
    "arkmpc_pt_mul": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_share_mul_public": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_mul_authenticated": [_vp, _i, _sz, _vp, _vp, _vp, _vp],
    "arkmpc_pt_mul_generator": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_mul_generator_public": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_pt_beaver_mask": [_vp, _i, _sz] + [_vp] * 6,
    "arkmpc_pt_beaver_recombine": [_vp, _i, _i, _vp, _sz] + [_vp] * 13,
    "arkmpc_pt_mac_check": [_vp, _i, _vp, _sz, _vp, _vp, _vp],
    "arkmpc_pt_sum_is_identity": [_vp, _i, _sz, _vp, _vp, C.POINTER(_i)],
    "arkmpc_pt_validate": [_vp, _i, _sz, _vp, C.POINTER(_i)],
    "arkmpc_pt_sum": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_pt_share_sum": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_pt_msm": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_msm_authenticated": [_vp, _i, _sz, _vp, _vp, _vp, _vp],
    "arkmpc_pt_share_split": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_share_join": [_vp, _i, _sz, _vp, _vp, _vp],
    "arkmpc_pt_normalize": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_pt_from_affine": [_vp, _i, _sz, _vp, _vp],
    "arkmpc_fr_batch_mul_begin_host": [_vp, _i, _i, _vp, _sz] + [_vp] * 6 + [C.POINTER(_vp)],
    "arkmpc_fr_batch_mul_begin_host_shares": [_vp, _i, _i, _vp, _sz] + [_vp] * 6 + [C.POINTER(_vp)],
    "arkmpc_fr_batch_mul_finish_host": [_vp, _vp, _vp, _vp],
    "arkmpc_fr_batch_mul_abort": [_vp],
    "arkmpc_fr_batch_mul_host_bytes": [_vp, _sz, _i, C.POINTER(_u64), C.POINTER(_u64)],
}
_RESTYPES = {
    "arkmpc_status_string": C.c_char_p,
    "arkmpc_last_error": C.c_char_p,
    "arkmpc_ctx_get_stream": _vp,
    "arkmpc_ctx_launch_count": _u64,
    "arkmpc_point_bytes": _sz,
}

# every symbol include/arkmpc_b200.h declares (checked by tests/test_abi.py against the header text)
EXPORTED_SYMBOLS = tuple(_PROTOS)


def load() -> C.CDLL:
    """Load the native library; raise (never fall back) if it is absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ArkMpcError(ERR_UNSUPPORTED, "load",
                                  f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, argtypes in _PROTOS.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = _RESTYPES.get(name, _i)
            _lib = lib
    return _lib


_STATUS_TEXT = {0: "ok", -1: "invalid argument", -2: "CUDA error", -3: "no usable CUDA device", -4: "out of device memory",
                -5: "unsupported in this build", -6: "NCCL error"}


def _status_string(status: int) -> str:
    # never calls load(): this runs while reporting that the library is MISSING (and load() holds a non-reentrant lock)
    if _lib is not None:
        try:
            return _lib.arkmpc_status_string(status).decode()
        except Exception:  # pragma: no cover
            pass
    return _STATUS_TEXT.get(status, "unknown status")


def check(status: int, where: str, ctx: Optional[int] = None) -> None:
    if status != OK:
        detail = ""
        if ctx:
            detail = (load().arkmpc_last_error(ctx) or b"").decode()
        raise ArkMpcError(status, where, detail)


def device_count() -> int:
    n = _i(0)
    load().arkmpc_device_count(C.byref(n))
    return n.value
