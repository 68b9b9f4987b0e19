"""ctypes binding of libmquant_b200.so (the C ABI declared in include/mquant_b200.h).

There is no CPU fallback: if the library cannot be loaded, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (MQUANT_B200_LIB: an experimental build of the same library, tools/ A/B runs only)
LIB_PATH = os.environ.get("MQUANT_B200_LIB") or os.path.join(_HERE, "libmquant_b200.so")

MQ_FLAG_SPLIT = 1
MQ_FLAG_FP32_HAD = 2
MQ_FLAG_FAST_HAD = 4

_lib = None

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_u = C.c_uint
_sz = C.c_size_t

# name -> (restype, argtypes); must list EVERY symbol of include/mquant_b200.h
SIGNATURES = {
    "mq_version": (_i, []),
    "mq_last_error": (C.c_char_p, []),
    "mq_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "mq_act_quant_i8": (_i, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i, _i, _u, _p]),
    "mq_act_fake_quant_f16": (_i, [_p, _i64, _p, _i64, _p, _p, _i, _i, _i64, _i, _p]),
    "mq_rmsn_quant_i8": (_i, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i, _i, C.c_float, C.c_float, _p]),
    "mq_w4a8_gemm_gated_ws": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, C.c_size_t, _p]),
    "mq_silu_mul_interleaved_f16": (_i, [_p, _i64, _p, _i64, _i64, _i64, _p]),
    "mq_gptq_hessian_update": (_i, [_p, _i64, _i64, _i, _p, C.c_float, C.c_float, _p]),
    "mq_gptq_block_f32": (_i, [_p, _i64, _p, _i64, _p, C.c_float, _p, _p, _p, _i, _i, _p]),
    "mq_act_quant_i8_dynamic": (_i, [_p, _i64, _p, _p, _i64, _i, _i, C.c_float, _p]),
    "mq_hadamard_quant_i8": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _u, _p]),
    "mq_hadamard_transform_f16": (_i, [_p, _i64, _p, _i64, _i64, _i, C.c_float, _p]),
    "mq_observer_minmax": (_i, [_p, _i64, _i64, _i, _i, _p, _p]),
    "mq_observer_minmax_msq": (_i, [_p, _i64, _i64, _i, _p, _p, _p, _p, _p]),
    "mq_observer_percentile_workspace": (_sz, []),
    "mq_observer_percentile": (_i, [_p, _i64, _i64, _i, C.c_double, _p, _p, _sz, _p]),
    "mq_weight_quant_rtn": (_i, [_p, _i64, _p, _p, _p, _i64, _i, _i, _p]),
    "mq_w4_packed_bytes": (_sz, [_i64, _i64]),
    "mq_pack_w4": (_i, [_p, _i64, _p, _i64, _i64, _u, _p]),
    "mq_repack_w4_from_ref": (_i, [_p, _p, _i64, _i64, _p]),
    "mq_unpack_w4": (_i, [_p, _p, _i64, _i64, _i64, _p]),
    "mq_w4a8_gemm": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p]),
    "mq_w4a8_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mq_w4a8_gemm_plan": (_i, [_i64, _i64, _i64, _sz]),
    "mq_debug_force_gemm_kernel": (_i, [_i]),
    "mq_debug_force_hadamard_kernel": (_i, [_i]),
    "mq_w4a8_gemm_ws": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _sz, _p]),
    "mq_w4a8_gemm_simt": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p]),
}


class MQuantLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MQuantLibraryError(
                f"{LIB_PATH} not found: build it with `python -m mquant_b200.build` "
                "(mquant_b200 has no CPU fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mq_last_error().decode("utf-8", "replace")
        raise MQuantLibraryError(f"{what} failed (status {rc}): {msg}")
