"""ctypes binding of libsinddm_b200.so (the C ABI declared in include/sinddm_b200.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing or a call fails,
the caller gets an exception -- never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libsinddm_b200.so"

MATH_FP32 = 0
MATH_TF32 = 1
MATH_TF32X3 = 2
NUM_PARAMS = 52

_vp = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_sz = C.c_size_t
_f = C.c_float


class ConvDesc(C.Structure):
    """struct sinddm_conv_desc"""
    _fields_ = [
        ("B", _i), ("H", _i), ("W", _i),
        ("inp", _vp), ("Cin", _i),
        ("w", _vp), ("ntaps", _i),
        ("in_res", _vp), ("Cres", _i),
        ("w_res", _vp), ("N", _i),
        ("bias", _vp), ("res_add", _vp), ("x3", _vp), ("w_res3", _vp),
        ("gelu", _i),
        ("out_pre", _vp), ("dgelu_z", _vp),
        ("w_final", _vp), ("b_final", _vp), ("out_final", _vp),
        ("round_tf32", _i),
        ("out", _vp),
    ]


class DdpmStepDesc(C.Structure):
    """struct sinddm_ddpm_step_desc"""
    _fields_ = [
        ("x_t", _vp), ("eps", _vp), ("x_tilde", _vp), ("noise", _vp), ("t", _vp), ("out", _vp),
        ("B", _i), ("per_sample", _ll), ("reblur_mode", _i), ("clip_denoised", _i), ("omega", _f),
        ("sqrt_recip_alphas_cumprod", _vp), ("sqrt_recipm1_alphas_cumprod", _vp),
        ("posterior_mean_coef1", _vp), ("posterior_mean_coef2", _vp), ("posterior_log_variance_clipped", _vp),
        ("alphas_cumprod", _vp), ("sqrt_alphas_cumprod", _vp), ("sqrt_one_minus_alphas_cumprod", _vp),
        ("gammas", _vp),
    ]


FUSED_MAX_WORLD = 8


class FusedStepDesc(C.Structure):
    """struct sinddm_fused_step_desc"""
    _fields_ = [
        ("world", _i), ("rank", _i), ("n", _ll),
        ("grads", _vp * FUSED_MAX_WORLD), ("flags", _vp * FUSED_MAX_WORLD),
        ("epoch", C.c_uint32),
        ("param", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("ema", _vp),
        ("lr", _f), ("beta1", _f), ("beta2", _f), ("eps", _f),
        ("step", _ll), ("ema_mode", _i), ("ema_beta", _f), ("wait_ns", _vp), ("mc_grads", _vp),
    ]


# name -> (restype, argtypes); must list every function include/sinddm_b200.h declares (tests check it)
SIGNATURES = {
    "sinddm_init": (_i, [_i]),
    "sinddm_last_error": (C.c_char_p, []),
    "sinddm_abi_version": (_i, []),
    "sinddm_launch_count": (C.c_ulonglong, []),
    "sinddm_profile_enable": (None, [_i]),
    "sinddm_profile_collect": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i)]),
    "sinddm_plan_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "sinddm_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp, _sz]),
    "sinddm_plan_destroy": (None, [_vp]),
    "sinddm_net_pack_weights": (_i, [_vp, _vp, _vp]),
    "sinddm_net_forward": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "sinddm_net_backward": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sinddm_conv_forward": (_i, [C.POINTER(ConvDesc), _i, _vp]),
    "sinddm_conv_epilogue_flavour": (_i, [C.POINTER(ConvDesc)]),
    "sinddm_pack_conv_weights": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sinddm_split3": (_i, [_vp, _ll, _i, _vp, _i, _vp]),
    "sinddm_conv_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "sinddm_conv_wgrad": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "sinddm_dw5x5": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sinddm_dw5x5_wgrad_workspace_bytes": (_sz, [_i, _i, _i]),
    "sinddm_dw5x5_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "sinddm_colsum_workspace_bytes": (_sz, [_i]),
    "sinddm_colsum": (_i, [_vp, _ll, _i, _vp, _vp, _sz, _vp]),
    "sinddm_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sinddm_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sinddm_qsample_mix": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _vp]),
    "sinddm_philox_normal_rows": (_i, [_vp, _ll, _ll, _ll, C.c_ulonglong, C.c_ulonglong, _vp]),
    "sinddm_l1_loss_workspace_bytes": (_sz, []),
    "sinddm_l1_loss": (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _sz, _vp]),
    "sinddm_ddpm_step": (_i, [C.POINTER(DdpmStepDesc), _vp]),
    "sinddm_fused_step": (_i, [C.POINTER(FusedStepDesc), _vp]),
}


class SinddmError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None
_inited_devices = set()


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise SinddmError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -m sinddm_b200.build or "
                f"__graft_entry__.build()); sinddm_b200 has no CPU / PyTorch fallback")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.sinddm_abi_version() != 3:
            raise SinddmError("libsinddm_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sinddm_last_error()
        raise SinddmError(f"{what or 'sinddm call'} failed (status {rc}): {msg.decode() if msg else ''}")


def init(device_index: int) -> None:
    if device_index in _inited_devices:
        return
    lib = load()
    check(lib.sinddm_init(int(device_index)), "sinddm_init")
    _inited_devices.add(device_index)


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SinddmError("sinddm_b200 runs on CUDA (sm_100a) tensors only; got a CPU tensor -- "
                              "there is no CPU fallback")
