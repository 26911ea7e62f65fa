"""ctypes binding of libshmfast.so (the C ABI declared in include/shmfast.h).

No CPU fallback: if the library is missing or no sm_100 device is present, every compute entry point
raises (ShmfastError / RuntimeError) instead of computing somewhere else.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

SHM_MAX_D = 16
SHM_MAX_L = 2

ENGINE_AUTO, ENGINE_FP32, ENGINE_TC_BF16X3 = 0, 1, 2
CNN_4DOF, CNN_OPENLAB = 0, 1
GEMM_SIMT, GEMM_TC_F16X3, GEMM_TC_BF16X3 = 0, 1, 2

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


class ShmfastError(RuntimeError):
    pass


class WindowSrc(C.Structure):
    _fields_ = [
        ("base", _vp), ("win_stride", C.c_int64), ("row_stride", C.c_int64), ("T", C.c_int32), ("D", C.c_int32),
        ("chan", C.c_int32 * SHM_MAX_D), ("normalize", C.c_int32), ("nan_to_zero", C.c_int32), ("clip", C.c_float),
        ("mean", C.c_float * SHM_MAX_D), ("std", C.c_float * SHM_MAX_D),
    ]


class VaeCfg(C.Structure):
    _fields_ = [("D", C.c_int32), ("H", C.c_int32), ("Z", C.c_int32), ("L", C.c_int32), ("has_ln", C.c_int32),
                ("ln_eps", C.c_float), ("engine", C.c_int32)]


class VaeWeights(C.Structure):
    _fields_ = [
        ("enc_w_ih", _vp * SHM_MAX_L), ("enc_w_hh", _vp * SHM_MAX_L), ("enc_b_ih", _vp * SHM_MAX_L), ("enc_b_hh", _vp * SHM_MAX_L),
        ("ln_w", _vp), ("ln_b", _vp), ("fc_mu_w", _vp), ("fc_mu_b", _vp), ("fc_lv_w", _vp), ("fc_lv_b", _vp),
        ("l2h_w", _vp), ("l2h_b", _vp),
        ("dec_w_ih", _vp * SHM_MAX_L), ("dec_w_hh", _vp * SHM_MAX_L), ("dec_b_ih", _vp * SHM_MAX_L), ("dec_b_hh", _vp * SHM_MAX_L),
        ("out_w", _vp), ("out_b", _vp),
    ]


class Cnn4dofWeights(C.Structure):
    _fields_ = [("conv_w", _vp * 2), ("conv_b", _vp * 2), ("bn_w", _vp * 2), ("bn_b", _vp * 2), ("bn_mean", _vp * 2),
                ("bn_var", _vp * 2), ("fc1_w", _vp), ("fc1_b", _vp), ("fc2_w", _vp), ("fc2_b", _vp), ("bn_eps", C.c_float)]


class ExtractCfg(C.Structure):
    _fields_ = [("T", C.c_int32), ("stride", C.c_int32), ("ma_window", C.c_int32), ("struct_channel_mask", C.c_int32),
                ("obstruction_sentinel", C.c_double), ("raw_diff_th", C.c_double), ("raw_abs_th", C.c_double),
                ("clean_max_jump", C.c_double), ("clean_max_abs", C.c_double), ("raw_invalid_ratio_fault", C.c_float),
                ("flat_var_eps", C.c_float), ("force_range_for_flatline", C.c_float), ("allow_max", C.c_float)]


class CnnOlWeights(C.Structure):
    _fields_ = [("conv_w", _vp * 4), ("conv_b", _vp * 4), ("gn_w", _vp * 4), ("gn_b", _vp * 4), ("fc1_w", _vp),
                ("fc1_b", _vp), ("fc2_w", _vp), ("fc2_b", _vp), ("gn_eps", C.c_float)]


# every symbol include/shmfast.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "shm_strerror": (C.c_char_p, [C.c_int]),
    "shm_last_cuda_error": (C.c_char_p, []),
    "shm_version": (C.c_int, []),
    "shm_device_check": (C.c_int, [C.c_int]),
    "shm_window_normalize": (C.c_int, [C.POINTER(WindowSrc), _vp, C.c_int64, _vp, _vp]),
    "shm_vae_create": (C.c_int, [C.POINTER(_vp), C.POINTER(VaeCfg), C.POINTER(VaeWeights), C.c_int]),
    "shm_vae_update_weights": (C.c_int, [_vp, C.POINTER(VaeWeights), _vp]),
    "shm_vae_destroy": (C.c_int, [_vp]),
    "shm_vae_engine": (C.c_int, [_vp]),
    "shm_vae_get_cfg": (C.c_int, [_vp, C.POINTER(VaeCfg)]),
    "shm_hybrid4dof_workspace_bytes": (C.c_int64, [_vp, C.c_int64, C.c_int64]),
    "shm_hybrid4dof_score": (C.c_int, [_vp, _vp, C.POINTER(WindowSrc), C.c_int64, _vp, _vp, C.c_float, C.c_int64] + [_vp] * 10 +
                             [C.c_int64, _vp]),
    "shm_hybridol_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "shm_hybridol_score": (C.c_int, [_vp, _vp, C.POINTER(WindowSrc), C.POINTER(WindowSrc), C.c_int64, _vp, C.c_float, C.c_double,
                                     C.c_int64] + [_vp] * 10 + [C.c_int64, _vp]),
    "shm_scatter_flagged_4dof": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "shm_scatter_flagged_openlab": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_double, C.c_int64, _vp, _vp, _vp, _vp]),
    "shm_vae_debug_counters": (C.c_int, [_vp, _vp, C.c_int]),
    "shm_vae_score": (C.c_int, [_vp, C.POINTER(WindowSrc), _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shm_vae_rescore": (C.c_int, [_vp, C.POINTER(WindowSrc), _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "shm_vae_decode": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp]),
    "shm_vae_param_count": (C.c_int64, [C.POINTER(VaeCfg)]),
    "shm_vae_trainer_create": (C.c_int, [C.POINTER(_vp), C.POINTER(VaeCfg), C.c_int32, C.c_int32, C.c_int]),
    "shm_vae_trainer_destroy": (C.c_int, [_vp]),
    "shm_vae_train_forward": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp]),
    "shm_vae_train_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shm_vae_elbo_grad": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "shm_adam_clip_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "shm_gemm_f32": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _vp,
                               C.c_int32, C.c_int32, _vp]),
    "shm_train_set_tensor_cores": (C.c_int, [C.c_int]),
    "shm_adamw_clip_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "shm_cnn_param_count": (C.c_int64, [C.c_int]),
    "shm_cnn_trainer_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int32, C.c_int]),
    "shm_cnn_trainer_destroy": (C.c_int, [_vp]),
    "shm_cnn_train_forward": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, C.c_float, _vp, C.c_float, _vp, _vp]),
    "shm_cnn_train_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "shm_cnn_loss_grad": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_float, _vp, _vp, _vp]),
    "shm_compact_workspace_bytes": (C.c_int64, [C.c_int64]),
    "shm_compact": (C.c_int, [_vp, C.c_float, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "shm_cnn4dof_create": (C.c_int, [C.POINTER(_vp), C.POINTER(Cnn4dofWeights), C.c_int]),
    "shm_cnn4dof_update_weights": (C.c_int, [_vp, C.POINTER(Cnn4dofWeights), _vp]),
    "shm_cnn4dof_destroy": (C.c_int, [_vp]),
    "shm_cnn4dof_forward": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "shm_cnnol_create": (C.c_int, [C.POINTER(_vp), C.POINTER(CnnOlWeights), C.c_int]),
    "shm_cnnol_update_weights": (C.c_int, [_vp, C.POINTER(CnnOlWeights), _vp]),
    "shm_cnnol_destroy": (C.c_int, [_vp]),
    "shm_cnnol_forward": (C.c_int, [_vp, C.POINTER(WindowSrc), _vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "shm_cnnol_set_engine": (C.c_int, [_vp, C.c_int]),
    "shm_cnnol_engine": (C.c_int, [_vp]),
    "shm_openlab_extract_workspace_bytes": (C.c_int64, [C.c_int64]),
    "shm_openlab_extract": (C.c_int, [_vp, C.c_int64, C.POINTER(ExtractCfg)] + [_vp] * 15),
    "shm_stitch_segment_rmse": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _vp, _vp, _vp,
                                          C.c_int32, _vp, _vp, _vp]),
    "shm_percentile_workspace_bytes": (C.c_int64, [C.c_int64]),
    "shm_percentile": (C.c_int, [_vp, C.c_int64, C.c_double, _vp, _vp, _vp]),
}

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libshmfast.so"
_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library.  The library is always checked against the digest of csrc/ + shmfast.h
    (shmfast.build writes the stamp only next to a freshly linked .so): a stale binary is rebuilt with nvcc, or
    refused when `build_if_missing` is False -- never dlopen'ed against newer ctypes signatures."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _b
    if build_if_missing:
        _b.build()                                  # no-op when the stamp matches the sources
    elif not LIB_PATH.exists():
        raise ShmfastError(f"{LIB_PATH} is missing: run `python -m shmfast.build` (no CPU fallback exists)")
    elif not _b.is_fresh():
        raise ShmfastError(f"{LIB_PATH} does not match csrc/ and include/shmfast.h: run `python -m shmfast.build`")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.shm_strerror(rc).decode()
    if rc == -3:
        msg += ": " + lib.shm_last_cuda_error().decode()
    raise ShmfastError(f"{what or 'libshmfast'} failed ({rc}): {msg}")
