"""ctypes binding of libpfm_b200.so (the C ABI of include/pfm_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpfm_b200.so")

PFM_PREC_FP32, PFM_PREC_BF16 = 0, 1
PFM_SOLVER_EULER, PFM_SOLVER_MIDPOINT = 0, 1
PFM_LOSS_FM_OT, PFM_LOSS_CFM, PFM_LOSS_DROID = 0, 1, 2
PFM_STEP_PF_ODE, PFM_STEP_DDIM, PFM_STEP_EM = 1, 2, 3
PFM_ACT = {"ELU": 0, "Tanh": 1, "ReLU": 2, "LeakyReLU": 3, "SiLU": 4}
PFM_POST_MAX_FEATS = 16


class PfmError(RuntimeError):
    pass


class EpicCfgC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("feats", "input_dim", "hid", "latent", "layers", "t_dim", "t_local_cat",
                                         "t_global_cat", "global_cond_dim", "local_cond_dim")] + \
               [("sum_scale", C.c_float), ("neg_slope", C.c_float)]


class MlpCfgC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("features", "t_dim", "cond_dim", "n_linears", "act")]


class TfCfgC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "feats", "t_dim", "cond_dim", "add_time_to_input", "model_dim", "num_layers",
                                         "num_heads", "ctxt_out", "embd_hddn", "dense_hddn", "num_tokens")] + \
               [("neg_slope", C.c_float), ("ln_eps", C.c_float)]


_F = C.c_void_p     # device pointers travel as integers (tensor.data_ptr())
_SIGNATURES = {
    "pfm_version": (C.c_int, []),
    "pfm_last_error": (C.c_char_p, []),
    "pfm_epic_create": (C.c_int, [C.POINTER(EpicCfgC), C.c_int, C.POINTER(C.c_void_p)]),
    "pfm_epic_destroy": (None, [C.c_void_p]),
    "pfm_epic_num_linears": (C.c_int, [C.c_void_p]),
    "pfm_epic_linear_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pfm_epic_set_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "pfm_epic_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "pfm_epic_set_train_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "pfm_epic_debug_copy": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]),
    "pfm_epic_forward": (C.c_int, [C.c_void_p, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_void_p]),
    "pfm_epic_sample": (C.c_int, [C.c_void_p, _F, _F, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pfm_epic_set_params": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pfm_epic_param_grads": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]),
    "pfm_epic_grad_size": (C.c_longlong, [C.c_void_p]),
    "pfm_epic_grad_chunks": (C.c_int, [C.c_void_p]),
    "pfm_epic_grad_chunk_range": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "pfm_epic_stream_wait_grad_chunk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pfm_epic_loss_fwd_bwd": (C.c_int, [C.c_void_p, _F, _F, _F, _F, _F, _F, _F, _F, C.c_int, C.c_float, _F, _F, C.c_int,
                                        C.c_int, C.c_void_p]),
    "pfm_epic_forward_train": (C.c_int, [C.c_void_p, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_void_p]),
    "pfm_epic_backward": (C.c_int, [C.c_void_p, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_void_p]),
    "pfm_tf_create": (C.c_int, [C.POINTER(TfCfgC), C.c_int, C.POINTER(C.c_void_p)]),
    "pfm_tf_destroy": (None, [C.c_void_p]),
    "pfm_tf_num_params": (C.c_int, [C.c_void_p]),
    "pfm_tf_param_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pfm_tf_set_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "pfm_tf_forward": (C.c_int, [C.c_void_p, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_void_p]),
    "pfm_tf_sample": (C.c_int, [C.c_void_p, _F, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pfm_tf_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "pfm_tf_last_launches": (C.c_int, [C.c_void_p]),
    "pfm_tf_grad_size": (C.c_longlong, [C.c_void_p]),
    "pfm_tf_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p]),
    "pfm_tf_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfm_tf_loss_fwd_bwd": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7 + [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p]),
    "pfm_epic_sample_diffusion": (C.c_int, [C.c_void_p] + [_F] * 8 + [C.c_int] * 5 + [C.c_void_p]),
    "pfm_clip_adamw": (C.c_int, [_F, _F, _F, _F, C.c_longlong] + [C.c_float] * 6 + [C.c_int, _F, C.c_float, _F, C.c_void_p]),
    "pfm_postprocess": (C.c_int, [_F, _F, _F, C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                  C.c_int, C.c_void_p]),
    "pfm_ot_assign": (C.c_int, [_F, _F, C.c_int, C.c_int, C.c_int, _F, _F, C.c_void_p]),
    "pfm_ot_gather": (C.c_int, [_F] * 5 + [C.c_int] * 3 + [_F] * 3 + [C.c_void_p]),
    "pfm_mlp_create": (C.c_int, [C.POINTER(MlpCfgC), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                 C.POINTER(C.c_void_p)]),
    "pfm_mlp_destroy": (None, [C.c_void_p]),
    "pfm_mlp_linear_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pfm_mlp_set_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "pfm_mlp_forward": (C.c_int, [C.c_void_p, _F, C.c_int, _F, _F, _F, C.c_int, C.c_void_p]),
    "pfm_mlp_sample": (C.c_int, [C.c_void_p, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pfm_epic_last_launches": (C.c_int, [C.c_void_p]),
    "pfm_epic_last_groups": (C.c_int, [C.c_void_p]),
    "pfm_epic_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "pfm_epic_last_kernel_ms": (C.c_float, [C.c_void_p]),
}

_lib = None


def load():
    """dlopen the library and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfmError(
            f"{LIB_PATH} not found. Build it with `python -m particle_fm_b200.build` (needs nvcc). "
            "particle_fm_b200 has no CPU / PyTorch fallback by design.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc: int, what: str):
    if rc != 0:
        msg = load().pfm_last_error().decode("utf-8", "replace")
        raise PfmError(f"{what} failed (status {rc}): {msg}")
