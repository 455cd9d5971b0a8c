"""ctypes binding of include/srhep.h.  No torch types cross this boundary: device pointers
travel as integers (``tensor.data_ptr()``), streams as ``cudaStream_t`` handles."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .config import SrDimsC

_LIB: Optional[C.CDLL] = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsrhep.so")
# the bounds-asserting debug build (nvcc -DSRHEP_BOUNDS: every global-memory index of the hot kernels is checked against the extent of its
# buffer and traps with a message); selected with SRHEP_LIB_VARIANT=bounds in the environment of the process (tests/test_gpu_bounds.py)
LIB_PATH_BOUNDS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsrhep_bounds.so")

PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
METHODS = {"euler": 0, "midpoint": 1, "rk4": 2, "dopri5": 3}
CATEGORIES = ("embed", "adaln", "feat0", "ln", "qkv", "attn", "out", "mlp1", "mlp2", "head", "chain")

# every symbol include/srhep.h declares
EXPORTS = (
    "srhep_version", "srhep_weight_count", "srhep_create", "srhep_destroy", "srhep_last_error",
    "srhep_set_pass_tokens", "srhep_set_use_graph", "srhep_bind_events", "srhep_velocity",
    "srhep_sample", "srhep_sample_dopri5", "srhep_set_debug", "srhep_get_tap", "srhep_launch_count",
    "srhep_profile",
)


PFLOW_EXPORTS = ("pflow_weight_count", "pflow_create", "pflow_destroy", "pflow_last_error", "pflow_forward", "pflow_launch_count")


class PflowDimsC(C.Structure):
    """Mirror of ``PflowDims`` in include/pflow.h."""
    _fields_ = [("h_dim", C.c_int32), ("heads", C.c_int32), ("enc_layers", C.c_int32), ("kin_layers", C.c_int32),
                ("layer_emb_dim", C.c_int32), ("max_particles", C.c_int32), ("part_emb_dim", C.c_int32),
                ("card_n_hidden", C.c_int32), ("card_hidden", C.c_int32 * 4), ("card_out", C.c_int32)]


class PflowVarTransformC(C.Structure):
    """Mirror of ``PflowVarTransform`` in include/pflow.h."""
    _fields_ = [("trans", C.c_int32), ("m", C.c_float), ("scale", C.c_int32), ("mean", C.c_float), ("std", C.c_float),
                ("min", C.c_float), ("max", C.c_float), ("lo", C.c_float), ("hi", C.c_float)]


POST_EXPORTS = ("srpost_ensemble_unscale", "srpost_select_cells", "srpost_last_error")


class SrpostTargetTransformC(C.Structure):
    """Mirror of ``SrpostTargetTransform`` in include/srhep_post.h."""
    _fields_ = [("standard", C.c_int32), ("mean", C.c_float), ("std", C.c_float), ("alpha", C.c_float), ("f", C.c_float)]


class SrpostPflowOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw", "layer")]


class PflowCells(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw", "layer")]


class SrhepCond(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("eta", "cosphi", "sinphi", "e_proxy", "layer")]


def load() -> C.CDLL:
    """Loads libsrhep.so.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    variant = os.environ.get("SRHEP_LIB_VARIANT")          # "bounds": the index-asserting build; any other name: libsrhep_<name>.so next to it (A/B builds)
    path = LIB_PATH_BOUNDS if variant == "bounds" else (os.path.join(os.path.dirname(LIB_PATH), f"libsrhep_{variant}.so") if variant else LIB_PATH)
    if not os.path.isfile(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m superresolutionhep_b200.build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback path.")
    lib = C.CDLL(path)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    lib.srhep_version.restype = C.c_char_p
    lib.srhep_version.argtypes = []
    lib.srhep_weight_count.restype = C.c_size_t
    lib.srhep_weight_count.argtypes = [C.POINTER(SrDimsC)]
    lib.srhep_create.restype = C.c_int
    lib.srhep_create.argtypes = [C.c_int, C.POINTER(SrDimsC), vp, C.c_size_t, C.c_int, C.POINTER(vp)]
    lib.srhep_destroy.restype = C.c_int
    lib.srhep_destroy.argtypes = [vp]
    lib.srhep_last_error.restype = C.c_char_p
    lib.srhep_last_error.argtypes = [vp]
    lib.srhep_set_pass_tokens.restype = C.c_int
    lib.srhep_set_pass_tokens.argtypes = [vp, i64]
    lib.srhep_set_use_graph.restype = C.c_int
    lib.srhep_set_use_graph.argtypes = [vp, C.c_int]
    lib.srhep_set_debug.restype = C.c_int
    lib.srhep_set_debug.argtypes = [vp, C.c_int]
    lib.srhep_bind_events.restype = C.c_int
    lib.srhep_bind_events.argtypes = [vp, C.POINTER(SrhepCond), vp, i32, vp]
    lib.srhep_velocity.restype = C.c_int
    lib.srhep_velocity.argtypes = [vp, vp, vp, vp, vp]
    lib.srhep_sample.restype = C.c_int
    lib.srhep_sample.argtypes = [vp, vp, vp, i32, i32, i32, vp, C.POINTER(i32), vp]
    lib.srhep_sample_dopri5.restype = C.c_int
    lib.srhep_sample_dopri5.argtypes = [vp, vp, vp, i32, f32, f32, i32, vp, C.POINTER(i32), vp]
    lib.srhep_get_tap.restype = C.c_int
    lib.srhep_get_tap.argtypes = [vp, C.c_char_p, vp, C.c_size_t, vp]
    lib.srhep_profile.restype = C.c_int
    lib.srhep_profile.argtypes = [vp, vp, f32, vp, C.POINTER(f32), C.POINTER(i32), vp]
    lib.srhep_launch_count.restype = u64
    lib.srhep_launch_count.argtypes = [vp]
    lib.pflow_weight_count.restype = C.c_size_t
    lib.pflow_weight_count.argtypes = [C.POINTER(PflowDimsC)]
    lib.pflow_create.restype = C.c_int
    lib.pflow_create.argtypes = [C.c_int, C.POINTER(PflowDimsC), vp, C.c_size_t, C.POINTER(PflowVarTransformC), C.POINTER(vp)]
    lib.pflow_destroy.restype = C.c_int
    lib.pflow_destroy.argtypes = [vp]
    lib.pflow_last_error.restype = C.c_char_p
    lib.pflow_last_error.argtypes = [vp]
    lib.pflow_forward.restype = C.c_int
    lib.pflow_forward.argtypes = [vp, C.POINTER(PflowCells), vp, i32, vp, vp, vp, vp, vp, vp]
    lib.pflow_launch_count.restype = u64
    lib.pflow_launch_count.argtypes = [vp]
    lib.srpost_ensemble_unscale.restype = C.c_int
    lib.srpost_ensemble_unscale.argtypes = [vp, i32, i32, i64, vp, C.POINTER(SrpostTargetTransformC), f32, vp, vp, vp, vp]
    lib.srpost_select_cells.restype = C.c_int
    lib.srpost_select_cells.argtypes = [vp, vp, vp, vp, vp, i32, f32, C.POINTER(PflowVarTransformC), C.POINTER(PflowVarTransformC),
                                        C.POINTER(SrpostPflowOut), vp, vp]
    lib.srpost_last_error.restype = C.c_char_p
    lib.srpost_last_error.argtypes = []
    _LIB = lib
    return lib


def check_post(lib: C.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.srpost_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def check_pflow(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.pflow_last_error(handle)
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def check(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.srhep_last_error(handle)
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
