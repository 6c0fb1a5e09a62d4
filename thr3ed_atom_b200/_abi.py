"""ctypes mirror of ``include/r3d_b200.h`` and the loader of ``_lib/libr3d_b200.so``.

The library is the product: there is no eager / CPU fallback.  ``lib()`` raises if the shared
object has not been built (``python -m thr3ed_atom_b200.build``), and every entry point raises
``RuntimeError`` with the library's message when it returns a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

# $R3D_LIB_PATH selects another build of the same library (tuning experiments: other -D flags); default = the in-tree build
LIB_PATH = Path(os.environ.get("R3D_LIB_PATH") or (Path(__file__).resolve().parent / "_lib" / "libr3d_b200.so"))
ABI_VERSION = 5

# enums of r3d_b200.h
PRE_IDENTITY, PRE_ABS = 0, 1
POST_IDENTITY, POST_RELU, POST_SOFTPLUS = 0, 1, 2
FLAG_PERTURB, FLAG_WHITE_BKGD, FLAG_DIFFUSE, FLAG_OPTIMIZED_SAMPLING = 1, 2, 4, 8

c_float_p = C.POINTER(C.c_float)


class R3dGrid(C.Structure):
    _fields_ = [
        ("densities", C.c_void_p),
        ("features", C.c_void_p),
        ("dims", C.c_int32 * 3),
        ("sh_degree", C.c_int32),
        ("num_features", C.c_int32),
        ("feature_stride", C.c_int32),
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("norm_scale", C.c_float * 3),
        ("norm_bias", C.c_float * 3),
        ("density_scale", C.c_float),
        ("density_pre", C.c_int32),
        ("density_post", C.c_int32),
        ("density_quads", C.c_void_p),
    ]


class R3dCamera(C.Structure):
    _fields_ = [
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("focal", C.c_float),
        ("rotation", C.c_float * 9),
        ("translation", C.c_float * 3),
    ]


class R3dViewSet(C.Structure):
    _fields_ = [
        ("rotations", C.c_void_p),
        ("translations", C.c_void_p),
        ("images", C.c_void_p),
        ("num_views", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("focal", C.c_float),
    ]


class R3dRays(C.Structure):
    _fields_ = [
        ("origins", C.c_void_p),
        ("directions", C.c_void_p),
        ("bounds", C.c_void_p),
        ("camera", C.POINTER(R3dCamera)),
        ("num_rays", C.c_int64),
        ("tile_width", C.c_int32),
        ("tile_height", C.c_int32),
    ]


class R3dRenderConfig(C.Structure):
    _fields_ = [
        ("num_samples", C.c_int32),
        ("near", C.c_float),
        ("far", C.c_float),
        ("flags", C.c_uint32),
        ("jitter", C.c_void_p),
        ("rng_seed", C.c_uint64),
        ("variant", C.c_int32),
    ]


class R3dRenderOut(C.Structure):
    _fields_ = [("colour", C.c_void_p), ("depth", C.c_void_p), ("acc", C.c_void_p), ("disparity", C.c_void_p), ("sample_cache", C.c_void_p),
                ("colour_diffuse", C.c_void_p), ("sample_cache_diffuse", C.c_void_p), ("sample_mask", C.c_void_p)]


class R3dRenderOutGrad(C.Structure):
    _fields_ = [("colour", C.c_void_p), ("depth", C.c_void_p), ("acc", C.c_void_p), ("disparity", C.c_void_p), ("colour_diffuse", C.c_void_p)]


class R3dGridGrad(C.Structure):
    _fields_ = [("densities", C.c_void_p), ("features", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/r3d_b200.h declares
SIGNATURES = {
    "r3d_abi_version": (C.c_int, []),
    "r3d_has_ab_variants": (C.c_int, []),
    "r3d_last_error": (C.c_char_p, []),
    "r3d_sample_mask_words": (C.c_int64, [C.POINTER(R3dRays)]),
    "r3d_render_fwd": (C.c_int, [C.POINTER(R3dGrid), C.POINTER(R3dRays), C.POINTER(R3dRenderConfig), C.POINTER(R3dRenderOut), C.c_void_p]),
    "r3d_render_bwd": (
        C.c_int,
        [C.POINTER(R3dGrid), C.POINTER(R3dRays), C.POINTER(R3dRenderConfig), C.POINTER(R3dRenderOut),
         C.POINTER(R3dRenderOutGrad), C.POINTER(R3dGridGrad), C.c_void_p],
    ),
    "r3d_cast_rays": (C.c_int, [C.POINTER(R3dCamera), C.c_void_p, C.c_void_p, C.c_void_p]),
    "r3d_sample_ray_batch": (C.c_int, [C.POINTER(R3dViewSet), C.c_int64, C.c_int32, C.c_int32, C.c_uint64] + [C.c_void_p] * 5),
    "r3d_grid_lookup_fwd": (C.c_int, [C.POINTER(R3dGrid), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "r3d_grid_lookup_bwd": (C.c_int, [C.POINTER(R3dGrid), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(R3dGridGrad), C.c_void_p]),
    "r3d_mark_touched_voxels": (C.c_int, [C.POINTER(R3dGrid), C.POINTER(R3dRays), C.POINTER(R3dRenderConfig), C.c_void_p, C.c_void_p]),
    "r3d_sample_statistics": (C.c_int, [C.POINTER(R3dGrid), C.POINTER(R3dRays), C.POINTER(R3dRenderConfig), C.c_void_p, C.c_void_p]),
    "r3d_density_quad_floats": (C.c_int64, [C.POINTER(C.c_int32 * 3)]),
    "r3d_build_density_quads": (C.c_int, [C.POINTER(R3dGrid), C.c_void_p, C.c_void_p]),
    "r3d_multimem_all_reduce": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "r3d_multimem_shard_floats": (C.c_int64, [C.c_int64, C.c_int32]),
    "r3d_multimem_adam_step": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32] + [C.c_float] * 7 + [C.c_int32, C.c_void_p],
    ),
    "r3d_peer_adam_step": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32] + [C.c_float] * 7 + [C.c_int32, C.c_void_p],
    ),
    "r3d_adam_step": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_float] * 7 + [C.c_void_p],
    ),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded C-ABI library (loads on first use; raises loudly if it is missing)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                "Run `python -m thr3ed_atom_b200.build` (needs nvcc). There is no CPU fallback."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError => header/library mismatch
            fn.restype, fn.argtypes = restype, argtypes
        got = handle.r3d_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libr3d_b200.so ABI version {got} != binding version {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().r3d_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")
