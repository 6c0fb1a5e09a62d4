"""Deterministic input builders shared by the golden generator (``make_golden.py``), the oracle
tests and the GPU parity tests.

Inputs are produced with ``numpy.random.RandomState`` (a frozen legacy stream), never with torch's
RNG, so every machine rebuilds bit-identical inputs from the case description alone; only the
*outputs* of the reference are stored in ``tests/golden/*.npz``.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np

HOTDOG_RADIUS = 4.031128406524658  # reference thre3d_atom/data/tests/test_datasets.py:48-52
HOTDOG_NEAR, HOTDOG_FAR = 1.8, 6.6  # [2, 6] * (0.9, 1.1): tools/convert...py:15, data/datasets.py:243-244


def relu_field_density_scale(world_size) -> float:
    """reference rendering/volumetric/utils/misc.py:68-78"""
    diag = math.sqrt(sum(e * e for e in world_size))
    return ((math.sqrt(27.0) * 100.0) / diag) / 3


def spherical_pose(yaw_deg: float, pitch_deg: float, radius: float) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 ``Rz(yaw) @ Rx(pitch) @ Tz(radius)`` (reference utils/imaging_utils.py:146-191),
    evaluated with fp32 matrix products in the same association order."""
    yaw, pitch = yaw_deg / 180.0 * np.pi, pitch_deg / 180.0 * np.pi
    tz = np.eye(4, dtype=np.float32)
    tz[2, 3] = radius
    rx = np.array(
        [[1, 0, 0, 0], [0, np.cos(pitch), -np.sin(pitch), 0], [0, np.sin(pitch), np.cos(pitch), 0], [0, 0, 0, 1]],
        dtype=np.float32,
    )
    rz = np.array(
        [[np.cos(yaw), -np.sin(yaw), 0, 0], [np.sin(yaw), np.cos(yaw), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
        dtype=np.float32,
    )
    c2w = (rz @ (rx @ tz).astype(np.float32)).astype(np.float32)
    return np.ascontiguousarray(c2w[:3, :3]), np.ascontiguousarray(c2w[:3, 3:])


@dataclasses.dataclass
class Case:
    name: str
    dims: Tuple[int, int, int]
    sh_degree: int
    voxel_size: Tuple[float, float, float]
    location: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    density_pre: str = "identity"
    density_post: str = "relu"
    density_scale: float = 1.0
    density_range: Tuple[float, float] = (-1.0, 1.0)
    feature_range: Tuple[float, float] = (-1.0, 1.0)
    # camera / rays
    image_hw: Optional[Tuple[int, int]] = None  # pinhole camera if set
    focal: float = 0.0
    pose: Tuple[float, float, float] = (30.0, 60.0, HOTDOG_RADIUS)  # yaw, pitch, radius
    num_random_rays: int = 0  # else random incoherent rays
    # render config
    num_samples: int = 32
    near: float = HOTDOG_NEAR
    far: float = HOTDOG_FAR
    white_bkgd: bool = True
    diffuse: bool = False
    optimized_sampling: bool = False
    jitter: bool = False
    with_depth_acc_grads: bool = False
    seed: int = 42


def _hotdog_focal(width: int) -> float:
    return 1111.11 * width / 800.0


CASES: Dict[str, Case] = {
    c.name: c
    for c in [
        # BASELINE.json configs[0]: 32^3 SH-deg-0 grid, one 64x64 posed camera, 32 samples/ray
        Case("c1_32cube_deg0", (32, 32, 32), 0, (3 / 32,) * 3, image_hw=(64, 64), focal=_hotdog_focal(64),
             num_samples=32, density_scale=relu_field_density_scale((3, 3, 3)), with_depth_acc_grads=True),
        # the reference's 16^3 deg-2 ReLU-field model (modules/tests/test_volumetric_model.py:31-61)
        Case("deg2_16cube", (16, 16, 16), 2, (3 / 16,) * 3, image_hw=(40, 40), focal=_hotdog_focal(40),
             num_samples=64, density_scale=relu_field_density_scale((3, 3, 3))),
        Case("deg2_diffuse", (16, 16, 16), 2, (3 / 16,) * 3, image_hw=(24, 24), focal=_hotdog_focal(24),
             num_samples=48, density_scale=relu_field_density_scale((3, 3, 3)), diffuse=True, pose=(200.0, 35.0, 4.5)),
        # stratified jitter (explicit U[0,1) tensor) + ray/AABB-clipped sampling
        Case("deg2_jitter_optimized", (16, 16, 16), 2, (3 / 16,) * 3, image_hw=(32, 32), focal=_hotdog_focal(32),
             num_samples=40, density_scale=relu_field_density_scale((3, 3, 3)), jitter=True,
             optimized_sampling=True, pose=(115.0, 75.0, HOTDOG_RADIUS), with_depth_acc_grads=True),
        Case("deg2_jitter", (20, 20, 20), 2, (3 / 20,) * 3, image_hw=(32, 32), focal=_hotdog_focal(32),
             num_samples=64, density_scale=relu_field_density_scale((3, 3, 3)), jitter=True,
             white_bkgd=False, pose=(300.0, 50.0, HOTDOG_RADIUS)),
        # anisotropic, off-centre grid, softplus field, degree 1
        Case("deg1_aniso_softplus", (10, 12, 14), 1, (0.25, 0.2, 0.15), location=(0.2, -0.1, 0.3),
             density_post="softplus", density_scale=7.5, image_hw=(28, 36), focal=44.0,
             num_samples=56, near=1.5, far=6.0, pose=(75.0, 40.0, 3.7), white_bkgd=False),
        # "traditional" field: abs pre-activation, identity post, scale 1, degree 3
        Case("deg3_abs", (12, 12, 12), 3, (0.25,) * 3, density_pre="abs", density_post="identity",
             density_scale=1.0, density_range=(-6.0, 6.0), image_hw=(30, 30), focal=_hotdog_focal(30),
             num_samples=50, with_depth_acc_grads=True),
        # the reference's 2x2x2 colour cube (thre3d_reprs/tests/test_voxels.py:88-134), one oblique view
        Case("cube2", (2, 2, 2), 0, (2.0, 2.0, 2.0), density_range=(-10.0, 10.0), image_hw=(36, 36),
             focal=75.0, num_samples=96, near=5.0, far=18.0, pose=(35.0, 55.0, 10.0)),
        # incoherent rays, some missing the grid entirely (NaN disparity), black background
        Case("deg2_random_rays", (16, 16, 16), 2, (3 / 16,) * 3, num_random_rays=1500,
             num_samples=64, near=0.5, far=7.0, density_scale=relu_field_density_scale((3, 3, 3)),
             white_bkgd=False),
        # sparse, "trained-like" occupancy: densities U(-1,1) - 0.5
        Case("deg2_sparse", (24, 24, 24), 2, (3 / 24,) * 3, density_range=(-1.5, 0.5), image_hw=(32, 32),
             focal=_hotdog_focal(32), num_samples=96, density_scale=relu_field_density_scale((3, 3, 3)),
             pose=(160.0, 65.0, HOTDOG_RADIUS)),
    ]
}

_CUBE2_FEATURES = np.array(
    # the eight corner colours of the reference cube fixture (test_voxels.py:105-118)
    [10, -10, -10, -10, 10, -10, -10, -10, 10, 10, 10, -10, -10, 10, 10, 10, -10, 10, 10, 10, 10, -10, -10, -10],
    dtype=np.float32,
)


def pinhole_rays(height: int, width: int, focal: float, rot: np.ndarray, trans: np.ndarray):
    """fp32 pixel-centre rays, flat row-major (reference rendering/volumetric/utils/misc.py:27-50).
    Built from integer pixel indices (``i + 0.5`` is exact in fp32 for every i < 2**22)."""
    xs = (np.arange(width, dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    ys = (np.arange(height, dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    xx, yy = np.meshgrid(xs, ys)  # [H, W]
    f = np.float32(focal)
    cam = np.stack(
        [
            ((xx - np.float32(width * 0.5)) / f).astype(np.float32),
            (-((yy - np.float32(height * 0.5)) / f)).astype(np.float32),
            -np.ones_like(xx),
        ],
        -1,
    ).reshape(-1, 3)
    return cam, rot, trans


def build_inputs(case: Case) -> Dict[str, np.ndarray]:
    """All arrays a case needs: grid values, rays, jitter, upstream gradients."""
    rng = np.random.RandomState(case.seed)
    w, d, h = case.dims
    nf = 3 * (case.sh_degree + 1) ** 2
    dens = rng.uniform(*case.density_range, size=(w, d, h, 1)).astype(np.float32)
    if case.name == "cube2":
        feat = _CUBE2_FEATURES.reshape(2, 2, 2, 3).copy()
    else:
        feat = rng.uniform(*case.feature_range, size=(w, d, h, nf)).astype(np.float32)

    if case.image_hw is not None:
        hh, ww = case.image_hw
        rot, trans = spherical_pose(*case.pose)
        cam, _, _ = pinhole_rays(hh, ww, case.focal, rot, trans)
        # d = R @ dir_cam evaluated per ray in fp32 (matmul of [3,3] by [3,1])
        dirs = np.einsum("ij,nj->ni", rot.astype(np.float32), cam).astype(np.float32)
        origins = np.broadcast_to(trans.reshape(1, 3), dirs.shape).astype(np.float32).copy()
    else:
        n = case.num_random_rays
        origins = rng.normal(size=(n, 3)).astype(np.float32)
        origins = (origins / np.linalg.norm(origins, axis=-1, keepdims=True) * rng.uniform(2.5, 4.5, (n, 1))).astype(np.float32)
        target = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32)
        dirs = target - origins
        dirs = (dirs / np.linalg.norm(dirs, axis=-1, keepdims=True) * rng.uniform(0.7, 1.6, (n, 1))).astype(np.float32)
        # a tenth of the rays look away from the grid
        away = rng.uniform(size=n) < 0.1
        dirs[away] *= -1.0
    n = origins.shape[0]
    out = {
        "densities": dens,
        "features": feat,
        "origins": np.ascontiguousarray(origins),
        "directions": np.ascontiguousarray(dirs),
        "grad_colour": rng.normal(size=(n, 3)).astype(np.float32),
    }
    if case.jitter:
        out["jitter"] = rng.uniform(0.0, 1.0, size=(n, case.num_samples)).astype(np.float32)
        out["jitter"] = np.minimum(out["jitter"], np.float32(1.0 - 2**-24))  # keep [0, 1)
    if case.with_depth_acc_grads:
        out["grad_depth"] = rng.normal(size=(n, 1)).astype(np.float32)
        out["grad_acc"] = rng.normal(size=(n, 1)).astype(np.float32)
    return out
