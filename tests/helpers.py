"""Shared test plumbing: load a golden case, run the oracles on it, compare results."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(GOLDEN)):
    if p not in sys.path:
        sys.path.insert(0, p)

from cases import CASES, Case, build_inputs  # noqa: E402  (tests/golden/cases.py)


def load_golden(name: str) -> Dict[str, np.ndarray]:
    with np.load(GOLDEN / f"{name}.npz") as f:
        return {k: f[k] for k in f.files}


def aabb_of(case: Case):
    out = []
    for n, s, c in zip(case.dims, case.voxel_size, case.location):
        half = (n * s) / 2
        out.append((c - half, c + half))
    return tuple(out)


def run_torch_port(case: Case, inp: Dict[str, np.ndarray], with_grads: bool = True):
    from oracle import torch_port as tp

    grid = tp.OracleGrid(
        densities=torch.from_numpy(inp["densities"]),
        features=torch.from_numpy(inp["features"]),
        voxel_size=case.voxel_size,
        location=case.location,
        density_scale=case.density_scale,
        density_pre=case.density_pre,
        density_post=case.density_post,
    )
    cfg = dict(
        num_samples=case.num_samples,
        near=case.near,
        far=case.far,
        jitter=torch.from_numpy(inp["jitter"]) if case.jitter else None,
        white_bkgd=case.white_bkgd,
        diffuse=case.diffuse,
        optimized_sampling=case.optimized_sampling,
    )
    o, d = torch.from_numpy(inp["origins"]), torch.from_numpy(inp["directions"])
    if not with_grads:
        with torch.no_grad():
            out = tp.render(grid, o, d, **cfg)
    else:
        out = tp.render_with_grads(
            grid,
            o,
            d,
            torch.from_numpy(inp["grad_colour"]),
            torch.from_numpy(inp["grad_depth"]) if "grad_depth" in inp else None,
            torch.from_numpy(inp["grad_acc"]) if "grad_acc" in inp else None,
            **cfg,
        )
    return {k: v.numpy() for k, v in out.items()}


def run_numpy_f64(case: Case, inp: Dict[str, np.ndarray], with_grads: bool = True):
    from oracle import numpy_f64 as nf

    return nf.render(
        inp["densities"],
        inp["features"],
        aabb_of(case),
        inp["origins"],
        inp["directions"],
        num_samples=case.num_samples,
        near=case.near,
        far=case.far,
        density_scale=case.density_scale,
        density_pre=case.density_pre,
        density_post=case.density_post,
        jitter=inp.get("jitter"),
        white_bkgd=case.white_bkgd,
        diffuse=case.diffuse,
        optimized_sampling=case.optimized_sampling,
        grad_colour=inp["grad_colour"] if with_grads else None,
        grad_depth=inp.get("grad_depth"),
        grad_acc=inp.get("grad_acc"),
    )


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def assert_outputs_close(got, want, *, atol=1e-5, rtol_depth=1e-5, what=""):
    """Forward tolerances of SURVEY.md 8(c): colour/acc abs <= 1e-5, depth rel <= 1e-5 (+ abs floor),
    disparity compared where finite, NaNs must coincide."""
    np.testing.assert_allclose(got["colour"], want["colour"], atol=atol, rtol=0, err_msg=f"{what} colour")
    np.testing.assert_allclose(got["acc"], want["acc"], atol=atol, rtol=0, err_msg=f"{what} acc")
    np.testing.assert_allclose(got["depth"], want["depth"], atol=atol * 10, rtol=rtol_depth, err_msg=f"{what} depth")
    gn, wn = np.isnan(got["disparity"]), np.isnan(want["disparity"])
    assert np.array_equal(gn, wn), f"{what}: NaN pattern of disparity differs ({gn.sum()} vs {wn.sum()})"
    # disparity = acc / depth amplifies relative error of tiny-weight rays; compare where acc is not negligible
    ok = (~wn) & (np.abs(want["acc"]) > 1e-3)
    np.testing.assert_allclose(got["disparity"][ok], want["disparity"][ok], rtol=1e-3, atol=1e-5, err_msg=f"{what} disparity")


# ---------------------------------------------------------------------------------------------
# CUDA side (product path, through the public API -> ctypes -> C ABI)
# ---------------------------------------------------------------------------------------------
_PRE = {"identity": lambda: torch.nn.Identity(), "abs": lambda: torch.abs}
_POST = {"identity": lambda: torch.nn.Identity(), "relu": lambda: torch.nn.ReLU(), "softplus": lambda: torch.nn.Softplus()}


def make_cuda_grid(case: Case, inp: Dict[str, np.ndarray], device, tunable: bool = True):
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize

    return VoxelGrid(
        densities=torch.from_numpy(inp["densities"]).to(device),
        features=torch.from_numpy(inp["features"]).to(device),
        voxel_size=VoxelSize(*case.voxel_size),
        grid_location=VoxelGridLocation(*case.location),
        density_preactivation=_PRE[case.density_pre](),
        density_postactivation=_POST[case.density_post](),
        expected_density_scale=case.density_scale,
        tunable=tunable,
    )


def make_cuda_config(case: Case, **overrides):
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

    kw = dict(
        num_samples_per_ray=case.num_samples,
        camera_bounds=CameraBounds(case.near, case.far),
        perturb_sampled_points=case.jitter,
        optimized_sampling=case.optimized_sampling,
        white_bkgd=case.white_bkgd,
        render_diffuse=case.diffuse,
    )
    kw.update(overrides)
    return SHVoxGridRenderConfig(**kw)


def run_cuda_case(case: Case, inp: Dict[str, np.ndarray], device, with_grads: bool = True, use_tile_hint: bool = False, grid=None):
    """Render a golden case with the CUDA path through ``render_sh_voxel_grid`` (+ autograd backward)."""
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    grid = make_cuda_grid(case, inp, device) if grid is None else grid
    rays = Rays(torch.from_numpy(inp["origins"]).to(device), torch.from_numpy(inp["directions"]).to(device))
    hints = {}
    if case.jitter:
        hints["jitter"] = torch.from_numpy(inp["jitter"]).to(device)
    if use_tile_hint and case.image_hw is not None:
        hints["image_hw"] = case.image_hw
    with render_hints(**hints), torch.set_grad_enabled(with_grads):
        out = render_sh_voxel_grid(grid, rays, make_cuda_config(case))
    res = {
        "colour": out.colour.detach().cpu().numpy(),
        "depth": out.depth.detach().cpu().numpy(),
        "acc": out.extra["accumulated_weight"].detach().cpu().numpy(),
        "disparity": out.extra["disparity"].detach().cpu().numpy(),
    }
    if with_grads:
        loss = (out.colour * torch.from_numpy(inp["grad_colour"]).to(device)).sum()
        if "grad_depth" in inp:
            loss = loss + (out.depth * torch.from_numpy(inp["grad_depth"]).to(device)).sum()
            loss = loss + (out.extra["accumulated_weight"] * torch.from_numpy(inp["grad_acc"]).to(device)).sum()
        grid.zero_grad()
        loss.backward()
        res["grad_densities"] = grid.densities.grad.detach().cpu().numpy()
        res["grad_features"] = grid.feature_storage.grad[..., : inp["features"].shape[-1]].detach().cpu().numpy()
        res["grad_feature_padding"] = grid.feature_storage.grad[..., inp["features"].shape[-1] :].detach().cpu().numpy()
    return res


def hash_jitter(seed: int, num_rays: int, num_samples: int, rays=None) -> np.ndarray:
    """NumPy restatement of the in-kernel counter-based jitter (csrc/r3d_device.cuh: mix32 /
    ray_rng_key / jitter_u) so that a run with the in-kernel RNG can be replayed through the oracle.
    ``rays``: explicit ray indices of the launch (default: all ``num_rays``)."""
    def mix32(h):
        h = h.astype(np.uint32)
        h ^= h >> np.uint32(16)
        h = (h * np.uint32(0x7FEB352D)).astype(np.uint32)
        h ^= h >> np.uint32(15)
        h = (h * np.uint32(0x846CA68B)).astype(np.uint32)
        h ^= h >> np.uint32(16)
        return h

    with np.errstate(over="ignore"):
        seed_lo, seed_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
        rays = np.arange(num_rays, dtype=np.uint64) if rays is None else np.asarray(rays, dtype=np.uint64)
        key = mix32((rays & np.uint64(0xFFFFFFFF)).astype(np.uint32) + seed_lo)
        key = mix32(key ^ seed_hi ^ (rays >> np.uint64(32)).astype(np.uint32))
        samples = (np.arange(num_samples, dtype=np.uint32) * np.uint32(0x9E3779B9)).astype(np.uint32)
        h = mix32((key[:, None] + samples[None, :]).astype(np.uint32))
    return ((h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
