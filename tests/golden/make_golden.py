"""Generate golden vectors by running the UNMODIFIED reference implementation on CPU.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [--reference /root/reference]

The reference (akanimax/thr3ed_atom @ 8695b5a) is imported as-is; the only accommodation is two
empty ``sys.modules`` stubs for ``matplotlib.pyplot`` and ``easydict`` which the hot path imports
at module load but never calls (utils/imaging_utils.py:4, utils/misc.py:6).  Inputs come from
``cases.py`` (NumPy RandomState); the outputs of ``render_sh_voxel_grid`` and of autograd's
backward into ``_densities`` / ``_features`` are written to ``tests/golden/<case>.npz``.

Stratified jitter: the reference draws ``torch.rand`` inside ``sample_uniform_points_on_rays``
(rendering/volumetric/sample.py:63); for jittered cases ``torch.rand`` is temporarily replaced by
a function returning the case's fixed ``U[0,1)`` tensor so that the golden is reproducible.

Nothing here is imported by the product, and nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import argparse
import contextlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from cases import CASES, Case, build_inputs, spherical_pose  # noqa: E402


def import_reference(ref_root: str):
    for name in ("matplotlib", "matplotlib.pyplot", "easydict"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["easydict"].EasyDict = dict
    sys.path.insert(0, ref_root)
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose

    return types.SimpleNamespace(**locals())


@contextlib.contextmanager
def fixed_torch_rand(value):
    if value is None:
        yield
        return
    real = torch.rand

    def fake(*shape, **kw):
        assert tuple(shape) == tuple(value.shape), (shape, value.shape)
        return value.clone()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = real


_PRE = {"identity": torch.nn.Identity(), "abs": torch.abs}
_POST = {"identity": torch.nn.Identity(), "relu": torch.nn.ReLU(), "softplus": torch.nn.Softplus()}


def run_case(ref, case: Case) -> dict:
    inp = build_inputs(case)
    grid = ref.VoxelGrid(
        densities=torch.from_numpy(inp["densities"]),
        features=torch.from_numpy(inp["features"]),
        voxel_size=ref.VoxelSize(*case.voxel_size),
        grid_location=ref.VoxelGridLocation(*case.location),
        density_preactivation=_PRE[case.density_pre],
        density_postactivation=_POST[case.density_post],
        expected_density_scale=case.density_scale,
        tunable=True,
    )
    rays = ref.Rays(torch.from_numpy(inp["origins"]), torch.from_numpy(inp["directions"]))
    cfg = ref.SHVoxGridRenderConfig(
        num_samples_per_ray=case.num_samples,
        camera_bounds=ref.CameraBounds(case.near, case.far),
        perturb_sampled_points=case.jitter,
        optimized_sampling=case.optimized_sampling,
        white_bkgd=case.white_bkgd,
        render_diffuse=case.diffuse,
    )
    jitter = torch.from_numpy(inp["jitter"]) if case.jitter else None
    with fixed_torch_rand(jitter):
        out = ref.render_sh_voxel_grid(grid, rays, cfg)
    loss = (out.colour * torch.from_numpy(inp["grad_colour"])).sum()
    if case.with_depth_acc_grads:
        loss = loss + (out.depth * torch.from_numpy(inp["grad_depth"])).sum()
        loss = loss + (out.extra["accumulated_weight"] * torch.from_numpy(inp["grad_acc"])).sum()
    loss.backward()
    res = {
        "colour": out.colour.detach().numpy(),
        "depth": out.depth.detach().numpy(),
        "acc": out.extra["accumulated_weight"].detach().numpy(),
        "disparity": out.extra["disparity"].detach().numpy(),
        "grad_densities": grid.densities.grad.numpy(),
        "grad_features": grid.features.grad.numpy(),
    }
    # VoxelGrid.forward (the point lookup, voxels.py:276-331) on a scatter of points, incl. outside
    rng = np.random.RandomState(case.seed + 1)
    lo = np.array([r[0] for r in grid.aabb], np.float32)
    hi = np.array([r[1] for r in grid.aabb], np.float32)
    pts = (lo + (hi - lo) * rng.uniform(-0.15, 1.15, size=(257, 3))).astype(np.float32)
    with torch.no_grad():
        res["lookup_points"] = pts
        res["lookup_values"] = grid(torch.from_numpy(pts)).numpy()
        res["lookup_inside"] = grid.test_inside_volume(torch.from_numpy(pts)).numpy()
    # cast_rays (rendering/volumetric/utils/misc.py:12-50) for the case's camera
    if case.image_hw is not None:
        rot, trans = spherical_pose(*case.pose)
        casted = ref.cast_rays(
            ref.CameraIntrinsics(case.image_hw[0], case.image_hw[1], case.focal),
            ref.CameraPose(rot, trans),
        )
        res["cast_origins"] = casted.origins.reshape(-1, 3).numpy().copy()
        res["cast_directions"] = casted.directions.reshape(-1, 3).numpy().copy()
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    ref = import_reference(args.reference)
    torch.manual_seed(42)
    torch.set_num_threads(1)  # single-threaded ATen => run-to-run reproducible summation order
    for name, case in CASES.items():
        if args.only and name != args.only:
            continue
        res = run_case(ref, case)
        path = HERE / f"{name}.npz"
        np.savez_compressed(path, **res)
        print(f"{name:28s} rays={res['colour'].shape[0]:6d}  colour.mean={res['colour'].mean():.6f} "
              f"|gF|={np.abs(res['grad_features']).sum():.4f}  nan_disp={int(np.isnan(res['disparity']).sum())} "
              f"-> {path.name} ({path.stat().st_size/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
