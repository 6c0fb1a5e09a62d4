#!/usr/bin/env python
"""The reference trainer's per-iteration work on its own batch shape (32 768 rays from 8 cached 400x400 views, 256^3 deg-2
grid, 256 spp; modules/trainers.py:278-341), three ways:
  a) reference recipe: cast_rays per view + collate + randperm subsample (utils/misc.py:117-129), then fwd + bwd
  b) device-side sampler, independent pixels (tile 1x1): one kernel draws rays + pixels, then fwd + bwd
  c) device-side sampler, 8x4 pixel tiles: every warp of the render kernels gets one coherent tile
Prints one JSON object (ms per iteration: batch assembly, forward + backward, total).

    python profiles/sampler_bench.py > profiles/r02_sampler_bench.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from cases import HOTDOG_RADIUS, relu_field_density_scale, spherical_pose  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.utils.misc import (cast_rays, collate_rays, flatten_rays, sample_random_rays_and_pixels_synchronously,  # noqa: E402
                                                               sample_training_ray_batch)
from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose  # noqa: E402


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    g, side, spp, batch, views = 256, 400, 256, 32768, 8
    shift = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
    dens = torch.empty((g, g, g, 1), device=dev).uniform_(-1, 1) - shift
    feat = torch.empty((g, g, g, 27), device=dev).uniform_(-1, 1)
    grid = VoxelGrid(dens, feat, VoxelSize(3 / g, 3 / g, 3 / g), density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                     expected_density_scale=relu_field_density_scale((3, 3, 3)), tunable=True)
    del feat, dens
    cfg = SHVoxGridRenderConfig(spp, CameraBounds(1.8, 6.6), perturb_sampled_points=True, white_bkgd=True)
    intr = CameraIntrinsics(side, side, 1111.11 * side / 800)
    poses = [CameraPose(*spherical_pose(45.0 * k, 60.0, HOTDOG_RADIUS)) for k in range(views)]
    images = torch.rand((views, side, side, 3), device=dev)
    params = list(grid.parameters())

    def reference_batch():
        rays = collate_rays([flatten_rays(cast_rays(intr, p, device=dev)) for p in poses])
        return sample_random_rays_and_pixels_synchronously(rays, images.reshape(-1, 3), batch)

    def fwd_bwd(rays, pixels):
        out = render_sh_voxel_grid(grid, rays, cfg)
        for p in params:
            p.grad = None
        torch.nn.functional.l1_loss(out.colour, pixels).backward()

    res = {"workload": f"{g}^3 deg-2, {batch} rays from {views} cached {side}x{side} views, {spp} spp, density shift {shift}"}
    for name, maker in (("reference_recipe_randperm", reference_batch),
                        ("device_sampler_pixels_1x1", lambda: sample_training_ray_batch(poses, intr, images, batch, tile=(1, 1))),
                        ("device_sampler_tiles_8x4", lambda: sample_training_ray_batch(poses, intr, images, batch, tile=(8, 4)))):
        rays, pixels = maker()
        res[name] = {"batch_assembly_ms": timed(maker), "fwd_bwd_ms": timed(lambda: fwd_bwd(rays, pixels))}
        res[name]["total_ms"] = res[name]["batch_assembly_ms"] + res[name]["fwd_bwd_ms"]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
