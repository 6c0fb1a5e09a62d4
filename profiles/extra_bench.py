#!/usr/bin/env python
"""The other BASELINE.json configurations, measured once for the record (not the bench line):
  c2: 128^3 deg-2 grid, 400x400, 128 spp, forward only through VolumetricModel.render (in-kernel ray generation)
  c5: 512^3 deg-3 grid, 1600x1600, 512 spp, forward + backward on ONE GPU (the per-GPU share of the 8-GPU stress config)

    python profiles/extra_bench.py [c2] [c5] > profiles/r01_extra_bench.json
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from cases import HOTDOG_RADIUS, relu_field_density_scale, spherical_pose  # noqa: E402
from thr3ed_atom_b200 import _kernels  # noqa: E402
from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, make_render_args, render_hints, render_sh_voxel_grid  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose  # noqa: E402


def make(grid_n, deg, side, spp, dev, perturb=True):
    nf = 3 * (deg + 1) ** 2
    dens = torch.empty((grid_n,) * 3 + (1,), device=dev).uniform_(-1, 1)
    feat = torch.empty((grid_n,) * 3 + (nf,), device=dev).uniform_(-1, 1)
    grid = VoxelGrid(dens, feat, VoxelSize(*(3 / grid_n,) * 3), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=relu_field_density_scale((3, 3, 3)), tunable=True)
    del feat
    cfg = SHVoxGridRenderConfig(spp, CameraBounds(1.8, 6.6), perturb_sampled_points=perturb, white_bkgd=True)
    rot, trans = spherical_pose(30.0, 60.0, HOTDOG_RADIUS)
    intr, pose = CameraIntrinsics(side, side, 1111.11 * side / 800), CameraPose(rot, trans)
    return grid, cfg, intr, pose


def events(fn, iters, warmup=3):
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
    which = set(sys.argv[1:]) or {"c2", "c5"}
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"peak_gbs": peak}
    if "c2" in which:
        grid, cfg, intr, pose = make(128, 2, 400, 128, dev)
        vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
        ms = events(lambda: vol_mod.render(pose, intr), iters=50)
        rays = flatten_rays(cast_rays(intr, pose, device=dev))
        touched = int(_kernels.mark_touched_voxels(grid.kernel_desc(), rays.origins, rays.directions, make_render_args(cfg)).sum().item())
        nbytes = touched * 112 + 48 * len(rays)
        out["c2_forward_only"] = {"workload": "128^3 deg-2, 400x400, 128 spp, VolumetricModel.render (in-kernel ray generation), forward only",
                                  "ms_per_frame": ms, "rays_per_s": len(rays) / (ms * 1e-3), "frames_per_s": 1e3 / ms, "unique_voxels_touched": touched,
                                  "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak}
        del grid, vol_mod, rays
        torch.cuda.empty_cache()
    if "c5" in which:
        grid, cfg, intr, pose = make(512, 3, 1600, 512, dev)
        vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
        rays = flatten_rays(cast_rays(intr, pose, device=dev))
        pixels = torch.rand((len(rays), 3), device=dev)
        params = list(grid.parameters())

        def step():
            with render_hints(image_hw=(1600, 1600)):
                o = vol_mod.render_rays(rays)
            loss = torch.nn.functional.l1_loss(o.colour, pixels)
            for p in params:
                p.grad = None
            loss.backward()

        ms = events(step, iters=3, warmup=1)
        touched = int(_kernels.mark_touched_voxels(grid.kernel_desc(), rays.origins, rays.directions, make_render_args(cfg)).sum().item())
        nbytes = 4 * touched * 196 + 108 * len(rays)
        out["c5_one_gpu_step"] = {"workload": "512^3 deg-3, 1600x1600, 512 spp, fwd + bwd on one GPU (sample cache over the 16 GiB cap -> backward re-gathers)",
                                  "ms_per_step": ms, "rays_per_s": len(rays) / (ms * 1e-3), "unique_voxels_touched": touched,
                                  "algorithmic_bytes_step": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                                  "max_memory_allocated_gb": torch.cuda.max_memory_allocated() / 1e9}
    if "train" in which:
        # the reference trainer's batch (modules/trainers.py:278-303): rays of 8 cached views pooled, a random subset of
        # 32768 drawn with randperm -> incoherent rays in list order; and the same rays sorted back into view/pixel order
        grid, cfg, intr, pose = make(256, 2, 400, 256, dev)
        vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
        pooled = []
        for k in range(8):
            rot, trans = spherical_pose(45.0 * k, 60.0, HOTDOG_RADIUS)
            pooled.append(flatten_rays(cast_rays(intr, CameraPose(rot, trans), device=dev)))
        o = torch.cat([r.origins for r in pooled])
        d = torch.cat([r.directions for r in pooled])
        perm = torch.randperm(o.shape[0], device=dev)[:32768]
        params = list(grid.parameters())
        from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays

        for label, sel in (("random order (as the trainer passes them)", perm), ("same rays sorted by view and pixel", perm.sort().values)):
            rays = Rays(o[sel].contiguous(), d[sel].contiguous())
            pixels = torch.rand((len(rays), 3), device=dev)

            def step():
                out_ = vol_mod.render_rays(rays)
                loss = torch.nn.functional.l1_loss(out_.colour, pixels)
                for p in params:
                    p.grad = None
                loss.backward()

            ms = events(step, iters=10)
            out.setdefault("train_batch_32768_rays", []).append(
                {"workload": "256^3 deg-2, 32768 rays drawn from 8 pooled 400x400 views, 256 spp, fwd + bwd (incl. 1.95 GB gradient zero-fill)",
                 "ray_order": label, "ms_per_step": ms, "rays_per_s": len(rays) / (ms * 1e-3)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
