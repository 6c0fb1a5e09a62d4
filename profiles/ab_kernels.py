#!/usr/bin/env python
"""A/B timing of kernel variants on the c3 workload (256^3 deg-2, 800x800, 256 spp) in ONE process:
per-variant forward / backward kernel times (CUDA events on the launching stream) and the whole step.

    python profiles/ab_kernels.py [--variants 0,8] [--pads 4,8] [--iters 10] > gpurun_out/ab.json

variant bits (include/r3d_b200.h R3dRenderConfig.variant): 0 default; 1 per-ray backward; 2 per-ray forward;
4 TMA-staged forward; 8 shared-memory staged forward (cp.async); 16 backward without the cell-merge (see r3d_render.cu).
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from cases import HOTDOG_RADIUS, relu_field_density_scale, spherical_pose  # noqa: E402
from thr3ed_atom_b200 import _kernels  # noqa: E402
from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays  # noqa: E402
import thr3ed_atom_b200.thre3d_reprs.renderers as renderers  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_hints, render_sh_voxel_grid  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,8")
    ap.add_argument("--pads", default="4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--deg", type=int, default=2)
    ap.add_argument("--side", type=int, default=800)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--density-shift", type=float, default=0.0, help="densities ~ U(-1,1) - shift (SURVEY 8d trained-like grid: 0.9)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    results = []
    ev = {"fwd": [], "bwd": []}
    real_fwd, real_bwd = _kernels.render_forward, _kernels.render_backward

    def timed(name, fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[name].append((e0, e1))
            return out
        return wrapper

    renderers._kernels.render_forward = timed("fwd", real_fwd)
    renderers._kernels.render_backward = timed("bwd", real_bwd)

    for pad in [int(x) for x in args.pads.split(",")]:
        os.environ["R3D_FEATURE_PAD"] = str(pad)
        torch.manual_seed(42)
        nf = 3 * (args.deg + 1) ** 2
        dens = torch.empty((args.grid,) * 3 + (1,), device=dev).uniform_(-1, 1) - args.density_shift
        feat = torch.empty((args.grid,) * 3 + (nf,), device=dev).uniform_(-1, 1)
        grid = VoxelGrid(dens, feat, VoxelSize(*(3 / args.grid,) * 3), density_preactivation=torch.nn.Identity(),
                         density_postactivation=torch.nn.ReLU(), expected_density_scale=relu_field_density_scale((3, 3, 3)), tunable=True)
        del feat, dens
        cfg = SHVoxGridRenderConfig(args.spp, CameraBounds(1.8, 6.6), perturb_sampled_points=True, white_bkgd=True)
        rot, trans = spherical_pose(30.0, 60.0, HOTDOG_RADIUS)
        intr, pose = CameraIntrinsics(args.side, args.side, 1111.11 * args.side / 800), CameraPose(rot, trans)
        vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
        rays = flatten_rays(cast_rays(intr, pose, device=dev))
        pixels = torch.rand((len(rays), 3), device=dev)
        params = list(grid.parameters())
        ref_colour = None
        for variant in [int(x) for x in args.variants.split(",")]:
            def step():
                with render_hints(image_hw=(args.side, args.side), variant=variant, rng_seed=1234):
                    o = vol_mod.render_rays(rays)
                loss = torch.nn.functional.l1_loss(o.colour, pixels)
                for p in params:
                    p.grad = None
                loss.backward()
                return o

            for _ in range(args.warmup):
                o = step()
            torch.cuda.synchronize()
            ev["fwd"].clear(), ev["bwd"].clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                o = step()
            e1.record()
            torch.cuda.synchronize()
            col = o.colour.detach()
            if ref_colour is None:
                ref_colour = col.clone()
            gf = grid.feature_storage.grad
            results.append({
                "pad": pad, "variant": variant, "step_ms": e0.elapsed_time(e1) / args.iters,
                "fwd_ms": sum(a.elapsed_time(b) for a, b in ev["fwd"]) / len(ev["fwd"]),
                "bwd_ms": sum(a.elapsed_time(b) for a, b in ev["bwd"]) / len(ev["bwd"]),
                "max_abs_colour_diff_vs_first": float((col - ref_colour).abs().max()),
                "grad_feat_l2": float(gf.double().norm()) if gf is not None else None,
                "colour_sum": float(col.double().sum()), "colour_sq": float((col.double() ** 2).sum()),
            })
            print(json.dumps(results[-1]), file=sys.stderr, flush=True)
        del grid, vol_mod, rays, pixels, params
        torch.cuda.empty_cache()
    print(json.dumps({"workload": f"{args.grid}^3 deg {args.deg}, {args.side}^2, {args.spp} spp, density shift {args.density_shift}", "results": results}, indent=1))


if __name__ == "__main__":
    main()
