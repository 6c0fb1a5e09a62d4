#!/usr/bin/env python
"""Training-loss step of the reference trainer (modules/trainers.py:306-341: specular render + diffuse render + L1 losses +
backward) done with two fused renders vs the single-pass specular + diffuse render, on the c3 shape and on the trainer's
own 32768-ray batch.    python profiles/dual_bench.py > gpurun_out/dual.json"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden"), str(ROOT / "profiles")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from extra_bench import events, make  # noqa: E402
from cases import HOTDOG_RADIUS, spherical_pose  # noqa: E402
from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays  # noqa: E402
from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid  # noqa: E402
from thr3ed_atom_b200.utils.imaging_utils import CameraPose  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    out = {}
    l1 = torch.nn.functional.l1_loss
    for label, side, batch in (("c3_800x800_all_rays", 800, None), ("trainer_batch_32768_random_rays", 400, 32768)):
        grid, cfg, intr, pose = make(256, 2, side, 256, dev)
        vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
        if batch is None:
            rays = flatten_rays(cast_rays(intr, pose, device=dev))
            hint = (side, side)
        else:
            pooled = [flatten_rays(cast_rays(intr, CameraPose(*spherical_pose(45.0 * k, 60.0, HOTDOG_RADIUS)), device=dev)) for k in range(8)]
            o, d = torch.cat([r.origins for r in pooled]), torch.cat([r.directions for r in pooled])
            sel = torch.randperm(o.shape[0], device=dev)[:batch]
            rays, hint = Rays(o[sel].contiguous(), d[sel].contiguous()), None
        pixels = torch.rand((len(rays), 3), device=dev)
        params = list(grid.parameters())

        def two_renders():
            with render_hints(image_hw=hint):
                spec = vol_mod.render_rays(rays)
                diff = vol_mod.render_rays(rays, render_diffuse=True)
            loss = l1(spec.colour, pixels) + l1(diff.colour, pixels)
            for p in params:
                p.grad = None
            loss.backward()

        def single_pass():
            with render_hints(image_hw=hint):
                spec, diff = vol_mod.render_rays_with_diffuse(rays)
            loss = l1(spec.colour, pixels) + l1(diff.colour, pixels)
            for p in params:
                p.grad = None
            loss.backward()

        a, b = events(two_renders, iters=10), events(single_pass, iters=10)
        out[label] = {"rays": len(rays), "two_fused_renders_ms": a, "single_pass_ms": b, "speedup": a / b,
                      "workload": "256^3 deg-2, 256 spp: specular + diffuse render, two L1 losses, backward (incl. gradient zero-fill)"}
        del grid, vol_mod, rays, pixels, params
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
