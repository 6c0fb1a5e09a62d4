#!/usr/bin/env python
"""Micro-benchmarks of the small kernels around the renderer, each timed with
CUDA events after warm-up and reported against its own HBM roofline (algorithmic bytes / time vs MEASURED_PEAKS.json).

    python profiles/aux_bench.py > profiles/r01_aux_bench.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from thr3ed_atom_b200 import _kernels  # noqa: E402
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402


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
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"peak_gbs": peak, "kernels": []}

    def add(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out["kernels"].append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "note": note})

    # cast_rays: 1600 x 1600 camera, 24 B written per ray
    import numpy as np
    rot, trans = np.eye(3, dtype=np.float32), np.array([0.0, 0.0, 4.0], np.float32)
    h = w = 1600
    add("cast_rays_kernel", timed(lambda: _kernels.cast_rays(h, w, 2222.0, rot, trans, dev)), h * w * 24,
        "1600x1600 rays; includes the two torch.empty allocations (caching allocator)")

    # fused Adam on a 256^3 x 28 grid tensor: 16 B read + 12 B written per element
    n = 256**3 * 28
    p, g = torch.randn(n, device=dev), torch.randn(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    add("adam_kernel", timed(lambda: _kernels.adam_step(p, g, m, v, lr=0.03, beta1=0.9, beta2=0.999, eps=1e-8, step=3), iters=10), n * 28,
        "256^3 x 28 floats (the c3 feature tensor): read p,g,m,v + write p,m,v")
    params = [p.clone().requires_grad_(True)]
    params[0].grad = g
    opt = torch.optim.Adam(params, lr=0.03)
    add("torch.optim.Adam (foreach, reference trainer's optimizer)", timed(lambda: opt.step(), iters=10), n * 28, "same tensor, for comparison")
    del p, g, m, v, params, opt
    torch.cuda.empty_cache()

    # VoxelGrid.forward on scattered points: 128^3 deg-2 grid, 4 M points: 8 records of 108 B + 8 densities gathered, 112 B written
    gen = torch.Generator().manual_seed(0)
    grid = VoxelGrid(torch.rand((128, 128, 128, 1), generator=gen).to(dev), torch.rand((128, 128, 128, 27), generator=gen).to(dev),
                     VoxelSize(3 / 128, 3 / 128, 3 / 128), density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(), tunable=True)
    pts = ((torch.rand((4_000_000, 3), generator=gen) - 0.5) * 3.2).to(dev)
    desc = grid.kernel_desc()
    add("lookup_fwd_kernel", timed(lambda: _kernels.grid_lookup_forward(desc, pts), iters=10), pts.shape[0] * (12 + 112),
        "4 M random points, 128^3 deg-2 grid; algorithmic = 12 B point in + 112 B out per point (the 235 MB grid is L2-resident-ish)")
    gout = torch.randn((pts.shape[0], 28), device=dev)
    gd, gf = torch.zeros_like(desc.densities), torch.zeros_like(desc.features)
    add("lookup_bwd_kernel", timed(lambda: _kernels.grid_lookup_backward(desc, pts, gout, gd, gf), iters=10), pts.shape[0] * (12 + 112),
        "same points; scatter of 8 x 28 atomics per point")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
