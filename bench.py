#!/usr/bin/env python
"""Benchmark of the SH-voxel-grid render hot path (BASELINE.json metric: rays/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): a 256^3 SH-degree-2
ReLU-field voxel grid (U(-1,1) init, 3x3x3 world, expected_density_scale 33.33), one 800x800 "hotdog
shape" pinhole camera (pose_spherical(30, 60, 4.031128), focal 1111.11, bounds 1.8..6.6), 256 stratified
(jittered) samples per ray, white background, L1 loss against U(0,1) pixels.  One STEP = one pass of
the hot path over the whole 640 000-ray batch through the public API:
    render_rays (fused forward kernel) -> l1_loss -> backward (gradient zero-fill + fused backward kernel)
At N > 1 every rank renders its own camera view of the same grid (weak scaling: 640 000 rays per GPU)
and the step ends with the NCCL all-reduce of the grid gradient (the path's only exchange).

Printed JSON (one line, rank 0): see the keys at the bottom; `value` is device-resident throughput,
`e2e` is the same step with rays/pixels coming from pinned host memory and the rendered colour +
loss going back, `roofline` is for the dominant kernel (the render kernel with the longer mean launch; both are also under
`roofline_fwd` / `roofline_bwd`), `cpu_baseline` is the CPU oracle
port (oracle/torch_port.py -- the reference's algorithm on ATen CPU kernels) on a bounded ray sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

def _baseline_metric() -> str:
    """The metric string of BASELINE.json (the contract names the metric; fall back to a literal copy if the file is absent)."""
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:  # noqa: BLE001
        return "rays/sec (fwd+bwd) at 256\u00b3 grid, 800\u00b2, 256 spp; HBM roofline %"


METRIC = _baseline_metric()
WORKLOADS = {
    # name: (grid, sh_degree, image side, samples/ray)
    "c3_256cube_deg2_800px_256spp": (256, 2, 800, 256),
    "c2_128cube_deg2_400px_128spp": (128, 2, 400, 128),
    "c1_32cube_deg0_64px_32spp": (32, 0, 64, 32),
    # BASELINE.json configs[4] (8-GPU stress shape; not the bench line): --workload c5_512cube_deg3_1600px_512spp
    "c5_512cube_deg3_1600px_512spp": (512, 3, 1600, 512),
}
HOTDOG_RADIUS, NEAR, FAR = 4.031128406524658, 1.8, 6.6
WORLD = (3.0, 3.0, 3.0)


def density_scale() -> float:
    diag = float(np.sqrt(sum(e * e for e in WORLD)))
    return ((float(np.sqrt(27.0)) * 100.0) / diag) / 3


def make_grid_values(grid: int, deg: int, seed: int = 42):
    """U(-1,1) densities / features, the train script's own init (train_...py:202-206), seed 42."""
    g = torch.Generator().manual_seed(seed)
    nf = 3 * (deg + 1) ** 2
    dens = torch.empty((grid, grid, grid, 1), dtype=torch.float32).uniform_(-1.0, 1.0, generator=g)
    feat = torch.empty((grid, grid, grid, nf), dtype=torch.float32).uniform_(-1.0, 1.0, generator=g)
    return dens, feat


def camera_for_rank(rank: int, side: int):
    from cases import spherical_pose

    yaw = 30.0 + 45.0 * rank  # canonical view, then 45-degree steps for further views
    rot, trans = spherical_pose(yaw, 60.0, HOTDOG_RADIUS)
    return rot, trans, 1111.11 * side / 800.0


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm, ATen CPU kernels), bounded ray sample
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rays_per_sec(workload: str, sample_rays: int, steps: int, warmup: int, threads: int):
    from oracle import torch_port as tp

    grid_n, deg, side, spp = WORKLOADS[workload]
    torch.set_num_threads(threads)
    dens, feat = make_grid_values(grid_n, deg)
    grid = tp.OracleGrid(dens, feat, tuple(w / grid_n for w in WORLD), (0.0, 0.0, 0.0), density_scale(), "identity", "relu")
    rot, trans, focal = camera_for_rank(0, side)
    origins, dirs = tp.cast_pinhole_rays(side, side, focal, rot, trans)
    n = origins.shape[0]
    gen = torch.Generator().manual_seed(7)
    pixels = torch.rand((n, 3), generator=gen)
    times = []
    for it in range(warmup + steps):
        # a different strided subset of the image every step (same geometry statistics as the full batch)
        idx = torch.arange(it % 7, n, max(1, n // sample_rays))[:sample_rays]
        o, d, px = origins[idx].contiguous(), dirs[idx].contiguous(), pixels[idx]
        jitter = torch.rand((idx.numel(), spp), generator=gen)
        t0 = time.perf_counter()
        dq = dens.detach().requires_grad_(True)
        fq = feat.detach().requires_grad_(True)
        import dataclasses

        out = tp.render(dataclasses.replace(grid, densities=dq, features=fq), o, d, num_samples=spp, near=NEAR, far=FAR,
                        jitter=jitter, white_bkgd=True)
        loss = torch.nn.functional.l1_loss(out["colour"], px)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    return idx.numel() / mean, mean, idx.numel()


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port (the reference is pure
    Python/PyTorch and cannot travel to the GPU box; oracle/torch_port.py restates it op by op on the same ATen kernels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, max(1, min(args.warmup, 2))
    rps, mean_s, sample = cpu_oracle_rays_per_sec(workload, args.cpu_sample_rays, steps, warmup, threads)
    grid_n, deg, side, spp = WORKLOADS[workload]
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "grid": grid_n, "sh_degree": deg, "image": [side, side], "samples_per_ray": spp,
                   "rays_per_step": sample, "device": "cpu"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} strided rays of the {side}x{side} view per step, fwd+bwd, chunked like the reference must be"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def roofline_entries(bytes_fwd, fwd_ms, bytes_bwd, bwd_ms, peak, peak_src, traffic_fwd=None, traffic_bwd=None):
    """HBM-roofline entries of the two render kernels: ALGORITHMIC bytes / mean launch duration / measured
    peak.  ``roofline`` is the entry of the DOMINANT kernel (the one with the longer mean launch); both kernels are also
    reported under their own keys."""
    def entry(kernel, nbytes, ms, traffic):
        achieved = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": nbytes, "kernel_ms": ms}

    fwd = entry("render_fwd_group_kernel", bytes_fwd, fwd_ms, traffic_fwd)
    bwd = entry("render_bwd_coop_kernel", bytes_bwd, bwd_ms, traffic_bwd)
    return {"roofline": dict(fwd if fwd_ms > bwd_ms else bwd), "roofline_fwd": fwd, "roofline_bwd": bwd}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_256cube_deg2_800px_256spp", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rays", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-perturb", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--flat-order", action="store_true", help="do not pass the image-shape hint (rays in list order)")
    ap.add_argument("--reducer", default=os.environ.get("R3D_BENCH_REDUCER", "auto"), choices=["auto", "nccl", "nvls"],
                    help="N>1: gradient exchange = NCCL all-reduce, or the in-switch multimem kernel (csrc/r3d_comm.cu). "
                         "auto = nvls at 8+ ranks (measured 4.04 vs 4.18 ms), nccl below (at 2 ranks the multicast path "
                         "moves 1.5x the ring's bytes: 4.9 vs 3.3 ms)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, make_render_args, render_hints, render_sh_voxel_grid
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL writes its version banner to stdout; stdout carries the one JSON line
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    grid_n, deg, side, spp = WORKLOADS[args.workload]
    nf = 3 * (deg + 1) ** 2
    if grid_n**3 * (nf + 1) * 4 > 8 * 2**30:
        # 512^3 deg 3 is 26 GB per replica: draw it on the device (same U(-1,1) law, same seed on every rank)
        gdev = torch.Generator(device=device).manual_seed(42)
        dens = torch.empty((grid_n, grid_n, grid_n, 1), dtype=torch.float32, device=device).uniform_(-1.0, 1.0, generator=gdev)
        feat = torch.empty((grid_n, grid_n, grid_n, nf), dtype=torch.float32, device=device).uniform_(-1.0, 1.0, generator=gdev)
        args.no_cpu_baseline = True
    else:
        dens, feat = make_grid_values(grid_n, deg)
    voxel_grid = VoxelGrid(
        densities=dens.to(device), features=feat.to(device), voxel_size=VoxelSize(*[w / grid_n for w in WORLD]),
        density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
        expected_density_scale=density_scale(), tunable=True,
    )
    del dens, feat
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=spp, camera_bounds=CameraBounds(NEAR, FAR),
                                perturb_sampled_points=not args.no_perturb, white_bkgd=True)
    vol_mod = VolumetricModel(voxel_grid, render_sh_voxel_grid, cfg, device=device)

    rot, trans, focal = camera_for_rank(rank, side)
    rays = flatten_rays(cast_rays(CameraIntrinsics(side, side, focal), CameraPose(rot, trans), device=device))
    n_rays = len(rays)
    gen = torch.Generator().manual_seed(7 + rank)
    pixels_host = torch.rand((n_rays, 3), generator=gen).pin_memory()
    pixels = pixels_host.to(device)
    origins_host, dirs_host = rays.origins.cpu().pin_memory(), rays.directions.cpu().pin_memory()
    colour_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
    hint = None if args.flat_order else (side, side)

    # ---- algorithmic bytes: exact count of touched voxels (untimed bitmap pass) ----
    with render_hints(image_hw=hint, rng_seed=1234):
        margs = make_render_args(cfg)
    bitmap = _kernels.mark_touched_voxels(voxel_grid.kernel_desc(), rays.origins, rays.directions, margs)
    touched = int(bitmap.sum().item())
    del bitmap
    rec_bytes = 4 * (nf + 1)  # one (density + SH) voxel record
    bytes_fwd = touched * rec_bytes + 48 * n_rays
    bytes_bwd = 3 * touched * rec_bytes + 60 * n_rays

    # kernel-level CUDA events (on the launching stream = torch's current stream)
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

    import thr3ed_atom_b200.thre3d_reprs.renderers as _renderers

    launches = {"n": 0}

    def counted(fn):
        def wrapper(*a, **k):
            launches["n"] += 1
            return fn(*a, **k)
        return wrapper

    _renderers._kernels.render_forward = counted(timed("fwd", real_fwd))
    _renderers._kernels.render_backward = counted(timed("bwd", real_bwd))

    params = list(voxel_grid.parameters())
    reducer = None
    want_nvls = args.reducer == "nvls" or (args.reducer == "auto" and world >= 8)
    if world > 1 and want_nvls:
        from thr3ed_atom_b200.distributed import NVLSGradientReducer

        # gradients live in symmetric memory; the NVSwitch reduces them in place.  Every rank must take the same branch,
        # so a failure anywhere (no multicast support) sends all ranks to NCCL.
        ok = torch.ones(1, device=device)
        try:
            reducer = NVLSGradientReducer(voxel_grid)
        except Exception as e:  # noqa: BLE001
            if args.reducer == "nvls":
                raise
            ok.zero_()
            print(f"[bench] rank {rank}: NVLS reducer unavailable ({e!r}); using NCCL", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0 and reducer is not None:
            reducer.close()
            for p in voxel_grid.parameters():
                p.grad = None
            reducer = None

    def step(o, d, px):
        with render_hints(image_hw=hint, variant=args.variant):
            out = vol_mod.render_rays(Rays(o, d))
        loss = torch.nn.functional.l1_loss(out.colour, px)
        if reducer is not None:
            reducer.zero_grad()  # same 1.95 GB memset as autograd's fresh zero-filled buffers
            loss.backward()  # the backward kernel accumulates straight into the symmetric buffers
            reducer.all_reduce()
            launches["n"] += 1  # r3d multimem all-reduce kernel
            return loss, out
        for p in params:
            p.grad = None
        loss.backward()
        if world > 1:
            # the path's only exchange: sum of the per-rank dense grid gradients
            for p in params:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
        return loss, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + timed region (device-resident inputs).  The clock sampler runs from the first warm-up step on:
    #      nvidia-smi needs ~100 ms to start, the timed region itself can be shorter than that. ----
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step(rays.origins, rays.directions, pixels)
        barrier()
        ev["fwd"].clear(), ev["bwd"].clear()
        launches["n"] = 0
        barrier()
        start.record()
        for _ in range(args.steps):
            step(rays.origins, rays.directions, pixels)
        stop.record()
        barrier()
    total_ms = start.elapsed_time(stop)
    gpu_launches = launches["n"]
    fwd_ms = float(np.mean([a.elapsed_time(b) for a, b in ev["fwd"]]))
    bwd_ms = float(np.mean([a.elapsed_time(b) for a, b in ev["bwd"]]))

    # ---- e2e: host buffers in, loss + colour out, every step.  The step's inputs (rays + pixels, 23 MB) are copied from
    #      pinned host memory on a copy stream while the previous step computes (what a data loader does); every step's
    #      copies, the colour read-back and the loss read-back (a host sync) are inside the timed region. ----
    copy_stream = torch.cuda.Stream(device=device)

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            staged = [t.to(device, non_blocking=True) for t in (origins_host, dirs_host, pixels_host)]
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return staged, ready

    def e2e_steps(count):
        nxt = stage_inputs()
        for k in range(count):
            (o, d, px), ready = nxt
            main = torch.cuda.current_stream(device)
            main.wait_event(ready)
            for t in (o, d, px):
                t.record_stream(main)
            if k + 1 < count:
                nxt = stage_inputs()  # next step's inputs fly while this step computes
            loss, out = step(o, d, px)
            colour_host.copy_(out.colour.detach(), non_blocking=True)
            float(loss.item())

    e2e_steps(1)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_s * 1e3, fwd_ms, bwd_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_ms, bwd_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        traffic_fwd = traffic_bwd = None
        tp_path = ROOT / "profiles" / "traffic.json"
        if tp_path.exists():
            try:
                measured = json.loads(tp_path.read_text()).get(args.workload, {})
                traffic_fwd, traffic_bwd = measured.get("render_fwd_dram_bytes"), measured.get("render_bwd_dram_bytes")
            except Exception:
                traffic_fwd = traffic_bwd = None
        ms_per_step = total_ms / args.steps
        value = world * n_rays / (ms_per_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": args.workload, "grid": grid_n, "sh_degree": deg, "image": [side, side], "samples_per_ray": spp,
                "rays_per_gpu_per_step": n_rays, "ray_order": "flat" if args.flat_order else "image(8x4 tiles)",
                "perturb": not args.no_perturb, "step": "render_rays fwd + l1_loss + backward (grad zero-fill + fused bwd)"
                + ((" + NCCL all-reduce(grid grad)" if reducer is None else " + in-switch NVLS all-reduce(grid grad)") if world > 1 else ""),
                "l2": f"inputs exceed L2: the grid is {grid_n**3 * rec_bytes / 1e6:.0f} MB vs 126 MB",
                "variant": args.variant,
            },
            "e2e": {"value": world * n_rays * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": 3 * n_rays * 12, "d2h_bytes_per_step": n_rays * 12 + 4},
            "gpu_launches": gpu_launches,
            **roofline_entries(bytes_fwd, fwd_ms, bytes_bwd, bwd_ms, peak, peak_src, traffic_fwd, traffic_bwd),
            "unique_voxels_touched": touched, "voxel_record_bytes": rec_bytes,
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            del vol_mod, voxel_grid, params
            torch.cuda.empty_cache()
            rps, mean_s, sample = cpu_oracle_rays_per_sec(args.workload, args.cpu_sample_rays, steps=2, warmup=1, threads=threads)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": f"{sample} strided rays of the same {side}x{side} view, fwd+bwd, {mean_s:.2f} s per pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
