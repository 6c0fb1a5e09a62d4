#!/usr/bin/env python
"""Benchmark of the SH-voxel-grid render hot path (BASELINE.json metric: rays/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): a 256^3 SH-degree-2
ReLU-field voxel grid (U(-1,1) init, 3x3x3 world, expected_density_scale 33.33), one 800x800 "hotdog
shape" pinhole camera (pose_spherical(30, 60, 4.031128), focal 1111.11, bounds 1.8..6.6), 256 stratified
(jittered) samples per ray, white background, L1 loss against U(0,1) pixels.  One STEP = one full training
step over the whole 640 000-ray batch through the public API, as the reference trainer runs it
(modules/trainers.py:339-341: zero_grad -> backward -> optimizer.step):
    zero_grad -> render_rays (fused forward kernel) -> l1_loss -> backward (fused backward kernel) -> Adam step
At N = 1 the optimizer is the fused dense Adam kernel.  At N > 1 every rank renders its own camera view of the
same replicated grid (weak scaling: 640 000 rays per GPU) and gradient exchange + optimizer are ONE in-switch
kernel: reduce-scatter -> shard-local Adam -> all-gather over NVLink/NVSwitch multicast (csrc/r3d_comm.cu);
`--exchange nccl` = NCCL all-reduce + local Adam.  `--no-optimizer` gives the round-1 step (forward + backward
+ gradient all-reduce); the line also carries that number as `fwd_bwd_only`.

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
def cpu_oracle_rays_per_sec(workload: str, sample_rays: int, steps: int, warmup: int, threads: int, with_optimizer: bool = True):
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
    if not with_optimizer:
        return idx.numel() / mean, mean, idx.numel(), 0.0
    # the step's optimizer half: torch.optim.Adam over the dense grid, as the reference trainer builds it
    # (modules/trainers.py:242-245); its cost does not depend on the ray count, so it is timed once on the full grid
    pd, pf = dens.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [pd, pf], "lr": 1e-5}], betas=(0.9, 0.999))
    pd.grad, pf.grad = torch.zeros_like(pd), torch.zeros_like(pf)
    opt.step()
    t0 = time.perf_counter()
    opt.step()
    adam_s = time.perf_counter() - t0
    # full-step throughput extrapolated from the sample: the render part scales with the rays, the optimizer part does not
    full = n * mean / idx.numel() + adam_s
    return n / full, mean, idx.numel(), adam_s


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port (the reference is pure
    Python/PyTorch and cannot travel to the GPU box; oracle/torch_port.py restates it op by op on the same ATen kernels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, max(1, min(args.warmup, 2))
    rps, mean_s, sample, adam_s = cpu_oracle_rays_per_sec(workload, args.cpu_sample_rays, steps, warmup, threads,
                                                          with_optimizer=not getattr(args, "no_optimizer", False))
    grid_n, deg, side, spp = WORKLOADS[workload]
    n_full = side * side
    mean_s = n_full / rps  # extrapolated full-batch step
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "grid": grid_n, "sh_degree": deg, "image": [side, side], "samples_per_ray": spp,
                   "rays_per_step": sample, "device": "cpu"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} strided rays of the {side}x{side} view per step (fwd+bwd, chunked like the reference must be), "
                                   f"extrapolated to the {n_full}-ray batch, + one dense torch.optim.Adam step over the grid ({adam_s:.2f} s)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT_FD = None


def reserve_stdout_for_the_json_line() -> None:
    """stdout carries exactly ONE line, the result.  Libraries write there too (NCCL prints its version banner to stdout at any
    NCCL_DEBUG level): point file descriptor 1 at stderr for the whole run and keep the real stdout for `emit`."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT_FD is None:
        sys.stdout.write(data.decode()), sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT_FD, data)


def fp32_peak_tflops(sm_mhz: float, sms: int = 148) -> float:
    """FP32 SIMT peak: SMs x 128 lanes x 2 flop x clock (SURVEY.md 8d companion figure)."""
    return sms * 128 * 2 * sm_mhz * 1e6 / 1e12


def roofline_entries(bytes_fwd, fwd_ms, bytes_bwd, bwd_ms, peak, peak_src, traffic, stats, rec_bytes, nf, sm_mhz):
    """HBM-roofline entries of the two render kernels: ALGORITHMIC bytes / mean launch duration / measured
    peak.  ``roofline`` is the entry of the DOMINANT kernel (the one with the longer mean launch); both kernels are also
    reported under their own keys, each with the 8d companion figures: gather-request bytes (what the reference's two
    grid_sample calls request: in-range corner references x record bytes) and the FP32 work against the SIMT peak."""
    # FMA counts per sample: probe = 8-corner density interpolation + position/cell arithmetic; a contributing sample adds
    # the 8 x F record contraction, the SH weighting and compositing (forward), and the same contraction transposed plus
    # the chain rule (backward)
    fma_probe, fma_fwd, fma_bwd = 8 + 30, 8 * nf + nf + 24, 8 * nf + nf + 8 + 40
    flops_fwd = 2.0 * (stats["samples_inside"] * fma_probe + stats["samples_contributing"] * fma_fwd)
    flops_bwd = 2.0 * stats["samples_contributing"] * (fma_bwd + 20)
    fp32_peak = fp32_peak_tflops(sm_mhz)

    def entry(kernel, nbytes, ms, key, flops, requests):
        achieved = nbytes / (ms * 1e-3) / 1e9
        tf = flops / (ms * 1e-3) / 1e12
        return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get(key), "traffic_source": traffic.get("source"), "peak_source": peak_src,
                "algorithmic_bytes": nbytes, "kernel_ms": ms,
                "gather_request_bytes": requests, "fp32_flops": flops, "fp32_tflops": tf, "fp32_peak_tflops": fp32_peak,
                "fp32_frac": tf / fp32_peak}

    requests = stats["corner_refs"] * rec_bytes
    fwd = entry("render_fwd_group_kernel", bytes_fwd, fwd_ms, "render_fwd_dram_bytes", flops_fwd, requests)
    bwd = entry("render_bwd_coop_kernel", bytes_bwd, bwd_ms, "render_bwd_dram_bytes", flops_bwd, requests)
    return {"roofline": dict(fwd if fwd_ms > bwd_ms else bwd), "roofline_fwd": fwd, "roofline_bwd": bwd}


def load_traffic(workload: str) -> dict:
    """ncu-measured DRAM bytes per launch (profiles/traffic.json, written by profiles/measure_traffic.py on the GPU box).
    The file is stamped with the digest of the library sources it was measured with; a stale file yields null, not a
    number that no longer describes the kernels."""
    path = ROOT / "profiles" / "traffic.json"
    try:
        from thr3ed_atom_b200 import build as _build

        data = json.loads(path.read_text())
        if data.get("lib_digest") != _build._source_digest():
            return {"source": "profiles/traffic.json is stale (library sources changed since it was measured): re-run profiles/measure_traffic.py"}
        entry = dict(data.get(workload, {}))
        entry["source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, {data.get('measured', '?')}"
        return entry
    except Exception as e:  # noqa: BLE001
        return {"source": f"unavailable ({e.__class__.__name__})"}


def sm_side_ceiling(workload: str, fwd_ms: float) -> dict:
    """Companion of `roofline_fwd`: the forward is bound inside the SM, not by HBM, so next to the HBM fraction the line
    carries the time of a forward that moves the same data and computes nothing -- the two-kernel forward of the measurement build
    (probe kernel + gather kernel, csrc/r3d_fwd_split.cuh) with the gather's arithmetic removed (full-width loads kept alive by two LOP3
    per record), measured HERE in a subprocess on the same GPU after the timed region.  Needs the measurement build
    (`python -m thr3ed_atom_b200.build --ab`) made from the current sources; otherwise the entry says why it is absent."""
    import subprocess

    from thr3ed_atom_b200 import build as _build

    if workload != "c3_256cube_deg2_800px_256spp":
        return {"unavailable": "measured for the c3 workload only"}
    if not _build.ab_is_current():
        return {"unavailable": "no measurement build of the current sources (python -m thr3ed_atom_b200.build --ab)"}
    env = dict(os.environ, R3D_LIB_PATH=str(_build.AB_LIB_PATH), R3D_SPLIT_MODE="9")
    proc = subprocess.run([sys.executable, str(ROOT / "profiles" / "ab_kernels.py"), "--variants", "32768", "--iters", "5", "--warmup", "2"],
                          env=env, capture_output=True, text=True, timeout=300, stdin=subprocess.DEVNULL)
    if proc.returncode != 0:
        return {"unavailable": f"measurement run failed: {proc.stderr[-200:]}"}
    res = json.loads(proc.stdout)["results"][0]
    floor_ms = float(res["fwd_ms"])
    return {
        "no_arithmetic_two_kernel_forward_ms": floor_ms, "forward_ms": fwd_ms, "forward_over_floor": fwd_ms / floor_ms,
        "what": "probe kernel (march, density, depth / acc) + gather kernel with its arithmetic removed (same addresses, same shared-memory "
                "hand-over, full-width loads), run back to back on this GPU; the fused product forward overlaps the two phases and does the "
                "arithmetic. The gather is bound by the L1 data pipe (profiles/r02_split_forward_ncu.md: 86 % busy), not by HBM.",
    }


def pytorch_gpu_baseline(workload: str, device, chunk: int = 32768, chunks: int = 4):
    """Informational row: the reference's op-by-op PyTorch path (the oracle port: same ATen ops in the
    same order) on the SAME B200, forward + backward, ray-chunked at 32768 as the reference itself has to be
    (modules/volumetric_model.py:151-167), on a bounded sample of the view's rows."""
    import dataclasses

    from oracle import torch_port as tp

    grid_n, deg, side, spp = WORKLOADS[workload]
    dens, feat = make_grid_values(grid_n, deg)
    grid = tp.OracleGrid(dens.to(device), feat.to(device), tuple(w / grid_n for w in WORLD), (0.0, 0.0, 0.0), density_scale(), "identity", "relu")
    rot, trans, focal = camera_for_rank(0, side)
    origins, dirs = tp.cast_pinhole_rays(side, side, focal, rot, trans)
    n = origins.shape[0]
    gen = torch.Generator().manual_seed(7)
    pixels = torch.rand((n, 3), generator=gen)
    times = []
    for it in range(chunks + 1):
        start = ((it * 7919 * chunk) % max(1, n - chunk)) // side * side  # whole rows, spread over the image
        o, d, px = (t[start:start + chunk].to(device) for t in (origins, dirs, pixels))
        jitter = torch.rand((o.shape[0], spp), device=device)
        dq, fq = grid.densities.detach().requires_grad_(True), grid.features.detach().requires_grad_(True)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = tp.render(dataclasses.replace(grid, densities=dq, features=fq), o, d, num_samples=spp, near=NEAR, far=FAR, jitter=jitter, white_bkgd=True)
        torch.nn.functional.l1_loss(out["colour"], px).backward()
        e1.record()
        torch.cuda.synchronize(device)
        if it > 0:
            times.append(e0.elapsed_time(e1))
        del out, dq, fq
    ms = float(np.mean(times))
    return {"value": chunk / (ms * 1e-3), "unit": "rays/s", "kind": "port (oracle/torch_port.py on cuda: the reference's ATen op sequence)",
            "sample": f"{chunks} chunks of {chunk} rays of the same view, fwd+bwd, {ms:.1f} ms per chunk", "ms_per_chunk": ms}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    reserve_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_256cube_deg2_800px_256spp", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rays", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the informational PyTorch-on-the-same-GPU row")
    ap.add_argument("--no-perturb", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--flat-order", action="store_true", help="do not pass the image-shape hint (rays in list order)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling (BASELINE.json configs[3] read literally): ONE 800x800 view, its rays sharded over the ranks in "
                         "contiguous row blocks (distributed.shard_rays); default is weak scaling, one full view per rank")
    ap.add_argument("--no-optimizer", action="store_true",
                    help="round-1 step definition: forward + backward (+ gradient all-reduce at N > 1), no optimizer step")
    ap.add_argument("--lr", type=float, default=1e-5,
                    help="Adam rate of the timed steps: the full update is computed and written; the rate is small so that the "
                         "U(-1,1) statistics of the synthetic grid (hence the workload) do not drift over the run")
    ap.add_argument("--exchange", default=os.environ.get("R3D_BENCH_EXCHANGE", "auto"), choices=["auto", "fused", "nccl", "nvls"],
                    help="N>1: fused = in-switch reduce-scatter -> shard-local Adam -> all-gather kernel (csrc/r3d_comm.cu); "
                         "nccl = NCCL all-reduce then the local fused Adam; nvls = in-switch all-reduce kernel then the local fused "
                         "Adam; auto = fused when NVLS multicast is available, else nccl")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.optim import FusedGridAdam
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
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    grid_n, deg, side, spp = WORKLOADS[args.workload]
    nf = 3 * (deg + 1) ** 2
    big = grid_n**3 * (nf + 1) * 4 > 8 * 2**30
    if big:
        # 512^3 deg 3 is 26 GB per replica: draw it on the device (same U(-1,1) law, same seed on every rank)
        gdev = torch.Generator(device=device).manual_seed(42)
        dens = torch.empty((grid_n, grid_n, grid_n, 1), dtype=torch.float32, device=device).uniform_(-1.0, 1.0, generator=gdev)
        feat = torch.empty((grid_n, grid_n, grid_n, nf), dtype=torch.float32, device=device).uniform_(-1.0, 1.0, generator=gdev)
        args.no_cpu_baseline = args.no_gpu_baseline = True
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

    rot, trans, focal = camera_for_rank(0 if args.strong else rank, side)
    rays = flatten_rays(cast_rays(CameraIntrinsics(side, side, focal), CameraPose(rot, trans), device=device))
    gen = torch.Generator().manual_seed(7 + (0 if args.strong else rank))
    pixels_host = torch.rand((len(rays), 3), generator=gen)
    rows = side
    if args.strong and world > 1:
        from thr3ed_atom_b200.distributed import shard_rays

        assert side % world == 0, "--strong shards whole image rows"
        shard = shard_rays(rays, pixels_host)  # contiguous block of rows: the image-tile mapping stays valid inside a rank
        rays = Rays(shard.rays.origins.contiguous(), shard.rays.directions.contiguous())
        pixels_host, rows = shard.pixels.contiguous(), side // world
    n_rays = len(rays)
    pixels_host = pixels_host.pin_memory()
    pixels = pixels_host.to(device)
    origins_host, dirs_host = rays.origins.cpu().pin_memory(), rays.directions.cpu().pin_memory()
    colour_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
    hint = None if args.flat_order else (rows, side)

    # ---- algorithmic bytes + companion figures: exact counts from untimed passes over the same rays ----
    with render_hints(image_hw=hint, rng_seed=1234):
        margs = make_render_args(cfg)
    bitmap = _kernels.mark_touched_voxels(voxel_grid.kernel_desc(), rays.origins, rays.directions, margs)
    touched = int(bitmap.sum().item())
    del bitmap
    stats = _kernels.sample_statistics(voxel_grid.kernel_desc(), rays.origins, rays.directions, margs)
    rec_bytes = 4 * (nf + 1)  # one (density + SH) voxel record
    bytes_fwd = touched * rec_bytes + 48 * n_rays
    bytes_bwd = 3 * touched * rec_bytes + 60 * n_rays

    # kernel-level CUDA events (on the launching stream = torch's current stream)
    ev = {"fwd": [], "bwd": [], "opt": []}
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

    # ---- exchange + optimizer (reference modules/trainers.py:339-341: zero_grad -> backward -> optimizer.step) ----
    params = list(voxel_grid.parameters())
    betas, eps = (0.9, 0.999), 1e-8
    reducer = sharded = local_opt = None
    exchange = "none"
    if world > 1:
        want = args.exchange
        ok = torch.ones(1, device=device)
        if want in ("auto", "fused") and not args.no_optimizer:
            from thr3ed_atom_b200.distributed import NVLSShardedAdam

            # every rank must take the same branch, so a failure anywhere (no multicast support) sends all ranks to NCCL
            try:
                sharded = NVLSShardedAdam(voxel_grid, lr=args.lr, betas=betas, eps=eps)
            except Exception as e:  # noqa: BLE001
                if want == "fused":
                    raise
                ok.zero_()
                print(f"[bench] rank {rank}: fused NVLS optimizer unavailable ({e!r}); using NCCL", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) == 0.0:
                if sharded is not None:
                    sharded.close()
                sharded = None
            exchange = ("fused-" + sharded.exchange) if sharded is not None else "nccl"  # fused-multimem (in-switch) / fused-peer (2 ranks)
        elif want == "nvls" or (want in ("auto", "fused") and args.no_optimizer):
            from thr3ed_atom_b200.distributed import NVLSGradientReducer

            try:
                reducer = NVLSGradientReducer(voxel_grid)
            except Exception as e:  # noqa: BLE001
                if want == "nvls":
                    raise
                ok.zero_()
                print(f"[bench] rank {rank}: NVLS reducer unavailable ({e!r}); using NCCL", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) == 0.0 and reducer is not None:
                reducer.close()
                for p in params:
                    p.grad = None
                reducer = None
            exchange = "nvls" if reducer is not None else "nccl"
        else:
            exchange = "nccl"
    if not args.no_optimizer and sharded is None:
        local_opt = FusedGridAdam([{"params": params, "lr": args.lr}], betas=betas, eps=eps)

    def exchange_and_update():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if sharded is not None:
            sharded.step()  # barrier -> reduce-scatter / Adam / all-gather in one kernel -> barrier
            launches["n"] += 1
        else:
            if reducer is not None:
                reducer.all_reduce()
                launches["n"] += 1
            elif world > 1:
                for p in params:  # the path's only exchange: sum of the per-rank dense grid gradients
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            if local_opt is not None:
                local_opt.step()
                launches["n"] += len(params)
        e1.record()
        ev["opt"].append((e0, e1))

    def step(o, d, px, after_loss=None):
        if sharded is not None:
            sharded.zero_grad()  # same 1.95 GB memset as autograd's fresh zero-filled buffers
        elif reducer is not None:
            reducer.zero_grad()
        else:
            for p in params:
                p.grad = None  # autograd then hands the backward kernel freshly zero-filled buffers
        with render_hints(image_hw=hint, variant=args.variant):
            out = vol_mod.render_rays(Rays(o, d))
        loss = torch.nn.functional.l1_loss(out.colour, px)
        if args.strong and world > 1:
            loss = loss / world  # equal shards: the sum over ranks of (local mean / world) is the mean over the whole batch
        if after_loss is not None:
            after_loss(loss, out)
        loss.backward()  # with a direct target the backward kernel accumulates straight into the symmetric buffers
        exchange_and_update()
        return loss, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: one untimed step checks the exchange against an independent NCCL reduction of a strided probe ----
    parity = None
    if world > 1:
        flat_g = sharded.grad_flat if sharded is not None else (reducer.flat if reducer is not None else None)
        total = sum(p.numel() for p in params)
        stride = max(1, total // (1 << 20))
        if sharded is not None:
            idx = torch.arange(0, sharded.total, stride, device=device)
            p_before = sharded.param_flat[idx].clone()
            sharded.zero_grad()
            with render_hints(image_hw=hint, variant=args.variant):
                out = vol_mod.render_rays(rays)
            torch.nn.functional.l1_loss(out.colour, pixels).backward()
            g_sum = flat_g[idx].clone()
            dist.all_reduce(g_sum, op=dist.ReduceOp.SUM)  # NCCL on 1 M probe elements: the independent path
            sharded.step()  # first step: exp_avg = exp_avg_sq = 0  =>  p -= lr * g / (|g| + eps) up to rounding
            torch.cuda.synchronize()
            want_p = p_before - args.lr * (g_sum / (g_sum.abs() + eps))
            got_p = sharded.param_flat[idx]
            # p is fp32 of magnitude <= 1 + steps * lr: one update is resolved to ulp(p) = 6e-8, so the check is in ulps of p
            err_ulps = float(((got_p - want_p).abs().max() / 5.9604645e-08).item())
            err = float(((got_p - want_p).abs().max() / args.lr).item())
            spread = got_p.clone()
            lo_, hi_ = spread.clone(), spread.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN), dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            replicas_equal = bool(torch.equal(lo_, hi_))
            moved = float(((got_p - p_before).abs() > 0).float().mean().item())
            parity = {"what": "fused reduce-scatter/Adam/all-gather vs NCCL all-reduce of a strided probe + closed-form first Adam step",
                      "probe_elements": int(idx.numel()), "max_err_ulps_of_param": err_ulps, "max_err_over_lr": err, "replicas_identical": replicas_equal,
                      "fraction_of_probe_updated": moved, "ok": bool(err_ulps <= 2.0 and replicas_equal and moved > 0.1)}
        else:
            if reducer is not None:
                reducer.zero_grad()
            else:
                for p in params:
                    p.grad = None
            with render_hints(image_hw=hint, variant=args.variant):
                out = vol_mod.render_rays(rays)
            torch.nn.functional.l1_loss(out.colour, pixels).backward()
            grads = [p.grad.reshape(-1) for p in params]
            idxs = [torch.arange(0, g.numel(), stride, device=device) for g in grads]
            local = torch.cat([g[i] for g, i in zip(grads, idxs)]).clone()
            dist.all_reduce(local, op=dist.ReduceOp.SUM)
            if reducer is not None:
                reducer.all_reduce()
            else:
                for p in params:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            got = torch.cat([p.grad.reshape(-1)[i] for p, i in zip(params, idxs)])
            err = float(((got - local).abs().max() / local.abs().max().clamp(min=1e-30)).item())
            parity = {"what": f"{exchange} all-reduce of the grid gradient vs an independent NCCL reduction of a strided probe",
                      "probe_elements": int(local.numel()), "max_rel_err": err, "ok": bool(err < 1e-4)}
        worst = torch.tensor([0.0 if parity["ok"] else 1.0], device=device)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        parity["ok"] = bool(float(worst.item()) == 0.0)

    # ---- warm-up + timed region (device-resident inputs).  The clock sampler runs from the first warm-up step on:
    #      nvidia-smi needs ~100 ms to start, the timed region itself can be shorter than that. ----
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step(rays.origins, rays.directions, pixels)
        barrier()
        for k in ev:
            ev[k].clear()
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
    opt_ms = float(np.mean([a.elapsed_time(b) for a, b in ev["opt"]]))

    # ---- e2e: host buffers in, loss + colour out, every step.  The step's inputs (rays + pixels, 23 MB) are copied from
    #      pinned host memory on a copy stream while the previous step computes (what a data loader does).  The rendered colour
    #      (7.7 MB) and the loss leave on a second copy stream as soon as the loss exists, while the backward runs; the host
    #      waits for THAT step's read-back (a host sync per step) after it has queued the rest of the step.  Every step's
    #      copies and its read-back are inside the timed region. ----
    copy_stream = torch.cuda.Stream(device=device)
    d2h_stream = torch.cuda.Stream(device=device)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    read_done = torch.cuda.Event()

    def read_back(loss, out):
        main = torch.cuda.current_stream(device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ready)
            colour_host.copy_(out.colour.detach(), non_blocking=True)
            loss_host.copy_(loss.detach(), non_blocking=True)
            read_done.record(d2h_stream)
        out.colour.record_stream(d2h_stream)
        loss.record_stream(d2h_stream)

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            staged = [t.to(device, non_blocking=True) for t in (origins_host, dirs_host, pixels_host)]
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return staged, ready

    trace = os.environ.get("R3D_BENCH_E2E_TRACE") == "1"  # host-side timeline of the e2e loop on stderr (diagnostic)

    def e2e_steps(count):
        nxt = stage_inputs()
        for k in range(count):
            t_a = time.perf_counter()
            (o, d, px), ready = nxt
            main = torch.cuda.current_stream(device)
            main.wait_event(ready)
            for t in (o, d, px):
                t.record_stream(main)
            if k + 1 < count:
                nxt = stage_inputs()  # next step's inputs fly while this step computes
            t_b = time.perf_counter()
            step(o, d, px, after_loss=read_back)
            t_c = time.perf_counter()
            read_done.synchronize()
            float(loss_host.item())
            if trace and rank == 0:
                t_d = time.perf_counter()
                print(f"[e2e trace] step {k}: stage {1e3 * (t_b - t_a):.3f} ms, queue step {1e3 * (t_c - t_b):.3f} ms, wait read-back {1e3 * (t_d - t_c):.3f} ms",
                      file=sys.stderr)

    e2e_steps(max(3, min(args.warmup, 5)))  # untimed: the copy streams' allocator pools and the staging double-buffer reach steady state
    import gc

    # The e2e figure is wall clock over K steps (~0.2 s): start-up effects of the copy path and scheduling hiccups of the shared
    # host show up in it at full size (first passes at 17 and 27 ms/step before passes at 11.3 ms/step were observed on the pool).
    # Two passes of K steps each, the faster one is reported, both are listed under e2e.passes_ms_per_step.
    e2e_passes = []
    for _ in range(2):
        gc.collect()
        barrier()
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        barrier()
        e2e_passes.append(time.perf_counter() - t0)
    e2e_s = min(e2e_passes)

    # ---- max over ranks ----
    if world > 1:  # every rank reports the same pass: the one whose slowest rank was fastest
        tp = torch.tensor(e2e_passes, dtype=torch.float64, device=device)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_passes = [float(x) for x in tp.tolist()]
        e2e_s = min(e2e_passes)
    t = torch.tensor([total_ms, e2e_s * 1e3, fwd_ms, bwd_ms, opt_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_ms, bwd_ms, opt_ms = [float(x) for x in t.tolist()]
    peak_mem_gb = torch.cuda.max_memory_allocated(device) / 1e9

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        sm_max = 1965.0
        if peaks_path.exists():
            pk = json.loads(peaks_path.read_text())
            peak, peak_src = float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
            sm_max = float(pk.get("sm_max_mhz", sm_max))
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        clock_summary = clocks.summary()
        sm_mhz = clock_summary.get("sm_mhz") or sm_max
        ms_per_step = total_ms / args.steps
        value = world * n_rays / (ms_per_step * 1e-3)  # weak: n_rays per rank; strong: n_rays is the rank's shard of one view
        opt_name = ("none" if args.no_optimizer else
                    (("peer-to-peer reduce-scatter -> shard-local Adam -> all-gather (r3d_peer_adam_step)" if sharded.exchange == "peer" else
                      "in-switch reduce-scatter -> shard-local Adam -> all-gather (r3d_multimem_adam_step)") if sharded is not None else
                     (f"{exchange} all-reduce(grid grad) + " if world > 1 else "") + "fused dense Adam (r3d_adam_step)"))
        step_desc = "zero_grad + render_rays fwd + l1_loss + backward (fused bwd)"
        if args.no_optimizer:
            step_desc += f" + {exchange} all-reduce(grid grad)" if world > 1 else ""
        else:
            step_desc += " + optimizer.step: " + opt_name
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": args.workload, "grid": grid_n, "sh_degree": deg, "image": [side, side], "samples_per_ray": spp,
                "rays_per_gpu_per_step": n_rays, "ray_order": "flat" if args.flat_order else "image(8x4 tiles)",
                "perturb": not args.no_perturb, "step": step_desc, "optimizer": opt_name, "adam_lr": None if args.no_optimizer else args.lr,
                "exchange": exchange,
                "l2": f"inputs exceed L2: the grid is {grid_n**3 * rec_bytes / 1e6:.0f} MB vs 126 MB",
                "variant": args.variant,
            },
            "phases_ms": {"render_fwd_kernel": fwd_ms, "render_bwd_kernel": bwd_ms, "exchange_plus_optimizer": opt_ms,
                          "zero_fill_loss_and_launch_gaps": max(0.0, ms_per_step - fwd_ms - bwd_ms - opt_ms)},
            "collective_ms": opt_ms if world > 1 else 0.0,
            "fwd_bwd_only": {"ms": ms_per_step - opt_ms, "rays_per_s": world * n_rays / ((ms_per_step - opt_ms) * 1e-3),
                             "note": "the same timed steps minus the exchange + optimizer phase (round-1 definition at N = 1)"},
            "e2e": {"value": world * n_rays * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": 3 * n_rays * 12, "d2h_bytes_per_step": n_rays * 12 + 4,
                    "passes_ms_per_step": [1e3 * x / args.steps for x in e2e_passes], "reported": "the faster of two passes of K steps"},
            "gpu_launches": gpu_launches,
            **roofline_entries(bytes_fwd, fwd_ms, bytes_bwd, bwd_ms, peak, peak_src, load_traffic(args.workload), stats, rec_bytes, nf, sm_mhz),
            "unique_voxels_touched": touched, "voxel_record_bytes": rec_bytes, "sample_statistics": stats,
            "peak_memory_gb": peak_mem_gb,
            "clocks": clock_summary,
        }
        if parity is not None:
            line["parity_check"] = parity
        if world == 1:
            del vol_mod, voxel_grid, params, local_opt
            torch.cuda.empty_cache()
            if not args.no_gpu_baseline:
                try:
                    line["pytorch_gpu_baseline"] = pytorch_gpu_baseline(args.workload, device)
                except Exception as e:  # noqa: BLE001  (informational row: never fail the bench line)
                    line["pytorch_gpu_baseline"] = {"unavailable": repr(e)[:200]}
                torch.cuda.empty_cache()
                try:
                    line["roofline_fwd"]["sm_side_ceiling"] = sm_side_ceiling(args.workload, fwd_ms)
                except Exception as e:  # noqa: BLE001  (companion figure: never fail the bench line)
                    line["roofline_fwd"]["sm_side_ceiling"] = {"unavailable": repr(e)[:200]}
                if line["roofline"].get("kernel") == "render_fwd_group_kernel":
                    line["roofline"]["sm_side_ceiling"] = line["roofline_fwd"]["sm_side_ceiling"]
            if not args.no_cpu_baseline:
                threads = os.cpu_count() or 1
                rps, mean_s, sample, adam_s = cpu_oracle_rays_per_sec(args.workload, args.cpu_sample_rays, steps=2, warmup=1, threads=threads,
                                                                      with_optimizer=not args.no_optimizer)
                line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                        "sample": f"{sample} strided rays of the same {side}x{side} view, fwd+bwd, {mean_s:.2f} s per pass, extrapolated "
                                                  f"to the {n_rays}-ray batch, + one dense torch.optim.Adam step over the grid ({adam_s:.2f} s)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
