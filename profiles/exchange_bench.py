#!/usr/bin/env python
"""Time the fused in-switch exchange + optimizer (NVLSShardedAdam.step) on a c3-sized grid for several launch shapes.
    torchrun --nproc-per-node N profiles/exchange_bench.py            (R3D_MULTIMEM_UNROLL=2|4|8 per process)"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from thr3ed_atom_b200.distributed import NVLSShardedAdam

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
g = int(os.environ.get("GRID", "256"))
mod = torch.nn.Module()
mod.d = torch.nn.Parameter(torch.rand(g, g, g, 1, device=dev)); mod.f = torch.nn.Parameter(torch.rand(g, g, g, 28, device=dev))
opt = NVLSShardedAdam(mod, lr=1e-5)
opt.grad_flat.normal_()
res = []
for blocks in [int(x) for x in os.environ.get("BLOCKS", "148,296,592,1184").split(",")]:
    for _ in range(3):
        opt.step(num_blocks=blocks)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step(num_blocks=blocks)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.append({"blocks": blocks, "ms": float(t.item())})
if rank == 0:
    print(json.dumps({"world": world, "unroll": os.environ.get("R3D_MULTIMEM_UNROLL", "4"), "bytes": opt.total * 4, "results": res}), flush=True)
dist.destroy_process_group()
