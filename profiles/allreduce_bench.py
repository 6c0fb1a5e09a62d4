#!/usr/bin/env python
"""All-reduce of a grid-gradient-sized fp32 buffer (256^3 x 28 + 256^3 floats = 1.95 GB): NCCL vs NVLS multimem
(torch symmetric memory) on the GPUs of one box.  torchrun --nproc-per-node N profiles/allreduce_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 256**3 * 29
    out = {"world": world, "bytes": n * 4}
    x = torch.ones(n, device=dev)
    ms = timed(lambda: dist.all_reduce(x))
    out["nccl_ms"] = ms
    out["nccl_algbw_gbs"] = n * 4 / ms / 1e6
    del x
    try:
        import torch.distributed._symmetric_memory as symm_mem

        group = dist.group.WORLD
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, group.group_name)
        out["multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
        t.fill_(1.0)
        torch.cuda.synchronize()
        dist.barrier()
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", group.group_name)
        torch.cuda.synchronize()
        out["multimem_correct"] = bool(torch.all(t[:1000] == world).item()) and bool(torch.all(t[-1000:] == world).item())
        ms = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", group.group_name))
        out["multimem_ms"] = ms
        out["multimem_algbw_gbs"] = n * 4 / ms / 1e6
        ms = timed(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", group.group_name))
        out["two_shot_ms"] = ms
        # the hand-written in-switch all-reduce of this repo (csrc/r3d_comm.cu), bracketed by the handle's barriers
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from thr3ed_atom_b200 import _kernels

        mc = int(hdl.multicast_ptr)
        t.fill_(1.0)
        torch.cuda.synchronize()
        dist.barrier()
        hdl.barrier(channel=0)
        _kernels.multimem_all_reduce(mc, n, rank, world, dev)
        hdl.barrier(channel=1)
        torch.cuda.synchronize()
        out["r3d_multimem_correct"] = bool(torch.all(t[:1000] == world).item()) and bool(torch.all(t[-1000:] == world).item()) and bool(torch.all(t[n // 2 - 500 : n // 2 + 500] == world).item())
        for blocks in (0, 148, 592, 1184):
            def run(blocks=blocks):
                hdl.barrier(channel=0)
                _kernels.multimem_all_reduce(mc, n, rank, world, dev, blocks)
                hdl.barrier(channel=1)
            ms = timed(run)
            out[f"r3d_multimem_ms_blocks{blocks}"] = ms
            out[f"r3d_multimem_algbw_gbs_blocks{blocks}"] = n * 4 / ms / 1e6
    except Exception as e:  # noqa: BLE001
        out["symm_mem_error"] = repr(e)[:400]
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
