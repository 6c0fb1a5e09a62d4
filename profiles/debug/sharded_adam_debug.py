"""2-GPU debug driver: where does NVLSShardedAdam leave torch.optim.Adam's trajectory?  (torchrun --nproc-per-node 2)"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
from helpers import CASES, build_inputs, make_cuda_config, make_cuda_grid
from thr3ed_atom_b200.distributed import NVLSShardedAdam, all_reduce_grid_gradients, shard_rays
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
case = CASES["deg2_16cube"]; inp = build_inputs(case)
rays = Rays(torch.from_numpy(inp["origins"]).to(dev), torch.from_numpy(inp["directions"]).to(dev))
gc = torch.from_numpy(inp["grad_colour"]).to(dev)
shard = shard_rays(rays, gc); lr = 0.03
def local_backward(grid):
    out = render_sh_voxel_grid(grid, shard.rays, make_cuda_config(case)); (out.colour * shard.pixels).sum().backward()
grid_a = make_cuda_grid(case, inp, dev); opt_a = torch.optim.Adam([{"params": list(grid_a.parameters()), "lr": lr}], betas=(0.9, 0.999))
grid_b = make_cuda_grid(case, inp, dev); opt_b = NVLSShardedAdam(grid_b, lr=lr, betas=(0.9, 0.999))
shardn = opt_b.state["exp_avg"].numel()
PHASE1 = os.environ.get("PHASE1", "1") == "1"
if PHASE1:
    full = shard_rays(rays, gc, rank=0, world_size=1)
    opt_a.zero_grad()
    if rank == 0:
        out = render_sh_voxel_grid(grid_a, full.rays, make_cuda_config(case)); (out.colour * full.pixels).sum().backward()
    else:
        for p in grid_a.parameters():
            p.grad = torch.zeros_like(p)
    opt_b.zero_grad()
    for pa, pb in zip(grid_a.parameters(), grid_b.parameters()):
        pb.grad.copy_(pa.grad if rank == 0 else torch.zeros_like(pa))
    all_reduce_grid_gradients(grid_a)
    opt_a.step(); opt_b.step(); torch.cuda.synchronize()
    ms = [torch.zeros_like(opt_b.state["exp_avg"]) for _ in range(world)]; vs = [torch.zeros_like(opt_b.state["exp_avg"]) for _ in range(world)]
    dist.all_gather(ms, opt_b.state["exp_avg"]); dist.all_gather(vs, opt_b.state["exp_avg_sq"])
    m_flat, v_flat = torch.cat(ms)[: opt_b.total], torch.cat(vs)[: opt_b.total]
    off = 0
    for name, pa, pb in zip(("dens", "feat"), grid_a.parameters(), grid_b.parameters()):
        n = pa.numel(); st = opt_a.state[pa]
        rel = lambda x, y: float((x - y).norm() / y.norm().clamp(min=1e-30))
        if rank == 0:
            print(f"phase1 {name}: m rel {rel(m_flat[off:off+n].view_as(pa), st['exp_avg']):.2e} v rel {rel(v_flat[off:off+n].view_as(pa), st['exp_avg_sq']):.2e} |p_a-p_b|max {float((pa-pb).abs().max()):.2e} "
                  f"grad_a norm {float(pa.grad.norm()):.4e} m_b norm {float(m_flat[off:off+n].norm()):.4e} m_t norm {float(st['exp_avg'].norm()):.4e}", flush=True)
        off += (n + 3) // 4 * 4
for step in range(3):
    prev = [p.detach().clone() for p in grid_b.parameters()]
    opt_a.zero_grad(); local_backward(grid_a); all_reduce_grid_gradients(grid_a)
    opt_b.zero_grad(); local_backward(grid_b)
    summed = opt_b.grad_flat.clone(); dist.all_reduce(summed)
    opt_a.step(); opt_b.step(); torch.cuda.synchronize()
    # gather the sharded state
    ms = [torch.zeros_like(opt_b.state["exp_avg"]) for _ in range(world)]; vs = [torch.zeros_like(opt_b.state["exp_avg"]) for _ in range(world)]
    dist.all_gather(ms, opt_b.state["exp_avg"]); dist.all_gather(vs, opt_b.state["exp_avg_sq"])
    m_flat, v_flat = torch.cat(ms)[: opt_b.total], torch.cat(vs)[: opt_b.total]
    off = 0
    for name, pa, pb, p0 in zip(("dens", "feat"), grid_a.parameters(), grid_b.parameters(), prev):
        n = pa.numel(); st = opt_a.state[pa]
        g = summed[off:off + n].view_as(pa)
        rel = lambda x, y: float((x - y).norm() / y.norm().clamp(min=1e-30))
        ua, ub = pa.detach() - p0, pb.detach() - p0
        if rank == 0:
            print(f"step {step} {name}: grad rel {rel(g, pa.grad):.2e}  m rel {rel(m_flat[off:off+n].view_as(pa), st['exp_avg']):.2e}  "
                  f"v rel {rel(v_flat[off:off+n].view_as(pa), st['exp_avg_sq']):.2e}  update rel {rel(ub, ua):.2e}  "
                  f"outliers {float(((ua-ub).abs() > 1e-2*lr).float().mean()):.4f}  |p_a-p_b|max {float((pa-pb).abs().max()):.2e}", flush=True)
            d = (ua - ub).abs().reshape(-1); k = int(d.argmax())
            print(f"    worst elem {k}: ua {float(ua.reshape(-1)[k]):.3e} ub {float(ub.reshape(-1)[k]):.3e} g {float(pa.grad.reshape(-1)[k]):.3e} "
                  f"m_t {float(st['exp_avg'].reshape(-1)[k]):.3e} m_b {float(m_flat[off+k]):.3e} v_t {float(st['exp_avg_sq'].reshape(-1)[k]):.3e} v_b {float(v_flat[off+k]):.3e}", flush=True)
        off += (n + 3) // 4 * 4
    with torch.no_grad():
        for pa, pb in zip(grid_a.parameters(), grid_b.parameters()):
            pa.copy_(pb)
dist.destroy_process_group()
