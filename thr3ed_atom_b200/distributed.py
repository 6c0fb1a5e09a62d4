"""Multi-GPU data parallelism of the render path: one process per GPU, rays sharded, ONE collective.

Rays are independent, so the forward pass needs no exchange; the backward pass needs exactly one --
the sum over ranks of the dense grid gradient (the reference itself is
single-device, the hook point is between ``total_loss.backward()`` and ``optimizer.step()`` at
reference modules/trainers.py:339-341).  Grid parameters are replicated on every rank.

    shard = shard_rays(rays, pixels)                  # contiguous block per rank (keeps image tiles coherent)
    out = vol_mod.render_rays(shard.rays)
    loss = l1_loss(out.colour, shard.pixels) * shard.loss_weight   # so that the sum over ranks is the global mean
    loss.backward()
    all_reduce_grid_gradients(vol_mod.thre3d_repr)    # NCCL over NVLink / NVSwitch
    optimizer.step()

``NVLSGradientReducer`` is the second implementation of the same exchange: the gradient lives in symmetric memory and is
summed inside the NVSwitch by this repo's own multimem kernel (``csrc/r3d_comm.cu``).
"""
from __future__ import annotations

import dataclasses
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """``[start, end)`` of the contiguous block owned by ``rank``; blocks differ by at most one item."""
    base, extra = divmod(num_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


@dataclasses.dataclass
class RayShard:
    rays: Rays
    pixels: Optional[Tensor]
    start: int
    end: int
    total: int

    @property
    def loss_weight(self) -> float:
        """A per-rank *mean* loss times this weight, summed over ranks, equals the mean over all rays."""
        return (self.end - self.start) / self.total if self.total else 0.0


def shard_rays(rays: Rays, pixels: Optional[Tensor] = None, rank: Optional[int] = None, world_size: Optional[int] = None, group=None) -> RayShard:
    """This rank's contiguous slice of a flat ray batch (and of the matching pixels)."""
    r, w = _world(group)
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    n = len(rays)
    start, end = shard_bounds(n, rank, world_size)
    return RayShard(rays[start:end], None if pixels is None else pixels[start:end], start, end, n)


def shard_views(num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None, group=None) -> List[int]:
    """Whole images per rank (weak scaling: every GPU renders full views of the replicated grid)."""
    r, w = _world(group)
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    start, end = shard_bounds(num_views, rank, world_size)
    return list(range(start, end))


def grid_gradients(module: torch.nn.Module) -> List[Tensor]:
    return [p.grad for p in module.parameters() if p.grad is not None]


def all_reduce_grid_gradients(module: torch.nn.Module, group=None, average: bool = False, async_op: bool = False):
    """Sum (or average) the grid gradient over all ranks, in place.  The padded feature gradient and the density
    gradient are reduced as they are stored (no flattening copy: at 256^3 deg 2 the message is 2.1 GB).
    Returns the list of async work handles when ``async_op`` is set."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    world_size = dist.get_world_size(group)
    handles = []
    for grad in grid_gradients(module):
        if average:
            grad.div_(world_size)
        work = dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            handles.append(work)
    return handles


def broadcast_grid(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s grid values (replicas must be identical)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=group)


class NVLSGradientReducer:
    """Grid gradients kept in symmetric memory and summed across ranks *inside the NVSwitch*.

    NCCL's ring all-reduce moves ``2 (n-1)/n`` x the gradient bytes per GPU and direction.  With NVLink-switch multicast
    (NVLS) every rank instead pulls the already-reduced sum of its 1/n slice (``multimem.ld_reduce``) and broadcasts it
    back (``multimem.st``): ~1x the bytes per direction (kernel: ``csrc/r3d_comm.cu``).  For that the gradient has to live
    in memory that is mapped into one multicast object on all ranks, so this class owns the gradient storage:

        reducer = NVLSGradientReducer(voxel_grid)      # p.grad of every grid parameter now aliases symmetric memory
        for step in ...:
            reducer.zero_grad()                        # instead of optimizer.zero_grad() (the buffers must stay)
            loss = ...; loss.backward()                # the backward kernel accumulates straight into the buffers
            reducer.all_reduce()                       # barrier -> in-switch reduction -> barrier, on the current stream
            optimizer.step()

    Raises at construction if symmetric memory / multicast is unavailable (callers fall back to
    ``all_reduce_grid_gradients``).
    """

    def __init__(self, module: torch.nn.Module, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from thr3ed_atom_b200.thre3d_reprs import renderers

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NVLSGradientReducer needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world_size = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise RuntimeError("no trainable grid parameters")
        device = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]  # every region stays 16-byte aligned
        self.total = sum(sizes)
        self.flat = symm_mem.empty(self.total, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.flat, self.group.group_name)
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if self.multicast_ptr == 0:
            raise RuntimeError("no NVLS multicast mapping for the gradient buffer on this system")
        self.flat.zero_()
        self._registry = renderers._direct_grad_targets
        self._keys = []
        offset = 0
        for p, size in zip(self.params, sizes):
            view = self.flat[offset : offset + p.numel()].view_as(p)
            p.grad = view
            self._registry[p.data_ptr()] = view
            self._keys.append(p.data_ptr())
            offset += size
        self.device = device

    def zero_grad(self) -> None:
        self.flat.zero_()

    def all_reduce(self, num_blocks: int = 0) -> None:
        from thr3ed_atom_b200 import _kernels

        self.handle.barrier(channel=0)  # every rank has finished accumulating its gradient
        _kernels.multimem_all_reduce(self.multicast_ptr, self.total, self.rank, self.world_size, self.device, num_blocks)
        self.handle.barrier(channel=1)  # every slice has been reduced and broadcast

    def close(self) -> None:
        for k in self._keys:
            self._registry.pop(k, None)
        self._keys = []
