"""Tensor-level wrappers of the C ABI (``include/r3d_b200.h``).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every function below packs
raw ``data_ptr()`` values into the ABI structs and enqueues the library's kernels on
``torch.cuda.current_stream()``.  Inputs must be fp32 CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from thr3ed_atom_b200 import _abi


def _require_cuda(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200 render path only runs on CUDA tensors "
            "(there is deliberately no CPU fallback)."
        )
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclasses.dataclass
class GridDesc:
    """Everything the kernels need to know about a voxel grid (mirrors ``R3dGrid``)."""

    densities: Tensor  # [W, D, H, 1] contiguous
    features: Tensor  # [W, D, H, stride] contiguous (first num_features channels used)
    num_features: int
    aabb: Tuple[Tuple[float, float], Tuple[float, float], Tuple[float, float]]
    norm_scale: Sequence[float]
    norm_bias: Sequence[float]
    density_scale: float
    density_pre: int
    density_post: int
    density_quads: Optional[Tensor] = None  # derived probe volume (R3dGrid.density_quads), built by build_density_quads

    def to_struct(self) -> _abi.R3dGrid:
        d = _require_cuda(self.densities, "voxel_grid.densities")
        f = _require_cuda(self.features, "voxel_grid.features")
        if not (d.is_contiguous() and f.is_contiguous()):
            raise RuntimeError("voxel grid storage must be contiguous")
        if d.device != f.device:
            raise RuntimeError("densities and features are not on the same device :(")
        k = self.num_features // 3
        deg = int(round(k**0.5)) - 1
        if 3 * (deg + 1) ** 2 != self.num_features:
            raise ValueError(f"number of features ({self.num_features}) is not 3 * (sh_degree + 1) ** 2")
        g = _abi.R3dGrid()
        g.densities, g.features = d.data_ptr(), f.data_ptr()
        g.dims[:] = list(f.shape[:3])
        g.sh_degree, g.num_features, g.feature_stride = deg, self.num_features, f.shape[3]
        g.aabb_min[:] = [float(np.float32(r[0])) for r in self.aabb]
        g.aabb_max[:] = [float(np.float32(r[1])) for r in self.aabb]
        g.norm_scale[:] = [float(s) for s in self.norm_scale]
        g.norm_bias[:] = [float(b) for b in self.norm_bias]
        g.density_scale = float(self.density_scale)
        g.density_pre, g.density_post = self.density_pre, self.density_post
        if self.density_quads is not None:
            q = _require_cuda(self.density_quads, "density_quads")
            if q.device != d.device or not q.is_contiguous() or q.numel() != density_quad_floats(f.shape[:3]):
                raise ValueError("density_quads does not belong to this grid")
            g.density_quads = q.data_ptr()
        return g


@dataclasses.dataclass
class RenderArgs:
    """Per-call render parameters (mirrors ``R3dRenderConfig`` + the ray hints of ``R3dRays``)."""

    num_samples: int
    near: float
    far: float
    perturb: bool = False
    white_bkgd: bool = False
    diffuse: bool = False
    optimized_sampling: bool = False
    jitter: Optional[Tensor] = None  # explicit [N, S] U[0,1) offsets
    rng_seed: int = 0
    ray_bounds: Optional[Tensor] = None  # [N, 2]
    image_hw: Optional[Tuple[int, int]] = None  # rays are a row-major H x W image (coherence hint)
    camera: Optional[Tuple[int, int, float, Sequence[float], Sequence[float]]] = None  # (H, W, focal, R9, t3)
    variant: int = 0
    keep_for_backward: bool = True  # False: no backward pass can follow (torch.no_grad()) -> no sample cache / ballots

    def flags(self) -> int:
        return (
            (_abi.FLAG_PERTURB if self.perturb else 0)
            | (_abi.FLAG_WHITE_BKGD if self.white_bkgd else 0)
            | (_abi.FLAG_DIFFUSE if self.diffuse else 0)
            | (_abi.FLAG_OPTIMIZED_SAMPLING if self.optimized_sampling else 0)
        )


def _pack_call(grid: GridDesc, origins: Optional[Tensor], directions: Optional[Tensor], num_rays: int, args: RenderArgs):
    g = grid.to_struct()
    keep = []  # keeps ctypes objects alive until the call returns
    r = _abi.R3dRays()
    r.num_rays = num_rays
    if args.camera is not None:
        h, w, focal, rot, trans = args.camera
        cam = _abi.R3dCamera()
        cam.height, cam.width, cam.focal = int(h), int(w), float(focal)
        cam.rotation[:] = [float(x) for x in rot]
        cam.translation[:] = [float(x) for x in trans]
        keep.append(cam)
        r.camera = C.pointer(cam)
    else:
        o = _require_cuda(origins, "rays.origins")
        d = _require_cuda(directions, "rays.directions")
        if not (o.is_contiguous() and d.is_contiguous()):
            raise RuntimeError("rays must be contiguous [N, 3] tensors")
        if o.device != grid.features.device:
            raise RuntimeError(f"rays are on {o.device} but the voxel grid is on {grid.features.device}")
        r.origins, r.directions = o.data_ptr(), d.data_ptr()
        if args.image_hw is not None:
            r.tile_height, r.tile_width = int(args.image_hw[0]), int(args.image_hw[1])
    if args.ray_bounds is not None:
        b = _require_cuda(args.ray_bounds, "ray_bounds")
        if tuple(b.shape) != (num_rays, 2) or not b.is_contiguous():
            raise ValueError("ray_bounds must be a contiguous [N, 2] tensor")
        r.bounds = b.data_ptr()
    c = _abi.R3dRenderConfig()
    c.num_samples, c.near, c.far = int(args.num_samples), float(args.near), float(args.far)
    c.flags = args.flags()
    if args.jitter is not None and args.perturb:
        j = _require_cuda(args.jitter, "jitter")
        if tuple(j.shape) != (num_rays, args.num_samples) or not j.is_contiguous():
            raise ValueError(f"jitter must be a contiguous [{num_rays}, {args.num_samples}] tensor")
        c.jitter = j.data_ptr()
    c.rng_seed = int(args.rng_seed) & 0xFFFFFFFFFFFFFFFF
    c.variant = int(args.variant)
    return g, r, c, keep


def density_quad_floats(dims) -> int:
    arr = (C.c_int32 * 3)(*[int(x) for x in dims])
    n = int(_abi.lib().r3d_density_quad_floats(C.byref(arr)))
    if n < 0:
        raise ValueError(f"bad grid dims {tuple(dims)}")
    return n


def build_density_quads(grid: GridDesc, quads: Optional[Tensor] = None) -> Tensor:
    """(Re)build the density quad volume of ``grid`` (see ``R3dGrid.density_quads``) and return it."""
    device = grid.features.device
    if quads is None:
        quads = torch.empty((density_quad_floats(grid.features.shape[:3]),), dtype=torch.float32, device=device)
    g = dataclasses.replace(grid, density_quads=None).to_struct()
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_build_density_quads(C.byref(g), quads.data_ptr(), _stream(device)), "r3d_build_density_quads")
    return quads


def sample_statistics(grid: GridDesc, origins: Optional[Tensor], directions: Optional[Tensor], args: RenderArgs) -> dict:
    """Measurement helper: samples visited / inside the AABB / in-range corner references / contributing samples."""
    device = grid.features.device
    n = origins.shape[0] if args.camera is None else int(args.camera[0]) * int(args.camera[1])
    counters = torch.zeros((4,), dtype=torch.int64, device=device)
    g, r, c, keep = _pack_call(grid, origins, directions, n, args)
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_sample_statistics(C.byref(g), C.byref(r), C.byref(c), counters.data_ptr(), _stream(device)), "r3d_sample_statistics")
    del keep
    v, i, k, s = [int(x) for x in counters.tolist()]
    return {"samples_visited": v, "samples_inside": i, "corner_refs": k, "samples_contributing": s}


def sample_cache_bytes(num_rays: int, num_samples: int) -> int:
    return 16 * num_rays * num_samples


def new_sample_cache(num_rays: int, num_samples: int, device) -> Tensor:
    """Uninitialised ``[S, N, 4]`` buffer for the forward's per-sample (sigmoid(raw) rgb, sigma) records."""
    return torch.empty((num_samples, num_rays, 4), dtype=torch.float32, device=device)


def sample_mask_supported(grid: GridDesc, args: RenderArgs) -> bool:
    """Will ``render_forward`` dispatch the lane-group kernel (the one that writes the per-step contribution ballots), and can
    the backward use them (ReLU density post-activation)?  Mirrors ``fwd_uses_group_kernel`` / ``mask_usable`` in
    ``csrc/r3d_render.cu``; the library refuses a mask it would not write."""
    f = grid.features
    # A/B builds only: the warp-specialised / cell-sorted forward variants write the ballots too, the per-ray and staged ones
    # (bits 2, 4, 8) do not; bit 1 and 128 select backward kernels
    return ((args.variant & (2 | 4 | 8)) == 0 and not args.diffuse and grid.density_post == _abi.POST_RELU and f.shape[3] % 4 == 0
            and f.data_ptr() % 16 == 0 and f.shape[0] * f.shape[1] * f.shape[2] * (f.shape[3] // 4) <= 0xFFFFFFFF)


def new_sample_mask(grid: GridDesc, origins: Optional[Tensor], directions: Optional[Tensor], num_rays: int, args: RenderArgs) -> Tensor:
    """Uninitialised ``[S, warps]`` int32 buffer for the forward's per-step contribution ballots."""
    _, r, _, keep = _pack_call(grid, origins, directions, num_rays, args)
    words = int(_abi.lib().r3d_sample_mask_words(C.byref(r)))
    del keep
    if words < 0:
        _abi.check(1, "r3d_sample_mask_words")
    return torch.empty((args.num_samples, words), dtype=torch.int32, device=grid.features.device)


def render_forward(grid: GridDesc, origins: Optional[Tensor], directions: Optional[Tensor], args: RenderArgs,
                   sample_cache: Optional[Tensor] = None, with_diffuse: bool = False, sample_cache_diffuse: Optional[Tensor] = None,
                   sample_mask: Optional[Tensor] = None):
    """Fused forward render.  Returns ``(colour [N,3], depth [N,1], acc [N,1], disparity [N,1])``.
    ``sample_cache`` (``new_sample_cache``) is filled for the backward pass when given.
    ``with_diffuse``: single-pass specular + diffuse render -- a fifth return value ``colour_diffuse [N,3]`` is the band-0
    image of the same samples (``sample_cache_diffuse``: its per-sample records for the backward)."""
    device = grid.features.device
    n = origins.shape[0] if args.camera is None else int(args.camera[0]) * int(args.camera[1])
    colour_diffuse = torch.empty((n, 3), dtype=torch.float32, device=device) if with_diffuse else None
    colour = torch.empty((n, 3), dtype=torch.float32, device=device)
    depth = torch.empty((n, 1), dtype=torch.float32, device=device)
    acc = torch.empty((n, 1), dtype=torch.float32, device=device)
    disparity = torch.empty((n, 1), dtype=torch.float32, device=device)
    g, r, c, keep = _pack_call(grid, origins, directions, n, args)
    if sample_cache is not None:
        _require_cuda(sample_cache, "sample_cache")
        if tuple(sample_cache.shape) != (args.num_samples, n, 4) or not sample_cache.is_contiguous():
            raise ValueError(f"sample_cache must be a contiguous [{args.num_samples}, {n}, 4] tensor")
    if sample_cache_diffuse is not None:
        _require_cuda(sample_cache_diffuse, "sample_cache_diffuse")
        if not with_diffuse or tuple(sample_cache_diffuse.shape) != (args.num_samples, n, 4) or not sample_cache_diffuse.is_contiguous():
            raise ValueError(f"sample_cache_diffuse needs with_diffuse and a contiguous [{args.num_samples}, {n}, 4] tensor")
    out = _abi.R3dRenderOut(colour.data_ptr(), depth.data_ptr(), acc.data_ptr(), disparity.data_ptr(), _ptr(sample_cache),
                            _ptr(colour_diffuse), _ptr(sample_cache_diffuse), _ptr(sample_mask))
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_render_fwd(C.byref(g), C.byref(r), C.byref(c), C.byref(out), _stream(device)), "r3d_render_fwd")
    del keep
    if with_diffuse:
        return colour, depth, acc, disparity, colour_diffuse
    return colour, depth, acc, disparity


def render_backward(
    grid: GridDesc,
    origins: Optional[Tensor],
    directions: Optional[Tensor],
    args: RenderArgs,
    saved: Tuple[Tensor, Tensor, Tensor],
    grads: Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor], Optional[Tensor]],
    grad_densities: Optional[Tensor],
    grad_features: Optional[Tensor],
    sample_cache: Optional[Tensor] = None,
    diffuse: Optional[Tuple[Tensor, Optional[Tensor], Optional[Tensor]]] = None,
    sample_mask: Optional[Tensor] = None,
) -> None:
    """Fused backward: accumulates into ``grad_densities`` / ``grad_features`` (same layout as the grid).
    ``sample_cache`` must be the buffer the matching forward call filled (else the radiance is re-gathered).
    ``diffuse`` = ``(colour_diffuse, grad_colour_diffuse, sample_cache_diffuse)`` of a single-pass specular + diffuse render."""
    device = grid.features.device
    colour, depth, acc = saved
    n = colour.shape[0]
    g, r, c, keep = _pack_call(grid, origins, directions, n, args)
    colour_d, grad_colour_d, cache_d = diffuse if diffuse is not None else (None, None, None)
    if grad_colour_d is not None:
        grad_colour_d = _require_cuda(grad_colour_d.contiguous(), "grad_output")
    sv = _abi.R3dRenderOut(colour.data_ptr(), depth.data_ptr(), acc.data_ptr(), None, _ptr(sample_cache), _ptr(colour_d), _ptr(cache_d),
                           _ptr(sample_mask))
    gs = [None if t is None else _require_cuda(t.contiguous(), "grad_output") for t in grads]
    go = _abi.R3dRenderOutGrad(*[_ptr(t) for t in gs], _ptr(grad_colour_d))
    for t, ref in ((grad_densities, grid.densities), (grad_features, grid.features)):
        if t is not None and (tuple(t.shape) != tuple(ref.shape) or not t.is_contiguous() or t.dtype != torch.float32):
            raise ValueError("gradient buffers must match the grid storage layout")
    gg = _abi.R3dGridGrad(_ptr(grad_densities), _ptr(grad_features))
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_render_bwd(C.byref(g), C.byref(r), C.byref(c), C.byref(sv), C.byref(go), C.byref(gg), _stream(device)),
            "r3d_render_bwd",
        )
    del keep, gs


def mark_touched_voxels(grid: GridDesc, origins: Optional[Tensor], directions: Optional[Tensor], args: RenderArgs) -> Tensor:
    """uint8 ``[W, D, H]`` bitmap of voxels referenced by in-volume samples (measurement helper)."""
    device = grid.features.device
    n = origins.shape[0] if args.camera is None else int(args.camera[0]) * int(args.camera[1])
    bitmap = torch.zeros(tuple(grid.features.shape[:3]), dtype=torch.uint8, device=device)
    g, r, c, keep = _pack_call(grid, origins, directions, n, args)
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_mark_touched_voxels(C.byref(g), C.byref(r), C.byref(c), bitmap.data_ptr(), _stream(device)), "r3d_mark_touched_voxels")
    del keep
    return bitmap


def cast_rays(height: int, width: int, focal: float, rotation, translation, device) -> Tuple[Tensor, Tensor]:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"cast_rays runs on CUDA only (requested {device})")
    rot = torch.as_tensor(rotation).detach().to("cpu", torch.float32).reshape(9).tolist()
    trans = torch.as_tensor(translation).detach().to("cpu", torch.float32).reshape(3).tolist()
    cam = _abi.R3dCamera()
    cam.height, cam.width, cam.focal = height, width, focal
    cam.rotation[:] = rot
    cam.translation[:] = trans
    origins = torch.empty((height * width, 3), dtype=torch.float32, device=device)
    directions = torch.empty((height * width, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_cast_rays(C.byref(cam), origins.data_ptr(), directions.data_ptr(), _stream(device)), "r3d_cast_rays")
    return origins, directions


def grid_lookup_forward(grid: GridDesc, points: Tensor, want_inside: bool = False):
    device = grid.features.device
    p = _require_cuda(points, "points").contiguous()
    n = p.shape[0]
    out = torch.empty((n, grid.num_features + 1), dtype=torch.float32, device=device)
    inside = torch.empty((n,), dtype=torch.uint8, device=device) if want_inside else None
    g = grid.to_struct()
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_grid_lookup_fwd(C.byref(g), p.data_ptr(), n, out.data_ptr(), _ptr(inside), _stream(device)), "r3d_grid_lookup_fwd")
    return out, inside


def grid_lookup_backward(grid: GridDesc, points: Tensor, grad_out: Tensor, grad_densities: Optional[Tensor], grad_features: Optional[Tensor]) -> None:
    device = grid.features.device
    p = _require_cuda(points, "points").contiguous()
    go = _require_cuda(grad_out, "grad_out").contiguous()
    g = grid.to_struct()
    gg = _abi.R3dGridGrad(_ptr(grad_densities), _ptr(grad_features))
    with torch.cuda.device(device):
        _abi.check(_abi.lib().r3d_grid_lookup_bwd(C.byref(g), p.data_ptr(), p.shape[0], go.data_ptr(), C.byref(gg), _stream(device)), "r3d_grid_lookup_bwd")


def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, *, lr: float, beta1: float, beta2: float,
              eps: float, step: int, grad_scale: float = 1.0) -> None:
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _require_cuda(t, name)
        if not t.is_contiguous() or t.numel() != param.numel():
            raise ValueError(f"{name} must be contiguous and match param")
    device = param.device
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_adam_step(
                param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                lr, beta1, beta2, eps, 1.0 - beta1**step, 1.0 - beta2**step, grad_scale, _stream(device),
            ),
            "r3d_adam_step",
        )


def multimem_all_reduce(multicast_ptr: int, num_floats: int, rank: int, world_size: int, device, num_blocks: int = 0) -> None:
    """Enqueue the in-switch (NVLS) sum all-reduce kernel on the current stream (see ``r3d_multimem_all_reduce``)."""
    device = torch.device(device)
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_multimem_all_reduce(C.c_void_p(multicast_ptr), num_floats, rank, world_size, num_blocks, _stream(device)),
            "r3d_multimem_all_reduce",
        )


def multimem_shard_floats(num_floats: int, world_size: int) -> int:
    n = int(_abi.lib().r3d_multimem_shard_floats(num_floats, world_size))
    if n < 0:
        raise ValueError("flat buffer must be a non-negative multiple of 4 floats")
    return n


def multimem_adam_step(grad_multicast_ptr: int, param_multicast_ptr: int, param_local: Tensor, exp_avg_shard: Tensor, exp_avg_sq_shard: Tensor,
                       rank: int, world_size: int, *, lr: float, beta1: float, beta2: float, eps: float, step: int, grad_scale: float = 1.0,
                       num_blocks: int = 0) -> None:
    """Enqueue the fused reduce-scatter -> shard-local Adam -> all-gather kernel (see ``r3d_multimem_adam_step``)."""
    for name, t in (("param_local", param_local), ("exp_avg_shard", exp_avg_shard), ("exp_avg_sq_shard", exp_avg_sq_shard)):
        _require_cuda(t, name)
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    shard = multimem_shard_floats(param_local.numel(), world_size)
    if exp_avg_shard.numel() != shard or exp_avg_sq_shard.numel() != shard:
        raise ValueError(f"optimizer state shards must hold {shard} floats")
    device = param_local.device
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_multimem_adam_step(
                C.c_void_p(grad_multicast_ptr), C.c_void_p(param_multicast_ptr), param_local.data_ptr(), exp_avg_shard.data_ptr(),
                exp_avg_sq_shard.data_ptr(), param_local.numel(), rank, world_size, lr, beta1, beta2, eps, 1.0 - beta1**step,
                1.0 - beta2**step, grad_scale, num_blocks, _stream(device),
            ),
            "r3d_multimem_adam_step",
        )


def peer_adam_step(grad_ptrs, param_ptrs, num_floats: int, exp_avg_shard: Tensor, exp_avg_sq_shard: Tensor, rank: int, world_size: int, *,
                   lr: float, beta1: float, beta2: float, eps: float, step: int, grad_scale: float = 1.0, num_blocks: int = 0) -> None:
    """Enqueue the fused reduce-scatter -> shard-local Adam -> all-gather kernel over peer-to-peer pointers (``r3d_peer_adam_step``).
    ``grad_ptrs`` / ``param_ptrs``: one device address per rank (symmetric-memory ``buffer_ptrs``)."""
    if len(grad_ptrs) != world_size or len(param_ptrs) != world_size:
        raise ValueError("one gradient and one parameter pointer per rank")
    for name, t in (("exp_avg_shard", exp_avg_shard), ("exp_avg_sq_shard", exp_avg_sq_shard)):
        _require_cuda(t, name)
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    shard = multimem_shard_floats(num_floats, world_size)
    if exp_avg_shard.numel() != shard or exp_avg_sq_shard.numel() != shard:
        raise ValueError(f"optimizer state shards must hold {shard} floats")
    device = exp_avg_shard.device
    g_arr = (C.c_void_p * world_size)(*[int(x) for x in grad_ptrs])
    p_arr = (C.c_void_p * world_size)(*[int(x) for x in param_ptrs])
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_peer_adam_step(
                g_arr, p_arr, exp_avg_shard.data_ptr(), exp_avg_sq_shard.data_ptr(), num_floats, rank, world_size, lr, beta1, beta2, eps,
                1.0 - beta1**step, 1.0 - beta2**step, grad_scale, num_blocks, _stream(device),
            ),
            "r3d_peer_adam_step",
        )


def has_ab_variants() -> bool:
    """Was the loaded library built with -DR3D_AB_VARIANTS (measurement-only kernel variants selectable)?"""
    return bool(_abi.lib().r3d_has_ab_variants())


def sample_ray_batch(rotations: Tensor, translations: Tensor, images: Optional[Tensor], height: int, width: int, focal: float, batch: int,
                     tile: Tuple[int, int] = (1, 1), seed: int = 0, want_indices: bool = False):
    """``batch`` training rays (+ target pixels) from random pixels / pixel tiles of ``V`` posed views (``r3d_sample_ray_batch``).
    ``rotations [V, 3, 3]``, ``translations [V, 3]`` and ``images [V, H, W, 3]`` are fp32 CUDA tensors; ``tile = (width, height)``."""
    rot = _require_cuda(rotations, "rotations").reshape(-1, 9).contiguous()
    tr = _require_cuda(translations, "translations").reshape(-1, 3).contiguous()
    v = rot.shape[0]
    if tr.shape[0] != v:
        raise ValueError("rotations and translations disagree on the number of views")
    device = rot.device
    if images is not None:
        images = _require_cuda(images, "images")
        if tuple(images.shape) != (v, height, width, 3) or not images.is_contiguous():
            raise ValueError(f"images must be a contiguous [{v}, {height}, {width}, 3] tensor")
    views = _abi.R3dViewSet(rot.data_ptr(), tr.data_ptr(), _ptr(images), v, int(height), int(width), float(focal))
    origins = torch.empty((batch, 3), dtype=torch.float32, device=device)
    directions = torch.empty((batch, 3), dtype=torch.float32, device=device)
    pixels = torch.empty((batch, 3), dtype=torch.float32, device=device) if images is not None else None
    indices = torch.empty((batch,), dtype=torch.int64, device=device) if want_indices else None
    with torch.cuda.device(device):
        _abi.check(
            _abi.lib().r3d_sample_ray_batch(C.byref(views), batch, int(tile[0]), int(tile[1]), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                            origins.data_ptr(), directions.data_ptr(), _ptr(pixels), _ptr(indices), _stream(device)),
            "r3d_sample_ray_batch",
        )
    return origins, directions, pixels, indices
