"""The SH voxel-grid render procedure -- the drop-in for the reference's
``thre3d_atom/thre3d_reprs/renderers.py`` (``RenderProcedure`` :25, ``SHVoxGridRenderConfig`` :28-45,
``render_sh_voxel_grid`` :48-102).

The reference binds three ``functools.partial`` stages (sampler, point processor, accumulator) and
lets autograd differentiate ~150 ATen launches over materialised ``[N*S, F+1]`` tensors.  Here the
same procedure is ONE fused CUDA kernel forward and ONE fused kernel backward behind the C ABI
(``include/r3d_b200.h``); per-ray state lives in registers, nothing of size ``N*S`` is ever stored.

``SHVoxGridRenderConfig`` keeps the reference's fields, order and defaults (it is pickled by
qualified name and rebuilt from ``dataclasses.asdict`` when checkpoints are loaded).
"""
from __future__ import annotations

import contextlib
import dataclasses
import os
import threading
from typing import Any, Callable, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from thr3ed_atom_b200 import _kernels
from thr3ed_atom_b200.rendering.volumetric.accumulate import density2occupancy_pb
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays, RenderOut, assert_flat_rays
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid
from thr3ed_atom_b200.utils.constants import EXTRA_ACCUMULATED_WEIGHTS, EXTRA_DISPARITY
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

RenderConfig = Any
RenderProcedure = Callable[[Module, Rays, RenderConfig, Optional[int]], RenderOut]


@dataclasses.dataclass
class SHVoxGridRenderConfig:
    # probing
    num_samples_per_ray: int
    camera_bounds: CameraBounds
    perturb_sampled_points: bool = True
    optimized_sampling: bool = False

    # accumulation
    density2occupancy: Callable[[Tensor, Tensor], Tensor] = density2occupancy_pb
    radiance_hdr_tone_map: Callable[[Tensor], Tensor] = torch.sigmoid
    stochastic_density_noise_std: float = 0.0
    white_bkgd: bool = False

    # misc render modes (consumed by the callers)
    render_diffuse: bool = False
    render_num_samples_per_ray: int = 1024
    parallel_rays_chunk_size: int = 32768


# ---------------------------------------------------------------------------------------------
# per-call hints that have no slot in the reference's config dataclass
# ---------------------------------------------------------------------------------------------
_hints = threading.local()


@contextlib.contextmanager
def render_hints(*, image_hw: Optional[Tuple[int, int]] = None, jitter: Optional[Tensor] = None, rng_seed: Optional[int] = None, variant: Optional[int] = None):
    """Optional side-channel for ``render_sh_voxel_grid`` calls made inside the ``with`` block.

    image_hw: the flat rays are a row-major ``H x W`` image -> threads are mapped to 8x4 pixel tiles
              (pure scheduling hint; per-ray results are unchanged).
    jitter:   explicit ``[N, S]`` stratified offsets in ``[0, 1)`` (what the reference draws with
              ``torch.rand``) instead of the in-kernel counter-based RNG; used for parity tests.
    rng_seed: seed of the in-kernel RNG (default: drawn from torch's CPU generator, so
              ``torch.manual_seed`` makes renders reproducible).
    variant:  kernel variant selector (A/B measurement only).
    """
    previous = getattr(_hints, "value", None)
    _hints.value = {"image_hw": image_hw, "jitter": jitter, "rng_seed": rng_seed, "variant": variant}
    try:
        yield
    finally:
        _hints.value = previous


def _current_hints() -> dict:
    return getattr(_hints, "value", None) or {}


# ---------------------------------------------------------------------------------------------
# autograd binding of the fused kernels
# ---------------------------------------------------------------------------------------------
# Optional direct gradient targets: {data_ptr of a grid storage tensor -> buffer of the same shape}.  When a target is
# registered for a parameter, the backward kernel accumulates straight into it (and autograd receives no gradient for that
# input) -- used by thr3ed_atom_b200.distributed.NVLSGradientReducer to keep the gradient in symmetric memory, where
# the NVSwitch can reduce it in place.
_direct_grad_targets = {}


def register_direct_grad_target(param: Tensor, buffer: Tensor) -> int:
    """Make the backward kernel accumulate the gradient of ``param`` straight into ``buffer`` (same shape).  The entry is
    keyed by the storage address the kernels see and remembers its owner weakly: it is ignored (and dropped) once the
    parameter is gone or has moved to other storage, so a stale address re-used by the caching allocator can never route
    another tensor's gradient into an old buffer."""
    import weakref

    if tuple(buffer.shape) != tuple(param.shape) or buffer.device != param.device:
        raise ValueError("direct gradient target must match the parameter's shape and device")
    key = param.data_ptr()
    _direct_grad_targets[key] = (weakref.ref(param), buffer)
    return key


def unregister_direct_grad_target(key: int) -> None:
    _direct_grad_targets.pop(key, None)


def _lookup_direct_grad_target(storage: Tensor) -> Optional[Tensor]:
    key = storage.data_ptr()
    entry = _direct_grad_targets.get(key)
    if entry is None:
        return None
    owner, buffer = entry[0](), entry[1]
    if owner is None or owner.data_ptr() != key or tuple(owner.shape) != tuple(storage.shape):
        _direct_grad_targets.pop(key, None)  # the registered parameter died or was re-allocated (.to(), stage upscale)
        return None
    return buffer


class _FusedSHVoxGridRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, densities: Tensor, features: Tensor, origins: Tensor, directions: Tensor, grid: VoxelGrid, args: _kernels.RenderArgs,
                with_diffuse: bool = False):
        desc = grid.kernel_desc(densities, features, _probe_volume(grid, densities, args))
        # When a backward pass will follow, the forward keeps (sigmoid(raw) rgb, sigma) of every contributing sample
        # ([S, N, 4] fp32) so that the backward does not gather the 8 corner records a second time.
        # With the cache goes one ballot word per warp and marching step (which rays' samples contributed): a ReLU-field
        # backward then marches by those instead of repeating the inside test and the density gather.
        cache = cache_d = mask = None
        n = origins.shape[0]
        # ctx.needs_input_grad is True for a tunable grid even under torch.no_grad(); the caller records whether autograd is
        # actually recording (args.keep_for_backward), so that inference / chunked no-grad renders store nothing of size N*S
        if args.keep_for_backward and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and n > 0:
            if _kernels.sample_cache_bytes(n, args.num_samples) * (2 if with_diffuse else 1) <= sample_cache_limit_bytes():
                cache = _kernels.new_sample_cache(n, args.num_samples, origins.device)
                cache_d = _kernels.new_sample_cache(n, args.num_samples, origins.device) if with_diffuse else None
                if _kernels.sample_mask_supported(desc, args) and os.environ.get("R3D_SAMPLE_MASK", "1") != "0":
                    mask = _kernels.new_sample_mask(desc, origins, directions, n, args)
        ctx.desc, ctx.args, ctx.cache, ctx.cache_d, ctx.mask, ctx.with_diffuse = desc, args, cache, cache_d, mask, with_diffuse
        ctx.set_materialize_grads(False)  # unused outputs arrive as None instead of zero tensors
        if with_diffuse:
            colour, depth, acc, disparity, colour_d = _kernels.render_forward(desc, origins, directions, args, cache, True, cache_d, mask)
            ctx.save_for_backward(origins, directions, colour, depth, acc, colour_d)
            return colour, depth, acc, disparity, colour_d
        colour, depth, acc, disparity = _kernels.render_forward(desc, origins, directions, args, cache, sample_mask=mask)
        ctx.save_for_backward(origins, directions, colour, depth, acc)
        return colour, depth, acc, disparity

    @staticmethod
    def backward(ctx, g_colour, g_depth, g_acc, g_disparity, g_colour_d=None):
        if ctx.with_diffuse:
            origins, directions, colour, depth, acc, colour_d = ctx.saved_tensors
        else:
            (origins, directions, colour, depth, acc), colour_d = ctx.saved_tensors, None
        desc = ctx.desc
        need_d, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        direct_d = _lookup_direct_grad_target(desc.densities) if need_d else None
        direct_f = _lookup_direct_grad_target(desc.features) if need_f else None
        grad_d = direct_d if direct_d is not None else (torch.zeros_like(desc.densities) if need_d else None)
        grad_f = direct_f if direct_f is not None else (torch.zeros_like(desc.features) if need_f else None)
        if need_d or need_f:
            _kernels.render_backward(
                desc, origins, directions, ctx.args, (colour, depth, acc), (g_colour, g_depth, g_acc, g_disparity), grad_d, grad_f,
                ctx.cache, diffuse=(colour_d, g_colour_d, ctx.cache_d) if (ctx.with_diffuse and g_colour_d is not None) else None,
                sample_mask=ctx.mask,
            )
        ctx.cache = ctx.cache_d = ctx.mask = None  # free the per-sample records as soon as they are consumed
        # gradients accumulated into a registered target are already where they belong
        return (None if direct_d is not None else grad_d), (None if direct_f is not None else grad_f), None, None, None, None, None


def _probe_volume(grid: VoxelGrid, densities: Optional[Tensor], args: _kernels.RenderArgs) -> Optional[Tensor]:
    """Density quad volume for the warp-specialised forward kernel (None when another kernel will run)."""
    if not (args.variant & (32 | 32768 | 65536)) or args.diffuse:
        return None
    return grid.density_quads(densities, fresh=args.keep_for_backward)


def sample_cache_limit_bytes() -> int:
    """Upper bound on the per-call sample cache (``16 * rays * samples`` bytes); above it the backward re-gathers.
    Default 48 GiB (a B200 has 180 GB); override with ``R3D_SAMPLE_CACHE_MAX_BYTES`` (0 disables the cache)."""
    return int(os.environ.get("R3D_SAMPLE_CACHE_MAX_BYTES", 48 * 2**30))


def _validate_config(cfg: SHVoxGridRenderConfig) -> None:
    # the fused kernels implement exactly the reference defaults of the accumulation stage; anything else is refused loudly
    if cfg.density2occupancy is not density2occupancy_pb and getattr(cfg.density2occupancy, "__name__", "") != "density2occupancy_pb":
        raise NotImplementedError("only density2occupancy_pb (1 - exp(-sigma * delta)) is implemented by the fused B200 renderer")
    if cfg.radiance_hdr_tone_map is not torch.sigmoid:
        raise NotImplementedError("only torch.sigmoid is implemented as radiance_hdr_tone_map by the fused B200 renderer")
    if cfg.stochastic_density_noise_std != 0.0:
        raise NotImplementedError("stochastic_density_noise_std != 0 is not implemented by the fused B200 renderer")


def _draw_seed() -> int:
    # consumes torch's default CPU generator => reproducible under torch.manual_seed
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


def make_render_args(cfg: SHVoxGridRenderConfig, **overrides) -> _kernels.RenderArgs:
    hints = _current_hints()
    near, far = cfg.camera_bounds
    args = _kernels.RenderArgs(
        num_samples=int(cfg.num_samples_per_ray),
        near=float(near),
        far=float(far),
        perturb=bool(cfg.perturb_sampled_points),
        white_bkgd=bool(cfg.white_bkgd),
        diffuse=bool(cfg.render_diffuse),
        optimized_sampling=bool(cfg.optimized_sampling),
        jitter=hints.get("jitter"),
        image_hw=hints.get("image_hw"),
        variant=hints.get("variant") or int(os.environ.get("R3D_VARIANT", "0")),
    )
    if args.perturb and args.jitter is None:
        seed = hints.get("rng_seed")
        args.rng_seed = _draw_seed() if seed is None else int(seed)
    args.keep_for_backward = torch.is_grad_enabled()
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def render_sh_voxel_grid(
    voxel_grid: VoxelGrid,
    rays: Rays,
    render_config: SHVoxGridRenderConfig,
    parallel_points_chunk_size: Optional[int] = None,
) -> RenderOut:
    """
    renders an SH-based voxel grid
    Args:
        voxel_grid: the VoxelGrid being rendered
        rays: flat ``[N, 3]`` rays (origins + not-necessarily-unit directions) on the grid's device
        render_config: SHVoxGridRenderConfig
        parallel_points_chunk_size: memory hint of the reference's op-by-op pipeline; the fused kernel keeps
            per-ray state in registers and needs no chunking, so it is accepted and ignored
    Returns: RenderOut(colour [N,3], depth [N,1], extra={disparity [N,1], accumulated_weight [N,1]}),
             differentiable w.r.t. the grid's densities / features when grad mode is on
    """
    assert_flat_rays(rays)
    _validate_config(render_config)
    args = make_render_args(render_config)
    if args.image_hw is not None and args.image_hw[0] * args.image_hw[1] != rays.origins.shape[0]:
        args.image_hw = None  # hint does not describe this batch (e.g. a chunk of an image)
    origins = rays.origins.detach().contiguous()
    directions = rays.directions.detach().contiguous()
    colour, depth, acc, disparity = _FusedSHVoxGridRender.apply(
        voxel_grid.densities, voxel_grid.feature_storage, origins, directions, voxel_grid, args
    )
    return RenderOut(colour=colour, depth=depth, extra={EXTRA_DISPARITY: disparity, EXTRA_ACCUMULATED_WEIGHTS: acc})


def render_sh_voxel_grid_with_diffuse(
    voxel_grid: VoxelGrid,
    rays: Rays,
    render_config: SHVoxGridRenderConfig,
    parallel_points_chunk_size: Optional[int] = None,
) -> Tuple[RenderOut, RenderOut]:
    """Single-pass specular + diffuse render.

    The reference trainer renders every ray batch twice -- ``vol_mod.render_rays(rays)`` and
    ``vol_mod.render_rays(rays, render_diffuse=True)`` (``modules/trainers.py:306-330``) -- to regularise the geometry with
    the view-independent image.  The diffuse radiance is ``C0 * coeff[ch][0]`` (``process.py:59-63``): the ``k = 0`` element
    of the very voxel records the specular render gathers.  This call produces both images from ONE march, ONE 8-corner
    gather per sample and ONE backward pass.  Returns ``(specular RenderOut, diffuse RenderOut)``; depth, disparity and
    accumulated weight are shared (they do not depend on the radiance).

    Semantic difference to two reference calls: both images see the SAME stratified sample positions (the reference draws
    a fresh ``torch.rand`` for each render); with ``perturb_sampled_points=False`` the results are identical.
    ``render_config.render_diffuse`` must be False."""
    assert_flat_rays(rays)
    _validate_config(render_config)
    if render_config.render_diffuse:
        raise ValueError("render_sh_voxel_grid_with_diffuse renders both images; render_config.render_diffuse must be False")
    args = make_render_args(render_config)
    if args.image_hw is not None and args.image_hw[0] * args.image_hw[1] != rays.origins.shape[0]:
        args.image_hw = None
    origins = rays.origins.detach().contiguous()
    directions = rays.directions.detach().contiguous()
    colour, depth, acc, disparity, colour_d = _FusedSHVoxGridRender.apply(
        voxel_grid.densities, voxel_grid.feature_storage, origins, directions, voxel_grid, args, True
    )
    extra = {EXTRA_DISPARITY: disparity, EXTRA_ACCUMULATED_WEIGHTS: acc}
    return RenderOut(colour=colour, depth=depth, extra=extra), RenderOut(colour=colour_d, depth=depth, extra=dict(extra))


def render_sh_voxel_grid_camera(voxel_grid: VoxelGrid, camera_intrinsics, camera_pose, render_config: SHVoxGridRenderConfig) -> RenderOut:
    """Whole-image inference render with IN-KERNEL ray generation (no ray tensors, no chunking):
    the fused form of ``cast_rays`` + ``flatten_rays`` + ``render_sh_voxel_grid`` used by
    ``VolumetricModel.render``.  Not differentiable (the reference's ``render`` runs under no_grad)."""
    _validate_config(render_config)
    height, width, focal = camera_intrinsics
    rot = torch.as_tensor(camera_pose.rotation).detach().to("cpu", torch.float32).reshape(9).tolist()
    trans = torch.as_tensor(camera_pose.translation).detach().to("cpu", torch.float32).reshape(3).tolist()
    args = make_render_args(render_config, camera=(int(height), int(width), float(focal), rot, trans), image_hw=None)
    with torch.no_grad():
        desc = voxel_grid.kernel_desc(density_quads=_probe_volume(voxel_grid, None, dataclasses.replace(args, keep_for_backward=False)))
        colour, depth, acc, disparity = _kernels.render_forward(desc, None, None, args)
    return RenderOut(colour=colour, depth=depth, extra={EXTRA_DISPARITY: disparity, EXTRA_ACCUMULATED_WEIGHTS: acc})
