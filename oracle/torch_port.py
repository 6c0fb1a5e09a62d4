"""fp32 PyTorch-CPU restatement of the reference SH-voxel-grid render path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): checker + CPU baseline, never product.

One flat function per stage, written from the algorithm, each citing the
reference lines (relative to ``/root/reference/``) whose arithmetic it reproduces.  The op
*order* of every fp32 expression is kept identical to the reference so that, on the same
machine, results agree with the reference to the last bit (checked by
``tests/test_oracle_golden.py`` against goldens produced by the reference itself).

Everything is differentiable through autograd, so ``torch.autograd.grad`` on the outputs gives
the reference's backward pass into ``densities`` / ``features``.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ZERO_PLUS = 1e-10  # thre3d_atom/utils/constants.py:7
INFINITY = 1e10  # thre3d_atom/utils/constants.py:8

# thre3d_atom/rendering/volumetric/utils/spherical_harmonics.py:33-50
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (
    1.0925484305920792,
    -1.0925484305920792,
    0.31539156525252005,
    -1.0925484305920792,
    0.5462742152960396,
)
SH_C3 = (
    -0.5900435899266435,
    2.890611442640554,
    -0.4570457994644658,
    0.3731763325901154,
    -0.4570457994644658,
    1.445305721320277,
    -0.5900435899266435,
)


@dataclasses.dataclass
class OracleGrid:
    """Plain description of a reference ``VoxelGrid`` (thre3d_reprs/voxels.py:46-124)."""

    densities: torch.Tensor  # [W, D, H, 1] fp32
    features: torch.Tensor  # [W, D, H, F] fp32, F = 3 * (deg + 1) ** 2, channel-major
    voxel_size: Tuple[float, float, float]
    location: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    density_scale: float = 1.0
    density_pre: str = "identity"  # "identity" | "abs"
    density_post: str = "relu"  # "identity" | "relu" | "softplus"

    @property
    def dims(self) -> Tuple[int, int, int]:
        return tuple(self.features.shape[:3])

    @property
    def aabb(self) -> Tuple[Tuple[float, float], ...]:
        # voxels.py:187-212 -- python-float (double) arithmetic, centre +- dims*size/2
        out = []
        for n, s, c in zip(self.dims, self.voxel_size, self.location):
            half = (n * s) / 2
            out.append((c - half, c + half))
        return tuple(out)


_PRE = {"identity": lambda x: x, "abs": torch.abs}
_POST = {"identity": lambda x: x, "relu": F.relu, "softplus": F.softplus}


def cast_pinhole_rays(height: int, width: int, focal: float, rotation, translation):
    """Pixel-centre pinhole rays, flat row-major (y then x).  utils/misc.py:27-50."""
    rot = torch.as_tensor(np.asarray(rotation), dtype=torch.float32)
    trans = torch.as_tensor(np.asarray(translation), dtype=torch.float32).reshape(3)
    xs = torch.linspace(0.5, width - 0.5, width, dtype=torch.float32)
    ys = torch.linspace(0.5, height - 0.5, height, dtype=torch.float32)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")  # [H, W] each
    cam = torch.stack(
        [(xx - width * 0.5) / focal, -(yy - height * 0.5) / focal, -torch.ones_like(xx)], -1
    )
    dirs = (rot @ cam[..., None])[..., 0]
    origins = trans.expand_as(dirs)
    return origins.reshape(-1, 3).contiguous(), dirs.reshape(-1, 3).contiguous()


def spherical_pose(yaw_deg: float, pitch_deg: float, radius: float):
    """Rz(yaw) @ Rx(pitch) @ Tz(radius), fp32.  utils/imaging_utils.py:146-191."""
    yaw, pitch = yaw_deg / 180.0 * np.pi, pitch_deg / 180.0 * np.pi
    tz = torch.eye(4, dtype=torch.float32)
    tz[2, 3] = radius
    rx = torch.tensor(
        [
            [1.0, 0.0, 0.0, 0.0],
            [0.0, np.cos(pitch), -np.sin(pitch), 0.0],
            [0.0, np.sin(pitch), np.cos(pitch), 0.0],
            [0.0, 0.0, 0.0, 1.0],
        ],
        dtype=torch.float32,
    )
    rz = torch.tensor(
        [
            [np.cos(yaw), -np.sin(yaw), 0.0, 0.0],
            [np.sin(yaw), np.cos(yaw), 0.0, 0.0],
            [0.0, 0.0, 1.0, 0.0],
            [0.0, 0.0, 0.0, 1.0],
        ],
        dtype=torch.float32,
    )
    c2w = rz @ (rx @ tz)
    return c2w[:3, :3].contiguous(), c2w[:3, 3:].contiguous()


def aabb_ray_bounds(origins, directions, near: float, far: float, aabb) -> torch.Tensor:
    """Per-ray [near, far] from the slab test.  rendering/volumetric/sample.py:71-183.

    Quirks kept: ``d + 1e-10`` denominators (:102-103), the miss test of axis k uses the
    interval accumulated over the axes before it (:122-129, :153-160), misses fall back to the
    camera bounds (:175-177), result clipped at 0 (:180); hits are *not* intersected with the
    camera near/far.
    """
    n = origins.shape[0]
    cam = torch.tensor([near, far], dtype=origins.dtype, device=origins.device).reshape(1, 2).repeat(n, 1)
    hit = torch.ones(n, dtype=torch.bool, device=origins.device)
    lo = hi = None
    for axis, (a_lo, a_hi) in enumerate(aabb):
        denom = directions[:, axis] + ZERO_PLUS
        t0 = (a_lo - origins[:, axis]) / denom
        t1 = (a_hi - origins[:, axis]) / denom
        swap = t0 > t1
        ta, tb = torch.where(swap, t1, t0), torch.where(swap, t0, t1)
        if lo is None:
            lo, hi = ta, tb
            continue
        hit = hit & ~((lo > tb) | (ta > hi))
        lo = torch.where(ta > lo, ta, lo)
        hi = torch.where(tb < hi, tb, hi)
    bounds = torch.stack([lo, hi], -1)
    bounds = torch.where(hit[:, None], bounds, cam)
    return torch.clip(bounds, min=0.0)


def sample_depths(
    num_rays: int,
    num_samples: int,
    near,
    far,
    jitter: Optional[torch.Tensor],
    device=None,
) -> torch.Tensor:
    """Depths ``z[N, S]``.  rendering/volumetric/sample.py:38-64.

    ``near``/``far`` are python floats (CameraBounds) or ``[N, 1]`` tensors (per-ray bounds).
    ``jitter`` is the ``U[0,1)`` tensor the reference draws with ``torch.rand`` (:63); ``None``
    means ``perturb=False``.
    """
    if not torch.is_tensor(near):
        near = torch.tensor([near], dtype=torch.float32, device=device).repeat(num_rays, 1)
        far = torch.tensor([far], dtype=torch.float32, device=device).repeat(num_rays, 1)
    t = torch.linspace(0.0, 1.0, num_samples, dtype=torch.float32, device=near.device)[None, :]
    z = near * (1.0 - t) + far * t
    if jitter is not None:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mid, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mid], -1)
        z = lower + (upper - lower) * jitter
    return z


def normalise_points(points: torch.Tensor, aabb) -> torch.Tensor:
    """World -> [-1, 1] with fp32 scale/bias.  voxels.py:214-223, imaging_utils.py:58-63."""
    out = torch.empty_like(points)
    for axis, (lo, hi) in enumerate(aabb):
        scale = (np.float32(1.0) - np.float32(-1.0)) / (np.float32(hi) - np.float32(lo))
        bias = np.float32(-1.0) - np.float32(lo) * scale
        out[:, axis] = points[:, axis] * scale + bias
    return out


def grid_lookup(grid: OracleGrid, points: torch.Tensor) -> torch.Tensor:
    """``VoxelGrid.forward``: ``[P, 3] -> [P, F + 1]`` (features then density).  voxels.py:276-331."""
    npts = normalise_points(points, grid.aabb)[None, None, None, :, :]
    pre_d = _PRE[grid.density_pre](grid.densities * grid.density_scale)
    sigma = F.grid_sample(pre_d[None].permute(0, 4, 3, 2, 1), npts, align_corners=False)
    sigma = sigma.permute(0, 2, 3, 4, 1).reshape(-1, 1)
    sigma = _POST[grid.density_post](sigma)
    feats = F.grid_sample(grid.features[None].permute(0, 4, 3, 2, 1), npts, align_corners=False)
    feats = feats.permute(0, 2, 3, 4, 1).reshape(-1, grid.features.shape[-1])
    return torch.cat([feats, sigma], dim=-1)


def inside_mask(grid: OracleGrid, points: torch.Tensor) -> torch.Tensor:
    """Strict ``lo < p < hi`` on every axis.  voxels.py:252-274."""
    m = torch.ones(points.shape[0], dtype=torch.bool, device=points.device)
    for axis, (lo, hi) in enumerate(grid.aabb):
        m = m & (points[:, axis] > lo) & (points[:, axis] < hi)
    return m[:, None]


def sh_radiance(coeffs: torch.Tensor, viewdirs: torch.Tensor) -> torch.Tensor:
    """``[P, 3, K]`` coefficients x unit dirs ``[P, 3]`` -> ``[P, 3]``.  spherical_harmonics.py:64-116."""
    k = coeffs.shape[-1]
    deg = int(np.sqrt(k)) - 1
    assert (deg + 1) ** 2 == k and 0 <= deg < 4
    out = SH_C0 * coeffs[..., 0]
    if deg > 0:
        x, y, z = viewdirs[..., 0:1], viewdirs[..., 1:2], viewdirs[..., 2:3]
        out = out - SH_C1 * y * coeffs[..., 1] + SH_C1 * z * coeffs[..., 2] - SH_C1 * x * coeffs[..., 3]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        out = (
            out
            + SH_C2[0] * xy * coeffs[..., 4]
            + SH_C2[1] * yz * coeffs[..., 5]
            + SH_C2[2] * (2.0 * zz - xx - yy) * coeffs[..., 6]
            + SH_C2[3] * xz * coeffs[..., 7]
            + SH_C2[4] * (xx - yy) * coeffs[..., 8]
        )
    if deg > 2:
        out = (
            out
            + SH_C3[0] * y * (3 * xx - yy) * coeffs[..., 9]
            + SH_C3[1] * xy * z * coeffs[..., 10]
            + SH_C3[2] * y * (4 * zz - xx - yy) * coeffs[..., 11]
            + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * coeffs[..., 12]
            + SH_C3[4] * x * (4 * zz - xx - yy) * coeffs[..., 13]
            + SH_C3[5] * z * (xx - yy) * coeffs[..., 14]
            + SH_C3[6] * x * (xx - 3 * yy) * coeffs[..., 15]
        )
    return out


def composite(raw_radiance, sigma, depths, directions, white_bkgd: bool, noise=None):
    """Front-to-back alpha compositing.  rendering/volumetric/accumulate.py:43-113.

    ``raw_radiance [N, S, 3]``, ``sigma [N, S]``, ``depths [N, S]``.  ``noise`` is the optional
    ``randn * std`` tensor of :59-62 (``None`` = std 0).
    """
    deltas = depths[..., 1:] - depths[..., :-1]
    deltas = torch.cat([deltas, torch.full_like(deltas[..., :1], INFINITY)], dim=-1)
    deltas = deltas * directions[..., None, :].norm(dim=-1)
    if noise is not None:
        sigma = sigma + noise
    alpha = 1.0 - torch.exp(-(sigma * deltas))
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype, device=alpha.device)
    weights = alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha], -1), -1)[:, :-1]
    colour = torch.sum(torch.sigmoid(raw_radiance) * weights[..., None], dim=-2)
    acc = torch.sum(weights, dim=-1, keepdim=True)
    if white_bkgd:
        colour = colour + (1 - acc)
    depth = (depths * weights).sum(dim=-1, keepdim=True)
    disparity = 1.0 / torch.maximum(torch.full_like(acc, ZERO_PLUS), depth / acc)
    return colour, depth, acc, disparity


def render(
    grid: OracleGrid,
    origins: torch.Tensor,
    directions: torch.Tensor,
    *,
    num_samples: int,
    near: float,
    far: float,
    jitter: Optional[torch.Tensor] = None,
    white_bkgd: bool = False,
    diffuse: bool = False,
    optimized_sampling: bool = False,
    noise: Optional[torch.Tensor] = None,
) -> Dict[str, torch.Tensor]:
    """The whole ``render_sh_voxel_grid`` pipeline (thre3d_reprs/renderers.py:48-102)."""
    n = origins.shape[0]
    if optimized_sampling:
        b = aabb_ray_bounds(origins, directions, near, far, grid.aabb)
        z = sample_depths(n, num_samples, b[:, :1], b[:, 1:], jitter, device=origins.device)
    else:
        z = sample_depths(n, num_samples, near, far, jitter, device=origins.device)
    points = origins[:, None, :] + directions[:, None, :] * z[..., None]  # sample.py:67
    flat = points.reshape(-1, 3)

    looked_up = grid_lookup(grid, flat)  # process.py:37
    coeffs, sigma = looked_up[:, :-1], looked_up[:, -1:]
    viewdirs = directions / directions.norm(dim=-1, keepdim=True)  # process.py:53
    viewdirs = viewdirs[:, None, :].repeat(1, num_samples, 1).reshape(-1, 3)
    coeffs = coeffs.reshape(coeffs.shape[0], 3, -1)  # channel-major, process.py:61,66
    if diffuse:
        coeffs = coeffs[..., :1]  # process.py:59-63
    raw = sh_radiance(coeffs, viewdirs)

    inside = inside_mask(grid, flat)  # process.py:80-84
    raw = torch.where(inside, raw, torch.full_like(raw, -INFINITY))
    sigma = torch.where(inside, sigma, torch.zeros_like(sigma))

    colour, depth, acc, disparity = composite(
        raw.reshape(n, num_samples, 3),
        sigma.reshape(n, num_samples),
        z,
        directions,
        white_bkgd,
        noise,
    )
    return {"colour": colour, "depth": depth, "acc": acc, "disparity": disparity, "z": z}


def render_with_grads(
    grid: OracleGrid,
    origins,
    directions,
    grad_colour: torch.Tensor,
    grad_depth: Optional[torch.Tensor] = None,
    grad_acc: Optional[torch.Tensor] = None,
    ray_chunk: Optional[int] = None,
    **cfg,
):
    """Forward + autograd backward into the grid (what ``total_loss.backward()`` does, trainers.py:340).

    ``ray_chunk`` renders and back-propagates chunk by chunk, accumulating the grid gradient, the
    way the reference has to be driven at large shapes (it materialises ``[N*S, F+1]`` tensors).
    """
    dens = grid.densities.detach().clone().requires_grad_(True)
    feat = grid.features.detach().clone().requires_grad_(True)
    g = dataclasses.replace(grid, densities=dens, features=feat)
    n = origins.shape[0]
    step = n if ray_chunk is None else ray_chunk
    outs = []
    jitter = cfg.pop("jitter", None)
    for s in range(0, n, step):
        e = min(n, s + step)
        out = render(
            g,
            origins[s:e],
            directions[s:e],
            jitter=None if jitter is None else jitter[s:e],
            **cfg,
        )
        loss = (out["colour"] * grad_colour[s:e]).sum()
        if grad_depth is not None:
            loss = loss + (out["depth"] * grad_depth[s:e]).sum()
        if grad_acc is not None:
            loss = loss + (out["acc"] * grad_acc[s:e]).sum()
        loss.backward()
        outs.append({k: v.detach() for k, v in out.items()})
    merged = {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
    merged["grad_densities"] = dens.grad
    merged["grad_features"] = feat.grad
    return merged
