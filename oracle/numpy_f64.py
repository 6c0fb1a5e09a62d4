"""fp64 NumPy restatement of the render path with a hand-written trilinear gather and the
*analytic* backward pass (no autograd).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Purpose: an implementation that shares no code with ATen's ``grid_sample`` nor with autograd, in
double precision, so that when the fp32 CUDA kernels and the fp32 torch port disagree in the last
digits (different summation order, fused multiply-adds, atomics) there is an arbiter.  It follows
the mathematical contract in SURVEY.md appendix A; the reference lines each block restates are
cited inline (paths relative to ``/root/reference/thre3d_atom/``).

Sample *positions* are taken in fp32 exactly as the reference computes them (they decide the
discontinuous parts: which cell a sample is in and whether it is inside the AABB); everything
downstream of the positions is fp64.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

ZERO_PLUS = 1e-10  # utils/constants.py:7
INFINITY = 1e10  # utils/constants.py:8

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (
    -0.5900435899266435,
    2.890611442640554,
    -0.4570457994644658,
    0.3731763325901154,
    -0.4570457994644658,
    1.445305721320277,
    -0.5900435899266435,
)

f32 = np.float32


def sh_basis(deg: int, v: np.ndarray) -> np.ndarray:
    """Signed real-SH basis ``Y[N, K]`` such that ``raw_ch = sum_k Y_k * coeff[ch][k]``.

    rendering/volumetric/utils/spherical_harmonics.py:86-116 (constants :33-50).
    """
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    cols = [np.full_like(x, C0)]
    if deg > 0:
        cols += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        cols += [
            C3[0] * y * (3 * xx - yy),
            C3[1] * xy * z,
            C3[2] * y * (4 * zz - xx - yy),
            C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
            C3[4] * x * (4 * zz - xx - yy),
            C3[5] * z * (xx - yy),
            C3[6] * x * (xx - 3 * yy),
        ]
    return np.stack(cols, -1)


def linspace01_f32(steps: int) -> np.ndarray:
    """``torch.linspace(0, 1, steps, dtype=float32)`` as ATen's *per-element* kernels compute it
    (RangeFactories: ``start + step*i`` for ``i < steps//2``, else ``end - step*(steps-1-i)``), with
    the multiply-add fused as on the CUDA device -- bit-exact versus ``torch.linspace(device="cuda")``
    (checked in the GPU tests).  ATen's vectorised CPU kernel builds the second half as
    ``arange(base, step)`` per SIMD vector instead and can differ from this by one ulp, i.e. the
    reference itself is not bit-stable across devices here."""
    if steps == 1:
        return np.zeros(1, f32)
    step = np.float64(f32(1.0) / f32(steps - 1))
    i = np.arange(steps)
    lo = (step * i).astype(f32)  # fma(step, i, 0) == round(step * i)
    hi = (1.0 - step * (steps - 1 - i)).astype(f32)  # fma(-step, steps-1-i, 1)
    return np.where(i < steps // 2, lo, hi).astype(f32)


def depths_f32(near: np.ndarray, far: np.ndarray, steps: int, jitter: Optional[np.ndarray]) -> np.ndarray:
    """fp32 depths, op order of rendering/volumetric/sample.py:46-64. near/far: ``[N, 1]`` fp32."""
    t = linspace01_f32(steps)[None, :]
    z = (near * (f32(1.0) - t)).astype(f32) + (far * t).astype(f32)
    z = z.astype(f32)
    if jitter is not None:
        mid = (f32(0.5) * (z[:, 1:] + z[:, :-1]).astype(f32)).astype(f32)
        upper = np.concatenate([mid, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mid], -1)
        z = (lower + ((upper - lower).astype(f32) * jitter.astype(f32)).astype(f32)).astype(f32)
    return z


def aabb_ray_bounds_f32(o, d, near, far, aabb) -> np.ndarray:
    """rendering/volumetric/sample.py:71-183 in fp32 (see torch_port.aabb_ray_bounds)."""
    n = o.shape[0]
    hit = np.ones(n, bool)
    lo = hi = None
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax, (a_lo, a_hi) in enumerate(aabb):
            den = (d[:, ax] + f32(ZERO_PLUS)).astype(f32)
            t0 = ((f32(a_lo) - o[:, ax]).astype(f32) / den).astype(f32)
            t1 = ((f32(a_hi) - o[:, ax]).astype(f32) / den).astype(f32)
            ta, tb = np.where(t0 > t1, t1, t0), np.where(t0 > t1, t0, t1)
            if lo is None:
                lo, hi = ta, tb
                continue
            hit &= ~((lo > tb) | (ta > hi))
            lo = np.where(ta > lo, ta, lo)
            hi = np.where(tb < hi, tb, hi)
    b = np.stack([lo, hi], -1)
    b = np.where(hit[:, None], b, np.array([[near, far]], f32))
    return np.maximum(b, f32(0.0)).astype(f32)


def _post(kind: str, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """value and derivative of the density post-activation (voxels.py:309)."""
    if kind == "relu":
        return np.maximum(x, 0.0), (x > 0).astype(x.dtype)
    if kind == "softplus":  # torch.nn.Softplus(beta=1, threshold=20)
        big = x > 20.0
        sp = np.where(big, x, np.log1p(np.exp(np.minimum(x, 20.0))))
        return sp, np.where(big, 1.0, 1.0 / (1.0 + np.exp(-x)))
    return x, np.ones_like(x)


def render(
    densities: np.ndarray,  # [W, D, H, 1]
    features: np.ndarray,  # [W, D, H, F]
    aabb,  # ((x0,x1),(y0,y1),(z0,z1)) python floats
    origins: np.ndarray,  # [N, 3] fp32
    directions: np.ndarray,  # [N, 3] fp32
    *,
    num_samples: int,
    near: float,
    far: float,
    density_scale: float = 1.0,
    density_pre: str = "identity",
    density_post: str = "relu",
    jitter: Optional[np.ndarray] = None,
    white_bkgd: bool = False,
    diffuse: bool = False,
    optimized_sampling: bool = False,
    grad_colour: Optional[np.ndarray] = None,  # [N, 3]
    grad_depth: Optional[np.ndarray] = None,  # [N, 1]
    grad_acc: Optional[np.ndarray] = None,  # [N, 1]
) -> Dict[str, np.ndarray]:
    o32, d32 = origins.astype(f32), directions.astype(f32)
    n, s = o32.shape[0], num_samples
    w_, d_, h_, nf = features.shape
    k_all = nf // 3
    deg = int(round(np.sqrt(k_all))) - 1
    dims = np.array([w_, d_, h_])

    # ---- fp32 sample positions (sample.py:38-67) ----
    if optimized_sampling:
        b = aabb_ray_bounds_f32(o32, d32, f32(near), f32(far), aabb)
        z32 = depths_f32(b[:, :1], b[:, 1:], s, jitter)
    else:
        z32 = depths_f32(np.full((n, 1), near, f32), np.full((n, 1), far, f32), s, jitter)
    p32 = (o32[:, None, :] + (d32[:, None, :] * z32[:, :, None]).astype(f32)).astype(f32)  # [N,S,3]

    # strict inside test on the fp32 points (voxels.py:252-274)
    inside = np.ones((n, s), bool)
    for ax, (lo, hi) in enumerate(aabb):
        inside &= (p32[..., ax] > f32(lo)) & (p32[..., ax] < f32(hi))

    # ---- continuous voxel index (voxels.py:214-223 + grid_sample align_corners=False) ----
    # n = p*scale + bias in fp32 (imaging_utils.py:58-63), then i = ((n + 1) * dim - 1) / 2
    idx = np.empty((n, s, 3), np.float64)
    for ax, (lo, hi) in enumerate(aabb):
        scale = (f32(1.0) - f32(-1.0)) / (f32(hi) - f32(lo))
        bias = f32(-1.0) - f32(lo) * scale
        nrm = ((p32[..., ax] * scale).astype(f32) + bias).astype(f32)
        idx[..., ax] = ((nrm.astype(np.float64) + 1.0) * dims[ax] - 1.0) / 2.0
    i0 = np.floor(idx).astype(np.int64)
    fr = idx - i0

    dens = densities.astype(np.float64)[..., 0]
    feat = features.astype(np.float64)
    pre_d = np.abs(dens * density_scale) if density_pre == "abs" else dens * density_scale

    sig_pre = np.zeros((n, s))
    coef = np.zeros((n, s, nf))
    corner_cache = []
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ii = i0 + np.array([cx, cy, cz])
                ok = np.all((ii >= 0) & (ii < dims), -1)  # zero padding
                wgt = (
                    (fr[..., 0] if cx else 1 - fr[..., 0])
                    * (fr[..., 1] if cy else 1 - fr[..., 1])
                    * (fr[..., 2] if cz else 1 - fr[..., 2])
                )
                wgt = np.where(ok, wgt, 0.0)
                ic = np.clip(ii, 0, dims - 1)
                sig_pre += wgt * pre_d[ic[..., 0], ic[..., 1], ic[..., 2]]
                coef += wgt[..., None] * feat[ic[..., 0], ic[..., 1], ic[..., 2]]
                corner_cache.append((ic, wgt))
    sigma, dpost = _post(density_post, sig_pre)

    # ---- SH radiance (process.py:53-72) ----
    dn = np.linalg.norm(d32.astype(np.float64), axis=-1, keepdims=True)
    v = (d32 / np.linalg.norm(d32, axis=-1, keepdims=True).astype(f32)).astype(np.float64)
    ybasis = sh_basis(deg, v)  # [N, K]
    if diffuse:
        ybasis = ybasis.copy()
        ybasis[:, 1:] = 0.0  # only k = 0 survives (process.py:59-63)
    coef3 = coef.reshape(n, s, 3, k_all)
    raw = np.einsum("nsck,nk->nsc", coef3, ybasis)

    # mask (process.py:80-84)
    raw = np.where(inside[..., None], raw, -INFINITY)
    sigma = np.where(inside, sigma, 0.0)

    # ---- compositing (accumulate.py:43-88) ----
    z = z32.astype(np.float64)
    delta = np.concatenate([z[:, 1:] - z[:, :-1], np.full((n, 1), INFINITY)], -1) * dn
    alpha = 1.0 - np.exp(-sigma * delta)
    trans = np.cumprod(np.concatenate([np.ones((n, 1)), 1.0 - alpha], -1), -1)  # T_0..T_S
    wts = alpha * trans[:, :-1]
    with np.errstate(over="ignore"):
        sg = 1.0 / (1.0 + np.exp(-raw))
    colour_fg = np.sum(wts[..., None] * sg, 1)
    acc = wts.sum(-1, keepdims=True)
    colour = colour_fg + (1.0 - acc) if white_bkgd else colour_fg
    depth = (wts * z).sum(-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = depth / acc
        disparity = 1.0 / np.where(np.isnan(ratio), ratio, np.maximum(ZERO_PLUS, ratio))
    out = {
        "colour": colour,
        "depth": depth,
        "acc": acc,
        "disparity": disparity,
        "z": z32,
        "inside": inside,
        "sigma": sigma,
        "weights": wts,
    }
    if grad_colour is None:
        return out

    # ---- analytic backward ----
    gc = grad_colour.astype(np.float64)
    gd = np.zeros((n, 1)) if grad_depth is None else grad_depth.astype(np.float64)
    ga = np.zeros((n, 1)) if grad_acc is None else grad_acc.astype(np.float64)
    if white_bkgd:
        ga = ga - gc.sum(-1, keepdims=True)
    q = np.einsum("nsc,nc->ns", sg, gc) + gd * z + ga  # per-sample scalar
    wq = wts * q
    suffix = np.cumsum(wq[:, ::-1], -1)[:, ::-1] - wq  # sum_{j>i} w_j q_j
    d_sigma = delta * (trans[:, 1:] * q - suffix)
    d_sigma = np.where(inside, d_sigma, 0.0)
    d_sig_pre = d_sigma * dpost  # through the post-activation
    d_raw = wts[..., None] * gc[:, None, :] * sg * (1.0 - sg)
    d_raw = np.where(inside[..., None], d_raw, 0.0)
    d_coef = (d_raw[..., :, None] * ybasis[:, None, None, :]).reshape(n, s, nf)

    g_pre = np.zeros_like(dens)
    g_feat = np.zeros_like(feat)
    for ic, wgt in corner_cache:
        np.add.at(g_pre, (ic[..., 0], ic[..., 1], ic[..., 2]), wgt * d_sig_pre)
        np.add.at(g_feat, (ic[..., 0], ic[..., 1], ic[..., 2]), wgt[..., None] * d_coef)
    if density_pre == "abs":
        g_dens = g_pre * np.sign(dens * density_scale) * density_scale
    else:
        g_dens = g_pre * density_scale
    out["grad_densities"] = g_dens[..., None]
    out["grad_features"] = g_feat
    return out
