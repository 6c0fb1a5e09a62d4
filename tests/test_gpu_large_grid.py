"""Parity at a grid whose FEATURE storage exceeds 2^31 floats (BASELINE.json configs[4] is 512^3 degree 3 = 6.4 G floats).

360^3 voxels x 48 features (degree 3) = 2.24 G floats = 9 GB: every float offset of the upper ~4 % of the grid is above
2^31, so the 32-bit float4 record indices of the forward kernels and the 32x32 -> 64-bit offset products of the backward
scatter are exercised, against the CPU oracle (oracle/torch_port.py, the reference algorithm on ATen CPU kernels), with
the in-kernel stratified jitter replayed through the oracle (tests/helpers.py::hash_jitter).

Replaces reference thre3d_reprs/voxels.py:296-318 (grid_sample over the [W, D, H, F] parameter) and its autograd backward.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from helpers import hash_jitter

pytestmark = pytest.mark.gpu

G, DEG, S = 360, 3, 96
NF = 3 * (DEG + 1) ** 2
WORLD = 3.0


def _grid_values():
    gen = torch.Generator().manual_seed(360)
    dens = torch.empty((G, G, G, 1), dtype=torch.float32).uniform_(-1.0, 1.0, generator=gen)
    # 9 GB of features: a random 90^3 block tiled 4x per axis (drawing 2.2 G randoms on one CPU thread takes a minute);
    # a 2^32-float wrap-around would move a record by 89 478 485.33 voxels, which no tiling period hides
    block = torch.empty((G // 4, G // 4, G // 4, NF), dtype=torch.float32).uniform_(-1.0, 1.0, generator=gen)
    feat = block.repeat(4, 4, 4, 1)
    return dens, feat


def _rays():
    from cases import HOTDOG_RADIUS, spherical_pose
    from oracle import torch_port as tp

    rot, trans = spherical_pose(30.0, 60.0, HOTDOG_RADIUS)
    o, d = tp.cast_pinhole_rays(32, 32, 1111.11 * 32 / 800.0, torch.from_numpy(rot), torch.from_numpy(trans))
    # 64 extra rays through the high-address corner (x, y, z all near +1.5), from outside the grid
    rng = np.random.RandomState(5)
    target = np.float32(WORLD / 2) - rng.uniform(0.0, 0.12, size=(64, 3)).astype(np.float32)
    origin = np.array([2.6, 3.1, 2.9], dtype=np.float32)[None] + rng.uniform(-0.2, 0.2, size=(64, 3)).astype(np.float32)
    dirs = target - origin
    dirs = dirs / np.linalg.norm(dirs, axis=1, keepdims=True) * rng.uniform(0.9, 1.2, size=(64, 1)).astype(np.float32)
    o = torch.cat([o.reshape(-1, 3), torch.from_numpy(origin.astype(np.float32))])
    d = torch.cat([d.reshape(-1, 3), torch.from_numpy(dirs.astype(np.float32))])
    return o.contiguous(), d.contiguous()


def test_grid_with_more_than_2_31_feature_floats(cuda_device):
    from cases import relu_field_density_scale
    from oracle import torch_port as tp
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_hints, render_sh_voxel_grid
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 40 * 2**30:
        pytest.skip("needs ~30 GB of device memory")
    dens, feat = _grid_values()
    assert feat.numel() > 2**31
    scale = relu_field_density_scale((WORLD,) * 3) * 0.05  # thin medium: rays reach the far (high-address) side of the grid
    near, far = 0.5, 7.0
    o, d = _rays()
    n = o.shape[0]
    assert n >= 1024
    seed = 0x0BADC0FFEE % (2**62)
    u = torch.from_numpy(hash_jitter(seed, n, S))
    gen = torch.Generator().manual_seed(3)
    gc = torch.rand((n, 3), generator=gen) - 0.5

    grid = VoxelGrid(densities=dens.to(cuda_device), features=feat.to(cuda_device), voxel_size=VoxelSize(*(WORLD / G,) * 3),
                     density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                     expected_density_scale=scale, tunable=True)
    assert grid.feature_storage.numel() > 2**31
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=CameraBounds(near, far), perturb_sampled_points=True, white_bkgd=True)
    with render_hints(rng_seed=seed):
        out = render_sh_voxel_grid(grid, Rays(o.to(cuda_device), d.to(cuda_device)), cfg)
        (out.colour * gc.to(cuda_device)).sum().backward()
    torch.cuda.synchronize()

    want = tp.render_with_grads(tp.OracleGrid(dens, feat, (WORLD / G,) * 3, (0.0, 0.0, 0.0), scale, "identity", "relu"), o, d, gc,
                                num_samples=S, near=near, far=far, jitter=u, white_bkgd=True)
    del feat
    np.testing.assert_allclose(out.colour.detach().cpu().numpy(), want["colour"].numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(out.extra["accumulated_weight"].detach().cpu().numpy(), want["acc"].numpy(), atol=2e-5, rtol=0)

    def rel_l2_cuda(got: torch.Tensor, ref_cpu: torch.Tensor) -> float:
        num = den = 0.0
        flat_g, flat_r = got.reshape(-1), ref_cpu.reshape(-1)
        for s0 in range(0, flat_g.numel(), 2**28):  # chunked: the tensors are 9 GB each
            r = flat_r[s0:s0 + 2**28].to(cuda_device).double()
            num += float(((flat_g[s0:s0 + 2**28].double() - r) ** 2).sum())
            den += float((r**2).sum())
        return (num / max(den, 1e-300)) ** 0.5

    gf, gd = grid.feature_storage.grad, grid.densities.grad
    assert gf.shape[-1] == NF  # degree 3: 48 floats, no padding lane
    assert rel_l2_cuda(gf, want["grad_features"]) < 1e-4
    assert rel_l2_cuda(gd, want["grad_densities"]) < 1e-4
    # the part of the gradient that lives above float offset 2^31 is populated (the extra rays cross it) and agrees by itself
    hi_got, hi_ref = gf.reshape(-1)[2**31:], want["grad_features"].reshape(-1)[2**31:]
    assert int((hi_ref != 0).sum()) > 10000
    assert rel_l2_cuda(hi_got, hi_ref) < 1e-4
