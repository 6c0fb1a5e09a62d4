"""Training-level parity: the same small ReLU-field fit run (a) through the CUDA path with the fused Adam and (b) through
the CPU oracle with torch.optim.Adam, from the same initial grid, on the same posed target images -- the structure of the
reference's training iteration (modules/trainers.py:306-341: specular L1 + diffuse L1, zero_grad / backward / Adam step).
BASELINE north star: rendered PSNR within 0.05 dB of the reference."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _psnr(mse: float) -> float:
    return -10.0 * math.log10(mse)


def test_short_training_run_matches_the_oracle_within_0p05_db(cuda_device):
    from cases import relu_field_density_scale, spherical_pose
    from oracle import torch_port as tp
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.optim import FusedGridAdam
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

    G, DEG, SIDE, S, VIEWS, ITERS, LR = 16, 1, 20, 48, 6, 40, 0.03
    nf = 3 * (DEG + 1) ** 2
    scale = relu_field_density_scale((3, 3, 3))
    vs = (3 / G,) * 3
    gen = torch.Generator().manual_seed(0)

    # ground-truth scene: a soft ball of density with view-dependent colour; targets rendered by the oracle
    ax = (torch.arange(G) + 0.5) / G * 3 - 1.5
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    gt_dens = (0.9 - torch.sqrt(xx**2 + yy**2 + zz**2))[..., None].float() * 0.8
    gt_feat = torch.empty((G, G, G, nf)).uniform_(-1.5, 1.5, generator=gen)
    gt = tp.OracleGrid(gt_dens, gt_feat, vs, (0, 0, 0), scale, "identity", "relu")
    rays_o, rays_d = [], []
    for k in range(VIEWS):
        rot, trans = spherical_pose(60.0 * k, 55.0, 4.0)
        o, d = tp.cast_pinhole_rays(SIDE, SIDE, 1111.11 * SIDE / 800, rot, trans)
        rays_o.append(o), rays_d.append(d)
    o, d = torch.cat(rays_o), torch.cat(rays_d)
    cfg = dict(num_samples=S, near=1.8, far=6.6, white_bkgd=True)
    with torch.no_grad():
        target = tp.render(gt, o, d, **cfg)["colour"]

    init_d = torch.empty((G, G, G, 1)).uniform_(-1, 1, generator=gen)
    init_f = torch.empty((G, G, G, nf)).uniform_(-1, 1, generator=gen)

    # (a) CPU oracle + torch Adam
    dens = init_d.clone().requires_grad_(True)
    feat = init_f.clone().requires_grad_(True)
    opt = torch.optim.Adam([dens, feat], lr=LR, betas=(0.9, 0.999))
    ref_curve = []
    for _ in range(ITERS):
        og = tp.OracleGrid(dens, feat, vs, (0, 0, 0), scale, "identity", "relu")
        spec = tp.render(og, o, d, **cfg)["colour"]
        diff = tp.render(og, o, d, diffuse=True, **cfg)["colour"]
        loss = torch.nn.functional.l1_loss(spec, target) + torch.nn.functional.l1_loss(diff, target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_curve.append(_psnr(float(torch.nn.functional.mse_loss(spec.detach(), target))))

    # (b) CUDA path + fused Adam, through VolumetricModel.render_rays like the trainer
    grid = VoxelGrid(init_d.to(cuda_device), init_f.to(cuda_device), VoxelSize(*vs), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=scale, tunable=True)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(S, CameraBounds(1.8, 6.6), perturb_sampled_points=False, white_bkgd=True),
                              device=cuda_device)
    rays = Rays(o.to(cuda_device), d.to(cuda_device))
    tgt = target.to(cuda_device)
    fopt = FusedGridAdam(grid.parameters(), lr=LR, betas=(0.9, 0.999))
    curve = []
    for _ in range(ITERS):
        spec = vol_mod.render_rays(rays).colour
        diff = vol_mod.render_rays(rays, render_diffuse=True).colour
        loss = torch.nn.functional.l1_loss(spec, tgt) + torch.nn.functional.l1_loss(diff, tgt)
        fopt.zero_grad()
        loss.backward()
        fopt.step()
        curve.append(_psnr(float(torch.nn.functional.mse_loss(spec.detach(), tgt))))

    assert curve[-1] > curve[0] + 3.0, (curve[0], curve[-1])  # it actually learns
    gaps = np.abs(np.array(curve) - np.array(ref_curve))
    assert gaps.max() < 0.05, f"PSNR gap {gaps.max():.4f} dB (final {curve[-1]:.3f} vs {ref_curve[-1]:.3f})"
    # and the fitted parameters themselves stay together
    rel = float((grid.features.detach().cpu() - feat.detach()).norm() / feat.detach().norm())
    assert rel < 2e-2, rel
