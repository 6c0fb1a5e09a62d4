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


def test_trainer_like_run_with_random_batches_and_jitter_matches_the_oracle_within_0p05_db(cuda_device):
    """The reference trainer's iteration at a less toy-like scale (modules/trainers.py:281-341): a 48^3 degree-2 ReLU field, 8 posed
    views, a fresh random ray batch and fresh stratified jitter every iteration (the in-kernel counter-based jitter, replayed
    through the oracle with tests/helpers.py::hash_jitter), specular + diffuse L1, Adam.  (a) CUDA path + fused Adam against
    (b) the oracle's ATen op sequence + torch.optim.Adam, both on the GPU.  Held-out view PSNR within 0.05 dB all along."""
    from cases import relu_field_density_scale, spherical_pose
    from helpers import hash_jitter
    from oracle import torch_port as tp
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.optim import FusedGridAdam
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_hints, render_sh_voxel_grid
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

    G, DEG, SIDE, S, VIEWS, ITERS, BATCH, LR = 48, 2, 40, 96, 8, 60, 4096, 0.03
    dev = cuda_device
    nf = 3 * (DEG + 1) ** 2
    scale = relu_field_density_scale((3, 3, 3))
    vs = (3 / G,) * 3
    gen = torch.Generator().manual_seed(7)

    # ground truth: two soft blobs with view-dependent colour; targets (8 training views + 1 held-out) rendered by the oracle
    ax = (torch.arange(G) + 0.5) / G * 3 - 1.5
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    blob = lambda cx, cy, cz, r: r - torch.sqrt((xx - cx) ** 2 + (yy - cy) ** 2 + (zz - cz) ** 2)  # noqa: E731
    gt_dens = torch.maximum(blob(0.3, 0.0, 0.1, 0.7), blob(-0.5, 0.3, -0.2, 0.5))[..., None].float() * 0.8
    gt_feat = torch.empty((G, G, G, nf)).uniform_(-1.5, 1.5, generator=gen)
    gt = tp.OracleGrid(gt_dens.to(dev), gt_feat.to(dev), vs, (0, 0, 0), scale, "identity", "relu")

    def view(yaw, pitch):
        rot, trans = spherical_pose(yaw, pitch, 4.0)
        o, d = tp.cast_pinhole_rays(SIDE, SIDE, 1111.11 * SIDE / 800, rot, trans)
        return o.to(dev), d.to(dev)

    train = [view(45.0 * k, 50.0 + 5.0 * (k % 3)) for k in range(VIEWS)]
    o_all, d_all = torch.cat([v[0] for v in train]), torch.cat([v[1] for v in train])
    o_test, d_test = view(22.5, 35.0)
    cfg = dict(num_samples=S, near=1.8, far=6.6, white_bkgd=True)
    with torch.no_grad():
        target_all = tp.render(gt, o_all, d_all, **cfg)["colour"]
        target_test = tp.render(gt, o_test, d_test, **cfg)["colour"]

    init_d = torch.empty((G, G, G, 1)).uniform_(-1, 1, generator=gen)
    init_f = torch.empty((G, G, G, nf)).uniform_(-1, 1, generator=gen)
    batches = [torch.randperm(o_all.shape[0], generator=gen)[:BATCH].to(dev) for _ in range(ITERS)]
    seed = lambda it, which: 1_000_003 * (2 * it + which) + 17  # noqa: E731  (spec / diffuse renders draw separate jitter, process-wide)

    # (b) oracle + torch Adam
    dens = init_d.clone().to(dev).requires_grad_(True)
    feat = init_f.clone().to(dev).requires_grad_(True)
    opt = torch.optim.Adam([dens, feat], lr=LR, betas=(0.9, 0.999))
    ref_curve = []
    for it in range(ITERS):
        idx = batches[it]
        o, d, tgt = o_all[idx], d_all[idx], target_all[idx]
        og = tp.OracleGrid(dens, feat, vs, (0, 0, 0), scale, "identity", "relu")
        j0 = torch.from_numpy(hash_jitter(seed(it, 0), BATCH, S)).to(dev)
        j1 = torch.from_numpy(hash_jitter(seed(it, 1), BATCH, S)).to(dev)
        spec = tp.render(og, o, d, jitter=j0, **cfg)["colour"]
        diff = tp.render(og, o, d, jitter=j1, diffuse=True, **cfg)["colour"]
        loss = torch.nn.functional.l1_loss(spec, tgt) + torch.nn.functional.l1_loss(diff, tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it % 10 == 9:
            with torch.no_grad():
                img = tp.render(og, o_test, d_test, **cfg)["colour"]
            ref_curve.append(_psnr(float(torch.nn.functional.mse_loss(img, target_test))))

    # (a) CUDA path + fused Adam
    grid = VoxelGrid(init_d.to(dev), init_f.to(dev), VoxelSize(*vs), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=scale, tunable=True)
    train_cfg = SHVoxGridRenderConfig(S, CameraBounds(1.8, 6.6), perturb_sampled_points=True, white_bkgd=True)
    test_cfg = SHVoxGridRenderConfig(S, CameraBounds(1.8, 6.6), perturb_sampled_points=False, white_bkgd=True)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, train_cfg, device=dev)
    fopt = FusedGridAdam(grid.parameters(), lr=LR, betas=(0.9, 0.999))
    curve = []
    for it in range(ITERS):
        idx = batches[it]
        rays, tgt = Rays(o_all[idx].contiguous(), d_all[idx].contiguous()), target_all[idx]
        with render_hints(rng_seed=seed(it, 0)):
            spec = vol_mod.render_rays(rays).colour
        with render_hints(rng_seed=seed(it, 1)):
            diff = vol_mod.render_rays(rays, render_diffuse=True).colour
        loss = torch.nn.functional.l1_loss(spec, tgt) + torch.nn.functional.l1_loss(diff, tgt)
        fopt.zero_grad()
        loss.backward()
        fopt.step()
        if it % 10 == 9:
            with torch.no_grad():
                img = render_sh_voxel_grid(grid, Rays(o_test, d_test), test_cfg).colour
            curve.append(_psnr(float(torch.nn.functional.mse_loss(img, target_test))))

    assert curve[-1] > curve[0] + 1.0, curve  # it learns
    gaps = np.abs(np.array(curve) - np.array(ref_curve))
    print(f"held-out PSNR every 10 iterations: cuda {np.round(curve, 4).tolist()} oracle {np.round(ref_curve, 4).tolist()} max gap {gaps.max():.5f} dB")
    assert gaps.max() < 0.05, f"held-out PSNR gap {gaps.max():.4f} dB: {curve} vs {ref_curve}"
