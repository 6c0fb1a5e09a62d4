"""GPU tests of the pieces around the fused renderer and of its edge cases: in-kernel ray generation,
the counter-based jitter, the point-lookup kernels, the unpadded (reference) feature layout, Adam,
opaque scenes (exact early termination), degenerate shapes, error behaviour -- and size-independent
properties at the BASELINE.json shapes (128^3 / 256^3 grids)."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import (
    CASES,
    aabb_of,
    assert_outputs_close,
    build_inputs,
    hash_jitter,
    load_golden,
    make_cuda_config,
    make_cuda_grid,
    rel_l2,
    run_cuda_case,
    run_numpy_f64,
)

pytestmark = pytest.mark.gpu


def _rays(inp, device):
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays

    return Rays(torch.from_numpy(inp["origins"]).to(device), torch.from_numpy(inp["directions"]).to(device))


# ---------------------------------------------------------------------------------------------
# ray generation (reference rendering/volumetric/utils/misc.py:12-50)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_32cube_deg0", "deg1_aniso_softplus", "cube2"])
def test_cast_rays_kernel_matches_reference(name, cuda_device):
    from cases import spherical_pose
    from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thr3ed_atom_b200.utils.imaging_utils import CameraIntrinsics, CameraPose

    case, gold = CASES[name], load_golden(name)
    rot, trans = spherical_pose(*case.pose)
    rays = cast_rays(CameraIntrinsics(case.image_hw[0], case.image_hw[1], case.focal), CameraPose(rot, trans), device=cuda_device)
    assert tuple(rays.origins.shape) == (case.image_hw[0], case.image_hw[1], 3)
    flat = flatten_rays(rays)
    np.testing.assert_array_equal(flat.origins.cpu().numpy(), gold["cast_origins"])
    np.testing.assert_allclose(flat.directions.cpu().numpy(), gold["cast_directions"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["c1_32cube_deg0", "deg2_16cube"])
def test_camera_render_equals_render_of_cast_rays(name, cuda_device):
    """VolumetricModel.render (in-kernel ray generation, one launch) == render_rays(cast_rays(...)) bit for bit."""
    from cases import spherical_pose
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid
    from thr3ed_atom_b200.utils.imaging_utils import CameraIntrinsics, CameraPose

    case = CASES[name]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, make_cuda_config(case), device=cuda_device)
    rot, trans = spherical_pose(*case.pose)
    intr, pose = CameraIntrinsics(case.image_hw[0], case.image_hw[1], case.focal), CameraPose(rot, trans)
    image = vol_mod.render(pose, intr)
    assert tuple(image.colour.shape) == (intr.height, intr.width, 3) and tuple(image.depth.shape) == (intr.height, intr.width, 1)
    with torch.no_grad():
        flat = vol_mod.render_rays(flatten_rays(cast_rays(intr, pose, device=cuda_device)))
    assert torch.equal(image.colour.reshape(-1, 3), flat.colour)
    assert torch.equal(image.depth.reshape(-1, 1), flat.depth)
    assert torch.equal(image.extra["accumulated_weight"].reshape(-1, 1), flat.extra["accumulated_weight"])
    # and the CPU copy path of the reference signature
    cpu_image = vol_mod.render(pose, intr, gpu_render=False, parallel_rays_chunk_size=1000)
    assert cpu_image.colour.device.type == "cpu" and torch.equal(cpu_image.colour, image.colour.cpu())
    # overriding an unknown config field is an error, as in the reference (volumetric_model.py:74-79)
    with pytest.raises(ValueError):
        vol_mod.render_rays(flatten_rays(cast_rays(intr, pose, device=cuda_device)), not_a_field=1)


# ---------------------------------------------------------------------------------------------
# stratified jitter from the in-kernel counter-based RNG
# ---------------------------------------------------------------------------------------------
def test_in_kernel_jitter_replays_through_the_oracle(cuda_device):
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    case = dataclasses.replace(CASES["deg2_jitter"], jitter=True)
    inp = build_inputs(case)
    seed = 0x1234_5678_9ABC_DEF0 % (2**62)
    n, s = inp["origins"].shape[0], case.num_samples
    u = hash_jitter(seed, n, s)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(float(u.mean()) - 0.5) < 0.01
    grid = make_cuda_grid(case, inp, cuda_device)
    cfg = make_cuda_config(case)
    with render_hints(rng_seed=seed):
        a = render_sh_voxel_grid(grid, _rays(inp, cuda_device), cfg)
        loss = (a.colour * torch.from_numpy(inp["grad_colour"]).to(cuda_device)).sum()
        loss.backward()
    ga = grid.feature_storage.grad.clone()
    grid.zero_grad()
    with render_hints(jitter=torch.from_numpy(u).to(cuda_device)):
        b = render_sh_voxel_grid(grid, _rays(inp, cuda_device), cfg)
        (b.colour * torch.from_numpy(inp["grad_colour"]).to(cuda_device)).sum().backward()
    assert torch.equal(a.colour, b.colour) and torch.equal(a.depth, b.depth)
    assert rel_l2(ga.cpu().numpy(), grid.feature_storage.grad.cpu().numpy()) < 1e-5  # backward re-derived the same offsets
    want = run_numpy_f64(case, {**inp, "jitter": u})
    got = {"colour": a.colour.detach().cpu().numpy(), "depth": a.depth.detach().cpu().numpy(),
           "acc": a.extra["accumulated_weight"].detach().cpu().numpy(), "disparity": a.extra["disparity"].detach().cpu().numpy()}
    assert_outputs_close(got, want, atol=1e-5, rtol_depth=1e-5, what="in-kernel jitter")
    # a different seed gives a different (but statistically equivalent) image; torch.manual_seed pins the default seed
    with render_hints(rng_seed=seed + 1), torch.no_grad():
        c = render_sh_voxel_grid(grid, _rays(inp, cuda_device), cfg)
    assert not torch.equal(a.colour, c.colour)
    assert abs(float(a.colour.detach().mean() - c.colour.mean())) < 5e-3
    with torch.no_grad():
        torch.manual_seed(11)
        d1 = render_sh_voxel_grid(grid, _rays(inp, cuda_device), cfg)
        torch.manual_seed(11)
        d2 = render_sh_voxel_grid(grid, _rays(inp, cuda_device), cfg)
    assert torch.equal(d1.colour, d2.colour)


# ---------------------------------------------------------------------------------------------
# VoxelGrid.forward / test_inside_volume (reference voxels.py:252-331)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_32cube_deg0", "deg1_aniso_softplus", "deg3_abs", "deg2_16cube"])
def test_point_lookup_matches_reference(name, cuda_device):
    from oracle import torch_port as tp

    case, gold = CASES[name], load_golden(name)
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    pts = torch.from_numpy(gold["lookup_points"]).to(cuda_device)
    values = grid(pts)
    np.testing.assert_allclose(values.detach().cpu().numpy(), gold["lookup_values"], atol=2e-5, rtol=1e-5)
    assert np.array_equal(grid.test_inside_volume(pts).cpu().numpy(), gold["lookup_inside"])
    # backward of the lookup against autograd of the fp32 torch port
    g_out = torch.from_numpy(np.random.RandomState(3).normal(size=gold["lookup_values"].shape).astype(np.float32))
    (values * g_out.to(cuda_device)).sum().backward()
    dens = torch.from_numpy(inp["densities"]).requires_grad_(True)
    feat = torch.from_numpy(inp["features"]).requires_grad_(True)
    og = tp.OracleGrid(dens, feat, case.voxel_size, case.location, case.density_scale, case.density_pre, case.density_post)
    (tp.grid_lookup(og, torch.from_numpy(gold["lookup_points"])) * g_out).sum().backward()
    nf = inp["features"].shape[-1]
    assert rel_l2(grid.densities.grad.cpu().numpy(), dens.grad.numpy()) < 1e-5
    assert rel_l2(grid.feature_storage.grad[..., :nf].cpu().numpy(), feat.grad.numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------
# reference (unpadded) feature layout through the raw C ABI wrappers
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["deg2_16cube", "c1_32cube_deg0"])
def test_unpadded_reference_layout_gives_the_same_result(name, cuda_device):
    """The C ABI accepts any record stride: whole 16-byte vectors (what VoxelGrid stores), whole 32-byte sectors
    (256-bit loads in the per-ray kernels) and the reference's unpadded layout (scalar loads).  All give the same result
    (the padded layouts take the lane-group forward, which sums in another order than the per-ray kernel the unpadded
    layout takes: equal to fp32 rounding, not bit for bit)."""
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.thre3d_reprs.renderers import make_render_args

    case = CASES[name]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    nf = inp["features"].shape[-1]
    padded = grid.kernel_desc()
    unpadded = dataclasses.replace(padded, features=torch.from_numpy(inp["features"]).to(cuda_device).contiguous())
    stride8 = (nf + 7) // 8 * 8
    vec4 = torch.zeros((*inp["features"].shape[:3], stride8), device=cuda_device)
    vec4[..., :nf] = unpadded.features
    vec4 = dataclasses.replace(padded, features=vec4)
    assert padded.features.shape[-1] % 4 == 0 and unpadded.features.shape[-1] % 4 != 0
    args = make_render_args(make_cuda_config(case))
    o, d = torch.from_numpy(inp["origins"]).to(cuda_device), torch.from_numpy(inp["directions"]).to(cuda_device)
    gc = torch.from_numpy(inp["grad_colour"]).to(cuda_device)
    results = []
    for desc in (padded, vec4, unpadded):
        out = _kernels.render_forward(desc, o, d, args)
        g_d, g_f = torch.zeros_like(desc.densities), torch.zeros_like(desc.features)
        _kernels.render_backward(desc, o, d, args, out[:3], (gc, None, None, None), g_d, g_f)
        results.append((out, g_d, g_f))
    (out_p, gp_d, gp_f) = results[0]
    for out_o, go_d, go_f in results[1:]:
        for a, b in zip(out_p[:3], out_o[:3]):
            assert ((a - b).abs() <= 2e-6 * b.abs().clamp(min=1.0)).all()
        assert rel_l2(gp_f[..., :nf].cpu().numpy(), go_f[..., :nf].cpu().numpy()) < 1e-5
        assert rel_l2(gp_d.cpu().numpy(), go_d.cpu().numpy()) < 1e-5
        assert not bool(go_f[..., nf:].any())


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def test_opaque_scene_exact_early_termination(cuda_device):
    """Large densities drive alpha to exactly 1 and T to exactly 0: the kernels stop the ray there, which must not
    change anything (all later weights are exactly 0).

    Density gradients in such a scene are vanishing (|g| ~ 5e-10 here, against ~1e-1 for a semi-transparent scene):
    they are carried by samples whose remaining transmittance is below fp32 resolution of the ray's total.  The
    single-pass backward forms the suffix sum as (total - prefix), clamped to its analytic bound T_{i+1} * max|q|, so its
    absolute error there is bounded by delta * scale * min(6e-8 |total|, T_{i+1} max|q|) -- checked on that absolute
    scale; feature gradients and everything in the forward pass are checked at the usual tolerances."""
    case = dataclasses.replace(CASES["deg2_16cube"], density_range=(20.0, 60.0), name="opaque")
    inp = build_inputs(case)
    want = run_numpy_f64(case, inp)
    got = run_cuda_case(case, inp, cuda_device)
    assert float(want["acc"].max()) > 0.999999
    assert_outputs_close(got, want, atol=1e-5, rtol_depth=1e-5, what="opaque")
    assert rel_l2(got["grad_features"], want["grad_features"]) < 5e-5
    delta = (case.far - case.near) / (case.num_samples - 1) * float(np.linalg.norm(inp["directions"], axis=-1).max())
    noise_floor = 6e-8 * delta * case.density_scale * float(np.abs(inp["grad_colour"]).sum(-1).max())
    assert float(np.abs(got["grad_densities"] - want["grad_densities"]).max()) < noise_floor


def test_far_plane_inside_the_grid_last_delta_is_infinite(cuda_device):
    """When the far bound ends inside the volume the last sample gets delta = 1e10*|d| (accumulate.py:49-53)."""
    case = dataclasses.replace(CASES["deg2_16cube"], far=4.2, num_samples=33, name="far_inside")
    inp = build_inputs(case)
    want = run_numpy_f64(case, inp)
    got = run_cuda_case(case, inp, cuda_device)
    assert_outputs_close(got, want, atol=1e-5, rtol_depth=1e-5, what="far_inside")
    assert rel_l2(got["grad_features"], want["grad_features"]) < 5e-5
    assert rel_l2(got["grad_densities"], want["grad_densities"]) < 5e-5


@pytest.mark.parametrize("num_samples", [1, 2, 3])
def test_tiny_sample_counts(num_samples, cuda_device):
    case = dataclasses.replace(CASES["deg2_16cube"], num_samples=num_samples, near=3.0, far=4.5, name=f"s{num_samples}")
    for jitter in (False, True):
        c = dataclasses.replace(case, jitter=jitter)
        inp = build_inputs(c)
        want = run_numpy_f64(c, inp)
        got = run_cuda_case(c, inp, cuda_device)
        # with 1-3 samples the intervals are huge (delta ~ 1.5 |d|), which amplifies the fp32 rounding of
        # near-zero interpolated densities (33.3 * a cancelling 8-term sum) into alpha: 5e-5 instead of 1e-5
        assert_outputs_close(got, want, atol=5e-5, rtol_depth=5e-5, what=c.name)
        assert rel_l2(got["grad_features"], want["grad_features"]) < 5e-5


def test_zero_rays_and_per_ray_bounds(cuda_device):
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import make_render_args, render_sh_voxel_grid

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    empty = Rays(torch.empty((0, 3), device=cuda_device), torch.empty((0, 3), device=cuda_device))
    out = render_sh_voxel_grid(grid, empty, make_cuda_config(case))
    assert tuple(out.colour.shape) == (0, 3) and tuple(out.depth.shape) == (0, 1)
    # explicit per-ray bounds tensor (reference sample.py:42-43) equal to the camera bounds => same result
    args = make_render_args(make_cuda_config(case))
    o, d = torch.from_numpy(inp["origins"]).to(cuda_device), torch.from_numpy(inp["directions"]).to(cuda_device)
    a = _kernels.render_forward(grid.kernel_desc(), o, d, args)
    bounds = torch.tensor([[case.near, case.far]], device=cuda_device).repeat(o.shape[0], 1).contiguous()
    b = _kernels.render_forward(grid.kernel_desc(), o, d, dataclasses.replace(args, ray_bounds=bounds))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_gradient_accumulates_across_renders_like_autograd(cuda_device):
    """Two renders -> two backward passes accumulate into .grad (the trainer's specular + diffuse losses)."""
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    rays = _rays(inp, cuda_device)
    gc = torch.from_numpy(inp["grad_colour"]).to(cuda_device)
    spec = render_sh_voxel_grid(grid, rays, make_cuda_config(case))
    diff = render_sh_voxel_grid(grid, rays, make_cuda_config(case, render_diffuse=True))
    ((spec.colour * gc).sum() + (diff.colour * gc).sum()).backward()
    total = grid.feature_storage.grad.clone()
    grid.zero_grad()
    (render_sh_voxel_grid(grid, rays, make_cuda_config(case)).colour * gc).sum().backward()
    g_spec = grid.feature_storage.grad.clone()
    grid.zero_grad()
    (render_sh_voxel_grid(grid, rays, make_cuda_config(case, render_diffuse=True)).colour * gc).sum().backward()
    assert rel_l2(total.cpu().numpy(), (g_spec + grid.feature_storage.grad).cpu().numpy()) < 1e-5


def test_disparity_gradient_matches_autograd(cuda_device):
    from oracle import torch_port as tp
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    dens = torch.from_numpy(inp["densities"]).requires_grad_(True)
    feat = torch.from_numpy(inp["features"]).requires_grad_(True)
    og = tp.OracleGrid(dens, feat, case.voxel_size, case.location, case.density_scale, case.density_pre, case.density_post)
    o, d = torch.from_numpy(inp["origins"]), torch.from_numpy(inp["directions"])
    ref = tp.render(og, o, d, num_samples=case.num_samples, near=case.near, far=case.far, white_bkgd=case.white_bkgd)
    mask = (ref["acc"] > 0.2).float()  # stay away from the 0/0 rays, whose gradient is NaN in the reference too
    gdisp = torch.from_numpy(np.random.RandomState(5).normal(size=(o.shape[0], 1)).astype(np.float32)) * mask
    (torch.nan_to_num(ref["disparity"]) * gdisp).sum().backward()
    grid = make_cuda_grid(case, inp, cuda_device)
    out = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case))
    (torch.nan_to_num(out.extra["disparity"]) * gdisp.to(cuda_device)).sum().backward()
    nf = inp["features"].shape[-1]
    assert rel_l2(grid.densities.grad.cpu().numpy(), dens.grad.numpy()) < 1e-4
    assert rel_l2(grid.feature_storage.grad[..., :nf].cpu().numpy(), feat.grad.numpy()) < 1e-4


def test_unsupported_configurations_fail_loudly(cuda_device):
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    rays = _rays(inp, cuda_device)
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(grid, rays, make_cuda_config(case, radiance_hdr_tone_map=torch.tanh))
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(grid, rays, make_cuda_config(case, stochastic_density_noise_std=1.0))
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(grid, rays, make_cuda_config(case, density2occupancy=lambda s, d: s * d))
    weird = VoxelGrid(torch.zeros(4, 4, 4, 1, device=cuda_device), torch.zeros(4, 4, 4, 3, device=cuda_device), VoxelSize(1, 1, 1),
                      density_preactivation=torch.exp)
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(weird, rays, make_cuda_config(case))
    with pytest.raises(AssertionError):  # flat rays only (reference render_interface.py:127-129)
        render_sh_voxel_grid(grid, Rays(rays.origins.reshape(40, 40, 3), rays.directions.reshape(40, 40, 3)), make_cuda_config(case))
    with pytest.raises(RuntimeError):  # CPU rays: no CPU fallback
        render_sh_voxel_grid(grid, Rays(rays.origins.cpu(), rays.directions.cpu()), make_cuda_config(case))
    bad = VoxelGrid(torch.zeros(4, 4, 4, 1, device=cuda_device), torch.zeros(4, 4, 4, 75, device=cuda_device), VoxelSize(1, 1, 1))
    with pytest.raises(RuntimeError, match="only degrees 0, 1, 2, and 3"):  # degree 4 (reference spherical_harmonics.py:79)
        render_sh_voxel_grid(bad, rays, make_cuda_config(case))


# ---------------------------------------------------------------------------------------------
# fused Adam (next-row f1) against torch.optim.Adam
# ---------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_adam(cuda_device):
    from thr3ed_atom_b200 import _kernels

    gen = torch.Generator(device="cpu").manual_seed(0)
    n = 4 * 1000 + 3
    p0 = torch.randn(n, generator=gen)
    p_ref = p0.clone().to(cuda_device).requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=0.03, betas=(0.9, 0.999))
    p = p0.clone().to(cuda_device)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, generator=gen).to(cuda_device) * (step % 3)
        # the kernel keeps the stores of 16-byte groups whose gradient and moments are all zero (identity update): elements
        # 400..799 never receive a gradient, 800..1199 only from step 3 on, 1200..1599 only at step 1 (their moments then decay),
        # and 1601..1602 sit in groups whose other elements are live
        g[400:800] = 0.0
        if step < 3:
            g[800:1200] = 0.0
        if step > 1:
            g[1200:1600] = 0.0
        g[1601:1603] = 0.0
        p_ref.grad = g.clone()
        opt.step()
        _kernels.adam_step(p, g, m, v, lr=0.03, beta1=0.9, beta2=0.999, eps=1e-8, step=step)
        np.testing.assert_allclose(p.cpu().numpy(), p_ref.detach().cpu().numpy(), rtol=2e-6, atol=1e-6)
    assert torch.equal(p[400:800].cpu(), p0[400:800]) and not bool(m[400:800].any()) and not bool(v[400:800].any())
    assert bool((m[1200:1600] != 0).all())  # decaying moments keep being written


# ---------------------------------------------------------------------------------------------
# BASELINE.json shapes: size-independent properties + oracle spot checks
# ---------------------------------------------------------------------------------------------
def _hotdog_setup(grid_n, deg, side, spp, device, density_shift=0.0, seed=42):
    from cases import relu_field_density_scale, spherical_pose, HOTDOG_RADIUS
    from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig
    from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose

    g = torch.Generator().manual_seed(seed)
    nf = 3 * (deg + 1) ** 2
    dens = torch.empty((grid_n,) * 3 + (1,)).uniform_(-1, 1, generator=g) + density_shift
    feat = torch.empty((grid_n,) * 3 + (nf,)).uniform_(-1, 1, generator=g)
    grid = VoxelGrid(dens.to(device), feat.to(device), VoxelSize(*(3 / grid_n,) * 3), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=relu_field_density_scale((3, 3, 3)), tunable=True)
    rot, trans = spherical_pose(30.0, 60.0, HOTDOG_RADIUS)
    rays = flatten_rays(cast_rays(CameraIntrinsics(side, side, 1111.11 * side / 800), CameraPose(rot, trans), device=device))
    cfg = SHVoxGridRenderConfig(spp, CameraBounds(1.8, 6.6), perturb_sampled_points=False, white_bkgd=True)
    return grid, rays, cfg, dens, feat


def test_config2_shape_against_oracle_on_a_ray_subset(cuda_device):
    """BASELINE configs[1]: 128^3 deg-2 grid, 400x400, 128 spp.  Full image on the GPU; 3000 strided rays re-rendered
    (forward + backward) by the fp32 CPU oracle on the same grid."""
    from oracle import torch_port as tp
    from cases import relu_field_density_scale
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    grid, rays, cfg, dens, feat = _hotdog_setup(128, 2, 400, 128, cuda_device)
    with torch.no_grad(), render_hints(image_hw=(400, 400)):
        full = render_sh_voxel_grid(grid, rays, cfg)
    assert float(full.extra["accumulated_weight"].max()) <= 1.0 + 1e-5 and float(full.colour.min()) >= -1e-6
    idx = torch.arange(17, 160000, 53)[:3000]
    o, d = rays.origins[idx.to(cuda_device)].contiguous(), rays.directions[idx.to(cuda_device)].contiguous()
    og = tp.OracleGrid(dens, feat, (3 / 128,) * 3, (0, 0, 0), relu_field_density_scale((3, 3, 3)), "identity", "relu")
    gc = torch.from_numpy(np.random.RandomState(0).normal(size=(idx.numel(), 3)).astype(np.float32))
    want = tp.render_with_grads(og, o.cpu(), d.cpu(), gc, num_samples=128, near=1.8, far=6.6, white_bkgd=True)
    # the same rays taken out of the full-image launch: per-ray results do not depend on the batch
    np.testing.assert_allclose(full.colour[idx.to(cuda_device)].cpu().numpy(), want["colour"].numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(full.depth[idx.to(cuda_device)].cpu().numpy(), want["depth"].numpy(), atol=2e-4, rtol=2e-5)
    sub = render_sh_voxel_grid(grid, Rays(o, d), cfg)
    assert torch.equal(sub.colour.detach(), full.colour[idx.to(cuda_device)])
    (sub.colour * gc.to(cuda_device)).sum().backward()
    assert rel_l2(grid.densities.grad.cpu().numpy(), want["grad_densities"].numpy()) < 1e-4
    assert rel_l2(grid.feature_storage.grad[..., :27].cpu().numpy(), want["grad_features"].numpy()) < 1e-4


def test_config3_shape_properties(cuda_device):
    """BASELINE configs[2]: 256^3 deg-2 grid, 800x800, 256 spp, forward + backward at full size.
    Size-independent properties: batch-split invariance (bit-exact forward, gradient of the halves sums to the
    gradient of the whole), linearity of the backward pass in the upstream gradient, exact zeros where nothing was
    sampled, bounded outputs, and an fp32-oracle spot check on 1024 rays."""
    from oracle import torch_port as tp
    from cases import relu_field_density_scale
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    grid, rays, cfg, dens, feat = _hotdog_setup(256, 2, 800, 256, cuda_device)
    n = len(rays)
    gen = torch.Generator().manual_seed(1)
    g1 = torch.randn((n, 3), generator=gen).to(cuda_device)
    g2 = torch.randn((n, 3), generator=gen).to(cuda_device)

    def grads(r, g, hint=None):
        grid.zero_grad()
        with render_hints(image_hw=hint):
            out = render_sh_voxel_grid(grid, r, cfg)
        (out.colour * g).sum().backward()
        return out, grid.densities.grad.clone(), grid.feature_storage.grad.clone()

    out, gd1, gf1 = grads(rays, g1, hint=(800, 800))
    acc = out.extra["accumulated_weight"]
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float(out.colour.min()) >= -1e-6 and float(out.colour.max()) <= 1.0 + 1e-5
    assert not bool(torch.isnan(out.colour).any()) and not bool(torch.isnan(gf1).any())
    assert not bool(gf1[..., 27:].any()), "padding lanes received gradient"

    # batch-split invariance
    half = n // 2
    oa, gda, gfa = grads(rays[:half], g1[:half])
    ob, gdb, gfb = grads(rays[half:], g1[half:])
    assert torch.equal(torch.cat([oa.colour, ob.colour]).detach(), out.colour.detach())
    assert float((gfa + gfb - gf1).norm() / gf1.norm()) < 1e-5
    assert float((gda + gdb - gd1).norm() / gd1.norm()) < 1e-5
    del gfa, gfb, gda, gdb, oa, ob

    # linearity in the upstream gradient
    _, gd2, gf2 = grads(rays, g2, hint=(800, 800))
    _, gd12, gf12 = grads(rays, g1 + g2, hint=(800, 800))
    assert float((gf1 + gf2 - gf12).norm() / gf12.norm()) < 1e-5
    assert float((gd1 + gd2 - gd12).norm() / gd12.norm()) < 1e-5
    del gf2, gf12, gd2, gd12

    # voxels no in-volume sample references get exactly zero gradient
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.thre3d_reprs.renderers import make_render_args

    touched = _kernels.mark_touched_voxels(grid.kernel_desc(), rays.origins, rays.directions, make_render_args(cfg)).bool()
    frac = float(touched.float().mean())
    assert 0.75 < frac < 0.85, frac  # SURVEY 8d measured 79.7 % for this view
    assert not bool(gf1[~touched].any()) and not bool(gd1[~touched].any())
    del gf1, gd1, touched
    grid.zero_grad()
    torch.cuda.empty_cache()

    # oracle spot check (forward + gradient of a 1024-ray sub-batch) on the same 256^3 grid
    idx = torch.arange(5, n, 611)[:1024]
    o, d = rays.origins[idx.to(cuda_device)].contiguous(), rays.directions[idx.to(cuda_device)].contiguous()
    og = tp.OracleGrid(dens, feat, (3 / 256,) * 3, (0, 0, 0), relu_field_density_scale((3, 3, 3)), "identity", "relu")
    gc = torch.from_numpy(np.random.RandomState(0).normal(size=(idx.numel(), 3)).astype(np.float32))
    want = tp.render_with_grads(og, o.cpu(), d.cpu(), gc, num_samples=256, near=1.8, far=6.6, white_bkgd=True)
    np.testing.assert_allclose(out.colour[idx.to(cuda_device)].detach().cpu().numpy(), want["colour"].numpy(), atol=2e-5, rtol=0)
    sub = render_sh_voxel_grid(grid, Rays(o, d), cfg)
    (sub.colour * gc.to(cuda_device)).sum().backward()
    assert rel_l2(grid.feature_storage.grad[..., :27].cpu().numpy(), want["grad_features"].numpy()) < 1e-4
    assert rel_l2(grid.densities.grad.cpu().numpy(), want["grad_densities"].numpy()) < 1e-4


def test_config3_shape_with_in_kernel_jitter_against_the_oracle(cuda_device):
    """The bench's own path at the bench's own shape (BASELINE configs[2], bench.py: perturb_sampled_points=True with the in-kernel
    counter-based jitter, image-tile ray order): 256^3 deg-2 grid, 800x800, 256 spp.  2048 strided rays of the full-image launch are
    replayed through the oracle with the same jitter (tests/helpers.py::hash_jitter, keyed by the ray's index in the launch):
    forward, and the grid gradient of a loss that only those rays feed."""
    from oracle import torch_port as tp
    from cases import relu_field_density_scale
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    grid, rays, cfg, dens, feat = _hotdog_setup(256, 2, 800, 256, cuda_device)
    cfg = dataclasses.replace(cfg, perturb_sampled_points=True)
    n, S, seed = len(rays), 256, 0x0BADC0FFEE
    idx = torch.arange(101, n, 307)[:2048]
    gc_sub = torch.from_numpy(np.random.RandomState(3).normal(size=(idx.numel(), 3)).astype(np.float32)).to(cuda_device)
    gc = torch.zeros((n, 3), device=cuda_device)
    gc[idx.to(cuda_device)] = gc_sub
    with render_hints(image_hw=(800, 800), rng_seed=seed):
        out = render_sh_voxel_grid(grid, rays, cfg)
        (out.colour * gc).sum().backward()
    u = torch.from_numpy(hash_jitter(seed, n, S, rays=idx.numpy())).to(cuda_device)
    assert 0.45 < float(u.mean()) < 0.55
    og = tp.OracleGrid(dens.to(cuda_device), feat.to(cuda_device), (3 / 256,) * 3, (0, 0, 0), relu_field_density_scale((3, 3, 3)), "identity", "relu")
    o, d = rays.origins[idx.to(cuda_device)].contiguous(), rays.directions[idx.to(cuda_device)].contiguous()
    want = tp.render_with_grads(og, o, d, gc_sub, num_samples=S, near=1.8, far=6.6, white_bkgd=True, jitter=u, ray_chunk=512)
    sel = idx.to(cuda_device)
    np.testing.assert_allclose(out.colour[sel].detach().cpu().numpy(), want["colour"].cpu().numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(out.depth[sel].detach().cpu().numpy(), want["depth"].cpu().numpy(), atol=2e-4, rtol=2e-5)
    assert rel_l2(grid.feature_storage.grad[..., :27].cpu().numpy(), want["grad_features"].cpu().numpy()) < 1e-4
    assert rel_l2(grid.densities.grad.cpu().numpy(), want["grad_densities"].cpu().numpy()) < 1e-4
    # the jitter really moved the samples: the deterministic render of the same rays differs
    with torch.no_grad(), render_hints(image_hw=(800, 800)):
        plain = render_sh_voxel_grid(grid, rays, dataclasses.replace(cfg, perturb_sampled_points=False))
    assert float((plain.colour[sel] - out.colour[sel].detach()).abs().max()) > 1e-3


@pytest.mark.parametrize("name", ["deg2_16cube", "deg3_abs", "deg1_aniso_softplus", "c1_32cube_deg0", "deg2_jitter_optimized"])
def test_cooperative_and_per_ray_kernels_agree(name, cuda_device, monkeypatch):
    """The warp-cooperative kernels against the thread-per-ray kernels (variant 3), with and without the forward's sample
    cache.  The staged forwards (variants 8, 4) do the per-ray arithmetic of the per-ray kernel => bit-identical images;
    the default lane-group forward sums the 8 corners and the SH terms in another order => equal to fp32 rounding.
    Gradients are equal up to summation order."""
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    from thr3ed_atom_b200 import _kernels

    if not _kernels.has_ab_variants():
        pytest.skip("product build: A/B kernel variants are compiled only with -DR3D_AB_VARIANTS (python -m thr3ed_atom_b200.build --ab)")
    case = CASES[name]
    inp = build_inputs(case)
    gc = torch.from_numpy(inp["grad_colour"]).to(cuda_device)
    results = {}
    for label, variant, cache_limit in (("group+cache", 0, None), ("per-ray+cache", 3, None), ("group", 0, "0"), ("per-ray", 3, "0"),
                                        ("staged+cache", 8, None), ("staged", 8, "0"), ("staged, TMA", 4, None),
                                        ("ws+cache", 32, None), ("ws", 32, "0"), ("ws-sort+cache", 96, None), ("ws, no quads", 32, None),
                                        ("ws-sort+L2 prefetch", 96 | 256, None), ("ws-sort+L1 prefetch", 96 | 512, None),
                                        ("group fwd, ws bwd", 128, None), ("ws fwd, ws bwd", 96 | 128, None),
                                        ("group-sort+cache", 2048, None), ("group-sort", 2048, "0"),
                                        ("ws-sort+dual streams", 96 | 1024, None), ("ws-sort+L2 prefetch+dual streams", 96 | 256 | 1024, None),
                                        ("split+cache", 32768, None), ("split, no quads", 32768, None), ("group-quads+cache", 65536, None)):
        if cache_limit is None:
            monkeypatch.delenv("R3D_SAMPLE_CACHE_MAX_BYTES", raising=False)
        else:
            monkeypatch.setenv("R3D_SAMPLE_CACHE_MAX_BYTES", cache_limit)
        monkeypatch.setenv("R3D_DENSITY_QUADS", "0" if "no quads" in label else "1")
        grid = make_cuda_grid(case, inp, cuda_device)
        hints = {"variant": variant}
        if case.jitter:
            hints["jitter"] = torch.from_numpy(inp["jitter"]).to(cuda_device)
        with render_hints(**hints):
            out = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case))
            (out.colour * gc).sum().backward()
        results[label] = (out.colour.detach().clone(), out.depth.detach().clone(), grid.densities.grad.clone(), grid.feature_storage.grad.clone())
    ref = results["per-ray"]
    for label, res in results.items():
        if label.startswith("ws"):  # warp-specialised forward: the lane-group arithmetic (up to FMA contraction), any publish order
            assert (res[0] - results["group"][0]).abs().max().item() < 1e-6, label
            assert (res[0] - results["ws"][0]).abs().max().item() < 1e-6, label  # with / without cache, sorted or not, quads or not
        if label.startswith("split") or label.startswith("group-quads"):  # the lane-group arithmetic (up to FMA contraction)
            assert (res[0] - results["group+cache"][0]).abs().max().item() < 1e-6, label
            assert ((res[1] - results["group+cache"][1]).abs() <= 1e-6 * results["group+cache"][1].abs().clamp(min=1.0)).all(), label
        if label.startswith("group") or label.startswith("ws") or label.startswith("split"):  # ("group-quads" included)
            assert (res[0] - ref[0]).abs().max().item() < 2e-6, label
            assert ((res[1] - ref[1]).abs() <= 2e-6 * ref[1].abs().clamp(min=1.0)).all(), label
        else:
            assert torch.equal(res[0], ref[0]) and torch.equal(res[1], ref[1]), label
        assert rel_l2(res[2].cpu().numpy(), ref[2].cpu().numpy()) < 1e-5, label
        assert rel_l2(res[3].cpu().numpy(), ref[3].cpu().numpy()) < 1e-5, label


@pytest.mark.parametrize("num_rays", [1, 31, 33, 129, 1000])
def test_ragged_batch_sizes(num_rays, cuda_device):
    """Batches that do not fill a warp / a CTA, incl. a single ray (the reference's .squeeze() breaks at one point,
    voxels.py:305; the kernels have no such restriction)."""
    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    idx = np.linspace(0, inp["origins"].shape[0] - 1, num_rays).astype(np.int64)
    sub = {k: (v[idx] if k in ("origins", "directions", "grad_colour") else v) for k, v in inp.items()}
    want = run_numpy_f64(case, sub)
    got = run_cuda_case(case, sub, cuda_device)
    assert_outputs_close(got, want, atol=1e-5, rtol_depth=1e-5, what=f"n={num_rays}")
    assert rel_l2(got["grad_features"], want["grad_features"]) < 5e-5
    assert rel_l2(got["grad_densities"], want["grad_densities"]) < 5e-5


def test_many_samples_per_ray_and_non_contiguous_ray_views(cuda_device):
    """1024 samples per ray (the reference's render default, renderers.py:44) on rays given as strided views."""
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid

    case = dataclasses.replace(CASES["deg2_16cube"], num_samples=1024, name="s1024")
    inp = build_inputs(case)
    keep = np.arange(0, inp["origins"].shape[0], 7)
    sub = {k: (v[keep] if k in ("origins", "directions", "grad_colour") else v) for k, v in inp.items()}
    want = run_numpy_f64(case, sub)
    grid = make_cuda_grid(case, inp, cuda_device)
    both = torch.from_numpy(np.concatenate([inp["origins"], inp["directions"]], -1)).to(cuda_device)  # [N, 6]
    rays = Rays(both[::7, :3], both[::7, 3:])  # non-contiguous views
    assert not rays.origins.is_contiguous()
    out = render_sh_voxel_grid(grid, rays, make_cuda_config(case))
    (out.colour * torch.from_numpy(sub["grad_colour"]).to(cuda_device)).sum().backward()
    got = {"colour": out.colour.detach().cpu().numpy(), "depth": out.depth.detach().cpu().numpy(),
           "acc": out.extra["accumulated_weight"].detach().cpu().numpy(), "disparity": out.extra["disparity"].detach().cpu().numpy()}
    assert_outputs_close(got, want, atol=1e-5, rtol_depth=1e-5, what="s1024")
    assert rel_l2(grid.feature_storage.grad[..., :27].cpu().numpy(), want["grad_features"]) < 5e-5


# ---------------------------------------------------------------------------------------------
# backward marching by the forward's contribution ballots (R3dRenderOut.sample_mask)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["deg2_16cube", "deg2_jitter", "deg2_sparse", "c1_32cube_deg0", "deg2_random_rays", "deg3_abs", "deg1_aniso_softplus",
                                  "deg3_abs+relu", "deg1_aniso_softplus+relu"])
def test_backward_by_contribution_ballots_equals_the_density_march(name, cuda_device, monkeypatch):
    """With a ReLU field the backward takes sigma from the forward's per-sample records and the set of contributing samples
    from its per-step ballots instead of repeating the inside test and the density gather: same samples, same sigma bits =>
    same gradients up to the order of the atomic sums.  Other post-activations keep the density march (the mask is then
    not even allocated); both are compared with R3D_SAMPLE_MASK=0."""
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.thre3d_reprs.renderers import make_render_args

    # "+relu": the degree-1 / degree-3 cases with a ReLU post-activation, so that those instantiations take the ballot march too
    case = dataclasses.replace(CASES[name[:-5]], density_post="relu") if name.endswith("+relu") else CASES[name]
    inp = build_inputs(case)
    res = {}
    for label, env in (("mask", "1"), ("march", "0")):
        monkeypatch.setenv("R3D_SAMPLE_MASK", env)
        res[label] = run_cuda_case(case, inp, cuda_device, with_grads=True, use_tile_hint=True)
    for key in ("colour", "depth", "acc"):
        assert np.array_equal(res["mask"][key], res["march"][key]), key
    assert rel_l2(res["mask"]["grad_features"], res["march"]["grad_features"]) < 1e-6
    assert rel_l2(res["mask"]["grad_densities"], res["march"]["grad_densities"]) < 1e-6
    grid = make_cuda_grid(case, inp, cuda_device)
    expect = case.density_post == "relu" and not case.diffuse
    assert _kernels.sample_mask_supported(grid.kernel_desc(), make_render_args(make_cuda_config(case))) == expect


def test_sample_mask_is_refused_when_another_forward_kernel_would_run(cuda_device):
    """The library never accepts a mask it would not write (per-ray / staged forward variants)."""
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.thre3d_reprs.renderers import make_render_args

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    desc = grid.kernel_desc()
    o, d = torch.from_numpy(inp["origins"]).to(cuda_device), torch.from_numpy(inp["directions"]).to(cuda_device)
    args = make_render_args(make_cuda_config(case))
    cache = _kernels.new_sample_cache(o.shape[0], args.num_samples, cuda_device)
    mask = _kernels.new_sample_mask(desc, o, d, o.shape[0], args)
    assert mask.shape == (args.num_samples, (o.shape[0] + 127) // 128 * 4)
    _kernels.render_forward(desc, o, d, args, cache, sample_mask=mask)  # default kernel: accepted
    args.variant = 8  # staged forward (A/B builds), refused outright by the product build
    with pytest.raises(RuntimeError, match="sample_mask needs|A/B variants are compiled only"):
        _kernels.render_forward(desc, o, d, args, cache, sample_mask=mask)
    args.variant = 0
    args.diffuse = True  # band-0-only renders use the per-ray kernel, which writes no ballots
    with pytest.raises(RuntimeError, match="sample_mask needs"):
        _kernels.render_forward(desc, o, d, args, cache, sample_mask=mask)
    args.diffuse = False
    with pytest.raises(RuntimeError, match="sample_mask needs"):
        _kernels.render_forward(desc, o, d, args, None, sample_mask=mask)


def test_no_grad_render_allocates_no_sample_cache(cuda_device, monkeypatch):
    """A render under torch.no_grad() on a TUNABLE grid (inference, chunked VolumetricModel.render) must not allocate the
    [S, N, 4] per-sample records or the ballots: no backward pass can follow (ctx.needs_input_grad alone is True there)."""
    import thr3ed_atom_b200.thre3d_reprs.renderers as renderers
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid

    case = CASES["deg2_16cube"]
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    assert grid.densities.requires_grad
    calls = {"cache": 0, "mask": 0}
    real_cache, real_mask = renderers._kernels.new_sample_cache, renderers._kernels.new_sample_mask

    def counting_cache(*a, **k):
        calls["cache"] += 1
        return real_cache(*a, **k)

    def counting_mask(*a, **k):
        calls["mask"] += 1
        return real_mask(*a, **k)

    monkeypatch.setattr(renderers._kernels, "new_sample_cache", counting_cache)
    monkeypatch.setattr(renderers._kernels, "new_sample_mask", counting_mask)
    with torch.no_grad():
        quiet = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case))
    assert calls == {"cache": 0, "mask": 0}
    loud = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case))
    assert calls["cache"] == 1
    assert torch.equal(quiet.colour, loud.colour.detach())


@pytest.mark.parametrize("tile", [(1, 1), (8, 4)])
def test_device_side_training_batch_sampler(tile, cuda_device):
    """one kernel replaces cast_rays on every view + randperm + gathers (reference trainers.py:281-303,
    utils/misc.py:117-129).  Every sampled ray / pixel equals cast_rays / the image at the pixel it names; pixels come in whole
    tiles; the draw is uniform over views and pixels and reproducible from the seed."""
    from cases import spherical_pose
    from thr3ed_atom_b200 import _kernels
    from thr3ed_atom_b200.rendering.volumetric.utils.misc import cast_rays, flatten_rays, sample_training_ray_batch
    from thr3ed_atom_b200.utils.imaging_utils import CameraIntrinsics, CameraPose

    v, h, w, focal = 5, 48, 64, 70.0
    poses = [CameraPose(*spherical_pose(37.0 * k, 30.0 + 7.0 * k, 4.0)) for k in range(v)]
    gen = torch.Generator().manual_seed(3)
    images = torch.rand((v, h, w, 3), generator=gen).to(cuda_device)
    intr = CameraIntrinsics(h, w, focal)
    batch = 8192
    rot = torch.stack([torch.as_tensor(p.rotation, dtype=torch.float32).reshape(3, 3) for p in poses]).to(cuda_device)
    tr = torch.stack([torch.as_tensor(p.translation, dtype=torch.float32).reshape(3) for p in poses]).to(cuda_device)
    o, d, px, idx = _kernels.sample_ray_batch(rot, tr, images, h, w, focal, batch, tile=tile, seed=99, want_indices=True)
    assert int(idx.min()) >= 0 and int(idx.max()) < v * h * w
    all_rays = [flatten_rays(cast_rays(intr, p, device=cuda_device)) for p in poses]
    ref_o = torch.cat([r.origins for r in all_rays])[idx]
    ref_d = torch.cat([r.directions for r in all_rays])[idx]
    assert torch.equal(o, ref_o) and torch.equal(d, ref_d)
    assert torch.equal(px, images.reshape(-1, 3)[idx])
    # whole tiles, row-major inside a tile, aligned to the tile grid
    x, y, view = idx % w, (idx // w) % h, idx // (h * w)
    tw, th = tile
    xs, ys, vs = x.reshape(-1, th, tw), y.reshape(-1, th, tw), view.reshape(-1, th, tw)
    assert bool((xs[:, :, :1] % tw == 0).all()) and bool((ys[:, :1, :] % th == 0).all())
    assert bool((xs == xs[:, :1, :1] + torch.arange(tw, device=cuda_device)[None, None, :]).all())
    assert bool((ys == ys[:, :1, :1] + torch.arange(th, device=cuda_device)[None, :, None]).all())
    assert bool((vs == vs[:, :1, :1]).all())
    # uniform over views and over the image (loose bounds: 8192 / (tw * th) independent draws)
    counts = torch.bincount(view, minlength=v).float() / batch
    assert float((counts - 1.0 / v).abs().max()) < (0.03 if tile == (1, 1) else 0.12)
    assert abs(float(x.float().mean()) / (w - 1) - 0.5) < 0.06 and abs(float(y.float().mean()) / (h - 1) - 0.5) < 0.06
    # reproducible from the seed, different with another one; the public helper returns Rays + pixels
    o2, _, _, idx2 = _kernels.sample_ray_batch(rot, tr, images, h, w, focal, batch, tile=tile, seed=99, want_indices=True)
    _, _, _, idx3 = _kernels.sample_ray_batch(rot, tr, images, h, w, focal, batch, tile=tile, seed=100, want_indices=True)
    assert torch.equal(idx, idx2) and not torch.equal(idx, idx3)
    rays, pixels = sample_training_ray_batch(poses, intr, images, 1024, tile=tile, seed=5)
    assert tuple(rays.origins.shape) == (1024, 3) and tuple(pixels.shape) == (1024, 3)
    with pytest.raises(RuntimeError, match="multiple of the tile size"):
        _kernels.sample_ray_batch(rot, tr, images, h, w, focal, 33, tile=(8, 4))
