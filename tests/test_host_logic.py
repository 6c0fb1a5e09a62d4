"""Host-side mirror of the reference interface, on CPU tensors (no kernels run here): VoxelGrid storage and
state-dict contract, config handling, VolumetricModel bookkeeping / checkpoints, camera helpers, ray glue,
the loud refusal to render on CPU, and the 2-process (gloo) gradient all-reduce plumbing."""
import dataclasses
import os
import pickle

import numpy as np
import pytest
import torch

from thr3ed_atom_b200 import _abi
from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel, create_volumetric_model_from_saved_model
from thr3ed_atom_b200.rendering.volumetric.accumulate import density2occupancy_pb
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays, RenderOut, render
from thr3ed_atom_b200.rendering.volumetric.utils.misc import (
    collate_rays,
    collate_rendered_output,
    compute_expected_density_scale_for_relu_field_grid,
    flatten_rays,
    reshape_rendered_output,
    sample_random_rays_and_pixels_synchronously,
)
from thr3ed_atom_b200.thre3d_reprs.renderers import SHVoxGridRenderConfig, make_render_args, render_hints, render_sh_voxel_grid
from thr3ed_atom_b200.thre3d_reprs.voxels import (
    VoxelGrid,
    VoxelGridLocation,
    VoxelSize,
    classify_density_activations,
    create_voxel_grid_from_saved_info_dict,
    padded_feature_stride,
    scale_voxel_grid_with_required_output_size,
)
from thr3ed_atom_b200.utils.imaging_utils import (
    CameraBounds,
    CameraIntrinsics,
    adjust_dynamic_range,
    get_thre360_animation_poses,
    get_thre360_spiral_animation_poses,
    pose_spherical,
    range_map_coefficients,
    scale_camera_intrinsics,
)


def _grid(deg=2, dims=(4, 5, 6), tunable=True, **kw):
    g = torch.Generator().manual_seed(0)
    nf = 3 * (deg + 1) ** 2
    return VoxelGrid(
        densities=torch.rand((*dims, 1), generator=g),
        features=torch.rand((*dims, nf), generator=g),
        voxel_size=VoxelSize(0.5, 0.4, 0.3),
        grid_location=VoxelGridLocation(0.1, -0.2, 0.3),
        density_preactivation=torch.nn.Identity(),
        density_postactivation=torch.nn.ReLU(),
        expected_density_scale=33.0,
        tunable=tunable,
        **kw,
    )


# ---------------------------------------------------------------------------- VoxelGrid
@pytest.mark.parametrize("deg,stride", [(0, 4), (1, 12), (2, 28), (3, 48)])
def test_feature_storage_is_padded_to_whole_16_byte_vectors(deg, stride):
    grid = _grid(deg)
    nf = 3 * (deg + 1) ** 2
    assert padded_feature_stride(nf) == stride
    assert tuple(grid.feature_storage.shape) == (4, 5, 6, stride) and grid.feature_storage.is_contiguous()
    assert tuple(grid.features.shape) == (4, 5, 6, nf)
    assert tuple(grid.densities.shape) == (4, 5, 6, 1)
    assert grid.grid_dims == (4, 5, 6) and (grid.width_x, grid.depth_y, grid.height_z) == (4, 5, 6)
    # parameters are the leaves Adam sees (reference trainers.py:238-245): densities + (padded) features
    assert [tuple(p.shape) for p in grid.parameters()] == [(4, 5, 6, 1), (4, 5, 6, stride)]
    # in-place initialisation through the getters (reference trainers.py:151-152) writes through the view
    with torch.no_grad():
        torch.nn.init.uniform_(grid.features, 2.0, 3.0)
        torch.nn.init.uniform_(grid.densities, -1.0, 1.0)
    assert float(grid.feature_storage[..., :nf].min()) >= 2.0
    if stride != nf:
        assert not bool(grid.feature_storage[..., nf:].any())


def test_state_dict_uses_the_reference_keys_and_shapes_and_round_trips():
    grid = _grid(2)
    sd = grid.state_dict()
    assert list(sd.keys()) == ["_densities", "_features"]  # reference thre3d_reprs/constants.py:10-11
    assert tuple(sd["_features"].shape) == (4, 5, 6, 27) and tuple(sd["_densities"].shape) == (4, 5, 6, 1)
    assert torch.equal(sd["_features"], grid.features.detach())
    other = _grid(2)
    with torch.no_grad():
        other.features.zero_()
    other.load_state_dict(sd)
    assert torch.equal(other.features, grid.features) and torch.equal(other.densities, grid.densities)
    assert not bool(other.feature_storage[..., 27:].any())
    # non-tunable grids keep plain tensors (not parameters / not in the state dict), as in the reference
    frozen = _grid(2, tunable=False)
    assert list(frozen.parameters()) == [] and list(frozen.state_dict().keys()) == []


def test_setters_check_shapes_and_rewrap_parameters():
    grid = _grid(2)
    new_f = torch.ones(4, 5, 6, 27)
    grid.features = new_f
    assert isinstance(grid.feature_storage, torch.nn.Parameter) and torch.equal(grid.features, new_f)
    grid.densities = torch.full((4, 5, 6, 1), 2.0)
    assert isinstance(grid.densities, torch.nn.Parameter) and float(grid.densities.mean()) == 2.0
    with pytest.raises(AssertionError):
        grid.features = torch.ones(4, 5, 6, 12)
    with pytest.raises(AssertionError):
        grid.densities = torch.ones(4, 5, 7, 1)
    with pytest.raises(AssertionError):
        VoxelGrid(torch.zeros(2, 2, 2), torch.zeros(2, 2, 2, 3), VoxelSize())


def test_aabb_config_dicts_and_kernel_descriptor():
    grid = _grid(2)
    aabb = grid.aabb
    assert aabb.x_range == pytest.approx((0.1 - 1.0, 0.1 + 1.0)) and aabb.y_range == pytest.approx((-1.2, 0.8))
    assert aabb.z_range == pytest.approx((0.3 - 0.9, 0.3 + 0.9))
    verts = grid.get_bounding_volume_vertices()
    assert tuple(verts.shape) == (8, 3) and float(verts[:, 0].min()) == pytest.approx(-0.9)
    cfg = grid.get_config_dict()
    assert set(cfg) == {"grid_location", "density_preactivation", "density_postactivation", "feature_preactivation",
                        "feature_postactivation", "radiance_transfer_function", "expected_density_scale", "tunable"}
    assert set(grid.get_save_config_dict()) == set(cfg) | {"voxel_size"}
    inside = grid.test_inside_volume(torch.tensor([[0.1, -0.2, 0.3], [-0.9, 0.0, 0.0], [5.0, 0.0, 0.0]]))
    assert inside.squeeze(-1).tolist() == [True, False, False]  # strict inequalities: a point ON the plane is outside
    desc = grid.kernel_desc()
    assert (desc.density_pre, desc.density_post) == (_abi.PRE_IDENTITY, _abi.POST_RELU)
    s, b = range_map_coefficients(aabb.x_range, (-1.0, 1.0))
    assert desc.norm_scale[0] == float(s) and desc.norm_bias[0] == float(b)
    # the voxel_size setter does not move the bounding box (reference voxels.py:166-168)
    grid.voxel_size = VoxelSize(1.0, 1.0, 1.0)
    assert grid.aabb == aabb and grid.voxel_size == VoxelSize(1.0, 1.0, 1.0)
    assert "grid_dims: (4, 5, 6)" in repr(grid)


def test_activation_classification():
    assert classify_density_activations(torch.nn.Identity(), torch.nn.ReLU()) == (_abi.PRE_IDENTITY, _abi.POST_RELU)
    assert classify_density_activations(torch.nn.Identity(), torch.nn.Softplus()) == (_abi.PRE_IDENTITY, _abi.POST_SOFTPLUS)
    assert classify_density_activations(torch.abs, torch.nn.Identity()) == (_abi.PRE_ABS, _abi.POST_IDENTITY)
    with pytest.raises(NotImplementedError):
        classify_density_activations(torch.exp, torch.nn.Identity())
    with pytest.raises(NotImplementedError):
        classify_density_activations(torch.nn.Identity(), torch.nn.Softplus(beta=2.0))
    with pytest.raises(NotImplementedError):
        _grid(2, feature_postactivation=torch.nn.Sigmoid()).kernel_desc()


def test_rescaling_a_grid_matches_trilinear_interpolate():
    grid = _grid(1, dims=(4, 4, 4))
    bigger = scale_voxel_grid_with_required_output_size(grid, (8, 6, 5))
    assert bigger.grid_dims == (8, 6, 5) and isinstance(bigger.densities, torch.nn.Parameter)
    assert bigger.voxel_size == pytest.approx((0.5 * 4 / 8, 0.4 * 4 / 6, 0.3 * 4 / 5))
    both = torch.cat([grid.features, grid.densities], -1).permute(3, 0, 1, 2)[None]
    want = torch.nn.functional.interpolate(both, size=(8, 6, 5), mode="trilinear", align_corners=False)[0].permute(1, 2, 3, 0)
    assert torch.allclose(bigger.features, want[..., :-1]) and torch.allclose(bigger.densities, want[..., -1:])
    assert bigger.aabb.x_range == pytest.approx(grid.aabb.x_range)


# ---------------------------------------------------------------------------- config / model facade
def test_render_config_has_the_reference_fields_in_order():
    names = [f.name for f in dataclasses.fields(SHVoxGridRenderConfig)]
    assert names == ["num_samples_per_ray", "camera_bounds", "perturb_sampled_points", "optimized_sampling", "density2occupancy",
                     "radiance_hdr_tone_map", "stochastic_density_noise_std", "white_bkgd", "render_diffuse",
                     "render_num_samples_per_ray", "parallel_rays_chunk_size"]  # reference renderers.py:28-45
    cfg = SHVoxGridRenderConfig(64, CameraBounds(1.0, 2.0))
    assert (cfg.perturb_sampled_points, cfg.optimized_sampling, cfg.white_bkgd, cfg.render_diffuse) == (True, False, False, False)
    assert cfg.density2occupancy is density2occupancy_pb and cfg.radiance_hdr_tone_map is torch.sigmoid
    assert (cfg.stochastic_density_noise_std, cfg.render_num_samples_per_ray, cfg.parallel_rays_chunk_size) == (0.0, 1024, 32768)
    with render_hints(image_hw=(2, 3), rng_seed=5):
        args = make_render_args(cfg)
    assert args.perturb and args.rng_seed == 5 and args.image_hw == (2, 3) and args.flags() == _abi.FLAG_PERTURB
    torch.manual_seed(3)
    a = make_render_args(cfg).rng_seed
    torch.manual_seed(3)
    assert make_render_args(cfg).rng_seed == a  # default seed follows torch.manual_seed
    flags = make_render_args(dataclasses.replace(cfg, perturb_sampled_points=False, white_bkgd=True, render_diffuse=True, optimized_sampling=True)).flags()
    assert flags == _abi.FLAG_WHITE_BKGD | _abi.FLAG_DIFFUSE | _abi.FLAG_OPTIMIZED_SAMPLING
    assert float(density2occupancy_pb(torch.tensor(2.0), torch.tensor(0.5))) == pytest.approx(1 - np.exp(-1.0))


def test_volumetric_model_bookkeeping_and_checkpoint_round_trip(tmp_path):
    grid = _grid(2)
    cfg = SHVoxGridRenderConfig(32, CameraBounds(1.8, 6.6), white_bkgd=True)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=torch.device("cpu"))
    assert vol_mod.render_procedure is render_sh_voxel_grid and vol_mod.thre3d_repr is grid and vol_mod.render_config is cfg
    updated = VolumetricModel._update_render_config(cfg, {"render_diffuse": True, "num_samples_per_ray": 8})
    assert updated.render_diffuse and updated.num_samples_per_ray == 8 and not cfg.render_diffuse  # original untouched
    with pytest.raises(ValueError, match="Unknown render configuration field"):
        VolumetricModel._update_render_config(cfg, {"nope": 1})
    info = vol_mod.get_save_info(extra_info={"camera_bounds": CameraBounds(1.8, 6.6), "hemispherical_radius": 4.03})
    assert set(info) == {"thre3d_repr", "render_procedure", "render_config_type", "render_config", "extra_info"}
    assert info["render_procedure"] is render_sh_voxel_grid and info["render_config_type"] is SHVoxGridRenderConfig
    path = tmp_path / "model.pth"
    torch.save(info, path)
    loaded, extra = create_volumetric_model_from_saved_model(path, create_voxel_grid_from_saved_info_dict, device=torch.device("cpu"))
    assert extra["hemispherical_radius"] == 4.03 and loaded.render_config == cfg
    assert torch.equal(loaded.thre3d_repr.features, grid.features) and torch.equal(loaded.thre3d_repr.densities, grid.densities)
    assert loaded.thre3d_repr.voxel_size == grid.voxel_size and loaded.render_procedure is render_sh_voxel_grid
    # the procedure and config are pickled by qualified name
    assert pickle.loads(pickle.dumps(render_sh_voxel_grid)) is render_sh_voxel_grid


def test_rendering_cpu_tensors_is_refused_not_emulated():
    grid = _grid(2)
    rays = Rays(torch.zeros(5, 3), torch.ones(5, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        render_sh_voxel_grid(grid, rays, SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        grid(torch.zeros(4, 3))
    with pytest.raises(AssertionError, match="FLAT RAYS"):
        render_sh_voxel_grid(grid, Rays(torch.zeros(2, 2, 3), torch.ones(2, 2, 3)), SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0)))


def test_single_pass_specular_and_diffuse_render_host_contract():
    """render_sh_voxel_grid_with_diffuse: same refusals as the single render (no CPU emulation, flat rays), the config
    must not ask for a diffuse-only render, VolumetricModel forwards config overrides and rejects unknown ones."""
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid_with_diffuse

    grid = _grid(2)
    rays = Rays(torch.zeros(5, 3), torch.ones(5, 3))
    cfg = SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        render_sh_voxel_grid_with_diffuse(grid, rays, cfg)
    with pytest.raises(ValueError, match="render_diffuse must be False"):
        render_sh_voxel_grid_with_diffuse(grid, rays, SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0), render_diffuse=True))
    with pytest.raises(AssertionError, match="FLAT RAYS"):
        render_sh_voxel_grid_with_diffuse(grid, Rays(torch.zeros(2, 2, 3), torch.ones(2, 2, 3)), cfg)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=torch.device("cpu"))
    with pytest.raises(ValueError, match="Unknown render configuration field"):
        vol_mod.render_rays_with_diffuse(rays, not_a_field=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vol_mod.render_rays_with_diffuse(rays, num_samples_per_ray=4)


def test_contribution_ballots_are_requested_only_where_the_library_writes_and_uses_them():
    """``_kernels.sample_mask_supported`` mirrors the dispatch in csrc/r3d_render.cu: the per-step ballots exist only for the
    default forward kernel (padded layout, all SH bands, variant 0) and only help a ReLU-field backward."""
    from thr3ed_atom_b200 import _kernels

    cfg = SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0))
    relu = _grid(2).kernel_desc()
    assert _kernels.sample_mask_supported(relu, make_render_args(cfg))
    assert not _kernels.sample_mask_supported(relu, make_render_args(dataclasses.replace(cfg, render_diffuse=True)))
    with render_hints(variant=8):
        assert not _kernels.sample_mask_supported(relu, make_render_args(cfg))
    softplus = VoxelGrid(
        densities=torch.rand(4, 5, 6, 1), features=torch.rand(4, 5, 6, 27), voxel_size=VoxelSize(0.5, 0.4, 0.3),
        density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.Softplus(), tunable=True,
    ).kernel_desc()
    assert not _kernels.sample_mask_supported(softplus, make_render_args(cfg))
    unpadded = dataclasses.replace(relu, features=torch.rand(4, 5, 6, 27))
    assert not _kernels.sample_mask_supported(unpadded, make_render_args(cfg))


# ---------------------------------------------------------------------------- ray / output glue, cameras
def test_ray_and_output_containers():
    rays = Rays(torch.arange(24.0).reshape(2, 4, 3), torch.ones(2, 4, 3))
    flat = flatten_rays(rays)
    assert len(flat) == 8 and tuple(flat[2:5].origins.shape) == (3, 3)
    both = collate_rays([flat, flat])
    assert len(both) == 16
    with pytest.raises(AssertionError):
        Rays(torch.zeros(3, 3), torch.zeros(4, 3))
    pix = torch.arange(16.0)[:, None].repeat(1, 3)
    sub_rays, sub_pix = sample_random_rays_and_pixels_synchronously(both, pix, 5)
    assert len(sub_rays) == 5 and tuple(sub_pix.shape) == (5, 3)
    idx = sub_pix[:, 0].long()
    assert torch.equal(sub_rays.origins, both.origins[idx])  # rays and pixels stay in sync
    out = RenderOut(torch.zeros(8, 3), torch.zeros(8, 1), {"disparity": torch.ones(8, 1)})
    merged = collate_rendered_output([out, out])
    assert tuple(merged.colour.shape) == (16, 3) and tuple(merged.extra["disparity"].shape) == (16, 1)
    img = reshape_rendered_output(merged, CameraIntrinsics(4, 4, 10.0))
    assert tuple(img.colour.shape) == (4, 4, 3) and tuple(img.extra["disparity"].shape) == (4, 4, 1)
    assert RenderOut(torch.zeros(1, 3), torch.zeros(1, 1)).extra == {}
    with pytest.raises(AssertionError):
        RenderOut(torch.zeros(2, 4), torch.zeros(2, 1))
    # the generic 3-stage driver still composes user stages
    res = render(flat, CameraBounds(0, 1), 2, lambda r, b, n: ("s", n), lambda s, r: ("p", s), lambda p, r: RenderOut(torch.zeros(8, 3), torch.zeros(8, 1)))
    assert isinstance(res, RenderOut)
    assert compute_expected_density_scale_for_relu_field_grid((3.0, 3.0, 3.0)) == pytest.approx(100.0 / 3.0)


def test_camera_helpers():
    from cases import spherical_pose

    pose = pose_spherical(30.0, 60.0, 4.031128)
    rot, trans = spherical_pose(30.0, 60.0, 4.031128)
    np.testing.assert_allclose(pose.rotation.numpy(), rot, atol=1e-7)
    np.testing.assert_allclose(pose.translation.numpy(), trans, atol=1e-6)
    assert float(torch.linalg.norm(pose.translation)) == pytest.approx(4.031128, rel=1e-6)
    assert scale_camera_intrinsics(CameraIntrinsics(800, 800, 1111.11), 0.5) == CameraIntrinsics(400, 400, 555.555)
    assert len(get_thre360_animation_poses(4.0, 60.0, 9)) == 8
    assert len(get_thre360_spiral_animation_poses((1.0, 3.0), 2.0, 2, 11)) == 10
    x = np.array([0.0, 5.0, 10.0], np.float32)
    np.testing.assert_allclose(adjust_dynamic_range(x, (0, 10), (-1, 1), slack=True), [-1, 0, 1], atol=1e-6)
    np.testing.assert_allclose(adjust_dynamic_range(x, (0, 5), (0, 1)), [0, 1, 1], atol=1e-6)  # slack=False clips


# ---------------------------------------------------------------------------- 2-process data parallelism (gloo)
def _dp_worker(rank, world, port, tmp):
    import torch.distributed as dist

    from helpers import CASES, build_inputs
    from oracle import torch_port as tp
    from thr3ed_atom_b200.distributed import all_reduce_grid_gradients, broadcast_grid, shard_rays, shard_views

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = CASES["deg2_16cube"]
        inp = build_inputs(case)
        n = 301  # not divisible by the world size
        o, d = torch.from_numpy(inp["origins"][:n]), torch.from_numpy(inp["directions"][:n])
        pixels = torch.from_numpy(np.random.RandomState(0).uniform(size=(n, 3)).astype(np.float32))
        grid = VoxelGrid(torch.from_numpy(inp["densities"]) + rank, torch.from_numpy(inp["features"]) + rank, VoxelSize(*case.voxel_size),
                         density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                         expected_density_scale=case.density_scale, tunable=True)
        broadcast_grid(grid, src=0)  # replicas must start identical
        shard = shard_rays(Rays(o, d), pixels)
        assert shard.end - shard.start in (150, 151) and shard.total == n
        assert shard_views(8) == ([0, 1, 2, 3] if rank == 0 else [4, 5, 6, 7])

        def oracle_loss(dens, feat, rays_o, rays_d, px):  # stand-in renderer for the CPU test: the oracle (test infra)
            og = tp.OracleGrid(dens, feat, case.voxel_size, case.location, case.density_scale, "identity", "relu")
            out = tp.render(og, rays_o, rays_d, num_samples=16, near=case.near, far=case.far, white_bkgd=True)
            return torch.nn.functional.l1_loss(out["colour"], px)

        loss = oracle_loss(grid.densities, grid.features, shard.rays.origins, shard.rays.directions, shard.pixels) * shard.loss_weight
        loss.backward()
        all_reduce_grid_gradients(grid)
        # single-process reference: the mean over ALL rays
        dens = torch.from_numpy(inp["densities"]).requires_grad_(True)
        feat = torch.from_numpy(inp["features"]).requires_grad_(True)
        oracle_loss(dens, feat, o, d, pixels).backward()
        assert torch.allclose(grid.densities.grad, dens.grad, atol=1e-7, rtol=1e-5)
        assert torch.allclose(grid.feature_storage.grad[..., :27], feat.grad, atol=1e-7, rtol=1e-5)
        assert not bool(grid.feature_storage.grad[..., 27:].any())
        torch.save(torch.tensor(1), os.path.join(tmp, f"ok{rank}"))
    finally:
        dist.destroy_process_group()


def test_two_process_ray_sharding_and_gradient_all_reduce(tmp_path):
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_cover_everything_once():
    from thr3ed_atom_b200.distributed import shard_bounds

    for n in (0, 1, 7, 640000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_reference_import_paths_resolve_through_the_compat_shim(monkeypatch):
    import importlib
    import sys
    from pathlib import Path

    monkeypatch.syspath_prepend(str(Path(__file__).resolve().parent.parent / "compat"))
    for name in [m for m in sys.modules if m == "thre3d_atom" or m.startswith("thre3d_atom.")]:
        monkeypatch.delitem(sys.modules, name)
    from thre3d_atom.modules.volumetric_model import VolumetricModel as VM  # noqa: E402
    from thre3d_atom.rendering.volumetric.render_interface import Rays as R  # noqa: E402
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig as C, render_sh_voxel_grid as f  # noqa: E402
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid as VG  # noqa: E402
    from thre3d_atom.utils.imaging_utils import CameraBounds as CB  # noqa: E402

    assert f is render_sh_voxel_grid and C is SHVoxGridRenderConfig and VG is VoxelGrid and VM is VolumetricModel
    assert R is Rays and CB is CameraBounds
    assert importlib.import_module("thre3d_atom.thre3d_reprs.constants").u_FEATURES == "_features"


def test_bench_reports_the_dominant_kernel_in_the_roofline_entry():
    """bench.py: ``roofline`` is the entry of the render kernel with the longer mean launch; both kernels keep their own keys."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("r3d_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    stats = {"samples_visited": 1000, "samples_inside": 600, "corner_refs": 4700, "samples_contributing": 300}
    traffic = {"render_fwd_dram_bytes": 3, "render_bwd_dram_bytes": 5, "source": "ncu"}
    r = bench.roofline_entries(1_500_000_000, 4.5, 4_500_000_000, 4.4, 6500.0, "measured", traffic, stats, 112, 27, 1965.0)
    assert r["roofline"]["kernel"] == "render_fwd_group_kernel" and r["roofline"] == r["roofline_fwd"]
    assert r["roofline_bwd"]["traffic"] == 5 and r["roofline_fwd"]["traffic"] == 3
    assert abs(r["roofline_bwd"]["achieved"] - 4.5e9 / 4.4e-3 / 1e9) < 1e-6 and abs(r["roofline_bwd"]["frac"] - r["roofline_bwd"]["achieved"] / 6500.0) < 1e-12
    # SURVEY.md 8d companion figures travel with every entry: gather-request bytes and the FP32 work against the SIMT peak
    assert r["roofline_fwd"]["gather_request_bytes"] == 4700 * 112
    assert abs(r["roofline_fwd"]["fp32_peak_tflops"] - 148 * 128 * 2 * 1.965e9 / 1e12) < 1e-9
    assert 0.0 < r["roofline_fwd"]["fp32_frac"] < 1.0 and r["roofline_fwd"]["fp32_flops"] > r["roofline_bwd"]["fp32_flops"] * 0.5
    r = bench.roofline_entries(1_500_000_000, 4.0, 4_500_000_000, 4.4, 6500.0, "measured", {"source": "stale"}, stats, 112, 27, 1965.0)
    assert r["roofline"]["kernel"] == "render_bwd_coop_kernel" and r["roofline"]["traffic"] is None
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r["roofline"]


def test_bench_stdout_carries_only_the_json_line():
    """bench.py reserves file descriptor 1 for its one JSON line: whatever a library prints to stdout during the run (NCCL's
    version banner did) lands on stderr."""
    import json
    import subprocess
    import sys

    code = ("import bench, os; bench.reserve_stdout_for_the_json_line(); print('python-level noise'); "
            "os.system('echo child-process noise'); bench.emit({'metric': 'x', 'value': 1.5})")
    repo = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=repo, timeout=300)
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout) == {"metric": "x", "value": 1.5} and p.stdout.count("\n") == 1
    assert "python-level noise" in p.stderr and "child-process noise" in p.stderr


def test_bench_sm_side_ceiling_is_absent_with_a_reason_not_with_an_error(monkeypatch):
    """The self-measured companion of roofline_fwd needs the measurement build of the CURRENT sources; without it (or for another
    workload) the bench line carries the reason, never a stale number and never an exception."""
    import bench
    from thr3ed_atom_b200 import build as _build

    assert "c3 workload only" in bench.sm_side_ceiling("c2_128cube_deg2_400px_128spp", 1.0)["unavailable"]
    monkeypatch.setattr(_build, "ab_is_current", lambda: False)
    out = bench.sm_side_ceiling("c3_256cube_deg2_800px_256spp", 4.4)
    assert "build --ab" in out["unavailable"] and "no_arithmetic_two_kernel_forward_ms" not in out


def test_bench_refuses_stale_dram_traffic_numbers(tmp_path, monkeypatch):
    """profiles/traffic.json is stamped with the digest of the library sources it was measured with; any other digest => null."""
    import importlib.util
    import json
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("r3d_bench2", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from thr3ed_atom_b200 import build as _build

    (tmp_path / "profiles").mkdir()
    monkeypatch.setattr(bench, "ROOT", tmp_path)
    path = tmp_path / "profiles" / "traffic.json"
    path.write_text(json.dumps({"lib_digest": "not-the-current-sources", "w": {"render_fwd_dram_bytes": 1}}))
    assert "render_fwd_dram_bytes" not in bench.load_traffic("w") and "stale" in bench.load_traffic("w")["source"]
    path.write_text(json.dumps({"lib_digest": _build._source_digest(), "measured": "r02", "w": {"render_fwd_dram_bytes": 1, "render_bwd_dram_bytes": 2}}))
    assert bench.load_traffic("w")["render_fwd_dram_bytes"] == 1 and bench.load_traffic("w")["render_bwd_dram_bytes"] == 2


def test_compat_shim_hosts_the_reference_trainer(tmp_path):
    """with $THRE3D_ATOM_REFERENCE the shim imports the reference's OWN trainer / datasets / visualisations
    under the ``thre3d_atom`` package while the hot-path modules resolve to the B200 implementation, so the identity checks
    of reference modules/trainers.py:116-122 hold for a B200 VolumetricModel.  Runs in a subprocess (import-system state);
    skipped where no reference checkout is present (the GPU box)."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    import pytest

    reference = os.environ.get("THRE3D_ATOM_REFERENCE", "/root/reference")
    if not (Path(reference) / "thre3d_atom" / "modules" / "trainers.py").is_file():
        pytest.skip("no reference checkout (set $THRE3D_ATOM_REFERENCE)")
    root = Path(__file__).resolve().parent.parent
    code = r'''
import sys, inspect
sys.path.insert(0, sys.argv[1] + "/compat"); sys.path.insert(0, sys.argv[1])
import thre3d_atom
from thre3d_atom.modules import trainers                     # the reference's own file ...
assert trainers.__file__.startswith(sys.argv[2]), trainers.__file__
import thr3ed_atom_b200.thre3d_reprs.renderers as b200_renderers
import thr3ed_atom_b200.thre3d_reprs.voxels as b200_voxels
import thr3ed_atom_b200.modules.volumetric_model as b200_vm
assert trainers.render_sh_voxel_grid is b200_renderers.render_sh_voxel_grid        # ... bound to the B200 hot path
assert trainers.VoxelGrid is b200_voxels.VoxelGrid and trainers.VolumetricModel is b200_vm.VolumetricModel
assert trainers.scale_voxel_grid_with_required_output_size is b200_voxels.scale_voxel_grid_with_required_output_size
src = inspect.getsource(trainers.train_sh_vox_grid_vol_mod_with_posed_images)
assert "render_procedure == render_sh_voxel_grid" in src.replace("\n", " ").replace("  ", " ") or "render_sh_voxel_grid" in src
# names the B200 modules do not define come from the checkout's module of the same name (visualisation helpers)
from thre3d_atom.utils.imaging_utils import postprocess_depth_map, CameraBounds
from thre3d_atom.rendering.volumetric.utils.misc import ndcize_rays, cast_rays
import thr3ed_atom_b200.utils.imaging_utils as b200_iu
assert CameraBounds is b200_iu.CameraBounds and postprocess_depth_map.__module__.startswith("thre3d_atom._reference")
import numpy as np
img = postprocess_depth_map(np.linspace(0, 1, 16, dtype=np.float32).reshape(4, 4, 1))
assert img.shape == (4, 4, 3) and img.dtype == np.uint8
from thre3d_atom.data.datasets import PosedImagesDataset       # reference data layer, reference visualisations
from thre3d_atom.visualizations.static import visualize_sh_vox_grid_vol_mod_rendered_feedback
from thre3d_atom.utils.misc import compute_thre3d_grid_sizes
assert compute_thre3d_grid_sizes((128, 128, 128), 4, 2.0)[-1] == (128, 128, 128)
# the identity asserts themselves, on a B200 model (no render call: this box has no GPU)
import torch
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid, SHVoxGridRenderConfig
from thre3d_atom.modules.volumetric_model import VolumetricModel
grid = VoxelGrid(torch.zeros(4, 4, 4, 1), torch.zeros(4, 4, 4, 3), VoxelSize(0.5, 0.5, 0.5), tunable=True)
vm = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0)), device=torch.device("cpu"))
assert isinstance(vm.thre3d_repr, trainers.VoxelGrid) and vm.render_procedure == trainers.render_sh_voxel_grid
print("SHIM_OK")
'''
    env = dict(os.environ, THRE3D_ATOM_REFERENCE=reference)
    out = subprocess.run([sys.executable, "-c", code, str(root), reference], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "SHIM_OK" in out.stdout, out.stdout + out.stderr


def test_checkpoints_saved_through_the_shim_use_reference_names(tmp_path):
    """reference modules/volumetric_model.py:83-97 pickles the render procedure, the config type and NamedTuples by qualified
    name: saved while the compat shim is active they must name ``thre3d_atom.*`` (loadable by the upstream framework), and
    load back here."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    code = r'''
import sys, pickletools, io
sys.path.insert(0, sys.argv[1] + "/compat"); sys.path.insert(0, sys.argv[1])
import torch, thre3d_atom
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid, SHVoxGridRenderConfig
from thre3d_atom.modules.volumetric_model import VolumetricModel, create_volumetric_model_from_saved_model
from thre3d_atom.utils.imaging_utils import CameraBounds
assert render_sh_voxel_grid.__module__ == "thre3d_atom.thre3d_reprs.renderers" and VoxelSize.__module__ == "thre3d_atom.thre3d_reprs.voxels"
grid = VoxelGrid(torch.rand(4, 4, 4, 1), torch.rand(4, 4, 4, 27), VoxelSize(0.5, 0.5, 0.5), tunable=True)
vm = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(8, CameraBounds(1.0, 2.0)), device=torch.device("cpu"))
path = sys.argv[2]
torch.save(vm.get_save_info({"camera_bounds": CameraBounds(1.0, 2.0)}), path)
import zipfile
raw = zipfile.ZipFile(path).read([n for n in zipfile.ZipFile(path).namelist() if n.endswith("data.pkl")][0])
names = {arg for op, arg, _ in pickletools.genops(raw) if op.name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE", "BINUNICODE") and isinstance(arg, str)}
assert not any("thr3ed_atom_b200" in n for n in names), sorted(n for n in names if "thr3ed" in n)
assert any(n.startswith("thre3d_atom.thre3d_reprs.renderers") for n in names)
back, extra = create_volumetric_model_from_saved_model(path, thre3d_repr_creator=__import__("thre3d_atom.thre3d_reprs.voxels", fromlist=["x"]).create_voxel_grid_from_saved_info_dict, device=torch.device("cpu"))
assert torch.equal(back.thre3d_repr.features, grid.features) and back.render_procedure is render_sh_voxel_grid
print("NAMES_OK")
'''
    out = subprocess.run([sys.executable, "-c", code, str(root), str(tmp_path / "ckpt.pth")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "NAMES_OK" in out.stdout, out.stdout + out.stderr


def test_reference_trainer_runs_unchanged_up_to_the_first_cuda_kernel(tmp_path):
    """The north star's "train_sh_based_voxel_grid_with_posed_images.py runs unchanged": through the compat shim the reference's
    OWN ``train_sh_vox_grid_vol_mod_with_posed_images`` (modules/trainers.py:49) is called on a B200 ``VolumetricModel`` and a
    synthetic posed-image dataset in the reference's on-disk format.  On this CPU-only box it has to get through the identity
    asserts (:116-122), the stage-size schedule, the dataset pyramid, ``scale_voxel_grid_with_required_output_size`` + re-init
    on the padded grid (:145-152), the feedback pose, data loaders, output directories, imageio and the TensorBoard writer, and
    stop exactly where the first CUDA kernel is asked to run on CPU tensors (``cast_rays``, :281-291) -- with this repo's
    deliberate "CUDA only" error, not with an interface mismatch.  Skipped where no reference checkout is present."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path

    import numpy as np
    import pytest
    from PIL import Image

    reference = os.environ.get("THRE3D_ATOM_REFERENCE", "/root/reference")
    if not (Path(reference) / "thre3d_atom" / "modules" / "trainers.py").is_file():
        pytest.skip("no reference checkout (set $THRE3D_ATOM_REFERENCE)")
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from cases import spherical_pose

    root = Path(__file__).resolve().parent.parent
    images = tmp_path / "images"
    images.mkdir()
    rng = np.random.RandomState(0)
    params = {}
    for k in range(3):  # reference data format: data/datasets.py:31-115, tools/convert_from_nerf_blender_dataset.py:64-81
        name = f"r_{k}.png"
        Image.fromarray(rng.randint(0, 255, size=(16, 16, 3), dtype=np.uint8)).save(images / name)
        rot, trans = spherical_pose(40.0 * k, 50.0, 4.0)
        params[name] = {"extrinsic": {"rotation": rot.tolist(), "translation": trans.tolist()},
                        "intrinsic": {"height": 16, "width": 16, "focal": 20.0, "bounds": [2.0, 6.0]}}
    (tmp_path / "train_camera_params.json").write_text(json.dumps(params))
    code = r'''
import sys, traceback
from pathlib import Path
sys.path.insert(0, sys.argv[1] + "/compat"); sys.path.insert(0, sys.argv[1])
import torch, thre3d_atom
from thre3d_atom.data.datasets import PosedImagesDataset                      # reference
from thre3d_atom.modules.trainers import train_sh_vox_grid_vol_mod_with_posed_images   # reference
from thre3d_atom.modules.volumetric_model import VolumetricModel               # B200
from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid, SHVoxGridRenderConfig
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
from thre3d_atom.rendering.volumetric.utils.misc import compute_expected_density_scale_for_relu_field_grid
data = Path(sys.argv[2])
ds = PosedImagesDataset(data / "images", data / "train_camera_params.json", downsample_factor=1.0, rgba_white_bkgd=True)
grid = VoxelGrid(torch.empty(16, 16, 16, 1).uniform_(-1, 1), torch.empty(16, 16, 16, 27).uniform_(-1, 1), VoxelSize(3 / 16, 3 / 16, 3 / 16),
                 density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                 expected_density_scale=compute_expected_density_scale_for_relu_field_grid((3.0, 3.0, 3.0)), tunable=True)
vm = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(16, ds.camera_bounds, white_bkgd=True), device=torch.device("cpu"))
try:
    train_sh_vox_grid_vol_mod_with_posed_images(vm, ds, output_dir=data / "out", num_stages=2, num_iterations_per_stage=1,
                                                image_batch_cache_size=2, ray_batch_size=64, num_workers=0, fast_debug_mode=True)
    print("UNEXPECTED: trained on CPU")
except RuntimeError as e:
    frames = traceback.extract_tb(e.__traceback__)
    where = [f for f in frames if "thr3ed_atom_b200" in f.filename]
    ref = [f for f in frames if "/thre3d_atom/modules/trainers.py" in f.filename]
    assert "CUDA" in str(e) and where and ref, (str(e), [f.filename for f in frames])
    assert vm.thre3d_repr.grid_dims == (8, 8, 8), vm.thre3d_repr.grid_dims          # stage-0 grid of the 2-stage schedule, rescaled by OUR voxels.py
    assert (data / "out" / "training_logs" / "rendered_output" / "1__real_log.png").is_file()
    print("STOPPED_AT", Path(where[-1].filename).name, where[-1].name, "| called from trainers.py line", ref[-1].lineno)
'''
    env = dict(os.environ, THRE3D_ATOM_REFERENCE=reference, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code, str(root), str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "STOPPED_AT" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    assert "cast_rays" in out.stdout or "_kernels.py" in out.stdout, out.stdout
