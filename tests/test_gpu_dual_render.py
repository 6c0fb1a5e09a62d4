"""Single-pass specular + diffuse render: ``render_sh_voxel_grid_with_diffuse`` against the oracle run
twice (all SH bands / band 0 only, reference ``modules/trainers.py:306-330`` + ``process.py:59-63``) and against two
separate fused renders of the same sample positions."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import CASES, build_inputs, make_cuda_config, make_cuda_grid, rel_l2, run_numpy_f64

pytestmark = pytest.mark.gpu

DUAL_CASES = ["deg2_16cube", "deg2_jitter", "deg3_abs", "deg1_aniso_softplus", "c1_32cube_deg0", "deg2_sparse"]


def _rays(inp, device):
    from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays

    return Rays(torch.from_numpy(inp["origins"]).to(device), torch.from_numpy(inp["directions"]).to(device))


def _diffuse_upstream(inp):
    # a second, different upstream gradient for the diffuse image (deterministic: the specular one, flipped and rescaled)
    return np.ascontiguousarray(inp["grad_colour"][::-1, ::-1] * np.float32(0.75) + np.float32(0.01))


def _dual(case, inp, device, with_cache=True, hints=None):
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid_with_diffuse

    grid = make_cuda_grid(case, inp, device)
    hints = dict(hints or {})
    if case.jitter and "jitter" not in hints and "rng_seed" not in hints:
        hints["jitter"] = torch.from_numpy(inp["jitter"]).to(device)
    gc = torch.from_numpy(inp["grad_colour"]).to(device)
    gcd = torch.from_numpy(_diffuse_upstream(inp)).to(device)
    with render_hints(**hints):
        spec, diff = render_sh_voxel_grid_with_diffuse(grid, _rays(inp, device), make_cuda_config(case, render_diffuse=False))
        ((spec.colour * gc).sum() + (diff.colour * gcd).sum()).backward()
    nf = inp["features"].shape[-1]
    return {
        "colour": spec.colour.detach().cpu().numpy(), "colour_diffuse": diff.colour.detach().cpu().numpy(),
        "depth": spec.depth.detach().cpu().numpy(), "acc": spec.extra["accumulated_weight"].detach().cpu().numpy(),
        "grad_densities": grid.densities.grad.cpu().numpy(), "grad_features": grid.feature_storage.grad[..., :nf].cpu().numpy(),
        "grad_feature_padding": grid.feature_storage.grad[..., nf:].cpu().numpy(),
    }


@pytest.mark.parametrize("name", DUAL_CASES)
def test_dual_render_matches_two_oracle_renders(name, cuda_device):
    """Forward: each image within 1e-5 of the fp64 oracle's render of that mode; backward: the gradient of
    sum(g_s * colour) + sum(g_d * colour_diffuse) equals the sum of the oracle's two gradients (rel-L2 <= 5e-5)."""
    case = dataclasses.replace(CASES[name], diffuse=False)
    inp = build_inputs(case)
    got = _dual(case, inp, cuda_device)
    inp_d = dict(inp)
    inp_d["grad_colour"] = _diffuse_upstream(inp)
    inp_d.pop("grad_depth", None), inp_d.pop("grad_acc", None)
    inp_s = dict(inp)
    inp_s.pop("grad_depth", None), inp_s.pop("grad_acc", None)
    want_s = run_numpy_f64(case, inp_s)
    want_d = run_numpy_f64(dataclasses.replace(case, diffuse=True), inp_d)
    np.testing.assert_allclose(got["colour"], want_s["colour"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(got["colour_diffuse"], want_d["colour"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(got["acc"], want_s["acc"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(got["depth"], want_s["depth"], atol=1e-4, rtol=1e-5)
    assert rel_l2(got["grad_features"], want_s["grad_features"] + want_d["grad_features"]) < 5e-5
    assert rel_l2(got["grad_densities"], want_s["grad_densities"] + want_d["grad_densities"]) < 5e-5
    assert not got["grad_feature_padding"].any()


@pytest.mark.parametrize("name", ["deg2_16cube", "deg2_jitter", "deg3_abs"])
@pytest.mark.parametrize("with_cache", [True, False])
def test_dual_render_equals_two_fused_renders(name, with_cache, cuda_device, monkeypatch):
    """Same sample positions => the two images of the single pass equal the two separate fused renders to fp32 rounding,
    and so does the gradient; with and without the forward's per-sample caches (re-gather path)."""
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_hints, render_sh_voxel_grid

    if with_cache:
        monkeypatch.delenv("R3D_SAMPLE_CACHE_MAX_BYTES", raising=False)
    else:
        monkeypatch.setenv("R3D_SAMPLE_CACHE_MAX_BYTES", "0")
    case = dataclasses.replace(CASES[name], diffuse=False)
    inp = build_inputs(case)
    hints = {"rng_seed": 4242} if case.jitter else {}
    got = _dual(case, inp, cuda_device, hints=hints)
    grid = make_cuda_grid(case, inp, cuda_device)
    gc = torch.from_numpy(inp["grad_colour"]).to(cuda_device)
    gcd = torch.from_numpy(_diffuse_upstream(inp)).to(cuda_device)
    with render_hints(**hints):
        spec = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case, render_diffuse=False))
    with render_hints(**hints):
        diff = render_sh_voxel_grid(grid, _rays(inp, cuda_device), make_cuda_config(case, render_diffuse=True))
    ((spec.colour * gc).sum() + (diff.colour * gcd).sum()).backward()
    nf = inp["features"].shape[-1]
    np.testing.assert_allclose(got["colour"], spec.colour.detach().cpu().numpy(), atol=2e-6, rtol=0)
    np.testing.assert_allclose(got["colour_diffuse"], diff.colour.detach().cpu().numpy(), atol=2e-6, rtol=0)
    assert rel_l2(got["grad_features"], grid.feature_storage.grad[..., :nf].cpu().numpy()) < 1e-5
    assert rel_l2(got["grad_densities"], grid.densities.grad.cpu().numpy()) < 1e-5


def test_only_one_image_needs_gradient(cuda_device):
    """Upstream gradient on the diffuse image only / the specular image only (the other arrives as None)."""
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid, render_sh_voxel_grid_with_diffuse

    case = dataclasses.replace(CASES["deg2_16cube"], diffuse=False)
    inp = build_inputs(case)
    gc = torch.from_numpy(inp["grad_colour"]).to(cuda_device)
    for which in ("diffuse", "specular"):
        grid = make_cuda_grid(case, inp, cuda_device)
        spec, diff = render_sh_voxel_grid_with_diffuse(grid, _rays(inp, cuda_device), make_cuda_config(case, render_diffuse=False))
        ((diff if which == "diffuse" else spec).colour * gc).sum().backward()
        ref = make_cuda_grid(case, inp, cuda_device)
        out = render_sh_voxel_grid(ref, _rays(inp, cuda_device), make_cuda_config(case, render_diffuse=(which == "diffuse")))
        (out.colour * gc).sum().backward()
        assert rel_l2(grid.feature_storage.grad.cpu().numpy(), ref.feature_storage.grad.cpu().numpy()) < 1e-5, which
        assert rel_l2(grid.densities.grad.cpu().numpy(), ref.densities.grad.cpu().numpy()) < 1e-5, which


def test_volumetric_model_entry_point_and_errors(cuda_device):
    from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel
    from thr3ed_atom_b200.thre3d_reprs.renderers import render_sh_voxel_grid, render_sh_voxel_grid_with_diffuse

    case = dataclasses.replace(CASES["deg2_16cube"], diffuse=False)
    inp = build_inputs(case)
    grid = make_cuda_grid(case, inp, cuda_device)
    vol_mod = VolumetricModel(grid, render_sh_voxel_grid, make_cuda_config(case, render_diffuse=False), device=cuda_device)
    with torch.no_grad():
        spec, diff = vol_mod.render_rays_with_diffuse(_rays(inp, cuda_device))
        one = vol_mod.render_rays(_rays(inp, cuda_device))
        two = vol_mod.render_rays(_rays(inp, cuda_device), render_diffuse=True)
    assert spec.colour.shape == diff.colour.shape == (inp["origins"].shape[0], 3)
    assert (spec.colour - one.colour).abs().max().item() < 2e-6
    assert (diff.colour - two.colour).abs().max().item() < 2e-6
    assert torch.equal(spec.depth, diff.depth)
    with pytest.raises(ValueError, match="render_diffuse must be False"):
        render_sh_voxel_grid_with_diffuse(grid, _rays(inp, cuda_device), make_cuda_config(case, render_diffuse=True))
    with pytest.raises(ValueError, match="Unknown render configuration field"):
        vol_mod.render_rays_with_diffuse(_rays(inp, cuda_device), no_such_field=1)
