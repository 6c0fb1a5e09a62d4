"""Shared test plumbing: load a golden case, run the oracles on it, compare results."""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(GOLDEN)):
    if p not in sys.path:
        sys.path.insert(0, p)

from cases import CASES, Case, build_inputs  # noqa: E402  (tests/golden/cases.py)


def load_golden(name: str) -> Dict[str, np.ndarray]:
    with np.load(GOLDEN / f"{name}.npz") as f:
        return {k: f[k] for k in f.files}


def aabb_of(case: Case):
    out = []
    for n, s, c in zip(case.dims, case.voxel_size, case.location):
        half = (n * s) / 2
        out.append((c - half, c + half))
    return tuple(out)


def run_torch_port(case: Case, inp: Dict[str, np.ndarray], with_grads: bool = True):
    from oracle import torch_port as tp

    grid = tp.OracleGrid(
        densities=torch.from_numpy(inp["densities"]),
        features=torch.from_numpy(inp["features"]),
        voxel_size=case.voxel_size,
        location=case.location,
        density_scale=case.density_scale,
        density_pre=case.density_pre,
        density_post=case.density_post,
    )
    cfg = dict(
        num_samples=case.num_samples,
        near=case.near,
        far=case.far,
        jitter=torch.from_numpy(inp["jitter"]) if case.jitter else None,
        white_bkgd=case.white_bkgd,
        diffuse=case.diffuse,
        optimized_sampling=case.optimized_sampling,
    )
    o, d = torch.from_numpy(inp["origins"]), torch.from_numpy(inp["directions"])
    if not with_grads:
        with torch.no_grad():
            out = tp.render(grid, o, d, **cfg)
    else:
        out = tp.render_with_grads(
            grid,
            o,
            d,
            torch.from_numpy(inp["grad_colour"]),
            torch.from_numpy(inp["grad_depth"]) if "grad_depth" in inp else None,
            torch.from_numpy(inp["grad_acc"]) if "grad_acc" in inp else None,
            **cfg,
        )
    return {k: v.numpy() for k, v in out.items()}


def run_numpy_f64(case: Case, inp: Dict[str, np.ndarray], with_grads: bool = True):
    from oracle import numpy_f64 as nf

    return nf.render(
        inp["densities"],
        inp["features"],
        aabb_of(case),
        inp["origins"],
        inp["directions"],
        num_samples=case.num_samples,
        near=case.near,
        far=case.far,
        density_scale=case.density_scale,
        density_pre=case.density_pre,
        density_post=case.density_post,
        jitter=inp.get("jitter"),
        white_bkgd=case.white_bkgd,
        diffuse=case.diffuse,
        optimized_sampling=case.optimized_sampling,
        grad_colour=inp["grad_colour"] if with_grads else None,
        grad_depth=inp.get("grad_depth"),
        grad_acc=inp.get("grad_acc"),
    )


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def assert_outputs_close(got, want, *, atol=1e-5, rtol_depth=1e-5, what=""):
    """Forward tolerances of SURVEY.md 8(c): colour/acc abs <= 1e-5, depth rel <= 1e-5 (+ abs floor),
    disparity compared where finite, NaNs must coincide."""
    np.testing.assert_allclose(got["colour"], want["colour"], atol=atol, rtol=0, err_msg=f"{what} colour")
    np.testing.assert_allclose(got["acc"], want["acc"], atol=atol, rtol=0, err_msg=f"{what} acc")
    np.testing.assert_allclose(got["depth"], want["depth"], atol=atol * 10, rtol=rtol_depth, err_msg=f"{what} depth")
    gn, wn = np.isnan(got["disparity"]), np.isnan(want["disparity"])
    assert np.array_equal(gn, wn), f"{what}: NaN pattern of disparity differs ({gn.sum()} vs {wn.sum()})"
    # disparity = acc / depth amplifies relative error of tiny-weight rays; compare where acc is not negligible
    ok = (~wn) & (np.abs(want["acc"]) > 1e-3)
    np.testing.assert_allclose(got["disparity"][ok], want["disparity"][ok], rtol=1e-3, atol=1e-5, err_msg=f"{what} disparity")
