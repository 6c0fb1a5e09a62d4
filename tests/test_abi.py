"""The C-ABI boundary without a GPU: the library loads, exports exactly the symbols include/r3d_b200.h
declares, the ctypes mirror has the C compiler's struct layouts, and argument validation (which runs
before any CUDA call) reports errors through r3d_last_error.  No compute call is made here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "r3d_b200.h"


@pytest.fixture(scope="module")
def lib():
    from thr3ed_atom_b200 import _abi, build

    build.build()  # no-op when current; nvcc cross-compiles without a GPU
    return _abi.lib()


def declared_functions():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"R3D_API\s+[\w\s\*]+?\b(r3d_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    assert names == sorted(
        ["r3d_abi_version", "r3d_last_error", "r3d_sample_mask_words", "r3d_render_fwd", "r3d_render_bwd", "r3d_cast_rays", "r3d_grid_lookup_fwd",
         "r3d_grid_lookup_bwd", "r3d_mark_touched_voxels", "r3d_adam_step", "r3d_multimem_all_reduce", "r3d_sample_statistics",
         "r3d_density_quad_floats", "r3d_build_density_quads", "r3d_multimem_shard_floats", "r3d_multimem_adam_step", "r3d_has_ab_variants", "r3d_sample_ray_batch", "r3d_peer_adam_step"]
    )


def test_library_exports_every_declared_symbol_and_nothing_else(lib):
    from thr3ed_atom_b200 import _abi

    out = subprocess.run(["nm", "-D", "--defined-only", str(_abi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert [s for s in exported if s.startswith("r3d_")] == declared_functions()
    assert not [s for s in exported if not s.startswith("r3d_") and not s.startswith("_")], exported
    assert sorted(_abi.SIGNATURES) == declared_functions()
    for name in declared_functions():
        assert getattr(lib, name) is not None
    assert lib.r3d_abi_version() == _abi.ABI_VERSION


def test_ctypes_structs_match_the_c_layout(tmp_path):
    from thr3ed_atom_b200 import _abi

    structs = ["R3dGrid", "R3dCamera", "R3dViewSet", "R3dRays", "R3dRenderConfig", "R3dRenderOut", "R3dRenderOutGrad", "R3dGridGrad"]
    fields = {n: [f[0] for f in getattr(_abi, n)._fields_] for n in structs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for n in structs:
        lines.append(f'printf("{n} %zu\\n", sizeof({n}));')
        for fld in fields[n]:
            lines.append(f'printf("{n}.{fld} %zu\\n", offsetof({n}, {fld}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)  # the header is plain C
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in structs:
        cls = getattr(_abi, n)
        assert int(got[n]) == C.sizeof(cls), n
        for fld in fields[n]:
            assert int(got[f"{n}.{fld}"]) == getattr(cls, fld).offset, f"{n}.{fld}"


def test_argument_validation_reports_through_last_error(lib):
    from thr3ed_atom_b200 import _abi

    # NULL grid: rejected before any CUDA call
    assert lib.r3d_render_fwd(None, None, None, None, None) == 1
    assert b"grid is NULL" in lib.r3d_last_error()
    g = _abi.R3dGrid()
    g.densities, g.features = 16, 16  # non-NULL dummies; validation fails before they are touched
    g.dims[:] = [4, 4, 4]
    g.aabb_min[:] = [-1, -1, -1]
    g.aabb_max[:] = [1, 1, 1]
    g.sh_degree, g.num_features, g.feature_stride = 4, 75, 76
    assert lib.r3d_render_fwd(C.byref(g), None, None, None, None) == 2
    assert b"only degrees 0, 1, 2, and 3 are supported" in lib.r3d_last_error()
    g.sh_degree, g.num_features, g.feature_stride = 2, 26, 28
    assert lib.r3d_render_fwd(C.byref(g), None, None, None, None) == 1
    assert b"does not match" in lib.r3d_last_error()
    g.num_features = 27
    assert lib.r3d_render_fwd(C.byref(g), None, None, None, None) == 1
    assert b"rays is NULL" in lib.r3d_last_error()
    r = _abi.R3dRays()
    r.num_rays = 5
    assert lib.r3d_render_fwd(C.byref(g), C.byref(r), None, None, None) == 1
    assert b"no camera was given" in lib.r3d_last_error()
    r.origins, r.directions = 16, 16
    c = _abi.R3dRenderConfig()
    c.num_samples = 0
    assert lib.r3d_render_fwd(C.byref(g), C.byref(r), C.byref(c), None, None) == 1
    assert b"num_samples_per_ray must be >= 1" in lib.r3d_last_error()
    with pytest.raises(RuntimeError, match="num_samples_per_ray"):
        _abi.check(1, "r3d_render_fwd")
    assert lib.r3d_adam_step(None, None, None, None, 8, 0.1, 0.9, 0.999, 1e-8, 0.1, 0.001, 1.0, None) == 1
    # fused exchange + optimizer entry points: every argument error is caught before a kernel is enqueued
    hyper = (0.1, 0.9, 0.999, 1e-8, 0.1, 0.001, 1.0)
    assert lib.r3d_multimem_adam_step(None, None, None, None, None, 8, 0, 2, *hyper, 0, None) == 1
    assert b"multicast pointer is NULL" in lib.r3d_last_error()
    assert lib.r3d_peer_adam_step(None, None, None, None, 8, 0, 2, *hyper, 0, None) == 1
    assert b"NULL argument" in lib.r3d_last_error()
    two = (C.c_void_p * 2)(32, 0)  # replica 1 missing
    assert lib.r3d_peer_adam_step(two, two, 32, 32, 8, 0, 2, *hyper, 0, None) == 1
    assert b"replica 1" in lib.r3d_last_error()
    ok2 = (C.c_void_p * 2)(32, 64)
    assert lib.r3d_peer_adam_step(ok2, ok2, 32, 32, 8, 2, 2, *hyper, 0, None) == 1  # rank out of range
    assert lib.r3d_peer_adam_step(ok2, ok2, 32, 32, 8, 0, 9, *hyper, 0, None) == 1  # more replicas than the kernel addresses
    assert b"1..8 ranks" in lib.r3d_last_error()
    assert lib.r3d_peer_adam_step(ok2, ok2, 32, 32, 6, 0, 2, *hyper, 0, None) == 1  # not a multiple of 4 floats
    assert lib.r3d_multimem_shard_floats(16, 3) == 8 and lib.r3d_multimem_shard_floats(6, 2) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from thr3ed_atom_b200 import _abi

    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.lib()
