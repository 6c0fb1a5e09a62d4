"""Import-path shim: ``sys.path.insert(0, "<repo>/compat")`` makes the reference's import paths for the render hot path
resolve to the B200 implementation (``thr3ed_atom_b200``), e.g.

    from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid, SHVoxGridRenderConfig
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.modules.volumetric_model import VolumetricModel

Only the modules on the path are mapped (INTEGRATION.md lists them); the reference's trainer, datasets, visualisations and
CLIs are not part of this repo.  Do not put this directory on the path together with a checkout of the reference itself.
"""
import importlib
import sys

_MAPPED = [
    "utils", "utils.constants", "utils.imaging_utils", "utils.metric_utils",
    "rendering", "rendering.volumetric", "rendering.volumetric.render_interface", "rendering.volumetric.accumulate",
    "rendering.volumetric.utils", "rendering.volumetric.utils.misc",
    "thre3d_reprs", "thre3d_reprs.constants", "thre3d_reprs.voxels", "thre3d_reprs.renderers",
    "modules", "modules.volumetric_model",
]
for _name in _MAPPED:
    sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"thr3ed_atom_b200.{_name}")
for _top in ("utils", "rendering", "thre3d_reprs", "modules"):
    setattr(sys.modules[__name__], _top, sys.modules[f"{__name__}.{_top}"])
