"""Import-path shim: ``sys.path.insert(0, "<repo>/compat")`` makes the reference's import paths resolve to the B200
implementation (``thr3ed_atom_b200``) for the render hot path, e.g.

    from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid, SHVoxGridRenderConfig
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.modules.volumetric_model import VolumetricModel

Hosting the reference's OWN trainer / datasets / visualisations / CLIs (which this repo deliberately does not re-implement,
SURVEY.md section 2): point ``$THRE3D_ATOM_REFERENCE`` at a checkout of akanimax/thr3ed_atom.  Every ``thre3d_atom.*`` module
that is not on the hot path (``modules.trainers``, ``modules.testers``, ``data.*``, ``visualizations.*``, ``utils.misc``,
``utils.logging`` ...) is then imported from that checkout *under this package*, so its own ``from thre3d_atom.thre3d_reprs...``
imports land on the B200 modules -- the identity checks of reference modules/trainers.py:116-122
(``isinstance(thre3d_repr, VoxelGrid)``, ``render_procedure == render_sh_voxel_grid``) hold for a B200 ``VolumetricModel`` and
``train_sh_based_voxel_grid_with_posed_images.py`` runs unchanged.  Names that a mapped module does not define
(``postprocess_depth_map``, ``ndcize_rays``: visualisation helpers) are taken from the checkout's module of the same name.
Third-party packages the reference imports but that may be absent offline (``easydict``, ``imageio``, ``matplotlib``,
``lpips``) get minimal stand-ins from ``compat/standins`` -- only when the real package cannot be imported.
"""
import importlib
import importlib.abc
import importlib.util
import os
import sys
import types
from pathlib import Path

_MAPPED = [
    "utils", "utils.constants", "utils.imaging_utils", "utils.metric_utils",
    "rendering", "rendering.volumetric", "rendering.volumetric.render_interface", "rendering.volumetric.accumulate",
    "rendering.volumetric.utils", "rendering.volumetric.utils.misc",
    "thre3d_reprs", "thre3d_reprs.constants", "thre3d_reprs.voxels", "thre3d_reprs.renderers",
    "modules", "modules.volumetric_model",
]
_PKG = __name__
_REFERENCE = os.environ.get("THRE3D_ATOM_REFERENCE")
_REF_ROOT = Path(_REFERENCE) / "thre3d_atom" if _REFERENCE else None
if _REF_ROOT is not None and not _REF_ROOT.is_dir():
    raise ImportError(f"$THRE3D_ATOM_REFERENCE={_REFERENCE!r} does not contain a thre3d_atom package")


def _ensure_standins() -> None:
    """easydict / imageio / matplotlib / lpips: real package if importable, else the stand-in directory (appended, so an
    installed package always wins)."""
    standins = str(Path(__file__).resolve().parent.parent / "standins")
    for name in ("easydict", "imageio", "matplotlib", "lpips"):
        if importlib.util.find_spec(name) is None and standins not in sys.path:
            sys.path.append(standins)


def _reference_file(relative: str):
    """(path, is_package) of module ``relative`` ("a.b.c") inside the reference checkout, or None."""
    if _REF_ROOT is None:
        return None
    base = _REF_ROOT.joinpath(*relative.split("."))
    if (base / "__init__.py").is_file():
        return base / "__init__.py", True
    if base.with_suffix(".py").is_file():
        return base.with_suffix(".py"), False
    return None


def _load_reference(relative: str, as_name: str):
    found = _reference_file(relative)
    if found is None:
        return None
    path, is_pkg = found
    spec = importlib.util.spec_from_file_location(as_name, path, submodule_search_locations=[str(path.parent)] if is_pkg else None)
    module = importlib.util.module_from_spec(spec)
    sys.modules[as_name] = module
    try:
        spec.loader.exec_module(module)
    except BaseException:
        sys.modules.pop(as_name, None)
        raise
    return module


class _MappedModule(types.ModuleType):
    """``thre3d_atom.x`` backed by ``thr3ed_atom_b200.x``; names the B200 module lacks come from the reference checkout."""

    def __init__(self, name: str, target: types.ModuleType, relative: str):
        super().__init__(name, target.__doc__)
        self.__dict__["_target"], self.__dict__["_relative"] = target, relative
        self.__dict__["__path__"] = list(getattr(target, "__path__", [])) if hasattr(target, "__path__") else None
        if self.__dict__["__path__"] is None:
            del self.__dict__["__path__"]
        self.__dict__["__file__"] = getattr(target, "__file__", None)

    def __getattr__(self, item):
        target = self.__dict__["_target"]
        try:
            return getattr(target, item)
        except AttributeError:
            pass
        if item.startswith("__"):
            raise AttributeError(item)
        private = f"{_PKG}._reference.{self.__dict__['_relative']}"
        fallback = sys.modules.get(private) or _load_reference(self.__dict__["_relative"], private)
        if fallback is not None and hasattr(fallback, item):
            return getattr(fallback, item)
        raise AttributeError(f"module {self.__name__!r} has no attribute {item!r}" + ("" if _REF_ROOT else " (set $THRE3D_ATOM_REFERENCE to delegate to a reference checkout)"))

    def __dir__(self):
        return sorted(set(dir(self.__dict__["_target"])))


class _ReferenceFinder(importlib.abc.MetaPathFinder):
    """Unmapped ``thre3d_atom.*`` modules: import them from the reference checkout under this package's name."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PKG + ".") or fullname.startswith(_PKG + "._reference"):
            return None
        relative = fullname[len(_PKG) + 1:]
        if relative in _MAPPED:
            return None
        found = _reference_file(relative)
        if found is None:
            return None
        file, is_pkg = found
        return importlib.util.spec_from_file_location(fullname, file, submodule_search_locations=[str(file.parent)] if is_pkg else None)


def _adopt_reference_names(target: types.ModuleType, shim_name: str) -> None:
    """Classes and functions DEFINED in a mapped module report the reference's module path while the shim is active, so that
    what ``VolumetricModel.get_save_info`` pickles by qualified name (``render_sh_voxel_grid``, ``SHVoxGridRenderConfig``,
    ``density2occupancy_pb``, the camera / voxel NamedTuples; reference modules/volumetric_model.py:83-97) is written under
    ``thre3d_atom.*`` -- checkpoints saved through the shim stay loadable by the upstream framework."""
    for attr, obj in list(vars(target).items()):
        if attr.startswith("_") or not (isinstance(obj, type) or isinstance(obj, types.FunctionType)):
            continue
        if getattr(obj, "__module__", None) == target.__name__:
            try:
                obj.__module__ = shim_name
            except (AttributeError, TypeError):
                pass


_ensure_standins()
for _name in _MAPPED:
    _target = importlib.import_module(f"thr3ed_atom_b200.{_name}")
    _adopt_reference_names(_target, f"{_PKG}.{_name}")
    sys.modules[f"{_PKG}.{_name}"] = _MappedModule(f"{_PKG}.{_name}", _target, _name) if _REF_ROOT is not None else _target
for _top in ("utils", "rendering", "thre3d_reprs", "modules"):
    setattr(sys.modules[_PKG], _top, sys.modules[f"{_PKG}.{_top}"])
if _REF_ROOT is not None:
    sys.modules[f"{_PKG}._reference"] = types.ModuleType(f"{_PKG}._reference")
    sys.meta_path.insert(0, _ReferenceFinder())
