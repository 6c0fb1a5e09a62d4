"""thr3ed_atom_b200 -- B200-native (sm_100a) implementation of the volumetric-rendering hot path of
akanimax/thr3ed_atom: ray generation -> stratified sampling -> trilinear lookup in a dense
(density + SH) voxel grid -> ReLU -> SH evaluation -> alpha compositing, and the backward pass into
the grid.  Host side: Python/PyTorch mirroring the reference's module layout
(``thre3d_reprs.voxels.VoxelGrid``, ``thre3d_reprs.renderers.render_sh_voxel_grid`` ...); device side:
hand-written CUDA behind the C ABI in ``include/r3d_b200.h`` (``_lib/libr3d_b200.so``).

There is no CPU / eager fallback: rendering requires the built library and CUDA tensors.
"""
__version__ = "0.1.0"

from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays, RenderOut  # noqa: F401
from thr3ed_atom_b200.thre3d_reprs.renderers import (  # noqa: F401
    SHVoxGridRenderConfig,
    render_hints,
    render_sh_voxel_grid,
)
from thr3ed_atom_b200.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize  # noqa: F401
from thr3ed_atom_b200.modules.volumetric_model import VolumetricModel  # noqa: F401
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds, CameraIntrinsics, CameraPose  # noqa: F401
