"""Keys of the checkpoint dictionary (``VolumetricModel.get_save_info``) and of a ``VoxelGrid`` state dict, under the
reference's names (reference thre3d_atom/thre3d_reprs/constants.py:1-11).  This is an on-disk format: do not change."""
from typing import Final

# top level of the saved dict
THRE3D_REPR: Final[str] = "thre3d_repr"  # -> {STATE_DICT: ..., CONFIG_DICT: ...}
RENDER_PROCEDURE: Final[str] = "render_procedure"  # the render function object (pickled by qualified name)
RENDER_CONFIG_TYPE: Final[str] = "render_config_type"  # the config dataclass type
RENDER_CONFIG: Final[str] = "render_config"  # dataclasses.asdict(config)

# inside THRE3D_REPR
STATE_DICT: Final[str] = "state_dict"
CONFIG_DICT: Final[str] = "config_dict"

# state-dict entries of a VoxelGrid
u_DENSITIES: Final[str] = "_densities"
u_FEATURES: Final[str] = "_features"
