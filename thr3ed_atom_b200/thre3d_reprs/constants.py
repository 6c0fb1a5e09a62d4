"""Keys of the checkpoint dictionary written by ``VolumetricModel.get_save_info``
(reference thre3d_atom/thre3d_reprs/constants.py:1-11; on-disk format, must not change)."""
THRE3D_REPR = "thre3d_repr"
RENDER_PROCEDURE = "render_procedure"
RENDER_CONFIG = "render_config"
RENDER_CONFIG_TYPE = "render_config_type"
STATE_DICT = "state_dict"
CONFIG_DICT = "config_dict"

# state_dict keys of a VoxelGrid
u_DENSITIES = "_densities"
u_FEATURES = "_features"
