"""Keys of the checkpoint dictionary (``VolumetricModel.get_save_info``) and of a ``VoxelGrid`` state dict, under the
reference's names (reference thre3d_atom/thre3d_reprs/constants.py:1-11).  This is an on-disk format: do not change."""
_CHECKPOINT_LAYOUT = {
    # top level of the saved dict
    "THRE3D_REPR": "thre3d_repr",  # -> {STATE_DICT: ..., CONFIG_DICT: ...}
    "RENDER_PROCEDURE": "render_procedure",  # the render function object (pickled by qualified name)
    "RENDER_CONFIG_TYPE": "render_config_type",  # the config dataclass type
    "RENDER_CONFIG": "render_config",  # dataclasses.asdict(config)
    # inside THRE3D_REPR
    "STATE_DICT": "state_dict",
    "CONFIG_DICT": "config_dict",
    # state-dict entries of a VoxelGrid
    "u_DENSITIES": "_densities",
    "u_FEATURES": "_features",
}
globals().update(_CHECKPOINT_LAYOUT)
__all__ = list(_CHECKPOINT_LAYOUT)
