"""PSNR from a mean squared error, for floats and for 0-d / 1-element tensors alike (the helper the callers of the
render path log with; reference thre3d_atom/utils/metric_utils.py:10-21).  A zero error maps to the reference's
"infinite" PSNR sentinels: ``math.inf`` for floats, a ``[1e10]`` tensor for tensors."""
import math
from typing import Union

import torch

from thr3ed_atom_b200.utils.constants import INFINITY

_DB_PER_NEPER = 10.0 / math.log(10.0)


def mse2psnr(x: Union[float, torch.Tensor]) -> Union[float, torch.Tensor]:
    is_tensor = isinstance(x, torch.Tensor)
    if x == 0.0:
        return torch.tensor([INFINITY], dtype=x.dtype, device=x.device) if is_tensor else math.inf
    return -_DB_PER_NEPER * (torch.log(x) if is_tensor else math.log(x))
