"""PSNR helper used by the callers of the render path (reference thre3d_atom/utils/metric_utils.py:10-21)."""
import math
from typing import Any

import torch
from torch import Tensor

from thr3ed_atom_b200.utils.constants import INFINITY


def mse2psnr(x: Any) -> Any:
    if isinstance(x, Tensor):
        if x == 0.0:
            return torch.tensor([INFINITY], dtype=x.dtype, device=x.device)
        return -10.0 * torch.log(x) / math.log(10.0)
    return -10.0 * math.log(x) / math.log(10.0) if x != 0.0 else math.inf
