"""Stand-in for ``lpips`` (reference modules/testers.py:26,45-49): the VGG weights cannot be fetched offline, so the metric is
reported as NaN (with one warning) instead of stopping the reference's evaluation loop.  Only used when the real package is
not installed."""
import warnings

import torch


class LPIPS(torch.nn.Module):
    def __init__(self, net: str = "vgg", **kwargs):
        super().__init__()
        self.net = net
        self._warned = False

    def forward(self, in0, in1, normalize: bool = False, **kwargs):
        if not self._warned:
            warnings.warn("lpips is not installed: LPIPS values are reported as NaN (compat/standins/lpips.py)")
            self._warned = True
        return torch.full((in0.shape[0], 1, 1, 1), float("nan"), device=in0.device)
