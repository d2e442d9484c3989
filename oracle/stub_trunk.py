"""TEST INFRASTRUCTURE -- a stand-in for the third-party EfficientNet trunk (efficientnet-pytorch 0.7.0 is not vendored
by the reference and absent from the image): the attribute surface Encoder.get_features_depth walks
(stp3/models/encoder.py:57-86: _conv_stem, _bn0, _swish, _blocks[i](x, drop_connect_rate=...), _global_params) with
cheap deterministic blocks that reproduce EfficientNet-b4's endpoint geometry: reductions of 24 / 32 / 56 / 160
channels at 1/2, 1/4, 1/8, 1/16 of the input, the 1/16 one after block 21.  Lets the reference's own
get_features_depth run end to end (endpoint bookkeeping + both heads) so the drop-in Encoder can be checked against it."""
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=True)
        self.stride = stride

    def forward(self, x, drop_connect_rate=None):
        if self.stride == 2:
            x = F.avg_pool2d(x, 2)
        return torch.tanh(self.conv(x))


class StubEfficientNetB4(nn.Module):
    """32 blocks like efficientnet-b4 (the encoders delete those after 21); strides at blocks 2, 6, 10, 22."""

    def __init__(self):
        super().__init__()
        self._conv_stem = nn.Conv2d(3, 48, 3, stride=2, padding=1, bias=False)
        self._bn0 = nn.BatchNorm2d(48)
        self._swish = nn.SiLU()
        chans = [48] + [24] * 2 + [32] * 4 + [56] * 4 + [112] * 6 + [160] * 6 + [272] * 8 + [448] * 2
        stride_at = {2, 6, 10, 22}
        self._blocks = nn.ModuleList(_Block(chans[i], chans[i + 1], 2 if i in stride_at else 1) for i in range(32))
        self._global_params = SimpleNamespace(drop_connect_rate=0.2)
        # attributes the encoders delete
        self._conv_head, self._bn1 = nn.Identity(), nn.Identity()
        self._avg_pooling, self._dropout, self._fc = nn.Identity(), nn.Identity(), nn.Identity()
