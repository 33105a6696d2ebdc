"""RegNetX-200MF / X-400MF / Y-400MF (CIFAR variant) with GroupNorm(32) and squeeze-excitation.

Architecture of reference ``Net/RegNet.py:10-141``: stem 3×3→64 + GN, four stages of bottleneck
blocks ``1×1 → grouped 3×3 (group width 8/16) → [SE] → 1×1`` with projection shortcuts, global
average pool, linear.  SE squeeze width derives from the block *input* width (``:41``).
RegNetY-400MF: 5 714 362 params / 303 tensors.
"""
import torch.nn as nn

from .layers import Conv2d, GroupNormAct, Linear
from .. import ops

_GROUPS = 32


def _groups(c: int) -> int:
    """GroupNorm(32, c) like the reference; for widths not divisible by 32 (RegNetX-200MF's
    24/56/152/368, which make the reference constructor raise) the largest divisor ≤ 32."""
    if c % _GROUPS == 0:
        return _GROUPS
    return max(g for g in range(1, _GROUPS + 1) if c % g == 0)


class SE(nn.Module):
    def __init__(self, in_planes, se_planes):
        super().__init__()
        self.se1 = Conv2d(in_planes, se_planes, kernel_size=1, bias=True)
        self.se2 = Conv2d(se_planes, in_planes, kernel_size=1, bias=True)

    def forward(self, x):
        # squeeze, both 1x1 projections, the gate and the scaling: one fused kernel per direction on CUDA (ops/se.py)
        from ..ops import se as fused_se
        return fused_se.squeeze_excite(x, self.se1.weight, self.se1.bias, self.se2.weight, self.se2.bias)


class Block(nn.Module):
    def __init__(self, w_in, w_out, stride, group_width, bottleneck_ratio, se_ratio):
        super().__init__()
        w_b = int(round(w_out * bottleneck_ratio))
        self.conv1 = Conv2d(w_in, w_b, kernel_size=1, bias=False)
        self.gn1 = GroupNormAct(_groups(w_b), w_b)
        self.conv2 = Conv2d(w_b, w_b, kernel_size=3, stride=stride, padding=1, groups=w_b // group_width, bias=False)
        self.gn2 = GroupNormAct(_groups(w_b), w_b)
        self.with_se = se_ratio > 0
        if self.with_se:
            self.se = SE(w_b, int(round(w_in * se_ratio)))
        self.conv3 = Conv2d(w_b, w_out, kernel_size=1, bias=False)
        self.gn3 = GroupNormAct(_groups(w_out), w_out, relu=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or w_in != w_out:
            self.shortcut = nn.Sequential(Conv2d(w_in, w_out, kernel_size=1, stride=stride, bias=False),
                                          GroupNormAct(_groups(w_out), w_out, relu=False))

    def forward(self, x):
        out = self.gn1(self.conv1(x))
        out = self.gn2(self.conv2(out))
        if self.with_se:
            out = self.se(out)
        return self.gn3(self.conv3(out), residual=self.shortcut(x))


class RegNet(nn.Module):
    input_shape = (3, 32, 32)

    def __init__(self, cfg, num_classes=10):
        super().__init__()
        self.cfg = cfg
        self.in_planes = 64
        self.conv1 = Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.gn1 = GroupNormAct(_GROUPS, 64)
        for i in range(4):
            setattr(self, f"layer{i + 1}", self._stage(i))
        self.linear = Linear(cfg["widths"][-1], num_classes)

    def _stage(self, idx):
        c = self.cfg
        blocks = []
        for i in range(c["depths"][idx]):
            blocks.append(Block(self.in_planes, c["widths"][idx], c["strides"][idx] if i == 0 else 1,
                                c["group_width"], c["bottleneck_ratio"], c["se_ratio"]))
            self.in_planes = c["widths"][idx]
        return nn.Sequential(*blocks)

    def forward(self, x):
        out = self.gn1(self.conv1(x))
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        return self.linear(ops.global_avg_pool2d(out).flatten(1))


def _cfg(depths, widths, group_width, se_ratio):
    return {"depths": depths, "widths": widths, "strides": [1, 1, 2, 2], "group_width": group_width,
            "bottleneck_ratio": 1, "se_ratio": se_ratio}


def RegNetX_200MF(num_classes=10):
    return RegNet(_cfg([1, 1, 4, 7], [24, 56, 152, 368], 8, 0), num_classes)


def RegNetX_400MF(num_classes=10):
    return RegNet(_cfg([1, 2, 7, 12], [32, 64, 160, 384], 16, 0), num_classes)


def RegNetY_400MF(num_classes=10):
    return RegNet(_cfg([1, 2, 7, 12], [32, 64, 160, 384], 16, 0.25), num_classes)
