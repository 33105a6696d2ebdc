"""GoogLeNet (Inception v1, CIFAR variant) with GroupNorm.

Capabilities of reference ``Net/GoogleNet.py:7-98`` — same widths, GN group counts (8 / 16), biased
convolutions, 258 parameter tensors / 6 166 250 parameters — with the reference's defect fixed: its
5×5 branch applies ``GroupNorm(8, n5x5red)`` *before* the conv that produces those channels
(``Net/GoogleNet.py:28-31``) and crashes at the first forward.  Here every branch is
the intended conv→GN→ReLU.
"""
import torch
import torch.nn as nn

from .layers import Conv2d, GroupNormAct, Linear, MaxPool2d
from .. import ops


def _unit(cin, cout, k, groups):
    """conv(k×k, bias) → GN(groups) → ReLU (fused)."""
    return [Conv2d(cin, cout, kernel_size=k, padding=k // 2), GroupNormAct(groups, cout)]


class Inception(nn.Module):
    def __init__(self, in_planes, n1x1, n3x3red, n3x3, n5x5red, n5x5, pool_planes):
        super().__init__()
        self.b1 = nn.Sequential(*_unit(in_planes, n1x1, 1, 8))
        self.b2 = nn.Sequential(*_unit(in_planes, n3x3red, 1, 8), *_unit(n3x3red, n3x3, 3, 16))
        self.b3 = nn.Sequential(*_unit(in_planes, n5x5red, 1, 8), *_unit(n5x5red, n5x5, 3, 8),
                                *_unit(n5x5, n5x5, 3, 8))
        self.b4 = nn.Sequential(MaxPool2d(3, stride=1, padding=1), *_unit(in_planes, pool_planes, 1, 8))

    def forward(self, x):
        return torch.cat([self.b1(x), self.b2(x), self.b3(x), self.b4(x)], 1)


class GoogLeNet(nn.Module):
    input_shape = (3, 32, 32)

    def __init__(self, num_classes=10):
        super().__init__()
        self.pre_layers = nn.Sequential(*_unit(3, 192, 3, 8))
        self.a3 = Inception(192, 64, 96, 128, 16, 32, 32)
        self.b3 = Inception(256, 128, 128, 192, 32, 96, 64)
        self.maxpool = MaxPool2d(3, stride=2, padding=1)
        self.a4 = Inception(480, 192, 96, 208, 16, 48, 64)
        self.b4 = Inception(512, 160, 112, 224, 24, 64, 64)
        self.c4 = Inception(512, 128, 128, 256, 24, 64, 64)
        self.d4 = Inception(512, 112, 144, 288, 32, 64, 64)
        self.e4 = Inception(528, 256, 160, 320, 32, 128, 128)
        self.a5 = Inception(832, 256, 160, 320, 32, 128, 128)
        self.b5 = Inception(832, 384, 192, 384, 48, 128, 128)
        self.avgpool = nn.AvgPool2d(8, stride=1)
        self.linear = Linear(1024, num_classes)

    def forward(self, x):
        out = self.b3(self.a3(self.pre_layers(x)))
        out = self.maxpool(out)
        out = self.e4(self.d4(self.c4(self.b4(self.a4(out)))))
        out = self.maxpool(out)
        out = self.b5(self.a5(out))
        return self.linear(ops.avg_pool2d(out, 8).flatten(1))     # AvgPool2d(8) on the 8x8 map (self.avgpool kept for parity)
