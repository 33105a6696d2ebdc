"""CIFAR-stem ResNet-18/34/50/101/152 with GroupNorm(32) everywhere.

Architecture of reference ``Net/Resnet.py:5-108``: 3×3 stride-1 stem without max-pool, basic /
bottleneck blocks, projection shortcut = 1×1 conv + GN, ``avg_pool2d(4)`` head.  Every
``GN → (+shortcut) → ReLU`` is a single fused op here.  Attribute names match the
reference so checkpoints interchange (ResNet-50: 23 520 842 params / 161 tensors; ResNet-101:
42 512 970 / 314).
"""
import torch.nn as nn

from .layers import Conv2d, GroupNormAct, Linear
from .. import ops

_GROUPS = 32


class _Shortcut(nn.Sequential):
    """index 0 = 1×1 conv, index 1 = GN (no ReLU): same state_dict keys as the reference."""

    def __init__(self, cin, cout, stride):
        super().__init__(Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False),
                         GroupNormAct(_GROUPS, cout, relu=False))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.gn1 = GroupNormAct(_GROUPS, planes, relu=True)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.gn2 = GroupNormAct(_GROUPS, planes, relu=True)          # relu applied after the residual add
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes * self.expansion:
            self.shortcut = _Shortcut(in_planes, planes * self.expansion, stride)

    def forward(self, x):
        out = self.gn1(self.conv1(x))
        return self.gn2(self.conv2(out), residual=self.shortcut(x))


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2d(in_planes, planes, kernel_size=1, bias=False)
        self.gn1 = GroupNormAct(_GROUPS, planes)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.gn2 = GroupNormAct(_GROUPS, planes)
        self.conv3 = Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.gn3 = GroupNormAct(_GROUPS, planes * self.expansion, relu=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes * self.expansion:
            self.shortcut = _Shortcut(in_planes, planes * self.expansion, stride)

    def forward(self, x):
        out = self.gn1(self.conv1(x))
        out = self.gn2(self.conv2(out))
        return self.gn3(self.conv3(out), residual=self.shortcut(x))


class ResNet(nn.Module):
    input_shape = (3, 32, 32)

    def __init__(self, block, num_blocks, num_classes=10):
        super().__init__()
        self.in_planes = 64
        self.conv1 = Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.gn1 = GroupNormAct(_GROUPS, 64)
        self.layer1 = self._stage(block, 64, num_blocks[0], 1)
        self.layer2 = self._stage(block, 128, num_blocks[1], 2)
        self.layer3 = self._stage(block, 256, num_blocks[2], 2)
        self.layer4 = self._stage(block, 512, num_blocks[3], 2)
        self.linear = Linear(512 * block.expansion, num_classes)

    def _stage(self, block, planes, n, stride):
        blocks = []
        for s in [stride] + [1] * (n - 1):
            blocks.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*blocks)

    def forward(self, x):
        out = self.gn1(self.conv1(x))
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        out = ops.avg_pool2d(out, 4).flatten(1)
        return self.linear(out)


def ResNet18(num_classes=10):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes)


def ResNet34(num_classes=10):
    return ResNet(BasicBlock, [3, 4, 6, 3], num_classes)


def ResNet50(num_classes=10):
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes)


def ResNet101(num_classes=10):
    return ResNet(Bottleneck, [3, 4, 23, 3], num_classes)


def ResNet152(num_classes=10):
    return ResNet(Bottleneck, [3, 8, 36, 3], num_classes)
