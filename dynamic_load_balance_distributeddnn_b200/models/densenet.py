"""DenseNet-BC 121/169/201/161 (CIFAR variant) with GroupNorm(32).

Architecture of reference ``Net/Densenet.py:9-100``: bare 3×3 stem, bottleneck
``GN→ReLU→1×1(→4g)→GN→ReLU→3×3(→g)``, NEW features are *prepended* (``cat([out, x])``, ``:20``),
transitions ``GN→ReLU→1×1(halve)→avg_pool2``, head ``GN→ReLU→avg_pool4→Linear``.  DenseNet-121:
6 956 298 params in 362 tensors.

Two execution paths over the same parameters:
  * ``forward`` — op-by-op through ``ops`` (fused GN+ReLU kernels, autograd, ``torch.cat``);
    this is the numerics reference.
  * dense blocks can run *concat-free* (``DenseStage.fused``; ``ops/dense_block.py``): one
    preallocated NHWC buffer per stage, each 3×3 conv writes its ``g`` channels straight into its
    slice, per-(sample, channel) GN statistics are computed once when a channel is produced and
    re-used by every later layer, and the backward accumulates into one gradient buffer in place
    (kills the 58 cat copies and their gradient adds).
"""
import math

import torch
import torch.nn as nn

from .. import ops
from .layers import Conv2d, GroupNormAct, Linear

_GROUPS = 32


class Bottleneck(nn.Module):
    def __init__(self, in_planes, growth_rate):
        super().__init__()
        self.gn1 = GroupNormAct(_GROUPS, in_planes)
        self.conv1 = Conv2d(in_planes, 4 * growth_rate, kernel_size=1, bias=False)
        self.gn2 = GroupNormAct(_GROUPS, 4 * growth_rate)
        self.conv2 = Conv2d(4 * growth_rate, growth_rate, kernel_size=3, padding=1, bias=False)

    def forward(self, x):
        out = self.conv1(self.gn1(x))
        out = self.conv2(self.gn2(out))
        return torch.cat([out, x], 1)


class DenseStage(nn.Sequential):
    """A dense block: ``nn.Sequential`` of bottlenecks (same keys as the reference) that can
    execute through the concat-free fused path."""

    fused = True

    def forward(self, x):
        if self.fused and x.is_cuda:
            try:
                from ..ops import dense_block
            except ImportError:
                dense_block = None
            if dense_block is not None and dense_block.supported(self, x):
                return dense_block.run(self, x)
        return super().forward(x)


class Transition(nn.Module):
    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.gn = GroupNormAct(_GROUPS, in_planes)
        self.conv = Conv2d(in_planes, out_planes, kernel_size=1, bias=False)

    def forward(self, x):
        # avg-pool and a 1x1 convolution are both linear and commute: pooling FIRST is the same function
        # with 4x fewer conv FLOPs / bytes (reference order: conv then pool, Net/Densenet.py:31-32)
        return self.conv(ops.avg_pool2d(self.gn(x), 2))


class DenseNet(nn.Module):
    input_shape = (3, 32, 32)

    def __init__(self, nblocks, growth_rate=12, reduction=0.5, num_classes=10):
        super().__init__()
        self.growth_rate = growth_rate
        planes = 2 * growth_rate
        self.conv1 = Conv2d(3, planes, kernel_size=3, padding=1, bias=False)
        for i, n in enumerate(nblocks):
            setattr(self, f"dense{i + 1}", self._dense(planes, n))
            planes += n * growth_rate
            if i < len(nblocks) - 1:
                out_planes = int(math.floor(planes * reduction))
                setattr(self, f"trans{i + 1}", Transition(planes, out_planes))
                planes = out_planes
        self.gn = GroupNormAct(_GROUPS, planes)
        self.linear = Linear(planes, num_classes)

    def _dense(self, in_planes, n):
        layers = []
        for _ in range(n):
            layers.append(Bottleneck(in_planes, self.growth_rate))
            in_planes += self.growth_rate
        return DenseStage(*layers)

    def forward(self, x):
        out = self.conv1(x)
        out = self.trans1(self.dense1(out))
        out = self.trans2(self.dense2(out))
        out = self.trans3(self.dense3(out))
        out = self.dense4(out)
        out = ops.avg_pool2d(self.gn(out), 4).flatten(1)
        return self.linear(out)


def DenseNet121(num_classes=10):
    return DenseNet([6, 12, 24, 16], growth_rate=32, num_classes=num_classes)


def DenseNet169(num_classes=10):
    return DenseNet([6, 12, 32, 32], growth_rate=32, num_classes=num_classes)


def DenseNet201(num_classes=10):
    return DenseNet([6, 12, 48, 32], growth_rate=32, num_classes=num_classes)


def DenseNet161(num_classes=10):
    return DenseNet([6, 12, 36, 24], growth_rate=48, num_classes=num_classes)
