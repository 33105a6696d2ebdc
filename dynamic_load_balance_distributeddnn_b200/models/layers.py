"""Building blocks shared by the model zoo.

Parameter names mirror ``torch.nn`` (``weight``/``bias``) and the zoo keeps the reference's module
attribute names, so ``state_dict`` keys are interchangeable with checkpoints of the reference
``Net/*`` classes (parity is asserted in ``tests/test_models.py``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class GroupNormAct(nn.Module):
    """``act(GroupNorm(x) [+ residual])`` as ONE fused op (``ops.group_norm_act``).

    Affine parameters are kept in fp32 even when the rest of the model runs in bf16
    (``_dlb_keep_fp32``): they are read directly by the fused kernels."""

    def __init__(self, num_groups: int, num_channels: int, relu: bool = True, eps: float = 1e-5):
        super().__init__()
        if num_channels % num_groups:
            raise ValueError("num_channels must be divisible by num_groups")
        self.num_groups, self.num_channels, self.relu, self.eps = num_groups, num_channels, relu, eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.weight._dlb_keep_fp32 = True
        self.bias._dlb_keep_fp32 = True

    def forward(self, x, residual=None):
        return ops.group_norm_act(x, self.num_groups, self.weight, self.bias, self.eps, self.relu, residual)

    def extra_repr(self):
        return f"{self.num_groups}, {self.num_channels}, relu={self.relu}"


class Conv2d(nn.Conv2d):
    """nn.Conv2d (same init, same parameters) routed through ``ops.conv2d``."""

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], self.groups)


class Linear(nn.Linear):
    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class MaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d routed through ``ops.max_pool2d`` (NHWC kernel with a gather backward on CUDA)."""

    def forward(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else self.kernel_size[0]
        s = self.stride if isinstance(self.stride, int) else self.stride[0]
        p = self.padding if isinstance(self.padding, int) else self.padding[0]
        return ops.max_pool2d(x, k, s, p)
