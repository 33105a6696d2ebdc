"""MnistNet: 2×(5×5 valid conv → 2×2 max-pool → ReLU) → FC(320,50) → FC(50,10) → log_softmax.

Same architecture, parameter names and shapes as reference ``Net/MnistNet.py:9-27`` (21 840 params in
8 tensors).  The reference trains it with ``F.cross_entropy`` on top of the
``log_softmax`` output (``dbs.py:374``) — idempotent, kept for parity.
"""
import torch.nn as nn
import torch.nn.functional as F

from .layers import Conv2d, Linear
from .. import ops


class MnistNet(nn.Module):
    input_shape = (1, 28, 28)

    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.conv1 = Conv2d(1, 10, kernel_size=5)
        self.conv2 = Conv2d(10, 20, kernel_size=5)
        self.conv2_drop = nn.Dropout2d()
        self.fc1 = Linear(320, 50)
        self.fc2 = Linear(50, num_classes)

    def forward(self, x):
        x = F.relu(ops.max_pool2d(self.conv1(x), 2))
        x = F.relu(ops.max_pool2d(self.conv2_drop(self.conv2(x)), 2))
        x = x.reshape(x.shape[0], -1) if x.is_contiguous() else x.contiguous().reshape(x.shape[0], -1)
        x = F.relu(self.fc1(x))
        x = F.dropout(x, training=self.training)
        return F.log_softmax(self.fc2(x).float(), dim=1)
