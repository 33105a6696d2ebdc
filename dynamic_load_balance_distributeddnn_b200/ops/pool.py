"""Non-overlapping average pooling on channels-last tensors (``csrc/pool.cu``), with a PyTorch fallback."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _native as nat


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        n, c, h, w = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((n, c, h // k, w // k), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        nat.check(nat.require().dlb_avgpool_nhwc(0, nat.dtype_code(x.dtype), x.data_ptr(), y.data_ptr(), n, h, w, c, k,
                                                 nat.stream_ptr(x.device)), "avgpool_fwd")
        ctx.cfg = (n, c, h, w, k)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, k = ctx.cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        nat.check(nat.require().dlb_avgpool_nhwc(1, nat.dtype_code(dy.dtype), dy.data_ptr(), dx.data_ptr(), n, h, w, c, k,
                                                 nat.stream_ptr(dy.device)), "avgpool_bwd")
        return dx, None


def avg_pool2d(x: torch.Tensor, k: int) -> torch.Tensor:
    """``F.avg_pool2d(x, k)`` (stride k)."""
    if x.is_cuda and nat.available() and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] > 1 \
            and x.shape[2] % k == 0 and x.shape[3] % k == 0:
        return _AvgPoolFn.apply(x, k)
    return F.avg_pool2d(x, k)
