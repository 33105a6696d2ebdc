"""Device-side image augmentation (uint8 NHWC → normalised channels-last float tensor).

Capability of the reference's torchvision pipeline — ``RandomCrop(32, padding=4)``,
``RandomHorizontalFlip``, ``ToTensor``, ``Normalize`` (reference ``dataloader.py:68-75``) — as
one kernel on the device (``csrc/misc.cu``), or a vectorised PyTorch version on CPU.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from . import _native as nat


def _reference(images_u8: torch.Tensor, mean, std, pad: int, flip: bool, seed: int, step: int, dtype) -> torch.Tensor:
    b, h, w, c = images_u8.shape
    x = images_u8.to(torch.float32) / 255.0
    if pad > 0 or flip:
        g = torch.Generator().manual_seed(int(seed) * 1000003 + int(step))
        if pad > 0:
            xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pad, pad, pad, pad)).permute(0, 2, 3, 1)
            oy = torch.randint(0, 2 * pad + 1, (b,), generator=g)
            ox = torch.randint(0, 2 * pad + 1, (b,), generator=g)
            rows = (oy[:, None] + torch.arange(h)[None, :])                     # [b,h]
            cols = (ox[:, None] + torch.arange(w)[None, :])                     # [b,w]
            x = xp[torch.arange(b)[:, None, None], rows[:, :, None], cols[:, None, :]]
        if flip:
            f = torch.rand(b, generator=g) < 0.5
            x = torch.where(f[:, None, None, None], x.flip(2), x)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 1, 1, c)
    s = torch.tensor(std, dtype=torch.float32).view(1, 1, 1, c)
    x = ((x - m) / s).to(dtype)
    return x.permute(0, 3, 1, 2)          # logical NCHW over NHWC memory (= channels_last)


def augment(images_u8: torch.Tensor, mean: Sequence[float], std: Sequence[float], pad: int = 0, flip: bool = False,
            seed: int = 0, step: int = 0, dtype: torch.dtype = torch.float32, out: torch.Tensor = None,
            step_tensor: torch.Tensor = None) -> torch.Tensor:
    """→ tensor of logical shape [B,C,H,W] in channels_last memory."""
    b, h, w, c = images_u8.shape
    if images_u8.is_cuda and nat.available():
        lib = nat.require()
        if out is None:
            out = torch.empty((b, c, h, w), dtype=dtype, device=images_u8.device, memory_format=torch.channels_last) \
                if c > 1 else torch.empty((b, c, h, w), dtype=dtype, device=images_u8.device)
        mean_a = (ctypes.c_float * 4)(*([float(m) for m in mean] + [0.0] * (4 - c)))
        std_a = (ctypes.c_float * 4)(*([float(s) for s in std] + [1.0] * (4 - c)))
        nat.check(lib.dlb_augment(images_u8.data_ptr(), out.data_ptr(), nat.dtype_code(out.dtype), b, h, w, c, int(pad),
                                  int(bool(flip)), ctypes.addressof(mean_a), ctypes.addressof(std_a),
                                  (int(seed) * 2654435761 + (0 if step_tensor is not None else int(step))) & 0xFFFFFFFF,
                                  nat.ptr(step_tensor), nat.stream_ptr(images_u8.device)), "augment")
        return out
    res = _reference(images_u8.cpu(), mean, std, pad, flip, seed, step, dtype).to(images_u8.device)
    if out is not None:
        out.copy_(res)
        return out
    return res
