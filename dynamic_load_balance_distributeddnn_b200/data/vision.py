"""Image datasets as flat uint8 NHWC arrays + device-side augmentation.

The reference builds ``torchvision`` datasets with ``download=True`` every epoch and runs PIL
transforms one image at a time in the training process (``dataloader.py:59-99,112-115``;
SURVEY D11/D12).  Here a dataset is a single ``uint8 [N,H,W,C]`` tensor (pinned when CUDA is
present) plus an ``int64 [N]`` label tensor, built once per run; a batch is a host gather into
a pinned staging buffer → one async H2D copy → ONE augmentation kernel on the device
(random-crop-with-padding + horizontal flip + normalise + cast to NHWC bf16/fp32;
``ops/augment.py``).  The test split is only normalised (the reference also applies the random
train-time transforms to it, D12).

Real data is used when the standard torchvision files are already on disk (no network here);
otherwise a synthetic set with the same shapes, class count and split sizes is generated:
class-conditional colour blobs + noise, so a model can actually learn it.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

SPECS = {
    # name: (H, W, C, classes, train N, test N, mean, std, crop padding, flip)
    "mnist": (28, 28, 1, 10, 60000, 10000, (0.1307,), (0.3081,), 0, False),
    "cifar10": (32, 32, 3, 10, 50000, 10000, (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010), 4, True),
    "cifar100": (32, 32, 3, 100, 50000, 10000, (0.5071, 0.4865, 0.4409), (0.2673, 0.2564, 0.2762), 4, True),
}


@dataclass
class ImageDataset:
    name: str
    images: torch.Tensor          # uint8 [N,H,W,C]
    labels: torch.Tensor          # int64 [N]
    num_classes: int
    mean: Tuple[float, ...]
    std: Tuple[float, ...]
    pad: int
    flip: bool
    synthetic: bool

    def __len__(self) -> int:
        return self.images.shape[0]

    @property
    def shape(self):
        return tuple(self.images.shape[1:])


def _synthetic_images(n: int, h: int, w: int, c: int, classes: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, classes, (n,), generator=g)
    # per-class low-frequency template (so the task is learnable) + per-sample noise
    tg = torch.Generator().manual_seed(977)
    templ = torch.rand(classes, 4, 4, c, generator=tg)
    templ = torch.nn.functional.interpolate(templ.permute(0, 3, 1, 2), size=(h, w), mode="bilinear",
                                            align_corners=False).permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, c, dtype=torch.uint8)
    chunk = 8192
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        noise = torch.rand(e - s, h, w, c, generator=g)
        img = 0.65 * templ[labels[s:e]] + 0.35 * noise
        out[s:e] = (img * 255.0).clamp_(0, 255).to(torch.uint8)
    return out, labels.to(torch.int64)


def _try_torchvision(name: str, root: str, train: bool):
    try:
        from torchvision import datasets
        cls = {"mnist": datasets.FashionMNIST, "cifar10": datasets.CIFAR10, "cifar100": datasets.CIFAR100}[name]
        ds = cls(root, train=train, download=False)
    except Exception:
        return None
    data = ds.data
    if isinstance(data, np.ndarray):
        data = torch.from_numpy(data)
    if data.dim() == 3:
        data = data.unsqueeze(-1)
    labels = torch.as_tensor(ds.targets, dtype=torch.int64)
    return data.contiguous().to(torch.uint8), labels


_DATASETS = {}


def load_image_dataset(name: str, train: bool, root: str = "./data", synthetic: Optional[bool] = None,
                       n_override: int = 0, seed: int = 1234, pin: bool = False) -> ImageDataset:
    """Build (once per process) the named dataset split."""
    key = (name, train, root, synthetic, n_override, seed)
    if key in _DATASETS:
        return _DATASETS[key]
    h, w, c, classes, n_train, n_test, mean, std, pad, flip = SPECS[name]
    real = None
    if not synthetic:
        real = _try_torchvision(name, root, train)
        if real is None and synthetic is False:
            raise FileNotFoundError(f"{name} not found under {root} (no network: pre-populate or use --synthetic true)")
    if real is not None:
        images, labels = real
        if n_override:
            images, labels = images[:n_override], labels[:n_override]
        is_syn = False
    else:
        n = n_override or (n_train if train else n_test)
        images, labels = _synthetic_images(n, h, w, c, classes, seed + (0 if train else 1))
        is_syn = True
    if pin and torch.cuda.is_available():
        images = images.pin_memory()
        labels = labels.pin_memory()
    ds = ImageDataset(name, images, labels, classes, mean, std, pad if train else 0, flip if train else False, is_syn)
    _DATASETS[key] = ds
    return ds


class BatchStager:
    """Host gather → pinned staging → async H2D on a copy stream, ``slots`` buffers deep (the host only blocks when it is
    ``slots`` steps ahead of the device: the wait inside ``stage`` is back-pressure, not host work).

    ``stage(indices)`` returns device tensors ``(uint8 [b,H,W,C], int64 [b])`` valid on the
    current stream.  On CPU it is a plain index_select."""

    def __init__(self, ds: ImageDataset, max_batch: int, device: torch.device, slots: int = 4):
        self.ds = ds
        self.slots = slots
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.bytes_per_step = 0
        self.blocked_s = 0.0          # host time spent waiting for a slot (back-pressure: the device is the bottleneck)
        if self.cuda:
            h, w, c = ds.shape
            self._host_img = [torch.empty(max_batch, h, w, c, dtype=torch.uint8).pin_memory() for _ in range(slots)]
            self._host_lab = [torch.empty(max_batch, dtype=torch.int64).pin_memory() for _ in range(slots)]
            self._dev_img = [torch.empty(max_batch, h, w, c, dtype=torch.uint8, device=self.device) for _ in range(slots)]
            self._dev_lab = [torch.empty(max_batch, dtype=torch.int64, device=self.device) for _ in range(slots)]
            self._stream = torch.cuda.Stream(self.device)
            self._done = [torch.cuda.Event() for _ in range(slots)]
            self._consumed = [torch.cuda.Event() for _ in range(slots)]
            self._slot = 0

    def stage(self, indices) -> Tuple[torch.Tensor, torch.Tensor]:
        idx = torch.as_tensor(indices, dtype=torch.int64)
        b = idx.numel()
        if not self.cuda:
            return self.ds.images.index_select(0, idx), self.ds.labels.index_select(0, idx)
        s = self._slot
        self._slot = (s + 1) % self.slots
        t0 = time.perf_counter()
        self._consumed[s].synchronize()           # previous user of this slot has finished
        self.blocked_s += time.perf_counter() - t0
        torch.index_select(self.ds.images, 0, idx, out=self._host_img[s][:b])
        torch.index_select(self.ds.labels, 0, idx, out=self._host_lab[s][:b])
        with torch.cuda.stream(self._stream):
            self._dev_img[s][:b].copy_(self._host_img[s][:b], non_blocking=True)
            self._dev_lab[s][:b].copy_(self._host_lab[s][:b], non_blocking=True)
            self._done[s].record(self._stream)
        torch.cuda.current_stream(self.device).wait_event(self._done[s])
        self.bytes_per_step = b * (self.ds.images[0].numel() + 8)
        self._last = s
        return self._dev_img[s][:b], self._dev_lab[s][:b]

    def release(self) -> None:
        """Mark the most recently staged slot as consumed on the current stream."""
        if self.cuda:
            self._consumed[self._last].record(torch.cuda.current_stream(self.device))
