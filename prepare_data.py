#!/usr/bin/env python
"""Pre-populate ./data like the reference's prepare_data.py (reference prepare_data.py:4-7).  There is no network in
the target environment, so this first tries torchvision's downloader and otherwise materialises the synthetic stand-ins
(same shapes / class counts / split sizes) that the trainer falls back to anyway."""
import sys

from dynamic_load_balance_distributeddnn_b200.data import load_image_dataset


def main(root: str = "./data") -> None:
    try:
        from torchvision import datasets
        for cls in (datasets.FashionMNIST, datasets.CIFAR10, datasets.CIFAR100):
            for train in (True, False):
                cls(root, train=train, download=True)
        print("downloaded FashionMNIST / CIFAR10 / CIFAR100 into", root)
    except Exception as e:  # noqa: BLE001
        print(f"download unavailable ({type(e).__name__}); using synthetic datasets of the same shape")
        for name in ("mnist", "cifar10", "cifar100"):
            ds = load_image_dataset(name, True, root, synthetic=True, n_override=1024)
            print(f"  {name}: images {tuple(ds.images.shape)} classes {ds.num_classes}")


if __name__ == "__main__":
    main(*sys.argv[1:])
