"""Model registry: CLI name → constructor (reference ``dbs.py:345-363`` plus explicit variants)."""
from __future__ import annotations

from typing import Callable, Dict

import torch.nn as nn

from . import densenet, googlenet, mnistnet, regnet, resnet, transformer

# transformer hyper-parameters are literals in the reference's run() (dbs.py:337-343)
LM_DEFAULTS = dict(ntoken=33278, ninp=200, nhead=2, nhid=200, nlayers=2, dropout=0.2)

_CNN: Dict[str, Callable[[int], nn.Module]] = {
    "mnistnet": lambda nc: mnistnet.MnistNet(),
    "resnet": resnet.ResNet101,            # the reference maps `-m resnet` to ResNet-101 (dbs.py:348-350)
    "resnet18": resnet.ResNet18, "resnet34": resnet.ResNet34, "resnet50": resnet.ResNet50,
    "resnet101": resnet.ResNet101, "resnet152": resnet.ResNet152,
    "densenet": densenet.DenseNet121, "densenet121": densenet.DenseNet121, "densenet169": densenet.DenseNet169,
    "densenet201": densenet.DenseNet201, "densenet161": densenet.DenseNet161,
    "googlenet": googlenet.GoogLeNet,
    "regnet": regnet.RegNetY_400MF, "regnetx200": regnet.RegNetX_200MF, "regnetx400": regnet.RegNetX_400MF,
    "regnety400": regnet.RegNetY_400MF,
}


def build_model(name: str, num_classes: int = 10, ntoken: int = 0) -> nn.Module:
    if name == "transformer":
        kw = dict(LM_DEFAULTS)
        if ntoken:
            kw["ntoken"] = ntoken
        return transformer.TransformerModel(**kw)
    if name not in _CNN:
        raise KeyError(f"unknown model {name!r}")
    return _CNN[name](num_classes)


def model_names():
    return ["transformer"] + sorted(_CNN)
