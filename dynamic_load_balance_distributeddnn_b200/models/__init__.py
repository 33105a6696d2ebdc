from .densenet import DenseNet, DenseNet121, DenseNet161, DenseNet169, DenseNet201
from .googlenet import GoogLeNet
from .mnistnet import MnistNet
from .regnet import RegNet, RegNetX_200MF, RegNetX_400MF, RegNetY_400MF
from .registry import LM_DEFAULTS, build_model, model_names
from .resnet import ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152
from .transformer import TransformerModel

__all__ = ["DenseNet", "DenseNet121", "DenseNet161", "DenseNet169", "DenseNet201", "GoogLeNet", "MnistNet",
           "RegNet", "RegNetX_200MF", "RegNetX_400MF", "RegNetY_400MF", "ResNet", "ResNet18", "ResNet34",
           "ResNet50", "ResNet101", "ResNet152", "TransformerModel", "build_model", "model_names", "LM_DEFAULTS"]
