"""Model zoo under the reference's module names: ``from dynamic_load_balance_distributeddnn_b200 import Net`` then
``Net.Densenet.DenseNet121(10)``, ``Net.Resnet.ResNet50(10)``, ... (reference ``Net/*.py``, selected in ``dbs.py:345-363``).

These are attribute aliases of ``models.*`` — deliberately NOT a top-level ``Net`` package: a top-level package of that name
would shadow the unmodified reference's own ``Net`` namespace package inside ``bench.py --impl reference``."""
from ..models import densenet as Densenet
from ..models import googlenet as GoogleNet
from ..models import mnistnet as MnistNet
from ..models import regnet as RegNet
from ..models import resnet as Resnet
from ..models import transformer as Transformer

__all__ = ["Densenet", "GoogleNet", "MnistNet", "RegNet", "Resnet", "Transformer"]
