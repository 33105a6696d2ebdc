from .lr_policy import lr_at_epoch
from .trainer import Trainer

__all__ = ["Trainer", "lr_at_epoch"]
