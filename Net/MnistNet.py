"""Reference-compatible import path (`import Net.MnistNet`) — re-exports the B200-native implementation."""
from dynamic_load_balance_distributeddnn_b200.models.mnistnet import *  # noqa: F401,F403
from dynamic_load_balance_distributeddnn_b200.models import mnistnet as _impl

__all__ = [n for n in dir(_impl) if not n.startswith('_')]
