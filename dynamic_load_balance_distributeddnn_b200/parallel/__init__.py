from .buckets import FlatState
from .comm import Comm, SingleComm, SymmComm, TorchComm, make_comm

__all__ = ["FlatState", "Comm", "SingleComm", "SymmComm", "TorchComm", "make_comm"]
