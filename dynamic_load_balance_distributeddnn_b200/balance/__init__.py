from .reallocator import Reallocator, get_size, integer_split, reference_split, throughput_shares
from .tracker import TimeTracker

__all__ = ["Reallocator", "get_size", "integer_split", "reference_split", "throughput_shares",
           "TimeTracker"]
