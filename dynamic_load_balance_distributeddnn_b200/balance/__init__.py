from .reallocator import AffineReallocator, Reallocator, get_size, integer_split, reference_split, throughput_shares
from .tracker import TimeTracker

__all__ = ["AffineReallocator", "Reallocator", "get_size", "integer_split", "reference_split", "throughput_shares",
           "TimeTracker"]
