"""dlb-b200: dynamic-batch-size data-parallel training, native to Blackwell (sm_100a)."""
__version__ = "0.1.0"

from .config import DBSConfig  # noqa: F401
