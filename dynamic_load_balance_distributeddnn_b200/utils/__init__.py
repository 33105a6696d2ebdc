from .checkpoint import load_checkpoint, save_checkpoint
from .clocks import ClockSampler
from .logging import done_marker, init_logger, log_path
from .stats import StatsRecorder, load_stats
from .tracing import Tracer

__all__ = ["load_checkpoint", "save_checkpoint", "ClockSampler", "done_marker", "init_logger", "log_path",
           "StatsRecorder", "load_stats", "Tracer"]


def print_layer(model, layer_name):
    """Return the named parameter (reference ``utils.py:1-4``)."""
    return dict(model.named_parameters()).get(layer_name)
