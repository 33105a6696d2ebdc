from . import _native
from .augment import augment
from .functional import conv2d, cross_entropy, group_norm_act, linear, nll_loss
from .norm import group_norm_act_reference
from .pool import avg_pool2d, global_avg_pool2d, max_pool2d
from .transformer_ops import (add_layer_norm, add_layer_norm_reference, causal_attention,
                              causal_attention_reference, linear_cross_entropy, linear_cross_entropy_reference)

__all__ = ["_native", "augment", "avg_pool2d", "global_avg_pool2d", "max_pool2d", "conv2d", "cross_entropy", "group_norm_act", "group_norm_act_reference",
           "linear", "nll_loss", "add_layer_norm", "add_layer_norm_reference", "causal_attention",
           "causal_attention_reference", "linear_cross_entropy", "linear_cross_entropy_reference"]
