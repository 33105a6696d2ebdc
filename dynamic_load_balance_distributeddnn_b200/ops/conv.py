"""tcgen05 implicit-GEMM convolution dispatch (filled in by the GEMM/conv build stage).

``tc_supported`` gates which shapes take the hand-written TMA→SMEM→tcgen05.mma→TMEM path
(``csrc/gemm_tc.cu``); everything else goes to the vendor library.
"""
from __future__ import annotations

import os


from . import _native as nat

_ENABLED = os.environ.get("DLB_TC_CONV", "1") == "1"


def tc_supported(x, weight, stride, padding, groups) -> bool:
    if not _ENABLED or not nat.available():
        return False
    try:
        from . import gemm_tc
    except Exception:
        return False
    return gemm_tc.conv_supported(x, weight, stride, padding, groups)


def conv2d_tc(x, weight, bias, stride, padding):
    from . import gemm_tc
    return gemm_tc.conv2d(x, weight, bias, stride, padding)
