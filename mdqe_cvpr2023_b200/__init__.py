"""B200-native (sm_100a) implementation of MDQE's data-parallel hot path: multi-scale deformable
attention forward/backward and the mask-logit contraction, behind the reference's operator API."""
import importlib
import os
import sys

from . import _lib  # noqa: F401
from .functions import (MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnFusedJointFunction,  # noqa: F401
                        MSDeformAttnGroupedFunction, mask_logits, tc_linear)
from .consumers import (aligned_bilinear, mask_losses, mask_match_cost, mask_nms_siou, mask_track_siou,  # noqa: F401
                        query_init_sample)
from .modules import MSDeformAttn  # noqa: F401
from .ops import (mask_logits_backward, mask_logits_forward, ms_deform_attn_backward,  # noqa: F401
                  ms_deform_attn_forward)

__all__ = ["MSDeformAttn", "MSDeformAttnFunction", "MSDeformAttnGroupedFunction", "MSDeformAttnFusedFunction", "MSDeformAttnFusedJointFunction", "mask_logits", "tc_linear", "ms_deform_attn_forward",
           "ms_deform_attn_backward", "mask_logits_forward", "mask_logits_backward", "install_dropin",
           "mask_match_cost", "mask_losses", "mask_nms_siou", "mask_track_siou", "aligned_bilinear", "query_init_sample"]


def install_dropin():
    """Register the module `MultiScaleDeformableAttention` (the name the reference imports at
    ms_deform_attn_func.py:19) in sys.modules, backed by this package's kernels."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
    if path not in sys.path:
        sys.path.insert(0, path)
    return importlib.import_module("MultiScaleDeformableAttention")
