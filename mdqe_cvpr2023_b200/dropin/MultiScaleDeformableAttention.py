"""Drop-in for the reference's compiled extension module of the same name.

The reference does ``import MultiScaleDeformableAttention as MSDA``
(/root/reference/mdqe/models/ops/functions/ms_deform_attn_func.py:19) and calls
``MSDA.ms_deform_attn_forward`` / ``MSDA.ms_deform_attn_backward`` (src/vision.cpp:13-16).  Put this
directory on PYTHONPATH (or call ``mdqe_cvpr2023_b200.install_dropin()``) and the reference's own
func.py / ms_deform_attn.py run unchanged on the B200 kernels.
"""
from mdqe_cvpr2023_b200.ops import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401

__all__ = ["ms_deform_attn_forward", "ms_deform_attn_backward"]
