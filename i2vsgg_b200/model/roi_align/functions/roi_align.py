"""`RoIAlignFunction` of lib/model/roi_align/functions/roi_align.py:7-51 over the sm_100a kernels."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .... import ops


class _LatticeRoIAlign(Function):
    """Lattice RoIAlign with an optional fused 2x2/stride-1 pool; one autograd node for the whole module."""

    @staticmethod
    def forward(ctx, features, rois, pooled_h, pooled_w, spatial_scale, pool):
        ctx.save_for_backward(rois, features if pool == "max" else None)
        ctx.cfg = (int(pooled_h), int(pooled_w), float(spatial_scale), pool)
        ctx.feature_size = tuple(features.shape)
        return ops.roi_align_forward(features, rois, int(pooled_h), int(pooled_w), float(spatial_scale), pool)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, features = ctx.saved_tensors
        ph, pw, scale, pool = ctx.cfg
        assert grad_output.is_cuda          # functions/roi_align.py:38
        grad_input = ops.roi_align_backward(grad_output, features, rois, ctx.feature_size, ph, pw, scale, pool)
        return grad_input, None, None, None, None, None


class RoIAlignFunction:
    """Keeps the 0.4-era call shape `RoIAlignFunction(ah, aw, scale)(features, rois)`."""

    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _LatticeRoIAlign.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale,
                                      "none")
