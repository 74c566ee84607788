"""`RoIPoolFunction` of lib/model/roi_pooling/functions/roi_pool.py:6-38 over the sm_100a kernels."""
from __future__ import annotations

from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .... import ops


class _RoIPool(Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_h, pooled_w, spatial_scale):
        out, argmax = ops.roi_pool_forward(features, rois, int(pooled_h), int(pooled_w), float(spatial_scale),
                                           ops.ARGMAX_FLAT)
        ctx.save_for_backward(rois, argmax)
        ctx.cfg = (int(pooled_h), int(pooled_w), float(spatial_scale))
        ctx.feature_size = tuple(features.shape)
        ctx.mark_non_differentiable(argmax)
        return out, argmax

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, _grad_argmax):
        rois, argmax = ctx.saved_tensors
        ph, pw, scale = ctx.cfg
        assert grad_output.is_cuda          # functions/roi_pool.py:31
        grad_input = ops.roi_pool_backward(grad_output, rois, argmax, ctx.feature_size, ph, pw, scale,
                                           ops.ARGMAX_FLAT)
        return grad_input, None, None, None, None


class RoIPoolFunction:
    """Keeps the 0.4-era call shape `RoIPoolFunction(ph, pw, scale)(features, rois)`; `.argmax` is set by the call."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.argmax = None

    def __call__(self, features, rois):
        out, self.argmax = _RoIPool.apply(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale)
        return out
