"""`ROIPool` / `roi_pool` of lib/model/roi_layers/roi_pool.py:11-63 (the op behind the absent `model._C`)."""
from __future__ import annotations

from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from ... import ops


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale):
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.input_shape = tuple(input.size())
        output, argmax = ops.roi_pool_forward(input, roi, ctx.output_size[0], ctx.output_size[1], spatial_scale,
                                              ops.ARGMAX_PLANE)
        ctx.save_for_backward(roi, argmax)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        grad_input = ops.roi_pool_backward(grad_output, rois, argmax, ctx.input_shape, ctx.output_size[0],
                                           ctx.output_size[1], ctx.spatial_scale, ops.ARGMAX_PLANE)
        return grad_input, None, None, None


roi_pool = _ROIPool.apply


class ROIPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois):
        return roi_pool(input, rois, self.output_size, self.spatial_scale)

    def __repr__(self):
        return f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale})"
