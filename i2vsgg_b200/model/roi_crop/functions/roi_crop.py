"""`RoICropFunction` of lib/model/roi_crop/functions/roi_crop.py:7-24 over the sm_100a kernels."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .... import ops


class _RoICropOp(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        ctx.save_for_backward(input2)
        ctx.feature_size = tuple(input1.shape)
        return ops.roi_crop_forward(input1, input2)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (grids,) = ctx.saved_tensors
        # functions/roi_crop.py:19-24: both gradients start at zero; the kernel only ever writes the features' one
        return ops.roi_crop_backward(grad_output, grids, ctx.feature_size), torch.zeros_like(grids)


class RoICropFunction:
    """Keeps the 0.4-era call shape `RoICropFunction()(input1, input2)`: input1 [B,C,H,W] features, input2 [N,oh,ow,2]
    sampling grid (y, x) in [-1, 1] -> [N,C,oh,ow]."""

    def __call__(self, input1, input2):
        return _RoICropOp.apply(input1, input2)
