"""`RoIAlign`, `RoIAlignAvg`, `RoIAlignMax` of lib/model/roi_align/modules/roi_align.py:6-42.

The Avg / Max variants run the (h+1)x(w+1) lattice and the 2x2/stride-1 pool in ONE kernel (forward and backward)
instead of RoIAlign followed by avg_pool2d / max_pool2d."""
from __future__ import annotations

from torch.nn.modules.module import Module

from ..functions.roi_align import RoIAlignFunction, _LatticeRoIAlign


class RoIAlign(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super().__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale)(features, rois)


class RoIAlignAvg(RoIAlign):
    def forward(self, features, rois):
        return _LatticeRoIAlign.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale,
                                      "avg")


class RoIAlignMax(RoIAlign):
    def forward(self, features, rois):
        return _LatticeRoIAlign.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale,
                                      "max")
