"""`_ProposalLayer` of lib/model/rpn/proposal_layer.py:26-163: decode, clip, sort, top-N, NMS and padding run as one
kernel chain on the device for the whole batch (no per-frame Python loop, no host round trip)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..utils.config import cfg
from .generate_anchors import generate_anchors
from ... import ops


class _ProposalLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super().__init__()
        self._feat_stride = feat_stride
        self._anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self._num_anchors = self._anchors.size(0)

    def forward(self, input, target=False, from_scores=False):
        """`input` as in the reference: (rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key).  With `from_scores` the first
        entry is the raw rpn_cls_score and the softmax of rpn.py:66-68 runs inside the decode kernel (model/rpn/rpn.py)."""
        cls_prob, bbox_deltas, im_info, cfg_key = input[0], input[1], input[2], input[3]
        pre_nms_topN = cfg[cfg_key].RPN_PRE_NMS_TOP_N
        post_nms_topN = cfg[cfg_key].RPN_POST_NMS_TOP_N
        if target:
            post_nms_topN = cfg[cfg_key].RPN_POST_NMS_TOP_N_TARGET
        nms_thresh = cfg[cfg_key].RPN_NMS_THRESH
        if self._anchors.device != cls_prob.device:
            self._anchors = self._anchors.to(cls_prob.device)
        with torch.no_grad():
            return ops.proposal_forward(cls_prob, bbox_deltas, im_info, self._anchors, int(self._feat_stride),
                                        int(pre_nms_topN), int(post_nms_topN), float(nms_thresh),
                                        from_scores=from_scores)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass

    def reshape(self, bottom, top):
        """Reshaping happens during the call to forward."""
        pass
