"""The part of `_RPN.forward` (lib/model/rpn/rpn.py:57-78) that follows the head's convolutions: the 2-way softmax over
each anchor's (background, foreground) scores and the proposal layer.  The three convolutions (`RPN_Conv`,
`RPN_cls_score`, `RPN_bbox_pred`, rpn.py:27-38) belong to the backbone side and are not part of this path; this module
takes their outputs."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..utils.config import cfg
from .proposal_layer import _ProposalLayer
from ... import ops


class _RPN(nn.Module):
    """Same constructor data as the reference (`din` is kept for signature compatibility; there are no convolutions
    here).  `forward_head_outputs` is rpn.py:66-78 from `rpn_cls_score` / `rpn_bbox_pred` on."""

    def __init__(self, din=None):
        super().__init__()
        self.din = din
        self.anchor_scales = cfg.ANCHOR_SCALES
        self.anchor_ratios = cfg.ANCHOR_RATIOS
        self.feat_stride = cfg.FEAT_STRIDE[0]
        self.nc_score_out = len(self.anchor_scales) * len(self.anchor_ratios) * 2     # rpn.py:30
        self.nc_bbox_out = len(self.anchor_scales) * len(self.anchor_ratios) * 4      # rpn.py:34
        self.RPN_proposal = _ProposalLayer(self.feat_stride, self.anchor_scales, self.anchor_ratios)

    @staticmethod
    def reshape(x, d):
        """rpn.py:46-55."""
        s = x.size()
        return x.view(s[0], int(d), int(float(s[1] * s[2]) / float(d)), s[3])

    @staticmethod
    def cls_prob(rpn_cls_score):
        """rpn.py:66-68: reshape to two channels, softmax over them, reshape back -- one kernel."""
        return ops.rpn_cls_prob(rpn_cls_score)

    def forward_head_outputs(self, rpn_cls_score, rpn_bbox_pred, im_info, target=False, fused=True):
        """rois of rpn.py:74-78.  `fused`: the probabilities are formed inside the decode kernel and never written;
        otherwise `cls_prob` is materialised first, as the reference does.  Both give the same rois bit for bit."""
        cfg_key = "TRAIN" if self.training else "TEST"
        with torch.no_grad():
            if fused:
                return self.RPN_proposal((rpn_cls_score, rpn_bbox_pred, im_info, cfg_key), target=target, from_scores=True)
            return self.RPN_proposal((self.cls_prob(rpn_cls_score), rpn_bbox_pred, im_info, cfg_key), target=target)
