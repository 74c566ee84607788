"""`_AnchorTargetLayer` of lib/model/rpn/anchor_target_layer.py:30-193 on the sm_100a kernels.

Same constructor and `forward((rpn_cls_score, gt_boxes, im_info, num_boxes))` -> [labels [B,1,A*H,W], bbox_targets,
bbox_inside_weights, bbox_outside_weights (each [B,4A,H,W])].  Overlaps, label rules, ordered positive / negative lists
and the final scatter run on the device (the [B, N, G] overlap tensor is never materialised); the sub-sampling draws
are numpy's, call for call as in :121-145, so a seeded run disables the anchors the reference disables.  Quirks kept:
the inside test uses the first image's size for the whole batch (:81-84) and the uniform example weight is computed from
the LAST image's label count (:156-158, the loop variable `i` after the loop).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import nn

from ... import _lib
from ..utils.config import cfg
from .generate_anchors import generate_anchors


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _AnchorTargetLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super().__init__()
        self._feat_stride = feat_stride
        self._scales = scales
        self._anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self._num_anchors = self._anchors.size(0)
        self._allowed_border = 0

    def forward(self, input):
        rpn_cls_score, gt_boxes, im_info = input[0], input[1], input[2]
        if not (rpn_cls_score.is_cuda and gt_boxes.is_cuda):
            raise _lib.I2VError("_AnchorTargetLayer: expected CUDA tensors (i2vsgg_b200 has no CPU path)")
        lib = _lib.load()
        dev = gt_boxes.device
        H, W = rpn_cls_score.size(2), rpn_cls_score.size(3)
        gt_boxes = gt_boxes.float().contiguous()
        B, G = gt_boxes.shape[:2]
        A = self._num_anchors
        total = A * H * W
        base = self._anchors.to(dev).contiguous()
        info0 = im_info[0].detach().cpu()
        im_h, im_w = float(int(info0[0])), float(int(info0[1]))                         # long(im_info[0][.]) at :83-84
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        max_ov, labels = torch.empty((B, total), **f32), torch.empty((B, total), **f32)
        argmax = torch.empty((B, total), **i32)
        gt_max = torch.empty((B, G), **i32)
        pos_list, neg_list = torch.empty((B, total), **i32), torch.empty((B, total), **i32)
        counts = torch.empty((B, 2), **i32)
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        stride = int(self._feat_stride)
        with torch.cuda.device(dev):
            _lib.check(lib.i2v_anchor_overlaps(_p(base), _p(gt_boxes), B, A, H, W, stride, G, float(self._allowed_border), im_w,
                                               im_h, _p(max_ov), _p(argmax), _p(gt_max), s), "i2v_anchor_overlaps")
            _lib.check(lib.i2v_anchor_labels(_p(base), _p(gt_boxes), B, A, H, W, stride, G, _p(max_ov), _p(gt_max),
                                             float(cfg.TRAIN.RPN_NEGATIVE_OVERLAP), float(cfg.TRAIN.RPN_POSITIVE_OVERLAP),
                                             int(bool(cfg.TRAIN.RPN_CLOBBER_POSITIVES)), _p(labels), s), "i2v_anchor_labels")
            # positives: label >= 1; negatives: 0 <= label < 1 (don't-care and outside anchors are -1)
            _lib.check(lib.i2v_fg_bg_select(_p(labels), B, total, 1.0, 1.0, 0.0, _p(pos_list), _p(neg_list), _p(counts), s),
                       "i2v_fg_bg_select")
        cnt = counts.cpu().numpy()
        num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)                # :119
        # :121-145, the reference's draws in the reference's order (per image: positives first, then negatives)
        dis_pos, dis_neg = [], []
        kept = np.zeros((B, 2), np.int64)
        for i in range(B):
            sum_fg, sum_bg = int(cnt[i, 0]), int(cnt[i, 1])
            dp = np.zeros((0,), np.int64)
            if sum_fg > num_fg:
                dp = np.random.permutation(sum_fg)[: sum_fg - num_fg]
            fg_left = sum_fg - len(dp)
            num_bg = int(cfg.TRAIN.RPN_BATCHSIZE) - fg_left
            dn = np.zeros((0,), np.int64)
            if sum_bg > num_bg:
                dn = np.random.permutation(sum_bg)[: sum_bg - num_bg]
            dis_pos.append(dp)
            dis_neg.append(dn)
            kept[i] = (fg_left, sum_bg - len(dn))
        for lst, dis in ((pos_list, dis_pos), (neg_list, dis_neg)):
            m = max((len(d) for d in dis), default=0)
            if m == 0:
                continue
            pos = np.zeros((B, m), np.int32)
            n = np.zeros((B,), np.int32)
            for i, d in enumerate(dis):
                pos[i, : len(d)] = d
                n[i] = len(d)
            pos_d, n_d = torch.from_numpy(pos).to(dev), torch.from_numpy(n).to(dev)
            with torch.cuda.device(dev):
                _lib.check(lib.i2v_anchor_disable(_p(labels), _p(lst), _p(pos_d), _p(n_d), B, total, m, s), "i2v_anchor_disable")
        if cfg.TRAIN.RPN_POSITIVE_WEIGHT < 0:                                            # :155-158
            num_examples = int(kept[B - 1].sum())
            pos_w = neg_w = 1.0 / num_examples
        else:
            raise NotImplementedError("RPN_POSITIVE_WEIGHT >= 0: the reference leaves the weights undefined on this branch "
                                      "(anchor_target_layer.py:159-161)")
        labels_out = torch.empty((B, 1, A * H, W), **f32)
        targets = torch.empty((B, 4 * A, H, W), **f32)
        inside, outside = torch.empty((B, 4 * A, H, W), **f32), torch.empty((B, 4 * A, H, W), **f32)
        with torch.cuda.device(dev):
            _lib.check(lib.i2v_anchor_targets_finalize(_p(base), _p(gt_boxes), B, A, H, W, stride, G, _p(max_ov), _p(argmax),
                                                       _p(labels), float(cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0]), float(pos_w),
                                                       float(neg_w), _p(labels_out), _p(targets), _p(inside), _p(outside), s),
                       "i2v_anchor_targets_finalize")
        return [labels_out, targets, inside, outside]

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""

    def reshape(self, bottom, top):
        """Reshaping happens during the call to forward."""
