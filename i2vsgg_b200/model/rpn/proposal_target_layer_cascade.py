"""`_ProposalTargetLayer` of lib/model/rpn/proposal_target_layer_cascade.py:20-212 on the sm_100a kernels.

Same constructor and `forward(all_rois, gt_boxes, num_boxes)` -> (rois, labels, bbox_targets, bbox_inside_weights,
bbox_outside_weights).  The overlap reduction, the ordered foreground / background lists and the final gather run on
the device; the random sample is drawn on the host with exactly the numpy calls of :151-180, in the same order, so that
a seeded run picks the RoIs the reference picks (the per-image counts it needs are one small device-to-host copy, where
the reference synchronises once per image on `.numel()`).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import nn

from ... import _lib
from ..utils.config import cfg


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _ProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super().__init__()
        self._num_classes = nclasses
        self.BBOX_NORMALIZE_MEANS = torch.FloatTensor(cfg.TRAIN.BBOX_NORMALIZE_MEANS)
        self.BBOX_NORMALIZE_STDS = torch.FloatTensor(cfg.TRAIN.BBOX_NORMALIZE_STDS)
        self.BBOX_INSIDE_WEIGHTS = torch.FloatTensor(cfg.TRAIN.BBOX_INSIDE_WEIGHTS)

    def forward(self, all_rois, gt_boxes, num_boxes):
        if not (all_rois.is_cuda and gt_boxes.is_cuda):
            raise _lib.I2VError("_ProposalTargetLayer: expected CUDA tensors (i2vsgg_b200 has no CPU path)")
        lib = _lib.load()
        dev = all_rois.device
        gt_boxes = gt_boxes.float().contiguous()
        B, K = gt_boxes.shape[:2]
        # :41-45 the ground-truth boxes join the candidates
        gt_append = gt_boxes.new_zeros(gt_boxes.shape)
        gt_append[:, :, 1:5] = gt_boxes[:, :, :4]
        rois = torch.cat([all_rois.float(), gt_append], 1).contiguous()
        R = rois.size(1)
        rois_per_image = int(cfg.TRAIN.BATCH_SIZE / 1)                                 # :47-48
        fg_rois_per_image = int(np.round(cfg.TRAIN.FG_FRACTION * rois_per_image)) or 1  # :49-50

        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        max_ov, labels = torch.empty((B, R), **f32), torch.empty((B, R), **f32)
        assign, fg_inds, bg_inds = torch.empty((B, R), **i32), torch.empty((B, R), **i32), torch.empty((B, R), **i32)
        counts = torch.empty((B, 2), **i32)
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(lib.i2v_roi_gt_overlaps(_p(rois), 5, _p(gt_boxes), B, R, K, _p(max_ov), _p(assign), _p(labels), s),
                       "i2v_roi_gt_overlaps")
            _lib.check(lib.i2v_fg_bg_select(_p(max_ov), B, R, float(cfg.TRAIN.FG_THRESH), float(cfg.TRAIN.BG_THRESH_HI),
                                            float(cfg.TRAIN.BG_THRESH_LO), _p(fg_inds), _p(bg_inds), _p(counts), s),
                       "i2v_fg_bg_select")
        cnt = counts.cpu().numpy()
        # :136-183 -- the reference's draws, call for call
        positions = np.zeros((B, rois_per_image), np.int32)
        fg_this = np.zeros((B,), np.int32)
        for i in range(B):
            fg_num, bg_num = int(cnt[i, 0]), int(cnt[i, 1])
            if fg_num > 0 and bg_num > 0:
                nfg = min(fg_rois_per_image, fg_num)
                positions[i, :nfg] = np.random.permutation(fg_num)[:nfg]
                positions[i, nfg:] = np.floor(np.random.rand(rois_per_image - nfg) * bg_num)
            elif fg_num > 0:
                nfg = rois_per_image
                positions[i] = np.floor(np.random.rand(rois_per_image) * fg_num)
            elif bg_num > 0:
                nfg = 0
                positions[i] = np.floor(np.random.rand(rois_per_image) * bg_num)
            else:
                raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
            fg_this[i] = nfg
        pos_d = torch.from_numpy(positions).to(dev)
        nfg_d = torch.from_numpy(fg_this).to(dev)
        S = rois_per_image
        rois_out, labels_out = torch.empty((B, S, 5), **f32), torch.empty((B, S), **f32)
        targets, inside, outside = torch.empty((B, S, 4), **f32), torch.empty((B, S, 4), **f32), torch.empty((B, S, 4), **f32)
        arr = lambda t: (ctypes.c_float * 4)(*[float(v) for v in t])
        means, stds, iw = arr(self.BBOX_NORMALIZE_MEANS), arr(self.BBOX_NORMALIZE_STDS), arr(self.BBOX_INSIDE_WEIGHTS)
        with torch.cuda.device(dev):
            _lib.check(lib.i2v_proposal_targets_gather(
                _p(rois), _p(gt_boxes), _p(assign), _p(labels), _p(fg_inds), _p(bg_inds), _p(pos_d), _p(nfg_d), B, R, K, S,
                ctypes.cast(means, ctypes.c_void_p), ctypes.cast(stds, ctypes.c_void_p), ctypes.cast(iw, ctypes.c_void_p),
                int(bool(cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)), _p(rois_out), _p(labels_out), _p(targets),
                _p(inside), _p(outside), s), "i2v_proposal_targets_gather")
        return rois_out, labels_out, targets, inside, outside

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""

    def reshape(self, bottom, top):
        """Reshaping happens during the call to forward."""
