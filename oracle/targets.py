"""Proposal-target assignment restated in numpy -- TEST INFRASTRUCTURE ONLY.

Follows lib/model/rpn/proposal_target_layer_cascade.py:33-212 with `bbox_overlaps_batch`
(lib/model/rpn/bbox_transform.py:215-257) and `bbox_transform_batch` (:54-75); all arithmetic in float32 in the
reference's operation order, random draws through `np.random` in the reference's order (seed it like the reference run).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def overlaps_batch(rois4, gt):
    """rois4 [B,R,4], gt [B,K,5] -> overlaps [B,R,K] (bbox_transform.py:224-257)."""
    rois4, gt = rois4.astype(F), gt.astype(F)
    gw = (gt[:, :, 2] - gt[:, :, 0] + F(1)); gh = (gt[:, :, 3] - gt[:, :, 1] + F(1))
    aw = (rois4[:, :, 2] - rois4[:, :, 0] + F(1)); ah = (rois4[:, :, 3] - rois4[:, :, 1] + F(1))
    ga, aa = (gw * gh)[:, None, :], (aw * ah)[:, :, None]
    b, q = rois4[:, :, None, :], gt[:, None, :, :4]
    iw = np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0]) + F(1)
    ih = np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1]) + F(1)
    iw[iw < 0] = 0
    ih[ih < 0] = 0
    ua = aa + ga - iw * ih
    ov = (iw * ih / ua).astype(F)
    ov[np.broadcast_to(((gw == 1) & (gh == 1))[:, None, :], ov.shape)] = 0
    ov[np.broadcast_to(((aw == 1) & (ah == 1))[:, :, None], ov.shape)] = -1
    return ov


def transform_batch(ex, gt):
    """bbox_transform.py:54-75 for [B,S,4] boxes."""
    ew = ex[..., 2] - ex[..., 0] + F(1); eh = ex[..., 3] - ex[..., 1] + F(1)
    ecx = ex[..., 0] + F(0.5) * ew; ecy = ex[..., 1] + F(0.5) * eh
    gw = gt[..., 2] - gt[..., 0] + F(1); gh = gt[..., 3] - gt[..., 1] + F(1)
    gcx = gt[..., 0] + F(0.5) * gw; gcy = gt[..., 1] + F(0.5) * gh
    return np.stack([(gcx - ecx) / ew, (gcy - ecy) / eh, np.log(gw / ew), np.log(gh / eh)], -1).astype(F)


def proposal_target_layer(all_rois, gt_boxes, batch_size=128, fg_fraction=0.25, fg_thresh=0.5, bg_hi=0.5, bg_lo=0.1,
                          means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2), inside=(1, 1, 1, 1), normalize=True):
    all_rois, gt_boxes = np.asarray(all_rois, F), np.asarray(gt_boxes, F)
    B, K = gt_boxes.shape[:2]
    app = np.zeros_like(gt_boxes)
    app[:, :, 1:5] = gt_boxes[:, :, :4]
    rois = np.concatenate([all_rois, app], 1)                                  # :41-45
    S = int(batch_size)
    fg_per = int(np.round(fg_fraction * S)) or 1
    ov = overlaps_batch(rois[:, :, 1:5], gt_boxes)
    max_ov, assign = ov.max(2), ov.argmax(2)                                   # :124 (first maximum)
    labels = np.take_along_axis(gt_boxes[:, :, 4], assign, 1)
    labels_b = np.zeros((B, S), F)
    rois_b = np.zeros((B, S, 5), F)
    gt_b = np.zeros((B, S, 5), F)
    for i in range(B):
        fg = np.nonzero(max_ov[i] >= F(fg_thresh))[0]
        bg = np.nonzero((max_ov[i] < F(bg_hi)) & (max_ov[i] >= F(bg_lo)))[0]
        if len(fg) > 0 and len(bg) > 0:
            nfg = min(fg_per, len(fg))
            fg = fg[np.random.permutation(len(fg))[:nfg]]
            bg = bg[np.floor(np.random.rand(S - nfg) * len(bg)).astype(np.int64)]
        elif len(fg) > 0:
            fg = fg[np.floor(np.random.rand(S) * len(fg)).astype(np.int64)]
            bg = bg[:0]
            nfg = S
        elif len(bg) > 0:
            bg = bg[np.floor(np.random.rand(S) * len(bg)).astype(np.int64)]
            fg = fg[:0]
            nfg = 0
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = np.concatenate([fg, bg])
        labels_b[i] = labels[i][keep]
        labels_b[i, nfg:] = 0
        rois_b[i] = rois[i][keep]
        rois_b[i, :, 0] = i
        gt_b[i] = gt_boxes[i][assign[i][keep]]
    t = transform_batch(rois_b[:, :, 1:5], gt_b[:, :, :4])
    if normalize:
        t = ((t - np.asarray(means, F)) / np.asarray(stds, F)).astype(F)
    pos = labels_b > 0
    targets = np.where(pos[..., None], t, F(0)).astype(F)
    inside_w = np.where(pos[..., None], np.asarray(inside, F), F(0)).astype(F)
    return rois_b, labels_b, targets, inside_w, (inside_w > 0).astype(F)
