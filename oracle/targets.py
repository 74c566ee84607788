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


def anchor_target_layer(gt_boxes, im_info, base_anchors, feat_h, feat_w, feat_stride=16, neg=0.3, pos=0.7, clobber=False,
                        fg_fraction=0.5, batchsize=256, inside_weight=1.0, allowed_border=0):
    """lib/model/rpn/anchor_target_layer.py:48-193 in numpy (float32, the reference's operation order and numpy draws)
    -> [labels [B,1,A*H,W], targets, inside, outside (each [B,4A,H,W])]."""
    gt = np.asarray(gt_boxes, F)
    B, G = gt.shape[:2]
    A = base_anchors.shape[0]
    sx, sy = np.meshgrid(np.arange(feat_w) * feat_stride, np.arange(feat_h) * feat_stride)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], 1).astype(F)
    allan = (np.asarray(base_anchors, F)[None] + shifts[:, None]).reshape(-1, 4)
    total = allan.shape[0]
    im_h, im_w = int(im_info[0][0]), int(im_info[0][1])
    keep = ((allan[:, 0] >= -allowed_border) & (allan[:, 1] >= -allowed_border) & (allan[:, 2] < im_w + allowed_border) &
            (allan[:, 3] < im_h + allowed_border))
    inds = np.nonzero(keep)[0]
    anchors = allan[inds]
    ov = overlaps_batch(np.broadcast_to(anchors[None], (B,) + anchors.shape), gt)            # :98
    max_ov, argmax = ov.max(2), ov.argmax(2)
    gt_max = ov.max(1)
    labels = np.full((B, len(inds)), -1, F)
    if not clobber:
        labels[max_ov < F(neg)] = 0
    gt_max[gt_max == 0] = F(1e-5)
    k = (ov == gt_max[:, None, :]).sum(2)
    if k.sum() > 0:
        labels[k > 0] = 1
    labels[max_ov >= F(pos)] = 1
    if clobber:
        labels[max_ov < F(neg)] = 0
    num_fg = int(fg_fraction * batchsize)
    sum_fg, sum_bg = (labels == 1).sum(1), (labels == 0).sum(1)
    for i in range(B):
        if sum_fg[i] > num_fg:
            fg = np.nonzero(labels[i] == 1)[0]
            labels[i][fg[np.random.permutation(len(fg))[: len(fg) - num_fg]]] = -1
        num_bg = batchsize - (labels[i] == 1).sum()
        if sum_bg[i] > num_bg:
            bg = np.nonzero(labels[i] == 0)[0]
            labels[i][bg[np.random.permutation(len(bg))[: len(bg) - num_bg]]] = -1
    gsel = np.take_along_axis(gt[:, :, :4], argmax[..., None], 1)
    t = transform_batch(np.broadcast_to(anchors[None], gsel.shape), gsel)
    inside = np.zeros_like(labels)
    inside[labels == 1] = F(inside_weight)
    w = F(1.0 / int((labels[B - 1] >= 0).sum()))                                           # :156-158, the last image
    outside = np.zeros_like(labels)
    outside[labels == 1] = w
    outside[labels == 0] = w

    def unmap(d, fill):
        out = np.full((B, total) + d.shape[2:], fill, F)
        out[:, inds] = d
        return out

    H, W = feat_h, feat_w
    lab = unmap(labels, -1).reshape(B, H, W, A).transpose(0, 3, 1, 2).reshape(B, 1, A * H, W)
    tg = unmap(t, 0).reshape(B, H, W, A * 4).transpose(0, 3, 1, 2)
    iw = np.repeat(unmap(inside, 0)[..., None], 4, 2).reshape(B, H, W, 4 * A).transpose(0, 3, 1, 2)
    ow = np.repeat(unmap(outside, 0)[..., None], 4, 2).reshape(B, H, W, 4 * A).transpose(0, 3, 1, 2)
    return [np.ascontiguousarray(a) for a in (lab, tg, iw, ow)]
