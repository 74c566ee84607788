"""SGG stage pieces of the hot path with the reference's shapes and dtypes.

``build_pairs``       faster_rcnn_SGG_emb.py:597-606 (ordered pairs) + :649-656 (union boxes, dual masks) in one launch
``detection_output``  lib/utils.py:584-627 (top-100 triplets of a frame) on the device
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)


def build_pairs(pred_boxes, im_h: float, im_w: float, device="cuda", margin: float = 10.0, want_masks: bool = True):
    """pred_boxes [N,4] -> (ixs [P] int64, ixo [P] int64, rel_boxes [P,5] fp32, SpatialFea [P,2,32,32] fp32) on `device`.

    P = N*(N-1) ordered pairs, i-major (`for i: for j: if i != j`), rel_boxes[:,0] = 0, union boxes grown by `margin`
    and clipped to [0, im_w] x [0, im_h] (resnet_SGG_emb.py:240-244), masks as resnet_SGG_emb.py:246-256."""
    boxes = _dev_f32(pred_boxes, device).reshape(-1, 4)
    return ops.pair_build(boxes, float(im_h), float(im_w), float(margin), want_masks)


_UNIQUE_CACHE = {}


def unordered_pairs(num_boxes: int, device="cuda"):
    """For the ordered pair list of `build_pairs` (p = i*(N-1) + j - (j > i)): the positions `rep` [U] of the pairs with
    i < j (one representative per unordered pair, U = N(N-1)/2) and `inverse` [P] mapping every ordered pair to its
    representative.  The union boxes of (i,j) and (j,i) are the same box, so everything `vrd` computes from rel_boxes
    alone (RoIPool, fc6, fc7, fc8) needs to run on `rel_boxes[rep]` only.  Pure index arithmetic, cached per N."""
    key = (int(num_boxes), str(device))
    if key not in _UNIQUE_CACHE:
        n = int(num_boxes)
        i, j = torch.triu_indices(n, n, 1)
        rep = i * (n - 1) + j - 1
        u = torch.arange(rep.numel())
        inverse = torch.empty((n * (n - 1),), dtype=torch.int64)
        inverse[rep] = u
        inverse[j * (n - 1) + i] = u              # the pair (j, i), j > i, sits at j*(N-1) + i
        _UNIQUE_CACHE[key] = (rep.to(device), inverse.to(device))
    return _UNIQUE_CACHE[key]


def frame_triplets(rel_score, confs, classes, boxes, ixs, ixo, top_k: int = 100):
    """-> (records [top_k,13] fp32, count int32[1]); record = (conf, cls_s, rel, cls_o, sub box, obj box, pair idx)."""
    dev = rel_score.device
    return ops.triplet_topk(rel_score, _dev_f32(confs, dev), torch.as_tensor(classes, device=dev).long(),
                            _dev_f32(boxes, dev), torch.as_tensor(ixs, device=dev).long(),
                            torch.as_tensor(ixo, device=dev).long(), top_k)


def detection_output(vrd_data):
    """Drop-in for lib/utils.py:584-627: same dict in, same five values out (numpy), selection done on the device."""
    if len(vrd_data["bboxes"]) <= 1:
        return None, None, None, None, None
    rel_score = vrd_data["rel_score"]
    if not isinstance(rel_score, torch.Tensor):
        rel_score = torch.as_tensor(np.asarray(rel_score, np.float32))
    rel_score = rel_score.detach().float().cuda()
    rec, cnt = frame_triplets(rel_score, vrd_data["scores"], np.asarray(vrd_data["classes"]),
                              np.asarray(vrd_data["bboxes"], np.float32), np.asarray(vrd_data["ixs"]),
                              np.asarray(vrd_data["ixo"]), 100)
    rec = rec.cpu().numpy()
    k = int(cnt.item())
    rlp_labels_im = np.zeros((100, 3), dtype=np.float64)
    sub_bboxes_im = np.zeros((100, 4), dtype=np.float64)
    obj_bboxes_im = np.zeros((100, 4), dtype=np.float64)
    rlp_labels_im[:k] = rec[:k, 1:4]
    sub_bboxes_im[:k] = rec[:k, 4:8]
    obj_bboxes_im[:k] = rec[:k, 8:12]
    return rlp_labels_im, rec[:k, 0].copy(), sub_bboxes_im, obj_bboxes_im, rec[:k, 12].astype(np.int64)
