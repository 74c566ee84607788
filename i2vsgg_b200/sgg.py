"""SGG stage pieces of the hot path with the reference's shapes and dtypes.

``build_pairs``       faster_rcnn_SGG_emb.py:597-606 (ordered pairs) + :649-656 (union boxes, dual masks) in one launch
``detection_output``  lib/utils.py:584-627 (top-100 triplets of a frame) on the device
``association``       lib/utils.py:461-526 + :134-182 (per-frame triplets -> video relations), greedy matching on the device
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


# ------------------------------------------------------------------------------------------------ temporal association
def _fill_empty_frames(counts, invalid_num: int = 4):
    """lib/utils.py:470-518, index arithmetic only: which frame position a frame without predictions borrows them from
    (nearest non-empty neighbour, the earlier one on ties), unless every frame within +-4 positions is empty too."""
    n = len(counts)
    empty = [c == 0 for c in counts]
    src = list(range(n))
    tmp = [-1] * n
    for i in range(n):
        if empty[i]:
            j = i - 1
            while j >= 0 and empty[j]:
                j -= 1
            left = 0 if j < 0 else i - j
            j = i + 1
            while j < n and empty[j]:
                j += 1
            right = 0 if j >= n else j - i
            if right == 0 or (left > 0 and left <= right):
                tmp[i] = i - left
            elif left == 0 or (right > 0 and left > right):
                tmp[i] = i + right
    for i in range(n):
        if tmp[i] >= 0:
            if i < invalid_num:
                start, end = 0, i + invalid_num
            elif i > n - invalid_num - 1:
                start, end = i - invalid_num, n - 1
            else:
                start, end = i - invalid_num, i + invalid_num
            lonely = all(tmp[j] != -1 for j in range(start, end + 1))
            src[i] = -1 if lonely else tmp[i]
    return src


def association(records, counts, frame_numbers=None, max_num_per_video: int = 200, min_length: int = 10,
                objects=None, predicates=None):
    """`association()` of lib/utils.py:461-526 for ONE video whose per-frame triplets are the gathered
    `records [F,100,13]`, `counts [F]` (video order; see `shard.all_gather_triplets`): fills frames without predictions,
    runs the greedy association on the device, keeps relations of at least `min_length` frames, sorts them by score
    (descending, stable) and returns the first `max_num_per_video` as the reference's dictionaries.  `objects` /
    `predicates` map class ids to names as lib/utils.py:34-35 does; without them the ids stay."""
    records = records if isinstance(records, torch.Tensor) else torch.as_tensor(np.asarray(records, np.float32))
    records = records.float().cuda() if not records.is_cuda else records.float()
    cnt_h = [int(c) for c in (counts.tolist() if isinstance(counts, torch.Tensor) else list(counts))]
    F, K = records.shape[0], records.shape[1]
    if F == 0 or all(c == 0 for c in cnt_h):
        return []
    src = _fill_empty_frames(cnt_h)
    fnos = list(range(F)) if frame_numbers is None else [int(x) for x in frame_numbers]
    rel_id, order, info, score, num = ops.greedy_association(records, cnt_h, fnos, src, max_traj=100)
    n = int(num.item())
    rel_id, order = rel_id.cpu().numpy(), order.cpu().numpy()
    info, score = info[:n].cpu().numpy(), score[:n].cpu().numpy()
    rec = records.cpu().numpy().astype(np.float64)
    keep = [r for r in range(n) if info[r, 5] >= min_length]
    keep.sort(key=lambda r: score[r], reverse=True)                  # lib/utils.py:522 (stable, like list.sort)
    keep = keep[:max_num_per_video]
    # the member rows of the kept relations, frame by frame and prediction by prediction (the order the reference appends
    # them in), gathered with array operations: a Python loop over 100 predictions x 1000 frames cost more than the kernel
    out = []
    if keep:
        keep_arr = np.asarray(keep, np.int64)
        fs, js = np.nonzero((rel_id >= 0) & np.isin(rel_id, keep_arr))     # row-major: frames, then predictions
        rs = rel_id[fs, js]
        rows = rec[np.asarray(src, np.int64)[fs], order[fs, js]]
        by_rel = np.argsort(rs, kind="stable")
        rs, rows = rs[by_rel], rows[by_rel]
        lo, hi = np.searchsorted(rs, keep_arr, "left"), np.searchsorted(rs, keep_arr, "right")
    for i, r in enumerate(keep):
        m = rows[lo[i]:hi[i]]
        s, p, o = int(info[r, 2]), int(info[r, 3]), int(info[r, 4])
        out.append({"triplet": [objects[s] if objects is not None else s, predicates[p] if predicates is not None else p,
                                objects[o] if objects is not None else o],
                    "score": float(score[r]), "duration": [int(info[r, 0]), int(info[r, 1])],
                    "sub_traj": m[:, 4:8].tolist(), "obj_traj": m[:, 8:12].tolist(),
                    "rel_idex": m[:, 12].astype(np.int64).tolist()})
    return out
