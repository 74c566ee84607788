"""Temporal association of per-frame triplets into video relations -- TEST INFRASTRUCTURE ONLY.

`greedy_relational_association` / `association` restate lib/utils.py:134-182 and :461-526 (with `VideoRelation`,
:37-98, and `_iou`, :20-32).  `reference_functions()` executes the reference's own definitions (extracted from the source
file, because importing lib/utils.py needs scipy .mat files and two JSON files at absolute paths, :34-35) and is what
`tests/golden/make_assoc_golden.py` and the `needs_reference` test pin this restatement with.

A frame's predictions are `[fno, [[conf, [s_cid, pid, o_cid], [sub_box, obj_box], rel_idx], ...]]`
(test_net_SGG_emb.py:209); relations come back as dicts with `triplet` (class ids here; the reference maps them through
its name lists), `score`, `duration`, `sub_traj`, `obj_traj`, `rel_idex`.
"""
from __future__ import annotations

import ast
import os

import numpy as np

REF_UTILS = "/root/reference/lib/utils.py"


def reference_functions(objects=None, predicates=None):
    """{name: object} of _iou, VideoRelation, greedy_relational_association, association, executed from the reference."""
    src = open(REF_UTILS).read()
    tree = ast.parse(src)
    keep = {"_iou", "VideoRelation", "greedy_relational_association", "association"}
    nodes = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in keep]
    ident = _Identity()
    env = {"np": np, "objects_list": objects if objects is not None else ident,
           "predicates_list": predicates if predicates is not None else ident, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), REF_UTILS, "exec"), env)
    return {k: env[k] for k in keep}


class _Identity:
    """objects_list / predicates_list stand-in: maps a class id to itself."""

    def __getitem__(self, i):
        return int(i)


def _iou(a, b):
    """lib/utils.py:20-32 (no +1, strict emptiness test), float64 like the reference's Python floats."""
    left, right = max(a[0], b[0]), min(a[2], b[2])
    up, down = max(a[1], b[1]), min(a[3], b[3])
    if left >= right or down <= up:
        return 0
    s1 = (a[2] - a[0]) * (a[3] - a[1])
    s2 = (b[2] - b[0]) * (b[3] - b[1])
    sc = (down - up) * (right - left)
    return sc / (s1 + s2 - sc)


class _Rel:
    __slots__ = ("trip", "straj", "otraj", "confs", "idex", "fstart", "fend")

    def __init__(self, trip, sbox, obox, fstart, conf, idex):
        self.trip, self.straj, self.otraj = [int(t) for t in trip], [sbox], [obox]
        self.confs, self.idex, self.fstart, self.fend = [conf], [idex], fstart, fstart + 1

    def mean(self):
        return np.mean(self.confs)        # numpy's pairwise float64 sum, as lib/utils.py:65-66


def greedy_relational_association(frame_relations, max_traj_num_in_clip=100, min_len=10):
    """lib/utils.py:134-182."""
    frame_relations = sorted(frame_relations, key=lambda x: int(x[0]))
    rels, last = [], []
    for i, (index, preds) in enumerate(frame_relations):
        preds = sorted(preds, key=lambda x: x[0], reverse=True)[:max_traj_num_in_clip]
        cur = []
        for pred in preds:
            conf, trip, (sbox, obox), idex = pred[0], pred[1], pred[2], pred[3]
            merged = False
            if i > 0:
                last.sort(key=lambda r: r.mean(), reverse=True)          # :159, stable
                for r in last:
                    if [int(t) for t in trip] == r.trip or list(trip) == r.trip:
                        if index == r.fend and _iou(r.straj[-1], sbox) >= 0.5 and _iou(r.otraj[-1], obox) >= 0.5:
                            r.straj.append(sbox)
                            r.otraj.append(obox)
                            r.confs.append(conf)
                            r.idex.append(idex)
                            r.fend += 1
                            last.remove(r)
                            cur.append(r)
                            merged = True
                            break
            if not merged:
                r = _Rel(trip, sbox, obox, index, conf, idex)
                rels.append(r)
                cur.append(r)
        last = cur
    return [{"triplet": r.trip, "score": float(r.mean()), "duration": [int(r.fstart), int(r.fend)],
             "sub_traj": r.straj, "obj_traj": r.otraj, "rel_idex": r.idex} for r in rels if len(r.straj) >= min_len]


def fill_empty_frames(pred, invalid_num=4):
    """lib/utils.py:470-518: frames without predictions borrow the nearest non-empty neighbour's, unless every frame in a
    +-4 window is empty.  Returns the source frame position per frame (-1: keep, -2: leave empty)."""
    mask = [0 if len(p[1]) == 0 else -1 for p in pred]
    n = len(pred)
    tmp = [-1] * n
    for i in range(n):
        if mask[i] == 0:
            j = i - 1
            while j >= 0 and mask[j] == 0:
                j -= 1
            left = 0 if j < 0 else i - j
            j = i + 1
            while j < n and mask[j] == 0:
                j += 1
            right = 0 if j >= n else j - i
            if right == 0 or (left > 0 and left <= right):
                tmp[i] = i - left
            elif left == 0 or (right > 0 and left > right):
                tmp[i] = i + right
    mask = tmp
    for i in range(n):
        if mask[i] >= 0:
            if i < invalid_num:
                start, end = 0, i + invalid_num
            elif i > n - invalid_num - 1:
                start, end = i - invalid_num, n - 1
            else:
                start, end = i - invalid_num, i + invalid_num
            if all(mask[j] != -1 for j in range(start, end + 1)):
                mask[i] = -2
    return mask


def association(frame_relations, max_num_per_video=200):
    """lib/utils.py:461-526 on {vid: [[fno, preds], ...]}."""
    out = {}
    for vid, pred in frame_relations.items():
        pred = sorted(pred, key=lambda x: int(x[0]))
        if all(len(p[1]) == 0 for p in pred):
            continue
        mask = fill_empty_frames(pred)
        pred = [[p[0], p[1]] for p in pred]
        for i, m in enumerate(mask):
            if m > -1:
                pred[i][1] = pred[m][1]
        rel = greedy_relational_association(pred)
        rel.sort(key=lambda x: x["score"], reverse=True)
        out[vid] = rel[:max_num_per_video]
    return out
