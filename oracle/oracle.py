"""Python face of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

ctypes bindings to ``oracle/liboracle.so`` (built from ``oracle/oracle.c`` by
``make -C oracle``) plus the numpy restatements of the host-side Python parts of the
path.  Every function cites the reference lines it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.c_float
_i = ctypes.c_int
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    """Compile oracle.c (and, if the reference tree is mounted, oracle/_ref)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/lib/model"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, ty=_fp):
    return a.ctypes.data_as(ty)


def default_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- RoIAlign (lattice)
def roi_align_forward(feat, rois, gh: int, gw: int, scale: float, nthreads: int = 1) -> np.ndarray:
    """roi_align.c:80-136 / roi_align_kernel.cu:15-70 -> [N,C,gh,gw]."""
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    _, C, H, W = feat.shape
    N = rois.shape[0]
    out = np.empty((N, C, gh, gw), np.float32)
    lib().orc_roi_align_forward(_ptr(feat), _f(scale), _i(N), _i(H), _i(W), _i(C), _i(gh), _i(gw),
                                _ptr(rois), _ptr(out), _i(nthreads))
    return out


def roi_align_backward(top_diff, rois, feat_shape, gh: int, gw: int, scale: float, nthreads: int = 1) -> np.ndarray:
    """roi_align_kernel.cu:94-143 -> [B,C,H,W] (exact sum of the float contributions)."""
    top_diff, rois = _f32(top_diff), _f32(rois).reshape(-1, 5)
    B, C, H, W = feat_shape
    N = rois.shape[0]
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_align_backward(_ptr(top_diff), _f(scale), _i(B), _i(N), _i(H), _i(W), _i(C), _i(gh), _i(gw),
                                 _ptr(rois), _ptr(out), _i(nthreads))
    return out


def pool2x2_forward(x, mode: str) -> np.ndarray:
    """avg_pool2d / max_pool2d(kernel 2, stride 1) of modules/roi_align.py:29,42."""
    x = _f32(x)
    N, C, gh, gw = x.shape
    out = np.empty((N, C, gh - 1, gw - 1), np.float32)
    lib().orc_pool2x2_forward(_ptr(x), ctypes.c_long(N * C), _i(gh), _i(gw), _i(1 if mode == "avg" else 2), _ptr(out))
    return out


def pool2x2_backward(grad_out, x, mode: str) -> np.ndarray:
    grad_out, x = _f32(grad_out), _f32(x)
    N, C, gh, gw = x.shape
    out = np.empty_like(x)
    lib().orc_pool2x2_backward(_ptr(grad_out), _ptr(x), ctypes.c_long(N * C), _i(gh), _i(gw),
                               _i(1 if mode == "avg" else 2), _ptr(out))
    return out


def roi_align_pooled_forward(feat, rois, ph: int, pw: int, scale: float, mode: str, nthreads: int = 1) -> np.ndarray:
    """RoIAlignAvg / RoIAlignMax (modules/roi_align.py:18-42): lattice (ph+1)x(pw+1) then 2x2/1 pool;
    mode 'none' is plain RoIAlign (modules/roi_align.py:6-16)."""
    if mode == "none":
        return roi_align_forward(feat, rois, ph, pw, scale, nthreads)
    return pool2x2_forward(roi_align_forward(feat, rois, ph + 1, pw + 1, scale, nthreads), mode)


def roi_align_pooled_backward(grad_out, feat, rois, ph: int, pw: int, scale: float, mode: str, nthreads: int = 1):
    """Autograd of the composition above: pool backward, then roi_align_kernel.cu:94-143."""
    feat = _f32(feat)
    if mode == "none":
        return roi_align_backward(grad_out, rois, feat.shape, ph, pw, scale, nthreads)
    lattice = roi_align_forward(feat, rois, ph + 1, pw + 1, scale, nthreads) if mode == "max" else \
        np.zeros((np.asarray(rois).reshape(-1, 5).shape[0], feat.shape[1], ph + 1, pw + 1), np.float32)
    g_lat = pool2x2_backward(grad_out, lattice, mode)
    return roi_align_backward(g_lat, rois, feat.shape, ph + 1, pw + 1, scale, nthreads)


# ----------------------------------------------------------------------------- RoIPool (cffi era)
def roi_pool_forward(feat, rois, ph: int, pw: int, scale: float, nthreads: int = 1):
    """roi_pooling_kernel.cu:24-93 -> (out [N,C,ph,pw], argmax int32 flat index into feat)."""
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    _, C, H, W = feat.shape
    N = rois.shape[0]
    out = np.empty((N, C, ph, pw), np.float32)
    arg = np.empty((N, C, ph, pw), np.int32)
    lib().orc_roi_pool_forward(_ptr(feat), _f(scale), _i(N), _i(H), _i(W), _i(C), _i(ph), _i(pw), _ptr(rois),
                               _ptr(out), _ptr(arg, _ip), _i(nthreads))
    return out, arg


def roi_pool_backward(top_diff, rois, argmax, feat_shape, ph: int, pw: int, scale: float) -> np.ndarray:
    """roi_pooling_kernel.cu:128-203."""
    top_diff, rois = _f32(top_diff), _f32(rois).reshape(-1, 5)
    argmax = np.ascontiguousarray(argmax, np.int32)
    B, C, H, W = feat_shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_pool_backward(_ptr(top_diff), _f(scale), _i(B), _i(rois.shape[0]), _i(H), _i(W), _i(C), _i(ph),
                                _i(pw), _ptr(rois), _ptr(out), _ptr(argmax, _ip))
    return out


# ----------------------------------------------------------------------------- model._C stand-ins
def c_roi_align_forward(feat, rois, ph: int, pw: int, scale: float, sampling_ratio: int, nthreads: int = 1):
    """The op behind roi_layers/roi_align.py:20 (maskrcnn-benchmark RoIAlign, == torchvision aligned=False)."""
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    _, C, H, W = feat.shape
    out = np.empty((rois.shape[0], C, ph, pw), np.float32)
    lib().orc_c_roi_align_forward(_ptr(feat), _f(scale), _i(rois.shape[0]), _i(H), _i(W), _i(C), _i(ph), _i(pw),
                                  _i(sampling_ratio), _ptr(rois), _ptr(out), _i(nthreads))
    return out


def c_roi_align_backward(grad_out, rois, feat_shape, ph: int, pw: int, scale: float, sampling_ratio: int):
    grad_out, rois = _f32(grad_out), _f32(rois).reshape(-1, 5)
    B, C, H, W = feat_shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_c_roi_align_backward(_ptr(grad_out), _f(scale), _i(B), _i(rois.shape[0]), _i(H), _i(W), _i(C), _i(ph),
                                   _i(pw), _i(sampling_ratio), _ptr(rois), _ptr(out))
    return out


def c_roi_pool_forward(feat, rois, ph: int, pw: int, scale: float, nthreads: int = 1):
    """The op behind roi_layers/roi_pool.py:17-19 (per-plane argmax h*W+w, -1 when empty)."""
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    _, C, H, W = feat.shape
    out = np.empty((rois.shape[0], C, ph, pw), np.float32)
    arg = np.empty((rois.shape[0], C, ph, pw), np.int32)
    lib().orc_c_roi_pool_forward(_ptr(feat), _f(scale), _i(rois.shape[0]), _i(H), _i(W), _i(C), _i(ph), _i(pw),
                                 _ptr(rois), _ptr(out), _ptr(arg, _ip), _i(nthreads))
    return out, arg


def c_roi_pool_backward(grad_out, rois, argmax, feat_shape, ph: int, pw: int):
    grad_out, rois = _f32(grad_out), _f32(rois).reshape(-1, 5)
    argmax = np.ascontiguousarray(argmax, np.int32)
    B, C, H, W = feat_shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_c_roi_pool_backward(_ptr(grad_out), _ptr(argmax, _ip), _i(B), _i(rois.shape[0]), _i(H), _i(W), _i(C),
                                  _i(ph), _i(pw), _ptr(rois), _ptr(out))
    return out


# ----------------------------------------------------------------------------- roi_crop
def roi_crop_forward(feat, grids) -> np.ndarray:
    """RoICropFunction.forward (functions/roi_crop.py:8-16 -> roi_crop_cuda_kernel.cu:47-108): feat [B,C,H,W], grids
    [N,oh,ow,2] = (y, x) in [-1,1] -> [N,C,oh,ow]."""
    feat, grids = _f32(feat), _f32(grids)
    B, C, H, W = feat.shape
    N, oh, ow, _ = grids.shape
    out = np.empty((N, C, oh, ow), np.float32)
    lib().orc_roi_crop_forward(_ptr(feat), _ptr(grids), _i(B), _i(C), _i(H), _i(W), _i(N), _i(oh), _i(ow), _ptr(out))
    return out


def roi_crop_backward(grad_out, grids, feat_shape) -> np.ndarray:
    """RoICropFunction.backward (functions/roi_crop.py:18-24 -> roi_crop_cuda_kernel.cu:111-195): the gradient of the
    features; the reference leaves the gradient of the grids at zero."""
    grad_out, grids = _f32(grad_out), _f32(grids)
    B, C, H, W = feat_shape
    N, oh, ow, _ = grids.shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_crop_backward(_ptr(grad_out), _ptr(grids), _i(B), _i(C), _i(H), _i(W), _i(N), _i(oh), _i(ow), _ptr(out))
    return out


# ----------------------------------------------------------------------------- NMS
def nms_sorted(boxes, thresh: float, max_keep: int = 0) -> np.ndarray:
    """Greedy NMS over rows already in descending-score order (nms_cpu.py:14-32 arithmetic)."""
    boxes = _f32(boxes)
    n, stride = boxes.shape
    keep = np.empty(max(n, 1), np.int32)
    k = lib().orc_nms_sorted(_ptr(boxes), _i(n), _i(stride), _f(np.float32(thresh)), _i(max_keep), _ptr(keep, _ip))
    return keep[:k].copy()


def nms(dets, thresh: float) -> np.ndarray:
    """nms_wrapper.py:13-21 -> nms_cpu.py:6-34: sort by column 4 descending, greedy, original indices."""
    dets = _f32(dets)
    if dets.shape[0] == 0:
        return np.empty(0, np.int32)
    order = np.argsort(-dets[:, 4], kind="stable")
    kept = nms_sorted(dets[order], thresh)
    return order[kept].astype(np.int32)


# ----------------------------------------------------------------------------- proposal layer
def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)) -> np.ndarray:
    """generate_anchors.py:45-105 (ratio enumeration around the 16x16 window, then scale enumeration)."""
    ratios = np.asarray(ratios, np.float64)
    scales = np.asarray(scales, np.float64)
    w = h = float(base_size)
    cx = cy = 0.5 * (base_size - 1)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    rows = []
    for rw, rh in zip(ws, hs):
        for s in scales:
            sw, sh = rw * s, rh * s
            rows.append([cx - 0.5 * (sw - 1), cy - 0.5 * (sh - 1), cx + 0.5 * (sw - 1), cy + 0.5 * (sh - 1)])
    return np.asarray(rows, np.float64)


def rpn_cls_prob(cls_score) -> np.ndarray:
    """rpn.py:63-69: rpn_cls_score [B,2A,H,W] viewed as [B,2,A*H,W], softmax over dim 1, viewed back: channel a is an
    anchor's background score, channel a + A its foreground score.  fp32 max / subtract / add / divide, exponentials
    correctly rounded (double exp rounded once) -- the arithmetic the CUDA kernel performs, so it matches bit for bit;
    torch's own softmax (vectorised expf) differs from it by at most a few ulp (tests/golden/rpn_golden.npz)."""
    s = _f32(cls_score)
    A = s.shape[1] // 2
    bg, fg = s[:, :A], s[:, A:]
    m = np.maximum(bg, fg)
    e_bg = np.exp((bg - m).astype(np.float64)).astype(np.float32)
    e_fg = np.exp((fg - m).astype(np.float64)).astype(np.float32)
    tot = e_bg + e_fg
    return np.concatenate([e_bg / tot, e_fg / tot], axis=1).astype(np.float32)


def proposal_decode(cls_prob, bbox_pred, im_info, base_anchors, feat_stride: int = 16):
    """proposal_layer.py:67,81-111 + bbox_transform.py:77-103,125-133 -> (boxes [B,KA,4], scores [B,KA])."""
    cls_prob, bbox_pred, im_info = _f32(cls_prob), _f32(bbox_pred), _f32(im_info)
    base = _f32(base_anchors)
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    boxes = np.empty((B, H * W * A, 4), np.float32)
    scores = np.empty((B, H * W * A), np.float32)
    lib().orc_proposal_decode(_ptr(cls_prob), _ptr(bbox_pred), _ptr(im_info), _ptr(base), _i(B), _i(A), _i(H), _i(W),
                              _i(feat_stride), _ptr(boxes), _ptr(scores))
    return boxes, scores


def proposal_layer(cls_prob, bbox_pred, im_info, pre_nms_top_n: int, post_nms_top_n: int, nms_thresh: float,
                   feat_stride: int = 16, scales=(8, 16, 32), ratios=(0.5, 1, 2)) -> np.ndarray:
    """_ProposalLayer.forward (proposal_layer.py:49-163) -> [B, post_nms_top_n, 5]."""
    base = generate_anchors(scales=scales, ratios=ratios).astype(np.float32)
    boxes, scores = proposal_decode(cls_prob, bbox_pred, im_info, base, feat_stride)
    B = boxes.shape[0]
    out = np.zeros((B, post_nms_top_n, 5), np.float32)
    for b in range(B):
        order = np.argsort(-scores[b], kind="stable")                       # :127
        if 0 < pre_nms_top_n < scores.size:                                  # :140 (batch-wide numel)
            order = order[:pre_nms_top_n]
        cand = boxes[b][order]
        keep = nms_sorted(cand, nms_thresh, post_nms_top_n if post_nms_top_n > 0 else 0)   # :150-154
        out[b, :, 0] = b                                                    # :160
        out[b, :keep.size, 1:] = cand[keep]                                 # :161
    return out


# ----------------------------------------------------------------------------- SGG pair stage
def enumerate_pairs(n: int):
    """faster_rcnn_SGG_emb.py:597-606: all ordered (i, j), i != j, i-major."""
    ixs = np.repeat(np.arange(n, dtype=np.int64), n - 1) if n > 1 else np.empty(0, np.int64)
    j = np.tile(np.arange(n - 1, dtype=np.int64), n) if n > 1 else np.empty(0, np.int64)
    ixo = j + (j >= ixs)
    return ixs, ixo


def union_boxes(boxes, ixs, ixo, ih: float, iw: float, margin: float = 10.0) -> np.ndarray:
    """resnet_SGG_emb.py:240-244 applied per pair as in faster_rcnn_SGG_emb.py:649-653 -> rel_boxes [P,5]."""
    b = np.asarray(boxes, np.float64)
    s, o = b[ixs], b[ixo]
    rel = np.zeros((len(ixs), 5), np.float64)
    rel[:, 1] = np.maximum(0, np.minimum(s[:, 0], o[:, 0]) - margin)
    rel[:, 2] = np.maximum(0, np.minimum(s[:, 1], o[:, 1]) - margin)
    rel[:, 3] = np.minimum(iw, np.maximum(s[:, 2], o[:, 2]) + margin)
    rel[:, 4] = np.minimum(ih, np.maximum(s[:, 3], o[:, 3]) + margin)
    return rel.astype(np.float32)                                            # resnet_SGG_emb.py:136


def dual_mask_extent(box, ih: float, iw: float):
    """resnet_SGG_emb.py:246-252: the (x1, x2, y1, y2) cell extents of one 32x32 mask (fp64 like numpy)."""
    rh, rw = 32.0 / ih, 32.0 / iw
    x1 = max(0, int(math.floor(float(box[0]) * rw)))
    x2 = min(32, int(math.ceil(float(box[2]) * rw)))
    y1 = max(0, int(math.floor(float(box[1]) * rh)))
    y2 = min(32, int(math.ceil(float(box[3]) * rh)))
    return x1, x2, y1, y2


def dual_masks(boxes, ixs, ixo, ih: float, iw: float) -> np.ndarray:
    """faster_rcnn_SGG_emb.py:654-655 -> SpatialFea [P,2,32,32] float32."""
    boxes = np.asarray(boxes)
    per_box = np.zeros((boxes.shape[0], 32, 32), np.float32)
    for k in range(boxes.shape[0]):
        x1, x2, y1, y2 = dual_mask_extent(boxes[k], ih, iw)
        per_box[k, y1:y2, x1:x2] = 1
    return np.stack([per_box[ixs], per_box[ixo]], axis=1) if len(ixs) else np.zeros((0, 2, 32, 32), np.float32)


def detection_output(rel_score, confs, classes, boxes, ixs, ixo, top: int = 100):
    """lib/utils.py:609-626: scale by both confidences, global argsort descending, first `top`.
    Returns (conf[K], labels[K,3], sub_boxes[K,4], obj_boxes[K,4], pair_idx[K]); ties -> lower flat index."""
    prob = np.array(rel_score, np.float32, copy=True)
    confs = np.asarray(confs, np.float32)
    for i in range(prob.shape[0]):
        prob[i] = prob[i] * confs[ixs[i]] * confs[ixo[i]]
    flat = np.argsort(-prob.ravel(), kind="stable")[:top]
    pair, rel = np.unravel_index(flat, prob.shape)
    boxes = np.asarray(boxes, np.float32)
    classes = np.asarray(classes)
    labels = np.stack([classes[ixs[pair]], rel, classes[ixo[pair]]], axis=1).astype(np.float32)
    return prob[pair, rel], labels, boxes[ixs[pair]], boxes[ixo[pair]], pair.astype(np.int64)


# ------------------------------------------------------------------ SGG projection (resnet_SGG_emb.py:128-221)
def _fc(x, params, name, relu=True):
    """FC of lib/model/faster_rcnn/utils.py:48-60: nn.Linear then optional ReLU, fp32."""
    y = x.astype(np.float32) @ params[name + ".weight"].T + params[name + ".bias"]
    return np.maximum(y, 0, out=y) if relu else y


def _conv2d(x, w, b, stride: int, pad: int, relu=True):
    """nn.Conv2d + ReLU of lib/model/faster_rcnn/utils.py:32-46 (bn=False), NCHW fp32, by explicit patches."""
    n, c, h, wd = x.shape
    o, _, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    oh, ow = (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
    s = xp.strides
    patches = np.lib.stride_tricks.as_strided(xp, (n, oh, ow, c, kh, kw),
                                              (s[0], s[2] * stride, s[3] * stride, s[1], s[2], s[3]))
    y = patches.reshape(n * oh * ow, c * kh * kw).astype(np.float32) @ w.reshape(o, -1).T + b
    y = y.reshape(n, oh, ow, o).transpose(0, 3, 1, 2)
    return np.maximum(y, 0) if relu else y


def _normalize(x):
    """F.normalize(p=2, dim=1): x / max(||x||, 1e-12)."""
    return x / np.maximum(np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True)), 1e-12).astype(np.float32)


def vrd_forward(params: dict, prd_vecs, fmap, boxes, rel_boxes, spatial, ix1, ix2, use_obj_visual=True,
                spatial_type=2, pool: int = 7, rows=None, nthreads: int = 1, dropout_masks=None, dropout_p: float = 0.5):
    """vrd.forward (resnet_SGG_emb.py:128-221) restated in numpy fp32 -> (scores [P,n_rel], feat [P,emb]).

    roi_pool is the model._C flavour (`c_roi_pool_forward`).  Eval mode by default: dropout is the identity (:148-149)
    and the scores are softmaxed (:215-219).  `dropout_masks` = the four keep masks of the F.dropout calls at :148, :149,
    :162, :163 ([N,h], [N,h], [P,h], [P,h]) switches to TRAINING mode: y = keep * y / (1 - p) behind fc6 / fc7 of both
    branches and raw cosine similarities as scores.
    `rows` restricts the computation to a subset of pair indices (full-size spot checks)."""
    fmap = _f32(fmap)
    boxes = _f32(boxes).reshape(-1, 5)
    rel_boxes = _f32(rel_boxes).reshape(-1, 5)
    ix1, ix2 = np.asarray(ix1, np.int64), np.asarray(ix2, np.int64)
    spatial = _f32(spatial)
    if rows is not None:
        rows = np.asarray(rows, np.int64)
        rel_boxes, ix1, ix2, spatial = rel_boxes[rows], ix1[rows], ix2[rows], spatial[rows]
    training = dropout_masks is not None
    keep = [np.asarray(m, np.float32) * np.float32(1.0 / (1.0 - dropout_p)) for m in dropout_masks] if training else None
    drop = (lambda y, k: y * keep[k]) if training else (lambda y, k: y)
    x_so, _ = c_roi_pool_forward(fmap, boxes, pool, pool, 1.0 / 16, nthreads)                    # :144
    x_so = drop(_fc(drop(_fc(x_so.reshape(len(boxes), -1), params, "fc6.fc"), 0), params, "fc7.fc"), 1)   # :146-149
    obj = _fc(x_so, params, "so_vis_embeddings.fc", relu=False)                                 # :150
    x_u, _ = c_roi_pool_forward(fmap, rel_boxes, pool, pool, 1.0 / 16, nthreads)                 # :158
    x = drop(_fc(drop(_fc(x_u.reshape(len(rel_boxes), -1), params, "fc6.fc"), 2), params, "fc7.fc"), 3)   # :160-163
    x = _fc(x, params, "fc8.fc")                                                                # :164
    parts = [x]
    if use_obj_visual:                                                                          # :166-170
        parts.append(_fc(np.concatenate([obj[ix1], obj[ix2]], 1), params, "fc_so.fc"))
    if spatial_type == 1:                                                                       # :172-174
        parts.append(_fc(spatial.reshape(len(rel_boxes), 8), params, "fc_lov.fc"))
    elif spatial_type == 2:                                                                     # :175-179
        lo = _conv2d(spatial.reshape(-1, 2, 32, 32), params["conv_lo.0.conv.weight"], params["conv_lo.0.conv.bias"], 2, 2)
        lo = _conv2d(lo, params["conv_lo.1.conv.weight"], params["conv_lo.1.conv.bias"], 2, 2)
        lo = _conv2d(lo, params["conv_lo.2.conv.weight"], params["conv_lo.2.conv.bias"], 1, 0)
        parts.append(_fc(lo.reshape(len(rel_boxes), -1), params, "fc_lov.fc"))
    x = _fc(_fc(np.concatenate(parts, 1), params, "fc_fusion.fc"), params, "fc_rel.fc", relu=False)  # :190-191
    prd = _f32(prd_vecs) @ params["prd_sem_embeddings.0.weight"].T + params["prd_sem_embeddings.0.bias"]   # :203-206
    prd = np.where(prd > 0, prd, np.float32(0.1) * prd)
    prd = prd @ params["prd_sem_embeddings.2.weight"].T + params["prd_sem_embeddings.2.bias"]
    sim = (_normalize(x) @ _normalize(prd).T).astype(np.float64)                                # :207-211
    if training:                                                                                # :215: no softmax
        return sim.astype(np.float32), x
    e = np.exp(sim - sim.max(1, keepdims=True))                                                 # :216-219 (eval)
    return (e / e.sum(1, keepdims=True)).astype(np.float32), x
