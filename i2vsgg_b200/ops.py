"""Tensor-level entry points over the C ABI: every function takes CUDA tensors, runs on torch's current
stream and returns new tensors.  torch is used for device memory and streams only; all arithmetic happens in
``libi2vsgg_b200.so``.  CPU tensors are rejected -- there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import (ARGMAX_FLAT, ARGMAX_PLANE, DT_BF16, DT_F32, DT_TF32, IMPL_AUTO, IMPL_BAND, IMPL_GATHER, IMPL_PHASE, IMPL_PLANE,
                   IMPL_CHAN, IMPL_EVEN, IMPL_ROWS, IMPL_SLAB,
                   POOL_AVG, POOL_MAX, POOL_NONE, check, load)

_POOLS = {"none": POOL_NONE, "avg": POOL_AVG, "max": POOL_MAX, POOL_NONE: POOL_NONE, POOL_AVG: POOL_AVG,
          POOL_MAX: POOL_MAX}
_IMPLS = {"auto": IMPL_AUTO, "gather": IMPL_GATHER, "plane": IMPL_PLANE, "rows": IMPL_ROWS, "phase": IMPL_PHASE, "band": IMPL_BAND, "slab": IMPL_SLAB, "even": IMPL_EVEN, "chan": IMPL_CHAN,
          IMPL_AUTO: IMPL_AUTO, IMPL_GATHER: IMPL_GATHER, IMPL_PLANE: IMPL_PLANE, IMPL_ROWS: IMPL_ROWS, IMPL_PHASE: IMPL_PHASE,
          IMPL_BAND: IMPL_BAND, IMPL_SLAB: IMPL_SLAB, IMPL_EVEN: IMPL_EVEN, IMPL_CHAN: IMPL_CHAN}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.I2VError(f"{what}: expected a CUDA tensor (i2vsgg_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _rois5(rois: torch.Tensor) -> torch.Tensor:
    rois = _f32(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 5:
        # roi_align_cuda.c:19-22 / roi_pooling_cuda.c:20-23 return 0 here; a drop-in that silently did nothing
        # would hide the bug, so this raises instead
        raise _lib.I2VError(f"rois must be [N,5], got {tuple(rois.shape)}")
    return rois


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------- lattice RoIAlign
def roi_align_forward(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float, pool="none", impl="auto"):
    """RoIAlign / RoIAlignAvg / RoIAlignMax forward (modules/roi_align.py:6-42) -> [N,C,pooled_h,pooled_w]."""
    features, rois = _f32(features, "features"), _rois5(rois)
    B, C, H, W = features.shape
    N = rois.size(0)
    out = torch.empty((N, C, pooled_h, pooled_w), dtype=torch.float32, device=features.device)
    lib = load()
    with torch.cuda.device(features.device):
        nb = lib.i2v_roi_align_workspace_bytes(B, N)
        ws = _workspace(nb, features.device)
        check(lib.i2v_roi_align_forward(_p(features), _p(rois), _p(out), B, C, H, W, N, pooled_h, pooled_w,
                                        float(spatial_scale), _POOLS[pool], _IMPLS[impl], _p(ws), ws.numel(),
                                        _stream()), "i2v_roi_align_forward")
    return out


def roi_align_backward(grad_out, features, rois, feat_shape, pooled_h: int, pooled_w: int, spatial_scale: float,
                       pool="none", impl="auto"):
    """Gradient of the above w.r.t. the features -> [B,C,H,W].  `features` is only read for pool='max'."""
    grad_out, rois = _f32(grad_out, "grad_out"), _rois5(rois)
    B, C, H, W = feat_shape
    N = rois.size(0)
    if _POOLS[pool] == POOL_MAX:
        features = _f32(features, "features")
    else:
        features = None
    grad_in = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
    lib = load()
    with torch.cuda.device(grad_out.device):
        ws = _workspace(lib.i2v_roi_align_workspace_bytes(B, N), grad_out.device)
        check(lib.i2v_roi_align_backward(_p(grad_out), _p(features), _p(rois), _p(grad_in), B, C, H, W, N, pooled_h,
                                         pooled_w, float(spatial_scale), _POOLS[pool], _IMPLS[impl], _p(ws),
                                         ws.numel(), _stream()), "i2v_roi_align_backward")
    return grad_in


# --------------------------------------------------------------------------------------- RoIPool
def roi_pool_forward(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float, argmax_mode=ARGMAX_FLAT):
    features, rois = _f32(features, "features"), _rois5(rois)
    B, C, H, W = features.shape
    N = rois.size(0)
    out = torch.empty((N, C, pooled_h, pooled_w), dtype=torch.float32, device=features.device)
    argmax = torch.empty((N, C, pooled_h, pooled_w), dtype=torch.int32, device=features.device)
    with torch.cuda.device(features.device):
        check(load().i2v_roi_pool_forward(_p(features), _p(rois), _p(out), _p(argmax), B, C, H, W, N, pooled_h,
                                          pooled_w, float(spatial_scale), argmax_mode, _stream()),
              "i2v_roi_pool_forward")
    return out, argmax


def roi_pool_backward(grad_out, rois, argmax, feat_shape, pooled_h: int, pooled_w: int, spatial_scale: float,
                      argmax_mode=ARGMAX_FLAT):
    grad_out, rois = _f32(grad_out, "grad_out"), _rois5(rois)
    argmax = argmax.contiguous()
    assert argmax.dtype == torch.int32 and argmax.is_cuda
    B, C, H, W = feat_shape
    grad_in = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(load().i2v_roi_pool_backward(_p(grad_out), _p(rois), _p(argmax), _p(grad_in), B, C, H, W, rois.size(0),
                                           pooled_h, pooled_w, float(spatial_scale), argmax_mode, _stream()),
              "i2v_roi_pool_backward")
    return grad_in


# --------------------------------------------------------------------------------------- model._C RoIAlign
def c_roi_align_forward(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float, sampling_ratio: int):
    features, rois = _f32(features, "features"), _rois5(rois)
    B, C, H, W = features.shape
    out = torch.empty((rois.size(0), C, pooled_h, pooled_w), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(load().i2v_c_roi_align_forward(_p(features), _p(rois), _p(out), B, C, H, W, rois.size(0), pooled_h,
                                             pooled_w, float(spatial_scale), int(sampling_ratio), _stream()),
              "i2v_c_roi_align_forward")
    return out


def c_roi_align_backward(grad_out, rois, feat_shape, pooled_h: int, pooled_w: int, spatial_scale: float,
                         sampling_ratio: int):
    grad_out, rois = _f32(grad_out, "grad_out"), _rois5(rois)
    B, C, H, W = feat_shape
    grad_in = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(load().i2v_c_roi_align_backward(_p(grad_out), _p(rois), _p(grad_in), B, C, H, W, rois.size(0), pooled_h,
                                              pooled_w, float(spatial_scale), int(sampling_ratio), _stream()),
              "i2v_c_roi_align_backward")
    return grad_in


# --------------------------------------------------------------------------------------- NMS
def nms_sorted(boxes, thresh: float, max_keep: int = 0):
    """boxes [N,>=4] or [B,N,>=4], rows already in descending score order -> (keep [.., K] int32, counts)."""
    boxes = _f32(boxes, "boxes")
    single = boxes.dim() == 2
    if single:
        boxes = boxes.unsqueeze(0)
    B, N, S = boxes.shape
    kcap = min(N, max_keep) if max_keep > 0 else N
    keep = torch.zeros((B, max(kcap, 1)), dtype=torch.int32, device=boxes.device)
    num = torch.zeros((B,), dtype=torch.int32, device=boxes.device)
    lib = load()
    with torch.cuda.device(boxes.device):
        ws = _workspace(lib.i2v_nms_workspace_bytes(B, N), boxes.device)
        check(lib.i2v_nms_sorted(_p(boxes), B, N, S, float(thresh), int(max_keep), _p(keep), keep.size(1), _p(num),
                                 _p(ws), ws.numel(), _stream()), "i2v_nms_sorted")
    if single:
        return keep[0, : int(num[0].item())], num
    return keep, num


def nms_dets(dets, thresh: float):
    """nms_wrapper.nms(): dets [N,5] (x1,y1,x2,y2,score) in any order -> kept original row indices, int32."""
    dets = _f32(dets, "dets")
    if dets.dim() != 2 or dets.size(1) != 5:
        raise _lib.I2VError(f"dets must be [N,5], got {tuple(dets.shape)}")
    N = dets.size(0)
    keep = torch.empty((max(N, 1),), dtype=torch.int32, device=dets.device)
    num = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    lib = load()
    with torch.cuda.device(dets.device):
        ws = _workspace(lib.i2v_nms_dets_workspace_bytes(N), dets.device)
        check(lib.i2v_nms_dets(_p(dets), N, float(thresh), _p(keep), _p(num), _p(ws), ws.numel(), _stream()),
              "i2v_nms_dets")
    return keep[: int(num.item())]


# --------------------------------------------------------------------------------------- proposal layer
def rpn_cls_prob(cls_score):
    """rpn.py:66-68: rpn_cls_prob [B,2A,H,W] = softmax over each anchor's (background, foreground) channel pair."""
    cls_score = _f32(cls_score, "cls_score")
    B, A2, H, W = cls_score.shape
    if A2 % 2:
        raise _lib.I2VError("rpn_cls_prob: the channel count must be even (2 per anchor)")
    out = torch.empty_like(cls_score)
    lib = load()
    with torch.cuda.device(cls_score.device):
        check(lib.i2v_rpn_cls_prob(_p(cls_score), _p(out), B, A2 // 2, H, W, _stream()), "i2v_rpn_cls_prob")
    return out


def proposal_forward(cls_prob, bbox_pred, im_info, base_anchors, feat_stride: int, pre_nms_top_n: int,
                     post_nms_top_n: int, nms_thresh: float, return_counts: bool = False, from_scores: bool = False):
    """_ProposalLayer.forward (proposal_layer.py:49-163) -> rois [B, post_nms_top_n, 5].  With `from_scores` the first
    argument is the RPN head's raw rpn_cls_score and the softmax of rpn.py:66-68 runs inside the decode kernel."""
    cls_prob, bbox_pred = _f32(cls_prob, "cls_prob"), _f32(bbox_pred, "bbox_pred")
    im_info, base_anchors = _f32(im_info, "im_info"), _f32(base_anchors, "base_anchors")
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    if bbox_pred.shape != (B, 4 * A, H, W) or base_anchors.shape != (A, 4) or im_info.shape != (B, 3):
        raise _lib.I2VError("proposal_forward: inconsistent shapes")
    out = torch.empty((B, post_nms_top_n, 5), dtype=torch.float32, device=cls_prob.device)
    counts = torch.empty((B,), dtype=torch.int32, device=cls_prob.device)
    lib = load()
    with torch.cuda.device(cls_prob.device):
        ws = _workspace(lib.i2v_proposal_workspace_bytes(B, A, H, W, pre_nms_top_n), cls_prob.device)
        fn = lib.i2v_proposal_forward_scores if from_scores else lib.i2v_proposal_forward
        check(fn(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(base_anchors), B, A, H, W,
                 int(feat_stride), int(pre_nms_top_n), int(post_nms_top_n), float(nms_thresh),
                 _p(out), _p(counts), _p(ws), ws.numel(), _stream()), "i2v_proposal_forward")
    return (out, counts) if return_counts else out


def proposal_stages(cls_prob, bbox_pred, im_info, base_anchors, feat_stride: int, want_order: bool = True):
    """Decoded+clipped boxes [B,KA,4], scores [B,KA] and the descending-score order [B,KA] (tests)."""
    cls_prob, bbox_pred = _f32(cls_prob, "cls_prob"), _f32(bbox_pred, "bbox_pred")
    im_info, base_anchors = _f32(im_info, "im_info"), _f32(base_anchors, "base_anchors")
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    KA = A * H * W
    dev = cls_prob.device
    boxes = torch.empty((B, KA, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((B, KA), dtype=torch.float32, device=dev)
    order = torch.empty((B, KA), dtype=torch.int32, device=dev) if want_order else None
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.i2v_proposal_workspace_bytes(B, A, H, W, 0), dev)
        check(lib.i2v_proposal_stages(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(base_anchors), B, A, H, W,
                                      int(feat_stride), _p(boxes), _p(scores), _p(order), _p(ws), ws.numel(),
                                      _stream()), "i2v_proposal_stages")
    return boxes, scores, order


# --------------------------------------------------------------------------------------- SGG pair stage
def roi_crop_forward(features, grids):
    """RoICropFunction.forward (functions/roi_crop.py:8-16): features [B,C,H,W], grids [N,oh,ow,2] = (y, x) in [-1,1]
    -> [N,C,oh,ow]; RoI n samples frame n // (N // B)."""
    features, grids = _f32(features, "features"), _f32(grids, "grids")
    if features.dim() != 4 or grids.dim() != 4 or grids.size(3) != 2:
        raise _lib.I2VError("roi_crop_forward: features [B,C,H,W] and grids [N,oh,ow,2] expected")
    B, C, H, W = features.shape
    N, oh, ow, _ = grids.shape
    out = torch.empty((N, C, oh, ow), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(load().i2v_roi_crop_forward(_p(features), _p(grids), _p(out), B, C, H, W, N, oh, ow, _stream()),
              "i2v_roi_crop_forward")
    return out


def roi_crop_backward(grad_out, grids, feat_shape):
    """RoICropFunction.backward (functions/roi_crop.py:18-24): the gradient of the features [B,C,H,W] (the gradient of the
    grids is zero in the reference)."""
    grad_out, grids = _f32(grad_out, "grad_out"), _f32(grids, "grids")
    B, C, H, W = (int(v) for v in feat_shape)
    N, oh, ow, _ = grids.shape
    if grad_out.shape != (N, C, oh, ow):
        raise _lib.I2VError(f"roi_crop_backward: grad_out {tuple(grad_out.shape)} != {(N, C, oh, ow)}")
    gin = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(load().i2v_roi_crop_backward(_p(grad_out), _p(grids), _p(gin), B, C, H, W, N, oh, ow, _stream()),
              "i2v_roi_crop_backward")
    return gin


def pair_build(boxes, im_h: float, im_w: float, margin: float = 10.0, want_masks: bool = True):
    """boxes [N,4] -> (ixs [P], ixo [P] int64, rel_boxes [P,5], masks [P,2,32,32] or None), P = N(N-1)."""
    boxes = _f32(boxes, "boxes")
    N = boxes.size(0)
    P = N * (N - 1) if N > 1 else 0
    dev = boxes.device
    ixs = torch.empty((P,), dtype=torch.int64, device=dev)
    ixo = torch.empty((P,), dtype=torch.int64, device=dev)
    rel = torch.empty((P, 5), dtype=torch.float32, device=dev)
    masks = torch.empty((P, 2, 32, 32), dtype=torch.float32, device=dev) if want_masks else None
    with torch.cuda.device(dev):
        check(load().i2v_pair_build(_p(boxes), N, float(im_h), float(im_w), float(margin), _p(ixs), _p(ixo), _p(rel),
                                    _p(masks), _stream()), "i2v_pair_build")
    return ixs, ixo, rel, masks


def triplet_topk(rel_score, conf, classes, boxes, ixs, ixo, top_k: int = 100):
    """lib/utils.py:609-626 on the device -> (records [top_k,13] fp32, count int32[1])."""
    rel_score, conf, boxes = _f32(rel_score, "rel_score"), _f32(conf, "conf"), _f32(boxes, "boxes")
    classes, ixs, ixo = classes.long().contiguous(), ixs.long().contiguous(), ixo.long().contiguous()
    P, R = rel_score.shape
    dev = rel_score.device
    rec = torch.empty((top_k, 13), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.i2v_triplet_topk_workspace_bytes(P, R), dev)
        check(lib.i2v_triplet_topk(_p(rel_score), _p(conf), _p(classes), _p(boxes), _p(ixs), _p(ixo), P, R,
                                   int(top_k), _p(rec), _p(cnt), _p(ws), ws.numel(), _stream()), "i2v_triplet_topk")
    return rec, cnt


def pair_build_frames(boxes, im_h: float, im_w: float, margin: float = 10.0, want_masks: bool = True):
    """boxes [F,N,4] -> (ixs [F*P], ixo [F*P] int64 = group rows f*N+i, rel_boxes [F*P,5] with the frame number in column
    0, obj_masks [F*N,32,32] or None): the pair stage of a whole frame group in one launch."""
    boxes = _f32(boxes, "boxes")
    if boxes.dim() != 3 or boxes.size(2) != 4:
        raise _lib.I2VError(f"pair_build_frames: boxes must be [F,N,4], got {tuple(boxes.shape)}")
    F, N = boxes.shape[:2]
    P = N * (N - 1) if N > 1 else 0
    dev = boxes.device
    ixs = torch.empty((F * P,), dtype=torch.int64, device=dev)
    ixo = torch.empty((F * P,), dtype=torch.int64, device=dev)
    rel = torch.empty((F * P, 5), dtype=torch.float32, device=dev)
    masks = torch.empty((F * N, 32, 32), dtype=torch.float32, device=dev) if want_masks else None
    with torch.cuda.device(dev):
        check(load().i2v_pair_build_frames(_p(boxes), F, N, float(im_h), float(im_w), float(margin), _p(ixs), _p(ixo),
                                           _p(rel), _p(masks), _stream()), "i2v_pair_build_frames")
    return ixs, ixo, rel, masks


def triplet_topk_frames(rel_score, conf, classes, boxes, ixs, ixo, top_k: int = 100):
    """lib/utils.py:609-626 for F frames of N detections each in three launches: rel_score [F*P,R], conf / classes [F,N],
    boxes [F,N,4], ixs / ixo [P] frame-local -> (records [F,top_k,13], counts [F] int32)."""
    rel_score, conf, boxes = _f32(rel_score, "rel_score"), _f32(conf, "conf"), _f32(boxes, "boxes")
    classes, ixs, ixo = classes.long().contiguous(), ixs.long().contiguous(), ixo.long().contiguous()
    F, N = conf.shape
    P, R = ixs.numel(), rel_score.size(1)
    if rel_score.size(0) != F * P or classes.shape != (F, N) or boxes.shape != (F, N, 4):
        raise _lib.I2VError("triplet_topk_frames: shapes do not describe F frames of N detections")
    dev = rel_score.device
    rec = torch.empty((F, top_k, 13), dtype=torch.float32, device=dev)
    cnt = torch.empty((F,), dtype=torch.int32, device=dev)
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.i2v_triplet_topk_frames_workspace_bytes(F, P, R), dev)
        check(lib.i2v_triplet_topk_frames(_p(rel_score), _p(conf), _p(classes), _p(boxes), _p(ixs), _p(ixo), F, N, P, R,
                                          int(top_k), _p(rec), _p(cnt), _p(ws), ws.numel(), _stream()),
              "i2v_triplet_topk_frames")
    return rec, cnt


# --------------------------------------------------------------------------------------- SGG projection
_TORCH_DT = {torch.float32: DT_F32, torch.bfloat16: DT_BF16}


def roi_pool_rows(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float, dtype=torch.bfloat16, out=None):
    """`roi_pool(fmap, boxes).view(N, -1)` (resnet_SGG_emb.py:144-146,158-160) -> [N, C*ph*pw] fp32 or bf16.
    `out` may be a row slice of a larger matrix (the object and union rows share one GEMM)."""
    features, rois = _f32(features, "features"), _rois5(rois)
    B, C, H, W = features.shape
    N, K = rois.size(0), C * pooled_h * pooled_w
    if out is None:
        out = torch.empty((N, K), dtype=dtype, device=features.device)
    if out.shape != (N, K) or out.stride(1) != 1 or out.dtype not in _TORCH_DT or not out.is_cuda:
        raise _lib.I2VError("roi_pool_rows: bad `out`")
    with torch.cuda.device(features.device):
        check(load().i2v_roi_pool_rows(_p(features), _p(rois), _p(out), B, C, H, W, N, pooled_h, pooled_w,
                                       float(spatial_scale), out.stride(0), _TORCH_DT[out.dtype], _stream()),
              "i2v_roi_pool_rows")
    return out


def linear(x, weight, bias=None, relu: bool = False, out=None, out_dtype=torch.float32, keep_mask=None, keep_scale: float = 1.0):
    """FC of lib/model/faster_rcnn/utils.py:48-60 on tcgen05: act(x @ weight.T + bias).  x [M,K] and weight [N,K] are both bf16
    (tensor cores in bf16) or both fp32 (tensor cores in tf32); rows may be strided.  `out` may be a column slice.
    `keep_mask` [M,N] uint8 applies inverted dropout in the epilogue: y = keep ? y * keep_scale : 0."""
    if not (x.is_cuda and weight.is_cuda) or x.dtype != weight.dtype or x.dtype not in _TORCH_DT:
        raise _lib.I2VError("linear: x and weight must be CUDA tensors, both bf16 or both fp32")
    if x.dim() != 2 or weight.dim() != 2 or x.size(1) != weight.size(1) or x.stride(1) != 1 or weight.stride(1) != 1:
        raise _lib.I2VError(f"linear: bad shapes {tuple(x.shape)} x {tuple(weight.shape)}")
    M, K = x.shape
    N = weight.size(0)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    if out.shape != (M, N) or out.stride(1) != 1 or out.dtype not in _TORCH_DT or not out.is_cuda:
        raise _lib.I2VError("linear: bad `out`")
    if bias is not None:
        bias = _f32(bias, "bias")
        if bias.numel() != N:
            raise _lib.I2VError("linear: bias size")
    in_dt = DT_BF16 if x.dtype == torch.bfloat16 else DT_TF32
    with torch.cuda.device(x.device):
        if keep_mask is not None:
            if keep_mask.dtype != torch.uint8 or keep_mask.shape != (M, N) or keep_mask.stride(1) != 1 or not keep_mask.is_cuda:
                raise _lib.I2VError("linear: keep_mask must be a [M,N] uint8 CUDA tensor with contiguous rows")
            check(load().i2v_linear_forward_dropout(_p(x), _p(weight), _p(bias), _p(out), M, N, K, x.stride(0), weight.stride(0),
                                                    out.stride(0), in_dt, _TORCH_DT[out.dtype], int(bool(relu)),
                                                    _p(keep_mask), keep_mask.stride(0), float(keep_scale), _stream()),
                  "i2v_linear_forward_dropout")
        else:
            check(load().i2v_linear_forward(_p(x), _p(weight), _p(bias), _p(out), M, N, K, x.stride(0), weight.stride(0),
                                            out.stride(0), in_dt, _TORCH_DT[out.dtype], int(bool(relu)), _stream()),
                  "i2v_linear_forward")
    return out


def round_tf32(x, out=None):
    """fp32 [rows, cols] (rows may be strided) -> the nearest tf32 values, as fp32 words (`out` may be `x`)."""
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1 or x.dtype != torch.float32:
        raise _lib.I2VError("round_tf32: expected a 2-d fp32 CUDA tensor with contiguous rows")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().i2v_round_tf32(_p(x), _p(out), x.size(0), x.size(1), x.stride(0), out.stride(0), _stream()),
              "i2v_round_tf32")
    return out


def cast_bf16(src, out=None):
    """fp32 [rows, cols] (rows may be strided) -> bf16, round to nearest even."""
    src = src if src.dtype == torch.float32 else src.float()
    if not src.is_cuda or src.dim() != 2 or src.stride(1) != 1:
        raise _lib.I2VError("cast_bf16: expected a 2-d CUDA tensor with contiguous rows")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        check(load().i2v_cast_bf16(_p(src), _p(out), src.size(0), src.size(1), src.stride(0), out.stride(0),
                                   _stream()), "i2v_cast_bf16")
    return out


def rel_scores(x, prd, softmax: bool = True):
    """resnet_SGG_emb.py:207-219: softmax(normalize(x) @ normalize(prd).T) -> [P, R] fp32."""
    x, prd = _f32(x, "x"), _f32(prd, "prd")
    P, E = x.shape
    R = prd.size(0)
    out = torch.empty((P, R), dtype=torch.float32, device=x.device)
    lib = load()
    with torch.cuda.device(x.device):
        ws = _workspace(lib.i2v_rel_scores_workspace_bytes(R, E), x.device)
        check(lib.i2v_rel_scores(_p(x), _p(prd), _p(out), P, R, E, int(bool(softmax)), _p(ws), ws.numel(), _stream()),
              "i2v_rel_scores")
    return out


def conv2d_nhwc_split(x, weight_taps, bias, kernel: int, pad: int, relu: bool = True, out_dtype=torch.bfloat16):
    """`conv2d_nhwc` with stride 2 on a parity-split activation x [N, 2, 2, H/2, W/2, C] bf16 (plane (y & 1, x & 1),
    position (y >> 1, x >> 1)) -> [N, H/2, W/2, O]; same bits as the strided call on the unsplit map."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 6 or not x.is_contiguous() or x.size(1) != 2 or x.size(2) != 2:
        raise _lib.I2VError("conv2d_nhwc_split: expected a contiguous [N,2,2,H/2,W/2,C] bf16 CUDA tensor")
    n, _, _, h2, w2, c = x.shape
    o = weight_taps.size(0)
    out = torch.empty((n, h2, w2, o), dtype=out_dtype, device=x.device)
    b = None if bias is None else _f32(bias, "bias")
    with torch.cuda.device(x.device):
        check(load().i2v_conv2d_nhwc_split_forward(_p(x), _p(weight_taps), _p(b), _p(out), n, 2 * h2, 2 * w2, c, o, kernel, pad,
                                                   weight_taps.stride(0), o, _TORCH_DT[out_dtype], int(bool(relu)), _stream()),
              "i2v_conv2d_nhwc_split_forward")
    return out


def im2col_bf16(x, kernel: int, stride: int, pad: int, layout: str, ld: int | None = None, out_dtype=torch.bfloat16):
    """Patches of a square-kernel convolution as bf16 (or fp32) rows [(n, oy, ox), (ky, kx, c)] (pitch `ld`, zero padded).
    x is [N,C,H,W] (`layout='nchw'`) or [N,H,W,C] (`layout='nhwc'`), fp32 or bf16, contiguous."""
    if not x.is_cuda or x.dim() != 4 or x.dtype not in _TORCH_DT or not x.is_contiguous():
        raise _lib.I2VError("im2col_bf16: expected a contiguous 4-d fp32/bf16 CUDA tensor")
    if layout == "nchw":
        n, c, h, w = x.shape
        sn, sc, sy, sx = c * h * w, h * w, w, 1
    elif layout == "nhwc":
        n, h, w, c = x.shape
        sn, sc, sy, sx = h * w * c, 1, w * c, c
    else:
        raise _lib.I2VError("im2col_bf16: layout must be nchw or nhwc")
    oh, ow = (h + 2 * pad - kernel) // stride + 1, (w + 2 * pad - kernel) // stride + 1
    k = kernel * kernel * c
    ld = k if ld is None else int(ld)
    out = torch.empty((n * oh * ow, ld), dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        fn = load().i2v_im2col_f32 if out_dtype == torch.float32 else load().i2v_im2col_bf16
        check(fn(_p(x), _TORCH_DT[x.dtype], n, c, h, w, sn, sc, sy, sx, kernel, kernel, stride, pad, _p(out), ld, _stream()),
              "i2v_im2col")
    return out, (n, oh, ow)


def pair_rows_bf16(obj, ixs, ixo, out=None):
    """cat(obj[ixs], obj[ixo], dim=1) as bf16 rows [P, 2E] (resnet_SGG_emb.py:150-151,169)."""
    obj = _f32(obj, "obj")
    ixs, ixo = ixs.long().contiguous(), ixo.long().contiguous()
    N, E = obj.shape
    P = ixs.numel()
    if out is None:
        out = torch.empty((P, 2 * E), dtype=torch.bfloat16, device=obj.device)
    if out.shape != (P, 2 * E) or out.stride(1) != 1 or out.dtype != torch.bfloat16:
        raise _lib.I2VError("pair_rows_bf16: bad `out`")
    with torch.cuda.device(obj.device):
        check(load().i2v_pair_rows_bf16(_p(obj), _p(ixs), _p(ixo), _p(out), N, P, E, out.stride(0), _stream()),
              "i2v_pair_rows_bf16")
    return out


def gather_rows_bf16(src, idx, out=None):
    """out[p] = src[idx[p]] for bf16 rows; `out` may be a column slice of a wider matrix."""
    if not src.is_cuda or src.dtype != torch.bfloat16 or src.dim() != 2 or src.stride(1) != 1:
        raise _lib.I2VError("gather_rows_bf16: expected a 2-d bf16 CUDA tensor with contiguous rows")
    idx = idx.long().contiguous()
    P, D = idx.numel(), src.size(1)
    if out is None:
        out = torch.empty((P, D), dtype=torch.bfloat16, device=src.device)
    if out.shape != (P, D) or out.stride(1) != 1 or out.dtype != torch.bfloat16:
        raise _lib.I2VError("gather_rows_bf16: bad `out`")
    with torch.cuda.device(src.device):
        check(load().i2v_gather_rows_bf16(_p(src), _p(idx), _p(out), src.size(0), P, D, src.stride(0), out.stride(0),
                                          _stream()), "i2v_gather_rows_bf16")
    return out


def greedy_association(records, counts, frame_numbers=None, source_frame=None, max_traj: int = 100):
    """lib/utils.py:134-182 on the device.  records [F,K,13] fp32, counts [F] int32 ->
    (rel_id [F,K] int32, order [F,K] int32, rel_info [F*K,6] int32, rel_score [F*K] float64, num_rel int32[1])."""
    records = _f32(records, "records")
    F, K, w = records.shape
    if w != 13:
        raise _lib.I2VError("greedy_association: records must be [F, K, 13]")
    dev = records.device
    i32 = lambda t: None if t is None else torch.as_tensor(t, device=dev).to(torch.int32).contiguous()
    counts, frame_numbers, source_frame = i32(counts), i32(frame_numbers), i32(source_frame)
    rel_id = torch.empty((F, K), dtype=torch.int32, device=dev)
    order = torch.empty((F, K), dtype=torch.int32, device=dev)
    info = torch.zeros((max(F * K, 1), 6), dtype=torch.int32, device=dev)
    score = torch.zeros((max(F * K, 1),), dtype=torch.float64, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.i2v_association_workspace_bytes(F, K), dev)
        check(lib.i2v_greedy_association(_p(records), _p(counts), _p(frame_numbers), _p(source_frame), F, K, int(max_traj),
                                         _p(rel_id), _p(order), _p(info), _p(score), _p(num), _p(ws), ws.numel(),
                                         _stream()), "i2v_greedy_association")
    return rel_id, order, info, score, num


def pair_conv1_bf16(obj_maps, ixs, ixo, bias, relu: bool = True, split_hw=None):
    """obj_maps [N, positions, 2C] fp32 (per-object single-channel convolutions, subject half then object half) ->
    relu(obj_maps[ixs][..., :C] + obj_maps[ixo][..., C:] + bias) as NHWC bf16 [P, positions, C]; with `split_hw` = (oh, ow)
    (positions = oh*ow, both even) in the parity-split layout [P, 2, 2, oh/2, ow/2, C] of `conv2d_nhwc_split`."""
    obj_maps = _f32(obj_maps, "obj_maps")
    N, positions, c2 = obj_maps.shape
    C = c2 // 2
    ixs, ixo = ixs.long().contiguous(), ixo.long().contiguous()
    P = ixs.numel()
    b = None if bias is None else _f32(bias, "bias")
    if split_hw is not None:
        oh, ow = split_hw
        if oh * ow != positions:
            raise _lib.I2VError("pair_conv1_bf16: split_hw does not match the number of positions")
        out = torch.empty((P, 2, 2, oh // 2, ow // 2, C), dtype=torch.bfloat16, device=obj_maps.device)
        with torch.cuda.device(obj_maps.device):
            check(load().i2v_pair_conv1_split_bf16(_p(obj_maps), _p(ixs), _p(ixo), _p(b), _p(out), N, P, oh, ow, C,
                                                   int(bool(relu)), _stream()), "i2v_pair_conv1_split_bf16")
        return out
    out = torch.empty((P, positions, C), dtype=torch.bfloat16, device=obj_maps.device)
    with torch.cuda.device(obj_maps.device):
        check(load().i2v_pair_conv1_bf16(_p(obj_maps), _p(ixs), _p(ixo), _p(b), _p(out), N, P, positions, C,
                                         int(bool(relu)), _stream()), "i2v_pair_conv1_bf16")
    return out


def conv2d_nhwc(x, weight_taps, bias, kernel: int, stride: int, pad: int, relu: bool = True, out_dtype=torch.bfloat16):
    """Implicit-GEMM convolution on tcgen05: x [N,H,W,C] bf16 NHWC, weight_taps [O, kernel*kernel*Cp] bf16 (taps in
    (ky, kx) order, channels of each tap padded to Cp = 64*ceil(C/64)) -> [N,OH,OW,O].  Raises I2VError (unsupported) for
    shapes it does not take; callers fall back to im2col_bf16 + linear."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4 or not x.is_contiguous():
        raise _lib.I2VError("conv2d_nhwc: expected a contiguous NHWC bf16 CUDA tensor")
    n, h, w, c = x.shape
    o = weight_taps.size(0)
    oh, ow = (h + 2 * pad - kernel) // stride + 1, (w + 2 * pad - kernel) // stride + 1
    out = torch.empty((n, oh, ow, o), dtype=out_dtype, device=x.device)
    b = None if bias is None else _f32(bias, "bias")
    with torch.cuda.device(x.device):
        check(load().i2v_conv2d_nhwc_forward(_p(x), _p(weight_taps), _p(b), _p(out), n, h, w, c, o, kernel, stride, pad,
                                             weight_taps.stride(0), o, _TORCH_DT[out_dtype], int(bool(relu)), _stream()),
              "i2v_conv2d_nhwc_forward")
    return out
