"""The reference itself, as far as it can be run -- TEST INFRASTRUCTURE ONLY.

Three things live here, all used to *pin* oracle/oracle.c and the CUDA path:

* ``cpu_*``   -- ``oracle/_ref/libref_cpu.so``: the reference's own ``roi_align.c`` and
  ``roi_pooling.c`` compiled unmodified (``make -C oracle ref``) behind the ``TH/TH.h`` shim.
* ``cuda_*``  -- ``oracle/_ref/libref_cuda.so``: the reference's own ``roi_align_kernel.cu``,
  ``roi_pooling_kernel.cu`` and ``nms_cuda_kernel.cu`` compiled unmodified for sm_100a; callable
  only on a GPU box (device pointers come from torch tensors).
* ``py_*``    -- the reference's Python (``nms_cpu.py``, ``proposal_layer.py`` ...) imported from
  ``/root/reference`` behind an in-memory ``easydict`` shim.  ``/root/reference`` exists only in
  the build container, so these are used by ``tests/golden/make_golden.py`` to write fixtures
  and by tests marked ``needs_reference``; nothing on the GPU box calls them.
"""
from __future__ import annotations

import ctypes
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

_f = ctypes.c_float
_i = ctypes.c_int
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p


def have_cpu_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cpu.so"))


def have_cuda_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cuda.so"))


def have_py_ref() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "model"))


# ------------------------------------------------------------------ reference CPU C (unmodified)
class _THFloatStorage(ctypes.Structure):
    _fields_ = [("data", _fp), ("numel", ctypes.c_long)]


class _THFloatTensor(ctypes.Structure):
    _fields_ = [("data", _fp), ("size", ctypes.c_long * 4), ("ndim", _i), ("storage", _THFloatStorage)]


def _th(a: np.ndarray) -> _THFloatTensor:
    t = _THFloatTensor()
    t.data = a.ctypes.data_as(_fp)
    for d in range(4):
        t.size[d] = a.shape[d] if d < a.ndim else 1
    t.ndim = a.ndim
    t.storage.data = t.data
    t.storage.numel = a.size
    return t


_cpu = None


def _cpu_lib():
    global _cpu
    if _cpu is None:
        _cpu = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_cpu.so"))
    return _cpu


def cpu_roi_align_forward(feat, rois, gh, gw, scale) -> np.ndarray:
    """roi_align.c:17-45 `roi_align_forward` (the cffi entry point itself)."""
    feat = np.ascontiguousarray(feat, np.float32)
    rois = np.ascontiguousarray(rois, np.float32).reshape(-1, 5)
    out = np.zeros((rois.shape[0], feat.shape[1], gh, gw), np.float32)
    tf, tr, to = _th(feat), _th(rois), _th(out)
    rc = _cpu_lib().roi_align_forward(_i(gh), _i(gw), _f(scale), ctypes.byref(tf), ctypes.byref(tr), ctypes.byref(to))
    assert rc == 1
    return out


def cpu_roi_pooling_forward_nhwc(feat_nhwc, rois, ph, pw, scale) -> np.ndarray:
    """roi_pooling.c:4-104 `roi_pooling_forward` (NHWC input, batch 1, no argmax)."""
    feat = np.ascontiguousarray(feat_nhwc, np.float32)
    rois = np.ascontiguousarray(rois, np.float32).reshape(-1, 5)
    out = np.zeros((rois.shape[0], feat.shape[3], ph, pw), np.float32)
    tf, tr, to = _th(feat), _th(rois), _th(out)
    rc = _cpu_lib().roi_pooling_forward(_i(ph), _i(pw), _f(scale), ctypes.byref(tf), ctypes.byref(tr), ctypes.byref(to))
    assert rc == 1
    return out


def cpu_roi_crop_forward_bhwd(feat_bhwc, grids) -> np.ndarray:
    """roi_crop.c:7-103 `BilinearSamplerBHWD_updateOutput`: the CPU twin samples image b with grid b (no RoIs-per-image
    division) from a channel-last tensor: feat [B,H,W,C], grids [B,oh,ow,2] -> [B,oh,ow,C]."""
    feat = np.ascontiguousarray(feat_bhwc, np.float32)
    grids = np.ascontiguousarray(grids, np.float32)
    out = np.zeros((grids.shape[0], grids.shape[1], grids.shape[2], feat.shape[3]), np.float32)
    tf, tg, to = _th(feat), _th(grids), _th(out)
    rc = _cpu_lib().BilinearSamplerBHWD_updateOutput(ctypes.byref(tf), ctypes.byref(tg), ctypes.byref(to))
    assert rc == 1
    return out


# ------------------------------------------------------------------ reference CUDA kernels (unmodified)
_cuda = None


def _cuda_lib():
    global _cuda
    if _cuda is None:
        _cuda = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_cuda.so"))
    return _cuda


def _stream():
    import torch
    return _vp(torch.cuda.current_stream().cuda_stream)


def cuda_roi_align_forward(feat, rois, gh, gw, scale):
    """ROIAlignForwardLaucher (roi_align_kernel.cu:73-91) on torch CUDA tensors."""
    import torch
    N, C = rois.shape[0], feat.shape[1]
    out = torch.zeros((N, C, gh, gw), device=feat.device, dtype=torch.float32)
    _cuda_lib().ROIAlignForwardLaucher(_vp(feat.data_ptr()), _f(scale), _i(N), _i(feat.shape[2]), _i(feat.shape[3]),
                                       _i(C), _i(gh), _i(gw), _vp(rois.data_ptr()), _vp(out.data_ptr()), _stream())
    return out


def cuda_roi_align_backward(top_diff, rois, feat_shape, gh, gw, scale):
    """ROIAlignBackwardLaucher (roi_align_kernel.cu:145-162)."""
    import torch
    B, C, H, W = feat_shape
    gin = torch.zeros((B, C, H, W), device=top_diff.device, dtype=torch.float32)
    _cuda_lib().ROIAlignBackwardLaucher(_vp(top_diff.data_ptr()), _f(scale), _i(B), _i(rois.shape[0]), _i(H), _i(W),
                                        _i(C), _i(gh), _i(gw), _vp(rois.data_ptr()), _vp(gin.data_ptr()), _stream())
    return gin


def cuda_roi_crop_forward(feat, grids):
    """BilinearSamplerBHWD_updateOutput_cuda_kernel (roi_crop_cuda_kernel.cu:200-253) with the arguments of
    roi_crop_cuda.c:21-44 on contiguous torch CUDA tensors; the output is pre-zeroed like functions/roi_crop.py:11."""
    import torch
    B, C, H, W = feat.shape
    N, oh, ow, _ = grids.shape
    out = torch.zeros((N, C, oh, ow), device=feat.device, dtype=torch.float32)
    rc = _cuda_lib().BilinearSamplerBHWD_updateOutput_cuda_kernel(
        _i(C), _i(ow), _i(oh), _i(N), _i(C), _i(H), _i(W), _i(B), _vp(feat.data_ptr()), _i(C * H * W), _i(H * W), _i(W), _i(1),
        _vp(grids.data_ptr()), _i(oh * ow * 2), _i(1), _i(ow * 2), _i(2), _vp(out.data_ptr()), _i(C * oh * ow), _i(oh * ow),
        _i(ow), _i(1), _stream())
    assert rc == 1
    return out


def cuda_roi_crop_backward(feat, grids, grad_out):
    """BilinearSamplerBHWD_updateGradInput_cuda_kernel (roi_crop_cuda_kernel.cu:255-335) -> (grad_input, grad_grids)."""
    import torch
    B, C, H, W = feat.shape
    N, oh, ow, _ = grids.shape
    gi, gg = torch.zeros_like(feat), torch.zeros_like(grids)
    rc = _cuda_lib().BilinearSamplerBHWD_updateGradInput_cuda_kernel(
        _i(C), _i(ow), _i(oh), _i(N), _i(C), _i(H), _i(W), _i(B), _vp(feat.data_ptr()), _i(C * H * W), _i(H * W), _i(W), _i(1),
        _vp(grids.data_ptr()), _i(oh * ow * 2), _i(1), _i(ow * 2), _i(2), _vp(gi.data_ptr()), _i(C * H * W), _i(H * W), _i(W),
        _i(1), _vp(gg.data_ptr()), _i(oh * ow * 2), _i(1), _i(ow * 2), _i(2), _vp(grad_out.data_ptr()), _i(C * oh * ow),
        _i(oh * ow), _i(ow), _i(1), _stream())
    assert rc == 1
    return gi, gg


def cuda_roi_pool_forward(feat, rois, ph, pw, scale):
    """ROIPoolForwardLaucher (roi_pooling_kernel.cu:95-125)."""
    import torch
    N, C = rois.shape[0], feat.shape[1]
    out = torch.zeros((N, C, ph, pw), device=feat.device, dtype=torch.float32)
    arg = torch.zeros((N, C, ph, pw), device=feat.device, dtype=torch.int32)
    _cuda_lib().ROIPoolForwardLaucher(_vp(feat.data_ptr()), _f(scale), _i(N), _i(feat.shape[2]), _i(feat.shape[3]),
                                      _i(C), _i(ph), _i(pw), _vp(rois.data_ptr()), _vp(out.data_ptr()),
                                      _vp(arg.data_ptr()), _stream())
    return out, arg


def cuda_roi_pool_backward(top_diff, rois, argmax, feat_shape, ph, pw, scale):
    """ROIPoolBackwardLaucher (roi_pooling_kernel.cu:205-234)."""
    import torch
    B, C, H, W = feat_shape
    gin = torch.zeros((B, C, H, W), device=top_diff.device, dtype=torch.float32)
    _cuda_lib().ROIPoolBackwardLaucher(_vp(top_diff.data_ptr()), _f(scale), _i(B), _i(rois.shape[0]), _i(H), _i(W),
                                       _i(C), _i(ph), _i(pw), _vp(rois.data_ptr()), _vp(gin.data_ptr()),
                                       _vp(argmax.data_ptr()), _stream())
    return gin


def cuda_nms(dets_sorted_host: np.ndarray, thresh: float):
    """nms_cuda_compute (nms_cuda_kernel.cu:87-161): host boxes in, device keep/num out."""
    import torch
    dets = np.ascontiguousarray(dets_sorted_host, np.float32)
    n, dim = dets.shape
    keep = torch.zeros(n, device="cuda", dtype=torch.int32)
    num = torch.zeros(1, device="cuda", dtype=torch.int32)
    torch.cuda.synchronize()
    _cuda_lib().nms_cuda_compute(_vp(keep.data_ptr()), _vp(num.data_ptr()), dets.ctypes.data_as(_fp), _i(n), _i(dim),
                                 _f(thresh))
    torch.cuda.synchronize()
    return keep[: int(num.item())].cpu().numpy()


# ------------------------------------------------------------------ reference Python (build container only)
class _EasyDict(dict):
    """What `from easydict import EasyDict` gives config.py:11: attribute access == item access, nested."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)
        super().__setattr__(k, v)

    __setattr__ = __setitem__


_py_ready = False


def _py_setup():
    global _py_ready
    if _py_ready:
        return
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    if not hasattr(np, "float"):      # lib/utils.py-era aliases some reference files still use
        np.float = float
    sys.path.insert(0, os.path.join(REF_ROOT, "lib"))
    _py_ready = True


def py_nms_cpu(dets: np.ndarray, thresh: float) -> np.ndarray:
    """Executes lib/model/nms/nms_cpu.py:6-34 unmodified."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("_ref_nms_cpu", os.path.join(REF_ROOT, "lib/model/nms/nms_cpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.nms_cpu(torch.from_numpy(np.ascontiguousarray(dets, np.float32)), thresh).numpy()


def py_generate_anchors(**kw) -> np.ndarray:
    """Executes lib/model/rpn/generate_anchors.py:45-56 unmodified."""
    _py_setup()
    from model.rpn.generate_anchors import generate_anchors
    return generate_anchors(**kw)


def py_proposal_layer(cls_prob, bbox_pred, im_info, cfg_key="TEST", target=False, overrides=None):
    """Executes _ProposalLayer.forward (lib/model/rpn/proposal_layer.py:49-163) unmodified on CPU tensors."""
    _py_setup()
    import torch
    from model.utils.config import cfg
    from model.rpn.proposal_layer import _ProposalLayer
    for k, v in (overrides or {}).items():
        cfg[cfg_key][k] = v
    layer = _ProposalLayer(cfg.FEAT_STRIDE[0], cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    with torch.no_grad():
        out = layer((torch.from_numpy(cls_prob), torch.from_numpy(bbox_pred), torch.from_numpy(im_info), cfg_key),
                    target)
    return out.numpy()


def py_bbox_transform_inv_clip(anchors, deltas, im_info):
    """Executes bbox_transform.py:77-103 and :125-133 unmodified."""
    _py_setup()
    import torch
    from model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    B = deltas.shape[0]
    p = bbox_transform_inv(torch.from_numpy(anchors), torch.from_numpy(deltas), B)
    return clip_boxes(p, torch.from_numpy(im_info), B).numpy()


def py_vrd_forward(params: dict, args, prd_vecs, fmap, boxes, rel_boxes, spatial, classes, ix1, ix2, train_masks=None):
    """Executes `vrd.forward` (lib/model/faster_rcnn/resnet_SGG_emb.py:128-221) UNMODIFIED on the CPU.

    What is stubbed, and only that: `model._C` has no source in the reference (lib/setup.py:19), so
    `model.roi_layers.ROIPool` is bound to torchvision.ops.roi_pool (the same maskrcnn-benchmark op); the detector base
    class `_fasterRCNN` (not on this path) is an empty nn.Module; `.cuda()` is the identity because the build container
    has no GPU; the three annotation pickles the constructor opens (:75-80) are empty temporaries.
    `params` maps the reference's state_dict keys to numpy arrays.  Returns (scores [P,n_rel], feat [P,emb]).

    `train_masks` (four keep masks, in the order of the F.dropout calls at :148, :149, :162, :163) runs the module in
    TRAINING mode with one more stub: the module's `F.dropout` applies those masks (x * keep / (1 - p), p = 0.5) instead of
    drawing its own, so that the result is reproducible."""
    _py_setup()
    import pickle
    import tempfile
    import torch
    import torchvision
    from torch import nn

    if "model.roi_layers" not in sys.modules or not getattr(sys.modules["model.roi_layers"], "_i2v_stub", False):
        rl = types.ModuleType("model.roi_layers")
        rl._i2v_stub = True

        class ROIPool(nn.Module):
            def __init__(self, output_size, spatial_scale):
                super().__init__()
                self.output_size, self.spatial_scale = output_size, spatial_scale

            def forward(self, input, rois):
                return torchvision.ops.roi_pool(input, rois, self.output_size, self.spatial_scale)

        rl.ROIPool = ROIPool
        rl.ROIAlign = None
        sys.modules["model.roi_layers"] = rl
        fr = types.ModuleType("model.faster_rcnn.faster_rcnn_SGG_emb")
        fr._fasterRCNN = type("_fasterRCNN", (nn.Module,), {})
        sys.modules["model.faster_rcnn.faster_rcnn_SGG_emb"] = fr
    from model.faster_rcnn import resnet_SGG_emb as ref_mod

    tmp = tempfile.mkdtemp()
    for name in ("source_so_prior_path", "source_gt_rels_path", "target_gt_rels_path"):
        path = os.path.join(tmp, name + ".pkl")
        with open(path, "wb") as f:
            pickle.dump([], f)
        setattr(args, name, path)
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        net = ref_mod.vrd(args, None, np.asarray(prd_vecs, np.float32))
        sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}
        missing = net.load_state_dict(sd, strict=True)
        net.eval()
        old_f = ref_mod.F
        if train_masks is not None:
            net.train()
            queue = [torch.from_numpy(np.asarray(m, np.float32)) for m in train_masks]

            class _F:
                def __getattr__(self, name):
                    return getattr(old_f, name)

                @staticmethod
                def dropout(x, p=0.5, training=True, inplace=False):
                    assert training and p == 0.5
                    return x * queue.pop(0) * 2.0

            ref_mod.F = _F()
        net.obj_vecs = np.zeros((args.num_classes + 1, 300), np.float32)
        with torch.no_grad():
            scores, feat = net(np.asarray(fmap, np.float32), np.asarray(boxes, np.float32),
                               np.asarray(rel_boxes, np.float32), np.asarray(spatial, np.float32),
                               list(classes), np.asarray(ix1), np.asarray(ix2))
    finally:
        torch.Tensor.cuda = old_cuda
        try:
            ref_mod.F = old_f
        except NameError:
            pass
    return scores.numpy(), feat


def py_proposal_target_layer(all_rois, gt_boxes, num_classes=21, seed=0, overrides=None):
    """Executes `_ProposalTargetLayer.forward` (lib/model/rpn/proposal_target_layer_cascade.py:33-59) unmodified on CPU
    tensors after `np.random.seed(seed)`.  Returns the five outputs as numpy arrays."""
    _py_setup()
    import torch
    from model.utils.config import cfg
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    for k, v in (overrides or {}).items():
        cfg.TRAIN[k] = v
    layer = _ProposalTargetLayer(num_classes)
    np.random.seed(seed)
    with torch.no_grad():
        out = layer(torch.from_numpy(np.ascontiguousarray(all_rois, np.float32)),
                    torch.from_numpy(np.ascontiguousarray(gt_boxes, np.float32)), None)
    return [o.numpy() for o in out]


def py_anchor_target_layer(gt_boxes, im_info, feat_h=38, feat_w=63, seed=0):
    """Executes `_AnchorTargetLayer.forward` (lib/model/rpn/anchor_target_layer.py:48-193) unmodified on CPU tensors after
    `np.random.seed(seed)`.  Returns [labels, bbox_targets, inside_weights, outside_weights] as numpy arrays."""
    _py_setup()
    import torch
    from model.utils.config import cfg
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    layer = _AnchorTargetLayer(cfg.FEAT_STRIDE[0], cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    gt = torch.from_numpy(np.ascontiguousarray(gt_boxes, np.float32))
    score = torch.zeros((gt.shape[0], 18, feat_h, feat_w))
    np.random.seed(seed)
    with torch.no_grad():
        out = layer((score, gt, torch.from_numpy(np.ascontiguousarray(im_info, np.float32)), None))
    return [o.numpy() for o in out]


def py_rpn_cls_prob(cls_score: np.ndarray) -> np.ndarray:
    """Executes rpn.py:66-68 on the CPU: `_RPN.reshape` (rpn.py:46-55, taken from the unmodified source file) around
    `F.softmax(., 1)`.  Only the static method is extracted: the class body needs the cffi-era extensions to import."""
    import ast
    import torch
    import torch.nn.functional as F
    src = open(os.path.join(REF_ROOT, "lib/model/rpn/rpn.py")).read()
    fn = next(n for c in ast.walk(ast.parse(src)) if isinstance(c, ast.ClassDef) and c.name == "_RPN"
              for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "reshape")
    fn.decorator_list = []
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "rpn.py:_RPN.reshape", "exec"), ns)
    reshape = ns["reshape"]
    x = torch.from_numpy(np.ascontiguousarray(cls_score, np.float32))
    nc_score_out = x.shape[1]
    with torch.no_grad():
        prob = reshape(F.softmax(reshape(x, 2), 1), nc_score_out)
    return prob.numpy()
