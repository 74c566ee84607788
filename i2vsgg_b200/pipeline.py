"""The hot path as one object with pre-allocated device buffers: proposal layer -> RoIAlignAvg forward ->
RoIAlignAvg backward for a batch of frames, either on device-resident inputs (`device_step`) or from/to pinned host
buffers (`host_step`, the end-to-end call).  All compute goes through the C ABI; torch only owns memory and streams.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, synth
from ._lib import IMPL_AUTO, POOL_AVG, check, load


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class StageTimer:
    """CUDA events at the stage boundaries of `device_step`, recorded on the launching stream."""

    STAGES = ("proposal", "roi_align_fwd", "roi_align_bwd")

    def __init__(self):
        self.events = []

    def mark(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def add(self, marks):
        self.events.append(marks)

    def mean_ms(self):
        torch.cuda.synchronize()
        out = {k: 0.0 for k in self.STAGES}
        for marks in self.events:
            for i, k in enumerate(self.STAGES):
                out[k] += marks[i].elapsed_time(marks[i + 1])
        n = max(1, len(self.events))
        return {k: v / n for k, v in out.items()}


class HostPipeline:
    def __init__(self, device, frames, channels, feat_h, feat_w, pooled, spatial_scale, pre_nms, post_nms, nms_thresh,
                 feat_stride=16, num_anchors=synth.NUM_ANCHORS):
        self.lib = load()
        self.dev = device
        self.B, self.C, self.H, self.W, self.P = frames, channels, feat_h, feat_w, pooled
        self.scale, self.pre, self.post, self.thr = float(spatial_scale), int(pre_nms), int(post_nms), float(nms_thresh)
        self.stride, self.A = int(feat_stride), int(num_anchors)
        self.N = frames * post_nms
        f32 = dict(dtype=torch.float32, device=device)
        self.anchors = torch.from_numpy(synth.BASE_ANCHORS).to(device)
        self.rois = torch.empty((frames, post_nms, 5), **f32)
        self.counts = torch.empty((frames,), dtype=torch.int32, device=device)
        self.pooled = torch.empty((self.N, channels, pooled, pooled), **f32)
        self.grad_in = torch.empty((frames, channels, feat_h, feat_w), **f32)
        self.ws_prop = torch.empty(self.lib.i2v_proposal_workspace_bytes(frames, self.A, feat_h, feat_w, pre_nms),
                                   dtype=torch.uint8, device=device)
        self.ws_roi = torch.empty(self.lib.i2v_roi_align_workspace_bytes(frames, self.N), dtype=torch.uint8,
                                  device=device)
        # kernels per step: decode, sort, nms | prep, bucket, forward | prep, (bucket,) backward (memsets not counted)
        self.launches_per_step = 9
        self._host = None

    # ------------------------------------------------------------------ device-resident inputs
    def device_step(self, cls_prob, bbox_pred, im_info, features, grad_out, timer: StageTimer | None = None):
        lib, s = self.lib, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        marks = [timer.mark()] if timer else None
        check(lib.i2v_proposal_forward(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(self.anchors), self.B, self.A,
                                       self.H, self.W, self.stride, self.pre, self.post, self.thr, _p(self.rois),
                                       _p(self.counts), _p(self.ws_prop), self.ws_prop.numel(), s), "proposal_forward")
        if timer:
            marks.append(timer.mark())
        check(lib.i2v_roi_align_forward(_p(features), _p(self.rois), _p(self.pooled), self.B, self.C, self.H, self.W,
                                        self.N, self.P, self.P, self.scale, POOL_AVG, IMPL_AUTO, _p(self.ws_roi),
                                        self.ws_roi.numel(), s), "roi_align_forward")
        if timer:
            marks.append(timer.mark())
        check(lib.i2v_roi_align_backward(_p(grad_out), None, _p(self.rois), _p(self.grad_in), self.B, self.C, self.H,
                                         self.W, self.N, self.P, self.P, self.scale, POOL_AVG, IMPL_AUTO,
                                         _p(self.ws_roi), self.ws_roi.numel(), s), "roi_align_backward")
        if timer:
            marks.append(timer.mark())
            timer.add(marks)
        return self.rois, self.pooled, self.grad_in

    # ------------------------------------------------------------------ pinned host inputs and outputs
    def _host_buffers(self, cls_prob, bbox_pred, im_info, features, grad_out):
        if self._host is None:
            d = {}
            for name, t in (("cls", cls_prob), ("reg", bbox_pred), ("info", im_info), ("feat", features),
                            ("grad", grad_out)):
                d[name] = torch.empty(t.shape, dtype=torch.float32, device=self.dev)
            d["rois_h"] = torch.empty(self.rois.shape, dtype=torch.float32).pin_memory()
            d["pooled_h"] = torch.empty(self.pooled.shape, dtype=torch.float32).pin_memory()
            d["grad_in_h"] = torch.empty(self.grad_in.shape, dtype=torch.float32).pin_memory()
            self._host = d
            self.h2d_bytes = sum(t.numel() * 4 for t in (cls_prob, bbox_pred, im_info, features, grad_out))
            self.d2h_bytes = sum(d[k].numel() * 4 for k in ("rois_h", "pooled_h", "grad_in_h"))
        return self._host

    def host_step(self, cls_prob, bbox_pred, im_info, features, grad_out):
        """Host buffers in, host buffers out: copies every input to the device, runs the step, copies every result
        (proposals, pooled features, feature gradient) back, and returns when they are in host memory."""
        d = self._host_buffers(cls_prob, bbox_pred, im_info, features, grad_out)
        d["cls"].copy_(cls_prob, non_blocking=True)
        d["reg"].copy_(bbox_pred, non_blocking=True)
        d["info"].copy_(im_info, non_blocking=True)
        d["feat"].copy_(features, non_blocking=True)
        d["grad"].copy_(grad_out, non_blocking=True)
        self.device_step(d["cls"], d["reg"], d["info"], d["feat"], d["grad"])
        d["rois_h"].copy_(self.rois, non_blocking=True)
        d["pooled_h"].copy_(self.pooled, non_blocking=True)
        d["grad_in_h"].copy_(self.grad_in, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return d["rois_h"], d["pooled_h"], d["grad_in_h"]

    h2d_bytes = 0
    d2h_bytes = 0
