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
        self._rois_pp = [self.rois, torch.empty((frames, post_nms, 5), **f32)]   # ping-pong for the pipelined step
        self._side = None            # second stream: the next batch's proposal layer
        self._pending = None         # (buffer index, event) of a proposal launched ahead
        self.counts = torch.empty((frames,), dtype=torch.int32, device=device)
        self.pooled = torch.empty((self.N, channels, pooled, pooled), **f32)
        self.grad_in = torch.empty((frames, channels, feat_h, feat_w), **f32)
        self.ws_prop = torch.empty(self.lib.i2v_proposal_workspace_bytes(frames, self.A, feat_h, feat_w, pre_nms),
                                   dtype=torch.uint8, device=device)
        self.ws_roi = torch.empty(self.lib.i2v_roi_align_workspace_bytes(frames, self.N), dtype=torch.uint8,
                                  device=device)
        # kernels per step: decode, top-K order, nms, gated full sort, gated nms | prep, bucket, slab tables, forward |
        # prep, bucket, phase tables, backward (memsets not counted)
        self.launches_per_step = 13
        self._host = None
        # the backward of a step takes the forward's RoI tables (same RoIs, same workspace, back to back) when the two
        # plane-resident kernels apply to this shape: two launches fewer per step
        self._share_tables = (channels % 16 == 0 and pooled == 7 and (feat_h * feat_w) % 4 == 2
                              and feat_h * feat_w * 16 * 4 + 70 * 1024 <= 227 * 1024)
        if self._share_tables:
            self.launches_per_step -= 2

    # ------------------------------------------------------------------ device-resident inputs
    def _run(self, cls_prob, bbox_pred, im_info, features, grad_out, rois, pooled, grad_in, frames, timer=None):
        """proposal layer -> RoIAlignAvg forward -> backward for `frames` frames (all arguments are that many frames)."""
        lib, s = self.lib, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = frames * self.post
        marks = [timer.mark()] if timer else None
        check(lib.i2v_proposal_forward(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(self.anchors), frames, self.A,
                                       self.H, self.W, self.stride, self.pre, self.post, self.thr, _p(rois),
                                       _p(self.counts), _p(self.ws_prop), self.ws_prop.numel(), s), "proposal_forward")
        if timer:
            marks.append(timer.mark())
        check(lib.i2v_roi_align_forward(_p(features), _p(rois), _p(pooled), frames, self.C, self.H, self.W,
                                        n, self.P, self.P, self.scale, POOL_AVG, IMPL_AUTO, _p(self.ws_roi),
                                        self.ws_roi.numel(), s), "roi_align_forward")
        if timer:
            marks.append(timer.mark())
        # rois = NULL: the forward call just above left this batch's tables and per-frame lists in ws_roi
        check(lib.i2v_roi_align_backward(_p(grad_out), None, None if self._share_tables else _p(rois), _p(grad_in), frames,
                                         self.C, self.H, self.W, n, self.P, self.P, self.scale, POOL_AVG, IMPL_AUTO,
                                         _p(self.ws_roi), self.ws_roi.numel(), s), "roi_align_backward")
        if timer:
            marks.append(timer.mark())
            timer.add(marks)

    def device_step(self, cls_prob, bbox_pred, im_info, features, grad_out, timer: StageTimer | None = None):
        self._run(cls_prob, bbox_pred, im_info, features, grad_out, self.rois, self.pooled, self.grad_in, self.B, timer)
        return self.rois, self.pooled, self.grad_in

    # ------------------------------------------------------------------ the same step, software-pipelined over batches
    def _launch_proposal(self, cls_prob, bbox_pred, im_info, rois, stream):
        check(self.lib.i2v_proposal_forward(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(self.anchors), self.B, self.A,
                                            self.H, self.W, self.stride, self.pre, self.post, self.thr, _p(rois),
                                            _p(self.counts), _p(self.ws_prop), self.ws_prop.numel(),
                                            ctypes.c_void_p(stream.cuda_stream)), "proposal_forward")

    def pipelined_step(self, cls_prob, bbox_pred, im_info, features, grad_out, next_rpn=None):
        """One step whose proposal layer was (or is now) launched on a second stream, and which launches the NEXT
        batch's proposal layer (`next_rpn = (cls_prob, bbox_pred, im_info)`) on that stream before its own RoIAlign
        backward.  The proposal chain is latency-bound (1-4 CTAs per frame, 0.25 ms on at most 128 of 148 SMs); its CTAs slot in
        between the waves of the backward kernel instead of holding the whole GPU.  Every step still runs all three
        stages; only their placement in time changes.  Returns (rois, pooled, grad_in) of THIS step."""
        main = torch.cuda.current_stream()
        if self._side is None:
            # high priority: when an SM frees up under the backward kernel (one 227 KB CTA per SM, many waves) the
            # proposal chain's CTAs are placed before the backward's next CTA, so the chain really runs underneath it
            # instead of at its tail
            self._side = torch.cuda.Stream(self.dev, priority=-1)
        if self._pending is None:                      # first step of a run: nothing was launched ahead
            self._side.wait_stream(main)
            self._launch_proposal(cls_prob, bbox_pred, im_info, self._rois_pp[0], self._side)
            self._pending = (0, self._side.record_event())
        cur, ready = self._pending
        rois = self._rois_pp[cur]
        main.wait_event(ready)
        lib, s = self.lib, ctypes.c_void_p(main.cuda_stream)
        check(lib.i2v_roi_align_forward(_p(features), _p(rois), _p(self.pooled), self.B, self.C, self.H, self.W, self.N,
                                        self.P, self.P, self.scale, POOL_AVG, IMPL_AUTO, _p(self.ws_roi),
                                        self.ws_roi.numel(), s), "roi_align_forward")
        if next_rpn is not None:
            # the other roi buffer was last read by the step before this one, which the main stream has finished issuing;
            # order the side stream after that work before it overwrites the buffer
            fence = main.record_event()
            self._side.wait_event(fence)
            self._launch_proposal(next_rpn[0], next_rpn[1], next_rpn[2], self._rois_pp[cur ^ 1], self._side)
            self._pending = (cur ^ 1, self._side.record_event())
        else:
            self._pending = None
        check(lib.i2v_roi_align_backward(_p(grad_out), None, None if self._share_tables else _p(rois), _p(self.grad_in),
                                         self.B, self.C, self.H, self.W, self.N, self.P, self.P, self.scale, POOL_AVG,
                                         IMPL_AUTO, _p(self.ws_roi), self.ws_roi.numel(), s), "roi_align_backward")
        return rois, self.pooled, self.grad_in

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
            d["s_in"], d["s_out"] = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
            self._host = d
            self.h2d_bytes = sum(t.numel() * 4 for t in (cls_prob, bbox_pred, im_info, features, grad_out))
            self.d2h_bytes = sum(d[k].numel() * 4 for k in ("rois_h", "pooled_h", "grad_in_h"))
        return self._host

    def host_step(self, cls_prob, bbox_pred, im_info, features, grad_out, chunk_frames: int = 4):
        """Host buffers in, host buffers out: every input is copied to the device, the step runs, every result
        (proposals, pooled features, feature gradient) is copied back; returns when they are in host memory.

        Frames are independent, so the batch moves in chunks of `chunk_frames`: chunk k+1 is on its way in (copy
        engine 1) while chunk k computes and chunk k-1 is on its way out (copy engine 2).  The PCIe link is full duplex,
        which makes the step cost max(bytes in, bytes out) / link instead of their sum."""
        out = self.host_step_async(cls_prob, bbox_pred, im_info, features, grad_out, chunk_frames, slot=0)
        out[3].synchronize()
        return out[:3]

    def host_step_async(self, cls_prob, bbox_pred, im_info, features, grad_out, chunk_frames: int = 4, slot: int = 0):
        """`host_step` without the final wait: returns (rois_h, pooled_h, grad_in_h, event); the results are in the host
        buffers of `slot` (0 or 1) once `event.synchronize()` returns.  A caller that keeps two steps in flight -- issue
        step i+1 into the other slot, then wait for step i -- lets the inputs of a step travel in while the results of
        the step before are still on their way out, so that the link's two directions stay busy across step boundaries
        (the first chunk in and the last chunk out of a step are otherwise alone on the link).  Device buffers are
        shared between the steps; per-chunk events keep a chunk's input from being overwritten before its compute has
        read it, and its result from being overwritten before the copy out of the step before has read it."""
        d = self._host_buffers(cls_prob, bbox_pred, im_info, features, grad_out)
        if slot and "rois_h1" not in d:
            for k in ("rois_h", "pooled_h", "grad_in_h"):
                d[k + "1"] = torch.empty(d[k].shape, dtype=torch.float32).pin_memory()
        sfx = "1" if slot else ""
        rois_h, pooled_h, grad_in_h = d["rois_h" + sfx], d["pooled_h" + sfx], d["grad_in_h" + sfx]
        main, s_in, s_out = torch.cuda.current_stream(), d["s_in"], d["s_out"]
        computed, copied = d.setdefault("computed", {}), d.setdefault("copied", {})
        if not computed:            # first use: order the copy streams after whatever the caller did before
            s_in.wait_stream(main)
            s_out.wait_stream(main)
        post = self.post
        for f0 in range(0, self.B, chunk_frames):
            f1 = min(self.B, f0 + chunk_frames)
            r0, r1 = f0 * post, f1 * post
            with torch.cuda.stream(s_in):
                if f0 in computed:
                    s_in.wait_event(computed[f0])          # the step before has read this chunk's device inputs
                d["cls"][f0:f1].copy_(cls_prob[f0:f1], non_blocking=True)
                d["reg"][f0:f1].copy_(bbox_pred[f0:f1], non_blocking=True)
                d["info"][f0:f1].copy_(im_info[f0:f1], non_blocking=True)
                d["feat"][f0:f1].copy_(features[f0:f1], non_blocking=True)
                d["grad"][r0:r1].copy_(grad_out[r0:r1], non_blocking=True)
                ready = s_in.record_event()
            main.wait_event(ready)
            if f0 in copied:
                main.wait_event(copied[f0])                # the step before has copied this chunk's results out
            self._run(d["cls"][f0:f1], d["reg"][f0:f1], d["info"][f0:f1], d["feat"][f0:f1], d["grad"][r0:r1],
                      self.rois[f0:f1], self.pooled[r0:r1], self.grad_in[f0:f1], f1 - f0)
            if f0:      # frame indices of the whole batch, as device_step writes them
                check(self.lib.i2v_rois_add_frame(_p(self.rois[f0:f1]), (f1 - f0) * post, f0,
                                                  ctypes.c_void_p(main.cuda_stream)), "rois_add_frame")
            done = computed[f0] = main.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                rois_h[f0:f1].copy_(self.rois[f0:f1], non_blocking=True)
                pooled_h[r0:r1].copy_(self.pooled[r0:r1], non_blocking=True)
                grad_in_h[f0:f1].copy_(self.grad_in[f0:f1], non_blocking=True)
                copied[f0] = s_out.record_event()
        return rois_h, pooled_h, grad_in_h, copied[f0]

    h2d_bytes = 0
    d2h_bytes = 0
