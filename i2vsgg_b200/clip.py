"""A clip through the SGG stage, sharded over ranks (BASELINE.json configs[4]).

Each rank takes a contiguous chunk of the clip's frames (`shard.frame_range`) and runs, a few frames per launch group:
pair enumeration + union boxes + dual masks (`sgg.build_pairs`), the relation head (`vrd.forward` with the
unordered-pair shortcut) and the per-frame top-100 triplet selection (`ops.triplet_topk`); the only exchange is the
all-gather of the [frames, 100, 13] records into video order (`shard.all_gather_triplets`), after which the temporal
association of lib/utils.py:461-526 can run on any rank.  Mirrors the per-frame loop of test_net_SGG_emb.py:150-215.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, sgg, shard


class ClipRunner:
    def __init__(self, head, im_h: float, im_w: float, frames_per_group: int = 4, top_k: int = shard.TOP_K,
                 graphs: bool = False):
        """`graphs`: capture the ~150 launches of a frame group in a CUDA graph per (frames, detections) shape and replay it
        for every later group of that shape (the group is launch-bound on the host otherwise)."""
        self.head, self.im_h, self.im_w = head, float(im_h), float(im_w)
        self.group, self.top_k = int(frames_per_group), int(top_k)
        self.graphs = bool(graphs)
        self._captured = {}
        self._side = None
        self._index = {}

    def _group_replayed(self, fmap, boxes, classes, conf):
        """`_group` through a CUDA graph: static input buffers are overwritten, the graph replayed, the outputs copied out."""
        key = (tuple(fmap.shape), tuple(boxes.shape), fmap.device.index)
        entry = self._captured.get(key)
        if entry is None:
            static = [t.clone() for t in (fmap, boxes, classes, conf)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # lazy initialisation (caches, function attributes) first
                for _ in range(2):
                    self._group(*static)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._group(*static)
            entry = self._captured[key] = (graph, static, out)
        graph, static, out = entry
        for dst, src in zip(static, (fmap, boxes, classes, conf)):
            dst.copy_(src)
        graph.replay()
        return out[0].clone(), out[1].clone()

    def _group(self, fmap, boxes, classes, conf):
        """fmap [F,C,H,W], boxes [F,N,4], classes [F,N], conf [F,N] (device) -> records [F,top_k,13], counts [F].

        No per-frame Python loop: ONE launch builds the pair lists, union boxes and object masks of the whole group
        (`ops.pair_build_frames`; the [P,2,32,32] pair masks of faster_rcnn_SGG_emb.py:654-655 are never written, the head
        takes one mask per object), the relation head runs once over the group's rows, and three launches select the top
        triplets of all frames (`ops.triplet_topk_frames`)."""
        F, N = boxes.shape[:2]
        dev = fmap.device
        P = N * (N - 1)
        key = (F, N, dev.index)
        if key not in self._index:      # per group shape: the frame column of the RoIs and the unordered-pair index maps
            rep1, inv1 = sgg.unordered_pairs(N, dev)
            offs = torch.arange(F, device=dev)
            self._index[key] = (torch.arange(F, device=dev, dtype=torch.float32).repeat_interleave(N)[:, None].contiguous(),
                                (rep1[None, :] + offs[:, None] * P).reshape(-1).contiguous(),
                                (inv1[None, :] + offs[:, None] * rep1.numel()).reshape(-1).contiguous())
        frame_col, rep, inv = self._index[key]
        boxes = boxes.contiguous()
        ixs, ixo, rel, obj_masks = ops.pair_build_frames(boxes, self.im_h, self.im_w)
        rois = torch.cat([frame_col, boxes.reshape(F * N, 4)], 1)
        scores, _ = self.head(fmap, rois, rel, None, None, ixs, ixo, return_numpy=False, rel_unique=(rep, inv),
                              obj_masks=obj_masks)
        return ops.triplet_topk_frames(scores, conf, classes, boxes, ixs[:P], ixo[:P], self.top_k)

    def run_full(self, rpn_cls, rpn_reg, im_info, fmaps, boxes, classes, conf, num_frames: int, rank: int = 0,
                 world: int = 1, group=None, pre_nms: int = 12000, post_nms: int = 300, nms_thresh: float = 0.7,
                 spatial_scale: float = 1.0 / 16, timings: dict = None):
        """BASELINE.json configs[4] end to end for this rank's frames: per frame group the RPN outputs (`rpn_cls`
        [f,2A,H,W], `rpn_reg` [f,4A,H,W], `im_info` [f,3]) go through proposal decode + NMS (`ops.proposal_forward`), the
        proposals through RoIAlignAvg 7x7 over `fmaps` (`ops.roi_align_forward`; the pooled rows are what the detector
        head, out of scope here, would classify), and the frame's detections (`boxes`, `classes`, `conf`) through the pair
        stage, the relation head and the top-100 selection (`_group`); then ONE all-gather puts the records of the whole
        clip on every rank.  Returns (records [num_frames,top_k,13], counts [num_frames], proposals kept per frame).
        `timings`, if given, receives the CUDA-event duration of the all-gather alone ("gather_ms")."""
        from . import synth
        lo, hi = shard.frame_range(num_frames, rank, world)
        assert boxes.shape[0] == hi - lo, "pass exactly the frames of shard.frame_range(num_frames, rank, world)"
        dev = boxes.device
        anchors = torch.from_numpy(synth.BASE_ANCHORS).to(dev)
        recs, cnts, kept = [], [], []
        # Software pipelining over frame groups: the detector side of group g+1 (proposal decode + NMS, RoIAlignAvg) is
        # issued on a high-priority side stream before the relation stage of group g on the main stream, which waits only
        # for the detector side of ITS group (in the full model the detector head, out of scope here, sits on that edge).
        # The proposal chain is latency-bound on a few SMs and slots in under the relation stage's GEMMs.
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(dev, priority=-1)
        side = self._side
        side.wait_stream(main)
        starts = list(range(0, hi - lo, self.group))

        def detector(f0):
            f1 = min(hi - lo, f0 + self.group)
            with torch.cuda.stream(side):
                fm = fmaps[f0:f1]
                rois, nkeep = ops.proposal_forward(rpn_cls[f0:f1], rpn_reg[f0:f1], im_info[f0:f1], anchors, 16, pre_nms,
                                                   post_nms, nms_thresh, return_counts=True)
                ops.roi_align_forward(fm, rois.reshape(-1, 5), 7, 7, spatial_scale, "avg")
                return nkeep, side.record_event()

        pending = detector(starts[0]) if starts else None
        for gi, f0 in enumerate(starts):
            f1 = min(hi - lo, f0 + self.group)
            nkeep, ready = pending
            pending = detector(starts[gi + 1]) if gi + 1 < len(starts) else None
            main.wait_event(ready)
            step = self._group_replayed if self.graphs else self._group
            r, c = step(fmaps[f0:f1], boxes[f0:f1], classes[f0:f1], conf[f0:f1])
            recs.append(r)
            cnts.append(c)
            kept.append(nkeep)
        main.wait_stream(side)
        rec = torch.cat(recs) if recs else torch.empty((0, self.top_k, shard.RECORD_WIDTH), device=dev)
        cnt = torch.cat(cnts) if cnts else torch.empty((0,), dtype=torch.int32, device=dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = shard.all_gather_triplets(rec, cnt, num_frames, group)
        b.record()
        if timings is not None:
            torch.cuda.synchronize()
            timings["gather_ms"] = a.elapsed_time(b)
        return out[0], out[1], (torch.cat(kept) if kept else torch.empty((0,), dtype=torch.int32, device=dev))

    def run(self, fmaps, boxes, classes, conf, num_frames: int, rank: int = 0, world: int = 1, group=None):
        """This rank's frames (`fmaps` [f,C,H,W], `boxes` [f,N,4], `classes` [f,N], `conf` [f,N], in clip order)
        -> (records [num_frames, top_k, 13], counts [num_frames]) of the WHOLE clip on every rank."""
        lo, hi = shard.frame_range(num_frames, rank, world)
        assert fmaps.shape[0] == hi - lo, "pass exactly the frames of shard.frame_range(num_frames, rank, world)"
        recs, cnts = [], []
        for f0 in range(0, hi - lo, self.group):
            f1 = min(hi - lo, f0 + self.group)
            step = self._group_replayed if self.graphs else self._group
            r, c = step(fmaps[f0:f1], boxes[f0:f1], classes[f0:f1], conf[f0:f1])
            recs.append(r)
            cnts.append(c)
        dev = fmaps.device
        rec = torch.cat(recs) if recs else torch.empty((0, self.top_k, shard.RECORD_WIDTH), device=dev)
        cnt = torch.cat(cnts) if cnts else torch.empty((0,), dtype=torch.int32, device=dev)
        return shard.all_gather_triplets(rec, cnt, num_frames, group)
