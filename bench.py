#!/usr/bin/env python
"""Benchmark of the region-level hot path (BASELINE.json configs[1]).

One step = one pass over a batch of 32 synthetic VidVRD-shaped frames:
    proposal decode + sort + NMS (12000 pre-NMS -> 300 post-NMS per frame)
    -> RoIAlignAvg 7x7 forward over conv4 features [32,1024,38,63]   (9600 RoIs)
    -> RoIAlignAvg backward of an upstream gradient [9600,1024,7,7]
Metric: frames/s.  `value` is timed on the device with the inputs resident in HBM, the K steps software-pipelined over
batches (proposal layer of step i+1 on a second stream under the RoIAlign backward of step i; `--no-pipeline` for strictly
sequential stages); stage durations / roofline come from a stage-by-stage pass.  `e2e` goes through the host-buffer
pipeline (pinned host inputs copied in, all results copied out, every step).

    python bench.py --gpus N --steps K --warmup W                (torchrun for N > 1; frames are sharded, weak scaling)
    python bench.py --impl reference ...                         (the CPU port of the reference path, all host cores)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FRAMES = 32
PRE_NMS, POST_NMS, NMS_THRESH = 12000, 300, 0.7
CHANNELS, FEAT_H, FEAT_W, POOLED = 1024, 38, 63, 7
SCALE = 1.0 / 16
METRIC = "frames/sec RoIAlign+NMS+pair-feature path; RoIAlign HBM GB/s vs peak"
WORKLOAD = ("configs[1]: 32 VidVRD-shaped frames (600x1000 -> conv4 1024x38x63), 12000 pre-NMS / 300 post-NMS RPN "
            "proposals per frame, proposal decode + NMS + RoIAlignAvg 7x7 fwd + bwd")


def config_dict(world: int):
    """What the workload is -- identical in the GPU arm and the reference arm (the driver compares the two)."""
    return {"workload": WORKLOAD, "frames_per_gpu": FRAMES, "rois_per_gpu": FRAMES * POST_NMS,
            "l2": "inputs larger than L2 (314 MB features, 1.9 GB pooled tensor and gradient per step)",
            "parallelism": f"frames sharded over {world} rank(s), no data-path collective"}


def algorithmic_bytes(frames: int, rois: int):
    """every feature byte once + rois + the pooled tensor (fwd); pooled gradient read +
    feature gradient written (bwd)."""
    feat = frames * CHANNELS * FEAT_H * FEAT_W * 4
    pooled = rois * CHANNELS * POOLED * POOLED * 4
    return {"roi_align_fwd": feat + rois * 20 + pooled, "roi_align_bwd": pooled + feat + rois * 20}


def measured_traffic():
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), or {}."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_path(frames: int, threads: int, seed: int = 0):
    """The reference path on the host: oracle/ (C port, OpenMP) -- decode + sort + nms_cpu semantics, lattice
    RoIAlign 8x8 + avg_pool (RoIAlignAvg) forward and its backward.  Returns seconds for `frames` frames."""
    from i2vsgg_b200 import synth
    from oracle import oracle
    cls, reg = synth.rpn_outputs(seed, batch=frames)
    info = synth.im_info(frames)
    feat = synth.feature_map(seed, frames)
    grad = np.random.default_rng(seed).standard_normal((frames * POST_NMS, CHANNELS, POOLED, POOLED), dtype=np.float32)
    oracle.lib()
    t0 = time.perf_counter()
    rois = oracle.proposal_layer(cls, reg, info, PRE_NMS, POST_NMS, NMS_THRESH).reshape(-1, 5)
    oracle.roi_align_pooled_forward(feat, rois, POOLED, POOLED, SCALE, "avg", nthreads=threads)
    oracle.roi_align_pooled_backward(grad, feat, rois, POOLED, POOLED, SCALE, "avg", nthreads=threads)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    threads = oracle.default_threads()
    sample_frames = 8
    for _ in range(min(args.warmup, 2)):
        cpu_path(1, threads)
    times = [cpu_path(sample_frames, threads, seed=s) for s in range(args.steps)]
    sec = float(np.sum(times))
    value = sample_frames * args.steps / sec
    sample = (f"{sample_frames} frames per step (of the 32-frame batch), oracle/ C port of the reference path "
              f"(nms_cpu.py semantics, roi_align.c loop + RoIAlignAvg pool, backward per roi_align_kernel.cu) with "
              f"OpenMP over {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from i2vsgg_b200 import ops, synth
    from i2vsgg_b200.pipeline import HostPipeline, StageTimer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs of the named shape, built on the host from seeds (one set of frames per rank)
    cls_h, reg_h = synth.rpn_outputs(1000 + rank, batch=FRAMES)
    info_h = synth.im_info(FRAMES)
    g = torch.Generator().manual_seed(rank)
    feat_h = torch.randn((FRAMES, CHANNELS, FEAT_H, FEAT_W), generator=g).pin_memory()
    grad_h = torch.randn((FRAMES * POST_NMS, CHANNELS, POOLED, POOLED), generator=g).pin_memory()
    cls_h, reg_h, info_h = (torch.from_numpy(a).pin_memory() for a in (cls_h, reg_h, info_h))
    pipe = HostPipeline(dev, FRAMES, CHANNELS, FEAT_H, FEAT_W, POOLED, SCALE, PRE_NMS, POST_NMS, NMS_THRESH)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident run: inputs are in HBM before the timed region starts
    cls_d, reg_d, info_d, feat_d, grad_d = (t.to(dev, non_blocking=True) for t in (cls_h, reg_h, info_h, feat_h, grad_h))
    timer = StageTimer()
    for _ in range(args.warmup):
        pipe.device_step(cls_d, reg_d, info_d, feat_d, grad_d)
    # stage durations (and with them the roofline figures) come from a stage-by-stage pass on one stream, so that a
    # kernel's time is its own; it is not part of `value`
    stage_steps = max(3, min(args.steps, 10))
    for _ in range(stage_steps):
        pipe.device_step(cls_d, reg_d, info_d, feat_d, grad_d, timer)
    stage_ms = timer.mean_ms()
    seq_ms = sum(stage_ms.values())
    if not args.no_pipeline:
        for _ in range(2):                               # warm the second stream
            pipe.pipelined_step(cls_d, reg_d, info_d, feat_d, grad_d, (cls_d, reg_d, info_d))
        pipe.pipelined_step(cls_d, reg_d, info_d, feat_d, grad_d, None)
    barrier()
    sampler = ClockSampler(local)
    with sampler:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for k in range(args.steps):
            if args.no_pipeline:
                pipe.device_step(cls_d, reg_d, info_d, feat_d, grad_d)
            else:
                nxt = (cls_d, reg_d, info_d) if k + 1 < args.steps else None
                pipe.pipelined_step(cls_d, reg_d, info_d, feat_d, grad_d, nxt)
        end.record()
        barrier()
    ms_total = reduce_max(start.elapsed_time(end))
    launches = pipe.launches_per_step * args.steps

    # ---- end to end through host buffers
    # Two steps are kept in flight (`host_step_async`, results into alternating pinned buffers, step i awaited after step
    # i+1 was issued): every step still copies all its inputs in and all its results out inside the timed region, but the
    # first chunk in and the last chunk out of a step no longer have the link to themselves.
    e_steps = 0 if args.no_e2e else max(1, min(args.steps, 10))
    for _ in range(min(args.warmup, 2) if e_steps else 0):
        pipe.host_step(cls_h, reg_h, info_h, feat_h, grad_h, chunk_frames=args.e2e_chunk)
    if e_steps:                                                        # allocates the second set of pinned result buffers
        pipe.host_step_async(cls_h, reg_h, info_h, feat_h, grad_h, chunk_frames=args.e2e_chunk, slot=1)[3].synchronize()
    barrier()
    with sampler:
        start2, end2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start2.record()
        prev = None
        for k in range(e_steps):
            cur = pipe.host_step_async(cls_h, reg_h, info_h, feat_h, grad_h, chunk_frames=args.e2e_chunk, slot=k & 1)
            if prev is not None:
                prev[3].synchronize()                                  # step k-1's results are in host memory
            prev = cur
        if prev is not None:
            torch.cuda.current_stream().wait_event(prev[3])
        end2.record()
        if prev is not None:
            prev[3].synchronize()
        barrier()
    e2e_ms = reduce_max(start2.elapsed_time(end2))

    # ---- the host<->device ceiling of this box under the same load: every rank copies 512 MiB each way at the same time
    # (pinned buffers, both copy engines); the end-to-end step cannot beat max(bytes in, bytes out) / this rate
    ceiling = None
    if e_steps:
        nb = 512 << 20
        hp_in, hp_out = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        dv_in, dv_out = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
        c1, c2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def duplex(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            c1.wait_stream(torch.cuda.current_stream())
            c2.wait_stream(torch.cuda.current_stream())
            for _ in range(reps):
                with torch.cuda.stream(c1):
                    dv_in.copy_(hp_in, non_blocking=True)
                with torch.cuda.stream(c2):
                    hp_out.copy_(dv_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(c1)
            torch.cuda.current_stream().wait_stream(c2)
            b.record()
            torch.cuda.synchronize()
            return nb * reps / (a.elapsed_time(b) * 1e-3) / 1e9
        duplex(1)
        barrier()
        mine = duplex(4)
        if world > 1:
            t = torch.tensor([mine, -mine, mine], device=dev, dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            lo_rank, agg = -float(tmax[1].item()), float(t[2].item())
        else:
            lo_rank, agg = mine, mine
        need = max(pipe.h2d_bytes, pipe.d2h_bytes)
        ceiling = {"gbs_each_direction_slowest_rank": lo_rank, "gbs_each_direction_all_ranks": agg,
                   "frames_per_s_at_ceiling": FRAMES * world / (need / (lo_rank * 1e9)),
                   "how": "pinned host <-> device, 512 MiB each way at once on two streams, all ranks at the same time"}
        del hp_in, hp_out, dv_in, dv_out

    # ---- parity of the very buffers that were timed (outside the timed regions): the proposals of frames 0 and 17
    # against the oracle's proposal layer, three of their pooled rows against the oracle's RoIAlignAvg
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle
        pipe.device_step(cls_d, reg_d, info_d, feat_d, grad_d)
        torch.cuda.synchronize()
        fr = [0, 17]
        want = oracle.proposal_layer(cls_h[fr].numpy(), reg_h[fr].numpy(), info_h[fr].numpy(), PRE_NMS, POST_NMS, NMS_THRESH)
        got = pipe.rois[fr].cpu().numpy().copy()
        got[:, :, 0] = np.arange(len(fr))[:, None]
        rois_equal = bool(np.array_equal(got, want))
        rows = [5, 150, 299]
        sub = feat_h[17:18].numpy()
        r3 = want[1, rows].copy()
        r3[:, 0] = 0
        pw = oracle.roi_align_pooled_forward(sub, r3, POOLED, POOLED, SCALE, "avg", nthreads=oracle.default_threads())
        pg = pipe.pooled[[17 * POST_NMS + k for k in rows]].cpu().numpy()
        err = float(np.abs(pg - pw).max() / max(np.abs(pw).max(), 1e-30))
        parity = {"rois_equal_oracle": rois_equal, "pooled_max_rel_err": err, "frames_checked": fr, "rows_checked": rows}
        assert rois_equal and err <= 1e-5, f"the timed buffers differ from the oracle: {parity}"

    extra = {} if args.no_configs else other_configs(dev, rank, world, args.clip_frames)

    if rank == 0:
        alg = algorithmic_bytes(FRAMES, FRAMES * POST_NMS)
        top = max(("roi_align_fwd", "roi_align_bwd"), key=lambda k: stage_ms[k])
        peak, which = measured_peak()
        achieved = alg[top] / (stage_ms[top] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": FRAMES * world * args.steps / (ms_total * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world),
            "pipelining": ("none" if args.no_pipeline else
                           "the proposal layer of step i+1 runs on a second stream under the RoIAlign backward of step i; "
                           "every step runs all three stages"),
            "ms_per_step_stage_by_stage": seq_ms,
            "clocks": sampler.summary(),
            "e2e": {"value": FRAMES * world * e_steps / (e2e_ms * 1e-3) if e_steps else None, "unit": "frames/s", "steps": e_steps,
                    "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
                    "ms_per_step": e2e_ms / max(e_steps, 1), "ceiling": ceiling,
                    "how": "HostPipeline.host_step_async, %d frames per copy/compute chunk, two steps in flight "
                           "(results into alternating pinned buffers; every step copies all inputs in and all results "
                           "out inside the timed region)" % args.e2e_chunk},
            "gpu_launches": launches,
            "stages_ms": stage_ms,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic().get(top),
                         "traffic_source": measured_traffic().get("source"), "peak_source": which,
                         "algorithmic_bytes": alg[top],
                         "other": {k: {"achieved": alg[k] / (stage_ms[k] * 1e-3) / 1e9,
                                       "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / peak} for k in alg}},
        }
        line["parity"] = parity
        line.update(extra)
        if not args.no_projection:
            line["projection"] = projection_side_measurement(dev)
        if not args.no_cpu and world >= 1:
            from oracle import oracle
            threads = oracle.default_threads()
            sample_frames, reps = FRAMES, 2
            cpu_path(1, threads)
            sec = sum(cpu_path(sample_frames, threads, seed=1 + r) for r in range(reps))
            line["cpu_baseline"] = {"value": sample_frames * reps / sec, "unit": "frames/s", "cores": threads,
                                    "kind": "port",
                                    "sample": f"{reps} x {sample_frames} frames (the whole batch) of the same workload "
                                              f"through oracle/ (C port of the reference CPU path, OpenMP, {threads} "
                                              f"threads), {sec:.1f} s of CPU work"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(dev, rank: int, world: int, clip_frames_per_rank: int):
    """BASELINE.json configs[2], [3] and [4], measured after the headline step (none of this is part of `value`).

    sgg_frame  configs[2]: 64 detections -> 4032 ordered pairs: pair build + vrd.forward (union RoIPool, fc6/fc7/fc8 on
               tcgen05, fusion, cosine scores) + top-100, ms per frame through ClipRunner (4 frames per launch group)
    train_bwd  configs[3]: RoIAlignAvg backward, 8 images x 256 RoIs, gradient [2048,1024,7,7]
    clip       configs[4]: a (128 x ranks)-frame clip, frames sharded over the ranks: proposal decode + NMS -> RoIAlignAvg
               forward -> pair stage -> relation head -> top-100 -> ONE NCCL all-gather of the records -> temporal
               association on rank 0.  Every rank takes part (the all-gather is a collective)."""
    import torch
    import torch.distributed as dist
    from i2vsgg_b200 import ops, sgg, shard, synth
    from i2vsgg_b200.clip import ClipRunner
    from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    tf_peak, hbm_peak = float(peaks.get("bf16_tflops", 1590.0)), float(peaks.get("hbm_gbs", 6650.0))

    def timed(fn, warm=2, reps=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = {}
    vargs = synth.VrdArgs()
    head = vrd(vargs, None, synth.prd_vectors(7))
    head.load_state_dict({k: torch.from_numpy(v) for k, v in synth.vrd_params(1234, vargs).items()})
    head = head.to(dev).eval().prepare()
    det = 64
    group = 4
    runner = ClipRunner(head, synth.IM_H, synth.IM_W, group)
    fmap1 = torch.from_numpy(synth.feature_map(100 + rank, 1)).to(dev)

    # ---- configs[2]
    if rank == 0:
        boxes, classes, conf = synth.clip_detections(5, group, det)
        b = torch.from_numpy(boxes).to(dev)
        c = torch.from_numpy(np.tile(classes, (group, 1))).to(dev)
        s = torch.from_numpy(np.tile(conf, (group, 1))).to(dev)
        fm = fmap1.expand(group, -1, -1, -1).contiguous()
        ms = timed(lambda: runner._group(fm, b, c, s)) / group
        P, N, U = det * (det - 1), det, det * (det - 1) // 2
        flop = 2 * (U + N) * 50176 * 4096 + 2 * (U + N) * 4096 * 4096 + 2 * U * 4096 * 256 + 2 * N * 4096 * 300 \
            + 2 * P * (600 + 768) * 256 + 2 * P * 256 * 300 + 2 * P * 300 * 132 \
            + 2 * P * (256 * 50 * 96 + 64 * 2400 * 128 + 8192 * 64 + 64 * 256)
        out["sgg_frame"] = {"workload": "configs[2]: 64 detections -> 4032 ordered pairs, pair build + vrd.forward + top-100, "
                                        "4 frames per launch group", "ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                            "roofline": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "peak": tf_peak,
                                         "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / tf_peak,
                                         "flop_per_frame": flop,
                                         "note": "whole frame (pooling, small layers, selection included) against the bf16 peak; "
                                                 "flops counted with the unordered-pair shortcut"}}
        del fm
        # ---- configs[3]
        B4, per = 8, 256
        rois = torch.from_numpy(synth.rois(401, B4 * per, batch=B4, sort_by_batch=True)).to(dev)
        grad = torch.randn((B4 * per, CHANNELS, POOLED, POOLED), device=dev)
        ms4 = timed(lambda: ops.roi_align_backward(grad, None, rois, (B4, CHANNELS, FEAT_H, FEAT_W), POOLED, POOLED, SCALE,
                                                   "avg"), warm=3, reps=20)
        nbytes = grad.numel() * 4 + B4 * CHANNELS * FEAT_H * FEAT_W * 4 + rois.numel() * 4
        out["train_bwd"] = {"workload": "configs[3]: RoIAlignAvg backward, 8 images x 256 RoIs, 1024 x 38 x 63", "ms": ms4,
                            "roofline": {"bound": "hbm", "achieved": nbytes / (ms4 * 1e-3) / 1e9, "peak": hbm_peak,
                                         "unit": "GB/s", "frac": nbytes / (ms4 * 1e-3) / 1e9 / hbm_peak,
                                         "algorithmic_bytes": nbytes}}
        del grad
        # ---- the module-API ops the SGG model constructs: ROIPool((7,7), 1/16) forward + backward
        # (resnet_SGG_emb.py:82) and the model._C RoIAlign (faster_rcnn_SGG_emb.py:47), 8 frames x 300 RoIs
        from i2vsgg_b200._lib import ARGMAX_PLANE
        Bm, Nm = 8, 2400
        featm = torch.randn((Bm, CHANNELS, FEAT_H, FEAT_W), device=dev)
        roism = torch.from_numpy(synth.rois(402, Nm, batch=Bm, sort_by_batch=True)).to(dev)
        gradm = torch.randn((Nm, CHANNELS, POOLED, POOLED), device=dev)
        _, argm = ops.roi_pool_forward(featm, roism, POOLED, POOLED, SCALE, ARGMAX_PLANE)
        fbytes, pbytes = featm.numel() * 4, gradm.numel() * 4

        def entry(ms, nbytes):
            return {"ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                    "unit": "GB/s", "algorithmic_bytes": nbytes}
        out["module_ops"] = {
            "workload": "8 frames x 300 RoIs, 1024 x 38 x 63, 7x7 (fractions of the measured HBM copy bandwidth)",
            "roi_pool_fwd": entry(timed(lambda: ops.roi_pool_forward(featm, roism, POOLED, POOLED, SCALE, ARGMAX_PLANE), reps=10),
                                  fbytes + 2 * pbytes),
            "roi_pool_bwd": entry(timed(lambda: ops.roi_pool_backward(gradm, roism, argm, featm.shape, POOLED, POOLED, SCALE,
                                                                      ARGMAX_PLANE), reps=10), fbytes + 2 * pbytes),
            "c_roi_align_fwd": entry(timed(lambda: ops.c_roi_align_forward(featm, roism, POOLED, POOLED, SCALE, 0), reps=10),
                                     fbytes + pbytes),
            "c_roi_align_bwd": entry(timed(lambda: ops.c_roi_align_backward(gradm, roism, featm.shape, POOLED, POOLED, SCALE, 0),
                                           reps=10), fbytes + pbytes),
            "note": "model._C has no source in the reference tree: parity of these two ops is pinned to torchvision only",
        }
        del featm, gradm, argm

    # ---- configs[4]
    frames = clip_frames_per_rank * world
    lo, hi = shard.frame_range(frames, rank, world)
    boxes, classes, conf = synth.clip_detections(5, frames, det)
    b = torch.from_numpy(boxes[lo:hi]).to(dev)
    c = torch.from_numpy(np.tile(classes, (hi - lo, 1))).to(dev)
    s = torch.from_numpy(np.tile(conf, (hi - lo, 1))).to(dev)
    nd = 8                                            # distinct RPN frames per rank, cycled over its chunk of the clip
    cls_h, reg_h = synth.rpn_outputs(2000 + rank, batch=nd)
    reps = (hi - lo + nd - 1) // nd
    cls_d = torch.from_numpy(cls_h).to(dev).repeat(reps, 1, 1, 1)[: hi - lo]
    reg_d = torch.from_numpy(reg_h).to(dev).repeat(reps, 1, 1, 1)[: hi - lo]
    info_d = torch.from_numpy(synth.im_info(hi - lo)).to(dev)

    class View:                                       # one seeded map per rank, handed out group by group
        def __getitem__(self, sl):
            n = len(range(*sl.indices(hi - lo)))
            return fmap1.expand(n, -1, -1, -1).contiguous()

    runner.run_full(cls_d[:group], reg_d[:group], info_d[:group], View(), b[:group], c[:group], s[:group], group * world,
                    rank, world, pre_nms=PRE_NMS, post_nms=POST_NMS, nms_thresh=NMS_THRESH)        # warm-up, same collective
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = {}
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rec, cnt, kept = runner.run_full(cls_d, reg_d, info_d, View(), b, c, s, frames, rank, world, pre_nms=PRE_NMS,
                                     post_nms=POST_NMS, nms_thresh=NMS_THRESH, timings=tm)
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e), tm["gather_ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        t0 = time.perf_counter()
        rels = sgg.association(rec, cnt)
        assoc_ms = 1e3 * (time.perf_counter() - t0)
        ms = float(t[0].item())
        out["clip"] = {"workload": f"configs[4]: {frames}-frame synthetic VidVRD clip over {world} rank(s): proposal decode + NMS "
                                   f"({PRE_NMS} -> {POST_NMS}), RoIAlignAvg 7x7 forward, {det} detections -> {det * (det - 1)} pairs, "
                                   f"relation head, top-100 triplets, one all-gather of the records, temporal association on rank 0",
                       "frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3),
                       "gather": {"collective": "all_gather_into_tensor" if world > 1 else "none (one rank)",
                                  "backend": "nccl" if world > 1 else None, "us": 1e3 * float(t[1].item()),
                                  "gather_bytes": int(rec.numel() * 4 + cnt.numel() * 4)},
                       "records_ok": bool(rec.shape == (frames, 100, 13) and int(cnt.min()) == 100),
                       "proposals_kept_min": int(kept.min()) if kept.numel() else None,
                       "association": {"ms": assoc_ms, "relations": len(rels)}}
    return out


def projection_side_measurement(dev):
    """Outside the timed step: the relation head's dominant GEMM (fc6, 64 objects + 2016 distinct union boxes = 2080 rows
    x 50176 -> 4096, bf16 on tcgen05) against the measured cuBLAS bf16 peak.  CUDA events, 3 warm-up + 10 launches."""
    import torch
    from i2vsgg_b200 import ops
    m, n, k = 2080, 4096, 50176
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path)).get("bf16_tflops", 1590.0)) if os.path.exists(peaks_path) else 1590.0
    x = torch.randn((m, k), device=dev).bfloat16()
    w = (torch.randn((n, k), device=dev) * 0.01).bfloat16()
    b = torch.zeros((n,), device=dev)
    y = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(x, w, b, relu=True, out=y)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.linear(x, w, b, relu=True, out=y)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    # the same GEMM on fp32 operands (tcgen05 kind::tf32), what vrd(..., precision="tf32") runs
    x32, w32 = x.float(), w.float()
    y32 = torch.empty((m, n), device=dev, dtype=torch.float32)
    for _ in range(2):
        ops.linear(x32, w32, b, relu=True, out=y32)
    torch.cuda.synchronize()
    s.record()
    for _ in range(5):
        ops.linear(x32, w32, b, relu=True, out=y32)
    e.record()
    torch.cuda.synchronize()
    ms32 = s.elapsed_time(e) / 5
    tf32 = {"dtype": "tf32 (fp32 operands)", "ms": ms32, "tflops": 2.0 * m * n * k / (ms32 * 1e-3) / 1e12,
            "note": "no measured tf32 peak in MEASURED_PEAKS.json; nominal dense tf32 is half the bf16 rate"}
    del x32, w32, y32
    return {"tf32": tf32, "kernel": "linear_tcgen05_pair_kernel, cta_group::2 (fc6 of vrd.forward, SURVEY 8 a19)", "shape": [m, n, k], "dtype": "bf16",
            "ms": ms, "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                   "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if os.path.exists(peaks_path)
                                   else "fallback (B200_PROFILING.md)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--e2e-chunk", type=int, default=2, help="frames per copy/compute chunk of the host-buffer leg")
    ap.add_argument("--no-projection", action="store_true", help="skip the relation-head side measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] / [3] / [4] blocks")
    ap.add_argument("--clip-frames", type=int, default=128, help="frames per rank of the configs[4] clip")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="run the three stages of a step strictly one after the other (no second stream)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
