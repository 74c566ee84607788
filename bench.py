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
            "config": {"workload": WORKLOAD, "frames_per_step": sample_frames},
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
    e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 2) if e_steps else 0):
        pipe.host_step(cls_h, reg_h, info_h, feat_h, grad_h, chunk_frames=args.e2e_chunk)
    barrier()
    with sampler:
        start2, end2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start2.record()
        for _ in range(e_steps):
            pipe.host_step(cls_h, reg_h, info_h, feat_h, grad_h, chunk_frames=args.e2e_chunk)
        end2.record()
        barrier()
    e2e_ms = reduce_max(start2.elapsed_time(end2))

    if rank == 0:
        alg = algorithmic_bytes(FRAMES, FRAMES * POST_NMS)
        top = max(("roi_align_fwd", "roi_align_bwd"), key=lambda k: stage_ms[k])
        peak, which = measured_peak()
        achieved = alg[top] / (stage_ms[top] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": FRAMES * world * args.steps / (ms_total * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": FRAMES, "rois_per_gpu": FRAMES * POST_NMS,
                       "l2": "inputs larger than L2 (314 MB features, 1.9 GB pooled tensor and gradient per step)",
                       "parallelism": f"frames sharded over {world} rank(s), no data-path collective",
                       "pipelining": ("none" if args.no_pipeline else
                                      "the proposal layer of step i+1 runs on a second stream under the RoIAlign backward of "
                                      "step i; every step runs all three stages"),
                       "ms_per_step_stage_by_stage": seq_ms},
            "clocks": sampler.summary(),
            "e2e": {"value": FRAMES * world * e_steps / (e2e_ms * 1e-3) if e_steps else None, "unit": "frames/s", "steps": e_steps,
                    "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
                    "ms_per_step": e2e_ms / max(e_steps, 1)},
            "gpu_launches": launches,
            "stages_ms": stage_ms,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic().get(top), "peak_source": which,
                         "algorithmic_bytes": alg[top],
                         "other": {k: {"achieved": alg[k] / (stage_ms[k] * 1e-3) / 1e9,
                                       "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / peak} for k in alg}},
        }
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
    return {"kernel": "linear_tcgen05_pair_kernel, cta_group::2 (fc6 of vrd.forward, SURVEY 8 a19)", "shape": [m, n, k], "dtype": "bf16",
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
    ap.add_argument("--no-pipeline", action="store_true",
                    help="run the three stages of a step strictly one after the other (no second stream)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
