#!/usr/bin/env python
"""Times the tcgen05 FC kernel on the projection shapes of SURVEY.md 8(d) (config 3: 64 objects + 4032 union boxes =
4096 rows; fc6 50176->4096, fc7 4096->4096, fc8 4096->256) and prints one JSON line per shape with TFLOP/s and the
fraction of the measured cuBLAS bf16 peak (MEASURED_PEAKS.json).  Inputs are random bf16, resident in HBM; CUDA
events on the launching stream, 3 warm-up + N timed launches.

    python profiles/bench_projection.py [--iters 10] [--tf32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from i2vsgg_b200 import ops  # noqa: E402

SHAPES = [("fc6", 4096, 4096, 50176), ("fc7", 4096, 4096, 4096), ("fc8", 4032, 256, 4096),
          ("fc_fusion", 4032, 256, 768)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    peaks = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peaks = json.load(open(path))
    peak = float(peaks.get("bf16_tflops", 1590.0))
    dt = torch.float32 if args.tf32 else torch.bfloat16
    for name, m, n, k in SHAPES:
        if args.only and name != args.only:
            continue
        x = torch.randn((m, k), device="cuda").to(dt)
        w = (torch.randn((n, k), device="cuda") * 0.02).to(dt)
        b = torch.randn((n,), device="cuda")
        y = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.linear(x, w, b, relu=True, out=y)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.iters):
            ops.linear(x, w, b, relu=True, out=y)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.iters
        tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
        print(json.dumps({"layer": name, "M": m, "N": n, "K": k, "dtype": "tf32" if args.tf32 else "bf16",
                          "ms": ms, "tflops": tf, "peak_tflops": peak, "frac_of_measured_bf16": tf / peak}), flush=True)
        del x, w, y


if __name__ == "__main__":
    main()
