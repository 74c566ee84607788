#!/usr/bin/env python
"""The module-API ops the SGG model constructs (ROIPool((7,7),1/16) forward / backward, model._C RoIAlign) at 8 frames x 300
RoIs x 1024 channels, CUDA events.  I2V_POOL_PER_ELEMENT=1 selects the per-element RoIPool kernels for comparison."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from i2vsgg_b200 import ops, synth  # noqa: E402
from i2vsgg_b200._lib import ARGMAX_FLAT, ARGMAX_PLANE  # noqa: E402


def timed(fn, warm=3, reps=10):
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


def main():
    B, N, C, H, W = 8, 2400, 1024, 38, 63
    feat = torch.randn((B, C, H, W), device="cuda")
    rois = torch.from_numpy(synth.rois(402, N, batch=B, sort_by_batch=True)).cuda()
    grad = torch.randn((N, C, 7, 7), device="cuda")
    out = {"per_element": bool(os.environ.get("I2V_POOL_PER_ELEMENT"))}
    for name, mode in (("flat", ARGMAX_FLAT), ("plane", ARGMAX_PLANE)):
        _, arg = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16, mode)
        out[f"roi_pool_fwd_{name}_ms"] = timed(lambda: ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16, mode))
        out[f"roi_pool_bwd_{name}_ms"] = timed(lambda: ops.roi_pool_backward(grad, rois, arg, feat.shape, 7, 7, 1 / 16, mode))
    out["roi_pool_rows_bf16_ms"] = timed(lambda: ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16))
    out["c_roi_align_fwd_ms"] = timed(lambda: ops.c_roi_align_forward(feat, rois, 7, 7, 1 / 16, 0))
    out["c_roi_align_bwd_ms"] = timed(lambda: ops.c_roi_align_backward(grad, rois, feat.shape, 7, 7, 1 / 16, 0))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
