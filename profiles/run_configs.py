#!/usr/bin/env python
"""The other BASELINE.json configurations on one B200 (CUDA events, 5 warm-up + 50 timed launches, inputs resident):
configs[0] single frame, 300 proposals: NMS@0.7 of 6000 candidates + RoIAlign 7x7 forward;
configs[3] training step of the instance discriminator: RoIAlignAvg backward, 8 images x 256 RoIs.
One JSON line each with GB/s against the measured HBM copy bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from i2vsgg_b200 import ops, synth  # noqa: E402


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # ---- configs[0]
    cls, reg = synth.rpn_outputs(0, batch=1)
    c, r, info, anc = cu(cls), cu(reg), cu(synth.im_info(1)), cu(synth.BASE_ANCHORS)
    feat = cu(synth.feature_map(0, 1))
    rois = ops.proposal_forward(c, r, info, anc, 16, 6000, 300, 0.7).reshape(-1, 5)
    ms_p = timed(lambda: ops.proposal_forward(c, r, info, anc, 16, 6000, 300, 0.7))
    ms_f = timed(lambda: ops.roi_align_forward(feat, rois, 7, 7, 1 / 16, "avg"))
    b0 = 1024 * 38 * 63 * 4 + 300 * 20 + 300 * 1024 * 49 * 4
    print(json.dumps({"config": "configs[0]: 1 frame, 6000 -> 300 proposals, RoIAlignAvg 7x7 forward",
                      "proposal_ms": ms_p, "roi_align_fwd_ms": ms_f, "frames_per_s": 1e3 / (ms_p + ms_f),
                      "fwd_gbs": b0 / ms_f / 1e6, "fwd_frac_of_measured_hbm": b0 / ms_f / 1e6 / peak}))
    # ---- configs[3]
    B, N = 8, 2048
    r4 = synth.rois(4, N, batch=B, sort_by_batch=True)
    g = torch.randn((N, 1024, 7, 7), device="cuda")
    rr = cu(r4)
    ms_b = timed(lambda: ops.roi_align_backward(g, None, rr, (B, 1024, 38, 63), 7, 7, 1 / 16, "avg"), 20)
    b3 = N * 1024 * 49 * 4 + B * 1024 * 38 * 63 * 4 + N * 20
    print(json.dumps({"config": "configs[3]: RoIAlignAvg backward, 8 images x 256 RoIs", "roi_align_bwd_ms": ms_b,
                      "images_per_s": B * 1e3 / ms_b, "bwd_gbs": b3 / ms_b / 1e6,
                      "bwd_frac_of_measured_hbm": b3 / ms_b / 1e6 / peak}))


if __name__ == "__main__":
    main()
