#!/usr/bin/env python
"""RoIAlignAvg forward on config 2 (32 frames x 300 proposals, 1024 x 38 x 63), every plane-resident kernel timed alone
(CUDA events, 5 warm-up + 30 launches, inputs resident, 2.2 GB of outputs per launch > L2)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from i2vsgg_b200 import ops, synth  # noqa: E402


def main():
    impls = sys.argv[1:] or ["slab", "chan", "even", "plane"]
    B, C, H, W = 32, 1024, 38, 63
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cls, reg = synth.rpn_outputs(0, batch=B)
    rois = ops.proposal_forward(cu(cls), cu(reg), cu(synth.im_info(B)), cu(synth.BASE_ANCHORS), 16, 12000, 300,
                                0.7).reshape(-1, 5)
    N = rois.size(0)
    feat = torch.randn((B, C, H, W), device="cuda")
    nbytes = N * C * 49 * 4 + B * C * H * W * 4 + N * 20
    ref = None
    for impl in impls:
        fn = lambda: ops.roi_align_forward(feat, rois, 7, 7, 1 / 16, "avg", impl)
        out = fn()
        if ref is None:
            ref = out
        err = float((out - ref).abs().max())
        del out
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(30):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 30
        print(json.dumps({"impl": impl, "ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "max_abs_diff_vs_first": err}))


if __name__ == "__main__":
    main()
