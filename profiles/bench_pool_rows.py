#!/usr/bin/env python
"""The relation head's RoIPool rows (fc6's A operand) and cosine-score kernels on one frame group (4 frames x 64
detections: 256 object boxes + 8064 union boxes), each timed alone with CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from i2vsgg_b200 import ops, sgg, synth  # noqa: E402

F, det = 4, 64
boxes, classes, conf = synth.clip_detections(5, F, det)
b = torch.from_numpy(boxes).cuda().contiguous()
fm = torch.from_numpy(synth.feature_map(100, 1)).cuda().expand(F, -1, -1, -1).contiguous()
ixs, ixo, rel, obj_masks = ops.pair_build_frames(b, synth.IM_H, synth.IM_W)
P = det * (det - 1)
rep1, inv1 = sgg.unordered_pairs(det, b.device)
offs = torch.arange(F, device=b.device)
rep = (rep1[None, :] + offs[:, None] * P).reshape(-1)
rois = torch.cat([torch.arange(F, device=b.device, dtype=torch.float32).repeat_interleave(det)[:, None], b.reshape(F * det, 4)], 1)
allb = torch.cat((rois, rel.index_select(0, rep)))
out = torch.empty((allb.shape[0], fm.shape[1] * 49), dtype=torch.bfloat16, device=b.device)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


res = {"rows": int(allb.shape[0]), "pool_ms": timed(lambda: ops.roi_pool_rows(fm, allb, 7, 7, 1.0 / 16, out=out))}
x = torch.randn((F * P, 300), device="cuda")
prd = torch.randn((132, 300), device="cuda")
res["rel_scores_ms"] = timed(lambda: ops.rel_scores(x, prd, softmax=True))
print(json.dumps(res))
