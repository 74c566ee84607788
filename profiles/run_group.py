#!/usr/bin/env python
"""One ClipRunner frame group (4 frames x 64 detections: pair stage + relation head + top-100) for launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/run_group.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from i2vsgg_b200 import synth  # noqa: E402
from i2vsgg_b200.clip import ClipRunner  # noqa: E402
from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd  # noqa: E402

args = synth.VrdArgs()
head = vrd(args, None, synth.prd_vectors(7))
head.load_state_dict({k: torch.from_numpy(v) for k, v in synth.vrd_params(1234, args).items()})
head = head.cuda().eval().prepare()
F, det = int(os.environ.get("GROUP", "4")), 64
boxes, classes, conf = synth.clip_detections(5, F, det)
b = torch.from_numpy(boxes).cuda()
c = torch.from_numpy(np.tile(classes, (F, 1))).cuda()
s = torch.from_numpy(np.tile(conf, (F, 1))).cuda()
fm = torch.from_numpy(synth.feature_map(100, 1)).cuda().expand(F, -1, -1, -1).contiguous()
runner = ClipRunner(head, synth.IM_H, synth.IM_W, F)
for _ in range(int(os.environ.get("REPS", "3"))):
    runner._group(fm, b, c, s)
torch.cuda.synchronize()
if os.environ.get("TIME"):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        runner._group(fm, b, c, s)
    e.record()
    torch.cuda.synchronize()
    print("ms per frame", a.elapsed_time(e) / 10 / F)
