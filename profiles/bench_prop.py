"""Times the proposal stage (decode + sort + NMS) of the config the bench line is quoted on.  Usage: bench_prop.py [frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from i2vsgg_b200 import ops, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
cls_h, reg_h = synth.rpn_outputs(1000, batch=frames)
cls, reg, info = (torch.from_numpy(x).to(dev) for x in (cls_h, reg_h, synth.im_info(frames)))
anchors = torch.from_numpy(synth.BASE_ANCHORS).to(dev)


def step():
    return ops.proposal_forward(cls, reg, info, anchors, 16, 12000, 300, 0.7, return_counts=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    rois, kept = step()
b.record()
torch.cuda.synchronize()
print(json.dumps({"frames": frames, "ms": a.elapsed_time(b) / 20, "kept_min": int(kept.min()), "kept_max": int(kept.max())}))
