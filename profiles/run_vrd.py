#!/usr/bin/env python
"""Config 3 (BASELINE.json configs[2]) through `vrd.forward`: 64 detections -> 4032 ordered pairs, union-box RoIPool,
fc6/fc7/fc8 + fusion + cosine scores on one B200.  Prints one JSON line: ms per frame (CUDA events, inputs resident),
the projection's FLOP count and TFLOP/s.   python profiles/run_vrd.py [--iters 5]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from i2vsgg_b200 import sgg, synth  # noqa: E402
from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--det", type=int, default=64)
    ap.add_argument("--full", action="store_true", help="pool / fc6-fc8 every ordered pair (no unordered-pair shortcut)")
    a = ap.parse_args()
    args = synth.VrdArgs()
    net = vrd(args, None, synth.prd_vectors(7))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.vrd_params(1234, args).items()})
    net = net.cuda().eval().prepare()
    fmap = torch.from_numpy(synth.feature_map(41, 1)).cuda()
    det, classes, conf = synth.detections(42, a.det)
    boxes = torch.from_numpy(np.concatenate([np.zeros((a.det, 1), np.float32), det], 1)).cuda()

    def frame():
        ixs, ixo, rel, masks = sgg.build_pairs(torch.from_numpy(det).cuda(), synth.IM_H, synth.IM_W)
        uniq = None if a.full else sgg.unordered_pairs(a.det)
        om = None if a.full else masks[torch.arange(a.det, device="cuda") * (a.det - 1), 0]
        return net(fmap, boxes, rel, masks, classes, ixs, ixo, return_numpy=False, rel_unique=uniq, obj_masks=om)

    for _ in range(2):
        frame()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        frame()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.iters
    P, N = a.det * (a.det - 1), a.det
    U = P if a.full else P // 2          # rows that really go through fc6 / fc7 / fc8
    flop = 2 * (U + N) * 50176 * 4096 + 2 * (U + N) * 4096 * 4096 + 2 * U * 4096 * 256 + 2 * N * 4096 * 300 \
        + 2 * P * (600 + 768) * 256 + 2 * P * 256 * 300 + 2 * P * 300 * 132 \
        + 2 * P * (256 * 50 * 96 + 64 * 2400 * 128 + 8192 * 64 + 64 * 256)
    print(json.dumps({"workload": "configs[2]: %d detections -> %d pairs, vrd.forward" % (N, P), "ms_per_frame": ms,
                      "frames_per_s": 1e3 / ms, "unordered_pair_shortcut": not a.full, "tflop_per_frame": flop / 1e12, "tflops": flop / (ms * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
