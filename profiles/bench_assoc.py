#!/usr/bin/env python
"""Temporal association of a clip's gathered triplet records (rank 0's last stage in configs[4]): the device kernel alone
and the whole `sgg.association` call (kernel + the host side that builds the reference's dictionaries)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from i2vsgg_b200 import ops, sgg, synth  # noqa: E402

for frames in (128, 1024):
    rec, cnt = synth.clip_records(3, frames=frames, tracks=40, clutter=60)
    rec_d = torch.from_numpy(rec).cuda()
    cnt_l = [int(c) for c in cnt]
    src = sgg._fill_empty_frames(cnt_l)
    fnos = list(range(frames))
    for _ in range(2):
        ops.greedy_association(rec_d, cnt_l, fnos, src, max_traj=100)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ops.greedy_association(rec_d, cnt_l, fnos, src, max_traj=100)
    e.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rels = sgg.association(rec_d, cnt)
    t1 = time.perf_counter()
    print(json.dumps({"frames": frames, "kernel_ms": a.elapsed_time(e) / 3, "association_ms": (t1 - t0) * 1e3,
                      "relations": len(rels)}))
