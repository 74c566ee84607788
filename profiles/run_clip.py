#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): a synthetic VidVRD clip through pair build + relation head + triplet top-k,
frames sharded over the ranks of one box, one NCCL all-gather of the per-frame records at the end.

    python profiles/run_clip.py --frames 128                                   (1 GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/run_clip.py --frames 1024

Prints one JSON line from rank 0: frames/s of the whole job (CUDA events, max over ranks, feature maps resident in HBM).
Feature maps are one seeded map per rank re-used for its frames (a 1024-frame clip of distinct maps is 40 GB of
host-side random numbers; the kernels' work does not depend on the values), detections drift per frame as in
`synth.clip_detections`."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from i2vsgg_b200 import shard, synth  # noqa: E402
from i2vsgg_b200.clip import ClipRunner  # noqa: E402
from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--det", type=int, default=64)
    ap.add_argument("--group", type=int, default=4)
    ap.add_argument("--graphs", action="store_true", help="replay each frame group as a CUDA graph and check it against the eager run")
    ap.add_argument("--associate", action="store_true", help="rank 0 also runs the temporal association on the gathered records")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args = synth.VrdArgs()
    head = vrd(args, None, synth.prd_vectors(7))
    head.load_state_dict({k: torch.from_numpy(v) for k, v in synth.vrd_params(1234, args).items()})
    head = head.to(dev).eval().prepare()
    lo, hi = shard.frame_range(a.frames, rank, world)
    boxes, classes, conf = synth.clip_detections(5, a.frames, a.det)
    boxes = torch.from_numpy(boxes[lo:hi]).to(dev)
    classes = torch.from_numpy(np.tile(classes, (hi - lo, 1))).to(dev)
    conf = torch.from_numpy(np.tile(conf, (hi - lo, 1))).to(dev)
    fmap1 = torch.from_numpy(synth.feature_map(100 + rank, 1)).to(dev)
    fmaps = fmap1.expand(hi - lo, -1, -1, -1)
    runner = ClipRunner(head, synth.IM_H, synth.IM_W, a.group, graphs=a.graphs)
    warm = min(hi - lo, a.group)
    eager = runner._group(fmaps[:warm].contiguous(), boxes[:warm], classes[:warm], conf[:warm])
    graph_equal = None
    if a.graphs:                     # capture here (outside the timed region) and check the replay against the eager group
        replay = runner._group_replayed(fmaps[:warm].contiguous(), boxes[:warm], classes[:warm], conf[:warm])
        graph_equal = bool(torch.equal(replay[0], eager[0]) and torch.equal(replay[1], eager[1]))

    class View:                      # hands out contiguous groups of the expanded map without materialising the clip
        shape = fmaps.shape
        device = dev

        def __getitem__(self, sl):
            return fmaps[sl].contiguous()

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rec, cnt = runner.run(View(), boxes, classes, conf, a.frames, rank, world)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = rec.shape == (a.frames, 100, 13) and int(cnt.min()) == 100
    assoc = None
    if rank == 0 and a.associate:
        from i2vsgg_b200 import sgg
        import time
        sgg.association(rec[:8], cnt[:8])                       # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rels = sgg.association(rec, cnt)
        assoc = {"ms": 1e3 * (time.perf_counter() - t0), "relations": len(rels),
                 "longest": max([r["duration"][1] - r["duration"][0] for r in rels] or [0])}
    if rank == 0:
        print(json.dumps({"workload": f"configs[4]: {a.frames}-frame clip, {a.det} detections -> {a.det * (a.det - 1)} pairs per "
                                      f"frame, pair build + vrd.forward + triplet top-100, all-gather of records",
                          "n_gpus": world, "frames": a.frames, "ms": float(ms.item()),
                          "frames_per_s": a.frames / (float(ms.item()) * 1e-3), "records_ok": bool(ok), "graphs": a.graphs, "graph_replay_equals_eager": graph_equal,
                          "association": assoc,
                          "gather_bytes": int(rec.numel() * 4)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
