"""Frame sharding over ranks and the one exchange step of the path: the all-gather of per-frame triplet records into
video order.  Frames are independent until the temporal association
(lib/utils.py:134-182 needs them in order, `fstart == r.fend` at :166), so each rank takes a contiguous chunk and the
only collective is a fixed-size all-gather issued once per clip.

torch.distributed is plumbing here (NCCL over NVLink on the GPUs, gloo in the CPU tests); the payload is 5.2 KB per
frame, so the collective is latency-bound and a single call per clip is the right granularity.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_WIDTH = 13   # conf, cls_s, rel, cls_o, sub box x4, obj box x4, pair idx  (test_net_SGG_emb.py:209)
TOP_K = 100


def frames_per_rank(num_frames: int, world: int) -> int:
    return (num_frames + world - 1) // world


def frame_range(num_frames: int, rank: int, world: int):
    """Contiguous chunk [lo, hi) of ceil(F/G) frames for `rank` (the last ranks may get fewer or none)."""
    per = frames_per_rank(num_frames, world)
    lo = min(num_frames, rank * per)
    return lo, min(num_frames, lo + per)


def all_gather_triplets(records: torch.Tensor, counts: torch.Tensor, num_frames: int, group=None):
    """records [frames_local, TOP_K, 13] fp32, counts [frames_local] int32 (this rank's frames, in order)
    -> (records [num_frames, TOP_K, 13], counts [num_frames]) on every rank, in video order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return records[:num_frames], counts[:num_frames]
    per = frames_per_rank(num_frames, world)
    k, w = records.shape[1], records.shape[2]
    # one buffer per rank: records followed by the counts (as float), padded to `per` frames -> a single collective
    send = torch.zeros((per, k * w + 1), dtype=torch.float32, device=records.device)
    n = records.shape[0]
    send[:n, : k * w] = records.reshape(n, k * w)
    send[:n, k * w] = counts.to(torch.float32)
    recv = torch.empty((world * per, k * w + 1), dtype=torch.float32, device=records.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv[:num_frames]
    return recv[:, : k * w].reshape(num_frames, k, w), recv[:, k * w].to(torch.int32)
