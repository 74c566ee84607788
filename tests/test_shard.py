"""CPU tests of the multi-rank path: world_size-2 gloo all-gather of per-frame triplet records, frame partitioning."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from i2vsgg_b200 import shard


def test_frame_range_partitions_in_order():
    for frames, world in [(1024, 8), (10, 4), (3, 8), (0, 2), (7, 1)]:
        got = []
        for r in range(world):
            lo, hi = shard.frame_range(frames, r, world)
            assert 0 <= lo <= hi <= frames and hi - lo <= shard.frames_per_rank(max(frames, 1), world)
            got += list(range(lo, hi))
        assert got == list(range(frames))


def _worker(rank, world, port, frames, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(123)
    full = torch.randn((frames, shard.TOP_K, shard.RECORD_WIDTH), generator=g)
    cnt = torch.randint(0, 101, (frames,), generator=g, dtype=torch.int32)
    lo, hi = shard.frame_range(frames, rank, world)
    rec, c = shard.all_gather_triplets(full[lo:hi].clone(), cnt[lo:hi].clone(), frames)
    ok = torch.equal(rec, full) and torch.equal(c, cnt)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("frames", [8, 7, 1])
def test_all_gather_triplets_gloo_world2(tmp_path, frames):
    port = 29500 + (os.getpid() % 500) + frames
    mp.spawn(_worker, args=(2, port, frames, str(tmp_path)), nprocs=2, join=True)
    assert all(open(tmp_path / f"ok{r}").read() == "1" for r in range(2))
