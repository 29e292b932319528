"""
world_size-2 (and 3) gloo tests of the multi-rank host logic (batch sharding, MAX timing reduction, rank-ordered
gather): run on CPU, no GPU needed.  The data path itself has no collective.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from playslot_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world_size, port, n_videos, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert shard.world() == (rank, world_size, rank)
        videos = torch.arange(n_videos * 6, dtype=torch.float32).view(n_videos, 2, 3)
        mine = shard.shard_batch(videos, rank, world_size)
        # every rank runs the same "pipeline" on its shard (stand-in: a per-video function)
        local = mine.sum(dim=(1, 2), keepdim=False).unsqueeze(1) * 2.0
        full = shard.gather_concat(local)
        t = shard.max_over_ranks(10.0 + rank)
        q.put((rank, mine.shape[0], full.flatten().tolist(), t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size,n_videos", [(2, 8), (2, 5), (3, 7)])
def test_sharded_run_equals_single_process(world_size, n_videos):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, n_videos, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    videos = torch.arange(n_videos * 6, dtype=torch.float32).view(n_videos, 2, 3)
    expect = (videos.sum(dim=(1, 2)) * 2.0).tolist()
    assert sum(r[1] for r in res) == n_videos
    for rank, n_local, full, t in res:
        assert full == expect                      # rank-ordered concatenation == single-process result
        assert t == 10.0 + world_size - 1          # MAX over ranks
        a, b = shard.shard_range(n_videos, rank, world_size)
        assert n_local == b - a


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
