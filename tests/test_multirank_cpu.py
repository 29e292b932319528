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


# ------------------------------------------------------------------------------------------------ training collectives

def _train_worker(rank, world_size, port, q):
    """ gradient bucket all-reduce and EMA-VQ statistic sync of playslot_b200.train / quantization on CPU tensors """
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from playslot_b200 import train
        from playslot_b200.models.BlocksUtils.quantization import EmaVectorQuantizer
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        frozen = torch.nn.Linear(3, 3).requires_grad_(False)
        bucket = train.GradientBucket(list(lin.parameters()) + list(frozen.parameters()))
        g = torch.Generator().manual_seed(100)
        x_all = torch.randn(6 * world_size + 1, 5, generator=g)      # shards of UNEQUAL size
        a, b = shard.shard_range(x_all.shape[0], rank, world_size)
        bucket.attach()                                   # .grad tensors are views of the flat bucket: no copy in / out
        lin(x_all[a:b]).pow(2).mean().backward()          # per-rank mean loss on the shard
        assert lin[0].weight.grad.data_ptr() == bucket.views[0].data_ptr()
        lin[1].bias.grad = None                           # a parameter without gradient on this rank must still take part
        nbytes = bucket.all_reduce_mean(local_count=b - a)
        grads = [None if p.grad is None else p.grad.clone() for p in lin.parameters()]
        # EMA codebook statistics on the global batch
        torch.manual_seed(1)
        vq = EmaVectorQuantizer(num_embs=4, emb_dim=3)
        z_all = torch.randn(4 * world_size, 5, 3, generator=g)
        idx_all = torch.randint(0, 4, (4 * world_size, 5, 1), generator=g)
        a2, b2 = shard.shard_range(z_all.shape[0], rank, world_size)
        vq.ema_update(z_all[a2:b2], idx_all[a2:b2])
        q.put((rank, nbytes, [None if t is None else t.tolist() for t in grads], vq.ema_count.tolist(), vq.ema_weight.tolist(),
               vq.vq.embeddings["0"].weight.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2])
def test_training_collectives_equal_single_process(world_size):
    from playslot_b200.models.BlocksUtils.quantization import EmaVectorQuantizer
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world_size, port, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the gradient of the mean loss over the GLOBAL (concatenated) batch
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(6 * world_size + 1, 5, generator=g)
    lin(x_all).pow(2).mean().backward()
    acc = [p.grad.clone() for p in lin.parameters()]
    acc[3].zero_()                                         # lin[1].bias was dropped on every rank
    torch.manual_seed(1)
    vq = EmaVectorQuantizer(num_embs=4, emb_dim=3)
    vq.sync_stats = False
    z_all = torch.randn(4 * world_size, 5, 3, generator=g)
    idx_all = torch.randint(0, 4, (4 * world_size, 5, 1), generator=g)
    vq.ema_update(z_all, idx_all)
    n_float = sum(p.numel() for p in lin.parameters())
    for rank, nbytes, grads, cnt, wgt, emb in res:
        assert nbytes == 4 * (n_float + 1)                 # one flat bucket: every trainable parameter + the sample count
        for got, ref in zip(grads, acc):
            assert got is not None
            assert torch.allclose(torch.tensor(got), ref, atol=1e-6)
        assert torch.allclose(torch.tensor(cnt), vq.ema_count, atol=1e-6)          # == single GPU on the global batch
        assert torch.allclose(torch.tensor(wgt), vq.ema_weight, atol=1e-6)
        assert torch.allclose(torch.tensor(emb), vq.vq.embeddings["0"].weight, atol=1e-6)
