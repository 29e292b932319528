"""
SURVEY.md section 4-v: N ranks == 1 GPU on the concatenated batch.  Launch under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multirank_parity.py

Every rank runs the seed + rollout pipeline (GridShapes5Objs, 6 + 15 frames) on its contiguous shard of a seeded global batch of
16 videos; the shards' predicted frames / slots are gathered in rank order and rank 0 compares them with its own run over the whole
batch.  Videos are independent, so the only difference allowed is summation order inside the SlotAttention partials (the tile
partition depends on the batch size): rel-L2 <= 1e-5.  Prints one JSON line on rank 0.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from playslot_b200 import configs, shard                  # noqa: E402
from playslot_b200.pipeline import seed_rollout           # noqa: E402


def main():
    rank, world, local = shard.world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = configs.experiment("GridShapes5Objs")
    torch.manual_seed(0)
    savi, pred = configs.build_models(cfg, "cuda")
    Bg, ns, npred = 16, 6, 15
    g = torch.Generator().manual_seed(4321)
    video = torch.rand(Bg, ns + npred, 3, 64, 64, generator=g)
    noise = torch.randn(Bg, ns + npred - 1, 6, 8, generator=g)
    init = torch.randn(Bg, 6, 128, generator=g)

    def run(lo, hi):
        vals = [init[lo:hi], noise[lo:hi]]                    # LearnedRandom draw (device), then the InvDyn draw (host)
        real = torch.randn

        def fake(*a, **k):
            t = vals.pop(0)
            return t.clone().to(k["device"]) if k.get("device") is not None else t.clone()
        torch.randn = fake
        try:
            return seed_rollout(savi, pred, video[lo:hi].cuda(), ns, npred)
        finally:
            torch.randn = real

    a, b = shard.shard_range(Bg, rank, world)
    mine = run(a, b)
    gathered = {k: shard.gather_concat(mine[k].contiguous()) for k in ("pred_imgs", "pred_slots", "pred_masks", "slot_history")}
    if rank == 0:
        full = run(0, Bg)
        out = {"world": world, "global_batch": Bg, "shards": [list(shard.shard_range(Bg, r, world)) for r in range(world)]}
        ok = True
        for k, v in gathered.items():
            ref = full[k]
            rel = float((v.double() - ref.double()).norm() / ref.double().norm())
            mx = float((v - ref).abs().max())
            out[k] = {"rel_l2": rel, "max_abs": mx, "bit_identical": bool(torch.equal(v, ref))}
            ok = ok and rel <= 1e-5
        out["ok"] = ok
        print(json.dumps(out))
        assert ok, out
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
