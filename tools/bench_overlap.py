"""
Sweep of the decode / roll-out overlap (pipeline.seed_rollout(overlap=...)): SMs reserved for the roll-out chain while a predicted
frame is decoded on the side stream.  GridShapes5Objs, 64 videos, 6 + 15 frames, CUDA graph, 10 timed replays each.

    python tools/bench_overlap.py [reserve values ...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from playslot_b200 import configs, pipeline          # noqa: E402


def main():
    reserves = [int(a) for a in sys.argv[1:]] or [-1, 0, 8, 12, 16, 24, 32]
    cfg = configs.experiment("GridShapes5Objs")
    torch.manual_seed(0)
    savi, pred = configs.build_models(cfg, "cuda")
    B, ns, npred = 64, 6, 15
    video = torch.rand(B, ns + npred, 3, 64, 64, device="cuda")
    res = {}
    ref = None
    for r in reserves:
        pipeline.OVERLAP_DECODE = r >= 0
        pipeline.SM_RESERVE = max(r, 0)
        g = pipeline.GraphedSeedRollout(savi, pred, B, ns, npred)
        noise = g.draw_noise()
        for _ in range(3):
            out = g(video, noise=noise)
        torch.cuda.synchronize()
        imgs = out["pred_imgs"].clone()
        if ref is None:
            ref = imgs
        same = bool(torch.equal(imgs, ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g(video, noise=noise)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["sequential" if r < 0 else f"reserve_{r}"] = {"ms_per_step": ms, "frames_per_s": B * (ns + npred) / ms * 1e3,
                                                           "launches": g.launches_per_replay, "bit_identical_to_first": same}
        print(r, res["sequential" if r < 0 else f"reserve_{r}"], flush=True)
        del g
    print(json.dumps(res))


if __name__ == "__main__":
    main()
