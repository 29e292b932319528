""" development tool: forward_multiple_samples(num_samples=10, prior actions), batched vs one roll-out per sample """
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import configs

cfg = configs.experiment("GridShapes5Objs")
torch.manual_seed(0)
savi, pred = configs.build_models(cfg, "cuda")
B = 64
sh = torch.randn(B, 6, savi.num_slots, 128, device="cuda")
for batched in (False, True):
    pred.batch_prior_samples = batched
    with torch.no_grad():
        for _ in range(2):
            pred.forward_multiple_samples(sh, 10, 6, 15, use_posterior=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            pred.forward_multiple_samples(sh, 10, 6, 15, use_posterior=False)
        torch.cuda.synchronize()
        print(f"batched={batched}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms for 10 samples x {B} videos x 15 frames (eager)")
