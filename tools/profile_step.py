""" development tool: per-kernel / per-shape device time of one seed+rollout step (CUDA events around every launch) """
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import configs, ops
from playslot_b200.pipeline import seed_rollout

exp = sys.argv[1] if len(sys.argv) > 1 else "GridShapes5Objs"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = configs.experiment(exp)
torch.manual_seed(0)
savi, pred = configs.build_models(cfg, "cuda")
video = torch.rand(B, 21, 3, 64, 64, device="cuda")
for _ in range(2):
    seed_rollout(savi, pred, video, 6, 15)
torch.cuda.synchronize()

# stage split with events
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
with torch.no_grad():
    ev[0].record()
    out = savi(video, num_imgs=21, decode=False)
    ev[1].record()
    ps, _ = pred.forward_multiple_samples(out["slot_history"], 1, 6, 15, use_posterior=True)
    ev[2].record()
    S, D = ps.shape[-2:]
    savi.decode(ps.reshape(B * 15, S, D))
    ev[3].record()
torch.cuda.synchronize()
print(f"{exp} B={B}: encode {ev[0].elapsed_time(ev[1]):.1f} ms | invdyn+rollout {ev[1].elapsed_time(ev[2]):.1f} ms | decode {ev[2].elapsed_time(ev[3]):.1f} ms")

ops.profile_start()
seed_rollout(savi, pred, video, 6, 15)
prof = ops.profile_stop()
tot = sum(v["ms"] for v in prof.values())
print(f"sum of kernel times {tot:.1f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:22s} {v['ms']:8.2f} ms  {v['calls']:5d} calls")
# linear by shape: re-run with a per-shape breakdown (events are per call, aggregate by meta)
ops.profile_start()
seed_rollout(savi, pred, video, 6, 15)
rec = ops._PROFILE
ops._PROFILE = None
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, e0, e1, meta in rec:
    if name.startswith("linear"):
        agg[meta][0] += 1
        agg[meta][1] += e0.elapsed_time(e1)
print("linear by (M, N, K):")
for m, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    fl = 2.0 * m[0] * m[1] * m[2] * n
    print(f"  {str(m):24s} {n:4d} calls {t:8.2f} ms  {fl / t / 1e9:8.1f} TFLOP/s" if t > 0 else m)
