""" development tool: per-kernel device time of the SlotAttention forward at BASELINE config 2
    (B=64, N=4096, D=128, 3 iterations) for both arithmetic modes and several tokens-per-CTA settings """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops
from playslot_b200.models.BlocksUtils.attention import SlotAttention

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S in (8, 4):
    torch.manual_seed(S)
    m = SlotAttention(128, 128, S, 3, 1, 256).eval().cuda()
    feats = torch.randn(64, 4096, 128, device="cuda")
    slots = torch.randn(64, S, 128, device="cuda")
    for mode in ("tc", "tc1", "fp32"):
        for tokens in ((128, 256, 512, 1024) if mode != "fp32" else (256,)):
            m.mode = mode
            SlotAttention.TOKENS_PER_CTA = tokens
            with torch.no_grad():
                for _ in range(3):
                    m(feats, slots, step=0)
                ts = []
                for _ in range(10):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); m(feats, slots, step=0); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                ops.profile_start()
                for _ in range(5):
                    m(feats, slots, step=0)
                prof = ops.profile_stop()
            det = "  ".join(f"{k} {1e3 * v['ms'] / v['calls']:.1f}us x{v['calls'] // 5}" for k, v in prof.items())
            by = 64 * 4096 * 128 * 4
            main = [v for k, v in prof.items() if k.startswith("sa_main")][0]
            gbs = by / (main["ms"] / main["calls"] * 1e-3) / 1e9
            print(f"S={S} mode={mode} tokens/CTA={tokens}: forward {ts[len(ts)//2]*1e3:.0f} us (L2 flushed) | {det} | main pass {gbs:.0f} GB/s (warm)")
