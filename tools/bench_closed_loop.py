"""
Single-step latency of the latent-imagination loop (src/11_evaluate_behavior_on_simulation.py:149-183): decompose_action_latent ->
enforce_window -> cOCVP.forward_single -> decode of the predicted frame, batch 1 (planning) and 16, eager launches vs one CUDA graph
per window length (pipeline.ClosedLoopStepper).   python tools/bench_closed_loop.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from playslot_b200 import configs                       # noqa: E402
from playslot_b200.pipeline import ClosedLoopStepper    # noqa: E402


def main():
    out = {}
    for exp in ("BlockPush", "GridShapes5Objs"):
        cfg = configs.experiment(exp)
        torch.manual_seed(0)
        savi, pred = configs.build_models(cfg, "cuda")
        A = cfg["predictor"]["predictor_params"]["action_dim"]
        S = cfg["model"]["model_params"]["num_slots"]
        per_slot = cfg["predictor"]["predictor_name"] == "SlotLatentPredictor"
        for B in (1, 16):
            for graphs in (False, True):
                st = ClosedLoopStepper(savi, pred, B, graphs=graphs)
                img = torch.rand(B, 3, 64, 64, device="cuda")
                lat = torch.randn((B, S, A) if per_slot else (B, A), device="cuda")
                for rep in range(2):                      # the first pass captures the graphs
                    st.reset(img)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(20):
                        o = st.step(lat)
                    torch.cuda.synchronize()
                    ms = (time.perf_counter() - t0) / 20 * 1e3
                out[f"{exp}_B{B}_{'graph' if graphs else 'eager'}"] = {"ms_per_step_wall": ms, "steps_per_s": 1e3 / ms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
