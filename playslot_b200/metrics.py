"""
On-device evaluation metrics of the predicted frames -- what `05_evaluate_PlaySlot.py:98-108` computes on the CPU after copying
every predicted frame to the host (`preds_eval[:, 0].cpu()` -> lib/metrics.py MetricTracker).  Here the per-frame squared error is
reduced on the GPU in one pass over predictions and targets (csrc/metrics.cu), so only (B, num_preds) numbers travel.

PSNR follows piqa.psnr.psnr (the function lib/metrics.py:232 calls; piqa is a third-party dependency absent from the reference
tree -- its published definition is restated): per image, mse = mean over (C, H, W) of (x - y)^2 with both clamped to [0, 1] by
the caller, psnr = 10 * log10(value_range^2 / (mse + 1e-8)), value_range = 1.
"""
import torch

from . import ops


def frame_mse(pred, target, clamp01=True):
    """ pred, target (B, F, C, H, W) on the GPU (target may be a time slice of the clip) -> (B, F) mean squared error per frame """
    elems = 1
    for d in pred.shape[2:]:
        elems *= d
    return ops.frame_sq_err(pred if pred.stride(-1) == 1 else pred.contiguous(), target, clamp01=clamp01) / elems


def psnr(pred, target, epsilon=1e-8, value_range=1.0):
    """ (B, F) PSNR per frame; `.mean()` = lib/metrics.py PSNR.aggregate()[0], `.mean(0)` = its frame-wise curve """
    return 10.0 * torch.log10(value_range ** 2 / (frame_mse(pred, target, clamp01=True) + epsilon))


@torch.no_grad()
def evaluate_rollout(savi, predictor, videos, num_seed, num_preds):
    """ `forward_eval` of src/05_evaluate_PlaySlot.py:66-108 for the posterior sample with the metric tail on the device:
    returns dict(psnr (B, np), mse (B, np)) -- 2 * B * np floats instead of B * np frames leave the GPU """
    from .pipeline import seed_rollout
    res = seed_rollout(savi, predictor, videos, num_seed, num_preds)
    pred = res["pred_imgs"][:, 0]
    tgt = videos[:, num_seed:num_seed + num_preds]
    if tgt.dtype == torch.uint8:
        tgt = tgt.float().div(255)
    mse = frame_mse(pred, tgt, clamp01=True)
    return {"mse": mse, "psnr": 10.0 * torch.log10(1.0 / (mse + 1e-8)), "pred_slots": res["pred_slots"]}
