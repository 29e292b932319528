"""
The seed + rollout path a user of the reference runs (src/05_evaluate_PlaySlot.py:112-139, `forward_model`):
    savi(videos, num_imgs=ns+np, decode=False) -> predictor.forward_multiple_samples(slot_history, num_samples,
    ns, np, use_posterior=True) -> savi.decode(pred_slots)
exposed as one call, plus a host-buffer variant used for end-to-end timing.
"""
import torch


@torch.no_grad()
def seed_rollout(savi, predictor, videos, num_seed, num_preds, num_samples=1, use_posterior=True, decode=True, **init_kwargs):
    """
    videos (B, >= num_seed+num_preds, 3, H, W) on the GPU ->
      dict(slot_history (B, ns+np, S, D), pred_slots (B, num_samples, np, S, D), pred_imgs (B, num_samples, np, 3, H, W),
           pred_objs, pred_masks)
    """
    T = num_seed + num_preds
    out = savi(videos, num_imgs=T, decode=False, **init_kwargs)
    slot_history = out["slot_history"]
    pred_slots, actions = predictor.forward_multiple_samples(
        slot_history=slot_history, num_samples=num_samples, num_seed=num_seed, num_preds=num_preds, use_posterior=use_posterior)
    res = {"slot_history": slot_history, "pred_slots": pred_slots, "actions": actions}
    if decode:
        B, n, P, S, D = pred_slots.shape
        imgs, (objs, masks) = savi.decode(pred_slots.reshape(B * n * P, S, D))
        res["pred_imgs"] = imgs.view(B, n, P, *imgs.shape[1:])
        res["pred_objs"] = objs.view(B, n, P, *objs.shape[1:])
        res["pred_masks"] = masks.view(B, n, P, *masks.shape[1:])
    return res


@torch.no_grad()
def seed_rollout_host(savi, predictor, videos_host, num_seed, num_preds, out_host=None, device="cuda", **kw):
    """
    Host-buffer entry point: `videos_host` is a (pinned) CPU tensor; the predicted frames come back in
    `out_host` (pinned CPU, (B, np, 3, H, W)).  Both copies are stream-ordered with the kernels.
    """
    v = videos_host.to(device, non_blocking=True)
    res = seed_rollout(savi, predictor, v, num_seed, num_preds, **kw)
    imgs = res["pred_imgs"][:, 0]
    if out_host is None:
        out_host = torch.empty(imgs.shape, dtype=imgs.dtype, pin_memory=True)
    out_host.copy_(imgs, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host, res
