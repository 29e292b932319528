"""
Joint InvDyn + cOCVP training step -- host-side mirror of `Trainer.forward_loss_metric`
(src/04_train_PlaySlot.py:32-122) and `BasePredictorTrainer.optimization` (src/base/basePredictorTrainer.py:320-333),
for batch-sharded data parallelism (one process per GPU):

    with no_grad:  slot_history = savi(videos, num_imgs=ns+np, decode=False)             SAVi frozen
    pred_slots, out = predictor(slots=slot_history, num_seed=ns, num_preds=np)           InvDyn + cOCVP, BPTT
    pred_imgs = savi.decode(pred_slots)                                                  data gradient only
    loss = MSE(pred_imgs, videos[:, 1:ns+np]) + MSE(pred_slots, slot_history[:, 1:ns+np])
           + mean(quant) + 0.25 * mean(commit)                                           (loss.py:66-92,201-238,319-340)
    loss.backward();  all-reduce(grads) / world;  clip_grad_norm_(0.05);  Adam.step()

The scalar losses, the clip and Adam are plain PyTorch like in the reference; every
forward / backward arithmetic node is a playslot_b200 kernel.  The only collective is the gradient all-reduce
(NCCL over NVLink), issued as ONE flat bucket, plus the (K + K*A + 1)-float EMA-VQ statistics inside the quantizer.
"""
import torch
import torch.distributed as dist


def predictor_losses(pred_slots, target_slots, pred_imgs, target_imgs, vq_losses, commit_weight=0.25):
    """ the three loss terms of the shipped PlaySlot configs; returns (total, dict of detached terms) """
    img = torch.nn.functional.mse_loss(pred_imgs, target_imgs)
    slot = torch.nn.functional.mse_loss(pred_slots, target_slots)
    quant, commit = vq_losses["quant_loss"].mean(), vq_losses["commit_loss"].mean()
    total = img + slot + quant + commit_weight * commit
    return total, {"pred_img_mse": img.detach(), "pred_slot_mse": slot.detach(), "quant": quant.detach(), "commit": commit.detach()}


def forward_loss(savi, predictor, videos, num_seed, num_preds, **init_kwargs):
    """ forward part of the step; returns (loss, terms, pred_slots) """
    T = num_seed + num_preds
    with torch.no_grad():
        slot_history = savi(videos, num_imgs=T, decode=False, **init_kwargs)["slot_history"]
    pred_slots, out = predictor(slots=slot_history, num_seed=num_seed, num_preds=num_preds)
    B, n, S, D = pred_slots.shape
    pred_imgs, _ = savi.decode(pred_slots.reshape(B * n, S, D))
    pred_imgs = pred_imgs.view(B, n, *pred_imgs.shape[1:])
    loss, terms = predictor_losses(pred_slots, slot_history[:, 1:T], pred_imgs, videos[:, 1:T], out["vq_losses"])
    return loss, terms, pred_slots


class GradientBucket:
    """ one flat fp32 buffer for all trainable gradients: a single NCCL all-reduce per step """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def all_reduce_mean(self):
        """ grads -> bucket -> all-reduce(SUM) / world -> grads; no-op bookkeeping when not distributed """
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat)
        self.flat.div_(dist.get_world_size())
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat.numel() * 4


def train_step(savi, predictor, optimizer, videos, num_seed, num_preds, bucket=None, clip_gradient=0.05, **init_kwargs):
    """ one optimisation step on this rank's shard of the batch; returns (loss value tensor, terms) """
    predictor.train()
    optimizer.zero_grad(set_to_none=True)
    loss, terms, _ = forward_loss(savi, predictor, videos, num_seed, num_preds, **init_kwargs)
    loss.backward()
    if bucket is not None:
        bucket.all_reduce_mean()
    if clip_gradient is not None:
        torch.nn.utils.clip_grad_norm_(predictor.parameters(), clip_gradient)
    optimizer.step()
    return loss.detach(), terms


def make_optimizer(predictor, lr=1e-4):
    """ Adam as in lib/setup_model.py:372 """
    return torch.optim.Adam([p for p in predictor.parameters() if p.requires_grad], lr=lr)
