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
    from . import autograd as AG
    # one fused reduction pass per term (csrc/metrics.cu); the targets are time slices of the clip / slot history, read in place
    img = AG.mse_loss(pred_imgs, target_imgs)
    slot = AG.mse_loss(pred_slots, target_slots)
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
    """
    One flat fp32 buffer holding ALL trainable gradients: the parameters' `.grad` tensors are views into it (autograd
    accumulates straight into the bucket), so a step issues a single NCCL all-reduce with no copy in or out.  The last
    element carries the rank's sample count: ranks whose shards differ in size (B % world != 0, shard.shard_range) weight
    their mean-loss gradients by it, which makes the result the gradient of the GLOBAL-batch mean loss
    (basePredictorTrainer.py:320-333 on the concatenated batch).
    """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + 1, device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.grads = self.flat[:n]
        self.count = self.flat[n:]

    def attach(self):
        """ make every parameter's .grad the bucket view (zeroed); call instead of optimizer.zero_grad(set_to_none=True) """
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
            p._ps_bucket_grad = v        # autograd._Linear accumulates its weight / bias gradients straight into the view

    def _gather_detached(self):
        """ gradients that autograd rebound (set_to_none users): copy them in; the fast path never takes this branch """
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def all_reduce_mean(self, local_count=None, async_op=False):
        """ in-place all-reduce of the bucket to the global-batch mean gradient; returns the bytes reduced (0 when not
        distributed) or, with async_op, the work handle (call .wait(), then `finish()`) """
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0
        self._gather_detached()
        if local_count is None:
            self.count.fill_(1.0)
        else:
            self.count.fill_(float(local_count))
            self.grads.mul_(float(local_count))
        if async_op:
            return dist.all_reduce(self.flat, async_op=True)
        dist.all_reduce(self.flat)
        self.finish()
        return self.flat.numel() * 4

    def finish(self):
        self.grads.div_(self.count)                 # device-side division by the global count: no host synchronisation


def train_step(savi, predictor, optimizer, videos, num_seed, num_preds, bucket=None, clip_gradient=0.05, **init_kwargs):
    """ one optimisation step on this rank's shard of the batch; returns (loss value tensor, terms) """
    predictor.train()
    if bucket is not None:
        bucket.attach()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss, terms, _ = forward_loss(savi, predictor, videos, num_seed, num_preds, **init_kwargs)
    from . import autograd as AG
    with AG.accumulate_into_bucket():           # (a no-op without an attached bucket)
        loss.backward()
    if bucket is not None:
        bucket.all_reduce_mean(local_count=videos.shape[0])
    if clip_gradient is not None:
        torch.nn.utils.clip_grad_norm_(predictor.parameters(), clip_gradient)
    optimizer.step()
    return loss.detach(), terms


def make_optimizer(predictor, lr=1e-4, capturable=False):
    """ Adam as in lib/setup_model.py:372; capturable=True keeps the step counter on the device (needed by GraphedTrainStep) """
    return torch.optim.Adam([p for p in predictor.parameters() if p.requires_grad], lr=lr, capturable=capturable)


def _trackers(module):
    """ the python-side codeword-usage meters of the quantizers (not in the state_dict) """
    for m in module.modules():
        tr = getattr(m, "trackers", None)
        if isinstance(tr, dict):
            yield from tr.values()


def _snapshot_training_state(predictor, optimizer):
    with torch.no_grad():
        return {
            "params": [p.detach().clone() for p in predictor.parameters()],
            "buffers": [b.detach().clone() for b in predictor.buffers()],
            "opt": {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in optimizer.state.items()},
            "trackers": [(tr.bins.data.clone(), tr._iters.clone()) for tr in _trackers(predictor)],
        }


def _restore_training_state(predictor, optimizer, snap):
    """ in place: the captured graph keeps reading and writing these very tensors """
    with torch.no_grad():
        for p, v in zip(predictor.parameters(), snap["params"]):
            p.copy_(v)
        for b, v in zip(predictor.buffers(), snap["buffers"]):
            b.copy_(v)
        for p, st in optimizer.state.items():
            old = snap["opt"].get(p)
            for k, v in st.items():
                if torch.is_tensor(v):
                    if old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()                  # a fresh Adam state: step 0, zero moments
        for tr, (bins, iters) in zip(_trackers(predictor), snap["trackers"]):
            tr.bins.data.copy_(bins)
            tr._iters.copy_(iters)
        optimizer.zero_grad(set_to_none=True)


class GraphedTrainStep:
    """
    `train_step` for a fixed batch shape captured in ONE CUDA graph (forward, backward, clip, Adam, EMA-VQ update): at the
    reference's batch size (8 videos per GPU) the ~3000 launches of a step are host-bound when issued from Python.
    With torch.distributed initialised the NCCL gradient all-reduce and the EMA-VQ statistic all-reduce are part of the graph
    (every rank must build and call this object in lock step).  The InvDyn noise is drawn
    on the host with the reference's RNG call and uploaded into a static buffer before every replay, like
    pipeline.GraphedSeedRollout.  The optimizer must be built with capturable=True.
    linear_mode="fp16x3" runs the forward and data-gradient GEMMs of the predictor's Linear layers on the tensor cores
    (weights repacked inside the graph every step); the default keeps them exact fp32.
    """

    def __init__(self, savi, predictor, optimizer, batch, num_seed, num_preds, resolution=(64, 64), clip_gradient=0.05, warmup=3,
                 bucket=None, linear_mode="fp32"):
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # Multi-GPU: the gradient all-reduce (one flat bucket the .grad tensors alias) and the EMA-VQ statistic all-reduce are
        # captured INSIDE the graph.  NCCL capture needs (a) a communicator that has already run on this device -- the
        # warm-up steps below do that -- and (b) capture_error_mode="thread_local", because the NCCL watchdog thread polls CUDA
        # events while the capture is open.
        # (single GPU too: the bucket views are what the Linear backward accumulates into without autograd add kernels)
        self.bucket = bucket if bucket is not None else GradientBucket(predictor.parameters())
        self.savi, self.predictor, self.optimizer = savi, predictor, optimizer
        self.num_seed, self.num_preds = num_seed, num_preds
        T = num_seed + num_preds
        H, W = resolution
        la = predictor.latent_action
        per_slot = type(la).__name__ == "VQSimpleMLPSlotLatentAction"
        S, A = savi.num_slots, la.action_dim
        self.noise_shape = (batch, T - 1, S, A) if per_slot else (batch, T - 1, A)
        dev = next(predictor.parameters()).device
        self.video = torch.zeros(batch, T, 3, H, W, device=dev)
        self.noise = torch.zeros(self.noise_shape, device=dev)
        from .pipeline import PinnedRing
        self._noise_ring = PinnedRing(self.noise_shape)
        self._la = la
        predictor.train()

        from . import autograd as AG
        if linear_mode not in ("fp32", "fp16x3", "bf16"):
            raise ValueError(f"linear_mode={linear_mode!r}; expected 'fp32', 'fp16x3' or 'bf16'")
        self.linear_mode = linear_mode

        def body():
            if self.bucket is not None:
                self.bucket.attach()                                               # zeroed flat bucket, .grad = views
            else:
                optimizer.zero_grad(set_to_none=True)
            old_mode, AG.TRAIN_LINEAR_MODE = AG.TRAIN_LINEAR_MODE, linear_mode     # arithmetic of the differentiable Linears
            try:
                loss, terms, _ = forward_loss(savi, predictor, self.video, num_seed, num_preds)
                with AG.accumulate_into_bucket():
                    loss.backward()
            finally:
                AG.TRAIN_LINEAR_MODE = old_mode
            if self.bucket is not None:
                self.bucket.all_reduce_mean(local_count=batch)
            if clip_gradient is not None:
                torch.nn.utils.clip_grad_norm_(predictor.parameters(), clip_gradient)
            optimizer.step()
            return loss.detach(), terms

        # The warm-up steps (they build lazy caches, workspaces and the optimizer state tensors the graph will update) are real
        # optimisation steps on a dummy batch: snapshot every piece of training state first and put it back, IN PLACE, after
        # the capture, so that constructing this object does not change the model.
        snap = _snapshot_training_state(predictor, optimizer)
        la._static_noise = self.noise
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            self.graph = torch.cuda.CUDAGraph()
            if self.bucket is None:
                optimizer.zero_grad(set_to_none=True)
            kw = {"capture_error_mode": "thread_local"} if self.world > 1 else {}
            from . import ops as _ops
            n0 = _ops.launch_count()
            with torch.cuda.graph(self.graph, **kw):
                self.loss, self.terms = body()
            self.launches_per_replay = _ops.launch_count() - n0       # playslot_b200 kernels inside the graph
        finally:
            la._static_noise = None
            _restore_training_state(predictor, optimizer, snap)
            if self.bucket is not None:
                self.bucket.flat.zero_()

    def draw_noise(self):
        return torch.randn(self.noise_shape, dtype=torch.float32)

    def __call__(self, videos, noise=None):
        """ one optimisation step on `videos` (B, T, 3, H, W); returns (loss tensor, terms) -- static tensors of the graph """
        self.video.copy_(videos, non_blocking=True)
        self._noise_ring.upload(self.draw_noise() if noise is None else noise, self.noise)
        self.graph.replay()
        # the replay advanced the parameters on the device without bumping their Python-side `_version`: invalidate every
        # derived-weight cache (packed tensor-core copies, concatenated q/k/v / embedding / head matrices) so that an eager
        # forward after this step sees the new weights
        from . import ops
        ops.bump_weight_generation()
        return self.loss, self.terms


# ------------------------------------------------------------------------------------------------ SAVi training (02_train_savi.py)

def savi_forward_loss(savi, videos, **init_kwargs):
    """ `Trainer.forward_loss_metric` of src/02_train_savi.py:21-70: model(videos) with decode, MSE(recons_imgs, videos.clamp(0, 1)) """
    from . import autograd as AG
    out = savi(videos, num_imgs=videos.shape[1], **init_kwargs)
    target = videos if videos.dtype == torch.float32 else videos.float().div(255)
    loss = AG.mse_loss(out["recons_imgs"], target.clamp(0, 1))
    return loss, out


def savi_train_step(savi, optimizer, videos, bucket=None, clip_gradient=0.05, **init_kwargs):
    """ one optimisation step of the decomposition model on this rank's shard (gradient all-reduce through `bucket` when
    torch.distributed is initialised), as base/baseTrainer.py drives 02_train_savi.py; returns (loss value tensor, model output) """
    savi.train()
    if bucket is not None:
        bucket.attach()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss, out = savi_forward_loss(savi, videos, **init_kwargs)
    from . import autograd as AG
    with AG.accumulate_into_bucket():
        loss.backward()
    if bucket is not None:
        bucket.all_reduce_mean(local_count=videos.shape[0])
    if clip_gradient is not None:
        torch.nn.utils.clip_grad_norm_(savi.parameters(), clip_gradient)
    optimizer.step()
    return loss.detach(), out


class GraphedSaviTrainStep:
    """
    `savi_train_step` (src/02_train_savi.py:21-70 as driven by base/baseTrainer.py) for a fixed clip shape captured in ONE CUDA
    graph: forward with decode, MSE loss, full backward (encoder, SlotAttention, transition, decoder), clip, Adam.  Launched from
    Python the ~1500 kernels of a step leave a third of the time to the host.  The convolution weights are repacked inside the
    graph every replay with the power-of-two scale chosen at their first packing (`ops.REUSE_CONV_PACK_SCALE`: no host
    synchronisation), the gradients accumulate in a flat GradientBucket (all-reduced inside the graph when torch.distributed is
    initialised).  The optimizer must be built with capturable=True.  A `LearnedRandom` initializer draws its noise from the CUDA
    generator inside the graph (graph-safe Philox offsets), like the reference draws it on the device.
    """

    def __init__(self, savi, optimizer, batch, frames, resolution=(64, 64), clip_gradient=0.05, warmup=3, video_dtype=torch.float32):
        from . import ops
        self.savi, self.optimizer = savi, optimizer
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.bucket = GradientBucket(savi.parameters())
        dev = next(savi.parameters()).device
        H, W = resolution
        self.video = torch.zeros(batch, frames, 3, H, W, device=dev, dtype=video_dtype)

        def body():
            return savi_train_step(savi, optimizer, self.video, bucket=self.bucket, clip_gradient=clip_gradient)

        snap = _snapshot_training_state(savi, optimizer)
        old_flag, ops.REUSE_CONV_PACK_SCALE = ops.REUSE_CONV_PACK_SCALE, True
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    body()                                   # builds caches, workspaces, Adam state and the remembered pack scales
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            self.graph = torch.cuda.CUDAGraph()
            kw = {"capture_error_mode": "thread_local"} if self.world > 1 else {}
            n0 = ops.launch_count()
            with torch.cuda.graph(self.graph, **kw):
                self.loss, self.out = body()
            self.launches_per_replay = ops.launch_count() - n0
        finally:
            ops.REUSE_CONV_PACK_SCALE = old_flag
            _restore_training_state(savi, optimizer, snap)
            self.bucket.flat.zero_()

    def __call__(self, videos):
        """ one optimisation step on `videos` (B, T, 3, H, W); returns (loss tensor, model output dict) -- static tensors of the graph """
        self.video.copy_(videos, non_blocking=True)
        self.graph.replay()
        from . import ops
        ops.bump_weight_generation()             # eager forwards after this step must rebuild their derived weights
        return self.loss, self.out
