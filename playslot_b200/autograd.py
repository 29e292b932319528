"""
Differentiable front-ends of the kernels (torch.autograd.Function plumbing around the C ABI).

The reference trains the predictor with plain PyTorch autograd (src/04_train_PlaySlot.py:32-122: SAVi frozen and
run without gradients, InvDyn + cOCVP forward, decoder with gradients w.r.t. its input slots, three losses,
`loss.backward()`).  Here every arithmetic node of that graph is one of our kernels with a hand-written backward
kernel (csrc/backward.cu); autograd only strings them together (views, concatenations, residual additions).

When no gradient is needed the functions below fall through to the fused inference kernels of `ops`.
"""
import torch

from . import ops


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def needs_grad(*tensors):
    return _needs_grad(*tensors)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# Set (only) by the training steps of playslot_b200.train around their loss.backward(): the Linear backward may then add its
# weight / bias gradients straight into the GradientBucket views.  Off everywhere else, so torch.autograd.grad(...) and
# hand-written loops that expect the gradients to be RETURNED keep plain autograd semantics.
ACCUMULATE_INTO_BUCKET = False


class accumulate_into_bucket:
    """ context manager: enable in-place accumulation of Linear weight gradients into the attached GradientBucket """

    def __enter__(self):
        global ACCUMULATE_INTO_BUCKET
        self._old, ACCUMULATE_INTO_BUCKET = ACCUMULATE_INTO_BUCKET, True

    def __exit__(self, *exc):
        global ACCUMULATE_INTO_BUCKET
        ACCUMULATE_INTO_BUCKET = self._old
        return False


def _bucket_view(param):
    """ the flat-bucket view train.GradientBucket.attach() installed as `param.grad` (a leaf Parameter still bound to it), or None """
    if not ACCUMULATE_INTO_BUCKET:
        return None
    v = getattr(param, "_ps_bucket_grad", None)
    if v is None or not param.is_leaf or param.grad is None or param.grad.data_ptr() != v.data_ptr():
        return None
    return v


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = _c(x)
        y, mean, rstd = ops.layernorm_fwd(x, _c(gamma.detach()), _c(beta.detach()), eps)
        ctx.save_for_backward(x, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(_c(dy), x, mean, rstd, _c(gamma.detach()))
        return dx, dg, db, None


# arithmetic of the differentiable Linear: "fp32" (exact FFMA kernels), "fp16x3" (tensor cores, fp32-grade) or "bf16" (tensor
# cores, one bf16 product per GEMM with fp32 accumulation, fp32 master weights and fp32 weight gradients: the "bf16" training
# mode of BASELINE.json configs[3])
TRAIN_LINEAR_MODE = "fp32"


class _Linear(torch.autograd.Function):
    """ y = x W^T + b.  Default: exact fp32.  TRAIN_LINEAR_MODE = "fp16x3": the forward GEMM and the data-gradient GEMM run on
    the tensor cores (fp32-grade split products); the weights change every optimiser step, so their packed copies are
    rebuilt per step with the power-of-two scale chosen at the first packing (no host synchronisation: capturable). """

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias = bias                 # the Parameter object (not a saved tensor): its bucket view receives the bias gradient
        ctx.mode = TRAIN_LINEAR_MODE
        w = weight if (ctx.mode != "fp32" and weight.is_contiguous()) else _c(weight.detach())     # the Parameter object carries the pack cache
        return ops.linear(x, w, None if bias is None else _c(bias.detach()), mode=ctx.mode, reuse_pack_scale=True)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        N, K = weight.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            M = dy.numel() // N
            if ctx.mode in ("fp16x3", "bf16") and M >= ops.LINEAR_TC_MIN_ROWS and N % 4 == 0 and weight.is_contiguous():
                dx = ops.linear_dgrad_tc(dy, weight, ctx.mode).view(x.shape)
            else:
                dx = ops.linear(dy, weight.detach().t().contiguous(), None, mode="fp32").view(x.shape)     # dX = dY W
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            # Parameters whose .grad is a view of the training step's flat GradientBucket accumulate straight into it: a weight
            # shared by the T roll-out steps would otherwise cost one autograd add kernel per use (~900 launches per step)
            gw = _bucket_view(weight) if ctx.needs_input_grad[1] else None
            gb = _bucket_view(ctx.bias) if ctx.has_bias else None
            if gw is not None and (not ctx.has_bias or gb is not None):
                ops.linear_wgrad_acc(dy.view(-1, N), x.view(-1, K), gw, gb)
            else:
                dw, db = ops.linear_wgrad(dy.view(-1, N), x.view(-1, K), want_bias=ctx.has_bias)
        return dx, dw, db


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, act):
        z = _c(z)
        ctx.save_for_backward(z)
        ctx.act = act
        return ops.act_fwd(z, act)

    @staticmethod
    def backward(ctx, dy):
        (z,) = ctx.saved_tensors
        return ops.act_bwd(_c(dy), z, ctx.act), None


class _MHA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, n_seq, L, n_heads, head_dim, spf):
        qkv = _c(qkv)
        ctx.save_for_backward(qkv)
        ctx.geom = (n_seq, L, n_heads, head_dim, spf)
        return ops.mha(qkv, n_seq, L, n_heads, head_dim, spf)

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        n_seq, L, H, dh, spf = ctx.geom
        return ops.mha_bwd(qkv, _c(dout), n_seq, L, H, dh, spf), None, None, None, None, None


def layer_norm(x, gamma, beta, eps):
    return _LayerNorm.apply(x, gamma, beta, eps)


def linear(x, weight, bias=None, ln=None, act=ops.ACT_NONE, res=None, res_div=1, res_mod=None):
    """ same contract as ops.linear; differentiable w.r.t. x, weight, bias, the LayerNorm parameters and res """
    ln_params = ln[:2] if ln is not None else ()
    if not _needs_grad(x, weight, bias, res, *ln_params):
        # inference: the fused kernel.  The parameter OBJECTS are passed through untouched -- ops caches the packed
        # tensor-core copy of a weight on the tensor object itself, a detach() would defeat it.
        return ops.linear(x, weight, bias, ln=ln, act=act, res=res, res_div=res_div, res_mod=res_mod)
    h = x
    if ln is not None:
        h = _LayerNorm.apply(h, ln[0], ln[1], ln[2])
    y = _Linear.apply(h, weight, bias)
    if act != ops.ACT_NONE:
        y = _Act.apply(y, act)
    if res is not None:
        N = y.shape[-1]
        r2 = res.reshape(-1, N)
        if res_div == 1 and (res_mod is None or res_mod == r2.shape[0]) and r2.shape[0] == y.numel() // N:
            y = y + r2.view(y.shape)
        else:
            M = y.numel() // N
            rows = (torch.arange(M, device=y.device) // res_div) % (res_mod or r2.shape[0])
            y = y + r2[rows].view(y.shape)
    return y


def mha(qkv, n_seq, L, n_heads, head_dim, slots_per_frame=0, res=None):
    if not _needs_grad(qkv, res):
        return ops.mha(qkv, n_seq, L, n_heads, head_dim, slots_per_frame, res=res)
    out = _MHA.apply(qkv, n_seq, L, n_heads, head_dim, slots_per_frame)
    if res is not None:
        out = out + res.reshape(out.shape)
    return out


class _InvDynTail(torch.autograd.Function):
    """ ps_invdyn_tail with gradients w.r.t. the mean / variance head outputs (and, for a trainable codebook, the
    commitment-loss gradient w.r.t. the codewords). """

    @staticmethod
    def forward(ctx, mv, noise, codebook, per_slot):
        # a private copy of the codebook: the EMA quantizer updates its codewords in place later in the same forward
        mv, codebook_c = _c(mv), codebook.detach().clone()
        out = ops.invdyn_tail(mv, noise, codebook_c, per_slot)
        ctx.save_for_backward(mv, noise, out["z"], out["idx"], codebook_c)
        ctx.per_slot = per_slot
        ctx.mark_non_differentiable(out["idx"])
        return out["action_dist"], out["z"], out["proto"], out["variability"], out["idx"], out["loss"], out["loss"].clone()

    @staticmethod
    def backward(ctx, g_dist, g_z, g_proto, g_var, _g_idx, g_quant, g_commit):
        mv, noise, z, idx, codebook = ctx.saved_tensors
        d_mv = ops.invdyn_tail_bwd(mv, noise, z, idx, codebook, g_dist, g_z, g_proto, g_var, g_quant, ctx.per_slot)
        d_cb = None
        if ctx.needs_input_grad[2]:
            # trainable (non-EMA) codebook
            A = codebook.shape[1]
            d_cb = torch.zeros_like(codebook)
            if g_commit is not None:
                zq = codebook[idx.reshape(-1)]
                contrib = (2.0 / A) * (zq - z.reshape(-1, A)) * g_commit.reshape(-1, 1)   # commit = mean_A (z_q - sg(z))^2
                d_cb.index_add_(0, idx.reshape(-1), contrib)
            if not ctx.per_slot and g_var is not None:
                # InvDynS: variability = z - E[idx] comes from a second, differentiable codebook lookup
                # (quantization.py:365-368 via LatentAction.py:118-125), so d E[idx] -= g_var.  (InvDynM passes the
                # straight-through prototype: variability = z - sg-equivalent value, no codebook gradient.)
                d_cb.index_add_(0, idx.reshape(-1), -g_var.reshape(-1, A))
        return d_mv, None, d_cb, None


def invdyn_tail(mv, noise, codebook, per_slot):
    """ dict like ops.invdyn_tail plus separate "quant" / "commit" per-vector losses; differentiable when needed """
    if not _needs_grad(mv, codebook):
        out = ops.invdyn_tail(mv.detach(), noise, codebook.detach(), per_slot)
        out["quant"], out["commit"] = out["loss"], out["loss"].clone()
        return out
    dist, z, proto, var, idx, quant, commit = _InvDynTail.apply(mv, noise, codebook, per_slot)
    return {"action_dist": dist, "z": z, "proto": proto, "variability": var, "idx": idx, "quant": quant, "commit": commit}


class _MSELoss(torch.autograd.Function):
    """ nn.MSELoss(pred, target) (mean over all elements; PredImgMSELoss / PredSlotMSELoss, src/lib/loss.py:201-238) as one
    reduction pass and one backward pass; the target may be a time slice of the clip (read in place), no gradient w.r.t. it """

    @staticmethod
    def forward(ctx, pred, target):
        pred = _c(pred)
        sq = ops.frame_sq_err(pred, target)
        ctx.save_for_backward(pred, target)
        ctx.scale = 1.0 / pred.numel()
        return sq.sum() * ctx.scale

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        return ops.mse_bwd(pred, target, _c(g.reshape(1)), ctx.scale), None


def mse_loss(pred, target):
    """ (B, F, ...) tensors -> scalar mean squared error, differentiable w.r.t. pred """
    if not _needs_grad(pred):
        return ops.frame_sq_err(_c(pred), target).sum() / pred.numel()
    return _MSELoss.apply(pred, target)
