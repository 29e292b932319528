"""
Differentiable SAVi (src/02_train_savi.py:21-70: `model(videos)` -> MSE(recons_imgs, videos) -> `loss.backward()`): the autograd
plumbing around the kernels for the parts of the model the predictor training keeps frozen -- encoder convolutions, SlotAttention
and the decoder's weights.  Forward kernels are the ones of the inference path (tcgen05 convolutions on padded split planes, fused
LayerNorm + Linear); the backward kernels live in csrc/backward.cu, csrc/conv_tc.cu (data gradients on tcgen05) and
csrc/savi_train.cu (weight gradients, SlotAttention core, GRU gates).

Weight-derived constants (the border-class tap sums and conv1(position embedding) of the collapsed first decoder convolution,
the projected position grids) are differentiated with plain torch autograd on their small defining expressions: they are
functions of the weights only, a few hundred kB, computed once per step.
"""
import torch
import torch.nn.functional as F

from . import ops
from . import autograd as AG


def _pow2_scale(g):
    """ power of two that brings max|g| near 1 (device-side, no host synchronisation): gradients of a mean loss are ~1e-6 per
    pixel, far below fp16's normal range; convolutions are linear, so the scale is undone exactly afterwards """
    amax = g.abs().max().clamp_min(1e-30)
    return torch.exp2(-torch.floor(torch.log2(amax))).reshape(1)


# ------------------------------------------------------------------------------------------------ SlotAttention pieces

class _SACore(torch.autograd.Function):
    """ t = attention-weighted, renormalised mean of the normalised features (attention.py:100-110 with k = v = xn folded out) """

    @staticmethod
    def forward(ctx, A, xn, eps):
        A, xn = AG._c(A), AG._c(xn)
        t, Z, masks = ops.sa_core_fwd(A, xn, eps, want_masks=True)
        ctx.save_for_backward(A, xn, t, Z)
        ctx.eps = eps
        ctx.mark_non_differentiable(masks)
        return t, masks

    @staticmethod
    def backward(ctx, dt, _dmasks):
        A, xn, t, Z = ctx.saved_tensors
        dA, dxn = ops.sa_core_bwd(A, xn, t, Z, AG._c(dt), ctx.eps)
        return dA, dxn, None


class _GRUGates(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, gh, h):
        gi, gh, h = AG._c(gi), AG._c(gh), AG._c(h)
        out, gates = ops.gru_gates_fwd(gi, gh, h)
        ctx.save_for_backward(gates, gh, h)
        return out

    @staticmethod
    def backward(ctx, dout):
        gates, gh, h = ctx.saved_tensors
        return ops.gru_gates_bwd(AG._c(dout), gates, gh, h)


def slot_attention_train(sa, feats, slots, num_iters):
    """ SlotAttention.forward (attention.py:76-119) with gradients w.r.t. features, incoming slots and every parameter.
    feats (B, N, 128), slots (B, S, 128) -> (slots, masks (B, S, N) of the last iteration) """
    B, N, Df = feats.shape
    ni, ns, nm = sa.norm_input, sa.norm_slot, sa.norm_mlp
    xn = AG.layer_norm(feats.reshape(B * N, Df), ni.weight, ni.bias, ni.eps).view(B, N, Df)
    wk_t = sa.to_k.weight.t().contiguous()                    # (feat, slot): A = q W_k as a Linear with weight W_k^T
    masks = None
    for _ in range(num_iters):
        h = slots
        q = AG.linear(slots, sa.to_q.weight, None, ln=(ns.weight, ns.bias, ns.eps))
        A = AG.linear(q, wk_t) * sa.scale
        t, masks = _SACore.apply(A, xn, sa.epsilon)
        u = AG.linear(t, sa.to_v.weight)                      # sum_n w v_n = (sum_n w xn_n) W_v^T
        gi = AG.linear(u, sa.gru.weight_ih, sa.gru.bias_ih)
        gh = AG.linear(h, sa.gru.weight_hh, sa.gru.bias_hh)
        slots = _GRUGates.apply(gi, gh, h)
        m = AG.linear(slots, sa.mlp[0].weight, sa.mlp[0].bias, ln=(nm.weight, nm.bias, nm.eps), act=ops.ACT_RELU)
        slots = AG.linear(m, sa.mlp[2].weight, sa.mlp[2].bias, res=slots)
    return slots, masks


# ------------------------------------------------------------------------------------------------ encoder convolutions

class _EncoderConvs(torch.autograd.Function):
    """ frames (M, 3, H, W) [fp32 or uint8] -> tokens (M, H*W, C) = conv stack + position embedding (SAVi.py:207-213) with
    gradients w.r.t. the four convolutions' weights / biases and the position-embedding table """

    @staticmethod
    def forward(ctx, x, posemb, savi, *params):
        blocks = list(savi.encoder.encoder)
        prec = ops.TC_PRECISION[savi.conv_mode]
        M, _, H, W = x.shape
        C = blocks[0].conv.out_channels
        acts = [ops.conv_tc_workspace(M, H, W, x.device) for _ in range(3)]              # h1, h2, h3 kept for the backward
        w0, b0 = savi._tc_first_conv(blocks[0])
        ops.conv2d_split(x, w0, b0, blocks[0].conv.kernel_size[0], blocks[0]._relu, acts[0], prec)
        ops.conv5x5_tc(acts[0], savi._tc_weight(blocks[1]), savi._tc_bias64(blocks[1]), M, H, W, True, prec, act_out=acts[1])
        ops.conv5x5_tc(acts[1], savi._tc_weight(blocks[2]), savi._tc_bias64(blocks[2]), M, H, W, True, prec, act_out=acts[2])
        pe64 = posemb if C == 64 else F.pad(posemb, (0, 64 - C)).contiguous()
        tokens = torch.empty(M, H * W, 64, device=x.device, dtype=torch.float32)
        ops.conv5x5_tc(acts[2], savi._tc_weight(blocks[3]), savi._tc_bias64(blocks[3]), M, H, W, False, prec, out_tokens=tokens, addend=pe64)
        ctx.savi, ctx.acts, ctx.x, ctx.geom = savi, acts, x, (M, H, W, C, prec)
        return tokens if C == 64 else tokens[..., :C].contiguous()

    @staticmethod
    def backward(ctx, g):
        savi, acts, x = ctx.savi, ctx.acts, ctx.x
        M, H, W, C, prec = ctx.geom
        blocks = list(savi.encoder.encoder)
        g = AG._c(g)
        d_posemb = ops.colsum(g.view(M, H * W * C)).view(H * W, C)
        g64 = g if C == 64 else F.pad(g, (0, 64 - C)).contiguous()
        sc = _pow2_scale(g64)
        inv = 1.0 / sc                                            # device tensor
        ga = ops.conv_tc_workspace(M, H, W, g.device)
        gb = ops.conv_tc_workspace(M, H, W, g.device)
        ops.nhwc_to_split_scaled(g64, ga, M, H, W, prec, sc)      # d (conv4 output): no ReLU after the last encoder convolution
        grads = []
        cur, other = ga, gb
        for li in (3, 2, 1):
            dW = ops.conv5x5_planes_wgrad(cur, acts[li - 1], M, H, W, prec, prec)[:C, :C] * inv
            db = ops.planes_colsum(cur, M, H, W, prec)[:C] * inv
            grads = [dW, db] + grads
            ops.conv5x5_tc_dgrad(cur, savi._tc_weight_flipped(blocks[li]), acts[li - 1], M, H, W, prec, act_out=other)
            cur, other = other, cur
        k = blocks[0].conv.kernel_size[0]
        dW1 = ops.conv_first_wgrad(cur, x, M, H, W, k, prec)[:C] * inv
        db1 = ops.planes_colsum(cur, M, H, W, prec)[:C] * inv
        grads = [dW1, db1] + grads
        return (None, d_posemb, None) + tuple(grads)


def encoder_tokens_train(savi, frames):
    """ differentiable `encoder(x)` + position embedding; frames (M, 3, H, W) """
    blocks = list(savi.encoder.encoder)
    if len(blocks) != 4 or not savi._tc_applicable(blocks, skip_first=True, channels=savi.enc_tc_channels) \
            or savi.conv_mode not in ("fp16x3", "fp16"):
        raise NotImplementedError("SAVi training covers the shipped encoders (four plain 5x5 ConvBlocks, 32 or 64 channels) on the "
                                  "tensor-core path (conv_mode fp16x3 / fp16)")
    pe = savi.encoder_pos_embedding
    H, W = pe.resolution
    grid = pe.grid_on(pe.projection.weight.device)[0].reshape(4, H * W)                                  # (4, HW)
    posemb = grid.t() @ pe.projection.weight[:, :, 0, 0].t() + pe.projection.bias                         # (HW, C), differentiable
    params = []
    for b in blocks:
        params += [b.conv.weight, b.conv.bias]
    x = frames if frames.dtype == torch.uint8 else AG._c(frames.float())
    return _EncoderConvs.apply(x.contiguous(), posemb.contiguous(), savi, *params)


# ------------------------------------------------------------------------------------------------ decoder

def _decoder_constants_diff(savi):
    """ differentiable restatement of SAVi._decoder_constants: (wsum (ncls^2 * C1, D), P_hwc (H*W, C1)) from conv1 / posemb params """
    blocks = savi.decoder.conv_blocks()
    conv1 = blocks[0].conv
    pe = savi.decoder_pos_embedding
    w = conv1.weight
    k = w.shape[-1]
    p = k // 2

    def taps(c):
        if c < p:
            return slice(p - c, k)
        if c > p:
            return slice(0, k - (c - p))
        return slice(0, k)
    rows = [w[:, :, taps(cy), taps(cx)].sum(dim=(2, 3)) for cy in range(2 * p + 1) for cx in range(2 * p + 1)]
    wsum = torch.stack(rows, dim=0).reshape(-1, w.shape[1])
    H, W = savi.decoder_resolution
    grid = pe.grid_on(w.device)[0]                                                                        # (4, H, W)
    pos = torch.einsum("ck,khw->chw", pe.projection.weight[:, :, 0, 0], grid) + pe.projection.bias[:, None, None]
    P = F.conv2d(pos[None], w, conv1.bias, padding=p)[0]                                                  # (C1, H, W)
    return wsum, P.permute(1, 2, 0).reshape(H * W, -1)


class _DecodeFull(torch.autograd.Function):
    """ SAVi.decode with gradients w.r.t. the slots AND the decoder's parameters (decoders.py:89-113, SAVi.py:220-248) """

    @staticmethod
    def forward(ctx, slots, savi, wsum, P_hwc, *params):
        recon, recons, masks, saved = savi._decode_impl(slots.detach(), keep=True)
        ctx.savi, ctx.saved_acts, ctx.slots = savi, saved, slots.detach()
        ctx.save_for_backward(recons, masks)
        ctx.set_materialize_grads(False)
        return recon, recons, masks

    @staticmethod
    def backward(ctx, g_recon, g_recons, g_masks):
        recons, masks = ctx.saved_tensors
        d_slots, wg = ctx.savi._decode_backward(ctx.saved_acts, recons, masks, g_recon, g_recons, g_masks, weight_grads=True,
                                                slots=ctx.slots)
        return (d_slots, None, wg["wsum"], wg["P_hwc"]) + tuple(wg["params"])


def decode_train(savi, slots):
    """ differentiable decode: slots (M, S, D) -> recon, recons, masks """
    blocks = savi.decoder.conv_blocks()
    final = savi.decoder.final_conv()
    wsum, P_hwc = _decoder_constants_diff(savi)
    params = []
    for b in blocks[1:]:
        params += [b.conv.weight, b.conv.bias]
    params += [final.weight, final.bias]
    return _DecodeFull.apply(AG._c(slots.float()), savi, wsum, P_hwc, *params)
