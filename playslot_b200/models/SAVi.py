"""
SAVi object-centric video decomposition -- mirror of src/models/SAVi.py (same constructor kwargs = the JSON
`model_params`, same attributes, forward / encode / decode / broadcast signatures and state_dict layout).

Kernel mapping (csrc/):
  encode : 4x ps_conv2d (the last one writes NHWC tokens and adds the weight-only position embedding in its
           epilogue) -> ps_linear (LayerNorm fused) + ReLU -> ps_linear                 [SAVi.py:207-217]
  slots  : SlotAttention (ps_sa_main / ps_sa_update), transition block (ps_linear / ps_mha)  [:172-195]
  decode : first conv collapsed on the broadcast input (ps_linear + ps_dec_expand), 3x ps_conv2d,
           ps_dec_composite (last conv + slot softmax + compositing)                       [:220-248]
The image encoder has no temporal dependency, so `forward` encodes all frames of the clip in one batched
pass (time-major) before running the recurrent slot loop.
"""
import math
import os

import torch
import torch.nn as nn

from .. import ops
from .. import autograd as AG
from ._common import DerivedWeights, require_cuda, no_grad_only, f32c
from .BlocksUtils.attention import SlotAttention
from .BlocksUtils.initializers import get_initializer
from .BlocksUtils.model_blocks import SoftPositionEmbed
from .BlocksUtils.model_utils import init_xavier_
from .BlocksUtils.transition_models import get_transition_module
from .EncodersDecoders.encoders import get_encoder
from .EncodersDecoders.decoders import get_decoder


class _DecodeFn(torch.autograd.Function):
    """ SAVi.decode with a data gradient w.r.t. the slots (the decoder is frozen in 04_train_PlaySlot.py:87-90) """

    @staticmethod
    def forward(ctx, slots, savi):
        # predictor training: the decoder is frozen, so the last block's activation is kept as split planes (its sign is all the
        # data gradient needs) and the 3x3 convolution + compositing run on the tensor cores like in inference
        recon, recons, masks, saved = savi._decode_impl(slots.detach(), keep="planes")
        ctx.savi, ctx.saved_acts = savi, saved
        ctx.save_for_backward(recons, masks)
        ctx.set_materialize_grads(False)
        return recon, recons, masks

    @staticmethod
    def backward(ctx, g_recon, g_recons, g_masks):
        recons, masks = ctx.saved_tensors
        return ctx.savi._decode_backward(ctx.saved_acts, recons, masks, g_recon, g_recons, g_masks), None


class SAVi(nn.Module):
    """
    Args mirror the reference (SAVi.py:29-56): num_slots, slot_dim, num_iterations, num_iterations_first,
    in_channels, mlp_hidden, mlp_encoder_dim, initializer (str), encoder / decoder (dicts),
    transition_module_params (dict).
    """

    ENCODE_CHUNK_BYTES = 6 << 30      # feature tokens kept per batched encode pass
    DECODE_CHUNK_IMAGES = 4096        # slot-images decoded per pass (1 MiB of fp32 activations each at 64x64x64)

    # arithmetic of the 64 -> 64 channel 5x5 convolutions:
    #   "fp16x3": tcgen05 tensor cores, fp16 hi + scaled fp16 lo operands, three products, fp32 accumulate
    #             (22 mantissa bits: fp32-grade, passes the 1e-4 parity bar)
    #   "fp16" / "bf16": tcgen05 tensor cores, single half-precision product (stated tolerances)
    #   "fp32"  : exact fp32 FFMA kernel on the CUDA cores
    CONV_MODES = ("fp16x3", "fp16", "bf16", "fp32")

    def __init__(self, num_slots, slot_dim, num_iterations=3, num_iterations_first=None,
                 in_channels=3, mlp_hidden=128, mlp_encoder_dim=128,
                 initializer=None, encoder=None, decoder=None, transition_module_params=None):
        super().__init__()
        self.num_slots = num_slots
        self.slot_dim = slot_dim
        self.num_iterations_first = num_iterations_first
        self.num_iterations = num_iterations
        self.in_channels = in_channels
        self.mlp_hidden = mlp_hidden
        self.mlp_encoder_dim = mlp_encoder_dim

        self.initializer = get_initializer(mode=initializer, slot_dim=slot_dim, num_slots=num_slots)
        self.transition_module = get_transition_module(slot_dim=slot_dim, num_slots=num_slots, **transition_module_params)
        self.build_encoder(encoder_params=encoder)
        self.build_decoder(decoder_params=decoder)
        self.slot_attention = SlotAttention(
                dim_feats=self.mlp_encoder_dim,
                dim_slots=self.slot_dim,
                num_slots=self.num_slots,
                num_iters_first=self.num_iterations_first,
                num_iters=self.num_iterations,
                mlp_hidden=self.mlp_hidden,
            )
        self._dec_cache = DerivedWeights()
        self.conv_mode = os.environ.get("PLAYSLOT_CONV_MODE", "fp16x3")
        if self.conv_mode not in self.CONV_MODES:
            raise ValueError(f"PLAYSLOT_CONV_MODE={self.conv_mode!r}; expected one of {self.CONV_MODES}")
        self._tc_ws = {}
        self._tc_w = {}
        self._init_model()

    def build_encoder(self, encoder_params):
        self.encoder = get_encoder(in_channels=self.in_channels, encoder=encoder_params)
        out_feats = self.encoder.out_features
        self.encoder_pos_embedding = SoftPositionEmbed(
                hidden_size=out_feats,
                resolution=encoder_params["encoder_params"].get("resolution")
            )
        self.encoder_mlp = nn.Sequential(
                nn.LayerNorm(out_feats),
                nn.Linear(out_feats, self.mlp_encoder_dim),
                nn.ReLU(),
                nn.Linear(self.mlp_encoder_dim, self.mlp_encoder_dim),
            )

    def build_decoder(self, decoder_params):
        self.decoder_resolution = decoder_params["decoder_params"].get("resolution")
        self.decoder_pos_embedding = SoftPositionEmbed(hidden_size=self.slot_dim, resolution=self.decoder_resolution)
        self.decoder = get_decoder(in_channels=self.slot_dim, decoder=decoder_params)

    # --------------------------------------------------------------------------------------------- forward
    def forward(self, x, num_imgs=10, decode=True, **kwargs):
        """
        x (B, T, C, H, W) -> dict(recons_imgs (B,num_imgs,C,H,W), recons_objs (B,num_imgs,S,C,H,W),
        masks (B,num_imgs,S,1,H,W), slot_history (B,num_imgs,S,D)); with decode=False the first three are
        stacks of empty tensors, as in the reference (SAVi.py:135-204).
        """
        require_cuda(x, "SAVi.forward")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(x, num_imgs, decode, **kwargs)          # SAVi training (02_train_savi.py:21-70)
        B = x.shape[0]
        predicted = self.initializer(batch_size=B, **kwargs)
        predicted = f32c(predicted.detach())

        slot_history, recons_h, objs_h, masks_h = [], [], [], []
        use_planes = self.slot_attention.mode == "tc"
        # use_planes: norm_input statistics + fp16 hi/lo split of every frame once per chunk (fused into the encoder MLP
        # kernel when that applies); the per-frame SlotAttention passes then stream these planes by TMA
        feats_all = self._encode_clip(x, num_imgs, planes=use_planes)     # list of (t0, shape (t1-t0, B, N, D), data) chunks
        for t in range(num_imgs):
            if use_planes:
                t0, shp, planes = next(c for c in feats_all if c[0] <= t < c[0] + c[1][0])
                feats = ("planes", planes, (t - t0) * shp[1] * shp[2], shp[1], shp[2])
            else:
                feats = next(f[t - t0] for t0, _, f in feats_all if t0 <= t < t0 + f.shape[0])
            slots = self.slot_attention(inputs=feats, slots=predicted, step=t)
            if decode:
                recon, (recons, masks) = self.decode(slots)
            else:
                recon, recons, masks = torch.tensor([]), torch.tensor([]), torch.tensor([])
            predicted = self.transition_module(slots)
            slot_history.append(slots)
            recons_h.append(recon)
            objs_h.append(recons)
            masks_h.append(masks)
        return {
            "recons_imgs": torch.stack(recons_h, dim=1),
            "recons_objs": torch.stack(objs_h, dim=1),
            "masks": torch.stack(masks_h, dim=1),
            "slot_history": torch.stack(slot_history, dim=1),
        }

    def _forward_train(self, x, num_imgs, decode, **kwargs):
        """ differentiable forward (every parameter receives a gradient): same recurrence as `forward`, composed from the
        autograd front-ends in playslot_b200.autograd / savi_autograd (fp32 feature tokens instead of the fused plane hand-over) """
        from .. import savi_autograd as SG
        B, _, C, H, W = x.shape
        N = H * W
        predicted = self.initializer(batch_size=B, **kwargs).float()
        frames = x[:, :num_imgs].transpose(0, 1).reshape(num_imgs * B, C, H, W)
        tokens = SG.encoder_tokens_train(self, frames)                                # (T*B, N, C)
        ln, fc1, fc2 = self.encoder_mlp[0], self.encoder_mlp[1], self.encoder_mlp[3]
        hdn = AG.linear(tokens, fc1.weight, fc1.bias, ln=(ln.weight, ln.bias, ln.eps), act=ops.ACT_RELU)
        feats = AG.linear(hdn, fc2.weight, fc2.bias).view(num_imgs, B, N, -1)
        slot_history, recons_h, objs_h, masks_h = [], [], [], []
        for t in range(num_imgs):
            slots = self.slot_attention(inputs=feats[t], slots=predicted, step=t)
            if decode:
                recon, (recons, masks) = self.decode(slots)
            else:
                recon, recons, masks = torch.tensor([]), torch.tensor([]), torch.tensor([])
            predicted = self.transition_module(slots)
            slot_history.append(slots)
            recons_h.append(recon)
            objs_h.append(recons)
            masks_h.append(masks)
        return {
            "recons_imgs": torch.stack(recons_h, dim=1),
            "recons_objs": torch.stack(objs_h, dim=1),
            "masks": torch.stack(masks_h, dim=1),
            "slot_history": torch.stack(slot_history, dim=1),
        }

    def _encode_clip(self, x, num_imgs, planes=False):
        """ batched, time-major encode of frames [0, num_imgs) in chunks bounded by ENCODE_CHUNK_BYTES; each chunk is
        (t0, (frames, B, N, D), fp32 features) or, with planes=True, (t0, shape, pre-normalised fp16 hi/lo planes) """
        B, _, C, H, W = x.shape
        N = H * W
        per_frame = B * N * self.mlp_encoder_dim * 4
        step = max(1, min(num_imgs, self.ENCODE_CHUNK_BYTES // max(per_frame, 1)))
        chunks = []
        for t0 in range(0, num_imgs, step):
            t1 = min(num_imgs, t0 + step)
            frames = x[:, t0:t1].transpose(0, 1).reshape((t1 - t0) * B, C, H, W)
            shape = (t1 - t0, B, N, self.mlp_encoder_dim)
            if planes:
                chunks.append((t0, shape, self.encode(frames, planes=True)))
            else:
                chunks.append((t0, shape, self.encode(frames).view(shape)))
        return chunks

    # --------------------------------------------------------------------------------------------- encode
    def encode(self, x, planes=False):
        """ images (M, C, H, W) -> feature tokens (M, H*W, mlp_encoder_dim)   (SAVi.py:207-217).
        planes=True (internal): the norm_input-normalised features as fp16 hi/lo planes (2, M*H*W, 128) for
        SlotAttention's "tc" mode instead """
        require_cuda(x, "SAVi.encode")
        no_grad_only("SAVi.encode", x, self.encoder_mlp[1].weight)
        # uint8 frames (0..255): the ToTensor normalisation x / 255 of the reference's DataLoader (data/load_data.py:86-91) is fused
        # into the first convolution when it runs as the split-plane producer; otherwise normalise here (same IEEE division)
        u8_fused = (x.dtype == torch.uint8 and x.shape[0] <= 65535 and x.shape[2] % 16 == 0 and x.shape[3] % 16 == 0
                    and self._tc_applicable(self.encoder.encoder, skip_first=True, channels=self.enc_tc_channels))
        if x.dtype == torch.uint8:
            x = x.contiguous() if u8_fused else x.float().div(255)
        if not u8_fused:
            x = f32c(x)
        M, _, H, W = x.shape
        posemb = self.encoder_pos_embedding.table(channels_last=True)
        ln, fc1, fc2 = self.encoder_mlp[0], self.encoder_mlp[1], self.encoder_mlp[3]
        C = fc1.in_features
        tc_convs = self._tc_applicable(self.encoder.encoder, skip_first=True, channels=self.enc_tc_channels)
        fused = (planes and ops.linear_mode() == "fp16x3" and self.mlp_encoder_dim == 128 and fc1.out_features == 128
                 and C <= 64 and C % 16 == 0 and M * H * W >= ops.LINEAR_TC_MIN_ROWS)
        paired = tc_convs and self._pairable(x)
        if paired:
            # 32-channel encoder: two images share the 64 channels of the tensor-core convolution (block-diagonal weights)
            tokens = self._encode_convs_tc_paired(x, posemb)                      # (ceil(M/2), H*W, 2 images, 32)
            if fused:
                planes = ops.enc_mlp_planes(tokens.view(-1, C), C, (ln.weight, ln.bias, ln.eps), (fc1.weight, fc1.bias),
                                            (fc2.weight, fc2.bias), self.slot_attention.norm_input.eps, pair_hw=H * W)
                return planes if M % 2 == 0 else planes[:, :M * H * W].contiguous()
            tokens = tokens.permute(0, 2, 1, 3).reshape(-1, H * W, C)[:M].contiguous()
        elif tc_convs:
            tokens = self._encode_convs_tc(x, posemb, keep_padded=fused)          # (M, H*W, 64 if fused else C)
        else:
            tokens = self.encoder.run(x, last_nhwc_addend=posemb)
        if fused:
            # LayerNorm -> Linear -> ReLU -> Linear -> norm_input statistics -> split, one kernel, activations stay on chip
            return ops.enc_mlp_planes(tokens.view(M * H * W, -1), C, (ln.weight, ln.bias, ln.eps), (fc1.weight, fc1.bias),
                                      (fc2.weight, fc2.bias), self.slot_attention.norm_input.eps)
        tokens = tokens.view(M, H * W, -1)
        h = ops.linear(tokens, fc1.weight, fc1.bias, ln=(ln.weight, ln.bias, ln.eps), act=ops.ACT_RELU)
        feats = ops.linear(h, fc2.weight, fc2.bias)
        return self.slot_attention.prepare(feats) if planes else feats

    # --------------------------------------------------------------------------------------------- tensor-core convs
    enc_tc_channels = (32, 64)      # encoder widths routed to the tcgen05 convolution (32 is zero-padded to 64)

    def _tc_applicable(self, blocks, skip_first, channels=(64,)):
        """ the tcgen05 kernel covers plain 5x5, 64 -> 64 channel ConvBlocks; 32-channel encoder stacks run on it with
        their weights zero-padded to 64 channels (4x the useful MMA work is still ~4x faster than the FFMA kernel) """
        if self.conv_mode not in ops.TC_PRECISION:
            return False
        blocks = list(blocks)[1:] if skip_first else list(blocks)
        if not blocks:
            return False
        ch = blocks[0].conv.in_channels if getattr(blocks[0], "conv", None) is not None else None
        if ch not in channels:
            return False
        for b in blocks:
            c = getattr(b, "conv", None)
            if c is None or not b._plain or c.kernel_size != (5, 5) or c.in_channels != ch or c.out_channels != ch:
                return False
        return True

    def _tc_workspace(self, tag, n_img, H, W, device):
        key = (tag, n_img, H, W, str(device))
        if key not in self._tc_ws:
            self._tc_ws[key] = ops.conv_tc_workspace(n_img, H, W, device)
        return self._tc_ws[key]

    @staticmethod
    def _pad64(w):
        """ zero-pad a (Cout, Cin, k, k) weight to 64 x 64 channels """
        if w.shape[0] == 64 and w.shape[1] == 64:
            return w
        out = torch.zeros(64, 64, *w.shape[2:], device=w.device, dtype=w.dtype)
        out[:w.shape[0], :w.shape[1]] = w
        return out

    def _tc_weight(self, block):
        prec = ops.TC_PRECISION[self.conv_mode]
        cache = self._tc_w.setdefault((id(block), prec), DerivedWeights())
        w = block.conv.weight
        return cache.get([w], lambda: ops.pack_conv_weight_tc(self._pad64(w.detach()), prec, owner=w))

    def _tc_first_conv(self, block):
        """ (packed weight, bias) of the first encoder conv with its output channels zero-padded to 64 """
        cache = self._tc_w.setdefault((id(block), "first"), DerivedWeights())
        c = block.conv

        def build():
            w = c.weight.detach()
            if w.shape[0] == 64:
                return block.packed_weight(), c.bias.detach()
            wp = torch.zeros(64, *w.shape[1:], device=w.device, dtype=w.dtype)
            wp[:w.shape[0]] = w
            bp = torch.zeros(64, device=w.device, dtype=w.dtype)
            bp[:w.shape[0]] = c.bias.detach()
            return ops.pack_conv_weight(wp), bp
        return cache.get([c.weight, c.bias], build)

    def _tc_bias64(self, block):
        cache = self._tc_w.setdefault((id(block), "bias64"), DerivedWeights())
        b = block.conv.bias

        def build():
            if b.shape[0] == 64:
                return b.detach()
            out = torch.zeros(64, device=b.device, dtype=b.dtype)
            out[:b.shape[0]] = b.detach()
            return out
        return cache.get([b], build)

    PAIR_32CH = True                # 32-channel encoders: convolve two images per 64-channel plane (False: zero-pad each image to 64)

    def _pairable(self, x):
        blocks = list(self.encoder.encoder)
        M, _, H, W = x.shape
        return (self.PAIR_32CH and blocks[0].conv.out_channels == 32 and M >= 2 and (M + 1) // 2 <= 65535
                and H % 16 == 0 and W % 16 == 0)

    def _pair_weights(self, block, first):
        """ block-diagonal 64-channel weights / duplicated bias of a 32-channel ConvBlock: image 2p of a pair lives in channels
        0..31, image 2p+1 in channels 32..63 (the zero blocks contribute exact zeros, so both images see their own convolution) """
        prec = ops.TC_PRECISION[self.conv_mode]
        cache = self._tc_w.setdefault((id(block), "pair", first, prec), DerivedWeights())
        c = block.conv

        def build():
            w = c.weight.detach()
            co, ci = w.shape[0], w.shape[1]
            wp = torch.zeros(2 * co, 2 * ci, *w.shape[2:], device=w.device, dtype=w.dtype)
            wp[:co, :ci] = w
            wp[co:, ci:] = w
            bp = torch.cat([c.bias.detach(), c.bias.detach()]).contiguous()
            return (ops.pack_conv_weight(wp) if first else ops.pack_conv_weight_tc(wp, prec, owner=c.weight)), bp
        return cache.get([c.weight, c.bias], build)

    def _encode_convs_tc_paired(self, x, posemb):
        """ the encoder stack of a 32-channel configuration on image PAIRS: (M, 3, H, W) is read as (M/2, 6, H, W), every layer runs
        with block-diagonal weights, and the tokens come back as (M/2, H*W, 2, 32) -- half the tensor-core work of padding each
        image to 64 channels.  An odd batch gets one zero image appended. """
        prec = ops.TC_PRECISION[self.conv_mode]
        blocks = list(self.encoder.encoder)
        M, Cin, H, W = x.shape
        if M % 2:
            x = torch.cat([x, torch.zeros_like(x[:1])], dim=0)
        Mp = x.shape[0] // 2
        xp = x.view(Mp, 2 * Cin, H, W)
        a = self._tc_workspace("enc_a", Mp, H, W, x.device)
        b = self._tc_workspace("enc_b", Mp, H, W, x.device)
        w0, b0 = self._pair_weights(blocks[0], True)
        ops.conv2d_split(xp, w0, b0, blocks[0].conv.kernel_size[0], blocks[0]._relu, a, prec)
        for blk in blocks[1:-1]:
            w, bias = self._pair_weights(blk, False)
            ops.conv5x5_tc(a, w, bias, Mp, H, W, blk._relu, prec, act_out=b)
            a, b = b, a
        last = blocks[-1]
        if not hasattr(self, "_posemb_pair"):
            self._posemb_pair = DerivedWeights()
        pe = self.encoder_pos_embedding.projection
        posemb2 = self._posemb_pair.get([pe.weight, pe.bias], lambda: torch.cat([posemb, posemb], dim=-1).contiguous())
        tokens = torch.empty(Mp, H * W, 64, device=x.device, dtype=torch.float32)
        w, bias = self._pair_weights(last, False)
        ops.conv5x5_tc(a, w, bias, Mp, H, W, last._relu, prec, out_tokens=tokens, addend=posemb2)
        return tokens.view(Mp, H * W, 2, 32)

    def _encode_convs_tc(self, x, posemb, keep_padded=False):
        """ first conv (3 -> C) on the CUDA cores writing the split planes, the C -> C layers on tcgen05; tokens + posemb out.
        C = 32 stacks are zero-padded to 64 channels. """
        prec = ops.TC_PRECISION[self.conv_mode]
        blocks = list(self.encoder.encoder)
        M, _, H, W = x.shape
        C = blocks[0].conv.out_channels
        a = self._tc_workspace("enc_a", M, H, W, x.device)
        b = self._tc_workspace("enc_b", M, H, W, x.device)
        c0 = blocks[0].conv
        if M <= 65535 and H % 16 == 0 and W % 16 == 0:
            # first conv (FFMA) writes the split planes of the tensor-core layers directly
            w0, b0 = self._tc_first_conv(blocks[0])
            ops.conv2d_split(x, w0, b0, c0.kernel_size[0], blocks[0]._relu, a, prec)
        else:
            if C != 64:
                raise NotImplementedError("32-channel encoders on the tensor-core path need H, W multiples of 16 and <= 65535 images per pass")
            h = blocks[0].run(x)                                      # fp32 NCHW (M, 64, H, W)
            ops.nchw_to_split(h, a, prec)
            del h
        for blk in blocks[1:-1]:
            ops.conv5x5_tc(a, self._tc_weight(blk), self._tc_bias64(blk), M, H, W, blk._relu, prec, act_out=b)
            a, b = b, a
        last = blocks[-1]
        if C != 64:
            if not hasattr(self, "_posemb64"):
                self._posemb64 = DerivedWeights()
            pe = self.encoder_pos_embedding.projection
            posemb = self._posemb64.get([pe.weight, pe.bias], lambda: torch.nn.functional.pad(posemb, (0, 64 - C)).contiguous())
        tokens = torch.empty(M, H * W, 64, device=x.device, dtype=torch.float32)
        ops.conv5x5_tc(a, self._tc_weight(last), self._tc_bias64(last), M, H, W, last._relu, prec, out_tokens=tokens, addend=posemb)
        return tokens if (C == 64 or keep_padded) else tokens[..., :C].contiguous()

    # --------------------------------------------------------------------------------------------- decode
    def _decoder_constants(self):
        """
        Weight-only tensors of the collapsed first decoder convolution (cached per weight version):
          Wsum (ncls*ncls*C1, D): tap-sums of conv1's weights for the (2p+1)^2 border classes,
          P (C1, H, W): conv1(position embedding) + bias.
        """
        blocks = self.decoder.conv_blocks()
        conv1 = blocks[0].conv
        pe = self.decoder_pos_embedding
        srcs = [conv1.weight, conv1.bias, pe.projection.weight, pe.projection.bias]

        def build():
            w = conv1.weight.detach()                      # (C1, D, k, k)
            k = w.shape[-1]
            p = k // 2

            def taps(c):                                   # valid kernel taps along one axis for border class c
                if c < p:
                    return slice(p - c, k)
                if c > p:
                    return slice(0, k - (c - p))
                return slice(0, k)
            rows = []
            for cy in range(2 * p + 1):
                for cx in range(2 * p + 1):
                    rows.append(w[:, :, taps(cy), taps(cx)].sum(dim=(2, 3)))     # (C1, D)
            wsum = torch.stack(rows, dim=0).reshape(-1, w.shape[1]).contiguous()  # (ncls^2 * C1, D)
            pos = pe.table(channels_last=False)[None].contiguous()               # (1, D, H, W)
            P = ops.conv2d(pos, blocks[0].packed_weight(), conv1.bias, k, relu=False)   # conv1(posemb) + bias
            P = P[0].contiguous()
            return wsum, P, k, P.permute(1, 2, 0).reshape(-1, P.shape[0]).contiguous()
        return self._dec_cache.get(srcs, build)

    def decode(self, slots):
        """
        slots (M, S, D) -> recon_combined (M, C, H, W), (recons (M, S, C, H, W) un-masked, masks (M, S, 1, H, W))
        (SAVi.py:220-234).  Differentiable w.r.t. `slots` (data gradient only: the training script keeps SAVi frozen).
        """
        require_cuda(slots, "SAVi.decode")
        if torch.is_grad_enabled() and (self.decoder.final_conv().weight.requires_grad
                                        or self.decoder_pos_embedding.projection.weight.requires_grad):
            # SAVi training: gradients w.r.t. the slots and the decoder's parameters
            from .. import savi_autograd as SG
            recon, recons, masks = SG.decode_train(self, slots)
            return recon, (recons, masks)
        if torch.is_grad_enabled() and slots.requires_grad:
            recon, recons, masks = _DecodeFn.apply(f32c(slots), self)         # predictor training: frozen decoder, data gradient
            return recon, (recons, masks)
        recon, recons, masks, _ = self._decode_impl(f32c(slots.detach()), keep=False)
        return recon, (recons, masks)

    def _decode_impl(self, slots, keep):
        """ forward kernels; with keep (True: SAVi training, "planes": predictor training) the per-layer activations the backward
        needs are returned.  In the fp16x3 mode the last block's activation stays in the split-plane format as well -- saved as
        ("planes", tensor) -- and the 3x3 convolution + compositing run on the tensor cores as in inference; the final data gradient
        takes its ReLU mask, and SAVi training its final weight gradient, from those planes.  Single-product modes keep an fp32 NCHW
        copy and the FFMA compositing kernel. """
        M, S, D = slots.shape
        H, W = self.decoder_resolution
        dec = self.decoder
        if dec.upsample not in (None, 1) or self.in_channels != 3:
            raise NotImplementedError("SAVi.decode kernels cover upsample in (None, 1) and RGB outputs "
                                      "(every shipped config); got upsample=%r in_channels=%r" % (dec.upsample, self.in_channels))
        blocks = dec.conv_blocks()
        wsum, P, k, P_hwc = self._decoder_constants()
        C1 = blocks[0].conv.out_channels
        use_tc = C1 == 64 and self._tc_applicable(blocks, skip_first=True)
        prec = ops.TC_PRECISION.get(self.conv_mode, 0)
        if keep and not (use_tc and prec in (2, 3) and len(blocks) == 4 and all(b._relu for b in blocks)):
            raise NotImplementedError("the decoder data gradient is implemented for the tensor-core path (conv_mode fp16x3 / fp16) "
                                      "of the 4-block 64-channel decoder of the shipped configs")
        final = dec.final_conv()
        wfin = self._final_packed(final)
        recon_l, recons_l, masks_l, saved = [], [], [], []
        step = max(1, self.DECODE_CHUNK_IMAGES // S)
        for m0 in range(0, M, step):
            m1 = min(M, m0 + step)
            sl = slots[m0:m1].reshape((m1 - m0) * S, D)
            G = ops.linear(sl, wsum)                                           # (n_img, ncls^2 * C1)
            n_img = (m1 - m0) * S
            if use_tc:
                planes_last = bool(keep) and prec == 3 and S <= 16      # both training paths: h4 stays in the split-plane format
                if keep:
                    acts = [self._acquire_train_ws(n_img, H, W, slots.device) for _ in range(4 if planes_last else 3)]     # h1, h2, h3 (, h4)
                else:
                    a = self._tc_workspace("dec_a", n_img, H, W, slots.device)
                    b = self._tc_workspace("dec_b", n_img, H, W, slots.device)
                    acts = [a if i % 2 == 0 else b for i in range(len(blocks) - 1)]
                ops.dec_expand_split(G, P_hwc, acts[0], n_img, H, W, k, prec)
                for i, blk in enumerate(blocks[1:-1]):
                    ops.conv5x5_tc(acts[i], self._tc_weight(blk), blk.conv.bias, n_img, H, W, blk._relu, prec, act_out=acts[i + 1])
                if planes_last:
                    ops.conv5x5_tc(acts[2], self._tc_weight(blocks[-1]), blocks[-1].conv.bias, n_img, H, W, blocks[-1]._relu, prec, act_out=acts[3])
                    r, rs, ms = ops.dec_composite_tc(acts[3], self._final_packed_tc(final), final.bias, m1 - m0, S, H, W)
                    recon_l.append(r); recons_l.append(rs); masks_l.append(ms)
                    saved.append((m0, m1, acts[:3], ("planes", acts[3])))
                    continue
                if not keep and prec == 3 and S <= 16:
                    # inference: the last block stays in the split format and the 3x3 conv + slot softmax + compositing
                    # run on the tensor cores too (no fp32 NCHW round trip of the 64-channel activation)
                    last_in = acts[len(blocks) - 2]
                    last_out = b if last_in is a else a                        # ping-pong workspaces
                    ops.conv5x5_tc(acts[len(blocks) - 2], self._tc_weight(blocks[-1]), blocks[-1].conv.bias, n_img, H, W, blocks[-1]._relu, prec, act_out=last_out)
                    r, rs, ms = ops.dec_composite_tc(last_out, self._final_packed_tc(final), final.bias, m1 - m0, S, H, W)
                    recon_l.append(r); recons_l.append(rs); masks_l.append(ms)
                    continue
                h = torch.empty(n_img, 64, H, W, device=slots.device, dtype=torch.float32)
                ops.conv5x5_tc(acts[len(blocks) - 2], self._tc_weight(blocks[-1]), blocks[-1].conv.bias, n_img, H, W, blocks[-1]._relu, prec, out_nchw=h)
            else:
                h = ops.dec_expand(G, P, n_img, C1, H, W, k)
                for blk in blocks[1:]:
                    h = blk.run(h)
            r, rs, ms = ops.dec_composite(h, wfin, final.bias, m1 - m0, S)
            recon_l.append(r); recons_l.append(rs); masks_l.append(ms)
            if keep:
                saved.append((m0, m1, acts, h))
        if len(recon_l) == 1:
            return recon_l[0], recons_l[0], masks_l[0], saved
        return torch.cat(recon_l), torch.cat(recons_l), torch.cat(masks_l), saved

    def _acquire_train_ws(self, n_img, H, W, device):
        """ activation planes kept for a backward pass: taken from a pool that `_decode_backward` refills, so a training step does
        not zero-fill ~10 GB of fresh workspaces (the producers keep the borders zero, like in the cached inference workspaces) """
        pool = self.__dict__.setdefault("_train_ws_pool", {}).setdefault((n_img, H, W, str(device)), [])
        return pool.pop() if pool else ops.conv_tc_workspace(n_img, H, W, device)

    def _release_train_ws(self, tensors, n_img, H, W, device):
        pool = self.__dict__.setdefault("_train_ws_pool", {}).setdefault((n_img, H, W, str(device)), [])
        for t in tensors:
            if all(t is not q for q in pool):
                pool.append(t)

    def _tc_weight_flipped(self, block):
        """ packed weights of the data gradient: W'[ci][co][ky][kx] = W[co][ci][K-1-ky][K-1-kx] """
        prec = ops.TC_PRECISION[self.conv_mode]
        cache = self._tc_w.setdefault((id(block), prec, "flip"), DerivedWeights())
        w = block.conv.weight
        return cache.get([w], lambda: ops.pack_conv_weight_tc(self._pad64(w.detach()).flip(2, 3).transpose(0, 1).contiguous(), prec, owner=w))

    def _decode_backward(self, saved, recons, masks, g_recon, g_recons, g_masks, weight_grads=False, slots=None):
        """ d loss / d slots through compositing, the 3x3 and the three 5x5 convolutions and the collapsed first one; with
        weight_grads also the gradients of every decoder parameter (returned as (d_slots, dict)) """
        M, S = recons.shape[:2]
        H, W = self.decoder_resolution
        blocks = self.decoder.conv_blocks()
        wsum, _, k, _ = self._decoder_constants()
        if not hasattr(self, "_wsum_t_cache"):
            self._wsum_t_cache = DerivedWeights()
        wsum_t = self._wsum_t_cache.get([wsum], lambda: wsum.t().contiguous())            # (D, ncls^2 * C1)
        wfin = self._final_packed(self.decoder.final_conv())
        prec = ops.TC_PRECISION[self.conv_mode]
        cg = lambda g, m0, m1: None if g is None else f32c(g[m0:m1])
        out = torch.empty(M, S, self.slot_dim, device=recons.device, dtype=torch.float32)
        wg = None
        if weight_grads:
            dev = recons.device
            wg = {"wsum": torch.zeros_like(wsum), "P_hwc": torch.zeros(H * W, 64, device=dev),
                  "conv": [[torch.zeros(64, 64, 5, 5, device=dev), torch.zeros(64, device=dev)] for _ in range(3)],
                  "final": [torch.zeros(4, 64, 3, 3, device=dev), torch.zeros(4, device=dev)]}
        for m0, m1, acts, h in saved:
            n_img = (m1 - m0) * S
            g_o = ops.dec_composite_bwd(f32c(recons[m0:m1]), f32c(masks[m0:m1]), cg(g_recon, m0, m1), cg(g_recons, m0, m1), cg(g_masks, m0, m1))
            ga = self._tc_workspace("dec_ga", n_img, H, W, recons.device)
            gb = self._tc_workspace("dec_gb", n_img, H, W, recons.device)
            # power-of-two gradient scale computed on the device (no host sync): MSE-mean gradients are ~1e-6 per pixel,
            # far below fp16's normal range; the convolutions are linear, so the scale is undone exactly at the end
            amax = g_o.abs().max().clamp_min(1e-30)
            g_scale = torch.exp2(-torch.floor(torch.log2(amax))).reshape(1)
            if isinstance(h, tuple):                                                               # ("planes", h4 as split planes)
                ops.dec_final_dgrad_planes(g_o, wfin, h[1], ga, n_img, H, W, prec, g_scale)
            else:
                ops.dec_final_dgrad(g_o, wfin, h, ga, prec, g_scale)                               # through conv3x3 and relu(h4)
            inv = 1.0 / g_scale
            if wg is not None:
                dwf, dbf = (ops.conv3x3_final_wgrad_planes(g_o, h[1], n_img, H, W, prec) if isinstance(h, tuple)
                            else ops.conv3x3_final_wgrad(g_o, h))
                wg["final"][0] += dwf; wg["final"][1] += dbf
                wg["conv"][2][0] += ops.conv5x5_planes_wgrad(ga, acts[2], n_img, H, W, prec, prec) * inv     # block 3: h3 -> h4
                wg["conv"][2][1] += ops.planes_colsum(ga, n_img, H, W, prec) * inv
            ops.conv5x5_tc_dgrad(ga, self._tc_weight_flipped(blocks[3]), acts[2], n_img, H, W, prec, act_out=gb)
            if wg is not None:
                wg["conv"][1][0] += ops.conv5x5_planes_wgrad(gb, acts[1], n_img, H, W, prec, prec) * inv     # block 2: h2 -> h3
                wg["conv"][1][1] += ops.planes_colsum(gb, n_img, H, W, prec) * inv
            ops.conv5x5_tc_dgrad(gb, self._tc_weight_flipped(blocks[2]), acts[1], n_img, H, W, prec, act_out=ga)
            if wg is not None:
                wg["conv"][0][0] += ops.conv5x5_planes_wgrad(ga, acts[0], n_img, H, W, prec, prec) * inv     # block 1: h1 -> h2
                wg["conv"][0][1] += ops.planes_colsum(ga, n_img, H, W, prec) * inv
            tokens = torch.empty(n_img, H * W, 64, device=recons.device, dtype=torch.float32)
            ops.conv5x5_tc_dgrad(ga, self._tc_weight_flipped(blocks[1]), acts[0], n_img, H, W, prec, out_tokens=tokens)
            dG = ops.dec_expand_bwd(tokens, n_img, H, W, k, g_scale)
            out[m0:m1] = ops.linear(dG, wsum_t, mode="fp32").view(m1 - m0, S, self.slot_dim)
            if wg is not None:
                # collapsed first convolution: G = slots @ wsum^T (+ P broadcast): d wsum = dG^T slots, d P = sum over images of d h1
                dws, _ = ops.linear_wgrad(dG, f32c(slots[m0:m1].reshape(n_img, self.slot_dim)), want_bias=False)
                wg["wsum"] += dws
                wg["P_hwc"] += ops.colsum(tokens.view(n_img, H * W * 64)).view(H * W, 64) * inv
        for m0, m1, acts, h in saved:                      # the activations are consumed: hand their buffers back to the pool
            extra = [h[1]] if isinstance(h, tuple) else []
            self._release_train_ws(list(acts) + extra, (m1 - m0) * S, H, W, recons.device)
        if wg is None:
            return out
        params = []
        for dw, db in wg["conv"]:
            params += [dw, db]
        params += wg["final"]
        return out, {"wsum": wg["wsum"], "P_hwc": wg["P_hwc"], "params": params}

    def _final_packed(self, conv):
        if not hasattr(self, "_fin_cache"):
            self._fin_cache = DerivedWeights()
        return self._fin_cache.get([conv.weight], lambda: ops.pack_conv_weight(conv.weight))

    def _final_packed_tc(self, conv):
        if not hasattr(self, "_fin_tc_cache"):
            self._fin_tc_cache = DerivedWeights()
        return self._fin_tc_cache.get([conv.weight], lambda: ops.pack_final_conv_tc(conv.weight, owner=conv.weight))

    def broadcast(self, slots):
        """ (B, S, D) -> (B*S, D, H, W): slot broadcast + position embedding (SAVi.py:237-248).  The decode
        kernels never materialise this tensor; the method exists for API parity (plain tensor ops). """
        D = slots.shape[-1]
        H, W = self.decoder_resolution
        grid = slots.reshape(-1, D, 1, 1).expand(-1, D, H, W)
        return grid + self.decoder_pos_embedding.table(channels_last=False)[None]

    @torch.no_grad()
    def _init_model(self):
        """ Xavier weights / zero biases, zero GRU biases, orthogonal recurrent GRU weights (SAVi.py:251-267) """
        init_xavier_(self)
        torch.nn.init.zeros_(self.slot_attention.gru.bias_ih)
        torch.nn.init.zeros_(self.slot_attention.gru.bias_hh)
        torch.nn.init.orthogonal_(self.slot_attention.gru.weight_hh)
        if hasattr(self.slot_attention, "slots_mu"):
            limit = math.sqrt(6.0 / (1 + self.slot_attention.dim_slots))
            torch.nn.init.uniform_(self.slot_attention.slots_mu, -limit, limit)
            torch.nn.init.uniform_(self.slot_attention.slots_sigma, -limit, limit)
