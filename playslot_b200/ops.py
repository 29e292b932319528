"""
Thin tensor-level wrappers around the C ABI (include/playslot_b200.h).

PyTorch is used for device memory, streams and (later) autograd plumbing only: every function here checks
its arguments, allocates the outputs with torch.empty and launches hand-written kernels on the current
CUDA stream.  CPU tensors are rejected -- there is no fallback.
"""
import math
import os

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_ABS = 0, 1, 2, 3

# arithmetic of the Linear layers: "fp16x3" (tcgen05, split fp16 operands, three products, fp32-grade),
# "fp16" / "bf16" (tcgen05, one half-precision product) or "fp32" (exact FFMA kernel on the CUDA cores)
LINEAR_MODES = ("fp16x3", "fp16", "bf16", "fp32")
_LINEAR_MODE = os.environ.get("PLAYSLOT_LINEAR_MODE", "fp16x3")
LINEAR_TC_MIN_ROWS = 128          # below one row tile the CUDA-core kernel is used


def set_linear_mode(mode):
    global _LINEAR_MODE
    if mode not in LINEAR_MODES:
        raise ValueError(f"linear mode {mode!r}; expected one of {LINEAR_MODES}")
    _LINEAR_MODE = mode


def linear_mode():
    return _LINEAR_MODE


# Weight generation: derived-weight caches (packed tensor-core copies, concatenated matrices) are keyed on the parameters'
# (data_ptr, _version) AND on this counter.  A CUDA-graph replay of a training step updates parameters on the device without
# touching the Python-side `_version`, so train.GraphedTrainStep bumps the counter after every replay.
_WEIGHT_GEN = 0


def weight_generation():
    return _WEIGHT_GEN


def bump_weight_generation():
    global _WEIGHT_GEN
    _WEIGHT_GEN += 1


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise _lib.PlaySlotKernelError(
            f"{name} lives on {t.device}: playslot_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


# --- optional per-kernel device timing (CUDA events on the launching stream); used by bench.py only --------
_PROFILE = None


def profile_start():
    """ start collecting (name, start_event, end_event, meta) for every launch made through this module """
    global _PROFILE
    _PROFILE = []


def profile_stop():
    """ stop collecting; returns {name: {"ms": total, "calls": n, "meta": [..]}} (synchronises the device) """
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, meta in rec or []:
        d = out.setdefault(name, {"ms": 0.0, "calls": 0, "meta": []})
        d["ms"] += e0.elapsed_time(e1)
        d["calls"] += 1
        if meta is not None:
            d["meta"].append(meta)
    return out


class _timed:
    def __init__(self, name, meta=None):
        self.name, self.meta = name, meta

    def __enter__(self):
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.e1.record()
            _PROFILE.append((self.name, self.e0, self.e1, self.meta))
        return False


def launch_count():
    return int(_lib.load().ps_launch_count())


def reset_launch_count():
    _lib.load().ps_reset_launch_count()


class sm_reserve:
    """ context manager: the persistent kernels launched inside leave `n` SMs free for the kernels of a concurrent stream """

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        lib = _lib.load()
        self.old = int(lib.ps_get_sm_reserve())
        _lib.check(lib.ps_set_sm_reserve(self.n), "ps_set_sm_reserve")
        return self

    def __exit__(self, *exc):
        _lib.load().ps_set_sm_reserve(self.old)
        return False


# ------------------------------------------------------------------------------------------------ dense

class Planes:
    """ an activation matrix (M, N) stored as fp16 [hi | lo] planes (int16 tensor (2, M, N), lo scaled by 2048): the operand
    format the tensor-core linear loads by TMA.  Produced by linear(..., out_planes=True) / mha(..., out_planes=True) and
    consumed by linear(Planes, ...), so chained linears skip the fp32 round trip and the conversion work. """
    __slots__ = ("data",)

    def __init__(self, data):
        self.data = data

    @property
    def shape(self):
        return self.data.shape[1:]


def planes_ok(rows, features=32):
    """ can a (rows, features) activation be handed over as Planes?  (fp16x3 tensor-core linears only) """
    return _LINEAR_MODE == "fp16x3" and rows >= LINEAR_TC_MIN_ROWS and features % 32 == 0


def linear(x, weight, bias=None, ln=None, act=ACT_NONE, res=None, res_div=1, res_mod=None, out=None, mode=None, out_planes=False,
           reuse_pack_scale=False):
    """
    y = act(LN(x) @ weight.T + bias) + res_rows.   x: (..., K) contiguous; weight (N, K).
    ln = (gamma, beta, eps) fuses a LayerNorm over K in front.  `res` is a 2-D (R, N) tensor; output
    row m receives res[(m // res_div) % res_mod] (defaults: plain residual of the same shape).
    `mode` overrides the global linear mode for this call (the training path passes "fp32": its weights change
    every optimiser step, so a packed tensor-core copy would be rebuilt per call).
    """
    lib = _lib.load()
    if isinstance(x, Planes) or out_planes:
        return _linear_planes(x, weight, bias, ln, act, res, res_div, res_mod, out_planes)
    _req(x, "x"); _req(weight, "weight")
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f"linear: weight {tuple(weight.shape)} does not match input features {K}")
    M = x.numel() // K
    x2 = x.view(M, K)
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
    y2 = out.view(M, N)
    g = b = None
    eps = 0.0
    if ln is not None:
        g, b, eps = ln
        _req(g, "ln gamma"); _req(b, "ln beta")
    if bias is not None:
        _req(bias, "bias")
    ldres = 0
    if res is not None:
        _req(res, "res")
        res = res.view(-1, N)
        ldres = N
        if res_mod is None:
            res_mod = res.shape[0]
    lmode = mode or _LINEAR_MODE
    if lmode != "fp32" and M >= LINEAR_TC_MIN_ROWS:
        prec = TC_PRECISION[lmode]
        w_tc, w_scale, Kp, Npad = _packed_linear_weight(weight, prec, reuse_scale=reuse_pack_scale or REUSE_CONV_PACK_SCALE)
        with _timed("linear_tc_p%d" % prec, (M, N, K)):
            st = lib.ps_linear_tc(_ptr(x2), K, _ptr(w_tc), float(w_scale), Kp, Npad, _ptr(bias), _ptr(g), _ptr(b), float(eps),
                                  _ptr(res), ldres, int(res_div), int(res_mod or 0), _ptr(y2), N, M, N, K, int(act), prec, _stream())
        _lib.check(st, "ps_linear_tc")
        return out
    with _timed("linear", (M, N, K)):
        st = lib.ps_linear(_ptr(x2), K, _ptr(weight), _ptr(bias), _ptr(g), _ptr(b), float(eps),
                           _ptr(res), ldres, int(res_div), int(res_mod or 0), _ptr(y2), N, M, N, K, int(act), _stream())
    _lib.check(st, "ps_linear")
    return out


def pack_linear_weight_tc(w, precision, scale=None):
    """
    nn.Linear weight (N, K) fp32 -> (16-bit tensor (2 [hi, lo], Npad, Kp), w_scale, Kp, Npad): K padded to a
    multiple of 64 and N to a multiple of 128 with zeros; weights multiplied by a power of two that puts max|w|
    near 1 (exact; undone in the kernel epilogue) so the fp16 parts stay away from the subnormal range.
    `scale`: reuse a previously chosen power of two instead of measuring max|w| (no host synchronisation: the form the
    training path uses to repack weights that change every step, also inside a CUDA-graph capture).
    """
    N, K = w.shape
    Kp, Npad = (K + 63) // 64 * 64, (N + 127) // 128 * 128
    wt = torch.zeros(Npad, Kp, device=w.device, dtype=torch.float32)
    wt[:N, :K] = w.detach().float()
    if precision == 1:
        scale = 1.0                                   # bf16 has the exponent range of fp32: no scaling (and no host sync)
    if scale is None:
        amax = float(wt.abs().max())
        scale = 1.0 if amax == 0.0 else 2.0 ** (-round(math.log2(amax)))
    wt = wt * scale
    if precision == 1:
        hi = wt.to(torch.bfloat16).view(torch.int16)
        lo = torch.zeros_like(hi)
    else:
        h = wt.to(torch.float16)
        hi = h.view(torch.int16)
        lo = ((wt - h.float()) * TC_LO_SCALE).to(torch.float16).view(torch.int16)
    return torch.stack([hi, lo], dim=0).contiguous(), scale, Kp, Npad


def _packed_linear_weight(weight, precision, transposed=False, reuse_scale=False):
    """ packed copy (of weight, or of weight.t() for the data-gradient GEMM) cached on the weight tensor object, rebuilt when
    the tensor is modified in place or moved; the power-of-two scale of the first packing is kept for the rebuilds """
    attr = "_ps_tc_pack_t" if transposed else "_ps_tc_pack"
    key = (weight._version, weight.data_ptr(), precision, _WEIGHT_GEN)
    c = getattr(weight, attr, None)
    if c is None or c[0] != key:
        with torch.no_grad():
            scale = getattr(weight, "_ps_tc_scale", None) if reuse_scale else None      # training: sync-free rebuilds
            if scale is None and reuse_scale and precision != 1:
                # no scale yet and no host synchronisation allowed (temporary weights such as a concatenated q/k/v matrix are
                # new tensor objects every call): the power of two that normalises a Xavier-sized weight of this shape.  The
                # scale only keeps the fp16 parts in range (6e-5 < |w| * scale < 65504), a few octaves off do not matter.
                scale = 2.0 ** (-round(math.log2(math.sqrt(6.0 / (weight.shape[0] + weight.shape[1])))))
            src = weight.detach().t() if transposed else weight
            c = (key, pack_linear_weight_tc(src, precision, scale=scale))
            if precision != 1:
                weight._ps_tc_scale = c[1][1]
        setattr(weight, attr, c)
    return c[1]


def linear_dgrad_tc(dy, weight, mode="fp16x3"):
    """ dX = dY @ W on the tensor cores (fp16x3, or one bf16 product with mode="bf16"): dy (M, N), weight (N, K) -> (M, K);
    packs W^T like a Linear weight """
    lib = _lib.load()
    _req(dy, "dy"); _req(weight, "weight")
    N, K = weight.shape
    M = dy.numel() // N
    prec = TC_PRECISION[mode]
    w_tc, w_scale, Kp, Npad = _packed_linear_weight(weight, prec, transposed=True, reuse_scale=True)   # a (K, N) "weight"
    out = torch.empty(M, K, device=dy.device, dtype=torch.float32)
    with _timed("linear_tc_p%d" % prec, (M, K, N)):
        st = lib.ps_linear_tc(_ptr(dy.view(M, N)), N, _ptr(w_tc), float(w_scale), Kp, Npad, None, None, None, 0.0,
                              None, 0, 1, 0, _ptr(out), K, M, K, N, 0, prec, _stream())
    _lib.check(st, "ps_linear_tc")
    return out


def _linear_planes(x, weight, bias, ln, act, res, res_div, res_mod, out_planes):
    """ linear() with Planes on the input and / or output side (fp16x3 tensor-core kernel only) """
    lib = _lib.load()
    _req(weight, "weight")
    N, K = weight.shape
    xin = x.data if isinstance(x, Planes) else x
    if isinstance(x, Planes):
        if not xin.is_cuda or xin.dtype != torch.int16 or not xin.is_contiguous():
            raise TypeError("linear: Planes operand must be a contiguous CUDA int16 tensor (2, M, K)")
    else:
        _req(xin, "x")
    if xin.shape[-1] != K:
        raise ValueError(f"linear: weight {tuple(weight.shape)} does not match input features {xin.shape[-1]}")
    M = xin.numel() // K // (2 if isinstance(x, Planes) else 1)
    if M < LINEAR_TC_MIN_ROWS:
        raise ValueError("linear: plane operands need at least %d rows" % LINEAR_TC_MIN_ROWS)
    g = b = None
    eps = 0.0
    if ln is not None:
        g, b, eps = ln
    ldres = 0
    if res is not None:
        _req(res, "res")
        res = res.view(-1, N)
        ldres = N
        if res_mod is None:
            res_mod = res.shape[0]
    w_tc, w_scale, Kp, Npad = _packed_linear_weight(weight, 3)
    if out_planes:
        y = torch.empty(2, M, N, device=xin.device, dtype=torch.int16)
        out = Planes(y)
    else:
        y = torch.empty(M, N, device=xin.device, dtype=torch.float32)
        out = y
    is_p = isinstance(x, Planes)
    with _timed("linear_tc_p3", (M, N, K)):
        st = lib.ps_linear_tc_planes(None if is_p else _ptr(xin), K, _ptr(xin) if is_p else None, M if is_p else 0, _ptr(w_tc),
                                     float(w_scale), Kp, Npad, _ptr(bias), _ptr(g), _ptr(b), float(eps), _ptr(res), ldres,
                                     int(res_div), int(res_mod or 0), None if out_planes else _ptr(y), _ptr(y) if out_planes else None,
                                     M if out_planes else 0, N, M, N, K, int(act), _stream())
    _lib.check(st, "ps_linear_tc_planes")
    return out


def mha(qkv, n_seq, L, n_heads, head_dim, slots_per_frame=0, res=None, out_planes=False):
    """ qkv: (n_seq*L, 3*n_heads*head_dim) -> (n_seq*L, n_heads*head_dim) (+ res of the same shape); out_planes: as Planes """
    lib = _lib.load()
    _req(qkv, "qkv")
    inner = n_heads * head_dim
    qkv2 = qkv.view(n_seq * L, 3 * inner)
    if out_planes:
        y = torch.empty(2, n_seq * L, inner, device=qkv.device, dtype=torch.int16)
        with _timed("mha"):
            st = lib.ps_mha_planes(_ptr(qkv2), 3 * inner, _ptr(res), _ptr(y), n_seq * L, inner, n_seq, L, n_heads, head_dim,
                                   int(slots_per_frame), _stream())
        _lib.check(st, "ps_mha_planes")
        return Planes(y)
    out = torch.empty(n_seq * L, inner, device=qkv.device, dtype=torch.float32)
    if res is not None:
        _req(res, "res")
        assert res.numel() == out.numel()
    with _timed("mha"):
        st = lib.ps_mha(_ptr(qkv2), 3 * inner, _ptr(res), _ptr(out), inner, n_seq, L, n_heads, head_dim, int(slots_per_frame), _stream())
    _lib.check(st, "ps_mha")
    return out


def pack_conv_weight(w):
    """ nn.Conv2d weight (Cout, Cin, K, K) -> packed (Cin, K*K, Cout) used by ps_conv2d / ps_dec_composite """
    Cout, Cin, K, _ = w.shape
    return w.detach().permute(1, 2, 3, 0).reshape(Cin, K * K, Cout).contiguous()


def conv2d(x, w_packed, bias, ksize, relu, out_nhwc=False, addend=None):
    """ x (n, Cin, H, W) -> (n, Cout, H, W) or, with out_nhwc, (n, H, W, Cout) (+ addend (H*W, Cout)) """
    lib = _lib.load()
    _req(x, "x"); _req(w_packed, "w_packed")
    n, Cin, H, W = x.shape
    Cout = w_packed.shape[-1]
    assert w_packed.shape[0] == Cin and w_packed.shape[1] == ksize * ksize
    shape = (n, H, W, Cout) if out_nhwc else (n, Cout, H, W)
    y = torch.empty(shape, device=x.device, dtype=torch.float32)
    MAXI = 65535
    for i0 in range(0, n, MAXI):
        i1 = min(n, i0 + MAXI)
        with _timed("conv2d_%dx%d_%dto%d" % (ksize, ksize, Cin, Cout), (i1 - i0, Cin, Cout, H, W, ksize)):
            st = lib.ps_conv2d(_ptr(x[i0:i1]), _ptr(w_packed), _ptr(bias), _ptr(addend), _ptr(y[i0:i1]),
                               i1 - i0, Cin, Cout, H, W, ksize, int(relu), int(out_nhwc), _stream())
        _lib.check(st, "ps_conv2d")
    return y


def conv2d_split(x, w_packed, bias, ksize, relu, act_out, precision):
    """ x (n, Cin, H, W) fp32 -- or uint8 images 0..255, normalised by 1/255 inside the kernel -- -> padded NHWC split planes
    (Cout == 64) in `act_out` """
    lib = _lib.load()
    _req(w_packed, "w_packed")
    n, Cin, H, W = x.shape
    assert w_packed.shape[-1] == 64 and n <= 65535
    if x.dtype == torch.uint8:
        if not x.is_cuda or not x.is_contiguous():
            raise ValueError("conv2d_split: uint8 images must be a contiguous CUDA tensor")
        with _timed("conv2d_%dx%d_%dto64_split" % (ksize, ksize, Cin), (n, Cin, 64, H, W, ksize)):
            st = lib.ps_conv2d_split_u8(_ptr(x), _ptr(w_packed), _ptr(bias), _ptr(act_out), n, Cin, H, W, ksize, int(relu), int(precision), _stream())
        _lib.check(st, "ps_conv2d_split_u8")
        return act_out
    _req(x, "x")
    with _timed("conv2d_%dx%d_%dto64_split" % (ksize, ksize, Cin), (n, Cin, 64, H, W, ksize)):
        st = lib.ps_conv2d_split(_ptr(x), _ptr(w_packed), _ptr(bias), _ptr(act_out), n, Cin, H, W, ksize, int(relu), int(precision), _stream())
    _lib.check(st, "ps_conv2d_split")
    return act_out


# ------------------------------------------------------------------------------------------------ SlotAttention

def sa_partial_floats(S):
    return int(_lib.load().ps_sa_partial_floats(S))


def sa_main(x, A, a0, a1, partial, masks, n_chunks, ln_eps, attn_eps):
    lib = _lib.load()
    _req(x, "feats"); _req(A, "A"); _req(partial, "partial")
    B, N, D = x.shape
    S = A.shape[1]
    with _timed("sa_main", (B, N, S, D, masks is not None)):
        st = lib.ps_sa_main(_ptr(x), _ptr(A), _ptr(a0), _ptr(a1), _ptr(partial), _ptr(masks), B, N, S, D, n_chunks,
                            float(ln_eps), float(attn_eps), _stream())
    _lib.check(st, "ps_sa_main")


def sa_main_tc(x, A, a0, partial, masks, n_chunks, ln_eps, attn_eps):
    lib = _lib.load()
    _req(x, "feats"); _req(A, "A"); _req(partial, "partial")
    B, N, D = x.shape
    S = A.shape[1]
    with _timed("sa_main_tc", (B, N, S, D, masks is not None)):
        st = lib.ps_sa_main_tc(_ptr(x), _ptr(A), _ptr(a0), _ptr(partial), _ptr(masks), B, N, S, D, n_chunks,
                               float(ln_eps), float(attn_eps), _stream())
    _lib.check(st, "ps_sa_main_tc")


def sa_prepare(x, ln_eps=1e-5):
    """ x (..., N, 128) fp32 features of one or several frame batches -> planes int16 (2 [hi, lo], n_tokens, 128): one pass
    applying norm_input's statistics (no affine part) and the fp16 hi/lo split """
    lib = _lib.load()
    _req(x, "feats")
    D = x.shape[-1]
    n_tok = x.numel() // D
    planes = torch.empty(2, n_tok, D, device=x.device, dtype=torch.int16)
    with _timed("sa_prepare", (n_tok,)):
        st = lib.ps_sa_prepare(_ptr(x), _ptr(planes), n_tok, D, float(ln_eps), _stream())
    _lib.check(st, "ps_sa_prepare")
    return planes


def enc_mlp_planes(tokens, K1, ln, fc1, fc2, out_eps, pair_hw=0):
    """ tokens (n_tok, ldx) fp32 (first K1 columns used), ln = (gamma, beta, eps) over K1, fc1 = (W1 (128,K1), b1),
    fc2 = (W2 (128,128), b2) -> planes int16 (2 [hi, lo], n_tok, 128) of norm_input-normalised encoder features: the
    encoder MLP and sa_prepare in one tensor-core kernel.  pair_hw > 0: the token rows are (pair, position, image-in-pair)
    (channel-paired 32-channel encoder); the planes come out in image order """
    lib = _lib.load()
    _req(tokens, "tokens")
    n_tok, ldx = tokens.shape
    w1, s1, kp1, np1 = _packed_linear_weight(fc1[0], 3)
    w2, s2, kp2, np2 = _packed_linear_weight(fc2[0], 3)
    assert (kp1, np1, kp2, np2) == (64, 128, 128, 128), "enc_mlp_planes: Linear(K1<=64, 128) -> Linear(128, 128) only"
    planes = torch.empty(2, n_tok, 128, device=tokens.device, dtype=torch.int16)
    with _timed("enc_mlp_planes", (n_tok, K1)):
        st = lib.ps_enc_mlp_planes_paired(_ptr(tokens), ldx, K1, _ptr(ln[0]), _ptr(ln[1]), float(ln[2]), _ptr(w1), float(s1), _ptr(fc1[1]),
                                          _ptr(w2), float(s2), _ptr(fc2[1]), float(out_eps), _ptr(planes), n_tok, int(pair_hw), _stream())
    _lib.check(st, "ps_enc_mlp_planes_paired")
    return planes


def sa_tc2_chunks(B, N):
    """ partial slots per frame sa_main_tc2 needs for a (B, N) batch """
    return int(_lib.load().ps_sa_tc2_chunks(int(B), int(N)))


def sa_main_tc2(planes, first_token, B, N, A, a0, partial, masks, n_chunks, attn_eps):
    """ planes: (2, n_tokens, 128) int16 from sa_prepare; the batch covers tokens [first_token, first_token + B*N) """
    lib = _lib.load()
    S = A.shape[1]
    n_tok = planes.shape[1]
    base = planes.data_ptr() + first_token * 128 * 2
    with _timed("sa_main_tc2", (B, N, S, 128, masks is not None)):
        st = lib.ps_sa_main_tc2(base, n_tok, _ptr(A), _ptr(a0), _ptr(partial), _ptr(masks), B, N, S, 128, n_chunks,
                                float(attn_eps), _stream())
    _lib.check(st, "ps_sa_main_tc2")


def sa_update(partial, n_chunks, slots_in, slots_out, w, A, a0, a1, mlp_hidden, scale, ln_eps, do_update, do_prep):
    """ w: dict of prepared SlotAttention weights (see models.BlocksUtils.attention.SlotAttention._weights) """
    lib = _lib.load()
    B, S, D = slots_in.shape
    with _timed("sa_update"):
      st = lib.ps_sa_update(_ptr(partial), n_chunks, _ptr(slots_in), _ptr(slots_out),
                          _ptr(w["ln_in_g"]), _ptr(w["ln_in_b"]), _ptr(w["WvT"]),
                          _ptr(w["WihT"]), _ptr(w["WhhT"]), _ptr(w["b_ih"]), _ptr(w["b_hh"]),
                          _ptr(w["ln_mlp_g"]), _ptr(w["ln_mlp_b"]),
                          _ptr(w["W1T"]), _ptr(w["b1"]), _ptr(w["W2T"]), _ptr(w["b2"]),
                          _ptr(w["ln_slot_g"]), _ptr(w["ln_slot_b"]), _ptr(w["WqT"]), _ptr(w["Wk"]),
                          _ptr(A), _ptr(a0), _ptr(a1), B, S, D, int(mlp_hidden), float(scale), float(ln_eps),
                          int(do_update), int(do_prep), _stream())
    _lib.check(st, "ps_sa_update")


def sa_update2(partial, n_chunks, slots_in, slots_out, w, A, a0, mlp_hidden, ln_eps, do_update, do_prep):
    """ latency-optimised slot update with the folded weights w["WivT"], w["MAT"], w["v0"] """
    lib = _lib.load()
    B, S, D = slots_in.shape
    with _timed("sa_update"):
        st = lib.ps_sa_update2(_ptr(partial), n_chunks, _ptr(slots_in), _ptr(slots_out),
                               _ptr(w["ln_in_g"]), _ptr(w["ln_in_b"]),
                               _ptr(w["WivT"]), _ptr(w["WhhT"]), _ptr(w["b_ih"]), _ptr(w["b_hh"]),
                               _ptr(w["ln_mlp_g"]), _ptr(w["ln_mlp_b"]),
                               _ptr(w["W1T"]), _ptr(w["b1"]), _ptr(w["W2T"]), _ptr(w["b2"]),
                               _ptr(w["ln_slot_g"]), _ptr(w["ln_slot_b"]), _ptr(w["MAT"]), _ptr(w["v0"]),
                               _ptr(A), _ptr(a0), B, S, D, int(mlp_hidden), float(ln_eps), int(do_update), int(do_prep), _stream())
    _lib.check(st, "ps_sa_update2")


# ------------------------------------------------------------------------------------------------ decoder

def dec_expand(G, P, n_img, C, H, W, ksize):
    lib = _lib.load()
    _req(G, "G"); _req(P, "P")
    out = torch.empty(n_img, C, H, W, device=G.device, dtype=torch.float32)
    with _timed("dec_expand"):
        st = lib.ps_dec_expand(_ptr(G), _ptr(P), _ptr(out), n_img, C, H, W, ksize, _stream())
    _lib.check(st, "ps_dec_expand")
    return out


def dec_composite(h, w_packed, bias, M, S):
    lib = _lib.load()
    _req(h, "h"); _req(w_packed, "w_packed"); _req(bias, "bias")
    _, Cin, H, W = h.shape
    recon = torch.empty(M, 3, H, W, device=h.device, dtype=torch.float32)
    recons = torch.empty(M, S, 3, H, W, device=h.device, dtype=torch.float32)
    masks = torch.empty(M, S, 1, H, W, device=h.device, dtype=torch.float32)
    MAXM = 65535
    for m0 in range(0, M, MAXM):
        m1 = min(M, m0 + MAXM)
        with _timed("dec_composite"):
            st = lib.ps_dec_composite(_ptr(h[m0 * S:m1 * S]), _ptr(w_packed), _ptr(bias), _ptr(recon[m0:m1]),
                                      _ptr(recons[m0:m1]), _ptr(masks[m0:m1]), m1 - m0, S, Cin, H, W, _stream())
        _lib.check(st, "ps_dec_composite")
    return recon, recons, masks


def pack_final_conv_tc(w, owner=None):
    """ final decoder conv weight (4, 64, 3, 3) -> (int16 (3 filter rows, 32 rows [hi: kx0 o0-3 | kx1 o0-3 | kx2 o0-3 | 4 zero; lo: same],
    64 cin), w_scale): the three taps of a filter row form one N = 32 operand of ps_dec_composite_tc """
    Cout, Cin, K, _ = w.shape
    assert (Cout, Cin, K) == (4, 64, 3), "the tensor-core compositing kernel covers the 64 -> 4 channel 3x3 convolution"
    wt = w.detach().float().permute(2, 3, 0, 1).reshape(3, 12, 64)          # [ky][kx * 4 + o][cin]
    scale = _conv_pack_scale(wt, owner, 3)
    wt = wt * scale
    h = wt.to(torch.float16)
    lo = ((wt - h.float()) * TC_LO_SCALE).to(torch.float16)
    out = torch.zeros(3, 32, 64, device=w.device, dtype=torch.float16)
    out[:, 0:12] = h
    out[:, 16:28] = lo
    return out.view(torch.int16).contiguous(), scale


def dec_composite_tc(act, w_packed, bias, M, S, H, W):
    """ act: padded split planes of the last 5x5 block's output for M*S slot images -> (recon, recons, masks) """
    lib = _lib.load()
    w_tc, w_scale = w_packed
    _req(bias, "bias")
    dev = act.device
    recon = torch.empty(M, 3, H, W, device=dev, dtype=torch.float32)
    recons = torch.empty(M, S, 3, H, W, device=dev, dtype=torch.float32)
    masks = torch.empty(M, S, 1, H, W, device=dev, dtype=torch.float32)
    with _timed("dec_composite_tc"):
        st = lib.ps_dec_composite_tc(_ptr(act), _ptr(w_tc), float(w_scale), _ptr(bias), _ptr(recon), _ptr(recons), _ptr(masks),
                                     M, S, H, W, _stream())
    _lib.check(st, "ps_dec_composite_tc")
    return recon, recons, masks


# ------------------------------------------------------------------------------------------------ InvDyn tail

def invdyn_tail(mv, noise, codebook, per_slot, eps=1e-6):
    """
    mv (B, T, R, 2A), noise (B, T-1, R, A), codebook (K, A) ->
    dict(action_dist (B,T-1,2,R,A), z, proto, variability (B,T-1,R,A), idx int64 (B,T-1,R), loss (B,T-1,R))
    """
    lib = _lib.load()
    _req(mv, "mv"); _req(noise, "noise"); _req(codebook, "codebook")
    B, T, R, A2 = mv.shape
    A = A2 // 2
    K = codebook.shape[0]
    dev = mv.device
    dist = torch.empty(B, T - 1, 2, R, A, device=dev, dtype=torch.float32)
    z = torch.empty(B, T - 1, R, A, device=dev, dtype=torch.float32)
    proto = torch.empty_like(z)
    var = torch.empty_like(z)
    idx = torch.empty(B, T - 1, R, device=dev, dtype=torch.int64)
    loss = torch.empty(B, T - 1, R, device=dev, dtype=torch.float32)
    st = lib.ps_invdyn_tail(_ptr(mv), _ptr(noise), _ptr(codebook), _ptr(dist), _ptr(z), _ptr(proto), _ptr(var),
                            _ptr(idx), _ptr(loss), B, T, R, A, K, int(per_slot), float(eps), _stream())
    _lib.check(st, "ps_invdyn_tail")
    return {"action_dist": dist, "z": z, "proto": proto, "variability": var, "idx": idx, "loss": loss}


def vq_lookup(z, codebook):
    """ z (..., A) -> (proto = z + (z_q - z), z_q, idx int64 (..., 1), loss (...)) """
    lib = _lib.load()
    _req(z, "z"); _req(codebook, "codebook")
    A = z.shape[-1]
    n = z.numel() // A
    K = codebook.shape[0]
    proto = torch.empty_like(z)
    zq = torch.empty_like(z)
    idx = torch.empty(*z.shape[:-1], 1, device=z.device, dtype=torch.int64)
    loss = torch.empty(z.shape[:-1], device=z.device, dtype=torch.float32)
    st = lib.ps_vq_lookup(_ptr(z), _ptr(codebook), _ptr(proto), _ptr(zq), _ptr(idx), _ptr(loss), n, A, K, _stream())
    _lib.check(st, "ps_vq_lookup")
    return proto, zq, idx, loss


# ------------------------------------------------------------------------------------------------ tcgen05 convolution

TC_PRECISION = {"fp16x3": 3, "fp16": 2, "bf16": 1}
TC_LO_SCALE = 2048.0


def conv_tc_workspace(n_img, H, W, device):
    """ zero-initialised activation buffer in the padded NHWC bf16 [hi | lo] format of ps_conv5x5_tc """
    elems = int(_lib.load().ps_conv_tc_plane_elems(n_img, H, W))
    return torch.zeros(2, elems // 64, 64, device=device, dtype=torch.int16)


# Training steps captured in a CUDA graph repack the convolution weights inside the graph every replay: the power-of-two scale
# (which only keeps the fp16 parts in range) must then come without a host synchronisation.  While this flag is set, the packers
# reuse the scale chosen at a tensor's first packing (stored on the parameter object) instead of measuring max|w| again (the
# tensor-core Linear layers do the same through _packed_linear_weight(reuse_scale=True)).
REUSE_CONV_PACK_SCALE = False


def _conv_pack_scale(wt, owner, precision):
    """ power of two that puts max|w| near 1; `owner`: the parameter the scale is remembered on (None: always measured) """
    if precision == 1:
        return 1.0
    if owner is not None and REUSE_CONV_PACK_SCALE:
        sc = getattr(owner, "_ps_conv_scale", None)
        if sc is not None:
            return sc
    amax = float(wt.abs().max())
    scale = 1.0 if amax == 0.0 else 2.0 ** (-round(math.log2(amax)))
    if owner is not None:
        owner._ps_conv_scale = scale
    return scale


def pack_conv_weight_tc(w, precision, owner=None):
    """
    nn.Conv2d weight (64, 64, 5, 5) fp32 -> (16-bit tensor (25 taps, 2 [hi, lo], 64 cout, 64 cin), w_scale).
    The weights are multiplied by a power of two w_scale that puts max|w| near 1 (exact; undone in the kernel's
    epilogue) so that the fp16 parts stay far away from the subnormal range.
    """
    Cout, Cin, K, _ = w.shape
    assert (Cout, Cin, K) == (64, 64, 5), "the tensor-core convolution covers 64 -> 64 channels, 5x5"
    wt = w.detach().float().permute(2, 3, 0, 1).reshape(K * K, Cout, Cin)
    scale = _conv_pack_scale(wt, owner, precision)
    wt = wt * scale
    if precision == 1:
        hi = wt.to(torch.bfloat16).view(torch.int16)
        lo = torch.zeros_like(hi)
    else:
        h = wt.to(torch.float16)
        hi = h.view(torch.int16)
        lo = ((wt - h.float()) * TC_LO_SCALE).to(torch.float16).view(torch.int16)
    return torch.stack([hi, lo], dim=1).contiguous(), scale


def conv5x5_tc(act_in, w_packed, bias, n_img, H, W, relu, precision, act_out=None, out_nchw=None, out_tokens=None, addend=None):
    """ w_packed = pack_conv_weight_tc(w, precision); exactly one of act_out (padded split), out_nchw (fp32
    [n,64,H,W]) or out_tokens (fp32 [n,H*W,64]) receives the result """
    lib = _lib.load()
    w_tc, w_scale = w_packed
    mode = 0 if act_out is not None else (1 if out_nchw is not None else 2)
    f32 = out_nchw if mode == 1 else out_tokens
    with _timed("conv5x5_tc_p%d" % precision, (n_img, 64, 64, H, W, 5)):
        st = lib.ps_conv5x5_tc(_ptr(act_in), _ptr(w_tc), _ptr(bias), float(w_scale), _ptr(act_out), _ptr(f32), _ptr(addend),
                               n_img, H, W, int(relu), mode, int(precision), _stream())
    _lib.check(st, "ps_conv5x5_tc")
    return act_out if mode == 0 else f32


def dec_expand_split(G, P_hwc, act_out, n_img, H, W, ksize, precision):
    lib = _lib.load()
    _req(G, "G"); _req(P_hwc, "P_hwc")
    with _timed("dec_expand_split"):
        st = lib.ps_dec_expand_split(_ptr(G), _ptr(P_hwc), _ptr(act_out), n_img, H, W, ksize, int(precision), _stream())
    _lib.check(st, "ps_dec_expand_split")
    return act_out


def nchw_to_split(x, act_out, precision):
    lib = _lib.load()
    _req(x, "x")
    n, C, H, W = x.shape
    assert C == 64
    with _timed("nchw_to_split"):
        for i0 in range(0, n, 65535):
            i1 = min(n, i0 + 65535)
            assert i0 == 0, "more than 65535 images per call are not supported"
            st = lib.ps_nchw_to_split(_ptr(x), _ptr(act_out), n, H, W, int(precision), _stream())
            _lib.check(st, "ps_nchw_to_split")
    return act_out


# ------------------------------------------------------------------------------------------------ training kernels

def layernorm_fwd(x, gamma, beta, eps):
    """ x (..., K) -> (y, mean (M,), rstd (M,)): un-fused LayerNorm of the training path (keeps the statistics) """
    lib = _lib.load()
    _req(x, "x"); _req(gamma, "gamma"); _req(beta, "beta")
    K = x.shape[-1]
    M = x.numel() // K
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    with _timed("layernorm_fwd"):
        st = lib.ps_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), float(eps), _ptr(y), _ptr(mean), _ptr(rstd), M, K, _stream())
    _lib.check(st, "ps_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma):
    lib = _lib.load()
    _req(dy, "dy"); _req(x, "x")
    K = x.shape[-1]
    M = x.numel() // K
    dx = torch.empty_like(x)
    dg = torch.empty(K, device=x.device, dtype=torch.float32)
    db = torch.empty(K, device=x.device, dtype=torch.float32)
    with _timed("layernorm_bwd"):
        st = lib.ps_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dx), _ptr(dg), _ptr(db), M, K, _stream())
    _lib.check(st, "ps_layernorm_bwd")
    return dx, dg, db


def linear_wgrad(dy, x, want_bias=True):
    """ dy (M, N), x (M, K) -> (dW (N, K), db (N,) or None) """
    lib = _lib.load()
    _req(dy, "dy"); _req(x, "x")
    M, N = dy.shape
    K = x.shape[1]
    dW = torch.empty(N, K, device=x.device, dtype=torch.float32)
    db = torch.empty(N, device=x.device, dtype=torch.float32) if want_bias else None
    with _timed("linear_wgrad", (M, N, K)):
        st = lib.ps_linear_wgrad(_ptr(dy), N, _ptr(x), K, _ptr(dW), _ptr(db), M, N, K, _stream())
    _lib.check(st, "ps_linear_wgrad")
    return dW, db


def linear_wgrad_acc(dy, x, dW, db=None):
    """ dW (N, K) += dy^T x, db (N,) += colsum(dy): accumulate into existing gradient buffers """
    lib = _lib.load()
    _req(dy, "dy"); _req(x, "x"); _req(dW, "dW")
    M, N = dy.shape
    K = x.shape[1]
    assert dW.shape == (N, K) and dW.is_contiguous() and (db is None or (db.shape == (N,) and db.is_contiguous()))
    with _timed("linear_wgrad", (M, N, K)):
        st = lib.ps_linear_wgrad_acc(_ptr(dy), N, _ptr(x), K, _ptr(dW), _ptr(db), M, N, K, _stream())
    _lib.check(st, "ps_linear_wgrad_acc")


def act_fwd(z, act):
    lib = _lib.load()
    _req(z, "z")
    y = torch.empty_like(z)
    with _timed("act_fwd"):
        st = lib.ps_act_fwd(_ptr(z), _ptr(y), z.numel(), int(act), _stream())
    _lib.check(st, "ps_act_fwd")
    return y


def act_bwd(dy, z, act):
    lib = _lib.load()
    _req(dy, "dy"); _req(z, "z")
    dz = torch.empty_like(z)
    with _timed("act_bwd"):
        st = lib.ps_act_bwd(_ptr(dy), _ptr(z), _ptr(dz), z.numel(), int(act), _stream())
    _lib.check(st, "ps_act_bwd")
    return dz


def mha_bwd(qkv, dout, n_seq, L, n_heads, head_dim, slots_per_frame=0):
    lib = _lib.load()
    _req(qkv, "qkv"); _req(dout, "dout")
    inner = n_heads * head_dim
    dqkv = torch.empty_like(qkv)
    with _timed("mha_bwd"):
        st = lib.ps_mha_bwd(_ptr(qkv), 3 * inner, _ptr(dout), inner, _ptr(dqkv), n_seq, L, n_heads, head_dim, int(slots_per_frame), _stream())
    _lib.check(st, "ps_mha_bwd")
    return dqkv


def invdyn_tail_bwd(mv, noise, z, idx, codebook, g_dist, g_z, g_proto, g_var, g_quant, per_slot, eps=1e-6):
    lib = _lib.load()
    B, T, R, A2 = mv.shape
    A = A2 // 2
    d_mv = torch.empty_like(mv)
    c = lambda t: None if t is None else _req(t.contiguous(), "grad")
    gs = [c(g) for g in (g_dist, g_z, g_proto, g_var, g_quant)]
    st = lib.ps_invdyn_tail_bwd(_ptr(mv), _ptr(noise), _ptr(z), _ptr(idx), _ptr(codebook), *[_ptr(g) for g in gs], _ptr(d_mv),
                                B, T, R, A, codebook.shape[0], int(per_slot), float(eps), _stream())
    _lib.check(st, "ps_invdyn_tail_bwd")
    return d_mv


def dec_composite_bwd(recons, masks, g_recon, g_recons, g_masks):
    """ gradients w.r.t. the 4-channel output of the last decoder convolution, (M*S, 4, H, W) """
    lib = _lib.load()
    M, S, _, H, W = recons.shape
    g_o = torch.empty(M * S, 4, H, W, device=recons.device, dtype=torch.float32)
    with _timed("dec_composite_bwd"):
        st = lib.ps_dec_composite_bwd(_ptr(recons), _ptr(masks), _ptr(g_recon), _ptr(g_recons), _ptr(g_masks), _ptr(g_o), M, S, H, W, _stream())
    _lib.check(st, "ps_dec_composite_bwd")
    return g_o


def dec_final_dgrad_planes(g_o, w_packed, h_planes, act_out, n_img, H, W, precision, g_scale=None):
    """ dec_final_dgrad with the saved last-block activation as padded split planes (predictor training: no fp32 NCHW copy) """
    lib = _lib.load()
    with _timed("dec_final_dgrad"):
        st = lib.ps_dec_final_dgrad_planes(_ptr(g_o), _ptr(w_packed), _ptr(h_planes), _ptr(act_out), _ptr(g_scale), n_img, H, W,
                                           int(precision), _stream())
    _lib.check(st, "ps_dec_final_dgrad_planes")
    return act_out


def dec_final_dgrad(g_o, w_packed, h, act_out, precision, g_scale=None):
    lib = _lib.load()
    n_img, _, H, W = h.shape
    with _timed("dec_final_dgrad"):
        st = lib.ps_dec_final_dgrad(_ptr(g_o), _ptr(w_packed), _ptr(h), _ptr(act_out), _ptr(g_scale), n_img, H, W, int(precision), _stream())
    _lib.check(st, "ps_dec_final_dgrad")
    return act_out


def conv5x5_tc_dgrad(g_in, w_packed_flipped, mask_act, n_img, H, W, precision, act_out=None, out_tokens=None):
    """ data gradient of a 64 -> 64 5x5 convolution + the ReLU in front of it: conv(g_in, flipped weights) masked by
    (mask_act > 0); result as padded split (act_out) or fp32 NHWC tokens (out_tokens) """
    lib = _lib.load()
    w_tc, w_scale = w_packed_flipped
    mode = 0 if act_out is not None else 2
    with _timed("conv5x5_tc_dgrad_p%d" % precision, (n_img, 64, 64, H, W, 5)):
        st = lib.ps_conv5x5_tc_dgrad(_ptr(g_in), _ptr(w_tc), float(w_scale), _ptr(mask_act), _ptr(act_out), _ptr(out_tokens),
                                     n_img, H, W, mode, int(precision), _stream())
    _lib.check(st, "ps_conv5x5_tc_dgrad")
    return act_out if mode == 0 else out_tokens


def dec_expand_bwd(g_tokens, n_img, H, W, ksize, g_scale=None):
    lib = _lib.load()
    dG = torch.empty(n_img, ksize * ksize * 64, device=g_tokens.device, dtype=torch.float32)
    with _timed("dec_expand_bwd"):
        st = lib.ps_dec_expand_bwd(_ptr(g_tokens), _ptr(dG), _ptr(g_scale), n_img, H, W, ksize, _stream())
    _lib.check(st, "ps_dec_expand_bwd")
    return dG


# ------------------------------------------------------------------------------------------------ fused loss / metric tail

def _frame_layout(t, name):
    """ (B, F, ...) tensor whose frames (all trailing dimensions) are contiguous -> (B, F, elems, video stride, frame stride) """
    _req_f32_cuda(t, name)
    if t.dim() < 3:
        raise ValueError(f"{name}: expected (B, F, ...), got {tuple(t.shape)}")
    B, Fr = t.shape[0], t.shape[1]
    elems = 1
    for d in t.shape[2:]:
        elems *= d
    expect = 1
    for d, st in zip(reversed(t.shape[2:]), reversed(t.stride()[2:])):
        if d != 1 and st != expect:
            raise ValueError(f"{name}: the trailing (per-frame) dimensions must be contiguous")
        expect *= d
    return B, Fr, elems, t.stride(0), t.stride(1)


def _req_f32_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise _lib.PlaySlotKernelError(f"{name} lives on {t.device}: playslot_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")


def frame_sq_err(pred, target, clamp01=False):
    """ pred, target (B, F, ...) (time slices of a clip are read in place) -> (B, F) sums of squared differences per frame """
    lib = _lib.load()
    B, Fr, elems, pv, pf = _frame_layout(pred, "pred")
    B2, F2, e2, tv, tf = _frame_layout(target, "target")
    if (B, Fr, elems) != (B2, F2, e2):
        raise ValueError(f"frame_sq_err: pred {tuple(pred.shape)} vs target {tuple(target.shape)}")
    out = torch.empty(B, Fr, device=pred.device, dtype=torch.float32)
    with _timed("frame_sq_err"):
        st = lib.ps_frame_sq_err(_ptr(pred), _ptr(target), _ptr(out), B, Fr, elems, pv, pf, tv, tf, int(clamp01), _stream())
    _lib.check(st, "ps_frame_sq_err")
    return out


def mse_bwd(pred, target, g, scale):
    """ 2 * scale * g * (pred - target) for a contiguous pred (B, F, ...) and a target addressed like in frame_sq_err """
    lib = _lib.load()
    _req(pred, "pred")
    B, Fr, elems, _, _ = _frame_layout(pred, "pred")
    _, _, _, tv, tf = _frame_layout(target, "target")
    d = torch.empty_like(pred)
    with _timed("mse_bwd"):
        st = lib.ps_mse_bwd(_ptr(pred), _ptr(target), _ptr(g), float(scale), _ptr(d), B, Fr, elems, tv, tf, _stream())
    _lib.check(st, "ps_mse_bwd")
    return d


# ------------------------------------------------------------------------------------------------ SAVi training kernels

def sa_core_fwd(A, xn, eps, want_masks=True):
    """ A (B,S,128) folded queries, xn (B,N,128) normalised features -> (t (B,S,128), Z (B,S), masks (B,S,N) or None) """
    lib = _lib.load()
    _req(A, "A"); _req(xn, "xn")
    B, S, D = A.shape
    N = xn.shape[1]
    assert D == 128 and xn.shape == (B, N, 128), "sa_core: 128 features"
    ch = int(lib.ps_sa_core_chunks(N))
    dev = A.device
    t = torch.empty(B, S, D, device=dev, dtype=torch.float32)
    Z = torch.empty(B, S, device=dev, dtype=torch.float32)
    masks = torch.empty(B, S, N, device=dev, dtype=torch.float32) if want_masks else None
    pt = torch.empty(B, ch, S, D, device=dev, dtype=torch.float32)
    pz = torch.empty(B, ch, S, device=dev, dtype=torch.float32)
    with _timed("sa_core_fwd"):
        st = lib.ps_sa_core_fwd(_ptr(A), _ptr(xn), _ptr(t), _ptr(Z), _ptr(masks), _ptr(pt), _ptr(pz), B, S, N, float(eps), _stream())
    _lib.check(st, "ps_sa_core_fwd")
    return t, Z, masks


def sa_core_bwd(A, xn, t, Z, dt, eps, dxn=None):
    """ -> (dA (B,S,128), dxn (B,N,128)); with `dxn` given the feature gradient is accumulated into it """
    lib = _lib.load()
    B, S, D = A.shape
    N = xn.shape[1]
    ch = int(lib.ps_sa_core_chunks(N))
    dev = A.device
    dA = torch.empty(B, S, D, device=dev, dtype=torch.float32)
    acc = dxn is not None
    if dxn is None:
        dxn = torch.empty_like(xn)
    pd = torch.empty(B, ch, S, D, device=dev, dtype=torch.float32)
    with _timed("sa_core_bwd"):
        st = lib.ps_sa_core_bwd(_ptr(A), _ptr(xn), _ptr(t), _ptr(Z), _ptr(_req(dt, "dt")), _ptr(dA), _ptr(dxn), _ptr(pd), B, S, N, float(eps),
                                int(acc), _stream())
    _lib.check(st, "ps_sa_core_bwd")
    return dA, dxn


def gru_gates_fwd(gi, gh, h):
    lib = _lib.load()
    _req(gi, "gi"); _req(gh, "gh"); _req(h, "h")
    D = h.shape[-1]
    M = h.numel() // D
    out = torch.empty_like(h)
    gates = torch.empty_like(gi)
    st = lib.ps_gru_gates_fwd(_ptr(gi), _ptr(gh), _ptr(h), _ptr(out), _ptr(gates), M, D, _stream())
    _lib.check(st, "ps_gru_gates_fwd")
    return out, gates


def gru_gates_bwd(dhout, gates, gh, h):
    lib = _lib.load()
    D = h.shape[-1]
    M = h.numel() // D
    dgi, dgh, dh = torch.empty_like(gates), torch.empty_like(gates), torch.empty_like(h)
    st = lib.ps_gru_gates_bwd(_ptr(_req(dhout, "dhout")), _ptr(gates), _ptr(gh), _ptr(h), _ptr(dgi), _ptr(dgh), _ptr(dh), M, D, _stream())
    _lib.check(st, "ps_gru_gates_bwd")
    return dgi, dgh, dh


WGRAD_IMPL = os.environ.get("PLAYSLOT_WGRAD_IMPL", "tc")      # "tc": tcgen05 kernel where it applies; "ffma": CUDA-core kernel


def conv5x5_planes_wgrad(dy_planes, x_planes, n_img, H, W, dy_precision, x_precision, scale=1.0, impl=None):
    """ -> dW (64 cout, 64 cin, 5, 5) = scale * correlation of the gradient planes with the activation planes """
    lib = _lib.load()
    dW = torch.zeros(25, 64, 64, device=dy_planes.device, dtype=torch.float32)
    impl = impl or WGRAD_IMPL
    if impl == "tc" and dy_precision == 3 and x_precision == 3 and W <= 68:
        with _timed("conv5x5_wgrad_tc", (n_img, 64, 64, H, W, 5)):
            st = lib.ps_conv5x5_planes_wgrad_tc(_ptr(dy_planes), _ptr(x_planes), _ptr(dW), n_img, H, W, 3, float(scale), _stream())
        _lib.check(st, "ps_conv5x5_planes_wgrad_tc")
        return dW.view(5, 5, 64, 64).permute(2, 3, 0, 1).contiguous()
    with _timed("conv5x5_wgrad", (n_img, 64, 64, H, W, 5)):
        st = lib.ps_conv5x5_planes_wgrad(_ptr(dy_planes), _ptr(x_planes), _ptr(dW), n_img, H, W, int(dy_precision), int(x_precision),
                                         float(scale), _stream())
    _lib.check(st, "ps_conv5x5_planes_wgrad")
    return dW.view(5, 5, 64, 64).permute(2, 3, 0, 1).contiguous()


def planes_colsum(planes, n_img, H, W, precision, scale=1.0):
    lib = _lib.load()
    out = torch.zeros(64, device=planes.device, dtype=torch.float32)
    st = lib.ps_planes_colsum(_ptr(planes), _ptr(out), n_img, H, W, int(precision), float(scale), _stream())
    _lib.check(st, "ps_planes_colsum")
    return out


def conv_first_wgrad(dh_planes, x, n_img, H, W, ksize, precision, scale=1.0):
    """ x (n, Cin, H, W) fp32 or uint8 -> dW (64, Cin, k, k) """
    lib = _lib.load()
    Cin = x.shape[1]
    dW = torch.zeros(64, Cin, ksize, ksize, device=x.device, dtype=torch.float32)
    u8 = x.dtype == torch.uint8
    st = lib.ps_conv_first_wgrad(_ptr(dh_planes), None if u8 else _ptr(x), _ptr(x) if u8 else None, _ptr(dW), n_img, Cin, H, W, ksize,
                                 int(precision), float(scale), _stream())
    _lib.check(st, "ps_conv_first_wgrad")
    return dW


def conv3x3_final_wgrad(g_o, h):
    """ g_o (n, 4, H, W), h (n, 64, H, W) -> (dW (4, 64, 3, 3), db (4,)) """
    lib = _lib.load()
    _req(g_o, "g_o"); _req(h, "h")
    n, _, H, W = h.shape
    dW = torch.zeros(4, 64, 3, 3, device=h.device, dtype=torch.float32)
    db = torch.zeros(4, device=h.device, dtype=torch.float32)
    with _timed("conv3x3_final_wgrad"):
        st = lib.ps_conv3x3_final_wgrad(_ptr(g_o), _ptr(h), _ptr(dW), _ptr(db), n, H, W, _stream())
    _lib.check(st, "ps_conv3x3_final_wgrad")
    return dW, db


def conv3x3_final_wgrad_planes(g_o, h_planes, n, H, W, precision):
    """ conv3x3_final_wgrad with h as padded split planes """
    lib = _lib.load()
    _req(g_o, "g_o")
    dW = torch.zeros(4, 64, 3, 3, device=g_o.device, dtype=torch.float32)
    db = torch.zeros(4, device=g_o.device, dtype=torch.float32)
    with _timed("conv3x3_final_wgrad"):
        st = lib.ps_conv3x3_final_wgrad_planes(_ptr(g_o), _ptr(h_planes), int(precision), _ptr(dW), _ptr(db), n, H, W, _stream())
    _lib.check(st, "ps_conv3x3_final_wgrad_planes")
    return dW, db


def colsum(x2d, scale=1.0):
    """ (M, N) -> (N,) column sums times scale """
    lib = _lib.load()
    _req(x2d, "x")
    M, N = x2d.shape
    out = torch.empty(N, device=x2d.device, dtype=torch.float32)
    st = lib.ps_colsum(_ptr(x2d), _ptr(out), M, N, float(scale), _stream())
    _lib.check(st, "ps_colsum")
    return out


def nhwc_to_split_scaled(tokens, planes, n_img, H, W, precision, scale=None):
    """ tokens (n_img, H*W, ld >= 64) fp32 * scale[0] -> padded split planes (pre-zeroed borders) """
    lib = _lib.load()
    _req(tokens, "tokens")
    st = lib.ps_nhwc_to_split_scaled(_ptr(tokens), tokens.shape[-1], _ptr(scale), _ptr(planes), n_img, H, W, int(precision), _stream())
    _lib.check(st, "ps_nhwc_to_split_scaled")
    return planes


# ------------------------------------------------------------------------------------------------ fused transformer stack

FUSED_TBLOCK = os.environ.get("PLAYSLOT_FUSED_TBLOCK", "1") != "0"


def _pow2_scale_of(t):
    amax = float(t.abs().max())
    return 1.0 if amax == 0.0 else 2.0 ** (-round(math.log2(amax)))


def _split_fp16(w):
    h = w.to(torch.float16)
    lo = ((w - h.float()) * TC_LO_SCALE).to(torch.float16)
    return torch.stack([h, lo], dim=0)


def tblock_stack_applicable(blocks, L):
    """ can this list of pre-LN transformer blocks run as ONE fused launch (csrc/tblock_tc.cu)? """
    if not FUSED_TBLOCK or _LINEAR_MODE != "fp16x3" or not blocks or L > 128:
        return False
    b0 = blocks[0]
    for b in blocks:
        a = b.attn
        if (b.embed_dim != 128 or b.mlp_size != 512 or a.num_heads not in (4, 8) or a.head_dim > 64 or a.num_heads != b0.attn.num_heads
                or a.head_dim != b0.attn.head_dim or getattr(b, "use_cross_attn", False) or b.mlp._act != ACT_GELU
                or (not a.has_out_proj() and a.inner_dim != 128)):
            return False
    return True


def pack_tblock_stack(blocks):
    """ packed operands of ps_tblock_stack for a list of blocks (attention.py TransformerEncoderBlock / TransformerDecoder mirrors) """
    a0 = blocks[0].attn
    H, dh = a0.num_heads, a0.head_dim
    dev = a0.q.weight.device
    WA, WB, inv, vecs = [], [], [], {k: [] for k in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "b_o", "b1", "b2")}
    for b in blocks:
        a = b.attn
        q, k, v = a.q.weight.detach().float(), a.k.weight.detach().float(), a.v.weight.detach().float()       # (H*dh, 128)
        fc1, fc2 = b.mlp.mlp[0], b.mlp.mlp[2]
        if a.has_out_proj():
            wo, bo = a.out_proj[0].weight.detach().float(), a.out_proj[0].bias.detach().float()             # (128, H*dh)
        else:
            wo, bo = torch.eye(128, device=dev), torch.zeros(128, device=dev)                                   # heads are concatenated
        s_qkv = _pow2_scale_of(torch.cat([q, k, v]))
        s_1, s_o, s_2 = _pow2_scale_of(fc1.weight.detach()), _pow2_scale_of(wo), _pow2_scale_of(fc2.weight.detach())
        rows = []
        for h in range(H):
            for w in (q, k, v):
                blk = torch.zeros(64, 128, device=dev)
                blk[:dh] = w[h * dh:(h + 1) * dh] * s_qkv
                rows.append(blk)
        rows.append(fc1.weight.detach().float() * s_1)
        WA.append(_split_fp16(torch.cat(rows, dim=0)))                                                        # (2, H*192 + 512, 128)
        cols = []
        for h in range(H):
            blk = torch.zeros(128, 64, device=dev)
            blk[:, :dh] = wo[:, h * dh:(h + 1) * dh] * s_o
            cols.append(blk)
        cols.append(fc2.weight.detach().float() * s_2)
        WB.append(_split_fp16(torch.cat(cols, dim=1)))                                                        # (2, 128, H*64 + 512)
        inv.append(torch.tensor([1.0 / s_qkv, 1.0 / s_1, 1.0 / s_o, 1.0 / s_2], device=dev))
        vecs["ln1_g"].append(b.ln_att.weight.detach().float()); vecs["ln1_b"].append(b.ln_att.bias.detach().float())
        vecs["ln2_g"].append(b.ln_mlp.weight.detach().float()); vecs["ln2_b"].append(b.ln_mlp.bias.detach().float())
        vecs["b_o"].append(bo); vecs["b1"].append(fc1.bias.detach().float()); vecs["b2"].append(fc2.bias.detach().float())
    out = {k: torch.stack(v).contiguous() for k, v in vecs.items()}
    out.update(WA=torch.stack(WA).view(torch.int16).contiguous(), WB=torch.stack(WB).view(torch.int16).contiguous(),
               inv_scale=torch.stack(inv).contiguous(), heads=H, qscale=dh ** -0.5, n_layers=len(blocks),
               eps1=blocks[0].ln_att.eps, eps2=blocks[0].ln_mlp.eps)
    return out


def tblock_stack(x, pack, n_seq, L, spf):
    """ x (n_seq * L, 128) -> the stack's output (new tensor) """
    lib = _lib.load()
    _req(x, "x")
    assert x.numel() == n_seq * L * 128
    y = x.clone()
    ws = torch.empty(int(lib.ps_tblock_stack_workspace_floats(int(n_seq), int(L))), device=x.device, dtype=torch.float32)
    with _timed("tblock_stack", (n_seq, L, pack["n_layers"])):
        st = lib.ps_tblock_stack(_ptr(y), int(n_seq), int(L), int(spf), pack["n_layers"], pack["heads"], float(pack["qscale"]),
                                 float(pack["eps1"]), float(pack["eps2"]), _ptr(pack["WA"]), _ptr(pack["WB"]), _ptr(pack["ln1_g"]),
                                 _ptr(pack["ln1_b"]), _ptr(pack["ln2_g"]), _ptr(pack["ln2_b"]), _ptr(pack["b_o"]), _ptr(pack["b1"]),
                                 _ptr(pack["b2"]), _ptr(pack["inv_scale"]), _ptr(ws), _stream())
    _lib.check(st, "ps_tblock_stack")
    return y
