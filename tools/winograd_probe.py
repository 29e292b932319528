""" development tool (CPU, offline): numerics of a Winograd F(2x2, 5x5) formulation of the 64 -> 64 channel 5x5 convolution in the
fp16x3 split-product arithmetic of conv5x5_tc, against an fp64 direct convolution -- VERDICT r1 item 4(c): "probe Winograd
F(2x2,5x5) numerics with hi/lo planes offline before writing any kernel; record the result either way".

The transform matrices are built by Cook-Toom from 6 interpolation points (5 finite + infinity): A^T = [p_j^i], G = [p_j^k], and
B^T is solved from the bilinear identity  sum_j A^T[i,j] G[j,k] B^T[j,l] = delta(l, i + k)  in fp64; rows of B^T are normalised
to max |.| = 1 with the scale moved into G (weights are transformed offline).  Arithmetic emulated as the kernel would do it:
input transform in fp32 -> fp16 hi/lo split -> three fp16 products with fp32 accumulation over the 64 input channels per Winograd
point -> output transform in fp32.   python tools/winograd_probe.py > profiles/r02_winograd_probe.md """
import itertools
import numpy as np
import torch
import torch.nn.functional as F

torch.manual_seed(0)
LO = 2048.0


def cook_toom(points, m=2, r=5):
    n = m + r - 1
    AT = np.zeros((m, n)); G = np.zeros((n, r))
    for j, p in enumerate(points):
        if p is None:                       # the point at infinity
            AT[m - 1, j] = 1.0; G[j, r - 1] = 1.0
        else:
            AT[:, j] = [p ** i for i in range(m)]; G[j, :] = [p ** k for k in range(r)]
    rows = [(i, k) for i in range(m) for k in range(r)]
    M = np.array([[AT[i, j] * G[j, k] for j in range(n)] for (i, k) in rows])
    BT = np.zeros((n, n))
    for l in range(n):
        rhs = np.array([1.0 if i + k == l else 0.0 for (i, k) in rows])
        BT[:, l] = np.linalg.lstsq(M, rhs, rcond=None)[0]
    assert np.abs(M @ BT - np.array([[1.0 if i + k == l else 0.0 for l in range(n)] for (i, k) in rows])).max() < 1e-10
    s = np.abs(BT).max(axis=1, keepdims=True)
    return AT, G * s, BT / s


def split(x):
    hi = x.half().float()
    lo = ((x - hi) * LO).half().float()
    return hi, lo


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


def direct_fp16x3(x, w):
    s = 2.0 ** -round(float(torch.log2(w.abs().max())))        # power-of-two weight scale, as ops.pack_conv_weight_tc
    xh, xl = split(x); wh, wl = split(w * s)
    y = F.conv2d(xh, wh, padding=2) + (F.conv2d(xh, wl, padding=2) + F.conv2d(xl, wh, padding=2)) / LO
    return y / s


def winograd_fp16x3(x, w, AT, G, BT, exact_products=False):
    B, C, H, W = x.shape
    O = w.shape[0]
    AT_, BT_ = torch.tensor(AT, dtype=torch.float32), torch.tensor(BT, dtype=torch.float32)
    U = torch.einsum("ak,ockl,bl->ocab", torch.tensor(G), w.double(), torch.tensor(G))          # offline, fp64 -> (O, C, 6, 6)
    xp = F.pad(x, (2, 2, 2, 2))
    tiles = xp.unfold(2, 6, 2).unfold(3, 6, 2)                                                  # (B, C, H/2, W/2, 6, 6)
    V = torch.einsum("ak,bcyxkl,dl->bcyxad", BT_, tiles, BT_)                                   # fp32 input transform
    if exact_products:
        Mm = torch.einsum("ocad,bcyxad->boyxad", U, V.double()).float()
    else:
        us = 2.0 ** -torch.round(torch.log2(U.abs().amax(dim=(0, 1), keepdim=True)))           # per-point power-of-two scale
        Uh, Ul = split((U * us).float())
        Vh, Vl = split(V)
        Mm = (torch.einsum("ocad,bcyxad->boyxad", Uh, Vh)
              + (torch.einsum("ocad,bcyxad->boyxad", Ul, Vh) + torch.einsum("ocad,bcyxad->boyxad", Uh, Vl)) / LO) / us.float().reshape(1, 1, 1, 1, 6, 6)
    Y = torch.einsum("ia,boyxad,jd->boyxij", AT_, Mm, AT_)                                      # fp32 output transform (2 x 2 per tile)
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(B, O, H, W)


def main():
    B, C, H, W = 4, 64, 64, 64
    # decoder-like operands: post-ReLU activations (half of them zero, heavy-ish tail), Xavier-scale weights
    x = torch.relu(torch.randn(B, C, H, W)) * torch.rand(B, C, 1, 1) * 2.0
    w = torch.randn(C, C, 5, 5) * (2.0 / (C * 25 + C * 25)) ** 0.5
    ref = F.conv2d(x.double(), w.double(), padding=2)
    print("# Winograd F(2x2, 5x5) numerics probe (CPU emulation of the fp16x3 arithmetic; tools/winograd_probe.py)\n")
    print(f"operands: x = relu(randn) * U(0,2) per channel, ({B},{C},{H},{W}); w = Xavier-scale randn (64,64,5,5); reference = fp64 direct convolution\n")
    print("| formulation | rel-L2 error | max-rel error |")
    print("|---|---|---|")
    print("| direct, torch fp32 | %.2e | %.2e |" % rel(F.conv2d(x, w, padding=2), ref))
    print("| direct, fp16x3 (what conv5x5_tc computes) | %.2e | %.2e |" % rel(direct_fp16x3(x, w), ref))
    xh = x.half().float()
    print("| direct, single fp16 product | %.2e | %.2e |" % rel(F.conv2d(xh, w.half().float(), padding=2), ref))
    best = None
    for name, pts in (("0, +-1, +-2, inf", [0, 1, -1, 2, -2, None]), ("0, +-1, +-1/2, inf", [0, 1, -1, 0.5, -0.5, None]),
                      ("0, +-1/2, +-2, inf", [0, 0.5, -0.5, 2, -2, None]), ("0, +-1, 2, 1/2, inf", [0, 1, -1, 2, 0.5, None])):
        AT, G, BT = cook_toom(pts)
        e_exact = rel(winograd_fp16x3(x, w, AT, G, BT, exact_products=True), ref)
        e_x3 = rel(winograd_fp16x3(x, w, AT, G, BT), ref)
        amp = float(np.abs(AT).sum(axis=1).max() ** 2 * np.abs(BT).sum(axis=1).max() ** 2)
        print("| Winograd, points {%s}: fp32 transforms, exact products | %.2e | %.2e |" % ((name,) + e_exact))
        print("| Winograd, points {%s}: fp32 transforms, fp16x3 products (|A^T|_1^2 |B^T|_1^2 = %.0f) | %.2e | %.2e |" % ((name, amp) + e_x3))
        if best is None or e_x3[0] < best[1][0]:
            best = (name, e_x3)
    d = rel(direct_fp16x3(x, w), ref)
    print(f"\nbest Winograd point set {{{best[0]}}}: {best[1][0] / d[0]:.1f}x the rel-L2 error and {best[1][1] / d[1]:.1f}x the max-rel error of the direct fp16x3 convolution")
    # operation counts per 2x2 output tile and channel (for the cost side of the argument)
    print("\nMultiplies per output pixel and (cin, cout) pair: direct 25, Winograd 36 / 4 = 9 (2.78x fewer tensor-core MACs);")
    print("CUDA-core work added per 6x6 input tile and channel: B^T d B = 2 x 6 x (6x6 mat-vec) ~ 2 x 6 x 30 = 360 flops + 36 fp16 hi/lo splits (~5 ops each),")
    print("per 2x2 output tile and channel: A^T m A ~ 2 x (2x6)(6x6) ~ 100 flops: ~640 CUDA-core operations per 4 output pixel-channels.")


if __name__ == "__main__":
    main()
