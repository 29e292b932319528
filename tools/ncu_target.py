""" development tool: a short program that launches the kernels we want to look at under ncu """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops
from playslot_b200.models.BlocksUtils.attention import SlotAttention

what = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
if what in ("all", "linear"):
    for (M, N, K, ln) in ((1 << 20, 128, 128, False), (3072, 1536, 128, True), (3072, 128, 512, False)):
        x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * K ** -0.5; b = torch.randn(N, device="cuda")
        g, be = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
        for _ in range(3):
            y = ops.linear(x, w, b, ln=(g, be, 1e-6) if ln else None)
        torch.cuda.synchronize()
if what in ("all", "sa"):
    S = 8
    m = SlotAttention(128, 128, S, 3, 1, 256).eval().cuda()
    feats = torch.randn(64, 4096, 128, device="cuda"); slots = torch.randn(64, S, 128, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            m(feats, slots, step=0)
    torch.cuda.synchronize()
if what in ("conv", "dec"):
    # decoder-sized launches: 4092 slot images 64x64 through the tensor-core convolution and the compositing kernel
    n, H, W = 4092, 64, 64
    a0 = ops.conv_tc_workspace(n, H, W, "cuda"); a1 = ops.conv_tc_workspace(n, H, W, "cuda")
    x = torch.randn(n, 64, H, W, device="cuda")
    ops.nchw_to_split(x, a0, 3)
    del x
    w = torch.randn(64, 64, 5, 5, device="cuda") * 0.03; b = torch.randn(64, device="cuda") * 0.1
    wp = ops.pack_conv_weight_tc(w, 3)
    for _ in range(3):
        ops.conv5x5_tc(a0, wp, b, n, H, W, True, 3, act_out=a1)
    torch.cuda.synchronize()
if what in ("encmlp",):
    n_tok = 1 << 20
    x = torch.randn(n_tok, 64, device="cuda")
    gam, bet = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    w1, b1 = torch.randn(128, 64, device="cuda") * 0.125, torch.zeros(128, device="cuda")
    w2, b2 = torch.randn(128, 128, device="cuda") * 0.09, torch.zeros(128, device="cuda")
    for _ in range(3):
        ops.enc_mlp_planes(x, 64, (gam, bet, 1e-5), (w1, b1), (w2, b2), 1e-5)
    torch.cuda.synchronize()
if what in ("mha",):
    qkv = torch.randn(3072, 1536, device="cuda")
    for _ in range(3):
        ops.mha(qkv, 64, 48, 8, 64, 6)
    torch.cuda.synchronize()
if what in ("composite",):
    from playslot_b200 import configs
    cfg = configs.experiment("GridShapes5Objs")
    savi, pred = configs.build_models(cfg, "cuda")
    slots = torch.randn(600, savi.num_slots, 128, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            savi.decode(slots)
    torch.cuda.synchronize()
if what in ("rolllin",):
    M = 3072
    x = torch.randn(M, 128, device="cuda"); g_, b_ = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    w1 = torch.randn(512, 128, device="cuda") * 0.09; b1 = torch.zeros(512, device="cuda")
    w2 = torch.randn(128, 512, device="cuda") * 0.05; b2 = torch.zeros(128, device="cuda")
    wq = torch.randn(1536, 128, device="cuda") * 0.09
    for _ in range(3):
        q = ops.linear(x, wq, None, ln=(g_, b_, 1e-6))
        h = ops.linear(x, w1, b1, ln=(g_, b_, 1e-6), act=ops.ACT_GELU, out_planes=True)
        y = ops.linear(h, w2, b2, res=x)
    torch.cuda.synchronize()
print("done")
