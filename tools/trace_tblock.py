""" development tool: per-phase timeline of the fused transformer-stack kernel (CTA 0, %globaltimer stamps) """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops, _lib
from playslot_b200.models.BlocksUtils.attention import TransformerDecoder

torch.manual_seed(0)
n_layers, n_seq, L, spf = 8, 64, 48, 6
blocks = [TransformerDecoder(embed_dim=128, head_dim=64, num_heads=8, mlp_size=512, project_out=True).cuda().eval() for _ in range(n_layers)]
pack = ops.pack_tblock_stack(blocks)
x = torch.randn(n_seq * L, 128, device="cuda")
dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
ws = torch.empty(int(_lib.load().ps_tblock_stack_workspace_floats(n_seq, L)), device="cuda")
lib = _lib.load()
for rep in range(3):
    y = x.clone()
    st = lib.ps_tblock_stack_dbg(y.data_ptr(), n_seq, L, spf, n_layers, pack["heads"], float(pack["qscale"]), float(pack["eps1"]), float(pack["eps2"]),
                                 pack["WA"].data_ptr(), pack["WB"].data_ptr(), pack["ln1_g"].data_ptr(), pack["ln1_b"].data_ptr(),
                                 pack["ln2_g"].data_ptr(), pack["ln2_b"].data_ptr(), pack["b_o"].data_ptr(), pack["b1"].data_ptr(),
                                 pack["b2"].data_ptr(), pack["inv_scale"].data_ptr(), ws.data_ptr(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0
torch.cuda.synchronize()
t = dbg.cpu().tolist()
t = [v for v in t if v > 0]
t0 = t[0]
print("stamps", len(t), "total us", (t[-1] - t0) / 1e3)
per_layer = len(t) // n_layers
names = []
for hh in range(2):
    for nm in ("qk", "v", "S", "PV"):
        names += [f"h{hh} wait {nm} begin", f"h{hh} wait {nm} end"]
names += ["red1 waitacc b", "red1 waitacc e", "red1 sync1 b", "red1 sync1 e", "red1 sync2 b", "red1 sync2 e",
          "fc1 wait b", "fc1 wait e", "red2 waitacc b", "red2 waitacc e", "red2 sync1 b", "red2 sync1 e", "red2 sync2 b", "red2 sync2 e"]
for layer in (0, 3):
    base = layer * per_layer
    print(f"--- layer {layer} ({per_layer} stamps per layer)")
    prev = t[base - 1] if base > 0 else t0
    for i in range(per_layer):
        print(f"{names[i] if i < len(names) else i:24s} +{(t[base + i] - prev) / 1e3:7.2f} us   (t = {(t[base + i] - t0) / 1e3:8.2f})")
        prev = t[base + i]
# wall time of the launch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.tblock_stack(x, pack, n_seq, L, spf)
e1.record(); torch.cuda.synchronize()
print("launch avg us", e0.elapsed_time(e1) / 20 * 1e3)
