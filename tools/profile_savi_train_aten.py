""" development tool: which ATen kernels (outside the C-ABI library) the SAVi training step launches -- torch.profiler """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from playslot_b200 import configs, train

cfg = configs.experiment("BlockPush")
torch.manual_seed(0)
savi, _ = configs.build_models(cfg, "cuda")
opt = torch.optim.Adam(savi.parameters(), lr=1e-4)
video = torch.rand(32, 8, 3, 64, 64, device="cuda")
for _ in range(3):
    train.savi_train_step(savi, opt, video)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train.savi_train_step(savi, opt, video)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=90))
