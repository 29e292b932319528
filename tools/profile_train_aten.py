""" development tool: which ATen ops (outside the C-ABI library) a training step launches, by call site -- torch.profiler with stacks """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from playslot_b200 import configs, train

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = configs.experiment("ButtonPress")
torch.manual_seed(0)
savi, pred = configs.build_models(cfg, "cuda")
for p in savi.parameters():
    p.requires_grad_(False)
opt = train.make_optimizer(pred)
bucket = train.GradientBucket(pred.parameters())
video = torch.rand(B, 14, 3, 64, 64, device="cuda")
for _ in range(3):
    train.train_step(savi, pred, opt, video, 6, 8, bucket=bucket)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=False) as prof:
    train.train_step(savi, pred, opt, video, 6, 8, bucket=bucket)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
print(prof.key_averages(group_by_stack_n=6).table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=50, max_src_column_width=110))
