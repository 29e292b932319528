import torch
x = torch.empty(1 << 29, device="cuda", dtype=torch.float32).normal_()   # 2 GB
y = torch.empty_like(x)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.sum())
print("read-only (torch.sum, 2 GB): %.0f GB/s" % (x.numel() * 4 / ms / 1e6))
ms = t(lambda: x.max())
print("read-only (torch.max, 2 GB): %.0f GB/s" % (x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x))
print("copy (read + write, 2+2 GB): %.0f GB/s" % (2 * x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.zero_())
print("write-only (memset 2 GB): %.0f GB/s" % (x.numel() * 4 / ms / 1e6))
