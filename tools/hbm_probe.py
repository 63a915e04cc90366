"""HBM write / read / copy rates on this box (GB/s), for judging the HBM-bound kernels."""
import torch
dev = "cuda:0"
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, bytes_, label):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:10s} {bytes_ / ms * 1e-6:8.1f} GB/s")
t(lambda: a.fill_(1.0), 4 * n, "write")
t(lambda: a.sum(), 4 * n, "read")
t(lambda: b.copy_(a), 8 * n, "copy")
