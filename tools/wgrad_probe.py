"""Times the tensor-core weight-gradient kernel (+ its split reduction) on the finest balloons scale."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch
from sinddm_b200 import ops

dev = "cuda:0"
B, H, W = 32, 186, 248
for Cx, Cy, ntaps in ((160, 160, 9), (160, 80, 9), (80, 160, 9), (80, 80, 9), (160, 80, 1), (80, 160, 1)):
    x = torch.randn(B, H, W, Cx, device=dev)
    dy = torch.randn(B, H, W, Cy, device=dev)
    for _ in range(2):
        ops.conv_wgrad(x, dy, ntaps, math=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.conv_wgrad(x, dy, ntaps, math=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * ntaps * Cx * Cy
    print(f"wgrad {Cx:3d}->{Cy:3d} taps {ntaps}: {ms:7.3f} ms  {fl / ms * 1e-9:7.1f} TFLOP/s (incl. split reduction)", flush=True)
