"""Where does tc_conv_kernel spend its time?  Times the 160->160 3x3 layer of the finest balloons scale with the
kernel's diagnostic switches (SINDDM_TC_DEBUG: 1 = no operand loads, 2 = no epilogue traffic, 4 = no MMAs; results are
garbage when set, only the duration is of interest).

    python tools/conv_probe.py [Cin N]
"""
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch
from sinddm_b200 import ops

dev = "cuda:0"
B, H, W = 32, 186, 248
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 160
N = int(sys.argv[2]) if len(sys.argv) > 2 else 160
x = torch.randn(B, H, W, Cin, device=dev)
w = torch.randn(N, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)
wf, _ = ops.pack_conv_weights(w, round_tf32=True)
bias = torch.randn(N, device=dev)
z = torch.randn(B, H, W, N, device=dev)
flops = 2.0 * B * H * W * N * 9 * Cin


def run(label, dbg, **kw):
    os.environ["SINDDM_TC_DEBUG"] = str(dbg)
    for _ in range(2):
        ops.conv_forward(x, wf, math=1, bias=bias, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.conv_forward(x, wf, math=1, bias=bias, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{label:44s} dbg={dbg}  {ms:7.3f} ms  {flops / ms * 1e-9:7.1f} TFLOP/s", flush=True)


for name, kw in (("fwd gelu+pre (net[0])", dict(gelu=True, save_pre=True, round_tf32=True)),
                 ("plain (bias only)", dict()),
                 ("dgrad (x gelu'(z))", dict(dgelu_z=z, round_tf32=True))):
    print(f"--- {Cin}->{N} {name}")
    run("full", 0, **kw)
    run("no operand loads", 1, **kw)
    run("no epilogue traffic", 2, **kw)
    run("no loads, no epilogue traffic", 3, **kw)
    run("no MMAs (loads + epilogue)", 4, **kw)
    run("loads only", 6, **kw)
    run("epilogue only", 5, **kw)
    run("epilogue only, staged but not stored", 13, **kw)
    run("barrier chain only", 7, **kw)
os.environ["SINDDM_TC_DEBUG"] = "0"
