"""Staged on-device diagnostics for the tcgen05 kernels (run on the B200 box, each stage in its own
subprocess under a timeout so a hung kernel cannot take the others down).

    python tools/gpu_probe.py            # all stages
    python tools/gpu_probe.py conv_tc    # one stage, in-process
"""
from __future__ import annotations

import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

STAGES = ["simt", "conv_tc", "wgrad_tc", "net"]


def stats(name, out, ref):
    import torch
    out, ref = out.double(), ref.double()
    diff = (out - ref).abs()
    idx = int(diff.argmax())
    print(f"[{name}] rel_l2={float((out - ref).norm() / (ref.norm() + 1e-30)):.3e} max_abs={float(diff.max()):.3e} "
          f"ref_max={float(ref.abs().max()):.3e} worst_idx={idx} out={float(out.flatten()[idx]):.5f} "
          f"ref={float(ref.flatten()[idx]):.5f} nan={int(torch.isnan(out).sum())}", flush=True)


def stage_simt():
    import torch, torch.nn.functional as F
    from sinddm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    x = torch.randn(2, 80, 19, 23, device=dev)
    w = torch.randn(80, 80, 3, 3, device=dev) / 27
    wf, wd = ops.pack_conv_weights(w)
    r = ops.conv_forward(x.permute(0, 2, 3, 1).contiguous(), wf, math=0)
    stats("simt conv 80->80", r["out"].permute(0, 3, 1, 2), F.conv2d(x.double(), w.double(), padding=1))
    dy = torch.randn(2, 80, 19, 23, device=dev)
    wz = torch.zeros(80, 80, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    (dw,) = torch.autograd.grad(F.conv2d(x.double(), wz, padding=1), wz, dy.double())
    out = ops.conv_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), 9, math=0)
    stats("simt wgrad 80x80", out, dw)


def stage_conv_tc():
    import torch, torch.nn.functional as F
    from sinddm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    for (B, H, W, Ci, Co) in [(1, 8, 16, 32, 16), (1, 8, 16, 80, 80), (2, 19, 23, 160, 160), (4, 94, 126, 80, 160)]:
        x = torch.randn(B, Ci, H, W, device=dev)
        w = torch.randn(Co, Ci, 3, 3, device=dev) / (3 * Ci ** 0.5)
        wf, wd = ops.pack_conv_weights(w, round_tf32=True)
        t0 = time.time()
        r = ops.conv_forward(x.permute(0, 2, 3, 1).contiguous(), wf, math=1)
        torch.cuda.synchronize()
        stats(f"tc conv {Ci}->{Co} {B}x{H}x{W} ({time.time() - t0:.3f}s)", r["out"].permute(0, 3, 1, 2),
              F.conv2d(x.double(), w.double(), padding=1))


def stage_wgrad_tc():
    import torch, torch.nn.functional as F
    from sinddm_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    for (B, H, W, Cx, Cy, nt) in [(1, 4, 32, 32, 16, 1), (1, 8, 32, 32, 32, 9), (2, 19, 23, 80, 80, 9),
                                   (2, 33, 70, 160, 160, 9), (2, 19, 23, 160, 80, 1)]:
        k = 3 if nt == 9 else 1
        x = torch.randn(B, Cx, H, W, device=dev)
        dy = torch.randn(B, Cy, H, W, device=dev)
        wz = torch.zeros(Cy, Cx, k, k, device=dev, dtype=torch.float64, requires_grad=True)
        (dw,) = torch.autograd.grad(F.conv2d(x.double(), wz, padding=k // 2), wz, dy.double())
        out = ops.conv_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), nt, math=1)
        torch.cuda.synchronize()
        stats(f"tc wgrad Cx={Cx} Cy={Cy} taps={nt} {B}x{H}x{W}", out, dw)


def stage_net():
    sys.path.insert(0, str(REPO))
    import __graft_entry__ as ge
    ge.smoke()


def main():
    if len(sys.argv) > 1:
        globals()["stage_" + sys.argv[1]]()
        return
    for st in STAGES:
        print(f"===== stage {st} =====", flush=True)
        try:
            p = subprocess.run([sys.executable, __file__, st], timeout=240, capture_output=True, text=True)
            print(p.stdout[-4000:])
            if p.returncode != 0:
                print(f"stage {st} exit code {p.returncode}\n{p.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"stage {st} TIMED OUT (hung kernel?)\n{(e.stdout or b'')[-2000:]}\n{(e.stderr or b'')[-2000:]}")


if __name__ == "__main__":
    main()
