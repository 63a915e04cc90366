"""Determinism stress: the same forward / backward repeated must give identical bits every time."""
import sys, os
import pathlib; REPO = pathlib.Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / 'tests'))
import torch
from test_gpu_net import build, rs_tensor, DEV
sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
net, dif = build("tf32", sizes=sizes, losses=[1.1, 0.78, 0.55, 0.39])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for B, s in ((32, 4), (1, 4), (4, 2), (16, 4)):
    h, w = sizes[s][1], sizes[s][0]
    x = rs_tensor(7 + B, (B, 3, h, w), 0.5).clamp(-1, 1).to(DEV)
    t = (torch.arange(B, device=DEV) * 3) % 100
    with torch.no_grad():
        ref = net(x, t, s).clone()
        for i in range(reps):
            y = net(x, t, s)
            if not torch.equal(y, ref):
                d = (y - ref).abs()
                nz = (d > 0).nonzero()
                bad += 1
                print(f"B={B} s={s} rep {i}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}, first {nz[:4].tolist()}", flush=True)
# training step determinism
x = rs_tensor(3, (8, 3, 94, 126), 0.5).clamp(-1, 1).to(DEV)
noise = rs_tensor(4, (8, 3, 94, 126)).to(DEV)
t = torch.arange(8, device=DEV) * 7
def grads():
    net.zero_grad()
    dif.p_losses(x, t, 2, noise=noise, x_orig=x).backward()
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
g0 = grads()
for i in range(reps // 2):
    g = grads()
    if not torch.equal(g, g0):
        bad += 1
        print(f"train rep {i}: {(g != g0).sum().item()} gradient elements differ, max {float((g - g0).abs().max()):.3e}", flush=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("SINDDM_")}, "nondeterministic results:", bad)
