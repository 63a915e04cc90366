"""CPU-side cost of one training step, piece by piece (no device syncs inside the timed pieces)."""
import sys, tempfile, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch
import bench
from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet

dev = "cuda:0"
tmp = Path(tempfile.mkdtemp())
bench.synthetic_pyramid(tmp, bench.BALLOONS_SIZES)
net = SinDDMNet(dim=160, multiscale=True, device=dev).to(dev)
dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=bench.BALLOONS_SIZES,
                                  timesteps=100, train_full_t=True, scale_losses=bench.BALLOONS_SCALE_LOSSES,
                                  device=dev, results_folder=str(tmp / "r")).to(dev)
tr = MultiscaleTrainer(dif, str(tmp) + "/", n_scales=5, image_sizes=bench.BALLOONS_SIZES, train_batch_size=32,
                       train_lr=1e-3, gradient_accumulate_every=1, avg_window=10 ** 9, results_folder=str(tmp / "r"),
                       device=dev)
tr._prepare_training()
tr.step = 1
for s in range(5):
    tr.train_step(s=s)
torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
N = 20
s = 0
for it in range(N):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sm = torch.multinomial(tr._s_weights, 1); si = int(sm); tick("multinomial+int", t0)
    t0 = time.perf_counter(); loss = tr.model(tr.data_list[s], s); tick("forward", t0)
    t0 = time.perf_counter(); loss.backward(); tick("backward", t0)
    t0 = time.perf_counter(); tr.bucket.all_reduce_mean(); tick("bucket", t0)
    t0 = time.perf_counter(); tr.opt.step(); tick("opt.step", t0)
    t0 = time.perf_counter(); tr.opt.zero_grad(); tick("zero_grad", t0)
    t0 = time.perf_counter(); tr.step_ema() if it % 10 == 0 else None; tick("ema", t0)
    t0 = time.perf_counter(); tr.scheduler.step(); tick("sched", t0)
    t0 = time.perf_counter(); torch.cuda.synchronize(); tick("gpu_tail_sync", t0)
for k, v in acc.items():
    print(f"{k:18s} {v / N * 1e3:8.3f} ms")
print("total cpu+sync", sum(acc.values()) / N * 1e3)
# whole train_step, async
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(N): tr.train_step(s=0)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"train_step(s=0): cpu issue {(t1 - t0) / N * 1e3:.3f} ms/step, with final sync {(t2 - t0) / N * 1e3:.3f} ms/step")
