"""A/B timing of one training step per scale and one sampling pass (device time, CUDA events) under the current
environment switches; prints one line.  usage: SINDDM_TC_ISSUERS=1 python tools/ab_probe.py"""
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch  # noqa: E402

import bench  # noqa: E402
from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
tmp = Path(tempfile.mkdtemp())
bench.synthetic_pyramid(tmp, bench.BALLOONS_SIZES)
net = SinDDMNet(dim=bench.DIM, multiscale=True, device=dev).to(dev)
dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=bench.BALLOONS_SIZES,
                                  timesteps=100, train_full_t=True, scale_losses=bench.BALLOONS_SCALE_LOSSES,
                                  device=dev, results_folder=str(tmp / "res")).to(dev)
dif.num_timesteps_ideal = list(bench.BALLOONS_T_IDEAL)
tr = MultiscaleTrainer(dif, str(tmp) + "/", n_scales=5, image_sizes=bench.BALLOONS_SIZES, train_batch_size=bench.BATCH,
                       train_lr=1e-3, gradient_accumulate_every=1, avg_window=10 ** 9, results_folder=str(tmp / "res"),
                       device=dev)
tr._prepare_training()
tr.step = 1
res = {}
for s in (0, 1, 2, 3, 4):
    for _ in range(3):
        tr.train_step(s=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        tr.train_step(s=s)
    e1.record()
    torch.cuda.synchronize()
    res[f"train_s{s}_ms"] = round(e0.elapsed_time(e1) / 8, 3)


def sample():
    return tr.sample_scales(scale_mul=(1, 1), custom_sample=True, batch_size=bench.SAMPLE_BATCH,
                            custom_t_list=bench.BALLOONS_T_IDEAL[1:], save_images=False)


sample()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sample()
sample()
e1.record()
torch.cuda.synchronize()
res["sample16_ms"] = round(e0.elapsed_time(e1) / 2, 2)
# per-scale sampling time (second pass: plans and step graphs exist)
per = []
img = None
for s, total in enumerate(bench.BALLOONS_T_IDEAL):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if s == 0:
        img = dif.sample(batch_size=bench.SAMPLE_BATCH)
    else:
        img = dif.sample_via_scale(bench.SAMPLE_BATCH, img, s=s, scale_mul=(1, 1), custom_sample=True, custom_img_size_idx=s,
                                   custom_t=total)
    b.record()
    torch.cuda.synchronize()
    per.append(round(a.elapsed_time(b), 2))
res["sample_per_scale_ms"] = per
res["env"] = {k: v for k, v in os.environ.items() if k.startswith("SINDDM_")}
print(res)
