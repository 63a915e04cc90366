"""Small driver for ncu: a few training steps at one balloons scale (and optionally a few sampling steps).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_step.py --scale 4 --steps 2
"""
import argparse
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=bench.BATCH)
    ap.add_argument("--sample-steps", type=int, default=0)
    args = ap.parse_args()
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet
    dev = "cuda:0"
    torch.manual_seed(0)
    tmp = Path(tempfile.mkdtemp())
    bench.synthetic_pyramid(tmp, bench.BALLOONS_SIZES)
    net = SinDDMNet(dim=bench.DIM, multiscale=True, device=dev).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=bench.BALLOONS_SIZES,
                                      timesteps=100, train_full_t=True, scale_losses=bench.BALLOONS_SCALE_LOSSES,
                                      device=dev, results_folder=str(tmp / "res")).to(dev)
    tr = MultiscaleTrainer(dif, str(tmp) + "/", n_scales=5, image_sizes=bench.BALLOONS_SIZES,
                           train_batch_size=args.batch, train_lr=1e-3, gradient_accumulate_every=1,
                           avg_window=10 ** 9, results_folder=str(tmp / "res"), device=dev)
    tr._prepare_training()
    tr.step = 1
    for _ in range(args.steps):
        tr.train_step(s=args.scale)
    torch.cuda.synchronize()
    if args.sample_steps:
        h, w = dif.image_sizes[args.scale]
        x = torch.randn(bench.SAMPLE_BATCH, 3, h, w, device=dev)
        dif.img_prev_upsample = torch.randn_like(x).clamp(-1, 1)
        for i in range(args.sample_steps):
            x = dif.p_sample(x, torch.full((bench.SAMPLE_BATCH,), 10, device=dev, dtype=torch.long), args.scale)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
