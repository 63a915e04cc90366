"""Command line of the B200-native SinDDM hot path -- flag-compatible with the reference's main.py.

    python main.py --scope balloons --mode train  --dataset_folder ./datasets/balloons/ --image_name balloons.png
    python main.py --scope balloons --mode sample --dataset_folder ./datasets/balloons/ --image_name balloons.png \
                   --load_milestone 12
    python -m torch.distributed.run --nproc-per-node 8 main.py --mode train ...     # data parallel over 8 B200

Every flag of the reference (main.py:15-58) is accepted.  Modes `train`, `sample`, `harmonization` and
`style_transfer` run on the sm_100a kernels; the CLIP / ROI modes (clip_content, clip_style_gen,
clip_style_trans, clip_roi, roi) are outside this repo's scope and exit with a
clear message instead of importing CLIP.
"""
from __future__ import annotations

import argparse
import sys

import torch

from SinDDM.functions import create_img_scales
from SinDDM.models import MultiScaleGaussianDiffusion, SinDDMNet
from SinDDM.trainer import MultiscaleTrainer
from sinddm_b200 import dist as spdist

OUT_OF_SCOPE_MODES = ("clip_content", "clip_style_gen", "clip_style_trans", "clip_roi", "roi")


def build_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scope", default="forest", help="choose training scope.")
    ap.add_argument("--mode", help="train | sample | harmonization | style_transfer (the CLIP / ROI modes of the reference are out of scope)")
    # harmonization / style transfer inputs; the CLIP / ROI flags are accepted so existing command lines parse
    ap.add_argument("--input_image", default="seascape_composite_dragon.png")
    ap.add_argument("--start_t_harm", default=5, type=int)
    ap.add_argument("--start_t_style", default=15, type=int)
    ap.add_argument("--harm_mask", default="seascape_mask_dragon.png")
    ap.add_argument("--clip_text", default="Fire in the Forest")
    ap.add_argument("--fill_factor", type=float)
    ap.add_argument("--strength", type=float)
    ap.add_argument("--roi_n_tar", default=1, type=int)
    # dataset
    ap.add_argument("--dataset_folder", default="./datasets/forest/")
    ap.add_argument("--image_name", default="forest.jpeg")
    ap.add_argument("--results_folder", default="./results/")
    # net / diffusion
    ap.add_argument("--dim", default=160, type=int)
    ap.add_argument("--scale_factor", default=1.411, type=float)
    ap.add_argument("--timesteps", default=100, type=int)
    # training
    ap.add_argument("--train_batch_size", default=32, type=int, help="GLOBAL batch (split over ranks under torchrun)")
    ap.add_argument("--grad_accumulate", default=1, type=int)
    ap.add_argument("--train_num_steps", default=120001, type=int)
    ap.add_argument("--save_and_sample_every", default=10000, type=int)
    ap.add_argument("--avg_window", default=100, type=int)
    ap.add_argument("--train_lr", default=1e-3, type=float)
    ap.add_argument("--sched_k_milestones", nargs="+", default=[20, 40, 70, 80, 90, 110], type=int)
    ap.add_argument("--load_milestone", default=0, type=int)
    # sampling
    ap.add_argument("--sample_batch_size", default=16, type=int)
    ap.add_argument("--scale_mul", nargs="+", default=[1, 1], type=float)
    ap.add_argument("--sample_t_list", nargs="+", type=int)
    ap.add_argument("--device_num", default=0, type=int)
    # dev flags of the reference
    ap.add_argument("--sample_limited_t", action="store_true")
    ap.add_argument("--omega", default=0, type=float)
    ap.add_argument("--loss_factor", default=1, type=float)
    # new: numerics of the dense convolutions (tf32 tensor cores | fp32 CUDA cores)
    ap.add_argument("--math", default=None, choices=["tf32", "tf32x3", "fp32"])
    # new: keep the pyramid in memory instead of writing scale_i/ and scale_i_recon/ next to the image
    ap.add_argument("--in_memory_pyramid", action="store_true")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.mode in OUT_OF_SCOPE_MODES:
        sys.exit(f"--mode {args.mode} is outside the scope of sinddm_b200 (train and sample are implemented)")
    if args.mode not in ("train", "sample", "harmonization", "style_transfer"):
        sys.exit("--mode must be train, sample, harmonization or style_transfer")

    rank, local_rank, world = spdist.init_process_group()
    if rank == 0:
        print("num devices: " + str(torch.cuda.device_count()))
    device = f"cuda:{local_rank if world > 1 else args.device_num}"
    torch.cuda.set_device(device)
    scale_mul = (args.scale_mul[0], args.scale_mul[1])
    results_folder = args.results_folder + "/" + args.scope

    # the pyramid is written next to the dataset image: only rank 0 creates it, everyone reads it
    pyramid = None
    if args.in_memory_pyramid:
        sizes, rescale_losses, scale_factor, n_scales, pyramid = create_img_scales(
            args.dataset_folder, args.image_name, scale_factor=args.scale_factor, create=False, auto_scale=50000,
            return_pyramid=True)
    else:
        if rank == 0:
            create_img_scales(args.dataset_folder, args.image_name, scale_factor=args.scale_factor, create=True,
                              auto_scale=50000)
        if world > 1:
            torch.distributed.barrier()
        sizes, rescale_losses, scale_factor, n_scales = create_img_scales(
            args.dataset_folder, args.image_name, scale_factor=args.scale_factor, create=False, auto_scale=50000)

    model = SinDDMNet(dim=args.dim, multiscale=True, device=device, math=args.math).to(device)
    diffusion = MultiScaleGaussianDiffusion(
        denoise_fn=model, save_interm=False, results_folder=results_folder, n_scales=n_scales,
        scale_factor=scale_factor, image_sizes=sizes, scale_mul=scale_mul, channels=3, timesteps=args.timesteps,
        train_full_t=True, scale_losses=rescale_losses, loss_factor=args.loss_factor, loss_type="l1", betas=None,
        device=device, reblurring=True, sample_limited_t=args.sample_limited_t, omega=args.omega).to(device)
    sample_t_list = diffusion.num_timesteps_ideal[1:] if args.sample_t_list is None else args.sample_t_list

    trainer = MultiscaleTrainer(
        diffusion, folder=args.dataset_folder, n_scales=n_scales, scale_factor=scale_factor, image_sizes=sizes,
        train_batch_size=args.train_batch_size, train_lr=args.train_lr, train_num_steps=args.train_num_steps,
        gradient_accumulate_every=args.grad_accumulate, ema_decay=0.995, fp16=False,
        save_and_sample_every=args.save_and_sample_every, avg_window=args.avg_window,
        sched_milestones=[k * 1000 for k in args.sched_k_milestones], results_folder=results_folder, device=device,
        pyramid=pyramid)

    if args.load_milestone > 0:
        trainer.load(milestone=args.load_milestone)
    if args.mode in ("harmonization", "style_transfer"):
        # reference main.py:294-320: start at the last scale from t = start_t_style / start_t_harm
        import os
        start_s = n_scales - 1
        use_hist = args.mode == "style_transfer"
        custom_t = [0] * (n_scales - 1) + [args.start_t_style if use_hist else args.start_t_harm]
        trainer.ema_model.reblurring = True
        trainer.image2image(input_folder=os.path.join(args.dataset_folder, "i2i"), input_file=args.input_image,
                            mask=args.harm_mask, hist_ref_path=f"{args.dataset_folder}scale_{start_s}/",
                            batch_size=args.sample_batch_size, image_name=args.image_name, start_s=start_s,
                            custom_t=custom_t, scale_mul=(1, 1), device=device, use_hist=use_hist, save_unbatched=True,
                            auto_scale=50000, mode=args.mode)
        return
    if args.mode == "train":
        trainer.train()
    trainer.sample_scales(scale_mul=scale_mul, custom_sample=True, image_name=args.image_name,
                          batch_size=args.sample_batch_size, custom_t_list=sample_t_list)


if __name__ == "__main__":
    main()
