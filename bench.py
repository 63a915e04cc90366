"""Benchmark of the SinDDM hot path on B200 (and the reference CPU arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline metric (BASELINE.json): training steps/s on the balloons pyramid (5 scales, 48x64 .. 186x248),
batch 32 per GPU, T=100, dim=160 -- one "step" = one MultiscaleTrainer optimizer step (q_sample, denoiser
forward + backward, L1 loss, [gradient all-reduce], Adam, EMA cadence, LR scheduler) at the scale the
step draws; scales are visited round-robin so every window of 5 steps is the expectation of the
reference's uniform multinomial.  Also reported: sampling images/s for sample_scales(16 images, balloons'
T list [100,52,41,31,22] = 246 denoiser evaluations per image) in the `sampling` object.

Prints ONE JSON line (see the task contract): value = whole-job batch-32 steps/s, e2e = the same through the
public API with host-resident inputs copied in every step and the loss read back, roofline for the dominant
tcgen05 kernel from CUDA events recorded around its launches during the timed region, cpu_baseline = the
CPU oracle (a torch-CPU restatement of the reference, the reference being pure PyTorch) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BALLOONS_SIZES = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]     # (W, H), SURVEY.md section 8
BALLOONS_T_IDEAL = [100, 52, 41, 31, 22]                                       # probed from the reference
BALLOONS_SCALE_LOSSES = [1.20, 0.85, 0.60, 0.42]   # any values: train_full_t=True, T list given explicitly
DIM = 160
BATCH = 32
SAMPLE_BATCH = 16
FWD_FLOP_PER_PX = 2.150e6
TRAIN_FLOP_PER_PX = 6.45e6


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                     0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
            while True:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
                if self.stop_flag.wait(0.05):       # at least one sample is always taken
                    break
        except Exception as e:  # NVML missing: report nulls rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def synthetic_pyramid(folder: Path, sizes, seed=0):
    """Synthetic balloons-shaped training data: one smooth random image per scale plus its blurry twin, in the
    folder layout MultiscaleTrainer reads (scale_i/, scale_i_recon/)."""
    from PIL import Image
    rs = np.random.RandomState(seed)
    w_f, h_f = sizes[-1]
    yy, xx = np.mgrid[0:h_f, 0:w_f].astype(np.float64)
    img = np.zeros((h_f, w_f, 3))
    for c in range(3):
        for _ in range(8):
            img[:, :, c] += rs.uniform(0.3, 1) * np.sin(rs.uniform(.01, .2) * xx + rs.uniform(.01, .2) * yy + rs.uniform(0, 6))
    img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    full = Image.fromarray(img)
    levels = [full.resize(s, Image.LANCZOS) for s in sizes]
    for i, lv in enumerate(levels):
        d = folder / f"scale_{i}"
        d.mkdir(parents=True, exist_ok=True)
        lv.save(d / "img.png")
        if i > 0:
            d = folder / f"scale_{i}_recon"
            d.mkdir(parents=True, exist_ok=True)
            levels[i - 1].resize(sizes[i], Image.BILINEAR).save(d / "img.png")


def mean_px():
    return float(np.mean([w * h for (w, h) in BALLOONS_SIZES]))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference) on a bounded sample
# ------------------------------------------------------------------------------------------------

def cpu_train_sample(reps=1, batch=1):
    """One training step (forward + backward + Adam) per scale at `batch` images; returns seconds per scale."""
    from oracle import sinddm_oracle as orc
    params = {k: v.clone().requires_grad_(True) for k, v in orc.synthetic_params(0, DIM).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    sch = orc.Schedule(5, BALLOONS_SCALE_LOSSES, timesteps=100, train_full_t=True)
    g = torch.Generator().manual_seed(0)
    times = []
    for s, (w, h) in enumerate(BALLOONS_SIZES):
        x = torch.rand(batch, 3, h, w, generator=g) * 2 - 1
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            t = torch.randint(0, 100, (batch,), generator=g)
            noise = torch.randn(batch, 3, h, w, generator=g)
            loss = orc.p_losses(params, sch, x, t, s, noise, x_orig=x)
            loss.backward()
            opt.step()
            opt.zero_grad()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        times.append(best)
    return times


def cpu_baseline_object():
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_train_sample(reps=1)                                   # warm-up (thread pools, allocator)
    times = cpu_train_sample(reps=2)
    step32 = float(np.mean(times)) * BATCH                    # CPU time is linear in batch
    return {"value": 1.0 / step32, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle (torch-CPU restatement of the reference) train step at batch 1 of {BATCH} on each of the "
                      f"5 balloons scales, best of 2, time scaled x{BATCH}; per-scale s: "
                      + ",".join(f"{t:.3f}" for t in times)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, restated (oracle), all host threads."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import sinddm_oracle as orc
    params = {k: v.clone().requires_grad_(True) for k, v in orc.synthetic_params(0, DIM).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    sch = orc.Schedule(5, BALLOONS_SCALE_LOSSES, timesteps=100, train_full_t=True)
    g = torch.Generator().manual_seed(0)
    data = [torch.rand(1, 3, h, w, generator=g) * 2 - 1 for (w, h) in BALLOONS_SIZES]

    def step(i):
        s = i % 5
        t = torch.randint(0, 100, (1,), generator=g)
        noise = torch.randn(data[s].shape, generator=g)
        loss = orc.p_losses(params, sch, data[s], t, s, noise, x_orig=data[s])
        loss.backward()
        opt.step()
        opt.zero_grad()

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = args.steps / (dt * BATCH)      # batch-32 steps/s: each timed step processed 1 of 32 images
    sample = (f"each step = one oracle train step at batch 1 of {BATCH} (scale = step mod 5); value scales the time "
              f"x{BATCH} (CPU time is linear in batch)")
    print(json.dumps({
        "impl": "reference", "metric": "train_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * BATCH * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "balloons full pyramid train, batch 32, T=100, dim=160 (configs[1])",
                   "device": "cpu", "note": "the reference is pure PyTorch; its CPU path restated in oracle/"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

def run_b200_arm(args):
    import ctypes as C

    import torch.distributed as tdist

    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet, _capi
    from sinddm_b200 import dist as spdist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION): send
    # everything written to fd 1 during initialisation to stderr so that stdout carries the JSON line only
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local_rank, world = spdist.init_process_group()
        torch.cuda.set_device(local_rank)
        dev = f"cuda:{local_rank}"
        if world > 1:
            tdist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    lib = _capi.load()
    _capi.init(local_rank)
    torch.manual_seed(0)

    tmp = Path(tempfile.mkdtemp(prefix=f"sinddm_bench_r{rank}_"))
    synthetic_pyramid(tmp, BALLOONS_SIZES)
    net = SinDDMNet(dim=DIM, multiscale=True, device=dev).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=BALLOONS_SIZES,
                                      timesteps=100, train_full_t=True, scale_losses=BALLOONS_SCALE_LOSSES,
                                      loss_type="l1", reblurring=True, omega=0, device=dev,
                                      results_folder=str(tmp / "res")).to(dev)
    dif.num_timesteps_ideal = list(BALLOONS_T_IDEAL)
    trainer = MultiscaleTrainer(dif, str(tmp) + "/", n_scales=5, scale_factor=1.403, image_sizes=BALLOONS_SIZES,
                                train_batch_size=BATCH * world, train_lr=1e-3, train_num_steps=10 ** 9,
                                gradient_accumulate_every=1, ema_decay=0.995, fp16=False, save_and_sample_every=10 ** 9,
                                avg_window=10 ** 9, sched_milestones=[20000, 40000, 70000, 80000, 90000, 110000],
                                results_folder=str(tmp / "res"), device=dev)
    trainer._prepare_training()
    trainer.step = 1            # skip the step-0 loss print / read-back

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------------------
    for i in range(args.warmup):
        trainer.train_step(s=i % 5)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.sinddm_profile_enable(1)
    launches0 = lib.sinddm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_scale = [[] for _ in range(5)]
    ev0.record()
    for i in range(args.steps):
        trainer.train_step(s=i % 5)
    ev1.record()
    barrier()
    ms_total = reduce_max(ev0.elapsed_time(ev1))
    launches = int(lib.sinddm_launch_count() - launches0)
    clocks = sampler.summary()
    prof = {}
    for kind, name in ((0, "tc_conv_kernel"), (1, "tc_wgrad_kernel")):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        _capi.check(lib.sinddm_profile_collect(kind, C.byref(ms), C.byref(fl), C.byref(n)), "profile_collect")
        prof[name] = {"ms": ms.value, "flops": fl.value, "launches": n.value}
    lib.sinddm_profile_enable(0)

    # per-scale step time (separate short loops; not part of `value`)
    for s in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        trainer.train_step(s=s)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            trainer.train_step(s=s)
        e1.record()
        torch.cuda.synchronize()
        per_scale[s] = e0.elapsed_time(e1) / 3

    # ---- end to end through the public API, inputs resident on the HOST -------------------------------
    host_data = [tuple(t.cpu().pin_memory() for t in pair) for pair in trainer.data_list]
    h2d = float(np.mean([sum(t.numel() * 4 for t in pair) for pair in host_data]))

    def e2e_step(i):
        s = i % 5
        trainer.data_list[s] = tuple(t.to(dev, non_blocking=True) for t in host_data[s])
        loss = trainer.train_step(s=s)
        return loss.item()                    # device -> host read of the step's result

    for i in range(max(args.warmup, 5)):
        e2e_step(i)
    barrier()
    e2e_per_scale = [0.0] * 5
    t0 = time.perf_counter()
    e2e_steps = []
    for i in range(args.steps):
        ts = time.perf_counter()
        e2e_step(i)
        e2e_steps.append(time.perf_counter() - ts)
        e2e_per_scale[i % 5] += e2e_steps[-1]
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_per_scale = [1e3 * v / max(1, len(range(k, args.steps, 5))) for k, v in enumerate(e2e_per_scale)]

    # ---- sampling: sample_scales, 16 images per GPU ----------------------------------------------------
    def sample_once():
        return trainer.sample_scales(scale_mul=(1, 1), custom_sample=True, batch_size=SAMPLE_BATCH * world,
                                     custom_t_list=BALLOONS_T_IDEAL[1:], save_images=False)
    sample_once()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import sinddm_b200.diffusion as sdiff
    l0, g0 = lib.sinddm_launch_count(), sdiff.graph_replayed_launches
    e0.record()
    out = sample_once()
    e1.record()
    barrier()
    sample_ms = reduce_max(e0.elapsed_time(e1))
    # kernels launched directly + kernels executed by CUDA-graph replays of the captured timestep
    sample_launches = int(lib.sinddm_launch_count() - l0) + int(sdiff.graph_replayed_launches - g0)
    t0 = time.perf_counter()
    final = sample_once()[-1]
    host_imgs = final.cpu()                                   # images read back to the host
    torch.cuda.synchronize()
    sample_e2e_s = reduce_max(time.perf_counter() - t0)
    finite = bool(torch.isfinite(host_imgs).all())

    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    bf16_sustained = peaks.get("bf16_tflops_sustained")
    tf32_peak = (bf16_sustained / 2.0) if bf16_sustained else 1400.0 / 2.0
    peak_src = "measured" if bf16_sustained else "fallback"
    conv = prof["tc_conv_kernel"]
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else None
    wg = prof["tc_wgrad_kernel"]
    wg_ach = wg["flops"] / (wg["ms"] * 1e-3) / 1e12 if wg["ms"] > 0 else None
    traffic = None
    tpath = REPO / "profiles" / "traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get("tc_conv_kernel_dram_bytes_per_launch")

    steps_per_s = world * args.steps / (ms_total * 1e-3)
    result = {
        "metric": "train_steps_per_sec",
        "value": steps_per_s,
        "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "tf32",
        "data": "synthetic",
        "config": {"workload": "balloons full pyramid train, batch 32 per GPU, T=100, dim=160 (configs[1])",
                   "scales_hw": [(h, w) for (w, h) in BALLOONS_SIZES], "s_schedule": "round_robin over 5 scales",
                   "global_batch": BATCH * world, "parallelism": f"dp{world}",
                   "value_definition": "batch-32 optimizer steps/s summed over GPUs (global batch 32*N per step)",
                   "l2": "per-step working set (0.7-11 GB of activations) exceeds the 126 MB L2; no flush needed",
                   "math": "TF32 operands, fp32 accumulate (tcgen05 kind::tf32); storage fp32"},
        "clocks": clocks,
        "e2e": {"value": world * args.steps / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "per_scale_ms_per_step": per_scale,
        "e2e_per_scale_ms_per_step": e2e_per_scale,
        "e2e_slowest_step": {"index": int(np.argmax(e2e_steps)), "ms": 1e3 * float(np.max(e2e_steps))},
        "finest_scale_steps_per_sec": 1e3 / per_scale[4] * world,
        "achieved_tflops_whole_step": TRAIN_FLOP_PER_PX * mean_px() * BATCH * steps_per_s / 1e12,
        "roofline": {"bound": "tensor", "kernel": "tc_conv_kernel (3x3/1x1 conv forward + data gradient)",
                     "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (achieved / tf32_peak) if achieved else None, "traffic": traffic,
                     "peak_note": f"dense TF32 = half of the {peak_src} sustained bf16 cuBLAS rate",
                     "launches_timed": conv["launches"], "share_of_step": conv["ms"] / ms_total,
                     "wgrad_kernel": {"achieved": wg_ach, "frac": (wg_ach / tf32_peak) if wg_ach else None,
                                      "launches_timed": wg["launches"], "share_of_step": wg["ms"] / ms_total}},
        "sampling": {"metric": "sample_images_per_sec", "value": SAMPLE_BATCH * world / (sample_ms * 1e-3),
                     "unit": "images/s", "ms_per_image": sample_ms / (SAMPLE_BATCH * world),
                     "e2e_value": SAMPLE_BATCH * world / sample_e2e_s, "images": SAMPLE_BATCH * world,
                     "net_evals_per_image": sum(BALLOONS_T_IDEAL), "gpu_launches": sample_launches,
                     "achieved_tflops": FWD_FLOP_PER_PX * sum(t * w * h for t, (w, h) in zip(BALLOONS_T_IDEAL, BALLOONS_SIZES))
                     * SAMPLE_BATCH * world / (sample_ms * 1e-3) / 1e12,
                     "finite": finite,
                     "config": "sample_scales(batch 16 per GPU, scale_mul=(1,1), T list [100,52,41,31,22]) (configs[2])"},
    }
    if world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline_object()
    print(json.dumps(result))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
