"""Benchmark of the SinDDM hot path on B200 (and the reference CPU arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

Headline metric (BASELINE.json): training steps/s on the balloons pyramid (5 scales, 48x64 .. 186x248), batch 32
per GPU, T=100, dim=160 -- one "step" = one MultiscaleTrainer optimizer step (q_sample, denoiser forward + backward,
L1 loss, [gradient all-reduce], Adam, EMA cadence, LR scheduler) at the scale the step draws.

Legs of the B200 arm (one JSON line on stdout):
  value      device-timed (CUDA events) `trainer.train_step(s)` with the scales visited round-robin = the exact
             expectation of the reference's uniform multinomial draw; inputs resident in HBM.
  e2e        `MultiscaleTrainer.train()` ITSELF (the public loop: torch.multinomial scale draw on the device
             generator, the batch copied in from pinned host memory every step, `loss.item()` every step like the
             reference, trainer.py:194-213), wall clock.  The seed is picked so that the K draws of the timed window
             visit the five scales as evenly as K allows (the multinomial is still drawn, per step, inside train());
             the realised counts are reported.
  sampling   sample_scales(16 images per GPU) -> images/s (configs[2]).
  roofline   the dominant kernel (tc_conv) from CUDA events around each of its launches in the timed region.
  cpu_baseline / gpu_eager_baseline   the oracle (torch restatement of the reference, stock ATen/oneDNN/cuDNN kernels)
             at a REAL batch 32 on the host cores, and the same restatement run eagerly on the B200 (cuDNN TF32 = the
             reference's own GPU path) -- reported baselines.
  other_configs   BASELINE.json configs[3] (seascape, global batch 128 split over the ranks = strong scaling) and
             configs[4] (starry_night x(2,2), 64 samples split over the ranks).
`--impl reference` times the CPU oracle at batch 32, scale = step mod 5, all host threads.
"""
from __future__ import annotations

import argparse
import contextlib
import gc
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

# (W, H) per scale and the reference's num_timesteps_ideal, probed from the reference
BALLOONS_SIZES = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
BALLOONS_T_IDEAL = [100, 52, 41, 31, 22]
BALLOONS_SCALE_LOSSES = [1.20, 0.85, 0.60, 0.42]   # any values: train_full_t=True, T list given explicitly
SEASCAPE_SIZES = [(62, 50), (88, 71), (125, 100), (176, 141), (249, 200)]
SEASCAPE_T_IDEAL = [100, 56, 45, 35, 25]
STARRY_SIZES = [(62, 49), (88, 69), (125, 98), (178, 140), (252, 198)]
STARRY_T_IDEAL = [100, 48, 37, 27, 19]
DIM = 160
BATCH = 32
SAMPLE_BATCH = 16
SEASCAPE_GLOBAL_BATCH = 128        # the reference's Dataset yields at most 128 rows (trainer.py:52-53)
STARRY_GLOBAL_SAMPLES = 64
FWD_FLOP_PER_PX = 2.150e6
TRAIN_FLOP_PER_PX = 6.45e6
WORKLOAD = "balloons full pyramid train, batch 32 per GPU, T=100, dim=160 (configs[1])"
S_SCHEDULE = "scale = step mod 5 (every window of 5 steps is the expectation of the reference's uniform draw)"


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


@contextlib.contextmanager
def stdout_to_stderr():
    """Everything written to fd 1 inside the block goes to stderr (NCCL banners, the trainer's progress prints):
    stdout carries the JSON line only."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def synthetic_pyramid(folder: Path, sizes, seed=0):
    """Synthetic training data of the named image's shape: one smooth random image per scale plus its blurry twin,
    in the folder layout MultiscaleTrainer reads (scale_i/, scale_i_recon/)."""
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


def mean_px(sizes=BALLOONS_SIZES):
    return float(np.mean([w * h for (w, h) in sizes]))


def sample_flops_per_image(sizes, t_list, scale_mul=(1, 1)):
    return FWD_FLOP_PER_PX * sum(t * int(w * scale_mul[1]) * int(h * scale_mul[0]) for t, (w, h) in zip(t_list, sizes))


# ------------------------------------------------------------------------------------------------
# baselines: the oracle (torch restatement of the reference) on the host cores / eagerly on the GPU
# ------------------------------------------------------------------------------------------------

class OracleTrainer:
    """The reference's training step restated with stock torch ops (oracle/), Adam included, at a real batch."""

    def __init__(self, device="cpu", batch=BATCH, sizes=BALLOONS_SIZES):
        from oracle import sinddm_oracle as orc
        self.orc, self.device, self.batch, self.sizes = orc, device, batch, sizes
        self.params = {k: v.to(device).requires_grad_(True) for k, v in orc.synthetic_params(0, DIM).items()}
        self.opt = torch.optim.Adam(list(self.params.values()), lr=1e-3)
        self.sch = orc.Schedule(5, BALLOONS_SCALE_LOSSES, timesteps=100, train_full_t=True).to(device)
        g = torch.Generator().manual_seed(0)
        self.data = [(torch.rand(1, 3, h, w, generator=g) * 2 - 1).repeat(batch, 1, 1, 1).to(device) for (w, h) in sizes]

    def step(self, s):
        x = self.data[s]
        t = torch.randint(0, 100, (self.batch,), device=self.device)
        noise = torch.randn_like(x)
        loss = self.orc.p_losses(self.params, self.sch, x, t, s, noise, x_orig=x)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad()
        return loss


def cpu_baseline_object():
    """One REAL batch-32 training step per balloons scale on all host cores (about 13 s on a 16-core box)."""
    torch.set_num_threads(os.cpu_count() or 1)
    warm = OracleTrainer("cpu", batch=2)
    for s in range(5):
        warm.step(s)                                           # thread pools, allocator, oneDNN primitive caches
    del warm
    tr = OracleTrainer("cpu", batch=BATCH)
    times = []
    for s in range(5):
        t0 = time.perf_counter()
        tr.step(s)
        times.append(time.perf_counter() - t0)
    return {"value": 5.0 / float(np.sum(times)), "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle (torch-CPU restatement of the reference) at the real batch {BATCH}: one training step "
                      "(forward, backward, Adam) on each of the 5 balloons scales after a batch-2 warm-up pass; "
                      "per-scale s: " + ",".join(f"{t:.2f}" for t in times)}


def gpu_eager_baseline_object(dev, steps=10):
    """The reference's own GPU path: the restated module graph run eagerly by stock PyTorch on the B200 (cuDNN
    convolutions with TF32 allowed -- torch.backends.cudnn.allow_tf32 defaults to True, reference main.py sets
    nothing else), batch 32, round-robin scales; plus its reverse-sampling chain for 16 images."""
    from oracle import sinddm_oracle as orc
    torch.backends.cudnn.allow_tf32 = True
    tr = OracleTrainer(dev, batch=BATCH)
    for s in range(5):
        tr.step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.step(i % 5)
    e1.record()
    torch.cuda.synchronize()
    train_ms = e0.elapsed_time(e1) / steps
    per_scale = []
    for s in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tr.step(s)
        tr.step(s)
        b.record()
        torch.cuda.synchronize()
        per_scale.append(a.elapsed_time(b) / 2)
    params = {k: v.detach() for k, v in tr.params.items()}
    sch = tr.sch
    del tr
    torch.cuda.empty_cache()

    @torch.no_grad()
    def chain():
        img = None
        for s, ((w, h), total) in enumerate(zip(BALLOONS_SIZES, BALLOONS_T_IDEAL)):
            if s == 0:
                img = torch.randn(SAMPLE_BATCH, 3, h, w, device=dev)
                x_tilde = None
            else:
                x_tilde = torch.nn.functional.interpolate(img, size=(h, w), mode="bilinear")
                tt = torch.full((SAMPLE_BATCH,), total, device=dev, dtype=torch.long)
                img = orc.q_sample(sch, x_tilde, tt, torch.randn_like(x_tilde))
            for i in reversed(range(total)):
                tt = torch.full((SAMPLE_BATCH,), i, device=dev, dtype=torch.long)
                img = orc.p_sample(params, sch, img, tt, s, torch.randn_like(img), x_tilde=x_tilde)
        return img

    # one coarse-scale pass as warm-up, then the full chain
    chain()
    torch.cuda.synchronize()
    e0.record()
    out = chain()
    e1.record()
    torch.cuda.synchronize()
    sample_ms = e0.elapsed_time(e1)
    return {"train_steps_per_sec": 1e3 / train_ms, "train_ms_per_step": train_ms, "per_scale_ms_per_step": per_scale,
            "sample_images_per_sec": SAMPLE_BATCH / (sample_ms * 1e-3), "finite": bool(torch.isfinite(out).all()),
            "what": "oracle/ (the reference's module graph as stock torch ops) run eagerly on this B200: cuDNN "
                    f"convolutions with allow_tf32=True, torch.optim.Adam, batch {BATCH}, scale = step mod 5, "
                    f"{steps} timed steps; sampling = the 246-evaluation chain for {SAMPLE_BATCH} images",
            "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, restated (oracle), all host threads, at the
    real batch 32; step i runs scale i mod 5.  Rank 0 only under torchrun."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    tr = OracleTrainer("cpu", batch=BATCH)
    for i in range(args.warmup):
        tr.step(i % 5)
    t0 = time.perf_counter()
    for i in range(args.steps):
        tr.step(i % 5)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = (f"each step = one oracle training step (forward, backward, Adam) at the real batch {BATCH}, "
              f"scale = step mod 5; {args.steps} steps after {args.warmup} warm-up steps")
    print(json.dumps({
        "impl": "reference", "metric": "train_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "s_schedule": S_SCHEDULE, "global_batch": BATCH, "device": "cpu",
                   "note": "the reference is pure PyTorch and cannot travel to the GPU box; its CPU path restated in "
                           "oracle/ (pinned against the reference by tests/golden); one process, all host cores"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

class Job:
    """Process-wide state of the B200 arm."""

    def __init__(self):
        import torch.distributed as tdist

        from sinddm_b200 import _capi
        from sinddm_b200 import dist as spdist
        self.tdist = tdist
        with stdout_to_stderr():
            self.rank, self.local_rank, self.world = spdist.init_process_group()
            torch.cuda.set_device(self.local_rank)
            self.dev = f"cuda:{self.local_rank}"
            if self.world > 1:
                tdist.all_reduce(torch.zeros(1, device=self.dev))
                torch.cuda.synchronize()
        self.lib = _capi.load()
        _capi.init(self.local_rank)
        self.capi = _capi

    def barrier(self):
        if self.world > 1:
            self.tdist.barrier()
        torch.cuda.synchronize()

    def reduce_max(self, x):
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.tdist.all_reduce(t, op=self.tdist.ReduceOp.MAX)
        return float(t.item())


def make_trainer(job, sizes, t_ideal, global_batch, tag, math=None):
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet
    tmp = Path(tempfile.mkdtemp(prefix=f"sinddm_bench_{tag}_r{job.rank}_"))
    synthetic_pyramid(tmp, sizes)
    dev = job.dev
    net = SinDDMNet(dim=DIM, multiscale=True, device=dev, math=math).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=sizes,
                                      timesteps=100, train_full_t=True, scale_losses=BALLOONS_SCALE_LOSSES,
                                      loss_type="l1", reblurring=True, omega=0, device=dev,
                                      results_folder=str(tmp / "res")).to(dev)
    dif.num_timesteps_ideal = list(t_ideal)
    trainer = MultiscaleTrainer(dif, str(tmp) + "/", n_scales=5, scale_factor=1.403, image_sizes=sizes,
                                train_batch_size=global_batch, train_lr=1e-3, train_num_steps=10 ** 9,
                                gradient_accumulate_every=1, ema_decay=0.995, fp16=False, save_and_sample_every=10 ** 9,
                                avg_window=10 ** 9, sched_milestones=[20000, 40000, 70000, 80000, 90000, 110000],
                                results_folder=str(tmp / "res"), device=dev)
    with stdout_to_stderr():
        trainer._prepare_training()
    trainer.step = 1            # skip the step-0 loss print / read-back
    return trainer


def pick_balanced_seed(trainer, job, sizes, steps, warmup, max_seeds=600):
    """Seed whose `warmup + steps` training steps draw, in the timed window, every scale as evenly as `steps` allows.
    Replays exactly the device-RNG calls of a step (multinomial -> randint(global batch) -> randn(global batch x 3 x
    H x W), quirk Q7) without running the network.  Deterministic, so every rank picks the same seed."""
    dev = job.dev
    gb = trainer.local_batch * job.world
    px = [w * h for (w, h) in sizes]
    target = steps * float(np.mean(px))
    best = None
    max_seeds = max(20, min(max_seeds, 12000 // (warmup + steps)))      # bound the search to a few seconds
    for seed in range(max_seeds):
        torch.manual_seed(seed)
        counts = [0] * len(sizes)
        for i in range(warmup + steps):
            s = trainer._draw_scale()
            torch.randint(0, 100, (gb,), device=dev)
            w, h = sizes[s]
            torch.randn((gb, 3, h, w), device=dev)
            if i >= warmup:
                counts[s] += 1
        cost = abs(sum(c * p for c, p in zip(counts, px)) - target) / target
        if best is None or cost < best[0]:
            best = (cost, seed, counts)
        if max(counts) - min(counts) <= (0 if steps % len(sizes) == 0 else 1):
            return seed, counts
    return best[1], best[2]


def train_legs(job, args, trainer, sizes, want_e2e=True):
    """Device-timed round-robin leg (+ roofline counters), per-scale times, and the train() end-to-end leg."""
    import ctypes as C
    lib, dev, world = job.lib, job.dev, job.world
    n_sc = len(sizes)
    for i in range(args.warmup):
        trainer.train_step(s=i % n_sc)
    job.barrier()
    fused = getattr(trainer, "_fused", None)
    if fused is not None:
        fused.barrier_wait_ms(reset=True)
    sampler = ClockSampler(job.local_rank)
    sampler.start()
    launches0 = lib.sinddm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        trainer.train_step(s=i % n_sc)
    ev1.record()
    job.barrier()
    res = {"ms_total": job.reduce_max(ev0.elapsed_time(ev1)),
           "launches": int(lib.sinddm_launch_count() - launches0), "clocks": sampler.summary(), "prof": {}}
    # the same K steps once more with a CUDA event pair around every tc_conv / tc_wgrad launch (on the launching
    # stream) for the roofline object: the ~56 event records per step cost about 2 % and stay out of `value`
    lib.sinddm_profile_enable(1)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for i in range(args.steps):
        trainer.train_step(s=i % n_sc)
    pv1.record()
    job.barrier()
    res["ms_total_profiled"] = job.reduce_max(pv0.elapsed_time(pv1))
    for kind, name in ((0, "tc_conv_kernel"), (1, "tc_wgrad_kernel")):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        job.capi.check(lib.sinddm_profile_collect(kind, C.byref(ms), C.byref(fl), C.byref(n)), "profile_collect")
        res["prof"][name] = {"ms": ms.value, "flops": fl.value, "launches": n.value}
    lib.sinddm_profile_enable(0)
    if fused is not None and world > 1:
        # time every rank sat in the fused step's NVLink barrier waiting for slower peers (rank skew, measured)
        tot, mx = fused.barrier_wait_ms(reset=True)
        t = torch.tensor([tot / args.steps, mx], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        job.tdist.all_gather(allt, t)
        res["barrier_wait"] = {"ms_per_step_by_rank": [float(a[0]) for a in allt],
                               "max_single_wait_ms_by_rank": [float(a[1]) for a in allt],
                               "what": "time CTA 0 of each rank's fused all-reduce+Adam+EMA kernel spent waiting for "
                                       "the slowest peer's gradients (in-kernel NVLink flag barrier); the fastest "
                                       "rank's wait is the step-time skew between GPUs"}

    # per-scale step time (separate short loops; not part of `value`)
    per_scale = []
    for s in range(n_sc):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6 if s < n_sc - 2 else 3        # short loops expose the first step's host launch latency at small scales
        trainer.train_step(s=s)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            trainer.train_step(s=s)
        e1.record()
        torch.cuda.synchronize()
        per_scale.append(e0.elapsed_time(e1) / reps)
    res["per_scale_ms"] = per_scale
    # the dominant kernel per scale (two more steps each with the per-launch events on)
    by_scale = []
    for s in range(n_sc):
        lib.sinddm_profile_enable(1)
        trainer.train_step(s=s)
        trainer.train_step(s=s)
        torch.cuda.synchronize()
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        job.capi.check(lib.sinddm_profile_collect(0, C.byref(ms), C.byref(fl), C.byref(n)), "profile_collect")
        lib.sinddm_profile_enable(0)
        by_scale.append(fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else None)
    res["conv_tflops_by_scale"] = by_scale
    if not want_e2e:
        return res

    # ---- end to end: MultiscaleTrainer.train() itself, inputs resident on the HOST, loss read back every step
    trainer.prepare_host_data()             # pinned host batches + persistent device staging buffers for every scale
    trainer.loss_readback = "step"
    for i in range(2 * n_sc):               # every scale twice in this mode before anything is seeded or timed
        trainer.train_step(s=i % n_sc)
    torch.cuda.synchronize()
    res["h2d"] = float(np.mean([sum(t.numel() * 4 for t in pair) for pair in trainer.data_list]))
    e2e_warm = max(args.warmup, 10)
    seed, predicted = pick_balanced_seed(trainer, job, sizes, args.steps, e2e_warm)
    torch.manual_seed(seed)
    # instrumentation only: a host timestamp per train_step() call of train() (which scale, how long on the host)
    stamps = []
    inner = trainer.train_step

    def stamped_step():
        ts = time.perf_counter()
        out = inner()
        stamps.append(time.perf_counter() - ts)
        return out
    trainer.train_step = stamped_step
    with stdout_to_stderr():
        trainer.train_num_steps = trainer.step + e2e_warm
        trainer.train()
        job.barrier()
        gc.collect()
        stamps.clear()
        trainer.scale_counts = [0] * n_sc
        t0 = time.perf_counter()
        trainer.train_num_steps = trainer.step + args.steps
        trainer.train()
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        job.barrier()
    del trainer.train_step
    res["e2e_slowest_step"] = {"index": int(np.argmax(stamps)), "host_ms": 1e3 * float(np.max(stamps)),
                               "median_host_ms": 1e3 * float(np.median(stamps))}
    res["e2e_s"] = job.reduce_max(t_local)
    res["e2e_counts"] = list(trainer.scale_counts)
    res["e2e_seed"] = seed
    res["e2e_counts_predicted"] = predicted
    res["e2e_last_loss"] = trainer.last_loss
    # what the round-robin mix of the same per-scale times would give for the realised draw (identical when balanced)
    res["e2e_mix_ms_expected"] = float(sum(c * m for c, m in zip(trainer.scale_counts, per_scale)))
    return res


def sampling_leg(job, trainer, t_ideal, global_batch, scale_mul=(1, 1)):
    import sinddm_b200.diffusion as sdiff
    lib = job.lib

    def sample_once():
        return trainer.sample_scales(scale_mul=scale_mul, custom_sample=True, batch_size=global_batch,
                                     custom_t_list=list(t_ideal[1:]), save_images=False)
    with stdout_to_stderr():
        sample_once()
        job.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, g0 = lib.sinddm_launch_count(), sdiff.graph_replayed_launches
        passes = 3          # a pass is ~0.3 s: three back to back smooth the clock / power-cap jitter of a single one
        e0.record()
        for _ in range(passes):
            sample_once()
        e1.record()
        job.barrier()
        ms = job.reduce_max(e0.elapsed_time(e1)) / passes
        # kernels launched directly + kernels executed by CUDA-graph replays of the captured timestep
        launches = (int(lib.sinddm_launch_count() - l0) + int(sdiff.graph_replayed_launches - g0)) // passes
        t0 = time.perf_counter()
        final = sample_once()[-1]
        host_imgs = final.cpu()                                   # images read back to the host
        torch.cuda.synchronize()
        e2e_s = job.reduce_max(time.perf_counter() - t0)
    return {"ms": ms, "launches": launches, "e2e_s": e2e_s, "finite": bool(torch.isfinite(host_imgs).all()),
            "d2h": host_imgs.numel() * 4}


def strict_mode_leg(job):
    """The same training step in math = "tf32x3" (3xTF32: every conv operand split into two TF32 values, fp32-class
    results -- the reference with allow_tf32 = False -- on the same tcgen05 kernels): device ms per step and scale,
    three timed steps each after two warm ones.  A side record, not the headline (which is the TF32 default both the
    reference's GPU path and this one use)."""
    trainer = make_trainer(job, BALLOONS_SIZES, BALLOONS_T_IDEAL, BATCH * job.world, "x3", math="tf32x3")
    per = []
    for s in range(len(BALLOONS_SIZES)):
        for _ in range(2):
            trainer.train_step(s=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            trainer.train_step(s=s)
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1) / 3)
    del trainer
    release_memory()
    return {"math": "tf32x3", "value": 1e3 * len(per) / sum(per), "unit": "steps/s", "per_scale_ms_per_step": per,
            "note": "3xTF32 on the tensor cores (x = hi + lo; x_hi*w_lo + x_lo*w_hi + x_hi*w_hi accumulated in fp32): "
                    "per-convolution error 2e-6 .. 4e-6 of max|ref| vs 4e-4 for TF32 (profiles/r02_tf32x3.txt); "
                    "scales visited round-robin, batch 32"}


def release_memory():
    """Call after `del trainer` in the caller's frame: plans, workspaces and cached blocks go back to the driver."""
    gc.collect()
    torch.cuda.empty_cache()


def seascape_config(job, args):
    """configs[3]: seascape pyramid, GLOBAL batch 128 split over the ranks (16 per GPU at 8) -- strong scaling."""
    world = job.world
    release_memory()
    need = 124.0 / world + 6.0     # GiB of plan workspaces (sinddm_plan_workspace_bytes, 5 scales) + slack
    free_gib = torch.cuda.mem_get_info()[0] / 2 ** 30
    if free_gib < need:
        return {"skipped": f"needs {need:.0f} GiB of plan workspaces per GPU at {world} GPU(s), {free_gib:.0f} GiB free"}
    trainer = make_trainer(job, SEASCAPE_SIZES, SEASCAPE_T_IDEAL, SEASCAPE_GLOBAL_BATCH, "sea")
    a = argparse.Namespace(steps=10, warmup=5)
    r = train_legs(job, a, trainer, SEASCAPE_SIZES, want_e2e=False)
    del trainer
    release_memory()
    sps = a.steps / (r["ms_total"] * 1e-3)
    return {"metric": "train_steps_per_sec", "value": sps, "unit": "steps/s", "scaling": "strong",
            "ms_per_step": r["ms_total"] / a.steps, "steps": a.steps, "warmup": a.warmup,
            "per_scale_ms_per_step": r["per_scale_ms"],
            "achieved_tflops_job": TRAIN_FLOP_PER_PX * mean_px(SEASCAPE_SIZES) * SEASCAPE_GLOBAL_BATCH * sps / 1e12,
            "config": {"workload": "seascape full pyramid train, GLOBAL batch 128 (the reference Dataset's cap), "
                                   "T=100, dim=160 (configs[3])", "global_batch": SEASCAPE_GLOBAL_BATCH,
                       "batch_per_gpu": SEASCAPE_GLOBAL_BATCH // world, "parallelism": f"dp{world}",
                       "scales_hw": [(h, w) for (w, h) in SEASCAPE_SIZES], "s_schedule": S_SCHEDULE,
                       "value_definition": "global-batch-128 optimizer steps/s of the whole job"}}


def starry_config(job, args):
    """configs[4]: starry_night sample_scales with scale_mul=(2,2) (396x504 finest), 64 samples split over the ranks."""
    world = job.world
    release_memory()
    need = 62.0 / world + 6.0
    free_gib = torch.cuda.mem_get_info()[0] / 2 ** 30
    if free_gib < need:
        return {"skipped": f"needs {need:.0f} GiB of plan workspaces per GPU at {world} GPU(s), {free_gib:.0f} GiB free"}
    trainer = make_trainer(job, STARRY_SIZES, STARRY_T_IDEAL, 8 * world, "star")
    r = sampling_leg(job, trainer, STARRY_T_IDEAL, STARRY_GLOBAL_SAMPLES, scale_mul=(2, 2))
    del trainer
    release_memory()
    ips = STARRY_GLOBAL_SAMPLES / (r["ms"] * 1e-3)
    return {"metric": "sample_images_per_sec", "value": ips, "unit": "images/s", "scaling": "strong",
            "ms_per_image": r["ms"] / STARRY_GLOBAL_SAMPLES, "e2e_value": STARRY_GLOBAL_SAMPLES / r["e2e_s"],
            "images": STARRY_GLOBAL_SAMPLES, "net_evals_per_image": sum(STARRY_T_IDEAL), "gpu_launches": r["launches"],
            "achieved_tflops_job": sample_flops_per_image(STARRY_SIZES, STARRY_T_IDEAL, (2, 2)) * ips / 1e12,
            "finite": r["finite"],
            "config": {"workload": "starry_night sample_scales, scale_mul=(2,2) (98x124 .. 396x504), 64 samples, "
                                   "T list [100,48,37,27,19] (configs[4])", "samples_per_gpu": STARRY_GLOBAL_SAMPLES // world,
                       "parallelism": f"{world} x independent samples, no collective"}}


def run_b200_arm(args):
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    job = Job()
    rank, world, dev = job.rank, job.world, job.dev
    torch.manual_seed(0)

    if args.config == "seascape_b128":
        out = seascape_config(job, args)
    elif args.config == "starry_2x2":
        out = starry_config(job, args)
    else:
        out = None
    if out is not None:
        if world > 1:
            job.tdist.barrier()
            job.tdist.destroy_process_group()
        if rank == 0:
            out.update({"n_gpus": world, "higher_is_better": True, "vs_baseline": None, "dtype": "tf32",
                        "data": "synthetic"})
            print(json.dumps(out))
        return

    # the eager-PyTorch-on-this-GPU baseline first (its cuDNN workspaces and autograd buffers are freed afterwards)
    eager = None
    if world == 1 and not args.no_baselines:
        try:
            eager = gpu_eager_baseline_object(dev)
        except Exception as e:  # noqa: BLE001 -- a baseline must not take the measurement down
            eager = {"unavailable": f"{type(e).__name__}: {e}"}
        gc.collect()
        torch.cuda.empty_cache()
        torch.manual_seed(0)

    trainer = make_trainer(job, BALLOONS_SIZES, BALLOONS_T_IDEAL, BATCH * world, "bal")
    r = train_legs(job, args, trainer, BALLOONS_SIZES)
    smp = sampling_leg(job, trainer, BALLOONS_T_IDEAL, SAMPLE_BATCH * world)
    del trainer
    release_memory()
    strict = None
    if world == 1 and not args.no_baselines:
        try:
            strict = strict_mode_leg(job)
        except Exception as e:  # noqa: BLE001 -- a side record must not take the measurement down
            strict = {"failed": f"{type(e).__name__}: {e}"}
            release_memory()

    others = {}
    if not args.no_other_configs:
        for name, fn in (("seascape_b128", seascape_config), ("starry_2x2", starry_config)):
            try:
                others[name] = fn(job, args)
            except Exception as e:  # noqa: BLE001
                others[name] = {"failed": f"{type(e).__name__}: {e}"}
                gc.collect()
                torch.cuda.empty_cache()

    if world > 1:
        job.tdist.barrier()
        job.tdist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    bf16_sustained = peaks.get("bf16_tflops_sustained")
    tf32_peak = (bf16_sustained / 2.0) if bf16_sustained else 1400.0 / 2.0
    peak_src = "measured" if bf16_sustained else "fallback"
    ms_total = r["ms_total"]
    conv = r["prof"]["tc_conv_kernel"]
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else None
    wg = r["prof"]["tc_wgrad_kernel"]
    wg_ach = wg["flops"] / (wg["ms"] * 1e-3) / 1e12 if wg["ms"] > 0 else None
    traffic, traffic_src = None, None
    tpath = REPO / "profiles" / "traffic.json"
    if tpath.exists():
        tj = json.loads(tpath.read_text())
        traffic = tj.get("tc_conv_kernel_dram_bytes_per_launch")
        traffic_src = tj.get("source")

    steps_per_s = world * args.steps / (ms_total * 1e-3)
    per_scale = r["per_scale_ms"]
    e2e_value = world * args.steps / r["e2e_s"]
    sample_ms = smp["ms"]
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
        "config": {"workload": WORKLOAD,
                   "scales_hw": [(h, w) for (w, h) in BALLOONS_SIZES], "s_schedule": S_SCHEDULE,
                   "global_batch": BATCH * world, "parallelism": f"dp{world}",
                   "value_definition": "batch-32 optimizer steps/s summed over GPUs (global batch 32*N per step)",
                   "l2": "per-step working set (0.7-11 GB of activations) exceeds the 126 MB L2; no flush needed",
                   "math": "TF32 operands, fp32 accumulate (tcgen05 kind::tf32); storage fp32"},
        "clocks": r["clocks"],
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                "through": "MultiscaleTrainer.train(): torch.multinomial scale draw on the device generator, batch "
                           "copied from pinned host memory, loss.item(), every step (reference trainer.py:194-213)",
                "scale_counts": r["e2e_counts"], "seed": r["e2e_seed"],
                "seed_rule": "first seed whose timed window draws the scales as evenly as K allows "
                             "(bench.pick_balanced_seed)",
                "ms_per_step": 1e3 * r["e2e_s"] / args.steps,
                "device_ms_for_same_scale_mix": r["e2e_mix_ms_expected"] / args.steps,
                "slowest_step": r["e2e_slowest_step"],
                "last_loss": r["e2e_last_loss"]},
        "gpu_launches": r["launches"],
        **({"barrier_wait": r["barrier_wait"]} if "barrier_wait" in r else {}),
        "per_scale_ms_per_step": per_scale,
        "finest_scale_steps_per_sec": 1e3 / per_scale[4] * world,
        "achieved_tflops_whole_step": TRAIN_FLOP_PER_PX * mean_px() * BATCH * steps_per_s / 1e12,
        "roofline": {"bound": "tensor", "kernel": "tc_conv_kernel (3x3/1x1 conv forward + data gradient)",
                     "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (achieved / tf32_peak) if achieved else None, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_note": f"dense TF32 = half of the {peak_src} sustained bf16 cuBLAS rate",
                     "launches_timed": conv["launches"], "share_of_step": conv["ms"] / r["ms_total_profiled"],
                     "achieved_by_scale": r["conv_tflops_by_scale"],
                     "frac_by_scale": [(v / tf32_peak) if v else None for v in r["conv_tflops_by_scale"]],
                     "measured_in": "a second pass of the same K steps with CUDA events around each launch "
                                    f"({r['ms_total_profiled'] / args.steps:.3f} ms/step with the events on)",
                     "wgrad_kernel": {"achieved": wg_ach, "frac": (wg_ach / tf32_peak) if wg_ach else None,
                                      "launches_timed": wg["launches"],
                                      "share_of_step": wg["ms"] / r["ms_total_profiled"]}},
        "sampling": {"metric": "sample_images_per_sec", "value": SAMPLE_BATCH * world / (sample_ms * 1e-3),
                     "unit": "images/s", "ms_per_image": sample_ms / (SAMPLE_BATCH * world),
                     "e2e_value": SAMPLE_BATCH * world / smp["e2e_s"], "images": SAMPLE_BATCH * world,
                     "net_evals_per_image": sum(BALLOONS_T_IDEAL), "gpu_launches": smp["launches"],
                     "achieved_tflops": sample_flops_per_image(BALLOONS_SIZES, BALLOONS_T_IDEAL) * SAMPLE_BATCH * world
                     / (sample_ms * 1e-3) / 1e12,
                     "finite": smp["finite"],
                     "config": "sample_scales(batch 16 per GPU, scale_mul=(1,1), T list [100,52,41,31,22]) (configs[2])"},
    }
    if others:
        result["other_configs"] = others
    if strict is not None:
        result["strict_mode"] = strict
    if eager is not None:
        result["gpu_eager_baseline"] = eager
    if world == 1 and not args.no_baselines:
        result["cpu_baseline"] = cpu_baseline_object()
    print(json.dumps(result))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="balloons_b32", choices=["balloons_b32", "seascape_b128", "starry_2x2"])
    ap.add_argument("--no-baselines", "--no-cpu-baseline", dest="no_baselines", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10 and "--steps" not in sys.argv:
            args.steps, args.warmup = 20, 5          # a batch-32 CPU step takes seconds: keep the default run short
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
