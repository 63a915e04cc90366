"""The fused optimizer step (sinddm_fused_step: gradient mean over NVLink peer memory + Adam + EMA in one kernel,
replacing SinDDM/trainer.py:208-213) against torch.optim.Adam + the reference EMA on the same gradients, the
trainer with and without it, and -- when the box has two GPUs -- the two-rank data-parallel path against NCCL."""
import copy
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import GOLDEN_SCALE_LOSSES, GOLDEN_SIZES

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]
DEV = "cuda:0"


def _nets(seed=0):
    from oracle import sinddm_oracle as orc
    from sinddm_b200 import SinDDMNet
    net = SinDDMNet(dim=160, multiscale=True, device=DEV)
    net.load_state_dict(orc.synthetic_params(seed=seed, dim=160), strict=True)
    net.to(DEV)
    return net, copy.deepcopy(net)


def test_fused_step_matches_torch_adam_and_reference_ema():
    from sinddm_b200.diffusion import EMA
    from sinddm_b200.fused_optim import FusedStep
    net, ema_net = _nets()
    ref_net, ref_ema = copy.deepcopy(net), copy.deepcopy(ema_net)
    keys_before = list(net.state_dict().keys())
    fused = FusedStep(net, ema_net, betas=(0.9, 0.999), eps=1e-8)
    assert list(net.state_dict().keys()) == keys_before            # re-pointing keeps the module surface
    for a, b in zip(net.parameters(), ref_net.parameters()):
        assert torch.equal(a, b)
    opt = torch.optim.Adam(ref_net.parameters(), lr=1e-3)
    ema = EMA(0.995)
    g = torch.Generator(device=DEV).manual_seed(1)
    lrs = [1e-3, 1e-3, 5e-4, 5e-4, 2.5e-4, 2.5e-4]
    for it, lr in enumerate(lrs):
        mode = (1 if it < 2 else 2) if it % 2 == 0 else 0          # copy, -, lerp, -, lerp, -
        bucket = fused.bucket()
        bucket.copy_(torch.randn(bucket.numel(), device=DEV, generator=g) * 1e-2)
        for p, gv in zip(ref_net.parameters(), fused.grad_views()):
            p.grad = gv.clone()
        fused.step(lr, mode, 0.995)
        for grp in opt.param_groups:
            grp["lr"] = lr
        opt.step()
        if mode == 1:
            ref_ema.load_state_dict(ref_net.state_dict())
        elif mode == 2:
            ema.update_model_average(ref_ema, ref_net)
    torch.cuda.synchronize()
    for (name, a), b in zip(net.named_parameters(), ref_net.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), (name, float((a - b).abs().max()))
    for (name, a), b in zip(ema_net.named_parameters(), ref_ema.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), ("ema " + name, float((a - b).abs().max()))


def _trainer(tmp, fused: bool):
    import bench
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet
    from oracle import sinddm_oracle as orc
    sizes = GOLDEN_SIZES[:3]
    data = Path(tmp) / "data"
    bench.synthetic_pyramid(data, sizes)
    net = SinDDMNet(dim=160, multiscale=True, device=DEV)
    net.load_state_dict(orc.synthetic_params(seed=4, dim=160), strict=True)
    net.to(DEV)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=3, scale_factor=1.36, image_sizes=sizes, timesteps=100,
                                      train_full_t=True, scale_losses=GOLDEN_SCALE_LOSSES[:2], device=DEV,
                                      results_folder=str(Path(tmp) / "r")).to(DEV)
    os.environ["SINDDM_FUSED_STEP"] = "1" if fused else "0"
    try:
        tr = MultiscaleTrainer(dif, str(data) + "/", n_scales=3, image_sizes=sizes, train_batch_size=4, train_lr=1e-3,
                               gradient_accumulate_every=1, step_start_ema=4, update_ema_every=2, avg_window=10 ** 9,
                               sched_milestones=[5], results_folder=str(Path(tmp) / "r"), device=DEV)
        tr._prepare_training()
    finally:
        os.environ.pop("SINDDM_FUSED_STEP", None)
    assert (tr._fused is not None) == fused
    return tr


def test_trainer_with_fused_step_tracks_the_torch_adam_trainer(tmp_path):
    """Same seeds, same kernels for the gradients: the two optimizer implementations must stay together (the
    only difference is Adam's rounding), through EMA copy / EMA average steps and an LR milestone."""
    losses = {}
    finals = {}
    for fused in (False, True):
        tr = _trainer(tmp_path / str(fused), fused)
        torch.manual_seed(7)
        ls = []
        for i in range(8):
            ls.append(float(tr.train_step(s=i % 3)))
        losses[fused] = ls
        finals[fused] = ([p.detach().clone() for p in tr.model.parameters()],
                         [p.detach().clone() for p in tr.ema_model.parameters()])
        assert tr.scheduler.get_last_lr()[0] == pytest.approx(5e-4)
        # sampling uses the EMA weights the fused kernel wrote
        out = tr.sample_scales(scale_mul=(1, 1), batch_size=2, save_images=False)
        assert torch.isfinite(out[-1]).all()
    print("fused vs torch Adam: losses", losses[True], losses[False], "max param diff / max|p|",
          max(float((a - b).abs().max() / max(1e-3, float(b.abs().max()))) for a, b in zip(finals[True][0], finals[False][0])))
    assert np.allclose(losses[True], losses[False], rtol=2e-3), (losses[True], losses[False])
    # eight Adam steps move a parameter by up to 8e-3; the two implementations differ in rounding only, but the L1
    # loss gradient (a sign) amplifies last-bit differences: allow 2 % of max|p| (~10 % of the movement)
    for a, b in zip(finals[True][0], finals[False][0]):
        assert float((a - b).abs().max()) <= 2e-2 * max(1e-3, float(b.abs().max()))
    for a, b in zip(finals[True][1], finals[False][1]):
        assert float((a - b).abs().max()) <= 2e-2 * max(1e-3, float(b.abs().max()))


@pytest.mark.parametrize("multimem", [0, 1])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_fused_step_equals_allreduce_plus_adam(world, multimem):
    """tools/fused_dp_check.py at every world size the box offers (the driver's 1-GPU box skips; `gpurun --gpus N`
    runs it; logs of the 2/4/8-rank runs are committed under profiles/).  multimem=1: the in-switch
    (multimem.ld_reduce) gradient sum instead of rank-ordered peer loads."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs of one NVLink box, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29517 + world), str(REPO / "tools" / "fused_dp_check.py")]
    env = dict(os.environ, SINDDM_FUSED_MULTIMEM=str(multimem))
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert f"FUSED_DP_OK world={world} multimem={multimem}" in proc.stdout


@pytest.mark.parametrize("kinds", [(True, True), (False, False), (True, False)])
def test_optimizer_state_checkpoint_resumes_adam(tmp_path, kinds):
    """optim-N.pt next to model-N.pt (whose keys stay the reference's).  A run resumed from the
    pair continues exactly like the uninterrupted run -- also across the two optimizer implementations -- whereas
    the reference's checkpoint alone restarts Adam's moments (quirk Q12)."""
    save_fused, load_fused = kinds
    a = _trainer(tmp_path / "a", save_fused)
    torch.manual_seed(3)
    for i in range(5):
        a.train_step(s=i % 3)
    a.save(1)
    assert set(torch.load(a.results_folder / "model-1.pt", map_location="cpu")) == {
        "step", "model", "ema", "sched", "running_loss", "running_scale"}
    torch.manual_seed(4)
    for i in range(3):
        a.train_step(s=i % 3)
    want = [p.detach().clone() for p in a.model.parameters()]

    b = _trainer(tmp_path / "a", load_fused)          # same results folder
    b.load(1)
    assert b.step == 5
    torch.manual_seed(4)
    for i in range(3):
        b.train_step(s=i % 3)
    worst = max(float((x - y).abs().max() / (y.abs().max() + 1e-12)) for x, y in zip(b.model.parameters(), want))
    assert worst <= (1e-6 if save_fused == load_fused else 1e-2), worst

    # without the optimizer file the moments restart and the continuation visibly differs
    (a.results_folder / "optim-1.pt").unlink()
    c = _trainer(tmp_path / "a", load_fused)
    c.load(1)
    torch.manual_seed(4)
    for i in range(3):
        c.train_step(s=i % 3)
    drift = max(float((x - y).abs().max() / (y.abs().max() + 1e-12)) for x, y in zip(c.model.parameters(), want))
    print(f"optimizer checkpoint {kinds}: resumed-vs-uninterrupted {worst:.3e}, without optimizer file {drift:.3e}")
    assert drift > 3 * max(worst, 1e-6)
