"""The drop-in path end to end on the GPU: main.py (flag-compatible CLI) -> create_img_scales ->
MultiscaleTrainer.train() -> sample_scales(), on a synthetic image, plus checkpoint round trip and the
trainer's bookkeeping (EMA cadence, LR schedule, loss log)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]


def _image(folder, w=124, h=93, seed=0):
    from PIL import Image
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([np.sin(0.07 * xx + c) + np.cos(0.05 * yy * (c + 1)) for c in range(3)], -1)
    img += 0.1 * rs.standard_normal(img.shape)
    img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    os.makedirs(folder, exist_ok=True)
    Image.fromarray(img).save(os.path.join(folder, "synth.png"))


def test_main_train_then_sample(tmp_path):
    sys.path.insert(0, str(REPO))
    import main as cli
    ds = str(tmp_path / "data") + "/"
    _image(ds)
    res = str(tmp_path / "results")
    torch.manual_seed(0)
    cli.main(["--scope", "synth", "--mode", "train", "--dataset_folder", ds, "--image_name", "synth.png",
              "--results_folder", res, "--train_num_steps", "12", "--train_batch_size", "4",
              "--save_and_sample_every", "6", "--avg_window", "4", "--sample_batch_size", "2"])
    out = Path(res) / "synth"
    assert (out / "model-1.pt").exists() and (out / "model-2.pt").exists()
    assert (out / "sample-2.png").exists()
    ck = torch.load(out / "model-2.pt", map_location="cpu")
    assert set(ck) == {"step", "model", "ema", "sched", "running_loss", "running_scale"}
    assert ck["step"] == 12 and len(ck["model"]) == 65 and len(ck["running_loss"]) == 3
    assert all(np.isfinite(ck["running_loss"]))
    finals = list((out).glob("final_samples_unbatched_*/*_out_b*.png"))
    assert len(finals) == 2
    # resume + sample only
    cli.main(["--scope", "synth", "--mode", "sample", "--dataset_folder", ds, "--image_name", "synth.png",
              "--results_folder", res, "--load_milestone", "2", "--sample_batch_size", "2"])


def test_trainer_bookkeeping_and_loss_decreases(tmp_path):
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet, create_img_scales
    ds = str(tmp_path / "data") + "/"
    _image(ds)
    sizes, losses, sf, ns = create_img_scales(ds, "synth.png", scale_factor=1.411, create=True, auto_scale=50000)
    dev = "cuda:0"
    torch.manual_seed(1)
    net = SinDDMNet(dim=160, multiscale=True, device=dev).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=ns, scale_factor=sf, image_sizes=sizes, timesteps=100,
                                      train_full_t=True, scale_losses=losses, device=dev,
                                      results_folder=str(tmp_path / "r")).to(dev)
    tr = MultiscaleTrainer(dif, ds, n_scales=ns, scale_factor=sf, image_sizes=sizes, train_batch_size=8,
                           train_lr=1e-3, train_num_steps=60, gradient_accumulate_every=1, step_start_ema=20,
                           update_ema_every=10, save_and_sample_every=10 ** 9, avg_window=20,
                           sched_milestones=[30, 50], results_folder=str(tmp_path / "r"), device=dev)
    assert len(tr.data_list) == ns and tr.data_list[1][0].shape[0] == 8
    tr.train()
    assert tr.step == 60
    assert tr.scheduler.get_last_lr()[0] == pytest.approx(1e-3 * 0.25)
    # running_loss[0] is one loss / window (quirk Q5); later entries are window means and must go down
    assert len(tr.running_loss) == 3
    assert tr.running_loss[2] < tr.running_loss[1]
    # EMA: hard copy until step 20, exponential average afterwards -> differs from the live model, stays finite
    diffs = [float((a - b).abs().max()) for a, b in zip(tr.model.parameters(), tr.ema_model.parameters())]
    assert max(diffs) > 0 and all(np.isfinite(diffs))


def test_harmonization_and_style_transfer_modes(tmp_path):
    """image2image (trainer.py:287-361) through the CLI after a short training run: files are written, the
    harmonized composite equals the input image away from the (dilated, blurred) mask and differs inside it."""
    from PIL import Image
    sys.path.insert(0, str(REPO))
    import main as cli
    ds = str(tmp_path / "data") + "/"
    _image(ds)
    res = str(tmp_path / "results")
    torch.manual_seed(0)
    cli.main(["--scope", "synth", "--mode", "train", "--dataset_folder", ds, "--image_name", "synth.png",
              "--results_folder", res, "--train_num_steps", "6", "--train_batch_size", "4",
              "--save_and_sample_every", "6", "--avg_window", "3", "--sample_batch_size", "2"])
    i2i = Path(ds) / "i2i"
    i2i.mkdir()
    base = np.asarray(Image.open(Path(ds) / "synth.png").convert("RGB")).copy()
    comp = base.copy()
    comp[30:50, 40:70] = [255, 40, 40]                       # pasted object
    Image.fromarray(comp).save(i2i / "composite.png")
    mask = np.zeros(base.shape[:2], dtype=np.uint8)
    mask[30:50, 40:70] = 255
    Image.fromarray(np.stack([mask] * 3, -1)).save(i2i / "mask.png")
    common = ["--scope", "synth", "--dataset_folder", ds, "--image_name", "synth.png", "--results_folder", res,
              "--load_milestone", "1", "--sample_batch_size", "2", "--input_image", "composite.png"]
    cli.main(common + ["--mode", "harmonization", "--harm_mask", "mask.png", "--start_t_harm", "5"])
    cli.main(common + ["--mode", "style_transfer", "--start_t_style", "8"])
    out = Path(res) / "synth"
    harm = sorted(out.glob("unbatched_i2i_s*_t_*_5_20*/composite.png_out_b0_i2i.png"))
    style = sorted(out.glob("unbatched_i2i_s*_t_*_8_20*/composite.png_out_b0_i2i.png"))
    assert len(harm) == 1 and len(style) == 1
    assert len(list((out / "i2i_final_samples").glob("composite_i2i_s_*_hist_off_*.png"))) == 1
    assert len(list((out / "i2i_final_samples").glob("composite_i2i_s_*_hist_on_*.png"))) == 1
    h = np.asarray(Image.open(harm[0]).convert("RGB")).astype(int)
    assert h.shape == comp.shape
    far = np.ones(mask.shape, bool)
    far[0:80, 10:100] = False                                 # > 7 px dilation + 4 sigma of blur away from the mask
    assert np.abs(h[far] - comp.astype(int)[far]).max() <= 1  # untouched outside the mask (8-bit rounding)
    assert np.abs(h[35:45, 45:65] - comp.astype(int)[35:45, 45:65]).max() > 1


def _small_trainer(tmp_path, **kw):
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet, create_img_scales
    ds = str(tmp_path / "data") + "/"
    if not os.path.exists(ds):
        _image(ds)
    sizes, losses, sf, ns = create_img_scales(ds, "synth.png", scale_factor=1.411, create=True, auto_scale=50000)
    dev = "cuda:0"
    net = SinDDMNet(dim=160, multiscale=True, device=dev).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=ns, scale_factor=sf, image_sizes=sizes, timesteps=100,
                                      train_full_t=True, scale_losses=losses, device=dev,
                                      results_folder=str(tmp_path / "r")).to(dev)
    args = dict(n_scales=ns, scale_factor=sf, image_sizes=sizes, train_batch_size=4, train_lr=1e-3,
                train_num_steps=10 ** 9, gradient_accumulate_every=1, step_start_ema=2, update_ema_every=1,
                save_and_sample_every=10 ** 9, avg_window=10 ** 9, results_folder=str(tmp_path / "r"), device=dev)
    args.update(kw)
    return MultiscaleTrainer(dif, ds, **args), (sizes, losses, sf, ns)


def test_side_stream_scale_draw_keeps_the_reference_rng_stream(tmp_path):
    """trainer.py:197 draws s with torch.multinomial from the device generator between the previous step's randn and
    this step's randint (quirk Q7).  The trainer issues that call on a side stream (no drain of the training stream);
    the values -- s, and the t / noise drawn after it, visible through the loss -- must be those of the plain
    in-stream call."""
    seqs, losses = [], []
    for side in (True, False):
        torch.manual_seed(11)
        tr, _ = _small_trainer(tmp_path)
        tr._prepare_training()
        if not side:
            tr._draw_stream = None                      # plain torch.multinomial on the training stream
        seq = []
        draw = tr._draw_scale
        tr._draw_scale = lambda draw=draw, seq=seq: (seq.append(draw()) or seq[-1])
        torch.manual_seed(5)
        ls = [float(tr.train_step()) for _ in range(8)]
        seqs.append(seq)
        losses.append(ls)
    assert seqs[0] == seqs[1] and len(set(seqs[0])) > 1, seqs
    assert losses[0] == losses[1], losses               # same t, same noise, same kernels: bit-identical


def test_host_data_and_per_step_loss_readback_match_the_device_resident_path(tmp_path):
    """host_data=True (batch copied in from pinned memory on a side stream every step) and loss_readback='step' (the
    reference's loss.item() per step) change where the bytes live, not the numbers."""
    out = []
    for host in (False, True):
        torch.manual_seed(3)
        tr, _ = _small_trainer(tmp_path, host_data=host, loss_readback="step" if host else "window", avg_window=4)
        tr.train_num_steps = 8
        torch.manual_seed(9)
        tr.train()
        assert tr.step == 8 and sum(tr.scale_counts) == 8
        if host:
            assert all(t.is_pinned() and not t.is_cuda for pair in tr.data_list for t in pair)
            assert isinstance(tr.last_loss, float)
        out.append(([p.detach().clone() for p in tr.model.parameters()], list(tr.running_loss)))
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b)
    assert np.allclose(out[0][1], out[1][1], rtol=1e-6)


def test_ema_sampling_sees_every_ema_update_on_the_torch_adam_path(tmp_path):
    """on the non-fused optimizer path EMA.update_model_average rebinds `.data`, which changes
    neither the parameter version nor (reliably) its address -- the key the packed conv weights were cached under.
    After an even number of EMA updates ema_model.sample() ran on stale packed weights.  Now the trainer / EMA mark
    the weights as updated: the EMA model's output must equal a freshly built net holding the same state_dict."""
    from sinddm_b200 import SinDDMNet
    torch.manual_seed(2)
    tr, _ = _small_trainer(tmp_path, gradient_accumulate_every=2)          # -> torch.optim.Adam + EMA class
    tr._prepare_training()
    assert tr._fused is None
    dev = "cuda:0"
    h, w = tr.model.image_sizes[1]
    x = torch.randn(2, 3, h, w, device=dev)
    t = torch.tensor([3, 70], device=dev)
    with torch.no_grad():
        tr.ema_model.denoise_fn(x, t, 1)               # packs the EMA weights once
    for n_updates in (2, 4, 5):
        for _ in range(n_updates):
            tr.train_step(s=1)
        fresh = SinDDMNet(dim=160, multiscale=True, device=dev).to(dev)
        fresh.load_state_dict(tr.ema_model.denoise_fn.state_dict())
        with torch.no_grad():
            got = tr.ema_model.denoise_fn(x, t, 1)
            want = fresh(x, t, 1)
        assert torch.equal(got, want), f"EMA model ran on stale packed weights after {n_updates} more updates"
    # and the EMA weights really moved
    assert any(float((a - b).abs().max()) > 0 for a, b in zip(tr.model.parameters(), tr.ema_model.parameters()))


def test_out_of_range_start_timestep_raises(tmp_path):
    """custom_t >= num_timesteps indexed the schedule tables out of bounds inside the kernels; the
    reference's gather raises for the same input."""
    tr, _ = _small_trainer(tmp_path)
    img = torch.zeros(1, 3, *tr.model.image_sizes[0], device="cuda:0")
    with pytest.raises(IndexError):
        tr.ema_model.sample_via_scale(1, img, s=1, custom_t=100)
    with pytest.raises(IndexError):
        tr.ema_model.sample_via_scale(1, img, s=1, custom_t=-1)
