"""Generate tests/golden/*.npz from the UNMODIFIED reference (fallenshock/SinDDM at /root/reference).

Runs only in the build container (the reference does not exist on the GPU box).  The reference is imported
read-only behind a stub shim for the modules this image lacks (skimage, matplotlib -- SURVEY.md 8c); its own
classes (SinDDMNet, MultiScaleGaussianDiffusion, create_img_scales) then compute every stored value on CPU
fp32.  Inputs and weights are regenerated at test time from numpy RandomState seeds (oracle.synthetic_params),
so only outputs are stored.

    python tools/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("SINDDM_REFERENCE", "/root/reference"))
OUT = REPO / "tests" / "golden"


def install_shim():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    noop = lambda *a, **k: None
    mod("skimage", morphology=mod("skimage.morphology"), filters=mod("skimage.filters"),
        exposure=mod("skimage.exposure", match_histograms=noop))
    mod("matplotlib", pyplot=mod("matplotlib.pyplot", plot=noop, grid=noop, ylim=noop, savefig=noop, clf=noop,
                                 rcParams={}))
    sys.path.insert(0, str(REF))


def rs_tensor(seed, shape, scale=1.0):
    return torch.from_numpy((np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32))


def synthetic_image(seed, w, h):
    """Smooth random RGB image (uint8) used for the pyramid golden."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(6):
            fx, fy, ph = rs.uniform(0.01, 0.15), rs.uniform(0.01, 0.15), rs.uniform(0, 6.28)
            img[:, :, c] += rs.uniform(0.3, 1.0) * np.sin(fx * xx + fy * yy + ph)
    img += 0.15 * rs.standard_normal(img.shape)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


def main():
    install_shim()
    # The repo root also holds a drop-in `SinDDM` package: keep it OFF sys.path so `SinDDM.*` below is the
    # reference's, and load the oracle (only for synthetic_params: weights are data, not algorithm) by path.
    import importlib.util
    sys.path[:] = [p for p in sys.path if p not in ("", str(REPO))]
    spec = importlib.util.spec_from_file_location("sinddm_oracle", REPO / "oracle" / "sinddm_oracle.py")
    orc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(orc)
    from SinDDM.functions import create_img_scales
    import SinDDM as _ref_pkg
    assert str(REF) in str(Path(_ref_pkg.__file__ or list(_ref_pkg.__path__)[0]).resolve()), "not the reference package"
    from SinDDM.models import MultiScaleGaussianDiffusion, SinDDMNet
    from PIL import Image

    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    dim = 160
    params = orc.synthetic_params(seed=11, dim=dim)

    net = SinDDMNet(dim=dim, multiscale=True, device="cpu")
    missing = net.load_state_dict(params, strict=True)
    print("load_state_dict:", missing)

    # ---- G1: denoiser forward ------------------------------------------------------------------
    g1 = {}
    x = rs_tensor(101, (2, 3, 19, 23))
    t = torch.tensor([7, 93], dtype=torch.long)
    with torch.no_grad():
        g1["y_s0"] = net(x, t, scale=0).numpy()
        g1["y_s3"] = net(x, t, scale=3).numpy()
        x2 = rs_tensor(102, (1, 3, 33, 17))
        g1["y2_s1"] = net(x2, torch.tensor([0], dtype=torch.long), scale=1).numpy()
    np.savez_compressed(OUT / "g1_net_forward.npz", **g1)

    # ---- G3: schedule ----------------------------------------------------------------------------
    n_scales = 5
    scale_losses = [1.1, 0.78, 0.55, 0.39]       # same order of magnitude as the real pyramids' (Q1)
    sizes = [(23, 19), (30, 25), (41, 34), (56, 46), (76, 62)]   # (W, H) as create_img_scales returns them
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=n_scales, scale_factor=1.36, image_sizes=sizes,
                                      timesteps=100, train_full_t=True, scale_losses=scale_losses,
                                      loss_type="l1", reblurring=True, omega=0, device="cpu",
                                      results_folder=tempfile.mkdtemp())
    g3 = {k: v.numpy() for k, v in dif.state_dict().items() if not k.startswith("denoise_fn.")}
    g3["num_timesteps_ideal"] = np.array(dif.num_timesteps_ideal)
    g3["num_timesteps_trained"] = np.array(dif.num_timesteps_trained)
    g3["image_sizes"] = np.array(dif.image_sizes)
    dif_nofull = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=n_scales, scale_factor=1.36,
                                             image_sizes=sizes, timesteps=100, train_full_t=False,
                                             scale_losses=scale_losses, device="cpu",
                                             results_folder=tempfile.mkdtemp())
    g3["num_timesteps_trained_nofull"] = np.array(dif_nofull.num_timesteps_trained)
    np.savez_compressed(OUT / "g3_schedule.npz", **g3)

    # ---- G2: training loss + gradients (t and noise injected by replaying the seeded RNG stream) --
    g2 = {}
    for s in (0, 2):
        h, w = sizes[s][1], sizes[s][0]
        x_orig = rs_tensor(200 + s, (3, 3, h, w), 0.5).clamp(-1, 1)
        x_blur = rs_tensor(210 + s, (3, 3, h, w), 0.5).clamp(-1, 1)
        net.zero_grad()
        torch.manual_seed(1234 + s)
        loss = dif((x_orig, x_blur), s)           # forward(): randint -> randn_like -> net -> l1
        loss.backward()
        # replay the stream to record what forward() drew (models.py:621,580)
        torch.manual_seed(1234 + s)
        t_drawn = torch.randint(0, dif.num_timesteps_trained[s], (3,)).long()
        noise_drawn = torch.randn_like(x_orig)
        g2[f"s{s}_t"] = t_drawn.numpy()
        g2[f"s{s}_noise"] = noise_drawn.numpy()
        g2[f"s{s}_loss"] = loss.detach().numpy()
        for name, prm in net.named_parameters():
            g = prm.grad.detach().numpy()
            if g.size <= 4096:
                g2[f"s{s}_grad/{name}"] = g
            else:                                   # big tensors: norm + a fixed strided sample
                g2[f"s{s}_gnorm/{name}"] = np.array(np.linalg.norm(g.astype(np.float64)))
                g2[f"s{s}_gsample/{name}"] = g.reshape(-1)[:: max(1, g.size // 512)][:512]
    np.savez_compressed(OUT / "g2_train_loss_grads.npz", **g2)

    # ---- G4: single reverse steps -----------------------------------------------------------------
    g4 = {}
    dif.eval()
    cases = [(0, 50), (0, 0), (2, 20), (2, 0), (4, 1)]
    for s, ti in cases:
        h, w = sizes[s][1], sizes[s][0]
        xt = rs_tensor(300 + 10 * s + ti, (2, 3, h, w))
        xtil = rs_tensor(400 + 10 * s + ti, (2, 3, h, w), 0.5).clamp(-1, 1)
        dif.img_prev_upsample = xtil
        tt = torch.full((2,), ti, dtype=torch.long)
        torch.manual_seed(77 + ti)
        out = dif.p_sample(xt, tt, s)
        torch.manual_seed(77 + ti)
        noise = torch.randn(xt.shape)
        g4[f"s{s}_t{ti}_out"] = out.detach().numpy()
        g4[f"s{s}_t{ti}_noise"] = noise.numpy()
    np.savez_compressed(OUT / "g4_p_sample.npz", **g4)

    # ---- G5: short seeded chains (sample at s=0 with a 12-step schedule; sample_via_scale at s=1) ----
    g5 = {}
    dif12 = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=n_scales, scale_factor=1.36, image_sizes=sizes,
                                        timesteps=12, train_full_t=True, scale_losses=scale_losses, device="cpu",
                                        results_folder=tempfile.mkdtemp())
    torch.manual_seed(5)
    s0 = dif12.sample(batch_size=2)
    g5["chain_s0"] = s0.detach().numpy()
    g5["T12_ideal"] = np.array(dif12.num_timesteps_ideal)
    torch.manual_seed(6)
    s1 = dif12.sample_via_scale(2, s0, s=1, scale_mul=(1, 1), custom_sample=True, custom_img_size_idx=1, custom_t=5)
    g5["chain_s1"] = s1.detach().numpy()
    np.savez_compressed(OUT / "g5_chains.npz", **g5)

    # ---- G6: pyramid builder ------------------------------------------------------------------------
    g6 = {}
    with tempfile.TemporaryDirectory() as td:
        img = synthetic_image(9, 248, 186)
        Image.fromarray(img).save(os.path.join(td, "synth.png"))
        szs, losses, sf, ns = create_img_scales(td + "/", "synth.png", scale_factor=1.411, create=True,
                                                auto_scale=50000)
        g6["sizes"] = np.array(szs)
        g6["rescale_losses"] = np.array(losses, dtype=np.float64)
        g6["scale_factor"] = np.array(sf)
        g6["n_scales"] = np.array(ns)
        for i in range(ns):
            g6[f"scale_{i}"] = np.asarray(Image.open(os.path.join(td, f"scale_{i}", "synth.png")))
            if i > 0:
                g6[f"scale_{i}_recon"] = np.asarray(Image.open(os.path.join(td, f"scale_{i}_recon", "synth.png")))
    # keep the fixture small: only checksums of the images
    for k in [k for k in g6 if k.startswith("scale_") and g6[k].ndim == 3]:
        a = g6.pop(k)
        g6[k + "/shape"] = np.array(a.shape)
        g6[k + "/sum"] = np.array(int(a.astype(np.int64).sum()))
        g6[k + "/wsum"] = np.array(int((a.astype(np.int64).reshape(-1) * (np.arange(a.size) % 251)).sum()))
    np.savez_compressed(OUT / "g6_pyramid.npz", **g6)

    make_r02_goldens(orc, net, params)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


def _store_grads(dst, prefix, named_params):
    """Small tensors whole; big ones as their L2 norm plus a fixed strided sample (as G2 does)."""
    for name, prm in named_params:
        g = prm.grad.detach().numpy() if prm.grad is not None else prm.detach().numpy()
        if g.size <= 4096:
            dst[f"{prefix}/{name}"] = g
        else:
            dst[f"{prefix}_norm/{name}"] = np.array(np.linalg.norm(g.astype(np.float64)))
            dst[f"{prefix}_sample/{name}"] = g.reshape(-1)[:: max(1, g.size // 512)][:512]


def make_r02_goldens(orc, net, params):
    """Round-2 fixtures: parity where the metric lives.
    G7  reference forward + loss + all 52 gradients at the BASELINE finest scale (186x248, B=2, s=4);
    G8  the authors' trained forest EMA weights (committed as a fixture) on the real noisy forest image at
        t in {2, 50, 95}, scales 0 and 3, plus the SURVEY 8c anchor input;
    G9  four steps of the reference MultiscaleTrainer.train() with every drawn (s, t, noise) recorded."""
    from PIL import Image
    from SinDDM.functions import create_img_scales
    from SinDDM.models import MultiScaleGaussianDiffusion, SinDDMNet
    from SinDDM.trainer import MultiscaleTrainer

    # ---- G7 --------------------------------------------------------------------------------------------------
    bal_sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
    bal_losses = [1.20, 0.85, 0.60, 0.42]
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.403, image_sizes=bal_sizes,
                                      timesteps=100, train_full_t=True, scale_losses=bal_losses, loss_type="l1",
                                      reblurring=True, omega=0, device="cpu", results_folder=tempfile.mkdtemp())
    g7 = {}
    s = 4
    h, w = bal_sizes[s][1], bal_sizes[s][0]
    x_orig = rs_tensor(700, (2, 3, h, w), 0.5).clamp(-1, 1)
    x_blur = rs_tensor(701, (2, 3, h, w), 0.5).clamp(-1, 1)
    net.zero_grad()
    torch.manual_seed(4321)
    loss = dif((x_orig, x_blur), s)
    loss.backward()
    torch.manual_seed(4321)
    t_drawn = torch.randint(0, dif.num_timesteps_trained[s], (2,)).long()
    noise_drawn = torch.randn_like(x_orig)
    with torch.no_grad():
        x_noisy = dif.q_sample(x_start=dif.gammas[s - 1][t_drawn].reshape(-1, 1, 1, 1) * x_blur +
                               (1 - dif.gammas[s - 1][t_drawn].reshape(-1, 1, 1, 1)) * x_orig, t=t_drawn,
                               noise=noise_drawn)
        pred = net(x_noisy, t_drawn, scale=s)
    assert abs(float((noise_drawn - pred).abs().mean()) - float(loss)) < 1e-6, "G7 replay does not match forward()"
    g7["t"] = t_drawn.numpy()
    g7["noise_seed_check"] = noise_drawn.reshape(-1)[::9973].numpy()      # noise itself: replayed by the test from the
    g7["noise"] = noise_drawn.numpy().astype(np.float16)                   # float16 copy is NOT used for parity; see below
    g7["loss"] = loss.detach().numpy()
    pn = pred.numpy()
    g7["pred_norm"] = np.array(np.linalg.norm(pn.astype(np.float64)))
    g7["pred_sample"] = pn.reshape(-1)[::53][:8192]
    g7["pred_corner"] = pn[:, :, :4, :4]
    g7["pred_border_row"] = pn[:, :, -1, :]
    _store_grads(g7, "grad", net.named_parameters())
    del g7["noise"]
    np.savez_compressed(OUT / "g7_finest_scale.npz", **g7)
    # the drawn noise is 2x3x186x248 fp32 = 1.1 MB: stored losslessly in its own file
    np.savez_compressed(OUT / "g7_noise.npz", noise=noise_drawn.numpy())

    # ---- G8 --------------------------------------------------------------------------------------------------
    ck = torch.load(str(REF / "results" / "forest" / "model-12.pt"), map_location="cpu", weights_only=False)
    ema_sd = ck["ema"]
    np.savez_compressed(OUT / "forest_ema_state.npz", **{k: v.numpy() for k, v in ema_sd.items()})
    g8 = {}
    with tempfile.TemporaryDirectory() as td:
        import shutil
        shutil.copy(str(REF / "datasets" / "forest" / "forest.png"), os.path.join(td, "forest.png"))
        szs, losses, sf, ns = create_img_scales(td + "/", "forest.png", scale_factor=1.411, create=True, auto_scale=50000)
        fnet = SinDDMNet(dim=160, multiscale=True, device="cpu")
        fdif = MultiScaleGaussianDiffusion(denoise_fn=fnet, n_scales=ns, scale_factor=sf, image_sizes=szs,
                                           timesteps=100, train_full_t=True, scale_losses=losses, loss_type="l1",
                                           reblurring=True, omega=0, device="cpu", results_folder=tempfile.mkdtemp())
        print("forest load_state_dict:", fdif.load_state_dict(ema_sd, strict=True))
        g8["sizes"] = np.array(szs)
        g8["n_scales"] = np.array(ns)
        from torchvision import transforms
        to_t = transforms.Compose([transforms.ToTensor(), transforms.Lambda(lambda t: (t * 2) - 1)])
        for sc in (0, 3):
            img_u8 = np.asarray(Image.open(os.path.join(td, f"scale_{sc}", "forest.png")).convert("RGB"))
            g8[f"s{sc}_img_u8"] = img_u8
            x0 = to_t(Image.fromarray(img_u8)).unsqueeze(0)
            if sc > 0:
                rec_u8 = np.asarray(Image.open(os.path.join(td, f"scale_{sc}_recon", "forest.png")).convert("RGB"))
                g8[f"s{sc}_recon_u8"] = rec_u8
                xb = to_t(Image.fromarray(rec_u8)).unsqueeze(0)
            for ti in (2, 50, 95):
                tt = torch.tensor([ti], dtype=torch.long)
                noise = rs_tensor(800 + 10 * sc + ti, x0.shape)
                with torch.no_grad():
                    if sc > 0:
                        gam = fdif.gammas[sc - 1][tt].reshape(-1, 1, 1, 1)
                        xin = fdif.q_sample(x_start=gam * xb + (1 - gam) * x0, t=tt, noise=noise)
                        lossv = fdif.p_losses(xb, tt, sc, noise=noise, x_orig=x0)
                    else:
                        xin = fdif.q_sample(x_start=x0, t=tt, noise=noise)
                        lossv = fdif.p_losses(x0, tt, sc, noise=noise)
                    eps = fnet(xin, tt, scale=sc)
                assert abs(float((noise - eps).abs().mean()) - float(lossv)) < 1e-6
                g8[f"s{sc}_t{ti}_eps"] = eps.numpy()
                g8[f"s{sc}_t{ti}_loss"] = lossv.numpy()
        # SURVEY.md 8c anchor input
        gen = torch.Generator().manual_seed(1234)
        xa = torch.randn(2, 3, 42, 75, generator=gen)
        ta = torch.tensor([7, 93], dtype=torch.long)
        with torch.no_grad():
            for sc in (0, 3):
                ya = fnet(xa, ta, scale=sc)
                g8[f"anchor_s{sc}"] = ya.numpy()
                print(f"anchor scale={sc}: mean {float(ya.mean()):.6f} std {float(ya.std()):.6f} y[0,0,0,:4] {ya[0, 0, 0, :4].tolist()}")
        g8["anchor_x"] = xa.numpy()
    np.savez_compressed(OUT / "g8_forest_weights.npz", **g8)

    # ---- G9 --------------------------------------------------------------------------------------------------
    g9 = {}
    with tempfile.TemporaryDirectory() as td:
        img = synthetic_image(21, 124, 93)
        Image.fromarray(img).save(os.path.join(td, "synth.png"))
        szs, losses, sf, ns = create_img_scales(td + "/", "synth.png", scale_factor=1.411, create=True, auto_scale=50000)
        tnet = SinDDMNet(dim=160, multiscale=True, device="cpu")
        tnet.load_state_dict(params, strict=True)
        tdif = MultiScaleGaussianDiffusion(denoise_fn=tnet, n_scales=ns, scale_factor=sf, image_sizes=szs, timesteps=100,
                                           train_full_t=True, scale_losses=losses, loss_type="l1", reblurring=True,
                                           omega=0, device="cpu", results_folder=os.path.join(td, "res"))
        torch.manual_seed(99)
        tr = MultiscaleTrainer(tdif, td + "/", n_scales=ns, scale_factor=sf, image_sizes=szs, train_batch_size=2,
                               train_lr=1e-3, train_num_steps=4, gradient_accumulate_every=1, ema_decay=0.995,
                               fp16=False, step_start_ema=2, update_ema_every=1, save_and_sample_every=10 ** 9,
                               avg_window=2, sched_milestones=[3], results_folder=os.path.join(td, "res"), device="cpu")
        g9["sizes"] = np.array(szs)
        g9["scale_losses"] = np.array(losses, dtype=np.float64)
        g9["scale_factor"] = np.array(sf)
        g9["n_scales"] = np.array(ns)
        for i in range(ns):
            # data_list rows are identical images: keep row 0 as uint8 (exactly recoverable: ToTensor()*2-1 of uint8)
            for j, name in enumerate(("orig", "blur")):
                u8 = ((tr.data_list[i][j][0] + 1) * 0.5 * 255).round().to(torch.uint8).numpy()
                assert torch.equal((torch.from_numpy(u8).float().div(255) * 2) - 1, tr.data_list[i][j][0])
                g9[f"data{i}_{name}_u8"] = u8
        # record everything train() draws (trainer.py:197, models.py:621,580)
        rec = {"s": [], "t": [], "noise": [], "loss": []}
        real = {"multinomial": torch.multinomial, "randint": torch.randint, "randn_like": torch.randn_like}

        def multinomial(*a, **k):
            out = real["multinomial"](*a, **k)
            rec["s"].append(int(out))
            return out

        def randint(*a, **k):
            out = real["randint"](*a, **k)
            rec["t"].append(out.clone())
            return out

        def randn_like(*a, **k):
            out = real["randn_like"](*a, **k)
            rec["noise"].append(out.clone())
            return out
        torch.multinomial, torch.randint, torch.randn_like = multinomial, randint, randn_like
        import builtins
        real_print = builtins.print

        def tee(*a, **k):
            msg = " ".join(str(x) for x in a)
            if msg.startswith("step:"):
                rec["loss"].append(msg)
            real_print(*a, **k)
        builtins.print = tee
        try:
            torch.manual_seed(100)
            tr.train()
        finally:
            torch.multinomial, torch.randint, torch.randn_like = real["multinomial"], real["randint"], real["randn_like"]
            builtins.print = real_print
        assert len(rec["s"]) == 4 and len(rec["t"]) == 4 and len(rec["noise"]) == 4, {k: len(v) for k, v in rec.items()}
        g9["s"] = np.array(rec["s"])
        for i in range(4):
            g9[f"t{i}"] = rec["t"][i].numpy()
            g9[f"noise{i}"] = rec["noise"][i].numpy()
        g9["running_loss"] = np.array(tr.running_loss, dtype=np.float64)
        g9["lr_final"] = np.array(tr.scheduler.get_last_lr()[0])
        g9["step_final"] = np.array(tr.step)
        _store_grads(g9, "model", [(n, p.detach().clone().requires_grad_(False)) for n, p in tr.model.denoise_fn.named_parameters()])
        _store_grads(g9, "ema", [(n, p.detach().clone().requires_grad_(False)) for n, p in tr.ema_model.denoise_fn.named_parameters()])
        print("G9 scales drawn:", rec["s"], "running_loss:", tr.running_loss)
    np.savez_compressed(OUT / "g9_trainer_steps.npz", **g9)


if __name__ == "__main__":
    main()
