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

    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
