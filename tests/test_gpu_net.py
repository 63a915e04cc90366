"""End-to-end parity of the CUDA path (public Python API -> C ABI -> sm_100a kernels) against
 (1) the committed outputs of the unmodified reference (tests/golden, CPU fp32), and
 (2) the CPU oracle on fresh seeded inputs,
plus size-independent properties at BASELINE.json's full sizes (balloons 186x248, batch 32).

Tolerances:
  math=fp32 : |err| <= 2e-4 * max|ref| for network outputs and gradients (different fp32 summation order
              across ~20 chained layers; no precision is dropped anywhere)
  math=tf32x3 : the fp32 bounds (3xTF32 on the tensor cores: operands split hi + lo, fp32 accumulation)
  math=tf32 : relative L2 <= 5e-3, |err| <= 3e-2 * max|ref| for a single network evaluation / gradient --
              TF32 operand rounding (2^-11) through 8 chained 3x3 convolutions; SURVEY.md H1 measured
              2.6e-4 .. 9.7e-3 abs on |y| <= 5 for the reference's own TF32 default.
"""
import copy
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN_SCALE_LOSSES, GOLDEN_SIZES, max_err_rel, rel_err, rs_tensor

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


STRICT = ("fp32", "tf32x3")     # math modes held to the fp32 bounds


def tol_check(out, ref, math, what=""):
    out, ref = out.detach().cpu(), ref.detach().cpu() if torch.is_tensor(ref) else torch.from_numpy(np.asarray(ref))
    if math in STRICT:
        assert max_err_rel(out, ref) <= 2e-4, (what, max_err_rel(out, ref))
    else:
        assert rel_err(out, ref) <= 5e-3 and max_err_rel(out, ref) <= 3e-2, (what, rel_err(out, ref), max_err_rel(out, ref))


def build(math, dim=160, timesteps=100, sizes=GOLDEN_SIZES, losses=GOLDEN_SCALE_LOSSES, seed=11):
    from oracle import sinddm_oracle as orc
    from sinddm_b200 import MultiScaleGaussianDiffusion, SinDDMNet
    net = SinDDMNet(dim=dim, multiscale=True, device=DEV, math=math)
    net.load_state_dict(orc.synthetic_params(seed=seed, dim=dim), strict=True)
    net.to(DEV)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=len(sizes), scale_factor=1.36, image_sizes=sizes,
                                      timesteps=timesteps, train_full_t=True, scale_losses=losses, loss_type="l1",
                                      reblurring=True, omega=0, device=DEV, results_folder=tempfile.mkdtemp()).to(DEV)
    return net, dif


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_net_forward_vs_reference_golden(golden, math):
    g = golden("g1_net_forward.npz")
    net, _ = build(math)
    x = rs_tensor(101, (2, 3, 19, 23)).to(DEV)
    t = torch.tensor([7, 93], device=DEV)
    with torch.no_grad():
        tol_check(net(x, t, scale=0), g["y_s0"], math, "s0")
        tol_check(net(x, t, scale=3), g["y_s3"], math, "s3")
        y2 = net(rs_tensor(102, (1, 3, 33, 17)).to(DEV), torch.tensor([0], device=DEV), scale=torch.tensor([1], device=DEV))
        tol_check(y2, g["y2_s1"], math, "y2")


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_train_loss_and_grads_vs_reference_golden(golden, math):
    g = golden("g2_train_loss_grads.npz")
    net, dif = build(math)
    for s in (0, 2):
        h, w = GOLDEN_SIZES[s][1], GOLDEN_SIZES[s][0]
        x_orig = rs_tensor(200 + s, (3, 3, h, w), 0.5).clamp(-1, 1).to(DEV)
        x_blur = rs_tensor(210 + s, (3, 3, h, w), 0.5).clamp(-1, 1).to(DEV)
        t = torch.from_numpy(g[f"s{s}_t"]).to(DEV)
        noise = torch.from_numpy(g[f"s{s}_noise"]).to(DEV)
        net.zero_grad()
        loss = dif.p_losses(x_blur if s > 0 else x_orig, t, s, noise=noise, x_orig=x_orig)
        loss.backward()
        assert loss.item() == pytest.approx(float(g[f"s{s}_loss"]), rel=1e-5 if math in STRICT else 2e-3)
        for name, prm in net.named_parameters():
            gr = prm.grad.detach().cpu()
            if gr.numel() <= 4096:
                tol_check(gr, g[f"s{s}_grad/{name}"], math, f"s{s} {name}")
            else:
                samp = gr.reshape(-1)[:: max(1, gr.numel() // 512)][:512]
                tol_check(samp, g[f"s{s}_gsample/{name}"], math, f"s{s} {name}")
                nrm = float(gr.double().norm())
                assert nrm == pytest.approx(float(g[f"s{s}_gnorm/{name}"]), rel=2e-4 if math in STRICT else 5e-3), name


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_p_sample_vs_reference_golden(golden, math):
    """The reference draws the step noise with torch.randn on ITS device (CPU there); here it is injected through
    the diffusion object's single noise entry point (_randn) so the fused ddpm_step sees the golden draw."""
    import sinddm_b200.diffusion as D
    g = golden("g4_p_sample.npz")
    net, dif = build(math)
    orig = D.noise_like
    try:
        for s, ti in [(0, 50), (0, 0), (2, 20), (2, 0), (4, 1)]:
            h, w = GOLDEN_SIZES[s][1], GOLDEN_SIZES[s][0]
            xt = rs_tensor(300 + 10 * s + ti, (2, 3, h, w)).to(DEV)
            dif.img_prev_upsample = rs_tensor(400 + 10 * s + ti, (2, 3, h, w), 0.5).clamp(-1, 1).to(DEV)
            noise = torch.from_numpy(g[f"s{s}_t{ti}_noise"]).to(DEV)
            D.noise_like = lambda shape, device, repeat=False: noise
            dif._randn = lambda shape, device, noise=noise: noise
            out = dif.p_sample(xt, torch.full((2,), ti, device=DEV, dtype=torch.long), s)
            tol_check(out, g[f"s{s}_t{ti}_out"], math, f"s={s} t={ti}")
    finally:
        D.noise_like = orig


def test_ddpm_step_vs_oracle_bitlevel():
    """ddpm_step alone (eps injected): same fp32 op order as the reference -> agreement to ~1 ulp."""
    from oracle import sinddm_oracle as orc
    from sinddm_b200 import ops
    net, dif = build("fp32")
    sch = orc.Schedule(5, GOLDEN_SCALE_LOSSES, timesteps=100, train_full_t=True)
    for s, ti in [(0, 99), (0, 1), (0, 0), (3, 30), (3, 1), (3, 0)]:
        shape = (3, 3, 17, 29)
        xt, eps, noise = rs_tensor(1, shape), rs_tensor(2, shape), rs_tensor(3, shape)
        xtil = rs_tensor(4, shape, 0.5).clamp(-1, 1)
        t = torch.full((3,), ti, dtype=torch.long)
        ref = orc.p_sample_update(sch, xt, eps, t, s, noise, xtil)
        reblur = s > 0
        out = ops.ddpm_step(xt.to(DEV), eps.to(DEV), noise.to(DEV), t.to(DEV), dif._tables(),
                            x_tilde=xtil.to(DEV) if reblur else None,
                            gammas_row=dif._gamma_row(s) if reblur else None, reblur_mode=reblur)
        assert torch.allclose(out.cpu(), ref, rtol=2e-6, atol=2e-6), (s, ti, (out.cpu() - ref).abs().max())


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_seeded_chain_vs_reference_golden(golden, math):
    """12-step chain at scale 0 and a 5-step sample_via_scale at scale 1, noise streams replayed from the
    reference's CPU generator (the CUDA generator is a different stream by construction)."""
    import sinddm_b200.diffusion as D
    g = golden("g5_chains.npz")
    net, dif = build(math, timesteps=12)
    dif.use_step_graph = False      # the replayed noise comes from a CPU generator: not capturable
    assert dif.num_timesteps_ideal == list(g["T12_ideal"])
    h, w = GOLDEN_SIZES[0][1], GOLDEN_SIZES[0][0]
    gen = torch.Generator().manual_seed(5)
    orig_nl, orig_randn = D.noise_like, dif._randn
    try:
        D.noise_like = lambda shape, device, repeat=False: torch.randn(shape, generator=gen).to(device)
        dif._randn = lambda shape, device: torch.randn(tuple(shape), generator=gen).to(device)
        s0 = dif.sample(batch_size=2)
        # chains amplify per-step differences; the 246-evaluation balloons chain measures 2.7e-6 (fp32) / 2.3e-3 (tf32)
        # max abs (tests/test_gpu_parity_r02.py): the 12-step schedule takes larger steps, bound 4x that
        atol = 1e-4 if math in STRICT else 1e-2
        print(f"12-step chain {math}: max abs err {float((s0.cpu() - torch.from_numpy(g['chain_s0'])).abs().max()):.2e}")
        assert (s0.cpu() - torch.from_numpy(g["chain_s0"])).abs().max() <= atol
        gen.manual_seed(6)
        s1 = dif.sample_via_scale(2, torch.from_numpy(g["chain_s0"]).to(DEV), s=1, scale_mul=(1, 1),
                                  custom_sample=True, custom_img_size_idx=1, custom_t=5)
        print(f"5-step via-scale chain {math}: max abs err {float((s1.cpu() - torch.from_numpy(g['chain_s1'])).abs().max()):.2e}")
        assert (s1.cpu() - torch.from_numpy(g["chain_s1"])).abs().max() <= atol
    finally:
        D.noise_like = orig_nl
        dif._randn = orig_randn


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_graph_replayed_sampling_equals_eager_sampling(math):
    """The sampling loops replay one captured CUDA graph per timestep; with the same CUDA generator seed the
    replayed chain (denoiser, noise draws, ddpm_step) must reproduce the eager chain bit for bit, at scale 0 and at
    a reblurring scale, also when the graph is reused with a new upsampled image and after a weight update."""
    import sinddm_b200.diffusion as D
    net, dif = build(math, timesteps=12)
    prev = rs_tensor(77, (2, 3, GOLDEN_SIZES[0][1], GOLDEN_SIZES[0][0]), 0.5).clamp(-1, 1).to(DEV)

    def run(use_graph, seed, img):
        dif.use_step_graph = use_graph
        torch.manual_seed(seed)
        a = dif.sample(batch_size=2)
        b = dif.sample_via_scale(2, img, s=1, scale_mul=(1, 1), custom_sample=True, custom_img_size_idx=1, custom_t=7)
        return a.clone(), b.clone()

    e0, e1 = run(False, 3, prev)
    before = D.graph_replayed_launches
    g0, g1 = run(True, 3, prev)
    assert D.graph_replayed_launches > before, "the graph path did not run"
    assert torch.equal(e0, g0) and torch.equal(e1, g1)
    # cached graphs, new seed and a different previous-scale image
    e0, e1 = run(False, 4, -prev)
    g0, g1 = run(True, 4, -prev)
    assert torch.equal(e0, g0) and torch.equal(e1, g1)
    # weights updated in place: the eager first step repacks them, the replays must see the new values
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(0.9)
    e0, e1 = run(False, 5, prev)
    g0, g1 = run(True, 5, prev)
    assert torch.equal(e0, g0) and torch.equal(e1, g1)
    assert torch.isfinite(g1).all()


@pytest.mark.parametrize("math", ["fp32", "tf32", "tf32x3"])
def test_fresh_inputs_vs_oracle(math):
    """New seeds, odd sizes, batch 5: forward and all 52 gradients against the CPU oracle."""
    from oracle import sinddm_oracle as orc
    sizes = [(31, 22), (45, 31)]
    net, dif = build(math, sizes=sizes, losses=[0.9], seed=23)
    params = orc.synthetic_params(seed=23, dim=160)
    sch = orc.Schedule(2, [0.9], timesteps=100, train_full_t=True)
    s = 1
    h, w = sizes[s][1], sizes[s][0]
    x_orig = rs_tensor(1, (5, 3, h, w), 0.5).clamp(-1, 1)
    x_blur = rs_tensor(2, (5, 3, h, w), 0.5).clamp(-1, 1)
    t = torch.tensor([0, 3, 40, 77, 99])
    noise = rs_tensor(3, (5, 3, h, w))
    ref_loss, ref_grads = orc.loss_and_grads(params, sch, x_blur, t, s, noise, x_orig=x_orig)
    net.zero_grad()
    loss = dif.p_losses(x_blur.to(DEV), t.to(DEV), s, noise=noise.to(DEV), x_orig=x_orig.to(DEV))
    loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-5 if math in STRICT else 2e-3)
    for name, prm in net.named_parameters():
        tol_check(prm.grad, ref_grads[name], math, name)


@pytest.mark.parametrize("math", ["tf32", "tf32x3"])
def test_tensor_core_path_vs_cuda_core_twin_full_size(math):
    """BASELINE size (balloons finest scale 186x248), batch 4: tcgen05 path (TF32, and 3xTF32 held to the fp32 bound)
    vs the exact fp32 twin."""
    sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
    losses = [1.1, 0.78, 0.55, 0.39]
    net_tc, dif_tc = build(math, sizes=sizes, losses=losses)
    net_32, dif_32 = build("fp32", sizes=sizes, losses=losses)
    s = 4
    B = 4
    x = rs_tensor(5, (B, 3, 186, 248), 0.5).clamp(-1, 1).to(DEV)
    t = torch.tensor([1, 20, 60, 99], device=DEV)
    noise = rs_tensor(6, (B, 3, 186, 248)).to(DEV)
    for net, dif in ((net_tc, dif_tc), (net_32, dif_32)):
        net.zero_grad()
        dif.p_losses(x, t, s, noise=noise, x_orig=x).backward()
    worst = 0.0
    for (name, a), (_, b) in zip(net_tc.named_parameters(), net_32.named_parameters()):
        worst = max(worst, max_err_rel(a.grad.cpu(), b.grad.cpu()))
        tol_check(a.grad, b.grad, math, name)
    with torch.no_grad():
        y, y32 = net_tc(x, t, s), net_32(x, t, s)
        print(f"full size {math} vs fp32 twin: forward max err / max|ref| {max_err_rel(y.cpu(), y32.cpu()):.2e}, "
              f"worst gradient {worst:.2e}")
        tol_check(y, y32, math, "forward")


@pytest.mark.parametrize("case", [("seascape finest, 16 per GPU (configs[3])", 16, 200, 249),
                                  ("starry_night x(2,2) finest, 8 per GPU (configs[4])", 8, 396, 504),
                                  ("starry_night x(2,2) coarsest", 8, 98, 124)])
def test_other_baseline_shapes_tensor_core_vs_cuda_core_twin(case):
    """The denoiser at the shapes of BASELINE.json's seascape / starry_night configs (sizes that are not multiples
    of the 16-pixel tiles, 400x500 images): tcgen05 path against its fp32 CUDA-core twin (first 2 images, the fp32 twin
    is slow), and batch rows independent of the batch size."""
    _, B, H, W = case
    sizes = [(W, H)]
    net_tc, _ = build("tf32", sizes=sizes, losses=[])
    net_32, _ = build("fp32", sizes=sizes, losses=[])
    x = rs_tensor(31, (B, 3, H, W), 0.5).clamp(-1, 1).to(DEV)
    t = (torch.arange(B, device=DEV) * 11) % 100
    with torch.no_grad():
        y = net_tc(x, t, 0)
        assert torch.isfinite(y).all()
        y32 = net_32(x[:2].contiguous(), t[:2].contiguous(), 0)
        tol_check(y[:2], y32, "tf32", "forward")
        y2 = net_tc(x[:2].contiguous(), t[:2].contiguous(), 0)
        assert torch.allclose(y[:2], y2, rtol=0, atol=1e-6 * float(y.abs().max()))


def test_full_size_batch32_properties():
    """Size-independent properties at the headline configuration (186x248, batch 32, TF32 path):
    identical rows give identical outputs; a batch row equals the same sample run alone; the gradient of a
    batch is the mean of its two half-batch gradients (what data parallelism relies on)."""
    sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
    net, dif = build("tf32", sizes=sizes, losses=[1.1, 0.78, 0.55, 0.39])
    B, s = 32, 4
    img = rs_tensor(7, (1, 3, 186, 248), 0.5).clamp(-1, 1).to(DEV)
    x = img.repeat(B, 1, 1, 1)
    t = torch.arange(B, device=DEV) * 3
    t[1] = t[0]
    with torch.no_grad():
        y = net(x, t, s)
        assert torch.equal(y[0], y[1])                                   # same (x, t) -> same bits
        y_single = net(x[5:6].contiguous(), t[5:6].contiguous(), s)
        d = (y[5:6] - y_single).abs()
        assert float(d.max()) <= 1e-6 * float(y.abs().max()), (float(d.max()), int((d > 0).sum()),
                                                                  (d > 0).nonzero()[:8].tolist())
    noise = rs_tensor(8, (B, 3, 186, 248)).to(DEV)
    net.zero_grad()
    dif.p_losses(x, t, s, noise=noise, x_orig=x).backward()
    full = [p.grad.clone() for p in net.parameters()]
    halves = []
    for lo in (0, B // 2):
        net.zero_grad()
        sl = slice(lo, lo + B // 2)
        dif.p_losses(x[sl].contiguous(), t[sl].contiguous(), s, noise=noise[sl].contiguous(),
                     x_orig=x[sl].contiguous()).backward()
        halves.append([p.grad.clone() for p in net.parameters()])
    for f, a, b in zip(full, halves[0], halves[1]):
        mean = 0.5 * (a + b)
        assert rel_err(f, mean) <= 2e-3 or float(f.abs().max()) < 1e-12


def test_error_paths():
    from sinddm_b200 import _capi
    net, dif = build("tf32")
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 19, 23), torch.zeros(1, dtype=torch.long), 0)          # CPU tensor: no fallback
    x = rs_tensor(1, (2, 3, 19, 23)).to(DEV)
    t = torch.tensor([1, 2], device=DEV)
    y1 = net(x, t, 0)
    _ = net(x, t, 0)                                                                  # overwrites saved activations
    with pytest.raises(RuntimeError):
        y1.sum().backward()
    # deepcopy (the trainer's EMA copy) gets its own runtime and gives the same result
    twin = copy.deepcopy(net)
    with torch.no_grad():
        assert torch.equal(twin(x, t, 0), net(x, t, 0))


def test_repeated_launches_are_bit_identical():
    """Every kernel on the path is deterministic (fixed reduction orders, no atomics): the same forward, the same
    training gradients and the same sampling step must reproduce bit for bit.  Also the short form of
    tools/determinism_stress.py, which caught a rare shared-memory race in the streamed epilogue operand."""
    sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
    net, dif = build("tf32", sizes=sizes, losses=[1.1, 0.78, 0.55, 0.39])
    for B, s in ((16, 4), (4, 2)):
        h, w = sizes[s][1], sizes[s][0]
        x = rs_tensor(40 + s, (B, 3, h, w), 0.5).clamp(-1, 1).to(DEV)
        t = (torch.arange(B, device=DEV) * 5) % 100
        with torch.no_grad():
            ref = net(x, t, s).clone()
            for _ in range(25):
                assert torch.equal(net(x, t, s), ref)
    x = rs_tensor(50, (8, 3, 94, 126), 0.5).clamp(-1, 1).to(DEV)
    noise = rs_tensor(51, (8, 3, 94, 126)).to(DEV)
    t = torch.arange(8, device=DEV) * 9

    def grads():
        net.zero_grad()
        dif.p_losses(x, t, 2, noise=noise, x_orig=x).backward()
        return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()

    g0 = grads()
    for _ in range(8):
        assert torch.equal(grads(), g0)


def test_determinism_stress_300():
    """tools/determinism_stress.py with 300 repetitions (forward at four shapes incl. 32x186x248, training gradients):
    the streamed epilogue operand of tc_conv is handed between the generic and the async proxy with explicit fences;
    a missing one showed up as a <= 2 % per-launch corruption in round 1, so the long form stays in the GPU suite."""
    import subprocess
    import sys
    from pathlib import Path
    tool = Path(__file__).resolve().parents[1] / "tools" / "determinism_stress.py"
    proc = subprocess.run([sys.executable, str(tool), "300"], capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "nondeterministic results: 0" in proc.stdout, proc.stdout[-2000:]
