"""Parity of every CUDA operator, called through the C ABI (sinddm_b200.ops -> libsinddm_b200.so), against
the same op stated with torch functional ops in float64 on the device (cuDNN/cuBLAS TF32 off).

Tolerances (stated here, used below):
  * math = fp32 (CUDA cores)      : max |err| <= 2e-5 * max|ref|   (fp32 accumulation-order noise only)
  * math = tf32 (tcgen05)         : operands carry 10 mantissa bits (2^-11 relative rounding each), fp32
                                    accumulate: relative L2 error <= 2e-3, max |err| <= 1e-2 * max|ref|
                                    -- the numerics class of the reference's own GPU default (cuDNN TF32).
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import max_err_rel, rel_err

pytestmark = pytest.mark.gpu

FP32_MAX = 2e-5
TF32_L2 = 2e-3
TF32_MAX = 1e-2


@pytest.fixture(scope="module")
def ops():
    from sinddm_b200 import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def dev():
    return torch.device("cuda:0")


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev())


def nhwc(x):   # NCHW -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def check(out, ref, math):
    if math == 0:
        assert max_err_rel(out, ref) <= FP32_MAX, (max_err_rel(out, ref), rel_err(out, ref))
    else:
        assert rel_err(out, ref) <= TF32_L2 and max_err_rel(out, ref) <= TF32_MAX, (rel_err(out, ref), max_err_rel(out, ref))


# ---------------------------------------------------------------------------------------------------

def test_layout_roundtrip(ops):
    x = randn(3, 3, 19, 23)
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y, nhwc(x))
    assert torch.equal(ops.nhwc_to_nchw(y), x)


CONV_CASES = [
    # B, H, W, Cin, Cout   (ragged sizes on purpose: tile = 8x16 pixels)
    (2, 19, 23, 80, 80),
    (1, 8, 16, 160, 160),
    (3, 33, 17, 80, 160),
    (2, 48, 64, 160, 80),
    (1, 67, 90, 160, 160),
]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_bias_gelu(ops, case, math):
    B, H, W, Ci, Co = case
    x = randn(B, Ci, H, W, seed=1)
    w = randn(Co, Ci, 3, 3, seed=2, scale=(Ci * 9) ** -0.5)
    b = randn(Co, seed=3, scale=0.1)
    wf, _ = ops.pack_conv_weights(w, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(x), wf, math=math, bias=b, gelu=True, save_pre=True)
    pre = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    check(nchw(r["pre"]), pre, math)
    check(nchw(r["out"]), F.gelu(pre), math)


@pytest.mark.parametrize("math", [0, 1])
def test_conv3x3_with_residual_conv_slices_and_final(ops, math):
    # block l4's second conv: 3x3 (80->80) + 1x1 residual conv from the 160-channel block input + final 1x1
    B, H, W, Cm, Cr = 2, 21, 37, 80, 160
    a1 = randn(B, Cm, H, W, seed=4)
    xin = randn(B, Cr, H, W, seed=5)
    w2 = randn(Cm, Cm, 3, 3, seed=6, scale=(Cm * 9) ** -0.5)
    wr = randn(Cm, Cr, 1, 1, seed=7, scale=Cr ** -0.5)
    bias = randn(Cm, seed=8, scale=0.1)
    wfin = randn(3, Cm, 1, 1, seed=9, scale=Cm ** -0.5)
    bfin = randn(3, seed=10, scale=0.1)
    w2p, _ = ops.pack_conv_weights(w2, round_tf32=bool(math))
    wrp, _ = ops.pack_conv_weights(wr, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(a1), w2p, math=math, bias=bias, in_res=nhwc(xin), w_res=wrp.reshape(Cm, Cr),
                         w_final=wfin.reshape(3, Cm).contiguous(), b_final=bfin)
    ref = F.conv2d(a1.double(), w2.double(), bias.double(), padding=1) + F.conv2d(xin.double(), wr.double())
    check(nchw(r["out"]), ref, math)
    check(r["final"], F.conv2d(ref, wfin.double(), bfin.double()), math)


@pytest.mark.parametrize("math", [0, 1])
def test_conv3x3_identity_residual_and_c3_residual(ops, math):
    B, H, W, Cc = 2, 25, 30, 160
    a1 = randn(B, Cc, H, W, seed=11)
    xin = randn(B, Cc, H, W, seed=12)
    w2 = randn(Cc, Cc, 3, 3, seed=13, scale=(Cc * 9) ** -0.5)
    w2p, _ = ops.pack_conv_weights(w2, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(a1), w2p, math=math, res_add=nhwc(xin))
    check(nchw(r["out"]), F.conv2d(a1.double(), w2.double(), padding=1) + xin.double(), math)
    # l1: residual 1x1 conv from the 3-channel image, computed in the epilogue
    Cm = 80
    a1 = randn(B, Cm, H, W, seed=14)
    x3 = randn(B, 3, H, W, seed=15)
    w2 = randn(Cm, Cm, 3, 3, seed=16, scale=(Cm * 9) ** -0.5)
    wr3 = randn(Cm, 3, 1, 1, seed=17)
    w2p, _ = ops.pack_conv_weights(w2, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(a1), w2p, math=math, x3=nhwc(x3), w_res3=wr3.reshape(Cm, 3).contiguous())
    check(nchw(r["out"]), F.conv2d(a1.double(), w2.double(), padding=1) + F.conv2d(x3.double(), wr3.double()), math)


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", [(2, 19, 23, 80, 160), (1, 40, 33, 160, 160), (2, 9, 70, 160, 80)])
def test_conv3x3_data_gradient_with_gelu_grad(ops, case, math):
    B, H, W, Ci, Co = case
    x = randn(B, Ci, H, W, seed=20).double().requires_grad_(True)
    w = randn(Co, Ci, 3, 3, seed=21, scale=(Ci * 9) ** -0.5)
    z = randn(B, Ci, H, W, seed=22)            # pre-activation that produced x = gelu(z) upstream
    dy = randn(B, Co, H, W, seed=23)
    y = F.conv2d(x, w.double(), padding=1)
    (dx,) = torch.autograd.grad(y, x, dy.double())
    zz = z.double().requires_grad_(True)
    (gz,) = torch.autograd.grad(F.gelu(zz).sum(), zz)
    _, wd = ops.pack_conv_weights(w, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(dy), wd, math=math, dgelu_z=nhwc(z))
    check(nchw(r["out"]), dx * gz, math)
    # 1x1 data gradient (residual conv)
    w1 = randn(Co, Ci, 1, 1, seed=24, scale=Ci ** -0.5)
    _, w1d = ops.pack_conv_weights(w1, round_tf32=bool(math))
    r = ops.conv_forward(nhwc(dy), w1d, math=math)
    (dx1,) = torch.autograd.grad(F.conv2d(x, w1.double()), x, dy.double())
    check(nchw(r["out"]), dx1, math)


def test_conv_small_channels_fp32_path(ops):
    # Cin = 3 (l1.net[0]) and Cout = 3 (data gradient into l1's depthwise output): CUDA-core kernel only
    B, H, W = 2, 19, 23
    x = randn(B, 3, H, W, seed=30)
    w = randn(80, 3, 3, 3, seed=31, scale=27 ** -0.5)
    b = randn(80, seed=32, scale=0.1)
    wf, wd = ops.pack_conv_weights(w)
    r = ops.conv_forward(nhwc(x), wf, math=0, bias=b, gelu=True)
    check(nchw(r["out"]), F.gelu(F.conv2d(x.double(), w.double(), b.double(), padding=1)), 0)
    dy = randn(B, 80, H, W, seed=33)
    xx = x.double().requires_grad_(True)
    (dx,) = torch.autograd.grad(F.conv2d(xx, w.double(), padding=1), xx, dy.double())
    r = ops.conv_forward(nhwc(dy), wd, math=0)
    check(nchw(r["out"]), dx, 0)
    with pytest.raises(RuntimeError):
        ops.conv_forward(nhwc(x), wf, math=1)      # no tensor-core path for Cin = 3: must say so, not fall back


WGRAD_CASES = [
    # B, H, W, Cx, Cy, ntaps
    (2, 19, 23, 80, 80, 9),
    (1, 33, 70, 160, 160, 9),
    (2, 24, 31, 80, 160, 9),
    (3, 17, 40, 160, 80, 9),
    (2, 19, 23, 80, 160, 1),
    (2, 30, 45, 160, 80, 1),
]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("case", WGRAD_CASES)
def test_conv_weight_gradient(ops, case, math):
    B, H, W, Cx, Cy, ntaps = case
    k = 3 if ntaps == 9 else 1
    x = randn(B, Cx, H, W, seed=40)
    dy = randn(B, Cy, H, W, seed=41)
    w = torch.zeros(Cy, Cx, k, k, device=dev(), dtype=torch.float64, requires_grad=True)
    (dw,) = torch.autograd.grad(F.conv2d(x.double(), w, padding=k // 2), w, dy.double())
    out = ops.conv_wgrad(nhwc(x), nhwc(dy), ntaps, math=math)
    check(out, dw, math)


def test_conv_weight_gradient_small_channels(ops):
    B, H, W = 2, 19, 23
    x = randn(B, 3, H, W, seed=42)
    dy = randn(B, 80, H, W, seed=43)
    w = torch.zeros(80, 3, 3, 3, device=dev(), dtype=torch.float64, requires_grad=True)
    (dw,) = torch.autograd.grad(F.conv2d(x.double(), w, padding=1), w, dy.double())
    check(ops.conv_wgrad(nhwc(x), nhwc(dy), 9, math=0), dw, 0)


@pytest.mark.parametrize("C", [3, 80, 160])
def test_depthwise5x5_forward_and_gradients(ops, C):
    B, H, W = 2, 21, 26
    x = randn(B, C, H, W, seed=50)
    w = randn(C, 1, 5, 5, seed=51, scale=0.2)
    b = randn(C, seed=52, scale=0.1)
    cond = randn(B, C, seed=53)
    out = ops.dw5x5(nhwc(x), w, b, cond)
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    cd = cond.double().requires_grad_(True)
    ref = F.conv2d(xd, wd, bd, padding=2, groups=C) + cd[:, :, None, None]
    check(nchw(out), ref, 0)
    dh = randn(B, C, H, W, seed=54)
    gx, gw, gb, gc = torch.autograd.grad(ref, (xd, wd, bd, cd), dh.double())
    add = randn(B, C, H, W, seed=55)
    dx = ops.dw5x5(nhwc(dh), w, None, None, nhwc(add), flip=True)
    check(nchw(dx), gx + add.double(), 0)
    dw, db, dcond = ops.dw5x5_wgrad(nhwc(x), nhwc(dh))
    check(dw, gw, 0)
    check(db, gb, 0)
    check(dcond, gc, 0)


@pytest.mark.parametrize("C", [3, 80, 160])
def test_colsum(ops, C):
    a = randn(5, 13, 17, C, seed=60)
    check(ops.colsum(a), a.double().sum(dim=(0, 1, 2)), 0)


def test_qsample_mix_l1_loss(ops):
    B, H, W = 4, 19, 23
    x_blur, x_orig, noise = randn(B, 3, H, W, seed=70), randn(B, 3, H, W, seed=71), randn(B, 3, H, W, seed=72)
    t = torch.tensor([0, 17, 50, 99], device=dev())
    sa = torch.linspace(0.99, 0.01, 100, device=dev())
    sb = (1 - sa * sa).sqrt()
    gam = torch.linspace(0.0, 1.0, 100, device=dev())
    out = ops.qsample_mix(x_blur, noise, t, sa, sb, x_orig=x_orig, gammas_row=gam)
    e = lambda a: a[t].reshape(B, 1, 1, 1)
    mix = e(gam) * x_blur + (1 - e(gam)) * x_orig
    ref = e(sa) * mix + e(sb) * noise
    assert torch.equal(out, ref)                       # same op order, separately rounded: bit exact
    out0 = ops.qsample_mix(x_blur, noise, t, sa, sb)
    assert torch.equal(out0, e(sa) * x_blur + e(sb) * noise)
    pred = randn(B, 3, H, W, seed=73).requires_grad_(True)
    loss, dpred = ops.l1_loss(noise, pred.detach(), want_grad=True)
    ref_loss = (noise - pred).abs().mean()
    (g,) = torch.autograd.grad(ref_loss, pred)
    assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-6)
    assert torch.allclose(dpred, g, rtol=0, atol=1e-12)


def test_streamed_epilogue_operand_is_deterministic(ops):
    """Regression test for a rare race in the tensor-core epilogue: the residual / saved-activation tile is streamed
    through ONE shared-memory tile per warp by TMA, and the next chunk's load used to be able to overtake the generic
    reads of the current chunk (a few pixels then received the operand of channel c+32).  It showed up in <= 2 % of
    launches at this size; tools/determinism_stress.py is the long version."""
    B, H, W, C = 16, 186, 248, 160
    x = torch.randn(B, H, W, C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    w = randn(C, C, 3, 3, seed=2).to("cuda") / (3 * C ** 0.5)
    wf, _ = ops.pack_conv_weights(w, round_tf32=True)
    bias = randn(C, seed=3).to("cuda")
    res = torch.randn(B, H, W, C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    plain = ops.conv_forward(x, wf, math=1, bias=bias)["out"]
    want = plain + res
    for _ in range(120):
        out = ops.conv_forward(x, wf, math=1, bias=bias, res_add=res)["out"]
        assert torch.equal(out, want)


@pytest.mark.parametrize("world,b,chw", [(2, 3, (3, 5, 7)), (8, 32, (3, 186, 248)), (4, 16, (3, 133, 177)), (8, 2, (3, 48, 64)),
                                          (3, 1, (1, 1, 17))])
def test_data_parallel_noise_shard_equals_rows_of_the_global_draw(world, b, chw):
    """ops.randn_rows (sinddm_philox_normal_rows): rank r's rows of torch.randn's global draw without generating the
    rest -- values bit-identical to draw-and-slice for every rank, and the generator ends in the same state (the next
    torch draw is unchanged).  This is what lets N ranks consume exactly the random numbers one GPU would at 1/N of the
    cost."""
    from sinddm_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    torch.randn(5, device=dev)                      # a non-zero starting offset
    state = torch.cuda.get_rng_state(dev)
    full = torch.randn((b * world, *chw), device=dev)
    after = torch.randint(0, 1000, (8,), device=dev)
    for rank in sorted({0, world - 1, world // 2}):
        torch.cuda.set_rng_state(state, dev)
        rows = ops.randn_rows((b, *chw), rank, world, dev)
        assert torch.equal(rows, full[rank * b:(rank + 1) * b]), (world, rank)
        assert torch.equal(torch.randint(0, 1000, (8,), device=dev), after)


def test_diffusion_randn_uses_the_shard_kernel_under_data_parallel():
    import tempfile
    from sinddm_b200 import MultiScaleGaussianDiffusion, SinDDMNet
    net = SinDDMNet(dim=16, multiscale=True)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=2, scale_factor=1.3, image_sizes=[(23, 19), (30, 25)],
                                      timesteps=100, train_full_t=True, scale_losses=[0.9], results_folder=tempfile.mkdtemp())
    dev = torch.device("cuda:0")
    for shard in (True, False):
        dif.dp_shard_rng = shard
        got = []
        for rank in range(4):
            dif.set_data_parallel(rank, 4)
            torch.manual_seed(77)
            got.append(dif._randn((2, 3, 25, 30), dev))
            tail = torch.randn(3, device=dev)
        torch.manual_seed(77)
        full = torch.randn((8, 3, 25, 30), device=dev)
        assert torch.equal(torch.cat(got), full)
        assert torch.equal(tail, torch.randn(3, device=dev))
