"""3xTF32 (math = "tf32x3", SINDDM_MATH_TF32X3): the tcgen05 kernels of the TF32 path fed with operands split into two
TF32 values (x = hi + lo), so that every contraction accumulates x_hi*w_hi + x_hi*w_lo + x_lo*w_hi in fp32 -- the
numerics class of the reference with torch.backends.cudnn.allow_tf32 = False (SinDDM/models.py:62-67,79-80 under
plain fp32), on the tensor cores.

Operator level (this file): the split (sinddm_split3) and split weight packing (sinddm_pack_conv_weights, round = 2) are
bit-exact against their definition; conv forward / weight gradient on split operands against float64:
  max |err| <= 2e-5 * max|ref| -- the bound the CUDA-core fp32 operators are held to in tests/test_gpu_ops.py
  (FP32_MAX); plain TF32 on the same inputs is printed beside it (~1e-3).
Network level: tests/test_gpu_net.py runs the golden / oracle parity cases with math = "tf32x3" under the fp32 bounds.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import max_err_rel, rel_err

pytestmark = pytest.mark.gpu

FP32_MAX = 2e-5
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from sinddm_b200 import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def tf32_rna(x):
    """round to nearest, ties away from zero, to 10 mantissa bits (cvt.rna.tf32.f32) on the fp32 bit pattern"""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def test_split3_is_bit_exact_against_its_definition(ops):
    x = nhwc(randn(2, 80, 9, 11, seed=1, scale=3.0))
    hi = tf32_rna(x)
    lo = tf32_rna(x - hi)
    # hi + lo reproduces x to 2^-22 relative; hi carries 10 mantissa bits
    assert float(((hi.double() + lo.double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max()) <= 2.0 ** -21
    m0 = ops.split3(x, 0)
    assert m0.shape == (2, 9, 11, 240)
    # cross terms first along the contraction axis (x: hi, lo, hi  against  w / dy: lo, hi, hi), hi*hi last
    assert torch.equal(m0[..., :80], hi) and torch.equal(m0[..., 80:160], lo) and torch.equal(m0[..., 160:], hi)
    m1, m2 = ops.split3(x, 1), ops.split3(x, 2)
    assert m1.shape == (6, 9, 11, 80)
    assert torch.equal(m1[:2], hi) and torch.equal(m1[2:4], lo) and torch.equal(m1[4:], hi)
    assert torch.equal(m2[:2], lo) and torch.equal(m2[2:4], hi) and torch.equal(m2[4:], hi)


def test_split_weight_packing_is_bit_exact(ops):
    w = randn(160, 80, 3, 3, seed=2, scale=0.05)
    fwd, dgr = ops.pack_conv_weights(w, round_tf32=2)
    plain_f, plain_d = ops.pack_conv_weights(w, round_tf32=False)
    assert fwd.shape == (9, 160, 240) and dgr.shape == (9, 80, 480)
    hi_f, hi_d = tf32_rna(plain_f), tf32_rna(plain_d)
    lo_f, lo_d = tf32_rna(plain_f - hi_f), tf32_rna(plain_d - hi_d)
    assert torch.equal(fwd[..., :80], lo_f) and torch.equal(fwd[..., 80:160], hi_f) and torch.equal(fwd[..., 160:], hi_f)
    assert torch.equal(dgr[..., :160], lo_d) and torch.equal(dgr[..., 160:320], hi_d) and torch.equal(dgr[..., 320:], hi_d)


@pytest.mark.parametrize("case", [(2, 19, 23, 80, 80), (1, 33, 17, 160, 160), (2, 48, 64, 160, 80), (1, 67, 90, 80, 160)])
def test_conv3x3_on_split_operands_reaches_fp32_accuracy(ops, case):
    B, H, W, Ci, Co = case
    x = randn(B, Ci, H, W, seed=1)
    w = randn(Co, Ci, 3, 3, seed=2, scale=(Ci * 9) ** -0.5)
    b = randn(Co, seed=3, scale=0.1)
    pre = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    w3, _ = ops.pack_conv_weights(w, round_tf32=2)
    r3 = ops.conv_forward(ops.split3(nhwc(x), 0), w3, math=1, bias=b, gelu=True, save_pre=True)
    w1, _ = ops.pack_conv_weights(w, round_tf32=True)
    r1 = ops.conv_forward(nhwc(x), w1, math=1, bias=b, gelu=True, save_pre=True)
    e3, e1 = max_err_rel(nchw(r3["pre"]), pre), max_err_rel(nchw(r1["pre"]), pre)
    print(f"conv {case}: max err / max|ref|  3xTF32 {e3:.2e} (L2 {rel_err(nchw(r3['pre']), pre):.2e})   TF32 {e1:.2e}")
    assert e3 <= FP32_MAX, e3
    assert max_err_rel(nchw(r3["out"]), F.gelu(pre)) <= FP32_MAX
    assert e3 < e1 / 20          # and it is the split that buys it


def test_conv_with_residual_slices_on_split_operands(ops):
    # block l2's second conv: 3x3 (160 -> 160) + 1x1 residual conv from the 80-channel block input as extra K slices
    B, H, W = 2, 21, 30
    a1, xin = randn(B, 160, H, W, seed=4), randn(B, 80, H, W, seed=5)
    w2, wr = randn(160, 160, 3, 3, seed=6, scale=0.03), randn(160, 80, 1, 1, seed=7, scale=0.1)
    bias = randn(160, seed=8, scale=0.1)
    ref = F.conv2d(a1.double(), w2.double(), bias.double(), padding=1) + F.conv2d(xin.double(), wr.double())
    w2p, _ = ops.pack_conv_weights(w2, round_tf32=2)
    wrp, _ = ops.pack_conv_weights(wr, round_tf32=2)
    r = ops.conv_forward(ops.split3(nhwc(a1), 0), w2p, math=1, bias=bias, in_res=ops.split3(nhwc(xin), 0), w_res=wrp[0])
    e = max_err_rel(nchw(r["out"]), ref)
    print(f"conv + residual slices, 3xTF32: max err / max|ref| {e:.2e}")
    assert e <= FP32_MAX, e


@pytest.mark.parametrize("case", [(2, 19, 23, 80, 80, 9), (3, 33, 17, 160, 160, 9), (2, 48, 64, 80, 160, 1)])
def test_weight_gradient_on_split_operands_reaches_fp32_accuracy(ops, case):
    B, H, W, Cx, Cy, ntaps = case
    x, dy = randn(B, Cx, H, W, seed=11), randn(B, Cy, H, W, seed=12)
    k = 3 if ntaps == 9 else 1
    w = torch.zeros(Cy, Cx, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x.double(), w, padding=k // 2), w, dy.double())
    dw3 = ops.conv_wgrad(ops.split3(nhwc(x), 1), ops.split3(nhwc(dy), 2), ntaps, math=1)
    dw1 = ops.conv_wgrad(nhwc(x), nhwc(dy), ntaps, math=1)
    e3, e1 = max_err_rel(dw3.reshape(ref.shape), ref), max_err_rel(dw1.reshape(ref.shape), ref)
    print(f"wgrad {case}: max err / max|ref|  3xTF32 {e3:.2e}   TF32 (unrounded operands, truncated by the MMA) {e1:.2e}")
    assert e3 <= FP32_MAX, e3
    assert e3 < e1 / 20


def test_plan_mode_is_rejected_by_single_operators(ops):
    from sinddm_b200._capi import MATH_TF32X3, SinddmError
    x = nhwc(randn(1, 80, 8, 8))
    w, _ = ops.pack_conv_weights(randn(80, 80, 3, 3, scale=0.05))
    with pytest.raises(SinddmError, match="plan mode"):
        ops.conv_forward(x, w, math=MATH_TF32X3)


def test_training_step_time_of_the_three_math_modes():
    """Not a bound, a record: one finest-scale (186x248) training step at batch 8 in each math mode (device time)."""
    from test_gpu_net import build
    from conftest import rs_tensor
    sizes = [(64, 48), (90, 67), (126, 94), (177, 133), (248, 186)]
    B = 8
    x = rs_tensor(5, (B, 3, 186, 248), 0.5).clamp(-1, 1).to(DEV)
    t = torch.randint(0, 100, (B,), generator=torch.Generator().manual_seed(1)).to(DEV)
    noise = rs_tensor(6, (B, 3, 186, 248)).to(DEV)
    times = {}
    for math in ("tf32", "tf32x3", "fp32"):
        net, dif = build(math, sizes=sizes, losses=[1.1, 0.78, 0.55, 0.39])
        for _ in range(2):
            net.zero_grad()
            dif.p_losses(x, t, 4, noise=noise, x_orig=x).backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            net.zero_grad()
            dif.p_losses(x, t, 4, noise=noise, x_orig=x).backward()
        e1.record()
        torch.cuda.synchronize()
        times[math] = e0.elapsed_time(e1) / 3
        del net, dif
        torch.cuda.empty_cache()
    print("forward + backward at 8x186x248, ms:", {k: round(v, 2) for k, v in times.items()})
    assert times["tf32"] < times["tf32x3"] < times["fp32"]
