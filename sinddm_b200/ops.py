"""Python entry points for the single CUDA operators of libsinddm_b200.so.

Thin wrappers: allocate outputs / workspaces as torch tensors, pass device pointers and the current stream
through the C ABI.  Used by the diffusion module (qsample_mix, l1_loss, ddpm_step) and by the parity tests
(the conv / depthwise / gradient operators are otherwise driven from C through the network plan).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from ._capi import MATH_FP32, MATH_TF32, ConvDesc, DdpmStepDesc, check, ptr


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _prep(*tensors):
    _capi.require_cuda(*tensors)
    dev = next(t for t in tensors if t is not None).device
    _capi.init(dev.index if dev.index is not None else torch.cuda.current_device())
    for t in tensors:
        if t is not None and not t.is_contiguous():
            raise _capi.SinddmError("sinddm_b200 operators need contiguous tensors")
    return _capi.load()


def _bytes_ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=like.device)


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------

def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    lib = _prep(x)
    B, Cc, H, W = x.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=x.device)
    check(lib.sinddm_nchw_to_nhwc(ptr(x), ptr(out), B, Cc, H, W, _stream(x)), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    lib = _prep(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib.sinddm_nhwc_to_nchw(ptr(x), ptr(out), B, Cc, H, W, _stream(x)), "nhwc_to_nchw")
    return out


# ------------------------------------------------------------------------------------------------
# dense convolution (NHWC) and its gradients
# ------------------------------------------------------------------------------------------------

def split3(x: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """3xTF32 operand split of an NHWC tensor [B,H,W,C]: mode 0 -> [B,H,W,3C] = [hi | lo | hi] (conv operand),
    mode 1 -> [3B,H,W,C] = hi, lo, hi (weight-gradient x operand), mode 2 -> [3B,H,W,C] = lo, hi, hi (its dy operand)."""
    lib = _prep(x)
    B, H, W, Cc = x.shape
    shape = (B, H, W, 3 * Cc) if mode == 0 else (3 * B, H, W, Cc)
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    check(lib.sinddm_split3(ptr(x), B * H * W, Cc, ptr(out), int(mode), _stream(x)), "split3")
    return out


def pack_conv_weights(w: torch.Tensor, round_tf32=False):
    """OIHW weight -> (forward operand [tap][Cout][Cin], data-gradient operand [tap][Cin][Cout]); round_tf32 = 2: the
    3xTF32 split layouts [tap][Cout][3 Cin] / [tap][Cin][3 Cout] ([lo | hi | hi] along the contraction axis)."""
    lib = _prep(w)
    co, ci, kh, kw = w.shape
    ntaps = kh * kw
    k3 = 3 if int(round_tf32) == 2 else 1
    fwd = torch.empty((ntaps, co, k3 * ci), dtype=torch.float32, device=w.device)
    dgr = torch.empty((ntaps, ci, k3 * co), dtype=torch.float32, device=w.device)
    check(lib.sinddm_pack_conv_weights(ptr(w), co, ci, ntaps, ptr(fwd), ptr(dgr), int(round_tf32), _stream(w)),
          "pack_conv_weights")
    return fwd, dgr


def conv_forward(x, w_packed, *, math=MATH_TF32, bias=None, gelu=False, res_add=None, in_res=None, w_res=None,
                 x3=None, w_res3=None, dgelu_z=None, w_final=None, b_final=None, save_pre=False,
                 round_tf32=False, want_out=True):
    """x [B,H,W,Cin] NHWC, w_packed [ntaps][N][Cin] -> dict(out=[B,H,W,N], pre=..., final=[B,3,H,W])."""
    lib = _prep(x, w_packed, bias, res_add, in_res, w_res, x3, w_res3, dgelu_z, w_final, b_final)
    B, H, W, Cin = x.shape
    ntaps, N, cin2 = w_packed.shape
    assert cin2 == Cin
    d = ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.inp, d.Cin, d.w, d.ntaps, d.N = ptr(x), Cin, ptr(w_packed), ntaps, N
    if in_res is not None:
        d.in_res, d.Cres, d.w_res = ptr(in_res), in_res.shape[-1], ptr(w_res)
    d.bias, d.res_add, d.x3, d.w_res3 = ptr(bias), ptr(res_add), ptr(x3), ptr(w_res3)
    d.gelu = int(gelu)
    res = {}
    if save_pre:
        res["pre"] = torch.empty((B, H, W, N), dtype=torch.float32, device=x.device)
        d.out_pre = ptr(res["pre"])
    d.dgelu_z = ptr(dgelu_z)
    if w_final is not None:
        res["final"] = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
        d.w_final, d.b_final, d.out_final = ptr(w_final), ptr(b_final), ptr(res["final"])
    d.round_tf32 = int(round_tf32)
    if want_out:
        res["out"] = torch.empty((B, H, W, N), dtype=torch.float32, device=x.device)
        d.out = ptr(res["out"])
    check(lib.sinddm_conv_forward(C.byref(d), int(math), _stream(x)), "conv_forward")
    return res


def conv_wgrad(x, dy, ntaps: int, *, math=MATH_TF32) -> torch.Tensor:
    """x [B,H,W,Cx], dy [B,H,W,Cy] -> dW [Cy, Cx, k, k] (PyTorch layout)."""
    lib = _prep(x, dy)
    B, H, W, Cx = x.shape
    Cy = dy.shape[-1]
    k = 3 if ntaps == 9 else 1
    dw = torch.empty((Cy, Cx, k, k), dtype=torch.float32, device=x.device)
    nbytes = lib.sinddm_conv_wgrad_workspace_bytes(B, H, W, Cx, Cy, ntaps, int(math))
    ws = _bytes_ws(nbytes, x)
    check(lib.sinddm_conv_wgrad(ptr(x), Cx, ptr(dy), Cy, B, H, W, ntaps, ptr(dw), ptr(ws), ws.numel(), int(math),
                                _stream(x)), "conv_wgrad")
    return dw


# ------------------------------------------------------------------------------------------------
# depthwise 5x5, column sums
# ------------------------------------------------------------------------------------------------

def dw5x5(x, w, bias=None, cond=None, add=None, *, flip=False, round_tf32=False) -> torch.Tensor:
    """x [B,H,W,C]; w [C,1,5,5] (or [C,25]); cond [B,C]."""
    lib = _prep(x, w, bias, cond, add)
    B, H, W, Cc = x.shape
    out = torch.empty_like(x)
    check(lib.sinddm_dw5x5(ptr(x), ptr(w), ptr(bias), ptr(cond), ptr(add), ptr(out), B, H, W, Cc, int(flip),
                           int(round_tf32), _stream(x)), "dw5x5")
    return out


def dw5x5_wgrad(x, dh):
    lib = _prep(x, dh)
    B, H, W, Cc = x.shape
    dw = torch.empty((Cc, 1, 5, 5), dtype=torch.float32, device=x.device)
    db = torch.empty((Cc,), dtype=torch.float32, device=x.device)
    dcond = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
    ws = _bytes_ws(lib.sinddm_dw5x5_wgrad_workspace_bytes(B, H, Cc), x)
    check(lib.sinddm_dw5x5_wgrad(ptr(x), ptr(dh), ptr(dw), ptr(db), ptr(dcond), ptr(ws), ws.numel(), B, H, W, Cc,
                                 _stream(x)), "dw5x5_wgrad")
    return dw, db, dcond


def colsum(a: torch.Tensor) -> torch.Tensor:
    """a [..., C] -> [C] sum over all leading dims."""
    lib = _prep(a)
    Cc = a.shape[-1]
    P = a.numel() // Cc
    out = torch.empty((Cc,), dtype=torch.float32, device=a.device)
    ws = _bytes_ws(lib.sinddm_colsum_workspace_bytes(Cc), a)
    check(lib.sinddm_colsum(ptr(a), P, Cc, ptr(out), ptr(ws), ws.numel(), _stream(a)), "colsum")
    return out


# ------------------------------------------------------------------------------------------------
# diffusion arithmetic
# ------------------------------------------------------------------------------------------------

def qsample_mix(x_start, noise, t, sqrt_ac, sqrt_1mac, x_orig=None, gammas_row=None) -> torch.Tensor:
    lib = _prep(x_start, noise, t, sqrt_ac, sqrt_1mac, x_orig, gammas_row)
    assert t.dtype == torch.int64
    B = x_start.shape[0]
    out = torch.empty_like(x_start)
    check(lib.sinddm_qsample_mix(ptr(x_start), ptr(x_orig), ptr(noise), ptr(t), ptr(sqrt_ac), ptr(sqrt_1mac),
                                 ptr(gammas_row), ptr(out), B, x_start.numel() // B, _stream(x_start)),
          "qsample_mix")
    return out


def l1_loss(noise, pred, want_grad: bool):
    """-> (loss [] , dpred or None) with dpred = d mean|noise-pred| / d pred."""
    lib = _prep(noise, pred)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pred) if want_grad else None
    ws = _bytes_ws(lib.sinddm_l1_loss_workspace_bytes(), pred)
    check(lib.sinddm_l1_loss(ptr(noise), ptr(pred), pred.numel(), ptr(loss), ptr(dpred), ptr(ws), ws.numel(),
                             _stream(pred)), "l1_loss")
    return loss, dpred


def ddpm_step(x_t, eps, noise, t, tables: dict, *, x_tilde=None, gammas_row=None, reblur_mode=False,
              clip_denoised=True, omega=0.0) -> torch.Tensor:
    lib = _prep(x_t, eps, noise, t, x_tilde, gammas_row, *tables.values())
    assert t.dtype == torch.int64
    d = DdpmStepDesc()
    out = torch.empty_like(x_t)
    B = x_t.shape[0]
    d.x_t, d.eps, d.x_tilde, d.noise, d.t, d.out = ptr(x_t), ptr(eps), ptr(x_tilde), ptr(noise), ptr(t), ptr(out)
    d.B, d.per_sample = B, x_t.numel() // B
    d.reblur_mode, d.clip_denoised, d.omega = int(reblur_mode), int(clip_denoised), float(omega)
    for name in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                 "posterior_mean_coef2", "posterior_log_variance_clipped", "alphas_cumprod", "sqrt_alphas_cumprod",
                 "sqrt_one_minus_alphas_cumprod"):
        setattr(d, name, ptr(tables[name]))
    d.gammas = ptr(gammas_row)
    check(lib.sinddm_ddpm_step(C.byref(d), _stream(x_t)), "ddpm_step")
    return out


def randn_rows(shape, rank: int, world: int, device) -> torch.Tensor:
    """Rows [rank*b, (rank+1)*b) of `torch.randn((b*world, *shape[1:]), device=device)` WITHOUT generating the other
    ranks' rows: the shard is produced by a kernel that replays torch's element <-> Philox-counter mapping
    (sinddm_philox_normal_rows), and the device generator's offset is advanced exactly as the global draw would have.
    `shape` is the LOCAL shape (b, ...).  Values and generator state are bit-identical to draw-everything-and-slice
    (tests/test_gpu_ops.py), so N ranks still consume the random numbers one GPU would."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _capi.SinddmError("randn_rows needs a CUDA device")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    _capi.init(index)
    lib = _capi.load()
    b = int(shape[0])
    per_row = 1
    for d in shape[1:]:
        per_row *= int(d)
    count = b * per_row
    numel = count * int(world)
    props = torch.cuda.get_device_properties(index)
    # ATen/native/cuda/DistributionTemplates.h::calc_execution_policy (block 256, unroll 4)
    grid = min(props.multi_processor_count * (props.max_threads_per_multi_processor // 256), (numel + 255) // 256)
    stride = 256 * grid
    increment = ((numel - 1) // (stride * 4) + 1) * 4
    gen = torch.cuda.default_generators[index]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    out = torch.empty(tuple(int(d) for d in shape), dtype=torch.float32, device=dev)
    check(lib.sinddm_philox_normal_rows(ptr(out), int(rank) * count, count, stride, seed, offset,
                                        torch.cuda.current_stream(dev).cuda_stream), "philox_normal_rows")
    gen.set_offset(offset + increment)
    return out
