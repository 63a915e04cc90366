"""SinDDMNet on sm_100a: the reference's module surface over the CUDA plan in libsinddm_b200.so.

Reference: SinDDM/models.py:34-151 (SinusoidalPosEmb, SinDDMConvBlock, SinDDMNet).

The module tree, parameter names, shapes and construction order are the reference's, so `state_dict()` keys,
checkpoint files and init-RNG streams line up.  The stock nn.Conv2d / nn.Linear objects
are *containers only*: `forward` never calls them.  It hands the 52 parameter pointers to
`sinddm_net_forward` (one C call: conditioning kernel, 4 x {depthwise, conv+GELU, conv+residual}, fused
final conv) and autograd sees a single Function whose backward is `sinddm_net_backward`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict

import torch
from torch import nn

from . import _capi
from ._capi import MATH_FP32, MATH_TF32, MATH_TF32X3, NUM_PARAMS, check
from .functions import default, exists

_MATH_NAMES = {"fp32": MATH_FP32, "tf32": MATH_TF32, "tf32x3": MATH_TF32X3}


def default_math() -> int:
    """TF32 tensor-core math unless SINDDM_MATH=tf32x3 (fp32-class results on the tensor cores: every operand split
    into two TF32 values, three MMAs per product) or SINDDM_MATH=fp32 (the CUDA-core twin, a test oracle)."""
    return _MATH_NAMES[os.environ.get("SINDDM_MATH", "tf32").lower()]


class SinusoidalPosEmb(nn.Module):
    """models.py:34-46.  Kept for API parity; the fused conditioning kernel computes the same embedding."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def frequencies(self, device):
        half = self.dim // 2
        step = math.log(10000) / (half - 1)
        return torch.exp(torch.arange(half, device=device) * -step)

    def forward(self, x):
        arg = x[:, None] * self.frequencies(x.device)[None, :]
        return torch.cat((arg.sin(), arg.cos()), dim=-1)


class SinDDMConvBlock(nn.Module):
    """Parameter container of one conv block, models.py:51-67 (same submodule names and creation order)."""

    def __init__(self, dim, dim_out, *, time_emb_dim=None, mult=1):
        super().__init__()
        if mult != 1:
            raise NotImplementedError("the CUDA plan implements mult=1 (the only value the reference uses)")
        self.mlp = nn.Sequential(nn.GELU(), nn.Linear(time_emb_dim, time_emb_dim)) if exists(time_emb_dim) else None
        self.time_reshape = nn.Conv2d(time_emb_dim, dim, 1)
        self.ds_conv = nn.Conv2d(dim, dim, 5, padding=2, groups=dim)
        self.net = nn.Sequential(
            nn.Conv2d(dim, dim_out * mult, 3, padding=1),
            nn.GELU(),
            nn.Conv2d(dim_out * mult, dim_out, 3, padding=1),
        )
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x, time_emb=None):
        raise RuntimeError("SinDDMConvBlock is a parameter container; run the whole SinDDMNet (CUDA plan)")


class _PlanHandle:
    """One (shape, mode) plan: C handle + the workspace tensor its TMA descriptors point into."""

    def __init__(self, lib, device, B, H, W, dim, channels, math_mode, training):
        self.lib = lib
        self.key = (B, H, W, training)
        nbytes = lib.sinddm_plan_workspace_bytes(B, H, W, dim, channels, math_mode, int(training))
        if nbytes == 0:
            raise _capi.SinddmError(f"no plan for shape B={B} H={H} W={W} dim={dim}")
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        base = self.workspace.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        handle = C.c_void_p()
        check(lib.sinddm_plan_create(C.byref(handle), B, H, W, dim, channels, math_mode, int(training),
                                     C.c_void_p(aligned), nbytes), "sinddm_plan_create")
        self.handle = handle
        self.generation = 0       # bumped by every forward; backward must match (activations live in the plan)
        self.packed_key = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.sinddm_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _Runtime:
    """Per-module CUDA state (plans, pointer arrays).  Deliberately NOT copied by deepcopy / pickling:
    `copy.deepcopy(model)` (trainer's EMA copy, trainer.py:100) gets a fresh, empty runtime."""

    MAX_PLANS = 24

    def __init__(self):
        self.plans = OrderedDict()
        self.freqs = {}
        self.weights_epoch = 0      # bumped when a kernel updated the parameters in place (fused optimizer step)
        self.grad_bucket = None     # flat fp32 tensor backward writes the 52 gradients into (fused optimizer step)
        self.param_offsets = None   # element offset of every parameter inside that bucket (parameter order)
        self.param_total = 0
        self.param_list = None      # cached list(self.parameters()) (the Parameter objects never change)

    def __deepcopy__(self, memo):
        return _Runtime()

    def __reduce__(self):
        return (_Runtime, ())

    def plan(self, lib, device, B, H, W, dim, channels, math_mode, training):
        key = (device.index, B, H, W, dim, math_mode, bool(training))
        p = self.plans.get(key)
        if p is None:
            while len(self.plans) >= self.MAX_PLANS:
                self.plans.popitem(last=False)
            p = _PlanHandle(lib, device, B, H, W, dim, channels, math_mode, training)
            self.plans[key] = p
        else:
            self.plans.move_to_end(key)
        return p


def _ptr_array(tensors):
    arr = (C.c_void_p * NUM_PARAMS)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _NetFunction(torch.autograd.Function):
    """out = SinDDMNet(x, t, scale); backward = gradients of all 52 parameters (none for x, t)."""

    @staticmethod
    def forward(ctx, net, x, time, scale, want_grad, *params):
        lib = _capi.load()
        B, ch, H, W = x.shape
        plan = net._runtime.plan(lib, x.device, B, H, W, net.dim, ch, net.math, want_grad)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        parr = _ptr_array(params)
        pkey = (net._runtime.weights_epoch, tuple((p.data_ptr(), p._version) for p in params))
        if plan.packed_key != pkey:
            check(lib.sinddm_net_pack_weights(plan.handle, parr, stream), "sinddm_net_pack_weights")
            plan.packed_key = pkey
        out = torch.empty_like(x)
        check(lib.sinddm_net_forward(plan.handle, parr, x.data_ptr(), time.data_ptr(), float(scale),
                                     net._freqs(x.device).data_ptr(), out.data_ptr(), stream), "sinddm_net_forward")
        plan.generation += 1
        if want_grad:
            ctx.plan = plan
            ctx.generation = plan.generation
            bucket = net._runtime.grad_bucket
            offsets = net._runtime.param_offsets
            if offsets is None or len(offsets) != len(params):
                offsets, total = [], 0
                for p in params:
                    offsets.append(total)
                    total += p.numel()
                net._runtime.param_offsets, net._runtime.param_total = offsets, total
            if bucket is not None and bucket.numel() != net._runtime.param_total:
                raise _capi.SinddmError("gradient bucket size does not match the parameters")
            ctx.grad_bucket = bucket
            ctx.param_offsets = offsets
            ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise _capi.SinddmError(
                "SinDDMNet backward after another forward on the same shape: the saved activations live in the "
                "plan workspace and were overwritten; run backward before the next forward of this shape")
        params = ctx.saved_tensors
        lib = _capi.load()
        dout = dout.contiguous()
        stream = torch.cuda.current_stream(dout.device).cuda_stream
        bucket = ctx.grad_bucket
        if bucket is not None:
            # fused optimizer step: the 52 gradients go straight into the flat bucket at fixed offsets (no tensor views:
            # this runs between the loss read-back and the first backward kernel, i.e. on the critical path of a step),
            # stay there for sinddm_fused_step, and autograd does not accumulate them into .grad
            offsets = ctx.param_offsets
            base = bucket.data_ptr()
            garr = (C.c_void_p * NUM_PARAMS)()
            for i, off in enumerate(offsets):
                garr[i] = base + 4 * off
            check(lib.sinddm_net_backward(plan.handle, _ptr_array(params), dout.data_ptr(), garr, stream),
                  "sinddm_net_backward")
            return (None,) * (5 + len(params))
        sizes = [p.numel() for p in params]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dout.device)
        grads = [v.view_as(p) for v, p in zip(flat.split(sizes), params)]
        check(lib.sinddm_net_backward(plan.handle, _ptr_array(params), dout.data_ptr(), _ptr_array(grads), stream),
              "sinddm_net_backward")
        return (None, None, None, None, None, *grads)


class SinDDMNet(nn.Module):
    """models.py:85-151.  Same constructor; `forward(x, time, scale)` runs the CUDA plan."""

    def __init__(self, dim, out_dim=None, channels=3, with_time_emb=True, multiscale=False, device=None, math=None):
        super().__init__()
        self.device = device
        self.channels = channels
        self.multiscale = multiscale
        self.dim = dim
        self.math = default_math() if math is None else (_MATH_NAMES[math] if isinstance(math, str) else int(math))

        if with_time_emb:
            time_dim = 32
            if multiscale:
                self.SinEmbTime = SinusoidalPosEmb(time_dim)
                self.SinEmbScale = SinusoidalPosEmb(time_dim)
                self.time_mlp = nn.Sequential(
                    nn.Linear(time_dim * 2, time_dim * 4), nn.GELU(), nn.Linear(time_dim * 4, time_dim))
            else:
                self.time_mlp = nn.Sequential(
                    SinusoidalPosEmb(time_dim), nn.Linear(time_dim, time_dim * 4), nn.GELU(),
                    nn.Linear(time_dim * 4, time_dim))
        else:
            time_dim = None
            self.time_mlp = None

        half_dim = int(dim / 2)
        self.l1 = SinDDMConvBlock(channels, half_dim, time_emb_dim=time_dim)
        self.l2 = SinDDMConvBlock(half_dim, dim, time_emb_dim=time_dim)
        self.l3 = SinDDMConvBlock(dim, dim, time_emb_dim=time_dim)
        self.l4 = SinDDMConvBlock(dim, half_dim, time_emb_dim=time_dim)

        out_dim = default(out_dim, channels)
        self.final_conv = nn.Sequential(nn.Conv2d(half_dim, out_dim, 1))
        self._out_dim = out_dim
        self._runtime = _Runtime()

    # -- runtime helpers ---------------------------------------------------------------------------
    def _freqs(self, device):
        f = self._runtime.freqs.get(device)
        if f is None:
            f = SinusoidalPosEmb(32).frequencies(device).float().contiguous()
            self._runtime.freqs[device] = f
        return f

    def mark_weights_updated(self):
        """Call after the parameters were modified outside autograd's version tracking (raw kernels): the packed
        convolution weights of every plan are rebuilt before the next forward."""
        self._runtime.weights_epoch += 1

    def set_grad_bucket(self, bucket):
        """Flat fp32 CUDA tensor (sum of parameter sizes) that backward fills instead of returning gradients to
        autograd; None restores the normal `.grad` path."""
        self._runtime.grad_bucket = bucket

    def _check_supported(self):
        # The reference itself only works with multiscale=True (`if exists(self.multiscale)` is always true,
        # models.py:136, and SinEmbTime only exists in that branch); main.py:80 always passes True.
        if not self.multiscale or self.time_mlp is None:
            raise NotImplementedError("SinDDMNet CUDA plan needs with_time_emb=True and multiscale=True "
                                      "(the only configuration the reference can run)")
        if self.channels != 3 or self._out_dim != 3:
            raise NotImplementedError("SinDDMNet CUDA plan supports channels = out_dim = 3")

    def forward(self, x, time, scale=None):
        self._check_supported()
        if not x.is_cuda:
            raise _capi.SinddmError("SinDDMNet.forward needs CUDA tensors: sinddm_b200 has no CPU fallback")
        _capi.init(x.device.index if x.device.index is not None else torch.cuda.current_device())
        params = self._runtime.param_list
        if params is not None and not (params[0] is self.time_mlp[0].weight and params[-1] is self.final_conv[0].bias):
            params = None               # a Parameter object was replaced: rebuild the cached list
        if params is None:
            params = list(self.parameters())
            if len(params) != NUM_PARAMS:
                raise _capi.SinddmError(f"expected {NUM_PARAMS} parameter tensors, found {len(params)}")
            self._runtime.param_list = params
        x = x.contiguous().float()
        time = time.contiguous().to(torch.int64)
        scale_val = float(scale.item()) if torch.is_tensor(scale) else float(scale)
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _NetFunction.apply(self, x, time, scale_val, want_grad, *params)
