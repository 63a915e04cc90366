"""Fused optimizer step: gradient all-reduce over NVLink peer memory + Adam + EMA in one kernel.

Replaces the tail of one MultiscaleTrainer step (reference SinDDM/trainer.py:208-213 -- `opt.step()`,
`opt.zero_grad()`, `step_ema()` -> models.py:18-31 -- plus, under torchrun, the NCCL all-reduce of the gradient
bucket that the data-parallel path added): `sinddm_fused_step` (csrc/fused_optim.cu) does all of it in ONE
launch.  The host side here only owns the memory:

  * the parameters of the trained denoiser and of its EMA copy are re-pointed at two flat fp32 buffers (each
    nn.Parameter becomes a view, so state_dict / checkpoints / the weight-packing kernels see no difference);
  * the gradient bucket is the buffer `sinddm_net_backward` writes its 52 gradients into (no autograd
    accumulation, no copy-in); under torchrun it lives in a symmetric allocation
    (torch.distributed._symmetric_memory: allocation + handle exchange only) so that every rank's bucket and
    flag array are mapped into every other rank's address space; two buckets alternate by step parity.

Adam state (exp_avg / exp_avg_sq / step count) starts at zero like a fresh torch.optim.Adam; the reference does
not checkpoint optimizer state either (quirk Q12).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from . import _capi
from ._capi import FusedStepDesc, check


def _flatten_params(params: List[torch.nn.Parameter], pad_to: int) -> torch.Tensor:
    """Moves the parameters into one flat fp32 buffer (values preserved) and makes each a view of it."""
    sizes = [p.numel() for p in params]
    n = sum(sizes)
    npad = (n + pad_to - 1) // pad_to * pad_to
    flat = torch.zeros(npad, dtype=torch.float32, device=params[0].device)
    off = 0
    with torch.no_grad():
        for p, k in zip(params, sizes):
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            off += k
    return flat


class FusedStep:
    """Owner of the flat parameter / moment / EMA buffers and the (symmetric) gradient buckets."""

    def __init__(self, net, ema_net, *, betas=(0.9, 0.999), eps=1e-8, group=None):
        import torch.distributed as dist
        self.net, self.ema_net = net, ema_net
        self.params = [p for p in net.parameters()]
        self.ema_params = [p for p in ema_net.parameters()] if ema_net is not None else None
        if not self.params or not self.params[0].is_cuda:
            raise _capi.SinddmError("FusedStep needs CUDA parameters: sinddm_b200 has no CPU fallback")
        if any(p.dtype != torch.float32 or not p.requires_grad for p in self.params):
            raise _capi.SinddmError("FusedStep handles trainable fp32 parameters only")
        dev = self.params[0].device
        _capi.init(dev.index if dev.index is not None else torch.cuda.current_device())
        self.device = dev
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.n = sum(p.numel() for p in self.params)
        self.flat_param = _flatten_params(self.params, 4)
        self.npad = self.flat_param.numel()
        self.flat_ema = _flatten_params(self.ema_params, 4) if self.ema_params is not None else None
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.t = 0            # Adam step count
        self.epoch = 0        # barrier epoch (== number of fused steps issued)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > _capi.FUSED_MAX_WORLD:
            raise _capi.SinddmError(f"FusedStep supports up to {_capi.FUSED_MAX_WORLD} ranks of one NVSwitch box")
        # symmetric buffer: [bucket 0 | bucket 1 | 64 uint32 flags]
        self._flag_off = 2 * self.npad
        total = 2 * self.npad + 64
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.symm = symm.empty(total, dtype=torch.float32, device=dev)
            self.symm.zero_()
            torch.cuda.synchronize(dev)
            self.handle = symm.rendezvous(self.symm, dist.group.WORLD if group is None else group)
            self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            # NVLink-SHARP (multimem.ld_reduce) gradient sum inside the switch: one request per 16 bytes instead of
            # `world` peer loads (8 ranks: 0.137 -> 0.068 ms per step; replicas stayed bit-identical over 1006 steps x
            # 1.9 M elements x 8 ranks, profiles/r02_fused_dp_multimem.txt).  Default: on where the symmetric
            # allocation has a multicast mapping and world >= 4 (no gain at 2); SINDDM_FUSED_MULTIMEM=0/1 overrides.
            # The summation order is then the switch's (NCCL's NVLS all-reduce agrees to 3e-8), not rank order; the
            # trainer's _check_replicas() guards the replicas at every milestone either way.
            mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
            want = os.environ.get("SINDDM_FUSED_MULTIMEM", "")
            use = (want == "1") if want in ("0", "1") else self.world >= 4
            self.mc_ptr = mc if use else 0
            dist.barrier(group)       # every rank's flags are zero before anyone signals
        else:
            self.symm = torch.zeros(total, dtype=torch.float32, device=dev)
            self.handle = None
            self.peer_ptrs = [self.symm.data_ptr()]
            self.mc_ptr = 0
        self._sizes = [p.numel() for p in self.params]
        # [0]: nanoseconds spent waiting for peers in the in-kernel barrier, summed over steps; [1]: the largest wait
        self.wait_ns = torch.zeros(2, dtype=torch.int64, device=dev) if self.world > 1 else None
        for m in (net, ema_net):
            if m is not None and hasattr(m, "mark_weights_updated"):
                m.mark_weights_updated()

    # the bucket sinddm_net_backward writes this step's gradients into
    def bucket(self) -> torch.Tensor:
        parity = (self.epoch & 1)
        return self.symm[parity * self.npad: parity * self.npad + self.n]

    def grad_views(self) -> List[torch.Tensor]:
        """Per-parameter views of the current bucket (tests / debugging)."""
        return [v.view_as(p) for v, p in zip(self.bucket().split(self._sizes), self.params)]

    def barrier_wait_ms(self, reset: bool = True):
        """(total, max) milliseconds this rank spent in the fused step's NVLink barrier waiting for slower peers."""
        if self.wait_ns is None:
            return 0.0, 0.0
        tot, mx = (float(v) * 1e-6 for v in self.wait_ns.tolist())
        if reset:
            self.wait_ns.zero_()
        return tot, mx

    def step(self, lr: float, ema_mode: int = 0, ema_beta: float = 0.995) -> None:
        """Consumes the current bucket: mean over ranks, Adam, EMA.  Switches to the other bucket."""
        lib = _capi.load()
        parity = self.epoch & 1
        self.epoch += 1
        self.t += 1
        d = FusedStepDesc()
        d.world, d.rank, d.n = self.world, self.rank, self.npad
        for r in range(self.world):
            d.grads[r] = self.peer_ptrs[r] + 4 * parity * self.npad
            d.flags[r] = self.peer_ptrs[r] + 4 * self._flag_off
        d.epoch = self.epoch & 0xFFFFFFFF
        d.param, d.exp_avg, d.exp_avg_sq = self.flat_param.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        d.ema = self.flat_ema.data_ptr() if self.flat_ema is not None else None
        d.lr, d.beta1, d.beta2, d.eps = float(lr), self.betas[0], self.betas[1], self.eps
        d.step = self.t
        d.ema_mode = int(ema_mode) if self.flat_ema is not None else 0
        d.ema_beta = float(ema_beta)
        d.wait_ns = self.wait_ns.data_ptr() if self.wait_ns is not None else None
        d.mc_grads = (self.mc_ptr + 4 * parity * self.npad) if self.mc_ptr else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.sinddm_fused_step(C.byref(d), stream), "sinddm_fused_step")
        # the kernels wrote the parameters behind autograd's back: invalidate the packed conv weights
        self.net.mark_weights_updated()
        if d.ema_mode and self.ema_net is not None:
            self.ema_net.mark_weights_updated()
