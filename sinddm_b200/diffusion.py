"""MultiScaleGaussianDiffusion on sm_100a (reference: SinDDM/models.py:18-31 EMA, :155-631 diffusion).

Same constructor, buffers, attributes and public methods as the reference, so main.py / the trainer and
the authors' checkpoints (13 buffers + denoise_fn.* keys) drop in.  What changed underneath:

  * training (`forward` -> `p_losses`): blur-mix + q_sample is one kernel, the denoiser is one fused
    forward/backward plan, the L1 loss and its gradient are one kernel;
  * sampling (`p_sample`): the ~40 element-wise launches and two host syncs per step of
    predict_start_from_noise / p_mean_variance / q_posterior collapse into `ddpm_step`; the branch the
    reference takes by reading `t[0]` on the host is taken on the device;
  * RNG: every torch.randint / torch.randn call is kept, in the reference's order and shapes (quirk Q7), so
    a fixed seed gives the reference's stream.

CLIP guidance and ROI-guided sampling are out of scope; their attributes
exist so external code can poke them, enabling them raises.
"""
from __future__ import annotations

import os
import warnings
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from tqdm import tqdm

from . import _capi, ops
from .functions import cosine_beta_schedule, default, exists, extract, noise_like

_TABLES = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
           "posterior_mean_coef2", "posterior_log_variance_clipped", "alphas_cumprod", "sqrt_alphas_cumprod",
           "sqrt_one_minus_alphas_cumprod")


class EMA:
    """models.py:18-31: ma = ma * beta + (1 - beta) * current, parameters only (quirk Q8)."""

    def __init__(self, beta):
        self.beta = beta

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new

    def update_model_average(self, ma_model, current_model):
        for cur, ma in zip(current_model.parameters(), ma_model.parameters()):
            ma.data = self.update_average(ma.data, cur.data)
        # rebinding `.data` changes neither the parameter's version counter nor, reliably, its address (the caching
        # allocator recycles blocks): tell the CUDA plans that their packed copies of these weights are stale
        net = getattr(ma_model, 'denoise_fn', ma_model)
        if hasattr(net, 'mark_weights_updated'):
            net.mark_weights_updated()


class _L1LossFn(torch.autograd.Function):
    """mean |noise - pred| with the gradient produced by the same kernel pass (models.py:594)."""

    @staticmethod
    def forward(ctx, noise, pred):
        loss, dpred = ops.l1_loss(noise.contiguous(), pred.contiguous(), want_grad=ctx.needs_input_grad[1])
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, gout):
        if ctx.dpred is None:
            return None, None
        return None, ctx.dpred * gout


class _StepGraph:
    """One reverse-diffusion step (denoiser evaluation + noise draw + ddpm_step + `t -= 1`) captured as a CUDA graph
    over static tensors: replaying it n times runs n consecutive timesteps with one host call each instead of
    ~35 kernel launches and a dozen tensor allocations (the coarse scales are launch-bound otherwise)."""

    def __init__(self):
        self.graph = None
        self.x = None          # state, updated in place by every replay
        self.t = None          # [B] int64, decremented by every replay
        self.prev = None       # upsampled previous-scale sample (reblurring), or None
        self.launches = 0      # library kernels per replay


# library kernels executed through graph replays (they bypass the C launch counter, which counts at capture time)
graph_replayed_launches = 0


class MultiScaleGaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, *, save_interm=False, results_folder='/Results', n_scales, scale_factor,
                 image_sizes, scale_mul=(1, 1), channels=3, timesteps=100, train_full_t=False, scale_losses=None,
                 loss_factor=1, loss_type='l1', betas=None, device=None, reblurring=True, sample_limited_t=False,
                 omega=0):
        super().__init__()
        self.device = device
        self.save_interm = save_interm
        self.results_folder = Path(results_folder)
        self.channels = channels
        self.n_scales = n_scales
        self.scale_factor = scale_factor
        self.scale_mul = scale_mul
        self.sample_limited_t = sample_limited_t
        self.reblurring = reblurring
        self.img_prev_upsample = None
        self.omega = omega
        self.loss_type = loss_type
        # sampling loops replay a captured CUDA graph per timestep (SINDDM_SAMPLE_GRAPH=0 keeps the eager loop)
        self.use_step_graph = os.environ.get('SINDDM_SAMPLE_GRAPH', '1') != '0'

        # guidance hooks of the reference (models.py:193-220): present, inert, refusing to be switched on
        self.clip_guided_sampling = False
        self.guidance_sub_iters = None
        self.stop_guidance = None
        self.quantile = 0.8
        self.clip_model = None
        self.clip_strength = None
        self.clip_text = ''
        self.text_embedds = None
        self.text_embedds_hr = None
        self.text_embedds_lr = None
        self.clip_text_features = None
        self.clip_score = []
        self.clip_mask = None
        self.llambda = 0
        self.x_recon_prev = None
        self.clip_roi_bb = []
        self.roi_guided_sampling = False
        self.roi_bbs = []
        self.roi_bbs_stat = []
        self.roi_target_patch = []

        # sizes arrive (W, H) and are stored (H, W) (quirk Q2, models.py:222-223)
        self.image_sizes = tuple((image_sizes[i][1], image_sizes[i][0]) for i in range(n_scales))
        self.denoise_fn = denoise_fn

        if exists(betas):
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else betas
        else:
            betas = cosine_beta_schedule(timesteps)
        alphas = 1. - betas
        alpha_bar = np.cumprod(alphas, axis=0)
        alpha_bar_prev = np.append(1., alpha_bar[:-1])
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)

        as_f32 = partial(torch.tensor, dtype=torch.float32)
        post_var = betas * (1. - alpha_bar_prev) / (1. - alpha_bar)
        # registration order = state_dict order of the reference (models.py:247-267)
        self.register_buffer('betas', as_f32(betas))
        self.register_buffer('alphas_cumprod', as_f32(alpha_bar))
        self.register_buffer('alphas_cumprod_prev', as_f32(alpha_bar_prev))
        self.register_buffer('sqrt_alphas_cumprod', as_f32(np.sqrt(alpha_bar)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', as_f32(np.sqrt(1. - alpha_bar)))
        self.register_buffer('log_one_minus_alphas_cumprod', as_f32(np.log(1. - alpha_bar)))
        self.register_buffer('sqrt_recip_alphas_cumprod', as_f32(np.sqrt(1. / alpha_bar)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', as_f32(np.sqrt(1. / alpha_bar - 1)))
        self.register_buffer('posterior_variance', as_f32(post_var))
        self.register_buffer('posterior_log_variance_clipped', as_f32(np.log(np.maximum(post_var, 1e-20))))
        self.register_buffer('posterior_mean_coef1', as_f32(betas * np.sqrt(alpha_bar_prev) / (1. - alpha_bar)))
        self.register_buffer('posterior_mean_coef2',
                             as_f32((1. - alpha_bar_prev) * np.sqrt(alphas) / (1. - alpha_bar)))

        # per-scale starting timestep and blur schedule (models.py:269-287)
        sigma_t = np.sqrt(1. - alpha_bar) / np.sqrt(alpha_bar)
        self.num_timesteps_trained = [self.num_timesteps]
        self.num_timesteps_ideal = [self.num_timesteps]
        if scale_losses is not None:
            for i in range(n_scales - 1):
                self.num_timesteps_ideal.append(int(np.argmax(sigma_t > loss_factor * scale_losses[i])))
                self.num_timesteps_trained.append(int(timesteps) if train_full_t else self.num_timesteps_ideal[i + 1])
        gammas = torch.zeros(size=(n_scales - 1, self.num_timesteps), device=self.device)
        for i in range(n_scales - 1):
            gammas[i, :] = (torch.tensor(sigma_t, device=self.device) / (loss_factor * scale_losses[i])).clamp(min=0, max=1)
        self.register_buffer('gammas', gammas)

    # ------------------------------------------------------------------------------------------
    # small reference methods kept as tensor code (not on the hot path)
    # ------------------------------------------------------------------------------------------
    def q_mean_variance(self, x_start, t):
        """models.py:300-304"""
        mean = extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = extract(1. - self.alphas_cumprod, t, x_start.shape)
        log_variance = extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def _gamma_row(self, s):
        """Row of `gammas` used at scale s > 0 (contiguous view; clamping, where the reference clamps, is
        done inside the kernels so in-place edits of the buffer (trainer.py:327) stay visible)."""
        return self.gammas[int(s) - 1].contiguous()

    def _tables(self):
        return {name: getattr(self, name) for name in _TABLES}

    # -- data-parallel RNG: every rank draws the GLOBAL batch's stream and keeps its shard, so an N-GPU run
    #    consumes exactly the random numbers the 1-GPU run would.  dp_world == 1: plain draws.
    dp_rank = 0
    dp_world = 1
    dp_shard_rng = os.environ.get('SINDDM_DP_SHARD_RNG', '1') != '0'

    def set_data_parallel(self, rank, world):
        self.dp_rank, self.dp_world = int(rank), int(world)

    def _shard(self, full, local_b):
        return full[self.dp_rank * local_b:(self.dp_rank + 1) * local_b].contiguous()

    def _randn(self, shape, device):
        if self.dp_world == 1:
            return torch.randn(shape, device=device)
        dev = torch.device(device)
        numel = int(np.prod(shape)) * self.dp_world
        if (dev.type == 'cuda' and self.dp_shard_rng and numel >= 16 and numel < 2 ** 31
                and not torch.cuda.is_current_stream_capturing()):
            # only this rank's rows of the global draw, same values, same generator advance (ops.randn_rows)
            return ops.randn_rows(tuple(shape), self.dp_rank, self.dp_world, dev)
        # CPU (gloo tests), CUDA-graph capture (the offset must come from the graph's generator state), tiny draws:
        # draw the global batch and keep the shard
        full = torch.randn((shape[0] * self.dp_world, *shape[1:]), device=device)
        return self._shard(full, shape[0])

    def _refuse_guidance(self):
        if self.clip_guided_sampling or self.roi_guided_sampling:
            raise NotImplementedError("CLIP / ROI guided sampling is outside the sinddm_b200 hot path "
                                      "(SURVEY.md section 2, rows 8-9)")

    # ------------------------------------------------------------------------------------------
    # forward diffusion
    # ------------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        """models.py:570-576 via the qsample_mix kernel (no blur mix)."""
        noise = default(noise, lambda: self._randn(x_start.shape, x_start.device))
        t = t.expand(x_start.shape[0]).contiguous().to(torch.int64)   # sample_via_scale passes an expanded scalar
        return ops.qsample_mix(x_start.contiguous(), noise.contiguous(), t,
                               self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)

    def p_losses(self, x_start, t, s, noise=None, x_orig=None):
        """models.py:578-611: noise draw -> (blur mix) -> q_sample -> denoiser -> loss."""
        noise = default(noise, lambda: self._randn(x_start.shape, x_start.device))
        x_start = x_start.contiguous()
        noise = noise.contiguous()
        s = int(s)
        if s > 0:
            x_noisy = ops.qsample_mix(x_start, noise, t, self.sqrt_alphas_cumprod,
                                      self.sqrt_one_minus_alphas_cumprod, x_orig=x_orig.contiguous(),
                                      gammas_row=self._gamma_row(s))       # gamma unclamped here (Q3)
        else:
            x_noisy = ops.qsample_mix(x_start, noise, t, self.sqrt_alphas_cumprod,
                                      self.sqrt_one_minus_alphas_cumprod)
        x_recon = self.denoise_fn(x_noisy, t, s)

        if self.loss_type == 'l1':
            return _L1LossFn.apply(noise, x_recon)
        if self.loss_type == 'l2':
            return F.mse_loss(noise, x_recon)
        if self.loss_type == 'l1_pred_img':
            if s > 0:
                cur_gammas = self.gammas[s - 1].reshape(-1)
                if t[0] > 0:
                    x_mix_prev = extract(cur_gammas, t - 1, x_start.shape) * x_start + \
                        (1 - extract(cur_gammas, t - 1, x_start.shape)) * x_orig
                else:
                    x_mix_prev = x_orig
            else:
                x_mix_prev = x_start
            return (x_mix_prev - x_recon).abs().mean()
        raise NotImplementedError()

    def forward(self, x, s, *args, **kwargs):
        """models.py:613-631: x = (orig batch, blurry batch); draws t then calls p_losses."""
        s = int(s)
        x_orig = x[0]
        b, c, h, w = x_orig.shape
        img_size = self.image_sizes[s]
        assert h == img_size[0] and w == img_size[1], f'height and width of image must be {img_size}'
        if not 0 < self.num_timesteps_trained[s] <= self.num_timesteps:
            raise IndexError(f'num_timesteps_trained[{s}] = {self.num_timesteps_trained[s]} is outside the schedule')
        t = torch.randint(0, self.num_timesteps_trained[s], (b * self.dp_world,), device=x_orig.device).long()
        if self.dp_world > 1:
            t = self._shard(t, b)
        if s > 0:
            return self.p_losses(x[1], t, s, x_orig=x_orig, *args, **kwargs)
        return self.p_losses(x_orig, t, s, *args, **kwargs)

    # ------------------------------------------------------------------------------------------
    # reverse diffusion
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def p_sample(self, x, t, s, clip_denoised=True, repeat_noise=False):
        """models.py:449-459 with p_mean_variance / predict_start_from_noise / q_posterior fused in ddpm_step.
        The denoiser runs grad-free (the reference's @enable_grad only serves CLIP guidance, quirk Q6)."""
        self._refuse_guidance()
        s = int(s)
        x = x.contiguous()
        eps = self.denoise_fn(x, t, scale=s)
        # drawn every step, also when unused (Q7)
        noise = noise_like(x.shape, x.device, True) if repeat_noise else self._randn(x.shape, x.device)
        reblur = bool(self.reblurring) and s > 0
        return ops.ddpm_step(x, eps, noise, t, self._tables(),
                             x_tilde=self.img_prev_upsample.contiguous() if reblur else None,
                             gammas_row=self._gamma_row(s) if reblur else None,
                             reblur_mode=reblur, clip_denoised=clip_denoised, omega=float(self.omega))

    # -- CUDA-graph replay of consecutive timesteps -------------------------------------------------
    def _step_graph_for(self, x, s, reblur):
        """Graph of one p_sample step for this (scale, shape); lives with the denoiser's inference plan so it
        dies with the workspace its kernels point into.  Returns None when capture is not possible."""
        net = self.denoise_fn
        lib = _capi.load()
        B, ch, H, W = x.shape
        plan = net._runtime.plan(lib, x.device, B, H, W, net.dim, ch, net.math, False)
        graphs = plan.__dict__.setdefault('step_graphs', {})
        key = (int(s), bool(reblur), float(self.omega), self.dp_rank, self.dp_world,
               tuple(p.data_ptr() for p in net.parameters()), self.gammas.data_ptr(), self.betas.data_ptr())
        sg = graphs.get(key)
        if sg is not None:
            return sg
        sg = _StepGraph()
        sg.x = torch.empty_like(x)
        sg.t = torch.zeros((B,), device=x.device, dtype=torch.long)
        sg.prev = torch.empty_like(x) if reblur else None
        saved_prev = self.img_prev_upsample
        try:
            if reblur:
                self.img_prev_upsample = sg.prev
            graph = torch.cuda.CUDAGraph()
            l0 = lib.sinddm_launch_count()
            with torch.cuda.graph(graph):
                out = self.p_sample(sg.x, sg.t, s)
                sg.x.copy_(out)
                sg.t.sub_(1)
            sg.launches = int(lib.sinddm_launch_count() - l0)
            sg.graph = graph
        except Exception as e:  # capture refused (e.g. an unsupported driver): keep the eager loop
            warnings.warn(f'sinddm_b200: CUDA-graph capture of the sampling step failed ({e}); using eager steps')
            self.use_step_graph = False
            sg = None
        finally:
            self.img_prev_upsample = saved_prev
        if sg is not None:
            while len(graphs) >= 4:
                graphs.pop(next(iter(graphs)))
            graphs[key] = sg
        return sg

    def _run_steps(self, img, s, t_hi, t_lo, total):
        """img <- p_sample(img, i, s) for i = t_hi-1 ... t_lo (the loop of models.py:480-485 / :540-545)."""
        global graph_replayed_launches
        device = img.device
        b = img.shape[0]
        steps = list(reversed(range(t_lo, t_hi)))
        reblur = bool(self.reblurring) and int(s) > 0
        use_graph = (self.use_step_graph and img.is_cuda and len(steps) >= 4 and not self.save_interm
                     and not self.clip_guided_sampling and not self.roi_guided_sampling)
        bar = tqdm(steps, desc='sampling loop time step', total=total, disable=None)
        if not use_graph:
            for i in bar:
                img = self.p_sample(img, torch.full((b,), i, device=device, dtype=torch.long), s)
            return img
        # first step eagerly: builds the plan and (re)packs the weights the graph's kernels read
        it = iter(bar)
        i0 = next(it)
        img = self.p_sample(img, torch.full((b,), i0, device=device, dtype=torch.long), s)
        sg = self._step_graph_for(img, s, reblur)
        if sg is None:
            for i in it:
                img = self.p_sample(img, torch.full((b,), i, device=device, dtype=torch.long), s)
            return img
        sg.x.copy_(img)
        sg.t.fill_(i0 - 1)
        if reblur:
            sg.prev.copy_(self.img_prev_upsample)
        for _ in it:
            sg.graph.replay()
            graph_replayed_launches += sg.launches
        return sg.x.clone()

    @torch.no_grad()
    def p_sample_loop(self, shape, s):
        """models.py:462-487"""
        device = self.betas.device
        img = self._randn(shape, device)
        t_min = self.num_timesteps_ideal[s + 1] if (self.sample_limited_t and s < (self.n_scales - 1)) else 0
        return self._run_steps(img, s, self.num_timesteps, t_min, self.num_timesteps)

    @torch.no_grad()
    def sample(self, batch_size=16, scale_0_size=None, s=0):
        """models.py:489-499"""
        image_size = scale_0_size if scale_0_size is not None else self.image_sizes[0]
        return self.p_sample_loop((batch_size, self.channels, image_size[0], image_size[1]), s=s)

    @torch.no_grad()
    def p_sample_via_scale_loop(self, batch_size, img, s, custom_t=None):
        """models.py:501-547"""
        device = self.betas.device
        total_t = self.num_timesteps_ideal[min(s, self.n_scales - 1)] - 1 if custom_t is None else custom_t
        b = batch_size
        # the kernels index the schedule / gamma tables with t and t-1 unchecked; the reference's gather raises for
        # the same input (custom_t, --sample_t_list, --start_t_style, --start_t_harm)
        if not 0 <= int(total_t) < self.num_timesteps:
            raise IndexError(f'start timestep {total_t} is outside the schedule [0, {self.num_timesteps})')
        self.img_prev_upsample = img
        img = self.q_sample(x_start=img, t=torch.Tensor.expand(torch.tensor(total_t, device=device), batch_size),
                            noise=None)
        t_min = self.num_timesteps_ideal[s + 1] if (self.sample_limited_t and s < (self.n_scales - 1)) else 0
        return self._run_steps(img, s, total_t, t_min, total_t)

    @torch.no_grad()
    def sample_via_scale(self, batch_size, img, s, scale_mul=(1, 1), custom_sample=False, custom_img_size_idx=0,
                         custom_t=None, custom_image_size=None):
        """models.py:549-568"""
        if custom_sample:
            if custom_img_size_idx >= self.n_scales:
                size = self.image_sizes[self.n_scales - 1]
                factor = self.scale_factor ** (custom_img_size_idx + 1 - self.n_scales)
                size = (int(size[0] * factor), int(size[1] * factor))
            else:
                size = self.image_sizes[custom_img_size_idx]
        else:
            size = self.image_sizes[s]
        image_size = (int(size[0] * scale_mul[0]), int(size[1] * scale_mul[1]))
        if custom_image_size is not None:
            image_size = custom_image_size
        img = F.interpolate(img, size=image_size, mode='bilinear')
        return self.p_sample_via_scale_loop(batch_size, img, s, custom_t=custom_t)
