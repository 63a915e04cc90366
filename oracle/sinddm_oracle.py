"""CPU oracle for the SinDDM hot path -- TEST INFRASTRUCTURE ONLY.

A plain, functional restatement (stock torch fp32/fp64 tensor ops, no nn.Module, none of this repo's kernels; it
follows its inputs' device, so bench.py can also time it as "the reference's eager PyTorch path on the GPU") of the reference
algorithm on the path BASELINE.json names: the denoiser forward (SinDDMNet), the training loss
(MultiScaleGaussianDiffusion.p_losses) with gradients by autograd of this restatement, the diffusion schedule
and the reverse-sampling update (p_sample).  Every function cites the reference file:line it follows
(paths relative to the reference repository root).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / the timed CPU baseline -- never as part of the product path.  The product
(sinddm_b200/) has no CPU fallback and never imports from oracle/.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference itself, generated in the build container by tools/make_golden.py (which
imports /root/reference with a small stub shim) and committed under tests/golden/;
tests/test_oracle_golden.py replays them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

BLOCKS = ("l1", "l2", "l3", "l4")


# ---------------------------------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------------------------------

def param_shapes(dim: int = 160, channels: int = 3, time_dim: int = 32) -> "List[Tuple[str, Tuple[int, ...]]]":
    """Names and shapes of SinDDMNet(dim, channels, multiscale=True).state_dict(), in registration order.

    SinDDM/models.py:104-132 (network) and :54-67 (block): 52 tensors, 1,106,772 values at dim=160.
    """
    half = dim // 2
    out: List[Tuple[str, Tuple[int, ...]]] = [
        ("time_mlp.0.weight", (time_dim * 4, time_dim * 2)),
        ("time_mlp.0.bias", (time_dim * 4,)),
        ("time_mlp.2.weight", (time_dim, time_dim * 4)),
        ("time_mlp.2.bias", (time_dim,)),
    ]
    chans = [(channels, half), (half, dim), (dim, dim), (dim, half)]
    for name, (ci, co) in zip(BLOCKS, chans):
        out += [
            (f"{name}.mlp.1.weight", (time_dim, time_dim)),
            (f"{name}.mlp.1.bias", (time_dim,)),
            (f"{name}.time_reshape.weight", (ci, time_dim, 1, 1)),
            (f"{name}.time_reshape.bias", (ci,)),
            (f"{name}.ds_conv.weight", (ci, 1, 5, 5)),
            (f"{name}.ds_conv.bias", (ci,)),
            (f"{name}.net.0.weight", (co, ci, 3, 3)),
            (f"{name}.net.0.bias", (co,)),
            (f"{name}.net.2.weight", (co, co, 3, 3)),
            (f"{name}.net.2.bias", (co,)),
        ]
        if ci != co:
            out += [(f"{name}.res_conv.weight", (co, ci, 1, 1)), (f"{name}.res_conv.bias", (co,))]
    out += [("final_conv.0.weight", (channels, half, 1, 1)), ("final_conv.0.bias", (channels,))]
    return out


def synthetic_params(seed: int, dim: int = 160, channels: int = 3, dtype=torch.float32) -> Params:
    """Deterministic stand-in weights that do not depend on torch's init RNG.

    numpy's legacy RandomState stream is frozen across versions, so tools/make_golden.py (reference side)
    and the tests (oracle / CUDA side) regenerate identical tensors from the seed alone.  Scaling follows
    nn.Conv2d / nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)).
    """
    rs = np.random.RandomState(seed)
    params: Params = {}
    for name, shape in param_shapes(dim, channels):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else None
        if fan_in is None:  # bias: use the fan-in of the matching weight
            fan_in = int(np.prod(params[name.replace("bias", "weight")].shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        arr = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        params[name] = torch.from_numpy(arr).to(dtype)
    return params


# ---------------------------------------------------------------------------------------------------
# denoiser
# ---------------------------------------------------------------------------------------------------

def sinusoidal_pos_emb(x: torch.Tensor, dim: int = 32) -> torch.Tensor:
    """SinusoidalPosEmb.forward, SinDDM/models.py:39-46."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=x.device) * -emb)
    emb = x[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def conv_block(p: Params, name: str, x: torch.Tensor, cond_vec: torch.Tensor) -> torch.Tensor:
    """SinDDMConvBlock.forward, SinDDM/models.py:69-80."""
    c = x.shape[1]
    h = F.conv2d(x, p[f"{name}.ds_conv.weight"], p[f"{name}.ds_conv.bias"], padding=2, groups=c)          # :70
    cond = F.linear(F.gelu(cond_vec), p[f"{name}.mlp.1.weight"], p[f"{name}.mlp.1.bias"])                # :74
    cond = F.conv2d(cond[:, :, None, None], p[f"{name}.time_reshape.weight"], p[f"{name}.time_reshape.bias"])  # :75-76
    h = h + cond                                                                                          # :77
    h = F.conv2d(h, p[f"{name}.net.0.weight"], p[f"{name}.net.0.bias"], padding=1)                        # :63
    h = F.gelu(h)                                                                                         # :64
    h = F.conv2d(h, p[f"{name}.net.2.weight"], p[f"{name}.net.2.bias"], padding=1)                        # :65
    if f"{name}.res_conv.weight" in p:                                                                    # :67
        res = F.conv2d(x, p[f"{name}.res_conv.weight"], p[f"{name}.res_conv.bias"])
    else:
        res = x
    return h + res                                                                                        # :80


def net_forward(p: Params, x: torch.Tensor, time: torch.Tensor, scale) -> torch.Tensor:
    """SinDDMNet.forward (multiscale=True), SinDDM/models.py:134-151."""
    dtype = x.dtype
    scale_tensor = torch.ones(time.shape, dtype=dtype, device=x.device) * float(scale)                    # :137
    t = sinusoidal_pos_emb(time.to(dtype) if dtype == torch.float64 else time, 32).to(dtype)              # :138
    s = sinusoidal_pos_emb(scale_tensor, 32).to(dtype)                                                    # :139
    ts = torch.cat((t, s), dim=1)                                                                         # :140
    h = F.linear(ts, p["time_mlp.0.weight"], p["time_mlp.0.bias"])                                        # :106-110
    cond_vec = F.linear(F.gelu(h), p["time_mlp.2.weight"], p["time_mlp.2.bias"])
    for name in BLOCKS:                                                                                   # :146-149
        x = conv_block(p, name, x, cond_vec)
    return F.conv2d(x, p["final_conv.0.weight"], p["final_conv.0.bias"])                                  # :151


# ---------------------------------------------------------------------------------------------------
# schedule
# ---------------------------------------------------------------------------------------------------

def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """SinDDM/functions.py:117-127 (note linspace(0, steps, steps), quirk Q4)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


class Schedule:
    """Buffers of MultiScaleGaussianDiffusion.__init__, SinDDM/models.py:227-287."""

    def __init__(self, n_scales: int, scale_losses: Sequence[float], timesteps: int = 100, loss_factor: float = 1,
                 train_full_t: bool = False):
        betas = cosine_beta_schedule(timesteps)                                                           # :230
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.betas = f32(betas)
        self.alphas_cumprod = f32(ac)
        self.alphas_cumprod_prev = f32(ac_prev)
        self.sqrt_alphas_cumprod = f32(np.sqrt(ac))
        self.sqrt_one_minus_alphas_cumprod = f32(np.sqrt(1.0 - ac))
        self.log_one_minus_alphas_cumprod = f32(np.log(1.0 - ac))
        self.sqrt_recip_alphas_cumprod = f32(np.sqrt(1.0 / ac))
        self.sqrt_recipm1_alphas_cumprod = f32(np.sqrt(1.0 / ac - 1))
        pv = betas * (1.0 - ac_prev) / (1.0 - ac)                                                        # :259
        self.posterior_variance = f32(pv)
        self.posterior_log_variance_clipped = f32(np.log(np.maximum(pv, 1e-20)))
        self.posterior_mean_coef1 = f32(betas * np.sqrt(ac_prev) / (1.0 - ac))
        self.posterior_mean_coef2 = f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac))
        sigma_t = np.sqrt(1.0 - ac) / np.sqrt(ac)                                                        # :269
        self.num_timesteps_trained = [self.num_timesteps]
        self.num_timesteps_ideal = [self.num_timesteps]
        for i in range(n_scales - 1):                                                                     # :272-280
            self.num_timesteps_ideal.append(int(np.argmax(sigma_t > loss_factor * scale_losses[i])))
            self.num_timesteps_trained.append(int(timesteps) if train_full_t else self.num_timesteps_ideal[i + 1])
        gammas = torch.zeros((n_scales - 1, self.num_timesteps))                                          # :283-285
        for i in range(n_scales - 1):
            gammas[i, :] = (torch.tensor(sigma_t) / (loss_factor * scale_losses[i])).clamp(min=0, max=1)
        self.gammas = gammas

    def to(self, device) -> "Schedule":
        """Moves the tables (bench.py's reference-on-GPU eager baseline runs this restatement on `cuda`)."""
        for n in list(self.buffers()):
            setattr(self, n, getattr(self, n).to(device))
        return self

    def buffers(self) -> Dict[str, torch.Tensor]:
        names = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                 "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                 "posterior_mean_coef1", "posterior_mean_coef2", "gammas"]
        return {n: getattr(self, n) for n in names}


def extract(a: torch.Tensor, t: torch.Tensor, x_shape) -> torch.Tensor:
    """SinDDM/functions.py:105-108."""
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


# ---------------------------------------------------------------------------------------------------
# training loss
# ---------------------------------------------------------------------------------------------------

def q_sample(sch: Schedule, x_start, t, noise):
    """SinDDM/models.py:570-576."""
    return (extract(sch.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
            extract(sch.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)


def noisy_input(sch: Schedule, x_start, t, s: int, noise, x_orig=None):
    """The network input p_losses builds, SinDDM/models.py:582-590 (gamma unclamped, quirk Q3)."""
    if int(s) > 0:
        g = sch.gammas[s - 1].reshape(-1)
        x_mix = extract(g, t, x_start.shape) * x_start + (1 - extract(g, t, x_start.shape)) * x_orig
        return q_sample(sch, x_mix, t, noise)
    return q_sample(sch, x_start, t, noise)


def p_losses(p: Params, sch: Schedule, x_start, t, s: int, noise, x_orig=None, loss_type: str = "l1"):
    """SinDDM/models.py:578-611 ('l1' and 'l2'; 'l1_pred_img' is unused by main.py:97)."""
    x_noisy = noisy_input(sch, x_start, t, s, noise, x_orig)
    x_recon = net_forward(p, x_noisy, t, s)                                                               # :587/:591
    if loss_type == "l1":
        return (noise - x_recon).abs().mean()                                                             # :594
    if loss_type == "l2":
        return F.mse_loss(noise, x_recon)                                                                 # :596
    raise NotImplementedError(loss_type)


def loss_and_grads(p: Params, sch: Schedule, x_start, t, s: int, noise, x_orig=None):
    """loss.backward() of trainer.py:200-202 on the restated forward."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss = p_losses(leaf, sch, x_start, t, s, noise, x_orig)
    loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in leaf.items()}


# ---------------------------------------------------------------------------------------------------
# reverse sampling
# ---------------------------------------------------------------------------------------------------

def p_sample_update(sch: Schedule, x, eps, t, s: int, noise, x_tilde=None, reblurring: bool = True,
                    omega: float = 0.0, clip_denoised: bool = True):
    """x_{t-1} from x_t and the predicted noise: predict_start_from_noise (:306-318), the re-blur mix and clamp
    of p_mean_variance (:434-442), q_posterior (:321-352) and p_sample's noise add (:455-459).
    CLIP / ROI branches (:367-431) are off on this path."""
    shape = x.shape
    x_ddpm = extract(sch.sqrt_recip_alphas_cumprod, t, shape) * x - extract(sch.sqrt_recipm1_alphas_cumprod, t, shape) * eps
    reblur = reblurring and int(s) > 0
    if not reblur:
        x_recon, x_t_mix = x_ddpm, x_ddpm
    else:
        g = sch.gammas[s - 1].reshape(-1).clamp(0, 0.55)
        x_recon = (x_ddpm - extract(g, t, shape) * x_tilde) / (1 - extract(g, t, shape))
        x_t_mix = x_ddpm
    if reblur and t[0] > 0:
        g = sch.gammas[s - 1].reshape(-1).clamp(0, 0.55)
        x_tm1_mix = extract(g, t - 1, shape) * x_tilde + (1 - extract(g, t - 1, shape)) * x_recon
    else:
        x_tm1_mix = x_recon
    if clip_denoised:
        x_tm1_mix = x_tm1_mix.clamp(-1.0, 1.0)
        x_t_mix = x_t_mix.clamp(-1.0, 1.0)
    if not reblur:
        mean = extract(sch.posterior_mean_coef1, t, shape) * x_tm1_mix + extract(sch.posterior_mean_coef2, t, shape) * x
        logvar = extract(sch.posterior_log_variance_clipped, t, shape)
    elif t[0] > 0:
        var_hi = 1 - extract(sch.alphas_cumprod, t - 1, shape)
        var = (1 - omega) * torch.zeros(shape, dtype=x.dtype, device=x.device) + omega * var_hi
        logvar = torch.log(var.clamp(1e-20, None))
        mean = (extract(sch.sqrt_alphas_cumprod, t - 1, shape) * x_tm1_mix +
                torch.sqrt(1 - extract(sch.alphas_cumprod, t - 1, shape) - var) *
                (x - extract(sch.sqrt_alphas_cumprod, t, shape) * x_t_mix) /
                extract(sch.sqrt_one_minus_alphas_cumprod, t, shape))
    else:
        mean = x_tm1_mix
        logvar = extract(sch.posterior_log_variance_clipped, t, shape)
    mask = (1 - (t == 0).to(x.dtype)).reshape(shape[0], *((1,) * (len(shape) - 1)))
    return mean + mask * (0.5 * logvar).exp() * noise


def p_sample(p: Params, sch: Schedule, x, t, s: int, noise, x_tilde=None, **kw):
    """MultiScaleGaussianDiffusion.p_sample with the noise draw injected, SinDDM/models.py:449-459."""
    eps = net_forward(p, x, t, s)                                                                         # :356
    return p_sample_update(sch, x, eps, t, s, noise, x_tilde, **kw)


def p_sample_loop(p: Params, sch: Schedule, shape, s: int, generator: torch.Generator):
    """SinDDM/models.py:462-487 with torch.randn drawn from `generator` in the reference's call order (Q7)."""
    img = torch.randn(shape, generator=generator)
    for i in reversed(range(0, sch.num_timesteps)):
        t = torch.full((shape[0],), i, dtype=torch.long)
        eps = net_forward(p, img, t, s)
        noise = torch.randn(shape, generator=generator)
        img = p_sample_update(sch, img, eps, t, s, noise)
    return img


def p_sample_via_scale_loop(p: Params, sch: Schedule, img, s: int, total_t: int, generator: torch.Generator):
    """SinDDM/models.py:501-547: img is the bilinearly upsampled previous-scale sample (x_tilde)."""
    x_tilde = img
    b = img.shape[0]
    t0 = torch.full((b,), total_t, dtype=torch.long)
    img = q_sample(sch, img, t0, torch.randn(img.shape, generator=generator))                             # :518
    for i in reversed(range(0, total_t)):                                                                 # :540
        t = torch.full((b,), i, dtype=torch.long)
        eps = net_forward(p, img, t, s)
        noise = torch.randn(img.shape, generator=generator)
        img = p_sample_update(sch, img, eps, t, s, noise, x_tilde)
    return img


# ---------------------------------------------------------------------------------------------------
# trainer step
# ---------------------------------------------------------------------------------------------------

def ema_update(ema: Params, cur: Params, beta: float) -> None:
    """EMA.update_model_average, SinDDM/models.py:23-31: old * beta + (1 - beta) * new, parameters only (Q8)."""
    for k in ema:
        ema[k] = ema[k] * beta + (1 - beta) * cur[k].detach()


def train_steps(p: Params, sch: Schedule, data_list, draws, *, lr: float, milestones: Sequence[int] = (),
                gamma: float = 0.5, ema_decay: float = 0.995, step_start_ema: int = 2000, update_ema_every: int = 10,
                avg_window: int = 100):
    """MultiscaleTrainer.train, SinDDM/trainer.py:189-214, for len(draws) steps with gradient_accumulate_every=1 and
    every random draw injected: draws[i] = (s, t, noise).  data_list[s] = (orig batch, blurry batch).
    Returns (model params, ema params, running_loss, last lr).  Adam / MultiStepLR are torch's own, as in the
    reference (trainer.py:134-136)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    ema = {k: v.detach().clone() for k, v in p.items()}                                                   # :100,153
    opt = torch.optim.Adam(list(leaf.values()), lr=lr)                                                    # :134
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=list(milestones), gamma=gamma)          # :136
    running_loss, loss_avg = [], 0.0
    for step, (s, t, noise) in enumerate(draws):
        x_orig, x_blur = data_list[s]
        if int(s) > 0:                                                                                    # models.py:624-631
            loss = p_losses(leaf, sch, x_blur, t, s, noise, x_orig=x_orig)
        else:
            loss = p_losses(leaf, sch, x_orig, t, s, noise)
        loss_avg += loss.item()                                                                           # :202
        loss.backward()
        if step % avg_window == 0:                                                                        # :204-207 (Q5)
            running_loss.append(loss_avg / avg_window)
            loss_avg = 0.0
        opt.step()                                                                                        # :208
        opt.zero_grad()
        if step % update_ema_every == 0:                                                                  # :211, :155-159
            if step < step_start_ema:
                ema = {k: v.detach().clone() for k, v in leaf.items()}
            else:
                ema_update(ema, leaf, ema_decay)
        sched.step()                                                                                      # :212
    return ({k: v.detach() for k, v in leaf.items()}, ema, running_loss, sched.get_last_lr()[0])
