"""MultiscaleTrainer (reference: SinDDM/trainer.py:35-285) over the sm_100a diffusion module.

Same constructor keywords, attributes (`model`, `ema_model`, `data_list`, `opt`, `scheduler`, `step`,
`running_loss`), checkpoint format (`step, model, ema, sched, running_loss, running_scale`) and RNG call
order per step (multinomial -> randint -> randn, quirk Q7).  New here: data-parallel training under torchrun
(sinddm_b200.dist) and a loss read-back only every `avg_window` steps instead of a host sync per step.

`image2image` (harmonization / style transfer) runs on the same sampler; clip_sampling /
clip_roi_sampling / roi_guided_sampling are out of scope and raise
NotImplementedError.
"""
from __future__ import annotations

import copy
import datetime
from functools import partial
from pathlib import Path

import torch
from PIL import Image
from torch.optim import Adam
from torch.optim.lr_scheduler import MultiStepLR
from torch.utils import data

from . import dist as spdist
from .diffusion import EMA
from .functions import loss_backwards, num_to_groups


def _to_model_range(img: Image.Image) -> torch.Tensor:
    """transforms.ToTensor() followed by t*2-1 (trainer.py:46-50), without the torchvision dependency."""
    import numpy as np
    arr = np.asarray(img.convert('RGB'), dtype=np.uint8)
    ten = torch.from_numpy(arr.copy()).permute(2, 0, 1).to(torch.float32).div(255)
    return (ten * 2) - 1


class Dataset(data.Dataset):
    """trainer.py:35-63: always yields the first image of the scale folder (and its blurry `_recon` twin)."""

    def __init__(self, folder, image_size, blurry_img=False, exts=['jpg', 'jpeg', 'png']):
        super().__init__()
        self.folder = folder
        self.image_size = image_size
        self.blurry_img = blurry_img
        self.paths = [p for ext in exts for p in Path(f'{folder}').glob(f'**/*.{ext}')]
        if blurry_img:
            self.folder_recon = folder + '_recon/'
            self.paths_recon = [p for ext in exts for p in Path(f'{self.folder_recon}').glob(f'**/*.{ext}')]
        self.transform = _to_model_range

    def __len__(self):
        return len(self.paths) * 128

    def __getitem__(self, index):
        img = self.transform(Image.open(self.paths[0]))
        if self.blurry_img:
            return img, self.transform(Image.open(self.paths_recon[0]))
        return img

    def batch(self, batch_size):
        """What one DataLoader batch of the reference contains: min(batch_size, len(self)) identical rows."""
        n = min(batch_size, len(self))
        item = self[0]
        if self.blurry_img:
            return tuple(t.unsqueeze(0).repeat(n, 1, 1, 1) for t in item)
        return item.unsqueeze(0).repeat(n, 1, 1, 1)


class MultiscaleTrainer(object):

    def __init__(self, ms_diffusion_model, folder, *, ema_decay=0.995, n_scales=None, scale_factor=1,
                 image_sizes=None, train_batch_size=32, train_lr=2e-5, train_num_steps=100000,
                 gradient_accumulate_every=2, fp16=False, step_start_ema=2000, update_ema_every=10,
                 save_and_sample_every=25000, avg_window=100, sched_milestones=None, results_folder='./results',
                 device=None, scale_draw='device', pyramid=None, host_data=False, loss_readback='window'):
        super().__init__()
        self.device = device
        self.sched_milestones = [10000, 30000, 60000, 80000, 90000] if sched_milestones is None else sched_milestones
        image_sizes = [] if image_sizes is None else image_sizes
        self.model = ms_diffusion_model
        self.ema = EMA(ema_decay)
        self.ema_model = copy.deepcopy(self.model)
        self.update_ema_every = update_ema_every
        self.step_start_ema = step_start_ema
        self.save_and_sample_every = save_and_sample_every
        self.avg_window = avg_window

        # 'device': torch.multinomial on the device like the reference (trainer.py:197; same RNG stream).  The call
        # validates its input with host read-backs, i.e. it synchronises the stream it runs on: issued on the
        # training stream it drained the whole previous step before the next one could be enqueued (6 ms per step in
        # round 1).  It now runs on a side stream (_draw_scale): same generator, same Philox offsets, same values, but
        # the syncs only wait for the draw itself.  'host': the same uniform draw from a CPU generator (the device
        # stream of t / noise then differs from the reference's)
        self.scale_draw = scale_draw
        # host_data: keep data_list in pinned HOST memory and copy the step's batch in every step (the reference keeps
        # it on the device, trainer.py:120-132; bench.py's end-to-end leg uses this).  loss_readback: 'window' reads
        # the loss back once per avg_window; 'step' does the reference's loss.item() per micro-step (trainer.py:202)
        self.host_data = bool(host_data)
        if loss_readback not in ('window', 'step'):
            raise ValueError("loss_readback must be 'window' or 'step'")
        self.loss_readback = loss_readback
        self.last_loss = None
        self._copy_stream = None         # host_data: side stream + persistent device staging buffers per scale
        self._staging = {}
        self._staging_pending = None
        self.batch_size = train_batch_size           # GLOBAL batch (split over ranks under torchrun)
        self.n_scales = n_scales
        self.scale_factor = scale_factor
        self.gradient_accumulate_every = gradient_accumulate_every
        self.train_num_steps = train_num_steps

        # data parallel setup (no-op outside torchrun)
        self.rank, self.world = spdist.rank(), spdist.world_size()
        self.local_batch = spdist.shard_batch(train_batch_size, self.world)
        for m in (self.model, self.ema_model):
            if hasattr(m, 'set_data_parallel'):
                m.set_data_parallel(self.rank, self.world)

        self.input_paths = []
        self.ds_list = []
        self.data_list = []
        self.results_folder = Path(results_folder)
        if self.rank == 0:
            self.results_folder.mkdir(parents=True, exist_ok=True)

        # one (orig, blurry) batch per scale, resident on the device for the whole run (trainer.py:120-132)
        if pyramid is not None:
            # in-memory pyramid from create_img_scales(..., return_pyramid=True): the same tensors the PNG round trip
            # through scale_i/ and scale_i_recon/ gives, without touching the dataset folder
            if len(pyramid) != n_scales:
                raise ValueError(f'pyramid has {len(pyramid)} levels, n_scales is {n_scales}')
            n = min(self.local_batch, 128)                  # the reference's Dataset yields at most 128 rows
            for i, (level, blurry) in enumerate(pyramid):
                self.input_paths.append((folder or '') + 'scale_' + str(i))
                self.ds_list.append(None)
                orig = _to_model_range(level).unsqueeze(0).repeat(n, 1, 1, 1)
                other = _to_model_range(blurry).unsqueeze(0).repeat(n, 1, 1, 1) if i > 0 else orig.clone()
                self.data_list.append((self._place(orig), self._place(other)))
        for i in range(n_scales if pyramid is None else 0):
            self.input_paths.append(folder + 'scale_' + str(i))
            ds = Dataset(self.input_paths[i], image_sizes[i] if i < len(image_sizes) else None, blurry_img=i > 0)
            self.ds_list.append(ds)
            if i > 0:
                orig, blur = ds.batch(self.local_batch)
                self.data_list.append((self._place(orig), self._place(blur)))
            else:
                orig = ds.batch(self.local_batch)
                self.data_list.append((self._place(orig), self._place(orig.clone())))

        if self.host_data:
            self.prepare_host_data()
        self.opt = Adam(ms_diffusion_model.parameters(), lr=train_lr)
        self.scheduler = MultiStepLR(self.opt, milestones=self.sched_milestones, gamma=0.5)
        self.bucket = spdist.GradientBucket(ms_diffusion_model.parameters())

        self.step = 0
        self._fused = None          # set by _prepare_training()
        self.running_loss = []
        self.running_scale = []
        self.avg_t = []
        self.scale_counts = [0] * (n_scales or 0)      # new: how often train_step ran each scale

        assert not fp16, 'Apex must be installed in order for mixed precision training to be turned on'
        self.fp16 = fp16
        self.reset_parameters()

    # ---------------------------------------------------------------------------------------------
    def _place(self, batch):
        """Where a scale's training batch lives between steps: on the device (reference behaviour) or, with
        host_data=True, in pinned host memory from which every step copies it in."""
        if self.host_data:
            return batch.contiguous().pin_memory() if torch.cuda.is_available() else batch.contiguous()
        return batch.to(self.device)

    def prepare_host_data(self):
        """host_data=True: move every scale's batch to pinned host memory (if it is not there yet) and allocate the
        persistent device staging buffers of all scales NOW, so that no step of the training loop allocates (a
        cudaMalloc in the middle of a run stalls the host for tens of milliseconds)."""
        self.host_data = True
        self.data_list = [tuple(t if (not t.is_cuda and t.is_pinned()) else self._place(t.cpu()) for t in pair)
                          for pair in self.data_list]
        if torch.cuda.is_available() and self.data_list and self.data_list[0][0].is_pinned():
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            for s, pair in enumerate(self.data_list):
                if s not in self._staging:
                    self._staging[s] = {'buf': [tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in pair)
                                                for _ in range(2)], 'read': [None, None], 'i': 0}

    def _batch(self, s):
        pair = self.data_list[s]
        if not self.host_data:
            return pair
        if not pair[0].is_pinned():
            return tuple(t.to(self.device) for t in pair)
        # copy on a side stream into PERSISTENT device buffers (two per scale, alternating): the transfer overlaps
        # the previous step's kernels still queued on the training stream, and no allocator call sits on the step's
        # path (allocating the destination per step made the caching allocator call cudaMalloc on the copy stream's
        # pool now and then: sporadic 25-50 ms host stalls, tools/e2e_probe.py)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        st = self._staging.get(s)
        if st is None:
            st = {'buf': [tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in pair) for _ in range(2)],
                  'read': [None, None], 'i': 0}
            self._staging[s] = st
        i = st['i']
        st['i'] = i ^ 1
        main = torch.cuda.current_stream(self.device)
        if st['read'][i] is not None:
            self._copy_stream.wait_event(st['read'][i])      # the step that last read this buffer has consumed it
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(st['buf'][i], pair):
                dst.copy_(src, non_blocking=True)
        main.wait_stream(self._copy_stream)
        self._staging_pending = (st, i)                       # train_step records the "consumed" event after the forward
        return st['buf'][i]

    def reset_parameters(self):
        self.ema_model.load_state_dict(self.model.state_dict())
        self._ema_weights_changed()

    def _ema_weights_changed(self):
        """The EMA copy's parameters were rewritten (load_state_dict copies in place, EMA.update_model_average
        rebinds `.data`): neither is guaranteed to change (data_ptr, _version), the key the packed conv weights are
        cached under, so say it explicitly (stale packed weights in ema_model.sample)."""
        net = getattr(self.ema_model, 'denoise_fn', None)
        if net is not None and hasattr(net, 'mark_weights_updated'):
            net.mark_weights_updated()

    def step_ema(self):
        """trainer.py:155-159: hard copy until step_start_ema, EMA afterwards (quirk Q8)."""
        if self.step < self.step_start_ema:
            self.reset_parameters()
            return
        self.ema.update_model_average(self.ema_model, self.model)
        self._ema_weights_changed()

    def save(self, milestone):
        """trainer.py:161-177 (rank 0 only; the loss plot needs matplotlib and is skipped when it is absent)."""
        if self.rank != 0:
            return
        payload = {
            'step': self.step,
            'model': self.model.state_dict(),
            'ema': self.ema_model.state_dict(),
            'sched': self.scheduler.state_dict(),
            'running_loss': self.running_loss,
            'running_scale': self.running_scale,
        }
        torch.save(payload, str(self.results_folder / f'model-{milestone}.pt'))
        # new: the reference's checkpoint has no optimizer state (quirk Q12), so a resumed run
        # restarts Adam's moments; the state goes into a SEPARATE file so that model-N.pt keeps the reference's keys
        torch.save(self._optimizer_state(), str(self.results_folder / f'optim-{milestone}.pt'))
        try:
            from matplotlib import pyplot as plt
        except Exception:
            return
        plt.rcParams['figure.figsize'] = [16, 8]
        plt.plot(self.running_loss)
        plt.grid(True)
        plt.ylim((0, 0.2))
        plt.savefig(str(self.results_folder / 'running_loss'))
        plt.clf()

    def load(self, milestone):
        """trainer.py:179-187 (optimizer state is not part of the checkpoint, quirk Q12)."""
        ckpt = torch.load(str(self.results_folder / f'model-{milestone}.pt'), map_location=self.device)
        self.step = ckpt['step']
        self.model.load_state_dict(ckpt['model'])
        self.ema_model.load_state_dict(ckpt['ema'])
        self.scheduler.load_state_dict(ckpt['sched'])
        self.running_loss = ckpt['running_loss']
        for m in (self.model, self.ema_model):      # packed conv weights are rebuilt from the loaded parameters
            net = getattr(m, 'denoise_fn', None)
            if net is not None and hasattr(net, 'mark_weights_updated'):
                net.mark_weights_updated()
        opt_path = self.results_folder / f'optim-{milestone}.pt'
        if opt_path.exists():       # absent for checkpoints written by the reference: Adam restarts, like there
            self._set_optimizer_state(torch.load(str(opt_path), map_location='cpu'))

    def _optimizer_state(self):
        """Adam state in one layout for both optimizer implementations: step count + flat moments (parameter order)."""
        sizes = [p.numel() for p in self.model.parameters()]
        if self._fused is not None:
            n = sum(sizes)
            return {'step': int(self._fused.t), 'lr': float(self.opt.param_groups[0]['lr']),
                    'exp_avg': self._fused.exp_avg[:n].detach().cpu().clone(),
                    'exp_avg_sq': self._fused.exp_avg_sq[:n].detach().cpu().clone()}
        pending = getattr(self, '_pending_optim', None)
        if pending is not None and not self.opt.state:
            return pending
        step, m, v = 0, [], []
        for p, k in zip(self.model.parameters(), sizes):
            st = self.opt.state.get(p, {})
            step = max(step, int(st['step']) if 'step' in st else 0)
            m.append(st['exp_avg'].detach().reshape(-1).cpu() if 'exp_avg' in st else torch.zeros(k))
            v.append(st['exp_avg_sq'].detach().reshape(-1).cpu() if 'exp_avg_sq' in st else torch.zeros(k))
        return {'step': step, 'lr': float(self.opt.param_groups[0]['lr']), 'exp_avg': torch.cat(m),
                'exp_avg_sq': torch.cat(v)}

    def _set_optimizer_state(self, state):
        self._pending_optim = state
        if getattr(self, '_fused_decided', False):
            self._apply_optimizer_state()

    def _apply_optimizer_state(self):
        state = getattr(self, '_pending_optim', None)
        if state is None:
            return
        self._pending_optim = None
        params = list(self.model.parameters())
        sizes = [p.numel() for p in params]
        if state['exp_avg'].numel() != sum(sizes):
            raise ValueError('optimizer checkpoint does not match the model parameters')
        if 'lr' in state:
            # the scheduler state alone does not restore the optimizer's learning rate (the reference resumes at
            # train_lr until the next milestone); with the optimizer file the run continues where it stopped
            for group in self.opt.param_groups:
                group['lr'] = float(state['lr'])
        if self._fused is not None:
            n = sum(sizes)
            self._fused.t = int(state['step'])
            self._fused.exp_avg[:n].copy_(state['exp_avg'])
            self._fused.exp_avg_sq[:n].copy_(state['exp_avg_sq'])
            return
        if int(state['step']) == 0:
            return
        for p, m, v in zip(params, state['exp_avg'].split(sizes), state['exp_avg_sq'].split(sizes)):
            self.opt.state[p] = {'step': torch.tensor(float(state['step'])),
                                 'exp_avg': m.view_as(p).to(p.device).clone(),
                                 'exp_avg_sq': v.view_as(p).to(p.device).clone()}

    # ---------------------------------------------------------------------------------------------
    def _draw_scale(self):
        """trainer.py:197: s ~ multinomial(num_timesteps_trained) from the device generator.  torch.multinomial checks
        its weights with host read-backs (stream synchronisations), so it runs on a side stream: the values depend on
        the generator's Philox offset, which advances at call time on the host, not on the stream -- the draw is the
        one the reference makes, but the training stream is not drained and the host keeps running ahead."""
        if self.scale_draw == 'host':
            return int(torch.multinomial(input=self._s_weights_host, num_samples=1, generator=self._host_gen))
        if self._draw_stream is None:
            return int(torch.multinomial(input=self._s_weights, num_samples=1))
        with torch.cuda.stream(self._draw_stream):
            return int(torch.multinomial(input=self._s_weights, num_samples=1))

    def train_step(self, s=None):
        """One optimizer step (trainer.py:196-214).  Returns the (device, fp32) loss of the last micro-batch."""
        s = self._draw_scale() if s is None else int(s)
        if s < len(self.scale_counts):
            self.scale_counts[s] += 1
        loss = None
        fused = self._fused
        if fused is not None:
            # backward writes the 52 gradients straight into this step's (peer-mapped) bucket
            self.model.denoise_fn.set_grad_bucket(fused.bucket())
        for _ in range(self.gradient_accumulate_every):
            batch = self._batch(s)
            loss = self.model(batch, s)
            pending = getattr(self, '_staging_pending', None)
            if pending is not None:      # the batch is only read by the forward pass (q_sample / blur mix)
                st, i = pending
                if st['read'][i] is None:
                    st['read'][i] = torch.cuda.Event()
                st['read'][i].record(torch.cuda.current_stream(self.device))
                self._staging_pending = None
            if self.loss_readback == 'step':
                self.last_loss = loss.item()                # the reference's per-micro-step host sync (trainer.py:202)
                self._loss_acc_host += self.last_loss
            else:
                self._loss_acc += loss.detach().double()
            # (loss / 1 is the identity in fp32: skip the extra kernel and autograd node on the critical path)
            loss_backwards(self.fp16, loss if self.gradient_accumulate_every == 1 else
                           loss / self.gradient_accumulate_every, self.opt)
        if fused is None:
            self.bucket.all_reduce_mean()
        else:
            self.model.denoise_fn.set_grad_bucket(None)     # manual backward calls outside the trainer keep `.grad`
        if self.step % self.avg_window == 0:
            if self.loss_readback == 'step':
                acc = torch.tensor(self._loss_acc_host, dtype=torch.float64, device=self.device)
                self._loss_acc_host = 0.0
            else:
                acc = self._loss_acc.clone()
                self._loss_acc.zero_()
            if self.world > 1:
                import torch.distributed as tdist
                tdist.all_reduce(acc)
                acc /= self.world
            avg = float(acc.item()) / self.avg_window     # first report divides one loss by the window (Q5)
            if self.rank == 0:
                print(f'step:{self.step} loss:{avg}')
            self.running_loss.append(avg)
        if fused is not None:
            # one kernel: gradient mean over the ranks (NVLink peer loads), Adam, EMA (trainer.py:208-213)
            ema_mode = 0
            if self.step % self.update_ema_every == 0:
                ema_mode = 1 if self.step < self.step_start_ema else 2
            fused.step(self.opt.param_groups[0]['lr'], ema_mode, self.ema.beta)
        else:
            self.opt.step()
            self.opt.zero_grad()
            if self.step % self.update_ema_every == 0:
                self.step_ema()
        self.scheduler.step()
        self.step += 1
        return loss

    def _make_fused_step(self):
        """The fused all-reduce + Adam + EMA step (sinddm_b200.fused_optim) when the configuration allows it:
        CUDA, one micro-batch per optimizer step, at most FUSED_MAX_WORLD ranks, SINDDM_FUSED_STEP != 0.  Otherwise
        (and when its construction fails on ANY rank: no peer access, symmetric-memory rendezvous refused)
        torch.optim.Adam + the NCCL bucket.  The decision is agreed across ranks."""
        import os
        import warnings

        from . import _capi
        net = getattr(self.model, 'denoise_fn', None)
        ok = (os.environ.get('SINDDM_FUSED_STEP', '1') != '0' and self.gradient_accumulate_every == 1
              and net is not None and hasattr(net, 'set_grad_bucket')
              and next(net.parameters()).is_cuda and self.world <= _capi.FUSED_MAX_WORLD)
        fused, why = None, ''
        if ok:
            try:
                from .fused_optim import FusedStep
                group = self.opt.param_groups[0]
                fused = FusedStep(net, self.ema_model.denoise_fn, betas=group['betas'], eps=group['eps'])
            except Exception as e:      # noqa: BLE001 -- any failure means "use the fallback", on every rank
                fused, why = None, f'{type(e).__name__}: {e}'
        if self.world > 1:
            import torch.distributed as tdist
            flag = torch.tensor([1 if fused is not None else 0], dtype=torch.int32,
                                device=self.device if tdist.get_backend() == 'nccl' else 'cpu')
            tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)
            if int(flag.item()) == 0 and fused is not None:
                fused, why = None, 'another rank could not build the fused step'
        if ok and fused is None:
            warnings.warn(f'sinddm_b200: fused optimizer step unavailable ({why}); using NCCL all-reduce + torch.optim.Adam')
            # FusedStep may have re-pointed the parameters at its flat buffers before failing: values are intact
        if fused is not None:
            self.opt._opt_called = True      # the scheduler only reads the learning rate from this optimizer now
        return fused

    def _sync_replicas(self):
        """Data-parallel start: every rank must hold bit-identical parameters, EMA parameters and RNG state, because
        nothing re-synchronises the replicas later (the fused step applies the same update everywhere) and the scale
        draw / global (t, noise) stream come from each rank's own generator.  Rank 0's state wins."""
        if self.world <= 1:
            return
        import torch.distributed as tdist
        on_gpu = tdist.get_backend() == 'nccl'
        with torch.no_grad():
            for m in (self.model, self.ema_model):
                for p in m.parameters():
                    tdist.broadcast(p.data, src=0)
                for b in m.buffers():
                    tdist.broadcast(b.data, src=0)
        for m in (self.model, self.ema_model):
            net = getattr(m, 'denoise_fn', None)
            if net is not None and hasattr(net, 'mark_weights_updated'):
                net.mark_weights_updated()
        cpu_state = torch.get_rng_state()
        st = cpu_state.to(self.device) if on_gpu else cpu_state.clone()
        tdist.broadcast(st, src=0)
        torch.set_rng_state(st.cpu())
        if on_gpu:
            st = torch.cuda.get_rng_state(self.device).to(self.device)
            tdist.broadcast(st, src=0)
            torch.cuda.set_rng_state(st.cpu(), self.device)

    def _check_replicas(self):
        """Turns silent replica drift into an error: nothing re-synchronises the data-parallel replicas after
        _sync_replicas (every rank applies the same fused update), so at every checkpoint milestone the ranks compare
        a 64-bit checksum of the parameter and EMA bits.  Collective; two scalars per call."""
        if self.world <= 1:
            return
        import torch.distributed as tdist
        sums = []
        with torch.no_grad():
            for m in (self.model, self.ema_model):
                acc = torch.zeros((), dtype=torch.int64, device=self.device)
                for k, p in enumerate(m.parameters()):
                    bits = p.detach().contiguous().view(torch.int32).to(torch.int64)
                    acc = acc + (bits.sum() * (2 * k + 1))
                sums.append(acc)
        mine = torch.stack(sums)
        lo, hi = mine.clone(), mine.clone()
        tdist.all_reduce(lo, op=tdist.ReduceOp.MIN)
        tdist.all_reduce(hi, op=tdist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise RuntimeError(f'data-parallel replicas diverged at step {self.step} (rank {self.rank}: parameter/EMA '
                               f'checksums {mine.tolist()}, min {lo.tolist()}, max {hi.tolist()}); the ranks no longer '
                               f'train the same model')

    def _prepare_training(self):
        if not getattr(self, '_replicas_synced', False):
            self._sync_replicas()
            self._replicas_synced = True
        if self._fused is None and not getattr(self, '_fused_decided', False):
            self._fused = self._make_fused_step()
            self._fused_decided = True
            self._apply_optimizer_state()      # state loaded before the optimizer implementation was chosen
        self._s_weights = torch.tensor(self.model.num_timesteps_trained, device=self.device, dtype=torch.float)
        self._s_weights_host = self._s_weights.cpu()
        on_cuda = self._s_weights.is_cuda
        if on_cuda:
            torch.cuda.current_stream(self._s_weights.device).synchronize()     # weights visible to the side stream
        if not hasattr(self, '_draw_stream'):
            self._draw_stream = torch.cuda.Stream(device=self._s_weights.device) if on_cuda else None
        if not hasattr(self, '_host_gen'):
            self._host_gen = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 63))
        if not hasattr(self, '_loss_acc'):
            self._loss_acc = torch.zeros((), dtype=torch.float64, device=self.device)
            self._loss_acc_host = 0.0

    def _gather_images(self, local):
        """Rank-ordered concatenation of every rank's image shard (the global noise stream is sharded in rank order,
        so this is the batch one GPU would have produced).  Collective: every rank must call it; the result is
        complete on every rank."""
        if self.world <= 1:
            return local
        import torch.distributed as tdist
        parts = [torch.empty_like(local) for _ in range(self.world)]
        tdist.all_gather(parts, local.contiguous())
        return torch.cat(parts, dim=0)

    def train(self):
        self._prepare_training()
        while self.step < self.train_num_steps:
            self.train_step()
            if self.step % self.save_and_sample_every == 0:
                milestone = self.step // self.save_and_sample_every
                batches = num_to_groups(16, self.batch_size)
                shard = all(n % self.world == 0 for n in batches)     # else every rank samples the whole group
                if not shard:
                    self.ema_model.set_data_parallel(0, 1)
                try:
                    images = torch.cat([self.ema_model.sample(batch_size=n // self.world if shard else n)
                                        for n in batches], dim=0)
                finally:
                    self.ema_model.set_data_parallel(self.rank, self.world)
                if shard:
                    images = self._gather_images(images)
                images = (images + 1) * 0.5
                if self.rank == 0:
                    from torchvision import utils
                    utils.save_image(images, str(self.results_folder / f'sample-{milestone}.png'), nrow=4)
                self._check_replicas()
                self.save(milestone)
        if self.rank == 0:
            print('training completed')

    # ---------------------------------------------------------------------------------------------
    def sample_scales(self, scale_mul=None, batch_size=16, custom_sample=False, custom_image_size_idxs=None,
                      custom_scales=None, image_name='', start_noise=True, custom_t_list=None, desc=None,
                      save_unbatched=True, save_images=True):
        """trainer.py:226-285: coarse-to-fine sampling driver, always on the EMA model.  Under torchrun each
        rank generates batch_size / world_size images with no collective on the sampling path; when images are
        written, the shards are gathered (rank order = the 1-GPU batch order) and rank 0 writes all batch_size of
        them.  Returns the list of per-scale batches of THIS rank; `save_images=False` skips gather and PNG writes."""
        ema = self.ema_model
        if desc is None:
            desc = f'sample_{str(datetime.datetime.now()).replace(":", "_")}'
        if ema.reblurring:
            desc = desc + '_rblr'
        if ema.sample_limited_t:
            desc = desc + '_t_lmtd'
        if custom_t_list is None:
            custom_t_list = ema.num_timesteps_ideal[1:]
        if custom_scales is None:
            custom_scales = [*range(self.n_scales)]
            n_scales = self.n_scales
        else:
            n_scales = len(custom_scales)
        if custom_image_size_idxs is None:
            custom_image_size_idxs = [*range(self.n_scales)]
        local_b = spdist.shard_batch(batch_size, self.world)

        samples = []
        out_dir = Path(str(self.results_folder / 'final_samples'))
        if save_images and self.rank == 0:
            out_dir.mkdir(parents=True, exist_ok=True)
        if scale_mul is not None:
            base = self.model.image_sizes[custom_image_size_idxs[0]]
            scale_0_size = (int(base[0] * scale_mul[0]), int(base[1] * scale_mul[1]))
        else:
            scale_0_size = None
        t_list = [ema.num_timesteps_trained[0]] + custom_t_list
        prefix = '_'.join(str(e) for e in t_list)
        final_img = None
        for i in range(n_scales):
            if start_noise and i == 0:
                samples.append(ema.sample(batch_size=local_b, scale_0_size=scale_0_size, s=custom_scales[i]))
            elif i == 0:       # inject the training image instead of noise
                first = Image.open(self.input_paths[custom_scales[i]] + '/' + image_name)
                samples.append(_to_model_range(first).repeat(local_b, 1, 1, 1).to(self.device))
            else:
                samples.append(ema.sample_via_scale(local_b, samples[i - 1], s=custom_scales[i],
                                                    scale_mul=scale_mul, custom_sample=custom_sample,
                                                    custom_img_size_idx=custom_image_size_idxs[i],
                                                    custom_t=custom_t_list[int(custom_scales[i]) - 1]))
            final_img = (samples[i] + 1) * 0.5
            if save_images:
                final_img = self._gather_images(final_img)
            if save_images and self.rank == 0:
                from torchvision import utils
                utils.save_image(final_img, str(out_dir / prefix) +
                                 f'_out_s{i}_{desc}_sm_{scale_mul[0]}_{scale_mul[1]}.png', nrow=4)
        if save_images and save_unbatched and self.rank == 0:
            from torchvision import utils
            out_dir = Path(str(self.results_folder / f'final_samples_unbatched_{desc}'))
            out_dir.mkdir(parents=True, exist_ok=True)
            for b in range(final_img.shape[0]):
                utils.save_image(final_img[b], str(out_dir / prefix) + f'_out_b{b}.png')
        return samples

    # ---------------------------------------------------------------------------------------------
    def image2image(self, input_folder='', input_file='', mask='', hist_ref_path='', image_name='', start_s=1,
                    custom_t=None, batch_size=16, scale_mul=(1, 1), device=None, use_hist=False, save_unbatched=True,
                    auto_scale=None, mode=None, save_images=True):
        """trainer.py:287-361: harmonization / style transfer = inject the (histogram-matched) input image at scale
        `start_s`, noise it to custom_t[s] and run the reverse chain of the remaining scales on the EMA model; for
        harmonization the result is blended with the input through the dilated, blurred mask.  Returns the list of
        per-scale batches (the last one is the final composite in [0, 1])."""
        import os

        import numpy as np

        from .functions import dilate_mask, match_histograms
        device = self.device if device is None else device
        if custom_t is None:
            custom_t = self.ema_model.num_timesteps_ideal
        input_img = Image.open(os.path.join(input_folder, input_file)).convert("RGB")
        image_size = input_img.size
        if auto_scale is not None:
            scaler = np.sqrt((image_size[0] * image_size[1]) / auto_scale)
            if scaler > 1:
                image_size = (int(image_size[0] / scaler), int(image_size[1] / scaler))
                input_img = input_img.resize(image_size, Image.LANCZOS)
        if mode == 'harmonization':
            mask_img = Image.open(os.path.join(input_folder, mask)).convert("RGB").resize(image_size, Image.LANCZOS)
            mask_ten = torch.from_numpy(np.asarray(mask_img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div(255)
            mask_img = torch.from_numpy(dilate_mask(mask_ten, mode=mode)).to(device=device, dtype=torch.float32)
        else:
            mask_img = 1
        if use_hist:
            image_name = image_name.rsplit(".", 1)[0] + '.png'
            orig_sample_0 = Image.open(hist_ref_path + image_name).convert("RGB")
            input_img = Image.fromarray(match_histograms(image=np.array(input_img), reference=np.array(orig_sample_0),
                                                         channel_axis=2))
        input_img_tensor = _to_model_range(input_img)
        input_size = input_img_tensor.shape[1:]
        input_img_batch = input_img_tensor.repeat(batch_size, 1, 1, 1).to(device)

        out_dir = Path(str(self.results_folder / 'i2i_final_samples'))
        if save_images and self.rank == 0:
            out_dir.mkdir(parents=True, exist_ok=True)
        t_string = '_'.join(str(e) for e in custom_t)
        time = str(datetime.datetime.now()).replace(":", "_")
        if start_s > 0:  # the starting scale has no mixing between blurry and clean images (in place, like the reference)
            self.ema_model.gammas[start_s - 1].clamp_(0, 0)
        samples, final_img = [], None
        for i in range(self.n_scales - start_s):
            s = i + start_s
            ds_factor = self.scale_factor ** (self.n_scales - s - 1)
            cur_size = (int(input_size[0] / ds_factor), int(input_size[1] / ds_factor))
            src = input_img_batch if i == 0 else samples[i - 1]
            samples.append(self.ema_model.sample_via_scale(batch_size, src, s=s, custom_t=custom_t[s],
                                                           scale_mul=scale_mul, custom_image_size=cur_size))
            final_img = (samples[i] + 1) * 0.5
            if i == self.n_scales - start_s - 1:
                denorm = ((input_img_batch + 1) * 0.5).clamp_(0.0, 1.0)
                final_img = mask_img * final_img + (1 - mask_img) * denorm
                samples[i] = final_img
            if save_images and self.rank == 0:
                from torchvision import utils
                stem = input_file.rsplit(".", 1)[0]
                utils.save_image(final_img, str(out_dir / f'{stem}_i2i_s_{start_s + i}_t_{t_string}_hist_'
                                                          f'{"on" if use_hist else "off"}_{time}.png'), nrow=4)
        if save_images and save_unbatched and self.rank == 0:
            from torchvision import utils
            out_dir = Path(str(self.results_folder / f'unbatched_i2i_s{start_s}_t_{t_string}_{time}'))
            out_dir.mkdir(parents=True, exist_ok=True)
            for b in range(batch_size):
                utils.save_image(final_img[b], os.path.join(out_dir, input_file + f'_out_b{b}_i2i.png'))
        return samples

    def clip_sampling(self, *args, **kwargs):
        raise NotImplementedError('CLIP-guided sampling is outside the sinddm_b200 hot path')

    def clip_roi_sampling(self, *args, **kwargs):
        raise NotImplementedError('CLIP-guided sampling is outside the sinddm_b200 hot path')

    def roi_guided_sampling(self, *args, **kwargs):
        raise NotImplementedError('ROI-guided sampling is outside the sinddm_b200 hot path')
