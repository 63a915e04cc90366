"""Two-or-more-rank check of the fused optimizer step (run under torchrun on one NVLink box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_dp_check.py

Every rank fills its gradient bucket with rank-dependent values; sinddm_fused_step (peer loads over NVLink +
Adam + EMA) must give
  (1) the parameters that (sum of the gathered buckets in rank order) / world + torch.optim.Adam give (<= 2e-5 of
      max|p|: same arithmetic, only fma contraction differs),
  (2) bit-identical replicas on every rank,
  (3) against NCCL all_reduce + torch.optim.Adam (whose ring / tree / in-switch summation order differs in the last
      bits of the gradient mean): agreement within a STATED bound.  Adam normalises the update to m/(sqrt(v)+eps), so
      where the ranks' gradients cancel to ~1e-8 a last-bit difference of the mean can flip the update's sign; an
      element can then differ by up to 2*lr per step.  Bound: >= 99.9 % of the elements within 1e-5 of max|p|, and no
      element off by more than 2 * lr * steps.
With SINDDM_FUSED_MULTIMEM=1 the gradient sum is one multimem.ld_reduce per 16 bytes (reduced inside the NVSwitch);
its summation order is the hardware's, so (1) is then checked with (3)'s bound, and (2) -- every rank must still get
the same bits -- is the property that decides whether that mode can be used at all.
Prints FUSED_DP_OK and the per-step device time of both paths."""
import copy
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sinddm_b200 import SinDDMNet  # noqa: E402
from sinddm_b200 import dist as spdist  # noqa: E402
from sinddm_b200.fused_optim import FusedStep  # noqa: E402


def main():
    rank, local_rank, world = spdist.init_process_group()
    assert world > 1, "run under torchrun with >= 2 ranks"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    torch.manual_seed(0)                       # same initial replicas
    net = SinDDMNet(dim=160, multiscale=True, device=dev).to(dev)
    ema_net = copy.deepcopy(net)
    ref_net = copy.deepcopy(net)
    nccl_net = copy.deepcopy(net)
    fused = FusedStep(net, ema_net)
    opt = torch.optim.Adam(ref_net.parameters(), lr=1e-3)
    opt_nccl = torch.optim.Adam(nccl_net.parameters(), lr=1e-3)
    nsteps = 6
    g = torch.Generator(device=dev).manual_seed(100 + rank)      # different gradients per rank
    for it in range(nsteps):
        bucket = fused.bucket()
        bucket.copy_(torch.randn(bucket.numel(), device=dev, generator=g) * 1e-2)
        # reference gradient mean with the kernel's association order (rank 0 + rank 1 + ...): NCCL's ring / tree /
        # in-switch orders differ in the last bits, which Adam's normalisation amplifies wherever the ranks'
        # gradients cancel -- that comparison is reported separately below
        gathered = [torch.empty_like(bucket) for _ in range(world)]
        dist.all_gather(gathered, bucket.clone())
        flat = gathered[0].clone()
        for r in range(1, world):
            flat += gathered[r]
        flat *= 1.0 / world
        nccl = bucket.clone()
        dist.all_reduce(nccl)
        nccl /= world
        nccl_grad_diff = max(locals().get("nccl_grad_diff", 0.0),
                             float((nccl - flat).abs().max() / (flat.abs().max() + 1e-30)))
        for p, v in zip(ref_net.parameters(), flat.split([q.numel() for q in ref_net.parameters()])):
            p.grad = v.view_as(p).clone()
        for p, v in zip(nccl_net.parameters(), nccl.split([q.numel() for q in nccl_net.parameters()])):
            p.grad = v.view_as(p).clone()
        fused.step(1e-3, 1 if it == 0 else 2, 0.995)
        opt.step()
        opt_nccl.step()
    torch.cuda.synchronize()
    worst = 0.0
    torch.set_grad_enabled(False)
    for a, b in zip(net.parameters(), ref_net.parameters()):
        worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-12)))
    multimem = bool(fused.mc_ptr)
    if os.environ.get("SINDDM_FUSED_MULTIMEM", "") == "1":
        assert multimem, f"rank {rank}: SINDDM_FUSED_MULTIMEM=1 but the symmetric allocation has no multicast mapping"
    if multimem:
        # in-switch (NVLink SHARP) summation order is the hardware's, not rank order: same stated bound as (3)
        d1 = torch.cat([(a - b).abs().reshape(-1) for a, b in zip(net.parameters(), ref_net.parameters())])
        pm = max(float(b.abs().max()) for b in ref_net.parameters())
        assert float(d1.max()) <= 2 * 1e-3 * nsteps and float((d1 > 1e-5 * pm).double().mean()) <= 1e-3, \
            f"rank {rank}: multimem fused step vs ordered sum + Adam: max |dp| {float(d1.max())}"
    else:
        assert worst <= 2e-5, f"rank {rank}: fused step differs from ordered sum + Adam by {worst}"
    # (3) against NCCL's summation order, within the Adam-amplified bound stated in the module docstring
    pmax = max(float(b.abs().max()) for b in nccl_net.parameters())
    diffs = torch.cat([(a - b).abs().reshape(-1) for a, b in zip(net.parameters(), nccl_net.parameters())])
    nccl_worst_abs = float(diffs.max())
    nccl_frac_off = float((diffs > 1e-5 * pmax).double().mean())
    assert nccl_worst_abs <= 2 * 1e-3 * nsteps, f"rank {rank}: fused vs NCCL + Adam: max |dp| {nccl_worst_abs}"
    assert nccl_frac_off <= 1e-3, f"rank {rank}: fused vs NCCL + Adam: {nccl_frac_off:.2e} of the elements differ"
    # replicas must be bit-identical across ranks (fixed summation order)
    mine = fused.flat_param.clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    assert torch.equal(mine, ref0), f"rank {rank}: parameter replica differs from rank 0"

    # long replica-identity run (SINDDM_DP_CHECK_REPLICA_STEPS=N): N more steps on fresh rank-dependent gradients, then
    # every rank must still hold rank 0's bits -- the property in-switch reduction has to keep over a whole training
    extra = int(os.environ.get("SINDDM_DP_CHECK_REPLICA_STEPS", "0"))
    for it in range(extra):
        bucket = fused.bucket()
        bucket.copy_(torch.randn(bucket.numel(), device=dev, generator=g) * 1e-2)
        fused.step(1e-3, 2, 0.995)
    if extra:
        torch.cuda.synchronize()
        mine = torch.cat([fused.flat_param.reshape(-1), torch.cat([p.reshape(-1) for p in ema_net.parameters()])])
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        assert torch.equal(mine, ref0), f"rank {rank}: replica differs from rank 0 after {extra} more steps"
        if rank == 0:
            print(f"REPLICAS_BIT_IDENTICAL after {nsteps + extra} steps (parameters and EMA, world={world}, "
                  f"multimem={int(multimem)})", flush=True)

    # timing: fused kernel vs NCCL all-reduce + torch Adam (device time, max over ranks)
    def timed(fn, n=30):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def nccl_path():
        dist.all_reduce(flat)
        flat.div_(world)
        opt.step()

    t_fused = timed(lambda: fused.step(1e-3, 2, 0.995))
    t_nccl = timed(nccl_path)
    if rank == 0:
        print(f"FUSED_DP_OK world={world} multimem={int(multimem)} max_rel_diff={worst:.2e} replicas_bit_identical=True "
              f"nccl_vs_ordered_sum_grad_rel_diff={nccl_grad_diff:.2e} fused_vs_nccl_adam_max_abs={nccl_worst_abs:.2e} "
              f"(bound {2 * 1e-3 * nsteps:.1e}) fused_vs_nccl_adam_frac_over_1e-5={nccl_frac_off:.2e} (bound 1e-3) "
              f"fused_step_ms={t_fused:.4f} "
              f"nccl_allreduce_plus_adam_ms={t_nccl:.4f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
