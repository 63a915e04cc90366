// Gradient all-reduce + Adam + EMA in ONE kernel over NVLink peer memory.
//
// Replaces, per optimizer step of MultiscaleTrainer.train (reference SinDDM/trainer.py:208-213):
//   [data parallel] all_reduce(grad bucket) / N            -- NCCL launch + copy-in / copy-out of the bucket
//   self.opt.step()                                        -- torch.optim.Adam: ~10 multi-tensor launches
//   self.opt.zero_grad()
//   if step % update_ema_every == 0: self.step_ema()       -- 52-tensor EMA loop or state_dict copy (models.py:18-31)
//
// Layout: every rank owns a flat fp32 gradient bucket (all 52 parameter gradients, written in place by
// sinddm_net_backward) inside a symmetric allocation that is peer-mapped into every other rank's address space
// (one process per GPU; the mapping is torch.distributed._symmetric_memory plumbing, the arithmetic and the
// synchronisation are this kernel).  At N = 1..8 ranks and 4.4 MB the collective is latency bound, so it is a
// ONE-SHOT all-reduce: after an in-kernel barrier over NVLink (epoch-stamped flags, st.release.sys /
// ld.acquire.sys) every rank loads all N buckets through NVLink P2P loads, adds them in rank order 0..N-1 (every
// rank computes the same bits, so the parameter replicas cannot drift) and applies Adam and the EMA update to its
// own replica in the same pass.  Buckets are double buffered by step parity: a bucket is rewritten two steps
// later, after the next step's barrier has proven that every peer finished reading it -- no second barrier.
#include "common.cuh"
#include "fused_optim.h"

namespace sinddm {

namespace {

SINDDM_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SINDDM_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer buckets are written by other GPUs between launches: bypass L1 and never use the read-only path
SINDDM_DEVINL float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

SINDDM_DEVINL unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

SINDDM_DEVINL float4 multimem_ld_reduce_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}

struct Args {
    FusedStepDesc d;
};

SINDDM_DEVINL void adam_one(float g, float& p, float& m, float& v, const FusedStepDesc& d) {
    // torch.optim.Adam (no weight decay, no amsgrad), same operation order as torch/optim/adam.py
    m = fmaf(g - m, 1.f - d.beta1, m);                         // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(g * g, 1.f - d.beta2, v * d.beta2);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / d.bias2_sqrt + d.eps;       // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
    p = p - d.step_size * (m / denom);                         // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
}

__global__ void __launch_bounds__(256) fused_allreduce_adam_ema_kernel(const Args a) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const FusedStepDesc& d = a.d;
    const int world = d.world, rank = d.rank;
    if (world > 1) {
        // -------- barrier over NVLink: every rank's bucket of this step is complete ---------------------------
        // (this rank's gradients were written by earlier kernels on this stream: visible device-wide at launch;
        //  peers read them from this GPU's memory through its L2, the coherence point)
        if (blockIdx.x == 0 && threadIdx.x < world && (int)threadIdx.x != rank) {
            __threadfence_system();
            st_release_sys(d.flags[threadIdx.x] + rank, d.epoch);      // flags[dst][src]
        }
        unsigned long long t0 = 0;
        if (d.wait_ns && blockIdx.x == 0 && threadIdx.x == 0) t0 = global_timer_ns();
        if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
            const uint32_t* f = d.flags[rank] + threadIdx.x;
            while ((int32_t)(ld_acquire_sys(f) - d.epoch) < 0) {
            }
        }
        __syncthreads();
        if (d.wait_ns && blockIdx.x == 0 && threadIdx.x == 0) {
            // how long this rank sat in the barrier = how far behind the slowest peer was (rank skew, not link time)
            const unsigned long long dt = global_timer_ns() - t0;
            d.wait_ns[0] += dt;
            if (dt > d.wait_ns[1]) d.wait_ns[1] = dt;
        }
    }
    const float inv_world = 1.f / (float)world;
    const long long n4 = d.n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.mc_grads && world > 1) {
            // NVLink SHARP: the switch adds the world buckets and returns the sum (one request instead of world loads)
            g = multimem_ld_reduce_f4(d.mc_grads + i * 4);
        } else {
#pragma unroll 1
            for (int r = 0; r < world; ++r) {          // fixed order: identical bits on every rank
                const float* src = d.grads[r] + i * 4;
                const float4 t = (r == rank) ? *reinterpret_cast<const float4*>(src) : ld_peer_f4(src);
                g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
            }
        }
        g.x *= inv_world; g.y *= inv_world; g.z *= inv_world; g.w *= inv_world;
        float4 p = reinterpret_cast<float4*>(d.param)[i];
        float4 m = reinterpret_cast<float4*>(d.exp_avg)[i];
        float4 v = reinterpret_cast<float4*>(d.exp_avg_sq)[i];
        adam_one(g.x, p.x, m.x, v.x, d);
        adam_one(g.y, p.y, m.y, v.y, d);
        adam_one(g.z, p.z, m.z, v.z, d);
        adam_one(g.w, p.w, m.w, v.w, d);
        reinterpret_cast<float4*>(d.param)[i] = p;
        reinterpret_cast<float4*>(d.exp_avg)[i] = m;
        reinterpret_cast<float4*>(d.exp_avg_sq)[i] = v;
        if (d.ema_mode == 1) {                     // hard copy before step_start_ema (trainer.py:156-158)
            reinterpret_cast<float4*>(d.ema)[i] = p;
        } else if (d.ema_mode == 2) {              // old * beta + (1 - beta) * new (models.py:28-31)
            float4 e = reinterpret_cast<float4*>(d.ema)[i];
            const float b = d.ema_beta, c = 1.f - d.ema_beta;
            e.x = e.x * b + c * p.x;
            e.y = e.y * b + c * p.y;
            e.z = e.z * b + c * p.z;
            e.w = e.w * b + c * p.w;
            reinterpret_cast<float4*>(d.ema)[i] = e;
        }
    }
}

}  // namespace

int fused_step_launch(const FusedStepDesc& d, cudaStream_t stream) {
    SINDDM_REQUIRE(d.world >= 1 && d.world <= kFusedMaxWorld, "fused_step: world=%d unsupported", d.world);
    SINDDM_REQUIRE(d.rank >= 0 && d.rank < d.world, "fused_step: bad rank");
    SINDDM_REQUIRE(d.n > 0 && d.n % 4 == 0, "fused_step: n must be a positive multiple of 4");
    SINDDM_REQUIRE(d.param && d.exp_avg && d.exp_avg_sq, "fused_step: NULL state");
    SINDDM_REQUIRE(d.ema_mode == 0 || d.ema, "fused_step: EMA requested without a buffer");
    for (int r = 0; r < d.world; ++r) {
        SINDDM_REQUIRE(d.grads[r] != nullptr, "fused_step: NULL gradient bucket for rank %d", r);
        SINDDM_REQUIRE(d.world == 1 || d.flags[r] != nullptr, "fused_step: NULL flag array for rank %d", r);
    }
    Args a;
    a.d = d;
    // every CTA waits on flags that only the peers' CTA 0 can set: the grid is kept within one wave so that
    // CTA 0 of every rank is always resident
    const int sms = device_info().initialized ? device_info().num_sms : 148;
    long long want = (d.n / 4 + 255) / 256;
    const int grid = (int)(want < sms ? want : sms);
    (void)launch_pdl(fused_allreduce_adam_ema_kernel, dim3(grid), dim3(256), (size_t)(0), stream, a);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
