// Element-wise pieces of MultiScaleGaussianDiffusion fused into single kernels (3-channel NCHW tensors).
//
//   qsample_mix  : p_losses' blur mix + q_sample            (reference SinDDM/models.py:583-586,570-576)
//   l1_loss      : (noise - pred).abs().mean() + gradient    (models.py:594)
//   ddpm_step    : predict_start_from_noise, re-blur mixing, clamp, q_posterior, noise add
//                  (models.py:306-318, 434-447, 321-352, 453-459) -- ~40 eager launches and two host
//                  syncs per sampling step in the reference, one launch and no sync here.
//
// The arithmetic keeps the reference's operation order with separately rounded multiplies and adds
// (__fmul_rn / __fadd_rn block FMA contraction) so results track the eager fp32 path to the last bits.
#include <curand_kernel.h>

#include "common.cuh"
#include "diffusion_ops.h"

namespace sinddm {

namespace {

SINDDM_DEVINL float mul(float a, float b) { return __fmul_rn(a, b); }
SINDDM_DEVINL float add(float a, float b) { return __fadd_rn(a, b); }
SINDDM_DEVINL float sub(float a, float b) { return __fsub_rn(a, b); }
SINDDM_DEVINL float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

__global__ void qsample_mix_kernel(const float* __restrict__ x_start, const float* __restrict__ x_orig,
                                   const float* __restrict__ noise, const long long* __restrict__ t,
                                   const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac,
                                   const float* __restrict__ gammas, float* __restrict__ out, int B,
                                   long long per_sample) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const long long total = (long long)B * per_sample;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_sample);
        const long long tb = t[b];
        float x = x_start[i];
        if (gammas) {
            const float g = gammas[tb];
            // x_mix = g * x_blur + (1 - g) * x_orig        (gamma is NOT clamped in training, Q3)
            x = add(mul(g, x), mul(sub(1.f, g), x_orig[i]));
        }
        out[i] = add(mul(sqrt_ac[tb], x), mul(sqrt_1mac[tb], noise[i]));
    }
}

constexpr int kLossBlocks = 296;

__global__ void __launch_bounds__(256)
l1_partial_kernel(const float* __restrict__ noise, const float* __restrict__ pred, long long n, float inv_n,
                  float* __restrict__ partial, float* __restrict__ dpred) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float wsum[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float d = noise[i] - pred[i];
        s += fabsf(d);
        // d|noise - pred| / dpred = -sgn(noise - pred), sgn(0) = 0 like torch.abs's backward
        if (dpred) dpred[i] = d > 0.f ? -inv_n : (d < 0.f ? inv_n : 0.f);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? wsum[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

__global__ void l1_final_kernel(const float* __restrict__ partial, int nblk, float inv_n, float* __restrict__ loss) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float wsum[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) *loss = v * inv_n;
    }
}

__global__ void ddpm_step_kernel(const DdpmStepArgs a) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const long long total = (long long)a.B * a.per_sample;
    const bool reblur = a.reblur_mode != 0;
    const bool t0_pos = a.t[0] > 0;  // the reference branches on t[0] (models.py:331,434)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / a.per_sample);
        const long long tb = a.t[b];
        const float xt = a.x_t[i];
        const float eps = a.eps[i];
        // predict_start_from_noise (models.py:308-309)
        const float x_ddpm = sub(mul(a.sqrt_recip_ac[tb], xt), mul(a.sqrt_recipm1_ac[tb], eps));
        float x_tm1_mix, x_t_mix = x_ddpm;
        float xtil = 0.f;
        if (reblur) {
            xtil = a.x_tilde[i];
            const float g = fminf(fmaxf(a.gammas[tb], 0.f), 0.55f);              // clamp(0, 0.55) (:314)
            x_tm1_mix = sub(x_ddpm, mul(g, xtil)) / sub(1.f, g);                  // (:315-316)
            if (t0_pos) {                                                        // (:434-436)
                const float gp = fminf(fmaxf(a.gammas[tb - 1], 0.f), 0.55f);
                x_tm1_mix = add(mul(gp, xtil), mul(sub(1.f, gp), x_tm1_mix));
            }
        } else {
            x_tm1_mix = x_ddpm;
        }
        if (a.clip_denoised) {                                                   // (:440-442)
            x_tm1_mix = clamp1(x_tm1_mix);
            x_t_mix = clamp1(x_t_mix);
        }
        // q_posterior (models.py:321-352)
        float mean, logvar;
        if (!reblur) {
            mean = add(mul(a.post_coef1[tb], x_tm1_mix), mul(a.post_coef2[tb], xt));
            logvar = a.post_logvar[tb];
        } else if (t0_pos) {
            const float ac_prev = a.ac[tb - 1];
            const float var_hi = sub(1.f, ac_prev);
            const float var = add(mul(sub(1.f, a.omega), 0.f), mul(a.omega, var_hi));
            logvar = logf(fmaxf(var, 1e-20f));
            const float num = mul(sqrtf(sub(sub(1.f, ac_prev), var)), sub(xt, mul(a.sqrt_ac[tb], x_t_mix)));
            mean = add(mul(a.sqrt_ac[tb - 1], x_tm1_mix), num / a.sqrt_1mac[tb]);
        } else {
            mean = x_tm1_mix;
            logvar = a.post_logvar[tb];
        }
        // p_sample (models.py:456-459)
        const float mask = tb == 0 ? 0.f : 1.f;
        a.out[i] = add(mean, mul(mul(mask, expf(mul(0.5f, logvar))), a.noise[i]));
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148ll * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int qsample_mix_launch(const float* x_start, const float* x_orig, const float* noise, const long long* t,
                       const float* sqrt_ac, const float* sqrt_1mac, const float* gammas, float* out, int B,
                       long long per_sample, cudaStream_t stream) {
    SINDDM_REQUIRE(gammas == nullptr || x_orig != nullptr, "qsample_mix: gammas given without x_orig");
    const long long total = (long long)B * per_sample;
    (void)launch_pdl(qsample_mix_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, x_start, x_orig, noise, t, sqrt_ac, sqrt_1mac, gammas,
                                                                 out, B, per_sample);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

size_t l1_loss_scratch_floats() { return kLossBlocks; }

int l1_loss_launch(const float* noise, const float* pred, long long n, float* loss, float* dpred, float* scratch,
                   cudaStream_t stream) {
    SINDDM_REQUIRE(n > 0, "l1_loss: empty input");
    const float inv_n = 1.0f / (float)n;
    int nblk = grid_for(n, 256);
    if (nblk > kLossBlocks) nblk = kLossBlocks;
    (void)launch_pdl(l1_partial_kernel, dim3(nblk), dim3(256), (size_t)(0), stream, noise, pred, n, inv_n, scratch, dpred);
    SINDDM_CUDA_OK(cudaGetLastError());
    (void)launch_pdl(l1_final_kernel, dim3(1), dim3(256), (size_t)(0), stream, scratch, nblk, inv_n, loss);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int ddpm_step_launch(const DdpmStepArgs& a, cudaStream_t stream) {
    SINDDM_REQUIRE(!a.reblur_mode || (a.x_tilde && a.gammas), "ddpm_step: re-blur mode needs x_tilde and gammas");
    const long long total = (long long)a.B * a.per_sample;
    (void)launch_pdl(ddpm_step_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, a);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

// ---------------------------------------------------------------------------------------------------
// rows of torch.randn's stream (data-parallel noise shards)
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
philox_normal_rows_kernel(float* __restrict__ out, long long first, long long count, long long stride,
                          unsigned long long seed, unsigned long long offset) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        const long long li = first + e;
        const long long q = li / stride;
        const long long idx = li - q * stride;       // torch's thread index = Philox subsequence
        const long long round = q >> 2;               // that thread's curand_normal4 call number
        const int comp = (int)(q & 3);
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)idx, offset + 4ull * (unsigned long long)round, &st);
        const float4 r = curand_normal4(&st);
        out[e] = comp == 0 ? r.x : comp == 1 ? r.y : comp == 2 ? r.z : r.w;
    }
}
}  // namespace

int philox_normal_rows_launch(float* out, long long first, long long count, long long stride, unsigned long long seed,
                              unsigned long long offset, cudaStream_t stream) {
    long long want = (count + 255) / 256;
    const long long cap = 148ll * 16;
    philox_normal_rows_kernel<<<(unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), 256, 0, stream>>>(out, first, count,
                                                                                                        stride, seed, offset);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
