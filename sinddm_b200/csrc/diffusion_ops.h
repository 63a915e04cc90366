// Launchers for the fused diffusion element-wise kernels (diffusion_ops.cu).
#pragma once

#include "host_common.h"

namespace sinddm {

int qsample_mix_launch(const float* x_start, const float* x_orig, const float* noise, const long long* t,
                       const float* sqrt_ac, const float* sqrt_1mac, const float* gammas, float* out, int B,
                       long long per_sample, cudaStream_t stream);

size_t l1_loss_scratch_floats();
int l1_loss_launch(const float* noise, const float* pred, long long n, float* loss, float* dpred, float* scratch,
                   cudaStream_t stream);

struct DdpmStepArgs {
    const float* x_t;        // [B, C*H*W] current sample
    const float* eps;        // predicted noise (denoiser output)
    const float* x_tilde;    // upsampled previous-scale sample (img_prev_upsample), re-blur mode only
    const float* noise;      // fresh N(0,1) draw for this step
    const long long* t;      // [B] timesteps (int64, as the reference passes them)
    float* out;              // x_{t-1}
    int B;
    long long per_sample;
    int reblur_mode;         // s > 0 and reblurring
    int clip_denoised;
    float omega;
    // schedule tables, each [T]
    const float *sqrt_recip_ac, *sqrt_recipm1_ac, *post_coef1, *post_coef2, *post_logvar, *ac, *sqrt_ac, *sqrt_1mac;
    const float* gammas;     // gammas[s-1] row, [T] (unclamped; the kernel clamps to [0, 0.55])
};
int ddpm_step_launch(const DdpmStepArgs& a, cudaStream_t stream);

// elements [first, first + count) of torch.randn's Philox stream (see include/sinddm_b200.h)
int philox_normal_rows_launch(float* out, long long first, long long count, long long stride, unsigned long long seed,
                              unsigned long long offset, cudaStream_t stream);

}  // namespace sinddm
