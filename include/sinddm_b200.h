/*
 * sinddm_b200 -- C ABI of the B200-native SinDDM hot path (libsinddm_b200.so).
 *
 * The reference (fallenshock/SinDDM) is pure PyTorch: its "FFI" for this path is the set of ATen/cuDNN
 * calls behind nn.Conv2d / nn.Linear / nn.GELU and the tensor arithmetic of MultiScaleGaussianDiffusion.
 * Each entry point below names the reference code it replaces (file:line under the reference repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors); the library never
 *     allocates or frees user-visible memory and keeps no global mutable state beyond sinddm_init();
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all work is
 *     enqueued asynchronously on it, no hidden synchronisation, CUDA-graph capturable;
 *   - return value: 0 on success, negative sinddm_status otherwise; sinddm_last_error() gives the text
 *     (thread-local);
 *   - 3-channel image tensors are NCHW fp32 like the reference's; internal activations are NHWC fp32;
 *   - `math`: SINDDM_MATH_TF32 runs the dense 3x3/1x1 convolutions on the tcgen05 tensor cores with TF32
 *     operands and fp32 accumulation (the numerics class of the reference's own GPU default,
 *     torch.backends.cudnn.allow_tf32 = True); SINDDM_MATH_FP32 runs them on CUDA cores in plain fp32;
 *     SINDDM_MATH_TF32X3 (plans only) runs them on the same tensor-core kernels with every operand split into
 *     two TF32 values (x = hi + lo): each contraction accumulates x_hi*w_hi + x_hi*w_lo + x_lo*w_hi in fp32, i.e.
 *     fp32-class results (the reference with allow_tf32 = False) at three times the tensor-core work.
 */
#ifndef SINDDM_B200_H_
#define SINDDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SINDDM_ABI_VERSION 3

typedef enum sinddm_status {
    SINDDM_STATUS_OK = 0,
    SINDDM_STATUS_INVALID = -1,
    SINDDM_STATUS_CUDA = -2,
    SINDDM_STATUS_NOT_INIT = -3,
    SINDDM_STATUS_WORKSPACE = -4
} sinddm_status;

enum { SINDDM_MATH_FP32 = 0, SINDDM_MATH_TF32 = 1, SINDDM_MATH_TF32X3 = 2 };

/* Number of parameter tensors of SinDDMNet(multiscale=True), in state_dict / .parameters() order
 * (SinDDM/models.py:104-132, :54-67): time_mlp.{0,2}.{weight,bias}; for l1..l4: mlp.1.{w,b},
 * time_reshape.{w,b}, ds_conv.{w,b}, net.0.{w,b}, net.2.{w,b}, res_conv.{w,b} (absent for l3);
 * final_conv.0.{w,b}. */
#define SINDDM_NUM_PARAMS 52

/* ---- library ---------------------------------------------------------------------------------- */

/* Binds the library to `device` (must be sm_100), resolves the TMA descriptor encoder.  Idempotent. */
int sinddm_init(int device);
const char* sinddm_last_error(void);
int sinddm_abi_version(void);

/* Number of kernels this library has launched so far in the process (bench.py's gpu_launches). */
unsigned long long sinddm_launch_count(void);
/* Per-kernel CUDA-event timing on the launching stream, off by default.  kind 0 = tcgen05 conv (forward /
 * data gradient), 1 = tcgen05 weight gradient.  collect() synchronises on the recorded events and returns the
 * summed duration, algorithmic FLOPs (2 x real pixels x N x K) and launch count since enable(1). */
void sinddm_profile_enable(int on);
int sinddm_profile_collect(int kind, double* total_ms, double* total_flops, int* launches);

/* ---- whole denoiser: SinDDMNet.forward and its autograd backward (SinDDM/models.py:134-151) ---- */

typedef struct sinddm_plan sinddm_plan;

/* Bytes of workspace a plan for x[B,channels,H,W] with SinDDMNet(dim) needs.  training != 0 keeps every
 * activation backward needs (what autograd would save) and the gradient scratch. */
size_t sinddm_plan_workspace_bytes(int B, int H, int W, int dim, int channels, int math, int training);

/* Creates a plan over a caller-owned, 1024-byte aligned workspace (TMA descriptors point into it, so the
 * workspace must outlive the plan and must not move). */
int sinddm_plan_create(sinddm_plan** out, int B, int H, int W, int dim, int channels, int math, int training,
                       void* workspace, size_t workspace_bytes);
void sinddm_plan_destroy(sinddm_plan* plan);

/* Re-layouts the conv weights into GEMM operand order (and TF32-rounds them in TF32 mode).  Call after
 * every parameter update and before sinddm_net_forward.  params: SINDDM_NUM_PARAMS device pointers. */
int sinddm_net_pack_weights(sinddm_plan* plan, const float* const* params, void* stream);

/* out[B,3,H,W] = SinDDMNet(x[B,3,H,W], time[B] (int64), scale).  `freqs` = the 16 sinusoidal frequencies
 * exp(-k ln(1e4)/15) as computed by SinusoidalPosEmb.forward (models.py:41-43).
 * Replaces models.py:134-151 (called from :356 when sampling and :587/:591 when training). */
int sinddm_net_forward(sinddm_plan* plan, const float* const* params, const float* x, const int64_t* time,
                       float scale, const float* freqs, float* out, void* stream);

/* Gradients of all SINDDM_NUM_PARAMS parameters given dout = dLoss/dout[B,3,H,W]; must follow a
 * sinddm_net_forward on the same (training) plan.  grads: SINDDM_NUM_PARAMS device pointers, each shaped
 * like its parameter; they are overwritten, not accumulated.  Replaces what `loss.backward()`
 * (SinDDM/trainer.py:202, functions.py:97-102) runs through cuDNN dgrad / wgrad for the denoiser. */
int sinddm_net_backward(sinddm_plan* plan, const float* const* params, const float* dout, float* const* grads,
                        void* stream);

/* ---- single operators (the same kernels the plan drives; used by the parity tests) -------------- */

/* Dense conv as a GEMM over NHWC activations: out[p][n] = sum_{tap,c} in[p(+)tap][c] * w[tap][n][c]
 * (+ optional 1x1 residual conv from a second input, + epilogue).  Replaces nn.Conv2d 3x3/1x1 forward and
 * its data gradient inside SinDDMConvBlock (models.py:62-67,79-80). */
typedef struct sinddm_conv_desc {
    int B, H, W;
    const float* in;        /* [B,H,W,Cin] */
    int Cin;
    const float* w;         /* packed [ntaps][N][Cin] (see sinddm_pack_conv_weights) */
    int ntaps;              /* 9 (3x3, zero pad 1) or 1 (1x1) */
    const float* in_res;    /* [B,H,W,Cres] input of a fused 1x1 residual conv, or NULL */
    int Cres;
    const float* w_res;     /* [N][Cres] */
    int N;
    const float* bias;      /* [N] or NULL */
    const float* res_add;   /* [B,H,W,N] identity residual, or NULL */
    const float* x3;        /* [B,H,W,3] + w_res3 [N][3]: 1x1 residual conv from a 3-channel input, or NULL */
    const float* w_res3;
    int gelu;               /* exact-erf GELU on the result */
    float* out_pre;         /* [B,H,W,N] pre-activation copy, or NULL */
    const float* dgelu_z;   /* [B,H,W,N]: multiply by gelu'(z) (data-gradient epilogue), or NULL */
    const float* w_final;   /* [3][N] + b_final [3] + out_final NCHW [B,3,H,W]: fused trailing 1x1 conv, or NULL */
    const float* b_final;
    float* out_final;
    int round_tf32;         /* round `out` to TF32 */
    float* out;             /* [B,H,W,N] or NULL */
} sinddm_conv_desc;
int sinddm_conv_forward(const sinddm_conv_desc* desc, int math, void* stream);

/* Host-only query (no GPU needed): the tensor-core kernel is compiled once per set of epilogue features ("flavour": a
 * bit mask -- 1 streamed operand, 2 second streamed operand, 4 identity residual, 8 gelu' multiply, 16 3-channel
 * residual, 32 GELU, 64 pre-activation copy, 128 final conv, 256 3-channel copy, 512 column sums, 1024 tf32 rounding)
 * and a launch runs the smallest instantiation covering the descriptor's epilogue.  Returns that instantiation's mask
 * (2047 = the generic one), or a negative status.  Same numerics whichever runs; this exists so tests and profiles
 * can name the kernel (`tc_conv_kernel<0, mask>` in ncu). */
int sinddm_conv_epilogue_flavour(const sinddm_conv_desc* desc);

/* PyTorch OIHW weight [Cout][Cin][ntaps] -> forward operand dst_fwd[tap][Cout][Cin] and/or data-gradient
 * operand dst_dgrad[tap][Cin][Cout] (taps flipped). Either destination may be NULL.
 * round_tf32: 0 = as is, 1 = rounded to TF32, 2 = the 3xTF32 split: the contraction axis is tripled as
 * [lo | hi | hi] (dst_fwd[tap][Cout][3*Cin], dst_dgrad[tap][Cin][3*Cout]) to meet activations split by
 * sinddm_split3(mode 0). */
int sinddm_pack_conv_weights(const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad,
                             int round_tf32, void* stream);

/* 3xTF32 operand split of x [P][C] (C % 4 == 0): hi = tf32(x) (round to nearest), lo = tf32(x - hi).
 * mode 0: out [P][3*C] = [hi | lo | hi]   (operand of sinddm_conv_forward with Cin = 3*C and split weights);
 * mode 1: out [3][P][C] = hi, lo, hi      (x operand of sinddm_conv_wgrad with B tripled);
 * mode 2: out [3][P][C] = lo, hi, hi      (dy operand of the same call).
 * The two cross terms come first along the contraction axis and hi*hi last: the tensor core truncates its fp32
 * accumulator at every MMA step, in proportion to what the accumulator holds at that step.
 * With SINDDM_MATH_TF32 those calls then return x_hi*w_lo + x_lo*w_hi + x_hi*w_hi accumulated in fp32: what a
 * SINDDM_MATH_TF32X3 plan does for every dense convolution of SinDDMNet (models.py:62-67,79-80 with
 * torch.backends.cudnn.allow_tf32 = False). */
int sinddm_split3(const float* x, long long P, int C, float* out, int mode, void* stream);

/* dW[Cy][Cx][ntaps] = sum_p x[p(+)tap][ci] * dy[p][co] (PyTorch weight-gradient layout).
 * Replaces cuDNN backward-filter for the same modules.  workspace: sinddm_conv_wgrad_workspace_bytes(). */
size_t sinddm_conv_wgrad_workspace_bytes(int B, int H, int W, int Cx, int Cy, int ntaps, int math);
int sinddm_conv_wgrad(const float* x, int Cx, const float* dy, int Cy, int B, int H, int W, int ntaps, float* dw,
                      void* workspace, size_t workspace_bytes, int math, void* stream);

/* out = add + dw5x5(in; w[C][25], zero pad 2) + bias[c] + cond[b][c]; NHWC.  flip != 0 uses the flipped
 * kernel (data gradient).  Replaces ds_conv + `h + condition` (models.py:61,70,77). bias/cond/add may be NULL. */
int sinddm_dw5x5(const float* in, const float* w, const float* bias, const float* cond, const float* add, float* out,
                 int B, int H, int W, int C, int flip, int round_tf32, void* stream);
/* Depthwise weight [C][25], bias [C] and conditioning [B][C] gradients. */
size_t sinddm_dw5x5_wgrad_workspace_bytes(int B, int H, int C);
int sinddm_dw5x5_wgrad(const float* x, const float* dh, float* dw, float* db, float* dcond, void* workspace,
                       size_t workspace_bytes, int B, int H, int W, int C, void* stream);

/* out[c] = sum_p a[p][c] (bias gradients). */
size_t sinddm_colsum_workspace_bytes(int C);
int sinddm_colsum(const float* a, long long P, int C, float* out, void* workspace, size_t workspace_bytes,
                  void* stream);

int sinddm_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream);
int sinddm_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream);

/* ---- diffusion arithmetic (MultiScaleGaussianDiffusion) ---------------------------------------- */

/* x_noisy = sqrt_ac[t]*x_mix + sqrt_1mac[t]*noise, x_mix = gammas[t]*x_start + (1-gammas[t])*x_orig when
 * gammas != NULL (scale > 0), else x_start.  Replaces p_losses' mix + q_sample (models.py:583-586,570-576). */
int sinddm_qsample_mix(const float* x_start, const float* x_orig, const float* noise, const int64_t* t,
                       const float* sqrt_ac, const float* sqrt_1mac, const float* gammas, float* out, int B,
                       long long per_sample, void* stream);

/* loss[0] = mean |noise - pred|; dpred (optional) = d loss / d pred.  Replaces models.py:594 and its
 * autograd.  workspace: sinddm_l1_loss_workspace_bytes(). */
size_t sinddm_l1_loss_workspace_bytes(void);
int sinddm_l1_loss(const float* noise, const float* pred, long long n, float* loss, float* dpred, void* workspace,
                   size_t workspace_bytes, void* stream);

/* One reverse-diffusion update x_t -> x_{t-1} given the denoiser output: predict_start_from_noise,
 * re-blur mixing with x_tilde, clamp, q_posterior and the noise add (models.py:306-318,434-447,321-352,
 * 453-459).  reblur_mode = (s > 0 and reblurring). Tables are the module's registered buffers. */
typedef struct sinddm_ddpm_step_desc {
    const float* x_t;
    const float* eps;
    const float* x_tilde;
    const float* noise;
    const int64_t* t;
    float* out;
    int B;
    long long per_sample;
    int reblur_mode;
    int clip_denoised;
    float omega;
    const float* sqrt_recip_alphas_cumprod;
    const float* sqrt_recipm1_alphas_cumprod;
    const float* posterior_mean_coef1;
    const float* posterior_mean_coef2;
    const float* posterior_log_variance_clipped;
    const float* alphas_cumprod;
    const float* sqrt_alphas_cumprod;
    const float* sqrt_one_minus_alphas_cumprod;
    const float* gammas;
} sinddm_ddpm_step_desc;
int sinddm_ddpm_step(const sinddm_ddpm_step_desc* desc, void* stream);

/* Rows of a data-parallel noise draw.  The reference draws `noise = torch.randn_like(x)` for its whole batch
 * (SinDDM/models.py:580, 455, 470); under data parallelism rank r needs rows [r*b, (r+1)*b) of the draw ONE GPU would
 * have made.  This fills `out[0 .. count)` with elements [first, first + count) of the tensor that
 * `torch.randn(numel_global floats)` produces on the CUDA generator state (seed, offset): torch's grid-stride kernel
 * gives element li to thread li % stride (Philox subsequence) in round (li / stride) / 4, component (li / stride) % 4 of
 * curand_normal4 (ATen/native/cuda/DistributionTemplates.h); stride = 256 * grid of torch's launch for numel_global
 * (the caller computes it from the device properties and advances the generator offset as torch would). */
int sinddm_philox_normal_rows(float* out, long long first, long long count, long long stride,
                              unsigned long long seed, unsigned long long offset, void* stream);

/* ---- optimizer step: gradient all-reduce + Adam + EMA in one kernel ------------------------------ */

/* One optimizer step of MultiscaleTrainer.train (SinDDM/trainer.py:208-213: opt.step(), opt.zero_grad(), step_ema()
 * -> models.py:18-31) for the flat fp32 parameter vector, preceded -- when world > 1 -- by the data-parallel
 * gradient mean the reference does not have.  grads[r] is rank r's gradient bucket of this step and flags[r] rank
 * r's array of `world` uint32 flags (zero-initialised once), both inside a symmetric allocation that is
 * peer-mapped into this process (NVLink P2P); the kernel synchronises the ranks through the flags
 * (flags[dst][src] = epoch), sums the buckets in rank order, divides by world and applies Adam
 * (torch.optim.Adam semantics, no weight decay / amsgrad) and the EMA update in the same pass.  A bucket may be
 * rewritten only after the NEXT step's call has been enqueued (double buffer by step parity).  world == 1:
 * grads[0] is a plain device buffer and flags are ignored. */
#define SINDDM_FUSED_MAX_WORLD 8
typedef struct sinddm_fused_step_desc {
    int world, rank;
    long long n;                 /* elements, multiple of 4 */
    const float* grads[SINDDM_FUSED_MAX_WORLD];
    uint32_t* flags[SINDDM_FUSED_MAX_WORLD];
    uint32_t epoch;              /* strictly increasing per call, starts at 1 */
    float* param;
    float* exp_avg;
    float* exp_avg_sq;
    float* ema;                  /* may be NULL when ema_mode == 0 */
    float lr, beta1, beta2, eps;
    long long step;              /* 1-based Adam step count t (bias corrections 1 - beta^t) */
    int ema_mode;                /* 0: none, 1: ema = param, 2: ema = ema * ema_beta + (1 - ema_beta) * param */
    float ema_beta;
    unsigned long long* wait_ns; /* optional (NULL = off): device array of 2; [0] += nanoseconds this rank's first CTA
                                  * spent in the NVLink barrier waiting for its peers (= rank skew), [1] = max of them */
    const float* mc_grads;       /* optional (NULL = peer loads): NVLink-SHARP multicast address of this step's bucket
                                  * (all ranks' buckets bound to one multicast object): the gradient sum is then ONE
                                  * multimem.ld_reduce per 16 bytes, reduced inside the NVSwitch, instead of world loads */
} sinddm_fused_step_desc;
int sinddm_fused_step(const sinddm_fused_step_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SINDDM_B200_H_ */
