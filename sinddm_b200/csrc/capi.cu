// extern "C" surface of libsinddm_b200.so -- thin argument checking over the internal launchers.
#include "../../include/sinddm_b200.h"

#include <math.h>
#include <string.h>

#include <new>

#include "diffusion_ops.h"
#include "fused_optim.h"
#include "net.h"

using namespace sinddm;

struct sinddm_plan {
    Plan p;
};

static_assert(SINDDM_NUM_PARAMS == kNumParams, "parameter count out of sync with net.h");
static_assert(SINDDM_FUSED_MAX_WORLD == kFusedMaxWorld, "fused step world limit out of sync");

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int sinddm_init(int device) { return init_device(device); }
const char* sinddm_last_error(void) { return last_error(); }
int sinddm_abi_version(void) { return SINDDM_ABI_VERSION; }
unsigned long long sinddm_launch_count(void) { return launch_count(); }
void sinddm_profile_enable(int on) { prof_enable(on); }
int sinddm_profile_collect(int kind, double* total_ms, double* total_flops, int* launches) {
    SINDDM_REQUIRE(total_ms && total_flops && launches, "profile_collect: NULL argument");
    return prof_collect(kind, total_ms, total_flops, launches);
}

size_t sinddm_plan_workspace_bytes(int B, int H, int W, int dim, int channels, int math, int training) {
    if (B < 1 || H < 1 || W < 1 || dim < 2 || channels < 1) return 0;
    return plan_workspace_bytes(B, H, W, dim, channels, math, training);
}

int sinddm_plan_create(sinddm_plan** out, int B, int H, int W, int dim, int channels, int math, int training,
                       void* workspace, size_t workspace_bytes) {
    SINDDM_REQUIRE(out != nullptr, "plan_create: out is NULL");
    *out = nullptr;
    SINDDM_REQUIRE(device_info().initialized, "sinddm_init() has not been called");
    sinddm_plan* h = new (std::nothrow) sinddm_plan;
    SINDDM_REQUIRE(h != nullptr, "plan_create: out of host memory");
    const int rc = plan_build(&h->p, B, H, W, dim, channels, math, training, workspace, workspace_bytes);
    if (rc != 0) {
        delete h;
        return rc;
    }
    *out = h;
    return SINDDM_OK;
}

void sinddm_plan_destroy(sinddm_plan* plan) { delete plan; }

int sinddm_net_pack_weights(sinddm_plan* plan, const float* const* params, void* stream) {
    SINDDM_REQUIRE(plan && params, "net_pack_weights: NULL argument");
    return net_pack_weights(&plan->p, params, as_stream(stream));
}

int sinddm_net_forward(sinddm_plan* plan, const float* const* params, const float* x, const int64_t* time,
                       float scale, const float* freqs, float* out, void* stream) {
    SINDDM_REQUIRE(plan && params && x && time && freqs && out, "net_forward: NULL argument");
    return net_forward(&plan->p, params, x, reinterpret_cast<const long long*>(time), scale, freqs, out,
                       as_stream(stream));
}

int sinddm_net_backward(sinddm_plan* plan, const float* const* params, const float* dout, float* const* grads,
                        void* stream) {
    SINDDM_REQUIRE(plan && params && dout && grads, "net_backward: NULL argument");
    return net_backward(&plan->p, params, dout, grads, as_stream(stream));
}

static void conv_problem_from_desc(const sinddm_conv_desc* d, ConvProblem* out) {
    ConvProblem& p = *out;
    memset(&p, 0, sizeof(p));
    p.B = d->B; p.H = d->H; p.W = d->W;
    p.in = d->in; p.Cin = d->Cin; p.w = d->w; p.ntaps = d->ntaps;
    p.in_res = d->in_res; p.Cres = d->Cres; p.w_res = d->w_res; p.N = d->N;
    p.w_blocked = 0;
    p.ep.bias = d->bias; p.ep.res_add = d->res_add; p.ep.x3 = d->x3; p.ep.w_res3 = d->w_res3;
    p.ep.gelu = d->gelu; p.ep.out_pre = d->out_pre; p.ep.dgelu_z = d->dgelu_z;
    p.ep.w_final = d->w_final; p.ep.b_final = d->b_final; p.ep.out_final = d->out_final;
    p.ep.round_tf32 = d->round_tf32; p.ep.out = d->out; p.ep.out3 = nullptr; p.ep.pre_grad = 0; p.ep.fast_math = 0;
}

int sinddm_conv_epilogue_flavour(const sinddm_conv_desc* d) {
    SINDDM_REQUIRE(d != nullptr, "conv_epilogue_flavour: NULL descriptor");
    ConvProblem p;
    conv_problem_from_desc(d, &p);
    return tc_conv_flavour_mask(p.ep);
}

int sinddm_conv_forward(const sinddm_conv_desc* d, int math, void* stream) {
    SINDDM_REQUIRE(d != nullptr, "conv_forward: NULL descriptor");
    SINDDM_REQUIRE(d->in && d->w, "conv_forward: in / w are required");
    SINDDM_REQUIRE(d->out || d->out_final, "conv_forward: no output requested");
    ConvProblem p;
    conv_problem_from_desc(d, &p);
    SINDDM_REQUIRE(p.B >= 1 && p.H >= 1 && p.W >= 1 && p.Cin >= 1 && p.N >= 1, "conv_forward: bad shape");
    SINDDM_REQUIRE(!p.ep.x3 || p.ep.w_res3, "conv_forward: x3 given without w_res3");
    SINDDM_REQUIRE(!p.ep.w_final || p.ep.out_final, "conv_forward: w_final given without out_final");
    if (math == MATH_TF32) {
        SINDDM_REQUIRE(tc_conv_supported(p), "conv_forward: shape Cin=%d Cres=%d N=%d ntaps=%d has no tensor-core path",
                       p.Cin, p.Cres, p.N, p.ntaps);
        TcConvOp op;
        SINDDM_TRY(tc_conv_prepare(p, &op));
        return tc_conv_launch(op, as_stream(stream));
    }
    SINDDM_REQUIRE(math == MATH_FP32, "conv_forward: math mode %d unknown here (SINDDM_MATH_TF32X3 is a plan mode: "
                   "single operators get it from sinddm_split3 + split-packed weights under SINDDM_MATH_TF32)", math);
    return simt_conv_launch(p, as_stream(stream));
}

int sinddm_pack_conv_weights(const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad,
                             int round_tf32, void* stream) {
    SINDDM_REQUIRE(w && (dst_fwd || dst_dgrad), "pack_conv_weights: NULL argument");
    SINDDM_REQUIRE(Cout >= 1 && Cin >= 1 && (ntaps == 9 || ntaps == 1), "pack_conv_weights: bad shape");
    SINDDM_REQUIRE(round_tf32 >= 0 && round_tf32 <= 2, "pack_conv_weights: round_tf32=%d unknown", round_tf32);
    return pack_conv_weights_launch(w, Cout, Cin, ntaps, dst_fwd, dst_dgrad, round_tf32, as_stream(stream));
}

int sinddm_split3(const float* x, long long P, int C, float* out, int mode, void* stream) {
    SINDDM_REQUIRE(x && out && P >= 1 && C >= 4, "split3: bad argument");
    return split3_launch(x, P, C, out, mode, as_stream(stream));
}

static int wgrad_nsplit_for(int B, int H, int W, int Cx, int Cy, int ntaps, int math) {
    if (math == MATH_TF32 && tc_wgrad_supported(Cx, Cy)) return tc_wgrad_nsplit(B, H, W, Cx, Cy, ntaps);
    return simt_wgrad_nsplit(B, H, W, Cx, Cy, ntaps);
}

size_t sinddm_conv_wgrad_workspace_bytes(int B, int H, int W, int Cx, int Cy, int ntaps, int math) {
    if (B < 1 || H < 1 || W < 1 || Cx < 1 || Cy < 1) return 0;
    return (size_t)wgrad_nsplit_for(B, H, W, Cx, Cy, ntaps, math) * ntaps * Cx * Cy * sizeof(float);
}

int sinddm_conv_wgrad(const float* x, int Cx, const float* dy, int Cy, int B, int H, int W, int ntaps, float* dw,
                      void* workspace, size_t workspace_bytes, int math, void* stream) {
    SINDDM_REQUIRE(x && dy && dw && workspace, "conv_wgrad: NULL argument");
    SINDDM_REQUIRE(ntaps == 9 || ntaps == 1, "conv_wgrad: ntaps must be 9 or 1");
    SINDDM_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Cx >= 1 && Cy >= 1, "conv_wgrad: bad shape");
    const size_t need = sinddm_conv_wgrad_workspace_bytes(B, H, W, Cx, Cy, ntaps, math);
    if (workspace_bytes < need) {
        set_error("conv_wgrad: workspace of %zu bytes < required %zu", workspace_bytes, need);
        return SINDDM_ERR_WORKSPACE;
    }
    WgradProblem p;
    p.B = B; p.H = H; p.W = W; p.x = x; p.Cx = Cx; p.dy = dy; p.Cy = Cy; p.ntaps = ntaps;
    p.partial = static_cast<float*>(workspace);
    p.nsplit = wgrad_nsplit_for(B, H, W, Cx, Cy, ntaps, math);
    if (math == MATH_TF32) {
        SINDDM_REQUIRE(tc_wgrad_supported(Cx, Cy), "conv_wgrad: Cx=%d Cy=%d has no tensor-core path", Cx, Cy);
        TcWgradOp op;
        SINDDM_TRY(tc_wgrad_prepare(p, &op));
        SINDDM_TRY(tc_wgrad_launch(op, as_stream(stream)));
    } else {
        SINDDM_REQUIRE(math == MATH_FP32, "conv_wgrad: unknown math mode %d", math);
        SINDDM_TRY(simt_wgrad_launch(p, as_stream(stream)));
    }
    return wgrad_reduce_launch(p.partial, p.nsplit, ntaps, Cx, Cy, dw, 0, as_stream(stream));
}

int sinddm_dw5x5(const float* in, const float* w, const float* bias, const float* cond, const float* add, float* out,
                 int B, int H, int W, int C, int flip, int round_tf32, void* stream) {
    SINDDM_REQUIRE(in && w && out, "dw5x5: NULL argument");
    SINDDM_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1, "dw5x5: bad shape");
    return dw5x5_launch(in, w, bias, cond, add, out, B, H, W, C, flip, round_tf32, as_stream(stream));
}

size_t sinddm_dw5x5_wgrad_workspace_bytes(int B, int H, int C) {
    if (B < 1 || H < 1 || C < 1) return 0;
    return dw5x5_wgrad_scratch_floats(B, H, C) * sizeof(float);
}

int sinddm_dw5x5_wgrad(const float* x, const float* dh, float* dw, float* db, float* dcond, void* workspace,
                       size_t workspace_bytes, int B, int H, int W, int C, void* stream) {
    SINDDM_REQUIRE(x && dh && workspace, "dw5x5_wgrad: NULL argument");
    SINDDM_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1, "dw5x5_wgrad: bad shape");
    if (workspace_bytes < sinddm_dw5x5_wgrad_workspace_bytes(B, H, C)) {
        set_error("dw5x5_wgrad: workspace too small");
        return SINDDM_ERR_WORKSPACE;
    }
    return dw5x5_wgrad_launch(x, dh, dw, db, dcond, static_cast<float*>(workspace), B, H, W, C, as_stream(stream));
}

size_t sinddm_colsum_workspace_bytes(int C) { return C < 1 ? 0 : colsum_scratch_floats(C) * sizeof(float); }

int sinddm_colsum(const float* a, long long P, int C, float* out, void* workspace, size_t workspace_bytes,
                  void* stream) {
    SINDDM_REQUIRE(a && out && workspace, "colsum: NULL argument");
    SINDDM_REQUIRE(P >= 1 && C >= 1, "colsum: bad shape");
    if (workspace_bytes < sinddm_colsum_workspace_bytes(C)) {
        set_error("colsum: workspace too small");
        return SINDDM_ERR_WORKSPACE;
    }
    return colsum_launch(a, P, C, out, static_cast<float*>(workspace), as_stream(stream));
}

int sinddm_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
    SINDDM_REQUIRE(src && dst && B >= 1 && C >= 1 && H >= 1 && W >= 1, "nchw_to_nhwc: bad argument");
    return nchw_to_nhwc_launch(src, dst, B, C, H, W, as_stream(stream));
}

int sinddm_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
    SINDDM_REQUIRE(src && dst && B >= 1 && C >= 1 && H >= 1 && W >= 1, "nhwc_to_nchw: bad argument");
    return nhwc_to_nchw_launch(src, dst, B, C, H, W, as_stream(stream));
}

int sinddm_philox_normal_rows(float* out, long long first, long long count, long long stride, unsigned long long seed,
                              unsigned long long offset, void* stream) {
    SINDDM_REQUIRE(out != nullptr, "philox_normal_rows: NULL output");
    SINDDM_REQUIRE(first >= 0 && count >= 1 && stride >= 256 && stride % 256 == 0, "philox_normal_rows: bad range");
    SINDDM_REQUIRE(offset % 4 == 0, "philox_normal_rows: the generator offset must be a multiple of 4");
    return philox_normal_rows_launch(out, first, count, stride, seed, offset, as_stream(stream));
}

int sinddm_qsample_mix(const float* x_start, const float* x_orig, const float* noise, const int64_t* t,
                       const float* sqrt_ac, const float* sqrt_1mac, const float* gammas, float* out, int B,
                       long long per_sample, void* stream) {
    SINDDM_REQUIRE(x_start && noise && t && sqrt_ac && sqrt_1mac && out, "qsample_mix: NULL argument");
    SINDDM_REQUIRE(B >= 1 && per_sample >= 1, "qsample_mix: bad shape");
    return qsample_mix_launch(x_start, x_orig, noise, reinterpret_cast<const long long*>(t), sqrt_ac, sqrt_1mac,
                              gammas, out, B, per_sample, as_stream(stream));
}

size_t sinddm_l1_loss_workspace_bytes(void) { return l1_loss_scratch_floats() * sizeof(float); }

int sinddm_l1_loss(const float* noise, const float* pred, long long n, float* loss, float* dpred, void* workspace,
                   size_t workspace_bytes, void* stream) {
    SINDDM_REQUIRE(noise && pred && loss && workspace, "l1_loss: NULL argument");
    if (workspace_bytes < sinddm_l1_loss_workspace_bytes()) {
        set_error("l1_loss: workspace too small");
        return SINDDM_ERR_WORKSPACE;
    }
    return l1_loss_launch(noise, pred, n, loss, dpred, static_cast<float*>(workspace), as_stream(stream));
}

int sinddm_ddpm_step(const sinddm_ddpm_step_desc* d, void* stream) {
    SINDDM_REQUIRE(d != nullptr, "ddpm_step: NULL descriptor");
    SINDDM_REQUIRE(d->x_t && d->eps && d->noise && d->t && d->out, "ddpm_step: NULL tensor");
    SINDDM_REQUIRE(d->sqrt_recip_alphas_cumprod && d->sqrt_recipm1_alphas_cumprod && d->posterior_mean_coef1 &&
                       d->posterior_mean_coef2 && d->posterior_log_variance_clipped && d->alphas_cumprod &&
                       d->sqrt_alphas_cumprod && d->sqrt_one_minus_alphas_cumprod,
                   "ddpm_step: NULL schedule table");
    SINDDM_REQUIRE(d->B >= 1 && d->per_sample >= 1, "ddpm_step: bad shape");
    DdpmStepArgs a;
    a.x_t = d->x_t; a.eps = d->eps; a.x_tilde = d->x_tilde; a.noise = d->noise;
    a.t = reinterpret_cast<const long long*>(d->t);
    a.out = d->out; a.B = d->B; a.per_sample = d->per_sample;
    a.reblur_mode = d->reblur_mode; a.clip_denoised = d->clip_denoised; a.omega = d->omega;
    a.sqrt_recip_ac = d->sqrt_recip_alphas_cumprod; a.sqrt_recipm1_ac = d->sqrt_recipm1_alphas_cumprod;
    a.post_coef1 = d->posterior_mean_coef1; a.post_coef2 = d->posterior_mean_coef2;
    a.post_logvar = d->posterior_log_variance_clipped; a.ac = d->alphas_cumprod;
    a.sqrt_ac = d->sqrt_alphas_cumprod; a.sqrt_1mac = d->sqrt_one_minus_alphas_cumprod;
    a.gammas = d->gammas;
    return ddpm_step_launch(a, as_stream(stream));
}

int sinddm_fused_step(const sinddm_fused_step_desc* d, void* stream) {
    SINDDM_REQUIRE(d != nullptr, "fused_step: NULL descriptor");
    SINDDM_REQUIRE(d->world >= 1 && d->world <= SINDDM_FUSED_MAX_WORLD, "fused_step: world=%d unsupported", d->world);
    SINDDM_REQUIRE(d->step >= 1, "fused_step: step must be >= 1");
    SINDDM_REQUIRE(d->beta1 >= 0.f && d->beta1 < 1.f && d->beta2 >= 0.f && d->beta2 < 1.f, "fused_step: bad betas");
    FusedStepDesc a;
    memset(&a, 0, sizeof(a));
    a.world = d->world; a.rank = d->rank; a.n = d->n;
    for (int r = 0; r < d->world; ++r) {
        a.grads[r] = d->grads[r];
        a.flags[r] = d->flags[r];
    }
    a.epoch = d->epoch;
    a.param = d->param; a.exp_avg = d->exp_avg; a.exp_avg_sq = d->exp_avg_sq; a.ema = d->ema;
    a.beta1 = d->beta1; a.beta2 = d->beta2; a.eps = d->eps;
    // bias corrections in double like torch.optim.Adam's python scalars
    const double bc1 = 1.0 - pow((double)d->beta1, (double)d->step);
    const double bc2 = 1.0 - pow((double)d->beta2, (double)d->step);
    a.step_size = (float)((double)d->lr / bc1);
    a.bias2_sqrt = (float)sqrt(bc2);
    a.ema_mode = d->ema_mode; a.ema_beta = d->ema_beta;
    a.wait_ns = d->wait_ns;
    a.mc_grads = d->mc_grads;
    return fused_step_launch(a, as_stream(stream));
}

}  // extern "C"
