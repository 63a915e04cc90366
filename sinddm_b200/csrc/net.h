// Whole-denoiser plan: shapes, workspace carve-up, prepared tensor-core ops, forward / backward drivers.
#pragma once

#include "ops.h"

namespace sinddm {

constexpr int kNumBlocks = 4;     // l1..l4 (reference SinDDM/models.py:124-127)
constexpr int kTimeDim = 32;      // time_dim (models.py:101)
constexpr int kNumParams = 52;    // SinDDMNet.parameters() for multiscale=True, in registration order

// Index of every parameter tensor inside the `params` / `grads` pointer arrays of the C ABI.  The order is
// exactly SinDDMNet.state_dict() order (models.py:104-132 / :54-67), so a Python caller passes
// [p.data_ptr() for p in net.parameters()].
enum ParamIndex {
    P_TM0_W = 0, P_TM0_B, P_TM2_W, P_TM2_B,   // time_mlp.0 / time_mlp.2
    P_BLOCK0 = 4,                              // then per block (12 tensors, l3 has no res_conv -> 10):
    // +0 mlp.1.weight  +1 mlp.1.bias  +2 time_reshape.weight  +3 time_reshape.bias  +4 ds_conv.weight
    // +5 ds_conv.bias  +6 net.0.weight  +7 net.0.bias  +8 net.2.weight  +9 net.2.bias  +10 res_conv.weight
    // +11 res_conv.bias
};

struct CondParams {
    const float *w0, *w0b, *w2, *w2b;
    const float* wm[kNumBlocks];
    const float* wmb[kNumBlocks];
    const float* wt[kNumBlocks];
    const float* wtb[kNumBlocks];
    int C[kNumBlocks];
};
struct CondGrads {
    float *w0, *w0b, *w2, *w2b;
    float* wm[kNumBlocks];
    float* wmb[kNumBlocks];
    float* wt[kNumBlocks];
    float* wtb[kNumBlocks];
};
struct CondSaved {
    float *emb, *h1, *cv, *m;  // [B,64] [B,128] [B,32] [4][B,32]
};
size_t cond_saved_floats(int B);
size_t cond_bwd_scratch_floats(int B);
CondSaved cond_saved_carve(float* base, int B);
int cond_fwd_launch(const CondParams& P, const long long* time, float scale, const float* freqs, int B,
                    const CondSaved& S, float* cond, cudaStream_t stream);
int cond_bwd_launch(const CondParams& P, const CondGrads& G, int B, const CondSaved& S, const float* dcond,
                    float* scratch, cudaStream_t stream);

// MATH_TF32X3: the tensor-core kernels of MATH_TF32 fed with hi/lo-split operands (x = hi + lo, both tf32): every
// contraction adds x_hi w_hi + x_hi w_lo + x_lo w_hi into its fp32 accumulator -- fp32-class accuracy at three times
// the MMA work (the reference's behaviour with torch.backends.cudnn.allow_tf32 = False, on the tensor cores).
enum MathMode { MATH_FP32 = 0, MATH_TF32 = 1, MATH_TF32X3 = 2 };

struct BlockBufs {
    int Ci, Co;
    bool has_res;         // res_conv is a 1x1 conv (Ci != Co); identity otherwise
    int pbase;            // index of mlp.1.weight in the param arrays
    // packed weights
    float *w0_f, *w0_d, *w2_f, *w2_d, *wr_f, *wr_d;
    float* bias2c;        // net[2].bias + res_conv.bias
    // l1 in TF32 mode: net[0] (3 -> Co, 3x3) runs as a 1x1 tensor-core GEMM over the im2col rows x27 [P,32]
    bool im2col;
    float* x27;
    // activations ([P, C] NHWC)
    float *h0, *z1, *a1, *o;
    const float* in;      // block input (previous block's o, or x_nhwc)
    float* cond;          // [B, Ci]
    float* dcond;         // [B, Ci]
    // prepared tensor-core ops (valid when the matching use_tc_* flag is set)
    bool tc_c1, tc_c2, tc_d2, tc_d1, tc_dr, tc_w2, tc_w0, tc_wr;
    TcConvOp c1, c2, d2, d1, dr;
    TcWgradOp wg2, wg0, wgr;
    ConvProblem pc1, pc2, pd2, pd1, pdr;
    WgradProblem pw2, pw0, pwr;
};

struct Plan {
    int B, H, W, dim, half, channels;
    int math, training;
    int blocked_weights;   // tensor-core layers use the blocked pre-swizzled weight layout (1-D bulk copies)
    long long P;
    float* ws;
    size_t ws_bytes;
    // shared buffers
    float *x_nhwc, *cond_all, *dcond_all, *cond_saved, *cond_scratch;
    float *wf_d;                 // final conv data-gradient weights [1][half][3]
    float *dout_nhwc, *d_a, *d_b, *dz1, *dh0, *dxres;
    float *partial, *dw_scratch, *colsum_scratch, *colsum_out;
    float *split_a, *split_b;    // MATH_TF32X3: the split operands of the next tensor-core launch ([P][3C] or [3P][C])
    BlockBufs blk[kNumBlocks];
    // final conv gradient problems
    ConvProblem pfd;
    WgradProblem pfw;
};

size_t plan_workspace_bytes(int B, int H, int W, int dim, int channels, int math, int training);
int plan_build(Plan* plan, int B, int H, int W, int dim, int channels, int math, int training, void* ws,
               size_t ws_bytes);
int net_pack_weights(Plan* plan, const float* const* params, cudaStream_t stream);
int net_forward(Plan* plan, const float* const* params, const float* x_nchw, const long long* time, float scale,
                const float* freqs, float* out_nchw, cudaStream_t stream);
int net_backward(Plan* plan, const float* const* params, const float* dout_nchw, float* const* grads,
                 cudaStream_t stream);

}  // namespace sinddm
