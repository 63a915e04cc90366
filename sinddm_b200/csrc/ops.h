// Internal operator interface between the network driver (net.cu / capi.cu) and the kernels.
//
// All activations inside the library are NHWC fp32 ([B,H,W,C], "P = B*H*W pixels x C channels"); only the
// 3-channel boundary tensors (network input / output, diffusion state) are NCHW like the reference's.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "host_common.h"

namespace sinddm {

// ------------------------------------------------------------------------------------------------
// Dense convolution as a GEMM:  out[p][n] = sum_{tap, c} in[p (+) tap][c] * w[tap][n][c]  (+ residual K-slices)
//   ntaps = 9 (3x3, pad 1, tap = ky*3+kx, offset (ky-1, kx-1)) or 1 (1x1).
// The same problem description drives the tcgen05 kernel (tc_conv.cu) and the fp32 CUDA-core kernel
// (simt_conv.cu); forward convolutions and data-gradients differ only in how the weights were packed.
// ------------------------------------------------------------------------------------------------
struct ConvEpilogue {
    const float* bias;      // [N] added to every pixel, or null
    const float* res_add;   // [P,N] identity residual added before activation, or null
    const float* x3;        // [P,3] NHWC: 1x1 residual conv from a 3-channel input ...
    const float* w_res3;    // ... with weights [N,3]; both null when unused
    int gelu;               // apply exact-erf GELU
    float* out_pre;         // [P,N] pre-activation copy (saved for backward), or null
    const float* dgelu_z;   // [P,N]: multiply result by gelu'(z) (data-gradient epilogue), or null
    const float* w_final;   // [3,N] fused trailing 1x1 conv to 3 channels ...
    const float* b_final;   // ... bias [3] ...
    float* out_final;       // ... written NCHW [B,3,H,W]; all null when unused
    int round_tf32;         // round `out` to tf32 (its only consumers are tensor-core operands)
    float* out;             // [P,N], or null when only out_final / out3 is wanted
    float* out3;            // tensor-core path only: [P,3] NHWC copy of result columns 0..2 (a 3-channel result computed
                            // with N padded to 16 zero weight rows), or null
    int pre_grad;           // tensor-core path: out_pre receives gelu'(pre-activation) instead of the pre-activation, and
                            // dgelu_z already holds gelu'(z) (the forward epilogue has cdf and pdf at hand: two extra
                            // instructions there replace ~25 + two MUFU per value in the data-gradient epilogue)
    int fast_math;          // CUDA-core kernels: use the TF32-mode GELU (gelu_fast) instead of erff (set with math = tf32)
    float* colsum_part;     // tensor-core path only: [grid * 4][N] per-(CTA, TMEM lane quarter) column sums of what the
                            // kernel stores to `out` (the bias gradient of the layer below: finish with
                            // colsum_final_launch(colsum_part, tc_conv_colsum_rows(op), N, ...)), or null
};

struct ConvProblem {
    int B, H, W;
    const float* in;      // [P,Cin]
    int Cin;
    const float* w;       // [ntaps][N][Cin]
    int ntaps;
    const float* in_res;  // [P,Cres] input of a fused 1x1 residual conv (extra K-slices), or null
    int Cres;
    const float* w_res;   // [N][Cres]
    int N;
    // w_blocked != 0: `w` / `w_res` are in the tensor-core staging layout [tap][chunk][N][32] with the 128-byte
    // rows pre-swizzled (pack_conv_weights_launch(..., blocked=1)); every (tap, 32-channel chunk) weight box is
    // then one contiguous N x 128 B block that a single 1-D bulk copy brings into shared memory.
    int w_blocked;
    ConvEpilogue ep;
};

// tcgen05 path: needs Cin % 8 == 0, Cres % 8 == 0, N % 16 == 0, 16 <= N <= 160.
struct TcConvOp {
    CUtensorMap tm_a, tm_ares, tm_b, tm_bres;
    CUtensorMap tm_out, tm_pre, tm_in;   // epilogue: result / pre-activation stores, streamed-operand loads
    ConvProblem p;
    int stage_bytes, nstages, smem_bytes;
    int tiles_w, tiles_h, ntiles, grid;
    int cs;   // cluster size (CTAs sharing each weight stage through TMA multicast)
    int halo, nh, abw, nsa, abytes, atx;   // activation ring: one halo box per chunk for all nine taps?  slots, slot stride, box bytes
    // A/B and diagnostic switches (SINDDM_TC_*), read from the environment ONCE by tc_conv_prepare
    int sw_peek, sw_l2pf, sw_issuers2, sw_dbg, sw_stage_release;
};
// rows of ConvEpilogue::colsum_part this op writes (a multiple of its grid size)
int tc_conv_colsum_rows(const TcConvOp& op);
// out[c] = sum over the nrows partial rows, fixed order
int colsum_final_launch(const float* part, int nrows, int C, float* out, cudaStream_t stream);
bool tc_conv_supported(const ConvProblem& p);
int tc_conv_prepare(const ConvProblem& p, TcConvOp* op);
int tc_conv_launch(const TcConvOp& op, cudaStream_t stream);
// compile-time epilogue-feature mask of the tc_conv_kernel instantiation a launch with this epilogue runs
// (tc_conv.cu: kFl*; 2047 = the generic instantiation)
int tc_conv_flavour_mask(const ConvEpilogue& ep);

int simt_conv_launch(const ConvProblem& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Weight gradient:  partial[split][tap][ci][co] = sum_{p in split} x[p (+) tap][ci] * dy[p][co]
// followed by a deterministic reduction over splits into the PyTorch parameter layout.
// ------------------------------------------------------------------------------------------------
struct WgradProblem {
    int B, H, W;
    const float* x;   // [P,Cx]
    int Cx;
    const float* dy;  // [P,Cy]
    int Cy;
    int ntaps;        // 9 or 1
    float* partial;   // [nsplit][ntaps][Cx][Cy] scratch
    int nsplit;       // chosen by *_plan
};

struct TcWgradOp {
    CUtensorMap tm_x, tm_dy;
    WgradProblem p;
    int cxk, cyk, nregions, nreg_cta, ntile, ngroups, wt, nks, pw, rb, col_stride, tmem_cols;
    int x_bytes, stage_bytes, nstages, smem_bytes;
};
bool tc_wgrad_supported(int Cx, int Cy);
int tc_wgrad_nsplit(int B, int H, int W, int Cx, int Cy, int ntaps);
int tc_wgrad_prepare(const WgradProblem& p, TcWgradOp* op);
int tc_wgrad_launch(const TcWgradOp& op, cudaStream_t stream);

int simt_wgrad_nsplit(int B, int H, int W, int Cx, int Cy, int ntaps);
int simt_wgrad_launch(const WgradProblem& p, cudaStream_t stream);

// dst (OIHW: [Cy][Cx][ntaps]) = sum_split partial;  keep_layout = 1 keeps [ntaps][Cx][Cy], 2 = im2col rows (below).
int wgrad_reduce_launch(const float* partial, int nsplit, int ntaps, int Cx, int Cy, float* dst, int keep_layout,
                        cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Weight packing (PyTorch OIHW -> GEMM operand layouts), tf32-rounded when round != 0.
//   fwd : dst[tap][co][ci] = w[co][ci][tap]
//   dgrad: dst[tap][ci][co] = w[co][ci][ntaps-1-tap]
// ------------------------------------------------------------------------------------------------
// dgrad_rows > Cin pads the data-gradient operand to that many rows (the extra rows are never written: clear them once).
int pack_conv_weights_launch(const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad, int round,
                             cudaStream_t stream, int blocked = 0, int dgrad_rows = 0);
// 3-channel 3x3 conv as a 1x1 GEMM over im2col rows: out[p][tap*3 + c] = x[p (+) tap][c] ([P,32], columns 27..31
// zero); weights [Co][3][3][3] -> [1][Co][32]; wgrad_reduce_launch(..., keep_layout = 2) maps the [32][Co]
// gradient of that GEMM back to OIHW.
int im2col3x3_c3_launch(const float* x, float* out, int B, int H, int W, int round, cudaStream_t stream);
int pack_im2col_weights_launch(const float* w, int Co, float* dst, int round, int blocked, cudaStream_t stream);
// All weight-packing work of one step in ONE launch (they used to be ~15 launches of a few microseconds each, every
// training step, at every scale).  kind 0: pack_conv_weights, 1: pack_im2col_weights, 2: o = a + b (bias sums).
struct PackJob {
    int kind;
    const float* w;        // source (kind 2: a)
    const float* w2;       // kind 2: b
    float* dst_fwd;        // kind 1: dst; kind 2: o
    float* dst_dgrad;
    int Cout, Cin, ntaps, round, blocked, dgrad_rows;
    int total;             // elements (threads) of this job
    int block0;            // first 256-thread block of this job (filled by the launcher)
};
constexpr int kMaxPackJobs = 20;
struct PackJobs {
    PackJob j[kMaxPackJobs];
    int n;
};
void pack_jobs_add_conv(PackJobs* jobs, const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad,
                        int round, int blocked = 0, int dgrad_rows = 0);
void pack_jobs_add_im2col(PackJobs* jobs, const float* w, int Co, float* dst, int round, int blocked);
void pack_jobs_add_sum(PackJobs* jobs, const float* a, const float* b, float* o, int n);
int pack_jobs_launch(PackJobs* jobs, cudaStream_t stream);

// 3xTF32 operand split (MATH_TF32X3): hi = tf32(x), lo = tf32(x - hi).  mode 0: out [P][3C] = [hi | lo | hi];
// mode 1: out [3][P][C] = hi, lo, hi;  mode 2: out [3][P][C] = lo, hi, hi.  Weights are packed to match with
// round == 2 in pack_conv_weights / pack_jobs_add_conv ([lo | hi | hi] along K, K tripled): cross terms first.
int split3_launch(const float* x, long long P, int C, float* out, int mode, cudaStream_t stream);

// floats needed by a packed weight buffer in either layout ([ntaps][N][K] or blocked with K padded to 32)
inline size_t packed_weight_floats(int ntaps, int N, int K) { return (size_t)ntaps * N * ((K + 31) / 32 * 32); }

// ------------------------------------------------------------------------------------------------
// Depthwise 5x5 (pad 2):  out[p][c] = add[p][c] + bias[c] + cond[b][c] + sum_tap w[c][tap(') ] * in[p (+) tap][c]
//   flip = 1 correlates with the flipped kernel (data gradient).
// ------------------------------------------------------------------------------------------------
// csum_out != null: also csum_out[c] (and csum_out2[c]) = sum_p out[p][c], accumulated by the same kernel pass;
// csum_scratch then needs dw5x5_csum_scratch_floats(B,H,W,C) floats.
size_t dw5x5_csum_scratch_floats(int B, int H, int W, int C);
int dw5x5_launch(const float* in, const float* w /*[C][25]*/, const float* bias, const float* cond /*[B][C]*/,
                 const float* add, float* out, int B, int H, int W, int C, int flip, int round_tf32,
                 cudaStream_t stream, float* csum_out = nullptr, float* csum_out2 = nullptr,
                 float* csum_scratch = nullptr);

// dW[c][tap] = sum_p x[p(+)tap][c] dh[p][c];  db[c] = sum_p dh[p][c];  dcond[b][c] = sum_hw dh[b,hw][c].
// scratch needs dw5x5_wgrad_scratch_floats(B,H,C) floats.
size_t dw5x5_wgrad_scratch_floats(int B, int H, int C);
int dw5x5_wgrad_launch(const float* x, const float* dh, float* dw, float* db, float* dcond, float* scratch, int B,
                       int H, int W, int C, cudaStream_t stream);

// final_conv (1x1, C -> 3) backward in one pass: d_o = W^T dout (tf32-rounded when round), dw [3][C], db [3], and
// db_prev [C] (+ optional copy db_prev2) = column sums of the UNROUNDED d_o = W^T db.  dout_nhwc3 [P,3], o / d_o [P,C],
// w [3][C].  scratch: final_conv_bwd_scratch_floats(C) floats.
size_t final_conv_bwd_scratch_floats(int C);
bool final_conv_bwd_supported(int C);
int final_conv_bwd_launch(const float* dout_nhwc3, const float* o, const float* w, float* d_o, long long P, int C,
                          int round, float* dw, float* db, float* db_prev, float* db_prev2, float* scratch,
                          cudaStream_t stream);

// weight gradient of a 1x1 conv FROM a 3-channel input (l1.res_conv): dw [C][3] = sum_p dy[p][c] * x3[p][j]; one read of
// dy [P,C]; scratch as for final_conv_bwd_launch.
int wgrad_from_c3_launch(const float* x3, const float* dy, long long P, int C, float* dw, float* scratch,
                         cudaStream_t stream);

// out[c] = sum_p a[p][c]; scratch needs colsum_scratch_floats(C) floats.
size_t colsum_scratch_floats(int C);
int colsum_launch(const float* a, long long P, int C, float* out, float* scratch, cudaStream_t stream);

// 3-channel boundary layout changes.
int nchw_to_nhwc_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t stream);
int nhwc_to_nchw_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t stream);

}  // namespace sinddm
