// SinDDMNet forward / backward as a fixed sequence of sm_100a kernels over one caller-owned workspace.
//
// Mirrors SinDDMNet.forward (reference SinDDM/models.py:134-151) and SinDDMConvBlock.forward (:69-80):
//
//   cond   = fused cond kernel (embeddings + time_mlp + 4 x (mlp, time_reshape))           [cond.cu]
//   per block l:  h0 = dw5x5(x) + b + cond_l                                               [simt_misc.cu]
//                 a1 = GELU(conv3x3(h0) + b)          (z1 = pre-activation kept when training)
//                 o  = conv3x3(a1) + b + res(x)        res = 1x1 conv as extra K-slices / epilogue FMAs / identity
//   out    = final 1x1 conv, fused into l4's second conv epilogue, written NCHW.
//
// The dense convs run on tcgen05 (tc_conv.cu) when math == MATH_TF32 and the layer is a real contraction
// (Cin >= 8 and N % 16 == 0); the 3-channel layers and math == MATH_FP32 use the CUDA-core twin.
#include "net.h"

#include <stdlib.h>
#include <string.h>

namespace sinddm {

namespace {

struct Carver {
    uint8_t* base;
    size_t off;
    float* take(size_t nfloats) {
        off = align_up(off, 1024);
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += nfloats * sizeof(float);
        return p;
    }
};

inline void block_channels(int l, int dim, int channels, int* ci, int* co) {
    const int half = dim / 2;
    const int cis[4] = {channels, half, dim, dim};
    const int cos[4] = {half, dim, dim, half};
    *ci = cis[l];
    *co = cos[l];
}

inline int max_i(int a, int b) { return a > b ? a : b; }

bool use_tc_conv(int math, int Cin, int Cres, bool has_res_slices, int N) {
    if (math == MATH_FP32) return false;
    ConvProblem p;
    memset(&p, 0, sizeof(p));
    p.Cin = Cin;
    p.N = N;
    p.ntaps = 9;
    p.in_res = has_res_slices ? reinterpret_cast<const float*>(1) : nullptr;
    p.Cres = Cres;
    return tc_conv_supported(p);
}

// Rows of the first-conv data-gradient operand: a 3-channel result (l1) runs on the tensor cores with N padded to 16
// zero weight rows (3 of 16 accumulator columns are stored) instead of a CUDA-core kernel.
inline int d1_rows(int math, int Ci, int Co) {
    return (math == MATH_TF32 && Ci < 16 && use_tc_conv(math, Co, 0, false, 16)) ? 16 : Ci;
}

// l1.net[0] as a 1x1 GEMM over im2col rows (27 patch values padded to one 32-channel chunk)
inline bool use_im2col(int math, int Ci, int Co) {
    ConvProblem p;
    memset(&p, 0, sizeof(p));
    p.Cin = 32;
    p.N = Co;
    p.ntaps = 1;
    return math == MATH_TF32 && Ci == 3 && tc_conv_supported(p) && tc_wgrad_supported(32, Co);
}

bool use_tc_wgrad(int math, int Cx, int Cy) { return math != MATH_FP32 && tc_wgrad_supported(Cx, Cy); }

// MATH_TF32X3 runs a tensor-core weight gradient over the batch-tripled split operands
inline int km(int math) { return math == MATH_TF32X3 ? 3 : 1; }

int wgrad_nsplit(int math, int B, int H, int W, int Cx, int Cy, int ntaps) {
    return use_tc_wgrad(math, Cx, Cy) ? tc_wgrad_nsplit(km(math) * B, H, W, Cx, Cy, ntaps)
                                      : simt_wgrad_nsplit(B, H, W, Cx, Cy, ntaps);
}

// Lays out every buffer; with base == nullptr only the total size is computed.
size_t carve(Plan* pl, uint8_t* base) {
    Carver cv{base, 0};
    const int B = pl->B, dim = pl->dim, half = pl->half, ch = pl->channels;
    const size_t P = (size_t)pl->P;
    const bool tr = pl->training != 0;
    int csum = 0;
    for (int l = 0; l < kNumBlocks; ++l) {
        int ci, co;
        block_channels(l, dim, ch, &ci, &co);
        csum += ci;
    }
    pl->x_nhwc = cv.take(P * ch);
    pl->cond_all = cv.take((size_t)B * csum);
    pl->cond_saved = cv.take(cond_saved_floats(B));
    pl->wf_d = cv.take((size_t)half * ch);

    // activation buffers: training keeps everything; inference ping-pongs through 4 max-size buffers
    float* pool[4] = {nullptr, nullptr, nullptr, nullptr};
    if (!tr)
        for (int i = 0; i < 4; ++i) pool[i] = cv.take(P * dim);

    size_t coff = 0;
    size_t partial_max = 0;
    for (int l = 0; l < kNumBlocks; ++l) {
        BlockBufs& b = pl->blk[l];
        block_channels(l, dim, ch, &b.Ci, &b.Co);
        b.has_res = b.Ci != b.Co;
        // packed weights (sized for the blocked layout: K padded to a multiple of 32)
        const int k3 = km(pl->math);      // MATH_TF32X3 packs [lo | hi | hi] along K
        b.w0_f = cv.take(packed_weight_floats(9, b.Co, k3 * b.Ci));
        b.w2_f = cv.take(packed_weight_floats(9, b.Co, k3 * b.Co));
        b.wr_f = b.has_res ? cv.take(packed_weight_floats(1, b.Co, k3 * b.Ci)) : nullptr;
        b.w0_d = tr ? cv.take(packed_weight_floats(9, d1_rows(pl->math, b.Ci, b.Co), k3 * b.Co)) : nullptr;
        b.w2_d = tr ? cv.take(packed_weight_floats(9, b.Co, k3 * b.Co)) : nullptr;
        b.wr_d = (tr && b.has_res) ? cv.take(packed_weight_floats(1, b.Ci, k3 * b.Co)) : nullptr;
        b.bias2c = b.has_res ? cv.take((size_t)b.Co) : nullptr;
        b.im2col = use_im2col(pl->math, b.Ci, b.Co);
        b.x27 = b.im2col ? cv.take(P * 32) : nullptr;
        if (tr) {
            b.h0 = cv.take(P * b.Ci);
            b.z1 = cv.take(P * b.Co);
            b.a1 = cv.take(P * b.Co);
            b.o = cv.take(P * b.Co);
        } else {
            // four rotating slots: in_l = o_{l-1} = slot 3l, h0 = 3l+1, a1 = 3l+2, o = 3l+3 (mod 4), so the
            // tensors alive at the same time (in, h0, a1, o) never alias
            b.h0 = pool[(3 * l + 1) % 4];
            b.a1 = pool[(3 * l + 2) % 4];
            b.o = pool[(3 * l + 3) % 4];
            b.z1 = nullptr;
        }
        b.cond = pl->cond_all ? pl->cond_all + coff : nullptr;
        coff += (size_t)B * b.Ci;
        if (tr) {
            const size_t n2 = (size_t)wgrad_nsplit(pl->math, B, pl->H, pl->W, b.Co, b.Co, 9) * 9 * b.Co * b.Co;
            const size_t n0 = b.im2col ? (size_t)wgrad_nsplit(pl->math, B, pl->H, pl->W, 32, b.Co, 1) * 32 * b.Co
                                       : (size_t)wgrad_nsplit(pl->math, B, pl->H, pl->W, b.Ci, b.Co, 9) * 9 * b.Ci * b.Co;
            const size_t nr = (size_t)wgrad_nsplit(pl->math, B, pl->H, pl->W, b.Ci, b.Co, 1) * b.Ci * b.Co;
            if (n2 > partial_max) partial_max = n2;
            if (n0 > partial_max) partial_max = n0;
            if (nr > partial_max) partial_max = nr;
        }
    }
    if (pl->math == MATH_TF32X3) {
        pl->split_a = cv.take(3 * P * dim);
        pl->split_b = cv.take(3 * P * dim);
    }
    if (tr) {
        pl->dcond_all = cv.take((size_t)B * csum);
        pl->cond_scratch = cv.take(cond_bwd_scratch_floats(B));
        pl->dout_nhwc = cv.take(P * ch);
        pl->d_a = cv.take(P * dim);
        pl->d_b = cv.take(P * dim);
        pl->dz1 = cv.take(P * dim);
        pl->dh0 = cv.take(P * dim);
        pl->dxres = cv.take(P * dim);
        const size_t nf = (size_t)simt_wgrad_nsplit(B, pl->H, pl->W, ch, half, 1) * ch * half;
        if (nf > partial_max) partial_max = nf;
        pl->partial = cv.take(partial_max);
        pl->dw_scratch = cv.take(dw5x5_wgrad_scratch_floats(B, pl->H, dim));
        {
            // shared by the depthwise kernels' fused column sums, colsum_launch and tc_conv's epilogue column sums
            size_t n = dw5x5_csum_scratch_floats(B, pl->H, pl->W, dim);
            // (tc_conv: one row per CTA and TMEM lane quarter, at most one CTA per SM; 192 >= any sm_100 part)
            const size_t n_tc = (size_t)192 * 4 * dim, n_cs = colsum_scratch_floats(dim);
            const size_t n_fb = final_conv_bwd_scratch_floats(half);
            if (n_tc > n) n = n_tc;
            if (n_cs > n) n = n_cs;
            if (n_fb > n) n = n_fb;
            pl->colsum_scratch = cv.take(n);
        }
        pl->colsum_out = cv.take(dim);
        size_t doff = 0;
        for (int l = 0; l < kNumBlocks; ++l) {
            pl->blk[l].dcond = pl->dcond_all ? pl->dcond_all + doff : nullptr;
            doff += (size_t)B * pl->blk[l].Ci;
        }
    }
    return align_up(cv.off, 1024);
}

int run_conv(bool tc, TcConvOp& op, const ConvProblem& p, cudaStream_t s) {
    if (tc) {
        op.p.ep = p.ep;  // epilogue pointers may be patched per call (out_final)
        return tc_conv_launch(op, s);
    }
    return simt_conv_launch(p, s);
}

int run_wgrad(bool tc, const TcWgradOp& op, const WgradProblem& p, float* dst, cudaStream_t s, int layout = 0) {
    if (tc) {
        SINDDM_TRY(tc_wgrad_launch(op, s));
    } else {
        SINDDM_TRY(simt_wgrad_launch(p, s));
    }
    return wgrad_reduce_launch(p.partial, p.nsplit, p.ntaps, p.Cx, p.Cy, dst, layout, s);
}

}  // namespace

size_t plan_workspace_bytes(int B, int H, int W, int dim, int channels, int math, int training) {
    Plan tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.B = B;
    tmp.H = H;
    tmp.W = W;
    tmp.dim = dim;
    tmp.half = dim / 2;
    tmp.channels = channels;
    tmp.math = math;
    tmp.training = training;
    tmp.P = (long long)B * H * W;
    return carve(&tmp, nullptr);
}

int plan_build(Plan* pl, int B, int H, int W, int dim, int channels, int math, int training, void* ws,
               size_t ws_bytes) {
    SINDDM_REQUIRE(B >= 1 && H >= 1 && W >= 1, "plan: bad shape B=%d H=%d W=%d", B, H, W);
    SINDDM_REQUIRE(dim >= 2 && dim % 2 == 0 && dim <= 256, "plan: dim=%d unsupported", dim);
    SINDDM_REQUIRE(channels == 3, "plan: channels=%d unsupported (the reference always uses 3)", channels);
    SINDDM_REQUIRE(math == MATH_FP32 || math == MATH_TF32 || math == MATH_TF32X3, "plan: math mode %d unknown", math);
    SINDDM_REQUIRE(B <= 65535, "plan: batch too large");
    memset(pl, 0, sizeof(*pl));
    pl->B = B;
    pl->H = H;
    pl->W = W;
    pl->dim = dim;
    pl->half = dim / 2;
    pl->channels = channels;
    pl->math = math;
    pl->training = training;
    pl->P = (long long)B * H * W;
    const size_t need = carve(pl, nullptr);
    if (ws == nullptr || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 1023u) != 0) {
        set_error("plan: workspace %p of %zu bytes does not satisfy %zu bytes, 1024-aligned", ws, ws_bytes, need);
        return SINDDM_ERR_WORKSPACE;
    }
    pl->ws = static_cast<float*>(ws);
    pl->ws_bytes = ws_bytes;
    carve(pl, static_cast<uint8_t*>(ws));
    // blocked weight boxes are padded to 32 channels: the padding must read as zero, so clear the packed-weight
    // buffers once (the pack kernels only ever write real elements)
    const char* e2sm = getenv("SINDDM_TC_2SM");
    pl->blocked_weights = (math != MATH_FP32) && !(e2sm && atoi(e2sm) != 0);
    const int k3 = km(math);
    for (int l = 0; l < kNumBlocks; ++l) {
        BlockBufs& b = pl->blk[l];
        SINDDM_CUDA_OK(cudaMemset(b.w0_f, 0, packed_weight_floats(9, b.Co, k3 * b.Ci) * sizeof(float)));
        SINDDM_CUDA_OK(cudaMemset(b.w2_f, 0, packed_weight_floats(9, b.Co, k3 * b.Co) * sizeof(float)));
        if (b.wr_f) SINDDM_CUDA_OK(cudaMemset(b.wr_f, 0, packed_weight_floats(1, b.Co, k3 * b.Ci) * sizeof(float)));
        if (b.w0_d)
            SINDDM_CUDA_OK(cudaMemset(b.w0_d, 0, packed_weight_floats(9, d1_rows(math, b.Ci, b.Co), k3 * b.Co) * sizeof(float)));
        if (b.w2_d) SINDDM_CUDA_OK(cudaMemset(b.w2_d, 0, packed_weight_floats(9, b.Co, k3 * b.Co) * sizeof(float)));
        if (b.wr_d) SINDDM_CUDA_OK(cudaMemset(b.wr_d, 0, packed_weight_floats(1, b.Ci, k3 * b.Co) * sizeof(float)));
    }

    const bool tr = training != 0;
    const int rnd = math == MATH_TF32 ? 1 : 0;
    const bool x3 = math == MATH_TF32X3;
    // MATH_TF32X3: a tensor-core problem reads the split copy of its operand (three times the channels, or for the
    // weight gradients three times the batch); the launch sites in net_forward / net_backward fill split_a / split_b
    auto conv_x3 = [&](bool tc, ConvProblem& p) {
        if (!(x3 && tc)) return;
        p.in = pl->split_a;
        p.Cin *= 3;
        if (p.in_res) { p.in_res = pl->split_b; p.Cres *= 3; }
    };
    auto wgrad_x3 = [&](bool tc, WgradProblem& p) {
        if (!(x3 && tc)) return;
        p.x = pl->split_a;
        p.dy = pl->split_b;
        p.B *= 3;
    };
    const float* prev_o = pl->x_nhwc;
    for (int l = 0; l < kNumBlocks; ++l) {
        BlockBufs& b = pl->blk[l];
        b.in = prev_o;
        b.pbase = l < 3 ? P_BLOCK0 + 12 * l : P_BLOCK0 + 12 + 12 + 10;   // l3 has no res_conv (10 tensors)
        const bool res_slices = b.has_res && b.Ci >= 8;          // 1x1 residual conv fused as extra K
        const bool res_c3 = b.has_res && !res_slices;            // Cin = 3: epilogue FMAs

        // ---- conv1: h0 -> a1 (GELU), z1 kept when training
        ConvProblem& c1 = b.pc1;
        memset(&c1, 0, sizeof(c1));
        c1.B = B; c1.H = H; c1.W = W;
        c1.in = b.h0; c1.Cin = b.Ci; c1.w = b.w0_f; c1.ntaps = 9; c1.N = b.Co;
        if (b.im2col) { c1.in = b.x27; c1.Cin = 32; c1.ntaps = 1; }
        b.tc_c1 = b.im2col || use_tc_conv(math, b.Ci, 0, false, b.Co);
        b.tc_c2 = use_tc_conv(math, b.Co, b.Ci, res_slices, b.Co);
        c1.ep.gelu = 1; c1.ep.out = b.a1; c1.ep.out_pre = tr ? b.z1 : nullptr; c1.ep.round_tf32 = rnd && b.tc_c2;
        c1.ep.fast_math = rnd;
        c1.w_blocked = b.tc_c1 && pl->blocked_weights;
        conv_x3(b.tc_c1, c1);
        if (b.tc_c1) SINDDM_TRY(tc_conv_prepare(c1, &b.c1));

        // ---- conv2: a1 -> o, + residual; l4 also carries the final 1x1 conv
        ConvProblem& c2 = b.pc2;
        memset(&c2, 0, sizeof(c2));
        c2.B = B; c2.H = H; c2.W = W;
        c2.in = b.a1; c2.Cin = b.Co; c2.w = b.w2_f; c2.ntaps = 9; c2.N = b.Co;
        if (res_slices) { c2.in_res = b.in; c2.Cres = b.Ci; c2.w_res = b.wr_f; }
        if (res_c3) c2.ep.x3 = b.in;
        if (!b.has_res) c2.ep.res_add = b.in;
        c2.ep.out = (l < 3 || tr) ? b.o : nullptr;
        c2.w_blocked = b.tc_c2 && pl->blocked_weights;
        conv_x3(b.tc_c2, c2);
        if (b.tc_c2) SINDDM_TRY(tc_conv_prepare(c2, &b.c2));

        if (tr) {
            // ---- data gradients
            ConvProblem& d2 = b.pd2;
            memset(&d2, 0, sizeof(d2));
            d2.B = B; d2.H = H; d2.W = W;
            d2.Cin = b.Co; d2.w = b.w2_d; d2.ntaps = 9; d2.N = b.Co;
            d2.ep.dgelu_z = b.z1; d2.ep.out = pl->dz1; d2.ep.round_tf32 = rnd;
            b.tc_d2 = use_tc_conv(math, b.Co, 0, false, b.Co);
            // both ends on the tensor-core path: z1 carries gelu'(pre-activation) instead of the pre-activation
            {
                const char* e = getenv("SINDDM_PRE_GRAD");     // A/B switch; default on
                if (b.tc_c1 && b.tc_d2 && !(e && atoi(e) == 0)) { c1.ep.pre_grad = 1; d2.ep.pre_grad = 1; }
            }

            ConvProblem& d1 = b.pd1;
            memset(&d1, 0, sizeof(d1));
            d1.B = B; d1.H = H; d1.W = W;
            d1.in = pl->dz1; d1.Cin = b.Co; d1.w = b.w0_d; d1.ntaps = 9; d1.N = d1_rows(math, b.Ci, b.Co);
            if (d1.N != b.Ci) d1.ep.out3 = pl->dh0;     // padded tensor-core problem: columns 0..2 -> [P,3]
            else d1.ep.out = pl->dh0;
            b.tc_d1 = use_tc_conv(math, b.Co, 0, false, d1.N);

            ConvProblem& dr = b.pdr;
            memset(&dr, 0, sizeof(dr));
            dr.B = B; dr.H = H; dr.W = W;
            dr.Cin = b.Co; dr.w = b.wr_d; dr.ntaps = 1; dr.N = b.Ci;
            dr.ep.out = pl->dxres;
            b.tc_dr = b.has_res && l > 0 && use_tc_conv(math, b.Co, 0, false, b.Ci);

            // ---- weight gradients
            WgradProblem& w2 = b.pw2;
            memset(&w2, 0, sizeof(w2));
            w2.B = B; w2.H = H; w2.W = W; w2.x = b.a1; w2.Cx = b.Co; w2.Cy = b.Co; w2.ntaps = 9;
            w2.partial = pl->partial;
            w2.nsplit = wgrad_nsplit(math, B, H, W, b.Co, b.Co, 9);
            b.tc_w2 = use_tc_wgrad(math, b.Co, b.Co);

            WgradProblem& w0 = b.pw0;
            memset(&w0, 0, sizeof(w0));
            w0.B = B; w0.H = H; w0.W = W; w0.x = b.h0; w0.Cx = b.Ci; w0.dy = pl->dz1; w0.Cy = b.Co; w0.ntaps = 9;
            if (b.im2col) { w0.x = b.x27; w0.Cx = 32; w0.ntaps = 1; }
            w0.partial = pl->partial;
            w0.nsplit = wgrad_nsplit(math, B, H, W, w0.Cx, b.Co, w0.ntaps);
            b.tc_w0 = use_tc_wgrad(math, w0.Cx, b.Co);

            WgradProblem& wr = b.pwr;
            memset(&wr, 0, sizeof(wr));
            wr.B = B; wr.H = H; wr.W = W; wr.x = b.in; wr.Cx = b.Ci; wr.Cy = b.Co; wr.ntaps = 1;
            wr.partial = pl->partial;
            wr.nsplit = wgrad_nsplit(math, B, H, W, b.Ci, b.Co, 1);
            b.tc_wr = b.has_res && use_tc_wgrad(math, b.Ci, b.Co);
        }
        prev_o = b.o;
    }

    if (tr) {
        // The incoming gradient d_o of block l alternates between d_a and d_b (l = 3 gets d_a).
        for (int l = kNumBlocks - 1; l >= 0; --l) {
            BlockBufs& b = pl->blk[l];
            float* d_o = ((kNumBlocks - 1 - l) % 2 == 0) ? pl->d_a : pl->d_b;
            b.pd2.in = d_o;
            b.pdr.in = d_o;
            b.pw2.dy = d_o;
            b.pwr.dy = d_o;
            b.pd2.w_blocked = b.tc_d2 && pl->blocked_weights;
            b.pd1.w_blocked = b.tc_d1 && pl->blocked_weights;
            b.pdr.w_blocked = b.tc_dr && pl->blocked_weights;
            // dz1's column sums (= net[0]'s bias gradient) come out of the data-gradient kernel's epilogue
            {
                const char* e = getenv("SINDDM_TC_COLSUM");    // A/B switch; default on
                if (b.tc_d2 && !(e && atoi(e) == 0)) b.pd2.ep.colsum_part = pl->colsum_scratch;
            }
            conv_x3(b.tc_d2, b.pd2);
            conv_x3(b.tc_d1, b.pd1);
            conv_x3(b.tc_dr, b.pdr);
            wgrad_x3(b.tc_w2, b.pw2);
            wgrad_x3(b.tc_w0, b.pw0);
            wgrad_x3(b.tc_wr, b.pwr);
            if (b.tc_d2) SINDDM_TRY(tc_conv_prepare(b.pd2, &b.d2));
            if (b.tc_d1) SINDDM_TRY(tc_conv_prepare(b.pd1, &b.d1));
            if (b.tc_dr) SINDDM_TRY(tc_conv_prepare(b.pdr, &b.dr));
            if (b.tc_w2) SINDDM_TRY(tc_wgrad_prepare(b.pw2, &b.wg2));
            if (b.tc_w0) SINDDM_TRY(tc_wgrad_prepare(b.pw0, &b.wg0));
            if (b.tc_wr) SINDDM_TRY(tc_wgrad_prepare(b.pwr, &b.wgr));
        }
        // final conv: d_o4 = W_f^T dout (1x1, 3 -> half); dW_f[3][half] = sum_p dout[p][j] o4[p][c]
        ConvProblem& fd = pl->pfd;
        memset(&fd, 0, sizeof(fd));
        fd.B = B; fd.H = H; fd.W = W;
        fd.in = pl->dout_nhwc; fd.Cin = channels; fd.w = pl->wf_d; fd.ntaps = 1; fd.N = pl->half;
        fd.ep.out = pl->d_a; fd.ep.round_tf32 = rnd;
        WgradProblem& fw = pl->pfw;
        memset(&fw, 0, sizeof(fw));
        fw.B = B; fw.H = H; fw.W = W;
        fw.x = pl->dout_nhwc; fw.Cx = channels; fw.dy = pl->blk[3].o; fw.Cy = pl->half; fw.ntaps = 1;
        fw.partial = pl->partial;
        fw.nsplit = simt_wgrad_nsplit(B, H, W, channels, pl->half, 1);
    }
    return SINDDM_OK;
}

int net_pack_weights(Plan* pl, const float* const* params, cudaStream_t s) {
    // tensor-core layers: 1 = round to tf32, 2 = MATH_TF32X3's [lo | hi | hi] split along K; CUDA-core layers: 0
    const int rnd = pl->math == MATH_TF32 ? 1 : (pl->math == MATH_TF32X3 ? 2 : 0);
    struct PdlScope {
        explicit PdlScope(long long px) { pdl_set(px <= pdl_max_pixels()); }
        ~PdlScope() { pdl_set(false); }
    } pdl_scope(pl->P);
    // every repack of the step goes into ONE launch (pack_jobs_kernel)
    PackJobs jobs;
    jobs.n = 0;
    for (int l = 0; l < kNumBlocks; ++l) {
        BlockBufs& b = pl->blk[l];
        // tensor-core layers read the blocked pre-swizzled layout, CUDA-core layers the plain [tap][N][K] one
        if (b.im2col) {
            pack_jobs_add_im2col(&jobs, params[b.pbase + 6], b.Co, b.w0_f, rnd, pl->blocked_weights);
            if (pl->training)
                pack_jobs_add_conv(&jobs, params[b.pbase + 6], b.Co, b.Ci, 9, nullptr, b.w0_d, b.tc_d1 ? rnd : 0,
                                   b.tc_d1 && pl->blocked_weights, b.pd1.N);
        } else if (pl->training && b.tc_d1 != b.tc_c1) {
            // l1: the forward conv (Cin = 3) runs on CUDA cores, its data gradient on the tensor cores
            pack_jobs_add_conv(&jobs, params[b.pbase + 6], b.Co, b.Ci, 9, b.w0_f, nullptr, b.tc_c1 ? rnd : 0,
                               b.tc_c1 && pl->blocked_weights);
            pack_jobs_add_conv(&jobs, params[b.pbase + 6], b.Co, b.Ci, 9, nullptr, b.w0_d, b.tc_d1 ? rnd : 0,
                               b.tc_d1 && pl->blocked_weights, b.pd1.N);
        } else {
            pack_jobs_add_conv(&jobs, params[b.pbase + 6], b.Co, b.Ci, 9, b.w0_f, b.w0_d, b.tc_c1 ? rnd : 0,
                               b.tc_c1 && pl->blocked_weights, pl->training ? b.pd1.N : 0);
        }
        pack_jobs_add_conv(&jobs, params[b.pbase + 8], b.Co, b.Co, 9, b.w2_f, b.w2_d, b.tc_c2 ? rnd : 0,
                           b.tc_c2 && pl->blocked_weights);
        if (b.has_res) {
            if (b.Ci >= 8)
                pack_jobs_add_conv(&jobs, params[b.pbase + 10], b.Co, b.Ci, 1, b.wr_f, b.wr_d, b.tc_c2 ? rnd : 0,
                                   b.tc_c2 && pl->blocked_weights);
            // net[2].bias + res_conv.bias enter the same epilogue
            pack_jobs_add_sum(&jobs, params[b.pbase + 9], params[b.pbase + 11], b.bias2c, b.Co);
        }
    }
    if (pl->training) {
        // final_conv.0.weight [3][half] -> data-gradient layout [1][half][3]
        pack_jobs_add_conv(&jobs, params[kNumParams - 2], pl->channels, pl->half, 1, nullptr, pl->wf_d, 0);
    }
    return pack_jobs_launch(&jobs, s);
}

static void fill_cond_params(const Plan* pl, const float* const* params, CondParams* cp) {
    cp->w0 = params[P_TM0_W];
    cp->w0b = params[P_TM0_B];
    cp->w2 = params[P_TM2_W];
    cp->w2b = params[P_TM2_B];
    for (int l = 0; l < kNumBlocks; ++l) {
        const BlockBufs& b = pl->blk[l];
        cp->wm[l] = params[b.pbase + 0];
        cp->wmb[l] = params[b.pbase + 1];
        cp->wt[l] = params[b.pbase + 2];
        cp->wtb[l] = params[b.pbase + 3];
        cp->C[l] = b.Ci;
    }
}

int net_forward(Plan* pl, const float* const* params, const float* x_nchw, const long long* time, float scale,
                const float* freqs, float* out_nchw, cudaStream_t s) {
    const int B = pl->B, H = pl->H, W = pl->W;
    const int rnd = pl->math == MATH_TF32 ? 1 : 0;
    const bool x3 = pl->math == MATH_TF32X3;
    struct PdlScope {   // programmatic dependent launches only where the launch overheads matter (small problems)
        explicit PdlScope(long long px) { pdl_set(px <= pdl_max_pixels()); }
        ~PdlScope() { pdl_set(false); }
    } pdl_scope(pl->P);
    SINDDM_TRY(nchw_to_nhwc_launch(x_nchw, pl->x_nhwc, B, pl->channels, H, W, s));
    CondParams cp;
    fill_cond_params(pl, params, &cp);
    SINDDM_TRY(cond_fwd_launch(cp, time, scale, freqs, B, cond_saved_carve(pl->cond_saved, B), pl->cond_all, s));

    for (int l = 0; l < kNumBlocks; ++l) {
        BlockBufs& b = pl->blk[l];
        // h0 = ds_conv(x) + bias + cond      (models.py:70-77)
        SINDDM_TRY(dw5x5_launch(b.in, params[b.pbase + 4], params[b.pbase + 5], b.cond, nullptr, b.h0, B, H, W, b.Ci,
                                0, b.tc_c1 ? rnd : 0, s));
        if (b.im2col) SINDDM_TRY(im2col3x3_c3_launch(b.h0, b.x27, B, H, W, 0, s));   // h0 is already tf32-rounded
        // a1 = GELU(net[0](h0))              (models.py:63-64)
        b.pc1.ep.bias = params[b.pbase + 7];
        if (x3 && b.tc_c1) SINDDM_TRY(split3_launch(b.h0, pl->P, b.Ci, pl->split_a, 0, s));
        SINDDM_TRY(run_conv(b.tc_c1, b.c1, b.pc1, s));
        // o = net[2](a1) + res_conv(x)       (models.py:65,80)
        ConvEpilogue& e2 = b.pc2.ep;
        e2.bias = b.has_res ? b.bias2c : params[b.pbase + 9];
        if (e2.x3) e2.w_res3 = params[b.pbase + 10];
        if (l == kNumBlocks - 1) {  // final_conv (models.py:130-132,151) rides in the epilogue
            e2.w_final = params[kNumParams - 2];
            e2.b_final = params[kNumParams - 1];
            e2.out_final = out_nchw;
        }
        if (x3 && b.tc_c2) {
            SINDDM_TRY(split3_launch(b.a1, pl->P, b.Co, pl->split_a, 0, s));
            if (b.pc2.in_res) SINDDM_TRY(split3_launch(b.in, pl->P, b.Ci, pl->split_b, 0, s));
        }
        SINDDM_TRY(run_conv(b.tc_c2, b.c2, b.pc2, s));
    }
    return SINDDM_OK;
}

int net_backward(Plan* pl, const float* const* params, const float* dout_nchw, float* const* grads, cudaStream_t s) {
    SINDDM_REQUIRE(pl->training, "net_backward called on an inference plan");
    const int B = pl->B, H = pl->H, W = pl->W;
    const long long P = pl->P;
    const int rnd = pl->math == MATH_TF32 ? 1 : 0;
    const bool x3 = pl->math == MATH_TF32X3;
    struct PdlScope {
        explicit PdlScope(long long px) { pdl_set(px <= pdl_max_pixels()); }
        ~PdlScope() { pdl_set(false); }
    } pdl_scope(P);
    // MATH_TF32X3: split operands of the next tensor-core launch (conv: channels tripled; wgrad: batch tripled)
    auto split_conv = [&](bool tc, const float* src, int C) -> int {
        return (x3 && tc) ? split3_launch(src, P, C, pl->split_a, 0, s) : SINDDM_OK;
    };
    auto split_wgrad = [&](bool tc, const float* x, int Cx, const float* dy, int Cy) -> int {
        if (!(x3 && tc)) return SINDDM_OK;
        SINDDM_TRY(split3_launch(x, P, Cx, pl->split_a, 1, s));
        return split3_launch(dy, P, Cy, pl->split_b, 2, s);
    };

    // ---- final_conv: bias / weight gradients and the gradient into l4's output
    SINDDM_TRY(nchw_to_nhwc_launch(dout_nchw, pl->dout_nhwc, B, pl->channels, H, W, s));
    const bool fused_final = pl->channels == 3 && final_conv_bwd_supported(pl->half);
    if (fused_final) {
        // one pass: d_o4 -> d_a, final_conv weight / bias gradients, and l4.net[2]'s (and res_conv's) bias gradient
        const BlockBufs& b4 = pl->blk[kNumBlocks - 1];
        SINDDM_TRY(final_conv_bwd_launch(pl->dout_nhwc, b4.o, params[kNumParams - 2], pl->d_a, P, pl->half, rnd,
                                         grads[kNumParams - 2], grads[kNumParams - 1], grads[b4.pbase + 9],
                                         b4.has_res ? grads[b4.pbase + 11] : nullptr, pl->colsum_scratch, s));
    } else {
        SINDDM_TRY(colsum_launch(pl->dout_nhwc, P, pl->channels, grads[kNumParams - 1], pl->colsum_scratch, s));
        SINDDM_TRY(simt_wgrad_launch(pl->pfw, s));
        SINDDM_TRY(wgrad_reduce_launch(pl->partial, pl->pfw.nsplit, 1, pl->channels, pl->half, grads[kNumParams - 2], 1, s));
        SINDDM_TRY(simt_conv_launch(pl->pfd, s));
    }

    for (int l = kNumBlocks - 1; l >= 0; --l) {
        BlockBufs& b = pl->blk[l];
        float* d_o = ((kNumBlocks - 1 - l) % 2 == 0) ? pl->d_a : pl->d_b;
        float* d_prev = (d_o == pl->d_a) ? pl->d_b : pl->d_a;

        // bias gradients of net[2] (and res_conv, identical sum): for l < 3 they were accumulated by the depthwise
        // data-gradient launch that produced d_o (end of the previous iteration)
        if (l == kNumBlocks - 1 && !fused_final) {
            SINDDM_TRY(colsum_launch(d_o, P, b.Co, grads[b.pbase + 9], pl->colsum_scratch, s));
            if (b.has_res)
                SINDDM_CUDA_OK(cudaMemcpyAsync(grads[b.pbase + 11], grads[b.pbase + 9], sizeof(float) * b.Co,
                                               cudaMemcpyDeviceToDevice, s));
        }
        // weight gradients of net[2] and res_conv
        SINDDM_TRY(split_wgrad(b.tc_w2, b.a1, b.Co, d_o, b.Co));
        SINDDM_TRY(run_wgrad(b.tc_w2, b.wg2, b.pw2, grads[b.pbase + 8], s));
        if (b.has_res) {
            if (!b.tc_wr && b.Ci == 3 && final_conv_bwd_supported(b.Co)) {  // l1.res_conv: one streaming pass over d_o
                SINDDM_TRY(wgrad_from_c3_launch(b.in, d_o, P, b.Co, grads[b.pbase + 10], pl->colsum_scratch, s));
            } else {
                SINDDM_TRY(split_wgrad(b.tc_wr, b.in, b.Ci, d_o, b.Co));
                SINDDM_TRY(run_wgrad(b.tc_wr, b.wgr, b.pwr, grads[b.pbase + 10], s));
            }
        }
        // dz1 = conv3x3^T(d_o) * gelu'(z1)
        SINDDM_TRY(split_conv(b.tc_d2, d_o, b.Co));
        SINDDM_TRY(run_conv(b.tc_d2, b.d2, b.pd2, s));
        if (b.tc_d2 && b.pd2.ep.colsum_part)
            SINDDM_TRY(colsum_final_launch(pl->colsum_scratch, tc_conv_colsum_rows(b.d2), b.Co, grads[b.pbase + 7], s));
        else
            SINDDM_TRY(colsum_launch(pl->dz1, P, b.Co, grads[b.pbase + 7], pl->colsum_scratch, s));
        SINDDM_TRY(split_wgrad(b.tc_w0, b.h0, b.Ci, pl->dz1, b.Co));      // (im2col never coexists with MATH_TF32X3)
        SINDDM_TRY(run_wgrad(b.tc_w0, b.wg0, b.pw0, grads[b.pbase + 6], s, b.im2col ? 2 : 0));
        // dh0 = conv3x3^T(dz1)
        SINDDM_TRY(split_conv(b.tc_d1, pl->dz1, b.Co));
        SINDDM_TRY(run_conv(b.tc_d1, b.d1, b.pd1, s));
        // depthwise weight / bias / conditioning gradients
        SINDDM_TRY(dw5x5_wgrad_launch(b.in, pl->dh0, grads[b.pbase + 4], grads[b.pbase + 5], b.dcond, pl->dw_scratch,
                                      B, H, W, b.Ci, s));
        if (l > 0) {
            // gradient into the block input: depthwise^T(dh0) + residual path
            const float* addp = d_o;
            if (b.has_res) {
                SINDDM_TRY(split_conv(b.tc_dr, d_o, b.Co));
                SINDDM_TRY(run_conv(b.tc_dr, b.dr, b.pdr, s));
                addp = pl->dxres;
            }
            const BlockBufs& pb = pl->blk[l - 1];   // d_prev is its upstream gradient
            SINDDM_TRY(dw5x5_launch(pl->dh0, params[b.pbase + 4], nullptr, nullptr, addp, d_prev, B, H, W, b.Ci, 1,
                                    rnd, s, grads[pb.pbase + 9], pb.has_res ? grads[pb.pbase + 11] : nullptr,
                                    pl->colsum_scratch));
        }
    }

    // ---- conditioning path
    CondParams cp;
    fill_cond_params(pl, params, &cp);
    CondGrads cg;
    cg.w0 = grads[P_TM0_W];
    cg.w0b = grads[P_TM0_B];
    cg.w2 = grads[P_TM2_W];
    cg.w2b = grads[P_TM2_B];
    for (int l = 0; l < kNumBlocks; ++l) {
        const BlockBufs& b = pl->blk[l];
        cg.wm[l] = grads[b.pbase + 0];
        cg.wmb[l] = grads[b.pbase + 1];
        cg.wt[l] = grads[b.pbase + 2];
        cg.wtb[l] = grads[b.pbase + 3];
    }
    SINDDM_TRY(cond_bwd_launch(cp, cg, B, cond_saved_carve(pl->cond_saved, B), pl->dcond_all, pl->cond_scratch, s));
    return SINDDM_OK;
}

}  // namespace sinddm
