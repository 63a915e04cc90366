// Conditioning path of the denoiser: sinusoidal (t, scale) embeddings -> time_mlp -> per-block mlp + 1x1
// time_reshape, producing one bias vector [B, C_l] per conv block (reference SinDDM/models.py:39-46,
// 106-110,137-141 and :54-60,74-76).  ~30 k parameters and B <= 128 rows: latency-bound, so the whole
// forward is ONE kernel (one CTA per sample) and the backward is two (per-sample chain + weight sums),
// instead of ~25 + ~50 library launches.
#include "common.cuh"
#include "net.h"

namespace sinddm {

namespace {

constexpr int TD = kTimeDim;      // 32
constexpr int ED = 2 * kTimeDim;  // 64  cat(t_emb, s_emb)
constexpr int HD = 4 * kTimeDim;  // 128 hidden of time_mlp

__global__ void __launch_bounds__(128)
cond_fwd_kernel(const CondParams P, const long long* __restrict__ time, float scale, const float* __restrict__ freqs,
                int B, CondSaved S, float* __restrict__ cond) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float emb[ED], g1[HD], g2[TD];
    const int b = blockIdx.x, t = threadIdx.x;

    if (t < ED) {
        const int half = TD / 2;                           // 16 frequencies
        const bool is_scale = t >= TD;
        const int k = (t % TD) % half;
        const bool is_cos = (t % TD) >= half;
        const float xv = is_scale ? scale : (float)time[b];
        const float arg = xv * freqs[k];
        const float e = is_cos ? cosf(arg) : sinf(arg);
        emb[t] = e;
        S.emb[(size_t)b * ED + t] = e;
    }
    __syncthreads();
    {   // time_mlp.0 : Linear(64 -> 128), then GELU
        float acc = P.w0b[t];
        const float* wr = P.w0 + (size_t)t * ED;
#pragma unroll 32
        for (int k = 0; k < ED; ++k) acc = fmaf(wr[k], emb[k], acc);
        S.h1[(size_t)b * HD + t] = acc;
        g1[t] = gelu_erf(acc);
    }
    __syncthreads();
    if (t < TD) {  // time_mlp.2 : Linear(128 -> 32) = cond_vec; every block starts with GELU(cond_vec)
        float acc = P.w2b[t];
        const float* wr = P.w2 + (size_t)t * HD;
#pragma unroll 32
        for (int k = 0; k < HD; ++k) acc = fmaf(wr[k], g1[k], acc);
        S.cv[(size_t)b * TD + t] = acc;
        g2[t] = gelu_erf(acc);
    }
    __syncthreads();
    // The four blocks' conditioning heads only depend on g2: warp l computes block l's mlp[1] (Linear 32 -> 32), then
    // all sum(C_l) time_reshape outputs (Conv1x1 32 -> C on a 1x1 image) are spread over the block.  Same summation
    // order per output as the sequential form (bias first, k ascending): bit-identical results, two phases instead of
    // eight (this kernel is on the critical path of every sampling step: 246 launches per image batch).
    static_assert(kNumBlocks * TD == 128, "cond_fwd_kernel runs 128 threads: one warp per block's mlp[1]");
    __shared__ float m_all[kNumBlocks][TD];
    {
        const int l = t >> 5, j = t & 31;   // blockDim.x == 128 == kNumBlocks * TD
        float acc = P.wmb[l][j];
        const float* wr = P.wm[l] + (size_t)j * TD;
        float wv[TD];
#pragma unroll
        for (int k = 0; k < TD; ++k) wv[k] = wr[k];
#pragma unroll
        for (int k = 0; k < TD; ++k) acc = fmaf(wv[k], g2[k], acc);
        m_all[l][j] = acc;
        S.m[((size_t)l * B + b) * TD + j] = acc;
    }
    __syncthreads();
    int ctot = 0;
    for (int l = 0; l < kNumBlocks; ++l) ctot += P.C[l];
    for (int o = t; o < ctot; o += blockDim.x) {
        int l = 0, c = o;
        size_t coff = 0;
        while (c >= P.C[l]) {
            c -= P.C[l];
            coff += (size_t)B * P.C[l];
            ++l;
        }
        const int C = P.C[l];
        float acc = P.wtb[l][c];
        const float* wr = P.wt[l] + (size_t)c * TD;
        float wv[TD];
#pragma unroll
        for (int k = 0; k < TD; ++k) wv[k] = wr[k];
#pragma unroll
        for (int k = 0; k < TD; ++k) acc = fmaf(wv[k], m_all[l][k], acc);
        cond[coff + (size_t)b * C + c] = acc;
    }
}

// per-sample backward chain: dcond -> dm_l, dcv, dh1  (written to scratch for the weight-sum kernel)
__global__ void __launch_bounds__(128)
cond_bwd_chain_kernel(const CondParams P, int B, CondSaved S, const float* __restrict__ dcond,
                      float* __restrict__ dm /*[4][B][32]*/, float* __restrict__ dcv /*[B][32]*/,
                      float* __restrict__ dh1 /*[B][128]*/) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float s_dm[TD], s_dg2[TD], s_dcv[TD], s_part[4][TD];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < TD) s_dg2[t] = 0.f;
    __syncthreads();
    size_t coff = 0;
    for (int l = 0; l < kNumBlocks; ++l) {
        const int C = P.C[l];
        {
            // dm = Wt^T dcond: the channel range is split over the block's four warps (this loop used to run on one
            // warp as a chain of C dependent global loads + FMAs: 80 us per training step at every scale), partial sums
            // are combined in a fixed order
            const int part = t >> 5, lane = t & 31;
            const int per = (C + 3) / 4;
            const int c0 = part * per, c1 = min(C, c0 + per);
            const float* g = dcond + coff + (size_t)b * C;
            const float* w = P.wt[l] + lane;
            float a0 = 0.f, a1 = 0.f;
            int c = c0;
            for (; c + 1 < c1; c += 2) {
                a0 = fmaf(w[(size_t)c * TD], g[c], a0);
                a1 = fmaf(w[(size_t)(c + 1) * TD], g[c + 1], a1);
            }
            if (c < c1) a0 = fmaf(w[(size_t)c * TD], g[c], a0);
            s_part[part][lane] = a0 + a1;
        }
        __syncthreads();
        if (t < TD) {
            const float acc = ((s_part[0][t] + s_part[1][t]) + s_part[2][t]) + s_part[3][t];
            s_dm[t] = acc;
            dm[((size_t)l * B + b) * TD + t] = acc;
        }
        __syncthreads();
        if (t < TD) {
            float acc = s_dg2[t];
            for (int j = 0; j < TD; ++j) acc = fmaf(P.wm[l][(size_t)j * TD + t], s_dm[j], acc);
            s_dg2[t] = acc;
        }
        __syncthreads();
        coff += (size_t)B * C;
    }
    if (t < TD) {
        const float v = s_dg2[t] * gelu_erf_grad(S.cv[(size_t)b * TD + t]);
        s_dcv[t] = v;
        dcv[(size_t)b * TD + t] = v;
    }
    __syncthreads();
    {
        float acc = 0.f;
        for (int j = 0; j < TD; ++j) acc = fmaf(P.w2[(size_t)j * HD + t], s_dcv[j], acc);
        dh1[(size_t)b * HD + t] = acc * gelu_erf_grad(S.h1[(size_t)b * HD + t]);
    }
}

// dW[r][k] = sum_b L[b][r] * f(R[b][k]),  db[r] = sum_b L[b][r];  f = identity or GELU
struct OuterJob {
    const float* L;
    const float* R;
    int rows, cols, gelu_r;
    float* dW;
    float* db;
};
struct OuterJobs {
    OuterJob j[2 + 2 * kNumBlocks];
};

__global__ void cond_bwd_weights_kernel(const OuterJobs jobs, int B) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const OuterJob J = jobs.j[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J.rows * (J.cols + 1)) return;
    const int r = i / (J.cols + 1), k = i % (J.cols + 1);
    float acc = 0.f;
    if (k == J.cols) {
        for (int b = 0; b < B; ++b) acc += J.L[(size_t)b * J.rows + r];
        J.db[r] = acc;
    } else {
        for (int b = 0; b < B; ++b) {
            float rv = J.R[(size_t)b * J.cols + k];
            if (J.gelu_r) rv = gelu_erf(rv);
            acc = fmaf(J.L[(size_t)b * J.rows + r], rv, acc);
        }
        J.dW[(size_t)r * J.cols + k] = acc;
    }
}

}  // namespace

size_t cond_saved_floats(int B) { return (size_t)B * (ED + HD + TD + kNumBlocks * TD); }
size_t cond_bwd_scratch_floats(int B) { return (size_t)B * (kNumBlocks * TD + TD + HD); }

CondSaved cond_saved_carve(float* base, int B) {
    CondSaved S;
    S.emb = base;
    S.h1 = S.emb + (size_t)B * ED;
    S.cv = S.h1 + (size_t)B * HD;
    S.m = S.cv + (size_t)B * TD;
    return S;
}

int cond_fwd_launch(const CondParams& P, const long long* time, float scale, const float* freqs, int B,
                    const CondSaved& S, float* cond, cudaStream_t stream) {
    (void)launch_pdl(cond_fwd_kernel, dim3(B), dim3(128), (size_t)(0), stream, P, time, scale, freqs, B, S, cond);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int cond_bwd_launch(const CondParams& P, const CondGrads& G, int B, const CondSaved& S, const float* dcond,
                    float* scratch, cudaStream_t stream) {
    float* dm = scratch;
    float* dcv = dm + (size_t)kNumBlocks * B * TD;
    float* dh1 = dcv + (size_t)B * TD;
    (void)launch_pdl(cond_bwd_chain_kernel, dim3(B), dim3(128), (size_t)(0), stream, P, B, S, dcond, dm, dcv, dh1);
    SINDDM_CUDA_OK(cudaGetLastError());

    OuterJobs jobs;
    int maxel = 0;
    auto add = [&](int idx, const float* L, const float* R, int rows, int cols, int gelu_r, float* dW, float* db) {
        jobs.j[idx] = OuterJob{L, R, rows, cols, gelu_r, dW, db};
        const int el = rows * (cols + 1);
        if (el > maxel) maxel = el;
    };
    add(0, dh1, S.emb, HD, ED, 0, G.w0, G.w0b);
    add(1, dcv, S.h1, TD, HD, 1, G.w2, G.w2b);
    size_t coff = 0;
    for (int l = 0; l < kNumBlocks; ++l) {
        add(2 + 2 * l, dm + (size_t)l * B * TD, S.cv, TD, TD, 1, G.wm[l], G.wmb[l]);
        add(3 + 2 * l, dcond + coff, S.m + (size_t)l * B * TD, P.C[l], TD, 0, G.wt[l], G.wtb[l]);
        coff += (size_t)B * P.C[l];
    }
    dim3 grid(ceil_div(maxel, 128), 2 + 2 * kNumBlocks);
    (void)launch_pdl(cond_bwd_weights_kernel, dim3(grid), dim3(128), (size_t)(0), stream, jobs, B);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
