// HBM-bound NHWC helpers: column sums (bias gradients) and the 3-channel NCHW <-> NHWC layout changes.
#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

// ---------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[c] = sum_p a[p][c]
// ---------------------------------------------------------------------------------------------------
constexpr int kColsumBlocks = 296;

__global__ void colsum_partial_kernel(const float* __restrict__ a, long long P, int C, float* __restrict__ scratch) {
    extern __shared__ float red[];  // [lanes][C]
    const int c = threadIdx.x, ly = threadIdx.y, lanes = blockDim.y;
    const long long p_begin = P * blockIdx.x / gridDim.x;
    const long long p_end = P * (blockIdx.x + 1) / gridDim.x;
    float s0 = 0.f, s1 = 0.f;
    long long q = p_begin + ly;
    for (; q + lanes < p_end; q += 2 * lanes) {
        s0 += __ldg(a + q * C + c);
        s1 += __ldg(a + (q + lanes) * C + c);
    }
    if (q < p_end) s0 += __ldg(a + q * C + c);
    red[ly * C + c] = s0 + s1;
    __syncthreads();
    if (ly == 0) {
        float s = 0.f;
        for (int y = 0; y < lanes; ++y) s += red[y * C + c];
        scratch[(size_t)blockIdx.x * C + c] = s;
    }
}

// block = 32 channels x 8 lanes over the partial sums (fixed order: deterministic)
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ scratch, int nblk, int C,
                                                           float* __restrict__ out) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float red[8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (int k = threadIdx.y; k < nblk; k += 8) s += scratch[(size_t)k * C + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float v = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) v += red[y][threadIdx.x];
        out[c] = v;
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const long long total = (long long)B * C * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const long long b = p / HW, hw = p % HW;
        dst[i] = src[(b * C + c) * HW + hw];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
    const long long total = (long long)B * C * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long hw = i % HW;
        const int c = (int)((i / HW) % C);
        const long long b = i / ((long long)HW * C);
        dst[i] = src[(b * HW + hw) * C + c];
    }
}

// ---------------------------------------------------------------------------------------------------
// final_conv backward (reference SinDDM/models.py:130-132,151: 1x1 conv C -> 3) in ONE pass over the C-channel tensors:
//     d_o[p][c]  = sum_j W[j][c] * dout[p][j]            (gradient into the last block's output; tf32-rounded on request)
//     dW[j][c]   = sum_p dout[p][j] * o[p][c]            (per-block partial sums, fixed-order finish below)
//     db[j]      = sum_p dout[p][j]
// reads o (C channels) once and writes d_o once: 2 * P * C * 4 bytes, HBM bound.  Replaces five launches (3-channel
// column sum, CUDA-core weight gradient + reduce, CUDA-core data gradient) that read / wrote the same bytes 2.3 times,
// and the column sum of d_o (the last block's bias gradient), which is W^T db exactly -- no pass over memory at all.
// A thread owns 4 consecutive channels (one 16-byte access) of every (blockDim.x / (C/4))-th pixel of its block's range.
// ---------------------------------------------------------------------------------------------------
constexpr int kFinalBwdBlocks = 592;   // 4 per SM
constexpr int kFinalBwdThreads = 240;

// WRITE = false: only the weight-gradient partial sums (the same reduction serves the 1x1 residual conv FROM a 3-channel
// input, l1.res_conv: dW[c][j] = sum_p dy[p][c] * x3[p][j]).
template <bool WRITE>
__global__ void __launch_bounds__(kFinalBwdThreads)
final_conv_bwd_kernel(const float* __restrict__ dout3, const float* __restrict__ o, const float* __restrict__ w,
                      float* __restrict__ d_o, long long P, int C, int round, float* __restrict__ part) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    extern __shared__ float red[];   // [groups][16]
    const int c4n = C >> 2;                                   // threads per pixel
    const int groups = blockDim.x / c4n;
    const int g = threadIdx.x / c4n, c4 = threadIdx.x - g * c4n;
    const bool active = g < groups;
    const long long p_begin = P * blockIdx.x / gridDim.x, p_end = P * (blockIdx.x + 1) / gridDim.x;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, w2 = w0;
    if (active && WRITE) {
        // scalar loads: parameters are views into a flat buffer, aligned to 4 bytes only
        const float* wp = w + 4 * c4;
        w0 = make_float4(__ldg(wp), __ldg(wp + 1), __ldg(wp + 2), __ldg(wp + 3));
        w1 = make_float4(__ldg(wp + C), __ldg(wp + C + 1), __ldg(wp + C + 2), __ldg(wp + C + 3));
        w2 = make_float4(__ldg(wp + 2 * C), __ldg(wp + 2 * C + 1), __ldg(wp + 2 * C + 2), __ldg(wp + 2 * C + 3));
    }
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (active) {
#pragma unroll 2
        for (long long q = p_begin + g; q < p_end; q += groups) {
            const float d0 = __ldg(dout3 + q * 3 + 0), d1 = __ldg(dout3 + q * 3 + 1), d2 = __ldg(dout3 + q * 3 + 2);
            const float4 x = __ldg(reinterpret_cast<const float4*>(o + q * C) + c4);
            a0.x = fmaf(d0, x.x, a0.x); a0.y = fmaf(d0, x.y, a0.y); a0.z = fmaf(d0, x.z, a0.z); a0.w = fmaf(d0, x.w, a0.w);
            a1.x = fmaf(d1, x.x, a1.x); a1.y = fmaf(d1, x.y, a1.y); a1.z = fmaf(d1, x.z, a1.z); a1.w = fmaf(d1, x.w, a1.w);
            a2.x = fmaf(d2, x.x, a2.x); a2.y = fmaf(d2, x.y, a2.y); a2.z = fmaf(d2, x.z, a2.z); a2.w = fmaf(d2, x.w, a2.w);
            b0 += d0; b1 += d1; b2 += d2;
            if (!WRITE) continue;
            float4 y;
            y.x = fmaf(w2.x, d2, fmaf(w1.x, d1, w0.x * d0));
            y.y = fmaf(w2.y, d2, fmaf(w1.y, d1, w0.y * d0));
            y.z = fmaf(w2.z, d2, fmaf(w1.z, d1, w0.z * d0));
            y.w = fmaf(w2.w, d2, fmaf(w1.w, d1, w0.w * d0));
            if (round) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
            reinterpret_cast<float4*>(d_o + q * C)[c4] = y;
        }
    }
    // block partials in a fixed order: for each of the 12 sums of this thread's channel quad, add the pixel groups 0..G-1
    float* mine = red + (size_t)threadIdx.x * 16;
    mine[0] = a0.x; mine[1] = a0.y; mine[2] = a0.z; mine[3] = a0.w;
    mine[4] = a1.x; mine[5] = a1.y; mine[6] = a1.z; mine[7] = a1.w;
    mine[8] = a2.x; mine[9] = a2.y; mine[10] = a2.z; mine[11] = a2.w;
    mine[12] = b0; mine[13] = b1; mine[14] = b2; mine[15] = 0.f;
    __syncthreads();
    // part[block][3*C + 3]: dW partial rows j = 0..2 then the 3 db partials
    float* dst = part + (size_t)blockIdx.x * (3 * C + 4);
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
        const int j = i / C, c = i - j * C;
        float s = 0.f;
        for (int gg = 0; gg < groups; ++gg) s += red[(size_t)(gg * c4n + (c >> 2)) * 16 + j * 4 + (c & 3)];
        dst[i] = s;
    }
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int gg = 0; gg < groups; ++gg) s += red[(size_t)(gg * c4n) * 16 + 12 + threadIdx.x];
        dst[3 * C + threadIdx.x] = s;
    }
}

// dW[j][c], db[j] = fixed-order sums of the block partials; db_prev[c] = sum_j W[j][c] * db[j].  One warp per output
// value (lanes stride over the partial rows, fixed shuffle tree); every block recomputes the three db sums it needs.
// transpose != 0: dw is [C][3] (a conv FROM 3 channels) instead of [3][C]; db / db_prev may be null.
__global__ void __launch_bounds__(256)
final_conv_bwd_finish_kernel(const float* __restrict__ part, int nblk, int C, const float* __restrict__ w,
                             float* __restrict__ dw, float* __restrict__ db, float* __restrict__ db_prev,
                             float* __restrict__ db_prev2, int transpose) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float sdb[3];
    const int stride = 3 * C + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    auto column = [&](int i) {
        float s = 0.f;
        for (int k = lane; k < nblk; k += 32) s += part[(size_t)k * stride + i];
        return warp_sum(s);
    };
    if (warp < 3) {
        const float s = column(3 * C + warp);
        if (lane == 0) {
            sdb[warp] = s;
            if (blockIdx.x == 0 && db) db[warp] = s;
        }
    }
    for (int i = blockIdx.x * nwarps + warp; i < 3 * C; i += gridDim.x * nwarps) {
        const float s = column(i);
        if (lane == 0) dw[transpose ? (i % C) * 3 + i / C : i] = s;
    }
    __syncthreads();
    if (!db_prev) return;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const float v = fmaf(w[2 * C + c], sdb[2], fmaf(w[C + c], sdb[1], w[c] * sdb[0]));
        db_prev[c] = v;
        if (db_prev2) db_prev2[c] = v;
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148ll * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

size_t colsum_scratch_floats(int C) { return (size_t)kColsumBlocks * C; }

int colsum_launch(const float* a, long long P, int C, float* out, float* scratch, cudaStream_t stream) {
    SINDDM_REQUIRE(C <= 256, "colsum: C=%d too large", C);
    int lanes = 512 / C;
    if (lanes < 1) lanes = 1;
    if (lanes > 32) lanes = 32;
    int nblk = kColsumBlocks;
    if ((long long)nblk * lanes > P) nblk = (int)((P + lanes - 1) / lanes);
    if (nblk < 1) nblk = 1;
    dim3 block(C, lanes);
    colsum_partial_kernel<<<nblk, block, (size_t)lanes * C * sizeof(float), stream>>>(a, P, C, scratch);
    SINDDM_CUDA_OK(cudaGetLastError());
    (void)launch_pdl(colsum_final_kernel, dim3(ceil_div(C, 32)), dim3(dim3(32, 8)), (size_t)(0), stream, scratch, nblk, C, out);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

size_t final_conv_bwd_scratch_floats(int C) { return (size_t)kFinalBwdBlocks * (3 * C + 4); }

bool final_conv_bwd_supported(int C) { return C % 4 == 0 && C >= 4 && C / 4 <= kFinalBwdThreads; }

int final_conv_bwd_launch(const float* dout_nhwc3, const float* o, const float* w, float* d_o, long long P, int C,
                          int round, float* dw, float* db, float* db_prev, float* db_prev2, float* scratch,
                          cudaStream_t stream) {
    SINDDM_REQUIRE(final_conv_bwd_supported(C), "final_conv_bwd: C=%d unsupported", C);
    const int c4n = C / 4;
    const int threads = (kFinalBwdThreads / c4n) * c4n;
    int nblk = kFinalBwdBlocks;
    if ((long long)nblk > P) nblk = (int)P;
    (void)launch_pdl(final_conv_bwd_kernel<true>, dim3(nblk), dim3(threads), (size_t)((size_t)threads * 16 * sizeof(float)), stream, dout_nhwc3, o, w, d_o, P,
                                                                                               C, round, scratch);
    SINDDM_CUDA_OK(cudaGetLastError());
    (void)launch_pdl(final_conv_bwd_finish_kernel, dim3(16), dim3(256), (size_t)(0), stream, scratch, nblk, C, w, dw, db, db_prev, db_prev2, 0);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int wgrad_from_c3_launch(const float* x3, const float* dy, long long P, int C, float* dw, float* scratch,
                         cudaStream_t stream) {
    SINDDM_REQUIRE(final_conv_bwd_supported(C), "wgrad_from_c3: C=%d unsupported", C);
    const int c4n = C / 4;
    const int threads = (kFinalBwdThreads / c4n) * c4n;
    int nblk = kFinalBwdBlocks;
    if ((long long)nblk > P) nblk = (int)P;
    (void)launch_pdl(final_conv_bwd_kernel<false>, dim3(nblk), dim3(threads), (size_t)((size_t)threads * 16 * sizeof(float)), stream, x3, dy, nullptr, nullptr, P,
                                                                                                C, 0, scratch);
    SINDDM_CUDA_OK(cudaGetLastError());
    (void)launch_pdl(final_conv_bwd_finish_kernel, dim3(16), dim3(256), (size_t)(0), stream, scratch, nblk, C, nullptr, dw, nullptr, nullptr, nullptr, 1);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int colsum_final_launch(const float* part, int nrows, int C, float* out, cudaStream_t stream) {
    (void)launch_pdl(colsum_final_kernel, dim3(ceil_div(C, 32)), dim3(dim3(32, 8)), (size_t)(0), stream, part, nrows, C, out);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int nchw_to_nhwc_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t stream) {
    const long long total = (long long)B * C * H * W;
    (void)launch_pdl(nchw_to_nhwc_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, src, dst, B, C, H * W);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int nhwc_to_nchw_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t stream) {
    const long long total = (long long)B * C * H * W;
    nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, dst, B, C, H * W);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
