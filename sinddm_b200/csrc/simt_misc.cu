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
    colsum_final_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, stream>>>(scratch, nblk, C, out);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int nchw_to_nhwc_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t stream) {
    const long long total = (long long)B * C * H * W;
    nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, dst, B, C, H * W);
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
