// HBM-bound NHWC kernels: depthwise 5x5 (+bias +conditioning), its gradients, column sums, layout changes.
//
// dw5x5 replaces `h = self.ds_conv(x); h = h + condition` (reference SinDDM/models.py:61,70,77); the
// gradient kernels replace what autograd derives for them.
#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kDwPix = 4;  // consecutive output pixels (along W) per thread; each thread owns 4 channels

// out[p][c] = add[p][c] + bias[c] + cond[b][c] + sum_{ky,kx} w[c][ky*5+kx (flipped if flip)] * in[(h+ky-2, w+kx-2)][c]
template <int VEC>
__global__ void __launch_bounds__(256)
dw5x5_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
             const float* __restrict__ cond, const float* __restrict__ add, float* __restrict__ out, int B, int H,
             int W, int C, int flip, int round) {
    extern __shared__ float wsm[];  // [25][C] tap-major so a thread reads VEC consecutive channels
    for (int i = threadIdx.x; i < 25 * C; i += blockDim.x) {
        const int c = i % C, tap = i / C;
        wsm[i] = wgt[c * 25 + (flip ? 24 - tap : tap)];
    }
    __syncthreads();

    const int CV = C / VEC;
    const int WG = (W + kDwPix - 1) / kDwPix;
    const long long total = (long long)B * H * WG * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        const int wg = (int)((idx / CV) % WG);
        const int h = (int)((idx / ((long long)CV * WG)) % H);
        const int b = (int)(idx / ((long long)CV * WG * H));
        const int c = cv * VEC;
        const int w0 = wg * kDwPix;

        float acc[kDwPix][VEC];
#pragma unroll
        for (int px = 0; px < kDwPix; ++px)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[px][v] = 0.f;

#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int hh = h + ky - 2;
            if (hh < 0 || hh >= H) continue;
            const float* rowp = in + (((size_t)b * H + hh) * W) * C + c;
            float xin[kDwPix + 4][VEC];
#pragma unroll
            for (int i = 0; i < kDwPix + 4; ++i) {
                const int ww = w0 + i - 2;
                if (ww >= 0 && ww < W) {
                    if (VEC == 4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(rowp + (size_t)ww * C));
                        xin[i][0] = t.x;
                        xin[i][1 % VEC] = t.y;
                        xin[i][2 % VEC] = t.z;
                        xin[i][3 % VEC] = t.w;
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) xin[i][v] = __ldg(rowp + (size_t)ww * C + v);
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xin[i][v] = 0.f;
                }
            }
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                float wv[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) wv[v] = wsm[(ky * 5 + kx) * C + c + v];
#pragma unroll
                for (int px = 0; px < kDwPix; ++px)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[px][v] = fmaf(xin[px + kx][v], wv[v], acc[px][v]);
            }
        }

        float bv[VEC], cdv[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            bv[v] = bias ? bias[c + v] : 0.f;
            cdv[v] = cond ? cond[(size_t)b * C + c + v] : 0.f;
        }
#pragma unroll
        for (int px = 0; px < kDwPix; ++px) {
            const int w = w0 + px;
            if (w >= W) break;
            const size_t off = (((size_t)b * H + h) * W + w) * C + c;
            float r[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // reference order: (conv + bias) + condition
                r[v] = (acc[px][v] + bv[v]) + cdv[v];
                if (add) r[v] += add[off + v];
                if (round) r[v] = round_tf32(r[v]);
            }
            if (VEC == 4) {
                *reinterpret_cast<float4*>(out + off) = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) out[off + v] = r[v];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// depthwise weight / bias / conditioning gradients.
// grid = (B * nchunk) CTAs, block = (C, kLanes): thread (c, ly) accumulates 26 sums over the pixels of rows
// h = chunk rows, pixel lanes ly; smem-reduced over ly -> scratch[b][chunk][26][C].
// ---------------------------------------------------------------------------------------------------
constexpr int kRowsPerChunk = 8;

__global__ void dw5x5_wgrad_partial_kernel(const float* __restrict__ x, const float* __restrict__ dh,
                                           float* __restrict__ scratch, int B, int H, int W, int C, int nchunk) {
    extern __shared__ float red[];  // [lanes][26][C]
    const int c = threadIdx.x;
    const int ly = threadIdx.y;
    const int lanes = blockDim.y;
    const int chunk = blockIdx.x % nchunk;
    const int b = blockIdx.x / nchunk;
    const int h_begin = chunk * kRowsPerChunk;
    const int h_end = min(H, h_begin + kRowsPerChunk);

    float acc[26];
#pragma unroll
    for (int i = 0; i < 26; ++i) acc[i] = 0.f;

    const int npix = (h_end - h_begin) * W;
    for (int q = ly; q < npix; q += lanes) {
        const int h = h_begin + q / W;
        const int w = q % W;
        const float g = __ldg(dh + (((size_t)b * H + h) * W + w) * C + c);
        acc[25] += g;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int hh = h + ky - 2;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int ww = w + kx - 2;
                if (ww < 0 || ww >= W) continue;
                acc[ky * 5 + kx] = fmaf(__ldg(x + (((size_t)b * H + hh) * W + ww) * C + c), g, acc[ky * 5 + kx]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 26; ++i) red[(ly * 26 + i) * C + c] = acc[i];
    __syncthreads();
    if (ly == 0) {
        for (int i = 0; i < 26; ++i) {
            float s = 0.f;
            for (int y = 0; y < lanes; ++y) s += red[(y * 26 + i) * C + c];
            scratch[(((size_t)b * nchunk + chunk) * 26 + i) * C + c] = s;
        }
    }
}

// stage 2: dcond[b][c] = sum_chunk s[b][chunk][25][c]; dw[c][tap] = sum_b sum_chunk s[..][tap][c]; db = sum_b dcond
__global__ void dw5x5_wgrad_final_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                         float* __restrict__ db, float* __restrict__ dcond, int B, int C, int nchunk) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int i = blockIdx.y;  // 0..25
    float tot = 0.f;
    for (int b = 0; b < B; ++b) {
        float s = 0.f;
        for (int k = 0; k < nchunk; ++k) s += scratch[(((size_t)b * nchunk + k) * 26 + i) * C + c];
        if (i == 25 && dcond) dcond[(size_t)b * C + c] = s;
        tot += s;
    }
    if (i == 25) {
        if (db) db[c] = tot;
    } else if (dw) {
        dw[c * 25 + i] = tot;
    }
}

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

__global__ void colsum_final_kernel(const float* __restrict__ scratch, int nblk, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += scratch[(size_t)k * C + c];
    out[c] = s;
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

int dw5x5_launch(const float* in, const float* w, const float* bias, const float* cond, const float* add, float* out,
                 int B, int H, int W, int C, int flip, int round_tf32, cudaStream_t stream) {
    const size_t smem = (size_t)25 * C * sizeof(float);
    const int WG = ceil_div(W, kDwPix);
    if (C % 4 == 0) {
        const long long total = (long long)B * H * WG * (C / 4);
        dw5x5_kernel<4><<<grid_for(total, 256), 256, smem, stream>>>(in, w, bias, cond, add, out, B, H, W, C, flip,
                                                                    round_tf32);
    } else {
        const long long total = (long long)B * H * WG * C;
        dw5x5_kernel<1><<<grid_for(total, 256), 256, smem, stream>>>(in, w, bias, cond, add, out, B, H, W, C, flip,
                                                                    round_tf32);
    }
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

size_t dw5x5_wgrad_scratch_floats(int B, int H, int C) {
    return (size_t)B * ceil_div(H, kRowsPerChunk) * 26 * C;
}

int dw5x5_wgrad_launch(const float* x, const float* dh, float* dw, float* db, float* dcond, float* scratch, int B,
                       int H, int W, int C, cudaStream_t stream) {
    SINDDM_REQUIRE(C <= 256, "dw5x5_wgrad: C=%d too large", C);
    const int nchunk = ceil_div(H, kRowsPerChunk);
    int lanes = 512 / C;
    if (lanes < 1) lanes = 1;
    if (lanes > 16) lanes = 16;
    dim3 block(C, lanes);
    const size_t smem = (size_t)lanes * 26 * C * sizeof(float);
    static int attr_set = 0;
    if (smem > 48 * 1024 && !attr_set) {
        SINDDM_CUDA_OK(cudaFuncSetAttribute(dw5x5_wgrad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            96 * 1024));
        attr_set = 1;
    }
    dw5x5_wgrad_partial_kernel<<<B * nchunk, block, smem, stream>>>(x, dh, scratch, B, H, W, C, nchunk);
    SINDDM_CUDA_OK(cudaGetLastError());
    dim3 grid2(ceil_div(C, 64), 26);
    dw5x5_wgrad_final_kernel<<<grid2, 64, 0, stream>>>(scratch, dw, db, dcond, B, C, nchunk);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

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
    colsum_final_kernel<<<ceil_div(C, 64), 64, 0, stream>>>(scratch, nblk, C, out);
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
