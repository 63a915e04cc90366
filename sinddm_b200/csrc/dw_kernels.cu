// Depthwise 5x5 (+bias +conditioning) and its gradients on NHWC fp32 -- HBM-bound kernels.
//
// dw5x5 replaces `h = self.ds_conv(x); h = h + condition` (reference SinDDM/models.py:61,70,77); with
// flip=1 it is the data gradient; dw5x5_wgrad replaces what autograd derives for ds_conv.weight / bias and
// for the conditioning vector.
//
// Layout of the work: one thread owns ONE channel of one row segment (b, h, 64 px) and slides a 5x5 register
// window along W.  Adjacent threads are adjacent channels, so every global access of a warp is one
// contiguous 128-byte line; per output pixel a thread loads 5 new inputs (the window's new column) instead
// of 25, and the 25 weights live in registers.
#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kSeg = 64;          // pixels per row segment
constexpr int kRowsPerBlock = 8;  // image rows per CTA (threadIdx.y)

constexpr int kRing = 8;  // window ring: 5 live columns + 3 columns of load-ahead (hides L2 latency)

struct RowWindow {
    // v[ky][slot]: slot = (image column - first column) modulo kRing
    float v[5][kRing];
};

// loads column `col` of the 5 input rows around h into slot (col mod 5); zero outside the image
SINDDM_DEVINL void load_column(RowWindow& win, const float* const (&rowp)[5], const bool (&rowok)[5], int col, int W,
                               int C, int slot) {
    const bool ok = col >= 0 && col < W;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) win.v[ky][slot] = (ok && rowok[ky]) ? __ldg(rowp[ky] + (size_t)col * C) : 0.f;
}

template <typename Body>
SINDDM_DEVINL void slide_row(const float* __restrict__ in, int b, int h, int w0, int w1, int c, int H, int W, int C,
                             Body&& body) {
    const float* rowp[5];
    bool rowok[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        const int hh = h + ky - 2;
        rowok[ky] = hh >= 0 && hh < H;
        rowp[ky] = in + (((size_t)b * H + (rowok[ky] ? hh : 0)) * W) * C + c;
    }
    RowWindow win;
    // output column wo = w0 + i reads image columns base+i .. base+i+4 (base = w0 - 2) = slots (i + kx) % kRing;
    // columns base .. base+6 are loaded up front, iteration i loads column base+i+7 (needed 3 outputs later)
    const int base = w0 - 2;
#pragma unroll
    for (int j = 0; j < kRing - 1; ++j) load_column(win, rowp, rowok, base + j, W, C, j);
    // kRing outputs per outer iteration so ring slots are compile-time constants
    for (int w = w0; w < w1; w += kRing) {
#pragma unroll
        for (int j = 0; j < kRing; ++j) {
            const int wo = w + j;
            if (wo < w1) {
                load_column(win, rowp, rowok, wo + 5, W, C, (j + kRing - 1) % kRing);
                float x[5][5];
#pragma unroll
                for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) x[ky][kx] = win.v[ky][(j + kx) % kRing];
                body(wo, x);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
dw5x5_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
             const float* __restrict__ cond, const float* __restrict__ add, float* __restrict__ out, int B, int H,
             int W, int C, int flip, int round) {
    // block = 32 channels x 8 image rows of one 64-px segment: the 8 row-warps read overlapping input rows
    // at the same time, so 12 input rows are fetched from L2 per 8 output rows (L1 serves the rest)
    const int nseg = (W + kSeg - 1) / kSeg;
    const int seg = blockIdx.x % nseg;
    const int c = (blockIdx.x / nseg) * 32 + threadIdx.x;
    const int h = blockIdx.y * kRowsPerBlock + threadIdx.y;
    const int b = blockIdx.z;
    if (c >= C || h >= H) return;

    float wr[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) wr[t / 5][t % 5] = __ldg(wgt + c * 25 + (flip ? 24 - t : t));
    const float bv = bias ? __ldg(bias + c) : 0.f;
    const float cv = cond ? __ldg(cond + (size_t)b * C + c) : 0.f;
    const int w0 = seg * kSeg;
    const int w1 = min(W, w0 + kSeg);
    const size_t rowoff = (((size_t)b * H + h) * W) * C + c;

    slide_row(in, b, h, w0, w1, c, H, W, C, [&](int wo, const float (&x)[5][5]) {
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) acc = fmaf(x[ky][kx], wr[ky][kx], acc);
        const size_t off = rowoff + (size_t)wo * C;
        float v = (acc + bv) + cv;               // reference order: (conv + bias) + condition
        if (add) v += __ldg(add + off);
        if (round) v = round_tf32(v);
        out[off] = v;
    });
}

// ---------------------------------------------------------------------------------------------------
// gradients of the depthwise weight / bias / conditioning vector
//   grid = B * nchunk CTAs (chunk = kRowsPerChunk image rows), block = (C, lanes)
//   thread (c, lane) walks the chunk's (row, segment) units lane, lane+lanes, ... with the sliding window,
//   accumulates 25 tap sums + 1 plain sum, smem-reduces over lanes -> scratch[b][chunk][26][C]
// ---------------------------------------------------------------------------------------------------
constexpr int kRowsPerChunk = 8;

__global__ void __launch_bounds__(256)
dw5x5_wgrad_partial_kernel(const float* __restrict__ x, const float* __restrict__ dh, float* __restrict__ scratch,
                           int B, int H, int W, int C, int nchunk) {
    __shared__ float red[kRowsPerChunk][26][32];
    const int lane = threadIdx.x;           // channel within the 32-channel group
    const int ry = threadIdx.y;             // image row within the chunk
    const int c = blockIdx.x * 32 + lane;
    const int chunk = blockIdx.y;
    const int b = blockIdx.z;
    const int h = chunk * kRowsPerChunk + ry;

    float acc[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t / 5][t % 5] = 0.f;
    float gsum = 0.f;

    if (c < C && h < H) {
        const float* dhrow = dh + (((size_t)b * H + h) * W) * C + c;
        slide_row(x, b, h, 0, W, c, H, W, C, [&](int wo, const float (&xv)[5][5]) {
            const float g = __ldg(dhrow + (size_t)wo * C);
            gsum += g;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) acc[ky][kx] = fmaf(xv[ky][kx], g, acc[ky][kx]);
        });
    }
#pragma unroll
    for (int t = 0; t < 25; ++t) red[ry][t][lane] = acc[t / 5][t % 5];
    red[ry][25][lane] = gsum;
    __syncthreads();
    // 256 threads reduce 26 x 32 sums over the 8 rows
    for (int i = threadIdx.y * 32 + lane; i < 26 * 32; i += 256) {
        const int t = i / 32, l = i % 32;
        const int cc = blockIdx.x * 32 + l;
        if (cc < C) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kRowsPerChunk; ++y) s += red[y][t][l];
            scratch[(((size_t)b * nchunk + chunk) * 26 + t) * C + cc] = s;
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
// TMA-staged versions (C % 4 == 0): the (8+4) x (32+4) pixel x 32 channel halo tile is brought into shared
// memory by ONE bulk-tensor copy whose out-of-image part TMA zero-fills (= the conv's zero padding, no
// bounds code at all); thread (channel, row) then slides its 5x5 window along the 32 columns reading
// 5 shared-memory words per output (a warp = 32 channels of one pixel = one conflict-free 128-byte row).
// ---------------------------------------------------------------------------------------------------
constexpr int kTW = 32;                                  // output columns per tile
constexpr int kTH = 8;                                   // output rows per tile (= threadIdx.y)
constexpr int kTileCols = kTW + 4, kTileRows = kTH + 4;
constexpr int kTileFloats = kTileRows * kTileCols * 32;  // 13824 floats = 55296 B

template <typename Body>
SINDDM_DEVINL void slide_smem(const float* tile, int ry, int lane, Body&& body) {
    // win[ky][slot], slot = tile column modulo 5
    float win[5][5];
    const float* base = tile + (size_t)ry * kTileCols * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) win[ky][j] = base[(ky * kTileCols + j) * 32];
#pragma unroll   // fully unrolled: `wo` is a compile-time constant for the caller (register arrays indexed by it)
    for (int w = 0; w < kTW; w += 5) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int wo = w + j;
            if (wo < kTW) {
#pragma unroll
                for (int ky = 0; ky < 5; ++ky) win[ky][(j + 4) % 5] = base[(ky * kTileCols + wo + 4) * 32];
                float x[5][5];
#pragma unroll
                for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) x[ky][kx] = win[ky][(j + kx) % 5];
                body(wo, x);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
dw5x5_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ wgt,
                 const float* __restrict__ bias, const float* __restrict__ cond, const float* __restrict__ add,
                 float* __restrict__ out, int H, int W, int C, int tiles_w, int flip, int round) {
    extern __shared__ __align__(128) uint8_t dsm[];
    float* tile = reinterpret_cast<float*>(dsm);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm + kTileFloats * sizeof(float));
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int tw = blockIdx.x % tiles_w;
    const int cg = blockIdx.x / tiles_w;
    const int th = blockIdx.y, b = blockIdx.z;
    const int tid = ry * 32 + lane;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, kTileFloats * sizeof(float));
        tma_load_4d(tile, &tm_in, bar, cg * 32, tw * kTW - 2, th * kTH - 2, b);
    }
    const int c = cg * 32 + lane;
    const int h = th * kTH + ry;
    const bool cok = c < C;
    float wr[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) wr[t / 5][t % 5] = cok ? __ldg(wgt + c * 25 + (flip ? 24 - t : t)) : 0.f;
    const float bv = (bias && cok) ? __ldg(bias + c) : 0.f;
    const float cv = (cond && cok) ? __ldg(cond + (size_t)b * C + c) : 0.f;
    mbar_wait(bar, 0);
    if (!cok || h >= H) return;
    const size_t rowoff = (((size_t)b * H + h) * W) * C + c;
    const int w0 = tw * kTW;
    slide_smem(tile, ry, lane, [&](int wo, const float (&x)[5][5]) {
        const int w = w0 + wo;
        if (w < W) {
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) acc = fmaf(x[ky][kx], wr[ky][kx], acc);
            const size_t off = rowoff + (size_t)w * C;
            float v = (acc + bv) + cv;           // reference order: (conv + bias) + condition
            if (add) v += __ldg(add + off);
            if (round) v = round_tf32(v);
            out[off] = v;
        }
    });
}

// (A two-channels-per-thread variant with 64-channel boxes and packed fp32x2 FMAs was measured slower --
// 1.03 ms vs 0.76 ms per 160-channel launch at 32x186x248 -- because its 125 registers halve the resident
// warps; the one-channel kernel above stays.)

// grid = (channel groups, row chunks, B); the CTA walks the column tiles of its 8-row chunk with a
// double-buffered TMA pipeline and keeps the 26 sums per (channel, row) in registers.
__global__ void __launch_bounds__(256)
dw5x5_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ dh,
                       float* __restrict__ scratch, int H, int W, int C, int tiles_w, int nchunk) {
    extern __shared__ __align__(128) uint8_t dsm[];
    float* tiles = reinterpret_cast<float*>(dsm);   // 2 buffers; reused for the final reduction
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm + 2 * kTileFloats * sizeof(float));
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int cg = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int tid = ry * 32 + lane;
    const int c = cg * 32 + lane;
    const int h = chunk * kTH + ry;
    const bool active = c < C && h < H;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar[0], kTileFloats * sizeof(float));
        tma_load_4d(tiles, &tm_x, &bar[0], cg * 32, -2, chunk * kTH - 2, b);
    }
    float acc[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t / 5][t % 5] = 0.f;
    float gsum = 0.f;
    const float* dhrow = dh + (((size_t)b * H + (h < H ? h : 0)) * W) * C + (c < C ? c : 0);

    for (int tw = 0; tw < tiles_w; ++tw) {
        const int buf = tw & 1;
        if (tid == 0 && tw + 1 < tiles_w) {
            // the other buffer was consumed in iteration tw-1 (all threads passed the __syncthreads below)
            mbar_arrive_expect_tx(&bar[buf ^ 1], kTileFloats * sizeof(float));
            tma_load_4d(tiles + (size_t)(buf ^ 1) * kTileFloats, &tm_x, &bar[buf ^ 1], cg * 32, (tw + 1) * kTW - 2,
                        chunk * kTH - 2, b);
        }
        mbar_wait(&bar[buf], (uint32_t)(tw >> 1) & 1u);
        if (active) {
            const int w0 = tw * kTW;
            // the 32 upstream-gradient values of this row segment: issued together (32 loads in flight) instead of
            // one dependent load per pixel inside the sliding loop
            float gv[kTW];
#pragma unroll
            for (int i = 0; i < kTW; ++i) gv[i] = (w0 + i < W) ? __ldg(dhrow + (size_t)(w0 + i) * C) : 0.f;
            slide_smem(tiles + (size_t)buf * kTileFloats, ry, lane, [&](int wo, const float (&xv)[5][5]) {
                const int w = w0 + wo;
                if (w < W) {
                    const float g = gv[wo];
                    gsum += g;
#pragma unroll
                    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) acc[ky][kx] = fmaf(xv[ky][kx], g, acc[ky][kx]);
                }
            });
        }
        __syncthreads();   // everyone is done with `buf` before it is refilled two iterations later
    }
    // reduce over the 8 rows through shared memory (the tile buffers are free now)
    float* red = tiles;    // [8][26][32]
#pragma unroll
    for (int t = 0; t < 25; ++t) red[(ry * 26 + t) * 32 + lane] = active ? acc[t / 5][t % 5] : 0.f;
    red[(ry * 26 + 25) * 32 + lane] = active ? gsum : 0.f;
    __syncthreads();
    for (int i = tid; i < 26 * 32; i += 256) {
        const int t = i / 32, l = i % 32;
        const int cc = cg * 32 + l;
        if (cc < C) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kTH; ++y) s += red[(y * 26 + t) * 32 + l];
            scratch[(((size_t)b * nchunk + chunk) * 26 + t) * C + cc] = s;
        }
    }
}

}  // namespace

int dw5x5_launch(const float* in, const float* w, const float* bias, const float* cond, const float* add, float* out,
                 int B, int H, int W, int C, int flip, int round_tf32, cudaStream_t stream) {
    SINDDM_REQUIRE(B <= 65535, "dw5x5: batch too large");
    if (C % 4 == 0 && device_info().initialized) {
        CUtensorMap tm;
        SINDDM_TRY(make_tmap_nhwc(&tm, in, B, H, W, C, 32, kTileCols, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE));
        const int tiles_w = ceil_div(W, kTW);
        const size_t smem = kTileFloats * sizeof(float) + 16;
        static int attr_set = 0;
        if (!attr_set) {
            SINDDM_CUDA_OK(cudaFuncSetAttribute(dw5x5_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            attr_set = 1;
        }
        dim3 grid(tiles_w * ceil_div(C, 32), ceil_div(H, kTH), B);
        dim3 block(32, kTH);
        dw5x5_tma_kernel<<<grid, block, smem, stream>>>(tm, w, bias, cond, add, out, H, W, C, tiles_w, flip,
                                                        round_tf32);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    const int nseg = ceil_div(W, kSeg);
    dim3 grid(nseg * ceil_div(C, 32), ceil_div(H, kRowsPerBlock), B);
    dim3 block(32, kRowsPerBlock);
    dw5x5_kernel<<<grid, block, 0, stream>>>(in, w, bias, cond, add, out, B, H, W, C, flip, round_tf32);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

size_t dw5x5_wgrad_scratch_floats(int B, int H, int C) {
    return (size_t)B * ceil_div(H, kRowsPerChunk) * 26 * C;
}

int dw5x5_wgrad_launch(const float* x, const float* dh, float* dw, float* db, float* dcond, float* scratch, int B,
                       int H, int W, int C, cudaStream_t stream) {
    SINDDM_REQUIRE(B <= 65535, "dw5x5_wgrad: batch too large");
    static_assert(kRowsPerChunk == kTH, "scratch layout assumes 8-row chunks in both kernels");
    const int nchunk = ceil_div(H, kRowsPerChunk);
    dim3 grid(ceil_div(C, 32), nchunk, B);
    dim3 block(32, kRowsPerChunk);
    if (C % 4 == 0 && device_info().initialized) {
        CUtensorMap tm;
        SINDDM_TRY(make_tmap_nhwc(&tm, x, B, H, W, C, 32, kTileCols, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE));
        const size_t smem = 2 * kTileFloats * sizeof(float) + 16;
        static int attr_set = 0;
        if (!attr_set) {
            SINDDM_CUDA_OK(cudaFuncSetAttribute(dw5x5_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            attr_set = 1;
        }
        dw5x5_wgrad_tma_kernel<<<grid, block, smem, stream>>>(tm, dh, scratch, H, W, C, ceil_div(W, kTW), nchunk);
    } else {
        dw5x5_wgrad_partial_kernel<<<grid, block, 0, stream>>>(x, dh, scratch, B, H, W, C, nchunk);
    }
    SINDDM_CUDA_OK(cudaGetLastError());
    dim3 grid2(ceil_div(C, 64), 26);
    dw5x5_wgrad_final_kernel<<<grid2, 64, 0, stream>>>(scratch, dw, db, dcond, B, C, nchunk);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
