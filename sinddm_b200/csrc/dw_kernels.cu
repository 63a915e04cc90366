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
// block = 32 channels x 8 batch lanes: the per-image sums are formed in parallel (each in chunk order), then one thread
// per channel adds them in batch order -- the same association as a single sequential loop, at an eighth of its latency
// (this kernel runs four times per training step and used to cost 40 us at EVERY scale).
constexpr int kFinalBatchLanes = 8;
__global__ void __launch_bounds__(32 * kFinalBatchLanes)
dw5x5_wgrad_final_kernel(const float* __restrict__ scratch, float* __restrict__ dw, float* __restrict__ db,
                         float* __restrict__ dcond, int B, int C, int nchunk) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    extern __shared__ float per_image[];   // [B][32]
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int i = blockIdx.y;  // 0..25
    const bool cok = c < C;
    for (int b = threadIdx.y; b < B; b += kFinalBatchLanes) {
        float s = 0.f;
        if (cok) {
            const float* src = scratch + (((size_t)b * nchunk) * 26 + i) * C + c;
            for (int k = 0; k < nchunk; ++k) s += src[(size_t)k * 26 * C];
            if (i == 25 && dcond) dcond[(size_t)b * C + c] = s;
        }
        per_image[b * 32 + threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.y == 0 && cok) {
        float tot = 0.f;
        for (int b = 0; b < B; ++b) tot += per_image[b * 32 + threadIdx.x];
        if (i == 25) {
            if (db) db[c] = tot;
        } else if (dw) {
            dw[c * 25 + i] = tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// C == 3 (l1.ds_conv on the network input): the channel-per-lane kernels above would leave 29 of 32 lanes idle.
// Here a thread owns a PIXEL and all three channels: the 75 weights sit in registers, neighbouring threads read
// neighbouring 12-byte pixels (every load instruction of a warp covers one contiguous 384-byte span, and the 25
// taps of a pixel hit lines its neighbours already pulled into L1).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dw5x5_c3_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                const float* __restrict__ cond, const float* __restrict__ add, float* __restrict__ out, int B, int H,
                int W, int flip, int round) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    float wr[25][3];
#pragma unroll
    for (int t = 0; t < 25; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) wr[t][c] = __ldg(wgt + c * 25 + (flip ? 24 - t : t));
    const long long P = (long long)B * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(p % W);
        const int h = (int)((p / W) % H);
        const int b = (int)(p / ((long long)W * H));
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int hh = h + ky - 2;
            if (hh < 0 || hh >= H) continue;
            const float* row = in + (p + (long long)(ky - 2) * W) * 3;
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int ww = w + kx - 2;
                if (ww < 0 || ww >= W) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] = fmaf(__ldg(row + (kx - 2) * 3 + c), wr[ky * 5 + kx][c], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = (acc[c] + (bias ? __ldg(bias + c) : 0.f)) + (cond ? __ldg(cond + b * 3 + c) : 0.f);
            if (add) v += __ldg(add + p * 3 + c);
            out[p * 3 + c] = round ? round_tf32(v) : v;
        }
    }
}

// gradients for C == 3: grid = (chunks, B); a thread accumulates the 75 tap sums + 3 plain sums over its pixels of
// image b, the block combines them with shuffles + shared memory -> scratch[b][chunk][26][3]
__global__ void __launch_bounds__(256)
dw5x5_wgrad_c3_kernel(const float* __restrict__ x, const float* __restrict__ dh, float* __restrict__ scratch, int H,
                      int W, int nchunk) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float red[8][78];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int HW = H * W;
    const float* xb = x + (size_t)b * HW * 3;
    const float* db = dh + (size_t)b * HW * 3;
    float acc[26][3];
#pragma unroll
    for (int t = 0; t < 26; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[t][c] = 0.f;
    for (int p = chunk * blockDim.x + threadIdx.x; p < HW; p += nchunk * blockDim.x) {
        const int w = p % W, h = p / W;
        float g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g[c] = __ldg(db + (size_t)p * 3 + c);
            acc[25][c] += g[c];
        }
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int hh = h + ky - 2;
            if (hh < 0 || hh >= H) continue;
            const float* row = xb + (size_t)(p + (ky - 2) * W) * 3;
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int ww = w + kx - 2;
                if (ww < 0 || ww >= W) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[ky * 5 + kx][c] = fmaf(__ldg(row + (kx - 2) * 3 + c), g[c], acc[ky * 5 + kx][c]);
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < 26; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = warp_sum(acc[t][c]);
            if (lane == 0) red[warp][t * 3 + c] = s;
        }
    __syncthreads();
    if (threadIdx.x < 78) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x];
        scratch[((size_t)b * nchunk + chunk) * 78 + threadIdx.x] = s;   // [26][3]: t * 3 + c
    }
}

// ---------------------------------------------------------------------------------------------------
// TMA-staged versions (C % 4 == 0).  These kernels are bound by instruction issue long before HBM (25 FMAs per
// output plus whatever surrounds them), so the layout of the work is chosen to minimise instructions per output:
//   * tile = 16 x 16 output pixels x 32 channels; its (16+4) x (16+4) halo box is ONE bulk-tensor copy whose
//     out-of-image part TMA zero-fills (= the conv's zero padding: no bounds code on the input side at all);
//   * thread (channel lane, row pair ry) produces output rows 2ry and 2ry+1 together and slides a 6 x 5 register
//     window along the 16 columns: 6 conflict-free shared-memory words (a warp = 32 channels of one pixel = one
//     128-byte row) feed 50 FMAs in two independent chains -- 3 loads per output instead of 5;
//   * a CTA walks up to kMaxTilesPerCta tiles of one tile row with a two-buffer TMA pipeline (the box of tile
//     i+1 is in flight while tile i is computed), so the 25 weights / bias / condition are loaded once per CTA
//     and two resident CTAs keep ~100 KB of loads in flight per SM;
//   * interior tiles run a store path without any bounds predicate; the epilogue flavour (residual add, tf32
//     rounding) is a template parameter.
// ---------------------------------------------------------------------------------------------------
constexpr int kTW = 16;                                  // output columns per tile
constexpr int kTH = 16;                                  // output rows per tile (two per thread row)
constexpr int kTileCols = kTW + 4, kTileRows = kTH + 4;
constexpr int kTileFloats = kTileRows * kTileCols * 32;  // 12800 floats = 51200 B
constexpr int kRowThreads = kTH / 2;                     // threadIdx.y extent
constexpr int kMaxTilesPerCta = 8;

// loads window column `col` (tile coordinates) of the six rows 2ry .. 2ry+5 into slot col % 5
SINDDM_DEVINL void load_window_column(float (&win)[6][5], const float* base, int col) {
#pragma unroll
    for (int k = 0; k < 6; ++k) win[k][col % 5] = base[(k * kTileCols + col) * 32];
}

template <bool ADD, bool ROUND, bool FULL, bool CSUM>
SINDDM_DEVINL void dw5x5_tile(const float* tile, int ry, int lane, const float (&wr)[5][5], float bv, float cv,
                              const float* __restrict__ addp, float* __restrict__ outp, size_t off, int C,
                              size_t rowstride, int wvalid, bool ok_a, bool ok_b, float& csum) {
    // residual operand of the whole tile: 32 independent loads in flight before the first FMA needs one
    float ad_a[kTW], ad_b[kTW];
    if (ADD) {
#pragma unroll
        for (int wo = 0; wo < kTW; ++wo) {
            const size_t o = off + (size_t)wo * C;
            ad_a[wo] = (FULL || (ok_a && wo < wvalid)) ? __ldg(addp + o) : 0.f;
            ad_b[wo] = (FULL || (ok_b && wo < wvalid)) ? __ldg(addp + o + rowstride) : 0.f;
        }
    }
    const float* base = tile + (size_t)(2 * ry) * kTileCols * 32 + lane;
    float win[6][5];
#pragma unroll
    for (int j = 0; j < 4; ++j) load_window_column(win, base, j);
#pragma unroll
    for (int wo = 0; wo < kTW; ++wo) {
        load_window_column(win, base, wo + 4);
        float acc_a = 0.f, acc_b = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                acc_a = fmaf(win[ky][(wo + kx) % 5], wr[ky][kx], acc_a);
                acc_b = fmaf(win[ky + 1][(wo + kx) % 5], wr[ky][kx], acc_b);
            }
        }
        float va = (acc_a + bv) + cv;            // reference order: (conv + bias) + condition
        float vb = (acc_b + bv) + cv;
        if (ADD) {
            va += ad_a[wo];
            vb += ad_b[wo];
        }
        if (ROUND) {
            va = round_tf32(va);
            vb = round_tf32(vb);
        }
        const size_t o = off + (size_t)wo * C;
        if (FULL || (ok_a && wo < wvalid)) {
            outp[o] = va;
            if (CSUM) csum += va;
        }
        if (FULL || (ok_b && wo < wvalid)) {
            outp[o + rowstride] = vb;
            if (CSUM) csum += vb;
        }
    }
}

// CSUM: the CTA also writes the per-channel sum of everything it stored to csum_part[part][C] (part = (seg, th, b));
// the data-gradient launch of block l produces the upstream gradient of block l-1, whose bias gradient is exactly
// that column sum -- one extra add per output instead of another full pass over the tensor.
template <bool ADD, bool ROUND, bool CSUM>
__global__ void __launch_bounds__(32 * kRowThreads, 2)
dw5x5_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ wgt,
                 const float* __restrict__ bias, const float* __restrict__ cond, const float* __restrict__ add,
                 float* __restrict__ out, int H, int W, int C, int tiles_w, int ncg, int tpc, int flip,
                 float* __restrict__ csum_part) {
    extern __shared__ __align__(128) uint8_t dsm[];
    float* tiles = reinterpret_cast<float*>(dsm);   // two halo boxes
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm + 2 * kTileFloats * sizeof(float));
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int cg = blockIdx.x % ncg, seg = blockIdx.x / ncg;
    const int th = blockIdx.y, b = blockIdx.z;
    const int tid = ry * 32 + lane;
    const int tw0 = seg * tpc;
    const int ntile = min(tpc, tiles_w - tw0);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    pdl_grid_sync();
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar[0], kTileFloats * sizeof(float));
        tma_load_4d(tiles, &tm_in, &bar[0], cg * 32, tw0 * kTW - 2, th * kTH - 2, b);
    }
    const int c = cg * 32 + lane;
    const bool cok = c < C;
    const int cc = cok ? c : 0;
    float wr[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) wr[t / 5][t % 5] = __ldg(wgt + cc * 25 + (flip ? 24 - t : t));
    const float bv = bias ? __ldg(bias + cc) : 0.f;
    const float cv = cond ? __ldg(cond + (size_t)b * C + cc) : 0.f;
    const int h = th * kTH + 2 * ry;
    const bool ok_a = cok && h < H, ok_b = cok && h + 1 < H;
    const bool rows_full = th * kTH + kTH <= H && cg * 32 + 32 <= C;   // CTA-uniform
    const size_t rowstride = (size_t)W * C;
    const size_t rowoff = ((size_t)b * H + (h < H ? h : 0)) * rowstride + cc;
    float csum = 0.f;

    for (int i = 0; i < ntile; ++i) {
        const int buf = i & 1;
        if (tid == 0 && i + 1 < ntile) {
            // the other buffer was consumed in iteration i-1 (all threads passed the __syncthreads below)
            mbar_arrive_expect_tx(&bar[buf ^ 1], kTileFloats * sizeof(float));
            tma_load_4d(tiles + (size_t)(buf ^ 1) * kTileFloats, &tm_in, &bar[buf ^ 1], cg * 32,
                        (tw0 + i + 1) * kTW - 2, th * kTH - 2, b);
        }
        const int w0 = (tw0 + i) * kTW;
        const size_t off = rowoff + (size_t)w0 * C;
        const float* tile = tiles + (size_t)buf * kTileFloats;
        if (rows_full && w0 + kTW <= W) {
            // (the residual operand loads are issued before the wait: they overlap the box's arrival)
            mbar_wait(&bar[buf], (uint32_t)(i >> 1) & 1u);
            dw5x5_tile<ADD, ROUND, true, CSUM>(tile, ry, lane, wr, bv, cv, add, out, off, C, rowstride, kTW, true, true,
                                                   csum);
        } else {
            mbar_wait(&bar[buf], (uint32_t)(i >> 1) & 1u);
            dw5x5_tile<ADD, ROUND, false, CSUM>(tile, ry, lane, wr, bv, cv, add, out, off, C, rowstride, W - w0, ok_a,
                                                    ok_b, csum);
        }
        // generic-proxy reads of `buf` are followed by an async-proxy (TMA) refill: proxy fence, then the barrier
        fence_proxy_async_smem();
        __syncthreads();   // everyone is done with `buf` before it is refilled
    }
    if (CSUM) {
        float* red = tiles;   // [8][32]; the boxes are dead after the loop's last __syncthreads
        red[ry * 32 + lane] = csum;
        __syncthreads();
        if (ry == 0 && cok) {
            float t = 0.f;
#pragma unroll
            for (int y = 0; y < kRowThreads; ++y) t += red[y * 32 + lane];
            const size_t part = ((size_t)seg * gridDim.y + th) * gridDim.z + b;
            csum_part[part * C + c] = t;
        }
    }
}

// out[c] = sum_k part[k][C]: block = 32 channels x 8 partial-sum lanes
__global__ void __launch_bounds__(256) csum_final_kernel(const float* __restrict__ part, int nparts, int C,
                                                         float* __restrict__ out, float* __restrict__ out2) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    __shared__ float red[8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float t = 0.f;
    if (c < C)
        for (int k = threadIdx.y; k < nparts; k += 8) t += part[(size_t)k * C + c];
    red[threadIdx.y][threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float v = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) v += red[y][threadIdx.x];
        out[c] = v;
        if (out2) out2[c] = v;
    }
}

// Weight / bias / conditioning gradients with the same tiling: thread (channel, row pair) keeps the 25 tap sums
// and the plain sum of its channel in registers while its CTA walks the tiles of (up to) half a tile row; the 32
// upstream-gradient values of a tile are fetched together before the halo box is waited for.  Out-of-image pixels
// contribute through a zero upstream gradient, so there is no bounds code in the FMA loop.
// scratch[b][chunk = th * nseg + seg][26][C]
__global__ void __launch_bounds__(32 * kRowThreads, 2)
dw5x5_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ dh,
                       float* __restrict__ scratch, int H, int W, int C, int tiles_w, int ncg, int tpc, int nseg) {
    extern __shared__ __align__(128) uint8_t dsm[];
    float* tiles = reinterpret_cast<float*>(dsm);   // 2 buffers; reused for the final reduction
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm + 2 * kTileFloats * sizeof(float));
    const int lane = threadIdx.x, ry = threadIdx.y;
    const int cg = blockIdx.x % ncg, seg = blockIdx.x / ncg;
    const int th = blockIdx.y, b = blockIdx.z;
    const int tid = ry * 32 + lane;
    const int c = cg * 32 + lane;
    const int h = th * kTH + 2 * ry;
    const bool cok = c < C;
    const bool ok_a = cok && h < H, ok_b = cok && h + 1 < H;
    const int tw0 = seg * tpc;
    const int ntile = min(tpc, tiles_w - tw0);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    pdl_grid_sync();
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar[0], kTileFloats * sizeof(float));
        tma_load_4d(tiles, &tm_x, &bar[0], cg * 32, tw0 * kTW - 2, th * kTH - 2, b);
    }
    float acc[5][5];
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t / 5][t % 5] = 0.f;
    float gsum = 0.f;
    const size_t rowstride = (size_t)W * C;
    const float* dhrow = dh + ((size_t)b * H + (h < H ? h : 0)) * rowstride + (cok ? c : 0);

    for (int i = 0; i < ntile; ++i) {
        const int buf = i & 1;
        if (tid == 0 && i + 1 < ntile) {
            mbar_arrive_expect_tx(&bar[buf ^ 1], kTileFloats * sizeof(float));
            tma_load_4d(tiles + (size_t)(buf ^ 1) * kTileFloats, &tm_x, &bar[buf ^ 1], cg * 32,
                        (tw0 + i + 1) * kTW - 2, th * kTH - 2, b);
        }
        const int w0 = (tw0 + i) * kTW;
        float g_a[kTW], g_b[kTW];
#pragma unroll
        for (int wo = 0; wo < kTW; ++wo) {
            const size_t o = (size_t)(w0 + wo) * C;
            g_a[wo] = (ok_a && w0 + wo < W) ? __ldg(dhrow + o) : 0.f;
            g_b[wo] = (ok_b && w0 + wo < W) ? __ldg(dhrow + o + rowstride) : 0.f;
        }
        mbar_wait(&bar[buf], (uint32_t)(i >> 1) & 1u);
        const float* base = tiles + (size_t)buf * kTileFloats + (size_t)(2 * ry) * kTileCols * 32 + lane;
        float win[6][5];
#pragma unroll
        for (int j = 0; j < 4; ++j) load_window_column(win, base, j);
#pragma unroll
        for (int wo = 0; wo < kTW; ++wo) {
            load_window_column(win, base, wo + 4);
            const float ga = g_a[wo], gb = g_b[wo];
            gsum += ga + gb;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 5; ++kx)
                    acc[ky][kx] = fmaf(win[ky + 1][(wo + kx) % 5], gb, fmaf(win[ky][(wo + kx) % 5], ga, acc[ky][kx]));
            }
        }
        fence_proxy_async_smem();   // generic reads -> async-proxy refill (see the forward kernel)
        __syncthreads();   // everyone is done with `buf` before it is refilled two iterations later
    }
    // reduce over the 8 row pairs through shared memory (the tile buffers are free now)
    float* red = tiles;    // [8][26][32]
#pragma unroll
    for (int t = 0; t < 25; ++t) red[(ry * 26 + t) * 32 + lane] = acc[t / 5][t % 5];
    red[(ry * 26 + 25) * 32 + lane] = gsum;
    __syncthreads();
    const int chunk = th * nseg + seg, nchunk = gridDim.y * nseg;
    for (int i = tid; i < 26 * 32; i += 32 * kRowThreads) {
        const int t = i / 32, l = i % 32;
        const int cc = cg * 32 + l;
        if (cc < C) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kRowThreads; ++y) s += red[(y * 26 + t) * 32 + l];
            scratch[(((size_t)b * nchunk + chunk) * 26 + t) * C + cc] = s;
        }
    }
}

}  // namespace

// tiles per CTA along a tile row: long walks amortise the per-CTA setup, but small images still need
// a few CTAs per SM
static int tiles_per_cta(int tiles_w, long long ctas_per_segment, int max_tpc) {
    const long long want = 4ll * (device_info().initialized ? device_info().num_sms : 148);
    int tpc = max_tpc < tiles_w ? max_tpc : tiles_w;
    while (tpc > 1 && ctas_per_segment * ceil_div(tiles_w, tpc) < want) tpc = (tpc + 1) / 2;
    return tpc;
}

template <bool ADD, bool ROUND, bool CSUM>
static int dw5x5_tma_launch(const CUtensorMap& tm, const float* w, const float* bias, const float* cond,
                            const float* add, float* out, int B, int H, int W, int C, int flip, cudaStream_t stream,
                            float* csum_out, float* csum_out2, float* csum_scratch) {
    const size_t smem = 2 * kTileFloats * sizeof(float) + 16;
    static int attr_set = 0;
    if (!attr_set) {
        SINDDM_CUDA_OK(cudaFuncSetAttribute(dw5x5_tma_kernel<ADD, ROUND, CSUM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = 1;
    }
    const int tiles_w = ceil_div(W, kTW), tiles_h = ceil_div(H, kTH), ncg = ceil_div(C, 32);
    const int tpc = tiles_per_cta(tiles_w, (long long)ncg * tiles_h * B, kMaxTilesPerCta);
    dim3 grid(ncg * ceil_div(tiles_w, tpc), tiles_h, B);
    dim3 block(32, kRowThreads);
    (void)launch_pdl(dw5x5_tma_kernel<ADD, ROUND, CSUM>, grid, block, (size_t)smem, stream, tm, w, bias, cond, add, out, H, W,
                     C, tiles_w, ncg, tpc, flip, csum_scratch);
    SINDDM_CUDA_OK(cudaGetLastError());
    if (CSUM) {
        const int nparts = (int)(grid.x / ncg) * tiles_h * B;
        (void)launch_pdl(csum_final_kernel, dim3(ncg), dim3(dim3(32, 8)), (size_t)(0), stream, csum_scratch, nparts, C, csum_out, csum_out2);
        SINDDM_CUDA_OK(cudaGetLastError());
    }
    return SINDDM_OK;
}

size_t dw5x5_csum_scratch_floats(int B, int H, int W, int C) {
    const size_t a = (size_t)ceil_div(W, kTW) * ceil_div(H, kTH) * B * C;
    const size_t b = colsum_scratch_floats(C);
    return a > b ? a : b;
}

int dw5x5_launch(const float* in, const float* w, const float* bias, const float* cond, const float* add, float* out,
                 int B, int H, int W, int C, int flip, int round_tf32, cudaStream_t stream, float* csum_out,
                 float* csum_out2, float* csum_scratch) {
    SINDDM_REQUIRE(B <= 65535, "dw5x5: batch too large");
    SINDDM_REQUIRE(!csum_out || csum_scratch, "dw5x5: column sums need a scratch buffer");
    if (C % 4 == 0 && device_info().initialized) {
        CUtensorMap tm;
        SINDDM_TRY(make_tmap_nhwc(&tm, in, B, H, W, C, 32, kTileCols, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE));
#define SINDDM_DW_ARGS tm, w, bias, cond, add, out, B, H, W, C, flip, stream, csum_out, csum_out2, csum_scratch
        if (add && csum_out)
            return round_tf32 ? dw5x5_tma_launch<true, true, true>(SINDDM_DW_ARGS)
                              : dw5x5_tma_launch<true, false, true>(SINDDM_DW_ARGS);
        if (add)
            return round_tf32 ? dw5x5_tma_launch<true, true, false>(SINDDM_DW_ARGS)
                              : dw5x5_tma_launch<true, false, false>(SINDDM_DW_ARGS);
        if (!csum_out)
            return round_tf32 ? dw5x5_tma_launch<false, true, false>(SINDDM_DW_ARGS)
                              : dw5x5_tma_launch<false, false, false>(SINDDM_DW_ARGS);
        // (column sums without a residual operand: not needed by the network; separate pass below)
        const int rc = round_tf32 ? dw5x5_tma_launch<false, true, false>(SINDDM_DW_ARGS)
                                  : dw5x5_tma_launch<false, false, false>(SINDDM_DW_ARGS);
        SINDDM_TRY(rc);
#undef SINDDM_DW_ARGS
        SINDDM_TRY(colsum_launch(out, (long long)B * H * W, C, csum_out, csum_scratch, stream));
        if (csum_out2)
            SINDDM_CUDA_OK(cudaMemcpyAsync(csum_out2, csum_out, sizeof(float) * C, cudaMemcpyDeviceToDevice, stream));
        return SINDDM_OK;
    }
    if (csum_out) {
        SINDDM_TRY(dw5x5_launch(in, w, bias, cond, add, out, B, H, W, C, flip, round_tf32, stream, nullptr, nullptr,
                                nullptr));
        SINDDM_TRY(colsum_launch(out, (long long)B * H * W, C, csum_out, csum_scratch, stream));
        if (csum_out2)
            SINDDM_CUDA_OK(cudaMemcpyAsync(csum_out2, csum_out, sizeof(float) * C, cudaMemcpyDeviceToDevice, stream));
        return SINDDM_OK;
    }
    if (C == 3) {
        const long long P = (long long)B * H * W;
        const long long cap = 8ll * (device_info().initialized ? device_info().num_sms : 148);
        const long long want = (P + 255) / 256;
        (void)launch_pdl(dw5x5_c3_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), (size_t)(0), stream, in, w, bias, cond, add, out, B, H, W,
                                                                                 flip, round_tf32);
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

// both weight-gradient kernels write at most ceil(H/8) + 1 chunks of 26 x C partial sums per image: the CUDA-core
// kernel one per 8 rows, the TMA kernel at most two (row halves) per 16 rows
size_t dw5x5_wgrad_scratch_floats(int B, int H, int C) {
    return (size_t)B * (ceil_div(H, kRowsPerChunk) + 1) * 26 * C;
}

int dw5x5_wgrad_launch(const float* x, const float* dh, float* dw, float* db, float* dcond, float* scratch, int B,
                       int H, int W, int C, cudaStream_t stream) {
    SINDDM_REQUIRE(B <= 65535, "dw5x5_wgrad: batch too large");
    static_assert(2 * kRowsPerChunk == kTH, "scratch bound assumes 8-row chunks vs 16-row tiles");
    int nchunk;
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
        const int tiles_w = ceil_div(W, kTW), tiles_h = ceil_div(H, kTH), ncg = ceil_div(C, 32);
        // at most two segments per tile row (the scratch bound above)
        int tpc = tiles_per_cta(tiles_w, (long long)ncg * tiles_h * B, tiles_w);
        if (tpc < ceil_div(tiles_w, 2)) tpc = ceil_div(tiles_w, 2);
        const int nseg = ceil_div(tiles_w, tpc);
        nchunk = tiles_h * nseg;
        dim3 grid(ncg * nseg, tiles_h, B);
        dim3 block(32, kRowThreads);
        (void)launch_pdl(dw5x5_wgrad_tma_kernel, grid, block, (size_t)smem, stream, tm, dh, scratch, H, W, C, tiles_w, ncg, tpc,
                         nseg);
    } else if (C == 3) {
        nchunk = ceil_div(H, kRowsPerChunk);
        if (nchunk > 16) nchunk = 16;
        (void)launch_pdl(dw5x5_wgrad_c3_kernel, dim3(dim3(nchunk, B)), dim3(256), (size_t)(0), stream, x, dh, scratch, H, W, nchunk);
    } else {
        nchunk = ceil_div(H, kRowsPerChunk);
        dim3 grid(ceil_div(C, 32), nchunk, B);
        dim3 block(32, kRowsPerChunk);
        dw5x5_wgrad_partial_kernel<<<grid, block, 0, stream>>>(x, dh, scratch, B, H, W, C, nchunk);
    }
    SINDDM_CUDA_OK(cudaGetLastError());
    SINDDM_REQUIRE((size_t)B * 32 * sizeof(float) <= 48 * 1024, "dw5x5_wgrad: batch %d too large for the final reduction", B);
    dim3 grid2(ceil_div(C, 32), 26);
    (void)launch_pdl(dw5x5_wgrad_final_kernel, grid2, dim3(32, kFinalBatchLanes), (size_t)B * 32 * sizeof(float), stream,
                     scratch, dw, db, dcond, B, C, nchunk);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
