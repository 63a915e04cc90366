// Convolution weight gradient on the sm_100a tensor cores (tcgen05 + TMEM + TMA), split-K over pixels.
//
// Replaces cuDNN's backward-filter behind autograd for SinDDMConvBlock.net[0], net[2], res_conv
// (reference SinDDM/models.py:62-67; trainer.py:202 `loss.backward()`).
//
//   dW[tap][ci][co] = sum_p x[p (+) tap][ci] * dy[p][co]
//
//   GEMM view      K = pixels, walked in steps of one 32-pixel row segment (b, h, 32 w)
//                  M = the flattened (tap, 32-channel chunk of x) axis, 4 chunks = one 128-row UMMA tile
//                      -> rows of one tile may come from different taps (different shifted boxes of x),
//                         which keeps M tiles full even though Cx is 80/160;
//                  N = Cy (all output channels of dy, one UMMA N).
//   operands       both are "MN-major" (channels contiguous, pixels strided): each 32px x 32ch fp32 box
//                  is TMA-loaded with the 128B/32B-atom swizzle, the only smem layout the tensor core
//                  accepts for MN-major tf32.  Shifted x boxes use out-of-bounds zero fill for padding.
//   accumulators   tpc M-tiles x Cy fp32 columns in TMEM (<= 512), resident for the whole K range.
//   grid           ngroups (M-tile groups) x nsplit (pixel ranges); each CTA writes its partial
//                  [tap][ci][co] block, a second kernel (wgrad_reduce) sums the splits in fixed order.
#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kKP = 32;                   // pixels per K step
constexpr int kCC = 32;                   // channels per chunk box
constexpr int kBoxBytes = kKP * kCC * 4;  // 4 KiB
constexpr int kThreads = 192;

struct KernelArgs {
    int B, H, W;
    int Cx, Cy, ntaps;
    int cxk, cyk, nchunks;       // chunks per tap (x), chunks of dy, ntaps*cxk
    int tpc, ngroups, nsplit;    // M tiles per CTA, groups, pixel splits
    int wt, nks;                 // 32-px segments per row, total K steps
    int col_stride;              // TMEM columns between M-tile accumulators
    int tmem_cols;
    int nstages, stage_bytes;
    uint32_t idesc;
    float* partial;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                const KernelArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + (size_t)a.nstages * a.stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* done_bar = empty_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int group = blockIdx.y;
    const int split = blockIdx.x;

    // K-step range of this split (balanced, contiguous)
    const int ks_begin = (int)(((long long)a.nks * split) / a.nsplit);
    const int ks_end = (int)(((long long)a.nks * (split + 1)) / a.nsplit);

    // chunk slots of this group: slot q <-> global chunk id g0 + q, valid while < nchunks
    const int nslots = a.tpc * 4;
    const int g0 = group * nslots;
    int nvalid = a.nchunks - g0;
    if (nvalid > nslots) nvalid = nslots;
    const int ntile_valid = (nvalid + 3) >> 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_dy);
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int a_bytes = nslots * kBoxBytes;  // x region of a stage; dy region follows

    // Roles 0 and 1 run with all 32 lanes in uniform control flow; the issuing lane is elected inside the asm.
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = (uint32_t)(nvalid + a.cyk) * kBoxBytes;
        for (int ks = ks_begin; ks < ks_end; ++ks) {
            const int per_img = a.H * a.wt;
            const int b = ks / per_img;
            const int r = ks - b * per_img;
            const int h = r / a.wt;
            const int w0 = (r - h * a.wt) * kKP;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sx = smem + (size_t)stage * a.stage_bytes;
            uint8_t* sy = sx + a_bytes;
            mbar_arrive_expect_tx_w(&full_bar[stage], tx_bytes);
            for (int j = 0; j < a.cyk; ++j)
                tma_load_4d_w(sy + j * kBoxBytes, &tm_dy, &full_bar[stage], j * kCC, w0, h, b);
            for (int q = 0; q < nvalid; ++q) {
                const int gq = g0 + q;
                const int tap = gq / a.cxk;
                const int j = gq - tap * a.cxk;
                int dy = 0, dx = 0;
                if (a.ntaps == 9) {
                    dy = tap / 3 - 1;
                    dx = tap % 3 - 1;
                }
                tma_load_4d_w(sx + q * kBoxBytes, &tm_x, &full_bar[stage], j * kCC, w0 + dx, h + dy, b);
            }
            if (++stage == a.nstages) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        // MN-major, 128B swizzle with 32B atoms: 4-pixel atoms are 512 B apart (SBO), 32-channel chunk boxes are
        // kBoxBytes apart (LBO, bits 16.. of the low word); 8 pixels per MMA = 1024 B = +64 in the start address.
        const uint64_t proto = umma_smem_desc(0, kBoxBytes, 512, UMMA_LAYOUT_SW128_B32);
        const uint32_t desc_hi = (uint32_t)(proto >> 32);
        const uint32_t lbo_lo = (uint32_t)proto;
        for (int ks = ks_begin; ks < ks_end; ++ks) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            const uint32_t sx = smem_u32(smem + (size_t)stage * a.stage_bytes);
            const uint32_t sy = sx + a_bytes;
            const uint32_t acc = (ks != ks_begin) ? 1u : 0u;
            const uint32_t b_lo = lbo_lo | ((sy >> 4) & 0x3FFFu);
            for (int i = 0; i < ntile_valid; ++i) {
                const uint32_t a_lo = lbo_lo | (((sx + (uint32_t)i * 4u * kBoxBytes) >> 4) & 0x3FFFu);
                umma_tf32_ss_x4(tmem_base + (uint32_t)(i * a.col_stride), a_lo, b_lo, desc_hi, 64u, a.idesc, acc,
                                kKP / 8);
            }
            umma_commit_elect(&empty_bar[stage]);
            if (++stage == a.nstages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        umma_commit_elect(done_bar);
    } else {
        // epilogue: one accumulator row (= one (tap, ci)) per thread, Cy contiguous floats each
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        mbar_wait(done_bar, 0);
        tc_fence_after_sync();
        for (int i = 0; i < ntile_valid; ++i) {
            const int v = (group * a.tpc + i) * 128 + row;   // virtual (tap, chunk, channel) row
            const int gq = v >> 5;
            const int tap = gq / a.cxk;
            const int ci = (gq - tap * a.cxk) * kCC + (v & 31);
            const bool valid = (gq < a.nchunks) && (ci < a.Cx);
            float* dst = a.partial + (((size_t)split * a.ntaps + tap) * a.Cx + ci) * a.Cy;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * a.col_stride);
            for (int cc = 0; cc < a.Cy; cc += 16) {
                float vals[16];
                tmem_ld16(taddr + cc, vals);
                if (valid) {
                    float4* o4 = reinterpret_cast<float4*>(dst + cc);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        o4[q] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

struct Shape {
    int cxk, cyk, nchunks, tpc, ngroups, col_stride, tmem_cols, wt, nks;
};

Shape make_shape(int B, int H, int W, int Cx, int Cy, int ntaps) {
    Shape s;
    s.cxk = ceil_div(Cx, kCC);
    s.cyk = ceil_div(Cy, kCC);
    s.nchunks = ntaps * s.cxk;
    s.col_stride = (int)align_up((size_t)Cy, 32);
    int tpc = 512 / s.col_stride;
    if (tpc > 3) tpc = 3;                    // smem: 3 tiles x 4 boxes + dy boxes per stage
    const int ntile = ceil_div(s.nchunks, 4);
    if (tpc > ntile) tpc = ntile;
    s.tpc = tpc;
    s.ngroups = ceil_div(ntile, tpc);
    int cols = tpc * s.col_stride;
    s.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    s.wt = ceil_div(W, kKP);
    s.nks = B * H * s.wt;
    return s;
}

}  // namespace

bool tc_wgrad_supported(int Cx, int Cy) {
    return Cx >= 8 && Cx % 4 == 0 && Cy >= 16 && Cy % 16 == 0 && Cy <= 160;
}

int tc_wgrad_nsplit(int B, int H, int W, int Cx, int Cy, int ntaps) {
    const Shape s = make_shape(B, H, W, Cx, Cy, ntaps);
    int sms = device_info().initialized ? device_info().num_sms : 148;
    int nsplit = sms / s.ngroups;
    if (nsplit < 1) nsplit = 1;
    // keep at least 8 K steps per CTA so the pipeline fill is amortised
    int cap = s.nks / 8;
    if (cap < 1) cap = 1;
    if (nsplit > cap) nsplit = cap;
    return nsplit;
}

int tc_wgrad_prepare(const WgradProblem& p, TcWgradOp* op) {
    SINDDM_REQUIRE(tc_wgrad_supported(p.Cx, p.Cy), "tc_wgrad: unsupported channels Cx=%d Cy=%d", p.Cx, p.Cy);
    SINDDM_REQUIRE(p.ntaps == 9 || p.ntaps == 1, "tc_wgrad: ntaps must be 9 or 1");
    SINDDM_REQUIRE(device_info().initialized, "sinddm_init() has not been called");
    const Shape s = make_shape(p.B, p.H, p.W, p.Cx, p.Cy, p.ntaps);
    SINDDM_REQUIRE(p.nsplit >= 1 && p.nsplit <= s.nks, "tc_wgrad: nsplit=%d out of range (K steps %d)", p.nsplit, s.nks);
    op->p = p;
    op->cxk = s.cxk;
    op->cyk = s.cyk;
    op->nchunks = s.nchunks;
    op->tpc = s.tpc;
    op->ngroups = s.ngroups;
    op->wt = s.wt;
    op->nks = s.nks;
    op->col_stride = s.col_stride;
    op->tmem_cols = s.tmem_cols;
    SINDDM_TRY(make_tmap_nhwc(&op->tm_x, p.x, p.B, p.H, p.W, p.Cx, kCC, kKP, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    SINDDM_TRY(make_tmap_nhwc(&op->tm_dy, p.dy, p.B, p.H, p.W, p.Cy, kCC, kKP, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    op->stage_bytes = (s.tpc * 4 + s.cyk) * kBoxBytes;
    const int tail = 8 * 8 * 2 + 8 + 16;
    int nst = (device_info().max_smem_optin - 1024 - tail) / op->stage_bytes;
    if (nst > 6) nst = 6;
    SINDDM_REQUIRE(nst >= 2, "tc_wgrad: not enough shared memory");
    op->nstages = nst;
    op->smem_bytes = nst * op->stage_bytes + tail + 1024;
    return SINDDM_OK;
}

int tc_wgrad_launch(const TcWgradOp& op, cudaStream_t stream) {
    static int smem_set = 0;
    if (!smem_set) {
        SINDDM_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            device_info().max_smem_optin));
        smem_set = 1;
    }
    const WgradProblem& p = op.p;
    KernelArgs a;
    a.B = p.B;
    a.H = p.H;
    a.W = p.W;
    a.Cx = p.Cx;
    a.Cy = p.Cy;
    a.ntaps = p.ntaps;
    a.cxk = op.cxk;
    a.cyk = op.cyk;
    a.nchunks = op.nchunks;
    a.tpc = op.tpc;
    a.ngroups = op.ngroups;
    a.nsplit = p.nsplit;
    a.wt = op.wt;
    a.nks = op.nks;
    a.col_stride = op.col_stride;
    a.tmem_cols = op.tmem_cols;
    a.nstages = op.nstages;
    a.stage_bytes = op.stage_bytes;
    a.idesc = umma_idesc_tf32(128, p.Cy, 1, 1);
    a.partial = p.partial;
    dim3 grid(p.nsplit, op.ngroups);
    prof_begin(stream, 1, 2.0 * (double)p.B * p.H * p.W * p.ntaps * p.Cx * p.Cy);
    tc_wgrad_kernel<<<grid, kThreads, op.smem_bytes, stream>>>(op.tm_x, op.tm_dy, a);
    prof_end(stream);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
