// Convolution weight gradient on the sm_100a tensor cores (tcgen05 + TMEM + TMA), split-K over pixels.
//
// Replaces cuDNN's backward-filter behind autograd for SinDDMConvBlock.net[0], net[2], res_conv
// (reference SinDDM/models.py:62-67; trainer.py:202 `loss.backward()`).
//
//   dW[tap][ci][co] = sum_p x[p (+) tap][ci] * dy[p][co]
//
//   GEMM view      K = pixels, walked in steps of one 32-pixel row segment (b, h, 32 w)
//                  M = channels of x (32-channel chunks), N = Cy (all channels of dy, one UMMA N).
//   operands       both are "MN-major" (channels contiguous, pixels strided): [pixels][32 ch] fp32 boxes, TMA-loaded
//                  with the 128B/32B-atom swizzle, the only smem layout the tensor core accepts for MN-major tf32.
//   halo sharing   the kernel is bound by the bytes entering shared memory, so the three horizontal taps share
//                  their data: a "region" is the (32 ch, 34 px) box of one image row h+dy starting at w0-1
//                  (out-of-image pixels zero-filled by TMA = the convolution's padding), and tap dx is the same
//                  box read from pixel 1+dx on -- the operand descriptor simply starts 128 B (one pixel) later.
//                  A CTA stages up to 4 regions = (vertical tap, channel chunk) pairs per K step and owns the
//                  3 x (4 x 32) x Cy accumulators of their three horizontal taps: 3 MMA tiles from 4 boxes
//                  instead of 12 boxes (1.8x fewer bytes per FLOP at 160 channels, 2x at Cy = 80).
//   accumulators   up to 3 tiles x Cy fp32 columns in TMEM, resident for the CTA's whole pixel range.
//   issue          one MMA-issuing warp per accumulator tile (= horizontal tap): a K step is 12 MMAs, and one warp needs
//                  ~60 cycles per MMA plus ~230 per mbarrier wait / tcgen05.commit (tools/pipe_bench.cu) -- ~1300 cycles
//                  per K step against 480 (Cy = 80) / 960 (Cy = 160) cycles of tensor-pipe work.  Three issuers (4 MMAs +
//                  one commit each, the next stage's barrier tested inside the MMA asm, no tcgen05 fence per step)
//                  bring the issue path under the pipe time.  Every accumulator still has ONE issuing thread walking K
//                  in order, so the summation order -- and the result -- stay bit-reproducible.
//   grid           nsplit (pixel ranges) x ngroups (region sets); each CTA writes its partial [tap][ci][co] block,
//                  a second kernel (wgrad_reduce) sums the splits in fixed order (deterministic).
#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kKP = 32;                   // pixels per K step
constexpr int kCC = 32;                   // channels per chunk box
constexpr int kRowBytes = kCC * 4;        // one pixel of a box: 128 B
constexpr int kDyBoxBytes = kKP * kRowBytes;   // 4 KiB
constexpr int kRegPerTile = 4;            // 32-channel regions per M = 128 tile
constexpr int kMaxTiles = 3;
constexpr int kThreads = 256;             // warp 0: TMA producer, warps 1-3: one MMA issuer per accumulator tile, warps 4-7: epilogue
constexpr int kIssuers = kMaxTiles;

struct KernelArgs {
    int B, H, W;
    int Cx, Cy, ntaps;
    int cxk, cyk;
    int nregions;                // 3 * cxk (3x3: one per vertical tap and chunk) or cxk (1x1)
    int nreg_cta;                // regions staged per CTA
    int ntile;                   // accumulator tiles per CTA: the 3 horizontal taps (3x3) or ceil(nreg_cta / 4) (1x1)
    int nsplit;
    int wt, nks;                 // 32-px segments per row, total K steps
    int rb;                      // bytes between region boxes in shared memory
    int col_stride;              // TMEM columns between accumulator tiles
    int tmem_cols;
    int nstages, stage_bytes, x_bytes;
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
    const bool k3 = a.ntaps == 9;

    // K-step range of this split (balanced, contiguous)
    const int ks_begin = (int)(((long long)a.nks * split) / a.nsplit);
    const int ks_end = (int)(((long long)a.nks * (split + 1)) / a.nsplit);

    // regions of this group: local q <-> global region r0 + q
    const int r0 = group * a.nreg_cta;
    int nvalid = a.nregions - r0;
    if (nvalid > a.nreg_cta) nvalid = a.nreg_cta;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_dy);
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], (uint32_t)a.ntile);   // one commit per issuing warp
        }
        mbar_init(done_bar, (uint32_t)a.ntile);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    pdl_grid_sync();   // the set-up above overlaps the previous kernel's tail; no global access before this point
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // Roles 0 and 1 run with all 32 lanes in uniform control flow; the issuing lane is elected inside the asm.
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int pw = k3 ? kKP + 2 : kKP;
        const uint32_t tx_bytes = (uint32_t)(nvalid * pw * kRowBytes + a.cyk * kDyBoxBytes);
        for (int ks = ks_begin; ks < ks_end; ++ks) {
            const int per_img = a.H * a.wt;
            const int b = ks / per_img;
            const int r = ks - b * per_img;
            const int h = r / a.wt;
            const int w0 = (r - h * a.wt) * kKP;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sx = smem + (size_t)stage * a.stage_bytes;
            uint8_t* sy = sx + a.x_bytes;
            mbar_arrive_expect_tx_w(&full_bar[stage], tx_bytes);
            for (int j = 0; j < a.cyk; ++j)
                tma_load_4d_w(sy + j * kDyBoxBytes, &tm_dy, &full_bar[stage], j * kCC, w0, h, b);
            for (int q = 0; q < nvalid; ++q) {
                const int rg = r0 + q;
                const int dyi = rg / a.cxk;              // vertical tap (3x3) -- always 0 for 1x1
                const int j = rg - dyi * a.cxk;
                tma_load_4d_w(sx + q * a.rb, &tm_x, &full_bar[stage], j * kCC, k3 ? w0 - 1 : w0, k3 ? h + dyi - 1 : h, b);
            }
            if (++stage == a.nstages) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else if (warp <= kIssuers) {
        // tile i = warp - 1:  3x3: horizontal tap i, the box read from pixel i on;  1x1: regions 4i..4i+3
        const int i = warp - 1;
        if (i < a.ntile) {
            int stage = 0;
            uint32_t phase = 0;
            // MN-major, 128B swizzle with 32B atoms: 4-pixel atoms are 512 B apart (SBO); the 32-channel boxes of one
            // operand are LBO apart (bits 16.. of the low word): one region stride for x, one box for dy.
            // 8 pixels per MMA = 1024 B = +64 in the start address.
            const uint64_t proto_x = umma_smem_desc(0, (uint32_t)a.rb, 512, UMMA_LAYOUT_SW128_B32);
            const uint64_t proto_y = umma_smem_desc(0, kDyBoxBytes, 512, UMMA_LAYOUT_SW128_B32);
            const uint32_t desc_hi = (uint32_t)(proto_y >> 32);
            const uint32_t lbo_x = (uint32_t)proto_x, lbo_y = (uint32_t)proto_y;
            const uint32_t dacc = tmem_base + (uint32_t)(i * a.col_stride);
            const uint32_t tile_off = k3 ? (uint32_t)i * kRowBytes : (uint32_t)(i * kRegPerTile * a.rb);
            bool ready = false;   // the next stage's full barrier was seen complete by the test inside the MMA asm
            for (int ks = ks_begin; ks < ks_end; ++ks) {
                if (!ready) mbar_wait(&full_bar[stage], phase);
                const uint32_t sx = smem_u32(smem + (size_t)stage * a.stage_bytes);
                const uint32_t sy = sx + (uint32_t)a.x_bytes;
                const uint32_t b_lo = lbo_y | ((sy >> 4) & 0x3FFFu);
                // A start that is a whole number of pixels (128 B) into the box needs no base-offset field: the
                // tensor core applies the swizzle to absolute shared-memory address bits, like TMA did on the way
                // in (verified on B200: base offset 0 is bit-compatible with the aligned kernel, non-zero is wrong)
                const uint32_t a_lo = lbo_x | (((sx + tile_off) >> 4) & 0x3FFFu);
                const int cur = stage;
                if (++stage == a.nstages) {
                    stage = 0;
                    phase ^= 1u;
                }
                // operands written by TMA and observed through the mbarrier need no tcgen05 fence (as in tc_conv)
                const uint32_t r = umma_tf32_ss_x4_test(dacc, a_lo, b_lo, desc_hi, 64u, a.idesc, (ks != ks_begin) ? 1u : 0u,
                                                        kKP / 8, &full_bar[stage], phase);
                umma_commit_elect(&empty_bar[cur]);
                ready = __all_sync(0xffffffffu, r != 0);
            }
            umma_commit_elect(done_bar);
        }
    } else {
        // epilogue: one accumulator row (= one (tap, ci)) per thread, Cy contiguous floats each
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        mbar_wait(done_bar, 0);
        tc_fence_after_sync();
        for (int i = 0; i < a.ntile; ++i) {
            const int q = (k3 ? 0 : i * kRegPerTile) + (row >> 5);   // local region of this row
            const int rg = r0 + q;
            const int dyi = rg / a.cxk;
            const int j = rg - dyi * a.cxk;
            const int tap = k3 ? dyi * 3 + i : 0;
            const int ci = j * kCC + (row & 31);
            const bool valid = (q < nvalid) && (ci < a.Cx);
            float* dst = a.partial + (((size_t)split * a.ntaps + tap) * a.Cx + ci) * a.Cy;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * a.col_stride);
            for (int cc = 0; cc < a.Cy; cc += 16) {
                float vals[16];
                tmem_ld16(taddr + cc, vals);
                if (valid) {
                    float4* o4 = reinterpret_cast<float4*>(dst + cc);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        o4[t] = make_float4(vals[4 * t], vals[4 * t + 1], vals[4 * t + 2], vals[4 * t + 3]);
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
    int cxk, cyk, nregions, nreg_cta, ntile, ngroups, col_stride, tmem_cols, wt, nks, pw, rb, x_bytes;
};

Shape make_shape(int B, int H, int W, int Cx, int Cy, int ntaps) {
    Shape s;
    s.cxk = ceil_div(Cx, kCC);
    s.cyk = ceil_div(Cy, kCC);
    s.col_stride = (int)align_up((size_t)Cy, 32);
    if (ntaps == 9) {
        s.nregions = 3 * s.cxk;
        // fewest CTAs groups first, then the fewest staged regions per group
        int best = kRegPerTile, best_groups = ceil_div(s.nregions, kRegPerTile);
        for (int n = kRegPerTile - 1; n >= 1; --n)
            if (ceil_div(s.nregions, n) <= best_groups) best = n;
        s.nreg_cta = best;
        s.ntile = 3;
        s.pw = kKP + 2;
        s.x_bytes = 0;   // set below
    } else {
        s.nregions = s.cxk;
        const int cap = kRegPerTile * kMaxTiles;
        const int groups = ceil_div(s.nregions, cap);
        s.nreg_cta = ceil_div(s.nregions, groups);
        s.ntile = ceil_div(s.nreg_cta, kRegPerTile);
        s.pw = kKP;
    }
    s.ngroups = ceil_div(s.nregions, s.nreg_cta);
    s.rb = (int)align_up((size_t)s.pw * kRowBytes, 1024);
    // every tile's descriptor spans 4 regions: keep that reach inside the x area of the stage
    s.x_bytes = (ntaps == 9 ? kRegPerTile : s.ntile * kRegPerTile) * s.rb;
    const int cols = s.ntile * s.col_stride;
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
    op->nregions = s.nregions;
    op->nreg_cta = s.nreg_cta;
    op->ntile = s.ntile;
    op->ngroups = s.ngroups;
    op->wt = s.wt;
    op->nks = s.nks;
    op->pw = s.pw;
    op->rb = s.rb;
    op->col_stride = s.col_stride;
    op->tmem_cols = s.tmem_cols;
    op->x_bytes = s.x_bytes;
    SINDDM_TRY(make_tmap_nhwc(&op->tm_x, p.x, p.B, p.H, p.W, p.Cx, kCC, s.pw, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    SINDDM_TRY(make_tmap_nhwc(&op->tm_dy, p.dy, p.B, p.H, p.W, p.Cy, kCC, kKP, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    op->stage_bytes = s.x_bytes + s.cyk * kDyBoxBytes;
    const int tail = 8 * 8 * 2 + 8 + 16;
    int nst = (device_info().max_smem_optin - 1024 - tail) / op->stage_bytes;
    if (nst > 8) nst = 8;
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
    a.nregions = op.nregions;
    a.nreg_cta = op.nreg_cta;
    a.ntile = op.ntile;
    a.nsplit = p.nsplit;
    a.wt = op.wt;
    a.nks = op.nks;
    a.rb = op.rb;
    a.col_stride = op.col_stride;
    a.tmem_cols = op.tmem_cols;
    a.nstages = op.nstages;
    a.stage_bytes = op.stage_bytes;
    a.x_bytes = op.x_bytes;
    a.idesc = umma_idesc_tf32(128, p.Cy, 1, 1);
    a.partial = p.partial;
    dim3 grid(p.nsplit, op.ngroups);
    prof_begin(stream, 1, 2.0 * (double)p.B * p.H * p.W * p.ntaps * p.Cx * p.Cy);
    (void)launch_pdl(tc_wgrad_kernel, grid, dim3(kThreads), (size_t)op.smem_bytes, stream, op.tm_x, op.tm_dy, a);
    prof_end(stream);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
