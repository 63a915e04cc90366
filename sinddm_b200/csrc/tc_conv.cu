// 3x3 / 1x1 convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the cuDNN implicit-GEMM calls behind nn.Conv2d in SinDDMConvBlock.net / res_conv
// (reference SinDDM/models.py:62-67,79-80) and, with data-gradient-packed weights, their backward.
//
//   GEMM view      M = pixels: one CTA tile = 16 rows x 16 cols of one image = two UMMA M=128 halves
//                  N = output channels (one UMMA N, 16..160)
//                  K = taps x input channels (+ an optional 1x1 residual conv as extra K from a second input)
//
// With fp32 (TF32) operands the wide layers are bound by the bytes that travel L2 -> shared memory (9 GB per
// 160 -> 160 launch at 8.8 TB/s, ~85 % of the chip-wide limit; profiles/r01d_ncu_tc_conv.txt) before the tensor pipe
// (76 % active), so the K walk is arranged for operand reuse INSIDE shared memory:
//   * one pipeline stage = (32-channel chunk c, horizontal tap kx).  Its A box is the (16+2) x 16 pixel halo
//     tile shifted by kx-1 columns (4-D TMA box (32 ch, 16 w, 18 h, 1 b); out-of-image pixels and the channel
//     tail are zero-filled by TMA = exact zero padding, no im2col buffer).  The three vertical taps ky and
//     the two M halves are just different 2 KiB-aligned row offsets into that one box:
//         A(ky, half) = box + (ky + 8*half) * 16 px * 128 B
//     so every loaded activation byte feeds 3 taps, and every weight byte (3 boxes [N][32] per stage) feeds
//     256 pixels: 2.3x fewer bytes per FLOP than a tap-by-tap 128-pixel walk.
//   * the halo boxes (36 KiB) and the weight boxes (N x 128 B, one per tap) travel through TWO independent
//     mbarrier rings (3 activation slots, 4-8 weight slots), so ~200 KiB of loads stay in flight and the
//     L2 latency under load is covered although one activation box feeds 24 MMAs.
//   * operands land in 128B-swizzled K-major tiles that the UMMA descriptors consume directly.
//   accumulators   fp32 in TMEM: three slots of N columns; tile pair p uses slots (2p, 2p+1) mod 3, so the
//                  main loop of pair p+1 only waits for the epilogue of the FIRST half of pair p.
//   roles          (warpgroup aligned) warp 0: TMA producer | warp 1: TMEM alloc | warps 1 and 2: MMA issue, one per
//                  M = 128 half of the tile (whole warps in uniform control flow, the issuing lane is elected inside
//                  the asm: this removed a ~250-cycle/MMA issue cost; two issuers because one cannot feed N = 80)
//                  warp 3: optional L2 prefetcher of the streamed epilogue operand (off: measured neutral)
//                  warps 4-11: epilogue (TMEM -> registers -> bias/residual/GELU/... -> smem tile -> TMA store), two
//                  warps per TMEM lane quarter splitting the columns; a streamed operand (residual / saved gelu'(z))
//                  arrives by TMA one chunk ahead.  The epilogue is bound by instruction issue, not latency: its
//                  math is spelled out instruction by instruction (common.cuh), the forward pass saves gelu'(z) so
//                  the data gradient only multiplies.
//   barriers       every mbarrier wait / tcgen05.commit / tcgen05.fence costs its warp 150-260 cycles
//                  (tools/pipe_bench.cu): waits are tested one box ahead inside the MMA asm, there is no per-box
//                  tcgen05 fence, and generic <-> async proxy hand-overs of a shared-memory tile carry explicit
//                  proxy fences in BOTH directions (the read -> TMA-refill direction was a real, rare race).
//   grid           persistent, min(#tiles, #SMs) CTAs, static round-robin over tiles.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kTileH = 16;
constexpr int kTileW = 16;
constexpr int kBoxH = kTileH + 2;            // halo rows
constexpr int kKC = 32;                      // channels per K chunk (32 fp32 = one 128B swizzle row)
constexpr int kRowBytes = kTileW * kKC * 4;  // one image row of the box: 16 px x 128 B = 2 KiB
constexpr int kABytes = kBoxH * kRowBytes;   // 36 KiB
constexpr int kMaxN = 160;
constexpr int kThreads = 384;       // warp 0 TMA, warps 1-2 MMA (one per half tile), warp 3 optional L2 prefetcher, warps 4-11 epilogue
constexpr int kEpiWarp0 = 4;        // roles are warpgroup aligned so that setmaxnreg can move registers between them
constexpr int kEpiThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kSlots = 3;        // accumulator slots rotated by two-half tiles
constexpr int kMaxSlots = 5;     // ... by three-half tiles (narrow layers: 5 x 96 columns)
constexpr int kStagesA = 3;   // activation (halo box) ring (halo mode: 2 slots of the larger box)
// "halo" mode (3x3 layers): ONE (32 ch, 18 w, 18 h) box per channel chunk serves all nine taps -- tap (ky, kx) of the
// M = 128 half tile hf (16 rows x 8 pixels) is the same box read from pixel (ky * 18 + kx + 8 * hf) on, 8-pixel core
// groups one box row (18 x 128 B) apart: the start address moves by whole pixels (128 B), the swizzle pattern TMA wrote
// and the tensor core reads are both functions of the absolute shared-memory address.  Activation bytes entering
// shared memory drop 2.6x (3 x 36 KiB -> 40.5 KiB per chunk); the kernel was bound by exactly those bytes.
// In halo mode the CTA tile is 16 rows x (8 * nh) pixels, nh = 2 or 3 M = 128 "halves" with one accumulator slot and
// one MMA-issuing warp each: with nh = 3 every weight box feeds 384 pixels instead of 256 (weights are 84 % of the
// bytes entering shared memory once the activations come as one box per chunk).
constexpr int kHaloStagesA = 2;
constexpr int kMaxHalves = 3;
constexpr int kMaxStagesB = 8;   // weight box ring
// epilogue staging: every epilogue warp owns two [32 px][16 ch] tiles (64B-swizzled, 2 KiB).  Layers that stream an
// operand in (residual / saved pre-activation) have one result: tile 0 receives the TMA load, tile 1 is drained by
// the TMA store.  Layers without one may have two results (pre-activation + activation): both tiles are store buffers.
constexpr int kEpiChunk = 16;
constexpr int kEpiTileBytes = 32 * kEpiChunk * 4;
constexpr int kEpiWarpBytes = 2 * kEpiTileBytes;
constexpr int kEpiBytes = (kEpiThreads / 32) * kEpiWarpBytes;   // 32 KiB

struct KernelArgs {
    int B, H, W;
    int Cin, ntaps, nchunks;   // main K walk: nchunks x (3 or 1) stages
    int Cres, nchunks_res;     // residual K walk: nchunks_res stages (centre tap only)
    int N;
    int tiles_w, tiles_h, ntiles;
    int nstages_b, bbox_bytes; // weight ring: slots and bytes per slot (N x 128 B rounded up to 1 KiB)
    int slot_stride;           // TMEM columns between accumulator slots
    const float* w_blk;        // non-null: weights in the blocked pre-swizzled layout [tap][chunk][N][32] ...
    const float* wres_blk;     // ... (and the residual 1x1 weights [chunk][N][32]): boxes come by 1-D bulk copy
    uint32_t idesc;
    int issuers2;              // single-CTA kernel: warps 1 and 2 each issue the MMAs of one half tile (SINDDM_TC_ISSUERS=1: warp 1 issues both)
    int peek;                  // MMA warps test the next weight box's barrier inside the MMA asm, no per-box tcgen05 fence (SINDDM_TC_PEEK=0: off)
    int halo;                  // one halo box per chunk for all nine taps; M halves split the tile by columns
    int nh;                    // M = 128 halves per tile (2; halo mode: 2 or 3), tile width = halo ? 8 * nh : 16 pixels
    int nslots;                // accumulator slots the halves rotate through (3; 5 with three halves)
    int abw;                   // halo mode: box width in pixels (8 * nh + 2)
    int nsa, abytes, atx;      // activation ring: slots, bytes between slots, bytes per box
    int stage_release;         // narrow layers (weight ring >= 2 2/3 stages): the MMA warps commit ONE barrier per stage (the
                               // halo box's empty barrier) that also releases the stage's weight boxes, instead of
                               // one commit per box -- every tcgen05.commit costs its warp ~230 cycles, as long as the
                               // four N = 80 MMAs of a box execute (SINDDM_TC_STAGE_RELEASE=0: off)
    int l2pf;                  // warp 3 prefetches the streamed epilogue operand into L2 one tile ahead (SINDDM_TC_L2PF=1)
    int dbg;                   // diagnostics (SINDDM_TC_DEBUG): 1 = no operand loads, 2 = no epilogue traffic, 4 = no MMAs, 8 = stage but do not store
    ConvEpilogue ep;
};

// smem tail (after the 1024-aligned stage ring):
//   uint64 fullA[3], emptyA[3], fullB[8], emptyB[8], tfull[3], tempty[3], epi_in[8]; uint32 tmem_slot[4];
//   float bias[kMaxN], wres3[kMaxN*3] (aliased by wfinal[3*kMaxN]: no layer has both), bfinal[4], fin[128*3]
constexpr int kTailBytes = (2 * kStagesA + 2 * kMaxStagesB + 2 * kMaxSlots + kEpiThreads / 32) * 8 + 16 +
                           (kMaxN + kMaxN * 3 + 4 + 128 * 3) * 4 + 64;

// Epilogue flavours.  The epilogue block (`process`) holds every fused epilogue of the network behind run-time flags
// and is unrolled five times for the drain-first register arrays: ~100 KB of code, of which one layer executes a
// fraction, against a 32 KB L1.5 instruction cache (ncu: `no_instructions` was 10-22 % of the stall samples).  FL is a
// compile-time mask of the features an instantiation MAY use: a feature whose bit is clear is compiled out, a feature
// whose bit is set is still governed by its run-time flag, so any instantiation whose mask covers the layer's features
// computes exactly what the generic one (kFlAll) does.  tc_conv_launch picks the smallest covering instantiation.
enum : int {
    kFlStream = 1,     // a streamed epilogue operand (res_add or dgelu_z) arrives through the per-warp TMA tile
    kFlPre2 = 2,       // ... and a second one through plain loads (res_add AND dgelu_z: operator tests only)
    kFlResAdd = 4,     // identity residual added
    kFlDgelu = 8,      // multiply by gelu'(z) / the saved gelu'
    kFlX3 = 16,        // 3-channel 1x1 residual conv as epilogue FMAs
    kFlGelu = 32,      // GELU (and, with pre_grad, the saved gelu')
    kFlPre = 64,       // pre-activation copy
    kFlFinal = 128,    // fused trailing 1x1 conv to 3 channels
    kFlOut3 = 256,     // [P,3] copy of columns 0..2
    kFlColsum = 512,   // column sums of the stored tile
    kFlRound = 1024,   // round the result to tf32
    kFlAll = 2047
};

// TWO = true: CTA pairs (cluster of 2, cta_group::2): one M=256 MMA covers the same half of BOTH CTAs' tiles,
// each CTA stages only half of the weight rows, and only the leader CTA issues MMAs.
template <bool TWO, int FL>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_ares,
               const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_bres,
               const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_pre,
               const __grid_constant__ CUtensorMap tm_in, const KernelArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    uint8_t* smem_b = smem + (size_t)a.nsa * a.abytes;
    uint8_t* smem_epi = smem_b + (size_t)a.nstages_b * a.bbox_bytes;
    uint8_t* tail = smem_epi + kEpiBytes;
    uint64_t* fulla_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* emptya_bar = fulla_bar + kStagesA;
    uint64_t* fullb_bar = emptya_bar + kStagesA;
    uint64_t* emptyb_bar = fullb_bar + kMaxStagesB;
    uint64_t* tfull_bar = emptyb_bar + kMaxStagesB;
    uint64_t* tempty_bar = tfull_bar + kMaxSlots;
    uint64_t* epi_in_bar = tempty_bar + kMaxSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_in_bar + kEpiThreads / 32);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_wres3 = s_bias + kMaxN;
    float* s_wfinal = s_wres3;   // the 3-channel residual weights and the fused final conv never meet in one layer
    float* s_bfinal = s_wfinal + 3 * kMaxN;

    // warp index through a shuffle so the compiler can prove the role branches are warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int N = a.N;

    // ---------------------------------------------------------------- one-time setup
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        if (!a.w_blk) tma_prefetch_desc(&tm_b);
        if (a.nchunks_res > 0) {
            tma_prefetch_desc(&tm_ares);
            if (!a.w_blk) tma_prefetch_desc(&tm_bres);
        }
        // empty barriers: one commit per MMA-issuing warp (two in the single-CTA kernel, see below)
        for (int i = 0; i < a.nsa; ++i) {
            mbar_init(&fulla_bar[i], TWO ? 2 : 1);   // one arrival per producing CTA
            mbar_init(&emptya_bar[i], (!TWO && a.issuers2) ? a.nh : 1);
        }
        for (int i = 0; i < a.nstages_b; ++i) {
            mbar_init(&fullb_bar[i], TWO ? 2 : 1);
            mbar_init(&emptyb_bar[i], (!TWO && a.issuers2) ? a.nh : 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (TWO ? 2 : 1) * (kEpiThreads / 32));   // one arrival per epilogue warp
        }
        for (int i = 0; i < kEpiThreads / 32; ++i) mbar_init(&epi_in_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (TWO) tmem_alloc_2sm(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    if (warp == 2 && lane == 0) tmem_slot[1] = 0u;   // tiles started by the MMA warps (read by the L2 prefetcher)
    // everything above (descriptor prefetch, barrier init, TMEM allocation) overlaps the previous kernel's tail
    pdl_grid_sync();
    if (warp >= kEpiWarp0) {
        const int t = threadIdx.x - kEpiWarp0 * 32;
        for (int i = t; i < N; i += kEpiThreads) s_bias[i] = a.ep.bias ? a.ep.bias[i] : 0.f;
        if ((FL & kFlX3) && a.ep.w_res3)
            for (int i = t; i < N * 3; i += kEpiThreads) s_wres3[i] = a.ep.w_res3[i];
        if ((FL & kFlFinal) && a.ep.w_final) {
            for (int i = t; i < N * 3; i += kEpiThreads) s_wfinal[i] = a.ep.w_final[i];
            if (t < 3) s_bfinal[t] = a.ep.b_final ? a.ep.b_final[t] : 0.f;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (TWO) cluster_sync_all();   // both CTAs' barriers and TMEM exist before any cross-CTA traffic
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int CS = TWO ? 2 : 1;
    const int crank = TWO ? (int)cluster_ctarank() : 0;
    const int pair_id = blockIdx.x / CS;
    const int npairs = gridDim.x / CS;
    const int nsuper = (a.ntiles + CS - 1) / CS;

    const bool halo = !TWO && a.halo;
    const int nkx = (a.ntaps == 9 && !halo) ? 3 : 1; // stages per chunk: one per horizontal tap (halo mode: one)
    const int nky = a.ntaps == 9 ? (halo ? 9 : 3) : 1;   // weight boxes per stage: vertical taps (halo mode: all nine)
    const int nst_main = a.nchunks * nkx;
    const int nst = nst_main + a.nchunks_res;        // stages per tile
    const uint32_t nbytes = (uint32_t)(N / CS) * kKC * 4;   // one weight box as staged by THIS CTA

    // Roles 0 and 1 run their loops with ALL 32 lanes (uniform control flow); TMA, MMA and commit instructions
    // elect their single issuing lane inside the asm (common.cuh) -- see elect_one_sync() for why.
    // the launch gives every thread 168 registers; the producer / MMA warpgroup hands most of its share to the
    // epilogue warpgroups, which keep a whole half tile's accumulator columns in registers (384*168 = 128*88 + 256*208)
    if (warp < kEpiWarp0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory");
      if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        int sa_i = 0, sb_i = 0;
        uint32_t pha = 0, phb = 0;
        // stage_release bookkeeping: stages are released in order through the halo boxes' empty barriers; `rel` stages
        // (holding `rel_boxes` weight boxes) are known to be released; jst / kbox count the stages / boxes loaded so far
        const bool srel = !TWO && a.stage_release;
        int rel = 0, rel_boxes = 0, rel_local = 0, rel_slot = 0, jst = 0, kbox = 0;
        uint32_t rel_par = 0;
        auto release_one = [&]() {
            mbar_wait(&emptya_bar[rel_slot], rel_par);
            rel_boxes += (rel_local < nst_main) ? nky : 1;
            if (++rel_local == nst) rel_local = 0;
            if (++rel_slot == a.nsa) {
                rel_slot = 0;
                rel_par ^= 1u;
            }
            ++rel;
        };
        for (int st = pair_id; st < nsuper; st += npairs) {
            // a tile index past the end (odd tile count, second CTA of the last pair) runs the same pipeline on
            // zeros: image index >= B is out of bounds for TMA, which zero-fills the box
            const int tile = st * CS + crank;
            const int tw = tile % a.tiles_w;
            const int th = (tile / a.tiles_w) % a.tiles_h;
            const int b = tile / (a.tiles_w * a.tiles_h);
            const int h0 = th * kTileH, w0 = tw * (halo ? 8 * a.nh : kTileW);
            for (int it = 0; it < nst; ++it) {
                const bool main = it < nst_main;
                const int c = main ? it / nkx : it - nst_main;
                const int kx = main ? it - c * nkx : 0;
                // activation halo box: rows h0-1 .. h0+16, columns shifted by the horizontal tap
                if (srel) {
                    while (rel < jst - a.nsa + 1) release_one();
                    ++jst;
                } else {
                    mbar_wait(&emptya_bar[sa_i], pha ^ 1u);
                }
                const int kys = main ? nky : 1;   // weight boxes of the vertical taps that read this halo box
                // stage_release: ONE full barrier per stage -- the halo box's -- also receives the bytes of the stage's
                // weight boxes, so the producer arrives once and the MMA warps wait once per stage
                uint64_t* const stage_full = &fulla_bar[sa_i];
                if (!TWO && (a.dbg & 1)) {
                    mbar_arrive_expect_tx_w(&fulla_bar[sa_i], 0);
                } else if (!TWO) {
                    mbar_arrive_expect_tx_w(&fulla_bar[sa_i], srel ? (uint32_t)a.atx + (uint32_t)kys * nbytes : (uint32_t)a.atx);
                    tma_load_4d_w(smem + (size_t)sa_i * a.abytes, main ? &tm_a : &tm_ares, &fulla_bar[sa_i], c * kKC,
                                  halo ? w0 - 1 : w0 + ((main && nkx == 3) ? kx - 1 : 0), h0 - 1, b);
                } else {
                    // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of both
                    if (crank == 0) mbar_arrive_expect_tx_w(&fulla_bar[sa_i], 2 * kABytes);
                    else mbar_arrive_remote_w(&fulla_bar[sa_i], 0);
                    tma_load_4d_2sm_w(smem + (size_t)sa_i * kABytes, main ? &tm_a : &tm_ares, &fulla_bar[sa_i], c * kKC,
                                      w0 + ((main && nkx == 3) ? kx - 1 : 0), h0 - 1, b);
                }
                if (++sa_i == a.nsa) {
                    sa_i = 0;
                    pha ^= 1u;
                }
                for (int ky = 0; ky < kys; ++ky) {
                    const int tap = !main ? 0 : halo ? ky : (nkx == 3 ? ky * 3 + kx : 0);
                    if (srel) {
                        // slot kbox % nstages_b was last used by box kbox - nstages_b: its whole stage must be done
                        while (rel_boxes < kbox - a.nstages_b + 1) release_one();
                        ++kbox;
                    } else {
                        mbar_wait(&emptyb_bar[sb_i], phb ^ 1u);
                    }
                    if (!TWO && (a.dbg & 1)) {
                        if (!srel) mbar_arrive_expect_tx_w(&fullb_bar[sb_i], 0);
                    } else if (!TWO) {
                        uint64_t* const box_full = srel ? stage_full : &fullb_bar[sb_i];
                        if (!srel) mbar_arrive_expect_tx_w(box_full, nbytes);
                        if (a.w_blk) {
                            // the whole (tap, chunk) weight box is one contiguous pre-swizzled block
                            const float* src = main ? a.w_blk + ((size_t)tap * a.nchunks + c) * N * kKC
                                                    : a.wres_blk + (size_t)c * N * kKC;
                            bulk_copy_g2s_w(smem_b + (size_t)sb_i * a.bbox_bytes, src, nbytes, box_full);
                        } else {
                            tma_load_2d_w(smem_b + (size_t)sb_i * a.bbox_bytes, main ? &tm_b : &tm_bres, box_full,
                                          c * kKC, tap * N);
                        }
                    } else {
                        // this CTA stages rows [crank*N/2, (crank+1)*N/2) of the tap's weight matrix
                        if (crank == 0) mbar_arrive_expect_tx_w(&fullb_bar[sb_i], 2 * nbytes);
                        else mbar_arrive_remote_w(&fullb_bar[sb_i], 0);
                        tma_load_2d_2sm_w(smem_b + (size_t)sb_i * a.bbox_bytes, main ? &tm_b : &tm_bres,
                                          &fullb_bar[sb_i], c * kKC, tap * N + crank * (N / 2));
                    }
                    if (++sb_i == a.nstages_b) {
                        sb_i = 0;
                        phb ^= 1u;
                    }
                }
            }
        }
      } else if (!TWO && a.issuers2 && warp >= 1 && warp <= a.nh) {
        // ------------------------------------------------------------ MMA issuers (single-CTA kernel)
        // Issuing a tcgen05.mma costs the issuing warp ~50-70 cycles (descriptor moves into uniform registers,
        // election, predicates), which is as long as an N = 80 MMA executes: one issuing warp cannot keep the
        // tensor pipe busy on the narrow layers and barely does on the wide ones.  The two M = 128 halves of a
        // tile are independent accumulators, so warp 1 issues every MMA of the first half and warp 2 every MMA
        // of the second half: both wait on the same full barriers, both commit to the empty barriers (count 2),
        // each commits its own accumulator slot -- the first half no longer waits for the second one either.
        const int half = warp - 1;
        int sa_i = 0, sb_i = 0;
        uint32_t pha = 0, phb = 0;
        uint32_t slot_uses[kMaxSlots] = {0, 0, 0, 0, 0};
        int titer = 0;
        // high words of the operand descriptors: 8-pixel core groups are 1024 B apart (16-pixel box rows: contiguous) or,
        // in halo mode, one 18-pixel box row apart; weight boxes are dense [N][128 B]
        const uint32_t desc_hi_b = (uint32_t)(umma_smem_desc(0, 0, 1024, UMMA_LAYOUT_SW128) >> 32);
        const uint32_t desc_hi =
            halo ? (uint32_t)(umma_smem_desc(0, 0, (uint32_t)a.abw * (kKC * 4), UMMA_LAYOUT_SW128) >> 32) : desc_hi_b;
        // byte offset of this half's operand inside the activation box for vertical tap / box index ky
        auto a_off = [&](bool main, int ky) -> uint32_t {
            if (halo) {
                const int ty = main ? ky / 3 : 1, tx = main ? ky - 3 * (ky / 3) : 1;
                return (uint32_t)((ty * a.abw + tx + 8 * half) * (kKC * 4));
            }
            return (uint32_t)((((main && nky == 3) ? ky : 1) + 8 * half) * kRowBytes);
        };
        // Every barrier / tcgen05 bookkeeping instruction stalls this warp for 180-260 cycles (tools/pipe_bench.cu),
        // about as long as the four N = 80 MMAs of a weight box execute.  So the NEXT box's full barrier is tested
        // inside the asm that issues the current box's MMAs (the answer is read after the MMAs were issued;
        // mbar_wait only when it was "not yet"), and there is no tcgen05.fence per box: operands written by TMA and
        // observed through the mbarrier need none (the fence after the accumulator-slot wait orders against the
        // epilogue's tcgen05.ld).  SINDDM_TC_PEEK=0 restores wait + fence per box.
        bool b_ready = false;
        bool st_ready = false;   // stage_release: the NEXT stage's (single) full barrier was seen complete by the peek
        const bool srel = a.stage_release && a.peek;
        for (int st = pair_id; st < nsuper; st += npairs, ++titer) {
            // the nh halves of tile t use accumulator slots nh*t .. nh*t + nh-1 (mod nslots): two halves rotate through
            // three slots, three halves through five, so a tile never waits for the epilogue of the tile before it
            const int slot = (a.nh * titer + half) % a.nslots;
            // the slot must have been drained by the epilogue of its previous use (by any half)
            mbar_wait(&tempty_bar[slot], (slot_uses[slot] & 1u) ^ 1u);
            for (int hh = 0; hh < a.nh; ++hh) ++slot_uses[(a.nh * titer + hh) % a.nslots];
            tc_fence_after_sync();
            if (half == 0 && lane == 0) *reinterpret_cast<volatile uint32_t*>(&tmem_slot[1]) = (uint32_t)titer + 1u;
            const uint32_t dacc = tmem_base + (uint32_t)(slot * a.slot_stride);
            for (int it = 0; it < nst; ++it) {
                const bool main = it < nst_main;
                const int cvalid = main ? min(kKC, a.Cin - (it / nkx) * kKC) : min(kKC, a.Cres - (it - nst_main) * kKC);
                const uint32_t nmma = (uint32_t)(cvalid >> 3);  // K = 8 tf32 per instruction, <= 4 per chunk
                const int kys = main ? nky : 1;
                if (!(srel && st_ready)) mbar_wait(&fulla_bar[sa_i], pha);
                const uint32_t sa = smem_u32(smem + (size_t)sa_i * a.abytes);
                if (srel) {
                    // one barrier per stage (halo box + its weight boxes), tested one stage ahead inside the asm that
                    // issues this stage's last MMAs; one commit per stage releases everything the stage read
                    const int sa_n = (sa_i + 1 == a.nsa) ? 0 : sa_i + 1;
                    const uint32_t ph_n = (sa_n == 0) ? pha ^ 1u : pha;
                    uint32_t r = 0;
                    for (int ky = 0; ky < kys; ++ky) {
                        const uint32_t ad = ((sa + a_off(main, ky)) >> 4) & 0x3FFFu;
                        const uint32_t bd = (smem_u32(smem_b + (size_t)sb_i * a.bbox_bytes) >> 4) & 0x3FFFu;
                        if (++sb_i == a.nstages_b) sb_i = 0;
                        const uint32_t acc = (it | ky) != 0 ? 1u : 0u;
                        if (a.dbg & 4) continue;
                        if (ky == kys - 1) r = umma_tf32_ss_x4_test_h2(dacc, ad, bd, desc_hi, desc_hi_b, 2u, a.idesc, acc, nmma, &fulla_bar[sa_n], ph_n);
                        else umma_tf32_ss_x4_h2(dacc, ad, bd, desc_hi, desc_hi_b, 2u, a.idesc, acc, nmma);
                    }
                    umma_commit_elect(&emptya_bar[sa_i]);
                    st_ready = __all_sync(0xffffffffu, r != 0);
                    if (++sa_i == a.nsa) {
                        sa_i = 0;
                        pha ^= 1u;
                    }
                    continue;
                }
                for (int ky = 0; ky < kys; ++ky) {
                    if (!b_ready) mbar_wait(&fullb_bar[sb_i], phb);
                    if (!a.peek) tc_fence_after_sync();
                    // vertical tap = row offset into the halo box; the 1x1 / residual case reads the centre rows
                    const uint32_t ad = ((sa + a_off(main, ky)) >> 4) & 0x3FFFu;
                    const uint32_t bd = (smem_u32(smem_b + (size_t)sb_i * a.bbox_bytes) >> 4) & 0x3FFFu;
                    const int sb_cur = sb_i;
                    if (++sb_i == a.nstages_b) {
                        sb_i = 0;
                        phb ^= 1u;
                    }
                    if (a.peek && !(a.dbg & 4)) {
                        const uint32_t r = umma_tf32_ss_x4_test_h2(dacc, ad, bd, desc_hi, desc_hi_b, 2u, a.idesc,
                                                                   (it | ky) != 0 ? 1u : 0u, nmma, &fullb_bar[sb_i], phb);
                        umma_commit_elect(&emptyb_bar[sb_cur]);
                        b_ready = __all_sync(0xffffffffu, r != 0);
                    } else {
                        if (!(a.dbg & 4)) umma_tf32_ss_x4_h2(dacc, ad, bd, desc_hi, desc_hi_b, 2u, a.idesc, (it | ky) != 0 ? 1u : 0u, nmma);
                        umma_commit_elect(&emptyb_bar[sb_cur]);
                        b_ready = false;
                    }
                }
                umma_commit_elect(&emptya_bar[sa_i]);
                if (++sa_i == a.nsa) {
                    sa_i = 0;
                    pha ^= 1u;
                }
            }
            umma_commit_elect(&tfull_bar[slot]);
        }
      } else if (!TWO && (FL & kFlStream) && a.issuers2 && a.l2pf && warp == 3 && (a.ep.res_add || a.ep.dgelu_z) &&
                 !(a.dbg & 2)) {
        // ------------------------------------------------------------ L2 prefetcher of the streamed epilogue operand
        // The epilogue warps stream the residual / saved pre-activation tile by tile with ONE box in flight per
        // warp (shared memory is spent on the operand rings), i.e. 16 KiB per SM against a DRAM latency of
        // ~2500 cycles: ~1.6 TB/s chip-wide, slower than the MMAs.  This otherwise idle warp requests the boxes of
        // a tile into L2 one tile ahead of the MMA warps (paced by their progress counter, so the ~24 MB in
        // flight chip-wide stay far below the L2 capacity); the epilogue's loads then pay L2 latency only.
        const CUtensorMap* pm = &tm_in;
        const int nch = N / kEpiChunk;
        int t_local = 0;
        for (int st = pair_id; st < nsuper; st += npairs, ++t_local) {
            while ((int)*reinterpret_cast<volatile uint32_t*>(&tmem_slot[1]) < t_local) __nanosleep(256);
            const int tile = st;
            const int tw = tile % a.tiles_w;
            const int th = (tile / a.tiles_w) % a.tiles_h;
            const int b = tile / (a.tiles_w * a.tiles_h);
            for (int idx = lane; idx < nch * (kTileH / 2); idx += 32) {
                const int hq = th * kTileH + (idx % (kTileH / 2)) * 2;
                if (hq < a.H) tma_prefetch_l2_4d(pm, (idx / (kTileH / 2)) * kEpiChunk, tw * kTileW, hq, b);
            }
        }
      } else if ((TWO || !a.issuers2) && warp == 1 && crank == 0) {
        // ------------------------------------------------------------ MMA issuer (leader CTA of a pair)
        int sa_i = 0, sb_i = 0;
        uint32_t pha = 0, phb = 0;
        uint32_t slot_uses[kSlots] = {0, 0, 0};
        int titer = 0;
        // high 32 bits of every operand descriptor: SBO = 1024 B, version 1, 128B swizzle
        const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 0, 1024, UMMA_LAYOUT_SW128) >> 32);
        for (int st = pair_id; st < nsuper; st += npairs, ++titer) {
            const int s0 = (2 * titer) % kSlots, s1 = (2 * titer + 1) % kSlots;
            // the slot must have been drained by the epilogue (of both CTAs) of its previous use
            if (TWO) {
                mbar_wait_cluster(&tempty_bar[s0], (slot_uses[s0] & 1u) ^ 1u);
                mbar_wait_cluster(&tempty_bar[s1], (slot_uses[s1] & 1u) ^ 1u);
            } else {
                mbar_wait(&tempty_bar[s0], (slot_uses[s0] & 1u) ^ 1u);
                mbar_wait(&tempty_bar[s1], (slot_uses[s1] & 1u) ^ 1u);
            }
            ++slot_uses[s0];
            ++slot_uses[s1];
            tc_fence_after_sync();
            const uint32_t d0 = tmem_base + (uint32_t)(s0 * a.slot_stride);
            const uint32_t d1 = tmem_base + (uint32_t)(s1 * a.slot_stride);
            for (int it = 0; it < nst; ++it) {
                const bool main = it < nst_main;
                const int cvalid = main ? min(kKC, a.Cin - (it / nkx) * kKC) : min(kKC, a.Cres - (it - nst_main) * kKC);
                const uint32_t nmma = (uint32_t)(cvalid >> 3);  // K = 8 tf32 per instruction, <= 4 per chunk
                const int kys = main ? nky : 1;
                mbar_wait(&fulla_bar[sa_i], pha);
                const uint32_t sa = smem_u32(smem + (size_t)sa_i * kABytes);
                // vertical tap = row offset into the halo box; the 1x1 / residual case reads the centre rows.
                // K slices advance by 32 B inside the 128B-swizzled rows: +2 in the encoded start address
                auto a_desc = [&](int ky, int half) {
                    const int row0 = ((main && nky == 3) ? ky : 1) + 8 * half;
                    return ((sa + (uint32_t)row0 * kRowBytes) >> 4) & 0x3FFFu;
                };
                auto b_desc = [&](int slot_b) {
                    return (smem_u32(smem_b + (size_t)slot_b * a.bbox_bytes) >> 4) & 0x3FFFu;
                };
                if (TWO || it != nst - 1) {
                    for (int ky = 0; ky < kys; ++ky) {
                        mbar_wait(&fullb_bar[sb_i], phb);
                        tc_fence_after_sync();
                        const uint32_t a0 = a_desc(ky, 0), a1 = a_desc(ky, 1), bb = b_desc(sb_i);
                        const uint32_t acc = (it | ky) != 0 ? 1u : 0u;
                        if (TWO) {
                            umma_tf32_ss_x4_2sm(d0, a0, bb, desc_hi, 2u, a.idesc, acc, nmma);
                            umma_tf32_ss_x4_2sm(d1, a1, bb, desc_hi, 2u, a.idesc, acc, nmma);
                            umma_commit_2sm_elect(&emptyb_bar[sb_i]);
                        } else {
                            if (!(a.dbg & 4)) {
                                umma_tf32_ss_x4(d0, a0, bb, desc_hi, 2u, a.idesc, acc, nmma);
                                umma_tf32_ss_x4(d1, a1, bb, desc_hi, 2u, a.idesc, acc, nmma);
                            }
                            umma_commit_elect(&emptyb_bar[sb_i]);
                        }
                        if (++sb_i == a.nstages_b) {
                            sb_i = 0;
                            phb ^= 1u;
                        }
                    }
                } else {
                    // last stage of the tile: finish the first half on its own, so that the epilogue drains its
                    // accumulator (and returns the TMEM slot the next tile needs) while the second half completes
                    int sb = sb_i;
                    uint32_t ph = phb;
                    for (int ky = 0; ky < kys; ++ky) {
                        mbar_wait(&fullb_bar[sb], ph);
                        tc_fence_after_sync();
                        if (!(a.dbg & 4))
                            umma_tf32_ss_x4(d0, a_desc(ky, 0), b_desc(sb), desc_hi, 2u, a.idesc, (it | ky) != 0 ? 1u : 0u, nmma);
                        if (++sb == a.nstages_b) {
                            sb = 0;
                            ph ^= 1u;
                        }
                    }
                    umma_commit_elect(&tfull_bar[s0]);
                    for (int ky = 0; ky < kys; ++ky) {
                        if (!(a.dbg & 4))
                            umma_tf32_ss_x4(d1, a_desc(ky, 1), b_desc(sb_i), desc_hi, 2u, a.idesc, (it | ky) != 0 ? 1u : 0u, nmma);
                        umma_commit_elect(&emptyb_bar[sb_i]);
                        if (++sb_i == a.nstages_b) {
                            sb_i = 0;
                            phb ^= 1u;
                        }
                    }
                }
                if (TWO) umma_commit_2sm_elect(&emptya_bar[sa_i]);
                else umma_commit_elect(&emptya_bar[sa_i]);
                if (++sa_i == kStagesA) {
                    sa_i = 0;
                    pha ^= 1u;
                }
            }
            if (TWO) {
                umma_commit_2sm_elect(&tfull_bar[s0]);
                umma_commit_2sm_elect(&tfull_bar[s1]);
            } else {
                umma_commit_elect(&tfull_bar[s1]);   // s0 was committed inside the last stage
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
        // ------------------------------------------------------------ epilogue warps (8)
        // warp w reads TMEM lane quarter (w & 3) = 32 pixels = a 16 x 2 pixel strip; the two warps of a quarter
        // split the 16-column chunks (even / odd).  Results never touch the LSU on their way out: a lane owns one
        // pixel, so a direct float4 store would scatter a warp instruction over 32 cache lines (the kernel used to
        // be bound by exactly that).  Instead each warp assembles a [32 px][16 ch] tile in shared memory and one
        // lane hands it to the TMA unit as a (16 ch, 16 w, 2 h) box store, which also clips the image border.
        // The streamed input operand (residual or saved pre-activation) arrives the same way, one chunk ahead.
        const int ew = warp - kEpiWarp0;
        const int quarter = warp & 3;
        const int cgrp = ew >> 2;                 // 0: chunks 0,2,4..  1: chunks 1,3,5..
        const int row = quarter * 32 + lane;      // accumulator row == pixel within the half tile
        const ConvEpilogue& ep = a.ep;
        float* s_fin = reinterpret_cast<float*>(s_bfinal + 4);   // [128][3] partial final-conv sums of group 1
        uint8_t* stg_in = smem_epi + (size_t)ew * kEpiWarpBytes;
        uint64_t* in_bar = &epi_in_bar[ew];
        // 64B swizzle: the 16-byte unit index of a row is XORed with bits 7-8 of the row's byte offset
        const uint32_t lrow = (uint32_t)lane * 64u, swz = (uint32_t)(lane >> 1) & 3u;
        // shared-window addresses of this lane's four 16-byte units in staging tile 0 (tile 1 = + kEpiTileBytes)
        const uint32_t stg_u32 = smem_u32(stg_in);
        const uint32_t so0 = stg_u32 + lrow + ((0u ^ swz) << 4), so1 = stg_u32 + lrow + ((1u ^ swz) << 4);
        const uint32_t so2 = stg_u32 + lrow + ((2u ^ swz) << 4), so3 = stg_u32 + lrow + ((3u ^ swz) << 4);
        uint32_t in_phase = 0;
        int obuf = 0;
        // run-time epilogue flags, each gated by its compile-time flavour bit (see kFl*)
        const bool e_resadd = (FL & kFlResAdd) ? ep.res_add != nullptr : false;
        const bool e_dgelu = (FL & kFlDgelu) ? ep.dgelu_z != nullptr : false;
        const bool e_x3 = (FL & kFlX3) ? ep.w_res3 != nullptr : false;
        const bool e_gelu = (FL & kFlGelu) ? ep.gelu != 0 : false;
        const bool e_pre = (FL & kFlPre) ? ep.out_pre != nullptr : false;
        const bool e_final = (FL & kFlFinal) ? ep.w_final != nullptr : false;
        const bool e_out3 = (FL & kFlOut3) ? ep.out3 != nullptr : false;
        const bool e_colsum = (FL & kFlColsum) ? ep.colsum_part != nullptr : false;
        const bool e_round = (FL & kFlRound) ? ep.round_tf32 != 0 : false;
        const float* gsrc = (FL & kFlStream) ? (e_resadd ? ep.res_add : (e_dgelu ? ep.dgelu_z : nullptr))
                                             : nullptr;   // streamed through TMA
        const float* gsrc2 = (FL & kFlPre2) ? ((e_resadd && e_dgelu) ? ep.dgelu_z : nullptr)
                                            : nullptr;    // rare second stream: plain loads
        uint32_t slot_uses[kMaxSlots] = {0, 0, 0, 0, 0};
        int titer = 0;
        // ep.colsum_part: column sums of everything this warp stores to `out` (its column share, its pixel quarter),
        // kept in lanes 0-15 (one column each) per chunk of the warp, written once at the end of the kernel
        float csum_acc[kMaxN / (2 * kEpiChunk)];
#pragma unroll
        for (int i = 0; i < kMaxN / (2 * kEpiChunk); ++i) csum_acc[i] = 0.f;
        for (int st = pair_id; st < nsuper; st += npairs, ++titer) {
            const int tile = st * CS + crank;
            const int tw = tile % a.tiles_w;
            const int th = (tile / a.tiles_w) % a.tiles_h;
            const int b = tile / (a.tiles_w * a.tiles_h);
            const bool live = (tile < a.ntiles) && !(a.dbg & 2);
#pragma unroll 1
            for (int half = 0; half < a.nh; ++half) {
                const int slot = (a.nh * titer + half) % a.nslots;
                // accumulator row -> pixel.  Stacked halves: half = 8 image rows x 16 px, this warp's strip = 2 rows x 16 px;
                // halo mode: half = 16 rows x 8 px (columns 8*half ..), this warp's strip = 4 rows x 8 px
                const int h = halo ? th * kTileH + (row >> 3) : th * kTileH + half * 8 + row / kTileW;
                const int w = halo ? tw * 8 * a.nh + half * 8 + (row & 7) : tw * kTileW + row % kTileW;
                const int hq = halo ? th * kTileH + quarter * 4 : th * kTileH + half * 8 + quarter * 2;
                const int w0 = halo ? tw * 8 * a.nh + half * 8 : tw * kTileW;   // this warp's strip
                const bool valid = live && (h < a.H) && (w < a.W);
                const size_t pix = ((size_t)b * a.H + h) * a.W + w;
                const unsigned vmask = e_colsum ? __ballot_sync(0xffffffffu, valid) : 0u;

                float x3v[3] = {0.f, 0.f, 0.f};
                if ((FL & kFlX3) && ep.x3 && valid) {
                    x3v[0] = ep.x3[pix * 3 + 0];
                    x3v[1] = ep.x3[pix * 3 + 1];
                    x3v[2] = ep.x3[pix * 3 + 2];
                }
                float fin[3] = {0.f, 0.f, 0.f};

                // the first chunk of the streamed operand is requested before the accumulator is even ready
                int cc = cgrp * kEpiChunk;
                if (gsrc && live && cc < N && lane == 0) {
                    mbar_arrive_expect_tx(in_bar, kEpiTileBytes);
                    tma_load_4d(stg_in, &tm_in, in_bar, cc, w0, hq, b);
                }

                mbar_wait(&tfull_bar[slot], slot_uses[slot] & 1u);
                ++slot_uses[slot];
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * a.slot_stride);

                // column sums of a staged [32 px][16 ch] tile: lanes 0-15 walk the even rows of "their" column, lanes 16-31
                // the odd rows (one 128-byte wavefront per step: conflict free under the 64B swizzle); pixels outside
                // the image are skipped.  0.08 instructions per value, against 5 for a shuffle reduction of registers.
                auto tile_colsum = [&](const uint32_t tile_u32, float& acc) {
                    // row r = 2i + odd: (r >> 1) & 3 = i & 3 for both parities, so the swizzled column offset only depends
                    // on i & 3 -> four lane-dependent base addresses, every load is base + immediate
                    const uint32_t j = (uint32_t)lane & 15u, odd = (uint32_t)lane >> 4;
                    const uint32_t base = tile_u32 + odd * 64u + ((j & 3u) << 2);
                    const uint32_t q = j >> 2;
                    const uint32_t b0 = base + ((q ^ 0u) << 4), b1 = base + ((q ^ 1u) << 4), b2 = base + ((q ^ 2u) << 4),
                                   b3 = base + ((q ^ 3u) << 4);
                    float x[16];
#pragma unroll
                    for (uint32_t i = 0; i < 16; i += 4) {
                        x[i + 0] = lds_f32_plain(b0 + (i + 0) * 128u);
                        x[i + 1] = lds_f32_plain(b1 + (i + 1) * 128u);
                        x[i + 2] = lds_f32_plain(b2 + (i + 2) * 128u);
                        x[i + 3] = lds_f32_plain(b3 + (i + 3) * 128u);
                    }
                    if (vmask != 0xffffffffu) {
#pragma unroll
                        for (uint32_t i = 0; i < 16; ++i)
                            if (!((vmask >> (2u * i + odd)) & 1u)) x[i] = 0.f;
                    }
                    // fixed association: four independent chains of four, then a tree
                    const float s0 = (x[0] + x[4]) + (x[8] + x[12]), s1 = (x[1] + x[5]) + (x[9] + x[13]);
                    const float s2 = (x[2] + x[6]) + (x[10] + x[14]), s3 = (x[3] + x[7]) + (x[11] + x[15]);
                    float s = (s0 + s1) + (s2 + s3);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    acc += s;
                };

                auto store_tile = [&](const CUtensorMap* map, const int cc, const float (&v)[16], float* csum = nullptr) {
                    // with a streamed operand tile 1 is the only store buffer, otherwise the two tiles alternate
                    uint8_t* buf = stg_in + (gsrc ? 1 : obuf) * kEpiTileBytes;
                    obuf ^= 1;
                    if (lane == 0) {   // the previous store that read this buffer must have drained it
                        if (gsrc) bulk_wait_group_read<0>();
                        else bulk_wait_group_read<1>();
                    }
                    __syncwarp();
                    const uint32_t bo = (uint32_t)(buf - stg_in);
                    sts_f4(so0 + bo, v[0], v[1], v[2], v[3]);
                    sts_f4(so1 + bo, v[4], v[5], v[6], v[7]);
                    sts_f4(so2 + bo, v[8], v[9], v[10], v[11]);
                    sts_f4(so3 + bo, v[12], v[13], v[14], v[15]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && !(a.dbg & 8)) {
                        tma_store_4d(map, buf, cc, w0, hq, b);
                        bulk_commit_group();
                    }
                    if (csum) tile_colsum(stg_u32 + bo, *csum);   // the tile stays intact until its next wait_group_read
                };

                // TMEM -> register loads are double buffered: chunk i+1 is in flight while chunk i is processed
                auto process = [&](const int cc, const uint32_t (&raw)[16], float& csum) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                    float4 cur[4], pre2[4];
                    if (gsrc && live) {
                        mbar_wait(in_bar, in_phase);
                        in_phase ^= 1u;
                        cur[0] = lds_f4(so0);
                        cur[1] = lds_f4(so1);
                        cur[2] = lds_f4(so2);
                        cur[3] = lds_f4(so3);
                        // The next chunk's TMA load overwrites this tile (async proxy) right after these generic-proxy
                        // reads.  Without a proxy fence between the two the hardware is free to perform the write first:
                        // a few pixels then got the NEXT chunk's operand -- rare (<= 2 % of launches), timing dependent,
                        // found by tools/determinism_stress.py (out = plain + res[c+32] exactly).  So: reads are
                        // ld.volatile (ptxas may neither sink nor repeat them), every lane fences generic -> async, and
                        // the warp vote on a predicate computed from the loaded registers (also the warp barrier) lets
                        // lane 0 issue the load only after every lane's data has returned.
                        fence_proxy_async_smem();
                        {
                            const uint32_t chk = __float_as_uint(cur[0].x) ^ __float_as_uint(cur[1].y) ^
                                                 __float_as_uint(cur[2].z) ^ __float_as_uint(cur[3].w);
                            const unsigned landed = __ballot_sync(0xffffffffu, chk != 0u);
                            asm volatile("" ::"r"(landed) : "memory");
                        }
                        if (lane == 0 && cc + 2 * kEpiChunk < N) {   // next chunk of this warp
                            mbar_arrive_expect_tx(in_bar, kEpiTileBytes);
                            tma_load_4d(stg_in, &tm_in, in_bar, cc + 2 * kEpiChunk, w0, hq, b);
                        }
                    }
                    if (gsrc2) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            pre2[q] = valid ? __ldg(reinterpret_cast<const float4*>(gsrc2 + pix * N + cc) + q)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    {   // s_bias is 16-byte aligned and cc a multiple of 16: four broadcast LDS.128 instead of sixteen LDS.32
                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + cc);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bb = b4[q];
                            v[4 * q + 0] += bb.x;
                            v[4 * q + 1] += bb.y;
                            v[4 * q + 2] += bb.z;
                            v[4 * q + 3] += bb.w;
                        }
                    }
                    if (e_x3) {
                        // the 16 columns' [3] weights are 48 consecutive floats (16-byte aligned: s_wres3 sits 976 B into
                        // the 1024-aligned tail, cc * 12 B is a multiple of 192): twelve broadcast LDS.128 through the
                        // shared window instead of 48 scalar loads through generic addresses
                        const uint32_t wb = smem_u32(s_wres3) + (uint32_t)cc * 12u;
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {          // 4 columns = 12 floats = 3 x float4 per step
                            const float4 w0 = lds_f4_plain(wb + g4 * 48u), w1 = lds_f4_plain(wb + g4 * 48u + 16u),
                                         w2 = lds_f4_plain(wb + g4 * 48u + 32u);
                            const float wr[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = g4 * 4 + jj;
                                v[j] = fmaf(x3v[2], wr[3 * jj + 2], fmaf(x3v[1], wr[3 * jj + 1], fmaf(x3v[0], wr[3 * jj], v[j])));
                            }
                        }
                    }
                    if (e_resadd) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            v[4 * q + 0] += cur[q].x;
                            v[4 * q + 1] += cur[q].y;
                            v[4 * q + 2] += cur[q].z;
                            v[4 * q + 3] += cur[q].w;
                        }
                    }
                    if (e_gelu && ep.pre_grad && e_pre) {
                        // save gelu'(z) for the backward pass instead of z: cdf and pdf are both at hand here
                        float g[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float cdf, pdf;
                            phi_cdf_pdf(v[j], &cdf, &pdf);
                            g[j] = fmaf(v[j], pdf, cdf);
                            v[j] *= cdf;
                        }
                        if (live) store_tile(&tm_pre, cc, g);
                    } else {
                        if (e_pre && live) store_tile(&tm_pre, cc, v);
                        if (e_gelu) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
                        }
                    }
                    if (e_dgelu) {
                        const float4* z = gsrc2 ? pre2 : cur;
                        if (ep.pre_grad) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[4 * q + 0] *= z[q].x;
                                v[4 * q + 1] *= z[q].y;
                                v[4 * q + 2] *= z[q].z;
                                v[4 * q + 3] *= z[q].w;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[4 * q + 0] *= gelu_grad_fast(z[q].x);
                                v[4 * q + 1] *= gelu_grad_fast(z[q].y);
                                v[4 * q + 2] *= gelu_grad_fast(z[q].z);
                                v[4 * q + 3] *= gelu_grad_fast(z[q].w);
                            }
                        }
                    }
                    if (e_final) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            fin[0] = fmaf(v[j], s_wfinal[0 * N + cc + j], fin[0]);
                            fin[1] = fmaf(v[j], s_wfinal[1 * N + cc + j], fin[1]);
                            fin[2] = fmaf(v[j], s_wfinal[2 * N + cc + j], fin[2]);
                        }
                    }
                    if (e_out3 && cc == 0 && valid) {
                        // a lane owns one pixel: the warp's 32 x 12 bytes are contiguous NHWC memory
                        float* o3 = ep.out3 + pix * 3;
                        o3[0] = v[0];
                        o3[1] = v[1];
                        o3[2] = v[2];
                    }
                    if (ep.out && live) {
                        if (e_round) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
                        }
                        store_tile(&tm_out, cc, v, e_colsum ? &csum : nullptr);
                    }
                };
                {
                    // Drain first, compute later: the whole column share of this warp (<= 5 chunks) moves to
                    // registers and the TMEM slot goes straight back to the MMA warp, so the next tile's main loop
                    // overlaps the epilogue arithmetic of BOTH halves instead of waiting for the first one.
                    uint32_t r[kMaxN / (2 * kEpiChunk)][16];
#pragma unroll
                    for (int i = 0; i < kMaxN / (2 * kEpiChunk); ++i)
                        if (cc + i * 2 * kEpiChunk < N) tmem_ld16_issue(taddr + cc + i * 2 * kEpiChunk, r[i]);
#pragma unroll
                    for (int i = 0; i < kMaxN / (2 * kEpiChunk); ++i)
                        if (cc + i * 2 * kEpiChunk < N) tmem_ld16_wait(r[i]);
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        if (TWO) mbar_arrive_remote(&tempty_bar[slot], 0);   // the leader's MMA warp waits for both CTAs
                        else mbar_arrive(&tempty_bar[slot]);
                    }
#pragma unroll
                    for (int i = 0; i < kMaxN / (2 * kEpiChunk); ++i)
                        if (cc + i * 2 * kEpiChunk < N) process(cc + i * 2 * kEpiChunk, r[i], csum_acc[i]);
                }

                if (e_final) {
                    // the two column groups of a pixel combine their partial 3-channel sums through smem
                    if (cgrp == 1) {
                        s_fin[row * 3 + 0] = fin[0];
                        s_fin[row * 3 + 1] = fin[1];
                        s_fin[row * 3 + 2] = fin[2];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (cgrp == 0 && valid) {
                        const size_t plane = (size_t)a.H * a.W;
                        float* o = ep.out_final + (size_t)b * 3 * plane + (size_t)h * a.W + w;
                        o[0] = fin[0] + s_fin[row * 3 + 0] + s_bfinal[0];
                        o[plane] = fin[1] + s_fin[row * 3 + 1] + s_bfinal[1];
                        o[2 * plane] = fin[2] + s_fin[row * 3 + 2] + s_bfinal[2];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
        if (lane == 0) bulk_wait_group_read<0>();   // the staging tiles are read until the last store has drained
        __syncwarp();
        if (e_colsum && lane < 16) {
            float* dst = ep.colsum_part + ((size_t)blockIdx.x * 4 + quarter) * N + cgrp * kEpiChunk + lane;
#pragma unroll
            for (int i = 0; i < kMaxN / (2 * kEpiChunk); ++i)
                if (cgrp * kEpiChunk + i * 2 * kEpiChunk < N) dst[i * 2 * kEpiChunk] = csum_acc[i];
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before_sync();
    __syncthreads();
    if (TWO) cluster_sync_all();   // neither CTA leaves while the other may still touch its smem / TMEM
    if (warp == 1) {
        tc_fence_after_sync();
        if (TWO) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// The instantiations, most specific first (tc_conv_launch takes the first whose mask covers the layer).  The masks are
// the epilogues SinDDMNet's plan uses (net.cu); anything else -- operator tests, future layers -- runs the generic one.
using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const CUtensorMap, const CUtensorMap, const KernelArgs);
struct Flavour {
    int mask;
    KernelFn fn;
    const char* what;
};
constexpr int kFlC1 = kFlGelu | kFlPre | kFlRound;                       // net[0]: GELU (+ saved gelu' when training)
constexpr int kFlRes = kFlStream | kFlResAdd;                            // l3.net[2]: + identity residual (streamed)
constexpr int kFlD2 = kFlStream | kFlDgelu | kFlColsum | kFlRound;       // data gradient through GELU + column sums
const Flavour kFlavours[] = {
    {0, tc_conv_kernel<false, 0>, "plain (bias only): residual-slice layers, first-conv / residual data gradients"},
    {kFlOut3, tc_conv_kernel<false, kFlOut3>, "3-channel data gradient (N padded to 16)"},
    {kFlX3, tc_conv_kernel<false, kFlX3>, "l1.net[2]: 3-channel residual conv in the epilogue"},
    {kFlFinal, tc_conv_kernel<false, kFlFinal>, "l4.net[2]: fused final 1x1 conv"},
    {kFlRes, tc_conv_kernel<false, kFlRes>, "l3.net[2]: streamed identity residual"},
    {kFlC1, tc_conv_kernel<false, kFlC1>, "net[0]: GELU, saved gelu', tf32 rounding"},
    {kFlD2, tc_conv_kernel<false, kFlD2>, "net[2] data gradient: streamed gelu', column sums, tf32 rounding"},
    {kFlAll, tc_conv_kernel<false, kFlAll>, "generic"},
};
constexpr int kNumFlavours = (int)(sizeof(kFlavours) / sizeof(kFlavours[0]));

int flavour_need(const ConvEpilogue& ep) {
    int need = 0;
    if (ep.res_add || ep.dgelu_z) need |= kFlStream;
    if (ep.res_add && ep.dgelu_z) need |= kFlPre2;
    if (ep.res_add) need |= kFlResAdd;
    if (ep.dgelu_z) need |= kFlDgelu;
    if (ep.x3 || ep.w_res3) need |= kFlX3;
    if (ep.gelu) need |= kFlGelu;
    if (ep.out_pre) need |= kFlPre;
    if (ep.w_final) need |= kFlFinal;
    if (ep.out3) need |= kFlOut3;
    if (ep.colsum_part) need |= kFlColsum;
    if (ep.round_tf32) need |= kFlRound;
    return need;
}

}  // namespace

// Flavour mask of the instantiation tc_conv_launch picks for this epilogue (kFlAll = the generic one).
int tc_conv_flavour_mask(const ConvEpilogue& ep) {
    const int need = flavour_need(ep);
    for (int i = 0; i < kNumFlavours; ++i)
        if ((need & ~kFlavours[i].mask) == 0) return kFlavours[i].mask;
    return kFlAll;
}

static int two_sm_setting() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SINDDM_TC_2SM");
        v = e ? (atoi(e) != 0) : 0;   // CTA pairs are correct but currently slower: off by default
    }
    return v;
}

int tc_conv_colsum_rows(const TcConvOp& op) { return op.grid * 4; }

bool tc_conv_supported(const ConvProblem& p) {
    if (p.Cin < 8 || p.Cin % 8 != 0) return false;
    if (p.in_res && (p.Cres < 8 || p.Cres % 8 != 0)) return false;
    if (p.N % 16 != 0 || p.N < 16 || p.N > kMaxN) return false;
    if (p.ntaps != 9 && p.ntaps != 1) return false;
    return true;
}

int tc_conv_prepare(const ConvProblem& p, TcConvOp* op) {
    SINDDM_REQUIRE(tc_conv_supported(p), "tc_conv: unsupported shape Cin=%d Cres=%d N=%d ntaps=%d", p.Cin, p.Cres,
                   p.N, p.ntaps);
    SINDDM_REQUIRE(device_info().initialized, "sinddm_init() has not been called");
    SINDDM_REQUIRE(!(p.ep.w_res3 && p.ep.w_final), "tc_conv: w_res3 and w_final cannot be fused into one layer");
    SINDDM_REQUIRE(!((p.ep.res_add || p.ep.dgelu_z) && p.ep.out_pre && p.ep.out),
                   "tc_conv: a streamed epilogue operand allows one result tensor");
    op->p = p;
    // CTA pairs (SINDDM_TC_2SM=0 disables): each CTA stages N/2 weight rows, which must be whole 8-row atoms
    op->cs = (two_sm_setting() && p.N % 16 == 0 && !p.w_blocked) ? 2 : 1;
    const int cs = op->cs;
    // halo mode: 3x3 layers of the single-CTA kernel (SINDDM_TC_HALO=0: the three-boxes-per-chunk walk)
    {
        const char* e = getenv("SINDDM_TC_HALO");
        const char* ei = getenv("SINDDM_TC_ISSUERS");      // the halo walk lives in the two-issuer loop
        op->halo = (p.ntaps == 9 && cs == 1 && !(e && atoi(e) == 0) && !(ei && atoi(ei) == 1)) ? 1 : 0;
    }
    // three halves need three accumulator slots of N (rounded to 32) columns each -- always true for N <= 160
    {
        const char* e = getenv("SINDDM_TC_HALVES");
        // three halves only where FIVE accumulator slots fit (N <= 96): with fewer the next tile's main loop would wait
        // for the epilogue of this one (measured: N = 160 with three fixed slots is 20 % slower)
        // (A/B on one box, 32x186x248: N = 80 layers -3 .. -11 %, the N = 16 data gradient -23 %; the layer whose epilogue
        //  adds the 3-channel residual per pixel, l1.net[2], +10 % -> keeps two halves)
        op->nh = (op->halo && !(e && atoi(e) == 2) && !p.ep.x3 &&
                  kMaxSlots * (int)align_up((size_t)p.N, 32) <= kTmemCols) ? 3 : 2;
    }
    const int abw = op->halo ? 8 * op->nh + 2 : kTileW;       // activation box width in pixels
    op->abw = abw;
    op->nsa = op->halo ? kHaloStagesA : kStagesA;
    op->atx = op->halo ? kBoxH * abw * kKC * 4 : kABytes;
    op->abytes = (int)align_up((size_t)op->atx, 1024);
    SINDDM_TRY(make_tmap_nhwc(&op->tm_a, p.in, p.B, p.H, p.W, p.Cin, kKC, abw, kBoxH, CU_TENSOR_MAP_SWIZZLE_128B));
    if (p.w_blocked) memset(&op->tm_b, 0, sizeof(op->tm_b));
    else SINDDM_TRY(make_tmap_2d(&op->tm_b, p.w, p.Cin, p.ntaps * p.N, kKC, p.N / cs, CU_TENSOR_MAP_SWIZZLE_128B));
    if (p.in_res) {
        SINDDM_TRY(make_tmap_nhwc(&op->tm_ares, p.in_res, p.B, p.H, p.W, p.Cres, kKC, abw, kBoxH,
                                  CU_TENSOR_MAP_SWIZZLE_128B));
        if (p.w_blocked) memset(&op->tm_bres, 0, sizeof(op->tm_bres));
        else SINDDM_TRY(make_tmap_2d(&op->tm_bres, p.w_res, p.Cres, p.N, kKC, p.N / cs, CU_TENSOR_MAP_SWIZZLE_128B));
    } else {
        op->tm_ares = op->tm_a;
        op->tm_bres = op->tm_b;
    }
    // epilogue tiles: (16 ch, 16 w, 2 h) boxes, 64B swizzle; unused maps alias the activation map
    {
        const float* in_stream = p.ep.res_add ? p.ep.res_add : p.ep.dgelu_z;
        const float* ptrs[3] = {p.ep.out, p.ep.out_pre, in_stream};
        CUtensorMap* maps[3] = {&op->tm_out, &op->tm_pre, &op->tm_in};
        for (int i = 0; i < 3; ++i) {
            if (ptrs[i])
                SINDDM_TRY(make_tmap_nhwc(maps[i], ptrs[i], p.B, p.H, p.W, p.N, kEpiChunk, op->halo ? 8 : kTileW,
                                          op->halo ? 4 : 2, CU_TENSOR_MAP_SWIZZLE_64B));
            else
                *maps[i] = op->tm_a;
        }
    }
    // weight ring: as many boxes as fit beside the 3 activation slots and the epilogue tiles (at most kMaxStagesB)
    op->stage_bytes = (int)align_up((size_t)(p.N / cs) * kKC * 4, 1024);
    const int budget =
        device_info().max_smem_optin - 1024 /*alignment slack*/ - kTailBytes - op->nsa * op->abytes - kEpiBytes;
    int nst = budget / op->stage_bytes;
    if (nst > kMaxStagesB) nst = kMaxStagesB;
    {
        const char* e = getenv("SINDDM_TC_MAX_BOXES");   // diagnostic: shallower weight ring
        if (e && atoi(e) >= 3 && atoi(e) < nst) nst = atoi(e);
    }
    SINDDM_REQUIRE(nst >= 3, "tc_conv: not enough shared memory for the weight ring");
    op->nstages = nst;
    op->smem_bytes = op->nsa * op->abytes + nst * op->stage_bytes + kEpiBytes + kTailBytes + 1024;
    op->tiles_w = ceil_div(p.W, op->halo ? 8 * op->nh : kTileW);
    op->tiles_h = ceil_div(p.H, kTileH);
    op->ntiles = op->tiles_w * op->tiles_h * p.B;
    const int nsuper = ceil_div(op->ntiles, cs);
    int npairs = device_info().num_sms / cs;
    if (npairs > nsuper) npairs = nsuper;
    op->grid = npairs * cs;
    // switches: read here (plan build / single-operator call), never per launch
    auto env_int = [](const char* name, int dflt) {
        const char* e = getenv(name);
        return e ? atoi(e) : dflt;
    };
    op->sw_peek = env_int("SINDDM_TC_PEEK", 1) != 0;
    op->sw_l2pf = env_int("SINDDM_TC_L2PF", 0) != 0 && !op->halo;   // measured neutral (profiles/): off unless asked for
    op->sw_issuers2 = env_int("SINDDM_TC_ISSUERS", 2) != 1;
    op->sw_dbg = env_int("SINDDM_TC_DEBUG", 0);                 // diagnostic runs only: results are wrong when set
    // stage-granular release needs the ring to hold two whole stages plus the boxes in flight behind them
    const int nky = p.ntaps == 9 ? (op->halo ? 9 : 3) : 1;
    op->sw_stage_release = env_int("SINDDM_TC_STAGE_RELEASE", 1) != 0 && cs == 1 && op->sw_issuers2 && op->sw_peek &&
                           nst >= 2 * nky + 2;
    return SINDDM_OK;
}

int tc_conv_launch(const TcConvOp& op, cudaStream_t stream) {
    // opt-in shared memory size: set once per process (function-local static: thread-safe initialisation)
    static const cudaError_t smem_set = []() {
        for (int i = 0; i < kNumFlavours; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kFlavours[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 device_info().max_smem_optin);
            if (e != cudaSuccess) return e;
        }
        return cudaFuncSetAttribute(tc_conv_kernel<true, kFlAll>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    device_info().max_smem_optin);
    }();
    SINDDM_CUDA_OK(smem_set);
    // the smallest instantiation whose compile-time flavour mask covers this launch's epilogue (SINDDM_TC_FLAVOURS=0:
    // always the generic one -- A/B switch and the reference the specialised ones are tested against)
    static const bool use_flavours = []() {
        const char* e = getenv("SINDDM_TC_FLAVOURS");
        return !(e && atoi(e) == 0);
    }();
    KernelFn kernel = kFlavours[kNumFlavours - 1].fn;
    if (use_flavours) {
        const int need = flavour_need(op.p.ep);
        for (int i = 0; i < kNumFlavours; ++i)
            if ((need & ~kFlavours[i].mask) == 0) {
                kernel = kFlavours[i].fn;
                break;
            }
    }
    const ConvProblem& p = op.p;
    KernelArgs a;
    a.B = p.B;
    a.H = p.H;
    a.W = p.W;
    a.Cin = p.Cin;
    a.ntaps = p.ntaps;
    a.nchunks = ceil_div(p.Cin, kKC);
    a.Cres = p.in_res ? p.Cres : 0;
    a.nchunks_res = p.in_res ? ceil_div(p.Cres, kKC) : 0;
    a.N = p.N;
    a.tiles_w = op.tiles_w;
    a.tiles_h = op.tiles_h;
    a.ntiles = op.ntiles;
    a.nstages_b = op.nstages;
    a.bbox_bytes = op.stage_bytes;
    a.slot_stride = (int)align_up((size_t)p.N, 32);
    a.idesc = umma_idesc_tf32(op.cs == 2 ? 256 : 128, p.N, 0, 0);
    a.w_blk = p.w_blocked ? p.w : nullptr;
    a.wres_blk = p.w_blocked ? p.w_res : nullptr;
    a.peek = op.sw_peek;
    a.l2pf = op.sw_l2pf;
    a.issuers2 = op.sw_issuers2;
    a.dbg = op.sw_dbg;
    a.stage_release = op.sw_stage_release;
    a.halo = op.halo;
    a.nh = op.nh;
    a.nslots = op.nh == 3 ? kMaxSlots : kSlots;
    a.abw = op.abw;
    a.nsa = op.nsa;
    a.abytes = op.abytes;
    a.atx = op.atx;
    a.ep = p.ep;
    // algorithmic work: real pixels x N x (taps*Cin + Cres) MACs
    prof_begin(stream, 0, 2.0 * (double)p.B * p.H * p.W * p.N * ((double)p.ntaps * p.Cin + a.Cres));
    if (op.cs == 2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(op.grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = op.smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t lerr = cudaLaunchKernelEx(&cfg, tc_conv_kernel<true, kFlAll>, op.tm_a, op.tm_ares, op.tm_b, op.tm_bres,
                                              op.tm_out, op.tm_pre, op.tm_in, a);
        if (lerr != cudaSuccess) {
            prof_end(stream);
            set_error("tc_conv (cta pair) launch failed: %s", cudaGetErrorString(lerr));
            return SINDDM_ERR_CUDA;
        }
    } else {
        (void)launch_pdl(kernel, dim3(op.grid), dim3(kThreads), (size_t)op.smem_bytes, stream, op.tm_a,
                         op.tm_ares, op.tm_b, op.tm_bres, op.tm_out, op.tm_pre, op.tm_in, a);
    }
    prof_end(stream);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
