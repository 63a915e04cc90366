// 3x3 / 1x1 convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the cuDNN implicit-GEMM calls behind nn.Conv2d in SinDDMConvBlock.net / res_conv
// (reference SinDDM/models.py:62-67,79-80) and, with data-gradient-packed weights, their backward.
//
//   GEMM view      M = pixels (tile = 8 rows x 16 cols = 128 pixels of one image)
//                  N = output channels (one UMMA N, 16..160)
//                  K = taps x input channels, walked as (32-channel chunk, tap) steps;
//                      an optional 1x1 residual conv rides along as extra K steps from a second input.
//   A operand      NHWC activations.  Each K step is ONE 4-D TMA box (32 ch, 16 w, 8 h, 1 b) whose
//                  start coordinate is shifted by the tap offset; the halo and the image border are
//                  out-of-bounds coordinates that TMA zero-fills, i.e. exact zero padding, no im2col.
//   B operand      packed weights [tap][N][Cin] (K-major), one 2-D TMA box (32 ch, N rows) per step.
//   both land in 128B-swizzled K-major smem tiles that the UMMA descriptors consume directly.
//   accumulator    fp32 in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps
//                  the main loop of tile i+1.
//   roles          warp 0: TMA producer (one lane) | warp 1: TMEM alloc + MMA issue (one lane)
//                  warps 2-5: epilogue (TMEM -> registers -> bias/residual/GELU/... -> global)
//   grid           persistent, min(#tiles, #SMs) CTAs, static round-robin over tiles.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kTileH = 8;
constexpr int kTileW = 16;
constexpr int kBM = kTileH * kTileW;      // 128 rows = UMMA M
constexpr int kKC = 32;                   // channels per K step (32 fp32 = one 128B swizzle row)
constexpr int kABytes = kBM * kKC * 4;    // 16 KiB
constexpr int kMaxN = 160;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;           // TMEM column offset between the two accumulator buffers

struct KernelArgs {
    int B, H, W;
    int Cin, ntaps, nchunks;   // main K walk: nchunks x ntaps steps
    int Cres, nchunks_res;     // residual K walk: nchunks_res steps (center tap)
    int N;
    int tiles_w, tiles_h, ntiles;
    int nstages, stage_bytes;
    int cs;                    // cluster size: CTAs of a cluster work on cs different tiles and share every
    int b_rows;                // weight stage -- each loads N/cs rows of it and multicasts them to all
    int nsuper;                // ceil(ntiles / cs)
    uint32_t idesc;
    ConvEpilogue ep;
};

// smem carve-up (after the 1024-aligned stage ring):
//   uint64 full[nstages], empty[nstages], tmem_full[2], tmem_empty[2]; uint32 tmem_slot;
//   float bias[kMaxN], wres3[kMaxN*3], wfinal[3*kMaxN], bfinal[4]
constexpr int kTailBytes = 16 * 8 * 2 + 4 * 8 + 16 + (kMaxN + kMaxN * 3 + 3 * kMaxN + 4) * 4;

__global__ void __launch_bounds__(kThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_ares,
               const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_bres,
               const KernelArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    uint8_t* tail = smem + (size_t)a.nstages * a.stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 16;
    uint64_t* tfull_bar = empty_bar + 16;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_wres3 = s_bias + kMaxN;
    float* s_wfinal = s_wres3 + kMaxN * 3;
    float* s_bfinal = s_wfinal + 3 * kMaxN;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int N = a.N;

    // ---------------------------------------------------------------- one-time setup
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        if (a.nchunks_res > 0) {
            tma_prefetch_desc(&tm_ares);
            tma_prefetch_desc(&tm_bres);
        }
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], a.cs);   // every CTA of the cluster releases the stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        for (int i = t; i < N; i += kEpiThreads) s_bias[i] = a.ep.bias ? a.ep.bias[i] : 0.f;
        if (a.ep.w_res3)
            for (int i = t; i < N * 3; i += kEpiThreads) s_wres3[i] = a.ep.w_res3[i];
        if (a.ep.w_final) {
            for (int i = t; i < N * 3; i += kEpiThreads) s_wfinal[i] = a.ep.w_final[i];
            if (t < 3) s_bfinal[t] = a.ep.b_final ? a.ep.b_final[t] : 0.f;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (a.cs > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int crank = a.cs > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / a.cs;
    const int nclusters = gridDim.x / a.cs;
    const uint16_t cmask = (uint16_t)((1u << a.cs) - 1u);

    const int nk_main = a.nchunks * a.ntaps;
    const int nk = nk_main + a.nchunks_res;
    const uint32_t b_bytes = (uint32_t)N * kKC * 4;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int st = cluster_id; st < a.nsuper; st += nclusters) {
                // tiles past the end (ragged last cluster) run the same pipeline on an all-zero tile:
                // image index B is out of bounds for TMA, which zero-fills the box
                const int tile = st * a.cs + crank;
                const int tw = tile % a.tiles_w;
                const int th = (tile / a.tiles_w) % a.tiles_h;
                const int b = tile / (a.tiles_w * a.tiles_h);
                const int h0 = th * kTileH, w0 = tw * kTileW;
                for (int it = 0; it < nk; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa = smem + (size_t)stage * a.stage_bytes;
                    uint8_t* sb = sa + kABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], kABytes + b_bytes);
                    if (it < nk_main) {
                        const int c = it / a.ntaps;
                        const int tap = it - c * a.ntaps;
                        int dy = 0, dx = 0;
                        if (a.ntaps == 9) {
                            dy = tap / 3 - 1;
                            dx = tap % 3 - 1;
                        }
                        tma_load_4d(sa, &tm_a, &full_bar[stage], c * kKC, w0 + dx, h0 + dy, b);
                        if (a.cs == 1)
                            tma_load_2d(sb, &tm_b, &full_bar[stage], c * kKC, tap * N);
                        else
                            tma_load_2d_mc(sb + (size_t)crank * a.b_rows * (kKC * 4), &tm_b, &full_bar[stage], c * kKC,
                                           tap * N + crank * a.b_rows, cmask);
                    } else {
                        const int c = it - nk_main;
                        tma_load_4d(sa, &tm_ares, &full_bar[stage], c * kKC, w0, h0, b);
                        if (a.cs == 1)
                            tma_load_2d(sb, &tm_bres, &full_bar[stage], c * kKC, 0);
                        else
                            tma_load_2d_mc(sb + (size_t)crank * a.b_rows * (kKC * 4), &tm_bres, &full_bar[stage],
                                           c * kKC, crank * a.b_rows, cmask);
                    }
                    if (++stage == a.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int titer = 0;
            for (int st = cluster_id; st < a.nsuper; st += nclusters, ++titer) {
                const int buf = titer & 1;
                mbar_wait(&tempty_bar[buf], (((uint32_t)titer >> 1) & 1u) ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccStride;
                for (int it = 0; it < nk; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * a.stage_bytes);
                    const uint32_t sb = sa + kABytes;
                    int cvalid;
                    if (it < nk_main) {
                        const int c = it / a.ntaps;
                        cvalid = min(kKC, a.Cin - c * kKC);
                    } else {
                        cvalid = min(kKC, a.Cres - (it - nk_main) * kKC);
                    }
                    const int nmma = cvalid >> 3;  // K = 8 tf32 per instruction
                    for (int k = 0; k < nmma; ++k) {
                        const uint64_t da = umma_smem_desc(sa + k * 32, 0, 1024, UMMA_LAYOUT_SW128);
                        const uint64_t db = umma_smem_desc(sb + k * 32, 0, 1024, UMMA_LAYOUT_SW128);
                        umma_tf32_ss(d_tmem, da, db, a.idesc, (it | k) != 0 ? 1u : 0u);
                    }
                    if (a.cs == 1)
                        umma_commit(&empty_bar[stage]);
                    else
                        umma_commit_mc(&empty_bar[stage], cmask);
                    if (++stage == a.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int quarter = warp & 3;             // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;      // accumulator row == pixel within the tile
        const int hl = row / kTileW, wl = row % kTileW;
        const ConvEpilogue& ep = a.ep;
        int titer = 0;
        for (int st = cluster_id; st < a.nsuper; st += nclusters, ++titer) {
            const int tile = st * a.cs + crank;
            const int tw = tile % a.tiles_w;
            const int th = (tile / a.tiles_w) % a.tiles_h;
            const int b = tile / (a.tiles_w * a.tiles_h);
            const int h = th * kTileH + hl, w = tw * kTileW + wl;
            const bool valid = (tile < a.ntiles) && (h < a.H) && (w < a.W);
            const size_t pix = ((size_t)b * a.H + h) * a.W + w;
            const int buf = titer & 1;

            float x3v[3] = {0.f, 0.f, 0.f};
            if (ep.x3 && valid) {
                x3v[0] = ep.x3[pix * 3 + 0];
                x3v[1] = ep.x3[pix * 3 + 1];
                x3v[2] = ep.x3[pix * 3 + 2];
            }
            float fin[3] = {0.f, 0.f, 0.f};

            mbar_wait(&tfull_bar[buf], ((uint32_t)titer >> 1) & 1u);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kAccStride;

            for (int cc = 0; cc < N; cc += 16) {
                float v[16];
                tmem_ld16(taddr + cc, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += s_bias[cc + j];
                if (ep.w_res3) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float* wr = &s_wres3[(cc + j) * 3];
                        v[j] = fmaf(x3v[2], wr[2], fmaf(x3v[1], wr[1], fmaf(x3v[0], wr[0], v[j])));
                    }
                }
                if (valid) {
                    const size_t off = pix * N + cc;
                    if (ep.res_add) {
                        const float4* r4 = reinterpret_cast<const float4*>(ep.res_add + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 r = __ldg(r4 + q);
                            v[4 * q + 0] += r.x;
                            v[4 * q + 1] += r.y;
                            v[4 * q + 2] += r.z;
                            v[4 * q + 3] += r.w;
                        }
                    }
                    if (ep.out_pre) {
                        float4* o4 = reinterpret_cast<float4*>(ep.out_pre + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                    if (ep.gelu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (ep.dgelu_z) {
                        const float4* z4 = reinterpret_cast<const float4*>(ep.dgelu_z + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 z = __ldg(z4 + q);
                            v[4 * q + 0] *= gelu_erf_grad(z.x);
                            v[4 * q + 1] *= gelu_erf_grad(z.y);
                            v[4 * q + 2] *= gelu_erf_grad(z.z);
                            v[4 * q + 3] *= gelu_erf_grad(z.w);
                        }
                    }
                    if (ep.w_final) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            fin[0] = fmaf(v[j], s_wfinal[0 * N + cc + j], fin[0]);
                            fin[1] = fmaf(v[j], s_wfinal[1 * N + cc + j], fin[1]);
                            fin[2] = fmaf(v[j], s_wfinal[2 * N + cc + j], fin[2]);
                        }
                    }
                    if (ep.out) {
                        if (ep.round_tf32) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
                        }
                        float4* o4 = reinterpret_cast<float4*>(ep.out + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
            }
            // every tcgen05.ld of this buffer has completed (wait::ld inside tmem_ld16): hand it back
            tc_fence_before_sync();
            mbar_arrive(&tempty_bar[buf]);

            if (ep.w_final && valid) {
                const size_t plane = (size_t)a.H * a.W;
                float* o = ep.out_final + (size_t)b * 3 * plane + (size_t)h * a.W + w;
                o[0] = fin[0] + s_bfinal[0];
                o[plane] = fin[1] + s_bfinal[1];
                o[2 * plane] = fin[2] + s_bfinal[2];
            }
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before_sync();
    __syncthreads();
    if (a.cs > 1) cluster_sync_all();   // nobody exits while a peer may still multicast into its smem
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

// SINDDM_TC_CLUSTER = 1 | 2 | 4 (default 2): CTAs per cluster sharing each weight stage by TMA multicast
static int cluster_size_setting() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SINDDM_TC_CLUSTER");
        v = e ? atoi(e) : 2;
        if (v != 1 && v != 2 && v != 4) v = 2;
    }
    return v;
}

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
    op->p = p;
    SINDDM_TRY(make_tmap_nhwc(&op->tm_a, p.in, p.B, p.H, p.W, p.Cin, kKC, kTileW, kTileH, CU_TENSOR_MAP_SWIZZLE_128B));
    // cluster size: weight slices must be whole 8-row (1024 B) swizzle atoms
    int cs = cluster_size_setting();
    while (cs > 1 && (p.N % (8 * cs) != 0)) cs >>= 1;
    op->cs = cs;
    SINDDM_TRY(make_tmap_2d(&op->tm_b, p.w, p.Cin, p.ntaps * p.N, kKC, p.N / cs, CU_TENSOR_MAP_SWIZZLE_128B));
    if (p.in_res) {
        SINDDM_TRY(make_tmap_nhwc(&op->tm_ares, p.in_res, p.B, p.H, p.W, p.Cres, kKC, kTileW, kTileH,
                                  CU_TENSOR_MAP_SWIZZLE_128B));
        SINDDM_TRY(make_tmap_2d(&op->tm_bres, p.w_res, p.Cres, p.N, kKC, p.N / cs, CU_TENSOR_MAP_SWIZZLE_128B));
    } else {
        op->tm_ares = op->tm_a;
        op->tm_bres = op->tm_b;
    }
    op->stage_bytes = kABytes + (int)align_up((size_t)p.N * kKC * 4, 1024);
    const int budget = device_info().max_smem_optin - 1024 /*alignment slack*/ - kTailBytes;
    int nst = budget / op->stage_bytes;
    if (nst > 8) nst = 8;
    SINDDM_REQUIRE(nst >= 2, "tc_conv: not enough shared memory for a 2-stage pipeline");
    op->nstages = nst;
    op->smem_bytes = nst * op->stage_bytes + kTailBytes + 1024;
    op->tiles_w = ceil_div(p.W, kTileW);
    op->tiles_h = ceil_div(p.H, kTileH);
    op->ntiles = op->tiles_w * op->tiles_h * p.B;
    const int nsuper = ceil_div(op->ntiles, cs);
    int nclusters = device_info().num_sms / cs;
    if (nclusters > nsuper) nclusters = nsuper;
    op->grid = nclusters * cs;
    return SINDDM_OK;
}

int tc_conv_launch(const TcConvOp& op, cudaStream_t stream) {
    static int smem_set = 0;
    if (smem_set < op.smem_bytes) {
        SINDDM_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            device_info().max_smem_optin));
        smem_set = device_info().max_smem_optin;
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
    a.nstages = op.nstages;
    a.stage_bytes = op.stage_bytes;
    a.idesc = umma_idesc_tf32(kBM, p.N, 0, 0);
    a.ep = p.ep;
    // algorithmic work: real pixels x N x (taps*Cin + Cres) MACs
    prof_begin(stream, 0, 2.0 * (double)p.B * p.H * p.W * p.N * ((double)p.ntaps * p.Cin + a.Cres));
    a.cs = op.cs;
    a.b_rows = p.N / op.cs;
    a.nsuper = ceil_div(op.ntiles, op.cs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(op.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = op.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = op.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SINDDM_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_conv_kernel, op.tm_a, op.tm_ares, op.tm_b, op.tm_bres, a));
    prof_end(stream);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
