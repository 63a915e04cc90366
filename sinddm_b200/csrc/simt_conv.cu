// fp32 CUDA-core versions of the dense convolution and its weight gradient.
//
// Two jobs: (1) the layers that are no dense contraction -- Cin = 3 (l1.net[0], l1.res_conv) and
// Cout = 3 (final_conv, the data gradient into l1's depthwise output) -- which the tensor-core path does not
// take (reference SinDDM/models.py:124,130-132); (2) `math = fp32` strict mode, in which EVERY convolution
// runs here with plain FFMA so results can be checked against the CPU oracle at fp32 tolerance and the
// tcgen05 kernels can be checked against an exact on-device twin.  Same ConvProblem / WgradProblem
// contract as the tensor-core kernels, including every epilogue option.
#include <string.h>

#include "common.cuh"
#include "ops.h"

namespace sinddm {

namespace {

constexpr int kTH = 8, kTW = 16;   // pixel tile, one pixel per thread
constexpr int kNB = 32;            // output channels per pass (register accumulators)
constexpr int kCK = 8;             // input channels per smem weight slab

// weights slab in smem: ws[tap][ci (kCK)][kNB]  (co contiguous -> float4 broadcast reads)
__global__ void __launch_bounds__(kTH* kTW)
simt_conv_kernel(const ConvProblem p, int tiles_w, int tiles_h) {
    __shared__ __align__(16) float ws[9 * kCK * kNB];

    const int tile = blockIdx.x;
    const int tw = tile % tiles_w;
    const int th = (tile / tiles_w) % tiles_h;
    const int b = tile / (tiles_w * tiles_h);
    const int hl = threadIdx.x / kTW, wl = threadIdx.x % kTW;
    const int h = th * kTH + hl, w = tw * kTW + wl;
    const bool valid = (h < p.H) && (w < p.W);
    const size_t pix = ((size_t)b * p.H + h) * p.W + w;
    const ConvEpilogue& ep = p.ep;
    const int N = p.N;

    float x3v[3] = {0.f, 0.f, 0.f};
    if (ep.x3 && valid) {
        x3v[0] = ep.x3[pix * 3 + 0];
        x3v[1] = ep.x3[pix * 3 + 1];
        x3v[2] = ep.x3[pix * 3 + 2];
    }
    float fin[3] = {0.f, 0.f, 0.f};

    // neighbour pixel base pointers (nullptr = zero padding)
    for (int n0 = 0; n0 < N; n0 += kNB) {
        float acc[kNB];
#pragma unroll
        for (int j = 0; j < kNB; ++j) acc[j] = 0.f;

        // ---- main taps, then (phase 1) the optional 1x1 residual conv as extra K
        for (int phase = 0; phase < 2; ++phase) {
            const float* in = phase == 0 ? p.in : p.in_res;
            const float* wt = phase == 0 ? p.w : p.w_res;
            const int C = phase == 0 ? p.Cin : p.Cres;
            const int ntaps = phase == 0 ? p.ntaps : 1;
            if (in == nullptr) continue;
            for (int c0 = 0; c0 < C; c0 += kCK) {
                __syncthreads();
                // stage weights: ws[tap][ci][j] = wt[tap][n0+j][c0+ci]
                for (int i = threadIdx.x; i < ntaps * kCK * kNB; i += blockDim.x) {
                    const int j = i % kNB;
                    const int ci = (i / kNB) % kCK;
                    const int tap = i / (kNB * kCK);
                    const int n = n0 + j, c = c0 + ci;
                    ws[i] = (n < N && c < C) ? wt[((size_t)tap * N + n) * C + c] : 0.f;
                }
                __syncthreads();
                if (valid) {
                    for (int tap = 0; tap < ntaps; ++tap) {
                        int hh = h, ww = w;
                        if (ntaps == 9) {
                            hh += tap / 3 - 1;
                            ww += tap % 3 - 1;
                        }
                        if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W) continue;
                        const float* src = in + (((size_t)b * p.H + hh) * p.W + ww) * C + c0;
                        const int cmax = min(kCK, C - c0);
                        for (int ci = 0; ci < cmax; ++ci) {
                            const float av = __ldg(src + ci);
                            const float4* w4 = reinterpret_cast<const float4*>(&ws[(tap * kCK + ci) * kNB]);
#pragma unroll
                            for (int q = 0; q < kNB / 4; ++q) {
                                const float4 wv = w4[q];
                                acc[4 * q + 0] = fmaf(av, wv.x, acc[4 * q + 0]);
                                acc[4 * q + 1] = fmaf(av, wv.y, acc[4 * q + 1]);
                                acc[4 * q + 2] = fmaf(av, wv.z, acc[4 * q + 2]);
                                acc[4 * q + 3] = fmaf(av, wv.w, acc[4 * q + 3]);
                            }
                        }
                    }
                }
            }
        }

        // ---- epilogue (identical semantics to the tensor-core kernel's)
        if (valid) {
#pragma unroll
            for (int j = 0; j < kNB; ++j) {
                const int n = n0 + j;
                if (n < N) {
                    float v = acc[j];
                    if (ep.bias) v += ep.bias[n];
                    if (ep.w_res3) {
                        const float* wr = ep.w_res3 + n * 3;
                        v = fmaf(x3v[2], wr[2], fmaf(x3v[1], wr[1], fmaf(x3v[0], wr[0], v)));
                    }
                    const size_t off = pix * N + n;
                    if (ep.res_add) v += ep.res_add[off];
                    if (ep.out_pre) ep.out_pre[off] = v;
                    if (ep.gelu) v = gelu_erf(v);
                    if (ep.dgelu_z) v *= gelu_erf_grad(ep.dgelu_z[off]);
                    if (ep.w_final) {
                        fin[0] = fmaf(v, ep.w_final[0 * N + n], fin[0]);
                        fin[1] = fmaf(v, ep.w_final[1 * N + n], fin[1]);
                        fin[2] = fmaf(v, ep.w_final[2 * N + n], fin[2]);
                    }
                    if (ep.out) ep.out[off] = ep.round_tf32 ? round_tf32(v) : v;
                }
            }
        }
    }

    if (ep.w_final && valid) {
        const size_t plane = (size_t)p.H * p.W;
        float* o = ep.out_final + (size_t)b * 3 * plane + (size_t)h * p.W + w;
        o[0] = fin[0] + (ep.b_final ? ep.b_final[0] : 0.f);
        o[plane] = fin[1] + (ep.b_final ? ep.b_final[1] : 0.f);
        o[2 * plane] = fin[2] + (ep.b_final ? ep.b_final[2] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient, CUDA cores.  grid = (nsplit, ntaps, ceil(Cx/kCB)); thread <-> output channel co.
// ---------------------------------------------------------------------------------------------------
constexpr int kCB = 8;
constexpr int kPY = 4;   // pixel lanes per CTA (threadIdx.y), reduced through smem at the end

__global__ void simt_wgrad_kernel(const WgradProblem p) {
    extern __shared__ float red[];   // [kPY][kCB][blockDim.x]
    const int split = blockIdx.x;
    const int tap = blockIdx.y;
    const int ci0 = blockIdx.z * kCB;
    const int co = threadIdx.x;
    const int py = threadIdx.y;
    const long long P = (long long)p.B * p.H * p.W;
    const long long p_begin = P * split / p.nsplit;
    const long long p_end = P * (split + 1) / p.nsplit;
    int dyo = 0, dxo = 0;
    if (p.ntaps == 9) {
        dyo = tap / 3 - 1;
        dxo = tap % 3 - 1;
    }
    float acc[kCB];
#pragma unroll
    for (int i = 0; i < kCB; ++i) acc[i] = 0.f;
    const int cmax = min(kCB, p.Cx - ci0);
    const bool act = co < p.Cy;

    for (long long q = p_begin + py; q < p_end; q += kPY) {
        const int w = (int)(q % p.W);
        const int h = (int)((q / p.W) % p.H);
        const int hh = h + dyo, ww = w + dxo;
        if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W || !act) continue;
        const float* xs = p.x + (q + (long long)dyo * p.W + dxo) * p.Cx + ci0;
        const float g = __ldg(p.dy + q * p.Cy + co);
#pragma unroll
        for (int i = 0; i < kCB; ++i)
            if (i < cmax) acc[i] = fmaf(__ldg(xs + i), g, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kCB; ++i) red[(py * kCB + i) * blockDim.x + co] = acc[i];
    __syncthreads();
    if (py == 0 && act) {
        for (int i = 0; i < cmax; ++i) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kPY; ++y) s += red[(y * kCB + i) * blockDim.x + co];
            p.partial[(((size_t)split * p.ntaps + tap) * p.Cx + ci0 + i) * p.Cy + co] = s;
        }
    }
}

// dst[co][ci][tap] (or [tap][ci][co] when keep_layout) = sum_s partial[s][tap][ci][co], fixed order.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int ntaps, int Cx, int Cy,
                                    float* __restrict__ dst, int keep_layout) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const int total = ntaps * Cx * Cy;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // eight loads in flight per thread; the additions keep the split order (deterministic)
    float s = 0.f;
    int k = 0;
    for (; k + 8 <= nsplit; k += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(partial + (size_t)(k + j) * total + i);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
    }
    for (; k < nsplit; ++k) s += __ldcs(partial + (size_t)k * total + i);
    if (keep_layout == 2) {
        // im2col rows of a 3-channel 3x3 conv: ci row index = tap*3 + c (rows 27..31 are padding) -> OIHW [Cy][3][3][3]
        const int co = i % Cy;
        const int k = i / Cy;
        if (k < 27) dst[((size_t)co * 3 + k % 3) * 9 + k / 3] = s;
    } else if (keep_layout) {
        dst[i] = s;
    } else {
        const int co = i % Cy;
        const int ci = (i / Cy) % Cx;
        const int tap = i / (Cy * Cx);
        dst[((size_t)co * Cx + ci) * ntaps + tap] = s;
    }
}

// blocked = 0: dst[tap][n][k];  blocked = 1: dst[tap][k/32][n][32] with 16-byte groups XOR-swizzled by (n & 7),
// i.e. exactly the shared-memory image of a 128B-swizzled K-major UMMA operand tile (padding pre-zeroed).
__device__ __forceinline__ size_t packed_index(int tap, int n, int k, int N, int K, int blocked) {
    if (!blocked) return ((size_t)tap * N + n) * K + k;
    const int nchunk = (K + 31) >> 5;
    const int chunk = k >> 5, kk = k & 31;
    const int grp = (kk >> 2) ^ (n & 7);
    return (((size_t)tap * nchunk + chunk) * N + n) * 32 + grp * 4 + (kk & 3);
}

__device__ __forceinline__ void pack_conv_one(int i, const float* __restrict__ w, int Cout, int Cin, int ntaps,
                                              float* __restrict__ dst_fwd, float* __restrict__ dst_dgrad, int round,
                                              int blocked, int dgrad_rows) {
    const int tap = i % ntaps;
    const int ci = (i / ntaps) % Cin;
    const int co = i / (ntaps * Cin);
    float v = w[i];
    if (round == 2) {
        // 3xTF32 (MATH_TF32X3): K is tripled as [lo | hi | hi], matching activations split as [hi | lo | hi]
        // (split3_launch), so that one pass over K adds x_hi w_lo + x_lo w_hi + x_hi w_hi into the fp32 accumulator.
        // The two small cross terms come FIRST: the tensor core truncates its fp32 accumulator at every MMA step (an
        // error proportional to the accumulator's magnitude at that step), so only the last third of the steps --
        // the ones that add x_hi w_hi -- still pay it.
        const float hi = round_tf32(v), lo = round_tf32(v - hi);
        if (dst_fwd) {
            dst_fwd[packed_index(tap, co, ci, Cout, 3 * Cin, blocked)] = lo;
            dst_fwd[packed_index(tap, co, Cin + ci, Cout, 3 * Cin, blocked)] = hi;
            dst_fwd[packed_index(tap, co, 2 * Cin + ci, Cout, 3 * Cin, blocked)] = hi;
        }
        if (dst_dgrad) {
            dst_dgrad[packed_index(ntaps - 1 - tap, ci, co, dgrad_rows, 3 * Cout, blocked)] = lo;
            dst_dgrad[packed_index(ntaps - 1 - tap, ci, Cout + co, dgrad_rows, 3 * Cout, blocked)] = hi;
            dst_dgrad[packed_index(ntaps - 1 - tap, ci, 2 * Cout + co, dgrad_rows, 3 * Cout, blocked)] = hi;
        }
        return;
    }
    if (round) v = round_tf32(v);
    if (dst_fwd) dst_fwd[packed_index(tap, co, ci, Cout, Cin, blocked)] = v;
    if (dst_dgrad) dst_dgrad[packed_index(ntaps - 1 - tap, ci, co, dgrad_rows, Cout, blocked)] = v;
}

// 3xTF32 operand split: hi = tf32(x) (round to nearest), lo = tf32(x - hi) (x - hi is exact in fp32), so that
// x = hi + lo up to 2^-22 |x|.  mode 0: out[p][3C] = [hi | lo | hi] (conv operand: channels tripled, meets weights
// packed [lo | hi | hi]); mode 1: out[3][P][C] = hi, lo, hi (weight-gradient x operand: batch tripled); mode 2:
// out[3][P][C] = lo, hi, hi (weight-gradient dy operand).  The cross terms hi*lo, lo*hi come first along the
// contraction axis, hi*hi last (see pack_conv_one).  One float4 per thread; C % 4 == 0.
__global__ void __launch_bounds__(256)
split3_kernel(const float4* __restrict__ x, float4* __restrict__ out, long long n4, int C4, int mode) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        float4 hi, lo;
        hi.x = round_tf32(v.x); hi.y = round_tf32(v.y); hi.z = round_tf32(v.z); hi.w = round_tf32(v.w);
        lo.x = round_tf32(v.x - hi.x); lo.y = round_tf32(v.y - hi.y);
        lo.z = round_tf32(v.z - hi.z); lo.w = round_tf32(v.w - hi.w);
        if (mode == 0) {
            const long long p = i / C4;
            const long long o = p * 3 * C4 + (i - p * C4);
            out[o] = hi;
            out[o + C4] = lo;
            out[o + 2 * C4] = hi;
        } else {
            out[i] = mode == 1 ? hi : lo;
            out[n4 + i] = mode == 1 ? lo : hi;
            out[2 * n4 + i] = hi;
        }
    }
}

__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJobs jobs) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    int k = 0;
    while (k + 1 < jobs.n && (int)blockIdx.x >= jobs.j[k + 1].block0) ++k;
    const PackJob& J = jobs.j[k];
    const int i = ((int)blockIdx.x - J.block0) * 256 + (int)threadIdx.x;
    if (i >= J.total) return;
    if (J.kind == 0) {
        pack_conv_one(i, J.w, J.Cout, J.Cin, J.ntaps, J.dst_fwd, J.dst_dgrad, J.round, J.blocked, J.dgrad_rows);
    } else if (J.kind == 1) {
        const int tap = i % 9, c = (i / 9) % 3, co = i / 27;
        float v = J.w[i];
        if (J.round) v = round_tf32(v);
        J.dst_fwd[packed_index(0, co, tap * 3 + c, J.Cout, 32, J.blocked)] = v;
    } else {
        J.dst_fwd[i] = J.w[i] + J.w2[i];
    }
}

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int ntaps,
                                         float* __restrict__ dst_fwd, float* __restrict__ dst_dgrad, int round,
                                         int blocked, int dgrad_rows) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const int total = Cout * Cin * ntaps;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    pack_conv_one(i, w, Cout, Cin, ntaps, dst_fwd, dst_dgrad, round, blocked, dgrad_rows);
}

// ---------------------------------------------------------------------------------------------------
// im2col for the one 3x3 convolution whose input has 3 channels (l1.net[0], reference SinDDM/models.py:62): its K is
// only 27, so instead of a CUDA-core kernel the 27 patch values of every pixel are laid out as ONE 32-channel
// (128-byte) row and the layer runs as a 1x1 tensor-core GEMM with the usual fused epilogue; the weight gradient
// reads the same rows.  out[p][tap*3 + c] = x[p (+) tap][c], columns 27..31 = 0; 8 threads (float4 each) per pixel.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col3x3_c3_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W, int round) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const long long total = (long long)B * H * W * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i & 7);
        const long long p = i >> 3;
        const int w = (int)(p % W), h = (int)((p / W) % H);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = q * 4 + j;
            const int tap = k / 3, c = k - tap * 3;
            const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
            float t = 0.f;
            if (k < 27 && hh >= 0 && hh < H && ww >= 0 && ww < W)
                t = __ldg(x + (p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * 3 + c);
            v[j] = round ? round_tf32(t) : t;
        }
        reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// w [Co][3][3][3] (OIHW) -> the 1x1 GEMM operand [1][Co][32] with k = tap*3 + c, in either packed layout
__global__ void pack_im2col_weights_kernel(const float* __restrict__ w, int Co, float* __restrict__ dst, int round,
                                           int blocked) {
    pdl_grid_sync();   // programmatic dependent launch: nothing above touches global memory
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Co * 27) return;
    const int tap = i % 9, c = (i / 9) % 3, co = i / 27;
    float v = w[i];
    if (round) v = round_tf32(v);
    dst[packed_index(0, co, tap * 3 + c, Co, 32, blocked)] = v;
}

// ---------------------------------------------------------------------------------------------------
// Specialisations for the 3-channel layers (they are HBM / LSU bound, not contractions):
//   small Cin  (l1.net[0] 3->80 3x3, final_conv data gradient 3->80 1x1): the whole K = ntaps*Cin <= 27
//              input patch of a pixel sits in registers, weights [K][N] in smem are read as float4 broadcasts;
//   small N    (data gradient into l1's depthwise output, 80->3 3x3): 3 accumulators per pixel, inputs read as
//              float4 along the channel dimension, weights [tap][ci][4] in smem.
// ---------------------------------------------------------------------------------------------------
SINDDM_DEVINL void conv_epilogue_one(const ConvEpilogue& ep, float v, int n, int N, size_t pix, const float (&x3v)[3],
                                     float (&fin)[3]) {
    if (ep.bias) v += ep.bias[n];
    if (ep.w_res3) {
        const float* wr = ep.w_res3 + n * 3;
        v = fmaf(x3v[2], wr[2], fmaf(x3v[1], wr[1], fmaf(x3v[0], wr[0], v)));
    }
    const size_t off = pix * N + n;
    if (ep.res_add) v += ep.res_add[off];
    if (ep.out_pre) ep.out_pre[off] = v;
    if (ep.gelu) v = gelu_erf(v);
    if (ep.dgelu_z) v *= gelu_erf_grad(ep.dgelu_z[off]);
    if (ep.w_final) {
        fin[0] = fmaf(v, ep.w_final[0 * N + n], fin[0]);
        fin[1] = fmaf(v, ep.w_final[1 * N + n], fin[1]);
        fin[2] = fmaf(v, ep.w_final[2 * N + n], fin[2]);
    }
    if (ep.out) ep.out[off] = ep.round_tf32 ? round_tf32(v) : v;
}

template <int NTAPS, int CIN>
__global__ void __launch_bounds__(128)
simt_conv_smallcin_kernel(const ConvProblem p) {
    constexpr int K = NTAPS * CIN;
    extern __shared__ __align__(16) float ws[];  // [K][Npad]
    const int N = p.N;
    const int Npad = (N + 3) & ~3;
    for (int i = threadIdx.x; i < K * Npad; i += blockDim.x) {
        const int n = i % Npad, k = i / Npad;
        const int tap = k / CIN, ci = k % CIN;
        ws[i] = n < N ? p.w[((size_t)tap * N + n) * CIN + ci] : 0.f;
    }
    __syncthreads();
    const long long P = (long long)p.B * p.H * p.W;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const int w = (int)(pix % p.W);
    const int h = (int)((pix / p.W) % p.H);
    float xin[K];
#pragma unroll
    for (int tap = 0; tap < NTAPS; ++tap) {
        int hh = h, ww = w;
        if (NTAPS == 9) {
            hh += tap / 3 - 1;
            ww += tap % 3 - 1;
        }
        const bool ok = hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
        const float* src = p.in + (pix + (long long)(hh - h) * p.W + (ww - w)) * CIN;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xin[tap * CIN + ci] = ok ? __ldg(src + ci) : 0.f;
    }
    float x3v[3] = {0.f, 0.f, 0.f};
    if (p.ep.x3) {
        x3v[0] = p.ep.x3[pix * 3 + 0];
        x3v[1] = p.ep.x3[pix * 3 + 1];
        x3v[2] = p.ep.x3[pix * 3 + 2];
    }
    float fin[3] = {0.f, 0.f, 0.f};
    for (int n0 = 0; n0 < N; n0 += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[k * Npad + n0]);
            acc.x = fmaf(xin[k], wv.x, acc.x);
            acc.y = fmaf(xin[k], wv.y, acc.y);
            acc.z = fmaf(xin[k], wv.z, acc.z);
            acc.w = fmaf(xin[k], wv.w, acc.w);
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        if ((N & 3) == 0) {
            // vector epilogue: one 16-byte store per output tensor instead of four scattered 4-byte stores
            const ConvEpilogue& ep = p.ep;
            const size_t off = (size_t)pix * N + n0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ep.bias) v[j] += ep.bias[n0 + j];
                if (ep.w_res3) {
                    const float* wr = ep.w_res3 + (n0 + j) * 3;
                    v[j] = fmaf(x3v[2], wr[2], fmaf(x3v[1], wr[1], fmaf(x3v[0], wr[0], v[j])));
                }
            }
            if (ep.res_add) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(ep.res_add + off));
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
            if (ep.out_pre) *reinterpret_cast<float4*>(ep.out_pre + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (ep.gelu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
            }
            if (ep.dgelu_z) {
                const float4 z = __ldg(reinterpret_cast<const float4*>(ep.dgelu_z + off));
                v[0] *= gelu_erf_grad(z.x); v[1] *= gelu_erf_grad(z.y);
                v[2] *= gelu_erf_grad(z.z); v[3] *= gelu_erf_grad(z.w);
            }
            if (ep.w_final) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    fin[0] = fmaf(v[j], ep.w_final[0 * N + n0 + j], fin[0]);
                    fin[1] = fmaf(v[j], ep.w_final[1 * N + n0 + j], fin[1]);
                    fin[2] = fmaf(v[j], ep.w_final[2 * N + n0 + j], fin[2]);
                }
            }
            if (ep.out) {
                if (ep.round_tf32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = round_tf32(v[j]);
                }
                *reinterpret_cast<float4*>(ep.out + off) = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + j < N) conv_epilogue_one(p.ep, v[j], n0 + j, N, (size_t)pix, x3v, fin);
        }
    }
    if (p.ep.w_final) {
        const size_t plane = (size_t)p.H * p.W;
        const int b = (int)(pix / (long long)plane);
        float* o = p.ep.out_final + (size_t)b * 3 * plane + (size_t)h * p.W + w;
        o[0] = fin[0] + (p.ep.b_final ? p.ep.b_final[0] : 0.f);
        o[plane] = fin[1] + (p.ep.b_final ? p.ep.b_final[1] : 0.f);
        o[2 * plane] = fin[2] + (p.ep.b_final ? p.ep.b_final[2] : 0.f);
    }
}

// N <= 4, Cin % 4 == 0
__global__ void __launch_bounds__(128)
simt_conv_smalln_kernel(const ConvProblem p) {
    extern __shared__ __align__(16) float ws[];  // [ntaps][Cin][4]
    const int N = p.N, C = p.Cin;
    for (int i = threadIdx.x; i < p.ntaps * C * 4; i += blockDim.x) {
        const int n = i & 3, ci = (i >> 2) % C, tap = (i >> 2) / C;
        ws[i] = n < N ? p.w[((size_t)tap * N + n) * C + ci] : 0.f;
    }
    __syncthreads();
    const long long P = (long long)p.B * p.H * p.W;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const int w = (int)(pix % p.W);
    const int h = (int)((pix / p.W) % p.H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < p.ntaps; ++tap) {
        int hh = h, ww = w;
        if (p.ntaps == 9) {
            hh += tap / 3 - 1;
            ww += tap % 3 - 1;
        }
        if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W) continue;
        const float4* src = reinterpret_cast<const float4*>(p.in + (pix + (long long)(hh - h) * p.W + (ww - w)) * C);
        const float4* wt = reinterpret_cast<const float4*>(ws + (size_t)tap * C * 4);
#pragma unroll 4
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 xv = __ldg(src + c4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 wv = wt[c4 * 4 + j];
                acc[0] = fmaf(xs[j], wv.x, acc[0]);
                acc[1] = fmaf(xs[j], wv.y, acc[1]);
                acc[2] = fmaf(xs[j], wv.z, acc[2]);
                acc[3] = fmaf(xs[j], wv.w, acc[3]);
            }
        }
    }
    float x3v[3] = {0.f, 0.f, 0.f};
    float fin[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < N; ++n) conv_epilogue_one(p.ep, acc[n], n, N, (size_t)pix, x3v, fin);
}

// Cin == 3, N % 4 == 0 (l1.net[0]; the data gradient of final_conv): thread <-> (pixel pair, 4 output channels).
// Consecutive threads own consecutive channel quads of one pixel pair, so every float4 store of a warp is
// contiguous NHWC memory (the per-pixel kernel above scatters each store instruction over 32 cache lines), and
// the quad's 4 x NTAPS*3 weights live in registers because the quad of a thread never changes.
template <int NTAPS>
__global__ void __launch_bounds__(256) simt_conv_cin3_kernel(const ConvProblem p, int n4, int ppb) {
    constexpr int K = NTAPS * 3;
    constexpr int R = NTAPS == 9 ? 3 : 1;      // window rows
    constexpr int CW = NTAPS == 9 ? 4 : 2;     // window columns of a pixel pair
    const int N = p.N;
    const int quad = threadIdx.x % n4, pl = threadIdx.x / n4;
    const int n0 = quad * 4;
    const ConvEpilogue& ep = p.ep;
    float wr[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) wr[k][j] = p.w[((size_t)(k / 3) * N + n0 + j) * 3 + (k % 3)];
    }
    float bias4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias4[j] = ep.bias ? ep.bias[n0 + j] : 0.f;

    const int wp = (p.W + 1) >> 1;                       // pixel pairs per row
    const long long npairs = (long long)p.B * p.H * wp;
    for (long long pg = blockIdx.x; pg * ppb < npairs; pg += gridDim.x) {
        const long long pp = pg * ppb + pl;
        if (pp >= npairs) continue;
        const int w = (int)(pp % wp) * 2;
        const long long row = pp / wp;                   // b * H + h
        const int h = (int)(row % p.H);
        // window of x: rows h-1..h+1, columns w-1..w+2 (3x3) or row h, columns w..w+1 (1x1)
        float xw[R][CW][3];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int hh = h + (NTAPS == 9 ? r - 1 : 0);
            const bool rok = hh >= 0 && hh < p.H;
            const float* xrow = p.in + (row + (NTAPS == 9 ? r - 1 : 0)) * p.W * 3;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const int ww = w + (NTAPS == 9 ? c - 1 : c);
                const bool ok = rok && ww >= 0 && ww < p.W;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) xw[r][c][ci] = ok ? __ldg(xrow + (size_t)ww * 3 + ci) : 0.f;
            }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            if (w + px >= p.W) break;
            float v[4] = {bias4[0], bias4[1], bias4[2], bias4[3]};
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int c = 0; c < (NTAPS == 9 ? 3 : 1); ++c) {
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        const float xv = xw[r][c + px][ci];
                        const int k = (r * (NTAPS == 9 ? 3 : 1) + c) * 3 + ci;
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = fmaf(xv, wr[k][j], v[j]);
                    }
                }
            }
            const size_t off = (size_t)(row * p.W + w + px) * N + n0;
            if (ep.res_add) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(ep.res_add + off));
                v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
            }
            if (ep.out_pre) __stcs(reinterpret_cast<float4*>(ep.out_pre + off), make_float4(v[0], v[1], v[2], v[3]));
            if (ep.gelu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = ep.fast_math ? gelu_fast(v[j]) : gelu_erf(v[j]);
            }
            if (ep.dgelu_z) {
                const float4 z = __ldg(reinterpret_cast<const float4*>(ep.dgelu_z + off));
                const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= ep.fast_math ? gelu_grad_fast(zz[j]) : gelu_erf_grad(zz[j]);
            }
            if (ep.out) {
                if (ep.round_tf32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = round_tf32(v[j]);
                }
                *reinterpret_cast<float4*>(ep.out + off) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

// N <= 3, Cin % 4 == 0, Cin <= 128 (the data gradient into l1's 3-channel depthwise output): one WARP per pixel,
// lane j owns input channels 4j..4j+3, so a tap is one contiguous float4 load per lane (the per-pixel kernel above
// strides every load instruction over 32 pixels) and the lane's 3 x NTAPS x 4 weights stay in registers; the three
// sums are combined with shuffles.  A block is a strip of 12 pixels walking down the image: the rows it re-reads
// for the vertical taps are still in L1.
constexpr int kN3Warps = 12;

template <int NTAPS>
__global__ void __launch_bounds__(kN3Warps * 32, 1) simt_conv_n3_kernel(const ConvProblem p, int nstrips, int hsplit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = p.Cin, N = p.N;
    const bool lact = lane * 4 < C;
    float wr[NTAPS][3][4];
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
        for (int n = 0; n < 3; ++n) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                wr[t][n][j] = (lact && n < N) ? p.w[((size_t)t * N + n) * C + lane * 4 + j] : 0.f;
        }
    }
    int bid = blockIdx.x;
    const int part = bid % hsplit;
    bid /= hsplit;
    const int strip = bid % nstrips;
    const int b = bid / nstrips;
    const int w = strip * kN3Warps + warp;
    if (w >= p.W) return;
    const int h_begin = (int)((long long)p.H * part / hsplit), h_end = (int)((long long)p.H * (part + 1) / hsplit);
    for (int h = h_begin; h < h_end; ++h) {
        const size_t pix = ((size_t)b * p.H + h) * p.W + w;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < NTAPS; ++t) {
            const int hh = h + (NTAPS == 9 ? t / 3 - 1 : 0), ww = w + (NTAPS == 9 ? t % 3 - 1 : 0);
            if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W || !lact) continue;
            const float4 xv = __ldg(reinterpret_cast<const float4*>(
                                        p.in + (((size_t)b * p.H + hh) * p.W + ww) * C) + lane);
#pragma unroll
            for (int n = 0; n < 3; ++n)
                acc[n] = fmaf(xv.w, wr[t][n][3], fmaf(xv.z, wr[t][n][2], fmaf(xv.y, wr[t][n][1], fmaf(xv.x, wr[t][n][0], acc[n]))));
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) acc[n] = warp_sum(acc[n]);
        if (lane == 0) {
            float x3v[3] = {0.f, 0.f, 0.f};
            float fin[3] = {0.f, 0.f, 0.f};
            for (int n = 0; n < N; ++n) conv_epilogue_one(p.ep, acc[n], n, N, pix, x3v, fin);
        }
    }
}

// Weight gradient for Cx == 3 (l1.net[0], l1.res_conv, final_conv): dy is read ONCE for all taps.
// One thread per output channel co walks row segments pixel by pixel; the 3x3 window of the three x channels
// lives in registers and slides (9 new broadcast loads per pixel instead of 27), dy is prefetched one block of
// six pixels ahead.  grid = nsplit CTAs of (Cy padded to 32, kPY) threads; walkers take (row, segment) items
// round robin; partial[split][tap][ci][co] like the other weight-gradient kernels.
constexpr int kCx3Block = 6;    // pixels per unrolled block = two rotations of the three window columns
constexpr int kCx3Seg = 66;     // pixels per row segment (multiple of kCx3Block)

template <int NTAPS>
__global__ void __launch_bounds__(384, 2) simt_wgrad_cx3_kernel(const WgradProblem p, int seg_len, int nseg) {
    constexpr int K = NTAPS * 3;
    constexpr int R = NTAPS == 9 ? 3 : 1;      // window rows
    extern __shared__ float red[];  // [kPY][K][blockDim.x]
    const int co = threadIdx.x, py = threadIdx.y;
    const bool act = co < p.Cy;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;

    const int items = p.B * p.H * nseg;
    const int nwalk = gridDim.x * kPY;
    for (int item = blockIdx.x * kPY + py; item < items && act; item += nwalk) {
        const int row = item / nseg, seg = item - row * nseg;
        const int h = row % p.H;
        const int w_begin = seg * seg_len;
        const int w_end = min(p.W, w_begin + seg_len);
        const float* dyrow = p.dy + (size_t)row * p.W * p.Cy + co;
        const float* xrow[R];
        bool rok[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int hh = h + (NTAPS == 9 ? r - 1 : 0);
            rok[r] = hh >= 0 && hh < p.H;
            xrow[r] = p.x + ((size_t)row + (NTAPS == 9 ? r - 1 : 0)) * p.W * 3;
        }
        auto load_col = [&](float (&c)[R][3], int w) {
            const bool wok = w >= 0 && w < p.W;
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) c[r][ci] = (wok && rok[r]) ? __ldg(xrow[r] + (size_t)w * 3 + ci) : 0.f;
            }
        };
        auto load_g = [&](float (&g)[kCx3Block], int w) {
#pragma unroll
            for (int u = 0; u < kCx3Block; ++u) g[u] = (w + u < w_end) ? __ldg(dyrow + (size_t)(w + u) * p.Cy) : 0.f;
        };
        if (NTAPS == 9) {
            float c0[R][3], c1[R][3], c2[R][3];
            load_col(c0, w_begin - 1);
            load_col(c1, w_begin);
            // columns ca / cb / cc are at w-1 / w / w+1; the right one is fetched here
            auto step = [&](float (&ca)[R][3], float (&cb)[R][3], float (&cc)[R][3], int w, float g) {
                load_col(cc, w + 1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        acc[(r * 3 + 0) * 3 + ci] = fmaf(ca[r][ci], g, acc[(r * 3 + 0) * 3 + ci]);
                        acc[(r * 3 + 1) * 3 + ci] = fmaf(cb[r][ci], g, acc[(r * 3 + 1) * 3 + ci]);
                        acc[(r * 3 + 2) * 3 + ci] = fmaf(cc[r][ci], g, acc[(r * 3 + 2) * 3 + ci]);
                    }
                }
            };
            // no software prefetch here: 27 FMAs per pixel and 24 resident warps cover the dy latency
            for (int w = w_begin; w < w_end; w += kCx3Block) {
                float g[kCx3Block];
                load_g(g, w);
                step(c0, c1, c2, w + 0, g[0]);
                step(c1, c2, c0, w + 1, g[1]);
                step(c2, c0, c1, w + 2, g[2]);
                step(c0, c1, c2, w + 3, g[3]);
                step(c1, c2, c0, w + 4, g[4]);
                step(c2, c0, c1, w + 5, g[5]);
            }
        } else {
            float gn[kCx3Block];
            load_g(gn, w_begin);
            for (int w = w_begin; w < w_end; w += kCx3Block) {
                float g[kCx3Block];
                float c[kCx3Block][R][3];
#pragma unroll
                for (int u = 0; u < kCx3Block; ++u) {
                    g[u] = gn[u];
                    load_col(c[u], w + u);
                }
                load_g(gn, w + kCx3Block);
#pragma unroll
                for (int u = 0; u < kCx3Block; ++u) {
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) acc[ci] = fmaf(c[u][0][ci], g[u], acc[ci]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) red[(py * K + k) * blockDim.x + co] = acc[k];
    __syncthreads();
    if (py == 0 && act) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kPY; ++y) s += red[(y * K + k) * blockDim.x + co];
            // k = tap*3 + ci  ->  partial[split][tap][ci][co]
            p.partial[((size_t)blockIdx.x * K + k) * p.Cy + co] = s;
        }
    }
}

}  // namespace

int simt_conv_launch(const ConvProblem& p, cudaStream_t stream) {
    SINDDM_REQUIRE(p.ntaps == 9 || p.ntaps == 1, "simt_conv: ntaps must be 9 or 1");
    SINDDM_REQUIRE(p.N >= 1 && p.Cin >= 1, "simt_conv: bad channel counts");
    SINDDM_REQUIRE(!p.ep.pre_grad && !p.ep.out3, "simt_conv: pre_grad / out3 are tensor-core epilogue options");
    const long long P = (long long)p.B * p.H * p.W;
    if (p.Cin == 3 && p.in_res == nullptr && p.N % 4 == 0 && p.N <= 512 && !p.ep.x3 && !p.ep.w_final) {
        const int n4 = p.N / 4;
        const int ppb = n4 >= 256 ? 1 : 256 / n4;
        const long long npairs = (long long)p.B * p.H * ((p.W + 1) / 2);
        const long long want = (npairs + ppb - 1) / ppb;
        const long long cap = (long long)(device_info().initialized ? device_info().num_sms : 148) * 8;
        const unsigned blocks = (unsigned)(want < cap ? want : cap);
        if (p.ntaps == 9)
            simt_conv_cin3_kernel<9><<<blocks, n4 * ppb, 0, stream>>>(p, n4, ppb);
        else
            simt_conv_cin3_kernel<1><<<blocks, n4 * ppb, 0, stream>>>(p, n4, ppb);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    if (p.N <= 3 && p.Cin % 4 == 0 && p.Cin <= 128 && p.in_res == nullptr && !p.ep.w_final && !p.ep.x3) {
        const int nstrips = ceil_div(p.W, kN3Warps);
        const int hsplit = p.H >= 32 ? 2 : 1;
        const unsigned blocks = (unsigned)(p.B * nstrips * hsplit);
        if (p.ntaps == 9)
            simt_conv_n3_kernel<9><<<blocks, kN3Warps * 32, 0, stream>>>(p, nstrips, hsplit);
        else
            simt_conv_n3_kernel<1><<<blocks, kN3Warps * 32, 0, stream>>>(p, nstrips, hsplit);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    if (p.Cin == 3 && p.in_res == nullptr && p.N <= 512) {
        const size_t smem = (size_t)p.ntaps * 3 * ((p.N + 3) & ~3) * sizeof(float);
        const unsigned blocks = (unsigned)((P + 127) / 128);
        if (p.ntaps == 9)
            simt_conv_smallcin_kernel<9, 3><<<blocks, 128, smem, stream>>>(p);
        else
            simt_conv_smallcin_kernel<1, 3><<<blocks, 128, smem, stream>>>(p);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    if (p.N <= 4 && p.Cin % 4 == 0 && p.in_res == nullptr && p.ntaps * p.Cin * 16 <= 48 * 1024 && !p.ep.w_final &&
        !p.ep.x3) {
        const size_t smem = (size_t)p.ntaps * p.Cin * 4 * sizeof(float);
        simt_conv_smalln_kernel<<<(unsigned)((P + 127) / 128), 128, smem, stream>>>(p);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    const int tiles_w = ceil_div(p.W, kTW), tiles_h = ceil_div(p.H, kTH);
    const int ntiles = tiles_w * tiles_h * p.B;
    simt_conv_kernel<<<ntiles, kTH * kTW, 0, stream>>>(p, tiles_w, tiles_h);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int simt_wgrad_nsplit(int B, int H, int W, int Cx, int Cy, int ntaps) {
    long long P = (long long)B * H * W;
    long long n = P / 1024;
    if (n < 1) n = 1;
    if (n > 1024) n = 1024;
    // keep the split-partial scratch of one layer under 64 MiB
    while (n > 1 && n * ntaps * Cx * Cy * 4 > (64ll << 20)) n /= 2;
    // the Cx == 3 row walker wants two resident CTAs per SM and a short split reduction
    const long long sms = device_info().initialized ? device_info().num_sms : 148;
    if (Cx == 3 && n > 2 * sms) n = 2 * sms;
    return (int)n;
}

int simt_wgrad_launch(const WgradProblem& p, cudaStream_t stream) {
    dim3 grid(p.nsplit, p.ntaps, ceil_div(p.Cx, kCB));
    const int tx = (int)align_up((size_t)p.Cy, 32);
    SINDDM_REQUIRE(tx * kPY <= 1024, "simt_wgrad: Cy=%d too large", p.Cy);
    dim3 block(tx, kPY);
    if (p.Cx == 3 && tx * kPY <= 384 && (size_t)kPY * p.ntaps * 3 * tx * sizeof(float) <= 48 * 1024) {
        const size_t smem = (size_t)kPY * p.ntaps * 3 * tx * sizeof(float);
        const int seg_len = p.W > kCx3Seg ? kCx3Seg : ceil_div(p.W, kCx3Block) * kCx3Block;
        const int nseg = ceil_div(p.W, seg_len);
        if (p.ntaps == 9)
            simt_wgrad_cx3_kernel<9><<<p.nsplit, block, smem, stream>>>(p, seg_len, nseg);
        else
            simt_wgrad_cx3_kernel<1><<<p.nsplit, block, smem, stream>>>(p, seg_len, nseg);
        SINDDM_CUDA_OK(cudaGetLastError());
        return SINDDM_OK;
    }
    simt_wgrad_kernel<<<grid, block, (size_t)kPY * kCB * tx * sizeof(float), stream>>>(p);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int wgrad_reduce_launch(const float* partial, int nsplit, int ntaps, int Cx, int Cy, float* dst, int keep_layout,
                        cudaStream_t stream) {
    const int total = ntaps * Cx * Cy;
    (void)launch_pdl(wgrad_reduce_kernel, dim3(ceil_div(total, 256)), dim3(256), (size_t)(0), stream, partial, nsplit, ntaps, Cx, Cy, dst, keep_layout);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int im2col3x3_c3_launch(const float* x, float* out, int B, int H, int W, int round, cudaStream_t stream) {
    const long long total = (long long)B * H * W * 8;
    const long long cap = 16ll * (device_info().initialized ? device_info().num_sms : 148);
    const long long want = (total + 255) / 256;
    (void)launch_pdl(im2col3x3_c3_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), (size_t)(0), stream, x, out, B, H, W, round);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int pack_im2col_weights_launch(const float* w, int Co, float* dst, int round, int blocked, cudaStream_t stream) {
    (void)launch_pdl(pack_im2col_weights_kernel, dim3(ceil_div(Co * 27, 256)), dim3(256), (size_t)(0), stream, w, Co, dst, round, blocked);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int pack_conv_weights_launch(const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad, int round,
                             cudaStream_t stream, int blocked, int dgrad_rows) {
    const int total = Cout * Cin * ntaps;
    (void)launch_pdl(pack_conv_weights_kernel, dim3(ceil_div(total, 256)), dim3(256), (size_t)(0), stream, w, Cout, Cin, ntaps, dst_fwd, dst_dgrad, round,
                                                                       blocked, dgrad_rows > Cin ? dgrad_rows : Cin);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

int split3_launch(const float* x, long long P, int C, float* out, int mode, cudaStream_t stream) {
    SINDDM_REQUIRE(C % 4 == 0 && mode >= 0 && mode <= 2, "split3: C=%d mode=%d unsupported", C, mode);
    const long long n4 = P * (C / 4);
    const long long cap = 32ll * (device_info().initialized ? device_info().num_sms : 148);
    const long long want = (n4 + 255) / 256;
    (void)launch_pdl(split3_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), (size_t)0, stream,
                     reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), n4, C / 4, mode);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

static PackJob* pack_jobs_next(PackJobs* jobs) {
    if (jobs->n >= kMaxPackJobs) return nullptr;
    PackJob* j = &jobs->j[jobs->n++];
    memset(j, 0, sizeof(*j));
    return j;
}

void pack_jobs_add_conv(PackJobs* jobs, const float* w, int Cout, int Cin, int ntaps, float* dst_fwd, float* dst_dgrad,
                        int round, int blocked, int dgrad_rows) {
    PackJob* j = pack_jobs_next(jobs);
    if (!j) return;
    j->kind = 0; j->w = w; j->Cout = Cout; j->Cin = Cin; j->ntaps = ntaps; j->dst_fwd = dst_fwd; j->dst_dgrad = dst_dgrad;
    j->round = round; j->blocked = blocked; j->dgrad_rows = dgrad_rows > Cin ? dgrad_rows : Cin;
    j->total = Cout * Cin * ntaps;
}

void pack_jobs_add_im2col(PackJobs* jobs, const float* w, int Co, float* dst, int round, int blocked) {
    PackJob* j = pack_jobs_next(jobs);
    if (!j) return;
    j->kind = 1; j->w = w; j->Cout = Co; j->dst_fwd = dst; j->round = round; j->blocked = blocked; j->total = Co * 27;
}

void pack_jobs_add_sum(PackJobs* jobs, const float* a, const float* b, float* o, int n) {
    PackJob* j = pack_jobs_next(jobs);
    if (!j) return;
    j->kind = 2; j->w = a; j->w2 = b; j->dst_fwd = o; j->total = n;
}

int pack_jobs_launch(PackJobs* jobs, cudaStream_t stream) {
    SINDDM_REQUIRE(jobs->n < kMaxPackJobs, "pack_jobs: too many jobs");
    if (jobs->n == 0) return SINDDM_OK;
    int blocks = 0;
    for (int k = 0; k < jobs->n; ++k) {
        jobs->j[k].block0 = blocks;
        blocks += ceil_div(jobs->j[k].total, 256);
    }
    (void)launch_pdl(pack_jobs_kernel, dim3(blocks), dim3(256), (size_t)0, stream, *jobs);
    SINDDM_CUDA_OK(cudaGetLastError());
    return SINDDM_OK;
}

}  // namespace sinddm
