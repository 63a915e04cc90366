// Micro-benchmark of the producer / MMA-warp handshake of tc_conv (no data movement): what do an mbarrier wait,
// a tcgen05.commit and the commit -> mbarrier -> producer -> mbarrier round trip cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench tools/pipe_bench.cu && /tmp/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>

#include "../sinddm_b200/csrc/common.cuh"

using namespace sinddm;

constexpr int kMaxRing = 16;

// mode 0: commits only | 1: ready waits only | 2: tcgen05 fences only | 3: x4 MMA (N) + commit per iteration
// mode 4: ring ping-pong, consumer = wait(full) + commit(empty), producer = wait(empty) + arrive(full)
// mode 5: like 4 with a x4 MMA (N) issued per iteration | 6: like 4 but the consumer arrives with a plain mbarrier.arrive
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, int depth, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[kMaxRing], empty[kMaxRing], done;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxRing; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(&done, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    const uint32_t hi = (uint32_t)(umma_smem_desc(0, 0, 1024, UMMA_LAYOUT_SW128) >> 32);
    const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 48 * 1024;
    if (warp == 1 && mode < 7) {            // "MMA warp"
        if (mode == 1) {        // make phase 0 of full[0] complete once
            if ((threadIdx.x & 31) == 0) mbar_arrive(&full[0]);
            __syncwarp();
        }
        const long long t0 = clock64();
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
            if (mode == 0) {
                umma_commit_elect(&empty[s]);
            } else if (mode == 1) {
                mbar_wait(&full[0], 0);
            } else if (mode == 2) {
                tc_fence_after_sync();
            } else if (mode == 3) {
                umma_tf32_ss_x4(tm, (sa >> 4) & 0x3FFF, (sb >> 4) & 0x3FFF, hi, 2, idesc, 1u, 4u);
                umma_commit_elect(&empty[s]);
            } else {
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                if (mode == 5) umma_tf32_ss_x4(tm, (sa >> 4) & 0x3FFF, (sb >> 4) & 0x3FFF, hi, 2, idesc, 1u, 4u);
                if (mode == 6) {
                    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
                    __syncwarp();
                } else {
                    umma_commit_elect(&empty[s]);
                }
            }
            if (++s == depth) {
                s = 0;
                ph ^= 1u;
            }
        }
        const long long t_issue = clock64();
        umma_commit_elect(&done);
        mbar_wait(&done, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
            out[0] = t1 - t0;
            out[1] = t_issue - t0;
        }
    } else if (warp == 2 && mode >= 7) {   // "epilogue warp" primitives
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 7) fence_proxy_async_smem();
            else if (mode == 8) { if ((threadIdx.x & 31) == 0) bulk_wait_group_read<0>(); __syncwarp(); }
            else if (mode == 9) { if ((threadIdx.x & 31) == 0) bulk_commit_group(); __syncwarp(); }
            else if (mode == 10) __syncwarp();
            else if (mode == 11) {      // the whole store sequence of one 2 KiB staging tile, without the TMA store itself
                if ((threadIdx.x & 31) == 0) bulk_wait_group_read<1>();
                __syncwarp();
                *reinterpret_cast<float4*>(smem + (threadIdx.x & 31) * 64) = make_float4(1.f, 2.f, 3.f, 4.f);
                fence_proxy_async_smem();
                __syncwarp();
                if ((threadIdx.x & 31) == 0) bulk_commit_group();
            } else if (mode == 12) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else if (mode == 13) {
                uint32_t r[16];
                tmem_ld16_issue(tm, r);
                tmem_ld16_wait(r);
                if (r[0] == 0x12345678u) out[3] = 1;
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
            out[0] = t1 - t0;
            out[1] = t1 - t0;
        }
    } else if (warp == 0 && mode >= 4 && mode < 7) {   // "producer"
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx_w(&full[s], 0);
            if (++s == depth) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc(tm, 512);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[14] = {"commit only", "ready mbarrier wait only", "tcgen05.fence::after only", "x4 MMA + commit",
                            "ring: wait + commit | wait + arrive", "ring: wait + x4 MMA + commit | wait + arrive",
                            "ring: wait + plain arrive | wait + arrive", "fence.proxy.async.shared::cta",
                            "cp.async.bulk.wait_group.read 0 (nothing pending)", "cp.async.bulk.commit_group (empty)",
                            "__syncwarp", "store sequence w/o TMA (wait_group, STS, fence, commit)", "tcgen05.wait::ld (nothing pending)",
                            "tcgen05.ld x16 + wait"};
    const int iters = 4096;
    cudaMalloc(&d, 64);
    for (int mode = 0; mode < 14; ++mode) {
        for (int N : {80, 160}) {
            if (N == 160 && !(mode == 3 || mode == 5)) continue;
            for (int depth : {3, 4, 8}) {
                if ((mode < 4 || mode >= 7) && depth != 8) continue;
                bench<<<148, 128, 100 * 1024>>>(mode, N, iters, depth, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long both[2] = {0, 0};
                cudaMemcpy(both, d, 16, cudaMemcpyDeviceToHost);
                printf("%-48s N=%3d ring=%d : issue %7.1f  total %7.1f cycles/iteration %s\n", names[mode], N, depth,
                       (double)both[1] / iters, (double)both[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
