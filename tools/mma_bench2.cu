// Micro-benchmark of cta_group::2 tf32 MMAs (cluster of 2 CTAs, leader issues): cycles per M=256 MMA.
#include <cstdio>
#include <cuda_runtime.h>
#include "../sinddm_b200/csrc/common.cuh"
using namespace sinddm;

__global__ void __launch_bounds__(128, 1) bench2(int N, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc_2sm(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    const int crank = (int)cluster_ctarank();
    if (warp == 1 && crank == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
        const uint32_t idesc = umma_idesc_tf32(256, N, 0, 0);
        const uint32_t hi = (uint32_t)(umma_smem_desc(0, 0, 1024, UMMA_LAYOUT_SW128) >> 32);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            const uint32_t dst = tm + (uint32_t)(((i >> 2) & 1) * 256);
            const uint32_t off = (uint32_t)((i >> 2) & 3) * 16384u;
            umma_tf32_ss_x4_2sm(dst, ((sa + off) >> 4) & 0x3FFF, ((sb + off) >> 4) & 0x3FFF, hi, 2, idesc, 1u, 4u);
        }
        const long long t_issue = clock64();
        umma_commit_2sm_elect(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
    } else if (warp == 1) {
        mbar_wait(&bar, 0);   // peer waits for the multicast commit
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_2sm(tm, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {80, 160, 256}) {
        const int iters = 4096;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench2, N, iters, d);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long both[2] = {0, 0};
        cudaMemcpy(both, d, 16, cudaMemcpyDeviceToHost);
        printf("cta_group::2 tf32 M=256 N=%3d : issue %6.1f total %6.1f cycles/MMA (%5.0f MAC/cycle/SM) %s %s\n", N,
               (double)both[1] / iters, (double)both[0] / iters, 256.0 * N * 8 / ((double)both[0] / iters) / 2,
               cudaGetErrorString(e), cudaGetErrorString(e2));
    }
    return 0;
}
