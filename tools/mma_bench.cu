// Micro-benchmark: issue-to-completion cost of back-to-back tcgen05.mma (SS operands) for the shapes the
// conv / wgrad kernels use.  One CTA per SM, operands are whatever is in shared memory (values irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bench tools/mma_bench.cu && /tmp/mma_bench
#include <cstdio>
#include <cuda_runtime.h>

#include "../sinddm_b200/csrc/common.cuh"

using namespace sinddm;

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

// mode 0: tf32 K-major SW128 (conv), 1: tf32 MN-major SW128_B32 (wgrad), 2: bf16 K-major SW128
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, int distinct, int elect, int nacc, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (elect ? (warp == 1) : (threadIdx.x == 32)) {
        // elect = 1: the whole warp runs the loop (uniform control flow), one elected lane issues
        // elect = 0: a single thread runs the loop (divergent)
        const uint32_t sa = smem_u32(smem);
        const uint32_t sb = sa + 64 * 1024;
        uint32_t idesc;
        if (mode == 0) idesc = umma_idesc_tf32(128, N, 0, 0);
        else if (mode == 1) idesc = umma_idesc_tf32(128, N, 1, 1);
        else idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const long long t0 = clock64();
        if (elect == 2) {
            // whole warp, one asm block per 4 MMAs (umma_tf32_ss_x4)
            const uint32_t hi = (uint32_t)(umma_smem_desc(0, 0, 1024, UMMA_LAYOUT_SW128) >> 32);
            for (int i = 0; i < iters; i += 4) {
                const uint32_t dst = tm + (uint32_t)(((i >> 2) & (nacc - 1)) * 256);
                const uint32_t off = (uint32_t)((i >> 2) & 3) * 16384u;
                umma_tf32_ss_x4(dst, ((sa + off) >> 4) & 0x3FFF, ((sb + off) >> 4) & 0x3FFF, hi, 2, idesc, 1u, 4u);
            }
        } else
        for (int i = 0; i < iters; ++i) {
            const int d = i & 3;   // walk over 4 different operand slices
            uint64_t da, db;
            if (mode == 1) {
                da = umma_smem_desc(sa + d * 1024, 4096, 512, UMMA_LAYOUT_SW128_B32);
                db = umma_smem_desc(sb + d * 1024, 4096, 512, UMMA_LAYOUT_SW128_B32);
            } else {
                da = umma_smem_desc(sa + (d & 3) * 32 + (d >> 2) * 16384, 0, 1024, UMMA_LAYOUT_SW128);
                db = umma_smem_desc(sb + (d & 3) * 32 + (d >> 2) * 20480, 0, 1024, UMMA_LAYOUT_SW128);
            }
            const uint32_t dst = tm + (uint32_t)((i & (nacc - 1)) * 256);   // rotate over `nacc` accumulators
            if (!elect || elect_one_sync()) {
                if (mode == 2) umma_f16_ss(dst, da, db, idesc, 1u);
                else umma_tf32_ss(dst, da, db, idesc, 1u);
            }
        }
        const long long t_issue = clock64();
        if (elect == 2) umma_commit_elect(&bar);
        else if (!elect || elect_one_sync()) umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
            out[0] = t1 - t0;
            out[1] = t_issue - t0;
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
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[3] = {"tf32 K-major  (K=8/instr)", "tf32 MN-major (K=8/instr)", "bf16 K-major  (K=16/instr)"};
    for (int grid : {148}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int N : {80, 160, 256}) {
                for (int elect : {1, 2}) {
                    const int distinct = 12, nacc = 2;
                    if (elect == 2 && mode != 0) continue;
                    if (mode == 1 && N == 256) continue;
                    const int iters = 4096;
                    bench<<<grid, 128, 200 * 1024>>>(mode, N, iters, distinct, elect, nacc, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long both[2] = {0, 0};
                    cudaMemcpy(both, d, 16, cudaMemcpyDeviceToHost);
                    const long long cyc = both[0];
                    const double per = (double)cyc / iters;
                    const double kk = mode == 2 ? 16 : 8;
                    printf("grid %3d  %s  M=128 N=%3d  elect=%d : issue %6.1f  total %7.1f cycles/MMA  (%5.0f MAC/cycle/SM)  %s\n", grid,
                           names[mode], N, elect, (double)both[1] / iters, per, 128.0 * N * kk / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
            }
        }
    }
    return 0;
}
