// Micro-benchmark of the epilogue's store path: how fast can one SM push staged tiles to an NHWC [P][C] tensor with
// TMA tensor stores, as a function of the box's inner (channel) extent?  tc_conv's epilogue stores (16 ch x 8 w x 4 h)
// boxes = 32 pieces of 64 B per store; this measures 64 / 128 / 256-byte pieces, 2 / 4 KiB boxes, 8 warps per CTA with
// two stores in flight per warp (wait_group.read 1), one CTA per SM -- no other work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tma_store_bench tools/tma_store_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../sinddm_b200/csrc/common.cuh"

using namespace sinddm;

// every warp owns rows [4 * warp_global .. +4) of an image-tile row band; walks tiles of (bw px x 4 rows)
__global__ void __launch_bounds__(256, 1)
store_bench(const __grid_constant__ CUtensorMap map, int box_c, int box_w, int C, int W, int H, int B, int inflight,
            int box_bytes, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* stg = smem + warp * 2 * box_bytes;
    for (int i = lane; i < 2 * box_bytes / 4; i += 32) reinterpret_cast<float*>(stg)[i] = (float)(i + warp);
    fence_proxy_async_smem();
    __syncthreads();
    const int cgroups = C / box_c;
    const int wt = (W + box_w - 1) / box_w, ht = H / 4;
    const long long boxes = (long long)B * ht * wt * cgroups;
    const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
    const long long t0 = clock64();
    int buf = 0;
    for (long long i = gw; i < boxes; i += nw) {
        const int cg = (int)(i % cgroups);
        long long r = i / cgroups;
        const int tw = (int)(r % wt);
        r /= wt;
        const int th = (int)(r % ht), b = (int)(r / ht);
        if (lane == 0) {
            if (inflight == 1) bulk_wait_group_read<0>();
            else if (inflight == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<3>();
        }
        __syncwarp();
        // (the real epilogue writes the tile here: four 16-byte stores per lane)
        sts_f4(smem_u32(stg + buf * box_bytes) + lane * 16, 1.f, 2.f, 3.f, (float)i);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_4d(&map, stg + buf * box_bytes, cg * box_c, tw * box_w, th * 4, b);
            bulk_commit_group();
        }
        buf ^= 1;
    }
    if (lane == 0) bulk_wait_group_read<0>();
    __syncwarp();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int B = 32, H = 184, W = 248, C = 160;
    const size_t n = (size_t)B * H * W * C;
    float* out;
    cudaMalloc(&out, n * 4);
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    cudaFuncSetAttribute(store_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8192 + 1024);
    struct Case { int box_c, box_w, inflight; };
    const Case cases[] = {{16, 8, 2}, {32, 8, 2}, {64, 8, 2}, {16, 16, 2}, {32, 4, 2}, {16, 8, 4}, {32, 8, 4}, {16, 8, 1}, {32, 8, 1}, {160, 2, 2}};
    printf("TMA tensor stores to NHWC [%d,%d,%d,%d] fp32 (%.0f MB), 148 CTAs x 8 warps, boxes (c, w, 4 rows)\n", B, H, W, C, n * 4 / 1e6);
    for (const Case& cs : cases) {
        CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)cs.box_c, (cuuint32_t)cs.box_w, 4, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int box_bytes = cs.box_c * cs.box_w * 4 * 4;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            store_bench<<<148, 256, 8 * 2 * 8192 + 1024>>>(map, cs.box_c, cs.box_w, C, W, H, B, cs.inflight, box_bytes, cyc);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        const double bytes = (double)B * H * ((W + cs.box_w - 1) / cs.box_w * cs.box_w > W ? W : W) * C * 4.0;
        printf("box %3d ch x %2d px x 4 rows (%4d B pieces, %5d B/box), %d in flight per warp: %.3f ms  %.2f TB/s  %s\n",
               cs.box_c, cs.box_w, cs.box_c * 4, box_bytes, cs.inflight, ms, bytes / ms / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
