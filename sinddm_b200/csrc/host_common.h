// Host-side helpers shared by the launchers: error reporting, launch checks, TMA descriptor encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sinddm {

// Error codes returned through the C ABI (include/sinddm_b200.h mirrors these).
enum {
    SINDDM_OK = 0,
    SINDDM_ERR_INVALID = -1,   // bad argument / unsupported shape
    SINDDM_ERR_CUDA = -2,      // CUDA runtime / driver error
    SINDDM_ERR_NOT_INIT = -3,  // sinddm_init() was not called
    SINDDM_ERR_WORKSPACE = -4, // workspace too small / misaligned
};

void set_error(const char* fmt, ...);
const char* last_error();

#define SINDDM_CUDA_OK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        ::sinddm::note_call(#expr);                                                            \
        if (_e != cudaSuccess) {                                                               \
            ::sinddm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ::sinddm::SINDDM_ERR_CUDA;                                                  \
        }                                                                                      \
    } while (0)

#define SINDDM_REQUIRE(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::sinddm::set_error(__VA_ARGS__);     \
            return ::sinddm::SINDDM_ERR_INVALID;  \
        }                                         \
    } while (0)

#define SINDDM_TRY(expr)             \
    do {                             \
        int _rc = (expr);            \
        if (_rc != 0) return _rc;    \
    } while (0)

// Per-process state set up by sinddm_init(): SM count, smem opt-in limit, driver entry point.
struct DeviceInfo {
    int initialized;
    int device;
    int num_sms;
    int max_smem_optin;
};
const DeviceInfo& device_info();
int init_device(int device);

// [B,H,W,C] fp32 NHWC tensor -> 4-D tiled TMA descriptor, box = (box_c, box_w, box_h, 1).
// Out-of-bounds box elements (negative or past-the-end coordinates, channel tails) are zero-filled.
int make_tmap_nhwc(CUtensorMap* out, const float* base, int B, int H, int W, int C, int box_c, int box_w, int box_h,
                   CUtensorMapSwizzle swizzle);

// [rows, inner] fp32 row-major matrix -> 2-D tiled TMA descriptor, box = (box_inner, box_rows).
int make_tmap_2d(CUtensorMap* out, const float* base, int inner, int rows, int box_inner, int box_rows,
                 CUtensorMapSwizzle swizzle);

// Every kernel launch is followed by SINDDM_CUDA_OK(cudaGetLastError()); note_call counts those, which
// gives bench.py its `gpu_launches` figure (sinddm_launch_count in the C ABI).
void note_call(const char* expr);
unsigned long long launch_count();

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's live roofline numbers).
// kind: 0 = tc_conv (forward / data gradient), 1 = tc_wgrad.  Off by default; zero cost when off.
void prof_enable(int on);
void prof_begin(cudaStream_t stream, int kind, double flops);
void prof_end(cudaStream_t stream);
int prof_collect(int kind, double* total_ms, double* total_flops, int* launches);

// Programmatic dependent launch: a kernel launched through launch_pdl may start (its CTAs become resident, run their
// prologue: barrier init, TMEM allocation, descriptor prefetch) while the previous kernel of the stream is still
// draining its last CTAs.  EVERY kernel launched this way executes pdl_grid_sync() (common.cuh: griddepcontrol.wait +
// griddepcontrol.launch_dependents) before its first access to global memory, so data dependencies between consecutive
// kernels hold exactly as with ordinary stream order.  ~95 launches per training step and ~35 per sampling step each
// save their prologue / the predecessor's tail (SINDDM_PDL=0 restores plain launches for A/B runs).
// (measured, tools/ab_probe.py: -6 % / -3 % / -1.5 % per training step at 98k / 193k / 379k pixels, +1 % beyond: the
// network driver switches it on per call for problems below pdl_max_pixels(); it is off everywhere else)
bool pdl_enabled();
void pdl_set(bool on);            // per-call switch used by net_forward / net_backward (single host thread per device)
long long pdl_max_pixels();       // SINDDM_PDL_MAX_PX, default 400000; SINDDM_PDL=0 disables PDL altogether
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace sinddm
