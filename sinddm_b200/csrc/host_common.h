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

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace sinddm
