// Error slot, device init and TMA descriptor encoding.
#include "host_common.h"

#include <stdlib.h>

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace sinddm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char* last_error() { return g_err; }

static unsigned long long g_launches = 0;

namespace {
bool g_pdl_call = false;   // off outside the network driver's small-problem scopes
}
bool pdl_enabled() {
    static const bool on = []() {
        const char* e = getenv("SINDDM_PDL");
        return !(e && atoi(e) == 0);
    }();
    return on && g_pdl_call;
}
void pdl_set(bool on) { g_pdl_call = on; }
long long pdl_max_pixels() {
    static const long long v = []() {
        const char* e = getenv("SINDDM_PDL_MAX_PX");
        return e ? atoll(e) : 400000ll;
    }();
    return v;
}

void note_call(const char* expr) {
    // "cudaGetLastError()" is what follows every <<<>>> launch in this library
    if (expr[0] == 'c' && expr[4] == 'G' && expr[7] == 'L') ++g_launches;
}

unsigned long long launch_count() { return g_launches; }

namespace {
struct ProfEntry {
    cudaEvent_t a, b;
    double flops;
    int kind;
};
constexpr int kProfMax = 8192;
ProfEntry g_prof[kProfMax];
int g_prof_n = 0, g_prof_made = 0, g_prof_on = 0, g_prof_open = -1;
}  // namespace

void prof_enable(int on) {
    g_prof_on = on;
    if (on) g_prof_n = 0;
}

void prof_begin(cudaStream_t stream, int kind, double flops) {
    g_prof_open = -1;
    if (!g_prof_on || g_prof_n >= kProfMax) return;
    if (g_prof_n >= g_prof_made) {
        if (cudaEventCreate(&g_prof[g_prof_made].a) != cudaSuccess) return;
        if (cudaEventCreate(&g_prof[g_prof_made].b) != cudaSuccess) return;
        ++g_prof_made;
    }
    g_prof_open = g_prof_n++;
    g_prof[g_prof_open].flops = flops;
    g_prof[g_prof_open].kind = kind;
    cudaEventRecord(g_prof[g_prof_open].a, stream);
}

void prof_end(cudaStream_t stream) {
    if (g_prof_open >= 0) cudaEventRecord(g_prof[g_prof_open].b, stream);
    g_prof_open = -1;
}

int prof_collect(int kind, double* total_ms, double* total_flops, int* launches) {
    double ms = 0, fl = 0;
    int n = 0;
    for (int i = 0; i < g_prof_n; ++i) {
        if (g_prof[i].kind != kind) continue;
        SINDDM_CUDA_OK(cudaEventSynchronize(g_prof[i].b));
        float t = 0.f;
        SINDDM_CUDA_OK(cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b));
        ms += t;
        fl += g_prof[i].flops;
        ++n;
    }
    *total_ms = ms;
    *total_flops = fl;
    *launches = n;
    return SINDDM_OK;
}

static DeviceInfo g_dev = {0, -1, 0, 0};
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

const DeviceInfo& device_info() { return g_dev; }

int init_device(int device) {
    SINDDM_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SINDDM_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("sinddm_b200 is built for sm_100a only; device %d reports sm_%d%d", device, prop.major, prop.minor);
        return SINDDM_ERR_INVALID;
    }
    if (!g_encode) {
        // libcuda is resolved at run time through the runtime so the .so links without a driver present.
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SINDDM_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return SINDDM_ERR_CUDA;
        }
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    g_dev.device = device;
    g_dev.num_sms = prop.multiProcessorCount;
    g_dev.max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    g_dev.initialized = 1;
    return SINDDM_OK;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    if (!g_encode) {
        set_error("sinddm_init() must be called before any tensor-core entry point");
        return SINDDM_ERR_NOT_INIT;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
        set_error("TMA base pointer %p is not 16-byte aligned", base);
        return SINDDM_ERR_INVALID;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu/%llu, box %u/%u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return SINDDM_ERR_CUDA;
    }
    return SINDDM_OK;
}

int make_tmap_nhwc(CUtensorMap* out, const float* base, int B, int H, int W, int C, int box_c, int box_w, int box_h,
                   CUtensorMapSwizzle swizzle) {
    if ((C * 4) % 16 != 0) {
        set_error("NHWC TMA descriptor needs C %% 4 == 0 (got C=%d)", C);
        return SINDDM_ERR_INVALID;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    return encode(out, base, 4, dims, strides, box, swizzle);
}

int make_tmap_2d(CUtensorMap* out, const float* base, int inner, int rows, int box_inner, int box_rows,
                 CUtensorMapSwizzle swizzle) {
    if ((inner * 4) % 16 != 0) {
        set_error("2-D TMA descriptor needs inner %% 4 == 0 (got %d)", inner);
        return SINDDM_ERR_INVALID;
    }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)inner * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    return encode(out, base, 2, dims, strides, box, swizzle);
}

}  // namespace sinddm
