// One-kernel gradient all-reduce (NVLink peer loads) + Adam + EMA (fused_optim.cu).
#pragma once

#include "host_common.h"

namespace sinddm {

constexpr int kFusedMaxWorld = 8;   // GPUs of one NVSwitch box

struct FusedStepDesc {
    int world, rank;
    long long n;                           // elements per bucket (multiple of 4)
    const float* grads[kFusedMaxWorld];    // rank r's gradient bucket of this step, mapped into this process
    uint32_t* flags[kFusedMaxWorld];       // rank r's flag array [world] (flags[dst][src] = epoch), peer mapped
    uint32_t epoch;                        // strictly increasing per step, starts at 1
    float* param;                          // this rank's flat parameters / Adam moments / EMA parameters
    float* exp_avg;
    float* exp_avg_sq;
    float* ema;                            // may be null when ema_mode == 0
    float beta1, beta2, eps;
    float step_size;                       // lr / (1 - beta1^t)
    float bias2_sqrt;                      // sqrt(1 - beta2^t)
    int ema_mode;                          // 0: leave, 1: ema = param, 2: ema = ema * beta + (1 - beta) * param
    float ema_beta;
    unsigned long long* wait_ns;           // optional: [0] += ns CTA 0 waited in the barrier, [1] = max (rank skew)
    const float* mc_grads;                 // optional: multicast address of this step's bucket (in-switch reduction)
};

int fused_step_launch(const FusedStepDesc& d, cudaStream_t stream);

}  // namespace sinddm
