// Shared device helpers for the sinddm_b200 kernels (sm_100a only).
//
// Everything here is hand-written PTX wrappers for the Blackwell async machinery
// (mbarrier, TMA bulk-tensor copies, tcgen05 MMA / TMEM) plus the few scalar math
// helpers (exact-erf GELU, TF32 rounding) every kernel shares.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef SINDDM_DEVINL
#define SINDDM_DEVINL __device__ __forceinline__
#endif

namespace sinddm {

// ----------------------------------------------------------------------------------------------
// scalar math
// ----------------------------------------------------------------------------------------------

// nn.GELU() default (approximate='none'): 0.5 x (1 + erf(x / sqrt(2)))  -- reference models.py:55,64,108
SINDDM_DEVINL float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// d/dx gelu(x) = Phi(x) + x * phi(x)
SINDDM_DEVINL float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Epilogue versions for the tensor-core (TF32) path: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e.
// at the level of erff's own rounding and four orders of magnitude below the TF32 operand rounding of the
// GEMM that feeds it).  The GELU epilogues are bound by instruction issue (ncu: 38 % of all warp instructions of a
// 160 -> 160 GELU layer were these lines), so everything is spelled out: bare ex2.approx / rcp.approx (no range
// fix-ups: the arguments are <= 0 resp. >= 1), 0.5 folded into the polynomial, and GELU itself without a select:
//   gelu(x) = x * Phi(x) = max(x, 0) - |x| * h(|x|),   h(a) = 0.5 * erfc(a / sqrt2) = t * p(t) * exp(-a^2 / 2)
// 13 instructions per value instead of ~25.
SINDDM_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
SINDDM_DEVINL float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// h = 0.5 * erfc(|x| / sqrt2) (the upper tail of the normal distribution), e = exp(-x^2 / 2)
SINDDM_DEVINL void phi_tail(float x, float* h, float* e) {
    const float ax = fabsf(x);
    *e = ex2_approx((x * -0.72134752044448170368f) * x);          // -0.5 * log2(e)
    const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    *h = (p * t) * *e;
}

SINDDM_DEVINL void phi_cdf_pdf(float x, float* cdf, float* pdf) {
    float h, e;
    phi_tail(x, &h, &e);
    *cdf = x >= 0.f ? 1.0f - h : h;
    *pdf = 0.39894228040143267794f * e;
}

SINDDM_DEVINL float gelu_fast(float x) {
    float h, e;
    phi_tail(x, &h, &e);
    return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}

SINDDM_DEVINL float gelu_grad_fast(float x) {
    float cdf, pdf;
    phi_cdf_pdf(x, &cdf, &pdf);
    return fmaf(x, pdf, cdf);
}

// Round-to-nearest (ties away from zero) fp32 -> tf32 (10-bit mantissa), returned as an fp32 bit pattern.  The tensor
// core truncates fp32 operands to tf32; rounding at the producer removes the truncation bias.  cvt.rna.tf32.f32
// compiles to three instructions; on the sign-magnitude bit pattern "add half an ulp, clear the low bits" is the
// same function in two (inf stays inf, the largest finite values round to inf like cvt.rna does).
SINDDM_DEVINL float round_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

SINDDM_DEVINL uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte shared-memory accesses by shared-window address (no generic-pointer arithmetic in the hot loops)
SINDDM_DEVINL void sts_f4(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// ld.VOLATILE: `asm volatile` only pins the statement for the front end; to ptxas a plain ld.shared is an ordinary
// load that it may sink towards its use or re-execute to save registers.  Tiles that the async proxy (TMA) overwrites
// behind ptxas's back must be read exactly once, where the program says so.
SINDDM_DEVINL float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// plain shared load for tiles only this warp's generic-proxy stores write (the "memory" clobber keeps program order)
SINDDM_DEVINL float lds_f32_plain(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

SINDDM_DEVINL float4 lds_f4_plain(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

SINDDM_DEVINL float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}

// True in exactly one lane of a fully converged warp.  tcgen05.mma / TMA / tcgen05.commit are issued through
// the uniform datapath: they must sit in WARP-UNIFORM control flow guarded by this predicate.  Guarding them
// with `lane == 0` instead makes the compiler wrap each one in an ELECT/branch loop over the active lanes,
// which costs ~200 cycles per instruction (measured with tools/mma_bench.cu).
SINDDM_DEVINL bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// Programmatic dependent launch (host_common.h: launch_pdl): wait until the previous kernel of the stream has completed
// and its memory is visible, then let the NEXT kernel's CTAs start as soon as this grid's CTAs retire.  Must precede the
// kernel's first global-memory access; everything before it (barrier init, TMEM allocation, descriptor prefetch, index
// arithmetic) overlaps the predecessor's tail.  A no-op when the kernel was launched without the attribute.
SINDDM_DEVINL void pdl_grid_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

SINDDM_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------

SINDDM_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

SINDDM_DEVINL void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

SINDDM_DEVINL void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

SINDDM_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

SINDDM_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

SINDDM_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// try_wait is a hardware-suspended wait with a bounded time slice; loop until the phase flips.
SINDDM_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) -- tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------------

SINDDM_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

SINDDM_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

SINDDM_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                               int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Brings a box into L2 only (no shared-memory destination, no completion tracking): later TMA loads of the same
// box hit L2 instead of paying the DRAM latency.
SINDDM_DEVINL void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// Tile store shared -> global (elements outside the tensor are clipped).  Completion is tracked per issuing
// thread in bulk async-groups: commit after the store, wait_group_read<N> before the staging buffer of all but
// the N most recent groups is overwritten.  The smem writes being stored need fence_proxy_async_smem() first.
SINDDM_DEVINL void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
SINDDM_DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
SINDDM_DEVINL void bulk_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Whole-warp ("_w") versions: executed by all 32 lanes of a converged warp, the lane election happens inside
// the asm by predication, so the surrounding loop stays in uniform control flow (see elect_one_sync()).
SINDDM_DEVINL void mbar_arrive_expect_tx_w(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(bytes)
        : "memory");
}

SINDDM_DEVINL void tma_load_2d_w(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

SINDDM_DEVINL void tma_load_4d_w(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                 int c3) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// 1-D bulk copy global -> shared of `bytes` contiguous bytes (multiple of 16, both addresses 16-byte aligned),
// completing on an mbarrier like a tensor copy.  One request stream of full lines instead of one 128-byte
// row per tensor-box row.
SINDDM_DEVINL void bulk_copy_g2s_w(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Multicast variant: the box is written at the same CTA-relative smem offset of every CTA in `cta_mask`
// and completes bytes on the mbarrier at the same offset in each of them.
SINDDM_DEVINL void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
        "[%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

SINDDM_DEVINL void tma_load_2d_mc_w(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                    uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
        "[%1, {%3, %4}], [%2], %5;\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// thread-block clusters
// ----------------------------------------------------------------------------------------------

SINDDM_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// all threads of all CTAs in the cluster
SINDDM_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM -> register loads
// ----------------------------------------------------------------------------------------------

// Whole-warp collective.  Writes the TMEM base address (lane 0, first column) to *smem_slot.
SINDDM_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

SINDDM_DEVINL void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}

SINDDM_DEVINL void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SINDDM_DEVINL void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], operands fp32 containers consumed as TF32, fp32 accumulate.
// Issued by ONE thread on behalf of the CTA.
SINDDM_DEVINL void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Up to four back-to-back tf32 MMAs along K issued from ONE asm statement executed by the WHOLE (converged)
// warp: the lane election and the per-K descriptor advance (+32 B = +2 in the >>4 encoded start address)
// happen inside, with predication instead of branches.  Issue cost per MMA drops from ~250 cycles (C++ loop
// with `if (elect)` around a single-MMA asm: convergence-barrier + R2UR chain per instruction, measured with
// tools/mma_bench.cu) to a few tens of cycles.
//   a_lo / b_lo : low 32 bits of the smem descriptors of the first K slice;  hi: shared high 32 bits
//   step_lo     : descriptor advance per K slice (2 for K-major 128B-swizzle rows, 64 for MN-major boxes)
//   nk          : number of K slices to issue (1..4);  acc_first: accumulate flag of the first one
SINDDM_DEVINL void umma_tf32_ss_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t step_lo,
                                   uint32_t idesc, uint32_t acc_first, uint32_t nk) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pone, p1, p2, p3;\n\t"
        ".reg .b32 rx, a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %6, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.gt.u32 p1, %7, 1;\n\t"
        "setp.gt.u32 p2, %7, 2;\n\t"
        "setp.gt.u32 p3, %7, 3;\n\t"
        "and.pred p1, p1, pe;\n\t"
        "and.pred p2, p2, pe;\n\t"
        "and.pred p3, p3, pe;\n\t"
        "add.u32 a1, %1, %4;\n\t"
        "add.u32 a2, a1, %4;\n\t"
        "add.u32 a3, a2, %4;\n\t"
        "add.u32 b1, %2, %4;\n\t"
        "add.u32 b2, b1, %4;\n\t"
        "add.u32 b3, b2, %4;\n\t"
        "mov.b64 da0, {%1, %3};\n\t"
        "mov.b64 da1, {a1, %3};\n\t"
        "mov.b64 da2, {a2, %3};\n\t"
        "mov.b64 da3, {a3, %3};\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da0, db0, %5, pacc;\n\t"
        "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%0], da1, db1, %5, pone;\n\t"
        "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%0], da2, db2, %5, pone;\n\t"
        "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%0], da3, db3, %5, pone;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(step_lo), "r"(idesc), "r"(acc_first), "r"(nk)
        : "memory");
}

// umma_tf32_ss_x4 that also TESTS another mbarrier (non-blocking) and returns the answer: the test is issued first and
// its predicate is only read after the MMAs have been issued, so the ~250-cycle round trip of the barrier unit
// (tools/pipe_bench.cu) overlaps the MMA issue instead of stalling the warp in front of it.  The result is per lane;
// make it warp-uniform (__all_sync) before branching on it.
SINDDM_DEVINL uint32_t umma_tf32_ss_x4_test(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t step_lo,
                                            uint32_t idesc, uint32_t acc_first, uint32_t nk, uint64_t* next_bar,
                                            uint32_t next_parity) {
    uint32_t ready;
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pone, p1, p2, p3, pnext;\n\t"
        ".reg .b32 rx, a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pnext, [%9], %10;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %7, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.gt.u32 p1, %8, 1;\n\t"
        "setp.gt.u32 p2, %8, 2;\n\t"
        "setp.gt.u32 p3, %8, 3;\n\t"
        "and.pred p1, p1, pe;\n\t"
        "and.pred p2, p2, pe;\n\t"
        "and.pred p3, p3, pe;\n\t"
        "add.u32 a1, %2, %5;\n\t"
        "add.u32 a2, a1, %5;\n\t"
        "add.u32 a3, a2, %5;\n\t"
        "add.u32 b1, %3, %5;\n\t"
        "add.u32 b2, b1, %5;\n\t"
        "add.u32 b3, b2, %5;\n\t"
        "mov.b64 da0, {%2, %4};\n\t"
        "mov.b64 da1, {a1, %4};\n\t"
        "mov.b64 da2, {a2, %4};\n\t"
        "mov.b64 da3, {a3, %4};\n\t"
        "mov.b64 db0, {%3, %4};\n\t"
        "mov.b64 db1, {b1, %4};\n\t"
        "mov.b64 db2, {b2, %4};\n\t"
        "mov.b64 db3, {b3, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%1], da0, db0, %6, pacc;\n\t"
        "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%1], da1, db1, %6, pone;\n\t"
        "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%1], da2, db2, %6, pone;\n\t"
        "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%1], da3, db3, %6, pone;\n\t"
        "selp.u32 %0, 1, 0, pnext;\n\t"
        "}\n"
        : "=r"(ready)
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(step_lo), "r"(idesc), "r"(acc_first), "r"(nk),
          "r"(smem_u32(next_bar)), "r"(next_parity)
        : "memory");
    return ready;
}

// Variants with separate high descriptor words for A (hi) and B (hi_b): operands whose core-group strides differ.
SINDDM_DEVINL void umma_tf32_ss_x4_h2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t hi_b, uint32_t step_lo,
                                   uint32_t idesc, uint32_t acc_first, uint32_t nk) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pone, p1, p2, p3;\n\t"
        ".reg .b32 rx, a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %6, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.gt.u32 p1, %7, 1;\n\t"
        "setp.gt.u32 p2, %7, 2;\n\t"
        "setp.gt.u32 p3, %7, 3;\n\t"
        "and.pred p1, p1, pe;\n\t"
        "and.pred p2, p2, pe;\n\t"
        "and.pred p3, p3, pe;\n\t"
        "add.u32 a1, %1, %4;\n\t"
        "add.u32 a2, a1, %4;\n\t"
        "add.u32 a3, a2, %4;\n\t"
        "add.u32 b1, %2, %4;\n\t"
        "add.u32 b2, b1, %4;\n\t"
        "add.u32 b3, b2, %4;\n\t"
        "mov.b64 da0, {%1, %3};\n\t"
        "mov.b64 da1, {a1, %3};\n\t"
        "mov.b64 da2, {a2, %3};\n\t"
        "mov.b64 da3, {a3, %3};\n\t"
        "mov.b64 db0, {%2, %8};\n\t"
        "mov.b64 db1, {b1, %8};\n\t"
        "mov.b64 db2, {b2, %8};\n\t"
        "mov.b64 db3, {b3, %8};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da0, db0, %5, pacc;\n\t"
        "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%0], da1, db1, %5, pone;\n\t"
        "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%0], da2, db2, %5, pone;\n\t"
        "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%0], da3, db3, %5, pone;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(step_lo), "r"(idesc), "r"(acc_first), "r"(nk), "r"(hi_b)
        : "memory");
}

SINDDM_DEVINL uint32_t umma_tf32_ss_x4_test_h2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t hi_b, uint32_t step_lo,
                                            uint32_t idesc, uint32_t acc_first, uint32_t nk, uint64_t* next_bar,
                                            uint32_t next_parity) {
    uint32_t ready;
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pone, p1, p2, p3, pnext;\n\t"
        ".reg .b32 rx, a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pnext, [%9], %10;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %7, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.gt.u32 p1, %8, 1;\n\t"
        "setp.gt.u32 p2, %8, 2;\n\t"
        "setp.gt.u32 p3, %8, 3;\n\t"
        "and.pred p1, p1, pe;\n\t"
        "and.pred p2, p2, pe;\n\t"
        "and.pred p3, p3, pe;\n\t"
        "add.u32 a1, %2, %5;\n\t"
        "add.u32 a2, a1, %5;\n\t"
        "add.u32 a3, a2, %5;\n\t"
        "add.u32 b1, %3, %5;\n\t"
        "add.u32 b2, b1, %5;\n\t"
        "add.u32 b3, b2, %5;\n\t"
        "mov.b64 da0, {%2, %4};\n\t"
        "mov.b64 da1, {a1, %4};\n\t"
        "mov.b64 da2, {a2, %4};\n\t"
        "mov.b64 da3, {a3, %4};\n\t"
        "mov.b64 db0, {%3, %11};\n\t"
        "mov.b64 db1, {b1, %11};\n\t"
        "mov.b64 db2, {b2, %11};\n\t"
        "mov.b64 db3, {b3, %11};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%1], da0, db0, %6, pacc;\n\t"
        "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%1], da1, db1, %6, pone;\n\t"
        "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%1], da2, db2, %6, pone;\n\t"
        "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%1], da3, db3, %6, pone;\n\t"
        "selp.u32 %0, 1, 0, pnext;\n\t"
        "}\n"
        : "=r"(ready)
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(step_lo), "r"(idesc), "r"(acc_first), "r"(nk),
          "r"(smem_u32(next_bar)), "r"(next_parity), "r"(hi_b)
        : "memory");
    return ready;
}

// Whole-warp version of umma_commit: one elected lane arrives.
SINDDM_DEVINL void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe;\n\t"
        ".reg .b32 rx;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
SINDDM_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Same, arriving on the mbarrier at this CTA-relative offset in every CTA of `cta_mask` (used when the smem
// stage being released is also written by the peers' multicast TMA).
SINDDM_DEVINL void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

SINDDM_DEVINL void umma_commit_mc_elect(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::
            "r"(smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
SINDDM_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a cluster run ONE M=256 MMA; each supplies its own 128 rows
// of A and half of the rows of B from its own shared memory and keeps its 128 accumulator rows in its own
// TMEM.  Only the leader CTA (cluster rank 0) issues MMAs; TMA loads of both CTAs complete on the leader's
// mbarrier (same smem offset, peer bit cleared), commits multicast to both CTAs' barriers.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of "the same offset in the even CTA"

SINDDM_DEVINL void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

SINDDM_DEVINL void tmem_dealloc_2sm(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}

SINDDM_DEVINL void tma_load_2d_2sm_w(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
        "[%1, {%3, %4}], [%2];\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}

SINDDM_DEVINL void tma_load_4d_2sm_w(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                     int c3) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
        "[%1, {%3, %4, %5, %6}], [%2];\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// arrive (no transaction bytes) on the barrier at this smem offset in CTA `cta` of the cluster
SINDDM_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

SINDDM_DEVINL void mbar_arrive_remote_w(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx, ra;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "@pe mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// cluster-scope acquire wait (the barrier is arrived on by the peer CTA)
SINDDM_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

SINDDM_DEVINL void umma_tf32_ss_x4_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t step_lo,
                                       uint32_t idesc, uint32_t acc_first, uint32_t nk) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pone, p1, p2, p3;\n\t"
        ".reg .b32 rx, a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "elect.sync rx|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %6, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.gt.u32 p1, %7, 1;\n\t"
        "setp.gt.u32 p2, %7, 2;\n\t"
        "setp.gt.u32 p3, %7, 3;\n\t"
        "and.pred p1, p1, pe;\n\t"
        "and.pred p2, p2, pe;\n\t"
        "and.pred p3, p3, pe;\n\t"
        "add.u32 a1, %1, %4;\n\t"
        "add.u32 a2, a1, %4;\n\t"
        "add.u32 a3, a2, %4;\n\t"
        "add.u32 b1, %2, %4;\n\t"
        "add.u32 b2, b1, %4;\n\t"
        "add.u32 b3, b2, %4;\n\t"
        "mov.b64 da0, {%1, %3};\n\t"
        "mov.b64 da1, {a1, %3};\n\t"
        "mov.b64 da2, {a2, %3};\n\t"
        "mov.b64 da3, {a3, %3};\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da0, db0, %5, pacc;\n\t"
        "@p1 tcgen05.mma.cta_group::2.kind::tf32 [%0], da1, db1, %5, pone;\n\t"
        "@p2 tcgen05.mma.cta_group::2.kind::tf32 [%0], da2, db2, %5, pone;\n\t"
        "@p3 tcgen05.mma.cta_group::2.kind::tf32 [%0], da3, db3, %5, pone;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(step_lo), "r"(idesc), "r"(acc_first), "r"(nk)
        : "memory");
}

// leader CTA: arrive on the barrier at this offset in both CTAs of the pair once all prior MMAs are done
SINDDM_DEVINL void umma_commit_2sm_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b32 rx;\n\telect.sync rx|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::
            "r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// Split version for software pipelining: issue the load, do other work, then wait.  The wait takes the
// destination registers as in/out operands so the compiler cannot move their first use above it.
SINDDM_DEVINL void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

SINDDM_DEVINL void tmem_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100 "version 1" shared-memory matrix descriptor and instruction descriptor)
// ----------------------------------------------------------------------------------------------

enum : uint32_t {
    UMMA_LAYOUT_SW128 = 2,       // 128-byte swizzle, 16-byte atoms  (K-major operands)
    UMMA_LAYOUT_SW128_B32 = 1,   // 128-byte swizzle, 32-byte atoms  (the only layout for MN-major tf32)
};

// start address / leading-dim byte offset / stride-dim byte offset are all encoded >> 4.
SINDDM_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version = 1 (Blackwell)
    d |= static_cast<uint64_t>(layout & 0x7u) << 61;
    return d;
}

// kind::tf32, fp32 accumulator.  a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ inline uint32_t umma_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;            // c_format  = F32
    d |= 2u << 7;            // a_format  = TF32
    d |= 2u << 10;           // b_format  = TF32
    d |= (a_mn & 1u) << 15;  // a_major
    d |= (b_mn & 1u) << 16;  // b_major
    d |= ((N >> 3) & 0x3Fu) << 17;
    d |= ((M >> 4) & 0x1Fu) << 24;
    return d;
}

}  // namespace sinddm
