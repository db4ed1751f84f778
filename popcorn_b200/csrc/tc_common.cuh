// tcgen05 / TMEM / mbarrier primitives shared by the tensor-core kernels (head_tc.cu, conv_tc.cu), sm_100a only.
//
// Conventions used by both kernels:
//   * UMMA M = 128 (one TMEM lane per pixel), cta_group::1, kind::f16 (default) or kind::tf32 with fp32 accumulation in TMEM;
//   * the A operand (activations) lives in TMEM and is written with tcgen05.st by the thread that owns the lane;
//   * the B operand (weights) lives in shared memory, pre-split (hi/lo) and pre-swizzled on the host into the
//     canonical K-major SWIZZLE_128B layout: 32-float K-atoms, 128 B per row, 8-row groups 1024 B apart;
//   * error-compensated split operands: x = hi + lo (fp16 pairs, or TF32 with hi = top 19 bits), D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
#pragma once
#include "common.cuh"

// Operand format of the split products D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (both tensor-core kernels):
//   0: TF32 halves (kind::tf32, K = 8 per UMMA, one TMEM column per A element)
//   1: fp16 halves (kind::f16, K = 16 per UMMA, two A elements per TMEM column): the same 11 + 11 significand bits in half the
//      tcgen05.st bytes and half the UMMAs (conv Cin = 8: two thirds, its K = 24 pads to 32); accuracy through the whole path:
//      tools/precision_split_study.py / profiles/r2_split_precision_study.txt
#ifndef PC_TC_F16
#define PC_TC_F16 1
#endif
#ifndef PC_SPLIT_F32X2
#define PC_SPLIT_F32X2 1     // residuals of a pair with one fma.rn.f32x2 (head loop 691 -> 675 instructions, conv stager 166 -> 154; bit-identical)
#endif

namespace pc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp.  UMMAs must be issued under elect.sync: ptxas then knows the region is
// single-lane and feeds UTCHMMA from uniform registers directly; under a plain `tid == 0` branch it wraps EVERY
// tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~10 instructions, ~50 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a value the compiler knows to be warp-uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// instruction descriptor, kind::tf32: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major A and B,
// N>>3 at bits 17-22, M>>4 at bits 24-28   (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// kind::f16 with fp16 operands (A = B = F16: format 0), fp32 accumulation: K = 16 per instruction at the rate kind::tf32 has for K = 8
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between 8-row groups |
// version 1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// D[tmem] (+)= A[tmem] * B[smem desc]^T ; accumulate == 0 overwrites D
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// the same with fp16 operands: a 32-bit TMEM column of A holds K elements 2c (low half) and 2c+1 (high half)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
#if PC_TC_F16
#define umma_ts umma_f16_ts
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t M, uint32_t N) { return umma_idesc_f16(M, N); }
#else
#define umma_ts umma_tf32_ts
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t M, uint32_t N) { return umma_idesc_tf32(M, N); }
#endif
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tbase, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void mbar_init1(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// one non-blocking probe of an mbarrier phase; the result can be consumed much later, which hides the ~100-cycle
// round trip of the synchronisation unit behind other work (software-pipelined waits)
__device__ __forceinline__ uint32_t mbar_test(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(mbar), "r"(parity)
                 : "memory");
    return done;
}
// mbarrier.try_wait suspends the thread in hardware for a bounded time instead of busy-polling shared memory
#ifndef PC_WAIT_NS
#define PC_WAIT_NS 0                 // > 0: back off with nanosleep between failed probes (A/B: power / clocks vs hand-off latency)
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint32_t mbar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
        if (done) return;
        if (PC_WAIT_NS > 0) __nanosleep(PC_WAIT_NS);
        if (spin > (1u << 24)) __trap();     // a protocol mistake must fault, never hang the GPU
    }
}

// Wait for THREE barrier phases with one round trip: the probes are issued back to back, so their ~200-cycle latencies
// through the synchronisation unit overlap instead of adding up (pass a barrier twice when fewer are needed).
__device__ __forceinline__ void mbar_wait3_sleep(uint32_t m1, uint32_t p1, uint32_t m2, uint32_t p2, uint32_t m3, uint32_t p3) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred q1, q2, q3;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q1, [%1], %2;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q2, [%3], %4;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q3, [%5], %6;\n\t"
            "and.pred q1, q1, q2;\n\tand.pred q1, q1, q3;\n\tselp.u32 %0, 1, 0, q1;\n\t}"
            : "=r"(done)
            : "r"(m1), "r"(p1), "r"(m2), "r"(p2), "r"(m3), "r"(p3)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) __trap();
    }
}

// bounded spin on an mbarrier phase: a descriptor mistake must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

// x = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// (x0, x1) = hi + lo with fp16 halves packed as f16x2 (x0 in the low half): 11 + 11 significand bits like the TF32 split, in half the
// operand bytes.  fp16 keeps 5 exponent bits:
//   * lo parts below 6e-5 are subnormal: absolute error <= 3e-8 (tools/precision_split_study.py);
//   * 65504 < |x| <= 131008: hi saturates at 65504 and lo carries the rest (relative error <= 2^-12 (|x| - 65504) / |x|);
//   * |x| > 131008: lo overflows to inf ON PURPOSE (plain cvt.rn, not .satfinite) — the pixel comes out inf / NaN instead of silently
//     clipped.  Such activations need the TF32 build (-DPC_TC_F16=0); BN-normalised networks stay orders of magnitude below.
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
#if PC_SPLIT_F32X2
    // residuals of both elements with ONE packed fp32 instruction: (r0, r1) = (h0, h1) * (-1, -1) + (x0, x1)
    unsigned long long xx, hh, rr;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(hh) : "f"(h0), "f"(h1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(hh), "l"(0xBF800000BF800000ull), "l"(xx));
    float r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rr));
#else
    const float r0 = x0 - h0, r1 = x1 - h1;
#endif
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

}  // namespace pc
