// Shared helpers for the popcorn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/popcorn_b200.h"

namespace pc {

void set_error(const char* fmt, ...);
void count_launch();

// optional per-kernel CUDA-event timing (pc_profile_*): a scope records start/stop events on the launch stream
int prof_register(const char* name);
struct ProfScope {
    int slot; cudaStream_t st;
    ProfScope(int cat, cudaStream_t stream, double units = 0.0);
    ~ProfScope();
};

#define PC_CHECK_ARG(cond, msg)                           \
    do {                                                  \
        if (!(cond)) {                                    \
            pc::set_error("%s: %s", __func__, msg);       \
            return PC_ERR_INVALID;                        \
        }                                                 \
    } while (0)

#define PC_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pc::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e));      \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

#define PC_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        pc::count_launch();                                                                   \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            pc::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e));  \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

int num_sms();

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit loads / stores (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ld_stream4i(const int* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// packed fp32x2 FMA (sm_100+): d.lo = a.lo*b.lo + d.lo ; d.hi = a.hi*b.hi + d.hi
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

}  // namespace pc
