// Standalone probe: sustained tcgen05.st / tcgen05.ld bandwidth of one SM as a function of the number of issuing warps and the
// instruction width (32x32b .x8 / .x16 / .x32), one CTA per SM.  (conv_tc.cu's stagers sustain ~125 B/clk with 8 warps.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I popcorn_b200/csrc -o tools/probe/tmem_bw_probe tools/probe/tmem_bw_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

#include "tc_common.cuh"

using namespace pc;

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// mode 0: st .x8, 1: st .x16, 2: st .x32, 3: ld .x16
__global__ void __launch_bounds__(1024) probe(int mode, int iters, long long* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&slot);
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = (uint32_t)((warp >> 2) * 64) & 511u;     // warps of one lane quarter use different column blocks
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int c = 0; c < 64; c += 8) tmem_st8(tbase + lane_off + col0 + c, reinterpret_cast<uint32_t(&)[8]>(r[c & 31]));
        } else if (mode == 1) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) tmem_st16(tbase + lane_off + col0 + c, reinterpret_cast<uint32_t(&)[16]>(r[c & 31]));
        } else if (mode == 2) {
#pragma unroll
            for (int c = 0; c < 64; c += 32) st32(tbase + lane_off + col0 + c, r);
        } else if (mode == 3) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) tmem_ld16(tbase + lane_off + col0 + c, reinterpret_cast<uint32_t(&)[16]>(r[c & 31]));
        } else {      // mode 4: the stagers' pattern — three .x16 stores, then tcgen05.wait::st, every iteration (48 columns)
#pragma unroll
            for (int c = 0; c < 48; c += 16) tmem_st16(tbase + lane_off + col0 + c, reinterpret_cast<uint32_t(&)[16]>(r[c & 31]));
            tc_wait_st();
        }
    }
    if (mode == 3) tc_wait_ld(); else tc_wait_st();
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (mode == 3 && r[0] == 0x7fffffff) out[1] = r[5];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const int iters = 2000;
    printf("mode warps  clk  bytes/clk/SM\n");
    for (int mode = 0; mode < 5; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            probe<<<1, warps * 32, 0>>>(mode, iters, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long c;
            cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * (mode == 4 ? 48 : 64) * 4;
            printf("%s %3d %9lld %8.1f   %.1f clk per iteration\n", mode == 0 ? "st.x8 " : mode == 1 ? "st.x16" : mode == 2 ? "st.x32" : mode == 3 ? "ld.x16" : "3xst.x16+wait", warps, c,
                   bytes / c, (double)c / iters);
        }
    return 0;
}
