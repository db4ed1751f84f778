// Standalone probe: does concurrent tcgen05.st / tcgen05.ld traffic of OTHER warps (other TMEM columns) slow a chain of tcgen05.mma
// (kind::tf32, M = 128, A in TMEM, B in shared memory)?  In the conv / head pipelines a UMMA retires every ~1.8x its isolated time.
//   umma_contention_probe N bg_mode bg_warps      bg_mode 0 none, 1 st.x16 loop, 2 ld.x16 loop, 3 shared-memory LDS loop
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace pc;

__global__ void __launch_bounds__(1024) probe(int N, int nmma, int bg_mode, long long* clk, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    const uint32_t mbar = smem_u32(&bar);
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    if (tid == 0) { mbar_init1(mbar); stop = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&slot);
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0;
    if (warp < 4) for (int c = 0; c < 512; c += 16) tmem_st16(tbase + lane_off + c, z);
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_tf32(128, N);
            const uint64_t bdesc = make_bdesc(smem_u32(sm));
            const long long t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < nmma; i += 3) {
#pragma unroll
                for (int u = 0; u < 3; ++u) umma_tf32_ts(tbase + 8u * u, tbase + 256 + 8u * u, bdesc + 2 * u, idesc, 1u);
            }
            umma_commit(mbar);
            mbar_wait(mbar, 0);
            clk[0] = clock64() - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (bg_mode) {
        // background traffic on columns 320.. (never the MMA's A / D columns)
        const uint32_t col = 320u + (uint32_t)(((warp >> 2) % 3) * 64);
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = tid + i;
        float acc = 0.f;
        const float* sp = reinterpret_cast<const float*>(sm) + 8192 + (tid & 31);
        while (!stop) {
            if (bg_mode == 1) {
#pragma unroll
                for (int c = 0; c < 64; c += 16) tmem_st16(tbase + lane_off + col + c, r);
                tc_wait_st();
            } else if (bg_mode == 2) {
#pragma unroll
                for (int c = 0; c < 64; c += 16) tmem_ld16(tbase + lane_off + col + c, r);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) acc += sp[c * 136];
            }
        }
        if (acc == 123.f || r[0] == 0x7fffffff) sink[tid] = acc + r[3];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 24;
    long long* clk; float* sink;
    cudaMalloc(&clk, 16); cudaMalloc(&sink, 4096 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int nmma = 3000;
    printf("N=%d: background traffic vs clk per UMMA\n", N);
    for (int mode = 0; mode < 4; ++mode)
        for (int bw : {4, 8, 16}) {
            if (mode == 0 && bw != 4) continue;
            probe<<<1, 32 * (1 + bw), 80 * 1024>>>(N, nmma, mode, clk, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long c;
            cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            printf("  %-14s bg warps %2d : %.2f clk per UMMA\n", mode == 0 ? "none" : mode == 1 ? "tcgen05.st x16" : mode == 2 ? "tcgen05.ld x16" : "LDS", mode ? bw : 0,
                   (double)c / nmma);
        }
    return 0;
}
