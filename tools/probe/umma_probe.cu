// Standalone tcgen05 probe: cycles per tcgen05.mma kind::tf32 (M=128, A operand in TMEM, B in shared memory,
// K = 8 per instruction) as a function of N, and the issue -> commit -> mbarrier round-trip latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I popcorn_b200/csrc -o tools/probe/umma_probe tools/probe/umma_probe.cu
//   umma_probe            -> table on stdout
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace pc;

__global__ void __launch_bounds__(128) probe(int N, int nmma, int same_d, int tmem_cols, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    const uint32_t mbar = smem_u32(&bar);
    if (warp == 0) tmem_alloc(smem_u32(&slot), tmem_cols);
    if (tid == 0) mbar_init1(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&slot);
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    // zero the TMEM block (A operand and accumulators): garbage could be NaN
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0;
    for (int c = 0; c < tmem_cols; c += 16) tmem_st16(tbase + lane_off + c, z);
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t bdesc = make_bdesc(smem_u32(sm));
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 0 && elect_one()) {
        tc_fence_after();
        // A at columns [0, 64) (8 k-steps), D from column 64
        t0 = clock64();
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                umma_tf32_ts(tbase + 64 + (same_d ? 0 : ((i + j) & 1) * N), tbase + 8 * j, bdesc, idesc, 1u);
        }
        t1 = clock64();
        umma_commit(mbar);
    }
    mbar_wait(mbar, 0);
    t2 = clock64();
    tc_fence_after();
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, tmem_cols);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
    printf("N  nmma  same_d  ctas_per_sm  issue_clk  total_clk  clk_per_mma\n");
    const int Ns[] = {16, 32, 48, 64, 96, 128, 192};
    for (int ctas = 1; ctas <= 2; ++ctas)
        for (int sd = 0; sd <= 1; ++sd)
            for (int N : Ns) {
                if (64 + 2 * N > 512 / ctas) continue;
                for (int nmma : {8, 512, 4096}) {
                    out[0] = out[1] = 0;
                    probe<<<148 * ctas, 128, 34 * 1024>>>(N, nmma, sd, 512 / ctas, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    printf("%3d %5d %d %d %8lld %8lld %8.2f\n", N, nmma, sd, ctas, out[0], out[1], (double)out[1] / nmma);
                }
            }
    return 0;
}
