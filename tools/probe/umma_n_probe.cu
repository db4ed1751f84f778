// Standalone tcgen05 probe: is a kind::tf32 UMMA with M = 128 and N = 8 / 24 / 40 (not a multiple of 16) legal on sm_100a, does it
// produce the right numbers, and what does it cost?  (conv_tc.cu wants N = 24 = 3 output rows x 8 channels in one instruction.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I popcorn_b200/csrc -o tools/probe/umma_n_probe tools/probe/umma_n_probe.cu
//   umma_n_probe N          (one N per process: an illegal instruction poisons the context)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

using namespace pc;

__global__ void __launch_bounds__(128) probe(int N, int nmma, float* dout, long long* clk, int varyA, int varyB, int varyD) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    float* B = reinterpret_cast<float*>(sm);                       // [N rows][32 floats], SWIZZLE_128B, K = 8 used
    for (int i = tid; i < 64 * 32; i += 128) B[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        const int pos = (((k / 4) ^ (n % 8)) * 4) + k % 4;
        B[n * 32 + pos] = (float)((n * 3 + k * 5) % 11 - 5);
    }
    const uint32_t mbar = smem_u32(&bar);
    if (warp == 0) tmem_alloc(smem_u32(&slot), 128);
    if (tid == 0) mbar_init1(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&slot);
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-777.f);   // sentinel in D columns: untouched columns must keep it
    for (int c = 0; c < 64; c += 16) tmem_st16(tbase + lane_off + c, z);
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint((float)((tid + 2 * k) % 7 - 3));
    tmem_st8(tbase + lane_off + 64, a);                            // A at columns 64..71
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t bdesc = make_bdesc(smem_u32(sm));
    long long t0 = 0, t1 = 0;
    if (warp == 0 && elect_one()) {
        tc_fence_after();
        umma_tf32_ts(tbase, tbase + 64, bdesc, idesc, 0u);         // D = A * B^T  (overwrite)
        umma_commit(mbar);
        mbar_wait(mbar, 0);
        t0 = clock64();
        for (int i = 0; i < nmma; ++i) umma_tf32_ts(tbase, tbase + 64, bdesc, idesc, (i > 0) ? 1u : 0u);
        umma_commit(mbar);
        mbar_wait(mbar, 1);
        t1 = clock64();
        clk[0] = t1 - t0;
        // second timing: operands vary from one MMA to the next (values are garbage; only the clock matters)
        //   varyA: A columns cycle through 3 positions; varyB: B start address cycles through 3 row blocks / k offsets; varyD: D columns cycle
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < nmma; i += 3) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const uint32_t a = tbase + 64 + (varyA ? 8u * u : 0u);
                const uint64_t b = bdesc + (uint64_t)(varyB ? ((varyB == 1 ? 2 * u : 64 * u)) : 0);   // 1: +32 B along K, 2: +1024 B (8 rows)
                const uint32_t d = tbase + (varyD ? 8u * u : 0u);
                umma_tf32_ts(d, a, b, idesc, 1u);
            }
        }
        umma_commit(mbar);
        mbar_wait(mbar, 0);
        t1 = clock64();
        clk[1] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    uint32_t d[16];
    for (int c = 0; c < 64; c += 16) {
        tmem_ld16(tbase + lane_off + c, d);
        tc_wait_ld();
        for (int i = 0; i < 16; ++i) dout[tid * 64 + c + i] = __uint_as_float(d[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 24, nmma = 510;
    const int vA = argc > 2 ? atoi(argv[2]) : 0, vB = argc > 3 ? atoi(argv[3]) : 0, vD = argc > 4 ? atoi(argv[4]) : 0;
    float* dout; long long* clk;
    cudaMalloc(&dout, 128 * 64 * 4); cudaMalloc(&clk, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    probe<<<1, 128, 16 * 1024>>>(N, nmma, dout, clk, vA, vB, vD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
    static float h[128 * 64]; long long c, c2;
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&c2, clk + 1, 8, cudaMemcpyDeviceToHost);
    int bad = 0, touched_beyond = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            const float got = h[m * 64 + n];
            if (n < N) {
                float ref = 0.f;
                for (int k = 0; k < 8; ++k) ref += (float)((m + 2 * k) % 7 - 3) * (float)((n * 3 + k * 5) % 11 - 5);
                ref *= (float)nmma;
                if (got != ref) { if (bad < 4) printf("  mismatch m=%d n=%d got %g ref %g\n", m, n, got, ref); ++bad; }
            } else if (got != -777.f) ++touched_beyond;
        }
    printf("N=%3d: %d mismatches, %d columns beyond N touched, %lld clk for %d MMAs -> %.2f clk per MMA | vary A=%d B=%d D=%d: %.2f clk per MMA\n",
           N, bad, touched_beyond, c, nmma, (double)c / nmma, vA, vB, vD, (double)c2 / nmma);
    return 0;
}
