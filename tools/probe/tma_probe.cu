// Standalone TMA probe: tma_probe <variant>; each variant in its own process (errors are sticky)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK, bool CTA, bool SYNC = true>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int x, int y) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = saddr(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (SYNC) __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4) : "memory");
        if (RANK == 2) {
            if (CTA) asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(saddr(sm)), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
            else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(saddr(sm)), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(saddr(sm)), "l"(&tm), "r"(x), "r"(y), "r"(0), "r"(b) : "memory");
        }
    }
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
        if (done) break;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int C = 8, H = 64, W = 64;
    float* d; cudaMalloc(&d, C * H * W * 4);
    float* h = new float[C * H * W]; for (int i = 0; i < C * H * W; ++i) h[i] = (float)i;
    cudaMemcpy(d, h, C * H * W * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 1 << 20); cudaMemset(out, 0, 1 << 20);
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)ptr;
    printf("variant %d entrypoint err=%d q=%d ptr=%p\n", variant, (int)ge, (int)q, ptr);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    int rank = (variant >= 3) ? 3 : 2;
    cuuint64_t dims[3] = {W, H, C}; cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {32, 32, 1}; cuuint32_t es[3] = {1, 1, 1};
    int cx = 0, cy = 0;
    CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (variant == 4) { cx = -1; cy = -1; }
    if (variant == 5) { box[1] = 34; box[2] = 4; }
    if (variant == 6) { box[0] = 36; box[1] = 34; box[2] = 4; cx = -1; cy = -1; }
    if (variant == 7) l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (variant == 9) { box[0] = 36; }
    if (variant == 12) { cx = 1; }
    if (variant == 13) { cy = -1; }
    if (variant == 14) { cx = -4; }
    if (variant == 15) { cx = -1; }
    if (variant == 16) { cx = 60; cy = 60; }
    if (variant == 17) { cx = -32; }
    if (variant == 18) { cx = -1; box[0] = 36; }
    if (variant == 10) { box[1] = 34; }
    if (variant == 11) { box[2] = 4; }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d first qwords %llx %llx %llx %llx\n", (int)r, (unsigned long long)tm.opaque[0], (unsigned long long)tm.opaque[1], (unsigned long long)tm.opaque[2], (unsigned long long)tm.opaque[3]);
    const int n = box[0] * box[1] * (rank == 3 ? box[2] : 1), smem = n * 4 + 1024;
    cudaFuncSetAttribute(k<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (variant == 0) k<2, false><<<1, 128, smem>>>(tm, out, n, 0, 0);
    if (variant == 1) k<2, true><<<1, 128, smem>>>(tm, out, n, 0, 0);
    if (variant == 2) {   // cluster launch attribute 1x1x1
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {1, 1, 1};
        cfg.attrs = at; cfg.numAttrs = 1;
        int x0 = 0, y0 = 0, nn = n;
        cudaLaunchKernelEx(&cfg, k<2, false>, tm, out, nn, x0, y0);
    }
    if (variant == 3 || (variant >= 4 && variant != 8)) k<3, false><<<1, 128, smem>>>(tm, out, n, cx, cy);
    if (variant == 8) k<3, false, false><<<1, 128, smem>>>(tm, out, n, cx, cy);
    cudaError_t e1 = cudaDeviceSynchronize();
    float v[4] = {0};
    if (e1 == cudaSuccess) cudaMemcpy(v, out + 33, 16, cudaMemcpyDeviceToHost);
    printf("run=%s sample=%g %g %g (expect 65 66 67)\n", cudaGetErrorString(e1), v[0], v[1], v[2]);
    return 0;
}
