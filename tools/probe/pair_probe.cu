// Standalone probe for the round-2 "fp16 pair" convolution design (profiles/r1c_precision_study.md):
//
//   * activations live in memory as 16-byte pixel chunks  [C/4][H][W][4 channels x (b1, b2)]  with x = h1 + h2, two fp16 in one
//     32-bit word (b1 in the low half = the even K element);
//   * a row of 130 pixels (128 + the kx halo) x C/4 chunks is brought to shared memory by ONE 4-D TMA box {4 words, 130, 1, C/4}
//     (out-of-bounds pixels zero-filled), which IS the canonical K-major no-swizzle ("interleave") UMMA operand layout:
//     16 bytes per row, 8-row groups SBO = 128 B apart, the two 16-byte K chunks of one K=16 instruction LBO = 130*16 B apart;
//   * the kx tap is a +16-byte shift of the descriptor start address, the ky taps are the N = 48 = 3 x 16 accumulator columns;
//   * weights: B[n = (ky, co)][k = (ci, dup)] with every weight duplicated over the (b1, b2) slots, split w = w1 + w2 into two
//     bf16 images, so that  D += A*B(w1) + A*B(w2) = (b1 + b2)(w1 + w2): all four partial products, two kind::f16 UMMAs per 8
//     input channels and tap.
//
// What it answers on a B200 (nothing here was run on hardware in round 1):
//   1. does the SS-form kind::f16 UMMA accept these descriptors (LBO/SBO meaning, start addresses that are only 16-byte
//      aligned) and produce the conv partial sums?  -> max |err| against a CPU reference on the same bf16 operands;
//   2. does the 4-D TMA box deliver that layout, including a negative start coordinate?        -> mode 1 vs mode 0;
//   3. what does an SS-form N = 48 / 96 UMMA cost when A (4 KB per instruction) comes from shared memory?   -> clk per MMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I popcorn_b200/csrc -o tools/probe/pair_probe tools/probe/pair_probe.cu -lcuda
//   pair_probe [swap_lbo_sbo] [variant]     (run each variant in its own process: a faulting descriptor poisons the context)
//     variant 0: SS form, two N = 48 UMMAs per (tap, 8 channels) — weight images w1 and w2 into the same accumulator
//     variant 1: SS form, ONE N = 96 UMMA per (tap, 8 channels) — [B(w1) | B(w2)] side by side, the read-out adds the halves
//                (A is fetched from shared memory once instead of twice)
//     variant 2: tcgen05.cp.128x256b copies each 128 x K16 slice of A from shared memory into TMEM, then two TS-form N = 48 UMMAs
//     variant 3: the 3xTF32 sibling (tools/probe/conv_ss.cu): fp32 chunks [C/4][H][W][4 floats], kind::tf32 (K = 8 = 2 chunks) reads
//                the RAW staged row as A_hi (the tensor core ignores the low 13 mantissa bits), a thread pass writes lo = x - hi
//                into a second buffer; D = A*B_hi + A_lo*B_hi + A*B_lo with tf32 hi/lo weight images
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"

using namespace pc;

constexpr int C = 8, CQ = C / 4;            // input channels, 16-byte chunks per pixel
constexpr int COUT = 16, N = 3 * COUT;      // accumulator columns = (ky, co)
constexpr int H = 8, W = 200;
constexpr int BOXW = 130;                   // 128 pixels + kx halo
constexpr int LBO_A = BOXW * 16;            // bytes between the 16-byte K chunks of A
constexpr int LBO_B = N * 16;               // same for B (N rows of 16 bytes per chunk)
constexpr int A_BYTES = CQ * BOXW * 16;
constexpr int B_IMG = 3 * CQ * N * 16;      // one weight image: [kx][chunk][n][16 B]

// instruction descriptor kind::f16: D = f32 (bits 4-5 = 1), A = B = fp16 (bits 7-9, 10-12 = 0; bf16 would be 1), K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t n) {      // (name kept; the operands are fp16 pairs, see r1c_precision_study.md)
    return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((M >> 4) << 24);
}
// K-major no-swizzle descriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46 | layout 0 @61
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// 128 lanes x 256 bits (= one K = 16 slice of a 16-bit A operand = 8 TMEM columns) from shared memory (matrix descriptor) to TMEM
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

struct Args {
    const uint32_t* act;      // [CQ][H][W][4] pair words
    const uint8_t* wimg;      // two weight images (w1, w2), B_IMG bytes each
    float* out;               // [128][N]
    long long* clk;           // [4]
    int mode;                 // 0: threads copy the row, 1: TMA
    int y, x0, swap, reps, variant;
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                                   // A_BYTES (4160), padded to 8 KB
    uint8_t* sAlo = sm + 8192;                          // variant 3: lo = x - trunc_tf32(x), same layout
    uint8_t* sB = sm + 16384;                           // 2 * B_IMG
    __shared__ __align__(8) unsigned long long bar_tma, bar_mma;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    for (int i = tid; i < 2 * B_IMG / 16; i += 128) reinterpret_cast<int4*>(sB)[i] = reinterpret_cast<const int4*>(a.wimg)[i];
    if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
    if (tid == 0) { mbar_init(smem_u32(&bar_tma), 1); mbar_init(smem_u32(&bar_mma), 1); mbar_init_fence(); }
    if (a.mode == 0) {                                  // the layout the TMA box is expected to produce, written by hand
        for (int i = tid; i < CQ * BOXW; i += 128) {
            const int q = i / BOXW, p = i - q * BOXW, x = a.x0 - 1 + p;
            int4 v = make_int4(0, 0, 0, 0);
            if (x >= 0 && x < W && a.y >= 0 && a.y < H)
                v = reinterpret_cast<const int4*>(a.act)[((size_t)q * H + a.y) * W + x];
            reinterpret_cast<int4*>(sA)[i] = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&slot);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (a.mode == 1) {
        if (tid == 0) {
            t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(A_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(smem_u32(sA)), "l"(&tm), "r"(0), "r"(a.x0 - 1), "r"(a.y), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
        }
        mbar_wait(smem_u32(&bar_tma), 0);
        if (tid == 0) { t1 = clock64(); a.clk[2] = t1 - t0; }
    }
    if (a.variant == 3) {
        for (int i = tid; i < CQ * BOXW * 4; i += 128) {
            const float v = reinterpret_cast<const float*>(sA)[i];
            reinterpret_cast<float*>(sAlo)[i] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0 && elect_one()) {
        const uint32_t lboA = a.swap ? 128u : (uint32_t)LBO_A, sboA = a.swap ? (uint32_t)LBO_A : 128u;
        const uint32_t lboB = a.swap ? 128u : (uint32_t)LBO_B, sboB = a.swap ? (uint32_t)LBO_B : 128u;
        const uint32_t id = idesc_bf16(128, N);
        t2 = clock64();
        const uint32_t id96 = idesc_bf16(128, 2 * N);
        const uint32_t tA = tbase + 128;                           // variant 2: A slices at TMEM columns [128, 128 + 8 * 3 * CQ/2)
        for (int rep = 0; rep < a.reps; ++rep) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int j = 0; j < CQ / 2; ++j) {                 // one K=16 instruction = 8 channels = 2 chunks
                    const uint64_t ad = make_desc_nosw(smem_u32(sA) + kx * 16 + 2 * j * LBO_A, lboA, sboA);
                    const uint32_t first = (rep | kx | j) ? 1u : 0u;
                    if (a.variant == 3) {
                        const uint32_t idt = umma_idesc_tf32(128, N);
                        const uint64_t alo = make_desc_nosw(smem_u32(sAlo) + kx * 16 + 2 * j * LBO_A, lboA, sboA);
                        const uint64_t bhi = make_desc_nosw(smem_u32(sB) + (kx * CQ + 2 * j) * LBO_B, lboB, sboB);
                        const uint64_t blo = make_desc_nosw(smem_u32(sB) + B_IMG + (kx * CQ + 2 * j) * LBO_B, lboB, sboB);
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                                     ::"r"(tbase), "l"(ad), "l"(bhi), "r"(idt), "r"(first), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                                     ::"r"(tbase), "l"(alo), "l"(bhi), "r"(idt), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                                     ::"r"(tbase), "l"(ad), "l"(blo), "r"(idt), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
                    } else if (a.variant == 1) {                          // B rows [0,48) = w1 image, [48,96) = w2 image: same chunk, 96 rows
                        const uint64_t bd = make_desc_nosw(smem_u32(sB) + (kx * CQ + 2 * j) * (2 * LBO_B), a.swap ? 128u : 2u * LBO_B,
                                                           a.swap ? 2u * LBO_B : 128u);
                        umma_bf16_ss(tbase, ad, bd, id96, first);
                    } else {
                        if (a.variant == 2) tmem_cp_128x256b(tA + 8 * (kx * (CQ / 2) + j), ad);
#pragma unroll
                        for (int img = 0; img < 2; ++img) {
                            const uint64_t bd = make_desc_nosw(smem_u32(sB) + img * B_IMG + (kx * CQ + 2 * j) * LBO_B, lboB, sboB);
                            if (a.variant == 2) umma_bf16_ts(tbase, tA + 8 * (kx * (CQ / 2) + j), bd, id, first | (uint32_t)img);
                            else umma_bf16_ss(tbase, ad, bd, id, first | (uint32_t)img);
                        }
                    }
                }
        }
        umma_commit(smem_u32(&bar_mma));
    }
    mbar_wait(smem_u32(&bar_mma), 0);
    t3 = clock64();
    tc_fence_after();
    if (tid == 0) { a.clk[0] = t3 - t2; a.clk[1] = (long long)a.reps * 3 * (CQ / 2) * (a.variant == 1 ? 1 : a.variant == 3 ? 3 : 2); }
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16], u[16];
        tmem_ld16(tbase + lane_off + c, v);
        tmem_ld16(tbase + lane_off + N + c, u);                      // variant 1: the w2 half of the accumulator
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            a.out[(size_t)tid * N + c + i] = __uint_as_float(v[i]) + (a.variant == 1 ? __uint_as_float(u[i]) : 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// fp16 pieces (the names say bf16 for historical reasons: the first version of this probe used bf16 pairs, which fail the parity bar)
static uint16_t bf16_rn(float f) { const __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float bf16_f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int swap = argc > 1 ? atoi(argv[1]) : 0;
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    // activations (fp32 truth, and their pair words)
    std::vector<float> act(C * H * W);
    std::vector<uint32_t> words((size_t)CQ * H * W * 4);
    for (int c = 0; c < C; ++c)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float v = rnd() * 3.f;
                act[(c * H + y) * W + x] = v;
                const uint16_t b1 = bf16_rn(v), b2 = bf16_rn(v - bf16_f(b1));
                words[(((size_t)(c / 4) * H + y) * W + x) * 4 + c % 4] = (uint32_t)b1 | ((uint32_t)b2 << 16);
            }
    // weights [co][ci][ky][kx] and the two duplicated bf16 images [kx][chunk][n][8 bf16 = 4 ch x (w, w)]
    std::vector<float> wt(COUT * C * 9);
    for (auto& v : wt) v = rnd() * 0.3f;
    std::vector<uint16_t> wimg(2 * B_IMG / 2, 0);
    for (int img = 0; img < 2; ++img)
        for (int kx = 0; kx < 3; ++kx)
            for (int q = 0; q < CQ; ++q)
                for (int ky = 0; ky < 3; ++ky)
                    for (int co = 0; co < COUT; ++co)
                        for (int e = 0; e < 4; ++e) {
                            const float w = wt[((co * C + q * 4 + e) * 3 + ky) * 3 + kx];
                            const uint16_t w1 = bf16_rn(w), w2 = bf16_rn(w - bf16_f(w1));
                            const size_t base = (size_t)img * (B_IMG / 2) + (((size_t)(kx * CQ + q) * N + ky * COUT + co) * 8) + 2 * e;
                            wimg[base] = wimg[base + 1] = img ? w2 : w1;
                        }
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    if (variant == 1) {                     // [kx][chunk][96 rows = w1 image | w2 image][16 B]
        std::vector<uint16_t> w96(wimg.size());
        for (int kq = 0; kq < 3 * CQ; ++kq)
            for (int img = 0; img < 2; ++img)
                for (int i = 0; i < N * 8; ++i)
                    w96[((size_t)kq * 2 + img) * N * 8 + i] = wimg[(size_t)img * (B_IMG / 2) + (size_t)kq * N * 8 + i];
        wimg = w96;
    }
    auto tf32_hi = [](float f) { uint32_t u; memcpy(&u, &f, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    if (variant == 3) {                     // fp32 chunks (same bytes as the pair words) + tf32 hi / lo weight images [kx][chunk][n][4 floats]
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) memcpy(&words[(((size_t)(c / 4) * H + y) * W + x) * 4 + c % 4], &act[(c * H + y) * W + x], 4);
        std::vector<float> w32(2 * B_IMG / 4, 0.f);
        for (int kx = 0; kx < 3; ++kx)
            for (int q = 0; q < CQ; ++q)
                for (int ky = 0; ky < 3; ++ky)
                    for (int co = 0; co < COUT; ++co)
                        for (int e = 0; e < 4; ++e) {
                            const float w = wt[((co * C + q * 4 + e) * 3 + ky) * 3 + kx], hi = tf32_hi(w);
                            const size_t at = ((size_t)(kx * CQ + q) * N + ky * COUT + co) * 4 + e;
                            w32[at] = hi; w32[B_IMG / 4 + at] = w - hi;
                        }
        memcpy(wimg.data(), w32.data(), 2 * B_IMG);
    }
    uint32_t* d_act; uint8_t* d_w; float* d_out; long long* d_clk;
    cudaMalloc(&d_act, words.size() * 4); cudaMalloc(&d_w, 2 * B_IMG); cudaMalloc(&d_out, 128 * N * 4); cudaMallocManaged(&d_clk, 64);
    cudaMemcpy(d_act, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, wimg.data(), 2 * B_IMG, cudaMemcpyHostToDevice);

    void* ptr = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[4] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CQ};
    cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    cuuint32_t box[4] = {4, BOXW, 1, CQ}, es[4] = {1, 1, 1, 1};
    CUresult er = ((EncodeTiledFn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, d_act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled (uint32, 4-D box {4,%d,1,%d}) -> %d ; swap_lbo_sbo=%d variant=%d\n", BOXW, CQ, (int)er, swap, variant);

    const int smem = 16384 + 2 * B_IMG + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(128 * N);
    const int cases[][2] = {{3, 0}, {0, 0}, {H - 1, 64}, {4, 128}};            // (row y, tile origin x0): x0 = 0 reads pixel -1, 128 runs past W
    for (int mode = 0; mode <= 1; ++mode)
        for (auto& cs : cases) {
            Args a{d_act, d_w, d_out, d_clk, mode, cs[0], cs[1], swap, 1, variant};
            cudaMemset(d_out, 0, 128 * N * 4);
            probe<<<1, 128, smem>>>(tm, a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d y %d x0 %d: %s\n", mode, cs[0], cs[1], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), d_out, 128 * N * 4, cudaMemcpyDeviceToHost);
            double err_pair = 0, err_fp32 = 0, ref_max = 0;
            for (int t = 0; t < 128; ++t)
                for (int ky = 0; ky < 3; ++ky)
                    for (int co = 0; co < COUT; ++co) {
                        double rp = 0, rf = 0;
                        for (int kx = 0; kx < 3; ++kx)
                            for (int ci = 0; ci < C; ++ci) {
                                const int x = cs[1] + t + kx - 1;
                                if (x < 0 || x >= W) continue;
                                const float v = act[(ci * H + cs[0]) * W + x], w = wt[((co * C + ci) * 3 + ky) * 3 + kx];
                                const uint16_t b1 = bf16_rn(v), b2 = bf16_rn(v - bf16_f(b1)), w1 = bf16_rn(w), w2 = bf16_rn(w - bf16_f(w1));
                                if (variant == 3) {
                                    const float vh = tf32_hi(v), vl = tf32_hi(v - vh), wh = tf32_hi(w), wl = tf32_hi(w - wh);
                                    rp += (double)vh * wh + (double)vl * wh + (double)vh * wl;
                                } else
                                rp += ((double)bf16_f(b1) + bf16_f(b2)) * ((double)bf16_f(w1) + bf16_f(w2));
                                rf += (double)v * w;
                            }
                        const double g = out[(size_t)t * N + ky * COUT + co];
                        err_pair = fmax(err_pair, fabs(g - rp)); err_fp32 = fmax(err_fp32, fabs(g - rf)); ref_max = fmax(ref_max, fabs(rf));
                    }
            printf("mode %d (%s) y=%d x0=%3d : max|err| vs same split operands %.3e, vs fp32 conv %.3e (max |ref| %.2f)%s\n", mode,
                   mode ? "TMA 4-D box" : "thread copy", cs[0], cs[1], err_pair, err_fp32, ref_max, err_pair < 1e-4 ? "  OK" : "  MISMATCH");
            if (mode == 1) printf("         TMA row (%d B) issue -> landed: %lld clk\n", A_BYTES, d_clk[2]);
        }
    // cost of SS-form UMMAs with A from shared memory
    for (int reps : {1, 64, 512}) {
        Args a{d_act, d_w, d_out, d_clk, 0, 3, 0, swap, reps, variant};
        probe<<<1, 128, smem>>>(tm, a);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("timing run failed\n"); return 1; }
        const int n = variant == 1 ? 2 * N : N;
        printf("variant %d kind::f16 M=128 N=%d K=16: %lld MMAs in %lld clk -> %.1f clk per MMA (N/2 = %d would be the tensor rate)\n", variant, n,
               d_clk[1], d_clk[0], (double)d_clk[0] / d_clk[1], n / 2);
    }
    return 0;
}
