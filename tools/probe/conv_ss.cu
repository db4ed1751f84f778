// ROUND-2 CANDIDATE — written without a GPU at the end of round 1, never run.  Built into its own library
// (tools/probe/build_conv_pair.sh -> popcorn_b200/libpopcorn_b200_probe.so), NOT into libpopcorn_b200.so.
//
// 3x3 convolution (+ folded BN bias + ReLU, optional 2x2 max-pool), 3xTF32 like csrc/conv_tc.cu — SAME operand precision and fp32
// range, so the parity tests apply unchanged — but with the A operand read by the UMMA straight from shared memory (SS form):
//   * activations live in memory as fp32 in 16-byte pixel chunks  [C/4][H][W][4 channels]  (same bytes as planar NCHW);
//   * one 4-D TMA box {4 floats, 136 px, 1 row, C/4 chunks} per input row lands in shared memory as the canonical K-major no-swizzle
//     UMMA operand: pixel stride 16 B, 8-row groups SBO = 128 B apart, 16-byte K chunks LBO = 136*16 B apart (zero-filled outside);
//   * UMMA kind::tf32 reads raw fp32 words and ignores the low 13 mantissa bits: the staged row IS A_hi, for free.  The kx tap is a
//     +16-byte shift of the descriptor start address (conv_tc.cu writes three shifted copies of every row into TMEM instead);
//   * the only register-level work left: one pass lo = x - trunc_tf32(x) per staged value into a second buffer of the same layout
//     (1 ld.shared.v4 + 1 st.shared.v4 per 4 channels, against 3*Cin ld.shared + splits + 6*Cin/8 tcgen05.st per pixel today);
//   * per (tap, 8 channels): A*B_hi + A*B_lo + A_lo*B_hi, K = 8, N = 48 = the three ky accumulators of output rows r-1, r, r+1 in an
//     8-slot TMEM ring, B rows [W_ky2 | W_ky1 | W_ky0]; TMEM carries accumulators only (no A traffic on its single port);
//   * the epilogue stores 16-byte pixel chunks (coalesced) instead of 8-16 scalar plane stores per thread.
// Roles: warp 0 = TMA producer (one lane), warp 1 = UMMA issuer (one lane), warps 2-9 = two epilogue groups alternating output-row
// pairs, warps 10-17 = the lo pass (two groups alternating input rows).
// Sibling of conv_pair.cu (fp16-pair operands: fewer tensor slots, but a 65504 range that the parity weights nearly exhaust —
// profiles/r1c_precision_study.md); this variant changes nothing numerically.  Hardware questions: tools/probe/pair_probe.cu.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"
#include "conv_ss.cuh"

namespace pc {

// K-major no-swizzle matrix descriptor: start>>4 | LBO>>4 @16 (between 16-byte K chunks) | SBO>>4 @32 (between 8-row groups) | version 1 @46
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

template <int CQ, int COUT>
struct SsGeom {
    static constexpr int KSTEPS = CQ / 2;                          // K = 8 instructions per tap: 8 channels = 2 chunks each
    static constexpr int IMG_HALF = 3 * CQ * PBROWS * 16;          // one weight image [kx][chunk][48 rows][4 floats]
    static constexpr int OFF_BIAS = 2 * IMG_HALF;                  // images: [hi | lo]
    static constexpr int W_BYTES = OFF_BIAS + 64;
    static constexpr int ROW = CQ * PCHUNK;                        // one input row, all chunks (multiple of 128 B)
    static constexpr int STAGE = 2 * ROW;                          // raw row (= A_hi as the tensor core reads it) + its lo row
    static constexpr int NS = CQ <= 2 ? 12 : CQ <= 4 ? 8 : 4;      // rows in flight
    static constexpr int OFF_STAGE = (W_BYTES + 127) / 128 * 128;
    static constexpr int OFF_BARS = OFF_STAGE + NS * STAGE;        // s_full[NS] a_ready[NS] s_empty[NS] d_full[PND] d_empty[PND]
    static constexpr int OFF_TMEM = OFF_BARS + 8 * (3 * NS + 2 * PND);
    static constexpr int SMEM_NEED = OFF_TMEM + 16 + 1024;
    static constexpr int SMEM_BYTES = SMEM_NEED > 116 * 1024 ? SMEM_NEED : 116 * 1024;   // one CTA per SM
    static_assert(CQ % 2 == 0 && SMEM_BYTES <= 227 * 1024, "geometry");
    // the two lo-pass groups alternate input rows (i & 1): with an even ring depth a stage is always served by the same group, so a
    // group has seen phase n-1 of a stage's barrier complete before it waits for phase n (parity waits alias two phases apart —
    // tools/probe/simulate_protocol.py finds the false wake-up with an odd NS)
    static_assert(NS % 2 == 0, "ring depth must be even");
};

template <int CQA, int CQB, int COUT, int EPI>
__global__ void __launch_bounds__(PTHREADS, 1) conv3x3_ss_kernel(const __grid_constant__ SsParams p) {
    constexpr int CQ = CQA + CQB;
    using G = SsGeom<CQ, COUT>;
    constexpr int NS = G::NS;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const SsJob& job = p.jobs[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const float* bias = reinterpret_cast<const float*>(sm + G::OFF_BIAS);
    const uint32_t bars = smem_u32(sm + G::OFF_BARS);
    auto s_full = [&](int i) { return bars + 8u * (uint32_t)i; };                    // TMA (tx bytes) -> lo pass
    auto a_ready = [&](int i) { return bars + 8u * (uint32_t)(NS + i); };            // lo pass (4 warps) -> issuer
    auto s_empty = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + i); };        // issuer (commit) -> TMA
    auto d_full = [&](int i) { return bars + 8u * (uint32_t)(3 * NS + i); };         // issuer (commit) -> epilogue, per output row slot
    auto d_empty = [&](int i) { return bars + 8u * (uint32_t)(3 * NS + PND + i); };  // epilogue (4 warps) -> issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + G::OFF_TMEM);

    for (int i = tid; i < G::W_BYTES / 16; i += PTHREADS)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(job.wimg) + i);
    constexpr uint32_t TCOLS = 128u;                                  // the accumulator ring is all that lives in TMEM
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TCOLS);
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(s_full(i), 1); mbar_init(a_ready(i), 4); mbar_init(s_empty(i), 1); }
        for (int i = 0; i < PND; ++i) { mbar_init(d_full(i), 1); mbar_init(d_empty(i), 4); }
        mbar_init_fence();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // weights + barriers -> visible to UMMA / TMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;      // a warp may touch TMEM lanes 32*(warp%4) .. +31
    if (warp >= 2 && warp < 6) {                                      // UMMAs only ever accumulate: all slots start at zero
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int i = 0; i < PND; ++i) tmem_st16(tbase + lane_off + 16 * i, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int H = p.H, W = p.W, TR = p.TR;                            // TR even
    const int ntiles = p.tiles_x * p.tiles_y;
    auto tile_rows = [&](int tile) {                                  // rounded up to even: an odd last row is computed, never stored
        const int y0 = (tile / p.tiles_x) * TR;
        const int n = (H - y0) < TR ? (H - y0) : TR;
        return (n + 1) & ~1;
    };

    if (warp == 0) {
        if (elect_one()) {
            // =========================== TMA producer ===========================
            const uint32_t stage_base = smem_u32(sm + G::OFF_STAGE);
            const CUtensorMap* tmA = &p.tmA[blockIdx.y];
            const CUtensorMap* tmB = &p.tmB[blockIdx.y];
            int i = 0;
#pragma unroll 1
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
                const int x0 = tx * 128, y0 = ty * TR, nrows = tile_rows(tile);
#pragma unroll 1
                for (int r = -1; r <= nrows; ++r, ++i) {
                    const int s = i % NS, n = i / NS;
                    if (n >= 1) mbar_wait_sleep(s_empty(s), (uint32_t)(n - 1) & 1u);
                    const uint32_t dst = stage_base + (uint32_t)s * G::STAGE;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_full(s)), "r"(G::ROW) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(dst), "l"(tmA), "r"(0), "r"(x0 - 1 - job.a_ox), "r"(y0 + r - job.a_oy), "r"(0), "r"(s_full(s)) : "memory");
                    if (CQB > 0)
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                     ::"r"(dst + CQA * PCHUNK), "l"(tmB), "r"(0), "r"(x0 - 1 - job.b_ox), "r"(y0 + r - job.b_oy), "r"(0), "r"(s_full(s)) : "memory");
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // =========================== UMMA issuer ===========================
            const uint32_t stage_base = smem_u32(sm + G::OFF_STAGE), sW = smem_u32(sm);
            int i = 0, g0 = 0;                                        // running input-row (ring) index / output-row index of the tile's row 0
#pragma unroll 1
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int nrows = tile_rows(tile);
#pragma unroll 1
                for (int r = -1; r <= nrows; ++r, ++i) {
                    const int s = i % NS;
                    {
                        // the row and its lo copy are staged (a_ready) + the slot of output row r+1, which this row opens, has been
                        // drained (d_empty): ONE merged probe — a blocking mbarrier probe costs ~200 cycles even when long complete
                        const uint32_t m1 = a_ready(s), p1 = (uint32_t)(i / NS) & 1u;
                        uint32_t m2 = m1, p2 = p1;
                        const int g = g0 + r + 1;
                        if (r + 1 <= nrows - 1 && g >= PND) { m2 = d_empty(g % PND); p2 = (uint32_t)(g / PND - 1) & 1u; }
                        mbar_wait3_sleep(m1, p1, m2, p2, m2, p2);
                    }
                    tc_fence_after();
                    const uint32_t sA = stage_base + (uint32_t)s * G::STAGE;
                    const int lo = r - 1 < 0 ? 0 : r - 1, hi = r + 1 > nrows - 1 ? nrows - 1 : r + 1;
                    int o = lo;
                    while (o <= hi) {                                 // runs of output rows whose ring slots are adjacent (the ring wraps)
                        const int slot = (g0 + o) % PND;
                        int n = hi - o + 1;
                        if (n > PND - slot) n = PND - slot;
                        const uint32_t d = tbase + 16u * (uint32_t)slot;
                        const uint32_t brow = 16u * (uint32_t)(o - r + 1);          // output row o takes tap ky = r - o + 1 = B block 2 - ky
                        const uint32_t id = umma_idesc_tf32(128, 16u * (uint32_t)n);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int j = 0; j < G::KSTEPS; ++j) {
                                const uint64_t a_raw = desc_nosw(sA + kx * 16 + 2 * j * PCHUNK, PCHUNK, 128);          // read as A_hi
                                const uint64_t a_lo = desc_nosw(sA + G::ROW + kx * 16 + 2 * j * PCHUNK, PCHUNK, 128);
                                const uint32_t boff = ((kx * CQ + 2 * j) * PBROWS + brow) * 16;
                                const uint64_t b_hi = desc_nosw(sW + boff, PBROWS * 16, 128), b_lo = desc_nosw(sW + G::IMG_HALF + boff, PBROWS * 16, 128);
                                umma_tf32_ss(d, a_raw, b_hi, id);
                                umma_tf32_ss(d, a_lo, b_hi, id);
                                umma_tf32_ss(d, a_raw, b_lo, id);
                            }
                        o += n;
                    }
                    umma_commit(s_empty(s));                                       // the staged row may be overwritten
                    if (r >= 1) umma_commit(d_full((g0 + r - 1) % PND));           // output row r-1 has all three taps
                }
                g0 += nrows;
            }
        }
    } else if (warp >= 10) {
        // =========================== lo pass: lo = x - trunc_tf32(x) for every staged value, same layout, second buffer ===========================
        const int lgroup = (warp - 10) >> 2;                                       // two groups of four warps alternate input rows: a blocking
        const int t = ((warp - 10) & 3) * 32 + lane;                               // barrier probe (~200 clk) per row must not serialise them
        uint8_t* stage0 = sm + G::OFF_STAGE;
        int i = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int nrows = tile_rows(tile);
#pragma unroll 1
            for (int r = -1; r <= nrows; ++r, ++i) {
                if ((i & 1) != lgroup) continue;
                const int s = i % NS;
                mbar_wait_sleep(s_full(s), (uint32_t)(i / NS) & 1u);
                const float4* raw = reinterpret_cast<const float4*>(stage0 + (size_t)s * G::STAGE);
                float4* lo = reinterpret_cast<float4*>(stage0 + (size_t)s * G::STAGE + G::ROW);
#pragma unroll 1
                for (int k = t; k < CQ * PBOX; k += 128) {
                    const float4 v = raw[k];
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    lo[k] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the UMMA's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready(s));
            }
        }
    } else {
        // =========================== epilogue: two groups alternate output-row pairs ===========================
        const int group = (warp - 2) >> 2;
        const int px = (warp & 3) * 32 + lane;                                     // TMEM lane == pixel; (warp & 3) is also the lane quarter
        const int cH = p.crop_H > 0 ? p.crop_H : H, cW = p.crop_H > 0 ? p.crop_W : W;
        float dotw[EPI == PEPI_DOT ? 8 : 1];
        float dotb = 0.f;
        if (EPI == PEPI_DOT) {
#pragma unroll
            for (int o = 0; o < 8; ++o) dotw[o] = __ldg(job.dotw + o);
            dotb = __ldg(job.dotw + 8);
        }
        int g0 = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            const int y0 = ty * TR, nrows = tile_rows(tile);
            const int vx = tx * 128 + px;
#pragma unroll 1
            for (int m = 0; m < nrows / 2; ++m) {
                const int g = g0 + 2 * m;                                          // even: the pair's slots are adjacent
                if (((g >> 1) & 1) != group) continue;
                const int slot = g % PND;
                mbar_wait_sleep(d_full(slot + 1), (uint32_t)(g / PND) & 1u);       // commits are ordered: row g is final as well
                tc_fence_after();
                uint32_t d[2][16];
                const uint32_t t = tbase + lane_off + 16u * (uint32_t)slot;
                tmem_ld16(t, d[0]);
                tmem_ld16(t + 16, d[1]);
                tc_wait_ld();
                {
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = 0u;
                    tmem_st16(t, z);
                    tmem_st16(t + 16, z);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(d_empty(slot)); mbar_arrive(d_empty(slot + 1)); }
                float acc[2][COUT];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        const float v = __uint_as_float(d[h][o]) + bias[o];
                        acc[h][o] = job.linear ? v : fmaxf(v, 0.f);
                    }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int oy = y0 + 2 * m + h;
                    if (oy >= H || vx >= W) continue;
                    if (job.out_c4) {
#pragma unroll
                        for (int q = 0; q < COUT / 4; ++q)
                            reinterpret_cast<float4*>(job.out_c4)[((size_t)q * H + oy) * W + vx] =
                                make_float4(acc[h][4 * q], acc[h][4 * q + 1], acc[h][4 * q + 2], acc[h][4 * q + 3]);
                    }
                    const int yy = oy - p.crop_y, xx = vx - p.crop_x;
                    const bool inside = yy >= 0 && yy < cH && xx >= 0 && xx < cW;
                    if (job.out_planar && inside) {
#pragma unroll
                        for (int o = 0; o < COUT; ++o) job.out_planar[(long long)o * job.out_cs + (long long)yy * job.out_rs + xx] = acc[h][o];
                    }
                    if (EPI == PEPI_DOT && inside) {
                        float sacc = 0.f;
#pragma unroll
                        for (int o = 0; o < 8; ++o) sacc = fmaf(acc[h][o], dotw[o], sacc);
                        if (job.dot_in) sacc += job.dot_in[(long long)yy * job.dot_in_rs + xx];
                        if (job.dot_final) {
                            sacc += dotb;
                            sacc = 1.f / (1.f + expf(-sacc));
                        }
                        job.dot_out[(long long)yy * job.dot_out_rs + xx] = sacc;
                    }
                }
                if (EPI == PEPI_POOL) {                               // 2x2 max over (rows 2m, 2m+1) x (lanes 2k, 2k+1)
                    const int pH = H >> 1, pW = W >> 1;
                    const int py = (y0 >> 1) + m, pxl = vx >> 1;
                    const bool stp = !(lane & 1) && py < pH && pxl < pW;
                    float hm[COUT];
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        const float vm = fmaxf(acc[0][o], acc[1][o]);
                        hm[o] = fmaxf(vm, __shfl_xor_sync(FULL, vm, 1));
                    }
                    if (stp) {
#pragma unroll
                        for (int q = 0; q < COUT / 4; ++q)
                            reinterpret_cast<float4*>(job.pool_c4)[((size_t)q * pH + py) * pW + pxl] =
                                make_float4(hm[4 * q], hm[4 * q + 1], hm[4 * q + 2], hm[4 * q + 3]);
                    }
                }
            }
            g0 += nrows;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, TCOLS);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int conv_ss_img_bytes(int cin, int /*cout*/) { return 2 * 3 * (cin / 4) * PBROWS * 16 + 64; }

// flat = [cin][ky][kx][cout] + bias[cout] (the SIMT pack, BN folded) -> [hi | lo][kx][chunk][48 rows][4 floats] + bias[16];
// row = 16 * (2 - ky) + co, hi = top 19 bits of w (exact TF32), lo = w - hi  (same split as conv_tc_pack_layer)
static void conv_ss_pack_layer(const float* flat, int cin, int cout, uint8_t* img) {
    const int cq = cin / 4, half = 3 * cq * PBROWS * 4;            // floats of one image
    memset(img, 0, conv_ss_img_bytes(cin, cout));
    float* w32 = reinterpret_cast<float*>(img);
    for (int kx = 0; kx < 3; ++kx)
        for (int q = 0; q < cq; ++q)
            for (int ky = 0; ky < 3; ++ky)
                for (int co = 0; co < cout; ++co)
                    for (int e = 0; e < 4; ++e) {
                        const float w = flat[(((q * 4 + e) * 3 + ky) * 3 + kx) * cout + co];
                        uint32_t bits;
                        memcpy(&bits, &w, 4);
                        bits &= 0xFFFFE000u;
                        float hi;
                        memcpy(&hi, &bits, 4);
                        const size_t at = ((size_t)(kx * cq + q) * PBROWS + (2 - ky) * 16 + co) * 4 + e;
                        w32[at] = hi;
                        w32[half + at] = w - hi;
                    }
    float* b = w32 + 2 * half;
    for (int co = 0; co < cout; ++co) b[co] = flat[cin * 9 * cout + co];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// chunked tensor [CQ][H][W][4 floats] -> 4-D map, box {4, PBOX, 1, CQ}
static bool make_tmap_c4(CUtensorMap* tm, const float* ptr, int cq, int H, int W) {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    if (!fn || !ptr || (((uintptr_t)ptr) & 15)) return false;
    cuuint64_t dims[4] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cq};
    cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    cuuint32_t box[4] = {4, PBOX, 1, (cuuint32_t)cq}, es[4] = {1, 1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CQA, int CQB, int COUT, int EPI>
static int launch_ss(SsParams& p, int njobs, cudaStream_t st) {
    using G = SsGeom<CQA + CQB, COUT>;
    auto k = conv3x3_ss_kernel<CQA, CQB, COUT, EPI>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int per_job = nsm / njobs;
    if (per_job < 1) per_job = 1;
    p.tiles_x = cdiv(p.W, 128);
    if (p.TR <= 0) p.TR = ((long long)p.tiles_x * cdiv(p.H, 64) >= 8ll * per_job) ? 64 : 32;
    p.tiles_y = cdiv(p.H, p.TR);
    const int ntiles = p.tiles_x * p.tiles_y;
    if (per_job > ntiles) per_job = ntiles;
    k<<<dim3(per_job, njobs), PTHREADS, G::SMEM_BYTES, st>>>(p);
    PC_CUDA(cudaGetLastError());
    return 0;
}

// ---- API for the layer schedule (dda_c4.cu) ----
int conv_ss_image_bytes(int cin, int cout) { return conv_ss_img_bytes(cin, cout); }
void conv_ss_pack(const float* flat, int cin, int cout, uint8_t* img) { conv_ss_pack_layer(flat, cin, cout, img); }
bool conv_ss_tmap(CUtensorMap* tm, const float* ptr, int cq, int H, int W) { return make_tmap_c4(tm, ptr, cq, H, W); }
int conv_ss_launch(int cqa, int cqb, int cout, int epi, SsParams& p, int njobs, cudaStream_t st) {
    const int key = ((cqa * 10 + cqb) * 100 + cout) * 10 + epi;
    switch (key) {
        case 20080 + PEPI_STORE: return launch_ss<2, 0, 8, PEPI_STORE>(p, njobs, st);
        case 20080 + PEPI_POOL: return launch_ss<2, 0, 8, PEPI_POOL>(p, njobs, st);
        case 20080 + PEPI_DOT: return launch_ss<2, 0, 8, PEPI_DOT>(p, njobs, st);
        case 20160 + PEPI_STORE: return launch_ss<2, 0, 16, PEPI_STORE>(p, njobs, st);
        case 40160 + PEPI_STORE: return launch_ss<4, 0, 16, PEPI_STORE>(p, njobs, st);
        case 40160 + PEPI_POOL: return launch_ss<4, 0, 16, PEPI_POOL>(p, njobs, st);
        case 44080 + PEPI_STORE: return launch_ss<4, 4, 8, PEPI_STORE>(p, njobs, st);
        case 22080 + PEPI_STORE: return launch_ss<2, 2, 8, PEPI_STORE>(p, njobs, st);
    }
    return PC_ERR_INVALID;
}

// pc::set_error for the probe library is defined in conv_pair.cu

}  // namespace pc

using namespace pc;

// One conv layer on chunked fp32 tensors (device pointers); `wflat` is a HOST pointer to [cin][3][3][cout] + bias[cout].
// a: [cqa][a_H][a_W][4] floats, b (optional): [cqb][b_H][b_W][4]; outputs any of: out_c4 [cout/4][H][W][4], out_planar [cout][H][W],
// pool_c4 [cout/4][H/2][W/2][4].  tile_rows: 0 = automatic.
extern "C" int pc_probe_conv3x3_ss(const float* a, int cqa, int a_H, int a_W, int a_oy, int a_ox, const float* b, int cqb,
                                   int b_H, int b_W, int b_oy, int b_ox, const float* wflat, int cout, int relu, int H, int W,
                                   float* out_c4, float* out_planar, float* pool_c4, int tile_rows, void* stream) {
    if (!a || !wflat || H < 1 || W < 1 || (cout != 8 && cout != 16)) return PC_ERR_INVALID;
    const int cin = 4 * (cqa + cqb);
    std::vector<uint8_t> img(conv_ss_img_bytes(cin, cout));
    conv_ss_pack_layer(wflat, cin, cout, img.data());
    uint8_t* d_img = nullptr;
    PC_CUDA(cudaMalloc(&d_img, img.size()));
    PC_CUDA(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    SsParams p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.TR = tile_rows & ~1;
    SsJob& J = p.jobs[0];
    J.wimg = d_img; J.out_c4 = out_c4; J.out_planar = out_planar; J.out_cs = (long long)H * W; J.out_rs = W; J.pool_c4 = pool_c4;
    J.a_oy = a_oy; J.a_ox = a_ox; J.b_oy = b_oy; J.b_ox = b_ox; J.linear = relu ? 0 : 1;
    if (!make_tmap_c4(&p.tmA[0], a, cqa, a_H, a_W)) return PC_ERR_INVALID;
    if (cqb > 0 && !make_tmap_c4(&p.tmB[0], b, cqb, b_H, b_W)) return PC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = conv_ss_launch(cqa, cqb, cout, pool_c4 ? PEPI_POOL : PEPI_STORE, p, 1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_img);
    if (rc) return rc;
    return e == cudaSuccess ? 0 : (int)e;
}
