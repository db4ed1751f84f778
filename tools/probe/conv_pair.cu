// ROUND-2 CANDIDATE — written without a GPU at the end of round 1, never run.  Built into its own library
// (tools/probe/build_conv_pair.sh -> popcorn_b200/libpopcorn_b200_probe.so), NOT into libpopcorn_b200.so.
//
// 3x3 convolution (+ folded BN bias + ReLU, optional 2x2 max-pool) on "fp16 pair" activations, without stager warps.
//
// Why (profiles/r1c_precision_study.md): the shipped tensor-core conv (csrc/conv_tc.cu) spends its time in the
// stager warps — 3*Cin ld.shared + a TF32 split + tcgen05.st per pixel — and on the single TMEM port.  On the parity tests' weights
// (reference init + real DDA checkpoint) the per-pixel bar (1e-2) needs >= 19 operand bits: two bf16 pieces (16 bits) FAIL it, two
// fp16 pieces per value (x = h1 + h2, 22 bits, |x| < 65504) sit at the fp32 noise floor like 3xTF32 does.  If an activation is
// STORED as its pair (h1 | h2 << 16, one 32-bit word), memory already holds the tensor-core operand:
//   * layout [C/4][H][W][4 words]: a pixel's 4 channels = 16 bytes = one row of a K-major no-swizzle UMMA core matrix;
//   * one 4-D TMA box {4 words, 136 px, 1 row, C/4 chunks} per input row lands in shared memory as the canonical operand:
//     pixel stride 16 B, 8-row groups SBO = 128 B apart, 16-byte K chunks LBO = 136*16 B apart (zero-filled outside the image =
//     the conv's zero padding / the Up block's F.pad);
//   * UMMA kind::f16 (fp16 x fp16 -> fp32 in TMEM), M = 128 pixels, K = 16 = 8 channels x (h1, h2); the kx tap is a +16-byte
//     shift of the A descriptor's start address; the ky taps are accumulator columns: input row r feeds output rows r-1, r, r+1
//     = three adjacent 16-column slots of an 8-slot TMEM ring through ONE N = 48 instruction with B rows [W_ky2 | W_ky1 | W_ky0];
//   * weights are duplicated over the (h1, h2) slots and split w = w1 + w2 into two fp16 images: A*B(w1) + A*B(w2) =
//     (h1 + h2)(w1 + w2), all four partial products, 2 UMMAs per (tap, 8 channels);
//   * the epilogue (tcgen05.ld -> bias -> ReLU) packs each value back into a pair and stores 16-byte pixel chunks (coalesced).
// Roles: warp 0 = TMA producer (one lane), warp 1 = UMMA issuer (one lane), warps 2-9 = two epilogue groups that alternate
// output-row pairs.  No register-level staging of the operands at all.
//
// Open questions the hardware has to answer first (tools/probe/pair_probe.cu): descriptor field meaning for this layout, cost of
// an SS-form N = 48 UMMA whose A (4 KB) comes from shared memory (if it is A-read bound: tcgen05.cp the row into TMEM once and use
// the TS form, or N = 96 with both weight images side by side).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"

namespace pc {

constexpr int PJOBS = 8;
constexpr int PBOX = 136;                      // staged pixels per row: x0-1 .. x0+134 (136 * 16 B = 17 * 128 B keeps every chunk 128-B aligned)
constexpr int PCHUNK = PBOX * 16;              // bytes of one 4-channel chunk of a staged row = LBO of the A descriptor
constexpr int PND = 8;                         // accumulator ring: output rows in flight (16 TMEM columns each)
constexpr int PBROWS = 48;                     // B rows: [W_ky2 | W_ky1 | W_ky0] x 16 output channels (Cout 8 zero-padded)
constexpr int PTHREADS = 10 * 32;
enum { PEPI_STORE = 0, PEPI_POOL = 1 };

struct PairJob {
    const uint8_t* wimg;                       // packed weights (conv_pair_pack_layer), device
    uint32_t* out_pair;                        // [COUT/4][H][W][4] pair words, or null
    float* out_planar; long long out_cs; int out_rs;   // planar fp32 output (the layer that feeds the head), or null
    uint32_t* pool_pair;                       // [COUT/4][H/2][W/2][4] pair words (PEPI_POOL)
    int a_oy, a_ox, b_oy, b_ox;                // source offsets (the Up block's zero-padded upsampled branch)
    int linear;                                // 1: no ReLU
};
struct alignas(64) PairParams {
    CUtensorMap tmA[PJOBS], tmB[PJOBS];
    int H, W, TR, tiles_x, tiles_y;
    PairJob jobs[PJOBS];
};

__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t n) {       // D f32 (bits 4-5 = 1), A = B = fp16 (format 0), K-major
    return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((M >> 4) << 24);
}
// K-major no-swizzle matrix descriptor: start>>4 | LBO>>4 @16 (between 16-byte K chunks) | SBO>>4 @32 (between 8-row groups) | version 1 @46
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// 128 lanes x 256 bits (one K = 16 slice of a 16-bit A operand = 8 TMEM columns): shared memory (matrix descriptor) -> TMEM
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// x -> (h1, h2): h1 = fp16(x) round-to-nearest-even, h2 = fp16(x - h1); packed h1 | h2 << 16 (h1 = the even K element).
// 22 significant bits while |x| < 65504 (values beyond saturate instead of becoming inf) and x - h1 is not subnormal.
__device__ __forceinline__ uint32_t pack_pair(float v) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h1 = __float2half_rn(v);
    const __half h2 = __float2half_rn(v - __half2float(h1));
    return (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
}

template <int CQ, int COUT>
struct PairGeom {
    static constexpr int KSTEPS = CQ / 2;                          // K = 16 instructions per tap: 8 channels each
    static constexpr int IMG_HALF = 3 * CQ * PBROWS * 16;          // one weight image [kx][chunk][48 rows][16 B]
    static constexpr int OFF_BIAS = 2 * IMG_HALF;
    static constexpr int W_BYTES = OFF_BIAS + 64;
    static constexpr int STAGE = CQ * PCHUNK;                      // one input row, all chunks (multiple of 128 B)
    static constexpr int NS = CQ <= 2 ? 16 : CQ <= 4 ? 10 : 5;     // rows in flight (~70-87 KB)
    static constexpr int OFF_STAGE = (W_BYTES + 127) / 128 * 128;
    static constexpr int OFF_BARS = OFF_STAGE + NS * STAGE;        // s_full[NS] s_empty[NS] d_full[PND] d_empty[PND]
    static constexpr int OFF_TMEM = OFF_BARS + 8 * (2 * NS + 2 * PND);
    static constexpr int SMEM_NEED = OFF_TMEM + 16 + 1024;
    static constexpr int SMEM_BYTES = SMEM_NEED > 116 * 1024 ? SMEM_NEED : 116 * 1024;   // one CTA per SM
    static_assert(CQ % 2 == 0 && SMEM_BYTES <= 227 * 1024, "geometry");
};

// A_CP = false: SS form, the UMMAs read A from shared memory (4 KB per instruction, twice per tap: once per weight image).
// A_CP = true : each 128 x K16 slice of the staged row is copied ONCE into TMEM by tcgen05.cp (no registers involved) and the UMMAs
//               use the TS form; two A buffers alternate between input rows.  Relies on the tensor pipe executing the cp / mma
//               stream of the single issuing thread in order (cp -> mma is documented; the mma -> cp WAR two rows later is the
//               thing to confirm on hardware: if not, wait on the s_empty commit of row i-2 before the copies).
template <int CQA, int CQB, int COUT, int EPI, bool A_CP>
__global__ void __launch_bounds__(PTHREADS, 1) conv3x3_pair_kernel(const __grid_constant__ PairParams p) {
    constexpr int CQ = CQA + CQB;
    using G = PairGeom<CQ, COUT>;
    constexpr int NS = G::NS;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const PairJob& job = p.jobs[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const float* bias = reinterpret_cast<const float*>(sm + G::OFF_BIAS);
    const uint32_t bars = smem_u32(sm + G::OFF_BARS);
    auto s_full = [&](int i) { return bars + 8u * (uint32_t)i; };                    // TMA (tx bytes) -> issuer
    auto s_empty = [&](int i) { return bars + 8u * (uint32_t)(NS + i); };            // issuer (commit) -> TMA
    auto d_full = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + i); };         // issuer (commit) -> epilogue, per output row slot
    auto d_empty = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + PND + i); };  // epilogue (4 warps) -> issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + G::OFF_TMEM);

    for (int i = tid; i < G::W_BYTES / 16; i += PTHREADS)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(job.wimg) + i);
    constexpr uint32_t TCOLS = A_CP ? 512u : 128u;                    // accumulator ring [0,128) (+ two A buffers of 24 * KSTEPS columns from 128)
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TCOLS);
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 1); }
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
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_full(s)), "r"(G::STAGE) : "memory");
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
                        // the row has landed (s_full) + the slot of output row r+1, which this row opens, has been drained (d_empty):
                        // ONE merged probe — a blocking mbarrier probe costs ~200 cycles even when the phase completed long ago
                        const uint32_t m1 = s_full(s), p1 = (uint32_t)(i / NS) & 1u;
                        uint32_t m2 = m1, p2 = p1;
                        const int g = g0 + r + 1;
                        if (r + 1 <= nrows - 1 && g >= PND) { m2 = d_empty(g % PND); p2 = (uint32_t)(g / PND - 1) & 1u; }
                        mbar_wait3_sleep(m1, p1, m2, p2, m2, p2);
                    }
                    tc_fence_after();
                    const uint32_t sA = stage_base + (uint32_t)s * G::STAGE;
                    const uint32_t tA = tbase + 128u + (uint32_t)(i & 1) * (24u * G::KSTEPS);
                    if (A_CP) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int j = 0; j < G::KSTEPS; ++j)
                                tmem_cp_128x256b(tA + 8u * (uint32_t)(kx * G::KSTEPS + j), desc_nosw(sA + kx * 16 + 2 * j * PCHUNK, PCHUNK, 128));
                    }
                    const int lo = r - 1 < 0 ? 0 : r - 1, hi = r + 1 > nrows - 1 ? nrows - 1 : r + 1;
                    int o = lo;
                    while (o <= hi) {                                 // runs of output rows whose ring slots are adjacent (the ring wraps)
                        const int slot = (g0 + o) % PND;
                        int n = hi - o + 1;
                        if (n > PND - slot) n = PND - slot;
                        const uint32_t d = tbase + 16u * (uint32_t)slot;
                        const uint32_t brow = 16u * (uint32_t)(o - r + 1);          // output row o takes tap ky = r - o + 1 = B block 2 - ky
                        const uint32_t id = idesc_f16(128, 16u * (uint32_t)n);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int j = 0; j < G::KSTEPS; ++j) {
                                const uint64_t ad = desc_nosw(sA + kx * 16 + 2 * j * PCHUNK, PCHUNK, 128);
#pragma unroll
                                for (int img = 0; img < 2; ++img) {
                                    const uint64_t bd = desc_nosw(sW + img * G::IMG_HALF + ((kx * CQ + 2 * j) * PBROWS + brow) * 16, PBROWS * 16, 128);
                                    if (A_CP) umma_f16_ts(d, tA + 8u * (uint32_t)(kx * G::KSTEPS + j), bd, id);
                                    else umma_f16_ss(d, ad, bd, id);
                                }
                            }
                        o += n;
                    }
                    umma_commit(s_empty(s));                                       // the staged row may be overwritten
                    if (r >= 1) umma_commit(d_full((g0 + r - 1) % PND));           // output row r-1 has all three taps
                }
                g0 += nrows;
            }
        }
    } else {
        // =========================== epilogue: two groups alternate output-row pairs ===========================
        const int group = (warp - 2) >> 2;
        const int px = (warp & 3) * 32 + lane;                                     // TMEM lane == pixel; (warp & 3) is also the lane quarter
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
                    if (job.out_pair) {
#pragma unroll
                        for (int q = 0; q < COUT / 4; ++q) {
                            const uint4 w4 = make_uint4(pack_pair(acc[h][4 * q]), pack_pair(acc[h][4 * q + 1]), pack_pair(acc[h][4 * q + 2]),
                                                        pack_pair(acc[h][4 * q + 3]));
                            reinterpret_cast<uint4*>(job.out_pair)[((size_t)q * H + oy) * W + vx] = w4;
                        }
                    }
                    if (job.out_planar) {
#pragma unroll
                        for (int o = 0; o < COUT; ++o) job.out_planar[(long long)o * job.out_cs + (long long)oy * job.out_rs + vx] = acc[h][o];
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
                        for (int q = 0; q < COUT / 4; ++q) {
                            const uint4 w4 = make_uint4(pack_pair(hm[4 * q]), pack_pair(hm[4 * q + 1]), pack_pair(hm[4 * q + 2]), pack_pair(hm[4 * q + 3]));
                            reinterpret_cast<uint4*>(job.pool_pair)[((size_t)q * pH + py) * pW + pxl] = w4;
                        }
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
static uint16_t h_f16_rn(float f) {
    const __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
static float h_f16_f(uint16_t u) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}

static int conv_pair_img_bytes(int cin, int /*cout*/) { return 2 * 3 * (cin / 4) * PBROWS * 16 + 64; }

// flat = [cin][ky][kx][cout] + bias[cout] (the SIMT pack, BN folded) -> [w1 | w2][kx][chunk][48 rows][4 ch x (w, w)] fp16 + bias[16] fp32
static void conv_pair_pack_layer(const float* flat, int cin, int cout, uint8_t* img) {
    const int cq = cin / 4, half = 3 * cq * PBROWS * 8;            // uint16 elements of one image
    memset(img, 0, conv_pair_img_bytes(cin, cout));
    uint16_t* w16 = reinterpret_cast<uint16_t*>(img);
    for (int kx = 0; kx < 3; ++kx)
        for (int q = 0; q < cq; ++q)
            for (int ky = 0; ky < 3; ++ky)
                for (int co = 0; co < cout; ++co)
                    for (int e = 0; e < 4; ++e) {
                        const float w = flat[(((q * 4 + e) * 3 + ky) * 3 + kx) * cout + co];
                        const uint16_t w1 = h_f16_rn(w), w2 = h_f16_rn(w - h_f16_f(w1));
                        const size_t at = ((size_t)(kx * cq + q) * PBROWS + (2 - ky) * 16 + co) * 8 + 2 * e;
                        w16[at] = w16[at + 1] = w1;
                        w16[half + at] = w16[half + at + 1] = w2;
                    }
    float* b = reinterpret_cast<float*>(img + 2 * half * 2);
    for (int co = 0; co < cout; ++co) b[co] = flat[cin * 9 * cout + co];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// pair tensor [CQ][H][W][4 words] -> 4-D map, box {4, PBOX, 1, CQ}
static bool make_tmap_pair(CUtensorMap* tm, const uint32_t* ptr, int cq, int H, int W) {
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
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<uint32_t*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CQA, int CQB, int COUT, int EPI>
static int launch_pair(PairParams& p, int njobs, cudaStream_t st) {
    using G = PairGeom<CQA + CQB, COUT>;
    static const bool a_cp = [] { const char* e = getenv("POPCORN_PAIR_A_CP"); return e && atoi(e) != 0; }();
    auto k = a_cp ? conv3x3_pair_kernel<CQA, CQB, COUT, EPI, true> : conv3x3_pair_kernel<CQA, CQB, COUT, EPI, false>;
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

void set_error(const char* fmt, ...) { (void)fmt; }     // probe library: errors are returned as codes only

}  // namespace pc

using namespace pc;

// One conv layer on pair tensors (device pointers); `wflat` is a HOST pointer to [cin][3][3][cout] + bias[cout].
// a: [cqa][a_H][a_W][4] words, b (optional): [cqb][b_H][b_W][4]; outputs any of: out_pair [cout/4][H][W][4], out_planar [cout][H][W]
// fp32, pool_pair [cout/4][H/2][W/2][4].  tile_rows: 0 = automatic.
extern "C" int pc_probe_conv3x3_pair(const uint32_t* a, int cqa, int a_H, int a_W, int a_oy, int a_ox, const uint32_t* b, int cqb,
                                     int b_H, int b_W, int b_oy, int b_ox, const float* wflat, int cout, int relu, int H, int W,
                                     uint32_t* out_pair, float* out_planar, uint32_t* pool_pair, int tile_rows, void* stream) {
    if (!a || !wflat || H < 1 || W < 1 || (cout != 8 && cout != 16)) return PC_ERR_INVALID;
    const int cin = 4 * (cqa + cqb);
    std::vector<uint8_t> img(conv_pair_img_bytes(cin, cout));
    conv_pair_pack_layer(wflat, cin, cout, img.data());
    uint8_t* d_img = nullptr;
    PC_CUDA(cudaMalloc(&d_img, img.size()));
    PC_CUDA(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.TR = tile_rows & ~1;
    PairJob& J = p.jobs[0];
    J.wimg = d_img; J.out_pair = out_pair; J.out_planar = out_planar; J.out_cs = (long long)H * W; J.out_rs = W; J.pool_pair = pool_pair;
    J.a_oy = a_oy; J.a_ox = a_ox; J.b_oy = b_oy; J.b_ox = b_ox; J.linear = relu ? 0 : 1;
    if (!make_tmap_pair(&p.tmA[0], a, cqa, a_H, a_W)) return PC_ERR_INVALID;
    if (cqb > 0 && !make_tmap_pair(&p.tmB[0], b, cqb, b_H, b_W)) return PC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int key = (cqa * 10 + cqb) * 100 + cout;
    int rc = PC_ERR_INVALID;
    if (pool_pair) {
        if (key == 2008) rc = launch_pair<2, 0, 8, PEPI_POOL>(p, 1, st);
        else if (key == 4016) rc = launch_pair<4, 0, 16, PEPI_POOL>(p, 1, st);
    } else {
        switch (key) {
            case 2008: rc = launch_pair<2, 0, 8, PEPI_STORE>(p, 1, st); break;
            case 2016: rc = launch_pair<2, 0, 16, PEPI_STORE>(p, 1, st); break;
            case 4016: rc = launch_pair<4, 0, 16, PEPI_STORE>(p, 1, st); break;
            case 4408: rc = launch_pair<4, 4, 8, PEPI_STORE>(p, 1, st); break;
            case 2208: rc = launch_pair<2, 2, 8, PEPI_STORE>(p, 1, st); break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_img);
    if (rc) return rc;
    return e == cudaSuccess ? 0 : (int)e;
}
