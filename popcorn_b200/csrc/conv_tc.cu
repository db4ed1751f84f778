// 3x3 convolutions of the DDA UNet as implicit GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), split operands
// (x = hi + lo, three products per MAC; fp16 halves by default — two K elements per TMEM column, K = 16 per UMMA — or TF32 halves with
// -DPC_TC_F16=0; the text below counts K in TF32 elements, halve the columns and k-steps for fp16).
//
// Mapping (one persistent CTA per SM walks tiles of 128 columns x TR rows of one (image, stream) job):
//   * UMMA M = 128 = the 128 pixels of one image row segment; pixel x0+t == TMEM lane t;
//   * K of one input row = (kx, ci): the row is written to TMEM THREE times, shifted by -1/0/+1 pixel, each value
//     split x = hi + lo -> A operand [128 x 3*Cin] hi and lo, in TMEM;
//   * the ky shift is NOT a lane shift: input row r feeds output rows r+1 (ky=0), r (ky=1), r-1 (ky=2), i.e. three
//     different fp32 accumulators D[128 x 16] that live in an 8-slot TMEM ring (slot = running output row % 8);
//   * B operand = the folded weights [W_ky2 | W_ky1 | W_ky0][co][(kx,ci)] (48 rows: 3 x Cout 16, or Cout 8 zero-padded),
//     pre-split and pre-swizzled on the host (K-major SWIZZLE_128B), copied to shared memory once per CTA; the three
//     accumulators of rows r-1, r, r+1 are adjacent ring slots, so ONE N=48 UMMA per k-step and split term feeds all
//     three (two UMMAs where the ring wraps; fewer rows at tile borders).  Accumulators are always accumulated into:
//     the epilogue zeroes a slot after reading it;
//   * per input row: 3*Cin/8 (TF32) | ceil(3*Cin/16) (fp16) k-steps x 3 (hi*hi + lo*hi + hi*lo) tcgen05.mma, N/2 = 24 cycles each
//     (tools/probe/umma_probe.cu) -> 27*Cin cycles per 128 pixels.
// Warp-specialised pipeline (576 threads), every hand-off is an mbarrier, nothing is block-synchronous:
//   warp 17     TMA     : one lane streams input rows (all channels, 136 floats: 128 px + halo, zero-filled outside
//                         the source = conv padding / the Up block's F.pad) into an NS-deep shared-memory ring
//   warps 0-7   stagers : two groups alternate rows: wait s_full -> 3*Cin ld.shared (x-1, x, x+1) -> split ->
//                         tcgen05.st into one of NA A buffers -> arrive s_empty, full_a
//   warps 16,17 MMA     : one lane each; issuer q owns the output row pairs with (pair & 1) == q (disjoint accumulators),
//                         waits full_a, issues its half of the batch's UMMAs, commits to empty_a and, for a finished
//                         pair, to d_full — while one issuer waits or commits the other keeps the tensor pipe fed
//   warps 8-15  epilogue: two groups alternate row pairs: wait d_full -> tcgen05.ld -> arrive d_empty -> bias + ReLU
//                         -> store | 2x2 maxpool | 1x1 logit dot + sigmoid (+ crop)
// Precision: ~21-bit operands, fp32 accumulation — what the 1e-2 per-pixel bar needs; plain
// single-pass TF32 fails it.  The CUDA cores only move data (3*Cin shared loads + splits per pixel instead of
// 9*Cin*Cout FMAs), so the layers become HBM-bound ((Cin+Cout)*4 B per pixel).
// The reflect-padded, channel-remapped first layer (Cin 2 | 4, HBM-bound at fp32 SIMT) stays on conv.cu's stencil.
//
// Replaces model/DDA_model/utils/networks.py:253-271 (DoubleConv: Conv2d 3x3 pad 1 + BatchNorm2d(eval) + ReLU),
// :284-295 (MaxPool2d in Down), :318 (skip concat), :323-330 (OutConv) and popcorn.py:317-320.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

#ifndef PC_TC_PROBE
#define PC_TC_PROBE 0              // 1: CTA (0,0) accumulates per-role cycle counters into g_tc_dbg (development only)
#endif

namespace pc {

#if PC_TC_PROBE
__device__ long long g_tc_dbg[32];
__device__ long long g_tc_trace[4096];   // [role 0..7][batch 0..63][event 0..7] clock64 stamps of CTA (0,0)
#define TCP_TRACE(role, batch, ev) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (batch) >= 64 && (batch) < 128) g_tc_trace[((role) * 64 + (batch) - 64) * 8 + (ev)] = clock64(); } while (0)
#define TCP_T(var) const long long var = clock64()
#define TCP_DECL long long tcp_acc[6] = {0, 0, 0, 0, 0, 0}
#define TCP_ADD(slot, a, b) tcp_acc[(slot) & 7] += (b) - (a)
#define TCP_FLUSH(base) do { if (blockIdx.x == 0 && blockIdx.y == 0) for (int q = 0; q < 6; ++q) g_tc_dbg[(base) + q] = tcp_acc[q]; } while (0)
#else
#define TCP_TRACE(role, batch, ev)
#define TCP_T(var)
#define TCP_DECL
#define TCP_ADD(slot, a, b)
#define TCP_FLUSH(base)
#endif

constexpr int TCM = 128;           // pixels per UMMA
constexpr int TCN = 16;            // UMMA N (output channels, zero-padded)
// timing experiments of the probe build (results are then WRONG on purpose): 1 = no UMMAs, 2 = no tcgen05.st in the stagers,
// 4 = stagers read no shared memory, 8 = epilogue skips tcgen05.ld / re-zeroing, 16 = no full-resolution stores, 32 = no pooled stores
#ifndef PC_TC_EXP
#define PC_TC_EXP 0
#endif
#ifndef PC_TC_ST16
#define PC_TC_ST16 1               // stagers write the A operand with 16-column tcgen05.st
#endif
#ifndef PC_TC_ND8
#define PC_TC_ND8 8                // accumulator ring depth (output rows in flight) of the Cout = 8 kernels
#endif
// Occupancy: <8,0,8,store> and <8,0,8,dot> run TWO CTAs per SM (15 warps each: 2 stager groups, ONE epilogue group, 2 MMA warps, TMA;
// 256 TMEM columns) — two complete pipelines whose hand-off latencies interleave, the same remedy as the head's two tile contexts:
// 9.8 -> 7.9 and 12.0 -> 11.5 ms per bench step.  Measured and NOT adopted for <8,0,8,pool> (+10 %: one epilogue group does the pooling,
// 64-register cap spills), <8,0,16> (+5 %) and <8,8,8> (+26 %: only two A buffers fit 256 columns); those keep one 19-warp CTA.
// PC_TC_OCC2 bit 0: store / dot, bit 1: pool, bit 2: <8,0,16>, bit 3: <8,8,8> (A/B builds).
#ifndef PC_TC_OCC2
#define PC_TC_OCC2 1
#endif
#ifndef PC_TC_POOL_SPLIT
#define PC_TC_POOL_SPLIT 1         // pooling epilogue: the two lanes of a pooled pixel split the channels
#endif
#ifndef PC_TC_PAIRS16
#define PC_TC_PAIRS16 1            // <8,8,8> (Cin 16, Cout 8): 1 = two issuers with pair ownership, 0 = one issuer with 3-row windows
#endif
__host__ __device__ constexpr int tc_occ(int cin_a, int cin_b, int cout, int epi) {
    return (cin_a <= 8 && cin_b == 0 && cout == 8 && (epi == EPI_STORE || epi == EPI_DOT) && (PC_TC_OCC2 & 1)) ? 2
         : (cin_a == 8 && cin_b == 0 && cout == 8 && epi == EPI_POOL && (PC_TC_OCC2 & 2)) ? 2
         : (cin_a == 8 && cin_b == 0 && cout == 16 && (PC_TC_OCC2 & 4)) ? 2
         : (cin_a == 8 && cin_b == 8 && cout == 8 && (PC_TC_OCC2 & 8)) ? 2 : 1;
}

template <int CIN, int COUT, int OCC = 1>
struct TcGeom {
    static constexpr int TMEM_COLS = 512 / OCC;                 // tensor-memory columns of one CTA
    static constexpr int NEG = OCC == 2 ? 1 : 2;                // epilogue groups (4 warps each: one per TMEM lane quarter)
    static constexpr int THREADS = (8 + 4 * NEG + 3) * 32;      // 2 stager groups + epilogue groups + 2 MMA warps + TMA warp
    static constexpr int W_EPI0 = 8, W_MMA = 8 + 4 * NEG, W_TMA = W_MMA + 2;
    static constexpr int SLOTW = COUT == 8 ? 8 : 16;            // accumulator columns per output row
    static constexpr int ND = COUT == 8 ? PC_TC_ND8 : 8;        // accumulator ring: output rows in flight
    // MMA issue: two issuers that own alternate output row pairs (more UMMAs in flight per SM: one thread sustains one UMMA per
    // ~24 clk in this pipeline, the tensor pipe takes one per 9-13), or ONE issuer with 3-row windows (fewer, wider UMMAs).
    // Measured per layer shape (profiles/r2_conv_pipeline.md): windows win for Cin 8 and Cin 32, pairs for Cin 16 and for Cout 16.
    static constexpr bool TWO_ISSUERS = COUT == 16 || (CIN == 16 && PC_TC_PAIRS16);
    static constexpr int BROWS = COUT == 8 ? 24 : 48;           // rows of a B matrix: [W_ky2 | W_ky1 | W_ky0] x Cout (see conv_tc_pack_layer)
    // A columns (32-bit) per half (hi | lo) of one input row: one per K element (TF32) or one per two (fp16, K padded to 16)
    static constexpr int KROW = PC_TC_F16 ? (3 * CIN + 15) / 16 * 8 : (3 * CIN + 7) / 8 * 8;
    static constexpr int KSTEPS = KROW / 8;                     // one UMMA consumes 8 columns = 32 B of K either way
    static constexpr int KATOMS = (KROW + 31) / 32;             // 128-byte swizzle atoms along K
    static constexpr int BATOM = BROWS * 128;                   // bytes of one K-atom
    static constexpr int BMAT = KATOMS * BATOM;                 // bytes of one swizzled [BROWS x KROW] matrix
    static constexpr int OFF_BIAS = 2 * BMAT;                   // matrices: [hi, lo]
    static constexpr int IMG_BYTES = OFF_BIAS + 64;             // + bias[16]
    static constexpr int A_COLS = 2 * KROW;                     // one A buffer: hi at [0, KROW), lo at [KROW, 2*KROW)
    static constexpr int NA_FIT = (TMEM_COLS - ND * SLOTW) / A_COLS;
    static constexpr int NA = NA_FIT >= 8 ? 8 : NA_FIT >= 4 ? 4 : 2;   // A buffers (power of two): Cin 8 -> 8 (four row pairs in flight:
                                                                // the stagers run ahead of the UMMAs of the two batches before), Cin 16 -> 4, Cin 32 -> 2
    static constexpr int D_COL0 = NA * A_COLS;                  // accumulator slots of 16 columns
    static constexpr int STAGE_BYTES = CIN * TC_BOXW * 4;       // one input row, all channels
    static constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;   // ring slot pitch: a TMA destination is 128-byte aligned (Cin 2: 1088 -> 1152)
    static constexpr int NS = CIN <= 8 ? 16 : CIN <= 16 ? 8 : 4;   // shared-memory ring depth (~70 KB in flight per SM)
    static constexpr int OFF_STAGE = (IMG_BYTES + 127) / 128 * 128;
    static constexpr int OFF_BARS = OFF_STAGE + NS * STAGE_STRIDE;   // s_full[NS] s_empty[NS] full_a[NA/2] empty_a[NA/2] d_full[ND/2] d_empty[ND/2]
    static constexpr int NBARS = 2 * NS + 2 * NA + 2 * ND;
    static constexpr int OFF_TMEM = OFF_BARS + 8 * NBARS;
    static constexpr int OFF_CTW = (OFF_TMEM + 16 + 15) / 16 * 16;  // EPI_CONVT: [COUT][4][COUT] + bias[COUT] floats of the transposed conv
    static constexpr int CTW_FLOATS = COUT * 4 * COUT + COUT;
    static constexpr int SMEM_NEED = OFF_CTW + CTW_FLOATS * 4 + 1024;
    // OCC 1: > half of the SM's shared memory, i.e. exactly one CTA per SM (it owns all 512 TMEM columns); OCC 2: what it needs
    static constexpr int SMEM_BYTES = OCC == 2 ? SMEM_NEED : (SMEM_NEED > 116 * 1024 ? SMEM_NEED : 116 * 1024);
    static_assert(NA_FIT >= 2, "TMEM budget");
    static_assert(D_COL0 % 16 == 0 && D_COL0 + ND * SLOTW <= TMEM_COLS, "accumulator ring placement");
    static_assert(SMEM_BYTES * OCC <= 226 * 1024, "shared memory layout");
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int c, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(c), "r"(bar) : "memory");
}

template <int CIN_A, int CIN_B, int COUT, int EPI>
__global__ void __launch_bounds__((TcGeom<CIN_A + CIN_B, COUT, tc_occ(CIN_A, CIN_B, COUT, EPI)>::THREADS), tc_occ(CIN_A, CIN_B, COUT, EPI))
conv3x3_tc_kernel(const __grid_constant__ TcConvParams p) {
    constexpr int CIN = CIN_A + CIN_B;
    using G = TcGeom<CIN, COUT, tc_occ(CIN_A, CIN_B, COUT, EPI)>;
    constexpr int WS_THREADS = G::THREADS, W_EPI0 = G::W_EPI0, W_MMA = G::W_MMA, W_TMA = G::W_TMA, NGROUP = G::NEG, TMEM_ALL = G::TMEM_COLS;
    constexpr int NA = G::NA, NS = G::NS, NP = NA / 2, ND = G::ND, NDP = ND / 2;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET into the shared array: a pointer that went through uintptr_t arithmetic loses its address space
    // and every access through it compiles to a generic LD/ST (the conv stagers' 24 loads per pixel: ~400 of their ~500 cycles per row)
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const ConvJob& job = p.jobs[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const float* bias = reinterpret_cast<const float*>(sm + G::OFF_BIAS);
    const uint32_t bars = smem_u32(sm + G::OFF_BARS);
    // rows move through the pipeline in PAIRS (input rows 2b-1, 2b of a tile; output rows 2m, 2m+1): one barrier
    // round trip of the single MMA-issuing thread then covers two rows of UMMAs
    auto s_full = [&](int i) { return bars + 8u * (uint32_t)i; };                        // TMA -> stagers, per row
    auto s_empty = [&](int i) { return bars + 8u * (uint32_t)(NS + i); };                // stagers -> TMA, per row
    auto full_a = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + i); };             // stagers -> MMA, per pair of A buffers
    auto empty_a = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + NP + i); };       // MMA (commit) -> stagers
    auto d_full = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + 2 * NP + i); };    // MMA (commit) -> epilogue, per pair of slots
    auto d_empty = [&](int i) { return bars + 8u * (uint32_t)(2 * NS + 2 * NP + NDP + i); };   // epilogue -> stagers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + G::OFF_TMEM);

    for (int i = tid; i < G::IMG_BYTES / 16; i += WS_THREADS)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(job.wtc) + i);
    if (EPI == EPI_CONVT)
        for (int i = tid; i < G::CTW_FLOATS / 4; i += WS_THREADS)
            reinterpret_cast<float4*>(sm + G::OFF_CTW)[i] = __ldg(reinterpret_cast<const float4*>(job.ctw) + i);
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_ALL);
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 4); }     // 4 = the stager warps of a group
        for (int i = 0; i < NP; ++i) { mbar_init(full_a(i), 8); mbar_init(empty_a(i), G::TWO_ISSUERS ? 2 : 1); }   // 8 = both stager groups; 2 = both MMA issuers commit (or the only one)
        for (int i = 0; i < NDP; ++i) { mbar_init(d_full(i), 1); mbar_init(d_empty(i), 4); }    // 4 = the epilogue warps of a group
        mbar_init_fence();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weights + barriers (generic stores) -> visible to UMMA / TMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tD = tbase + G::D_COL0;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;    // a warp may touch TMEM lanes 32*(warp%4) .. +31
    if (warp >= W_EPI0 && warp < W_EPI0 + 4) {                      // UMMAs only ever accumulate: start from zero
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int i = 0; i < ND * G::SLOTW / 16; ++i) tmem_st16(tD + lane_off + 16 * i, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int px = (warp & 3) * 32 + lane;                          // pixel of this thread inside a tile row

    const int H = p.H, W = p.W, TR = p.TR;                          // TR is even
    const int ntiles = p.tiles_x * p.tiles_y;
    // rows of a tile, rounded up to even: an odd last row is computed (from zero-filled input) and never stored
    auto tile_rows = [&](int tile) {
        const int y0 = (tile / p.tiles_x) * TR;
        const int n = (H - y0) < TR ? (H - y0) : TR;
        return (n + 1) & ~1;
    };

    if (warp < W_EPI0) {
        // =========================== stagers: shared-memory row -> TMEM (A operand) ===========================
        const int group = warp >> 2;                                 // group 0 stages input rows 2b-1, group 1 rows 2b
        const float* stage0 = reinterpret_cast<const float*>(sm + G::OFF_STAGE) + px + 3;   // box column 3 = image column x-1
        int B = 0, P0 = 0;                                           // running batch index / first output pair of the tile
        TCP_DECL;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int nb = tile_rows(tile) / 2 + 1;                  // batches of this tile: (nrows + 2) / 2
#pragma unroll 1
            for (int b = 0; b < nb; ++b, ++B) {
                const int i = 2 * B + group;                         // running input-row index = ring position
                const int s = i % NS, pb = B & (NP - 1), n = B / NP;
                TCP_T(t0);
                if ((warp & 3) == 0 && lane == 0) TCP_TRACE(group, B, 0);
                {
                    // the row has landed (s_full) + the UMMAs that read this A buffer pair are done (empty_a) + the accumulator
                    // slots of the output pair this batch opens are drained (d_empty): one merged wait
                    const uint32_t m1 = s_full(s), p1 = (uint32_t)(i / NS) & 1u;
                    uint32_t m2 = m1, p2 = p1, m3 = m1, p3 = p1;
                    if (n >= 1) { m2 = empty_a(pb); p2 = (uint32_t)(n - 1) & 1u; }
                    const int P = P0 + b;
                    if (b < nb - 1 && P >= NDP) { m3 = d_empty(P & (NDP - 1)); p3 = (uint32_t)(P / NDP - 1) & 1u; }
#if PC_TC_PROBE
                    {   // probe build: the three waits one after the other, timed separately (slot 5 = empty_a, slot 1 = d_empty)
                        mbar_wait_sleep(m1, p1);
                        const long long ta = clock64();
                        mbar_wait_sleep(m2, p2);
                        const long long tb = clock64();
                        mbar_wait_sleep(m3, p3);
                        const long long tc_ = clock64();
                        tcp_acc[5] += tb - ta; tcp_acc[1] += tc_ - tb; tcp_acc[0] -= tc_ - ta;
                    }
#else
                    mbar_wait3_sleep(m1, p1, m2, p2, m3, p3);
#endif
                }
                TCP_T(t1);
                tc_fence_after();
                TCP_T(t2);
                if ((warp & 3) == 0 && lane == 0) TCP_TRACE(group, B, 3);
                const float* st = stage0 + s * (G::STAGE_STRIDE / 4);
                const uint32_t tA = tbase + (uint32_t)(2 * pb + group) * G::A_COLS + lane_off;
                auto a_elem = [&](int k) {                                                   // A element k = kx * CIN + ci
                    return (k < 3 * CIN) ? ((PC_TC_EXP & 4) ? (float)(k + px) : st[(k % CIN) * TC_BOXW + k / CIN]) : 0.f;
                };
                auto load_split = [&](int col, uint32_t& hi, uint32_t& lo) {
                    if (PC_TC_F16 && 2 * col >= 3 * CIN) { hi = 0u; lo = 0u; }                // K padding (the asm split is opaque to constant folding)
                    else if (PC_TC_F16) split_f16x2(a_elem(2 * col), a_elem(2 * col + 1), hi, lo);
                    else split_tf32(a_elem(col), hi, lo);
                };
#if PC_TC_ST16
                // 16-column tcgen05.st (64 B per lane and instruction): the .x8 form sustains only ~126 B/clk per SM here
                if (G::KROW % 16 == 0) {
#pragma unroll
                    for (int j = 0; j < G::KROW / 16; ++j) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) load_split(16 * j + q, hi[q], lo[q]);
                        if (PC_TC_EXP & 2) { asm volatile("" ::"r"(hi[0] ^ hi[5] ^ hi[15] ^ lo[0] ^ lo[7] ^ lo[15])); continue; }
                        tmem_st16(tA + 16 * j, hi);
                        tmem_st16(tA + G::KROW + 16 * j, lo);
                    }
                } else {                                            // KROW = 24 | 8: [hi | lo] = 48 | 16 contiguous columns, three stores | one
                    static_assert(G::KROW % 16 == 0 || G::KROW == 24 || G::KROW == 8, "stager store shapes");
                    uint32_t w[2 * G::KROW];
#pragma unroll
                    for (int q = 0; q < G::KROW; ++q) load_split(q, w[q], w[G::KROW + q]);
                    if (!(PC_TC_EXP & 2)) {
#pragma unroll
                        for (int j = 0; j < 2 * G::KROW / 16; ++j) tmem_st16(tA + 16 * j, reinterpret_cast<uint32_t(&)[16]>(w[16 * j]));
                    } else {
                        asm volatile("" ::"r"(w[0] ^ w[G::KROW - 1] ^ w[G::KROW] ^ w[2 * G::KROW - 1]));
                    }
                }
#else
#pragma unroll
                for (int j = 0; j < G::KSTEPS; ++j) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) load_split(8 * j + q, hi[q], lo[q]);
                    if (PC_TC_EXP & 2) { asm volatile("" ::"r"(hi[0] ^ hi[1] ^ hi[2] ^ hi[3] ^ hi[4] ^ hi[5] ^ hi[6] ^ hi[7] ^ lo[0] ^ lo[1] ^ lo[2] ^ lo[3] ^ lo[4] ^ lo[5] ^ lo[6] ^ lo[7])); continue; }
                    tmem_st8(tA + 8 * j, hi);
                    tmem_st8(tA + G::KROW + 8 * j, lo);
                }
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty(s));                               // the ring slot may be refilled
                TCP_T(t3);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a(pb));
                TCP_T(t4);
                if ((warp & 3) == 0 && lane == 0) TCP_TRACE(group, B, 4);
                TCP_ADD(0, t0, t1); TCP_ADD(2, t2, t3); TCP_ADD(3, t3, t4); TCP_ADD(4, t4 - 1, t4);
            }
            P0 += nb - 1;
        }
        if (tid == 0) TCP_FLUSH(8);
    } else if (warp < W_MMA) {
        // =========================== epilogue: TMEM accumulators -> bias + ReLU -> global ===========================
        const int group = (warp - W_EPI0) >> 2;
        float dotw[EPI == EPI_DOT ? 8 : 1];
        float dotb = 0.f;
        if (EPI == EPI_DOT) {
#pragma unroll
            for (int o = 0; o < 8; ++o) dotw[o] = __ldg(job.dotw + o);
            dotb = __ldg(job.dotw + 8);
        }
        int P0 = 0;
        TCP_DECL;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            const int y0 = ty * TR, npairs = tile_rows(tile) / 2;
            const int vx = tx * TCM + px;
            const bool in_x = vx < W;
#pragma unroll 1
            for (int m = group; m < npairs; m += NGROUP) {                           // row pairs alternate between the groups
                const int P = P0 + m, ps = P & (NDP - 1);
                TCP_T(t0);
                if (((warp - W_EPI0) & 3) == 0 && lane == 0) TCP_TRACE(4 + group, P, 0);
                // EPI_DOT, second stream: the first stream's partial logits are fetched BEFORE the wait for the accumulators, so the DRAM round
                // trip overlaps the UMMAs instead of sitting in the (single, at OCC 2) epilogue group's per-pair critical path
                float din[2] = {0.f, 0.f};
                if (EPI == EPI_DOT && job.dot_in) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int oy = y0 + 2 * m + h, yy = oy - p.crop_y, xx = vx - p.crop_x;
                        if (in_x && oy < H && yy >= 0 && yy < p.crop_H && xx >= 0 && xx < p.crop_W)
                            din[h] = __ldg(job.dot_in + (long long)yy * job.dot_in_rs + xx);
                    }
                }
                mbar_wait_sleep(d_full(ps), (uint32_t)(P / NDP) & 1u);
                tc_fence_after();
                TCP_T(t1);
                if (((warp - W_EPI0) & 3) == 0 && lane == 0) TCP_TRACE(4 + group, P, 1);
                uint32_t d[2][COUT];
                const uint32_t tpair = tD + lane_off + 2 * G::SLOTW * (uint32_t)ps;      // the pair's two slots are adjacent
                if (COUT == 16) {
                    tmem_ld16(tpair, reinterpret_cast<uint32_t(&)[16]>(d[0]));
                    tmem_ld16(tpair + 16, reinterpret_cast<uint32_t(&)[16]>(d[1]));
                } else {
                    tmem_ld16(tpair, reinterpret_cast<uint32_t(&)[16]>(d[0]));             // d[0][0..7] = row 2m, d[1][0..7] = row 2m+1
                }
                tc_wait_ld();
                {
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = 0u;
                    tmem_st16(tpair, z);                                                   // the next rows using the slots accumulate from zero
                    if (COUT == 16) tmem_st16(tpair + 16, z);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty(ps));            // the slot pair may be re-opened
                TCP_T(t2);
                if (((warp - W_EPI0) & 3) == 0 && lane == 0) TCP_TRACE(4 + group, P, 2);
                TCP_ADD(0, t0, t1); TCP_ADD(1, t1, t2); TCP_ADD(2, t2 - 1, t2);
                float acc[2][COUT];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) { const float t = __uint_as_float(d[h][o]) + bias[o]; acc[h][o] = job.linear ? t : fmaxf(t, 0.f); }
                const int xx = vx - p.crop_x;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int oy = y0 + 2 * m + h;
                    const int yy = oy - p.crop_y;
                    const bool inside = in_x && oy < H && yy >= 0 && yy < p.crop_H && xx >= 0 && xx < p.crop_W;
                    if (EPI == EPI_DOT) {
                        if (inside) {
                            float sacc = 0.f;
#pragma unroll
                            for (int o = 0; o < 8; ++o) sacc = fmaf(acc[h][o], dotw[o], sacc);
                            sacc += din[h];
                            if (job.dot_final) {
                                sacc += dotb;
                                sacc = 1.f / (1.f + expf(-sacc));
                            }
                            job.dot_out[(long long)yy * job.dot_out_rs + xx] = sacc;
                        }
                    } else if (job.out && inside && !(PC_TC_EXP & 16)) {
                        float* dst = job.out + (long long)yy * job.out_rs + xx;
#pragma unroll
                        for (int o = 0; o < COUT; ++o) dst[(long long)o * job.out_cs] = acc[h][o];
                    }
                }
                if (EPI == EPI_CONVT) {
                    // ConvTranspose2d(k=2, s=2) of the Up block on the two activated rows this thread holds (networks.py:302):
                    // up[co][2y+dy][2x+dx] = b[co] + sum_ci act[ci][y][x] * w[ci][dy*2+dx][co] — purely per pixel, so it rides in
                    // the epilogue (FP32 pipe, otherwise idle here) and the low-resolution activation never travels to HBM.
                    const float* ctw = reinterpret_cast<const float*>(sm + G::OFF_CTW);
                    const float* ctb = ctw + COUT * 4 * COUT;
                    const bool row_ok[2] = {in_x && (y0 + 2 * m) < H, in_x && (y0 + 2 * m + 1) < H};
#pragma unroll 1
                    for (int cg = 0; cg < COUT; cg += 8) {
#pragma unroll 1
                        for (int dy = 0; dy < 2; ++dy) {
                            // packed fp32 (fma.rn.f32x2): two adjacent output channels per instruction, weights read as 64-bit pairs
                            unsigned long long u2[2][2][4];          // [row h][dx][co pair]
#pragma unroll
                            for (int o = 0; o < 4; ++o) {
                                const unsigned long long bo = *reinterpret_cast<const unsigned long long*>(ctb + cg + 2 * o);
                                u2[0][0][o] = bo; u2[0][1][o] = bo; u2[1][0][o] = bo; u2[1][1][o] = bo;
                            }
#pragma unroll
                            for (int ci = 0; ci < COUT; ++ci) {
                                const unsigned long long a0 = pack2(acc[0][ci], acc[0][ci]), a1 = pack2(acc[1][ci], acc[1][ci]);
#pragma unroll
                                for (int dx = 0; dx < 2; ++dx) {
                                    const float* wr = ctw + (ci * 4 + dy * 2 + dx) * COUT + cg;
                                    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wr);
                                    const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(wr + 4);
                                    fma2(u2[0][dx][0], a0, wa.x); fma2(u2[0][dx][1], a0, wa.y); fma2(u2[0][dx][2], a0, wb.x); fma2(u2[0][dx][3], a0, wb.y);
                                    fma2(u2[1][dx][0], a1, wa.x); fma2(u2[1][dx][1], a1, wa.y); fma2(u2[1][dx][2], a1, wb.x); fma2(u2[1][dx][3], a1, wb.y);
                                }
                            }
                            float u[2][2][8];                        // [row h][dx][co]
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                                    for (int o = 0; o < 4; ++o) unpack2(u2[h][dx][o], u[h][dx][2 * o], u[h][dx][2 * o + 1]);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (!row_ok[h]) continue;
                                const int oy = y0 + 2 * m + h;
                                float* dst = job.ct_out + (long long)cg * job.ct_cs + (long long)(2 * oy + dy) * job.ct_rs + 2 * vx;
#pragma unroll
                                for (int o = 0; o < 8; ++o)
                                    *reinterpret_cast<float2*>(dst + (long long)o * job.ct_cs) = make_float2(u[h][0][o], u[h][1][o]);
                            }
                        }
                    }
                }
                if (EPI == EPI_POOL) {                               // 2x2 max over (rows 2m, 2m+1) x (lanes 2k, 2k+1)
                    const int py = (y0 >> 1) + m, pxl = vx >> 1;
#if PC_TC_POOL_SPLIT
                    // the two lanes of a pooled pixel share the channels: the even lane finishes and stores channels [0, COUT/2), the odd
                    // lane [COUT/2, COUT) — half the shuffles and stores of "every lane reduces all channels, even lanes store"
                    const bool stp = py < (H >> 1) && pxl < (W >> 1);
                    const bool odd = lane & 1;
                    float* pdst = job.pool + (long long)(odd ? COUT / 2 : 0) * job.pool_cs + (long long)py * job.pool_rs + pxl;
#pragma unroll
                    for (int o = 0; o < COUT / 2; ++o) {
                        const float va = fmaxf(acc[0][o], acc[1][o]), vb = fmaxf(acc[0][o + COUT / 2], acc[1][o + COUT / 2]);
                        const float other = __shfl_xor_sync(FULL, odd ? va : vb, 1);      // the partner's half of MY channels
                        const float hm = fmaxf(odd ? vb : va, other);
                        if (stp && !(PC_TC_EXP & 32)) pdst[(long long)o * job.pool_cs] = hm;
                    }
#else
                    const bool stp = !(lane & 1) && py < (H >> 1) && pxl < (W >> 1);
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        const float vm = fmaxf(acc[0][o], acc[1][o]);
                        const float hm = fmaxf(vm, __shfl_xor_sync(FULL, vm, 1));
                        if (stp && !(PC_TC_EXP & 32)) job.pool[(long long)o * job.pool_cs + (long long)py * job.pool_rs + pxl] = hm;
                    }
#endif
                }
            }
            P0 += npairs;
        }
        if (tid == W_EPI0 * 32) TCP_FLUSH(16);
    } else if (warp < W_TMA) {
        if (elect_one()) {
            // =========================== MMA issuers: one lane each of two warps ===========================
            // Issuer q owns the output row pairs with (running pair index & 1) == q, i.e. its own accumulator slots, so the two
            // never accumulate into the same TMEM columns; while one waits on a barrier or commits, the other keeps the tensor
            // pipe fed.  A batch (input rows 2b-1, 2b) feeds exactly two pairs:
            //   pair b-1 (rows 2b-2, 2b-1): row 2b-1 -> N=32 with B blocks [ky2, ky1];  row 2b -> N=16 (row 2b-1) with block ky2
            //   pair b   (rows 2b, 2b+1)  : row 2b-1 -> N=16 (row 2b) with block ky0;   row 2b -> N=32 with B blocks [ky1, ky0]
            const int q = warp - W_MMA;
            const uint32_t sW = smem_u32(sm);
            constexpr uint32_t ID16 = tc_idesc(TCM, TCN), ID32 = tc_idesc(TCM, 2 * TCN);
            const uint64_t bd_hi = make_bdesc(sW), bd_lo = make_bdesc(sW + G::BMAT);
            constexpr uint64_t BLK = (uint64_t)((TCN * 128) >> 4);              // one 16-row block of B, in descriptor address units
            auto issue = [&](uint32_t d, uint32_t tAhi, uint64_t boff, uint32_t idesc) {
                const uint32_t tAlo = tAhi + G::KROW;
#pragma unroll
                for (int j = 0; j < G::KSTEPS; ++j) {
                    const uint64_t koff = boff + (uint64_t)(((j >> 2) * G::BATOM + (j & 3) * 32) >> 4);   // address field: 16-byte units
                    if (PC_TC_EXP & 1) continue;
                    umma_ts(d, tAhi + 8 * j, bd_hi + koff, idesc, 1u);
                    umma_ts(d, tAlo + 8 * j, bd_hi + koff, idesc, 1u);
                    umma_ts(d, tAhi + 8 * j, bd_lo + koff, idesc, 1u);
                }
            };
            int B = 0, P0 = 0;
            const bool active = G::TWO_ISSUERS || q == 0;                   // windows: issuer 1 has no work (see below)
            TCP_DECL;
#pragma unroll 1
            for (int tile = blockIdx.x; active && tile < ntiles; tile += gridDim.x) {
                const int npairs = tile_rows(tile) / 2, nb = npairs + 1;
#pragma unroll 1
                for (int b = 0; b < nb; ++b, ++B) {
                    const int pb = B & (NP - 1);
                    TCP_T(t0);
                    TCP_TRACE(2 + q, B, 0);
                    mbar_wait_sleep(full_a(pb), (uint32_t)(B / NP) & 1u);     // both rows staged (and the slots they open drained)
                    tc_fence_after();
                    TCP_T(t1);
                    TCP_TRACE(2 + q, B, 1);
                    const uint32_t tA0 = tbase + (uint32_t)(2 * pb) * G::A_COLS, tA1 = tA0 + G::A_COLS;
                    const int Pl = P0 + b - 1, Pu = P0 + b;                   // lower / upper pair fed by this batch
                    if (COUT == 16) {
                        // B rows [W_ky2 | W_ky1 | W_ky0] (16 each): two adjacent rows of a pair = one N=32 UMMA
                        if ((Pl & 1) == q) {
                            if (b >= 1) {
                                const uint32_t d = tD + 32 * (uint32_t)(Pl & (NDP - 1));
                                issue(d, tA0, 0, ID32);                        // rows 2b-2 (ky2), 2b-1 (ky1)
                                issue(d + 16, tA1, 0, ID16);                   // row 2b-1 (ky2)
                            }
                        } else if (b < npairs) {
                            const uint32_t d = tD + 32 * (uint32_t)(Pu & (NDP - 1));
                            issue(d, tA0, 2 * BLK, ID16);                      // row 2b (ky0)
                            issue(d, tA1, BLK, ID32);                          // rows 2b (ky1), 2b+1 (ky0)
                        }
                    } else if (G::TWO_ISSUERS) {
                        // Cout 8, pair ownership: 8-column slots, a pair = 16 columns; B rows [W_ky2 | W_ky1 | W_ky0] (8 each)
                        constexpr uint32_t ID8 = tc_idesc(TCM, 8);
                        constexpr uint64_t BLK8 = (uint64_t)((8 * 128) >> 4);   // one 8-row block of B, in descriptor address units
                        if ((Pl & 1) == q) {
                            if (b >= 1) {
                                const uint32_t d = tD + 16 * (uint32_t)(Pl & (NDP - 1));
                                issue(d, tA0, 0, ID16);                        // rows 2b-2 (ky2), 2b-1 (ky1)
                                issue(d + 8, tA1, 0, ID8);                     // row 2b-1 (ky2)
                            }
                        } else if (b < npairs) {
                            const uint32_t d = tD + 16 * (uint32_t)(Pu & (NDP - 1));
                            issue(d, tA0, 2 * BLK8, ID8);                      // row 2b (ky0)
                            issue(d, tA1, BLK8, ID16);                         // rows 2b (ky1), 2b+1 (ky0)
                        }
                    } else if (q == 0) {
                        // Cout 8, windows: 8-column slots; ONE issuer (the accumulation order of an output row is then program order, i.e.
                        // reproducible).  Input row i feeds output rows i-1 (ky 2), i (ky 1), i+1 (ky 0): three ADJACENT ring slots, so
                        // one N = 24 UMMA per k-step and split term against B = [W_ky2 | W_ky1 | W_ky0] (12.5 clk on the tensor pipe,
                        // where two N = 16 UMMAs cost 19: tools/probe/umma_n_probe.cu — M = 128 accepts N = 8 / 24 on sm_100a).  The
                        // window is clipped at the tile's first / last row and split in two where the 8-row ring wraps.
                        constexpr uint32_t ID8 = tc_idesc(TCM, 8), ID24 = tc_idesc(TCM, 24);
                        constexpr uint64_t BLK8 = (uint64_t)((8 * 128) >> 4);   // one 8-row block of B, in descriptor address units
                        const int nrows = 2 * npairs;
                        auto window = [&](int i, uint32_t tA) {
                            const int lo = i - 1 < 0 ? 0 : i - 1, hi = i + 1 > nrows - 1 ? nrows - 1 : i + 1;
                            const int n = hi - lo + 1, blk0 = lo - (i - 1);
                            const int slot = (2 * P0 + lo) & (ND - 1);
                            const int run1 = n < ND - slot ? n : ND - slot;
                            issue(tD + 8 * (uint32_t)slot, tA, (uint64_t)blk0 * BLK8, run1 == 1 ? ID8 : run1 == 2 ? ID16 : ID24);
                            if (run1 < n) issue(tD, tA, (uint64_t)(blk0 + run1) * BLK8, (n - run1) == 1 ? ID8 : ID16);
                        };
                        window(2 * b - 1, tA0);
                        window(2 * b, tA1);
                    }
                    TCP_T(t2);
                    TCP_TRACE(2 + q, B, 2);
                    if (G::TWO_ISSUERS || q == 0) umma_commit(empty_a(pb));    // the A buffer pair may be refilled (both issuers commit)
                    if (b >= 1 && (G::TWO_ISSUERS ? (Pl & 1) == q : q == 0)) umma_commit(d_full(Pl & (NDP - 1)));   // output rows 2b-2, 2b-1 are final
                    TCP_T(t3);
                    TCP_TRACE(2 + q, B, 3);
                    TCP_ADD(0, t0, t1); TCP_ADD(1, t1, t2); TCP_ADD(2, t2, t3); TCP_ADD(3, t3 - 1, t3); TCP_ADD(4, t3 - 1, t3);
                }
                P0 += npairs;
            }
            if (q == 0) TCP_FLUSH(0);
        }
    } else if (elect_one()) {
        // =========================== TMA producer: global rows -> shared-memory ring ===========================
        const uint32_t stage_base = smem_u32(sm + G::OFF_STAGE);
        const CUtensorMap* tmA = &p.tmA[blockIdx.y];
        const CUtensorMap* tmB = &p.tmB[blockIdx.y];
        int i = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            const int x0 = tx * TCM, y0 = ty * TR, nrows = tile_rows(tile);
#pragma unroll 1
            for (int r = -1; r <= nrows; ++r, ++i) {
                const int s = i % NS, n = i / NS;
                if (n >= 1) mbar_wait_sleep(s_empty(s), (uint32_t)(n - 1) & 1u);
                const uint32_t dst = stage_base + (uint32_t)s * G::STAGE_STRIDE;
                mbar_expect_tx(s_full(s), G::STAGE_BYTES);
                tma_load_3d(dst, tmA, x0 - 4 - job.a_ox, y0 + r - job.a_oy, 0, s_full(s));
                if (CIN_B > 0) tma_load_3d(dst + CIN_A * TC_BOXW * 4, tmB, x0 - 4 - job.b_ox, y0 + r - job.b_oy, 0, s_full(s));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tbase, TMEM_ALL);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int tc_brows(int cout) { return cout == 8 ? 24 : 48; }

static int tc_krow(int cin) { return PC_TC_F16 ? (3 * cin + 15) / 16 * 8 : (3 * cin + 7) / 8 * 8; }   // TcGeom::KROW

static int tc_geom_img_floats(int cin, int cout) {
    const int krow = tc_krow(cin), katoms = (krow + 31) / 32;
    return (2 * katoms * tc_brows(cout) * 128 + 64) / 4;
}

int conv_tc_layer_floats(int cin, int cout) { return (int)round_up(tc_geom_img_floats(cin, cout), 64); }   // 256-B multiple

// flat = [cin][ky][kx][cout] + bias[cout] (the SIMT pack)  ->  [hi|lo][katom][rows][32 floats] swizzled + bias[16], k = kx*cin + ci.
//   rows [W_ky2 | W_ky1 | W_ky0] (row cout*(2-ky) + co): 48 rows for cout 16, 24 for cout 8 — the issuers address 8- or 16-row
//   aligned windows of it for one, two or three adjacent output rows (conv3x3_tc_kernel).
void conv_tc_pack_layer(const float* flat, int cin, int cout, float* img) {
    const int krow = tc_krow(cin), katoms = (krow + 31) / 32;
    const int rows = tc_brows(cout);
    const int mat = katoms * rows * 32;                 // floats per matrix
    const int total = conv_tc_layer_floats(cin, cout);
    memset(img, 0, sizeof(float) * total);
    // (row offset, ky) placements of the 8/16-row weight blocks
    int place[3][2], nplace = 0;
    for (int ky = 0; ky < 3; ++ky) { place[nplace][0] = (2 - ky) * cout; place[nplace][1] = ky; ++nplace; }
    for (int pi = 0; pi < nplace; ++pi) {
        const int r0 = place[pi][0], ky = place[pi][1];
        for (int co = 0; co < cout; ++co)
            for (int kx = 0; kx < 3; ++kx)
                for (int ci = 0; ci < cin; ++ci) {
                    const float w = flat[((ci * 3 + ky) * 3 + kx) * cout + co];
                    const int n = r0 + co;
                    const int k = kx * cin + ci;
#if PC_TC_F16
                    // fp16 halves, 64 per 128-byte row of an atom; Swizzle<3,4,3>: 16-byte chunk (8 halves) ^= row % 8
                    const __half hi = __float2half_rn(w);
                    const __half lo = __float2half_rn(w - __half2float(hi));
                    const int atom = k / 64, kk = k % 64;
                    const int pos = (((kk / 8) ^ (n % 8)) * 8) + kk % 8;
                    const int idx = atom * (rows * 64) + n * 64 + pos;
                    __half* himg = reinterpret_cast<__half*>(img);
                    himg[idx] = hi;
                    himg[2 * mat + idx] = lo;                                // mat counts floats: the lo matrix starts 2 * mat halves in
#else
                    uint32_t bits;
                    memcpy(&bits, &w, 4);
                    bits &= 0xFFFFE000u;
                    float hi;
                    memcpy(&hi, &bits, 4);
                    const float lo = w - hi;
                    const int atom = k / 32, kk = k % 32;
                    const int pos = (((kk / 4) ^ (n % 8)) * 4) + kk % 4;      // Swizzle<3,4,3>: 16-B chunk ^= row % 8
                    const int idx = atom * (rows * 32) + n * 32 + pos;
                    img[idx] = hi;
                    img[mat + idx] = lo;
#endif
                }
    }
    for (int n = 0; n < cout; ++n) img[2 * mat + n] = flat[cin * 9 * cout + n];
}

#if PC_TC_PROBE
}  // namespace pc
extern "C" int pc_debug_tc_counters(long long* out32, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, pc::g_tc_dbg, sizeof(long long) * 32);
    if (reset == 2) cudaMemcpyFromSymbol(out32, pc::g_tc_trace, sizeof(long long) * 4096);
    if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(pc::g_tc_dbg, z, sizeof(z)); }
    return 0;
}
namespace pc {
#endif

bool conv_tc_enabled() {
    static const bool on = [] {
        const char* e = getenv("POPCORN_CONV_TC");
        return e ? atoi(e) != 0 : true;
    }();
    return on;
}

static int conv_tc_rows() {
    static const int tr = [] {
        const char* e = getenv("POPCORN_CONV_TC_ROWS");
        int v = e ? atoi(e) : 0;                        // 0 = choose per launch
        if (v <= 0) return 0;
        if (v < 2) v = 2;
        return v & ~1;                                  // even: rows move through the pipeline in pairs
    }();
    return tr;
}

template <int CIN_A, int CIN_B, int COUT, int EPI>
static int launch_tc_impl(TcConvParams& p, int njobs, cudaStream_t st) {
    using G = TcGeom<CIN_A + CIN_B, COUT, tc_occ(CIN_A, CIN_B, COUT, EPI)>;
    constexpr int OCC = tc_occ(CIN_A, CIN_B, COUT, EPI);
    static const int cat = [] {
        char nm[64];
        snprintf(nm, sizeof(nm), "conv3x3_tc<%d,%d,%d,%s>", CIN_A, CIN_B, COUT, EPI == EPI_STORE ? "store" : EPI == EPI_POOL ? "pool" : EPI == EPI_DOT ? "dot" : "convt");
        return prof_register(nm);
    }();
    auto k = conv3x3_tc_kernel<CIN_A, CIN_B, COUT, EPI>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int per_job = OCC * num_sms() / njobs;                            // persistent CTAs of one job, OCC CTAs per SM
    if (per_job < 1) per_job = 1;
    p.tiles_x = cdiv(p.W, TCM);
    p.TR = conv_tc_rows();
    if (p.TR == 0)                                                    // auto: tall tiles (3 % halo rows) once every CTA still gets >= 8 of them
        p.TR = ((long long)p.tiles_x * cdiv(p.H, 64) >= 8ll * per_job) ? 64 : 32;
    p.tiles_y = cdiv(p.H, p.TR);
    const int ntiles = p.tiles_x * p.tiles_y;
    if (per_job > ntiles) per_job = ntiles;
    {
        ProfScope prof(cat, st, (double)p.H * p.W * njobs);
        k<<<dim3(per_job, njobs), G::THREADS, G::SMEM_BYTES, st>>>(p);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

int launch_conv_tc(int cin_a, int cin_b, int cout, int epi, TcConvParams& p, int njobs, cudaStream_t st) {
    const int key = ((cin_a * 100 + cin_b) * 100 + cout) * 10 + epi;
    switch (key) {
        case ((2 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<2, 0, 8, EPI_STORE>(p, njobs, st);   // first layer, SAR stream
        case ((4 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<4, 0, 8, EPI_STORE>(p, njobs, st);   // first layer, optical stream
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<8, 0, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_POOL: return launch_tc_impl<8, 0, 8, EPI_POOL>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_DOT: return launch_tc_impl<8, 0, 8, EPI_DOT>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 16) * 10 + EPI_STORE: return launch_tc_impl<8, 0, 16, EPI_STORE>(p, njobs, st);
        case ((16 * 100 + 0) * 100 + 16) * 10 + EPI_STORE: return launch_tc_impl<16, 0, 16, EPI_STORE>(p, njobs, st);
        case ((16 * 100 + 0) * 100 + 16) * 10 + EPI_POOL: return launch_tc_impl<16, 0, 16, EPI_POOL>(p, njobs, st);
        case ((16 * 100 + 16) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<16, 16, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 8) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<8, 8, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_CONVT: return launch_tc_impl<8, 0, 8, EPI_CONVT>(p, njobs, st);
        case ((16 * 100 + 0) * 100 + 16) * 10 + EPI_CONVT: return launch_tc_impl<16, 0, 16, EPI_CONVT>(p, njobs, st);
    }
    set_error("launch_conv_tc: no instantiation for (%d,%d,%d,%d)", cin_a, cin_b, cout, epi);
    return PC_ERR_INVALID;
}

}  // namespace pc
