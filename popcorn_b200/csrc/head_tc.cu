// Occupancy head on the 5th-gen tensor cores (tcgen05 + TMEM), error-compensated split operands (x = hi + lo, three products per MAC).
//
// Per 128-pixel tile (UMMA M = 128, N = 64, cta_group::1), 256 worker threads:
//   * a pixel = one TMEM lane; TWO warps share each 32-lane quarter (warp w and w+4) and split the per-pixel work by
//     columns (features 0-7 | 8-15, hidden units 0-31 | 32-63), which halves every epilogue between two layers' UMMAs;
//   * activations are the A operand and LIVE IN TMEM: the epilogue reads the fp32 accumulator D with tcgen05.ld, applies ReLU,
//     splits x = hi + lo into fp16 pairs (PC_TC_F16, default; two K elements per TMEM column) or TF32 halves and writes both
//     halves back with tcgen05.st — they never touch shared or global memory;
//   * weights are the B operand in shared memory, pre-split (hi/lo) and pre-swizzled on the host into the
//     canonical K-major SWIZZLE_128B layout, copied in verbatim once per CTA;
//   * every k-step (16 fp16 / 8 tf32 elements) issues three tcgen05.mma (hi*hi + lo*hi + hi*lo) into the same fp32
//     accumulator: 22-bit operands, which is what the per-pixel 1e-2 bar needs (tools/precision_split_study.py);
//     fp16 build: the layer's bias is one more k-step (constant (1, 0, ..) columns of A x the bias row of the weight image);
//   * the 64->1 output layer, ReLU, x builtup, stores and the census partial sums run in the last epilogue.
// One persistent CTA per SM with three (fp16) or two (tf32) tile contexts and one UMMA-issuing warp: see "Kernel structure" below.
// Replaces model/popcorn.py:79-88, 160-190, 195-228 (same contract as head.cu's SIMT kernel).
#include "head_common.cuh"
#include "tc_common.cuh"

namespace pc {

#ifndef PC_HEAD_SPIN
#define PC_HEAD_SPIN 0           // 1: poll the hand-off barriers with mbarrier.test_wait instead of the suspending try_wait
#endif
#if PC_HEAD_SPIN
#define HEAD_WAIT mbar_wait
#else
#define HEAD_WAIT mbar_wait_sleep
#endif
#ifndef PC_HEAD_EXP
#define PC_HEAD_EXP 0            // timing experiments (results wrong on purpose): 1 = no feature loads, 2 = no output stores, 4 = no builtup load
#endif
#ifndef PC_HEAD_PROBE
#define PC_HEAD_PROBE 0          // 1: CTA 0 accumulates per-role cycle counters (development only)
#endif
#if PC_HEAD_PROBE
__device__ long long g_head_dbg[32];
#define HP_T(v) const long long v = clock64()
#define HP_ADD(slot, a, b) do { if (blockIdx.x == 0) hp_acc[slot] += (b) - (a); } while (0)
#else
#define HP_T(v)
#define HP_ADD(slot, a, b)
#endif
constexpr int TM = 128;          // pixels per tile
constexpr int HT = 256;          // threads per CTA: two warps per TMEM lane quarter
constexpr uint32_t IDESC = tc_idesc(128, 64);       // kind::tf32 or kind::f16 (PC_TC_F16, tc_common.cuh)

// byte offsets inside the packed TC weight image (host: weights.pack_head_tc)
constexpr int OFF_W1HI = 0, OFF_W1LO = 8192, OFF_W2HI = 16384, OFF_W2LO = 32768, OFF_W3HI = 49152, OFF_W3LO = 65536;
constexpr int OFF_VEC = 81920;                 // b1[64] b2[64] b3[64] w4[64] b4[4]
constexpr int TC_PACK_BYTES = OFF_VEC + 260 * 4;
constexpr int OFF_MBAR = TC_PACK_BYTES;        // 8-byte mbarrier
constexpr int OFF_TMEM = OFF_MBAR + 8;         // 4-byte TMEM base address slot
constexpr int OFF_PART = OFF_TMEM + 8;         // float[128]: output-layer partial dot of the upper-half warps
constexpr int TC_SMEM_BYTES = OFF_PART + 512 + 1024;   // + slack to align the base to 1024 B

// one hidden layer's UMMAs: D[128x64] = A[128xK] * W[64xK]^T, K in steps of 8 (TF32) or 16 (fp16) = 8 TMEM columns of A and 32 bytes
// of a B row either way, three split terms per step
// fp16 build: the layer's bias rides in the UMMAs — the A operand carries a constant column pair (1, 0) behind its K columns and the
// weight image the bias (hi | lo) as K row Kpad — instead of 32 FADDs per thread and layer in epilogues that are bound by instruction
// issue (the tensor pipe has the headroom: 12 -> 14 UMMAs per hidden layer)
constexpr bool BIAS_MMA = PC_TC_F16 != 0;
template <int K>
__device__ __forceinline__ void issue_layer(uint32_t tD, uint32_t tAhi, uint32_t tAlo, uint32_t sWhi, uint32_t sWlo,
                                            uint32_t mbar) {
#pragma unroll
    for (int j = 0; j < (PC_TC_F16 ? (K + 15) / 16 : K / 8); ++j) {
        const uint32_t koff = (uint32_t)((j >> 2) * 8192 + (j & 3) * 32);   // 128-byte swizzle atoms along K
        const uint64_t bhi = make_bdesc(sWhi + koff), blo = make_bdesc(sWlo + koff);
        umma_ts(tD, tAhi + 8 * j, bhi, IDESC, j > 0 ? 1u : 0u);
        umma_ts(tD, tAlo + 8 * j, bhi, IDESC, 1u);
        umma_ts(tD, tAhi + 8 * j, blo, IDESC, 1u);
    }
    if (BIAS_MMA) {
        constexpr int j = (K + 15) / 16;                                    // the k-step behind the layer's own K: (1, 0 | zeros) x bias row
        const uint32_t koff = (uint32_t)((j >> 2) * 8192 + (j & 3) * 32);
        umma_ts(tD, tAhi + 8 * j, make_bdesc(sWhi + koff), IDESC, 1u);
        umma_ts(tD, tAhi + 8 * j, make_bdesc(sWlo + koff), IDESC, 1u);
    }
    umma_commit(mbar);
}

// hidden-layer epilogue of one thread: its 32 columns of D -> relu(D + bias) -> (hi, lo) -> A operand of the next layer
__device__ __forceinline__ void epilogue_hidden(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float* bias) {
    // 16 columns at a time, hi split in place: 32 live registers (the kernel is capped at 96 by its 17 warps: registers are
    // allocated per four warps, 65 536 / 640 threads)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t v[16];
        tmem_ld16(tD + 16 * q, v);
        tc_wait_ld();
#if PC_TC_F16
        // fp16 halves: hidden units (2i, 2i+1) share a column — the thread's 32 units become 16 + 16 columns (tAhi / tAlo point at them)
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            split_f16x2(fmaxf(__uint_as_float(v[2 * i]), 0.f), fmaxf(__uint_as_float(v[2 * i + 1]), 0.f), hi[i], lo[i]);   // bias: BIAS_MMA
        tmem_st8(tAhi + 8 * q, hi);
        tmem_st8(tAlo + 8 * q, lo);
#else
        uint32_t lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_tf32(fmaxf(__uint_as_float(v[i]) + bias[16 * q + i], 0.f), v[i], lo[i]);
        tmem_st16(tAhi + 16 * q, v);
        tmem_st16(tAlo + 16 * q, lo);
#endif
    }
}

// Kernel structure: ONE persistent CTA per SM, 8 * NCTX + 1 warps.  NCTX tile contexts (each: accumulator D, A_hi, A_lo, next tile's
// layer-1 operand = 160 TMEM columns with fp16 halves, 224 with TF32) are worked on by their own 8 worker warps; the last warp is the
// only UMMA issuer and serves the contexts in strict rotation (ctx0 layer 1, ctx1 layer 1, ctx2 layer 1, ctx0 layer 2, ...), so one
// context's epilogue always runs under the other contexts' UMMAs instead of all drifting into the same phase.  Hand-offs are mbarriers: workers -> issuer a_ready[c]
// (8 warp arrivals), issuer -> workers d_ready[c] (tcgen05.commit).
constexpr int HWORK = 256;                       // worker threads per context
#ifndef PC_HEAD_NCTX
#define PC_HEAD_NCTX (PC_TC_F16 ? 3 : 2)   // tile contexts per CTA: three fit the tensor memory with fp16 operands (160 columns each)
#endif
constexpr int NCTX = PC_HEAD_NCTX;
constexpr int HTHREADS = NCTX * HWORK + 32;
// tensor-memory columns of one context: accumulator D | A_hi | A_lo of the hidden layers | A1_hi | A1_lo (layer-1 features of the next tile)
// (fp16 build: the hi halves end with the 8 columns of the constant bias k-step)
constexpr uint32_t AHI_COLS = PC_TC_F16 ? 32 + 8 : 64, ALO_COLS = PC_TC_F16 ? 32 : 64, A1HI_COLS = PC_TC_F16 ? 8 + 8 : 16, A1LO_COLS = PC_TC_F16 ? 8 : 16;
constexpr uint32_t C_AHI = 64, C_ALO = C_AHI + AHI_COLS, C_A1HI = C_ALO + ALO_COLS, C_A1LO = C_A1HI + A1HI_COLS;
constexpr uint32_t CTX_COLS = (C_A1LO + A1LO_COLS + 31) / 32 * 32;
static_assert(NCTX * CTX_COLS <= 512 && NCTX >= 2 && NCTX <= 4, "tensor-memory budget of the tile contexts");
constexpr int OFF_BARS2 = OFF_MBAR;              // a_ready[2], d_ready[2]
constexpr int OFF_TMEM2 = OFF_BARS2 + 64;
constexpr int OFF_PART2 = OFF_TMEM2 + 16;        // float[NCTX][128]
constexpr int TC2_SMEM_BYTES = (OFF_PART2 + 512 * NCTX + 1024) > 116 * 1024 ? (OFF_PART2 + 512 * NCTX + 1024) : 116 * 1024;   // > half an SM: one CTA per SM

template <int K1, bool SPARSE, bool SMALL>
__global__ void __launch_bounds__(HTHREADS, 1) head_tc_kernel(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET into the shared array: a pointer that went through uintptr_t arithmetic loses its address space
    // and every access through it compiles to a generic LD/ST (the conv stagers' 24 loads per pixel: ~400 of their ~500 cycles per row)
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
    const float* b1 = vec, *b2 = vec + 64, *b3 = vec + 128, *w4 = vec + 192, *b4 = vec + 256;
    const uint32_t bars = smem_u32(sm + OFF_BARS2);
    auto a_ready = [&](int c) { return bars + 8u * (uint32_t)c; };
    auto d_ready = [&](int c) { return bars + 8u * NCTX + 8u * (uint32_t)c; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM2);
    const int tid = threadIdx.x, warp = uniform_warp_idx();

    for (int i = tid; i < TC_PACK_BYTES / 16; i += HTHREADS)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(a.pack) + i);
    if (warp == 8 * NCTX) tmem_alloc(smem_u32(tmem_slot), 512);
    if (tid == 0) {
        for (int c = 0; c < NCTX; ++c) { mbar_init(a_ready(c), 8); mbar_init(d_ready(c), 1); }
        mbar_init_fence();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weight image (generic stores) -> visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t sW = smem_u32(sm);

    const long long HW = SPARSE ? a.HW : (long long)a.H * a.W;
    const long long total = SPARSE ? (long long)__ldg(a.n_dev) : HW * a.B;
    const long long ntiles = (total + TM - 1) / TM;
    const long long stride = (long long)NCTX * gridDim.x;                       // tiles are dealt to (CTA, context) round-robin

    if (warp < 8 * NCTX) {
        // =========================== workers of context c ===========================
        const int c = warp >> 3;
        const int wl = warp & 7, lane = tid & 31;
        const int px = (wl & 3) * 32 + lane;            // pixel (= TMEM lane) of this thread inside a tile
        const int half = wl >> 2;                       // which half of the columns this thread works on
        constexpr int CH = K1 >= 16 ? K1 / 2 : K1;      // layer-1 feature channels per thread (K1 = 8: lower half stages all)
        const bool stager = K1 >= 16 || half == 0;
        const int c_lo = K1 >= 16 ? half * CH : 0;
        const uint32_t lane_off = (uint32_t)((wl & 3) * 32) << 16;     // a warp may touch TMEM lanes 32*(warp%4) .. +31
        const uint32_t tD = tbase + CTX_COLS * c, tAhi = tD + C_AHI, tAlo = tD + C_ALO;
        const uint32_t tA1hi = tD + C_A1HI, tA1lo = tD + C_A1LO;      // layer-1 operand (features) of the NEXT tile: own columns, staged early
        const uint32_t col_off = (uint32_t)(32 * half);                 // this thread's accumulator columns = hidden units
        const uint32_t acol_off = PC_TC_F16 ? col_off / 2 : col_off;    // ... and where they go in the next layer's A operand
        float* part = reinterpret_cast<float*>(sm + OFF_PART2) + 128 * c;
        uint32_t ph = 0;                                 // phase counter of both barriers of this context
        if (BIAS_MMA && half == 0) {                     // the constant k-step of both A operands: K element Kpad = 1.0, the rest zero
            const uint32_t one[8] = {0x00003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};      // fp16 pair (1.0, 0.0)
            tmem_st8(tAhi + lane_off + 32, one);
            tmem_st8(tA1hi + lane_off + 8, one);
            tc_wait_st();
        }

        // Where a thread's pixel is: (image, row, column) for the dense map, (image, pixel inside the image) for the sparse index list.
        // The dense position ADVANCES by a constant number of pixels per tile (stride * 128), so it is updated with two compare-and-wrap
        // steps instead of being re-derived from a linear index: the divisions (three uses x two divisions per tile and thread, ~100
        // instructions of a loop that is bound by instruction issue) happen once per thread.  The sparse position comes from the index
        // list and costs one 32-bit division per tile (the host picks SMALL whenever the batch has fewer than 2^31 pixels — always for
        // the int32 index list; the 64-bit software division costs ~150 instructions).
        struct Pos { int b, y, x; };
        int step_b = 0, step_y = 0, step_x = 0;
        if (!SPARSE) {
            const long long S = stride * TM;
            step_b = (int)(S / HW);
            const long long r = S - (long long)step_b * HW;
            step_y = (int)(r / a.W);
            step_x = (int)(r - (long long)step_y * a.W);
        }
        auto advance = [&](Pos q) {                      // dense: the same lane's pixel of the context's next tile
            q.x += step_x; if (q.x >= a.W) { q.x -= a.W; ++q.y; }
            q.y += step_y; if (q.y >= a.H) { q.y -= a.H; ++q.b; }
            q.b += step_b;
            return q;
        };
        auto locate_first = [&](long long tile, bool& valid) {
            Pos q{0, 0, 0};
            const long long i = tile * TM + px;
            if (SPARSE) {
                valid = tile < ntiles && i < total;
                if (valid) {
                    const uint32_t p = (uint32_t)__ldg(a.idx + i);
                    q.b = (int)(p / (uint32_t)HW);
                    q.x = (int)(p - (uint32_t)q.b * (uint32_t)HW);
                }
            } else {
                if (SMALL) {
                    q.b = (int)((uint32_t)i / (uint32_t)HW);
                    const uint32_t r = (uint32_t)i - (uint32_t)q.b * (uint32_t)HW;
                    q.y = (int)(r / (uint32_t)a.W);
                    q.x = (int)(r - (uint32_t)q.y * (uint32_t)a.W);
                } else {
                    const long long b = i / HW, r = i - b * HW;
                    q.b = b > a.B ? a.B : (int)b;
                    q.y = (int)(r / a.W);
                    q.x = (int)(r - (long long)q.y * a.W);
                }
                valid = q.b < a.B;
            }
            return q;
        };
        auto locate_next = [&](const Pos& cur, long long tile, bool& valid) {
            if (SPARSE) return locate_first(tile, valid);
            const Pos q = advance(cur);
            valid = q.b < a.B;                           // p < B * H * W  <=>  image index < B
            return q;
        };
        auto feat_offset = [&](const Pos& q) -> long long {
            return SPARSE ? q.b * a.f_bs + q.x : q.b * a.f_bs + (long long)q.y * a.f_rs + q.x;
        };
        auto hand_over = [&]() {                         // TMEM writes of this warp are done -> one arrival on a_ready[c]
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(c));
        };
        auto wait_d = [&]() { HEAD_WAIT(d_ready(c), ph & 1u); ++ph; tc_fence_after(); };

#if PC_HEAD_PROBE
        long long hp_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
        float fcur[CH];        // ONE register set: consumed at the top of a tile, refilled (next tile) while layer 2 runs — no copy, a
                               // MOV of a register an outstanding load still has to fill would wait for DRAM
        long long tile = (long long)NCTX * blockIdx.x + c;
        bool valid;
        Pos cur = locate_first(tile, valid);
        {
            const long long foff0 = valid ? feat_offset(cur) : 0;
#pragma unroll
            for (int k = 0; k < CH; ++k) fcur[k] = (valid && stager) ? __ldg(a.feats + foff0 + (long long)(c_lo + k) * a.f_cs) : 0.f;
        }

        // layer-1 A operand: this pixel's features (this thread's channel half), split, into the context's A1 columns
        auto stage_features = [&]() {
#if PC_TC_F16
            if (stager) {                                // CH = 8 channels -> 4 columns of fp16 pairs; K1 = 8 pads K to 16 with 4 zero columns
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) { split_f16x2(fcur[2 * k], fcur[2 * k + 1], hi[k], lo[k]); hi[4 + k] = 0u; lo[4 + k] = 0u; }
                if (K1 >= 16) {
                    tmem_st4(tA1hi + lane_off + c_lo / 2, reinterpret_cast<uint32_t(&)[4]>(hi[0]));
                    tmem_st4(tA1lo + lane_off + c_lo / 2, reinterpret_cast<uint32_t(&)[4]>(lo[0]));
                } else {
                    tmem_st8(tA1hi + lane_off, hi);
                    tmem_st8(tA1lo + lane_off, lo);
                }
            }
#else
            if (stager) {
#pragma unroll
                for (int c0 = 0; c0 < CH; c0 += 8) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) split_tf32(fcur[c0 + k], hi[k], lo[k]);
                    tmem_st8(tA1hi + lane_off + c_lo + c0, hi);
                    tmem_st8(tA1lo + lane_off + c_lo + c0, lo);
                }
            }
#endif
        };
        if (tile < ntiles) { stage_features(); hand_over(); }      // first tile of this context: layer 1 may run

#pragma unroll 1
        for (; tile < ntiles; tile += stride) {
            HP_T(h0);
            HP_T(h1);
            wait_d();
            HP_T(h2);
            epilogue_hidden(tD + lane_off + col_off, tAhi + lane_off + acol_off, tAlo + lane_off + acol_off, b1 + col_off);
            HP_T(g0);
            hand_over();
            HP_T(g1);
            HP_ADD(8, h2, g0); HP_ADD(9, g0, g1);
            // ---- while layer 2's UMMAs run (this warp would only wait): locate the next tile, start loading its features, and fetch
            //      THIS tile's builtup score, so that neither the address arithmetic nor a DRAM round trip sits between two UMMA phases
            bool valid_n;
            const Pos nxt = locate_next(cur, tile + stride, valid_n);
            {
                const long long foffn = valid_n ? feat_offset(nxt) : 0;
#pragma unroll
                for (int k = 0; k < CH; ++k) fcur[k] = (valid_n && stager) ? ((PC_HEAD_EXP & 1) ? (float)(foffn & 7) : __ldg(a.feats + foffn + (long long)(c_lo + k) * a.f_cs)) : 0.f;
            }
            HP_T(g2);
            HP_ADD(10, g1, g2);
            float bu_cur = 1.f;
            if (half == 0 && valid && a.builtup) {
                const Pos& q = cur;
                bu_cur = (PC_HEAD_EXP & 4) ? (float)q.x : __ldg(a.builtup + (SPARSE ? q.b * HW + q.x : q.b * a.bu_bs + (long long)q.y * a.bu_rs + q.x));
            }
            HP_T(h3);
            wait_d();
            HP_T(h4);
            epilogue_hidden(tD + lane_off + col_off, tAhi + lane_off + acol_off, tAlo + lane_off + acol_off, b2 + col_off);
            hand_over();
            // ---- while layer 3's UMMAs run: the NEXT tile's features (loaded during layer 2) go into the A1 columns, which no UMMA in
            //      flight reads; the issuer learns about them only after this tile's accumulator has been read (below)
            const bool more = tile + stride < ntiles;
            if (more) stage_features();
            HP_T(h5);
            wait_d();
            HP_T(h6);
            HP_ADD(0, h0, h1); HP_ADD(1, h1, h2); HP_ADD(2, h2, h3); HP_ADD(3, h3, h4); HP_ADD(4, h4, h5); HP_ADD(5, h5, h6); HP_ADD(7, h6 - 1, h6);
            // ---- output layer on the CUDA cores: o = b4 + sum_n relu(D3[n] + b3[n]) * w4[n]; each thread sums its 32 hidden
            //      units, the upper half hands its partial sum over through shared memory ----
            float o = 0.f;
            {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tD + lane_off + col_off + 16 * q, v);
                    tc_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        o = fmaf(fmaxf(BIAS_MMA ? __uint_as_float(v[k]) : __uint_as_float(v[k]) + b3[col_off + 16 * q + k], 0.f), w4[col_off + 16 * q + k], o);
                }
            }
            // D has been read and the next tile's A1 is staged: layer 1 of the next tile may run NOW, under this tile's combine + stores
            if (more) hand_over(); else tc_fence_before();
            HP_T(f0);
            if (half == 1) part[px] = o;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + c), "r"(HWORK) : "memory");      // the 8 warps of this context only
            HP_T(f1);
            HP_ADD(11, h6, f0); HP_ADD(12, f0, f1);
            if (half == 0) {
                const float s = fmaxf(b4[0] + o + part[px], 0.f);
                float d = 0.f; int bin = -1;
                if (valid) {
                    const Pos& q = cur;
                    const long long ooff = SPARSE ? q.b * HW + q.x : q.b * a.o_bs + (long long)q.y * a.o_rs + q.x;
                    d = a.builtup ? s * bu_cur : s;
                    if (!(PC_HEAD_EXP & 2) || d == 123.456f) {
                    a.dens[ooff] = d;
                    if (SPARSE) { if (a.scale_sel) a.scale_sel[tile * TM + px] = s; }
                    else if (a.scale) a.scale[ooff] = s;
                    }
                    if (a.sums) {
                        const long long ioff = SPARSE ? 0 : q.b * a.id_bs + (long long)q.y * a.id_rs + q.x;
                        if (SPARSE) bin = q.b;
                        else if (a.census_idx) bin = (a.ids == nullptr || __ldg(a.ids + ioff) == __ldg(a.census_idx + q.b)) ? q.b : -1;
                        else if (a.ids) { const int id = __ldg(a.ids + ioff); bin = (id >= 0 && id < a.R) ? id : -1; }
                        else bin = q.b;
                    }
                }
                if (a.sums) bin_add(a.sums, bin, d);
            }
            HP_T(f2);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + c), "r"(HWORK) : "memory");      // part[] may be rewritten by the next tile
            HP_T(f3);
            HP_ADD(13, f1, f2); HP_ADD(14, f2, f3);
            cur = nxt;
            valid = valid_n;
            HP_T(h7);
            HP_ADD(6, h6, h7);
        }
#if PC_HEAD_PROBE
        if (blockIdx.x == 0 && tid == 0) for (int q = 0; q < 16; ++q) g_head_dbg[q] = hp_acc[q];
#endif
    } else if (elect_one()) {
        // =========================== UMMA issuer: strict alternation between the two contexts ===========================
        long long n[NCTX], rounds = 0;
        uint32_t ph[NCTX];
        for (int c = 0; c < NCTX; ++c) {
            const long long first = (long long)NCTX * blockIdx.x + c;
            n[c] = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
            rounds = n[c] > rounds ? n[c] : rounds;
            ph[c] = 0;
        }
#if PC_HEAD_PROBE
        long long hp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
#pragma unroll 1
        for (long long k = 0; k < rounds; ++k) {
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                // not unrolled over the contexts / layers: six inlined copies of issue_layer need more descriptor registers than the
                // kernel's 96 and the issuing thread then reloads spilled operands from local memory between two UMMAs
#pragma unroll 1
                for (int c = 0; c < NCTX; ++c) {
                    if (k >= n[c]) continue;
                    const uint32_t tD = tbase + CTX_COLS * c, tAhi = tD + C_AHI, tAlo = tD + C_ALO;
                    HP_T(i0);
                    HEAD_WAIT(a_ready(c), ph[c] & 1u); ++ph[c];
                    tc_fence_after();
                    HP_T(i1);
                    if (layer == 0) issue_layer<K1>(tD, tD + C_A1HI, tD + C_A1LO, sW + OFF_W1HI, sW + OFF_W1LO, d_ready(c));
                    else if (layer == 1) issue_layer<HN>(tD, tAhi, tAlo, sW + OFF_W2HI, sW + OFF_W2LO, d_ready(c));
                    else issue_layer<HN>(tD, tAhi, tAlo, sW + OFF_W3HI, sW + OFF_W3LO, d_ready(c));
                    HP_T(i2);
                    HP_ADD(layer == 0 ? 0 : 2, i0, i1); HP_ADD(layer == 0 ? 1 : 3, i1, i2); HP_ADD(7, i2 - 1, i2);
                }
            }
        }
#if PC_HEAD_PROBE
            if (blockIdx.x == 0) for (int q = 0; q < 8; ++q) g_head_dbg[16 + q] = hp_acc[q];
#endif
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8 * NCTX) tmem_dealloc(tbase, 512);
}

template <int K1, bool SPARSE, bool SMALL>
static int launch_head_tc_impl(const HeadArgs& a, long long total_bound, cudaStream_t st) {
    auto k = head_tc_kernel<K1, SPARSE, SMALL>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
    long long tiles = (total_bound + TM - 1) / TM;
    const int maxg = num_sms();              // persistent: one CTA per SM (it owns all 512 TMEM columns), two tile contexts each
    int grid = (int)((tiles + NCTX - 1) / NCTX < maxg ? (tiles + NCTX - 1) / NCTX : maxg);
    if (grid < 1) grid = 1;
    {
        static const int cat = prof_register(SPARSE ? "head_tc<sparse>" : "head_tc<dense>");
        ProfScope prof(cat, st, (double)total_bound);
        k<<<grid, HTHREADS, TC2_SMEM_BYTES, st>>>(a);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

template <int K1, bool SPARSE>
static int launch_head_tc(const HeadArgs& a, long long total_bound, cudaStream_t st) {
    const long long HW = SPARSE ? a.HW : (long long)a.H * a.W;
    const bool small = total_bound < 0x7fffffffll && HW * (SPARSE ? 1 : a.B) < 0x7fffffffll && HW < 0x7fffffffll;
    return small ? launch_head_tc_impl<K1, SPARSE, true>(a, total_bound, st) : launch_head_tc_impl<K1, SPARSE, false>(a, total_bound, st);
}

}  // namespace pc

using namespace pc;

extern "C" int pc_head_tc_pack_bytes(void) { return TC_PACK_BYTES; }
extern "C" int pc_tc_operand_format(void) { return PC_TC_F16; }
#if PC_HEAD_PROBE
extern "C" int pc_debug_head_counters(long long* out32) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, pc::g_head_dbg, sizeof(long long) * 32);
    return 0;
}
#endif

extern "C" int pc_head_dense_forward_tc(const void* tcpack, int head_in, const float* feats, long long f_bstride,
                                        long long f_cstride, int f_rstride, const float* builtup, long long bu_bstride,
                                        int bu_rstride, int B, int H, int W, float* dens, float* scale,
                                        long long o_bstride, int o_rstride, const int32_t* ids, long long id_bstride,
                                        int id_rstride, const int32_t* census_idx, double* sums, int R,
                                        pc_stream_t stream) {
    PC_CHECK_ARG(tcpack && feats && dens, "null pointer");
    PC_CHECK_ARG((((uintptr_t)tcpack) & 15) == 0, "tc pack must be 16-byte aligned");
    PC_CHECK_ARG(head_in == 16 || head_in == 8, "head_in must be 16 or 8");
    PC_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "bad shape");
    PC_CHECK_ARG(!(ids && !census_idx && sums) || R >= 1, "R must be >= 1 with an id raster");
    HeadArgs a{};
    a.pack = reinterpret_cast<const float*>(tcpack); a.feats = feats; a.f_bs = f_bstride; a.f_cs = f_cstride; a.f_rs = f_rstride;
    a.builtup = builtup; a.bu_bs = bu_bstride; a.bu_rs = bu_rstride; a.B = B; a.H = H; a.W = W;
    a.dens = dens; a.scale = scale; a.o_bs = o_bstride; a.o_rs = o_rstride;
    a.ids = ids; a.id_bs = id_bstride; a.id_rs = id_rstride; a.census_idx = census_idx; a.sums = sums; a.R = R;
    const long long total = (long long)B * H * W;
    return head_in == 16 ? launch_head_tc<16, false>(a, total, (cudaStream_t)stream)
                         : launch_head_tc<8, false>(a, total, (cudaStream_t)stream);
}

extern "C" int pc_head_sparse_forward_tc(const void* tcpack, int head_in, const float* feats, long long f_bstride,
                                         long long f_cstride, const float* builtup, const int32_t* idx,
                                         const int32_t* n_dev, long long n_max, long long HW, float* dens,
                                         float* scale_sel, double* popcount, pc_stream_t stream) {
    PC_CHECK_ARG(tcpack && feats && idx && n_dev && dens, "null pointer");
    PC_CHECK_ARG((((uintptr_t)tcpack) & 15) == 0, "tc pack must be 16-byte aligned");
    PC_CHECK_ARG(head_in == 16 || head_in == 8, "head_in must be 16 or 8");
    PC_CHECK_ARG(HW >= 1 && n_max >= 0, "bad shape");
    HeadArgs a{};
    a.pack = reinterpret_cast<const float*>(tcpack); a.feats = feats; a.f_bs = f_bstride; a.f_cs = f_cstride; a.builtup = builtup;
    a.dens = dens; a.scale_sel = scale_sel; a.sums = popcount; a.idx = idx; a.n_dev = n_dev; a.HW = HW;
    a.B = 1; a.H = 1; a.W = 1;
    return head_in == 16 ? launch_head_tc<16, true>(a, n_max, (cudaStream_t)stream)
                         : launch_head_tc<8, true>(a, n_max, (cudaStream_t)stream);
}
