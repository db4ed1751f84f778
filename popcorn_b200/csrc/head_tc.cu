// Occupancy head on the 5th-gen tensor cores (tcgen05 + TMEM), error-compensated 3xTF32.
//
// Per 128-pixel tile (UMMA M = 128, N = 64, cta_group::1), 256 threads:
//   * a pixel = one TMEM lane; TWO warps share each 32-lane quarter (warp w and w+4) and split the per-pixel work by
//     columns (features 0-7 | 8-15, hidden units 0-31 | 32-63), which halves every epilogue between two layers' UMMAs;
//   * activations are the A operand and LIVE IN TMEM: the epilogue reads the fp32 accumulator D with
//     tcgen05.ld, applies bias + ReLU, splits x = hi + lo (hi = top 19 bits, exactly a TF32 number) and
//     writes both halves back with tcgen05.st — they never touch shared or global memory;
//   * weights are the B operand in shared memory, pre-split (hi/lo) and pre-swizzled on the host into the
//     canonical K-major SWIZZLE_128B layout, copied in verbatim once per CTA;
//   * every 8-wide k-step issues three tcgen05.mma.kind::tf32 (hi*hi + lo*hi + hi*lo) into the same fp32
//     accumulator: ~21-bit operands, which is what the per-pixel 1e-2 bar needs;
//   * the 64->1 output layer, ReLU, x builtup, stores and the census partial sums run in the last epilogue.
// Two CTAs per SM (256 TMEM columns, 85 KB smem each) overlap one tile's epilogue with the other's MMAs.
// Replaces model/popcorn.py:79-88, 160-190, 195-228 (same contract as head.cu's SIMT kernel).
#include "head_common.cuh"
#include "tc_common.cuh"

namespace pc {

constexpr int TM = 128;          // pixels per tile
constexpr int HT = 256;          // threads per CTA: two warps per TMEM lane quarter
constexpr uint32_t IDESC = umma_idesc_tf32(128, 64);

// byte offsets inside the packed TC weight image (host: weights.pack_head_tc)
constexpr int OFF_W1HI = 0, OFF_W1LO = 8192, OFF_W2HI = 16384, OFF_W2LO = 32768, OFF_W3HI = 49152, OFF_W3LO = 65536;
constexpr int OFF_VEC = 81920;                 // b1[64] b2[64] b3[64] w4[64] b4[4]
constexpr int TC_PACK_BYTES = OFF_VEC + 260 * 4;
constexpr int OFF_MBAR = TC_PACK_BYTES;        // 8-byte mbarrier
constexpr int OFF_TMEM = OFF_MBAR + 8;         // 4-byte TMEM base address slot
constexpr int OFF_PART = OFF_TMEM + 8;         // float[128]: output-layer partial dot of the upper-half warps
constexpr int TC_SMEM_BYTES = OFF_PART + 512 + 1024;   // + slack to align the base to 1024 B

// one hidden layer's UMMAs: D[128x64] = A[128xK] * W[64xK]^T, K in steps of 8, three split terms per step
template <int K>
__device__ __forceinline__ void issue_layer(uint32_t tD, uint32_t tAhi, uint32_t tAlo, uint32_t sWhi, uint32_t sWlo,
                                            uint32_t mbar) {
#pragma unroll
    for (int j = 0; j < K / 8; ++j) {
        const uint32_t koff = (uint32_t)((j >> 2) * 8192 + (j & 3) * 32);   // 32-float swizzle atoms along K
        const uint64_t bhi = make_bdesc(sWhi + koff), blo = make_bdesc(sWlo + koff);
        umma_tf32_ts(tD, tAhi + 8 * j, bhi, IDESC, j > 0 ? 1u : 0u);
        umma_tf32_ts(tD, tAlo + 8 * j, bhi, IDESC, 1u);
        umma_tf32_ts(tD, tAhi + 8 * j, blo, IDESC, 1u);
    }
    umma_commit(mbar);
}

// hidden-layer epilogue of one thread: its 32 columns of D -> relu(D + bias) -> (hi, lo) -> A operand of the next layer
__device__ __forceinline__ void epilogue_hidden(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float* bias) {
    uint32_t v[2][16];
#pragma unroll
    for (int q = 0; q < 2; ++q) tmem_ld16(tD + 16 * q, v[q]);      // both loads in flight, one wait
    tc_wait_ld();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_tf32(fmaxf(__uint_as_float(v[q][i]) + bias[16 * q + i], 0.f), hi[i], lo[i]);
        tmem_st16(tAhi + 16 * q, hi);
        tmem_st16(tAlo + 16 * q, lo);
    }
}

template <int K1, bool SPARSE>
__global__ void __launch_bounds__(HT, 2) head_tc_kernel(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const float* vec = reinterpret_cast<const float*>(sm + OFF_VEC);
    const float* b1 = vec, *b2 = vec + 64, *b3 = vec + 128, *w4 = vec + 192, *b4 = vec + 256;
    const uint32_t mbar = smem_u32(sm + OFF_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
    float* part = reinterpret_cast<float*>(sm + OFF_PART);
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const int px = tid & (TM - 1);                   // pixel (= TMEM lane) of this thread inside a tile
    const int half = warp >> 2;                      // which half of the columns this thread works on
    constexpr int CH = K1 >= 16 ? K1 / 2 : K1;       // layer-1 feature channels per thread (K1 = 8: lower half stages all)
    const bool stager = K1 >= 16 || half == 0;
    const int c_lo = K1 >= 16 ? half * CH : 0;

    for (int i = tid; i < TC_PACK_BYTES / 16; i += HT)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(a.pack) + i);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
    if (tid == 0) mbar_init1(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weight image (generic stores) -> visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;    // a warp may touch TMEM lanes 32*(warp%4) .. +31
    const uint32_t tD = tbase, tAhi = tbase + 64, tAlo = tbase + 128;  // column offsets inside the 256-col block
    const uint32_t col_off = (uint32_t)(32 * half);                  // this thread's 32 hidden units
    const uint32_t sW = smem_u32(sm);
    uint32_t phase = 0;

    const long long HW = SPARSE ? a.HW : (long long)a.H * a.W;
    const long long total = SPARSE ? (long long)__ldg(a.n_dev) : HW * a.B;

    // pixel bookkeeping + feature fetch of one tile; the NEXT tile's features are fetched while the current tile
    // runs its MMAs / epilogues, so the global-load latency is off the critical path
    struct Px { long long i; bool valid; int b; long long boff, ooff, ioff; };
    auto locate = [&](long long base, long long& foff) {
        Px q;
        q.i = base + px;
        q.valid = q.i < total;
        q.b = 0; q.boff = 0; q.ooff = 0; q.ioff = 0; foff = 0;
        if (q.valid) {
            const long long p = SPARSE ? (long long)__ldg(a.idx + q.i) : q.i;
            q.b = (int)(p / HW);
            const long long r = p - (long long)q.b * HW;
            if (SPARSE) {
                foff = q.b * a.f_bs + r; q.boff = p; q.ooff = p;
            } else {
                const int y = (int)(r / a.W), x = (int)(r - (long long)y * a.W);
                foff = q.b * a.f_bs + (long long)y * a.f_rs + x;
                q.boff = q.b * a.bu_bs + (long long)y * a.bu_rs + x;
                q.ooff = q.b * a.o_bs + (long long)y * a.o_rs + x;
                q.ioff = q.b * a.id_bs + (long long)y * a.id_rs + x;
            }
        }
        return q;
    };
    float fcur[CH], fnext[CH];
    long long foff0;
    Px cur = locate((long long)blockIdx.x * TM, foff0);
#pragma unroll
    for (int c = 0; c < CH; ++c) fcur[c] = (cur.valid && stager) ? __ldg(a.feats + foff0 + (long long)(c_lo + c) * a.f_cs) : 0.f;

    for (long long base = (long long)blockIdx.x * TM; base < total; base += (long long)gridDim.x * TM) {
        const long long i = cur.i;
        const bool valid = cur.valid;
        const int b = cur.b;
        const long long boff = cur.boff, ooff = cur.ooff, ioff = cur.ioff;
        // ---- layer-1 A operand: this pixel's features (this thread's channel half), split, into TMEM ----
        if (stager) {
#pragma unroll
            for (int c0 = 0; c0 < CH; c0 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) split_tf32(fcur[c0 + c], hi[c], lo[c]);
                tmem_st8(tAhi + lane_off + c_lo + c0, hi);
                tmem_st8(tAlo + lane_off + c_lo + c0, lo);
            }
        }
        // ---- prefetch the next tile's features (consumed at the top of the next iteration) ----
        long long foffn;
        const Px nxt = locate(base + (long long)gridDim.x * TM, foffn);
#pragma unroll
        for (int c = 0; c < CH; ++c) fnext[c] = (nxt.valid && stager) ? __ldg(a.feats + foffn + (long long)(c_lo + c) * a.f_cs) : 0.f;
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer<K1>(tD, tAhi, tAlo, sW + OFF_W1HI, sW + OFF_W1LO, mbar); }
        mbar_wait_sleep(mbar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_hidden(tD + lane_off + col_off, tAhi + lane_off + col_off, tAlo + lane_off + col_off, b1 + col_off);
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer<HN>(tD, tAhi, tAlo, sW + OFF_W2HI, sW + OFF_W2LO, mbar); }
        mbar_wait_sleep(mbar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_hidden(tD + lane_off + col_off, tAhi + lane_off + col_off, tAlo + lane_off + col_off, b2 + col_off);
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer<HN>(tD, tAhi, tAlo, sW + OFF_W3HI, sW + OFF_W3LO, mbar); }
        mbar_wait_sleep(mbar, phase); phase ^= 1;
        tc_fence_after();
        // ---- output layer on the CUDA cores: o = b4 + sum_n relu(D3[n] + b3[n]) * w4[n]; each thread sums its 32 hidden
        //      units, the upper half hands its partial sum over through shared memory ----
        float o = 0.f;
        {
            uint32_t v[2][16];
            tmem_ld16(tD + lane_off + col_off, v[0]);
            tmem_ld16(tD + lane_off + col_off + 16, v[1]);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    o = fmaf(fmaxf(__uint_as_float(v[q][k]) + b3[col_off + 16 * q + k], 0.f), w4[col_off + 16 * q + k], o);
        }
        tc_fence_before();   // D is overwritten by the next tile's first UMMA after the next __syncthreads
        if (half == 1) part[px] = o;
        __syncthreads();
        if (half == 0) {
            const float s = fmaxf(b4[0] + o + part[px], 0.f);
            float d = 0.f; int bin = -1;
            if (valid) {
                d = a.builtup ? s * __ldg(a.builtup + boff) : s;
                a.dens[ooff] = d;
                if (SPARSE) { if (a.scale_sel) a.scale_sel[i] = s; }
                else if (a.scale) a.scale[ooff] = s;
                if (a.sums) {
                    if (SPARSE) bin = b;
                    else if (a.census_idx) bin = (a.ids == nullptr || __ldg(a.ids + ioff) == __ldg(a.census_idx + b)) ? b : -1;
                    else if (a.ids) { const int id = __ldg(a.ids + ioff); bin = (id >= 0 && id < a.R) ? id : -1; }
                    else bin = b;
                }
            }
            if (a.sums) bin_add(a.sums, bin, d);
        }
        cur = nxt;
#pragma unroll
        for (int c = 0; c < CH; ++c) fcur[c] = fnext[c];
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

template <int K1, bool SPARSE>
static int launch_head_tc(const HeadArgs& a, long long total_bound, cudaStream_t st) {
    auto k = head_tc_kernel<K1, SPARSE>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    long long tiles = (total_bound + TM - 1) / TM;
    const int maxg = num_sms() * 2;          // persistent: 2 CTAs per SM, each walks tiles grid-stride
    int grid = (int)(tiles < maxg ? tiles : maxg);
    if (grid < 1) grid = 1;
    {
        static const int cat = prof_register(SPARSE ? "head_tc<sparse>" : "head_tc<dense>");
        ProfScope prof(cat, st, (double)total_bound);
        k<<<grid, HT, TC_SMEM_BYTES, st>>>(a);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_head_tc_pack_bytes(void) { return TC_PACK_BYTES; }

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
