// 3x3 convolutions of the DDA UNet as implicit GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), 3xTF32.
//
// Mapping (per CTA: a tile of 128 columns x TR rows of one (image, stream) job, 128 threads):
//   * UMMA M = 128 = the 128 pixels of one image row segment; thread t owns pixel x0+t == TMEM lane t;
//   * K of one input row = (kx, ci): the row is written to TMEM THREE times, shifted by -1/0/+1 pixel (the thread
//     loads its own pixel coalesced, gets the neighbours by warp shuffle; lanes 0/31 load the halo pixel), each
//     value split x = hi + lo (hi = top 19 bits = exact TF32) -> A operand [128 x 3*Cin] hi and lo, in TMEM;
//   * the ky shift is NOT a lane shift: input row r feeds output rows r+1 (ky=0), r (ky=1), r-1 (ky=2), i.e. three
//     different fp32 accumulators D[128 x 16] that live in a 4-slot TMEM ring (slot = output row & 3);
//   * B operand = the folded weights W_ky[co][(kx,ci)] (N = 16: Cout 16, or Cout 8 zero-padded), pre-split and
//     pre-swizzled on the host (K-major SWIZZLE_128B), copied to shared memory once per CTA;
//   * per input row, one thread issues 3 (ky) x 3*Cin/8 (k-steps) x 3 (hi*hi + lo*hi + hi*lo) tcgen05.mma
//     kind::tf32 and commits them to an mbarrier; while they run, the CTA prefetches the next input row into
//     registers and runs the epilogue of output row r-2 (tcgen05.ld -> bias + ReLU -> store | 2x2 maxpool |
//     1x1 logit dot + sigmoid).  Several CTAs per SM (TMEM 128/256 columns each) keep the tensor pipe busy.
// Precision: ~21-bit operands, fp32 accumulation — what the 1e-2 per-pixel bar needs; plain
// single-pass TF32 fails it.  The CUDA cores only move data: Cin loads + 2*Cin shuffles + 3*Cin splits per pixel
// instead of 9*Cin*Cout FMAs.
//
// Replaces model/DDA_model/utils/networks.py:253-271 (DoubleConv: Conv2d 3x3 pad 1 + BatchNorm2d(eval) + ReLU),
// :284-295 (MaxPool2d in Down), :318 (skip concat), :323-330 (OutConv) and popcorn.py:244,296-300,317-320.
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace pc {

constexpr int TCM = 128;           // pixels per UMMA = threads per CTA
constexpr int TCN = 16;            // UMMA N (output channels, zero-padded)
constexpr int DSLOTS = 4;          // accumulator ring

template <int CIN>
struct TcGeom {
    static constexpr int KROW = (3 * CIN + 7) / 8 * 8;          // A columns per half (hi | lo) of one input row
    static constexpr int KSTEPS = KROW / 8;
    static constexpr int KATOMS = (KROW + 31) / 32;             // 32-float swizzle atoms along K
    static constexpr int BMAT = KATOMS * TCN * 128;             // bytes of one swizzled [16 x KROW] matrix
    static constexpr int OFF_BIAS = 6 * BMAT;                   // matrices: [ky][hi, lo]
    static constexpr int IMG_BYTES = OFF_BIAS + 64;             // + bias[16]
    static constexpr int A_COLS = 2 * KROW;                     // hi at [0, KROW), lo at [KROW, 2*KROW)
    static constexpr int D_COL0 = A_COLS;                       // 4 accumulator slots of 16 columns
    static constexpr int TMEM_COLS = (A_COLS + DSLOTS * TCN) <= 128 ? 128 : 256;
    static constexpr int CTAS = 512 / TMEM_COLS;                // resident CTAs per SM (TMEM is the limit)
    static constexpr int OFF_MBAR = IMG_BYTES;
    static constexpr int OFF_TMEM = OFF_MBAR + 8;
    // shared-memory request: padded so that no more than CTAS CTAs fit on an SM (a CTA beyond the TMEM capacity
    // would block in tcgen05.alloc while holding an SM slot)
    static constexpr int SMEM_MIN = 227 * 1024 / (CTAS + 1) + 1024;
    static constexpr int SMEM_BYTES = (OFF_TMEM + 8 + 1024) > SMEM_MIN ? (OFF_TMEM + 8 + 1024) : SMEM_MIN;
    static_assert(A_COLS + DSLOTS * TCN <= 256, "TMEM budget");
    static_assert(A_COLS % 16 == 0, "accumulator slots must start on a 16-column boundary");
};

template <int CIN_A, int CIN_B, int COUT, int EPI>
__global__ void __launch_bounds__(TCM, TcGeom<CIN_A + CIN_B>::CTAS)
conv3x3_tc_kernel(const __grid_constant__ TcConvParams p) {
    constexpr int CIN = CIN_A + CIN_B;
    using G = TcGeom<CIN>;
    constexpr uint32_t IDESC = umma_idesc_tf32(TCM, TCN);
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const ConvJob& job = p.jobs[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const float* bias = reinterpret_cast<const float*>(sm + G::OFF_BIAS);
    const uint32_t mbar = smem_u32(sm + G::OFF_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + G::OFF_TMEM);

    for (int i = tid; i < G::IMG_BYTES / 16; i += TCM)
        reinterpret_cast<int4*>(sm)[i] = __ldg(reinterpret_cast<const int4*>(job.wtc) + i);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), G::TMEM_COLS);
    if (tid == 0) mbar_init1(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weight image (generic stores) -> visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;          // this warp's 32 TMEM lanes
    const uint32_t tAhi = tbase, tAlo = tbase + G::KROW, tD = tbase + G::D_COL0;
    const uint32_t sW = smem_u32(sm);
    uint32_t phase = 0;

    const int H = p.H, W = p.W;
    float dotw[EPI == EPI_DOT ? 8 : 1];
    float dotb = 0.f;
    if (EPI == EPI_DOT) {
#pragma unroll
        for (int o = 0; o < 8; ++o) dotw[o] = __ldg(job.dotw + o);
        dotb = __ldg(job.dotw + 8);
    }

    const int ntiles = p.tiles_x * p.tiles_y;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int x0 = tx * TCM, y0 = ty * p.TR;
        const int nrows = (H - y0) < p.TR ? (H - y0) : p.TR;
        const int vx = x0 + tid;
        // ---- column bookkeeping of this thread: own pixel and (lanes 0 / 31) the halo pixel of its warp ----
        const bool is_edge = (lane == 0) || (lane == 31);
        const int vxe = (lane == 0) ? vx - 1 : vx + 1;
        const bool in_x = vx < W, in_xe = is_edge && vxe >= 0 && vxe < W;
        int a_sx = vx - job.a_ox, a_sxe = vxe - job.a_ox;
        bool a_ok = in_x, a_oke = in_xe;
        if (job.a_reflect) {
            a_sx = a_sx < 0 ? -a_sx : a_sx;     a_sx = a_sx >= job.a_W ? 2 * (job.a_W - 1) - a_sx : a_sx;
            a_sxe = a_sxe < 0 ? -a_sxe : a_sxe; a_sxe = a_sxe >= job.a_W ? 2 * (job.a_W - 1) - a_sxe : a_sxe;
        } else {
            a_ok = a_ok && a_sx >= 0 && a_sx < job.a_W;
            a_oke = a_oke && a_sxe >= 0 && a_sxe < job.a_W;
        }
        const int b_sx = vx - job.b_ox, b_sxe = vxe - job.b_ox;
        const bool b_ok = CIN_B > 0 && in_x && b_sx >= 0 && b_sx < job.b_W;
        const bool b_oke = CIN_B > 0 && in_xe && b_sxe >= 0 && b_sxe < job.b_W;

        float v[CIN], e[CIN];     // input row in flight: own pixel / halo pixel (lanes 0, 31)
        auto load_row = [&](int r) {
            const int vy = y0 + r;
            const bool rin = vy >= 0 && vy < H;
            {
                int sy = vy - job.a_oy;
                bool rok = rin;
                if (job.a_reflect) { sy = sy < 0 ? -sy : sy; sy = sy >= job.a_H ? 2 * (job.a_H - 1) - sy : sy; }
                else rok = rok && sy >= 0 && sy < job.a_H;
                const float* rowp = job.a + (rok ? (long long)sy * job.a_rs : 0ll);
                const bool ok = rok && a_ok, oke = rok && a_oke;
#pragma unroll
                for (int c = 0; c < CIN_A; ++c) {
                    int plane = c;
                    if (CIN_A <= 4) plane = (job.a_chmap >> (8 * c)) & 0xff;
                    const float* pp = rowp + plane * job.a_cs;
                    v[c] = 0.f; e[c] = 0.f;
                    if (ok) v[c] = __ldg(pp + a_sx);
                    if (oke) e[c] = __ldg(pp + a_sxe);
                }
            }
            if (CIN_B > 0) {
                const int sy = vy - job.b_oy;
                const bool rok = rin && sy >= 0 && sy < job.b_H;
                const float* rowp = job.b + (rok ? (long long)sy * job.b_rs : 0ll);
                const bool ok = rok && b_ok, oke = rok && b_oke;
#pragma unroll
                for (int c = 0; c < CIN_B; ++c) {
                    const float* pp = rowp + c * job.b_cs;
                    v[CIN_A + c] = 0.f; e[CIN_A + c] = 0.f;
                    if (ok) v[CIN_A + c] = __ldg(pp + b_sx);
                    if (oke) e[CIN_A + c] = __ldg(pp + b_sxe);
                }
            }
        };
        // registers -> TMEM: A[lane][kx*CIN + ci] = in[ci][x + kx - 1], split into hi / lo
        auto stage_row = [&]() {
#pragma unroll
            for (int j = 0; j < G::KSTEPS; ++j) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int col = 8 * j + q;
                    float val = 0.f;
                    if (col < 3 * CIN) {
                        const int kx = col / CIN, ci = col % CIN;
                        if (kx == 1) {
                            val = v[ci];
                        } else if (kx == 0) {
                            const float t = __shfl_up_sync(FULL, v[ci], 1);
                            val = lane == 0 ? e[ci] : t;
                        } else {
                            const float t = __shfl_down_sync(FULL, v[ci], 1);
                            val = lane == 31 ? e[ci] : t;
                        }
                    }
                    split_tf32(val, hi[q], lo[q]);
                }
                tmem_st8(tAhi + lane_off + 8 * j, hi);
                tmem_st8(tAlo + lane_off + 8 * j, lo);
            }
        };
        // all UMMAs of input row r (tile-local, -1 .. nrows), issued by one thread
        auto issue_row = [&](int r) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int y = r - ky + 1;
                if (y < 0 || y >= nrows) continue;
                const uint32_t d = tD + TCN * (uint32_t)(y & (DSLOTS - 1));
                const uint32_t whi = sW + (2 * ky) * G::BMAT, wlo = sW + (2 * ky + 1) * G::BMAT;
#pragma unroll
                for (int j = 0; j < G::KSTEPS; ++j) {
                    const uint32_t koff = (uint32_t)((j >> 2) * (TCN * 128) + (j & 3) * 32);
                    const uint64_t bhi = make_bdesc(whi + koff), blo = make_bdesc(wlo + koff);
                    umma_tf32_ts(d, tAhi + 8 * j, bhi, IDESC, (ky == 0 && j == 0) ? 0u : 1u);   // ky = 0 opens the row
                    umma_tf32_ts(d, tAlo + 8 * j, bhi, IDESC, 1u);
                    umma_tf32_ts(d, tAhi + 8 * j, blo, IDESC, 1u);
                }
            }
            umma_commit(mbar);
        };
        float prev[EPI == EPI_POOL ? COUT : 1];   // horizontally pooled even row, waiting for the odd row
        auto epilogue = [&](int y) {
            uint32_t d[COUT];
            if (COUT == 16) tmem_ld16(tD + lane_off + TCN * (uint32_t)(y & (DSLOTS - 1)), reinterpret_cast<uint32_t(&)[16]>(d));
            else tmem_ld8(tD + lane_off + TCN * (uint32_t)(y & (DSLOTS - 1)), reinterpret_cast<uint32_t(&)[8]>(d));
            tc_wait_ld();
            float acc[COUT];
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[o] = fmaxf(__uint_as_float(d[o]) + bias[o], 0.f);
            const int oy = y0 + y;
            const int yy = oy - p.crop_y, xx = vx - p.crop_x;
            const bool inside = in_x && yy >= 0 && yy < p.crop_H && xx >= 0 && xx < p.crop_W;
            if (EPI == EPI_DOT) {
                if (inside) {
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < 8; ++o) s = fmaf(acc[o], dotw[o], s);
                    if (job.dot_in) s += job.dot_in[(long long)yy * job.dot_in_rs + xx];
                    if (job.dot_final) {
                        s += dotb;
                        s = 1.f / (1.f + expf(-s));
                    }
                    job.dot_out[(long long)yy * job.dot_out_rs + xx] = s;
                }
                return;
            }
            if (job.out && inside) {
                float* dst = job.out + (long long)yy * job.out_rs + xx;
#pragma unroll
                for (int o = 0; o < COUT; ++o) dst[(long long)o * job.out_cs] = acc[o];
            }
            if (EPI == EPI_POOL) {
                const int py = oy >> 1, px = vx >> 1;
                const bool st = (y & 1) && !(lane & 1) && py < (H >> 1) && px < (W >> 1);
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const float hm = fmaxf(acc[o], __shfl_xor_sync(FULL, acc[o], 1));
                    if (y & 1) {
                        if (st) job.pool[(long long)o * job.pool_cs + (long long)py * job.pool_rs + px] = fmaxf(prev[o], hm);
                    } else {
                        prev[o] = hm;
                    }
                }
            }
        };

        // ---- software pipeline over the input rows of the tile ----
        load_row(-1);
#pragma unroll 1
        for (int r = -1; r <= nrows; ++r) {
            if (r > -1) {                       // UMMAs of row r-1 done: the A buffer is free, output row r-2 is final
                mbar_wait(mbar, phase); phase ^= 1;
                tc_fence_after();
            }
            stage_row();
            tc_wait_st();
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) { tc_fence_after(); issue_row(r); }
            if (r < nrows) load_row(r + 1);     // in flight while the UMMAs and the epilogue below run
            if (r >= 2) epilogue(r - 2);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        epilogue(nrows - 1);
        tc_fence_before();                      // the next tile's first UMMA follows its first __syncthreads
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, G::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int tc_geom_img_floats(int cin) {
    const int krow = (3 * cin + 7) / 8 * 8, katoms = (krow + 31) / 32;
    return (6 * katoms * TCN * 128 + 64) / 4;
}

int conv_tc_layer_floats(int cin) { return (int)round_up(tc_geom_img_floats(cin), 64); }   // 256-B multiple

// flat = [cin][ky][kx][cout] + bias[cout] (the SIMT pack)  ->  [ky][hi|lo][katom][16 rows][32 floats] swizzled + bias[16]
void conv_tc_pack_layer(const float* flat, int cin, int cout, float* img) {
    const int krow = (3 * cin + 7) / 8 * 8, katoms = (krow + 31) / 32;
    const int mat = katoms * TCN * 32;                  // floats per matrix
    const int total = conv_tc_layer_floats(cin);
    memset(img, 0, sizeof(float) * total);
    for (int ky = 0; ky < 3; ++ky)
        for (int n = 0; n < cout; ++n)
            for (int kx = 0; kx < 3; ++kx)
                for (int ci = 0; ci < cin; ++ci) {
                    const float w = flat[((ci * 3 + ky) * 3 + kx) * cout + n];
                    uint32_t bits;
                    memcpy(&bits, &w, 4);
                    bits &= 0xFFFFE000u;
                    float hi;
                    memcpy(&hi, &bits, 4);
                    const float lo = w - hi;
                    const int k = kx * cin + ci;
                    const int atom = k / 32, kk = k % 32;
                    const int pos = (((kk / 4) ^ (n % 8)) * 4) + kk % 4;      // Swizzle<3,4,3>: 16-B chunk ^= row % 8
                    const int idx = atom * (TCN * 32) + n * 32 + pos;
                    img[(2 * ky) * mat + idx] = hi;
                    img[(2 * ky + 1) * mat + idx] = lo;
                }
    for (int n = 0; n < cout; ++n) img[6 * mat + n] = flat[cin * 9 * cout + n];
}

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
        int v = e ? atoi(e) : 32;
        if (v < 2) v = 2;
        return v & ~1;                                  // even: the 2x2 pool pairs rows inside a tile
    }();
    return tr;
}

template <int CIN_A, int CIN_B, int COUT, int EPI>
static int launch_tc_impl(TcConvParams& p, int njobs, cudaStream_t st) {
    using G = TcGeom<CIN_A + CIN_B>;
    static const int cat = [] {
        char nm[64];
        snprintf(nm, sizeof(nm), "conv3x3_tc<%d,%d,%d,%s>", CIN_A, CIN_B, COUT, EPI == EPI_STORE ? "store" : EPI == EPI_POOL ? "pool" : "dot");
        return prof_register(nm);
    }();
    auto k = conv3x3_tc_kernel<CIN_A, CIN_B, COUT, EPI>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    p.TR = conv_tc_rows();
    p.tiles_x = cdiv(p.W, TCM);
    p.tiles_y = cdiv(p.H, p.TR);
    const int ntiles = p.tiles_x * p.tiles_y;
    int per_job = cdiv((long long)num_sms() * G::CTAS, njobs);       // persistent CTAs of one job
    if (per_job > ntiles) per_job = ntiles;
    if (per_job < 1) per_job = 1;
    {
        ProfScope prof(cat, st, (double)p.H * p.W * njobs);
        k<<<dim3(per_job, njobs), TCM, G::SMEM_BYTES, st>>>(p);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

int launch_conv_tc(int cin_a, int cin_b, int cout, int epi, TcConvParams& p, int njobs, cudaStream_t st) {
    const int key = ((cin_a * 100 + cin_b) * 100 + cout) * 10 + epi;
    switch (key) {
        case ((2 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<2, 0, 8, EPI_STORE>(p, njobs, st);
        case ((4 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<4, 0, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<8, 0, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_POOL: return launch_tc_impl<8, 0, 8, EPI_POOL>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 8) * 10 + EPI_DOT: return launch_tc_impl<8, 0, 8, EPI_DOT>(p, njobs, st);
        case ((8 * 100 + 0) * 100 + 16) * 10 + EPI_STORE: return launch_tc_impl<8, 0, 16, EPI_STORE>(p, njobs, st);
        case ((16 * 100 + 0) * 100 + 16) * 10 + EPI_STORE: return launch_tc_impl<16, 0, 16, EPI_STORE>(p, njobs, st);
        case ((16 * 100 + 0) * 100 + 16) * 10 + EPI_POOL: return launch_tc_impl<16, 0, 16, EPI_POOL>(p, njobs, st);
        case ((16 * 100 + 16) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<16, 16, 8, EPI_STORE>(p, njobs, st);
        case ((8 * 100 + 8) * 100 + 8) * 10 + EPI_STORE: return launch_tc_impl<8, 8, 8, EPI_STORE>(p, njobs, st);
    }
    set_error("launch_conv_tc: no instantiation for (%d,%d,%d,%d)", cin_a, cin_b, cout, epi);
    return PC_ERR_INVALID;
}

}  // namespace pc
