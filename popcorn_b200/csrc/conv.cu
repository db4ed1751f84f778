// DDA dual-stream UNet backbone on sm_100a: fused 3x3 conv (+folded BN +ReLU [+2x2 maxpool | +1x1
// logit dot +sigmoid]) stencil kernels, 2x2 transposed conv, and the host-side layer schedule.
//
// Why SIMT fp32 and not tcgen05 here: the layers have N = 8/16 output channels
// and K = 9*Cin; single-pass TF32/BF16 operands fail the 1e-2 per-pixel bar, and with
// 3xTF32 splitting an M=128,N=8 UMMA is bound by re-reading its A operand from shared memory (4 KB per
// 8192 MAC), i.e. below the FP32 FMA rate.  A register-tiled stencil re-uses every staged input value
// for 9 taps x 8 channels out of registers instead.
//
// Replaces model/DDA_model/utils/networks.py:121-151 (UNet.forward), :253-330 (DoubleConv/Down/Up/
// OutConv) and the padding / reorder / sigmoid / crop wrappers of model/popcorn.py:126-158, 279-322.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"

namespace pc {

constexpr int TILE = 32;            // output tile edge per CTA
constexpr int SROWS = TILE + 2;     // staged rows (1-px halo)
constexpr int SPITCH = 40;          // staged row pitch (floats) = TMA box width (inner TMA coordinate must be 16-B aligned)
constexpr int XOFF = 3;             // staged column of image column x0-1: the box starts at x0-4

struct alignas(64) ConvParams {
    CUtensorMap tmA[MAX_JOBS];           // TMA descriptors of the job's sources (TMA staging only)
    CUtensorMap tmB[MAX_JOBS];
    int H, W;                            // virtual image == output extent
    int crop_y, crop_x, crop_H, crop_W;  // stores go to (y-crop_y, x-crop_x) if inside [0,crop_H)x[0,crop_W)
    ConvJob jobs[MAX_JOBS];
};

template <int CIN>
__host__ __device__ constexpr int conv_cc() { return CIN < 4 ? CIN : 4; }       // channels per staged chunk

template <int CIN, int COUT>
__host__ __device__ constexpr int conv_wfloats_padded() { return (CIN * 9 * COUT + COUT + 31) / 32 * 32; }   // 128-B multiple

// The Cin = 2 first layer (SAR stream) walks PC_CONV_MT consecutive tiles of a row per CTA and stages tile t+1 while tile t is computed
// (two buffers): the copies' DRAM latency, the weight load and the CTA launch are paid once per run instead of once per tile
// (layer micro-benchmark 0.340 -> 0.308 ms per 33.5 Mpx).  Measured and NOT adopted for Cin = 4 (0.489 -> 0.527 ms: its tile already
// carries twice the FFMA2 work to hide the copies behind).
#ifndef PC_CONV_MT
#define PC_CONV_MT 4
#endif
#ifndef PC_CONV_MT4
#define PC_CONV_MT4 1
#endif
template <int CIN, bool TMA>
__host__ __device__ constexpr int conv_mt() { return TMA ? 1 : CIN == 2 ? PC_CONV_MT : CIN == 4 ? PC_CONV_MT4 : 1; }

template <int CIN, int COUT, int EPI>
constexpr int conv_smem_floats() {
    constexpr int CC = conv_cc<CIN>();
    constexpr int NBUF = ((CIN / CC) > 1 || conv_mt<CIN, false>() > 1) ? 2 : 1;
    return conv_wfloats_padded<CIN, COUT>() + NBUF * CC * SROWS * SPITCH + 8 /* 2 mbarriers */ + 32 /* align slack */;
}

// ---- TMA + mbarrier primitives (staging of whole [channels][34][36] boxes by one thread) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 22)) __trap();   // a bad descriptor must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int c, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(c), "r"(bar) : "memory");
}

// 4-byte async copy global -> shared with zero fill when !ok (src-size 0 reads nothing)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CIN_A, int CIN_B, int COUT, int EPI, bool F32X2, bool TMA>
__global__ void __launch_bounds__(128 * (COUT / 8), (COUT == 8) ? 4 : 2)
conv3x3_kernel(const __grid_constant__ ConvParams p) {
    constexpr int CIN = CIN_A + CIN_B;
    constexpr int CC = conv_cc<CIN>();
    constexpr int NCHUNK = CIN / CC;
    constexpr int MT = conv_mt<CIN, TMA>();               // tiles per CTA along x
    constexpr int NBUF = (NCHUNK > 1 || MT > 1) ? 2 : 1;
    static_assert(MT == 1 || EPI == EPI_STORE || EPI == EPI_POOL, "multi-tile CTAs: epilogues without an early return only");
    constexpr int NT = 128 * (COUT / 8);
    constexpr int NWARP = NT / 32;
    constexpr int WFLOATS = CIN * 9 * COUT + COUT;
    constexpr int XBUF = CC * SROWS * SPITCH;
    static_assert(CIN % CC == 0 && (CIN_A % CC == 0 || CIN_B == 0), "chunks must not straddle sources");

    extern __shared__ __align__(16) float smem_dyn[];
    // 128-byte alignment (TMA destination) as an OFFSET into the shared array: a pointer that went through uintptr_t arithmetic loses its
    // address space and every access through it compiles to a generic LD / ST instead of LDS / STS (cuobjdump: 32 LD, 0 LDS before)
    float* smem = smem_dyn + (((128u - (smem_addr(smem_dyn) & 127u)) & 127u) >> 2);
    float* ws = smem;
    float* xs = smem + conv_wfloats_padded<CIN, COUT>();
    const uint32_t bar0 = smem_addr(xs + NBUF * XBUF);      // two 8-byte mbarriers (TMA path)

    const ConvJob& job = p.jobs[blockIdx.z];
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = tx + 8 * ty + 128 * tz;
    const int lane = tid & 31, warp = tid >> 5;
    int x0 = blockIdx.x * (MT * TILE);                    // advances by TILE per tile of this CTA's run
    const int y0 = blockIdx.y * TILE;
    const int H = p.H, W = p.W;

    // ---- stage CC channels of the (virtual) input tile with a 1-px halo: all copies are issued
    //      back-to-back as cp.async (no register staging, no branches), zero-filled outside ----
    auto stage = [&](int chunk, float* dst) {
        const bool fromA = (chunk * CC) < CIN_A;
        const float* sp; long long cs; int rs, sH, sW, oy, ox, refl, ch0;
        if (fromA) { sp = job.a; cs = job.a_cs; rs = job.a_rs; sH = job.a_H; sW = job.a_W; oy = job.a_oy; ox = job.a_ox; refl = job.a_reflect; ch0 = chunk * CC; }
        else       { sp = job.b; cs = job.b_cs; rs = job.b_rs; sH = job.b_H; sW = job.b_W; oy = job.b_oy; ox = job.b_ox; refl = 0; ch0 = chunk * CC - CIN_A; }
        // source column / validity of a staged column c (image column x0 - 1 + c), source row offset / validity of a staged row r
        auto col_of = [&](int c, int& sx, bool& ok) {
            const int vx = x0 - 1 + c;
            ok = (c < SROWS) && vx >= 0 && vx < W;
            sx = vx - ox;
            if (refl) { sx = sx < 0 ? -sx : sx; sx = sx >= sW ? 2 * (sW - 1) - sx : sx; }
            else ok = ok && sx >= 0 && sx < sW;
            sx = ok ? sx : 0;
        };
        auto row_of = [&](int r, long long& roff, bool& rok) {
            const int vy = y0 - 1 + r;
            rok = r < SROWS && vy >= 0 && vy < H;
            int sy = vy - oy;
            if (refl) { sy = sy < 0 ? -sy : sy; sy = sy >= sH ? 2 * (sH - 1) - sy : sy; }
            else rok = rok && sy >= 0 && sy < sH;
            roff = rok ? (long long)sy * rs : 0;
        };
        int sx0; bool ok0;
        col_of(lane, sx0, ok0);
        const float* planes[CC];
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            int plane = ch0 + c;
            if (CIN_A <= 4 && fromA) plane = (job.a_chmap >> (8 * plane)) & 0xff;
            planes[c] = sp + plane * cs;
        }
        // columns 0..31 of every row: one full-warp copy per (row, channel)
#pragma unroll 3
        for (int r = warp; r < SROWS; r += NWARP) {
            long long roff; bool rok;
            row_of(r, roff, rok);
#pragma unroll
            for (int c = 0; c < CC; ++c)
                cp_async4(dst + (c * SROWS + r) * SPITCH + XOFF + lane, planes[c] + roff + sx0, rok && ok0);
        }
        // the last two columns (32, 33) of 16 rows per copy: lane = (row within the group, column) — instead of one copy per row with
        // two active lanes (half of the kernel's copy instructions were 94 % empty)
        static_assert(SROWS - 32 == 2, "halo column copy assumes two extra columns");
        for (int g = warp; g * 16 < SROWS; g += NWARP) {
            const int r = g * 16 + (lane >> 1), cx = 32 + (lane & 1);
            long long roff; bool rok;
            int sx1; bool ok1;
            row_of(r, roff, rok);
            col_of(cx, sx1, ok1);
            if (r < SROWS) {
#pragma unroll
                for (int c = 0; c < CC; ++c)
                    cp_async4(dst + (c * SROWS + r) * SPITCH + XOFF + cx, planes[c] + roff + sx1, rok && ok1);
            }
        }
    };

    // TMA staging: one thread arms the chunk's mbarrier with the box size and issues ONE bulk tensor copy; the
    // hardware zero-fills everything outside the source tensor (= the conv's zero padding / the Up block's F.pad)
    auto stage_tma = [&](int chunk) {
        const bool fromA = (chunk * CC) < CIN_A;
        const uint32_t bar = bar0 + 8 * (chunk & (NBUF - 1));
        const uint32_t dst = smem_addr(xs + (chunk & (NBUF - 1)) * XBUF);
        mbar_expect_tx(bar, XBUF * 4);
        if (fromA) tma_load_3d(dst, &p.tmA[blockIdx.z], x0 - 1 - XOFF - job.a_ox, y0 - 1 - job.a_oy, chunk * CC, bar);
        else tma_load_3d(dst, &p.tmB[blockIdx.z], x0 - 1 - XOFF - job.b_ox, y0 - 1 - job.b_oy, chunk * CC - CIN_A, bar);
    };

    if (TMA) {
        if (tid == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            stage_tma(0);
        }
    } else {
        stage(0, xs);
        cp_async_commit();
    }
    for (int i = tid; i < WFLOATS / 4; i += NT)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(job.w) + i);

#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt, x0 += TILE) {
    if (MT > 1) {
        if (x0 >= W) break;
        if (mt > 0) __syncthreads();        // every thread is done with the buffer the next prefetch overwrites
    }
    float acc[2][4][8];
    unsigned long long acc2[2][4][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[r][c][o] = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) acc2[r][c][o] = 0ull;
        }

#pragma unroll 1
    for (int chunk = 0; chunk < NCHUNK; ++chunk) {
        float* cur = xs + ((MT > 1 ? mt : chunk) & (NBUF - 1)) * XBUF;
        if (MT > 1) {
            if (mt + 1 < MT && x0 + TILE < W) {         // stage the run's next tile into the other buffer, then wait for this one
                x0 += TILE;
                stage(0, xs + ((mt + 1) & 1) * XBUF);
                x0 -= TILE;
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
        } else if (TMA) {
            if (chunk == 0) __syncthreads();   // weights + barrier init visible to every thread
            if (chunk + 1 < NCHUNK && tid == 0) stage_tma(chunk + 1);   // prefetch into the other buffer
            mbar_wait_parity(bar0 + 8 * (chunk & (NBUF - 1)), (chunk / NBUF) & 1);
        } else {
            if (chunk + 1 < NCHUNK) {           // prefetch the next chunk into the other buffer
                stage(chunk + 1, xs + ((chunk + 1) & (NBUF - 1)) * XBUF);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
        }
        // ---------------- register-tiled stencil: 2x4 pixels x 8 output channels per thread ----------------
        const float* xt = cur + (2 * ty) * SPITCH + 4 * tx;
#pragma unroll 1
        for (int c = 0; c < CC; ++c) {
            float xin[4][6];   // image columns x-1 .. x+4 of the thread's 4-pixel strip = staged columns 4tx+3 .. 4tx+8
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* row = xt + (c * SROWS + r) * SPITCH;
                const float4 v1 = *reinterpret_cast<const float4*>(row + 4);
                xin[r][0] = row[3]; xin[r][1] = v1.x; xin[r][2] = v1.y; xin[r][3] = v1.z; xin[r][4] = v1.w; xin[r][5] = row[8];
            }
            const float* wc = ws + (chunk * CC + c) * 9 * COUT + 8 * tz;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    if (F32X2) {
                        const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wc + (ky * 3 + kx) * COUT);
                        const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(wc + (ky * 3 + kx) * COUT + 4);
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float xv = xin[r + ky][q + kx];
                                const unsigned long long xx = pack2(xv, xv);
                                fma2(acc2[r][q][0], xx, wa.x);
                                fma2(acc2[r][q][1], xx, wa.y);
                                fma2(acc2[r][q][2], xx, wb.x);
                                fma2(acc2[r][q][3], xx, wb.y);
                            }
                    } else {
                        const float4 wa = *reinterpret_cast<const float4*>(wc + (ky * 3 + kx) * COUT);
                        const float4 wb = *reinterpret_cast<const float4*>(wc + (ky * 3 + kx) * COUT + 4);
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float xv = xin[r + ky][q + kx];
                                acc[r][q][0] = fmaf(xv, wa.x, acc[r][q][0]);
                                acc[r][q][1] = fmaf(xv, wa.y, acc[r][q][1]);
                                acc[r][q][2] = fmaf(xv, wa.z, acc[r][q][2]);
                                acc[r][q][3] = fmaf(xv, wa.w, acc[r][q][3]);
                                acc[r][q][4] = fmaf(xv, wb.x, acc[r][q][4]);
                                acc[r][q][5] = fmaf(xv, wb.y, acc[r][q][5]);
                                acc[r][q][6] = fmaf(xv, wb.z, acc[r][q][6]);
                                acc[r][q][7] = fmaf(xv, wb.w, acc[r][q][7]);
                            }
                    }
                }
        }
        if (NBUF > 1 && MT == 1) __syncthreads();  // `cur` is refilled by the prefetch issued at the top of the next-but-one chunk
    }

    // ---------------- epilogue: bias + ReLU, then store / pool / logit dot ----------------
    const float* bias = ws + CIN * 9 * COUT + 8 * tz;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (F32X2) {
#pragma unroll
                for (int o = 0; o < 4; ++o) unpack2(acc2[r][q][o], acc[r][q][2 * o], acc[r][q][2 * o + 1]);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) { const float t = acc[r][q][o] + bias[o]; acc[r][q][o] = job.linear ? t : fmaxf(t, 0.f); }
        }

    const int oy = y0 + 2 * ty, ox = x0 + 4 * tx;
    if (EPI == EPI_DOT) {
        float dw[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) dw[o] = __ldg(job.dotw + o);
        const float db = __ldg(job.dotw + 8);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int yy = oy + r - p.crop_y, xx = ox + q - p.crop_x;
                if (yy < 0 || yy >= p.crop_H || xx < 0 || xx >= p.crop_W) continue;
                float d = 0.f;
#pragma unroll
                for (int o = 0; o < 8; ++o) d = fmaf(acc[r][q][o], dw[o], d);
                if (job.dot_in) d += job.dot_in[(long long)yy * job.dot_in_rs + xx];
                if (job.dot_final) {
                    d += db;
                    d = 1.f / (1.f + expf(-d));
                }
                job.dot_out[(long long)yy * job.dot_out_rs + xx] = d;
            }
        return;
    }

    if (job.out) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int yy = oy + r - p.crop_y;
            if (oy + r >= H || yy < 0 || yy >= p.crop_H) continue;
            const int xx = ox - p.crop_x;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float* dst = job.out + (long long)(8 * tz + o) * job.out_cs + (long long)yy * job.out_rs + xx;
                if (job.out_vec && xx >= 0 && xx + 3 < p.crop_W + 0 * W) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0][o], acc[r][1][o], acc[r][2][o], acc[r][3][o]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (xx + q >= 0 && xx + q < p.crop_W && ox + q < W) dst[q] = acc[r][q][o];
                }
            }
        }
    }
    if (EPI == EPI_POOL) {
        const int py = oy >> 1, px = ox >> 1;
        const int Hp = H >> 1, Wp = W >> 1;
        if (py < Hp) {
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const float m0 = fmaxf(fmaxf(acc[0][0][o], acc[0][1][o]), fmaxf(acc[1][0][o], acc[1][1][o]));
                const float m1 = fmaxf(fmaxf(acc[0][2][o], acc[0][3][o]), fmaxf(acc[1][2][o], acc[1][3][o]));
                float* dst = job.pool + (long long)(8 * tz + o) * job.pool_cs + (long long)py * job.pool_rs + px;
                if (px + 1 < Wp) *reinterpret_cast<float2*>(dst) = make_float2(m0, m1);
                else if (px < Wp) dst[0] = m0;
            }
        }
    }
    }   // tiles of this CTA's run (MT)
}

// ---------------------------------------------------------------------------------------------------
// ConvTranspose2d(k=2, s=2): out[co, 2y+dy, 2x+dx] = b[co] + sum_ci in[ci,y,x] * w[ci][dy*2+dx][co]
// (networks.py:302).  One thread per low-res pixel; weights broadcast from shared memory.
// ---------------------------------------------------------------------------------------------------
struct ConvTJob {
    const float* in; long long in_cs; int in_rs;
    const float* w;  // [C][4][C] then bias[C]
    float* out; long long out_cs; int out_rs;
};
struct ConvTParams {
    int Hl, Wl;
    ConvTJob jobs[MAX_JOBS];
};

template <int C>
__global__ void __launch_bounds__(128) convt2x2_kernel(const __grid_constant__ ConvTParams p) {
    __shared__ __align__(16) float ws[C * 4 * C + C];
    const ConvTJob& job = p.jobs[blockIdx.z];
    for (int i = threadIdx.x + 32 * threadIdx.y; i < (C * 4 * C + C) / 4; i += 128)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(job.w) + i);
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x >= p.Wl || y >= p.Hl) return;
    float xin[C];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) xin[ci] = __ldg(job.in + ci * job.in_cs + (long long)y * job.in_rs + x);
#pragma unroll 1
    for (int cg = 0; cg < C; cg += 8) {  // 8 output channels at a time: 32 accumulators
        float acc[4][8];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[t][o] = ws[C * 4 * C + cg + o];
#pragma unroll
        for (int ci = 0; ci < C; ++ci)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 wa = *reinterpret_cast<const float4*>(ws + (ci * 4 + t) * C + cg);
                const float4 wb = *reinterpret_cast<const float4*>(ws + (ci * 4 + t) * C + cg + 4);
                acc[t][0] = fmaf(xin[ci], wa.x, acc[t][0]); acc[t][1] = fmaf(xin[ci], wa.y, acc[t][1]);
                acc[t][2] = fmaf(xin[ci], wa.z, acc[t][2]); acc[t][3] = fmaf(xin[ci], wa.w, acc[t][3]);
                acc[t][4] = fmaf(xin[ci], wb.x, acc[t][4]); acc[t][5] = fmaf(xin[ci], wb.y, acc[t][5]);
                acc[t][6] = fmaf(xin[ci], wb.z, acc[t][6]); acc[t][7] = fmaf(xin[ci], wb.w, acc[t][7]);
            }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float* dst = job.out + (long long)(cg + o) * job.out_cs + (long long)(2 * y) * job.out_rs + 2 * x;
            *reinterpret_cast<float2*>(dst) = make_float2(acc[0][o], acc[1][o]);
            *reinterpret_cast<float2*>(dst + job.out_rs) = make_float2(acc[2][o], acc[3][o]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Host side: packed-weight layout and the layer schedule
// ---------------------------------------------------------------------------------------------------
static bool first_layer_tc_enabled() {     // POPCORN_CONV_TC_L0=0: keep the first layer on the fp32 stencil even where the tensor-core kernel could run it
    static const bool on = [] {
        const char* e = getenv("POPCORN_CONV_TC_L0");
        return e ? atoi(e) != 0 : true;
    }();
    return on;
}
constexpr int PC_NO_TC = -7001;   // launch_conv<..., EPI_CONVT>: no tensor-core launch was possible (internal, never returned to callers)
struct LayerSpec { int cin, cout, is_t; };
static const LayerSpec kLayers[12] = {
    {-1, 8, 0}, {8, 8, 0}, {8, 16, 0}, {16, 16, 0}, {16, 16, 0}, {16, 16, 0},
    {16, 16, 1}, {32, 8, 0}, {8, 8, 0}, {8, 8, 1}, {16, 8, 0}, {8, 8, 0}};

static int layer_floats(int stream, int layer) {
    const LayerSpec& L = kLayers[layer];
    const int cin = L.cin < 0 ? (stream == 0 ? 2 : 4) : L.cin;
    return cin * (L.is_t ? 4 : 9) * L.cout + L.cout;
}
static int pack_offset(int stream, int layer) {
    int off = 0;
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < 12; ++l) {
            if (s == stream && l == layer) return off;
            off += layer_floats(s, l);
        }
    // out convs: fusion (20), sar (12), optical (12)
    if (layer == 12) return off + (stream == 0 ? 0 : stream == 1 ? 20 : 32);
    return off + 44;
}

// tensor-core weight images (conv_tc.cu) follow the fp32 pack, 256-byte aligned: one image per 3x3 conv layer
static int tc_pack_base() { return (int)round_up(pack_offset(3, 13), 64); }
static int tc_pack_offset(int stream, int layer) {     // floats from the start of the TC section; layer 12 = end
    int off = 0;
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < 12; ++l) {
            if (s == stream && l == layer) return off;
            if (kLayers[l].is_t) continue;
            const int cin = kLayers[l].cin < 0 ? (s == 0 ? 2 : 4) : kLayers[l].cin;
            off += conv_tc_layer_floats(cin, kLayers[l].cout);
        }
    return off;
}

struct Plane {
    float* p; int C, H, W, rs; long long cs;
};

struct Carver {
    char* base; size_t off, cap;
    Plane take(int C, int H, int W) {
        Plane t;
        t.C = C; t.H = H; t.W = W;
        t.rs = (int)round_up(W > 0 ? W : 1, 32);
        t.cs = (long long)t.rs * (H > 0 ? H : 1);
        t.p = reinterpret_cast<float*>(base ? base + off : nullptr);
        off += (size_t)round_up((long long)C * t.cs * 4, 256);
        return t;
    }
};

struct StreamBufs { Plane F0, F1, F2, HA, HB, HC, HD, QA, QB; };

static size_t carve(char* base, int B, int nstream, int Hv, int Wv, StreamBufs* bufs, Plane* dot_tmp) {
    Carver cv{base, 0, 0};
    const int H2 = Hv / 2, W2 = Wv / 2, H4 = H2 / 2, W4 = W2 / 2;
    for (int i = 0; i < B * nstream; ++i) {
        StreamBufs sb;
        sb.F0 = cv.take(8, Hv, Wv); sb.F1 = cv.take(8, Hv, Wv); sb.F2 = cv.take(8, Hv, Wv);
        sb.HA = cv.take(8, H2, W2); sb.HB = cv.take(16, H2, W2); sb.HC = cv.take(16, H2, W2); sb.HD = cv.take(16, H2, W2);
        sb.QA = cv.take(16, H4, W4); sb.QB = cv.take(16, H4, W4);
        if (bufs) bufs[i] = sb;
    }
    for (int b = 0; b < B; ++b) {
        Plane t = cv.take(1, Hv, Wv);
        if (dot_tmp) dot_tmp[b] = t;
    }
    return cv.off;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    return fn;
}

bool make_tmap3d(CUtensorMap* tm, const float* ptr, int C, int H, int W, int rs, long long cs, int boxw, int boxh, int boxc) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !ptr || (((uintptr_t)ptr) & 15) || (rs & 3) || (cs & 3) || C < boxc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C};
    cuuint64_t strides[2] = {(cuuint64_t)rs * 4, (cuuint64_t)cs * 4};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, (cuuint32_t)boxc};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the SIMT kernels stage [cc][34][40] boxes
static bool make_tmap(CUtensorMap* tm, const float* ptr, int C, int H, int W, int rs, long long cs, int cc) {
    return make_tmap3d(tm, ptr, C, H, W, rs, cs, SPITCH, SROWS, cc);
}

template <int CIN_A, int CIN_B, int COUT, int EPI, bool X2, bool TMA>
static int launch_conv_impl(const ConvParams& p, int njobs, cudaStream_t st) {
    constexpr int smem = conv_smem_floats<CIN_A + CIN_B, COUT, EPI>() * 4;
    constexpr int MT = conv_mt<CIN_A + CIN_B, TMA>();
    dim3 grid(cdiv(cdiv(p.W, TILE), MT), cdiv(p.H, TILE), njobs), block(8, 16, COUT / 8);
    auto k = conv3x3_kernel<CIN_A, CIN_B, COUT, EPI, X2, TMA>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, block, smem, st>>>(p);
    return 0;
}

template <int CIN_A, int CIN_B, int COUT, int EPI>
static int launch_conv(ConvParams& p, int njobs, cudaStream_t st) {
    static const bool use_x2 = [] {
        const char* e = getenv("POPCORN_CONV_F32X2");
        return e ? atoi(e) != 0 : true;   // packed fma.rn.f32x2 (FFMA2) is ~13% faster on B200
    }();
    static const bool allow_tma = [] {
        const char* e = getenv("POPCORN_CONV_TMA");
        return e ? atoi(e) != 0 : true;
    }();
    static const int cat = [] {
        char nm[64];
        snprintf(nm, sizeof(nm), "conv3x3<%d,%d,%d,%s>", CIN_A, CIN_B, COUT, EPI == EPI_STORE ? "store" : EPI == EPI_POOL ? "pool" : EPI == EPI_DOT ? "dot" : "convt");
        return prof_register(nm);
    }();
    // tensor-core path (conv_tc.cu): every job carries a pre-swizzled weight image and its sources can be read by TMA
    // (plain planes: not the reflect-padded, channel-remapped first layer, which stays on the fp32 stencil)
    // The first layer (Cin 2 | 4, Cout 8, plain store) qualifies when its source needs no reflection and its channel map is a contiguous
    // run of planes: (c0, c0 + 1) for the SAR stream, the optical stream's (c0 + 2, c0 + 1, c0, c0 + 3) — that permutation is folded into
    // the layer's tensor-core weight image (pc_dda_tc_pack), so ONE TMA box of Cin planes starting at c0 feeds it.
    constexpr bool first_layer = CIN_A < 8;
    bool tc = conv_tc_enabled() && (CIN_A >= 8 || (CIN_B == 0 && COUT == 8 && EPI == EPI_STORE && first_layer_tc_enabled()));
    int c0[MAX_JOBS] = {0};
    for (int j = 0; tc && j < njobs; ++j) {
        const ConvJob& J = p.jobs[j];
        tc = J.wtc != nullptr && !J.a_reflect;
        if (tc && first_layer) {
            c0[j] = (CIN_A == 2) ? (int)(J.a_chmap & 0xff) : (int)((J.a_chmap >> 16) & 0xff);
            const unsigned want = (CIN_A == 2) ? (0x00000100u + 0x00000101u * (unsigned)c0[j]) : (0x03000102u + 0x01010101u * (unsigned)c0[j]);
            tc = J.a_chmap == want && (J.a_ox & 3) == 0;
        }
    }
    if (tc) {
        TcConvParams tp;
        memset(&tp, 0, sizeof(tp));
        tp.H = p.H; tp.W = p.W; tp.crop_y = p.crop_y; tp.crop_x = p.crop_x; tp.crop_H = p.crop_H; tp.crop_W = p.crop_W;
        for (int j = 0; tc && j < njobs; ++j) {
            const ConvJob& J = p.jobs[j];
            tp.jobs[j] = J;
            tc = make_tmap3d(&tp.tmA[j], J.a + (long long)c0[j] * J.a_cs, CIN_A, J.a_H, J.a_W, J.a_rs, J.a_cs, TC_BOXW, 1, CIN_A);
            if (tc && CIN_B > 0) tc = make_tmap3d(&tp.tmB[j], J.b, CIN_B, J.b_H, J.b_W, J.b_rs, J.b_cs, TC_BOXW, 1, CIN_B);
        }
        if (tc) return launch_conv_tc(CIN_A, CIN_B, COUT, EPI, tp, njobs, st);
    }
    if constexpr (EPI == EPI_CONVT) {
        return PC_NO_TC;       // the fused transposed-conv epilogue exists on the tensor-core path only: the caller runs conv + convT kernels
    } else {
    constexpr int CC = conv_cc<CIN_A + CIN_B>();
    // TMA staging needs plain (non-reflected, identity-channel) 16-byte-aligned sources: every layer but the first
    bool tma = allow_tma && CIN_A >= 8;
    for (int j = 0; tma && j < njobs; ++j) {
        const ConvJob& J = p.jobs[j];
        tma = !J.a_reflect && (J.a_ox & 3) == 0 && make_tmap(&p.tmA[j], J.a, CIN_A, J.a_H, J.a_W, J.a_rs, J.a_cs, CC);
        if (tma && CIN_B > 0) tma = (J.b_ox & 3) == 0 && make_tmap(&p.tmB[j], J.b, CIN_B, J.b_H, J.b_W, J.b_rs, J.b_cs, CC);
    }
    ProfScope prof(cat, st, (double)p.H * p.W * njobs);
    int rc;
    if (tma) rc = use_x2 ? launch_conv_impl<CIN_A, CIN_B, COUT, EPI, true, true>(p, njobs, st)
                         : launch_conv_impl<CIN_A, CIN_B, COUT, EPI, false, true>(p, njobs, st);
    else rc = use_x2 ? launch_conv_impl<CIN_A, CIN_B, COUT, EPI, true, false>(p, njobs, st)
                     : launch_conv_impl<CIN_A, CIN_B, COUT, EPI, false, false>(p, njobs, st);
    if (rc) return rc;
    PC_LAUNCH_CHECK();
    return 0;
    }
}

static void set_a(ConvJob& j, const Plane& t) {
    j.a = t.p; j.a_cs = t.cs; j.a_rs = t.rs; j.a_H = t.H; j.a_W = t.W; j.a_oy = 0; j.a_ox = 0; j.a_reflect = 0;
    j.a_chmap = 0x03020100u;
}
static void set_out(ConvJob& j, const Plane& t) {
    j.out = t.p; j.out_cs = t.cs; j.out_rs = t.rs; j.out_vec = 1;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_dda_pack_floats(void) { return pack_offset(3, 13); }
extern "C" int pc_dda_pack_offset(int stream, int layer) {
    if (layer < 0 || layer > 12 || stream < 0 || stream > 2 || (layer < 12 && stream > 1)) return -1;
    return pack_offset(stream, layer);
}

extern "C" size_t pc_dda_workspace_bytes(int B, int C, int Hv, int Wv) {
    const int ns = (C == 6) ? 2 : 1;
    return carve(nullptr, B, ns, Hv, Wv, nullptr, nullptr) + 256;
}

extern "C" int pc_dda_tc_pack_base(void) { return tc_pack_base(); }
extern "C" int pc_dda_tc_pack_floats(void) { return tc_pack_offset(2, 12); }
extern "C" int pc_dda_tc_pack(const float* flat_host, float* img_host) {
    PC_CHECK_ARG(flat_host && img_host, "null pointer");
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < 12; ++l) {
            if (kLayers[l].is_t) continue;
            const int cin = kLayers[l].cin < 0 ? (s == 0 ? 2 : 4) : kLayers[l].cin;
            const float* flat = flat_host + pack_offset(s, l);
            float perm[4 * 9 * 8 + 8];
            if (s == 1 && l == 0) {
                // optical first layer: the tensor-core kernel reads the planes in MEMORY order (R, G, B, NIR) with one TMA box, the fp32 pack
                // is in the network's channel order (B, G, R, NIR = planes 2, 1, 0, 3): plane q holds logical channel (2, 1, 0, 3)[q]
                static const int logical_of_plane[4] = {2, 1, 0, 3};
                const int per = 9 * kLayers[l].cout;
                for (int q = 0; q < 4; ++q) memcpy(perm + q * per, flat + logical_of_plane[q] * per, sizeof(float) * per);
                memcpy(perm + 4 * per, flat + 4 * per, sizeof(float) * kLayers[l].cout);
                flat = perm;
            }
            conv_tc_pack_layer(flat, cin, kLayers[l].cout, img_host + tc_pack_offset(s, l));
        }
    return 0;
}
extern "C" int pc_conv_tc_layer_floats(int cin, int cout) { return conv_tc_layer_floats(cin, cout); }
extern "C" int pc_conv_tc_pack_layer(const float* flat_host, int cin, int cout, float* img_host) {
    PC_CHECK_ARG(flat_host && img_host && cin >= 1 && cin <= 32 && (cout == 8 || cout == 16), "bad argument");
    conv_tc_pack_layer(flat_host, cin, cout, img_host);
    return 0;
}

extern "C" int pc_dda_forward(const float* wpack, long long wpack_floats, const float* x, int B, int C, int H, int W, long long x_bstride,
                              long long x_cstride, int x_rstride, int pad_top, int pad_bottom, int pad_left,
                              int pad_right, int mode, float* out, long long out_bstride, long long out_cstride,
                              int out_rstride, void* workspace, size_t workspace_bytes, pc_stream_t stream) {
    PC_CHECK_ARG(wpack && x && out && workspace, "null pointer");
    PC_CHECK_ARG(C == 6 || C == 2 || C == 4, "input channels must be 6 (S1+S2+NIR), 2 (S1) or 4 (S2+NIR)");
    PC_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "bad shape");
    PC_CHECK_ARG(pad_top >= 0 && pad_bottom >= 0 && pad_left >= 0 && pad_right >= 0, "negative padding");
    PC_CHECK_ARG(pad_top < H && pad_bottom < H && pad_left < W && pad_right < W, "reflect padding must be < dim");
    PC_CHECK_ARG(mode == PC_DDA_FEATURES || mode == PC_DDA_BUILTUP, "bad mode");
    const int Hv = H + pad_top + pad_bottom, Wv = W + pad_left + pad_right;
    PC_CHECK_ARG(Hv >= 4 && Wv >= 4, "window too small for two 2x2 pools");
    const bool S1 = (C != 4), S2 = (C != 2);
    const int ns = (S1 ? 1 : 0) + (S2 ? 1 : 0);
    const size_t need = pc_dda_workspace_bytes(B, C, Hv, Wv);
    if (workspace_bytes < need) {
        set_error("pc_dda_forward: workspace %zu < required %zu", workspace_bytes, need);
        return PC_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* wsb = reinterpret_cast<char*>(round_up((long long)(uintptr_t)workspace, 256));
    const int H2 = Hv / 2, W2 = Wv / 2, H4 = H2 / 2, W4 = W2 / 2;

    // stream ids present, in feature-channel order (sar first): networks.py:197-207
    int sids[2]; int n = 0;
    if (S1) sids[n++] = 0;
    if (S2) sids[n++] = 1;

    for (int b0 = 0; b0 < B; b0 += MAX_JOBS / 2) {
        const int nb = (B - b0) < MAX_JOBS / 2 ? (B - b0) : MAX_JOBS / 2;
        StreamBufs bufs[MAX_JOBS]; Plane dot_tmp[MAX_JOBS];
        carve(wsb, nb, ns, Hv, Wv, bufs, dot_tmp);
        const int nj = nb * ns;
        auto W_ = [&](int s, int l) { return wpack + pack_offset(s, l); };
        const bool have_tc = wpack_floats >= (long long)tc_pack_base() + tc_pack_offset(2, 12);
        auto WT_ = [&](int s, int l) -> const float* { return have_tc ? wpack + tc_pack_base() + tc_pack_offset(s, l) : nullptr; };
        auto jb = [&](int j) { return b0 + j / ns; };     // batch index of job j
        auto js = [&](int j) { return sids[j % ns]; };    // stream id of job j
        ConvParams p;
        ConvTParams pt;
        auto reset = [&](int Hh, int Ww) {
            memset(&p, 0, sizeof(p));
            p.H = Hh; p.W = Ww; p.crop_y = 0; p.crop_x = 0; p.crop_H = Hh; p.crop_W = Ww;
        };
        int rc;
        // ---- L0 inc.conv.0 : virtual reflect-padded, channel-reordered input -> F0 (per stream: Cin 2 | 4)
        for (int si = 0; si < ns; ++si) {
            reset(Hv, Wv);
            const int s = sids[si];
            for (int k = 0; k < nb; ++k) {
                ConvJob& j = p.jobs[k];
                j.a = x + (long long)(b0 + k) * x_bstride; j.a_cs = x_cstride; j.a_rs = x_rstride;
                j.a_H = H; j.a_W = W; j.a_oy = pad_top; j.a_ox = pad_left;
                // without virtual padding the virtual image IS the source: "outside" = the conv's zero padding, which a plain (TMA) load gives too
                j.a_reflect = (pad_top | pad_bottom | pad_left | pad_right) ? 1 : 0;
                // [R,G,B,NIR,VV,VH] -> sar (VV,VH) | optical (B,G,R,NIR)   popcorn.py:130-134
                if (C == 6) j.a_chmap = (s == 0) ? 0x00000504u : 0x03000102u;
                else if (C == 2) j.a_chmap = 0x00000100u;
                else j.a_chmap = 0x03000102u;
                j.w = W_(s, 0); j.wtc = WT_(s, 0);
                set_out(j, bufs[k * ns + si].F0);
            }
            rc = (s == 0) ? launch_conv<2, 0, 8, EPI_STORE>(p, nb, st) : launch_conv<4, 0, 8, EPI_STORE>(p, nb, st);
            if (rc) return rc;
        }
        // ---- L1 inc.conv.3 : F0 -> F1 (+ pooled HA)
        reset(Hv, Wv);
        for (int j = 0; j < nj; ++j) {
            ConvJob& J = p.jobs[j]; set_a(J, bufs[j].F0); J.w = W_(js(j), 1); J.wtc = WT_(js(j), 1); set_out(J, bufs[j].F1);
            J.pool = bufs[j].HA.p; J.pool_cs = bufs[j].HA.cs; J.pool_rs = bufs[j].HA.rs;
        }
        if ((rc = launch_conv<8, 0, 8, EPI_POOL>(p, nj, st))) return rc;
        // ---- L2 down1.conv.0 : HA -> HB(16)
        reset(H2, W2);
        for (int j = 0; j < nj; ++j) { ConvJob& J = p.jobs[j]; set_a(J, bufs[j].HA); J.w = W_(js(j), 2); J.wtc = WT_(js(j), 2); set_out(J, bufs[j].HB); }
        if ((rc = launch_conv<8, 0, 16, EPI_STORE>(p, nj, st))) return rc;
        // ---- L3 down1.conv.3 : HB -> HC (+ pooled QA)
        reset(H2, W2);
        for (int j = 0; j < nj; ++j) {
            ConvJob& J = p.jobs[j]; set_a(J, bufs[j].HB); J.w = W_(js(j), 3); J.wtc = WT_(js(j), 3); set_out(J, bufs[j].HC);
            J.pool = bufs[j].QA.p; J.pool_cs = bufs[j].QA.cs; J.pool_rs = bufs[j].QA.rs;
        }
        if ((rc = launch_conv<16, 0, 16, EPI_POOL>(p, nj, st))) return rc;
        // ---- L4 down2.conv.0 : QA -> QB ; L5 down2.conv.3 : QB -> QA
        reset(H4, W4);
        for (int j = 0; j < nj; ++j) { ConvJob& J = p.jobs[j]; set_a(J, bufs[j].QA); J.w = W_(js(j), 4); J.wtc = WT_(js(j), 4); set_out(J, bufs[j].QB); }
        if ((rc = launch_conv<16, 0, 16, EPI_STORE>(p, nj, st))) return rc;
        // L5 + L6: down2.conv.3 with the Up block's transposed conv in its epilogue (QB -> HD [16, 2*H4, 2*W4]; the quarter-resolution
        // activation never reaches HBM); without the tensor-core path: conv -> QA, then the convT kernel
        static const bool fuse_ct = [] { const char* e = getenv("POPCORN_FUSE_CONVT"); return e ? atoi(e) != 0 : true; }();
        rc = PC_NO_TC;
        if (fuse_ct && have_tc) {
            reset(H4, W4);
            for (int j = 0; j < nj; ++j) {
                ConvJob& J = p.jobs[j]; set_a(J, bufs[j].QB); J.w = W_(js(j), 5); J.wtc = WT_(js(j), 5);
                J.ctw = W_(js(j), 6); J.ct_out = bufs[j].HD.p; J.ct_cs = bufs[j].HD.cs; J.ct_rs = bufs[j].HD.rs;
            }
            rc = launch_conv<16, 0, 16, EPI_CONVT>(p, nj, st);
            if (rc && rc != PC_NO_TC) return rc;
        }
        if (rc == PC_NO_TC) {
            reset(H4, W4);
            for (int j = 0; j < nj; ++j) { ConvJob& J = p.jobs[j]; set_a(J, bufs[j].QB); J.w = W_(js(j), 5); J.wtc = WT_(js(j), 5); set_out(J, bufs[j].QA); }
            if ((rc = launch_conv<16, 0, 16, EPI_STORE>(p, nj, st))) return rc;
            // ---- L6 up2.up : QA -> HD [16, 2*H4, 2*W4]
            memset(&pt, 0, sizeof(pt)); pt.Hl = H4; pt.Wl = W4;
            for (int j = 0; j < nj; ++j) {
                ConvTJob& J = pt.jobs[j]; J.in = bufs[j].QA.p; J.in_cs = bufs[j].QA.cs; J.in_rs = bufs[j].QA.rs;
                J.w = W_(js(j), 6); J.out = bufs[j].HD.p; J.out_cs = bufs[j].HD.cs; J.out_rs = bufs[j].HD.rs;
            }
            {
                static const int cat = prof_register("convt2x2<16>");
                ProfScope prof(cat, st, (double)H4 * W4 * nj);
                convt2x2_kernel<16><<<dim3(cdiv(W4, 32), cdiv(H4, 4), nj), dim3(32, 4), 0, st>>>(pt);
            }
            PC_LAUNCH_CHECK();
        }
        // ---- L7 up2.conv.0 : cat[HC(16), pad(HD)(16)] -> HA(8) ; L8 up2.conv.3 : HA -> HB(8)
        reset(H2, W2);
        for (int j = 0; j < nj; ++j) {
            ConvJob& J = p.jobs[j]; set_a(J, bufs[j].HC);
            J.b = bufs[j].HD.p; J.b_cs = bufs[j].HD.cs; J.b_rs = bufs[j].HD.rs; J.b_H = 2 * H4; J.b_W = 2 * W4;
            J.b_oy = (H2 - 2 * H4) / 2; J.b_ox = (W2 - 2 * W4) / 2;   // F.pad split, networks.py:309-312
            J.w = W_(js(j), 7); J.wtc = WT_(js(j), 7); set_out(J, bufs[j].HA);
        }
        if ((rc = launch_conv<16, 16, 8, EPI_STORE>(p, nj, st))) return rc;
        // L8 + L9: up2.conv.3 with up1's transposed conv in its epilogue (HA -> F2 [8, 2*H2, 2*W2])
        rc = PC_NO_TC;
        if (fuse_ct && have_tc) {
            reset(H2, W2);
            for (int j = 0; j < nj; ++j) {
                ConvJob& J = p.jobs[j]; set_a(J, bufs[j].HA); J.w = W_(js(j), 8); J.wtc = WT_(js(j), 8);
                J.ctw = W_(js(j), 9); J.ct_out = bufs[j].F2.p; J.ct_cs = bufs[j].F2.cs; J.ct_rs = bufs[j].F2.rs;
            }
            rc = launch_conv<8, 0, 8, EPI_CONVT>(p, nj, st);
            if (rc && rc != PC_NO_TC) return rc;
        }
        if (rc == PC_NO_TC) {
            reset(H2, W2);
            for (int j = 0; j < nj; ++j) { ConvJob& J = p.jobs[j]; set_a(J, bufs[j].HA); J.w = W_(js(j), 8); J.wtc = WT_(js(j), 8); set_out(J, bufs[j].HB); }
            if ((rc = launch_conv<8, 0, 8, EPI_STORE>(p, nj, st))) return rc;
            // ---- L9 up1.up : HB(8) -> F2 [8, 2*H2, 2*W2]
            memset(&pt, 0, sizeof(pt)); pt.Hl = H2; pt.Wl = W2;
            for (int j = 0; j < nj; ++j) {
                ConvTJob& J = pt.jobs[j]; J.in = bufs[j].HB.p; J.in_cs = bufs[j].HB.cs; J.in_rs = bufs[j].HB.rs;
                J.w = W_(js(j), 9); J.out = bufs[j].F2.p; J.out_cs = bufs[j].F2.cs; J.out_rs = bufs[j].F2.rs;
            }
            {
                static const int cat = prof_register("convt2x2<8>");
                ProfScope prof(cat, st, (double)H2 * W2 * nj);
                convt2x2_kernel<8><<<dim3(cdiv(W2, 32), cdiv(H2, 4), nj), dim3(32, 4), 0, st>>>(pt);
            }
            PC_LAUNCH_CHECK();
        }
        // ---- L10 up1.conv.0 : cat[F1(8), pad(F2)(8)] -> F0(8)
        reset(Hv, Wv);
        for (int j = 0; j < nj; ++j) {
            ConvJob& J = p.jobs[j]; set_a(J, bufs[j].F1);
            J.b = bufs[j].F2.p; J.b_cs = bufs[j].F2.cs; J.b_rs = bufs[j].F2.rs; J.b_H = 2 * H2; J.b_W = 2 * W2;
            J.b_oy = (Hv - 2 * H2) / 2; J.b_ox = (Wv - 2 * W2) / 2;
            J.w = W_(js(j), 10); J.wtc = WT_(js(j), 10); set_out(J, bufs[j].F0);
        }
        if ((rc = launch_conv<8, 8, 8, EPI_STORE>(p, nj, st))) return rc;
        // ---- L11 up1.conv.3 : F0 -> features (cropped) | logit dot (+sigmoid, cropped)
        if (mode == PC_DDA_FEATURES) {
            reset(Hv, Wv);
            p.crop_y = pad_top; p.crop_x = pad_left; p.crop_H = H; p.crop_W = W;
            for (int j = 0; j < nj; ++j) {
                ConvJob& J = p.jobs[j]; set_a(J, bufs[j].F0); J.w = W_(js(j), 11); J.wtc = WT_(js(j), 11);
                J.out = out + (long long)jb(j) * out_bstride + (long long)(8 * (j % ns)) * out_cstride;
                J.out_cs = out_cstride; J.out_rs = out_rstride;
                J.out_vec = (pad_left % 4 == 0) && (out_rstride % 4 == 0) && (out_cstride % 4 == 0) &&
                            (out_bstride % 4 == 0) && (((uintptr_t)out) % 16 == 0);
            }
            if ((rc = launch_conv<8, 0, 8, EPI_STORE>(p, nj, st))) return rc;
        } else {
            for (int si = 0; si < ns; ++si) {
                reset(Hv, Wv);
                p.crop_y = pad_top; p.crop_x = pad_left; p.crop_H = H; p.crop_W = W;
                const int s = sids[si];
                const bool last = (si == ns - 1);
                for (int k = 0; k < nb; ++k) {
                    ConvJob& J = p.jobs[k]; set_a(J, bufs[k * ns + si].F0); J.w = W_(s, 11); J.wtc = WT_(s, 11);
                    // fusion_out_conv weights [sar 0:8 | optical 8:16] + bias; single-modality: own out conv
                    const float* oc = (ns == 2) ? wpack + pack_offset(0, 12) : wpack + pack_offset(s + 1, 12);
                    J.dotw = (ns == 2) ? oc + 8 * si : oc;
                    // bias sits after the weights: fusion at +16, single at +8 -> pass via a 9-float view
                    J.dot_final = last ? 1 : 0;
                    J.dot_in = (si > 0) ? dot_tmp[k].p : nullptr; J.dot_in_rs = dot_tmp[k].rs;
                    if (last) { J.dot_out = out + (long long)(b0 + k) * out_bstride; J.dot_out_rs = out_rstride; }
                    else { J.dot_out = dot_tmp[k].p; J.dot_out_rs = dot_tmp[k].rs; }
                }
                if ((rc = launch_conv<8, 0, 8, EPI_DOT>(p, nb, st))) return rc;
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Unit-test hooks: one conv / convT layer on plain contiguous tensors (tests/test_gpu_kernels.py).
// ---------------------------------------------------------------------------------------------------
extern "C" int pc_test_conv3x3(const float* a, int cin_a, int a_H, int a_W, int a_oy, int a_ox, int a_reflect,
                               const float* b, int cin_b, int b_H, int b_W, int b_oy, int b_ox, const float* w,
                               int cout, int H, int W, float* out, float* pool, const float* wtc, pc_stream_t stream) {
    PC_CHECK_ARG(a && w && out, "null pointer");
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.crop_H = H; p.crop_W = W;
    ConvJob& J = p.jobs[0];
    J.a = a; J.a_cs = (long long)a_H * a_W; J.a_rs = a_W; J.a_H = a_H; J.a_W = a_W; J.a_oy = a_oy; J.a_ox = a_ox;
    J.a_reflect = a_reflect; J.a_chmap = 0x03020100u;
    J.b = b; J.b_cs = (long long)b_H * b_W; J.b_rs = b_W; J.b_H = b_H; J.b_W = b_W; J.b_oy = b_oy; J.b_ox = b_ox;
    J.w = w; J.wtc = wtc; J.out = out; J.out_cs = (long long)H * W; J.out_rs = W;
    J.out_vec = (W % 4 == 0) && (((uintptr_t)out) % 16 == 0);
    J.pool = pool; J.pool_cs = (long long)(H / 2) * (W / 2); J.pool_rs = W / 2;
    cudaStream_t st = (cudaStream_t)stream;
    const int key = cin_a * 10000 + cin_b * 100 + cout;
    if (pool) {
        PC_CHECK_ARG((W / 2) % 2 == 0, "pooled row stride must be even for the float2 stores");
        if (key == 80008) return launch_conv<8, 0, 8, EPI_POOL>(p, 1, st);
        if (key == 160016) return launch_conv<16, 0, 16, EPI_POOL>(p, 1, st);
        PC_CHECK_ARG(false, "no pooled instantiation for this shape");
    }
    switch (key) {
        case 20008: return launch_conv<2, 0, 8, EPI_STORE>(p, 1, st);
        case 40008: return launch_conv<4, 0, 8, EPI_STORE>(p, 1, st);
        case 80008: return launch_conv<8, 0, 8, EPI_STORE>(p, 1, st);
        case 80016: return launch_conv<8, 0, 16, EPI_STORE>(p, 1, st);
        case 160016: return launch_conv<16, 0, 16, EPI_STORE>(p, 1, st);
        case 161608: return launch_conv<16, 16, 8, EPI_STORE>(p, 1, st);
        case 80808: return launch_conv<8, 8, 8, EPI_STORE>(p, 1, st);
    }
    PC_CHECK_ARG(false, "no instantiation for this (cin_a, cin_b, cout)");
}

// ---------------------------------------------------------------------------------------------------
// Layer-level entry points with explicit strides (the training path drives the UNet layer by layer from
// Python so that autograd can keep every activation: popcorn_b200/model/unet_train.py).
// ---------------------------------------------------------------------------------------------------
extern "C" int pc_conv3x3_layer(const float* a, int cin_a, long long a_cs, int a_rs, int a_H, int a_W, int a_oy, int a_ox,
                                int a_reflect, unsigned a_chmap, const float* b, int cin_b, long long b_cs, int b_rs, int b_H,
                                int b_W, int b_oy, int b_ox, const float* w, const float* wtc, int cout, int relu, int H,
                                int W, float* out, long long out_cs, int out_rs, float* pool, long long pool_cs, int pool_rs,
                                pc_stream_t stream) {
    PC_CHECK_ARG(a && w && out, "null pointer");
    PC_CHECK_ARG(H >= 1 && W >= 1, "bad shape");
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.crop_H = H; p.crop_W = W;
    ConvJob& J = p.jobs[0];
    J.a = a; J.a_cs = a_cs; J.a_rs = a_rs; J.a_H = a_H; J.a_W = a_W; J.a_oy = a_oy; J.a_ox = a_ox;
    J.a_reflect = a_reflect; J.a_chmap = a_chmap;
    J.b = b; J.b_cs = b_cs; J.b_rs = b_rs; J.b_H = b_H; J.b_W = b_W; J.b_oy = b_oy; J.b_ox = b_ox;
    J.w = w; J.wtc = wtc; J.linear = relu ? 0 : 1;
    J.out = out; J.out_cs = out_cs; J.out_rs = out_rs;
    J.out_vec = (out_rs % 4 == 0) && (out_cs % 4 == 0) && (((uintptr_t)out) % 16 == 0);
    J.pool = pool; J.pool_cs = pool_cs; J.pool_rs = pool_rs;
    cudaStream_t st = (cudaStream_t)stream;
    const int key = cin_a * 10000 + cin_b * 100 + cout;
    if (pool) {
        PC_CHECK_ARG(pool_rs % 2 == 0 && pool_cs % 2 == 0 && (((uintptr_t)pool) % 8 == 0), "pooled planes must be 8-byte aligned with even strides");
        if (key == 80008) return launch_conv<8, 0, 8, EPI_POOL>(p, 1, st);
        if (key == 160016) return launch_conv<16, 0, 16, EPI_POOL>(p, 1, st);
        PC_CHECK_ARG(false, "no pooled instantiation for this shape");
    }
    switch (key) {
        case 20008: return launch_conv<2, 0, 8, EPI_STORE>(p, 1, st);
        case 40008: return launch_conv<4, 0, 8, EPI_STORE>(p, 1, st);
        case 80008: return launch_conv<8, 0, 8, EPI_STORE>(p, 1, st);
        case 80016: return launch_conv<8, 0, 16, EPI_STORE>(p, 1, st);
        case 160016: return launch_conv<16, 0, 16, EPI_STORE>(p, 1, st);
        case 161608: return launch_conv<16, 16, 8, EPI_STORE>(p, 1, st);
        case 80808: return launch_conv<8, 8, 8, EPI_STORE>(p, 1, st);
    }
    PC_CHECK_ARG(false, "no instantiation for this (cin_a, cin_b, cout)");
}

extern "C" int pc_convt2x2_layer(const float* in, int C, long long in_cs, int in_rs, int Hl, int Wl, const float* w, float* out,
                                 long long out_cs, int out_rs, pc_stream_t stream) {
    PC_CHECK_ARG(in && w && out, "null pointer");
    PC_CHECK_ARG(C == 8 || C == 16, "C must be 8 or 16");
    PC_CHECK_ARG(out_rs % 2 == 0 && out_cs % 2 == 0 && (((uintptr_t)out) % 8 == 0), "output planes must be 8-byte aligned with even strides");
    if (Hl <= 0 || Wl <= 0) return 0;
    ConvTParams pt;
    memset(&pt, 0, sizeof(pt));
    pt.Hl = Hl; pt.Wl = Wl;
    ConvTJob& J = pt.jobs[0];
    J.in = in; J.in_cs = in_cs; J.in_rs = in_rs; J.w = w; J.out = out; J.out_cs = out_cs; J.out_rs = out_rs;
    dim3 grid(cdiv(Wl, 32), cdiv(Hl, 4), 1), block(32, 4);
    if (C == 8) convt2x2_kernel<8><<<grid, block, 0, (cudaStream_t)stream>>>(pt);
    else convt2x2_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>(pt);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_test_convt2x2(const float* in, int C, int Hl, int Wl, const float* w, float* out, pc_stream_t stream) {
    PC_CHECK_ARG(in && w && out, "null pointer");
    PC_CHECK_ARG(C == 8 || C == 16, "C must be 8 or 16");
    ConvTParams pt;
    memset(&pt, 0, sizeof(pt));
    pt.Hl = Hl; pt.Wl = Wl;
    ConvTJob& J = pt.jobs[0];
    J.in = in; J.in_cs = (long long)Hl * Wl; J.in_rs = Wl; J.w = w;
    J.out = out; J.out_cs = 4ll * Hl * Wl; J.out_rs = 2 * Wl;
    dim3 grid(cdiv(Wl, 32), cdiv(Hl, 4), 1), block(32, 4);
    if (C == 8) convt2x2_kernel<8><<<grid, block, 0, (cudaStream_t)stream>>>(pt);
    else convt2x2_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>(pt);
    PC_LAUNCH_CHECK();
    return 0;
}
