// Backward of the DDA UNet layers (fine-tuning `unetmodel`, the reference's default for
// batches below 9 M pixels, run_train.py:191-202).  Replaces the autograd backward of
// model/DDA_model/utils/networks.py:253-271 (Conv2d+BN(eval)+ReLU), :289 (MaxPool2d), :302 (ConvTranspose2d).
//
//   dgrad of a 3x3 conv   = the forward conv kernels (conv.cu / conv_tc.cu) on the ReLU-masked gradient with the
//                           transposed, tap-flipped weights and the `linear` epilogue — scheduled from Python;
//   wgrad of a 3x3 conv   = conv_wgrad_kernel: dW[ci][tap][co] = sum_px in[ci](y+ky-1, x+kx-1) * g[co](y, x),
//                           db[co] = sum_px g[co]; per-CTA partials over 32x32-pixel tiles, fixed-order second stage
//                           (no float atomics: bitwise reproducible, comparable at 1e-3);
//   ReLU / max-pool       = relu_mask_kernel, pool_bwd_kernel (first-maximum rule of ATen's max_pool2d backward);
//   ConvTranspose2d k2 s2 = convt_dgrad_kernel, convt_wgrad_kernel.
// All tensors are fp32 planes with explicit strides; gradients w.r.t. the FOLDED weights come back in the packed
// layouts of include/popcorn_b200.h and are unfolded to (W, b, gamma, beta) on the host (tiny tensors).
#include <stdlib.h>

#include "common.cuh"

namespace pc {

constexpr int WT = 32;             // wgrad tile edge
constexpr int WIP = WT + 3;        // staged input pitch: 34 columns used, odd pitch -> distinct banks per channel
constexpr int WGP = WT * (WT + 1) + 1;   // gradient plane pitch: odd, so the `cout` planes a warp reads at one (r, c) sit in distinct banks
                                   // (32 * 33 is a multiple of 32: every channel hit the same bank, an 8- to 16-way conflict on a quarter of the loads)

struct WgradArgs {
    const float* a; long long a_cs; int a_rs; int a_H, a_W, a_oy, a_ox, a_reflect; unsigned a_chmap; int cin_a;
    const float* b; long long b_cs; int b_rs; int b_H, b_W, b_oy, b_ox; int cin_b;
    const float* g; long long g_cs; int g_rs;     // [cout][H][W] ReLU-masked output gradient
    int cout, H, W, tiles_x;
    float* partial;                                // [tiles][cin*9*cout + cout]
};

// one CTA = one 32x32 tile; thread owns one (ci, co) pair (and a slice of the rows when there are fewer pairs than threads)
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ float sm[];
    const int cin = a.cin_a + a.cin_b, cout = a.cout;
    const int NP = cin * cout;                      // <= 256
    const int G = 256 / NP;                         // row groups (1, 2 or 4)
    float* xs = sm;                                 // [cin][34][WIP]
    float* gs = xs + cin * (WT + 2) * WIP;          // [cout][32][33]
    float* red = gs + cout * WGP;                   // [G][NP][10]
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
    const int x0 = tx * WT, y0 = ty * WT;
    // ---- stage the input tile with a 1-px halo (zero outside the virtual image; reflect / channel map / concat as the forward loader)
    // unrolled: eight independent global loads in flight per thread (the tile is latency-bound otherwise: one DRAM round trip per iteration)
#pragma unroll 8
    for (int i = tid; i < cin * (WT + 2) * (WT + 2); i += 256) {
        const int c = i / ((WT + 2) * (WT + 2)), r = (i / (WT + 2)) % (WT + 2), col = i % (WT + 2);
        const int vy = y0 - 1 + r, vx = x0 - 1 + col;
        float v = 0.f;
        if (vy >= 0 && vy < a.H && vx >= 0 && vx < a.W) {
            if (c < a.cin_a) {
                int sy = vy - a.a_oy, sx = vx - a.a_ox;
                bool ok = true;
                if (a.a_reflect) {
                    sy = sy < 0 ? -sy : sy; sy = sy >= a.a_H ? 2 * (a.a_H - 1) - sy : sy;
                    sx = sx < 0 ? -sx : sx; sx = sx >= a.a_W ? 2 * (a.a_W - 1) - sx : sx;
                } else ok = sy >= 0 && sy < a.a_H && sx >= 0 && sx < a.a_W;
                int plane = c;
                if (a.cin_a <= 4) plane = (a.a_chmap >> (8 * c)) & 0xff;
                if (ok) v = __ldg(a.a + plane * a.a_cs + (long long)sy * a.a_rs + sx);
            } else {
                const int sy = vy - a.b_oy, sx = vx - a.b_ox;
                if (sy >= 0 && sy < a.b_H && sx >= 0 && sx < a.b_W) v = __ldg(a.b + (c - a.cin_a) * a.b_cs + (long long)sy * a.b_rs + sx);
            }
        }
        xs[(c * (WT + 2) + r) * WIP + col] = v;
    }
#pragma unroll 8
    for (int i = tid; i < cout * WT * WT; i += 256) {
        const int c = i / (WT * WT), r = (i / WT) % WT, col = i % WT;
        const int vy = y0 + r, vx = x0 + col;
        gs[c * WGP + r * (WT + 1) + col] = (vy < a.H && vx < a.W) ? __ldg(a.g + c * a.g_cs + (long long)vy * a.g_rs + vx) : 0.f;
    }
    __syncthreads();
    float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float accb = 0.f;
    const int pair = tid % NP, grp = tid / NP;
    const int co = pair % cout, ci = pair / cout;
    if (grp < G) {
        const int rows = WT / G;
        const float* xc = xs + ci * (WT + 2) * WIP;
        const float* gc = gs + co * WGP;
        for (int r = grp * rows; r < (grp + 1) * rows; ++r) {
            const float* x0r = xc + r * WIP;          // input rows r, r+1, r+2 hold image rows y-1, y, y+1
            const float* x1r = x0r + WIP;
            const float* x2r = x1r + WIP;
            float w0 = x0r[0], w1 = x0r[1], m0 = x1r[0], m1 = x1r[1], s0 = x2r[0], s1 = x2r[1];
#pragma unroll 4
            for (int c = 0; c < WT; ++c) {
                const float w2 = x0r[c + 2], m2 = x1r[c + 2], s2 = x2r[c + 2];
                const float gv = gc[r * (WT + 1) + c];
                acc[0] = fmaf(w0, gv, acc[0]); acc[1] = fmaf(w1, gv, acc[1]); acc[2] = fmaf(w2, gv, acc[2]);
                acc[3] = fmaf(m0, gv, acc[3]); acc[4] = fmaf(m1, gv, acc[4]); acc[5] = fmaf(m2, gv, acc[5]);
                acc[6] = fmaf(s0, gv, acc[6]); acc[7] = fmaf(s1, gv, acc[7]); acc[8] = fmaf(s2, gv, acc[8]);
                accb += gv;
                w0 = w1; w1 = w2; m0 = m1; m1 = m2; s0 = s1; s1 = s2;
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) red[(grp * NP + pair) * 10 + t] = acc[t];
        red[(grp * NP + pair) * 10 + 9] = accb;
    }
    __syncthreads();
    // ---- fixed-order sum over the row groups, written in the packed layout [ci][tap][co] (+ bias[co]) ----
    const int NW = cin * 9 * cout;
    float* out = a.partial + (long long)blockIdx.x * (NW + cout);
    for (int i = tid; i < NW + cout; i += 256) {
        float s = 0.f;
        if (i < NW) {
            const int c_i = i / (9 * cout), t = (i / cout) % 9, c_o = i % cout;
            for (int g = 0; g < G; ++g) s += red[(g * NP + c_i * cout + c_o) * 10 + t];
        } else {
            const int c_o = i - NW;                 // bias: the ci == 0 threads saw every pixel once
            for (int g = 0; g < G; ++g) s += red[(g * NP + c_o) * 10 + 9];
        }
        out[i] = s;
    }
}

// Register-tiled wgrad (round 2): one CTA = one 32x32 tile; a thread owns ONE input channel and EIGHT output channels, i.e. 72 weight
// accumulators (+ 8 bias sums), and walks its share of the tile's pixels with a sliding 3x3 window of the input: 3 scalar LDS + 2
// LDS.128 (the pixel's eight gradients, stored [row][col][cout]) for 72 FMAs — the first version's (ci, co)-pair threads issued 4 LDS
// per 9 FMAs and ran at ~1/8 of the FP32 pipe.  The pixels are split over G = 256 / (cin * cout/8) thread groups (rows; for the thin
// first layer also column blocks); the groups' partials are summed in a fixed order through shared memory (the staging buffers are
// reused), then written as this tile's partial in the packed layout.  Bitwise reproducible: no atomics anywhere.
constexpr int WG_ACC = 80;         // floats a thread contributes: 9 taps x 8 output channels + 8 bias sums

__global__ void __launch_bounds__(256) conv_wgrad2_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int cin = a.cin_a + a.cin_b, cout = a.cout;
    const int NO = cout >> 3;                       // output-channel octets
    const int NP = cin * NO;                        // owner slots (<= 64)
    const int G = 256 / NP;                         // pixel groups (>= 4)
    float* xs = sm;                                 // [cin][34][WIP]
    float* gs = xs + cin * (WT + 2) * WIP;          // [32][32][cout]
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
    const int x0 = tx * WT, y0 = ty * WT;
    // ---- stage the input tile with a 1-px halo (zero outside the virtual image; reflect / channel map / concat as the forward loader)
#pragma unroll 4
    for (int i = tid; i < cin * (WT + 2) * (WT + 2); i += 256) {
        const int c = i / ((WT + 2) * (WT + 2)), r = (i / (WT + 2)) % (WT + 2), col = i % (WT + 2);
        const int vy = y0 - 1 + r, vx = x0 - 1 + col;
        float v = 0.f;
        if (vy >= 0 && vy < a.H && vx >= 0 && vx < a.W) {
            if (c < a.cin_a) {
                int sy = vy - a.a_oy, sx = vx - a.a_ox;
                bool ok = true;
                if (a.a_reflect) {
                    sy = sy < 0 ? -sy : sy; sy = sy >= a.a_H ? 2 * (a.a_H - 1) - sy : sy;
                    sx = sx < 0 ? -sx : sx; sx = sx >= a.a_W ? 2 * (a.a_W - 1) - sx : sx;
                } else ok = sy >= 0 && sy < a.a_H && sx >= 0 && sx < a.a_W;
                int plane = c;
                if (a.cin_a <= 4) plane = (a.a_chmap >> (8 * c)) & 0xff;
                if (ok) v = __ldg(a.a + plane * a.a_cs + (long long)sy * a.a_rs + sx);
            } else {
                const int sy = vy - a.b_oy, sx = vx - a.b_ox;
                if (sy >= 0 && sy < a.b_H && sx >= 0 && sx < a.b_W) v = __ldg(a.b + (c - a.cin_a) * a.b_cs + (long long)sy * a.b_rs + sx);
            }
        }
        xs[(c * (WT + 2) + r) * WIP + col] = v;
    }
#pragma unroll 4
    for (int i = tid; i < cout * WT * WT; i += 256) {          // coalesced along the image row, transposed into [r][col][c]
        const int c = i / (WT * WT), r = (i / WT) % WT, col = i % WT;
        const int vy = y0 + r, vx = x0 + col;
        gs[(r * WT + col) * cout + c] = (vy < a.H && vx < a.W) ? __ldg(a.g + c * a.g_cs + (long long)vy * a.g_rs + vx) : 0.f;
    }
    __syncthreads();
    const int slot = tid % NP, grp = tid / NP;
    const int ci = slot / NO, oct = slot - ci * NO;
    float acc[9][8];
    float accb[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        accb[o] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t][o] = 0.f;
    }
    if (grp < G) {
        // pixel share of this group: whole rows while there are at most 32 groups, else one row and a column block
        int r0, r1, c0, c1;
        if (G <= WT) { const int rows = WT / G; r0 = grp * rows; r1 = r0 + rows; c0 = 0; c1 = WT; }
        else { const int cs = G / WT, cw = WT / cs; r0 = grp % WT; r1 = r0 + 1; c0 = (grp / WT) * cw; c1 = c0 + cw; }
        const float* xc = xs + ci * (WT + 2) * WIP;
        for (int r = r0; r < r1; ++r) {
            const float* x0r = xc + r * WIP + c0;     // input rows r, r+1, r+2 hold image rows y-1, y, y+1; column c0 = image column x-1
            const float* x1r = x0r + WIP;
            const float* x2r = x1r + WIP;
            const float4* gp = reinterpret_cast<const float4*>(gs + (r * WT + c0) * cout + 8 * oct);
            float w0 = x0r[0], w1 = x0r[1], m0 = x1r[0], m1 = x1r[1], s0 = x2r[0], s1 = x2r[1];
#pragma unroll 2
            for (int c = 0; c < c1 - c0; ++c) {
                const float w2 = x0r[c + 2], m2 = x1r[c + 2], s2 = x2r[c + 2];
                const float4 ga = gp[(c * cout) >> 2], gb = gp[((c * cout) >> 2) + 1];
                const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
                const float xv[9] = {w0, w1, w2, m0, m1, m2, s0, s1, s2};
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int o = 0; o < 8; ++o) acc[t][o] = fmaf(xv[t], gv[o], acc[t][o]);
#pragma unroll
                for (int o = 0; o < 8; ++o) accb[o] += gv[o];
                w0 = w1; w1 = w2; m0 = m1; m1 = m2; s0 = s1; s1 = s2;
            }
        }
    }
    __syncthreads();                                  // the staging buffers become the reduction buffer
    float* red = sm;                                  // [G][NP][WG_ACC]
    if (grp < G) {
        float* rp = red + (grp * NP + slot) * WG_ACC;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int o = 0; o < 8; ++o) rp[t * 8 + o] = acc[t][o];
#pragma unroll
        for (int o = 0; o < 8; ++o) rp[72 + o] = accb[o];
    }
    __syncthreads();
    // ---- fixed-order sum over the groups, written in the packed layout [ci][tap][co] (+ bias[co]) ----
    const int NW = cin * 9 * cout;
    float* out = a.partial + (long long)blockIdx.x * (NW + cout);
    for (int i = tid; i < NW + cout; i += 256) {
        float s = 0.f;
        if (i < NW) {
            const int c_i = i / (9 * cout), t = (i / cout) % 9, c_o = i % cout;
            const int sl = c_i * NO + (c_o >> 3), e = t * 8 + (c_o & 7);
            for (int g = 0; g < G; ++g) s += red[(g * NP + sl) * WG_ACC + e];
        } else {
            const int c_o = i - NW;                 // bias: the ci == 0 threads saw every pixel once
            const int sl = c_o >> 3, e = 72 + (c_o & 7);
            for (int g = 0; g < G; ++g) s += red[(g * NP + sl) * WG_ACC + e];
        }
        out[i] = s;
    }
}

// grad[i] = sum over parts (fixed order); two-level: each block of 256 parts is summed by one thread, then the block sums
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ partial, int nparts, int n,
                                                             float* __restrict__ grad, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;                                  // fp64 second stage: thousands of tile partials
    for (int c = 0; c < nparts; ++c) s += (double)partial[(long long)c * n + i];
    grad[i] = (accumulate ? grad[i] : 0.f) + (float)s;
}

__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ act,
                                                        const float* __restrict__ add, float* __restrict__ out,
                                                        int C, int H, int W, long long g_cs, int g_rs, long long a_cs, int a_rs,
                                                        long long d_cs, int d_rs, long long o_cs, int o_rs) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    for (int c = 0; c < C; ++c) {
        float v = g[c * g_cs + (long long)y * g_rs + x];
        if (add) v += add[c * d_cs + (long long)y * d_rs + x];
        out[c * o_cs + (long long)y * o_rs + x] = act[c * a_cs + (long long)y * a_rs + x] > 0.f ? v : 0.f;
    }
}

// out(y,x) = [act(y,x) > 0] * ( skip(y,x) + (act(y,x) is the FIRST maximum of its 2x2 window ? g_pool(y/2,x/2) : 0) )
// (ATen max_pool2d keeps the first maximum in row-major window order; rows/cols beyond 2*floor(n/2) are not pooled)
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ skip, const float* __restrict__ gpool,
                                                       const float* __restrict__ act, float* __restrict__ out, int C, int H,
                                                       int W, long long s_cs, int s_rs, long long p_cs, int p_rs,
                                                       long long a_cs, int a_rs, long long o_cs, int o_rs) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int py = y >> 1, px = x >> 1;
    const bool pooled = py < (H >> 1) && px < (W >> 1);
    for (int c = 0; c < C; ++c) {
        const float* ap = act + c * a_cs;
        const float me = ap[(long long)y * a_rs + x];
        float v = skip ? skip[c * s_cs + (long long)y * s_rs + x] : 0.f;
        if (pooled) {
            const float* w0 = ap + (long long)(2 * py) * a_rs + 2 * px;
            const float v00 = w0[0], v01 = w0[1], v10 = w0[a_rs], v11 = w0[a_rs + 1];
            int best = 0; float bv = v00;
            if (v01 > bv) { bv = v01; best = 1; }
            if (v10 > bv) { bv = v10; best = 2; }
            if (v11 > bv) { bv = v11; best = 3; }
            if (best == ((y & 1) * 2 + (x & 1))) v += gpool[c * p_cs + (long long)py * p_rs + px];
        }
        out[c * o_cs + (long long)y * o_rs + x] = me > 0.f ? v : 0.f;
    }
}

// ConvTranspose2d(k2,s2) dgrad: gin[ci](y,x) = sum_co sum_t gu[co](2y+dy, 2x+dx) * w[ci][t][co]
template <int C>
__global__ void __launch_bounds__(128) convt_dgrad_kernel(const float* __restrict__ gu, long long gu_cs, int gu_rs,
                                                          const float* __restrict__ w, float* __restrict__ gin, long long gi_cs,
                                                          int gi_rs, int Hl, int Wl) {
    __shared__ float ws[C * 4 * C];
    for (int i = threadIdx.x + 32 * threadIdx.y; i < C * 4 * C; i += 128) ws[i] = __ldg(w + i);
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= Wl || y >= Hl) return;
    float acc[C];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) acc[ci] = 0.f;
    for (int co = 0; co < C; ++co) {
        const float* gp = gu + co * gu_cs + (long long)(2 * y) * gu_rs + 2 * x;
        const float g0 = gp[0], g1 = gp[1], g2 = gp[gu_rs], g3 = gp[gu_rs + 1];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            const float* wp = ws + ci * 4 * C + co;
            acc[ci] = fmaf(g0, wp[0], fmaf(g1, wp[C], fmaf(g2, wp[2 * C], fmaf(g3, wp[3 * C], acc[ci]))));
        }
    }
#pragma unroll
    for (int ci = 0; ci < C; ++ci) gin[ci * gi_cs + (long long)y * gi_rs + x] = acc[ci];
}

// ConvTranspose2d wgrad partials over 16x16 low-res tiles: dW[ci][t][co] = sum in[ci](y,x) * gu[co](2y+dy,2x+dx); db[co] = sum gu[co]
template <int C>
__global__ void __launch_bounds__(256) convt_wgrad_kernel(const float* __restrict__ in, long long in_cs, int in_rs,
                                                          const float* __restrict__ gu, long long gu_cs, int gu_rs, int Hl,
                                                          int Wl, int tiles_x, float* __restrict__ partial) {
    extern __shared__ float sm[];
    float* xs = sm;                    // [C][256]
    float* gs = xs + C * 257;          // [C][32*32] (+pad)
    constexpr int GP = 32 * 32 + 1;
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * 16, y0 = ty * 16;
    for (int i = tid; i < C * 256; i += 256) {
        const int c = i >> 8, r = (i >> 4) & 15, col = i & 15;
        const int y = y0 + r, x = x0 + col;
        xs[c * 257 + (i & 255)] = (y < Hl && x < Wl) ? __ldg(in + c * in_cs + (long long)y * in_rs + x) : 0.f;
    }
    for (int i = tid; i < C * 1024; i += 256) {
        const int c = i >> 10, r = (i >> 5) & 31, col = i & 31;
        const int y = 2 * y0 + r, x = 2 * x0 + col;
        gs[c * GP + (i & 1023)] = (y < 2 * Hl && x < 2 * Wl) ? __ldg(gu + c * gu_cs + (long long)y * gu_rs + x) : 0.f;
    }
    __syncthreads();
    constexpr int NW = C * 4 * C;
    float* out = partial + (long long)blockIdx.x * (NW + C);
    for (int o = tid; o < NW + C; o += 256) {
        float s = 0.f;
        if (o < NW) {
            const int ci = o / (4 * C), t = (o / C) & 3, co = o % C;
            const float* xp = xs + ci * 257;
            const float* gp = gs + co * GP + (t >> 1) * 32 + (t & 1);
            for (int r = 0; r < 16; ++r)
#pragma unroll 4
                for (int c = 0; c < 16; ++c) s = fmaf(xp[r * 16 + c], gp[(2 * r) * 32 + 2 * c], s);
        } else {
            const float* gp = gs + (o - NW) * GP;
            for (int i = 0; i < 1024; ++i) s += gp[i];
        }
        out[o] = s;
    }
}

}  // namespace pc

using namespace pc;

extern "C" size_t pc_conv_wgrad_workspace_bytes(int cin, int cout, int H, int W) {
    return (size_t)cdiv(H, WT) * cdiv(W, WT) * (size_t)(cin * 9 * cout + cout) * sizeof(float) + 256;
}

extern "C" int pc_conv3x3_wgrad(const float* a, int cin_a, long long a_cs, int a_rs, int a_H, int a_W, int a_oy, int a_ox,
                                int a_reflect, unsigned a_chmap, const float* b, int cin_b, long long b_cs, int b_rs, int b_H,
                                int b_W, int b_oy, int b_ox, const float* g, long long g_cs, int g_rs, int cout, int H, int W,
                                float* grad_pack, int accumulate, void* workspace, size_t workspace_bytes, pc_stream_t stream) {
    PC_CHECK_ARG(a && g && grad_pack && workspace, "null pointer");
    const int cin = cin_a + cin_b;
    PC_CHECK_ARG(cin >= 1 && (cout == 8 || cout == 16) && cin * cout <= 256 && 256 % (cin * cout) == 0, "unsupported channel counts");
    PC_CHECK_ARG(cin_b == 0 || b, "null pointer");
    PC_CHECK_ARG(H >= 1 && W >= 1, "bad shape");
    if (workspace_bytes < pc_conv_wgrad_workspace_bytes(cin, cout, H, W)) {
        set_error("pc_conv3x3_wgrad: workspace too small");
        return PC_ERR_WORKSPACE;
    }
    WgradArgs A{};
    A.a = a; A.a_cs = a_cs; A.a_rs = a_rs; A.a_H = a_H; A.a_W = a_W; A.a_oy = a_oy; A.a_ox = a_ox; A.a_reflect = a_reflect;
    A.a_chmap = a_chmap; A.cin_a = cin_a;
    A.b = b; A.b_cs = b_cs; A.b_rs = b_rs; A.b_H = b_H; A.b_W = b_W; A.b_oy = b_oy; A.b_ox = b_ox; A.cin_b = cin_b;
    A.g = g; A.g_cs = g_cs; A.g_rs = g_rs; A.cout = cout; A.H = H; A.W = W; A.tiles_x = cdiv(W, WT);
    A.partial = reinterpret_cast<float*>(round_up((long long)(uintptr_t)workspace, 256));
    const int tiles = A.tiles_x * cdiv(H, WT);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool v1 = [] { const char* e = getenv("POPCORN_WGRAD_V1"); return e && atoi(e) != 0; }();
    if (v1) {                                          // the first version (one (ci, co) pair per thread), kept for A/B runs
        const int G = 256 / (cin * cout);
        const int smem = (cin * (WT + 2) * WIP + cout * WGP + G * cin * cout * 10) * 4;
        PC_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        static const int cat = prof_register("conv3x3_wgrad");
        ProfScope prof(cat, st, (double)H * W);
        conv_wgrad_kernel<<<tiles, 256, smem, st>>>(A);
    } else {
        const int stage = cin * (WT + 2) * WIP + cout * WT * WT;          // floats: input tile + transposed gradient tile
        const int red = 256 * WG_ACC;                                     // floats: every thread's 80 partials (G * NP == 256 slots at most)
        const int smem = (stage > red ? stage : red) * 4;
        PC_CUDA(cudaFuncSetAttribute(conv_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        static const int cat = prof_register("conv3x3_wgrad");
        ProfScope prof(cat, st, (double)H * W);
        conv_wgrad2_kernel<<<tiles, 256, smem, st>>>(A);
    }
    PC_LAUNCH_CHECK();
    const int n = cin * 9 * cout + cout;
    partial_reduce_kernel<<<cdiv(n, 256), 256, 0, st>>>(A.partial, tiles, n, grad_pack, accumulate);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_relu_backward(const float* g, long long g_cs, int g_rs, const float* act, long long a_cs, int a_rs,
                                const float* add, long long d_cs, int d_rs, float* out, long long o_cs, int o_rs, int C, int H,
                                int W, pc_stream_t stream) {
    PC_CHECK_ARG(g && act && out, "null pointer");
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    relu_mask_kernel<<<dim3(cdiv(W, 256), H), 256, 0, (cudaStream_t)stream>>>(g, act, add, out, C, H, W, g_cs, g_rs, a_cs, a_rs,
                                                                             d_cs, d_rs, o_cs, o_rs);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_maxpool2x2_relu_backward(const float* skip, long long s_cs, int s_rs, const float* gpool, long long p_cs,
                                           int p_rs, const float* act, long long a_cs, int a_rs, float* out, long long o_cs,
                                           int o_rs, int C, int H, int W, pc_stream_t stream) {
    PC_CHECK_ARG(gpool && act && out, "null pointer");
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    pool_bwd_kernel<<<dim3(cdiv(W, 256), H), 256, 0, (cudaStream_t)stream>>>(skip, gpool, act, out, C, H, W, s_cs, s_rs, p_cs,
                                                                            p_rs, a_cs, a_rs, o_cs, o_rs);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_convt2x2_dgrad(const float* gu, long long gu_cs, int gu_rs, const float* w, int C, int Hl, int Wl, float* gin,
                                 long long gi_cs, int gi_rs, pc_stream_t stream) {
    PC_CHECK_ARG(gu && w && gin, "null pointer");
    PC_CHECK_ARG(C == 8 || C == 16, "C must be 8 or 16");
    if (Hl <= 0 || Wl <= 0) return 0;
    dim3 grid(cdiv(Wl, 32), cdiv(Hl, 4)), block(32, 4);
    if (C == 8) convt_dgrad_kernel<8><<<grid, block, 0, (cudaStream_t)stream>>>(gu, gu_cs, gu_rs, w, gin, gi_cs, gi_rs, Hl, Wl);
    else convt_dgrad_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>(gu, gu_cs, gu_rs, w, gin, gi_cs, gi_rs, Hl, Wl);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t pc_convt_wgrad_workspace_bytes(int C, int Hl, int Wl) {
    return (size_t)cdiv(Hl, 16) * cdiv(Wl, 16) * (size_t)(C * 4 * C + C) * sizeof(float) + 256;
}

extern "C" int pc_convt2x2_wgrad(const float* in, long long in_cs, int in_rs, const float* gu, long long gu_cs, int gu_rs, int C,
                                 int Hl, int Wl, float* grad_pack, int accumulate, void* workspace, size_t workspace_bytes,
                                 pc_stream_t stream) {
    PC_CHECK_ARG(in && gu && grad_pack && workspace, "null pointer");
    PC_CHECK_ARG(C == 8 || C == 16, "C must be 8 or 16");
    PC_CHECK_ARG(Hl >= 1 && Wl >= 1, "bad shape");
    if (workspace_bytes < pc_convt_wgrad_workspace_bytes(C, Hl, Wl)) {
        set_error("pc_convt2x2_wgrad: workspace too small");
        return PC_ERR_WORKSPACE;
    }
    float* partial = reinterpret_cast<float*>(round_up((long long)(uintptr_t)workspace, 256));
    const int tiles_x = cdiv(Wl, 16), tiles = tiles_x * cdiv(Hl, 16);
    const int smem = (C * 257 + C * 1025) * 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 8) {
        PC_CUDA(cudaFuncSetAttribute(convt_wgrad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        convt_wgrad_kernel<8><<<tiles, 256, smem, st>>>(in, in_cs, in_rs, gu, gu_cs, gu_rs, Hl, Wl, tiles_x, partial);
    } else {
        PC_CUDA(cudaFuncSetAttribute(convt_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        convt_wgrad_kernel<16><<<tiles, 256, smem, st>>>(in, in_cs, in_rs, gu, gu_cs, gu_rs, Hl, Wl, tiles_x, partial);
    }
    PC_LAUNCH_CHECK();
    const int n = C * 4 * C + C;
    partial_reduce_kernel<<<cdiv(n, 256), 256, 0, st>>>(partial, tiles, n, grad_pack, accumulate);
    PC_LAUNCH_CHECK();
    return 0;
}
