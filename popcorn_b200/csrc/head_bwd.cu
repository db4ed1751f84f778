// Census-supervised training step, occupancy-head backward on sm_100a (SIMT fp32).
// One persistent CTA per SM walks 128-pixel tiles of the compacted pixel list: it RECOMPUTES the
// forward activations in shared memory (nothing is saved by the forward), then runs dgrad and wgrad
// against them, accumulating the weight gradients in registers across all of its tiles.  Per-CTA
// partials are reduced in a fixed order by a second kernel -> bitwise reproducible, no float atomics.
// Replaces: autograd backward of model/popcorn.py:162-187 with unet_no_grad=True (run_train.py:201-230).
#include "common.cuh"

namespace pc {

constexpr int BM = 128;    // pixels per tile
constexpr int BP = 132;    // row pitch of the [row][pixel] shared buffers (conflict-free LDS.128 for wgrad)
constexpr int BN = 64;

__host__ __device__ constexpr int bwd_pack_floats(int K1) { return K1 * BN + BN + 2 * (BN * BN + BN) + BN + 4; }

struct HeadBwdArgs {
    const float* pack; const float* feats; long long f_bs, f_cs; const float* builtup;
    const int32_t* idx; const int32_t* n_dev; long long HW;
    const float* g_pop; float g_coef; const float* g_sel;
    float* partial;   // [gridDim.x][PF]
    float* g_feats;   // optional dL/dfeats, same strides as feats (zeroed by the caller)
};

// dst[n][m] = epilogue( sum_k src[k][m] * Wk[k][n] ),  n < 64, m < 128; thread: 4 pixels x 8 outputs.
//  FWD : dst = relu(acc + bias[n])           DGRAD : dst = (dst_old > 0) ? acc : 0
template <int K, bool FWD>
__device__ __forceinline__ void tile_gemm(const float* src, float* dst, const float* Wk, const float* bias, int lane, int warp) {
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float bj = FWD ? bias[8 * warp + j] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = bj;
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(src + k * BP + 4 * lane);
        const float4 wa = *reinterpret_cast<const float4*>(Wk + k * BN + 8 * warp);
        const float4 wb = *reinterpret_cast<const float4*>(Wk + k * BN + 8 * warp + 4);
        const float x[4] = {xa.x, xa.y, xa.z, xa.w};
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
    if (!FWD) __syncthreads();  // dst still holds activations other threads read in the preceding wgrad
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4* d = reinterpret_cast<float4*>(dst + (8 * warp + j) * BP + 4 * lane);
        if (FWD) {
            *d = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
        } else {
            const float4 h = *d;
            *d = make_float4(h.x > 0.f ? acc[0][j] : 0.f, h.y > 0.f ? acc[1][j] : 0.f, h.z > 0.f ? acc[2][j] : 0.f,
                             h.w > 0.f ? acc[3][j] : 0.f);
        }
    }
    __syncthreads();
}

// gw[i][j] += sum_m dz[a+16i][m] * h[b+16j][m]  (i<4, j<NJ);  gb[i] += sum_m dz[a+16i][m] on the b==0 threads
template <int NJ>
__device__ __forceinline__ void tile_wgrad(const float* dz, const float* h, float (&gw)[4][4], float (&gb)[4], int a, int b) {
#pragma unroll 2
    for (int m = 0; m < BM; m += 4) {
        float4 d[4], x[NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = *reinterpret_cast<const float4*>(dz + (a + 16 * i) * BP + m);
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[j] = *reinterpret_cast<const float4*>(h + (b + 16 * j) * BP + m);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                gw[i][j] = fmaf(d[i].x, x[j].x, gw[i][j]);
                gw[i][j] = fmaf(d[i].y, x[j].y, gw[i][j]);
                gw[i][j] = fmaf(d[i].z, x[j].z, gw[i][j]);
                gw[i][j] = fmaf(d[i].w, x[j].w, gw[i][j]);
            }
            if (b == 0) gb[i] += (d[i].x + d[i].y) + (d[i].z + d[i].w);
        }
    }
}

template <int K1>
__global__ void __launch_bounds__(256, 1) head_backward_kernel(const __grid_constant__ HeadBwdArgs a) {
    constexpr int PF = bwd_pack_floats(K1);
    extern __shared__ __align__(16) float smem[];
    float* A0 = smem;                    // h0 [K1][BP]
    float* A1 = A0 + K1 * BP;            // h1 -> dz1
    float* A2 = A1 + BN * BP;            // h2 -> dz2
    float* A3 = A2 + BN * BP;            // h3 -> dz3
    float* dO = A3 + BN * BP;            // [BM]
    float* wp = dO + BM;                 // packed weights (forward layout)
    float* W2n = wp + PF;                // W2[n][k]  (dgrad operand)
    float* W3n = W2n + BN * BN;          // W3[n][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < PF / 4; i += 256)
        reinterpret_cast<float4*>(wp)[i] = __ldg(reinterpret_cast<const float4*>(a.pack) + i);
    const float* W1t = wp;
    const float* b1 = W1t + K1 * BN;
    const float* W2t = b1 + BN;
    const float* b2 = W2t + BN * BN;
    const float* W3t = b2 + BN;
    const float* b3 = W3t + BN * BN;
    const float* w4 = b3 + BN;
    const float* b4 = w4 + BN;
    __syncthreads();
    for (int i = tid; i < BN * BN; i += 256) {   // W[n][k] = Wt[k][n]
        const int n = i >> 6, k = i & 63;
        W2n[i] = W2t[k * BN + n];
        W3n[i] = W3t[k * BN + n];
    }
    __syncthreads();

    const int wa = tid >> 4, wb = tid & 15;      // wgrad ownership: rows {wa+16i} x cols {wb+16j}
    float g3[4][4] = {}, g2[4][4] = {}, g1[4][4] = {};
    float gb3[4] = {}, gb2[4] = {}, gb1[4] = {};
    float g4 = 0.f, gb4 = 0.f;                    // lane j<8 of warp w owns dw4[8w+j]; thread 0 owns db4

    const long long total = (long long)__ldg(a.n_dev);
    for (long long base = (long long)blockIdx.x * BM; base < total; base += (long long)gridDim.x * BM) {
        // ---- gather: thread (m = tid & 127) loads channel half (tid >> 7) of pixel m ----
        {
            const int m = tid & 127, half = tid >> 7;
            const long long i = base + m;
            const bool valid = i < total;
            long long foff = 0;
            if (valid) {
                const long long p = (long long)__ldg(a.idx + i);
                const int b = (int)(p / a.HW);
                foff = b * a.f_bs + (p - (long long)b * a.HW);
            }
#pragma unroll
            for (int c = 0; c < K1 / 2; ++c) {
                const int ch = half * (K1 / 2) + c;
                A0[ch * BP + m] = valid ? __ldg(a.feats + foff + ch * a.f_cs) : 0.f;
            }
        }
        __syncthreads();
        tile_gemm<K1, true>(A0, A1, W1t, b1, lane, warp);
        tile_gemm<BN, true>(A1, A2, W2t, b2, lane, warp);
        tile_gemm<BN, true>(A2, A3, W3t, b3, lane, warp);
        // ---- output scalar and dL/do per pixel ----
        if (tid < BM) {
            const long long i = base + tid;
            float g = 0.f;
            if (i < total) {
                float o = b4[0];
#pragma unroll 16
                for (int k = 0; k < BN; ++k) o = fmaf(A3[k * BP + tid], w4[k], o);
                if (o > 0.f) {
                    const long long p = (long long)__ldg(a.idx + i);
                    const int b = (int)(p / a.HW);
                    const float bu = a.builtup ? __ldg(a.builtup + p) : 1.f;
                    g = __ldg(a.g_pop + b) * bu + a.g_coef + (a.g_sel ? __ldg(a.g_sel + i) : 0.f);
                }
            }
            dO[tid] = g;
        }
        __syncthreads();
        // ---- dw4 / db4, then dz3 = do * w4 * [h3 > 0] in place ----
        {
            const float4 dv = *reinterpret_cast<const float4*>(dO + 4 * lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 hv = *reinterpret_cast<const float4*>(A3 + (8 * warp + j) * BP + 4 * lane);
                float part = (hv.x * dv.x + hv.y * dv.y) + (hv.z * dv.z + hv.w * dv.w);
                part = warp_sum(part);
                if (lane == j) g4 += part;
            }
            if (warp == 0) {
                const float s = warp_sum((dv.x + dv.y) + (dv.z + dv.w));
                if (lane == 0) gb4 += s;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float wj = w4[8 * warp + j];
                float4* hp = reinterpret_cast<float4*>(A3 + (8 * warp + j) * BP + 4 * lane);
                const float4 hv = *hp;
                *hp = make_float4(hv.x > 0.f ? dv.x * wj : 0.f, hv.y > 0.f ? dv.y * wj : 0.f,
                                  hv.z > 0.f ? dv.z * wj : 0.f, hv.w > 0.f ? dv.w * wj : 0.f);
            }
        }
        __syncthreads();
        tile_wgrad<4>(A3, A2, g3, gb3, wa, wb);                 // dW3 += dz3 (x) h2
        tile_gemm<BN, false>(A3, A2, W3n, nullptr, lane, warp); // dz2 = (W3^T dz3) * [h2 > 0]   (syncs)
        tile_wgrad<4>(A2, A1, g2, gb2, wa, wb);                 // dW2 += dz2 (x) h1
        tile_gemm<BN, false>(A2, A1, W2n, nullptr, lane, warp); // dz1 = (W2^T dz2) * [h1 > 0]
        if (wb < K1) tile_wgrad<1>(A1, A0, g1, gb1, wa, wb);    // dW1 += dz1 (x) h0
        if (a.g_feats) {                                         // dfeats[ch][m] = sum_n W1t[ch][n] * dz1[n][m], scattered
            const int m = tid & 127, half = tid >> 7;
            const long long i = base + m;
            if (i < total) {
                const long long p = (long long)__ldg(a.idx + i);
                const int b = (int)(p / a.HW);
                const long long foff = b * a.f_bs + (p - (long long)b * a.HW);
                float acc[K1 / 2];
#pragma unroll
                for (int c = 0; c < K1 / 2; ++c) acc[c] = 0.f;
#pragma unroll 4
                for (int n = 0; n < BN; ++n) {
                    const float dz = A1[n * BP + m];
#pragma unroll
                    for (int c = 0; c < K1 / 2; ++c) acc[c] = fmaf(W1t[(half * (K1 / 2) + c) * BN + n], dz, acc[c]);
                }
#pragma unroll
                for (int c = 0; c < K1 / 2; ++c) a.g_feats[foff + (half * (K1 / 2) + c) * a.f_cs] = acc[c];
            }
        }
        __syncthreads();
    }

    // ---- write this CTA's partial gradient in pack layout (Wt[k][n] <- dW[n][k]) ----
    float* out = a.partial + (long long)blockIdx.x * PF;
    float* o_W1t = out;
    float* o_b1 = o_W1t + K1 * BN;
    float* o_W2t = o_b1 + BN;
    float* o_b2 = o_W2t + BN * BN;
    float* o_W3t = o_b2 + BN;
    float* o_b3 = o_W3t + BN * BN;
    float* o_w4 = o_b3 + BN;
    float* o_b4 = o_w4 + BN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = wa + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = wb + 16 * j;
            o_W3t[k * BN + n] = g3[i][j];
            o_W2t[k * BN + n] = g2[i][j];
        }
        if (wb < K1) o_W1t[wb * BN + n] = g1[i][0];
        if (wb == 0) { o_b3[n] = gb3[i]; o_b2[n] = gb2[i]; o_b1[n] = gb1[i]; }
    }
    if (lane < 8) o_w4[8 * warp + lane] = g4;
    if (tid == 0) { o_b4[0] = gb4; o_b4[1] = 0.f; o_b4[2] = 0.f; o_b4[3] = 0.f; }
}

__global__ void __launch_bounds__(256) head_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int PF,
                                                              float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PF) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(long long)c * PF + i];   // fixed order
    grad[i] = s;
}

static int bwd_grid(long long n_max) {
    const long long tiles = (n_max + BM - 1) / BM;
    const int sms = num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < sms ? tiles : sms));
}

}  // namespace pc

using namespace pc;

extern "C" size_t pc_head_bwd_workspace_bytes(int head_in) {
    if (head_in != 8 && head_in != 16) return 0;
    return (size_t)num_sms() * bwd_pack_floats(head_in) * sizeof(float) + 256;
}

extern "C" int pc_head_sparse_backward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                                       long long f_cstride, const float* builtup, const int32_t* idx,
                                       const int32_t* n_dev, long long n_max, long long HW, const float* g_popcount,
                                       float g_scale_coef, const float* g_scale_sel, float* grad_pack, void* workspace,
                                       size_t workspace_bytes, float* g_feats, pc_stream_t stream) {
    PC_CHECK_ARG(hpack && feats && idx && n_dev && g_popcount && grad_pack && workspace, "null pointer");
    PC_CHECK_ARG(head_in == 16 || head_in == 8, "head_in must be 16 or 8");
    PC_CHECK_ARG(HW >= 1 && n_max >= 0, "bad shape");
    if (workspace_bytes < pc_head_bwd_workspace_bytes(head_in)) {
        set_error("pc_head_sparse_backward: workspace too small");
        return PC_ERR_WORKSPACE;
    }
    HeadBwdArgs a{};
    a.pack = hpack; a.feats = feats; a.f_bs = f_bstride; a.f_cs = f_cstride; a.builtup = builtup;
    a.idx = idx; a.n_dev = n_dev; a.HW = HW; a.g_pop = g_popcount; a.g_coef = g_scale_coef; a.g_sel = g_scale_sel;
    a.partial = reinterpret_cast<float*>(round_up((long long)(uintptr_t)workspace, 256));
    a.g_feats = g_feats;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = bwd_grid(n_max);
    const int PF = bwd_pack_floats(head_in);
    static const int cat = prof_register("head_backward");
    ProfScope prof(cat, st, (double)n_max);
    if (head_in == 16) {
        constexpr int smem = ((16 + 3 * BN) * BP + BM + bwd_pack_floats(16) + 2 * BN * BN) * 4;
        auto k = head_backward_kernel<16>;
        PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<grid, 256, smem, st>>>(a);
    } else {
        constexpr int smem = ((8 + 3 * BN) * BP + BM + bwd_pack_floats(8) + 2 * BN * BN) * 4;
        auto k = head_backward_kernel<8>;
        PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<grid, 256, smem, st>>>(a);
    }
    PC_LAUNCH_CHECK();
    head_bwd_reduce_kernel<<<cdiv(PF, 256), 256, 0, st>>>(a.partial, grid, PF, grad_pack);
    PC_LAUNCH_CHECK();
    return 0;
}
