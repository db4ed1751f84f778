// Fused occupancy head on sm_100a (SIMT fp32 path): per-pixel MLP 16->64->64->64->1 with the
// activations held in shared memory, ReLU, x builtup, store / scatter, census partial sums.
// Replaces model/popcorn.py:79-88 (head), :160-190 (relu, x building_counts, region sum) and
// :195-228 (sparse_module_forward gather / scatter).
#include "head_common.cuh"

namespace pc {

// one dense layer on the CTA's tile: act[k][m] (k < K) -> act[n][m] (n < 64), in place.
// thread (lane, warp): pixels {4*lane..+3} and {128+4*lane..+3}, outputs n = 8*warp..+7.
template <int K>
__device__ __forceinline__ void mlp_layer(float* act, const float* Wt, const float* bias, int lane, int warp) {
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float bj = bias[8 * warp + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = bj;
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(act + k * HM + 4 * lane);
        const float4 xb = *reinterpret_cast<const float4*>(act + k * HM + 128 + 4 * lane);
        const float4 wa = *reinterpret_cast<const float4*>(Wt + k * HN + 8 * warp);
        const float4 wb = *reinterpret_cast<const float4*>(Wt + k * HN + 8 * warp + 4);
        const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
    __syncthreads();  // every thread has finished reading the previous activations
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float* row = act + (8 * warp + j) * HM;
        *reinterpret_cast<float4*>(row + 4 * lane) =
            make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
        *reinterpret_cast<float4*>(row + 128 + 4 * lane) =
            make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f), fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
    }
    __syncthreads();
}

template <int K1, bool SPARSE>
__global__ void __launch_bounds__(256, 2) head_forward_kernel(const __grid_constant__ HeadArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                 // [64][HM]
    float* wp = smem + HN * HM;        // packed weights
    constexpr int PF = head_pack_floats(K1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < PF / 4; i += 256)
        reinterpret_cast<float4*>(wp)[i] = __ldg(reinterpret_cast<const float4*>(a.pack) + i);
    const float* W1t = wp;
    const float* b1 = W1t + K1 * HN;
    const float* W2t = b1 + HN;
    const float* b2 = W2t + HN * HN;
    const float* W3t = b2 + HN;
    const float* b3 = W3t + HN * HN;
    const float* w4 = b3 + HN;
    const float* b4 = w4 + HN;

    const long long HW = SPARSE ? a.HW : (long long)a.H * a.W;
    const long long total = SPARSE ? (long long)__ldg(a.n_dev) : HW * a.B;

    for (long long base = (long long)blockIdx.x * HM; base < total; base += (long long)gridDim.x * HM) {
        // ---- gather this thread's pixel ----
        const long long i = base + tid;
        const bool valid = i < total;
        long long p = 0; int b = 0; long long foff = 0, boff = 0, ooff = 0, ioff = 0;
        if (valid) {
            p = SPARSE ? (long long)__ldg(a.idx + i) : i;
            b = (int)(p / HW);
            const long long q = p - (long long)b * HW;
            if (SPARSE) {
                foff = b * a.f_bs + q; boff = p; ooff = p;
            } else {
                const int y = (int)(q / a.W), x = (int)(q - (long long)y * a.W);
                foff = b * a.f_bs + (long long)y * a.f_rs + x;
                boff = b * a.bu_bs + (long long)y * a.bu_rs + x;
                ooff = b * a.o_bs + (long long)y * a.o_rs + x;
                ioff = b * a.id_bs + (long long)y * a.id_rs + x;
            }
        }
#pragma unroll
        for (int c = 0; c < K1; ++c) act[c * HM + tid] = valid ? __ldg(a.feats + foff + c * a.f_cs) : 0.f;
        __syncthreads();
        mlp_layer<K1>(act, W1t, b1, lane, warp);
        mlp_layer<HN>(act, W2t, b2, lane, warp);
        mlp_layer<HN>(act, W3t, b3, lane, warp);
        // ---- last layer (row 0 of head.6) + ReLU + x builtup ----
        float o = b4[0];
#pragma unroll 16
        for (int k = 0; k < HN; ++k) o = fmaf(act[k * HM + tid], w4[k], o);
        const float s = fmaxf(o, 0.f);
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
        __syncthreads();  // act is re-used by the next tile
    }
}

template <int K1, bool SPARSE>
static int launch_head(const HeadArgs& a, long long total_bound, cudaStream_t st) {
    constexpr int smem = (HN * HM + head_pack_floats(K1)) * 4;
    auto k = head_forward_kernel<K1, SPARSE>;
    PC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long tiles = (total_bound + HM - 1) / HM;
    const int maxg = num_sms() * 2 * 8;
    int grid = (int)(tiles < maxg ? tiles : maxg);
    if (grid < 1) grid = 1;
    {
        static const int cat = prof_register(SPARSE ? "head_forward_simt<sparse>" : "head_forward_simt<dense>");
        ProfScope prof(cat, st, (double)total_bound);
        k<<<grid, 256, smem, st>>>(a);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_head_pack_floats(int head_in) { return (head_in == 8 || head_in == 16) ? head_pack_floats(head_in) : -1; }

extern "C" int pc_head_dense_forward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                                     long long f_cstride, int f_rstride, const float* builtup, long long bu_bstride,
                                     int bu_rstride, int B, int H, int W, float* dens, float* scale, long long o_bstride,
                                     int o_rstride, const int32_t* ids, long long id_bstride, int id_rstride,
                                     const int32_t* census_idx, double* sums, int R, pc_stream_t stream) {
    PC_CHECK_ARG(hpack && feats && dens, "null pointer");
    PC_CHECK_ARG(head_in == 16 || head_in == 8, "head_in must be 16 or 8");
    PC_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "bad shape");
    PC_CHECK_ARG(!(ids && !census_idx && sums) || R >= 1, "R must be >= 1 with an id raster");
    HeadArgs a{};
    a.pack = hpack; a.feats = feats; a.f_bs = f_bstride; a.f_cs = f_cstride; a.f_rs = f_rstride;
    a.builtup = builtup; a.bu_bs = bu_bstride; a.bu_rs = bu_rstride; a.B = B; a.H = H; a.W = W;
    a.dens = dens; a.scale = scale; a.o_bs = o_bstride; a.o_rs = o_rstride;
    a.ids = ids; a.id_bs = id_bstride; a.id_rs = id_rstride; a.census_idx = census_idx; a.sums = sums; a.R = R;
    const long long total = (long long)B * H * W;
    return head_in == 16 ? launch_head<16, false>(a, total, (cudaStream_t)stream)
                         : launch_head<8, false>(a, total, (cudaStream_t)stream);
}

extern "C" int pc_head_sparse_forward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                                      long long f_cstride, const float* builtup, const int32_t* idx,
                                      const int32_t* n_dev, long long n_max, long long HW, float* dens,
                                      float* scale_sel, double* popcount, pc_stream_t stream) {
    PC_CHECK_ARG(hpack && feats && idx && n_dev && dens, "null pointer");
    PC_CHECK_ARG(head_in == 16 || head_in == 8, "head_in must be 16 or 8");
    PC_CHECK_ARG(HW >= 1 && n_max >= 0, "bad shape");
    HeadArgs a{};
    a.pack = hpack; a.feats = feats; a.f_bs = f_bstride; a.f_cs = f_cstride; a.builtup = builtup;
    a.dens = dens; a.scale_sel = scale_sel; a.sums = popcount; a.idx = idx; a.n_dev = n_dev; a.HW = HW;
    a.B = 1; a.H = 1; a.W = 1;
    return head_in == 16 ? launch_head<16, true>(a, n_max, (cudaStream_t)stream)
                         : launch_head<8, true>(a, n_max, (cudaStream_t)stream);
}
