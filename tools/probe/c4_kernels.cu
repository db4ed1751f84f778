// ROUND-2 CANDIDATE helpers (never run): the SIMT kernels around conv_ss.cu in the 16-byte pixel-chunk layout [C/4][H][W][4] fp32.
//   * planar <-> chunk converters: let a single layer of the shipped schedule be swapped for conv_ss.cu and measured in place
//     (they cost a pass each and disappear once every layer of a stream speaks the chunk layout);
//   * ConvTranspose2d k2 s2 on chunks (model/DDA_model/utils/networks.py:302): one thread per low-res pixel reads C/4 float4 and
//     writes the 2x2 output pixels' chunks as float4 — same arithmetic order as csrc/conv.cu's convt2x2_kernel (bias first, then
//     ci ascending), so results are bit-identical to the planar kernel.
// Built into popcorn_b200/libpopcorn_b200_probe.so by tools/probe/build_conv_pair.sh.
#include "common.cuh"

namespace pc {

__global__ void __launch_bounds__(256) planar_to_c4_kernel(const float* __restrict__ in, long long cs, int rs, int CQ, int H, int W,
                                                           float4* __restrict__ out) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, q = blockIdx.z;
    if (x >= W) return;
    const float* s = in + (long long)(4 * q) * cs + (long long)y * rs + x;
    out[((size_t)q * H + y) * W + x] = make_float4(__ldg(s), __ldg(s + cs), __ldg(s + 2 * cs), __ldg(s + 3 * cs));
}

__global__ void __launch_bounds__(256) c4_to_planar_kernel(const float4* __restrict__ in, int CQ, int H, int W, float* __restrict__ out,
                                                           long long cs, int rs) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, q = blockIdx.z;
    if (x >= W) return;
    const float4 v = __ldg(in + ((size_t)q * H + y) * W + x);
    float* d = out + (long long)(4 * q) * cs + (long long)y * rs + x;
    d[0] = v.x; d[cs] = v.y; d[2 * cs] = v.z; d[3 * cs] = v.w;
}

// w: [ci][dy*2+dx][co] then bias[co]  (the SIMT pack of weights.py::_pack_convt)
template <int C>
__global__ void __launch_bounds__(128) convt2x2_c4_kernel(const float4* __restrict__ in, const float* __restrict__ w, int Hl, int Wl,
                                                          float4* __restrict__ out) {
    __shared__ __align__(16) float ws[C * 4 * C + C];
    for (int i = threadIdx.x + 32 * threadIdx.y; i < (C * 4 * C + C) / 4; i += 128)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= Wl || y >= Hl) return;
    const int Ho = 2 * Hl, Wo = 2 * Wl;
    float xin[C];
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        const float4 v = __ldg(in + ((size_t)q * Hl + y) * Wl + x);
        xin[4 * q] = v.x; xin[4 * q + 1] = v.y; xin[4 * q + 2] = v.z; xin[4 * q + 3] = v.w;
    }
#pragma unroll 1
    for (int cg = 0; cg < C; cg += 8) {                     // 8 output channels (two chunks) at a time: 32 accumulators
        float acc[4][8];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[t][o] = ws[C * 4 * C + cg + o];
#pragma unroll
        for (int ci = 0; ci < C; ++ci)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[t][o] = fmaf(xin[ci], ws[(ci * 4 + t) * C + cg + o], acc[t][o]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {                       // t = dy*2 + dx
            const size_t px = (size_t)(2 * y + (t >> 1)) * Wo + 2 * x + (t & 1);
            out[(size_t)(cg / 4) * Ho * Wo + px] = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
            out[(size_t)(cg / 4 + 1) * Ho * Wo + px] = make_float4(acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
        }
    }
}

// host launcher used by dda_c4.cu
int convt2x2_c4_launch(int C, const float* in, const float* w, int Hl, int Wl, float* out, cudaStream_t st) {
    const dim3 grid(cdiv(Wl, 32), cdiv(Hl, 4)), block(32, 4);
    if (C == 8) convt2x2_c4_kernel<8><<<grid, block, 0, st>>>(reinterpret_cast<const float4*>(in), w, Hl, Wl, reinterpret_cast<float4*>(out));
    else if (C == 16) convt2x2_c4_kernel<16><<<grid, block, 0, st>>>(reinterpret_cast<const float4*>(in), w, Hl, Wl, reinterpret_cast<float4*>(out));
    else return PC_ERR_INVALID;
    return (int)cudaGetLastError();
}

}  // namespace pc

using namespace pc;

extern "C" int pc_probe_planar_to_c4(const float* in, long long cs, int rs, int C, int H, int W, float* out, void* stream) {
    if (!in || !out || C % 4 || H < 1 || W < 1 || H > 65535) return PC_ERR_INVALID;
    planar_to_c4_kernel<<<dim3(cdiv(W, 256), H, C / 4), 256, 0, (cudaStream_t)stream>>>(in, cs, rs, C / 4, H, W, reinterpret_cast<float4*>(out));
    return (int)cudaGetLastError();
}
extern "C" int pc_probe_c4_to_planar(const float* in, int C, int H, int W, float* out, long long cs, int rs, void* stream) {
    if (!in || !out || C % 4 || H < 1 || W < 1 || H > 65535) return PC_ERR_INVALID;
    c4_to_planar_kernel<<<dim3(cdiv(W, 256), H, C / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(in), C / 4, H, W, out, cs, rs);
    return (int)cudaGetLastError();
}
// in: [C/4][Hl][Wl][4], w: device [C][4][C] + bias[C], out: [C/4][2Hl][2Wl][4]
extern "C" int pc_probe_convt2x2_c4(const float* in, const float* w, int C, int Hl, int Wl, float* out, void* stream) {
    if (!in || !w || !out || (C != 8 && C != 16) || Hl < 1 || Wl < 1) return PC_ERR_INVALID;
    return convt2x2_c4_launch(C, in, w, Hl, Wl, out, (cudaStream_t)stream);
}
