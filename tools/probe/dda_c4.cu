// ROUND-2 CANDIDATE (never run): one DDA feature pass (both streams, pads = 0: the country engine's case) in the 16-byte pixel-chunk
// layout, on conv_ss.cu + c4_kernels.cu — the A/B partner of pc_dda_forward(mode = PC_DDA_FEATURES) for the first round-2 timing.
//
// Layer schedule = csrc/conv.cu's (model/DDA_model/utils/networks.py:121-151), every intermediate in [C/4][H][W][4] fp32:
//   L0  inc.conv0   planar input (2 | 4 planes, channel-remapped) -> F0(8)        first_layer_c4_kernel (plain fp32 stencil)
//   L1  inc.conv3   F0 -> F1(8) + pooled HA(8)                                     conv_ss <2,0,8,pool>
//   L2  down1.conv0 HA -> HB(16)                                                   conv_ss <2,0,16>
//   L3  down1.conv3 HB -> HC(16) + pooled QA(16)                                   conv_ss <4,0,16,pool>
//   L4/5 down2      QA -> QB -> QA                                                 conv_ss <4,0,16> x2
//   L6  up2.up      QA -> HD(16)                                                   convt2x2_c4<16>
//   L7  up2.conv0   cat[HC, pad(HD)] -> HA(8)                                      conv_ss <4,4,8>
//   L8  up2.conv3   HA -> HB(8)                                                    conv_ss <2,0,8>
//   L9  up1.up      HB -> F2(8)                                                    convt2x2_c4<8>
//   L10 up1.conv0   cat[F1, pad(F2)] -> F0(8)                                      conv_ss <2,2,8>
//   L11 up1.conv3   F0 -> planar features, channels 8*stream ..                    conv_ss <2,0,8> (planar output)
// Both streams run as two jobs of one launch.  The first layer here is a plain one-thread-per-pixel fp32 stencil (no reflection,
// no staging): it is HBM/L2-bound either way, and is only here so that the pass is self-contained.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_ss.cuh"

namespace pc {

// c4_kernels.cu
int convt2x2_c4_launch(int C, const float* in, const float* w, int Hl, int Wl, float* out, cudaStream_t st);

// planar x [C_in planes selected by chmap] -> F0 [2][H][W][4]: conv3x3 pad 1 + bias + ReLU, CIN 2 | 4, 8 outputs
template <int CIN>
__global__ void __launch_bounds__(256) first_layer_c4_kernel(const float* __restrict__ x, long long cs, int rs, unsigned chmap, const float* __restrict__ w,
                                                             int H, int W, float4* __restrict__ out) {
    __shared__ float ws[CIN * 9 * 8 + 8];
    for (int i = threadIdx.x; i < CIN * 9 * 8 + 8; i += 256) ws[i] = __ldg(w + i);
    __syncthreads();
    const int xo = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (xo >= W) return;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = ws[CIN * 9 * 8 + o];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* plane = x + (long long)((chmap >> (8 * ci)) & 0xFF) * cs;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = xo + kx - 1;
                const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(plane + (long long)yy * rs + xx) : 0.f;
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(v, ws[((ci * 3 + ky) * 3 + kx) * 8 + o], acc[o]);
            }
        }
    }
    const size_t px = (size_t)y * W + xo;
    out[px] = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
    out[(size_t)H * W + px] = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
}

// the flat fp32 pack of one DDA copy (weights.py::_pack_dda_fp32 == csrc/conv.cu pack_offset): per stream 12 layers
struct LSpec { int cin, cout, is_t; };
static const LSpec kL[12] = {{-1, 8, 0}, {8, 8, 0}, {8, 16, 0}, {16, 16, 0}, {16, 16, 0}, {16, 16, 0},
                             {16, 16, 1}, {32, 8, 0}, {8, 8, 0}, {8, 8, 1}, {16, 8, 0}, {8, 8, 0}};
static int l_floats(int s, int l) {
    const int cin = kL[l].cin < 0 ? (s == 0 ? 2 : 4) : kL[l].cin;
    return cin * (kL[l].is_t ? 4 : 9) * kL[l].cout + kL[l].cout;
}
static int l_offset(int s, int l) {
    int off = 0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 12; ++b) {
            if (a == s && b == l) return off;
            off += l_floats(a, b);
        }
    return off;
}

}  // namespace pc

using namespace pc;

#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (int)_e; } while (0)

// wpack_host: the fp32 section of weights.pack_dda(...) on the HOST (29 804 floats); x: device planar [6][H][W] normalised, reference
// channel order [R,G,B,NIR,VV,VH]; out: device planar [16][H][W]; H, W multiples of 4.  Runs the pass `iters` times and returns
// the average ms in *ms_out (weights are packed and uploaded once, outside the timed region).
extern "C" int pc_probe_dda_features_c4(const float* wpack_host, const float* x, int H, int W, float* out, int iters, float* ms_out, void* stream) {
    if (!wpack_host || !x || !out || H < 4 || W < 4 || (H & 3) || (W & 3) || H > 65535) return PC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
    // ---- weights: conv images for layers 1-5, 7, 8, 10, 11; flat packs for layers 0, 6, 9 ----
    std::vector<uint8_t> himg;
    size_t img_off[2][12] = {};
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < 12; ++l) {
            img_off[s][l] = himg.size();
            if (l == 0 || kL[l].is_t) {
                const size_t n = (size_t)l_floats(s, l) * 4, pad = (n + 255) / 256 * 256;
                himg.resize(himg.size() + pad, 0);
                memcpy(himg.data() + img_off[s][l], wpack_host + l_offset(s, l), n);
            } else {
                const size_t n = (size_t)conv_ss_image_bytes(kL[l].cin, kL[l].cout), pad = (n + 255) / 256 * 256;
                himg.resize(himg.size() + pad, 0);
                conv_ss_pack(wpack_host + l_offset(s, l), kL[l].cin, kL[l].cout, himg.data() + img_off[s][l]);
            }
        }
    uint8_t* dimg = nullptr;
    CK(cudaMalloc(&dimg, himg.size()));
    CK(cudaMemcpy(dimg, himg.data(), himg.size(), cudaMemcpyHostToDevice));
    // ---- activations: per stream F0, F1, F2 (8 ch, full), HA (8), HB, HC, HD (16, half), QA, QB (16, quarter) ----
    const size_t full8 = (size_t)8 * H * W, half8 = (size_t)8 * H2 * W2, half16 = 2 * half8, q16 = (size_t)16 * H4 * W4;
    const size_t per_stream = 3 * full8 + half8 + 3 * half16 + 2 * q16;
    float* ws = nullptr;
    CK(cudaMalloc(&ws, 2 * per_stream * sizeof(float)));
    struct Bufs { float *F0, *F1, *F2, *HA, *HB, *HC, *HD, *QA, *QB; } B[2];
    for (int s = 0; s < 2; ++s) {
        float* p = ws + s * per_stream;
        B[s].F0 = p; p += full8; B[s].F1 = p; p += full8; B[s].F2 = p; p += full8;
        B[s].HA = p; p += half8; B[s].HB = p; p += half16; B[s].HC = p; p += half16; B[s].HD = p; p += half16;
        B[s].QA = p; p += q16; B[s].QB = p;
    }
    auto img = [&](int s, int l) { return dimg + img_off[s][l]; };
    SsParams p;
    auto conv = [&](int l, int cqa, int cqb, int cout, int epi, int Hh, int Ww, float* Bufs::*a, int aH, int aW, float* Bufs::*b, int bH, int bW,
                    float* Bufs::*o, float* Bufs::*pool, bool planar_out) -> int {
        memset(&p, 0, sizeof(p));
        p.H = Hh; p.W = Ww;
        for (int s = 0; s < 2; ++s) {
            SsJob& J = p.jobs[s];
            J.wimg = img(s, l);
            if (planar_out) { J.out_planar = out + (size_t)8 * s * H * W; J.out_cs = (long long)H * W; J.out_rs = W; }
            else J.out_c4 = B[s].*o;
            if (pool) J.pool_c4 = B[s].*pool;
            if (!conv_ss_tmap(&p.tmA[s], B[s].*a, cqa, aH, aW)) return PC_ERR_INVALID;
            if (cqb) {
                if (!conv_ss_tmap(&p.tmB[s], B[s].*b, cqb, bH, bW)) return PC_ERR_INVALID;
                J.b_oy = (Hh - bH) / 2; J.b_ox = (Ww - bW) / 2;            // F.pad split of the Up block (networks.py:309-312)
            }
        }
        return conv_ss_launch(cqa, cqb, cout, epi, p, 2, st);
    };
    auto pass = [&]() -> int {
        int rc;
        for (int s = 0; s < 2; ++s) {                                      // L0: sar = planes (VV,VH) = 4,5 ; optical = (B,G,R,NIR) = 2,1,0,3
            const dim3 grid(cdiv(W, 256), H);
            if (s == 0) first_layer_c4_kernel<2><<<grid, 256, 0, st>>>(x, (long long)H * W, W, 0x00000504u, reinterpret_cast<const float*>(img(0, 0)), H, W, reinterpret_cast<float4*>(B[0].F0));
            else first_layer_c4_kernel<4><<<grid, 256, 0, st>>>(x, (long long)H * W, W, 0x03000102u, reinterpret_cast<const float*>(img(1, 0)), H, W, reinterpret_cast<float4*>(B[1].F0));
        }
        if ((rc = conv(1, 2, 0, 8, PEPI_POOL, H, W, &Bufs::F0, H, W, nullptr, 0, 0, &Bufs::F1, &Bufs::HA, false))) return rc;
        if ((rc = conv(2, 2, 0, 16, PEPI_STORE, H2, W2, &Bufs::HA, H2, W2, nullptr, 0, 0, &Bufs::HB, nullptr, false))) return rc;
        if ((rc = conv(3, 4, 0, 16, PEPI_POOL, H2, W2, &Bufs::HB, H2, W2, nullptr, 0, 0, &Bufs::HC, &Bufs::QA, false))) return rc;
        if ((rc = conv(4, 4, 0, 16, PEPI_STORE, H4, W4, &Bufs::QA, H4, W4, nullptr, 0, 0, &Bufs::QB, nullptr, false))) return rc;
        if ((rc = conv(5, 4, 0, 16, PEPI_STORE, H4, W4, &Bufs::QB, H4, W4, nullptr, 0, 0, &Bufs::QA, nullptr, false))) return rc;
        for (int s = 0; s < 2; ++s)
            if ((rc = convt2x2_c4_launch(16, B[s].QA, reinterpret_cast<const float*>(img(s, 6)), H4, W4, B[s].HD, st))) return rc;
        if ((rc = conv(7, 4, 4, 8, PEPI_STORE, H2, W2, &Bufs::HC, H2, W2, &Bufs::HD, 2 * H4, 2 * W4, &Bufs::HA, nullptr, false))) return rc;
        if ((rc = conv(8, 2, 0, 8, PEPI_STORE, H2, W2, &Bufs::HA, H2, W2, nullptr, 0, 0, &Bufs::HB, nullptr, false))) return rc;
        for (int s = 0; s < 2; ++s)
            if ((rc = convt2x2_c4_launch(8, B[s].HB, reinterpret_cast<const float*>(img(s, 9)), H2, W2, B[s].F2, st))) return rc;
        if ((rc = conv(10, 2, 2, 8, PEPI_STORE, H, W, &Bufs::F1, H, W, &Bufs::F2, 2 * H2, 2 * W2, &Bufs::F0, nullptr, false))) return rc;
        if ((rc = conv(11, 2, 0, 8, PEPI_STORE, H, W, &Bufs::F0, H, W, nullptr, 0, 0, nullptr, nullptr, true))) return rc;
        return (int)cudaGetLastError();
    };
    int rc = pass();                                                       // warm-up (also the correctness run)
    if (!rc) rc = (int)cudaStreamSynchronize(st);
    float ms = 0.f;
    if (!rc && iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int it = 0; it < iters && !rc; ++it) rc = pass();
        cudaEventRecord(e1, st);
        if (!rc) rc = (int)cudaEventSynchronize(e1);
        if (!rc) cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        ms /= iters;
    }
    if (ms_out) *ms_out = ms;
    cudaFree(ws);
    cudaFree(dimg);
    return rc;
}
