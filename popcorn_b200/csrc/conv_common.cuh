// Declarations shared by the SIMT stencil kernels (conv.cu) and the tcgen05 implicit-GEMM kernels (conv_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pc {

constexpr int MAX_JOBS = 8;
constexpr int TC_BOXW = 136;         // floats per staged row: box starts at x0-4 (TMA needs a 16-byte aligned inner coordinate), 128 px + halo

enum { EPI_STORE = 0, EPI_POOL = 1, EPI_DOT = 2, EPI_CONVT = 3 };

// One (image, stream) instance of a conv layer.  The input is the channel concatenation of source A (first CIN_A
// channels; optional reflect folding and plane remap for the first layer) and source B (next CIN_B channels, placed
// at an offset and zero outside: the zero-padded upsampled branch of an Up block).
struct ConvJob {
    const float* a; long long a_cs; int a_rs; int a_H, a_W; int a_oy, a_ox; int a_reflect; unsigned a_chmap;
    const float* b; long long b_cs; int b_rs; int b_H, b_W; int b_oy, b_ox;
    const float* w;                      // SIMT: [CIN][9][COUT] then bias[COUT]
    const float* wtc;                    // tcgen05: pre-split, pre-swizzled image of the layer (conv_tc.cu), or null
    float* out; long long out_cs; int out_rs; int out_vec;
    int linear;                          // 1: no ReLU in the epilogue (the dgrad convolutions of the UNet backward)
    float* pool; long long pool_cs; int pool_rs;
    const float* dotw;                   // [8] weights + [1] bias of the 1x1 out conv slice (EPI_DOT)
    const float* dot_in; int dot_in_rs;  // partial logits of the other stream (or null)
    float* dot_out; int dot_out_rs; int dot_final;
    // EPI_CONVT (tensor-core kernel only): the Up block's ConvTranspose2d(k=2, s=2) applied to this layer's activated output
    // in the epilogue; ctw = [COUT][4 taps dy*2+dx][COUT] then bias[COUT] (the SIMT pack of the transposed conv),
    // ct_out = [COUT][2H][2W] planes.  The layer's own output is not stored unless `out` is set.
    const float* ctw; float* ct_out; long long ct_cs; int ct_rs;
};

// geometry of one launch of the tensor-core conv (all jobs share it)
struct alignas(64) TcConvParams {
    CUtensorMap tmA[MAX_JOBS];           // TMA descriptors of the jobs' sources: box = [channels][1 row][TC_BOXW floats]
    CUtensorMap tmB[MAX_JOBS];
    int H, W;                            // virtual image == output extent
    int crop_y, crop_x, crop_H, crop_W;  // stores go to (y-crop_y, x-crop_x) if inside [0,crop_H)x[0,crop_W)
    int TR, tiles_x, tiles_y;            // tile = 128 columns x TR rows
    ConvJob jobs[MAX_JOBS];
};

// conv.cu: [C][H][W] fp32 planes (row stride rs, plane stride cs, in floats) -> 3-D tensor map with a [boxc][boxh][boxw] box;
// false when the source does not meet TMA's alignment rules (16-byte base and strides)
bool make_tmap3d(CUtensorMap* tm, const float* ptr, int C, int H, int W, int rs, long long cs, int boxw, int boxh, int boxc);

// conv_tc.cu
int conv_tc_layer_floats(int cin, int cout);                                    // floats of one layer image
void conv_tc_pack_layer(const float* flat, int cin, int cout, float* img);       // host: [cin][9][cout]+bias -> image
bool conv_tc_enabled();
int launch_conv_tc(int cin_a, int cin_b, int cout, int epi, TcConvParams& p, int njobs, cudaStream_t st);

}  // namespace pc
