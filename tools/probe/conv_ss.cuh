// Shared declarations of the round-2 candidate conv_ss.cu (3xTF32, raw fp32 row as the SS-form A operand) — see its header comment.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pc {

constexpr int PJOBS = 8;
constexpr int PBOX = 136;                      // staged pixels per row: x0-1 .. x0+134 (136 * 16 B = 17 * 128 B keeps every chunk 128-B aligned)
constexpr int PCHUNK = PBOX * 16;              // bytes of one 4-channel chunk of a staged row = LBO of the A descriptor
constexpr int PND = 8;                         // accumulator ring: output rows in flight (16 TMEM columns each)
constexpr int PBROWS = 48;                     // B rows: [W_ky2 | W_ky1 | W_ky0] x 16 output channels (Cout 8 zero-padded)
constexpr int PTHREADS = 18 * 32;
enum { PEPI_STORE = 0, PEPI_POOL = 1, PEPI_DOT = 2 };

struct SsJob {
    const uint8_t* wimg;                       // packed weights (conv_ss_pack_layer), device
    float* out_c4;                             // [COUT/4][H][W][4] fp32 chunks, or null
    float* out_planar; long long out_cs; int out_rs;   // planar fp32 output (the layer that feeds the head), or null
    float* pool_c4;                            // [COUT/4][H/2][W/2][4] fp32 chunks (PEPI_POOL)
    int a_oy, a_ox, b_oy, b_ox;                // source offsets (the Up block's zero-padded upsampled branch)
    int linear;                                // 1: no ReLU
    // PEPI_DOT (builtup copy, last layer): 1x1 out-conv slice over the 8 outputs (+ the other stream's partial logits, + bias and
    // sigmoid when dot_final), planar fp32 [crop_H][crop_W]   (model/popcorn.py:301, 317-320; networks.py:323-330)
    const float* dotw;                         // [8] weights + [1] bias
    const float* dot_in; int dot_in_rs;
    float* dot_out; int dot_out_rs; int dot_final;
};
struct alignas(64) SsParams {
    CUtensorMap tmA[PJOBS], tmB[PJOBS];
    int H, W, TR, tiles_x, tiles_y;
    int crop_y, crop_x, crop_H, crop_W;        // planar / dot outputs go to (y - crop_y, x - crop_x) if inside [0,crop_H) x [0,crop_W); crop_H = 0: no crop
    SsJob jobs[PJOBS];
};

// host API (conv_ss.cu)
int conv_ss_image_bytes(int cin, int cout);                                     // bytes of one packed layer image
void conv_ss_pack(const float* flat, int cin, int cout, uint8_t* img);          // [cin][ky][kx][cout] + bias -> image (host)
bool conv_ss_tmap(CUtensorMap* tm, const float* ptr, int cq, int H, int W);     // [cq][H][W][4] fp32 chunks -> 4-D tensor map
int conv_ss_launch(int cqa, int cqb, int cout, int epi, SsParams& p, int njobs, cudaStream_t st);

}  // namespace pc
