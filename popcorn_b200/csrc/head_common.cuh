// Shared declarations of the occupancy-head kernels (SIMT fp32 and tcgen05 split-operand variants).
#pragma once
#include "common.cuh"

namespace pc {

constexpr int HM = 256;   // pixels per CTA tile
constexpr int HN = 64;    // hidden width

__host__ __device__ constexpr int head_pack_floats(int K1) { return K1 * HN + HN + 2 * (HN * HN + HN) + HN + 4; }

struct HeadArgs {
    const float* pack;
    const float* feats; long long f_bs, f_cs; int f_rs;
    const float* builtup; long long bu_bs; int bu_rs;
    int B, H, W;
    float* dens; float* scale; long long o_bs; int o_rs;
    const int32_t* ids; long long id_bs; int id_rs;
    const int32_t* census_idx;
    double* sums; int R;
    // sparse
    const int32_t* idx; const int32_t* n_dev; long long HW; float* scale_sel;
};

// warp-aggregated census partial sum: one fp64 atomic per warp when the whole warp shares a bin
__device__ __forceinline__ void bin_add(double* sums, int bin, float v) {
    int same;
    __match_all_sync(0xffffffffu, bin, &same);
    if (same) {
        const float s = warp_sum(v);
        if ((threadIdx.x & 31) == 0 && bin >= 0) atomicAdd(sums + bin, (double)s);
    } else if (bin >= 0) {
        atomicAdd(sums + bin, (double)v);
    }
}

}  // namespace pc
