// Census aggregation, sparsity-mask compaction and country-map accumulation kernels (HBM-bound).
// Replaces data/PopulationDataset.py:696-725 (per-region crop/mask/sum loop), :842-850 (dasymetric
// rescale), model/popcorn.py:361-377 (sparsity mask) + :214-226 (boolean-index compaction) and
// run_eval.py:127-154 (centre-masked accumulation, mean / std finalisation).
#include "common.cuh"

namespace pc {

// ---------------------------------------------------------------------------------------------------
// region segment-sum.  Each warp owns a contiguous span of pixels and walks it 128 px at a time
// (float4 + int4 per lane, coalesced 512 B).  While the whole warp sees a single id the lanes only
// accumulate privately; the warp flushes ONE fp64 atomic when the id changes (census regions are
// spatially coherent, so runs are thousands of pixels long).  Mixed 128-px groups fall back to
// per-lane run-length merging + atomics.   Algorithmic traffic: 8 B / pixel.
// ---------------------------------------------------------------------------------------------------
// With R <= SMEM_BINS the flushes go to per-CTA fp32 bins in shared memory and each CTA adds only the bins it
// touched to the global fp64 sums at the end: neighbouring warps walk the same region, so direct fp64 atomics
// on ~R hot addresses serialise in L2 (ncu r1: 54 % long-scoreboard + 43 % MIO stalls at 0.3-1.1 TB/s).
constexpr int SMEM_BINS = 8192;

template <bool SMEM>
__device__ __forceinline__ void bin_flush(float* bins, double* sums, int id, float v) {
    if (SMEM) atomicAdd(bins + id, v);
    else atomicAdd(sums + id, (double)v);
}

template <bool SMEM>
__device__ __forceinline__ void region_sum_body(const float* __restrict__ dens, const int32_t* __restrict__ ids,
                                                long long npix, int R, double* __restrict__ sums, long long span,
                                                float* bins) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long beg = warp * span;
    long long end = beg + span < npix ? beg + span : npix;
    if (beg >= npix) return;
    int cur = -1;        // warp-uniform id of the open run
    float run = 0.f;     // this lane's private partial of the open run
    const bool vec_ok = ((((uintptr_t)dens) | ((uintptr_t)ids)) & 15) == 0;
    constexpr int U = 4;   // independent 128-px groups in flight per warp (8 x 16 B loads per lane)
    for (long long g0 = beg; g0 < end; g0 += 128 * U) {
        float v[U][4]; int id[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = g0 + 128 * u + 4 * lane;
            if (vec_ok && p + 3 < end) {
                const float4 fv = ld_stream4(dens + p);
                const int4 iv = ld_stream4i(ids + p);
                v[u][0] = fv.x; v[u][1] = fv.y; v[u][2] = fv.z; v[u][3] = fv.w;
                id[u][0] = iv.x; id[u][1] = iv.y; id[u][2] = iv.z; id[u][3] = iv.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = p + k < end;
                    v[u][k] = ok ? dens[p + k] : 0.f;
                    id[u][k] = ok ? ids[p + k] : -1;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (g0 + 128 * u >= end) break;   // warp-uniform
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (id[u][k] < 0 || id[u][k] >= R) { id[u][k] = -1; v[u][k] = 0.f; }
            const bool mine_uniform = (id[u][0] == id[u][1]) && (id[u][1] == id[u][2]) && (id[u][2] == id[u][3]);
            int all_same;
            __match_all_sync(0xffffffffu, mine_uniform ? id[u][0] : -2 - lane, &all_same);
            if (all_same) {
                if (id[u][0] != cur) {           // warp-uniform branch
                    const float s = warp_sum(run);
                    if (lane == 0 && cur >= 0) bin_flush<SMEM>(bins, sums, cur, s);
                    cur = id[u][0];
                    run = 0.f;
                }
                run += (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]);
            } else {
                // mixed group: merge equal neighbours privately, then one atomic per private run
                float s = v[u][0]; int c = id[u][0];
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    if (id[u][k] == c) s += v[u][k];
                    else {
                        if (c >= 0) bin_flush<SMEM>(bins, sums, c, s);
                        c = id[u][k]; s = v[u][k];
                    }
                }
                if (c >= 0) bin_flush<SMEM>(bins, sums, c, s);
            }
        }
    }
    const float s = warp_sum(run);
    if (lane == 0 && cur >= 0) bin_flush<SMEM>(bins, sums, cur, s);
}

template <bool SMEM>
__global__ void __launch_bounds__(256) region_sum_kernel(const float* __restrict__ dens, const int32_t* __restrict__ ids,
                                                         long long npix, int R, double* __restrict__ sums,
                                                         long long span) {
    extern __shared__ float bins[];
    if (SMEM) {
        for (int i = threadIdx.x; i < R; i += 256) bins[i] = 0.f;
        __syncthreads();
    }
    region_sum_body<SMEM>(dens, ids, npix, R, sums, span, bins);
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < R; i += 256) {
            const float v = bins[i];
            if (v != 0.f) atomicAdd(sums + i, (double)v);
        }
    }
}

__global__ void __launch_bounds__(256) region_gather_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                            long long npix, int R, float* __restrict__ data, int multiply) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const int id = ids[p];
        const bool ok = id >= 0 && id < R;
        if (multiply) { if (ok) data[p] *= __ldg(table + id); }
        else data[p] = ok ? __ldg(table + id) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------
// sparsity mask + compaction (three passes, deterministic row-major order)
// ---------------------------------------------------------------------------------------------------
constexpr int CBLOCK = 1024;  // pixels per compaction block (256 threads x 4)

struct CompactArgs {
    const float* builtup; const float* admin; const int32_t* census_idx;
    const uint8_t* grid_rows; const uint8_t* grid_cols; int use_builtup;
    int B, H, W; long long npix;
    uint8_t* mask; int32_t* idx; int32_t* n_out;
    int32_t* counts;   // [2][nblocks]: selected, region
    int32_t* offsets;  // [nblocks]
    int32_t* flags;    // [0] = use region fallback
};

__device__ __forceinline__ void mask_eval(const CompactArgs& a, long long p, bool& sel, bool& region) {
    const long long HW = (long long)a.H * a.W;
    const int b = (int)(p / HW);
    const long long q = p - (long long)b * HW;
    const int y = (int)(q / a.W), x = (int)(q - (long long)y * a.W);
    region = a.admin[p] == (float)a.census_idx[b];                                  // popcorn.py:362 / 364
    const bool built = a.use_builtup ? (a.builtup[p] > 0.f) : true;
    const bool grid = a.grid_rows[y] && a.grid_cols[x];                             // popcorn.py:369
    sel = region && (built || grid);                                                // popcorn.py:372
}

__global__ void __launch_bounds__(256) compact_count_kernel(const __grid_constant__ CompactArgs a, int nblocks) {
    __shared__ int s_sel[8], s_reg[8];
    const long long p0 = (long long)blockIdx.x * CBLOCK + 4 * threadIdx.x;
    int csel = 0, creg = 0;
    unsigned m4 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long p = p0 + k;
        bool sel = false, region = false;
        if (p < a.npix) mask_eval(a, p, sel, region);
        csel += sel; creg += region;
        m4 |= (unsigned)(sel ? 1 : 0) << (8 * k) | (unsigned)(region ? 2 : 0) << (8 * k);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (p0 + k < a.npix) a.mask[p0 + k] = (m4 >> (8 * k)) & 0xff;   // bit0 = selected, bit1 = region (fixed up in pass 3)
    csel = __reduce_add_sync(0xffffffffu, csel);
    creg = __reduce_add_sync(0xffffffffu, creg);
    if ((threadIdx.x & 31) == 0) { s_sel[threadIdx.x >> 5] = csel; s_reg[threadIdx.x >> 5] = creg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ts = 0, tr = 0;
        for (int w = 0; w < 8; ++w) { ts += s_sel[w]; tr += s_reg[w]; }
        a.counts[blockIdx.x] = ts;
        a.counts[nblocks + blockIdx.x] = tr;
    }
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(const __grid_constant__ CompactArgs a, int nblocks) {
    __shared__ int s_part[32];
    __shared__ int s_total_sel;
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // total of the selected mask first (decides the fallback, popcorn.py:374-375)
    int t = 0;
    for (int i = tid; i < nblocks; i += 1024) t += a.counts[i];
    t = __reduce_add_sync(0xffffffffu, t);
    if (lane == 0) s_part[warp] = t;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < 32; ++w) s += s_part[w];
        s_total_sel = s;
        s_carry = 0;
        a.flags[0] = (s == 0) ? 1 : 0;
    }
    __syncthreads();
    const int32_t* cnt = a.counts + (s_total_sel == 0 ? nblocks : 0);
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + tid;
        const int v = i < nblocks ? cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        __syncthreads();
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_part[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            s_part[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_part[warp - 1] : 0) + inc - v;
        if (i < nblocks) a.offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_part[31];
        __syncthreads();
    }
    if (tid == 0) a.n_out[0] = s_carry;
}

__global__ void __launch_bounds__(256) compact_write_kernel(const __grid_constant__ CompactArgs a) {
    __shared__ int s_warp[8];
    const int bit = a.flags[0] ? 2 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = (long long)blockIdx.x * CBLOCK + 4 * threadIdx.x;
    bool m[4]; int c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        m[k] = (p0 + k < a.npix) && (a.mask[p0 + k] & bit);
        c += m[k];
    }
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    int pos = a.offsets[blockIdx.x] + wbase + inc - c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (p0 + k < a.npix) a.mask[p0 + k] = m[k] ? 1 : 0;
        if (m[k]) a.idx[pos++] = (int32_t)(p0 + k);
    }
}

// ---------------------------------------------------------------------------------------------------
// country-map accumulation
// ---------------------------------------------------------------------------------------------------
// One thread = 4 consecutive pixels of a row (VEC: 16-byte loads / stores of the tile and the maps, 8-byte of the int16 counts; the host
// checks the alignment) or 1 pixel.  Per element the arithmetic is the scalar one: separately rounded mul / add (no FMA contraction) —
// the reference squares, then adds (run_eval.py:111,128).
template <bool VEC>
__global__ void __launch_bounds__(256) accumulate_kernel(const float* __restrict__ dens, const float* __restrict__ scale,
                                                         int t_rs, int r0, int r1, int c0, int c1, float* map,
                                                         float* map_sq, float* smap, float* smap_sq, int16_t* count,
                                                         int m_rs, int y0, int x0) {
    constexpr int PX = VEC ? 4 : 1;
    const int c = c0 + (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const int r = r0 + blockIdx.y;
    if (c >= c1 || r >= r1) return;
    const long long t = (long long)r * t_rs + c;
    const long long m = (long long)(y0 + r) * m_rs + (x0 + c);
    if (VEC && c + 3 < c1) {
        const float4 d = *reinterpret_cast<const float4*>(dens + t);
        float4 v = *reinterpret_cast<float4*>(map + m);
        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
        *reinterpret_cast<float4*>(map + m) = v;
        if (map_sq) {
            float4 q = *reinterpret_cast<float4*>(map_sq + m);
            q.x = __fadd_rn(q.x, __fmul_rn(d.x, d.x)); q.y = __fadd_rn(q.y, __fmul_rn(d.y, d.y));
            q.z = __fadd_rn(q.z, __fmul_rn(d.z, d.z)); q.w = __fadd_rn(q.w, __fmul_rn(d.w, d.w));
            *reinterpret_cast<float4*>(map_sq + m) = q;
        }
        if (scale) {
            const float4 sc = *reinterpret_cast<const float4*>(scale + t);
            if (smap) {
                float4 u = *reinterpret_cast<float4*>(smap + m);
                u.x += sc.x; u.y += sc.y; u.z += sc.z; u.w += sc.w;
                *reinterpret_cast<float4*>(smap + m) = u;
            }
            if (smap_sq) {
                float4 q = *reinterpret_cast<float4*>(smap_sq + m);
                q.x = __fadd_rn(q.x, __fmul_rn(sc.x, sc.x)); q.y = __fadd_rn(q.y, __fmul_rn(sc.y, sc.y));
                q.z = __fadd_rn(q.z, __fmul_rn(sc.z, sc.z)); q.w = __fadd_rn(q.w, __fmul_rn(sc.w, sc.w));
                *reinterpret_cast<float4*>(smap_sq + m) = q;
            }
        }
        if (count) {
            short4 n = *reinterpret_cast<short4*>(count + m);
            n.x += 1; n.y += 1; n.z += 1; n.w += 1;
            *reinterpret_cast<short4*>(count + m) = n;
        }
        return;
    }
    for (int k = 0; k < PX && c + k < c1; ++k) {
        const float d = dens[t + k];
        map[m + k] += d;
        if (map_sq) map_sq[m + k] = __fadd_rn(map_sq[m + k], __fmul_rn(d, d));
        if (scale) {
            const float sc = scale[t + k];
            if (smap) smap[m + k] += sc;
            if (smap_sq) smap_sq[m + k] = __fadd_rn(smap_sq[m + k], __fmul_rn(sc, sc));
        }
        if (count) count[m + k] += 1;
    }
}

// finalize one pixel: op-by-op IEEE arithmetic in the reference's order (run_eval.py:143-154): the variance formula cancels
// catastrophically where tile results nearly coincide, so FMA contraction would change NaN / 0 outcomes
__device__ __forceinline__ void finalize_pixel(float* map, float* map_sq, float* smap, float* smap_sq, long long p, int n) {
    const float nf = (float)n;
    const float mean = __fdiv_rn(map[p], nf);
    map[p] = mean;
    if (map_sq)
        map_sq[p] = __fsqrt_rn(__fdiv_rn(__fsub_rn(map_sq[p], __fmul_rn(__fmul_rn(mean, mean), nf)), nf - 1.f));
    if (smap) {
        const float sm = __fdiv_rn(smap[p], nf);
        smap[p] = sm;
        if (smap_sq)
            smap_sq[p] = __fsqrt_rn(__fdiv_rn(__fsub_rn(smap_sq[p], __fmul_rn(__fmul_rn(sm, sm), nf)), nf - 1.f));
    }
}

// The visit counts are read 8 at a time (16 bytes; VEC: `count` 16-byte aligned): almost every pixel has count <= 1 and needs nothing else
// (run_eval.py:140, div_mask = count > 1), so the kernel is a 2 B/px scan with rare read-modify-writes.
template <bool VEC>
__global__ void __launch_bounds__(256) finalize_kernel(float* map, float* map_sq, float* smap, float* smap_sq,
                                                       const int16_t* __restrict__ count, long long npix) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        const long long nvec = npix / 8;
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
            const int4 w = __ldg(reinterpret_cast<const int4*>(count) + v);
            const int words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int n = (int)(short)((words[k >> 1] >> (16 * (k & 1))) & 0xffff);
                if (n > 1) finalize_pixel(map, map_sq, smap, smap_sq, 8 * v + k, n);
            }
        }
        for (long long p = nvec * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
            const int n = count[p];
            if (n > 1) finalize_pixel(map, map_sq, smap, smap_sq, p, n);
        }
    } else {
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
            const int n = count[p];
            if (n > 1) finalize_pixel(map, map_sq, smap, smap_sq, p, n);
        }
    }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_region_sum(const float* dens, const int32_t* ids, long long npix, int R, double* sums,
                             pc_stream_t stream) {
    PC_CHECK_ARG(npix >= 0 && R >= 1, "bad shape");
    if (npix == 0) return 0;
    PC_CHECK_ARG(dens && ids && sums, "null pointer");
    // ~16 warps per SM resident x 4 waves of spans; each span a multiple of 128 px
    const long long want_warps = (long long)num_sms() * 16 * 4;
    long long span = round_up(cdiv(npix, want_warps), 512);
    if (span < 2048) span = 2048;
    const long long nwarps = cdiv(npix, span);
    const int grid = cdiv(nwarps * 32, 256);
    static const int cat = prof_register("region_sum");
    ProfScope prof(cat, (cudaStream_t)stream, (double)npix);
    if (R <= SMEM_BINS)
        region_sum_kernel<true><<<grid, 256, R * sizeof(float), (cudaStream_t)stream>>>(dens, ids, npix, R, sums, span);
    else
        region_sum_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dens, ids, npix, R, sums, span);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_region_sum_backward(const float* g_sums, const int32_t* ids, long long npix, int R, float* g_dens,
                                      pc_stream_t stream) {
    PC_CHECK_ARG(npix >= 0 && R >= 1, "bad shape");
    if (npix == 0) return 0;
    PC_CHECK_ARG(g_sums && ids && g_dens, "null pointer");
    const int grid = (int)(cdiv(npix, 256) < num_sms() * 16 ? cdiv(npix, 256) : num_sms() * 16);
    region_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g_sums, ids, npix, R, g_dens, 0);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_region_scale(float* dens, const int32_t* ids, long long npix, int R, const float* factor,
                               pc_stream_t stream) {
    PC_CHECK_ARG(npix >= 0 && R >= 1, "bad shape");
    if (npix == 0) return 0;
    PC_CHECK_ARG(dens && ids && factor, "null pointer");
    const int grid = (int)(cdiv(npix, 256) < num_sms() * 16 ? cdiv(npix, 256) : num_sms() * 16);
    region_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(factor, ids, npix, R, dens, 1);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t pc_compact_workspace_bytes(long long npix) {
    const long long nb = cdiv(npix > 0 ? npix : 1, CBLOCK);
    return (size_t)(3 * nb + 8) * sizeof(int32_t) + 256;
}

extern "C" int pc_sparse_mask_compact(const float* builtup, const float* admin, const int32_t* census_idx,
                                      const uint8_t* grid_rows, const uint8_t* grid_cols, int use_builtup, int B, int H,
                                      int W, uint8_t* mask_out, int32_t* idx_out, int32_t* n_out, void* workspace,
                                      size_t workspace_bytes, pc_stream_t stream) {
    PC_CHECK_ARG(admin && census_idx && grid_rows && grid_cols && mask_out && idx_out && n_out && workspace, "null pointer");
    PC_CHECK_ARG(!use_builtup || builtup, "builtup required");
    PC_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "bad shape");
    const long long npix = (long long)B * H * W;
    PC_CHECK_ARG(npix < (1ll << 31), "more than 2^31 pixels in one batch");
    if (workspace_bytes < pc_compact_workspace_bytes(npix)) {
        set_error("pc_sparse_mask_compact: workspace too small");
        return PC_ERR_WORKSPACE;
    }
    const int nb = cdiv(npix, CBLOCK);
    CompactArgs a{};
    a.builtup = builtup; a.admin = admin; a.census_idx = census_idx; a.grid_rows = grid_rows; a.grid_cols = grid_cols;
    a.use_builtup = use_builtup; a.B = B; a.H = H; a.W = W; a.npix = npix;
    a.mask = mask_out; a.idx = idx_out; a.n_out = n_out;
    int32_t* w = reinterpret_cast<int32_t*>(round_up((long long)(uintptr_t)workspace, 256));
    a.counts = w; a.offsets = w + 2 * nb; a.flags = w + 3 * nb;
    cudaStream_t st = (cudaStream_t)stream;
    static const int cat = prof_register("sparse_mask_compact");
    ProfScope prof(cat, st, (double)npix);
    compact_count_kernel<<<nb, 256, 0, st>>>(a, nb);
    PC_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(a, nb);
    PC_LAUNCH_CHECK();
    compact_write_kernel<<<nb, 256, 0, st>>>(a);
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_accumulate_tile(const float* dens, const float* scale, int t_rstride, int r0, int r1, int c0, int c1,
                                  float* map, float* map_sq, float* smap, float* smap_sq, int16_t* count, int m_rstride,
                                  int y0, int x0, pc_stream_t stream) {
    PC_CHECK_ARG(dens && map, "null pointer");
    if (r1 <= r0 || c1 <= c0) return 0;
    static const int cat = prof_register("accumulate");
    ProfScope prof(cat, (cudaStream_t)stream, (double)(r1 - r0) * (c1 - c0));
    auto al = [](const void* p, int b) { return p == nullptr || (((uintptr_t)p) % b) == 0; };
    const bool vec = (t_rstride % 4 == 0) && (m_rstride % 4 == 0) && (c0 % 4 == 0) && (x0 % 4 == 0) && al(dens, 16) && al(scale, 16) &&
                     al(map, 16) && al(map_sq, 16) && al(smap, 16) && al(smap_sq, 16) && al(count, 8);
    if (vec) {
        dim3 grid(cdiv(cdiv(c1 - c0, 4), 256), r1 - r0);
        accumulate_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dens, scale, t_rstride, r0, r1, c0, c1, map, map_sq, smap,
                                                                      smap_sq, count, m_rstride, y0, x0);
    } else {
        dim3 grid(cdiv(c1 - c0, 256), r1 - r0);
        accumulate_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dens, scale, t_rstride, r0, r1, c0, c1, map, map_sq, smap,
                                                                       smap_sq, count, m_rstride, y0, x0);
    }
    PC_LAUNCH_CHECK();
    return 0;
}

extern "C" int pc_finalize_map(float* map, float* map_sq, float* smap, float* smap_sq, const int16_t* count,
                               long long npix, pc_stream_t stream) {
    PC_CHECK_ARG(map && count, "null pointer");
    if (npix <= 0) return 0;
    static const int cat = prof_register("finalize");
    ProfScope prof(cat, (cudaStream_t)stream, (double)npix);
    const bool vec = (((uintptr_t)count) % 16) == 0;
    const long long work = vec ? cdiv(npix, 8) : npix;
    const int grid = (int)(cdiv(work, 256) < num_sms() * 16 ? cdiv(work, 256) : num_sms() * 16);
    if (vec) finalize_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(map, map_sq, smap, smap_sq, count, npix);
    else finalize_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(map, map_sq, smap, smap_sq, count, npix);
    PC_LAUNCH_CHECK();
    return 0;
}
