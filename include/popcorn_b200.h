/*
 * popcorn_b200 — C-ABI of the B200-native (sm_100a) POPCORN dense-prediction hot path.
 *
 * Drop-in boundary: the reference has no FFI — its "kernels" are the
 * torch.nn calls inside model/popcorn.py and model/DDA_model/utils/networks.py.  Each entry
 * point below names the reference code it replaces (paths relative to the reference root).
 * The Python host (popcorn_b200/model/popcorn.py) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers, fp32, planar NCHW unless stated; strides in elements
 *   - every call enqueues on `stream` (a cudaStream_t) and returns immediately; no internal threads,
 *     no allocation: scratch comes from the caller-provided workspace (torch's caching allocator)
 *   - return value 0 = ok, otherwise a cudaError_t value or PC_ERR_*; pc_last_error() gives the text
 *   - no exceptions cross the ABI; re-entrant per (workspace, stream)
 */
#ifndef POPCORN_B200_H
#define POPCORN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pc_stream_t; /* cudaStream_t */

#define PC_ERR_INVALID 10001   /* bad argument */
#define PC_ERR_WORKSPACE 10002 /* workspace too small */

#define PC_DDA_FEATURES 0 /* DualStreamUNet(..., return_features=True): [B,F,H,W], F = 8 per present stream */
#define PC_DDA_BUILTUP 1  /* sigmoid(logits) of create_building_score: [B,1,H,W] */

/* Library version (major*100+minor) and last error text of the calling thread. */
int pc_version(void);
const char* pc_last_error(void);
/* Number of kernels this library has launched in this process (all threads); reset != 0 zeroes it after reading.
 * bench.py reports it as "gpu_launches". */
long long pc_launch_count(int reset);
/* Optional per-kernel timing: while enabled, every launch of this library is bracketed by CUDA events on its own
 * stream.  pc_profile_enable(1) resets and starts, (0) stops; pc_profile_get(cat) synchronises the events of
 * category `cat` (0 <= cat < pc_profile_num(), named by pc_profile_name) and returns summed ms and launch count.
 * bench.py uses it for the live `roofline` numbers. */
int pc_profile_enable(int on);
int pc_profile_num(void);
const char* pc_profile_name(int cat);
int pc_profile_get(int cat, double* ms, long long* launches, double* units); /* units = pixels processed */
/* Strided window copy between a (pinned) host raster and device memory on `stream` (cudaMemcpy2DAsync);
 * kind 1 = host->device, 2 = device->host.  Replaces to_cuda_inplace (utils/utils.py:22) / the per-tile .cpu()
 * of run_eval.py:127-135 for row/column windows that are not contiguous in the raster. */
int pc_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                      int kind, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Packed weights.  The host folds BatchNorm(eval) into each conv and lays the
 * floats out in the order the kernels stage them: per stream (sar, optical) 12 layer blocks
 *   conv block  : w[cin][ky*3+kx][cout], bias[cout]
 *   convT block : w[cin][dy*2+dx][cout], bias[cout]
 * in the order inc.0 inc.3 down1.0 down1.3 down2.0 down2.3 up2.up up2.0 up2.3 up1.up up1.0 up1.3,
 * followed by fusion_out_conv (16 w + 1 b, padded to 20), sar_out_conv (8+1 -> 12), optical_out_conv
 * (8+1 -> 12).  pc_dda_pack_floats() returns the total length; pc_dda_pack_offset(stream, layer)
 * the offset of a block (layer 12 with stream 0/1/2 = fusion/sar/optical out conv).
 * Replaces: the nn.Module parameter storage of model/DDA_model/utils/networks.py:154-181.
 * --------------------------------------------------------------------------------------------- */
int pc_dda_pack_floats(void);
int pc_dda_pack_offset(int stream, int layer);

/* Tensor-core weight section (csrc/conv_tc.cu): the ten 3x3 conv layers of each stream again, as tcgen05 B
 * operands — per layer [hi|lo] matrices with rows [W_ky2 | W_ky1 | W_ky0] (3*Cout rows: row Cout*(2-ky) + co, k = kx*cin + ci,
 * K padded to 128-byte atoms) in the UMMA K-major SWIZZLE_128B layout, each weight split w = hi + lo for the three-product scheme
 * D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo — fp16 halves (64 per 128-byte row; default build) or TF32 halves (hi = top 19 bits, 32 per
 * row; build with TC_OPERANDS=tf32), see pc_tc_operand_format() — then bias[16].
 * It is appended to the fp32 pack at float offset pc_dda_tc_pack_base() (the fp32 pack rounded up to 256 B) and is
 * pc_dda_tc_pack_floats() long; pc_dda_tc_pack converts a HOST fp32 pack into a HOST image (pure host code).
 * pc_dda_forward uses the tensor-core kernels when the pack it is given is long enough to hold the section.
 * The image of the optical stream's FIRST layer holds its four input channels in memory plane order (R, G, B, NIR = network channels
 * 2, 1, 0, 3): where the source needs no reflection that layer reads its planes with one TMA box (csrc/conv.cu launch_conv). */
int pc_dda_tc_pack_base(void);
int pc_dda_tc_pack_floats(void);
int pc_dda_tc_pack(const float* flat_host, float* img_host);
int pc_conv_tc_layer_floats(int cin, int cout);
int pc_conv_tc_pack_layer(const float* flat_host, int cin, int cout, float* img_host);

/* Head pack: W1t[k=Cin][64], b1[64], W2t[64][64], b2[64], W3t[64][64], b3[64], w4[64] (row 0 of
 * head.6.weight — only channel 0 is used, model/popcorn.py:162-164), b4 (+3 pad).  Cin = 16 or 8. */
int pc_head_pack_floats(int head_in);

/* ---------------------------------------------------------------------------------------------
 * pc_dda_forward — one DualStreamUNet copy on a (reflect-padded) window.
 * Replaces: POPCORN.add_padding + channel reorder + DualStreamUNet.forward (+ fusion_out_conv +
 *           sigmoid + revert_padding)   model/popcorn.py:126-158 and 279-322,
 *           model/DDA_model/utils/networks.py:121-151, 192-237, 253-330.
 *   x          [B,C,H,W] view (C = 6: R,G,B,NIR,VV,VH | 2: VV,VH | 4: R,G,B,NIR), strides given
 *   pad_*      reflect padding applied virtually by the loader (14/14/14/14 for the builtup pass,
 *              the pad-to-64 amounts for the feature pass, 0 otherwise); nothing is materialised
 *   mode       PC_DDA_FEATURES -> out[B,F,H,W] ; PC_DDA_BUILTUP -> out[B,1,H,W]; cropped back to HxW
 *   wpack      device pack of wpack_floats floats: the fp32 section, optionally followed by the tensor-core
 *              section (then the 3x3 convs run on tcgen05 with split operands; otherwise the fp32 SIMT stencils run)
 * --------------------------------------------------------------------------------------------- */
size_t pc_dda_workspace_bytes(int B, int C, int Hv, int Wv);
int pc_dda_forward(const float* wpack, long long wpack_floats, const float* x, int B, int C, int H, int W, long long x_bstride,
                   long long x_cstride, int x_rstride, int pad_top, int pad_bottom, int pad_left, int pad_right,
                   int mode, float* out, long long out_bstride, long long out_cstride, int out_rstride,
                   void* workspace, size_t workspace_bytes, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * pc_head_dense_forward — fused occupancy head on every pixel.
 * Replaces: self.head(headin)[:,0] -> relu -> * building_counts -> region sum
 *           model/popcorn.py:164, 170, 178, 184-190 (and the per-region loop of
 *           data/PopulationDataset.py:705-712 when `ids` is a full id raster).
 *   feats [B,Cin,H,W] (strided), builtup [B,1,H,W] or NULL (occupancymodel=False: dens = relu(out))
 *   dens, scale   [B,H,W] outputs (scale may be NULL)
 *   ids           optional int32 [B,H,W] region ids; sums (double[R]) += dens where 0 <= id < R
 *   census_idx    optional int32 [B]: if given, sums has B entries and pixel p of image b contributes to
 *                 sums[b] iff ids[b,p] == census_idx[b]  (popcount of model/popcorn.py:186-187);
 *                 with ids == NULL and sums != NULL: sums[b] += all pixels of image b (:190)
 * --------------------------------------------------------------------------------------------- */
int pc_head_dense_forward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                          long long f_cstride, int f_rstride, const float* builtup, long long bu_bstride,
                          int bu_rstride, int B, int H, int W, float* dens, float* scale, long long o_bstride,
                          int o_rstride, const int32_t* ids, long long id_bstride, int id_rstride,
                          const int32_t* census_idx, double* sums, int R, pc_stream_t stream);

/* Tensor-core variants of the two head forwards (tcgen05.mma with hi/lo operand splitting — three products per MAC, activations
 * as the A operand in TMEM; csrc/head_tc.cu).  Same contract and arguments as pc_head_dense_forward /
 * pc_head_sparse_forward, except that `tcpack` is the pc_head_tc_pack_bytes()-byte weight image built by
 * popcorn_b200.weights.pack_head_tc (hi/lo split, K-major SWIZZLE_128B). */
int pc_head_tc_pack_bytes(void);
/* Operand format the library's tensor-core kernels were built for (csrc/tc_common.cuh PC_TC_F16): 0 = TF32 halves (4-byte elements,
 * 32 per 128-byte swizzle row), 1 = fp16 halves (2-byte elements, 64 per row).  Both weight images (pc_dda_tc_pack, pack_head_tc) keep
 * their sizes and section offsets; only the element format inside a matrix differs. */
int pc_tc_operand_format(void);
int pc_head_dense_forward_tc(const void* tcpack, int head_in, const float* feats, long long f_bstride,
                             long long f_cstride, int f_rstride, const float* builtup, long long bu_bstride,
                             int bu_rstride, int B, int H, int W, float* dens, float* scale, long long o_bstride,
                             int o_rstride, const int32_t* ids, long long id_bstride, int id_rstride,
                             const int32_t* census_idx, double* sums, int R, pc_stream_t stream);
int pc_head_sparse_forward_tc(const void* tcpack, int head_in, const float* feats, long long f_bstride,
                              long long f_cstride, const float* builtup, const int32_t* idx, const int32_t* n_dev,
                              long long n_max, long long HW, float* dens, float* scale_sel, double* popcount,
                              pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * pc_infer_tile_fused — the whole dense inference of one tile batch in ONE call:
 *   builtup pass (reflect 14, building_extractor, fusion logits, sigmoid) -> feature pass (unetmodel, pad-to-64 rule of
 *   add_padding) -> tcgen05 occupancy head -> relu x builtup -> optional census partial sums.
 * Replaces: POPCORN.forward in eval mode, model/popcorn.py:100-193, as run_eval.py:109 calls it (padding=False).
 *   bext_pack / unet_pack  device DDA packs WITH the tensor-core section (pc_dda_pack_floats .. + pc_dda_tc_pack), `pack_floats` long
 *   head_tcpack            pc_head_tc_pack_bytes() image
 *   x [B,6|2|4,H,W] fp32 view; dens / scale [B,H,W] (scale may be NULL); builtup [B,1,H,W] (output; contiguous)
 *   ids / census_idx / sums / R: as pc_head_dense_forward (all NULL / 0: no census sums)
 *   workspace: pc_infer_tile_workspace_bytes(B, C, H, W) = DDA activations of the larger (reflect-padded) pass + the feature map.
 * "Fused" at the boundary: one call, no host round trip and no allocation between the three launch groups; the activations still
 * travel layer by layer through the workspace (DESIGN.md §5 explains why the layers are not merged into one kernel).
 * --------------------------------------------------------------------------------------------- */
size_t pc_infer_tile_workspace_bytes(int B, int C, int H, int W);
int pc_infer_tile_fused(const float* bext_pack, const float* unet_pack, long long pack_floats, const void* head_tcpack,
                        const float* x, int B, int C, int H, int W, long long x_bstride, long long x_cstride, int x_rstride,
                        float* dens, float* scale, float* builtup, const int32_t* ids, const int32_t* census_idx,
                        double* sums, int R, void* workspace, size_t workspace_bytes, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse occupancy head (training path).
 * pc_sparse_mask_compact — get_sparsity_mask live branch + row-major compaction.
 * Replaces: model/popcorn.py:361-377 and the boolean-index gather of :214-226.
 *   builtup [B,H,W], admin [B,H,W] fp32 ids (reference dtype) , census_idx int32 [B]
 *   grid_rows uint8[H], grid_cols uint8[W]: the CPU-RNG 60x60 grid (host draws it, popcorn.py:367-368)
 *   mask_out uint8 [B,H,W]; idx_out int32 [B*H*W] (first n valid, row-major (b,h,w) order);
 *   n_out   int32 device scalar.  If the mask is empty the region mask is used (popcorn.py:374-375).
 *   workspace: pc_compact_workspace_bytes(B*H*W)
 * --------------------------------------------------------------------------------------------- */
size_t pc_compact_workspace_bytes(long long npix);
int pc_sparse_mask_compact(const float* builtup, const float* admin, const int32_t* census_idx,
                           const uint8_t* grid_rows, const uint8_t* grid_cols, int use_builtup, int B, int H, int W,
                           uint8_t* mask_out, int32_t* idx_out, int32_t* n_out, void* workspace,
                           size_t workspace_bytes, pc_stream_t stream);

/* pc_head_sparse_forward — gather + MLP + scatter (+ popcount) on the n compacted pixels.
 * Replaces: sparse_module_forward + relu + *builtup + masked sum, model/popcorn.py:195-228, 170-187.
 *   feats [B,Cin,H,W] contiguous per image (f_bstride, f_cstride); idx int32[n] flat (b*H*W + p)
 *   n_dev: device int32 (n); n_max: upper bound used to size the grid (B*H*W is always valid)
 *   dens [B*H*W] must be zero-filled by the caller (scatter target); scale_sel [n] = relu(out) row-major
 *   popcount double[B] += dens over selected pixels of image b (mask is inside the region by construction)
 */
int pc_head_sparse_forward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                           long long f_cstride, const float* builtup, const int32_t* idx, const int32_t* n_dev,
                           long long n_max, long long HW, float* dens, float* scale_sel, double* popcount,
                           pc_stream_t stream);

/* pc_head_sparse_backward — gradients of the head parameters for the census-supervised step.
 * Replaces: autograd backward of model/popcorn.py:162-187 under unet_no_grad=True (run_train.py:201-230).
 *   g_popcount float[B]  : dL/dpopcount[b]
 *   g_scale_coef         : coefficient c of the scale regulariser, dL/dscale_sel[i] += c * sign(scale_sel[i])
 *                          (utils/losses.py:74-76: lam * scale_regularization / n)
 *   g_scale_sel float[n] or NULL : additional explicit dL/dscale_sel
 *   grad_pack            : device float[pc_head_pack_floats(head_in)] gradient in hpack layout (w4/b4 slots =
 *                          row 0 of head.6; row 1 has zero gradient), overwritten
 *   workspace            : pc_head_bwd_workspace_bytes(head_in)
 *   g_feats              : NULL, or device float [B, head_in, H, W] with the strides of feats, ZEROED by the caller:
 *                          dL/dfeats is scattered to the selected pixels (needed when unetmodel is fine-tuned, N4)
 *  Deterministic: fixed grid, fixed tile->CTA assignment, two-stage fixed-order reduction, no float atomics.
 */
size_t pc_head_bwd_workspace_bytes(int head_in);
int pc_head_sparse_backward(const float* hpack, int head_in, const float* feats, long long f_bstride,
                            long long f_cstride, const float* builtup, const int32_t* idx, const int32_t* n_dev,
                            long long n_max, long long HW, const float* g_popcount, float g_scale_coef,
                            const float* g_scale_sel, float* grad_pack, void* workspace, size_t workspace_bytes,
                            float* g_feats, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Census aggregation.
 * pc_region_sum — sums[id[p]] += dens[p] for 0 <= id < R (ids int32; id < 0 or >= R ignored).
 * Replaces: the R-iteration crop/mask/sum loop of data/PopulationDataset.py:696-725 (and :842-846).
 * pc_region_sum_backward — g_dens[p] = (0 <= id[p] < R) ? g_sums[id[p]] : 0   (autograd of the above).
 * pc_region_scale — dens[p] *= factor[id[p]]  (adjust_map_to_census, data/PopulationDataset.py:846-850).
 * --------------------------------------------------------------------------------------------- */
int pc_region_sum(const float* dens, const int32_t* ids, long long npix, int R, double* sums, pc_stream_t stream);
int pc_region_sum_backward(const float* g_sums, const int32_t* ids, long long npix, int R, float* g_dens,
                           pc_stream_t stream);
int pc_region_scale(float* dens, const int32_t* ids, long long npix, int R, const float* factor, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Country-map accumulation (run_eval.py:127-154).
 * pc_accumulate_tile — map[y0+r, x0+c] += dens[r,c]; map_sq += dens^2; scale maps likewise; count += 1
 *                      for the tile's centre window rows [r0,r1) x cols [c0,c1).
 * pc_finalize_map    — where count > 1: mean = sum / count; std = sqrt((sumsq - mean^2*count)/(count-1)).
 * --------------------------------------------------------------------------------------------- */
int pc_accumulate_tile(const float* dens, const float* scale, int t_rstride, int r0, int r1, int c0, int c1,
                       float* map, float* map_sq, float* smap, float* smap_sq, int16_t* count, int m_rstride,
                       int y0, int x0, pc_stream_t stream);
int pc_finalize_map(float* map, float* map_sq, float* smap, float* smap_sq, const int16_t* count, long long npix,
                    pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * pc_ingest_normalize — raw bands (on-disk dtypes) -> the normalised fp32 input window [n_s2+n_s1, h, w].
 * Replaces: the host-side .astype(np.float32) of data/PopulationDataset.py:594-604, to_cuda_inplace
 *           (utils/utils.py:22) of 24 B/px and apply_normalize + concatenate (utils/utils.py:105-127, 162-171):
 *           out[c] = (float(band[c]) - mean[c]) / std[c], fp32, bit-identical to the reference tensor.
 *   s2            DEVICE planes, uint16 (s2_is_u16 = 1) or float32; element strides; output channel c reads plane
 *                 (s2_plane_map >> 8c) & 0xff  (file order B02,B03,B04,B08 -> R,G,B,NIR is 0x03000102)
 *   s1            DEVICE float32 planes (VV, VH)
 *   mean, stdv    HOST arrays of n_s2 + n_s1 floats (data/config/dataset_stats.json)
 * --------------------------------------------------------------------------------------------- */
int pc_ingest_normalize(const void* s2, int s2_is_u16, int n_s2, long long s2_cstride, int s2_rstride,
                        unsigned s2_plane_map, const float* s1, int n_s1, long long s1_cstride, int s1_rstride, int h,
                        int w, const float* mean, const float* stdv, float* out, long long out_cstride, int out_rstride,
                        pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * UNet fine-tuning (the reference's default for batches < 9 M px, run_train.py:191-202).
 * The training path runs the DDA UNet LAYER BY LAYER so that autograd (popcorn_b200/model/unet_train.py) can keep
 * every activation.  All planes are fp32 with explicit element strides (cs = plane, rs = row).
 *   pc_conv3x3_layer   one Conv2d 3x3 pad 1 + folded BN (+ReLU if relu) [+ fused 2x2 max-pool output]; source A
 *                      (optional reflect folding / plane map), optional concatenated source B at an offset, weights
 *                      in the packed block layout (w) and optionally as tcgen05 image (wtc).  With relu = 0 and the
 *                      transposed, tap-flipped weights this is also the DGRAD of the layer.      networks.py:258-267
 *   pc_convt2x2_layer  ConvTranspose2d k2 s2 forward                                              networks.py:302
 *   pc_conv3x3_wgrad   dW[cin][tap][cout], db[cout] of the folded conv from its input (A | B) and the ReLU-masked output
 *                      gradient g; deterministic two-stage reduction; accumulate != 0 adds to grad_pack
 *   pc_relu_backward   out = (act > 0) ? g (+ add) : 0                                            (ReLU backward)
 *   pc_maxpool2x2_relu_backward  out = (act > 0) ? skip + [act is the first max of its window] * gpool : 0
 *                      (MaxPool2d(2) backward + the skip-connection sum + ReLU backward, networks.py:289, 318)
 *   pc_convt2x2_dgrad / pc_convt2x2_wgrad   backward of the transposed conv (packed layout [cin][dy*2+dx][cout], bias)
 * --------------------------------------------------------------------------------------------- */
int pc_conv3x3_layer(const float* a, int cin_a, long long a_cs, int a_rs, int a_H, int a_W, int a_oy, int a_ox,
                     int a_reflect, unsigned a_chmap, const float* b, int cin_b, long long b_cs, int b_rs, int b_H, int b_W,
                     int b_oy, int b_ox, const float* w, const float* wtc, int cout, int relu, int H, int W, float* out,
                     long long out_cs, int out_rs, float* pool, long long pool_cs, int pool_rs, pc_stream_t stream);
int pc_convt2x2_layer(const float* in, int C, long long in_cs, int in_rs, int Hl, int Wl, const float* w, float* out,
                      long long out_cs, int out_rs, pc_stream_t stream);
size_t pc_conv_wgrad_workspace_bytes(int cin, int cout, int H, int W);
int pc_conv3x3_wgrad(const float* a, int cin_a, long long a_cs, int a_rs, int a_H, int a_W, int a_oy, int a_ox,
                     int a_reflect, unsigned a_chmap, const float* b, int cin_b, long long b_cs, int b_rs, int b_H, int b_W,
                     int b_oy, int b_ox, const float* g, long long g_cs, int g_rs, int cout, int H, int W, float* grad_pack,
                     int accumulate, void* workspace, size_t workspace_bytes, pc_stream_t stream);
int pc_relu_backward(const float* g, long long g_cs, int g_rs, const float* act, long long a_cs, int a_rs, const float* add,
                     long long d_cs, int d_rs, float* out, long long o_cs, int o_rs, int C, int H, int W, pc_stream_t stream);
int pc_maxpool2x2_relu_backward(const float* skip, long long s_cs, int s_rs, const float* gpool, long long p_cs, int p_rs,
                                const float* act, long long a_cs, int a_rs, float* out, long long o_cs, int o_rs, int C,
                                int H, int W, pc_stream_t stream);
int pc_convt2x2_dgrad(const float* gu, long long gu_cs, int gu_rs, const float* w, int C, int Hl, int Wl, float* gin,
                      long long gi_cs, int gi_rs, pc_stream_t stream);
size_t pc_convt_wgrad_workspace_bytes(int C, int Hl, int Wl);
int pc_convt2x2_wgrad(const float* in, long long in_cs, int in_rs, const float* gu, long long gu_cs, int gu_rs, int C, int Hl,
                      int Wl, float* grad_pack, int accumulate, void* workspace, size_t workspace_bytes, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Unit-test hooks (tests/test_gpu_kernels.py): ONE fused conv3x3(+folded BN)+ReLU layer, optionally with a
 * second concatenated source placed at an offset (Up block) and a fused 2x2 max-pool output, and ONE
 * ConvTranspose2d(k2,s2) layer, on plain contiguous tensors.  w uses the packed block layouts above.
 * wtc != NULL (a device copy of the layer's pc_conv_tc_pack_layer image) selects the tensor-core kernel.
 * Same kernels pc_dda_forward schedules; replaces networks.py:258-267 / :289 / :302 one layer at a time.
 * --------------------------------------------------------------------------------------------- */
int pc_test_conv3x3(const float* a, int cin_a, int a_H, int a_W, int a_oy, int a_ox, int a_reflect, const float* b,
                    int cin_b, int b_H, int b_W, int b_oy, int b_ox, const float* w, int cout, int H, int W,
                    float* out, float* pool, const float* wtc, pc_stream_t stream);
int pc_test_convt2x2(const float* in, int C, int Hl, int Wl, const float* w, float* out, pc_stream_t stream);
/* FP32 SIMT ceiling probe: every thread runs 32*iters dependent-chain FMAs (scalar, or packed fma.rn.f32x2);
 * returns the number of threads launched (negative = error is impossible: errors are > 0 codes <= 10002, thread
 * counts are >= 37888).  Measured denominator for the stencil kernels' FP32 roofline. */
int pc_test_fma_peak(int use_x2, int iters, int blocks_per_sm, float* out, pc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POPCORN_B200_H */
