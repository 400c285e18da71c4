/* waveformer_b200 - C ABI of the B200 (sm_100a) kernels behind the WaveFormer 3D-segmentation hot path.
 *
 * The reference has NO native boundary on this path: it is pure Python that calls ptwt / torch.  The
 * entry points below are what a maintainer binds (ctypes stub in INTEGRATION.md) at the Python call sites named on
 * each function.  Conventions:
 *   - plain device pointers + sizes; no torch / C++ types; every call is asynchronous on `stream` (a cudaStream_t
 *     passed as void*; NULL = the legacy default stream);
 *   - returns 0 on success, a negative wf_status otherwise (wf_error_string gives the text); nothing aborts;
 *   - dtype: WF_F32 or WF_BF16 is the storage type of activations; all arithmetic accumulates in fp32;
 *   - no global state except a lazily created per-device attribute cache (max dynamic shared memory opt-in).
 */
#ifndef WAVEFORMER_B200_H
#define WAVEFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    WF_OK = 0,
    WF_ERR_BAD_DTYPE = -1,     /* dtype (or the dtype combination) is not one the entry point is built for */
    WF_ERR_BAD_SHAPE = -2,     /* odd extent, zero size, head_dim unsupported, window does not tile the grid ... */
    WF_ERR_NULL_POINTER = -3,
    WF_ERR_MISALIGNED = -4,    /* a pointer / stride breaks the 16-byte alignment the vector path needs */
    WF_ERR_CUDA = -5,          /* a CUDA runtime call failed; wf_last_cuda_error() has the code */
    WF_ERR_WORKSPACE = -6,     /* workspace too small; query with the *_workspace_bytes function */
    WF_ERR_UNSUPPORTED = -7
} wf_status;

/* Storage / operand formats.  WF_BF16 and WF_F16 are interchangeable 16-bit formats everywhere a 16-bit activation, weight
 * or tensor-core operand appears (same tcgen05 rate; fp16 carries 3 more mantissa bits and is the 16-bit precision policy's
 * default: every such tensor on this path is normalised or one GEMM away from a normalisation, so its range suffices). */
typedef enum { WF_F32 = 0, WF_BF16 = 1, WF_F16 = 2 } wf_dtype;

const char *wf_version(void);
const char *wf_error_string(int status);
int wf_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel group 1: 3D Haar analysis, one level.
 * Replaces ptwt.wavedec3(x, 'db1', level=1, mode='zero') as called by WaveletTransform3D.forward
 * (reference network_models/wave_helper.py:349-353) from Block.multi_scale_forward (wave_helper.py:484-486),
 * including the two permute().contiguous() copies around it when the channels-last entry point is used.
 *
 * Sub-band k (0..6) = aad, ada, add, daa, dad, dda, ddd (letter order D,H,W; a = low, d = high), each shaped like
 * the LL output; band k starts at hf + k * hf_band_stride elements.  hf == NULL skips the seven detail writes (the
 * caller discards them: only the last block of a stage keeps its details, reference waveformer.py:287-288).
 * Extents D, H, W must be even.
 * ---------------------------------------------------------------------------------------------------------- */

/* x: [n, D, H, W] contiguous (n = every leading dim folded, as ptwt does); ll and each band: [n, D/2, H/2, W/2]. */
int wf_dwt3d_ncdhw(const void *x, void *ll, void *hf, int dtype, int64_t n, int D, int H, int W,
                   int64_t hf_band_stride, void *stream);

/* x: [B, D, H, W, C], channel stride 1, voxel stride x_vox_stride elements (>= C; lets the input be a channel
 * slice of a wider buffer); ll: [B, D/2, H/2, W/2, C] with voxel stride ll_vox_stride; bands: voxel stride C. */
int wf_dwt3d_ndhwc(const void *x, void *ll, void *hf, int dtype, int hf_dtype, int B, int D, int H, int W, int C,
                   int64_t x_vox_stride, int64_t ll_vox_stride, int64_t hf_band_stride, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel group 3: 3D Haar synthesis, one level, fused with the optional high-frequency gate and with the write into
 * a wider (concatenation) buffer.
 * Replaces ptwt.waverec3((ll,) + details, 'db1') in UnetrIDWTBlock.forward (reference
 * network_models/idwt_upsample.py:159-160), the per-sub-band gate multiply of HFRefinementRes.forward
 * (idwt_upsample.py:49, `x * refined`) and the torch.cat((out, skip), 1) at idwt_upsample.py:163 (the synthesis
 * writes channels [0, C) of the concat buffer directly).  Also the adjoint (= inverse) used for DWT gradients.
 * hf == NULL means all-zero details.  gate (optional, same layout as hf) multiplies each detail before synthesis.
 * ---------------------------------------------------------------------------------------------------------- */
int wf_idwt3d_ncdhw(const void *ll, const void *hf, const void *gate, void *x, int dtype, int64_t n, int d, int h,
                    int w, int64_t hf_band_stride, void *stream);

int wf_idwt3d_ndhwc(const void *ll, const void *hf, const void *gate, void *x, int dtype, int B, int d, int h, int w,
                    int C, int64_t ll_vox_stride, int64_t hf_band_stride, int64_t x_vox_stride, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel group 2: multiscale window attention on the LL band.
 * Replaces Block.window_partition (reference network_models/wave_helper.py:450-461), Attention.forward
 * (network_models/attention.py:83-104) and the reshape-only "window reverse" (wave_helper.py:498-499) in one call.
 *
 * x:    [B, D1, H1, W1, C] channels-last, contiguous.  Windows of ws^3 tokens tile the grid, window order
 *       (b, zblk, yblk, xblk), token order (dz, dy, dx).
 * out:  [B * nW * ws^3, C] contiguous in WINDOW order; the reference re-reads exactly this buffer as
 *       [B, D1, H1, W1, C] without an inverse permute, so `out` viewed with that shape IS the reference result.
 * bias_t:   dense relative-position bias produced by wf_relpos_bias_expand, fp32 [heads, N, N] stored TRANSPOSED
 *           (bias_t[h][j][i] = table[index[i][j]][h]) so a warp of queries reads it coalesced (CUDA-core kernels).
 * bias_img: dense 16-bit image produced by wf_relpos_bias_image for the tensor-core kernels (N = 512 only):
 *           img[h][i][520] = fmt(table[index[i][j]][h] * log2 e), rows padded to 520 so the (head, 128-query) slab a
 *           CTA needs is ONE contiguous 133120-byte bulk copy whose rows are bank-conflict free in shared memory.
 * `dtype` is the weight type and the operand format of the GEMMs:
 *   WF_F32  fp32 CUDA-core kernels (x, weights, out fp32; needs bias_t);
 *   WF_BF16 tcgen05 tensor-core kernels when bias_img != NULL and the geometry is 512-token windows with head_dim 16
 *           (wf_window_attn_tc_supported), else the CUDA-core kernels (need bias_t, out_dtype == WF_BF16);
 *   WF_F16  the same tensor-core kernels with fp16 operands (10-bit mantissa: q, k, v, P and O are rounded 8x finer than
 *           in bf16 at the same tensor-core rate); tensor-core geometry only.
 * x may be stored as x_dtype = WF_F32 or WF_BF16 (converted to the operand format while the tile is staged);
 * out_dtype is `dtype` or WF_F32 (the projection epilogue writes its fp32 accumulators).
 * Weights are in `dtype`; qkv_w [3C, C], qkv_b [3C], proj_w [C, C], proj_b [C] (PyTorch Linear layout).
 * head_dim = C / heads must be 8, 16, 32 or 64; scale multiplies q after its bias (attention.py:88).
 * ---------------------------------------------------------------------------------------------------------- */
int wf_relpos_bias_expand(const void *table, int table_dtype, const int64_t *index, float *bias_t, int heads, int N,
                          int table_rows, void *stream);

size_t wf_relpos_bias_image_bytes(int heads, int N);
int wf_relpos_bias_image(const void *table, int table_dtype, const int64_t *index, void *img, int fmt, int heads, int N,
                         int table_rows, void *stream);

int wf_window_attn_tc_supported(int D1, int H1, int W1, int C, int heads, int ws);
size_t wf_window_attn_workspace_bytes(int dtype, int B, int D1, int H1, int W1, int C, int heads, int ws);

int wf_window_attn_fwd(const void *x, int x_dtype, const void *qkv_w, const void *qkv_b, const void *proj_w,
                       const void *proj_b, const float *bias_t, const void *bias_img, void *out, int out_dtype,
                       void *workspace, size_t workspace_bytes, int dtype, int B, int D1, int H1, int W1, int C,
                       int heads, int ws, float scale, void *stream);

/* Error-compensated fp16 tensor-core attention (tensor-core geometry only: wf_window_attn_tc_supported).  Every GEMM
 * operand that feeds the softmax is carried as a pair hi = fp16(v), lo = fp16(v - hi): x and the weights in both Linear
 * layers (passed as two fp16 images each, qkv_w_hi / qkv_w_lo [3C, C], proj_w_hi / proj_w_lo [C, C]), q and k between the
 * QKV projection and the scores, O between the core and the output projection; products are accumulated as
 * hi*hi + lo*hi + hi*lo (three tcgen05.mma instead of one, on a tensor pipe that is ~5 % busy in this operator), biases are
 * fp32, x and out are fp32.  Same semantics as wf_window_attn_fwd (attention.py:83-104); what changes is that the
 * scores are exact to ~2^-21 |q||k| instead of 2^-11: with plain fp16 operands the score error, amplified by the softmax,
 * is the largest single term of the 16-bit policy's logit error.  P and v remain plain fp16 (their errors average out
 * over the 512 keys).  bias_img: the fp16 image of wf_relpos_bias_image. */
size_t wf_window_attn_split_workspace_bytes(int B, int D1, int H1, int W1, int C, int heads, int ws);
int wf_window_attn_fwd_split(const float *x, const void *qkv_w_hi, const void *qkv_w_lo, const float *qkv_b,
                             const void *proj_w_hi, const void *proj_w_lo, const float *proj_b, const void *bias_img,
                             float *out, void *workspace, size_t workspace_bytes, int B, int D1, int H1, int W1, int C,
                             int heads, int ws, float scale, void *stream);

/* Gradient of the attention core between the two Linear layers (training; reference network_models/attention.py:87-101,
 * differentiated): S = (scale q) k^T + table[index], P = softmax(S), O = P v.
 * workspace : the fp32 workspace wf_window_attn_fwd(dtype = WF_F32) left behind (q pre-scaled, k, v head-major
 *             [windows][heads][N][hd], then O [windows * N, C]);
 * d_o       : gradient wrt O, [windows * N, C] (= grad_out @ proj_w, a library GEMM done by the caller);
 * d_qkv     : receives the gradient wrt the qkv Linear's output, [windows * N, 3C] (q part already multiplied by scale);
 * d_table   : [table_rows, heads], ACCUMULATED into (the caller zeroes it);
 * stats     : scratch of wf_window_attn_bwd_stats_bytes (log-sum-exp and delta of every score row).
 * Probabilities are recomputed, never stored.  The gradients of the two Linear layers themselves are plain GEMMs on
 * d_qkv / d_o and stay with the caller. */
size_t wf_window_attn_bwd_stats_bytes(int64_t windows, int N, int heads);
int wf_window_attn_bwd(const float *workspace, const float *bias_t, const float *table, const int64_t *index,
                       const float *d_o, float *d_qkv, float *d_table, float *stats, int64_t windows, int N, int C,
                       int heads, int table_rows, float scale, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Block glue that dominated the step as library calls, channels-last, fp32 accumulate.
 * ---------------------------------------------------------------------------------------------------------- */

/* PatchMerging front half: y[b, z, y, x, s*C + c] = LayerNorm_{8C}(x[b, 2z+i_s, 2y+j_s, 2x+k_s, c]) for the 8 octants
 * s = 0..7, (i_s, j_s, k_s) = bits 2, 1, 0 of ((octants >> 3s) & 7).  Replaces the eight strided slices + torch.cat + self.norm
 * of PatchMerging.forward (reference network_models/wave_helper.py:170-194; MONAI 0.9's order repeats two octants, which
 * is why the order is an argument) - the [.., 8C] concatenation is never stored.  x: fp32 [B, D, H, W, C] dense, D/H/W even,
 * C % 16 == 0 and C <= 192; gamma / beta: fp32 [8C]; y: [B, D/2, H/2, W/2, 8C] dense, out_dtype WF_F32 or WF_BF16. */
int wf_patch_merge_layernorm(const float *x, const float *gamma, const float *beta, void *y, int out_dtype, int B, int D,
                             int H, int W, int C, uint32_t octants, float eps, void *stream);

/* Depthwise 3x3x3 convolution, zero padding 1: y[b,z,y,x,c] = bias[c] + sum_taps w27[tap][c] * x[b,z+dz,y+dy,x+dx,c].
 * Replaces CCF_FFN.dwconv (reference network_models/wave_helper.py:231-232,283) and ProjectionUpsample.conv1[1]
 * (wave_helper.py:44).  x, y: [B, D, H, W, C] dense channels-last; w27: fp32 [27][C], tap = (dz+1)*9+(dy+1)*3+(dx+1),
 * i.e. the [C,1,3,3,3] weight transposed; bias: fp32 [C] or NULL. */
int wf_dwconv3d_ndhwc(const void *x, const float *w27, const float *bias, void *y, int dtype, int B, int D, int H, int W,
                      int C, void *stream);

/* The same convolution (bf16, C % 8 == 0, W >= 8) that also reduces the statistics of the normalisation that follows it
 * (ProjectionUpsample.norm, reference network_models/wave_helper.py:59,74): mean_rstd receives fp32 [B][C][2] = (mean,
 * 1/sqrt(var + eps)) of the ROUNDED result, sums is fp64 scratch [B][C][2].  Other geometries: WF_ERR_UNSUPPORTED (call
 * wf_dwconv3d_ndhwc + wf_instnorm_stats_ndhwc). */
int wf_dwconv3d_ndhwc_stats(const void *x, const float *w27, const float *bias, void *y, double *sums, float *mean_rstd,
                            float eps, int dtype, int B, int D, int H, int W, int C, void *stream);

/* Affine-free InstanceNorm3d statistics of x [B, S voxels, C] (voxel stride x_vox_stride): sums is scratch
 * (fp64 [B][C][2]), mean_rstd receives fp32 [B][C][2] = (mean, 1/sqrt(var + eps)), biased variance.
 * Replaces the statistics half of nn.InstanceNorm3d in MONAI UnetResBlock (monai/networks/blocks/dynunet_block.py:
 * 98-111) and ChannelCalibration (reference network_models/network_backbone.py:118-121). */
int wf_instnorm_stats_ndhwc(const void *x, double *sums, float *mean_rstd, int dtype, int B, int64_t S, int C,
                            int64_t x_vox_stride, float eps, void *stream);

/* y = act(((x - mean) * rstd) * gamma + beta + R) with R = 0 (res NULL), res (res_mean_rstd NULL) or
 * (res - mean_r) * rstd_r; gamma / beta (fp32 [C]) optional.  act: 0 none, 1 ReLU, 2 LeakyReLU(slope).
 * Fuses norm + residual add + activation of dynunet_block.py:100-110; with gamma / beta it is GroupNorm(num_groups = C)
 * of ProjectionUpsample.norm (reference network_models/wave_helper.py:59,74).  x and res share `dtype`; y_dtype is `dtype`, or
 * WF_BF16 with dtype WF_F32 / WF_F16 (a block kept in fp32-TF32 or fp16 by the precision policy writing into a bf16 concat
 * buffer). */
int wf_instnorm_apply_ndhwc(const void *x, const float *mean_rstd, const void *res, const float *res_mean_rstd,
                            const float *gamma, const float *beta, void *y, int act, float slope, int dtype, int y_dtype,
                            int B, int64_t S, int C, int64_t x_vox_stride, int64_t res_vox_stride, int64_t y_vox_stride,
                            void *stream);

/* y = act((x - mean) * rstd + (conv3(xin) - mean_r) * rstd_r), where conv3 is the 1x1x1 shortcut convolution of a FOUR-channel
 * input: the last pass of the network's first residual block (MONAI UnetResBlock.forward, monai/networks/blocks/
 * dynunet_block.py:104-110: `residual = norm3(conv3(inp)); out += residual; out = lrelu(out)`) with the shortcut recomputed per
 * voxel instead of being read back - together with wf_conv3d_c4_in_stats(y1 = NULL), which still delivers res_mean_rstd, the
 * C-channel shortcut tensor never exists in memory.  x, y: [B, S, C] 16-bit (`dtype` WF_BF16 / WF_F16) with voxel strides;
 * xin: [B, S, 4] dense, xin_dtype = WF_F32 (rounded to `dtype` per value, as the convolution kernel does) or `dtype`;
 * w4: fp32 [C][4] = the operand-format weights of conv3 widened to fp32.  act as in wf_instnorm_apply_ndhwc. */
int wf_instnorm_apply_shortcut4_ndhwc(const void *x, const float *mean_rstd, const void *xin, int xin_dtype, const float *w4,
                                      const float *res_mean_rstd, void *y, int act, float slope, int dtype, int B, int64_t S, int C,
                                      int64_t x_vox_stride, int64_t y_vox_stride, void *stream);

/* InstanceNorm statistics (mean, 1/sqrt(var + eps), biased variance) of the 1x1x1 shortcut convolution y_c = sum_k w4[c][k] x_k of a
 * FOUR-channel volume x [B, S, 4], derived from the 4 first and 10 second moments of x per sample (mean_c = w_c . m, var_c =
 * w_c^T Cov w_c) - the statistics half of `norm3(conv3(inp))` (monai/networks/blocks/dynunet_block.py:104-106) without computing
 * conv3.  x_dtype WF_F32 (each value rounded to op_dtype first, as the convolution kernel does) or op_dtype (WF_BF16 / WF_F16);
 * w4 fp32 [C][4]; sums: fp64 scratch [B * 14]; mean_rstd: fp32 [B * C * 2]. */
int wf_shortcut4_stats(const void *x, int x_dtype, int op_dtype, const float *w4, double *sums, float *mean_rstd, float eps, int B,
                       int64_t S, int C, void *stream);

/* out[b, v, k] = head_b[k] + sum_c head_w[k, c] * act(norm(x)[b, v, c] + R): wf_instnorm_apply_ndhwc fused with the 1x1x1
 * output convolution that is its only consumer (Waveformer.out, reference network_models/network_backbone.py:407;
 * UnetOutBlock, monai/networks/blocks/dynunet_block.py:266), so the last C-channel activation is never written.
 * head_w fp32 [K, C], head_b fp32 [K] or NULL, out [B, S, K] dense (out_dtype WF_F32, or WF_BF16 with dtype WF_BF16). */
int wf_instnorm_apply_head_ndhwc(const void *x, const float *mean_rstd, const void *res, const float *res_mean_rstd,
                                 const float *head_w, const float *head_b, void *out, int act, float slope, int dtype,
                                 int out_dtype, int B, int64_t S, int C, int K, int64_t x_vox_stride,
                                 int64_t res_vox_stride, void *stream);

/* y[r, :] = LayerNorm(x[r, :C]) * gamma + beta (gamma / beta fp32 [C] or NULL), optionally followed by GELU(erf).
 * Rows are voxels of a channels-last tensor (row strides in elements); input and output storage types are independent.
 * y2 (optional, dense [rows, C], y2_dtype WF_BF16 or WF_F16) receives the same result rounded to 16 bits: the GEMM
 * operand, while y keeps the fp32 copy that CCF_FFN's own residual adds back (wave_helper.py:293).
 * Replaces Block.norm1 / norm2 (reference network_models/wave_helper.py:477,509), CCF_FFN.norm1 / norm2 + act
 * (wave_helper.py:278,286), PatchMerging.norm (wave_helper.py:192) and proj_out (network_models/waveformer.py:193-204). */
int wf_layernorm_ndhwc(const void *x, const float *gamma, const float *beta, void *y, void *y2, int y2_dtype, int in_dtype,
                       int out_dtype, int64_t rows, int C, int64_t x_row_stride, int64_t y_row_stride, float eps,
                       int gelu, void *stream);

/* out = a + b + c + bias[channel]: the tail of a transformer block, x + LN(x) + ffn(LN(x)) (+ the fc bias) in ONE fp32 pass
 * (reference network_models/wave_helper.py:293 and :509).  a, b, out fp32 [rows, C]; c fp32 or bf16 [rows, C]; bias fp32 [C]
 * or NULL; C % 4 == 0. */
int wf_residual_sum(const float *a, const float *b, const void *c, int c_dtype, const float *bias, float *out, int64_t rows,
                    int C, void *stream);

/* GroupNorm(num_groups = C) + 1x1x1 convolution folded into one per-sample linear map (ProjectionUpsample.norm -> conv2,
 * reference network_models/wave_helper.py:59-60,74-75): with (mean, rstd) = mean_rstd[b][c] of the normalised tensor,
 * a = rstd * gamma, d = beta - mean * a:  w_folded[b][n][c] = w[n][c] * a[b][c],  b_folded[b][n] = bias[n] + sum_c w[n][c] d[b][c].
 * w [N, C] and bias [N] (or NULL) in w_dtype, outputs in out_dtype (same, or WF_BF16 from WF_F32); gamma / beta fp32 or NULL. */
int wf_groupnorm_fold_linear(const float *mean_rstd, const float *gamma, const float *beta, const void *w, const void *bias,
                             void *w_folded, void *b_folded, int w_dtype, int out_dtype, int B, int C, int N, void *stream);

/* x = GELU(x) (exact erf form) in place on n elements (n % 8 == 0 for bf16, % 4 for fp32): the activation between the
 * 1x1x1 convolutions of ProjectionUpsample (reference network_models/wave_helper.py:47-63). */
int wf_gelu_inplace(void *x, int dtype, int64_t n, void *stream);

/* y = base + sum_s trilinear_upsample(srcs[s]) (sources summed in order, then added to base; base may be NULL).
 * srcs[s]: [B, d_s, h_s, w_s, C] dense channels-last of src_dtype, src_dims = int[3 * nsrc]; base / y: [B, D, H, W, C]
 * of io_dtype with voxel strides.  align_corners = 0 reproduces F.interpolate(size=(D,H,W), mode='trilinear') + the
 * level sum + the shortcut add of Block.multi_scale_forward (reference network_models/wave_helper.py:500-508);
 * align_corners = 1 reproduces nn.Upsample(scale_factor, 'trilinear', align_corners=True) of ProjectionUpsample
 * (wave_helper.py:42,63).  srcs and src_dims are HOST arrays (read during the call). */
int wf_upsample_trilinear_add_ndhwc(const void *const *srcs, const int *src_dims, int nsrc, const void *base, void *y,
                                    int src_dtype, int io_dtype, int align_corners, int B, int D, int H, int W, int C,
                                    int64_t base_vox_stride, int64_t y_vox_stride, void *stream);

/* hi = fp16(x), lo = fp16(x - hi) for n fp32 values (n % 4 == 0): an error-compensated fp16 pair.  conv(hi) + conv(lo) is the
 * convolution of the unrounded input; used for the first convolution of the skip blocks encoder2..4 (reference
 * network_models/network_backbone.py:387-389 -> monai/networks/blocks/dynunet_block.py:98). */
int wf_split_f16(const float *x, void *hi, void *lo, int64_t n, void *stream);

/* Patch embedding for 4 input channels: Conv3d(4 -> Cout, kernel = stride = 2, bias) of a channels-last window x
 * [B, D, H, W, 4] (x_dtype fp32 / bf16 / fp16) written as the channels-last fp32 residual stream y [B, D/2, H/2, W/2, Cout]
 * (exact fp32 arithmetic).  Replaces PatchEmbed.proj + the rearrange at the top of MultiscaleTransformer.forward_features
 * (reference monai/networks/blocks/patchembedding.py:196-225, network_models/waveformer.py:260-270).
 * wpack: fp32 [8 taps = (dz, dy, dx)][4][Cout]; bias fp32 [Cout] or NULL; Cout a multiple of 12. */
int wf_patch_embed_k2s2_c4(const void *x, int x_dtype, const float *wpack, const float *bias, float *y, int B, int D, int H,
                           int W, int Cout, void *stream);

/* CCF_FFN's two pointwise GEMMs fused with the normalisations around them (tcgen05; C = 48 or 96, i.e. encoder stages 1
 * and 2).  Reference: `x = x + drop_path(mlp(norm2(x)))` (network_models/wave_helper.py:509) with CCF_FFN.forward
 * (wave_helper.py:260-294) = n + fc(GELU(LN(dwconv(GELU(LN(pwconv(n))))))), n = norm2(x).
 *   wf_ffn_front: t1[r, :4C] = GELU(LayerNorm_4C(pw_w . LayerNorm_C(x[r, :C]) + pw_b))     x fp32 -> t1 16-bit (dtype)
 *   wf_ffn_back : out[r, :C] = x[r] + LayerNorm_C(x[r]) + fc_w . GELU(LayerNorm_4C(t2[r, :4C])) + fc_b      -> fp32
 * (the depthwise 3^3 stencil between them is wf_dwconv3d_ndhwc).  pw_w [4C, C] and fc_w [C, 4C] are in `dtype`
 * (WF_BF16 / WF_F16), every bias / gamma / beta is fp32 (NULL = identity), rows are voxels (dense). */
int wf_ffn_front(const float *x, const float *norm2_w, const float *norm2_b, float norm2_eps, const void *pw_w,
                 const float *pw_b, const float *ln_w, const float *ln_b, float ln_eps, void *t1, int dtype, int64_t rows,
                 int C, void *stream);
int wf_ffn_back(const void *t2, int dtype, const float *ln_w, const float *ln_b, float ln_eps, const void *fc_w,
                const float *fc_b, const float *x, const float *norm2_w, const float *norm2_b, float norm2_eps, float *out,
                int64_t rows, int C, void *stream);

/* Tail of ProjectionUpsample (reference network_models/wave_helper.py:75-81):
 *     out[r, :N] = w1 . GELU(h[r, :K1]) + b1  +  w2 . u[r, :K2] + b2  +  addend[r, :N]
 * - the last 1^3 convolution of `conv3` applied to the (exact, erf) GELU of its input, plus the residual branch `res_conv` (Upsample +
 * 1^3 convolution of the module's input) either as a second product on the upsampled input u (K2 > 0) or, because a 1^3
 * convolution commutes with trilinear interpolation, as the fp32 `addend` = Upsample(res_conv's convolution at LOW resolution)
 * (K2 = 0: u, w2, b2 ignored) - written straight into a channel slice of the decoder's concatenation buffer (row pitch
 * out_row_stride elements).  h, u, w1 [N, K1], w2 [N, K2] and out are 16-bit (dtype: WF_BF16 or WF_F16), biases fp32 or NULL, addend
 * fp32 dense [rows, N] or NULL, accumulation fp32.  One tcgen05 kernel instead of a GELU pass, two library GEMMs and an add.
 * (K1, K2, N) in {(192, 0, 48), (192, 96, 48), (192, 192, 48)} - the reference configuration's learnable_up3 / learnable_up4 -
 * otherwise WF_ERR_UNSUPPORTED. */
int wf_pw_gelu_dual(const void *h, const void *u, int dtype, const void *w1, const float *b1, const void *w2, const float *b2,
                    const float *addend, void *out, int64_t rows, int K1, int K2, int N, int64_t out_row_stride, void *stream);

/* 3x3x3 convolution (padding 1, no bias) of a 4-channel channels-last volume x [B, D, H, W, 4] (op_dtype WF_BF16 or
 * WF_F16 = format of the tensor-core operands, of wpack and of both results; x_dtype = WF_F32 - converted while gathered -
 * or op_dtype; fp32 accumulation), fused with an optional 1x1x1 convolution of the same input and with the InstanceNorm
 * statistics of both 16-bit results.  Replaces conv1 + norm1's statistics and conv3 + norm3's
 * statistics of the first residual block, Waveformer.encoder1 (reference network_models/network_backbone.py:247-255 ->
 * monai/networks/blocks/dynunet_block.py:98-111).
 * wpack: bf16 [n0 + n1][112], k = tap * 4 + channel with tap = (dz+1)*9 + (dy+1)*3 + (dx+1), zero padded to 112; rows
 *        >= n0 carry the 1x1x1 weights in the centre tap (k = 52..55).
 * y0 / y1: bf16 [B, D, H, W, n0 / n1] with voxel strides (channel slices of wider buffers allowed); n1 may be 0; y1 may be
 *        NULL with n1 > 0: only the statistics of the 1x1x1 result are produced (see wf_instnorm_apply_shortcut4_ndhwc).
 * sums0 / sums1: fp64 scratch [B * n * 2]; mean_rstd0 / mean_rstd1: fp32 [B * n * 2] = (mean, 1/sqrt(var + eps)). */
int wf_conv3d_c4_in_stats(const void *x, int x_dtype, int op_dtype, const void *wpack, void *y0, int64_t y0_vox_stride, int n0,
                          void *y1, int64_t y1_vox_stride, int n1, double *sums0, double *sums1, float *mean_rstd0,
                          float *mean_rstd1, float eps, int B, int D, int H, int W, void *stream);

/* 3x3x3 convolution (padding 1, no bias) 48 -> 48 channels on channels-last bf16 rows of W = 128 voxels: a producer /
 * consumer tcgen05 implicit GEMM (input rows streamed through a shared-memory ring, weights resident, double-buffered TMEM
 * accumulators) that also (i) applies InstanceNorm + LeakyReLU(slope) to its INPUT while staging it when in_mean_rstd is
 * given, and (ii) returns the InstanceNorm statistics of its bf16 OUTPUT.  Replaces `conv2(lrelu(norm1(.)))` + norm2's
 * statistics in the 128^3 residual blocks (reference monai/networks/blocks/dynunet_block.py:98-111, used by
 * Waveformer.encoder1 / decoder1, network_models/network_backbone.py:386,405).
 * wpack: bf16 [3 dz][3 dx][3 k-steps][2 chunks][144 = 3 dy x 48 out][8], element = w[out, 16 ks + 8 chunk + e, dz, dy, dx]
 * (the three dy taps of a (dz, dx, k-step) are adjacent along N so that one tcgen05.mma feeds up to three output rows).
 * dtype: WF_BF16 or WF_F16 - the one 16-bit format of x, wpack and y.  sums: fp64 scratch [B * 48 * 2]; mean_rstd (out) /
 * in_mean_rstd (in, optional): fp32 [B * 48 * 2] = (mean, 1/sqrt(var + eps)) pairs. */
int wf_conv3d_k3_c48_in_stats(const void *x, int dtype, const void *wpack, void *y, double *sums, float *mean_rstd,
                              const float *in_mean_rstd, float slope, float eps, int B, int D, int H, int W,
                              int64_t x_vox_stride, int64_t y_vox_stride, void *stream);

/* The same 48 -> 48 convolution (rolling-row kernel, no input normalisation) with an optional 16-bit `addend` laid out like y
 * that is added to the fp32 accumulators before rounding: y = conv(x) + addend.  Two launches compute a 3x3x3 convolution whose
 * 96 input channels are two 48-channel halves of a concatenation buffer - y1 = conv_a(cat[..., :48]), y = conv_b(cat[..., 48:])
 * + y1 - which is how decoder1's first convolution runs (reference monai/networks/blocks/dynunet_block.py:98 on the
 * torch.cat of monai/networks/blocks/unetr_block.py:83-84).  addend may be NULL; statistics are those of the final y. */
int wf_conv3d_k3_c48_add_stats(const void *x, int dtype, const void *wpack, const void *addend, void *y, double *sums,
                               float *mean_rstd, float eps, int B, int D, int H, int W, int64_t x_vox_stride,
                               int64_t add_vox_stride, int64_t y_vox_stride, void *stream);

/* Diagnostic twin of wf_conv3d_k3_c48_in_stats (fp16 only): same computation, and every CTA also reports how many clocks each of
 * its warp roles spent in each phase.  Rolling-row kernel (in_mean_rstd == NULL): clocks [grid][3 roles][8] with
 *   producer {wait: free ring slot, -, -, -, -, -, -, total}
 *   issuer   {wait: free accumulator slot, wait: staged row, tcgen05.mma issue + commit, -, -, -, -, total}
 *   epilogue {wait: finished row, tcgen05.ld, zero + release, staging tile free, pack + statistics + st.shared, fence + barrier +
 *             bulk store, -, total};
 * block kernel (in_mean_rstd != NULL): clocks [grid][8] = {loader: wait slot, wait copies, total; issuer: wait accumulator, wait
 * row, total; epilogue: wait accumulators, total}.  The stage that never waits is the kernel's limiter. */
int wf_conv3d_k3_c48_stage_clocks(const void *x, int dtype, const void *wpack, void *y, double *sums, float *mean_rstd,
                                  const float *in_mean_rstd, float slope, float eps, int B, int D, int H, int W,
                                  int64_t x_vox_stride, int64_t y_vox_stride, long long *clocks, void *stream);

/* ConvTranspose3d(kernel 2, stride 2, no bias) on channels-last 16-bit activations (dtype WF_BF16 or WF_F16: x, wpack
 * and y share it) as one tensor-core GEMM whose epilogue
 * writes every output voxel in place, e.g. into channels [0, Cout) of a concatenation buffer (y_vox_stride = 2 * Cout).
 * Replaces UnetrUpBlock.transp_conv and the torch.cat that follows it (reference monai/networks/blocks/unetr_block.py:57-86,
 * Waveformer.decoder1, network_models/network_backbone.py:405).
 * x: [B, D, H, W, Cin] (voxel stride x_vox_stride); wpack: bf16 [8 * Cout, Cin], row (dz*4 + dy*2 + dx) * Cout + co holds
 * w[:, co, dz, dy, dx]; y: [B, 2D, 2H, 2W, Cout] with voxel stride y_vox_stride.  Cin, Cout multiples of 16. */
int wf_convtranspose3d_k2s2_ndhwc(const void *x, const void *wpack, void *y, int dtype, int B, int D, int H, int W, int Cin,
                                  int Cout, int64_t x_vox_stride, int64_t y_vox_stride, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sliding-window stitching (re-hosted MONAI inferer, reference monai/inferers/utils.py:216-299).
 * ---------------------------------------------------------------------------------------------------------- */

/* Gather `nwin` windows (roi r0 x r1 x r2, starts in `starts` = int32 [nwin][4] = {batch index, z0, y0, x0}, device
 * memory) from vol [Bv, C, D, H, W] fp32 into win [nwin, C, r0, r1, r2] (dtype, NCDHW contiguous) or, when
 * channels_last != 0, [nwin, r0, r1, r2, C].  Replaces torch.cat([inputs[s] ...]) at inferers/utils.py:223.
 * flip (bit 0 / 1 / 2 = mirror z / y / x): the window is cut from the MIRRORED volume V'(p) = V(flip(p)) without that
 * copy ever being built - the mirror test-time augmentation of light_training/prediction.py:129-156
 * (`torch.flip(x, dims)` in front of the inferer) as an index transform.  0 = plain gather. */
int wf_sw_gather(const float *vol, void *win, const int32_t *starts, int nwin, int dtype, int channels_last, int C,
                 int D, int H, int W, int r0, int r1, int r2, int flip, void *stream);

/* acc[b, k, z0+z, y0+y, x0+x] += max(gz[z]*gy[y]*gx[x], floor) * seg[n, k, z, y, x] for every window n (atomic
 * adds: windows of one call may overlap).  seg: [nwin, K, r0, r1, r2] NCDHW or NDHWC (channels_last).  gz/gy/gx are
 * the 1-D gaussian factors of compute_importance_map (monai/data/utils.py:1121-1138).  Replaces
 * `seg *= w; out[slice] += seg` at inferers/utils.py:287-289 / :351-360.
 * flip != 0: `starts` are positions in the mirrored volume (see wf_sw_gather); the contribution lands at flip(p), i.e.
 * the `torch.flip(..., dims)` of the stitched result at prediction.py:135-155 is part of the scatter. */
int wf_sw_accumulate(const void *seg, float *acc, const int32_t *starts, const float *gz, const float *gy,
                     const float *gx, float floor_w, int nwin, int dtype, int channels_last, int K, int D, int H,
                     int W, int r0, int r1, int r2, int flip, void *stream);

/* acc[b, k, z, y, x] /= sum over every window w of `all_starts` (int32 [nall][4], the FULL window list, not just this
 * rank's; an entry whose volume slot is < 0 applies to EVERY volume of the call - volumes of one call share their
 * geometry, so nall is normally one volume's window count; nall <= 3072) covering the voxel of max(gz*gy*gx, floor).  The count map is geometry only,
 * so it is recomputed here instead of being stored and reduced (inferers/utils.py:265-276, :298-299).
 * labels (optional, uint8 [Bv, D, H, W]) receives argmax over k (4_predict.py:241).
 * Only planes z in [z_begin, z_end) of every volume are normalised, so a slab whose windows are all accumulated can be
 * finished (and copied to the host) while later windows still run; (0, D) = the whole volume.
 * flip: the volume was stitched from a mirrored pass (its count map is the plain one read at flip(q)).
 * dst (optional, same shape as acc): dst = (dst_add ? dst : 0) + dst_scale * normalised acc - the running mean over the
 * 2^k mirrored passes of prediction.py:129-157 without a pass of its own. */
int wf_sw_finalize(float *acc, uint8_t *labels, const int32_t *all_starts, int nall, const float *gz, const float *gy,
                   const float *gx, float floor_w, int Bv, int K, int D, int H, int W, int r0, int r1, int r2,
                   int z_begin, int z_end, int flip, float *dst, float dst_scale, int dst_add, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEFORMER_B200_H */
