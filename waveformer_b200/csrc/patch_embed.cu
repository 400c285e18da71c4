// Patch embedding: Conv3d(in_chans -> C, kernel = stride = 2, bias) of a channels-last window, written straight into the
// encoder's channels-last fp32 residual stream.
//
// Reference: MONAI PatchEmbed.forward (monai/networks/blocks/patchembedding.py:196-225) as used by
// MultiscaleTransformer.forward_features (network_models/waveformer.py:260-270): `x = self.proj(x)` followed by the
// rearrange to [B, D, H, W, C].  With kernel == stride every output voxel reads its own 2x2x2 input cell (32 values for 4
// channels): a 48 x 32 matrix-vector product per voxel, HBM-bound.  The library path was a TF32 tensor-core convolution (the
// raw image rounded to 10 bits), a separate bias pass over the 100 MB result and a layout copy (0.2 ms per batch-2 window);
// here it is one exact-fp32 pass: read the window once, write the stream once.
#include "wf_common.cuh"

namespace wf {

// thread = (output voxel, group of 12 output channels); CIN = 4, COUT % 12 == 0.
// wpack: [tap = (dz*2+dy)*2+dx][cin][COUT] fp32, bias [COUT] fp32, both staged in shared memory.
template <typename TIN>
__global__ void __launch_bounds__(256) patch_embed_k2s2_c4_kernel(const TIN *__restrict__ x, const float *__restrict__ wpack,
                                                                  const float *__restrict__ bias, float *__restrict__ y,
                                                                  int64_t total, int d, int h, int w, int COUT) {
    extern __shared__ float sW[];          // [32][COUT] + [COUT]
    for (int i = threadIdx.x; i < 33 * COUT; i += blockDim.x) sW[i] = i < 32 * COUT ? wpack[i] : (bias ? bias[i - 32 * COUT] : 0.f);
    __syncthreads();
    const int groups = COUT / 12;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const int64_t vox = idx / groups;      // ((b*d + z)*h + yy)*w + xx
    const int xx = (int)(vox % w);
    int64_t t = vox / w;
    const int yy = (int)(t % h);
    t /= h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const int64_t v000 = ((b * (2 * d) + 2 * z) * H + 2 * yy) * (int64_t)W + 2 * xx;
    float in[32];
#pragma unroll
    for (int tap = 0; tap < 8; ++tap) {
        const int64_t v = v000 + ((tap >> 2) & 1) * (int64_t)H * W + ((tap >> 1) & 1) * W + (tap & 1);
        if constexpr (sizeof(TIN) == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(x) + v);
            in[4 * tap] = q.x; in[4 * tap + 1] = q.y; in[4 * tap + 2] = q.z; in[4 * tap + 3] = q.w;
        } else {
            float f[4];
            load4<TIN>(x + v * 4, f);
            in[4 * tap] = f[0]; in[4 * tap + 1] = f[1]; in[4 * tap + 2] = f[2]; in[4 * tap + 3] = f[3];
        }
    }
    float acc[12];
    const float *bw = sW + 32 * COUT + g * 12;
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = bw[c];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const float4 *wr = reinterpret_cast<const float4 *>(sW + k * COUT + g * 12);
        const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2];
        const float v = in[k];
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        acc[8] = fmaf(v, w2.x, acc[8]); acc[9] = fmaf(v, w2.y, acc[9]); acc[10] = fmaf(v, w2.z, acc[10]); acc[11] = fmaf(v, w2.w, acc[11]);
    }
    float4 *dst = reinterpret_cast<float4 *>(y + vox * COUT + g * 12);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    dst[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
}

// Four output voxels (consecutive along x) per thread: the one-voxel kernel above issues three 16-byte shared-memory weight loads per
// 12 FMAs and is bound by the L1 / shared-memory pipe (94 % busy in ncu, 0.35 ms for a 0.5 GB problem); here a weight packet feeds
// 48 FMAs.  The taps are walked one at a time so that only four input voxels (16 values) are live next to the 48 accumulators.
template <typename TIN>
__global__ void __launch_bounds__(256) patch_embed_k2s2_c4_x4_kernel(const TIN *__restrict__ x, const float *__restrict__ wpack,
                                                                     const float *__restrict__ bias, float *__restrict__ y,
                                                                     int64_t total, int d, int h, int w4, int COUT) {
    extern __shared__ float sW[];          // [32][COUT] + [COUT]
    for (int i = threadIdx.x; i < 33 * COUT; i += blockDim.x) sW[i] = i < 32 * COUT ? wpack[i] : (bias ? bias[i - 32 * COUT] : 0.f);
    __syncthreads();
    const int groups = COUT / 12;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const int64_t quad = idx / groups;     // ((b*d + z)*h + yy)*w4 + xq
    const int xq = (int)(quad % w4);
    int64_t t = quad / w4;
    const int yy = (int)(t % h);
    t /= h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int w = 4 * w4, H = 2 * h, W = 2 * w;
    const int64_t v000 = ((b * (2 * d) + 2 * z) * H + 2 * yy) * (int64_t)W + 8 * xq;
    float acc[4][12];
    const float *bw = sW + 32 * COUT + g * 12;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[j][c] = bw[c];
#pragma unroll
    for (int tap = 0; tap < 8; ++tap) {
        float in[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t v = v000 + ((tap >> 2) & 1) * (int64_t)H * W + ((tap >> 1) & 1) * W + (tap & 1) + 2 * j;
            if constexpr (sizeof(TIN) == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(x) + v);
                in[j][0] = q.x; in[j][1] = q.y; in[j][2] = q.z; in[j][3] = q.w;
            } else {
                load4<TIN>(x + v * 4, in[j]);
            }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const float4 *wr = reinterpret_cast<const float4 *>(sW + (tap * 4 + ci) * COUT + g * 12);
            const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v = in[j][ci];
                acc[j][0] = fmaf(v, w0.x, acc[j][0]); acc[j][1] = fmaf(v, w0.y, acc[j][1]); acc[j][2] = fmaf(v, w0.z, acc[j][2]); acc[j][3] = fmaf(v, w0.w, acc[j][3]);
                acc[j][4] = fmaf(v, w1.x, acc[j][4]); acc[j][5] = fmaf(v, w1.y, acc[j][5]); acc[j][6] = fmaf(v, w1.z, acc[j][6]); acc[j][7] = fmaf(v, w1.w, acc[j][7]);
                acc[j][8] = fmaf(v, w2.x, acc[j][8]); acc[j][9] = fmaf(v, w2.y, acc[j][9]); acc[j][10] = fmaf(v, w2.z, acc[j][10]); acc[j][11] = fmaf(v, w2.w, acc[j][11]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 *dst = reinterpret_cast<float4 *>(y + (quad * 4 + j) * COUT + g * 12);
        dst[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        dst[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        dst[2] = make_float4(acc[j][8], acc[j][9], acc[j][10], acc[j][11]);
    }
}

}  // namespace wf

extern "C" int wf_patch_embed_k2s2_c4(const void *x, int x_dtype, const float *wpack, const float *bias, float *y, int B, int D,
                                      int H, int W, int Cout, void *stream) {
    if (!x || !wpack || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || (D | H | W) & 1) return WF_ERR_BAD_SHAPE;
    if (Cout <= 0 || Cout % 12 != 0 || Cout > 384) return WF_ERR_UNSUPPORTED;
    if (!wf::aligned16(x) || !wf::aligned16(y) || !wf::aligned16(wpack)) return WF_ERR_MISALIGNED;
    const size_t smem = (size_t)33 * Cout * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (!wf::ab_old() && (W / 2) % 4 == 0) {     // four output voxels per thread
        const int64_t total4 = (int64_t)B * (D / 2) * (H / 2) * (W / 8) * (Cout / 12);
        const unsigned grid4 = (unsigned)((total4 + 255) / 256);
        if (x_dtype == WF_F32)
            wf::patch_embed_k2s2_c4_x4_kernel<float><<<grid4, 256, smem, st>>>((const float *)x, wpack, bias, y, total4, D / 2, H / 2, W / 8, Cout);
        else if (x_dtype == WF_BF16)
            wf::patch_embed_k2s2_c4_x4_kernel<__nv_bfloat16><<<grid4, 256, smem, st>>>((const __nv_bfloat16 *)x, wpack, bias, y, total4, D / 2, H / 2, W / 8, Cout);
        else if (x_dtype == WF_F16)
            wf::patch_embed_k2s2_c4_x4_kernel<__half><<<grid4, 256, smem, st>>>((const __half *)x, wpack, bias, y, total4, D / 2, H / 2, W / 8, Cout);
        else
            return WF_ERR_BAD_DTYPE;
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    const int64_t total = (int64_t)B * (D / 2) * (H / 2) * (W / 2) * (Cout / 12);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (x_dtype == WF_F32)
        wf::patch_embed_k2s2_c4_kernel<float><<<grid, 256, smem, st>>>((const float *)x, wpack, bias, y, total, D / 2, H / 2, W / 2, Cout);
    else if (x_dtype == WF_BF16)
        wf::patch_embed_k2s2_c4_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)x, wpack, bias, y, total, D / 2, H / 2, W / 2, Cout);
    else if (x_dtype == WF_F16)
        wf::patch_embed_k2s2_c4_kernel<__half><<<grid, 256, smem, st>>>((const __half *)x, wpack, bias, y, total, D / 2, H / 2, W / 2, Cout);
    else
        return WF_ERR_BAD_DTYPE;
    WF_LAUNCH_CHECK();
    return WF_OK;
}
