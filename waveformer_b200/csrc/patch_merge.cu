// PatchMerging front half: gather the 8 octants of every 2x2x2 cell and LayerNorm the concatenated 8C-vector, one kernel.
//
// Replaces `x0..x7 = x[:, i::2, j::2, k::2, :]; torch.cat([...], -1); self.norm(x)` of the reference's PatchMerging /
// PatchMergingV2 (network_models/wave_helper.py:125-194).  The concatenated [B, D/2, H/2, W/2, 8C] tensor (100 MB at stage
// 1) is never written: one warp owns one output row, reads its eight C-channel segments straight from the fp32 stream
// (16-byte packets, each segment is contiguous), reduces mean / variance in registers (two-pass, values stay resident) and
// stores the normalised row in the GEMM operand type.  The octant ORDER is an argument: MONAI 0.9's list repeats two
// octants (wave_helper.py:170-194), which is the order the reference trains with.
#include "wf_common.cuh"

namespace wf {

// NPL = 16-byte packets (4 fp32 channels) per lane = 8C / 128
template <typename TO, int NPL>
__global__ void __launch_bounds__(256) patch_merge_ln_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, TO *__restrict__ y,
                                                             int64_t rows, int d, int h, int w, int C, uint32_t octants,
                                                             float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    int64_t t = row;
    const int xo = (int)(t % w); t /= w;
    const int yo = (int)(t % h); t /= h;
    const int zo = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const int ppc = C >> 2;             // packets per segment
    float v[NPL][4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int p = i * 32 + lane;    // packet index within the 8C row
        const int seg = p / ppc, off = (p - seg * ppc) * 4;
        const uint32_t o = (octants >> (3 * seg)) & 7u;      // (i, j, k) = bits 2, 1, 0
        const int64_t vox = ((b * 2 * d + 2 * zo + ((o >> 2) & 1)) * H + 2 * yo + ((o >> 1) & 1)) * W + 2 * xo + (o & 1);
        const float4 f = __ldg(reinterpret_cast<const float4 *>(x + vox * C + off));
        v[i][0] = f.x; v[i][1] = f.y; v[i][2] = f.z; v[i][3] = f.w;
        sum += (f.x + f.y) + (f.z + f.w);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float n = (float)(8 * C);
    const float mean = sum / n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float dlt = v[i][e] - mean;
            sq = fmaf(dlt, dlt, sq);
        }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
    const float rstd = rsqrtf(sq / n + eps);
    TO *yr = y + row * (int64_t)(8 * C);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c0));
        const float4 bt = __ldg(reinterpret_cast<const float4 *>(beta + c0));
        const float o0 = fmaf((v[i][0] - mean) * rstd, g.x, bt.x), o1 = fmaf((v[i][1] - mean) * rstd, g.y, bt.y);
        const float o2 = fmaf((v[i][2] - mean) * rstd, g.z, bt.z), o3 = fmaf((v[i][3] - mean) * rstd, g.w, bt.w);
        const float o4[4] = {o0, o1, o2, o3};
        store4<TO>(yr + c0, o4);
    }
}

template <typename TO>
static int patch_merge_launch(const float *x, const float *gamma, const float *beta, TO *y, int B, int d, int h, int w, int C,
                              uint32_t octants, float eps, cudaStream_t st) {
    const int64_t rows = (int64_t)B * d * h * w;
    const unsigned grid = (unsigned)((rows + 7) / 8);
#define WF_PM(NPL_) patch_merge_ln_kernel<TO, NPL_><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, d, h, w, C, octants, eps)
    switch (8 * C / 128) {
        case 1: WF_PM(1); break;
        case 2: WF_PM(2); break;
        case 3: WF_PM(3); break;
        case 4: WF_PM(4); break;
        case 6: WF_PM(6); break;
        case 8: WF_PM(8); break;
        case 12: WF_PM(12); break;
        default: return WF_ERR_UNSUPPORTED;
    }
#undef WF_PM
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" int wf_patch_merge_layernorm(const float *x, const float *gamma, const float *beta, void *y, int out_dtype, int B,
                                        int D, int H, int W, int C, uint32_t octants, float eps, void *stream) {
    if (!x || !gamma || !beta || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0 || (D | H | W) & 1) return WF_ERR_BAD_SHAPE;
    if (C % 16 != 0 || (8 * C) % 128 != 0) return WF_ERR_UNSUPPORTED;       // whole 16-byte packets, 32 lanes x NPL packets
    if (!wf::aligned16(x) || !wf::aligned16(gamma) || !wf::aligned16(beta) || !wf::aligned16(y)) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == WF_F32)
        return wf::patch_merge_launch<float>(x, gamma, beta, (float *)y, B, D / 2, H / 2, W / 2, C, octants, eps, st);
    if (out_dtype == WF_BF16)
        return wf::patch_merge_launch<__nv_bfloat16>(x, gamma, beta, (__nv_bfloat16 *)y, B, D / 2, H / 2, W / 2, C, octants, eps, st);
    if (out_dtype == WF_F16)
        return wf::patch_merge_launch<__half>(x, gamma, beta, (__half *)y, B, D / 2, H / 2, W / 2, C, octants, eps, st);
    return WF_ERR_BAD_DTYPE;
}
