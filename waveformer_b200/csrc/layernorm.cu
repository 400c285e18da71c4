// LayerNorm over the channel dimension of channels-last activations, optionally fused with GELU(erf) (sm_100a).
//
// Reference call sites: Block.norm1 / norm2 (network_models/wave_helper.py:477,509), CCF_FFN.norm1/norm2 followed by
// GELU (wave_helper.py:278,286), PatchMerging.norm (wave_helper.py:192) and the affine-free proj_out LayerNorm
// (network_models/waveformer.py:193-204).  One streaming pass: a row (C <= 2048 channels) is held in registers by
// TPR = 8 / 16 / 32 cooperating lanes, mean and centred variance are exact two-pass fp32, and input / output storage
// types are independent (fp32 residual stream in, bf16 GEMM operand out).
#include "wf_common.cuh"

namespace wf {

// GELU(x) = x * Phi(x) with the exact (erf) definition nn.GELU() uses.  erf by Abramowitz-Stegun 7.1.26 (|error| <=
// 1.5e-7, i.e. fp32 round-off level): one reciprocal, one ex2 and six FMAs instead of erff's ~25 instructions - this
// kernel is ALU-bound on the exact routine.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-p, e, 1.f);           // erf(|x| / sqrt 2)
    return 0.5f * x + 0.5f * fabsf(x) * erf_abs;      // = x * (1 + sign(x) erf) / 2
}

// Each thread owns NV groups of 4 consecutive channels: channel index = (j * TPR + sub) * 4 + e.
// y2 (optional) receives the same values as bf16 - the GEMM operand - while y keeps the fp32 copy the residual needs.
template <typename TI, typename TO, typename T2, int TPR, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI *__restrict__ x, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, TO *__restrict__ y,
                                                        T2 *__restrict__ y2, int64_t rows, int C, int64_t xs,
                                                        int64_t ys, float eps, int gelu) {
    constexpr int RPB = 256 / TPR;  // rows per block
    const int sub = threadIdx.x % TPR;
    const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / TPR;
    const bool live = row < rows;
    const int groups = C >> 2;  // groups of 4 channels
    float v[NV][4];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int g = j * TPR + sub;
        if (live && g < groups) {
            load4<TI>(x + row * xs + g * 4, v[j]);
            sum += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
        } else {
            v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0.f;
        }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int g = j * TPR + sub;
        if (g < groups) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = v[j][e] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
    if (!live) return;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int g = j * TPR + sub;
        if (g >= groups) continue;
        float o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t = (v[j][e] - mean) * rstd;
            if (gamma != nullptr) t = fmaf(t, __ldg(gamma + g * 4 + e), beta != nullptr ? __ldg(beta + g * 4 + e) : 0.f);
            o4[e] = gelu ? gelu_erf(t) : t;
        }
        store4<TO>(y + row * ys + g * 4, o4);
        if (y2 != nullptr) store4<T2>(y2 + row * (int64_t)C + g * 4, o4);
    }
}

// Wide variant (C % 8 == 0): a row is split into chunks of 8 channels; LPR lanes cooperate on a row and each lane owns
// NCH chunks (chunk index = j * LPR + lane-in-row), i.e. 16-byte (bf16) / 2 x 16-byte (fp32) accesses, NCH * 8 values of
// independent work per thread and log2(LPR) shuffle steps per statistic.  A warp covers 32 / LPR rows.
template <typename TI, typename TO, typename T2, int LPR, int NCH>
__global__ void __launch_bounds__(256) layernorm_wide_kernel(const TI *__restrict__ x, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, TO *__restrict__ y,
                                                             T2 *__restrict__ y2, int64_t rows, int C,
                                                             int64_t xs, int64_t ys, float eps, int gelu) {
    constexpr int RPB = 256 / LPR;
    const int sub = threadIdx.x % LPR;
    const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
    const bool live = row < rows;
    const int chunks = C >> 3;
    float v[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int ch = j * LPR + sub;
        if (live && ch < chunks) {
            load8<TI>(x + row * xs + ch * 8, v[j]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[j][e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        if (j * LPR + sub < chunks) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[j][e] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
    if (!live) return;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int ch = j * LPR + sub;
        if (ch >= chunks) continue;
        float o8[8];
        float g8[8], b8[8];
        if (gamma != nullptr) {
            const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + ch * 8)), gb = __ldg(reinterpret_cast<const float4 *>(gamma + ch * 8 + 4));
            g8[0] = ga.x; g8[1] = ga.y; g8[2] = ga.z; g8[3] = ga.w; g8[4] = gb.x; g8[5] = gb.y; g8[6] = gb.z; g8[7] = gb.w;
            if (beta != nullptr) {
                const float4 ba = __ldg(reinterpret_cast<const float4 *>(beta + ch * 8)), bb = __ldg(reinterpret_cast<const float4 *>(beta + ch * 8 + 4));
                b8[0] = ba.x; b8[1] = ba.y; b8[2] = ba.z; b8[3] = ba.w; b8[4] = bb.x; b8[5] = bb.y; b8[6] = bb.z; b8[7] = bb.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) b8[e] = 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float t = (v[j][e] - mean) * rstd;
            if (gamma != nullptr) t = fmaf(t, g8[e], b8[e]);
            o8[e] = gelu ? gelu_erf(t) : t;
        }
        store8<TO>(y + row * ys + ch * 8, o8);
        if (y2 != nullptr) store8<T2>(y2 + row * (int64_t)C + ch * 8, o8);
    }
}

template <typename TI, typename TO, typename T2>
static bool layernorm_wide_launch(const TI *x, const float *gamma, const float *beta, TO *y, T2 *y2, int64_t rows,
                                  int C, int64_t xs, int64_t ys, float eps, int gelu, cudaStream_t st) {
    if (C % 8 != 0 || (xs * sizeof(TI)) % 16 != 0 || (ys * sizeof(TO)) % 16 != 0 || !aligned16(x) || !aligned16(y) ||
        (y2 != nullptr && !aligned16(y2)) || (gamma != nullptr && !aligned16(gamma)) || (beta != nullptr && !aligned16(beta)))
        return false;
    const int chunks = C / 8;
#define WF_LNW(LPR_, NCH_)                                                                                            \
    do {                                                                                                              \
        const int rpb = 256 / LPR_;                                                                                   \
        layernorm_wide_kernel<TI, TO, T2, LPR_, NCH_><<<(unsigned)((rows + rpb - 1) / rpb), 256, 0, st>>>(                \
            x, gamma, beta, y, y2, rows, C, xs, ys, eps, gelu);                                                       \
        return true;                                                                                                  \
    } while (0)
    if (chunks <= 6) WF_LNW(2, 3);
    if (chunks <= 12) WF_LNW(4, 3);
    if (chunks <= 24) WF_LNW(8, 3);
    if (chunks <= 48) WF_LNW(16, 3);
    if (chunks <= 96) WF_LNW(32, 3);
    if (chunks <= 192) WF_LNW(32, 6);
#undef WF_LNW
    return false;
}

template <typename TI, typename TO, typename T2>
static int layernorm_launch(const TI *x, const float *gamma, const float *beta, TO *y, T2 *y2, int64_t rows,
                            int C, int64_t xs, int64_t ys, float eps, int gelu, cudaStream_t st) {
    if (layernorm_wide_launch<TI, TO, T2>(x, gamma, beta, y, y2, rows, C, xs, ys, eps, gelu, st)) {
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    if (C % 4 != 0 || C > 2048) return WF_ERR_UNSUPPORTED;
    if ((xs * sizeof(TI)) % (sizeof(TI) == 4 ? 16 : 8) != 0 || (ys * sizeof(TO)) % (sizeof(TO) == 4 ? 16 : 8) != 0)
        return WF_ERR_MISALIGNED;
    const int groups = C / 4;
#define WF_LN(TPR_, NV_)                                                                                         \
    do {                                                                                                         \
        const int rpb = 256 / TPR_;                                                                              \
        layernorm_kernel<TI, TO, T2, TPR_, NV_><<<(unsigned)((rows + rpb - 1) / rpb), 256, 0, st>>>(x, gamma, beta, y, y2, \
                                                                                                rows, C, xs, ys, eps, gelu); \
    } while (0)
    if (groups <= 8) WF_LN(8, 1);
    else if (groups <= 16) WF_LN(16, 1);
    else if (groups <= 32) WF_LN(32, 1);
    else if (groups <= 64) WF_LN(32, 2);
    else if (groups <= 128) WF_LN(32, 4);
    else if (groups <= 256) WF_LN(32, 8);
    else WF_LN(32, 16);
#undef WF_LN
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" int wf_layernorm_ndhwc(const void *x, const float *gamma, const float *beta, void *y, void *y2, int y2_dtype,
                                  int in_dtype, int out_dtype, int64_t rows, int C, int64_t x_row_stride,
                                  int64_t y_row_stride, float eps, int gelu, void *stream) {
    if (!x || !y) return WF_ERR_NULL_POINTER;
    if (rows <= 0 || C <= 0 || x_row_stride < C || y_row_stride < C) return WF_ERR_BAD_SHAPE;
    if (y2 != nullptr && y2_dtype != WF_BF16 && y2_dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    using hf = __half;
#define WF_LN_CASE(IC_, OC_, TI_, TO_)                                                                                      \
    if (in_dtype == IC_ && out_dtype == OC_) {                                                                              \
        if (y2 != nullptr && y2_dtype == WF_F16)                                                                            \
            return wf::layernorm_launch<TI_, TO_, hf>((const TI_ *)x, gamma, beta, (TO_ *)y, (hf *)y2, rows, C, x_row_stride, \
                                                      y_row_stride, eps, gelu, st);                                         \
        return wf::layernorm_launch<TI_, TO_, bf>((const TI_ *)x, gamma, beta, (TO_ *)y, (bf *)y2, rows, C, x_row_stride,     \
                                                  y_row_stride, eps, gelu, st);                                             \
    }
    WF_LN_CASE(WF_F32, WF_F32, float, float)
    WF_LN_CASE(WF_F32, WF_BF16, float, bf)
    WF_LN_CASE(WF_BF16, WF_BF16, bf, bf)
    WF_LN_CASE(WF_BF16, WF_F32, bf, float)
    WF_LN_CASE(WF_F32, WF_F16, float, hf)
    WF_LN_CASE(WF_F16, WF_F16, hf, hf)
    WF_LN_CASE(WF_F16, WF_F32, hf, float)
#undef WF_LN_CASE
    return WF_ERR_BAD_DTYPE;
}
