// out = a + b + c + bias (per channel): the tail of a transformer block, x_new = x + LN(x) + ffn(LN(x)) with the fc bias
// (reference network_models/wave_helper.py:293 `x + self.fc(...)` inside CCF_FFN and :509 `x = x + drop_path(mlp(norm2(x)))`),
// as ONE fp32 pass instead of three elementwise kernels.  a, b, out: fp32 [rows, C]; c: fp32 or bf16 [rows, C]; bias fp32 [C].
#include "wf_common.cuh"

namespace wf {

template <typename TC>
__global__ void __launch_bounds__(256) residual_sum_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                           const TC *__restrict__ c, const float *__restrict__ bias,
                                                           float *__restrict__ out, int64_t total4, int c4s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index of a 4-channel packet
    if (i >= total4) return;
    const float4 va = __ldcs(reinterpret_cast<const float4 *>(a) + i), vb = __ldcs(reinterpret_cast<const float4 *>(b) + i);
    float4 vc;
    if constexpr (sizeof(TC) == 4) {
        vc = __ldcs(reinterpret_cast<const float4 *>(c) + i);
    } else {
        const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(c) + i);
        float f[8];
        Pack<TC>::unpack(make_uint4(t.x, t.y, 0u, 0u), f);
        vc = make_float4(f[0], f[1], f[2], f[3]);
    }
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) bb = __ldg(reinterpret_cast<const float4 *>(bias) + (int)(i % c4s));
    float4 o;
    o.x = ((va.x + vb.x) + vc.x) + bb.x; o.y = ((va.y + vb.y) + vc.y) + bb.y;
    o.z = ((va.z + vb.z) + vc.z) + bb.z; o.w = ((va.w + vb.w) + vc.w) + bb.w;
    reinterpret_cast<float4 *>(out)[i] = o;
}

// GELU (exact erf form, nn.GELU() default) in place on a bf16 / fp32 buffer: the activation between the 1x1x1 convolutions
// of ProjectionUpsample (reference network_models/wave_helper.py:47-63).  Same erf evaluation as the LayerNorm kernel
// (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7); 16-byte streaming accesses.
__device__ __forceinline__ float gelu_erf_as(float x) {
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
    return 0.5f * x + 0.5f * fabsf(x) * fmaf(-p, e, 1.f);
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_inplace_kernel(T *__restrict__ x, int64_t packets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= packets) return;
    constexpr int V = Pack<T>::VEC;
    float f[V];
    Pack<T>::unpack(__ldcs(reinterpret_cast<const typename Pack<T>::raw *>(x) + i), f);
#pragma unroll
    for (int e = 0; e < V; ++e) f[e] = gelu_erf_as(f[e]);
    reinterpret_cast<typename Pack<T>::raw *>(x)[i] = Pack<T>::pack(f);
}

// x (fp32) -> hi = fp16(x), lo = fp16(x - hi): an error-compensated fp16 pair (22 significant bits).  A convolution is linear
// in its input, so conv(hi) + conv(lo) reproduces conv(x) without the input's rounding error; used for the first
// convolution of the skip blocks encoder2..4, whose InstanceNorm amplifies exactly that error.
__global__ void __launch_bounds__(256) split_f16_kernel(const float4 *__restrict__ x, uint2 *__restrict__ hi, uint2 *__restrict__ lo,
                                                        int64_t packets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= packets) return;
    const float4 v = __ldcs(x + i);
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    hi[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
    lo[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
}

// GroupNorm(num_groups = C) followed by a 1x1x1 convolution = one per-sample linear map: with (mean, rstd) of every
// (sample, channel), a = rstd * gamma and d = beta - mean * a,   W'[b] = W diag(a[b]),   b'[b] = bias + W d[b].
// One block per (output row n, sample b): replaces ~20 tiny library launches per ProjectionUpsample call.
template <typename TW, typename TO>
__global__ void __launch_bounds__(128) groupnorm_fold_kernel(const float *__restrict__ mr, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, const TW *__restrict__ w,
                                                             const TW *__restrict__ bias, TO *__restrict__ wf_,
                                                             TO *__restrict__ bf_, int C, int N) {
    __shared__ float red[4];
    const int n = blockIdx.x, b = blockIdx.y;
    float part = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) {
        const float2 m = *reinterpret_cast<const float2 *>(mr + ((int64_t)b * C + c) * 2);
        const float a = m.y * (gamma ? gamma[c] : 1.f);
        const float d = (beta ? beta[c] : 0.f) - m.x * a;
        const float wv = to_f32(w[(int64_t)n * C + c]);
        wf_[((int64_t)b * N + n) * C + c] = from_f32<TO>(wv * a);
        part = fmaf(wv, d, part);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0)
        bf_[(int64_t)b * N + n] = from_f32<TO>((bias ? to_f32(bias[n]) : 0.f) + ((red[0] + red[1]) + (red[2] + red[3])));
}

}  // namespace wf

extern "C" int wf_groupnorm_fold_linear(const float *mean_rstd, const float *gamma, const float *beta, const void *w,
                                        const void *bias, void *w_folded, void *b_folded, int w_dtype, int out_dtype, int B,
                                        int C, int N, void *stream) {
    if (!mean_rstd || !w || !w_folded || !b_folded) return WF_ERR_NULL_POINTER;
    if (B <= 0 || C <= 0 || N <= 0 || B > 65535) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)N, (unsigned)B);
    using bf = __nv_bfloat16;
    if (w_dtype == WF_F32 && out_dtype == WF_F32)
        wf::groupnorm_fold_kernel<float, float><<<grid, 128, 0, st>>>(mean_rstd, gamma, beta, (const float *)w, (const float *)bias,
                                                                      (float *)w_folded, (float *)b_folded, C, N);
    else if (w_dtype == WF_BF16 && out_dtype == WF_BF16)
        wf::groupnorm_fold_kernel<bf, bf><<<grid, 128, 0, st>>>(mean_rstd, gamma, beta, (const bf *)w, (const bf *)bias,
                                                                (bf *)w_folded, (bf *)b_folded, C, N);
    else if (w_dtype == WF_F32 && out_dtype == WF_BF16)
        wf::groupnorm_fold_kernel<float, bf><<<grid, 128, 0, st>>>(mean_rstd, gamma, beta, (const float *)w, (const float *)bias,
                                                                   (bf *)w_folded, (bf *)b_folded, C, N);
    else if (w_dtype == WF_F16 && out_dtype == WF_F16)
        wf::groupnorm_fold_kernel<__half, __half><<<grid, 128, 0, st>>>(mean_rstd, gamma, beta, (const __half *)w, (const __half *)bias,
                                                                        (__half *)w_folded, (__half *)b_folded, C, N);
    else if (w_dtype == WF_F32 && out_dtype == WF_F16)
        wf::groupnorm_fold_kernel<float, __half><<<grid, 128, 0, st>>>(mean_rstd, gamma, beta, (const float *)w, (const float *)bias,
                                                                       (__half *)w_folded, (__half *)b_folded, C, N);
    else
        return WF_ERR_BAD_DTYPE;
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_split_f16(const float *x, void *hi, void *lo, int64_t n, void *stream) {
    if (!x || !hi || !lo) return WF_ERR_NULL_POINTER;
    if (n <= 0 || n % 4) return WF_ERR_BAD_SHAPE;
    if (!wf::aligned16(x) || (reinterpret_cast<uintptr_t>(hi) & 7u) || (reinterpret_cast<uintptr_t>(lo) & 7u)) return WF_ERR_MISALIGNED;
    wf::split_f16_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(x), reinterpret_cast<uint2 *>(hi), reinterpret_cast<uint2 *>(lo), n / 4);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_gelu_inplace(void *x, int dtype, int64_t n, void *stream) {
    if (!x) return WF_ERR_NULL_POINTER;
    if (n <= 0) return WF_ERR_BAD_SHAPE;
    if (!wf::aligned16(x)) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32) {
        if (n % 4) return WF_ERR_BAD_SHAPE;
        wf::gelu_inplace_kernel<float><<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>((float *)x, n / 4);
    } else if (dtype == WF_BF16) {
        if (n % 8) return WF_ERR_BAD_SHAPE;
        wf::gelu_inplace_kernel<__nv_bfloat16><<<(unsigned)((n / 8 + 255) / 256), 256, 0, st>>>((__nv_bfloat16 *)x, n / 8);
    } else if (dtype == WF_F16) {
        if (n % 8) return WF_ERR_BAD_SHAPE;
        wf::gelu_inplace_kernel<__half><<<(unsigned)((n / 8 + 255) / 256), 256, 0, st>>>((__half *)x, n / 8);
    } else {
        return WF_ERR_BAD_DTYPE;
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_residual_sum(const float *a, const float *b, const void *c, int c_dtype, const float *bias, float *out,
                               int64_t rows, int C, void *stream) {
    if (!a || !b || !c || !out) return WF_ERR_NULL_POINTER;
    if (rows <= 0 || C <= 0 || C % 4) return WF_ERR_BAD_SHAPE;
    if (!wf::aligned16(a) || !wf::aligned16(b) || !wf::aligned16(out) || (reinterpret_cast<uintptr_t>(c) & 7u) ||
        (bias && !wf::aligned16(bias)))
        return WF_ERR_MISALIGNED;
    const int64_t total4 = rows * (C / 4);
    const unsigned grid = (unsigned)((total4 + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (c_dtype == WF_F32)
        wf::residual_sum_kernel<float><<<grid, 256, 0, st>>>(a, b, (const float *)c, bias, out, total4, C / 4);
    else if (c_dtype == WF_BF16)
        wf::residual_sum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, b, (const __nv_bfloat16 *)c, bias, out, total4, C / 4);
    else if (c_dtype == WF_F16)
        wf::residual_sum_kernel<__half><<<grid, 256, 0, st>>>(a, b, (const __half *)c, bias, out, total4, C / 4);
    else
        return WF_ERR_BAD_DTYPE;
    WF_LAUNCH_CHECK();
    return WF_OK;
}
