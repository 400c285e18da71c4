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
        vc = make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                         __uint_as_float(t.y & 0xffff0000u));
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

}  // namespace wf

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
    else
        return WF_ERR_BAD_DTYPE;
    WF_LAUNCH_CHECK();
    return WF_OK;
}
