// Depthwise 3x3x3 convolution on channels-last activations (sm_100a).
//
// Used by CCF_FFN.dwconv (reference network_models/wave_helper.py:231-232,283: Conv3d(hid, hid, 3, padding=1,
// groups=hid)) and ProjectionUpsample.conv1[1] (wave_helper.py:44).  cuDNN executes these grouped convolutions one
// group at a time (2976 sgemm launches per forward, 65 % of the step); as a stencil it is a bandwidth problem:
// each thread owns one 16-byte channel packet and TX consecutive voxels along W, walks the 9 (dz, dy) input rows, and
// keeps TX accumulators per channel, so an input row is loaded once per TX outputs.  Neighbouring rows / planes are
// served by L1/L2 (a three-plane working set is a few MB).
//
// x, y: [B, D, H, W, C] channels-last, dense; w27: fp32 [27][C] (tap-major repack of the [C,1,3,3,3] weight);
// bias: fp32 [C] or NULL.  Accumulation in fp32.
#include "wf_common.cuh"

namespace wf {

template <typename T, int VEC> struct CVec {
    __device__ static inline void load(const T *p, float (&v)[VEC]) {
        if constexpr (VEC == 1) {
            v[0] = to_f32(__ldg(p));
        } else {
            Pack<T>::unpack(__ldg(reinterpret_cast<const typename Pack<T>::raw *>(p)), v);
        }
    }
    __device__ static inline void store(T *p, const float (&v)[VEC]) {
        if constexpr (VEC == 1) {
            *p = from_f32<T>(v[0]);
        } else {
            *reinterpret_cast<typename Pack<T>::raw *>(p) = Pack<T>::pack(v);
        }
    }
};

template <int VEC> __device__ inline void load_w(const float *p, float (&w)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + i);
            w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) w[i] = __ldg(p + i);
    }
}

template <typename T, int VEC, int TX>
__global__ void __launch_bounds__(256) dwconv3d_ndhwc_kernel(const T *__restrict__ x, const float *__restrict__ w27,
                                                             const float *__restrict__ bias, T *__restrict__ y,
                                                             int64_t total, int D, int H, int W, int C, int cvecs,
                                                             int xtiles) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    int64_t t = idx / cvecs;
    const int xt = (int)(t % xtiles); t /= xtiles;
    const int yy0 = (int)(t % H); t /= H;
    const int zz0 = (int)(t % D);
    const int64_t b = t / D;
    const int x0 = xt * TX;
    const int c0 = cv * VEC;
    float acc[TX][VEC];
    {
        float bv[VEC];
        if (bias != nullptr) load_w<VEC>(bias + c0, bv);
#pragma unroll
        for (int o = 0; o < TX; ++o)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[o][e] = bias != nullptr ? bv[e] : 0.f;
    }
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        const int zz = zz0 + dz;
        if ((unsigned)zz >= (unsigned)D) continue;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = yy0 + dy;
            if ((unsigned)yy >= (unsigned)H) continue;
            const T *row = x + (((b * D + zz) * H + yy) * (int64_t)W) * C + c0;
            float in[TX + 2][VEC];
#pragma unroll
            for (int i = 0; i < TX + 2; ++i) {
                const int xx = x0 - 1 + i;
                if ((unsigned)xx < (unsigned)W) {
                    CVec<T, VEC>::load(row + (int64_t)xx * C, in[i]);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) in[i][e] = 0.f;
                }
            }
            const float *wp = w27 + ((dz + 1) * 9 + (dy + 1) * 3) * (int64_t)C + c0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float wk[VEC];
                load_w<VEC>(wp + (int64_t)k * C, wk);
#pragma unroll
                for (int o = 0; o < TX; ++o)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[o][e] = fmaf(in[o + k][e], wk[e], acc[o][e]);
            }
        }
    }
    T *orow = y + (((b * D + zz0) * H + yy0) * (int64_t)W) * C + c0;
#pragma unroll
    for (int o = 0; o < TX; ++o)
        if (x0 + o < W) CVec<T, VEC>::store(orow + (int64_t)(x0 + o) * C, acc[o]);
}

template <typename T>
static int dwconv_launch(const T *x, const float *w27, const float *bias, T *y, int B, int D, int H, int W, int C,
                         cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    constexpr int TX = 4;
    const int xtiles = (W + TX - 1) / TX;
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(y) && aligned16(w27) && (bias == nullptr || aligned16(bias));
    if (vec) {
        const int cvecs = C / V;
        const int64_t total = (int64_t)B * D * H * xtiles * cvecs;
        dwconv3d_ndhwc_kernel<T, V, TX><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, w27, bias, y, total, D, H, W, C, cvecs, xtiles);
    } else {
        const int64_t total = (int64_t)B * D * H * xtiles * C;
        dwconv3d_ndhwc_kernel<T, 1, TX><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, w27, bias, y, total, D, H, W, C, C, xtiles);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" int wf_dwconv3d_ndhwc(const void *x, const float *w27, const float *bias, void *y, int dtype, int B, int D,
                                 int H, int W, int C, void *stream) {
    if (!x || !w27 || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32) return wf::dwconv_launch<float>((const float *)x, w27, bias, (float *)y, B, D, H, W, C, st);
    if (dtype == WF_BF16)
        return wf::dwconv_launch<__nv_bfloat16>((const __nv_bfloat16 *)x, w27, bias, (__nv_bfloat16 *)y, B, D, H, W, C, st);
    return WF_ERR_BAD_DTYPE;
}
