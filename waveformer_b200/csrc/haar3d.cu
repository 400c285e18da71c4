// Kernel groups 1 and 3: one-level 3D Haar analysis / synthesis (sm_100a).
//
// Replaces ptwt.wavedec3 / ptwt.waverec3 at reference network_models/wave_helper.py:350 and
// network_models/idwt_upsample.py:160 (plus the layout copies and the torch.cat around them, see
// include/waveformer_b200.h).  Both directions are pure streaming: every input byte is read once and every output byte
// written once with 16-byte evict-first accesses, so the bound is HBM bandwidth (2 * N * sizeof(T) bytes per call).
//
// Two layouts, each with a 16-byte vector kernel and a scalar fallback for shapes the vector path cannot take:
//   NCDHW  : a thread turns 4 input rows x 8 samples into 4 coefficients of each of the 8 sub-bands.
//   NDHWC  : a thread owns one 2x2x2 cell x one 16-byte channel packet (4 fp32 / 8 bf16 channels).
#include <type_traits>

#include "wf_common.cuh"

namespace wf {

template <typename T> struct Quad;  // 4 consecutive elements of T (the NCDHW kernels' output granule)
template <> struct Quad<float> {
    using raw = float4;
    __device__ static inline raw pack(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
    __device__ static inline void unpack(const raw &r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <> struct Quad<__nv_bfloat16> {
    using raw = uint2;
    __device__ static inline raw pack(const float (&v)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        return make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
    __device__ static inline void unpack(const raw &r, float (&v)[4]) {
        v[0] = __uint_as_float(r.x << 16);
        v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16);
        v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
};
template <> struct Quad<__half> {
    using raw = uint2;
    __device__ static inline raw pack(const float (&v)[4]) {
        __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        return make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
};

// 8 consecutive elements of T -> fp32
__device__ inline void ld8s(const float *p, float (&v)[8]) {
    float4 a = ld_stream<float4>(p), b = ld_stream<float4>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ inline void ld8s(const __nv_bfloat16 *p, float (&v)[8]) {
    uint4 a = ld_stream<uint4>(p);
    Pack<__nv_bfloat16>::unpack(a, v);
}
__device__ inline void st8s(float *p, const float (&v)[8]) {
    st_stream(p, make_float4(v[0], v[1], v[2], v[3]));
    st_stream(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ inline void st8s(__nv_bfloat16 *p, const float (&v)[8]) { st_stream(p, Pack<__nv_bfloat16>::pack(v)); }
__device__ inline void ld8s(const __half *p, float (&v)[8]) {
    uint4 a = ld_stream<uint4>(p);
    Pack<__half>::unpack(a, v);
}
__device__ inline void st8s(__half *p, const float (&v)[8]) { st_stream(p, Pack<__half>::pack(v)); }

// ------------------------------------------------------------------------------------------------ NCDHW -------
template <typename T>
__global__ void __launch_bounds__(256) dwt_ncdhw_vec_kernel(const T *__restrict__ x, T *__restrict__ ll,
                                                            T *__restrict__ hf, int64_t total, int d, int h, int w,
                                                            int64_t band_stride) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int wq = w >> 2;  // quads per output row
    const int xq = (int)(idx % wq);
    const int64_t row = idx / wq;  // (b*d + z)*h + y
    const int y = (int)(row % h);
    const int64_t t = row / h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const T *r00 = x + (((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 8 * xq);
    float f[4][8];
    ld8s(r00, f[0]);
    ld8s(r00 + W, f[1]);
    ld8s(r00 + (int64_t)H * W, f[2]);
    ld8s(r00 + (int64_t)H * W + W, f[3]);
    float out[8][4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float v[8] = {f[0][2 * o], f[0][2 * o + 1], f[1][2 * o], f[1][2 * o + 1],
                            f[2][2 * o], f[2][2 * o + 1], f[3][2 * o], f[3][2 * o + 1]};
        float c[8];
        haar8(v, c);
#pragma unroll
        for (int n = 0; n < 8; ++n) out[n][o] = c[n];
    }
    const int64_t off = row * w + 4 * xq;
    st_stream(ll + off, Quad<T>::pack(out[0]));
    if (hf != nullptr) {
#pragma unroll
        for (int k = 0; k < 7; ++k) st_stream(hf + k * band_stride + off, Quad<T>::pack(out[k + 1]));
    }
}

// Raw global words: BYTES = 8, 16 or 32 per thread and access, streaming (evict-first) both ways.
template <int BYTES> struct Words { uint32_t w[BYTES / 4]; };
template <int BYTES> __device__ inline Words<BYTES> ld_words(const void *p) {
    Words<BYTES> r;
    if constexpr (BYTES == 8) {
        const uint2 t = ld_stream<uint2>(p);
        r.w[0] = t.x; r.w[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) {
            const uint4 t = ld_stream<uint4>(reinterpret_cast<const uint4 *>(p) + i);
            r.w[4 * i] = t.x; r.w[4 * i + 1] = t.y; r.w[4 * i + 2] = t.z; r.w[4 * i + 3] = t.w;
        }
    }
    return r;
}
template <int BYTES> __device__ inline void st_words(void *p, const Words<BYTES> &r) {
    if constexpr (BYTES == 8) {
        st_stream(p, make_uint2(r.w[0], r.w[1]));
    } else {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i)
            st_stream(reinterpret_cast<uint4 *>(p) + i, make_uint4(r.w[4 * i], r.w[4 * i + 1], r.w[4 * i + 2], r.w[4 * i + 3]));
    }
}
template <typename T, int N> __device__ inline void words_to_f32(const Words<N * (int)sizeof(T)> &r, float (&v)[N]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r.w[i]);
    } else if constexpr (std::is_same<T, __half>::value) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r.w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            v[2 * i] = __uint_as_float(r.w[i] << 16);
            v[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
        }
    }
}
template <typename T, int N> __device__ inline Words<N * (int)sizeof(T)> f32_to_words(const float (&v)[N]) {
    Words<N * (int)sizeof(T)> r;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N; ++i) r.w[i] = __float_as_uint(v[i]);
    } else if constexpr (std::is_same<T, __half>::value) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            r.w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            r.w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
    }
    return r;
}

// One thread = CO consecutive coefficients of every sub-band -> 4 output rows x 2*CO samples.  The eight (fifteen with
// a gate) sub-band loads are issued back to back before the first use, so each thread keeps 8 x CO x sizeof(T) bytes
// (128 B at CO = 16 / sizeof(T)) in flight; HF / GATE are compile-time so nothing blocks the hoisting.
template <typename T, int CO, bool HF, bool GATE>
__global__ void __launch_bounds__(256) idwt_ncdhw_vec_kernel(const T *__restrict__ ll, const T *__restrict__ hf,
                                                             const T *__restrict__ gate, T *__restrict__ x,
                                                             int64_t total, int d, int h, int w, int64_t band_stride) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    constexpr int NB = CO * (int)sizeof(T);
    const int wq = w / CO;
    const int xq = (int)(idx % wq);
    const int64_t row = idx / wq;
    const int y = (int)(row % h);
    const int64_t t = row / h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const int64_t off = row * w + CO * xq;
    Words<NB> raw[8], graw[7];
    raw[0] = ld_words<NB>(ll + off);
    if constexpr (HF) {
#pragma unroll
        for (int k = 0; k < 7; ++k) raw[k + 1] = ld_words<NB>(hf + k * band_stride + off);
        if constexpr (GATE) {
#pragma unroll
            for (int k = 0; k < 7; ++k) graw[k] = ld_words<NB>(gate + k * band_stride + off);
        }
    }
    float cin[8][CO];
    words_to_f32<T, CO>(raw[0], cin[0]);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if constexpr (HF) {
            words_to_f32<T, CO>(raw[k + 1], cin[k + 1]);
            if constexpr (GATE) {
                float g[CO];
                words_to_f32<T, CO>(graw[k], g);
#pragma unroll
                for (int o = 0; o < CO; ++o) cin[k + 1][o] *= g[o];
            }
        } else {
#pragma unroll
            for (int o = 0; o < CO; ++o) cin[k + 1][o] = 0.f;
        }
    }
    float f[4][2 * CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        const float c[8] = {cin[0][o], cin[1][o], cin[2][o], cin[3][o], cin[4][o], cin[5][o], cin[6][o], cin[7][o]};
        float v[8];
        haar8(c, v);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            f[r][2 * o] = v[2 * r];
            f[r][2 * o + 1] = v[2 * r + 1];
        }
    }
    T *r00 = x + (((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 2 * CO * xq);
    st_words<2 * NB>(r00, f32_to_words<T, 2 * CO>(f[0]));
    st_words<2 * NB>(r00 + W, f32_to_words<T, 2 * CO>(f[1]));
    st_words<2 * NB>(r00 + (int64_t)H * W, f32_to_words<T, 2 * CO>(f[2]));
    st_words<2 * NB>(r00 + (int64_t)H * W + W, f32_to_words<T, 2 * CO>(f[3]));
}

// scalar fallbacks: one thread per 2x2x2 cell
template <typename T>
__global__ void dwt_ncdhw_scalar_kernel(const T *__restrict__ x, T *__restrict__ ll, T *__restrict__ hf, int64_t total,
                                        int d, int h, int w, int64_t band_stride) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int xx = (int)(idx % w);
    const int64_t row = idx / w;
    const int y = (int)(row % h);
    const int64_t t = row / h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const T *p = x + (((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 2 * xx);
    const int64_t pl = (int64_t)H * W;
    const float v[8] = {to_f32(p[0]),      to_f32(p[1]),      to_f32(p[W]),      to_f32(p[W + 1]),
                        to_f32(p[pl]),     to_f32(p[pl + 1]), to_f32(p[pl + W]), to_f32(p[pl + W + 1])};
    float c[8];
    haar8(v, c);
    ll[idx] = from_f32<T>(c[0]);
    if (hf != nullptr)
        for (int k = 0; k < 7; ++k) hf[k * band_stride + idx] = from_f32<T>(c[k + 1]);
}

template <typename T>
__global__ void idwt_ncdhw_scalar_kernel(const T *__restrict__ ll, const T *__restrict__ hf, const T *__restrict__ gate,
                                         T *__restrict__ x, int64_t total, int d, int h, int w, int64_t band_stride) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int xx = (int)(idx % w);
    const int64_t row = idx / w;
    const int y = (int)(row % h);
    const int64_t t = row / h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    float c[8];
    c[0] = to_f32(ll[idx]);
    for (int k = 0; k < 7; ++k) {
        float v = hf ? to_f32(hf[k * band_stride + idx]) : 0.f;
        if (hf && gate) v *= to_f32(gate[k * band_stride + idx]);
        c[k + 1] = v;
    }
    float v[8];
    haar8(c, v);
    T *p = x + (((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 2 * xx);
    const int64_t pl = (int64_t)H * W;
    p[0] = from_f32<T>(v[0]); p[1] = from_f32<T>(v[1]); p[W] = from_f32<T>(v[2]); p[W + 1] = from_f32<T>(v[3]);
    p[pl] = from_f32<T>(v[4]); p[pl + 1] = from_f32<T>(v[5]); p[pl + W] = from_f32<T>(v[6]); p[pl + W + 1] = from_f32<T>(v[7]);
}

// ------------------------------------------------------------------------------------------------ NDHWC -------
// VEC = Pack<T>::VEC channels per thread (16-byte packets) or 1 (scalar fallback).
template <typename T, int VEC> struct ChanIO {
    __device__ static inline void load(const T *p, float (&v)[VEC]) {
        if constexpr (VEC == 1) {
            v[0] = to_f32(*p);
        } else {
            Pack<T>::unpack(ld_stream<typename Pack<T>::raw>(p), v);
        }
    }
    __device__ static inline void store(T *p, const float (&v)[VEC]) {
        if constexpr (VEC == 1) {
            *p = from_f32<T>(v[0]);
        } else {
            st_stream(p, Pack<T>::pack(v));
        }
    }
};

// THF: storage type of the seven detail bands (may be bf16 while x / LL stay fp32: the encoder's residual-stream
// precision for the attention input, the decoder's activation type for the details).
template <typename T, typename THF, int VEC>
__global__ void __launch_bounds__(256) dwt_ndhwc_kernel(const T *__restrict__ x, T *__restrict__ ll, THF *__restrict__ hf,
                                                        int64_t total, int d, int h, int w, int cchunks,
                                                        int64_t xs, int64_t lls, int64_t band_stride, int C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cc = (int)(idx % cchunks);
    const int64_t vox = idx / cchunks;  // ((b*d + z)*h + y)*w + x
    const int xx = (int)(vox % w);
    int64_t t = vox / w;
    const int y = (int)(t % h);
    t /= h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    const int64_t v000 = ((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 2 * xx;
    const T *p = x + v000 * xs + cc * VEC;
    const int64_t sx = xs, sy = (int64_t)W * xs, sz = (int64_t)H * W * xs;
    float f[8][VEC];
#pragma unroll
    for (int m = 0; m < 8; ++m)
        ChanIO<T, VEC>::load(p + ((m >> 2) & 1) * sz + ((m >> 1) & 1) * sy + (m & 1) * sx, f[m]);
    float o[8][VEC];
#pragma unroll
    for (int ch = 0; ch < VEC; ++ch) {
        const float v[8] = {f[0][ch], f[1][ch], f[2][ch], f[3][ch], f[4][ch], f[5][ch], f[6][ch], f[7][ch]};
        float c[8];
        haar8(v, c);
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n][ch] = c[n];
    }
    ChanIO<T, VEC>::store(ll + vox * lls + cc * VEC, o[0]);
    if (hf != nullptr) {
        THF *q = hf + vox * C + cc * VEC;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if constexpr (VEC == 1 || sizeof(THF) == sizeof(T)) {
                ChanIO<THF, VEC>::store(q + k * band_stride, o[k + 1]);
            } else {  // fp32 packet of 4 channels -> 4 sixteen-bit values (8 bytes)
                static_assert(VEC == 4, "mixed storage: fp32 in, 16-bit details");
                st_stream(q + k * band_stride, Quad<THF>::pack(o[k + 1]));
            }
        }
    }
}

template <typename T, int VEC, bool HF, bool GATE>
__global__ void __launch_bounds__(256) idwt_ndhwc_kernel(const T *__restrict__ ll, const T *__restrict__ hf,
                                                         const T *__restrict__ gate, T *__restrict__ x, int64_t total,
                                                         int d, int h, int w, int cchunks, int64_t lls,
                                                         int64_t band_stride, int64_t xs, int C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cc = (int)(idx % cchunks);
    const int64_t vox = idx / cchunks;
    const int xx = (int)(vox % w);
    int64_t t = vox / w;
    const int y = (int)(t % h);
    t /= h;
    const int z = (int)(t % d);
    const int64_t b = t / d;
    const int H = 2 * h, W = 2 * w;
    // every load first (compile-time HF / GATE: no branch between them), then the butterflies
    float cin[8][VEC], g[7][VEC];
    ChanIO<T, VEC>::load(ll + vox * lls + cc * VEC, cin[0]);
    if constexpr (HF) {
        const T *q = hf + vox * C + cc * VEC;
#pragma unroll
        for (int k = 0; k < 7; ++k) ChanIO<T, VEC>::load(q + k * band_stride, cin[k + 1]);
        if constexpr (GATE) {
            const T *gq = gate + vox * C + cc * VEC;
#pragma unroll
            for (int k = 0; k < 7; ++k) ChanIO<T, VEC>::load(gq + k * band_stride, g[k]);
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int ch = 0; ch < VEC; ++ch) cin[k + 1][ch] *= g[k][ch];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int ch = 0; ch < VEC; ++ch) cin[k + 1][ch] = 0.f;
    }
    float o[8][VEC];
#pragma unroll
    for (int ch = 0; ch < VEC; ++ch) {
        const float c[8] = {cin[0][ch], cin[1][ch], cin[2][ch], cin[3][ch], cin[4][ch], cin[5][ch], cin[6][ch], cin[7][ch]};
        float v[8];
        haar8(c, v);
#pragma unroll
        for (int m = 0; m < 8; ++m) o[m][ch] = v[m];
    }
    const int64_t v000 = ((b * (2 * d) + 2 * z) * H + 2 * y) * (int64_t)W + 2 * xx;
    T *p = x + v000 * xs + cc * VEC;
    const int64_t sx = xs, sy = (int64_t)W * xs, sz = (int64_t)H * W * xs;
#pragma unroll
    for (int m = 0; m < 8; ++m)
        ChanIO<T, VEC>::store(p + ((m >> 2) & 1) * sz + ((m >> 1) & 1) * sy + (m & 1) * sx, o[m]);
}

// ------------------------------------------------------------------------------------------------ hosts -------
static inline int grid_for(int64_t total, int block) { return (int)((total + block - 1) / block); }

template <typename T>
static int dwt_ncdhw_launch(const T *x, T *ll, T *hf, int64_t n, int D, int H, int W, int64_t bs, cudaStream_t st) {
    const int d = D / 2, h = H / 2, w = W / 2;
    const bool vec = (W % 8 == 0) && aligned16(x) && aligned16(ll) && (hf == nullptr || (aligned16(hf) && (bs * sizeof(T)) % 16 == 0));
    if (vec) {
        const int64_t total = n * d * h * (w / 4);
        dwt_ncdhw_vec_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(x, ll, hf, total, d, h, w, bs);
    } else {
        const int64_t total = n * d * h * w;
        dwt_ncdhw_scalar_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(x, ll, hf, total, d, h, w, bs);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T>
static int idwt_ncdhw_launch(const T *ll, const T *hf, const T *gate, T *x, int64_t n, int d, int h, int w, int64_t bs,
                             cudaStream_t st) {
    constexpr int CO = 16 / (int)sizeof(T);  // 16-byte sub-band loads: 4 fp32 / 8 bf16 coefficients per thread
    const bool vec = (w % CO == 0) && aligned16(x) && aligned16(ll) &&
                     (hf == nullptr || (aligned16(hf) && (bs * sizeof(T)) % 16 == 0)) && (gate == nullptr || aligned16(gate));
    if (gate != nullptr && hf == nullptr) return WF_ERR_NULL_POINTER;
    if (vec) {
        const int64_t total = n * d * h * (w / CO);
        const int grid = grid_for(total, 256);
        if (hf == nullptr)
            idwt_ncdhw_vec_kernel<T, CO, false, false><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, bs);
        else if (gate == nullptr)
            idwt_ncdhw_vec_kernel<T, CO, true, false><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, bs);
        else
            idwt_ncdhw_vec_kernel<T, CO, true, true><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, bs);
    } else {
        const int64_t total = n * d * h * w;
        idwt_ncdhw_scalar_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, bs);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T, typename THF>
static int dwt_ndhwc_launch(const T *x, T *ll, THF *hf, int B, int D, int H, int W, int C, int64_t xs, int64_t lls,
                            int64_t bs, cudaStream_t st) {
    const int d = D / 2, h = H / 2, w = W / 2;
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(ll) && (xs * e) % 16 == 0 && (lls * e) % 16 == 0 &&
                     (hf == nullptr || (aligned16(hf) && (bs * sizeof(THF)) % 16 == 0 && (C * sizeof(THF)) % 8 == 0));
    const int64_t vox = (int64_t)B * d * h * w;
    if (vec) {
        const int64_t total = vox * (C / V);
        dwt_ndhwc_kernel<T, THF, V><<<grid_for(total, 256), 256, 0, st>>>(x, ll, hf, total, d, h, w, C / V, xs, lls, bs, C);
    } else {
        const int64_t total = vox * C;
        dwt_ndhwc_kernel<T, THF, 1><<<grid_for(total, 256), 256, 0, st>>>(x, ll, hf, total, d, h, w, C, xs, lls, bs, C);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T, int V>
static void idwt_ndhwc_dispatch(const T *ll, const T *hf, const T *gate, T *x, int64_t total, int d, int h, int w,
                                int cchunks, int64_t lls, int64_t bs, int64_t xs, int C, cudaStream_t st) {
    const int grid = grid_for(total, 256);
    if (hf == nullptr)
        idwt_ndhwc_kernel<T, V, false, false><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, cchunks, lls, bs, xs, C);
    else if (gate == nullptr)
        idwt_ndhwc_kernel<T, V, true, false><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, cchunks, lls, bs, xs, C);
    else
        idwt_ndhwc_kernel<T, V, true, true><<<grid, 256, 0, st>>>(ll, hf, gate, x, total, d, h, w, cchunks, lls, bs, xs, C);
}

template <typename T>
static int idwt_ndhwc_launch(const T *ll, const T *hf, const T *gate, T *x, int B, int d, int h, int w, int C,
                             int64_t lls, int64_t bs, int64_t xs, cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    if (gate != nullptr && hf == nullptr) return WF_ERR_NULL_POINTER;
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(ll) && (xs * e) % 16 == 0 && (lls * e) % 16 == 0 &&
                     (hf == nullptr || (aligned16(hf) && (bs * e) % 16 == 0)) && (gate == nullptr || aligned16(gate));
    const int64_t vox = (int64_t)B * d * h * w;
    if (vec)
        idwt_ndhwc_dispatch<T, V>(ll, hf, gate, x, vox * (C / V), d, h, w, C / V, lls, bs, xs, C, st);
    else
        idwt_ndhwc_dispatch<T, 1>(ll, hf, gate, x, vox * C, d, h, w, C, lls, bs, xs, C, st);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

using namespace wf;

extern "C" int wf_dwt3d_ncdhw(const void *x, void *ll, void *hf, int dtype, int64_t n, int D, int H, int W,
                              int64_t hf_band_stride, void *stream) {
    if (!x || !ll) return WF_ERR_NULL_POINTER;
    if (n <= 0 || D <= 0 || H <= 0 || W <= 0 || (D | H | W) & 1) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32) return dwt_ncdhw_launch<float>((const float *)x, (float *)ll, (float *)hf, n, D, H, W, hf_band_stride, st);
    if (dtype == WF_BF16)
        return dwt_ncdhw_launch<__nv_bfloat16>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)ll, (__nv_bfloat16 *)hf, n, D, H, W, hf_band_stride, st);
    if (dtype == WF_F16)
        return dwt_ncdhw_launch<__half>((const __half *)x, (__half *)ll, (__half *)hf, n, D, H, W, hf_band_stride, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_idwt3d_ncdhw(const void *ll, const void *hf, const void *gate, void *x, int dtype, int64_t n, int d,
                               int h, int w, int64_t hf_band_stride, void *stream) {
    if (!x || !ll) return WF_ERR_NULL_POINTER;
    if (n <= 0 || d <= 0 || h <= 0 || w <= 0) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32)
        return idwt_ncdhw_launch<float>((const float *)ll, (const float *)hf, (const float *)gate, (float *)x, n, d, h, w, hf_band_stride, st);
    if (dtype == WF_BF16)
        return idwt_ncdhw_launch<__nv_bfloat16>((const __nv_bfloat16 *)ll, (const __nv_bfloat16 *)hf, (const __nv_bfloat16 *)gate,
                                                (__nv_bfloat16 *)x, n, d, h, w, hf_band_stride, st);
    if (dtype == WF_F16)
        return idwt_ncdhw_launch<__half>((const __half *)ll, (const __half *)hf, (const __half *)gate, (__half *)x, n, d, h, w,
                                         hf_band_stride, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_dwt3d_ndhwc(const void *x, void *ll, void *hf, int dtype, int hf_dtype, int B, int D, int H, int W,
                              int C, int64_t x_vox_stride, int64_t ll_vox_stride, int64_t hf_band_stride, void *stream) {
    if (!x || !ll) return WF_ERR_NULL_POINTER;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || (D | H | W) & 1) return WF_ERR_BAD_SHAPE;
    if (x_vox_stride < C || ll_vox_stride < C) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    if (dtype == WF_F32 && hf_dtype == WF_F32)
        return dwt_ndhwc_launch<float, float>((const float *)x, (float *)ll, (float *)hf, B, D, H, W, C, x_vox_stride, ll_vox_stride, hf_band_stride, st);
    if (dtype == WF_F32 && hf_dtype == WF_BF16)
        return dwt_ndhwc_launch<float, bf>((const float *)x, (float *)ll, (bf *)hf, B, D, H, W, C, x_vox_stride, ll_vox_stride, hf_band_stride, st);
    if (dtype == WF_BF16 && hf_dtype == WF_BF16)
        return dwt_ndhwc_launch<bf, bf>((const bf *)x, (bf *)ll, (bf *)hf, B, D, H, W, C, x_vox_stride, ll_vox_stride, hf_band_stride, st);
    if (dtype == WF_F32 && hf_dtype == WF_F16)
        return dwt_ndhwc_launch<float, __half>((const float *)x, (float *)ll, (__half *)hf, B, D, H, W, C, x_vox_stride, ll_vox_stride, hf_band_stride, st);
    if (dtype == WF_F16 && hf_dtype == WF_F16)
        return dwt_ndhwc_launch<__half, __half>((const __half *)x, (__half *)ll, (__half *)hf, B, D, H, W, C, x_vox_stride, ll_vox_stride, hf_band_stride, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_idwt3d_ndhwc(const void *ll, const void *hf, const void *gate, void *x, int dtype, int B, int d, int h,
                               int w, int C, int64_t ll_vox_stride, int64_t hf_band_stride, int64_t x_vox_stride,
                               void *stream) {
    if (!x || !ll) return WF_ERR_NULL_POINTER;
    if (B <= 0 || C <= 0 || d <= 0 || h <= 0 || w <= 0) return WF_ERR_BAD_SHAPE;
    if (x_vox_stride < C || ll_vox_stride < C) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32)
        return idwt_ndhwc_launch<float>((const float *)ll, (const float *)hf, (const float *)gate, (float *)x, B, d, h, w, C,
                                        ll_vox_stride, hf_band_stride, x_vox_stride, st);
    if (dtype == WF_BF16)
        return idwt_ndhwc_launch<__nv_bfloat16>((const __nv_bfloat16 *)ll, (const __nv_bfloat16 *)hf, (const __nv_bfloat16 *)gate,
                                                (__nv_bfloat16 *)x, B, d, h, w, C, ll_vox_stride, hf_band_stride, x_vox_stride, st);
    if (dtype == WF_F16)
        return idwt_ndhwc_launch<__half>((const __half *)ll, (const __half *)hf, (const __half *)gate, (__half *)x, B, d, h, w, C,
                                         ll_vox_stride, hf_band_stride, x_vox_stride, st);
    return WF_ERR_BAD_DTYPE;
}
