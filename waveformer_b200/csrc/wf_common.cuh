// Shared device/host helpers for the waveformer_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>

#include "../../include/waveformer_b200.h"

namespace wf {

extern int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return WF_ERR_CUDA;
}

#define WF_CUDA_CHECK(expr)                                      \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return ::wf::cuda_fail(_e);       \
    } while (0)

// cudaGetLastError (not Peek): a failed launch must not poison every later call of the library
#define WF_LAUNCH_CHECK() WF_CUDA_CHECK(cudaGetLastError())

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200

// WF_AB_OLD=1 routes the kernels that have a newer variant (persistent convT, 4-voxel patch embedding, 2-voxel output head, z-walking
// upsample) back to their predecessors: same-box A/B timing only (the boxes differ by +-3 % in the clocks their power cap allows).
inline bool ab_old() {
    static const bool v = getenv("WF_AB_OLD") != nullptr && getenv("WF_AB_OLD")[0] == '1';
    return v;
}

// Function attributes (the > 48 KB dynamic shared-memory opt-in) are per DEVICE: one flag word per call site, one bit per
// device ordinal, so a process that drives several GPUs opts in on each of them.  Returns true the first time the calling
// site runs on the current device.
inline bool first_use_on_current_device(unsigned long long &mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;   // unknown ordinal: just opt in again
    const unsigned long long bit = 1ull << dev;
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// ---- 16-byte packets of activations ---------------------------------------------------------------------------
template <typename T> struct Pack;  // VEC elements of T in one 16-byte global transaction
template <> struct Pack<float> {
    static constexpr int VEC = 4;
    using raw = float4;
    __device__ static inline void unpack(const raw &r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
    __device__ static inline raw pack(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Pack<__nv_bfloat16> {
    static constexpr int VEC = 8;
    using raw = uint4;
    __device__ static inline void unpack(const raw &r, float (&v)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static inline raw pack(const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <> struct Pack<__half> {   // fp16 activations of the skip blocks (precision policy: 10-bit mantissa at bf16 cost)
    static constexpr int VEC = 8;
    using raw = uint4;
    __device__ static inline void unpack(const raw &r, float (&v)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    __device__ static inline raw pack(const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// 4 / 8 consecutive elements of T <-> fp32 registers with plain (cached) vector accesses: fp32 16 / 2 x 16 bytes,
// 16-bit types 8 / 16 bytes
template <typename T> __device__ inline void load4(const T *p, float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const uint2 t = *reinterpret_cast<const uint2 *>(p);
        float f[8];
        Pack<T>::unpack(make_uint4(t.x, t.y, 0u, 0u), f);
        v[0] = f[0]; v[1] = f[1]; v[2] = f[2]; v[3] = f[3];
    }
}
template <typename T> __device__ inline void store4(T *p, const float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        const float f[8] = {v[0], v[1], v[2], v[3], 0.f, 0.f, 0.f, 0.f};
        const uint4 t = Pack<T>::pack(f);
        *reinterpret_cast<uint2 *>(p) = make_uint2(t.x, t.y);
    }
}
template <typename T> __device__ inline void load8(const T *p, float (&v)[8]) {
    if constexpr (sizeof(T) == 4) {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        Pack<T>::unpack(*reinterpret_cast<const uint4 *>(p), v);
    }
}
template <typename T> __device__ inline void store8(T *p, const float (&v)[8]) {
    if constexpr (sizeof(T) == 4) {
        reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        *reinterpret_cast<uint4 *>(p) = Pack<T>::pack(v);
    }
}

// streaming (evict-first) 16-byte global accesses: every byte of the Haar kernels is touched exactly once
template <typename R> __device__ inline R ld_stream(const void *p) { return __ldcs(reinterpret_cast<const R *>(p)); }
template <typename R> __device__ inline void st_stream(void *p, const R &v) { __stcs(reinterpret_cast<R *>(p), v); }

template <typename T> __device__ inline float to_f32(T v);
template <> __device__ inline float to_f32<float>(float v) { return v; }
template <> __device__ inline float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ inline float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ inline T from_f32(float v);
template <> __device__ inline __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ inline float from_f32<float>(float v) { return v; }
template <> __device__ inline __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- the 2x2x2 Haar butterfly (orthonormal, symmetric => the same routine analyses and synthesises) -----------
// in : v[m], m = (i<<2)|(j<<1)|k  = sample at (2z+i, 2y+j, 2x+k)
// out: c[n], n = (p<<2)|(q<<1)|r  = sub-band (D-kind p, H-kind q, W-kind r): aaa,aad,ada,add,daa,dad,dda,ddd
__device__ inline void haar8(const float (&v)[8], float (&c)[8]) {
    const float s = 0.35355339059327378f;  // 1 / (2*sqrt(2))
    // W axis (bit 0)
    const float a00 = v[0] + v[1], d00 = v[0] - v[1];
    const float a01 = v[2] + v[3], d01 = v[2] - v[3];
    const float a10 = v[4] + v[5], d10 = v[4] - v[5];
    const float a11 = v[6] + v[7], d11 = v[6] - v[7];
    // H axis (bit 1)
    const float aa0 = a00 + a01, da0 = a00 - a01, ad0 = d00 + d01, dd0 = d00 - d01;
    const float aa1 = a10 + a11, da1 = a10 - a11, ad1 = d10 + d11, dd1 = d10 - d11;
    // D axis (bit 2)
    c[0] = s * (aa0 + aa1);
    c[1] = s * (ad0 + ad1);
    c[2] = s * (da0 + da1);
    c[3] = s * (dd0 + dd1);
    c[4] = s * (aa0 - aa1);
    c[5] = s * (ad0 - ad1);
    c[6] = s * (da0 - da1);
    c[7] = s * (dd0 - dd1);
}

}  // namespace wf
