// InstanceNorm3d (affine-free) on channels-last activations, fused with the activation and the residual add that
// follow it in the reference's conv blocks (sm_100a).
//
// Reference call sites: MONAI UnetResBlock.forward (monai/networks/blocks/dynunet_block.py:98-111: norm1 + lrelu,
// norm2, norm3, `out += residual`, lrelu) and ChannelCalibration.forward (network_models/network_backbone.py:118-128).
// torch's instance_norm copies channels-last tensors to NCDHW and back (35 % of the step before this kernel).
// Two streaming passes: per-(b, c) sum / sum-of-squares (fp32 per thread, fp64 across blocks, so E[x^2]-E[x]^2 does
// not cancel), then y = act((x - mean) * rstd [+ (r - mean_r) * rstd_r | + r]).
#include "wf_common.cuh"

namespace wf {

template <typename T, int VEC> struct NVec {
    __device__ static inline void load(const T *p, float (&v)[VEC]) {
        if constexpr (VEC == 1) v[0] = to_f32(*p);
        else Pack<T>::unpack(*reinterpret_cast<const typename Pack<T>::raw *>(p), v);
    }
    __device__ static inline void store(T *p, const float (&v)[VEC]) {
        if constexpr (VEC == 1) *p = from_f32<T>(v[0]);
        else *reinterpret_cast<typename Pack<T>::raw *>(p) = Pack<T>::pack(v);
    }
};

// grid (chunks, B); each block reduces `vox_per_block` voxels of one batch element for all channels.
// Sums are taken of (x - pivot) with pivot = the channel's value at voxel 0 of the batch element, so that
// E[d^2] - E[d]^2 stays well conditioned when |mean| >> std.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const T *__restrict__ x, double *__restrict__ sums,
                                                             int64_t S, int C, int cvecs, int64_t vox_per_block,
                                                             int64_t xs) {
    extern __shared__ float red[];  // [256][2*VEC]
    const int b = blockIdx.y;
    const int group = 256 / cvecs;  // voxels handled per sweep
    const int tid = threadIdx.x;
    const int cv = tid % cvecs;
    const int vg = tid / cvecs;
    const int64_t v0 = (int64_t)blockIdx.x * vox_per_block;
    const int64_t v1 = min(S, v0 + vox_per_block);
    float s[VEC], q[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) s[e] = q[e] = 0.f;
    if (vg < group) {
        const T *base = x + (int64_t)b * S * xs + cv * VEC;
        float piv[VEC];
        NVec<T, VEC>::load(base, piv);
        for (int64_t v = v0 + vg; v < v1; v += group) {
            float f[VEC];
            NVec<T, VEC>::load(base + v * xs, f);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float d = f[e] - piv[e];
                s[e] += d;
                q[e] = fmaf(d, d, q[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        red[tid * 2 * VEC + e] = s[e];
        red[tid * 2 * VEC + VEC + e] = q[e];
    }
    __syncthreads();
    // thread t < cvecs*2*VEC sums column t over the voxel groups
    const int ncol = cvecs * 2 * VEC;
    for (int col = tid; col < ncol; col += 256) {
        const int ccv = col / (2 * VEC), e2 = col % (2 * VEC);
        double a = 0.0;
        for (int g = 0; g < group; ++g) a += (double)red[(g * cvecs + ccv) * 2 * VEC + e2];
        const int c = ccv * VEC + (e2 % VEC);
        atomicAdd(sums + ((int64_t)b * C + c) * 2 + (e2 / VEC), a);
    }
}

template <typename T>
__global__ void instnorm_finalize_kernel(const T *__restrict__ x, const double *__restrict__ sums, float *__restrict__ mr,
                                         int n, int C, int64_t S, int64_t xs, double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // b*C + c
    if (i >= n) return;
    const int b = i / C, c = i % C;
    const double piv = (double)to_f32(x[(int64_t)b * S * xs + c]);
    const double inv_s = 1.0 / (double)S;
    const double md = sums[2 * i] * inv_s;  // mean of (x - pivot)
    double var = sums[2 * i + 1] * inv_s - md * md;
    var = var < 0.0 ? 0.0 : var;
    mr[2 * i] = (float)(piv + md);
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

// act: 0 none, 1 relu, 2 leaky relu (slope)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                             const T *__restrict__ res, const float *__restrict__ res_mr,
                                                             T *__restrict__ y, int64_t total, int64_t S, int C,
                                                             int cvecs, int64_t xs, int64_t rs, int64_t ys, int act,
                                                             float slope, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % cvecs);
    const int64_t vox = idx / cvecs;  // b*S + v
    const int64_t b = vox / S;
    const int c0 = cv * VEC;
    float f[VEC];
    NVec<T, VEC>::load(x + vox * xs + c0, f);
    const float *m = mr + ((int64_t)b * C + c0) * 2;
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = (f[e] - __ldg(m + 2 * e)) * __ldg(m + 2 * e + 1);
    if (gamma != nullptr) {  // GroupNorm(num_groups = C) = InstanceNorm + per-channel affine
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = fmaf(f[e], __ldg(gamma + c0 + e), beta != nullptr ? __ldg(beta + c0 + e) : 0.f);
    }
    if (res != nullptr) {
        float r[VEC];
        NVec<T, VEC>::load(res + vox * rs + c0, r);
        if (res_mr != nullptr) {
            const float *m2 = res_mr + ((int64_t)b * C + c0) * 2;
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] = (r[e] - __ldg(m2 + 2 * e)) * __ldg(m2 + 2 * e + 1);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] += r[e];
    }
    if (act == 1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = fmaxf(f[e], 0.f);
    } else if (act == 2) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * slope;
    }
    NVec<T, VEC>::store(y + vox * ys + c0, f);
}

// Last block of the network: y = act((x - mean) * rstd + [(r - mean_r) * rstd_r | r]) is consumed only by the 1x1x1
// output convolution (Waveformer.out, reference network_models/network_backbone.py:407 -> UnetOutBlock,
// monai/networks/blocks/dynunet_block.py:266), so the C-channel activation is never written: each lane normalises one
// 16-byte channel packet, multiplies it with its slice of the [K, C] head weight, the packets of a voxel are summed
// through shared memory and one thread per voxel stores the K logits.
//   block = 32 voxels x cpv packets (cpv = C / VEC), persistent over voxel groups; KP = padded K (4, 8 or 16).
template <typename T, typename TO, int KP>
__global__ void __launch_bounds__(1024) instnorm_apply_head_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                                   const T *__restrict__ res, const float *__restrict__ res_mr,
                                                                   const float *__restrict__ hw, const float *__restrict__ hb,
                                                                   TO *__restrict__ out, int64_t S, int64_t total_vox, int C,
                                                                   int K, int cpv, int64_t xs, int64_t rs, int act, float slope) {
    constexpr int VEC = Pack<T>::VEC;
    extern __shared__ float part[];  // [32 * cpv][KP]
    const int tid = threadIdx.x;
    const int vl = tid / cpv, cv = tid - vl * cpv;   // voxel in the group, channel packet
    const int c0 = cv * VEC;
    float w[KP][VEC];
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int e = 0; e < VEC; ++e) w[k][e] = k < K ? __ldg(hw + (int64_t)k * C + c0 + e) : 0.f;
    const int64_t groups = (total_vox + 31) / 32;
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t vox = g * 32 + vl;
        float p[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) p[k] = 0.f;
        if (vox < total_vox) {
            const int64_t b = vox / S;
            float f[VEC];
            NVec<T, VEC>::load(x + vox * xs + c0, f);
            const float *m = mr + ((int64_t)b * C + c0) * 2;
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[e] = (f[e] - __ldg(m + 2 * e)) * __ldg(m + 2 * e + 1);
            if (res != nullptr) {
                float r[VEC];
                NVec<T, VEC>::load(res + vox * rs + c0, r);
                if (res_mr != nullptr) {
                    const float *m2 = res_mr + ((int64_t)b * C + c0) * 2;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) r[e] = (r[e] - __ldg(m2 + 2 * e)) * __ldg(m2 + 2 * e + 1);
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] += r[e];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (act == 1) f[e] = fmaxf(f[e], 0.f);
                else if (act == 2) f[e] = f[e] > 0.f ? f[e] : f[e] * slope;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k)
#pragma unroll
                for (int e = 0; e < VEC; ++e) p[k] = fmaf(f[e], w[k][e], p[k]);
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) part[(size_t)tid * KP + k] = p[k];
        __syncthreads();
        if (tid < 32 && g * 32 + tid < total_vox) {
            float o[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) o[k] = (hb != nullptr && k < K) ? __ldg(hb + k) : 0.f;
            for (int j = 0; j < cpv; ++j)
#pragma unroll
                for (int k = 0; k < KP; ++k) o[k] += part[((size_t)tid * cpv + j) * KP + k];
            TO *dst = out + (g * 32 + tid) * (int64_t)K;
            for (int k = 0; k < K; ++k) dst[k] = from_f32<TO>(o[k]);
        }
        __syncthreads();
    }
}

template <typename T, typename TO>
static int apply_head_launch(const T *x, const float *mr, const T *res, const float *res_mr, const float *hw, const float *hb,
                             TO *out, int B, int64_t S, int C, int K, int64_t xs, int64_t rs, int act, float slope,
                             cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    if (C % V != 0 || C / V > 32 || K < 1 || K > 16) return WF_ERR_UNSUPPORTED;
    if (!aligned16(x) || (xs * e) % 16 != 0 || (res != nullptr && (!aligned16(res) || (rs * e) % 16 != 0))) return WF_ERR_MISALIGNED;
    const int cpv = C / V;
    const int threads = 32 * cpv;
    const int64_t total = (int64_t)B * S;
    const int64_t groups = (total + 31) / 32;
    const int per_sm = 2048 / threads > 0 ? 2048 / threads : 1;
    const unsigned grid = (unsigned)min(groups, (int64_t)kNumSMs * per_sm);
#define WF_HEAD(KP_) instnorm_apply_head_kernel<T, TO, KP_><<<grid, threads, (size_t)threads * KP_ * sizeof(float), st>>>( \
        x, mr, res, res_mr, hw, hb, out, S, total, C, K, cpv, xs, rs, act, slope)
    if (K <= 4) WF_HEAD(4);
    else if (K <= 8) WF_HEAD(8);
    else WF_HEAD(16);
#undef WF_HEAD
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T>
static int stats_launch(const T *x, double *sums, float *mr, int B, int64_t S, int C, int64_t xs, float eps, cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    WF_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * C, st));
    const bool vec = (C % V == 0) && (C / V <= 256) && aligned16(x) && (xs * sizeof(T)) % 16 == 0;
    const int64_t vpb = 2048;
    dim3 grid((unsigned)((S + vpb - 1) / vpb), (unsigned)B);
    if (vec) {
        instnorm_stats_kernel<T, V><<<grid, 256, 256 * 2 * V * sizeof(float), st>>>(x, sums, S, C, C / V, vpb, xs);
    } else {
        if (C > 256) return WF_ERR_UNSUPPORTED;
        instnorm_stats_kernel<T, 1><<<grid, 256, 256 * 2 * sizeof(float), st>>>(x, sums, S, C, C, vpb, xs);
    }
    WF_LAUNCH_CHECK();
    const int n = B * C;
    instnorm_finalize_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(x, sums, mr, n, C, S, xs, (double)eps);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T>
static int apply_launch(const T *x, const float *mr, const T *res, const float *res_mr, T *y, int B, int64_t S, int C,
                        int64_t xs, int64_t rs, int64_t ys, int act, float slope, const float *gamma, const float *beta,
                        cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    const bool vec = (C % V == 0) && aligned16(x) && aligned16(y) && (xs * e) % 16 == 0 && (ys * e) % 16 == 0 &&
                     (res == nullptr || (aligned16(res) && (rs * e) % 16 == 0));
    if (vec) {
        const int64_t total = (int64_t)B * S * (C / V);
        instnorm_apply_kernel<T, V><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, mr, res, res_mr, y, total, S, C, C / V, xs, rs, ys, act, slope, gamma, beta);
    } else {
        const int64_t total = (int64_t)B * S * C;
        instnorm_apply_kernel<T, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, mr, res, res_mr, y, total, S, C, C, xs, rs, ys, act, slope, gamma, beta);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" int wf_instnorm_stats_ndhwc(const void *x, double *sums, float *mean_rstd, int dtype, int B, int64_t S, int C,
                                       int64_t x_vox_stride, float eps, void *stream) {
    if (!x || !sums || !mean_rstd) return WF_ERR_NULL_POINTER;
    if (B <= 0 || S <= 0 || C <= 0 || x_vox_stride < C) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32) return wf::stats_launch<float>((const float *)x, sums, mean_rstd, B, S, C, x_vox_stride, eps, st);
    if (dtype == WF_BF16)
        return wf::stats_launch<__nv_bfloat16>((const __nv_bfloat16 *)x, sums, mean_rstd, B, S, C, x_vox_stride, eps, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_instnorm_apply_ndhwc(const void *x, const float *mean_rstd, const void *res, const float *res_mean_rstd,
                                       const float *gamma, const float *beta, void *y, int act, float slope, int dtype,
                                       int B, int64_t S, int C, int64_t x_vox_stride, int64_t res_vox_stride,
                                       int64_t y_vox_stride, void *stream) {
    if (!x || !mean_rstd || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || S <= 0 || C <= 0 || x_vox_stride < C || y_vox_stride < C || (res && res_vox_stride < C)) return WF_ERR_BAD_SHAPE;
    if (act < 0 || act > 2) return WF_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F32)
        return wf::apply_launch<float>((const float *)x, mean_rstd, (const float *)res, res_mean_rstd, (float *)y, B, S, C,
                                       x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_BF16)
        return wf::apply_launch<__nv_bfloat16>((const __nv_bfloat16 *)x, mean_rstd, (const __nv_bfloat16 *)res, res_mean_rstd,
                                               (__nv_bfloat16 *)y, B, S, C, x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_instnorm_apply_head_ndhwc(const void *x, const float *mean_rstd, const void *res, const float *res_mean_rstd,
                                            const float *head_w, const float *head_b, void *out, int act, float slope,
                                            int dtype, int out_dtype, int B, int64_t S, int C, int K, int64_t x_vox_stride,
                                            int64_t res_vox_stride, void *stream) {
    if (!x || !mean_rstd || !head_w || !out) return WF_ERR_NULL_POINTER;
    if (B <= 0 || S <= 0 || C <= 0 || x_vox_stride < C || (res && res_vox_stride < C)) return WF_ERR_BAD_SHAPE;
    if (act < 0 || act > 2) return WF_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    if (dtype == WF_F32 && out_dtype == WF_F32)
        return wf::apply_head_launch<float, float>((const float *)x, mean_rstd, (const float *)res, res_mean_rstd, head_w, head_b,
                                                   (float *)out, B, S, C, K, x_vox_stride, res_vox_stride, act, slope, st);
    if (dtype == WF_BF16 && out_dtype == WF_F32)
        return wf::apply_head_launch<bf, float>((const bf *)x, mean_rstd, (const bf *)res, res_mean_rstd, head_w, head_b,
                                                (float *)out, B, S, C, K, x_vox_stride, res_vox_stride, act, slope, st);
    if (dtype == WF_BF16 && out_dtype == WF_BF16)
        return wf::apply_head_launch<bf, bf>((const bf *)x, mean_rstd, (const bf *)res, res_mean_rstd, head_w, head_b,
                                             (bf *)out, B, S, C, K, x_vox_stride, res_vox_stride, act, slope, st);
    return WF_ERR_BAD_DTYPE;
}
