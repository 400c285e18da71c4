// InstanceNorm3d (affine-free) on channels-last activations, fused with the activation and the residual add that
// follow it in the reference's conv blocks (sm_100a).
//
// Reference call sites: MONAI UnetResBlock.forward (monai/networks/blocks/dynunet_block.py:98-111: norm1 + lrelu,
// norm2, norm3, `out += residual`, lrelu) and ChannelCalibration.forward (network_models/network_backbone.py:118-128).
// torch's instance_norm copies channels-last tensors to NCDHW and back (35 % of the step before this kernel).
// Two streaming passes: per-(b, c) sum / sum-of-squares (fp32 per thread, fp64 across blocks, so E[x^2]-E[x]^2 does
// not cancel), then y = act((x - mean) * rstd [+ (r - mean_r) * rstd_r | + r]).
#include <type_traits>

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

// four fp32 values -> four 16-bit values of type TO in one 8-byte store (fp32 blocks writing a 16-bit concat slice)
template <typename TO> __device__ __forceinline__ void store_narrow4(TO *p, const float (&f)[4]) {
    uint32_t lo, hi;
    if constexpr (std::is_same<TO, __half>::value) {
        __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
        lo = *reinterpret_cast<uint32_t *>(&a); hi = *reinterpret_cast<uint32_t *>(&b);
    } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
        lo = *reinterpret_cast<uint32_t *>(&a); hi = *reinterpret_cast<uint32_t *>(&b);
    }
    *reinterpret_cast<uint2 *>(p) = make_uint2(lo, hi);
}

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

template <int VEC> __device__ __forceinline__ void load_consts(const float *p, float (&v)[2 * VEC]) {
    if constexpr (VEC % 2 == 0) {     // 2 * VEC floats, 16-byte aligned (c0 is a multiple of VEC)
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2 * VEC; ++i) v[i] = __ldg(p + i);
    }
}

// act: 0 none, 1 relu, 2 leaky relu (slope).  A thread moves U packets (16 bytes of T each, 256 packets apart so that every
// load instruction of a warp is contiguous); all of its loads are issued before the first result is needed.  U = 1 is what runs:
// 4.3 TB/s at 2 x 48 x 128^3, bound by the ~60 instructions per packet (unpack, 2 ops normalise, 2 ops LeakyReLU, pack, index math).
template <typename T, typename TO, int VEC, int U>
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                             const T *__restrict__ res, const float *__restrict__ res_mr,
                                                             TO *__restrict__ y, int64_t total, int64_t S, int C,
                                                             int cvecs, int64_t xs, int64_t rs, int64_t ys, int act,
                                                             float slope, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta) {
    const int64_t base = (int64_t)blockIdx.x * (256 * U) + threadIdx.x;
    int cv[U];
    int64_t vox[U], b[U];   // vox = b*S + v
    float f[U][VEC], r[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t idx = base + u * 256;
        if (idx >= total) { cv[u] = -1; continue; }
        if (total < 0x7fffffffLL) {   // 32-bit divisions: index math must stay cheap
            const uint32_t i32 = (uint32_t)idx, q = i32 / (uint32_t)cvecs;
            cv[u] = (int)(i32 - q * (uint32_t)cvecs);
            vox[u] = q;
            b[u] = q / (uint32_t)S;
        } else {
            cv[u] = (int)(idx % cvecs);
            vox[u] = idx / cvecs;
            b[u] = vox[u] / S;
        }
        NVec<T, VEC>::load(x + vox[u] * xs + cv[u] * VEC, f[u]);
        if (res != nullptr) NVec<T, VEC>::load(res + vox[u] * rs + cv[u] * VEC, r[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (cv[u] < 0) continue;
        const int c0 = cv[u] * VEC;
        // (mean, rstd) pairs of this packet's channels: 2 * VEC consecutive floats, fetched as 16-byte loads when VEC allows
        float mrv[2 * VEC];
        load_consts<VEC>(mr + (b[u] * C + c0) * 2, mrv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[u][e] = (f[u][e] - mrv[2 * e]) * mrv[2 * e + 1];
        if (gamma != nullptr) {  // GroupNorm(num_groups = C) = InstanceNorm + per-channel affine
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[u][e] = fmaf(f[u][e], __ldg(gamma + c0 + e), beta != nullptr ? __ldg(beta + c0 + e) : 0.f);
        }
        if (res != nullptr) {
            if (res_mr != nullptr) {
                load_consts<VEC>(res_mr + (b[u] * C + c0) * 2, mrv);
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[u][e] = (r[u][e] - mrv[2 * e]) * mrv[2 * e + 1];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[u][e] += r[u][e];
        }
        if (act == 1) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[u][e] = fmaxf(f[u][e], 0.f);
        } else if (act == 2) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[u][e] = f[u][e] > 0.f ? f[u][e] : f[u][e] * slope;
        }
        if constexpr (sizeof(TO) == sizeof(T)) {
            NVec<TO, VEC>::store(y + vox[u] * ys + c0, f[u]);
        } else {  // fp32 in, 16-bit out: 4 channels -> 8 bytes
            if constexpr (VEC == 1) {
                y[vox[u] * ys + c0] = from_f32<TO>(f[u][0]);
            } else {
                store_narrow4<TO>(y + vox[u] * ys + c0, f[u]);
            }
        }
    }
}

// The same map for the common case (vector packets, fewer than 2^32 / cvecs packets per sample): grid (packets of a sample / 256, B),
// the sample comes from blockIdx.y and the packet -> (voxel, channel packet) split is one multiply-high by a host-computed
// reciprocal.  The general kernel above spends ~100 of its ~170 executed instructions on index arithmetic (five divisions through
// the reciprocal unit and a 64-bit path): at one 16-byte packet per thread that, not HBM, set its 4.3 TB/s.
template <typename T, typename TO, int VEC, int U>
__global__ void __launch_bounds__(256) instnorm_apply_fast_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                                  const T *__restrict__ res, const float *__restrict__ res_mr,
                                                                  TO *__restrict__ y, uint32_t per_sample, uint32_t S, int C, uint32_t cvecs,
                                                                  uint32_t magic, int64_t xs, int64_t rs, int64_t ys, int act, float slope,
                                                                  const float *__restrict__ gamma, const float *__restrict__ beta) {
    // U packets per thread, 256 apart (every load instruction of a warp stays contiguous); all loads are issued before the first
    // result is needed: one 16-byte packet per thread leaves ~32 KB in flight per SM, short of what 6.4 TB/s needs
    const uint32_t p0 = blockIdx.x * (256u * U) + threadIdx.x;
    const int64_t b = blockIdx.y;
    typename Pack<T>::raw xr[U], rr[U];
    uint32_t v[U];
    int c0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t p = min(p0 + u * 256u, per_sample - 1);     // clamped: out-of-range lanes recompute the last packet, never store
        v[u] = cvecs == 1 ? p : __umulhi(p, magic);                // p / cvecs
        c0[u] = (int)(p - v[u] * cvecs) * VEC;
        xr[u] = *reinterpret_cast<const typename Pack<T>::raw *>(x + (b * S + v[u]) * xs + c0[u]);
        if (res != nullptr) rr[u] = *reinterpret_cast<const typename Pack<T>::raw *>(res + (b * S + v[u]) * rs + c0[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t vox = b * S + v[u];
        float f[VEC];
        Pack<T>::unpack(xr[u], f);
        float mrv[2 * VEC];
        load_consts<VEC>(mr + (b * C + c0[u]) * 2, mrv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = (f[e] - mrv[2 * e]) * mrv[2 * e + 1];
        if (gamma != nullptr) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[e] = fmaf(f[e], __ldg(gamma + c0[u] + e), beta != nullptr ? __ldg(beta + c0[u] + e) : 0.f);
        }
        if (res != nullptr) {
            float r[VEC];
            Pack<T>::unpack(rr[u], r);
            if (res_mr != nullptr) {
                load_consts<VEC>(res_mr + (b * C + c0[u]) * 2, mrv);
#pragma unroll
                for (int e = 0; e < VEC; ++e) r[e] = (r[e] - mrv[2 * e]) * mrv[2 * e + 1];
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
        if (p0 + u * 256u < per_sample) {
            if constexpr (sizeof(TO) == sizeof(T)) {
                NVec<TO, VEC>::store(y + vox * ys + c0[u], f);
            } else {  // fp32 in, 16-bit out: 4 channels -> 8 bytes
                store_narrow4<TO>(y + vox * ys + c0[u], f);
            }
        }
    }
}

// Register-constant, software-pipelined form of the same map (same-width input and output packets, at most 256 channel packets per
// voxel).  A block has group * cvecs threads (group = 256 / cvecs voxels per sweep), so a thread keeps ONE channel packet for its whole
// life: the folded constants  y = x * (rstd * gamma) + (beta - mean * rstd * gamma - mean_r * rstd_r) + r * rstd_r  sit in registers
// instead of being re-read for every packet (four 16-byte constant loads per 16-byte data packet in the kernels above), the index
// arithmetic is one add per packet, and the loads of sweep i + U are issued before sweep i is processed, so a thread always has U
// packets in flight.
// RES: 0 none, 1 residual tensor (raw, or InstanceNorm'd when res_mr is given), 2 the residual is the InstanceNorm'd 1x1x1 shortcut
// convolution of a FOUR-channel input (MONAI UnetResBlock.conv3 / norm3 of the network's first block, dynunet_block.py:104-108),
// recomputed per voxel from the 8- or 16-byte input voxel and w4[C][4] (fp32 copies of the operand-format weights) with rstd_r folded
// in: the 48-channel shortcut tensor is then neither written by the convolution kernel nor read here.
template <typename T, typename TO, typename XT, int VEC, int RES, int U>
__global__ void __launch_bounds__(256) instnorm_apply_reg_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                                 const T *__restrict__ res, const float *__restrict__ res_mr,
                                                                 const XT *__restrict__ xin4, const float *__restrict__ w4,
                                                                 TO *__restrict__ y, uint32_t S, int C, uint32_t cvecs, uint32_t vpb,
                                                                 int64_t xs, int64_t rs, int64_t ys, int act, float slope,
                                                                 const float *__restrict__ gamma, const float *__restrict__ beta) {
    using Raw = typename Pack<T>::raw;
    using XRaw = typename std::conditional<sizeof(XT) == 4, float4, uint2>::type;
    const uint32_t vg = threadIdx.x / cvecs, cv = threadIdx.x - vg * cvecs, group = blockDim.x / cvecs;
    const int64_t b = blockIdx.y;
    const int c0 = (int)cv * VEC;
    float sc[VEC], sh[VEC], rsc[VEC];
    float w[RES == 2 ? VEC : 1][4];
    {
        const float *m = mr + (b * C + c0) * 2;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float r = m[2 * e + 1] * (gamma != nullptr ? gamma[c0 + e] : 1.f);
            sc[e] = r;
            sh[e] = fmaf(-m[2 * e], r, beta != nullptr ? beta[c0 + e] : 0.f);
            rsc[e] = 1.f;
        }
        if (RES != 0 && res_mr != nullptr) {
            const float *m2 = res_mr + (b * C + c0) * 2;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                rsc[e] = m2[2 * e + 1];
                sh[e] = fmaf(-m2[2 * e], rsc[e], sh[e]);
            }
        }
        if constexpr (RES == 2) {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) w[e][k] = w4[(c0 + e) * 4 + k] * rsc[e];
        }
    }
    const uint32_t v_begin = blockIdx.x * vpb, v_end = min(S, v_begin + vpb);
    const uint32_t nsweeps = (v_end - v_begin + group - 1) / group;
    const T *xb = x + b * (int64_t)S * xs + c0;
    const T *rb = RES == 1 ? res + b * (int64_t)S * rs + c0 : nullptr;
    const XT *ib = RES == 2 ? xin4 + b * (int64_t)S * 4 : nullptr;
    TO *yb = y + b * (int64_t)S * ys + c0;
    Raw cur[U], rcur[RES == 1 ? U : 1];
    XRaw icur[RES == 2 ? U : 1];
    auto load = [&](uint32_t sweep, Raw &xr, Raw &rr, XRaw &ir) {
        const int64_t v = min(v_begin + vg + sweep * group, S - 1);      // clamped: lanes past the end recompute the last voxel, never store
        xr = *reinterpret_cast<const Raw *>(xb + v * xs);
        if constexpr (RES == 1) rr = *reinterpret_cast<const Raw *>(rb + v * rs);
        if constexpr (RES == 2) ir = *reinterpret_cast<const XRaw *>(ib + v * 4);
    };
#pragma unroll
    for (int u = 0; u < U; ++u) load(u, cur[u], rcur[RES == 1 ? u : 0], icur[RES == 2 ? u : 0]);
    for (uint32_t s0 = 0; s0 < nsweeps; s0 += U) {
        Raw nxt[U], rnxt[RES == 1 ? U : 1];
        XRaw inxt[RES == 2 ? U : 1];
        if (s0 + U < nsweeps) {
#pragma unroll
            for (int u = 0; u < U; ++u) load(s0 + U + u, nxt[u], rnxt[RES == 1 ? u : 0], inxt[RES == 2 ? u : 0]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t v = v_begin + vg + (s0 + u) * group;
            float f[VEC];
            Pack<T>::unpack(cur[u], f);
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
            if constexpr (RES == 1) {
                float r[VEC];
                Pack<T>::unpack(rcur[u], r);
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] = fmaf(r[e], rsc[e], f[e]);
            }
            if constexpr (RES == 2) {
                float xi[4];
                if constexpr (sizeof(XT) == 4) {      // fp32 input: rounded to the operand format exactly as the convolution kernel does
                    const float4 t = *reinterpret_cast<const float4 *>(&icur[u]);
                    xi[0] = to_f32(from_f32<T>(t.x)); xi[1] = to_f32(from_f32<T>(t.y)); xi[2] = to_f32(from_f32<T>(t.z)); xi[3] = to_f32(from_f32<T>(t.w));
                } else {
                    const uint2 t = *reinterpret_cast<const uint2 *>(&icur[u]);
                    if constexpr (std::is_same<T, __half>::value) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x)), c = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
                        xi[0] = a.x; xi[1] = a.y; xi[2] = c.x; xi[3] = c.y;
                    } else {
                        xi[0] = __uint_as_float(t.x << 16); xi[1] = __uint_as_float(t.x & 0xffff0000u);
                        xi[2] = __uint_as_float(t.y << 16); xi[3] = __uint_as_float(t.y & 0xffff0000u);
                    }
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    f[e] += fmaf(w[e][3], xi[3], fmaf(w[e][2], xi[2], fmaf(w[e][1], xi[1], w[e][0] * xi[0])));
            }
            if (act == 1) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] = fmaxf(f[e], 0.f);
            } else if (act == 2) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * slope;
            }
            if (v < v_end) *reinterpret_cast<typename Pack<TO>::raw *>(yb + (int64_t)v * ys) = Pack<TO>::pack(f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cur[u] = nxt[u];
            if constexpr (RES == 1) rcur[u] = rnxt[u];
            if constexpr (RES == 2) icur[u] = inxt[u];
        }
    }
}

// Last block of the network: y = act((x - mean) * rstd + [(r - mean_r) * rstd_r | r]) is consumed only by the 1x1x1
// output convolution (Waveformer.out, reference network_models/network_backbone.py:407 -> UnetOutBlock,
// monai/networks/blocks/dynunet_block.py:266), so the C-channel activation is never written: each lane normalises one
// 16-byte channel packet and multiplies it with its slice of the [K, C] head weight; the cpv = C / VEC packets of a voxel
// sit in consecutive lanes of one warp (floor(32 / cpv) voxels per warp) and are summed with a segmented shuffle tree;
// lane 0 of each voxel stores the K logits.  No shared memory, no block barrier; two voxel groups per iteration keep
// four 16-byte loads in flight per lane; mean / rstd and the head weights live in registers.
template <typename T, typename TO, int KP>
__global__ void __launch_bounds__(256, KP <= 4 ? 2 : 1) instnorm_apply_head_kernel(const T *__restrict__ x, const float *__restrict__ mr,
                                                                  const T *__restrict__ res, const float *__restrict__ res_mr,
                                                                  const float *__restrict__ hw, const float *__restrict__ hb,
                                                                  TO *__restrict__ out, int64_t S, int64_t total_vox, int C,
                                                                  int K, int cpv, int64_t xs, int64_t rs, int act, float slope) {
    constexpr int VEC = Pack<T>::VEC;
    const int lane = threadIdx.x & 31;
    const int vpw = 32 / cpv;                       // voxels per warp
    const int vl = lane / cpv, cv = lane - vl * cpv;
    const bool lane_live = vl < vpw;
    const int c0 = cv * VEC;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float w[KP][VEC];
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int e = 0; e < VEC; ++e) w[k][e] = (k < K && lane_live) ? __ldg(hw + (int64_t)k * C + c0 + e) : 0.f;
    float bias_k[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) bias_k[k] = (hb != nullptr && k < K) ? __ldg(hb + k) : 0.f;
    // folded normalisation: t = x * sx + r * sr + sh   (sx = rstd, sr = rstd_r or 1, sh = -mean * rstd - mean_r * rstd_r)
    float sx[VEC], sr[VEC], sh[VEC];
    int64_t cur_b = -1;
    auto load_stats = [&](int64_t b) {
        const float *m = mr + (b * C + c0) * 2;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            sx[e] = __ldg(m + 2 * e + 1);
            sh[e] = -__ldg(m + 2 * e) * sx[e];
            sr[e] = 1.f;
        }
        if (res != nullptr && res_mr != nullptr) {
            const float *m2 = res_mr + (b * C + c0) * 2;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                sr[e] = __ldg(m2 + 2 * e + 1);
                sh[e] = fmaf(-__ldg(m2 + 2 * e), sr[e], sh[e]);
            }
        }
        cur_b = b;
    };
    const int64_t groups = (total_vox + vpw - 1) / vpw;          // one group = the voxels one warp covers at once
    const bool small = total_vox < 0x7fffffffLL;                  // 32-bit batch-index division
    using Raw = typename Pack<T>::raw;
    Raw xr[2], rr[2], xn[2], rn[2];                               // current / prefetched packets of two voxel groups
    auto fetch = [&](int64_t g0, Raw (&xd)[2], Raw (&rd)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t vx = (g0 + u) * vpw + vl;
            if (lane_live && (g0 + u) < groups && vx < total_vox) {
                xd[u] = *reinterpret_cast<const Raw *>(x + vx * xs + c0);
                if (res != nullptr) rd[u] = *reinterpret_cast<const Raw *>(res + vx * rs + c0);
            }
        }
    };
    int64_t g0 = warp_id * 2;
    if (g0 < groups) fetch(g0, xr, rr);
    for (; g0 < groups; g0 += nwarps * 2) {
        const int64_t gn = g0 + nwarps * 2;
        if (gn < groups) fetch(gn, xn, rn);                       // next iteration's loads fly while this one computes
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t vx = (g0 + u) * vpw + vl;
            const bool ok = lane_live && (g0 + u) < groups && vx < total_vox;
            float p[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) p[k] = 0.f;
            if (ok) {
                const int64_t b = small ? (int64_t)((uint32_t)vx / (uint32_t)S) : vx / S;
                if (b != cur_b) load_stats(b);
                float f[VEC];
                Pack<T>::unpack(xr[u], f);
#pragma unroll
                for (int e = 0; e < VEC; ++e) f[e] = fmaf(f[e], sx[e], sh[e]);
                if (res != nullptr) {
                    float r[VEC];
                    Pack<T>::unpack(rr[u], r);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) f[e] = fmaf(r[e], sr[e], f[e]);
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
            // segmented tree over the cpv lanes of a voxel (all 32 lanes take part in the shuffles)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                if (o < 2 * cpv) {
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const float t = __shfl_down_sync(0xffffffffu, p[k], o);
                        if (cv + o < cpv) p[k] += t;
                    }
                }
            }
            if (ok && cv == 0) {
                TO *dst = out + vx * (int64_t)K;
                if constexpr (KP == 4 && sizeof(TO) == 4) {
                    if (K == 4) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(p[0] + bias_k[0], p[1] + bias_k[1], p[2] + bias_k[2], p[3] + bias_k[3]);
                        continue;
                    }
                }
                for (int k = 0; k < K; ++k) dst[k] = from_f32<TO>(p[k] + bias_k[k]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) { xr[u] = xn[u]; rr[u] = rn[u]; }
    }
}

// bf16 fast path of the fused head: one thread = one voxel.  All CPV 16-byte packets of x and of the residual are loaded
// up front (2 * CPV independent loads in flight per thread), everything else is thread-local: no shuffles, no barriers in
// the loop.  Per-channel constants live in shared memory as two float4 per channel - (scale_x, scale_r, shift, 0) and the
// K <= 4 head weights - read as warp-wide broadcasts.  grid = (blocks, B): a block stays inside one batch element.
template <typename T, typename TO, int CPV>
__global__ void __launch_bounds__(256) instnorm_apply_head_voxel_kernel(
    const T *__restrict__ x, const float *__restrict__ mr, const T *__restrict__ res,
    const float *__restrict__ res_mr, const float *__restrict__ hw, const float *__restrict__ hb, TO *__restrict__ out,
    int64_t S, int K, int64_t xs, int64_t rs, int act, float slope) {
    constexpr int C = CPV * 8;
    __shared__ float4 s_norm[C], s_w[C];
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float m = mr[((int64_t)b * C + c) * 2], r = mr[((int64_t)b * C + c) * 2 + 1];
        float sr = 0.f, sh = -m * r;
        if (res != nullptr) {
            sr = 1.f;
            if (res_mr != nullptr) {
                sr = res_mr[((int64_t)b * C + c) * 2 + 1];
                sh = fmaf(-res_mr[((int64_t)b * C + c) * 2], sr, sh);
            }
        }
        s_norm[c] = make_float4(r, sr, sh, 0.f);
        s_w[c] = make_float4(hw[c], K > 1 ? hw[C + c] : 0.f, K > 2 ? hw[2 * C + c] : 0.f, K > 3 ? hw[3 * C + c] : 0.f);
    }
    __syncthreads();
    const float4 bias = make_float4(hb ? hb[0] : 0.f, (hb && K > 1) ? hb[1] : 0.f, (hb && K > 2) ? hb[2] : 0.f,
                                    (hb && K > 3) ? hb[3] : 0.f);
    const T *xb = x + (int64_t)b * S * xs;
    const T *rb = res != nullptr ? res + (int64_t)b * S * rs : nullptr;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < S; v += (int64_t)gridDim.x * blockDim.x) {
        uint4 xr[CPV], rr[CPV];
#pragma unroll
        for (int j = 0; j < CPV; ++j) xr[j] = __ldg(reinterpret_cast<const uint4 *>(xb + v * xs) + j);
        if (rb != nullptr) {
#pragma unroll
            for (int j = 0; j < CPV; ++j) rr[j] = __ldg(reinterpret_cast<const uint4 *>(rb + v * rs) + j);
        }
        float a0 = bias.x, a1 = bias.y, a2 = bias.z, a3 = bias.w;
#pragma unroll
        for (int j = 0; j < CPV; ++j) {
            float f[8], r[8];
            Pack<T>::unpack(xr[j], f);
            if (rb != nullptr) Pack<T>::unpack(rr[j], r);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 n = s_norm[j * 8 + e], w = s_w[j * 8 + e];
                float t = fmaf(f[e], n.x, n.z);
                if (rb != nullptr) t = fmaf(r[e], n.y, t);
                if (act == 1) t = fmaxf(t, 0.f);
                else if (act == 2) t = fmaxf(t, t * slope);   // LeakyReLU for 0 <= slope <= 1
                a0 = fmaf(t, w.x, a0); a1 = fmaf(t, w.y, a1); a2 = fmaf(t, w.z, a2); a3 = fmaf(t, w.w, a3);
            }
        }
        TO *dst = out + ((int64_t)b * S + v) * K;
        if constexpr (sizeof(TO) == 4) {
            if (K == 4) {
                *reinterpret_cast<float4 *>(dst) = make_float4(a0, a1, a2, a3);
                continue;
            }
        }
        dst[0] = from_f32<TO>(a0);
        if (K > 1) dst[1] = from_f32<TO>(a1);
        if (K > 2) dst[2] = from_f32<TO>(a2);
        if (K > 3) dst[3] = from_f32<TO>(a3);
    }
}

// Two voxels per thread (v and v + 256): a per-channel constant pair read from shared memory (two 16-byte broadcasts = eight
// passes of the shared-memory pipe) now serves two voxels.  The one-voxel kernel above kept that pipe 91 % busy (ncu) and ran at
// 5.1 TB/s; arithmetic and its order per voxel are unchanged.
template <typename T, typename TO, int CPV>
__global__ void __launch_bounds__(256) instnorm_apply_head_voxel2_kernel(
    const T *__restrict__ x, const float *__restrict__ mr, const T *__restrict__ res,
    const float *__restrict__ res_mr, const float *__restrict__ hw, const float *__restrict__ hb, TO *__restrict__ out,
    int64_t S, int K, int64_t xs, int64_t rs, int act, float slope) {
    constexpr int C = CPV * 8;
    __shared__ float4 s_norm[C], s_w[C];
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float m = mr[((int64_t)b * C + c) * 2], r = mr[((int64_t)b * C + c) * 2 + 1];
        float sr = 0.f, sh = -m * r;
        if (res != nullptr) {
            sr = 1.f;
            if (res_mr != nullptr) {
                sr = res_mr[((int64_t)b * C + c) * 2 + 1];
                sh = fmaf(-res_mr[((int64_t)b * C + c) * 2], sr, sh);
            }
        }
        s_norm[c] = make_float4(r, sr, sh, 0.f);
        s_w[c] = make_float4(hw[c], K > 1 ? hw[C + c] : 0.f, K > 2 ? hw[2 * C + c] : 0.f, K > 3 ? hw[3 * C + c] : 0.f);
    }
    __syncthreads();
    const float4 bias = make_float4(hb ? hb[0] : 0.f, (hb && K > 1) ? hb[1] : 0.f, (hb && K > 2) ? hb[2] : 0.f,
                                    (hb && K > 3) ? hb[3] : 0.f);
    const T *xb = x + (int64_t)b * S * xs;
    const T *rb = res != nullptr ? res + (int64_t)b * S * rs : nullptr;
    for (int64_t v0 = (int64_t)blockIdx.x * 512 + threadIdx.x; v0 < S; v0 += (int64_t)gridDim.x * 512) {
        const int64_t v1 = v0 + 256 < S ? v0 + 256 : v0;      // past the end: recompute v0, never store
        uint4 xr[2][CPV], rr[2][CPV];
#pragma unroll
        for (int j = 0; j < CPV; ++j) {
            xr[0][j] = __ldg(reinterpret_cast<const uint4 *>(xb + v0 * xs) + j);
            xr[1][j] = __ldg(reinterpret_cast<const uint4 *>(xb + v1 * xs) + j);
        }
        if (rb != nullptr) {
#pragma unroll
            for (int j = 0; j < CPV; ++j) {
                rr[0][j] = __ldg(reinterpret_cast<const uint4 *>(rb + v0 * rs) + j);
                rr[1][j] = __ldg(reinterpret_cast<const uint4 *>(rb + v1 * rs) + j);
            }
        }
        float a[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) { a[u][0] = bias.x; a[u][1] = bias.y; a[u][2] = bias.z; a[u][3] = bias.w; }
#pragma unroll
        for (int j = 0; j < CPV; ++j) {
            float f[2][8], r[2][8];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                Pack<T>::unpack(xr[u][j], f[u]);
                if (rb != nullptr) Pack<T>::unpack(rr[u][j], r[u]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 n = s_norm[j * 8 + e], w = s_w[j * 8 + e];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float t = fmaf(f[u][e], n.x, n.z);
                    if (rb != nullptr) t = fmaf(r[u][e], n.y, t);
                    if (act == 1) t = fmaxf(t, 0.f);
                    else if (act == 2) t = fmaxf(t, t * slope);   // LeakyReLU for 0 <= slope <= 1
                    a[u][0] = fmaf(t, w.x, a[u][0]); a[u][1] = fmaf(t, w.y, a[u][1]); a[u][2] = fmaf(t, w.z, a[u][2]); a[u][3] = fmaf(t, w.w, a[u][3]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && v0 + 256 >= S) break;
            TO *dst = out + ((int64_t)b * S + (u == 0 ? v0 : v1)) * K;
            if constexpr (sizeof(TO) == 4) {
                if (K == 4) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(a[u][0], a[u][1], a[u][2], a[u][3]);
                    continue;
                }
            }
            dst[0] = from_f32<TO>(a[u][0]);
            if (K > 1) dst[1] = from_f32<TO>(a[u][1]);
            if (K > 2) dst[2] = from_f32<TO>(a[u][2]);
            if (K > 3) dst[3] = from_f32<TO>(a[u][3]);
        }
    }
}

template <typename T, typename TO>
static bool apply_head_voxel_launch(const T *x, const float *mr, const T *res, const float *res_mr,
                                    const float *hw, const float *hb, TO *out, int B, int64_t S, int C, int K, int64_t xs,
                                    int64_t rs, int act, float slope, cudaStream_t st) {
    if (K > 4 || C % 8 != 0 || B > 65535 || (act == 2 && (slope < 0.f || slope > 1.f))) return false;
    const int64_t want = (S + 255) / 256;
    dim3 grid((unsigned)min(want, (int64_t)kNumSMs * 6), (unsigned)B);
#define WF_HV(CPV_) instnorm_apply_head_voxel_kernel<T, TO, CPV_><<<grid, 256, 0, st>>>(x, mr, res, res_mr, hw, hb, out, S, K, xs, rs, act, slope)
    if (!ab_old() && C / 8 <= 6 && C / 8 % 2 == 0 && S >= 512 * (int64_t)kNumSMs) {      // large maps: two voxels per thread
        const int64_t want2 = (S + 511) / 512;
        dim3 grid2((unsigned)min(want2, (int64_t)kNumSMs * 4), (unsigned)B);
#define WF_HV2(CPV_) instnorm_apply_head_voxel2_kernel<T, TO, CPV_><<<grid2, 256, 0, st>>>(x, mr, res, res_mr, hw, hb, out, S, K, xs, rs, act, slope)
        if (C / 8 == 2) WF_HV2(2); else if (C / 8 == 4) WF_HV2(4); else WF_HV2(6);
#undef WF_HV2
        return true;
    }
    switch (C / 8) {
        case 2: WF_HV(2); break;
        case 4: WF_HV(4); break;
        case 6: WF_HV(6); break;
        case 8: WF_HV(8); break;
        case 12: WF_HV(12); break;
        default: return false;
    }
#undef WF_HV
    return true;
}

template <typename T, typename TO>
static int apply_head_launch(const T *x, const float *mr, const T *res, const float *res_mr, const float *hw, const float *hb,
                             TO *out, int B, int64_t S, int C, int K, int64_t xs, int64_t rs, int act, float slope,
                             cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    if (C % V != 0 || C / V > 32 || K < 1 || K > 16) return WF_ERR_UNSUPPORTED;
    if (!aligned16(x) || (xs * e) % 16 != 0 || (res != nullptr && (!aligned16(res) || (rs * e) % 16 != 0))) return WF_ERR_MISALIGNED;
    if constexpr (sizeof(T) == 2) {
        if (apply_head_voxel_launch<T, TO>(x, mr, res, res_mr, hw, hb, out, B, S, C, K, xs, rs, act, slope, st)) {
            WF_LAUNCH_CHECK();
            return WF_OK;
        }
    }
    const int cpv = C / V;
    const int64_t total = (int64_t)B * S;
    const int vpw = 32 / cpv;
    const int64_t groups = (total + vpw - 1) / vpw;
    const int64_t want_blocks = (groups / 2 + 7) / 8 + 1;                    // 8 warps per block, 2 groups per iteration
    const unsigned grid = (unsigned)min(want_blocks, (int64_t)kNumSMs * 8);
#define WF_HEAD(KP_) instnorm_apply_head_kernel<T, TO, KP_><<<grid, 256, 0, st>>>( \
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
    // voxels per block: enough blocks to fill the GPU several times over even for the 64^3 / 32^3 maps (a fixed 2048 left a
    // 2 x 64^3 volume with 256 blocks on 148 SMs: 77 us for a 100 MB read)
    int64_t vpb = (S * B + (int64_t)kNumSMs * 16 - 1) / ((int64_t)kNumSMs * 16);
    vpb = vpb < 128 ? 128 : (vpb > 2048 ? 2048 : vpb);
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

template <typename T, typename TO, typename XT, int RES>
static int apply_reg_launch(const T *x, const float *mr, const T *res, const float *res_mr, const XT *xin4, const float *w4, TO *y, int B,
                            int64_t S, int C, int64_t xs, int64_t rs, int64_t ys, int act, float slope, const float *gamma,
                            const float *beta, cudaStream_t st) {
    constexpr int V = Pack<T>::VEC, U = 4;
    const uint32_t cvecs = (uint32_t)(C / V), group = 256u / cvecs;
    // sweeps per block: enough blocks for ~16 per SM when the tensor is large, never fewer than one pipeline step
    int64_t steps = (S * B) / ((int64_t)group * U * kNumSMs * 16);
    steps = steps < 1 ? 1 : (steps > 8 ? 8 : steps);
    const uint32_t vpb = group * U * (uint32_t)steps;
    const dim3 grid((unsigned)((S + vpb - 1) / vpb), (unsigned)B);
    instnorm_apply_reg_kernel<T, TO, XT, V, RES, U><<<grid, group * cvecs, 0, st>>>(x, mr, res, res_mr, xin4, w4, y, (uint32_t)S, C, cvecs, vpb, xs,
                                                                                   rs, ys, act, slope, gamma, beta);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <typename T, typename TO>
static int apply_launch(const T *x, const float *mr, const T *res, const float *res_mr, TO *y, int B, int64_t S, int C,
                        int64_t xs, int64_t rs, int64_t ys, int act, float slope, const float *gamma, const float *beta,
                        cudaStream_t st) {
    constexpr int V = Pack<T>::VEC;
    const size_t e = sizeof(T);
    // a packet is V elements: 16 bytes of T on the input side, V * sizeof(TO) bytes (16, or 8 for fp32 -> bf16) on the output
    const size_t out_packet = V * sizeof(TO);
    const bool vec = (C % V == 0) && aligned16(x) && (xs * e) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(y) % out_packet) == 0 && (ys * sizeof(TO)) % out_packet == 0 &&
                     (res == nullptr || (aligned16(res) && (rs * e) % 16 == 0));
    const int64_t per_sample = S * (C / V);
    if constexpr (sizeof(TO) == sizeof(T) && V > 1) {
        if (vec && B <= 65535 && C / V <= 256 && S < 0x7fffffffLL) {
            if (res == nullptr) return apply_reg_launch<T, TO, uint16_t, 0>(x, mr, nullptr, nullptr, nullptr, nullptr, y, B, S, C, xs, rs, ys, act, slope, gamma, beta, st);
            return apply_reg_launch<T, TO, uint16_t, 1>(x, mr, res, res_mr, nullptr, nullptr, y, B, S, C, xs, rs, ys, act, slope, gamma, beta, st);
        }
    }
    if (vec && V > 1 && B <= 65535 && per_sample * (C / V) < 0xffffffffLL && per_sample + 255 < 0xffffffffLL) {
        const uint32_t cvecs = (uint32_t)(C / V);
        const uint32_t magic = cvecs == 1 ? 0u : (uint32_t)((0x100000000ULL + cvecs - 1) / cvecs);   // exact for p * cvecs < 2^32
        constexpr int U = 4;
        const dim3 grid((unsigned)((per_sample + 256 * U - 1) / (256 * U)), (unsigned)B);
        instnorm_apply_fast_kernel<T, TO, V, U><<<grid, 256, 0, st>>>(x, mr, res, res_mr, y, (uint32_t)per_sample, (uint32_t)S, C, cvecs, magic, xs, rs,
                                                                    ys, act, slope, gamma, beta);
    } else if (vec) {
        const int64_t total = (int64_t)B * S * (C / V);
        constexpr int U = 1;     // measured: 4 packets per thread 0.38 ms, 1 packet 0.19 ms at 2 x 48 x 128^3 (the kernel is instruction-bound:
                                 // ~60 instructions per 16-byte packet, not latency-bound)
        instnorm_apply_kernel<T, TO, V, U><<<(unsigned)((total + 256 * U - 1) / (256 * U)), 256, 0, st>>>(x, mr, res, res_mr, y, total, S, C, C / V, xs, rs, ys, act, slope, gamma, beta);
    } else {
        const int64_t total = (int64_t)B * S * C;
        instnorm_apply_kernel<T, TO, 1, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, mr, res, res_mr, y, total, S, C, C, xs, rs, ys, act, slope, gamma, beta);
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
    if (dtype == WF_F16)
        return wf::stats_launch<__half>((const __half *)x, sums, mean_rstd, B, S, C, x_vox_stride, eps, st);
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_instnorm_apply_ndhwc(const void *x, const float *mean_rstd, const void *res, const float *res_mean_rstd,
                                       const float *gamma, const float *beta, void *y, int act, float slope, int dtype,
                                       int y_dtype, int B, int64_t S, int C, int64_t x_vox_stride, int64_t res_vox_stride,
                                       int64_t y_vox_stride, void *stream) {
    if (!x || !mean_rstd || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || S <= 0 || C <= 0 || x_vox_stride < C || y_vox_stride < C || (res && res_vox_stride < C)) return WF_ERR_BAD_SHAPE;
    if (act < 0 || act > 2) return WF_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    if (dtype == WF_F32 && y_dtype == WF_F32)
        return wf::apply_launch<float, float>((const float *)x, mean_rstd, (const float *)res, res_mean_rstd, (float *)y, B, S, C,
                                              x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_F32 && y_dtype == WF_BF16)   // fp32 block (TF32 convolutions) writing a bf16 concat slice
        return wf::apply_launch<float, bf>((const float *)x, mean_rstd, (const float *)res, res_mean_rstd, (bf *)y, B, S, C,
                                           x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_BF16 && y_dtype == WF_BF16)
        return wf::apply_launch<bf, bf>((const bf *)x, mean_rstd, (const bf *)res, res_mean_rstd, (bf *)y, B, S, C, x_vox_stride,
                                        res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_F16 && y_dtype == WF_F16)    // fp16 skip block, intermediate activation
        return wf::apply_launch<__half, __half>((const __half *)x, mean_rstd, (const __half *)res, res_mean_rstd, (__half *)y, B,
                                                S, C, x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_F16 && y_dtype == WF_BF16)   // fp16 skip block writing its bf16 concat slice
        return wf::apply_launch<__half, bf>((const __half *)x, mean_rstd, (const __half *)res, res_mean_rstd, (bf *)y, B, S, C,
                                            x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    if (dtype == WF_F32 && y_dtype == WF_F16)    // fp32 block (TF32 convolutions) writing an fp16 concat slice
        return wf::apply_launch<float, __half>((const float *)x, mean_rstd, (const float *)res, res_mean_rstd, (__half *)y, B, S,
                                               C, x_vox_stride, res_vox_stride, y_vox_stride, act, slope, gamma, beta, st);
    return WF_ERR_BAD_DTYPE;
}

namespace wf {

// First and second moments of a FOUR-channel volume per sample: sums[b][0..3] = sum x_k, sums[b][4..13] = sum x_k x_l (k <= l, row-major
// upper triangle), fp64.  The InstanceNorm statistics of ANY 1x1x1 convolution of that volume follow from these 14 numbers
// (mean_c = w_c . m, var_c = w_c^T Cov w_c), so the first block's shortcut needs neither its own accumulator columns in the convolution
// kernel nor a pass over its 48-channel result.  x is rounded to the operand format T first (what the convolution multiplies).
template <typename T, typename XT>
__global__ void __launch_bounds__(256) moments4_kernel(const XT *__restrict__ x, double *__restrict__ sums, int64_t S, int64_t vpb) {
    using XRaw = typename std::conditional<sizeof(XT) == 4, float4, uint2>::type;
    const int64_t b = blockIdx.y;
    const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = min(S, v0 + vpb);
    const XT *xb = x + b * S * 4;
    float m[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) m[i] = 0.f;
    for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {      // <= vpb / 256 values per thread and fp32 partial (vpb = 8192: 32)
        const XRaw raw = *reinterpret_cast<const XRaw *>(xb + v * 4);
        float f[4];
        if constexpr (sizeof(XT) == 4) {
            const float4 t = *reinterpret_cast<const float4 *>(&raw);
            f[0] = to_f32(from_f32<T>(t.x)); f[1] = to_f32(from_f32<T>(t.y)); f[2] = to_f32(from_f32<T>(t.z)); f[3] = to_f32(from_f32<T>(t.w));
        } else {
            const uint2 t = *reinterpret_cast<const uint2 *>(&raw);
            if constexpr (std::is_same<T, __half>::value) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x)), c = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
                f[0] = a.x; f[1] = a.y; f[2] = c.x; f[3] = c.y;
            } else {
                f[0] = __uint_as_float(t.x << 16); f[1] = __uint_as_float(t.x & 0xffff0000u);
                f[2] = __uint_as_float(t.y << 16); f[3] = __uint_as_float(t.y & 0xffff0000u);
            }
        }
        int i = 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m[k] += f[k];
#pragma unroll
            for (int l = k; l < 4; ++l) { m[i] = fmaf(f[k], f[l], m[i]); ++i; }
        }
    }
    __shared__ double red[8][14];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        double d = (double)m[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) red[warp][i] = d;
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += red[w][threadIdx.x];
        atomicAdd(sums + b * 14 + threadIdx.x, d);
    }
}

// mr[(b * C + c) * 2] = mean, + 1 = 1 / sqrt(var + eps) of y_c = sum_k w4[c][k] x_k from the moments above
__global__ void shortcut4_stats_finalize_kernel(const double *__restrict__ sums, const float *__restrict__ w4, float *__restrict__ mr,
                                                int n, int C, double inv_s, double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // b * C + c
    if (i >= n) return;
    const int b = i / C, c = i - b * C;
    const double *sm = sums + (int64_t)b * 14;
    double mean[4], cov[4][4];
    for (int k = 0; k < 4; ++k) mean[k] = sm[k] * inv_s;
    int j = 4;
    for (int k = 0; k < 4; ++k)
        for (int l = k; l < 4; ++l) {
            cov[k][l] = cov[l][k] = sm[j] * inv_s - mean[k] * mean[l];
            ++j;
        }
    double w[4];
    for (int k = 0; k < 4; ++k) w[k] = (double)w4[c * 4 + k];
    double mu = 0.0, var = 0.0;
    for (int k = 0; k < 4; ++k) {
        mu += w[k] * mean[k];
        for (int l = 0; l < 4; ++l) var += w[k] * w[l] * cov[k][l];
    }
    var = var < 0.0 ? 0.0 : var;
    mr[2 * i] = (float)mu;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

}  // namespace wf

extern "C" int wf_shortcut4_stats(const void *x, int x_dtype, int op_dtype, const float *w4, double *sums, float *mean_rstd, float eps,
                                  int B, int64_t S, int C, void *stream) {
    if (!x || !w4 || !sums || !mean_rstd) return WF_ERR_NULL_POINTER;
    if (B <= 0 || B > 65535 || S <= 0 || C <= 0) return WF_ERR_BAD_SHAPE;
    if (op_dtype != WF_BF16 && op_dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (x_dtype != WF_F32 && x_dtype != op_dtype) return WF_ERR_BAD_DTYPE;
    if (!wf::aligned16(x)) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    WF_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 14 * (size_t)B, st));
    const int64_t vpb = 8192;
    const dim3 grid((unsigned)((S + vpb - 1) / vpb), (unsigned)B);
    using bf = __nv_bfloat16;
    if (op_dtype == WF_F16) {
        if (x_dtype == WF_F32) wf::moments4_kernel<__half, float><<<grid, 256, 0, st>>>((const float *)x, sums, S, vpb);
        else wf::moments4_kernel<__half, uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)x, sums, S, vpb);
    } else {
        if (x_dtype == WF_F32) wf::moments4_kernel<bf, float><<<grid, 256, 0, st>>>((const float *)x, sums, S, vpb);
        else wf::moments4_kernel<bf, uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)x, sums, S, vpb);
    }
    WF_LAUNCH_CHECK();
    const int n = B * C;
    wf::shortcut4_stats_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(sums, w4, mean_rstd, n, C, 1.0 / (double)S, (double)eps);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_instnorm_apply_shortcut4_ndhwc(const void *x, const float *mean_rstd, const void *xin, int xin_dtype, const float *w4,
                                                 const float *res_mean_rstd, void *y, int act, float slope, int dtype, int B, int64_t S,
                                                 int C, int64_t x_vox_stride, int64_t y_vox_stride, void *stream) {
    if (!x || !mean_rstd || !xin || !w4 || !res_mean_rstd || !y) return WF_ERR_NULL_POINTER;
    if (B <= 0 || B > 65535 || S <= 0 || S >= 0x7fffffffLL || C <= 0 || C % 8 || C / 8 > 256 || x_vox_stride < C || y_vox_stride < C ||
        x_vox_stride % 8 || y_vox_stride % 8)
        return WF_ERR_BAD_SHAPE;
    if (act < 0 || act > 2) return WF_ERR_UNSUPPORTED;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (xin_dtype != WF_F32 && xin_dtype != dtype) return WF_ERR_BAD_DTYPE;
    if (!wf::aligned16(x) || !wf::aligned16(y) || !wf::aligned16(xin)) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
#define WF_SC4(T_, XT_)                                                                                                               \
    return wf::apply_reg_launch<T_, T_, XT_, 2>((const T_ *)x, mean_rstd, nullptr, res_mean_rstd, (const XT_ *)xin, w4, (T_ *)y, B, S, C, \
                                                x_vox_stride, C, y_vox_stride, act, slope, nullptr, nullptr, st)
    if (dtype == WF_F16) {
        if (xin_dtype == WF_F32) WF_SC4(__half, float);
        WF_SC4(__half, uint16_t);
    }
    if (xin_dtype == WF_F32) WF_SC4(bf, float);
    WF_SC4(bf, uint16_t);
#undef WF_SC4
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
    if (dtype == WF_F16 && out_dtype == WF_F32)
        return wf::apply_head_launch<__half, float>((const __half *)x, mean_rstd, (const __half *)res, res_mean_rstd, head_w,
                                                    head_b, (float *)out, B, S, C, K, x_vox_stride, res_vox_stride, act, slope, st);
    if (dtype == WF_F16 && out_dtype == WF_F16)
        return wf::apply_head_launch<__half, __half>((const __half *)x, mean_rstd, (const __half *)res, res_mean_rstd, head_w,
                                                     head_b, (__half *)out, B, S, C, K, x_vox_stride, res_vox_stride, act, slope, st);
    return WF_ERR_BAD_DTYPE;
}
