// Sliding-window stitching kernels (re-hosted MONAI inferer; reference monai/inferers/utils.py:216-299,
// monai/data/utils.py:1121-1138).  All are streaming, HBM-bound, fp32 accumulation.
#include "wf_common.cuh"

namespace wf {

// vol [Bv, C, D, H, W] fp32 -> win [nwin, C, r0, r1, r2] (or [nwin, r0, r1, r2, C]) in T
template <typename T, bool CL>
__global__ void __launch_bounds__(256) sw_gather_kernel(const float *__restrict__ vol, T *__restrict__ win,
                                                        const int32_t *__restrict__ starts, int64_t total, int C, int D,
                                                        int H, int W, int r0, int r1, int r2, int flip) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int64_t t = idx;
    int c, x, y, z;
    if (CL) {
        c = (int)(t % C); t /= C;
        x = (int)(t % r2); t /= r2;
        y = (int)(t % r1); t /= r1;
        z = (int)(t % r0); t /= r0;
    } else {
        x = (int)(t % r2); t /= r2;
        y = (int)(t % r1); t /= r1;
        z = (int)(t % r0); t /= r0;
        c = (int)(t % C); t /= C;
    }
    const int n = (int)t;
    const int b = starts[4 * n], z0 = starts[4 * n + 1], y0 = starts[4 * n + 2], x0 = starts[4 * n + 3];
    // mirror test-time augmentation (light_training/prediction.py:129-156): the window lives in the MIRRORED volume
    // V'(p) = V(flip(p)); flip bit 0 / 1 / 2 = z / y / x.  The mirrored copy itself is never built.
    int pz = z0 + z, py = y0 + y, px = x0 + x;
    if (flip & 1) pz = D - 1 - pz;
    if (flip & 2) py = H - 1 - py;
    if (flip & 4) px = W - 1 - px;
    const float v = __ldg(vol + ((((int64_t)b * C + c) * D + pz) * H + py) * (int64_t)W + px);
    win[idx] = from_f32<T>(v);
}

template <typename T, bool CL>
__global__ void __launch_bounds__(256) sw_accumulate_kernel(const T *__restrict__ seg, float *__restrict__ acc,
                                                            const int32_t *__restrict__ starts,
                                                            const float *__restrict__ gz, const float *__restrict__ gy,
                                                            const float *__restrict__ gx, float floor_w, int64_t total,
                                                            int K, int D, int H, int W, int r0, int r1, int r2, int flip) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int64_t t = idx;
    int c, x, y, z;
    if (CL) {
        c = (int)(t % K); t /= K;
        x = (int)(t % r2); t /= r2;
        y = (int)(t % r1); t /= r1;
        z = (int)(t % r0); t /= r0;
    } else {
        x = (int)(t % r2); t /= r2;
        y = (int)(t % r1); t /= r1;
        z = (int)(t % r0); t /= r0;
        c = (int)(t % K); t /= K;
    }
    const int n = (int)t;
    const int b = starts[4 * n], z0 = starts[4 * n + 1], y0 = starts[4 * n + 2], x0 = starts[4 * n + 3];
    // importance weight exactly as compute_importance_map builds it: ((gz*gy)*gx) in fp32, clamped from below
    const float wgt = fmaxf((gz[z] * gy[y]) * gx[x], floor_w);
    const float v = to_f32(seg[idx]) * wgt;
    int pz = z0 + z, py = y0 + y, px = x0 + x;      // mirrored pass: scatter straight back into the un-mirrored volume
    if (flip & 1) pz = D - 1 - pz;
    if (flip & 2) py = H - 1 - py;
    if (flip & 4) px = W - 1 - px;
    atomicAdd(acc + ((((int64_t)b * K + c) * D + pz) * H + py) * (int64_t)W + px, v);
}

// Channels-last fast paths: one thread per window VOXEL (grid.y = window), 32-bit index arithmetic, the C channel planes read /
// updated as C coalesced accesses (consecutive threads = consecutive x) and the voxel's C values moved as one packet.  The
// per-element kernels above (one thread per value, 64-bit divisions, a warp touching every channel plane in each load) ran at
// 1.2 TB/s: 0.33 ms per 6-window gather where the traffic is worth 0.07 ms.
template <typename T, int C>
__global__ void __launch_bounds__(256) sw_gather_voxel_kernel(const float *__restrict__ vol, T *__restrict__ win,
                                                              const int32_t *__restrict__ starts, int D, int H, int W, int r0, int r1,
                                                              int r2, int flip) {
    const uint32_t nvox = (uint32_t)r0 * r1 * r2;
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    if (v >= nvox) return;
    const int n = blockIdx.y;
    const uint32_t x = v % (uint32_t)r2, q = v / (uint32_t)r2, y = q % (uint32_t)r1, z = q / (uint32_t)r1;
    const int b = starts[4 * n];
    int pz = starts[4 * n + 1] + (int)z, py = starts[4 * n + 2] + (int)y, px = starts[4 * n + 3] + (int)x;
    if (flip & 1) pz = D - 1 - pz;
    if (flip & 2) py = H - 1 - py;
    if (flip & 4) px = W - 1 - px;
    const int64_t plane = (int64_t)D * H * W;
    const float *src = vol + (int64_t)b * C * plane + ((int64_t)pz * H + py) * W + px;
    float f[C];
#pragma unroll
    for (int c = 0; c < C; ++c) f[c] = __ldg(src + c * plane);
    T *dst = win + ((int64_t)n * nvox + v) * C;
    if constexpr (C == 4 && sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    } else if constexpr (C == 4 && sizeof(T) == 2) {
        T h[4] = {from_f32<T>(f[0]), from_f32<T>(f[1]), from_f32<T>(f[2]), from_f32<T>(f[3])};
        *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(h);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = from_f32<T>(f[c]);
    }
}

template <typename T, int K>
__global__ void __launch_bounds__(256) sw_accumulate_voxel_kernel(const T *__restrict__ seg, float *__restrict__ acc,
                                                                  const int32_t *__restrict__ starts, const float *__restrict__ gz,
                                                                  const float *__restrict__ gy, const float *__restrict__ gx,
                                                                  float floor_w, int D, int H, int W, int r0, int r1, int r2, int flip) {
    const uint32_t nvox = (uint32_t)r0 * r1 * r2;
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    if (v >= nvox) return;
    const int n = blockIdx.y;
    const uint32_t x = v % (uint32_t)r2, q = v / (uint32_t)r2, y = q % (uint32_t)r1, z = q / (uint32_t)r1;
    const int b = starts[4 * n];
    // importance weight exactly as compute_importance_map builds it: ((gz*gy)*gx) in fp32, clamped from below
    const float wgt = fmaxf((gz[z] * gy[y]) * gx[x], floor_w);
    const T *src = seg + ((int64_t)n * nvox + v) * K;
    float f[K];
    if constexpr (K == 4 && sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(src);
        f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    } else if constexpr (K == 4 && sizeof(T) == 2) {
        const uint2 t = *reinterpret_cast<const uint2 *>(src);
        const T *h = reinterpret_cast<const T *>(&t);
#pragma unroll
        for (int c = 0; c < 4; ++c) f[c] = to_f32(h[c]);
    } else {
#pragma unroll
        for (int c = 0; c < K; ++c) f[c] = to_f32(src[c]);
    }
    int pz = starts[4 * n + 1] + (int)z, py = starts[4 * n + 2] + (int)y, px = starts[4 * n + 3] + (int)x;
    if (flip & 1) pz = D - 1 - pz;
    if (flip & 2) py = H - 1 - py;
    if (flip & 4) px = W - 1 - px;
    const int64_t plane = (int64_t)D * H * W;
    float *dst = acc + (int64_t)b * K * plane + ((int64_t)pz * H + py) * W + px;
#pragma unroll
    for (int c = 0; c < K; ++c) atomicAdd(dst + c * plane, f[c] * wgt);   // same product, same rounding as the per-element kernel
}

// one thread per voxel: count = sum of window weights covering it; acc[:, k] /= count; optional argmax
__global__ void __launch_bounds__(256) sw_finalize_kernel(float *__restrict__ acc, uint8_t *__restrict__ labels,
                                                          const int32_t *__restrict__ all_starts, int nall,
                                                          const float *__restrict__ gz, const float *__restrict__ gy,
                                                          const float *__restrict__ gx, float floor_w, int64_t total,
                                                          int K, int D, int H, int W, int r0, int r1, int r2, int z_begin,
                                                          int z_end, int flip, float *__restrict__ dst, float dst_scale,
                                                          int dst_add) {
    extern __shared__ int32_t s_starts[];
    for (int i = threadIdx.x; i < 4 * nall; i += blockDim.x) s_starts[i] = all_starts[i];
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int64_t t = idx;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H); t /= H;
    const int nz = z_end - z_begin;                 // only planes [z_begin, z_end) of every volume (streamed output)
    const int z = z_begin + (int)(t % nz); t /= nz;
    const int b = (int)t;
    // a mirrored pass was stitched in the mirrored frame: its count map is the un-mirrored one read at flip(q)
    const int cz = (flip & 1) ? D - 1 - z : z, cy = (flip & 2) ? H - 1 - y : y, cx = (flip & 4) ? W - 1 - x : x;
    float count = 0.f;
    for (int n = 0; n < nall; ++n) {  // same order as the reference's `for __s in slices: count_map[__s] += w`
        if (s_starts[4 * n] >= 0 && s_starts[4 * n] != b) continue;   // slot < 0: the window applies to every volume
        const int lz = cz - s_starts[4 * n + 1], ly = cy - s_starts[4 * n + 2], lx = cx - s_starts[4 * n + 3];
        if ((unsigned)lz < (unsigned)r0 && (unsigned)ly < (unsigned)r1 && (unsigned)lx < (unsigned)r2)
            count += fmaxf((gz[lz] * gy[ly]) * gx[lx], floor_w);
    }
    const int64_t plane = (int64_t)D * H * W;
    const int64_t sp = ((int64_t)z * H + y) * W + x;
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
        float *p = acc + ((int64_t)b * K + k) * plane + sp;
        const float v = *p / count;
        *p = v;
        if (dst) {          // running mean over the mirrored passes: dst (+)= scale * stitched
            float *d = dst + ((int64_t)b * K + k) * plane + sp;
            *d = dst_add ? fmaf(dst_scale, v, *d) : dst_scale * v;
        }
        if (v > best) { best = v; arg = k; }
    }
    if (labels) labels[(int64_t)b * plane + sp] = (uint8_t)arg;
}

}  // namespace wf

using namespace wf;

static inline unsigned blocks_for(int64_t total) { return (unsigned)((total + 255) / 256); }

extern "C" int wf_sw_gather(const float *vol, void *win, const int32_t *starts, int nwin, int dtype, int channels_last,
                            int C, int D, int H, int W, int r0, int r1, int r2, int flip, void *stream) {
    if (!vol || !win || !starts) return WF_ERR_NULL_POINTER;
    if (flip < 0 || flip > 7) return WF_ERR_BAD_SHAPE;
    if (nwin <= 0 || C <= 0 || r0 <= 0 || r1 <= 0 || r2 <= 0 || r0 > D || r1 > H || r2 > W) return WF_ERR_BAD_SHAPE;
    const int64_t total = (int64_t)nwin * C * r0 * r1 * r2;
    cudaStream_t st = (cudaStream_t)stream;
    if (channels_last && C == 4 && nwin <= 65535 && (int64_t)r0 * r1 * r2 < 0x7fffffffLL &&
        (reinterpret_cast<uintptr_t>(win) & 15u) == 0) {
        const dim3 grid((unsigned)(((int64_t)r0 * r1 * r2 + 255) / 256), (unsigned)nwin);
        if (dtype == WF_F32) sw_gather_voxel_kernel<float, 4><<<grid, 256, 0, st>>>(vol, (float *)win, starts, D, H, W, r0, r1, r2, flip);
        else if (dtype == WF_BF16) sw_gather_voxel_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(vol, (__nv_bfloat16 *)win, starts, D, H, W, r0, r1, r2, flip);
        else if (dtype == WF_F16) sw_gather_voxel_kernel<__half, 4><<<grid, 256, 0, st>>>(vol, (__half *)win, starts, D, H, W, r0, r1, r2, flip);
        else return WF_ERR_BAD_DTYPE;
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
#define WF_G(T_, CL_) sw_gather_kernel<T_, CL_><<<blocks_for(total), 256, 0, st>>>(vol, (T_ *)win, starts, total, C, D, H, W, r0, r1, r2, flip)
    if (dtype == WF_F32) { if (channels_last) WF_G(float, true); else WF_G(float, false); }
    else if (dtype == WF_BF16) { if (channels_last) WF_G(__nv_bfloat16, true); else WF_G(__nv_bfloat16, false); }
    else if (dtype == WF_F16) { if (channels_last) WF_G(__half, true); else WF_G(__half, false); }
    else return WF_ERR_BAD_DTYPE;
#undef WF_G
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_sw_accumulate(const void *seg, float *acc, const int32_t *starts, const float *gz, const float *gy,
                                const float *gx, float floor_w, int nwin, int dtype, int channels_last, int K, int D,
                                int H, int W, int r0, int r1, int r2, int flip, void *stream) {
    if (!seg || !acc || !starts || !gz || !gy || !gx) return WF_ERR_NULL_POINTER;
    if (flip < 0 || flip > 7) return WF_ERR_BAD_SHAPE;
    if (nwin <= 0 || K <= 0 || r0 <= 0 || r1 <= 0 || r2 <= 0 || r0 > D || r1 > H || r2 > W) return WF_ERR_BAD_SHAPE;
    const int64_t total = (int64_t)nwin * K * r0 * r1 * r2;
    cudaStream_t st = (cudaStream_t)stream;
    if (channels_last && K == 4 && nwin <= 65535 && (int64_t)r0 * r1 * r2 < 0x7fffffffLL &&
        (reinterpret_cast<uintptr_t>(seg) & 15u) == 0) {
        const dim3 grid((unsigned)(((int64_t)r0 * r1 * r2 + 255) / 256), (unsigned)nwin);
        if (dtype == WF_F32) sw_accumulate_voxel_kernel<float, 4><<<grid, 256, 0, st>>>((const float *)seg, acc, starts, gz, gy, gx, floor_w, D, H, W, r0, r1, r2, flip);
        else if (dtype == WF_BF16) sw_accumulate_voxel_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)seg, acc, starts, gz, gy, gx, floor_w, D, H, W, r0, r1, r2, flip);
        else if (dtype == WF_F16) sw_accumulate_voxel_kernel<__half, 4><<<grid, 256, 0, st>>>((const __half *)seg, acc, starts, gz, gy, gx, floor_w, D, H, W, r0, r1, r2, flip);
        else return WF_ERR_BAD_DTYPE;
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
#define WF_A(T_, CL_) sw_accumulate_kernel<T_, CL_><<<blocks_for(total), 256, 0, st>>>((const T_ *)seg, acc, starts, gz, gy, gx, floor_w, total, K, D, H, W, r0, r1, r2, flip)
    if (dtype == WF_F32) { if (channels_last) WF_A(float, true); else WF_A(float, false); }
    else if (dtype == WF_BF16) { if (channels_last) WF_A(__nv_bfloat16, true); else WF_A(__nv_bfloat16, false); }
    else if (dtype == WF_F16) { if (channels_last) WF_A(__half, true); else WF_A(__half, false); }
    else return WF_ERR_BAD_DTYPE;
#undef WF_A
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_sw_finalize(float *acc, uint8_t *labels, const int32_t *all_starts, int nall, const float *gz,
                              const float *gy, const float *gx, float floor_w, int Bv, int K, int D, int H, int W,
                              int r0, int r1, int r2, int z_begin, int z_end, int flip, float *dst, float dst_scale,
                              int dst_add, void *stream) {
    if (!acc || !all_starts || !gz || !gy || !gx) return WF_ERR_NULL_POINTER;
    // the table is staged in static-limit shared memory (16 B per entry): 3072 entries = 48 KB.  Volumes of one call share
    // their geometry, so callers pass ONE volume's windows with slot -1 (18 entries for the BraTS case) whatever Bv is.
    if (nall <= 0 || nall > 3072 || Bv <= 0 || K <= 0) return WF_ERR_BAD_SHAPE;
    if (z_begin < 0 || z_end > D || z_begin >= z_end || flip < 0 || flip > 7) return WF_ERR_BAD_SHAPE;
    const int64_t total = (int64_t)Bv * (z_end - z_begin) * H * W;
    sw_finalize_kernel<<<blocks_for(total), 256, (size_t)nall * 16, (cudaStream_t)stream>>>(
        acc, labels, all_starts, nall, gz, gy, gx, floor_w, total, K, D, H, W, r0, r1, r2, z_begin, z_end, flip, dst, dst_scale, dst_add);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
