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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// bf16 fast path: shared-memory tile + mixed-precision FMA.
//   CTA = output tile of TZ x TY x TX = 4 x 4 x 16 voxels x 64 channels; the 6 x 6 x 18 x 64 haloed input tile (81 KB,
//   zero padded at the volume border) is staged once with cp.async; lane = one bf16x2 channel pair (a warp's 32 lanes
//   read 128 contiguous bytes: conflict-free LDS.32), warp = one pair of adjacent output rows (z, y..y+1) walking along
//   x with a 12-row x 3-column register window: 12 LDS.32 per 2 output voxels instead of 54 LDG.128 per 4.
//   Taps are kept as 27 packed bf16x2 registers and applied with FHFMA (fma.rn.f32.bf16: both 16-bit operands are
//   widened inside the fp32 FMA, no unpack instructions); accumulation is fp32.
constexpr int kDwTZ = 4, kDwTY = 4, kDwTX = 16, kDwCG = 64;
constexpr int kDwHZ = kDwTZ + 2, kDwHY = kDwTY + 2, kDwHX = kDwTX + 2;
constexpr int kDwSmem = kDwHZ * kDwHY * kDwHX * kDwCG * 2;  // 82944

template <bool F16> __device__ __forceinline__ void fhfma2(float &a0, float &a1, uint32_t v, uint32_t w) {
    // a0 += lo(v) * lo(w); a1 += hi(v) * hi(w)   (16-bit x 16-bit + fp32: FHFMA.BF16 / FHFMA.F16)
    if constexpr (F16)
        asm("{\n\t.reg .b16 vl, vh, wl, wh;\n\tmov.b32 {vl, vh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
            "fma.rn.f32.f16 %0, vl, wl, %0;\n\tfma.rn.f32.f16 %1, vh, wh, %1;\n\t}"
            : "+f"(a0), "+f"(a1)
            : "r"(v), "r"(w));
    else
        asm("{\n\t.reg .b16 vl, vh, wl, wh;\n\tmov.b32 {vl, vh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, vl, wl, %0;\n\tfma.rn.f32.bf16 %1, vh, wh, %1;\n\t}"
            : "+f"(a0), "+f"(a1)
            : "r"(v), "r"(w));
}
// two fp32 -> one packed pair of the 16-bit storage type, and back
template <bool F16> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (F16) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}
template <bool F16> __device__ __forceinline__ float2 unpack2(uint32_t w) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2 *>(&w));
    else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// STATS: additionally reduce per-(sample, channel) sum / sum of squares of the ROUNDED outputs into `sums` (fp64 [B][C][2],
// zeroed by the launcher) - the statistics of the GroupNorm / InstanceNorm that follows (ProjectionUpsample.norm), so the
// separate statistics pass over the result disappears.
template <bool STATS, bool F16>
__global__ void __launch_bounds__(256, 2) dwconv3d_bf16_tile_kernel(const uint16_t *__restrict__ x,
                                                                    const float *__restrict__ w27,
                                                                    const float *__restrict__ bias,
                                                                    uint16_t *__restrict__ y, int D, int H, int W,
                                                                    int C, int tiles_x, int tiles_y, int tiles_z,
                                                                    double *__restrict__ sums) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[STATS ? 8 * 32 * 4 : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    const int tz = t % tiles_z;
    const int64_t b = t / tiles_z;
    const int c0 = blockIdx.y * kDwCG;
    const int x0 = tx * kDwTX, y0 = ty * kDwTY, z0 = tz * kDwTZ;
    const int cchunks = min(8, (C - c0) >> 3);  // live 16-byte channel chunks of this group

    // ---- stage the haloed tile: 6 * 6 * 18 voxels x 8 chunks of 16 bytes ----
    const uint16_t *xb = x + (int64_t)b * D * H * W * C + c0;
    for (int i = tid; i < kDwHZ * kDwHY * kDwHX * 8; i += 256) {
        const int ch = i & 7, v = i >> 3;
        const int xi = v % kDwHX, yi = (v / kDwHX) % kDwHY, zi = v / (kDwHX * kDwHY);
        const int gx = x0 - 1 + xi, gy = y0 - 1 + yi, gz = z0 - 1 + zi;
        uint8_t *dst = smem + (size_t)v * 128 + ch * 16;
        if (ch < cchunks && (unsigned)gx < (unsigned)W && (unsigned)gy < (unsigned)H && (unsigned)gz < (unsigned)D)
            cp_async16(dst, xb + (((int64_t)gz * H + gy) * W + gx) * C + ch * 8);
        else
            *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    // ---- this lane's taps (two channels) while the copies are in flight ----
    const int c = c0 + 2 * lane;
    const bool live_c = c < C;
    uint32_t wt[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const float2 f = live_c ? __ldg(reinterpret_cast<const float2 *>(w27 + (int64_t)k * C + c)) : make_float2(0.f, 0.f);
        wt[k] = pack2<F16>(f.x, f.y);
    }
    float b0 = 0.f, b1 = 0.f;
    if (bias != nullptr && live_c) {
        const float2 f = __ldg(reinterpret_cast<const float2 *>(bias + c));
        b0 = f.x; b1 = f.y;
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- warp = output rows (z0 + zl, y0 + 2*yp) and (.., y0 + 2*yp + 1) ----
    const int zl = warp >> 1, yp = warp & 1;
    const uint32_t *tile = reinterpret_cast<const uint32_t *>(smem) + lane;
    // input row r of the window: zi = zl + r / 4, yi = 2 * yp + r % 4   (halo coordinates)
    auto ld = [&](int r, int xi) -> uint32_t {
        const int zi = zl + (r >> 2), yi = 2 * yp + (r & 3);
        return tile[((zi * kDwHY + yi) * kDwHX + xi) * 32];
    };
    uint32_t win[12][3];
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        win[r][0] = ld(r, 0);
        win[r][1] = ld(r, 1);
    }
    const int gz = z0 + zl, gy = y0 + 2 * yp;
    uint16_t *yrow = y + ((((int64_t)b * D + gz) * H + gy) * W + x0) * C + c;
    const bool row0 = gz < D && gy < H && live_c, row1 = gz < D && gy + 1 < H && live_c;
    float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;   // STATS: this lane's two channels over its outputs
#pragma unroll
    for (int j = 0; j < kDwTX; ++j) {
#pragma unroll
        for (int r = 0; r < 12; ++r) win[r][(j + 2) % 3] = ld(r, j + 2);
        float a00 = b0, a01 = b1, a10 = b0, a11 = b1;   // output row 0 / 1, channel 0 / 1
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t wv = wt[dz * 9 + dy * 3 + dx];
                    fhfma2<F16>(a00, a01, win[dz * 4 + dy][(j + dx) % 3], wv);       // output row 0 reads window rows dy
                    fhfma2<F16>(a10, a11, win[dz * 4 + dy + 1][(j + dx) % 3], wv);   // output row 1 reads window rows dy + 1
                }
        if (x0 + j < W) {
            if (row0) {
                const uint32_t h = pack2<F16>(a00, a01);
                *reinterpret_cast<uint32_t *>(yrow + (int64_t)j * C) = h;
                if constexpr (STATS) {
                    const float2 f = unpack2<F16>(h);
                    st_s0 += f.x; st_s1 += f.y; st_q0 = fmaf(f.x, f.x, st_q0); st_q1 = fmaf(f.y, f.y, st_q1);
                }
            }
            if (row1) {
                const uint32_t h = pack2<F16>(a10, a11);
                *reinterpret_cast<uint32_t *>(yrow + ((int64_t)W + j) * C) = h;
                if constexpr (STATS) {
                    const float2 f = unpack2<F16>(h);
                    st_s0 += f.x; st_s1 += f.y; st_q0 = fmaf(f.x, f.x, st_q0); st_q1 = fmaf(f.y, f.y, st_q1);
                }
            }
        }
    }
    if constexpr (STATS) {
        float *r = s_red + (warp * 32 + lane) * 4;
        r[0] = st_s0; r[1] = st_s1; r[2] = st_q0; r[3] = st_q1;
        __syncthreads();
        if (tid < 2 * kDwCG) {              // thread = (channel of the group, sum | sum of squares)
            const int ch = tid >> 1, which = tid & 1;
            float a = 0.f;
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) a += s_red[(wv * 32 + (ch >> 1)) * 4 + which * 2 + (ch & 1)];
            if (c0 + ch < C) atomicAdd(sums + ((int64_t)b * C + c0 + ch) * 2 + which, (double)a);
        }
    }
}

__global__ void dwconv_stats_finalize_kernel(const double *__restrict__ sums, float *__restrict__ mr, int n, double inv_s,
                                             double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = sums[2 * i] * inv_s;
    double var = sums[2 * i + 1] * inv_s - m * m;
    var = var < 0.0 ? 0.0 : var;
    mr[2 * i] = (float)m;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

template <bool STATS, bool F16>
static int dwconv_bf16_tile_launch(const uint16_t *x, const float *w27, const float *bias, uint16_t *y, int B,
                                   int D, int H, int W, int C, cudaStream_t st, double *sums = nullptr) {
    static unsigned long long attr_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attr_done)) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(dwconv3d_bf16_tile_kernel<STATS, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem));
    }
    const int tiles_x = (W + kDwTX - 1) / kDwTX, tiles_y = (H + kDwTY - 1) / kDwTY, tiles_z = (D + kDwTZ - 1) / kDwTZ;
    const int64_t tiles = (int64_t)B * tiles_x * tiles_y * tiles_z;
    if (tiles > 0x7fffffff) return WF_ERR_UNSUPPORTED;
    dim3 grid((unsigned)tiles, (unsigned)((C + kDwCG - 1) / kDwCG));
    dwconv3d_bf16_tile_kernel<STATS, F16><<<grid, 256, kDwSmem, st>>>(x, w27, bias, y, D, H, W, C, tiles_x, tiles_y, tiles_z, sums);
    WF_LAUNCH_CHECK();
    return WF_OK;
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
    if (dtype == WF_BF16 || dtype == WF_F16) {
        // shared-memory tile kernel when the geometry gives it enough work; the register-tiled kernel otherwise
        const char *impl = getenv("WF_DWCONV_IMPL");
        const bool force_old = impl != nullptr && impl[0] == 'r';
        if (!force_old && C % 8 == 0 && W >= 8 && wf::aligned16(x) && wf::aligned16(w27) && (bias == nullptr || wf::aligned16(bias))) {
            if (dtype == WF_F16)
                return wf::dwconv_bf16_tile_launch<false, true>((const uint16_t *)x, w27, bias, (uint16_t *)y, B, D, H, W, C, st);
            return wf::dwconv_bf16_tile_launch<false, false>((const uint16_t *)x, w27, bias, (uint16_t *)y, B, D, H, W, C, st);
        }
        if (dtype == WF_F16)
            return wf::dwconv_launch<__half>((const __half *)x, w27, bias, (__half *)y, B, D, H, W, C, st);
        return wf::dwconv_launch<__nv_bfloat16>((const __nv_bfloat16 *)x, w27, bias, (__nv_bfloat16 *)y, B, D, H, W, C, st);
    }
    return WF_ERR_BAD_DTYPE;
}

extern "C" int wf_dwconv3d_ndhwc_stats(const void *x, const float *w27, const float *bias, void *y, double *sums,
                                       float *mean_rstd, float eps, int dtype, int B, int D, int H, int W, int C,
                                       void *stream) {
    if (!x || !w27 || !y || !sums || !mean_rstd) return WF_ERR_NULL_POINTER;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0) return WF_ERR_BAD_SHAPE;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_UNSUPPORTED;
    if (C % 8 != 0 || W < 8) return WF_ERR_UNSUPPORTED;
    if (!wf::aligned16(x) || !wf::aligned16(w27) || (bias != nullptr && !wf::aligned16(bias))) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    WF_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * C, st));
    const int rc = dtype == WF_F16
        ? wf::dwconv_bf16_tile_launch<true, true>((const uint16_t *)x, w27, bias, (uint16_t *)y, B, D, H, W, C, st, sums)
        : wf::dwconv_bf16_tile_launch<true, false>((const uint16_t *)x, w27, bias, (uint16_t *)y, B, D, H, W, C, st, sums);
    if (rc != WF_OK) return rc;
    const int n = B * C;
    wf::dwconv_stats_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(sums, mean_rstd, n, 1.0 / ((double)D * H * W), (double)eps);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
