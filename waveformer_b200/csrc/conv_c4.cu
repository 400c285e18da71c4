// 3x3x3 convolution (padding 1, stride 1, no bias) of a 4-channel channels-last volume, fused with the block's 1x1x1
// shortcut convolution and with the InstanceNorm statistics of both results (tcgen05 / TMEM, sm_100a).
//
// Reference: the first residual block of the U-Net, Waveformer.encoder1 = UnetrBasicBlock(in_chans -> 48, res_block)
// (network_models/network_backbone.py:247-255,386) -> MONAI UnetResBlock.forward
// (monai/networks/blocks/dynunet_block.py:98-111): conv1 (3^3, in -> out), norm1, lrelu, conv2, norm2, and the shortcut
// conv3 (1^3, in -> out), norm3.  With in_chans = 4 the library convolution runs at 13 TFLOP/s (3.3 ms per batch-2
// window, 15 % of the forward): K = 27 * 4 is too thin for its tiling.  Here it is an implicit GEMM whose im2col rows
// are gathered by the threads themselves:
//   tile = 128 consecutive voxels (one TMEM lane each); thread = voxel: 27 neighbour loads of 8 bytes (4 x bf16; fp32
//   input is converted on the fly), written as the K-major no-swizzle UMMA A image [14 chunks][128 rows][8];
//   B = packed weights [N][112] (k = tap * 4 + channel, zero padded; rows >= n0 hold the 1^3 shortcut's weights in the
//   centre tap), resident in shared memory for the CTA's lifetime; 7 x tcgen05.mma 128 x N x 16 into TMEM;
//   epilogue: TMEM -> bf16 -> shared staging tile -> coalesced 16-byte stores to the two outputs, and per-channel
//   sum / sum of squares of the ROUNDED outputs (thread t < N owns channel t; fp32 per tile, fp64 per CTA, one fp64
//   atomicAdd per CTA and channel at the end), so InstanceNorm needs no separate statistics pass.  y1 == NULL: only the
//   statistics of the shortcut are produced - its values are a 4-term dot product per channel that the block's last
//   InstanceNorm pass recomputes from the input voxel (wf_instnorm_apply_shortcut4_ndhwc), which saves one write and one
//   read of a full-resolution 48-channel tensor.
// Persistent CTAs (grid = min(tiles, 4 per SM)); traffic = one read of the input (L2-resident) + one write of the output.
#include <type_traits>

#include "tc_common.cuh"
#include "wf_common.cuh"

namespace wf {

using namespace tc;

constexpr int kC4K = 112;        // 27 taps * 4 channels, padded to 7 k-steps of 16
constexpr int kC4Chunks = 14;    // 16-byte K chunks per row

struct C4Geom {
    int B, D, H, W;
    int64_t total;  // B * D * H * W
};


// dynamic smem: [B image 14 * N * 16][A image 14 * 2048, re-used as the bf16 staging tile 128 * (N + 8) * 2 once the MMAs
// of the tile have completed]
// TIN: float (converted to the operand format while gathered) or uint16_t (already in the operand format);
// F16: operand / output format fp16 instead of bf16.
template <typename TIN, bool F16>
__global__ void __launch_bounds__(128, 4) conv3d_c4_kernel(const TIN *__restrict__ x, const uint16_t *__restrict__ wpack,
                                                           uint16_t *__restrict__ y0, int64_t ys0, int n0,
                                                           uint16_t *__restrict__ y1, int64_t ys1, int n1,
                                                           double *__restrict__ sums0, double *__restrict__ sums1,
                                                           C4Geom g, int64_t ntiles, uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int N = n0 + n1;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t *sB = smem;
    uint8_t *sA = smem + (size_t)kC4Chunks * N * 16;
    // staging: two DENSE sub-tiles [128][n0] and [128][n1] (so that, for dense outputs, tile -> global is a linear copy)
    uint16_t *sOut0 = reinterpret_cast<uint16_t *>(sA);
    uint16_t *sOut1 = sOut0 + 128 * n0;

    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    // weights: [N][112] bf16 row-major in global -> canonical K-major image [chunk][row][8]
    for (int idx = tid; idx < N * kC4Chunks; idx += 128) {
        const int r = idx % N, kc = idx / N;
        *reinterpret_cast<uint4 *>(sB + ((size_t)kc * N + r) * 16) =
            __ldg(reinterpret_cast<const uint4 *>(wpack + (int64_t)r * kC4K) + kc);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = instr_desc_h16<F16>(128, N, false);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int64_t S = (int64_t)g.D * g.H * g.W;

    // statistics: thread t < N owns channel pair (t % (N/2)) for the rows of half (t / (N/2)) of every tile; fp32 per
    // tile, fp64 across this CTA's tiles, one batch element at a time
    const int half_n = N >> 1;
    const int cpair = tid % half_n, rhalf = tid / half_n;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // sum c, sum c+1, sumsq c, sumsq c+1
    int64_t acc_b = -1;
    auto flush = [&]() {
        if (tid < N && acc_b >= 0) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 2 * cpair + e;
                double *dst = c < n0 ? sums0 + (acc_b * n0 + c) * 2 : sums1 + (acc_b * n1 + (c - n0)) * 2;
                atomicAdd(dst, acc[e]);
                atomicAdd(dst + 1, acc[2 + e]);
            }
        }
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    };

    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t v = tile * 128 + tid;
        const bool live = v < g.total;
        // ---- gather this voxel's 27 x 4 neighbourhood into the A image ----
        {
            int xx = 0, yy = 0, zz = 0;
            int64_t b = 0;
            if (live) {   // total < 2^31 (checked by the host wrapper): 32-bit divisions, ~5x cheaper than 64-bit ones
                const uint32_t v32 = (uint32_t)v;
                uint32_t t = v32 / (uint32_t)g.W;
                xx = (int)(v32 - t * (uint32_t)g.W);
                uint32_t t2 = t / (uint32_t)g.H;
                yy = (int)(t - t2 * (uint32_t)g.H);
                const uint32_t t3 = t2 / (uint32_t)g.D;
                zz = (int)(t2 - t3 * (uint32_t)g.D);
                b = t3;
            }
            const TIN *base = x + (b * S) * 4;
            uint2 tap[28];
            // unconditional loads (coordinates clamped into the volume, result masked afterwards) so that they are in
            // flight together instead of one predicated load at a time: all 27 for bf16 input (54 registers), nine per z
            // plane for fp32 input (36 registers)
            if constexpr (sizeof(TIN) == 2) {
                // per-axis clamped coordinates and validity, then 32-bit voxel offsets (total < 2^31): one IMAD.WIDE per load
                int zo[3], yo[3], xo[3];
                bool zok[3], yok[3], xok[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int z2 = zz + d - 1, y2 = yy + d - 1, x2 = xx + d - 1;
                    zok[d] = (unsigned)z2 < (unsigned)g.D; yok[d] = (unsigned)y2 < (unsigned)g.H; xok[d] = (unsigned)x2 < (unsigned)g.W;
                    zo[d] = (zok[d] ? z2 : zz) * g.H; yo[d] = yok[d] ? y2 : yy; xo[d] = xok[d] ? x2 : xx;
                }
                const uint2 *vb = reinterpret_cast<const uint2 *>(base);
                uint2 raw[27];
#pragma unroll
                for (int t = 0; t < 27; ++t) {
                    const int dz = t / 9, dy = (t / 3) % 3, dx = t % 3;
                    raw[t] = __ldg(vb + (uint32_t)((zo[dz] + yo[dy]) * g.W + xo[dx]));
                }
#pragma unroll
                for (int t = 0; t < 27; ++t) {
                    const int dz = t / 9, dy = (t / 3) % 3, dx = t % 3;
                    tap[t] = (live && zok[dz] && yok[dy] && xok[dx]) ? raw[t] : make_uint2(0u, 0u);
                }
            } else {
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz) {
                    float4 raw[9];
                    bool in[9];
                    const int z2 = zz + dz;
                    const int zc = min(max(z2, 0), g.D - 1);
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int y2 = yy + dy;
                        const int yc = min(max(y2, 0), g.H - 1);
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int x2 = xx + dx;
                            const int xc = min(max(x2, 0), g.W - 1);
                            const int i = (dy + 1) * 3 + (dx + 1);
                            in[i] = live && z2 == zc && y2 == yc && x2 == xc;
                            raw[i] = __ldg(reinterpret_cast<const float4 *>(base + (((int64_t)zc * g.H + yc) * g.W + xc) * 4));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        tap[(dz + 1) * 9 + i] = in[i] ? make_uint2(pack_h16<F16>(raw[i].x, raw[i].y), pack_h16<F16>(raw[i].z, raw[i].w))
                                                      : make_uint2(0u, 0u);
                }
            }
            tap[27] = make_uint2(0u, 0u);
            __syncthreads();   // the previous tile's staging tile (same shared memory) has been stored and summed
#pragma unroll
            for (int kc = 0; kc < kC4Chunks; ++kc)
                *reinterpret_cast<uint4 *>(sA + kc * 2048 + tid * 16) =
                    make_uint4(tap[2 * kc].x, tap[2 * kc].y, tap[2 * kc + 1].x, tap[2 * kc + 1].y);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();   // A image complete
        if ((tid >> 5) == 0) {     // warp 0, converged: one elected lane issues (tc_common.cuh, "warp-uniform issue")
            tc_fence_after();
            const uint32_t al = smem_desc_lo(a0, 2048), bl = smem_desc_lo(b0, N * 16), dh = smem_desc_hi(128);
#pragma unroll
            for (int ks = 0; ks < kC4K / 16; ++ks)
                mma_ss_w(tmem, al + (uint32_t)(ks * 2 * 2048 / 16), dh, bl + (uint32_t)(ks * 2 * N), dh, idesc, ks > 0 ? 1u : 0u);
            mma_commit_w(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue: TMEM lane -> bf16 -> staging row ----
        for (int c = 0; c < N; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + lane_base + c, r);
            tmem_wait_ld();
            uint4 lo, hi;
            lo.x = pack_h16<F16>(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_h16<F16>(__uint_as_float(r[2]), __uint_as_float(r[3]));
            lo.z = pack_h16<F16>(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_h16<F16>(__uint_as_float(r[6]), __uint_as_float(r[7]));
            hi.x = pack_h16<F16>(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_h16<F16>(__uint_as_float(r[10]), __uint_as_float(r[11]));
            hi.z = pack_h16<F16>(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_h16<F16>(__uint_as_float(r[14]), __uint_as_float(r[15]));
            // 16 columns never straddle the two outputs (n0 % 16 == 0 is required by the host wrapper when n1 > 0)
            uint4 *dst = reinterpret_cast<uint4 *>(c < n0 ? sOut0 + (size_t)tid * n0 + c : sOut1 + (size_t)tid * n1 + (c - n0));
            dst[0] = lo;
            dst[1] = hi;
        }
        tc_fence_before();
        __syncthreads();   // staging complete; TMEM is free for the next tile
        // ---- coalesced stores: voxel rows of n0 (y0) and n1 (y1) channels ----
        const int64_t v0 = tile * 128;
        const int rows = (int)min((int64_t)128, g.total - v0);
        if (ys0 == n0 && (n1 == 0 || y1 == nullptr || ys1 == n1)) {   // dense outputs: the tile is contiguous in global memory too
            const uint4 *s0 = reinterpret_cast<const uint4 *>(sOut0);
            uint4 *d0 = reinterpret_cast<uint4 *>(y0 + v0 * n0);
            for (int i = tid; i < rows * (n0 >> 3); i += 128) d0[i] = s0[i];
            if (n1 > 0 && y1 != nullptr) {
                const uint4 *s1 = reinterpret_cast<const uint4 *>(sOut1);
                uint4 *d1 = reinterpret_cast<uint4 *>(y1 + v0 * n1);
                for (int i = tid; i < rows * (n1 >> 3); i += 128) d1[i] = s1[i];
            }
        } else {
            const int per0 = n0 >> 3;
            for (int i = tid; i < rows * per0; i += 128) {
                const int r = i / per0, p = i % per0;
                *reinterpret_cast<uint4 *>(y0 + (v0 + r) * ys0 + p * 8) = *reinterpret_cast<const uint4 *>(sOut0 + (size_t)r * n0 + p * 8);
            }
            const int per1 = y1 != nullptr ? n1 >> 3 : 0;
            for (int i = tid; i < rows * per1; i += 128) {
                const int r = i / per1, p = i % per1;
                *reinterpret_cast<uint4 *>(y1 + (v0 + r) * ys1 + p * 8) = *reinterpret_cast<const uint4 *>(sOut1 + (size_t)r * n1 + p * 8);
            }
        }
        // ---- statistics of the rounded outputs: FHADD / FHFMA consume the bf16 halves directly ----
        if (tid < N) {
            // a tile may straddle two batch elements only if S % 128 != 0; handled row by row in that (rare) case
            const int64_t b_first = (uint32_t)v0 / (uint32_t)S, b_last = (uint32_t)(v0 + rows - 1) / (uint32_t)S;
            const int c2 = 2 * cpair;                                  // first channel of this thread's pair
            const uint32_t *col = c2 < n0 ? reinterpret_cast<const uint32_t *>(sOut0) + (c2 >> 1)
                                          : reinterpret_cast<const uint32_t *>(sOut1) + ((c2 - n0) >> 1);
            const int wpitch = (c2 < n0 ? n0 : n1) >> 1;
            const int r0 = rhalf * 64, r1 = min(rows, r0 + 64);
            if (b_first == b_last) {
                if (b_first != acc_b) { flush(); acc_b = b_first; }
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
                for (int r = r0; r < r1; ++r) {
                    const uint32_t w = col[(size_t)r * wpitch];
                    stat_h16x2<F16>(w, s0, s1, q0, q1);
                }
                acc[0] += (double)s0; acc[1] += (double)s1; acc[2] += (double)q0; acc[3] += (double)q1;
            } else {
                for (int r = r0; r < r1; ++r) {
                    const int64_t bb = (uint32_t)(v0 + r) / (uint32_t)S;
                    if (bb != acc_b) { flush(); acc_b = bb; }
                    const uint32_t w = col[(size_t)r * wpitch];
                    const float2 ff = unpack_h16<F16>(w);
                    const double f0 = (double)ff.x, f1 = (double)ff.y;
                    acc[0] += f0; acc[1] += f1; acc[2] += f0 * f0; acc[3] += f1 * f1;
                }
            }
        }
        // the next iteration's first __syncthreads orders these staging reads before the next gather overwrites them
    }
    flush();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// mean / rstd from raw (unshifted) fp64 sums: mr[2i] = mean, mr[2i+1] = 1 / sqrt(var + eps)
__global__ void stats_finalize_raw_kernel(const double *__restrict__ sums, float *__restrict__ mr, int n, double inv_s,
                                          double eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = sums[2 * i] * inv_s;
    double var = sums[2 * i + 1] * inv_s - m * m;
    var = var < 0.0 ? 0.0 : var;
    mr[2 * i] = (float)m;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

}  // namespace wf

using namespace wf;

extern "C" int wf_conv3d_c4_in_stats(const void *x, int x_dtype, int op_dtype, const void *wpack, void *y0, int64_t y0_vox_stride,
                                     int n0, void *y1, int64_t y1_vox_stride, int n1, double *sums0, double *sums1,
                                     float *mean_rstd0, float *mean_rstd1, float eps, int B, int D, int H, int W,
                                     void *stream) {
    if (!x || !wpack || !y0 || !sums0 || !mean_rstd0) return WF_ERR_NULL_POINTER;
    if (n1 > 0 && (!sums1 || !mean_rstd1)) return WF_ERR_NULL_POINTER;   // y1 == NULL: statistics of the second output only
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return WF_ERR_BAD_SHAPE;
    const int N = n0 + n1;
    if (n0 <= 0 || n1 < 0 || n0 % 8 || n1 % 8 || N % 16 || N > 128 || (n1 > 0 && n0 % 16)) return WF_ERR_BAD_SHAPE;
    if (y0_vox_stride < n0 || (n1 > 0 && y1 && y1_vox_stride < n1) || y0_vox_stride % 8 || (n1 > 0 && y1 && y1_vox_stride % 8)) return WF_ERR_BAD_SHAPE;
    if (op_dtype != WF_BF16 && op_dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (x_dtype != WF_F32 && x_dtype != op_dtype) return WF_ERR_BAD_DTYPE;
    if (!aligned16(x) || !aligned16(wpack) || !aligned16(y0) || (n1 > 0 && y1 && !aligned16(y1))) return WF_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    C4Geom g;
    g.B = B; g.D = D; g.H = H; g.W = W;
    g.total = (int64_t)B * D * H * W;
    if (g.total >= 0x7fffffffLL) return WF_ERR_UNSUPPORTED;   // the kernel uses 32-bit voxel arithmetic
    const int64_t ntiles = (g.total + 127) / 128;
    const size_t stage = (size_t)128 * N * 2, aimg = (size_t)kC4Chunks * 2048;
    const size_t smem = (size_t)kC4Chunks * N * 16 + (stage > aimg ? stage : aimg);
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    static unsigned long long attrs_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attrs_done)) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_c4_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_c4_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_c4_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_c4_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    WF_CUDA_CHECK(cudaMemsetAsync(sums0, 0, sizeof(double) * 2 * (size_t)B * n0, st));
    if (n1 > 0) WF_CUDA_CHECK(cudaMemsetAsync(sums1, 0, sizeof(double) * 2 * (size_t)B * n1, st));
    const int per_sm = (int)min((size_t)4, (size_t)(200 * 1024) / smem);
    const int grid = (int)min(ntiles, (int64_t)kNumSMs * (per_sm < 1 ? 1 : per_sm));
#define WF_C4(TIN_, F16_)                                                                                              \
    conv3d_c4_kernel<TIN_, F16_><<<grid, 128, smem, st>>>((const TIN_ *)x, (const uint16_t *)wpack, (uint16_t *)y0, y0_vox_stride, \
                                                          n0, (uint16_t *)y1, y1_vox_stride, n1, sums0, sums1, g, ntiles, cols)
    if (op_dtype == WF_F16) {
        if (x_dtype == WF_F32) WF_C4(float, true); else WF_C4(uint16_t, true);
    } else {
        if (x_dtype == WF_F32) WF_C4(float, false); else WF_C4(uint16_t, false);
    }
#undef WF_C4
    WF_LAUNCH_CHECK();
    const double inv_s = 1.0 / ((double)D * H * W);
    stats_finalize_raw_kernel<<<(B * n0 + 127) / 128, 128, 0, st>>>(sums0, mean_rstd0, B * n0, inv_s, (double)eps);
    WF_LAUNCH_CHECK();
    if (n1 > 0) {
        stats_finalize_raw_kernel<<<(B * n1 + 127) / 128, 128, 0, st>>>(sums1, mean_rstd1, B * n1, inv_s, (double)eps);
        WF_LAUNCH_CHECK();
    }
    return WF_OK;
}
