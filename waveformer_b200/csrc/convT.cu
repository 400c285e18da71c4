// ConvTranspose3d with kernel 2, stride 2 (no overlap, no bias) on channels-last activations as ONE tcgen05 GEMM whose
// epilogue scatters each output voxel straight into (a channel slice of) the decoder's concatenation buffer.
//
// Reference: UnetrUpBlock.transp_conv + torch.cat((out, skip), dim=1) (monai/networks/blocks/unetr_block.py:57-86), used
// by Waveformer.decoder1 (network_models/network_backbone.py:352-361,405): 144 -> 48 channels, 64^3 -> 128^3.
// With stride == kernel every output voxel (2z+dz, 2y+dy, 2x+dx) depends on exactly one input voxel:
//     y[b, 2z+dz, 2y+dy, 2x+dx, co] = sum_ci x[b, z, y, x, ci] * w[ci, co, dz, dy, dx]
// i.e. D[M = input voxels, N = 8 * Cout] = X[M, Cin] * Wp[N, Cin]^T with n = (dz*4 + dy*2 + dx) * Cout + co.
// The library path was a strided-dgrad kernel (0.28 ms) plus a 0.45 ms copy into the concat buffer; here the output is
// written once, in place.
//   grid (ceil(M / 128), N / NT); CTA: A tile [128 x Cin] and B tile [NT x Cin] staged as K-major no-swizzle UMMA images,
//   Cin / 16 x tcgen05.mma 128 x NT x 16, fp32 accumulators in TMEM; thread = input voxel (TMEM lane), 16 channels
//   (32 bytes, one full sector) per store.
#include "tc_common.cuh"
#include "wf_common.cuh"

namespace wf {

using namespace tc;

__global__ void __launch_bounds__(128) convT_k2s2_kernel(const __nv_bfloat16 *__restrict__ x, const uint16_t *__restrict__ wp,
                                                         __nv_bfloat16 *__restrict__ y, int64_t M, int K, int NT, int Cout,
                                                         int D, int H, int W, int64_t xs, int64_t ys, uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kchunks = K >> 3;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)kchunks * 2048;
    const int64_t m = (int64_t)blockIdx.x * 128 + tid;
    const int n0 = blockIdx.y * NT;
    const bool live = m < M;

    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(x + (live ? m : 0) * xs);
        for (int kc = 0; kc < kchunks; ++kc)
            *reinterpret_cast<uint4 *>(sA + (size_t)kc * 2048 + tid * 16) = live ? __ldg(src + kc) : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int idx = tid; idx < NT * kchunks; idx += 128) {
        const int r = idx % NT, kc = idx / NT;
        *reinterpret_cast<uint4 *>(sB + ((size_t)kc * NT + r) * 16) =
            __ldg(reinterpret_cast<const uint4 *>(wp + (int64_t)(n0 + r) * K) + kc);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_bf16(128, NT, false);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < (K >> 4); ++ks)
            mma_ss(tmem, smem_desc(a0 + ks * 2 * 2048, 2048, 128), smem_desc(b0 + ks * 2 * NT * 16, NT * 16, 128), idesc,
                   ks > 0 ? 1u : 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    // input voxel -> (b, z, y, x)
    int xx = 0, yy = 0, zz = 0;
    int64_t b = 0;
    if (live) {
        xx = (int)(m % W);
        int64_t t = m / W;
        yy = (int)(t % H); t /= H;
        zz = (int)(t % D);
        b = t / D;
    }
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int H2 = 2 * H, W2 = 2 * W;
    for (int c = 0; c < NT; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + c, r);
        tmem_wait_ld();
        if (!live) continue;
        const int n = n0 + c;
        const int pos = n / Cout, co = n - pos * Cout;
        const int dz = pos >> 2, dy = (pos >> 1) & 1, dx = pos & 1;
        const int64_t ov = ((b * (2 * D) + 2 * zz + dz) * H2 + 2 * yy + dy) * (int64_t)W2 + 2 * xx + dx;
        uint4 lo, hi;
        lo.x = pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3]));
        lo.z = pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7]));
        hi.x = pack_bf16(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_bf16(__uint_as_float(r[10]), __uint_as_float(r[11]));
        hi.z = pack_bf16(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_bf16(__uint_as_float(r[14]), __uint_as_float(r[15]));
        uint4 *dst = reinterpret_cast<uint4 *>(y + ov * ys + co);
        dst[0] = lo;
        dst[1] = hi;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace wf

using namespace wf;

extern "C" int wf_convtranspose3d_k2s2_ndhwc(const void *x, const void *wpack, void *y, int dtype, int B, int D, int H,
                                             int W, int Cin, int Cout, int64_t x_vox_stride, int64_t y_vox_stride,
                                             void *stream) {
    if (!x || !wpack || !y) return WF_ERR_NULL_POINTER;
    if (dtype != WF_BF16) return WF_ERR_BAD_DTYPE;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return WF_ERR_BAD_SHAPE;
    if (Cin % 16 || Cin > 512 || Cout % 16 || x_vox_stride < Cin || y_vox_stride < Cout || x_vox_stride % 8 || y_vox_stride % 8)
        return WF_ERR_BAD_SHAPE;
    if (!aligned16(x) || !aligned16(wpack) || !aligned16(y)) return WF_ERR_MISALIGNED;
    const int N = 8 * Cout;
    // N tile: the largest multiple of 16 (<= 128) that divides N; Cout % 16 == 0 keeps every 16-channel store group
    // inside one output position
    int NT = 0;
    for (int nt = 128; nt >= 16; nt -= 16)
        if (N % nt == 0) { NT = nt; break; }
    if (!NT) return WF_ERR_UNSUPPORTED;
    const int kchunks = Cin / 8;
    const size_t smem = (size_t)kchunks * 2048 + (size_t)kchunks * NT * 16;
    if (smem > 200 * 1024) return WF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    uint32_t cols = 32;
    while ((int)cols < NT) cols <<= 1;
    const int64_t M = (int64_t)B * D * H * W;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)(N / NT));
    convT_k2s2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, (const uint16_t *)wpack,
                                                                (__nv_bfloat16 *)y, M, Cin, NT, Cout, D, H, W, x_vox_stride,
                                                                y_vox_stride, cols);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
