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

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(128) convT_k2s2_kernel(const __nv_bfloat16 *__restrict__ x, const uint16_t *__restrict__ wp,
                                                         __nv_bfloat16 *__restrict__ y, int64_t M, int K, int NT, int Cout,
                                                         int D, int H, int W, int64_t xs, int64_t ys, uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int64_t s_ov[128];   // output voxel index of position (0, 0, 0) for each input voxel of the tile; -1 = none
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kchunks = K >> 3;
    {
        const int64_t mm = (int64_t)blockIdx.x * 128 + tid;
        int64_t ov = -1;
        if (mm < M) {   // M < 2^31 (host check): 32-bit divisions
            const uint32_t v32 = (uint32_t)mm;
            const uint32_t t = v32 / (uint32_t)W, xx = v32 - t * (uint32_t)W;
            const uint32_t t2 = t / (uint32_t)H, yy = t - t2 * (uint32_t)H;
            const uint32_t b = t2 / (uint32_t)D, zz = t2 - b * (uint32_t)D;
            ov = (((int64_t)b * (2 * D) + 2 * zz) * (2 * H) + 2 * yy) * (int64_t)(2 * W) + 2 * xx;
        }
        s_ov[tid] = ov;
    }
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)kchunks * 2048;
    const int n0 = blockIdx.y * NT;

    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    // Operand staging.  Global reads want a warp to touch few 128-byte lines, the K-major image wants the 8 lanes of a
    // quarter warp on 8 different rows (16-byte bank groups): lane = (row % 8) + 8 * (chunk % 4), i.e. one instruction
    // reads 64 contiguous bytes of 8 rows and writes conflict-free.
    {
        const int64_t m_base = (int64_t)blockIdx.x * 128;
        const int kq = (kchunks + 3) >> 2;                      // groups of 4 chunks
        for (int idx = tid; idx < 16 * kq * 32; idx += 128) {   // 16 row groups x kq chunk groups x 32 lanes
            const int l = idx & 31, grp = idx >> 5;
            const int rg = grp % 16, cg = grp / 16;
            const int r = rg * 8 + (l & 7), kc = cg * 4 + (l >> 3);
            if (kc < kchunks) {
                const int64_t mm = m_base + r;
                uint8_t *dst = sA + (size_t)kc * 2048 + r * 16;
                if (mm < M) cp_async16(dst, reinterpret_cast<const uint4 *>(x + mm * xs) + kc);   // all copies in flight at once
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        const int ngr = NT >> 3;
        for (int idx = tid; idx < ngr * kq * 32; idx += 128) {
            const int l = idx & 31, grp = idx >> 5;
            const int rg = grp % ngr, cg = grp / ngr;
            const int r = rg * 8 + (l & 7), kc = cg * 4 + (l >> 3);
            if (kc < kchunks)
                cp_async16(sB + ((size_t)kc * NT + r) * 16, reinterpret_cast<const uint4 *>(wp + (int64_t)(n0 + r) * K) + kc);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
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
    // ---- epilogue: TMEM lane -> bf16 -> staging tile [128 rows][NT + 8] (aliases the operand images, which the MMAs
    // have finished reading) -> cooperative stores: 16-byte pieces of the Cout-channel run of each output voxel ----
    __nv_bfloat16 *sOut = reinterpret_cast<__nv_bfloat16 *>(smem);
    const int pitch = NT + 8;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < NT; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + c, r);
        tmem_wait_ld();
        uint4 lo, hi;
        lo.x = pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3]));
        lo.z = pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7]));
        hi.x = pack_bf16(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_bf16(__uint_as_float(r[10]), __uint_as_float(r[11]));
        hi.z = pack_bf16(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_bf16(__uint_as_float(r[14]), __uint_as_float(r[15]));
        uint4 *dst = reinterpret_cast<uint4 *>(sOut + (size_t)tid * pitch + c);
        dst[0] = lo;
        dst[1] = hi;
    }
    __syncthreads();
    {
        // a (row, position) pair owns `per` = gcd-free run of 16-byte pieces: positions inside this N tile are
        // n0 / Cout ... ; piece p of (row, pos) -> channels [8p, 8p + 8) of that output voxel
        const int per = Cout >> 3;                       // 16-byte pieces per output voxel
        const int pos_in_tile = NT / Cout > 0 ? NT / Cout : 1;
        const int H2 = 2 * H, W2 = 2 * W;
        if (NT % Cout == 0) {
            const int total = 128 * pos_in_tile * per;
            for (int i = tid; i < total; i += 128) {
                const int p = i % per;
                const int t2 = i / per;
                const int row = t2 % 128, pl = t2 / 128;     // consecutive lanes: pieces of one voxel, then the next input x
                if (s_ov[row] < 0) continue;
                const int pos = n0 / Cout + pl;
                const int dz = pos >> 2, dy = (pos >> 1) & 1, dx = pos & 1;
                const int64_t ov = s_ov[row] + ((int64_t)dz * H2 + dy) * W2 + dx;
                *reinterpret_cast<uint4 *>(y + ov * ys + p * 8) =
                    *reinterpret_cast<const uint4 *>(sOut + (size_t)row * pitch + pl * Cout + p * 8);
            }
        } else {   // the N tile covers part of one position's channels (Cout > NT)
            const int pos = n0 / Cout, co0 = n0 - pos * Cout;
            const int pieces = NT >> 3;
            const int dz = pos >> 2, dy = (pos >> 1) & 1, dx = pos & 1;
            for (int i = tid; i < 128 * pieces; i += 128) {
                const int p = i % pieces, row = i / pieces;
                if (s_ov[row] < 0) continue;
                const int64_t ov = s_ov[row] + ((int64_t)dz * H2 + dy) * W2 + dx;
                *reinterpret_cast<uint4 *>(y + ov * ys + co0 + p * 8) =
                    *reinterpret_cast<const uint4 *>(sOut + (size_t)row * pitch + p * 8);
            }
        }
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
    // N tile: the largest multiple of 16 (<= 128) that divides N and is a multiple or a divisor of Cout (an N tile then
    // covers whole output positions, or a slice of one)
    int NT = 0;
    for (int nt = 128; nt >= 16; nt -= 16)
        if (N % nt == 0 && (nt % Cout == 0 || Cout % nt == 0)) { NT = nt; break; }
    if (!NT) return WF_ERR_UNSUPPORTED;
    const int kchunks = Cin / 8;
    const size_t images = (size_t)kchunks * 2048 + (size_t)kchunks * NT * 16, stage = (size_t)128 * (NT + 8) * 2;
    const size_t smem = images > stage ? images : stage;
    if (smem > 200 * 1024) return WF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    uint32_t cols = 32;
    while ((int)cols < NT) cols <<= 1;
    const int64_t M = (int64_t)B * D * H * W;
    if (M >= 0x7fffffffLL) return WF_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)(N / NT));
    convT_k2s2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, (const uint16_t *)wpack,
                                                                (__nv_bfloat16 *)y, M, Cin, NT, Cout, D, H, W, x_vox_stride,
                                                                y_vox_stride, cols);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
