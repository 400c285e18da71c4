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

// One CTA = 128 input voxels x ALL 8 * Cout output columns: the A tile is staged once, the whole packed weight matrix is
// staged next to it (L2-resident), all N / NT accumulator tiles live in TMEM at once (N <= 512 columns), one commit covers
// every MMA, then the eight warps drain TMEM (warp w: lane quadrant w % 4, tiles of parity w / 4) through a bf16 staging
// area that aliases the operand images, and the block writes the output voxels with coalesced 16-byte stores.
template <bool F16>
__global__ void __launch_bounds__(256) convT_k2s2_kernel(const uint16_t *__restrict__ x, const uint16_t *__restrict__ wp,
                                                         uint16_t *__restrict__ y, int64_t M, int K, int NT, int ntiles,
                                                         int Cout, int D, int H, int W, int64_t xs, int64_t ys) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int64_t s_ov[128];   // output voxel index of position (0, 0, 0) for each input voxel of the tile; -1 = none
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kchunks = K >> 3;
    const int N = NT * ntiles;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)kchunks * 2048;   // [ntiles][kchunks][NT][16 bytes]
    if (tid < 128) {
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
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    // Operand staging with cp.async (every copy in flight at once).  Global reads want a warp to touch few 128-byte lines,
    // the K-major image wants the 8 lanes of a quarter warp on 8 different rows (16-byte bank groups):
    // lane = (row % 8) + 8 * (chunk % 4), i.e. one instruction reads 64 contiguous bytes of 8 rows, conflict-free writes.
    {
        const int64_t m_base = (int64_t)blockIdx.x * 128;
        const int kq = (kchunks + 3) >> 2;                      // groups of 4 chunks
        const int l = tid & 31, lr = l & 7, lc = l >> 3;        // lane -> (row within the 8-row group, chunk within the group)
        for (int cg = 0; cg < kq; ++cg) {
            const int kc = cg * 4 + lc;
            if (kc >= kchunks) continue;
            for (int rg = warp; rg < 16; rg += 8) {             // A: 16 row groups
                const int r = rg * 8 + lr;
                const int64_t mm = m_base + r;
                uint8_t *dst = sA + (size_t)kc * 2048 + r * 16;
                if (mm < M) cp_async16(dst, reinterpret_cast<const uint4 *>(x + mm * xs) + kc);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
            const int rg_per_tile = NT >> 3;
            for (int t = 0; t < ntiles; ++t)                    // B: N / 8 row groups, tile by tile
                for (int rg = warp; rg < rg_per_tile; rg += 8) {
                    const int r = rg * 8 + lr;
                    cp_async16(sB + (((size_t)t * kchunks + kc) * NT + r) * 16,
                               reinterpret_cast<const uint4 *>(wp + (int64_t)(t * NT + r) * K) + kc);
                }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if ((tid >> 5) == 0) {     // warp 0, converged: one elected lane issues (tc_common.cuh, "warp-uniform issue")
        const uint32_t idesc = instr_desc_h16<F16>(128, NT, false);
        const uint32_t al = smem_desc_lo(smem_u32(sA), 2048), bl = smem_desc_lo(smem_u32(sB), NT * 16), dh = smem_desc_hi(128);
        for (int t = 0; t < ntiles; ++t)
            for (int ks = 0; ks < (K >> 4); ++ks)
                mma_ss_w(tmem + t * NT, al + (uint32_t)(ks * 2 * 2048 / 16), dh, bl + (uint32_t)((t * kchunks + ks * 2) * NT), dh, idesc,
                         ks > 0 ? 1u : 0u);
        mma_commit_w(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    __syncthreads();   // every thread has seen the MMAs complete: the operand images may be overwritten
    // ---- TMEM -> bf16 staging [128 rows][N + 8] ----
    uint16_t *sOut = reinterpret_cast<uint16_t *>(smem);
    const int pitch = N + 8;
    {
        const int row = (warp & 3) * 32 + (tid & 31);
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int t = warp >> 2; t < ntiles; t += 2) {
            for (int c = 0; c < NT; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem + lane_base + t * NT + c, r);
                tmem_wait_ld();
                uint4 lo, hi;
                lo.x = pack_h16<F16>(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_h16<F16>(__uint_as_float(r[2]), __uint_as_float(r[3]));
                lo.z = pack_h16<F16>(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_h16<F16>(__uint_as_float(r[6]), __uint_as_float(r[7]));
                hi.x = pack_h16<F16>(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_h16<F16>(__uint_as_float(r[10]), __uint_as_float(r[11]));
                hi.z = pack_h16<F16>(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_h16<F16>(__uint_as_float(r[14]), __uint_as_float(r[15]));
                uint4 *dst = reinterpret_cast<uint4 *>(sOut + (size_t)row * pitch + t * NT + c);
                dst[0] = lo;
                dst[1] = hi;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    // ---- coalesced stores: piece p (16 bytes = 8 channels) of output position pos of input voxel `row`; a thread keeps
    // its (row, piece) pairs for all eight positions, so the index arithmetic is done once ----
    {
        const int per = Cout >> 3;
        const int64_t sz = (int64_t)(2 * H) * (2 * W), sy = 2 * W;
        for (int q = tid; q < 128 * per; q += 256) {   // consecutive lanes: pieces of one voxel, then the next input x
            const int row = q / per, p = q - row * per;
            const int64_t base = s_ov[row];
            if (base < 0) continue;
            const uint16_t *src = sOut + (size_t)row * pitch + p * 8;
            uint16_t *dst = y + base * ys + p * 8;
#pragma unroll
            for (int pos = 0; pos < 8; ++pos) {
                const int64_t off = (pos >> 2) * sz + ((pos >> 1) & 1) * sy + (pos & 1);
                *reinterpret_cast<uint4 *>(dst + off * ys) = *reinterpret_cast<const uint4 *>(src + pos * Cout);
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Persistent form (grid = number of SMs, tiles round-robin): the packed weight matrix is staged ONCE per CTA instead of once per
// 128-voxel tile (110 KB of L2 reads next to a 36 KB A tile made the one-shot kernel a serial 11.6 us per tile at one CTA per SM:
// 0.97 ms for a 1.63 GB problem); the A image of tile i + 1 is fetched with cp.async as soon as the MMAs of tile i have completed,
// i.e. underneath tile i's drain; accumulator tile t (NT columns = 8 / ntiles output positions) goes through a [128][NT + 8]
// staging area of its own (it cannot alias the resident weights any more) and out with the same coalesced 16-byte stores.
template <bool F16>
__global__ void __launch_bounds__(256, 1) convT_k2s2_persist_kernel(const uint16_t *__restrict__ x, const uint16_t *__restrict__ wp,
                                                                    uint16_t *__restrict__ y, int64_t M, int K, int NT, int ntiles,
                                                                    int Cout, int D, int H, int W, int64_t xs, int64_t ys,
                                                                    int64_t mtiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int64_t s_ov[128];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kchunks = K >> 3;
    const int N = NT * ntiles;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)kchunks * 2048;                 // [ntiles][kchunks][NT][16 bytes]
    uint16_t *sOut = reinterpret_cast<uint16_t *>(sB + (size_t)kchunks * N * 16);
    const int pitch = NT + 8;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    const int kq = (kchunks + 3) >> 2;
    const int l = tid & 31, lr = l & 7, lc = l >> 3;            // lane -> (row within the 8-row group, chunk within the group of 4)
    auto stage_a = [&](int64_t tile) {
        const int64_t m_base = tile * 128;
        for (int cg = 0; cg < kq; ++cg) {
            const int kc = cg * 4 + lc;
            if (kc >= kchunks) continue;
            for (int rg = warp; rg < 16; rg += 8) {
                const int r = rg * 8 + lr;
                const int64_t mm = m_base + r;
                uint8_t *dst = sA + (size_t)kc * 2048 + r * 16;
                if (mm < M) cp_async16(dst, reinterpret_cast<const uint4 *>(x + mm * xs) + kc);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    {   // resident weights
        const int rg_per_tile = NT >> 3;
        for (int cg = 0; cg < kq; ++cg) {
            const int kc = cg * 4 + lc;
            if (kc >= kchunks) continue;
            for (int t = 0; t < ntiles; ++t)
                for (int rg = warp; rg < rg_per_tile; rg += 8) {
                    const int r = rg * 8 + lr;
                    cp_async16(sB + (((size_t)t * kchunks + kc) * NT + r) * 16, reinterpret_cast<const uint4 *>(wp + (int64_t)(t * NT + r) * K) + kc);
                }
        }
    }
    if ((int64_t)blockIdx.x < mtiles) stage_a(blockIdx.x);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = instr_desc_h16<F16>(128, NT, false);
    const uint32_t al = smem_desc_lo(smem_u32(sA), 2048), bl = smem_desc_lo(smem_u32(sB), NT * 16), dh = smem_desc_hi(128);
    const int per = Cout >> 3;                  // 16-byte pieces per output voxel
    const int pos_per_tile = NT / Cout;         // output positions held by one accumulator tile
    const int64_t sz = (int64_t)(2 * H) * (2 * W), sy = 2 * W;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
        if (tid < 128) {
            const int64_t mm = tile * 128 + tid;
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
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();       // A image (and, the first time, the weights) complete; the previous tile's TMEM reads and stores are done
        tc_fence_after();
        if (warp == 0) {       // converged warp, one elected lane issues (tc_common.cuh, "warp-uniform issue")
            for (int t = 0; t < ntiles; ++t)
                for (int ks = 0; ks < (K >> 4); ++ks)
                    mma_ss_w(tmem + t * NT, al + (uint32_t)(ks * 2 * 2048 / 16), dh, bl + (uint32_t)((t * kchunks + ks * 2) * NT), dh, idesc,
                             ks > 0 ? 1u : 0u);
            mma_commit_w(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        // the A image is free: the next tile's copies run underneath this tile's drain
        const int64_t next = tile + gridDim.x;
        if (next < mtiles) stage_a(next);
        const int row = (warp & 3) * 32 + (tid & 31);
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int chalf = NT >> 1;               // warps 0-3 drain the first half of the tile's columns, warps 4-7 the second
        for (int t = 0; t < ntiles; ++t) {
            for (int c = (warp >> 2) * chalf; c < ((warp >> 2) + 1) * chalf; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem + lane_base + t * NT + c, r);
                tmem_wait_ld();
                uint4 lo, hi;
                lo.x = pack_h16<F16>(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_h16<F16>(__uint_as_float(r[2]), __uint_as_float(r[3]));
                lo.z = pack_h16<F16>(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_h16<F16>(__uint_as_float(r[6]), __uint_as_float(r[7]));
                hi.x = pack_h16<F16>(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_h16<F16>(__uint_as_float(r[10]), __uint_as_float(r[11]));
                hi.z = pack_h16<F16>(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_h16<F16>(__uint_as_float(r[14]), __uint_as_float(r[15]));
                uint4 *dst = reinterpret_cast<uint4 *>(sOut + (size_t)row * pitch + c);
                dst[0] = lo;
                dst[1] = hi;
            }
            tc_fence_before();
            __syncthreads();   // staging of tile t complete
            for (int q = tid; q < 128 * per; q += 256) {   // consecutive lanes: pieces of one voxel, then the next input x
                const int rr = q / per, p = q - rr * per;
                const int64_t base = s_ov[rr];
                if (base < 0) continue;
                const uint16_t *src = sOut + (size_t)rr * pitch + p * 8;
                uint16_t *dst = y + base * ys + p * 8;
                for (int pl = 0; pl < pos_per_tile; ++pl) {
                    const int pos = t * pos_per_tile + pl;
                    const int64_t off = (pos >> 2) * sz + ((pos >> 1) & 1) * sy + (pos & 1);
                    *reinterpret_cast<uint4 *>(dst + off * ys) = *reinterpret_cast<const uint4 *>(src + pl * Cout);
                }
            }
            __syncthreads();   // staging may be overwritten (next accumulator tile / next voxel tile)
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace wf

using namespace wf;

extern "C" int wf_convtranspose3d_k2s2_ndhwc(const void *x, const void *wpack, void *y, int dtype, int B, int D, int H,
                                             int W, int Cin, int Cout, int64_t x_vox_stride, int64_t y_vox_stride,
                                             void *stream) {
    if (!x || !wpack || !y) return WF_ERR_NULL_POINTER;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return WF_ERR_BAD_SHAPE;
    if (Cin % 16 || Cin > 512 || Cout % 16 || x_vox_stride < Cin || y_vox_stride < Cout || x_vox_stride % 8 || y_vox_stride % 8)
        return WF_ERR_BAD_SHAPE;
    if (!aligned16(x) || !aligned16(wpack) || !aligned16(y)) return WF_ERR_MISALIGNED;
    const int N = 8 * Cout;
    if (N > 512) return WF_ERR_UNSUPPORTED;                     // all accumulator tiles live in TMEM at once
    // N tile: the largest multiple of 16 (<= 256) that divides N
    int NT = 0;
    for (int nt = 256; nt >= 16; nt -= 16)
        if (N % nt == 0) { NT = nt; break; }
    if (!NT) return WF_ERR_UNSUPPORTED;
    const int kchunks = Cin / 8;
    const size_t images = (size_t)kchunks * 2048 + (size_t)kchunks * N * 16, stage = (size_t)128 * (N + 8) * 2;
    const size_t smem = images > stage ? images : stage;
    if (smem > 220 * 1024) return WF_ERR_UNSUPPORTED;
    static unsigned long long attr_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attr_done)) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    const int64_t M = (int64_t)B * D * H * W;
    if (M >= 0x7fffffffLL) return WF_ERR_UNSUPPORTED;
    // persistent kernel: weights resident, needs a staging area of its own and whole output positions per accumulator tile
    const size_t smem_p = images + (size_t)128 * (NT + 8) * 2;
    const int64_t mtiles = (M + 127) / 128;
    if (!ab_old() && smem_p <= 220 * 1024 && NT % Cout == 0 && (NT / 2) % 16 == 0 && mtiles > kNumSMs) {
        static unsigned long long attr_p = 0;
        if (first_use_on_current_device(attr_p)) {
            WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            WF_CUDA_CHECK(cudaFuncSetAttribute(convT_k2s2_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        }
        if (dtype == WF_F16)
            convT_k2s2_persist_kernel<true><<<kNumSMs, 256, smem_p, (cudaStream_t)stream>>>(
                (const uint16_t *)x, (const uint16_t *)wpack, (uint16_t *)y, M, Cin, NT, N / NT, Cout, D, H, W, x_vox_stride, y_vox_stride, mtiles);
        else
            convT_k2s2_persist_kernel<false><<<kNumSMs, 256, smem_p, (cudaStream_t)stream>>>(
                (const uint16_t *)x, (const uint16_t *)wpack, (uint16_t *)y, M, Cin, NT, N / NT, Cout, D, H, W, x_vox_stride, y_vox_stride, mtiles);
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    const unsigned grid = (unsigned)((M + 127) / 128);
    if (dtype == WF_F16)
        convT_k2s2_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(
            (const uint16_t *)x, (const uint16_t *)wpack, (uint16_t *)y, M, Cin, NT, N / NT, Cout, D, H, W, x_vox_stride, y_vox_stride);
    else
        convT_k2s2_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(
            (const uint16_t *)x, (const uint16_t *)wpack, (uint16_t *)y, M, Cin, NT, N / NT, Cout, D, H, W, x_vox_stride, y_vox_stride);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
