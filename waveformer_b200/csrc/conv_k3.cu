// 3x3x3 convolution (padding 1, stride 1, no bias), 48 -> 48 channels, channels-last bf16, rows of W = 128 voxels:
// a producer / consumer tcgen05 implicit GEMM with the InstanceNorm statistics of the result fused into the epilogue and,
// optionally, the InstanceNorm + LeakyReLU of the INPUT applied on the fly to the staged rows.
//
// Reference: conv2 of the two 128^3 residual blocks, Waveformer.encoder1 / decoder1.conv_block (MONAI UnetResBlock.forward,
// monai/networks/blocks/dynunet_block.py:98-111: `out = self.conv2(self.lrelu(self.norm1(self.conv1(inp))))`, then norm2).
// The library runs this convolution at 0.77 ms per batch-2 window and needs one more pass for norm1 + lrelu (0.18 ms) and
// one for norm2's statistics (0.08 ms).
//
// Work decomposition.  One CTA owns a block of 4 output rows (y0 .. y0+3) of one (b, z) plane; a row is 128 voxels = the
// M dimension (one TMEM lane per voxel), the 48 output channels are N, and K runs over 27 taps x 48 input channels.
//   * An INPUT row (z', y') is staged once as a K-major no-swizzle image [6 chunks][132 rows][16 B]: row r holds voxel
//     x = r - 1, rows 0 and 129 are the zero halo.  Because rows are 16 bytes apart, the operand for tap dx is the same
//     image with its start address advanced by dx rows - no im2col copy.  The row feeds up to 3 output rows (dy); their
//     accumulators are adjacent in TMEM and their weight tiles adjacent along N, so each (dx, k-step) is ONE
//     tcgen05.mma 128 x 144 x 16: 9 instructions per input row instead of 27 (the A operand streams at the same ~100 clk
//     per instruction whatever N is).
//   * The 18 input rows of a block (3 planes x 6 rows) stream through a 4-slot ring (cp.async -> mbarrier), out-of-volume
//     rows are skipped (zero contribution).  All 27 weight tiles [144 x 16] (124 KB) stay resident in shared memory.
//   * Accumulators: 4 output rows x 48 columns, double buffered in TMEM (2 x 192 columns), so the epilogue of a block
//     (TMEM -> bf16 -> staging -> coalesced stores + per-channel sum / sum of squares) overlaps the MMAs of the next one.
// Warp roles (288 threads): warps 0-3 loaders (+ optional input normalisation), warp 4 MMA issuer, warps 5-8 epilogue.
#include <type_traits>

#include <cuda.h>

#include "tc_common.cuh"
#include "wf_common.cuh"

namespace wf {

using namespace tc;

constexpr int kK3C = 48;                       // channels in = channels out
constexpr int kK3Chunks = kK3C / 8;            // 6 sixteen-byte K chunks per voxel
constexpr int kK3Rows = 132;                   // image rows: 1 halo + 128 voxels + 1 halo + 2 pad
constexpr int kK3RowImg = kK3Chunks * kK3Rows * 16;        // 12672 bytes per staged input row
#ifndef WF_K3_RING
#define WF_K3_RING 7
#endif
#ifndef WF_K3_LAG
#define WF_K3_LAG 3
#endif
#ifndef WF_K3_SR
#define WF_K3_SR 1
#endif
constexpr int kK3Ring = WF_K3_RING;           // staged input rows (measured: 7 slots / 1-row epilogue staging 645 us, 5 / 2 rows 686 us)
constexpr int kK3Lag = WF_K3_LAG;                       // rows of cp.async copies in flight per loader thread
constexpr int kK3WTile3 = 2 * 3 * kK3C * 16;               // one [144 = 3 dy x 48 out][16] weight tile: 4608 bytes
constexpr int kK3WBytes = 27 * kK3WTile3;                  // 3 dz x 3 dx x 3 k-steps tiles = 124416
constexpr int kK3StageRows = WF_K3_SR;                      // output rows per epilogue staging pass
constexpr int kK3StageBytes = kK3StageRows * 128 * (kK3C + 8) * 2;   // 14336 per row
constexpr int kK3Smem = kK3WBytes + kK3Ring * kK3RowImg + kK3StageBytes;   // 227456 with 7 slots

__device__ __forceinline__ void cp_async16_k3(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mbar_arrive_k3(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct K3Args {
    const uint16_t *x;          // [B, D, H, W = 128, 48] bf16 or fp16, voxel stride xs
    const uint16_t *wpack;      // [3 dz][3 dx][3 ks][2 chunks][144 = 3 dy x 48 out][8] in the same 16-bit format
    uint16_t *y;                // [B, D, H, 128, 48], voxel stride ys
    double *sums;               // [B][48][2] (sum, sum of squares of the rounded outputs); zeroed by the host wrapper
    const float *in_mr;         // optional (mean, rstd) [B][48][2] of the input: x is normalised + LeakyReLU'd while staged
    float slope;
    int64_t xs, ys;
    int B, D, H;
    long long *prof;            // PROF instantiation only: [grid][8] stage clocks (see wf_conv3d_k3_c48_stage_clocks)
};

// F16: activations / weights / result are fp16 instead of bf16 (same tensor-core rate, 10-bit mantissa)
template <bool F16, bool PROF = false>
__global__ void __launch_bounds__(288, 1) conv3d_k3_c48_kernel(K3Args a) {
    using T16 = typename std::conditional<F16, __half, __nv_bfloat16>::type;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[kK3Ring], bar_empty[kK3Ring], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t *sW = smem;
    uint8_t *sRing = smem + kK3WBytes;
    uint16_t *sStage = reinterpret_cast<uint16_t *>(smem + kK3WBytes + kK3Ring * kK3RowImg);
    const int tid = threadIdx.x, warp = (int)warp_idx_uniform(), lane = tid & 31;
    constexpr int W = 128;
    const int yblocks = (a.H + 3) >> 2;
    const int64_t nblocks = (int64_t)a.B * a.D * yblocks;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < kK3Ring; ++i) {
            mbar_init(&bar_full[i], 128);    // one deferred arrival per loader thread
            mbar_init(&bar_empty[i], 1);     // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_acc_full[i], 1);  // tcgen05.commit
            mbar_init(&bar_acc_empty[i], 4); // one per epilogue warp
        }
        mbar_fence_init();
    }
    // resident weights (every thread helps) and the zero halo rows of the ring images
    for (int i = tid; i < kK3WBytes / 16; i += 288)
        reinterpret_cast<uint4 *>(sW)[i] = __ldg(reinterpret_cast<const uint4 *>(a.wpack) + i);
    for (int i = tid; i < kK3Ring * kK3Chunks * 4; i += 288) {
        const int slot = i / (kK3Chunks * 4), r = i % (kK3Chunks * 4);
        const int ch = r >> 2, which = r & 3;                       // rows 0, 129, 130, 131 of every chunk
        const int row = which == 0 ? 0 : 128 + which;
        *reinterpret_cast<uint4 *>(sRing + slot * kK3RowImg + (ch * kK3Rows + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // PROF: clocks this thread spent inside a wait, accumulated per cause (a stage that never waits is the pipeline's limiter)
    long long pw[3] = {0, 0, 0};
    const long long p_t0 = PROF ? clock64() : 0;
#define K3_TIMED(idx, ...) do { if constexpr (PROF) { const long long t_ = clock64(); __VA_ARGS__; pw[idx] += clock64() - t_; } else { __VA_ARGS__; } } while (0)

    if (warp < 4) {
        // ================================================= loaders ====================================================
        // thread t stages voxel x = t of each input row: 6 chunks of 16 bytes, written to rows t + 1 of the chunk planes
        uint32_t n_loaded = 0;   // rows staged so far by this CTA (ring position)
        uint32_t n_arrived = 0;  // rows whose copies have landed and been announced (cp.async path lags by <= 2 rows)
        int64_t pend_b[kK3Lag + 1];   // batch element of each in-flight row (selects the normalisation constants)
        float sc[kK3C], sh[kK3C];     // input normalisation as x * sc + sh (only used with in_mr)
        int64_t norm_b = -1;
        auto announce_upto = [&](uint32_t upto) {
            // the copies of rows [n_arrived, upto) have completed (cp.async.wait_group).  Optionally normalise + LeakyReLU
            // this thread's own voxel (the 6 chunks it copied itself: no cross-thread dependency), make the row visible to
            // the tensor core's async proxy, then arrive on the slot's barrier.
            for (; n_arrived < upto; ++n_arrived) {
                if (a.in_mr != nullptr) {
                    uint8_t *img = sRing + (n_arrived % kK3Ring) * kK3RowImg;
                    const int64_t rb = pend_b[n_arrived % (kK3Lag + 1)];
                    if (rb != norm_b) {      // (scale, shift) of the 48 channels live in registers; reloaded per batch element
                        const float2 *mr = reinterpret_cast<const float2 *>(a.in_mr) + rb * kK3C;
#pragma unroll
                        for (int c = 0; c < kK3C; ++c) {
                            const float2 m = __ldg(mr + c);
                            sc[c] = m.y;
                            sh[c] = -m.x * m.y;
                        }
                        norm_b = rb;
                    }
#pragma unroll
                    for (int ch = 0; ch < kK3Chunks; ++ch) {
                        uint4 *cell = reinterpret_cast<uint4 *>(img + (ch * kK3Rows + tid + 1) * 16);
                        float f[8];
                        Pack<T16>::unpack(*cell, f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float t = fmaf(f[e], sc[ch * 8 + e], sh[ch * 8 + e]);
                            f[e] = fmaxf(t, t * a.slope);
                        }
                        *cell = Pack<T16>::pack(f);
                    }
                }
                fence_proxy_async();
                mbar_arrive_k3(&bar_full[n_arrived % kK3Ring]);
            }
        };
        for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const int yb = (int)(blk % yblocks);
            const int64_t t2 = blk / yblocks;
            const int z = (int)(t2 % a.D);
            const int64_t b = t2 / a.D;
            const int y0 = yb * 4;
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = z + dz;
                if ((unsigned)zz >= (unsigned)a.D) continue;
                for (int iy = -1; iy <= 4; ++iy) {
                    const int yy = y0 + iy;
                    if ((unsigned)yy >= (unsigned)a.H) continue;
                    const int slot = n_loaded % kK3Ring;
                    const uint32_t use = n_loaded / kK3Ring;
                    if (use > 0) K3_TIMED(0, mbar_wait(&bar_empty[slot], (use - 1) & 1));   // the MMAs that read this slot are done
                    uint8_t *img = sRing + slot * kK3RowImg;
                    const uint16_t *src = a.x + ((((int64_t)b * a.D + zz) * a.H + yy) * W + tid) * a.xs;
#pragma unroll
                    for (int ch = 0; ch < kK3Chunks; ++ch)
                        cp_async16_k3(img + (ch * kK3Rows + tid + 1) * 16, src + ch * 8);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    pend_b[n_loaded % (kK3Lag + 1)] = b;
                    if (n_loaded + 1 - n_arrived > kK3Lag) {      // keep at most kK3Lag rows in flight per thread
                        K3_TIMED(1, asm volatile("cp.async.wait_group %0;" ::"n"(kK3Lag) : "memory"));
                        announce_upto(n_loaded + 1 - kK3Lag);
                    }
                    ++n_loaded;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        announce_upto(n_loaded);
        if constexpr (PROF)
            if (tid == 0) { a.prof[blockIdx.x * 8 + 0] = pw[0]; a.prof[blockIdx.x * 8 + 1] = pw[1]; a.prof[blockIdx.x * 8 + 2] = clock64() - p_t0; }
    } else if (warp == 4) {
        // ================================================= issuer =====================================================
        // All 32 lanes run this loop with warp-uniform values; one elected lane issues each tcgen05 instruction (tc_common.cuh:
        // from inside an `if (lane == 0)` block every MMA cost ~20 dependent instructions of R2UR moves = 140 clk, twice the
        // tensor pipe's own interval).
        {
            const uint32_t idesc1 = instr_desc_h16<F16>(128, kK3C, false), idesc2 = instr_desc_h16<F16>(128, 2 * kK3C, false),
                           idesc3 = instr_desc_h16<F16>(128, 3 * kK3C, false);
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(sRing), kK3Rows * 16), a_hi = smem_desc_hi(128);   // slot 0, chunk 0, row 0
            const uint32_t w_lo0 = smem_desc_lo(smem_u32(sW), 3 * kK3C * 16), w_hi = smem_desc_hi(128);      // tile (dz, dx, ks) = [2][144 = dy x out][8]
            uint32_t n_used = 0, nblk = 0;
            for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++nblk) {
                const int yb = (int)(blk % yblocks);
                const int z = (int)((blk / yblocks) % a.D);
                const int y0 = yb * 4;
                const int buf = nblk & 1;
                // the epilogue has zeroed (first use) or drained and re-zeroed (later uses) this accumulator buffer
                K3_TIMED(0, mbar_wait_warp(&bar_acc_empty[buf], (nblk >> 1) & 1));
                tc_fence_after();
                for (int dz = -1; dz <= 1; ++dz) {
                    if ((unsigned)(z + dz) >= (unsigned)a.D) continue;
                    for (int iy = -1; iy <= 4; ++iy) {
                        const int yy = y0 + iy;
                        if ((unsigned)yy >= (unsigned)a.H) continue;
                        const int slot = n_used % kK3Ring;
                        K3_TIMED(1, mbar_wait_warp(&bar_full[slot], (n_used / kK3Ring) & 1));
                        tc_fence_after();
                        // descriptors: a base per ring slot / for the weights, plus compile-time offsets in 16-byte units
                        // added to the low word (the address field holds addr >> 4; every shared-memory address fits its 14 bits)
                        const uint32_t a_lo = a_lo0 + (uint32_t)(slot * (kK3RowImg / 16));
                        // One MMA per (dx, k-step) covers EVERY output row this input row feeds: the rows' accumulators are
                        // adjacent in TMEM (descending row order) and their dy weight tiles are adjacent along N, so
                        // N = 48 x (rows fed) = up to 144.  Every MMA accumulates: the epilogue warps zero an accumulator
                        // buffer before handing it back.
                        const int r_hi = min(3, min(iy + 1, a.H - 1 - y0)), r_lo = max(0, iy - 1);
                        if (r_hi >= r_lo) {
                            const int nrows = r_hi - r_lo + 1;
                            const uint32_t idesc = nrows == 3 ? idesc3 : (nrows == 2 ? idesc2 : idesc1);
                            const uint32_t acc = tmem + buf * 256 + (3 - r_hi) * kK3C;
                            // first weight row: dy of r_hi = iy - r_hi (in -1..1) -> N offset (dy + 1) * 48
                            const uint32_t w_lo = w_lo0 + (uint32_t)((dz + 1) * 9 * (kK3WTile3 / 16) + (iy - r_hi + 1) * kK3C);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int ks = 0; ks < 3; ++ks)
                                    mma_ss_w(acc, a_lo + (uint32_t)(ks * 2 * kK3Rows + dx), a_hi,
                                             w_lo + (uint32_t)((dx * 3 + ks) * (kK3WTile3 / 16)), w_hi, idesc, 1u);
                        }
                        mma_commit_w(&bar_empty[slot]);   // slot reusable once these MMAs have read it
                        ++n_used;
                    }
                }
                mma_commit_w(&bar_acc_full[buf]);
            }
            if constexpr (PROF)
                if (lane == 0) { a.prof[blockIdx.x * 8 + 3] = pw[0]; a.prof[blockIdx.x * 8 + 4] = pw[1]; a.prof[blockIdx.x * 8 + 5] = clock64() - p_t0; }
        }
    } else {
        // ================================================ epilogue ====================================================
        const int quad = warp & 3;                          // a warp may touch TMEM lanes 32 * (warp id % 4) .. + 31 only
        const int et = quad * 32 + lane;                    // 0..127 = voxel x of this thread's TMEM lane (warps 5..8 -> 1,2,3,0)
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        constexpr int pitch = kK3C + 8;
        double acc_s[2] = {0.0, 0.0}, acc_q[2] = {0.0, 0.0};   // threads et < 96: channel pair et % 24, row quarter et / 24
        int64_t acc_b = -1;
        const int cpair = et % 24, rq = et / 24;
        auto flush = [&]() {
            if (et < 96 && acc_b >= 0) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double *dst = a.sums + (acc_b * kK3C + 2 * cpair + e) * 2;
                    atomicAdd(dst, acc_s[e]);
                    atomicAdd(dst + 1, acc_q[e]);
                }
            }
            acc_s[0] = acc_s[1] = acc_q[0] = acc_q[1] = 0.0;
        };
        auto zero_and_release = [&](int buf) {
            // every MMA accumulates, so an accumulator buffer is handed to the issuer zeroed
            uint32_t zeros[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zeros[i] = 0u;
#pragma unroll
            for (int c = 0; c < 4 * kK3C; c += 16) tmem_st16(tmem + lane_base + buf * 256 + c, zeros);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_k3(&bar_acc_empty[buf]);
        };
        zero_and_release(0);
        zero_and_release(1);
        uint32_t nblk = 0;
        for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++nblk) {
            const int yb = (int)(blk % yblocks);
            const int64_t t2 = blk / yblocks;
            const int z = (int)(t2 % a.D);
            const int64_t b = t2 / a.D;
            const int y0 = yb * 4;
            const int buf = nblk & 1;
            if (b != acc_b) { flush(); acc_b = b; }
            K3_TIMED(0, mbar_wait(&bar_acc_full[buf], (nblk >> 1) & 1));
            tc_fence_after();
            for (int half = 0; half < 4 / kK3StageRows; ++half) {   // kK3StageRows output rows at a time through the staging tile
                const int rows_here = min(kK3StageRows, a.H - (y0 + kK3StageRows * half));
                if (rows_here <= 0) break;
                for (int rr = 0; rr < rows_here; ++rr) {
                    const uint32_t acc = tmem + lane_base + buf * 256 + (3 - (kK3StageRows * half + rr)) * kK3C;   // descending row order
#pragma unroll
                    for (int c = 0; c < kK3C; c += 16) {
                        uint32_t r[16];
                        tmem_ld16(acc + c, r);
                        tmem_wait_ld();
                        uint4 lo, hi;
                        lo.x = pack_h16<F16>(__uint_as_float(r[0]), __uint_as_float(r[1]));   lo.y = pack_h16<F16>(__uint_as_float(r[2]), __uint_as_float(r[3]));
                        lo.z = pack_h16<F16>(__uint_as_float(r[4]), __uint_as_float(r[5]));   lo.w = pack_h16<F16>(__uint_as_float(r[6]), __uint_as_float(r[7]));
                        hi.x = pack_h16<F16>(__uint_as_float(r[8]), __uint_as_float(r[9]));   hi.y = pack_h16<F16>(__uint_as_float(r[10]), __uint_as_float(r[11]));
                        hi.z = pack_h16<F16>(__uint_as_float(r[12]), __uint_as_float(r[13])); hi.w = pack_h16<F16>(__uint_as_float(r[14]), __uint_as_float(r[15]));
                        uint4 *dst = reinterpret_cast<uint4 *>(sStage + (size_t)(rr * 128 + et) * pitch + c);
                        dst[0] = lo;
                        dst[1] = hi;
                    }
                }
                if (half == 4 / kK3StageRows - 1 || a.H - (y0 + kK3StageRows * (half + 1)) <= 0)
                    zero_and_release(buf);                          // accumulators fully read
                asm volatile("bar.sync 2, 128;" ::: "memory");      // staging complete (epilogue warps only)
                // coalesced stores: the rows of one pass are contiguous voxels in global memory
                const int nvox = rows_here * 128;
                const int64_t v0 = (((int64_t)b * a.D + z) * a.H + y0 + kK3StageRows * half) * W;
                for (int i = et; i < nvox * kK3Chunks; i += 128) {
                    const int vx = i / kK3Chunks, p = i - vx * kK3Chunks;
                    *reinterpret_cast<uint4 *>(a.y + (v0 + vx) * a.ys + p * 8) =
                        *reinterpret_cast<const uint4 *>(sStage + (size_t)vx * pitch + p * 8);
                }
                // statistics of the rounded outputs: 24 channel pairs x 4 row quarters
                if (et < 96) {
                    const uint32_t *col = reinterpret_cast<const uint32_t *>(sStage) + cpair;
                    constexpr int wpitch = pitch / 2;
                    const int q0 = rq * (nvox / 4), q1 = q0 + nvox / 4;
                    float s0 = 0.f, s1 = 0.f, qq0 = 0.f, qq1 = 0.f;
#pragma unroll 8
                    for (int r = q0; r < q1; ++r) {
                        const uint32_t w = col[(size_t)r * wpitch];
                        stat_h16x2<F16>(w, s0, s1, qq0, qq1);
                    }
                    acc_s[0] += (double)s0; acc_s[1] += (double)s1; acc_q[0] += (double)qq0; acc_q[1] += (double)qq1;
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");      // staging consumed before the next half overwrites it
            }
        }
        flush();
        if constexpr (PROF)
            if (et == 0) { a.prof[blockIdx.x * 8 + 6] = pw[0]; a.prof[blockIdx.x * 8 + 7] = clock64() - p_t0; }
    }
#undef K3_TIMED
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------------------
// Rolling-row variant (the default when the input needs no normalisation).  What the stage clocks of the block kernel above
// showed (scripts/k3_stage_clocks.py, profiles/r02_k3_stage_clocks.txt): every role was busy - the SM's load / store pipe was the
// shared limiter (cp.async copies of 16 bytes per thread, a fence + mbarrier arrival per thread and row, ld.shared + st.global of
// the staged result, 32 ld.shared per thread and row for the statistics), and a third of the MMAs were edge rows of a 4-row block
// with N = 48 / 96, which occupy the tensor pipe as long as N = 144 does (scripts/mma_probe.cu: max(66, N / 2) clk per MMA, whatever
// the operand layout).  Here:
//   * a CTA owns a contiguous RUN of output rows (B * D * H rows split evenly over the grid) and walks it in y.  The accumulators
//     of the output rows in flight form a ring of 10 TMEM slots (48 columns each); the input row (z + dz, y') feeds the output
//     rows y' - 1, y', y' + 1, whose slots are adjacent, so EVERY MMA is 128 x 144 x 16 (two smaller ones where the ring wraps or a
//     run starts / ends): 27 MMAs and 3 staged input rows per output row instead of 40.5 and 4.5;
//   * an input row is ONE cp.async.bulk.tensor (TMA) instruction issued by one thread: box [64 channels x 130 voxels] at x = -1
//     of a tensor map [48 channels][128 x][rows]: the two halo rows and channels 48..63 are the out-of-bounds zero fill, and the
//     image lands 128-byte swizzled, K-major, 128 bytes per voxel - the operand of tap dx / k-step ks is the same image with its
//     descriptor start advanced by dx * 128 + ks * 32 bytes (scripts/tma_probe.cu checks exactly this);
//   * an output row is complete one step after its own y; four epilogue warps drain it (TMEM -> 16 bit, per-thread running
//     sum / sum of squares of its 48 channels in registers, dense staging tile, ONE bulk store of the 12 KB row) while the tensor
//     pipe is rows ahead, zero the slot and hand it back.
// `addend` (optional): a second 16-bit tensor laid out like y that is added to the accumulators before rounding - the second
// pass of a convolution whose input channels are split over two launches (decoder1.conv1: 96 = 48 + 48 input channels).
constexpr int kRollSlots = 10;                                 // accumulator ring: output rows in flight
// A staged input row = two TMA boxes: channels 0..31 as a 64-byte-swizzled image [130 voxels][64 B] (k-steps 0 and 1) and channels
// 32..47 as a 32-byte-swizzled image [130][32 B] (k-step 2).  One 128-byte-swizzled image [130][128 B] with channels 48..63 zero
// filled works the same (first version) but a quarter of it is padding: 4 ring slots instead of 6 beside the 124 KB of weights.
constexpr int kRollImgA = 130 * 64, kRollImgB = 130 * 32;      // bytes the two boxes deliver
constexpr int kRollImgBytes = kRollImgA + kRollImgB;
constexpr int kRollOffB = 17 * 512;                            // 8704: the 32-channel image padded to whole 8-row swizzle atoms (512 B)
constexpr int kRollImg = 26 * 512;                             // slot pitch 13312: keeps both images aligned to their atoms (512 / 256 B)
constexpr int kRollStage = 128 * kK3C * 2;                     // one dense output row: 12288 bytes
constexpr int roll_smem(int ring, int nstage) { return ring * kRollImg + kK3WBytes + nstage * kRollStage; }   // 6 slots, 2 tiles: 228864
constexpr int kRollThreads = 192;                              // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

struct K3RollArgs {
    const uint16_t *wpack;      // as K3Args
    uint16_t *y;                // [B, D, H, 128, 48 of ys]
    const uint16_t *addend;     // optional, voxel stride as_
    double *sums;
    int64_t ys, as_;
    int B, D, H;
    long long *prof;
};

__device__ __forceinline__ void tma_load_row(void *dst, const CUtensorMap *map, int c0, int row, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(-1), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}

// kRollRing staged input rows, kRollNStage output staging tiles
template <bool F16, int kRollRing, int kRollNStage, bool PROF = false>
__global__ void __launch_bounds__(kRollThreads, 1) conv3d_k3_c48_roll_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap_hi,
                                                                             K3RollArgs a) {
    unsigned long long gt_entry = 0;
    if constexpr (PROF) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[kRollRing], bar_empty[kRollRing], bar_row_full[kRollSlots], bar_row_empty[kRollSlots];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_red[4][2 * kK3C];
    uint8_t *sRing = smem;
    uint8_t *sW = smem + kRollRing * kRollImg;
    uint8_t *sStage = sW + kK3WBytes;
    const int tid = threadIdx.x, warp = (int)warp_idx_uniform(), lane = tid & 31;
    constexpr int W = 128;
    // This CTA's output rows, in (b, z, y) order, come in pieces: first `whole` entire planes (plane j * grid + cta for j < whole), walked
    // by all CTAs in lockstep - neighbouring CTAs work on neighbouring planes at the same y, so the input rows they share (each row is
    // needed by the planes z - 1, z, z + 1) are fetched from DRAM once and hit in L2 twice - then an even share of the rows of the
    // remaining planes.  (One contiguous run of rows per CTA - 1.73 planes at BASELINE size - leaves the CTAs of neighbouring planes 93
    // rows out of phase: ncu showed the input being read three times from DRAM.)
    const int64_t planes = (int64_t)a.B * a.D;
    const int64_t whole = planes / gridDim.x;
    const int64_t rem_base = whole * gridDim.x * a.H, rem_rows = (planes - whole * gridDim.x) * a.H;
#define ROLL_PIECES_BEGIN                                                                                                             \
    for (int64_t piece_ = 0; piece_ <= whole; ++piece_) {                                                                              \
        const int64_t r0 = piece_ < whole ? (piece_ * gridDim.x + blockIdx.x) * a.H : rem_base + rem_rows * blockIdx.x / gridDim.x;    \
        const int64_t r1 = piece_ < whole ? r0 + a.H : rem_base + rem_rows * (blockIdx.x + 1) / gridDim.x;
#define ROLL_PIECES_END }

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < kRollRing; ++i) {
            mbar_init(&bar_full[i], 1);          // the producer's arrive.expect_tx; the TMA completes the transaction bytes
            mbar_init(&bar_empty[i], 1);         // tcgen05.commit
        }
        for (int i = 0; i < kRollSlots; ++i) {
            mbar_init(&bar_row_full[i], 1);      // tcgen05.commit
            mbar_init(&bar_row_empty[i], 4);     // one per epilogue warp
        }
        mbar_fence_init();
    }
    for (int i = tid; i < kK3WBytes / 16; i += kRollThreads)
        reinterpret_cast<uint4 *>(sW)[i] = __ldg(reinterpret_cast<const uint4 *>(a.wpack) + i);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long pw[7] = {0, 0, 0, 0, 0, 0, 0};
    const long long p_t0 = PROF ? clock64() : 0;
#define K3_TIMED(idx, ...) do { if constexpr (PROF) { const long long t_ = clock64(); __VA_ARGS__; pw[idx] += clock64() - t_; } else { __VA_ARGS__; } } while (0)

    if (warp == 0) {
        // ================================================= producer ===================================================
        int slot = 0;
        uint32_t ph = 1;                 // parity of the "slot is free" phase to wait for; the first pass over the ring is free
        bool first_pass = true;
        ROLL_PIECES_BEGIN
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int yb = ya + n - 1;
            const int z = (int)(p % a.D);
            // Every step loads three rows, whatever the position: a row outside the volume (z + dz or yy out of range) is fetched
            // at row coordinate -1, i.e. entirely from the tensor map's out-of-bounds zero fill, and contributes nothing.  The
            // pipeline therefore has ONE shape (3 rows and 27 MMAs per step) from the first step to the last.
            for (int yy = ya - 1; yy <= yb + 1; ++yy)
                for (int dz = -1; dz <= 1; ++dz) {
                    const bool inside = (unsigned)(z + dz) < (unsigned)a.D && (unsigned)yy < (unsigned)a.H;
                    if (!first_pass) K3_TIMED(0, mbar_wait_warp_relaxed(&bar_empty[slot], ph));   // the MMAs that read this slot are done
                    if (elect_one()) {
                        mbar_expect_tx(&bar_full[slot], kRollImgBytes);
                        const int row = inside ? (int)((p + dz) * a.H + yy) : -1;
                        tma_load_row(sRing + slot * kRollImg, &xmap, 0, row, &bar_full[slot]);
                        tma_load_row(sRing + slot * kRollImg + kRollOffB, &xmap_hi, 32, row, &bar_full[slot]);
                    }
                    __syncwarp();
                    if (++slot == kRollRing) { slot = 0; ph ^= 1; first_pass = false; }
                }
            r += n;
        }
        ROLL_PIECES_END
        if constexpr (PROF)
            if (lane == 0) {
                for (int q = 0; q < 6; ++q) a.prof[(blockIdx.x * 3 + 0) * 8 + q] = pw[q];
                a.prof[(blockIdx.x * 3 + 0) * 8 + 6] = (long long)gt_entry;                 // wall clock (ns) at kernel entry
                a.prof[(blockIdx.x * 3 + 0) * 8 + 7] = clock64() - p_t0;
            }
    } else if (warp == 1) {
        // ================================================= issuer =====================================================
        // The issue loop is kept SMALL on purpose (one copy of the nine tcgen05.mma of a (dz, part), runtime loops around it): with
        // the three dz taps and the ring-wrap split unrolled the kernel had 102 MMA sites in 200 KB of code, and CTAs fell for
        // hundreds of steps into a mode of ~550 clk per MMA (profiles/r02_k3_roll_notes.txt) - the issuing warp starved for
        // instructions - while scripts/mma_probe.cu runs the same stream at ~90 clk per MMA.  The MMA queue is shallow: whatever the
        // issuing warp does between two MMAs (barrier polls, address arithmetic) is exposed as tensor-pipe idle time.
        const uint32_t idesc0 = instr_desc_h16<F16>(128, 0, false);                  // N field (bits 17..22, N >> 3) added per call
        // A, k-steps 0 / 1: 64-byte swizzle, K-major, 8-row groups 512 bytes apart (layout type 4 in bits 61..63; the leading-dimension
        // field is unused); k-step 2: 32-byte swizzle, groups 256 bytes apart (type 6).  Tap dx = start advanced by dx rows, k-step 1 = by
        // 32 bytes inside the row.  Weights: no swizzle, tile (dz, dx, ks) = [2 chunks][144 = dy x out][8]
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(sRing), 16), a_hi = smem_desc_hi(512) | (4u << 29), a_hi2 = smem_desc_hi(256) | (6u << 29);
        const uint32_t w_lo0 = smem_desc_lo(smem_u32(sW), 3 * kK3C * 16), w_hi = smem_desc_hi(128);
        // ring positions are kept as (index, phase) pairs advanced by hand: no 64-bit division in the issue loop
        int ring_i = 0;                  // staged-row slot to consume next
        uint32_t ring_ph = 0;
        int c0m = 0;                     // (output-row counter of the segment's first row) mod kRollSlots
        int ready_m = 0, ready_rows = 0; // counter of the next row whose slot must be confirmed free: mod kRollSlots / absolute
        uint32_t ready_ph = 0;
        int c0 = 0;                      // output-row counter (rows per CTA fit 31 bits)
        uint32_t peek = mbar_peek(&bar_full[0], 0);   // try_wait result for the slot consumed NEXT (see mbar_peek)
        ROLL_PIECES_BEGIN
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int yb = ya + n - 1;
            for (int yy = ya - 1; yy <= yb + 1; ++yy) {
                // output rows fed by the input rows at yy: lo .. hi; row hi sits in the LOWEST slot (slots descend with the row
                // counter so that ascending TMEM columns meet the weight tile's ascending dy order)
                const int lo = max(yy - 1, ya), hi = min(yy + 1, yb);
                const int chi = c0 + (hi - ya);
                while (ready_rows <= chi) {
                    K3_TIMED(0, mbar_wait_lean(&bar_row_empty[kRollSlots - 1 - ready_m], ready_ph));
                    ++ready_rows;
                    if (++ready_m == kRollSlots) { ready_m = 0; ready_ph ^= 1; }
                }
                tc_fence_after();
                const int nrows = hi - lo + 1;
                const int him = (c0m + (hi - ya)) % kRollSlots;
                const int s_hi = kRollSlots - 1 - him;
                const int n1 = min(nrows, kRollSlots - s_hi);            // rows before the ring wraps
                const uint32_t wrow = (uint32_t)((yy - hi + 1) * kK3C);    // first weight row: dy of row hi = yy - hi
#pragma unroll 1
                for (int dz = 0; dz < 3; ++dz) {
                    K3_TIMED(1, mbar_wait_peeked(peek, &bar_full[ring_i], ring_ph));     // TMA data: the barrier's acquire is all it needs
                    const int cur_i = ring_i;
                    if (++ring_i == kRollRing) { ring_i = 0; ring_ph ^= 1; }
                    peek = mbar_peek(&bar_full[ring_i], ring_ph);                // the next slot's poll runs under this group's MMAs
                    const uint32_t a_lo = a_lo0 + (uint32_t)(cur_i * (kRollImg / 16));
                    const uint32_t w_lo = w_lo0 + (uint32_t)(dz * 9 * (kK3WTile3 / 16)) + wrow;
                    K3_TIMED(2, {
#pragma unroll 1
                        for (int part = 0; part < (n1 < nrows ? 2 : 1); ++part) {
                            const uint32_t acc = part ? tmem : tmem + (uint32_t)(s_hi * kK3C);
                            const int nr = part ? nrows - n1 : n1;
                            const uint32_t wl = w_lo + (uint32_t)(part ? n1 * kK3C : 0);
                            const uint32_t idesc = idesc0 + ((uint32_t)(nr * (kK3C >> 3)) << 17);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int ks = 0; ks < 3; ++ks)
                                    mma_ss_w(acc, a_lo + (uint32_t)(ks < 2 ? dx * 4 + ks * 2 : kRollOffB / 16 + dx * 2), ks < 2 ? a_hi : a_hi2,
                                             wl + (uint32_t)((dx * 3 + ks) * (kK3WTile3 / 16)), w_hi, idesc, 1u);
                        }
                        mma_commit_w(&bar_empty[cur_i]);
                    });
                }
                // row yy - 1 has received its three input rows
                if (yy - 1 >= ya) mma_commit_w(&bar_row_full[kRollSlots - 1 - (c0m + (yy - 1 - ya)) % kRollSlots]);
            }
            c0 += n;
            c0m = (c0m + n) % kRollSlots;
            r += n;
        }
        ROLL_PIECES_END
        if constexpr (PROF)
            if (lane == 0) {
                for (int q = 0; q < 6; ++q) a.prof[(blockIdx.x * 3 + 1) * 8 + q] = pw[q];
                unsigned long long gt_now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_now));
                a.prof[(blockIdx.x * 3 + 1) * 8 + 6] = (long long)gt_now;                   // wall clock (ns) when the issuer is done
                a.prof[(blockIdx.x * 3 + 1) * 8 + 7] = clock64() - p_t0;
            }
    } else {
        // ================================================ epilogue ====================================================
        const int quad = warp & 3;                          // a warp may touch TMEM lanes 32 * (warp id % 4) .. + 31 only
        const int et = quad * 32 + lane;                    // voxel x of this thread's TMEM lane
        const int ew = warp - 2;                            // 0..3
        const bool leader = warp == 2 && lane == 0;         // issues the bulk stores
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const bool dense = a.ys == kK3C;
        float st_s[kK3C], st_q[kK3C];                       // running sum / sum of squares of this thread's voxel column, per channel
#pragma unroll
        for (int i = 0; i < kK3C; ++i) st_s[i] = st_q[i] = 0.f;
        int64_t acc_b = -1;
        auto flush = [&]() {
            // reduce over the 128 voxel columns: butterfly inside each warp, the four warps through shared memory, one fp64 atomic
            // per (channel, moment) and CTA
            if (acc_b >= 0) {
#pragma unroll
                for (int i = 0; i < kK3C; ++i) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        st_s[i] += __shfl_xor_sync(0xffffffffu, st_s[i], o);
                        st_q[i] += __shfl_xor_sync(0xffffffffu, st_q[i], o);
                    }
                    if (lane == 0) { s_red[ew][2 * i] = st_s[i]; s_red[ew][2 * i + 1] = st_q[i]; }
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const int idx = ew * 32 + lane;
                if (idx < 2 * kK3C)
                    atomicAdd(a.sums + acc_b * 2 * kK3C + idx, (double)s_red[0][idx] + (double)s_red[1][idx] + (double)s_red[2][idx] + (double)s_red[3][idx]);
                asm volatile("bar.sync 2, 128;" ::: "memory");
            }
#pragma unroll
            for (int i = 0; i < kK3C; ++i) st_s[i] = st_q[i] = 0.f;
        };
        auto zero_and_release = [&](int slot) {
            uint32_t zeros[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zeros[i] = 0u;
#pragma unroll
            for (int c = 0; c < kK3C; c += 16) tmem_st16(tmem + lane_base + slot * kK3C + c, zeros);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_k3(&bar_row_empty[slot]);
        };
        for (int sl = 0; sl < kRollSlots; ++sl) zero_and_release(sl);
        int cm = 0, cstage = 0;          // output-row counter mod kRollSlots / mod kRollNStage
        uint32_t cph = 0;
        ROLL_PIECES_BEGIN
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int64_t b = p / a.D;
            if (b != acc_b) { flush(); acc_b = b; }
            for (int y = ya; y < ya + n; ++y) {
                const int slot = kRollSlots - 1 - cm;
                const uint32_t slot_ph = cph;
                uint16_t *stage = reinterpret_cast<uint16_t *>(sStage + cstage * kRollStage);
                if (++cm == kRollSlots) { cm = 0; cph ^= 1; }
                if (++cstage == kRollNStage) cstage = 0;
                const int64_t v0 = (p * a.H + y) * W;              // first voxel of the output row
                uint4 add[kK3Chunks];
                if (a.addend != nullptr) {
#pragma unroll
                    for (int k = 0; k < kK3Chunks; ++k) add[k] = *reinterpret_cast<const uint4 *>(a.addend + (v0 + et) * a.as_ + k * 8);
                }
                K3_TIMED(0, mbar_wait_warp_relaxed(&bar_row_full[slot], slot_ph));
                tc_fence_after();
                uint32_t rr[3][16];
                K3_TIMED(1, {
#pragma unroll
                    for (int k = 0; k < 3; ++k) tmem_ld16(tmem + lane_base + slot * kK3C + k * 16, rr[k]);
                    tmem_wait_ld();
                });
                K3_TIMED(2, zero_and_release(slot));
                // the staging tile this row goes to was last read by the bulk store of row c - kRollNStage
                K3_TIMED(3, {
                    if (leader) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kRollNStage - 1) : "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                });
                K3_TIMED(4, {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float f[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(rr[k][e]);
                        if (a.addend != nullptr) {
                            const uint4 a0 = add[2 * k], a1 = add[2 * k + 1];
                            const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float2 t = unpack_h16<F16>(aw[e]);
                                f[2 * e] += t.x;
                                f[2 * e + 1] += t.y;
                            }
                        }
                        uint32_t w8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            w8[e] = pack_h16<F16>(f[2 * e], f[2 * e + 1]);
                            stat_h16x2<F16>(w8[e], st_s[k * 16 + 2 * e], st_s[k * 16 + 2 * e + 1], st_q[k * 16 + 2 * e], st_q[k * 16 + 2 * e + 1]);
                        }
                        uint4 *dst = reinterpret_cast<uint4 *>(stage + (size_t)et * kK3C + k * 16);
                        dst[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                        dst[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                    }
                });
                K3_TIMED(5, {
                    if (dense) {
                        fence_proxy_async();                                // my staging writes -> visible to the bulk-copy engine
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (leader) {
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.y + v0 * kK3C), "r"(smem_u32(stage)),
                                         "r"(kRollStage)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else {
                        // y is a channel slice of a wider buffer: 96-byte pieces, copied by the threads
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        for (int i = et; i < 128 * kK3Chunks; i += 128) {
                            const int vx = i / kK3Chunks, pc = i - vx * kK3Chunks;
                            *reinterpret_cast<uint4 *>(a.y + (v0 + vx) * a.ys + pc * 8) = *reinterpret_cast<const uint4 *>(stage + (size_t)vx * kK3C + pc * 8);
                        }
                        if constexpr (kRollNStage == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
                    }
                });
            }
            r += n;
        }
        ROLL_PIECES_END
        flush();
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if constexpr (PROF)
            if (et == 0) {
                for (int q = 0; q < 6; ++q) a.prof[(blockIdx.x * 3 + 2) * 8 + q] = pw[q];
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                a.prof[(blockIdx.x * 3 + 2) * 8 + 6] = (long long)smid;
                a.prof[(blockIdx.x * 3 + 2) * 8 + 7] = clock64() - p_t0;
            }
    }
#undef K3_TIMED
#undef ROLL_PIECES_BEGIN
#undef ROLL_PIECES_END
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void k3_finalize_kernel(const double *__restrict__ sums, float *__restrict__ mr, int n, double inv_s, double eps) {
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

// cuTensorMapEncodeTiled is a DRIVER entry point: it is looked up through the runtime at first use, so the library carries no
// link-time dependency on libcuda.so.1 and still loads (for the symbol check) on a machine without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int k3_launch(const void *x, int dtype, const void *wpack, const void *addend, void *y, double *sums, float *mean_rstd,
                     const float *in_mean_rstd, float slope, float eps, int B, int D, int H, int W,
                     int64_t x_vox_stride, int64_t add_vox_stride, int64_t y_vox_stride, long long *prof, void *stream) {
    if (!x || !wpack || !y || !sums || !mean_rstd) return WF_ERR_NULL_POINTER;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (B <= 0 || D <= 0 || H <= 0 || W != 128) return WF_ERR_BAD_SHAPE;
    if (x_vox_stride < kK3C || y_vox_stride < kK3C || x_vox_stride % 8 || y_vox_stride % 8) return WF_ERR_BAD_SHAPE;
    if (!aligned16(x) || !aligned16(wpack) || !aligned16(y)) return WF_ERR_MISALIGNED;
    if (addend && (in_mean_rstd || add_vox_stride < kK3C || add_vox_stride % 8)) return WF_ERR_BAD_SHAPE;
    if (addend && !aligned16(addend)) return WF_ERR_MISALIGNED;
    if (prof && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long attr_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attr_done)) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<false, 6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, roll_smem(6, 2)));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<true, 6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, roll_smem(6, 2)));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<true, 6, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, roll_smem(6, 2)));
    }
    WF_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * kK3C, st));
    if (in_mean_rstd == nullptr) {
        // rolling-row kernel: the run of B * D * H output rows is split evenly over one CTA per SM.  The input is described by a
        // TMA tensor map [48 channels][128 x][B * D * H rows] (fp16 and bf16 are both "16-bit, no conversion" to the copy engine).
        const int64_t rows = (int64_t)B * D * H;
        if (rows > 0x7fffffff) return WF_ERR_BAD_SHAPE;
        CUtensorMap xmap, xmap_hi;
        const cuuint64_t gdim[3] = {(cuuint64_t)kK3C, 128, (cuuint64_t)rows};
        const cuuint64_t gstr[2] = {(cuuint64_t)x_vox_stride * 2, (cuuint64_t)x_vox_stride * 2 * 128};
        const cuuint32_t box_lo[3] = {32, 130, 1}, box_hi[3] = {16, 130, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const EncodeTiledFn encode = encode_tiled_fn();
        if (encode == nullptr ||
            encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(x), gdim, gstr, box_lo, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
            encode(&xmap_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(x), gdim, gstr, box_hi, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return WF_ERR_CUDA;
        K3RollArgs r;
        r.wpack = (const uint16_t *)wpack; r.y = (uint16_t *)y; r.addend = (const uint16_t *)addend;
        r.sums = sums; r.ys = y_vox_stride; r.as_ = add_vox_stride; r.B = B; r.D = D; r.H = H; r.prof = prof;
        const int grid = (int)(rows < kNumSMs ? rows : kNumSMs);
        // 6 ring slots + 2 staging tiles
        if (prof != nullptr)
            conv3d_k3_c48_roll_kernel<true, 6, 2, true><<<grid, kRollThreads, roll_smem(6, 2), st>>>(xmap, xmap_hi, r);
        else if (dtype == WF_F16)
            conv3d_k3_c48_roll_kernel<true, 6, 2><<<grid, kRollThreads, roll_smem(6, 2), st>>>(xmap, xmap_hi, r);
        else
            conv3d_k3_c48_roll_kernel<false, 6, 2><<<grid, kRollThreads, roll_smem(6, 2), st>>>(xmap, xmap_hi, r);
    } else {
        K3Args a;
        a.x = (const uint16_t *)x; a.wpack = (const uint16_t *)wpack; a.y = (uint16_t *)y; a.sums = sums;
        a.in_mr = in_mean_rstd; a.slope = slope; a.xs = x_vox_stride; a.ys = y_vox_stride; a.B = B; a.D = D; a.H = H; a.prof = prof;
        const int64_t nblocks = (int64_t)B * D * ((H + 3) / 4);
        const int grid = (int)(nblocks < kNumSMs ? nblocks : kNumSMs);
        if (prof != nullptr)
            conv3d_k3_c48_kernel<true, true><<<grid, 288, kK3Smem, st>>>(a);
        else if (dtype == WF_F16)
            conv3d_k3_c48_kernel<true><<<grid, 288, kK3Smem, st>>>(a);
        else
            conv3d_k3_c48_kernel<false><<<grid, 288, kK3Smem, st>>>(a);
    }
    WF_LAUNCH_CHECK();
    const int n = B * kK3C;
    k3_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(sums, mean_rstd, n, 1.0 / ((double)D * H * W), (double)eps);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_conv3d_k3_c48_in_stats(const void *x, int dtype, const void *wpack, void *y, double *sums, float *mean_rstd,
                                         const float *in_mean_rstd, float slope, float eps, int B, int D, int H, int W,
                                         int64_t x_vox_stride, int64_t y_vox_stride, void *stream) {
    return k3_launch(x, dtype, wpack, nullptr, y, sums, mean_rstd, in_mean_rstd, slope, eps, B, D, H, W, x_vox_stride, 0, y_vox_stride,
                     nullptr, stream);
}

extern "C" int wf_conv3d_k3_c48_add_stats(const void *x, int dtype, const void *wpack, const void *addend, void *y, double *sums,
                                          float *mean_rstd, float eps, int B, int D, int H, int W, int64_t x_vox_stride,
                                          int64_t add_vox_stride, int64_t y_vox_stride, void *stream) {
    return k3_launch(x, dtype, wpack, addend, y, sums, mean_rstd, nullptr, 0.f, eps, B, D, H, W, x_vox_stride, add_vox_stride, y_vox_stride,
                     nullptr, stream);
}

extern "C" int wf_conv3d_k3_c48_stage_clocks(const void *x, int dtype, const void *wpack, void *y, double *sums, float *mean_rstd,
                                             const float *in_mean_rstd, float slope, float eps, int B, int D, int H, int W,
                                             int64_t x_vox_stride, int64_t y_vox_stride, long long *clocks, void *stream) {
    if (!clocks) return WF_ERR_NULL_POINTER;
    return k3_launch(x, dtype, wpack, nullptr, y, sums, mean_rstd, in_mean_rstd, slope, eps, B, D, H, W, x_vox_stride, 0, y_vox_stride, clocks,
                     stream);
}
