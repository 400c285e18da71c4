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
// showed (scripts/k3_stage_clocks.py): its loader warps are busy 89 % of the time - a thread copying "its" voxel makes every
// warp-wide cp.async touch 24 different 128-byte lines - the issuer waits for staged rows, and a third of the MMAs are edge
// rows of a 4-row block with N = 48 / 96, which cost the same 66-72 clk as N = 144 (scripts/mma_probe.cu).  Here:
//   * a CTA owns a contiguous RUN of output rows (B * D * H rows split evenly over the grid) and walks it in y.  The accumulators
//     of the output rows in flight form a ring of 10 TMEM slots (48 columns each); the input row (z + dz, y') feeds the output
//     rows y' - 1, y', y' + 1, whose slots are adjacent, so EVERY MMA is 128 x 144 x 16 (two smaller ones where the ring wraps or a
//     run starts / ends): 27 MMAs and 3 staged input rows per output row instead of 40.5 and 4.5;
//   * an output row is complete one step after its own y, the epilogue drains it (TMEM -> 16 bit -> staging -> coalesced stores +
//     statistics) while the tensor pipe is three rows ahead, zeroes the slot and hands it back;
//   * loader threads copy CONSECUTIVE 16-byte chunks of the row (chunk q = voxel * 6 + channel chunk), so a warp-wide cp.async
//     reads 512 contiguous bytes; the issuer is one converged warp (tc_common.cuh, "warp-uniform issue").
// `addend` (optional): a second 16-bit tensor laid out like y that is added to the accumulators before rounding - the second
// pass of a convolution whose input channels are split over two launches (decoder1.conv1: 96 = 48 + 48 input channels).
constexpr int kRollSlots = 10;                                 // accumulator ring: output rows in flight
constexpr int kRollRing = 6;                                   // staged input rows
constexpr int kRollStage = 128 * (kK3C + 8) * 2;               // one output row, 16 bit, 112-byte pitch: 14336 bytes
constexpr int kRollSmem = kK3WBytes + kRollRing * kK3RowImg + 2 * kRollStage;   // 229120

struct K3RollArgs {
    const uint16_t *x;          // [B, D, H, 128, 48 of xs]
    const uint16_t *wpack;      // as K3Args
    uint16_t *y;                // [B, D, H, 128, 48 of ys]
    const uint16_t *addend;     // optional, voxel stride as_
    double *sums;
    int64_t xs, ys, as_;
    int B, D, H;
    long long *prof;
};

template <bool F16, bool PROF = false>
__global__ void __launch_bounds__(288, 1) conv3d_k3_c48_roll_kernel(K3RollArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[kRollRing], bar_empty[kRollRing], bar_row_full[kRollSlots], bar_row_empty[kRollSlots];
    __shared__ uint32_t tmem_slot;
    uint8_t *sW = smem;
    uint8_t *sRing = smem + kK3WBytes;
    uint8_t *sStage = smem + kK3WBytes + kRollRing * kK3RowImg;
    const int tid = threadIdx.x, warp = (int)warp_idx_uniform(), lane = tid & 31;
    constexpr int W = 128;
    // this CTA's run of output rows [r0, r1) in (b, z, y) order
    const int64_t R = (int64_t)a.B * a.D * a.H;
    const int64_t r0 = R * blockIdx.x / gridDim.x, r1 = R * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < kRollRing; ++i) {
            mbar_init(&bar_full[i], 128);        // one deferred arrival per loader thread
            mbar_init(&bar_empty[i], 1);         // tcgen05.commit
        }
        for (int i = 0; i < kRollSlots; ++i) {
            mbar_init(&bar_row_full[i], 1);      // tcgen05.commit
            mbar_init(&bar_row_empty[i], 4);     // one per epilogue warp
        }
        mbar_fence_init();
    }
    for (int i = tid; i < kK3WBytes / 16; i += 288)
        reinterpret_cast<uint4 *>(sW)[i] = __ldg(reinterpret_cast<const uint4 *>(a.wpack) + i);
    for (int i = tid; i < kRollRing * kK3Chunks * 4; i += 288) {        // zero halo rows 0, 129, 130, 131 of every chunk plane
        const int slot = i / (kK3Chunks * 4), r = i % (kK3Chunks * 4);
        const int ch = r >> 2, which = r & 3;
        const int row = which == 0 ? 0 : 128 + which;
        *reinterpret_cast<uint4 *>(sRing + slot * kK3RowImg + (ch * kK3Rows + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long pw[7] = {0, 0, 0, 0, 0, 0, 0};
    const long long p_t0 = PROF ? clock64() : 0;
#define K3_TIMED(idx, ...) do { if constexpr (PROF) { const long long t_ = clock64(); __VA_ARGS__; pw[idx] += clock64() - t_; } else { __VA_ARGS__; } } while (0)

    if (warp < 4) {
        // ================================================= loaders ====================================================
        // chunk q = it * 128 + tid of a row: voxel q / 6, channel chunk q % 6 (consecutive threads -> consecutive 16 bytes)
        int soff[kK3Chunks], doff[kK3Chunks];
#pragma unroll
        for (int it = 0; it < kK3Chunks; ++it) {
            const int q = it * 128 + tid, vx = q / kK3Chunks, ch = q - vx * kK3Chunks;
            soff[it] = (int)(vx * a.xs) + ch * 8;
            doff[it] = (ch * kK3Rows + vx + 1) * 16;
        }
        uint32_t n_loaded = 0, n_arrived = 0;
        auto announce_upto = [&](uint32_t upto) {
            for (; n_arrived < upto; ++n_arrived) {
                fence_proxy_async();
                mbar_arrive_k3(&bar_full[n_arrived % kRollRing]);
            }
        };
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int yb = ya + n - 1;
            const int64_t b = p / a.D;
            const int z = (int)(p - b * a.D);
            const int ys = max(ya - 1, 0), ye = min(yb + 1, a.H - 1);
            for (int yy = ys; yy <= ye; ++yy)
                for (int dz = -1; dz <= 1; ++dz) {
                    const int zz = z + dz;
                    if ((unsigned)zz >= (unsigned)a.D) continue;
                    const int slot = n_loaded % kRollRing;
                    const uint32_t use = n_loaded / kRollRing;
                    if (use > 0) K3_TIMED(0, mbar_wait(&bar_empty[slot], (use - 1) & 1));
                    uint8_t *img = sRing + slot * kK3RowImg;
                    const uint16_t *src = a.x + ((((int64_t)b * a.D + zz) * a.H + yy) * W) * a.xs;
                    K3_TIMED(2, {
#pragma unroll
                    for (int it = 0; it < kK3Chunks; ++it) cp_async16_k3(img + doff[it], src + soff[it]);
                    asm volatile("cp.async.commit_group;" ::: "memory"); });
                    if (n_loaded + 1 - n_arrived > kK3Lag) {
                        K3_TIMED(1, asm volatile("cp.async.wait_group %0;" ::"n"(kK3Lag) : "memory"));
                        K3_TIMED(3, announce_upto(n_loaded + 1 - kK3Lag));
                    }
                    ++n_loaded;
                }
            r += n;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        announce_upto(n_loaded);
        if constexpr (PROF)
            if (tid == 0) {
                for (int q = 0; q < 7; ++q) a.prof[(blockIdx.x * 3 + 0) * 8 + q] = pw[q];
                a.prof[(blockIdx.x * 3 + 0) * 8 + 7] = clock64() - p_t0;
            }
    } else if (warp == 4) {
        // ================================================= issuer =====================================================
        const uint32_t idesc1 = instr_desc_h16<F16>(128, kK3C, false), idesc2 = instr_desc_h16<F16>(128, 2 * kK3C, false),
                       idesc3 = instr_desc_h16<F16>(128, 3 * kK3C, false);
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(sRing), kK3Rows * 16), a_hi = smem_desc_hi(128);
        const uint32_t w_lo0 = smem_desc_lo(smem_u32(sW), 3 * kK3C * 16), w_hi = smem_desc_hi(128);
        auto mma_set = [&](uint32_t a_lo, uint32_t w_lo, uint32_t acc, int nrows) {
            const uint32_t idesc = nrows == 3 ? idesc3 : (nrows == 2 ? idesc2 : idesc1);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int ks = 0; ks < 3; ++ks)
                    mma_ss_w(acc, a_lo + (uint32_t)(ks * 2 * kK3Rows + dx), a_hi, w_lo + (uint32_t)((dx * 3 + ks) * (kK3WTile3 / 16)), w_hi,
                             idesc, 1u);
        };
        uint32_t n_used = 0;
        int64_t c0 = 0, c_ready = 0;     // output-row counters of this CTA: first row of the segment / rows whose slot is known free
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int yb = ya + n - 1;
            const int z = (int)(p % a.D);
            const int ys = max(ya - 1, 0), ye = min(yb + 1, a.H - 1);
            for (int yy = ys; yy <= ye; ++yy) {
                // output rows fed by the input rows at yy: lo .. hi; row hi sits in the LOWEST slot (slots descend with the row
                // counter so that ascending TMEM columns meet the weight tile's ascending dy order)
                const int lo = max(yy - 1, ya), hi = min(yy + 1, yb);
                const int64_t chi = c0 + (hi - ya);
                for (; c_ready <= chi; ++c_ready)
                    K3_TIMED(0, mbar_wait_warp(&bar_row_empty[kRollSlots - 1 - (int)(c_ready % kRollSlots)], (uint32_t)(c_ready / kRollSlots) & 1));
                tc_fence_after();
                const int nrows = hi - lo + 1;
                const int s_hi = kRollSlots - 1 - (int)(chi % kRollSlots);
                const int n1 = min(nrows, kRollSlots - s_hi);            // rows before the ring wraps
                const uint32_t wrow = (uint32_t)((yy - hi + 1) * kK3C);    // first weight row: dy of row hi = yy - hi
                for (int dz = -1; dz <= 1; ++dz) {
                    if ((unsigned)(z + dz) >= (unsigned)a.D) continue;
                    const int slot = n_used % kRollRing;
                    K3_TIMED(1, mbar_wait_warp(&bar_full[slot], (n_used / kRollRing) & 1));
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)(slot * (kK3RowImg / 16));
                    const uint32_t w_lo = w_lo0 + (uint32_t)((dz + 1) * 9 * (kK3WTile3 / 16)) + wrow;
                    K3_TIMED(2, {
                    mma_set(a_lo, w_lo, tmem + (uint32_t)(s_hi * kK3C), n1);
                    if (n1 < nrows) mma_set(a_lo, w_lo + (uint32_t)(n1 * kK3C), tmem, nrows - n1);
                    mma_commit_w(&bar_empty[slot]); });
                    ++n_used;
                }
                if (yy - 1 >= ya) mma_commit_w(&bar_row_full[kRollSlots - 1 - (int)((c0 + (yy - 1 - ya)) % kRollSlots)]);
                if (yy == ye && yy == yb) mma_commit_w(&bar_row_full[kRollSlots - 1 - (int)((c0 + (yb - ya)) % kRollSlots)]);
            }
            c0 += n;
            r += n;
        }
        if constexpr (PROF)
            if (lane == 0) {
                for (int q = 0; q < 7; ++q) a.prof[(blockIdx.x * 3 + 1) * 8 + q] = pw[q];
                a.prof[(blockIdx.x * 3 + 1) * 8 + 7] = clock64() - p_t0;
            }
    } else {
        // ================================================ epilogue ====================================================
        const int quad = warp & 3;
        const int et = quad * 32 + lane;                    // voxel x of this thread's TMEM lane
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        constexpr int pitch = kK3C + 8;
        double acc_s[2] = {0.0, 0.0}, acc_q[2] = {0.0, 0.0};
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
        int64_t c = 0;
        for (int64_t r = r0; r < r1;) {
            const int64_t p = r / a.H;
            const int ya = (int)(r - p * a.H);
            const int n = (int)min((int64_t)(a.H - ya), r1 - r);
            const int64_t b = p / a.D;
            if (b != acc_b) { flush(); acc_b = b; }
            for (int y = ya; y < ya + n; ++y, ++c) {
                const int slot = kRollSlots - 1 - (int)(c % kRollSlots);
                uint16_t *stage = reinterpret_cast<uint16_t *>(sStage + (c & 1) * kRollStage);
                const int64_t v0 = (p * a.H + y) * W;              // first voxel of the output row
                uint4 add[kK3Chunks];
                if (a.addend != nullptr) {
#pragma unroll
                    for (int k = 0; k < kK3Chunks; ++k) add[k] = *reinterpret_cast<const uint4 *>(a.addend + (v0 + et) * a.as_ + k * 8);
                }
                K3_TIMED(0, mbar_wait(&bar_row_full[slot], (uint32_t)(c / kRollSlots) & 1));
                tc_fence_after();
                uint32_t rr[3][16];
                K3_TIMED(1, {
#pragma unroll
                for (int k = 0; k < 3; ++k) tmem_ld16(tmem + lane_base + slot * kK3C + k * 16, rr[k]);
                tmem_wait_ld(); });
                K3_TIMED(2, zero_and_release(slot));
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
                    uint4 lo4, hi4;
                    lo4.x = pack_h16<F16>(f[0], f[1]);   lo4.y = pack_h16<F16>(f[2], f[3]);   lo4.z = pack_h16<F16>(f[4], f[5]);   lo4.w = pack_h16<F16>(f[6], f[7]);
                    hi4.x = pack_h16<F16>(f[8], f[9]);   hi4.y = pack_h16<F16>(f[10], f[11]); hi4.z = pack_h16<F16>(f[12], f[13]); hi4.w = pack_h16<F16>(f[14], f[15]);
                    uint4 *dst = reinterpret_cast<uint4 *>(stage + (size_t)et * pitch + k * 16);
                    dst[0] = lo4;
                    dst[1] = hi4;
                }
                // staging of this row complete (epilogue warps only); the other staging buffer's readers all passed this barrier too
                K3_TIMED(3, asm volatile("bar.sync 2, 128;" ::: "memory"));
                K3_TIMED(4, {
                for (int i = et; i < 128 * kK3Chunks; i += 128) {
                    const int vx = i / kK3Chunks, pc = i - vx * kK3Chunks;
                    *reinterpret_cast<uint4 *>(a.y + (v0 + vx) * a.ys + pc * 8) = *reinterpret_cast<const uint4 *>(stage + (size_t)vx * pitch + pc * 8);
                } });
                const long long t_stats = PROF ? clock64() : 0;
                if (et < 96) {
                    const uint32_t *col = reinterpret_cast<const uint32_t *>(stage) + cpair;
                    constexpr int wpitch = pitch / 2;
                    float s0 = 0.f, s1 = 0.f, qq0 = 0.f, qq1 = 0.f;
#pragma unroll 8
                    for (int q = rq * 32; q < rq * 32 + 32; ++q) stat_h16x2<F16>(col[(size_t)q * wpitch], s0, s1, qq0, qq1);
                    acc_s[0] += (double)s0; acc_s[1] += (double)s1; acc_q[0] += (double)qq0; acc_q[1] += (double)qq1;
                }
                if constexpr (PROF) pw[5] += clock64() - t_stats;
            }
            r += n;
        }
        flush();
        if constexpr (PROF)
            if (et == 0) {
                for (int q = 0; q < 7; ++q) a.prof[(blockIdx.x * 3 + 2) * 8 + q] = pw[q];
                a.prof[(blockIdx.x * 3 + 2) * 8 + 7] = clock64() - p_t0;
            }
    }
#undef K3_TIMED
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
    if ((int64_t)127 * x_vox_stride + 40 > 0x7fffffff) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long attr_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attr_done)) {
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK3Smem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRollSmem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRollSmem));
        WF_CUDA_CHECK(cudaFuncSetAttribute(conv3d_k3_c48_roll_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRollSmem));
    }
    WF_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * kK3C, st));
    if (in_mean_rstd == nullptr) {
        // rolling-row kernel: the run of B * D * H output rows is split evenly over one CTA per SM
        K3RollArgs r;
        r.x = (const uint16_t *)x; r.wpack = (const uint16_t *)wpack; r.y = (uint16_t *)y; r.addend = (const uint16_t *)addend;
        r.sums = sums; r.xs = x_vox_stride; r.ys = y_vox_stride; r.as_ = add_vox_stride; r.B = B; r.D = D; r.H = H; r.prof = prof;
        const int64_t rows = (int64_t)B * D * H;
        const int grid = (int)(rows < kNumSMs ? rows : kNumSMs);
        if (prof != nullptr)
            conv3d_k3_c48_roll_kernel<true, true><<<grid, 288, kRollSmem, st>>>(r);
        else if (dtype == WF_F16)
            conv3d_k3_c48_roll_kernel<true><<<grid, 288, kRollSmem, st>>>(r);
        else
            conv3d_k3_c48_roll_kernel<false><<<grid, 288, kRollSmem, st>>>(r);
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
