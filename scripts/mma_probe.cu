// tcgen05.mma operand-layout probe (sm_100a).  Two questions the 3^3 convolution kernel's design hangs on:
//   1. timing: clocks per tcgen05.mma (M = 128, K = 16, fp16, A and B from shared memory) as a function of N and of the
//      shared-memory layout (no swizzle with two chunk arrangements, 32 / 64 / 128-byte swizzle);
//   2. semantics: does an operand whose descriptor start address is advanced by ONE ROW (not a whole 8-row atom) still read
//      the rows a writer placed with the address-based XOR swizzle?  (That is what lets the three dx taps of a 3^3
//      convolution share one staged image.)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/mma_probe scripts/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include "../waveformer_b200/csrc/tc_common.cuh"

using namespace wf::tc;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

struct TimingCfg {
    uint32_t layout_a, lbo_a, sbo_a, slot_stride_a;   // A: slots cycle so consecutive MMAs read different addresses
    uint32_t layout_b, lbo_b, sbo_b, tile_stride_b;
    int N, iters;
};

__global__ void __launch_bounds__(128, 1) timing_kernel(TimingCfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    // some finite data (zeros would let a clever pipe skip nothing, but keep values tame anyway)
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid < 32) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_h16<true>(128, c.N, false);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        // warm-up
        for (int i = 0; i < 16; ++i)
            mma_ss(tmem, make_desc(a0, c.lbo_a, c.sbo_a, c.layout_a, 0), make_desc(b0, c.lbo_b, c.sbo_b, c.layout_b, 0), idesc, 1u);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) {
            const uint32_t sa = a0 + (i & 3) * c.slot_stride_a, sb = b0 + (i % 9) * c.tile_stride_b;
            mma_ss(tmem + ((i & 1) ? 256 : 0), make_desc(sa, c.lbo_a, c.sbo_a, c.layout_a, 0),
                   make_desc(sb, c.lbo_b, c.sbo_b, c.layout_b, 0), idesc, 1u);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// Lean issue loop: four precomputed descriptor pairs, unrolled - separates the tensor pipe's minimum interval from the issuing
// thread's own instruction count.  ts != 0: the A operand comes from TMEM (columns 448..) instead of shared memory.
__global__ void __launch_bounds__(128, 1) timing_lean_kernel(TimingCfg c, int ts, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid < 32) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_h16<true>(128, c.N, false);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        uint64_t da[4], db[4];
        for (int i = 0; i < 4; ++i) {
            da[i] = make_desc(a0 + i * c.slot_stride_a, c.lbo_a, c.sbo_a, c.layout_a, 0);
            db[i] = make_desc(b0 + i * c.tile_stride_b, c.lbo_b, c.sbo_b, c.layout_b, 0);
        }
        const long long t0 = clock64();
        for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ts) mma_ts(tmem + (j & 1) * 224, tmem + 448 + j * 8, db[j], idesc, 1u);
                else mma_ss(tmem + (j & 1) * 256, da[j], db[j], idesc, 1u);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// Pacing study: which of {issue pacing, operand address cycling, accumulator cycling} moves the per-MMA interval.
// na / nb = number of distinct A slots / B tiles cycled (1, 2 or 4), nacc = accumulators cycled (1 or 2), gap = minimum clocks
// between two issues (0 = back to back).
__global__ void __launch_bounds__(128, 1) timing_pace_kernel(TimingCfg c, int na, int nb, int nacc, int gap, long long *out,
                                                             int a_off = 0, int b_off = 0, int d_off = 0, int a_cycle = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid < 32) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_h16<true>(128, c.N, false);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        uint64_t da[4], db[4];
        uint32_t dd[4];
        for (int i = 0; i < 4; ++i) {
            da[i] = make_desc(a0 + (i % na) * c.slot_stride_a + a_off + (a_cycle ? (i % 3) * a_cycle : 0), c.lbo_a, c.sbo_a, c.layout_a, 0);
            db[i] = make_desc(b0 + (i % nb) * c.tile_stride_b + b_off, c.lbo_b, c.sbo_b, c.layout_b, 0);
            dd[i] = tmem + (i % nacc) * 256 + d_off;
        }
        const long long t0 = clock64();
        long long next = t0;
        for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (gap) { while (clock64() < next) {} next += gap; }
                mma_ss(dd[j], da[j], db[j], idesc, 1u);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// Interference study at N = 144 (the convolution's shape): tcgen05.commit every `commit_every` MMAs (arrivals on a barrier nobody
// waits for), and `bg_warps` other warps hammering shared memory (mode 1: 16-byte stores, 2: 16-byte loads, 3: cp.async 16-byte
// copies from global memory) while the MMAs run.
__global__ void __launch_bounds__(288, 1) timing_interf_kernel(TimingCfg c, int commit_every, int bg_warps, int bg_mode, const uint4 *gsrc,
                                                               long long *out, unsigned *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 200 * 1024 / 4; i += 288) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid < 32) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1u << 20); mbar_fence_init(); done = 0; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_h16<true>(128, c.N, false);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        uint64_t da[4], db[4];
        for (int i = 0; i < 4; ++i) {
            da[i] = make_desc(a0 + i * c.slot_stride_a + (i % 3) * 16, c.lbo_a, c.sbo_a, c.layout_a, 0);
            db[i] = make_desc(b0 + i * c.tile_stride_b, c.lbo_b, c.sbo_b, c.layout_b, 0);
        }
        const long long t0 = clock64();
        int since = 0;
        for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                mma_ss(tmem + (j & 1) * 256, da[j], db[j], idesc, 1u);
                if (commit_every && ++since == commit_every) { mma_commit(&bar2); since = 0; }
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        done = 1;
    } else if (warp >= 1 && warp <= bg_warps) {
        // background traffic on a region the MMAs do not read: [128 KB, 192 KB)
        uint4 *region = reinterpret_cast<uint4 *>(smem + 128 * 1024);
        uint4 v = make_uint4(tid, 1, 2, 3);
        unsigned acc = 0;
        int k = 0;
        while (!done) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = ((warp - 1) * 512 + ((k + u) & 15) * 32 + lane) & 4095;
                if (bg_mode == 1) region[idx] = v;
                else if (bg_mode == 2) { const uint4 r = region[idx]; acc += r.x; }
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(region + idx)), "l"(gsrc + (size_t)blockIdx.x * 4096 + idx) : "memory");
            }
            if (bg_mode == 3) { asm volatile("cp.async.commit_group;" ::: "memory"); asm volatile("cp.async.wait_group 2;" ::: "memory"); }
            k += 8;
        }
        if (bg_mode == 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (acc == 0xdeadbeef) sink[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// Rolling-accumulator study: the 3^3 convolution's MMA stream in isolation.  27 MMAs of N = 144 accumulate into one 144-column
// window of a 480-column ring; the window then moves on by 48 columns (splitting in two where the ring wraps).  Options (bits):
//   1: tcgen05.commit after every 9 MMAs and after every 27 (arrivals on barriers nobody waits for)
//   2: four other warps keep reading (tcgen05.ld) and zeroing (tcgen05.st) 48-column slots the MMAs are not using
//   4: a bulk copy (TMA engine) of 16 KB global -> shared lands in a free operand slot every 9 MMAs (fire and forget)
//   8: the window does NOT move (control)
//  16: tcgen05.fence::after_thread_sync before every group of 9 MMAs;  32: a (satisfied) mbarrier wait by the elected lane + __syncwarp too
__global__ void __launch_bounds__(192, 1) roll_probe_kernel(int opts, int steps, const uint8_t *gsrc, long long *out, unsigned *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar_sink, bar_tma, bar_done0;
    __shared__ uint32_t slot;
    __shared__ volatile int done, cur_step;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 200 * 1024 / 4; i += 192) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid < 32) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar_sink, 1u << 20); mbar_init(&bar_tma, 1u << 20); mbar_init(&bar_done0, 1); mbar_fence_init(); done = 0; cur_step = 0;
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_done0)) : "memory"); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        // converged warp, elected lane issues (as the convolution does)
        const uint32_t idesc3 = instr_desc_h16<true>(128, 144, false), idesc2 = instr_desc_h16<true>(128, 96, false), idesc1 = instr_desc_h16<true>(128, 48, false);
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16), a_hi = smem_desc_hi(1024) | (2u << 29);             // SW128 images, 17 KB apart
        const uint32_t w_lo0 = smem_desc_lo(smem_u32(smem) + 4 * 17408, 144 * 16), w_hi = smem_desc_hi(128);       // 27 weight tiles
        const long long t0 = clock64();
        int ring = 0, cm = 0;
        for (int st = 0; st < steps; ++st) {
            const int s_hi = (opts & 8) ? 0 : 9 - cm;
            const int n1 = min(3, 10 - s_hi);
            for (int dz = 0; dz < 3; ++dz) {
                if (opts & 32) mbar_wait_warp(&bar_done0, 0);
                if (opts & 16) tc_fence_after();
                const uint32_t a_lo = a_lo0 + ring * (17408 / 16), w_lo = w_lo0 + dz * 9 * (4608 / 16);
#pragma unroll
                for (int j = 0; j < 9; ++j)
                    mma_ss_w(tmem + s_hi * 48, a_lo + (j / 3) * 8 + (j % 3) * 2, a_hi, w_lo + j * (4608 / 16), w_hi, n1 == 3 ? idesc3 : (n1 == 2 ? idesc2 : idesc1), 1u);
                if (n1 < 3) {
#pragma unroll
                    for (int j = 0; j < 9; ++j)
                        mma_ss_w(tmem, a_lo + (j / 3) * 8 + (j % 3) * 2, a_hi, w_lo + j * (4608 / 16) + n1 * 48, w_hi, n1 == 1 ? idesc2 : idesc1, 1u);
                }
                if (opts & 1) mma_commit_w(&bar_sink);
                if ((opts & 4) && elect_one()) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(smem) + ((ring + 2) & 3) * 17408),
                                 "l"(gsrc + (size_t)blockIdx.x * 65536 + ((st * 3 + dz) & 3) * 16384), "r"(16384), "r"(smem_u32(&bar_tma))
                                 : "memory");
                }
                __syncwarp();
                ring = (ring + 1) & 3;
            }
            if (opts & 1) mma_commit_w(&bar_sink);
            if (++cm == 10) cm = 0;
            if (lane == 0) cur_step = st;
        }
        mma_commit_w(&bar);
        mbar_wait_warp(&bar, 0);
        if (lane == 0) { out[blockIdx.x] = clock64() - t0; done = 1; }
    } else if (warp >= 2 && (opts & 2)) {
        // epilogue-like traffic on the slot 5 positions behind the window
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        unsigned acc = 0;
        while (!done) {
            const int sl = (9 - (cur_step % 10) + 5) % 10;
            uint32_t r[16];
            for (int c = 0; c < 48; c += 16) { tmem_ld16(tmem + lane_base + sl * 48 + c, r); tmem_wait_ld(); acc += r[3]; }
            for (int i = 0; i < 16; ++i) r[i] = 0;
            for (int c = 0; c < 48; c += 16) tmem_st16(tmem + lane_base + sl * 48 + c, r);
            tmem_wait_st();
            __nanosleep(500);
        }
        if (acc == 0xdeadbeef) sink[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// ---- semantics --------------------------------------------------------------------------------------------------------------
struct SemCfg {
    uint32_t layout_a, lbo_a, sbo_a, start_off_a, base_off_a;   // A descriptor (start = image base + start_off_a)
    uint32_t layout_b, lbo_b, sbo_b;
    uint32_t a_bytes, b_bytes;
};

// D[128 x 16] = A[128 x 16] * B[16 x 16]^T with B = identity, so D[r][n] = A[r][n] as the tensor core read it
__global__ void __launch_bounds__(128, 1) sem_kernel(SemCfg c, const uint8_t *a_img, const uint8_t *b_img, float *d_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sA = smem, *sB = smem + 64 * 1024;
    for (uint32_t i = tid; i < c.a_bytes; i += 128) sA[i] = a_img[i];
    for (uint32_t i = tid; i < c.b_bytes; i += 128) sB[i] = b_img[i];
    if (tid < 32) tmem_alloc(&slot, 32);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_h16<true>(128, 16, false);
        mma_ss(tmem, make_desc(smem_u32(sA) + c.start_off_a, c.lbo_a, c.sbo_a, c.layout_a, c.base_off_a),
               make_desc(smem_u32(sB), c.lbo_b, c.sbo_b, c.layout_b, 0), idesc, 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_wait_ld();
    for (int n = 0; n < 16; ++n) d_out[(warp * 32 + lane) * 16 + n] = __uint_as_float(r[n]);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 32);
}

static uint32_t swz(uint32_t p, int xbytes) {   // address-based XOR swizzle of a byte offset inside a 1024-aligned region
    const uint32_t mask = xbytes / 16 - 1;      // 128 B: 7, 64 B: 3, 32 B: 1, 16 (none): 0
    return p ^ (((p >> 7) & mask) << 4);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char **argv) {
    CK(cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(timing_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(timing_pace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(timing_interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(roll_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(sem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    long long *d_out;
    CK(cudaMalloc(&d_out, 148 * sizeof(long long)));
    {
        uint8_t *gsrc; unsigned *sink;
        CK(cudaMalloc(&gsrc, (size_t)148 * 65536)); CK(cudaMemset(gsrc, 0x3c, (size_t)148 * 65536)); CK(cudaMalloc(&sink, 4));
        const int steps = 222;
        for (int rep = 0; rep < 2; ++rep)
            for (int opts : {8, 7, 7 + 16, 7 + 32, 7 + 48}) {
                roll_probe_kernel<<<148, 192, 200 * 1024>>>(opts, steps, gsrc, d_out, sink);
                CK(cudaDeviceSynchronize());
                std::vector<long long> t(148);
                CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0, mn = 1ll << 60; double mean = 0;
                for (long long v : t) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; mean += (double)v / 148; }
                printf("roll opts %2d (%s%s%s%s%s%s): clocks per step min / mean / max %7.0f / %7.0f / %7.0f  (27 MMAs of N = 144: floor 1944)\n", opts,
                       (opts & 8) ? "fixed window " : "moving window ", (opts & 1) ? "+commits " : "", (opts & 2) ? "+tcgen05.ld/st traffic " : "",
                       (opts & 4) ? "+bulk copies into operand slots " : "", (opts & 16) ? "+fence::after_thread_sync per 9 " : "", (opts & 32) ? "+mbarrier wait per 9" : "", (double)mn / steps, mean / steps, (double)mx / steps);
            }
        if (argc > 1) return 0;
    }
    // ------------------------------------------------ semantics ------------------------------------------------------
    // A image: logical rows R = 0..143 of 16 fp16 (32 bytes), value (R * 16 + k) & 2047; row pitch X in {32, 64, 128} bytes with
    // the X-byte swizzle (the k-step occupies bytes [koff, koff + 32) of the row), or the no-swizzle chunk-plane layout.
    struct Mode { const char *name; int x; uint32_t layout; };
    const Mode modes[] = {{"none(planes)", 16, 0}, {"swizzle32", 32, 6}, {"swizzle64", 64, 4}, {"swizzle128", 128, 2}};
    std::vector<uint8_t> bimg(16 * 32, 0);       // B identity, no swizzle: [2 chunks][16 rows n][8 k] -> lbo = 256, sbo = 128
    for (int n = 0; n < 16; ++n) {
        __half one = __float2half(1.0f);
        memcpy(&bimg[((n / 8) * 16 + n) * 16 + (n % 8) * 2], &one, 2);   // chunk n/8 (k 8..15 in chunk 1), row n, element n%8
    }
    uint8_t *d_a, *d_b; float *d_d;
    CK(cudaMalloc(&d_a, 64 * 1024)); CK(cudaMalloc(&d_b, 4096)); CK(cudaMalloc(&d_d, 128 * 16 * 4));
    CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    const int ROWS = 144;
    for (const Mode &m : modes) {
        for (int koff = 0; koff < (m.x >= 32 ? m.x : 32); koff += 32) {
            if (m.x == 16 && koff) break;
            for (int shift = 0; shift <= 9; shift += (shift < 2 ? 1 : 7)) {      // 0, 1, 2, 9 rows
                for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                    std::vector<uint8_t> img(64 * 1024, 0);
                    SemCfg c{};
                    if (m.x == 16) {
                        // [2 chunks][ROWS][16 B]
                        for (int R = 0; R < ROWS; ++R)
                            for (int k = 0; k < 16; ++k) {
                                __half v = __float2half((float)((R * 16 + k) & 2047));
                                memcpy(&img[((k / 8) * ROWS + R) * 16 + (k % 8) * 2], &v, 2);
                            }
                        c.layout_a = 0; c.lbo_a = ROWS * 16; c.sbo_a = 128; c.start_off_a = shift * 16;
                    } else {
                        for (int R = 0; R < ROWS; ++R)
                            for (int k = 0; k < 16; ++k) {
                                __half v = __float2half((float)((R * 16 + k) & 2047));
                                const uint32_t p = swz((uint32_t)(R * m.x + koff + k * 2), m.x);
                                memcpy(&img[p], &v, 2);
                            }
                        c.layout_a = m.layout; c.lbo_a = 16; c.sbo_a = 8 * m.x; c.start_off_a = shift * m.x + koff;
                    }
                    c.base_off_a = bo_mode ? ((c.start_off_a >> 7) & 7) : 0;
                    if (bo_mode && c.base_off_a == 0) continue;
                    c.layout_b = 0; c.lbo_b = 256; c.sbo_b = 128;
                    c.a_bytes = 64 * 1024; c.b_bytes = 512;
                    CK(cudaMemcpy(d_a, img.data(), img.size(), cudaMemcpyHostToDevice));
                    CK(cudaMemset(d_d, 0xff, 128 * 16 * 4));
                    sem_kernel<<<1, 128, 128 * 1024>>>(c, d_a, d_b, d_d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("sem %s koff %d shift %d: CUDA error %s\n", m.name, koff, shift, cudaGetErrorString(e)); return 1; }
                    std::vector<float> d(128 * 16);
                    CK(cudaMemcpy(d.data(), d_d, d.size() * 4, cudaMemcpyDeviceToHost));
                    int bad = 0, first = -1;
                    for (int r = 0; r < 128; ++r)
                        for (int n = 0; n < 16; ++n) {
                            const float want = (float)(((r + shift) * 16 + n) & 2047);
                            if (d[r * 16 + n] != want) { if (first < 0) first = r * 16 + n; ++bad; }
                        }
                    printf("sem %-13s koff %3d shift %d base_off %d : %s (%d wrong", m.name, koff, shift, c.base_off_a, bad ? "MISMATCH" : "ok", bad);
                    if (bad) printf("; first at r=%d n=%d got %.0f want %.0f", first / 16, first % 16, d[first], (float)((((first / 16) + shift) * 16 + first % 16) & 2047));
                    printf(")\n");
                }
            }
        }
    }
    // ------------------------------------------------- timing --------------------------------------------------------
    struct L { const char *name; uint32_t layout, lbo, sbo, slot; };
    // A operand layouts (128 rows used; slot stride chosen >= the image size, 1024-aligned)
    const L la[] = {
        {"none planes[chunk][132 rows][16B]", 0, 132 * 16, 128, 13312},
        {"none atoms[row/8][chunk][8][16B]", 0, 128, 256, 13312},
        {"swizzle32 [rows][32B]", 6, 16, 256, 13312},
        {"swizzle64 [rows][64B]", 4, 16, 512, 13312},
        {"swizzle128 [rows][128B]", 2, 16, 1024, 17408},
    };
    const int Ns[] = {16, 48, 96, 144, 192, 224};
    for (int grid : {1, 148}) {
        for (const L &l : la) {
            for (int N : Ns) {
                TimingCfg c{};
                c.layout_a = l.layout; c.lbo_a = l.lbo; c.sbo_a = l.sbo; c.slot_stride_a = l.slot;
                // B in the same family of layout ([N rows] instead of 128)
                c.layout_b = l.layout; c.sbo_b = l.sbo;
                c.lbo_b = (l.layout == 0 && l.lbo != 128) ? (uint32_t)N * 16 : l.lbo;
                c.tile_stride_b = 9216;    // >= 256 rows x 32 B, 1024-aligned
                c.N = N; c.iters = 4096;
                timing_kernel<<<grid, 128, 200 * 1024>>>(c, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("timing %s N %d: CUDA error %s\n", l.name, N, cudaGetErrorString(e)); return 1; }
                std::vector<long long> t(grid);
                CK(cudaMemcpy(t.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
                printf("time grid %3d  A/B %-36s N %3d : %7.1f clk per MMA (floor %5.1f)\n", grid, l.name, N, (double)mx / c.iters, 128.0 * N / 256.0);
            }
        }
    }
    // mixed: A in each layout, B always no-swizzle planes (what the convolution's resident weights use today)
    for (const L &l : la) {
        TimingCfg c{};
        c.layout_a = l.layout; c.lbo_a = l.lbo; c.sbo_a = l.sbo; c.slot_stride_a = l.slot;
        c.layout_b = 0; c.lbo_b = 144 * 16; c.sbo_b = 128; c.tile_stride_b = 4608;
        c.N = 144; c.iters = 4096;
        timing_kernel<<<148, 128, 200 * 1024>>>(c, d_out);
        CK(cudaDeviceSynchronize());
        std::vector<long long> t(148);
        CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
        printf("time grid 148  A %-36s B none planes, N 144 : %7.1f clk per MMA\n", l.name, (double)mx / c.iters);
    }
    // lean issue loop (SS) and A-from-TMEM (TS), no-swizzle planes
    for (int ts = 0; ts < 2; ++ts)
        for (int N : {16, 48, 96, 144, 192, 224}) {
            TimingCfg c{};
            c.layout_a = 0; c.lbo_a = 132 * 16; c.sbo_a = 128; c.slot_stride_a = 13312;
            c.layout_b = 0; c.lbo_b = (uint32_t)N * 16; c.sbo_b = 128; c.tile_stride_b = 9216;
            c.N = N; c.iters = 4096;
            timing_lean_kernel<<<148, 128, 200 * 1024>>>(c, ts, d_out);
            CK(cudaDeviceSynchronize());
            std::vector<long long> t(148);
            CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
            printf("lean grid 148 %s N %3d : %7.1f clk per MMA (floor %5.1f)\n", ts ? "A from TMEM" : "A from smem", N, (double)mx / c.iters, 128.0 * N / 256.0);
        }
    // pacing study
    for (int N : {96, 144, 192, 256})
        for (int cfg = 0; cfg < 10; ++cfg) {
            const int na_[]   = {4, 1, 4, 4, 1, 4, 4, 4, 4, 2};
            const int nb_[]   = {4, 4, 1, 4, 1, 4, 4, 4, 4, 2};
            const int nacc_[] = {2, 2, 2, 1, 1, 2, 2, 2, 2, 2};
            const int gap_[]  = {0, 0, 0, 0, 0, 64, 96, 128, 160, 0};
            TimingCfg c{};
            c.layout_a = 0; c.lbo_a = 132 * 16; c.sbo_a = 128; c.slot_stride_a = 13312;
            c.layout_b = 0; c.lbo_b = (uint32_t)N * 16; c.sbo_b = 128; c.tile_stride_b = 9216;
            c.N = N; c.iters = 4096;
            timing_pace_kernel<<<148, 128, 200 * 1024>>>(c, na_[cfg], nb_[cfg], nacc_[cfg], gap_[cfg], d_out);
            CK(cudaDeviceSynchronize());
            std::vector<long long> t(148);
            CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
            printf("pace N %3d  A slots %d  B tiles %d  accumulators %d  gap %3d : %7.1f clk per MMA (floor %5.1f)\n", N, na_[cfg], nb_[cfg],
                   nacc_[cfg], gap_[cfg], (double)mx / c.iters, 128.0 * N / 256.0);
        }
    // operand start-address alignment study (the convolution's dx taps start 1 row = 16 B off the 128-byte core-matrix grid)
    for (int lay = 0; lay < 3; ++lay)
        for (int N : {48, 96, 144})
            for (int cfg = 0; cfg < 7; ++cfg) {
                const uint32_t layout[] = {0, 6, 2}, rowb[] = {16, 32, 128}, sbo[] = {128, 256, 1024};
                const char *lname[] = {"none planes", "swizzle32", "swizzle128"};
                //                 aligned  A+1row  A+2rows  cycle dx  B+48rows  D+48cols  all
                const int aoff[] = {0, 1, 2, 0, 0, 0, 0};
                const int acyc[] = {0, 0, 0, 1, 0, 0, 1};
                const int boff[] = {0, 0, 0, 0, 48, 0, 48};
                const int doff[] = {0, 0, 0, 0, 0, 48, 48};
                TimingCfg c{};
                c.layout_a = layout[lay]; c.lbo_a = lay == 0 ? 132 * 16 : 16; c.sbo_a = sbo[lay]; c.slot_stride_a = lay == 2 ? 17408 : 13312;
                c.layout_b = 0; c.lbo_b = 192 * 16; c.sbo_b = 128; c.tile_stride_b = 9216;     // B: no-swizzle planes of 192 rows
                c.N = N; c.iters = 4096;
                timing_pace_kernel<<<148, 128, 200 * 1024>>>(c, 4, 4, 2, 0, d_out, aoff[cfg] * rowb[lay], boff[cfg] * 16, doff[cfg], acyc[cfg] * rowb[lay]);
                CK(cudaDeviceSynchronize());
                std::vector<long long> t(148);
                CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
                printf("align A %-11s N %3d  A start +%d rows%s  B start +%2d rows  D +%2d cols : %7.1f clk per MMA (floor %5.1f)\n", lname[lay], N,
                       aoff[cfg], acyc[cfg] ? " (cycling 0/1/2)" : "", boff[cfg], doff[cfg], (double)mx / c.iters, 128.0 * N / 256.0);
            }
    // interference study
    {
        uint4 *gsrc; unsigned *sink;
        CK(cudaMalloc(&gsrc, (size_t)148 * 4096 * 16)); CK(cudaMemset(gsrc, 0, (size_t)148 * 4096 * 16)); CK(cudaMalloc(&sink, 4));
        for (int N : {48, 144})
            for (int cfg = 0; cfg < 13; ++cfg) {
                const int ce[] = {0, 9, 3, 1, 0, 0, 0, 0, 0, 0, 0, 0, 9};
                const int bw[] = {0, 0, 0, 0, 1, 4, 8, 1, 4, 8, 4, 8, 4};
                const int bm[] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3};
                TimingCfg c{};
                c.layout_a = 0; c.lbo_a = 132 * 16; c.sbo_a = 128; c.slot_stride_a = 13312;
                c.layout_b = 0; c.lbo_b = (uint32_t)N * 16; c.sbo_b = 128; c.tile_stride_b = 9216;
                c.N = N; c.iters = 4096;
                timing_interf_kernel<<<148, 288, 200 * 1024>>>(c, ce[cfg], bw[cfg], bm[cfg], gsrc, d_out, sink);
                CK(cudaDeviceSynchronize());
                std::vector<long long> t(148);
                CK(cudaMemcpy(t.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0; for (long long v : t) mx = v > mx ? v : mx;
                const char *mn[] = {"-", "st.shared.v4", "ld.shared.v4", "cp.async 16B"};
                printf("interf N %3d  commit every %d  background warps %d (%s) : %7.1f clk per MMA (floor %5.1f)\n", N, ce[cfg], bw[cfg], mn[bm[cfg]],
                       (double)mx / c.iters, 128.0 * N / 256.0);
            }
    }
    return 0;
}
