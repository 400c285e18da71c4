// Blackwell (sm_100a) tensor-core plumbing: mbarrier, bulk async copies (TMA engine, UBLKCP), tcgen05 / TMEM.
// Inline PTX only; descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spins on try_wait.  Watchdog: a wait that lasts longer than ~4 s of wall clock (a protocol bug - no kernel of this library runs
// that long) traps, so a deadlock surfaces as a CUDA error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, polls = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && (++polls & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    } while (!done);
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar` (complete_tx bytes)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------------------------------
// one full warp: allocates `ncols` (power of two >= 32) columns, writes the base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// all previously issued tcgen05.mma of this thread arrive (once) on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- warp-uniform issue ------------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit are warp-level ("uniform datapath") instructions.  Issued from inside an `if (lane == 0)` block
// the compiler keeps their operands in per-thread registers and wraps EVERY instruction in an ELECT loop with ~7
// R2UR.BROADCAST moves: ~20 dependent instructions = ~140 clk per MMA from the single issuing thread, above the 72-clk tensor
// floor of a 128 x 144 x 16 MMA (scripts/mma_probe.cu measures max(66, N / 2) clk per MMA for every operand layout when the
// descriptors already sit in uniform registers).  The helpers below are called by ALL 32 lanes of a converged warp with
// warp-uniform arguments; one elected lane issues.  Descriptors travel as (lo, hi) words so that advancing the 14-bit address
// field is one 32-bit add.
__device__ __forceinline__ uint32_t warp_idx_uniform() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e != 0;
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity) {   // all lanes call; one polls
    if (elect_one()) mbar_wait(bar, parity);
    __syncwarp();
}
// Issuer flavour: one try_wait by every lane (no election, no loop) when the barrier has usually completed already; the polling
// path only when it has not.  The MMA queue is shallow, so every instruction the issuing warp spends between two MMAs is
// tensor-pipe idle time.
__device__ __forceinline__ void mbar_wait_lean(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (__any_sync(0xffffffffu, done == 0)) mbar_wait_warp(bar, parity);
}
// Split form for software pipelining: mbar_peek issues the try_wait (every lane) and returns its raw result; mbar_wait_peeked
// completes the wait later - a vote, and the polling path only if the barrier had not completed when peeked.  The issuer peeks at
// the NEXT operand slot's barrier before issuing the current group of MMAs, so the poll's latency runs under queued MMAs.
__device__ __forceinline__ uint32_t mbar_peek(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait_peeked(uint32_t peeked, uint64_t *bar, uint32_t parity) {
    if (__any_sync(0xffffffffu, peeked == 0)) mbar_wait_warp(bar, parity);
}
// The same for roles with slack (producers waiting for a free slot, epilogues waiting for a result): the polling lane sleeps
// between tries.  128 threads spinning on try_wait compete with the tensor core's operand fetch for the shared-memory pipe; with
// every thread of the four epilogue warps polling, a few CTAs per launch of the rolling-row convolution fell into a slow mode
// (2.8x the clocks of their neighbours, MMAs completing every ~200 clk) and set the kernel's duration.
__device__ __forceinline__ void mbar_wait_warp_relaxed(uint64_t *bar, uint32_t parity) {
    if (elect_one()) {
        const uint32_t addr = smem_u32(bar);
        uint32_t done, polls = 0;
        unsigned long long t0 = 0;
        for (;;) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity)
                : "memory");
            if (done) break;
            __nanosleep(64);
            if ((++polls & 0x3ffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) __trap();
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }   // version 1, no swizzle
__device__ __forceinline__ void mma_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], warp-uniform issue
__device__ __forceinline__ void mma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" core matrices of 8 rows x 16 bytes, 128 B each).
//   K-major operand  [rows x K]:  lbo = byte distance between the two 16-byte K chunks of one K=16 step,
//                                 sbo = byte distance between consecutive 8-row groups.
//   MN-major operand [K x cols]:  lbo = byte distance between consecutive groups of 8 K,
//                                 sbo = byte distance between consecutive groups of 8 columns (16 bytes).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);             // [0,14)  start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // [16,30) leading-dimension byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // [32,46) stride-dimension byte offset >> 4
    d |= (uint64_t)1 << 46;                              // [46,48) descriptor version 1 (sm_100)
    return d;                                            // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

// Instruction descriptor for kind::f16 with bf16 A/B, fp32 accumulate, A K-major.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, bool b_mn_major) {
    return (1u << 4)                          // [4,6)   D format  = F32
           | (1u << 7)                        // [7,10)  A format  = BF16
           | (1u << 10)                       // [10,13) B format  = BF16
           | ((b_mn_major ? 1u : 0u) << 16)   // [16]    B major   (0 = K, 1 = MN)
           | ((uint32_t)(N >> 3) << 17)       // [17,23) N >> 3
           | ((uint32_t)(M >> 4) << 24);      // [24,29) M >> 4
}

// D[tmem] (+)= A[smem] * B[smem]     (issued by ONE thread)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// TMEM <-> registers, shape 32x32b: lane i of the warp touches TMEM lane (base lane + i), consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
    return r;
}
// 16-bit operand formats of kind::f16: bf16 (F16 = false) or fp16 (F16 = true: 10-bit mantissa at the same tensor rate)
template <bool F16> __device__ __forceinline__ uint32_t pack_h16(float lo, float hi) {
    uint32_t r;
    if constexpr (F16)
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// per-channel-pair statistics straight from a packed 16-bit pair: s += v, q += v * v (the halves are widened inside the
// fp32 add / FMA: FHADD / FHFMA, no unpack instructions)
template <bool F16> __device__ __forceinline__ void stat_h16x2(uint32_t w, float &s0, float &s1, float &q0, float &q1) {
    if constexpr (F16)
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %4;\n\t"
            "add.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t"
            "fma.rn.f32.f16 %2, lo, lo, %2;\n\tfma.rn.f32.f16 %3, hi, hi, %3;\n\t}"
            : "+f"(s0), "+f"(s1), "+f"(q0), "+f"(q1) : "r"(w));
    else
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %4;\n\t"
            "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t"
            "fma.rn.f32.bf16 %2, lo, lo, %2;\n\tfma.rn.f32.bf16 %3, hi, hi, %3;\n\t}"
            : "+f"(s0), "+f"(s1), "+f"(q0), "+f"(q1) : "r"(w));
}
template <bool F16> __device__ __forceinline__ float2 unpack_h16(uint32_t w) {
    if constexpr (F16) {
        float a, b;
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(a), "=f"(b) : "r"(w));
        return make_float2(a, b);
    } else {
        return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
}
template <bool F16> __host__ __device__ constexpr uint32_t instr_desc_h16(int M, int N, bool b_mn_major) {
    // A / B format fields [7,10) and [10,13): 0 = F16, 1 = BF16
    return F16 ? (instr_desc_bf16(M, N, b_mn_major) & ~((7u << 7) | (7u << 10))) : instr_desc_bf16(M, N, b_mn_major);
}

}  // namespace tc
}  // namespace wf
