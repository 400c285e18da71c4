// Kernel group 2, bf16 tensor-core path (tcgen05 + TMEM, sm_100a): window partition + QKV projection, attention core,
// output projection, for windows of 8^3 = 512 tokens and head_dim 16 (every stage of the reference configuration).
//
// Replaces Block.window_partition (reference network_models/wave_helper.py:450-461), Attention.forward
// (network_models/attention.py:83-104) and the reshape-only window reverse (wave_helper.py:498-499).
//
// Three launches:
//   1. linear_tc_kernel<QKV>   D[128 x NT] = X_tile[128 x C] * Wqkv_tile[NT x C]^T  (tcgen05.mma, fp32 accum in TMEM);
//      the window-partition gather is folded into the A-tile load; the epilogue adds the bias, folds scale*log2(e)
//      into q and writes q/k/v per (window, head) in the 16-byte-chunked image [2][512][8] that IS the canonical
//      no-swizzle UMMA operand layout, so the core kernel fetches whole operands with one bulk copy each.
//   2. attn_core_tc_kernel     per (head, 128-query tile), persistent over windows: S = Q K^T (2 x MMA 128x256x16) into
//      all 512 TMEM columns; 8 softmax warps (one TMEM lane = one query row; two warps per row split the key halves)
//      do an exact two-pass softmax against a bf16 relative-position-bias tile that stays resident in shared memory
//      across windows; P (bf16) overwrites S in place in TMEM and feeds O = P V as the TMEM A operand (32 x MMA
//      128x16x16); Q/K/V of the next window are prefetched by bulk copies behind an mbarrier.
//   3. linear_tc_kernel<PROJ>  out = O * Wproj^T + b, written in window order (= the reference's reshape-only reverse).
//
// With head_dim 16 the core is bound by the 512x512 exponentials per (window, head), not by the tensor pipe: per
// 128x512 tile the MMAs need ~512 cycles while 65536 ex2 need >= 4096 cycles of the SM's 16/clk MUFU pipe.
#include <type_traits>

#include "tc_common.cuh"
#include "wf_common.cuh"

namespace wf {

using namespace tc;

struct TcWindowMap {
    int D1, H1, W1, nWy, nWx, nW;  // ws = 8, N = 512
    __device__ inline int64_t voxel(int64_t m) const {
        const int tok = (int)(m & 511);
        const int64_t win = m >> 9;
        const int widx = (int)(win % nW);
        const int64_t b = win / nW;
        const int xb = widx % nWx, yb = (widx / nWx) % nWy, zb = widx / (nWx * nWy);
        const int dx = tok & 7, dy = (tok >> 3) & 7, dz = tok >> 6;
        return ((b * D1 + zb * 8 + dz) * H1 + yb * 8 + dy) * (int64_t)W1 + xb * 8 + dx;
    }
};

constexpr float kLog2e = 1.4426950408889634f;

// ---- 16-bit operand formats: bf16 (FMT16 = false) or fp16 (FMT16 = true; 10-bit mantissa, used when the caller wants
// tighter numerics than bf16 at the same tensor-core rate) --------------------------------------------------------
template <bool F16> __device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    uint32_t r;
    if constexpr (F16)
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// t0 = lo16(w) + s0, t1 = hi16(w) + s1 with the 16-bit halves widened inside the add (FHADD: no unpack instruction)
template <bool F16> __device__ __forceinline__ void add16x2(uint32_t w, float s0, float s1, float &t0, float &t1) {
    if constexpr (F16)
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %3;\n\tadd.rn.f32.f16 %1, hi, %4;\n\t}"
            : "=f"(t0), "=f"(t1) : "r"(w), "f"(s0), "f"(s1));
    else
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %3;\n\tadd.rn.f32.bf16 %1, hi, %4;\n\t}"
            : "=f"(t0), "=f"(t1) : "r"(w), "f"(s0), "f"(s1));
}
template <bool F16> __device__ __forceinline__ float widen16(uint16_t h) {
    float t;
    if constexpr (F16)
        asm("cvt.f32.f16 %0, %1;" : "=f"(t) : "h"(h));
    else
        t = __uint_as_float((uint32_t)h << 16);
    return t;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
template <bool F16> __host__ __device__ constexpr uint32_t instr_desc16(int M, int N, bool b_mn_major) {
    return F16 ? (instr_desc_bf16(M, N, b_mn_major) & ~((7u << 7) | (7u << 10))) : instr_desc_bf16(M, N, b_mn_major);
}

// ------------------------------------------------------------------------------------------------ projections ----
// grid (M/128, Nout/NT), 128 threads.  smem: A image [kpc][128][8], B image [kpc][NT][8] 16-bit (no-swizzle canonical),
// kpc = K chunks (of 8 elements) staged per pass; K is walked in K / (8 kpc) passes that accumulate into the same TMEM tile.
// TA: float (converted to the operand format while staging), __nv_bfloat16 (converted unless the format is bf16), or
//     uint16_t (already in the operand format: the core kernel's output).   TO: float or uint16_t (operand format).
// SPLIT (fp16 only): error-compensated operands.  Every operand is carried as hi + lo, hi = fp16(v), lo = fp16(v - hi)
//     (22 significant bits together), and the product is accumulated as A_hi B_hi + A_lo B_hi + A_hi B_lo - three
//     tcgen05.mma per k-step instead of one, on a tensor pipe that idles anyway at these sizes (K = 48 .. 384).  What the
//     softmax exponentiates is then accurate to fp32 level: with plain fp16 operands the |q||k| 2^-11 error of the scores
//     is the largest single contribution to the 16-bit policy's logit error (scripts/precision_policy2_r02.py).  The QKV
//     epilogue stores q and k as hi / lo pairs too (planes 0 / 1 hi, 3 / 4 lo; v in plane 2), bias is added in fp32.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    const float ra = a - hf.x, rb = b - hf.y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

template <bool QKV, typename TA, bool F16, typename TO, bool SPLIT>
__global__ void __launch_bounds__(128) linear_tc_kernel(const TA *__restrict__ A, const TA *__restrict__ A_lo,
                                                        const uint16_t *__restrict__ Wt, const uint16_t *__restrict__ Wt_lo,
                                                        const uint16_t *__restrict__ bias, const float *__restrict__ bias32,
                                                        TO *__restrict__ out, int K, int NT, int Nout, TcWindowMap map,
                                                        int heads, int64_t B_, float qscale, uint32_t tmem_cols, int kpc) {
    static_assert(!SPLIT || F16, "compensated operands are an fp16 feature");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kchunks = K >> 3;
    const int npass = kchunks / kpc;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)kpc * 2048 * (SPLIT ? 2 : 1);
    uint8_t *sAl = smem + (size_t)kpc * 2048;            // SPLIT only
    uint8_t *sBl = sB + (size_t)kpc * NT * 16;           // SPLIT only
    const int64_t m0 = (int64_t)blockIdx.x * 128;
    const int n0 = blockIdx.y * NT;

    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    const int64_t src_row = QKV ? map.voxel(m0 + tid) : (m0 + tid);
    uint32_t tmem = 0;
    for (int p = 0; p < npass; ++p) {
        const int kc0 = p * kpc;
        // A tile: row r = tid, every 16-byte K chunk.  A warp writes 32 consecutive rows of one chunk: conflict-free.
        // (pass p > 0: every thread has waited for the previous pass's MMAs below, so the images may be overwritten)
        // Chunks are handled four at a time with all their global loads issued before the first conversion: the
        // single-window stages (C = 192 / 384) launch a handful of CTAs whose time IS this loop's load latency.
        {
            constexpr bool raw = sizeof(TA) == 2 && (std::is_same<TA, uint16_t>::value || !F16);
            if constexpr (raw) {
                const uint4 *src = reinterpret_cast<const uint4 *>(A + src_row * K) + kc0;
                const uint4 *srl = SPLIT ? reinterpret_cast<const uint4 *>(A_lo + src_row * K) + kc0 : nullptr;
                for (int kc = 0; kc < kpc; kc += 4) {
                    uint4 h[4], l[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) {
                            h[u] = __ldg(src + kc + u);
                            if constexpr (SPLIT) l[u] = __ldg(srl + kc + u);
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) {
                            *reinterpret_cast<uint4 *>(sA + (size_t)(kc + u) * 2048 + tid * 16) = h[u];
                            if constexpr (SPLIT) *reinterpret_cast<uint4 *>(sAl + (size_t)(kc + u) * 2048 + tid * 16) = l[u];
                        }
                }
            } else if constexpr (sizeof(TA) == 2) {  // bf16 activations, fp16 operands
                const uint4 *src = reinterpret_cast<const uint4 *>(A + src_row * K) + kc0;
                for (int kc = 0; kc < kpc; kc += 4) {
                    uint4 h[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) h[u] = __ldg(src + kc + u);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) {
                            const uint32_t w[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
                            uint4 o;
                            uint32_t *ow = &o.x;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                ow[e] = pack16<F16>(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
                            *reinterpret_cast<uint4 *>(sA + (size_t)(kc + u) * 2048 + tid * 16) = o;
                        }
                }
            } else {  // fp32 activations (residual-stream precision): converted to the operand format while staging
                const float4 *src = reinterpret_cast<const float4 *>(A + src_row * K) + 2 * kc0;
                for (int kc = 0; kc < kpc; kc += 4) {
                    float4 va[4], vb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) {
                            va[u] = __ldg(src + 2 * (kc + u));
                            vb[u] = __ldg(src + 2 * (kc + u) + 1);
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kc + u < kpc) {
                            const float4 a = va[u], b = vb[u];
                            uint4 o;
                            if constexpr (SPLIT) {
                                uint4 l;
                                split_pair(a.x, a.y, o.x, l.x); split_pair(a.z, a.w, o.y, l.y);
                                split_pair(b.x, b.y, o.z, l.z); split_pair(b.z, b.w, o.w, l.w);
                                *reinterpret_cast<uint4 *>(sAl + (size_t)(kc + u) * 2048 + tid * 16) = l;
                            } else {
                                o.x = pack16<F16>(a.x, a.y); o.y = pack16<F16>(a.z, a.w);
                                o.z = pack16<F16>(b.x, b.y); o.w = pack16<F16>(b.z, b.w);
                            }
                            *reinterpret_cast<uint4 *>(sA + (size_t)(kc + u) * 2048 + tid * 16) = o;
                        }
                }
            }
        }
        // B tile: rows n0 .. n0+NT-1 of the [Nout, K] weight (four cells per thread and step, loads first)
        for (int base = tid; base < NT * kpc; base += 4 * 128) {
            uint4 h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 128;
                if (idx < NT * kpc) {
                    const int r = idx % NT, kc = idx / NT;
                    h[u] = __ldg(reinterpret_cast<const uint4 *>(Wt + (int64_t)(n0 + r) * K) + kc0 + kc);
                    if constexpr (SPLIT) l[u] = __ldg(reinterpret_cast<const uint4 *>(Wt_lo + (int64_t)(n0 + r) * K) + kc0 + kc);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 128;
                if (idx < NT * kpc) {
                    const int r = idx % NT, kc = idx / NT;
                    *reinterpret_cast<uint4 *>(sB + ((size_t)kc * NT + r) * 16) = h[u];
                    if constexpr (SPLIT) *reinterpret_cast<uint4 *>(sBl + ((size_t)kc * NT + r) * 16) = l[u];
                }
            }
        }
        fence_proxy_async();  // st.shared above -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem = tmem_slot;
        if (tid == 0) {
            const uint32_t idesc = instr_desc16<F16>(128, NT, false);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int ks = 0; ks < (kpc >> 1); ++ks) {
                const uint64_t da = smem_desc(a0 + ks * 2 * 2048, 2048, 128);
                const uint64_t db = smem_desc(b0 + ks * 2 * NT * 16, NT * 16, 128);
                mma_ss(tmem, da, db, idesc, (p > 0 || ks > 0) ? 1u : 0u);
                if constexpr (SPLIT) {
                    const uint64_t dal = smem_desc(smem_u32(sAl) + ks * 2 * 2048, 2048, 128);
                    const uint64_t dbl = smem_desc(smem_u32(sBl) + ks * 2 * NT * 16, NT * 16, 128);
                    mma_ss(tmem, dal, db, idesc, 1u);
                    mma_ss(tmem, da, dbl, idesc, 1u);
                }
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, p & 1);
        tc_fence_after();
    }
    // epilogue: thread = output row; 16 columns (= one head's q, k or v slice when QKV) per step
    const int64_t m = m0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < NT; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + c, r);
        tmem_wait_ld();
        const int n = n0 + c;
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
            v[e] = __uint_as_float(r[e]) + (bias32 ? __ldg(bias32 + n + e) : (bias ? widen16<F16>(bias[n + e]) : 0.f));
        if constexpr (QKV) {
            const int C = heads * 16;
            const int which = n / C, hh = (n % C) >> 4;
            if (which == 0) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] *= qscale;
            }
            const int64_t win = m >> 9;
            const int tok = (int)(m & 511);
            const int64_t plane = B_ * heads * 8192;
            uint16_t *dst = reinterpret_cast<uint16_t *>(out) + ((int64_t)win * heads + hh) * 8192 + tok * 8 + which * plane;
            uint4 lo, hi;
            if constexpr (SPLIT) {
                uint4 llo, lhi;   // the lo parts of head dims 0..7 / 8..15
                split_pair(v[0], v[1], lo.x, llo.x); split_pair(v[2], v[3], lo.y, llo.y);
                split_pair(v[4], v[5], lo.z, llo.z); split_pair(v[6], v[7], lo.w, llo.w);
                split_pair(v[8], v[9], hi.x, lhi.x); split_pair(v[10], v[11], hi.y, lhi.y);
                split_pair(v[12], v[13], hi.z, lhi.z); split_pair(v[14], v[15], hi.w, lhi.w);
                if (which < 2) {      // q and k: the operands of the scores
                    uint16_t *dl = dst + 3 * plane;      // plane 3 (q lo) / 4 (k lo)
                    *reinterpret_cast<uint4 *>(dl) = llo;
                    *reinterpret_cast<uint4 *>(dl + 4096) = lhi;
                }
            } else {
                lo.x = pack16<F16>(v[0], v[1]); lo.y = pack16<F16>(v[2], v[3]); lo.z = pack16<F16>(v[4], v[5]); lo.w = pack16<F16>(v[6], v[7]);
                hi.x = pack16<F16>(v[8], v[9]); hi.y = pack16<F16>(v[10], v[11]); hi.z = pack16<F16>(v[12], v[13]); hi.w = pack16<F16>(v[14], v[15]);
            }
            *reinterpret_cast<uint4 *>(dst) = lo;          // chunk 0: head dims 0..7
            *reinterpret_cast<uint4 *>(dst + 4096) = hi;   // chunk 1: head dims 8..15
        } else if constexpr (sizeof(TO) == 4) {
            float4 *dst = reinterpret_cast<float4 *>(out + m * Nout + n);
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        } else {
            uint4 lo, hi;
            lo.x = pack16<F16>(v[0], v[1]); lo.y = pack16<F16>(v[2], v[3]); lo.z = pack16<F16>(v[4], v[5]); lo.w = pack16<F16>(v[6], v[7]);
            hi.x = pack16<F16>(v[8], v[9]); hi.y = pack16<F16>(v[10], v[11]); hi.z = pack16<F16>(v[12], v[13]); hi.w = pack16<F16>(v[14], v[15]);
            uint4 *dst = reinterpret_cast<uint4 *>(out + m * Nout + n);
            dst[0] = lo;
            dst[1] = hi;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------ core ------
constexpr int kBiasPitch = 520;                       // 16-bit elements per bias row (512 + 8 pad: conflict-free LDS.128)
constexpr int kBiasBytes = 128 * kBiasPitch * 2;      // 133120 per (head, query tile)
constexpr int kStageBytes = 4096 + 16384 + 16384;     // Q tile [2][128][8] + K [2][512][8] + V [2][512][8]
constexpr int kCoreSmem = kBiasBytes + 2 * kStageBytes + 2 * 4 * 128 * 4;  // + row max / row sum exchange (4 key quarters)
// SPLIT (compensated scores): Q hi / lo tiles + K hi / lo in ONE buffer that is refilled for the next window as soon as this
// window's score MMAs have completed (the whole softmax phase hides the copy), V double-buffered as before
constexpr int kQKSplitBytes = 2 * 4096 + 2 * 16384;                        // 40960
constexpr int kCoreSmemSplit = kBiasBytes + kQKSplitBytes + 2 * 16384 + 2 * 4 * 128 * 4;   // 210944

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Dense bias image for the core kernel: img[h][qt][i][kBiasPitch] = fmt16(table[index[qt*128+i][j]][h] * log2 e).
// Built once per table version; the core kernel then fetches its (head, query tile) slab with ONE bulk copy.
template <bool F16>
__global__ void relpos_bias_image_kernel(const void *__restrict__ table, int table_dtype,
                                         const int64_t *__restrict__ index, uint16_t *__restrict__ img, int heads,
                                         int table_rows) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over heads * 512 * kBiasPitch/2 (pairs)
    const int64_t total = (int64_t)heads * 512 * (kBiasPitch / 2);
    if (idx >= total) return;
    const int jp = (int)(idx % (kBiasPitch / 2));
    const int64_t r = idx / (kBiasPitch / 2);
    const int i = (int)(r % 512), h = (int)(r / 512);
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int j = 2 * jp + e;
        if (j < 512) {
            int64_t row = index[(int64_t)i * 512 + j];
            row = row < 0 ? 0 : (row >= table_rows ? table_rows - 1 : row);
            v[e] = (table_dtype == WF_F32 ? reinterpret_cast<const float *>(table)[row * heads + h]
                                          : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(table)[row * heads + h])) * kLog2e;
        }
    }
    reinterpret_cast<uint32_t *>(img)[idx] = pack16<F16>(v[0], v[1]);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void softmax_warps_sync() {   // named barrier 1: the 16 softmax warps only
    asm volatile("bar.sync 1, 512;" ::: "memory");
}

// grid = combos * groups, combos = heads * 4 (head, 128-query tile); CTA (combo, g) walks windows g, g+groups, ...
// 544 threads = 16 softmax warps + 1 issuer warp (warp specialisation):
//   softmax warp w: TMEM lanes 32*(w%4).. (query rows), key quarter w/4 (128 keys); pass 1 row max, pass 2 exponentials;
//   every finished 32-key chunk of P (16 packed TMEM columns, written over the first half of its own S chunk) is
//   announced on an mbarrier (one arrive per warp, four warps per chunk);
//   issuer (one thread): bulk loads of the next window, S = Q K^T (2 x tcgen05.mma 128x256x16), and - as each P chunk is
//   announced - its two O += P V k-steps (tcgen05.mma 128x16x16, A from TMEM), so the 32 small PV instructions run
//   underneath the exponentials instead of after them (they were 24 % of the kernel as a serial phase, profiles/).
// TMEM columns: S = [0, 512) fp32; P chunk (q, c) = [128q + 32c, 128q + 32c + 16); O = [16, 32) (the second half of
// chunk (0, 0), free once that chunk's P is written - the first PV instruction waits for exactly that chunk).
// SPLIT: q and k arrive as hi / lo pairs (planes 0 / 1 and 3 / 4 of `qkv`) and the scores are accumulated as
// Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T (6 instead of 2 tcgen05.mma per tile: the tensor pipe is ~5 % busy here), so the
// exponent's argument is exact to ~2^-21 |q||k|; O is written as a hi / lo pair (second plane at o + B_ * 512 * C) for the
// compensated output projection.
template <bool F16, bool SPLIT>
__global__ void __launch_bounds__(544, 1) attn_core_tc_kernel(const uint16_t *__restrict__ qkv,
                                                              const uint16_t *__restrict__ bias_img,
                                                              uint16_t *__restrict__ o, int heads, int64_t B_,
                                                              int groups) {
    static_assert(!SPLIT || F16, "compensated operands are an fp16 feature");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[2];
    __shared__ __align__(8) uint64_t bar_s, bar_o, bar_bias, bar_epi, bar_qk;
    __shared__ __align__(8) uint64_t bar_p[16];   // [quarter][chunk]
    __shared__ uint32_t tmem_slot;
    uint16_t *sBias = reinterpret_cast<uint16_t *>(smem);
    uint8_t *sStage = smem + kBiasBytes;
    float *sMax = reinterpret_cast<float *>(smem + kBiasBytes + (SPLIT ? kQKSplitBytes + 2 * 16384 : 2 * kStageBytes));  // [4][128]
    float *sSum = sMax + 512;                                                       // [4][128]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int combo = blockIdx.x % (heads * 4), g = blockIdx.x / (heads * 4);
    const int hh = combo >> 2, qt = combo & 3;
    const int C = heads * 16;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar_full[0], 1);
        mbar_init(&bar_full[1], 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_o, 1);
        mbar_init(&bar_bias, 1);
        mbar_init(&bar_epi, 4);
        mbar_init(&bar_qk, 1);
        for (int i = 0; i < 16; ++i) mbar_init(&bar_p[i], 4);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int64_t per_which = B_ * heads * 8192;  // elements between the q, k and v planes

    if (warp == 16) {
        // ================================================= issuer ====================================================
        if (g < B_) {      // the whole (converged) warp runs the issue loop, one elected lane issues each instruction
            // relative-position bias slab of this (head, query tile): one 130 KB bulk copy, resident for every window
            if (lane == 0) {
                mbar_expect_tx(&bar_bias, kBiasBytes);
                bulk_g2s(sBias, bias_img + ((int64_t)hh * 4 + qt) * (kBiasBytes / 2), kBiasBytes, &bar_bias);
            }
            __syncwarp();
            const uint32_t idesc_s = instr_desc16<F16>(128, 256, false);
            const uint32_t idesc_o = instr_desc16<F16>(128, 16, true);
            if constexpr (!SPLIT) {
                // (warp-uniform issue like the compensated branch below: all 32 lanes run the loop, one elected lane issues)
                auto issue_loads = [&](int64_t win, int stage) {
                    if (elect_one()) {
                        uint8_t *dst = sStage + stage * kStageBytes;
                        const uint16_t *qb = qkv + (win * heads + hh) * 8192;
                        mbar_expect_tx(&bar_full[stage], kStageBytes);
                        bulk_g2s(dst, qb + qt * 128 * 8, 2048, &bar_full[stage]);                 // Q chunk 0 (head dims 0..7)
                        bulk_g2s(dst + 2048, qb + 4096 + qt * 128 * 8, 2048, &bar_full[stage]);   // Q chunk 1
                        bulk_g2s(dst + 4096, qb + per_which, 16384, &bar_full[stage]);            // K
                        bulk_g2s(dst + 4096 + 16384, qb + 2 * per_which, 16384, &bar_full[stage]);  // V
                    }
                    __syncwarp();
                };
                issue_loads(g, 0);
                const uint32_t qk_hi = smem_desc_hi(128), v_hi = smem_desc_hi(8192);
                int it = 0;
                for (int64_t win = g; win < B_; win += groups, ++it) {
                    const int stage = it & 1;
                    const uint32_t ph_full = (it >> 1) & 1, ph = it & 1;
                    const uint32_t sQ = smem_u32(sStage + stage * kStageBytes), sK = sQ + 4096, sV = sK + 16384;
                    if (it > 0) {   // previous tile: O read back and its V / P consumed -> TMEM and the other stage are free
                        mbar_wait_warp(&bar_epi, (it - 1) & 1);
                        tc_fence_after();
                    }
                    if (win + groups < B_) issue_loads(win + groups, stage ^ 1);
                    mbar_wait_warp(&bar_full[stage], ph_full);
                    tc_fence_after();
                    // S[128 x 512] = Q[128 x 16] K^T : two N = 256 halves, TMEM columns [0,256) and [256,512)
                    const uint32_t dq = smem_desc_lo(sQ, 2048);
                    mma_ss_w(tmem, dq, qk_hi, smem_desc_lo(sK, 8192), qk_hi, idesc_s, 0u);
                    mma_ss_w(tmem + 256, dq, qk_hi, smem_desc_lo(sK + 256 * 16, 8192), qk_hi, idesc_s, 0u);
                    mma_commit_w(&bar_s);
                    // O[128 x 16] += P_chunk[128 x 32] V_chunk[32 x 16] as the chunks arrive (chunk index fastest over quarters)
                    const uint32_t dv = smem_desc_lo(sV, 128);
                    uint32_t first = 1;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c)
#pragma unroll 1
                        for (int q = 0; q < 4; ++q) {
                            mbar_wait_lean(&bar_p[q * 4 + c], ph);
                            tc_fence_after();
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int ks = q * 8 + c * 2 + j;   // k-step = keys [16 ks, 16 ks + 16)
                                mma_ts_w(tmem + 16, tmem + q * 128 + c * 32 + j * 8, dv + (uint32_t)(ks * 256 / 16), v_hi, idesc_o, first ? 0u : 1u);
                                first = 0;
                            }
                        }
                    mma_commit_w(&bar_o);
                }
            } else {
                uint8_t *sQK = sStage;                      // [Q hi 4096][Q lo 4096][K hi 16384][K lo 16384]
                uint8_t *sVb = sStage + kQKSplitBytes;      // 2 x 16384
                // Warp-uniform issue (tc_common.cuh): stage clocks of the single-thread issuer (profiles/r02_attn_stage_clocks_before.txt)
                // showed the 32 small PV instructions lagging ~2 400 clk behind the softmax warps at the end of every tile - every
                // tcgen05.mma issued from inside `if (lane == 0)` is wrapped in an ELECT loop with ~7 R2UR moves (~140 clk apiece).
                auto issue_qk = [&](int64_t win) {
                    if (elect_one()) {
                        const uint16_t *qb = qkv + (win * heads + hh) * 8192;
                        mbar_expect_tx(&bar_qk, kQKSplitBytes);
                        bulk_g2s(sQK, qb + qt * 128 * 8, 2048, &bar_qk);                                   // Q hi, head dims 0..7
                        bulk_g2s(sQK + 2048, qb + 4096 + qt * 128 * 8, 2048, &bar_qk);                     // Q hi, head dims 8..15
                        bulk_g2s(sQK + 4096, qb + 3 * per_which + qt * 128 * 8, 2048, &bar_qk);            // Q lo
                        bulk_g2s(sQK + 6144, qb + 3 * per_which + 4096 + qt * 128 * 8, 2048, &bar_qk);
                        bulk_g2s(sQK + 8192, qb + per_which, 16384, &bar_qk);                              // K hi
                        bulk_g2s(sQK + 8192 + 16384, qb + 4 * per_which, 16384, &bar_qk);                  // K lo
                    }
                    __syncwarp();
                };
                auto issue_v = [&](int64_t win, int stage) {
                    if (elect_one()) {
                        mbar_expect_tx(&bar_full[stage], 16384);
                        bulk_g2s(sVb + stage * 16384, qkv + (win * heads + hh) * 8192 + 2 * per_which, 16384, &bar_full[stage]);
                    }
                    __syncwarp();
                };
                issue_qk(g);
                issue_v(g, 0);
                const uint32_t sQh = smem_u32(sQK), sQl = sQh + 4096, sKh = sQh + 8192, sKl = sKh + 16384;
                const uint32_t q_hi = smem_desc_hi(128), dqh = smem_desc_lo(sQh, 2048), dql = smem_desc_lo(sQl, 2048);   // Q: [2][128][8]
                const uint32_t k_hi = smem_desc_hi(128), v_hi = smem_desc_hi(8192);
                int it = 0;
                for (int64_t win = g; win < B_; win += groups, ++it) {
                    const int stage = it & 1;
                    const uint32_t ph_v = (it >> 1) & 1, ph = it & 1;
                    const uint32_t sV = smem_u32(sVb + stage * 16384);
                    if (it > 0) {   // previous tile: O read back and its V / P consumed -> TMEM and the other V stage are free
                        mbar_wait_warp(&bar_epi, (it - 1) & 1);
                        tc_fence_after();
                    }
                    const bool more = win + groups < B_;
                    if (more) issue_v(win + groups, stage ^ 1);
                    mbar_wait_warp(&bar_qk, ph);
                    tc_fence_after();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t dkh = smem_desc_lo(sKh + half * 256 * 16, 8192), dkl = smem_desc_lo(sKl + half * 256 * 16, 8192);
                        mma_ss_w(tmem + half * 256, dqh, q_hi, dkh, k_hi, idesc_s, 0u);
                        mma_ss_w(tmem + half * 256, dql, q_hi, dkh, k_hi, idesc_s, 1u);
                        mma_ss_w(tmem + half * 256, dqh, q_hi, dkl, k_hi, idesc_s, 1u);
                    }
                    mma_commit_w(&bar_s);
                    // the score MMAs were the only readers of the Q / K buffer: refill it for the next window as soon as
                    // they have completed - the copy lands long before the softmax of this tile is through
                    mbar_wait_warp(&bar_s, ph);
                    if (more) issue_qk(win + groups);
                    mbar_wait_warp(&bar_full[stage], ph_v);
                    tc_fence_after();
                    const uint32_t dv = smem_desc_lo(sV, 128);
                    uint32_t first = 1;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c)
#pragma unroll 1
                        for (int q = 0; q < 4; ++q) {
                            mbar_wait_lean(&bar_p[q * 4 + c], ph);
                            tc_fence_after();
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int ks = q * 8 + c * 2 + j;   // k-step = keys [16 ks, 16 ks + 16)
                                mma_ts_w(tmem + 16, tmem + q * 128 + c * 32 + j * 8, dv + (uint32_t)(ks * 256 / 16), v_hi, idesc_o, first ? 0u : 1u);
                                first = 0;
                            }
                        }
                    mma_commit_w(&bar_o);
                }
            }
        }
    } else {
        // ============================================== softmax warps ================================================
        const int quarter = warp >> 2;                 // key quarter (128 keys) handled by this warp
        const int row = (warp & 3) * 32 + lane;        // query row inside the tile = TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t scol = tmem + lane_base + quarter * 128;
        const uint16_t *brow = sBias + row * kBiasPitch + quarter * 128;
        int it = 0;
        for (int64_t win = g; win < B_; win += groups, ++it) {
            const uint32_t ph = it & 1;
            if (it == 0) mbar_wait(&bar_bias, 0);
            mbar_wait(&bar_s, ph);
            tc_fence_after();
            // ---- pass 1: row maximum of s + bias over this warp's 128 keys (FHADD + FMNMX3: 1.5 ALU ops per key) ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(scol + c * 32, r);
                tmem_wait_ld();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 b = *reinterpret_cast<const uint4 *>(brow + c * 32 + q4 * 8);
                    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t0, t1;
                        add16x2<F16>(bw[e], __uint_as_float(r[q4 * 8 + 2 * e]), __uint_as_float(r[q4 * 8 + 2 * e + 1]), t0, t1);
                        mx = max3(mx, t0, t1);
                    }
                }
            }
            sMax[quarter * 128 + row] = mx;
            softmax_warps_sync();
            mx = fmaxf(fmaxf(sMax[row], sMax[128 + row]), fmaxf(sMax[256 + row], sMax[384 + row]));
            // ---- pass 2: p = 2^(s + bias - max), row sum; each 32-key chunk of P is packed over the first half of its
            // own S chunk and announced to the issuer ----
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(scol + c * 32, r);
                tmem_wait_ld();
                uint32_t pk[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 b = *reinterpret_cast<const uint4 *>(brow + c * 32 + q4 * 8);
                    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t0, t1;
                        add16x2<F16>(bw[e], __uint_as_float(r[q4 * 8 + 2 * e]) - mx, __uint_as_float(r[q4 * 8 + 2 * e + 1]) - mx, t0, t1);
                        const float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
                        sum += p0 + p1;
                        pk[q4 * 4 + e] = pack16<F16>(p0, p1);
                    }
                }
                tmem_st16(scol + c * 32, pk);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_p[quarter * 4 + c]);
            }
            sSum[quarter * 128 + row] = sum;
            softmax_warps_sync();   // every row sum is visible; also orders this tile's sMax reads before the next tile's writes
            if (warp < 4) {
                mbar_wait(&bar_o, ph);
                tc_fence_after();
                uint32_t r[16];
                tmem_ld16(tmem + lane_base + 16, r);
                tmem_wait_ld();
                const float inv = 1.f / ((sSum[row] + sSum[128 + row]) + (sSum[256 + row] + sSum[384 + row]));
                uint4 lo, hi;
                lo.x = pack16<F16>(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                lo.y = pack16<F16>(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                lo.z = pack16<F16>(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                lo.w = pack16<F16>(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                hi.x = pack16<F16>(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
                hi.y = pack16<F16>(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
                hi.z = pack16<F16>(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
                hi.w = pack16<F16>(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
                uint4 *dst = reinterpret_cast<uint4 *>(o + (win * 512 + qt * 128 + row) * (int64_t)C + hh * 16);
                if constexpr (SPLIT) {
                    uint4 llo, lhi;
                    split_pair(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv, lo.x, llo.x);
                    split_pair(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv, lo.y, llo.y);
                    split_pair(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv, lo.z, llo.z);
                    split_pair(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv, lo.w, llo.w);
                    split_pair(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv, hi.x, lhi.x);
                    split_pair(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv, hi.y, lhi.y);
                    split_pair(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv, hi.z, lhi.z);
                    split_pair(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv, hi.w, lhi.w);
                    uint4 *dl = reinterpret_cast<uint4 *>(o + B_ * 512 * (int64_t)C + (win * 512 + qt * 128 + row) * (int64_t)C + hh * 16);
                    dl[0] = llo;
                    dl[1] = lhi;
                }
                dst[0] = lo;
                dst[1] = hi;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_epi);   // O consumed: the issuer may overwrite TMEM with the next S
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// -------------------------------------------------------------------------------------------------- launcher ------
static int pick_ntile(int n) {
    for (int nt = 128; nt >= 16; nt -= 16)
        if (n % nt == 0) return nt;
    return 0;
}
static uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

bool attn_tc_supported(int D1, int H1, int W1, int C, int heads, int ws) {
    return ws == 8 && C == heads * 16 && C % 16 == 0 && C <= 384 && D1 % 8 == 0 && H1 % 8 == 0 && W1 % 8 == 0;
}

size_t attn_tc_bias_image_bytes(int heads) { return (size_t)heads * 4 * kBiasBytes; }

int attn_tc_bias_image(const void *table, int table_dtype, const int64_t *index, void *img, bool f16, int heads,
                       int table_rows, cudaStream_t st) {
    const int64_t total = (int64_t)heads * 512 * (kBiasPitch / 2);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (f16)
        relpos_bias_image_kernel<true><<<grid, 256, 0, st>>>(table, table_dtype, index, (uint16_t *)img, heads, table_rows);
    else
        relpos_bias_image_kernel<false><<<grid, 256, 0, st>>>(table, table_dtype, index, (uint16_t *)img, heads, table_rows);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

// K chunks per pass and N tile of a projection so that the operand images fit in shared memory
struct LinearPlan { int nt, kpc; size_t smem; };
static LinearPlan plan_linear(int C, int Nout, bool split) {
    LinearPlan p;
    p.nt = pick_ntile(Nout);
    p.kpc = C / 8;
    const size_t mult = split ? 2 : 1;
    auto bytes = [&](int kpc) { return mult * ((size_t)kpc * 2048 + (size_t)kpc * p.nt * 16); };
    while (p.nt && bytes(p.kpc) > 196 * 1024 && p.kpc % 4 == 0) p.kpc /= 2;   // passes of an even number of chunks
    p.smem = bytes(p.kpc);
    return p;
}

template <bool F16>
static int attn_tc_run(const void *x, int x_dtype, const uint16_t *qkv_w, const uint16_t *qkv_b, const uint16_t *proj_w,
                       const uint16_t *proj_b, const uint16_t *bias_img, void *out, bool out_f32, void *workspace, int B,
                       int D1, int H1, int W1, int C, int heads, float scale, cudaStream_t st) {
    TcWindowMap map;
    map.D1 = D1; map.H1 = H1; map.W1 = W1;
    map.nWy = H1 / 8; map.nWx = W1 / 8; map.nW = (D1 / 8) * map.nWy * map.nWx;
    const int64_t B_ = (int64_t)B * map.nW;
    const int64_t M = B_ * 512;
    uint16_t *qkv = reinterpret_cast<uint16_t *>(workspace);  // [3][B_][heads][2][512][8]
    uint16_t *obuf = qkv + 3 * M * C;                         // [M][C]
    static unsigned long long attrs_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attrs_done)) {
        const int big = 200 * 1024;
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<true, float, F16, uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<true, __nv_bfloat16, F16, uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<false, uint16_t, F16, uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<false, uint16_t, F16, float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(attn_core_tc_kernel<F16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCoreSmem));
    }
    {
        const LinearPlan lp = plan_linear(C, 3 * C, false);
        if (!lp.nt || lp.smem > 200 * 1024) return WF_ERR_BAD_SHAPE;
        dim3 grid((unsigned)(M / 128), (unsigned)(3 * C / lp.nt));
        if (x_dtype == WF_F32)
            linear_tc_kernel<true, float, F16, uint16_t, false><<<grid, 128, lp.smem, st>>>((const float *)x, nullptr, qkv_w, nullptr, qkv_b, nullptr, qkv, C, lp.nt, 3 * C, map, heads, B_, scale * kLog2e, pow2_cols(lp.nt), lp.kpc);
        else
            linear_tc_kernel<true, __nv_bfloat16, F16, uint16_t, false><<<grid, 128, lp.smem, st>>>((const __nv_bfloat16 *)x, nullptr, qkv_w, nullptr, qkv_b, nullptr, qkv, C, lp.nt, 3 * C, map, heads, B_, scale * kLog2e, pow2_cols(lp.nt), lp.kpc);
        WF_LAUNCH_CHECK();
    }
    {
        const int combos = heads * 4;
        int groups = kNumSMs / combos;
        if (groups < 1) groups = 1;
        if (groups > B_) groups = (int)B_;
        attn_core_tc_kernel<F16, false><<<combos * groups, 544, kCoreSmem, st>>>(qkv, bias_img, obuf, heads, B_, groups);
        WF_LAUNCH_CHECK();
    }
    {
        const LinearPlan lp = plan_linear(C, C, false);
        if (!lp.nt || lp.smem > 200 * 1024) return WF_ERR_BAD_SHAPE;
        dim3 grid((unsigned)(M / 128), (unsigned)(C / lp.nt));
        if (out_f32)
            linear_tc_kernel<false, uint16_t, F16, float, false><<<grid, 128, lp.smem, st>>>(obuf, nullptr, proj_w, nullptr, proj_b, nullptr, (float *)out, C, lp.nt, C, map, heads, B_, 1.f, pow2_cols(lp.nt), lp.kpc);
        else
            linear_tc_kernel<false, uint16_t, F16, uint16_t, false><<<grid, 128, lp.smem, st>>>(obuf, nullptr, proj_w, nullptr, proj_b, nullptr, (uint16_t *)out, C, lp.nt, C, map, heads, B_, 1.f, pow2_cols(lp.nt), lp.kpc);
        WF_LAUNCH_CHECK();
    }
    return WF_OK;
}

size_t attn_tc_split_workspace_bytes(int64_t tokens, int C) { return 7 * (size_t)tokens * (size_t)C * 2 + 256; }

// Compensated ("split") fp16 path: x fp32, weights as hi / lo fp16 pairs, fp32 biases, fp32 result.
// workspace: [5][B_][heads][2][512][8] (q hi, k hi, v, q lo, k lo) + [2][M][C] (o hi, o lo), 16-bit.
int attn_tc_forward_split(const float *x, const uint16_t *qkv_w_hi, const uint16_t *qkv_w_lo, const float *qkv_b,
                          const uint16_t *proj_w_hi, const uint16_t *proj_w_lo, const float *proj_b, const uint16_t *bias_img,
                          float *out, void *workspace, int B, int D1, int H1, int W1, int C, int heads, float scale,
                          cudaStream_t st) {
    TcWindowMap map;
    map.D1 = D1; map.H1 = H1; map.W1 = W1;
    map.nWy = H1 / 8; map.nWx = W1 / 8; map.nW = (D1 / 8) * map.nWy * map.nWx;
    const int64_t B_ = (int64_t)B * map.nW;
    const int64_t M = B_ * 512;
    uint16_t *qkv = reinterpret_cast<uint16_t *>(workspace);
    uint16_t *obuf = qkv + 5 * M * C;                         // [2][M][C]
    static unsigned long long attrs_done = 0;   // per-device opt-in bits
    if (first_use_on_current_device(attrs_done)) {
        const int big = 200 * 1024;
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<true, float, true, uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<false, uint16_t, true, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WF_CUDA_CHECK(cudaFuncSetAttribute(attn_core_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCoreSmemSplit));
    }
    {
        const LinearPlan lp = plan_linear(C, 3 * C, true);
        if (!lp.nt || lp.smem > 200 * 1024) return WF_ERR_BAD_SHAPE;
        dim3 grid((unsigned)(M / 128), (unsigned)(3 * C / lp.nt));
        linear_tc_kernel<true, float, true, uint16_t, true><<<grid, 128, lp.smem, st>>>(x, nullptr, qkv_w_hi, qkv_w_lo, nullptr, qkv_b, qkv, C, lp.nt, 3 * C, map, heads, B_, scale * kLog2e, pow2_cols(lp.nt), lp.kpc);
        WF_LAUNCH_CHECK();
    }
    {
        const int combos = heads * 4;
        int groups = kNumSMs / combos;
        if (groups < 1) groups = 1;
        if (groups > B_) groups = (int)B_;
        attn_core_tc_kernel<true, true><<<combos * groups, 544, kCoreSmemSplit, st>>>(qkv, bias_img, obuf, heads, B_, groups);
        WF_LAUNCH_CHECK();
    }
    {
        const LinearPlan lp = plan_linear(C, C, true);
        if (!lp.nt || lp.smem > 200 * 1024) return WF_ERR_BAD_SHAPE;
        dim3 grid((unsigned)(M / 128), (unsigned)(C / lp.nt));
        linear_tc_kernel<false, uint16_t, true, float, true><<<grid, 128, lp.smem, st>>>(obuf, obuf + M * C, proj_w_hi, proj_w_lo, nullptr, proj_b, out, C, lp.nt, C, map, heads, B_, 1.f, pow2_cols(lp.nt), lp.kpc);
        WF_LAUNCH_CHECK();
    }
    return WF_OK;
}

int attn_tc_forward(const void *x, int x_dtype, const void *qkv_w, const void *qkv_b, const void *proj_w,
                    const void *proj_b, const void *bias_img, void *out, bool out_f32, void *workspace, bool f16, int B,
                    int D1, int H1, int W1, int C, int heads, float scale, cudaStream_t st) {
    using u16 = uint16_t;
    if (f16)
        return attn_tc_run<true>(x, x_dtype, (const u16 *)qkv_w, (const u16 *)qkv_b, (const u16 *)proj_w, (const u16 *)proj_b,
                                 (const u16 *)bias_img, out, out_f32, workspace, B, D1, H1, W1, C, heads, scale, st);
    return attn_tc_run<false>(x, x_dtype, (const u16 *)qkv_w, (const u16 *)qkv_b, (const u16 *)proj_w, (const u16 *)proj_b,
                              (const u16 *)bias_img, out, out_f32, workspace, B, D1, H1, W1, C, heads, scale, st);
}

}  // namespace wf
