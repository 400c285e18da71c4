// CCF_FFN's two pointwise GEMMs fused with the normalisations around them (tcgen05 / TMEM, sm_100a).
//
// Reference: Block.multi_scale_forward's tail `x = x + drop_path(mlp(norm2(x)))` (network_models/wave_helper.py:509) with
// CCF_FFN.forward (wave_helper.py:260-294):  t = pwconv(n) -> LayerNorm(4C) -> GELU -> depthwise 3^3 -> LayerNorm(4C) -> GELU
// -> fc -> n + .   (n = norm2(x); CCF_FFN adds its own input back, the block adds x).
// Unfused, every arrow is a pass over a [voxels, 4C] tensor: norm2 (read x, write n twice), the pwconv GEMM (write 4C), LN +
// GELU (read 4C, write 4C), the depthwise stencil, LN + GELU again, the fc GEMM, and a 3-operand residual pass.  A voxel's
// row never needs another voxel's data outside the depthwise stencil, and one TMEM lane holds one voxel's row, so:
//
//   ffn_front_kernel   x (fp32 stream) -> norm2 in registers -> A operand -> tcgen05.mma [128 x 4C x C] -> epilogue reads the
//                      row from TMEM twice (statistics, then normalise + GELU) -> t1 (16-bit) : ONE read of x, ONE write of t1
//   (depthwise 3^3 stencil kernel, unchanged)
//   ffn_back_kernel    t2 (16-bit) staged as the K-major A image -> LayerNorm + GELU in place (thread = row) ->
//                      tcgen05.mma [128 x C x 4C] -> epilogue adds the fc bias, x and norm2(x) (recomputed from x) -> fp32 stream
//
// Both are HBM-bound (K = C resp. N = C is tiny); persistent CTAs keep the weights resident in shared memory.
// C = 48 or 96 (stages 1 / 2: 4C <= 384 fp32 accumulator columns fit TMEM; the 16^3 / 8^3 stages keep the library GEMMs).
#include "tc_common.cuh"
#include "wf_common.cuh"

namespace wf {

using namespace tc;

// GELU(x) = x * Phi(x), exact (erf) form via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7), as in layernorm.cu
__device__ __forceinline__ float ffn_gelu(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    return 0.5f * x + 0.5f * fabsf(x) * fmaf(-p, e, 1.f);
}

__device__ __forceinline__ void ffn_cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ffn_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ------------------------------------------------------------------------------------------------------ front ------
struct FfnFrontArgs {
    const float *x;          // [M, C] fp32 residual stream (after the attention branch)
    const float *g2, *b2;    // Block.norm2 (LayerNorm over C), fp32 [C]
    const uint16_t *w1;      // pwconv weight [4C, C] in the 16-bit operand format
    const float *bias1;      // pwconv bias fp32 [4C]
    const float *g1, *be1;   // CCF_FFN.norm1 (LayerNorm over 4C), fp32 [4C]
    uint16_t *t1;            // [M, 4C] 16-bit: GELU(LN(pwconv(norm2(x))))
    int64_t M, ntiles;
    float eps2, eps1;
};

// 256 threads.  Load phase: thread pair (2r, 2r+1) owns row r, half of the channels each.  Epilogue: warp w reads TMEM
// lane quadrant w % 4 (rows), column half w / 4.
template <bool F16, int C>
__global__ void __launch_bounds__(256, C == 48 ? 2 : 1) ffn_front_kernel(FfnFrontArgs a) {
    constexpr int N = 4 * C;                 // 192 / 384 accumulator columns
    constexpr int NT = 192;                  // one tcgen05.mma covers 192 columns
    constexpr int KCH = C / 8;               // 16-byte K chunks per row
    constexpr int HC = C / 2;                // channels per thread in the load phase
    constexpr uint32_t TMEM_COLS = N <= 256 ? 256 : 512;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sB = smem;                                     // [KCH][N][16 B]
    uint8_t *sA = smem + (size_t)KCH * N * 16;              // [KCH][128][16 B]
    float *sConst = reinterpret_cast<float *>(sA + (size_t)KCH * 2048);   // bias1 | g1 | be1, N floats each
    float2 *sStat = reinterpret_cast<float2 *>(sConst + 3 * N);           // [2 column halves][128 rows] (sum, sum of squares)
    float *sN2 = reinterpret_cast<float *>(sStat + 256);                  // norm2 gamma | beta, C floats each
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    for (int idx = tid; idx < N * KCH; idx += 256) {        // resident weights: [N][C] row-major -> K-major chunk image
        const int r = idx % N, kc = idx / N;
        *reinterpret_cast<uint4 *>(sB + ((size_t)kc * N + r) * 16) = __ldg(reinterpret_cast<const uint4 *>(a.w1 + (int64_t)r * C) + kc);
    }
    for (int i = tid; i < N; i += 256) {
        sConst[i] = a.bias1 ? a.bias1[i] : 0.f;
        sConst[N + i] = a.g1 ? a.g1[i] : 1.f;
        sConst[2 * N + i] = a.be1 ? a.be1[i] : 0.f;
    }
    for (int i = tid; i < C; i += 256) {
        sN2[i] = a.g2 ? a.g2[i] : 1.f;
        sN2[C + i] = a.b2 ? a.b2[i] : 0.f;
    }
    __syncthreads();
    const int row = tid >> 1, half = tid & 1;
    const float *g2r = sN2 + half * HC, *b2r = sN2 + C + half * HC;      // this thread's slice of norm2's affine parameters
    auto load_row = [&](int64_t tile, float (&v)[HC]) {
        const int64_t m = tile * 128 + row;
        if (m < a.M) {
            const float4 *src = reinterpret_cast<const float4 *>(a.x + m * C + half * HC);
#pragma unroll
            for (int i = 0; i < HC / 4; ++i) {
                const float4 t = __ldg(src + i);
                v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < HC; ++i) v[i] = 0.f;
        }
    };
    float cur[HC];
    if ((int64_t)blockIdx.x < a.ntiles) load_row(blockIdx.x, cur);
    uint32_t phase = 0;
    uint32_t tmem = 0;
    const uint32_t idesc = instr_desc_h16<F16>(128, NT, false);
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        // ---- norm2 of the row (two threads per row), A image ----
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) s += cur[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        const float mean = s * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) {
            const float d = cur[i] - mean;
            q = fmaf(d, d, q);
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        const float rstd = rsqrtf(q * (1.f / C) + a.eps2);
#pragma unroll
        for (int kc = 0; kc < HC / 8; ++kc) {
            uint4 u;
            uint32_t *uw = &u.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = kc * 8 + 2 * e;
                uw[e] = pack_h16<F16>(fmaf((cur[c] - mean) * rstd, g2r[c], b2r[c]), fmaf((cur[c + 1] - mean) * rstd, g2r[c + 1], b2r[c + 1]));
            }
            *reinterpret_cast<uint4 *>(sA + (size_t)(half * (HC / 8) + kc) * 2048 + row * 16) = u;
        }
        // prefetch the next tile's row while this tile is multiplied and normalised
        const int64_t next = tile + gridDim.x;
        if (next < a.ntiles) load_row(next, cur);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();      // A image complete; every warp has finished reading the previous tile's accumulators
        tc_fence_after();
        tmem = tmem_slot;
        if (tid == 0) {
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int nt = 0; nt < N / NT; ++nt)
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks)
                    mma_ss(tmem + nt * NT, smem_desc(a0 + ks * 2 * 2048, 2048, 128),
                           smem_desc(b0 + (ks * 2 * N + nt * NT) * 16, N * 16, 128), idesc, ks > 0 ? 1u : 0u);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue: LayerNorm over the 4C columns of the row + GELU, straight out of TMEM ----
        const int quad = warp & 3, chalf = warp >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + chalf * (N / 2);
        const float *cb = sConst + chalf * (N / 2);
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < N / 2; c += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float v = __uint_as_float(r[e]) + cb[c + e];
                sum += v;
                sq = fmaf(v, v, sq);
            }
        }
        sStat[chalf * 128 + erow] = make_float2(sum, sq);
        __syncthreads();
        const float2 s0 = sStat[erow], s1 = sStat[128 + erow];
        const float mu = (s0.x + s1.x) * (1.f / N);
        const float var = fmaxf((s0.y + s1.y) * (1.f / N) - mu * mu, 0.f);
        const float rs = rsqrtf(var + a.eps1);
        const int64_t m = tile * 128 + erow;
        uint16_t *dst = a.t1 + m * N + chalf * (N / 2);
#pragma unroll 1
        for (int c = 0; c < N / 2; c += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            tmem_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int c0 = c + 2 * e;
                const float v0 = fmaf((__uint_as_float(r[2 * e]) + cb[c0] - mu) * rs, cb[N + c0], cb[2 * N + c0]);
                const float v1 = fmaf((__uint_as_float(r[2 * e + 1]) + cb[c0 + 1] - mu) * rs, cb[N + c0 + 1], cb[2 * N + c0 + 1]);
                pk[e] = pack_h16<F16>(ffn_gelu(v0), ffn_gelu(v1));
            }
            if (m < a.M) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<uint4 *>(dst + c + 8 * i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
        }
        tc_fence_before();
        // (the next iteration's __syncthreads orders these TMEM reads and the sStat reads before the next tile's writes)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------- back ------
struct FfnBackArgs {
    const uint16_t *t2;      // [M, 4C] 16-bit: the depthwise stencil's result
    const float *g, *be;     // CCF_FFN.norm2 (LayerNorm over 4C), fp32 [4C]
    const uint16_t *wfc;     // fc weight [C, 4C] in the 16-bit operand format
    const float *bfc;        // fc bias fp32 [C]
    const float *x;          // [M, C] fp32 residual stream (the front kernel's input)
    const float *g2, *b2;    // Block.norm2, fp32 [C]
    float *out;              // [M, C] fp32: x + norm2(x) + fc(GELU(LN(t2))) + bias
    int64_t M, ntiles;
    float eps, eps2;
};

// 128 threads, thread = row (= TMEM lane).
template <bool F16, int C>
__global__ void __launch_bounds__(128) ffn_back_kernel(FfnBackArgs a) {
    using T16 = typename std::conditional<F16, __half, __nv_bfloat16>::type;
    constexpr int K = 4 * C;
    constexpr int KCH = K / 8;               // 24 / 48 chunks per row
    constexpr uint32_t TMEM_COLS = C <= 64 ? 64 : 128;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sB = smem;                                      // [KCH][C][16 B]
    uint8_t *sA = smem + (size_t)KCH * C * 16;               // [KCH][128][16 B]
    float *sConst = reinterpret_cast<float *>(sA + (size_t)KCH * 2048);   // g | be (K floats each) | bfc | g2 | b2 (C each)
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    for (int idx = tid; idx < C * KCH; idx += 128) {
        const int r = idx % C, kc = idx / C;
        *reinterpret_cast<uint4 *>(sB + ((size_t)kc * C + r) * 16) = __ldg(reinterpret_cast<const uint4 *>(a.wfc + (int64_t)r * K) + kc);
    }
    for (int i = tid; i < K; i += 128) {
        sConst[i] = a.g ? a.g[i] : 1.f;
        sConst[K + i] = a.be ? a.be[i] : 0.f;
    }
    for (int i = tid; i < C; i += 128) {
        sConst[2 * K + i] = a.bfc ? a.bfc[i] : 0.f;
        sConst[2 * K + C + i] = a.g2 ? a.g2[i] : 1.f;
        sConst[2 * K + 2 * C + i] = a.b2 ? a.b2[i] : 0.f;
    }
    // Tile of t2 -> A image (asynchronous).  One instruction of a warp copies 64 contiguous bytes (two whole sectors) of each of 8
    // rows: lane = (row % 8) + 8 * (chunk % 4), i.e. conflict-free shared-memory writes and no half-used sectors (a thread copying
    // its own row fetched every 32-byte sector twice, 16 bytes at a time).
    const int st_lr = tid & 7, st_lc = (tid >> 3) & 3;
    auto stage_row = [&](int64_t tile) {
        const int64_t m0 = tile * 128;
#pragma unroll 2
        for (int rg = warp; rg < 16; rg += 4) {
            const int r = rg * 8 + st_lr;
            const bool live_r = m0 + r < a.M;
            const uint16_t *src = a.t2 + (m0 + r) * K;
#pragma unroll 4
            for (int cg = 0; cg < KCH / 4; ++cg) {
                const int kc = cg * 4 + st_lc;
                uint8_t *dst = sA + (size_t)kc * 2048 + r * 16;
                if (live_r) ffn_cp_async16(dst, src + kc * 8);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t phase = 0;
    uint32_t tmem = 0;
    const uint32_t idesc = instr_desc_h16<F16>(128, C, false);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    if ((int64_t)blockIdx.x < a.ntiles) stage_row(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t m = tile * 128 + tid;
        const bool live = m < a.M;
        // residual-stream row: issued now, consumed in the epilogue
        float xr[C];
        if (live) {
            const float4 *src = reinterpret_cast<const float4 *>(a.x + m * C);
#pragma unroll
            for (int i = 0; i < C / 4; ++i) {
                const float4 t = __ldg(src + i);
                xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < C; ++i) xr[i] = 0.f;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();      // the tile was copied cooperatively: every thread's copies have landed
        // ---- LayerNorm(4C) + GELU of this thread's own row, in place in the A image ----
        float sum = 0.f, sq = 0.f;
#pragma unroll 4
        for (int kc = 0; kc < KCH; ++kc) {
            float f[8];
            Pack<T16>::unpack(*reinterpret_cast<const uint4 *>(sA + (size_t)kc * 2048 + tid * 16), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sum += f[e];
                sq = fmaf(f[e], f[e], sq);
            }
        }
        const float mu = sum * (1.f / K);
        const float rs = rsqrtf(fmaxf(sq * (1.f / K) - mu * mu, 0.f) + a.eps);
#pragma unroll 2
        for (int kc = 0; kc < KCH; ++kc) {
            uint4 *cell = reinterpret_cast<uint4 *>(sA + (size_t)kc * 2048 + tid * 16);
            float f[8];
            Pack<T16>::unpack(*cell, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = ffn_gelu(fmaf((f[e] - mu) * rs, sConst[kc * 8 + e], sConst[K + kc * 8 + e]));
            *cell = Pack<T16>::pack(f);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();      // A image complete; the previous tile's accumulators have been read by every warp
        tc_fence_after();
        tmem = tmem_slot;
        if (tid == 0) {
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll 4
            for (int ks = 0; ks < K / 16; ++ks)
                mma_ss(tmem, smem_desc(a0 + ks * 2 * 2048, 2048, 128), smem_desc(b0 + ks * 2 * C * 16, C * 16, 128), idesc,
                       ks > 0 ? 1u : 0u);
            mma_commit(&bar);
        }
        // norm2(x) of this row while the tensor core works
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) s2 += xr[i];
        const float mean2 = s2 * (1.f / C);
        float q2 = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float d = xr[i] - mean2;
            q2 = fmaf(d, d, q2);
        }
        const float rstd2 = rsqrtf(q2 * (1.f / C) + a.eps2);
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        // the A image is free again: start the next tile's copies underneath the epilogue
        const int64_t next = tile + gridDim.x;
        if (next < a.ntiles) stage_row(next);
        float *dst = a.out + m * C;
        const float *cb = sConst + 2 * K;
#pragma unroll
        for (int c = 0; c < C; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + lane_base + c, r);
            tmem_wait_ld();
            if (live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = c + 4 * i + e;
                        const float n = fmaf((xr[cc] - mean2) * rstd2, cb[C + cc], cb[2 * C + cc]);
                        o[e] = (xr[cc] + n) + (__uint_as_float(r[4 * i + e]) + cb[cc]);
                    }
                    *reinterpret_cast<float4 *>(dst + c + 4 * i) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        tc_fence_before();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, TMEM_COLS);
}

// --------------------------------------------------------------------------------------- ProjectionUpsample tail ------
// out[r, :N] = W1 . GELU(h[r, :K1]) + b1  +  W2 . u[r, :K2] + b2        (16-bit in, fp32 accumulate, 16-bit out, row pitch `os`)
//
// Reference: the end of ProjectionUpsample.forward (network_models/wave_helper.py:75-81): `x = self.conv3(self.act(x))` (single
// conv, or the last conv of the double-conv Sequential after its own GELU) and `x = x + self.res_conv(input)`, where res_conv =
// Upsample + 1^3 convolution - its upsampled operand u is shared with the main branch here.  Unfused this is a GELU pass over h
// (read + write), two library GEMMs with dense temporaries and an add into the strided concatenation slice: ~3.4 GB per six windows
// for learnable_up3 where h + u + out are 1.06 GB.  One persistent kernel, 256 threads: the two A images are staged with
// coalesced cp.async (a thread copies chunks q, q + 256, ...), the GELU is applied in place to the cells a thread copied itself,
// one converged warp issues K1 / 16 + K2 / 16 tcgen05.mma of N columns, the epilogue (thread = row) adds the biases and writes
// the row into the destination slice while the next tile's images are already loading.
struct PwTailArgs {
    const uint16_t *h, *u;       // [M, K1], [M, K2] dense
    const uint16_t *w1, *w2;     // [N, K1], [N, K2]
    const float *b1, *b2;        // fp32 [N] or NULL
    const float *addend;         // optional fp32 [M, N] dense, added before rounding
    uint16_t *out;               // [M, N], row pitch os elements
    int64_t M, ntiles, os;
};

template <bool F16, int K1, int K2, int N>
__global__ void __launch_bounds__(256) pw_gelu_dual_kernel(PwTailArgs a) {
    using T16 = typename std::conditional<F16, __half, __nv_bfloat16>::type;
    constexpr int KCH1 = K1 / 8, KCH2 = K2 / 8;
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sB1 = smem;                                     // [KCH1][N][16 B]
    uint8_t *sB2 = sB1 + (size_t)KCH1 * N * 16;              // [KCH2][N][16 B]
    uint8_t *sA1 = sB2 + (size_t)KCH2 * N * 16;              // [KCH1][128][16 B]
    uint8_t *sA2 = sA1 + (size_t)KCH1 * 2048;                // [KCH2][128][16 B]
    float *sBias = reinterpret_cast<float *>(sA2 + (size_t)KCH2 * 2048);   // b1 + b2, N floats
    const int tid = threadIdx.x, warp = (int)warp_idx_uniform();

    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    for (int idx = tid; idx < N * KCH1; idx += 256) {
        const int r = idx % N, kc = idx / N;
        *reinterpret_cast<uint4 *>(sB1 + ((size_t)kc * N + r) * 16) = __ldg(reinterpret_cast<const uint4 *>(a.w1 + (int64_t)r * K1) + kc);
    }
    for (int idx = tid; idx < N * KCH2; idx += 256) {
        const int r = idx % N, kc = idx / N;
        *reinterpret_cast<uint4 *>(sB2 + ((size_t)kc * N + r) * 16) = __ldg(reinterpret_cast<const uint4 *>(a.w2 + (int64_t)r * K2) + kc);
    }
    for (int i = tid; i < N; i += 256) sBias[i] = (a.b1 ? a.b1[i] : 0.f) + (a.b2 ? a.b2[i] : 0.f);
    // chunk q of a tile's h block: row q / KCH1, 16-byte chunk q % KCH1 (consecutive threads -> consecutive 16 bytes of global memory)
    // one instruction of a warp copies 64 contiguous bytes of each of 8 rows: lane = (row % 8) + 8 * (chunk % 4) - whole sectors on
    // the global side, conflict-free 16-byte writes on the shared side (consecutive lanes on consecutive chunks of one row wrote
    // 2048 bytes apart: the same bank group)
    const int st_lr = tid & 7, st_lc = (tid >> 3) & 3;
    auto stage = [&](int64_t tile) {
        const int64_t m0 = tile * 128;
        for (int rg = warp; rg < 16; rg += 8) {
            const int row = rg * 8 + st_lr;
            const bool live_r = m0 + row < a.M;
#pragma unroll 2
            for (int cg = 0; cg < KCH1 / 4; ++cg) {
                const int kc = cg * 4 + st_lc;
                uint8_t *dst = sA1 + (size_t)kc * 2048 + row * 16;
                if (live_r) ffn_cp_async16(dst, a.h + (m0 + row) * K1 + kc * 8);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll 2
            for (int cg = 0; cg < KCH2 / 4; ++cg) {
                const int kc = cg * 4 + st_lc;
                uint8_t *dst = sA2 + (size_t)kc * 2048 + row * 16;
                if (live_r) ffn_cp_async16(dst, a.u + (m0 + row) * K2 + kc * 8);
                else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t phase = 0;
    const uint32_t idesc = instr_desc_h16<F16>(128, N, false);
    if ((int64_t)blockIdx.x < a.ntiles) stage(blockIdx.x);
    fence_proxy_async();          // the weight images (generic-proxy stores) -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // GELU of h, in place, on the cells this thread copied itself (elementwise: no cross-thread dependency)
        for (int rg = warp; rg < 16; rg += 8) {
            const int row = rg * 8 + st_lr;
            for (int cg = 0; cg < KCH1 / 4; ++cg) {
                const int kc = cg * 4 + st_lc;
                uint4 *cell = reinterpret_cast<uint4 *>(sA1 + (size_t)kc * 2048 + row * 16);
                float f[8];
                Pack<T16>::unpack(*cell, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = ffn_gelu(f[e]);
                *cell = Pack<T16>::pack(f);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();      // both images complete; the previous tile's accumulators have been read by the epilogue warps
        tc_fence_after();
        if (warp == 4) {      // one converged warp issues (tc_common.cuh, "warp-uniform issue")
            const uint32_t a1 = smem_desc_lo(smem_u32(sA1), 2048), a2 = smem_desc_lo(smem_u32(sA2), 2048), ah = smem_desc_hi(128);
            const uint32_t b1 = smem_desc_lo(smem_u32(sB1), N * 16), b2 = smem_desc_lo(smem_u32(sB2), N * 16), bh = smem_desc_hi(128);
#pragma unroll 1
            for (int ks = 0; ks < K1 / 16; ++ks)
                mma_ss_w(tmem, a1 + (uint32_t)(ks * 2 * 2048 / 16), ah, b1 + (uint32_t)(ks * 2 * N), bh, idesc, ks > 0 ? 1u : 0u);
#pragma unroll 1
            for (int ks = 0; ks < K2 / 16; ++ks)
                mma_ss_w(tmem, a2 + (uint32_t)(ks * 2 * 2048 / 16), ah, b2 + (uint32_t)(ks * 2 * N), bh, idesc, 1u);
            mma_commit_w(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        // the images are free again: the next tile's copies run underneath the epilogue
        const int64_t next = tile + gridDim.x;
        if (next < a.ntiles) stage(next);
        if (warp < 4) {
            const int64_t m = tile * 128 + (warp & 3) * 32 + (tid & 31);
            uint16_t *dst = a.out + m * a.os;
#pragma unroll
            for (int c = 0; c < N; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem + lane_base + c, r);
                tmem_wait_ld();
                if (m < a.M) {
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(r[e]) + sBias[c + e];
                    if (a.addend != nullptr) {
                        const float4 *ad = reinterpret_cast<const float4 *>(a.addend + m * N + c);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 t = __ldg(ad + i);
                            f[4 * i] += t.x; f[4 * i + 1] += t.y; f[4 * i + 2] += t.z; f[4 * i + 3] += t.w;
                        }
                    }
                    uint4 lo, hi;
                    lo.x = pack_h16<F16>(f[0], f[1]);   lo.y = pack_h16<F16>(f[2], f[3]);   lo.z = pack_h16<F16>(f[4], f[5]);   lo.w = pack_h16<F16>(f[6], f[7]);
                    hi.x = pack_h16<F16>(f[8], f[9]);   hi.y = pack_h16<F16>(f[10], f[11]); hi.z = pack_h16<F16>(f[12], f[13]); hi.w = pack_h16<F16>(f[14], f[15]);
                    *reinterpret_cast<uint4 *>(dst + c) = lo;
                    *reinterpret_cast<uint4 *>(dst + c + 8) = hi;
                }
            }
        }
        tc_fence_before();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, TMEM_COLS);
}

template <bool F16, int K1, int K2, int N>
static int pw_tail_launch(const PwTailArgs &a, cudaStream_t st) {
    const size_t smem = (size_t)(K1 / 8 + K2 / 8) * N * 16 + (size_t)(K1 / 8 + K2 / 8) * 2048 + N * sizeof(float);
    static unsigned long long attr_done = 0;
    if (first_use_on_current_device(attr_done))
        WF_CUDA_CHECK(cudaFuncSetAttribute(pw_gelu_dual_kernel<F16, K1, K2, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)((224 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((224 * 1024) / (smem + 1024));
    const int64_t want = (int64_t)kNumSMs * (per_sm > 3 ? 3 : per_sm);
    const int64_t grid = a.ntiles < want ? a.ntiles : want;
    pw_gelu_dual_kernel<F16, K1, K2, N><<<(unsigned)grid, 256, smem, st>>>(a);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <bool F16, int C>
static int ffn_front_launch(const FfnFrontArgs &a, cudaStream_t st) {
    constexpr int N = 4 * C, KCH = C / 8;
    const size_t smem = (size_t)KCH * N * 16 + (size_t)KCH * 2048 + 3 * N * sizeof(float) + 2 * 128 * sizeof(float2) + 2 * C * sizeof(float);
    static unsigned long long attr_done = 0;
    if (first_use_on_current_device(attr_done))
        WF_CUDA_CHECK(cudaFuncSetAttribute(ffn_front_kernel<F16, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = N <= 256 ? 2 : 1;         // TMEM: 256 / 512 accumulator columns per CTA
    const int64_t grid = a.ntiles < (int64_t)kNumSMs * per_sm ? a.ntiles : (int64_t)kNumSMs * per_sm;
    ffn_front_kernel<F16, C><<<(unsigned)grid, 256, smem, st>>>(a);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template <bool F16, int C>
static int ffn_back_launch(const FfnBackArgs &a, cudaStream_t st) {
    constexpr int K = 4 * C, KCH = K / 8;
    const size_t smem = (size_t)KCH * C * 16 + (size_t)KCH * 2048 + (2 * K + 3 * C) * sizeof(float);
    static unsigned long long attr_done = 0;
    if (first_use_on_current_device(attr_done))
        WF_CUDA_CHECK(cudaFuncSetAttribute(ffn_back_kernel<F16, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)((220 * 1024) / smem) < 1 ? 1 : (int)((220 * 1024) / smem);
    const int64_t want = (int64_t)kNumSMs * (per_sm > 4 ? 4 : per_sm);
    const int64_t grid = a.ntiles < want ? a.ntiles : want;
    ffn_back_kernel<F16, C><<<(unsigned)grid, 128, smem, st>>>(a);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

using namespace wf;

extern "C" int wf_ffn_front(const float *x, const float *norm2_w, const float *norm2_b, float norm2_eps, const void *pw_w,
                            const float *pw_b, const float *ln_w, const float *ln_b, float ln_eps, void *t1, int dtype,
                            int64_t rows, int C, void *stream) {
    if (!x || !pw_w || !t1) return WF_ERR_NULL_POINTER;
    if (rows <= 0) return WF_ERR_BAD_SHAPE;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (C != 48 && C != 96) return WF_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(pw_w) || !aligned16(t1)) return WF_ERR_MISALIGNED;
    FfnFrontArgs a;
    a.x = x; a.g2 = norm2_w; a.b2 = norm2_b; a.w1 = (const uint16_t *)pw_w; a.bias1 = pw_b; a.g1 = ln_w; a.be1 = ln_b;
    a.t1 = (uint16_t *)t1; a.M = rows; a.ntiles = (rows + 127) / 128; a.eps2 = norm2_eps; a.eps1 = ln_eps;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F16) return C == 48 ? ffn_front_launch<true, 48>(a, st) : ffn_front_launch<true, 96>(a, st);
    return C == 48 ? ffn_front_launch<false, 48>(a, st) : ffn_front_launch<false, 96>(a, st);
}

extern "C" int wf_ffn_back(const void *t2, int dtype, const float *ln_w, const float *ln_b, float ln_eps, const void *fc_w,
                           const float *fc_b, const float *x, const float *norm2_w, const float *norm2_b, float norm2_eps,
                           float *out, int64_t rows, int C, void *stream) {
    if (!t2 || !fc_w || !x || !out) return WF_ERR_NULL_POINTER;
    if (rows <= 0) return WF_ERR_BAD_SHAPE;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (C != 48 && C != 96) return WF_ERR_UNSUPPORTED;
    if (!aligned16(t2) || !aligned16(fc_w) || !aligned16(x) || !aligned16(out)) return WF_ERR_MISALIGNED;
    FfnBackArgs a;
    a.t2 = (const uint16_t *)t2; a.g = ln_w; a.be = ln_b; a.wfc = (const uint16_t *)fc_w; a.bfc = fc_b; a.x = x; a.g2 = norm2_w;
    a.b2 = norm2_b; a.out = out; a.M = rows; a.ntiles = (rows + 127) / 128; a.eps = ln_eps; a.eps2 = norm2_eps;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WF_F16) return C == 48 ? ffn_back_launch<true, 48>(a, st) : ffn_back_launch<true, 96>(a, st);
    return C == 48 ? ffn_back_launch<false, 48>(a, st) : ffn_back_launch<false, 96>(a, st);
}

extern "C" int wf_pw_gelu_dual(const void *h, const void *u, int dtype, const void *w1, const float *b1, const void *w2, const float *b2,
                               const float *addend, void *out, int64_t rows, int K1, int K2, int N, int64_t out_row_stride, void *stream) {
    if (!h || !w1 || !out || (K2 > 0 && (!u || !w2))) return WF_ERR_NULL_POINTER;
    if (rows <= 0 || out_row_stride < N || out_row_stride % 8) return WF_ERR_BAD_SHAPE;
    if (dtype != WF_BF16 && dtype != WF_F16) return WF_ERR_BAD_DTYPE;
    if (!aligned16(h) || !aligned16(w1) || !aligned16(out) || (K2 > 0 && (!aligned16(u) || !aligned16(w2))) || (addend && !aligned16(addend)))
        return WF_ERR_MISALIGNED;
    PwTailArgs a;
    a.h = (const uint16_t *)h; a.u = (const uint16_t *)u; a.w1 = (const uint16_t *)w1; a.w2 = (const uint16_t *)w2; a.b1 = b1; a.b2 = K2 > 0 ? b2 : nullptr;
    a.addend = addend; a.out = (uint16_t *)out; a.M = rows; a.ntiles = (rows + 127) / 128; a.os = out_row_stride;
    cudaStream_t st = (cudaStream_t)stream;
    const bool f16 = dtype == WF_F16;
    if (K1 == 192 && K2 == 0 && N == 48) return f16 ? pw_tail_launch<true, 192, 0, 48>(a, st) : pw_tail_launch<false, 192, 0, 48>(a, st);
    if (K1 == 192 && K2 == 96 && N == 48) return f16 ? pw_tail_launch<true, 192, 96, 48>(a, st) : pw_tail_launch<false, 192, 96, 48>(a, st);
    if (K1 == 192 && K2 == 192 && N == 48) return f16 ? pw_tail_launch<true, 192, 192, 48>(a, st) : pw_tail_launch<false, 192, 192, 48>(a, st);
    return WF_ERR_UNSUPPORTED;
}
