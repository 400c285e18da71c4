// Kernel group 2, fp32-accurate SIMT path: window partition + QKV projection, attention core, output projection.
//
// Replaces Block.window_partition (reference network_models/wave_helper.py:450-461), Attention.forward
// (network_models/attention.py:83-104) and the reshape-only window reverse (wave_helper.py:498-499).
// This path computes every product in fp32 on the CUDA cores; it is the path used for dtype = WF_F32 (the
// fp32 parity gate, max-relative logit error <= 1e-4, cannot be met through TF32/BF16 tensor-core products) and the
// on-device cross-check of the tcgen05 path in attn_tc.cu (dtype = WF_BF16).
#include "wf_common.cuh"

namespace wf {

// ---- row -> voxel map of the window partition ---------------------------------------------------------------------
struct WindowMap {
    int D1, H1, W1, ws, N, nWy, nWx, nW;  // N = ws^3 tokens, nW windows per batch element
    __device__ inline int64_t voxel(int64_t m) const {  // m = (b*nW + widx)*N + token
        const int tok = (int)(m % N);
        const int64_t win = m / N;
        const int widx = (int)(win % nW);
        const int64_t b = win / nW;
        const int xb = widx % nWx, yb = (widx / nWx) % nWy, zb = widx / (nWx * nWy);
        const int dx = tok % ws, dy = (tok / ws) % ws, dz = tok / (ws * ws);
        return ((b * D1 + zb * ws + dz) * H1 + yb * ws + dy) * (int64_t)W1 + xb * ws + dx;
    }
};

// ---- out[m, n] = sum_k A[row(m), k] * Wt[n, k] + bias[n] -----------------------------------------------------------
// QKV = true : rows gathered through the window map; result scattered to head-major q/k/v buffers
//              [3][B_][heads][N][hd], q additionally multiplied by `scale`.
// QKV = false: plain [M, Nout] output.
template <typename TA, typename T, bool QKV>
__global__ void __launch_bounds__(256) linear_kernel(const TA *__restrict__ A, const T *__restrict__ Wt,
                                                     const T *__restrict__ bias, T *__restrict__ out, int64_t M,
                                                     int Nout, int K, WindowMap map, int heads, int hd, float scale,
                                                     int64_t B_) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    // each thread stages 4 elements of A and 4 of W per k-step: element (r, kk) with r = tid / 4, kk = (tid % 4) * 4 ..
    const int lr = tid / 4, lk = (tid % 4) * 4;
    const int64_t am = m0 + lr;
    const int64_t arow = (am < M) ? (QKV ? map.voxel(am) : am) : -1;
    const int wn = n0 + lr;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + lk + i;
            As[lk + i][lr] = (arow >= 0 && k < K) ? to_f32(A[arow * K + k]) : 0.f;
            Ws[lk + i][lr] = (wn < Nout && k < K) ? to_f32(Wt[(int64_t)wn * K + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[kk][ty * 4 + i];
                w[i] = Ws[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= Nout) continue;
            float v = acc[i][j] + (bias ? to_f32(bias[n]) : 0.f);
            if (QKV) {
                const int C = heads * hd;
                const int which = n / C, hh = (n % C) / hd, e = n % hd;
                if (which == 0) v *= scale;
                const int64_t win = m / map.N;
                const int tok = (int)(m % map.N);
                out[(((which * B_ + win) * heads + hh) * map.N + tok) * hd + e] = from_f32<T>(v);
            } else {
                out[m * Nout + n] = from_f32<T>(v);
            }
        }
    }
}

// ---- softmax(q k^T + bias) v for one (window, head), 128 queries per block, one query per thread -------------------
template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_core_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                        const T *__restrict__ v, const float *__restrict__ bias_t,
                                                        T *__restrict__ o, int heads, int N, int C, int KC) {
    extern __shared__ float smem[];
    float *Ks = smem;            // [KC][HD]
    float *Vs = smem + KC * HD;  // [KC][HD]
    const int wh = blockIdx.x;   // win*heads + h
    const int hh = wh % heads;
    const int64_t win = wh / heads;
    const int i = blockIdx.y * 128 + threadIdx.x;
    const bool active = i < N;
    const T *qp = q + ((int64_t)wh * N + (active ? i : 0)) * HD;
    float qr[HD], acc[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) {
        qr[e] = to_f32(qp[e]);
        acc[e] = 0.f;
    }
    float mrun = -INFINITY, lrun = 0.f;
    const float *bcol = bias_t + (int64_t)hh * N * N + (active ? i : 0);  // bias_t[h][j][i]
    for (int j0 = 0; j0 < N; j0 += KC) {
        const int kc = min(KC, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < kc * HD; t += 128) {
            Ks[t] = to_f32(k[((int64_t)wh * N + j0) * HD + t]);
            Vs[t] = to_f32(v[((int64_t)wh * N + j0) * HD + t]);
        }
        __syncthreads();
        for (int jj = 0; jj < kc; jj += 8) {
            float s[8];
            float cmax = -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = jj + u;
                float d = -INFINITY;
                if (j < kc) {
                    d = bcol[(int64_t)(j0 + j) * N];
#pragma unroll
                    for (int e = 0; e < HD; ++e) d = fmaf(qr[e], Ks[j * HD + e], d);
                }
                s[u] = d;
                cmax = fmaxf(cmax, d);
            }
            const float mnew = fmaxf(mrun, cmax);
            const float corr = __expf(mrun - mnew);  // exp(-inf) = 0 on the first group
            lrun *= corr;
#pragma unroll
            for (int e = 0; e < HD; ++e) acc[e] *= corr;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = jj + u;
                if (j < kc) {
                    const float p = __expf(s[u] - mnew);
                    lrun += p;
#pragma unroll
                    for (int e = 0; e < HD; ++e) acc[e] = fmaf(p, Vs[j * HD + e], acc[e]);
                }
            }
            mrun = mnew;
        }
    }
    if (active) {
        const float inv = 1.f / lrun;
        T *op = o + (win * N + i) * (int64_t)C + hh * HD;
#pragma unroll
        for (int e = 0; e < HD; ++e) op[e] = from_f32<T>(acc[e] * inv);
    }
}

__global__ void relpos_bias_expand_kernel(const void *table, int table_dtype, const int64_t *__restrict__ index,
                                          float *__restrict__ bias_t, int heads, int N, int table_rows) {
    // bias_t[h][j][i] = table[index[i][j]][h]
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)heads * N * N;
    if (idx >= total) return;
    const int i = (int)(idx % N);
    const int j = (int)((idx / N) % N);
    const int hh = (int)(idx / ((int64_t)N * N));
    int64_t r = index[(int64_t)i * N + j];
    r = r < 0 ? 0 : (r >= table_rows ? table_rows - 1 : r);
    float val;
    if (table_dtype == WF_F32)
        val = reinterpret_cast<const float *>(table)[r * heads + hh];
    else
        val = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(table)[r * heads + hh]);
    bias_t[idx] = val;
}

template <typename TA, typename T>
int attn_simt_forward(const TA *x, const T *qkv_w, const T *qkv_b, const T *proj_w, const T *proj_b,
                      const float *bias_t, T *out, void *workspace, int B, int D1, int H1, int W1, int C, int heads,
                      int ws, float scale, cudaStream_t st) {
    WindowMap map;
    map.D1 = D1; map.H1 = H1; map.W1 = W1; map.ws = ws; map.N = ws * ws * ws;
    map.nWy = H1 / ws; map.nWx = W1 / ws; map.nW = (D1 / ws) * map.nWy * map.nWx;
    const int hd = C / heads;
    const int64_t B_ = (int64_t)B * map.nW;
    const int64_t M = B_ * map.N;
    T *qkv = reinterpret_cast<T *>(workspace);
    T *q = qkv, *k = qkv + M * C, *v = qkv + 2 * M * C, *o = qkv + 3 * M * C;
    dim3 g1((unsigned)((M + 63) / 64), (unsigned)((3 * C + 63) / 64));
    linear_kernel<TA, T, true><<<g1, 256, 0, st>>>(x, qkv_w, qkv_b, qkv, M, 3 * C, C, map, heads, hd, scale, B_);
    WF_LAUNCH_CHECK();
    const int KC = map.N < 256 ? map.N : 256;
    const size_t smem = (size_t)2 * KC * hd * sizeof(float);
    dim3 g2((unsigned)(B_ * heads), (unsigned)((map.N + 127) / 128));
#define WF_CORE(HD_)                                                                                               \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            WF_CUDA_CHECK(cudaFuncSetAttribute(attn_core_kernel<T, HD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        attn_core_kernel<T, HD_><<<g2, 128, smem, st>>>(q, k, v, bias_t, o, heads, map.N, C, KC);                   \
    } while (0)
    switch (hd) {
        case 8: WF_CORE(8); break;
        case 16: WF_CORE(16); break;
        case 32: WF_CORE(32); break;
        case 64: WF_CORE(64); break;
        default: return WF_ERR_BAD_SHAPE;
    }
#undef WF_CORE
    WF_LAUNCH_CHECK();
    dim3 g3((unsigned)((M + 63) / 64), (unsigned)((C + 63) / 64));
    linear_kernel<T, T, false><<<g3, 256, 0, st>>>(o, proj_w, proj_b, out, M, C, C, map, heads, hd, 1.f, B_);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

template int attn_simt_forward<float, float>(const float *, const float *, const float *, const float *, const float *,
                                             const float *, float *, void *, int, int, int, int, int, int, int, float, cudaStream_t);
template int attn_simt_forward<__nv_bfloat16, __nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, const __nv_bfloat16 *,
                                                             const __nv_bfloat16 *, const __nv_bfloat16 *, const float *,
                                                             __nv_bfloat16 *, void *, int, int, int, int, int, int, int, float, cudaStream_t);
template int attn_simt_forward<float, __nv_bfloat16>(const float *, const __nv_bfloat16 *, const __nv_bfloat16 *,
                                                     const __nv_bfloat16 *, const __nv_bfloat16 *, const float *,
                                                     __nv_bfloat16 *, void *, int, int, int, int, int, int, int, float, cudaStream_t);

}  // namespace wf

extern "C" int wf_relpos_bias_expand(const void *table, int table_dtype, const int64_t *index, float *bias_t, int heads,
                                     int N, int table_rows, void *stream) {
    if (!table || !index || !bias_t) return WF_ERR_NULL_POINTER;
    if (heads <= 0 || N <= 0 || table_rows <= 0) return WF_ERR_BAD_SHAPE;
    if (table_dtype != WF_F32 && table_dtype != WF_BF16) return WF_ERR_BAD_DTYPE;
    const int64_t total = (int64_t)heads * N * N;
    wf::relpos_bias_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        table, table_dtype, index, bias_t, heads, N, table_rows);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
