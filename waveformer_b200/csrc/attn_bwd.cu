// Kernel group 2, gradient of the window-attention core (training, BASELINE config 5), fp32 on the CUDA cores.
//
// Differentiates what Attention.forward computes between its two Linear layers (reference network_models/attention.py:
// 87-101): S = (scale q) k^T + table[index], P = softmax(S), O = P v.  Given dO it produces the gradient of the qkv
// Linear's OUTPUT (d_qkv[M, 3C], row = window-major token, ready for the two library GEMMs of that Linear's backward) and
// the gradient of relative_position_bias_table.  Probabilities are never stored: both kernels recompute S from q, k and
// the bias, flash-attention style, so the working set per (window, head) is N x head_dim, not N x N.
//
//   row kernel    one thread per query i:   lse_i = logsumexp_j S_ij,  delta_i = sum_e dO_ie O_ie,
//                                           dq_i  = sum_j dS_ij k_j                  with dS_ij = P_ij (dP_ij - delta_i),
//   column kernel one thread per key j:     dk_j  = sum_i dS_ij q_i,   dv_j = sum_i P_ij dO_i,        dP_ij = dO_i . v_j
//                                           d_table[index[i][j]][h] += dS_ij   (shared-memory histogram per CTA, one
//                                           global atomic per touched table row per CTA)
// q in the workspace is already multiplied by `scale` (attention.py:88), so dk uses it as is and dq is scaled on the way
// out.  The two kernels are separate launches ordered by the stream (the column kernel reads lse / delta of every row).
#include "wf_common.cuh"

namespace wf {

template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_row_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                           const float *__restrict__ v, const float *__restrict__ bias_t,
                                                           const float *__restrict__ o, const float *__restrict__ d_o,
                                                           float *__restrict__ d_qkv, float *__restrict__ lse,
                                                           float *__restrict__ delta, int heads, int N, int C, int KC,
                                                           float scale) {
    extern __shared__ float smem[];
    float *Ks = smem;            // [KC][HD]
    float *Vs = smem + KC * HD;  // [KC][HD]
    const int wh = blockIdx.x;   // win * heads + h
    const int hh = wh % heads;
    const int64_t win = wh / heads;
    const int i = blockIdx.y * 128 + threadIdx.x;
    const bool active = i < N;
    const int ia = active ? i : 0;
    const float *qp = q + ((int64_t)wh * N + ia) * HD;
    const int64_t row = (win * N + ia) * (int64_t)C + hh * HD;   // this head's slice of token (win, i) in [M, C] buffers
    float qr[HD], dor[HD], dq[HD];
    float dl = 0.f;
#pragma unroll
    for (int e = 0; e < HD; ++e) {
        qr[e] = qp[e];
        dor[e] = d_o[row + e];
        dl = fmaf(dor[e], o[row + e], dl);
        dq[e] = 0.f;
    }
    const float *bcol = bias_t + (int64_t)hh * N * N + ia;   // bias_t[h][j][i]
    // pass 1: log-sum-exp of the score row
    float mrun = -INFINITY, lrun = 0.f;
    for (int j0 = 0; j0 < N; j0 += KC) {
        const int kc = min(KC, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < kc * HD; t += 128) Ks[t] = k[((int64_t)wh * N + j0) * HD + t];
        __syncthreads();
        for (int j = 0; j < kc; ++j) {
            float s = bcol[(int64_t)(j0 + j) * N];
#pragma unroll
            for (int e = 0; e < HD; ++e) s = fmaf(qr[e], Ks[j * HD + e], s);
            const float mnew = fmaxf(mrun, s);
            lrun = lrun * __expf(mrun - mnew) + __expf(s - mnew);
            mrun = mnew;
        }
    }
    const float l = mrun + __logf(lrun);
    // pass 2: dq
    for (int j0 = 0; j0 < N; j0 += KC) {
        const int kc = min(KC, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < kc * HD; t += 128) {
            Ks[t] = k[((int64_t)wh * N + j0) * HD + t];
            Vs[t] = v[((int64_t)wh * N + j0) * HD + t];
        }
        __syncthreads();
        for (int j = 0; j < kc; ++j) {
            float s = bcol[(int64_t)(j0 + j) * N], dp = 0.f;
#pragma unroll
            for (int e = 0; e < HD; ++e) {
                s = fmaf(qr[e], Ks[j * HD + e], s);
                dp = fmaf(dor[e], Vs[j * HD + e], dp);
            }
            const float ds = __expf(s - l) * (dp - dl);
#pragma unroll
            for (int e = 0; e < HD; ++e) dq[e] = fmaf(ds, Ks[j * HD + e], dq[e]);
        }
    }
    if (active) {
        float *dst = d_qkv + (win * N + i) * (int64_t)(3 * C) + hh * HD;
#pragma unroll
        for (int e = 0; e < HD; ++e) dst[e] = dq[e] * scale;
        lse[(int64_t)wh * N + i] = l;
        delta[(int64_t)wh * N + i] = dl;
    }
}

template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_col_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                           const float *__restrict__ v, const float *__restrict__ table,
                                                           const int64_t *__restrict__ index,
                                                           const float *__restrict__ d_o, const float *__restrict__ lse,
                                                           const float *__restrict__ delta, float *__restrict__ d_qkv,
                                                           float *__restrict__ d_table, int heads, int N, int C, int QC,
                                                           int table_rows) {
    extern __shared__ float smem[];
    float *Qs = smem;                    // [QC][HD]
    float *Gs = Qs + QC * HD;            // [QC][HD] dO rows
    float *Ls = Gs + QC * HD;            // [QC] lse
    float *Ds = Ls + QC;                 // [QC] delta
    float *tab = Ds + QC;                // [table_rows] this head's bias column
    float *dtab = tab + table_rows;      // [table_rows] its gradient, accumulated by this CTA
    const int wh = blockIdx.x;
    const int hh = wh % heads;
    const int64_t win = wh / heads;
    const int j = blockIdx.y * 128 + threadIdx.x;
    const bool active = j < N;
    const int ja = active ? j : 0;
    for (int r = threadIdx.x; r < table_rows; r += 128) {
        tab[r] = table[(int64_t)r * heads + hh];
        dtab[r] = 0.f;
    }
    float kr[HD], vr[HD], dk[HD], dv[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) {
        kr[e] = k[((int64_t)wh * N + ja) * HD + e];
        vr[e] = v[((int64_t)wh * N + ja) * HD + e];
        dk[e] = dv[e] = 0.f;
    }
    for (int i0 = 0; i0 < N; i0 += QC) {
        const int qc = min(QC, N - i0);
        __syncthreads();
        for (int t = threadIdx.x; t < qc * HD; t += 128) {
            const int ii = t / HD, e = t % HD;
            Qs[t] = q[((int64_t)wh * N + i0) * HD + t];
            Gs[t] = d_o[(win * N + i0 + ii) * (int64_t)C + hh * HD + e];
        }
        for (int t = threadIdx.x; t < qc; t += 128) {
            Ls[t] = lse[(int64_t)wh * N + i0 + t];
            Ds[t] = delta[(int64_t)wh * N + i0 + t];
        }
        __syncthreads();
        if (active) {
            for (int ii = 0; ii < qc; ++ii) {
                int64_t r = index[(int64_t)(i0 + ii) * N + j];
                r = r < 0 ? 0 : (r >= table_rows ? table_rows - 1 : r);
                float s = tab[r], dp = 0.f;
#pragma unroll
                for (int e = 0; e < HD; ++e) {
                    s = fmaf(Qs[ii * HD + e], kr[e], s);
                    dp = fmaf(Gs[ii * HD + e], vr[e], dp);
                }
                const float p = __expf(s - Ls[ii]);
                const float ds = p * (dp - Ds[ii]);
#pragma unroll
                for (int e = 0; e < HD; ++e) {
                    dk[e] = fmaf(ds, Qs[ii * HD + e], dk[e]);
                    dv[e] = fmaf(p, Gs[ii * HD + e], dv[e]);
                }
                atomicAdd(&dtab[r], ds);   // keys of one warp differ in (dy, dx): distinct rows, no intra-warp conflict
            }
        }
    }
    if (active) {
        float *dst = d_qkv + (win * N + j) * (int64_t)(3 * C) + hh * HD;
#pragma unroll
        for (int e = 0; e < HD; ++e) {
            dst[C + e] = dk[e];
            dst[2 * C + e] = dv[e];
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < table_rows; r += 128) {
        const float g = dtab[r];
        if (g != 0.f) atomicAdd(&d_table[(int64_t)r * heads + hh], g);
    }
}

template <int HD>
static int attn_bwd_launch(const float *ws, const float *bias_t, const float *table, const int64_t *index,
                           const float *d_o, float *d_qkv, float *d_table, float *stats, int64_t B_, int N, int C,
                           int heads, int table_rows, float scale, cudaStream_t st) {
    const int64_t M = B_ * N;
    const float *q = ws, *k = ws + M * C, *v = ws + 2 * M * C, *o = ws + 3 * M * C;
    float *lse = stats, *delta = stats + B_ * heads * N;
    const int KC = N < 256 ? N : 256, QC = N < 128 ? N : 128;
    const size_t smem_row = (size_t)2 * KC * HD * sizeof(float);
    const size_t smem_col = ((size_t)2 * QC * HD + 2 * QC + 2 * table_rows) * sizeof(float);
    if (smem_row > 200 * 1024 || smem_col > 200 * 1024) return WF_ERR_UNSUPPORTED;
    if (smem_row > 48 * 1024)
        WF_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_row_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
    if (smem_col > 48 * 1024)
        WF_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_col_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
    dim3 grid((unsigned)(B_ * heads), (unsigned)((N + 127) / 128));
    attn_bwd_row_kernel<HD><<<grid, 128, smem_row, st>>>(q, k, v, bias_t, o, d_o, d_qkv, lse, delta, heads, N, C, KC, scale);
    WF_LAUNCH_CHECK();
    attn_bwd_col_kernel<HD><<<grid, 128, smem_col, st>>>(q, k, v, table, index, d_o, lse, delta, d_qkv, d_table, heads, N,
                                                         C, QC, table_rows);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" size_t wf_window_attn_bwd_stats_bytes(int64_t windows, int N, int heads) {
    if (windows <= 0 || N <= 0 || heads <= 0) return 0;
    return (size_t)2 * windows * heads * N * sizeof(float);
}

extern "C" int wf_window_attn_bwd(const float *workspace, const float *bias_t, const float *table, const int64_t *index,
                                  const float *d_o, float *d_qkv, float *d_table, float *stats, int64_t windows, int N,
                                  int C, int heads, int table_rows, float scale, void *stream) {
    if (!workspace || !bias_t || !table || !index || !d_o || !d_qkv || !d_table || !stats) return WF_ERR_NULL_POINTER;
    if (windows <= 0 || N <= 0 || C <= 0 || heads <= 0 || table_rows <= 0 || C % heads != 0) return WF_ERR_BAD_SHAPE;
    if (windows * heads > 0x7fffffffLL) return WF_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
#define WF_BWD(HD_) \
    return wf::attn_bwd_launch<HD_>(workspace, bias_t, table, index, d_o, d_qkv, d_table, stats, windows, N, C, heads, table_rows, scale, st)
    switch (C / heads) {
        case 8: WF_BWD(8);
        case 16: WF_BWD(16);
        case 32: WF_BWD(32);
        case 64: WF_BWD(64);
        default: return WF_ERR_BAD_SHAPE;
    }
#undef WF_BWD
}
