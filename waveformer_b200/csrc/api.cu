// C-ABI glue: version / errors and the window-attention dispatcher (see include/waveformer_b200.h).
#include <stdlib.h>

#include "wf_common.cuh"

namespace wf {
int g_last_cuda_error = 0;

template <typename TA, typename T>
int attn_simt_forward(const TA *x, const T *qkv_w, const T *qkv_b, const T *proj_w, const T *proj_b,
                      const float *bias_t, T *out, void *workspace, int B, int D1, int H1, int W1, int C, int heads,
                      int ws, float scale, cudaStream_t st);
bool attn_tc_supported(int D1, int H1, int W1, int C, int heads, int ws);
size_t attn_tc_bias_image_bytes(int heads);
int attn_tc_bias_image(const void *table, int table_dtype, const int64_t *index, void *img, bool f16, int heads,
                       int table_rows, cudaStream_t st);
size_t attn_tc_split_workspace_bytes(int64_t tokens, int C);
int attn_tc_forward_split(const float *x, const uint16_t *qkv_w_hi, const uint16_t *qkv_w_lo, const float *qkv_b,
                          const uint16_t *proj_w_hi, const uint16_t *proj_w_lo, const float *proj_b, const uint16_t *bias_img,
                          float *out, void *workspace, int B, int D1, int H1, int W1, int C, int heads, float scale,
                          cudaStream_t st);
int attn_tc_forward(const void *x, int x_dtype, const void *qkv_w, const void *qkv_b, const void *proj_w,
                    const void *proj_b, const void *bias_img, void *out, bool out_f32, void *workspace, bool f16, int B,
                    int D1, int H1, int W1, int C, int heads, float scale, cudaStream_t st);
}  // namespace wf

extern "C" const char *wf_version(void) { return "waveformer_b200 0.1.0 (sm_100a)"; }

extern "C" int wf_last_cuda_error(void) { return wf::g_last_cuda_error; }

extern "C" const char *wf_error_string(int status) {
    switch (status) {
        case WF_OK: return "ok";
        case WF_ERR_BAD_DTYPE: return "unsupported dtype or dtype combination (WF_F32, WF_BF16, WF_F16)";
        case WF_ERR_BAD_SHAPE: return "bad shape (odd extent, empty tensor, window does not tile the grid, or unsupported head_dim)";
        case WF_ERR_NULL_POINTER: return "null pointer";
        case WF_ERR_MISALIGNED: return "pointer or stride not 16-byte aligned";
        case WF_ERR_CUDA: return cudaGetErrorString((cudaError_t)wf::g_last_cuda_error);
        case WF_ERR_WORKSPACE: return "workspace too small";
        case WF_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

static int check_attn_shape(int B, int D1, int H1, int W1, int C, int heads, int ws) {
    if (B <= 0 || C <= 0 || heads <= 0 || ws <= 0 || D1 <= 0 || H1 <= 0 || W1 <= 0) return WF_ERR_BAD_SHAPE;
    if (C % heads) return WF_ERR_BAD_SHAPE;
    if (D1 % ws || H1 % ws || W1 % ws) return WF_ERR_BAD_SHAPE;
    const int hd = C / heads;
    if (hd != 8 && hd != 16 && hd != 32 && hd != 64) return WF_ERR_BAD_SHAPE;
    return WF_OK;
}

extern "C" size_t wf_window_attn_workspace_bytes(int dtype, int B, int D1, int H1, int W1, int C, int heads, int ws) {
    if (check_attn_shape(B, D1, H1, W1, C, heads, ws) != WF_OK) return 0;
    const size_t e = dtype == WF_F32 ? 4 : 2;
    const size_t tokens = (size_t)B * D1 * H1 * W1;
    return 4 * tokens * (size_t)C * e + 256;  // q, k, v (head-major) and the pre-projection output
}

extern "C" size_t wf_relpos_bias_image_bytes(int heads, int N) {
    if (heads <= 0 || N != 512) return 0;
    return wf::attn_tc_bias_image_bytes(heads);
}

extern "C" int wf_relpos_bias_image(const void *table, int table_dtype, const int64_t *index, void *img, int fmt,
                                    int heads, int N, int table_rows, void *stream) {
    if (!table || !index || !img) return WF_ERR_NULL_POINTER;
    if (heads <= 0 || table_rows <= 0) return WF_ERR_BAD_SHAPE;
    if (N != 512) return WF_ERR_UNSUPPORTED;
    if ((table_dtype != WF_F32 && table_dtype != WF_BF16) || (fmt != WF_BF16 && fmt != WF_F16)) return WF_ERR_BAD_DTYPE;
    return wf::attn_tc_bias_image(table, table_dtype, index, img, fmt == WF_F16, heads, table_rows, (cudaStream_t)stream);
}

extern "C" int wf_window_attn_tc_supported(int D1, int H1, int W1, int C, int heads, int ws) {
    return check_attn_shape(1, D1, H1, W1, C, heads, ws) == WF_OK && wf::attn_tc_supported(D1, H1, W1, C, heads, ws) ? 1 : 0;
}

extern "C" size_t wf_window_attn_split_workspace_bytes(int B, int D1, int H1, int W1, int C, int heads, int ws) {
    if (check_attn_shape(B, D1, H1, W1, C, heads, ws) != WF_OK || !wf::attn_tc_supported(D1, H1, W1, C, heads, ws)) return 0;
    return wf::attn_tc_split_workspace_bytes((int64_t)B * D1 * H1 * W1, C);
}

extern "C" int wf_window_attn_fwd_split(const float *x, const void *qkv_w_hi, const void *qkv_w_lo, const float *qkv_b,
                                        const void *proj_w_hi, const void *proj_w_lo, const float *proj_b,
                                        const void *bias_img, float *out, void *workspace, size_t workspace_bytes, int B,
                                        int D1, int H1, int W1, int C, int heads, int ws, float scale, void *stream) {
    if (!x || !qkv_w_hi || !qkv_w_lo || !qkv_b || !proj_w_hi || !proj_w_lo || !proj_b || !bias_img || !out || !workspace)
        return WF_ERR_NULL_POINTER;
    const int rc = check_attn_shape(B, D1, H1, W1, C, heads, ws);
    if (rc != WF_OK) return rc;
    if (!wf::attn_tc_supported(D1, H1, W1, C, heads, ws)) return WF_ERR_UNSUPPORTED;
    if (workspace_bytes < wf_window_attn_split_workspace_bytes(B, D1, H1, W1, C, heads, ws)) return WF_ERR_WORKSPACE;
    using u16 = uint16_t;
    return wf::attn_tc_forward_split(x, (const u16 *)qkv_w_hi, (const u16 *)qkv_w_lo, qkv_b, (const u16 *)proj_w_hi,
                                     (const u16 *)proj_w_lo, proj_b, (const u16 *)bias_img, out, workspace, B, D1, H1, W1, C,
                                     heads, scale, (cudaStream_t)stream);
}

extern "C" int wf_window_attn_fwd(const void *x, int x_dtype, const void *qkv_w, const void *qkv_b, const void *proj_w,
                                  const void *proj_b, const float *bias_t, const void *bias_img, void *out,
                                  int out_dtype, void *workspace, size_t workspace_bytes, int dtype, int B, int D1,
                                  int H1, int W1, int C, int heads, int ws, float scale, void *stream) {
    if (!x || !qkv_w || !proj_w || !proj_b || !out || !workspace) return WF_ERR_NULL_POINTER;
    const int rc = check_attn_shape(B, D1, H1, W1, C, heads, ws);
    if (rc != WF_OK) return rc;
    if ((dtype != WF_F32 && dtype != WF_BF16 && dtype != WF_F16) || (x_dtype != WF_F32 && x_dtype != WF_BF16))
        return WF_ERR_BAD_DTYPE;
    if (dtype == WF_F32 && x_dtype != WF_F32) return WF_ERR_BAD_DTYPE;
    if (out_dtype != dtype && out_dtype != WF_F32) return WF_ERR_BAD_DTYPE;
    if (workspace_bytes < wf_window_attn_workspace_bytes(dtype, B, D1, H1, W1, C, heads, ws)) return WF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    if (dtype == WF_F32) {
        if (!bias_t) return WF_ERR_NULL_POINTER;
        return wf::attn_simt_forward<float, float>((const float *)x, (const float *)qkv_w, (const float *)qkv_b,
                                                   (const float *)proj_w, (const float *)proj_b, bias_t, (float *)out,
                                                   workspace, B, D1, H1, W1, C, heads, ws, scale, st);
    }
    // 16-bit operands: tcgen05 / TMEM path for the reference geometry (512-token windows, head_dim 16) when the dense
    // bias image is supplied; WF_ATTN_IMPL=simt forces the CUDA-core kernels (the tests cross-check the two on the device)
    const char *impl = getenv("WF_ATTN_IMPL");
    const bool force_simt = impl != nullptr && impl[0] == 's';
    const bool tc_ok = bias_img != nullptr && qkv_b != nullptr && wf::attn_tc_supported(D1, H1, W1, C, heads, ws);
    if (dtype == WF_F16 && (!tc_ok || force_simt)) return WF_ERR_UNSUPPORTED;  // fp16 exists as a tensor-core operand format only
    if (tc_ok && !force_simt)
        return wf::attn_tc_forward(x, x_dtype, qkv_w, qkv_b, proj_w, proj_b, bias_img, out, out_dtype == WF_F32, workspace,
                                   dtype == WF_F16, B, D1, H1, W1, C, heads, scale, st);
    if (!bias_t) return WF_ERR_NULL_POINTER;
    if (out_dtype != WF_BF16) return WF_ERR_UNSUPPORTED;
    if (x_dtype == WF_F32)
        return wf::attn_simt_forward<float, bf>((const float *)x, (const bf *)qkv_w, (const bf *)qkv_b, (const bf *)proj_w,
                                                (const bf *)proj_b, bias_t, (bf *)out, workspace, B, D1, H1, W1, C, heads,
                                                ws, scale, st);
    return wf::attn_simt_forward<bf, bf>((const bf *)x, (const bf *)qkv_w, (const bf *)qkv_b, (const bf *)proj_w,
                                         (const bf *)proj_b, bias_t, (bf *)out, workspace, B, D1, H1, W1, C, heads, ws,
                                         scale, st);
}
