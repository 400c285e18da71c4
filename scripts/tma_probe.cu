// TMA -> tcgen05 probe for the rolling-row 3^3 convolution (sm_100a): one input row [128 voxels][48 channels] (voxel stride xs) is
// fetched by ONE cp.async.bulk.tensor instruction into a 128-byte-swizzled K-major image [130 rows][128 B] (row 0 and row 129 are the
// zero halo produced by the out-of-bounds fill, channels 48..63 of every row are zero-filled too), and the tensor core reads the
// dx-shifted, k-step-advanced operand straight from it.  Checks D[r][n] = x[r + dx - 1][16 ks + n] for every (dx, ks).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/tma_probe scripts/tma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>

#include "../waveformer_b200/csrc/tc_common.cuh"

using namespace wf::tc;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap, int row, const uint8_t *b_img, float *d_out,
                                                       uint8_t *raw_out, uint16_t *y_bulk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sA = smem, *sB = smem + 32 * 1024, *sOut = smem + 40 * 1024;
    for (int i = tid; i < 17 * 1024; i += 128) sA[i] = 0xEE;     // poison: everything the MMA reads must come from the TMA
    for (int i = tid; i < 512; i += 128) sB[i] = b_img[i];
    if (tid < 32) tmem_alloc(&slot, 32);
    if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        mbar_expect_tx(&bar_tma, 130 * 128);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         smem_u32(sA)),
                     "l"(&tmap), "r"(0), "r"(-1), "r"(row), "r"(smem_u32(&bar_tma))
                     : "memory");
    }
    mbar_wait(&bar_tma, 0);
    for (int i = tid; i < 130 * 128; i += 128) raw_out[i] = sA[i];
    uint32_t phase = 0;
    for (int dx = 0; dx < 3; ++dx)
        for (int ks = 0; ks < 3; ++ks) {
            if (tid == 0) {
                const uint32_t idesc = instr_desc_h16<true>(128, 16, false);
                mma_ss(tmem, make_desc(smem_u32(sA) + dx * 128 + ks * 32, 16, 1024, 2), make_desc(smem_u32(sB), 256, 128, 0), idesc, 0u);
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, phase);
            phase ^= 1;
            tc_fence_after();
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
            tmem_wait_ld();
            for (int n = 0; n < 16; ++n) d_out[((dx * 3 + ks) * 128 + warp * 32 + lane) * 16 + n] = __uint_as_float(r[n]);
            tc_fence_before();
            __syncthreads();
        }
    // bulk store: a dense [128][48] 16-bit row staged in shared memory -> global, one instruction
    for (int i = tid; i < 128 * 48; i += 128) reinterpret_cast<uint16_t *>(sOut)[i] = (uint16_t)(i ^ 0x5a5a);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y_bulk), "r"(smem_u32(sOut)), "r"(128 * 48 * 2) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 32);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static float xval(int row, int vx, int ch) { return (float)((vx * 7 + ch * 131 + row * 17) & 2047); }

int main() {
    CK(cudaFree(0));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int ROWS = 5;
    std::vector<uint8_t> bimg(512, 0);
    for (int n = 0; n < 16; ++n) {
        __half one = __float2half(1.0f);
        memcpy(&bimg[((n / 8) * 16 + n) * 16 + (n % 8) * 2], &one, 2);
    }
    uint8_t *d_b, *d_raw; float *d_d; uint16_t *d_y;
    CK(cudaMalloc(&d_b, 512)); CK(cudaMemcpy(d_b, bimg.data(), 512, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_raw, 130 * 128)); CK(cudaMalloc(&d_d, 9 * 128 * 16 * 4)); CK(cudaMalloc(&d_y, 128 * 48 * 2));
    for (int xs : {48, 96})
        for (int variant = 0; variant < 2; ++variant) {
            // x: [ROWS][128][xs] fp16; channels >= 48 of a voxel (xs = 96) hold a different pattern that must never show up
            std::vector<__half> x((size_t)ROWS * 128 * xs);
            for (int r = 0; r < ROWS; ++r)
                for (int v = 0; v < 128; ++v)
                    for (int c = 0; c < xs; ++c) x[((size_t)r * 128 + v) * xs + c] = __float2half(c < 48 ? xval(r, v, c) : -1.0f);
            __half *d_x;
            CK(cudaMalloc(&d_x, x.size() * 2 + 256));
            CK(cudaMemcpy(d_x, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
            CUtensorMap tmap;
            // variant 0: the channel dimension is declared with its true extent 48 and the box overhangs it (64): elements 48..63 are
            // out-of-bounds zeros.  variant 1: declared extent 64 over a 96-byte voxel stride (overlapping reads of the next voxel).
            const cuuint64_t gdim[3] = {(cuuint64_t)(variant == 0 ? 48 : 64), 128, (cuuint64_t)ROWS};
            const cuuint64_t gstr[2] = {(cuuint64_t)xs * 2, (cuuint64_t)xs * 2 * 128};
            const cuuint32_t box[3] = {64, 130, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult cr = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d_x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) {
                const char *es = nullptr; cuGetErrorString(cr, &es);
                printf("xs %d variant %d: cuTensorMapEncodeTiled failed: %s\n", xs, variant, es ? es : "?");
                continue;
            }
            for (int row : {0, 2, ROWS - 1}) {
                CK(cudaMemset(d_d, 0xff, 9 * 128 * 16 * 4));
                CK(cudaMemset(d_y, 0, 128 * 48 * 2));
                probe_kernel<<<1, 128, 64 * 1024>>>(tmap, row, d_b, d_d, d_raw, d_y);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("xs %d variant %d row %d: CUDA error %s\n", xs, variant, row, cudaGetErrorString(e)); return 1; }
                std::vector<float> d(9 * 128 * 16);
                CK(cudaMemcpy(d.data(), d_d, d.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0, first = -1;
                for (int dx = 0; dx < 3; ++dx)
                    for (int ks = 0; ks < 3; ++ks)
                        for (int r = 0; r < 128; ++r)
                            for (int n = 0; n < 16; ++n) {
                                const int v = r + dx - 1;
                                const float want = (v < 0 || v > 127) ? 0.0f : xval(row, v, ks * 16 + n);
                                const int idx = ((dx * 3 + ks) * 128 + r) * 16 + n;
                                if (d[idx] != want) { if (first < 0) first = idx; ++bad; }
                            }
                std::vector<uint16_t> yb(128 * 48);
                CK(cudaMemcpy(yb.data(), d_y, yb.size() * 2, cudaMemcpyDeviceToHost));
                int ybad = 0;
                for (int i = 0; i < 128 * 48; ++i) ybad += yb[i] != (uint16_t)(i ^ 0x5a5a);
                printf("xs %d variant %d row %d: MMA over the TMA image %s (%d wrong", xs, variant, row, bad ? "MISMATCH" : "ok", bad);
                if (bad) printf("; first at dx %d ks %d r %d n %d got %.0f", first / (3 * 128 * 16), (first / (128 * 16)) % 3, (first / 16) % 128, first % 16, d[first]);
                printf("), bulk store %s\n", ybad ? "MISMATCH" : "ok");
                if (bad && row == 0) {
                    std::vector<uint8_t> raw(130 * 128);
                    CK(cudaMemcpy(raw.data(), d_raw, raw.size(), cudaMemcpyDeviceToHost));
                    for (int r = 0; r < 4; ++r) {
                        printf("  smem row %d:", r);
                        for (int c = 0; c < 64; c += 8) { __half h; memcpy(&h, &raw[r * 128 + c * 2], 2); printf(" %6.0f", __half2float(h)); }
                        printf("\n");
                    }
                }
            }
            CK(cudaFree(d_x));
        }
    return 0;
}
