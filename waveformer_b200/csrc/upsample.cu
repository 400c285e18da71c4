// Trilinear upsampling of channels-last activations, fused with the multi-level sum and the shortcut add of
// Block.multi_scale_forward (sm_100a).
//
// Reference: `attn_fused = attn_fused + F.interpolate(attn_windows, size=(D,H,W), mode='trilinear')` per level and
// `attn_fused = shortcut + attn_fused` (network_models/wave_helper.py:500-508), i.e. up to three upsampled coarse maps
// summed in level order and added to the block input; and nn.Upsample(scale_factor=s, mode='trilinear',
// align_corners=True) in ProjectionUpsample (wave_helper.py:42,63).  Index arithmetic follows PyTorch's
// upsample_trilinear3d (area_pixel_compute_source_index): align_corners=False: src = max(0, (dst+0.5)*in/out - 0.5);
// align_corners=True: src = dst*(in-1)/(out-1); i0 = floor(src), i1 = min(i0+1, in-1).
// The coarse sources are small (<= 1/8 of the output) and stay in L1/L2; the kernel streams `base` in and `y` out once.
#include "wf_common.cuh"

namespace wf {

struct UpSrc {
    const void *ptr;
    int d, h, w;
    float sz, sy, sx;  // source-coordinate scale per axis
};
struct UpArgs {
    UpSrc src[3];
    int nsrc;
    int align;
};

template <typename T, int N> __device__ __forceinline__ void load_n(const T *p, float (&v)[N]) {
    if constexpr (N == 1) {
        v[0] = to_f32(*p);
    } else if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = reinterpret_cast<const float4 *>(p)[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            float f[4];
            load4<T>(p + 4 * i, f);
            v[4 * i] = f[0]; v[4 * i + 1] = f[1]; v[4 * i + 2] = f[2]; v[4 * i + 3] = f[3];
        }
    }
}
template <typename T, int N> __device__ __forceinline__ void store_n(T *p, const float (&v)[N]) {
    if constexpr (N == 1) {
        *p = from_f32<T>(v[0]);
    } else if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float f[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
            store4<T>(p + 4 * i, f);
        }
    }
}

__device__ __forceinline__ void src_index(int dst, float scale, int in, int align, int &i0, int &i1, float &l1) {
    float s = align ? scale * dst : fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    if (l1 < 0.f) l1 = 0.f;
    if (l1 > 1.f) l1 = 1.f;
}

// grid: x = ceil(W * cvecs / 256) (threads flat over (x, channel packet) of one row), y = H, z = B * D: the z / y source
// indices and weights depend on blockIdx only (uniform datapath); a thread pays one division and the x-axis arithmetic.
template <typename TS, typename TB, typename TO, int VEC>
__global__ void __launch_bounds__(256) upsample_add_kernel(UpArgs a, const TB *__restrict__ base, TO *__restrict__ y,
                                                           int64_t total, int D, int H, int W, int C, int cvecs,
                                                           int64_t bs, int64_t ys) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= W * cvecs) return;
    const int xx = i / cvecs, cv = i - xx * cvecs;
    const int yy = blockIdx.y;
    const int zz = blockIdx.z % D;
    const int64_t b = blockIdx.z / D;
    const int c0 = cv * VEC;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int s = 0; s < a.nsrc; ++s) {
        const UpSrc &u = a.src[s];
        int z0, z1, y0, y1, x0, x1;
        float lz, ly, lx;
        src_index(zz, u.sz, u.d, a.align, z0, z1, lz);
        src_index(yy, u.sy, u.h, a.align, y0, y1, ly);
        src_index(xx, u.sx, u.w, a.align, x0, x1, lx);
        const TS *p = reinterpret_cast<const TS *>(u.ptr) + (int64_t)b * u.d * u.h * u.w * C + c0;
        float lvl[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) lvl[e] = 0.f;
        // same association as PyTorch: sum over the 8 corners of w_z * w_y * w_x * value
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int zi = (k & 4) ? z1 : z0, yi = (k & 2) ? y1 : y0, xi = (k & 1) ? x1 : x0;
            const float wgt = ((k & 4) ? lz : 1.f - lz) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
            float f[VEC];
            const TS *q = p + (((int64_t)zi * u.h + yi) * u.w + xi) * C;
            if constexpr (VEC == 1) {
                f[0] = to_f32(__ldg(q));
            } else {
                Pack<TS>::unpack(__ldg(reinterpret_cast<const typename Pack<TS>::raw *>(q)), f);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) lvl[e] = fmaf(wgt, f[e], lvl[e]);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += lvl[e];
    }
    const int64_t vox = ((b * D + zz) * H + yy) * (int64_t)W + xx;
    if (base != nullptr) {
        float bv[VEC];
        load_n<TB, VEC>(base + vox * bs + c0, bv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = bv[e] + acc[e];
    }
    store_n<TO, VEC>(y + vox * ys + c0, acc);
}

// ---- cell variant: one thread = a 2 x 2 x 2 block of output voxels x 4 channels -------------------------------------
// For upscaling by >= 2 (source step per output voxel <= 0.5) the two outputs of an aligned pair read at most three
// distinct source points per axis, so the cell needs 3 x 3 x 3 source loads per level instead of 8 x 8, and the index
// arithmetic is paid once per 8 outputs.  Interpolation is separable: x pass (3 -> 2), y pass (3 -> 2), z pass (3 -> 2).
struct AxisPair {
    int p[3];        // source points
    float wa[3];     // weights of output 2k on the points
    float wb[3];     // weights of output 2k + 1
};

__device__ __forceinline__ AxisPair axis_pair(int o, float scale, int in, int align) {
    int i0a, i1a, i0b, i1b;
    float la, lb;
    src_index(o, scale, in, align, i0a, i1a, la);
    src_index(o + 1, scale, in, align, i0b, i1b, lb);
    AxisPair r;
    r.p[0] = i0a; r.p[1] = i1a; r.p[2] = i1b;
    // weights land on the FIRST slot holding the same source index (clamped borders repeat indices); static indexing only
    const bool a_same = i1a == i0a;
    r.wa[0] = (1.f - la) + (a_same ? la : 0.f);
    r.wa[1] = a_same ? 0.f : la;
    r.wa[2] = 0.f;
    const int s0 = i0b == r.p[0] ? 0 : (i0b == r.p[1] ? 1 : 2);
    const int s1 = i1b == r.p[0] ? 0 : (i1b == r.p[1] ? 1 : 2);
#pragma unroll
    for (int k = 0; k < 3; ++k) r.wb[k] = (s0 == k ? 1.f - lb : 0.f) + (s1 == k ? lb : 0.f);
    return r;
}


// grid: x = ceil((W / 2) * (C / 4) / 128), y = H / 2, z = B * D / 2; 128 threads (~148 registers each)
template <typename TS, typename TB, typename TO>
__global__ void __launch_bounds__(128) upsample_cell_kernel(UpArgs a, const TB *__restrict__ base, TO *__restrict__ y, int D,
                                                            int H, int W, int C, int c4s, int64_t bs, int64_t ys) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= (W >> 1) * c4s) return;
    const int xc = i / c4s, cv = i - xc * c4s;
    const int yc = blockIdx.y;
    const int zc = blockIdx.z % (D >> 1);
    const int64_t b = blockIdx.z / (D >> 1);
    const int c0 = cv * 4;
    float out[2][2][2][4];
#pragma unroll
    for (int zo = 0; zo < 2; ++zo)
#pragma unroll
        for (int yo = 0; yo < 2; ++yo)
#pragma unroll
            for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                for (int e = 0; e < 4; ++e) out[zo][yo][xo][e] = 0.f;
#pragma unroll
    for (int s = 0; s < 3; ++s) {   // static index into the by-value argument struct (a dynamic one would copy it to local memory)
        if (s >= a.nsrc) break;
        const UpSrc &u = a.src[s];
        const AxisPair az = axis_pair(2 * zc, u.sz, u.d, a.align);
        const AxisPair ay = axis_pair(2 * yc, u.sy, u.h, a.align);
        const AxisPair ax = axis_pair(2 * xc, u.sx, u.w, a.align);
        const TS *p = reinterpret_cast<const TS *>(u.ptr) + (int64_t)b * u.d * u.h * u.w * C + c0;
#pragma unroll
        for (int iz = 0; iz < 3; ++iz) {
            float ty[2][2][4];   // [y output][x output][channel] for this z point
#pragma unroll
            for (int yo = 0; yo < 2; ++yo)
#pragma unroll
                for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                    for (int e = 0; e < 4; ++e) ty[yo][xo][e] = 0.f;
#pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
                const TS *row = p + ((int64_t)az.p[iz] * u.h + ay.p[iy]) * u.w * C;
                float v[3][4];
#pragma unroll
                for (int ix = 0; ix < 3; ++ix) load4<TS>(row + (int64_t)ax.p[ix] * C, v[ix]);
                float tx[2][4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    tx[0][e] = fmaf(ax.wa[2], v[2][e], fmaf(ax.wa[1], v[1][e], ax.wa[0] * v[0][e]));
                    tx[1][e] = fmaf(ax.wb[2], v[2][e], fmaf(ax.wb[1], v[1][e], ax.wb[0] * v[0][e]));
                }
#pragma unroll
                for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ty[0][xo][e] = fmaf(ay.wa[iy], tx[xo][e], ty[0][xo][e]);
                        ty[1][xo][e] = fmaf(ay.wb[iy], tx[xo][e], ty[1][xo][e]);
                    }
            }
#pragma unroll
            for (int yo = 0; yo < 2; ++yo)
#pragma unroll
                for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        out[0][yo][xo][e] = fmaf(az.wa[iz], ty[yo][xo][e], out[0][yo][xo][e]);
                        out[1][yo][xo][e] = fmaf(az.wb[iz], ty[yo][xo][e], out[1][yo][xo][e]);
                    }
        }
    }
#pragma unroll
    for (int zo = 0; zo < 2; ++zo)
#pragma unroll
        for (int yo = 0; yo < 2; ++yo)
#pragma unroll
            for (int xo = 0; xo < 2; ++xo) {
                const int64_t vox = ((b * D + 2 * zc + zo) * H + 2 * yc + yo) * (int64_t)W + 2 * xc + xo;
                if (base != nullptr) {
                    float bv[4];
                    load_n<TB, 4>(base + vox * bs + c0, bv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) out[zo][yo][xo][e] = bv[e] + out[zo][yo][xo][e];
                }
                store_n<TO, 4>(y + vox * ys + c0, out[zo][yo][xo]);
            }
}

// ---- z-walking cell variant: plane rings in shared memory -----------------------------------------------------------------------
// The cell kernel above re-interpolates all three source planes of every cell and level (27 loads, ~450 FMAs per level and 32 outputs
// x 4 channels) and is bound by instruction issue (SM busy 70-81 % in ncu).  Consecutive cells along z share source planes, so a
// thread walks ZC cells and keeps the x/y-interpolated planes ([2 y][2 x][4 channels] each) of every level in a ring in shared memory:
// plane p of level s in slot p % 3, as four 16-byte words per thread in thread-private columns ([level][slot][word][thread]:
// conflict-free, no barrier needed).  The planes a cell needs are consecutive source indices, so they never collide in the ring; a
// plane is interpolated only when its slot does not hold it yet - every cell at scale 2, every second at scale 4, every fourth at
// scale 8 - and the z pass reads three slots per level.  The z coordinate is uniform over the block: uniform control flow.  Same
// operations in the same order per output as the cell kernel: bit-identical results.  (Register caches instead of the ring were
// measured first: one level 329 us against the ring's 246 us and the cell kernel's 399 us - the slot rotation costs ~100 register
// moves per cell; two / three levels need 254 / 255 registers + spills and were no faster / slower than the cell kernel.)
template <typename TS, typename TB, typename TO, int ZC, int NSRC>
__global__ void __launch_bounds__(128) upsample_cell_ring_kernel(UpArgs a, const TB *__restrict__ base, TO *__restrict__ y, int D,
                                                                 int H, int W, int C, int c4s, int64_t bs, int64_t ys) {
    extern __shared__ __align__(16) float4 ring[];       // [NSRC][3 slots][4 words][128 threads]
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= (W >> 1) * c4s) return;                     // (no block-wide barrier below)
    const int xc = i / c4s, cv = i - xc * c4s;
    const int yc = blockIdx.y;
    const int nchunk = ((D >> 1) + ZC - 1) / ZC;
    const int zchunk = blockIdx.z % nchunk;
    const int64_t b = blockIdx.z / nchunk;
    const int c0 = cv * 4;
    int have[NSRC][3];
#pragma unroll
    for (int s = 0; s < NSRC; ++s) have[s][0] = have[s][1] = have[s][2] = -1;
    auto plane = [&](const UpSrc &u, int zs, float4 *dst) {      // dst: this thread's column of the slot (stride 128 float4 per word)
        const AxisPair ay = axis_pair(2 * yc, u.sy, u.h, a.align);
        const AxisPair ax = axis_pair(2 * xc, u.sx, u.w, a.align);
        const TS *p = reinterpret_cast<const TS *>(u.ptr) + (int64_t)b * u.d * u.h * u.w * C + c0;
        float ty[2][2][4];
#pragma unroll
        for (int yo = 0; yo < 2; ++yo)
#pragma unroll
            for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                for (int e = 0; e < 4; ++e) ty[yo][xo][e] = 0.f;
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const TS *row = p + ((int64_t)zs * u.h + ay.p[iy]) * u.w * C;
            float v[3][4];
#pragma unroll
            for (int ix = 0; ix < 3; ++ix) load4<TS>(row + (int64_t)ax.p[ix] * C, v[ix]);
            float tx[2][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                tx[0][e] = fmaf(ax.wa[2], v[2][e], fmaf(ax.wa[1], v[1][e], ax.wa[0] * v[0][e]));
                tx[1][e] = fmaf(ax.wb[2], v[2][e], fmaf(ax.wb[1], v[1][e], ax.wb[0] * v[0][e]));
            }
#pragma unroll
            for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ty[0][xo][e] = fmaf(ay.wa[iy], tx[xo][e], ty[0][xo][e]);
                    ty[1][xo][e] = fmaf(ay.wb[iy], tx[xo][e], ty[1][xo][e]);
                }
        }
#pragma unroll
        for (int yo = 0; yo < 2; ++yo)
#pragma unroll
            for (int xo = 0; xo < 2; ++xo) dst[(yo * 2 + xo) * 128] = make_float4(ty[yo][xo][0], ty[yo][xo][1], ty[yo][xo][2], ty[yo][xo][3]);
    };
    const int zc_end = min(D >> 1, (zchunk + 1) * ZC);
    for (int zc = zchunk * ZC; zc < zc_end; ++zc) {
        float out[2][2][2][4];
#pragma unroll
        for (int zo = 0; zo < 2; ++zo)
#pragma unroll
            for (int yo = 0; yo < 2; ++yo)
#pragma unroll
                for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                    for (int e = 0; e < 4; ++e) out[zo][yo][xo][e] = 0.f;
#pragma unroll
        for (int s = 0; s < NSRC; ++s) {
            const UpSrc &u = a.src[s];
            const AxisPair az = axis_pair(2 * zc, u.sz, u.d, a.align);
            float4 *lvl = ring + (size_t)s * 3 * 4 * 128 + threadIdx.x;
#pragma unroll
            for (int iz = 0; iz < 3; ++iz) {
                const int want = az.p[iz], slot = want % 3;
                // (static indexing of have[][]: the slot is found by comparison, not by a dynamic register index)
                const bool hit = (slot == 0 && have[s][0] == want) || (slot == 1 && have[s][1] == want) || (slot == 2 && have[s][2] == want);
                if (!hit) {
                    plane(u, want, lvl + slot * 4 * 128);
                    if (slot == 0) have[s][0] = want; else if (slot == 1) have[s][1] = want; else have[s][2] = want;
                }
            }
            // z pass, accumulated across the sources in level order: the association of the cell kernel (bit-identical results)
#pragma unroll
            for (int iz = 0; iz < 3; ++iz) {
                const float4 *src = lvl + (az.p[iz] % 3) * 4 * 128;
#pragma unroll
                for (int yo = 0; yo < 2; ++yo)
#pragma unroll
                    for (int xo = 0; xo < 2; ++xo) {
                        const float4 t = src[(yo * 2 + xo) * 128];
                        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            out[0][yo][xo][e] = fmaf(az.wa[iz], tv[e], out[0][yo][xo][e]);
                            out[1][yo][xo][e] = fmaf(az.wb[iz], tv[e], out[1][yo][xo][e]);
                        }
                    }
            }
        }
#pragma unroll
        for (int zo = 0; zo < 2; ++zo)
#pragma unroll
            for (int yo = 0; yo < 2; ++yo)
#pragma unroll
                for (int xo = 0; xo < 2; ++xo) {
                    const int64_t vox = ((b * D + 2 * zc + zo) * H + 2 * yc + yo) * (int64_t)W + 2 * xc + xo;
                    if (base != nullptr) {
                        float bv[4];
                        load_n<TB, 4>(base + vox * bs + c0, bv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) out[zo][yo][xo][e] = bv[e] + out[zo][yo][xo][e];
                    }
                    store_n<TO, 4>(y + vox * ys + c0, out[zo][yo][xo]);
                }
    }
}

// the cell kernel applies when every axis of every source is upscaled by at least 2 and the output extents are even
static bool cell_ok(const UpArgs &a, int D, int H, int W, int C) {
    if ((D | H | W) & 1 || C % 4 != 0 || (D >> 1) > 65535 || (H >> 1) > 65535) return false;
    for (int s = 0; s < a.nsrc; ++s)
        if (a.src[s].sz > 0.5f || a.src[s].sy > 0.5f || a.src[s].sx > 0.5f) return false;
    return true;
}

template <typename TS, typename TB, typename TO>
static int upsample_launch(const UpArgs &a, const TB *base, TO *y, int B, int D, int H, int W, int C, int64_t bs,
                           int64_t ys, cudaStream_t st) {
    constexpr int V = Pack<TS>::VEC;
    bool vec = (C % V == 0) && aligned16(y) && (base == nullptr || aligned16(base)) && (bs * sizeof(TB)) % 16 == 0 &&
               (ys * sizeof(TO)) % 16 == 0;
    for (int s = 0; s < a.nsrc; ++s) vec = vec && aligned16(a.src[s].ptr);
    if (!ab_old() && vec && cell_ok(a, D, H, W, C) && (D >> 1) >= 16) {      // deep volume: walk along z with plane rings in shared memory
        constexpr int ZC = 8;
        const int nchunk = ((D >> 1) + ZC - 1) / ZC;
        if ((int64_t)B * nchunk <= 65535) {
            dim3 grid((unsigned)(((W >> 1) * (C / 4) + 127) / 128), (unsigned)(H >> 1), (unsigned)(B * nchunk));
            const size_t smem = (size_t)a.nsrc * 3 * 4 * 128 * sizeof(float4);
            if (a.nsrc == 1) {
                upsample_cell_ring_kernel<TS, TB, TO, ZC, 1><<<grid, 128, smem, st>>>(a, base, y, D, H, W, C, C / 4, bs, ys);
            } else if (a.nsrc == 2) {
                upsample_cell_ring_kernel<TS, TB, TO, ZC, 2><<<grid, 128, smem, st>>>(a, base, y, D, H, W, C, C / 4, bs, ys);
            } else {
                static unsigned long long attr_done = 0;
                if (first_use_on_current_device(attr_done))
                    WF_CUDA_CHECK(cudaFuncSetAttribute(upsample_cell_ring_kernel<TS, TB, TO, ZC, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                upsample_cell_ring_kernel<TS, TB, TO, ZC, 3><<<grid, 128, smem, st>>>(a, base, y, D, H, W, C, C / 4, bs, ys);
            }
            WF_LAUNCH_CHECK();
            return WF_OK;
        }
    }
    if (vec && cell_ok(a, D, H, W, C) && (int64_t)B * (D >> 1) <= 65535) {
        dim3 grid((unsigned)(((W >> 1) * (C / 4) + 127) / 128), (unsigned)(H >> 1), (unsigned)(B * (D >> 1)));
        upsample_cell_kernel<TS, TB, TO><<<grid, 128, 0, st>>>(a, base, y, D, H, W, C, C / 4, bs, ys);
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    if (vec) {
        const int64_t total = (int64_t)B * D * H * W * (C / V);
        dim3 grid((unsigned)((W * (C / V) + 255) / 256), (unsigned)H, (unsigned)(B * D));
        upsample_add_kernel<TS, TB, TO, V><<<grid, 256, 0, st>>>(a, base, y, total, D, H, W, C, C / V, bs, ys);
    } else {
        const int64_t total = (int64_t)B * D * H * W * C;
        dim3 grid((unsigned)((W * C + 255) / 256), (unsigned)H, (unsigned)(B * D));
        upsample_add_kernel<TS, TB, TO, 1><<<grid, 256, 0, st>>>(a, base, y, total, D, H, W, C, C, bs, ys);
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace wf

extern "C" int wf_upsample_trilinear_add_ndhwc(const void *const *srcs, const int *src_dims, int nsrc, const void *base,
                                               void *y, int src_dtype, int io_dtype, int align_corners, int B, int D,
                                               int H, int W, int C, int64_t base_vox_stride, int64_t y_vox_stride,
                                               void *stream) {
    if (!srcs || !src_dims || !y) return WF_ERR_NULL_POINTER;
    if (nsrc < 1 || nsrc > 3 || B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0) return WF_ERR_BAD_SHAPE;
    if (H > 65535 || (int64_t)B * D > 65535) return WF_ERR_UNSUPPORTED;
    wf::UpArgs a;
    a.nsrc = nsrc;
    a.align = align_corners ? 1 : 0;
    const int out[3] = {D, H, W};
    for (int s = 0; s < nsrc; ++s) {
        if (!srcs[s]) return WF_ERR_NULL_POINTER;
        a.src[s].ptr = srcs[s];
        const int *dm = src_dims + 3 * s;
        if (dm[0] <= 0 || dm[1] <= 0 || dm[2] <= 0) return WF_ERR_BAD_SHAPE;
        a.src[s].d = dm[0]; a.src[s].h = dm[1]; a.src[s].w = dm[2];
        float sc[3];
        for (int k = 0; k < 3; ++k)
            sc[k] = align_corners ? (out[k] > 1 ? (float)(dm[k] - 1) / (float)(out[k] - 1) : 0.f) : (float)dm[k] / (float)out[k];
        a.src[s].sz = sc[0]; a.src[s].sy = sc[1]; a.src[s].sx = sc[2];
    }
    cudaStream_t st = (cudaStream_t)stream;
    using bf = __nv_bfloat16;
    // sources and base/output may differ in storage type (bf16 attention output, fp32 residual stream)
    if (src_dtype == WF_F32 && io_dtype == WF_F32)
        return wf::upsample_launch<float, float, float>(a, (const float *)base, (float *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    if (src_dtype == WF_BF16 && io_dtype == WF_BF16)
        return wf::upsample_launch<bf, bf, bf>(a, (const bf *)base, (bf *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    if (src_dtype == WF_BF16 && io_dtype == WF_F32)
        return wf::upsample_launch<bf, float, float>(a, (const float *)base, (float *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    if (src_dtype == WF_F32 && io_dtype == WF_BF16)
        return wf::upsample_launch<float, bf, bf>(a, (const bf *)base, (bf *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    using hf = __half;
    if (src_dtype == WF_F16 && io_dtype == WF_F16)
        return wf::upsample_launch<hf, hf, hf>(a, (const hf *)base, (hf *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    if (src_dtype == WF_F16 && io_dtype == WF_F32)
        return wf::upsample_launch<hf, float, float>(a, (const float *)base, (float *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    if (src_dtype == WF_F32 && io_dtype == WF_F16)
        return wf::upsample_launch<float, hf, hf>(a, (const hf *)base, (hf *)y, B, D, H, W, C, base_vox_stride, y_vox_stride, st);
    return WF_ERR_BAD_DTYPE;
}
