/* Oracle (TEST INFRASTRUCTURE ONLY - never linked into the product): plain-C 3D Haar analysis / synthesis.
 *
 * Restates what the reference obtains from ptwt==0.1.9 at network_models/wave_helper.py:350 (wavedec3, 'db1',
 * mode='zero', level 1) and network_models/idwt_upsample.py:160 (waverec3) for even extents, in the closed form
 *     c[pqr][z,y,x] = (1/(2*sqrt(2))) * sum_{i,j,k in {0,1}} (-1)^(p*i+q*j+r*k) * x[2z+i, 2y+j, 2x+k]
 * with sub-bands ordered aaa,aad,ada,add,daa,dad,dda,ddd (letter order = D,H,W; a = low, d = high).
 * ptwt is third-party and absent from /root/reference: PARITY UNPINNED at this boundary (see oracle/__init__.py);
 * the closed form is checked against the conv3d statement in oracle/haar.py by tests/test_oracle_haar.py.
 *
 * Layout: x is [n, D, H, W] contiguous (n = all leading dims folded, as ptwt does); coefficients are
 * [n, 8, D/2, H/2, W/2] contiguous.  Built by oracle/Makefile into oracle/_build/libhaar3d_oracle.so.
 */
#include <stddef.h>
#include <math.h>

#define HAAR_DEFINE(NAME, T)                                                                                   \
    void haar3d_dwt_##NAME(const T *x, T *c, long n, long D, long H, long W)                                   \
    {                                                                                                          \
        const long d = D / 2, h = H / 2, w = W / 2;                                                            \
        const T s = (T)(1.0 / (2.0 * sqrt(2.0)));                                                              \
        for (long b = 0; b < n; ++b)                                                                           \
            for (long z = 0; z < d; ++z)                                                                       \
                for (long y = 0; y < h; ++y) {                                                                 \
                    const T *r00 = x + ((b * D + 2 * z) * H + 2 * y) * W;                                      \
                    const T *r01 = r00 + W, *r10 = r00 + H * W, *r11 = r10 + W;                                \
                    T *o = c + ((b * 8 * d + z) * h + y) * w;                                                  \
                    const long band = d * h * w;                                                               \
                    for (long xx = 0; xx < w; ++xx) {                                                          \
                        const T v000 = r00[2 * xx], v001 = r00[2 * xx + 1];                                    \
                        const T v010 = r01[2 * xx], v011 = r01[2 * xx + 1];                                    \
                        const T v100 = r10[2 * xx], v101 = r10[2 * xx + 1];                                    \
                        const T v110 = r11[2 * xx], v111 = r11[2 * xx + 1];                                    \
                        /* W axis */                                                                           \
                        const T a00 = v000 + v001, d00 = v000 - v001;                                          \
                        const T a01 = v010 + v011, d01 = v010 - v011;                                          \
                        const T a10 = v100 + v101, d10 = v100 - v101;                                          \
                        const T a11 = v110 + v111, d11 = v110 - v111;                                          \
                        /* H axis: index = (z-bit)(h-kind)(w-kind) */                                          \
                        const T aa0 = a00 + a01, da0 = a00 - a01, ad0 = d00 + d01, dd0 = d00 - d01;            \
                        const T aa1 = a10 + a11, da1 = a10 - a11, ad1 = d10 + d11, dd1 = d10 - d11;            \
                        /* D axis; sub-band name = (D-kind)(H-kind)(W-kind) */                                 \
                        o[0 * band + xx] = s * (aa0 + aa1); /* aaa */                                          \
                        o[1 * band + xx] = s * (ad0 + ad1); /* aad */                                          \
                        o[2 * band + xx] = s * (da0 + da1); /* ada */                                          \
                        o[3 * band + xx] = s * (dd0 + dd1); /* add */                                          \
                        o[4 * band + xx] = s * (aa0 - aa1); /* daa */                                          \
                        o[5 * band + xx] = s * (ad0 - ad1); /* dad */                                          \
                        o[6 * band + xx] = s * (da0 - da1); /* dda */                                          \
                        o[7 * band + xx] = s * (dd0 - dd1); /* ddd */                                          \
                    }                                                                                          \
                }                                                                                              \
    }                                                                                                          \
                                                                                                               \
    void haar3d_idwt_##NAME(const T *c, T *x, long n, long d, long h, long w)                                  \
    {                                                                                                          \
        const long D = 2 * d, H = 2 * h, W = 2 * w;                                                            \
        const T s = (T)(1.0 / (2.0 * sqrt(2.0)));                                                              \
        for (long b = 0; b < n; ++b)                                                                           \
            for (long z = 0; z < d; ++z)                                                                       \
                for (long y = 0; y < h; ++y) {                                                                 \
                    const T *o = c + ((b * 8 * d + z) * h + y) * w;                                            \
                    const long band = d * h * w;                                                               \
                    T *r00 = x + ((b * D + 2 * z) * H + 2 * y) * W;                                            \
                    T *r01 = r00 + W, *r10 = r00 + H * W, *r11 = r10 + W;                                      \
                    for (long xx = 0; xx < w; ++xx) {                                                          \
                        const T aaa = o[0 * band + xx], aad = o[1 * band + xx];                                \
                        const T ada = o[2 * band + xx], add = o[3 * band + xx];                                \
                        const T daa = o[4 * band + xx], dad = o[5 * band + xx];                                \
                        const T dda = o[6 * band + xx], ddd = o[7 * band + xx];                                \
                        /* undo D axis: plane 0 = sum, plane 1 = difference */                                 \
                        const T aa0 = aaa + daa, aa1 = aaa - daa, ad0 = aad + dad, ad1 = aad - dad;            \
                        const T da0 = ada + dda, da1 = ada - dda, dd0 = add + ddd, dd1 = add - ddd;            \
                        /* undo H axis */                                                                      \
                        const T a00 = aa0 + da0, a01 = aa0 - da0, d00 = ad0 + dd0, d01 = ad0 - dd0;            \
                        const T a10 = aa1 + da1, a11 = aa1 - da1, d10 = ad1 + dd1, d11 = ad1 - dd1;            \
                        /* undo W axis */                                                                      \
                        r00[2 * xx] = s * (a00 + d00); r00[2 * xx + 1] = s * (a00 - d00);                      \
                        r01[2 * xx] = s * (a01 + d01); r01[2 * xx + 1] = s * (a01 - d01);                      \
                        r10[2 * xx] = s * (a10 + d10); r10[2 * xx + 1] = s * (a10 - d10);                      \
                        r11[2 * xx] = s * (a11 + d11); r11[2 * xx + 1] = s * (a11 - d11);                      \
                    }                                                                                          \
                }                                                                                              \
    }

HAAR_DEFINE(f32, float)
HAAR_DEFINE(f64, double)

/* Single-threaded on purpose (no OpenMP runtime in the image); oracle/haar_c.py splits the leading dimension over a
 * thread pool - ctypes releases the GIL - when a multi-core CPU baseline is wanted. */
