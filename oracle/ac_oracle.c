/*
 * ac_oracle.c -- CPU restatement of the Anime4KCPP v3 CNN upscaling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the B200 backend:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product library never links or calls it.
 *
 * Every function follows the reference's *Generic* CPU backend (the backend the
 * reference's own ProcessorTest.cpp:129 uses as ground truth, `create("cpu", 1, ...)`),
 * including its fp32 summation order.  Build WITHOUT fast-math and WITHOUT FMA contraction
 * (see oracle/Makefile) so the arithmetic is the reference's.  It is pinned bit-for-bit
 * against the compiled reference (oracle/_ref, built by oracle/build_ref.sh from the
 * sources under /root/reference) by tests/test_oracle_vs_ref.py and by the committed
 * golden vectors under tests/golden/ (minted from that compiled reference).
 *
 * Parity status of the individual stages:
 *   - rgb2yuv / rgba2yuva / yuv2rgb / yuva2rgba, the luma networks (ACNetLegacy, ACNet<8>,
 *     ARNet<8>), multi-pass 4x: pinned against the compiled reference.
 *   - ARNet: reference *code* pinned; the weights are synthetic (ARNet.p is a missing blob).
 *   - Catmull-Rom chroma resize: "parity unpinned".  The reference delegates it to
 *     nothings/stb stb_image_resize2.h (cmake/dependency/stb.cmake:15-19, GIT_TAG master),
 *     which is not vendored and not in this container.  orc_resize_catmull_rom restates the
 *     published stb gather-upsample algorithm behind the reference's call site
 *     core/src/ImageResize.cpp:136-272 (kernel :62-83, support 2.0 :202-205, edge clamp :269).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_U8 0x001
#define ORC_U16 0x002
#define ORC_F32 0x204

#define ORC_FAMILY_LEGACY 0
#define ORC_FAMILY_ACNET 1
#define ORC_FAMILY_ARNET 2

#define ORC_ACT_IDENTITY 0
#define ORC_ACT_RELU 1
#define ORC_ACT_PRELU 2

/*
 * Summation order.  The reference's CPU processor picks one of several per-ISA backends at run time
 * (core/src/processor/cpu/CPUProcessor.cpp:761-813) and they do not round identically:
 *   ORC_ORDER_GENERIC (0): OpImplGeneric, CPU/Generic.hpp:55-81 -- per input channel a left fold over the taps,
 *       products and sums rounded separately; `create("cpu", 1, ...)`, the ground truth of the reference's own tests.
 *   ORC_ORDER_FMA (1): OpImplX86SIMD256<true>, CPU/X86/AVX.hpp:24-146 -- what the auto-ISA choice executes for the
 *       8-channel layers on every FMA-capable x86 (the AVX512 backend defers to it for cin < 16,
 *       X86/AVX512.hpp:116): per input channel an FMA chain over the taps, the 8 chains reduced by the hsum tree
 *       ((s0+s4)+(s2+s6))+((s1+s5)+(s3+s7)), added onto the bias; scalar epilogue expressions (`v*scale + id`,
 *       `sat*max + 0.5f`) are contracted into FMAs by g++ in those translation units (-mfma, -ffp-contract=fast).
 */
#define ORC_ORDER_GENERIC 0
#define ORC_ORDER_FMA 1
static int g_order = ORC_ORDER_GENERIC;
void orc_set_order(int order) { g_order = order == ORC_ORDER_FMA ? ORC_ORDER_FMA : ORC_ORDER_GENERIC; }
int orc_get_order(void) { return g_order; }

static inline float hsum8(const float *s) { return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7])); }
/* `a*b + c` as the backend's translation unit rounds it */
static inline float muladd(float a, float b, float c) { return g_order == ORC_ORDER_FMA ? fmaf(a, b, c) : a * b + c; }

/* ---- core/internal/AC/Core/Internal/Util.hpp:50-78 ------------------------------------ */
static inline float to_float_u8(uint8_t v) { return (float)v / 255.0f; }
static inline float to_float_u16(uint16_t v) { return (float)v / 65535.0f; }
static inline float saturate(float v) { return v < 0.0f ? 0.0f : (v < 1.0f ? v : 1.0f); }
/* net_* variants: fromFloat as instantiated inside a backend translation unit (layer tails) */
static inline uint8_t from_float_u8(float v) { return (uint8_t)(saturate(v) * 255.0f + 0.5f); }
static inline uint16_t from_float_u16(float v) { return (uint16_t)(saturate(v) * 65535.0f + 0.5f); }
static inline uint8_t net_from_float_u8(float v) { return (uint8_t)muladd(saturate(v), 255.0f, 0.5f); }
static inline uint16_t net_from_float_u16(float v) { return (uint16_t)muladd(saturate(v), 65535.0f, 0.5f); }

static inline float load_elem(const uint8_t *row, int x, int type)
{
    switch (type)
    {
    case ORC_U8: return to_float_u8(row[x]);
    case ORC_U16: return to_float_u16(((const uint16_t *)row)[x]);
    default: return ((const float *)row)[x];
    }
}
static inline void store_elem(uint8_t *row, int x, int type, float v)
{
    switch (type)
    {
    case ORC_U8: row[x] = from_float_u8(v); break;
    case ORC_U16: ((uint16_t *)row)[x] = from_float_u16(v); break;
    default: ((float *)row)[x] = saturate(v); break;
    }
}

static inline void net_store_elem(uint8_t *row, int x, int type, float v)
{
    switch (type)
    {
    case ORC_U8: row[x] = net_from_float_u8(v); break;
    case ORC_U16: ((uint16_t *)row)[x] = net_from_float_u16(v); break;
    default: ((float *)row)[x] = saturate(v); break;
    }
}

/* ---- activations, core/internal/AC/Core/Internal/Processor/CPU/Common.hpp:21-50 -------- */
static inline float activate(float v, int act, const float *alphas, int c)
{
    if (act == ORC_ACT_RELU) return v > 0.0f ? v : 0.0f;
    if (act == ORC_ACT_PRELU) return (v > 0.0f ? v : 0.0f) + alphas[c] * (v < 0.0f ? v : 0.0f);
    return v;
}

/* ---- OpImplGeneric::conv<8,cout,9>, CPU/Generic.hpp:71-80: per input channel a left fold
 *      over the 9 taps, channel partial sums accumulated onto 0, bias added last ---------- */
static inline void conv9_c8(const float *const rptr[9], float *out, int cout, const float *kernels, const float *biases)
{
    for (int n = 0; n < cout; n++)
    {
        const float *k = kernels + n * 8 * 9;
        if (g_order == ORC_ORDER_FMA)
        {
            float s[8];
            for (int c = 0; c < 8; c++)
            {
                s[c] = 0.0f;
                for (int p = 0; p < 9; p++) s[c] = fmaf(rptr[p][c], k[p * 8 + c], s[c]);
            }
            out[n] = biases[n] + hsum8(s);
            continue;
        }
        float sum = 0.0f;
        for (int c = 0; c < 8; c++)
        {
            float s = rptr[0][c] * k[0 * 8 + c];
            for (int p = 1; p < 9; p++) s = s + rptr[p][c] * k[p * 8 + c];
            sum += s;
        }
        out[n] = sum + biases[n];
    }
}
/* OpImplGeneric::conv<8,8,1> (the ARNet 1x1), same shape with one tap */
static inline void conv1_c8(const float *in, float *out, const float *kernels, const float *biases)
{
    for (int n = 0; n < 8; n++)
    {
        if (g_order == ORC_ORDER_FMA)
        {
            float s[8];
            for (int c = 0; c < 8; c++) s[c] = in[c] * kernels[n * 8 + c];
            out[n] = biases[n] + hsum8(s);
            continue;
        }
        float sum = 0.0f;
        for (int c = 0; c < 8; c++) sum += in[c] * kernels[n * 8 + c];
        out[n] = sum + biases[n];
    }
}
/* OpImplGeneric::dot<8>, CPU/Generic.hpp:55-59 */
static inline float dot8(const float *a, const float *b)
{
    if (g_order == ORC_ORDER_FMA)
    {
        float t[8];
        for (int i = 0; i < 8; i++) t[i] = a[i] * b[i];
        return hsum8(t);
    }
    float s = a[0] * b[0];
    for (int i = 1; i < 8; i++) s = s + a[i] * b[i];
    return s;
}

static inline void gather9(const float *src, int w, int h, int i, int j, const float *rptr[9])
{
    int tp = i > 0 ? 1 : 0, bp = i < h - 1 ? 1 : 0, lp = j > 0 ? 1 : 0, rp = j < w - 1 ? 1 : 0;
    const float *r0 = src + (size_t)(i - tp) * w * 8, *r1 = src + (size_t)i * w * 8, *r2 = src + (size_t)(i + bp) * w * 8;
    rptr[0] = r0 + (j - lp) * 8; rptr[1] = r0 + j * 8; rptr[2] = r0 + (j + rp) * 8;
    rptr[3] = r1 + (j - lp) * 8; rptr[4] = r1 + j * 8; rptr[5] = r1 + (j + rp) * 8;
    rptr[6] = r2 + (j - lp) * 8; rptr[7] = r2 + j * 8; rptr[8] = r2 + (j + rp) * 8;
}

/* ---- conv3x3_cin1, CPU/Common.hpp:166-197 + OpImplGeneric::conv_cin1 Generic.hpp:61-69 -- */
static void layer_cin1(const uint8_t *src, int w, int h, int stride, int type, float *dst,
                       const float *kernels, const float *biases, int act, const float *alphas)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
    {
        int tp = i > 0 ? 1 : 0, bp = i < h - 1 ? 1 : 0;
        const uint8_t *r0 = src + (size_t)(i - tp) * stride, *r1 = src + (size_t)i * stride, *r2 = src + (size_t)(i + bp) * stride;
        for (int j = 0; j < w; j++)
        {
            int lp = j > 0 ? 1 : 0, rp = j < w - 1 ? 1 : 0;
            float r[9] = {
                load_elem(r0, j - lp, type), load_elem(r0, j, type), load_elem(r0, j + rp, type),
                load_elem(r1, j - lp, type), load_elem(r1, j, type), load_elem(r1, j + rp, type),
                load_elem(r2, j - lp, type), load_elem(r2, j, type), load_elem(r2, j + rp, type) };
            float *out = dst + ((size_t)i * w + j) * 8;
            for (int n = 0; n < 8; n++)
            {
                const float *k = kernels + n * 9;
                float s;
                if (g_order == ORC_ORDER_FMA)
                {
                    /* conv_cin1<8,9>, X86/AVX.hpp:95-124: 8 products through the hsum tree, 9th tap added as a scalar */
                    float t[8];
                    for (int p = 0; p < 8; p++) t[p] = r[p] * k[p];
                    s = muladd(r[8], k[8], hsum8(t));
                }
                else
                {
                    s = r[0] * k[0];
                    for (int p = 1; p < 9; p++) s = s + r[p] * k[p];
                }
                out[n] = activate(s + biases[n], act, alphas, n);
            }
        }
    }
}

/* ---- conv3x3_float<8,8>, CPU/Common.hpp:116-164: act, then optional `sum*scale + id` ---- */
static void layer_8to8(const float *src, int w, int h, float *dst, const float *kernels, const float *biases,
                       int act, const float *alphas, const float *residual, float scale)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float sum[8];
            gather9(src, w, h, i, j, rptr);
            conv9_c8(rptr, sum, 8, kernels, biases);
            float *out = dst + ((size_t)i * w + j) * 8;
            const float *id = residual ? residual + ((size_t)i * w + j) * 8 : 0;
            for (int n = 0; n < 8; n++)
            {
                float v = activate(sum[n], act, alphas, n);
                if (id) v = muladd(v, scale, id[n]);
                out[n] = v;
            }
        }
}

/* ---- conv3x3_conv1x1_float<8,8,8,false,false>, CPU/Common.hpp:223-288 (ARNet body end) -- */
static void layer_8to8_res_1x1(const float *src, int w, int h, float *dst,
                               const float *k3, const float *b3, const float *id3, float scale3,
                               const float *k1, const float *b1, const float *alphas1, const float *feat)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float buf[8], sum[8];
            gather9(src, w, h, i, j, rptr);
            conv9_c8(rptr, buf, 8, k3, b3);
            size_t o = ((size_t)i * w + j) * 8;
            for (int n = 0; n < 8; n++) buf[n] = muladd(buf[n], scale3, id3[o + n]);
            conv1_c8(buf, sum, k1, b1);
            for (int n = 0; n < 8; n++)
            {
                float v = activate(sum[n], ORC_ACT_PRELU, alphas1, n);
                dst[o + n] = v * 1.0f + feat[o + n];
            }
        }
}

/* ---- conv3x3_deconv2x2_float<8,8,1,ReLU>, CPU/Common.hpp:344-393 ------------------------ */
static void tail_deconv(const float *src, int w, int h, uint8_t *dst, int dst_stride, int type,
                        const float *k1, const float *b1, const float *k2)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float sum[8];
            gather9(src, w, h, i, j, rptr);
            conv9_c8(rptr, sum, 8, k1, b1);
            for (int n = 0; n < 8; n++) sum[n] = sum[n] > 0.0f ? sum[n] : 0.0f;
            for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++)
                    net_store_elem(dst + (size_t)(2 * i + dy) * dst_stride, 2 * j + dx, type, dot8(sum, k2 + 8 * (dy * 2 + dx)));
        }
}

/* ---- conv3x3_pixelshuffle_float<8,2,Identity,ResidualArg>, CPU/Common.hpp:290-342 ------- */
static void tail_pixelshuffle(const float *src, int w, int h, uint8_t *dst, int dst_stride, int type,
                              const float *kernels, const float *biases, const uint8_t *luma, int luma_stride)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float sum[4];
            gather9(src, w, h, i, j, rptr);
            conv9_c8(rptr, sum, 4, kernels, biases);
            float id = load_elem(luma + (size_t)i * luma_stride, j, type);
            for (int n = 0; n < 4; n++)
                net_store_elem(dst + (size_t)(2 * i + (n >> 1)) * dst_stride, 2 * j + (n & 1), type, sum[n] * 1.0f + id);
        }
}

/* ======================================================================================
 * ArtCNN<16/32> and FSRCNNX<8/16>: the same layer templates with F features.
 * OpImpl::conv<cin,cout,cpos> for any cin (multiple of 8): Generic.hpp:71-80 (per input channel a left fold over the
 * taps, channel sums accumulated onto 0, bias last) / X86/AVX.hpp:32-58,126-146 (256-bit FMA backend: lane c%8 runs one
 * FMA chain over (tap, 8-channel chunk) in that order, then the hsum tree, added onto the bias).
 * ====================================================================================== */
#define ORC_FAMILY_ARTCNN 3
#define ORC_FAMILY_FSRCNNX 4

static void conv_any(const float *const *rptr, int cpos, int cin, float *out, int cout, const float *kernels, const float *biases)
{
    for (int n = 0; n < cout; n++)
    {
        const float *k = kernels + (size_t)n * cin * cpos;
        if (g_order == ORC_ORDER_FMA)
        {
            float s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
            for (int p = 0; p < cpos; p++)
                for (int idx = 0; idx < cin / 8; idx++)
                    for (int c = 0; c < 8; c++) s[c] = fmaf(rptr[p][idx * 8 + c], k[p * cin + idx * 8 + c], s[c]);
            out[n] = biases[n] + hsum8(s);
            continue;
        }
        float sum = 0.0f;
        for (int c = 0; c < cin; c++)
        {
            float t = rptr[0][c] * k[c];
            for (int p = 1; p < cpos; p++) t = t + rptr[p][c] * k[p * cin + c];
            sum += t;
        }
        out[n] = sum + biases[n];
    }
}

/* conv3x3_cin1 / conv5x5_cin1 (Common.hpp:166-197, 199-221): ks x ks window with clamp-to-edge coordinates;
 * OpImpl::conv_cin1<cout,cpos>: Generic.hpp:61-69 (left fold) / X86/AVX.hpp:95-124 (8-tap vectors as FMA chains per lane,
 * hsum, the remaining taps added as scalars, bias last). */
static void layer_head_any(const uint8_t *src, int w, int h, int stride, int type, float *dst, int F, int ks,
                           const float *kernels, const float *biases)
{
    const int cpos = ks * ks, half = ks / 2;
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            float r[25];
            for (int a = 0; a < ks; a++)
                for (int b = 0; b < ks; b++)
                {
                    int y = i + a - half, x = j + b - half;
                    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
                    x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
                    r[a * ks + b] = load_elem(src + (size_t)y * stride, x, type);
                }
            float *out = dst + ((size_t)i * w + j) * F;
            for (int n = 0; n < F; n++)
            {
                const float *k = kernels + n * cpos;
                float s;
                if (g_order == ORC_ORDER_FMA)
                {
                    float t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
                    const int count = cpos / 8;
                    for (int idx = 0; idx < count; idx++)
                        for (int c = 0; c < 8; c++) t[c] = fmaf(r[idx * 8 + c], k[idx * 8 + c], t[c]);
                    s = hsum8(t);
                    for (int q = count * 8; q < cpos; q++) s = fmaf(r[q], k[q], s);
                }
                else
                {
                    s = r[0] * k[0];
                    for (int q = 1; q < cpos; q++) s = s + r[q] * k[q];
                }
                out[n] = s + biases[n];     /* Identity activation in both families */
            }
        }
}

static inline void gather9_any(const float *src, int w, int h, int F, int i, int j, const float *rptr[9])
{
    int tp = i > 0 ? 1 : 0, bp = i < h - 1 ? 1 : 0, lp = j > 0 ? 1 : 0, rp = j < w - 1 ? 1 : 0;
    const float *r0 = src + (size_t)(i - tp) * w * F, *r1 = src + (size_t)i * w * F, *r2 = src + (size_t)(i + bp) * w * F;
    rptr[0] = r0 + (j - lp) * F; rptr[1] = r0 + j * F; rptr[2] = r0 + (j + rp) * F;
    rptr[3] = r1 + (j - lp) * F; rptr[4] = r1 + j * F; rptr[5] = r1 + (j + rp) * F;
    rptr[6] = r2 + (j - lp) * F; rptr[7] = r2 + j * F; rptr[8] = r2 + (j + rp) * F;
}

/* conv3x3_float<F,F> (Common.hpp:116-164): activation, then `sum*scale + id` (scale 1.0 here) */
static void layer_conv_any(const float *src, int w, int h, int F, float *dst, const float *kernels, const float *biases,
                           int act, const float *alphas, const float *residual)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float sum[32];
            gather9_any(src, w, h, F, i, j, rptr);
            conv_any(rptr, 9, F, sum, F, kernels, biases);
            size_t o = ((size_t)i * w + j) * F;
            for (int n = 0; n < F; n++)
            {
                float v = activate(sum[n], act, alphas, n);
                if (residual) v = muladd(v, 1.0f, residual[o + n]);
                dst[o + n] = v;
            }
        }
}

/* conv3x3_conv1x1_float<F,F,F,false,true> (Common.hpp:223-288, Backend.hpp:239-250,296-307): conv3x3 + PReLU, then the
 * 1x1 + bias, `+ feat` (scale 1.0) and the second PReLU last (postactive1x1) */
static void layer_conv_1x1_any(const float *src, int w, int h, int F, float *dst,
                               const float *k3, const float *b3, const float *a3,
                               const float *k1, const float *b1, const float *a1, const float *feat)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float buf[32], sum[32];
            gather9_any(src, w, h, F, i, j, rptr);
            conv_any(rptr, 9, F, buf, F, k3, b3);
            for (int n = 0; n < F; n++) buf[n] = activate(buf[n], ORC_ACT_PRELU, a3, n);
            const float *one[1] = { buf };
            conv_any(one, 1, F, sum, F, k1, b1);
            size_t o = ((size_t)i * w + j) * F;
            for (int n = 0; n < F; n++) dst[o + n] = activate(muladd(sum[n], 1.0f, feat[o + n]), ORC_ACT_PRELU, a1, n);
        }
}

/* conv3x3_pixelshuffle_float<OUT,F,2,Identity,nullptr> (Common.hpp:290-342): no nearest-neighbour residual */
static void tail_pixelshuffle_any(const float *src, int w, int h, int F, uint8_t *dst, int dst_stride, int type,
                                  const float *kernels, const float *biases)
{
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++)
        {
            const float *rptr[9];
            float sum[4];
            gather9_any(src, w, h, F, i, j, rptr);
            conv_any(rptr, 9, F, sum, 4, kernels, biases);
            for (int n = 0; n < 4; n++)
                net_store_elem(dst + (size_t)(2 * i + (n >> 1)) * dst_stride, 2 * j + (n & 1), type, sum[n]);
        }
}

/* One 2x pass of ArtCNN<F> (CPUProcessor.cpp:1519-1549 / :1570-1600) or FSRCNNX<F> (:1621-1655 / :1676-1710);
 * weight offsets: Model/ArtCNN.hpp:33-60, Model/FSRCNNX.hpp:33-85. */
static int luma_pass_wide(int family, int blocks, int F, const float *k, const float *b, const float *a,
                          const uint8_t *src, int w, int h, int src_stride, int type, uint8_t *dst, int dst_stride)
{
    if (!(F == 8 || F == 16 || F == 32)) return -1;
    size_t n = (size_t)w * h * F;
    float *t1 = (float *)malloc(n * sizeof(float)), *t2 = (float *)malloc(n * sizeof(float)), *feat = (float *)malloc(n * sizeof(float));
    if (!t1 || !t2 || !feat) { free(t1); free(t2); free(feat); return -1; }
    float *in = t2, *out = t1, *t;
    if (family == ORC_FAMILY_ARTCNN)
    {
        const int KH = F * 9, KL = F * F * 9;
        int l = 0;
        layer_head_any(src, w, h, src_stride, type, feat, F, 3, k, b); l++;
        layer_conv_any(feat, w, h, F, out, k + KH + KL * (l - 1), b + F * l, ORC_ACT_RELU, 0, 0); l++;
        t = in; in = out; out = t;
        for (int i = 0; i < blocks - 1; i++)
        {
            layer_conv_any(in, w, h, F, out, k + KH + KL * (l - 1), b + F * l, ORC_ACT_RELU, 0, 0); l++;
            t = in; in = out; out = t;
        }
        layer_conv_any(in, w, h, F, out, k + KH + KL * (l - 1), b + F * l, ORC_ACT_IDENTITY, 0, feat); l++;
        t = in; in = out; out = t;
        tail_pixelshuffle_any(in, w, h, F, dst, dst_stride, type, k + KH + KL * (l - 1), b + F * l);
    }
    else
    {
        const int KH = F * 25, KL = F * F * 9;
        int l = 0;
        layer_head_any(src, w, h, src_stride, type, feat, F, 5, k, b); l++;
        layer_conv_any(feat, w, h, F, out, k + KH + KL * (l - 1), b + F * l, ORC_ACT_PRELU, a + F * (l - 1), 0); l++;
        t = in; in = out; out = t;
        for (int i = 0; i < blocks - 2; i++)
        {
            layer_conv_any(in, w, h, F, out, k + KH + KL * (l - 1), b + F * l, ORC_ACT_PRELU, a + F * (l - 1), 0); l++;
            t = in; in = out; out = t;
        }
        layer_conv_1x1_any(in, w, h, F, out, k + KH + KL * (l - 1), b + F * l, a + F * (l - 1),
                           k + KH + KL * blocks, b + F * (l + 1), a + F * l, feat);
        l += 2;
        t = in; in = out; out = t;
        tail_pixelshuffle_any(in, w, h, F, dst, dst_stride, type, k + KH + KL * blocks + F * F, b + F * l);
    }
    free(t1); free(t2); free(feat);
    return 0;
}

/*
 * One 2x luma pass.  Layer sequencing follows core/src/processor/cpu/CPUProcessor.cpp:
 * ACNetLegacy :1371-1390, ACNet<8> :1417-1436, ARNet<8> :1464-1491; weight offsets follow
 * core/include/AC/Core/Model/ACNet.hpp:34-60,78-115 and ARNet.hpp:32-71.
 * Returns 0, or -1 on allocation failure / bad arguments.
 */
int orc_luma_pass(int family, int blocks, const float *k, const float *b, const float *a,
                  const void *src_, int w, int h, int src_stride, int type, void *dst_, int dst_stride)
{
    const uint8_t *src = (const uint8_t *)src_;
    uint8_t *dst = (uint8_t *)dst_;
    if (w <= 0 || h <= 0 || !src || !dst) return -1;
    if (family == ORC_FAMILY_ARTCNN || family == ORC_FAMILY_FSRCNNX)    /* blocks | (features << 16) */
        return luma_pass_wide(family, blocks & 0xffff, blocks >> 16, k, b, a, src, w, h, src_stride, type, dst, dst_stride);
    size_t n = (size_t)w * h * 8;
    float *t1 = (float *)malloc(n * sizeof(float)), *t2 = (float *)malloc(n * sizeof(float)), *feat = 0;
    if (!t1 || !t2) { free(t1); free(t2); return -1; }
    int rc = 0;
    if (family == ORC_FAMILY_LEGACY)
    {
        float *in = t2, *out = t1, *t;
        int l = 0;
        layer_cin1(src, w, h, src_stride, type, out, k, b, ORC_ACT_RELU, 0); l++;
        t = in; in = out; out = t;
        for (int i = 0; i < blocks - 1; i++)
        {
            layer_8to8(in, w, h, out, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_RELU, 0, 0, 0.0f); l++;
            t = in; in = out; out = t;
        }
        tail_deconv(in, w, h, dst, dst_stride, type, k + 72 + 576 * (l - 1), b + 8 * l, k + 72 + 576 * l);
    }
    else if (family == ORC_FAMILY_ACNET)
    {
        float *in = t2, *out = t1, *t;
        int l = 0;
        layer_cin1(src, w, h, src_stride, type, out, k, b, ORC_ACT_PRELU, a); l++;
        t = in; in = out; out = t;
        for (int i = 0; i < blocks; i++)
        {
            layer_8to8(in, w, h, out, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_PRELU, a + 8 * l, 0, 0.0f); l++;
            t = in; in = out; out = t;
        }
        tail_pixelshuffle(in, w, h, dst, dst_stride, type, k + 72 + 576 * (l - 1), b + 8 * l, src, src_stride);
    }
    else if (family == ORC_FAMILY_ARNET)
    {
        feat = (float *)malloc(n * sizeof(float));
        if (!feat) rc = -1;
        else
        {
            /* alpha(l) for odd l lives at 8*((l-1)/2), ARNet.hpp:45,65-71 */
            int l = 0;
            layer_cin1(src, w, h, src_stride, type, feat, k, b, ORC_ACT_IDENTITY, 0); l++;
            layer_8to8(feat, w, h, t1, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_PRELU, a + 8 * ((l - 1) / 2), 0, 0.0f); l++;
            layer_8to8(t1, w, h, t2, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_IDENTITY, 0, feat, 0.2f); l++;
            for (int i = 0; i < blocks - 2; i++)
            {
                layer_8to8(t2, w, h, t1, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_PRELU, a + 8 * ((l - 1) / 2), 0, 0.0f); l++;
                layer_8to8(t1, w, h, t2, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_IDENTITY, 0, t2, 0.2f); l++;
            }
            layer_8to8(t2, w, h, t1, k + 72 + 576 * (l - 1), b + 8 * l, ORC_ACT_PRELU, a + 8 * ((l - 1) / 2), 0, 0.0f); l++;
            /* l == 2*blocks: last 3x3, then the 1x1 at l+1 */
            layer_8to8_res_1x1(t1, w, h, t2, k + 72 + 576 * (l - 1), b + 8 * l, t2, 0.2f,
                               k + 72 + 576 * l, b + 8 * (l + 1), a + 8 * (l / 2), feat);
            l += 2;
            tail_pixelshuffle(t2, w, h, dst, dst_stride, type, k + 72 + 576 * (2 * blocks) + 64, b + 8 * l, src, src_stride);
        }
    }
    else rc = -1;
    free(t1); free(t2); free(feat);
    return rc;
}

/* ---- detail::rgb2yuv 2-plane, core/src/ImageProcess.cpp:38-61; rgba2yuva :113-138 -------- */
void orc_rgb2yuv(const void *src_, int w, int h, int c, int src_stride, int type,
                 void *y_, int y_stride, void *uv_, int uv_stride)
{
    const uint8_t *src = (const uint8_t *)src_;
    uint8_t *yp = (uint8_t *)y_, *uvp = (uint8_t *)uv_;
    int uvc = c - 1; /* 2 for RGB, 3 (u,v,a) for RGBA */
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
    {
        const uint8_t *in = src + (size_t)i * src_stride;
        uint8_t *yo = yp + (size_t)i * y_stride, *uvo = uvp + (size_t)i * uv_stride;
        for (int j = 0; j < w; j++)
        {
            float r = load_elem(in, j * c + 0, type), g = load_elem(in, j * c + 1, type), bl = load_elem(in, j * c + 2, type);
            float al = 1.0f;
            if (c == 4)
            {
                al = load_elem(in, j * c + 3, type);
                r = r * al; g = g * al; bl = bl * al;
            }
            float y = 0.299f * r + 0.587f * g + 0.114f * bl;
            float u = 0.564f * (bl - y) + 0.5f;
            float v = 0.713f * (r - y) + 0.5f;
            store_elem(yo, j, type, y);
            store_elem(uvo, j * uvc + 0, type, u);
            store_elem(uvo, j * uvc + 1, type, v);
            if (c == 4) store_elem(uvo, j * uvc + 2, type, al);
        }
    }
}

/* ---- detail::yuv2rgb 2-plane, core/src/ImageProcess.cpp:191-215; yuva2rgba :275-308 ------ */
void orc_yuv2rgb(const void *y_, int y_stride, const void *uv_, int uv_stride, int w, int h, int c, int type,
                 void *dst_, int dst_stride)
{
    const uint8_t *yp = (const uint8_t *)y_, *uvp = (const uint8_t *)uv_;
    uint8_t *dst = (uint8_t *)dst_;
    int uvc = c - 1;
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
    {
        const uint8_t *yi = yp + (size_t)i * y_stride, *uvi = uvp + (size_t)i * uv_stride;
        uint8_t *out = dst + (size_t)i * dst_stride;
        for (int j = 0; j < w; j++)
        {
            float y = load_elem(yi, j, type);
            float u = load_elem(uvi, j * uvc + 0, type) - 0.5f;
            float v = load_elem(uvi, j * uvc + 1, type) - 0.5f;
            float r = y + 1.403f * v;
            float g = y - 0.344f * u - 0.714f * v;
            float bl = y + 1.773f * u;
            if (c == 4)
            {
                float al = load_elem(uvi, j * uvc + 2, type);
                if (al > 1e-6f) { r /= al; g /= al; bl /= al; }
                else r = g = bl = 0.0f;
                store_elem(out, j * c + 3, type, al);
            }
            store_elem(out, j * c + 0, type, r);
            store_elem(out, j * c + 1, type, g);
            store_elem(out, j * c + 2, type, bl);
        }
    }
}

/* ---- Catmull-Rom, core/src/ImageResize.cpp:58-83 (bicubic<0,1,1,2>, Horner poly3) -------- */
static inline float poly3(float x, float c0, float c1, float c2, float c3) { return c0 + x * (c1 + x * (c2 + x * c3)); }
static float catmull_rom(float v)
{
    const float b = 0.0f, c = 0.5f;
    const float p0 = (6.0f - 2.0f * b) / 6.0f, p1 = 0.0f, p2 = (-18.0f + 12.0f * b + 6.0f * c) / 6.0f, p3 = (12.0f - 9.0f * b - 6.0f * c) / 6.0f;
    const float q0 = (8.0f * b + 24.0f * c) / 6.0f, q1 = (-12.0f * b - 48.0f * c) / 6.0f, q2 = (6.0f * b + 30.0f * c) / 6.0f, q3 = (-b - 6.0f * c) / 6.0f;
    float x = fabsf(v);
    if (x < 1.0f) return poly3(x, p0, p1, p2, p3);
    if (x < 2.0f) return poly3(x, q0, q1, q2, q3);
    return 0.0f;
}

typedef struct { int n0, n1; float c[12]; } orc_contrib; /* taps n0..n1 inclusive (already clamped in range) */

/*
 * stb_image_resize2 gather-upsample coefficients for one axis (published algorithm,
 * stbir__calculate_coefficients_for_gather_upsample + stbir__cleanup_gathered_coefficients):
 * output centre (n+0.5)/scale, taps with |distance| < support, coefficients normalised to
 * sum 1, taps outside [0,in_size) folded onto the clamped edge pixel (STBIR_EDGE_CLAMP).
 */
static void make_contribs(orc_contrib *out, int in_size, int out_size, float scale)
{
    const float support = 2.0f;
    float inv_scale = 1.0f / scale;
    float out_radius = support * scale;
    /* Down-scaling (scale < 1; the post-network luma resize of non-power-of-two factors, Processor.cpp:237,249,265): the
     * filter is stretched by 1/scale in input space -- output pixel n gathers every input pixel within 2/scale of its centre
     * with weight k((centre distance) * scale), normalised to sum 1, out-of-range taps folded onto the edge pixel.  This is the
     * gather form of stb_image_resize2's down-sampler (unvendored: parity unpinned, like the up-sampler). */
    const int down = scale < 1.0f;
    for (int n = 0; n < out_size; n++)
    {
        float out_center = (float)n + 0.5f;
        float in_center = out_center * inv_scale;
        float lo = down ? in_center - support * inv_scale : (out_center - out_radius) * inv_scale;
        float hi = down ? in_center + support * inv_scale : (out_center + out_radius) * inv_scale;
        int first = (int)floorf(lo + 0.5f), last = (int)floorf(hi - 0.5f);
        if (last < first) last = first;
        if (last - first > 10) last = first + 10;
        float raw[12], total = 0.0f;
        for (int i = 0; i <= last - first; i++)
        {
            float in_pixel_center = (float)(first + i) + 0.5f;
            raw[i] = down ? catmull_rom((in_center - in_pixel_center) * scale) : catmull_rom(in_center - in_pixel_center);
            total += raw[i];
        }
        float fs = 1.0f / total;
        for (int i = 0; i <= last - first; i++) raw[i] *= fs;
        int n0 = first < 0 ? 0 : first, n1 = last > in_size - 1 ? in_size - 1 : last;
        if (n1 < n0) n1 = n0 = (first < 0 ? 0 : in_size - 1);
        orc_contrib *o = &out[n];
        o->n0 = n0; o->n1 = n1;
        for (int i = 0; i < 12; i++) o->c[i] = 0.0f;
        for (int i = 0; i <= last - first; i++)
        {
            int p = first + i;
            if (p >= n0 && p <= n1) o->c[p - n0] = raw[i];
        }
        for (int i = 0; i <= last - first; i++)
        {
            int p = first + i;
            if (p < n0) o->c[0] += raw[i];
            else if (p > n1) o->c[n1 - n0] += raw[i];
        }
    }
}

/*
 * Catmull-Rom upscale of an interleaved `c`-channel plane by (out_w/in_w, out_h/in_h):
 * decode (u8: * 1/255), horizontal pass, vertical pass, encode (v*max + 0.5, clamp, truncate;
 * float: stored as is).  Returns 0 / -1.
 */
int orc_resize_catmull_rom(const void *src_, int w, int h, int c, int src_stride, int type,
                           void *dst_, int ow, int oh, int dst_stride)
{
    const uint8_t *src = (const uint8_t *)src_;
    uint8_t *dst = (uint8_t *)dst_;
    if (ow * 2 < w || oh * 2 < h || ow <= 0 || oh <= 0) return -1; /* down to 1/2 (what Processor::process needs), any up-scale */
    orc_contrib *hc = (orc_contrib *)malloc(sizeof(orc_contrib) * ow), *vc = (orc_contrib *)malloc(sizeof(orc_contrib) * oh);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)h * ow * c);
    if (!hc || !vc || !tmp) { free(hc); free(vc); free(tmp); return -1; }
    make_contribs(hc, w, ow, (float)ow / (float)w);
    make_contribs(vc, h, oh, (float)oh / (float)h);
    const float inv_u8 = 1.0f / 255.0f, inv_u16 = 1.0f / 65535.0f;
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < h; i++)
    {
        const uint8_t *row = src + (size_t)i * src_stride;
        float *trow = tmp + (size_t)i * ow * c;
        for (int x = 0; x < ow; x++)
        {
            const orc_contrib *k = &hc[x];
            for (int ch = 0; ch < c; ch++)
            {
                float s = 0.0f;
                for (int p = k->n0; p <= k->n1; p++)
                {
                    float d = type == ORC_U8 ? (float)row[p * c + ch] * inv_u8
                            : type == ORC_U16 ? (float)((const uint16_t *)row)[p * c + ch] * inv_u16
                                              : ((const float *)row)[p * c + ch];
                    float t = k->c[p - k->n0] * d;
                    s = (p == k->n0) ? t : s + t;
                }
                trow[x * c + ch] = s;
            }
        }
    }
#pragma omp parallel for schedule(guided)
    for (int y = 0; y < oh; y++)
    {
        const orc_contrib *k = &vc[y];
        uint8_t *orow = dst + (size_t)y * dst_stride;
        for (int x = 0; x < ow * c; x++)
        {
            float s = 0.0f;
            for (int p = k->n0; p <= k->n1; p++)
            {
                float t = k->c[p - k->n0] * tmp[(size_t)p * ow * c + x];
                s = (p == k->n0) ? t : s + t;
            }
            if (type == ORC_U8)
            {
                float f = s * 255.0f + 0.5f;
                f = f < 0.0f ? 0.0f : (f > 255.0f ? 255.0f : f);
                orow[x] = (uint8_t)f;
            }
            else if (type == ORC_U16)
            {
                float f = s * 65535.0f + 0.5f;
                f = f < 0.0f ? 0.0f : (f > 65535.0f ? 65535.0f : f);
                ((uint16_t *)orow)[x] = (uint16_t)f;
            }
            else ((float *)orow)[x] = s;
        }
    }
    free(hc); free(vc); free(tmp);
    return 0;
}

static int elem_size(int type) { return type & 0xff; }
static int align4(int v) { return (v + 3) & ~3; }

/* ceilLog2, core/internal/AC/Core/Internal/Util.hpp:80-89 */
static int ceil_log2(double v)
{
    uint64_t d;
    memcpy(&d, &v, 8);
    return (int)((((d >> 52) & 0x7ff) - 1023) + ((d << 12) != 0));
}

/*
 * The whole-image driver, core/src/processor/Processor.cpp:199-276: colour split, `power` 2x passes (luma re-quantised to
 * the image type between passes), for factors that are not powers of two a Catmull-Rom down-scale of the luma by
 * fxy = factor / 2^power (:237, :249), ONE Catmull-Rom chroma resize by the full factor, merge.
 * dst must be int(w*factor) x int(h*factor) x c of the same element type.  Returns 0 / -1.
 */
int orc_process(int family, int blocks, const float *k, const float *b, const float *a,
                const void *src, int w, int h, int c, int src_stride, int type, double factor,
                void *dst, int dst_stride)
{
    int power = factor > 2.0 ? ceil_log2(factor) : 1;
    double fxy = factor / (double)(1 << power);
    if (!(c == 1 || c == 3 || c == 4) || !(factor >= 1.0)) return -1;
    const int dw = (int)(w * factor), dh = (int)(h * factor);      /* ImageResize.cpp:153-154 / the caller-sized dst */
    int es = elem_size(type);
    void *y = 0, *uv = 0;
    int y_stride = src_stride, uv_stride = 0, uvc = c - 1, rc = 0;
    const void *cur = src;
    int cw = w, chh = h, cstride = src_stride;
    if (c > 1)
    {
        y_stride = align4(w * es); uv_stride = align4(w * uvc * es);
        y = malloc((size_t)y_stride * h); uv = malloc((size_t)uv_stride * h);
        if (!y || !uv) { free(y); free(uv); return -1; }
        orc_rgb2yuv(src, w, h, c, src_stride, type, y, y_stride, uv, uv_stride);
        cur = y; cstride = y_stride;
    }
    void *owned = y;
    for (int i = 0; i < power && rc == 0; i++)
    {
        int nw = cw * 2, nh = chh * 2;
        int last = (i == power - 1) && c == 1 && fxy == 1.0;
        int nstride = last ? dst_stride : align4(nw * es);
        void *out = last ? dst : malloc((size_t)nstride * nh);
        if (!out) { rc = -1; break; }
        rc = orc_luma_pass(family, blocks, k, b, a, cur, cw, chh, cstride, type, out, nstride);
        if (owned) free(owned);
        owned = last ? 0 : out;
        cur = out; cw = nw; chh = nh; cstride = nstride;
    }
    if (rc == 0 && fxy != 1.0)
    {
        /* gray: resize(out, dst, 0, 0) straight into dst; colour: resize(out, out, fxy, fxy) into a new luma plane */
        int nstride = c == 1 ? dst_stride : align4(dw * es);
        void *out = c == 1 ? dst : malloc((size_t)nstride * dh);
        if (!out) rc = -1;
        else
        {
            rc = orc_resize_catmull_rom(cur, cw, chh, 1, cstride, type, out, dw, dh, nstride);
            if (owned) free(owned);
            owned = c == 1 ? 0 : out;
            cur = out; cw = dw; chh = dh; cstride = nstride;
        }
    }
    if (rc == 0 && c > 1)
    {
        int uvs = align4(cw * uvc * es);
        void *uv2 = malloc((size_t)uvs * chh);
        if (!uv2) rc = -1;
        else
        {
            rc = orc_resize_catmull_rom(uv, w, h, uvc, uv_stride, type, uv2, cw, chh, uvs);
            if (rc == 0) orc_yuv2rgb(cur, cstride, uv2, uvs, cw, chh, c, type, dst, dst_stride);
            free(uv2);
        }
    }
    if (owned) free(owned);
    free(uv);
    return rc;
}
