// Device-side pieces of the colour path that more than one translation unit needs (no kernels here): the Catmull-Rom contributor
// record, the RGB -> YUV arithmetic, and the per-lane form of the chroma resize + merge that the TMEM engine's tail runs.
//
// Reference semantics: core/src/ImageProcess.cpp:38-61 (rgb2yuv), :191-215 (yuv2rgb); core/src/ImageResize.cpp:136-272 (Catmull-Rom through
// stb_image_resize2); core/src/processor/Processor.cpp:207-213, 251-253 (where the driver runs them).
#pragma once

#include "acb200_common.cuh"

namespace acb
{
    // one output sample's taps along one axis (already edge-folded): src index n0 .. n0+cnt-1
    struct Contrib
    {
        int n0, cnt;
        float c[6];
    };

    struct YuvFromRgb { float y, u, v; };
    __device__ __forceinline__ YuvFromRgb rgb_to_yuv(float r, float g, float b)
    {
        YuvFromRgb o;
        o.y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
        o.u = __fadd_rn(__fmul_rn(0.564f, __fsub_rn(b, o.y)), 0.5f);
        o.v = __fadd_rn(__fmul_rn(0.713f, __fsub_rn(r, o.y)), 0.5f);
        return o;
    }

    // ---- colour handling fused into the TMEM engine's segment kernels (8-bit RGB, exactly 2x) -----------------------------------
    // The same arithmetic as rgb2yuv_u8x4_kernel and chroma_merge_u8_kernel<3>, value for value and rounding for rounding, in
    // a form a single lane can run for the source pixel it owns in the network's tail: the horizontal pass of the lane's two
    // output columns over a sliding window of five source rows in registers, the vertical pass of its 2 x 2 output pixels, the
    // re-quantisation (Processor.cpp:251-253 materialises the resized plane) and the YUV->RGB merge.  Nothing is shared
    // between lanes, so the network's epilogue warps stay free of barriers.
    constexpr float CM_MAGIC = 8388608.0f;      // 2^23: float(b) = (0x4B000000 | b) - 2^23; trunc(f) = (f +rz 2^23) - 2^23 for 0 <= f < 2^23

    struct HTaps2
    {
        int off[5];         // byte offsets of source columns n0 .. n0 + 4 (clamped to the row) inside a (u, v) row; n0 = first tap of column a
        int d;              // output column b = 2 gx + 1 starts at source column n0 + d, d in {0, 1}
        bool all_d1;        // every lane of the warp has d == 1 (true away from the left / right image edge)
        float ca[4], cb[4];
    };
    // pix_bytes: bytes per pixel of the chroma plane, 2 for (u, v) and 3 for (u, v, a)
    __device__ __forceinline__ HTaps2 load_htaps2(const Contrib* __restrict__ htab, int ox, int sw_img, int pix_bytes)
    {
        const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(htab + ox)), b0 = __ldg(reinterpret_cast<const uint4*>(htab + ox + 1));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(&htab[ox].c[2])), b1 = __ldg(reinterpret_cast<const float2*>(&htab[ox + 1].c[2]));
        HTaps2 k;
        const int n0 = static_cast<int>(a0.x);
        k.d = static_cast<int>(b0.x) - n0;
        k.all_d1 = __all_sync(0xffffffffu, k.d == 1);
#pragma unroll
        for (int j = 0; j < 5; j++) k.off[j] = pix_bytes * min(n0 + j, sw_img - 1);
        k.ca[0] = __uint_as_float(a0.z); k.ca[1] = __uint_as_float(a0.w); k.ca[2] = a1.x; k.ca[3] = a1.y;
        k.cb[0] = __uint_as_float(b0.z); k.cb[1] = __uint_as_float(b0.w); k.cb[2] = b1.x; k.cb[3] = b1.y;
        return k;
    }
    // left-to-right sum of four separately rounded products (catmull_sample's order)
    __device__ __forceinline__ float tap4(float c0, float c1, float c2, float c3, float t0, float t1, float t2, float t3)
    {
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0), __fmul_rn(c1, t1)), __fmul_rn(c2, t2)), __fmul_rn(c3, t3));
    }
    // stb decode of a byte b, float(b) * (1/255) rounded once, from the magic form m = 2^23 + b (exact): fma(m, r, -2^23 r) -- the product
    // m r is exact inside the FMA and 2^23 r is a float, so the single rounding is that of b r: equal to __fmul_rn(float(b), 1/255) for every b
    __device__ __forceinline__ float decode_u8_magic(uint32_t magic_bits)
    {
        constexpr float R = 1.0f / 255.0f;
        return __fmaf_rn(__uint_as_float(magic_bits), R, -CM_MAGIC * R);
    }
    // horizontal pass of one source row of the interleaved (u, v) plane for the lane's two output columns: (u_a, v_a, u_b, v_b).
    // Columns past the image carry zero coefficients and are read clamped (finite), as the tiled kernel zero-fills them.
    // MODE 0: (u, v) plane, one 16-bit load per tap.  MODE 1: the (u, v) of a (u, v, a) plane, two byte loads per tap.  MODE 2: the alpha of
    // a (u, v, a) plane in BOTH halves of the pair (the caller keeps one machinery for every pair of channels).
    template<int MODE>
    __device__ __forceinline__ float4 chroma_hrow2(const uint8_t* __restrict__ row, const HTaps2& k)
    {
        float tu[5], tv[5];
#pragma unroll
        for (int j = 0; j < 5; j++)
        {
            if constexpr (MODE == 0)
            {
                const uint32_t raw = __ldg(reinterpret_cast<const unsigned short*>(row + k.off[j]));
                tu[j] = decode_u8_magic(__byte_perm(raw, 0x4B000000u, 0x7640));     // bytes: raw.b0, 0x00, 0x00, 0x4B
                tv[j] = decode_u8_magic(__byte_perm(raw, 0x4B000000u, 0x7641));
            }
            else if constexpr (MODE == 1)
            {
                tu[j] = decode_u8_magic(0x4B000000u | __ldg(row + k.off[j]));
                tv[j] = decode_u8_magic(0x4B000000u | __ldg(row + k.off[j] + 1));
            }
            else
            {
                tu[j] = decode_u8_magic(0x4B000000u | __ldg(row + k.off[j] + 2));
                tv[j] = tu[j];
            }
        }
        float4 o;
        o.x = tap4(k.ca[0], k.ca[1], k.ca[2], k.ca[3], tu[0], tu[1], tu[2], tu[3]);
        o.y = tap4(k.ca[0], k.ca[1], k.ca[2], k.ca[3], tv[0], tv[1], tv[2], tv[3]);
        if (k.all_d1)
        {
            o.z = tap4(k.cb[0], k.cb[1], k.cb[2], k.cb[3], tu[1], tu[2], tu[3], tu[4]);
            o.w = tap4(k.cb[0], k.cb[1], k.cb[2], k.cb[3], tv[1], tv[2], tv[3], tv[4]);
        }
        else
        {
            const bool d = k.d != 0;
            o.z = tap4(k.cb[0], k.cb[1], k.cb[2], k.cb[3], d ? tu[1] : tu[0], d ? tu[2] : tu[1], d ? tu[3] : tu[2], d ? tu[4] : tu[3]);
            o.w = tap4(k.cb[0], k.cb[1], k.cb[2], k.cb[3], d ? tv[1] : tv[0], d ? tv[2] : tv[1], d ? tv[3] : tv[2], d ? tv[4] : tv[3]);
        }
        return o;
    }
    // One output pixel: stb encode of the resized (u, v) (x 255 + 0.5, clamp, truncate), toFloat of the stored byte, YUV -> RGB with
    // the luma `yv` (already toFloat of ITS stored byte), quant_u8.  Returns the three channel bytes in the low bytes of rb / gb / bb.
    __device__ __forceinline__ void chroma_merge_px(float su, float sv, float yv, uint32_t& rb, uint32_t& gb, uint32_t& bb)
    {
        const float fu = fminf(fmaxf(__fadd_rn(__fmul_rn(su, 255.0f), 0.5f), 0.0f), 255.0f);
        const float fv = fminf(fmaxf(__fadd_rn(__fmul_rn(sv, 255.0f), 0.5f), 0.0f), 255.0f);
        const float qu = unit_from_int<255>(__fsub_rn(__fadd_rz(fu, CM_MAGIC), CM_MAGIC));
        const float qv = unit_from_int<255>(__fsub_rn(__fadd_rz(fv, CM_MAGIC), CM_MAGIC));
        const float u = __fsub_rn(qu, 0.5f), v = __fsub_rn(qv, 0.5f);
        const float r = __fadd_rn(yv, __fmul_rn(1.403f, v));
        const float g = __fsub_rn(__fsub_rn(yv, __fmul_rn(0.344f, u)), __fmul_rn(0.714f, v));
        const float b = __fadd_rn(yv, __fmul_rn(1.773f, u));
        rb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(r), 255.0f), 0.5f), CM_MAGIC));
        gb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(g), 255.0f), 0.5f), CM_MAGIC));
        bb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(b), 255.0f), 0.5f), CM_MAGIC));
    }
    // the two RGB pixels of one output row of a lane (6 bytes at an even address) as three 16-bit stores
    __device__ __forceinline__ void store_rgb2(uint8_t* o, uint32_t r0, uint32_t g0, uint32_t b0, uint32_t r1, uint32_t g1, uint32_t b1)
    {
        unsigned short* o16 = reinterpret_cast<unsigned short*>(o);
        o16[0] = static_cast<unsigned short>(__byte_perm(r0, g0, 0x0040));
        o16[1] = static_cast<unsigned short>(__byte_perm(b0, r1, 0x0040));
        o16[2] = static_cast<unsigned short>(__byte_perm(g1, b1, 0x0040));
    }
    // rgba2yuva (ImageProcess.cpp:113-138): colour premultiplied by alpha, alpha kept as the third channel of the chroma plane
    __device__ __forceinline__ float luma_from_rgba_u8(uint32_t r, uint32_t g, uint32_t b, uint32_t a, uint8_t& qy, uint8_t& qu, uint8_t& qv, uint8_t& qa)
    {
        const float af = unit_from_int<255>(static_cast<float>(a));
        const YuvFromRgb o = rgb_to_yuv(__fmul_rn(unit_from_int<255>(static_cast<float>(r)), af), __fmul_rn(unit_from_int<255>(static_cast<float>(g)), af),
                                        __fmul_rn(unit_from_int<255>(static_cast<float>(b)), af));
        qy = quant_u8(o.y); qu = quant_u8(o.u); qv = quant_u8(o.v); qa = quant_u8(af);
        return unit_from_int<255>(static_cast<float>(qy));
    }
    // quantised luma byte of the colour split, as the network's toFloat reads it back (rgb2yuv_u8x4_kernel + load_elem)
    __device__ __forceinline__ float luma_from_rgb_u8(uint32_t r, uint32_t g, uint32_t b, uint8_t& qy, uint8_t& qu, uint8_t& qv)
    {
        const YuvFromRgb o = rgb_to_yuv(unit_from_int<255>(static_cast<float>(r)), unit_from_int<255>(static_cast<float>(g)), unit_from_int<255>(static_cast<float>(b)));
        qy = quant_u8(o.y); qu = quant_u8(o.u); qv = quant_u8(o.v);
        return unit_from_int<255>(static_cast<float>(qy));
    }
}
