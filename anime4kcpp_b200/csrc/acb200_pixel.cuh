// Colour split / merge and Catmull-Rom resize kernels of the hot path (HBM-bound, elementwise).
//
// Reference semantics:
//   rgb2yuv / rgba2yuva (2-plane)   core/src/ImageProcess.cpp:38-61, 113-138
//   yuv2rgb / yuva2rgba (2-plane)   core/src/ImageProcess.cpp:191-215, 275-308
//   resize(..., RESIZE_CATMULL_ROM) core/src/ImageResize.cpp:136-272 -> stb_image_resize2 gather upsample
// Every multiply/add below that the reference rounds separately uses __fmul_rn/__fadd_rn so nvcc cannot
// contract it into an FMA: the quantised planes must come out bit-identical to the CPU processor's.
#pragma once

#include "acb200_common.cuh"
#include "acb200_colour.cuh"

namespace acb
{
    // ystep / uvstep: element distance between horizontally adjacent samples of the Y and (U,V[,A]) outputs:
    // 1 and c-1 for the 2-plane form, c and c (uvp = yp + 1 element) for the packed YUV[A] form.
    __global__ void rgb2yuv_kernel(const void* __restrict__ src, int src_pitch, int w, int h, int c, int type,
                                   void* __restrict__ yp, int y_pitch, int ystep, void* __restrict__ uvp, int uv_pitch, int uvstep)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= w || y >= h) return;
        const uint8_t* in = static_cast<const uint8_t*>(src) + static_cast<size_t>(y) * src_pitch;
        float r = load_elem(in, x * c + 0, type), g = load_elem(in, x * c + 1, type), b = load_elem(in, x * c + 2, type);
        float a = 1.0f;
        if (c == 4)
        {
            a = load_elem(in, x * c + 3, type);
            r = __fmul_rn(r, a); g = __fmul_rn(g, a); b = __fmul_rn(b, a);
        }
        const YuvFromRgb o = rgb_to_yuv(r, g, b);
        uint8_t* yo = static_cast<uint8_t*>(yp) + static_cast<size_t>(y) * y_pitch;
        uint8_t* uvo = static_cast<uint8_t*>(uvp) + static_cast<size_t>(y) * uv_pitch;
        store_elem(yo, x * ystep, type, o.y);
        store_elem(uvo, x * uvstep + 0, type, o.u);
        store_elem(uvo, x * uvstep + 1, type, o.v);
        if (c == 4) store_elem(uvo, x * uvstep + 2, type, a);
    }

    // 8-bit RGB -> quantised Y plane + interleaved (u, v) plane, four pixels per thread: three 32-bit loads, one 32-bit and one
    // 64-bit store.  Same arithmetic as rgb2yuv_kernel value for value.  Requires 4-byte aligned rows (pointers and pitches).
    __global__ void __launch_bounds__(256) rgb2yuv_u8x4_kernel(const uint8_t* __restrict__ src, int src_pitch, int w, int h,
                                                               uint8_t* __restrict__ yp, int y_pitch, uint8_t* __restrict__ uvp, int uv_pitch)
    {
        const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x4 >= w || y >= h) return;
        const uint8_t* in = src + static_cast<size_t>(y) * src_pitch + 3 * x4;
        uint8_t* yo = yp + static_cast<size_t>(y) * y_pitch + x4;
        uint8_t* uvo = uvp + static_cast<size_t>(y) * uv_pitch + 2 * x4;
        uint8_t px[12], qy[4], quv[8];
        const int n = min(4, w - x4);
        if (n == 4)
        {
            const uint32_t* in4 = reinterpret_cast<const uint32_t*>(in);
            const uint32_t w0 = __ldg(in4), w1 = __ldg(in4 + 1), w2 = __ldg(in4 + 2);
#pragma unroll
            for (int k = 0; k < 4; k++) { px[k] = (w0 >> (8 * k)) & 0xff; px[4 + k] = (w1 >> (8 * k)) & 0xff; px[8 + k] = (w2 >> (8 * k)) & 0xff; }
        }
        else
            for (int k = 0; k < 12; k++) px[k] = k < 3 * n ? in[k] : 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const YuvFromRgb o = rgb_to_yuv(unit_from_int<255>(static_cast<float>(px[3 * k])), unit_from_int<255>(static_cast<float>(px[3 * k + 1])),
                                            unit_from_int<255>(static_cast<float>(px[3 * k + 2])));
            qy[k] = quant_u8(o.y); quv[2 * k] = quant_u8(o.u); quv[2 * k + 1] = quant_u8(o.v);
        }
        if (n == 4)
        {
            *reinterpret_cast<uint32_t*>(yo) = qy[0] | (qy[1] << 8) | (qy[2] << 16) | (static_cast<uint32_t>(qy[3]) << 24);
            *reinterpret_cast<uint2*>(uvo) = make_uint2(quv[0] | (quv[1] << 8) | (quv[2] << 16) | (static_cast<uint32_t>(quv[3]) << 24),
                                                        quv[4] | (quv[5] << 8) | (quv[6] << 16) | (static_cast<uint32_t>(quv[7]) << 24));
        }
        else
            for (int k = 0; k < n; k++) { yo[k] = qy[k]; uvo[2 * k] = quv[2 * k]; uvo[2 * k + 1] = quv[2 * k + 1]; }
    }

    // r,g,b (+a) from y and the already-decoded u,v(,a) of the same element type
    __device__ __forceinline__ void yuv_to_rgb_store(void* out_row, int x, int c, int type, float y, float uq, float vq, float aq)
    {
        const float u = __fsub_rn(uq, 0.5f), v = __fsub_rn(vq, 0.5f);
        float r = __fadd_rn(y, __fmul_rn(1.403f, v));
        float g = __fsub_rn(__fsub_rn(y, __fmul_rn(0.344f, u)), __fmul_rn(0.714f, v));
        float b = __fadd_rn(y, __fmul_rn(1.773f, u));
        if (c == 4)
        {
            if (aq > 1e-6f) { r = __fdiv_rn(r, aq); g = __fdiv_rn(g, aq); b = __fdiv_rn(b, aq); }
            else r = g = b = 0.0f;
            store_elem(out_row, x * c + 3, type, aq);
        }
        store_elem(out_row, x * c + 0, type, r);
        store_elem(out_row, x * c + 1, type, g);
        store_elem(out_row, x * c + 2, type, b);
    }

    __global__ void yuv2rgb_kernel(const void* __restrict__ yp, int y_pitch, int ystep, const void* __restrict__ uvp, int uv_pitch, int uvstep,
                                   int w, int h, int c, int type, void* __restrict__ dst, int dst_pitch)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= w || y >= h) return;
        const uint8_t* yi = static_cast<const uint8_t*>(yp) + static_cast<size_t>(y) * y_pitch;
        const uint8_t* uvi = static_cast<const uint8_t*>(uvp) + static_cast<size_t>(y) * uv_pitch;
        const float a = c == 4 ? load_elem(uvi, x * uvstep + 2, type) : 1.0f;
        yuv_to_rgb_store(static_cast<uint8_t*>(dst) + static_cast<size_t>(y) * dst_pitch, x, c, type,
                         load_elem(yi, x * ystep, type), load_elem(uvi, x * uvstep + 0, type), load_elem(uvi, x * uvstep + 1, type), a);
    }

    // stb_image_resize2 decode: integers are scaled by a reciprocal *multiply* (unlike toFloat's division)
    __device__ __forceinline__ float resize_decode(const void* row, int x, int type)
    {
        switch (type)
        {
        case ACB200_UINT8: return __fmul_rn(static_cast<float>(static_cast<const uint8_t*>(row)[x]), 1.0f / 255.0f);
        case ACB200_UINT16: return __fmul_rn(static_cast<float>(static_cast<const uint16_t*>(row)[x]), 1.0f / 65535.0f);
        case ACB200_FLOAT16: return __half2float(static_cast<const __half*>(row)[x]);
        default: return static_cast<const float*>(row)[x];
        }
    }
    // stb encode: v*max + 0.5, clamp to [0,max], truncate; floats are stored unclamped.
    // Returns the value as the next stage's toFloat() will read it back.
    __device__ __forceinline__ float resize_encode_store(void* row, int x, int type, float s, bool do_store)
    {
        switch (type)
        {
        case ACB200_UINT8:
        {
            float f = __fadd_rn(__fmul_rn(s, 255.0f), 0.5f);
            f = f < 0.0f ? 0.0f : (f > 255.0f ? 255.0f : f);
            const uint8_t q = static_cast<uint8_t>(f);
            if (do_store) static_cast<uint8_t*>(row)[x] = q;
            return unit_from_int<255>(static_cast<float>(q));
        }
        case ACB200_UINT16:
        {
            float f = __fadd_rn(__fmul_rn(s, 65535.0f), 0.5f);
            f = f < 0.0f ? 0.0f : (f > 65535.0f ? 65535.0f : f);
            const uint16_t q = static_cast<uint16_t>(f);
            if (do_store) static_cast<uint16_t*>(row)[x] = q;
            return unit_from_int<65535>(static_cast<float>(q));
        }
        case ACB200_FLOAT16:
        {
            const __half q = __float2half_rn(s);
            if (do_store) static_cast<__half*>(row)[x] = q;
            return __half2float(q);
        }
        default:
            if (do_store) static_cast<float*>(row)[x] = s;
            return s;
        }
    }

    // horizontal pass then vertical pass for one output sample of channel `ch` (left-to-right sums,
    // products and sums rounded separately, as the CPU restatement does)
    __device__ __forceinline__ float catmull_sample(const void* __restrict__ src, int src_pitch, int c, int ch, int type,
                                                    const Contrib& hc, const Contrib& vc)
    {
        float s = 0.0f;
        for (int i = 0; i < vc.cnt; i++)
        {
            const uint8_t* row = static_cast<const uint8_t*>(src) + static_cast<size_t>(vc.n0 + i) * src_pitch;
            float hsum = 0.0f;
            for (int j = 0; j < hc.cnt; j++)
            {
                const float t = __fmul_rn(hc.c[j], resize_decode(row, (hc.n0 + j) * c + ch, type));
                hsum = j == 0 ? t : __fadd_rn(hsum, t);
            }
            const float t = __fmul_rn(vc.c[i], hsum);
            s = i == 0 ? t : __fadd_rn(s, t);
        }
        return s;
    }

    __global__ void resize_catmull_kernel(const void* __restrict__ src, int src_pitch, int c, int type,
                                          const Contrib* __restrict__ htab, const Contrib* __restrict__ vtab,
                                          void* __restrict__ dst, int ow, int oh, int dst_pitch)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= ow || y >= oh) return;
        const Contrib hc = htab[x], vc = vtab[y];
        void* row = static_cast<uint8_t*>(dst) + static_cast<size_t>(y) * dst_pitch;
        for (int ch = 0; ch < c; ch++)
            resize_encode_store(row, x * c + ch, type, catmull_sample(src, src_pitch, c, ch, type, hc, vc), true);
    }

    // ac::core::shl / shr on an integer plane (core/src/ImageProcess.cpp:601-616): `a << n` / `a >> n` evaluated in int and
    // truncated to the element type on store.  left != 0: shl from src into dst; left == 0: shr (src may equal dst).
    __global__ void shift_kernel(const void* __restrict__ src, int src_pitch, void* __restrict__ dst, int dst_pitch, int n_elems, int h, int es, int n, int left)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= n_elems || y >= h) return;
        const uint8_t* in = static_cast<const uint8_t*>(src) + static_cast<size_t>(y) * src_pitch;
        uint8_t* out = static_cast<uint8_t*>(dst) + static_cast<size_t>(y) * dst_pitch;
        if (es == 1)
        {
            const int a = in[x];
            out[x] = static_cast<uint8_t>(left ? a << n : a >> n);
        }
        else
        {
            const int a = reinterpret_cast<const uint16_t*>(in)[x];
            reinterpret_cast<uint16_t*>(out)[x] = static_cast<uint16_t>(left ? a << n : a >> n);
        }
    }

    // Contributors of the general resize (up to 10 taps): the post-network luma down-scale of non-power-of-two factors
    // (Processor.cpp:237,249: resize(out, ..., fxy) with 1/2 < fxy < 1 gathers up to 9 source pixels per axis)
    struct ContribW
    {
        int n0, cnt;
        float c[10];
    };
    __global__ void resize_wide_kernel(const void* __restrict__ src, int src_pitch, int c, int type,
                                       const ContribW* __restrict__ htab, const ContribW* __restrict__ vtab,
                                       void* __restrict__ dst, int ow, int oh, int dst_pitch)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= ow || y >= oh) return;
        const ContribW hc = htab[x], vc = vtab[y];
        void* row = static_cast<uint8_t*>(dst) + static_cast<size_t>(y) * dst_pitch;
        for (int ch = 0; ch < c; ch++)
        {
            // horizontal pass then vertical pass, left-to-right sums with separately rounded products (as catmull_sample)
            float s = 0.0f;
            for (int i = 0; i < vc.cnt; i++)
            {
                const uint8_t* srow = static_cast<const uint8_t*>(src) + static_cast<size_t>(vc.n0 + i) * src_pitch;
                float hsum = 0.0f;
                for (int j = 0; j < hc.cnt; j++)
                {
                    const float t = __fmul_rn(hc.c[j], resize_decode(srow, (hc.n0 + j) * c + ch, type));
                    hsum = j == 0 ? t : __fadd_rn(hsum, t);
                }
                const float t = __fmul_rn(vc.c[i], hsum);
                s = i == 0 ? t : __fadd_rn(s, t);
            }
            resize_encode_store(row, x * c + ch, type, s, true);
        }
    }

    // Processor.cpp:251-253 fused: Catmull-Rom upscale of the (u,v[,a]) plane by the full factor, re-quantised to
    // the element type exactly where the reference materialises the resized plane, then YUV->RGB merge with the
    // network's luma.
    __global__ void chroma_merge_kernel(const void* __restrict__ yp, int y_pitch, const void* __restrict__ uvp, int uv_pitch,
                                        const Contrib* __restrict__ htab, const Contrib* __restrict__ vtab,
                                        int ow, int oh, int c, int type, void* __restrict__ dst, int dst_pitch)
    {
        const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= ow || y >= oh) return;
        const Contrib hc = htab[x], vc = vtab[y];
        const int uvc = c - 1;
        float q[3] = { 0.0f, 0.0f, 1.0f };
        for (int ch = 0; ch < uvc; ch++)
            q[ch] = resize_encode_store(nullptr, 0, type, catmull_sample(uvp, uv_pitch, uvc, ch, type, hc, vc), false);
        const float yv = load_elem(static_cast<const uint8_t*>(yp) + static_cast<size_t>(y) * y_pitch, x, type);
        yuv_to_rgb_store(static_cast<uint8_t*>(dst) + static_cast<size_t>(y) * dst_pitch, x, c, type, yv, q[0], q[1], q[2]);
    }

    // ---- 8-bit fast path of the fused chroma resize + merge ------------------------------------------------------------
    // One CTA produces a CM_OW x CM_OH tile of the output: it stages the decoded source (u,v[,a]) tile in shared memory, runs
    // the horizontal pass ONCE per source row (not once per output row), then each thread does the vertical pass, the
    // re-quantisation and the YUV->RGB merge for 4 adjacent output pixels and writes them as 32-bit words.
    // Same arithmetic, rounding points and summation order as chroma_merge_kernel (the general path); taps beyond a
    // contributor's count carry a zero coefficient, which leaves every partial sum unchanged.  Requires cnt <= 4 (true for
    // every upscale: the Catmull-Rom support spans 4 source pixels).
    constexpr int CM_OW = 64, CM_OH = 16, CM_THREADS = 256;
    constexpr int CM_SRC_W = CM_OW + 8, CM_SRC_H = CM_OH + 8;      // worst case (scale -> 1): tile + support + unrolled-tap slack

    template<int C>
    __global__ void __launch_bounds__(CM_THREADS) chroma_merge_u8_kernel(const uint8_t* __restrict__ yp, int y_pitch, const uint8_t* __restrict__ uvp, int uv_pitch,
                                                                      int sw_img, int sh_img,
                                                                      const Contrib* __restrict__ htab, const Contrib* __restrict__ vtab,
                                                                      int ow, int oh, uint8_t* __restrict__ dst, int dst_pitch)
    {
        constexpr int UVC = C - 1;
        __shared__ __align__(16) Contrib sh_v[CM_OH];
        __shared__ __align__(16) float s_src[CM_SRC_H][CM_SRC_W * UVC];
        __shared__ __align__(16) float s_hp[CM_SRC_H][CM_OW * UVC];
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int ox0 = blockIdx.x * CM_OW, oy0 = blockIdx.y * CM_OH;
        const int ncols = min(CM_OW, ow - ox0), nrows = min(CM_OH, oh - oy0);
        if (tid < nrows * 8) reinterpret_cast<uint32_t*>(sh_v)[tid] = reinterpret_cast<const uint32_t*>(vtab + oy0)[tid];    // the tile's vertical taps, once
        // source window of the tile (tables are monotonic in n0)
        // this thread's luma samples of the merge (4 output pixels): requested now, consumed after both resize passes
        uint8_t ylum[CM_OW * CM_OH / CM_THREADS];
#pragma unroll
        for (int k = 0; k < CM_OW * CM_OH / CM_THREADS; k++)
        {
            const int idx = tid + k * CM_THREADS, orow = min(idx / CM_OW, nrows - 1), col = min(idx % CM_OW, ncols - 1);
            ylum[k] = __ldg(yp + static_cast<size_t>(oy0 + orow) * y_pitch + ox0 + col);
        }
        const Contrib hfirst = htab[ox0], hlast = htab[ox0 + ncols - 1], vfirst = vtab[oy0], vlast = vtab[oy0 + nrows - 1];
        const int sx0 = hfirst.n0, sy0 = vfirst.n0;
        const int sw = min(hlast.n0 + 4, sw_img) - sx0 + 0, shh = min(vlast.n0 + 4, sh_img) - sy0;
        // decoded source tile (stb decode: q * (1/255), a multiply); columns / rows past the image stay zero-weighted
        const int drows = min(vlast.n0 + 4 - sy0, CM_SRC_H), dcols = min(hlast.n0 + 4 - sx0, CM_SRC_W) * UVC;
        for (int row = warp; row < drows; row += CM_THREADS / 32)
            for (int e = lane; e < dcols; e += 32)
            {
                float v = 0.0f;
                if (row < shh && e < sw * UVC) v = __fmul_rn(static_cast<float>(uvp[static_cast<size_t>(sy0 + row) * uv_pitch + sx0 * UVC + e]), 1.0f / 255.0f);
                s_src[row][e] = v;
            }
        // this lane's two output columns of the horizontal pass
        Contrib hk[2];
#pragma unroll
        for (int j = 0; j < 2; j++)
        {
            const int col = min(lane + 32 * j, ncols - 1);
            hk[j] = htab[ox0 + col];
            hk[j].n0 -= sx0;
        }
        __syncthreads();
        // rows / columns past the image are zero in s_src and carry zero coefficients: run the pass over them too so the
        // vertical pass never multiplies 0 by uninitialised shared memory
        const int hrows = min(vlast.n0 + 4 - sy0, CM_SRC_H);
        for (int row = warp; row < hrows; row += CM_THREADS / 32)
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const float* srow = &s_src[row][hk[j].n0 * UVC];
                if constexpr (UVC == 2)
                {
                    // both channels of a tap in one 8-byte access (half the shared-memory wavefronts of two scalar ones)
                    const float2 t0 = *reinterpret_cast<const float2*>(srow), t1 = *reinterpret_cast<const float2*>(srow + 2);
                    const float2 t2 = *reinterpret_cast<const float2*>(srow + 4), t3 = *reinterpret_cast<const float2*>(srow + 6);
                    float2 hs;
                    hs.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hk[j].c[0], t0.x), __fmul_rn(hk[j].c[1], t1.x)), __fmul_rn(hk[j].c[2], t2.x)), __fmul_rn(hk[j].c[3], t3.x));
                    hs.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hk[j].c[0], t0.y), __fmul_rn(hk[j].c[1], t1.y)), __fmul_rn(hk[j].c[2], t2.y)), __fmul_rn(hk[j].c[3], t3.y));
                    *reinterpret_cast<float2*>(&s_hp[row][(lane + 32 * j) * 2]) = hs;
                }
                else
                {
#pragma unroll
                    for (int ch = 0; ch < UVC; ch++)
                    {
                        float hsum = __fmul_rn(hk[j].c[0], srow[ch]);
                        hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[1], srow[UVC + ch]));
                        hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[2], srow[2 * UVC + ch]));
                        hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[3], srow[3 * UVC + ch]));
                        s_hp[row][(lane + 32 * j) * UVC + ch] = hsum;
                    }
                }
            }
        __syncthreads();
        // vertical pass + re-quantise + merge: consecutive lanes take consecutive output columns (conflict-free reads of the
        // horizontal-pass rows); the finished bytes are staged in shared memory and leave as 16-byte vectors
        __shared__ __align__(16) uint8_t s_out[CM_OH][CM_OW * C];
        if constexpr (C == 3)
        {
            // RGB: the hot path, written for instruction count (every step of the arithmetic is the one of the general loop
            // below, value for value).  The contributor comes in as two vector loads; clamps are min / max; trunc(f), the
            // float -> byte conversions and byte -> float run on the FMA pipe through the 2^23 trick instead of the quarter-rate
            // conversion unit (exact for 0 <= f < 2^23: a round-toward-zero add of 2^23 leaves trunc(f) in the low mantissa bits);
            // no per-pixel branch: rows / columns outside a partial tile are computed from clamped indices and not stored.
            constexpr float MAGIC = 8388608.0f;
            const int col = tid & (CM_OW - 1), rg = tid / CM_OW;
            const bool col_ok = col < ncols;
            const float2* hbase = reinterpret_cast<const float2*>(&s_hp[0][0]) + col;
#pragma unroll
            for (int it = 0; it < CM_OW * CM_OH / CM_THREADS; it++)
            {
                const int orow = rg + (CM_THREADS / CM_OW) * it;
                const Contrib* kp = &sh_v[min(orow, nrows - 1)];
                const uint4 ka = *reinterpret_cast<const uint4*>(kp);               // n0, cnt, c[0], c[1]
                const float2 kb = *reinterpret_cast<const float2*>(&kp->c[2]);      // c[2], c[3]
                const float c0 = __uint_as_float(ka.z), c1 = __uint_as_float(ka.w), c2 = kb.x, c3 = kb.y;
                const float2* h = hbase + (static_cast<int>(ka.x) - sy0) * CM_OW;
                const float2 t0 = h[0], t1 = h[CM_OW], t2 = h[2 * CM_OW], t3 = h[3 * CM_OW];
                const float su = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.x), __fmul_rn(c1, t1.x)), __fmul_rn(c2, t2.x)), __fmul_rn(c3, t3.x));
                const float sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.y), __fmul_rn(c1, t1.y)), __fmul_rn(c2, t2.y)), __fmul_rn(c3, t3.y));
                // stb encode (x 255 + 0.5, clamp, truncate), then toFloat of the stored byte
                const float fu = fminf(fmaxf(__fadd_rn(__fmul_rn(su, 255.0f), 0.5f), 0.0f), 255.0f);
                const float fv = fminf(fmaxf(__fadd_rn(__fmul_rn(sv, 255.0f), 0.5f), 0.0f), 255.0f);
                const float qu = unit_from_int<255>(__fsub_rn(__fadd_rz(fu, MAGIC), MAGIC));
                const float qv = unit_from_int<255>(__fsub_rn(__fadd_rz(fv, MAGIC), MAGIC));
                const float yv = unit_from_int<255>(__fsub_rn(__uint_as_float(0x4B000000u | ylum[it]), MAGIC));
                const float u = __fsub_rn(qu, 0.5f), v = __fsub_rn(qv, 0.5f);
                const float r = __fadd_rn(yv, __fmul_rn(1.403f, v));
                const float gch = __fsub_rn(__fsub_rn(yv, __fmul_rn(0.344f, u)), __fmul_rn(0.714f, v));
                const float b = __fadd_rn(yv, __fmul_rn(1.773f, u));
                // quant_u8: sat01, x 255 + 0.5, truncate -- the byte is the low mantissa byte of the magic sum
                const uint32_t rb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(r), 255.0f), 0.5f), MAGIC));
                const uint32_t gb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(gch), 255.0f), 0.5f), MAGIC));
                const uint32_t bb = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(b), 255.0f), 0.5f), MAGIC));
                if (col_ok && orow < nrows)
                {
                    uint8_t* o = &s_out[orow][col * 3];
                    o[0] = static_cast<uint8_t>(rb); o[1] = static_cast<uint8_t>(gb); o[2] = static_cast<uint8_t>(bb);
                }
            }
        }
        else
#pragma unroll
        for (int it = 0; it < CM_OW * CM_OH / CM_THREADS; it++)
        {
            const int idx = tid + it * CM_THREADS;
            const int orow = idx / CM_OW, col = idx % CM_OW;
            if (orow >= nrows || col >= ncols) continue;
            const Contrib& k = sh_v[orow];
            const int r0 = k.n0 - sy0;
            float q[3] = { 0.0f, 0.0f, 1.0f };
            float vs[UVC];
            if constexpr (UVC == 2)
            {
                const float* hcol = &s_hp[r0][col * 2];
                const float2 t0 = *reinterpret_cast<const float2*>(hcol), t1 = *reinterpret_cast<const float2*>(hcol + CM_OW * 2);
                const float2 t2 = *reinterpret_cast<const float2*>(hcol + 2 * CM_OW * 2), t3 = *reinterpret_cast<const float2*>(hcol + 3 * CM_OW * 2);
                vs[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(k.c[0], t0.x), __fmul_rn(k.c[1], t1.x)), __fmul_rn(k.c[2], t2.x)), __fmul_rn(k.c[3], t3.x));
                vs[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(k.c[0], t0.y), __fmul_rn(k.c[1], t1.y)), __fmul_rn(k.c[2], t2.y)), __fmul_rn(k.c[3], t3.y));
            }
            else
            {
#pragma unroll
                for (int ch = 0; ch < UVC; ch++)
                {
                    const float* hcol = &s_hp[r0][col * UVC + ch];
                    float sum = __fmul_rn(k.c[0], hcol[0]);
                    sum = __fadd_rn(sum, __fmul_rn(k.c[1], hcol[CM_OW * UVC]));
                    sum = __fadd_rn(sum, __fmul_rn(k.c[2], hcol[2 * CM_OW * UVC]));
                    sum = __fadd_rn(sum, __fmul_rn(k.c[3], hcol[3 * CM_OW * UVC]));
                    vs[ch] = sum;
                }
            }
#pragma unroll
            for (int ch = 0; ch < UVC; ch++)
            {
                float f = __fadd_rn(__fmul_rn(vs[ch], 255.0f), 0.5f);   // stb encode, then toFloat of the stored byte
                f = f < 0.0f ? 0.0f : (f > 255.0f ? 255.0f : f);
                q[ch] = unit_from_int<255>(truncf(f));                  // toFloat<u8>: a true division (Util.hpp:53-54), computed without one
            }
            const float yv = unit_from_int<255>(static_cast<float>(ylum[it]));
            const float u = __fsub_rn(q[0], 0.5f), v = __fsub_rn(q[1], 0.5f);
            float r = __fadd_rn(yv, __fmul_rn(1.403f, v));
            float gch = __fsub_rn(__fsub_rn(yv, __fmul_rn(0.344f, u)), __fmul_rn(0.714f, v));
            float b = __fadd_rn(yv, __fmul_rn(1.773f, u));
            uint8_t* o = &s_out[orow][col * C];
            if (C == 4)
            {
                if (q[2] > 1e-6f) { r = __fdiv_rn(r, q[2]); gch = __fdiv_rn(gch, q[2]); b = __fdiv_rn(b, q[2]); }
                else r = gch = b = 0.0f;
                o[3] = quant_u8(q[2]);
            }
            o[0] = quant_u8(r); o[1] = quant_u8(gch); o[2] = quant_u8(b);
        }
        __syncthreads();
        uint8_t* tile_dst = dst + static_cast<size_t>(oy0) * dst_pitch + static_cast<size_t>(ox0) * C;
        if (ncols == CM_OW && ((reinterpret_cast<uintptr_t>(tile_dst) | static_cast<uintptr_t>(dst_pitch)) & 15) == 0)
        {
            constexpr int VEC_PER_ROW = CM_OW * C / 16;
            for (int i = tid; i < nrows * VEC_PER_ROW; i += CM_THREADS)
            {
                const int row = i / VEC_PER_ROW, v = i % VEC_PER_ROW;
                reinterpret_cast<uint4*>(tile_dst + static_cast<size_t>(row) * dst_pitch)[v] = reinterpret_cast<const uint4*>(s_out[row])[v];
            }
        }
        else
        {
            for (int i = tid; i < nrows * ncols * C; i += CM_THREADS)
            {
                const int row = i / (ncols * C), e = i % (ncols * C);
                tile_dst[static_cast<size_t>(row) * dst_pitch + e] = s_out[row][e];
            }
        }
    }
    // ---- integer fast path of the stand-alone Catmull-Rom upscale (video chroma planes) -------------------------------------
    // Same tiling and arithmetic as chroma_merge_u8_kernel without the merge: T = uint8_t / uint16_t samples, NC = 1 (planar
    // U or V) or 2 (interleaved UV) channels; blockIdx.z selects one of up to two equally sized planes, so the U and V
    // planes of an I420 / I444 frame are ONE launch.  Bit-identical to resize_catmull_kernel (same products, sums and rounding
    // points; the extra taps of a short contributor carry zero coefficients).
    struct ResizePlanes
    {
        const void* src[2];
        void* dst[2];
        int src_pitch[2], dst_pitch[2];
    };

    template<typename T, int NC>
    __global__ void __launch_bounds__(CM_THREADS) resize_tile_kernel(const ResizePlanes pl, int sw_img, int sh_img,
                                                                  const Contrib* __restrict__ htab, const Contrib* __restrict__ vtab, int ow, int oh)
    {
        constexpr float MAXV = sizeof(T) == 1 ? 255.0f : 65535.0f;
        constexpr float RCP = 1.0f / MAXV;
        __shared__ Contrib sh_v[CM_OH];
        __shared__ float s_src[CM_SRC_H][CM_SRC_W * NC];
        __shared__ float s_hp[CM_SRC_H][CM_OW * NC];
        __shared__ __align__(16) T s_out[CM_OH][CM_OW * NC];
        const T* __restrict__ src = static_cast<const T*>(pl.src[blockIdx.z]);
        T* __restrict__ dst = static_cast<T*>(pl.dst[blockIdx.z]);
        const int src_pitch = pl.src_pitch[blockIdx.z], dst_pitch = pl.dst_pitch[blockIdx.z];
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int ox0 = blockIdx.x * CM_OW, oy0 = blockIdx.y * CM_OH;
        const int ncols = min(CM_OW, ow - ox0), nrows = min(CM_OH, oh - oy0);
        if (tid < nrows * 8) reinterpret_cast<uint32_t*>(sh_v)[tid] = reinterpret_cast<const uint32_t*>(vtab + oy0)[tid];
        const Contrib hfirst = htab[ox0], hlast = htab[ox0 + ncols - 1], vfirst = vtab[oy0], vlast = vtab[oy0 + nrows - 1];
        const int sx0 = hfirst.n0, sy0 = vfirst.n0;
        const int sw = min(hlast.n0 + 4, sw_img) - sx0, shh = min(vlast.n0 + 4, sh_img) - sy0;
        const int drows = min(vlast.n0 + 4 - sy0, CM_SRC_H), dcols = min(hlast.n0 + 4 - sx0, CM_SRC_W) * NC;
        for (int row = warp; row < drows; row += CM_THREADS / 32)
        {
            const T* srow = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(src) + static_cast<size_t>(sy0 + row) * src_pitch) + sx0 * NC;
            for (int e = lane; e < dcols; e += 32)
                s_src[row][e] = (row < shh && e < sw * NC) ? __fmul_rn(static_cast<float>(srow[e]), RCP) : 0.0f;
        }
        Contrib hk[2];
#pragma unroll
        for (int j = 0; j < 2; j++)
        {
            hk[j] = htab[ox0 + min(lane + 32 * j, ncols - 1)];
            hk[j].n0 -= sx0;
        }
        __syncthreads();
        for (int row = warp; row < drows; row += CM_THREADS / 32)
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const float* srow = &s_src[row][hk[j].n0 * NC];
#pragma unroll
                for (int ch = 0; ch < NC; ch++)
                {
                    float hsum = __fmul_rn(hk[j].c[0], srow[ch]);
                    hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[1], srow[NC + ch]));
                    hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[2], srow[2 * NC + ch]));
                    hsum = __fadd_rn(hsum, __fmul_rn(hk[j].c[3], srow[3 * NC + ch]));
                    s_hp[row][(lane + 32 * j) * NC + ch] = hsum;
                }
            }
        __syncthreads();
        for (int idx = tid; idx < CM_OW * CM_OH; idx += CM_THREADS)
        {
            const int orow = idx / CM_OW, col = idx % CM_OW;
            if (orow >= nrows || col >= ncols) continue;
            const Contrib& k = sh_v[orow];
            const int r0 = k.n0 - sy0;
#pragma unroll
            for (int ch = 0; ch < NC; ch++)
            {
                const float* hcol = &s_hp[r0][col * NC + ch];
                float sum = __fmul_rn(k.c[0], hcol[0]);
                sum = __fadd_rn(sum, __fmul_rn(k.c[1], hcol[CM_OW * NC]));
                sum = __fadd_rn(sum, __fmul_rn(k.c[2], hcol[2 * CM_OW * NC]));
                sum = __fadd_rn(sum, __fmul_rn(k.c[3], hcol[3 * CM_OW * NC]));
                float f = __fadd_rn(__fmul_rn(sum, MAXV), 0.5f);
                f = f < 0.0f ? 0.0f : (f > MAXV ? MAXV : f);
                s_out[orow][col * NC + ch] = static_cast<T>(f);
            }
        }
        __syncthreads();
        uint8_t* tile_dst = reinterpret_cast<uint8_t*>(dst) + static_cast<size_t>(oy0) * dst_pitch + static_cast<size_t>(ox0) * NC * sizeof(T);
        if (ncols == CM_OW && ((reinterpret_cast<uintptr_t>(tile_dst) | static_cast<uintptr_t>(dst_pitch)) & 15) == 0)
        {
            constexpr int VEC_PER_ROW = CM_OW * NC * static_cast<int>(sizeof(T)) / 16;
            for (int i = tid; i < nrows * VEC_PER_ROW; i += CM_THREADS)
            {
                const int row = i / VEC_PER_ROW, v = i % VEC_PER_ROW;
                reinterpret_cast<uint4*>(tile_dst + static_cast<size_t>(row) * dst_pitch)[v] = reinterpret_cast<const uint4*>(s_out[row])[v];
            }
        }
        else
        {
            for (int i = tid; i < nrows * ncols * NC; i += CM_THREADS)
            {
                const int row = i / (ncols * NC), e = i % (ncols * NC);
                reinterpret_cast<T*>(tile_dst + static_cast<size_t>(row) * dst_pitch)[e] = s_out[row][e];
            }
        }
    }
}
