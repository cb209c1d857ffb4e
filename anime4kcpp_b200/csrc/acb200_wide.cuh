// ArtCNN<16/32> and FSRCNNX<8/16>: the reference's remaining model families, first version.
//
// These nets have F = 8 / 16 / 32 features, so one layer's activations no longer fit the fused shared-memory scheme of the
// 8-feature engines.  Each layer is one (or, for 32 output channels, two) launches over fp32 maps `[h][w][F]` in HBM -- the
// structure of the reference's own CUDA backend (core/src/processor/cuda/Kernel.cu) -- computed in fp32 FFMA in exactly the
// summation order of the reference's 256-bit FMA backend (X86/AVX.hpp:32-58,95-146): lane c%8 runs one FMA chain over
// (tap, 8-channel chunk), then the hsum tree, then the bias.  Results are bit-identical to the reference's `create("cpu", 4)`.
//
//  * wide_head_kernel      conv3x3_cin1 / conv5x5_cin1, Identity            (Common.hpp:166-221)
//  * wide_conv_kernel      conv3x3_float<F,F> + ReLU / PReLU / Identity (+ residual, scale 1) or the F->4 pixel-shuffle tail
//                          (Common.hpp:116-164, 290-342)
//  * wide_pointwise_kernel the 1x1 + feat + PReLU that ends the FSRCNNX body (Common.hpp:223-288 with postactive1x1)
//
// A thread owns one pixel and four output channels at a time: 32 lane accumulators, the inputs come from a shared-memory tile
// as two 16-byte loads per (tap, chunk) and feed 32 FFMAs, the weights are kernel-parameter constants.
#pragma once

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"

namespace acb
{
    constexpr int WIDE_TW = 32, WIDE_TH = 32;       // output pixels per CTA
    constexpr int WIDE_THREADS = 256;               // 32 columns x 8 rows per sweep, 4 sweeps
    constexpr int WIDE_GROUP = 4;                   // output channels a thread accumulates at once

    enum { WIDE_STORE = 0, WIDE_SHUFFLE = 1 };

    template<int F, int NCO>
    struct WideConvParams
    {
        const float* in;        // [h][w][F]
        float* out;             // [h][w][F] (WIDE_STORE)
        const float* res;       // added after the activation with scale 1.0 (ArtCNN long skip) or null
        void* dst;              // output plane (WIDE_SHUFFLE)
        int dst_pitch, type;
        int w, h;
        int co0;                // first output channel of this launch
        int act;
        float k[NCO * 9 * F];   // [NCO][tap][cin]
        float b[NCO];
        float a[NCO];
    };

    template<int F>
    constexpr size_t wide_smem_bytes() { return static_cast<size_t>(WIDE_TH + 2) * (F / 4) * (WIDE_TW + 2) * sizeof(float4); }

    template<int F, int NCO, int MODE>
    __global__ void __launch_bounds__(WIDE_THREADS) wide_conv_kernel(const __grid_constant__ WideConvParams<F, NCO> prm)
    {
        constexpr int NCH = F / 4, SW = WIDE_TW + 2, SH = WIDE_TH + 2;
        extern __shared__ __align__(16) unsigned char wide_smem[];
        float4* tile = reinterpret_cast<float4*>(wide_smem);        // [SH][NCH][SW]
        const int x0 = blockIdx.x * WIDE_TW, y0 = blockIdx.y * WIDE_TH;
        // input tile with clamp-to-edge coordinates (Common.hpp:122-143)
        for (int i = threadIdx.x; i < SH * SW * NCH; i += WIDE_THREADS)
        {
            const int c = i % NCH, t = i / NCH, tx = t % SW, ty = t / SW;
            const int gx = clampi(x0 - 1 + tx, 0, prm.w - 1), gy = clampi(y0 - 1 + ty, 0, prm.h - 1);
            tile[(ty * NCH + c) * SW + tx] = __ldg(reinterpret_cast<const float4*>(prm.in + (static_cast<size_t>(gy) * prm.w + gx) * F) + c);
        }
        __syncthreads();
        const int lx = threadIdx.x & 31, ry = threadIdx.x >> 5;
        const int es = prm.type & 0xff;
        const bool aligned = MODE == WIDE_SHUFFLE && ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        for (int ly = ry; ly < WIDE_TH; ly += WIDE_THREADS / 32)
        {
            const int gx = x0 + lx, gy = y0 + ly;
            if (gx >= prm.w || gy >= prm.h) continue;
#pragma unroll 1
            for (int cg = 0; cg < NCO; cg += WIDE_GROUP)
            {
                float s[WIDE_GROUP][8];
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
#pragma unroll
                        for (int idx = 0; idx < F / 8; idx++)
                        {
                            const float4 v0 = tile[((ly + dy) * NCH + 2 * idx) * SW + lx + dx], v1 = tile[((ly + dy) * NCH + 2 * idx + 1) * SW + lx + dx];
                            const float r[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
                            const bool first = dy == 0 && dx == 0 && idx == 0;
#pragma unroll
                            for (int q = 0; q < WIDE_GROUP; q++)
                            {
                                const float* __restrict__ wk = prm.k + (cg + q) * 9 * F + (dy * 3 + dx) * F + idx * 8;
#pragma unroll
                                for (int c = 0; c < 8; c++) s[q][c] = fmaf(r[c], wk[c], first ? 0.0f : s[q][c]);
                            }
                        }
                float v[WIDE_GROUP];
#pragma unroll
                for (int q = 0; q < WIDE_GROUP; q++) v[q] = __fadd_rn(prm.b[cg + q], hsum8(s[q]));
                if (MODE == WIDE_SHUFFLE)
                {
                    // F -> 4, Identity, no nearest-neighbour residual; channel n lands at (2x + (n & 1), 2y + (n >> 1))
                    uint8_t* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy) * prm.dst_pitch;
                    net_store2(row, 2 * gx, prm.type, v[0], v[1], aligned);
                    net_store2(row + prm.dst_pitch, 2 * gx, prm.type, v[2], v[3], aligned);
                }
                else
                {
                    const size_t o = (static_cast<size_t>(gy) * prm.w + gx) * F + prm.co0 + cg;
#pragma unroll
                    for (int q = 0; q < WIDE_GROUP; q++)
                    {
                        if (prm.act == ACT_RELU) v[q] = fmaxf(v[q], 0.0f);
                        else if (prm.act == ACT_PRELU) v[q] = prelu(v[q], prm.a[cg + q]);
                    }
                    if (prm.res)
                    {
                        const float4 id = __ldg(reinterpret_cast<const float4*>(prm.res + o));
                        v[0] = fmaf(v[0], 1.0f, id.x); v[1] = fmaf(v[1], 1.0f, id.y); v[2] = fmaf(v[2], 1.0f, id.z); v[3] = fmaf(v[3], 1.0f, id.w);
                    }
                    *reinterpret_cast<float4*>(prm.out + o) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }

    // conv3x3_cin1 / conv5x5_cin1 with Identity activation: one thread per pixel, all F output channels
    struct WideHeadParams
    {
        const void* src;
        int src_pitch, type;
        float* out;             // [h][w][F]
        int w, h, F;
        float k[32 * 25];       // [F][ks*ks]
        float b[32];
    };
    template<int KS>
    __global__ void __launch_bounds__(256) wide_head_kernel(const __grid_constant__ WideHeadParams prm)
    {
        constexpr int CPOS = KS * KS, HALF = KS / 2, COUNT = CPOS / 8;
        const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
        if (x >= prm.w || y >= prm.h) return;
        float r[CPOS];
#pragma unroll
        for (int a = 0; a < KS; a++)
        {
            const uint8_t* row = static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(clampi(y + a - HALF, 0, prm.h - 1)) * prm.src_pitch;
#pragma unroll
            for (int b = 0; b < KS; b++) r[a * KS + b] = load_elem(row, clampi(x + b - HALF, 0, prm.w - 1), prm.type);
        }
        float* out = prm.out + (static_cast<size_t>(y) * prm.w + x) * prm.F;
#pragma unroll 1
        for (int n0 = 0; n0 < prm.F; n0 += 4)
        {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const float* __restrict__ k = prm.k + (n0 + q) * CPOS;
                // conv_cin1<cout, cpos>, X86/AVX.hpp:95-124: 8-tap vectors as per-lane FMA chains, hsum, remaining taps as scalars
                float t[8];
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    t[c] = __fmul_rn(r[c], k[c]);
#pragma unroll
                    for (int idx = 1; idx < COUNT; idx++) t[c] = fmaf(r[idx * 8 + c], k[idx * 8 + c], t[c]);
                }
                float s = hsum8(t);
#pragma unroll
                for (int p = COUNT * 8; p < CPOS; p++) s = fmaf(r[p], k[p], s);
                v[q] = __fadd_rn(s, prm.b[n0 + q]);
            }
            *reinterpret_cast<float4*>(out + n0) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }

    // FSRCNNX body end after its conv3x3 + PReLU: 1x1 + bias, + feat (scale 1.0), PReLU (conv<F,F,1>, Common.hpp:266-284)
    template<int F>
    struct WidePointParams
    {
        const float* in;        // conv3x3 + PReLU output
        const float* feat;
        float* out;
        int n_pixels;
        float k[F * F];
        float b[F];
        float a[F];
    };
    template<int F>
    __global__ void __launch_bounds__(256) wide_pointwise_kernel(const __grid_constant__ WidePointParams<F> prm)
    {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i >= prm.n_pixels) return;
        float r[F];
#pragma unroll
        for (int c = 0; c < F / 4; c++)
        {
            const float4 v = __ldg(reinterpret_cast<const float4*>(prm.in + static_cast<size_t>(i) * F) + c);
            r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
        }
#pragma unroll 1
        for (int n0 = 0; n0 < F; n0 += 4)
        {
            const float4 id = __ldg(reinterpret_cast<const float4*>(prm.feat + static_cast<size_t>(i) * F + n0));
            const float ft[4] = { id.x, id.y, id.z, id.w };
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const float* __restrict__ k = prm.k + (n0 + q) * F;
                float s[8];
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    s[c] = __fmul_rn(r[c], k[c]);
#pragma unroll
                    for (int idx = 1; idx < F / 8; idx++) s[c] = fmaf(r[idx * 8 + c], k[idx * 8 + c], s[c]);
                }
                const float sum = __fadd_rn(prm.b[n0 + q], hsum8(s));
                v[q] = prelu(fmaf(sum, 1.0f, ft[q]), prm.a[n0 + q]);
            }
            *reinterpret_cast<float4*>(prm.out + static_cast<size_t>(i) * F + n0) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}
