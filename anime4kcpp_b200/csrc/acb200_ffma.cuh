// fp32 CUDA-core (FFMA) engine of the fused luma network.
//
// One CTA computes a T x T tile of the input-resolution luma through a *segment* of the network
// with every intermediate 8-channel feature map held in shared memory:
//
//   [luma tile 58x58 | 8-ch map 56x56 from HBM] -> conv3x3 ... conv3x3 -> [2x luma tile | 8-ch map to HBM]
//
// The whole ACNetLegacy / ACNet-B4 / ACNet-B8 network is ONE segment (one launch, luma in, 2x luma
// out); the deeper ACNet-B18 and ARNet-B8..B64 are chains of segments (halo R = #layers per segment).
//
// What it replaces in the reference: core/src/processor/cuda/Kernel.cu:246-514 (five per-layer kernel
// templates, one launch + one fp16 HBM round trip per layer) and the layer sequencing of
// core/src/processor/cuda/CUDAProcessor.cpp:383-422, :451-490, :520-566.  Arithmetic follows the CPU
// processor (the parity oracle): fp32 activations everywhere (core/internal/.../CPU/Common.hpp:116-393),
// clamp-to-edge borders at every layer (Common.hpp:121-141), weights `[cout][tap][cin]`
// (CPU/Generic.hpp:42-46).
//
// Design notes
//  * Weights ride in the kernel parameter block (__grid_constant__, constant bank 0).  All layers of a
//    segment are unrolled at compile time, so every weight is a compile-time constant-bank offset and
//    each MAC is a single `FFMA Rd, Ra, c[0x0][imm], Rd` -- no weight loads, no weight registers.
//  * Feature maps are stored as two float4 planes (channels 0-3 / 4-7) of a fixed 56x56 frame, so a
//    warp reading 32 consecutive pixels issues conflict-free LDS.128.
//  * Each thread produces 2 vertically adjacent pixels x 8 output channels (16 accumulators) and
//    re-uses each loaded input row for both output rows: 24 LDS.128 per 1152 FFMA.
//  * Layer l's output region is the frame shrunk by l; positions outside the image are never
//    computed -- reads clamp their coordinates to the image instead (replicate padding).
#pragma once

#include "acb200_common.cuh"

namespace acb
{
    constexpr int FT = 56;      // feature-map frame edge (pixels)
    constexpr int LT = 58;      // luma tile edge = FT + 2
#ifndef ACB_FFMA_THREADS
#define ACB_FFMA_THREADS 256
#endif
    constexpr int FFMA_THREADS = ACB_FFMA_THREADS;

    enum { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_PRELU = 2 };

    // Compile-time description of one segment.
    //   HEAD : the segment starts from the luma plane with the 1->8 conv (else it loads an 8-ch map)
    //   NCONV: number of plain 8->8 body convs in the segment
    //   TAIL : the segment ends with the family's tail and writes the 2x luma (else it stores the map)
    template<int FAM_, bool HEAD_, int NCONV_, bool TAIL_>
    struct Seg
    {
        static constexpr int FAM = FAM_;
        static constexpr bool HEAD = HEAD_;
        static constexpr int NCONV = NCONV_;
        static constexpr bool TAIL = TAIL_;
        // 3x3 layers applied after the first 8-ch map: body convs + tail convs
        // (ARNet tail = PReLU conv, residual conv fused with the 1x1, pixel-shuffle conv)
        static constexpr int TAIL_LAYERS = TAIL ? (FAM == ACB200_FAMILY_ARNET ? 3 : 1) : 0;
        static constexpr int R = NCONV + TAIL_LAYERS;
        static constexpr int T = FT - 2 * R;
        static constexpr int NK = (HEAD ? 72 : 0) + 576 * NCONV +
            (TAIL ? (FAM == ACB200_FAMILY_ACNET_LEGACY ? 576 + 32 : FAM == ACB200_FAMILY_ACNET ? 288 : 576 + 576 + 64 + 288) : 0);
        static constexpr int NB = (HEAD ? 8 : 0) + 8 * NCONV +
            (TAIL ? (FAM == ACB200_FAMILY_ACNET_LEGACY ? 8 : FAM == ACB200_FAMILY_ACNET ? 4 : 8 + 8 + 8 + 4) : 0);
        static constexpr int NA = FAM == ACB200_FAMILY_ACNET_LEGACY ? 0
            : FAM == ACB200_FAMILY_ACNET ? (HEAD ? 8 : 0) + 8 * NCONV
            : (NCONV / 2) * 8 + (TAIL ? 16 : 0);
        static constexpr bool NEEDS_LUMA = HEAD || (TAIL && FAM != ACB200_FAMILY_ACNET_LEGACY);
        static_assert(T >= 8, "segment too deep for the 56x56 frame");
        static_assert(FAM != ACB200_FAMILY_ARNET || (NCONV % 2) == 0, "ARNet segments start and end on block boundaries");
    };

    template<class S>
    struct SegParams
    {
        const void* src;        // luma plane of the pass input (element type `type`)
        const float* map_in;    // [h][w][8] fp32, when !HEAD
        float* map_out;         // [h][w][8] fp32, when !TAIL
        const float* feat_in;   // ARNet: head output, consumed by the tail segment
        float* feat_out;        // ARNet: written by the head segment
        void* dst;              // 2x luma plane, when TAIL
        int src_pitch, dst_pitch;   // bytes
        int w, h;                   // pass input size
        int type;                   // ACB200_* element type of src/dst
        int tiles_x;
        float k[S::NK];
        float b[S::NB];
        float a[S::NA > 0 ? S::NA : 1];
    };

    struct TileGeom
    {
        int ox, oy;             // image coordinates of frame position (0,0)
        int ix0, ix1, iy0, iy1; // image bounds in frame coordinates (inclusive)
    };

    // ---- 1->8 head conv from the luma tile: Common.hpp:166-197 ------------------------------------
    template<int ACT, int KOFF, int BOFF, int AOFF, class P>
    __device__ __forceinline__ void head_layer(const P& prm, const float* __restrict__ luma, float4* __restrict__ out, const TileGeom& g)
    {
        const int xa = max(0, g.ix0), xb = min(FT, g.ix1 + 1), ya = max(0, g.iy0), yb = min(FT, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * (yb - ya);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            const int x = xa + i % ncols, y = ya + i / ncols;
            float r[9];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) r[dy * 3 + dx] = luma[(y + dy) * LT + x + dx]; // luma frame origin = feature origin - 1
            float v[8];
#pragma unroll
            for (int co = 0; co < 8; co++)
            {
                float s = prm.b[BOFF + co];
#pragma unroll
                for (int p = 0; p < 9; p++) s = fmaf(r[p], prm.k[KOFF + co * 9 + p], s);
                if (ACT == ACT_RELU) s = fmaxf(s, 0.0f);
                else if (ACT == ACT_PRELU) s = fmaf(prm.a[AOFF + co], fminf(s, 0.0f), fmaxf(s, 0.0f));
                v[co] = s;
            }
            out[y * FT + x] = make_float4(v[0], v[1], v[2], v[3]);
            out[FT * FT + y * FT + x] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }

    // 2 vertically adjacent pixels x COUT channels of a 3x3 conv over an 8-ch map, accumulators only
    template<int COUT, int KOFF, int BOFF, class P>
    __device__ __forceinline__ void conv_pair(const P& prm, const float4* __restrict__ in, const TileGeom& g, int x, int y, float (&acc)[2][COUT])
    {
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int co = 0; co < COUT; co++) acc[p][co] = prm.b[BOFF + co];
        const int cx[3] = { clampi(x - 1, g.ix0, g.ix1), x, clampi(x + 1, g.ix0, g.ix1) };
#pragma unroll
        for (int iy = 0; iy < 4; iy++)
        {
            const int ry = clampi(y - 1 + iy, g.iy0, g.iy1) * FT;
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
            {
                const float4 v0 = in[ry + cx[dx]], v1 = in[FT * FT + ry + cx[dx]];
                const float a[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
#pragma unroll
                for (int p = 0; p < 2; p++)
                {
                    const int dy = iy - p;
                    if (dy < 0 || dy > 2) continue;
#pragma unroll
                    for (int co = 0; co < COUT; co++)
#pragma unroll
                        for (int ci = 0; ci < 8; ci++)
                            acc[p][co] = fmaf(a[ci], prm.k[KOFF + co * 72 + (dy * 3 + dx) * 8 + ci], acc[p][co]);
                }
            }
        }
    }

    // ---- 8->8 body conv: Common.hpp:116-164 ---------------------------------------------------------
    // RES: `v*0.2 + out_old` in place (ARNet residual, CPUProcessor.cpp:1479,1483), rounded as mul then add.
    template<int L, int ACT, bool RES, int KOFF, int BOFF, int AOFF, class P>
    __device__ __forceinline__ void conv_layer(const P& prm, const float4* __restrict__ in, float4* __restrict__ out, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * ((yb - ya + 1) >> 1);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            const int x = xa + i % ncols, y = ya + 2 * (i / ncols);
            float acc[2][8];
            conv_pair<8, KOFF, BOFF>(prm, in, g, x, y, acc);
#pragma unroll
            for (int p = 0; p < 2; p++)
            {
                if (y + p >= yb) break;
                const int o = (y + p) * FT + x;
                float v[8];
#pragma unroll
                for (int co = 0; co < 8; co++)
                {
                    float s = acc[p][co];
                    if (ACT == ACT_RELU) s = fmaxf(s, 0.0f);
                    else if (ACT == ACT_PRELU) s = fmaf(prm.a[AOFF + co], fminf(s, 0.0f), fmaxf(s, 0.0f));
                    v[co] = s;
                }
                if (RES)
                {
                    const float4 r0 = out[o], r1 = out[FT * FT + o];
                    const float id[8] = { r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w };
#pragma unroll
                    for (int co = 0; co < 8; co++) v[co] = __fadd_rn(__fmul_rn(v[co], 0.2f), id[co]);
                }
                out[o] = make_float4(v[0], v[1], v[2], v[3]);
                out[FT * FT + o] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }

    // ---- ARNet end of body: conv3x3 *0.2 + x, then 1x1 + bias, PReLU, + feat: Common.hpp:223-288 -----
    template<int L, int KOFF, int BOFF, int AOFF, class P>
    __device__ __forceinline__ void arnet_end_layer(const P& prm, const float4* __restrict__ in, float4* __restrict__ out, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * ((yb - ya + 1) >> 1);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            const int x = xa + i % ncols, y = ya + 2 * (i / ncols);
            float acc[2][8];
            conv_pair<8, KOFF, BOFF>(prm, in, g, x, y, acc);
#pragma unroll
            for (int p = 0; p < 2; p++)
            {
                if (y + p >= yb) break;
                const int o = (y + p) * FT + x;
                const float4 r0 = out[o], r1 = out[FT * FT + o];
                const float id[8] = { r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w };
                float t[8], v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) t[c] = __fadd_rn(__fmul_rn(acc[p][c], 0.2f), id[c]);
                const float* f = prm.feat_in + (static_cast<size_t>(g.oy + y + p) * prm.w + (g.ox + x)) * 8;
                const float4 f0 = *reinterpret_cast<const float4*>(f), f1 = *reinterpret_cast<const float4*>(f + 4);
                const float ft[8] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w };
#pragma unroll
                for (int co = 0; co < 8; co++)
                {
                    float s = prm.b[BOFF + 8 + co];
#pragma unroll
                    for (int ci = 0; ci < 8; ci++) s = fmaf(t[ci], prm.k[KOFF + 576 + co * 8 + ci], s);
                    s = fmaf(prm.a[AOFF + co], fminf(s, 0.0f), fmaxf(s, 0.0f));
                    v[co] = __fadd_rn(s, ft[co]);
                }
                out[o] = make_float4(v[0], v[1], v[2], v[3]);
                out[FT * FT + o] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }

    // ---- ACNetLegacy tail: conv3x3 8->8 + ReLU, 2x2 stride-2 deconv 8->1: Common.hpp:344-393 ---------
    template<int L, int KOFF, int BOFF, class P>
    __device__ __forceinline__ void tail_deconv_layer(const P& prm, const float4* __restrict__ in, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * ((yb - ya + 1) >> 1);
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            const int x = xa + i % ncols, y = ya + 2 * (i / ncols);
            float acc[2][8];
            conv_pair<8, KOFF, BOFF>(prm, in, g, x, y, acc);
#pragma unroll
            for (int p = 0; p < 2; p++)
            {
                if (y + p >= yb) break;
                float t[8];
#pragma unroll
                for (int c = 0; c < 8; c++) t[c] = fmaxf(acc[p][c], 0.0f);
                const int gx = g.ox + x, gy = g.oy + y + p;
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
                {
                    float o[2];
#pragma unroll
                    for (int dx = 0; dx < 2; dx++)
                    {
                        float s = 0.0f;
#pragma unroll
                        for (int c = 0; c < 8; c++) s = fmaf(t[c], prm.k[KOFF + 576 + (dy * 2 + dx) * 8 + c], s);
                        o[dx] = s;
                    }
                    void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy + dy) * prm.dst_pitch;
                    store_elem2(row, 2 * gx, prm.type, o[0], o[1], aligned);
                }
            }
        }
    }

    // ---- ACNet / ARNet tail: conv3x3 8->4, + nearest-upsampled input luma, pixel shuffle: Common.hpp:290-342
    template<int L, int KOFF, int BOFF, class P>
    __device__ __forceinline__ void tail_shuffle_layer(const P& prm, const float4* __restrict__ in, const float* __restrict__ luma, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * ((yb - ya + 1) >> 1);
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            const int x = xa + i % ncols, y = ya + 2 * (i / ncols);
            float acc[2][4];
            conv_pair<4, KOFF, BOFF>(prm, in, g, x, y, acc);
#pragma unroll
            for (int p = 0; p < 2; p++)
            {
                if (y + p >= yb) break;
                const float id = luma[(y + p + 1) * LT + x + 1];
                const int gx = g.ox + x, gy = g.oy + y + p;
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
                {
                    void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy + dy) * prm.dst_pitch;
                    store_elem2(row, 2 * gx, prm.type, __fadd_rn(acc[p][dy * 2], id), __fadd_rn(acc[p][dy * 2 + 1], id), aligned);
                }
            }
        }
    }

    // compile-time walk over the body convs of a segment
    template<class S, int I, class P>
    __device__ __forceinline__ void run_body(const P& prm, float4* bufA, float4* bufB, const TileGeom& g)
    {
        if constexpr (I < S::NCONV)
        {
            constexpr int KOFF = (S::HEAD ? 72 : 0) + 576 * I;
            constexpr int BOFF = (S::HEAD ? 8 : 0) + 8 * I;
            float4* in = (I & 1) ? bufB : bufA;
            float4* out = (I & 1) ? bufA : bufB;
            if constexpr (S::FAM == ACB200_FAMILY_ACNET_LEGACY)
                conv_layer<I + 1, ACT_RELU, false, KOFF, BOFF, 0>(prm, in, out, g);
            else if constexpr (S::FAM == ACB200_FAMILY_ACNET)
                conv_layer<I + 1, ACT_PRELU, false, KOFF, BOFF, (S::HEAD ? 8 : 0) + 8 * I>(prm, in, out, g);
            else if constexpr ((I & 1) == 0)
                conv_layer<I + 1, ACT_PRELU, false, KOFF, BOFF, (I / 2) * 8>(prm, in, out, g);
            else
                conv_layer<I + 1, ACT_IDENTITY, true, KOFF, BOFF, 0>(prm, in, out, g);
            __syncthreads();
            run_body<S, I + 1>(prm, bufA, bufB, g);
        }
    }

    template<class S>
    __global__ void __launch_bounds__(FFMA_THREADS, 1) segment_ffma_kernel(const __grid_constant__ SegParams<S> prm)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        float4* bufA = reinterpret_cast<float4*>(smem_raw);              // 2 planes x 56x56 float4
        float4* bufB = bufA + 2 * FT * FT;
        float* luma = reinterpret_cast<float*>(bufB + 2 * FT * FT);      // 58x58 float

        const int tx = blockIdx.x % prm.tiles_x, ty = blockIdx.x / prm.tiles_x;
        TileGeom g;
        g.ox = tx * S::T - S::R;
        g.oy = ty * S::T - S::R;
        g.ix0 = -g.ox; g.ix1 = prm.w - 1 - g.ox;
        g.iy0 = -g.oy; g.iy1 = prm.h - 1 - g.oy;

        if constexpr (S::NEEDS_LUMA)
        {
            for (int i = threadIdx.x; i < LT * LT; i += FFMA_THREADS)
            {
                const int lx = i % LT, ly = i / LT;
                const int gx = clampi(g.ox - 1 + lx, 0, prm.w - 1), gy = clampi(g.oy - 1 + ly, 0, prm.h - 1);
                luma[i] = load_elem(static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch, gx, prm.type);
            }
        }
        if constexpr (!S::HEAD)
        {
            for (int i = threadIdx.x; i < FT * FT; i += FFMA_THREADS)
            {
                const int fx = i % FT, fy = i / FT;
                const int gx = clampi(g.ox + fx, 0, prm.w - 1), gy = clampi(g.oy + fy, 0, prm.h - 1);
                const float4* p = reinterpret_cast<const float4*>(prm.map_in + (static_cast<size_t>(gy) * prm.w + gx) * 8);
                bufA[i] = __ldg(p);
                bufA[FT * FT + i] = __ldg(p + 1);
            }
        }
        __syncthreads();
        if constexpr (S::HEAD)
        {
            constexpr int ACT = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? ACT_RELU : S::FAM == ACB200_FAMILY_ACNET ? ACT_PRELU : ACT_IDENTITY;
            head_layer<ACT, 0, 0, 0>(prm, luma, bufA, g);
            __syncthreads();
            if constexpr (S::FAM == ACB200_FAMILY_ARNET)
            {
                // keep the head output (`feat`) for the tail segment: centre T x T of this tile
                const int xa = max(S::R, g.ix0), xb = min(FT - S::R, g.ix1 + 1), ya = max(S::R, g.iy0), yb = min(FT - S::R, g.iy1 + 1);
                const int ncols = xb - xa, n = ncols * (yb - ya);
                for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
                {
                    const int x = xa + i % ncols, y = ya + i / ncols;
                    float4* p = reinterpret_cast<float4*>(prm.feat_out + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8);
                    p[0] = bufA[y * FT + x];
                    p[1] = bufA[FT * FT + y * FT + x];
                }
            }
        }
        run_body<S, 0>(prm, bufA, bufB, g);
        float4* cur = (S::NCONV & 1) ? bufB : bufA;
        float4* oth = (S::NCONV & 1) ? bufA : bufB;
        constexpr int KT = (S::HEAD ? 72 : 0) + 576 * S::NCONV;
        constexpr int BT = (S::HEAD ? 8 : 0) + 8 * S::NCONV;
        if constexpr (!S::TAIL)
        {
            const int xa = max(S::R, g.ix0), xb = min(FT - S::R, g.ix1 + 1), ya = max(S::R, g.iy0), yb = min(FT - S::R, g.iy1 + 1);
            const int ncols = xb - xa, n = ncols * (yb - ya);
            for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
            {
                const int x = xa + i % ncols, y = ya + i / ncols;
                float4* p = reinterpret_cast<float4*>(prm.map_out + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8);
                p[0] = cur[y * FT + x];
                p[1] = cur[FT * FT + y * FT + x];
            }
        }
        else if constexpr (S::FAM == ACB200_FAMILY_ACNET_LEGACY)
            tail_deconv_layer<S::NCONV + 1, KT, BT>(prm, cur, g);
        else if constexpr (S::FAM == ACB200_FAMILY_ACNET)
            tail_shuffle_layer<S::NCONV + 1, KT, BT>(prm, cur, luma, g);
        else
        {
            // ARNet: NCONV is even, so `cur` == bufA holds x
            constexpr int AT = (S::NCONV / 2) * 8;
            conv_layer<S::NCONV + 1, ACT_PRELU, false, KT, BT, AT>(prm, cur, oth, g);
            __syncthreads();
            arnet_end_layer<S::NCONV + 2, KT + 576, BT + 8, AT + 8>(prm, oth, cur, g);
            __syncthreads();
            tail_shuffle_layer<S::NCONV + 3, KT + 576 + 576 + 64, BT + 8 + 8 + 8>(prm, cur, luma, g);
        }
    }

    constexpr size_t FFMA_SMEM_BYTES = 2 * 2 * FT * FT * sizeof(float4) + LT * LT * sizeof(float);
}
