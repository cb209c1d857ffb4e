// fp32 CUDA-core (FFMA) engine of the fused luma network -- the EXACT engine.
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
// core/src/processor/cuda/CUDAProcessor.cpp:383-422, :451-490, :520-566.
//
// Arithmetic: bit-identical to the reference CPU processor's auto-ISA backend on FMA-capable x86
// (OpImplX86SIMD256<true>, core/internal/AC/Core/Internal/Processor/CPU/X86/AVX.hpp:24-146, which the AVX512
// backend also executes for these 8-channel layers, X86/AVX512.hpp:116).  Per output channel: eight FMA chains
// (one per input channel) over the nine taps in tap order, reduced by the hsum tree
// ((s0+s4)+(s2+s6))+((s1+s5)+(s3+s7)), added onto the bias; `v*scale + id` and `sat*max + 0.5f` are single FMAs
// exactly where g++ contracts them in that backend's translation unit.  fp32 activations everywhere
// (CPU/Common.hpp:116-393), clamp-to-edge borders at every layer (Common.hpp:121-141), weights `[cout][tap][cin]`.
//
// Design notes
//  * Weights ride in the kernel parameter block (__grid_constant__, constant bank 0).  All layers of a
//    segment are unrolled at compile time, so every weight sits at a compile-time constant-bank offset and
//    reaches the FFMA through a uniform register (ULDC) or an LDC -- no shared-memory traffic for weights.
//  * Feature maps are stored as two float4 planes (channels 0-3 / 4-7) of a fixed 56x56 frame, so a
//    warp reading 32 consecutive pixels issues conflict-free LDS.128.
//  * Each thread produces FFMA_P vertically adjacent pixels x 8 output channels with its (P+2) x 3 x 8 input
//    window held in registers, so every weight fetched feeds P FFMAs and every input is loaded once per layer.
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

    // i / d and i % d for the pixel loops over a tile region (0 <= i < 3300, 1 <= d <= 56): one multiply and a shift instead of the
    // ~35-instruction integer division sequence.  rcp = floor(2^20 / d) + 1 over-estimates 2^20 / d by less than 1, so the
    // quotient estimate is high by less than i / 2^20 < 0.0032 < 1 / 56 <= 1 - frac(i / d): the floor is exact.
    struct RegionDiv
    {
        uint32_t rcp;
        int d;
        __device__ __forceinline__ explicit RegionDiv(int d_) : rcp((1u << 20) / static_cast<uint32_t>(max(d_, 1)) + 1u), d(d_) {}
        __device__ __forceinline__ void split(int i, int& quo, int& rem) const
        {
            quo = static_cast<int>((static_cast<uint32_t>(i) * rcp) >> 20);
            rem = i - quo * d;
        }
    };

    struct TileGeom
    {
        int ox, oy;             // image coordinates of frame position (0,0)
        int ix0, ix1, iy0, iy1; // image bounds in frame coordinates (inclusive)
    };

#ifndef ACB_FFMA_P
#define ACB_FFMA_P 4
#endif
    constexpr int FFMA_P = ACB_FFMA_P;   // vertically adjacent pixels per thread

    // ((s0+s4)+(s2+s6))+((s1+s5)+(s3+s7)): OpImplX86SIMD256::hsum, X86/AVX.hpp:24-30
    __device__ __forceinline__ float hsum8(const float (&s)[8])
    {
        return __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[4]), __fadd_rn(s[2], s[6])), __fadd_rn(__fadd_rn(s[1], s[5]), __fadd_rn(s[3], s[7])));
    }
    __device__ __forceinline__ float prelu(float v, float alpha) { return fmaf(alpha, fminf(v, 0.0f), fmaxf(v, 0.0f)); }

    // fromFloat as the network tails round it in the FMA backend: one FMA for `sat*max + 0.5f`
    __device__ __forceinline__ void net_store1(void* row, int x, int type, float v)
    {
        switch (type)
        {
        case ACB200_UINT8: static_cast<uint8_t*>(row)[x] = static_cast<uint8_t>(fmaf(sat01(v), 255.0f, 0.5f)); return;
        case ACB200_UINT16: static_cast<uint16_t*>(row)[x] = static_cast<uint16_t>(fmaf(sat01(v), 65535.0f, 0.5f)); return;
        default: store_elem(row, x, type, v);
        }
    }
    __device__ __forceinline__ void net_store2(void* row, int x, int type, float v0, float v1, bool aligned)
    {
        switch (type)
        {
        case ACB200_UINT8:
        {
            const uint8_t q0 = static_cast<uint8_t>(fmaf(sat01(v0), 255.0f, 0.5f)), q1 = static_cast<uint8_t>(fmaf(sat01(v1), 255.0f, 0.5f));
            if (aligned) *reinterpret_cast<uchar2*>(static_cast<uint8_t*>(row) + x) = make_uchar2(q0, q1);
            else { static_cast<uint8_t*>(row)[x] = q0; static_cast<uint8_t*>(row)[x + 1] = q1; }
            return;
        }
        case ACB200_UINT16:
        {
            const uint16_t q0 = static_cast<uint16_t>(fmaf(sat01(v0), 65535.0f, 0.5f)), q1 = static_cast<uint16_t>(fmaf(sat01(v1), 65535.0f, 0.5f));
            if (aligned) *reinterpret_cast<ushort2*>(static_cast<uint16_t*>(row) + x) = make_ushort2(q0, q1);
            else { static_cast<uint16_t*>(row)[x] = q0; static_cast<uint16_t*>(row)[x + 1] = q1; }
            return;
        }
        default:
            store_elem2(row, x, type, v0, v1, aligned);
        }
    }

    // ---- 1->8 head conv from the luma tile: Common.hpp:166-197 ------------------------------------
    template<int ACT, int KOFF, int BOFF, int AOFF, class P>
    __device__ __forceinline__ void head_layer(const P& prm, const float* __restrict__ luma, float4* __restrict__ out, const TileGeom& g)
    {
        const int xa = max(0, g.ix0), xb = min(FT, g.ix1 + 1), ya = max(0, g.iy0), yb = min(FT, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * (yb - ya);
        const RegionDiv rd(ncols);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            int qy, qx;
            rd.split(i, qy, qx);
            const int x = xa + qx, y = ya + qy;
            float r[9];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) r[dy * 3 + dx] = luma[(y + dy) * LT + x + dx]; // luma frame origin = feature origin - 1
            float v[8];
#pragma unroll
            for (int co = 0; co < 8; co++)
            {
                // conv_cin1<8,9>, X86/AVX.hpp:95-124: taps 0-7 as products through the hsum tree, tap 8 as a scalar FMA, bias last
                float q[8];
#pragma unroll
                for (int p = 0; p < 8; p++) q[p] = __fmul_rn(r[p], prm.k[KOFF + co * 9 + p]);
                float s = __fadd_rn(fmaf(r[8], prm.k[KOFF + co * 9 + 8], hsum8(q)), prm.b[BOFF + co]);
                if (ACT == ACT_RELU) s = fmaxf(s, 0.0f);
                else if (ACT == ACT_PRELU) s = fmaf(prm.a[AOFF + co], fminf(s, 0.0f), fmaxf(s, 0.0f));
                v[co] = s;
            }
            out[y * FT + x] = make_float4(v[0], v[1], v[2], v[3]);
            out[FT * FT + y * FT + x] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }

    // FFMA_P vertically adjacent pixels x COUT channels of a 3x3 conv over an 8-ch map,
    // OpImplX86SIMD256<true>::conv<8,COUT,9>, X86/AVX.hpp:32-58,126-146.  The output-channel loop stays rolled
    // (runtime `co`) so ptxas cannot hoist a whole layer's worth of constant loads into the 63 uniform registers
    // and spill them.  `emit(co, v)` receives the FFMA_P finished sums (bias included) of output channel `co`.
#ifndef ACB_FFMA2
#define ACB_FFMA2 1
#endif
    // packed fp32 FMA of sm_100 (FFMA2): two independent IEEE fused multiply-adds per instruction, each rounded exactly like fmaf
    __device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c)
    {
        unsigned long long ra, rb, rc, rd;
        ra = *reinterpret_cast<const unsigned long long*>(&a); rb = *reinterpret_cast<const unsigned long long*>(&b); rc = *reinterpret_cast<const unsigned long long*>(&c);
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
        return *reinterpret_cast<float2*>(&rd);
    }

    // Vertically adjacent pixels per thread for the layer whose region is (FT - 2 L)^2 (interior tiles): a layer's pixel groups are dealt
    // over FFMA_THREADS threads in rounds, and the last round is usually partial -- 48 x 48 pixels are 576 groups of 4 rows (2.25 rounds
    // of 256: three rounds of 4 rows) but 768 groups of 3 rows (exactly three rounds of 3 rows).  The choice with the fewest rounds x rows;
    // ties go to the default (fewer operand loads per pixel).
    __host__ __device__ constexpr int ffma_rows_for(int L)
    {
        const int side = FT - 2 * L;
        int best = FFMA_P;
        long best_cost = 1L << 40;
        for (int p = 3; p <= FFMA_P; p++)
        {
            const long items = static_cast<long>(side) * ((side + p - 1) / p);
            const long rounds = (items + FFMA_THREADS - 1) / FFMA_THREADS;
            const long cost = rounds * p * 16 + (FFMA_P - p);
            if (cost < best_cost) { best_cost = cost; best = p; }
        }
        return best;
    }

    template<int COUT, int KOFF, int BOFF, int PP, class P, class Emit>
    __device__ __forceinline__ void conv_cols_rolled(const P& prm, const float4* __restrict__ in, const TileGeom& g, int x, int y, Emit&& emit)
    {
#if ACB_FFMA2
        // Packed form: input channels (2c, 2c + 1) of a pixel are one register pair, their weights one 64-bit constant load, their
        // accumulators one pair -- 36 FFMA2 instead of 72 FFMA per pixel and output channel.  Every accumulator lane still sees the same
        // products in the same order, so the sums are bit-identical to the scalar form (X86/AVX.hpp:32-58 lane for lane).
        float2 r2[PP + 2][3][4];
        const int cx2[3] = { clampi(x - 1, g.ix0, g.ix1), x, clampi(x + 1, g.ix0, g.ix1) };
#pragma unroll
        for (int iy = 0; iy < PP + 2; iy++)
        {
            const int ry = clampi(y - 1 + iy, g.iy0, min(g.iy1, FT - 1)) * FT;
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
            {
                const float4 v0 = in[ry + cx2[dx]], v1 = in[FT * FT + ry + cx2[dx]];
                r2[iy][dx][0] = make_float2(v0.x, v0.y); r2[iy][dx][1] = make_float2(v0.z, v0.w);
                r2[iy][dx][2] = make_float2(v1.x, v1.y); r2[iy][dx][3] = make_float2(v1.z, v1.w);
            }
        }
        // Output channels in groups of FOUR per (rolled) iteration: the caller stores a pixel's four channels as one float4 -- scalar
        // stores of one channel hit every fourth bank only (a four-way conflict on each of them, 40 % of the kernel's shared-memory wavefronts).
        static_assert(COUT % 4 == 0, "output channels are emitted in groups of four");
#pragma unroll 1
        for (int co4 = 0; co4 < COUT; co4 += 4)
        {
            float v4[4][PP];
            // (the four channels of a group run through ONE copy of the loop body: unrolled four times it is 12 % faster -- 0.760 ms for ACNetLegacy -- but the translation unit takes 14 minutes to compile, 7 minutes unrolled twice)
#pragma unroll 1
            for (int j = 0; j < 4; j++)
            {
                const int co = co4 + j;
                const float2* __restrict__ wk2 = reinterpret_cast<const float2*>(prm.k + KOFF + co * 72);
                float2 s2[PP][4];
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
#pragma unroll
                        for (int c2 = 0; c2 < 4; c2++)
                        {
                            const float2 wgt = wk2[(dy * 3 + dx) * 4 + c2];
#pragma unroll
                            for (int p = 0; p < PP; p++)
                                s2[p][c2] = fma2(r2[p + dy][dx][c2], wgt, (dy == 0 && dx == 0) ? make_float2(0.0f, 0.0f) : s2[p][c2]);
                        }
                const float bias = prm.b[BOFF + co];
#pragma unroll
                for (int p = 0; p < PP; p++)
                {
                    const float s8[8] = { s2[p][0].x, s2[p][0].y, s2[p][1].x, s2[p][1].y, s2[p][2].x, s2[p][2].y, s2[p][3].x, s2[p][3].y };
                    const float t = __fadd_rn(bias, hsum8(s8));
                    if (j == 0) v4[0][p] = t; else if (j == 1) v4[1][p] = t; else if (j == 2) v4[2][p] = t; else v4[3][p] = t;
                }
            }
            emit(co4, v4);
        }
        return;
#endif
        float r[PP + 2][3][8];
        const int cx[3] = { clampi(x - 1, g.ix0, g.ix1), x, clampi(x + 1, g.ix0, g.ix1) };
#pragma unroll
        for (int iy = 0; iy < PP + 2; iy++)
        {
            // (rows past the region's last group of PP feed sums that are dropped: keep their reads inside the frame)
            const int ry = clampi(y - 1 + iy, g.iy0, min(g.iy1, FT - 1)) * FT;
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
            {
                const float4 v0 = in[ry + cx[dx]], v1 = in[FT * FT + ry + cx[dx]];
                r[iy][dx][0] = v0.x; r[iy][dx][1] = v0.y; r[iy][dx][2] = v0.z; r[iy][dx][3] = v0.w;
                r[iy][dx][4] = v1.x; r[iy][dx][5] = v1.y; r[iy][dx][6] = v1.z; r[iy][dx][7] = v1.w;
            }
        }
#pragma unroll 1
        for (int co4 = 0; co4 < COUT; co4 += 4)
        {
            float v4[4][PP];
#pragma unroll 1
            for (int j = 0; j < 4; j++)
            {
                const int co = co4 + j;
                const float* __restrict__ wk = prm.k + KOFF + co * 72;
                float s[PP][8];
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
#pragma unroll
                        for (int ci = 0; ci < 8; ci++)
                        {
                            const float wgt = wk[(dy * 3 + dx) * 8 + ci];
#pragma unroll
                            for (int p = 0; p < PP; p++)
                                s[p][ci] = fmaf(r[p + dy][dx][ci], wgt, (dy == 0 && dx == 0) ? 0.0f : s[p][ci]);
                        }
                const float bias = prm.b[BOFF + co];
#pragma unroll
                for (int p = 0; p < PP; p++)
                {
                    const float t = __fadd_rn(bias, hsum8(s[p]));
                    if (j == 0) v4[0][p] = t; else if (j == 1) v4[1][p] = t; else if (j == 2) v4[2][p] = t; else v4[3][p] = t;
                }
            }
            emit(co4, v4);
        }
    }

    // ---- 8->8 body conv: Common.hpp:116-164 ---------------------------------------------------------
    // RES: `v*0.2 + out_old` in place (ARNet residual, CPUProcessor.cpp:1479,1483), one FMA as in the FMA backend.
    template<int L, int ACT, bool RES, int KOFF, int BOFF, int AOFF, int COUT = 8, class P>
    __device__ __forceinline__ void conv_layer(const P& prm, const float4* __restrict__ in, float4* __restrict__ out, const TileGeom& g)
    {
        constexpr int PP = ffma_rows_for(L);
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * ((yb - ya + PP - 1) / PP);
        const RegionDiv rdiv(ncols);
        float* __restrict__ outf = reinterpret_cast<float*>(out);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            int qy, qx;
            rdiv.split(i, qy, qx);
            const int x = xa + qx, y = ya + PP * qy;
            conv_cols_rolled<COUT, KOFF, BOFF, PP>(prm, in, g, x, y, [&](const int co4, const float (&v)[4][PP]) {
                // channels co4 .. co4 + 3 are the float4 of plane co4 / 4 at the pixel: one 16-byte store per pixel
                float4* dst = out + (co4 >> 2) * FT * FT + y * FT + x;
                float alpha[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                if (ACT == ACT_PRELU)
                {
#pragma unroll
                    for (int j = 0; j < 4; j++) alpha[j] = prm.a[AOFF + co4 + j];
                }
#pragma unroll
                for (int p = 0; p < PP; p++)
                {
                    if (y + p >= yb) break;
                    float s[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        s[j] = v[j][p];
                        if (ACT == ACT_RELU) s[j] = fmaxf(s[j], 0.0f);
                        else if (ACT == ACT_PRELU) s[j] = prelu(s[j], alpha[j]);
                    }
                    if (RES)
                    {
                        const float4 old = dst[p * FT];
                        s[0] = fmaf(s[0], 0.2f, old.x); s[1] = fmaf(s[1], 0.2f, old.y); s[2] = fmaf(s[2], 0.2f, old.z); s[3] = fmaf(s[3], 0.2f, old.w);
                    }
                    dst[p * FT] = make_float4(s[0], s[1], s[2], s[3]);
                }
            });
            (void)outf;
        }
    }

    // ---- ARNet end of body: conv3x3 *0.2 + x (done by conv_layer<IDENTITY, RES>), then this pointwise pass:
    //      1x1 + bias, PReLU, + feat: Common.hpp:223-288, conv<8,8,1> in X86/AVX.hpp order --------------------------
    template<int L, int KOFF1, int BOFF1, int AOFF, class P>
    __device__ __forceinline__ void arnet_fuse_pass(const P& prm, float4* __restrict__ buf, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * (yb - ya);
        const RegionDiv rdiv(ncols);
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            int qy, qx;
            rdiv.split(i, qy, qx);
            const int x = xa + qx, y = ya + qy, o = y * FT + x;
            const float4 t0 = buf[o], t1 = buf[FT * FT + o];
            const float t[8] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w };
            const float* f = prm.feat_in + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8;
            const float4 f0 = *reinterpret_cast<const float4*>(f), f1 = *reinterpret_cast<const float4*>(f + 4);
            const float ft[8] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w };
            float v[8];
#pragma unroll
            for (int co = 0; co < 8; co++)
            {
                float q[8];
#pragma unroll
                for (int ci = 0; ci < 8; ci++) q[ci] = __fmul_rn(t[ci], prm.k[KOFF1 + co * 8 + ci]);
                v[co] = __fadd_rn(prelu(__fadd_rn(prm.b[BOFF1 + co], hsum8(q)), prm.a[AOFF + co]), ft[co]);
            }
            buf[o] = make_float4(v[0], v[1], v[2], v[3]);
            buf[FT * FT + o] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }

    // ---- ACNetLegacy tail after its conv3x3 + ReLU: 2x2 stride-2 deconv 8->1, Common.hpp:344-393; dot<8> in
    //      X86/AVX.hpp:61-90 order (products through the hsum tree) --------------------------------------------------
    template<int L, int KOFF2, class P>
    __device__ __forceinline__ void deconv_pass(const P& prm, const float4* __restrict__ in, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * (yb - ya);
        const RegionDiv rdiv(ncols);
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            int qy, qx;
            rdiv.split(i, qy, qx);
            const int x = xa + qx, y = ya + qy, o = y * FT + x;
            const float4 t0 = in[o], t1 = in[FT * FT + o];
            const float t[8] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w };
            const int gx = g.ox + x, gy = g.oy + y;
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
            {
                float ov[2];
#pragma unroll
                for (int dx = 0; dx < 2; dx++)
                {
                    float q[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) q[c] = __fmul_rn(t[c], prm.k[KOFF2 + (dy * 2 + dx) * 8 + c]);
                    ov[dx] = hsum8(q);
                }
                void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy + dy) * prm.dst_pitch;
                net_store2(row, 2 * gx, prm.type, ov[0], ov[1], aligned);
            }
        }
    }

    // ---- ACNet / ARNet tail after its conv3x3 8->4: + nearest-upsampled input luma, pixel shuffle: Common.hpp:290-342
    template<int L, class P>
    __device__ __forceinline__ void shuffle_pass(const P& prm, const float4* __restrict__ in, const float* __restrict__ luma, const TileGeom& g)
    {
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int ncols = xb - xa, n = ncols * (yb - ya);
        const RegionDiv rdiv(ncols);
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
        {
            int qy, qx;
            rdiv.split(i, qy, qx);
            const int x = xa + qx, y = ya + qy;
            const float4 v = in[y * FT + x];
            const float id = luma[(y + 1) * LT + x + 1];
            const int gx = g.ox + x, gy = g.oy + y;
            uint8_t* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy) * prm.dst_pitch;
            net_store2(row, 2 * gx, prm.type, __fadd_rn(v.x, id), __fadd_rn(v.y, id), aligned);
            net_store2(row + prm.dst_pitch, 2 * gx, prm.type, __fadd_rn(v.z, id), __fadd_rn(v.w, id), aligned);
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
                const RegionDiv rdiv(ncols);
                for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
                {
                    int qy, qx;
                    rdiv.split(i, qy, qx);
                    const int x = xa + qx, y = ya + qy;
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
            const RegionDiv rdiv(ncols);
            for (int i = threadIdx.x; i < n; i += FFMA_THREADS)
            {
                int qy, qx;
                rdiv.split(i, qy, qx);
                const int x = xa + qx, y = ya + qy;
                float4* p = reinterpret_cast<float4*>(prm.map_out + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8);
                p[0] = cur[y * FT + x];
                p[1] = cur[FT * FT + y * FT + x];
            }
        }
        else if constexpr (S::FAM == ACB200_FAMILY_ACNET_LEGACY)
        {
            conv_layer<S::NCONV + 1, ACT_RELU, false, KT, BT, 0>(prm, cur, oth, g);
            __syncthreads();
            deconv_pass<S::NCONV + 1, KT + 576>(prm, oth, g);
        }
        else if constexpr (S::FAM == ACB200_FAMILY_ACNET)
        {
            conv_layer<S::NCONV + 1, ACT_IDENTITY, false, KT, BT, 0, 4>(prm, cur, oth, g);
            __syncthreads();
            shuffle_pass<S::NCONV + 1>(prm, oth, luma, g);
        }
        else
        {
            // ARNet: NCONV is even, so `cur` == bufA holds x
            constexpr int AT = (S::NCONV / 2) * 8;
            conv_layer<S::NCONV + 1, ACT_PRELU, false, KT, BT, AT>(prm, cur, oth, g);
            __syncthreads();
            conv_layer<S::NCONV + 2, ACT_IDENTITY, true, KT + 576, BT + 8, 0>(prm, oth, cur, g);
            __syncthreads();
            arnet_fuse_pass<S::NCONV + 2, KT + 576 + 576, BT + 8 + 8, AT + 8>(prm, cur, g);
            __syncthreads();
            conv_layer<S::NCONV + 3, ACT_IDENTITY, false, KT + 576 + 576 + 64, BT + 8 + 8 + 8, 0, 4>(prm, cur, oth, g);
            __syncthreads();
            shuffle_pass<S::NCONV + 3>(prm, oth, luma, g);
        }
    }

    constexpr size_t FFMA_SMEM_BYTES = 2 * 2 * FT * FT * sizeof(float4) + LT * LT * sizeof(float);
}
