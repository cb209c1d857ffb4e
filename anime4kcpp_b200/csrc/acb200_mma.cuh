// Tensor-core engine of the fused luma network: split-fp16 implicit GEMM on mma.sync (HMMA).
//
// Same tiling, segment chain and reference semantics as the exact FFMA engine (acb200_ffma.cuh), but the 8->8
// (and 8->4) 3x3 convolutions run on the tensor cores as implicit GEMMs  D[16 px x 8 cout] += A[16 px x K] B[K x 8],
// K = 9 taps x 8 cin = 72, five k-steps (four m16n8k16 covering two taps each + one m16n8k8).
//
// Plain fp16/bf16/TF32 operands cannot meet the ">= 99.9 % of 8-bit samples bit-exact" bar (SURVEY.md finding 4:
// fp16 activations leave 4-9 % of samples 1 LSB off), so every fp32 activation a and weight w is carried as an
// fp16 pair (hi, lo) with hi = fp16(v), lo = fp16(v - hi) (22+ significant bits), and each product is evaluated as
//     a*w  ~=  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo          (three MMAs, fp32 accumulation in the tensor core)
// which leaves rounding noise of the same order as re-ordering an fp32 sum.
//
// Shared memory holds each 8-channel map as two planes of [56x56 pixels][8 x fp16] = 16 bytes per pixel (hi plane and
// lo plane).  An 8x8 ldmatrix tile is then 8 consecutive pixels x 8 channels = 128 contiguous bytes (conflict-free), a
// 3x3 tap is a pixel offset of the row addresses, and one ldmatrix.x4 yields the whole A fragment of a k-step
// (16 pixels x two taps).  The D fragment of thread (g = lane/4, t = lane%4) is pixels {g, g+8} x couts {2t, 2t+1}:
// bias, activation, residual and the hi/lo re-split happen in registers, and the 4-byte half2 stores of a warp cover
// 8 pixels x 16 bytes contiguously.  Weights are pre-packed on the host into B-fragment order (one coalesced 128-byte
// load per fragment register per warp) and stay in 18 registers for a whole layer.
//
// The 1->8 head conv (72 MAC / pixel) and the deconvolution / pixel-shuffle / 1x1 epilogues stay in fp32 FFMA.
#pragma once

#include <cuda.h>        // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_fp16.h>

#include <type_traits>

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"

namespace acb
{
#ifndef ACB_MMA_THREADS
#define ACB_MMA_THREADS 512
#endif
    constexpr int MMA_THREADS = ACB_MMA_THREADS;
    constexpr int MMA_WARPS = MMA_THREADS / 32;
    constexpr int FRAG_WORDS_3X3 = 18 * 32;     // uint32 per packed 3x3 layer: (4 x 2 + 1) registers x {hi, lo} x 32 lanes
    constexpr int FRAG_WORDS_1X1 = 2 * 32;      // the ARNet 1x1: one k8 register x {hi, lo}
#ifndef ACB_MMA_CHAINS
#define ACB_MMA_CHAINS 3
#endif
    constexpr int MMA_CHAINS = ACB_MMA_CHAINS;
    constexpr int PLANE_BYTES = FT * FT * 16;   // distance between the hi and the lo plane of a map
    constexpr size_t MMA_SMEM_DATA_BYTES = 4 * FT * FT * sizeof(uint4) + LT * LT * sizeof(float);

    // Inter-segment maps of this engine are stored ALREADY SPLIT: two planes (hi, lo) of [h][w][8 x fp16] = 16 bytes per pixel and
    // plane -- the shared-memory layout of the frame.  The next segment's 56 x 56 frame (both planes, 100 352 bytes) is then ONE TMA
    // tensor copy: a 3-D box {56 x 16 bytes, 56 rows, 2 planes} at (ox, oy, 0) lands as [plane][y][x][8] = {A.hi, A.lo}; coordinates outside
    // the image are zero-filled by the TMA unit, and border CTAs never read them (replicate padding clamps READ coordinates).
    constexpr uint32_t MAP_BOX_BYTES = 2u * FT * FT * 16u;
    template<class S>
    struct MmaParams
    {
        alignas(64) CUtensorMap tmap;   // the previous segment's map (!HEAD segments), see above
        const void* src;
        const float* map_in;
        float* map_out;
        const float* feat_in;
        float* feat_out;
        void* dst;
        int src_pitch, dst_pitch;
        int w, h;
        int type;
        int tiles_x;
        const uint32_t* frags;  // packed B fragments of this segment's 3x3 convs (and the ARNet 1x1), in layer order
        float k[(S::HEAD ? 72 : 0) + 64 + 32];  // fp32 weights used outside the MMAs: head (72) | legacy deconv (32)
        float b[S::NB];
        float a[S::NA > 0 ? S::NA : 1];
    };

    struct HalfPlanes
    {
        uint4* hi;  // [FT*FT] pixels x 8 fp16
        uint4* lo;
    };

    __device__ __forceinline__ uint32_t pack_half2(__half a, __half b)
    {
        return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
    }
    // v -> (hi, lo) fp16 pair
    __device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo)
    {
        const __half2 h = __floats2half2_rn(v0, v1);        // one packed conversion
        hi = *reinterpret_cast<const uint32_t*>(&h);
        // v - float(hi) in ONE instruction per value: the sm_100 mixed-precision FMA (FHFMA) takes the fp16 halves of `hi`
        // directly, hi * (-1) + v with a single rounding (the difference is exactly representable, so it is exact)
        const unsigned short h0 = static_cast<unsigned short>(hi & 0xffffu), h1 = static_cast<unsigned short>(hi >> 16), m1 = 0xBC00;
        float l0, l1;
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"(h0), "h"(m1), "f"(v0));
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"(h1), "h"(m1), "f"(v1));
        const __half2 l = __floats2half2_rn(l0, l1);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    }
    __device__ __forceinline__ float2 join_pair(uint32_t hi, uint32_t lo)
    {
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi)), l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
        return make_float2(h.x + l.x, h.y + l.y);
    }
    __device__ __forceinline__ void store_pixel_split(const HalfPlanes& p, int o, const float (&v)[8])
    {
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x); split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z); split_pair(v[6], v[7], hi.w, lo.w);
        p.hi[o] = hi; p.lo[o] = lo;
    }
    __device__ __forceinline__ void load_pixel_joined(const HalfPlanes& p, int o, float (&v)[8])
    {
        const uint4 hi = p.hi[o], lo = p.lo[o];
        float2 f;
        f = join_pair(hi.x, lo.x); v[0] = f.x; v[1] = f.y;
        f = join_pair(hi.y, lo.y); v[2] = f.x; v[3] = f.y;
        f = join_pair(hi.z, lo.z); v[4] = f.x; v[5] = f.y;
        f = join_pair(hi.w, lo.w); v[6] = f.x; v[7] = f.y;
    }

    __device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr)
    {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
    }
    __device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t saddr)
    {
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
    }
    __device__ __forceinline__ void mma_k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
    {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    __device__ __forceinline__ void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0)
    {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
    }

    __device__ __forceinline__ void ldmatrix_x4_off(uint32_t (&r)[4], uint32_t saddr, int imm_plane)
    {
        // second plane (lo) sits at a fixed byte distance from the first: fold it into the address immediate
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr), "n"(FT * FT * 16));
        (void)imm_plane;
    }
    __device__ __forceinline__ void ldmatrix_x2_off(uint32_t (&r)[2], uint32_t saddr)
    {
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2+%3];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr), "n"(FT * FT * 16));
    }

    // One 3x3 conv layer (8 -> 8, or 8 -> 4 with the upper output channels' weights zero) on the tensor cores.
    //
    // The layer's output region (image-clipped, Wr x Hr pixels) is tiled FLAT: M-tile j is the 16 row-major-consecutive
    // region pixels 16j .. 16j+15, wrapping across region rows.  ldmatrix takes one row address per lane, so a wrap costs
    // nothing but an occasional 2-way bank conflict, and no MMA row is spent on padding a region width up to a multiple of 16.
    //
    //   epi(off, v0, v1, valid): called by every lane twice per tile -- for the D-fragment rows g and g+8 -- with the
    //   finished fp32 sums (bias0 / bias1 included: they are the initial accumulator) of output channels 2t, 2t+1 at frame pixel off = py * FT + px.  `valid` is
    //   false only in the overhang of the last tile (where `off` aliases the last valid pixel): the epilogue may use warp
    //   collectives and may read, but must not store for those.
    //   L = number of 3x3 layers between the frame and this layer's output (its region is the frame shrunk by L).
    //   The hi plane and the lo plane of `in` must be FT*FT*16 bytes apart (they are: see the kernel's smem carve-up).
    // B fragments of one 3x3 layer, one register per tap: 0-8 = w_hi, 9-17 = w_lo.  The kernel loads them one layer AHEAD (while
    // the previous layer's MMAs run), so no layer starts by waiting ~700 cycles for 18 L2 loads.
    __device__ __forceinline__ void mma_load_bfrag(uint32_t (&bf)[18], const uint32_t* __restrict__ frag)
    {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < 18; i++) bf[i] = __ldg(frag + i * 32 + lane);
    }

    template<bool BORDER, int NT, class Epi>
    __device__ __forceinline__ void mma_conv3x3_impl(const int L, const HalfPlanes& in, const uint32_t (&bf)[18], const TileGeom& g, const float bias0, const float bias1, Epi&& epi)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        const int wr = xb - xa, npix = wr * (yb - ya), tiles = (npix + 15) >> 4;
        const uint32_t rcp = (1u << 20) / static_cast<uint32_t>(max(wr, 1)) + 1u;     // q / wr == (q * rcp) >> 20 for q < 3300, wr <= 56
        const int cx_lo = max(g.ix0, 0), cx_hi = min(g.ix1, FT - 1), cy_lo = max(g.iy0, 0), cy_hi = min(g.iy1, FT - 1);
        const int m = lane >> 3, r = lane & 7, h = m >> 1;      // ldmatrix: this lane supplies row r of matrix m
        const uint32_t hi_base = static_cast<uint32_t>(__cvta_generic_to_shared(in.hi));
        // byte offset of this lane's tap for k-step s relative to its pixel (interior tiles: no clamping needed, every
        // read stays inside the frame because the region is the frame shrunk by L >= 1)
        int toff[5];
#pragma unroll
        for (int s = 0; s < 5; s++)
        {
            const int tap = (h && s < 4) ? 2 * s + 1 : 2 * s;
            toff[s] = ((tap / 3 - 1) * FT + (tap % 3 - 1)) * 16;
        }
        const int arow = r + ((m & 1) << 3), drow = lane >> 2;
        // NT M-tiles of 16 pixels are in flight per warp (their ldmatrix / MMA sequences interleave).  NT = 1 is what ships:
        // measured with NT = 2 (-DACB_MMA_TILES=2) ACNet-B8 gains 3 % but ACNetLegacy and ARNet lose 3-4 % (coarser work units
        // balance worse over the 16 warps than the extra instruction-level parallelism buys).
        // Interior CTAs with one tile in flight walk their pixel INCREMENTALLY: a warp's next tile lies 16 * MMA_WARPS region
        // pixels ahead, i.e. `dq` region rows and `rq` columns further (one conditional wrap), so the frame offset moves by a
        // constant plus a conditional constant -- five instructions instead of the multiply / shift / multiply-subtract / multiply-
        // add of the direct mapping.  Lanes past the region's last pixel (last tile only) are not clamped: they read at most one
        // row and one pixel below the region, which is inside the kernel's shared memory, and their results are dropped.
        constexpr bool WALK = !BORDER && NT == 1;
        int wqx = 0, woff = 0, step_off = 0, rq = 0;
        if (WALK)
        {
            const int q = warp * 16 + arow;
            const int qy = static_cast<int>((static_cast<uint32_t>(q) * rcp) >> 20);
            wqx = q - qy * wr;
            woff = (ya + qy) * FT + xa + wqx;
            const int dq = static_cast<int>((static_cast<uint32_t>(16 * MMA_WARPS) * rcp) >> 20);
            rq = 16 * MMA_WARPS - dq * wr;
            step_off = dq * FT + rq;
        }
        for (int it0 = warp * NT; it0 < tiles; it0 += MMA_WARPS * NT)
        {
            int px[NT], py[NT], poff[NT];
            uint32_t addr[NT][5];
            float c0[NT][4], c1[NT][4], c2[NT][4];
#pragma unroll
            for (int t = 0; t < NT; t++)
            {
                if (WALK)
                {
                    poff[t] = woff;
                    px[t] = py[t] = 0;      // (only the clamped border path needs them)
                    wqx += rq; woff += step_off;
                    if (wqx >= wr) { wqx -= wr; woff += FT - wr; }
                }
                else
                {
                    const int q = min((it0 + t) * 16 + arow, npix - 1);
                    const int qy = static_cast<int>((static_cast<uint32_t>(q) * rcp) >> 20), qx = q - qy * wr;
                    px[t] = xa + qx; py[t] = ya + qy;
                    poff[t] = py[t] * FT + px[t];
                }
#pragma unroll
                for (int e = 0; e < 4; e++) c1[t][e] = c2[t][e] = 0.0f;
                c0[t][0] = c0[t][2] = bias0; c0[t][1] = c0[t][3] = bias1;      // the bias rides in as the first chain's C operand
                if (BORDER)
                {
                    const int cx[3] = { clampi(px[t] - 1, cx_lo, cx_hi), px[t], clampi(px[t] + 1, cx_lo, cx_hi) };
                    const int ry[3] = { clampi(py[t] - 1, cy_lo, cy_hi) * FT, py[t] * FT, clampi(py[t] + 1, cy_lo, cy_hi) * FT };
#pragma unroll
                    for (int s = 0; s < 5; s++)
                    {
                        const int t0 = 2 * s, t1 = 2 * s + 1;
                        addr[t][s] = hi_base + (((h && s < 4) ? ry[t1 / 3] + cx[t1 % 3] : ry[t0 / 3] + cx[t0 % 3]) << 4);
                    }
                }
                else
                {
                    const uint32_t pix = hi_base + (poff[t] << 4);
#pragma unroll
                    for (int s = 0; s < 5; s++) addr[t][s] = pix + toff[s];
                }
            }
            // MMA_CHAINS independent accumulator chains per tile (3: a_hi*w_hi, a_lo*w_hi, a_hi*w_lo; 2: the a_lo*w_hi products
            // alternate between the other two chains, seven MMAs each, which saves the four adds that join c1).  An m16n8k8 occupies the tensor pipe
            // exactly as long as an m16n8k16 (8 cycles per SM partition, tools/microbench_hmma_latency.cu), so the odd ninth tap is
            // loaded as ONE fragment {hi | lo} and multiplied by [w_hi ; w_hi] in a single k16: 14 MMA slots per tile instead of 15.
#pragma unroll
            for (int s = 0; s < 4; s++)
            {
                uint32_t ah[NT][4], al[NT][4];
#pragma unroll
                for (int t = 0; t < NT; t++)
                {
                    ldmatrix_x4(ah[t], addr[t][s]);
                    ldmatrix_x4_off(al[t], addr[t][s], 0);
                }
#pragma unroll
                for (int t = 0; t < NT; t++)
                {
                    mma_k16(c0[t], ah[t], bf[2 * s], bf[2 * s + 1]);
                    mma_k16(c2[t], ah[t], bf[9 + 2 * s], bf[9 + 2 * s + 1]);
                    if (MMA_CHAINS == 3) mma_k16(c1[t], al[t], bf[2 * s], bf[2 * s + 1]);
                    else if (s & 1) mma_k16(c2[t], al[t], bf[2 * s], bf[2 * s + 1]);
                    else mma_k16(c0[t], al[t], bf[2 * s], bf[2 * s + 1]);
                }
            }
            {
                uint32_t f8[NT][4];
#pragma unroll
                for (int t = 0; t < NT; t++)
                    ldmatrix_x4(f8[t], addr[t][4] + (h ? static_cast<uint32_t>(FT * FT * 16) : 0u));   // matrices 0,1: hi plane; 2,3: lo plane
#pragma unroll
                for (int t = 0; t < NT; t++)
                {
                    mma_k16(MMA_CHAINS == 3 ? c1[t] : c0[t], f8[t], bf[8], bf[8]);
                    mma_k8(c2[t], f8[t][0], f8[t][1], bf[17]);
                }
            }
#pragma unroll
            for (int t = 0; t < NT; t++)
            {
                if (MMA_CHAINS == 3)
                {
#pragma unroll
                    for (int e = 0; e < 4; e++) c0[t][e] += c1[t][e];
                }
                // D fragment rows are region pixels 16*it + g and + 8: exactly the pixels lanes g and g + 8 addressed above
                const int d0 = __shfl_sync(0xffffffffu, poff[t], drow), d1 = __shfl_sync(0xffffffffu, poff[t], drow + 8);
                const int q0 = (it0 + t) * 16 + drow;
                epi(d0, c0[t][0] + c2[t][0], c0[t][1] + c2[t][1], q0 < npix);
                epi(d1, c0[t][2] + c2[t][2], c0[t][3] + c2[t][3], q0 + 8 < npix);
            }
        }
    }
    template<class Epi>
    __device__ __forceinline__ void mma_conv3x3(const int L, const HalfPlanes& in, const uint32_t (&bf)[18], const TileGeom& g, const float bias0, const float bias1, Epi&& epi)
    {
        // interior CTA: the image covers the whole frame, no replicate padding anywhere in this tile (uniform branch)
#ifndef ACB_MMA_TILES
#define ACB_MMA_TILES 1
#endif
        if (g.ix0 <= 0 && g.iy0 <= 0 && g.ix1 >= FT - 1 && g.iy1 >= FT - 1) mma_conv3x3_impl<false, ACB_MMA_TILES>(L, in, bf, g, bias0, bias1, epi);
        else mma_conv3x3_impl<true, 1>(L, in, bf, g, bias0, bias1, epi);
    }

    template<class S>
    __global__ void __launch_bounds__(MMA_THREADS, 1) segment_mma_kernel(const __grid_constant__ MmaParams<S> prm)
    {
        extern __shared__ __align__(128) unsigned char smem_mma[];
        uint4* base = reinterpret_cast<uint4*>(smem_mma);
        HalfPlanes A{ base, base + FT * FT }, B{ base + 2 * FT * FT, base + 3 * FT * FT };
        float* luma = reinterpret_cast<float*>(base + 4 * FT * FT);
        const uint32_t tma_bar = static_cast<uint32_t>(__cvta_generic_to_shared(smem_mma + MMA_SMEM_DATA_BYTES));   // mbarrier of the map copy

        const int tx = blockIdx.x % prm.tiles_x, ty = blockIdx.x / prm.tiles_x;
        TileGeom g;
        g.ox = tx * S::T - S::R;
        g.oy = ty * S::T - S::R;
        g.ix0 = -g.ox; g.ix1 = prm.w - 1 - g.ox;
        g.iy0 = -g.oy; g.iy1 = prm.h - 1 - g.oy;
        const int lane = threadIdx.x & 31, tq = lane & 3;
        if constexpr (!S::HEAD)
        {
            // one thread starts the tensor copy of the whole frame; everybody waits for it after the barrier below
            if (threadIdx.x == 0)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tma_bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tma_bar), "r"(MAP_BOX_BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(A.hi))), "l"(reinterpret_cast<uint64_t>(&prm.tmap)),
                                "r"(4 * g.ox), "r"(g.oy), "r"(0), "r"(tma_bar) : "memory");
            }
        }
        uint32_t bf[18], nb[18];        // B fragments of the current / the next 3x3 layer
        mma_load_bfrag(bf, prm.frags);  // consumed after the input load (and the head): the latency is hidden

        if constexpr (S::NEEDS_LUMA)
        {
            if (prm.type == ACB200_UINT8)
            {
                // all of this thread's pixel loads are issued before the first one is consumed (one exposed HBM/L2 latency
                // instead of one per loop iteration)
                constexpr int PER = (LT * LT + MMA_THREADS - 1) / MMA_THREADS;
                uint8_t px[PER];
#pragma unroll
                for (int k = 0; k < PER; k++)
                {
                    const int i = min(threadIdx.x + k * MMA_THREADS, LT * LT - 1);
                    const int lx = i % LT, ly = i / LT;
                    const int gx = clampi(g.ox - 1 + lx, 0, prm.w - 1), gy = clampi(g.oy - 1 + ly, 0, prm.h - 1);
                    px[k] = __ldg(static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch + gx);
                }
#pragma unroll
                for (int k = 0; k < PER; k++)
                {
                    const int i = threadIdx.x + k * MMA_THREADS;
                    if (i < LT * LT) luma[i] = unit_from_int<255>(static_cast<float>(px[k]));     // toFloat<u8>, exact, no division
                }
            }
            else
                for (int i = threadIdx.x; i < LT * LT; i += MMA_THREADS)
                {
                    const int lx = i % LT, ly = i / LT;
                    const int gx = clampi(g.ox - 1 + lx, 0, prm.w - 1), gy = clampi(g.oy - 1 + ly, 0, prm.h - 1);
                    luma[i] = load_elem(static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch, gx, prm.type);
                }
        }
        __syncthreads();
        if constexpr (!S::HEAD)
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(tma_bar), "r"(0) : "memory");
        }
        if constexpr (S::HEAD)
        {
            // 1 -> 8 head conv in fp32 (Common.hpp:166-197), one pixel per thread
            constexpr int ACT = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? ACT_RELU : S::FAM == ACB200_FAMILY_ACNET ? ACT_PRELU : ACT_IDENTITY;
            const int xa = max(0, g.ix0), xb = min(FT, g.ix1 + 1), ya = max(0, g.iy0), yb = min(FT, g.iy1 + 1);
            const int ncols = xb - xa, n = ncols * (yb - ya);
            const RegionDiv rd(ncols);
            for (int i = threadIdx.x; i < n; i += MMA_THREADS)
            {
                int qy, qx;
                rd.split(i, qy, qx);
                const int x = xa + qx, y = ya + qy;
                float r[9];
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) r[dy * 3 + dx] = luma[(y + dy) * LT + x + dx];
                float v[8];
#pragma unroll
                for (int co = 0; co < 8; co++)
                {
                    // plain FMA chain: this engine is held to the 8-bit tolerance, not to the reference's summation order
                    float s = prm.b[co];
#pragma unroll
                    for (int p = 0; p < 9; p++) s = fmaf(r[p], prm.k[co * 9 + p], s);
                    if (ACT == ACT_RELU) s = fmaxf(s, 0.0f);
                    else if (ACT == ACT_PRELU) s = prelu(s, prm.a[co]);
                    v[co] = s;
                }
                store_pixel_split(A, y * FT + x, v);
                if constexpr (S::FAM == ACB200_FAMILY_ARNET)
                {
                    if (x >= S::R && x < FT - S::R && y >= S::R && y < FT - S::R)
                    {
                        float4* p = reinterpret_cast<float4*>(prm.feat_out + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8);
                        p[0] = make_float4(v[0], v[1], v[2], v[3]);
                        p[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
            }
            __syncthreads();
        }

        constexpr int B0 = S::HEAD ? 8 : 0;     // bias / alpha offsets of the first body conv inside prm.b / prm.a
        constexpr int A0 = (S::FAM == ACB200_FAMILY_ACNET && S::HEAD) ? 8 : 0;
        HalfPlanes cur = A, oth = B;
        // ---- body convs -------------------------------------------------------------------------------------------------
        // A segment without a tail hands its last body conv's results to the next segment: that layer's epilogue stores them
        // straight into the global map (already split, see MmaParams) instead of into shared memory, which spreads the 80 KB per
        // CTA over the layer instead of a store burst at the end of every wave, and saves a pass over the tile.
        auto body_layer = [&](const int i, auto to_map_tag) {
            constexpr bool TO_MAP = decltype(to_map_tag)::value;
            const bool more = i + 1 < S::NCONV || S::TAIL;      // a further 3x3 layer follows in this segment
            if (more) mma_load_bfrag(nb, prm.frags + (i + 1) * FRAG_WORDS_3X3);
            const float b0 = prm.b[B0 + 8 * i + 2 * tq], b1 = prm.b[B0 + 8 * i + 2 * tq + 1];
            int act = ACT_RELU;
            bool res = false;
            float a0 = 0.0f, a1 = 0.0f;
            if constexpr (S::FAM == ACB200_FAMILY_ACNET)
            {
                act = ACT_PRELU;
                a0 = prm.a[A0 + 8 * i + 2 * tq]; a1 = prm.a[A0 + 8 * i + 2 * tq + 1];
            }
            else if constexpr (S::FAM == ACB200_FAMILY_ARNET)
            {
                if ((i & 1) == 0) { act = ACT_PRELU; a0 = prm.a[(i >> 1) * 8 + 2 * tq]; a1 = prm.a[(i >> 1) * 8 + 2 * tq + 1]; }
                else { act = ACT_IDENTITY; res = true; }
            }
            uint32_t* const out_q = reinterpret_cast<uint32_t*>(oth.hi) + tq;      // this lane's channel pair of pixel 0, hi plane
            // TO_MAP: this lane's channel pair of image pixel (g.ox, g.oy) in the hi plane of the global map (32-bit words)
            uint32_t* const map_q = reinterpret_cast<uint32_t*>(prm.map_out) + (static_cast<long long>(g.oy) * prm.w + g.ox) * 4 + tq;
            const size_t map_plane_words = static_cast<size_t>(prm.w) * prm.h * 4;
            auto epi = [&](const int off, float v0, float v1, const bool valid) {
                // no early exit for the overhang lanes: straight-line code with predicated stores (one address, the lo plane
                // is a constant distance away)
                if (act == ACT_RELU) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                else if (act == ACT_PRELU) { v0 = prelu(v0, a0); v1 = prelu(v1, a1); }
                uint32_t* ph = out_q + 4 * off;
                uint32_t* pl = ph + PLANE_BYTES / 4;
                if (res)
                {
                    // (overhang lanes alias the last valid pixel, which its owner rewrites: they must not read it either)
                    const float2 id = valid ? join_pair(*ph, *pl) : make_float2(0.0f, 0.0f);
                    v0 = fmaf(v0, 0.2f, id.x); v1 = fmaf(v1, 0.2f, id.y);
                }
                uint32_t hi, lo;
                split_pair(v0, v1, hi, lo);
                if constexpr (TO_MAP)
                {
                    const int py = off / FT, px = off - py * FT;
                    uint32_t* gp = map_q + (static_cast<long long>(py) * prm.w + px) * 4;
                    if (valid) { gp[0] = hi; gp[map_plane_words] = lo; }
                }
                else if (valid) { *ph = hi; *pl = lo; }
            };
            mma_conv3x3(i + 1, cur, bf, g, b0, b1, epi);
            if constexpr (!TO_MAP)
            {
                __syncthreads();
                if (more)
                {
#pragma unroll
                    for (int e = 0; e < 18; e++) bf[e] = nb[e];
                }
                const HalfPlanes t = cur; cur = oth; oth = t;
            }
        };
        constexpr int NLOOP = S::TAIL ? S::NCONV : S::NCONV - 1;
#pragma unroll 1
        for (int i = 0; i < NLOOP; i++) body_layer(i, std::false_type{});
        if constexpr (!S::TAIL) body_layer(S::NCONV - 1, std::true_type{});

        const uint32_t* tfrag = prm.frags + S::NCONV * FRAG_WORDS_3X3;
        constexpr int BT = B0 + 8 * S::NCONV;
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
        if constexpr (!S::TAIL)
        {
            // (the map went out with the last body conv's epilogue)
        }
        else if constexpr (S::FAM == ACB200_FAMILY_ACNET_LEGACY)
        {
            // conv3x3 + ReLU on the tensor cores, then the 2x2 deconv: each quad of lanes holds the 8 channels of a pixel
            // (2 per lane) -- four partial dots per lane (one per output sub-pixel), reduce-scattered over the quad in three shuffles:
            // lane t ends with sub-pixel t.  The sub-pixels' weights are held in the order {t, t^2, t^1, t^3} so that a lane always
            // keeps its first two slots and sends the other two (no selects): after the exchange with lane t^1 slots 0 / 1 hold the
            // pair sums of sub-pixels t and t^2, the exchange with lane t^2 completes sub-pixel t.
            const float b0 = prm.b[BT + 2 * tq], b1 = prm.b[BT + 2 * tq + 1];
            float kd[4][2];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                const int q = tq ^ (j == 1 ? 2 : j == 2 ? 1 : j == 3 ? 3 : 0);
                kd[j][0] = prm.k[(S::HEAD ? 72 : 0) + q * 8 + 2 * tq]; kd[j][1] = prm.k[(S::HEAD ? 72 : 0) + q * 8 + 2 * tq + 1];
            }
            // 8-bit output: the 2T x 2T result tile is staged in shared memory (the luma tile is dead after the head) and leaves
            // as 16-byte vectors; other element types are stored directly
            constexpr int OT = 2 * S::T, OPITCH = ((OT + 15) / 16) * 16;
            static_assert(OPITCH * OT <= LT * LT * 4, "output staging does not fit the luma tile");
            uint8_t* s_out = reinterpret_cast<uint8_t*>(luma);
#ifndef ACB_STAGE_HEADLESS_TAIL
#define ACB_STAGE_HEADLESS_TAIL 0
#endif
            // (this family's tail never reads the luma tile, so the space is free with or without a head -- but measured on the
            // head-less second segment of the two-segment chain the direct byte stores are 2.5 % faster than staging)
            const bool staged = prm.type == ACB200_UINT8 && (S::HEAD || ACB_STAGE_HEADLESS_TAIL);
            auto epi = [&](const int off, float v0, float v1, const bool valid) {
                const int py = off / FT, px = off - py * FT;
                v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f);
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = fmaf(v1, kd[j][1], v0 * kd[j][0]);
                o[0] += __shfl_xor_sync(0xffffffffu, o[2], 1);
                o[1] += __shfl_xor_sync(0xffffffffu, o[3], 1);
                const float mine = o[0] + __shfl_xor_sync(0xffffffffu, o[1], 2);
                if (valid)
                {
                    if (staged)
                        s_out[(2 * (py - S::R) + (tq >> 1)) * OPITCH + 2 * (px - S::R) + (tq & 1)] = static_cast<uint8_t>(fmaf(sat01(mine), 255.0f, 0.5f));
                    else
                    {
                        void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * (g.oy + py) + (tq >> 1)) * prm.dst_pitch;
                        net_store1(row, 2 * (g.ox + px) + (tq & 1), prm.type, mine);
                    }
                }
            };
            if (prm.type == ACB200_UINT8 && !staged)
            {
                // direct 8-bit stores, written for instruction count: no element-type dispatch per pixel, one 64-bit base per lane
                // and 32-bit offsets, saturate / FMA / convert
                uint8_t* const dq = static_cast<uint8_t*>(prm.dst) + static_cast<long long>(2 * g.oy + (tq >> 1)) * prm.dst_pitch + (2 * g.ox + (tq & 1));
                const int pitch2 = 2 * prm.dst_pitch;
                auto epi8 = [&](const int off, float v0, float v1, const bool valid) {
                    const int py = off / FT, px = off - py * FT;
                    v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f);
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) o[j] = fmaf(v1, kd[j][1], v0 * kd[j][0]);
                    o[0] += __shfl_xor_sync(0xffffffffu, o[2], 1);
                    o[1] += __shfl_xor_sync(0xffffffffu, o[3], 1);
                    const float mine = o[0] + __shfl_xor_sync(0xffffffffu, o[1], 2);
                    if (valid) dq[py * pitch2 + 2 * px] = static_cast<uint8_t>(fmaf(__saturatef(mine), 255.0f, 0.5f));
                };
                mma_conv3x3(S::NCONV + 1, cur, bf, g, b0, b1, epi8);
            }
            else mma_conv3x3(S::NCONV + 1, cur, bf, g, b0, b1, epi);
            if (staged)
            {
                __syncthreads();
                const int vw = min(OT, 2 * (prm.w - (g.ox + S::R))), vh = min(OT, 2 * (prm.h - (g.oy + S::R)));     // valid part of the output tile
                uint8_t* tile_dst = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * (g.oy + S::R)) * prm.dst_pitch + 2 * (g.ox + S::R);
                if (((reinterpret_cast<uintptr_t>(tile_dst) | static_cast<uintptr_t>(prm.dst_pitch)) & 15) == 0)
                {
                    const int vecs = vw >> 4;
                    for (int i = threadIdx.x; i < vh * (OPITCH / 16); i += MMA_THREADS)
                    {
                        const int row = i / (OPITCH / 16), v = i % (OPITCH / 16);
                        if (v < vecs) reinterpret_cast<uint4*>(tile_dst + static_cast<size_t>(row) * prm.dst_pitch)[v] = reinterpret_cast<const uint4*>(s_out + row * OPITCH)[v];
                        else
                            for (int e = v * 16; e < min(vw, v * 16 + 16); e++) tile_dst[static_cast<size_t>(row) * prm.dst_pitch + e] = s_out[row * OPITCH + e];
                    }
                }
                else
                    for (int i = threadIdx.x; i < vh * vw; i += MMA_THREADS)
                        tile_dst[static_cast<size_t>(i / vw) * prm.dst_pitch + (i % vw)] = s_out[(i / vw) * OPITCH + (i % vw)];
            }
        }
        else
        {
            int LPS = S::NCONV + 1;
            if constexpr (S::FAM == ACB200_FAMILY_ARNET)
            {
                // last block: PReLU conv, then conv *0.2 + x fused with the 1x1 (an m16n8k8 on the re-split sums: the D
                // fragment layout of the 3x3 IS the A fragment layout of a k8 MMA), PReLU, + feat
                constexpr int AT = (S::NCONV / 2) * 8;
                {
                    const float b0 = prm.b[BT + 2 * tq], b1 = prm.b[BT + 2 * tq + 1], a0 = prm.a[AT + 2 * tq], a1 = prm.a[AT + 2 * tq + 1];
                    const HalfPlanes out = oth;
                    auto epi = [&](const int off, float v0, float v1, const bool valid) {
                        if (!valid) return;
                        v0 = prelu(v0, a0); v1 = prelu(v1, a1);
                        uint32_t hi, lo;
                        split_pair(v0, v1, hi, lo);
                        reinterpret_cast<uint32_t*>(out.hi + off)[tq] = hi;
                        reinterpret_cast<uint32_t*>(out.lo + off)[tq] = lo;
                    };
                    mma_load_bfrag(nb, tfrag + FRAG_WORDS_3X3);
                    mma_conv3x3(S::NCONV + 1, cur, bf, g, b0, b1, epi);
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 18; e++) bf[e] = nb[e];
                }
                {
                    const uint32_t* f3 = tfrag + FRAG_WORDS_3X3;
                    const uint32_t* f1 = f3 + FRAG_WORDS_3X3;
                    const uint32_t w1h = __ldg(f1 + lane), w1l = __ldg(f1 + 32 + lane);
                    const float b0 = prm.b[BT + 8 + 2 * tq], b1 = prm.b[BT + 8 + 2 * tq + 1];
                    const float c0 = prm.b[BT + 16 + 2 * tq], c1 = prm.b[BT + 16 + 2 * tq + 1];
                    const float a0 = prm.a[AT + 8 + 2 * tq], a1 = prm.a[AT + 8 + 2 * tq + 1];
                    const HalfPlanes out = cur;     // x, updated in place
                    // the epilogue is called for (px, py) then (px + 8, py): collect both halves of the fragment, then run
                    // the 1x1 as tensor-core MMAs on the pair
                    float keep[2];
                    int koff = 0, phase = 0;
                    bool kvalid = false;
                    auto epi = [&](const int off, float v0, float v1, const bool valid) {
                        uint32_t* ph = reinterpret_cast<uint32_t*>(out.hi + off) + tq;
                        uint32_t* pl = reinterpret_cast<uint32_t*>(out.lo + off) + tq;
                        // overhang lanes of the last tile alias its last valid pixel: they must not touch it
                        const float2 id = valid ? join_pair(*ph, *pl) : make_float2(0.0f, 0.0f);
                        v0 = fmaf(v0, 0.2f, id.x); v1 = fmaf(v1, 0.2f, id.y);
                        if (phase == 0) { keep[0] = v0; keep[1] = v1; koff = off; kvalid = valid; phase = 1; return; }
                        phase = 0;
                        uint32_t h0, l0, h1, l1;
                        split_pair(keep[0], keep[1], h0, l0);
                        split_pair(v0, v1, h1, l1);
                        float d[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                        mma_k8(d, h0, h1, w1h);
                        mma_k8(d, l0, l1, w1h);
                        mma_k8(d, h0, h1, w1l);
                        const int offs[2] = { koff, off };
                        const bool oks[2] = { kvalid, valid };
#pragma unroll
                        for (int half = 0; half < 2; half++)
                        {
                            const int qo = offs[half], qy = qo / FT, qx = qo - qy * FT;
                            if (!oks[half]) continue;
                            float u0 = prelu(d[2 * half] + c0, a0), u1 = prelu(d[2 * half + 1] + c1, a1);
                            const int gx = clampi(g.ox + qx, 0, prm.w - 1), gy = clampi(g.oy + qy, 0, prm.h - 1);
                            const float2 ft = *reinterpret_cast<const float2*>(prm.feat_in + (static_cast<size_t>(gy) * prm.w + gx) * 8 + 2 * tq);
                            u0 += ft.x; u1 += ft.y;
                            uint32_t hi, lo;
                            split_pair(u0, u1, hi, lo);
                            reinterpret_cast<uint32_t*>(out.hi + qo)[tq] = hi;
                            reinterpret_cast<uint32_t*>(out.lo + qo)[tq] = lo;
                        }
                    };
                    mma_load_bfrag(nb, f1 + FRAG_WORDS_1X1);
                    mma_conv3x3(S::NCONV + 2, oth, bf, g, b0, b1, epi);
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 18; e++) bf[e] = nb[e];
                    LPS = S::NCONV + 3;
                }
            }
            // pixel-shuffle tail: conv3x3 8->4 (couts 4-7 are zero weights), + nearest-upsampled luma; lane t=0 owns output
            // row 2y (couts 0,1), lane t=1 owns row 2y+1 (couts 2,3)
            constexpr int BPS = S::FAM == ACB200_FAMILY_ARNET ? BT + 24 : BT;
            const float b0 = tq < 2 ? prm.b[BPS + 2 * tq] : 0.0f, b1 = tq < 2 ? prm.b[BPS + 2 * tq + 1] : 0.0f;
            constexpr int LT_PS = S::FAM == ACB200_FAMILY_ARNET ? S::NCONV + 3 : S::NCONV + 1;
            auto epi = [&](const int off, float v0, float v1, const bool valid) {
                if (tq < 2 && valid)
                {
                    const int py = off / FT, px = off - py * FT;
                    const float id = luma[(py + 1) * LT + px + 1];
                    void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * (g.oy + py) + tq) * prm.dst_pitch;
                    net_store2(row, 2 * (g.ox + px), prm.type, v0 + id, v1 + id, aligned);
                }
            };
            (void)LPS;
            if (prm.type == ACB200_UINT8 && aligned)
            {
                // 8-bit fast path: no element-type dispatch per pixel, one base pointer per lane and 32-bit offsets
                uint8_t* const dq = static_cast<uint8_t*>(prm.dst) + static_cast<long long>(2 * g.oy + tq) * prm.dst_pitch + 2 * g.ox;
                const int pitch2 = 2 * prm.dst_pitch;
                auto epi8 = [&](const int off, float v0, float v1, const bool valid) {
                    if (tq < 2 && valid)
                    {
                        const int py = off / FT, px = off - py * FT;
                        const float id = luma[(py + 1) * LT + px + 1];
                        const uint8_t q0 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v0 + id), 255.0f), 0.5f));
                        const uint8_t q1 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v1 + id), 255.0f), 0.5f));
                        *reinterpret_cast<uchar2*>(dq + py * pitch2 + 2 * px) = make_uchar2(q0, q1);
                    }
                };
                mma_conv3x3(LT_PS, cur, bf, g, b0, b1, epi8);
            }
            else mma_conv3x3(LT_PS, cur, bf, g, b0, b1, epi);
        }
    }

    constexpr size_t MMA_SMEM_BYTES = MMA_SMEM_DATA_BYTES + 16;     // + the mbarrier of the map copy
}
