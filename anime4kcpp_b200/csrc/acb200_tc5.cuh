// tcgen05 engine of the fused luma network: split-fp16 implicit GEMM on the 5th-generation tensor cores, accumulators in
// TMEM, operands read straight from the shared-memory activation planes through matrix descriptors.
//
// Same tiling, segment chain, weights and reference semantics as the other two engines (acb200_ffma.cuh, acb200_mma.cuh);
// what changes is how a 3x3 conv layer over the 56x56 frame is evaluated:
//
//  * The activation planes are [flat frame pixel][8 ch fp16] (16 B / pixel), hi plane followed by lo plane.  In the
//    no-swizzle K-major canonical layout an 8x16-byte core matrix is 8 consecutive pixels, SBO = 128 B walks 128 consecutive
//    pixels (one MMA's M rows = TMEM lanes), and LBO = the plane distance makes the two K chunks of a k-step the hi and lo
//    halves of the same pixel.  A vertical tap is nothing but the descriptor start address moved by +-56 pixels.
//  * A measured SS-mode tcgen05.mma (M=128, K=16) costs 57 cycles for ANY N <= 112 (tools/microbench_tcgen05.cu), so the three
//    horizontal taps and both weight halves are folded into N = 48:
//        D[q][dx*16 + j]   j < 8 : sum_dy (a_hi + a_lo)[q + 56 dy] . w_hi[dy][dx][:, co=j]
//                          j >= 8: sum_dy  a_hi       [q + 56 dy] . w_lo[dy][dx][:, co=j-8]
//    i.e. THREE MMAs per 128 pixels and layer (dy = -1, 0, +1 accumulate into the same TMEM columns).
//  * out[p] = C_-1[p-1] + C_0[p] + C_+1[p+1] with C_dx = the two halves of column group dx added.  TMEM lane = pixel, so the
//    epilogue thread of lane L finalises pixel base+L-1 from its own C_+1 and two warp shuffles; the first two lanes of a
//    warp take their neighbours' values from a 96-byte shared-memory mailbox, and consecutive tiles overlap by two pixels
//    (tile stride 126) so no tile ever waits for another.
//  * Replicate padding: the MMA cannot clamp coordinates, so CTAs that touch the image border copy the edge values one pixel
//    outwards after every layer (border CTAs only, uniform branch).
//
// Warp roles (640 threads): lane 0 of warps 0 and 1 issue the MMAs (one thread per TMEM buffer); warps 2-3 prefetch the next
// layer's B operand into shared memory; warps 4-19 are four epilogue groups (one warp per TMEM lane quadrant each), group g
// taking tile g of every batch.  Two batches of four tiles are in flight in TMEM (2 x 4 x 64 columns) with full/empty mbarriers
// per batch buffer; ONE tcgen05.commit per batch signals completion (a commit drains the tensor pipe, so it is amortised).
#pragma once

#include <cuda_fp16.h>

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"
#include "acb200_mma.cuh"

namespace acb
{
    constexpr int TC_GROUPS = 4;                // epilogue groups of 4 warps (one warp per TMEM lane quadrant)
    constexpr int TC_THREADS = 128 + TC_GROUPS * 128;
    constexpr int TC_SLOTS = 8;                 // TMEM accumulator slots of 64 columns
    constexpr int TC_N = 48;                    // 3 horizontal taps x (8 couts with w_hi | 8 couts with w_lo)
    constexpr int TC_TILE_STRIDE = 126;         // 128 MMA rows, 126 finished pixels
    constexpr int TC_PLANE_BYTES = FT * FT * 16;
    constexpr int TC_B_BYTES_DY = 2 * TC_N * 16;            // one dy: [2 K chunks][48 rows][8 fp16]
    constexpr int TC_B_BYTES_LAYER = 3 * TC_B_BYTES_DY;     // 4608
    constexpr int TC_B_WORDS_LAYER = TC_B_BYTES_LAYER / 4;

    template<class S>
    struct Tc5Params
    {
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
        const uint32_t* bops;   // packed B operands of this segment's 3x3 convs, TC_B_WORDS_LAYER words each, in layer order
        float k[(S::HEAD ? 72 : 0) + 64 + 32];  // fp32 weights used outside the MMAs: head (72) | ARNet 1x1 (64) | legacy deconv (32)
        float b[S::NB];
        float a[S::NA > 0 ? S::NA : 1];
    };

    __device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
    __device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
    {
        return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
               (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
    }
    __device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    }
    __device__ __forceinline__ void tc_commit(uint32_t bar)
    {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    __device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity)
    {
        uint32_t ok = 0, spins = 0;
        do
        {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (!ok && ++spins > (1u << 24)) __trap();      // a protocol bug must fault, never hang the device
        } while (!ok);
    }
    __device__ __forceinline__ void tc_mbar_arrive(uint32_t bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
    }
    __device__ __forceinline__ void tc_ld16(uint32_t (&v)[16], uint32_t taddr)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    }
    __device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

    // shared-memory carve-up (bytes from the dynamic smem base)
    constexpr int TC_OFF_A = 0;                                     // buffer A: hi plane, lo plane
    constexpr int TC_OFF_B = 2 * TC_PLANE_BYTES;                    // buffer B: hi plane, lo plane
    constexpr int TC_OFF_LUMA = 4 * TC_PLANE_BYTES;                 // 58x58 float luma tile (also the read slack past the last plane)
    constexpr int TC_OFF_BOP = TC_OFF_LUMA + ((LT * LT * 4 + 127) / 128) * 128;     // 2 x B operand of one layer
    constexpr int TC_OFF_XCH = TC_OFF_BOP + 2 * TC_B_BYTES_LAYER;   // mailboxes: [groups][2 parities][4 warps][3 vectors][8 floats]
    constexpr int TC_OFF_BAR = TC_OFF_XCH + TC_GROUPS * 2 * 4 * 3 * 8 * 4;  // full[8], empty[8] mbarriers, tmem slot
    constexpr int TC_SMEM_BYTES = TC_OFF_BAR + 16 * 8 + 16;
    static_assert(TC_SMEM_BYTES <= 232448, "tcgen05 engine exceeds the 227 KB shared-memory limit");

    struct Tc5Ctx
    {
        unsigned char* smem;
        uint32_t tmem;
        uint32_t tile_counter;      // tile batches issued / consumed so far (same sequence in the MMA warp and the epilogue warps)
        bool border;
    };

    // copy the image-edge pixels one pixel outwards inside the frame (replicate padding for the next layer's MMA reads)
    __device__ __forceinline__ void tc_replicate_border(const HalfPlanes& p, const TileGeom& g)
    {
        const int x0 = g.ix0, x1 = g.ix1, y0 = g.iy0, y1 = g.iy1;       // image bounds in frame coordinates (may lie outside the frame)
        const int ya = max(y0, 0), yb = min(y1, FT - 1), xa = max(x0 - 1, 0), xb = min(x1 + 1, FT - 1);
        // columns first (in-image rows), then rows (including the freshly written column pixels)
        for (int i = threadIdx.x; i < 2 * FT; i += TC_THREADS)
        {
            const int y = i >> 1, side = i & 1;
            if (y < ya || y > yb) continue;
            if (side == 0 && x0 >= 1 && x0 <= FT - 1) { p.hi[y * FT + x0 - 1] = p.hi[y * FT + x0]; p.lo[y * FT + x0 - 1] = p.lo[y * FT + x0]; }
            if (side == 1 && x1 >= 0 && x1 <= FT - 2) { p.hi[y * FT + x1 + 1] = p.hi[y * FT + x1]; p.lo[y * FT + x1 + 1] = p.lo[y * FT + x1]; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * FT; i += TC_THREADS)
        {
            const int x = i >> 1, side = i & 1;
            if (x < xa || x > xb) continue;
            if (side == 0 && y0 >= 1 && y0 <= FT - 1) { p.hi[(y0 - 1) * FT + x] = p.hi[y0 * FT + x]; p.lo[(y0 - 1) * FT + x] = p.lo[y0 * FT + x]; }
            if (side == 1 && y1 >= 0 && y1 <= FT - 2) { p.hi[(y1 + 1) * FT + x] = p.hi[y1 * FT + x]; p.lo[(y1 + 1) * FT + x] = p.lo[y1 * FT + x]; }
        }
    }

    // One 3x3 conv layer on tcgen05.  `in` must already carry replicate padding if the CTA touches the image border.
    //   fin(px, py, v[8], valid): called by every epilogue thread once per tile with the finished fp32 sums (bias NOT included)
    //   of all 8 output channels of frame pixel (px, py); valid == false for pixels outside the layer's image-clipped region.
    //   bop_smem: this layer's B operand in shared memory (3 x [2][48][8] fp16).
    template<class Fin>
    __device__ __forceinline__ void tc5_conv3x3(Tc5Ctx& ctx, const int L, const HalfPlanes& in, const uint32_t bop_smem, const TileGeom& g, Fin&& fin)
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int xa = max(L, g.ix0), xb = min(FT - L, g.ix1 + 1), ya = max(L, g.iy0), yb = min(FT - L, g.iy1 + 1);
        // flat range of output pixels of the (unclipped) layer region; tiles of 126 finished pixels
        const int p_first = L * FT + L, p_last = (FT - 1 - L) * FT + (FT - 1 - L);
        const int tiles = (p_last - p_first + TC_TILE_STRIDE) / TC_TILE_STRIDE;
        const uint32_t bars = tc_smem_u32(ctx.smem + TC_OFF_BAR);
        // Tiles are issued in batches of TC_GROUPS (one per epilogue group) with ONE tcgen05.commit per batch: a commit drains
        // the tensor pipe (~240 cycles measured, tools/microbench_tcgen05.cu), so it is amortised over 12 MMAs.  Two batches
        // (2 x 4 TMEM slots of 64 columns) are in flight: full[buf] / empty[buf] mbarriers.
        const int batches = (tiles + TC_GROUPS - 1) / TC_GROUPS;
        const uint32_t b0 = ctx.tile_counter;       // batches issued so far in this kernel
        if (warp < 2)
        {
            // two issuing threads (lane 0 of warps 0 and 1), one per TMEM buffer: while one thread sits in its mbarrier wait or
            // its commit drain, the other thread's MMAs keep the tensor pipe busy
            if (lane == 0)
            {
                const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(TC_N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
                const uint32_t a_base = tc_smem_u32(in.hi);
                for (int bi = 0; bi < batches; bi++)
                {
                    const uint32_t bt = b0 + bi, buf = bt & 1, use = bt >> 1;
                    if (static_cast<int>(buf) != warp) continue;
                    if (use > 0) tc_mbar_wait(bars + (2 + buf) * 8, (use - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int k = 0; k < TC_GROUPS; k++)
                    {
                        const int j = bi * TC_GROUPS + k;
                        if (j >= tiles) break;
                        const int base = p_first - 1 + j * TC_TILE_STRIDE;
#pragma unroll
                        for (int dy = 0; dy < 3; dy++)
                            tc_mma(ctx.tmem + (buf * TC_GROUPS + k) * 64, tc_desc(a_base + (base + (dy - 1) * FT) * 16, TC_PLANE_BYTES, 128),
                                   tc_desc(bop_smem + dy * TC_B_BYTES_DY, TC_N * 16, 128), idesc, dy > 0);
                    }
                    tc_commit(bars + buf * 8);
                }
            }
        }
        else if (warp >= 4)
        {
            const int group = (warp - 4) >> 2, quad = warp & 3;
            float4* xch_base = reinterpret_cast<float4*>(ctx.smem + TC_OFF_XCH) + group * (2 * 4 * 3 * 2);
            for (int bi = 0; bi < batches; bi++)
            {
                const int j = bi * TC_GROUPS + group;
                float4* xch = xch_base + (bi & 1) * (4 * 3 * 2);     // two mailboxes per group, alternating: one barrier per tile
                const uint32_t bt = b0 + bi, buf = bt & 1, use = bt >> 1;
                tc_mbar_wait(bars + buf * 8, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (j >= tiles)
                {
                    // this group has no tile in the last batch: just release the buffer
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    tc_mbar_arrive(bars + (2 + buf) * 8);
                    continue;
                }
                uint32_t r[3][16];
                const uint32_t taddr = ctx.tmem + (buf * TC_GROUPS + group) * 64 + (static_cast<uint32_t>(quad * 32) << 16);
                tc_ld16(r[0], taddr);
                tc_ld16(r[1], taddr + 16);
                tc_ld16(r[2], taddr + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tc_mbar_arrive(bars + (2 + buf) * 8);
                float cm[8], c0[8], v[8];      // C_-1, C_0 of this lane's pixel; v starts as its C_+1
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    cm[c] = __uint_as_float(r[0][c]) + __uint_as_float(r[0][8 + c]);
                    c0[c] = __uint_as_float(r[1][c]) + __uint_as_float(r[1][8 + c]);
                    v[c] = __uint_as_float(r[2][c]) + __uint_as_float(r[2][8 + c]);
                }
                // mailbox for the next warp's first two lanes: C_-1[30], C_-1[31], C_0[31]
                if (lane >= 30)
                {
                    float4* box = xch + (quad * 3 + (lane - 30)) * 2;
                    box[0] = make_float4(cm[0], cm[1], cm[2], cm[3]);
                    box[1] = make_float4(cm[4], cm[5], cm[6], cm[7]);
                    if (lane == 31)
                    {
                        box[2] = make_float4(c0[0], c0[1], c0[2], c0[3]);
                        box[3] = make_float4(c0[4], c0[5], c0[6], c0[7]);
                    }
                }
                float ca[8], cb[8];    // C_-1[q-2], C_0[q-1] (lanes 0,1 of a warp: replaced from the mailbox below)
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    ca[c] = __shfl_up_sync(0xffffffffu, cm[c], 2);
                    cb[c] = __shfl_up_sync(0xffffffffu, c0[c], 1);
                }
                asm volatile("bar.sync %0, 128;" :: "r"(1 + group) : "memory");
                if (lane < 2 && quad > 0)
                {
                    const float4* prev = xch + ((quad - 1) * 3) * 2;
                    const float4 m0 = prev[lane * 2], m1 = prev[lane * 2 + 1];             // C_-1[prev 30] (lane 0) / C_-1[prev 31] (lane 1)
                    ca[0] = m0.x; ca[1] = m0.y; ca[2] = m0.z; ca[3] = m0.w; ca[4] = m1.x; ca[5] = m1.y; ca[6] = m1.z; ca[7] = m1.w;
                    if (lane == 0)
                    {
                        const float4 n0 = prev[4], n1 = prev[5];                           // C_0[prev 31]
                        cb[0] = n0.x; cb[1] = n0.y; cb[2] = n0.z; cb[3] = n0.w; cb[4] = n1.x; cb[5] = n1.y; cb[6] = n1.z; cb[7] = n1.w;
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = (ca[c] + cb[c]) + v[c];
                const int p = p_first - 1 + j * TC_TILE_STRIDE + quad * 32 + lane - 1;
                const int py = p / FT, px = p - py * FT;
                const bool valid = (quad > 0 || lane >= 2) && px >= xa && px < xb && py >= ya && py < yb && p <= p_last;
                fin(px, py, v, valid);
            }
        }
        ctx.tile_counter = b0 + batches;
    }

    template<class S>
    __global__ void __launch_bounds__(TC_THREADS, 1) segment_tc5_kernel(const __grid_constant__ Tc5Params<S> prm)
    {
        extern __shared__ __align__(1024) unsigned char smem[];
        HalfPlanes A{ reinterpret_cast<uint4*>(smem + TC_OFF_A), reinterpret_cast<uint4*>(smem + TC_OFF_A + TC_PLANE_BYTES) };
        HalfPlanes B{ reinterpret_cast<uint4*>(smem + TC_OFF_B), reinterpret_cast<uint4*>(smem + TC_OFF_B + TC_PLANE_BYTES) };
        float* luma = reinterpret_cast<float*>(smem + TC_OFF_LUMA);
        uint32_t* bop = reinterpret_cast<uint32_t*>(smem + TC_OFF_BOP);
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
        const int warp = threadIdx.x >> 5;

        const int tx = blockIdx.x % prm.tiles_x, ty = blockIdx.x / prm.tiles_x;
        TileGeom g;
        g.ox = tx * S::T - S::R;
        g.oy = ty * S::T - S::R;
        g.ix0 = -g.ox; g.ix1 = prm.w - 1 - g.ox;
        g.iy0 = -g.oy; g.iy1 = prm.h - 1 - g.oy;

        // ---- one-time setup: barriers, TMEM, first layer's B operand, input tile ------------------------------------------------
        if (threadIdx.x == 0)
        {
            for (int i = 0; i < 2; i++)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(bars + i)));                           // full[buf]
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tc_smem_u32(bars + 2 + i)), "r"(TC_GROUPS * 128));  // empty[buf]
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0)
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        constexpr int NLAYERS = S::NCONV + S::TAIL_LAYERS;      // 3x3 convs in this segment
        for (int i = threadIdx.x; i < TC_B_WORDS_LAYER; i += TC_THREADS) bop[i] = __ldg(prm.bops + i);
        if constexpr (S::NEEDS_LUMA)
        {
            for (int i = threadIdx.x; i < LT * LT; i += TC_THREADS)
            {
                const int lx = i % LT, ly = i / LT;
                const int gx = clampi(g.ox - 1 + lx, 0, prm.w - 1), gy = clampi(g.oy - 1 + ly, 0, prm.h - 1);
                luma[i] = load_elem(static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch, gx, prm.type);
            }
        }
        if constexpr (!S::HEAD)
        {
            for (int i = threadIdx.x; i < FT * FT; i += TC_THREADS)
            {
                const int fx = i % FT, fy = i / FT;
                const int gx = clampi(g.ox + fx, 0, prm.w - 1), gy = clampi(g.oy + fy, 0, prm.h - 1);
                const float4* p = reinterpret_cast<const float4*>(prm.map_in + (static_cast<size_t>(gy) * prm.w + gx) * 8);
                const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
                const float v[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
                store_pixel_split(A, i, v);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        Tc5Ctx ctx;
        ctx.smem = smem;
        ctx.tmem = *tmem_slot;
        ctx.tile_counter = 0;
        ctx.border = !(g.ix0 <= 0 && g.iy0 <= 0 && g.ix1 >= FT - 1 && g.iy1 >= FT - 1);

        if constexpr (S::HEAD)
        {
            // 1 -> 8 head conv in fp32, one pixel per thread; the luma tile is replicate-padded by construction, so the head
            // output is written for the WHOLE frame (out-of-image positions get the value of their clamped pixel's window,
            // which the border pass below overwrites with the true replicated values)
            constexpr int ACT = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? ACT_RELU : S::FAM == ACB200_FAMILY_ACNET ? ACT_PRELU : ACT_IDENTITY;
            const int xa = max(0, g.ix0), xb = min(FT, g.ix1 + 1), ya = max(0, g.iy0), yb = min(FT, g.iy1 + 1);
            const int ncols = xb - xa, n = ncols * (yb - ya);
            for (int i = threadIdx.x; i < n; i += TC_THREADS)
            {
                const int x = xa + i % ncols, y = ya + i / ncols;
                float r[9];
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) r[dy * 3 + dx] = luma[(y + dy) * LT + x + dx];
                float v[8];
#pragma unroll
                for (int co = 0; co < 8; co++)
                {
                    float q[8];
#pragma unroll
                    for (int p = 0; p < 8; p++) q[p] = __fmul_rn(r[p], prm.k[co * 9 + p]);
                    float s = __fadd_rn(fmaf(r[8], prm.k[co * 9 + 8], hsum8(q)), prm.b[co]);
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
        if (ctx.border) { tc_replicate_border(A, g); }
        tc_fence_async_smem();
        __syncthreads();

        constexpr int B0 = S::HEAD ? 8 : 0;
        constexpr int A0 = (S::FAM == ACB200_FAMILY_ACNET && S::HEAD) ? 8 : 0;
        constexpr int BT = B0 + 8 * S::NCONV;
        HalfPlanes cur = A, oth = B;
        const int es = prm.type & 0xff;
        const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;

#pragma unroll 1
        for (int li = 0; li < NLAYERS; li++)
        {
            const uint32_t bop_smem = tc_smem_u32(bop) + (li & 1) * TC_B_BYTES_LAYER;
            // helper warps: next layer's B operand into the other half of the double buffer
            if (warp >= 2 && warp < 4 && li + 1 < NLAYERS)
                for (int i = threadIdx.x - 64; i < TC_B_WORDS_LAYER; i += 64)
                    bop[((li + 1) & 1) * TC_B_WORDS_LAYER + i] = __ldg(prm.bops + (li + 1) * TC_B_WORDS_LAYER + i);

            const bool is_body = li < S::NCONV;
            int act = ACT_RELU;
            bool res = false;
            int boff = B0 + 8 * li, aoff = 0;
            // what the layer's finished sums turn into
            enum { K_STORE, K_DECONV, K_SHUFFLE, K_ARNET_END } kind = K_STORE;
            if (is_body)
            {
                if (S::FAM == ACB200_FAMILY_ACNET) { act = ACT_PRELU; aoff = A0 + 8 * li; }
                else if (S::FAM == ACB200_FAMILY_ARNET)
                {
                    if ((li & 1) == 0) { act = ACT_PRELU; aoff = (li >> 1) * 8; }
                    else { act = ACT_IDENTITY; res = true; }
                }
            }
            else if (S::FAM == ACB200_FAMILY_ACNET_LEGACY) { kind = K_DECONV; boff = BT; }
            else if (S::FAM == ACB200_FAMILY_ACNET) { kind = K_SHUFFLE; boff = BT; act = ACT_IDENTITY; }
            else
            {
                const int tl = li - S::NCONV;       // ARNet tail: 0 PReLU conv, 1 residual conv + 1x1, 2 pixel-shuffle conv
                constexpr int AT = (S::NCONV / 2) * 8;
                if (tl == 0) { act = ACT_PRELU; boff = BT; aoff = AT; }
                else if (tl == 1) { kind = K_ARNET_END; boff = BT + 8; aoff = AT + 8; }
                else { kind = K_SHUFFLE; boff = BT + 24; act = ACT_IDENTITY; }
            }
            // Buffers: every map-writing layer reads `cur`, writes `oth`, then the two swap.  An ARNet residual layer reads t
            // (= `cur` after the swap of the preceding PReLU conv) and updates x in place -- x is exactly what `oth` holds then.
            const HalfPlanes src_planes = cur;
            const HalfPlanes dst_planes = oth;
            auto fin = [&](const int px, const int py, float (&v)[8], const bool valid) {
                if (!valid) return;
                const int o = py * FT + px;
                if (kind == K_STORE)
                {
#pragma unroll
                    for (int c = 0; c < 8; c++)
                    {
                        float s = v[c] + prm.b[boff + c];
                        if (act == ACT_RELU) s = fmaxf(s, 0.0f);
                        else if (act == ACT_PRELU) s = prelu(s, prm.a[aoff + c]);
                        v[c] = s;
                    }
                    if (res)
                    {
                        float id[8];
                        load_pixel_joined(dst_planes, o, id);
#pragma unroll
                        for (int c = 0; c < 8; c++) v[c] = fmaf(v[c], 0.2f, id[c]);
                    }
                    store_pixel_split(dst_planes, o, v);
                }
                else if (kind == K_DECONV)
                {
                    float t[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) t[c] = fmaxf(v[c] + prm.b[boff + c], 0.0f);
                    constexpr int KD = (S::HEAD ? 72 : 0) + 64;
#pragma unroll
                    for (int dy = 0; dy < 2; dy++)
                    {
                        float o2[2];
#pragma unroll
                        for (int dx = 0; dx < 2; dx++)
                        {
                            float s = 0.0f;
#pragma unroll
                            for (int c = 0; c < 8; c++) s = fmaf(t[c], prm.k[KD + (dy * 2 + dx) * 8 + c], s);
                            o2[dx] = s;
                        }
                        void* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * (g.oy + py) + dy) * prm.dst_pitch;
                        net_store2(row, 2 * (g.ox + px), prm.type, o2[0], o2[1], aligned);
                    }
                }
                else if (kind == K_SHUFFLE)
                {
                    const float id = luma[(py + 1) * LT + px + 1];
                    uint8_t* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * (g.oy + py)) * prm.dst_pitch;
                    net_store2(row, 2 * (g.ox + px), prm.type, (v[0] + prm.b[boff + 0]) + id, (v[1] + prm.b[boff + 1]) + id, aligned);
                    net_store2(row + prm.dst_pitch, 2 * (g.ox + px), prm.type, (v[2] + prm.b[boff + 2]) + id, (v[3] + prm.b[boff + 3]) + id, aligned);
                }
                else
                {
                    // ARNet end of body: *0.2 + x, 1x1 + bias, PReLU, + feat (Common.hpp:223-288); all 8 channels sit in this thread
                    float id[8], t[8], u[8];
                    load_pixel_joined(dst_planes, o, id);
#pragma unroll
                    for (int c = 0; c < 8; c++) t[c] = fmaf(v[c] + prm.b[boff + c], 0.2f, id[c]);
                    const float* f = prm.feat_in + (static_cast<size_t>(g.oy + py) * prm.w + (g.ox + px)) * 8;
                    const float4 f0 = *reinterpret_cast<const float4*>(f), f1 = *reinterpret_cast<const float4*>(f + 4);
                    const float ft[8] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w };
                    constexpr int K1 = (S::HEAD ? 72 : 0);
#pragma unroll
                    for (int co = 0; co < 8; co++)
                    {
                        float s = prm.b[boff + 8 + co];
#pragma unroll
                        for (int ci = 0; ci < 8; ci++) s = fmaf(t[ci], prm.k[K1 + co * 8 + ci], s);
                        u[co] = prelu(s, prm.a[aoff + co]) + ft[co];
                    }
                    store_pixel_split(dst_planes, o, u);
                }
            };
            tc5_conv3x3(ctx, li + 1, src_planes, bop_smem, g, fin);
            // layer done: epilogue stores visible to the async proxy, then (border CTAs) replicate padding of the new map
            tc_fence_async_smem();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool writes_map = kind == K_STORE || kind == K_ARNET_END;
            if (writes_map && ctx.border && li + 1 < NLAYERS)
            {
                tc_replicate_border(dst_planes, g);
                tc_fence_async_smem();
                __syncthreads();
            }
            if (writes_map) { const HalfPlanes t = cur; cur = oth; oth = t; }
        }

        if constexpr (!S::TAIL)
        {
            const int xa = max(S::R, g.ix0), xb = min(FT - S::R, g.ix1 + 1), ya = max(S::R, g.iy0), yb = min(FT - S::R, g.iy1 + 1);
            const int ncols = xb - xa, n = ncols * (yb - ya);
            for (int i = threadIdx.x; i < n; i += TC_THREADS)
            {
                const int x = xa + i % ncols, y = ya + i / ncols;
                float v[8];
                load_pixel_joined(cur, y * FT + x, v);
                float4* p = reinterpret_cast<float4*>(prm.map_out + (static_cast<size_t>(g.oy + y) * prm.w + (g.ox + x)) * 8);
                p[0] = make_float4(v[0], v[1], v[2], v[3]);
                p[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(ctx.tmem), "r"(512u));
    }
}
