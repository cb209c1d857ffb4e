// Tensor engine of the F -> F conv layers of ArtCNN<16/32> / FSRCNNX<16>: split-fp16 implicit GEMM on tcgen05 (UTCHMMA) with
// the accumulators in TMEM.  With 16 / 32 features this contraction finally has a tensor-core shape (K = 9 F = 144 / 288,
// N = F), and the accumulator read-back that bounds the 8-feature tcgen05 engine is amortised over 4-16x more math per pixel.
//
// One launch per layer over the same fp32 `[h][w][F]` maps as the exact kernels (acb200_wide.cuh), so head, 1x1 and
// pixel-shuffle tail stay shared.  A CTA owns a 64 x TH tile of output pixels:
//  * it loads the (TH+2) x 66 input tile with clamp-to-edge coordinates (= the layer's replicate padding), splits every fp32
//    value into an fp16 (hi, lo) pair and stores it as [8-channel chunk][hi | lo plane][flat tile pixel][8 x fp16] (16 B per
//    pixel): the no-swizzle K-major canonical layout, in which 8 consecutive pixels are one core matrix (SBO = 128 B), the two
//    K halves of a k16 step are the hi and lo plane of one chunk (LBO = plane distance), and a 3x3 tap is the descriptor start
//    address moved by dy*66 + dx pixels;
//  * an M-tile is 128 flat tile pixels (the two pad columns of each row are computed and dropped); one MMA per (tap, chunk):
//        D[q][j]     += (a_hi + a_lo)[q + dy*66 + dx][chunk] . w_hi[tap][chunk][:, j]          j <  F
//        D[q][F + j] +=  a_hi        [q + dy*66 + dx][chunk] . w_lo[tap][chunk][:, j]
//    i.e. N = 2F, 9 F/8 MMAs per M-tile, all accumulating into the same 2F TMEM columns;
//  * warp 0 issues; the other fifteen warps load the tile band by band (an mbarrier per M-tile tells the issuer that the rows
//    this tile reads have landed, so the MMAs of the first tiles overlap the loads of the later rows); warps 4-7 (one per TMEM
//    lane quadrant) then read the finished tiles back (tcgen05.ld), add the two halves, apply bias / activation / residual and
//    store fp32.  Eight TMEM slots hold the finished tiles; one tcgen05.commit per 2 (F = 32) or 4 (F = 16) tiles.
#pragma once

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"
#include "acb200_mma.cuh"
#include "acb200_tc5.cuh"

namespace acb
{
    constexpr int WTC_TW = 64, WTC_PITCH = WTC_TW + 2;
    constexpr int WTC_SLOT_COLS = 64;           // TMEM columns per finished M-tile

    template<int F>
    struct WideTc
    {
        static constexpr int NCH = F / 8;                                   // 8-channel chunks
        static constexpr int N = 2 * F;                                     // w_hi columns | w_lo columns
        // F = 32: one CTA per SM (the B operand alone is 74 KB), 512 threads, all 512 TMEM columns.
        // F = 16: two CTAs per SM (93 KB each) with 256 threads and 256 TMEM columns, so that one CTA's tile load runs under the
        //         other's MMAs -- operand fetch and tile stores inside one CTA only slow each other down (measured).
        static constexpr int THREADS = F == 16 ? 256 : 512;
        static constexpr int CTAS_PER_SM = F == 16 ? 2 : 1;
        static constexpr int SLOTS = F == 16 ? 4 : 8;                       // M-tiles resident in TMEM
        static constexpr int TH = 15;                                       // output rows per CTA (15 * 66 = 7.7 M-tiles of 128)
        static constexpr int NT = (TH * WTC_PITCH + 127) / 128;             // M-tiles
        static constexpr int BT = 2;                                        // M-tiles per tcgen05.commit (a commit drains the tensor pipe)
        static constexpr int NPIX = NT * 128 + 2 * WTC_PITCH + 8;           // pixels per plane incl. the read slack of the last M-tile
        static constexpr int PLANE_BYTES = NPIX * 16;
        static constexpr int A_BYTES = NCH * 2 * PLANE_BYTES;
        static constexpr int B_BYTES_TAP = 2 * N * 16;                      // one (tap, chunk): [2 K chunks][N rows][8 fp16]
        static constexpr int B_BYTES = 9 * NCH * B_BYTES_TAP;
        static constexpr int OFF_B = A_BYTES;
        static constexpr int OFF_BAR = OFF_B + B_BYTES;
        static constexpr int SMEM_BYTES = OFF_BAR + (2 * SLOTS + NT) * 8 + 16;      // full[], empty[], band[NT] mbarriers, tmem slot
        static_assert((TH + 2) * WTC_PITCH <= NPIX, "input tile does not fit its plane");
        static_assert(SMEM_BYTES <= 232448, "tile exceeds the 227 KB shared-memory limit");
        static_assert(N <= WTC_SLOT_COLS, "accumulator does not fit a TMEM slot");
    };

    template<int F>
    struct WideTcParams
    {
        const float* in;
        float* out;
        const float* res;           // added after the activation (scale 1.0) or null
        const uint32_t* bop;        // this layer's packed B operand (WideTc<F>::B_BYTES)
        int w, h, act;
        float b[F];
        float a[F];
    };

    template<int F>
    __global__ void __launch_bounds__(WideTc<F>::THREADS, WideTc<F>::CTAS_PER_SM) wide_tc_kernel(const __grid_constant__ WideTcParams<F> prm)
    {
        using G = WideTc<F>;
        constexpr int WTC_SLOTS = G::SLOTS, WTC_THREADS = G::THREADS;
        constexpr int BT = G::BT, NBS = WTC_SLOTS / BT;          // tiles per commit, batches in flight
        constexpr int LOADERS = WTC_THREADS - 32;                // every warp but the issuing one
        constexpr uint32_t TMEM_COLS = WTC_SLOTS * WTC_SLOT_COLS;
        extern __shared__ __align__(1024) unsigned char wtc_smem[];
        uint64_t* bars = reinterpret_cast<uint64_t*>(wtc_smem + G::OFF_BAR);            // full[NBS], empty[NBS], band[NT]
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WTC_SLOTS + G::NT);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int x0 = blockIdx.x * WTC_TW, y0 = blockIdx.y * G::TH;

        if (threadIdx.x == 0)
        {
            for (int i = 0; i < NBS; i++)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(bars + i)));                    // full[batch]: one commit
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(tc_smem_u32(bars + WTC_SLOTS + i)));      // empty[batch]: 128 readers
            }
            for (int i = 0; i < G::NT; i++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tc_smem_u32(bars + 2 * WTC_SLOTS + i)), "r"(LOADERS));   // band[tile]
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0)
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = *tmem_slot;
        const uint32_t bar0 = tc_smem_u32(bars);
        // rows of the input tile that M-tile j needs in addition to the earlier tiles: [band_lo(j), band_hi(j))
        auto band_hi = [](int j) { return min(G::TH + 2, (j * 128 + 127) / WTC_PITCH + 3); };

        if (warp == 0)
        {
            // ---- MMA issue: tile j starts as soon as its rows have landed; one commit per BT tiles -------------------------------
            if (lane == 0)
            {
                const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(G::N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
                const uint32_t a_base = tc_smem_u32(wtc_smem), b_base = tc_smem_u32(wtc_smem + G::OFF_B);
                for (int j = 0; j < G::NT; j++)
                {
                    const int slot = j % WTC_SLOTS, batch = j / BT, bslot = batch % NBS, use = batch / NBS;
                    if (j % BT == 0 && use > 0) tc_mbar_wait(bar0 + (WTC_SLOTS + bslot) * 8, (use - 1) & 1);
                    tc_mbar_wait(bar0 + (2 * WTC_SLOTS + j) * 8, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    int first = 1;
#pragma unroll 1
                    for (int tap = 0; tap < 9; tap++)
                    {
                        const int off = j * 128 + (tap / 3) * WTC_PITCH + (tap % 3);        // flat pixel of row 0 of this tap's A operand
#pragma unroll
                        for (int c = 0; c < G::NCH; c++)
                        {
                            tc_mma(tmem + slot * WTC_SLOT_COLS,
                                   tc_desc(a_base + (2 * c) * G::PLANE_BYTES + off * 16, G::PLANE_BYTES, 128),
                                   tc_desc(b_base + (tap * G::NCH + c) * G::B_BYTES_TAP, G::N * 16, 128), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                    }
                    if (j % BT == BT - 1 || j == G::NT - 1) tc_commit(bar0 + bslot * 8);
                }
            }
            __syncwarp();
        }
        else
        {
            // ---- loaders (15 warps): B operand, then the input tile band by band ----------------------------------------------
            const int lt = threadIdx.x - 32;
            {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(prm.bop);
                const uint32_t dst = tc_smem_u32(wtc_smem + G::OFF_B);
                for (int i = lt * 16; i < G::B_BYTES; i += LOADERS * 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + i), "l"(src + i) : "memory");
            }
            // the read slack behind the tile only feeds dropped rows, but keep it finite
            for (int i = lt; i < (G::NPIX - (G::TH + 2) * WTC_PITCH) * G::NCH * 2; i += LOADERS)
            {
                const int plane = i % (G::NCH * 2), t = (G::TH + 2) * WTC_PITCH + i / (G::NCH * 2);
                reinterpret_cast<uint4*>(wtc_smem + plane * G::PLANE_BYTES)[t] = make_uint4(0u, 0u, 0u, 0u);
            }
            int row_lo = 0;
            for (int j = 0; j < G::NT; j++)
            {
                const int row_hi = band_hi(j);
                const int i0 = row_lo * WTC_PITCH * G::NCH, i1 = row_hi * WTC_PITCH * G::NCH;
                constexpr int BATCH = 4;
                // the global loads of a batch are all issued before the first one is consumed: one memory latency per batch
                for (int base = i0; base < i1; base += BATCH * LOADERS)
                {
                    float4 v0[BATCH], v1[BATCH];
#pragma unroll
                    for (int k = 0; k < BATCH; k++)
                    {
                        const int i = min(base + k * LOADERS + lt, i1 - 1);
                        const int c = i % G::NCH, t = i / G::NCH, tx = t % WTC_PITCH, ty = t / WTC_PITCH;
                        const int gx = clampi(x0 - 1 + tx, 0, prm.w - 1), gy = clampi(y0 - 1 + ty, 0, prm.h - 1);     // clamp-to-edge = replicate padding
                        const float4* p = reinterpret_cast<const float4*>(prm.in + (static_cast<size_t>(gy) * prm.w + gx) * F + c * 8);
                        v0[k] = __ldg(p); v1[k] = __ldg(p + 1);
                    }
#pragma unroll
                    for (int k = 0; k < BATCH; k++)
                    {
                        const int i = base + k * LOADERS + lt;
                        if (i >= i1) continue;
                        const int c = i % G::NCH, t = i / G::NCH;
                        const float v[8] = { v0[k].x, v0[k].y, v0[k].z, v0[k].w, v1[k].x, v1[k].y, v1[k].z, v1[k].w };
                        HalfPlanes pl{ reinterpret_cast<uint4*>(wtc_smem + (2 * c) * G::PLANE_BYTES), reinterpret_cast<uint4*>(wtc_smem + (2 * c + 1) * G::PLANE_BYTES) };
                        store_pixel_split(pl, t, v);
                    }
                }
                if (j == 0) asm volatile("cp.async.wait_all;" ::: "memory");       // the B operand is needed from the first MMA on
                tc_fence_async_smem();                                              // generic-proxy stores -> visible to the MMA's async-proxy reads
                tc_mbar_arrive(bar0 + (2 * WTC_SLOTS + j) * 8);
                row_lo = row_hi;
            }
            // ---- epilogue (warps 4-7, one per TMEM lane quadrant) ----------------------------------------------------------------
            if (warp >= 4 && warp < 8)
            {
                const int quad = warp & 3;
                for (int j = 0; j < G::NT; j++)
                {
                    const int slot = j % WTC_SLOTS, batch = j / BT, bslot = batch % NBS, use = batch / NBS;
                    const int q = j * 128 + quad * 32 + lane;
                    const int ty = q / WTC_PITCH, tx = q - ty * WTC_PITCH;
                    const int gx = x0 + tx, gy = y0 + ty;
                    const bool valid = tx < WTC_TW && ty < G::TH && gx < prm.w && gy < prm.h;
                    const size_t o = valid ? (static_cast<size_t>(gy) * prm.w + gx) * F : 0;
                    // the long skip's values are requested before the wait, so their latency hides behind the MMAs
                    float4 id[F / 4];
                    if (prm.res)
                    {
#pragma unroll
                        for (int c4 = 0; c4 < F / 4; c4++) id[c4] = __ldg(reinterpret_cast<const float4*>(prm.res + o) + c4);
                    }
                    if (j % BT == 0) tc_mbar_wait(bar0 + bslot * 8, use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t r[G::N / 16][16];
                    const uint32_t taddr = tmem + slot * WTC_SLOT_COLS + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
                    for (int i = 0; i < G::N / 16; i++) tc_ld16(r[i], taddr + 16 * i);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (j % BT == BT - 1 || j == G::NT - 1)
                    {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        tc_mbar_arrive(bar0 + (WTC_SLOTS + bslot) * 8);
                    }
                    if (!valid) continue;
#pragma unroll
                    for (int c4 = 0; c4 < F / 4; c4++)
                    {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const int c = c4 * 4 + e;
                            // column c: (a_hi + a_lo) w_hi, column F + c: a_hi w_lo
                            float sacc = __uint_as_float(r[c / 16][c % 16]) + __uint_as_float(r[(F + c) / 16][(F + c) % 16]);
                            sacc += prm.b[c];
                            if (prm.act == ACT_RELU) sacc = fmaxf(sacc, 0.0f);
                            else if (prm.act == ACT_PRELU) sacc = prelu(sacc, prm.a[c]);
                            v[e] = sacc;
                        }
                        if (prm.res) { v[0] += id[c4].x; v[1] += id[c4].y; v[2] += id[c4].z; v[3] += id[c4].w; }
                        reinterpret_cast<float4*>(prm.out + o)[c4] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
    }
}
