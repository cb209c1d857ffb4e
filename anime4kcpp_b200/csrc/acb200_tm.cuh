// TMEM-resident tensor engine of the fused luma network (tcgen05, sm_100a): activations AND accumulators live in tensor
// memory for a whole segment; shared memory holds only the weights and the luma tile.
//
// What the reference does per 2x pass (core/src/processor/cuda/Kernel.cu:246-514, CUDAProcessor.cpp:383-566): one launch per
// layer, fp16 activations through HBM.  What the round-1 engines do: 56x56 frames in shared memory, the 3x3 taps gathered by
// ldmatrix (mma.sync) or by SS-mode descriptors (tcgen05, 44 cycles per MMA: the A fetch from shared memory is the bound).
// Measured facts this engine is built on (tools/microbench_tcgen05_ts.cu, profiles/r02_microbench_tcgen05_ts.txt):
//   * tcgen05.mma with the A operand in TMEM (M = 128, K = 16) costs N / 2 cycles: 24 cycles for N = 48, no operand fetch.
//   * the .ashift qualifier (multiply, then shift the A rows by one lane inside every 32-lane quadrant: new[r] = old[r + 1])
//     is free.  A horizontal tap is therefore a lane shift of the operand -- no data movement by any thread.
//   * a vertical tap is a different A operand (another 8-column group of TMEM) -- free as well.
//
// Layout.  A CTA owns four independent STRIPS, one per TMEM lane quadrant: 32 pixels wide (lane = x) and G <= 48 rows tall.
// Row y of the current layer's 8-channel map is ONE A operand: columns 8y .. 8y+7 of every lane hold the pixel's eight
// channels as split fp16 (hi words 0-3, lo words 4-7), i.e. K = 16 = {a_hi | a_lo}.  Per row r of the input map the issuing
// thread runs three alignments (dx = -1, 0, +1 through two .ashift's), each ONE MMA with N = 48 = 3 output rows x (8 couts
// with w_hi | 8 couts with w_lo): row r contributes to output rows r+1, r, r-1 with the weights of dy = -1, 0, +1, whose
// accumulators are adjacent 16-column slots of a ring of eight.  All nine taps and all three split-precision products
// (a_hi w_hi + a_lo w_hi + a_hi w_lo) are accumulated inside the tensor core: 72 tensor cycles per 128 pixels and layer.
//   The accumulate flag is always set: the epilogue that drains a slot writes the NEXT user's bias back into it (one
// tcgen05.st), which also makes the bias add free.
//   Epilogue (warps 4.., one warp per lane quadrant and row): tcgen05.ld of 16 columns, hi + lo, activation, fp16 split,
// tcgen05.st of the 8 operand columns of the next layer IN PLACE (row y of layer l overwrites row y of layer l - 1, which is
// dead once row y + 1 has been multiplied).  Nothing is exchanged between threads and no shared memory is touched.
//   The lane shift drifts the map by one lane per layer (lane j of layer l is pixel x0 + l + j), which consumes exactly the
// halo a fused segment loses anyway: after R layers lanes 0 .. 31 - 2R hold the strip's 32 - 2R output columns.
//   Layers are not separated by barriers: rows flow (mbarrier per row / per accumulator slot), so the tensor pipe runs
// across layer boundaries and under the head conv / the map load of the first rows.
//   Replicate padding (border CTAs): the epilogue copies the edge values one pixel outwards -- a warp shuffle in x, a second
// tcgen05.st in y.  Rows outside the image are skipped.
#pragma once

#include <cuda_fp16.h>

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"
#include "acb200_mma.cuh"

namespace acb
{
#ifndef ACB_TM_EPI_SETS
#define ACB_TM_EPI_SETS 4
#endif
    // TMEM budget: 8 columns per frame row (the A operands) + 8 columns per accumulator slot = 512: G + 4 * groups <= 64.  The host
    // picks the frame height G and with it the ring depth (groups of four slots): a deeper ring lets the issuers run further ahead of
    // the epilogue, a taller frame wastes less on the vertical halo.
    constexpr int TM_GMAX = 48;                     // rows of a strip frame (ring of 4 groups); 40 rows -> 6 groups
    constexpr int TM_MAX_GROUPS = 12 * 8;           // one-shot `full` barriers: groups of a segment (<= 12 per layer)
#ifndef ACB_TM_ISSUERS
#define ACB_TM_ISSUERS 4
#endif
    constexpr int TM_ISSUERS_DECL = ACB_TM_ISSUERS;
    constexpr int TM_SETS = ACB_TM_EPI_SETS;        // epilogue warp sets (4 warps = 4 lane quadrants each)
    constexpr int TM_THREADS = 32 * TM_ISSUERS_DECL + 128 * TM_SETS;
    constexpr int TM_MAX_R = 8;
#ifndef ACB_TM_ISSUERS
#define ACB_TM_ISSUERS 4
#endif
    constexpr int TM_ISSUERS = ACB_TM_ISSUERS;      // issuer warps (warps 0 .. TM_ISSUERS - 1), each takes every TM_ISSUERS-th chunk of the step program
    constexpr int TM_EPI_WARP0 = TM_ISSUERS;        // first epilogue warp (a multiple of 4: warp w works on TMEM lane quadrant w % 4)
    constexpr int TM_CHUNK = 4;                     // consecutive steps (input rows) per chunk, >= 3
    constexpr int TM_B_BYTES_HALF = 2 * 24 * 16;    // one B matrix: [2 K chunks][24 rows][8 fp16]
    constexpr int TM_B_BYTES_AL = 2 * TM_B_BYTES_HALF;      // one alignment: the w_hi matrix, then the w_lo matrix
    constexpr int TM_B_BYTES_LAYER = 3 * TM_B_BYTES_AL;     // 4608
    constexpr int TM_B_WORDS_LAYER = TM_B_BYTES_LAYER / 4;
    constexpr int TM_LP = 34;                       // luma tile pitch (floats): lanes 0..31 read columns j .. j + 2
    constexpr int TM_OFF_LUMA = TM_MAX_R * TM_B_BYTES_LAYER;
    constexpr int TM_OFF_BAR = TM_OFF_LUMA + 4 * (TM_GMAX + 2) * TM_LP * 4;
    constexpr int TM_N_BARS = TM_GMAX + TM_MAX_GROUPS + 16 + 1;
    constexpr int TM_OFF_GEOM = TM_OFF_BAR + TM_N_BARS * 8 + 8;
    constexpr int TM_OFF_STEPS = ((TM_OFF_GEOM + 3 * (TM_MAX_R + 2) * 4 + 4 + 15) / 16) * 16;      // the issuers' step program, 64 bytes per input row and layer
    constexpr int TM_MAX_STEPS = TM_MAX_R * TM_GMAX;
#ifdef ACB_TM_TRACE
    constexpr int TM_OFF_TRACE = TM_OFF_STEPS + TM_MAX_STEPS * 64;
    constexpr int TM_SMEM_BYTES = TM_OFF_TRACE + 5 * TM_MAX_STEPS * 8;
#else
    constexpr int TM_SMEM_BYTES = TM_OFF_STEPS + TM_MAX_STEPS * 64;
#endif

    template<class S>
    struct TmParams
    {
        const void* src;
        const uint4* map_in;    // previous segment's map, two planes (hi, lo) of [h][w][8 x fp16] (the mma engine's format)
        uint4* map_out;
        const float* feat_in;
        float* feat_out;
        void* dst;
        int src_pitch, dst_pitch;
        int w, h;
        int type;
        int tiles_x, strips_x, G;
        int ring_groups;        // accumulator ring depth in groups of four 8-column slots: 8 G + 32 ring_groups <= 512
        int issuers;            // issuer warps used (1 .. TM_ISSUERS)
        const uint32_t* bops;   // B operands of this segment's 3x3 convs, TM_B_WORDS_LAYER words each, in layer order
        float k[(S::HEAD ? 72 : 0) + 64 + 32];  // fp32 weights used outside the MMAs: head (72) | ARNet 1x1 (64) | legacy deconv (32)
        float b[S::NB];
        float a[S::NA > 0 ? S::NA : 1];
    };

    __device__ __forceinline__ bool tm_elect_one()
    {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
        return pred != 0;
    }
    __device__ __forceinline__ void tm_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc)
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    __device__ __forceinline__ void tm_mma_ashift(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc)
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.ashift [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    __device__ __forceinline__ void tm_commit(uint32_t bar)
    {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    __device__ __forceinline__ void tm_wait(uint32_t bar, uint32_t parity)
    {
        uint32_t ok = 0, spins = 0;
        do
        {
            // (the suspend-time hint keeps a waiting warp asleep instead of spinning through the sub-partition's issue slots)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
            if (!ok && ++spins > (1u << 20)) __trap();      // a protocol bug must fault, never hang the device
        } while (!ok);
    }
    __device__ __forceinline__ void tm_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
    __device__ __forceinline__ void tm_ld16(uint32_t (&v)[16], uint32_t taddr)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    }
    __device__ __forceinline__ void tm_ld8(uint32_t (&v)[8], uint32_t taddr)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    }
    __device__ __forceinline__ void tm_ld32(uint32_t (&v)[32], uint32_t taddr)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                       "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
    }
    __device__ __forceinline__ void tm_st32(uint32_t taddr, const uint32_t (&v)[32])
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
                     :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr) : "memory");
    }
    __device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8])
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                     :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr) : "memory");
    }
    __device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16])
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                     :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr) : "memory");
    }
#define ACB_TM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define ACB_TM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define ACB_TM_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define ACB_TM_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

    // rows of layer l's output that exist: the frame shrunk by l, clipped to the image (frame row y = image row y0 + y)
    __device__ __forceinline__ void tm_rows(int l, int G, int y0, int h, int& ya, int& yb)
    {
        ya = max(l, -y0);
        yb = min(G - 1 - l, h - 1 - y0);
    }

    // Progress words: one byte per lane quadrant (= per epilogue warp working on the row / slot group), holding a MONOTONIC count.
    // (mbarrier parity waits cannot be used where a waiter may be two phases ahead of the barrier -- several issuer warps run layers
    // apart on the same row index -- because a phase parity only distinguishes adjacent phases.)
    __device__ __forceinline__ void tm_wait_bytes(uint32_t addr, uint32_t need)
    {
        // every byte of the word >= need (all values < 128): ((b | 0x80) - need) keeps its top bit exactly when b >= need, and no
        // byte borrows from its neighbour
        const uint32_t want = need * 0x01010101u;
        uint32_t spins = 0;
        for (;;)
        {
            uint32_t w;
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr) : "memory");
            if ((((w | 0x80808080u) - want) & 0x80808080u) == 0x80808080u) return;
            if (++spins > (1u << 24)) __trap();     // a protocol bug must fault, never hang the device
        }
    }
    // publish: plain byte stores.  What they publish are TMEM writes whose COMPLETION the publishing thread has already waited for
    // (tcgen05.wait::st / wait::ld, then tcgen05.fence::before_thread_sync), so no membar is needed in front of the flag store.
    __device__ __forceinline__ void tm_publish_byte(uint32_t addr, uint32_t value)
    {
        asm volatile("st.volatile.shared.u8 [%0], %1;" :: "r"(addr), "r"(value) : "memory");
    }

    // Row bookkeeping.  The rows a layer produces are handled in GROUPS of four consecutive rows (the last group of a layer may be
    // shorter): a group is one epilogue work item, one `full` mbarrier phase and one drain count.  Rows are numbered densely across
    // layers with every layer padded to whole groups: row y of layer l has index T = tb[l] + (y - ya[l]), accumulator slot T % 16,
    // group T / 4, slot group (T / 4) % 4.
    template<class S>
    __global__ void __launch_bounds__(TM_THREADS, 1) segment_tm_kernel(const __grid_constant__ TmParams<S> prm)
    {
        constexpr int R = S::R;                 // 3x3 convs of this segment (all on the tensor cores)
        constexpr int SW = 32 - 2 * R;          // output columns of a strip
        static_assert(R >= 1 && R <= TM_MAX_R && SW >= 8, "segment too deep for 32-pixel strips");
        static_assert(S::FAM == ACB200_FAMILY_ACNET_LEGACY || S::FAM == ACB200_FAMILY_ACNET, "family not on this engine yet");
        static_assert(TM_SETS == 4, "groups are dealt to the epilogue sets by group index % 4");
        extern __shared__ __align__(128) unsigned char smem_tm[];
        float* luma_all = reinterpret_cast<float*>(smem_tm + TM_OFF_LUMA);
        const uint32_t bars = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm + TM_OFF_BAR));
        // flag_a[row], flag_e[slot group]: progress words (see tm_wait_bytes); bar_full[group] (one-shot: every group of every layer has its
        // own mbarrier, so no phase parity can alias), bar_bop: mbarriers
        const uint32_t flag_a = bars, bar_full = bars + 8 * TM_GMAX, flag_e = bar_full + 8 * TM_MAX_GROUPS, bar_bop = flag_e + 8 * 16;
        const int NRG = prm.ring_groups, NRS = 4 * NRG;                 // accumulator ring: groups / slots
        const uint32_t d_col0 = 512u - 32u * static_cast<uint32_t>(NRG);    // ... at the top of the TMEM columns
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_tm + TM_OFF_BAR + TM_N_BARS * 8);
        int* s_ya = reinterpret_cast<int*>(smem_tm + TM_OFF_GEOM);      // per layer 0..R+1: first / last existing row, padded dense index of the first
        int* s_yb = s_ya + TM_MAX_R + 2;
        int* s_tb = s_yb + TM_MAX_R + 2;
        int* s_nsteps = s_tb + TM_MAX_R + 2;
        uint4* steps = reinterpret_cast<uint4*>(smem_tm + TM_OFF_STEPS);
#ifdef ACB_TM_TRACE
        long long* trace = reinterpret_cast<long long*>(smem_tm + TM_OFF_TRACE);    // per chunk: [0] start, [1] issued; per group: [2] full seen, [3] drained, [4] rows published
        const bool traced = blockIdx.x == gridDim.x / 2 + 3;
#endif
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int G = prm.G, SH = G - 2 * R;
        const int tile_x = blockIdx.x % prm.tiles_x, sy = blockIdx.x / prm.tiles_x;
        const int y0 = sy * SH - R;

        // ---- setup ---------------------------------------------------------------------------------------------------------------
        if (threadIdx.x == 0)
        {
            for (int i = 0; i < TM_GMAX; i++) asm volatile("st.shared.u32 [%0], %1;" :: "r"(flag_a + 8 * i), "r"(0));
            for (int i = 0; i < 16; i++) asm volatile("st.shared.u32 [%0], %1;" :: "r"(flag_e + 8 * i), "r"(0));
            {
                // one `full` barrier per group: two commits per row
                int gi = 0;
                for (int l = 1; l <= R; l++)
                {
                    int ya, yb;
                    tm_rows(l, G, y0, prm.h, ya, yb);
                    for (int y = ya; y <= yb; y += 4, gi++)
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_full + 8 * gi), "r"(2 * min(4, yb - y + 1)));
                }
            }
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_bop));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // the segment's B operands: one bulk copy (UBLKCP) on an mbarrier
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_bop), "r"(R * TM_B_BYTES_LAYER) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm))), "l"(reinterpret_cast<uint64_t>(prm.bops)), "r"(R * TM_B_BYTES_LAYER), "r"(bar_bop) : "memory");
            int tb = 0, ns = 0;
            for (int l = 0; l <= R; l++)
            {
                int ya, yb;
                tm_rows(l, G, y0, prm.h, ya, yb);
                s_ya[l] = ya; s_yb[l] = yb; s_tb[l] = tb;
                if (l >= 1 && yb >= ya) { tb += 4 * ((yb - ya + 4) >> 2); ns += yb - ya + 3; }
            }
            s_ya[R + 1] = 0; s_yb[R + 1] = -1; s_tb[R + 1] = tb;
            *s_nsteps = ns;
        }
        if (warp == 0)
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(tmem_slot))), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        if constexpr (S::NEEDS_LUMA)
        {
            // luma tile of every strip: frame columns -1 .. 32, frame rows -1 .. G, clamp-to-edge (the head conv's own padding)
            const int n = 4 * (G + 2) * TM_LP;
            for (int i = threadIdx.x; i < n; i += TM_THREADS)
            {
                const int q = i / ((G + 2) * TM_LP), rem = i - q * (G + 2) * TM_LP, ly = rem / TM_LP, lx = rem - ly * TM_LP;
                const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
                const int gx = clampi(strip * SW - R - 1 + lx, 0, prm.w - 1), gy = clampi(y0 - 1 + ly, 0, prm.h - 1);
                luma_all[(q * (TM_GMAX + 2) + ly) * TM_LP + lx] = load_elem(static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch, gx, prm.type);
            }
        }
        ACB_TM_FENCE_BEFORE();
        __syncthreads();
        ACB_TM_FENCE_AFTER();
        const uint32_t tmem = *tmem_slot;
        constexpr int B0 = S::HEAD ? 8 : 0;     // bias / alpha offsets of the segment's first 3x3 conv inside prm.b / prm.a
        constexpr int A0 = (S::FAM == ACB200_FAMILY_ACNET && S::HEAD) ? 8 : 0;
        // bias of the segment's tensor layer ln (1-based) as the initial accumulator of output channel c (the ACNet tail has 4 couts)
        auto bias_of = [&](const int ln, const int c) -> float {
            if (S::TAIL && S::FAM == ACB200_FAMILY_ACNET && ln == R && c >= 4) return 0.0f;
            return prm.b[B0 + 8 * (ln - 1) + c];
        };
        {
            // The issuers' step program: one record of four uint4 per (layer, input row), built by all threads (one step each), with every operand the
            // issue loop needs already in its final form:
            //   [0] up to four waits (progress word | count << 24; 0 = none): the A row, then the slot groups this chunk touches first
            //   [1] A operand, first run: accumulator / B descriptor low word / instruction descriptor
            //   [2] second run (only when the accumulator ring wraps inside this row's outputs): accumulator (0 = none) / B / instruction descriptor; layer
            //   [3] up to three `full` barriers to commit to (barrier | number of commits << 24; 0 = none)
            // Steps are issued in CHUNKS of TM_CHUNK consecutive steps, chunk c by issuer warp c % issuers: every UTCHMMA holds a
            // scoreboard on its uniform-register operands until the tensor core dequeues it, so a single issuing thread can never run
            // ahead of the pipe and each of its waits becomes a bubble (120-150 cycles, profiles/r02_microbench_tcgen05_ts.txt).
            // With several issuers one warp's waits run under the other warps' queued MMAs.  All MMAs accumulate (the epilogue
            // re-initialises drained accumulators), so their order across rows does not matter; an output row's three input rows lie
            // in at most two chunks, hence every row contributes exactly two commits to its group's `full` barrier.
            const uint32_t bop_s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm));
            constexpr uint32_t IDESC0 = (1u << 4) | (static_cast<uint32_t>(128 >> 4) << 24);       // D f32, A / B f16 K-major, M = 128
            constexpr uint32_t DESC_HI = ((384u >> 4) & 0x3FFF) << 16;                              // LBO (K chunk distance: 24 rows) in the low word
            int base = 0;
            for (int l = 1; l <= R; l++)
            {
                const int ya = s_ya[l], yb = s_yb[l], tb = s_tb[l];
                if (yb < ya) continue;
                const int n = yb - ya + 3;
                for (int k = threadIdx.x; k < n; k += TM_THREADS)
                {
                    const int i = base + k, r = ya - 1 + k;
                    const int oa = max(r - 1, ya), ob = min(r + 1, yb), n_rows = ob - oa + 1;
                    const uint32_t ta = static_cast<uint32_t>(tb + oa - ya);
                    const int s0 = static_cast<int>(ta % static_cast<uint32_t>(NRS)), first = min(n_rows, NRS - s0), jb0 = oa - (r - 1);
                    const int c0 = (i / TM_CHUNK) * TM_CHUNK, c1 = c0 + TM_CHUNK - 1;       // this step's chunk
                    const uint32_t w0 = (flag_a + 8 * r) | (static_cast<uint32_t>(l) << 24);      // layer l - 1 publishes l
                    uint32_t w1 = 0u, w2 = 0u, cv0 = 0u, cv1 = 0u, cv2 = 0u, cv3 = 0u;
                    int last_group = -1;
#pragma unroll
                    for (int jj = 0; jj < 3; jj++)
                    {
                        if (jj >= n_rows) break;
                        const int o = oa + jj, rel = o - ya;
                        const uint32_t t = static_cast<uint32_t>(tb + rel), gidx = t >> 2;
                        // the group's rows are touched by this layer's steps 4j .. 4j + 5 (j = rel / 4); the first of them inside this chunk waits
                        // for the slot group's previous user to have been drained
                        if (static_cast<int>(gidx) != last_group)
                        {
                            last_group = static_cast<int>(gidx);
                            if (i == max(base + (rel & ~3), c0) && gidx >= static_cast<uint32_t>(NRG))
                            {
                                const uint32_t wq = (flag_e + 8 * (gidx % NRG)) | ((gidx / NRG) << 24);
                                if (w1 == 0u) w1 = wq; else w2 = wq;
                            }
                        }
                        const int i_a = base + rel, i_c = i_a + 2;                          // steps of input rows o - 1 and o + 1
                        if (i == min(i_c, c1))                                              // this chunk's last contribution to row o
                        {
                            const uint32_t m = (i_a >= c0 && i_c <= c1) ? 2u : 1u;
                            const uint32_t cq = (bar_full + 8 * gidx) | (m << 24);
                            if (jj == 0) cv0 = cq; else if (jj == 1) cv1 = cq; else cv2 = cq;
                        }
                    }
                    const uint32_t b1 = (((bop_s + (l - 1) * TM_B_BYTES_LAYER + jb0 * 128) >> 4) & 0x3FFF) | DESC_HI;
                    uint4 m1, m2;
                    m1.x = tmem + 8 * r;
                    m1.y = tmem + d_col0 + 8 * s0;
                    m1.z = b1;
                    m1.w = IDESC0 | (static_cast<uint32_t>(first) << 17);
                    m2.x = first < n_rows ? tmem + d_col0 : 0u;
                    m2.y = b1 + first * 8;
                    m2.z = IDESC0 | (static_cast<uint32_t>(n_rows - first) << 17);
                    m2.w = static_cast<uint32_t>(l);
                    // A chunk is REGULAR when its four steps are interior rows of one layer (three output rows each, weights at row block
                    // 0) and none of them wraps around the accumulator ring: the issue loop then derives every operand from the leading
                    // step's record by uniform arithmetic.  The leading record carries the flag and the dense index of the first touched row.
                    if (i == c0)
                    {
                        bool regular = k >= 2 && k + (TM_CHUNK - 1) <= n - 3;
                        for (int pp = 0; pp < TM_CHUNK && regular; pp++) regular = static_cast<int>((ta + pp) % static_cast<uint32_t>(NRS)) + 3 <= NRS;
                        cv3 = (regular ? 0x80000000u : 0u) | ta;
                    }
                    steps[4 * i] = make_uint4(w0, w1, w2, 0u);
                    steps[4 * i + 1] = m1;
                    steps[4 * i + 2] = m2;
                    steps[4 * i + 3] = make_uint4(cv0, cv1, cv2, cv3);
                }
                base += n;
            }
        }
        if (warp >= TM_EPI_WARP0 && warp < TM_EPI_WARP0 + 4)
        {
            // accumulator ring: every slot starts with its first user's bias
            // (slot s is first used by dense row index s, which belongs to a later layer when the frame has few rows)
            const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
            for (int s = 0; s < NRS; s++)
            {
                int ln = 1;
                while (ln <= R && s >= s_tb[ln + 1]) ln++;
                if (ln > R) break;
                uint32_t init[8];
#pragma unroll
                for (int c = 0; c < 8; c++) init[c] = __float_as_uint(bias_of(ln, c));
                tm_st8(tmem + lane_base + d_col0 + 8 * s, init);
            }
            ACB_TM_WAIT_ST();
        }
        ACB_TM_FENCE_BEFORE();
        __syncthreads();
        ACB_TM_FENCE_AFTER();

        if (warp < prm.issuers)
        {
            // ==== MMA issuers: warp w takes chunks w, w + issuers, ... of the step program =============================================
            tm_wait(bar_bop, 0);
            constexpr uint64_t DESC_TOP = static_cast<uint64_t>(((128u >> 4) & 0x3FFF) | (1u << 14)) << 32;     // SBO | descriptor version, high word
            const int nsteps = *s_nsteps;
            const uint32_t* wait_words = reinterpret_cast<const uint32_t*>(steps);
#pragma unroll 1
            for (int c = warp * TM_CHUNK; c < nsteps; c += prm.issuers * TM_CHUNK)
            {
                // A chunk that lies inside one layer polls all of its waits at once, one per lane (a chunk that spans two layers of a very
                // short frame can depend on its own earlier steps: it waits step by step).
                const int i_last = min(c + TM_CHUNK, nsteps) - 1;
#ifdef ACB_TM_TRACE
                if (lane == 0) trace[c] = clock64();
#endif
                const bool batched = steps[4 * c + 2].w == steps[4 * i_last + 2].w;
                if (batched)
                {
                    const int i = c + (lane >> 2);
                    const uint32_t w = (lane < 4 * TM_CHUNK && i <= i_last) ? wait_words[16 * i + (lane & 3)] : 0u;
                    if (w) tm_wait_bytes(w & 0xffffffu, w >> 24);
                    __syncwarp();
                }
                if (tm_elect_one())
                {
#ifdef ACB_TM_TRACE
                    trace[TM_MAX_STEPS + c] = clock64();
#endif
                    ACB_TM_FENCE_AFTER();
                    const uint4 lead1 = steps[4 * c + 1];
                    const uint32_t lead = steps[4 * c + 3].w;
                    if (batched && (lead & 0x80000000u))
                    {
                        // regular chunk: 24 MMAs with operands in uniform registers.  Step p multiplies A row r0 + p into the accumulators of
                        // output rows T0 + p .. T0 + p + 2 (consecutive slots); the commits follow the fixed pattern of a chunk: row T0 + p
                        // gets this chunk's last contribution at step p (both of its commits when all three input rows are in the chunk),
                        // rows T0 + 4 and T0 + 5 at the last step.
                        const uint32_t T0 = lead & 0x7fffffffu;
                        const uint32_t bl = lead1.z, idesc = lead1.w;
#pragma unroll
                        for (int pp = 0; pp < TM_CHUNK; pp++)
                        {
                            const uint32_t a = lead1.x + 8 * pp, dd = lead1.y + 8 * pp;
#pragma unroll
                            for (int al = 0; al < 3; al++)
                            {
                                const uint32_t o_hi = al * (TM_B_BYTES_AL >> 4), o_lo = o_hi + (TM_B_BYTES_HALF >> 4);
                                tm_mma(dd, a, DESC_TOP | (bl + o_hi), idesc);
                                if (al < 2) tm_mma_ashift(dd, a, DESC_TOP | (bl + o_lo), idesc); else tm_mma(dd, a, DESC_TOP | (bl + o_lo), idesc);
                            }
                            const uint32_t b0 = bar_full + 8 * ((T0 + pp) >> 2);
                            tm_commit(b0);
                            if (pp >= 2) tm_commit(b0);
                            if (pp == TM_CHUNK - 1)
                            {
                                tm_commit(bar_full + 8 * ((T0 + pp + 1) >> 2));
                                tm_commit(bar_full + 8 * ((T0 + pp + 2) >> 2));
                            }
                        }
                    }
                    else
                    {
#pragma unroll 1
                    for (int i = c; i <= i_last; i++)
                    {
                        const uint4 m1 = steps[4 * i + 1], m2 = steps[4 * i + 2], cv = steps[4 * i + 3];
                        if (!batched)
                        {
                            const uint4 w = steps[4 * i];
                            tm_wait_bytes(w.x & 0xffffffu, w.x >> 24);
                            if (w.y) tm_wait_bytes(w.y & 0xffffffu, w.y >> 24);
                            if (w.z) tm_wait_bytes(w.z & 0xffffffu, w.z >> 24);
                            if (w.w) tm_wait_bytes(w.w & 0xffffffu, w.w >> 24);
                            ACB_TM_FENCE_AFTER();
                        }
                        // per alignment: the w_hi matrix ((a_hi + a_lo) w_hi), then the w_lo matrix (a_hi w_lo) into the SAME 8 columns per
                        // output row; the operand shift rides on the alignment's last MMA
#pragma unroll
                        for (int al = 0; al < 3; al++)
                        {
                            const uint32_t o_hi = al * (TM_B_BYTES_AL >> 4), o_lo = o_hi + (TM_B_BYTES_HALF >> 4);
                            if (m2.x == 0u)
                            {
                                tm_mma(m1.y, m1.x, DESC_TOP | (m1.z + o_hi), m1.w);
                                if (al < 2) tm_mma_ashift(m1.y, m1.x, DESC_TOP | (m1.z + o_lo), m1.w); else tm_mma(m1.y, m1.x, DESC_TOP | (m1.z + o_lo), m1.w);
                            }
                            else
                            {
                                tm_mma(m1.y, m1.x, DESC_TOP | (m1.z + o_hi), m1.w);
                                tm_mma(m1.y, m1.x, DESC_TOP | (m1.z + o_lo), m1.w);
                                tm_mma(m2.x, m1.x, DESC_TOP | (m2.y + o_hi), m2.z);
                                if (al < 2) tm_mma_ashift(m2.x, m1.x, DESC_TOP | (m2.y + o_lo), m2.z); else tm_mma(m2.x, m1.x, DESC_TOP | (m2.y + o_lo), m2.z);
                            }
                        }
                        // output rows that have received this chunk's last contribution
                        for (uint32_t m = cv.x >> 24; m > 0; m--) tm_commit(cv.x & 0xffffffu);
                        for (uint32_t m = cv.y >> 24; m > 0; m--) tm_commit(cv.y & 0xffffffu);
                        for (uint32_t m = cv.z >> 24; m > 0; m--) tm_commit(cv.z & 0xffffffu);
                    }
                    }
                }
                __syncwarp();
            }
        }
        else if (warp >= TM_EPI_WARP0)
        {
            // ==== producers of layer 0 and epilogue of every layer: one warp per lane quadrant (strip) and group of four rows ===============
            const int set = (warp - TM_EPI_WARP0) >> 2, q = warp & 3;
            const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
            const int x0 = strip * SW - R;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            const float* luma = luma_all + q * (TM_GMAX + 2) * TM_LP;
            const int pad_top = -y0 - 1, pad_bot = prm.h - y0;          // frame rows of image rows -1 and h (replicate padding), if inside the frame
            const uint32_t my_a = tmem + lane_base;
            const uint32_t my_flag_a = flag_a + q;
            const bool pads = (pad_top >= 0) || (pad_bot <= G - 1);     // the frame reaches over the top / bottom image edge

            // one row of layer l's output map becomes the next layer's A operand (x-clamped in border strips); its padding copies in y
            auto put_row = [&](const int l, const int y, uint32_t (&w8)[8]) {
                const int L0 = -x0 - l, L1 = prm.w - 1 - x0 - l;       // lanes of image columns 0 and w - 1 in layer l's map
                if (L0 > 0 || L1 < 31)
                {
                    const int srcl = min(max(lane, L0), L1);
#pragma unroll
                    for (int c = 0; c < 8; c++) w8[c] = __shfl_sync(0xffffffffu, w8[c], srcl);
                }
                tm_st8(my_a + 8 * y, w8);
                if (pads)
                {
                    if (y == pad_top + 1 && pad_top >= 0) tm_st8(my_a + 8 * pad_top, w8);
                    if (y == pad_bot - 1 && pad_bot <= G - 1) tm_st8(my_a + 8 * pad_bot, w8);
                }
            };
            // ... and is published (after tcgen05.wait::st): every lane stores the progress byte (same address, same value: one store, no branch)
            auto publish_rows = [&](const int l, const int ya_, const int k) {
                for (int j = 0; j < k; j++) tm_publish_byte(my_flag_a + 8 * (ya_ + j), l + 1);
                if (pads)
                {
                    if (ya_ == pad_top + 1 && pad_top >= 0) tm_publish_byte(my_flag_a + 8 * pad_top, l + 1);
                    if (ya_ + k - 1 == pad_bot - 1 && pad_bot <= G - 1) tm_publish_byte(my_flag_a + 8 * pad_bot, l + 1);
                }
            };

            // ---- layer 0: the head conv (fp32 FFMA) or the previous segment's map --------------------------------------------------------
            {
                const int ya = s_ya[0], yb = s_yb[0];
                if constexpr (S::HEAD)
                {
                    constexpr int ACT = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? ACT_RELU : S::FAM == ACB200_FAMILY_ACNET ? ACT_PRELU : ACT_IDENTITY;
                    for (int yg = ya + 4 * set; yg <= yb; yg += 4 * TM_SETS)
                    {
                        const int k = min(4, yb - yg + 1);
                        for (int j = 0; j < k; j++)
                        {
                            const int y = yg + j;
                            float r9[9];
#pragma unroll
                            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                                for (int dx = 0; dx < 3; dx++) r9[dy * 3 + dx] = luma[(y + dy) * TM_LP + lane + dx];
                            float v[8];
#pragma unroll
                            for (int co = 0; co < 8; co++)
                            {
                                float s = prm.b[co];
#pragma unroll
                                for (int p = 0; p < 9; p++) s = fmaf(r9[p], prm.k[co * 9 + p], s);
                                if (ACT == ACT_RELU) s = fmaxf(s, 0.0f);
                                else if (ACT == ACT_PRELU) s = prelu(s, prm.a[co]);
                                v[co] = s;
                            }
                            uint32_t w8[8];
                            split_pair(v[0], v[1], w8[0], w8[4]); split_pair(v[2], v[3], w8[1], w8[5]);
                            split_pair(v[4], v[5], w8[2], w8[6]); split_pair(v[6], v[7], w8[3], w8[7]);
                            put_row(0, y, w8);
                        }
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        publish_rows(0, yg, k);
                    }
                }
                else
                {
                    // the map is read with clamped coordinates: padding included, every frame row inside the image +- 1 exists.  Four rows
                    // (eight 16-byte loads) are requested before the first is stored, and published together.
                    const int la = max(ya - 1, 0), lb = min(yb + 1, G - 1);
                    const int gx = clampi(x0 + lane, 0, prm.w - 1);
                    const size_t plane = static_cast<size_t>(prm.w) * prm.h;
                    for (int yg = la + 4 * set; yg <= lb; yg += 4 * TM_SETS)
                    {
                        uint4 hi[4], lo[4];
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            const int gy = clampi(y0 + min(yg + k, lb), 0, prm.h - 1);
                            hi[k] = __ldg(prm.map_in + static_cast<size_t>(gy) * prm.w + gx);
                            lo[k] = __ldg(prm.map_in + plane + static_cast<size_t>(gy) * prm.w + gx);
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            const uint32_t w8[8] = { hi[k].x, hi[k].y, hi[k].z, hi[k].w, lo[k].x, lo[k].y, lo[k].z, lo[k].w };
                            if (yg + k <= lb) tm_st8(my_a + 8 * (yg + k), w8);
                        }
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
#pragma unroll
                        for (int k = 0; k < 4; k++) if (yg + k <= lb) tm_publish_byte(my_flag_a + 8 * (yg + k), 1);
                    }
                }
            }

            // ---- layers 1 .. R ---------------------------------------------------------------------------------------------------------------
            const int es = prm.type & 0xff;
            const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
            const uint32_t my_flag_e = flag_e + q;
            (void)aligned;
#pragma unroll 1
            for (int l = 1; l <= R; l++)
            {
                const int ya = s_ya[l], yb = s_yb[l], tb = s_tb[l], t_end = s_tb[l + 1], t_end2 = l < R ? s_tb[l + 2] : 0;
                const bool last = l == R;
                const bool xclamp = (-x0 - l > 0) || (prm.w - 1 - x0 - l < 31);      // the strip reaches over the left / right image edge in this layer's map
                // accumulator re-initialisation blocks: this layer's bias and the next layer's (the slot group's next user is one ring ahead)
                uint32_t init_c[8], init_n[8];
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    init_c[c] = __float_as_uint(bias_of(l, c));
                    init_n[c] = __float_as_uint(bias_of(min(l + 1, R), c));
                }
                float alpha[8];
#pragma unroll
                for (int c = 0; c < 8; c++) alpha[c] = S::FAM == ACB200_FAMILY_ACNET ? prm.a[A0 + 8 * (min(l, S::NCONV) - 1) + c] : 0.0f;
                // this set's groups: group index gidx = tb / 4 + j with gidx % 4 == set
                const int g0 = tb >> 2;
                for (int j = (set - g0) & 3; ya + 4 * j <= yb; j += 4)
                {
                    const int yg = ya + 4 * j, k = min(4, yb - yg + 1);
                    const uint32_t gidx = static_cast<uint32_t>(g0 + j), use = gidx / static_cast<uint32_t>(NRG), sg = gidx - use * NRG;
                    const uint32_t d_addr = my_a + d_col0 + 32 * sg;
                    tm_wait(bar_full + 8 * gidx, 0);
                    ACB_TM_FENCE_AFTER();
#ifdef ACB_TM_TRACE
                    if (q == 0 && lane == 0) trace[2 * TM_MAX_STEPS + gidx] = clock64();
#endif
                    uint32_t d[32];
                    tm_ld32(d, d_addr);
                    ACB_TM_WAIT_LD();
                    {
                        // hand the slot group to its next user (group gidx + ring depth) with that layer's bias
                        const int tn = static_cast<int>(gidx + NRG) * 4;
                        auto reinit = [&](const uint32_t (&b8)[8]) { tm_st8(d_addr, b8); tm_st8(d_addr + 8, b8); tm_st8(d_addr + 16, b8); tm_st8(d_addr + 24, b8); };
                        if (tn < t_end) reinit(init_c);
                        else if (!last && tn < t_end2) reinit(init_n);
                        else if (!last)
                        {
                            int ln = l + 2;     // (frames with fewer rows per layer than the ring has slots)
                            while (ln <= R && tn >= s_tb[ln + 1]) ln++;
                            if (ln <= R)
                            {
                                uint32_t init_x[8];
#pragma unroll
                                for (int c = 0; c < 8; c++) init_x[c] = __float_as_uint(bias_of(ln, c));
                                reinit(init_x);
                            }
                        }
                    }
                    // the accumulators are free again as soon as they have been read and re-initialised: publish that BEFORE the arithmetic
                    // (the ring's turn-around time is what bounds how far the issuers can run ahead)
                    ACB_TM_WAIT_ST();
                    ACB_TM_FENCE_BEFORE();
                    tm_publish_byte(my_flag_e + 8 * sg, use + 1);
#ifdef ACB_TM_TRACE
                    if (q == 0 && lane == 0) trace[3 * TM_MAX_STEPS + gidx] = clock64();
#endif
                    if (!last && k == 4 && !pads && !xclamp)
                    {
                        // the common case -- a whole group of a body layer inside an interior strip -- as straight-line code: activation and
                        // split of the four rows, ONE 32-column store of the next layer's operands, four progress bytes
                        uint32_t w[32];
#pragma unroll
                        for (int jr = 0; jr < 4; jr++)
                        {
                            float v[8];
#pragma unroll
                            for (int c = 0; c < 8; c++)
                            {
                                v[c] = __uint_as_float(d[8 * jr + c]);
                                v[c] = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? fmaxf(v[c], 0.0f) : prelu(v[c], alpha[c]);
                            }
                            split_pair(v[0], v[1], w[8 * jr + 0], w[8 * jr + 4]); split_pair(v[2], v[3], w[8 * jr + 1], w[8 * jr + 5]);
                            split_pair(v[4], v[5], w[8 * jr + 2], w[8 * jr + 6]); split_pair(v[6], v[7], w[8 * jr + 3], w[8 * jr + 7]);
                        }
                        tm_st32(my_a + 8 * yg, w);
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        const uint32_t fa = my_flag_a + 8 * yg;
                        tm_publish_byte(fa, l + 1); tm_publish_byte(fa + 8, l + 1); tm_publish_byte(fa + 16, l + 1); tm_publish_byte(fa + 24, l + 1);
                        continue;
                    }
#pragma unroll
                    for (int jr = 0; jr < 4; jr++)
                    {
                        if (jr >= k) break;
                        const int y = yg + jr;
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) v[c] = __uint_as_float(d[8 * jr + c]);
                        if (!last || !S::TAIL)
                        {
                            // body conv: activation, split, next layer's operand (or the segment's output map)
#pragma unroll
                            for (int c = 0; c < 8; c++) v[c] = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? fmaxf(v[c], 0.0f) : prelu(v[c], alpha[c]);
                            uint32_t w8[8];
                            split_pair(v[0], v[1], w8[0], w8[4]); split_pair(v[2], v[3], w8[1], w8[5]);
                            split_pair(v[4], v[5], w8[2], w8[6]); split_pair(v[6], v[7], w8[3], w8[7]);
                            if (!last) put_row(l, y, w8);
                            else
                            {
                                const int gx = x0 + R + lane, gy = y0 + y;
                                if (lane < SW && gx < prm.w)
                                {
                                    const size_t o = static_cast<size_t>(gy) * prm.w + gx, plane = static_cast<size_t>(prm.w) * prm.h;
                                    prm.map_out[o] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                                    prm.map_out[plane + o] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                                }
                            }
                        }
                        else if constexpr (S::TAIL && S::FAM == ACB200_FAMILY_ACNET_LEGACY)
                        {
                            // conv + ReLU, then the 2x2 deconvolution (Common.hpp:344-393): four dots of 8, no bias
                            constexpr int KD = S::HEAD ? 72 : 0;
#pragma unroll
                            for (int c = 0; c < 8; c++) v[c] = fmaxf(v[c], 0.0f);
                            float o4[4];
#pragma unroll
                            for (int jo = 0; jo < 4; jo++)
                            {
                                float s = v[0] * prm.k[KD + jo * 8];
#pragma unroll
                                for (int c = 1; c < 8; c++) s = fmaf(v[c], prm.k[KD + jo * 8 + c], s);
                                o4[jo] = s;
                            }
                            const int gx = x0 + R + lane, gy = y0 + y;
                            if (lane < SW && gx < prm.w)
                            {
                                uint8_t* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy) * prm.dst_pitch;
                                if (prm.type == ACB200_UINT8 && aligned)
                                {
                                    const uint8_t q0 = static_cast<uint8_t>(fmaf(__saturatef(o4[0]), 255.0f, 0.5f)), q1 = static_cast<uint8_t>(fmaf(__saturatef(o4[1]), 255.0f, 0.5f));
                                    const uint8_t q2 = static_cast<uint8_t>(fmaf(__saturatef(o4[2]), 255.0f, 0.5f)), q3 = static_cast<uint8_t>(fmaf(__saturatef(o4[3]), 255.0f, 0.5f));
                                    *reinterpret_cast<uchar2*>(row + 2 * gx) = make_uchar2(q0, q1);
                                    *reinterpret_cast<uchar2*>(row + prm.dst_pitch + 2 * gx) = make_uchar2(q2, q3);
                                }
                                else
                                {
                                    net_store2(row, 2 * gx, prm.type, o4[0], o4[1], aligned);
                                    net_store2(row + prm.dst_pitch, 2 * gx, prm.type, o4[2], o4[3], aligned);
                                }
                            }
                        }
                        else if constexpr (S::TAIL && S::FAM == ACB200_FAMILY_ACNET)
                        {
                            // conv 8 -> 4 (+ bias, already in the accumulator), + nearest-upsampled luma, pixel shuffle (Common.hpp:290-342)
                            const int gx = x0 + R + lane, gy = y0 + y;
                            const float id = luma[(y + 1) * TM_LP + R + lane + 1];
                            if (lane < SW && gx < prm.w)
                            {
                                uint8_t* row = static_cast<uint8_t*>(prm.dst) + static_cast<size_t>(2 * gy) * prm.dst_pitch;
                                if (prm.type == ACB200_UINT8 && aligned)
                                {
                                    const uint8_t q0 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v[0] + id), 255.0f), 0.5f)), q1 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v[1] + id), 255.0f), 0.5f));
                                    const uint8_t q2 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v[2] + id), 255.0f), 0.5f)), q3 = static_cast<uint8_t>(__fadd_rn(__fmul_rn(__saturatef(v[3] + id), 255.0f), 0.5f));
                                    *reinterpret_cast<uchar2*>(row + 2 * gx) = make_uchar2(q0, q1);
                                    *reinterpret_cast<uchar2*>(row + prm.dst_pitch + 2 * gx) = make_uchar2(q2, q3);
                                }
                                else
                                {
                                    net_store2(row, 2 * gx, prm.type, v[0] + id, v[1] + id, aligned);
                                    net_store2(row + prm.dst_pitch, 2 * gx, prm.type, v[2] + id, v[3] + id, aligned);
                                }
                            }
                        }
                    }
                    if (!last)
                    {
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        publish_rows(l, yg, k);
                    }
#ifdef ACB_TM_TRACE
                    if (q == 0 && lane == 0) trace[4 * TM_MAX_STEPS + gidx] = clock64();
#endif
                }
            }
        }
        ACB_TM_FENCE_BEFORE();
        __syncthreads();
#ifdef ACB_TM_TRACE
        if (traced && threadIdx.x == 0)
        {
            const long long t0 = trace[0];
            for (int i = 0; i < *s_nsteps; i += TM_CHUNK) printf("S %d %lld %lld\n", i, trace[i] - t0, trace[TM_MAX_STEPS + i] - t0);
            for (int g = 0; g < s_tb[R + 1] / 4; g++) printf("R %d %lld %lld %lld\n", g, trace[2 * TM_MAX_STEPS + g] - t0, trace[3 * TM_MAX_STEPS + g] - t0, trace[4 * TM_MAX_STEPS + g] - t0);
        }
#endif
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
    }
}
