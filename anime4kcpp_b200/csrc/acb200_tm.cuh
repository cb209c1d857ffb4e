// TMEM-resident tensor engine of the fused luma network (tcgen05, sm_100a): activations AND accumulators live in tensor
// memory for a whole segment; shared memory holds only the weights and the luma tile.
//
// What the reference does per 2x pass (core/src/processor/cuda/Kernel.cu:246-514, CUDAProcessor.cpp:383-566): one launch per
// layer, fp16 activations through HBM.  What the round-1 engines do: 56x56 frames in shared memory, the 3x3 taps gathered by
// ldmatrix (mma.sync) or by SS-mode descriptors (tcgen05, 44 cycles per MMA: the A fetch from shared memory is the bound).
// Measured facts this engine is built on (tools/microbench_tcgen05_ts.cu, profiles/r02_microbench_tcgen05_ts.txt):
//   * tcgen05.mma with the A operand in TMEM (M = 128, K = 16) costs N / 2 cycles, no operand fetch.
//   * the .ashift qualifier (multiply, then shift the A rows by one lane inside every 32-lane quadrant: new[r] = old[r + 1])
//     is free.  A horizontal tap is therefore a lane shift of the operand -- no data movement by any thread.
//   * a vertical tap is a different A operand (another 8-column group of TMEM) -- free as well.
//   * the issuing thread cannot run ahead of the tensor pipe (every UTCHMMA holds a scoreboard on its uniform-register
//     operands until it is dequeued), so anything it waits for becomes a pipe bubble: several issuer warps take turns.
//
// Layout.  A CTA owns four independent STRIPS, one per TMEM lane quadrant: 32 pixels wide (lane = x) and G <= 32 rows tall.
// TMEM holds two frame buffers of 8 columns per row.  Row y of a layer's 8-channel map is ONE A operand: the pixel's eight
// channels as split fp16 (hi words 0-3, lo words 4-7), i.e. K = 16 = {a_hi | a_lo}.  Per row r of the input map the issuing
// thread runs three alignments (dx = -1, 0, +1 through two .ashift's), each two MMAs with N = 24 = 3 output rows x 8 couts:
// (a_hi + a_lo) w_hi and a_hi w_lo accumulate into the SAME columns; row r contributes to output rows r+1, r, r-1 with the
// weights of dy = -1, 0, +1, whose accumulators are adjacent rows of the OTHER frame buffer.  All nine taps and all three
// split-precision products are accumulated inside the tensor core: 76 tensor cycles per 128 pixels and layer.
//   The accumulate flag is always set: whoever writes row y of a layer also pre-loads row y of the other buffer (dead by then)
// with the next layer's bias, which makes the bias add free and the order of the MMAs across rows irrelevant.
//   Epilogue (warps 4.., one warp per lane quadrant and group of four rows): tcgen05.ld of 32 columns, activation, fp16 split,
// tcgen05.st of the next layer's operands IN PLACE.  Nothing is exchanged between threads and no shared memory is touched.
//   The lane shift drifts the map by one lane per layer (lane j of layer l is pixel x0 + l + j), which consumes exactly the
// halo a fused segment loses anyway: after R layers lanes 0 .. 31 - 2R hold the strip's 32 - 2R output columns.
//   Layers are not separated by barriers: rows flow (a progress byte per row and quadrant, a one-shot mbarrier per group of
// accumulator rows), so the tensor pipe runs across layer boundaries and under the head conv / the map load.
//   Replicate padding (border CTAs): the epilogue copies the edge values one pixel outwards -- a warp shuffle in x, a second
// tcgen05.st in y.  Rows outside the image are skipped.
#pragma once

#include <cuda_fp16.h>

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"
#include "acb200_mma.cuh"
#include "acb200_colour.cuh"

namespace acb
{
#ifndef ACB_TM_EPI_SETS
#define ACB_TM_EPI_SETS 4
#endif
    // TMEM: two frame buffers of 8 columns per frame row (columns 0 .. 255 and 256 .. 511).  Layer l reads its A operands from
    // buffer (l - 1) % 2 and accumulates into buffer l % 2; the epilogue converts an accumulator row IN PLACE into the next layer's
    // operand row and pre-loads the same row of the other buffer (whose operand is dead by then) with the next layer's bias.
    // Frame row y lives at columns 8 (y + 1) of a buffer: rows -1 and G are scratch (the edge steps of a layer add into the row just
    // outside its output range, so that every MMA has the same shape: three output rows, weights at row block 0).
    constexpr int TM_GMAX = 30;                     // rows of a strip frame
    constexpr int TM_MAX_GROUPS = 8 * 8;            // one-shot `full` barriers: groups of four rows of a segment (<= 8 per layer)
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
    constexpr int TM_ISSUERS = ACB_TM_ISSUERS;      // issuer warps, each takes every TM_ISSUERS-th chunk of the step program
    // Warp roles: the epilogue warps come FIRST (warp w works on TMEM lane quadrant w % 4), the issuer warps LAST: the warp scheduler
    // favours the higher warp id among eligible warps, and an issuer that cannot get an issue slot starves the tensor pipe.
    constexpr int TM_EPI_WARP0 = 0;
    constexpr int TM_ISS_WARP0 = 4 * ACB_TM_EPI_SETS;
#ifndef ACB_TM_CHUNK
#define ACB_TM_CHUNK 4
#endif
    constexpr int TM_CHUNK = ACB_TM_CHUNK;          // consecutive steps (input rows) per chunk: 3 or 4 (a chunk record holds TM_CHUNK + 2 <= 6 waits)
    static_assert(TM_CHUNK == 3 || TM_CHUNK == 4, "chunk records are laid out for three or four steps");
#ifndef ACB_TM_LAG
#define ACB_TM_LAG 4
#endif
    constexpr int TM_LAG = ACB_TM_LAG;              // rounds between a layer's chunk k and the next layer's chunk k in the issue order, >= 3
    constexpr int TM_B_BYTES_HALF = 2 * 24 * 16;    // one B matrix: [2 K chunks][24 rows][8 fp16]
    constexpr int TM_B_BYTES_AL = 2 * TM_B_BYTES_HALF;      // one alignment: the w_hi matrix, then the w_lo matrix
    constexpr int TM_B_BYTES_LAYER = 3 * TM_B_BYTES_AL;     // 4608
    constexpr int TM_B_WORDS_LAYER = TM_B_BYTES_LAYER / 4;
    constexpr int TM_LP = 34;                       // luma tile pitch (floats): lanes 0..31 read columns j .. j + 2
    constexpr int TM_OFF_LUMA = TM_MAX_R * TM_B_BYTES_LAYER;
    constexpr int TM_OFF_BAR = TM_OFF_LUMA + 4 * (TM_GMAX + 2) * TM_LP * 4;
    constexpr int TM_N_BARS = TM_GMAX + TM_MAX_GROUPS + 1;
    constexpr int TM_OFF_GEOM = TM_OFF_BAR + TM_N_BARS * 8 + 8;
    constexpr int TM_OFF_STEPS = ((TM_OFF_GEOM + 3 * (TM_MAX_R + 2) * 4 + 4 + 15) / 16) * 16;      // the issuers' step program, 48 bytes per input row and layer
    constexpr int TM_MAX_CHUNKS = TM_MAX_R * ((TM_GMAX + 2 + TM_CHUNK - 1) / TM_CHUNK);     // chunks of up to TM_CHUNK steps, never across layers
    static_assert(TM_MAX_CHUNKS <= 32 * TM_ISSUERS_DECL, "one issuer-warp thread builds one chunk record");
    constexpr int TM_MAX_STEPS = 4 * TM_MAX_CHUNKS;
    constexpr int TM_CHUNK_WORDS = 20;              // one record per chunk: 4 operand words, 6 waits, 8 commits, 2 spare
#ifdef ACB_TM_TRACE
    constexpr int TM_OFF_TRACE = TM_OFF_STEPS + TM_MAX_CHUNKS * TM_CHUNK_WORDS * 4;
    constexpr int TM_SMEM_BYTES = TM_OFF_TRACE + 4 * TM_MAX_STEPS * 8;
#else
    constexpr int TM_SMEM_BYTES = TM_OFF_STEPS + TM_MAX_CHUNKS * TM_CHUNK_WORDS * 4;
#endif

#ifndef ACB_TM_HEAD_FFMA2
#define ACB_TM_HEAD_FFMA2 1
#endif
#ifndef ACB_TM_CHROMA_LUT
#define ACB_TM_CHROMA_LUT 1
#endif
    // fused chroma merge: two tables of 256 float2 after the engine's own shared memory (see the tail)
    constexpr int TM_OFF_LUT = (TM_SMEM_BYTES + 15) / 16 * 16;
    // progress barriers (ACB_TM_PROGRESS_MBAR): one one-shot mbarrier per (published layer 1 .. R, frame row), four arrivals (one per lane quadrant)
    constexpr int TM_OFF_PROG = TM_OFF_LUT + 2 * 256 * 8;
    // RGBA fusion: the resized alpha bytes of a tail group, one uint2 per (epilogue warp, row of the group, lane) -- parked in shared memory
    // between the precompute and the merge so that the RGB path carries no registers for them across the wait
    // (it lies BEHIND ARNet's residual store, so that only the opt-in RGBA launches of ARNet pay for both: every kilobyte of shared memory is
    // a kilobyte less L1 for the map loads -- 16 KB more cost the ARNet chain 5 %)
    constexpr int TM_AQ_BYTES = 4 * TM_SETS * 4 * 32 * 8;
    constexpr int TM_OFF_X = (TM_OFF_PROG + TM_MAX_R * TM_GMAX * 8 + 127) / 128 * 128;
    constexpr int TM_X_BYTES = 4 * (TM_GMAX + 2) * 32 * 32;
    constexpr int TM_SMEM_BYTES_FUSED = TM_OFF_X + TM_AQ_BYTES;
    // ARNet: the block input x of the residual `conv * 0.2 + x` (CPUProcessor.cpp:1479,1483), fp32, one 32-byte slot per frame row, lane and
    // quadrant.  Written by the epilogue that produces x, read two layers later by the epilogue of the block's second conv -- two lanes to
    // the left, because the map drifts by one lane per layer.
    constexpr int TM_SMEM_BYTES_ARNET = TM_OFF_X + TM_X_BYTES;
    constexpr int TM_SMEM_BYTES_ARNET_RGBA = TM_SMEM_BYTES_ARNET + TM_AQ_BYTES;

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
        int issuers;            // issuer warps used (1 .. TM_ISSUERS)
        // Colour handling fused into the segments (8-bit RGB, exactly 2x; all null: luma plane in, luma plane out).  Reference:
        // Processor.cpp:207-213 (rgb2yuv before the network), :251-253 (chroma resize + yuv2rgb after it), run there as separate CPU steps.
        const uint8_t* rgb_src; // NEEDS_LUMA segments: the luma tile is computed from this packed RGB image instead of read from `src`
        uint8_t* uv_out;        // HEAD segments: every pixel's quantised (u, v) is written here by the CTA that owns the pixel
        uint8_t* y_out;         // HEAD segments, optional: the quantised luma plane as well (families whose tail adds the source luma read it back)
        const uint8_t* uv_in;   // TAIL segments: Catmull-Rom of this (u, v) plane, re-quantise, YUV -> RGB merge, RGB to rgb_dst
        uint8_t* rgb_dst;
        const Contrib* htab;    // contributors of the 2x chroma resize, per output column / row
        const Contrib* vtab;
        int rgb_pitch, uv_pitch, rgb_dst_pitch, y_pitch;
        int uvc;                // channels of the chroma plane: 2 (u, v) for RGB, 3 (u, v, a) for RGBA (the source then has four bytes per pixel)
        const uint32_t* bops;   // B operands of this segment's 3x3 convs, TM_B_WORDS_LAYER words each, in layer order
        alignas(8) float k[(S::HEAD ? 72 : 0) + 64 + 32];  // fp32 weights used outside the MMAs: head (72, tap-major [9][8]) | ARNet 1x1 (64) | legacy deconv (32)
        alignas(8)
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
// nanoseconds an epilogue warp sleeps after an unsuccessful mbarrier try_wait (the try_wait loop was 13 % of all executed instructions;
// measured 0 -> 20 .. 200 ns: luma pass 0.2238 -> 0.2230 ms, fused RGB frame 0.2943 -> 0.2925 ms)
#ifndef ACB_TM_WAIT_SLEEP
#define ACB_TM_WAIT_SLEEP 50
#endif
    __device__ __forceinline__ void tm_wait(uint32_t bar, uint32_t parity)
    {
        uint32_t ok = 0, spins = 0;
        do
        {
            // (the suspend-time hint keeps a waiting warp asleep instead of spinning through the sub-partition's issue slots)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
            if (!ok && ++spins > (1u << 20)) __trap();      // a protocol bug must fault, never hang the device
            if (ACB_TM_WAIT_SLEEP > 0 && !ok) __nanosleep(ACB_TM_WAIT_SLEEP);
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
#ifndef ACB_TM_PROGRESS_MBAR
#define ACB_TM_PROGRESS_MBAR 0
#endif
// Nanoseconds an issuer warp sleeps between two polls of a progress word / a layer ticket.  Measured (tools/time_tm.py, 1080p luma pass,
// same box): 0 (spin) 0.2372 ms, 5 .. 40 ns 0.2245 ms, 200 ns 0.2417 ms -- a spinning issuer takes issue slots from the epilogue warps
// of its scheduler (the polling loop was 18 % of all executed instructions), a long sleep turns into tensor-pipe bubbles.  Progress
// mbarriers instead of polled bytes (ACB_TM_PROGRESS_MBAR) measured 0.2337 ms alone and 0.2257 ms with the ticket sleep: no better.
#ifndef ACB_TM_POLL_SLEEP
#define ACB_TM_POLL_SLEEP 20
#endif
#ifndef ACB_TM_TICKET_SLEEP
#define ACB_TM_TICKET_SLEEP ACB_TM_POLL_SLEEP
#endif
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
            if (ACB_TM_POLL_SLEEP > 0) __nanosleep(ACB_TM_POLL_SLEEP);
        }
    }
    // publish: plain byte stores.  What they publish are TMEM writes whose COMPLETION the publishing thread has already waited for
    // (tcgen05.wait::st / wait::ld, then tcgen05.fence::before_thread_sync), so no membar is needed in front of the flag store.
    __device__ __forceinline__ void tm_publish_byte(uint32_t addr, uint32_t value)
    {
        asm volatile("st.volatile.shared.u8 [%0], %1;" :: "r"(addr), "r"(value) : "memory");
    }

    // Row bookkeeping.  The rows a layer produces are handled in GROUPS of four consecutive rows (the last group of a layer may be
    // shorter): a group is one epilogue work item and one one-shot `full` mbarrier.
    // RGBA: the fused colour path works on four-channel images ((u, v, a) chroma plane, un-premultiplying merge); a compile-time switch so
    // that the RGB instantiation carries none of it (as a run-time branch it cost the RGB tail 170 bytes of spills and 3.6 % of the frame)
    template<class S, bool RGBA = false>
    __global__ void __launch_bounds__(TM_THREADS, 1) segment_tm_kernel(const __grid_constant__ TmParams<S> prm)
    {
        constexpr int R = S::R;                 // 3x3 convs of this segment (all on the tensor cores)
        constexpr int SW = 32 - 2 * R;          // output columns of a strip
        static_assert(R >= 1 && R <= TM_MAX_R && SW >= 8, "segment too deep for 32-pixel strips");
        constexpr bool ARNET = S::FAM == ACB200_FAMILY_ARNET;
        static_assert(!ARNET || (S::NCONV % 2) == 0, "ARNet segments start and end on block boundaries");
        static_assert(TM_SETS == 4, "groups are dealt to the epilogue sets by group index % 4");
        extern __shared__ __align__(128) unsigned char smem_tm[];
        float* luma_all = reinterpret_cast<float*>(smem_tm + TM_OFF_LUMA);
        const uint32_t bars = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm + TM_OFF_BAR));
        // flag_a[row]: progress words (see tm_wait_bytes); bar_full[group] (one-shot: every group of every layer has its own mbarrier,
        // so no phase parity can alias), bar_bop: mbarriers
        const uint32_t flag_a = bars, bar_full = bars + 8 * TM_GMAX, bar_bop = bar_full + 8 * TM_MAX_GROUPS;
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_tm + TM_OFF_BAR + TM_N_BARS * 8);
        uint4* steps = reinterpret_cast<uint4*>(smem_tm + TM_OFF_STEPS);
#ifdef ACB_TM_TRACE
        long long* trace = reinterpret_cast<long long*>(smem_tm + TM_OFF_TRACE);    // per chunk: [0] start, [1] issued; per group: [2] full seen, [3] rows published
        const bool traced = blockIdx.x == gridDim.x / 2 + 3;
        if (threadIdx.x == 0) trace[4 * TM_MAX_STEPS - 1] = clock64();
#endif
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int G = prm.G, SH = G - 2 * R;
        const int tile_x = blockIdx.x % prm.tiles_x, sy = blockIdx.x / prm.tiles_x;
        const int y0 = sy * SH - R;

        // ---- setup ---------------------------------------------------------------------------------------------------------------
        // geometry: every thread derives it for itself (a handful of integer operations per layer), nothing serial
        int g_ya[TM_MAX_R + 2], g_yb[TM_MAX_R + 2], g_gb[TM_MAX_R + 2];
        int nchunks = 0;
        {
            int gb = 0;
#pragma unroll
            for (int l = 0; l <= R; l++)
            {
                tm_rows(l, G, y0, prm.h, g_ya[l], g_yb[l]);
                g_gb[l] = gb;
                if (l >= 1 && g_yb[l] >= g_ya[l]) { gb += (g_yb[l] - g_ya[l] + 4) >> 2; nchunks += (g_yb[l] - g_ya[l] + 3 + TM_CHUNK - 1) / TM_CHUNK; }
            }
            g_ya[R + 1] = 0; g_yb[R + 1] = -1; g_gb[R + 1] = gb;
        }
        int* s_geom = reinterpret_cast<int*>(smem_tm + TM_OFF_GEOM);     // [l] = ya | yb << 8 | first group << 16, for loops that are not unrolled
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int l = 0; l <= R + 1; l++) s_geom[l] = (g_ya[l] & 0xff) | ((g_yb[l] & 0xff) << 8) | (g_gb[l] << 16);
        }
        if (threadIdx.x < TM_GMAX) asm volatile("st.shared.u32 [%0], %1;" :: "r"(flag_a + 8 * threadIdx.x), "r"(0));
        // tickets[l]: chunks of layer l whose MMAs have been issued (see the issuers)
        const uint32_t tickets = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm + TM_OFF_GEOM)) + 4 * (TM_MAX_R + 2);
        if (threadIdx.x >= 32 && threadIdx.x < 32 + TM_MAX_R + 2) asm volatile("st.shared.u32 [%0], %1;" :: "r"(tickets + 4 * (threadIdx.x - 32)), "r"(0));
        if (threadIdx.x >= 64 && threadIdx.x < 64 + TM_MAX_GROUPS)
        {
            // one `full` barrier per group: two commits per row
            const int gi = threadIdx.x - 64;
#pragma unroll
            for (int l = 1; l <= R; l++)
                if (gi >= g_gb[l] && gi < g_gb[l + 1])
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_full + 8 * gi), "r"(2 * min(4, g_yb[l] - (g_ya[l] + 4 * (gi - g_gb[l])) + 1)));
        }
        if (threadIdx.x == 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_bop));
        [[maybe_unused]] const uint32_t prog = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm + TM_OFF_PROG));
#if ACB_TM_PROGRESS_MBAR
        for (int i = threadIdx.x; i < R * TM_GMAX; i += TM_THREADS) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" :: "r"(prog + 8 * i));
#endif
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (warp == TM_ISS_WARP0)
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(tmem_slot))), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        if constexpr (S::NEEDS_LUMA)
        {
            // luma tile of every strip: frame columns -1 .. 32, frame rows -1 .. G, clamp-to-edge (the head conv's own padding).  One warp
            // per tile row: lane = column (coalesced), lanes 0 / 1 also take columns 32 / 33.
            constexpr int NWARPS = TM_THREADS / 32, PER = (4 * (TM_GMAX + 2) + NWARPS - 1) / NWARPS;
            if (prm.rgb_src != nullptr)
            {
                // Fused colour split (ImageProcess.cpp:38-61): luma = toFloat(quantised Y of the packed RGB pixel); the CTA that OWNS a pixel
                // (its output region, not the halo) also stores the pixel's quantised (u, v) for the tail segment's chroma resize.
                constexpr int BATCH = 4;
                const int own_y0 = y0 + R, own_y1 = min(own_y0 + G - 2 * R, prm.h);
#pragma unroll 1
                for (int i0 = 0; i0 < PER; i0 += BATCH)
                {
                    uint32_t c0[BATCH];
#pragma unroll
                    for (int i = 0; i < BATCH; i++)
                    {
                        const int row = min(warp + (i0 + i) * NWARPS, 4 * (G + 2) - 1);
                        const int q = row / (G + 2), ly = row - q * (G + 2);
                        const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
                        const int gy = clampi(y0 - 1 + ly, 0, prm.h - 1), gx0 = strip * SW - R - 1;
                        const uint8_t* srow = prm.rgb_src + static_cast<size_t>(gy) * prm.rgb_pitch;
                        const int xc = clampi(gx0 + lane, 0, prm.w - 1);
                        const uint8_t* pa = srow + 3 * xc;
                        // (RGBA sources are read as aligned 32-bit pixels; the host side only fuses them when rows and base are 4-byte aligned)
                        c0[i] = RGBA ? __ldg(reinterpret_cast<const uint32_t*>(srow) + xc)
                                             : __ldg(pa) | (static_cast<uint32_t>(__ldg(pa + 1)) << 8) | (static_cast<uint32_t>(__ldg(pa + 2)) << 16);
                    }
#pragma unroll
                    for (int i = 0; i < BATCH; i++)
                    {
                        const int row = warp + (i0 + i) * NWARPS;
                        if (row >= 4 * (G + 2)) break;
                        const int q = row / (G + 2), ly = row - q * (G + 2);
                        float* drow = luma_all + (q * (TM_GMAX + 2) + ly) * TM_LP;
                        uint8_t qy, qu, qv, qa = 0;
                        drow[lane] = RGBA ? luma_from_rgba_u8(c0[i] & 0xffu, (c0[i] >> 8) & 0xffu, (c0[i] >> 16) & 0xffu, c0[i] >> 24, qy, qu, qv, qa)
                                                  : luma_from_rgb_u8(c0[i] & 0xffu, (c0[i] >> 8) & 0xffu, (c0[i] >> 16) & 0xffu, qy, qu, qv);
                        if (S::HEAD && prm.uv_out != nullptr)
                        {
                            const int strip = tile_x * 4 + q, gy = y0 - 1 + ly, gx = strip * SW - R - 1 + lane;
                            if (strip < prm.strips_x && gy >= own_y0 && gy < own_y1 && gx >= strip * SW && gx < min(strip * SW + SW, prm.w))
                            {
                                uint8_t* uvp = prm.uv_out + static_cast<size_t>(gy) * prm.uv_pitch;
                                if (RGBA) { uvp[3 * gx] = qu; uvp[3 * gx + 1] = qv; uvp[3 * gx + 2] = qa; }
                                else *reinterpret_cast<uchar2*>(uvp + 2 * gx) = make_uchar2(qu, qv);
                                if (prm.y_out != nullptr) prm.y_out[static_cast<size_t>(gy) * prm.y_pitch + gx] = qy;
                            }
                        }
                    }
                }
                // the two extra tile columns (32, 33) of every row: one pixel per thread in ONE pass (done per row by lanes 0 / 1 they cost
                // every warp a full conversion per row)
                for (int t = threadIdx.x; t < 8 * (G + 2); t += TM_THREADS)
                {
                    const int row = t >> 1, q = row / (G + 2), ly = row - q * (G + 2);
                    const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
                    const int gy = clampi(y0 - 1 + ly, 0, prm.h - 1), gx = clampi(strip * SW - R - 1 + 32 + (t & 1), 0, prm.w - 1);
                    const uint8_t* pb = prm.rgb_src + static_cast<size_t>(gy) * prm.rgb_pitch + (RGBA ? 4 : 3) * gx;
                    uint8_t qy, qu, qv, qa;
                    luma_all[(q * (TM_GMAX + 2) + ly) * TM_LP + 32 + (t & 1)] = RGBA ? luma_from_rgba_u8(__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), __ldg(pb + 3), qy, qu, qv, qa)
                                                                                                : luma_from_rgb_u8(__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), qy, qu, qv);
                }
            }
            else if (prm.type == ACB200_UINT8)
            {
                // all of this warp's rows are requested before the first is consumed (one exposed memory latency, not one per row)
                uint8_t p0[PER], p1[PER];
#pragma unroll
                for (int i = 0; i < PER; i++)
                {
                    const int row = min(warp + i * NWARPS, 4 * (G + 2) - 1);
                    const int q = row / (G + 2), ly = row - q * (G + 2);
                    const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
                    const int gy = clampi(y0 - 1 + ly, 0, prm.h - 1), gx0 = strip * SW - R - 1;
                    const uint8_t* srow = static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch;
                    p0[i] = __ldg(srow + clampi(gx0 + lane, 0, prm.w - 1));
                    p1[i] = __ldg(srow + clampi(gx0 + 32 + (lane & 1), 0, prm.w - 1));
                }
#pragma unroll
                for (int i = 0; i < PER; i++)
                {
                    const int row = warp + i * NWARPS;
                    if (row >= 4 * (G + 2)) break;
                    const int q = row / (G + 2), ly = row - q * (G + 2);
                    float* drow = luma_all + (q * (TM_GMAX + 2) + ly) * TM_LP;
                    drow[lane] = unit_from_int<255>(static_cast<float>(p0[i]));     // toFloat<u8>, exact, no division
                    if (lane < 2) drow[32 + lane] = unit_from_int<255>(static_cast<float>(p1[i]));
                }
            }
            else
                for (int row = warp; row < 4 * (G + 2); row += NWARPS)
                {
                    const int q = row / (G + 2), ly = row - q * (G + 2);
                    const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
                    const int gy = clampi(y0 - 1 + ly, 0, prm.h - 1), gx0 = strip * SW - R - 1;
                    const uint8_t* srow = static_cast<const uint8_t*>(prm.src) + static_cast<size_t>(gy) * prm.src_pitch;
                    float* drow = luma_all + (q * (TM_GMAX + 2) + ly) * TM_LP;
                    drow[lane] = load_elem(srow, clampi(gx0 + lane, 0, prm.w - 1), prm.type);
                    if (lane < 2) drow[32 + lane] = load_elem(srow, clampi(gx0 + 32 + lane, 0, prm.w - 1), prm.type);
                }
        }
        if constexpr (S::TAIL)
            if (prm.uv_in != nullptr && threadIdx.x < 512)
            {
                // Fused merge: everything the YUV -> RGB step derives from a quantised chroma byte q, tabulated once per CTA with the very
                // operations the merge kernel applies per pixel (toFloat<u8>, - 0.5, the four products of ImageProcess.cpp:191-215):
                // lut[q] = (0.344 u, 1.773 u), lut[256 + q] = (1.403 v, 0.714 v)
                const float c = __fsub_rn(unit_from_int<255>(static_cast<float>(threadIdx.x & 255)), 0.5f);
                reinterpret_cast<float2*>(smem_tm + TM_OFF_LUT)[threadIdx.x] =
                    threadIdx.x < 256 ? make_float2(__fmul_rn(0.344f, c), __fmul_rn(1.773f, c)) : make_float2(__fmul_rn(1.403f, c), __fmul_rn(0.714f, c));
            }
        if (threadIdx.x == 32)
        {
            // the segment's B operands: one bulk copy (UBLKCP) on an mbarrier
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_bop), "r"(R * TM_B_BYTES_LAYER) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm))), "l"(reinterpret_cast<uint64_t>(prm.bops)), "r"(R * TM_B_BYTES_LAYER), "r"(bar_bop) : "memory");
        }
#ifdef ACB_TM_TRACE
        if (threadIdx.x == 0) trace[4 * TM_MAX_STEPS - 2] = clock64();
        if (threadIdx.x == 300) trace[4 * TM_MAX_STEPS - 3] = clock64();
#endif
        ACB_TM_FENCE_BEFORE();
        __syncthreads();
        ACB_TM_FENCE_AFTER();
        const uint32_t tmem = *tmem_slot;
#ifdef ACB_TM_TRACE
        if (threadIdx.x == 0) trace[4 * TM_MAX_STEPS - 4] = clock64() + (tmem & 1);
#endif
        constexpr int B0 = S::HEAD ? 8 : 0;     // bias / alpha offsets of the segment's first 3x3 conv inside prm.b / prm.a
        constexpr int A0 = (S::FAM == ACB200_FAMILY_ACNET && S::HEAD) ? 8 : 0;
        // bias of the segment's tensor layer ln (1-based) as the initial accumulator of output channel c (the ACNet tail has 4 couts)
        auto bias_of = [&](const int ln, const int c) -> float {
            if (S::TAIL && S::FAM != ACB200_FAMILY_ACNET_LEGACY && ln == R && c >= 4) return 0.0f;       // the pixel-shuffle conv has four couts
            // (ARNet's tail: PReLU conv, residual conv, [1x1 bias], pixel-shuffle conv -- the 1x1's bias sits between the last two)
            return prm.b[B0 + 8 * (ln - 1) + ((ARNET && S::TAIL && ln == R) ? 8 : 0) + c];
        };
        if (warp >= TM_ISS_WARP0)
        {
            // The issuers' program: one record of TM_CHUNK_WORDS words per CHUNK of up to four consecutive input rows of one layer, built by
            // the issuer warps (the other warps are already producing layer 0) -- one thread per chunk:
            //   [0] A operand of the first step   [1] accumulator of the first step (output row r - 1)   [2] B descriptor low word (alignment 0)
            //   [3] steps in the chunk            [4..9] waits (progress word | count << 24; 0 = none): rows r0 - 1 .. r0 + steps of the previous
            //   layer -- the input rows, and the rows whose producers pre-loaded the accumulators the chunk adds to
            //   [10..17] `full` barriers to commit to after the chunk's MMAs (0 = none)   [18] the layer's ticket | chunk index in the layer << 24
            // Every step has the same shape -- A row r into output rows r - 1 .. r + 1 -- because the first and last steps of a layer simply add
            // into the scratch rows outside the layer's output range.  Chunk c of the program is issued by issuer warp c % issuers.  All MMAs accumulate, so
            // their order across rows does not matter; an output row's three input rows lie in at most two chunks, hence every row
            // contributes exactly two commits to its group's `full` barrier.
            const uint32_t bop_s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tm));
            constexpr uint32_t DESC_HI = ((384u >> 4) & 0x3FFF) << 16;                              // LBO (K chunk distance: 24 rows) in the low word
            const int c_mine = threadIdx.x - 32 * TM_ISS_WARP0;
            if (c_mine < nchunks)
            {
                // which layer, which chunk of it
                int l = 1, cb = 0;
                bool found = false;
#pragma unroll
                for (int ll = 1; ll <= R; ll++)
                {
                    const int nc = g_yb[ll] >= g_ya[ll] ? (g_yb[ll] - g_ya[ll] + 3 + TM_CHUNK - 1) / TM_CHUNK : 0;
                    if (!found)
                    {
                        if (c_mine >= cb + nc) { cb += nc; l = ll + 1; } else found = true;
                    }
                }
                int ya = 0, yb = -1, gb = 0;
#pragma unroll
                for (int ll = 1; ll <= R; ll++) if (ll == l) { ya = g_ya[ll]; yb = g_yb[ll]; gb = g_gb[ll]; }
                const int n = yb - ya + 3;                      // steps of the layer: input rows ya - 1 .. yb + 1
                const int k0 = (c_mine - cb) * TM_CHUNK, nst = min(TM_CHUNK, n - k0);
                const int r0 = ya - 1 + k0;
                const uint32_t buf_a = tmem + 256u * static_cast<uint32_t>((l - 1) & 1), buf_d = tmem + 256u * static_cast<uint32_t>(l & 1);
                // Program order: a WAVEFRONT over the layers, not layer after layer.  Chunk k of layer l needs the rows that chunks <= k + 2 of
                // layer l - 1 produce (MMAs, commit, epilogue, publish: ~2500 cycles); a frame of 30 rows is only eight chunks per layer,
                // so issued layer by layer the dependent chunk would follow its producer too closely and every layer would start by waiting.
                // In round t the program holds chunk t - (l - 1) LAG of every layer l (deepest layer first): producer and consumer are
                // LAG - 2 rounds apart, with the other layers' chunks in between.
                int pos = 0;
                {
                    const int t = (c_mine - cb) + (l - 1) * TM_LAG;
#pragma unroll
                    for (int ll = 1; ll <= R; ll++)
                    {
                        const int nc = g_yb[ll] >= g_ya[ll] ? (g_yb[ll] - g_ya[ll] + 3 + TM_CHUNK - 1) / TM_CHUNK : 0;
                        const int kk = t - (ll - 1) * TM_LAG;           // layer ll's chunk of round t
                        pos += min(nc, max(0, kk));                     // ... its chunks of earlier rounds
                        if (ll > l && kk >= 0 && kk < nc) pos++;        // ... and, within the round, the deeper layers come first
                    }
                }
                uint32_t* rec = reinterpret_cast<uint32_t*>(steps) + TM_CHUNK_WORDS * pos;
                rec[0] = buf_a + 8 * (r0 + 1);
                rec[1] = buf_d + 8 * (r0 - 1 + 1);
                rec[2] = (((bop_s + (l - 1) * TM_B_BYTES_LAYER) >> 4) & 0x3FFF) | DESC_HI;
                rec[3] = static_cast<uint32_t>(nst);
                const uint32_t need = static_cast<uint32_t>(l) << 24;       // layer l - 1 publishes l
#pragma unroll
                for (int w = 0; w < 6; w++)
                {
                    const int row = r0 - 1 + w;                 // needed as an input row (r0 .. r0 + nst - 1) or as a pre-loaded accumulator row (inside [ya, yb])
                    const bool input = w >= 1 && w <= nst, accum = w <= nst + 1 && row >= ya && row <= yb;
#if ACB_TM_PROGRESS_MBAR
                    rec[4 + w] = (input || accum) ? (prog + 8 * ((l - 1) * TM_GMAX + row)) : 0u;
#else
                    rec[4 + w] = (input || accum) ? ((flag_a + 8 * row) | need) : 0u;
#endif
                }
                // commits (issued after the chunk's last MMA): output row rel (relative to ya) is touched by the layer's steps rel, rel + 1,
                // rel + 2; every chunk that touches it commits once -- twice when all three steps are its own
                int ne = 0;
                const int k1 = k0 + nst - 1;
#pragma unroll
                for (int t = 0; t < TM_CHUNK + 2; t++)
                {
                    const int rel = k0 - 2 + t;
                    if (rel < 0 || rel > yb - ya || rel > k1) continue;
                    const uint32_t bar = bar_full + 8 * (gb + (rel >> 2));
                    rec[10 + ne++] = bar;
                    if (rel >= k0 && rel + 2 <= k1) rec[10 + ne++] = bar;
                }
                for (; ne < 8; ne++) rec[10 + ne] = 0u;
                rec[18] = (tickets + 4 * l) | (static_cast<uint32_t>(c_mine - cb) << 24);      // the layer's ticket | chunks of the layer before this one
            }
#ifdef ACB_TM_TRACE
            if (threadIdx.x == 32 * TM_ISS_WARP0) trace[4 * TM_MAX_STEPS - 5] = clock64();
#endif
            asm volatile("bar.sync 1, %0;" :: "n"(32 * TM_ISSUERS) : "memory");     // the issuer warps only
        }

        if (warp >= TM_ISS_WARP0 && warp - TM_ISS_WARP0 < prm.issuers)
        {
            // ==== MMA issuers: warp w takes chunks w, w + issuers, ... of the program ==================================================
            // Neighbouring chunks of a layer add into the two accumulator rows at their boundary, and fp32 sums depend on the order of the
            // additions: left alone, two warps would race and the last bit of those rows would change from run to run.  So a layer's
            // chunks are STARTED in order and every chunk issues its rows last to first: chunk k + 1 waits for the layer's ticket to reach
            // k + 1, which chunk k's issuer sets once its LAST row's MMAs -- the ones that add into the shared rows -- are in the queue, and
            // chunk k + 1 reaches ITS contribution to those rows (its first row) only after eighteen MMAs of its own.  The tensor pipe
            // executes in issue order, so the additions into a shared row always happen in the same order, while the bulk of
            // neighbouring chunks is still issued concurrently by different warps.
            tm_wait(bar_bop, 0);
            constexpr uint64_t DESC_TOP = static_cast<uint64_t>(((128u >> 4) & 0x3FFF) | (1u << 14)) << 32;     // SBO | descriptor version, high word
            constexpr uint32_t IDESC = (1u << 4) | (static_cast<uint32_t>(128 >> 4) << 24) | (3u << 17);       // D f32, A / B f16 K-major, M = 128, N = 24
            const uint32_t* prog = reinterpret_cast<const uint32_t*>(steps);
#pragma unroll 1
            for (int c = warp - TM_ISS_WARP0; c < nchunks; c += prm.issuers)
            {
                const uint32_t* rec = prog + TM_CHUNK_WORDS * c;
#ifdef ACB_TM_TRACE
                if (lane == 0) trace[c] = clock64();
#endif
                // all of the chunk's waits are polled at once, one per lane
                {
                    const uint32_t w = lane < 6 ? rec[4 + lane] : (lane == 6 ? rec[18] : 0u);
#if ACB_TM_PROGRESS_MBAR
                    if (lane < 6) { if (w) tm_wait(w, 0); }
#else
                    if (lane < 6) { if (w) tm_wait_bytes(w & 0xffffffu, w >> 24); }
#endif
                    else if (w >> 24)
                    {
                        uint32_t t;
                        do
                        {
                            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(t) : "r"(w & 0xffffffu) : "memory");
                            if (ACB_TM_TICKET_SLEEP > 0 && t < (w >> 24)) __nanosleep(ACB_TM_TICKET_SLEEP);
                        } while (t < (w >> 24));
                    }
                    __syncwarp();
                }
                if (tm_elect_one())
                {
#ifdef ACB_TM_TRACE
                    trace[TM_MAX_STEPS + c] = clock64();
#endif
                    ACB_TM_FENCE_AFTER();
                    const uint4 op = *reinterpret_cast<const uint4*>(rec);
                    const uint2 c0 = *reinterpret_cast<const uint2*>(rec + 10);
                    const uint4 c1 = *reinterpret_cast<const uint4*>(rec + 12);
                    const uint2 c2 = *reinterpret_cast<const uint2*>(rec + 16);
                    const uint32_t bl = op.z;
                    const int nst = static_cast<int>(op.w);
                    const uint32_t tk = rec[18];
                    // step p: A row r0 + p into the accumulators of output rows r0 + p - 1 .. r0 + p + 1, operands in uniform registers; per
                    // alignment the w_hi matrix ((a_hi + a_lo) w_hi), then the w_lo matrix (a_hi w_lo) into the SAME 8 columns per output row;
                    // the operand shift rides on the alignment's last MMA
#pragma unroll
                    for (int pq = 0; pq < TM_CHUNK; pq++)
                    {
                        const int pp = TM_CHUNK - 1 - pq;       // last row first (see above)
                        if (pp >= nst) continue;
                        const uint32_t a = op.x + 8 * pp, dd = op.y + 8 * pp;
#pragma unroll
                        for (int al = 0; al < 3; al++)
                        {
                            const uint32_t o_hi = al * (TM_B_BYTES_AL >> 4), o_lo = o_hi + (TM_B_BYTES_HALF >> 4);
                            tm_mma(dd, a, DESC_TOP | (bl + o_hi), IDESC);
                            if (al < 2) tm_mma_ashift(dd, a, DESC_TOP | (bl + o_lo), IDESC); else tm_mma(dd, a, DESC_TOP | (bl + o_lo), IDESC);
                        }
                        // the chunk's last row is in the queue: the layer's next chunk may start
                        if (pp == nst - 1) asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(tk & 0xffffffu), "r"((tk >> 24) + 1) : "memory");
                    }
                    // output rows that have received this chunk's last contribution
                    if (c0.x) tm_commit(c0.x);
                    if (c0.y) tm_commit(c0.y);
                    if (c1.x) tm_commit(c1.x);
                    if (c1.y) tm_commit(c1.y);
                    if (c1.z) tm_commit(c1.z);
                    if (c1.w) tm_commit(c1.w);
                    if (c2.x) tm_commit(c2.x);
                    if (c2.y) tm_commit(c2.y);
                }
                __syncwarp();
            }
        }
        else if (warp < TM_ISS_WARP0)
        {
            // ==== producers of layer 0 and epilogue of every layer: one warp per lane quadrant (strip) and group of four rows ===============
            const int set = (warp - TM_EPI_WARP0) >> 2, q = warp & 3;
            const int strip = min(tile_x * 4 + q, prm.strips_x - 1);
            const int x0 = strip * SW - R;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            const float* luma = luma_all + q * (TM_GMAX + 2) * TM_LP;
            const int pad_top = -y0 - 1, pad_bot = prm.h - y0;          // frame rows of image rows -1 and h (replicate padding), if inside the frame
            const uint32_t my_t = tmem + lane_base;
            const uint32_t my_flag_a = flag_a + q;
            // a row of the map that layer `value - 1` produced is complete in this quadrant
            auto tm_publish = [&](const uint32_t flag_addr, const uint32_t value) {
#if ACB_TM_PROGRESS_MBAR
                if (lane == 0) tm_arrive(prog + 8 * ((value - 1) * TM_GMAX + ((flag_addr - my_flag_a) >> 3)));
#else
                tm_publish_byte(flag_addr, value);
#endif
            };
            const bool pads = (pad_top >= 0) || (pad_bot <= G - 1);     // the frame reaches over the top / bottom image edge

            // ARNet: the residual store (see TM_OFF_X): slot [frame row + 1][lane] of this quadrant, two float4
            [[maybe_unused]] float4* const sx = reinterpret_cast<float4*>(smem_tm + TM_OFF_X) + static_cast<size_t>(q) * (TM_GMAX + 2) * 64;
            [[maybe_unused]] auto sx_store = [&](const int y, const float (&v)[8]) {
                float4* p = sx + ((y + 1) * 32 + lane) * 2;
                p[0] = make_float4(v[0], v[1], v[2], v[3]);
                p[1] = make_float4(v[4], v[5], v[6], v[7]);
            };
            // x of the pixel this lane holds TWO layers after the one that stored it: two lanes to the right (lanes 30 / 31 lie in the consumed halo)
            [[maybe_unused]] auto sx_load = [&](const int y, float (&x)[8]) {
                const float4* p = sx + ((y + 1) * 32 + min(lane + 2, 31)) * 2;
                const float4 a = p[0], b = p[1];
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
            };
            // one row of layer l's output map becomes the next layer's A operand in buffer l % 2 (x-clamped in border strips), with its
            // padding copies in y
            auto put_row = [&](const int l, const int y, uint32_t (&w8)[8]) {
                const int L0 = -x0 - l, L1 = prm.w - 1 - x0 - l;       // lanes of image columns 0 and w - 1 in layer l's map
                if (L0 > 0 || L1 < 31)
                {
                    const int srcl = min(max(lane, L0), L1);
#pragma unroll
                    for (int c = 0; c < 8; c++) w8[c] = __shfl_sync(0xffffffffu, w8[c], srcl);
                }
                const uint32_t buf = my_t + 256u * static_cast<uint32_t>(l & 1);
                tm_st8(buf + 8 * (y + 1), w8);
                if (pads)
                {
                    if (y == pad_top + 1 && pad_top >= 0) tm_st8(buf + 8 * (pad_top + 1), w8);
                    if (y == pad_bot - 1 && pad_bot <= G - 1) tm_st8(buf + 8 * (pad_bot + 1), w8);
                }
            };
            // ... and is published (after tcgen05.wait::st): every lane stores the progress byte (same address, same value: one store, no branch)
            auto publish_rows = [&](const int l, const int ya_, const int k) {
                for (int j = 0; j < k; j++) tm_publish(my_flag_a + 8 * (ya_ + j), l + 1);
                if (pads)
                {
                    if (ya_ == pad_top + 1 && pad_top >= 0) tm_publish(my_flag_a + 8 * pad_top, l + 1);
                    if (ya_ + k - 1 == pad_bot - 1 && pad_bot <= G - 1) tm_publish(my_flag_a + 8 * pad_bot, l + 1);
                }
            };

            // ---- layer 0: the head conv (fp32 FFMA) or the previous segment's map, into buffer 0; buffer 1 gets layer 1's bias ----------------
            {
                const int ya = g_ya[0], yb = g_yb[0];
                uint32_t bias1[8];
#pragma unroll
                for (int c = 0; c < 8; c++) bias1[c] = __float_as_uint(bias_of(1, c));
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
                            // two output channels per packed FMA (FFMA2; the launcher stores the head's weights tap-major: k[p * 8 + co]); every
                            // channel still adds its nine products in tap order onto its bias
                            float v[8];
#if ACB_TM_HEAD_FFMA2
#pragma unroll
                            for (int c2 = 0; c2 < 4; c2++)
                            {
                                float2 s = *reinterpret_cast<const float2*>(&prm.b[2 * c2]);
#pragma unroll
                                for (int p = 0; p < 9; p++) s = fma2(make_float2(r9[p], r9[p]), *reinterpret_cast<const float2*>(&prm.k[p * 8 + 2 * c2]), s);
                                v[2 * c2] = s.x; v[2 * c2 + 1] = s.y;
                            }
#else
#pragma unroll
                            for (int co = 0; co < 8; co++)
                            {
                                float s = prm.b[co];
#pragma unroll
                                for (int p = 0; p < 9; p++) s = fmaf(r9[p], prm.k[p * 8 + co], s);
                                v[co] = s;
                            }
#endif
#pragma unroll
                            for (int co = 0; co < 8; co++)
                            {
                                if (ACT == ACT_RELU) v[co] = fmaxf(v[co], 0.0f);
                                else if (ACT == ACT_PRELU) v[co] = prelu(v[co], prm.a[co]);
                            }
                            if constexpr (ARNET)
                            {
                                // feat: the first block's residual input, and (through global memory) the long skip the last segment adds
                                sx_store(y, v);
                                const int su = tile_x * 4 + q, gx = x0 + lane, gy = y0 + y;
                                if (su < prm.strips_x && gx >= su * SW && gx < min(su * SW + SW, prm.w) && gy >= y0 + R && gy < min(y0 + G - R, prm.h))
                                {
                                    float4* f = reinterpret_cast<float4*>(prm.feat_out + (static_cast<size_t>(gy) * prm.w + gx) * 8);
                                    f[0] = make_float4(v[0], v[1], v[2], v[3]);
                                    f[1] = make_float4(v[4], v[5], v[6], v[7]);
                                }
                            }
                            uint32_t w8[8];
                            split_pair(v[0], v[1], w8[0], w8[4]); split_pair(v[2], v[3], w8[1], w8[5]);
                            split_pair(v[4], v[5], w8[2], w8[6]); split_pair(v[6], v[7], w8[3], w8[7]);
                            put_row(0, y, w8);
                            tm_st8(my_t + 256u + 8 * (y + 1), bias1);
                        }
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        if constexpr (ARNET) __threadfence_block();      // the residual store is read by another warp, ordered through the progress flags
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
                            if (yg + k <= lb) { tm_st8(my_t + 8 * (yg + k + 1), w8); tm_st8(my_t + 256u + 8 * (yg + k + 1), bias1); }
                            if constexpr (ARNET)
                                if (yg + k <= lb)
                                {
                                    const float2 f0 = join_pair(hi[k].x, lo[k].x), f1 = join_pair(hi[k].y, lo[k].y), f2 = join_pair(hi[k].z, lo[k].z), f3 = join_pair(hi[k].w, lo[k].w);
                                    const float xv[8] = { f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y };
                                    sx_store(yg + k, xv);
                                }
                        }
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        if constexpr (ARNET) __threadfence_block();
#pragma unroll
                        for (int k = 0; k < 4; k++) if (yg + k <= lb) tm_publish(my_flag_a + 8 * (yg + k), 1);
                    }
                }
            }

            // ---- layers 1 .. R ---------------------------------------------------------------------------------------------------------------
            const int es = prm.type & 0xff;
            const bool aligned = ((reinterpret_cast<uintptr_t>(prm.dst) | static_cast<uintptr_t>(prm.dst_pitch)) & (2 * es - 1)) == 0;
            (void)aligned;
#pragma unroll 1
            for (int l = 1; l <= R; l++)
            {
                const int geom = s_geom[l];
                const int ya = geom & 0xff, yb = static_cast<int>(static_cast<int8_t>((geom >> 8) & 0xff)), g0 = geom >> 16;
                const bool last = l == R;
                const bool xclamp = (-x0 - l > 0) || (prm.w - 1 - x0 - l < 31);      // the strip reaches over the left / right image edge in this layer's map
                const uint32_t buf_d = my_t + 256u * static_cast<uint32_t>(l & 1), buf_o = my_t + 256u * static_cast<uint32_t>((l + 1) & 1);
                uint32_t bias_n[8];         // the next layer's bias: pre-loaded into the other buffer's rows (its operands are dead)
#pragma unroll
                for (int c = 0; c < 8; c++) bias_n[c] = __float_as_uint(bias_of(min(l + 1, R), c));
                float alpha[8];
#pragma unroll
                for (int c = 0; c < 8; c++)
                    alpha[c] = S::FAM == ACB200_FAMILY_ACNET ? prm.a[A0 + 8 * (min(l, S::NCONV) - 1) + c]
                             : ARNET ? prm.a[((l - 1) >> 1) * 8 + c] : 0.0f;      // ARNet: the PReLU of the block's first conv (odd l; in bounds for every l <= R)
                // ARNet layer roles: odd l = conv + PReLU; even l = conv * 0.2 + x (x = the block's input, from the residual store); the tail's
                // even layer continues with the 1x1 conv, PReLU and the long skip (CPUProcessor.cpp:1479-1483)
                [[maybe_unused]] const bool is_res = ARNET && (l & 1) == 0;
                [[maybe_unused]] const bool is_1x1 = ARNET && S::TAIL && l == S::NCONV + 2;
                // this set's groups: group index g0 + j with (g0 + j) % 4 == set
                for (int j = (set - g0) & 3; ya + 4 * j <= yb; j += 4)
                {
                    const int yg = ya + 4 * j, k = min(4, yb - yg + 1);
                    // Fused chroma resize + merge (tail segments, prm.uv_in).  Everything that does not depend on the network -- the horizontal
                    // pass of the lane's two output columns over the group's source rows, the vertical pass, the re-quantisation of the resized
                    // (u, v) -- runs BEFORE the wait for the group's accumulators, in time the warp would otherwise spend asleep; what is kept
                    // is one word per output row: the quantised (u_a, v_a, u_b, v_b) bytes.  After the wait only the merge with the luma is left.
                    [[maybe_unused]] uint2 cq0 = make_uint2(0u, 0u), cq1 = cq0, cq2 = cq0, cq3 = cq0;
                    // RGBA: the resized alpha in the same form (bytes a_a, a_a, a_b, a_b), this warp's slots: [row of the group][lane]
                    [[maybe_unused]] uint2* const s_aq = reinterpret_cast<uint2*>(smem_tm + TM_OFF_X + (ARNET ? TM_X_BYTES : 0)) + (warp * 4) * 32 + lane;
                    const bool fused = S::TAIL && last && prm.uv_in != nullptr;
                    if constexpr (S::TAIL)
                        if (fused)
                        {
                            constexpr int PIX = RGBA ? 3 : 2;       // bytes per pixel of the chroma plane: (u, v) or (u, v, a)
                            const HTaps2 hk = load_htaps2(prm.htab, 2 * min(x0 + R + lane, prm.w - 1), prm.w, PIX);
                            const int gy0 = y0 + yg;
                            auto row_ptr = [&](const int gyr) { return prm.uv_in + static_cast<size_t>(clampi(gyr, 0, prm.h - 1)) * prm.uv_pitch; };
                            auto encode4 = [](const float4 sv) {
                                // stb encode (x 255 + 0.5, clamp, truncate): the byte is the low mantissa byte of the round-toward-zero magic sum
                                const uint32_t b0 = __float_as_uint(__fadd_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(sv.x, 255.0f), 0.5f), 0.0f), 255.0f), CM_MAGIC));
                                const uint32_t b1 = __float_as_uint(__fadd_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(sv.y, 255.0f), 0.5f), 0.0f), 255.0f), CM_MAGIC));
                                const uint32_t b2 = __float_as_uint(__fadd_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(sv.z, 255.0f), 0.5f), 0.0f), 255.0f), CM_MAGIC));
                                const uint32_t b3 = __float_as_uint(__fadd_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(sv.w, 255.0f), 0.5f), 0.0f), 255.0f), CM_MAGIC));
                                return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
                            };
                            // the whole chroma side of the group for one pair of channels: `hrow_at(image row)` is the horizontal pass of that pair
                            auto precompute = [&](auto hrow_at, uint2& q0, uint2& q1, uint2& q2, uint2& q3) {
                                auto vrows = [&](const int jr, const float4& w0, const float4& w1, const float4& w2, const float4& w3, const float4& w4) {
                                    const int gy = gy0 + jr;
                                    const Contrib* vp = prm.vtab + 2 * gy;
                                    const uint4 va0 = __ldg(reinterpret_cast<const uint4*>(vp)), vb0 = __ldg(reinterpret_cast<const uint4*>(vp + 1));
                                    const float2 va1 = __ldg(reinterpret_cast<const float2*>(&vp[0].c[2])), vb1 = __ldg(reinterpret_cast<const float2*>(&vp[1].c[2]));
                                    const float ca[4] = { __uint_as_float(va0.z), __uint_as_float(va0.w), va1.x, va1.y };
                                    const float cb[4] = { __uint_as_float(vb0.z), __uint_as_float(vb0.w), vb1.x, vb1.y };
                                    float4 sa, sb;      // vertical pass: the pair in columns a, b of output rows 2 gy (sa) and 2 gy + 1 (sb)
                                    if (static_cast<int>(va0.x) == gy - 2 && static_cast<int>(vb0.x) == gy - 1)
                                    {
                                        // interior rows: the window holds exactly the rows both contributors read
                                        sa.x = tap4(ca[0], ca[1], ca[2], ca[3], w0.x, w1.x, w2.x, w3.x); sa.y = tap4(ca[0], ca[1], ca[2], ca[3], w0.y, w1.y, w2.y, w3.y);
                                        sa.z = tap4(ca[0], ca[1], ca[2], ca[3], w0.z, w1.z, w2.z, w3.z); sa.w = tap4(ca[0], ca[1], ca[2], ca[3], w0.w, w1.w, w2.w, w3.w);
                                        sb.x = tap4(cb[0], cb[1], cb[2], cb[3], w1.x, w2.x, w3.x, w4.x); sb.y = tap4(cb[0], cb[1], cb[2], cb[3], w1.y, w2.y, w3.y, w4.y);
                                        sb.z = tap4(cb[0], cb[1], cb[2], cb[3], w1.z, w2.z, w3.z, w4.z); sb.w = tap4(cb[0], cb[1], cb[2], cb[3], w1.w, w2.w, w3.w, w4.w);
                                    }
                                    else
                                    {
                                        // rows at the top image edge (folded taps start at another row): the contributors' own rows, recomputed.  Rows past
                                        // the image carry zero coefficients and are read clamped.
                                        float4 t[4];
#pragma unroll
                                        for (int i = 0; i < 4; i++) t[i] = hrow_at(static_cast<int>(va0.x) + i);
                                        sa.x = tap4(ca[0], ca[1], ca[2], ca[3], t[0].x, t[1].x, t[2].x, t[3].x); sa.y = tap4(ca[0], ca[1], ca[2], ca[3], t[0].y, t[1].y, t[2].y, t[3].y);
                                        sa.z = tap4(ca[0], ca[1], ca[2], ca[3], t[0].z, t[1].z, t[2].z, t[3].z); sa.w = tap4(ca[0], ca[1], ca[2], ca[3], t[0].w, t[1].w, t[2].w, t[3].w);
#pragma unroll
                                        for (int i = 0; i < 4; i++) t[i] = hrow_at(static_cast<int>(vb0.x) + i);
                                        sb.x = tap4(cb[0], cb[1], cb[2], cb[3], t[0].x, t[1].x, t[2].x, t[3].x); sb.y = tap4(cb[0], cb[1], cb[2], cb[3], t[0].y, t[1].y, t[2].y, t[3].y);
                                        sb.z = tap4(cb[0], cb[1], cb[2], cb[3], t[0].z, t[1].z, t[2].z, t[3].z); sb.w = tap4(cb[0], cb[1], cb[2], cb[3], t[0].w, t[1].w, t[2].w, t[3].w);
                                    }
                                    return make_uint2(encode4(sa), encode4(sb));
                                };
                                // a window of five rows slides down the group (row gy needs rows gy - 2 .. gy + 2): eight horizontal passes per group
                                float4 h0 = hrow_at(gy0 - 2), h1 = hrow_at(gy0 - 1), h2 = hrow_at(gy0), h3 = hrow_at(gy0 + 1), h4 = hrow_at(gy0 + 2);
                                q0 = vrows(0, h0, h1, h2, h3, h4);
                                if (k > 1) { h0 = hrow_at(gy0 + 3); q1 = vrows(1, h1, h2, h3, h4, h0); }
                                if (k > 2) { h1 = hrow_at(gy0 + 4); q2 = vrows(2, h2, h3, h4, h0, h1); }
                                if (k > 3) { h2 = hrow_at(gy0 + 5); q3 = vrows(3, h3, h4, h0, h1, h2); }
                            };
                            if constexpr (!RGBA) precompute([&](const int gyr) { return chroma_hrow2<0>(row_ptr(gyr), hk); }, cq0, cq1, cq2, cq3);
                            else
                            {
                                precompute([&](const int gyr) { return chroma_hrow2<1>(row_ptr(gyr), hk); }, cq0, cq1, cq2, cq3);
                                uint2 a0 = make_uint2(0u, 0u), a1 = a0, a2 = a0, a3 = a0;
                                precompute([&](const int gyr) { return chroma_hrow2<2>(row_ptr(gyr), hk); }, a0, a1, a2, a3);
                                s_aq[0] = a0; s_aq[32] = a1; s_aq[64] = a2; s_aq[96] = a3;      // read back by the same thread: no synchronisation
                            }
                        }
                    // yl[dy * 2 + dx]: the lane's four luma results as the value BEFORE truncation to the byte (x 255 + 0.5 applied)
                    [[maybe_unused]] auto fused_store = [&](const int jr, const int gx, const int gy, const float (&yl)[4], const bool ok) {
                        const uint2 cq = jr == 0 ? cq0 : jr == 1 ? cq1 : jr == 2 ? cq2 : cq3;
                        constexpr bool rgba = RGBA;
                        uint2 aq = make_uint2(0u, 0u);
                        if constexpr (RGBA) aq = s_aq[32 * jr];
                        uint8_t* o = prm.rgb_dst + static_cast<size_t>(2 * gy) * prm.rgb_dst_pitch + (rgba ? 8 : 6) * gx;
#pragma unroll
                        for (int dy = 0; dy < 2; dy++)
                        {
                            const uint32_t c4 = dy ? cq.y : cq.x;
                            uint32_t ch[6];
                            uint32_t ab[2] = { 0u, 0u };
#pragma unroll
                            for (int dx = 0; dx < 2; dx++)
                            {
                                // toFloat of the stored luma byte; the chroma terms of YUV -> RGB (ImageProcess.cpp:191-215) from the per-CTA tables
                                const float yv = unit_from_int<255>(__fsub_rn(__fadd_rz(yl[2 * dy + dx], CM_MAGIC), CM_MAGIC));
#if ACB_TM_CHROMA_LUT
                                const float2 tu = *reinterpret_cast<const float2*>(smem_tm + TM_OFF_LUT + 8 * __byte_perm(c4, 0u, dx ? 0x4442 : 0x4440));
                                const float2 tv = *reinterpret_cast<const float2*>(smem_tm + TM_OFF_LUT + 2048 + 8 * __byte_perm(c4, 0u, dx ? 0x4443 : 0x4441));
                                float r = __fadd_rn(yv, tv.x);
                                float g = __fsub_rn(__fsub_rn(yv, tu.x), tv.y);
                                float b = __fadd_rn(yv, tu.y);
#else
                                const float qu = unit_from_int<255>(__fsub_rn(__uint_as_float(__byte_perm(c4, 0x4B000000u, dx ? 0x7642 : 0x7640)), CM_MAGIC));
                                const float qv = unit_from_int<255>(__fsub_rn(__uint_as_float(__byte_perm(c4, 0x4B000000u, dx ? 0x7643 : 0x7641)), CM_MAGIC));
                                const float u = __fsub_rn(qu, 0.5f), v = __fsub_rn(qv, 0.5f);
                                float r = __fadd_rn(yv, __fmul_rn(1.403f, v));
                                float g = __fsub_rn(__fsub_rn(yv, __fmul_rn(0.344f, u)), __fmul_rn(0.714f, v));
                                float b = __fadd_rn(yv, __fmul_rn(1.773f, u));
#endif
                                if (rgba)
                                {
                                    // yuva2rgba (ImageProcess.cpp:275-308): un-premultiply by the resized alpha, which is stored as it is
                                    const float al = unit_from_int<255>(__fsub_rn(__uint_as_float(__byte_perm(dy ? aq.y : aq.x, 0x4B000000u, dx ? 0x7642 : 0x7640)), CM_MAGIC));
                                    if (al > 1e-6f) { r = __fdiv_rn(r, al); g = __fdiv_rn(g, al); b = __fdiv_rn(b, al); }
                                    else r = g = b = 0.0f;
                                    ab[dx] = quant_u8(al);
                                }
                                ch[3 * dx + 0] = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(r), 255.0f), 0.5f), CM_MAGIC));
                                ch[3 * dx + 1] = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(g), 255.0f), 0.5f), CM_MAGIC));
                                ch[3 * dx + 2] = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(__saturatef(b), 255.0f), 0.5f), CM_MAGIC));
                            }
                            if (ok)
                            {
                                if (rgba)
                                {
                                    uint32_t* o32 = reinterpret_cast<uint32_t*>(o + dy * prm.rgb_dst_pitch);
                                    o32[0] = __byte_perm(__byte_perm(ch[0], ch[1], 0x0040), __byte_perm(ch[2], ab[0], 0x0040), 0x5410);
                                    o32[1] = __byte_perm(__byte_perm(ch[3], ch[4], 0x0040), __byte_perm(ch[5], ab[1], 0x0040), 0x5410);
                                }
                                else store_rgb2(o + dy * prm.rgb_dst_pitch, ch[0], ch[1], ch[2], ch[3], ch[4], ch[5]);
                            }
                        }
                    };
                    tm_wait(bar_full + 8 * (g0 + j), 0);
                    ACB_TM_FENCE_AFTER();
#ifdef ACB_TM_TRACE
                    if (q == 0 && lane == 0) trace[2 * TM_MAX_STEPS + g0 + j] = clock64();
#endif
                    if (!last && k == 4 && !pads && !xclamp && !is_1x1)
                    {
                        // the common case -- a whole group of a body layer inside an interior strip -- as straight-line code: one 32-column
                        // load, activation and split of the four rows, ONE 32-column store of the next layer's operands in place, the next
                        // layer's bias into the other buffer, four progress bytes
                        uint32_t d[32];
                        tm_ld32(d, buf_d + 8 * (yg + 1));
                        ACB_TM_WAIT_LD();
                        tm_st8(buf_o + 8 * yg + 8, bias_n); tm_st8(buf_o + 8 * yg + 16, bias_n); tm_st8(buf_o + 8 * yg + 24, bias_n); tm_st8(buf_o + 8 * yg + 32, bias_n);
                        uint32_t w[32];
#pragma unroll
                        for (int jr = 0; jr < 4; jr++)
                        {
                            float v[8];
#pragma unroll
                            for (int c = 0; c < 8; c++) v[c] = __uint_as_float(d[8 * jr + c]);
                            if (is_res)
                            {
                                float x[8];
                                sx_load(yg + jr, x);
                                __syncwarp();           // every lane has read its slot before its owner overwrites it
#pragma unroll
                                for (int c = 0; c < 8; c++) v[c] = fmaf(v[c], 0.2f, x[c]);
                                sx_store(yg + jr, v);
                            }
                            else
                            {
#pragma unroll
                                for (int c = 0; c < 8; c++) v[c] = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? fmaxf(v[c], 0.0f) : prelu(v[c], alpha[c]);
                            }
                            split_pair(v[0], v[1], w[8 * jr + 0], w[8 * jr + 4]); split_pair(v[2], v[3], w[8 * jr + 1], w[8 * jr + 5]);
                            split_pair(v[4], v[5], w[8 * jr + 2], w[8 * jr + 6]); split_pair(v[6], v[7], w[8 * jr + 3], w[8 * jr + 7]);
                        }
                        tm_st32(buf_d + 8 * (yg + 1), w);
                        ACB_TM_WAIT_ST();
                        ACB_TM_FENCE_BEFORE();
                        if constexpr (ARNET) __threadfence_block();
                        const uint32_t fa = my_flag_a + 8 * yg;
                        tm_publish(fa, l + 1); tm_publish(fa + 8, l + 1); tm_publish(fa + 16, l + 1); tm_publish(fa + 24, l + 1);
#ifdef ACB_TM_TRACE
                        if (q == 0 && lane == 0) trace[3 * TM_MAX_STEPS + g0 + j] = clock64();
#endif
                        continue;
                    }
                    for (int jr = 0; jr < k; jr++)
                    {
                        const int y = yg + jr;
                        uint32_t d[8];
                        tm_ld8(d, buf_d + 8 * (y + 1));
                        ACB_TM_WAIT_LD();
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) v[c] = __uint_as_float(d[c]);
                        if (!last || !S::TAIL)
                        {
                            // body conv: activation, split, next layer's operand (or the segment's output map)
                            if (is_res)
                            {
                                float x[8];
                                sx_load(y, x);
                                __syncwarp();
#pragma unroll
                                for (int c = 0; c < 8; c++) v[c] = fmaf(v[c], 0.2f, x[c]);
                                if constexpr (ARNET && S::TAIL)
                                    if (is_1x1)
                                    {
                                        // 1x1 conv over the eight channels, PReLU, + feat (the head's output, kept in global memory by the first segment)
                                        constexpr int K1 = S::HEAD ? 72 : 0, BT = B0 + 8 * S::NCONV, AT = (S::NCONV / 2) * 8;
                                        const int gx = clampi(x0 + l + lane, 0, prm.w - 1), gy = y0 + y;
                                        const float4* f = reinterpret_cast<const float4*>(prm.feat_in + (static_cast<size_t>(gy) * prm.w + gx) * 8);
                                        const float4 f0 = __ldg(f), f1 = __ldg(f + 1);
                                        const float ft[8] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w };
                                        float u[8];
#pragma unroll
                                        for (int co = 0; co < 8; co++)
                                        {
                                            float t = prm.b[BT + 16 + co];
#pragma unroll
                                            for (int ci = 0; ci < 8; ci++) t = fmaf(v[ci], prm.k[K1 + co * 8 + ci], t);
                                            u[co] = prelu(t, prm.a[AT + 8 + co]) + ft[co];
                                        }
#pragma unroll
                                        for (int c = 0; c < 8; c++) v[c] = u[c];
                                    }
                                if (xclamp)
                                {
                                    // border strips: lanes outside the image take the edge pixel's value here already, so that what the residual
                                    // store keeps for them is the replicate-padded map (put_row repeats the shuffle on the packed words)
                                    const int srcl = min(max(lane, -x0 - l), prm.w - 1 - x0 - l);
#pragma unroll
                                    for (int c = 0; c < 8; c++) v[c] = __shfl_sync(0xffffffffu, v[c], srcl);
                                }
                                if (!is_1x1) sx_store(y, v);
                            }
                            else
                            {
#pragma unroll
                                for (int c = 0; c < 8; c++) v[c] = S::FAM == ACB200_FAMILY_ACNET_LEGACY ? fmaxf(v[c], 0.0f) : prelu(v[c], alpha[c]);
                            }
                            uint32_t w8[8];
                            split_pair(v[0], v[1], w8[0], w8[4]); split_pair(v[2], v[3], w8[1], w8[5]);
                            split_pair(v[4], v[5], w8[2], w8[6]); split_pair(v[6], v[7], w8[3], w8[7]);
                            if (!last)
                            {
                                put_row(l, y, w8);
                                tm_st8(buf_o + 8 * (y + 1), bias_n);
                            }
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
                            if (fused)
                            {
                                const float yl[4] = { fmaf(__saturatef(o4[0]), 255.0f, 0.5f), fmaf(__saturatef(o4[1]), 255.0f, 0.5f),
                                                      fmaf(__saturatef(o4[2]), 255.0f, 0.5f), fmaf(__saturatef(o4[3]), 255.0f, 0.5f) };
                                fused_store(jr, gx, gy, yl, lane < SW && gx < prm.w);
                            }
                            else if (lane < SW && gx < prm.w)
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
                        else if constexpr (S::TAIL && S::FAM != ACB200_FAMILY_ACNET_LEGACY)
                        {
                            // conv 8 -> 4 (+ bias, already in the accumulator), + nearest-upsampled luma, pixel shuffle (Common.hpp:290-342)
                            const int gx = x0 + R + lane, gy = y0 + y;
                            const float id = luma[(y + 1) * TM_LP + R + lane + 1];
                            if (fused)
                            {
                                const float yl[4] = { __fadd_rn(__fmul_rn(__saturatef(v[0] + id), 255.0f), 0.5f), __fadd_rn(__fmul_rn(__saturatef(v[1] + id), 255.0f), 0.5f),
                                                      __fadd_rn(__fmul_rn(__saturatef(v[2] + id), 255.0f), 0.5f), __fadd_rn(__fmul_rn(__saturatef(v[3] + id), 255.0f), 0.5f) };
                                fused_store(jr, gx, gy, yl, lane < SW && gx < prm.w);
                            }
                            else if (lane < SW && gx < prm.w)
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
                        if constexpr (ARNET) __threadfence_block();
                        publish_rows(l, yg, k);
                    }
                }
            }
        }
        ACB_TM_FENCE_BEFORE();
        __syncthreads();
#ifdef ACB_TM_TRACE
        if (traced && threadIdx.x == 0)
        {
            const long long t0 = trace[4 * TM_MAX_STEPS - 1];
            printf("E %lld\n", clock64() - t0);
            printf("P thread 0 at barrier 1: %lld, thread 300: %lld, after barrier 1: %lld, steps built: %lld\n", trace[4 * TM_MAX_STEPS - 2] - t0, trace[4 * TM_MAX_STEPS - 3] - t0, trace[4 * TM_MAX_STEPS - 4] - t0, trace[4 * TM_MAX_STEPS - 5] - t0);
            for (int i = 0; i < nchunks; i++) printf("S %d %lld %lld\n", i, trace[i] - t0, trace[TM_MAX_STEPS + i] - t0);
            for (int g = 0; g < g_gb[R + 1]; g++) printf("R %d %lld %lld\n", g, trace[2 * TM_MAX_STEPS + g] - t0, trace[3 * TM_MAX_STEPS + g] - t0);
        }
#endif
        if (warp == TM_ISS_WARP0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
    }
}
