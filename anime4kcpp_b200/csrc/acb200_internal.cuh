// Declarations shared by the translation units of the CUDA layer (acb200.cu and one TU per engine: the engines' kernel
// templates are instantiated in their own TUs so the library builds in parallel).
#pragma once

#include <atomic>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "acb200_common.cuh"
#include "acb200_ffma.cuh"

namespace acbh
{
    using namespace acb;

    enum SegKind
    {
        SEG_LEGACY_FULL,    // <LEGACY, head, 7, tail>
        SEG_ACNET_B4,       // <ACNET, head, 4, tail>
        SEG_ACNET_B8,       // <ACNET, head, 8, tail>
        SEG_ACNET_B18_A,    // <ACNET, head, 9, ->
        SEG_ACNET_B18_B,    // <ACNET, -, 9, tail>
        SEG_ARNET_FIRST,    // <ARNET, head, ARNET_SEG, ->
        SEG_ARNET_MID,      // <ARNET, -, ARNET_SEG, ->
        SEG_ARNET_LAST,     // <ARNET, -, ARNET_SEG - 2, tail>
        SEG_LEGACY_A,       // <LEGACY, head, LEGACY_SPLIT, ->
        SEG_LEGACY_B,       // <LEGACY, -, 7 - LEGACY_SPLIT, tail>
        SEG_ACNET_B8_A,     // <ACNET, head, 4, ->
        SEG_ACNET_B8_B,     // <ACNET, -, 4, tail>
        SEG_ACNET_MID5,     // <ACNET, -, 5, ->
    };
    using SegLegacyFull = Seg<ACB200_FAMILY_ACNET_LEGACY, true, 7, true>;
    using SegAcnetB4 = Seg<ACB200_FAMILY_ACNET, true, 4, true>;
    using SegAcnetB8 = Seg<ACB200_FAMILY_ACNET, true, 8, true>;
    using SegAcnetB18A = Seg<ACB200_FAMILY_ACNET, true, 9, false>;
    using SegAcnetB18B = Seg<ACB200_FAMILY_ACNET, false, 9, true>;
    // ARNet: ARNET_SEG body convs per segment (4: T = 48, halo recompute 1.13x; 8: T = 40, 1.39x and half the map traffic)
#ifndef ACB_ARNET_SEG
#define ACB_ARNET_SEG 4
#endif
    constexpr int ARNET_SEG = ACB_ARNET_SEG;
    using SegArnetFirst = Seg<ACB200_FAMILY_ARNET, true, ARNET_SEG, false>;
    using SegArnetMid = Seg<ACB200_FAMILY_ARNET, false, ARNET_SEG, false>;
    using SegArnetLast = Seg<ACB200_FAMILY_ARNET, false, ARNET_SEG - 2, true>;
    // ACNetLegacy's seven body convs (+ the tail's conv) as head + LEGACY_SPLIT | (7 - LEGACY_SPLIT) + tail
#ifndef ACB_LEGACY_SPLIT
#define ACB_LEGACY_SPLIT 4
#endif
    constexpr int LEGACY_SPLIT = ACB_LEGACY_SPLIT;
    using SegLegacyA = Seg<ACB200_FAMILY_ACNET_LEGACY, true, LEGACY_SPLIT, false>;
    using SegLegacyB = Seg<ACB200_FAMILY_ACNET_LEGACY, false, 7 - LEGACY_SPLIT, true>;
    // ACNet-B8 as head + ACNET_SPLIT | (8 - ACNET_SPLIT) + tail; ACNet-B18 as head + ACNET_SPLIT | 5 | 5 | (8 - ACNET_SPLIT) + tail
#ifndef ACB_ACNET_SPLIT
#define ACB_ACNET_SPLIT 4
#endif
    constexpr int ACNET_SPLIT = ACB_ACNET_SPLIT;
    using SegAcnetB8A = Seg<ACB200_FAMILY_ACNET, true, ACNET_SPLIT, false>;
    using SegAcnetB8B = Seg<ACB200_FAMILY_ACNET, false, 8 - ACNET_SPLIT, true>;
    using SegAcnetMid5 = Seg<ACB200_FAMILY_ACNET, false, 5, false>;
#ifndef ACB_SPLIT_CHAINS
#define ACB_SPLIT_CHAINS 1
#endif

    struct SegSpec
    {
        SegKind kind;
        int koff, boff, aoff;   // slice starts inside the model's flat arrays (contiguous by construction)
        int frag_off = 0;       // start of this segment's packed B fragments inside acb200_model::frags (uint32 units)
        int bop_off = 0;        // start of this segment's tcgen05 B operands inside acb200_model::bops (uint32 units)
        int tm_off = -1;        // start of this segment's B operands of the TMEM-resident engine inside acb200_model::tmops (uint32 units)
    };

    // every segment type of every chain: X(kind, type)
#define ACB_FOR_EACH_SEG(X) \
    X(SEG_LEGACY_FULL, SegLegacyFull) X(SEG_ACNET_B4, SegAcnetB4) X(SEG_ACNET_B8, SegAcnetB8) X(SEG_ACNET_B18_A, SegAcnetB18A) \
    X(SEG_ACNET_B18_B, SegAcnetB18B) X(SEG_ARNET_FIRST, SegArnetFirst) X(SEG_ARNET_MID, SegArnetMid) X(SEG_ARNET_LAST, SegArnetLast) \
    X(SEG_LEGACY_A, SegLegacyA) X(SEG_LEGACY_B, SegLegacyB) X(SEG_ACNET_B8_A, SegAcnetB8A) X(SEG_ACNET_B8_B, SegAcnetB8B) X(SEG_ACNET_MID5, SegAcnetMid5)
}

struct acb200_model
{
    int family = 0, blocks = 0, features = 8;
    std::vector<float> k, b, a;
    std::vector<acbh::SegSpec> chain;
    // tensor-core engine: B fragments (split fp16) of every segment, concatenated; chain[i].frag_off indexes into it
    std::vector<uint32_t> frags;
    // tcgen05 engine: B operands (split fp16, no-swizzle K-major canonical layout), TC_B_WORDS_LAYER words per 3x3 conv
    std::vector<uint32_t> bops;
    // TMEM-resident engine (acb200_tm.cuh): B operands, TM_B_WORDS_LAYER words per 3x3 conv
    std::vector<uint32_t> tmops;
    unsigned long long uid = 0;
};

struct acb200_session
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // tensor engine implementation: 0 mma.sync (HMMA), 1 tcgen05 SS (UTCHMMA, maps in shared memory), 2 tcgen05 TMEM-resident (default; the
    // families it does not cover -- ARNet, segments deeper than eight convs -- run on mma.sync)
    int tensor_impl = 2;
    int sm_count = 0;
    int engine = 2;     // 0 exact FFMA, 1 tensor-core MMA, 2 auto: exact for every 2x pass but the last, tensor for the last
    int fuse = 1;       // colour split / chroma resize / merge inside the TMEM engine's segment kernels: 0 never, 1 8-bit RGB at 2x, 2 RGBA as well
    std::string error = "NO ERROR";
    // grow-only device scratch
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf src, dst, y[2], uv, map[2], feat, htab, vtab;
    Buf pin[3], pout[3];    // planar video frames: staged source / result planes (host entry)
    Buf wide[3];            // ArtCNN / FSRCNNX: feat + two ping-pong maps, [h][w][F] fp32
    Buf dhtab, dvtab;       // down-scaling contributor tables of the post-network luma resize (non-power-of-two factors)
    int dtab_in_w = 0, dtab_in_h = 0, dtab_out_w = 0, dtab_out_h = 0;
    int wide_smem_configured = 0;
    // device copies of models' packed fragments, keyed by acb200_model::uid
    std::map<unsigned long long, void*> dev_frags;
    std::map<unsigned long long, void*> dev_bops;
    std::map<unsigned long long, void*> dev_tmops;
    int tab_in_w = 0, tab_in_h = 0, tab_out_w = 0, tab_out_h = 0, tab_max_cnt = 0;
    int smem_configured = 0;
};

namespace acbh
{
    extern std::atomic<unsigned long long> g_launches;

    inline int fail(acb200_session* s, int code, const char* what, cudaError_t e = cudaSuccess)
    {
        if (s)
        {
            s->error = what;
            if (e != cudaSuccess) { s->error += ": "; s->error += cudaGetErrorString(e); }
        }
        return code;
    }
    // the dynamic shared memory opt-in of a kernel, once per kernel and device (the attribute belongs to the device's context): `mask` is
    // a static of the launching function template, one bit per device
    inline int smem_optin_once(acb200_session* s, const void* kernel, size_t bytes, std::atomic<unsigned long long>& mask)
    {
        const unsigned long long bit = 1ull << (s->device & 63);
        if (mask.load(std::memory_order_acquire) & bit) return ACB200_OK;
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return fail(s, ACB200_ECUDA, "cudaFuncSetAttribute(max dynamic smem)", e);
        mask.fetch_or(bit, std::memory_order_release);
        return ACB200_OK;
    }
    // returned by an engine's launcher for a segment it has no kernel for (the caller picks another engine); never leaves the library
    constexpr int ACB_SEG_UNSUPPORTED = -1000;
#define ACB_CUDA(s, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return acbh::fail((s), ACB200_ECUDA, #call, e__); } while (0)

    // grow-only scratch, allocated and freed in the order of stream `st` (the stream the buffer's users run on)
    int ensure(acb200_session* s, cudaStream_t st, acb200_session::Buf& b, size_t bytes);
    // device copy of one of a model's packed tables (uploaded once per session and model)
    int device_table(acb200_session* s, cudaStream_t st, std::map<unsigned long long, void*>& cache, unsigned long long uid, const std::vector<uint32_t>& host,
                     const char* what, const uint32_t** out);

    // what one segment launch works on
    struct SegLaunch
    {
        const void* src; int src_pitch;
        void* dst; int dst_pitch;
        int w, h, type;
        const float* map_in; float* map_out; float* feat;
        // Colour handling fused into the TMEM engine's segments (8-bit RGB, 2x; see TmParams): all null = luma plane in / luma plane out.
        // Only launch_seg_tm understands these; the caller checks seg_tm_chain_supported() before it sets them.
        const uint8_t* rgb_src = nullptr; int rgb_pitch = 0;
        uint8_t* uv_out = nullptr; const uint8_t* uv_in = nullptr; int uv_pitch = 0;
        uint8_t* y_out = nullptr; int y_pitch = 0;
        int uvc = 2;            // channels of the chroma plane: 2 for RGB, 3 (u, v, a) for RGBA
        const void* htab = nullptr; const void* vtab = nullptr;
        uint8_t* rgb_dst = nullptr; int rgb_dst_pitch = 0;
    };
    // one fused segment on one engine (each defined in its own TU)
    int launch_seg_ffma(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a);
    int launch_seg_mma(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a);
    int launch_seg_tc5(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a);
    int launch_seg_tm(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a);
    bool seg_tm_supported(const acb200_model& m);
    // every segment of the model's chain has a kernel on the TMEM engine (no per-segment fall-back to mma.sync would happen)
    bool seg_tm_chain_supported(const acb200_model& m);
    // ArtCNN / FSRCNNX: one 2x luma pass, one launch per layer (acb200_seg_wide.cu)
    int luma_pass_wide_any(acb200_session* s, cudaStream_t st, const acb200_model& m, const void* src, int src_pitch, void* dst, int dst_pitch,
                           int w, int h, int type, bool tensor);
}
