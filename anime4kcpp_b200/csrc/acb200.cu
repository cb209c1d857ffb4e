// Thin C-ABI CUDA layer of the B200 backend (see include/acb200.h for the contract and the
// reference interfaces each entry point replaces).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "acb200_internal.cuh"
#include "acb200_mma.cuh"       // packed-table layouts only: the engines' kernels are instantiated in their own translation units
#include "acb200_tc5.cuh"
#include "acb200_tm.cuh"
#include "acb200_wide_tc.cuh"
#include "acb200_pixel.cuh"

namespace acbh
{
    std::atomic<unsigned long long> g_launches{ 0 };

    // Scratch is allocated and freed in the order of the stream its users run on.  (Round-1 allocated on the session's own
    // stream even when the caller passed another one: a regrown buffer could be handed out again while kernels queued on the
    // caller's stream were still reading it.)
    int ensure(acb200_session* s, cudaStream_t st, acb200_session::Buf& b, size_t bytes)
    {
        if (b.cap >= bytes) return ACB200_OK;
        if (b.p) ACB_CUDA(s, cudaFreeAsync(b.p, st));
        b.p = nullptr; b.cap = 0;
        ACB_CUDA(s, cudaMallocAsync(&b.p, bytes, st));
        b.cap = bytes;
        return ACB200_OK;
    }
    int device_table(acb200_session* s, cudaStream_t st, std::map<unsigned long long, void*>& cache, unsigned long long uid, const std::vector<uint32_t>& host,
                     const char* what, const uint32_t** out)
    {
        auto it = cache.find(uid);
        if (it == cache.end())
        {
            // Bounded: a long-lived session that has seen many (possibly destroyed) models drops ALL of its cached tables once there are
            // 32 of them -- after the device has drained, since launches in flight may still read them -- and re-uploads what it meets
            // again (a table is a few KB; models are identified by uid, which is never reused).
            if (cache.size() >= 32)
            {
                ACB_CUDA(s, cudaDeviceSynchronize());
                for (auto& kv : cache) cudaFree(kv.second);
                cache.clear();
            }
            void* p = nullptr;
            ACB_CUDA(s, cudaMalloc(&p, host.size() * sizeof(uint32_t)));
            cudaError_t e = cudaMemcpyAsync(p, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(p); return fail(s, ACB200_ECUDA, what, e); }
            it = cache.emplace(uid, p).first;
        }
        *out = static_cast<const uint32_t*>(it->second);
        return ACB200_OK;
    }
}
using namespace acbh;

namespace
{
    bool expected_lengths(int family, int blocks, int& nk, int& nb, int& na, int F = 8)
    {
        switch (family)
        {
        case ACB200_FAMILY_ARTCNN: // core/include/AC/Core/Model/ArtCNN.hpp:33-36
            nk = F * 9 + F * F * 9 * (blocks + 1) + F * 4 * 9; nb = F * (blocks + 2) + 4; na = 0;
            return (F == 16 || F == 32) && blocks >= 1 && blocks <= 16;
        case ACB200_FAMILY_FSRCNNX: // core/include/AC/Core/Model/FSRCNNX.hpp:33-38
            nk = F * 25 + F * F * 9 * blocks + F * F + F * 4 * 9; nb = F * (blocks + 2) + 4; na = F * (blocks + 1);
            return (F == 8 || F == 16) && blocks >= 2 && blocks <= 16;
        case ACB200_FAMILY_ACNET_LEGACY: // core/include/AC/Core/Model/ACNet.hpp:34-37
            nk = 72 + 576 * blocks + 32; nb = 8 + 8 * blocks; na = 0;
            return blocks == 8;
        case ACB200_FAMILY_ACNET: // ACNet.hpp:78-83
            nk = 72 + 576 * blocks + 288; nb = 8 + 8 * blocks + 4; na = 8 * (blocks + 1);
            return blocks == 4 || blocks == 8 || blocks == 18;
        case ACB200_FAMILY_ARNET: // ARNet.hpp:32-37
            nk = 72 + 576 * blocks * 2 + 64 + 288; nb = 8 + 8 * (blocks * 2 + 1) + 4; na = 8 * (blocks + 1);
            return blocks >= 8 && (blocks % 4) == 0;
        default:
            return false;
        }
    }

    void build_chain(acb200_model& m)
    {
        m.chain.clear();
        if (m.family >= ACB200_FAMILY_ARTCNN) return;       // per-layer kernels (acb200_wide.cuh), no fused segments
        // A tile's halo grows by one pixel per 3x3 layer, so a fused segment recomputes (56 - 2l)^2 / T^2 of layer l per tile: 1.39x
        // over the eight tensor-core layers of ACNetLegacy in ONE segment (T = 40), 1.15x when the network is cut in two (T = 50
        // and 46) at the price of one [h][w][8] fp32 map through L2 / HBM.  ACB_SPLIT_CHAINS selects the two-segment chains.
        if (m.family == ACB200_FAMILY_ACNET_LEGACY)
        {
            if (ACB_SPLIT_CHAINS)
            {
                m.chain.push_back({ SEG_LEGACY_A, 0, 0, 0 });
                m.chain.push_back({ SEG_LEGACY_B, 72 + 576 * LEGACY_SPLIT, 8 + 8 * LEGACY_SPLIT, 0 });
            }
            else m.chain.push_back({ SEG_LEGACY_FULL, 0, 0, 0 });
        }
        else if (m.family == ACB200_FAMILY_ACNET)
        {
            if (m.blocks == 4) m.chain.push_back({ SEG_ACNET_B4, 0, 0, 0 });
            else if (m.blocks == 8 && ACB_SPLIT_CHAINS)
            {
                m.chain.push_back({ SEG_ACNET_B8_A, 0, 0, 0 });
                m.chain.push_back({ SEG_ACNET_B8_B, 72 + 576 * ACNET_SPLIT, 8 + 8 * ACNET_SPLIT, 8 + 8 * ACNET_SPLIT });
            }
            else if (m.blocks == 8) m.chain.push_back({ SEG_ACNET_B8, 0, 0, 0 });
            else if (ACB_SPLIT_CHAINS)
            {
                // 18 body convs as head + 4 | 5 | 5 | 4 + tail (T = 48, 46, 46, 46) instead of head + 9 | 9 + tail (T = 38, 36)
                m.chain.push_back({ SEG_ACNET_B8_A, 0, 0, 0 });
                constexpr int S1 = ACNET_SPLIT, S2 = S1 + 5, S3 = S2 + 5;
                static_assert(S3 + (8 - ACNET_SPLIT) == 18, "ACNet-B18: head + S | 5 | 5 | (8 - S) + tail");
                m.chain.push_back({ SEG_ACNET_MID5, 72 + 576 * S1, 8 + 8 * S1, 8 + 8 * S1 });
                m.chain.push_back({ SEG_ACNET_MID5, 72 + 576 * S2, 8 + 8 * S2, 8 + 8 * S2 });
                m.chain.push_back({ SEG_ACNET_B8_B, 72 + 576 * S3, 8 + 8 * S3, 8 + 8 * S3 });
            }
            else
            {
                m.chain.push_back({ SEG_ACNET_B18_A, 0, 0, 0 });
                m.chain.push_back({ SEG_ACNET_B18_B, 72 + 576 * 9, 8 + 8 * 9, 8 + 8 * 9 });
            }
        }
        else
        {
            // 2(B-1) body convs ahead of the last block (B = 8, 16, 32, 64: always 2 short of a multiple of 8 and of 4):
            // ARNET_SEG with the head, ARNET_SEG per middle segment, ARNET_SEG - 2 in the tail segment
            const int body = 2 * (m.blocks - 1);
            m.chain.push_back({ SEG_ARNET_FIRST, 0, 0, 0 });
            int c0 = ARNET_SEG;
            for (; c0 + ARNET_SEG - 2 < body; c0 += ARNET_SEG) m.chain.push_back({ SEG_ARNET_MID, 72 + 576 * c0, 8 + 8 * c0, (c0 / 2) * 8 });
            m.chain.push_back({ SEG_ARNET_LAST, 72 + 576 * c0, 8 + 8 * c0, (c0 / 2) * 8 });
        }
    }

    // ------------------------------------------------------------------------------------------------
    // Catmull-Rom contributor tables (stb_image_resize2 gather upsample as called from
    // core/src/ImageResize.cpp:167-272: centre (n+0.5)/scale, support 2, coefficients normalised, CLAMP edge
    // folded onto the border pixel).  Kernel polynomial: ImageResize.cpp:58-83 with b = 0, c = 1/2.
    // ------------------------------------------------------------------------------------------------
    float catmull_rom(float v)
    {
        volatile float x = std::fabs(v); // volatile: keep every intermediate rounded to fp32, no contraction
        auto poly3 = [](float x, float c0, float c1, float c2, float c3) {
            volatile float t = x * c3; t = c2 + t; t = x * t; t = c1 + t; t = x * t; t = c0 + t; return static_cast<float>(t);
        };
        if (x < 1.0f) return poly3(x, 1.0f, 0.0f, -2.5f, 1.5f);
        if (x < 2.0f) return poly3(x, 2.0f, -4.0f, 2.5f, -0.5f);
        return 0.0f;
    }
    bool make_contribs(std::vector<Contrib>& out, int in_size, int out_size)
    {
        out.resize(out_size);
        const float scale = static_cast<float>(out_size) / static_cast<float>(in_size);
        const float inv_scale = 1.0f / scale, out_radius = 2.0f * scale;
        for (int n = 0; n < out_size; n++)
        {
            volatile float out_center = static_cast<float>(n) + 0.5f;
            volatile float in_center = out_center * inv_scale;
            volatile float lo = out_center - out_radius; lo = lo * inv_scale;
            volatile float hi = out_center + out_radius; hi = hi * inv_scale;
            int first = static_cast<int>(std::floor(lo + 0.5f)), last = static_cast<int>(std::floor(hi - 0.5f));
            if (last < first) last = first;
            if (last - first > 10) last = first + 10;
            float raw[12];
            volatile float total = 0.0f;
            for (int i = 0; i <= last - first; i++)
            {
                volatile float pc = static_cast<float>(first + i) + 0.5f;
                raw[i] = catmull_rom(in_center - pc);
                total = total + raw[i];
            }
            volatile float fs = 1.0f / total;
            for (int i = 0; i <= last - first; i++) { volatile float t = raw[i] * fs; raw[i] = t; }
            int n0 = std::max(first, 0), n1 = std::min(last, in_size - 1);
            if (n1 < n0) n1 = n0 = (first < 0 ? 0 : in_size - 1);
            float c[12] = {};
            for (int i = 0; i <= last - first; i++) { int p = first + i; if (p >= n0 && p <= n1) c[p - n0] = raw[i]; }
            for (int i = 0; i <= last - first; i++)
            {
                int p = first + i;
                if (p < n0) { volatile float t = c[0] + raw[i]; c[0] = t; }
                else if (p > n1) { volatile float t = c[n1 - n0] + raw[i]; c[n1 - n0] = t; }
            }
            // Drop trailing zero taps so the device loops stay short.  Leading zeros are kept: the tiled kernels take the
            // first / last contributor of a tile as its source window, which needs n0 to be monotonic in n (an integer
            // up-scale such as 3x has phases whose outer taps are exactly zero).
            while (n1 > n0 && c[n1 - n0] == 0.0f) n1--;
            Contrib& o = out[n];
            o.n0 = n0; o.cnt = n1 - n0 + 1;
            if (o.cnt > 6) return false;
            for (int i = 0; i < 6; i++) o.c[i] = i < o.cnt ? c[i] : 0.0f;
        }
        return true;
    }
    // Down-scaling contributors (scale in [1/2, 1)): the filter stretched by 1/scale in input space, weights normalised,
    // out-of-range taps folded onto the edge pixel -- the same arithmetic, in the same order, as oracle/ac_oracle.c make_contribs.
    bool make_contribs_down(std::vector<ContribW>& out, int in_size, int out_size)
    {
        out.resize(out_size);
        const float scale = static_cast<float>(out_size) / static_cast<float>(in_size);
        if (!(scale < 1.0f) || scale < 0.5f) return false;
        const float inv_scale = 1.0f / scale;
        for (int n = 0; n < out_size; n++)
        {
            volatile float out_center = static_cast<float>(n) + 0.5f;
            volatile float in_center = out_center * inv_scale;
            volatile float reach = 2.0f * inv_scale;
            volatile float lo = in_center - reach, hi = in_center + reach;
            int first = static_cast<int>(std::floor(lo + 0.5f)), last = static_cast<int>(std::floor(hi - 0.5f));
            if (last < first) last = first;
            if (last - first > 10) last = first + 10;
            float raw[12];
            volatile float total = 0.0f;
            for (int i = 0; i <= last - first; i++)
            {
                volatile float pc = static_cast<float>(first + i) + 0.5f;
                volatile float d = in_center - pc; d = d * scale;
                raw[i] = catmull_rom(d);
                total = total + raw[i];
            }
            volatile float fs = 1.0f / total;
            for (int i = 0; i <= last - first; i++) { volatile float t = raw[i] * fs; raw[i] = t; }
            int n0 = std::max(first, 0), n1 = std::min(last, in_size - 1);
            if (n1 < n0) n1 = n0 = (first < 0 ? 0 : in_size - 1);
            float c[12] = {};
            for (int i = 0; i <= last - first; i++) { int p = first + i; if (p >= n0 && p <= n1) c[p - n0] = raw[i]; }
            for (int i = 0; i <= last - first; i++)
            {
                int p = first + i;
                if (p < n0) { volatile float t = c[0] + raw[i]; c[0] = t; }
                else if (p > n1) { volatile float t = c[n1 - n0] + raw[i]; c[n1 - n0] = t; }
            }
            ContribW& o = out[n];
            o.n0 = n0; o.cnt = n1 - n0 + 1;
            if (o.cnt > 10) return false;
            for (int i = 0; i < 10; i++) o.c[i] = i < o.cnt ? c[i] : 0.0f;
        }
        return true;
    }
}

namespace
{
    // ---- host-side packing of the split-fp16 B fragments (layout documented in acb200_mma.cuh) ------------------------
    uint32_t half_bits(float v) { return static_cast<uint32_t>(__half_as_ushort(__float2half_rn(v))); }
    void split_w(float w, uint32_t& hi, uint32_t& lo)
    {
        const __half h = __float2half_rn(w);
        hi = static_cast<uint32_t>(__half_as_ushort(h));
        lo = half_bits(w - __half2float(h));
    }
    // W: [cout][9 taps][8 cin] fp32 (reference layout); cout <= 8, missing output channels are zero
    void pack_conv3x3(const float* W, int cout, std::vector<uint32_t>& out)
    {
        const size_t base = out.size();
        out.resize(base + FRAG_WORDS_3X3, 0u);
        for (int lane = 0; lane < 32; lane++)
        {
            const int n = lane >> 2, t = lane & 3;      // B fragment: column n (= cout), rows k = 2t, 2t+1 (+8)
            for (int s = 0; s < 5; s++)
                for (int half = 0; half < (s < 4 ? 2 : 1); half++)
                {
                    const int tap = 2 * s + half;
                    uint32_t h0 = 0, l0 = 0, h1 = 0, l1 = 0;
                    if (n < cout)
                    {
                        split_w(W[(n * 9 + tap) * 8 + 2 * t], h0, l0);
                        split_w(W[(n * 9 + tap) * 8 + 2 * t + 1], h1, l1);
                    }
                    const int reg = 2 * s + half;       // k-step 4 has a single register (index 8)
                    out[base + reg * 32 + lane] = h0 | (h1 << 16);
                    out[base + (9 + reg) * 32 + lane] = l0 | (l1 << 16);
                }
        }
    }
    // W: [cout 8][cin 8] fp32 (the ARNet 1x1)
    void pack_conv1x1(const float* W, std::vector<uint32_t>& out)
    {
        const size_t base = out.size();
        out.resize(base + FRAG_WORDS_1X1, 0u);
        for (int lane = 0; lane < 32; lane++)
        {
            const int n = lane >> 2, t = lane & 3;
            uint32_t h0, l0, h1, l1;
            split_w(W[n * 8 + 2 * t], h0, l0);
            split_w(W[n * 8 + 2 * t + 1], h1, l1);
            out[base + lane] = h0 | (h1 << 16);
            out[base + 32 + lane] = l0 | (l1 << 16);
        }
    }
    // tcgen05 B operand of one 3x3 conv: for dy in 0..2 a [N = 48][K = 16] fp16 matrix, K-major canonical layout
    // (element (n, k) at byte (k/8)*768 + n*16 + (k%8)*2).  Row n = dx*16 + j: j < 8 -> cout j, K chunk 0 (times a_hi) and chunk 1
    // (times a_lo) both carry w_hi; j >= 8 -> cout j-8, chunk 0 carries w_lo, chunk 1 is zero.
    void pack_bop3x3(const float* W, int cout, std::vector<uint32_t>& out)
    {
        const size_t base = out.size();
        out.resize(base + TC_B_WORDS_LAYER, 0u);
        uint16_t* h = reinterpret_cast<uint16_t*>(out.data() + base);
        for (int dy = 0; dy < 3; dy++)
            for (int dx = 0; dx < 3; dx++)
                for (int co = 0; co < cout; co++)
                    for (int ci = 0; ci < 8; ci++)
                    {
                        uint32_t hi, lo;
                        split_w(W[(co * 9 + dy * 3 + dx) * 8 + ci], hi, lo);
                        uint16_t* m = h + dy * (TC_B_BYTES_DY / 2);
                        const int n_hi = dx * 16 + co, n_lo = dx * 16 + 8 + co;
                        m[0 * (TC_N * 8) + n_hi * 8 + ci] = static_cast<uint16_t>(hi);     // chunk 0 (a_hi) x w_hi
                        m[1 * (TC_N * 8) + n_hi * 8 + ci] = static_cast<uint16_t>(hi);     // chunk 1 (a_lo) x w_hi
                        m[0 * (TC_N * 8) + n_lo * 8 + ci] = static_cast<uint16_t>(lo);     // chunk 0 (a_hi) x w_lo
                    }
    }
    // B operands of one 3x3 conv for the TMEM-resident engine (acb200_tm.cuh): for every alignment al (dx = al - 1) TWO [N = 24][K = 16]
    // fp16 matrices in the no-swizzle K-major canonical layout (element (n, k) at byte (k/8)*384 + n*16 + (k%8)*2).  Row n = jo*8 + co:
    // jo selects the output row relative to the input row (o = r - 1 + jo, i.e. dy = 1 - jo), co the output channel.  First matrix: both
    // K chunks (a_hi, a_lo) carry w_hi; second matrix: chunk 0 carries w_lo, chunk 1 is zero.  Both accumulate into the same columns.
    void pack_bop_tm(const float* W, int cout, std::vector<uint32_t>& out)
    {
        const size_t base = out.size();
        out.resize(base + TM_B_WORDS_LAYER, 0u);
        uint16_t* h = reinterpret_cast<uint16_t*>(out.data() + base);
        for (int al = 0; al < 3; al++)
            for (int jo = 0; jo < 3; jo++)
                for (int co = 0; co < cout; co++)
                    for (int ci = 0; ci < 8; ci++)
                    {
                        const int dy = 1 - jo, dx = al - 1;
                        uint32_t hi, lo;
                        split_w(W[(co * 9 + (dy + 1) * 3 + (dx + 1)) * 8 + ci], hi, lo);
                        uint16_t* m1 = h + al * (TM_B_BYTES_AL / 2);
                        uint16_t* m2 = m1 + TM_B_BYTES_HALF / 2;
                        const int n = jo * 8 + co;
                        m1[0 * (24 * 8) + n * 8 + ci] = static_cast<uint16_t>(hi);
                        m1[1 * (24 * 8) + n * 8 + ci] = static_cast<uint16_t>(hi);
                        m2[0 * (24 * 8) + n * 8 + ci] = static_cast<uint16_t>(lo);
                    }
    }
    template<class S>
    void pack_segment(acb200_model& m, SegSpec& sp)
    {
        {
            sp.tm_off = static_cast<int>(m.tmops.size());
            const float* kk = m.k.data() + sp.koff + (S::HEAD ? 72 : 0);
            for (int i = 0; i < S::NCONV; i++, kk += 576) pack_bop_tm(kk, 8, m.tmops);
            if (S::TAIL && S::FAM == ACB200_FAMILY_ARNET)
            {
                // PReLU conv, residual conv, [1x1: 64 weights, on the CUDA cores], pixel-shuffle conv
                pack_bop_tm(kk, 8, m.tmops);
                pack_bop_tm(kk + 576, 8, m.tmops);
                pack_bop_tm(kk + 1152 + 64, 4, m.tmops);
            }
            else if (S::TAIL) pack_bop_tm(kk, S::FAM == ACB200_FAMILY_ACNET_LEGACY ? 8 : 4, m.tmops);
        }
        {
            sp.bop_off = static_cast<int>(m.bops.size());
            const float* kk = m.k.data() + sp.koff + (S::HEAD ? 72 : 0);
            for (int i = 0; i < S::NCONV; i++, kk += 576) pack_bop3x3(kk, 8, m.bops);
            if (S::TAIL)
            {
                if (S::FAM == ACB200_FAMILY_ACNET_LEGACY) pack_bop3x3(kk, 8, m.bops);
                else if (S::FAM == ACB200_FAMILY_ACNET) pack_bop3x3(kk, 4, m.bops);
                else { pack_bop3x3(kk, 8, m.bops); pack_bop3x3(kk + 576, 8, m.bops); pack_bop3x3(kk + 1152 + 64, 4, m.bops); }
            }
        }
        sp.frag_off = static_cast<int>(m.frags.size());
        const float* k = m.k.data() + sp.koff + (S::HEAD ? 72 : 0);
        for (int i = 0; i < S::NCONV; i++, k += 576) pack_conv3x3(k, 8, m.frags);
        if (!S::TAIL) return;
        if (S::FAM == ACB200_FAMILY_ACNET_LEGACY) pack_conv3x3(k, 8, m.frags);
        else if (S::FAM == ACB200_FAMILY_ACNET) pack_conv3x3(k, 4, m.frags);
        else
        {
            pack_conv3x3(k, 8, m.frags);            // PReLU conv of the last block
            pack_conv3x3(k + 576, 8, m.frags);      // residual conv
            pack_conv1x1(k + 1152, m.frags);        // 1x1
            pack_conv3x3(k + 1152 + 64, 4, m.frags);// pixel-shuffle conv
        }
    }
    // tcgen05 B operand of one F -> F 3x3 conv of the wide families (acb200_wide_tc.cuh): for every (tap, 8-channel chunk) a
    // [N = 2F][K = 16] fp16 matrix in the K-major canonical layout (element (n, k) at byte (k/8)*N*16 + n*16 + (k%8)*2).
    // Row n < F: cout n, both K chunks carry w_hi (they multiply a_hi and a_lo); row F + n: K chunk 0 carries w_lo, chunk 1 is zero.
    void pack_bop_wide(const float* W, int F, std::vector<uint32_t>& out)
    {
        const int NCH = F / 8, N = 2 * F, tap_bytes = 2 * N * 16;
        const size_t base = out.size();
        out.resize(base + static_cast<size_t>(9) * NCH * tap_bytes / 4, 0u);
        uint16_t* h = reinterpret_cast<uint16_t*>(out.data() + base);
        for (int t = 0; t < 9; t++)
            for (int c = 0; c < NCH; c++)
            {
                uint16_t* blk = h + static_cast<size_t>(t * NCH + c) * tap_bytes / 2;
                for (int n = 0; n < F; n++)
                    for (int k = 0; k < 8; k++)
                    {
                        uint32_t hi, lo;
                        split_w(W[(static_cast<size_t>(n) * 9 + t) * F + c * 8 + k], hi, lo);
                        blk[0 * N * 8 + n * 8 + k] = static_cast<uint16_t>(hi);           // x a_hi
                        blk[1 * N * 8 + n * 8 + k] = static_cast<uint16_t>(hi);           // x a_lo
                        blk[0 * N * 8 + (F + n) * 8 + k] = static_cast<uint16_t>(lo);     // x a_hi
                    }
            }
    }
    void pack_model(acb200_model& m)
    {
        m.frags.clear();
        m.bops.clear();
        m.tmops.clear();
        if (m.family >= ACB200_FAMILY_ARTCNN && m.features >= 16)
        {
            const int F = m.features, ks = m.family == ACB200_FAMILY_ARTCNN ? 3 : 5;
            const int convs = m.family == ACB200_FAMILY_ARTCNN ? m.blocks + 1 : m.blocks;
            for (int l = 0; l < convs; l++) pack_bop_wide(m.k.data() + F * ks * ks + static_cast<size_t>(F) * F * 9 * l, F, m.bops);
        }
        for (SegSpec& sp : m.chain)
            switch (sp.kind)
            {
            case SEG_LEGACY_FULL: pack_segment<SegLegacyFull>(m, sp); break;
            case SEG_ACNET_B4: pack_segment<SegAcnetB4>(m, sp); break;
            case SEG_ACNET_B8: pack_segment<SegAcnetB8>(m, sp); break;
            case SEG_ACNET_B18_A: pack_segment<SegAcnetB18A>(m, sp); break;
            case SEG_ACNET_B18_B: pack_segment<SegAcnetB18B>(m, sp); break;
            case SEG_ARNET_FIRST: pack_segment<SegArnetFirst>(m, sp); break;
            case SEG_ARNET_MID: pack_segment<SegArnetMid>(m, sp); break;
            case SEG_ARNET_LAST: pack_segment<SegArnetLast>(m, sp); break;
            case SEG_LEGACY_A: pack_segment<SegLegacyA>(m, sp); break;
            case SEG_LEGACY_B: pack_segment<SegLegacyB>(m, sp); break;
            case SEG_ACNET_B8_A: pack_segment<SegAcnetB8A>(m, sp); break;
            case SEG_ACNET_B8_B: pack_segment<SegAcnetB8B>(m, sp); break;
            case SEG_ACNET_MID5: pack_segment<SegAcnetMid5>(m, sp); break;
            }
    }
    std::atomic<unsigned long long> g_model_uid{ 1 };
}

namespace
{
    size_t pitch_of(int w, int c, int es) { return (static_cast<size_t>(w) * c * es + 255) & ~static_cast<size_t>(255); }

    // Staging a host image on the device: when the host rows are 16-byte multiples and not much wider than a line, the device copy
    // keeps the HOST pitch, so the transfer is one contiguous block instead of a row-by-row (2-D) copy -- a 2-D copy of short rows
    // (960-byte chroma rows, 1920-byte luma rows) reaches well under the PCIe rate.  Every kernel takes arbitrary pitches.
    // `to_host`: the block is WRITTEN to host memory, so it may only be used for tight rows -- the bytes between the rows of a strided
    // view belong to someone else.  Towards the device a padded host pitch is fine (the padding is read, never used).
    size_t staging_pitch(int host_stride, size_t line, int w, int c, int es, bool to_host)
    {
        static const bool off = [] { const char* e = std::getenv("ACB200_CONTIG_COPY"); return e && e[0] == '0'; }();     // A/B switch
        const size_t hs = static_cast<size_t>(host_stride);
        const bool ok = to_host ? (hs == line && line % 4 == 0) : (hs >= line && hs % 16 == 0 && hs <= line + line / 4 + 256);
        return (!off && ok) ? hs : pitch_of(w, c, es);
    }
    // rows of `line` bytes between host memory (pitch hp) and device memory (pitch dp): one block when the pitches agree
    cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t line, int rows, cudaMemcpyKind kind, cudaStream_t st)
    {
        // (towards the host only tight rows: equal pitches alone do not make the bytes between the rows ours to write)
        if (dpitch == spitch && rows > 0 && (kind == cudaMemcpyHostToDevice || spitch == line))
            return cudaMemcpyAsync(dst, src, spitch * static_cast<size_t>(rows - 1) + line, kind, st);
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, line, rows, kind, st);
    }

    // what the fused colour path hands to the first and last segment of a chain (see TmParams)
    struct FusedColour
    {
        const uint8_t* rgb_src; int rgb_pitch;
        uint8_t* uv; int uv_pitch;
        int uvc;                            // channels of the chroma plane (2: RGB, 3: RGBA)
        uint8_t* y; int y_pitch;            // quantised luma plane, written by the head segment for a tail that adds the source luma (ACNet); else null
        const void* htab; const void* vtab;
        uint8_t* rgb_dst; int rgb_dst_pitch;
    };

    int luma_pass(acb200_session* s, cudaStream_t st, const acb200_model& m, const void* src, int src_pitch,
                  void* dst, int dst_pitch, int w, int h, int type, bool tensor, const FusedColour* fz = nullptr)
    {
        if (m.family >= ACB200_FAMILY_ARTCNN) return luma_pass_wide_any(s, st, m, src, src_pitch, dst, dst_pitch, w, h, type, tensor);
        float* maps[2] = { nullptr, nullptr };
        float* feat = nullptr;
        if (m.chain.size() > 1)
        {
            const size_t bytes = static_cast<size_t>(w) * h * 8 * sizeof(float);
            int rc;
            if ((rc = ensure(s, st, s->map[0], bytes)) != ACB200_OK) return rc;
            if ((rc = ensure(s, st, s->map[1], bytes)) != ACB200_OK) return rc;
            maps[0] = static_cast<float*>(s->map[0].p); maps[1] = static_cast<float*>(s->map[1].p);
            if (m.family == ACB200_FAMILY_ARNET)
            {
                if ((rc = ensure(s, st, s->feat, bytes)) != ACB200_OK) return rc;
                feat = static_cast<float*>(s->feat.p);
            }
        }
        int cur = 0;
        for (size_t i = 0; i < m.chain.size(); i++)
        {
            const SegSpec& sp = m.chain[i];
            const float* in = maps[cur];
            float* out = maps[cur ^ 1];
            // map / feat pointers by segment shape: a segment without a head reads `in`, one without a tail writes `out`
            const bool head = sp.kind == SEG_LEGACY_FULL || sp.kind == SEG_ACNET_B4 || sp.kind == SEG_ACNET_B8 || sp.kind == SEG_ACNET_B18_A || sp.kind == SEG_ARNET_FIRST ||
                              sp.kind == SEG_LEGACY_A || sp.kind == SEG_ACNET_B8_A;
            const bool tail = sp.kind == SEG_LEGACY_FULL || sp.kind == SEG_ACNET_B4 || sp.kind == SEG_ACNET_B8 || sp.kind == SEG_ACNET_B18_B || sp.kind == SEG_ARNET_LAST ||
                              sp.kind == SEG_LEGACY_B || sp.kind == SEG_ACNET_B8_B;
            SegLaunch a;
            a.src = src; a.src_pitch = src_pitch; a.dst = dst; a.dst_pitch = dst_pitch; a.w = w; a.h = h; a.type = type;
            a.map_in = head ? nullptr : in; a.map_out = tail ? nullptr : out; a.feat = feat;
            int rc;
            if (fz)
            {
                // fused colour handling: TMEM engine only (the caller has checked that every segment has a kernel there)
                a.uv_pitch = fz->uv_pitch; a.uvc = fz->uvc;
                if (head) { a.rgb_src = fz->rgb_src; a.rgb_pitch = fz->rgb_pitch; a.uv_out = fz->uv; a.y_out = fz->y; a.y_pitch = fz->y_pitch; }
                else { a.src = fz->y; a.src_pitch = fz->y_pitch; }
                if (tail) { a.uv_in = fz->uv; a.htab = fz->htab; a.vtab = fz->vtab; a.rgb_dst = fz->rgb_dst; a.rgb_dst_pitch = fz->rgb_dst_pitch; }
                if ((rc = launch_seg_tm(s, st, m, sp, a)) != ACB200_OK) return rc == ACB_SEG_UNSUPPORTED ? fail(s, ACB200_EINVAL, "fused colour path: segment without a TMEM kernel") : rc;
                cur ^= 1;
                continue;
            }
            if (!tensor) rc = launch_seg_ffma(s, st, m, sp, a);
            else if (s->tensor_impl == 1) rc = launch_seg_tc5(s, st, m, sp, a);
            else
            {
                // the TMEM-resident engine where it has a kernel for the segment, mma.sync for the rest (ARNet, segments deeper than 8 convs)
                rc = (s->tensor_impl == 2 && seg_tm_supported(m)) ? launch_seg_tm(s, st, m, sp, a) : ACB_SEG_UNSUPPORTED;
                if (rc == ACB_SEG_UNSUPPORTED) rc = launch_seg_mma(s, st, m, sp, a);
            }
            if (rc != ACB200_OK) return rc;
            cur ^= 1;
        }
        return ACB200_OK;
    }

    int ensure_tables(acb200_session* s, cudaStream_t st, int w, int h, int ow, int oh)
    {
        if (s->tab_in_w == w && s->tab_in_h == h && s->tab_out_w == ow && s->tab_out_h == oh) return ACB200_OK;
        std::vector<Contrib> ht, vt;
        if (!make_contribs(ht, w, ow) || !make_contribs(vt, h, oh)) return fail(s, ACB200_EINVAL, "resize: unsupported scale");
        int rc;
        if ((rc = ensure(s, st, s->htab, ht.size() * sizeof(Contrib))) != ACB200_OK) return rc;
        if ((rc = ensure(s, st, s->vtab, vt.size() * sizeof(Contrib))) != ACB200_OK) return rc;
        // pageable -> device: the copy is staged before the call returns, so the vectors may die
        ACB_CUDA(s, cudaMemcpyAsync(s->htab.p, ht.data(), ht.size() * sizeof(Contrib), cudaMemcpyHostToDevice, st));
        ACB_CUDA(s, cudaMemcpyAsync(s->vtab.p, vt.data(), vt.size() * sizeof(Contrib), cudaMemcpyHostToDevice, st));
        ACB_CUDA(s, cudaStreamSynchronize(st));
        s->tab_in_w = w; s->tab_in_h = h; s->tab_out_w = ow; s->tab_out_h = oh;
        s->tab_max_cnt = 0;
        for (const Contrib& k : ht) s->tab_max_cnt = std::max(s->tab_max_cnt, k.cnt);
        for (const Contrib& k : vt) s->tab_max_cnt = std::max(s->tab_max_cnt, k.cnt);
        return ACB200_OK;
    }

    bool valid_type(int t) { return t == ACB200_UINT8 || t == ACB200_UINT16 || t == ACB200_FLOAT16 || t == ACB200_FLOAT32; }

    // factor -> number of 2x passes; only powers of two (fxy == 1, Processor.cpp:204-205)
    int passes_for(double factor)
    {
        for (int p = 1; p <= 6; p++) if (factor == static_cast<double>(1 << p)) return p;
        return 0;
    }

    // Processor.cpp:203-204: power = factor > 2 ? ceilLog2(factor) : 1 passes of 2x, then (factor not a power of two) a luma
    // down-scale by fxy = factor / 2^power in (1/2, 1); the result is int(w * factor) x int(h * factor) (ImageResize.cpp:153-154).
    struct FactorPlan { int power = 0, dw = 0, dh = 0; };
    bool plan_factor(double factor, int w, int h, FactorPlan& p)
    {
        p.power = passes_for(factor);
        if (p.power) { p.dw = w << p.power; p.dh = h << p.power; return true; }
        if (!(factor >= 1.0) || factor > 64.0) return false;
        p.power = 1;
        while (static_cast<double>(1 << p.power) < factor) p.power++;
        p.dw = static_cast<int>(w * factor); p.dh = static_cast<int>(h * factor);
        return p.dw > 0 && p.dh > 0;
    }
    int ensure_down_tables(acb200_session* s, cudaStream_t st, int w, int h, int ow, int oh)
    {
        if (s->dtab_in_w == w && s->dtab_in_h == h && s->dtab_out_w == ow && s->dtab_out_h == oh) return ACB200_OK;
        std::vector<ContribW> ht, vt;
        if (!make_contribs_down(ht, w, ow) || !make_contribs_down(vt, h, oh)) return fail(s, ACB200_EINVAL, "resize: unsupported down-scale");
        int rc;
        if ((rc = ensure(s, st, s->dhtab, ht.size() * sizeof(ContribW))) != ACB200_OK) return rc;
        if ((rc = ensure(s, st, s->dvtab, vt.size() * sizeof(ContribW))) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaMemcpyAsync(s->dhtab.p, ht.data(), ht.size() * sizeof(ContribW), cudaMemcpyHostToDevice, st));
        ACB_CUDA(s, cudaMemcpyAsync(s->dvtab.p, vt.data(), vt.size() * sizeof(ContribW), cudaMemcpyHostToDevice, st));
        ACB_CUDA(s, cudaStreamSynchronize(st));
        s->dtab_in_w = w; s->dtab_in_h = h; s->dtab_out_w = ow; s->dtab_out_h = oh;
        return ACB200_OK;
    }

    // The whole Processor::process on device-resident planes (Processor.cpp:199-276).
    int process_on_device(acb200_session* s, const acb200_model* m, cudaStream_t st,
                          const void* d_src, int w, int h, int c, int src_pitch, int type, const FactorPlan& plan, void* d_dst, int dst_pitch)
    {
        const int power = plan.power;
        const bool down = plan.dw != (w << power) || plan.dh != (h << power);
        const int es = type & 0xff;
        const dim3 blk(32, 8);
        const void* cur = d_src;
        int cur_pitch = src_pitch, cw = w, ch = h, rc;
        // 8-bit RGB, exactly 2x, TMEM engine, chains of two or more segments: colour split inside the first segment's tile load, chroma
        // resize + merge inside the last segment's tail -- two launches per frame, no Y plane, bit-identical to the separate kernels
        // (RGB results leave as 16-bit stores; RGBA pixels are read and written as 32-bit words)
        const uintptr_t align_bits = (reinterpret_cast<uintptr_t>(d_dst) | static_cast<uintptr_t>(dst_pitch)) |
                                     (c == 4 ? (reinterpret_cast<uintptr_t>(d_src) | static_cast<uintptr_t>(src_pitch)) : 0);
        if (((c == 3 && s->fuse >= 1) || (c == 4 && s->fuse >= 2)) && type == ACB200_UINT8 && power == 1 && !down && s->engine != 0 && s->tensor_impl == 2 && m->chain.size() >= 2 &&
            seg_tm_chain_supported(*m) && (align_bits & (c == 4 ? 3 : 1)) == 0)
        {
            const size_t uvp = pitch_of(w, c - 1, 1);
            if ((rc = ensure(s, st, s->uv, uvp * h)) != ACB200_OK) return rc;
            if ((rc = ensure_tables(s, st, w, h, 2 * w, 2 * h)) != ACB200_OK) return rc;
            // ACNet's tail adds the nearest-upsampled source luma (Common.hpp:290-342): the head segment leaves the quantised Y plane for it
            const bool needs_y = m->family != ACB200_FAMILY_ACNET_LEGACY;
            const size_t yp = pitch_of(w, 1, 1);
            if (needs_y && (rc = ensure(s, st, s->y[0], yp * h)) != ACB200_OK) return rc;
            if (s->tab_max_cnt <= 4)
            {
                const FusedColour fz{ static_cast<const uint8_t*>(d_src), src_pitch, static_cast<uint8_t*>(s->uv.p), static_cast<int>(uvp), c - 1,
                                      needs_y ? static_cast<uint8_t*>(s->y[0].p) : nullptr, static_cast<int>(yp),
                                      s->htab.p, s->vtab.p, static_cast<uint8_t*>(d_dst), dst_pitch };
                return luma_pass(s, st, *m, nullptr, 0, nullptr, 0, w, h, type, true, &fz);
            }
        }
        if (c > 1)
        {
            const size_t yp = pitch_of(w, 1, es), uvp = pitch_of(w, c - 1, es);
            if ((rc = ensure(s, st, s->y[0], yp * h)) != ACB200_OK) return rc;
            if ((rc = ensure(s, st, s->uv, uvp * h)) != ACB200_OK) return rc;
            if (type == ACB200_UINT8 && c == 3 && ((reinterpret_cast<uintptr_t>(d_src) | static_cast<uintptr_t>(src_pitch)) & 3) == 0)
                rgb2yuv_u8x4_kernel<<<dim3((w + 127) / 128, (h + 7) / 8), blk, 0, st>>>(static_cast<const uint8_t*>(d_src), src_pitch, w, h,
                    static_cast<uint8_t*>(s->y[0].p), static_cast<int>(yp), static_cast<uint8_t*>(s->uv.p), static_cast<int>(uvp));
            else
                rgb2yuv_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), blk, 0, st>>>(d_src, src_pitch, w, h, c, type, s->y[0].p, static_cast<int>(yp), 1, s->uv.p, static_cast<int>(uvp), c - 1);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
            cur = s->y[0].p; cur_pitch = static_cast<int>(yp);
        }
        int slot = 1;
        for (int i = 0; i < power; i++)
        {
            const int nw = cw * 2, nh = ch * 2;
            const bool to_dst = (c == 1) && (i == power - 1) && !down;
            void* out; int out_pitch;
            if (to_dst) { out = d_dst; out_pitch = dst_pitch; }
            else
            {
                const size_t p = pitch_of(nw, 1, es);
                if ((rc = ensure(s, st, s->y[slot], p * nh)) != ACB200_OK) return rc;
                out = s->y[slot].p; out_pitch = static_cast<int>(p);
                slot ^= 1;
            }
            const bool tensor = s->engine == 1 || (s->engine == 2 && i == power - 1);
            if ((rc = luma_pass(s, st, *m, cur, cur_pitch, out, out_pitch, cw, ch, type, tensor)) != ACB200_OK) return rc;
            cur = out; cur_pitch = out_pitch; cw = nw; ch = nh;
        }
        if (down)
        {
            // gray: resize(out, dst, 0, 0) straight into dst; colour: resize(out, out, fxy, fxy) into the other luma plane
            if ((rc = ensure_down_tables(s, st, cw, ch, plan.dw, plan.dh)) != ACB200_OK) return rc;
            void* out; int out_pitch;
            if (c == 1) { out = d_dst; out_pitch = dst_pitch; }
            else
            {
                const size_t p = pitch_of(plan.dw, 1, es);
                if ((rc = ensure(s, st, s->y[slot], p * plan.dh)) != ACB200_OK) return rc;
                out = s->y[slot].p; out_pitch = static_cast<int>(p);
            }
            resize_wide_kernel<<<dim3((plan.dw + 31) / 32, (plan.dh + 7) / 8), blk, 0, st>>>(cur, cur_pitch, 1, type,
                static_cast<const ContribW*>(s->dhtab.p), static_cast<const ContribW*>(s->dvtab.p), out, plan.dw, plan.dh, out_pitch);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
            cur = out; cur_pitch = out_pitch; cw = plan.dw; ch = plan.dh;
        }
        if (c > 1)
        {
            if ((rc = ensure_tables(s, st, w, h, cw, ch)) != ACB200_OK) return rc;
            const int uv_pitch = static_cast<int>(pitch_of(w, c - 1, es));
            const Contrib* ht = static_cast<const Contrib*>(s->htab.p);
            const Contrib* vt = static_cast<const Contrib*>(s->vtab.p);
            const dim3 cgrid((cw + CM_OW - 1) / CM_OW, (ch + CM_OH - 1) / CM_OH);
            if (type == ACB200_UINT8 && c == 3 && s->tab_max_cnt <= 4)
                chroma_merge_u8_kernel<3><<<cgrid, CM_THREADS, 0, st>>>(static_cast<const uint8_t*>(cur), cur_pitch, static_cast<const uint8_t*>(s->uv.p), uv_pitch,
                                                                         w, h, ht, vt, cw, ch, static_cast<uint8_t*>(d_dst), dst_pitch);
            else if (type == ACB200_UINT8 && c == 4 && s->tab_max_cnt <= 4)
                chroma_merge_u8_kernel<4><<<cgrid, CM_THREADS, 0, st>>>(static_cast<const uint8_t*>(cur), cur_pitch, static_cast<const uint8_t*>(s->uv.p), uv_pitch,
                                                                         w, h, ht, vt, cw, ch, static_cast<uint8_t*>(d_dst), dst_pitch);
            else
                chroma_merge_kernel<<<dim3((cw + 31) / 32, (ch + 7) / 8), blk, 0, st>>>(cur, cur_pitch, s->uv.p, uv_pitch, ht, vt, cw, ch, c, type, d_dst, dst_pitch);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
        }
        return ACB200_OK;
    }

    int frame_stride(const acb200_plane& p, int es) { const int line = p.width * p.channel * es; return p.stride < line ? line : p.stride; }

    // cli/src/Main.cpp:183-206 on device-resident planes: plane 0 -> [shl] network [shr]; planes 1.. -> Catmull-Rom resize.
    int process_frame_on_device(acb200_session* s, const acb200_model* m, cudaStream_t st, const acb200_plane* src, const acb200_plane* dst,
                                int planes, int type, int shift, const FactorPlan& plan)
    {
        const int es = type & 0xff;
        const dim3 blk(32, 8);
        const bool integer = type == ACB200_UINT8 || type == ACB200_UINT16;
        const int sh = integer ? shift : 0;
        int rc;
        // ---- luma ----------------------------------------------------------------------------------------------------------
        const void* cur = src[0].data;
        int cur_pitch = frame_stride(src[0], es), cw = src[0].width, ch = src[0].height;
        int slot = 0;
        if (sh)
        {
            const size_t p = pitch_of(cw, 1, es);
            if ((rc = ensure(s, st, s->y[slot], p * ch)) != ACB200_OK) return rc;
            shift_kernel<<<dim3((cw + 31) / 32, (ch + 7) / 8), blk, 0, st>>>(cur, cur_pitch, s->y[slot].p, static_cast<int>(p), cw, ch, es, sh, 1);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
            cur = s->y[slot].p; cur_pitch = static_cast<int>(p);
        }
        // processor->process(srcy, dsty, factor): the 1-channel driver (passes, and the luma down-scale of non-2^k factors).  Its
        // intermediates start in y[1], so the shifted copy in y[0] is consumed before it can be overwritten.
        const int dpitch = frame_stride(dst[0], es);
        if ((rc = process_on_device(s, m, st, cur, cw, ch, 1, cur_pitch, type, plan, dst[0].data, dpitch)) != ACB200_OK) return rc;
        cw = plan.dw; ch = plan.dh; cur_pitch = dpitch;
        if (sh)
        {
            shift_kernel<<<dim3((cw + 31) / 32, (ch + 7) / 8), blk, 0, st>>>(dst[0].data, cur_pitch, dst[0].data, cur_pitch, cw, ch, es, sh, 0);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
        }
        // ---- chroma --------------------------------------------------------------------------------------------------------
        for (int i = 1; i < planes; i++)
        {
            const acb200_plane& a = src[i];
            const acb200_plane& b = dst[i];
            if ((rc = ensure_tables(s, st, a.width, a.height, b.width, b.height)) != ACB200_OK) return rc;
            if (integer && s->tab_max_cnt <= 4)
            {
                // tiled integer kernel; two equally shaped planes (the U and V of a planar frame) share one launch
                const bool pair = i + 1 < planes && src[i + 1].width == a.width && src[i + 1].height == a.height && src[i + 1].channel == a.channel &&
                                  dst[i + 1].width == b.width && dst[i + 1].height == b.height;
                ResizePlanes pl;
                for (int k = 0; k < 2; k++)
                {
                    const int j = (k == 1 && pair) ? i + 1 : i;
                    pl.src[k] = src[j].data; pl.dst[k] = dst[j].data;
                    pl.src_pitch[k] = frame_stride(src[j], es); pl.dst_pitch[k] = frame_stride(dst[j], es);
                }
                const dim3 grid((b.width + CM_OW - 1) / CM_OW, (b.height + CM_OH - 1) / CM_OH, pair ? 2 : 1);
                const Contrib* ht = static_cast<const Contrib*>(s->htab.p);
                const Contrib* vt = static_cast<const Contrib*>(s->vtab.p);
                if (es == 1 && a.channel == 1) resize_tile_kernel<uint8_t, 1><<<grid, CM_THREADS, 0, st>>>(pl, a.width, a.height, ht, vt, b.width, b.height);
                else if (es == 1) resize_tile_kernel<uint8_t, 2><<<grid, CM_THREADS, 0, st>>>(pl, a.width, a.height, ht, vt, b.width, b.height);
                else if (a.channel == 1) resize_tile_kernel<uint16_t, 1><<<grid, CM_THREADS, 0, st>>>(pl, a.width, a.height, ht, vt, b.width, b.height);
                else resize_tile_kernel<uint16_t, 2><<<grid, CM_THREADS, 0, st>>>(pl, a.width, a.height, ht, vt, b.width, b.height);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                ACB_CUDA(s, cudaGetLastError());
                if (pair) i++;
                continue;
            }
            resize_catmull_kernel<<<dim3((b.width + 31) / 32, (b.height + 7) / 8), blk, 0, st>>>(a.data, frame_stride(a, es), a.channel, type,
                static_cast<const Contrib*>(s->htab.p), static_cast<const Contrib*>(s->vtab.p), b.data, b.width, b.height, frame_stride(b, es));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
        }
        return ACB200_OK;
    }

    int check_frame_args(acb200_session* s, const acb200_model* m, const acb200_plane* src, const acb200_plane* dst, int planes, int type, int shift,
                         double factor, FactorPlan& plan)
    {
        if (!s) return ACB200_EINVAL;
        if (!m || !src || !dst) return fail(s, ACB200_EINVAL, "null argument");
        if (planes < 1 || planes > 3 || !valid_type(type)) return fail(s, ACB200_EINVAL, "frame: 1 to 3 planes of a supported element type");
        if (shift < 0 || shift >= 8 * (type & 0xff)) return fail(s, ACB200_EINVAL, "frame: shift outside the element width");
        if (!src[0].data || src[0].width <= 0 || src[0].height <= 0) return fail(s, ACB200_EINVAL, "frame: empty plane");
        if (!plan_factor(factor, src[0].width, src[0].height, plan)) return fail(s, ACB200_EINVAL, "factor must be at least 1 (and at most 64)");
        const int power = plan.power;
        for (int i = 0; i < planes; i++)
        {
            const acb200_plane& a = src[i];
            const acb200_plane& b = dst[i];
            if (!a.data || !b.data || a.width <= 0 || a.height <= 0) return fail(s, ACB200_EINVAL, "frame: empty plane");
            if (i == 0)
            {
                if (a.channel != 1 || b.channel != 1) return fail(s, ACB200_EINVAL, "frame: plane 0 must be the 1-channel luma plane");
                if (b.width != plan.dw || b.height != plan.dh) return fail(s, ACB200_EINVAL, "frame: destination luma plane must be factor x the source");
                if ((static_cast<long long>(a.width) << power) > 0x7fffffffLL / 16 || (static_cast<long long>(a.height) << power) > 0x7fffffffLL / 16) return fail(s, ACB200_EINVAL, "image too large");
            }
            else
            {
                if (a.channel < 1 || a.channel > 2 || b.channel != a.channel) return fail(s, ACB200_EINVAL, "frame: chroma planes have 1 or 2 channels");
                if (b.width < a.width || b.height < a.height) return fail(s, ACB200_EINVAL, "frame: chroma resize is upscale only");
            }
        }
        return ACB200_OK;
    }

    int check_args(acb200_session* s, const acb200_model* m, const void* src, int w, int h, int c, int type, double factor, void* dst, FactorPlan& plan,
                   bool power_of_two_only = false)
    {
        if (!s) return ACB200_EINVAL;
        if (!m || !src || !dst) return fail(s, ACB200_EINVAL, "null argument");
        if (w <= 0 || h <= 0 || !(c == 1 || c == 3 || c == 4) || !valid_type(type)) return fail(s, ACB200_EINVAL, "unsupported image shape or element type");
        if (power_of_two_only && !passes_for(factor)) return fail(s, ACB200_EINVAL, "factor must be a power of two >= 2");
        if (!plan_factor(factor, w, h, plan)) return fail(s, ACB200_EINVAL, "factor must be at least 1 (and at most 64)");
        if ((static_cast<long long>(w) << plan.power) > 0x7fffffffLL / 16 || (static_cast<long long>(h) << plan.power) > 0x7fffffffLL / 16) return fail(s, ACB200_EINVAL, "image too large");
        return ACB200_OK;
    }
}

extern "C"
{
    int acb200_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
        return n;
    }
    int acb200_device_info(int device, char* name, int name_len, size_t* vram_bytes, int* cc, int* sm_count, int* clock_khz)
    {
        if (device < 0 || device >= acb200_device_count()) return ACB200_ENODEVICE;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return ACB200_ECUDA; }
        if (name && name_len > 0) { std::strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
        if (vram_bytes) *vram_bytes = p.totalGlobalMem;
        if (cc) *cc = p.major * 10 + p.minor;
        if (sm_count) *sm_count = p.multiProcessorCount;
        if (clock_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); *clock_khz = khz; }
        return ACB200_OK;
    }

    int acb200_model_create(int family, int blocks, const float* kernels, int n_kernels, const float* biases, int n_biases,
                            const float* alphas, int n_alphas, acb200_model** out)
    {
        if (family < ACB200_FAMILY_ACNET_LEGACY || family > ACB200_FAMILY_ARNET) return ACB200_EINVAL;
        return acb200_model_create_wide(family, 8, blocks, kernels, n_kernels, biases, n_biases, alphas, n_alphas, out);
    }
    int acb200_model_create_wide(int family, int features, int blocks, const float* kernels, int n_kernels, const float* biases, int n_biases,
                                 const float* alphas, int n_alphas, acb200_model** out)
    {
        if (!out || !kernels || !biases) return ACB200_EINVAL;
        int nk, nb, na;
        if (family <= ACB200_FAMILY_ARNET && features != 8) return ACB200_EINVAL;
        if (!expected_lengths(family, blocks, nk, nb, na, features)) return ACB200_EINVAL;
        if (n_kernels != nk || n_biases != nb || n_alphas != na || (na > 0 && !alphas)) return ACB200_EINVAL;
        acb200_model* m = new (std::nothrow) acb200_model;
        if (!m) return ACB200_ENOMEM;
        m->family = family; m->blocks = blocks; m->features = features;
        m->k.assign(kernels, kernels + nk);
        m->b.assign(biases, biases + nb);
        if (na) m->a.assign(alphas, alphas + na);
        build_chain(*m);
        pack_model(*m);
        m->uid = g_model_uid.fetch_add(1);
        *out = m;
        return ACB200_OK;
    }
    void acb200_model_destroy(acb200_model* model) { delete model; }

    int acb200_session_create(int device, acb200_session** out)
    {
        if (!out) return ACB200_EINVAL;
        *out = nullptr;
        if (device < 0 || device >= acb200_device_count()) return ACB200_ENODEVICE;
        acb200_session* s = new (std::nothrow) acb200_session;
        if (!s) return ACB200_ENOMEM;
        s->device = device;
        // Engine selection for callers that only see the reference's API (ac::core::Processor / the C binding / pyac never pass a
        // session): ACB200_ENGINE = exact | tensor | auto, ACB200_TENSOR_IMPL = mma | tc5 | tm.  Unknown values are an error, not a default.
        if (const char* e = std::getenv("ACB200_ENGINE"))
        {
            if (!std::strcmp(e, "exact") || !std::strcmp(e, "0")) s->engine = 0;
            else if (!std::strcmp(e, "tensor") || !std::strcmp(e, "1")) s->engine = 1;
            else if (!std::strcmp(e, "auto") || !std::strcmp(e, "2")) s->engine = 2;
            else { delete s; return ACB200_EINVAL; }
        }
        if (const char* e = std::getenv("ACB200_TENSOR_IMPL"))
        {
            if (!std::strcmp(e, "mma") || !std::strcmp(e, "0")) s->tensor_impl = 0;
            else if (!std::strcmp(e, "tc5") || !std::strcmp(e, "1")) s->tensor_impl = 1;
            else if (!std::strcmp(e, "tm") || !std::strcmp(e, "2")) s->tensor_impl = 2;
            else { delete s; return ACB200_EINVAL; }
        }
        if (const char* e = std::getenv("ACB200_FUSE"))
        {
            if (!std::strcmp(e, "0")) s->fuse = 0;
            else if (!std::strcmp(e, "1")) s->fuse = 1;
            else if (!std::strcmp(e, "2")) s->fuse = 2;
            else { delete s; return ACB200_EINVAL; }
        }
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess)
        {
            cudaGetLastError();
            delete s;
            return ACB200_ECUDA;
        }
        // keep freed scratch in the pool instead of returning it to the OS between frames
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
        {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        *out = s;
        return ACB200_OK;
    }
    void acb200_session_destroy(acb200_session* s)
    {
        if (!s) return;
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        acb200_session::Buf* bufs[] = { &s->src, &s->dst, &s->y[0], &s->y[1], &s->uv, &s->map[0], &s->map[1], &s->feat, &s->htab, &s->vtab,
                                        &s->pin[0], &s->pin[1], &s->pin[2], &s->pout[0], &s->pout[1], &s->pout[2],
                                        &s->wide[0], &s->wide[1], &s->wide[2], &s->dhtab, &s->dvtab };
        for (auto* b : bufs) if (b->p) cudaFreeAsync(b->p, s->stream);
        cudaStreamSynchronize(s->stream);
        for (auto& kv : s->dev_frags) cudaFree(kv.second);
        for (auto& kv : s->dev_bops) cudaFree(kv.second);
        for (auto& kv : s->dev_tmops) cudaFree(kv.second);
        cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1);
        cudaStreamDestroy(s->stream);
        cudaGetLastError();
        delete s;
    }
    int acb200_session_device(const acb200_session* s) { return s ? s->device : ACB200_EINVAL; }
    const char* acb200_session_error(const acb200_session* s) { return s ? s->error.c_str() : "invalid session"; }
    void acb200_session_clear_error(acb200_session* s) { if (s) s->error = "NO ERROR"; }
    int acb200_session_set_tensor_impl(acb200_session* s, int impl) { if (!s || impl < 0 || impl > 2) return ACB200_EINVAL; s->tensor_impl = impl; return ACB200_OK; }
    int acb200_session_set_fusion(acb200_session* s, int on) { if (!s || on < 0 || on > 2) return ACB200_EINVAL; s->fuse = on; return ACB200_OK; }
    int acb200_session_set_engine(acb200_session* s, int engine) { if (!s || engine < 0 || engine > 2) return ACB200_EINVAL; s->engine = engine; return ACB200_OK; }

    int acb200_process_device(acb200_session* s, const acb200_model* m, const void* d_src, int w, int h, int c, int src_stride, int type,
                              double factor, void* d_dst, int dst_stride, void* stream)
    {
        FactorPlan plan;
        int rc;
        if ((rc = check_args(s, m, d_src, w, h, c, type, factor, d_dst, plan)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaSetDevice(s->device));
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : s->stream;
        const int es = type & 0xff;
        if (src_stride < w * c * es) src_stride = w * c * es;
        if (dst_stride < plan.dw * c * es) dst_stride = plan.dw * c * es;
        return process_on_device(s, m, st, d_src, w, h, c, src_stride, type, plan, d_dst, dst_stride);
    }

    // rows [out_y0, out_y1) of the result only; `dst` points at output row out_y0
    static int process_host_rows(acb200_session* s, const acb200_model* m, const void* src, int w, int h, int c, int src_stride, int type,
                                 double factor, int out_y0, int out_y1, void* dst, int dst_stride, bool whole)
    {
        FactorPlan plan;
        int rc;
        if ((rc = check_args(s, m, src, w, h, c, type, factor, dst, plan, !whole)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaSetDevice(s->device));
        const int es = type & 0xff, ow = plan.dw, oh = plan.dh;
        if (whole) { out_y0 = 0; out_y1 = oh; }
        if (out_y0 < 0 || out_y1 > oh || out_y0 >= out_y1) return fail(s, ACB200_EINVAL, "row range outside the result");
        const size_t line_in = static_cast<size_t>(w) * c * es, line_out = static_cast<size_t>(ow) * c * es;
        if (src_stride < static_cast<int>(line_in)) src_stride = static_cast<int>(line_in);
        if (dst_stride < static_cast<int>(line_out)) dst_stride = static_cast<int>(line_out);
        const size_t sp = staging_pitch(src_stride, line_in, w, c, es, false), dp = staging_pitch(dst_stride, line_out, ow, c, es, true);
        if ((rc = ensure(s, s->stream, s->src, sp * h)) != ACB200_OK) return rc;
        if ((rc = ensure(s, s->stream, s->dst, dp * oh)) != ACB200_OK) return rc;
        ACB_CUDA(s, copy_rows(s->src.p, sp, src, src_stride, line_in, h, cudaMemcpyHostToDevice, s->stream));
        ACB_CUDA(s, cudaEventRecord(s->ev0, s->stream));
        if ((rc = process_on_device(s, m, s->stream, s->src.p, w, h, c, static_cast<int>(sp), type, plan, s->dst.p, static_cast<int>(dp))) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaEventRecord(s->ev1, s->stream));
        s->timed = true;
        ACB_CUDA(s, copy_rows(dst, dst_stride, static_cast<const uint8_t*>(s->dst.p) + static_cast<size_t>(out_y0) * dp, dp, line_out, out_y1 - out_y0,
                              cudaMemcpyDeviceToHost, s->stream));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }
    int acb200_process_host(acb200_session* s, const acb200_model* m, const void* src, int w, int h, int c, int src_stride, int type,
                            double factor, void* dst, int dst_stride)
    {
        return process_host_rows(s, m, src, w, h, c, src_stride, type, factor, 0, 0, dst, dst_stride, true);
    }

    // ---- planar / semi-planar video frames ---------------------------------------------------------------------------------
    int acb200_process_frame_device(acb200_session* s, const acb200_model* m, const acb200_plane* d_src, const acb200_plane* d_dst, int planes,
                                    int type, int shift, double factor, void* stream)
    {
        FactorPlan plan;
        int rc;
        if ((rc = check_frame_args(s, m, d_src, d_dst, planes, type, shift, factor, plan)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaSetDevice(s->device));
        return process_frame_on_device(s, m, stream ? static_cast<cudaStream_t>(stream) : s->stream, d_src, d_dst, planes, type, shift, plan);
    }
    int acb200_process_frame_host(acb200_session* s, const acb200_model* m, const acb200_plane* src, const acb200_plane* dst, int planes,
                                  int type, int shift, double factor)
    {
        FactorPlan plan;
        int rc;
        if ((rc = check_frame_args(s, m, src, dst, planes, type, shift, factor, plan)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaSetDevice(s->device));
        const int es = type & 0xff;
        acb200_plane din[3], dout[3];
        for (int i = 0; i < planes; i++)
        {
            const size_t iline = static_cast<size_t>(src[i].width) * src[i].channel * es, oline = static_cast<size_t>(dst[i].width) * dst[i].channel * es;
            const size_t ip = staging_pitch(frame_stride(src[i], es), iline, src[i].width, src[i].channel, es, false);
            const size_t op = staging_pitch(frame_stride(dst[i], es), oline, dst[i].width, dst[i].channel, es, true);
            if ((rc = ensure(s, s->stream, s->pin[i], ip * src[i].height)) != ACB200_OK) return rc;
            if ((rc = ensure(s, s->stream, s->pout[i], op * dst[i].height)) != ACB200_OK) return rc;
            din[i] = src[i]; din[i].data = static_cast<unsigned char*>(s->pin[i].p); din[i].stride = static_cast<int>(ip);
            dout[i] = dst[i]; dout[i].data = static_cast<unsigned char*>(s->pout[i].p); dout[i].stride = static_cast<int>(op);
            ACB_CUDA(s, copy_rows(din[i].data, ip, src[i].data, frame_stride(src[i], es), iline, src[i].height, cudaMemcpyHostToDevice, s->stream));
        }
        ACB_CUDA(s, cudaEventRecord(s->ev0, s->stream));
        if ((rc = process_frame_on_device(s, m, s->stream, din, dout, planes, type, shift, plan)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaEventRecord(s->ev1, s->stream));
        s->timed = true;
        for (int i = 0; i < planes; i++)
            ACB_CUDA(s, copy_rows(dst[i].data, frame_stride(dst[i], es), dout[i].data, dout[i].stride, static_cast<size_t>(dst[i].width) * dst[i].channel * es, dst[i].height,
                                  cudaMemcpyDeviceToHost, s->stream));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }

    // ---- row bands (one very large image over several GPUs) --------------------------------------------------------------
    int acb200_model_halo(const acb200_model* m)
    {
        if (!m) return ACB200_EINVAL;
        // 3x3 layers on the path of one 2x pass = rows of input context each output row depends on
        if (m->family >= ACB200_FAMILY_ARTCNN) return m->blocks + 3;      // ArtCNN: head + (blocks + 1) + tail; FSRCNNX: 5x5 head (2) + blocks + tail
        return m->family == ACB200_FAMILY_ACNET_LEGACY ? m->blocks + 1 : m->family == ACB200_FAMILY_ACNET ? m->blocks + 2 : 2 * m->blocks + 2;
    }
    int acb200_band_plan(int h, double factor, int halo, int n_bands, int band, int* src_y0, int* src_y1, int* out_y0, int* out_y1)
    {
        const int power = passes_for(factor);
        if (!power || h <= 0 || n_bands <= 0 || band < 0 || band >= n_bands || halo < 0 || !src_y0 || !src_y1 || !out_y0 || !out_y1) return ACB200_EINVAL;
        // contiguous bands of source rows, as even as possible; every band but possibly the last few is non-empty
        const int base = h / n_bands, extra = h % n_bands;
        const int y0 = band * base + std::min(band, extra), y1 = y0 + base + (band < extra ? 1 : 0);
        *out_y0 = y0 << power; *out_y1 = y1 << power;
        // Context per 2x pass is `halo` rows at that pass's input resolution: halo * (1 + 1/2 + 1/4 ...) < 2 * halo source rows
        // over all passes, plus 2 rows of Catmull-Rom support for the chroma plane and 1 for rounding.  At the true image
        // border the band simply ends there and the network's own replicate padding applies -- that is what makes the bands
        // reproduce the whole-image result bit for bit.
        const int ctx = power == 1 ? halo + 3 : 2 * halo + 3;
        *src_y0 = std::max(0, y0 - ctx); *src_y1 = std::min(h, y1 + ctx);
        return ACB200_OK;
    }
    int acb200_process_host_band(acb200_session* s, const acb200_model* m, const void* src, int w, int h, int c, int src_stride, int type,
                                 double factor, int n_bands, int band, void* dst, int dst_stride)
    {
        if (!s) return ACB200_EINVAL;
        if (!m || !src || !dst) return fail(s, ACB200_EINVAL, "null argument");
        int sy0, sy1, oy0, oy1;
        if (acb200_band_plan(h, factor, acb200_model_halo(m), n_bands, band, &sy0, &sy1, &oy0, &oy1) != ACB200_OK) return fail(s, ACB200_EINVAL, "bad band request");
        if (oy0 == oy1) return ACB200_OK;   // more bands than rows
        const int power = passes_for(factor), es = type & 0xff;
        if (src_stride < w * c * es) src_stride = w * c * es;
        if (dst_stride < (w << power) * c * es) dst_stride = (w << power) * c * es;
        const uint8_t* sub = static_cast<const uint8_t*>(src) + static_cast<size_t>(sy0) * src_stride;
        uint8_t* out = static_cast<uint8_t*>(dst) + static_cast<size_t>(oy0) * dst_stride;
        return process_host_rows(s, m, sub, w, sy1 - sy0, c, src_stride, type, factor, oy0 - (sy0 << power), oy1 - (sy0 << power), out, dst_stride, false);
    }
    int acb200_session_sync(acb200_session* s)
    {
        if (!s) return ACB200_EINVAL;
        ACB_CUDA(s, cudaSetDevice(s->device));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }
    float acb200_session_last_kernel_ms(acb200_session* s)
    {
        float ms = -1.0f;
        if (s && s->timed && cudaEventElapsedTime(&ms, s->ev0, s->ev1) != cudaSuccess) { cudaGetLastError(); ms = -1.0f; }
        return ms;
    }

    // packed != 0: `y` is a packed YUV[A] image (c channels) and `uv` is ignored (1-plane forms,
    // core/src/ImageProcess.cpp:15-37,87-112); otherwise the 2-plane forms (:38-61,113-138)
    static int rgb2yuv_host_impl(acb200_session* s, const void* src, int w, int h, int c, int src_stride, int type, void* y, int y_stride, void* uv, int uv_stride, bool packed)
    {
        if (!s) return ACB200_EINVAL;
        if (!src || !y || (!packed && !uv) || w <= 0 || h <= 0 || !(c == 3 || c == 4) || !valid_type(type)) return fail(s, ACB200_EINVAL, "rgb2yuv: bad argument");
        ACB_CUDA(s, cudaSetDevice(s->device));
        const int es = type & 0xff;
        const size_t line = static_cast<size_t>(w) * c * es;
        const size_t sp = pitch_of(w, c, es), yp = pitch_of(w, packed ? c : 1, es), uvp = pitch_of(w, c - 1, es);
        int rc;
        if ((rc = ensure(s, s->stream, s->src, sp * h)) != ACB200_OK) return rc;
        if ((rc = ensure(s, s->stream, s->y[0], yp * h)) != ACB200_OK) return rc;
        if (!packed && (rc = ensure(s, s->stream, s->uv, uvp * h)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaMemcpy2DAsync(s->src.p, sp, src, std::max<size_t>(src_stride, line), line, h, cudaMemcpyHostToDevice, s->stream));
        if (packed)
            rgb2yuv_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->src.p, static_cast<int>(sp), w, h, c, type,
                s->y[0].p, static_cast<int>(yp), c, static_cast<uint8_t*>(s->y[0].p) + es, static_cast<int>(yp), c);
        else
            rgb2yuv_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->src.p, static_cast<int>(sp), w, h, c, type,
                s->y[0].p, static_cast<int>(yp), 1, s->uv.p, static_cast<int>(uvp), c - 1);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        const size_t yline = static_cast<size_t>(w) * (packed ? c : 1) * es;
        ACB_CUDA(s, cudaMemcpy2DAsync(y, std::max<size_t>(y_stride, yline), s->y[0].p, yp, yline, h, cudaMemcpyDeviceToHost, s->stream));
        if (!packed)
            ACB_CUDA(s, cudaMemcpy2DAsync(uv, std::max<size_t>(uv_stride, static_cast<size_t>(w) * (c - 1) * es), s->uv.p, uvp, static_cast<size_t>(w) * (c - 1) * es, h, cudaMemcpyDeviceToHost, s->stream));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }
    static int yuv2rgb_host_impl(acb200_session* s, const void* y, int y_stride, const void* uv, int uv_stride, int w, int h, int c, int type, void* dst, int dst_stride, bool packed)
    {
        if (!s) return ACB200_EINVAL;
        if (!dst || !y || (!packed && !uv) || w <= 0 || h <= 0 || !(c == 3 || c == 4) || !valid_type(type)) return fail(s, ACB200_EINVAL, "yuv2rgb: bad argument");
        ACB_CUDA(s, cudaSetDevice(s->device));
        const int es = type & 0xff;
        const size_t dp = pitch_of(w, c, es), yp = pitch_of(w, packed ? c : 1, es), uvp = pitch_of(w, c - 1, es);
        int rc;
        if ((rc = ensure(s, s->stream, s->dst, dp * h)) != ACB200_OK) return rc;
        if ((rc = ensure(s, s->stream, s->y[0], yp * h)) != ACB200_OK) return rc;
        if (!packed && (rc = ensure(s, s->stream, s->uv, uvp * h)) != ACB200_OK) return rc;
        const size_t yline = static_cast<size_t>(w) * (packed ? c : 1) * es;
        ACB_CUDA(s, cudaMemcpy2DAsync(s->y[0].p, yp, y, std::max<size_t>(y_stride, yline), yline, h, cudaMemcpyHostToDevice, s->stream));
        if (packed)
            yuv2rgb_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->y[0].p, static_cast<int>(yp), c,
                static_cast<uint8_t*>(s->y[0].p) + es, static_cast<int>(yp), c, w, h, c, type, s->dst.p, static_cast<int>(dp));
        else
        {
            ACB_CUDA(s, cudaMemcpy2DAsync(s->uv.p, uvp, uv, std::max<size_t>(uv_stride, static_cast<size_t>(w) * (c - 1) * es), static_cast<size_t>(w) * (c - 1) * es, h, cudaMemcpyHostToDevice, s->stream));
            yuv2rgb_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->y[0].p, static_cast<int>(yp), 1, s->uv.p, static_cast<int>(uvp), c - 1,
                w, h, c, type, s->dst.p, static_cast<int>(dp));
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        ACB_CUDA(s, cudaMemcpy2DAsync(dst, std::max<size_t>(dst_stride, static_cast<size_t>(w) * c * es), s->dst.p, dp, static_cast<size_t>(w) * c * es, h, cudaMemcpyDeviceToHost, s->stream));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }
    int acb200_rgb2yuv_host(acb200_session* s, const void* src, int w, int h, int c, int src_stride, int type, void* y, int y_stride, void* uv, int uv_stride)
    {
        return rgb2yuv_host_impl(s, src, w, h, c, src_stride, type, y, y_stride, uv, uv_stride, false);
    }
    int acb200_rgb2yuv_packed_host(acb200_session* s, const void* src, int w, int h, int c, int src_stride, int type, void* yuv, int yuv_stride)
    {
        return rgb2yuv_host_impl(s, src, w, h, c, src_stride, type, yuv, yuv_stride, nullptr, 0, true);
    }
    int acb200_yuv2rgb_host(acb200_session* s, const void* y, int y_stride, const void* uv, int uv_stride, int w, int h, int c, int type, void* dst, int dst_stride)
    {
        return yuv2rgb_host_impl(s, y, y_stride, uv, uv_stride, w, h, c, type, dst, dst_stride, false);
    }
    int acb200_yuv2rgb_packed_host(acb200_session* s, const void* yuv, int yuv_stride, int w, int h, int c, int type, void* dst, int dst_stride)
    {
        return yuv2rgb_host_impl(s, yuv, yuv_stride, nullptr, 0, w, h, c, type, dst, dst_stride, true);
    }
    int acb200_resize_catmull_rom_host(acb200_session* s, const void* src, int w, int h, int c, int src_stride, int type, void* dst, int ow, int oh, int dst_stride)
    {
        if (!s) return ACB200_EINVAL;
        if (!src || !dst || w <= 0 || h <= 0 || c < 1 || c > 4 || !valid_type(type) || ow <= 0 || oh <= 0 || 2 * ow < w || 2 * oh < h)
            return fail(s, ACB200_EINVAL, "resize: bad argument (any up-scale, down-scale to no less than 1/2)");
        ACB_CUDA(s, cudaSetDevice(s->device));
        const int es = type & 0xff;
        const size_t sp = pitch_of(w, c, es), dp = pitch_of(ow, c, es);
        int rc;
        if ((rc = ensure(s, s->stream, s->src, sp * h)) != ACB200_OK) return rc;
        if ((rc = ensure(s, s->stream, s->dst, dp * oh)) != ACB200_OK) return rc;
        if (ow < w || oh < h)
        {
            // at least one axis shrinks: general 10-tap contributors per axis (an axis that grows keeps its up-scaling taps)
            auto axis = [](std::vector<ContribW>& out, int in_size, int out_size) {
                if (out_size < in_size) return make_contribs_down(out, in_size, out_size);
                std::vector<Contrib> up;
                if (!make_contribs(up, in_size, out_size)) return false;
                out.resize(up.size());
                for (size_t i = 0; i < up.size(); i++)
                {
                    out[i].n0 = up[i].n0; out[i].cnt = up[i].cnt;
                    for (int k = 0; k < 10; k++) out[i].c[k] = k < 6 ? up[i].c[k] : 0.0f;
                }
                return true;
            };
            std::vector<ContribW> ht, vt;
            if (!axis(ht, w, ow) || !axis(vt, h, oh)) return fail(s, ACB200_EINVAL, "resize: unsupported scale");
            if ((rc = ensure(s, s->stream, s->dhtab, ht.size() * sizeof(ContribW))) != ACB200_OK) return rc;
            if ((rc = ensure(s, s->stream, s->dvtab, vt.size() * sizeof(ContribW))) != ACB200_OK) return rc;
            s->dtab_in_w = s->dtab_in_h = s->dtab_out_w = s->dtab_out_h = 0;       // the cached luma down-scale tables are gone
            ACB_CUDA(s, cudaMemcpyAsync(s->dhtab.p, ht.data(), ht.size() * sizeof(ContribW), cudaMemcpyHostToDevice, s->stream));
            ACB_CUDA(s, cudaMemcpyAsync(s->dvtab.p, vt.data(), vt.size() * sizeof(ContribW), cudaMemcpyHostToDevice, s->stream));
            ACB_CUDA(s, cudaStreamSynchronize(s->stream));
            ACB_CUDA(s, cudaMemcpy2DAsync(s->src.p, sp, src, std::max<size_t>(src_stride, static_cast<size_t>(w) * c * es), static_cast<size_t>(w) * c * es, h, cudaMemcpyHostToDevice, s->stream));
            resize_wide_kernel<<<dim3((ow + 31) / 32, (oh + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->src.p, static_cast<int>(sp), c, type,
                static_cast<const ContribW*>(s->dhtab.p), static_cast<const ContribW*>(s->dvtab.p), s->dst.p, ow, oh, static_cast<int>(dp));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
            ACB_CUDA(s, cudaMemcpy2DAsync(dst, std::max<size_t>(dst_stride, static_cast<size_t>(ow) * c * es), s->dst.p, dp, static_cast<size_t>(ow) * c * es, oh, cudaMemcpyDeviceToHost, s->stream));
            ACB_CUDA(s, cudaStreamSynchronize(s->stream));
            return ACB200_OK;
        }
        if ((rc = ensure_tables(s, s->stream, w, h, ow, oh)) != ACB200_OK) return rc;
        ACB_CUDA(s, cudaMemcpy2DAsync(s->src.p, sp, src, std::max<size_t>(src_stride, static_cast<size_t>(w) * c * es), static_cast<size_t>(w) * c * es, h, cudaMemcpyHostToDevice, s->stream));
        resize_catmull_kernel<<<dim3((ow + 31) / 32, (oh + 7) / 8), dim3(32, 8), 0, s->stream>>>(s->src.p, static_cast<int>(sp), c, type,
            static_cast<const Contrib*>(s->htab.p), static_cast<const Contrib*>(s->vtab.p), s->dst.p, ow, oh, static_cast<int>(dp));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        ACB_CUDA(s, cudaMemcpy2DAsync(dst, std::max<size_t>(dst_stride, static_cast<size_t>(ow) * c * es), s->dst.p, dp, static_cast<size_t>(ow) * c * es, oh, cudaMemcpyDeviceToHost, s->stream));
        ACB_CUDA(s, cudaStreamSynchronize(s->stream));
        return ACB200_OK;
    }

    unsigned long long acb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

    void* acb200_host_alloc(size_t bytes)
    {
        static const bool usable = acb200_device_count() > 0;
        if (!usable || bytes == 0) return nullptr;
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return p;
    }
    void acb200_host_free(void* p)
    {
        if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
    }

    const char* acb200_error_string(int code)
    {
        switch (code)
        {
        case ACB200_OK: return "success";
        case ACB200_EINVAL: return "invalid argument";
        case ACB200_ENODEVICE: return "no CUDA device";
        case ACB200_ENOMEM: return "out of memory";
        case ACB200_ECUDA: return "CUDA error";
        default: return "unknown error";
        }
    }
    const char* acb200_version(void) { return "acb200 0.1 (sm_100a)"; }
}
