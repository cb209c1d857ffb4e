// ArtCNN<16/32>, FSRCNNX<8/16>: one launch per layer over fp32 maps in HBM (acb200_wide.cuh, acb200_wide_tc.cuh).
#include "acb200_internal.cuh"
#include "acb200_wide.cuh"
#include "acb200_wide_tc.cuh"

namespace acbh
{
    // ---- ArtCNN<16/32>, FSRCNNX<8/16>: one launch per layer (two for 32 output channels) over fp32 maps in HBM -------------
    template<int F, int NCO, int MODE>
    int launch_wide_conv(acb200_session* s, cudaStream_t st, const float* in, float* out, const float* res, void* dst, int dst_pitch, int type,
                         int w, int h, int co0, int act, const float* k, const float* b, const float* a)
    {
        static_assert(sizeof(WideConvParams<F, NCO>) <= 32764, "kernel parameter block too large");
        WideConvParams<F, NCO> prm;
        prm.in = in; prm.out = out; prm.res = res; prm.dst = dst; prm.dst_pitch = dst_pitch; prm.type = type;
        prm.w = w; prm.h = h; prm.co0 = co0; prm.act = act;
        std::memcpy(prm.k, k + static_cast<size_t>(co0) * 9 * F, sizeof(prm.k));
        std::memcpy(prm.b, b + co0, sizeof(prm.b));
        if (a) std::memcpy(prm.a, a + co0, sizeof(prm.a)); else std::memset(prm.a, 0, sizeof(prm.a));
        static std::atomic<unsigned long long> optin{0};
        if (int rc = smem_optin_once(s, reinterpret_cast<const void*>(wide_conv_kernel<F, NCO, MODE>), wide_smem_bytes<F>(), optin)) return rc;
        wide_conv_kernel<F, NCO, MODE><<<dim3((w + WIDE_TW - 1) / WIDE_TW, (h + WIDE_TH - 1) / WIDE_TH), WIDE_THREADS, wide_smem_bytes<F>(), st>>>(prm);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        return ACB200_OK;
    }
    // the same layer on the tensor cores (tcgen05, split fp16): one launch, B operand from the model's packed table
    template<int F>
    int wide_conv_layer_tc(acb200_session* s, cudaStream_t st, const acb200_model& m, int conv_index, const float* in, float* out, const float* res,
                           int w, int h, int act, const float* b, const float* a)
    {
        const uint32_t* dbops = nullptr;
        int rc = device_table(s, st, s->dev_bops, m.uid, m.bops, "upload of tcgen05 B operands", &dbops);
        if (rc != ACB200_OK) return rc;
        WideTcParams<F> prm;
        prm.in = in; prm.out = out; prm.res = res; prm.w = w; prm.h = h; prm.act = act;
        prm.bop = dbops + static_cast<size_t>(conv_index) * (WideTc<F>::B_BYTES / 4);
        std::memcpy(prm.b, b, sizeof(prm.b));
        if (a) std::memcpy(prm.a, a, sizeof(prm.a)); else std::memset(prm.a, 0, sizeof(prm.a));
        static std::atomic<unsigned long long> optin{0};
        if ((rc = smem_optin_once(s, reinterpret_cast<const void*>(wide_tc_kernel<F>), WideTc<F>::SMEM_BYTES, optin)) != ACB200_OK) return rc;
        wide_tc_kernel<F><<<dim3((w + WTC_TW - 1) / WTC_TW, (h + WideTc<F>::TH - 1) / WideTc<F>::TH), WideTc<F>::THREADS, WideTc<F>::SMEM_BYTES, st>>>(prm);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        return ACB200_OK;
    }
    // F -> F conv layer: output channels in launches of at most 16 (the weights travel as kernel parameters)
    template<int F>
    int wide_conv_layer(acb200_session* s, cudaStream_t st, const float* in, float* out, const float* res, int w, int h, int act,
                        const float* k, const float* b, const float* a)
    {
        constexpr int NCO = F < 16 ? F : 16;
        for (int co0 = 0; co0 < F; co0 += NCO)
        {
            const int rc = launch_wide_conv<F, NCO, WIDE_STORE>(s, st, in, out, res, nullptr, 0, 0, w, h, co0, act, k, b, a);
            if (rc != ACB200_OK) return rc;
        }
        return ACB200_OK;
    }
    template<int F>
    int luma_pass_wide(acb200_session* s, cudaStream_t st, const acb200_model& m, const void* src, int src_pitch, void* dst, int dst_pitch,
                       int w, int h, int type, bool tensor)
    {
        // F -> F conv number `ci` (0-based) of the model: tensor engine for 16 / 32 features, exact FFMA kernels otherwise
        auto conv = [&](int ci, const float* in_, float* out_, const float* res_, int act, const float* k_, const float* b_, const float* a_) -> int {
            if constexpr (F >= 16)
            {
                if (tensor) return wide_conv_layer_tc<F>(s, st, m, ci, in_, out_, res_, w, h, act, b_, a_);
            }
            return wide_conv_layer<F>(s, st, in_, out_, res_, w, h, act, k_, b_, a_);
        };
        const size_t bytes = static_cast<size_t>(w) * h * F * sizeof(float);
        int rc;
        for (int i = 0; i < 3; i++) if ((rc = ensure(s, st, s->wide[i], bytes)) != ACB200_OK) return rc;
        float* feat = static_cast<float*>(s->wide[0].p);
        float* in = static_cast<float*>(s->wide[2].p);
        float* out = static_cast<float*>(s->wide[1].p);
        const bool art = m.family == ACB200_FAMILY_ARTCNN;
        const int ks = art ? 3 : 5, KH = F * ks * ks, KL = F * F * 9, B = m.blocks;
        const float* k = m.k.data();
        const float* b = m.b.data();
        const float* a = m.a.empty() ? nullptr : m.a.data();
        // head (CPUProcessor.cpp:1533 / :1635)
        {
            WideHeadParams hp;
            hp.src = src; hp.src_pitch = src_pitch; hp.type = type; hp.out = feat; hp.w = w; hp.h = h; hp.F = F;
            std::memcpy(hp.k, k, sizeof(float) * KH);
            std::memcpy(hp.b, b, sizeof(float) * F);
            const dim3 grid((w + 31) / 32, (h + 7) / 8);
            if (art) wide_head_kernel<3><<<grid, 256, 0, st>>>(hp); else wide_head_kernel<5><<<grid, 256, 0, st>>>(hp);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
        }
        int l = 1;
        const float* cur = feat;
        auto layer_k = [&](int layer) { return k + KH + static_cast<size_t>(KL) * (layer - 1); };
        if (art)
        {
            // blocks x (conv + ReLU), then conv + Identity + feat (CPUProcessor.cpp:1535-1545)
            for (int i = 0; i < B; i++, l++)
            {
                if ((rc = conv(l - 1, cur, out, nullptr, ACT_RELU, layer_k(l), b + F * l, nullptr)) != ACB200_OK) return rc;
                cur = out; std::swap(in, out);
            }
            if ((rc = conv(l - 1, cur, out, feat, ACT_IDENTITY, layer_k(l), b + F * l, nullptr)) != ACB200_OK) return rc;
            cur = out; std::swap(in, out); l++;
        }
        else
        {
            // (blocks - 1) x (conv + PReLU), then conv + PReLU -> 1x1 -> + feat -> PReLU (CPUProcessor.cpp:1637-1651)
            for (int i = 0; i < B; i++, l++)
            {
                if ((rc = conv(l - 1, cur, out, nullptr, ACT_PRELU, layer_k(l), b + F * l, a + F * (l - 1))) != ACB200_OK) return rc;
                cur = out; std::swap(in, out);
            }
            if constexpr (F <= 16)
            {
                WidePointParams<F> pp;
                pp.in = cur; pp.feat = feat; pp.out = out; pp.n_pixels = w * h;
                std::memcpy(pp.k, k + KH + static_cast<size_t>(KL) * B, sizeof(pp.k));
                std::memcpy(pp.b, b + F * l, sizeof(pp.b));
                std::memcpy(pp.a, a + F * (l - 1), sizeof(pp.a));
                wide_pointwise_kernel<F><<<(w * h + 255) / 256, 256, 0, st>>>(pp);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                ACB_CUDA(s, cudaGetLastError());
            }
            cur = out; std::swap(in, out); l++;
        }
        // F -> 4 + pixel shuffle (CPUProcessor.cpp:1548 / :1654)
        const float* kt = art ? layer_k(l) : k + KH + static_cast<size_t>(KL) * B + F * F;
        // launch_wide_conv copies NCO * 9 * F weights starting at co0 = 0
        return launch_wide_conv<F, 4, WIDE_SHUFFLE>(s, st, cur, nullptr, nullptr, dst, dst_pitch, type, w, h, 0, ACT_IDENTITY, kt, b + F * l, nullptr);
    }


    int luma_pass_wide_any(acb200_session* s, cudaStream_t st, const acb200_model& m, const void* src, int src_pitch, void* dst, int dst_pitch,
                           int w, int h, int type, bool tensor)
    {
        if (static_cast<long long>(w) * h * m.features > 0x7fffffffLL / 4) return fail(s, ACB200_EINVAL, "image too large for this model family");
        switch (m.features)
        {
        case 8: return luma_pass_wide<8>(s, st, m, src, src_pitch, dst, dst_pitch, w, h, type, tensor);
        case 16: return luma_pass_wide<16>(s, st, m, src, src_pitch, dst, dst_pitch, w, h, type, tensor);
        default: return luma_pass_wide<32>(s, st, m, src, src_pitch, dst, dst_pitch, w, h, type, tensor);
        }
    }
}
