// Launches of the TMEM-resident tcgen05 engine (acb200_tm.cuh).
#include <algorithm>
#include <cstdlib>

#include "acb200_internal.cuh"
#include "acb200_tm.cuh"

namespace acbh
{
    // Frame height: the CTA count is tiles_x * ceil(h / (G - 2R)) and every CTA costs about G + c row times, so the best G is the one
    // with the fewest (waves of SM-count CTAs) x (G + c) -- not necessarily the tallest frame.
    int tm_pick_rows(int tiles_x, int h, int R, int sms)
    {
        static const int force_g = [] { const char* e = std::getenv("ACB200_TM_G"); return e ? std::atoi(e) : 0; }();
        int best = TM_GMAX;
        double best_cost = 1e30;
        for (int G = 2 * R + 4; G <= TM_GMAX; G++)
        {
            const int tiles_y = (h + G - 2 * R - 1) / (G - 2 * R);
            const long long tiles = static_cast<long long>(tiles_x) * tiles_y;
            const double waves = static_cast<double>((tiles + sms - 1) / sms);
            const double cost = waves * (G + 8.0);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = G; }
        }
        if (force_g >= 2 * R + 4 && force_g <= TM_GMAX) best = force_g;
        return best;
    }

    template<class S>
    int launch_segment_tm(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a)
    {
        const void* src = a.src; const int src_pitch = a.src_pitch, dst_pitch = a.dst_pitch, w = a.w, h = a.h, type = a.type;
        void* dst = a.dst; const float* map_in = a.map_in; float* map_out = a.map_out; float* feat = a.feat;
        if constexpr (S::R > TM_MAX_R) return ACB_SEG_UNSUPPORTED;
        else
        {
            static_assert(sizeof(TmParams<S>) <= 32764, "kernel parameter block too large");
            const uint32_t* dops = nullptr;
            int rc = device_table(s, st, s->dev_tmops, m.uid, m.tmops, "upload of the TMEM engine's B operands", &dops);
            if (rc != ACB200_OK) return rc;
            if (!s->sm_count)
            {
                ACB_CUDA(s, cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device));
                if (s->sm_count <= 0) s->sm_count = 148;
            }
            TmParams<S> prm;
            prm.src = src; prm.map_in = reinterpret_cast<const uint4*>(map_in); prm.map_out = reinterpret_cast<uint4*>(map_out);
            prm.feat_in = feat; prm.feat_out = feat; prm.dst = dst;
            prm.src_pitch = src_pitch; prm.dst_pitch = dst_pitch; prm.w = w; prm.h = h; prm.type = type;
            prm.rgb_src = S::HEAD ? a.rgb_src : nullptr; prm.rgb_pitch = a.rgb_pitch;
            prm.y_out = S::HEAD ? a.y_out : nullptr; prm.y_pitch = a.y_pitch;
            prm.uv_out = S::HEAD ? a.uv_out : nullptr; prm.uv_in = S::TAIL ? a.uv_in : nullptr; prm.uv_pitch = a.uv_pitch;
            prm.htab = static_cast<const Contrib*>(a.htab); prm.vtab = static_cast<const Contrib*>(a.vtab);
            prm.rgb_dst = a.rgb_dst; prm.rgb_dst_pitch = a.rgb_dst_pitch; prm.uvc = a.uvc;
            constexpr int SW = 32 - 2 * S::R;
            prm.strips_x = (w + SW - 1) / SW;
            prm.tiles_x = (prm.strips_x + 3) / 4;
            prm.G = tm_pick_rows(prm.tiles_x, h, S::R, s->sm_count);
            const int tiles_y = (h + prm.G - 2 * S::R - 1) / (prm.G - 2 * S::R);
            prm.bops = dops + spec.tm_off;
            static const int issuers_env = [] { const char* e = std::getenv("ACB200_TM_ISSUERS"); const int v = e ? std::atoi(e) : TM_ISSUERS; return v < 1 ? 1 : (v > TM_ISSUERS ? TM_ISSUERS : v); }();
            prm.issuers = issuers_env;
            std::memset(prm.k, 0, sizeof(prm.k));
            constexpr int K0 = S::HEAD ? 72 : 0;
            if (S::HEAD)        // the 1 -> 8 head conv, transposed to [tap][cout] so that two couts' weights of a tap are one 64-bit operand
                for (int co = 0; co < 8; co++)
                    for (int p = 0; p < 9; p++) prm.k[p * 8 + co] = m.k[spec.koff + co * 9 + p];
            if (S::TAIL && S::FAM == ACB200_FAMILY_ACNET_LEGACY)
                std::memcpy(prm.k + K0, m.k.data() + spec.koff + K0 + 576 * (S::NCONV + 1), sizeof(float) * 32);
            if (S::TAIL && S::FAM == ACB200_FAMILY_ARNET)       // the 1x1 between the tail's residual conv and the pixel-shuffle conv
                std::memcpy(prm.k + K0, m.k.data() + spec.koff + K0 + 576 * (S::NCONV + 2), sizeof(float) * 64);
            std::memcpy(prm.b, m.b.data() + spec.boff, sizeof(float) * S::NB);
            if (S::NA > 0) std::memcpy(prm.a, m.a.data() + spec.aoff, sizeof(float) * S::NA);
            else prm.a[0] = 0.0f;
            static std::atomic<unsigned long long> optin{0};
            constexpr int SMEM_MAX = S::FAM == ACB200_FAMILY_ARNET ? TM_SMEM_BYTES_ARNET_RGBA : TM_SMEM_BYTES_FUSED;
            const int smem = S::FAM == ACB200_FAMILY_ARNET ? ((prm.uv_in && a.uvc == 3) ? TM_SMEM_BYTES_ARNET_RGBA : TM_SMEM_BYTES_ARNET)
                                                           : (prm.uv_in && a.uvc == 3) ? TM_SMEM_BYTES_FUSED
                                                           : (prm.uv_in || ACB_TM_PROGRESS_MBAR) ? TM_OFF_X : TM_SMEM_BYTES;
            if constexpr (S::NEEDS_LUMA || S::TAIL)
                if (a.uvc == 3)
                {
                    // four-channel images: the instantiation with the RGBA colour path
                    static std::atomic<unsigned long long> optin4{0};
                    if ((rc = smem_optin_once(s, reinterpret_cast<const void*>(segment_tm_kernel<S, true>), SMEM_MAX, optin4)) != ACB200_OK) return rc;
                    segment_tm_kernel<S, true><<<prm.tiles_x * tiles_y, TM_THREADS, smem, st>>>(prm);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                    ACB_CUDA(s, cudaGetLastError());
                    return ACB200_OK;
                }
            if ((rc = smem_optin_once(s, reinterpret_cast<const void*>(segment_tm_kernel<S>), SMEM_MAX, optin)) != ACB200_OK) return rc;
            segment_tm_kernel<S><<<prm.tiles_x * tiles_y, TM_THREADS, smem, st>>>(prm);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ACB_CUDA(s, cudaGetLastError());
            return ACB200_OK;
        }
    }


    int launch_seg_tm(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a)
    {
        switch (spec.kind)
        {
#define ACB_CASE(KIND, TYPE) case KIND: return launch_segment_tm<TYPE>(s, st, m, spec, a);
        ACB_FOR_EACH_SEG(ACB_CASE)
#undef ACB_CASE
        }
        return ACB200_EINVAL;
    }
    bool seg_tm_supported(const acb200_model& m) { return m.family == ACB200_FAMILY_ACNET_LEGACY || m.family == ACB200_FAMILY_ACNET || m.family == ACB200_FAMILY_ARNET; }
    bool seg_tm_chain_supported(const acb200_model& m)
    {
        if (!seg_tm_supported(m)) return false;
        for (const SegSpec& sp : m.chain)
            switch (sp.kind)
            {
#define ACB_CASE(KIND, TYPE) case KIND: if (TYPE::R > TM_MAX_R) return false; break;
            ACB_FOR_EACH_SEG(ACB_CASE)
#undef ACB_CASE
            }
        return true;
    }
}
