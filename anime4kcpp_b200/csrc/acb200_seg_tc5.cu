// Launches of the SS-mode tcgen05 engine (acb200_tc5.cuh).
#include "acb200_internal.cuh"
#include "acb200_tc5.cuh"

namespace acbh
{
    template<class S>
    int launch_segment_tc5(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec,
                           const void* src, int src_pitch, void* dst, int dst_pitch, int w, int h, int type,
                           const float* map_in, float* map_out, float* feat)
    {
        static_assert(sizeof(Tc5Params<S>) <= 32764, "kernel parameter block too large");
        const uint32_t* dbops = nullptr;
        int rc = device_table(s, st, s->dev_bops, m.uid, m.bops, "upload of tcgen05 B operands", &dbops);
        if (rc != ACB200_OK) return rc;
        Tc5Params<S> prm;
        prm.src = src; prm.map_in = map_in; prm.map_out = map_out; prm.feat_in = feat; prm.feat_out = feat; prm.dst = dst;
        prm.src_pitch = src_pitch; prm.dst_pitch = dst_pitch; prm.w = w; prm.h = h; prm.type = type;
        prm.tiles_x = (w + S::T - 1) / S::T;
        const int tiles_y = (h + S::T - 1) / S::T;
        prm.bops = dbops + spec.bop_off;
        std::memset(prm.k, 0, sizeof(prm.k));
        constexpr int K0 = S::HEAD ? 72 : 0;
        if (S::HEAD) std::memcpy(prm.k, m.k.data() + spec.koff, sizeof(float) * 72);
        if (S::TAIL && S::FAM == ACB200_FAMILY_ACNET_LEGACY)
            std::memcpy(prm.k + K0 + 64, m.k.data() + spec.koff + K0 + 576 * (S::NCONV + 1), sizeof(float) * 32);
        if (S::TAIL && S::FAM == ACB200_FAMILY_ARNET)
            std::memcpy(prm.k + K0, m.k.data() + spec.koff + K0 + 576 * (S::NCONV + 2), sizeof(float) * 64);
        std::memcpy(prm.b, m.b.data() + spec.boff, sizeof(float) * S::NB);
        if (S::NA > 0) std::memcpy(prm.a, m.a.data() + spec.aoff, sizeof(float) * S::NA);
        else prm.a[0] = 0.0f;
        static std::atomic<unsigned long long> optin{0};
        if (int rc2 = smem_optin_once(s, reinterpret_cast<const void*>(segment_tc5_kernel<S>), TC_SMEM_BYTES, optin)) return rc2;
        segment_tc5_kernel<S><<<prm.tiles_x * tiles_y, TC_THREADS, TC_SMEM_BYTES, st>>>(prm);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        return ACB200_OK;
    }


    int launch_seg_tc5(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a)
    {
        switch (spec.kind)
        {
#define ACB_CASE(KIND, TYPE) case KIND: return launch_segment_tc5<TYPE>(s, st, m, spec, a.src, a.src_pitch, a.dst, a.dst_pitch, a.w, a.h, a.type, a.map_in, a.map_out, a.feat);
        ACB_FOR_EACH_SEG(ACB_CASE)
#undef ACB_CASE
        }
        return ACB200_EINVAL;
    }
}
