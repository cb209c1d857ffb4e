// Launches of the exact fp32 FFMA engine (acb200_ffma.cuh), one kernel instantiation per segment type.
#include "acb200_internal.cuh"

namespace acbh
{
    template<class S>
    int launch_segment(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec,
                       const void* src, int src_pitch, void* dst, int dst_pitch, int w, int h, int type,
                       const float* map_in, float* map_out, float* feat)
    {
        static_assert(sizeof(SegParams<S>) <= 32764, "kernel parameter block too large");
        SegParams<S> prm;
        prm.src = src; prm.map_in = map_in; prm.map_out = map_out; prm.feat_in = feat; prm.feat_out = feat; prm.dst = dst;
        prm.src_pitch = src_pitch; prm.dst_pitch = dst_pitch; prm.w = w; prm.h = h; prm.type = type;
        prm.tiles_x = (w + S::T - 1) / S::T;
        const int tiles_y = (h + S::T - 1) / S::T;
        std::memcpy(prm.k, m.k.data() + spec.koff, sizeof(float) * S::NK);
        std::memcpy(prm.b, m.b.data() + spec.boff, sizeof(float) * S::NB);
        if (S::NA > 0) std::memcpy(prm.a, m.a.data() + spec.aoff, sizeof(float) * S::NA);
        else prm.a[0] = 0.0f;
        static std::atomic<unsigned long long> optin{0};
        if (int rc = smem_optin_once(s, reinterpret_cast<const void*>(segment_ffma_kernel<S>), FFMA_SMEM_BYTES, optin)) return rc;
        segment_ffma_kernel<S><<<prm.tiles_x * tiles_y, FFMA_THREADS, FFMA_SMEM_BYTES, st>>>(prm);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        return ACB200_OK;
    }


    int launch_seg_ffma(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a)
    {
        switch (spec.kind)
        {
#define ACB_CASE(KIND, TYPE) case KIND: return launch_segment<TYPE>(s, st, m, spec, a.src, a.src_pitch, a.dst, a.dst_pitch, a.w, a.h, a.type, a.map_in, a.map_out, a.feat);
        ACB_FOR_EACH_SEG(ACB_CASE)
#undef ACB_CASE
        }
        return ACB200_EINVAL;
    }
}
