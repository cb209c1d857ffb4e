// Launches of the mma.sync tensor engine (acb200_mma.cuh).
#include "acb200_internal.cuh"
#include "acb200_mma.cuh"

namespace acbh
{
    // TMA descriptor of an inter-segment map of the mma engine (layout in acb200_mma.cuh): 32-bit words, dims {4 w, h, 2 planes},
    // box {4 * 56, 56, 2}, no swizzle, out-of-bounds coordinates read as zero.  cuTensorMapEncodeTiled comes from the driver
    // through the runtime (no libcuda link dependency).
    using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    int encode_map_tmap(acb200_session* s, CUtensorMap* tm, const void* map, int w, int h)
    {
        static std::atomic<EncodeTiledFn> cached{ nullptr };
        EncodeTiledFn fn = cached.load(std::memory_order_acquire);
        if (!fn)
        {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
            const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
            if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return fail(s, ACB200_ECUDA, "cuTensorMapEncodeTiled is not available from the driver", e);
            fn = reinterpret_cast<EncodeTiledFn>(p);
            cached.store(fn, std::memory_order_release);
        }
        // a row of a plane is contiguous over (x, channel): described as 32-bit words so that the 56-pixel box row is ONE 896-byte
        // extent (224 words, the box limit is 256 elements) instead of 56 extents of 16 bytes
        const cuuint64_t dims[3] = { static_cast<cuuint64_t>(w) * 4, static_cast<cuuint64_t>(h), 2 };
        const cuuint64_t strides[2] = { static_cast<cuuint64_t>(w) * 16, static_cast<cuuint64_t>(w) * h * 16 };
        const cuuint32_t box[3] = { FT * 4, FT, 2 }, estr[3] = { 1, 1, 1 };
        const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(map), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(s, ACB200_ECUDA, "cuTensorMapEncodeTiled failed");
        return ACB200_OK;
    }

    template<class S>
    int launch_segment_mma(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec,
                           const void* src, int src_pitch, void* dst, int dst_pitch, int w, int h, int type,
                           const float* map_in, float* map_out, float* feat)
    {
        static_assert(sizeof(MmaParams<S>) <= 32764, "kernel parameter block too large");
        const uint32_t* dfrags = nullptr;
        int rc = device_table(s, st, s->dev_frags, m.uid, m.frags, "upload of weight fragments", &dfrags);
        if (rc != ACB200_OK) return rc;
        MmaParams<S> prm;
        prm.src = src; prm.map_in = map_in; prm.map_out = map_out; prm.feat_in = feat; prm.feat_out = feat; prm.dst = dst;
        prm.src_pitch = src_pitch; prm.dst_pitch = dst_pitch; prm.w = w; prm.h = h; prm.type = type;
        prm.tiles_x = (w + S::T - 1) / S::T;
        const int tiles_y = (h + S::T - 1) / S::T;
        prm.frags = dfrags + spec.frag_off;
        std::memset(&prm.tmap, 0, sizeof(prm.tmap));
        if (!S::HEAD && (rc = encode_map_tmap(s, &prm.tmap, map_in, w, h)) != ACB200_OK) return rc;
        std::memset(prm.k, 0, sizeof(prm.k));
        if (S::HEAD) std::memcpy(prm.k, m.k.data() + spec.koff, sizeof(float) * 72);
        if (S::TAIL && S::FAM == ACB200_FAMILY_ACNET_LEGACY)
            std::memcpy(prm.k + (S::HEAD ? 72 : 0), m.k.data() + spec.koff + (S::HEAD ? 72 : 0) + 576 * (S::NCONV + 1), sizeof(float) * 32);
        std::memcpy(prm.b, m.b.data() + spec.boff, sizeof(float) * S::NB);
        if (S::NA > 0) std::memcpy(prm.a, m.a.data() + spec.aoff, sizeof(float) * S::NA);
        else prm.a[0] = 0.0f;
        static std::atomic<unsigned long long> optin{0};
        if (int rc2 = smem_optin_once(s, reinterpret_cast<const void*>(segment_mma_kernel<S>), MMA_SMEM_BYTES, optin)) return rc2;
        segment_mma_kernel<S><<<prm.tiles_x * tiles_y, MMA_THREADS, MMA_SMEM_BYTES, st>>>(prm);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ACB_CUDA(s, cudaGetLastError());
        return ACB200_OK;
    }


    int launch_seg_mma(acb200_session* s, cudaStream_t st, const acb200_model& m, const SegSpec& spec, const SegLaunch& a)
    {
        switch (spec.kind)
        {
#define ACB_CASE(KIND, TYPE) case KIND: return launch_segment_mma<TYPE>(s, st, m, spec, a.src, a.src_pitch, a.dst, a.dst_pitch, a.w, a.h, a.type, a.map_in, a.map_out, a.feat);
        ACB_FOR_EACH_SEG(ACB_CASE)
#undef ACB_CASE
        }
        return ACB200_EINVAL;
    }
}
