// TEST INFRASTRUCTURE.  C entry points over the *compiled reference* (oracle/_ref).
//
// oracle/build_ref.sh compiles the reference's own CPU hot-path sources where they lie under
// /root/reference (core/src/{Alloc,Image,ImageProcess,Model}.cpp, processor/Processor.cpp,
// processor/cpu/CPUProcessor.cpp and the per-ISA Backend.hpp translation units) together with
// this shim into oracle/_ref/libac_ref.so.  The shim adds only what the reference's build
// system would have generated or fetched:
//   * ac::core::simd::support*()  (core/src/SIMD.cpp needs the un-vendored `ruapu`)
//   * ac::core::resize()          (core/src/ImageResize.cpp needs the un-vendored
//                                  stb_image_resize2.h): the identity shortcuts of
//                                  ImageResize.cpp:140-165 verbatim in behaviour, and Catmull-Rom
//                                  upscaling delegated to the oracle restatement
//                                  (orc_resize_catmull_rom) -- so chroma-resize parity stays
//                                  "unpinned"; everything else below is the reference's code.
// and a flat C API so tests / bench.py can call the reference through ctypes.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "AC/Core.hpp"
#include "AC/Core/SIMD.hpp"

extern "C" int orc_resize_catmull_rom(const void* src, int w, int h, int c, int src_stride, int type,
                                      void* dst, int ow, int oh, int dst_stride);

namespace ac::core::simd
{
    bool supportSSE() noexcept { return __builtin_cpu_supports("sse"); }
    bool supportSSE2() noexcept { return __builtin_cpu_supports("sse2"); }
    bool supportAVX() noexcept { return __builtin_cpu_supports("avx"); }
    bool supportAVX2() noexcept { return __builtin_cpu_supports("avx2"); }
    bool supportAVX512() noexcept { return __builtin_cpu_supports("avx512f"); }
    bool supportFMA() noexcept { return __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx"); }
    bool supportNEON() noexcept { return false; }
    bool supportRVV() noexcept { return false; }
    bool supportLSX() noexcept { return false; }
    bool supportLASX() noexcept { return false; }
    bool supportMSA() noexcept { return false; }
    bool supportAltiVec() noexcept { return false; }
    bool supportVSX() noexcept { return false; }
}

namespace
{
    void resizeImpl(const ac::core::Image& src, ac::core::Image& dst, const double fx, const double fy, const int mode) noexcept
    {
        using ac::core::Image;
        if (src.empty()) return;
        if (fx > 0.0 && fy > 0.0)
        {
            if (fx == 1.0 && fy == 1.0) { dst = src; return; }
            auto dstW = static_cast<int>(src.width() * fx);
            auto dstH = static_cast<int>(src.height() * fy);
            if ((dst.width() != dstW) || (dst.height() != dstH) || (dst.channels() != src.channels()) || (dst.type() != src.type()))
                dst.create(dstW, dstH, src.channels(), src.type());
        }
        else
        {
            if (dst.empty()) return;
            if (dst.width() == src.width() && dst.height() == src.height()) { dst = src; return; }
            if ((dst.channels() != src.channels()) || (dst.type() != src.type()))
                dst.create(dst.width(), dst.height(), src.channels(), src.type());
        }
        if (mode != ac::core::RESIZE_CATMULL_ROM || src.type() == Image::Float16 ||
            orc_resize_catmull_rom(src.ptr(), src.width(), src.height(), src.channels(), src.stride(), src.type(),
                                   dst.ptr(), dst.width(), dst.height(), dst.stride()) != 0)
            std::fprintf(stderr, "ref_shim: resize mode/shape not available without stb_image_resize2\n");
    }
}

void ac::core::resize(const Image& src, Image& dst, const double fx, const double fy, const int mode) noexcept
{
    if (src == dst)
    {
        Image tmp{};
        resizeImpl(src, tmp, fx, fy, mode);
        if (!tmp.empty()) dst = tmp;
    }
    else resizeImpl(src, dst, fx, fy, mode);
}
ac::core::Image ac::core::resize(const Image& src, const double fx, const double fy, const int mode) noexcept
{
    if (fx <= 0.0 || fy <= 0.0) return src;
    Image dst{};
    resizeImpl(src, dst, fx, fy, mode);
    return dst;
}

namespace
{
    std::shared_ptr<ac::core::Processor> cached(const char* model, int arch)
    {
        static thread_local std::string key;
        static thread_local std::shared_ptr<ac::core::Processor> proc;
        std::string k = std::string(model ? model : "") + "#" + std::to_string(arch);
        if (!proc || k != key)
        {
            proc = ac::core::Processor::create("cpu", arch, model);
            key = k;
        }
        return proc;
    }
}

extern "C"
{
    // Processor::process(src, dst, factor) on the reference CPU processor; arch 0 = auto ISA, 1 = Generic
    // (core/src/processor/cpu/CPUProcessor.cpp:17-61).  dst must be preallocated by the caller.
    int ref_process(const char* model, int arch, const void* src, int w, int h, int c, int stride, int type,
                    double factor, void* dst, int dst_stride)
    {
        auto proc = cached(model, arch);
        if (!proc || !proc->ok()) return -1;
        ac::core::Image s{ w, h, c, type, const_cast<void*>(src), stride };
        ac::core::Image d{ static_cast<int>(w * factor), static_cast<int>(h * factor), c, type, dst, dst_stride };
        proc->process(s, d, factor);
        return proc->ok() ? 0 : -1;
    }
    const char* ref_processor_name(const char* model, int arch)
    {
        auto proc = cached(model, arch);
        return proc ? proc->name() : "";
    }
    void ref_rgb2yuv(const void* src, int w, int h, int c, int stride, int type, void* y, int y_stride, void* uv, int uv_stride)
    {
        ac::core::Image s{ w, h, c, type, const_cast<void*>(src), stride };
        ac::core::Image yi{ w, h, 1, type, y, y_stride };
        ac::core::Image uvi{ w, h, c - 1, type, uv, uv_stride };
        if (c == 4) ac::core::rgba2yuva(s, yi, uvi); else ac::core::rgb2yuv(s, yi, uvi);
    }
    void ref_yuv2rgb(const void* y, int y_stride, const void* uv, int uv_stride, int w, int h, int c, int type, void* dst, int dst_stride)
    {
        ac::core::Image yi{ w, h, 1, type, const_cast<void*>(y), y_stride };
        ac::core::Image uvi{ w, h, c - 1, type, const_cast<void*>(uv), uv_stride };
        ac::core::Image d{ w, h, c, type, dst, dst_stride };
        if (c == 4) ac::core::yuva2rgba(yi, uvi, d); else ac::core::yuv2rgb(yi, uvi, d);
    }
    // Flat weight arrays exactly as the reference's model objects expose them (core/src/Model.cpp).
    int ref_model_arrays(const char* model, const float** k, int* nk, const float** b, int* nb, const float** a, int* na)
    {
        std::string m = model ? model : "";
        auto pick = [&](auto&& mdl) {
            *k = mdl.kernel(0); *nk = mdl.kernelLength();
            *b = mdl.bias(0); *nb = mdl.biasLength();
            *a = mdl.alphaLength() ? mdl.alpha(0) : nullptr; *na = mdl.alphaLength();
        };
        using namespace ac::core::model;
        struct { const char* name; int fam; int var; } table[] = {
            { "acnet-legacy-gan", 0, (int)ACNetLegacy::Variant::GAN }, { "acnet-legacy-hdn0", 0, (int)ACNetLegacy::Variant::HDN0 },
            { "acnet-legacy-hdn1", 0, (int)ACNetLegacy::Variant::HDN1 }, { "acnet-legacy-hdn2", 0, (int)ACNetLegacy::Variant::HDN2 },
            { "acnet-legacy-hdn3", 0, (int)ACNetLegacy::Variant::HDN3 },
            { "acnet-f8b4", 1, (int)ACNet<8>::Variant::B4_NORMAL }, { "acnet-f8b4-hdn", 1, (int)ACNet<8>::Variant::B4_HDN },
            { "acnet-f8b4-box", 1, (int)ACNet<8>::Variant::B4_BOX }, { "acnet-f8b4-box-hdn", 1, (int)ACNet<8>::Variant::B4_BOX_HDN },
            { "acnet-f8b8", 1, (int)ACNet<8>::Variant::B8_NORMAL }, { "acnet-f8b8-hdn", 1, (int)ACNet<8>::Variant::B8_HDN },
            { "acnet-f8b8-box", 1, (int)ACNet<8>::Variant::B8_BOX }, { "acnet-f8b8-box-hdn", 1, (int)ACNet<8>::Variant::B8_BOX_HDN },
            { "acnet-f8b18", 1, (int)ACNet<8>::Variant::B18_NORMAL }, { "acnet-f8b18-hdn", 1, (int)ACNet<8>::Variant::B18_HDN },
            { "acnet-f8b18-box", 1, (int)ACNet<8>::Variant::B18_BOX }, { "acnet-f8b18-box-hdn", 1, (int)ACNet<8>::Variant::B18_BOX_HDN },
            { "arnet-f8b8", 2, (int)ARNet<8>::Variant::B8_NORMAL }, { "arnet-f8b8-hdn", 2, (int)ARNet<8>::Variant::B8_HDN },
            { "arnet-f8b8-box", 2, (int)ARNet<8>::Variant::B8_BOX }, { "arnet-f8b8-box-hdn", 2, (int)ARNet<8>::Variant::B8_BOX_HDN },
            { "arnet-f8b16", 2, (int)ARNet<8>::Variant::B16_NORMAL }, { "arnet-f8b16-hdn", 2, (int)ARNet<8>::Variant::B16_HDN },
            { "arnet-f8b16-box", 2, (int)ARNet<8>::Variant::B16_BOX }, { "arnet-f8b16-box-hdn", 2, (int)ARNet<8>::Variant::B16_BOX_HDN },
            { "arnet-f8b32", 2, (int)ARNet<8>::Variant::B32_NORMAL }, { "arnet-f8b32-hdn", 2, (int)ARNet<8>::Variant::B32_HDN },
            { "arnet-f8b32-box", 2, (int)ARNet<8>::Variant::B32_BOX }, { "arnet-f8b32-box-hdn", 2, (int)ARNet<8>::Variant::B32_BOX_HDN },
            { "arnet-f8b64", 2, (int)ARNet<8>::Variant::B64_NORMAL }, { "arnet-f8b64-hdn", 2, (int)ARNet<8>::Variant::B64_HDN },
            { "arnet-f8b64-box", 2, (int)ARNet<8>::Variant::B64_BOX }, { "arnet-f8b64-box-hdn", 2, (int)ARNet<8>::Variant::B64_BOX_HDN },
        };
        for (auto& t : table)
            if (m == t.name)
            {
                if (t.fam == 0) pick(ACNetLegacy{ (ACNetLegacy::Variant)t.var });
                else if (t.fam == 1) pick(ACNet<8>{ (ACNet<8>::Variant)t.var });
                else pick(ARNet<8>{ (ARNet<8>::Variant)t.var });
                return 0;
            }
        return -1;
    }
    // tools/benchmark/src/Benchmark.cpp:45-66 in library form: warm-up max(1, 5% of batch), then `batch`
    // process(img, 2.0) calls on 1-channel u8 noise, `threads` images in flight.  Returns seconds.
    double ref_benchmark(const char* model, int arch, int w, int h, int channels, int batch, int threads, unsigned seed)
    {
        auto proc = ac::core::Processor::create("cpu", arch, model);
        if (!proc || !proc->ok()) return -1.0;
        std::mt19937 gen{ seed };
        int pool = std::max(1, std::min(batch, std::max(threads, 4)));
        std::vector<ac::core::Image> images;
        for (int idx = 0; idx < pool; idx++)
        {
            ac::core::Image image{ w, h, channels, ac::core::Image::UInt8 };
            auto length = image.size() & -4;
            for (int i = 0; i < length; i += 4) { auto px = gen(); std::memcpy(image.data() + i, &px, 4); }
            images.emplace_back(image);
        }
        for (int i = 0; i < std::max(static_cast<int>(batch * 0.05), 1); i++) proc->process(images[i % pool], 2.0);
        auto t0 = std::chrono::steady_clock::now();
        if (threads > 1)
        {
            std::atomic<int> next{ 0 };
            std::vector<std::thread> workers;
            for (int t = 0; t < threads; t++)
                workers.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < batch;) proc->process(images[i % pool], 2.0); });
            for (auto& wk : workers) wk.join();
        }
        else for (int i = 0; i < batch; i++) proc->process(images[i % pool], 2.0);
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
}
