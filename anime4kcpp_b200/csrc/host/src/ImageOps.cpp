// Free image operations of the hot path (reference: core/src/ImageProcess.cpp:346-610 dispatchers,
// core/src/ImageResize.cpp:274-293).  Each call runs on the GPU through the thin CUDA layer with the calling
// thread's session; outputs are created when empty and trusted otherwise, as in the reference.
#include <cstdio>

#include "AC/Core/Image.hpp"

#include "Internal.hpp"

namespace ac::core::internal
{
    acb200_session* threadSession() noexcept
    {
        struct Holder
        {
            acb200_session* s = nullptr;
            bool tried = false;
            ~Holder() { if (s) acb200_session_destroy(s); }
        };
        static thread_local Holder holder;
        if (!holder.tried)
        {
            holder.tried = true;
            const int n = acb200_device_count();
            int best = -1; long long bestScore = -1;
            for (int i = 0; i < n; i++)
            {
                int sms = 0, khz = 0;
                if (acb200_device_info(i, nullptr, 0, nullptr, nullptr, &sms, &khz) != ACB200_OK) continue;
                if (static_cast<long long>(sms) * khz > bestScore) { bestScore = static_cast<long long>(sms) * khz; best = i; }
            }
            if (best < 0 || acb200_session_create(best, &holder.s) != ACB200_OK)
            {
                holder.s = nullptr;
                std::fprintf(stderr, "ac::core: no usable CUDA device; image operations are unavailable (no CPU fallback)\n");
            }
        }
        return holder.s;
    }
}

namespace ac::core::internal
{
    int& lastOpStatus() noexcept
    {
        static thread_local int status = 0;
        return status;
    }
}

namespace
{
    using ac::core::Image;
    using ac::core::internal::threadSession;
    using ac::core::internal::lastOpStatus;

    void ensure(Image& img, int w, int h, int c, int type)
    {
        if (img.empty()) img.create(w, h, c, type);
    }
    void report(acb200_session* s, int rc, const char* what)
    {
        lastOpStatus() = rc;
        if (rc != ACB200_OK) std::fprintf(stderr, "ac::core::%s failed: %s\n", what, s ? acb200_session_error(s) : acb200_error_string(rc));
    }
    void splitPlanes(const Image& src, Image& y, Image& uv, const char* what)
    {
        if (src.empty()) return;
        ensure(y, src.width(), src.height(), 1, src.type());
        ensure(uv, src.width(), src.height(), src.channels() - 1, src.type());
        acb200_session* s = threadSession();
        if (!s) return;
        report(s, acb200_rgb2yuv_host(s, src.ptr(), src.width(), src.height(), src.channels(), src.stride(), src.type(), y.ptr(), y.stride(), uv.ptr(), uv.stride()), what);
    }
    void splitPacked(const Image& src, Image& dst, const char* what)
    {
        if (src.empty()) return;
        ensure(dst, src.width(), src.height(), src.channels(), src.type());
        acb200_session* s = threadSession();
        if (!s) return;
        report(s, acb200_rgb2yuv_packed_host(s, src.ptr(), src.width(), src.height(), src.channels(), src.stride(), src.type(), dst.ptr(), dst.stride()), what);
    }
    void mergePlanes(const Image& y, const Image& uv, Image& dst, const char* what)
    {
        if (y.empty() || uv.empty()) return;
        const int c = uv.channels() + 1;
        ensure(dst, y.width(), y.height(), c, y.type());
        acb200_session* s = threadSession();
        if (!s) return;
        report(s, acb200_yuv2rgb_host(s, y.ptr(), y.stride(), uv.ptr(), uv.stride(), y.width(), y.height(), c, y.type(), dst.ptr(), dst.stride()), what);
    }
    void mergePacked(const Image& src, Image& dst, const char* what)
    {
        if (src.empty()) return;
        ensure(dst, src.width(), src.height(), src.channels(), src.type());
        acb200_session* s = threadSession();
        if (!s) return;
        report(s, acb200_yuv2rgb_packed_host(s, src.ptr(), src.stride(), src.width(), src.height(), src.channels(), src.type(), dst.ptr(), dst.stride()), what);
    }

    void resizeInto(const Image& src, Image& dst, const double fx, const double fy, const int mode) noexcept
    {
        lastOpStatus() = ACB200_OK;
        if (src.empty()) return;
        if (mode != ac::core::RESIZE_CATMULL_ROM)
        {
            // nothing is allocated and dst is left untouched: an empty dst stays empty, never "allocated but unwritten"
            lastOpStatus() = ACB200_EINVAL;
            std::fprintf(stderr, "ac::core::resize: only RESIZE_CATMULL_ROM (any up-scale, down-scale to 1/2) is on the accelerated path\n");
            return;
        }
        if (fx > 0.0 && fy > 0.0)
        {
            if (fx == 1.0 && fy == 1.0) { dst = src; return; }
            const int w = static_cast<int>(src.width() * fx), h = static_cast<int>(src.height() * fy);
            if (dst.width() != w || dst.height() != h || dst.channels() != src.channels() || dst.type() != src.type())
                dst.create(w, h, src.channels(), src.type());
        }
        else
        {
            if (dst.empty()) return;
            if (dst.width() == src.width() && dst.height() == src.height()) { dst = src; return; }
            if (dst.channels() != src.channels() || dst.type() != src.type())
                dst.create(dst.width(), dst.height(), src.channels(), src.type());
        }
        acb200_session* s = threadSession();
        if (!s) { lastOpStatus() = ACB200_ENODEVICE; return; }
        report(s, acb200_resize_catmull_rom_host(s, src.ptr(), src.width(), src.height(), src.channels(), src.stride(), src.type(),
                                                 dst.ptr(), dst.width(), dst.height(), dst.stride()), "resize");
    }
}

int ac::core::lastImageOpStatus() noexcept { return ac::core::internal::lastOpStatus(); }
void ac::core::rgb2yuv(const Image& rgb, Image& yuv) { splitPacked(rgb, yuv, "rgb2yuv"); }
void ac::core::rgb2yuv(const Image& rgb, Image& y, Image& uv) { splitPlanes(rgb, y, uv, "rgb2yuv"); }
void ac::core::rgba2yuva(const Image& rgba, Image& yuva) { splitPacked(rgba, yuva, "rgba2yuva"); }
void ac::core::rgba2yuva(const Image& rgba, Image& y, Image& uva) { splitPlanes(rgba, y, uva, "rgba2yuva"); }
void ac::core::yuv2rgb(const Image& yuv, Image& rgb) { mergePacked(yuv, rgb, "yuv2rgb"); }
void ac::core::yuv2rgb(const Image& y, const Image& uv, Image& rgb) { mergePlanes(y, uv, rgb, "yuv2rgb"); }
void ac::core::yuva2rgba(const Image& yuva, Image& rgba) { mergePacked(yuva, rgba, "yuva2rgba"); }
void ac::core::yuva2rgba(const Image& y, const Image& uva, Image& rgba) { mergePlanes(y, uva, rgba, "yuva2rgba"); }

void ac::core::resize(const Image& src, Image& dst, const double fx, const double fy, const int mode) noexcept
{
    if (src == dst)
    {
        Image tmp{};
        resizeInto(src, tmp, fx, fy, mode);
        if (!tmp.empty()) dst = tmp;
    }
    else resizeInto(src, dst, fx, fy, mode);
}
ac::core::Image ac::core::resize(const Image& src, const double fx, const double fy, const int mode) noexcept
{
    if (fx <= 0.0 || fy <= 0.0) return src;
    Image dst{};
    resizeInto(src, dst, fx, fy, mode);
    return dst;
}
