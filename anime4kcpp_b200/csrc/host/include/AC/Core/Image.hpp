// ac::core::Image and the image operations of the upscaling hot path -- the B200 drop-in's restatement of the
// public interface in the reference's core/include/AC/Core/Image.hpp (class :359-419, free functions :21-356).
// Signatures, element-type codes, stride rules and ownership semantics are kept so reference callers compile
// and behave unchanged; storage rules follow core/src/Image.cpp:39-110.
//
// Off-path operations of the reference header (shl/shr, astype, crop, extract, insert, pixelShuffle, the other
// sixteen resize filters, image file I/O) are outside this build's scope and are not declared.
#pragma once

#include <cstdint>
#include <memory>

#include "ACCoreExport.hpp"

namespace ac::core
{
    class AC_CORE_EXPORT Image
    {
        struct ImageData;

    public:
        // (kind << 8) | sizeof(element); kind 0 = unsigned, 1 = signed, 2 = float
        using ElementType = int;
        static constexpr ElementType UInt8 = 0 << 8 | 1;
        static constexpr ElementType UInt16 = 0 << 8 | 2;
        static constexpr ElementType Float16 = 2 << 8 | 2;
        static constexpr ElementType Float32 = 2 << 8 | 4;

        AC_CORE_EXPORT Image() noexcept;
        AC_CORE_EXPORT Image(int w, int h, int c, ElementType elementType, int stride = 0);
        AC_CORE_EXPORT Image(int w, int h, int c, ElementType elementType, void* data, int stride = 0);
        AC_CORE_EXPORT Image(const Image&) noexcept;
        AC_CORE_EXPORT Image(Image&&) noexcept;
        AC_CORE_EXPORT ~Image() noexcept;
        AC_CORE_EXPORT Image& operator=(const Image&) noexcept;
        AC_CORE_EXPORT Image& operator=(Image&&) noexcept;

        // allocate (owned, ref-counted); row pitch = max(stride, line) rounded as Image.cpp:43
        AC_CORE_EXPORT void create(int w, int h, int c, ElementType elementType, int stride = 0);
        // wrap caller memory, never owns
        AC_CORE_EXPORT void map(int w, int h, int c, ElementType elementType, void* data, int stride = 0) noexcept;
        // allocate and copy rows from caller memory
        AC_CORE_EXPORT void from(int w, int h, int c, ElementType elementType, const void* data, int stride = 0);
        AC_CORE_EXPORT void to(void* data, int stride = 0) const noexcept;
        // sub-rectangle sharing this image's storage (clipped to the image)
        AC_CORE_EXPORT Image view(int x, int y, int w, int h) const noexcept;
        AC_CORE_EXPORT Image clone() const;

        int width() const noexcept { return w; }
        int height() const noexcept { return h; }
        int channels() const noexcept { return c; }
        int stride() const noexcept { return pitch; }
        int size() const noexcept { return h * pitch; }
        int elementSize() const noexcept { return elementType & 0xff; }
        int pixelSize() const noexcept { return c * elementSize(); }
        ElementType type() const noexcept { return elementType; }
        std::uint8_t* data() const noexcept { return static_cast<std::uint8_t*>(pixels); }
        std::uint8_t* line(const int y) const noexcept { return data() + static_cast<std::ptrdiff_t>(y) * pitch; }
        std::uint8_t* pixel(const int x, const int y) const noexcept { return line(y) + x * pixelSize(); }
        void* ptr() const noexcept { return pixels; }
        void* ptr(const int y) const noexcept { return line(y); }
        void* ptr(const int x, const int y) const noexcept { return pixel(x, y); }
        bool empty() const noexcept { return pixels == nullptr; }
        bool isUint() const noexcept { return (elementType >> 8) == 0; }
        bool isInt() const noexcept { return (elementType >> 8) == 1; }
        bool isFloat() const noexcept { return (elementType >> 8) == 2; }
        bool ownership() const noexcept { return dptr != nullptr; }
        // same buffer?
        bool operator==(const Image& other) const noexcept { return (ownership() && other.ownership()) ? (dptr == other.dptr) : (pixels == other.pixels); }
        bool operator!=(const Image& other) const noexcept { return !operator==(other); }

    private:
        int w, h, c;
        ElementType elementType;
        int pitch;
        void* pixels;
        std::shared_ptr<ImageData> dptr;
    };

    // colour split / merge; outputs are allocated when empty, otherwise trusted to have the right shape
    AC_CORE_EXPORT void rgb2yuv(const Image& rgb, Image& yuv);
    AC_CORE_EXPORT void rgb2yuv(const Image& rgb, Image& y, Image& uv);
    AC_CORE_EXPORT void rgba2yuva(const Image& rgba, Image& yuva);
    AC_CORE_EXPORT void rgba2yuva(const Image& rgba, Image& y, Image& uva);
    AC_CORE_EXPORT void yuv2rgb(const Image& yuv, Image& rgb);
    AC_CORE_EXPORT void yuv2rgb(const Image& y, const Image& uv, Image& rgb);
    AC_CORE_EXPORT void yuva2rgba(const Image& yuva, Image& rgba);
    AC_CORE_EXPORT void yuva2rgba(const Image& y, const Image& uva, Image& rgba);

    // numbering of the reference's enum ResizeModes (Image.hpp:272-291); this build implements the hot path's
    // RESIZE_CATMULL_ROM upscale only, other modes leave dst untouched.
    enum ResizeModes
    {
        RESIZE_POINT, RESIZE_CATMULL_ROM, RESIZE_MITCHELL_NETRAVALI, RESIZE_BICUBIC_0_60, RESIZE_BICUBIC_0_75,
        RESIZE_BICUBIC_0_100, RESIZE_BICUBIC_20_50, RESIZE_SOFTCUBIC50, RESIZE_SOFTCUBIC75, RESIZE_SOFTCUBIC100,
        RESIZE_LANCZOS2, RESIZE_LANCZOS3, RESIZE_LANCZOS4, RESIZE_SPLINE16, RESIZE_SPLINE36, RESIZE_SPLINE64, RESIZE_BILINEAR,
    };
    enum ImreadModes { IMREAD_UNCHANGED = 0, IMREAD_GRAYSCALE = 1, IMREAD_COLOR = 3, IMREAD_RGB = 3, IMREAD_RGBA = 4 };

    // Image file I/O (reference core/src/ImageIO.cpp:20-87, there on stb_image / stb_image_write): own codecs on zlib.  Reads PNG
    // (non-interlaced), BMP, binary PNM and uncompressed TGA into an 8-bit image with `mode` channels; writes .png / .bmp / .tga by
    // extension.  JPEG is not supported (imread gives an empty image, imwrite false).  Never throws.
    AC_CORE_EXPORT Image imdecode(const void* buffer, int size, int mode = IMREAD_UNCHANGED) noexcept;
    AC_CORE_EXPORT Image imread(const char* filename, int mode = IMREAD_UNCHANGED) noexcept;
    AC_CORE_EXPORT bool imwrite(const char* filename, const Image& image) noexcept;

    // fx, fy > 0: scale factors; otherwise the size of a non-empty dst decides (ImageResize.cpp:136-165)
    AC_CORE_EXPORT void resize(const Image& src, Image& dst, double fx, double fy, int mode = RESIZE_CATMULL_ROM) noexcept;
    AC_CORE_EXPORT Image resize(const Image& src, double fx, double fy, int mode = RESIZE_CATMULL_ROM) noexcept;
    // Extension (not in the reference): status of the calling thread's most recent free image function above -- 0 when it ran, non-zero
    // when the operation is not on the accelerated path (e.g. a resize mode other than RESIZE_CATMULL_ROM) or the GPU call failed.  The
    // reference's functions return void; the C and Python bindings use this to report an error instead of returning unwritten memory.
    AC_CORE_EXPORT int lastImageOpStatus() noexcept;
}
