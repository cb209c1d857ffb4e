// ac::core::Image storage rules, restated from the reference's core/src/Image.cpp:39-110 and
// core/src/Alloc.cpp (aligned allocation): pitch = caller stride if it covers a line, else the line rounded up
// to 4 bytes; map() never owns; view() shares the owner; from()/to() copy row by row.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "AC/Core/Image.hpp"

namespace
{
    constexpr int kStrideAlign = 4;     // AC_CORE_STRIDE_ALIGN default
    constexpr std::size_t kMallocAlign = 64;

    inline int lineBytes(int w, int c, int elementType) { return w * c * (elementType & 0xff); }
    inline int pickPitch(int stride, int line, bool alignUp)
    {
        if (stride >= line) return stride;
        return alignUp ? (line + kStrideAlign - 1) / kStrideAlign * kStrideAlign : line;
    }
}

struct ac::core::Image::ImageData
{
    void* data;
    explicit ImageData(std::size_t size) noexcept
        : data(std::aligned_alloc(kMallocAlign, (size + kMallocAlign - 1) / kMallocAlign * kMallocAlign)) {}
    ~ImageData() noexcept { std::free(data); }
    ImageData(const ImageData&) = delete;
    ImageData& operator=(const ImageData&) = delete;
};

ac::core::Image::Image() noexcept : w(0), h(0), c(0), elementType(UInt8), pitch(0), pixels(nullptr), dptr(nullptr) {}
ac::core::Image::Image(const int w, const int h, const int c, const ElementType elementType, const int stride) : Image()
{
    create(w, h, c, elementType, stride);
}
ac::core::Image::Image(const int w, const int h, const int c, const ElementType elementType, void* const data, const int stride) : Image()
{
    if (data) map(w, h, c, elementType, data, stride);
    else create(w, h, c, elementType, stride);
}
ac::core::Image::Image(const Image&) noexcept = default;
ac::core::Image::Image(Image&&) noexcept = default;
ac::core::Image::~Image() noexcept = default;
ac::core::Image& ac::core::Image::operator=(const Image&) noexcept = default;
ac::core::Image& ac::core::Image::operator=(Image&&) noexcept = default;

void ac::core::Image::create(const int w, const int h, const int c, const ElementType elementType, const int stride)
{
    const int line = lineBytes(w, c, elementType);
    if (h <= 0 || line <= 0) return;
    const int p = pickPitch(stride, line, true);
    auto block = std::make_shared<ImageData>(static_cast<std::size_t>(h) * p);
    this->w = w; this->h = h; this->c = c;
    this->elementType = elementType;
    this->pitch = p;
    this->pixels = block->data;
    this->dptr = std::move(block);
}
void ac::core::Image::map(const int w, const int h, const int c, const ElementType elementType, void* const data, const int stride) noexcept
{
    const int line = lineBytes(w, c, elementType);
    if (h <= 0 || line <= 0 || !data) return;
    this->w = w; this->h = h; this->c = c;
    this->elementType = elementType;
    this->pitch = pickPitch(stride, line, false);
    this->pixels = data;
    this->dptr.reset();
}
void ac::core::Image::from(const int w, const int h, const int c, const ElementType elementType, const void* const data, const int stride)
{
    const int line = lineBytes(w, c, elementType);
    if (h <= 0 || line <= 0 || !data) return;
    const int srcPitch = pickPitch(stride, line, false);
    create(w, h, c, elementType);
    const auto* in = static_cast<const std::uint8_t*>(data);
    for (int y = 0; y < h; y++) std::memcpy(this->line(y), in + static_cast<std::ptrdiff_t>(y) * srcPitch, line);
}
void ac::core::Image::to(void* const data, const int stride) const noexcept
{
    const int line = width() * pixelSize();
    if (height() <= 0 || line <= 0 || !data) return;
    const int dstPitch = pickPitch(stride, line, false);
    auto* out = static_cast<std::uint8_t*>(data);
    for (int y = 0; y < height(); y++) std::memcpy(out + static_cast<std::ptrdiff_t>(y) * dstPitch, this->line(y), line);
}
ac::core::Image ac::core::Image::view(const int x, const int y, const int w, const int h) const noexcept
{
    Image sub{};
    const int x0 = std::max(x, 0), y0 = std::max(y, 0);
    const int x1 = std::min(width(), x + w), y1 = std::min(height(), y + h);
    if (x1 > x0 && y1 > y0)
    {
        sub.w = x1 - x0; sub.h = y1 - y0; sub.c = channels();
        sub.elementType = type();
        sub.pitch = stride();
        sub.pixels = ptr(x0, y0);
        sub.dptr = dptr;
    }
    return sub;
}
ac::core::Image ac::core::Image::clone() const
{
    Image copy{};
    copy.from(width(), height(), channels(), type(), ptr(), stride());
    return copy;
}
