// ac::core::Image storage rules, restated from the reference's core/src/Image.cpp:39-110 and
// core/src/Alloc.cpp (aligned allocation): pitch = caller stride if it covers a line, else the line rounded up
// to 4 bytes; map() never owns; view() shares the owner; from()/to() copy row by row.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "AC/Core/Image.hpp"

#include "Internal.hpp"

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

namespace
{
    // Image storage comes from a pool of page-locked memory when a CUDA device is present: the host-fed path is PCIe-bound and a
    // pageable buffer is copied at about a quarter of the pinned rate (and synchronously, through the driver's staging buffer).
    // Page-locking is slow and serialises with the GPU work in flight (measured: tens of milliseconds per call under load), so the
    // pool pins SLABS of 128 MB or more and carves blocks out of them; a released block goes to its capacity bucket's free list and
    // is handed out again -- tools/benchmark's loop allocates a result image per call (the reference's Processor::process creates
    // dst, core/src/processor/Processor.cpp:199-276).  Pinned memory is never returned while the process lives (bounded by kMaxPinned;
    // beyond it, for small images, and on boxes without a GPU: malloc).
    class HostPool
    {
    public:
        static HostPool& get()
        {
            static HostPool* pool = new HostPool;       // never destroyed: static teardown must not call into a CUDA runtime that is gone
            return *pool;
        }
        // capacity buckets: powers of two and the three eighth-steps between them (at most 25 % slack)
        static std::size_t bucket(std::size_t n)
        {
            std::size_t p = kMinPinned;
            while (p < n) p <<= 1;
            const std::size_t q = p >> 3;
            for (std::size_t c = (p >> 1) + q; c < p; c += q) if (c >= n) return c;
            return p;
        }
        void* alloc(const std::size_t size, std::size_t& cap, bool& pinned)
        {
            pinned = false;
            cap = (size + kMallocAlign - 1) / kMallocAlign * kMallocAlign;
            if (enabled && size >= kMinPinned)
            {
                const std::size_t b = bucket(size);
                std::lock_guard<std::mutex> lock(m);
                auto it = cache.find(b);
                if (it != cache.end())
                {
                    void* p = it->second;
                    cache.erase(it);
                    cap = b; pinned = true;
                    return p;
                }
                if (slabLeft < b && total + std::max(kSlab, b) <= kMaxPinned)
                {
                    // what is left of the old slab stays unused (at most one block's worth per slab)
                    const std::size_t bytes = std::max(kSlab, b);
                    if (void* p = acb200_host_alloc(bytes)) { slab = static_cast<unsigned char*>(p); slabLeft = bytes; total += bytes; }
                }
                if (slabLeft >= b)
                {
                    void* p = slab;
                    slab += b; slabLeft -= b;
                    cap = b; pinned = true;
                    return p;
                }
            }
            return std::aligned_alloc(kMallocAlign, cap);
        }
        void release(void* p, const std::size_t cap, const bool pinned)
        {
            if (!p) return;
            if (!pinned) { std::free(p); return; }
            std::lock_guard<std::mutex> lock(m);
            cache.emplace(cap, p);
        }
    private:
        HostPool()
        {
            const char* e = std::getenv("ACB200_PINNED_IMAGES");
            enabled = !(e && e[0] == '0') && acb200_device_count() > 0;
        }
        static constexpr std::size_t kMinPinned = 64 * 1024, kSlab = std::size_t{ 128 } << 20, kMaxPinned = std::size_t{ 8 } << 30;
        std::mutex m;
        std::multimap<std::size_t, void*> cache;
        unsigned char* slab = nullptr;
        std::size_t slabLeft = 0, total = 0;
        bool enabled = false;
    };
}

struct ac::core::Image::ImageData
{
    void* data = nullptr;
    std::size_t cap = 0;
    bool pinned = false;
    explicit ImageData(std::size_t size) noexcept { data = HostPool::get().alloc(size, cap, pinned); }
    ~ImageData() noexcept { HostPool::get().release(data, cap, pinned); }
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
