// libac_c-compatible C binding over the B200 drop-in's ac::core (contract: include/AC/Core/*.h; reference
// behaviour: binding/c/src/Binding.cpp:19-247 -- NULL arguments give -AC_EINVAL / NULL, handles are heap objects
// owned by the library, dst's plain fields are refreshed after every call that may (re)allocate it).
#include <cstdlib>
#include <cstring>
#include <memory>

#include "AC/Core.hpp"

#include "AC/Core.h"

struct ACImageHandle { ac::core::Image image{}; };
struct ACProcessorHandle { std::shared_ptr<ac::core::Processor> processor{}; };

namespace
{
    void publish(const ac::core::Image& img, ACImage* out)
    {
        out->width = img.width(); out->height = img.height(); out->channels = img.channels();
        out->stride = img.stride(); out->element_type = img.type(); out->ptr = img.ptr();
    }
    ac::core::Image& handleOf(ACImage* image)
    {
        if (!image->hptr) image->hptr = new ACImageHandle{};
        return image->hptr->image;
    }
    void assign(ACImage* dst, ac::core::Image value)
    {
        handleOf(dst) = std::move(value);
        publish(dst->hptr->image, dst);
    }
    template<typename T> T* zeroed() { return static_cast<T*>(std::calloc(1, sizeof(T))); }
    bool usable(const ACImage* image) { return image && image->hptr; }
}

ACImage* ac_image_alloc(void) { return zeroed<ACImage>(); }
void ac_image_free(ACImage** const image)
{
    if (!image || !*image) return;
    ac_image_unref(*image);
    std::free(*image);
    *image = nullptr;
}
int ac_image_ref(const ACImage* const src, ACImage* const dst)
{
    if (!usable(src) || !dst) return AC_ERROR(AC_EINVAL);
    assign(dst, src->hptr->image);
    return AC_SUCCESS;
}
void ac_image_unref(ACImage* const image)
{
    if (!usable(image)) return;
    delete image->hptr;
    std::memset(image, 0, sizeof(ACImage));
}
int ac_image_create(ACImage* const image)
{
    if (!image) return AC_ERROR(AC_EINVAL);
    handleOf(image).create(image->width, image->height, image->channels, image->element_type, image->stride);
    publish(image->hptr->image, image);
    return AC_SUCCESS;
}
int ac_image_map(ACImage* const image)
{
    if (!image) return AC_ERROR(AC_EINVAL);
    handleOf(image).map(image->width, image->height, image->channels, image->element_type, image->ptr, image->stride);
    publish(image->hptr->image, image);
    return AC_SUCCESS;
}
int ac_image_from(ACImage* const image, const void* const data)
{
    if (!image) return AC_ERROR(AC_EINVAL);
    handleOf(image).from(image->width, image->height, image->channels, image->element_type, data, image->stride);
    publish(image->hptr->image, image);
    return AC_SUCCESS;
}
int ac_image_view(const ACImage* const src, ACImage* const dst, const int x, const int y, const int w, const int h)
{
    if (!usable(src) || !dst) return AC_ERROR(AC_EINVAL);
    assign(dst, src->hptr->image.view(x, y, w, h));
    return AC_SUCCESS;
}
int ac_image_clone(const ACImage* const src, ACImage* const dst)
{
    if (!usable(src) || !dst) return AC_ERROR(AC_EINVAL);
    assign(dst, src->hptr->image.clone());
    return AC_SUCCESS;
}
int ac_image_to(const ACImage* const image, void* const data, const int stride)
{
    if (!usable(image) || !data) return AC_ERROR(AC_EINVAL);
    image->hptr->image.to(data, stride);
    return AC_SUCCESS;
}

int ac_imread(const char* const filename, const int mode, ACImage* const image)
{
    if (!image || !filename) return AC_ERROR(AC_EINVAL);
    ac::core::Image loaded = ac::core::imread(filename, mode);
    if (loaded.empty()) return AC_ERROR(AC_EIO);
    assign(image, loaded);
    return AC_SUCCESS;
}
int ac_imwrite(const char* const filename, const ACImage* const image)
{
    if (!usable(image) || !filename) return AC_ERROR(AC_EINVAL);
    return ac::core::imwrite(filename, image->hptr->image) ? AC_SUCCESS : AC_ERROR(AC_EIO);
}

int ac_resize(const ACImage* const src, ACImage* const dst, const double fx, const double fy, const int mode)
{
    if (!usable(src) || !usable(dst)) return AC_ERROR(AC_EINVAL);
    ac::core::resize(src->hptr->image, dst->hptr->image, fx, fy, mode);
    // an unsupported mode (only RESIZE_CATMULL_ROM is on this path) or a failed GPU call is an error, not a silently unwritten image
    if (ac::core::lastImageOpStatus() != 0) return AC_ERROR(AC_EINVAL);
    publish(dst->hptr->image, dst);
    return AC_SUCCESS;
}
int ac_rgb2yuv(const ACImage* const rgb, ACImage* const yuv)
{
    if (!usable(rgb) || !usable(yuv)) return AC_ERROR(AC_EINVAL);
    ac::core::rgb2yuv(rgb->hptr->image, yuv->hptr->image);
    publish(yuv->hptr->image, yuv);
    return AC_SUCCESS;
}
int ac_rgba2yuva(const ACImage* const rgba, ACImage* const yuva)
{
    if (!usable(rgba) || !usable(yuva)) return AC_ERROR(AC_EINVAL);
    ac::core::rgba2yuva(rgba->hptr->image, yuva->hptr->image);
    publish(yuva->hptr->image, yuva);
    return AC_SUCCESS;
}
int ac_yuv2rgb(const ACImage* const yuv, ACImage* const rgb)
{
    if (!usable(yuv) || !usable(rgb)) return AC_ERROR(AC_EINVAL);
    ac::core::yuv2rgb(yuv->hptr->image, rgb->hptr->image);
    publish(rgb->hptr->image, rgb);
    return AC_SUCCESS;
}
int ac_yuva2rgba(const ACImage* const yuva, ACImage* const rgba)
{
    if (!usable(yuva) || !usable(rgba)) return AC_ERROR(AC_EINVAL);
    ac::core::yuva2rgba(yuva->hptr->image, rgba->hptr->image);
    publish(rgba->hptr->image, rgba);
    return AC_SUCCESS;
}

ACProcessor* ac_processor_alloc(void) { return zeroed<ACProcessor>(); }
void ac_processor_free(ACProcessor** const processor)
{
    if (!processor || !*processor) return;
    ac_processor_unref(*processor);
    std::free(*processor);
    *processor = nullptr;
}
int ac_processor_ref(const ACProcessor* const src, ACProcessor* const dst)
{
    if (!src || !src->hptr || !dst) return AC_ERROR(AC_EINVAL);
    if (!dst->hptr) dst->hptr = new ACProcessorHandle{};
    dst->hptr->processor = src->hptr->processor;
    dst->device = src->device; dst->type = src->type; dst->model = src->model;
    return AC_SUCCESS;
}
void ac_processor_unref(ACProcessor* const processor)
{
    if (!processor || !processor->hptr) return;
    delete processor->hptr;
    std::memset(processor, 0, sizeof(ACProcessor));
}
int ac_processor_create(ACProcessor* const processor)
{
    if (!processor) return AC_ERROR(AC_EINVAL);
    if (!processor->hptr) processor->hptr = new ACProcessorHandle{};
    processor->hptr->processor = ac::core::Processor::create(processor->type, processor->device, processor->model);
    return ac_processor_ok(processor);
}
int ac_processor_process(ACProcessor* const processor, const ACImage* const src, ACImage* const dst, const double factor)
{
    if (!processor || !processor->hptr || !processor->hptr->processor || !usable(src) || !usable(dst)) return AC_ERROR(AC_EINVAL);
    processor->hptr->processor->process(src->hptr->image, dst->hptr->image, factor);
    publish(dst->hptr->image, dst);
    return ac_processor_ok(processor);
}
extern "C" int ac_b200_process_frame(ac::core::Processor* processor, const struct acb200_plane* src, const struct acb200_plane* dst, int planes,
                                     int elementType, int shift, double factor);
int ac_processor_process_frame(ACProcessor* const processor, const ACPlane* const src, const ACPlane* const dst, const int planes, const int element_type,
                               const int shift, const double factor)
{
    if (!processor || !processor->hptr || !processor->hptr->processor || !src || !dst) return AC_ERROR(AC_EINVAL);
    // ACPlane and acb200_plane are the same POD (both mirror ac::video::Frame::plane)
    if (ac_b200_process_frame(processor->hptr->processor.get(), reinterpret_cast<const acb200_plane*>(src), reinterpret_cast<const acb200_plane*>(dst), planes,
                              element_type, shift, factor) != 0 && processor->hptr->processor->ok())
        return AC_ERROR(AC_EPROCESSOR);     // not a B200 processor: nothing ran
    return ac_processor_ok(processor);
}
int ac_processor_ok(const ACProcessor* const processor)
{
    if (!processor || !processor->hptr || !processor->hptr->processor) return AC_ERROR(AC_EINVAL);
    return processor->hptr->processor->ok() ? AC_SUCCESS : AC_ERROR(AC_EPROCESSOR);
}
const char* ac_processor_error(const ACProcessor* const processor)
{
    return (processor && processor->hptr && processor->hptr->processor) ? processor->hptr->processor->error() : nullptr;
}
const char* ac_processor_name(const ACProcessor* const processor)
{
    return (processor && processor->hptr && processor->hptr->processor) ? processor->hptr->processor->name() : nullptr;
}
int ac_processor_type(const ACProcessor* const processor)
{
    if (!processor || !processor->hptr || !processor->hptr->processor) return AC_ERROR(AC_EINVAL);
    return processor->hptr->processor->type();
}
const char* ac_processor_type_name(const ACProcessor* const processor)
{
    return (processor && processor->hptr && processor->hptr->processor) ? processor->hptr->processor->typeName() : nullptr;
}
const char* ac_processor_info(const int processor_type)
{
    switch (processor_type)
    {
    case AC_PROCESSOR_CPU: return ac::core::Processor::info<ac::core::Processor::CPU>();
    case AC_PROCESSOR_CUDA: return ac::core::Processor::info<ac::core::Processor::CUDA>();
    default: return "unsupported processor";
    }
}
const char* ac_processor_list_info(void) { return ac::core::Processor::listInfo(); }
