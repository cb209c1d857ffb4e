/*
 * C image API of the B200 drop-in.  Binary-compatible with the reference's libac_c
 * (binding/c/include/AC/Core/Image.h:15-127, implemented in binding/c/src/Binding.cpp:19-159): same POD
 * layout, enum values, function names and ownership rules -- the caller fills the plain fields, the library
 * owns `hptr`.  File I/O (ac_imread / ac_imwrite) runs on this build's own PNG / BMP / PNM / TGA codecs.
 */
#ifndef AC_BINDING_C_CORE_IMAGE_H
#define AC_BINDING_C_CORE_IMAGE_H

#include <stddef.h>
#include <stdint.h>

#ifndef AC_C_EXPORT
#   if defined(__GNUC__)
#       define AC_C_EXPORT __attribute__((visibility("default")))
#   else
#       define AC_C_EXPORT
#   endif
#endif
#ifdef __cplusplus
#   define AC_C_API extern "C" AC_C_EXPORT
#else
#   define AC_C_API AC_C_EXPORT
#endif

/* (kind << 8) | bytes per element */
enum ACImageElementType { AC_IMAGE_UINT8 = 0x001, AC_IMAGE_UINT16 = 0x002, AC_IMAGE_FLOAT32 = 0x204 };
enum ACImreadModes { AC_IMREAD_UNCHANGED = 0, AC_IMREAD_GRAYSCALE = 1, AC_IMREAD_COLOR = 3, AC_IMREAD_RGB = 3, AC_IMREAD_RGBA = 4 };
enum ACResizeModes
{
    AC_RESIZE_POINT, AC_RESIZE_CATMULL_ROM, AC_RESIZE_MITCHELL_NETRAVALI, AC_RESIZE_BICUBIC_0_60, AC_RESIZE_BICUBIC_0_75,
    AC_RESIZE_BICUBIC_0_100, AC_RESIZE_BICUBIC_20_50, AC_RESIZE_SOFTCUBIC50, AC_RESIZE_SOFTCUBIC75, AC_RESIZE_SOFTCUBIC100,
    AC_RESIZE_LANCZOS2, AC_RESIZE_LANCZOS3, AC_RESIZE_LANCZOS4, AC_RESIZE_SPLINE16, AC_RESIZE_SPLINE36, AC_RESIZE_SPLINE64,
    AC_RESIZE_BILINEAR
};

typedef struct ACImage
{
    int width;
    int height;
    int channels;
    int stride;                 /* bytes per row */
    int element_type;           /* enum ACImageElementType */
    void* ptr;                  /* first pixel */
    struct ACImageHandle* hptr; /* library-owned */
} ACImage;

/* lifetime: alloc = zeroed malloc; free = unref + free + NULL the caller's pointer */
AC_C_API ACImage* ac_image_alloc(void);
AC_C_API void ac_image_free(ACImage** image);
AC_C_API int ac_image_ref(const ACImage* src, ACImage* dst);
AC_C_API void ac_image_unref(ACImage* image);
/* storage: each reads width/height/channels/element_type/stride (and ptr for map) from the struct */
AC_C_API int ac_image_create(ACImage* image);
AC_C_API int ac_image_map(ACImage* image);
AC_C_API int ac_image_from(ACImage* image, const void* data);
AC_C_API int ac_image_view(const ACImage* src, ACImage* dst, int x, int y, int w, int h);
AC_C_API int ac_image_clone(const ACImage* src, ACImage* dst);
AC_C_API int ac_image_to(const ACImage* image, void* data, int stride);
/* file I/O (binding/c/include/AC/Core/Image.h:75-76): PNG / BMP / PNM / TGA in, .png / .bmp / .tga out; -AC_EIO when the file cannot
 * be read, decoded or written (JPEG is not supported by this build's codecs) */
AC_C_API int ac_imread(const char* filename, int mode, ACImage* image);
AC_C_API int ac_imwrite(const char* filename, const ACImage* image);
/* image operations; both images need a handle (create/map/from/...) */
AC_C_API int ac_resize(const ACImage* src, ACImage* dst, double fx, double fy, int mode);
AC_C_API int ac_rgb2yuv(const ACImage* rgb, ACImage* yuv);
AC_C_API int ac_rgba2yuva(const ACImage* rgba, ACImage* yuva);
AC_C_API int ac_yuv2rgb(const ACImage* yuv, ACImage* rgb);
AC_C_API int ac_yuva2rgba(const ACImage* yuva, ACImage* rgba);

static inline int ac_image_size(const ACImage* image) { return image->height * image->stride; }
static inline int ac_image_element_size(const ACImage* image) { return image->element_type & 0xff; }
static inline int ac_image_pixel_size(const ACImage* image) { return image->channels * ac_image_element_size(image); }
static inline uint8_t* ac_image_data(const ACImage* image) { return (uint8_t*)image->ptr; }
static inline uint8_t* ac_image_line(const ACImage* image, int y) { return ac_image_data(image) + (ptrdiff_t)image->stride * y; }
static inline uint8_t* ac_image_pixel(const ACImage* image, int x, int y) { return ac_image_line(image, y) + x * ac_image_pixel_size(image); }
static inline int ac_image_empty(const ACImage* image) { return image->ptr == NULL; }
static inline int ac_image_is_uint(const ACImage* image) { return (image->element_type >> 8) == 0; }
static inline int ac_image_is_int(const ACImage* image) { return (image->element_type >> 8) == 1; }
static inline int ac_image_is_float(const ACImage* image) { return (image->element_type >> 8) == 2; }

#endif
