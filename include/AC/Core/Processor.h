/*
 * C processor API of the B200 drop-in.  Binary-compatible with the reference's libac_c
 * (binding/c/include/AC/Core/Processor.h:8-36, implemented in binding/c/src/Binding.cpp:161-247).
 */
#ifndef AC_BINDING_C_CORE_PROCESSOR_H
#define AC_BINDING_C_CORE_PROCESSOR_H

#include <stdint.h>

#include "AC/Core/Image.h"

enum ACProcessorType { AC_PROCESSOR_CPU = 0, AC_PROCESSOR_OPENCL = 1, AC_PROCESSOR_CUDA = 2 };

typedef struct ACProcessor
{
    int device;                     /* device index; out of range = fastest */
    const char* type;               /* "auto" | "cuda" (this build has no "cpu" / "opencl") */
    const char* model;              /* e.g. "acnet-legacy-hdn0", "acnet-f8b8-hdn", "arnet-f8b64" */
    struct ACProcessorHandle* hptr; /* library-owned */
} ACProcessor;

AC_C_API ACProcessor* ac_processor_alloc(void);
AC_C_API void ac_processor_free(ACProcessor** processor);
AC_C_API int ac_processor_ref(const ACProcessor* src, ACProcessor* dst);
AC_C_API void ac_processor_unref(ACProcessor* processor);
/* builds the processor from device/type/model; returns ac_processor_ok() */
AC_C_API int ac_processor_create(ACProcessor* processor);
/* src and dst both need a handle; dst's plain fields are refreshed afterwards; returns ac_processor_ok() */
AC_C_API int ac_processor_process(ACProcessor* processor, const ACImage* src, ACImage* dst, double factor);
/*
 * Extension (not in the reference's libac_c): one planar / semi-planar YUV video frame -- plane 0 (luma) through the network
 * with the shl / shr bit-depth normalisation, the other planes through the Catmull-Rom resize -- as ONE GPU submission;
 * the body of the reference's per-frame video callback (cli/src/Main.cpp:183-206).  ACPlane mirrors one entry of
 * ac::video::Frame::plane.  dst planes are caller-allocated.  Returns ac_processor_ok().
 */
typedef struct ACPlane
{
    int width, height, channel, stride;
    uint8_t* data;
} ACPlane;
AC_C_API int ac_processor_process_frame(ACProcessor* processor, const ACPlane* src, const ACPlane* dst, int planes, int element_type, int shift, double factor);
AC_C_API int ac_processor_ok(const ACProcessor* processor);
AC_C_API const char* ac_processor_error(const ACProcessor* processor);
AC_C_API const char* ac_processor_name(const ACProcessor* processor);
AC_C_API int ac_processor_type(const ACProcessor* processor);
AC_C_API const char* ac_processor_type_name(const ACProcessor* processor);
AC_C_API const char* ac_processor_info(int processor_type);
AC_C_API const char* ac_processor_list_info(void);

#endif
