// Symbol visibility for the B200 drop-in of libac_core (the reference generates this header with CMake,
// core/CMakeLists.txt:356-360).
#pragma once
#if defined(__GNUC__)
#   define AC_CORE_EXPORT __attribute__((visibility("default")))
#   define AC_CORE_NO_EXPORT __attribute__((visibility("hidden")))
#else
#   define AC_CORE_EXPORT
#   define AC_CORE_NO_EXPORT
#endif
#define AC_CORE_WITH_CUDA 1
#define AC_CORE_DISABLE_IMAGE_IO 1
#ifndef AC_CORE_VERSION_STR
#   define AC_CORE_VERSION_STR "3.2.0-b200"
#endif
