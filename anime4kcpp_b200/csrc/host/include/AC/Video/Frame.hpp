// ac::video::Frame as the reference defines it (video/include/AC/Video/Pipeline.hpp:18-35) and the per-frame upscale the
// reference's video callers write by hand (cli/src/Main.cpp:183-206, filter/vapoursynth/src/Filter.cpp:31-42): luma plane
// through Processor::process (with the shl / shr bit-depth normalisation), every other plane through the Catmull-Rom
// ac::core::resize.  The decode / encode pipeline itself (FFmpeg) is outside this backend; this is the part of it that runs
// on the GPU, as one submission.
#pragma once

#include <cstdint>
#include <memory>

#include "AC/Core/Processor.hpp"

#include "ACCoreExport.hpp"

namespace ac::video
{
    struct Frame
    {
        struct {
            int width, height, channel, stride;
            std::uint8_t* data;
        } plane[3];
        int planes;
        // same encoding as ac::core::Image::ElementType
        int elementType;
        // one based number
        std::int64_t number;
        // keeps decoder-owned memory alive; untouched here
        std::shared_ptr<struct FrameData> dptr;

        bool operator<(const Frame& other) const noexcept { return number < other.number; }
        bool operator>(const Frame& other) const noexcept { return number > other.number; }
    };

    // dst planes are caller-allocated (Pipeline::request does that in the reference).  Returns processor.ok(); never throws.
    AC_CORE_EXPORT bool upscale(core::Processor& processor, const Frame& src, Frame& dst, double factor, int shift = 0) noexcept;
}
