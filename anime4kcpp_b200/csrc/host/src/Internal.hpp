// Host-internal helpers shared by the ac::core translation units.
#pragma once

#include "../../../../include/acb200.h"

namespace ac::core::internal
{
    // per-thread session on the fastest device for the free image functions (rgb2yuv, resize, ...); nullptr
    // when no CUDA device is usable -- there is no CPU path behind them
    acb200_session* threadSession() noexcept;
    // status of the calling thread's most recent free image function (0 = ok): the reference's free functions return void, so the
    // C and Python bindings read this to turn a failed or unsupported operation into an error code / exception instead of handing
    // back an image nobody wrote
    int& lastOpStatus() noexcept;
}
