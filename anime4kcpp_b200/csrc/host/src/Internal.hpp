// Host-internal helpers shared by the ac::core translation units.
#pragma once

#include "../../../../include/acb200.h"

namespace ac::core::internal
{
    // per-thread session on the fastest device for the free image functions (rgb2yuv, resize, ...); nullptr
    // when no CUDA device is usable -- there is no CPU path behind them
    acb200_session* threadSession() noexcept;
}
