// Umbrella header, same role as the reference's core/include/AC/Core.hpp.
#pragma once
#include "AC/Core/Image.hpp"
#include "AC/Core/Model.hpp"
#include "AC/Core/Processor.hpp"
