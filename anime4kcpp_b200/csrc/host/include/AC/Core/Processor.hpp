// ac::core::Processor -- the drop-in front door (reference: core/include/AC/Core/Processor.hpp:15-48).
//
// Public surface identical to the reference; one deliberate, non-public difference: the whole-image call
// process(src, dst, factor) is routed through a protected virtual (`processImage`) so the CUDA backend can own
// the colour split, chroma resize and merge on the GPU in the same submission as the luma network (the
// reference runs those three on the CPU in its non-virtual driver, core/src/processor/Processor.cpp:199-276).
#pragma once

#include <memory>

#include "AC/Core/Image.hpp"

#include "ACCoreExport.hpp"

namespace ac::core
{
    class AC_CORE_EXPORT Processor
    {
    public:
        static constexpr int CPU = 0;
        static constexpr int OpenCL = 1;
        static constexpr int CUDA = 2;

        AC_CORE_EXPORT Processor() noexcept;
        AC_CORE_EXPORT virtual ~Processor();

        AC_CORE_EXPORT Image process(const Image& src, double factor);
        // a non-empty `dst` is trusted to be correctly allocated and is written in place
        AC_CORE_EXPORT void process(const Image& src, Image& dst, double factor);

        AC_CORE_EXPORT virtual bool ok() noexcept;
        AC_CORE_EXPORT virtual const char* error() noexcept;
        AC_CORE_EXPORT virtual const char* name() const noexcept = 0;
        AC_CORE_EXPORT virtual int type() const noexcept = 0;
        AC_CORE_EXPORT virtual const char* typeName() const noexcept = 0;

        AC_CORE_EXPORT static std::shared_ptr<Processor> create(const char* type, int device, const char* model);
        AC_CORE_EXPORT static const char* listInfo();

        template<int type, typename Model> static std::shared_ptr<Processor> create(int idx, const Model& model);
        template<int type> static const char* info();

    protected:
        // dst is non-empty and correctly sized when this is called
        virtual void processImage(const Image& src, Image& dst, double factor) = 0;

        int idx;
    };
}
