// Model descriptors: ACNetLegacy, ACNet<8>, ARNet<8> (the hot path) and ArtCNN<16/32>, FSRCNNX<8/16>.
//
// Same public accessors as the reference's CRTP descriptors (core/include/AC/Core/Model/Base.hpp:15-63,
// ACNet.hpp:18-116, ARNet.hpp:16-72): flat fp32 arrays plus per-layer lengths / offsets, so code written against
// `model.kernel(l)`, `model.bias(l)`, `model.alpha(l)`, `kernels()`, `blocks()` keeps working.  The numbers behind
// ACNetLegacy and ACNet<8> are the reference's own tables (dumped into weights/acnet.bin by tools/gen_weights.cpp);
// ARNet<8> gets seeded stand-ins because the reference's ARNet.p is a missing blob (csrc/synth_weights.h).
// ArtCNN / FSRCNNX (core/include/AC/Core/Model/{ArtCNN,FSRCNNX}.hpp) use the reference's tables too.
#pragma once

#include <algorithm>
#include <cstddef>

#include "ACCoreExport.hpp"

namespace ac::core::model
{
    namespace detail
    {
        // what the three families share; Derived supplies the layer arithmetic
        template<typename Derived>
        class Descriptor
        {
        public:
            std::size_t kernelSize() const noexcept { return self().kernelLength() * sizeof(float); }
            std::size_t kernelSize(const int layer) const noexcept { return self().kernelLength(layer) * sizeof(float); }
            std::size_t biasSize() const noexcept { return self().biasLength() * sizeof(float); }
            std::size_t biasSize(const int layer) const noexcept { return self().biasLength(layer) * sizeof(float); }
            std::size_t alphaSize() const noexcept { return self().alphaLength() * sizeof(float); }
            std::size_t alphaSize(const int layer) const noexcept { return self().alphaLength(layer) * sizeof(float); }
            int blocks() const noexcept { return blockNum; }
            const float* kernel(const int layer = 0) const noexcept { return kptr + self().kernelOffset(layer); }
            const float* bias(const int layer = 0) const noexcept { return bptr + self().biasOffset(layer); }
            const float* alpha(const int layer = 0) const noexcept { return aptr + self().alphaOffset(layer); }
            // canonical model string ("acnet-f8b8-hdn", ...)
            const char* name() const noexcept { return canonical; }

        protected:
            const Derived& self() const noexcept { return *static_cast<const Derived*>(this); }
            // binds kptr/bptr/aptr to the named variant's arrays; false if unknown
            AC_CORE_EXPORT bool bind(const char* canonicalName) noexcept;

            int blockNum = 0;
            const float* kptr = nullptr;
            const float* bptr = nullptr;
            const float* aptr = nullptr;
            const char* canonical = "";
        };
    }

    class ACNetLegacy : public detail::Descriptor<ACNetLegacy>
    {
    public:
        enum class Variant { GAN, HDN0, HDN1, HDN2, HDN3 };
        AC_CORE_EXPORT ACNetLegacy(Variant v) noexcept;

        int kernels() const noexcept { return blockNum + 2; }   // head, blockNum 8->8 convs, deconv
        int biases() const noexcept { return blockNum + 1; }
        int alphas() const noexcept { return 0; }
        int kernelLength() const noexcept { return 72 + 576 * blockNum + 32; }
        int kernelLength(const int l) const noexcept { return l == 0 ? 72 : (l >= 1 && l <= blockNum) ? 576 : (l == blockNum + 1 ? 32 : 0); }
        int biasLength() const noexcept { return 8 * (blockNum + 1); }
        int biasLength(const int l) const noexcept { return (l >= 0 && l <= blockNum) ? 8 : 0; }
        int alphaLength() const noexcept { return 0; }
        int alphaLength(const int) const noexcept { return 0; }
        int kernelIndex(const int l) const noexcept { return std::clamp(l, 0, kernels() - 1); }
        int biasIndex(const int l) const noexcept { return std::clamp(l, 0, biases() - 1); }
        int alphaIndex(const int) const noexcept { return 0; }
        int kernelLayer(const int i) const noexcept { return std::clamp(i, 0, kernels() - 1); }
        int biasLayer(const int i) const noexcept { return std::clamp(i, 0, biases() - 1); }
        int alphaLayer(const int) const noexcept { return 0; }
        int kernelOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < kernels() ? 72 + 576 * (l - 1) : kernelLength()); }
        int biasOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < biases() ? 8 * l : biasLength()); }
        int alphaOffset(const int) const noexcept { return 0; }
    };

    template<int F>
    class ACNet : public detail::Descriptor<ACNet<F>>
    {
    public:
        enum class Variant
        {
            B4_NORMAL, B4_HDN, B4_BOX, B4_BOX_HDN,
            B8_NORMAL, B8_HDN, B8_BOX, B8_BOX_HDN,
            B18_NORMAL, B18_HDN, B18_BOX, B18_BOX_HDN
        };
        AC_CORE_EXPORT ACNet(Variant v) noexcept;

        int kernels() const noexcept { return this->blockNum + 2; }  // head, blockNum convs, 8->4 shuffle conv
        int biases() const noexcept { return this->blockNum + 2; }
        int alphas() const noexcept { return this->blockNum + 1; }
        int kernelLength() const noexcept { return F * 9 + F * F * 9 * this->blockNum + F * 4 * 9; }
        int kernelLength(const int l) const noexcept { return l == 0 ? F * 9 : (l >= 1 && l <= this->blockNum) ? F * F * 9 : (l == this->blockNum + 1 ? F * 4 * 9 : 0); }
        int biasLength() const noexcept { return F * (this->blockNum + 1) + 4; }
        int biasLength(const int l) const noexcept { return (l >= 0 && l <= this->blockNum) ? F : (l == this->blockNum + 1 ? 4 : 0); }
        int alphaLength() const noexcept { return F * (this->blockNum + 1); }
        int alphaLength(const int l) const noexcept { return (l >= 0 && l <= this->blockNum) ? F : 0; }
        int kernelIndex(const int l) const noexcept { return std::clamp(l, 0, kernels() - 1); }
        int biasIndex(const int l) const noexcept { return std::clamp(l, 0, biases() - 1); }
        int alphaIndex(const int l) const noexcept { return std::clamp(l, 0, alphas() - 1); }
        int kernelLayer(const int i) const noexcept { return std::clamp(i, 0, kernels() - 1); }
        int biasLayer(const int i) const noexcept { return std::clamp(i, 0, biases() - 1); }
        int alphaLayer(const int i) const noexcept { return std::clamp(i, 0, alphas() - 1); }
        int kernelOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < kernels() ? F * 9 + F * F * 9 * (l - 1) : kernelLength()); }
        int biasOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < biases() ? F * l : biasLength()); }
        int alphaOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < alphas() ? F * l : alphaLength()); }
    };

    template<int F>
    class ARNet : public detail::Descriptor<ARNet<F>>
    {
    public:
        enum class Variant
        {
            B8_NORMAL, B8_HDN, B8_BOX, B8_BOX_HDN,
            B16_NORMAL, B16_HDN, B16_BOX, B16_BOX_HDN,
            B32_NORMAL, B32_HDN, B32_BOX, B32_BOX_HDN,
            B64_NORMAL, B64_HDN, B64_BOX, B64_BOX_HDN
        };
        AC_CORE_EXPORT ARNet(Variant v) noexcept;

        // layers: head, 2*blockNum 3x3 convs, the 1x1 fuse, the 8->4 shuffle conv
        int kernels() const noexcept { return this->blockNum * 2 + 3; }
        int biases() const noexcept { return this->blockNum * 2 + 3; }
        int alphas() const noexcept { return this->blockNum + 1; }
        int kernelLength() const noexcept { return F * 9 + F * F * 9 * this->blockNum * 2 + F * F + F * 4 * 9; }
        int kernelLength(const int l) const noexcept
        {
            const int body = this->blockNum * 2;
            return l == 0 ? F * 9 : (l >= 1 && l <= body) ? F * F * 9 : (l == body + 1 ? F * F : (l == body + 2 ? F * 4 * 9 : 0));
        }
        int biasLength() const noexcept { return F * (this->blockNum * 2 + 2) + 4; }
        int biasLength(const int l) const noexcept { const int body = this->blockNum * 2; return (l >= 0 && l <= body + 1) ? F : (l == body + 2 ? 4 : 0); }
        int alphaLength() const noexcept { return F * (this->blockNum + 1); }
        int alphaLength(const int l) const noexcept { return (l >= 1 && l <= this->blockNum * 2 + 1 && (l & 1)) ? F : 0; }
        int kernelIndex(const int l) const noexcept { return std::clamp(l, 0, kernels() - 1); }
        int biasIndex(const int l) const noexcept { return std::clamp(l, 0, biases() - 1); }
        int alphaIndex(const int l) const noexcept { return (std::clamp(l, 1, this->blockNum * 2 + 1) - 1) / 2; }  // PReLU sits on the odd layers
        int kernelLayer(const int i) const noexcept { return std::clamp(i, 0, kernels() - 1); }
        int biasLayer(const int i) const noexcept { return std::clamp(i, 0, biases() - 1); }
        int alphaLayer(const int i) const noexcept { return std::clamp(i, 0, alphas() - 1) * 2 + 1; }
        int kernelOffset(const int l) const noexcept
        {
            const int body = this->blockNum * 2;
            if (l <= 0) return 0;
            if (l <= body + 1) return F * 9 + F * F * 9 * (l - 1);
            if (l < kernels()) return F * 9 + F * F * 9 * body + F * F;
            return kernelLength();
        }
        int biasOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < biases() ? F * l : biasLength()); }
        int alphaOffset(const int l) const noexcept { const int i = alphaIndex(l); return i <= 0 ? 0 : (i < alphas() ? F * i : alphaLength()); }
    };
    // head 1->F (3x3), blockNum + 1 convs F->F (the last one adds the head output), F->4 shuffle conv
    // (reference: core/include/AC/Core/Model/ArtCNN.hpp:16-61)
    template<int F>
    class ArtCNN : public detail::Descriptor<ArtCNN<F>>
    {
    public:
        enum class Variant { C4_NORMAL, C4_DN, C4_DS };
        AC_CORE_EXPORT ArtCNN(Variant v) noexcept;

        int kernels() const noexcept { return this->blockNum + 3; }
        int biases() const noexcept { return this->blockNum + 3; }
        int alphas() const noexcept { return 0; }
        int kernelLength() const noexcept { return F * 9 + F * F * 9 * (this->blockNum + 1) + F * 4 * 9; }
        int kernelLength(const int l) const noexcept { return l == 0 ? F * 9 : (l >= 1 && l <= this->blockNum + 1) ? F * F * 9 : (l == this->blockNum + 2 ? F * 4 * 9 : 0); }
        int biasLength() const noexcept { return F * (this->blockNum + 2) + 4; }
        int biasLength(const int l) const noexcept { return (l >= 0 && l <= this->blockNum + 1) ? F : (l == this->blockNum + 2 ? 4 : 0); }
        int alphaLength() const noexcept { return 0; }
        int alphaLength(const int) const noexcept { return 0; }
        int kernelIndex(const int l) const noexcept { return std::clamp(l, 0, kernels() - 1); }
        int biasIndex(const int l) const noexcept { return std::clamp(l, 0, biases() - 1); }
        int alphaIndex(const int) const noexcept { return 0; }
        int kernelLayer(const int i) const noexcept { return std::clamp(i, 0, kernels() - 1); }
        int biasLayer(const int i) const noexcept { return std::clamp(i, 0, biases() - 1); }
        int alphaLayer(const int) const noexcept { return 0; }
        int kernelOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < kernels() ? F * 9 + F * F * 9 * (l - 1) : kernelLength()); }
        int biasOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < biases() ? F * l : biasLength()); }
        int alphaOffset(const int) const noexcept { return 0; }
    };

    // head 1->F (5x5), blockNum convs F->F with PReLU, a 1x1 F->F (+ head output, PReLU), F->4 shuffle conv
    // (reference: core/include/AC/Core/Model/FSRCNNX.hpp:16-85)
    template<int F>
    class FSRCNNX : public detail::Descriptor<FSRCNNX<F>>
    {
    public:
        enum class Variant { B4_NORMAL, B4_DISTORT_PLUS };
        AC_CORE_EXPORT FSRCNNX(Variant v) noexcept;

        int kernels() const noexcept { return this->blockNum + 3; }
        int biases() const noexcept { return this->blockNum + 3; }
        int alphas() const noexcept { return this->blockNum + 1; }
        int kernelLength() const noexcept { return F * 25 + F * F * 9 * this->blockNum + F * F + F * 4 * 9; }
        int kernelLength(const int l) const noexcept
        {
            return l == 0 ? F * 25 : (l >= 1 && l <= this->blockNum) ? F * F * 9 : (l == this->blockNum + 1 ? F * F : (l == this->blockNum + 2 ? F * 4 * 9 : 0));
        }
        int biasLength() const noexcept { return F * (this->blockNum + 2) + 4; }
        int biasLength(const int l) const noexcept { return (l >= 0 && l <= this->blockNum + 1) ? F : (l == this->blockNum + 2 ? 4 : 0); }
        int alphaLength() const noexcept { return F * (this->blockNum + 1); }
        int alphaLength(const int l) const noexcept { return (l >= 1 && l <= this->blockNum + 1) ? F : 0; }
        int kernelIndex(const int l) const noexcept { return std::clamp(l, 0, kernels() - 1); }
        int biasIndex(const int l) const noexcept { return std::clamp(l, 0, biases() - 1); }
        int alphaIndex(const int l) const noexcept { return std::clamp(l, 1, alphas()) - 1; }
        int kernelLayer(const int i) const noexcept { return std::clamp(i, 0, kernels() - 1); }
        int biasLayer(const int i) const noexcept { return std::clamp(i, 0, biases() - 1); }
        int alphaLayer(const int i) const noexcept { return std::clamp(i, 0, alphas() - 1) + 1; }
        int kernelOffset(const int l) const noexcept
        {
            if (l <= 0) return 0;
            if (l <= this->blockNum + 1) return F * 25 + F * F * 9 * (l - 1);
            if (l < kernels()) return F * 25 + F * F * 9 * this->blockNum + F * F;
            return kernelLength();
        }
        int biasOffset(const int l) const noexcept { return l <= 0 ? 0 : (l < biases() ? F * l : biasLength()); }
        int alphaOffset(const int l) const noexcept { const int i = alphaIndex(l); return i <= 0 ? 0 : (i < alphas() ? F * i : alphaLength()); }
    };
}
