// Binds the model descriptors to their weight arrays (reference: core/src/Model.cpp:8-230 binds variants to the
// constexpr tables of core/internal/AC/Core/Internal/Model/Param/*.p).  Here the ACNet tables come from the
// embedded blob written by tools/gen_weights.cpp; the ARNet arrays are generated once from synth_weights.h
// because the reference's ARNet.p is a missing blob.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "AC/Core/Model.hpp"

#include "../../synth_weights.h"

extern "C" const unsigned char acb200_weights_blob[];
extern "C" const unsigned char acb200_weights_blob_end[];

namespace
{
    struct Arrays
    {
        int blocks = 0;
        const float* k = nullptr; const float* b = nullptr; const float* a = nullptr;
        std::vector<float> owned;   // ARNet stand-ins live here
    };

    struct BlobEntry
    {
        char name[48];
        std::uint32_t family, blocks, nk, nb, na, offset;
    };

    const Arrays* lookup(const char* canonical)
    {
        static std::mutex mutex;
        static std::map<std::string, std::unique_ptr<Arrays>> table;
        std::lock_guard<std::mutex> lock(mutex);
        if (table.empty())
        {
            const unsigned char* p = acb200_weights_blob;
            if (acb200_weights_blob_end - p >= 16 && std::memcmp(p, "ACB2WTS1", 8) == 0)
            {
                std::uint32_t n;
                std::memcpy(&n, p + 8, 4);
                const auto* entries = reinterpret_cast<const BlobEntry*>(p + 16);
                const auto* data = reinterpret_cast<const float*>(p + 16 + sizeof(BlobEntry) * n);
                for (std::uint32_t i = 0; i < n; i++)
                {
                    auto arr = std::make_unique<Arrays>();
                    arr->blocks = static_cast<int>(entries[i].blocks & 0xffffu);       // upper half: feature count of ArtCNN / FSRCNNX
                    arr->k = data + entries[i].offset;
                    arr->b = arr->k + entries[i].nk;
                    arr->a = entries[i].na ? arr->b + entries[i].nb : nullptr;
                    table[entries[i].name] = std::move(arr);
                }
            }
        }
        auto it = table.find(canonical);
        if (it != table.end()) return it->second.get();
        int blocks = 0;
        if (std::sscanf(canonical, "arnet-f8b%d", &blocks) == 1 && (blocks == 8 || blocks == 16 || blocks == 32 || blocks == 64))
        {
            auto arr = std::make_unique<Arrays>();
            const int nk = acsw_arnet_kernel_len(blocks), nb = acsw_arnet_bias_len(blocks), na = acsw_arnet_alpha_len(blocks);
            arr->owned.resize(static_cast<std::size_t>(nk) + nb + na);
            acsw_fill_arnet(canonical, blocks, arr->owned.data(), arr->owned.data() + nk, arr->owned.data() + nk + nb);
            arr->blocks = blocks;
            arr->k = arr->owned.data(); arr->b = arr->k + nk; arr->a = arr->b + nb;
            auto* raw = arr.get();
            table[canonical] = std::move(arr);
            return raw;
        }
        return nullptr;
    }

    // canonical strings are interned so Descriptor::name() stays valid for the process lifetime
    const char* intern(const std::string& s)
    {
        static std::mutex mutex;
        static std::map<std::string, std::unique_ptr<std::string>> pool;
        std::lock_guard<std::mutex> lock(mutex);
        auto& slot = pool[s];
        if (!slot) slot = std::make_unique<std::string>(s);
        return slot->c_str();
    }
}

template<typename Derived>
bool ac::core::model::detail::Descriptor<Derived>::bind(const char* canonicalName) noexcept
{
    const Arrays* arr = lookup(canonicalName);
    if (!arr) return false;
    blockNum = arr->blocks;
    kptr = arr->k; bptr = arr->b; aptr = arr->a;
    canonical = intern(canonicalName);
    return true;
}

ac::core::model::ACNetLegacy::ACNetLegacy(const Variant v) noexcept
{
    static const char* const names[] = { "acnet-legacy-gan", "acnet-legacy-hdn0", "acnet-legacy-hdn1", "acnet-legacy-hdn2", "acnet-legacy-hdn3" };
    bind(names[static_cast<int>(v)]);
}

template<int F>
ac::core::model::ACNet<F>::ACNet(const Variant v) noexcept
{
    static_assert(F == 8, "only the 8-feature ACNet exists");
    static const char* const size[] = { "b4", "b8", "b18" };
    static const char* const flavour[] = { "", "-hdn", "-box", "-box-hdn" };
    const int i = static_cast<int>(v);
    this->bind((std::string("acnet-f8") + size[i / 4] + flavour[i % 4]).c_str());
}

template<int F>
ac::core::model::ARNet<F>::ARNet(const Variant v) noexcept
{
    static_assert(F == 8, "only the 8-feature ARNet exists");
    static const char* const size[] = { "b8", "b16", "b32", "b64" };
    static const char* const flavour[] = { "", "-hdn", "-box", "-box-hdn" };
    const int i = static_cast<int>(v);
    this->bind((std::string("arnet-f8") + size[i / 4] + flavour[i % 4]).c_str());
}

template<int F>
ac::core::model::ArtCNN<F>::ArtCNN(const Variant v) noexcept
{
    static_assert(F == 16 || F == 32, "ArtCNN exists with 16 and 32 features");
    static const char* const flavour[] = { "", "-dn", "-ds" };
    this->bind((std::string("artcnn-c4f") + std::to_string(F) + flavour[static_cast<int>(v)]).c_str());
}

template<int F>
ac::core::model::FSRCNNX<F>::FSRCNNX(const Variant v) noexcept
{
    static_assert(F == 8 || F == 16, "FSRCNNX exists with 8 and 16 features");
    static const char* const flavour[] = { "", "-distort-plus" };
    this->bind((std::string("fsrcnnx-f") + std::to_string(F) + "b4" + flavour[static_cast<int>(v)]).c_str());
}

template class ac::core::model::detail::Descriptor<ac::core::model::ACNetLegacy>;
template class ac::core::model::detail::Descriptor<ac::core::model::ACNet<8>>;
template class ac::core::model::detail::Descriptor<ac::core::model::ARNet<8>>;
template class ac::core::model::detail::Descriptor<ac::core::model::ArtCNN<16>>;
template class ac::core::model::detail::Descriptor<ac::core::model::ArtCNN<32>>;
template class ac::core::model::detail::Descriptor<ac::core::model::FSRCNNX<8>>;
template class ac::core::model::detail::Descriptor<ac::core::model::FSRCNNX<16>>;
template class ac::core::model::ACNet<8>;
template class ac::core::model::ARNet<8>;
template class ac::core::model::ArtCNN<16>;
template class ac::core::model::ArtCNN<32>;
template class ac::core::model::FSRCNNX<8>;
template class ac::core::model::FSRCNNX<16>;
