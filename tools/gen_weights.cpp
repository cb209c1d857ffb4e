// Builds anime4kcpp_b200/weights/acnet.bin from the reference's weight tables (ACNet.p, ArtCNN.p, FSRCNNX.p).
//
// The reference defines its trained numbers as `constexpr float X_NHWC_{kernels,biases,alphas}[]`
// in core/internal/AC/Core/Internal/Model/Param/ACNet.p (bound to model variants in
// core/src/Model.cpp:8-127).  This tool #includes that file from the reference tree where it
// lies (so the compiler's own literal parsing yields the exact fp32 values the reference
// uses) and dumps the arrays into a small binary container.  Only the resulting data blob
// is committed; no reference source is copied.
//
//   g++ -std=c++17 -I/root/reference/core/internal tools/gen_weights.cpp -o /tmp/gen_weights
//   /tmp/gen_weights anime4kcpp_b200/weights/acnet.bin
//
// Container (little endian):
//   char     magic[8] = "ACB2WTS1"
//   uint32   n_models, reserved
//   n_models x { char name[48]; uint32 family, blocks, nk, nb, na, offset }   (offset in floats)
//   family: 0 ACNetLegacy, 1 ACNet<8>, 3 ArtCNN<F>, 4 FSRCNNX<F>; for families 3 and 4 `blocks` = blocks | (F << 16)
//   float    data[]
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define AC_CORE_PARAM_ALIGN 64
namespace p
{
#include "AC/Core/Internal/Model/Param/ACNet.p"
#include "AC/Core/Internal/Model/Param/ArtCNN.p"
#include "AC/Core/Internal/Model/Param/FSRCNNX.p"
}

struct Entry
{
    char name[48];
    std::uint32_t family, blocks, nk, nb, na, offset;
};

static std::vector<Entry> entries;
static std::vector<float> data;

template <std::size_t NK, std::size_t NB>
static void addLegacy(const char* name, const float (&k)[NK], const float (&b)[NB])
{
    Entry e{};
    std::strncpy(e.name, name, sizeof(e.name) - 1);
    e.family = 0; e.blocks = 8; e.nk = NK; e.nb = NB; e.na = 0; e.offset = static_cast<std::uint32_t>(data.size());
    data.insert(data.end(), k, k + NK);
    data.insert(data.end(), b, b + NB);
    entries.push_back(e);
}
template <std::size_t NK, std::size_t NB, std::size_t NA>
static void addACNet(const char* name, int blocks, const float (&k)[NK], const float (&b)[NB], const float (&a)[NA])
{
    Entry e{};
    std::strncpy(e.name, name, sizeof(e.name) - 1);
    e.family = 1; e.blocks = blocks; e.nk = NK; e.nb = NB; e.na = NA; e.offset = static_cast<std::uint32_t>(data.size());
    data.insert(data.end(), k, k + NK);
    data.insert(data.end(), b, b + NB);
    data.insert(data.end(), a, a + NA);
    entries.push_back(e);
}

template <std::size_t NK, std::size_t NB>
static void addArtCNN(const char* name, int channels, const float (&k)[NK], const float (&b)[NB])
{
    Entry e{};
    std::strncpy(e.name, name, sizeof(e.name) - 1);
    e.family = 3; e.blocks = 4u | (static_cast<std::uint32_t>(channels) << 16); e.nk = NK; e.nb = NB; e.na = 0; e.offset = static_cast<std::uint32_t>(data.size());
    data.insert(data.end(), k, k + NK);
    data.insert(data.end(), b, b + NB);
    entries.push_back(e);
}
template <std::size_t NK, std::size_t NB, std::size_t NA>
static void addFSRCNNX(const char* name, int channels, const float (&k)[NK], const float (&b)[NB], const float (&a)[NA])
{
    Entry e{};
    std::strncpy(e.name, name, sizeof(e.name) - 1);
    e.family = 4; e.blocks = 4u | (static_cast<std::uint32_t>(channels) << 16); e.nk = NK; e.nb = NB; e.na = NA; e.offset = static_cast<std::uint32_t>(data.size());
    data.insert(data.end(), k, k + NK);
    data.insert(data.end(), b, b + NB);
    data.insert(data.end(), a, a + NA);
    entries.push_back(e);
}

#define ARTCNN(name, F, V) addArtCNN(name, F, p::ArtCNN_##V##_NHWC_kernels, p::ArtCNN_##V##_NHWC_biases)
#define FSRCNNX(name, F, V) addFSRCNNX(name, F, p::FSRCNNX_##V##_NHWC_kernels, p::FSRCNNX_##V##_NHWC_biases, p::FSRCNNX_##V##_NHWC_alphas)
#define LEGACY(name, V) addLegacy(name, p::ACNetLegacy_##V##_NHWC_kernels, p::ACNetLegacy_##V##_NHWC_biases)
#define ACNET(name, B, V) addACNet(name, B, p::ACNet_##V##_NHWC_kernels, p::ACNet_##V##_NHWC_biases, p::ACNet_##V##_NHWC_alphas)

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 1; }
    LEGACY("acnet-legacy-gan", GAN);
    LEGACY("acnet-legacy-hdn0", HDN0);
    LEGACY("acnet-legacy-hdn1", HDN1);
    LEGACY("acnet-legacy-hdn2", HDN2);
    LEGACY("acnet-legacy-hdn3", HDN3);
    ACNET("acnet-f8b4", 4, F8B4);
    ACNET("acnet-f8b4-hdn", 4, F8B4_HDN);
    ACNET("acnet-f8b4-box", 4, F8B4_Box);
    ACNET("acnet-f8b4-box-hdn", 4, F8B4_Box_HDN);
    ACNET("acnet-f8b8", 8, F8B8);
    ACNET("acnet-f8b8-hdn", 8, F8B8_HDN);
    ACNET("acnet-f8b8-box", 8, F8B8_Box);
    ACNET("acnet-f8b8-box-hdn", 8, F8B8_Box_HDN);
    ACNET("acnet-f8b18", 18, F8B18);
    ACNET("acnet-f8b18-hdn", 18, F8B18_HDN);
    ACNET("acnet-f8b18-box", 18, F8B18_Box);
    ACNET("acnet-f8b18-box-hdn", 18, F8B18_Box_HDN);
    ARTCNN("artcnn-c4f16", 16, C4F16);
    ARTCNN("artcnn-c4f16-dn", 16, C4F16_DN);
    ARTCNN("artcnn-c4f16-ds", 16, C4F16_DS);
    ARTCNN("artcnn-c4f32", 32, C4F32);
    ARTCNN("artcnn-c4f32-dn", 32, C4F32_DN);
    ARTCNN("artcnn-c4f32-ds", 32, C4F32_DS);
    FSRCNNX("fsrcnnx-f8b4", 8, F8);
    FSRCNNX("fsrcnnx-f8b4-distort-plus", 8, F8_DistortPlus);
    FSRCNNX("fsrcnnx-f16b4", 16, F16);
    FSRCNNX("fsrcnnx-f16b4-distort-plus", 16, F16_DistortPlus);

    std::FILE* f = std::fopen(argv[1], "wb");
    if (!f) { std::perror("fopen"); return 1; }
    const char magic[8] = { 'A','C','B','2','W','T','S','1' };
    std::uint32_t n = static_cast<std::uint32_t>(entries.size()), reserved = 0;
    std::fwrite(magic, 1, 8, f);
    std::fwrite(&n, 4, 1, f);
    std::fwrite(&reserved, 4, 1, f);
    std::fwrite(entries.data(), sizeof(Entry), entries.size(), f);
    std::fwrite(data.data(), sizeof(float), data.size(), f);
    std::fclose(f);
    std::printf("%u models, %zu floats\n", n, data.size());
    for (auto& e : entries) std::printf("  %-22s family %u blocks %2u  k %5u b %3u a %3u  @%u\n", e.name, e.family, e.blocks, e.nk, e.nb, e.na, e.offset);
    return 0;
}
