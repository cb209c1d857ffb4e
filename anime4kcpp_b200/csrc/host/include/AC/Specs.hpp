// Model / processor catalogue of the B200 drop-in (reference: core/include/AC/Specs.hpp:22-291): the model
// families on the accelerated path, with the reference's names and parameter counts.
#pragma once

namespace ac::specs
{
    struct Model
    {
        const char* name;
        const char* description;
        int parameterCount;
        const char* version = nullptr;
        const char* author = nullptr;
        const char* homepage = nullptr;
    };
    struct Processor
    {
        const char* name;
        const char* description;
    };

    constexpr Model ModelList[] = {
        { "acnet-legacy-gan", "ACNetLegacy (ReLU, deconvolution tail), detail enhancement.", 4784 },
        { "acnet-legacy-hdn0", "ACNetLegacy, moderate denoising.", 4784 },
        { "acnet-legacy-hdn1", "ACNetLegacy, strong denoising.", 4784 },
        { "acnet-legacy-hdn2", "ACNetLegacy, aggressive denoising.", 4784 },
        { "acnet-legacy-hdn3", "ACNetLegacy, extreme denoising.", 4784 },
        { "acnet-f8b4", "ACNet 8 features x 4 blocks, neutral.", 2748 },
        { "acnet-f8b4-hdn", "ACNet f8b4, mild denoising.", 2748 },
        { "acnet-f8b4-box", "ACNet f8b4, box-degradation training.", 2748 },
        { "acnet-f8b4-box-hdn", "ACNet f8b4, box-degradation training, mild denoising.", 2748 },
        { "acnet-f8b8", "ACNet 8 features x 8 blocks, neutral.", 5116 },
        { "acnet-f8b8-hdn", "ACNet f8b8, mild denoising.", 5116 },
        { "acnet-f8b8-box", "ACNet f8b8, box-degradation training.", 5116 },
        { "acnet-f8b8-box-hdn", "ACNet f8b8, box-degradation training, mild denoising.", 5116 },
        { "acnet-f8b18", "ACNet 8 features x 18 blocks, neutral.", 11036 },
        { "acnet-f8b18-hdn", "ACNet f8b18, mild denoising.", 11036 },
        { "acnet-f8b18-box", "ACNet f8b18, box-degradation training.", 11036 },
        { "acnet-f8b18-box-hdn", "ACNet f8b18, box-degradation training, mild denoising.", 11036 },
        { "arnet-f8b8", "ARNet 8 features x 8 residual blocks (synthetic weights: ARNet.p is absent).", 9860 },
        { "arnet-f8b8-hdn", "ARNet f8b8, denoising variant (synthetic weights).", 9860 },
        { "arnet-f8b8-box", "ARNet f8b8, box variant (synthetic weights).", 9860 },
        { "arnet-f8b8-box-hdn", "ARNet f8b8, box + denoising variant (synthetic weights).", 9860 },
        { "arnet-f8b16", "ARNet 8 features x 16 residual blocks (synthetic weights).", 19268 },
        { "arnet-f8b16-hdn", "ARNet f8b16, denoising variant (synthetic weights).", 19268 },
        { "arnet-f8b16-box", "ARNet f8b16, box variant (synthetic weights).", 19268 },
        { "arnet-f8b16-box-hdn", "ARNet f8b16, box + denoising variant (synthetic weights).", 19268 },
        { "arnet-f8b32", "ARNet 8 features x 32 residual blocks (synthetic weights).", 38084 },
        { "arnet-f8b32-hdn", "ARNet f8b32, denoising variant (synthetic weights).", 38084 },
        { "arnet-f8b32-box", "ARNet f8b32, box variant (synthetic weights).", 38084 },
        { "arnet-f8b32-box-hdn", "ARNet f8b32, box + denoising variant (synthetic weights).", 38084 },
        { "arnet-f8b64", "ARNet 8 features x 64 residual blocks (synthetic weights).", 75716 },
        { "arnet-f8b64-hdn", "ARNet f8b64, denoising variant (synthetic weights).", 75716 },
        { "arnet-f8b64-box", "ARNet f8b64, box variant (synthetic weights).", 75716 },
        { "arnet-f8b64-box-hdn", "ARNet f8b64, box + denoising variant (synthetic weights).", 75716 },
        { "artcnn-c4f16", "ArtCNN C4F16, neutral.", 12340, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "artcnn-c4f16-dn", "ArtCNN C4F16, denoise and soften.", 12340, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "artcnn-c4f16-ds", "ArtCNN C4F16, denoise and sharpen.", 12340, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "artcnn-c4f32", "ArtCNN C4F32, neutral.", 47716, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "artcnn-c4f32-dn", "ArtCNN C4F32, denoise and soften.", 47716, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "artcnn-c4f32-ds", "ArtCNN C4F32, denoise and sharpen.", 47716, nullptr, "Artoriuz", "https://github.com/Artoriuz/ArtCNN" },
        { "fsrcnnx-f8b4", "FSRCNNX x2 8-0-4-1, slight denoising.", 2948, nullptr, "igv" },
        { "fsrcnnx-f8b4-distort-plus", "FSRCNNX x2 8-0-4-1 distort-plus, strong denoising.", 2948, nullptr, "nessotrin" },
        { "fsrcnnx-f16b4", "FSRCNNX x2 16-0-4-1, slight denoising.", 10628, nullptr, "igv" },
        { "fsrcnnx-f16b4-distort-plus", "FSRCNNX x2 16-0-4-1 distort-plus, strong denoising.", 10628, nullptr, "nessotrin" },
    };

    constexpr Processor ProcessorList[] = {
        { "auto", "Pick the fastest CUDA device." },
        { "cuda", "NVIDIA B200 (sm_100a) fused CNN backend." },
    };
}
