// The processor front door and the B200 CUDA backend's host class.
//
// Front door: type / model string parsing, create dispatch and listInfo follow the reference's
// core/src/processor/Processor.cpp (findProcessorType :12-25, findModel :26-187, create :286-317, listInfo
// :319-333).  Backend host class: device choice, sticky per-thread status and naming follow
// core/src/processor/cuda/CUDAProcessor.cpp (:45-50 performance score, :233-251 device index rule, :260-283).
//
// There is no CPU or OpenCL backend in this library and no fallback to one: `create("cpu", ...)` yields a
// processor whose ok() is false.
#include <algorithm>
#include <cctype>
#include <cstddef>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <thread>

#include "AC/Core/Model.hpp"
#include "AC/Core/Processor.hpp"
#include "AC/Video/Frame.hpp"

#include "../../../../include/acb200.h"
#include "Internal.hpp"

namespace
{
    using namespace ac::core;

    std::string lower(const char* s)
    {
        std::string out = s ? s : "";
        for (char& ch : out) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
        return out;
    }
    bool has(const std::string& s, const char* needle) { return s.find(needle) != std::string::npos; }

    int parseType(const char* type)
    {
        const std::string t = lower(type);
        if (t == "auto") return -1;
        if (t == "opencl") return Processor::OpenCL;
        if (t == "cuda") return Processor::CUDA;
        return Processor::CPU;
    }

    struct ModelChoice
    {
        int family;     // ACB200_FAMILY_*
        int variant;    // index into the family's Variant enum
        int features;   // F of the model template (8 except ArtCNN<16/32>, FSRCNNX<16>)
    };
    // first size token present wins, in the reference's test order
    int firstOf(const std::string& s, std::initializer_list<const char*> tokens, int fallback)
    {
        int i = 0;
        for (const char* t : tokens) { if (has(s, t)) return i; i++; }
        return fallback;
    }
    ModelChoice parseModel(const char* model)
    {
        const std::string m = lower(model);
        if (model)
        {
            // same test order as the reference (core/src/processor/Processor.cpp:39-76)
            if (has(m, "fsrcnnx")) return { ACB200_FAMILY_FSRCNNX, (has(m, "distort") || has(m, "dp")) ? 1 : 0, has(m, "f16") ? 16 : 8 };
            if (has(m, "artcnn")) return { ACB200_FAMILY_ARTCNN, has(m, "dn") ? 1 : (has(m, "ds") ? 2 : 0), has(m, "f32") ? 32 : 16 };
            const int flavour = (has(m, "box") ? 2 : 0) + (has(m, "hdn") ? 1 : 0);   // NORMAL, HDN, BOX, BOX_HDN
            if (has(m, "arnet")) return { ACB200_FAMILY_ARNET, firstOf(m, { "b8", "b16", "b32", "b64" }, 0) * 4 + flavour, 8 };
            if (has(m, "acnet"))
            {
                if (has(m, "legacy"))
                {
                    if (!has(m, "hdn")) return { ACB200_FAMILY_ACNET_LEGACY, 0, 8 };   // GAN
                    for (char ch : m) if (ch >= '0' && ch <= '3') return { ACB200_FAMILY_ACNET_LEGACY, 1 + (ch - '0'), 8 };
                    return { ACB200_FAMILY_ACNET_LEGACY, 1, 8 };                          // HDN0
                }
                return { ACB200_FAMILY_ACNET, firstOf(m, { "b4", "b8", "b18" }, 1) * 4 + flavour, 8 };
            }
        }
        return { ACB200_FAMILY_ACNET_LEGACY, 0, 8 };
    }

    // ---- a processor that only reports why it cannot run ---------------------------------------------------------
    class UnavailableProcessor final : public Processor
    {
    public:
        UnavailableProcessor(int type, const char* why) : kind(type), reason(why) {}
        bool ok() noexcept override { return false; }
        const char* error() noexcept override { return reason.c_str(); }
        const char* name() const noexcept override { return "unavailable"; }
        int type() const noexcept override { return kind; }
        const char* typeName() const noexcept override { return kind == Processor::CPU ? "CPU" : kind == Processor::OpenCL ? "OpenCL" : "CUDA"; }
    protected:
        void processImage(const Image&, Image&, double) override {}
    private:
        int kind;
        std::string reason;
    };

    // ---- the B200 backend ----------------------------------------------------------------------------------------
    class B200Processor final : public Processor
    {
    public:
        B200Processor(int device, int family, int features, int blocks, const float* k, int nk, const float* b, int nb, const float* a, int na)
        {
            const int count = acb200_device_count();
            if (count <= 0) { createError = "no CUDA device"; return; }
            idx = device;
            if (!(device >= 0 && device < count))
            {
                // out-of-range index = fastest device by clock x SM count
                long long best = -1;
                for (int i = 0; i < count; i++)
                {
                    int sms = 0, khz = 0;
                    if (acb200_device_info(i, nullptr, 0, nullptr, nullptr, &sms, &khz) != ACB200_OK) continue;
                    const long long score = static_cast<long long>(sms) * khz;
                    if (score > best) { best = score; idx = i; }
                }
            }
            char buf[256] = {};
            if (acb200_device_info(idx, buf, sizeof(buf), nullptr, nullptr, nullptr, nullptr) != ACB200_OK) { createError = "cannot query CUDA device"; return; }
            deviceName = buf;
            const int rc = acb200_model_create_wide(family, features, blocks, k, nk, b, nb, a, na, &model);
            if (rc != ACB200_OK) { createError = std::string("model rejected: ") + acb200_error_string(rc); model = nullptr; return; }
            // touch the device once so construction-time failures surface through ok(), as in the reference
            State& st = local();
            if (!st.session && createError.empty()) createError = st.error;
        }
        ~B200Processor() override
        {
            for (auto& kv : states) if (kv.second.session) acb200_session_destroy(kv.second.session);
            if (model) acb200_model_destroy(model);
        }

        bool ok() noexcept override
        {
            if (!createError.empty()) return false;
            return local().good;
        }
        const char* error() noexcept override
        {
            if (!createError.empty()) return createError.c_str();
            State& st = local();
            return st.good ? "NO ERROR" : st.error.c_str();
        }
        const char* name() const noexcept override { return deviceName.c_str(); }
        int type() const noexcept override { return Processor::CUDA; }
        const char* typeName() const noexcept override { return "CUDA"; }

        // one planar / semi-planar video frame as one submission (acb200_process_frame_host)
        bool processFrame(const acb200_plane* src, const acb200_plane* dst, const int planes, const int elementType, const int shift, const double factor)
        {
            if (!createError.empty()) return false;
            State& st = local();
            if (!st.session) return false;
            const int rc = acb200_process_frame_host(st.session, model, src, dst, planes, elementType, shift, factor);
            st.good = rc == ACB200_OK;
            if (!st.good) st.error = acb200_session_error(st.session);
            return st.good;
        }

    protected:
        void processImage(const Image& src, Image& dst, const double factor) override
        {
            if (!createError.empty()) return;
            State& st = local();
            if (!st.session) return;
            const int rc = acb200_process_host(st.session, model, src.ptr(), src.width(), src.height(), src.channels(), src.stride(), src.type(),
                                               factor, dst.ptr(), dst.stride());
            st.good = rc == ACB200_OK;
            if (!st.good) st.error = acb200_session_error(st.session);
        }

    private:
        // one session (stream + scratch) and one sticky status per calling thread, like the reference's
        // util::ThreadLocal members (CUDAProcessor.cpp:284, :374-377)
        struct State
        {
            acb200_session* session = nullptr;
            bool good = true;
            std::string error = "NO ERROR";
        };
        State& local()
        {
            const auto id = std::this_thread::get_id();
            {
                std::shared_lock<std::shared_mutex> lock(mutex);
                auto it = states.find(id);
                if (it != states.end()) return it->second;
            }
            std::unique_lock<std::shared_mutex> lock(mutex);
            State& st = states[id];
            if (!st.session)
            {
                const int rc = acb200_session_create(idx, &st.session);
                if (rc != ACB200_OK) { st.session = nullptr; st.good = false; st.error = std::string("cannot create CUDA session: ") + acb200_error_string(rc); }
            }
            return st;
        }

        std::string deviceName = "unavailable";
        std::string createError;
        acb200_model* model = nullptr;
        std::shared_mutex mutex;
        std::map<std::thread::id, State> states;
    };

    template<typename Model>
    std::shared_ptr<Processor> makeB200(int device, int family, const Model& m, int features = 8)
    {
        if (!m.kernel()) return std::make_shared<UnavailableProcessor>(Processor::CUDA, "model weights are not available in this build");
        return std::make_shared<B200Processor>(device, family, features, m.blocks(), m.kernel(), m.kernelLength(), m.bias(), m.biasLength(),
                                               m.alphaLength() ? m.alpha() : nullptr, m.alphaLength());
    }
}

ac::core::Processor::Processor() noexcept : idx(0) {}
ac::core::Processor::~Processor() = default;

ac::core::Image ac::core::Processor::process(const Image& src, const double factor)
{
    Image dst{};
    process(src, dst, factor);
    return dst;
}
void ac::core::Processor::process(const Image& src, Image& dst, const double factor)
{
    if (src.empty()) return;
    // an empty dst is allocated to (int)(w*factor) x (int)(h*factor); a non-empty one is trusted (Processor.hpp:27-29)
    if (dst.empty()) dst.create(static_cast<int>(src.width() * factor), static_cast<int>(src.height() * factor), src.channels(), src.type());
    processImage(src, dst, factor);
}
bool ac::core::Processor::ok() noexcept { return true; }
const char* ac::core::Processor::error() noexcept { return "NO ERROR"; }

template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::ACNetLegacy>(const int idx, const model::ACNetLegacy& model)
{
    return makeB200(idx, ACB200_FAMILY_ACNET_LEGACY, model);
}
template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::ACNet<8>>(const int idx, const model::ACNet<8>& model)
{
    return makeB200(idx, ACB200_FAMILY_ACNET, model);
}
template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::ARNet<8>>(const int idx, const model::ARNet<8>& model)
{
    return makeB200(idx, ACB200_FAMILY_ARNET, model);
}

template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::ArtCNN<16>>(const int idx, const model::ArtCNN<16>& model)
{
    return makeB200(idx, ACB200_FAMILY_ARTCNN, model, 16);
}
template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::ArtCNN<32>>(const int idx, const model::ArtCNN<32>& model)
{
    return makeB200(idx, ACB200_FAMILY_ARTCNN, model, 32);
}
template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::FSRCNNX<8>>(const int idx, const model::FSRCNNX<8>& model)
{
    return makeB200(idx, ACB200_FAMILY_FSRCNNX, model, 8);
}
template<> AC_CORE_EXPORT std::shared_ptr<ac::core::Processor> ac::core::Processor::create<ac::core::Processor::CUDA, ac::core::model::FSRCNNX<16>>(const int idx, const model::FSRCNNX<16>& model)
{
    return makeB200(idx, ACB200_FAMILY_FSRCNNX, model, 16);
}

std::shared_ptr<ac::core::Processor> ac::core::Processor::create(const char* type, const int device, const char* const model)
{
    const int kind = parseType(type);
    if (kind == Processor::CPU) return std::make_shared<UnavailableProcessor>(Processor::CPU, "the B200 drop-in carries no CPU backend (use \"cuda\" or \"auto\")");
    if (kind == Processor::OpenCL) return std::make_shared<UnavailableProcessor>(Processor::OpenCL, "the B200 drop-in carries no OpenCL backend (use \"cuda\" or \"auto\")");
    const int dev = kind == Processor::CUDA ? device : -1;   // auto = fastest device
    const ModelChoice choice = parseModel(model);
    switch (choice.family)
    {
    case ACB200_FAMILY_ACNET_LEGACY: return create<Processor::CUDA>(dev, model::ACNetLegacy{ static_cast<model::ACNetLegacy::Variant>(choice.variant) });
    case ACB200_FAMILY_ACNET: return create<Processor::CUDA>(dev, model::ACNet<8>{ static_cast<model::ACNet<8>::Variant>(choice.variant) });
    case ACB200_FAMILY_ARNET: return create<Processor::CUDA>(dev, model::ARNet<8>{ static_cast<model::ARNet<8>::Variant>(choice.variant) });
    case ACB200_FAMILY_ARTCNN:
        if (choice.features == 32) return create<Processor::CUDA>(dev, model::ArtCNN<32>{ static_cast<model::ArtCNN<32>::Variant>(choice.variant) });
        return create<Processor::CUDA>(dev, model::ArtCNN<16>{ static_cast<model::ArtCNN<16>::Variant>(choice.variant) });
    case ACB200_FAMILY_FSRCNNX:
        if (choice.features == 16) return create<Processor::CUDA>(dev, model::FSRCNNX<16>{ static_cast<model::FSRCNNX<16>::Variant>(choice.variant) });
        return create<Processor::CUDA>(dev, model::FSRCNNX<8>{ static_cast<model::FSRCNNX<8>::Variant>(choice.variant) });
    default: return std::make_shared<UnavailableProcessor>(Processor::CUDA, "unknown model family");
    }
}

template<> AC_CORE_EXPORT const char* ac::core::Processor::info<ac::core::Processor::CUDA>()
{
    static const std::string text = []() {
        std::ostringstream out;
        out << "CUDA:\n";
        const int count = acb200_device_count();
        for (int i = 0; i < count; i++)
        {
            char name[256] = {};
            std::size_t vram = 0;
            int cc = 0;
            if (acb200_device_info(i, name, sizeof(name), &vram, &cc, nullptr, nullptr) != ACB200_OK) continue;
            out << "  [" << i << "] " << name << " (" << (vram >> 20) << "MB, CC " << cc / 10.0 << ")\n";
        }
        return out.str();
    }();
    return text.c_str();
}
template<> AC_CORE_EXPORT const char* ac::core::Processor::info<ac::core::Processor::CPU>()
{
    return "CPU:\n  (no CPU backend in the B200 drop-in)\n";
}
template<> AC_CORE_EXPORT const char* ac::core::Processor::info<ac::core::Processor::OpenCL>()
{
    return "OpenCL:\n  (no OpenCL backend in the B200 drop-in)\n";
}
const char* ac::core::Processor::listInfo()
{
    static const std::string text = std::string(info<Processor::CPU>()) + info<Processor::CUDA>();
    return text.c_str();
}

// exported for bindings / tests: the canonical name a model string resolves to (empty = out of scope)
// planes == one ac::video::Frame::plane array; 0 on success.  Processors without the B200 backend report failure (there is
// no CPU path in this library).
extern "C" AC_CORE_EXPORT int ac_b200_process_frame(ac::core::Processor* processor, const acb200_plane* src, const acb200_plane* dst, int planes,
                                                    int elementType, int shift, double factor)
{
    auto* p = dynamic_cast<B200Processor*>(processor);
    if (!p || !src || !dst) return -1;
    return p->processFrame(src, dst, planes, elementType, shift, factor) ? 0 : -1;
}

bool ac::video::upscale(core::Processor& processor, const Frame& src, Frame& dst, const double factor, const int shift) noexcept
{
    static_assert(sizeof(acb200_plane) == sizeof(src.plane[0]) && offsetof(acb200_plane, data) == 16, "acb200_plane must mirror Frame::plane");
    if (src.planes < 1 || src.planes > 3 || dst.planes != src.planes) return false;
    acb200_plane a[3], b[3];
    for (int i = 0; i < src.planes; i++)
    {
        a[i] = { src.plane[i].width, src.plane[i].height, src.plane[i].channel, src.plane[i].stride, src.plane[i].data };
        b[i] = { dst.plane[i].width, dst.plane[i].height, dst.plane[i].channel, dst.plane[i].stride, dst.plane[i].data };
    }
    return ac_b200_process_frame(&processor, a, b, src.planes, src.elementType, shift, factor) == 0;
}

extern "C" AC_CORE_EXPORT const char* ac_b200_resolve_model(const char* model)
{
    const ModelChoice c = parseModel(model);
    switch (c.family)
    {
    case ACB200_FAMILY_ACNET_LEGACY: return model::ACNetLegacy{ static_cast<model::ACNetLegacy::Variant>(c.variant) }.name();
    case ACB200_FAMILY_ACNET: return model::ACNet<8>{ static_cast<model::ACNet<8>::Variant>(c.variant) }.name();
    case ACB200_FAMILY_ARNET: return model::ARNet<8>{ static_cast<model::ARNet<8>::Variant>(c.variant) }.name();
    case ACB200_FAMILY_ARTCNN:
        return c.features == 32 ? model::ArtCNN<32>{ static_cast<model::ArtCNN<32>::Variant>(c.variant) }.name()
                                : model::ArtCNN<16>{ static_cast<model::ArtCNN<16>::Variant>(c.variant) }.name();
    case ACB200_FAMILY_FSRCNNX:
        return c.features == 16 ? model::FSRCNNX<16>{ static_cast<model::FSRCNNX<16>::Variant>(c.variant) }.name()
                                : model::FSRCNNX<8>{ static_cast<model::FSRCNNX<8>::Variant>(c.variant) }.name();
    default: return "";
    }
}
// feature count F of the model a string resolves to (8, 16 or 32)
extern "C" AC_CORE_EXPORT int ac_b200_model_features(const char* model) { return parseModel(model).features; }

// exported for bindings / tests: the flat weight arrays behind a model string, exactly as handed to the CUDA layer.
// Returns the ACB200_FAMILY_* code, or -1 when the family is out of scope.
extern "C" AC_CORE_EXPORT int ac_b200_model_arrays(const char* model, int* blocks, const float** k, int* nk, const float** b, int* nb, const float** a, int* na)
{
    const ModelChoice c = parseModel(model);
    auto fill = [&](const auto& m) {
        if (blocks) *blocks = m.blocks();
        if (k) *k = m.kernel();
        if (nk) *nk = m.kernelLength();
        if (b) *b = m.bias();
        if (nb) *nb = m.biasLength();
        if (a) *a = m.alphaLength() ? m.alpha() : nullptr;
        if (na) *na = m.alphaLength();
    };
    switch (c.family)
    {
    case ACB200_FAMILY_ACNET_LEGACY: fill(model::ACNetLegacy{ static_cast<model::ACNetLegacy::Variant>(c.variant) }); break;
    case ACB200_FAMILY_ACNET: fill(model::ACNet<8>{ static_cast<model::ACNet<8>::Variant>(c.variant) }); break;
    case ACB200_FAMILY_ARNET: fill(model::ARNet<8>{ static_cast<model::ARNet<8>::Variant>(c.variant) }); break;
    case ACB200_FAMILY_ARTCNN:
        if (c.features == 32) fill(model::ArtCNN<32>{ static_cast<model::ArtCNN<32>::Variant>(c.variant) });
        else fill(model::ArtCNN<16>{ static_cast<model::ArtCNN<16>::Variant>(c.variant) });
        break;
    case ACB200_FAMILY_FSRCNNX:
        if (c.features == 16) fill(model::FSRCNNX<16>{ static_cast<model::FSRCNNX<16>::Variant>(c.variant) });
        else fill(model::FSRCNNX<8>{ static_cast<model::FSRCNNX<8>::Variant>(c.variant) });
        break;
    default: return -1;
    }
    return c.family;
}
