// `pyac` for the B200 drop-in: the Python surface of the reference's binding (binding/python/src/Binding.cpp:15-191)
// -- pyac.core.Processor(type="auto", device=0, model="acnet-f8b8-hdn"), .process(src, factor=2.0), __call__ (raises
// RuntimeError(error()) when !ok()), .ok/.error/.name/__str__, InfoList/CPU/OpenCL/CUDA, pyac.core.resize,
// ResizeModes / ImreadModes, pyac.specs.ModelList / ProcessorList -- over this library's ac::core.
// imread / imwrite run on the drop-in's own PNG / BMP / PNM / TGA codecs (no JPEG).
#include <cstdint>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include "AC/Core.hpp"
#include "AC/Specs.hpp"

namespace py = pybind11;
using ac::core::Image;
using ac::core::Processor;

namespace
{
    struct ArrayView
    {
        int w, h, c, type, stride;
        void* data;
        bool planar; // ndim == 2
    };

    ArrayView describe(const py::buffer_info& info)
    {
        if (info.ndim != 2 && info.ndim != 3) throw py::buffer_error{ "Incompatible dimension: expected 2 or 3." };
        ArrayView v{};
        v.h = static_cast<int>(info.shape[0]);
        v.w = static_cast<int>(info.shape[1]);
        v.c = info.ndim == 2 ? 1 : static_cast<int>(info.shape[2]);
        v.planar = info.ndim == 2;
        if (info.format == py::format_descriptor<std::uint8_t>::format()) v.type = Image::UInt8;
        else if (info.format == py::format_descriptor<std::uint16_t>::format()) v.type = Image::UInt16;
        else if (info.format == py::format_descriptor<float>::format()) v.type = Image::Float32;
        else if (info.format == "e") v.type = Image::Float16;
        else throw py::buffer_error{ "Incompatible type: expected uint8, uint16, float16 or float32." };
        v.stride = static_cast<int>(info.strides[0]);
        v.data = info.ptr;
        return v;
    }
    py::array allocate(const py::array& like, int h, int w, int c, bool planar)
    {
        return planar ? py::array{ like.dtype(), py::array::ShapeContainer{ h, w } } : py::array{ like.dtype(), py::array::ShapeContainer{ h, w, c } };
    }

    py::array upscale(Processor& self, const py::array& in, const double factor)
    {
        const ArrayView s = describe(in.request());
        const int ow = static_cast<int>(s.w * factor), oh = static_cast<int>(s.h * factor);
        if (ow <= 0 || oh <= 0) throw py::value_error{ "empty result size" };
        // The result lives in an ac::core::Image with tight rows -- page-locked memory from the image pool on a GPU box, so the
        // device-to-host copy runs at the PCIe rate -- and the returned array keeps that image alive (as imread's does).  The GIL is
        // released for the call: Python threads sharing one processor overlap like the C++ callers of tools/benchmark.
        const int es = s.type & 0xff;
        auto* dst = new Image{ ow, oh, s.c, static_cast<Image::ElementType>(s.type), ow * s.c * es };
        py::capsule owner{ dst, [](void* v) { delete static_cast<Image*>(v); } };
        if (dst->empty()) throw std::bad_alloc{};
        {
            Image src{ s.w, s.h, s.c, static_cast<Image::ElementType>(s.type), s.data, s.stride };
            py::gil_scoped_release release;
            self.process(src, *dst, factor);
        }
        if (s.planar) return py::array{ in.dtype(), { oh, ow }, { dst->stride(), es }, dst->data(), owner };
        return py::array{ in.dtype(), { oh, ow, s.c }, { dst->stride(), s.c * es, es }, dst->data(), owner };
    }
}

PYBIND11_MODULE(pyac, m)
{
    m.doc() = "Anime4KCPP CNN upscaling on NVIDIA B200 (drop-in for the reference's pyac).";

    auto core = m.def_submodule("core");

    py::class_<Processor, std::shared_ptr<Processor>>(core, "Processor")
        .def(py::init([](const char* type, const int device, const char* model) { return Processor::create(type, device, model); }),
             py::arg("type") = "auto", py::arg("device") = 0, py::arg("model") = "acnet-f8b8-hdn")
        .def("process", &upscale, py::arg("src"), py::arg("factor") = 2.0)
        .def("__call__", [](Processor& self, const py::array& in, const double factor) {
                py::array out = upscale(self, in, factor);
                if (!self.ok()) throw std::runtime_error{ self.error() };
                return out;
             }, py::arg("src"), py::arg("factor") = 2.0)
        .def("ok", &Processor::ok)
        .def("error", &Processor::error)
        .def("name", &Processor::name)
        .def("__str__", [](Processor& self) { return std::string(self.name()); })
        .def_property_readonly_static("InfoList", [](py::object) {
                return std::make_tuple(std::string(Processor::info<Processor::CPU>()), std::string(Processor::info<Processor::CUDA>()));
             })
        .def_readonly_static("CPU", &Processor::CPU)
        .def_readonly_static("OpenCL", &Processor::OpenCL)
        .def_readonly_static("CUDA", &Processor::CUDA);

    py::enum_<ac::core::ResizeModes>(core, "ResizeModes")
        .value("RESIZE_POINT", ac::core::RESIZE_POINT)
        .value("RESIZE_CATMULL_ROM", ac::core::RESIZE_CATMULL_ROM)
        .value("RESIZE_MITCHELL_NETRAVALI", ac::core::RESIZE_MITCHELL_NETRAVALI)
        .value("RESIZE_BICUBIC_0_60", ac::core::RESIZE_BICUBIC_0_60)
        .value("RESIZE_BICUBIC_0_75", ac::core::RESIZE_BICUBIC_0_75)
        .value("RESIZE_BICUBIC_0_100", ac::core::RESIZE_BICUBIC_0_100)
        .value("RESIZE_BICUBIC_20_50", ac::core::RESIZE_BICUBIC_20_50)
        .value("RESIZE_SOFTCUBIC50", ac::core::RESIZE_SOFTCUBIC50)
        .value("RESIZE_SOFTCUBIC75", ac::core::RESIZE_SOFTCUBIC75)
        .value("RESIZE_SOFTCUBIC100", ac::core::RESIZE_SOFTCUBIC100)
        .value("RESIZE_LANCZOS2", ac::core::RESIZE_LANCZOS2)
        .value("RESIZE_LANCZOS3", ac::core::RESIZE_LANCZOS3)
        .value("RESIZE_LANCZOS4", ac::core::RESIZE_LANCZOS4)
        .value("RESIZE_SPLINE16", ac::core::RESIZE_SPLINE16)
        .value("RESIZE_SPLINE36", ac::core::RESIZE_SPLINE36)
        .value("RESIZE_SPLINE64", ac::core::RESIZE_SPLINE64)
        .value("RESIZE_BILINEAR", ac::core::RESIZE_BILINEAR)
        .export_values();

    py::enum_<ac::core::ImreadModes>(core, "ImreadModes")
        .value("IMREAD_UNCHANGED", ac::core::IMREAD_UNCHANGED)
        .value("IMREAD_GRAYSCALE", ac::core::IMREAD_GRAYSCALE)
        .value("IMREAD_COLOR", ac::core::IMREAD_COLOR)
        .value("IMREAD_RGB", ac::core::IMREAD_RGB)
        .value("IMREAD_RGBA", ac::core::IMREAD_RGBA)
        .export_values();

    // binding/python/src/Binding.cpp:142-157: imread returns an (H,W[,C]) uint8 array that owns its image, imwrite takes such an array
    core.def("imread", [](const char* filename, const ac::core::ImreadModes mode) {
            auto* img = new Image{ ac::core::imread(filename, mode) };
            py::capsule owner{ img, [](void* v) { delete static_cast<Image*>(v); } };
            if (img->empty()) throw std::runtime_error{ std::string{ "pyac.core.imread: cannot read or decode " } + filename };
            if (img->channels() == 1)
                return py::array{ py::dtype::of<std::uint8_t>(), { img->height(), img->width() }, { img->stride(), img->pixelSize() }, img->data(), owner };
            return py::array{ py::dtype::of<std::uint8_t>(), { img->height(), img->width(), img->channels() }, { img->stride(), img->pixelSize(), img->elementSize() }, img->data(), owner };
        }, py::arg("filename"), py::arg("mode") = ac::core::IMREAD_UNCHANGED);
    core.def("imwrite", [](const char* filename, const py::array_t<std::uint8_t>& in) {
            const py::buffer_info src = in.request();
            if (src.ndim != 2 && src.ndim != 3) throw py::buffer_error{ "Incompatible dimension: expected 2 or 3." };
            return ac::core::imwrite(filename, Image{ static_cast<int>(src.shape[1]), static_cast<int>(src.shape[0]), src.ndim == 3 ? static_cast<int>(src.shape[2]) : 1,
                                                       Image::UInt8, src.ptr, static_cast<int>(src.strides[0]) });
        }, py::arg("filename"), py::arg("image"));

    // the reference defaults `mode` to RESIZE_BILINEAR; only RESIZE_CATMULL_ROM upscaling exists on this path
    core.def("resize", [](const py::array& in, const py::object& dsize, const double fx, const double fy, const ac::core::ResizeModes mode) {
            const ArrayView s = describe(in.request());
            int w = static_cast<int>(s.w * fx), h = static_cast<int>(s.h * fy);
            if (!dsize.is_none())
            {
                const py::tuple t = dsize.cast<py::tuple>();
                if (t.size() != 2) throw py::value_error{ "dsize should be (width, height)" };
                w = t[0].cast<int>();
                h = t[1].cast<int>();
            }
            if (w <= 0 || h <= 0) throw py::value_error{ "empty destination size" };
            if (mode != ac::core::RESIZE_CATMULL_ROM)
                throw py::value_error{ "pyac.core.resize: only RESIZE_CATMULL_ROM is implemented on the B200 path (pass mode=pyac.core.RESIZE_CATMULL_ROM)" };
            py::array out = allocate(in, h, w, s.c, s.c == 1);
            const py::buffer_info oinfo = out.request();
            Image src{ s.w, s.h, s.c, s.type, s.data, s.stride };
            Image dst{ w, h, s.c, s.type, oinfo.ptr, static_cast<int>(oinfo.strides[0]) };
            ac::core::resize(src, dst, 0.0, 0.0, mode);
            if (ac::core::lastImageOpStatus() != 0) throw std::runtime_error{ "pyac.core.resize failed on the GPU path (see stderr)" };
            return out;
        }, py::arg("src"), py::arg("dsize"), py::arg("fx") = 0.0, py::arg("fy") = 0.0, py::arg("mode") = ac::core::RESIZE_BILINEAR);

    auto specs = m.def_submodule("specs");
    {
        py::tuple models{ std::size(ac::specs::ModelList) };
        for (std::size_t i = 0; i < std::size(ac::specs::ModelList); i++)
        {
            const auto& e = ac::specs::ModelList[i];
            py::dict d{};
            d["name"] = e.name;
            d["description"] = e.description;
            d["parameter_count"] = e.parameterCount;
            d["version"] = e.version;
            d["author"] = e.author;
            d["homepage"] = e.homepage;
            models[i] = d;
        }
        specs.attr("ModelList") = models;
        py::tuple processors{ std::size(ac::specs::ProcessorList) };
        for (std::size_t i = 0; i < std::size(ac::specs::ProcessorList); i++)
        {
            py::dict d{};
            d["name"] = ac::specs::ProcessorList[i].name;
            d["description"] = ac::specs::ProcessorList[i].description;
            processors[i] = d;
        }
        specs.attr("ProcessorList") = processors;
    }
}
