// Host-descriptor consistency checks in the shape of the reference's tests/core/src/ModelTest.cpp:5-60 (offsets and lengths of
// every layer add up; first offsets are zero; arrays are bound), run over every variant of every family, plus the per-layer
// structure each family's header documents.  Built and run by tests/test_host_cpu.py against libac_b200.so.
#include <cstdio>
#include <cstring>

#include "AC/Core.hpp"

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL %s:%d: %s [%s]\n", __FILE__, __LINE__, #cond, what); failures++; } } while (0)

template<typename Model>
static void checkBase(const Model& m, const char* what)
{
    EXPECT(m.blocks() > 0);
    EXPECT(m.kernels() > 0 && m.biases() > 0);
    EXPECT(m.kernel(0) != nullptr && m.bias(0) != nullptr);
    EXPECT(m.kernelOffset(0) == 0 && m.biasOffset(0) == 0);
    const int lk = m.kernels() - 1, lb = m.biases() - 1;
    EXPECT(m.kernelOffset(lk) + m.kernelLength(lk) == m.kernelLength());
    EXPECT(m.biasOffset(lb) + m.biasLength(lb) == m.biasLength());
    int ks = 0, bs = 0;
    for (int i = 0; i < m.kernels(); i++) { EXPECT(m.kernelOffset(i) == ks); ks += m.kernelLength(i); }
    for (int i = 0; i < m.biases(); i++) { EXPECT(m.biasOffset(i) == bs); bs += m.biasLength(i); }
    EXPECT(ks == m.kernelLength() && bs == m.biasLength());
    EXPECT(m.kernelSize() == m.kernelLength() * sizeof(float) && m.biasSize() == m.biasLength() * sizeof(float));
    EXPECT(m.kernel(1) == m.kernel(0) + m.kernelLength(0) && m.bias(1) == m.bias(0) + m.biasLength(0));
    EXPECT(std::strlen(m.name()) > 0);
}

int main()
{
    using namespace ac::core::model;
    const char* what = "";
    for (int v = 0; v < 5; v++)
    {
        ACNetLegacy m{ static_cast<ACNetLegacy::Variant>(v) };
        what = m.name();
        checkBase(m, what);
        EXPECT(m.alphas() == 0 && m.kernels() == 10 && m.biases() == 9 && m.kernelLength() == 4712 && m.kernelLength(9) == 32);
    }
    for (int v = 0; v < 12; v++)
    {
        ACNet<8> m{ static_cast<ACNet<8>::Variant>(v) };
        what = m.name();
        checkBase(m, what);
        EXPECT(m.alphas() == m.blocks() + 1 && m.alpha(0) != nullptr && m.alphaOffset(0) == 0 && m.kernelLength(m.kernels() - 1) == 288);
        EXPECT(m.blocks() == (v < 4 ? 4 : v < 8 ? 8 : 18));
    }
    for (int v = 0; v < 16; v++)
    {
        ARNet<8> m{ static_cast<ARNet<8>::Variant>(v) };
        what = m.name();
        checkBase(m, what);
        EXPECT(m.alphas() == m.blocks() + 1 && m.alpha(0) != nullptr && m.kernels() == 2 * m.blocks() + 3);
        EXPECT(m.kernelLength(2 * m.blocks() + 1) == 64 && m.alphaLength(1) == 8 && m.alphaLength(2) == 0 && m.alphaOffset(3) == 8);
    }
    for (int v = 0; v < 3; v++)
    {
        ArtCNN<16> a{ static_cast<ArtCNN<16>::Variant>(v) };
        what = a.name();
        checkBase(a, what);
        EXPECT(a.blocks() == 4 && a.kernels() == 7 && a.alphas() == 0 && a.kernelLength() == 12240 && a.kernelLength(0) == 144 && a.kernelLength(6) == 576);
        ArtCNN<32> b{ static_cast<ArtCNN<32>::Variant>(v) };
        what = b.name();
        checkBase(b, what);
        EXPECT(b.kernelLength() == 47520 && b.biasLength() == 196 && b.kernelLength(1) == 9216);
    }
    for (int v = 0; v < 2; v++)
    {
        FSRCNNX<8> a{ static_cast<FSRCNNX<8>::Variant>(v) };
        what = a.name();
        checkBase(a, what);
        EXPECT(a.blocks() == 4 && a.kernels() == 7 && a.alphas() == 5 && a.kernelLength() == 2856 && a.kernelLength(0) == 200 && a.kernelLength(5) == 64);
        EXPECT(a.alphaOffset(1) == 0 && a.alphaOffset(2) == 8 && a.alphaLength(0) == 0 && a.alphaLength(5) == 8 && a.alphaLength() == 40);
        FSRCNNX<16> b{ static_cast<FSRCNNX<16>::Variant>(v) };
        what = b.name();
        checkBase(b, what);
        EXPECT(b.kernelLength() == 10448 && b.alphaLength() == 80 && b.kernelOffset(6) == 400 + 2304 * 4 + 256);
    }
    // the front door resolves every family (reference tests/core/src/ProcessorTest.cpp:105-111 checks listInfo the same way)
    what = "listInfo";
    EXPECT(ac::core::Processor::listInfo() != nullptr && std::strlen(ac::core::Processor::listInfo()) > 0);
    std::printf(failures ? "%d checks failed\n" : "model checks ok\n", failures);
    return failures ? 1 : 0;
}
