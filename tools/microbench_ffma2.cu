// FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a: 16 independent accumulator chains per thread, N warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ffma2 tools/microbench_ffma2.cu && tools/microbench_ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template<int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float w0, float w1)
{
    float2 s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 a = make_float2(1.0001f, 0.9999f), w = make_float2(w0, w1);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                if (MODE == 0) { s[i].x = fmaf(s[i].x, a.x, w.x); s[i].y = fmaf(s[i].y, a.y, w.y); }
                else s[i] = fma2(s[i], a, w);
            }
    }
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += s[i].x + s[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main()
{
    float* d;
    cudaMalloc(&d, 148 * 4 * 512 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : { 128, 256, 512 })
        for (int mode = 0; mode < 2; mode++)
        {
            for (int rep = 0; rep < 2; rep++)
            {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, threads>>>(d, iters, 0.5f, 0.25f); else k<1><<<148, threads>>>(d, iters, 0.5f, 0.25f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * threads * iters * 64.0 * 2.0;     // scalar FMAs
            printf("%s  %3d threads/SM: %.3f ms, %.1f TFLOP/s fp32, %.1f scalar FMA / clk / SM at 1.965 GHz\n", mode ? "FFMA2" : "FFMA ", threads, ms,
                   2.0 * fma / ms / 1e9, fma / (ms * 1e-3) / 148.0 / 1.965e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
