// Measures the per-SM issue rates that decide the luma-network engine on this B200:
// fp32 FFMA, legacy mma.sync fp16 (m16n8k16 / m16n8k8, fp32 accumulate), mma.sync tf32 m16n8k8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench.cu -o gpurun_out/microbench && gpurun_out/microbench
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__global__ void ffma_kernel(float* out, int iters, float a, float b)
{
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = fmaf(acc[i], a, b);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int CHAINS>
__global__ void hmma_k16_kernel(float* out, int iters)
{
    unsigned a[4] = { 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u }, b[2] = { 0x38003800u, 0x38003800u };
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int CHAINS>
__global__ void hmma_k8_kernel(float* out, int iters)
{
    unsigned a[2] = { 0x3c003c00u, 0x3c003c00u }, b[1] = { 0x38003800u };
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int CHAINS>
__global__ void tf32_k8_kernel(float* out, int iters)
{
    unsigned a[4] = { 0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u }, b[2] = { 0x3f000000u, 0x3f000000u };
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<class F>
float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, khz / 1000);
    const int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2)
    {
        const int threads = warps * 32, blocks = sms;
        float ms = time_ms([&] { ffma_kernel<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        double fma = double(blocks) * threads * iters * 32;
        printf("FFMA            warps/SM %2d : %8.1f TFLOP/s  (%.1f FMA/clk/SM @max clock)\n", warps, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3));
        ms = time_ms([&] { hmma_k16_kernel<8><<<blocks, threads>>>(out, iters); });
        double mac = double(blocks) * warps * iters * 8 * (16 * 8 * 16);
        printf("HMMA m16n8k16   warps/SM %2d : %8.1f TFLOP/s  (%.0f MAC/clk/SM)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / (khz * 1e3));
        ms = time_ms([&] { hmma_k8_kernel<8><<<blocks, threads>>>(out, iters); });
        mac = double(blocks) * warps * iters * 8 * (16 * 8 * 8);
        printf("HMMA m16n8k8    warps/SM %2d : %8.1f TFLOP/s  (%.0f MAC/clk/SM)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / (khz * 1e3));
        ms = time_ms([&] { tf32_k8_kernel<8><<<blocks, threads>>>(out, iters); });
        printf("TF32 m16n8k8    warps/SM %2d : %8.1f TFLOP/s  (%.0f MAC/clk/SM)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / (khz * 1e3));
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
