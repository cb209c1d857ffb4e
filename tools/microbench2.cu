// Do ldmatrix (shared-memory pipe) and legacy mma.sync (tensor pipe) overlap on sm_100?  Times HMMA-only, LDSM-only and a
// mixed kernel with the same instruction counts.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mix(float* out, int iters, int do_mma, int do_ldsm)
{
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(smem)[i] = 0x3c003c00u;
    __syncthreads();
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(smem)) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 2048;
    unsigned a[4] = { 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u }, b[2] = { 0x38003800u, 0x38003800u };
    float c[3][4] = {};
    unsigned acc = 0;
    for (int it = 0; it < iters; it++)
    {
        if (do_ldsm)
        {
#pragma unroll
            for (int j = 0; j < 10; j++)
            {
                unsigned r0, r1, r2, r3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + ((it + j) & 3) * 512));
                acc ^= r0 ^ r1 ^ r2 ^ r3;
            }
        }
        if (do_mma)
        {
#pragma unroll
            for (int j = 0; j < 15; j++)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j % 3][0]), "+f"(c[j % 3][1]), "+f"(c[j % 3][2]), "+f"(c[j % 3][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + __uint_as_float(acc & 1);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; cudaMalloc(&out, 4 * 148 * 1024);
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    const int iters = 2000;
    for (int warps = 8; warps <= 32; warps *= 2)
        for (int mode = 1; mode <= 3; mode++)
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_mix<<<148, warps * 32, 48 * 1024>>>(out, iters, mode & 1, mode >> 1);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_mix<<<148, warps * 32, 48 * 1024>>>(out, iters, mode & 1, mode >> 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("warps %2d  %-10s : %.3f ms  -> %.1f cycles per (10 LDSM.x4 + 15 HMMA) per SM at 1.965 GHz\n", warps,
                   mode == 1 ? "HMMA only" : mode == 2 ? "LDSM only" : "both", ms, ms * 1e-3 * 1.965e9 / (double(iters) * warps));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
