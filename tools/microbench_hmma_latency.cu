// Dependent-issue latency and per-warp throughput of mma.sync m16n8k16 / m16n8k8 (f16 x f16 -> f32) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_hmma_latency tools/microbench_hmma_latency.cu
// For C independent accumulator chains per warp and W warps per SM partition (SMSP), prints cycles per MMA per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template<int CHAINS, bool K8>
__global__ void chain_kernel(float* out, long long* cycles, int iters)
{
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
        {
            if (K8)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template<int CHAINS, bool K8>
void run(int warps_per_smsp, float* out, long long* dcyc)
{
    const int iters = 2000;
    chain_kernel<CHAINS, K8><<<148, warps_per_smsp * 4 * 32>>>(out, dcyc, iters);
    long long cyc = 0;
    cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("%s  chains/warp %d  warps/SMSP %d : %.1f cycles per MMA per warp, %.2f cycles per MMA per SMSP\n", K8 ? "m16n8k8 " : "m16n8k16", CHAINS, warps_per_smsp,
           static_cast<double>(cyc) / (iters * CHAINS), static_cast<double>(cyc) / (iters * CHAINS * warps_per_smsp));
}

int main()
{
    float* out; long long* dcyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&dcyc, sizeof(long long));
    for (int w : { 1, 2, 4, 6, 8 })
    {
        run<1, false>(w, out, dcyc); run<2, false>(w, out, dcyc); run<3, false>(w, out, dcyc); run<6, false>(w, out, dcyc);
        run<1, true>(w, out, dcyc); run<3, true>(w, out, dcyc); run<6, true>(w, out, dcyc);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
