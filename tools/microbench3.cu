// Does ordinary ALU work (the epilogue's FADD / F2FP / address arithmetic) issue in the shadow of legacy mma.sync on sm_100,
// or do the two add up?  Per loop iteration a warp issues 14 HMMA (three accumulator chains) and / or NALU independent FFMAs
// (eight chains).  Modes: HMMA only, ALU only, both in every warp, and "split": even warps run 2x the HMMAs, odd warps 2x the
// ALU work (same totals per SM partition, the two kinds never share a warp).
#include <cstdio>
#include <cuda_runtime.h>
#ifndef NALU
#define NALU 80
#endif
__device__ __forceinline__ void hmma14(float (&c)[3][4], const unsigned (&a)[4], const unsigned (&b)[2])
{
#pragma unroll
    for (int j = 0; j < 14; j++)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j % 3][0]), "+f"(c[j % 3][1]), "+f"(c[j % 3][2]), "+f"(c[j % 3][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// OP 0: fma.rn.f32 (FMA pipe)   1: xor.b32 (ALU pipe)   2: add.f32   3: mad.lo.s32 (IMAD)   4: cvt.rn.f16x2.f32 (F2FP)
template<int OP>
__device__ __forceinline__ void alu(unsigned (&f)[8], unsigned m)
{
#pragma unroll
    for (int j = 0; j < NALU; j++)
    {
        if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(f[j & 7]) : "r"(m));
        else if (OP == 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(f[j & 7]) : "r"(f[(j + 3) & 7]));
        else if (OP == 2) asm volatile("add.f32 %0, %0, %1;" : "+r"(f[j & 7]) : "r"(m));
        else if (OP == 3) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(f[j & 7]) : "r"(m));
        else asm volatile("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %0, %1; mov.b32 %0, t; }" : "+r"(f[j & 7]) : "r"(m));
    }
}
template<int OP>
__global__ void k_mix(float* out, int iters, int mode)
{
    unsigned a[4] = { 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u }, b[2] = { 0x38003800u, 0x38003800u };
    float c[3][4] = {};
    unsigned f[8] = { 0x3f800000u, 0x3f810000u, 0x3f820000u, 0x3f830000u, 0x3f840000u, 0x3f850000u, 0x3f860000u, 0x3f870000u };
    const unsigned m = 0x3f800000u + threadIdx.x;
    const bool even = ((threadIdx.x >> 7) & 1) == 0;     // warps 0-3, 8-11, ...: one warp per SM partition in each group
    for (int it = 0; it < iters; it++)
    {
        if (mode == 1) hmma14(c, a, b);
        else if (mode == 2) alu<OP>(f, m);
        else if (mode == 3) { hmma14(c, a, b); alu<OP>(f, m); }
        else if (even) { hmma14(c, a, b); hmma14(c, a, b); }
        else { alu<OP>(f, m); alu<OP>(f, m); }
    }
    float s = 0;
    for (int j = 0; j < 8; j++) s += __uint_as_float(f[j] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + s;
}
template<int OP>
void run(const char* opname, float* out);
int main()
{
    float* out; cudaMalloc(&out, 4 * 148 * 1024);
    run<0>("fma.rn.f32", out); run<1>("xor.b32", out); run<2>("add.f32", out); run<3>("mad.lo.s32", out); run<4>("cvt.rn.f16x2.f32", out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
template<int OP>
void run(const char* opname, float* out)
{
    printf("---- other instruction: %s\n", opname);
    const int iters = 2000;
    for (int warps = 8; warps <= 16; warps *= 2)
        for (int mode = 1; mode <= 4; mode++)
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_mix<OP><<<148, warps * 32>>>(out, iters, mode);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_mix<OP><<<148, warps * 32>>>(out, iters, mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("warps/SM %2d  %-22s : %.3f ms -> %.1f cycles per (14 HMMA + %d other) per warp of an SM partition at 1.965 GHz\n", warps,
                   mode == 1 ? "HMMA only" : mode == 2 ? "ALU only" : mode == 3 ? "both, same warp" : "split by warp parity", ms,
                   ms * 1e-3 * 1.965e9 / (double(iters) * warps / 4.0), NALU);
        }
}
