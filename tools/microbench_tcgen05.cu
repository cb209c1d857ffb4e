// tcgen05 feasibility probe for the luma network's implicit GEMM on B200 (sm_100a):
//   D[128 px x N] (+)= A[128 px x 16] * B[16 x N], fp16 operands from shared memory (SS), fp32 accumulator in TMEM.
// A is the activation plane [pixel][8 ch fp16] (16 B / pixel): with the no-swizzle K-major canonical layout one
// 8x16-byte core matrix is 8 consecutive pixels, the two K chunks of a k-step are two 3x3 taps (LBO = tap distance) and
// a tap shift is just a different descriptor start address -- no im2col, no data movement.
// Checks one MMA against a host reference, then measures issue-to-completion cycles per MMA for N = 8..32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench_tcgen05.cu -o tools/microbench_tcgen05.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// bounded spin: a bad descriptor must not hang the box
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    uint32_t ok = 0;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok && clock64() - t0 < 2000000000LL);
    return ok != 0;
}

constexpr int A_PIXELS = 4096;          // 64 KB activation plane
constexpr int SHIFT = 57;               // second tap = pixel + 57 (one row of a 56-wide frame + 1)

template<int N, int ACCS>
__global__ void __launch_bounds__(128, 1) probe(const __half* __restrict__ a_init, const __half* __restrict__ b_init, float* __restrict__ d_out,
                                                long long* __restrict__ cycles, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);                               // [A_PIXELS][8]
    __half* sB = reinterpret_cast<__half*>(smem + A_PIXELS * 16);               // [2 k-chunks][N][8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_PIXELS * 16 + 2 * N * 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < A_PIXELS * 8; i += 128) sA[i] = a_init[i];
    for (int i = tid; i < 2 * N * 8; i += 128) sB[i] = b_init[i];
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // generic-proxy smem writes -> async proxy (MMA reads)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    const uint64_t db = make_desc(b_base, N * 16, 128);                         // K chunks N*16 B apart, 8-row groups 128 B apart
    uint32_t parity = 0;
    long long t0 = 0, t1 = 0;
    if (tid == 0)
    {
        // ---- one MMA for the numerical check -------------------------------------------------------------------------
        mma_f16_ss(tmem, make_desc(a_base + 5 * 16, SHIFT * 16, 128), db, idesc, 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(smem_u32(bar), parity); parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t v[32];
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        if (N <= 16)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                           "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        else
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                           "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                           "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0)
            for (int j = 0; j < (N < 32 ? N : 32); j++) d_out[tid * N + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- throughput: `reps` groups of 10 MMAs (5 tap pairs x {hi, lo} planes) per 128-pixel tile, like one conv layer ------
    if (tid == 0)
    {
        t0 = clock64();
        for (int r = 0; r < reps; r++)
        {
            const uint32_t tile = a_base + ((r * 128) % (A_PIXELS - 512)) * 16;
#pragma unroll
            for (int s = 0; s < 10; s++)
                mma_f16_ss(tmem + (s % ACCS) * N, make_desc(tile + (s % 5) * 64 * 16 + (s / 5) * 16, SHIFT * 16, 128), db, idesc, s >= ACCS);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(smem_u32(bar), parity); parity ^= 1;
    if (tid == 0) { t1 = clock64(); cycles[blockIdx.x] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

// Pipeline probe: per "tile" 3 accumulating MMAs into one of 8 TMEM slots + one tcgen05.commit to that slot's mbarrier,
// optionally waiting (try_wait) on the barrier of the tile issued 8 tiles earlier.  LBO selects the K-chunk distance.
__global__ void __launch_bounds__(128, 1) pipe_probe(long long* __restrict__ cycles, int tiles, uint32_t lbo_bytes, int do_wait, int mmas_per_tile)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0)
    {
        for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(48 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 150 * 1024);
    if (tid == 0)
    {
        const long long t0 = clock64();
        for (int t = 0; t < tiles; t++)
        {
            const uint32_t slot = t & 7, use = t >> 3;
            if (do_wait && use > 0) mbar_wait(smem_u32(bars + slot), (use - 1) & 1);
            if (do_wait) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = a_base + ((t * 126) % 2000) * 16 + 57 * 16;
            for (int dy = 0; dy < mmas_per_tile; dy++)
                mma_f16_ss(tmem + slot * 64, make_desc(base + (dy - 1) * 56 * 16, lbo_bytes, 128), make_desc(b_base + dy * 1536, 768, 128), idesc, dy > 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bars + slot)) : "memory");
        }
        mbar_wait(smem_u32(bars + ((tiles - 1) & 7)), ((tiles - 1) >> 3) & 1);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}
void run_pipe(uint32_t lbo, int do_wait, int mmas)
{
    long long* dc; cudaMalloc(&dc, 148 * 8);
    const size_t smem = 200 * 1024 + 128;
    cudaFuncSetAttribute(pipe_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 4000;
    pipe_probe<<<148, 128, smem>>>(dc, tiles, lbo, do_wait, mmas);
    cudaError_t err = cudaDeviceSynchronize();
    std::vector<long long> hc(148); cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : hc) avg += double(c); avg /= 148;
    printf("pipe: LBO %6u B, %d MMA/tile, wait %d: %s, %.1f cycles per tile\n", lbo, mmas, do_wait, cudaGetErrorString(err), avg / tiles);
    cudaFree(dc);
}

template<int N, int ACCS>
void run(const std::vector<__half>& ha, int reps)
{
    std::vector<__half> hb(2 * N * 8);
    for (int kc = 0; kc < 2; kc++) for (int n = 0; n < N; n++) for (int k = 0; k < 8; k++) hb[(kc * N + n) * 8 + k] = __float2half(float((n + 2 * (kc * 8 + k)) % 5 - 2));
    __half *da, *db; float* dd; long long* dc;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * N * 4); cudaMalloc(&dc, 148 * 8);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = A_PIXELS * 16 + 2 * N * 16 + 64;
    cudaFuncSetAttribute(probe<N, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<N, ACCS><<<148, 128, smem>>>(da, db, dd, dc, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<N, ACCS><<<148, 128, smem>>>(da, db, dd, dc, reps);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> hd(128 * N); std::vector<long long> hc(148);
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < (N < 32 ? N : 32); n++)
    {
        float ref = 0;
        for (int k = 0; k < 16; k++)
        {
            const int px = 5 + m + (k >= 8 ? SHIFT : 0);
            ref += __half2float(ha[px * 8 + (k & 7)]) * __half2float(hb[((k >> 3) * N + n) * 8 + (k & 7)]);
        }
        if (ref != hd[m * N + n]) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, hd[m * N + n], ref); bad++; }
    }
    double avg = 0; for (auto c : hc) avg += double(c); avg /= 148;
    printf("N=%3d accs=%d: %s, check %s (%d bad); %d x 10 MMAs: %.1f cycles per MMA (M=128,K=16), %.1f cycles per 128-px layer tile; kernel %.3f ms\n", N, ACCS,
           cudaGetErrorString(err), bad ? "FAILED" : "ok", bad, reps, avg / (reps * 10.0), avg / reps, ms);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
}

int main()
{
    std::vector<__half> ha(A_PIXELS * 8);
    for (int p = 0; p < A_PIXELS; p++) for (int c = 0; c < 8; c++) ha[p * 8 + c] = __float2half(float((p * 3 + c) % 7));
    run_pipe(912, 0, 3); run_pipe(50176, 0, 3); run_pipe(50176 + 16, 0, 3); run_pipe(50176 + 64, 0, 3); run_pipe(50176, 1, 3); run_pipe(50176, 0, 1); run_pipe(50176, 0, 6);
    run<16, 1>(ha, 2000);
    run<16, 2>(ha, 2000);
    run<16, 5>(ha, 2000);
    run<48, 1>(ha, 2000);
    run<48, 2>(ha, 2000);
    run<144, 1>(ha, 2000);
    run<144, 2>(ha, 2000);
    run<256, 1>(ha, 2000);
    return 0;
}
