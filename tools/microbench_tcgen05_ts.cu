// tcgen05 probe, round 2: the A operand in TENSOR MEMORY (TS mode), the .ashift qualifier / tcgen05.shift, and the
// TMEM load / store rates -- the facts the TMEM-resident luma engine (acb200_tm.cuh) is designed around.
//
//   semantics (one CTA, checked on the host):
//     1. D = A_tmem * B_smem with A written by tcgen05.st.32x32b.x8 (lane = row, column c = K elements 2c | 2c+1)
//     2. tcgen05.mma ... .ashift : is A shifted before or after the multiply, in which direction, over which lanes
//     3. tcgen05.shift.down      : the same for the stand-alone shift
//   rates (148 CTAs, one issuing thread, clock64 around issue .. commit arrival):
//     4. cycles per TS-mode MMA for N = 8 .. 96, with and without .ashift, A operands rotating over TMEM columns
//     5. cost of a tcgen05.commit every g MMAs
//     6. tcgen05.ld / tcgen05.st bytes per cycle and SM with 4 / 8 / 16 warps
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench_tcgen05_ts.cu -o tools/microbench_tcgen05_ts.bin
#include <algorithm>
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
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts_ashift(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.ashift [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tm_shift(uint32_t tmem_a) { asm volatile("tcgen05.shift.cta_group::1.down [%0];" :: "r"(tmem_a) : "memory"); }
__device__ __forceinline__ void tm_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    uint32_t ok = 0;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok && clock64() - t0 < 400000000LL);
    return ok != 0;
}
__device__ __forceinline__ void tm_ld8(uint32_t (&v)[8], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tm_ld16(uint32_t (&v)[16], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                 :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                    "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
#define TM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TM_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TM_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

__host__ __device__ inline int a_val(int row, int k) { return ((row * 3 + k * 5) % 7) - 3; }
__host__ __device__ inline int b_val(int k, int n) { return ((n + 2 * k) % 5) - 2; }

#ifndef N_CHK_VALUE
#define N_CHK_VALUE 48
#endif
constexpr int N_CHK = N_CHK_VALUE;
constexpr int A_COL = 256;      // A operand columns used by the checks
// out: [4][128][N_CHK] floats (D0 plain, D1 ashift MMA, D2 plain after ashift, D3 plain after tcgen05.shift), then [2][128][8] u32 (A after 2, after 3)
__global__ void __launch_bounds__(128, 1) semantics(float* __restrict__ d_out, uint32_t* __restrict__ a_out, int* __restrict__ status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sB = reinterpret_cast<__half*>(smem);                   // [2 k-chunks][N][8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * N_CHK * 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * N_CHK * 8; i += 128)
    {
        const int kc = i / (N_CHK * 8), n = (i / 8) % N_CHK, kk = i % 8;
        sB[i] = __float2half(static_cast<float>(b_val(kc * 8 + kk, n)));
    }
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N_CHK >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t db = make_desc(smem_u32(sB), N_CHK * 16, 128);
    uint32_t parity = 0;
    int ok = 1;
    auto write_a = [&]() {
        uint32_t v[8];
#pragma unroll
        for (int c = 0; c < 8; c++)
        {
            const __half2 h = __floats2half2_rn(static_cast<float>(a_val(tid, 2 * c)), static_cast<float>(a_val(tid, 2 * c + 1)));
            v[c] = *reinterpret_cast<const uint32_t*>(&h);
        }
        tm_st8(tmem + lane_base + A_COL, v);
        TM_WAIT_ST();
        TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    };
    auto read_d = [&](int which, int col) {
        for (int c0 = 0; c0 < N_CHK; c0 += 8)
        {
            uint32_t v[8];
            tm_ld8(v, tmem + lane_base + col + c0);
            TM_WAIT_LD();
            for (int j = 0; j < 8; j++) d_out[(which * 128 + tid) * N_CHK + c0 + j] = __uint_as_float(v[j]);
        }
    };
    auto read_a = [&](int which) {
        uint32_t v[8];
        tm_ld8(v, tmem + lane_base + A_COL);
        TM_WAIT_LD();
        for (int j = 0; j < 8; j++) a_out[(which * 128 + tid) * 8 + j] = v[j];
    };
    // 1. plain TS MMA; D at an odd multiple of 16 columns
    write_a();
    if (tid == 0) { mma_ts(tmem + 80, tmem + A_COL, db, idesc, 0); tm_commit(smem_u32(bar)); }
    ok &= mbar_wait(smem_u32(bar), parity); parity ^= 1;
    TM_FENCE_AFTER();
    read_d(0, 80);
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    // 2. ashift MMA, then a plain MMA on the same A
    if (tid == 0) { mma_ts_ashift(tmem + 0, tmem + A_COL, db, idesc, 0); mma_ts(tmem + 64, tmem + A_COL, db, idesc, 0); tm_commit(smem_u32(bar)); }
    ok &= mbar_wait(smem_u32(bar), parity); parity ^= 1;
    TM_FENCE_AFTER();
    read_d(1, 0); read_d(2, 64); read_a(0);
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    // 3. stand-alone shift, then a plain MMA
    write_a();
    if (tid == 0) { tm_shift(tmem + A_COL); mma_ts(tmem + 128, tmem + A_COL, db, idesc, 0); tm_commit(smem_u32(bar)); }
    ok &= mbar_wait(smem_u32(bar), parity); parity ^= 1;
    TM_FENCE_AFTER();
    read_d(3, 128); read_a(1);
    if (tid == 0) *status = ok;
    TM_FENCE_BEFORE(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

// Tight issue loop (everything but the loop counter is loop-invariant or an immediate): per iteration 12 MMAs.
//   PATTERN 0: 12 plain TS MMAs, same A, same D (accumulating)
//   PATTERN 1: 12 plain TS MMAs, A and D both rotate per MMA over 4 operands / 4 accumulators
//   PATTERN 2: the engine's row step, 4 rows per iteration: per row {mma.ashift, mma.ashift, mma} on that row's A into that row's D
//   PATTERN 3: like 2 but the three alignments of FOUR rows interleaved (row 0..3 first alignment, then second, then third)
//   PATTERN 4: 12 SS MMAs (A from shared memory, K-major no-swizzle), same D: the round-1 engine's instruction, for reference
//   PATTERN 5: like 1 with accumulate = 0
// COMMITS: tcgen05.commit every 12 MMAs (1) or only at the end (0)
template<int N, int PATTERN, int COMMITS>
__global__ void __launch_bounds__(128, 1) rate_mma(long long* __restrict__ cycles, int iters)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0)
    {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    const uint32_t tmem = *tmem_slot;
    {
        uint32_t v[16];
        for (int j = 0; j < 16; j++) v[j] = 0x3c003c00u;
        for (int c = 0; c < 512; c += 16) tm_st16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        TM_WAIT_ST();
    }
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t b_base = smem_u32(smem);
    if (warp == 0 && elect_one())
    {
        uint64_t db[3];
        for (int j = 0; j < 3; j++) db[j] = make_desc(b_base + j * N * 32, N * 16, 128);
        const uint64_t da = make_desc(b_base + 16 * 1024, 8 * 1024, 128);      // SS reference: A tile of 128 rows x 2 K chunks
        const uint32_t a0 = tmem + 384, d0 = tmem;          // A operands: columns 384.. (8 each); D: columns 0.. (N <= 96 each, 4 of them)
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++)
        {
#pragma unroll
            for (int j = 0; j < 12; j++)
            {
                if (PATTERN == 0) mma_ts(d0, a0, db[j % 3], idesc, 1);
                else if (PATTERN == 1) mma_ts(d0 + (j % 4) * 96, a0 + (j % 4) * 8, db[j % 3], idesc, 1);
                else if (PATTERN == 5) mma_ts(d0 + (j % 4) * 96, a0 + (j % 4) * 8, db[j % 3], idesc, 0);
                else if (PATTERN == 2)
                {
                    const int row = j / 3, al = j % 3;
                    if (al < 2) mma_ts_ashift(d0 + row * 96, a0 + row * 8, db[al], idesc, 1);
                    else mma_ts(d0 + row * 96, a0 + row * 8, db[al], idesc, 1);
                }
                else if (PATTERN == 3)
                {
                    const int row = j % 4, al = j / 4;
                    if (al < 2) mma_ts_ashift(d0 + row * 96, a0 + row * 8, db[al], idesc, 1);
                    else mma_ts(d0 + row * 96, a0 + row * 8, db[al], idesc, 1);
                }
                else
                {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(d0), "l"(da), "l"(db[j % 3]), "r"(idesc), "r"(1u) : "memory");
                }
            }
            if (COMMITS) tm_commit(smem_u32(bars + 1 + (it & 1)));
        }
        tm_commit(smem_u32(bars));
        mbar_wait(smem_u32(bars), 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    TM_FENCE_BEFORE(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}
// What the issuing thread pays per "row" of the engine besides the MMAs: VARIANT bit 0 = tcgen05.commit per row, bit 1 = one mbarrier
// try_wait on an ALREADY COMPLETE barrier per row, bit 2 = tcgen05.fence::after_thread_sync per row, bit 3 = a second try_wait.
template<int VARIANT>
__global__ void __launch_bounds__(128, 1) issue_cost(long long* __restrict__ cycles, int rows)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0)
    {
        for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bars + 4)) : "memory");     // bars[4], bars[5]: phase 0 complete
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bars + 5)) : "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    const uint32_t tmem = *tmem_slot;
    {
        uint32_t v[16];
        for (int j = 0; j < 16; j++) v[j] = 0x3c003c00u;
        for (int c = 0; c < 512; c += 16) tm_st16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        TM_WAIT_ST();
    }
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(48 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t b_base = smem_u32(smem);
    if (warp == 0 && elect_one())
    {
        uint64_t db[3];
        for (int j = 0; j < 3; j++) db[j] = make_desc(b_base + j * 1536, 768, 128);
        const long long t0 = clock64();
        for (int r = 0; r < rows; r++)
        {
            if (VARIANT & 2) mbar_wait(smem_u32(bars + 4), 0);
            if (VARIANT & 8) mbar_wait(smem_u32(bars + 5), 0);
            if (VARIANT & 16)
            {
                // plain shared-memory flag poll (what a producer would publish with st.shared after its fences)
                uint32_t f;
                do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(f) : "r"(smem_u32(bars + 6)) : "memory"); } while (f == 0xdeadbeefu);
            }
            if (VARIANT & 32)
            {
                uint32_t ok;
                do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bars + 4)), "r"(0) : "memory"); } while (!ok);
            }
            if (VARIANT & 64)
            {
                // software-pipelined flag: the load for the NEXT row is issued before this row's MMAs, consumed a row later
                static_assert(true, "");
            }
            if (VARIANT & 4) TM_FENCE_AFTER();
            const uint32_t a = tmem + 8 * (r % 48), d = tmem + 384 + 16 * (r & 3);
            mma_ts_ashift(d, a, db[0], idesc, 1);
            mma_ts_ashift(d, a, db[1], idesc, 1);
            mma_ts(d, a, db[2], idesc, 1);
            if (VARIANT & 1) tm_commit(smem_u32(bars + 1 + (r & 1)));
        }
        tm_commit(smem_u32(bars));
        mbar_wait(smem_u32(bars), 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    TM_FENCE_BEFORE(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}
template<int VARIANT>
void run_issue(long long* dc)
{
    const int rows = 4000;
    cudaFuncSetAttribute(issue_cost<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 128);
    issue_cost<VARIANT><<<148, 128, 48 * 1024 + 128>>>(dc, rows);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> hc(148);
    cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double a = 0; for (auto c : hc) a += double(c);
    printf("issuer row = {ashift, ashift, plain} N=48%s%s%s%s%s%s: %s, %.1f cycles per row\n", (VARIANT & 16) ? " + ld.volatile.shared poll" : "", (VARIANT & 32) ? " + test_wait(complete)" : "", (VARIANT & 1) ? " + commit" : "", (VARIANT & 2) ? " + try_wait(complete)" : "",
           (VARIANT & 8) ? " + try_wait(complete)" : "", (VARIANT & 4) ? " + fence::after_thread_sync" : "", cudaGetErrorString(e), a / 148 / rows);
}

// W issuer warps at once, each with its own A operands and accumulators (N = 24, per row {plain, ashift} x 3 as the engine issues them):
// do MMAs from different warps stream through the tensor pipe as well as MMAs from one warp?
__global__ void __launch_bounds__(640, 1) multi_issuer(long long* __restrict__ cycles, int rows, int nwarps, int commit_every, int d_slide = 0, int a_rotate = 8, int side_traffic = 0)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 48 * 1024 + 200);
    if (tid == 0) *stop = 0;
    if (tid == 0)
    {
        for (int i = 0; i < 16; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    const uint32_t tmem = *tmem_slot;
    {
        uint32_t v[16];
        for (int j = 0; j < 16; j++) v[j] = 0x3c003c00u;
        if (warp < 4) for (int c = 0; c < 512; c += 16) tm_st16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        TM_WAIT_ST();
    }
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    if (warp >= 4)
    {
        // side traffic of the engine's epilogue warps: tcgen05.ld / tcgen05.st on columns no MMA touches, `side_traffic` = idle cycles between items
        if (side_traffic > 0)
        {
            uint32_t v[8];
            const uint32_t col = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 320 + 8 * ((warp >> 2) & 3);
            long long n = 0;
            while (*stop < nwarps)
            {
                tm_ld8(v, col); TM_WAIT_LD();
                for (int j = 0; j < 8; j++) v[j] += 1;
                tm_st8(col, v); TM_WAIT_ST();
                const long long t1 = clock64();
                while (clock64() - t1 < side_traffic) { }
                n++;
            }
            if (tid == 128) cycles[148 * 4 + blockIdx.x] = n;
        }
        TM_FENCE_BEFORE(); __syncthreads();
        return;
    }
    constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(24 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t b_base = smem_u32(smem);
    long long t0 = 0;
    if (warp < nwarps && elect_one())
    {
        uint64_t db[6];
        for (int j = 0; j < 6; j++) db[j] = make_desc(b_base + j * 768, 384, 128);
        t0 = clock64();
        for (int r = 0; r < rows; r++)
        {
            // d_slide: the accumulator window moves by d_slide columns per row inside a ring of 128 columns (the engine: 8 -> consecutive rows overlap in 16 of 24 columns)
            const uint32_t a = tmem + 64 * warp + 8 * (r % a_rotate), d = d_slide ? tmem + 384 + ((r * d_slide) % 104) : tmem + 384 + 32 * warp + 8 * (r & 1);
            mma_ts(d, a, db[0], idesc, 1); mma_ts_ashift(d, a, db[1], idesc, 1);
            mma_ts(d, a, db[2], idesc, 1); mma_ts_ashift(d, a, db[3], idesc, 1);
            mma_ts(d, a, db[4], idesc, 1); mma_ts(d, a, db[5], idesc, 1);
            if (commit_every && (r + 1) % commit_every == 0) tm_commit(smem_u32(bars + 4 + warp * 2 + (r & 1)));
        }
        tm_commit(smem_u32(bars + warp));
        mbar_wait(smem_u32(bars + warp), 0);
        cycles[blockIdx.x * 4 + warp] = clock64() - t0;
        atomicAdd(const_cast<int*>(stop), 1);
    }
    else if (warp < 4 && warp >= nwarps) { }
    TM_FENCE_BEFORE(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}
void run_multi(long long* dc4, int nwarps, int commit_every, int d_slide = 0, int side_traffic = 0)
{
    const int rows = 2000;
    cudaFuncSetAttribute(multi_issuer, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 256);
    cudaMemset(dc4, 0, 148 * 5 * 8);
    multi_issuer<<<148, 640, 48 * 1024 + 256>>>(dc4, rows, nwarps, commit_every, d_slide, 8, side_traffic);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> hc(148 * 5);
    cudaMemcpy(hc.data(), dc4, 148 * 5 * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int b = 0; b < 148; b++) { double m = 0; for (int w = 0; w < nwarps; w++) m = std::max(m, double(hc[b * 4 + w])); worst += m; }
    if (side_traffic) printf("[16 side warps: ld8 + st8 every %d cycles, %.0f items per warp] ", side_traffic, double(hc[148 * 4]));
    printf("%d issuer warp(s), D window slides %d columns per row, commit every %d rows: %s, %.1f cycles per row of 6 MMAs (N=24) per warp, %.1f cycles per row overall\n", nwarps, d_slide, commit_every, cudaGetErrorString(e),
           worst / 148 / rows, worst / 148 / rows / nwarps);
}

template<int N, int PATTERN, int COMMITS>
void run_rate(long long* dc)
{
    const int iters = 500;
    cudaFuncSetAttribute(rate_mma<N, PATTERN, COMMITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 128);
    rate_mma<N, PATTERN, COMMITS><<<148, 128, 48 * 1024 + 128>>>(dc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> hc(148);
    cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double a = 0; for (auto c : hc) a += double(c);
    printf("mma N=%3d pattern %d commits %d: %s, %.1f cycles per MMA\n", N, PATTERN, COMMITS, cudaGetErrorString(e), a / 148 / (iters * 12.0));
}

// what: 0 ld.x8, 1 ld.x16, 2 st.x8, 3 st.x16, 4 ld.x16 + st.x8 (the epilogue's mix); `batch` operations between waits
template<int WHAT>
__global__ void __launch_bounds__(512, 1) rate_ldst(long long* __restrict__ cycles, uint32_t* __restrict__ sink, int reps, int batch)
{
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    TM_FENCE_BEFORE(); __syncthreads(); TM_FENCE_AFTER();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    uint32_t v16[16], v8[8];
    for (int j = 0; j < 16; j++) v16[j] = tid + j;
    for (int j = 0; j < 8; j++) v8[j] = tid * 3 + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++)
    {
        for (int b = 0; b < batch; b++)
        {
            const uint32_t col = ((r * batch + b) * 16 + (warp >> 2) * 64) & 511 & ~15u;
            if (WHAT == 0) { tm_ld8(v8, tmem + lane_base + col); }
            else if (WHAT == 1) { tm_ld16(v16, tmem + lane_base + col); }
            else if (WHAT == 2) { tm_st8(tmem + lane_base + col, v8); }
            else if (WHAT == 3) { tm_st16(tmem + lane_base + col, v16); }
            else { tm_ld16(v16, tmem + lane_base + col); tm_st8(tmem + lane_base + ((col + 256) & 511), v8); }
        }
        if (WHAT <= 1 || WHAT == 4) { TM_WAIT_LD(); acc += v16[0] + v8[0] + v16[15] + v8[7]; }
        if (WHAT >= 2) TM_WAIT_ST();
    }
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    TM_FENCE_BEFORE(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

static double avg_cycles(long long* dc, int n)
{
    std::vector<long long> hc(n);
    cudaMemcpy(hc.data(), dc, n * 8, cudaMemcpyDeviceToHost);
    double a = 0; for (auto c : hc) a += double(c);
    return a / n;
}

int main()
{
    // ---- semantics ---------------------------------------------------------------------------------------------------------
    {
        float* dd; uint32_t* da; int* ds;
        cudaMalloc(&dd, 4 * 128 * N_CHK * 4); cudaMalloc(&da, 2 * 128 * 8 * 4); cudaMalloc(&ds, 4);
        cudaMemset(dd, 0xff, 4 * 128 * N_CHK * 4); cudaMemset(da, 0xff, 2 * 128 * 8 * 4);
        semantics<<<1, 128, 2 * N_CHK * 16 + 64>>>(dd, da, ds);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> hd(4 * 128 * N_CHK); std::vector<uint32_t> ha(2 * 128 * 8); int st = 0;
        cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ha.data(), da, ha.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
        printf("semantics: %s, barriers %s\n", cudaGetErrorString(e), st ? "ok" : "TIMED OUT");
        auto ref = [&](int row, int n) { float s = 0; for (int k = 0; k < 16; k++) s += float(a_val(row, k) * b_val(k, n)); return s; };
        const char* names[4] = { "D0 plain TS (D at column 80)", "D1 mma.ashift", "D2 plain after mma.ashift", "D3 plain after tcgen05.shift" };
        for (int w = 0; w < 4; w++)
        {
            // for every row find the source row whose reference it equals (s = src - row), tally
            int tally[5] = { 0, 0, 0, 0, 0 }, other = 0;     // shifts -2..+2
            std::vector<int> odd;
            for (int row = 0; row < 128; row++)
            {
                int found = 99;
                for (int s = -2; s <= 2 && found == 99; s++)
                {
                    const int src = row + s;
                    if (src < 0 || src > 127) continue;
                    bool eq = true;
                    for (int n = 0; n < N_CHK && eq; n++) eq = hd[(w * 128 + row) * N_CHK + n] == ref(src, n);
                    if (eq) found = s;
                }
                if (found == 99) { other++; odd.push_back(row); } else tally[found + 2]++;
            }
            printf("  %-32s rows equal to ref(row+s): s=-2:%d s=-1:%d s=0:%d s=+1:%d s=+2:%d other:%d;", names[w], tally[0], tally[1], tally[2], tally[3], tally[4], other);
            for (size_t i = 0; i < odd.size() && i < 8; i++) printf(" row %d (%.0f vs %.0f)", odd[i], hd[(w * 128 + odd[i]) * N_CHK], ref(odd[i], 0));
            printf("\n");
        }
        for (int w = 0; w < 2; w++)
        {
            int tally[5] = { 0, 0, 0, 0, 0 }, other = 0;
            printf("  A read back after %s: ", w == 0 ? "mma.ashift" : "tcgen05.shift");
            std::vector<int> odd;
            for (int row = 0; row < 128; row++)
            {
                int found = 99;
                for (int s = -2; s <= 2 && found == 99; s++)
                {
                    const int src = row + s;
                    if (src < 0 || src > 127) continue;
                    bool eq = true;
                    for (int c = 0; c < 8 && eq; c++)
                    {
                        const __half2 h = __floats2half2_rn(float(a_val(src, 2 * c)), float(a_val(src, 2 * c + 1)));
                        eq = ha[(w * 128 + row) * 8 + c] == *reinterpret_cast<const uint32_t*>(&h);
                    }
                    if (eq) found = s;
                }
                if (found == 99) { other++; odd.push_back(row); } else { tally[found + 2]++; if (found != 1) odd.push_back(row); }
            }
            printf("rows holding A[row+s]: s=-2:%d s=-1:%d s=0:%d s=+1:%d s=+2:%d other:%d; rows not at s=+1:", tally[0], tally[1], tally[2], tally[3], tally[4], other);
            for (size_t i = 0; i < odd.size() && i < 10; i++) printf(" %d", odd[i]);
            printf("\n");
        }
    }
    // ---- MMA rates ---------------------------------------------------------------------------------------------------------
    {
        long long* dc; cudaMalloc(&dc, 148 * 8);
        printf("patterns: 0 TS same A same D; 1 TS A and D rotate; 2 per row {ashift, ashift, plain}; 3 the same, 4 rows interleaved; 4 SS same D; 5 TS rotate, no accumulate\n");
        run_rate<8, 0, 0>(dc); run_rate<16, 0, 0>(dc); run_rate<24, 0, 0>(dc); run_rate<48, 0, 0>(dc); run_rate<96, 0, 0>(dc);
        run_rate<8, 1, 0>(dc); run_rate<16, 1, 0>(dc); run_rate<24, 1, 0>(dc); run_rate<48, 1, 0>(dc); run_rate<96, 1, 0>(dc);
        run_rate<16, 2, 0>(dc); run_rate<24, 2, 0>(dc); run_rate<48, 2, 0>(dc); run_rate<96, 2, 0>(dc);
        run_rate<16, 3, 0>(dc); run_rate<24, 3, 0>(dc); run_rate<48, 3, 0>(dc); run_rate<96, 3, 0>(dc);
        run_rate<16, 4, 0>(dc); run_rate<48, 4, 0>(dc); run_rate<96, 4, 0>(dc);
        run_rate<48, 5, 0>(dc);
        run_rate<24, 3, 1>(dc); run_rate<48, 3, 1>(dc); run_rate<48, 2, 1>(dc); run_rate<48, 4, 1>(dc); run_rate<48, 1, 1>(dc);
        { long long* dc4; cudaMalloc(&dc4, 148 * 5 * 8); for (int w = 1; w <= 4; w++) run_multi(dc4, w, 0); run_multi(dc4, 1, 1); run_multi(dc4, 4, 1); run_multi(dc4, 2, 1); run_multi(dc4, 1, 0, 8); run_multi(dc4, 1, 0, 24); run_multi(dc4, 1, 0, 16); run_multi(dc4, 4, 0, 8); run_multi(dc4, 4, 1, 8); run_multi(dc4, 4, 1, 8, 1); run_multi(dc4, 4, 1, 8, 200); run_multi(dc4, 4, 1, 8, 1000); cudaFree(dc4); }
        run_issue<0>(dc); run_issue<1>(dc); run_issue<2>(dc); run_issue<3>(dc); run_issue<4>(dc); run_issue<7>(dc); run_issue<15>(dc); run_issue<16 + 1 + 4>(dc); run_issue<32 + 1 + 4>(dc);
        cudaFree(dc);
    }
    // ---- TMEM load / store rates ---------------------------------------------------------------------------------------------
    {
        long long* dc; uint32_t* sink; cudaMalloc(&dc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
        const int reps = 2000;
        for (int warps : { 4, 8, 16 })
            for (int batch : { 1, 4 })
            {
                const int threads = warps * 32;
                double c[5];
                rate_ldst<0><<<148, threads>>>(dc, sink, reps, batch); cudaDeviceSynchronize(); c[0] = avg_cycles(dc, 148);
                rate_ldst<1><<<148, threads>>>(dc, sink, reps, batch); cudaDeviceSynchronize(); c[1] = avg_cycles(dc, 148);
                rate_ldst<2><<<148, threads>>>(dc, sink, reps, batch); cudaDeviceSynchronize(); c[2] = avg_cycles(dc, 148);
                rate_ldst<3><<<148, threads>>>(dc, sink, reps, batch); cudaDeviceSynchronize(); c[3] = avg_cycles(dc, 148);
                rate_ldst<4><<<148, threads>>>(dc, sink, reps, batch); cudaError_t e = cudaDeviceSynchronize(); c[4] = avg_cycles(dc, 148);
                const double ops = double(reps) * batch * warps;      // warp-level operations per SM
                printf("TMEM %2d warps, %d op(s) per wait: %s; ld.x8 %.0f B/clk/SM, ld.x16 %.0f, st.x8 %.0f, st.x16 %.0f, (ld.x16 + st.x8) %.1f cycles per warp pair\n", warps, batch,
                       cudaGetErrorString(e), ops * 1024 / c[0], ops * 2048 / c[1], ops * 1024 / c[2], ops * 2048 / c[3], c[4] / (double(reps) * batch));
            }
        cudaFree(dc); cudaFree(sink);
    }
    return 0;
}
