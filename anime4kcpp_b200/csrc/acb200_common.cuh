// Shared device helpers for the B200 backend: element-type codecs with the reference's exact
// rounding rules, and the tile geometry constants.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/acb200.h"

namespace acb
{
    // ---- element codecs: core/internal/AC/Core/Internal/Util.hpp:50-78 ----------------------------
    // toFloat<u8/u16> is a true division by max (not a reciprocal multiply); fromFloat saturates to
    // [0,1] then `v*max + 0.5f` truncated, with the multiply and add rounded separately (the
    // reference's translation units are built without FMA contraction).
    __device__ __forceinline__ float sat01(float v) { return v < 0.0f ? 0.0f : (v < 1.0f ? v : 1.0f); }

    // q / MAXV correctly rounded for every integer q in [0, MAXV], MAXV = 255 or 65535, without the division sequence:
    // one Newton correction of q * (1/MAXV).  Equal to __fdiv_rn(q, MAXV) for all 256 / 65536 inputs (checked exhaustively,
    // tests/test_oracle_cpu.py::test_division_free_to_float_is_exact restates the check in numpy).
    template<int MAXV>
    __device__ __forceinline__ float unit_from_int(float q)
    {
        constexpr float RCP = 1.0f / static_cast<float>(MAXV);
        const float y = __fmul_rn(q, RCP);
        return __fmaf_rn(__fmaf_rn(-static_cast<float>(MAXV), y, q), RCP, y);
    }

    __device__ __forceinline__ float load_elem(const void* row, int x, int type)
    {
        switch (type)
        {
        case ACB200_UINT8: return unit_from_int<255>(static_cast<float>(static_cast<const uint8_t*>(row)[x]));
        case ACB200_UINT16: return unit_from_int<65535>(static_cast<float>(static_cast<const uint16_t*>(row)[x]));
        case ACB200_FLOAT16: return __half2float(static_cast<const __half*>(row)[x]);
        default: return static_cast<const float*>(row)[x];
        }
    }
    __device__ __forceinline__ uint8_t quant_u8(float v) { return static_cast<uint8_t>(__fadd_rn(__fmul_rn(sat01(v), 255.0f), 0.5f)); }
    __device__ __forceinline__ uint16_t quant_u16(float v) { return static_cast<uint16_t>(__fadd_rn(__fmul_rn(sat01(v), 65535.0f), 0.5f)); }
    __device__ __forceinline__ void store_elem(void* row, int x, int type, float v)
    {
        switch (type)
        {
        case ACB200_UINT8: static_cast<uint8_t*>(row)[x] = quant_u8(v); break;
        case ACB200_UINT16: static_cast<uint16_t*>(row)[x] = quant_u16(v); break;
        case ACB200_FLOAT16: static_cast<__half*>(row)[x] = __float2half_rn(sat01(v)); break;
        default: static_cast<float*>(row)[x] = sat01(v); break;
        }
    }
    // value an element of this type holds after a store/load round trip (what the next stage sees)
    __device__ __forceinline__ float requant(float v, int type)
    {
        switch (type)
        {
        case ACB200_UINT8: return unit_from_int<255>(static_cast<float>(quant_u8(v)));
        case ACB200_UINT16: return unit_from_int<65535>(static_cast<float>(quant_u16(v)));
        case ACB200_FLOAT16: return __half2float(__float2half_rn(sat01(v)));
        default: return sat01(v);
        }
    }
    // store two horizontally adjacent elements (x even); vectorised when the row is suitably aligned
    __device__ __forceinline__ void store_elem2(void* row, int x, int type, float v0, float v1, bool aligned)
    {
        if (aligned)
        {
            switch (type)
            {
            case ACB200_UINT8: *reinterpret_cast<uchar2*>(static_cast<uint8_t*>(row) + x) = make_uchar2(quant_u8(v0), quant_u8(v1)); return;
            case ACB200_UINT16: *reinterpret_cast<ushort2*>(static_cast<uint16_t*>(row) + x) = make_ushort2(quant_u16(v0), quant_u16(v1)); return;
            case ACB200_FLOAT16: *reinterpret_cast<__half2*>(static_cast<__half*>(row) + x) = __floats2half2_rn(sat01(v0), sat01(v1)); return;
            default: *reinterpret_cast<float2*>(static_cast<float*>(row) + x) = make_float2(sat01(v0), sat01(v1)); return;
            }
        }
        store_elem(row, x, type, v0);
        store_elem(row, x + 1, type, v1);
    }

    __host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
}
