// unpack.cuh -- the one definition of i2f on the device.
#pragma once
#include "common.cuh"

// rtlsdr.rs:159:  i as f32 / 127.0 - 1.0   -- one IEEE division, one IEEE subtraction.
//
// The division by the constant 127 is done with the FMA residual-correction sequence
//     q0 = RN(a * r),  r = RN(1/127);   rem = fma(-127, q0, a)  (exact);   q = fma(rem, r, q0)
// which yields the correctly rounded quotient a/127 for every a in {0..255} (checked exhaustively with
// exact rational arithmetic when this was written, and on the device by
// tests/test_gpu_core.py::test_unpack_all_256_values_bit_exact against the CPU restatement's real
// division).  Three full-rate FP32 instructions instead of the ~10-instruction IEEE division routine.
// All operations are explicit _rn intrinsics: never contracted or reassociated.
__device__ __forceinline__ float lr_div127(float a)
{
    const float r = 0.007874015718698502f;            // RN(1/127)
    const float q0 = __fmul_rn(a, r);
    const float rem = __fmaf_rn(-127.0f, q0, a);
    return __fmaf_rn(rem, r, q0);
}

__device__ __forceinline__ float lr_i2f(uint32_t b)
{
    return __fsub_rn(lr_div127(__uint2float_rn(b)), 1.0f);
}

// byte `k` (0..3) of word w as an exact float: 0x4B0000bb is the float 2^23 + bb
__device__ __forceinline__ float lr_byte_to_float(uint32_t w, int k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + k)), 8388608.0f);
}

__device__ __forceinline__ float lr_i2f_byte(uint32_t w, int k)
{
    return __fsub_rn(lr_div127(lr_byte_to_float(w, k)), 1.0f);
}
