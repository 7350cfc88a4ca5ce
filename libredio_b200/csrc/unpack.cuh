// unpack.cuh -- the one definition of i2f on the device.
#pragma once
#include "common.cuh"

// rtlsdr.rs:159:  i as f32 / 127.0 - 1.0   -- one IEEE division, one IEEE subtraction.
// __fdiv_rn / __fsub_rn are the correctly rounded operations and are never contracted or
// replaced by reciprocal approximations whatever the compile flags, so this matches the CPU
// restatement (oracle/restated.c orc_i2f) on all 256 inputs; tests check that exhaustively.
__device__ __forceinline__ float lr_i2f(uint32_t b)
{
    return __fsub_rn(__fdiv_rn(__uint2float_rn(b), 127.0f), 1.0f);
}
