// k_firg_n128_d10.cu -- fir_gentile_kernel<128, 10, R, 128> (fir_gentile.cuh): one instance per translation unit
#include "fir_gentile.cuh"

LRC_FIRG_DEFINE(128, 10)
