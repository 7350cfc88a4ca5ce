// k_firg_n128_d4.cu -- fir_gentile_kernel<128, 4, R, 128> (fir_gentile.cuh): one instance per translation unit
#include "fir_gentile.cuh"

LRC_FIRG_DEFINE(128, 4)
