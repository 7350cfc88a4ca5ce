// k_firg_n64_d10.cu -- fir_gentile_kernel<64, 10, R, 128> (fir_gentile.cuh): one instance per translation unit
#include "fir_gentile.cuh"

LRC_FIRG_DEFINE(64, 10)
