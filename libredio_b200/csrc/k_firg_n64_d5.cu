// k_firg_n64_d5.cu -- fir_gentile_kernel<64, 5, R, 128> (fir_gentile.cuh): one instance per translation unit
#include "fir_gentile.cuh"

LRC_FIRG_DEFINE(64, 5)
