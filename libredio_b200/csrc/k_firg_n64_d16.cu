// k_firg_n64_d16.cu -- fir_gentile_kernel<64, 16, R, 128> (fir_gentile.cuh): one instance per translation unit
#include "fir_gentile.cuh"

LRC_FIRG_DEFINE(64, 16)
