// k_chaing_d8_n128_l10.cu -- chain_gen_kernel<128, 8, 10, R> (chain_generic.cuh): one instance per translation unit
#include "chain_generic.cuh"

LRC_CHAING_DEFINE1(8, 128, 10)
