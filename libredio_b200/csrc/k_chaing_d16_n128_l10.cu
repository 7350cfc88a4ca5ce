// k_chaing_d16_n128_l10.cu -- chain_gen_kernel<128, 16, 10, R> (chain_generic.cuh): one instance per translation unit
#include "chain_generic.cuh"

LRC_CHAING_DEFINE1(16, 128, 10)
