// k_chaing_d10_n128_l10.cu -- chain_gen_kernel<128, 10, 10, R> (chain_generic.cuh): one instance per translation unit
#include "chain_generic.cuh"

LRC_CHAING_DEFINE1(10, 128, 10)
