// k_chaing_d16_n128_l11.cu -- chain_gen_kernel<128, 16, 11, R> (chain_generic.cuh): one instance per translation unit
#include "chain_generic.cuh"

LRC_CHAING_DEFINE1(16, 128, 11)
