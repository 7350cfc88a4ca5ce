// k_chaing_d5_n128_l9.cu -- chain_gen_kernel<128, 5, 9, R> (chain_generic.cuh): one instance per translation unit
#include "chain_generic.cuh"

LRC_CHAING_DEFINE1(5, 128, 9)
