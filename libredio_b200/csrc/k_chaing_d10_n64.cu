// k_chaing_d10_n64.cu -- chain_gen_kernel<64, 10, 9|10|11, R> (chain_generic.cuh): the instances of one (decimation, NTAPS) pair
#include "chain_generic.cuh"

LRC_CHAING_DEFINE(10, 64)
