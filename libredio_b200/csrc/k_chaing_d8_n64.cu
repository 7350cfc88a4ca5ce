// k_chaing_d8_n64.cu -- chain_gen_kernel<64, 8, 9|10|11, R> (chain_generic.cuh): the instances of one (decimation, NTAPS) pair
#include "chain_generic.cuh"

LRC_CHAING_DEFINE(8, 64)
