// k_chaing_d4.cu -- instances of the generic fused chain (chain_generic.cuh) for decimation 4: ntaps <= 64 / <= 128,
// nfft 512 / 1024 / 2048.  One translation unit per decimation so that the instances compile in parallel.
#include "chain_generic.cuh"

int lrc_chaing_launch_d4(const chaing::Args &a, int log2n) { return chaing::launch_decim<4>(a, log2n); }
bool lrc_chaing_has_d4(int ntaps, int log2n) { return chaing::has_decim<4>(ntaps, log2n); }
