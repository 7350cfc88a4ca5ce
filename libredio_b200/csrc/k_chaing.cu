// k_chaing.cu -- dispatch of the generic fused chain (chain_generic.cuh; instances in k_chaing_d*_n*.cu)
#include "chain_generic.cuh"

bool lrc_chaing_has(int ntaps, int decim, int log2n, bool all)
{
    switch (decim) {
        case 4:  return chaing::has_decim<4>(ntaps, log2n, all);
        case 5:  return chaing::has_decim<5>(ntaps, log2n, all);
        case 8:  return chaing::has_decim<8>(ntaps, log2n, all);
        case 10: return chaing::has_decim<10>(ntaps, log2n, all);
        case 16: return chaing::has_decim<16>(ntaps, log2n, all);
        default: return false;
    }
}

// LRC_OK = launched, -1 = no instance for this shape (caller runs unfused), > 0 = LRC error
int lrc_chaing_launch(const chaing::Args &a, int decim, int log2n)
{
    if (a.ntaps < 1 || a.ntaps > 128 || log2n < 9 || log2n > 11) return -1;
    if (a.ntaps <= 64) {
        switch (decim) {
            case 4:  return lrc_chaing_launch_d4_n64(a, log2n);
            case 5:  return lrc_chaing_launch_d5_n64(a, log2n);
            case 8:  return lrc_chaing_launch_d8_n64(a, log2n);
            case 10: return lrc_chaing_launch_d10_n64(a, log2n);
            case 16: return lrc_chaing_launch_d16_n64(a, log2n);
            default: return -1;
        }
    }
#define BIG(D_) case D_: return log2n == 9 ? lrc_chaing_launch_d##D_##_n128_l9(a) : log2n == 10 ? lrc_chaing_launch_d##D_##_n128_l10(a) : lrc_chaing_launch_d##D_##_n128_l11(a)
    switch (decim) { BIG(4); BIG(5); BIG(8); BIG(10); BIG(16); default: return -1; }
#undef BIG
}
