// k_fir_gentile.cu -- dispatch of the generic-shape FIR tile kernel (fir_gentile.cuh; instances in k_firg_n*_d*.cu)
#include "fir_gentile.cuh"

bool lrc_fir_gentile_has(int ntaps, int decim)
{
    return ntaps >= 1 && ntaps <= 128 && (decim == 4 || decim == 5 || decim == 8 || decim == 10 || decim == 16);
}

// LRC_OK = launched; > 0 = error.  Caller checked lrc_fir_gentile_has.
int lrc_fir_gentile_launch(int n_sm, const float *h_taps, int ntaps, int decim, const float *d_in, size_t n_ch, size_t n_in,
                           size_t in_stride, float *d_out, size_t n_out, size_t out_stride, cudaStream_t s)
{
#define FG(NT_, D_) lrc_firg_launch_n##NT_##_d##D_(n_sm, h_taps, ntaps, d_in, n_ch, n_in, in_stride, d_out, n_out, out_stride, s)
    if (ntaps <= 64) {
        switch (decim) { case 4: return FG(64, 4); case 5: return FG(64, 5); case 8: return FG(64, 8); case 10: return FG(64, 10); case 16: return FG(64, 16); }
    } else {
        switch (decim) { case 4: return FG(128, 4); case 5: return FG(128, 5); case 8: return FG(128, 8); case 10: return FG(128, 10); case 16: return FG(128, 16); }
    }
#undef FG
    lrc_set_error("lrc_fir_gentile_launch: no instance for ntaps %d decim %d", ntaps, decim);
    return LRC_ERR_UNSUPPORTED;
}
