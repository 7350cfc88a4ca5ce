/*
 * oracle/restated.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C11, strict IEEE-754 binary32, no FMA contraction) of the
 * LibRedio sample-stream DSP hot path.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (libredio_b200/libredio_cuda.so) never links or calls it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - orc_fft / orc_fastfir  : pinned against the UNMODIFIED vendored kissfft C compiled
 *                              into oracle/_ref (bit-exact for power-of-two sizes) and the
 *                              golden vector of test/fft.py:95-98.
 *   - unpack / convolve / OOK: the reference holds no test, fixture or golden vector for
 *                              these; pinned only by hand-computed
 *                              micro-cases -> "parity pinned by restatement + KATs".
 *   - resampler              : libsamplerate is an un-vendored, un-pinned system library
 *                              -> PARITY UNPINNED (see oracle/defined_f64.py).
 */
#ifndef ORACLE_RESTATED_H
#define ORACLE_RESTATED_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float r, i; } orc_cpx;

/* ---- (1) unpack: src/rtlsdr/src/rtlsdr.rs:159-162 ------------------------------------ */
float  orc_i2f(uint8_t b);
/* returns number of complex samples written, or (size_t)-1 on odd length (the Rust indexes
 * i[1] of a 1-element chunk and panics) */
size_t orc_data_to_samples(const uint8_t *data, size_t nbytes, orc_cpx *out);

/* ---- (2) FIR: src/dsputils/src/dsputils.rs:30-32 ------------------------------------- */
/* valid-mode correlation, taps not reversed, left fold from 0; returns nu-nv+1 (0 if nu<nv) */
size_t orc_convolve_f32(const float *u, size_t nu, const float *v, size_t nv, float *y);
/* cf32 with real taps = convolve on the re and im planes independently, then keep every
 * d-th output (decimation is defined by the north-star).  `full` != 0 computes
 * every output first and strides afterwards (what composing the reference's blocks would do);
 * the results are identical either way. returns number of outputs */
size_t orc_fir_decimate_cf32(const orc_cpx *x, size_t n, const float *taps, size_t m,
                             size_t d, orc_cpx *z, int full);
/* tap designer: dsputils.rs:38-71.  `faithful` != 0 reproduces window()'s swapped arguments
 * (NaN at index 1); 0 gives the intended Blackman-Nuttall * sinc low-pass (documented
 * deviation). out has m entries */
void   orc_window(size_t m, float *out_m_plus_1, int faithful);
void   orc_sinc(size_t m, float fc, float *out_m);
void   orc_lpf(size_t m, float fc, float *out_m, int faithful);
void   orc_hpf(size_t m, float fc, float *out_m, int faithful);
void   orc_bsf(size_t m, float fc1, float fc2, float *out_m, int faithful);
void   orc_bpf(size_t m, float fc1, float fc2, float *out_m, int faithful);

/* ---- (3) FFT: src/kissfft/libkissfft/kiss_fft.c:238-388 ------------------------------ */
/* unscaled mixed-radix DIT, forward e^{-j..}, inverse e^{+j..}; fin != fout */
int    orc_fft(int nfft, int inverse, const orc_cpx *fin, orc_cpx *fout);
/* overlap-scrap fast FIR: tools/kiss_fastfir.c:65-245.  Processes every full block of the
 * n input samples, then (if flush) the zero-padded remainder. nfft==0 -> auto size.
 * returns number of outputs written */
size_t orc_fastfir(const orc_cpx *h, size_t nh, size_t nfft, const orc_cpx *in, size_t n,
                   orc_cpx *out, int flush);

/* ---- (5) OOK chain ------------------------------------------------------------------ */
/* envelope |x|: ratpak.rs:64-68 -> num::Complex::norm = hypot(re, im); defined as
 * (float)sqrt((double)re*re + (double)im*im) */
float  orc_norm(float re, float im);

/* trigger state machine, bitfount.rs:36-85.  Feed blocks of 512 envelope samples. */
typedef struct {
    long   trigger;        /* isize trigger            :41 */
    float  threshold;      /* f32 threshold            :44 */
    float *buf;            /* sample_buffer            :43 (starts as [0.0]) */
    size_t len, cap;
} orc_trigger;
void   orc_trigger_init(orc_trigger *t);
void   orc_trigger_free(orc_trigger *t);
void   orc_test_set_trigger_guard(size_t samples);   /* test hook, 0 = the reference's 1000*50*512 */
/* returns 1 and hands out the burst (*burst malloc'd, caller frees) when a burst is sent */
int    orc_trigger_block(orc_trigger *t, const float *samples, size_t n, float **burst,
                         size_t *burst_len, float *block_sum_out);

/* discretize, bitfount.rs:87-96: bits[i] = x[i] > max/2 */
void   orc_discretize(const float *burst, size_t n, uint8_t *bits);

/* packets out of one stream */
typedef struct {
    /* proto A: 36-bit packets (ratpak.rs:102-105), proto B: 24-bit (:107-110) */
    uint8_t *a_bits; size_t a_count;          /* a_count packets x 36 bits */
    uint8_t *b_bits; size_t b_count;          /* b_count packets x 24 bits */
    /* intermediate products, for stage-by-stage parity */
    float   *block_sums; size_t n_blocks;     /* s per 512-block */
    uint8_t *bits; size_t n_bits;             /* flattened discretize output */
    uint32_t *run_val; uint32_t *run_len; size_t n_runs;   /* rle output (last run not flushed) */
    size_t  n_bursts;
} orc_ook_result;

/* whole chain on a finite capture of u8 IQ (nbytes = n_blocks*1024):
 * data_to_samples -> norm -> trigger -> discretize -> rle -> dle(s_rate) -> matchers A/B
 * -> shaper_optional(36 / 24).  ratpak.rs:60-111, kpn.rs:17-38,148-150,266-275 */
int    orc_ook_decode(const uint8_t *iq, size_t n_blocks, unsigned s_rate, orc_ook_result *res);
void   orc_ook_free(orc_ook_result *res);

/* kpn.rs:111-124 */
size_t orc_b2d(const uint8_t *bits, size_t n);
void   orc_eat(const uint8_t *bits, const size_t *widths, size_t n_widths, size_t *out);

/* ---- CPU baseline of the headline chain (bench.py only) ------------------------------ */
/* cf32 -> FIR m taps / d -> frames of nfft -> Hann -> FFT (kiss_fft function pointer, so the
 * caller can pass the vendored reference build) -> |X|^2 accumulated into psd[nfft] (f64).
 * returns number of frames */
typedef void *(*orc_kiss_alloc_fn)(int, int, void *, size_t *);
typedef void  (*orc_kiss_fft_fn)(void *, const orc_cpx *, orc_cpx *);
size_t orc_chain_psd(const orc_cpx *x, size_t n, const float *taps, size_t m, size_t d,
                     int nfft, const float *window, double *psd,
                     orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn);
size_t orc_chain_psd_mode(const orc_cpx *x, size_t n, const float *taps, size_t m, size_t d,
                          int nfft, const float *window, double *psd,
                          orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn, int full);

#ifdef __cplusplus
}
#endif
size_t orc_kissfft_batch(const orc_cpx *x, orc_cpx *y, int nfft, int inverse, size_t batch,
                         orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn);
size_t orc_chain_psd_mode(const orc_cpx *x, size_t n, const float *taps, size_t m, size_t d,
                          int nfft, const float *window, double *psd,
                          orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn, int full);

#endif
