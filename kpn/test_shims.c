/* A C program written against the reference's own FFI declarations -- the prototypes of
 * libkissfft/kiss_fft.h:81-102 (what src/kissfft/src/kissfft.rs:11-16 binds) and of libsamplerate
 * (src/samplerate/src/samplerate.rs:32-42) -- linked with -lkissfft -lsamplerate from libredio_b200/:
 * the drop-in seams, exercised exactly as the unmodified reference would. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

typedef struct { float r, i; } kiss_fft_cpx;
typedef struct kiss_fft_state *kiss_fft_cfg;
kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);
void kiss_fft_cleanup(void);
int kiss_fft_next_fast_size(int n);
/* tools/kiss_fftr.h:21-43 */
typedef struct kiss_fftr_state *kiss_fftr_cfg;
kiss_fftr_cfg kiss_fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void kiss_fftr(kiss_fftr_cfg cfg, const float *timedata, kiss_fft_cpx *freqdata);
void kiss_fftri(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, float *timedata);

typedef struct {
    const float *data_in; float *data_out;
    long input_frames, output_frames, input_frames_used, output_frames_gen;
    int end_of_input; double src_ratio;
} SRC_DATA;
typedef struct SRC_STATE SRC_STATE;
SRC_STATE *src_new(int converter_type, int channels, int *error);
SRC_STATE *src_delete(SRC_STATE *state);
int src_process(SRC_STATE *state, SRC_DATA *data);
const char *src_strerror(int error);

#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(void)
{
    /* kissfft.rs:19-27: alloc once, then kiss_fft per frame */
    const int n = 1024;
    kiss_fft_cfg cfg = kiss_fft_alloc(n, 0, NULL, NULL);
    CHECK(cfg != NULL);
    kiss_fft_cpx *in = malloc(sizeof(kiss_fft_cpx) * n), *out = malloc(sizeof(kiss_fft_cpx) * n);
    for (int k = 0; k < n; ++k) { in[k].r = (float)cos(2 * M_PI * 37 * k / n) + 0.25f; in[k].i = (float)sin(2 * M_PI * 37 * k / n); }
    kiss_fft(cfg, in, out);
    CHECK(fabs(out[37].r - n) < 1e-2 && fabs(out[37].i) < 1e-2);       /* unit tone in bin 37, unscaled */
    CHECK(fabs(out[0].r - 0.25 * n) < 1e-2);
    CHECK(fabs(out[38].r) < 1e-2 && fabs(out[500].i) < 1e-2);
    kiss_fft(cfg, in, in);                                               /* fin == fout (kiss_fft.c:373-379) */
    CHECK(fabs(in[37].r - n) < 1e-2);
    size_t need = 0;
    CHECK(kiss_fft_alloc(n, 1, NULL, &need) == NULL && need > 0);        /* size query protocol, kiss_fft.h:66-78 */
    kiss_fft_cleanup();
    free(cfg);                                                           /* "can be simply free()d" kiss_fft.h:100-102 */

    /* a size kissfft factors as 2^3 5^3 (kf_factor, kiss_fft.c:309-330) and the next-fast-size helper */
    CHECK(kiss_fft_next_fast_size(997) == 1000 && kiss_fft_next_fast_size(1024) == 1024);
    {
        const int m = 1000;
        kiss_fft_cfg c2 = kiss_fft_alloc(m, 0, NULL, NULL);
        CHECK(c2 != NULL);
        for (int k = 0; k < m; ++k) { in[k].r = (float)cos(2 * M_PI * 11 * k / m); in[k].i = (float)sin(2 * M_PI * 11 * k / m); }
        kiss_fft(c2, in, out);
        CHECK(fabs(out[11].r - m) < 1e-2 && fabs(out[12].r) < 1e-2 && fabs(out[999].i) < 1e-2);
        free(c2);
    }
    /* the real-input pair the way tools/psdpng.c:139,165 and test/test_real.c use it */
    {
        const int m = 512;
        kiss_fftr_cfg fr = kiss_fftr_alloc(m, 0, NULL, NULL), ir = kiss_fftr_alloc(m, 1, NULL, NULL);
        CHECK(fr != NULL && ir != NULL);
        CHECK(kiss_fftr_alloc(511, 0, NULL, NULL) == NULL);              /* "Real FFT optimization must be even." */
        float t[512], back[512];
        kiss_fft_cpx F[257];
        for (int k = 0; k < m; ++k) t[k] = (float)cos(2 * M_PI * 5 * k / m) + 0.5f;
        kiss_fftr(fr, t, F);
        CHECK(fabs(F[5].r - m / 2) < 1e-2 && fabs(F[5].i) < 1e-2 && fabs(F[0].r - 0.5 * m) < 1e-2 && F[0].i == 0 && F[256].i == 0);
        kiss_fftri(ir, F, back);
        for (int k = 0; k < m; ++k) CHECK(fabs(back[k] - m * t[k]) < 2e-2);   /* unscaled round trip */
        free(fr); free(ir);
    }

    /* samplerate.rs:89-96: resample a 1000-sample sine by 2.0 and look at the length */
    int err = -1;
    SRC_STATE *st = src_new(1, 1, &err);
    CHECK(st != NULL && err == 0);
    float v[1000], o[2001];
    for (int k = 0; k < 1000; ++k) v[k] = sinf((float)k / 1000.0f);
    SRC_DATA d = { v, o, 1000, 2001, 0, 0, 0, 2.0 };
    CHECK(src_process(st, &d) == 0);
    CHECK(d.input_frames_used == 1000 && d.output_frames_gen == 2000);
    /* streaming: a second chunk continues the same stream */
    d.data_in = v; d.input_frames = 10; d.output_frames = 2001;
    CHECK(src_process(st, &d) == 0 && d.output_frames_gen == 20);
    /* a too-small output buffer consumes less input instead of overflowing */
    d.input_frames = 100; d.output_frames = 50;
    CHECK(src_process(st, &d) == 0 && d.output_frames_gen <= 50 && d.input_frames_used == 25);
    /* what samplerate.rs:64-84 meets when the library takes less than it was given: input_frames_used < input_frames; a
     * caller that re-submits the remainder gets the same stream as one call with enough room */
    {
        SRC_STATE *a = src_new(1, 1, &err), *b = src_new(1, 1, &err);
        CHECK(a && b);
        static float x[6000], ya[3100], yb[3100];
        for (int k = 0; k < 6000; ++k) x[k] = sinf(0.01f * (float)k) + 0.25f * sinf(0.37f * (float)k);
        SRC_DATA da = { x, ya, 6000, 3100, 0, 0, 0, 0.5 };
        CHECK(src_process(a, &da) == 0 && da.input_frames_used == 6000 && da.output_frames_gen == 3000);
        long used = 0, made = 0;
        int calls = 0;
        while (used < 6000 && calls < 100) {
            SRC_DATA db = { x + used, yb + made, 6000 - used, 700, 0, 0, 0, 0.5 };       /* room for 700 frames per call */
            CHECK(src_process(b, &db) == 0);
            CHECK(db.input_frames_used > 0 && db.output_frames_gen <= 700);
            used += db.input_frames_used; made += db.output_frames_gen; ++calls;
        }
        CHECK(used == 6000 && made == 3000 && calls >= 5);
        CHECK(memcmp(ya, yb, 3000 * sizeof(float)) == 0);
        src_delete(a); src_delete(b);
    }
    /* chunks may grow: libsamplerate puts no bound on input_frames (a first chunk of 10, then 100 000) */
    {
        SRC_STATE *g = src_new(1, 1, &err);
        static float big[100000], bo[50002];
        for (int k = 0; k < 100000; ++k) big[k] = sinf(0.001f * (float)k);
        SRC_DATA dg = { big, bo, 10, 50002, 0, 0, 0, 0.5 };
        CHECK(src_process(g, &dg) == 0 && dg.input_frames_used == 10);
        dg.data_in = big + 10; dg.input_frames = 99990;
        CHECK(src_process(g, &dg) == 0 && dg.input_frames_used == 99990 && dg.output_frames_gen == 49995);
        src_delete(g);
    }
    /* kiss_fft_alloc / free cycles: the cfg block is the caller's, the device plan is the shim's (no growth per cycle) */
    for (int k = 0; k < 200; ++k) {
        kiss_fft_cfg c3 = kiss_fft_alloc(256, k & 1, NULL, NULL);
        CHECK(c3 != NULL);
        free(c3);
    }
    d.src_ratio = 1e-9;
    int rc = src_process(st, &d);
    CHECK(rc != 0 && src_strerror(rc) != NULL);                          /* samplerate.rs:77-83 would panic with this text */
    src_delete(st);
    CHECK(src_new(1, 2, &err) == NULL && err != 0);
    printf("shims OK\n");
    return 0;
}
