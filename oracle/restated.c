/*
 * oracle/restated.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see restated.h).
 *
 * Plain-C restatement of LibRedio's DSP hot path, written from the semantics of the cited
 * reference lines (paths relative to /root/reference).  Build: oracle/Makefile
 * (-O2 -ffp-contract=off -fno-fast-math: every float op rounds exactly once, in the order
 * written, like the reference's scalar Rust / -O0 C).
 */
#include "restated.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ====================================================================================== */
/* (1) unpack -- src/rtlsdr/src/rtlsdr.rs:159  fn i2f(i: u8) -> f32 {i as f32/127.0 - 1.0} */
/* ====================================================================================== */
float orc_i2f(uint8_t b)
{
    float q = (float)b / 127.0f;     /* one IEEE f32 division */
    return q - 1.0f;                 /* one IEEE f32 subtraction */
}

/* src/rtlsdr/src/rtlsdr.rs:160-162  data.chunks(2).map(|i| Complex{re:i2f(i[0]), im:i2f(i[1])}) */
size_t orc_data_to_samples(const uint8_t *data, size_t nbytes, orc_cpx *out)
{
    if (nbytes & 1) return (size_t)-1;          /* i[1] out of bounds -> panic in the Rust */
    for (size_t k = 0; k < nbytes / 2; ++k) {
        out[k].r = orc_i2f(data[2 * k]);
        out[k].i = orc_i2f(data[2 * k + 1]);
    }
    return nbytes / 2;
}

/* ====================================================================================== */
/* (2) FIR -- src/dsputils/src/dsputils.rs:30-32                                           */
/*   u.windows(v.len()).map(|x| x.zip(v).map(|(x,y)| x*y).fold(0, |a,b| a+b))              */
/* ====================================================================================== */
size_t orc_convolve_f32(const float *u, size_t nu, const float *v, size_t nv, float *y)
{
    if (nv == 0 || nu < nv) return 0;
    size_t ny = nu - nv + 1;
    for (size_t i = 0; i < ny; ++i) {
        float acc = 0.0f;                        /* Float::zero() */
        for (size_t j = 0; j < nv; ++j) {
            float p = u[i + j] * v[j];           /* product rounds first ... */
            acc = acc + p;                       /* ... then the running sum, left to right */
        }
        y[i] = acc;
    }
    return ny;
}

size_t orc_fir_decimate_cf32(const orc_cpx *x, size_t n, const float *taps, size_t m,
                             size_t d, orc_cpx *z, int full)
{
    if (m == 0 || d == 0 || n < m) return 0;
    size_t ny = n - m + 1;
    size_t nz = (ny - 1) / d + 1;                /* k in [0, floor((n-m)/d)] */
    if (full) {
        /* the composition a LibRedio user would write: convolve every lag, then stride */
        orc_cpx *y = (orc_cpx *)malloc(ny * sizeof(orc_cpx));
        for (size_t i = 0; i < ny; ++i) {
            float ar = 0.0f, ai = 0.0f;
            for (size_t j = 0; j < m; ++j) {
                float pr = x[i + j].r * taps[j];
                float pi = x[i + j].i * taps[j];
                ar = ar + pr;
                ai = ai + pi;
            }
            y[i].r = ar; y[i].i = ai;
        }
        for (size_t k = 0; k < nz; ++k) z[k] = y[k * d];
        free(y);
    } else {
        for (size_t k = 0; k < nz; ++k) {
            const orc_cpx *w = x + k * d;
            float ar = 0.0f, ai = 0.0f;
            for (size_t j = 0; j < m; ++j) {
                float pr = w[j].r * taps[j];
                float pi = w[j].i * taps[j];
                ar = ar + pr;
                ai = ai + pi;
            }
            z[k].r = ar; z[k].i = ai;
        }
    }
    return nz;
}

/* src/dsputils/src/dsputils.rs:38-51.  The reference evaluates cos(2*pi*n/(nn-1)) with
 * n = m (the length) and nn = the running index -- the arguments are swapped, and the last
 * term divides by cos(nn-1) instead of taking the cosine of the quotient.  faithful=1
 * reproduces that literally (index 1 -> division by zero -> NaN/inf); faithful=0 is the
 * evident intent, Blackman-Nuttall a0 - a1 cos(2 pi k/(m)) + a2 cos(4 pi k/m) - a3 cos(6 pi k/m)
 * over k = 0..m (m+1 points, as the reference's 0..m+1 range). */
void orc_window(size_t m, float *out, int faithful)
{
    const float a0 = 0.3635819f, a1 = 0.4891775f, a2 = 0.1365995f, a3 = 0.0106411f;
    const float pi = 3.14159265358979323846f;
    float n = (float)m;
    for (size_t x = 0; x < m + 1; ++x) {
        float nn = (float)x;
        if (faithful) {
            float t1 = a1 * cosf(2.0f * pi * n / (nn - 1.0f));
            float t2 = a2 * cosf(4.0f * pi * n / (nn - 1.0f));
            float t3 = a3 * (6.0f * pi * n / cosf(nn - 1.0f));
            out[x] = a0 - t1 + t2 - t3;
        } else {
            float t1 = a1 * cosf(2.0f * pi * nn / n);
            float t2 = a2 * cosf(4.0f * pi * nn / n);
            float t3 = a3 * cosf(6.0f * pi * nn / n);
            out[x] = a0 - t1 + t2 - t3;
        }
    }
}

/* src/dsputils/src/dsputils.rs:53-63 */
void orc_sinc(size_t m, float fc, float *out)
{
    const float pi = 3.14159265358979323846f;
    for (size_t x = 0; x < m; ++x) {
        float n = (float)x - (float)m / 2.0f;
        float r = 2.0f * fc;
        if (n != 0.0f) r = sinf(2.0f * pi * fc * n) / (pi * n);
        out[x] = r;
    }
}

/* src/dsputils/src/dsputils.rs:66-71 : zip(window(m), sinc(m, fc)) -> m products */
void orc_lpf(size_t m, float fc, float *out, int faithful)
{
    float *w = (float *)malloc((m + 1) * sizeof(float));
    float *s = (float *)malloc(m * sizeof(float));
    orc_window(m, w, faithful);
    orc_sinc(m, fc, s);
    for (size_t i = 0; i < m; ++i) out[i] = w[i] * s[i];
    free(w); free(s);
}

/* src/dsputils/src/dsputils.rs:74-79 : -lpf with 1.0 added at index m/2 - 1 (the sinc peaks at m/2: the
 * reference's delta sits one tap early; kept) */
void orc_hpf(size_t m, float fc, float *out, int faithful)
{
    orc_lpf(m, fc, out, faithful);
    for (size_t i = 0; i < m; ++i) out[i] = -out[i];
    out[m / 2 - 1] += 1.0f;
}

/* src/dsputils/src/dsputils.rs:82-88 : lpf(fc1) + hpf(fc2); the `-= 0.0` of :86 is a no-op */
void orc_bsf(size_t m, float fc1, float fc2, float *out, int faithful)
{
    float *lp = (float *)malloc(m * sizeof(float));
    float *hp = (float *)malloc(m * sizeof(float));
    orc_lpf(m, fc1, lp, faithful);
    orc_hpf(m, fc2, hp, faithful);
    for (size_t i = 0; i < m; ++i) out[i] = lp[i] + hp[i];
    out[m / 2 - 1] -= 0.0f;
    free(lp); free(hp);
}

/* src/dsputils/src/dsputils.rs:91-94 : -bsf */
void orc_bpf(size_t m, float fc1, float fc2, float *out, int faithful)
{
    orc_bsf(m, fc1, fc2, out, faithful);
    for (size_t i = 0; i < m; ++i) out[i] = -out[i];
}

/* ====================================================================================== */
/* (3) FFT -- src/kissfft/libkissfft/kiss_fft.c                                            */
/* ====================================================================================== */
typedef struct {
    int n, inverse;
    int radix[64], rest[64];     /* (p, m) pairs as kf_factor writes them, :309-330 */
    int nstages;
    orc_cpx *tw;                 /* tw[i] = exp(-+ 2 pi j i / n), double -> float, :357-363 */
} fft_plan;

static void plan_factor(fft_plan *pl)
{
    /* kiss_fft.c:309-330: pull out 4s, then 2s, then 3, 5, 7, ...; a candidate above
     * floor(sqrt(n)) means the remainder is prime */
    int n = pl->n, p = 4, k = 0;
    double root = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2; else if (p == 2) p = 3; else p += 2;
            if (p > root) p = n;
        }
        n /= p;
        pl->radix[k] = p; pl->rest[k] = n; ++k;
    } while (n > 1);
    pl->nstages = k;
}

static fft_plan *plan_make(int n, int inverse)
{
    fft_plan *pl = (fft_plan *)calloc(1, sizeof(fft_plan));
    pl->n = n; pl->inverse = inverse;
    pl->tw = (orc_cpx *)malloc(sizeof(orc_cpx) * (size_t)n);
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < n; ++i) {
        double ph = -2 * pi * i / n;             /* :359 */
        if (inverse) ph *= -1;
        pl->tw[i].r = (float)cos(ph);            /* _kiss_fft_guts.h:136-137: cast to float */
        pl->tw[i].i = (float)sin(ph);
    }
    plan_factor(pl);
    return pl;
}
static void plan_free(fft_plan *pl) { free(pl->tw); free(pl); }

static inline orc_cpx cmul(orc_cpx a, orc_cpx b)   /* C_MUL, _kiss_fft_guts.h:87-89 */
{
    orc_cpx m;
    m.r = a.r * b.r - a.i * b.i;
    m.i = a.r * b.i + a.i * b.r;
    return m;
}
static inline orc_cpx cadd(orc_cpx a, orc_cpx b) { orc_cpx c = { a.r + b.r, a.i + b.i }; return c; }
static inline orc_cpx csub(orc_cpx a, orc_cpx b) { orc_cpx c = { a.r - b.r, a.i - b.i }; return c; }

/* combine p sub-transforms of length m that sit back to back in F; `stride` is the twiddle
 * step of this level.  Radix 2 and 4 follow kf_bfly2 (:21-42) / kf_bfly4 (:44-90) operation
 * by operation so power-of-two sizes are bit-identical to the vendored library; every other
 * radix uses the O(p^2) form of kf_bfly_generic (:198-236) (kf_bfly3/5 are algebraic
 * shortcuts of the same sums and differ only in the last bits). */
static void combine(const fft_plan *pl, orc_cpx *F, int stride, int p, int m)
{
    const orc_cpx *tw = pl->tw;
    if (p == 2) {
        for (int k = 0; k < m; ++k) {
            orc_cpx t = cmul(F[m + k], tw[k * stride]);
            F[m + k] = csub(F[k], t);
            F[k] = cadd(F[k], t);
        }
    } else if (p == 4) {
        for (int k = 0; k < m; ++k) {
            orc_cpx b1 = cmul(F[m + k],     tw[k * stride]);
            orc_cpx b2 = cmul(F[2 * m + k], tw[2 * k * stride]);
            orc_cpx b3 = cmul(F[3 * m + k], tw[3 * k * stride]);
            orc_cpx dif02 = csub(F[k], b2);
            orc_cpx sum02 = cadd(F[k], b2);
            orc_cpx sum13 = cadd(b1, b3);
            orc_cpx dif13 = csub(b1, b3);
            F[2 * m + k] = csub(sum02, sum13);
            F[k]         = cadd(sum02, sum13);
            if (pl->inverse) {
                F[m + k].r     = dif02.r - dif13.i;  F[m + k].i     = dif02.i + dif13.r;
                F[3 * m + k].r = dif02.r + dif13.i;  F[3 * m + k].i = dif02.i - dif13.r;
            } else {
                F[m + k].r     = dif02.r + dif13.i;  F[m + k].i     = dif02.i - dif13.r;
                F[3 * m + k].r = dif02.r - dif13.i;  F[3 * m + k].i = dif02.i + dif13.r;
            }
        }
    } else {
        orc_cpx *tmp = (orc_cpx *)malloc(sizeof(orc_cpx) * (size_t)p);
        for (int u = 0; u < m; ++u) {
            for (int q = 0; q < p; ++q) tmp[q] = F[u + q * m];
            for (int q1 = 0; q1 < p; ++q1) {
                int k = u + q1 * m, twidx = 0;
                orc_cpx acc = tmp[0];
                for (int q = 1; q < p; ++q) {
                    twidx += stride * k;
                    if (twidx >= pl->n) twidx -= pl->n;
                    while (twidx >= pl->n) twidx -= pl->n;
                    acc = cadd(acc, cmul(tmp[q], tw[twidx]));
                }
                F[k] = acc;
            }
        }
        free(tmp);
    }
}

/* decimation in time, kf_work (:238-302): level `lv` splits into p interleaved
 * sub-sequences of length m; leaves copy the strided input (:276-280) */
static void dit(const fft_plan *pl, orc_cpx *out, const orc_cpx *in, int stride, int lv)
{
    int p = pl->radix[lv], m = pl->rest[lv];
    if (m == 1) {
        for (int q = 0; q < p; ++q) out[q] = in[(size_t)q * stride];
    } else {
        for (int q = 0; q < p; ++q)
            dit(pl, out + (size_t)q * m, in + (size_t)q * stride, stride * p, lv + 1);
    }
    combine(pl, out, stride, p, m);
}

int orc_fft(int nfft, int inverse, const orc_cpx *fin, orc_cpx *fout)
{
    if (nfft < 1 || fin == fout) return -1;
    if (nfft == 1) { fout[0] = fin[0]; return 0; }
    fft_plan *pl = plan_make(nfft, inverse);
    dit(pl, fout, fin, 1, 0);
    plan_free(pl);
    return 0;
}

/* tools/kiss_fastfir.c.  alloc :65-171, one block :173-187, block loop :192-206,
 * flush :208-226 */
size_t orc_fastfir(const orc_cpx *h, size_t nh, size_t nfft, const orc_cpx *in, size_t n,
                   orc_cpx *out, int flush)
{
    if (nh == 0) return 0;
    if (nfft == 0) {                              /* :81-93 next pow2 >= 2*nh, at least 1024 */
        size_t i = nh - 1; nfft = 2;
        do { nfft <<= 1; } while (i >>= 1);
        if (nfft < 1024) nfft = 1024;
    }
    if (nfft < nh) return 0;
    size_t ngood = nfft - nh + 1;                 /* :127 */
    fft_plan *fwd = plan_make((int)nfft, 0), *inv = plan_make((int)nfft, 1);
    orc_cpx *H = (orc_cpx *)malloc(sizeof(orc_cpx) * nfft);
    orc_cpx *tmp = (orc_cpx *)calloc(nfft, sizeof(orc_cpx));
    orc_cpx *freq = (orc_cpx *)malloc(sizeof(orc_cpx) * nfft);
    orc_cpx *blk = (orc_cpx *)malloc(sizeof(orc_cpx) * nfft);
    /* :148-154 rotate h left so the scrap lands at the tail */
    tmp[0] = h[nh - 1];
    for (size_t i = 0; i + 1 < nh; ++i) tmp[nfft - nh + 1 + i] = h[i];
    dit(fwd, H, tmp, 1, 0);
    float scale = (float)(1.0 / (double)nfft);    /* :159  scale = 1.0 / st->nfft (float) */
    for (size_t i = 0; i < nfft; ++i) { H[i].r *= scale; H[i].i *= scale; }

    size_t done = 0, nout = 0;
    while (n - done >= nfft) {                    /* kff_nocopy :199-204 */
        dit(fwd, freq, in + done, 1, 0);
        for (size_t i = 0; i < nfft; ++i) freq[i] = cmul(freq[i], H[i]);
        dit(inv, blk, freq, 1, 0);
        memcpy(out + nout, blk, sizeof(orc_cpx) * ngood);
        done += ngood; nout += ngood;
    }
    if (flush) {                                  /* kff_flush :213-225 */
        size_t rem = n - done, zpad = nfft - rem;
        memset(tmp, 0, sizeof(orc_cpx) * nfft);
        memcpy(tmp, in + done, sizeof(orc_cpx) * rem);
        dit(fwd, freq, tmp, 1, 0);
        for (size_t i = 0; i < nfft; ++i) freq[i] = cmul(freq[i], H[i]);
        dit(inv, blk, freq, 1, 0);
        if (ngood > zpad) {
            memcpy(out + nout, blk, sizeof(orc_cpx) * (ngood - zpad));
            nout += ngood - zpad;
        }
    }
    free(H); free(tmp); free(freq); free(blk);
    plan_free(fwd); plan_free(inv);
    return nout;
}

/* ====================================================================================== */
/* (5) OOK chain                                                                           */
/* ====================================================================================== */
float orc_norm(float re, float im)
{
    /* ratpak.rs:67 |x| x.norm(); num 0.1.22 Complex::norm = re.hypot(im).  Defined
     *  as the classic libm formulation below: two exact f64 products, one f64
     * add, one correctly rounded f64 sqrt, one narrowing -- reproducible on any IEEE machine */
    double s = (double)re * (double)re + (double)im * (double)im;
    return (float)sqrt(s);
}

void orc_trigger_init(orc_trigger *t)
{
    t->trigger = 0;                               /* bitfount.rs:42 */
    t->threshold = 0.0f;                          /* :44 */
    t->cap = 1 << 16;
    t->buf = (float *)malloc(t->cap * sizeof(float));
    t->buf[0] = 0.0f; t->len = 1;                 /* :43 vec!(0.0) */
}
void orc_trigger_free(orc_trigger *t) { free(t->buf); t->buf = NULL; t->len = t->cap = 0; }

/* Test hook: the OOM guard of bitfount.rs:52 is `1000*trigger_duration*block_size` = 25.6 M samples (100 s of capture).  A parity
 * test of the guard path shrinks it ON BOTH SIDES (here and LRC_OOK_TEST_GUARD_BLOCKS in the CUDA library); 0 restores the
 * reference's constant. */
static size_t g_trigger_guard_samples = 0;
void orc_test_set_trigger_guard(size_t samples) { g_trigger_guard_samples = samples; }

int orc_trigger_block(orc_trigger *t, const float *samples, size_t n, float **burst,
                      size_t *burst_len, float *block_sum_out)
{
    const long   trigger_duration = 50;           /* :40 */
    const size_t block_size = 512;                /* :38 */
    t->trigger -= 1;                              /* :46 */
    float s = 0.0f;                               /* :48 sum(), sequential from 0.0 */
    for (size_t i = 0; i < n; ++i) s = s + samples[i];
    if (block_sum_out) *block_sum_out = s;
    const size_t guard = g_trigger_guard_samples ? g_trigger_guard_samples : 1000u * (size_t)trigger_duration * block_size;
    if (t->len > guard) {                         /* :52-54 */
        t->buf[0] = 0.0f; t->len = 1;
    }
    if (t->threshold == 0.0f) t->threshold = s;   /* :57-59 */
    if (t->trigger < 0) {                         /* :62-65 */
        t->threshold = t->threshold + s / 1000.0f;
        t->threshold = t->threshold - t->threshold * 0.002f;
    }
    if (s > t->threshold * 4.0f) t->trigger = trigger_duration;     /* :68-70 */
    if (t->trigger > 1) {                         /* :73-75 push_all */
        if (t->len + n > t->cap) {
            while (t->len + n > t->cap) t->cap *= 2;
            t->buf = (float *)realloc(t->buf, t->cap * sizeof(float));
        }
        memcpy(t->buf + t->len, samples, n * sizeof(float));
        t->len += n;
    }
    if (t->trigger == 0) {                        /* :78-81 send, buffer = vec!() */
        *burst = (float *)malloc((t->len ? t->len : 1) * sizeof(float));
        memcpy(*burst, t->buf, t->len * sizeof(float));
        *burst_len = t->len;
        t->len = 0;
        return 1;
    }
    return 0;
}

void orc_discretize(const float *burst, size_t n, uint8_t *bits)
{
    float mx = 0.0f;                              /* bitfount.rs:90 fold(0.0, max) */
    for (size_t i = 0; i < n; ++i) mx = (burst[i] > mx) ? burst[i] : mx;
    float half = mx / 2.0f;                       /* :91 */
    for (size_t i = 0; i < n; ++i) bits[i] = (uint8_t)(burst[i] > half);
}

size_t orc_b2d(const uint8_t *bits, size_t n)     /* kpn.rs:111-113, MSB first */
{
    size_t v = 0;
    for (size_t i = 0; i < n; ++i) v += ((size_t)1 << (n - i - 1)) * bits[i];
    return v;
}

void orc_eat(const uint8_t *bits, const size_t *widths, size_t n_widths, size_t *out)
{                                                 /* kpn.rs:116-124 */
    size_t i = 0;
    for (size_t w = 0; w < n_widths; ++w) { out[w] = orc_b2d(bits + i, widths[w]); i += widths[w]; }
}

/* growable byte / u32 vectors */
typedef struct { uint8_t *p; size_t n, cap; } bytev;
static void bv_push(bytev *v, uint8_t b)
{
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (uint8_t *)realloc(v->p, v->cap); }
    v->p[v->n++] = b;
}
typedef struct { uint32_t *p; size_t n, cap; } u32v;
static void uv_push(u32v *v, uint32_t x)
{
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (uint32_t *)realloc(v->p, v->cap * 4); }
    v->p[v->n++] = x;
}

static inline int in_rng(float d, float lo, float hi) { return d >= lo && d <= hi; }  /* a...b inclusive */

/* shaper_optional, kpn.rs:266-275 */
typedef struct { uint8_t cur[64]; size_t n; size_t want; bytev out; size_t count; } shaper;
static void shaper_feed(shaper *s, int opt)        /* opt: 0/1 = Some(bit), -1 = None */
{
    if (opt >= 0) {
        if (s->n < sizeof(s->cur)) s->cur[s->n] = (uint8_t)opt;
        s->n++;                                    /* longer runs can never equal `want` */
    } else {
        if (s->n == s->want) {
            for (size_t i = 0; i < s->want; ++i) bv_push(&s->out, s->cur[i]);
            s->count++;
        }
        s->n = 0;
    }
}

int orc_ook_decode(const uint8_t *iq, size_t n_blocks, unsigned s_rate, orc_ook_result *res)
{
    memset(res, 0, sizeof(*res));
    orc_trigger trg; orc_trigger_init(&trg);
    bytev bits = {0}; u32v rv = {0}, rl = {0};
    res->block_sums = (float *)malloc((n_blocks ? n_blocks : 1) * sizeof(float));
    res->n_blocks = n_blocks;
    orc_cpx cs[512]; float env[512];
    for (size_t b = 0; b < n_blocks; ++b) {
        /* bitfount.rs:24 data_to_samples on 1024 bytes; ratpak.rs:64-68 norm per sample */
        orc_data_to_samples(iq + b * 1024, 1024, cs);
        for (int i = 0; i < 512; ++i) env[i] = orc_norm(cs[i].r, cs[i].i);
        float *burst = NULL; size_t blen = 0;
        if (orc_trigger_block(&trg, env, 512, &burst, &blen, &res->block_sums[b])) {
            /* discretize sends one usize per sample into ONE continuous stream, :92-94 */
            uint8_t *bb = (uint8_t *)malloc(blen ? blen : 1);
            orc_discretize(burst, blen, bb);
            for (size_t i = 0; i < blen; ++i) bv_push(&bits, bb[i]);
            free(bb); free(burst);
            res->n_bursts++;
        }
    }
    orc_trigger_free(&trg);

    /* rle, kpn.rs:17-29: a run is emitted when the value changes; the final run never is */
    if (bits.n > 0) {
        uint8_t x = bits.p[0]; uint32_t i = 1;
        for (size_t k = 1; k < bits.n; ++k) {
            uint8_t y = bits.p[k];
            if (y != x) { uv_push(&rv, x); uv_push(&rl, i); i = 1; } else i = i + 1;
            x = y;
        }
    }

    /* dle, kpn.rs:32-38: (x, ct as f32 / s_rate as f32); then fork -> two loopers
     * (ratpak.rs:88-92 proto A, :93-97 proto B) -> shaper_optional 36 / 24 (:102-110) */
    shaper sa; memset(&sa, 0, sizeof sa); sa.want = 36;
    shaper sb; memset(&sb, 0, sizeof sb); sb.want = 24;
    size_t nr = rv.n;
    float *dur = (float *)malloc((nr ? nr : 1) * sizeof(float));
    for (size_t k = 0; k < nr; ++k) dur[k] = (float)rl.p[k] / (float)s_rate;

    for (size_t k = 0; k < nr; ) {                 /* proto A */
        uint32_t v = rv.p[k]; float d = dur[k]; ++k;
        if (v == 1 && in_rng(d, 2e-4f, 6e-4f)) {
            if (k >= nr) break;                    /* a.next().unwrap() on a closed port: panic */
            uint32_t v2 = rv.p[k]; float e = dur[k]; ++k;
            if (v2 == 0 && in_rng(e, 1.5e-3f, 2.5e-3f)) shaper_feed(&sa, 0);
            else if (v2 == 0 && in_rng(e, 3.5e-3f, 4.5e-3f)) shaper_feed(&sa, 1);
            else shaper_feed(&sa, -1);
        } else shaper_feed(&sa, -1);
    }
    for (size_t k = 0; k < nr; ) {                 /* proto B */
        uint32_t v = rv.p[k]; float d = dur[k]; ++k;
        if (v == 1 && (in_rng(d, 125e-6f, 250e-6f) || in_rng(d, 500e-6f, 650e-6f))) {
            if (k >= nr) break;
            uint32_t v2 = rv.p[k]; float e = dur[k]; ++k;
            if (v2 == 0 && (in_rng(e, 500e-6f, 650e-6f) || in_rng(e, 125e-6f, 250e-6f)))
                shaper_feed(&sb, d > e ? 1 : 0);
            else shaper_feed(&sb, -1);
        } else shaper_feed(&sb, -1);
    }
    free(dur);

    res->a_bits = sa.out.p; res->a_count = sa.count;
    res->b_bits = sb.out.p; res->b_count = sb.count;
    res->bits = bits.p; res->n_bits = bits.n;
    res->run_val = rv.p; res->run_len = rl.p; res->n_runs = nr;
    return 0;
}

void orc_ook_free(orc_ook_result *res)
{
    free(res->a_bits); free(res->b_bits); free(res->block_sums); free(res->bits);
    free(res->run_val); free(res->run_len);
    memset(res, 0, sizeof(*res));
}

/* ====================================================================================== */
/* CPU baseline of the headline chain (bench.py only)                                      */
/* ====================================================================================== */
size_t orc_chain_psd(const orc_cpx *x, size_t n, const float *taps, size_t m, size_t d,
                     int nfft, const float *window, double *psd,
                     orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn)
{
    return orc_chain_psd_mode(x, n, taps, m, d, nfft, window, psd, alloc_fn, fft_fn, 0);
}

/* full = 1: the FIR is dsputils::convolve as written (dsputils.rs:30-32: EVERY lag is computed) followed by
 * keeping every d-th output -- the composition a LibRedio user has today; full = 0 computes kept outputs only. */
size_t orc_chain_psd_mode(const orc_cpx *x, size_t n, const float *taps, size_t m, size_t d,
                          int nfft, const float *window, double *psd,
                          orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn, int full)
{
    if (n < m) return 0;
    size_t nz = (n - m) / d + 1;
    size_t nframes = nz / (size_t)nfft;
    void *cfg = alloc_fn ? alloc_fn(nfft, 0, NULL, NULL) : NULL;
    orc_cpx *frame = (orc_cpx *)malloc(sizeof(orc_cpx) * (size_t)nfft);
    orc_cpx *spec = (orc_cpx *)malloc(sizeof(orc_cpx) * (size_t)nfft);
    for (size_t f = 0; f < nframes; ++f) {
        orc_fir_decimate_cf32(x + f * (size_t)nfft * d, (size_t)(nfft - 1) * d + m, taps, m, d, frame, full);
        for (int i = 0; i < nfft; ++i) { frame[i].r *= window[i]; frame[i].i *= window[i]; }
        if (fft_fn) fft_fn(cfg, frame, spec); else orc_fft(nfft, 0, frame, spec);
        for (int i = 0; i < nfft; ++i)            /* tools/psdpng.c:165-166 */
            psd[i] += (double)(spec[i].r * spec[i].r + spec[i].i * spec[i].i);
    }
    free(frame); free(spec); free(cfg);
    return nframes;
}

/* batch of frames through the vendored kiss_fft in a C loop (CPU baseline timing only: a Python loop over
 * kiss_fft() calls measures the interpreter, not the library) */
size_t orc_kissfft_batch(const orc_cpx *x, orc_cpx *y, int nfft, int inverse, size_t batch,
                         orc_kiss_alloc_fn alloc_fn, orc_kiss_fft_fn fft_fn)
{
    void *cfg = alloc_fn ? alloc_fn(nfft, inverse, NULL, NULL) : NULL;
    for (size_t f = 0; f < batch; ++f) {
        if (fft_fn) fft_fn(cfg, x + f * (size_t)nfft, y + f * (size_t)nfft);
        else orc_fft(nfft, inverse, x + f * (size_t)nfft, y + f * (size_t)nfft);
    }
    free(cfg);
    return batch;
}
