// k_fm_resample.cu -- quadrature FM discriminator and rational polyphase resampler.
//
// FM discriminator: absent from the reference (north-star stage): d[n] = arg(x[n] conj(x[n-1])).
// Resampler: replaces samplerate::resample (src/samplerate/src/samplerate.rs:59-87), whose arithmetic
// lives in libsamplerate (un-vendored, un-pinned -> parity UNPINNED).  Our definition, mirrored in f64
// by oracle/defined_f64.py:
//     ratio = L/M;  prototype h: ntaps = 2*32*max(L,M)+1, sinc(2 fc (i-c)) * kaiser(beta=12),
//     fc = 0.45/max(L,M), sum(h) = L;
//     y[m] = sum_j h[(mM mod L) + jL] * x[floor(mM/L) - j],  x[<0] = 0   (streaming-causal).
#include "common.cuh"
#include <cmath>
#include <vector>

// ---------------------------------------------------------------------------------------------
// FM discriminator
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fmdemod_kernel(const float2 *__restrict__ in, size_t n_ch, size_t n, size_t in_stride,
               const float2 *__restrict__ state, float *__restrict__ out, size_t out_stride)
{
    const size_t total = n_ch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / n, k = i % n;
        const float2 *x = in + c * in_stride;
        const float2 cur = x[k];
        float2 prev;
        if (k > 0) prev = x[k - 1];
        else prev = state ? state[c] : make_float2(0.f, 0.f);
        const float2 z = cmul_conjb(cur, prev);
        out[c * out_stride + k] = atan2f(z.y, z.x);
    }
}

__global__ void fmdemod_state_kernel(const float2 *__restrict__ in, size_t n_ch, size_t n, size_t in_stride,
                                     float2 *__restrict__ state)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_ch) state[c] = in[c * in_stride + n - 1];
}

extern "C" int lrc_fmdemod_run(lrc_ctx *ctx, const float *d_in, size_t n_ch, size_t n, size_t in_stride,
                               float *d_state, float *d_out, size_t out_stride, void *stream)
{
    LRC_BIND(ctx);
    if (n_ch == 0 || n == 0) return LRC_OK;
    LRC_REQUIRE(d_in && d_out, LRC_ERR_INVALID, "lrc_fmdemod_run: null buffer");
    LRC_REQUIRE(in_stride >= n && out_stride >= n, LRC_ERR_INVALID, "lrc_fmdemod_run: stride shorter than length");
    LRC_REQUIRE(((uintptr_t)d_in & 7) == 0, LRC_ERR_INVALID, "lrc_fmdemod_run: input must be 8-byte aligned");
    cudaStream_t s = lrc_stream(ctx, stream);
    size_t blocks = ceil_div(n_ch * n, 256);
    const size_t cap = (size_t)ctx->n_sm * 16;
    if (blocks > cap) blocks = cap;
    fmdemod_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float2 *)d_in, n_ch, n, in_stride, (const float2 *)d_state,
                                                   d_out, out_stride);
    LRC_CUDA(cudaGetLastError());
    if (d_state) {
        fmdemod_state_kernel<<<(unsigned)ceil_div(n_ch, 256), 256, 0, s>>>((const float2 *)d_in, n_ch, n, in_stride,
                                                                         (float2 *)d_state);
        LRC_CUDA(cudaGetLastError());
    }
    return LRC_OK;
}

// ---------------------------------------------------------------------------------------------
// resampler
// ---------------------------------------------------------------------------------------------
static const int    RS_ZERO_CROSSINGS = 32;
static const double RS_KAISER_BETA = 12.0;
static const double RS_BANDWIDTH = 0.9;
static const int    RS_MAX_DEN = 4096;

struct lrc_resampler {
    lrc_ctx *ctx;
    int      L, M;
    size_t   n_ch, max_chunk, tpp, ntaps, cap;   // cap: floats per staging row
    std::vector<double> h;                       // prototype (f64)
    float   *d_hp;                               // [L][tpp] polyphase rows, f32
    float   *d_buf;                              // [n_ch][cap]: tpp-1 history + chunk
    float   *d_carry;                            // [n_ch][tpp]
    float   *d_hin, *d_hout; size_t hin_cap, hout_cap;   // staging for the host entry point (floats)
    unsigned long long n_total, m_next;          // inputs consumed / next output index (same for all channels)
};

static double bessel_i0(double x)
{
    double s = 1.0, t = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        t *= q / ((double)k * (double)k);
        s += t;
        if (t < 1e-18 * s) break;
    }
    return s;
}

// best rational approximation with bounded denominator (continued fractions, as Fraction.limit_denominator)
static bool rational(double x, int max_den, int *num, int *den)
{
    if (!(x > 0) || !std::isfinite(x)) return false;
    long long p0 = 0, q0 = 1, p1 = 1, q1 = 0;
    double r = x;
    for (int it = 0; it < 64; ++it) {
        const double fl = floor(r);
        if (fl > 1e15) break;
        const long long a = (long long)fl;
        const long long p2 = a * p1 + p0, q2 = a * q1 + q0;
        if (q2 > max_den || p2 > max_den) break;
        p0 = p1; q0 = q1; p1 = p2; q1 = q2;
        const double frac = r - fl;
        if (frac < 1e-15) break;
        r = 1.0 / frac;
    }
    if (q1 <= 0 || p1 <= 0) return false;
    if (fabs((double)p1 / (double)q1 - x) > 1e-12 * x) return false;
    *num = (int)p1; *den = (int)q1;
    return true;
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
resample_kernel(const float *__restrict__ buf, size_t cap, size_t n_ch, const float *__restrict__ hp, int L, int M,
                int tpp, unsigned long long n_total, unsigned long long m_next, size_t n_out,
                float *__restrict__ out, size_t out_stride)
{
    const size_t total = n_ch * n_out;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / n_out, k = i % n_out;
        const unsigned long long t = (m_next + k) * (unsigned long long)M;
        const unsigned long long base = t / (unsigned long long)L;
        const int phase = (int)(t % (unsigned long long)L);
        // buffer index of x[base]: the row starts at global input index n_total - (tpp - 1)
        const float *x = buf + c * cap + (size_t)(base - n_total) + (size_t)(tpp - 1);
        const float *h = hp + (size_t)phase * tpp;
        float acc[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc[u] = 0.f;
        int j = 0;
        for (; j + UNROLL <= tpp; j += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc[u] = fmaf(__ldg(h + j + u), x[-(j + u)], acc[u]);
        }
        for (; j < tpp; ++j) acc[0] = fmaf(__ldg(h + j), x[-j], acc[0]);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) s += acc[u];
        out[c * out_stride + k] = s;
    }
}

extern "C" int lrc_resampler_create(lrc_ctx *ctx, double ratio, size_t n_ch, size_t max_chunk, lrc_resampler **out)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(out && n_ch >= 1 && max_chunk >= 1, LRC_ERR_INVALID, "lrc_resampler_create: bad arguments");
    int L = 0, M = 0;
    if (!rational(ratio, RS_MAX_DEN, &L, &M)) {
        lrc_set_error("lrc_resampler_create: ratio %.17g is not L/M with L, M <= %d (libsamplerate accepts any "
                      "ratio in [1/256, 256])", ratio, RS_MAX_DEN);
        return LRC_ERR_UNSUPPORTED;
    }
    lrc_resampler *r = new (std::nothrow) lrc_resampler();
    LRC_REQUIRE(r != nullptr, LRC_ERR_NOMEM, "out of host memory");
    r->ctx = ctx; r->L = L; r->M = M; r->n_ch = n_ch; r->max_chunk = max_chunk;
    const int q = L > M ? L : M;
    r->ntaps = (size_t)2 * RS_ZERO_CROSSINGS * q + 1;
    r->tpp = (r->ntaps + L - 1) / L;
    r->h.resize(r->ntaps);
    const double pi = 3.14159265358979323846264338327950288;
    const double c = (double)(r->ntaps - 1) / 2.0, fc = 0.5 * RS_BANDWIDTH / q, i0b = bessel_i0(RS_KAISER_BETA);
    double sum = 0.0;
    for (size_t i = 0; i < r->ntaps; ++i) {
        const double u = 2.0 * fc * ((double)i - c);
        const double snc = (u == 0.0) ? 1.0 : sin(pi * u) / (pi * u);
        const double a = 2.0 * (double)i / (double)(r->ntaps - 1) - 1.0;
        const double arg = 1.0 - a * a;
        const double kw = bessel_i0(RS_KAISER_BETA * sqrt(arg > 0 ? arg : 0.0)) / i0b;
        r->h[i] = snc * kw;
        sum += r->h[i];
    }
    for (size_t i = 0; i < r->ntaps; ++i) r->h[i] *= (double)L / sum;
    std::vector<float> hp((size_t)L * r->tpp, 0.f);
    for (int p = 0; p < L; ++p)
        for (size_t j = 0; j < r->tpp; ++j) {
            const size_t idx = (size_t)p + j * (size_t)L;
            if (idx < r->ntaps) hp[(size_t)p * r->tpp + j] = (float)r->h[idx];
        }
    r->cap = (r->tpp + max_chunk + 3) / 4 * 4;
    r->d_hp = r->d_buf = r->d_carry = nullptr;
    r->d_hin = r->d_hout = nullptr; r->hin_cap = r->hout_cap = 0;
    r->n_total = 0; r->m_next = 0;
    if (cudaMalloc(&r->d_hp, hp.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&r->d_buf, n_ch * r->cap * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&r->d_carry, n_ch * r->tpp * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(r->d_hp, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        lrc_set_error("lrc_resampler_create: %s", cudaGetErrorString(cudaGetLastError()));
        lrc_resampler_destroy(r);
        return LRC_ERR_CUDA;
    }
    *out = r;
    return lrc_resampler_reset(r);
}

extern "C" int lrc_resampler_destroy(lrc_resampler *r)
{
    if (!r) return LRC_OK;
    cudaSetDevice(r->ctx->device);
    cudaFree(r->d_hp); cudaFree(r->d_buf); cudaFree(r->d_carry); cudaFree(r->d_hin); cudaFree(r->d_hout);
    delete r;
    return LRC_OK;
}

extern "C" int lrc_resampler_reset(lrc_resampler *r)
{
    LRC_REQUIRE(r != nullptr, LRC_ERR_INVALID, "null plan");
    LRC_BIND(r->ctx);
    r->n_total = 0; r->m_next = 0;
    LRC_CUDA(cudaMemsetAsync(r->d_buf, 0, r->n_ch * r->cap * sizeof(float), r->ctx->stream));   // x[<0] = 0
    LRC_CUDA(cudaStreamSynchronize(r->ctx->stream));
    return LRC_OK;
}

extern "C" int lrc_resampler_get_taps(const lrc_resampler *r, double *h_taps, size_t cap, size_t *ntaps, int *L, int *M)
{
    LRC_REQUIRE(r != nullptr, LRC_ERR_INVALID, "null plan");
    if (ntaps) *ntaps = r->ntaps;
    if (L) *L = r->L;
    if (M) *M = r->M;
    if (h_taps) {
        LRC_REQUIRE(cap >= r->ntaps, LRC_ERR_CAPACITY, "lrc_resampler_get_taps: buffer too small");
        memcpy(h_taps, r->h.data(), r->ntaps * sizeof(double));
    }
    return LRC_OK;
}

extern "C" size_t lrc_resampler_next_out_len(const lrc_resampler *r, size_t n_in)
{
    if (!r || n_in == 0) return 0;
    // all m with floor(mM/L) <= N-1  <=>  m <= (N L - 1)/M
    const unsigned long long N = r->n_total + n_in;
    const unsigned long long m_end = (N * (unsigned long long)r->L - 1) / (unsigned long long)r->M + 1;
    return (size_t)(m_end - r->m_next);
}

extern "C" int lrc_resampler_process(lrc_resampler *r, const float *d_in, size_t n_in, size_t in_stride,
                                     float *d_out, size_t out_stride, size_t *n_out, void *stream)
{
    LRC_REQUIRE(r && n_out, LRC_ERR_INVALID, "lrc_resampler_process: null argument");
    LRC_BIND(r->ctx);
    *n_out = 0;
    if (n_in == 0) return LRC_OK;
    LRC_REQUIRE(n_in <= r->max_chunk, LRC_ERR_CAPACITY, "lrc_resampler_process: chunk longer than max_chunk");
    LRC_REQUIRE(d_in && in_stride >= n_in, LRC_ERR_INVALID, "lrc_resampler_process: bad input");
    cudaStream_t s = lrc_stream(r->ctx, stream);
    const size_t hist = r->tpp - 1, row = r->cap * sizeof(float);
    LRC_CUDA(cudaMemcpy2DAsync(r->d_buf + hist, row, d_in, in_stride * sizeof(float), n_in * sizeof(float), r->n_ch,
                               cudaMemcpyDeviceToDevice, s));
    const size_t no = lrc_resampler_next_out_len(r, n_in);
    if (no) {
        LRC_REQUIRE(d_out && out_stride >= no, LRC_ERR_CAPACITY, "lrc_resampler_process: output too small "
                    "(the reference sizes it ratio*len + 1, samplerate.rs:64)");
        size_t blocks = ceil_div(r->n_ch * no, 256);
        const size_t cap = (size_t)r->ctx->n_sm * 16;
        if (blocks > cap) blocks = cap;
        resample_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(r->d_buf, r->cap, r->n_ch, r->d_hp, r->L, r->M, (int)r->tpp,
                                                           r->n_total, r->m_next, no, d_out, out_stride);
        LRC_CUDA(cudaGetLastError());
    }
    // new history = the last tpp-1 floats of [history | chunk]
    if (hist) {
        const size_t crow = r->tpp * sizeof(float);
        LRC_CUDA(cudaMemcpy2DAsync(r->d_carry, crow, r->d_buf + n_in, row, hist * sizeof(float), r->n_ch,
                                   cudaMemcpyDeviceToDevice, s));
        LRC_CUDA(cudaMemcpy2DAsync(r->d_buf, row, r->d_carry, crow, hist * sizeof(float), r->n_ch,
                                   cudaMemcpyDeviceToDevice, s));
    }
    r->n_total += n_in;
    r->m_next += no;
    *n_out = no;
    return LRC_OK;
}

extern "C" int lrc_resampler_process_host(lrc_resampler *r, const float *h_in, size_t n_in, float *h_out,
                                          size_t out_cap, size_t *n_out)
{
    LRC_REQUIRE(r && n_out, LRC_ERR_INVALID, "lrc_resampler_process_host: null argument");
    LRC_BIND(r->ctx);
    *n_out = 0;
    if (n_in == 0) return LRC_OK;
    LRC_REQUIRE(h_in != nullptr, LRC_ERR_INVALID, "lrc_resampler_process_host: null input");
    const size_t no = lrc_resampler_next_out_len(r, n_in);
    LRC_REQUIRE(no == 0 || (h_out && out_cap >= no), LRC_ERR_CAPACITY, "lrc_resampler_process_host: output too small");
    if (r->hin_cap < r->n_ch * n_in) {
        cudaFree(r->d_hin); r->d_hin = nullptr; r->hin_cap = 0;
        LRC_CUDA(cudaMalloc(&r->d_hin, r->n_ch * n_in * sizeof(float)));
        r->hin_cap = r->n_ch * n_in;
    }
    if (r->hout_cap < r->n_ch * (no + 1)) {
        cudaFree(r->d_hout); r->d_hout = nullptr; r->hout_cap = 0;
        LRC_CUDA(cudaMalloc(&r->d_hout, r->n_ch * (no + 1) * sizeof(float)));
        r->hout_cap = r->n_ch * (no + 1);
    }
    cudaStream_t s = r->ctx->stream;
    LRC_CUDA(cudaMemcpyAsync(r->d_hin, h_in, r->n_ch * n_in * sizeof(float), cudaMemcpyHostToDevice, s));
    size_t got = 0;
    int rc = lrc_resampler_process(r, r->d_hin, n_in, n_in, r->d_hout, no + 1, &got, s);
    if (rc) return rc;
    if (got)
        LRC_CUDA(cudaMemcpy2DAsync(h_out, out_cap * sizeof(float), r->d_hout, (no + 1) * sizeof(float), got * sizeof(float),
                                   r->n_ch, cudaMemcpyDeviceToHost, s));
    LRC_CUDA(cudaStreamSynchronize(s));
    *n_out = got;
    return LRC_OK;
}
