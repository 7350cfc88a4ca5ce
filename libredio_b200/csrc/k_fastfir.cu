// k_fastfir.cu -- long FIR by FFT overlap-save ("overlap-scrap"), one kernel per call.
//
// Replaces kiss_fastfir_alloc / kiss_fastfir (libkissfft/tools/kiss_fastfir.c:65-245):
//   H = FFT(h rotated left by nh-1) / nfft                      (:148-169)
//   per block b: out[b*ngood .. +ngood) = IFFT(FFT(in[b*ngood .. +nfft)) .* H)[0 .. ngood)   (:173-204)
//   flush: the remainder zero-padded to nfft, ngood - zpad outputs kept               (:208-226)
// One CTA owns one block at a time: forward FFT, spectrum multiply and inverse FFT all happen in
// registers/shared memory (the forward output layout "thread t holds bin t + e*T" is exactly the
// inverse input layout), so HBM sees nfft reads and ngood writes per block and nothing else.
#include "fft_core.cuh"
#include <vector>

using namespace lrfft;

int lrc_make_twiddles(int nfft, float2 **d_tw);
int lrc_log2_exact(int n);

struct lrc_fastfir {
    lrc_ctx *ctx;
    size_t   nh, nfft, ngood;
    int      log2n;
    float2  *d_tw;
    float2  *d_H;
};

template <int LOG2N>
__global__ void __launch_bounds__(CtaFFT<LOG2N, false>::T)
fastfir_kernel(const float2 *__restrict__ in, size_t n_in, float2 *__restrict__ out, size_t n_blocks_full,
               size_t n_blocks, size_t ngood, size_t flush_keep, const float2 *__restrict__ tw,
               const float2 *__restrict__ H)
{
    using FF = CtaFFT<LOG2N, false>;
    using FI = CtaFFT<LOG2N, true>;
    constexpr int N = FF::N, E = FF::E, T = FF::T;
    extern __shared__ float2 sm[];
    const int t = threadIdx.x;
    for (size_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const size_t s0 = b * ngood;
        const size_t avail = n_in - s0;              // < N only for the flush block
        float2 v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const size_t i = (size_t)t + (size_t)e * T;
            v[e] = (i < avail) ? __ldcs(in + s0 + i) : make_float2(0.f, 0.f);
        }
        FF::run(v, sm, tw, t, SyncCta{});
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = t + e * T;
            v[e] = cmulf(v[e], __ldg(H + i));        // C_MUL(freqbuf[i], fir_freq_resp[i])  :180-184
        }
        FI::run(v, sm, tw, t, SyncCta{});
        const size_t keep = (b < n_blocks_full) ? ngood : flush_keep;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const size_t i = (size_t)t + (size_t)e * T;
            if (i < keep) __stcs(out + s0 + i, v[e]);
        }
    }
}

extern "C" int lrc_fastfir_create(lrc_ctx *ctx, const float *h_taps_cpx, size_t nh, size_t nfft, lrc_fastfir **out)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(out && h_taps_cpx && nh >= 1, LRC_ERR_INVALID, "lrc_fastfir_create: bad arguments");
    if (nfft == 0) {
        // kiss_fastfir.c:81-93: next power of two at least twice the impulse response, at least 1024
        size_t i = nh - 1;
        nfft = 2;
        do { nfft <<= 1; } while (i >>= 1);
        if (nfft < 1024) nfft = 1024;
    }
    const int l2 = lrc_log2_exact((int)nfft);
    if (l2 < 1 || l2 > 13 || nfft > 8192) {
        lrc_set_error("lrc_fastfir_create: nfft=%zu: only powers of two in [2, 8192] (nh <= 4096 with the "
                      "automatic size)", nfft);
        return LRC_ERR_UNSUPPORTED;
    }
    LRC_REQUIRE(nfft >= nh, LRC_ERR_INVALID, "lrc_fastfir_create: nfft shorter than the impulse response");
    lrc_fastfir *f = new (std::nothrow) lrc_fastfir{ctx, nh, nfft, nfft - nh + 1, l2, nullptr, nullptr};
    LRC_REQUIRE(f != nullptr, LRC_ERR_NOMEM, "out of host memory");
    int rc = lrc_make_twiddles((int)nfft, &f->d_tw);
    if (rc) { delete f; return rc; }
    // rotated impulse response (:148-154), transformed with our own FFT kernel, scaled by 1/nfft (:159-169)
    const float2 *h = reinterpret_cast<const float2 *>(h_taps_cpx);
    std::vector<float2> rot(nfft, make_float2(0.f, 0.f));
    rot[0] = h[nh - 1];
    for (size_t i = 0; i + 1 < nh; ++i) rot[nfft - nh + 1 + i] = h[i];
    lrc_fft *plan = nullptr;
    rc = lrc_fft_create(ctx, (int)nfft, 0, &plan);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMalloc(&f->d_H, nfft * sizeof(float2));
    if (!rc && e == cudaSuccess) e = cudaMemcpy(f->d_H, rot.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice);
    if (!rc && e == cudaSuccess) rc = lrc_fft_run(plan, (const float *)f->d_H, (float *)f->d_H, 1, ctx->stream);
    if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e == cudaSuccess) e = cudaMemcpy(rot.data(), f->d_H, nfft * sizeof(float2), cudaMemcpyDeviceToHost);
    if (!rc && e == cudaSuccess) {
        const float scale = (float)(1.0 / (double)nfft);
        for (size_t i = 0; i < nfft; ++i) { rot[i].x *= scale; rot[i].y *= scale; }
        e = cudaMemcpy(f->d_H, rot.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice);
    }
    lrc_fft_destroy(plan);
    if (rc || e != cudaSuccess) {
        if (e != cudaSuccess) { lrc_set_error("lrc_fastfir_create: %s", cudaGetErrorString(e)); rc = LRC_ERR_CUDA; }
        lrc_fastfir_destroy(f);
        return rc;
    }
    *out = f;
    return LRC_OK;
}

extern "C" int lrc_fastfir_destroy(lrc_fastfir *f)
{
    if (!f) return LRC_OK;
    cudaSetDevice(f->ctx->device);
    cudaFree(f->d_tw); cudaFree(f->d_H);
    delete f;
    return LRC_OK;
}

extern "C" size_t lrc_fastfir_nfft(const lrc_fastfir *f) { return f ? f->nfft : 0; }

static void fastfir_counts(const lrc_fastfir *f, size_t n_in, int flush, size_t *full, size_t *flush_keep)
{
    // kff_nocopy :199-204: while (n >= nfft) { ...; n -= ngood; }
    *full = n_in >= f->nfft ? (n_in - f->nfft) / f->ngood + 1 : 0;
    const size_t rem = n_in - *full * f->ngood;
    // kff_flush :213-225: zpad = nfft - rem; keep ngood - zpad = rem - (nh - 1) samples (none if negative)
    *flush_keep = (flush && rem + 1 > f->nh) ? rem + 1 - f->nh : 0;
}

extern "C" size_t lrc_fastfir_out_len(const lrc_fastfir *f, size_t n_in, int flush)
{
    if (!f) return 0;
    size_t full, keep;
    fastfir_counts(f, n_in, flush, &full, &keep);
    return full * f->ngood + keep;
}

template <int LOG2N>
static int launch_fastfir(lrc_fastfir *f, const float2 *in, size_t n_in, float2 *out, size_t full, size_t nblk,
                          size_t keep, cudaStream_t s)
{
    using FF = CtaFFT<LOG2N, false>;
    const int threads = FF::T;
    const size_t smem = (size_t)FF::SMEM_CPX * sizeof(float2);
    auto kern = fastfir_kernel<LOG2N>;
    if (smem > 48 * 1024) LRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    LRC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    size_t blocks = (size_t)f->ctx->n_sm * occ;
    if (blocks > nblk) blocks = nblk;
    kern<<<(unsigned)blocks, threads, smem, s>>>(in, n_in, out, full, nblk, f->ngood, keep, f->d_tw, f->d_H);
    LRC_CUDA(cudaGetLastError());
    return LRC_OK;
}

extern "C" int lrc_fastfir_run(lrc_fastfir *f, const float *d_in, size_t n_in, float *d_out, int flush,
                               size_t *n_out, void *stream)
{
    LRC_REQUIRE(f != nullptr, LRC_ERR_INVALID, "null plan");
    LRC_BIND(f->ctx);
    size_t full, keep;
    fastfir_counts(f, n_in, flush, &full, &keep);
    if (n_out) *n_out = full * f->ngood + keep;
    const size_t nblk = full + (keep ? 1 : 0);
    if (nblk == 0) return LRC_OK;
    LRC_REQUIRE(d_in && d_out && d_in != d_out, LRC_ERR_INVALID, "lrc_fastfir_run: null or aliased buffers");
    LRC_REQUIRE(((uintptr_t)d_in & 7) == 0 && ((uintptr_t)d_out & 7) == 0, LRC_ERR_INVALID, "lrc_fastfir_run: misaligned");
    cudaStream_t s = lrc_stream(f->ctx, stream);
    const float2 *in = (const float2 *)d_in;
    float2 *out = (float2 *)d_out;
    switch (f->log2n) {
#define FF_CASE(L) case L: return launch_fastfir<L>(f, in, n_in, out, full, nblk, keep, s);
        FF_CASE(1) FF_CASE(2) FF_CASE(3) FF_CASE(4) FF_CASE(5) FF_CASE(6) FF_CASE(7)
        FF_CASE(8) FF_CASE(9) FF_CASE(10) FF_CASE(11) FF_CASE(12) FF_CASE(13)
#undef FF_CASE
    }
    return LRC_ERR_UNSUPPORTED;
}
