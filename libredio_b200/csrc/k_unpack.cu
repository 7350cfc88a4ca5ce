// k_unpack.cu -- u8 IQ -> interleaved complex f32, bit-exact with rtlsdr::i2f
// (src/rtlsdr/src/rtlsdr.rs:159-162):  i2f(b) = fl(fl(b)/127) - 1.
//
// HBM-bound: 2 B in, 8 B out per sample.  Each thread converts 16 input bytes (one 128-bit load)
// into 8 complex samples (four 128-bit stores).
#include "common.cuh"
#include "unpack.cuh"

__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t n_bytes)
{
    const size_t n_vec = n_bytes / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const uint4 q = ldg_stream_u4(reinterpret_cast<const uint4 *>(in) + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float4 *dst = reinterpret_cast<float4 *>(out) + i * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = lr_i2f(w[k] & 0xffu);
            o.y = lr_i2f((w[k] >> 8) & 0xffu);
            o.z = lr_i2f((w[k] >> 16) & 0xffu);
            o.w = lr_i2f(w[k] >> 24);
            stg_stream_f4(dst + k, o);
        }
    }
    // tail (< 16 bytes) and unaligned heads are handled by the scalar kernel below
}

__global__ void unpack_scalar_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t first,
                                     size_t n_bytes)
{
    const size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_bytes) out[i] = lr_i2f(in[i]);
}

extern "C" int lrc_unpack_u8_cf32(lrc_ctx *ctx, const uint8_t *d_iq, size_t n_bytes, float *d_out, void *stream)
{
    LRC_BIND(ctx);
    if (n_bytes & 1) {
        lrc_set_error("lrc_unpack_u8_cf32: %zu bytes is odd; the reference indexes i[1] out of bounds and "
                      "panics (rtlsdr.rs:161)", n_bytes);
        return LRC_ERR_ODD_LENGTH;
    }
    if (n_bytes == 0) return LRC_OK;
    LRC_REQUIRE(d_iq && d_out, LRC_ERR_INVALID, "lrc_unpack_u8_cf32: null buffer");
    cudaStream_t s = lrc_stream(ctx, stream);
    size_t vec_bytes = 0;
    if (((uintptr_t)d_iq & 15) == 0 && ((uintptr_t)d_out & 15) == 0) {
        vec_bytes = n_bytes / 16 * 16;
        if (vec_bytes) {
            size_t blocks = ceil_div(vec_bytes / 16, 256);
            const size_t cap = (size_t)ctx->n_sm * 8;
            if (blocks > cap) blocks = cap;
            unpack_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_iq, d_out, vec_bytes);
            LRC_CUDA(cudaGetLastError());
        }
    }
    if (vec_bytes < n_bytes) {
        const size_t rest = n_bytes - vec_bytes;
        unpack_scalar_kernel<<<(unsigned)ceil_div(rest, 256), 256, 0, s>>>(d_iq, d_out, vec_bytes, n_bytes);
        LRC_CUDA(cudaGetLastError());
    }
    return LRC_OK;
}
