// fm_core.cuh -- device pieces shared by the FM discriminator / resampler kernels (k_fm_resample.cu) and the fused
// config-3 receiver (k_fmrx.cu): the discriminator arithmetic and the packed window loop of the L = 1 decimators.
#pragma once
#include "common.cuh"

// atan2 for the discriminator: octant reduction to t = min/max in [0, 1] with one fast division, then
// atan(t) = t + t^3 P(t^2), P a degree-6 minimax fit (max abs error 1.1e-7 rad in f32 evaluation, i.e.
// > 120 dB below any usable deviation; the stage's bar is 100 dB SNR vs the f64 definition).  About
// 20 instructions instead of the ~45 of atan2f: the kernel drops from issue-bound to HBM-bound.
__device__ __forceinline__ float lr_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;        // atan2(0, 0) = 0 like atan2f / numpy
    const float u = __fmul_rn(t, t);
    float r = -0.0043553938157856464f;
    r = __fmaf_rn(r, u, 0.02304009348154068f);
    r = __fmaf_rn(r, u, -0.05777352675795555f);
    r = __fmaf_rn(r, u, 0.0979423001408577f);
    r = __fmaf_rn(r, u, -0.13976579904556274f);
    r = __fmaf_rn(r, u, 0.19962704181671143f);
    r = __fmaf_rn(r, u, -0.3333165943622589f);
    r = __fmaf_rn(__fmul_rn(r, u), t, t);
    if (ay > ax) r = __fsub_rn(1.57079632679489661923f, r);
    if (x < 0.f) r = __fsub_rn(3.14159265358979323846f, r);
    return copysignf(r, y);
}

// x[n] * conj(x[n-1]) with the rounding pinned (explicit fma/mul): the vectorised body and the scalar tail
// must give the same bits, or the output would depend on how a stream is chunked
__device__ __forceinline__ float2 fm_mul_conj(float2 a, float2 b)
{
    return make_float2(__fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)), __fmaf_rn(a.y, b.x, -__fmul_rn(a.x, b.y)));
}

// taps of an L = 1 decimator as a kernel parameter (constant bank -> uniform registers)
template <int NT_>
struct RsTaps { float g[NT_]; };

// The window loop, unrolled by template recursion (a `#pragma unroll` over ~175 iterations of packed builtins is
// not honoured by the compiler, and thousands of inline-asm statements in one block take minutes to compile):
// step J loads samples J, J+1 of both tiles with one LDS.128 and feeds the accumulators they belong to.
// PADP pairs of padding follow every STEPP pairs of the buffer (STEPP = a thread's window advance, so a window starts
// at a multiple of STEPP and sample J sits at J + PADP (J / STEPP) from it: compile-time offsets).  PADP = 0: dense.
template <int M, int R, int TPP, int J, int JEND, int STEPP = 1 << 30, int PADP = 0>
struct RsDec2Steps {
    template <class Taps>
    __device__ __forceinline__ static void run(const float2 *sx, const Taps &taps, float2 *acc)
    {
        if constexpr (J < JEND) {
            const float4 x = *reinterpret_cast<const float4 *>(sx + J + PADP * (J / STEPP));
            const float2 x0 = make_float2(x.x, x.y), x1 = make_float2(x.z, x.w);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k0 = J - r * M, k1 = k0 + 1;
                if (k0 >= 0 && k0 < TPP) acc[r] = __ffma2_rn(x0, make_float2(taps.g[k0], taps.g[k0]), acc[r]);
                if (k1 >= 0 && k1 < TPP) acc[r] = __ffma2_rn(x1, make_float2(taps.g[k1], taps.g[k1]), acc[r]);
            }
            RsDec2Steps<M, R, TPP, J + 2, JEND, STEPP, PADP>::run(sx, taps, acc);
        }
    }
};
