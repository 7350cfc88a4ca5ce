// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2 / add.rn.f32x2, sm_100a) issue and
// pipe throughput.  Decides whether complex arithmetic in the FFT / FIR kernels should be written packed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/f32x2 tools/ubench/f32x2.cu && gpurun_out/f32x2
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64 *)&r) : "l"(*(u64 *)&a), "l"(*(u64 *)&b), "l"(*(u64 *)&c));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64 *)&r) : "l"(*(u64 *)&a), "l"(*(u64 *)&b));
    return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c)
{
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float add1(float a, float b)
{
    float r;
    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

constexpr int CH = 8;        // independent chains per thread (complex values)
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) kern(float2 *out, float2 seed)
{
    float2 v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 w = make_float2(seed.x * 0.5f, seed.y * 0.25f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) { v[i].x = fma1(v[i].x, w.x, w.y); v[i].y = fma1(v[i].y, w.x, w.y); }     // 2 FFMA
            if (MODE == 1) { v[i] = fma2(v[i], w, w); }                                            // 1 FFMA2
            if (MODE == 2) { v[i].x = add1(v[i].x, w.x); v[i].y = add1(v[i].y, w.y); }             // 2 FADD
            if (MODE == 3) { v[i] = add2(v[i], w); }                                               // 1 FADD2
            if (MODE == 4) { v[i] = fma2(v[i], w, w); v[i].x = fma1(v[i].x, w.x, w.y); }           // FFMA2 + FFMA
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FIR-shaped use: acc[r] = fma2(x, {h[r], h[r]}, acc[r]) with the taps h[] uniform (kernel parameters).  ptxas
// encodes the broadcast tap as a uniform-register operand (FFMA2 R, R, UR.F32, R); the scalar form reads the
// constant bank directly (FFMA R, R, c[0x0][..], R).  MODE 5: packed + uniform tap, MODE 6: scalar + constant tap,
// MODE 7: packed with the tap copied to a per-thread register pair first.
struct Taps { float h[CH]; };
template <int MODE>
__global__ void __launch_bounds__(256) kern_taps(float2 *out, float2 seed, const __grid_constant__ Taps taps)
{
    float2 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    float2 x = make_float2(seed.x * 0.5f + threadIdx.x, seed.y * 0.25f);
    float2 hr[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) hr[i] = make_float2(taps.h[i] + (MODE == 7 ? 1e-9f * threadIdx.x : 0.f), taps.h[i]);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 5) acc[i] = __ffma2_rn(x, make_float2(taps.h[i], taps.h[i]), acc[i]);
            if (MODE == 6) { acc[i].x = fmaf(x.x, taps.h[i], acc[i].x); acc[i].y = fmaf(x.y, taps.h[i], acc[i].y); }
            if (MODE == 7) acc[i] = __ffma2_rn(x, hr[i], acc[i]);
        }
        x.x = -x.x;
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run_taps(const char *name, double flop_per_inner, double inst_per_inner)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8, threads = 256;
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * blocks * threads);
    Taps t;
    for (int i = 0; i < CH; ++i) t.h[i] = 0.001f * (i + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_taps<MODE><<<blocks, threads>>>(out, make_float2(1.0f, 0.5f), t);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) kern_taps<MODE><<<blocks, threads>>>(out, make_float2(1.0f, 0.5f), t);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)reps * blocks * threads * ITERS * CH;
    printf("{\"mode\": \"%s\", \"TFLOP/s\": %.2f, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f, \"ms\": %.3f, \"err\": \"%s\"}\n", name,
           n * flop_per_inner / (ms * 1e-3) / 1e12, n * inst_per_inner / 32.0 / (ms * 1e-3) / ((double)khz * 1e3) / sms,
           ms / reps, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

template <int MODE>
void run(const char *name, double flop_per_inner, double inst_per_inner)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8, threads = 256;
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<MODE><<<blocks, threads>>>(out, make_float2(1.0f, 0.5f));
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) kern<MODE><<<blocks, threads>>>(out, make_float2(1.0f, 0.5f));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)reps * blocks * threads * ITERS * CH;
    printf("{\"mode\": \"%s\", \"TFLOP/s\": %.2f, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f, \"ms\": %.3f, \"err\": \"%s\"}\n", name,
           n * flop_per_inner / (ms * 1e-3) / 1e12, n * inst_per_inner / 32.0 / (ms * 1e-3) / ((double)khz * 1e3) / sms,
           ms / reps, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    run<0>("2xFFMA scalar", 4, 2);
    run<1>("FFMA2 packed", 4, 1);
    run<2>("2xFADD scalar", 2, 2);
    run<3>("FADD2 packed", 2, 1);
    run<4>("FFMA2+FFMA", 6, 2);
    run_taps<5>("FFMA2 packed, tap broadcast from a uniform register (FIR form)", 4, 1);
    run_taps<6>("2xFFMA scalar, tap from the constant bank (FIR form)", 4, 2);
    run_taps<7>("FFMA2 packed, tap pair in per-thread registers (FIR form)", 4, 1);
    return 0;
}
