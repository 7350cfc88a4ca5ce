// fir_core.cuh -- register-blocked polyphase FIR + decimate on a shared-memory tile.
//
// Replaces the arithmetic of dsputils::convolve (src/dsputils/src/dsputils.rs:30-32) on the re and
// im planes:  z[k] = sum_j x[k*DECIM + j] * taps[j]  (correlation, taps not reversed, j ascending).
//
// A thread owns R consecutive outputs.  Its input window is the contiguous run of
// (R-1)*DECIM + NTAPS samples starting at R*DECIM*t; the window is streamed once through 128-bit
// shared loads and every loaded sample feeds the <= ceil(NTAPS/DECIM) accumulators it belongs to, so
// a sample costs one quarter of an LDS.128 for up to 2*ceil(NTAPS/DECIM) FFMAs.  With R odd the
// per-thread stride R*DECIM*8 bytes is an odd multiple of 16 bytes, which makes the 8 threads of an
// LDS.128 phase hit 8 distinct 16-byte bank groups (conflict-free) without padding the TMA-written
// tile.  Taps are kernel parameters (constant bank, immediate offsets): FFMA reads them for free.
#pragma once
#include "common.cuh"

template <int NTAPS>
struct FirTaps { float h[NTAPS]; };

template <int NTAPS, int DECIM, int R>
struct FirTile {
    static constexpr int WIN = (R - 1) * DECIM + NTAPS;        // samples a thread reads
    static constexpr int STEP = R * DECIM;                      // samples between thread windows
    static_assert(WIN % 2 == 0, "window must be a whole number of 16-byte loads");
    static_assert((STEP * 8) % 16 == 0, "thread windows must start 16-byte aligned");

    // sx: this thread's window in shared memory (cf32), 16-byte aligned
    __device__ __forceinline__ static void run_cf32(const float2 *sx, const FirTaps<NTAPS> &taps, float2 *acc)
    {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < WIN; j += 2) {
            const float4 x = *reinterpret_cast<const float4 *>(sx + j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k0 = j - r * DECIM, k1 = k0 + 1;
                if (k0 >= 0 && k0 < NTAPS) {
                    acc[r].x = fmaf(x.x, taps.h[k0], acc[r].x);
                    acc[r].y = fmaf(x.y, taps.h[k0], acc[r].y);
                }
                if (k1 >= 0 && k1 < NTAPS) {
                    acc[r].x = fmaf(x.z, taps.h[k1], acc[r].x);
                    acc[r].y = fmaf(x.w, taps.h[k1], acc[r].y);
                }
            }
        }
    }

    // u8 IQ window (2 bytes per sample), 4-byte aligned.  The unpack i2f(b) = b/127 - 1 is folded:
    // sum_j h_j (b_j/127 - 1) = sum_j (h_j/127) (b_j - 127), and b - 127 is exact in f32, so the
    // caller passes taps already divided by 127 (in double, on the host).
    //
    // This path is FP32-issue bound, not HBM bound (2 B in per sample): the scalar form spent 896 FFMA + 248 FADD
    // + 248 PRMT + 62 LDS = 1454 issue slots per thread and tile.  Here the I and Q lanes of a sample go through
    // ONE packed instruction (FFMA2 for the tap product, FADD2 for the bias), the tap broadcast to both lanes:
    // 448 FFMA2 + 124 FADD2 + 248 PRMT + 62 LDS = 882 slots, which leaves the FMA pipe (1144 lane-cycles) as the
    // bound.  Each lane still performs exactly fma(x, h, acc) in ascending tap order: results are bit-identical
    // to the scalar form.
    __device__ __forceinline__ static void run_u8(const uint32_t *sw, const FirTaps<NTAPS> &taps127, float2 *acc)
    {
        run_u8_with(sw, taps127, acc, [](int) {});
    }

    // the same with a caller-supplied piece of independent work `side(j)` placed after the loads of step j (j = 0, 2, ...,
    // WIN - 2, a compile-time constant once unrolled): the fused receiver (k_fmrx.cu) threads the previous round's
    // discriminator through the filter this way, so its dependent chains issue between the packed FMAs.  The filter's own
    // operation order is untouched: results are bit-identical to run_u8.
    // TAPS_IN_REGS = false reads every tap where it is used: the compiler then brings it from the constant bank into a
    // UNIFORM register (LDCU) for the packed instruction, 64 fewer vector registers per thread for one LDCU.128 per
    // four taps -- what lets the fused receiver keep three CTAs per SM.
    template <class Side, bool TAPS_IN_REGS = true>
    __device__ __forceinline__ static void run_u8_with(const uint32_t *sw, const FirTaps<NTAPS> &taps127, float2 *acc, Side side)
    {
        // taps live in registers: a packed instruction takes its broadcast operand from a register, not from
        // the constant bank
        float h[TAPS_IN_REGS ? NTAPS : 1];
        if (TAPS_IN_REGS) {
#pragma unroll
            for (int k = 0; k < NTAPS; ++k) h[k] = taps127.h[k];
        }
        auto tap = [&](int k) { return TAPS_IN_REGS ? h[TAPS_IN_REGS ? k : 0] : taps127.h[k]; };
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
        const float2 bias = make_float2(8388735.0f, 8388735.0f);
#pragma unroll
        for (int j = 0; j < WIN; j += 2) {
            const uint32_t w = sw[j / 2];
            // 0x4B0000bb is the float 2^23 + bb; subtracting 2^23 + 127 gives bb - 127 exactly
            const float2 x0 = sub2(make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)),
                                               __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441))), bias);
            const float2 x1 = sub2(make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)),
                                               __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7443))), bias);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k0 = j - r * DECIM, k1 = k0 + 1;
                if (k0 >= 0 && k0 < NTAPS) acc[r] = fma2(x0, make_float2(tap(k0), tap(k0)), acc[r]);
                if (k1 >= 0 && k1 < NTAPS) acc[r] = fma2(x1, make_float2(tap(k1), tap(k1)), acc[r]);
            }
            side(j);
        }
    }
};
