// fft_core.cuh -- register/shared-memory Stockham FFT building blocks (power-of-two sizes).
//
// Replaces the arithmetic of kf_work / kf_bfly2 / kf_bfly4 (libkissfft/kiss_fft.c:21-90,238-302):
// same unscaled DFT, forward exp(-j 2 pi k n / N), inverse exp(+j ...).  kissfft is a recursive
// out-of-place radix-4/2 DIT; here each thread owns E = min(16, N) points in registers, a pass is a
// radix-E butterfly done entirely in registers, and passes exchange data through padded shared
// memory in Stockham (auto-sort) order, so input and output are both in natural order and both use
// the coalesced layout "thread t holds x[t + e*T]", T = N/E threads per transform.
#pragma once
#include "common.cuh"

namespace lrfft {

// complex add/sub as ONE packed instruction each (FADD2, see common.cuh)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }

// forward: a * (-j);  inverse: a * (+j).  Pure lane swap + one negation: ptxas folds it into the operand
// modifier of the consuming packed instruction.
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a)
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void bfly2(float2 &a, float2 &b)
{
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// natural-order 4-point DFT in place: 8 packed adds
template <bool INV>
__device__ __forceinline__ void bfly4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    float2 s13 = cadd(a1, a3), d13 = rot90<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

// a * W16^k, W16 = exp(-+ 2 pi j / 16); k is a compile-time constant after unrolling.  Two packed
// instructions (FMUL2 + FFMA2) for the non-trivial k.
template <bool INV>
__device__ __forceinline__ float2 mul_w16(float2 a, int k)
{
    constexpr float C1 = 0.92387953251128673848f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508978178f;   // sin(pi/8)
    constexpr float H = 0.70710678118654752440f;    // cos(pi/4)
    k &= 15;
    if (k == 0) return a;
    if (k == 4) return rot90<INV>(a);
    if (k == 8) return make_float2(-a.x, -a.y);
    if (k == 12) return rot90<!INV>(a);
    float c, s;   // W = c - j s (forward)
    switch (k) {
        case 1:  c = C1;  s = S1;  break;
        case 2:  c = H;   s = H;   break;
        case 3:  c = S1;  s = C1;  break;
        case 5:  c = -S1; s = C1;  break;
        case 6:  c = -H;  s = H;   break;
        case 7:  c = -C1; s = S1;  break;
        case 9:  c = -C1; s = -S1; break;
        case 10: c = -H;  s = -H;  break;
        case 11: c = -S1; s = -C1; break;
        case 13: c = S1;  s = -C1; break;
        case 14: c = H;   s = -H;  break;
        default: c = C1;  s = -S1; break;   // 15
    }
    if (INV) s = -s;
    // (a.x + j a.y)(c - j s) = (a.x c + a.y s, a.y c - a.x s)
    return fma2(a, make_float2(c, c), mul2(make_float2(a.y, a.x), make_float2(s, -s)));
}

// a * W32^k, W32 = exp(-+ 2 pi j / 32); even k is a W16 power, odd k takes cos/sin from a literal table
// (k is a compile-time constant after unrolling, so the lookups fold to immediates)
template <bool INV>
__device__ __forceinline__ float2 mul_w32(float2 a, int k)
{
    k &= 31;
    if ((k & 1) == 0) return mul_w16<INV>(a, k >> 1);
    constexpr float C1 = 0.98078528040323043f, C3 = 0.83146961230254524f;   // cos(pi/16), cos(3 pi/16)
    constexpr float C5 = 0.55557023301960229f, C7 = 0.19509032201612833f;   // cos(5 pi/16), cos(7 pi/16)
    const float cq[4] = {C1, C3, C5, C7}, sq[4] = {C7, C5, C3, C1};        // first quadrant, odd k = 1, 3, 5, 7
    const int q = k >> 3, i = (k & 7) >> 1;
    float c = cq[i], s = sq[i];                                             // W = c - j s (forward)
    if (q == 1) { const float t = c; c = -s; s = t; }                       // angle + pi/2
    else if (q == 2) { c = -c; s = -s; }
    else if (q == 3) { const float t = c; c = s; s = -t; }
    if (INV) s = -s;
    return fma2(a, make_float2(c, c), mul2(make_float2(a.y, a.x), make_float2(s, -s)));
}

// natural-order R-point DFT of v[0..R) held in registers, R in {1, 2, 4, 8, 16, 32}
template <int R, bool INV>
struct RegFFT {
    static_assert(R == 8 || R == 16 || R == 32, "radix");
    __device__ __forceinline__ static void run(float2 *v)
    {
        constexpr int R2 = R / 4;    // R = 4 * R2 ; input index s = R2*a + b, output index r = c + 4*d
#pragma unroll
        for (int b = 0; b < R2; ++b) bfly4<INV>(v[b], v[R2 + b], v[2 * R2 + b], v[3 * R2 + b]);
#pragma unroll
        for (int c = 1; c < 4; ++c)
#pragma unroll
            for (int b = 1; b < R2; ++b) v[R2 * c + b] = mul_w32<INV>(v[R2 * c + b], b * c * (32 / R));
#pragma unroll
        for (int c = 0; c < 4; ++c) RegFFT<R2, INV>::run(v + R2 * c);
        float2 tmp[R];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int d = 0; d < R2; ++d) tmp[c + 4 * d] = v[R2 * c + d];
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = tmp[i];
    }
};
template <bool INV> struct RegFFT<1, INV> { __device__ __forceinline__ static void run(float2 *) {} };
template <bool INV> struct RegFFT<2, INV> {
    __device__ __forceinline__ static void run(float2 *v) { bfly2<INV>(v[0], v[1]); }
};
template <bool INV> struct RegFFT<4, INV> {
    __device__ __forceinline__ static void run(float2 *v) { bfly4<INV>(v[0], v[1], v[2], v[3]); }
};

// synchronisation among the T threads of one transform
struct SyncCta   { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct SyncNamed { int id, n; __device__ __forceinline__ void operator()() const { named_bar_sync(id, n); } };
struct SyncWarp  { unsigned mask; __device__ __forceinline__ void operator()() const { __syncwarp(mask); } };

template <int LOG2N, bool INV>
struct CtaFFT {
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = N >= 16 ? 16 : N;      // points per thread
    static constexpr int T = N / E;                  // threads per transform
    static constexpr int SMEM_CPX = N + (N >> 4);    // padded exchange buffer (float2 units)
    __device__ __forceinline__ static int pad(int i) { return i + (i >> 4); }

    // ---- twiddle bookkeeping: stage NS needs NB*(R-1) twiddles per thread, all fixed for a given t ----
    template <int NS>
    static constexpr int tw_count()
    {
        constexpr int REM = N / NS, R = REM >= E ? E : REM, NB = E / R;
        return NS > 1 ? NB * (R - 1) : 0;
    }
    template <int NS>
    static constexpr int tw_off()
    {
        if constexpr (NS <= 1) return 0;
        else return tw_off<NS / E>() + tw_count<NS / E>();      // every stage before the last has radix E
    }
    template <int NS>
    static constexpr int tw_total_from()
    {
        constexpr int REM = N / NS, R = REM >= E ? E : REM;
        if constexpr (NS * R == N) return tw_off<NS>() + tw_count<NS>();
        else return tw_total_from<NS * R>();
    }
    static constexpr int TW_REGS = tw_total_from<1>() > 0 ? tw_total_from<1>() : 1;   // 27 for N = 1024

    template <int NS>
    __device__ __forceinline__ static void load_stage_twiddles(const float2 *__restrict__ tw, int t, float2 *twr)
    {
        constexpr int REM = N / NS, R = REM >= E ? E : REM, NB = E / R;
        if constexpr (NS > 1) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int k = (t + q * T) & (NS - 1);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    float2 w = __ldg(&tw[r * k * (N / (NS * R))]);
                    if (INV) w.y = -w.y;
                    twr[tw_off<NS>() + q * (R - 1) + (r - 1)] = w;
                }
            }
        }
        if constexpr (NS * R != N) load_stage_twiddles<NS * R>(tw, t, twr);
    }
    // a persistent thread (fixed t) can fetch its twiddles once and keep them in registers
    __device__ __forceinline__ static void load_twiddles(const float2 *__restrict__ tw, int t, float2 *twr)
    {
        load_stage_twiddles<1>(tw, t, twr);
    }

    template <int NS, bool TWREG, class Sync>
    __device__ __forceinline__ static void stage(float2 *v, float2 *sm, const float2 *__restrict__ tw,
                                                 const float2 *twr, int t, const Sync &sync)
    {
        constexpr int REM = N / NS;
        constexpr int R = REM >= E ? E : REM;   // radix of this pass
        constexpr int NB = E / R;               // butterflies per thread
        constexpr bool LAST = (NS * R == N);
        if (NS > 1) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int k = (t + q * T) & (NS - 1);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    float2 w;
                    if constexpr (TWREG) {
                        w = twr[tw_off<NS>() + q * (R - 1) + (r - 1)];
                    } else {
                        w = __ldg(&tw[r * k * (N / (NS * R))]);
                        if (INV) w.y = -w.y;
                    }
                    v[q + r * NB] = cmulf(v[q + r * NB], w);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            float2 u[R];
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = v[q + r * NB];
            RegFFT<R, INV>::run(u);
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * NB] = u[r];
        }
        if constexpr (!LAST) {
            // not the last pass => R == E == 16, one butterfly per thread (j = t).  The padded addresses
            // are written as one base plus compile-time offsets (pad() is linear across multiples of 16),
            // so every access is a single LDS/STS with an immediate offset and no per-access integer math.
            const int base = (t / NS) * (NS * R) + (t & (NS - 1));
            float2 *wp = sm + pad(base);
            if constexpr (NS == 1) {
                // base = 16 t: pad(16 t + r) = pad(16 t) + r for r < 16
#pragma unroll
                for (int r = 0; r < R; ++r) wp[r] = v[r];
            } else {
                // NS is a multiple of 16: pad(base + r NS) = pad(base) + r (NS + NS/16)
#pragma unroll
                for (int r = 0; r < R; ++r) wp[r * (NS + NS / 16)] = v[r];
            }
            sync();
            if constexpr (T % 16 == 0) {
                const float2 *rp = sm + pad(t);          // pad(t + e T) = pad(t) + e (T + T/16)
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = rp[e * (T + T / 16)];
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = sm[pad(t + e * T)];
            }
            sync();
            stage<NS * R, TWREG, Sync>(v, sm, tw, twr, t, sync);
        }
    }

    // in: v[e] = x[t + e*T];  out: v[e] = X[t + e*T].  tw[k] = exp(-2 pi j k / N) (f64 -> f32 table).
    template <class Sync>
    __device__ __forceinline__ static void run(float2 *v, float2 *sm, const float2 *__restrict__ tw, int t,
                                               const Sync &sync)
    {
        stage<1, false, Sync>(v, sm, tw, nullptr, t, sync);
    }
    // same with the thread's twiddles already in registers (load_twiddles)
    template <class Sync>
    __device__ __forceinline__ static void run_twreg(float2 *v, float2 *sm, const float2 *twr, int t, const Sync &sync)
    {
        stage<1, true, Sync>(v, sm, nullptr, twr, t, sync);
    }
};

// ---------------------------------------------------------------------------------------------------------
// 1024-point transform by ONE warp (32 x 32): lane n2 holds x[n2 + 32 e], e < 32, in registers.
//   pass 1: DFT-32 over e in registers            A[n2][k1]
//   twiddle W1024^(n2 k1)                          (k1 = register index, n2 = lane)
//   ONE exchange through a 33-padded [32][32] shared tile (both directions conflict-free), __syncwarp only
//   pass 2: DFT-32 over n2 in registers            X[k1 + 32 k2], lane = k1, register = k2
// Input and output both use the coalesced layout "lane t holds element t + 32 e".  Against the CTA-level
// CtaFFT<10> (16 x 16 x 4, two exchanges, named barriers) this moves half the shared-memory bytes per point and
// has no barrier between warps: the 1024-point kernels are bound by shared-memory bandwidth and barrier
// stalls, not by FP32 issue.
// The 31 per-lane twiddles W^(lane r) live in a [31][32] shared table (row r - 1, column lane: conflict-free
// LDS.64) filled once per CTA -- 62 registers otherwise, or extra FP32 work if rebuilt from fewer stored powers,
// and this kernel is FP32-pipe bound once the exchanges are halved.
template <bool INV>
struct WarpFFT1024 {
    static constexpr int N = 1024;
    static constexpr int SMEM_CPX = 32 * 33;           // per-warp exchange tile
    static constexpr int TW_CPX = 31 * 32;             // per-CTA twiddle table

    // all threads of the CTA; caller synchronises afterwards
    __device__ __forceinline__ static void fill_twiddles(const float2 *__restrict__ tw, float2 *tws)
    {
        for (int i = threadIdx.x; i < TW_CPX; i += blockDim.x) {
            const int r = i / 32 + 1, lane = i % 32;
            float2 w = __ldg(&tw[r * lane]);             // 31 * 31 < 1024
            if (INV) w.y = -w.y;
            tws[i] = w;
        }
    }

    // v[e] = x[lane + 32 e] on entry, X[lane + 32 e] on exit.  xb: this warp's SMEM_CPX float2 tile.
    __device__ __forceinline__ static void run(float2 *v, float2 *xb, const float2 *tws, int lane)
    {
        RegFFT<32, INV>::run(v);
        const float2 *twl = tws + lane;
#pragma unroll
        for (int r = 1; r < 32; ++r) v[r] = cmulf(v[r], twl[32 * (r - 1)]);
        float2 *wp = xb + 33 * lane;                       // row n2 = lane, column k1 = r
        __syncwarp();                                      // previous frame's reads are done
#pragma unroll
        for (int r = 0; r < 32; ++r) wp[r] = v[r];
        __syncwarp();
        const float2 *rp = xb + lane;                      // column k1 = lane, rows n2 = e
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = rp[33 * e];
        RegFFT<32, INV>::run(v);
    }
};

}  // namespace lrfft
