// chain_generic.cuh -- the fused chain (k_chain.cu) for shapes other than the BASELINE one:
//   cf32 -> FIR(ntaps <= NTAPS)/DECIM (dsputils::convolve, dsputils.rs:30-32, + decimation)
//        -> frames of N = 2^LOG2N -> window -> FFT (kiss_fft.c:238-388) -> |X|^2 averaged (psdpng.c:165-177)
// one kernel, the decimated signal and the spectra never in HBM -- same warp-specialised layout as chain_kernel
// (FIR producer warps, FFT consumer warps, named barriers between them, mbarrier for the TMA), with three
// generalisations:
//
//  * ntaps is a RUN-TIME value <= NTAPS: the taps are zero-padded to NTAPS (a zero tap adds +-0 to the
//    accumulator: the sum is the ntaps-tap sum), the TMA brings exactly the (N-1) DECIM + ntaps samples
//    lrc_chain_frames promises, and the rest of the tile is zeroed once per CTA so the padded taps only ever meet zeros.
//
//  * the tile is written by the TMA in CHUNKS of STEP = R DECIM samples (one thread window advance), chunk c at byte
//    c (8 STEP + PADB): when 8 STEP is an even multiple of 16 bytes -- every power-of-two DECIM -- the thread windows of a
//    dense tile would start in the same one or two bank groups and the FIR's LDS.128 would run 4 to 8 ways conflicted
//    (the BASELINE shape escapes with R = 7, DECIM = 10: 560 bytes = 35 x 16); PADB = 16 makes the window pitch an odd
//    multiple of 16 bytes again.  A thread's sample j sits at 8 j + PADB (j / STEP) from its window start: compile-time
//    offsets after unrolling.  FIR thread t issues the bulk copy of chunk t, so the 70-300 copies of a tile are issued
//    in parallel, all completing on one mbarrier.
//
//  * N in {512, 1024, 2048}; shapes whose tile does not fit 227 KB (N DECIM > ~25 k samples) have no instance and run
//    unfused (FIR kernel -> HBM -> PSD kernel).
#pragma once
#include "fir_core.cuh"
#include "fft_core.cuh"

namespace chaing {
using namespace lrfft;

template <int NTAPS, int DECIM, int R>
struct GenTile {
    static constexpr int STEP = R * DECIM;
    static_assert(STEP % 2 == 0, "thread windows must start 16-byte aligned");
    static constexpr int WIN = (R - 1) * DECIM + NTAPS;
    static constexpr int WINL = (WIN + 1) & ~1;                       // loaded as pairs; a pad sample meets no tap
    static constexpr int PADB = ((STEP / 2) % 2 == 0) ? 16 : 0;
    static constexpr int PITCH = STEP * 8 + PADB;                      // bytes between thread windows = between chunks
    __host__ __device__ static constexpr int off(int j) { return j * 8 + PADB * (j / STEP); }

    __device__ __forceinline__ static void run(const uint8_t *wb, const FirTaps<NTAPS> &taps, float2 *acc)
    {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < WINL; j += 2) {
            const float4 x = *reinterpret_cast<const float4 *>(wb + off(j));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k0 = j - r * DECIM, k1 = k0 + 1;
                // one packed FFMA2 per sample and tap (re and im lanes, the tap broadcast from a uniform register): the
                // shapes with a small decimation are FP32-issue bound, not HBM bound
                if (k0 >= 0 && k0 < NTAPS) acc[r] = fma2(make_float2(x.x, x.y), make_float2(taps.h[k0], taps.h[k0]), acc[r]);
                if (k1 >= 0 && k1 < NTAPS) acc[r] = fma2(make_float2(x.z, x.w), make_float2(taps.h[k1], taps.h[k1]), acc[r]);
            }
        }
    }
};

template <int NTAPS, int DECIM, int LOG2N, int R>
struct GenCfg {
    using Tile = GenTile<NTAPS, DECIM, R>;
    using FFT = CtaFFT<LOG2N, false>;
    static constexpr int N = 1 << LOG2N;
    static constexpr int FRAME_ADV = N * DECIM;
    static constexpr int TILE_IN_MAX = (N - 1) * DECIM + NTAPS;
    static constexpr int NFIR = (N + R - 1) / R;
    static constexpr int NFIR_T = (NFIR + 31) / 32 * 32;
    static constexpr int NFFT_T = FFT::T;
    static constexpr int NT = NFIR_T + NFFT_T;
    static constexpr int N_CHUNKS_MAX = (TILE_IN_MAX + Tile::STEP - 1) / Tile::STEP;
    static constexpr int WIN_END = (NFIR - 1) * Tile::PITCH + Tile::off(Tile::WINL - 1) + 16;   // last byte a FIR thread reads
    static constexpr int CHUNK_END = N_CHUNKS_MAX * Tile::PITCH;
    static constexpr int TILE_BYTES = ((WIN_END > CHUNK_END ? WIN_END : CHUNK_END) + 127) / 128 * 128;
    static constexpr int OFF_HAND = TILE_BYTES;
    static constexpr int OFF_XCHG = OFF_HAND + FFT::SMEM_CPX * 8;
    static constexpr int OFF_BAR = OFF_XCHG + FFT::SMEM_CPX * 8;
    static constexpr int SMEM_BYTES = OFF_BAR + 16;
    static constexpr bool FITS = SMEM_BYTES <= 227 * 1024 && NT <= 1024;
    // two CTAs per SM whenever shared memory allows it (the register budget then follows from the launch bounds): the TMA
    // of one CTA overlaps the arithmetic of the other, as in the BASELINE instance
    // (at 128 registers per thread, which the 16-point-per-thread FFT needs)
    static constexpr int BY_SMEM = (227 * 1024) / (SMEM_BYTES + 1024), BY_REGS = 65536 / (NT * 128);
    static constexpr int MIN_CTAS_RAW = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
    static constexpr int MIN_CTAS = MIN_CTAS_RAW < 1 ? 1 : (MIN_CTAS_RAW > 8 ? 8 : MIN_CTAS_RAW);
    static_assert((FRAME_ADV * 8) % 16 == 0 && (Tile::STEP * 8) % 16 == 0, "TMA alignment");
    static_assert(NFFT_T % 32 == 0, "FFT threads must be whole warps");
};

enum { BAR_FULL = 1, BAR_EMPTY = 2, BAR_TILE = 3, BAR_FFT = 4 };

template <int NTAPS, int DECIM, int LOG2N, int R>
__global__ void __launch_bounds__(GenCfg<NTAPS, DECIM, LOG2N, R>::NT, GenCfg<NTAPS, DECIM, LOG2N, R>::MIN_CTAS)
chain_gen_kernel(const float2 *__restrict__ in, const float2 *__restrict__ tw, const float *__restrict__ win,
                 float *__restrict__ partial, size_t k_avg, size_t fpi, size_t ipr, size_t n_items, int ntaps_rt,
                 const __grid_constant__ FirTaps<NTAPS> taps)
{
    using Cfg = GenCfg<NTAPS, DECIM, LOG2N, R>;
    using Tile = typename Cfg::Tile;
    using FFT = typename Cfg::FFT;
    constexpr int N = Cfg::N, E = FFT::E, T = FFT::T;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *tile = smem;
    float2 *hand = reinterpret_cast<float2 *>(smem + Cfg::OFF_HAND);
    float2 *xchg = reinterpret_cast<float2 *>(smem + Cfg::OFF_XCHG);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + Cfg::OFF_BAR);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // the padded taps (ntaps_rt..NTAPS-1 are zero) must only ever meet finite values: zero the tile once
    for (int i = tid; i < Cfg::TILE_BYTES / 16; i += Cfg::NT) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();                                   // ... before the TMA (async proxy) writes over it
    __syncthreads();
    if ((size_t)blockIdx.x >= n_items) return;

    auto item_range = [&](size_t item, size_t &f0, size_t &f1) {
        const size_t row = item / ipr, c = item % ipr;
        f0 = row * k_avg + c * fpi;
        f1 = f0 + fpi;
        if (f1 > (row + 1) * k_avg) f1 = (row + 1) * k_avg;
    };

    if (tid < Cfg::NFIR_T) {
        // ================= FIR producers =================
        const int tile_in = (N - 1) * DECIM + ntaps_rt;                 // samples of a frame that exist in the input
        const int n_chunks = (tile_in + Tile::STEP - 1) / Tile::STEP;
        const int tma_samples = tile_in & ~1;                           // bulk copies move whole 16-byte units ...
        uint32_t phase = 0;
        bool first = true;
        size_t item = blockIdx.x, f0, f1;
        item_range(item, f0, f1);
        size_t f = f0;
        auto load_frame = [&](size_t fr) {                              // every producer thread: its chunk(s)
            const float2 *src = in + fr * (size_t)Cfg::FRAME_ADV;
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)tma_samples * 8u);
            for (int c = tid; c < n_chunks; c += Cfg::NFIR_T) {
                const int s0 = c * Tile::STEP;
                int ns = tma_samples - s0;
                if (ns > Tile::STEP) ns = Tile::STEP;
                if (ns > 0) tma_load_1d_evict_first(tile + (size_t)c * Tile::PITCH, src + s0, (uint32_t)ns * 8u, bar);
                // ... and an odd last sample goes by hand (its owner stores it before it arrives on BAR_TILE)
                if ((tile_in & 1) && s0 + Tile::STEP >= tile_in && s0 < tile_in)
                    *reinterpret_cast<float2 *>(tile + (size_t)c * Tile::PITCH + (size_t)(tile_in - 1 - s0) * 8) = __ldg(src + tile_in - 1);
            }
        };
        load_frame(f);
        // the hand-copied sample must be visible to the thread whose window holds it
        named_bar_sync(BAR_TILE, Cfg::NFIR_T);
        while (true) {
            mbar_wait(bar, phase);
            phase ^= 1;
            float2 acc[R];
            if (tid < Cfg::NFIR) Tile::run(tile + (size_t)tid * Tile::PITCH, taps, acc);
            size_t nf = f + 1, nitem = item, nf1 = f1;
            if (nf == f1) {
                nitem = item + gridDim.x;
                if (nitem < n_items) item_range(nitem, nf, nf1);
            }
            const bool has_next = nitem < n_items;
            // every FIR thread is done with the tile -> the chunks of the next frame
            named_bar_sync(BAR_TILE, Cfg::NFIR_T);
            if (has_next) load_frame(nf);
            if (!first) named_bar_sync(BAR_EMPTY, Cfg::NT);
            first = false;
            if (tid < Cfg::NFIR) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int o = tid * R + r;
                    if (o < N) hand[FFT::pad(o)] = acc[r];
                }
            }
            __threadfence_block();
            named_bar_arrive(BAR_FULL, Cfg::NT);
            if (!has_next) break;
            if (tile_in & 1) named_bar_sync(BAR_TILE, Cfg::NFIR_T);    // the hand-copied last sample (see load_frame)
            item = nitem; f = nf; f1 = nf1;
        }
    } else {
        // ================= FFT consumers =================
        const int t = tid - Cfg::NFIR_T;
        float w[E], acc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { w[e] = win[t + e * T]; acc[e] = 0.f; }
        size_t item = blockIdx.x, f0, f1;
        item_range(item, f0, f1);
        size_t f = f0;
        while (true) {
            size_t nf = f + 1, nitem = item, nf1 = f1;
            if (nf == f1) {
                nitem = item + gridDim.x;
                if (nitem < n_items) item_range(nitem, nf, nf1);
            }
            const bool has_next = nitem < n_items;
            named_bar_sync(BAR_FULL, Cfg::NT);
            float2 v[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float2 x = hand[FFT::pad(t + e * T)];
                v[e] = make_float2(x.x * w[e], x.y * w[e]);
            }
            if (has_next) named_bar_arrive(BAR_EMPTY, Cfg::NT);
            FFT::run(v, xchg, tw, t, SyncNamed{BAR_FFT, T});
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(v[e].x, v[e].x, fmaf(v[e].y, v[e].y, acc[e]));
            if (f + 1 == f1) {
                float *dst = partial + item * N + t;
#pragma unroll
                for (int e = 0; e < E; ++e) { dst[e * T] = acc[e]; acc[e] = 0.f; }
            }
            if (!has_next) break;
            item = nitem; f = nf; f1 = nf1;
        }
    }
}

// what a launch needs besides the template arguments
struct Args {
    const float2 *in, *tw;
    const float *win, *taps;     // taps: host pointer, ntaps values
    float *partial;
    size_t k_avg, fpi, ipr, n_items;
    int ntaps, n_sm;
    cudaStream_t stream;
};

// LRC_OK = launched, -1 = no instance for this shape (caller runs unfused), > 0 = LRC error
template <int NTAPS, int DECIM, int LOG2N, int R>
static int launch_one(const Args &a)
{
    using Cfg = GenCfg<NTAPS, DECIM, LOG2N, R>;
    if constexpr (!Cfg::FITS) {
        return -1;
    } else {
        auto kern = chain_gen_kernel<NTAPS, DECIM, LOG2N, R>;
        LRC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int occ = 1;
        LRC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::NT, Cfg::SMEM_BYTES));
        if (occ < 1) occ = 1;
        size_t blocks = (size_t)a.n_sm * occ;
        if (blocks > a.n_items) blocks = a.n_items;
        FirTaps<NTAPS> taps;
        for (int i = 0; i < NTAPS; ++i) taps.h[i] = i < a.ntaps ? a.taps[i] : 0.0f;
        kern<<<(unsigned)blocks, Cfg::NT, Cfg::SMEM_BYTES, a.stream>>>(a.in, a.tw, a.win, a.partial, a.k_avg, a.fpi, a.ipr,
                                                                      a.n_items, a.ntaps, taps);
        LRC_CUDA(cudaGetLastError());
        return 0;
    }
}

// R per decimation: window advance R DECIM of 28..64 samples (enough producer threads, 2-3x shared-memory read amplification)
template <int DECIM> struct RFor;
template <> struct RFor<4>  { static constexpr int R = 8; };
template <> struct RFor<5>  { static constexpr int R = 8; };
template <> struct RFor<8>  { static constexpr int R = 7; };
template <> struct RFor<10> { static constexpr int R = 7; };
template <> struct RFor<16> { static constexpr int R = 4; };

template <int DECIM>
static int launch_decim(const Args &a, int log2n)
{
    constexpr int R = RFor<DECIM>::R;
    if (a.ntaps < 1 || a.ntaps > 128) return -1;
    if (a.ntaps <= 64) {
        switch (log2n) {
            case 9:  return launch_one<64, DECIM, 9, R>(a);
            case 10: return launch_one<64, DECIM, 10, R>(a);
            case 11: return launch_one<64, DECIM, 11, R>(a);
            default: return -1;
        }
    }
    switch (log2n) {
        case 9:  return launch_one<128, DECIM, 9, R>(a);
        case 10: return launch_one<128, DECIM, 10, R>(a);
        case 11: return launch_one<128, DECIM, 11, R>(a);
        default: return -1;
    }
}

template <int DECIM>
static bool has_decim(int ntaps, int log2n)
{
    constexpr int R = RFor<DECIM>::R;
    if (ntaps < 1 || ntaps > 128 || log2n < 9 || log2n > 11) return false;
    if (ntaps <= 64) return log2n == 9 ? GenCfg<64, DECIM, 9, R>::FITS : log2n == 10 ? GenCfg<64, DECIM, 10, R>::FITS : GenCfg<64, DECIM, 11, R>::FITS;
    return log2n == 9 ? GenCfg<128, DECIM, 9, R>::FITS : log2n == 10 ? GenCfg<128, DECIM, 10, R>::FITS : GenCfg<128, DECIM, 11, R>::FITS;
}

}  // namespace chaing

// one translation unit per decimation (k_chaing_d*.cu), so the 28 instances compile in parallel
int  lrc_chaing_launch_d4(const chaing::Args &a, int log2n);
int  lrc_chaing_launch_d5(const chaing::Args &a, int log2n);
int  lrc_chaing_launch_d8(const chaing::Args &a, int log2n);
int  lrc_chaing_launch_d10(const chaing::Args &a, int log2n);
int  lrc_chaing_launch_d16(const chaing::Args &a, int log2n);
bool lrc_chaing_has_d4(int ntaps, int log2n);
bool lrc_chaing_has_d5(int ntaps, int log2n);
bool lrc_chaing_has_d8(int ntaps, int log2n);
bool lrc_chaing_has_d10(int ntaps, int log2n);
bool lrc_chaing_has_d16(int ntaps, int log2n);
