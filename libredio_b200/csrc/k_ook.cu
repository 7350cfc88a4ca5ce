// k_ook.cu -- 433 MHz OOK packet decode, bit-exact with the reference chain of src/ratpak.rs:60-111:
//
//   rtlsdr::data_to_samples (rtlsdr.rs:160)  -> |x| = hypot (ratpak.rs:64-68)
//   -> bitfount::trigger (bitfount.rs:36-85)  -> bitfount::discretize (:87-96)
//   -> kpn::rle (kpn.rs:17-29) -> kpn::dle (:32-38) -> pulse-pair matchers (ratpak.rs:88-97)
//   -> kpn::shaper_optional 36 / 24 (kpn.rs:266-275)
//
// The reference runs this as eight threads exchanging one message per SAMPLE.  Here it is four kernels
// over all streams at once; every float operation that feeds a comparison is the same IEEE operation in
// the same order as the reference (explicit __f*_rn / __d*_rn intrinsics, never contracted):
//
//   K-A ook_block_kernel : per 512-sample block, envelope of every sample, the strictly sequential f32
//                          block sum `s` (bitfount.rs:48) and the block max (order-free, exact)
//   K-B ook_trigger_kernel: one thread per stream walks its blocks through the trigger state machine
//                          (:46-81), tags every block with the burst it is collected into, and keeps
//                          max/2 per burst (discretize :90-91)
//   K-C ook_rle_kernel   : one warp per stream re-derives the envelope of collected blocks, slices it
//                          against the burst's max/2 into bit masks and emits the positions where the
//                          continuous bit stream changes value (rle: runs span burst boundaries, the last
//                          run is never flushed)
//   K-D ook_match_kernel : one thread per stream: run lengths -> seconds (dle, IEEE f32 division) ->
//                          matcher A and B -> shaper_optional -> packed packets
#include "common.cuh"
#include "unpack.cuh"
#include <algorithm>
#include <vector>

static const int OOK_BLOCK = 512;            // bitfount.rs:38
static const int OOK_TRIGGER_DURATION = 50;  // bitfount.rs:40

// |i2f(b0) + j i2f(b1)| = (float)sqrt((double)re*re + (double)im*im)  -- SURVEY 8c definition of
// num::Complex::norm (hypot).  Both products are exact in f64, so this is one rounded add, one
// correctly rounded sqrt and one narrowing, exactly as oracle/restated.c orc_norm.
__device__ __forceinline__ float lr_envelope(uint32_t b0, uint32_t b1)
{
    const double re = (double)lr_i2f(b0), im = (double)lr_i2f(b1);
    const double s = __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
    return __double2float_rn(__dsqrt_rn(s));
}

// The envelope depends on the two bytes only and is symmetric in them (the f64 add commutes), so the hot
// kernels read it from a triangular table (32896 entries, rows padded: 139.7 KB) held in shared memory.
// The table is filled ON THE DEVICE by lr_envelope above when the plan is created; the u8 domain being
// finite, table == formula for all 65536 pairs is a proof of equivalence, checked exhaustively by
// tests/test_gpu_ook_fastfir.py::test_envelope_exhaustive_65536_pairs_bit_exact (formula vs CPU) and by the
// bit-exact block sums of every OOK test (table vs CPU).
// Row hi of the triangle starts at hi (hi + 17) / 2 = tri(hi) + 8 hi: eight unused words per row, so that
// consecutive rows start ~8 banks apart.  (With the plain tri(hi) the rows next to hi = 127 -- where a noise
// floor lives -- start 0 or 1 bank apart and a warp's 32 lookups pile onto a handful of banks.)
constexpr int OOK_LUT_N = 255 * (255 + 17) / 2 + 256;         // 34936 floats
constexpr int OOK_LUT_BYTES = OOK_LUT_N * 4;
static_assert(OOK_LUT_N % 4 == 0, "table is copied as float4");

__device__ __forceinline__ uint32_t lut_index(uint32_t hi, uint32_t lo) { return ((hi * (hi + 17u)) >> 1) + lo; }

__device__ __forceinline__ float lut_envelope(const float *lut, uint32_t b0, uint32_t b1)
{
    return lut[lut_index(max(b0, b1), min(b0, b1))];
}

// The block-sum kernel is bound by the ALU pipe (byte extraction, max/min, index and address arithmetic run
// there at half the FMA pipe's rate), so its lookup is written to keep that pipe short: one PRMT per byte,
// max, min, and the byte address  lut + 4 (hi (hi + 17) / 2 + lo)  =  lut + hi (2 hi + 34) + 4 lo  as two
// integer multiply-adds (FMA pipe) and one scaled add.
__device__ __forceinline__ float lut_envelope_s(uint32_t lut_s, uint32_t b0, uint32_t b1)
{
    const uint32_t hi = max(b0, b1), lo = min(b0, b1);
    uint32_t a, b;
    asm("mad.lo.u32 %0, %1, 2, 34;" : "=r"(a) : "r"(hi));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(hi), "r"(a), "r"(lut_s));
    float e;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e) : "r"(b + 4u * lo));
    return e;
}

__device__ __forceinline__ void lut_load(float *s_lut, const float *__restrict__ g_lut)
{
    for (int i = threadIdx.x; i < OOK_LUT_N / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(s_lut)[i] = __ldg(reinterpret_cast<const float4 *>(g_lut) + i);
    __syncthreads();
}

__global__ void ook_build_lut_kernel(float *__restrict__ lut)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536u) return;
    const uint32_t hi = i >> 8, lo = i & 0xffu;
    if (lo <= hi) lut[lut_index(hi, lo)] = lr_envelope(hi, lo);
}

struct lrc_ook {
    lrc_ctx *ctx;
    size_t   n_streams, n_blocks, max_runs, max_packets, max_bursts;
    unsigned sample_rate;
    float    *d_sum, *d_max;          // [n_streams][n_blocks]
    int32_t  *d_tag;                  // [n_streams][n_blocks] burst index the block is collected into, -1 = none
    float    *d_half;                 // [n_streams][max_bursts]  max/2 of the burst
    uint8_t  *d_bflags;               // [n_streams][max_bursts]  bit0 = emitted, bit1 = leading 0.0 sample
    uint32_t *d_nbursts;              // [n_streams]
    uint32_t *d_trans;                // [n_streams][max_runs] positions where the bit stream changes value
    uint32_t *d_ntrans;               // [n_streams]  (may exceed max_runs: overflow is detected on fetch)
    uint32_t *d_nbits;                // [n_streams]  length of the flattened bit stream
    unsigned long long *d_packets;    // [n_streams][2][max_packets] packets packed MSB-first
    uint32_t *d_npackets;             // [n_streams][2]
    uint32_t *d_runs_dbg;             // [n_streams][max_runs] (value << 31 | length), filled by K-D
    float    *d_lut;                  // triangular envelope table, OOK_LUT_N floats
    uint16_t *d_rank;                 // [65536] rank of the pair's envelope among the distinct envelope values
    float    *d_uniq;                 // [n_uniq] the distinct envelope values, ascending
    uint32_t  n_uniq;
};

// ---------------------------------------------------------------------------------------------
// K-A: envelope, sequential block sum, block max.  One warp handles 32 consecutive blocks of one
// stream and lane b owns block b: its 512 envelopes must be added one after the other (bitfount.rs:48),
// so the lane walks its own 1024 bytes.  The bytes reach it through shared memory: the warp copies
// 64-byte slabs of its 32 block rows with cp.async (16 bytes per lane, 4 lanes per row: coalesced, no
// registers, the next slab in flight while this one is summed) into rows of pitch 80 bytes, which lane b
// then reads back with LDS.128 -- 5 x 16 bytes between lanes, so the eight lanes of a quarter-warp phase
// hit eight different bank groups.  No envelope is ever written back to shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int KA_WARPS = 16;
constexpr int KA_SLAB = 32;                          // samples per block row per slab (64 bytes)
constexpr int KA_PITCH = KA_SLAB * 2 + 16;           // bytes between block rows in the staging tile
constexpr int KA_STAGE_BYTES = 32 * KA_PITCH;
constexpr int KA_STAGES = 2;
constexpr int KA_SMEM_BYTES = OOK_LUT_BYTES + KA_WARPS * KA_STAGES * KA_STAGE_BYTES;

__global__ void __launch_bounds__(KA_WARPS * 32, 1)
ook_block_kernel(const uint8_t *__restrict__ iq, size_t stream_stride, size_t n_streams, size_t n_blocks,
                 const float *__restrict__ g_lut, float *__restrict__ d_sum, float *__restrict__ d_max)
{
    extern __shared__ __align__(16) float ka_smem[];
    float *lut = ka_smem;
    lut_load(lut, g_lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stg = reinterpret_cast<uint8_t *>(ka_smem + OOK_LUT_N) + warp * (KA_STAGES * KA_STAGE_BYTES);
    const uint32_t stg_s = smem_u32(stg);
    const uint32_t lut_s = smem_u32(lut);
    const size_t groups_per_stream = (n_blocks + 31) / 32;
    const size_t n_groups = groups_per_stream * n_streams;
    const size_t warps_total = (size_t)gridDim.x * KA_WARPS;
    constexpr int N_SLABS = OOK_BLOCK / KA_SLAB;
    for (size_t grp = (size_t)blockIdx.x * KA_WARPS + warp; grp < n_groups; grp += warps_total) {
        const size_t st = grp / groups_per_stream, b0 = (grp % groups_per_stream) * 32;
        const int nb = (int)((n_blocks - b0) < 32 ? (n_blocks - b0) : 32);
        const uint8_t *base = iq + st * stream_stride + b0 * (size_t)(OOK_BLOCK * 2);
        auto issue = [&](int slab) {
            const uint32_t dst0 = stg_s + (slab & 1) * KA_STAGE_BYTES;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + (lane >> 2), col = lane & 3;
                if (row < nb)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                 :: "r"(dst0 + row * KA_PITCH + col * 16),
                                    "l"(base + (size_t)row * (OOK_BLOCK * 2) + slab * (KA_SLAB * 2) + col * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float s = 0.0f, mx = 0.0f;
        issue(0);
        for (int slab = 0; slab < N_SLABS; ++slab) {
            if (slab + 1 < N_SLABS) {
                issue(slab + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();                                   // every lane's copies of this slab have landed
            if (lane < nb) {
                const uint4 *r = reinterpret_cast<const uint4 *>(stg + (slab & 1) * KA_STAGE_BYTES + lane * KA_PITCH);
#pragma unroll
                for (int q = 0; q < KA_SLAB * 2 / 16; ++q) {
                    const uint4 v = r[q];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float e0 = lut_envelope_s(lut_s, __byte_perm(w[k], 0, 0x4440), __byte_perm(w[k], 0, 0x4441));
                        const float e1 = lut_envelope_s(lut_s, __byte_perm(w[k], 0, 0x4442), __byte_perm(w[k], 0, 0x4443));
                        s = __fadd_rn(s, e0);                // samples.iter().sum(): left to right from 0.0
                        s = __fadd_rn(s, e1);
                        mx = fmaxf(mx, fmaxf(e0, e1));
                    }
                }
            }
            __syncwarp();                                   // the stage is rewritten two slabs from now
        }
        if (lane < nb) {
            d_sum[st * n_blocks + b0 + lane] = s;
            d_max[st * n_blocks + b0 + lane] = mx;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K-B: trigger state machine, one thread per stream (bitfount.rs:41-81, statement by statement)
// ---------------------------------------------------------------------------------------------
constexpr int KB_THREADS = 64;                // streams per CTA (static shared memory stays under 48 KB)
constexpr int KB_TILE = 32;                   // blocks per staged tile
constexpr int KB_LD = KB_TILE + 1;

__global__ void __launch_bounds__(KB_THREADS)
ook_trigger_kernel(const float *__restrict__ d_sum, const float *__restrict__ d_max, size_t n_streams,
                   size_t n_blocks, size_t max_bursts, int32_t *__restrict__ d_tag, float *__restrict__ d_half,
                   uint8_t *__restrict__ d_bflags, uint32_t *__restrict__ d_nbursts)
{
    // The state machine is sequential per stream (one thread each); its inputs are not: the CTA stages
    // tiles of [128 streams][32 blocks] sums and maxima through shared memory with coalesced row reads
    // (a warp reads 32 consecutive floats of one stream), double-buffered so the next tile is in flight
    // while this one is walked, and the tags leave the same way.
    __shared__ float s_sum[2][KB_THREADS * KB_LD];
    __shared__ float s_max[2][KB_THREADS * KB_LD];
    __shared__ int32_t s_tag[KB_THREADS * KB_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t st0 = (size_t)blockIdx.x * KB_THREADS;
    const size_t st = st0 + tid;
    const bool live = st < n_streams;
    int32_t *tag = d_tag + st * n_blocks;
    float *half = d_half + st * max_bursts;
    uint8_t *flags = d_bflags + st * max_bursts;
    int trigger = 0;                          // :41 (isize there; |trigger| <= n_blocks here)
    float threshold = 0.0f;                   // :44
    uint32_t buf_len = 1;                     // :43 sample_buffer = vec!(0.0); capped at 25.6 M + 512 by the guard below
    bool lead0 = true;                        // the buffer currently starts with that 0.0
    float cur_max = 0.0f;
    uint32_t burst = 0;                       // index of the burst being collected
    size_t burst_first_block = 0;
    bool burst_has_blocks = false;
    const size_t n_tiles = (n_blocks + KB_TILE - 1) / KB_TILE;
    // cp.async (LDGSTS) writes the staged tile straight into shared memory: the loads of tile t+1 are in
    // flight while the state machine walks tile t, and nothing waits on a register
    auto stage = [&](size_t tile, int buf) {
        const size_t b = tile * KB_TILE + lane;
#pragma unroll 4
        for (int r = warp; r < KB_THREADS; r += KB_THREADS / 32) {
            const size_t s = st0 + r;
            float *ds = &s_sum[buf][r * KB_LD + lane], *dm = &s_max[buf][r * KB_LD + lane];
            if (s < n_streams && b < n_blocks) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(ds)), "l"(d_sum + s * n_blocks + b) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dm)), "l"(d_max + s * n_blocks + b) : "memory");
            } else {
                *ds = 0.0f; *dm = 0.0f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);
    for (size_t tile = 0; tile < n_tiles; ++tile) {
        const int buf = (int)(tile & 1);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                                   // tile `tile` is in shared memory for every thread
        if (tile + 1 < n_tiles) stage(tile + 1, buf ^ 1);  // buffer buf^1 was last read two barriers ago
        const size_t b0 = tile * KB_TILE;
        const int nb = (int)((n_blocks - b0) < (size_t)KB_TILE ? (n_blocks - b0) : (size_t)KB_TILE);
        if (live) {
            // s / 1000 (:63) for the whole tile first: 32 independent IEEE divisions pipeline, inside the
            // state machine each would sit on the threshold's dependency chain.  s_tag doubles as the buffer.
            float *s_q = reinterpret_cast<float *>(s_tag) + tid * KB_LD;
#pragma unroll 8
            for (int u = 0; u < KB_TILE; ++u) s_q[u] = __fdiv_rn(s_sum[buf][tid * KB_LD + u], 1000.0f);
            for (int u = 0; u < nb; ++u) {
                const size_t b = b0 + u;
                trigger -= 1;                                                       // :46
                const float s = s_sum[buf][tid * KB_LD + u];                       // :48
                const float s_over_1000 = s_q[u];
                // :52-54 OOM guard (a burst longer than 50 000 blocks = 100 s at 256 ksps; no fixture reaches it).  KNOWN
                // DEVIATION in one corner: when the guard fires with trigger == 1, the reference sends the reset buffer
                // [0.0] at the next block (a burst of one 0 bit); here that burst has no tagged block and contributes no
                // bit, so transition positions after it are one lower than the reference's.  DESIGN.md section 7.
                if (buf_len > 1000u * OOK_TRIGGER_DURATION * OOK_BLOCK) {
                    if (burst_has_blocks) {
                        // blocks of this burst tagged in earlier tiles are already in global memory
                        for (size_t k = burst_first_block; k < b0; ++k) if (tag[k] == (int32_t)burst) tag[k] = -1;
                        for (int k = 0; k < u; ++k) if (s_tag[tid * KB_LD + k] == (int32_t)burst) s_tag[tid * KB_LD + k] = -1;
                    }
                    buf_len = 1; lead0 = true; cur_max = 0.0f; burst_has_blocks = false;
                }
                if (threshold == 0.0f) threshold = s;                               // :57-59
                if (trigger < 0) {                                                  // :62-65
                    threshold = __fadd_rn(threshold, s_over_1000);
                    threshold = __fsub_rn(threshold, __fmul_rn(threshold, 0.002f));
                }
                if (s > __fmul_rn(threshold, 4.0f)) trigger = OOK_TRIGGER_DURATION; // :68-70
                int32_t tg = -1;
                if (trigger > 1) {                                                  // :73-75 push_all
                    if (burst < max_bursts) {
                        tg = (int32_t)burst;
                        if (!burst_has_blocks) { burst_first_block = b; burst_has_blocks = true; }
                    }
                    buf_len += OOK_BLOCK;
                    cur_max = fmaxf(cur_max, s_max[buf][tid * KB_LD + u]);
                }
                s_tag[tid * KB_LD + u] = tg;
                if (trigger == 0) {                                                 // :78-81 send, buffer = vec!()
                    if (burst < max_bursts) {
                        const float h = __fdiv_rn(cur_max, 2.0f);                   // discretize :90-91 max/2f32
                        half[burst] = h;
                        flags[burst] = (uint8_t)(1u | (lead0 ? 2u : 0u));
                    }
                    burst += 1;
                    buf_len = 0; lead0 = false; cur_max = 0.0f; burst_has_blocks = false;
                }
            }
        }
        __syncthreads();
        // tags of this tile out, coalesced
        {
            const size_t b = b0 + lane;
            for (int r = warp; r < KB_THREADS; r += KB_THREADS / 32) {
                const size_t s = st0 + r;
                if (s < n_streams && b < n_blocks) d_tag[s * n_blocks + b] = s_tag[r * KB_LD + lane];
            }
        }
        __syncthreads();
    }
    if (!live) return;
    // a burst still open when the capture ends is never sent: un-tag its blocks
    if (burst_has_blocks)
        for (size_t k = burst_first_block; k < n_blocks; ++k) if (tag[k] == (int32_t)burst) tag[k] = -1;
    d_nbursts[st] = burst;                    // may exceed max_bursts -> reported by fetch
}

// ---------------------------------------------------------------------------------------------
// K-C: discretize + rle as bit masks, one warp per stream.  Lane l owns samples [16 l, 16 l + 16) of a
// collected block (32 contiguous bytes).  The positions (in the flattened bit stream) where the value
// changes are appended to the stream's transition list in order.
//
// The slicer `x > max/2` (bitfount.rs:91) never needs the envelope VALUE here, only its order against the
// threshold: the plan holds rank[b0 | b1 << 8] = index of the pair's envelope among the 32 k distinct
// envelope values in ascending order (a 128 KB u16 table indexed by the raw byte pair -- no max/min, no
// triangular index); a burst's max/2 becomes the rank threshold #{values <= max/2} by a warp-wide 32-ary
// search of the sorted values when the walk enters the burst; then x > max/2  <=>  rank(x) >= threshold,
// exactly, for every byte pair.  Neighbouring byte pairs (a noise floor sits within a few codes of 127/127)
// would all fall into the banks of byte0 >> 1 with a pitch of 256 entries, so a byte1 row is 264 entries long:
// 132 words = 4 banks further per row, which spreads an 8 x 8 neighbourhood over all 32 banks.  Like the
// block-sum kernel this one is ALU-pipe bound, so the per-sample work is kept to two PRMT, one integer
// multiply-add (FMA pipe) and one scaled add for the address, one subtraction and one funnel shift that
// moves the comparison's sign bit into the mask (bits arrive reversed; one BREV per 16 samples).
// ---------------------------------------------------------------------------------------------
constexpr int OOK_RANK_PITCH = 264;                           // u16 entries per byte1 row (256 used)
constexpr int OOK_RANK_N = 256 * OOK_RANK_PITCH;
constexpr int OOK_RANK_BYTES = OOK_RANK_N * 2;                // 135 168
static_assert(OOK_RANK_BYTES % 16 == 0, "table is copied as uint4");
__host__ __device__ __forceinline__ uint32_t ook_rank_slot(uint32_t b0, uint32_t b1) { return b0 + OOK_RANK_PITCH * b1; }

// #{i : uniq[i] <= h} for ascending uniq[0..n): every lane probes one position per round
__device__ __forceinline__ uint32_t warp_upper_bound(const float *__restrict__ uniq, uint32_t n, float h, int lane)
{
    uint32_t lo = 0, hi = n;          // the answer is in [lo, hi]: uniq[i] <= h for i < lo, uniq[i] > h for i >= hi
    while (lo < hi) {
        const uint32_t step = (hi - lo + 31) / 32;
        const uint32_t p = lo + (lane + 1) * step - 1;
        const bool le = p < hi ? (__ldg(uniq + p) <= h) : false;
        const uint32_t c = __popc(__ballot_sync(0xffffffffu, le));
        const uint32_t nlo = lo + c * step;
        const uint32_t cap = lo + (c + 1) * step - 1;
        hi = cap < hi ? cap : hi;
        lo = nlo;
    }
    return lo;
}

constexpr int KC_THREADS = 1024;             // 32 streams per CTA, one CTA per SM: 4736 warps cover 4096 streams in ONE wave

__global__ void __launch_bounds__(KC_THREADS, 1)
ook_rle_kernel(const uint8_t *__restrict__ iq, size_t stream_stride, size_t n_streams, size_t n_blocks,
               size_t max_bursts, size_t max_runs, const uint16_t *__restrict__ g_rank,
               const int32_t *__restrict__ d_tag,
               const float *__restrict__ d_half, const float *__restrict__ uniq, uint32_t n_uniq,
               const uint8_t *__restrict__ d_bflags,
               uint32_t *__restrict__ d_trans, uint32_t *__restrict__ d_ntrans, uint32_t *__restrict__ d_nbits)
{
    extern __shared__ __align__(16) uint16_t kc_rank[];
    for (int i = threadIdx.x; i < OOK_RANK_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(kc_rank)[i] = __ldg(reinterpret_cast<const uint4 *>(g_rank) + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t rank_s = smem_u32(kc_rank);
    const size_t st = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (st >= n_streams) return;
    const int32_t *tag = d_tag + st * n_blocks;
    const float *half = d_half + st * max_bursts;
    const uint8_t *flags = d_bflags + st * max_bursts;
    uint32_t *trans = d_trans + st * max_runs;
    const uint8_t *base = iq + st * stream_stride;
    uint32_t pos = 0;            // length of the bit stream so far
    uint32_t ntr = 0;            // transitions so far
    uint32_t prev = 0;           // value of the last bit (meaningful once pos > 0)
    int32_t cur_burst = -1;
    uint32_t h = 0;              // rank threshold of the current burst
    // one collected block: 16 samples per lane -> bit mask -> transitions appended in order
    auto process = [&](int32_t tg, const uint4 &q0, const uint4 &q1) {
        if (tg != cur_burst) {
            cur_burst = tg;
            h = warp_upper_bound(uniq, n_uniq, half[tg], lane);
            if (flags[tg] & 2u) {
                // the burst starts with the literal 0.0 of vec!(0.0): 0.0 > max/2 is false -> bit 0
                if (pos > 0 && prev != 0u) { if (lane == 0 && ntr < max_runs) trans[ntr] = pos; ntr++; }
                prev = 0u; pos += 1;
            }
        }
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t m = 0;
        const uint32_t hm1 = h - 1u;                                         // rank >= h  <=>  (h - 1) - rank < 0
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t a0, a1;
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a0) : "r"(__byte_perm(w[i], 0, 0x4441)), "r"(2u * OOK_RANK_PITCH), "r"(rank_s));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a1) : "r"(__byte_perm(w[i], 0, 0x4443)), "r"(2u * OOK_RANK_PITCH), "r"(rank_s));
            unsigned short r0, r1;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r0) : "r"(a0 + 2u * __byte_perm(w[i], 0, 0x4440)));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r1) : "r"(a1 + 2u * __byte_perm(w[i], 0, 0x4442)));
            m = __funnelshift_l(hm1 - (uint32_t)r0, m, 1);                   // (x > max/2f32) as usize  :91
            m = __funnelshift_l(hm1 - (uint32_t)r1, m, 1);
        }
        m = __brev(m) >> 16;                                                 // sample i -> bit i
        // previous bit of this lane's first sample
        uint32_t pb = __shfl_up_sync(0xffffffffu, m >> 15, 1) & 1u;
        const bool has_prev = (lane > 0) || (pos > 0);
        if (lane == 0) pb = prev;
        uint32_t tm = (m ^ ((m << 1) | pb)) & 0xffffu;                       // bit i set: sample i differs from i-1
        if (!has_prev) tm &= ~1u;                                            // very first bit of the stream
        const uint32_t cnt = __popc(tm);
        uint32_t off = cnt;                                                  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, off, d);
            if (lane >= d) off += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, off, 31);
        uint32_t o = ntr + off - cnt;
        while (tm) {
            const int i = __ffs(tm) - 1;
            tm &= tm - 1;
            if (o < max_runs) trans[o] = pos + lane * 16 + i;
            ++o;
        }
        ntr += total;
        prev = __shfl_sync(0xffffffffu, m >> 15, 31) & 1u;
        pos += OOK_BLOCK;
    };
    // collected blocks are fetched PF at a time, one batch ahead of the one being sliced: a warp keeps
    // 2 x PF KB in flight instead of one block (the walk along a stream is sequential)
    constexpr int PF = 2;
    int32_t tg_next = lane < (int)n_blocks ? tag[lane] : -1;
    for (size_t b0 = 0; b0 < n_blocks; b0 += 32) {
        const int32_t tg_l = tg_next;
        {
            const size_t bn = b0 + 32 + lane;                                 // the next group's tags are on their way
            tg_next = bn < n_blocks ? tag[bn] : -1;
        }
        unsigned rem = __ballot_sync(0xffffffffu, tg_l >= 0);
        int kA[PF], kB[PF];
        uint4 dA[PF][2], dB[PF][2];
        auto take = [&](int (&ks)[PF], uint4 (&d)[PF][2]) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                ks[i] = rem ? __ffs(rem) - 1 : -1;
                if (rem) rem &= rem - 1;
                if (ks[i] >= 0) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(base + (b0 + ks[i]) * (size_t)(OOK_BLOCK * 2) + lane * 32);
                    d[i][0] = ldg_stream_u4(p); d[i][1] = ldg_stream_u4(p + 1);
                }
            }
        };
        take(kA, dA);
        while (kA[0] >= 0) {
            take(kB, dB);
#pragma unroll
            for (int i = 0; i < PF; ++i)
                if (kA[i] >= 0) process(__shfl_sync(0xffffffffu, tg_l, kA[i]), dA[i][0], dA[i][1]);
#pragma unroll
            for (int i = 0; i < PF; ++i) { kA[i] = kB[i]; dA[i][0] = dB[i][0]; dA[i][1] = dB[i][1]; }
        }
    }
    if (lane == 0) { d_ntrans[st] = ntr; d_nbits[st] = pos; }
}

// ---------------------------------------------------------------------------------------------
// K-D: dle + matchers + shaper_optional, one thread per stream
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_rng(float d, float lo, float hi) { return d >= lo && d <= hi; }

struct Shaper {
    unsigned long long acc; uint32_t n, want, count;
    unsigned long long *out; size_t cap;
    __device__ void feed(int opt)
    {
        if (opt >= 0) {                          // Some(y) => x.push(y)
            if (n < 64) acc = (acc << 1) | (unsigned long long)opt;
            n++;
        } else {                                 // None if x.len() == l => send; None => clear
            if (n == want) { if (count < cap) out[count] = acc; count++; }
            n = 0; acc = 0;
        }
    }
};

// one run at a time through the pulse-pair matchers of ratpak.rs:88-97 (looper bodies) and shaper_optional
struct Matcher {
    int proto; bool pending; float d; Shaper sh;
    __device__ void feed(uint32_t v, float dur)
    {
        if (!pending) {
            bool first;
            if (proto == 0) first = (v == 1u) && in_rng(dur, 2e-4f, 6e-4f);                                         // ratpak.rs:91
            else first = (v == 1u) && (in_rng(dur, 125e-6f, 250e-6f) || in_rng(dur, 500e-6f, 650e-6f));            // :96
            if (!first) sh.feed(-1);
            else { pending = true; d = dur; }
            return;
        }
        pending = false;                                    // a.next().unwrap(): the following run is consumed
        const float e = dur;
        if (proto == 0) {
            if (v == 0u && in_rng(e, 1.5e-3f, 2.5e-3f)) sh.feed(0);
            else if (v == 0u && in_rng(e, 3.5e-3f, 4.5e-3f)) sh.feed(1);
            else sh.feed(-1);
        } else {
            if (v == 0u && (in_rng(e, 500e-6f, 650e-6f) || in_rng(e, 125e-6f, 250e-6f))) sh.feed(d > e ? 1 : 0);
            else sh.feed(-1);
        }
        // a pulse still pending when the runs end: the reference block dies in unwrap() with nothing sent
    }
};

// One warp per stream.  Run lengths -> seconds (dle) is data-parallel: the lanes turn a chunk of 256
// transitions into durations in shared memory with coalesced reads; the matchers are sequential, so lane 0
// (proto A) and lane 1 (proto B) then walk the chunk from shared memory instead of chasing global loads.
constexpr int KD_WARPS = 4, KD_CHUNK = 256;

__global__ void __launch_bounds__(KD_WARPS * 32)
ook_match_kernel(const uint32_t *__restrict__ d_trans, const uint32_t *__restrict__ d_ntrans, size_t n_streams,
                 size_t max_runs, size_t max_packets, float s_rate_f, unsigned long long *__restrict__ d_packets,
                 uint32_t *__restrict__ d_npackets, uint32_t *__restrict__ d_runs_dbg)
{
    __shared__ float s_dur[KD_WARPS][KD_CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t st = (size_t)blockIdx.x * KD_WARPS + warp;
    if (st >= n_streams) return;
    const uint32_t *tr = d_trans + st * max_runs;
    uint32_t nr = d_ntrans[st];
    if (nr > max_runs) nr = (uint32_t)max_runs;           // overflow is reported by fetch
    // run k: value = k & 1 (the stream starts with the 0 bit of vec!(0.0)), length = tr[k] - tr[k-1]
    uint32_t *dbg = d_runs_dbg + st * max_runs;
    Matcher mt{lane, false, 0.0f,
               Shaper{0ull, 0u, lane == 0 ? 36u : 24u, 0u, d_packets + (st * 2 + (lane & 1)) * max_packets, max_packets}};
    float *dur = s_dur[warp];
    for (uint32_t k0 = 0; k0 < nr; k0 += KD_CHUNK) {
        const uint32_t nk = nr - k0 < (uint32_t)KD_CHUNK ? nr - k0 : (uint32_t)KD_CHUNK;
        for (uint32_t i = lane; i < nk; i += 32) {
            const uint32_t k = k0 + i;
            const uint32_t len = tr[k] - (k ? tr[k - 1] : 0u);
            dbg[k] = ((k & 1u) << 31) | len;
            dur[i] = __fdiv_rn((float)len, s_rate_f);                                   // dle kpn.rs:35
        }
        __syncwarp();
        if (lane < 2)
            for (uint32_t i = 0; i < nk; ++i) mt.feed((k0 + i) & 1u, dur[i]);
        __syncwarp();
    }
    if (lane < 2) d_npackets[st * 2 + lane] = mt.sh.count;
}

__global__ void ook_envelope_table_kernel(float *__restrict__ table)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536u) table[i] = lr_envelope(i >> 8, i & 0xffu);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
extern "C" int lrc_ook_create(lrc_ctx *ctx, size_t n_streams, size_t n_blocks, unsigned sample_rate,
                              size_t max_runs, size_t max_packets, lrc_ook **out)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(out && n_streams >= 1 && n_blocks >= 1 && sample_rate >= 1, LRC_ERR_INVALID, "lrc_ook_create: bad arguments");
    LRC_REQUIRE(n_blocks * (size_t)OOK_BLOCK < 0xffffffffull, LRC_ERR_UNSUPPORTED, "lrc_ook_create: capture too long (2^32 samples)");
    lrc_ook *o = new (std::nothrow) lrc_ook();
    LRC_REQUIRE(o != nullptr, LRC_ERR_NOMEM, "out of host memory");
    memset(o, 0, sizeof(*o));
    o->ctx = ctx; o->n_streams = n_streams; o->n_blocks = n_blocks; o->sample_rate = sample_rate;
    o->max_runs = max_runs ? max_runs : 4096;
    o->max_packets = max_packets ? max_packets : 64;
    o->max_bursts = n_blocks / 49 + 2;                    // a burst is at least 49 collected blocks
    const size_t sb = n_streams * n_blocks;
    cudaError_t e = cudaSuccess;
#define OOK_ALLOC(ptr, count) if (e == cudaSuccess) e = cudaMalloc(&o->ptr, (count) * sizeof(*o->ptr))
    OOK_ALLOC(d_sum, sb); OOK_ALLOC(d_max, sb); OOK_ALLOC(d_tag, sb);
    OOK_ALLOC(d_half, n_streams * o->max_bursts); OOK_ALLOC(d_bflags, n_streams * o->max_bursts);
    OOK_ALLOC(d_nbursts, n_streams);
    OOK_ALLOC(d_trans, n_streams * o->max_runs); OOK_ALLOC(d_ntrans, n_streams); OOK_ALLOC(d_nbits, n_streams);
    OOK_ALLOC(d_packets, n_streams * 2 * o->max_packets); OOK_ALLOC(d_npackets, n_streams * 2);
    OOK_ALLOC(d_runs_dbg, n_streams * o->max_runs);
    OOK_ALLOC(d_lut, (size_t)OOK_LUT_N);
    OOK_ALLOC(d_rank, (size_t)OOK_RANK_N); OOK_ALLOC(d_uniq, (size_t)65536);
#undef OOK_ALLOC
    if (e == cudaSuccess) {
        e = cudaMemsetAsync(o->d_lut, 0, OOK_LUT_BYTES, ctx->stream);           // the row padding is never read
        ook_build_lut_kernel<<<256, 256, 0, ctx->stream>>>(o->d_lut);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ook_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KA_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ook_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OOK_RANK_BYTES);
    }
    if (e == cudaSuccess) {
        // rank table for the slicer: the envelope of every byte pair, computed ON THE DEVICE by the routine the
        // block-sum kernel uses, ordered on the host (a plain sort of 65536 floats, no arithmetic)
        ook_envelope_table_kernel<<<256, 256, 0, ctx->stream>>>(o->d_uniq);          // d_uniq as scratch: index b0*256 + b1
        e = cudaGetLastError();
        std::vector<float> tab(65536);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpy(tab.data(), o->d_uniq, 65536 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) {
            std::vector<float> uq(tab);
            std::sort(uq.begin(), uq.end());
            uq.erase(std::unique(uq.begin(), uq.end()), uq.end());
            std::vector<uint16_t> rank(OOK_RANK_N, 0);
            // tab index = b0 * 256 + b1; the table is symmetric in the two bytes
            for (uint32_t i = 0; i < 65536; ++i)
                rank[ook_rank_slot(i >> 8, i & 0xffu)] = (uint16_t)(std::lower_bound(uq.begin(), uq.end(), tab[i]) - uq.begin());
            o->n_uniq = (uint32_t)uq.size();
            e = cudaMemcpy(o->d_uniq, uq.data(), uq.size() * sizeof(float), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(o->d_rank, rank.data(), rank.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        lrc_set_error("lrc_ook_create: %s", cudaGetErrorString(e));
        lrc_ook_destroy(o);
        return LRC_ERR_CUDA;
    }
    *out = o;
    return LRC_OK;
}

extern "C" int lrc_ook_destroy(lrc_ook *o)
{
    if (!o) return LRC_OK;
    cudaSetDevice(o->ctx->device);
    cudaFree(o->d_sum); cudaFree(o->d_max); cudaFree(o->d_tag); cudaFree(o->d_half); cudaFree(o->d_bflags);
    cudaFree(o->d_nbursts); cudaFree(o->d_trans); cudaFree(o->d_ntrans); cudaFree(o->d_nbits);
    cudaFree(o->d_packets); cudaFree(o->d_npackets); cudaFree(o->d_runs_dbg); cudaFree(o->d_lut);
    cudaFree(o->d_rank); cudaFree(o->d_uniq);
    delete o;
    return LRC_OK;
}

extern "C" int lrc_ook_decode(lrc_ook *o, const uint8_t *d_iq, size_t stream_stride_bytes, void *stream)
{
    LRC_REQUIRE(o != nullptr, LRC_ERR_INVALID, "null plan");
    LRC_BIND(o->ctx);
    LRC_REQUIRE(d_iq != nullptr, LRC_ERR_INVALID, "lrc_ook_decode: null input");
    LRC_REQUIRE(stream_stride_bytes >= o->n_blocks * (size_t)OOK_BLOCK * 2, LRC_ERR_INVALID, "lrc_ook_decode: stride too short");
    LRC_REQUIRE(((uintptr_t)d_iq & 15) == 0 && (stream_stride_bytes & 15) == 0, LRC_ERR_INVALID,
                "lrc_ook_decode: input and stream stride must be 16-byte aligned");
    cudaStream_t s = lrc_stream(o->ctx, stream);
    const size_t groups = ((o->n_blocks + 31) / 32) * o->n_streams;
    size_t blocks = ceil_div(groups, (size_t)KA_WARPS);
    const size_t cap = (size_t)o->ctx->n_sm;           // one persistent CTA per SM (the table fills its shared memory)
    if (blocks > cap) blocks = cap;
    ook_block_kernel<<<(unsigned)blocks, KA_WARPS * 32, KA_SMEM_BYTES, s>>>(d_iq, stream_stride_bytes, o->n_streams,
                                                                           o->n_blocks, o->d_lut, o->d_sum, o->d_max);
    LRC_CUDA(cudaGetLastError());
    ook_trigger_kernel<<<(unsigned)ceil_div(o->n_streams, (size_t)KB_THREADS), KB_THREADS, 0, s>>>(
        o->d_sum, o->d_max, o->n_streams, o->n_blocks, o->max_bursts, o->d_tag, o->d_half, o->d_bflags, o->d_nbursts);
    LRC_CUDA(cudaGetLastError());
    ook_rle_kernel<<<(unsigned)ceil_div(o->n_streams * 32, (size_t)KC_THREADS), KC_THREADS, OOK_RANK_BYTES, s>>>(
        d_iq, stream_stride_bytes, o->n_streams, o->n_blocks, o->max_bursts, o->max_runs, o->d_rank, o->d_tag, o->d_half,
        o->d_uniq, o->n_uniq, o->d_bflags, o->d_trans, o->d_ntrans, o->d_nbits);
    LRC_CUDA(cudaGetLastError());
    ook_match_kernel<<<(unsigned)ceil_div(o->n_streams, (size_t)KD_WARPS), KD_WARPS * 32, 0, s>>>(
        o->d_trans, o->d_ntrans, o->n_streams, o->max_runs, o->max_packets, (float)o->sample_rate, o->d_packets,
        o->d_npackets, o->d_runs_dbg);
    LRC_CUDA(cudaGetLastError());
    return LRC_OK;
}

extern "C" int lrc_ook_fetch_packets(lrc_ook *o, lrc_ook_packet *h_packets, size_t cap, size_t *n_packets)
{
    LRC_REQUIRE(o && n_packets, LRC_ERR_INVALID, "lrc_ook_fetch_packets: null argument");
    LRC_BIND(o->ctx);
    LRC_CUDA(cudaDeviceSynchronize());
    const size_t ns = o->n_streams;
    std::vector<uint32_t> np(ns * 2), ntr(ns), nb(ns);
    LRC_CUDA(cudaMemcpy(np.data(), o->d_npackets, np.size() * 4, cudaMemcpyDeviceToHost));
    LRC_CUDA(cudaMemcpy(ntr.data(), o->d_ntrans, ns * 4, cudaMemcpyDeviceToHost));
    LRC_CUDA(cudaMemcpy(nb.data(), o->d_nbursts, ns * 4, cudaMemcpyDeviceToHost));
    size_t total = 0;
    for (size_t s = 0; s < ns; ++s) {
        if (ntr[s] > o->max_runs || np[2 * s] > o->max_packets || np[2 * s + 1] > o->max_packets || nb[s] > o->max_bursts) {
            lrc_set_error("lrc_ook_fetch_packets: stream %zu overflowed (runs %u/%zu, packets %u,%u/%zu, bursts %u/%zu)",
                          s, ntr[s], o->max_runs, np[2 * s], np[2 * s + 1], o->max_packets, nb[s], o->max_bursts);
            *n_packets = 0;
            return LRC_ERR_CAPACITY;
        }
        total += np[2 * s] + np[2 * s + 1];
    }
    *n_packets = total;
    if (total == 0) return LRC_OK;
    if (cap < total || !h_packets) {
        lrc_set_error("lrc_ook_fetch_packets: %zu packets, capacity %zu", total, cap);
        return LRC_ERR_CAPACITY;
    }
    std::vector<unsigned long long> pk(ns * 2 * o->max_packets);
    LRC_CUDA(cudaMemcpy(pk.data(), o->d_packets, pk.size() * 8, cudaMemcpyDeviceToHost));
    size_t w = 0;
    for (size_t s = 0; s < ns; ++s)
        for (int proto = 0; proto < 2; ++proto) {
            const uint32_t nbits = proto == 0 ? 36u : 24u;
            for (uint32_t k = 0; k < np[2 * s + proto]; ++k) {
                lrc_ook_packet &p = h_packets[w++];
                p.stream = (uint32_t)s; p.proto = (uint32_t)proto; p.seq = k; p.nbits = nbits;
                memset(p.bits, 0, sizeof(p.bits));
                const unsigned long long v = pk[(s * 2 + proto) * o->max_packets + k];
                for (uint32_t i = 0; i < nbits; ++i) p.bits[i] = (uint8_t)((v >> (nbits - 1 - i)) & 1ull);
            }
        }
    return LRC_OK;
}

extern "C" int lrc_ook_debug_ptrs(lrc_ook *o, const float **d_block_sums, const uint32_t **d_run_counts,
                                  const uint32_t **d_runs, const uint32_t **d_n_bits)
{
    LRC_REQUIRE(o != nullptr, LRC_ERR_INVALID, "null plan");
    if (d_block_sums) *d_block_sums = o->d_sum;
    if (d_run_counts) *d_run_counts = o->d_ntrans;
    if (d_runs) *d_runs = o->d_runs_dbg;
    if (d_n_bits) *d_n_bits = o->d_nbits;
    return LRC_OK;
}

extern "C" int lrc_eat(const uint8_t *bits, size_t nbits, const size_t *widths, size_t n_widths, size_t *out)
{
    // kpn::eat / kpn::b2d (kpn.rs:111-124): consecutive MSB-first fields
    LRC_REQUIRE(bits && widths && out, LRC_ERR_INVALID, "lrc_eat: null argument");
    size_t i = 0;
    for (size_t w = 0; w < n_widths; ++w) {
        if (i + widths[w] > nbits) {
            lrc_set_error("lrc_eat: fields need %zu bits, packet has %zu (the reference slices out of bounds and panics)",
                          i + widths[w], nbits);
            return LRC_ERR_LENGTH;
        }
        size_t v = 0;
        for (size_t k = 0; k < widths[w]; ++k) v += ((size_t)1 << (widths[w] - k - 1)) * bits[i + k];
        out[w] = v;
        i += widths[w];
    }
    return LRC_OK;
}

extern "C" int lrc_ook_envelope_table(lrc_ctx *ctx, float *d_table, void *stream)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(d_table != nullptr, LRC_ERR_INVALID, "lrc_ook_envelope_table: null output");
    ook_envelope_table_kernel<<<256, 256, 0, lrc_stream(ctx, stream)>>>(d_table);
    LRC_CUDA(cudaGetLastError());
    return LRC_OK;
}
