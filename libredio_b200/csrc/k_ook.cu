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
// kernels read it from a triangular table of 256*257/2 = 32896 floats (131.6 KB) held in shared memory.
// The table is filled ON THE DEVICE by lr_envelope above when the plan is created; the u8 domain being
// finite, table == formula for all 65536 pairs is a proof of equivalence, checked exhaustively by
// tests/test_gpu_ook_fastfir.py::test_envelope_exhaustive_65536_pairs_bit_exact (formula vs CPU) and by the
// bit-exact block sums of every OOK test (table vs CPU).
constexpr int OOK_LUT_N = 256 * 257 / 2;
constexpr int OOK_LUT_BYTES = OOK_LUT_N * 4;

__device__ __forceinline__ float lut_envelope(const float *lut, uint32_t b0, uint32_t b1)
{
    const uint32_t hi = max(b0, b1), lo = min(b0, b1);
    return lut[((hi * (hi + 1u)) >> 1) + lo];
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
    if (lo <= hi) lut[((hi * (hi + 1u)) >> 1) + lo] = lr_envelope(hi, lo);
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
};

// ---------------------------------------------------------------------------------------------
// K-A: envelope, sequential block sum, block max.  One warp handles 32 consecutive blocks of one
// stream: 32-sample slabs are loaded coalesced (4 lanes x 16 B per block row), turned into envelopes
// by the loading lane, parked in a padded shared tile, and lane b then adds row b in sample order.
// ---------------------------------------------------------------------------------------------
constexpr int KA_WARPS = 16;
constexpr int KA_SLAB = 32;                  // samples per block row per slab
constexpr int KA_LPR = KA_SLAB * 2 / 16;     // lanes (16-byte loads) per block row
constexpr int KA_RPI = 32 / KA_LPR;          // block rows per load iteration
constexpr int KA_LD = KA_SLAB + 1;           // padded row length (floats)

constexpr int KA_SMEM_BYTES = OOK_LUT_BYTES + KA_WARPS * 32 * KA_LD * 4;

__global__ void __launch_bounds__(KA_WARPS * 32, 1)
ook_block_kernel(const uint8_t *__restrict__ iq, size_t stream_stride, size_t n_streams, size_t n_blocks,
                 const float *__restrict__ g_lut, float *__restrict__ d_sum, float *__restrict__ d_max)
{
    extern __shared__ __align__(16) float ka_smem[];
    float *lut = ka_smem;
    lut_load(lut, g_lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *env = ka_smem + OOK_LUT_N + warp * (32 * KA_LD);
    const size_t groups_per_stream = (n_blocks + 31) / 32;
    const size_t n_groups = groups_per_stream * n_streams;
    const size_t warps_total = (size_t)gridDim.x * KA_WARPS;
    for (size_t grp = (size_t)blockIdx.x * KA_WARPS + warp; grp < n_groups; grp += warps_total) {
        const size_t st = grp / groups_per_stream, b0 = (grp % groups_per_stream) * 32;
        const int nb = (int)((n_blocks - b0) < 32 ? (n_blocks - b0) : 32);
        const uint8_t *base = iq + st * stream_stride + b0 * (size_t)(OOK_BLOCK * 2);
        float s = 0.0f, mx = 0.0f;
        // slab loads are software-pipelined: slab i+1 is in flight while slab i is converted and summed
        constexpr int NIT = 32 / KA_RPI;
        uint4 nxt[NIT];
        auto load_slab = [&](int slab) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int row = it * KA_RPI + lane / KA_LPR, col = lane % KA_LPR;
                nxt[it] = row < nb ? ldg_stream_u4(reinterpret_cast<const uint4 *>(
                                         base + (size_t)row * (OOK_BLOCK * 2) + slab * (KA_SLAB * 2) + col * 16))
                                   : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        load_slab(0);
        for (int slab = 0; slab < OOK_BLOCK / KA_SLAB; ++slab) {
            uint4 cur[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) cur[it] = nxt[it];
            if (slab + 1 < OOK_BLOCK / KA_SLAB) load_slab(slab + 1);
            // 32 rows x 64 B: 4 iterations of (8 rows x 4 lanes x 16 B); the padded tile makes both the
            // envelope stores here and the row-wise reads below bank-conflict free
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int row = it * KA_RPI + lane / KA_LPR, col = lane % KA_LPR;
                const uint32_t w[4] = {cur[it].x, cur[it].y, cur[it].z, cur[it].w};
                float *dst = env + row * KA_LD + col * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dst[2 * k]     = lut_envelope(lut, w[k] & 0xffu, (w[k] >> 8) & 0xffu);
                    dst[2 * k + 1] = lut_envelope(lut, (w[k] >> 16) & 0xffu, w[k] >> 24);
                }
            }
            __syncwarp();
            if (lane < nb) {
                const float *r = env + lane * KA_LD;
#pragma unroll
                for (int j = 0; j < KA_SLAB; ++j) {
                    const float e = r[j];
                    s = __fadd_rn(s, e);                 // samples.iter().sum(): left to right from 0.0
                    mx = fmaxf(mx, e);
                }
            }
            __syncwarp();
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
__global__ void __launch_bounds__(128)
ook_trigger_kernel(const float *__restrict__ d_sum, const float *__restrict__ d_max, size_t n_streams,
                   size_t n_blocks, size_t max_bursts, int32_t *__restrict__ d_tag, float *__restrict__ d_half,
                   uint8_t *__restrict__ d_bflags, uint32_t *__restrict__ d_nbursts)
{
    const size_t st = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= n_streams) return;
    const float *sum = d_sum + st * n_blocks, *bmax = d_max + st * n_blocks;
    int32_t *tag = d_tag + st * n_blocks;
    float *half = d_half + st * max_bursts;
    uint8_t *flags = d_bflags + st * max_bursts;
    long long trigger = 0;                    // :41
    float threshold = 0.0f;                   // :44
    unsigned long long buf_len = 1;           // :43 sample_buffer = vec!(0.0)
    bool lead0 = true;                        // the buffer currently starts with that 0.0
    float cur_max = 0.0f;
    uint32_t burst = 0;                       // index of the burst being collected
    size_t burst_first_block = 0;
    bool burst_has_blocks = false;
    // blocks are walked four at a time; the next four sums/maxima are already in flight (the state machine
    // itself is sequential, the loads are not)
    const bool vec = (n_blocks % 4 == 0);
    auto load4 = [&](size_t b, float *s4, float *m4) {
        if (vec && b + 4 <= n_blocks) {
            const float4 a = *reinterpret_cast<const float4 *>(sum + b), c = *reinterpret_cast<const float4 *>(bmax + b);
            s4[0] = a.x; s4[1] = a.y; s4[2] = a.z; s4[3] = a.w;
            m4[0] = c.x; m4[1] = c.y; m4[2] = c.z; m4[3] = c.w;
        } else {
            for (int u = 0; u < 4; ++u) {
                s4[u] = (b + u < n_blocks) ? sum[b + u] : 0.0f;
                m4[u] = (b + u < n_blocks) ? bmax[b + u] : 0.0f;
            }
        }
    };
    float ns[4], nm[4];
    load4(0, ns, nm);
    for (size_t b4 = 0; b4 < n_blocks; b4 += 4) {
        float cs[4], cm[4];
        for (int u = 0; u < 4; ++u) { cs[u] = ns[u]; cm[u] = nm[u]; }
        if (b4 + 4 < n_blocks) load4(b4 + 4, ns, nm);
        for (int u = 0; u < 4 && b4 + u < n_blocks; ++u) {
            const size_t b = b4 + u;
            trigger -= 1;                                                       // :46
            const float s = cs[u];                                             // :48
            if (buf_len > 1000ull * OOK_TRIGGER_DURATION * OOK_BLOCK) {         // :52-54 OOM guard
                if (burst_has_blocks)
                    for (size_t k = burst_first_block; k < b; ++k) if (tag[k] == (int32_t)burst) tag[k] = -1;
                buf_len = 1; lead0 = true; cur_max = 0.0f; burst_has_blocks = false;
            }
            if (threshold == 0.0f) threshold = s;                               // :57-59
            if (trigger < 0) {                                                  // :62-65
                threshold = __fadd_rn(threshold, __fdiv_rn(s, 1000.0f));
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
                cur_max = fmaxf(cur_max, cm[u]);
            }
            tag[b] = tg;
            if (trigger == 0) {                                                 // :78-81 send, buffer = vec!()
                if (burst < max_bursts) {
                    half[burst] = __fdiv_rn(cur_max, 2.0f);                     // discretize :90-91 max/2f32
                    flags[burst] = (uint8_t)(1u | (lead0 ? 2u : 0u));
                }
                burst += 1;
                buf_len = 0; lead0 = false; cur_max = 0.0f; burst_has_blocks = false;
            }
        }
    }
    // a burst still open when the capture ends is never sent: un-tag its blocks
    if (burst_has_blocks)
        for (size_t k = burst_first_block; k < n_blocks; ++k) if (tag[k] == (int32_t)burst) tag[k] = -1;
    d_nbursts[st] = burst;                    // may exceed max_bursts -> reported by fetch
}

// ---------------------------------------------------------------------------------------------
// K-C: discretize + rle as bit masks, one warp per stream.  Lane l owns samples [16 l, 16 l + 16) of a
// collected block (32 contiguous bytes).  The positions (in the flattened bit stream) where the value
// changes are appended to the stream's transition list in order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
ook_rle_kernel(const uint8_t *__restrict__ iq, size_t stream_stride, size_t n_streams, size_t n_blocks,
               size_t max_bursts, size_t max_runs, const float *__restrict__ g_lut,
               const int32_t *__restrict__ d_tag,
               const float *__restrict__ d_half, const uint8_t *__restrict__ d_bflags,
               uint32_t *__restrict__ d_trans, uint32_t *__restrict__ d_ntrans, uint32_t *__restrict__ d_nbits)
{
    extern __shared__ __align__(16) float kc_lut[];
    lut_load(kc_lut, g_lut);
    const float *lut = kc_lut;
    const int lane = threadIdx.x & 31;
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
    float h = 0.0f;
    for (size_t b0 = 0; b0 < n_blocks; b0 += 32) {
        // fetch 32 tags at once; skip quickly over untriggered stretches
        const size_t bi = b0 + lane;
        const int32_t tg_l = bi < n_blocks ? tag[bi] : -1;
        unsigned live = __ballot_sync(0xffffffffu, tg_l >= 0);
        uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
        if (live) {                                     // first live block of this group
            const uint4 *p = reinterpret_cast<const uint4 *>(base + (b0 + __ffs(live) - 1) * (size_t)(OOK_BLOCK * 2) + lane * 32);
            n0 = ldg_stream_u4(p); n1 = ldg_stream_u4(p + 1);
        }
        while (live) {
            const int k = __ffs(live) - 1;
            live &= live - 1;
            const int32_t tg = __shfl_sync(0xffffffffu, tg_l, k);
            const size_t b = b0 + k;
            const uint4 q0 = n0, q1 = n1;
            if (live) {                                 // the next live block is already on its way
                const uint4 *p = reinterpret_cast<const uint4 *>(base + (b0 + __ffs(live) - 1) * (size_t)(OOK_BLOCK * 2) + lane * 32);
                n0 = ldg_stream_u4(p); n1 = ldg_stream_u4(p + 1);
            }
            if (tg != cur_burst) {
                cur_burst = tg;
                h = half[tg];
                if (flags[tg] & 2u) {
                    // the burst starts with the literal 0.0 of vec!(0.0): 0.0 > max/2 is false -> bit 0
                    if (pos > 0 && prev != 0u) { if (lane == 0 && ntr < max_runs) trans[ntr] = pos; ntr++; }
                    prev = 0u; pos += 1;
                }
            }
            (void)b;
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float e0 = lut_envelope(lut, w[i] & 0xffu, (w[i] >> 8) & 0xffu);
                const float e1 = lut_envelope(lut, (w[i] >> 16) & 0xffu, w[i] >> 24);
                m |= (e0 > h ? 1u : 0u) << (2 * i);                          // (x > max/2f32) as usize  :91
                m |= (e1 > h ? 1u : 0u) << (2 * i + 1);
            }
            // previous bit of this lane's first sample
            uint32_t pb = __shfl_up_sync(0xffffffffu, m >> 15, 1) & 1u;
            const bool has_prev = (lane > 0) || (pos > 0);
            if (lane == 0) pb = prev;
            uint32_t tm = (m ^ ((m << 1) | pb)) & 0xffffu;                   // bit i set: sample i differs from i-1
            if (!has_prev) tm &= ~1u;                                        // very first bit of the stream
            const uint32_t cnt = __popc(tm);
            uint32_t off = cnt;                                              // inclusive warp scan
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

__global__ void __launch_bounds__(128)
ook_match_kernel(const uint32_t *__restrict__ d_trans, const uint32_t *__restrict__ d_ntrans, size_t n_streams,
                 size_t max_runs, size_t max_packets, float s_rate_f, unsigned long long *__restrict__ d_packets,
                 uint32_t *__restrict__ d_npackets, uint32_t *__restrict__ d_runs_dbg)
{
    const size_t st = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= n_streams) return;
    const uint32_t *tr = d_trans + st * max_runs;
    uint32_t nr = d_ntrans[st];
    if (nr > max_runs) nr = (uint32_t)max_runs;           // overflow is reported by fetch
    // run k: value = k & 1 (the stream starts with the 0 bit of vec!(0.0)), length = tr[k] - tr[k-1]
    uint32_t *dbg = d_runs_dbg + st * max_runs;
    for (uint32_t k = 0; k < nr; ++k) dbg[k] = ((k & 1u) << 31) | (tr[k] - (k ? tr[k - 1] : 0u));
    for (int proto = 0; proto < 2; ++proto) {
        Shaper sh{0ull, 0u, proto == 0 ? 36u : 24u, 0u, d_packets + (st * 2 + proto) * max_packets, max_packets};
        uint32_t k = 0;
        while (k < nr) {
            const uint32_t v = k & 1u;
            const float d = __fdiv_rn((float)(tr[k] - (k ? tr[k - 1] : 0u)), s_rate_f);   // dle kpn.rs:35
            ++k;
            bool first;
            if (proto == 0) first = (v == 1u) && in_rng(d, 2e-4f, 6e-4f);                                  // ratpak.rs:91
            else first = (v == 1u) && (in_rng(d, 125e-6f, 250e-6f) || in_rng(d, 500e-6f, 650e-6f));          // :96
            if (!first) { sh.feed(-1); continue; }
            if (k >= nr) break;                                // a.next().unwrap() on a drained port: the block dies
            const uint32_t v2 = k & 1u;
            const float e = __fdiv_rn((float)(tr[k] - tr[k - 1]), s_rate_f);
            ++k;
            if (proto == 0) {
                if (v2 == 0u && in_rng(e, 1.5e-3f, 2.5e-3f)) sh.feed(0);
                else if (v2 == 0u && in_rng(e, 3.5e-3f, 4.5e-3f)) sh.feed(1);
                else sh.feed(-1);
            } else {
                if (v2 == 0u && (in_rng(e, 500e-6f, 650e-6f) || in_rng(e, 125e-6f, 250e-6f))) sh.feed(d > e ? 1 : 0);
                else sh.feed(-1);
            }
        }
        d_npackets[st * 2 + proto] = sh.count;
    }
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
#undef OOK_ALLOC
    if (e == cudaSuccess) {
        ook_build_lut_kernel<<<256, 256, 0, ctx->stream>>>(o->d_lut);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ook_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KA_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ook_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OOK_LUT_BYTES);
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
    ook_trigger_kernel<<<(unsigned)ceil_div(o->n_streams, 128), 128, 0, s>>>(
        o->d_sum, o->d_max, o->n_streams, o->n_blocks, o->max_bursts, o->d_tag, o->d_half, o->d_bflags, o->d_nbursts);
    LRC_CUDA(cudaGetLastError());
    ook_rle_kernel<<<(unsigned)ceil_div(o->n_streams * 32, 512), 512, OOK_LUT_BYTES, s>>>(
        d_iq, stream_stride_bytes, o->n_streams, o->n_blocks, o->max_bursts, o->max_runs, o->d_lut, o->d_tag, o->d_half,
        o->d_bflags, o->d_trans, o->d_ntrans, o->d_nbits);
    LRC_CUDA(cudaGetLastError());
    ook_match_kernel<<<(unsigned)ceil_div(o->n_streams, 128), 128, 0, s>>>(
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
