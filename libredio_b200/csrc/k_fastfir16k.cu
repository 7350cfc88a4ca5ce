// k_fastfir16k.cu -- overlap-save with nfft = 16384 (the default block size for 2048 < nh <= 8192 on request, and the
// size lrc_fastfir_create picks itself for the config-5 shape when the caller does not insist on kiss_fastfir's).
//
// Why: with nh = 4096 an 8192-point block keeps 4097 of 8192 outputs (50 %), a 16384-point block keeps 12289 of 16384
// (75 %): 1.39x fewer butterflies per output sample (kiss_fastfir.c:81-93 picks the block size by a cost model of
// its own; the result is the same convolution whatever the block size).
//
//   16384 = 16 x 1024,  n = 1024 n1 + n2,  k = k1 + 16 k2
//   P1   column n2: DFT16 over n1 straight from global memory, times W_16384^(n2 k1), into shared row k1
//   P2   warp k1 owns row k1: 1024 points as 32 x 32 in registers (the row buffer itself is the exchange tile),
//        .* H[k1 + 16 k2], inverse transform, back into the row
//   P1'  column n2: times conj(W^(n2 k1)), IDFT16 over k1, outputs 1024 n1 + n2 < ngood to global memory; fused per
//        thread with P1 of the CTA's next block (its loads are in flight during the inverse butterflies) -- no barrier
//        between the two blocks' P1 phases because a column is private to its thread in both.
// Two CTA barriers per block.
//
// Shared-memory layout (round 2).  ncu on the first version (profiles/r2_ff16k_v1_ncu_keys.txt): the FP32 pipe (packed
// FADD2/FFMA2, 2 cycles each) and the shared-memory pipe both need ~12.5 k cycles per block and the top stall was
// mio_throttle (2.4 per issue) -- the LSU instruction queue full of 64-bit shared accesses arriving in bursts of 32.
// So every shared access that can be is now 128 bits wide (half the LSU instructions for the same bytes):
//   row k1, element n2 = l + 32 e  lives at  l * 34 + e   (a 32-entry slab per lane l, padded to 34 so that slabs
//   start 16 bytes apart modulo 128: LDS.128/STS.128 of 8 consecutive lanes hit 8 distinct bank groups)
//   - P2's lane l reads/writes its 32 points x[l + 32 e] as 16 LDS.128 / STS.128; the transposing reads of the two
//     exchanges stay 64-bit (column reads, conflict-free)
//   - a P1 thread (lane l, warp j) owns the column PAIR n2 = l + 64 j and n2 + 32, i.e. e = 2 j and 2 j + 1: adjacent
//     entries of slab l, one STS.128 / LDS.128 per row.  The pair's second twiddle is W_512^k1 (a literal) times the first.
//   - twiddle tables are stored per thread / per lane contiguously (pitch 18 / 34 entries) and read as LDS.128;
//     the inverse uses the same tables conjugated in the multiply; H is stored so that a lane's two consecutive
//     bins are one coalesced LDG.128.
// Shared memory: 16 rows x 1088 complex (136 KB) + 512 x 18 P1 twiddles (72 KB) + 32 x 34 warp-transform twiddles
// (8.5 KB) = 216.5 KB, one 512-thread CTA per SM.
#include "fft_core.cuh"

using namespace lrfft;

namespace ff16k {
constexpr int N = 16384, NPAIR = 512, ROWS = 16;
constexpr int LP = 34;                          // entries per lane slab (32 used)
constexpr int RP = 32 * LP;                     // entries per row
constexpr int DATA_CPX = ROWS * RP;
constexpr int T1P = 18;                         // P1 twiddles per thread: q = 0..15 (+2 pad)
constexpr int TW1_CPX = NPAIR * T1P;
constexpr int TWW_CPX = 32 * LP;                // [lane][r] : W_1024^(lane r)
constexpr int SMEM_BYTES = (DATA_CPX + TW1_CPX + TWW_CPX) * 8;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// exp(-2 pi j q / 512), q < 16 (compile-time index after unrolling): W_16384^(32 q)
__device__ __forceinline__ float2 w512(int q)
{
    constexpr float C[16] = {1.00000000000000000000f, 0.99992470183914450299f, 0.99969881869620424997f,
                             0.99932238458834954375f, 0.99879545620517240501f, 0.99811811290014917919f,
                             0.99729045667869020697f, 0.99631261218277800129f, 0.99518472667219692873f,
                             0.99390697000235606051f, 0.99247953459870996706f, 0.99090263542778000971f,
                             0.98917650996478101444f, 0.98730141815785843473f, 0.98527764238894122162f,
                             0.98310548743121628501f};
    constexpr float S[16] = {0.00000000000000000000f, 0.01227153828571992539f, 0.02454122852291228812f,
                             0.03680722294135883171f, 0.04906767432741801493f, 0.06132073630220857829f,
                             0.07356456359966742631f, 0.08579731234443989385f, 0.09801714032956060363f,
                             0.11022220729388305938f, 0.12241067519921619566f, 0.13458070850712616773f,
                             0.14673047445536174793f, 0.15885814333386144570f, 0.17096188876030121717f,
                             0.18303988795514095078f};
    return make_float2(C[q], -S[q]);
}

__device__ __forceinline__ float2 lo(float4 x) { return make_float2(x.x, x.y); }
__device__ __forceinline__ float2 hi(float4 x) { return make_float2(x.z, x.w); }
__device__ __forceinline__ float4 pack(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float2 ldg_nc_f2(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// one 1024-point transform by the calling warp, in place in its lane slabs: on entry and exit v[e] = x[lane + 32 e]
// (entry) / X[lane + 32 e] (exit).  `row`: the warp's RP-entry buffer, free to be overwritten; tww: [lane][r] twiddles.
template <bool INV>
__device__ __forceinline__ void warp_fft1024(float2 *v, float2 *row, const float2 *tww, int lane)
{
    RegFFT<32, INV>::run(v);
    const float4 *tw4 = reinterpret_cast<const float4 *>(tww + lane * LP);
    float4 *slab4 = reinterpret_cast<float4 *>(row + lane * LP);
    __syncwarp();                                        // every lane is done reading the tile (previous exchange)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 w = tw4[i];
        const float2 a = i ? (INV ? cmul_conjb(v[2 * i], lo(w)) : cmulf(v[2 * i], lo(w))) : v[0];
        const float2 b = INV ? cmul_conjb(v[2 * i + 1], hi(w)) : cmulf(v[2 * i + 1], hi(w));
        slab4[i] = pack(a, b);                           // tile[n2' = lane][k1' = 2 i, 2 i + 1]
    }
    __syncwarp();
    const float2 *col = row + lane;                      // tile[e][k1' = lane]
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = col[e * LP];
    RegFFT<32, INV>::run(v);
}
}  // namespace ff16k

// Measured and rejected for the P1 phase (whose global loads cannot be held in registers across P2 at 128 registers/thread):
//   - staging the next block's first column in shared memory with cp.async during P2 (P1 twiddles factorised into two
//     small tables to make room): 121 Gsamples/s against 140 -- 8192 LDGSTS per CTA sit in the LSU queue in front of the
//     barrier (lg_throttle 0.12 -> 0.79, barrier 0.63 -> 1.07 per issue), profiles/r2_ff16k_staged_ncu_phases.txt;
//   - a CTA-uniform fast path with unpredicated loads: 133 against 140.
// A TMA tensor copy would take the loads off the LSU, but block starts (multiples of ngood = 12289 samples) are only
// 8-byte aligned every other block.
//
// NT = 512: one column pair and one row per thread / warp, 128 registers.  NT = 256: two pairs and two rows, 255 registers --
// room to request a whole pair of the next block BEFORE P2 and the other pair at the head of the P1 phase, so no load is
// ever waited for; half the warps to hide everything else.  Chosen by measurement (lrc_fastfir16k_launch).
template <int NT>
__global__ void __launch_bounds__(NT, 1)
fastfir16k_kernel(const float2 *__restrict__ in, size_t n_in, float2 *__restrict__ out, size_t n_blocks_full,
                  size_t n_blocks, size_t ngood, size_t flush_keep, const float2 *__restrict__ tw /* W_16384^k */,
                  const float2 *__restrict__ tw1k /* W_1024^k */, const float4 *__restrict__ Hq)
{
    using namespace ff16k;
    constexpr int NW = NT / 32, PAIRS = NPAIR / NT, RPW = ROWS / NW;          // warps, pairs per thread, rows per warp
    extern __shared__ __align__(16) float2 ff16k_smem[];
    float2 *sd = ff16k_smem;                         // [16][RP]
    float2 *tw1 = sd + DATA_CPX;                     // [512 pairs][T1P]
    float2 *tww = tw1 + TW1_CPX;                     // [32][LP]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < TW1_CPX; i += NT) {
        const int pr = i / T1P, q = i % T1P;
        const int col = (pr & 31) + 64 * (pr >> 5);
        tw1[i] = q < 16 ? __ldg(tw + col * q) : make_float2(0.f, 0.f);        // 991 * 15 < 16384
    }
    for (int i = t; i < TWW_CPX; i += NT) {
        const int ln = i / LP, r = i % LP;
        tww[i] = r < 32 ? __ldg(tw1k + ln * r) : make_float2(0.f, 0.f);       // 31 * 31 < 1024
    }
    __syncthreads();

    // pair p of this thread: pair index pr = t + NT p = lane + 32 j, columns ca = lane + 64 j and ca + 32,
    // cells (row q) at sd[q RP + lane LP + 2 j] (.lo = column ca, .hi = column ca + 32)
    auto col_a = [&](int p) { return lane + 64 * (warp + NW * p); };
    auto twp_of = [&](int p) { return reinterpret_cast<const float4 *>(tw1 + (t + NT * p) * T1P); };
    auto cell_of = [&](int p) { return reinterpret_cast<float4 *>(sd + lane * LP + 2 * (warp + NW * p)); };

    // the 2 x 16 inputs of a column pair of block `blk`; volatile loads: they stay where they are written
    auto p1_load = [&](size_t blk, int ca, float2 *va, float2 *vb) {
        const size_t s0 = blk * ngood;
        const float2 *src = in + s0 + ca;
        if (n_in - s0 >= (size_t)N) {                            // every block but a ragged last one (CTA-uniform)
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                va[r] = ldg_nc_f2(src + 1024 * r);
                vb[r] = ldg_nc_f2(src + 1024 * r + 32);
            }
        } else {
            const int avail = (int)(n_in - s0);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                va[r] = ca + 1024 * r < avail ? ldg_nc_f2(src + 1024 * r) : make_float2(0.f, 0.f);
                vb[r] = ca + 32 + 1024 * r < avail ? ldg_nc_f2(src + 1024 * r + 32) : make_float2(0.f, 0.f);
            }
        }
    };
    // forward butterflies of a column pair; results into the pair's 16 shared cells
    auto p1_forward = [&](int p, float2 *va, float2 *vb) {
        const float4 *twp = twp_of(p);
        float4 *cell = cell_of(p);
        RegFFT<16, false>::run(va);
        RegFFT<16, false>::run(vb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 w = twp[i];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * i + h;
                const float2 wq = h ? hi(w) : lo(w);
                const float2 a = q ? cmulf(va[q], wq) : va[0];
                const float2 b = q ? cmulf(cmulf(vb[q], w512(q)), wq) : vb[0];
                cell[q * (RP / 2)] = pack(a, b);
            }
        }
    };
    // inverse side of a pair, part 1: cells -> registers, conjugated twiddles
    auto p1_unload = [&](int p, float2 *ua, float2 *ub) {
        const float4 *twp = twp_of(p);
        const float4 *cell = cell_of(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 w = twp[i];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * i + h;
                const float2 wq = h ? hi(w) : lo(w);
                const float4 x = cell[q * (RP / 2)];
                ua[q] = q ? cmul_conjb(lo(x), wq) : lo(x);
                ub[q] = q ? cmul_conjb(cmul_conjb(hi(x), w512(q)), wq) : hi(x);
            }
        }
    };
    // part 2: IDFT16 of one column, outputs 1024 r + col < keep to global memory
    auto p1_emit = [&](float2 *u, float2 *dst, int col, int keep) {
        RegFFT<16, true>::run(u);
        const int cnt = keep > col ? ((keep - 1 - col) >> 10) + 1 : 0;
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (r < cnt) __stcs(dst + 1024 * r, u[r]);
    };

    size_t b = blockIdx.x;
    if (b < n_blocks) {
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) {
            float2 va[16], vb[16];
            p1_load(b, col_a(p), va, vb);
            p1_forward(p, va, vb);
        }
    }
    for (; b < n_blocks; b += gridDim.x) {
        const size_t nb = b + gridDim.x;
        const bool more = nb < n_blocks;
        // the next block's input (1024 lines) is asked into L2 now; its P1 loads come later
        if (more) {
            const char *nsrc = reinterpret_cast<const char *>(in + nb * ngood);
#pragma unroll
            for (int k = 0; k < 1024 / NT; ++k) {
                const size_t off = (size_t)(t + NT * k) * 128;
                if (nb * ngood * 8 + off < n_in * 8) asm volatile("prefetch.global.L2 [%0];" :: "l"(nsrc + off));
            }
        }
        float2 nxa[16], nxb[16];
        if (NT == 256 && more) p1_load(nb, col_a(0), nxa, nxb);   // in flight across the whole of P2
        __syncthreads();
        // ---- P2: a warp transforms its rows, multiplies by its slice of H, transforms back -------------------
#pragma unroll 1
        for (int rr = 0; rr < RPW; ++rr) {
            const int k1 = warp + NW * rr;
            float2 *row = sd + RP * k1;
            float4 *slab4 = reinterpret_cast<float4 *>(row + lane * LP);
            float2 v[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 x = slab4[i];
                v[2 * i] = lo(x);
                v[2 * i + 1] = hi(x);
            }
            warp_fft1024<false>(v, row, tww, lane);              // X[k1 + 16 (lane + 32 e)]
            const float4 *hq = Hq + (size_t)k1 * 512 + lane;
#pragma unroll
            for (int i = 0; i < 16; ++i) {                       // C_MUL  kiss_fastfir.c:180-184
                const float4 h = __ldg(hq + 32 * i);
                v[2 * i] = cmulf(v[2 * i], lo(h));
                v[2 * i + 1] = cmulf(v[2 * i + 1], hi(h));
            }
            warp_fft1024<true>(v, row, tww, lane);
            __syncwarp();                                        // the tile reads of the last exchange are done
#pragma unroll
            for (int i = 0; i < 16; ++i) slab4[i] = pack(v[2 * i], v[2 * i + 1]);
        }
        __syncthreads();
        // ---- P1 inverse of this block fused with P1 forward of the CTA's next block -----------------------
        const int keep = (int)((b < n_blocks_full) ? ngood : flush_keep);
        float2 *dstg = out + b * ngood;
        if (NT == 256) {
            // pair 1 of the next block is requested now and used after all of pair 0's butterflies
            float2 nya[16], nyb[16];
            if (more) p1_load(nb, col_a(1), nya, nyb);
            asm volatile("" ::: "memory");
            {
                float2 ua[16], ub[16];
                p1_unload(0, ua, ub);
                p1_emit(ua, dstg + col_a(0), col_a(0), keep);
                p1_emit(ub, dstg + col_a(0) + 32, col_a(0) + 32, keep);
            }
            if (more) p1_forward(0, nxa, nxb);
            {
                float2 ua[16], ub[16];
                p1_unload(1, ua, ub);
                p1_emit(ua, dstg + col_a(1), col_a(1), keep);
                p1_emit(ub, dstg + col_a(1) + 32, col_a(1) + 32, keep);
            }
            if (more) p1_forward(1, nya, nyb);
        } else {
            // 128 registers: ua, ub, nxa, nxb are 32 each, so a column of the next block is requested only once a
            // column of this block has left
            const int ca = col_a(0);
            const float2 *src = in + nb * ngood + ca;
            // (one predicated form for whole and ragged blocks: splitting it into a CTA-uniform fast path with
            // unpredicated loads measured 5 % SLOWER, 133 vs 140 Gsamples/s -- the phase waits on memory, not on issue)
            const bool whole = more && n_in - nb * ngood >= (size_t)N;
            const int avail = more ? (whole ? N : (int)(n_in - nb * ngood)) : 0;
            float2 ua[16], ub[16];
            p1_unload(0, ua, ub);
            p1_emit(ua, dstg + ca, ca, keep);
            asm volatile("" ::: "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r) nxa[r] = (whole || ca + 1024 * r < avail) ? ldg_nc_f2(src + 1024 * r) : make_float2(0.f, 0.f);
            p1_emit(ub, dstg + ca + 32, ca + 32, keep);
            asm volatile("" ::: "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r)
                nxb[r] = (whole || ca + 32 + 1024 * r < avail) ? ldg_nc_f2(src + 1024 * r + 32) : make_float2(0.f, 0.f);
            if (more) p1_forward(0, nxa, nxb);
        }
    }
}

// H (natural order, already scaled by 1/nfft) -> Hq[k1][i][lane][c] = H[k1 + 16 (lane + 32 (2 i + c))]: what warp k1's
// lane holds in registers 2 i, 2 i + 1 after the forward transform, as one 16-byte word; a warp's load is 512 contiguous bytes
void lrc_fastfir16k_permute_H(const float2 *H, float2 *Hq)
{
    for (int k1 = 0; k1 < 16; ++k1)
        for (int i = 0; i < 16; ++i)
            for (int lane = 0; lane < 32; ++lane)
                for (int c = 0; c < 2; ++c)
                    Hq[((k1 * 16 + i) * 32 + lane) * 2 + c] = H[k1 + 16 * (lane + 32 * (2 * i + c))];
}

int lrc_fastfir16k_prepare(void)
{
    LRC_CUDA(cudaFuncSetAttribute(fastfir16k_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, ff16k::SMEM_BYTES));
    LRC_CUDA(cudaFuncSetAttribute(fastfir16k_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, ff16k::SMEM_BYTES));
    return LRC_OK;
}

int lrc_fastfir16k_launch(int n_sm, const float2 *in, size_t n_in, float2 *out, size_t full, size_t nblk, size_t ngood,
                          size_t keep, const float2 *d_tw16k, const float2 *d_tw1k, const float2 *d_Hq, cudaStream_t s)
{
    size_t blocks = (size_t)n_sm;                    // one CTA per SM (216.5 KB of shared memory)
    if (blocks > nblk) blocks = nblk;
    static const int nt = getenv("LRC_FASTFIR16K_THREADS") ? atoi(getenv("LRC_FASTFIR16K_THREADS")) : 512;
    const float4 *hq = reinterpret_cast<const float4 *>(d_Hq);
    if (nt == 256)
        fastfir16k_kernel<256><<<(unsigned)blocks, 256, ff16k::SMEM_BYTES, s>>>(in, n_in, out, full, nblk, ngood, keep, d_tw16k, d_tw1k, hq);
    else
        fastfir16k_kernel<512><<<(unsigned)blocks, 512, ff16k::SMEM_BYTES, s>>>(in, n_in, out, full, nblk, ngood, keep, d_tw16k, d_tw1k, hq);
    LRC_CUDA(cudaGetLastError());
    return LRC_OK;
}
