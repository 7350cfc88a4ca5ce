// k_fastfir16k.cu -- overlap-save with nfft = 16384.  STAGED: written against the verified numpy model
// tools/models/fastfir16k_model.py and checked ONCE on a B200 at the very end of round 1 (tools/staged_fastfir16k_check.py,
// profiles/r1_s8_fastfir16k_check.json: max error 6.2e-7 x RMS on windows incl. block seams and the flush block,
// 121.0 Gsamples/s against 100.2 for fastfir8k_kernel on 4096 taps), but the pytest cases for it have not run on
// hardware yet, so lrc_fastfir_create only builds this plan when LRC_FASTFIR_STAGED=1 is set.
//
// Why: with nh = 4096 an 8192-point block keeps 4097 of 8192 outputs (50 %), a 16384-point block keeps 12289 of 16384
// (75 %): 1.39x fewer butterflies per output sample (kiss_fastfir.c:81-93 picks the block size by a cost model of
// its own; the result is the same convolution whatever the block size).
//
//   16384 = 16 x 1024,  n = 1024 n1 + n2,  k = k1 + 16 k2
//   P1   column n2 (one thread, two columns per thread): DFT16 over n1 straight from global memory, times
//        W_16384^(n2 k1), stored to shared row k1 at padded position n2 + n2/32
//   P2   warp k1 owns row k1: 1024 points as 32 x 32 in registers (WarpFFT1024, the row buffer itself is the
//        exchange tile), .* Hp[k1][k2] = H[k1 + 16 k2], inverse transform, back into the row
//   P1'  column n2: times conj(W^(n2 k1)), IDFT16 over k1, outputs 1024 n1 + n2 < ngood to global memory; fused per
//        column with P1 of the CTA's next block (its loads are in flight during the inverse butterfly) -- no barrier
//        between the two blocks' P1 phases because a column is private to its thread in both.
// Two CTA barriers per block.  Shared memory: 16 rows x 1056 complex (132 KB) + W_16384^(n2 k1) for n2 < 512
// (60 KB; the upper half is a W_32^k1 constant away) + the two 31 x 32 warp-transform twiddle tables: 207.5 KB,
// one 512-thread CTA per SM.
#include "fft_core.cuh"

using namespace lrfft;

namespace ff16k {
constexpr int N = 16384, NT = 512, ROWS = 16, PITCH = 33 * 32;
constexpr int DATA_CPX = ROWS * PITCH;                      // 16896
constexpr int TW1_CPX = 15 * 512;                           // [k1 - 1][n2], n2 < 512 : W_16384^(n2 k1)
constexpr int TWW_CPX = WarpFFT1024<false>::TW_CPX;         // 31 * 32
constexpr int SMEM_BYTES = (DATA_CPX + TW1_CPX + 2 * TWW_CPX) * 8;
__host__ __device__ constexpr int pad(int n2) { return n2 + (n2 >> 5); }

// exp(-2 pi j k / 32), k < 16 (compile-time index after unrolling)
__device__ __forceinline__ float2 w32(int k)
{
    constexpr float C[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                             0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                             -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                             -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float S[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                             0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                             0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                             0.38268343236508977173f, 0.19509032201612826785f};
    return make_float2(C[k], -S[k]);
}
}  // namespace ff16k

__global__ void __launch_bounds__(ff16k::NT, 1)
fastfir16k_kernel(const float2 *__restrict__ in, size_t n_in, float2 *__restrict__ out, size_t n_blocks_full,
                  size_t n_blocks, size_t ngood, size_t flush_keep, const float2 *__restrict__ tw /* W_16384^k */,
                  const float2 *__restrict__ tw1k /* W_1024^k */, const float2 *__restrict__ Hp)
{
    using namespace ff16k;
    extern __shared__ __align__(16) float2 ff16k_smem[];
    float2 *sd = ff16k_smem;                         // [16][PITCH]
    float2 *tw1 = sd + DATA_CPX;                     // [15][512]
    float2 *twf = tw1 + TW1_CPX;                     // warp transform, forward
    float2 *twi = twf + TWW_CPX;                     // warp transform, inverse (conjugated)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < TW1_CPX; i += NT) tw1[i] = __ldg(tw + ((i >> 9) + 1) * (i & 511));     // 15 * 511 < 16384
    WarpFFT1024<false>::fill_twiddles(tw1k, twf);
    WarpFFT1024<true>::fill_twiddles(tw1k, twi);
    __syncthreads();

    auto p1_load = [&](size_t blk, int j, float2 *v) {
        const size_t s0 = blk * ngood, avail = n_in - s0;        // avail < N only for the flush block
        const float2 *src = in + s0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const size_t i = (size_t)(j + 1024 * r);
            v[r] = i < avail ? __ldg(src + i) : make_float2(0.f, 0.f);
        }
    };
    // forward butterfly of column j = t + 512 h; results into the column's 16 shared-memory slots
    auto p1_forward = [&](int h, float2 *v) {
        RegFFT<16, false>::run(v);
        float2 *dst = sd + pad(t + 512 * h);             // row k1 lives PITCH entries further
        dst[0] = v[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) {
            float2 w = tw1[(q - 1) * 512 + t];
            if (h) w = cmulf(w, w32(q));                 // W_16384^(512 q) = W_32^q
            dst[PITCH * q] = cmulf(v[q], w);
        }
    };

    size_t b = blockIdx.x;
    if (b < n_blocks) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            float2 v[16];
            p1_load(b, t + 512 * h, v);
            p1_forward(h, v);
        }
    }
    for (; b < n_blocks; b += gridDim.x) {
        __syncthreads();
        // ---- P2: warp `warp` transforms row `warp`, multiplies by its slice of H, transforms back -----------
        {
            float2 *row = sd + PITCH * warp;
            const float2 *hp = Hp + 1024 * warp + lane;
            float2 v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = row[lane + 33 * e];
            WarpFFT1024<false>::run(v, row, twf, lane);          // X[warp + 16 (lane + 32 e)]
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = cmulf(v[e], __ldg(hp + 32 * e));      // C_MUL  kiss_fastfir.c:180-184
            WarpFFT1024<true>::run(v, row, twi, lane);
            __syncwarp();                                        // the tile reads of the last exchange are done
#pragma unroll
            for (int e = 0; e < 32; ++e) row[lane + 33 * e] = v[e];
        }
        __syncthreads();
        // ---- P1 inverse of this block fused per column with P1 forward of the CTA's next block -------------
        const size_t nb = b + gridDim.x;
        const bool more = nb < n_blocks;
        const size_t s0 = b * ngood;
        const size_t keep = (b < n_blocks_full) ? ngood : flush_keep;
        float2 *dstg = out + s0;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int j = t + 512 * h;
            float2 nx[16];
            if (more) p1_load(nb, j, nx);
            {
                const float2 *p = sd + pad(j);
                float2 v[16];
                v[0] = p[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) {
                    float2 w = tw1[(q - 1) * 512 + t];
                    if (h) w = cmulf(w, w32(q));
                    v[q] = cmul_conjb(p[PITCH * q], w);
                }
                RegFFT<16, true>::run(v);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const size_t i = (size_t)(j + 1024 * r);
                    if (i < keep) __stcs(dstg + i, v[r]);
                }
            }
            if (more) p1_forward(h, nx);
        }
    }
}

// H (natural order, already scaled by 1/nfft) -> Hp[k1][k2] = H[k1 + 16 k2]
void lrc_fastfir16k_permute_H(const float2 *H, float2 *Hp)
{
    for (int k1 = 0; k1 < 16; ++k1)
        for (int k2 = 0; k2 < 1024; ++k2) Hp[k1 * 1024 + k2] = H[k1 + 16 * k2];
}

int lrc_fastfir16k_prepare(void)
{
    LRC_CUDA(cudaFuncSetAttribute(fastfir16k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ff16k::SMEM_BYTES));
    return LRC_OK;
}

int lrc_fastfir16k_launch(int n_sm, const float2 *in, size_t n_in, float2 *out, size_t full, size_t nblk, size_t ngood,
                          size_t keep, const float2 *d_tw16k, const float2 *d_tw1k, const float2 *d_Hp, cudaStream_t s)
{
    size_t blocks = (size_t)n_sm;                    // one 512-thread CTA per SM (207.5 KB of shared memory)
    if (blocks > nblk) blocks = nblk;
    fastfir16k_kernel<<<(unsigned)blocks, ff16k::NT, ff16k::SMEM_BYTES, s>>>(in, n_in, out, full, nblk, ngood, keep, d_tw16k,
                                                                           d_tw1k, d_Hp);
    LRC_CUDA(cudaGetLastError());
    return LRC_OK;
}
