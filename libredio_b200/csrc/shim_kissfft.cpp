// shim_kissfft.cpp -> libkissfft.so : the three symbols LibRedio's kissfft crate binds
// (src/kissfft/src/kissfft.rs:11-16, link name "kissfft" from src/kissfft/build.rs:6), with the ABI of
// libkissfft/kiss_fft.h:81-114, served by the GPU FFT.  An unmodified LibRedio build that finds this
// library instead of the one `make -C libkissfft install` copies to /usr/local/lib runs its FFT block on
// the B200.  One PCIe round trip per kiss_fft() call: this is the compatibility seam, not the fast path
// (kpn_gpu::fft batches frames; lrc_fft_run works on device-resident batches).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "../../include/libredio_cuda.h"

extern "C" {

typedef struct { float r, i; } kiss_fft_cpx;           // kiss_fft.h:51-54 with kiss_fft_scalar = float (Makefile:4)

struct kiss_fft_state {                                  // opaque to callers; one malloc block so free() works
    unsigned magic;
    int nfft, inverse;
    lrc_fft *plan;                                       // device plan (leaks by design if the caller free()s the cfg)
};
typedef struct kiss_fft_state *kiss_fft_cfg;

static lrc_ctx *g_ctx = nullptr;
static std::once_flag g_once;

static lrc_ctx *shim_ctx()
{
    std::call_once(g_once, [] {
        const char *d = getenv("LIBREDIO_DEVICE");
        int rc = lrc_ctx_create(d ? atoi(d) : 0, &g_ctx);
        if (rc != LRC_OK) {
            // the reference cannot report errors from kiss_fft_alloc other than NULL; be loud
            fprintf(stderr, "libkissfft (libredio_b200 shim): %s [%s]\n", lrc_strerror(rc), lrc_last_error());
            g_ctx = nullptr;
        }
    });
    return g_ctx;
}

kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem)
{
    const size_t memneeded = sizeof(struct kiss_fft_state);
    kiss_fft_cfg st = nullptr;
    if (lenmem == nullptr) {
        st = (kiss_fft_cfg)malloc(memneeded);            // kiss_fft.c:345-346
    } else {
        if (mem != nullptr && *lenmem >= memneeded) st = (kiss_fft_cfg)mem;     // :348-350
        *lenmem = memneeded;
    }
    if (!st) return nullptr;
    lrc_ctx *ctx = shim_ctx();
    lrc_fft *plan = nullptr;
    if (!ctx || lrc_fft_create(ctx, nfft, inverse_fft, &plan) != LRC_OK) {
        fprintf(stderr, "libkissfft (libredio_b200 shim): kiss_fft_alloc(%d) failed: %s\n", nfft, lrc_last_error());
        if (lenmem == nullptr) free(st);
        return nullptr;
    }
    st->magic = 0x4b495353u; st->nfft = nfft; st->inverse = inverse_fft; st->plan = plan;
    return st;
}

void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout)
{
    if (!cfg || cfg->magic != 0x4b495353u) { fprintf(stderr, "kiss_fft: bad cfg\n"); abort(); }
    int rc = lrc_fft_run_host(cfg->plan, (const float *)fin, (float *)fout, (size_t)cfg->nfft);   // fin == fout allowed
    if (rc != LRC_OK) { fprintf(stderr, "kiss_fft: %s [%s]\n", lrc_strerror(rc), lrc_last_error()); abort(); }
}

void kiss_fft_stride(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout, int in_stride)
{
    if (in_stride == 1) { kiss_fft(cfg, fin, fout); return; }
    kiss_fft_cpx *tmp = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * (size_t)cfg->nfft);
    for (int k = 0; k < cfg->nfft; ++k) tmp[k] = fin[(size_t)k * in_stride];
    kiss_fft(cfg, tmp, fout);
    free(tmp);
}

void kiss_fft_cleanup(void) { /* nothing needed any more (kiss_fft.c:391-394) */ }

int kiss_fft_next_fast_size(int n)
{
    // the reference returns the next 2^a 3^b 5^c; this library only transforms powers of two
    int m = 1;
    while (m < n) m <<= 1;
    return m;
}

}  // extern "C"
