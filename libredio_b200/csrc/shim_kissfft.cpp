// shim_kissfft.cpp -> libkissfft.so : the three symbols LibRedio's kissfft crate binds
// (src/kissfft/src/kissfft.rs:11-16, link name "kissfft" from src/kissfft/build.rs:6), with the ABI of
// libkissfft/kiss_fft.h:81-114, served by the GPU FFT.  An unmodified LibRedio build that finds this
// library instead of the one `make -C libkissfft install` copies to /usr/local/lib runs its FFT block on
// the B200.  One PCIe round trip per kiss_fft() call: this is the compatibility seam, not the fast path
// (kpn_gpu::fft batches frames; lrc_fft_run works on device-resident batches).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include "../../include/libredio_cuda.h"

// Device plans are shared and cached per (nfft, direction, real?): a kiss_fft_cfg is ONE malloc block that callers release
// with free() (kiss_fft.h:102; LibRedio's Rust never releases it at all, kissfft.rs:19), so the block itself may only
// hold a pointer into this cache -- alloc/free cycles then cost no device memory, whatever the caller does with the cfg.
// The cache is bounded by the number of distinct sizes a process asks for.  A plan owns scratch, so calls through one plan
// are serialised by its mutex (kiss_fft is documented thread-safe, README:103; different sizes still run concurrently).
namespace {
struct CachedPlan { void *plan = nullptr; std::mutex mu; };
std::mutex g_cache_mu;
std::map<std::pair<int, int>, CachedPlan *> g_cache;      // key: (nfft, inverse | real << 1)
}

extern "C" {

typedef struct { float r, i; } kiss_fft_cpx;           // kiss_fft.h:51-54 with kiss_fft_scalar = float (Makefile:4)

struct kiss_fft_state {                                  // opaque to callers; one malloc block so free() works
    unsigned magic;
    int nfft, inverse;
    CachedPlan *entry;                                   // shared device plan (shim-owned: free(cfg) leaks nothing)
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

static CachedPlan *cached_plan(int nfft, int inverse, int real)
{
    lrc_ctx *ctx = shim_ctx();
    if (!ctx) return nullptr;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const std::pair<int, int> key(nfft, (inverse ? 1 : 0) | (real ? 2 : 0));
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    void *plan = nullptr;
    const int rc = real ? lrc_rfft_create(ctx, nfft, inverse, (lrc_rfft **)&plan) : lrc_fft_create(ctx, nfft, inverse, (lrc_fft **)&plan);
    if (rc != LRC_OK) return nullptr;
    CachedPlan *e = new CachedPlan();
    e->plan = plan;
    g_cache[key] = e;
    return e;
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
    CachedPlan *e = cached_plan(nfft, inverse_fft, 0);
    if (!e) {
        fprintf(stderr, "libkissfft (libredio_b200 shim): kiss_fft_alloc(%d) failed: %s\n", nfft, lrc_last_error());
        if (lenmem == nullptr) free(st);
        return nullptr;
    }
    st->magic = 0x4b495353u; st->nfft = nfft; st->inverse = inverse_fft; st->entry = e;
    return st;
}

void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout)
{
    if (!cfg || cfg->magic != 0x4b495353u) { fprintf(stderr, "kiss_fft: bad cfg\n"); abort(); }
    std::lock_guard<std::mutex> lk(cfg->entry->mu);
    int rc = lrc_fft_run_host((lrc_fft *)cfg->entry->plan, (const float *)fin, (float *)fout, (size_t)cfg->nfft);   // fin == fout allowed
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
    // kiss_fft.c:396-408: the next size whose only factors are 2, 3 and 5 (all run on the GPU up to 8192)
    while (1) {
        int m = n;
        while ((m % 2) == 0) m /= 2;
        while ((m % 3) == 0) m /= 3;
        while ((m % 5) == 0) m /= 5;
        if (m <= 1) break;
        n++;
    }
    return n;
}

// ---- tools/kiss_fftr.h:21-43: real-input pair (used by psdpng and kiss_fastfir's real build) ----------
struct kiss_fftr_state {
    unsigned magic;
    int nfft, inverse;
    CachedPlan *entry;
};
typedef struct kiss_fftr_state *kiss_fftr_cfg;

kiss_fftr_cfg kiss_fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem)
{
    if (nfft & 1) {
        fprintf(stderr, "Real FFT optimization must be even.\n");   // kiss_fftr.c:34-37
        return nullptr;
    }
    const size_t memneeded = sizeof(struct kiss_fftr_state);
    kiss_fftr_cfg st = nullptr;
    if (lenmem == nullptr) {
        st = (kiss_fftr_cfg)malloc(memneeded);
    } else {
        if (mem != nullptr && *lenmem >= memneeded) st = (kiss_fftr_cfg)mem;
        *lenmem = memneeded;
    }
    if (!st) return nullptr;
    CachedPlan *e = cached_plan(nfft, inverse_fft, 1);
    if (!e) {
        fprintf(stderr, "libkissfft (libredio_b200 shim): kiss_fftr_alloc(%d) failed: %s\n", nfft, lrc_last_error());
        if (lenmem == nullptr) free(st);
        return nullptr;
    }
    st->magic = 0x4b465452u; st->nfft = nfft; st->inverse = inverse_fft; st->entry = e;
    return st;
}

void kiss_fftr(kiss_fftr_cfg cfg, const float *timedata, kiss_fft_cpx *freqdata)
{
    if (!cfg || cfg->magic != 0x4b465452u) { fprintf(stderr, "kiss_fftr: bad cfg\n"); abort(); }
    if (cfg->inverse) { fprintf(stderr, "kiss fft usage error: improper alloc\n"); exit(1); }   // kiss_fftr.c:73-76
    std::lock_guard<std::mutex> lk(cfg->entry->mu);
    int rc = lrc_rfft_run_host((lrc_rfft *)cfg->entry->plan, timedata, (float *)freqdata, 1);
    if (rc != LRC_OK) { fprintf(stderr, "kiss_fftr: %s [%s]\n", lrc_strerror(rc), lrc_last_error()); abort(); }
}

void kiss_fftri(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, float *timedata)
{
    if (!cfg || cfg->magic != 0x4b465452u) { fprintf(stderr, "kiss_fftri: bad cfg\n"); abort(); }
    if (!cfg->inverse) { fprintf(stderr, "kiss fft usage error: improper alloc\n"); exit(1); }  // kiss_fftr.c:126-129
    std::lock_guard<std::mutex> lk(cfg->entry->mu);
    int rc = lrc_rfft_run_host((lrc_rfft *)cfg->entry->plan, (const float *)freqdata, timedata, 1);
    if (rc != LRC_OK) { fprintf(stderr, "kiss_fftri: %s [%s]\n", lrc_strerror(rc), lrc_last_error()); abort(); }
}

}  // extern "C"
