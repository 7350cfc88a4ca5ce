// shim_samplerate.cpp -> libsamplerate.so : the symbols LibRedio's samplerate crate declares
// (src/samplerate/src/samplerate.rs:32-42: src_new, src_delete, src_process, src_get_name,
// src_get_description, src_get_version, src_set_ratio, src_is_valid_ratio, src_strerror), with the C ABI
// of libsamplerate's public header, served by the GPU polyphase resampler.
//
// NOT libsamplerate's arithmetic: the filter is ours (DESIGN.md "resampler", parity unpinned).  The
// struct below is the real C layout (long frames counts); the reference's Rust struct uses u64 for C
// `long` without #[repr(C)] (samplerate.rs:15-24) -- identical on LP64 Linux, which is all that runs here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "../../include/libredio_cuda.h"

extern "C" {

typedef struct {
    const float *data_in;
    float       *data_out;
    long         input_frames, output_frames;
    long         input_frames_used, output_frames_gen;
    int          end_of_input;
    double       src_ratio;
} SRC_DATA;

enum { SRC_ERR_NO_ERROR = 0, SRC_ERR_MALLOC_FAILED = 1, SRC_ERR_BAD_STATE = 2, SRC_ERR_BAD_DATA = 3,
       SRC_ERR_BAD_DATA_PTR = 4, SRC_ERR_BAD_SRC_RATIO = 6, SRC_ERR_BAD_CONVERTER = 10,
       SRC_ERR_BAD_CHANNEL_COUNT = 11, SRC_ERR_GPU = 100, SRC_ERR_RATIO_CHANGE = 101 };

struct SRC_STATE {
    unsigned magic;
    int channels, converter, last_error;
    double ratio;
    lrc_resampler *rs;
    size_t max_chunk;
};

static lrc_ctx *g_ctx = nullptr;
static std::once_flag g_once;
static lrc_ctx *shim_ctx()
{
    std::call_once(g_once, [] {
        const char *d = getenv("LIBREDIO_DEVICE");
        if (lrc_ctx_create(d ? atoi(d) : 0, &g_ctx) != LRC_OK) {
            fprintf(stderr, "libsamplerate (libredio_b200 shim): %s\n", lrc_last_error());
            g_ctx = nullptr;
        }
    });
    return g_ctx;
}

SRC_STATE *src_new(int converter_type, int channels, int *error)
{
    if (error) *error = SRC_ERR_NO_ERROR;
    if (channels != 1) { if (error) *error = SRC_ERR_BAD_CHANNEL_COUNT; return nullptr; }   // the reference uses 1 (:61)
    if (converter_type < 0 || converter_type > 4) { if (error) *error = SRC_ERR_BAD_CONVERTER; return nullptr; }
    if (!shim_ctx()) { if (error) *error = SRC_ERR_GPU; return nullptr; }
    SRC_STATE *st = (SRC_STATE *)calloc(1, sizeof(SRC_STATE));
    if (!st) { if (error) *error = SRC_ERR_MALLOC_FAILED; return nullptr; }
    st->magic = 0x53524331u; st->channels = channels; st->converter = converter_type;
    return st;
}

SRC_STATE *src_delete(SRC_STATE *st)
{
    if (st && st->magic == 0x53524331u) { lrc_resampler_destroy(st->rs); st->magic = 0; free(st); }
    return nullptr;
}

int src_is_valid_ratio(double ratio) { return (ratio >= 1.0 / 256.0 && ratio <= 256.0) ? 1 : 0; }

int src_process(SRC_STATE *st, SRC_DATA *data)
{
    if (!st || st->magic != 0x53524331u) return SRC_ERR_BAD_STATE;
    if (!data) return SRC_ERR_BAD_DATA;
    if ((!data->data_in && data->input_frames > 0) || (!data->data_out && data->output_frames > 0)) return SRC_ERR_BAD_DATA_PTR;
    if (!src_is_valid_ratio(data->src_ratio)) return SRC_ERR_BAD_SRC_RATIO;
    if (data->input_frames < 0) data->input_frames = 0;
    if (data->output_frames < 0) data->output_frames = 0;
    data->input_frames_used = data->output_frames_gen = 0;
    const size_t n_in_all = (size_t)data->input_frames;
    if (st->rs && std::fabs(st->ratio - data->src_ratio) > 1e-12 * st->ratio) return st->last_error = SRC_ERR_RATIO_CHANGE;
    if (!st->rs) {
        // libsamplerate puts no bound on input_frames and a caller may grow its chunks at will (samplerate.rs:64-84 passes
        // whatever Vec arrives): lrc_resampler sizes nothing from max_chunk, it only guards against a wild length
        st->max_chunk = (size_t)1 << 40;
        st->ratio = data->src_ratio;
        if (lrc_resampler_create(shim_ctx(), st->ratio, 1, st->max_chunk, &st->rs) != LRC_OK) {
            st->rs = nullptr;
            return st->last_error = SRC_ERR_BAD_SRC_RATIO;
        }
    }
    // take as many input frames as the output buffer can absorb
    size_t n_in = n_in_all;
    const size_t cap = (size_t)data->output_frames;
    if (lrc_resampler_next_out_len(st->rs, n_in) > cap) {
        size_t lo = 0, hi = n_in;                                 // largest n with out_len(n) <= cap
        while (lo < hi) {
            const size_t mid = (lo + hi + 1) / 2;
            if (lrc_resampler_next_out_len(st->rs, mid) <= cap) lo = mid; else hi = mid - 1;
        }
        n_in = lo;
    }
    size_t n_out = 0;
    if (n_in) {
        int rc = lrc_resampler_process_host(st->rs, data->data_in, n_in, data->data_out, cap ? cap : 1, &n_out);
        if (rc != LRC_OK) { fprintf(stderr, "src_process: %s\n", lrc_last_error()); return st->last_error = SRC_ERR_GPU; }
    }
    data->input_frames_used = (long)n_in;
    data->output_frames_gen = (long)n_out;
    return SRC_ERR_NO_ERROR;
}

int src_set_ratio(SRC_STATE *st, double new_ratio)
{
    if (!st || st->magic != 0x53524331u) return SRC_ERR_BAD_STATE;
    if (!src_is_valid_ratio(new_ratio)) return SRC_ERR_BAD_SRC_RATIO;
    if (st->rs && std::fabs(st->ratio - new_ratio) > 1e-12 * st->ratio) return SRC_ERR_RATIO_CHANGE;
    return SRC_ERR_NO_ERROR;
}

int src_reset(SRC_STATE *st)
{
    if (!st || st->magic != 0x53524331u) return SRC_ERR_BAD_STATE;
    if (st->rs && lrc_resampler_reset(st->rs) != LRC_OK) return SRC_ERR_GPU;
    return SRC_ERR_NO_ERROR;
}

int src_error(SRC_STATE *st) { return st ? st->last_error : SRC_ERR_BAD_STATE; }

const char *src_strerror(int error)
{
    switch (error) {
        case SRC_ERR_NO_ERROR: return "No error.";
        case SRC_ERR_MALLOC_FAILED: return "Malloc failed.";
        case SRC_ERR_BAD_STATE: return "SRC_STATE pointer is NULL.";
        case SRC_ERR_BAD_DATA: return "SRC_DATA pointer is NULL.";
        case SRC_ERR_BAD_DATA_PTR: return "SRC_DATA->data_out or SRC_DATA->data_in is NULL.";
        case SRC_ERR_BAD_SRC_RATIO: return "SRC ratio outside [1/256, 256] range (or not L/M with L, M <= 4096 on this GPU build).";
        case SRC_ERR_BAD_CONVERTER: return "Bad converter number.";
        case SRC_ERR_BAD_CHANNEL_COUNT: return "Channel count must be 1 on this GPU build.";
        case SRC_ERR_GPU: return "libredio_b200: CUDA failure (no CPU fallback).";
        case SRC_ERR_RATIO_CHANGE: return "libredio_b200: time-varying ratio is not implemented.";
    }
    return nullptr;
}

const char *src_get_name(int converter_type)
{
    return (converter_type >= 0 && converter_type <= 4) ? "libredio_b200 GPU polyphase sinc" : nullptr;
}
const char *src_get_description(int converter_type)
{
    return (converter_type >= 0 && converter_type <= 4)
               ? "Kaiser-windowed sinc, 32 zero crossings, 90% bandwidth, rational L/M polyphase on sm_100a" : nullptr;
}
const char *src_get_version(void) { return "libredio_b200-samplerate-shim 1.0 (not libsamplerate)"; }

}  // extern "C"
