// TEMPORARY: entry points not implemented yet fail loudly (replaced file by file).
#include "common.cuh"
#define TODO(name) { lrc_set_error(#name ": not implemented yet"); return LRC_ERR_UNSUPPORTED; }
extern "C" {
int    lrc_fastfir_create(lrc_ctx *, const float *, size_t, size_t, lrc_fastfir **) TODO(lrc_fastfir_create)
int    lrc_fastfir_destroy(lrc_fastfir *) { return LRC_OK; }
size_t lrc_fastfir_nfft(const lrc_fastfir *) { return 0; }
size_t lrc_fastfir_out_len(const lrc_fastfir *, size_t, int) { return 0; }
int    lrc_fastfir_run(lrc_fastfir *, const float *, size_t, float *, int, size_t *, void *) TODO(lrc_fastfir_run)
int    lrc_ook_create(lrc_ctx *, size_t, size_t, unsigned, size_t, size_t, lrc_ook **) TODO(lrc_ook_create)
int    lrc_ook_destroy(lrc_ook *) { return LRC_OK; }
int    lrc_ook_decode(lrc_ook *, const uint8_t *, size_t, void *) TODO(lrc_ook_decode)
int    lrc_ook_fetch_packets(lrc_ook *, lrc_ook_packet *, size_t, size_t *) TODO(lrc_ook_fetch_packets)
int    lrc_ook_debug_ptrs(lrc_ook *, const float **, const uint32_t **, const uint32_t **, const uint32_t **) TODO(lrc_ook_debug_ptrs)
int    lrc_eat(const uint8_t *, size_t, const size_t *, size_t, size_t *) TODO(lrc_eat)
int    lrc_ook_envelope_table(lrc_ctx *, float *, void *) TODO(lrc_ook_envelope_table)
}
