// ctx.cu -- context, status strings, pinned host memory.
#include "common.cuh"

static thread_local char g_err[512] = "";

void lrc_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int lrc_version(void) { return LRC_VERSION; }

extern "C" const char *lrc_last_error(void) { return g_err; }

extern "C" const char *lrc_strerror(int status)
{
    switch (status) {
        case LRC_OK:              return "ok";
        case LRC_ERR_INVALID:     return "invalid argument";
        case LRC_ERR_CUDA:        return "CUDA error (no usable GPU or a runtime failure); there is no CPU fallback";
        case LRC_ERR_UNSUPPORTED: return "valid for the reference but not implemented by this library";
        case LRC_ERR_NOMEM:       return "out of memory";
        case LRC_ERR_CAPACITY:    return "output capacity too small";
        case LRC_ERR_ODD_LENGTH:  return "odd byte count into the u8 IQ unpack";
        case LRC_ERR_LENGTH:      return "frame length does not match block_size";
    }
    return "unknown status";
}

extern "C" int lrc_ctx_create(int device, lrc_ctx **out)
{
    LRC_REQUIRE(out != nullptr, LRC_ERR_INVALID, "lrc_ctx_create: null out");
    int n = 0;
    LRC_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        lrc_set_error("lrc_ctx_create: device %d out of range (%d visible)", device, n);
        return LRC_ERR_INVALID;
    }
    LRC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LRC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        lrc_set_error("lrc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                      device, prop.major, prop.minor);
        return LRC_ERR_UNSUPPORTED;
    }
    lrc_ctx *c = new (std::nothrow) lrc_ctx();
    LRC_REQUIRE(c != nullptr, LRC_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    LRC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    LRC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LRC_CUDA(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
    *out = c;
    return LRC_OK;
}

extern "C" int lrc_ctx_destroy(lrc_ctx *c)
{
    if (!c) return LRC_OK;
    cudaSetDevice(c->device);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->out_stream);
    delete c;
    return LRC_OK;
}

extern "C" int lrc_ctx_sync(lrc_ctx *c)
{
    LRC_BIND(c);
    LRC_CUDA(cudaStreamSynchronize(c->stream));
    return LRC_OK;
}

extern "C" int lrc_ctx_sm_count(lrc_ctx *c, int *n_sm)
{
    LRC_REQUIRE(c && n_sm, LRC_ERR_INVALID, "null argument");
    *n_sm = c->n_sm;
    return LRC_OK;
}

extern "C" int lrc_host_alloc(lrc_ctx *c, size_t bytes, void **h_ptr)
{
    LRC_BIND(c);
    LRC_REQUIRE(h_ptr != nullptr, LRC_ERR_INVALID, "null out pointer");
    LRC_CUDA(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return LRC_OK;
}

extern "C" int lrc_host_free(lrc_ctx *c, void *h_ptr)
{
    LRC_BIND(c);
    if (h_ptr) LRC_CUDA(cudaFreeHost(h_ptr));
    return LRC_OK;
}

extern "C" int lrc_copy_to_host(lrc_ctx *c, void *h_dst, const void *d_src, size_t bytes)
{
    LRC_BIND(c);
    if (bytes == 0) return LRC_OK;
    LRC_REQUIRE(h_dst && d_src, LRC_ERR_INVALID, "lrc_copy_to_host: null pointer");
    LRC_CUDA(cudaDeviceSynchronize());
    LRC_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return LRC_OK;
}
