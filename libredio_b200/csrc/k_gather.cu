// k_gather.cu -- output gather over NVLink with the copy engines (the path shards into independent
// units; the ONLY inter-GPU traffic is moving each rank's output block to its peers).
//
// The reference has no multi-GPU story at all (threads + mpsc channels, kpn.rs:127-131); this is the piece that
// replaces "send the result Vec down the channel" when the producer blocks live on different GPUs.
//
// Why not ncclAllGather: the compute kernels are persistent and fill every SM (2 CTAs/SM of the chain kernel), so an
// NCCL kernel launched beside them only starts once they drain -- the gather serialises with the compute and costs
// 8 % of a step at 2 GPUs.  Here every rank owns a receive buffer [slot][src_rank][bytes_per_rank] that its peers
// map through CUDA IPC (one process per GPU) or plain peer access (several contexts in one process), and a push is
// `world` cudaMemcpyAsync D2D copies on a dedicated stream: copy engines over NVLink 5 / NVSwitch, zero SMs, fully
// overlapped with the next step's kernel.  Arrival is signalled by a 32-bit sequence number written into the
// receiver's flag word by the same stream (cuStreamWriteValue32, ordered after the data copy); consumers order a
// stream behind it with cuStreamWaitValue32 -- no host round trip, no collective.
#include "common.cuh"
#include <cuda.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>
#include <string>

constexpr int GATHER_MAX_WORLD = 64;
constexpr size_t GATHER_ALIGN = 256;

typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct lrc_gather {
    lrc_ctx     *ctx;
    int          rank, world, slots;
    int          root;                          // -1: every rank receives every block (all-gather); r: only rank r receives;
                                                // -2: the receiver rotates, push n of slot s lands on rank (n - 1 + s) % world
    size_t       bytes_per_rank, block_stride, slot_stride, flags_off, total_bytes;
    uint8_t     *base;                          // this rank's receive buffer (+ flag words)
    uint8_t     *peer[GATHER_MAX_WORLD];        // peers' receive buffers mapped into this process
    bool         opened[GATHER_MAX_WORLD];      // mapped with cudaIpcOpenMemHandle (must be closed)
    bool         connected;
    cudaStream_t push_stream;
    cudaEvent_t  ev_src;                        // producer stream -> push stream
    cudaEvent_t *ev_sent;                       // per slot: this rank's pushes have left the source buffer
    uint32_t    *seq;                           // per slot: pushes issued so far
    write32_fn   write32;
    wait32_fn    wait32;
    // host mode (lrc_gather_create_host): `base` is then the DEVICE alias of a page-locked POSIX shared-memory segment that
    // every rank of the node maps; pushes are D2H copies into it
    bool         is_host;
    uint8_t     *host_base;                     // this process's mapping of the segment
    std::string  shm_name;
    bool         shm_owner;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint8_t *slot_block(const lrc_gather *g, uint8_t *base, int slot, int src)
{
    return base + (size_t)slot * g->slot_stride + (size_t)src * g->block_stride;
}
static uint8_t *flag_word(const lrc_gather *g, uint8_t *base, int slot, int src)
{
    return base + g->flags_off + ((size_t)slot * g->world + src) * sizeof(uint32_t);
}

extern "C" int lrc_gather_create(lrc_ctx *ctx, int rank, int world, size_t bytes_per_rank, int slots, lrc_gather **out)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(out != nullptr, LRC_ERR_INVALID, "lrc_gather_create: null out");
    LRC_REQUIRE(world >= 1 && world <= GATHER_MAX_WORLD && rank >= 0 && rank < world, LRC_ERR_INVALID,
                "lrc_gather_create: need 0 <= rank < world <= 64");
    LRC_REQUIRE(slots >= 1 && slots <= 64 && bytes_per_rank > 0, LRC_ERR_INVALID,
                "lrc_gather_create: need 1 <= slots <= 64 and bytes_per_rank > 0");
    lrc_gather *g = new (std::nothrow) lrc_gather();
    LRC_REQUIRE(g != nullptr, LRC_ERR_NOMEM, "out of host memory");
    g->ctx = ctx; g->rank = rank; g->world = world; g->slots = slots; g->root = -1;
    g->is_host = false; g->host_base = nullptr; g->shm_owner = false;
    g->bytes_per_rank = bytes_per_rank;
    g->block_stride = align_up(bytes_per_rank, GATHER_ALIGN);
    g->slot_stride = g->block_stride * world;
    g->flags_off = g->slot_stride * slots;
    g->total_bytes = g->flags_off + align_up((size_t)slots * world * sizeof(uint32_t), GATHER_ALIGN);
    g->connected = false;
    for (int i = 0; i < GATHER_MAX_WORLD; ++i) { g->peer[i] = nullptr; g->opened[i] = false; }
    void *fw = nullptr, *fq = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &fw, cudaEnableDefault, &qr);
    if (e == cudaSuccess) e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fq, cudaEnableDefault, &qr);
    if (e != cudaSuccess || !fw || !fq) {
        delete g;
        lrc_set_error("lrc_gather_create: the driver does not export cuStreamWriteValue32/cuStreamWaitValue32");
        return LRC_ERR_UNSUPPORTED;
    }
    g->write32 = reinterpret_cast<write32_fn>(fw);
    g->wait32 = reinterpret_cast<wait32_fn>(fq);
    // plain cudaMalloc: IPC handles cannot be taken from pool / virtual-memory allocations
    e = cudaMalloc(reinterpret_cast<void **>(&g->base), g->total_bytes);
    if (e != cudaSuccess) {
        delete g;
        lrc_set_error("lrc_gather_create: cudaMalloc(%zu) -> %s", g->total_bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? LRC_ERR_NOMEM : LRC_ERR_CUDA;
    }
    // from here on a failure releases what exists: every member lrc_gather_destroy touches is initialised first
    g->push_stream = nullptr; g->ev_src = nullptr;
    g->ev_sent = new cudaEvent_t[slots]();
    g->seq = new uint32_t[slots]();
    e = cudaMemset(g->base, 0, g->total_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->push_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_src, cudaEventDisableTiming);
    for (int s = 0; s < slots && e == cudaSuccess; ++s) e = cudaEventCreateWithFlags(&g->ev_sent[s], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        lrc_set_error("lrc_gather_create: %s", cudaGetErrorString(e));
        lrc_gather_destroy(g);
        return LRC_ERR_CUDA;
    }
    g->peer[rank] = g->base;
    if (world == 1) g->connected = true;
    *out = g;
    return LRC_OK;
}

// Gather into HOST memory.  Measured at 8 GPUs (profiles/r2_u_bench_n8.json): rows pushed INTO a GPU that is running a
// bandwidth-bound kernel cost that kernel 4.7 % -- 13 us as soon as anything arrives plus the NVLink time of every 4 MB burst,
// with the peers idle, so it is the inbound traffic itself.  When the consumer of the rows is a CPU block anyway (the reference's
// PSD rows end in vidsink / psdpng, kpn blocks on the host), the rows need not visit another GPU at all: every rank copies its
// block D2H over ITS OWN PCIe link into one page-locked POSIX shared-memory segment [slot][src_rank][bytes] that all ranks of the
// node map, and raises its flag word there with the same stream-ordered 32-bit write; the root orders a stream behind the flags
// (lrc_gather_wait) or reads them from the CPU.  No GPU's HBM or NVLink port sees another rank's rows.
// Rank `root` creates the segment (an existing one of that name is replaced), the others open it (retrying for up to 30 s).
extern "C" int lrc_gather_create_host(lrc_ctx *ctx, int rank, int world, size_t bytes_per_rank, int slots, const char *shm_name,
                                      int root, lrc_gather **out)
{
    LRC_BIND(ctx);
    LRC_REQUIRE(out && shm_name && shm_name[0] == '/', LRC_ERR_INVALID, "lrc_gather_create_host: need out and a POSIX shm name \"/...\"");
    LRC_REQUIRE(world >= 1 && world <= GATHER_MAX_WORLD && rank >= 0 && rank < world && root >= 0 && root < world, LRC_ERR_INVALID,
                "lrc_gather_create_host: need 0 <= rank, root < world <= 64");
    LRC_REQUIRE(slots >= 1 && slots <= 64 && bytes_per_rank > 0, LRC_ERR_INVALID,
                "lrc_gather_create_host: need 1 <= slots <= 64 and bytes_per_rank > 0");
    lrc_gather *g = new (std::nothrow) lrc_gather();
    LRC_REQUIRE(g != nullptr, LRC_ERR_NOMEM, "out of host memory");
    g->ctx = ctx; g->rank = rank; g->world = world; g->slots = slots; g->root = root;
    g->is_host = true; g->host_base = nullptr; g->shm_name = shm_name; g->shm_owner = false; g->base = nullptr;
    g->bytes_per_rank = bytes_per_rank;
    g->block_stride = align_up(bytes_per_rank, GATHER_ALIGN);
    g->slot_stride = g->block_stride * world;
    g->flags_off = g->slot_stride * slots;
    g->total_bytes = align_up(g->flags_off + align_up((size_t)slots * world * sizeof(uint32_t), GATHER_ALIGN), 4096);
    g->connected = false;
    for (int i = 0; i < GATHER_MAX_WORLD; ++i) { g->peer[i] = nullptr; g->opened[i] = false; }
    g->push_stream = nullptr; g->ev_src = nullptr;
    g->ev_sent = new cudaEvent_t[slots]();
    g->seq = new uint32_t[slots]();
    void *fw = nullptr, *fq = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &fw, cudaEnableDefault, &qr);
    if (e == cudaSuccess) e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fq, cudaEnableDefault, &qr);
    if (e != cudaSuccess || !fw || !fq) {
        lrc_gather_destroy(g);
        lrc_set_error("lrc_gather_create_host: the driver does not export cuStreamWriteValue32/cuStreamWaitValue32");
        return LRC_ERR_UNSUPPORTED;
    }
    g->write32 = reinterpret_cast<write32_fn>(fw);
    g->wait32 = reinterpret_cast<wait32_fn>(fq);
    int fd = -1;
    if (rank == root) {
        shm_unlink(shm_name);
        fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)g->total_bytes) != 0) { close(fd); fd = -1; shm_unlink(shm_name); }
        g->shm_owner = fd >= 0;
    } else {
        for (int tries = 0; tries < 3000 && fd < 0; ++tries) {                 // the root may not have created it yet
            fd = shm_open(shm_name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) != 0 || (size_t)st.st_size != g->total_bytes) { close(fd); fd = -1; }
            }
            if (fd < 0) usleep(10000);
        }
    }
    if (fd < 0) {
        lrc_set_error("lrc_gather_create_host: cannot %s shared memory %s (%zu bytes)", rank == root ? "create" : "open", shm_name, g->total_bytes);
        lrc_gather_destroy(g);
        return LRC_ERR_UNSUPPORTED;
    }
    void *m = mmap(nullptr, g->total_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
        lrc_set_error("lrc_gather_create_host: mmap of %s failed", shm_name);
        lrc_gather_destroy(g);
        return LRC_ERR_NOMEM;
    }
    g->host_base = static_cast<uint8_t *>(m);
    if (rank == root) memset(g->host_base, 0, g->total_bytes);                 // flags start at 0 (ftruncate zero-fills; be explicit)
    e = cudaHostRegister(g->host_base, g->total_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    void *dp = nullptr;
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&dp, g->host_base, 0);
    if (e != cudaSuccess) {
        munmap(g->host_base, g->total_bytes); g->host_base = nullptr;          // never registered: destroy must not unregister
        lrc_set_error("lrc_gather_create_host: cudaHostRegister -> %s", cudaGetErrorString(e));
        lrc_gather_destroy(g);
        return LRC_ERR_CUDA;
    }
    g->base = static_cast<uint8_t *>(dp);
    e = cudaStreamCreateWithFlags(&g->push_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_src, cudaEventDisableTiming);
    for (int s = 0; s < slots && e == cudaSuccess; ++s) e = cudaEventCreateWithFlags(&g->ev_sent[s], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        lrc_set_error("lrc_gather_create_host: %s", cudaGetErrorString(e));
        lrc_gather_destroy(g);
        return LRC_ERR_CUDA;
    }
    g->connected = true;                                                       // nothing to map: the segment IS the connection
    *out = g;
    return LRC_OK;
}

extern "C" int lrc_gather_destroy(lrc_gather *g)
{
    if (!g) return LRC_OK;
    cudaSetDevice(g->ctx->device);
    if (g->push_stream) cudaStreamSynchronize(g->push_stream);
    for (int p = 0; p < g->world; ++p)
        if (g->opened[p]) cudaIpcCloseMemHandle(g->peer[p]);
    for (int s = 0; s < g->slots; ++s) if (g->ev_sent && g->ev_sent[s]) cudaEventDestroy(g->ev_sent[s]);
    if (g->ev_src) cudaEventDestroy(g->ev_src);
    if (g->push_stream) cudaStreamDestroy(g->push_stream);
    if (g->is_host) {
        if (g->host_base) {
            cudaHostUnregister(g->host_base);
            munmap(g->host_base, g->total_bytes);
        }
        if (g->shm_owner) shm_unlink(g->shm_name.c_str());
    } else {
        cudaFree(g->base);
    }
    delete[] g->ev_sent;
    delete[] g->seq;
    delete g;
    return LRC_OK;
}

extern "C" size_t lrc_gather_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int lrc_gather_export(lrc_gather *g, void *h_handle, size_t cap)
{
    LRC_REQUIRE(g && h_handle, LRC_ERR_INVALID, "lrc_gather_export: null argument");
    LRC_BIND(g->ctx);
    LRC_REQUIRE(cap >= sizeof(cudaIpcMemHandle_t), LRC_ERR_CAPACITY, "lrc_gather_export: handle buffer too small");
    LRC_REQUIRE(!g->is_host, LRC_ERR_INVALID, "lrc_gather_export: a host gather has nothing to export (the shared-memory name is the handle)");
    cudaIpcMemHandle_t h;
    LRC_CUDA(cudaIpcGetMemHandle(&h, g->base));
    memcpy(h_handle, &h, sizeof(h));
    return LRC_OK;
}

extern "C" int lrc_gather_connect(lrc_gather *g, const void *h_handles)
{
    LRC_REQUIRE(g && (h_handles || g->world == 1), LRC_ERR_INVALID, "lrc_gather_connect: null argument");
    LRC_BIND(g->ctx);
    LRC_REQUIRE(!g->connected || g->world == 1, LRC_ERR_INVALID, "lrc_gather_connect: already connected");
    const uint8_t *hb = static_cast<const uint8_t *>(h_handles);
    for (int p = 0; p < g->world; ++p) {
        if (p == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + (size_t)p * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        LRC_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        g->peer[p] = static_cast<uint8_t *>(ptr);
        g->opened[p] = true;
    }
    g->connected = true;
    return LRC_OK;
}

extern "C" int lrc_gather_connect_local(lrc_gather *g, lrc_gather *const *all)
{
    LRC_REQUIRE(g && all, LRC_ERR_INVALID, "lrc_gather_connect_local: null argument");
    LRC_BIND(g->ctx);
    for (int p = 0; p < g->world; ++p) {
        const lrc_gather *o = all[p];
        LRC_REQUIRE(o && o->world == g->world && o->rank == p && o->slots == g->slots &&
                    o->bytes_per_rank == g->bytes_per_rank, LRC_ERR_INVALID,
                    "lrc_gather_connect_local: peers must be created with the same geometry, all[p]->rank == p");
        if (o->ctx->device != g->ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            LRC_CUDA(e);
        }
        g->peer[p] = o->base;
    }
    g->connected = true;
    return LRC_OK;
}

// Who receives: every rank (root < 0, the default: an all-gather, `world - 1` outbound and inbound blocks per rank and
// push) or one rank only (a gather: one outbound block per rank, world - 1 inbound blocks at the root -- what a KPN graph
// whose consumer lives on one device needs; measured at 8 GPUs the all-gather's copy-engine traffic slows the HBM-bound
// chain kernel of EVERY rank by 4.2 %, profiles/r2_i_bench_n8.json).  SPMD: every rank sets the same root before its
// first push.  root = -2 (LRC_GATHER_ROTATE): the receiver rotates -- push n of slot s lands on rank (n - 1 + s) % world -- for
// graphs whose consumers are sharded like their producers: the same bytes cross NVLink, but every rank takes the inbound
// writes of one step in `world` (it is the RECEIVER's kernel that inbound P2P writes slow down, by 4.7 % for 28 MB a step).
extern "C" int lrc_gather_set_root(lrc_gather *g, int root)
{
    LRC_REQUIRE(g && root >= -2 && root < g->world, LRC_ERR_INVALID, "lrc_gather_set_root: need -2 <= root < world");
    LRC_REQUIRE(!g->is_host, LRC_ERR_INVALID, "lrc_gather_set_root: a host gather's receiver is fixed at creation");
    for (int s = 0; s < g->slots; ++s)
        LRC_REQUIRE(g->seq[s] == 0, LRC_ERR_INVALID, "lrc_gather_set_root: pushes were already issued");
    g->root = root;
    return LRC_OK;
}

// Order `stream` (the producer of the NEXT block that will live in the same source buffer) behind the copies of
// this rank's previous push from `slot`: the source buffer may be overwritten afterwards.
extern "C" int lrc_gather_wait_sent(lrc_gather *g, int slot, void *stream)
{
    LRC_REQUIRE(g && slot >= 0 && slot < g->slots, LRC_ERR_INVALID, "lrc_gather_wait_sent: bad slot");
    LRC_BIND(g->ctx);
    if (g->seq[slot] == 0) return LRC_OK;
    LRC_CUDA(cudaStreamWaitEvent(lrc_stream(g->ctx, stream), g->ev_sent[slot], 0));
    return LRC_OK;
}

extern "C" int lrc_gather_push(lrc_gather *g, int slot, const void *d_src, void *stream)
{
    LRC_REQUIRE(g && d_src && slot >= 0 && slot < g->slots, LRC_ERR_INVALID, "lrc_gather_push: bad argument");
    LRC_BIND(g->ctx);
    LRC_REQUIRE(g->connected, LRC_ERR_INVALID, "lrc_gather_push: peers are not connected yet");
    const uint32_t seq = ++g->seq[slot];
    LRC_CUDA(cudaEventRecord(g->ev_src, lrc_stream(g->ctx, stream)));
    LRC_CUDA(cudaStreamWaitEvent(g->push_stream, g->ev_src, 0));
    if (g->is_host) {
        // one D2H copy over this GPU's own PCIe link into the shared segment, then the flag (both through the device alias)
        LRC_CUDA(cudaMemcpyAsync(slot_block(g, g->host_base, slot, g->rank), d_src, g->bytes_per_rank, cudaMemcpyDeviceToHost,
                                 g->push_stream));
        const CUresult r = g->write32(reinterpret_cast<CUstream>(g->push_stream),
                                      reinterpret_cast<CUdeviceptr>(flag_word(g, g->base, slot, g->rank)), seq, 0);
        if (r != CUDA_SUCCESS) {
            lrc_set_error("lrc_gather_push: cuStreamWriteValue32 to the host segment -> CUresult %d", (int)r);
            return LRC_ERR_CUDA;
        }
        LRC_CUDA(cudaEventRecord(g->ev_sent[slot], g->push_stream));
        return LRC_OK;
    }
    for (int i = 0; i < g->world; ++i) {
        const int p = (g->rank + 1 + i) % g->world;          // start at the right-hand neighbour: spreads the NVSwitch ports
        if (g->root >= 0 && p != g->root) continue;
        if (g->root == -2 && p != (int)((seq - 1 + (uint32_t)slot) % (uint32_t)g->world)) continue;
        LRC_CUDA(cudaMemcpyAsync(slot_block(g, g->peer[p], slot, g->rank), d_src, g->bytes_per_rank,
                                 cudaMemcpyDeviceToDevice, g->push_stream));
        const CUresult r = g->write32(reinterpret_cast<CUstream>(g->push_stream),
                                      reinterpret_cast<CUdeviceptr>(flag_word(g, g->peer[p], slot, g->rank)), seq, 0);
        if (r != CUDA_SUCCESS) {
            lrc_set_error("lrc_gather_push: cuStreamWriteValue32 to rank %d -> CUresult %d", p, (int)r);
            return LRC_ERR_CUDA;
        }
    }
    LRC_CUDA(cudaEventRecord(g->ev_sent[slot], g->push_stream));
    return LRC_OK;
}

// Order `stream` behind the arrival of every rank's most recent push into this rank's `slot` (SPMD use: every rank
// pushes each slot the same number of times, so "most recent" is this rank's own push count for the slot).
extern "C" int lrc_gather_wait(lrc_gather *g, int slot, void *stream)
{
    LRC_REQUIRE(g && slot >= 0 && slot < g->slots, LRC_ERR_INVALID, "lrc_gather_wait: bad slot");
    LRC_BIND(g->ctx);
    const uint32_t seq = g->seq[slot];
    if (seq == 0) return LRC_OK;
    if (g->root >= 0 && g->rank != g->root) return LRC_OK;   // nothing is sent here
    if (g->root == -2 && g->rank != (int)((seq - 1 + (uint32_t)slot) % (uint32_t)g->world)) return LRC_OK;
    cudaStream_t s = lrc_stream(g->ctx, stream);
    for (int p = 0; p < g->world; ++p) {
        const CUresult r = g->wait32(reinterpret_cast<CUstream>(s),
                                     reinterpret_cast<CUdeviceptr>(flag_word(g, g->base, slot, p)), seq,
                                     CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) {
            lrc_set_error("lrc_gather_wait: cuStreamWaitValue32 -> CUresult %d", (int)r);
            return LRC_ERR_CUDA;
        }
    }
    return LRC_OK;
}

// Host gather only: block the CALLING CPU THREAD until every rank's latest push into `slot` has arrived (the flag words live in
// the shared host segment; a push writes its flag after its copy in stream order), or until timeout_ms have passed
// (LRC_ERR_CAPACITY then: nothing is wrong yet, the rows are just not there).  For consumer blocks that run on the CPU and have
// no CUDA stream to order behind the flags.
extern "C" int lrc_gather_wait_host(lrc_gather *g, int slot, unsigned timeout_ms)
{
    LRC_REQUIRE(g && slot >= 0 && slot < g->slots, LRC_ERR_INVALID, "lrc_gather_wait_host: bad slot");
    LRC_REQUIRE(g->is_host, LRC_ERR_INVALID, "lrc_gather_wait_host: not a host gather");
    const uint32_t seq = g->seq[slot];
    if (seq == 0) return LRC_OK;
    const volatile uint32_t *flags = reinterpret_cast<const volatile uint32_t *>(flag_word(g, g->host_base, slot, 0));
    struct timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned spins = 0;; ++spins) {
        bool all = true;
        for (int p = 0; p < g->world; ++p) all = all && (int32_t)(flags[p] - seq) >= 0;
        if (all) { __sync_synchronize(); return LRC_OK; }
        if ((spins & 1023u) == 1023u) {
            struct timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
            if (ms > (double)timeout_ms) {
                lrc_set_error("lrc_gather_wait_host: slot %d: not every rank's push %u arrived within %u ms", slot, seq, timeout_ms);
                return LRC_ERR_CAPACITY;
            }
            usleep(50);
        }
    }
}

extern "C" int lrc_gather_buffer(lrc_gather *g, int slot, void **d_ptr, size_t *block_stride)
{
    LRC_REQUIRE(g && d_ptr && slot >= 0 && slot < g->slots, LRC_ERR_INVALID, "lrc_gather_buffer: bad argument");
    *d_ptr = (g->is_host ? g->host_base : g->base) + (size_t)slot * g->slot_stride;     // host mode: a HOST pointer
    if (block_stride) *block_stride = g->block_stride;
    return LRC_OK;
}
