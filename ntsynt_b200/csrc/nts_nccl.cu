// nts_nccl.cu -- the one inter-GPU exchange of the path: merging per-GPU Bloom filters.
//
// NCCL has no bitwise reduce op, so the bit array is merged as COUNTERS: a chunk of bits is expanded
// to c-bit fields (c = 2, 4 or 8, the smallest with 2^c > world) packed in uint32 lanes -- no carry can
// cross a field because every rank contributes 0 or 1 -- summed with ONE ncclAllReduce(ncclSum) per
// chunk, and thresholded back to bits: AND <=> count == world, OR <=> count > 0.  Chunks are double
// buffered on two streams so expand/threshold of one chunk overlaps the all-reduce of the other.
// NCCL is dlopen()ed at run time (libnccl.so.2, the copy torch already mapped if bench.py imported
// torch), so the library has no link-time NCCL dependency.  The reference has nothing like this: it is
// a single process (src/ntsynt_make_common_bf.cpp:136-160 cascades genome after genome).
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <new>

#include <nccl.h>

#include "nts_internal.h"

namespace nts {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.handle) return NTS_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(NTS_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define NTS_SYM(field, name)                                                                     \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                                   \
    if (!g_nccl.field) return fail(NTS_ERR_NCCL, std::string("libnccl is missing ") + name)
    NTS_SYM(GetUniqueId, "ncclGetUniqueId");
    NTS_SYM(CommInitRank, "ncclCommInitRank");
    NTS_SYM(CommDestroy, "ncclCommDestroy");
    NTS_SYM(AllReduce, "ncclAllReduce");
    NTS_SYM(AllGather, "ncclAllGather");
    NTS_SYM(GetErrorString, "ncclGetErrorString");
#undef NTS_SYM
    g_nccl.handle = h;
    return NTS_OK;
}

#define NTS_NCCL(call)                                                                                        \
    do {                                                                                                      \
        ncclResult_t _r = (call);                                                                             \
        if (_r != ncclSuccess) return ::nts::fail(NTS_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

// one input uint32 (32 bits) -> C output uint32 words of (32/C) fields... here: FIELD bits per input bit
template <int FIELD>
__global__ void expand_bits_kernel(const uint32_t* __restrict__ bits, uint64_t n_words, uint32_t* __restrict__ out)
{
    constexpr int PER = 32 / FIELD;                 // input bits per output word
    const uint64_t n_out = n_words * FIELD;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = bits[o / FIELD] >> ((o % FIELD) * PER);
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) r |= ((x >> i) & 1u) << (i * FIELD);
        out[o] = r;
    }
}

// counters -> bits.  op 0: AND (count == world), op 1: OR (count > 0)
template <int FIELD>
__global__ void threshold_bits_kernel(const uint32_t* __restrict__ cnt, uint64_t n_words, uint32_t world, int op,
                                      uint32_t* __restrict__ bits)
{
    constexpr int PER = 32 / FIELD;
    constexpr uint32_t MASK = (FIELD == 32) ? 0xFFFFFFFFu : ((1u << FIELD) - 1);
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < FIELD; ++j) {
            const uint32_t c = cnt[wi * FIELD + j];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const uint32_t f = (c >> (i * FIELD)) & MASK;
                const uint32_t b = op == 0 ? (f == world) : (f != 0);
                r |= b << (j * PER + i);
            }
        }
        bits[wi] = r;
    }
}

}  // namespace nts

using namespace nts;

struct nts_comm {
    nts_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    DevBuf<uint32_t> stage[2];
    DevBuf<uint32_t> bar;         // one word for nts_nccl_barrier
    uint64_t chunk_words = 0;     // input words per chunk
};

extern "C" {

int nts_nccl_unique_id(uint8_t id_out[128])
{
    if (!id_out) return fail(NTS_ERR_ARG, "null argument");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NTS_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    memcpy(id_out, &id, 128);
    return NTS_OK;
}

int nts_nccl_init(nts_ctx* ctx, const uint8_t id_bytes[128], int rank, int world, nts_comm** out)
{
    if (!ctx || !id_bytes || !out || world < 1 || rank < 0 || rank >= world) return fail(NTS_ERR_ARG, "bad argument");
    if (world > 255) return fail(NTS_ERR_ARG, "at most 255 ranks");
    int rc = load_nccl();
    if (rc) return rc;
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_comm* c = new (std::nothrow) nts_comm();
    if (!c) return fail(NTS_ERR_NOMEM, "host allocation failed");
    c->ctx = ctx; c->rank = rank; c->world = world;
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return fail(NTS_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    for (int i = 0; i < 2; ++i) {
        NTS_CUDA(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
        NTS_CUDA(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
    }
    *out = c;
    return NTS_OK;
}

void nts_nccl_destroy(nts_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int i = 0; i < 2; ++i) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
    }
    delete c;
}

/* Stream-ordered barrier: a one-word all-reduce on the context's stream.  Work queued after it on any rank starts
 * only when every rank's stream has reached its own call -- the inter-GPU ordering the peer-memory merge kernels need
 * (nts_p2p_*), without a host round trip. */
int nts_nccl_barrier(nts_comm* c)
{
    if (!c) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(c->ctx->device));
    if (!c->bar.p) {
        if (c->bar.alloc(1) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (barrier word)");
        NTS_CUDA(cudaMemsetAsync(c->bar.p, 0, 4, c->ctx->stream));
    }
    ProfScope prof(c->ctx, PROF_NCCL, 0.0);
    NTS_NCCL(g_nccl.AllReduce(c->bar.p, c->bar.p, 1, ncclUint32, ncclSum, c->comm, c->ctx->stream));
    c->ctx->launches++;
    return NTS_OK;
}

int nts_nccl_world(const nts_comm* c) { return c ? c->world : 0; }
int nts_nccl_rank(const nts_comm* c) { return c ? c->rank : -1; }

}  // extern "C"

static int field_bits(int world) { return world <= 3 ? 2 : world <= 15 ? 4 : 8; }

template <int FIELD>
static int allreduce_bits(nts_comm* c, nts_bf* bf, int op)
{
    nts_ctx* ctx = bf->ctx;
    const uint64_t n_words = bf->alloc_bytes / 4;
    // chunk: 32 Mi input words (128 MB of bits) -> FIELD x 128 MB of counters per staging buffer
    const uint64_t chunk = std::min<uint64_t>(n_words, 32ull << 20);
    if (c->chunk_words < chunk || c->stage[0].n < chunk * FIELD) {
        for (int i = 0; i < 2; ++i)
            if (c->stage[i].alloc(chunk * FIELD) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (NCCL staging)");
        c->chunk_words = chunk;
    }
    // the filter was produced on ctx->stream: make both side streams wait for it
    cudaEvent_t ready;
    NTS_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    NTS_CUDA(cudaEventRecord(ready, ctx->stream));
    for (int i = 0; i < 2; ++i) NTS_CUDA(cudaStreamWaitEvent(c->streams[i], ready, 0));
    const int grid = ctx->sm_count * 8;
    uint64_t off = 0;
    int which = 0;
    ProfScope prof(ctx, PROF_NCCL, (double)bf->alloc_bytes);
    while (off < n_words) {
        const uint64_t n = std::min(chunk, n_words - off);
        cudaStream_t st = c->streams[which];
        uint32_t* stage = c->stage[which].p;
        expand_bits_kernel<FIELD><<<grid, 256, 0, st>>>(bf->words.p + off, n, stage);
        ctx->launches++;
        NTS_NCCL(g_nccl.AllReduce(stage, stage, n * FIELD, ncclUint32, ncclSum, c->comm, st));
        threshold_bits_kernel<FIELD><<<grid, 256, 0, st>>>(stage, n, (uint32_t)c->world, op, bf->words.p + off);
        ctx->launches++;
        off += n;
        which ^= 1;
    }
    NTS_CUDA(cudaGetLastError());
    for (int i = 0; i < 2; ++i) {
        NTS_CUDA(cudaEventRecord(c->done[i], c->streams[i]));
        NTS_CUDA(cudaStreamWaitEvent(ctx->stream, c->done[i], 0));
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventDestroy(ready);
    return NTS_OK;
}

extern "C" {

static int bf_allreduce(nts_comm* c, nts_bf* bf, int op)
{
    if (!c || !bf) return fail(NTS_ERR_ARG, "null argument");
    if (bf->ctx != c->ctx) return fail(NTS_ERR_ARG, "filter and communicator live on different contexts");
    NTS_CUDA(cudaSetDevice(c->ctx->device));
    if (c->world == 1) return NTS_OK;
    switch (field_bits(c->world)) {
    case 2: return allreduce_bits<2>(c, bf, op);
    case 4: return allreduce_bits<4>(c, bf, op);
    default: return allreduce_bits<8>(c, bf, op);
    }
}

int nts_bf_allreduce_and(nts_comm* c, nts_bf* bf) { return bf_allreduce(c, bf, 0); }
int nts_bf_allreduce_or(nts_comm* c, nts_bf* bf) { return bf_allreduce(c, bf, 1); }

/* gather variable-length minimizer tables: every rank contributes `mine`; out[r] receives rank r's table
 * (a new nts_mxs on this rank's context).  counts[] (world entries, host) must be known to the caller
 * (exchange them over the host side channel). */
int nts_mxs_allgather(nts_comm* c, const nts_mxs* mine, const uint64_t* counts, nts_mxs** out)
{
    if (!c || !mine || !counts || !out) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = c->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    uint64_t mx = 0;
    for (int r = 0; r < c->world; ++r) mx = std::max(mx, counts[r]);
    if (counts[c->rank] != mine->count) return fail(NTS_ERR_ARG, "counts[rank] does not match the local table");
    if (mx == 0) mx = 1;
    // padded all-gather of three columns (h1 as 2 x u32)
    DevBuf<uint32_t> send, recv;
    const uint64_t per = mx * 4;      // u32 words per rank: 2 (h1) + 1 (pos) + 1 (contig)
    if (send.alloc(per) != cudaSuccess || recv.alloc(per * c->world) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (allgather)");
    NTS_CUDA(cudaMemsetAsync(send.p, 0, per * 4, ctx->stream));
    if (mine->count) {
        NTS_CUDA(cudaMemcpyAsync(send.p, mine->h1.p, mine->count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        NTS_CUDA(cudaMemcpyAsync(send.p + mx * 2, mine->pos.p, mine->count * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        NTS_CUDA(cudaMemcpyAsync(send.p + mx * 3, mine->contig.p, mine->count * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    {
        ProfScope prof(ctx, PROF_NCCL, (double)(per * 4 * c->world));
        NTS_NCCL(g_nccl.AllGather(send.p, recv.p, per, ncclUint32, c->comm, ctx->stream));
    }
    for (int r = 0; r < c->world; ++r) {
        nts_mxs* t = new (std::nothrow) nts_mxs();
        if (!t) return fail(NTS_ERR_NOMEM, "host allocation failed");
        t->ctx = ctx; t->count = counts[r]; t->n_contigs = 0;
        const uint64_t n = std::max<uint64_t>(counts[r], 1);
        if (t->h1.alloc(n) != cudaSuccess || t->pos.alloc(n) != cudaSuccess || t->contig.alloc(n) != cudaSuccess) { delete t; return fail(NTS_ERR_NOMEM, "device allocation failed (allgather out)"); }
        const uint32_t* base = recv.p + (uint64_t)r * per;
        if (counts[r]) {
            NTS_CUDA(cudaMemcpyAsync(t->h1.p, base, counts[r] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            NTS_CUDA(cudaMemcpyAsync(t->pos.p, base + mx * 2, counts[r] * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            NTS_CUDA(cudaMemcpyAsync(t->contig.p, base + mx * 3, counts[r] * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        out[r] = t;
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

}  // extern "C"
