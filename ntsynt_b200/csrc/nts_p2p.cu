// nts_p2p.cu -- Bloom-filter merge over NVLink peer memory (the B200-native alternative to the NCCL
// counter merge of nts_nccl.cu).  One process per GPU: every rank exports its filter with
// CUDA IPC, maps the peers' filters, and then
//   reduce-scatter: rank r ANDs (or ORs) slice r of every peer's array into its own slice r with 128-bit
//                   loads straight from peer HBM -- compute and transfer are one kernel;
//   all-gather    : rank r copies the reduced slice s from peer s for every s != r.
// Wire volume per rank: 2 * (P-1)/P * filter bytes (26 GB at P = 8 for a 14.8 GB filter), 4x less than
// the 4-bit counter all-reduce.  The two phases are separated by host barriers supplied by the caller
// (the launcher's side channel); results are bit-identical to nts_bf_allreduce_and (tests/test_gpu_multi.py).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "nts_internal.h"

namespace nts {
// csrc/nts_bf_part.cu: hash-range owned builds
int owned_prepare(nts_ctx* ctx, int slot, uint64_t m, uint64_t plan_valid);
int owned_bin(nts_ctx* ctx, int slot, nts_bf* sized_like, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid);
int owned_buffers(nts_ctx* ctx, int slot, void** items, void** cursor);
int owned_overflow(nts_ctx* ctx, int slot, uint64_t* n);
int owned_apply(nts_ctx* ctx, int slot, const uint32_t* items, const unsigned int* cursor, nts_bf* bf, uint64_t off16, uint64_t n16);
int bf_range_op(nts_ctx* ctx, nts_bf* dst, const nts_bf* src, uint64_t off16, uint64_t n16, int op);

constexpr int P2P_MAX_RANKS = 16;
struct PeerPtrs { const uint4* p[P2P_MAX_RANKS]; };

// op 0 = AND, 1 = OR
__global__ void p2p_reduce_slice_kernel(uint4* __restrict__ mine, PeerPtrs peers, int rank, int world, uint64_t off16,
                                        uint64_t n16, int op)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 acc = mine[off16 + i];
        uint4 v[P2P_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < P2P_MAX_RANKS; ++p)
            if (p < world && p != rank) v[p] = peers.p[p][off16 + i];          // peer HBM over NVLink
#pragma unroll
        for (int p = 0; p < P2P_MAX_RANKS; ++p)
            if (p < world && p != rank) {
                if (op == 0) { acc.x &= v[p].x; acc.y &= v[p].y; acc.z &= v[p].z; acc.w &= v[p].w; }
                else         { acc.x |= v[p].x; acc.y |= v[p].y; acc.z |= v[p].z; acc.w |= v[p].w; }
            }
        mine[off16 + i] = acc;
    }
}

__global__ void p2p_gather_slice_kernel(uint4* __restrict__ mine, const uint4* __restrict__ peer, uint64_t off16, uint64_t n16)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = peer[off16 + i], b = peer[off16 + i + stride], c = peer[off16 + i + 2 * stride], d = peer[off16 + i + 3 * stride];
        mine[off16 + i] = a; mine[off16 + i + stride] = b; mine[off16 + i + 2 * stride] = c; mine[off16 + i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) mine[off16 + i] = peer[off16 + i];
}

// Contig-sharded ownership: every rank holds, per genome g, the bits of ITS contigs of g.
// common = AND over genomes of (OR over ranks of those partial filters): rank r produces slice r of `out` reading
// slice r of every (genome, rank) partial array straight from peer HBM -- the cascade of
// src/ntsynt_make_common_bf.cpp:136-160 and the cross-GPU merge in one kernel.
constexpr int P2P_MAX_SETS = 8;
struct MultiPtrs { const uint4* p[P2P_MAX_SETS][P2P_MAX_RANKS]; };

__global__ void p2p_and_of_or_kernel(uint4* __restrict__ out, MultiPtrs mp, int n_sets, int world, uint64_t off16, uint64_t n16)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 acc = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (int g = 0; g < n_sets; ++g) {
            uint4 v[P2P_MAX_RANKS];
#pragma unroll
            for (int p = 0; p < P2P_MAX_RANKS; ++p)
                if (p < world) v[p] = mp.p[g][p][off16 + i];                    // all of a genome's loads in flight together
            uint4 t = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int p = 0; p < P2P_MAX_RANKS; ++p)
                if (p < world) { t.x |= v[p].x; t.y |= v[p].y; t.z |= v[p].z; t.w |= v[p].w; }
            acc.x &= t.x; acc.y &= t.y; acc.z &= t.z; acc.w &= t.w;
        }
        out[off16 + i] = acc;
    }
}

}  // namespace nts

using namespace nts;

// csrc/nts_api.cu (defined inside its extern "C" block; not part of the public header)
extern "C" int genome_plain_view(const nts_genome* g, uint32_t k, nts::GenomeView* gv, uint64_t* total_valid, const nts::HashTables** tabs);

struct nts_p2p {
    nts_ctx* ctx = nullptr;
    nts_bf* mine = nullptr;
    int rank = 0, world = 1;
    void* peer[P2P_MAX_RANKS] = {nullptr};
    uint64_t n16 = 0;
};

extern "C" {

int nts_bf_ipc_handle(nts_bf* bf, uint8_t handle_out[64])
{
    if (!bf || !handle_out) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    NTS_CUDA(cudaIpcGetMemHandle(&h, bf->words.p));
    memcpy(handle_out, &h, 64);
    return NTS_OK;
}

int nts_p2p_open(nts_bf* mine, const uint8_t* handles /* world x 64 bytes, rank order */, int rank, int world, nts_p2p** out)
{
    if (!mine || !handles || !out || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world) return fail(NTS_ERR_ARG, "bad argument");
    nts_ctx* ctx = mine->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_p2p* p = new (std::nothrow) nts_p2p();
    if (!p) return fail(NTS_ERR_NOMEM, "host allocation failed");
    p->ctx = ctx; p->mine = mine; p->rank = rank; p->world = world; p->n16 = mine->alloc_bytes / 16;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { p->peer[r] = mine->words.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&p->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != rank && p->peer[q]) cudaIpcCloseMemHandle(p->peer[q]);
            delete p;
            return fail(NTS_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer filter): ") + cudaGetErrorString(e));
        }
    }
    *out = p;
    return NTS_OK;
}

void nts_p2p_close(nts_p2p* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    for (int r = 0; r < p->world; ++r) if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
    delete p;
}

static void slice_of(const nts_p2p* p, int s, uint64_t* off16, uint64_t* n16)
{
    const uint64_t per = (p->n16 + p->world - 1) / p->world;
    *off16 = std::min<uint64_t>(p->n16, per * s);
    *n16 = std::min<uint64_t>(p->n16, per * (s + 1)) - *off16;
}

/* phase 1: my slice of every peer's filter is reduced into my filter.  Call after a barrier that
 * guarantees every rank's filter is complete; synchronous. */
int nts_p2p_reduce_scatter(nts_p2p* p, int op)
{
    if (!p || (op != 0 && op != 1)) return fail(NTS_ERR_ARG, "bad argument");
    nts_ctx* ctx = p->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    uint64_t off, n;
    slice_of(p, p->rank, &off, &n);
    if (n) {
        PeerPtrs pp;
        for (int r = 0; r < P2P_MAX_RANKS; ++r) pp.p[r] = r < p->world ? reinterpret_cast<const uint4*>(p->peer[r]) : nullptr;
        ProfScope prof(ctx, PROF_NCCL, (double)(n * 16) * (p->world - 1));
        p2p_reduce_slice_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(p->mine->words.p), pp, p->rank,
                                                                           p->world, off, n, op);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* Contig-sharded merge, phase 1: slice `rank` of out->mine = AND over the n_sets genomes of (OR over ranks of slice
 * `rank` of sets[g]'s filters).  sets[g] maps genome g's partial filter of every rank, `out` maps the common filter
 * (its phase 2 is nts_p2p_all_gather(out)).  Call after a barrier that guarantees every partial filter is complete. */
int nts_p2p_reduce_and_of_or(nts_p2p* const* sets, uint32_t n_sets, nts_p2p* out)
{
    if (!sets || !out || n_sets < 1 || n_sets > (uint32_t)P2P_MAX_SETS) return fail(NTS_ERR_ARG, "between 1 and 8 genome filter sets are supported");
    nts_ctx* ctx = out->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    MultiPtrs mp;
    memset(&mp, 0, sizeof(mp));
    for (uint32_t g = 0; g < n_sets; ++g) {
        if (!sets[g] || sets[g]->ctx != ctx || sets[g]->world != out->world || sets[g]->rank != out->rank || sets[g]->n16 != out->n16)
            return fail(NTS_ERR_ARG, "filter sets disagree with the output set");
        for (int r = 0; r < out->world; ++r) mp.p[g][r] = reinterpret_cast<const uint4*>(sets[g]->peer[r]);
    }
    uint64_t off, n;
    slice_of(out, out->rank, &off, &n);
    if (n) {
        ProfScope prof(ctx, PROF_NCCL, (double)(n * 16) * (out->world - 1) * n_sets);
        p2p_and_of_or_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(out->mine->words.p), mp, (int)n_sets,
                                                                        out->world, off, n);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* phase 2: fetch every other rank's reduced slice.  Call after a barrier that guarantees every rank
 * finished phase 1; synchronous.  A final barrier must follow before any filter is modified again. */
int nts_p2p_all_gather(nts_p2p* p)
{
    if (!p) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = p->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    {
        ProfScope prof(ctx, PROF_NCCL, (double)p->mine->alloc_bytes);
        // rank r fetches from r+1, r+2, ...: at every step each source serves exactly one reader.  (Fetching in
        // plain rank order makes all P-1 readers pull from the same peer at once -- an incast that divides that
        // peer's NVLink egress by P-1: 105 ms instead of ~40 ms for the whole merge at P = 8.)
        for (int i = 1; i < p->world; ++i) {
            const int s = (p->rank + i) % p->world;
            uint64_t off, n;
            slice_of(p, s, &off, &n);
            if (!n) continue;
            p2p_gather_slice_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(p->mine->words.p),
                                                                               reinterpret_cast<const uint4*>(p->peer[s]), off, n);
            ctx->launches++;
        }
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}


/* ------------------------------------------------------------------------------------------------------------------
 * Hash-range OWNED build of the common filter (multi-GPU): instead of every GPU building whole (partial) filters and the
 * filters being merged -- G x (P-1)/P x 14.8 GB over NVLink per GPU in a contig-sharded run --, every GPU only BINS its
 * k-mers (pass 1 of the partitioned insert, same plan everywhere), and the GPU that owns a byte range of the filter
 * applies the buckets of EVERY GPU to that range, reading them over NVLink peer memory: 4 bytes per k-mer on the wire.
 * The owned slices are the slices of nts_p2p_all_gather, which then distributes the finished common filter. */
struct nts_binpeer {
    nts_ctx* ctx = nullptr;
    int rank = 0, world = 1, slot = 0;
    void* items[P2P_MAX_RANKS] = {nullptr};
    void* cursor[P2P_MAX_RANKS] = {nullptr};
};

int nts_bin_prepare(nts_bf* sized_like, int slot, uint64_t plan_valid)
{
    if (!sized_like || slot < 0 || slot > 7) return fail(NTS_ERR_ARG, "bad argument");
    NTS_CUDA(cudaSetDevice(sized_like->ctx->device));
    return owned_prepare(sized_like->ctx, slot, sized_like->bytes * 8, plan_valid);
}

int nts_bin_genome(nts_bf* sized_like, const nts_genome* g, uint32_t k, int slot)
{
    if (!sized_like || !g || g->ctx != sized_like->ctx) return fail(NTS_ERR_ARG, "bad argument");
    nts_ctx* ctx = sized_like->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    GenomeView gv;
    uint64_t total_valid = 0;
    const HashTables* tabs = nullptr;
    int rc = genome_plain_view(g, k, &gv, &total_valid, &tabs);
    if (rc) return rc;
    return owned_bin(ctx, slot, sized_like, gv, tabs, total_valid);
}

int nts_bin_overflow(nts_ctx* ctx, int slot, uint64_t* n)
{
    if (!ctx || !n) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    return owned_overflow(ctx, slot, n);
}

int nts_bin_ipc_handles(nts_ctx* ctx, int slot, uint8_t items_handle[64], uint8_t cursor_handle[64])
{
    if (!ctx || !items_handle || !cursor_handle) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    void *it = nullptr, *cu = nullptr;
    int rc = owned_buffers(ctx, slot, &it, &cu);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    NTS_CUDA(cudaIpcGetMemHandle(&h, it));
    memcpy(items_handle, &h, 64);
    NTS_CUDA(cudaIpcGetMemHandle(&h, cu));
    memcpy(cursor_handle, &h, 64);
    return NTS_OK;
}

int nts_binpeer_open(nts_ctx* ctx, const uint8_t* items_handles, const uint8_t* cursor_handles, int rank, int world, int slot,
                     nts_binpeer** out)
{
    if (!ctx || !items_handles || !cursor_handles || !out || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world)
        return fail(NTS_ERR_ARG, "bad argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_binpeer* p = new (std::nothrow) nts_binpeer();
    if (!p) return fail(NTS_ERR_NOMEM, "host allocation failed");
    p->ctx = ctx; p->rank = rank; p->world = world; p->slot = slot;
    int rc = owned_buffers(ctx, slot, &p->items[rank], &p->cursor[rank]);
    if (rc) { delete p; return rc; }
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, items_handles + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&p->items[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) {
            memcpy(&h, cursor_handles + (size_t)r * 64, 64);
            e = cudaIpcOpenMemHandle(&p->cursor[r], h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e != cudaSuccess) {
            nts_binpeer_close(p);
            return fail(NTS_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer buckets): ") + cudaGetErrorString(e));
        }
    }
    *out = p;
    return NTS_OK;
}

void nts_binpeer_close(nts_binpeer* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        if (p->items[r]) cudaIpcCloseMemHandle(p->items[r]);
        if (p->cursor[r]) cudaIpcCloseMemHandle(p->cursor[r]);
    }
    delete p;
}

/* bf[off16 .. off16 + n16) |= bits of the k-mers rank `source` binned (asynchronous on the context's stream; the caller
 * orders it after the source's binning pass with a stream-ordered barrier) */
int nts_bf_apply_owned(nts_bf* bf, nts_binpeer* p, int source, uint64_t off16, uint64_t n16)
{
    if (!bf || !p || source < 0 || source >= p->world || bf->ctx != p->ctx) return fail(NTS_ERR_ARG, "bad argument");
    if ((off16 + n16) * 16 > bf->alloc_bytes) return fail(NTS_ERR_ARG, "range outside the filter");
    NTS_CUDA(cudaSetDevice(p->ctx->device));
    return owned_apply(p->ctx, p->slot, static_cast<const uint32_t*>(p->items[source]), static_cast<const unsigned int*>(p->cursor[source]),
                       bf, off16, n16);
}

/* dst[off16 .. off16 + n16) op= src[same range], in units of 16 bytes: op 0 AND, 1 OR, 2 COPY; src null with op 2 = zero
 * (asynchronous on the context's stream) */
int nts_bf_range_op(nts_bf* dst, const nts_bf* src, uint64_t off16, uint64_t n16, int op)
{
    if (!dst || op < 0 || op > 2 || (!src && op != 2)) return fail(NTS_ERR_ARG, "bad argument");
    if (src && (src->ctx != dst->ctx || src->bytes != dst->bytes)) return fail(NTS_ERR_ARG, "filters differ");
    if ((off16 + n16) * 16 > dst->alloc_bytes) return fail(NTS_ERR_ARG, "range outside the filter");
    NTS_CUDA(cudaSetDevice(dst->ctx->device));
    return bf_range_op(dst->ctx, dst, src, off16, n16, op);
}

/* the slice of the filter rank `s` owns in nts_p2p_reduce_scatter / nts_p2p_all_gather (units of 16 bytes) */
int nts_p2p_slice(const nts_p2p* p, int s, uint64_t* off16, uint64_t* n16)
{
    if (!p || !off16 || !n16 || s < 0 || s >= p->world) return fail(NTS_ERR_ARG, "bad argument");
    slice_of(p, s, off16, n16);
    return NTS_OK;
}

}  // extern "C"
