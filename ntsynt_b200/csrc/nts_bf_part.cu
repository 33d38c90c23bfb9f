// nts_bf_part.cu -- host side of the partitioned Bloom insert (see nts_bin.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "nts_internal.h"
#include "nts_bin.cuh"

namespace nts {

// scratch kept per (context, slot) so that repeated inserts do not re-allocate 12 GB; two slots let the
// pipelined build (nts_bf_build_common) bin one genome while the previous one is being applied
struct PartScratch {
    DevBuf<uint32_t> items;
    DevBuf<uint64_t> bucket_off, chunk_first;
    DevBuf<uint32_t> bucket_cap;
    DevBuf<unsigned int> cursor;
    // plan of the current use
    uint32_t P = 0, shift = 0;
    uint64_t n_chunks = 0, n_items = 0;
};

static std::map<std::pair<nts_ctx*, int>, PartScratch*> g_scratch;
static constexpr int BIN_THREADS = 512, BIN_ITEMS = 16, BIN_TILE = BIN_THREADS * BIN_ITEMS;
static constexpr uint32_t CHUNK_ITEMS = 4096;

void part_scratch_release(nts_ctx* ctx)
{
    for (auto it = g_scratch.begin(); it != g_scratch.end();) {
        if (it->first.first == ctx) { delete it->second; it = g_scratch.erase(it); } else ++it;
    }
}

static void scratch_drop(nts_ctx* ctx, int slot)
{
    auto it = g_scratch.find({ctx, slot});
    if (it != g_scratch.end()) { delete it->second; g_scratch.erase(it); }
}

// Plan the buckets of one insert (m filter bits, total_valid k-mers) and upload the tables of `slot`.
// *ok = false when the partitioned path does not apply (small filter, too little work, no memory).
// Synchronises ctx->stream (the host tables go out of scope).
int part_prepare(nts_ctx* ctx, int slot, uint64_t m, uint64_t total_valid, bool* ok)
{
    *ok = false;
    const char* env = getenv("NTS_BF_PARTITION");
    const bool force = env && env[0] == '1';
    if (env && env[0] == '0') return NTS_OK;
    // only worth it when the filter is far larger than L2 and there is enough work
    if (!force && (m < (1ull << 32) || total_valid < (1ull << 26))) return NTS_OK;
    uint32_t shift = 28;                                   // 32 MB regions
    while (((m + (1ull << shift) - 1) >> shift) > 1024 && shift < 32) ++shift;
    const uint64_t P64 = (m + (1ull << shift) - 1) >> shift;
    if (P64 > 1024) return NTS_OK;                         // filter too large for the 16-bit bucket field / smem histogram
    const uint32_t P = (uint32_t)P64;
    if (total_valid / BIN_TILE > 0x7FFFFFF0ull) return NTS_OK;
    // bucket capacities: expectation + 6 sigma + slack (heavy hitters overflow to direct atomics)
    std::vector<uint64_t> off(P + 1, 0), chunk_first(P + 1, 0);
    std::vector<uint32_t> cap(P);
    for (uint32_t b = 0; b < P; ++b) {
        const uint64_t bits_b = std::min<uint64_t>(1ull << shift, m - ((uint64_t)b << shift));
        const double expect = (double)total_valid * (double)bits_b / (double)m;
        uint64_t c = (uint64_t)(expect * 1.02 + 6.0 * std::sqrt(expect) + 4096.0);
        c = (c + 3) & ~3ull;                               // 16-byte aligned buckets (bf_apply_kernel loads uint4)
        if (c > 0xFFFFFFF0ull) return NTS_OK;
        cap[b] = (uint32_t)c;
        off[b + 1] = off[b] + c;
        chunk_first[b + 1] = chunk_first[b] + (c + CHUNK_ITEMS - 1) / CHUNK_ITEMS;
    }
    if (chunk_first[P] > 0x7FFFFFF0ull || off[P] > 0xFFFFFFF0ull) return NTS_OK;   // 32-bit item indices in bf_bin_kernel
    PartScratch*& sc = g_scratch[{ctx, slot}];
    if (!sc) sc = new PartScratch();
    if (sc->items.n < off[P] && sc->items.alloc(off[P]) != cudaSuccess) { scratch_drop(ctx, slot); return NTS_OK; }   // no memory: direct path
    if (sc->bucket_off.n < 1025) {
        if (sc->bucket_off.alloc(1025) != cudaSuccess || sc->chunk_first.alloc(1025) != cudaSuccess ||
            sc->bucket_cap.alloc(1024) != cudaSuccess || sc->cursor.alloc(1024) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (partition tables)");
    }
    NTS_CUDA(cudaMemcpyAsync(sc->bucket_off.p, off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(sc->chunk_first.p, chunk_first.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(sc->bucket_cap.p, cap.data(), P * 4, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));          // host vectors go out of scope
    sc->P = P; sc->shift = shift; sc->n_chunks = chunk_first[P];
    *ok = true;
    return NTS_OK;
}

// pass 1 of a prepared slot on `st`: hash the genome, bin the bit indices by filter region
int part_bin(nts_ctx* ctx, int slot, cudaStream_t st, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid)
{
    PartScratch* sc = g_scratch[{ctx, slot}];
    const uint64_t m = bf->bytes * 8;
    sc->n_items = total_valid;
    NTS_CUDA(cudaMemsetAsync(sc->cursor.p, 0, sc->P * 4, st));
    BinParams bp;
    bp.items = sc->items.p; bp.bucket_off = sc->bucket_off.p; bp.bucket_cap = sc->bucket_cap.p; bp.cursor = sc->cursor.p;
    bp.n_buckets = sc->P; bp.region_shift = sc->shift;
    const size_t smem = sizeof(HashTables) + (size_t)BIN_TILE * 12 + (size_t)sc->P * 12 + 4;
    NTS_CUDA(cudaFuncSetAttribute(bf_bin_kernel<BIN_THREADS, BIN_ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t mprime = 0xFFFFFFFFFFFFFFFFull / m;
    const unsigned blocks = (unsigned)((total_valid + BIN_TILE - 1) / BIN_TILE);
    ProfScope prof(ctx, PROF_BF_BIN, (double)total_valid, false, st);
    bf_bin_kernel<BIN_THREADS, BIN_ITEMS><<<blocks, BIN_THREADS, smem, st>>>(gv, tabs, bf->words.p, m, mprime, total_valid, bp);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// pass 2 of a binned slot on `st`: apply the buckets region by region
int part_apply(nts_ctx* ctx, int slot, cudaStream_t st, nts_bf* bf)
{
    PartScratch* sc = g_scratch[{ctx, slot}];
    ProfScope prof(ctx, PROF_BF_APPLY, (double)sc->n_items, false, st);
    bf_apply_kernel<<<(unsigned)sc->n_chunks, 256, 0, st>>>(sc->items.p, sc->bucket_off.p, sc->bucket_cap.p, sc->cursor.p,
                                                           sc->chunk_first.p, sc->P, sc->shift, CHUNK_ITEMS, bf->words.p);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// returns NTS_OK and sets *done = true when the partitioned path ran
int bf_insert_partitioned(nts_ctx* ctx, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid,
                          bool* done)
{
    *done = false;
    bool ok = false;
    int rc = part_prepare(ctx, 0, bf->bytes * 8, total_valid, &ok);
    if (rc || !ok) return rc;
    {
        ProfScope prof(ctx, PROF_BF_INSERT, (double)total_valid);
        rc = part_bin(ctx, 0, ctx->stream, bf, gv, tabs, total_valid);
        if (rc) return rc;
        rc = part_apply(ctx, 0, ctx->stream, bf);
        if (rc) return rc;
    }
    *done = true;
    return NTS_OK;
}

}  // namespace nts
