// nts_bf_part.cu -- host side of the partitioned Bloom insert (see nts_bin.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "nts_internal.h"
#include "nts_bin.cuh"

namespace nts {

// scratch kept per context so that repeated inserts do not re-allocate 12 GB
struct PartScratch {
    DevBuf<uint32_t> items;
    DevBuf<uint64_t> bucket_off, chunk_first;
    DevBuf<uint32_t> bucket_cap;
    DevBuf<unsigned int> cursor;
};

static std::map<nts_ctx*, PartScratch*> g_scratch;

void part_scratch_release(nts_ctx* ctx)
{
    auto it = g_scratch.find(ctx);
    if (it != g_scratch.end()) { delete it->second; g_scratch.erase(it); }
}

// returns NTS_OK and sets *done = true when the partitioned path ran
int bf_insert_partitioned(nts_ctx* ctx, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid,
                          bool* done)
{
    *done = false;
    const uint64_t m = bf->bytes * 8;
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
    constexpr int THREADS = 512, ITEMS = 16, TILE = THREADS * ITEMS;
    if (total_valid / TILE > 0x7FFFFFF0ull) return NTS_OK;
    // bucket capacities: expectation + 6 sigma + slack (heavy hitters overflow to direct atomics)
    std::vector<uint64_t> off(P + 1, 0), chunk_first(P + 1, 0);
    std::vector<uint32_t> cap(P);
    const uint32_t chunk_items = 4096;
    for (uint32_t b = 0; b < P; ++b) {
        const uint64_t bits_b = std::min<uint64_t>(1ull << shift, m - ((uint64_t)b << shift));
        const double expect = (double)total_valid * (double)bits_b / (double)m;
        uint64_t c = (uint64_t)(expect * 1.02 + 6.0 * std::sqrt(expect) + 4096.0);
        c = (c + 3) & ~3ull;                               // 16-byte aligned buckets (bf_apply_kernel loads uint4)
        if (c > 0xFFFFFFF0ull) return NTS_OK;
        cap[b] = (uint32_t)c;
        off[b + 1] = off[b] + c;
        chunk_first[b + 1] = chunk_first[b] + (c + chunk_items - 1) / chunk_items;
    }
    if (chunk_first[P] > 0x7FFFFFF0ull) return NTS_OK;
    PartScratch*& sc = g_scratch[ctx];
    if (!sc) sc = new PartScratch();
    if (sc->items.n < off[P] && sc->items.alloc(off[P]) != cudaSuccess) { part_scratch_release(ctx); return NTS_OK; }   // no memory: direct path
    if (sc->bucket_off.n < P + 1) {
        if (sc->bucket_off.alloc(1025) != cudaSuccess || sc->chunk_first.alloc(1025) != cudaSuccess ||
            sc->bucket_cap.alloc(1024) != cudaSuccess || sc->cursor.alloc(1024) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (partition tables)");
    }
    NTS_CUDA(cudaMemcpyAsync(sc->bucket_off.p, off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(sc->chunk_first.p, chunk_first.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(sc->bucket_cap.p, cap.data(), P * 4, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(sc->cursor.p, 0, P * 4, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));          // host vectors go out of scope
    BinParams bp;
    bp.items = sc->items.p; bp.bucket_off = sc->bucket_off.p; bp.bucket_cap = sc->bucket_cap.p; bp.cursor = sc->cursor.p;
    bp.n_buckets = P; bp.region_shift = shift;
    const size_t smem = sizeof(HashTables) + (size_t)TILE * 12 + (size_t)P * 12;
    NTS_CUDA(cudaFuncSetAttribute(bf_bin_kernel<THREADS, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t mprime = 0xFFFFFFFFFFFFFFFFull / m;
    const unsigned blocks = (unsigned)((total_valid + TILE - 1) / TILE);
    {
        ProfScope prof(ctx, PROF_BF_INSERT, (double)total_valid);
        bf_bin_kernel<THREADS, ITEMS><<<blocks, THREADS, smem, ctx->stream>>>(gv, tabs, bf->words.p, m, mprime, total_valid, bp);
        ctx->launches++;
        bf_apply_kernel<<<(unsigned)chunk_first[P], 256, 0, ctx->stream>>>(sc->items.p, sc->bucket_off.p, sc->bucket_cap.p,
                                                                          sc->cursor.p, sc->chunk_first.p, P, shift,
                                                                          chunk_items, bf->words.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    *done = true;
    return NTS_OK;
}

}  // namespace nts
