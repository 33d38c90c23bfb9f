// nts_bf_part.cu -- host side of the partitioned Bloom insert (see nts_part.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "nts_internal.h"
#include "nts_part.cuh"
#include "nts_bin.cuh"
#include "nts_rank.cuh"

namespace nts {

static constexpr int PT_THREADS = 512, PT_ITEMS = 16, PT_PMAX = 1024, PT_TILE = PT_THREADS * PT_ITEMS;
static constexpr int AP_MAXT = 512, AP_MAXI = 8;

// scratch kept per context so that repeated inserts do not re-allocate ~25 GB
struct PartScratch {
    DevBuf<uint32_t> items1, items2;
    DevBuf<unsigned int> cursor1, cursor2;
    DevBuf<uint64_t> ovf;
    DevBuf<unsigned long long> ovf_count;
    unsigned long long* h_flag = nullptr;      // pinned: [0] overflow count, [1] error flag of the last insert
    bool flag_pending = false;
};

static std::map<nts_ctx*, PartScratch*> g_scratch;
static void pair_scratch_release(nts_ctx* ctx);

void part_scratch_release(nts_ctx* ctx)
{
    pair_scratch_release(ctx);
    auto it = g_scratch.find(ctx);
    if (it == g_scratch.end()) return;
    if (it->second->h_flag) cudaFreeHost(it->second->h_flag);
    delete it->second;
    g_scratch.erase(it);
}

static double env_double(const char* name, double dflt)
{
    const char* e = getenv(name);
    return (e && *e) ? atof(e) : dflt;
}

// ================================================================================================================
// (A) the production path: one shared-memory-atomic binning pass + one L2-resident RED.OR apply pass (nts_bin.cuh)
// ================================================================================================================
// scratch kept per (context, slot) so that repeated inserts do not re-allocate 12 GB; two slots let the
// pipelined build (nts_bf_build_common) bin one genome while the previous one is being applied
struct PairScratch {
    DevBuf<uint32_t> items;
    DevBuf<uint64_t> bucket_off, chunk_first;
    DevBuf<uint32_t> bucket_cap;
    DevBuf<unsigned int> cursor;
    DevBuf<unsigned int> ovf;                  // owned builds: items that did not fit their bucket (counted, not applied)
    std::vector<uint64_t> h_chunk_first;       // host copy of chunk_first
    uint64_t plan_m = 0, plan_valid = 0;       // what the tables were planned for
    // plan of the current use
    uint32_t P = 0, shift = 0;
    uint64_t n_chunks = 0, n_items = 0;
};

static std::map<std::pair<nts_ctx*, int>, PairScratch*> g_pair_scratch;
static constexpr int BIN_THREADS = 512, BIN_ITEMS = 16, BIN_TILE = BIN_THREADS * BIN_ITEMS;
static constexpr uint32_t CHUNK_ITEMS = 4096;

static void pair_scratch_release(nts_ctx* ctx)
{
    for (auto it = g_pair_scratch.begin(); it != g_pair_scratch.end();) {
        if (it->first.first == ctx) { delete it->second; it = g_pair_scratch.erase(it); } else ++it;
    }
}

static void scratch_drop(nts_ctx* ctx, int slot)
{
    auto it = g_pair_scratch.find({ctx, slot});
    if (it != g_pair_scratch.end()) { delete it->second; g_pair_scratch.erase(it); }
}

// Plan the buckets of one insert (m filter bits, total_valid k-mers) and upload the tables of `slot`.
// *ok = false when the partitioned path does not apply (small filter, too little work, no memory).
// Synchronises ctx->stream (the host tables go out of scope).
static int pair_prepare(nts_ctx* ctx, int slot, uint64_t m, uint64_t total_valid, bool* ok, bool always = false)
{
    *ok = false;
    const char* env = getenv("NTS_BF_PARTITION");
    const bool force = always || (env && env[0] == '1');
    if (!always && env && env[0] == '0') return NTS_OK;
    // only worth it when the filter is far larger than L2 and there is enough work
    if (!force && (m < (1ull << 32) || total_valid < (1ull << 26))) return NTS_OK;
    if (m >= (1ull << 42)) return NTS_OK;
    uint32_t shift = 28;                                   // 32 MB regions
    if (const char* es = getenv("NTS_BF_REGION_SHIFT")) { const int x = atoi(es); if (x >= 12 && x <= 32) shift = (uint32_t)x; }   // test knob
    const double cap_scale = env_double("NTS_BF_CAP_SCALE", 1.0);                           // test knob: < 1 forces the overflow branch
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
        uint64_t c = (uint64_t)((expect * 1.02 + 6.0 * std::sqrt(expect) + 4096.0) * cap_scale) + 4;
        c = (c + 3) & ~3ull;                               // 16-byte aligned buckets (bf_apply_kernel loads uint4)
        if (c > 0xFFFFFFF0ull) return NTS_OK;
        cap[b] = (uint32_t)c;
        off[b + 1] = off[b] + c;
        chunk_first[b + 1] = chunk_first[b] + (c + CHUNK_ITEMS - 1) / CHUNK_ITEMS;
    }
    if (chunk_first[P] > 0x7FFFFFF0ull || off[P] > 0xFFFFFFF0ull) return NTS_OK;   // 32-bit item indices in bf_bin_kernel
    PairScratch*& sc = g_pair_scratch[{ctx, slot}];
    if (!sc) sc = new PairScratch();
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
    sc->h_chunk_first = chunk_first; sc->plan_m = m; sc->plan_valid = total_valid;
    *ok = true;
    return NTS_OK;
}

// pass 1 of a prepared slot on `st`: hash the genome, bin the bit indices by filter region
static int pair_bin(nts_ctx* ctx, int slot, cudaStream_t st, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid,
                    const std::vector<UploadStage>* stages = nullptr, unsigned int* ovf_count = nullptr)
{
    PairScratch* sc = g_pair_scratch[{ctx, slot}];
    const uint64_t m = bf->bytes * 8;
    sc->n_items = total_valid;
    NTS_CUDA(cudaMemsetAsync(sc->cursor.p, 0, sc->P * 4, st));
    BinParams bp;
    bp.items = sc->items.p; bp.bucket_off = sc->bucket_off.p; bp.bucket_cap = sc->bucket_cap.p; bp.cursor = sc->cursor.p;
    bp.n_buckets = sc->P; bp.region_shift = sc->shift; bp.ovf_count = ovf_count;
    const uint64_t mprime = 0xFFFFFFFFFFFFFFFFull / m;
    const unsigned n_tiles = (unsigned)((total_valid + BIN_TILE - 1) / BIN_TILE);
    // NTS_BF_BIN=2: ranking without shared-memory atomics (nts_rank.cuh; same output up to the order inside a bucket)
    // where the item word has room for the high bucket digit.  Measured slower on B200 (34.8 against 21.2 ms per 3 Gbp
    // genome, profiles/r02c_ncu_rank_bin.md): its two counting-sort passes issue 285 instructions and ~65 shared-memory
    // wavefronts per k-mer row, which costs more than the one ATOMS they replace -- so the atomicAdd ranking stays the default.
    const char* eb = getenv("NTS_BF_BIN");
    const bool old_rank = !(eb && eb[0] == '2');
    const size_t smem_r = RankSmem<BIN_THREADS, BIN_ITEMS>::bytes(sc->P);
    const size_t smem = sizeof(HashTables) + (size_t)BIN_TILE * 12 + (size_t)sc->P * 12 + 4;
    const int which = (!old_rank && sc->P <= 512 && sc->shift <= 28) ? 4 : (!old_rank && sc->P <= 1024 && sc->shift <= 27) ? 5 : 0;
    if (which == 4)
        NTS_CUDA(cudaFuncSetAttribute(bf_rank_bin_kernel<BIN_THREADS, BIN_ITEMS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    else if (which == 5)
        NTS_CUDA(cudaFuncSetAttribute(bf_rank_bin_kernel<BIN_THREADS, BIN_ITEMS, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    else
        NTS_CUDA(cudaFuncSetAttribute(bf_bin_kernel<BIN_THREADS, BIN_ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto launch = [&](unsigned block0, unsigned blocks) {
        if (which == 4)
            bf_rank_bin_kernel<BIN_THREADS, BIN_ITEMS, 4><<<blocks, BIN_THREADS, smem_r, st>>>(gv, tabs, bf->words.p, m, mprime, total_valid, bp, block0);
        else if (which == 5)
            bf_rank_bin_kernel<BIN_THREADS, BIN_ITEMS, 5><<<blocks, BIN_THREADS, smem_r, st>>>(gv, tabs, bf->words.p, m, mprime, total_valid, bp, block0);
        else
            bf_bin_kernel<BIN_THREADS, BIN_ITEMS><<<blocks, BIN_THREADS, smem, st>>>(gv, tabs, bf->words.p, m, mprime, total_valid, bp, block0);
        ctx->launches++;
    };
    ProfScope prof(ctx, PROF_BF_PART1, (double)total_valid, false, st);
    if (!stages) {
        launch(0, n_tiles);
    } else {
        // the genome is still being uploaded: after every stage's event, the tiles whose k-mers are all on the device
        unsigned from = 0;
        for (size_t i = 0; i < stages->size(); ++i) {
            const UploadStage& sg = (*stages)[i];
            const unsigned to = i + 1 == stages->size() ? n_tiles : (unsigned)std::min<uint64_t>(n_tiles, sg.v_end / BIN_TILE);
            if (to <= from && i + 1 < stages->size()) continue;
            NTS_CUDA(cudaStreamWaitEvent(st, sg.ev, 0));
            if (to > from) launch(from, to - from);
            from = std::max(from, to);
        }
    }
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// pass 2 of a binned slot on `st`: apply the buckets region by region
static int pair_apply(nts_ctx* ctx, int slot, cudaStream_t st, nts_bf* bf)
{
    PairScratch* sc = g_pair_scratch[{ctx, slot}];
    ProfScope prof(ctx, PROF_BF_APPLY, (double)sc->n_items, false, st);
    bf_apply_kernel<<<(unsigned)sc->n_chunks, 256, 0, st>>>(sc->items.p, sc->bucket_off.p, sc->bucket_cap.p, sc->cursor.p,
                                                           sc->chunk_first.p, sc->P, sc->shift, CHUNK_ITEMS, bf->words.p);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// OR bits(genome) into bf with the pair; *done = false (nothing launched) when the pair does not apply
int pair_insert(nts_ctx* ctx, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid, bool* done,
                const std::vector<UploadStage>* stages)
{
    *done = false;
    bool ok = false;
    int rc = pair_prepare(ctx, 0, bf->bytes * 8, total_valid, &ok);
    if (rc || !ok) return rc;
    {
        ProfScope prof(ctx, PROF_BF_INSERT, (double)total_valid);
        if ((rc = pair_bin(ctx, 0, ctx->stream, bf, gv, tabs, total_valid, stages)) || (rc = pair_apply(ctx, 0, ctx->stream, bf))) return rc;
    }
    ctx->part_inserts++;
    *done = true;
    return NTS_OK;
}

// ================================================================================================================
// (A') hash-range OWNED builds (multi-GPU, csrc/nts_p2p.cu): every GPU bins its k-mers with the same plan; the GPU that
//      owns a range of the filter applies the buckets of every GPU -- read over NVLink peer memory -- to that range
// ================================================================================================================
int owned_prepare(nts_ctx* ctx, int slot, uint64_t m, uint64_t plan_valid)
{
    bool ok = false;
    auto it = g_pair_scratch.find({ctx, slot});
    if (it != g_pair_scratch.end() && it->second->plan_m == m && it->second->plan_valid == plan_valid && it->second->ovf.p) return NTS_OK;
    int rc = pair_prepare(ctx, slot, m, plan_valid, &ok, true);
    if (rc) return rc;
    if (!ok) return fail(NTS_ERR_ARG, "owned build: this filter size / amount of work cannot be planned (too many regions or no memory)");
    PairScratch* sc = g_pair_scratch[{ctx, slot}];
    if (!sc->ovf.p && sc->ovf.alloc(4) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (overflow counter)");
    NTS_CUDA(cudaMemsetAsync(sc->ovf.p, 0, 16, ctx->stream));
    return NTS_OK;
}

int owned_bin(nts_ctx* ctx, int slot, nts_bf* sized_like, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid)
{
    auto it = g_pair_scratch.find({ctx, slot});
    if (it == g_pair_scratch.end() || it->second->plan_m != sized_like->bytes * 8) return fail(NTS_ERR_STATE, "owned build: slot not prepared for this filter");
    PairScratch* sc = it->second;
    if (total_valid > sc->plan_valid) return fail(NTS_ERR_ARG, "owned build: more k-mers than the plan was made for");
    if (!total_valid) { NTS_CUDA(cudaMemsetAsync(sc->cursor.p, 0, sc->P * 4, ctx->stream)); return NTS_OK; }
    return pair_bin(ctx, slot, ctx->stream, sized_like, gv, tabs, total_valid, nullptr, sc->ovf.p);
}

int owned_buffers(nts_ctx* ctx, int slot, void** items, void** cursor)
{
    auto it = g_pair_scratch.find({ctx, slot});
    if (it == g_pair_scratch.end()) return fail(NTS_ERR_STATE, "owned build: slot not prepared");
    *items = it->second->items.p; *cursor = it->second->cursor.p;
    return NTS_OK;
}

int owned_overflow(nts_ctx* ctx, int slot, uint64_t* n)
{
    auto it = g_pair_scratch.find({ctx, slot});
    if (it == g_pair_scratch.end()) return fail(NTS_ERR_STATE, "owned build: slot not prepared");
    unsigned int h = 0;
    NTS_CUDA(cudaMemcpyAsync(&h, it->second->ovf.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *n = h;
    return NTS_OK;
}

// apply the buckets (`items`, `cursor`: this GPU's or a peer's, same plan) to words [off16, off16 + n16) x 16 bytes of bf
int owned_apply(nts_ctx* ctx, int slot, const uint32_t* items, const unsigned int* cursor, nts_bf* bf, uint64_t off16, uint64_t n16)
{
    auto it = g_pair_scratch.find({ctx, slot});
    if (it == g_pair_scratch.end() || it->second->plan_m != bf->bytes * 8) return fail(NTS_ERR_STATE, "owned build: slot not prepared for this filter");
    PairScratch* sc = it->second;
    if (!n16) return NTS_OK;
    const uint64_t bit_lo = off16 * 128, bit_hi = std::min<uint64_t>((off16 + n16) * 128, bf->alloc_bytes * 8);
    const uint32_t b_lo = (uint32_t)(bit_lo >> sc->shift);
    const uint32_t b_hi = (uint32_t)std::min<uint64_t>(sc->P, ((bit_hi - 1) >> sc->shift) + 1);       // one past the last region
    if (b_lo >= b_hi) return NTS_OK;
    const uint64_t c0 = sc->h_chunk_first[b_lo], c1 = sc->h_chunk_first[b_hi];
    if (c1 <= c0) return NTS_OK;
    ProfScope prof(ctx, PROF_BF_APPLY, 0.0);
    bf_apply_owned_kernel<<<(unsigned)(c1 - c0), 256, 0, ctx->stream>>>(items, sc->bucket_off.p, sc->bucket_cap.p, cursor,
                                                                       sc->chunk_first.p, sc->P, sc->shift, CHUNK_ITEMS, bf->words.p,
                                                                       c0, bit_lo, bit_hi);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

int bf_range_op(nts_ctx* ctx, nts_bf* dst, const nts_bf* src, uint64_t off16, uint64_t n16, int op)
{
    if (!n16) return NTS_OK;
    ProfScope prof(ctx, src ? PROF_BF_COMBINE : PROF_FILL, (double)n16 * 16);
    bf_range_kernel<<<(unsigned)std::min<uint64_t>((n16 + 255) / 256, (uint64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        reinterpret_cast<uint4*>(dst->words.p), src ? reinterpret_cast<const uint4*>(src->words.p) : nullptr, off16, n16, op);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// ================================================================================================================
// (B) the atomics-free three-pass variant (nts_part.cuh), selected with NTS_BF_IMPL=3: measured slower on B200
//     (profiles/r02a_ncu_partition3pass.md) but it writes whole filters (SET), fuses the AND and needs no zero-fill
// ================================================================================================================
struct PartPlan {
    PartParams pp;
    uint32_t chunks_per_bucket = 0, apply_threads = 0, apply_ctas_per_sm = 1;
    size_t apply_smem = 0;
};

// Plan the two partition levels for a filter of m bits and total_valid k-mers.  *ok = false when the partitioned
// path does not apply (small filter, too little work, region too large for shared memory, index overflow).
// Test knobs: NTS_BF_PARTITION=0/1 (never / always), NTS_BF_P1MAX, NTS_BF_P2 (bucket counts), NTS_BF_CAP_SCALE
// (scales the bucket capacities: < 1 forces the overflow list), NTS_BF_OVF_CAP (overflow list entries).
static bool plan_partition(nts_ctx* ctx, uint64_t m, uint64_t total_valid, PartPlan* plan)
{
    const char* env = getenv("NTS_BF_PARTITION");
    const bool force = env && env[0] == '1';
    if (env && env[0] == '0') return false;
    if (!force && (m < (1ull << 30) || total_valid < (1ull << 24))) return false;      // the direct kernel is fine in L2
    if (m >= (1ull << 38) || total_valid == 0) return false;
    uint32_t P1max = (uint32_t)env_double("NTS_BF_P1MAX", PT_PMAX);
    uint32_t P2 = (uint32_t)env_double("NTS_BF_P2", PT_PMAX);
    P1max = std::max(1u, std::min<uint32_t>(P1max, PT_PMAX));
    uint32_t log2P2 = 0;
    while ((2u << log2P2) <= std::min<uint32_t>(std::max(P2, 1u), PT_PMAX)) ++log2P2;
    P2 = 1u << log2P2;
    int max_optin = 0;
    if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device) != cudaSuccess) return false;
    uint64_t R = (m + (uint64_t)P1max * P2 - 1) / ((uint64_t)P1max * P2);
    R = std::max<uint64_t>(128, (R + 127) / 128 * 128);
    if (R > (uint64_t)max_optin - 1024) return false;                                   // byte flags of a region must fit one CTA
    PartParams& pp = plan->pp;
    pp.R = (uint32_t)R;
    pp.P2 = P2; pp.log2P2 = log2P2;
    pp.R1 = R * P2;
    const uint64_t n_regions = (m + R - 1) / R;
    pp.n_regions = (uint32_t)n_regions;
    pp.P1 = (uint32_t)((n_regions + P2 - 1) / P2);
    const uint32_t R7 = (uint32_t)(R >> 7);
    uint32_t lg = 0;
    while ((2u << lg) <= R7) ++lg;
    pp.shift = 31 + lg;
    pp.magic = (uint32_t)(((1ull << pp.shift) / R7) + 1);
    const double scale = env_double("NTS_BF_CAP_SCALE", 1.0);
    auto cap_for = [&](double bits, double slack) -> uint64_t {
        const double expect = (double)total_valid * bits / (double)m;
        uint64_t c = (uint64_t)((expect * 1.02 + 6.0 * std::sqrt(expect) + slack) * scale) + 4;
        return (c + 3) & ~3ull;
    };
    const uint64_t cap1 = cap_for((double)pp.R1, 1024.0), cap2 = cap_for((double)R, 64.0);
    if ((uint64_t)pp.P1 * cap1 > 0xFFFFFFF0ull || (uint64_t)P2 * cap2 > 0xFFFFFFF0ull || cap1 > 0x7FFFFFF0ull) return false;
    pp.cap1 = (uint32_t)cap1; pp.cap2 = (uint32_t)cap2;
    plan->chunks_per_bucket = (uint32_t)((cap1 + PT_TILE - 1) / PT_TILE);
    if ((uint64_t)pp.P1 * plan->chunks_per_bucket > 0x7FFFFFF0ull || (total_valid + PT_TILE - 1) / PT_TILE > 0x7FFFFFF0ull) return false;
    pp.ovf_cap = (uint64_t)env_double("NTS_BF_OVF_CAP", (double)(total_valid / 8 + (1u << 20)));
    // apply: one thread per 128 flags, as many passes over the region as needed with balanced threads
    const uint32_t q_per = (uint32_t)(R >> 7);
    const uint32_t iters = (q_per + AP_MAXT - 1) / AP_MAXT;
    plan->apply_threads = std::min<uint32_t>(AP_MAXT, std::max<uint32_t>(64, (((q_per + iters - 1) / iters) + 31) / 32 * 32));
    plan->apply_smem = (size_t)R;
    int smem_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device);
    plan->apply_ctas_per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(4, (size_t)smem_sm / (plan->apply_smem + 1024)));
    return true;
}

static int scratch_for(nts_ctx* ctx, PartPlan* plan, PartScratch** out)
{
    PartScratch*& sc = g_scratch[ctx];
    if (!sc) sc = new PartScratch();
    PartParams& pp = plan->pp;
    const size_t n1 = (size_t)pp.P1 * pp.cap1, n2 = (size_t)pp.P1 * pp.P2 * pp.cap2, nc2 = (size_t)pp.P1 * pp.P2;
    bool ok = true;
    if (sc->items1.n < n1) ok = ok && sc->items1.alloc(n1) == cudaSuccess;
    if (ok && sc->items2.n < n2) ok = sc->items2.alloc(n2) == cudaSuccess;
    if (ok && sc->cursor1.n < pp.P1) ok = sc->cursor1.alloc(PT_PMAX) == cudaSuccess;
    if (ok && sc->cursor2.n < nc2) ok = sc->cursor2.alloc(nc2) == cudaSuccess;
    if (ok && sc->ovf.n < pp.ovf_cap) ok = sc->ovf.alloc(pp.ovf_cap) == cudaSuccess;
    if (ok && sc->ovf_count.n < 2) ok = sc->ovf_count.alloc(2) == cudaSuccess;
    if (ok && !sc->h_flag) ok = cudaHostAlloc(reinterpret_cast<void**>(&sc->h_flag), 16, cudaHostAllocPortable) == cudaSuccess;
    if (!ok) {                     // no memory for the scratch: the caller takes the direct path
        cudaGetLastError();
        part_scratch_release(ctx);
        return NTS_ERR_NOMEM;
    }
    pp.items1 = sc->items1.p; pp.items2 = sc->items2.p; pp.cursor1 = sc->cursor1.p; pp.cursor2 = sc->cursor2.p;
    pp.ovf = sc->ovf.p; pp.ovf_count = sc->ovf_count.p;
    *out = sc;
    return NTS_OK;
}

// One partitioned insert on `st`: out = bits(genome) [SET], prev & bits [AND, prev != out], out | bits [OR, prev == out].
// *done = false (and nothing launched) when the partitioned path does not apply.  The overflow-list error flag of the
// insert is copied to pinned memory; part_check() reads it after the stream has been synchronised.
int part_insert(nts_ctx* ctx, cudaStream_t st, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid, uint64_t m,
                const uint32_t* prev, uint32_t* out, uint64_t alloc_bytes, int mode, bool* done)
{
    *done = false;
    const char* impl = getenv("NTS_BF_IMPL");
    if (!impl || impl[0] != '3') return NTS_OK;
    PartPlan plan;
    if (!plan_partition(ctx, m, total_valid, &plan)) return NTS_OK;
    PartScratch* sc = nullptr;
    if (scratch_for(ctx, &plan, &sc) != NTS_OK) return NTS_OK;
    const PartParams& pp = plan.pp;
    NTS_CUDA(cudaMemsetAsync(pp.cursor1, 0, (size_t)pp.P1 * 4, st));
    NTS_CUDA(cudaMemsetAsync(pp.cursor2, 0, (size_t)pp.P1 * pp.P2 * 4, st));
    NTS_CUDA(cudaMemsetAsync(pp.ovf_count, 0, 16, st));
    const char* env_match = getenv("NTS_BF_MATCH");                  // ranking by match.any instead of ballots (same result)
    const bool use_match = env_match && env_match[0] == '1';
    using L = PartSmem<PT_THREADS, PT_ITEMS, PT_PMAX>;
    const size_t smem1 = sizeof(HashTables) + L::BYTES, smem2 = L::BYTES;
    const uint64_t mprime = 0xFFFFFFFFFFFFFFFFull / m;
    const unsigned blocks1 = (unsigned)((total_valid + PT_TILE - 1) / PT_TILE);
    const unsigned blocks2 = pp.P1 * plan.chunks_per_bucket;
    {
        ProfScope prof(ctx, PROF_BF_PART1, (double)total_valid, false, st);
        if (use_match) {
            NTS_CUDA(cudaFuncSetAttribute(bf_part1_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            bf_part1_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, true><<<blocks1, PT_THREADS, smem1, st>>>(gv, tabs, m, mprime, total_valid, pp);
        } else {
            NTS_CUDA(cudaFuncSetAttribute(bf_part1_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            bf_part1_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, false><<<blocks1, PT_THREADS, smem1, st>>>(gv, tabs, m, mprime, total_valid, pp);
        }
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    {
        ProfScope prof(ctx, PROF_BF_PART2, (double)total_valid, false, st);
        if (use_match) {
            NTS_CUDA(cudaFuncSetAttribute(bf_part2_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            bf_part2_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, true><<<blocks2, PT_THREADS, smem2, st>>>(pp, plan.chunks_per_bucket);
        } else {
            NTS_CUDA(cudaFuncSetAttribute(bf_part2_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            bf_part2_kernel<PT_THREADS, PT_ITEMS, PT_PMAX, false><<<blocks2, PT_THREADS, smem2, st>>>(pp, plan.chunks_per_bucket);
        }
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    {
        ProfScope prof(ctx, PROF_BF_APPLY, (double)total_valid, false, st);
        NTS_CUDA(cudaFuncSetAttribute(bf_apply3_kernel<AP_MAXT, AP_MAXI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.apply_smem));
        const unsigned grid = (unsigned)std::min<uint64_t>(pp.n_regions, (uint64_t)ctx->sm_count * plan.apply_ctas_per_sm);
        bf_apply3_kernel<AP_MAXT, AP_MAXI><<<grid, plan.apply_threads, plan.apply_smem, st>>>(
            pp, reinterpret_cast<const uint4*>(prev), reinterpret_cast<uint4*>(out), alloc_bytes / 16, mode);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
        bf_overflow_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(pp.ovf, pp.ovf_count, pp.ovf_cap, prev, out, mode);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    if (sc->flag_pending) {        // an earlier insert's flag has not been looked at: keep a sticky OR of the error flags
        NTS_CUDA(cudaStreamSynchronize(st));
        if (sc->h_flag[1]) return fail(NTS_ERR_OVERFLOW, "partitioned Bloom insert: overflow list exhausted (set NTS_BF_PARTITION=0)");
    }
    NTS_CUDA(cudaMemcpyAsync(sc->h_flag, pp.ovf_count, 16, cudaMemcpyDeviceToHost, st));
    sc->flag_pending = true;
    ctx->part_inserts++;
    *done = true;
    return NTS_OK;
}

// after the stream of the last part_insert has been synchronised: *overflowed = its overflow list was exhausted
// (the filter is then incomplete and the caller must redo the insert with the direct kernel)
void part_check(nts_ctx* ctx, bool* overflowed, uint64_t* ovf_items)
{
    *overflowed = false;
    if (ovf_items) *ovf_items = 0;
    auto it = g_scratch.find(ctx);
    if (it == g_scratch.end() || !it->second->flag_pending) return;
    it->second->flag_pending = false;
    *overflowed = it->second->h_flag[1] != 0;
    if (ovf_items) *ovf_items = it->second->h_flag[0];
    ctx->part_overflow_items += it->second->h_flag[0];
}

}  // namespace nts
