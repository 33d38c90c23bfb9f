// nts_bf_part.cu -- host side of the partitioned Bloom insert (see nts_part.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "nts_internal.h"
#include "nts_part.cuh"

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

void part_scratch_release(nts_ctx* ctx)
{
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
        NTS_CUDA(cudaFuncSetAttribute(bf_apply_kernel<AP_MAXT, AP_MAXI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.apply_smem));
        const unsigned grid = (unsigned)std::min<uint64_t>(pp.n_regions, (uint64_t)ctx->sm_count * plan.apply_ctas_per_sm);
        bf_apply_kernel<AP_MAXT, AP_MAXI><<<grid, plan.apply_threads, plan.apply_smem, st>>>(
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
