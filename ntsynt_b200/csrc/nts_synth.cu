// nts_synth.cu -- synthetic genomes materialised directly in HBM (bench.py workloads).
//
// A genome is described by a host-built segment table (ntsynt_b200/synth.py): every output base is
// either a copy of an ancestor base (forward or reverse-complement), a random inserted base, or N.
// The ancestor is never stored: its base at (contig, position) is a pure function of a seed --
// i.i.d. bases with P(A)=P(T)=0.295, P(C)=P(G)=0.205 overlaid with copies of a few repeat families
// (each copy mutated at 10 %), so any shard can be regenerated on any GPU.  Counter-based hashing
// (splitmix64 finaliser) stands in for Philox; nothing here is on the timed path.
#include <new>

#include "nts_internal.h"

namespace nts {

__host__ __device__ __forceinline__ uint64_t smix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b)
{
    return smix(smix(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b);
}

__host__ __device__ __forceinline__ unsigned iid_base(uint64_t h)
{
    unsigned u = (unsigned)(h & 0xFFFF);
    // A .295 | C .205 | G .205 | T .295
    return u < 19333 ? 0u : u < 32768 ? 1u : u < 46203 ? 2u : 3u;
}

struct SynthParams {
    uint64_t anc_seed, genome_seed;
    uint32_t sub_thresh;       // substitution probability * 2^32
    uint32_t n_fam;            // repeat families (first half 300 bp, second half 6000 bp)
    uint32_t rep_thresh16;     // P(slot hosts a repeat copy) * 65536
};

constexpr uint32_t SLOT_BITS = 13;   // 8192-base slots

__host__ __device__ __forceinline__ unsigned ancestor_base(const SynthParams& p, uint32_t contig, uint64_t pos)
{
    if (p.n_fam) {
        const uint64_t slot = pos >> SLOT_BITS;
        const uint64_t h = hash3(p.anc_seed ^ 0x5157ull, contig, slot);
        if ((uint32_t)(h & 0xFFFF) < p.rep_thresh16) {
            const uint32_t fam = (uint32_t)((h >> 16) % p.n_fam);
            const uint32_t flen = fam < p.n_fam / 2 ? 300u : 6000u;
            const uint32_t off = (uint32_t)((h >> 32) % ((1u << SLOT_BITS) - flen));
            const int64_t t = (int64_t)(pos & ((1u << SLOT_BITS) - 1)) - (int64_t)off;
            if (t >= 0 && t < (int64_t)flen) {
                unsigned b = iid_base(hash3(p.anc_seed ^ 0xFA17ull, fam, (uint64_t)t));
                const uint64_t hm = hash3(p.anc_seed ^ 0x3117ull, contig, pos);
                if ((hm & 0x3FF) < 102) b = (b + 1 + (unsigned)((hm >> 10) % 3)) & 3u;   // 10 % divergence per copy
                return b;
            }
        }
    }
    return iid_base(hash3(p.anc_seed, contig, pos));
}

// one thread = one packed 64-bit word (32 bases) of one contig
__global__ void synth_kernel(uint64_t* __restrict__ packed, const nts_synth_seg* __restrict__ segs,
                             const uint64_t* __restrict__ seg_first /*[n_contigs+1]*/, uint32_t contig,
                             uint64_t contig_len, uint64_t word_off, SynthParams p)
{
    const uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b0 = wi * 32;
    if (b0 >= contig_len) return;
    const uint64_t lo0 = seg_first[contig], hi0 = seg_first[contig + 1];
    // segment covering b0: last s in [lo0, hi0) with dst_start <= b0
    uint64_t lo = lo0, hi = hi0;
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (segs[mid].dst_start <= b0) lo = mid; else hi = mid;
    }
    uint64_t s = lo;
    nts_synth_seg sg = segs[s];
    uint64_t word = 0;
    const uint64_t b1 = min(contig_len, b0 + 32);
    for (uint64_t b = b0; b < b1; ++b) {
        while (b >= sg.dst_start + sg.len && s + 1 < hi0) sg = segs[++s];
        unsigned base = 0;
        const uint64_t t = b - sg.dst_start;
        if (sg.anc_contig >= 0) {
            if (sg.strand >= 0) base = ancestor_base(p, (uint32_t)sg.anc_contig, sg.anc_start + t);
            else                base = 3u - ancestor_base(p, (uint32_t)sg.anc_contig, sg.anc_start - t);
            const uint64_t hs = hash3(p.genome_seed, contig, b);
            if ((uint32_t)hs < p.sub_thresh) base = (base + 1 + (unsigned)((hs >> 32) % 3)) & 3u;
        } else if (sg.anc_contig == -1) {
            base = iid_base(hash3(p.genome_seed ^ 0x1A5E27ull, contig, b));
        }   // -2: N run, packed as 0
        word |= (uint64_t)base << ((b - b0) * 2);
    }
    packed[word_off + wi] = word;
}

}  // namespace nts

using namespace nts;

extern "C" {

int nts_genome_synthesize(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const nts_synth_seg* segs,
                          uint64_t n_seg, uint64_t anc_seed, uint64_t genome_seed, double sub_rate,
                          uint32_t n_repeat_fam, double repeat_slot_prob, nts_genome** out)
{
    if (!ctx || !contig_len || !segs || !out || n_seg == 0) return fail(NTS_ERR_ARG, "null/empty argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_genome* g = new (std::nothrow) nts_genome();
    if (!g) return fail(NTS_ERR_NOMEM, "host allocation failed");
    struct Guard { nts_genome* p; ~Guard() { delete p; } } guard{g};
    g->ctx = ctx;
    g->n_contigs = n_contigs;
    g->contig_len.assign(contig_len, contig_len + n_contigs);
    g->contig_word_off.resize(n_contigs);
    uint64_t woff = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
        if (contig_len[c] > 0xFFFFFFF0ull) return fail(NTS_ERR_ARG, "contig longer than 2^32 bases");
        g->contig_word_off[c] = woff;
        woff += nts_packed_words(contig_len[c]);
        g->total_bases += contig_len[c];
    }
    g->n_words = woff;
    // segment index per contig + N runs (segments with anc_contig == -2), validated on the way
    std::vector<uint64_t> seg_first(n_contigs + 1, 0);
    g->nrun_off.assign(n_contigs + 1, 0);
    {
        uint64_t s = 0;
        for (uint32_t c = 0; c < n_contigs; ++c) {
            seg_first[c] = s;
            g->nrun_off[c] = g->nrun_start.size();
            uint64_t expect = 0;
            while (s < n_seg && segs[s].dst_contig == c) {
                if (segs[s].dst_start != expect || segs[s].len == 0)
                    return fail(NTS_ERR_ARG, "segment table must tile every contig without gaps");
                if (segs[s].anc_contig == -2) {
                    if (!g->nrun_start.empty() && g->nrun_off[c] < g->nrun_start.size() &&
                        g->nrun_start.back() + g->nrun_len.back() == segs[s].dst_start)
                        g->nrun_len.back() += segs[s].len;      // merge adjacent N segments
                    else { g->nrun_start.push_back(segs[s].dst_start); g->nrun_len.push_back(segs[s].len); }
                }
                expect += segs[s].len;
                ++s;
            }
            if (expect != contig_len[c]) return fail(NTS_ERR_ARG, "segment table does not cover contig");
        }
        seg_first[n_contigs] = s;
        g->nrun_off[n_contigs] = g->nrun_start.size();
        if (s != n_seg) return fail(NTS_ERR_ARG, "segment table is not sorted by (contig, dst_start)");
    }
    if (g->packed.alloc(g->n_words + 2) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (genome)");
    DevBuf<nts_synth_seg> d_segs;
    DevBuf<uint64_t> d_first;
    if (d_segs.alloc(n_seg) != cudaSuccess || d_first.alloc(n_contigs + 1) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (segments)");
    NTS_CUDA(cudaMemcpyAsync(d_segs.p, segs, n_seg * sizeof(nts_synth_seg), cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(d_first.p, seg_first.data(), (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(g->packed.p, 0, (g->n_words + 2) * 8, ctx->stream));
    SynthParams p;
    p.anc_seed = anc_seed; p.genome_seed = genome_seed;
    p.sub_thresh = (uint32_t)(sub_rate * 4294967296.0 > 4294967295.0 ? 4294967295.0 : sub_rate * 4294967296.0);
    p.n_fam = n_repeat_fam;
    p.rep_thresh16 = (uint32_t)(repeat_slot_prob * 65536.0);
    ProfScope prof(ctx, PROF_SYNTH, (double)g->total_bases);
    for (uint32_t c = 0; c < n_contigs; ++c) {
        const uint64_t nw = (contig_len[c] + 31) / 32;
        if (!nw) continue;
        synth_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, ctx->stream>>>(g->packed.p, d_segs.p, d_first.p, c, contig_len[c],
                                                                          g->contig_word_off[c], p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = g; guard.p = nullptr;
    return NTS_OK;
}

/* host evaluation of the same generator (tests: device genome == host formula) */
int nts_synth_ancestor_base(uint64_t anc_seed, uint32_t n_repeat_fam, double repeat_slot_prob, uint32_t contig,
                            uint64_t pos)
{
    SynthParams p;
    p.anc_seed = anc_seed; p.genome_seed = 0; p.sub_thresh = 0; p.n_fam = n_repeat_fam;
    p.rep_thresh16 = (uint32_t)(repeat_slot_prob * 65536.0);
    return (int)ancestor_base(p, contig, pos);
}

}  // extern "C"
