/* synth_oracle.c -- CPU statement of the synthetic-genome generator (TEST INFRASTRUCTURE: bench.py's CPU arm and
 * the tests use it; the product generates its workloads on the device, ntsynt_b200/csrc/nts_synth.cu).
 *
 * The workload is defined in SURVEY.md 8d: an ancestor whose base at (contig, position) is a pure function of a
 * seed (i.i.d. bases with P(A)=P(T)=0.295, P(C)=P(G)=0.205 overlaid with copies of repeat families, each copy
 * mutated at 10 %), and per genome a segment table (ntsynt_b200/synth_layout.py) of ancestor copies (forward /
 * reverse complement), random insertions and N runs, with per-base substitutions.  This file writes the ASCII
 * of a contig prefix from that definition, so the CPU arm can make its sample without touching the CUDA library;
 * tests/test_gpu_parity.py checks that it equals the device genome base for base. */
#include <stddef.h>
#include <stdint.h>

typedef struct {
    uint32_t dst_contig; int32_t anc_contig; uint64_t dst_start, anc_start, len; int32_t strand; uint32_t pad;
} orc_seg;

static inline uint64_t smix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) { return smix(smix(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b); }
static inline unsigned iid_base(uint64_t h)
{
    const unsigned u = (unsigned)(h & 0xFFFF);
    return u < 19333 ? 0u : u < 32768 ? 1u : u < 46203 ? 2u : 3u;
}

#define SLOT_BITS 13
static unsigned ancestor_base(uint64_t anc_seed, uint32_t n_fam, uint32_t rep_thresh16, uint32_t contig, uint64_t pos)
{
    if (n_fam) {
        const uint64_t slot = pos >> SLOT_BITS;
        const uint64_t h = hash3(anc_seed ^ 0x5157ull, contig, slot);
        if ((uint32_t)(h & 0xFFFF) < rep_thresh16) {
            const uint32_t fam = (uint32_t)((h >> 16) % n_fam);
            const uint32_t flen = fam < n_fam / 2 ? 300u : 6000u;
            const uint32_t off = (uint32_t)((h >> 32) % ((1u << SLOT_BITS) - flen));
            const int64_t t = (int64_t)(pos & ((1u << SLOT_BITS) - 1)) - (int64_t)off;
            if (t >= 0 && t < (int64_t)flen) {
                unsigned b = iid_base(hash3(anc_seed ^ 0xFA17ull, fam, (uint64_t)t));
                const uint64_t hm = hash3(anc_seed ^ 0x3117ull, contig, pos);
                if ((hm & 0x3FF) < 102) b = (b + 1 + (unsigned)((hm >> 10) % 3)) & 3u;
                return b;
            }
        }
    }
    return iid_base(hash3(anc_seed, contig, pos));
}

/* ASCII of bases [0, n) of contig `contig`, whose segments are segs[0 .. n_seg) (sorted by dst_start, tiling it) */
void orc_synth_contig(const orc_seg* segs, size_t n_seg, uint32_t contig, uint64_t n, uint64_t anc_seed, uint64_t genome_seed,
                      double sub_rate, uint32_t n_fam, double repeat_slot_prob, char* out)
{
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    double st = sub_rate * 4294967296.0;
    const uint32_t sub_thresh = (uint32_t)(st > 4294967295.0 ? 4294967295.0 : st);
    const uint32_t rep_thresh16 = (uint32_t)(repeat_slot_prob * 65536.0);
    const uint64_t CH = 1u << 16;
    const int64_t n_chunks = (int64_t)((n + CH - 1) / CH);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
        const uint64_t b0 = (uint64_t)ch * CH, b1 = b0 + CH < n ? b0 + CH : n;
        size_t lo = 0, hi = n_seg;                 /* last segment with dst_start <= b0 */
        while (hi - lo > 1) { const size_t mid = (lo + hi) >> 1; if (segs[mid].dst_start <= b0) lo = mid; else hi = mid; }
        size_t s = lo;
        for (uint64_t b = b0; b < b1; ++b) {
            while (b >= segs[s].dst_start + segs[s].len && s + 1 < n_seg) ++s;
            const orc_seg sg = segs[s];
            const uint64_t t = b - sg.dst_start;
            unsigned base;
            if (sg.anc_contig >= 0) {
                base = sg.strand >= 0 ? ancestor_base(anc_seed, n_fam, rep_thresh16, (uint32_t)sg.anc_contig, sg.anc_start + t)
                                      : 3u - ancestor_base(anc_seed, n_fam, rep_thresh16, (uint32_t)sg.anc_contig, sg.anc_start - t);
                const uint64_t hs = hash3(genome_seed, contig, b);
                if ((uint32_t)hs < sub_thresh) base = (base + 1 + (unsigned)((hs >> 32) % 3)) & 3u;
            } else if (sg.anc_contig == -1) {
                base = iid_base(hash3(genome_seed ^ 0x1A5E27ull, contig, b));
            } else {
                out[b] = 'N';
                continue;
            }
            out[b] = acgt[base];
        }
    }
}
