/*
 * oracle/ntsynt_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * A plain-C restatement of the sketch / Bloom-filter half of ntSynt's hot path.
 * Nothing under ntsynt_b200/ may link, import or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * What it restates (reference = /root/reference, bcgsc/ntSynt v1.0.4):
 *   - src/ntsynt_make_common_bf.cpp:28-40   approximate_bf_size
 *   - src/ntsynt_make_common_bf.cpp:122-131 level-1 BF: bf->insert(record.seq)
 *   - src/ntsynt_make_common_bf.cpp:136-160 cascade: contains(prev) -> insert(next)
 *   - bin/ntsynt_run_pipeline.smk:74-85     indexlr -k -w --long --seq --pos -s bf
 *   - subprojects/ntJoin/bin/ntjoin_utils.py:195-202 run_indexlr (refinement rounds)
 *
 * The arithmetic itself lives in btllib (github.com/bcgsc/btllib, pinned by the
 * reference only as "v1.6.2+", README.md:111), which is NOT in /root/reference and
 * not installable here.  The published ntHash2 / KmerBloomFilter / Indexlr
 * algorithms are restated below; parity is PINNED by the reference's own golden
 * indexlr outputs tests/expected_result/ *.k24.w1000.tsv and *.k20.w1000.tsv
 * (295 028 hash:pos:seq triples; see tests/test_oracle_golden.py), which fix the hash
 * function, canonical form (fwd+rev), h1 extension, BF size formula, bit mapping,
 * AND semantics and the rightmost tie-break.  NOT pinned by any golden: whether a
 * minimizer window spans an N run (we follow btllib's loop: windows count valid
 * k-mers only, `restart_on_gap` = 0) -- "parity unpinned" for that one switch.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX UINT64_C(0xFFFFFFFFFFFFFFFF)

/* ntHash2 seeds (btllib nthash_consts) */
static const uint64_t SEED_A = UINT64_C(0x3c8bfbb395c60474);
static const uint64_t SEED_C = UINT64_C(0x3193c18562a02b4c);
static const uint64_t SEED_G = UINT64_C(0x20323ed082572324);
static const uint64_t SEED_T = UINT64_C(0x295549f54be24456);
static const uint64_t MULTISEED = UINT64_C(0x90b45d39fb6da1fa);
#define MULTISHIFT 27

/* code: A0 C1 G2 T3, 4 = anything else; complement of code c is 3-c */
static inline int base_code(unsigned char ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

static inline uint64_t seed_of(int code)
{
    switch (code) {
    case 0: return SEED_A;
    case 1: return SEED_C;
    case 2: return SEED_G;
    default: return SEED_T;
    }
}

/* split rotate left by one: bits 0..32 rotate within 33 bits, bits 33..63 within 31 */
uint64_t orc_srol(uint64_t x)
{
    uint64_t m = ((x & UINT64_C(0x8000000000000000)) >> 30) | ((x & UINT64_C(0x100000000)) >> 32);
    return ((x << 1) & UINT64_C(0xFFFFFFFDFFFFFFFF)) | m;
}

/* inverse of orc_srol */
uint64_t orc_sror(uint64_t x)
{
    uint64_t m = ((x & UINT64_C(0x200000000)) << 30) | ((x & UINT64_C(1)) << 32);
    return ((x >> 1) & UINT64_C(0xFFFFFFFEFFFFFFFF)) | m;
}

static inline uint64_t srol_n(uint64_t x, unsigned n)
{
    for (unsigned i = 0; i < n; ++i) x = orc_srol(x);
    return x;
}

/* direct (non-rolling) definition: canonical hash of one k-mer; returns 0 if the
 * k-mer contains a non-ACGT character, else 1 and *h0 = fwd + rev (mod 2^64). */
int orc_kmer_hash(const char* kmer, unsigned k, uint64_t* h0)
{
    uint64_t fwd = 0, rev = 0;
    for (unsigned i = 0; i < k; ++i) {
        int c = base_code((unsigned char)kmer[i]);
        if (c > 3) return 0;
        fwd = orc_srol(fwd) ^ seed_of(c);
    }
    for (unsigned i = k; i-- > 0;) {
        int c = base_code((unsigned char)kmer[i]);
        rev = orc_srol(rev) ^ seed_of(3 - c);
    }
    *h0 = fwd + rev;
    return 1;
}

/* i-th extension hash (i >= 1): t = h0 * (i ^ k*MULTISEED); t ^= t >> 27 */
uint64_t orc_ext_hash(uint64_t h0, unsigned i, unsigned k)
{
    uint64_t t = h0 * ((uint64_t)i ^ ((uint64_t)k * MULTISEED));
    t ^= t >> MULTISHIFT;
    return t;
}

/* ------------------------------------------------------------------ rolling iterator */
typedef struct {
    const char* seq;
    size_t n;
    unsigned k;
    size_t pos;       /* position of the current k-mer (valid after a successful roll) */
    int inited;
    uint64_t fwd, rev;
    uint64_t out_f[4]; /* srol^k(seed[c])      : contribution of the leaving base, fwd */
    uint64_t in_r[4];  /* srol^k(seed[3-c])    : contribution of the entering base, rev (before sror) */
} orc_roller;

static void roller_init(orc_roller* r, const char* seq, size_t n, unsigned k)
{
    r->seq = seq; r->n = n; r->k = k; r->pos = 0; r->inited = 0; r->fwd = r->rev = 0;
    for (int c = 0; c < 4; ++c) {
        r->out_f[c] = srol_n(seed_of(c), k);
        r->in_r[c] = srol_n(seed_of(3 - c), k);
    }
}

/* (re)initialise at the first all-ACGT k-mer at or after `from`; 0 if none */
static int roller_seek(orc_roller* r, size_t from)
{
    const unsigned k = r->k;
    if (r->n < k) return 0;
    size_t p = from;
    while (p + k <= r->n) {
        /* find a bad base in [p, p+k), scanning from the right so we can skip past it */
        size_t bad = (size_t)-1;
        for (size_t i = p + k; i-- > p;) {
            if (base_code((unsigned char)r->seq[i]) > 3) { bad = i; break; }
        }
        if (bad == (size_t)-1) {
            uint64_t fwd = 0, rev = 0;
            for (size_t i = p; i < p + k; ++i) fwd = orc_srol(fwd) ^ seed_of(base_code((unsigned char)r->seq[i]));
            for (size_t i = p + k; i-- > p;) rev = orc_srol(rev) ^ seed_of(3 - base_code((unsigned char)r->seq[i]));
            r->fwd = fwd; r->rev = rev; r->pos = p; r->inited = 1;
            return 1;
        }
        p = bad + 1;
    }
    return 0;
}

/* advance to the next valid k-mer; 1 on success (btllib NtHash::roll semantics) */
static int roller_roll(orc_roller* r)
{
    if (!r->inited) return roller_seek(r, 0);
    size_t np = r->pos + 1;
    if (np + r->k > r->n) return 0;
    int cin = base_code((unsigned char)r->seq[np + r->k - 1]);
    if (cin > 3) return roller_seek(r, np + r->k);
    int cout = base_code((unsigned char)r->seq[r->pos]);
    r->fwd = orc_srol(r->fwd) ^ seed_of(cin) ^ r->out_f[cout];
    r->rev = orc_sror(r->rev ^ seed_of(3 - cout) ^ r->in_r[cin]);
    r->pos = np;
    return 1;
}

/* h0 of every k-mer of seq: h0_out[p], valid[p] for p in [0, n-k]; returns #valid */
size_t orc_hash_seq(const char* seq, size_t n, unsigned k, uint64_t* h0_out, uint8_t* valid)
{
    size_t nk = n >= k ? n - k + 1 : 0, cnt = 0;
    memset(valid, 0, nk);
    orc_roller r;
    roller_init(&r, seq, n, k);
    while (roller_roll(&r)) {
        h0_out[r.pos] = r.fwd + r.rev;
        valid[r.pos] = 1;
        ++cnt;
    }
    return cnt;
}

/* ------------------------------------------------------------------ Bloom filter */
/* src/ntsynt_make_common_bf.cpp:28-40 + btllib's round-up of the byte count to a
 * multiple of 8 (pinned by the goldens: +8 bytes breaks parity). Returns bytes. */
uint64_t orc_bf_bytes(long long genome_size, double fpr)
{
    long long size_bits = (long long)ceil(((double)(-1 * genome_size)) / log(1 - fpr));
    uint64_t bytes = (uint64_t)(size_bits / 8);
    bytes = (uint64_t)ceil((double)bytes / 8.0) * 8;
    return bytes;
}

static inline void bf_set(uint8_t* bits, uint64_t m, uint64_t h0, int atomic)
{
    uint64_t idx = h0 % m;
    uint8_t mask = (uint8_t)(1u << (idx & 7));
    if (atomic)
        __atomic_fetch_or(&bits[idx >> 3], mask, __ATOMIC_RELAXED);
    else
        bits[idx >> 3] |= mask;
}

static inline int bf_get(const uint8_t* bits, uint64_t m, uint64_t h0)
{
    uint64_t idx = h0 % m;
    return (bits[idx >> 3] >> (idx & 7)) & 1;
}

/* insert every valid k-mer of seq (KmerBloomFilter::insert(seq), 1 hash fn) */
void orc_bf_insert_seq(uint8_t* bits, uint64_t m_bits, const char* seq, size_t n, unsigned k, int atomic)
{
    orc_roller r;
    roller_init(&r, seq, n, k);
    while (roller_roll(&r)) bf_set(bits, m_bits, r.fwd + r.rev, atomic);
}

/* cascade level: insert into next only k-mers present in prev (cpp:145-153) */
void orc_bf_cascade_seq(const uint8_t* prev, uint8_t* next, uint64_t m_bits, const char* seq, size_t n,
                        unsigned k, int atomic)
{
    orc_roller r;
    roller_init(&r, seq, n, k);
    while (roller_roll(&r)) {
        uint64_t h = r.fwd + r.rev;
        if (bf_get(prev, m_bits, h)) bf_set(next, m_bits, h, atomic);
    }
}

void orc_bf_and(uint8_t* dst, const uint8_t* src, uint64_t bytes)
{
    for (uint64_t i = 0; i < bytes; ++i) dst[i] &= src[i];
}

uint64_t orc_bf_popcount(const uint8_t* bits, uint64_t bytes)
{
    uint64_t c = 0;
    for (uint64_t i = 0; i < bytes; ++i) c += (uint64_t)__builtin_popcount(bits[i]);
    return c;
}

/* repeat BF (bin/ntsynt_make_repeat_bfs.py:56-69): k-mers seen >= 2x within this genome */
void orc_bf_repeat_seq(uint8_t* genome_bits, uint8_t* rep_bits, uint64_t m_bits, const char* seq, size_t n, unsigned k)
{
    orc_roller r;
    roller_init(&r, seq, n, k);
    while (roller_roll(&r)) {
        uint64_t h = r.fwd + r.rev;
        if (bf_get(genome_bits, m_bits, h)) bf_set(rep_bits, m_bits, h, 0);
        else bf_set(genome_bits, m_bits, h, 0);
    }
}

/* ------------------------------------------------------------------ minimizers (indexlr) */
typedef struct { uint64_t min_hash, out_hash; size_t pos; } hashed_kmer;

/*
 * One FASTA record through btllib Indexlr::minimize as recalled:
 * ring buffer of w+1 hashed k-mers; a full rescan with `<=` when the current minimum
 * slid out of the window, else a single `<=` comparison with the entering k-mer;
 * emit when the position advances and the minimum is not the UINT64_MAX sentinel.
 * `common` may be NULL (no -s filter); `repeat` may be NULL (no -r filter).
 * restart_on_gap: 0 = windows count valid k-mers and span N runs (btllib as recalled);
 *                 1 = test-only alternative that restarts the window after every gap.
 * Returns the number of minimizers (may exceed cap; only cap are written).
 */
size_t orc_minimize2(const char* seq, size_t n, unsigned k, unsigned w, const uint8_t* common, uint64_t m_bits,
                     const uint8_t* repeat, uint64_t rep_bits, int restart_on_gap, uint64_t* out_h1, uint64_t* out_pos, size_t cap);

size_t orc_minimize(const char* seq, size_t n, unsigned k, unsigned w, const uint8_t* common, const uint8_t* repeat,
                    uint64_t m_bits, int restart_on_gap, uint64_t* out_h1, uint64_t* out_pos, size_t cap)
{
    return orc_minimize2(seq, n, k, w, common, m_bits, repeat, m_bits, restart_on_gap, out_h1, out_pos, cap);
}

/* the same with the repeat filter's own size (indexlr -s common.bf -r repeat.bf: two independent filters) */
size_t orc_minimize2(const char* seq, size_t n, unsigned k, unsigned w, const uint8_t* common, uint64_t m_bits,
                     const uint8_t* repeat, uint64_t rep_bits, int restart_on_gap, uint64_t* out_h1, uint64_t* out_pos, size_t cap)
{
    if (k > n || w > n - k + 1) return 0;
    size_t nout = 0;
    const size_t bufn = (size_t)w + 1;
    hashed_kmer* buf = (hashed_kmer*)malloc(bufn * sizeof(hashed_kmer));
    const hashed_kmer* cur = NULL;
    long long pos_prev = -1;
    size_t idx = 0;
    size_t last_pos = 0;
    int have_last = 0;
    orc_roller r;
    roller_init(&r, seq, n, k);
    while (roller_roll(&r)) {
        if (restart_on_gap && have_last && r.pos != last_pos + 1) { idx = 0; cur = NULL; }
        last_pos = r.pos; have_last = 1;
        uint64_t h0 = r.fwd + r.rev;
        hashed_kmer* hk = &buf[idx % bufn];
        hk->min_hash = h0;
        hk->out_hash = orc_ext_hash(h0, 1, k);
        hk->pos = r.pos;
        if (common && !bf_get(common, m_bits, h0)) hk->min_hash = ORC_MAX;
        if (repeat && bf_get(repeat, rep_bits, h0)) hk->min_hash = ORC_MAX;
        if (idx + 1 >= w) {
            size_t left = idx + 1 - w, right = idx + 1;
            const hashed_kmer* ml = &buf[left % bufn];
            const hashed_kmer* mr = &buf[(right - 1) % bufn];
            if (cur == NULL || cur->pos < ml->pos) {
                cur = ml;
                for (size_t i = left; i < right; ++i) {
                    const hashed_kmer* mi = &buf[i % bufn];
                    if (mi->min_hash <= cur->min_hash) cur = mi;
                }
            } else if (mr->min_hash <= cur->min_hash) {
                cur = mr;
            }
            if ((long long)cur->pos > pos_prev && cur->min_hash != ORC_MAX) {
                pos_prev = (long long)cur->pos;
                if (nout < cap) { out_h1[nout] = cur->out_hash; out_pos[nout] = cur->pos; }
                ++nout;
            }
        }
        ++idx;
    }
    free(buf);
    return nout;
}

/* ------------------------------------------------------------------ reference-structure drivers (CPU baseline)
 * Same threading structure as the reference: OpenMP over FASTA records only, atomic
 * byte-OR inserts (cpp:128-131,145-153); indexlr: one worker per record. */
void orc_common_bf_level1(uint8_t* bits, uint64_t m_bits, const char* const* seqs, const size_t* lens, int nrec,
                          unsigned k, int threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < nrec; ++i) orc_bf_insert_seq(bits, m_bits, seqs[i], lens[i], k, 1);
}

void orc_common_bf_cascade(const uint8_t* prev, uint8_t* next, uint64_t m_bits, const char* const* seqs,
                           const size_t* lens, int nrec, unsigned k, int threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < nrec; ++i) orc_bf_cascade_seq(prev, next, m_bits, seqs[i], lens[i], k, 1);
}

/* sketch every record of one genome; out arrays are per record (caller allocates cap each);
 * counts[i] receives the record's minimizer count */
void orc_sketch_records(const char* const* seqs, const size_t* lens, int nrec, unsigned k, unsigned w,
                        const uint8_t* common, uint64_t m_bits, uint64_t* const* out_h1, uint64_t* const* out_pos,
                        const size_t* caps, size_t* counts, int threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < nrec; ++i)
        counts[i] = orc_minimize(seqs[i], lens[i], k, w, common, NULL, m_bits, 0, out_h1[i], out_pos[i], caps[i]);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
