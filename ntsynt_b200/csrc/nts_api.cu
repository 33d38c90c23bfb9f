// nts_api.cu -- C-ABI of libntsynt_b200.so: context, ingest, Bloom filter and sketch entry points.
// See include/ntsynt_b200.h for the contract and the reference call sites each entry replaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "nts_internal.h"
#include "nts_kernels.cuh"

namespace nts {

void part_scratch_release(nts_ctx* ctx);
int part_insert(nts_ctx* ctx, cudaStream_t st, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid, uint64_t m,
                const uint32_t* prev, uint32_t* out, uint64_t alloc_bytes, int mode, bool* done);
void part_check(nts_ctx* ctx, bool* overflowed, uint64_t* ovf_items);
int pair_insert(nts_ctx* ctx, nts_bf* bf, const GenomeView& gv, const HashTables* tabs, uint64_t total_valid, bool* done,
                const std::vector<UploadStage>* stages = nullptr);
enum { APPLY_SET = 0, APPLY_AND = 1, APPLY_OR = 2 };      // nts_part.cuh

static thread_local std::string g_err;

// ---- caching device allocator (see nts_internal.h)
namespace {
struct PoolState {
    std::mutex mu;
    std::map<int, std::multimap<size_t, void*>> free_blocks;   // device -> size -> block
    std::map<void*, std::pair<int, size_t>> live;              // block -> (device, size)
};
PoolState& pool() { static PoolState* s = new PoolState(); return *s; }
size_t pool_round(size_t b)
{
    if (b < 4096) return 4096;
    if (b < (1u << 20)) { size_t r = 4096; while (r < b) r <<= 1; return r; }
    const size_t g = 2u << 20;                                  // 2 MB granules
    return (b + g - 1) / g * g;
}
}  // namespace

cudaError_t pool_alloc(void** out, size_t bytes)
{
    const size_t want = pool_round(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    PoolState& ps = pool();
    {
        std::lock_guard<std::mutex> lk(ps.mu);
        auto& fl = ps.free_blocks[dev];
        auto it = fl.lower_bound(want);
        if (it != fl.end() && it->first <= want + want / 4 + (1u << 20)) {     // close enough in size: reuse
            *out = it->second;
            ps.live[*out] = {dev, it->first};
            fl.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim(dev);                                          // out of memory: drop the cache and retry once
        e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(ps.mu);
        ps.live[*out] = {dev, want};
    }
    return e;
}

void pool_free(void* p)
{
    if (!p) return;
    PoolState& ps = pool();
    std::lock_guard<std::mutex> lk(ps.mu);
    auto it = ps.live.find(p);
    if (it == ps.live.end()) { cudaFree(p); return; }
    ps.free_blocks[it->second.first].emplace(it->second.second, p);
    ps.live.erase(it);
}

void pool_trim(int device)
{
    PoolState& ps = pool();
    std::vector<void*> blocks;
    {
        std::lock_guard<std::mutex> lk(ps.mu);
        auto& fl = ps.free_blocks[device];
        for (auto& kv : fl) blocks.push_back(kv.second);
        fl.clear();
    }
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    for (void* b : blocks) cudaFree(b);
    if (cur != device) cudaSetDevice(cur);
}

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) { g_err = msg; return code; }

static uint64_t srol_n(uint64_t x, unsigned n) { for (unsigned i = 0; i < n; ++i) x = srol(x); return x; }

static void make_hash_tables(uint32_t k, HashTables* t)
{
    std::memset(t, 0, sizeof(*t));
    for (unsigned in = 0; in < 4; ++in)
        for (unsigned out = 0; out < 4; ++out) {
            const uint64_t rf = seed_of(in) ^ srol_n(seed_of(out), k);
            const uint64_t rr = seed_of(3 - out) ^ srol_n(seed_of(3 - in), k);
            t->roll[in * 4 + out] = make_uint4((uint32_t)rf, (uint32_t)(rf >> 32), (uint32_t)rr, (uint32_t)(rr >> 32));
        }
    for (unsigned i = 0; i < k; ++i)
        for (unsigned c = 0; c < 4; ++c) {
            t->init_f[i * 4 + c] = srol_n(seed_of(c), k - 1 - i);
            t->init_r[i * 4 + c] = srol_n(seed_of(3 - c), i);
        }
}

static int get_tables(nts_ctx* ctx, uint32_t k, const HashTables** out)
{
    if (k < 1 || k > MAX_K) return fail(NTS_ERR_ARG, "k must be in [1, 64]");
    auto it = ctx->tables.find(k);
    if (it == ctx->tables.end()) {
        HashTables host;
        make_hash_tables(k, &host);
        HashTables* dev = nullptr;
        NTS_CUDA(cudaMalloc(reinterpret_cast<void**>(&dev), sizeof(HashTables)));
        NTS_CUDA(cudaMemcpyAsync(dev, &host, sizeof(HashTables), cudaMemcpyHostToDevice, ctx->stream));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        it = ctx->tables.emplace(k, dev).first;
    }
    *out = it->second;
    return NTS_OK;
}

// Build the valid-k-mer islands of a genome for (k, optional mask) and upload them.
static int build_view(const nts_genome* g, uint32_t k, const uint64_t* mask_off, const uint64_t* mask_start,
                      const uint64_t* mask_end, nts_view** out)
{
    nts_view* v = new (std::nothrow) nts_view();
    if (!v) return fail(NTS_ERR_NOMEM, "host allocation failed");
    v->k = k;
    std::vector<uint64_t> seg_v, seg_base;
    seg_v.push_back(0);
    v->contig_v.assign(g->n_contigs + 1, 0);
    uint64_t total = 0;
    for (uint32_t c = 0; c < g->n_contigs; ++c) {
        v->contig_v[c] = total;
        const uint64_t len = g->contig_len[c];
        const uint64_t cbase = g->contig_word_off[c] * 32;
        // merge the contig's N runs and mask intervals (both sorted) into blocked intervals
        size_t ni = g->nrun_off[c], ne = g->nrun_off[c + 1];
        size_t mi = mask_off ? mask_off[c] : 0, me = mask_off ? mask_off[c + 1] : 0;
        uint64_t clean_from = 0;   // start of the current clean stretch
        auto emit_clean = [&](uint64_t a, uint64_t b) {   // clean bases [a, b)
            if (b > a && b - a >= k) {
                uint64_t n = b - a - k + 1;
                seg_base.push_back(cbase + a);
                total += n;
                seg_v.push_back(total);
            }
        };
        while (ni < ne || mi < me) {
            uint64_t s, e;
            bool take_n = (mi >= me) || (ni < ne && g->nrun_start[ni] <= mask_start[mi]);
            if (take_n) { s = g->nrun_start[ni]; e = s + g->nrun_len[ni]; ++ni; }
            else { s = mask_start[mi]; e = mask_end[mi]; ++mi; }
            if (e > len) e = len;
            if (s >= e) continue;
            if (s > clean_from) emit_clean(clean_from, s);
            if (e > clean_from) clean_from = e;
        }
        if (len > clean_from) emit_clean(clean_from, len);
    }
    v->contig_v[g->n_contigs] = total;
    v->total_valid = total;
    v->n_seg = (uint32_t)seg_base.size();
    v->h_seg_v = seg_v; v->h_seg_base = seg_base;
    if (seg_base.size() > 0xFFFFFFF0ull) { delete v; return fail(NTS_ERR_ARG, "too many islands"); }
    cudaError_t e1 = v->seg_v.alloc(seg_v.size());
    cudaError_t e2 = v->seg_base.alloc(seg_base.size() ? seg_base.size() : 1);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { delete v; return fail(NTS_ERR_NOMEM, "device allocation failed (view)"); }
    cudaStream_t st = g->ctx->stream;
    cudaError_t e = cudaMemcpyAsync(v->seg_v.p, seg_v.data(), seg_v.size() * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !seg_base.empty())
        e = cudaMemcpyAsync(v->seg_base.p, seg_base.data(), seg_base.size() * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete v; return fail(NTS_ERR_CUDA, std::string("view upload: ") + cudaGetErrorString(e)); }
    *out = v;
    return NTS_OK;
}

// order ctx->stream after an asynchronous upload of the genome (once; later work on the stream is ordered by it)
static int wait_ready(const nts_genome* g)
{
    nts_genome* gm = const_cast<nts_genome*>(g);
    if (gm->ready && !gm->ready_waited) {
        NTS_CUDA(cudaStreamWaitEvent(gm->ctx->stream, gm->ready, 0));
        gm->ready_waited = true;
    }
    return NTS_OK;
}

static int get_plain_view(const nts_genome* g, uint32_t k, const nts_view** out, bool wait = true)
{
    if (wait) { int rc = wait_ready(g); if (rc) return rc; }
    nts_genome* gm = const_cast<nts_genome*>(g);
    auto it = gm->views.find(k);
    if (it == gm->views.end()) {
        nts_view* v = nullptr;
        int rc = build_view(g, k, nullptr, nullptr, nullptr, &v);
        if (rc) return rc;
        it = gm->views.emplace(k, v).first;
    }
    *out = it->second;
    return NTS_OK;
}

static GenomeView device_view(const nts_genome* g, const nts_view* v)
{
    GenomeView gv;
    gv.packed = g->packed.p;
    gv.seg_v = v->seg_v.p;
    gv.seg_base = v->seg_base.p;
    gv.n_seg = v->n_seg;
    gv.k = v->k;
    return gv;
}

static int grid_for(const nts_ctx* ctx, uint64_t items, int threads, int per_sm)
{
    uint64_t want = (items + threads - 1) / threads;
    uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
    return (int)std::max<uint64_t>(1, std::min(want, cap));
}

}  // namespace nts

using namespace nts;

extern "C" {

const char* nts_version(void) { return "ntsynt_b200 0.1 (sm_100a; parity target: ntSynt v1.0.4)"; }
const char* nts_last_error(void) { return g_err.c_str(); }

int nts_device_count(int* out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return fail(NTS_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *out = n;
    return NTS_OK;
}

// ------------------------------------------------------------------------------------ context
int nts_ctx_create(int device, nts_ctx** out)
{
    if (!out) return fail(NTS_ERR_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(NTS_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(NTS_ERR_ARG, "device index out of range");
    NTS_CUDA(cudaSetDevice(device));
    nts_ctx* ctx = new (std::nothrow) nts_ctx();
    if (!ctx) return fail(NTS_ERR_NOMEM, "host allocation failed");
    ctx->device = device;
    cudaDeviceProp prop;
    NTS_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    NTS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    NTS_CUDA(cudaEventCreate(&ctx->ev0));
    NTS_CUDA(cudaEventCreate(&ctx->ev1));
    NTS_CUDA(cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming));
    *out = ctx;
    return NTS_OK;
}

void nts_ctx_destroy(nts_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    part_scratch_release(ctx);
    pool_trim(ctx->device);
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int nts_ctx_sync(nts_ctx* ctx)
{
    NTS_CUDA(cudaSetDevice(ctx->device));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    bool over = false;
    part_check(ctx, &over, nullptr);      // an asynchronous partitioned insert whose overflow list ran out
    if (over) return fail(NTS_ERR_OVERFLOW, "partitioned Bloom insert: overflow list exhausted; the filter is incomplete "
                                            "(use nts_bf_insert_genome, or set NTS_BF_PARTITION=0)");
    return NTS_OK;
}

int nts_timer_start(nts_ctx* ctx)
{
    NTS_CUDA(cudaSetDevice(ctx->device));
    NTS_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return NTS_OK;
}

int nts_timer_stop(nts_ctx* ctx, float* ms_out)
{
    NTS_CUDA(cudaSetDevice(ctx->device));
    NTS_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    NTS_CUDA(cudaEventSynchronize(ctx->ev1));
    NTS_CUDA(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    return NTS_OK;
}

uint64_t nts_launch_count(const nts_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t nts_sketch_escalated(const nts_ctx* ctx) { return ctx ? ctx->sketch_escalated : 0; }
uint64_t nts_sketch_queried_all(const nts_ctx* ctx) { return ctx ? ctx->sketch_qall : 0; }

int nts_mem_info(nts_ctx* ctx, uint64_t* free_b, uint64_t* total_b)
{
    NTS_CUDA(cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    NTS_CUDA(cudaMemGetInfo(&f, &t));
    *free_b = f; *total_b = t;
    return NTS_OK;
}

int nts_prof_enable(nts_ctx* ctx, int on)
{
    if (!ctx) return fail(NTS_ERR_ARG, "null argument");
    ctx->prof_enabled = on != 0;
    return NTS_OK;
}

int nts_prof_reset(nts_ctx* ctx)
{
    if (!ctx) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& p : ctx->prof_pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    ctx->prof_pending.clear();
    for (int i = 0; i < PROF_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_units[i] = 0; ctx->prof_launches[i] = 0; }
    ctx->h2d_bytes = ctx->d2h_bytes = 0;
    return NTS_OK;
}

static const char* const PROF_NAMES[PROF_COUNT] = {"fill", "bf_insert", "bf_combine", "sketch", "sketch_post", "join",
                                                   "synth", "popcount", "bf_repeat", "edges", "nccl", "bf_build", "bf_part1", "bf_part2",
                                                   "bf_apply", "graph"};

int nts_prof_count(void) { return PROF_COUNT; }
const char* nts_prof_name(int id) { return (id >= 0 && id < PROF_COUNT) ? PROF_NAMES[id] : ""; }

int nts_prof_get(nts_ctx* ctx, int id, double* ms, double* units, uint64_t* launches)
{
    if (!ctx || id < 0 || id >= PROF_COUNT) return fail(NTS_ERR_ARG, "bad argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->prof_pending.empty()) {
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        for (auto& p : ctx->prof_pending) {
            float t = 0;
            if (cudaEventElapsedTime(&t, p.e0, p.e1) == cudaSuccess) { ctx->prof_ms[p.id] += t; ctx->prof_units[p.id] += p.units; }
            cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
        }
        ctx->prof_pending.clear();
    }
    if (ms) *ms = ctx->prof_ms[id];
    if (units) *units = ctx->prof_units[id];
    if (launches) *launches = ctx->prof_launches[id];
    return NTS_OK;
}

int nts_xfer_bytes(const nts_ctx* ctx, uint64_t* h2d, uint64_t* d2h)
{
    if (!ctx) return fail(NTS_ERR_ARG, "null argument");
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    return NTS_OK;
}

int nts_host_alloc(uint64_t bytes, void** out)
{
    if (!out) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return NTS_OK;
}

void nts_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------ ingest (host)
uint64_t nts_packed_words(uint64_t n_bases) { return ((n_bases + 63) / 64) * 2; }

int nts_pack_ascii(const char* seq, uint64_t n, uint64_t* words_out, uint64_t* nrun_start, uint64_t* nrun_len,
                   uint64_t nrun_cap, uint64_t* n_nruns)
{
    if (!seq || !words_out || !n_nruns) return fail(NTS_ERR_ARG, "null argument");
    static const auto lut = [] {
        struct L { uint8_t v[256]; } l;
        for (int i = 0; i < 256; ++i) l.v[i] = 4;
        l.v['A'] = l.v['a'] = 0; l.v['C'] = l.v['c'] = 1; l.v['G'] = l.v['g'] = 2; l.v['T'] = l.v['t'] = 3;
        return l;
    }();
    const uint64_t nw = nts_packed_words(n);
    uint64_t runs = 0;
    bool in_run = false;
    uint64_t run_start = 0;
    for (uint64_t wi = 0; wi < nw; ++wi) {
        uint64_t word = 0;
        const uint64_t b0 = wi * 32, b1 = std::min(n, b0 + 32);
        for (uint64_t b = b0; b < b1; ++b) {
            uint8_t c = lut.v[(unsigned char)seq[b]];
            if (c > 3) {
                if (!in_run) { in_run = true; run_start = b; }
                c = 0;
            } else if (in_run) {
                if (runs < nrun_cap) { nrun_start[runs] = run_start; nrun_len[runs] = b - run_start; }
                ++runs;
                in_run = false;
            }
            word |= (uint64_t)c << ((b - b0) * 2);
        }
        words_out[wi] = word;
    }
    if (in_run) {
        if (runs < nrun_cap) { nrun_start[runs] = run_start; nrun_len[runs] = n - run_start; }
        ++runs;
    }
    *n_nruns = runs;
    return NTS_OK;
}

int nts_unpack_ascii(const uint64_t* words, uint64_t start, uint64_t n, char* out)
{
    if (!words || !out) return fail(NTS_ERR_ARG, "null argument");
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t b = start + i;
        out[i] = acgt[(words[b >> 5] >> ((b & 31) * 2)) & 3];
    }
    return NTS_OK;
}

static int genome_upload_impl(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const uint64_t* contig_word_off,
                              const uint64_t* words, uint64_t n_words, const uint64_t* nrun_off, const uint64_t* nrun_start,
                              const uint64_t* nrun_len, bool async_copy, nts_genome** out)
{
    if (!ctx || !out || !contig_len || !contig_word_off || !nrun_off) return fail(NTS_ERR_ARG, "null argument");
    if (n_words && !words) return fail(NTS_ERR_ARG, "words is null");
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_genome* g = new (std::nothrow) nts_genome();
    if (!g) return fail(NTS_ERR_NOMEM, "host allocation failed");
    g->ctx = ctx;
    g->n_contigs = n_contigs;
    g->contig_len.assign(contig_len, contig_len + n_contigs);
    g->contig_word_off.assign(contig_word_off, contig_word_off + n_contigs);
    g->n_words = n_words;
    for (uint32_t c = 0; c < n_contigs; ++c) {
        g->total_bases += contig_len[c];
        if (contig_len[c] > 0xFFFFFFF0ull) { delete g; return fail(NTS_ERR_ARG, "contig longer than 2^32 bases"); }
        if (contig_word_off[c] + nts_packed_words(contig_len[c]) > n_words || (contig_word_off[c] & 1)) {
            delete g;
            return fail(NTS_ERR_ARG, "contig_word_off/n_words inconsistent with contig_len (16-byte aligned contigs)");
        }
    }
    g->nrun_off.assign(nrun_off, nrun_off + n_contigs + 1);
    const uint64_t nr = nrun_off[n_contigs];
    if (nr) { g->nrun_start.assign(nrun_start, nrun_start + nr); g->nrun_len.assign(nrun_len, nrun_len + nr); }
    // two guard words so that cursor pre-loads one word past the last base stay in bounds
    if (g->packed.alloc(n_words + 2) != cudaSuccess) { delete g; return fail(NTS_ERR_NOMEM, "device allocation failed (genome)"); }
    cudaError_t e = cudaSuccess;
    if (async_copy) {
        // copy on the context's H2D stream; consumers order themselves after `ready` (wait_ready)
        if (!ctx->stream_copy) e = cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking);
        // the block may have been recycled from something the main stream is still using (the pool frees without
        // tracking streams): the copy starts after the work queued on the main stream so far
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_side, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream_copy, ctx->ev_side, 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->packed.p + n_words, 0, 16, ctx->stream_copy);
        if (e == cudaSuccess && n_words) {
            ctx->h2d_bytes += n_words * 8;
            // 1/8, 1/8, 1/4, 1/2 of the words: the first Bloom insert starts after an eighth of the copy
            const uint64_t cuts[4] = {n_words / 8, n_words / 4, n_words / 2, n_words};
            uint64_t from = 0;
            for (int c = 0; c < 4 && e == cudaSuccess; ++c) {
                const uint64_t to = c == 3 ? n_words : (cuts[c] & ~1ull);
                if (to <= from) continue;
                e = cudaMemcpyAsync(g->packed.p + from, words + from, (to - from) * 8, cudaMemcpyHostToDevice, ctx->stream_copy);
                uint64_t stage_min = 1ull << 24;                 // stage events only where they can pay (>= 128 MB of words)
                if (const char* es = getenv("NTS_UPLOAD_STAGE_MIN_WORDS")) stage_min = strtoull(es, nullptr, 10);   // test knob
                if (c < 3 && n_words >= stage_min) {
                    cudaEvent_t ev = nullptr;
                    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
                    if (e == cudaSuccess) e = cudaEventRecord(ev, ctx->stream_copy);
                    if (e == cudaSuccess) { g->chunk_ev.push_back(ev); g->chunk_end_word.push_back(to); }
                }
                from = to;
            }
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(g->ready, ctx->stream_copy);
        g->ready_waited = false;
    } else {
        e = cudaMemsetAsync(g->packed.p + n_words, 0, 16, ctx->stream);
        if (e == cudaSuccess && n_words)
            e = copy_h2d(ctx, g->packed.p, words, n_words * 8);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        if (g->ready) cudaEventDestroy(g->ready);
        for (cudaEvent_t ev : g->chunk_ev) cudaEventDestroy(ev);
        delete g;
        return fail(NTS_ERR_CUDA, std::string("genome upload: ") + cudaGetErrorString(e));
    }
    *out = g;
    return NTS_OK;
}

int nts_genome_upload(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const uint64_t* contig_word_off,
                      const uint64_t* words, uint64_t n_words, const uint64_t* nrun_off, const uint64_t* nrun_start,
                      const uint64_t* nrun_len, nts_genome** out)
{
    return genome_upload_impl(ctx, n_contigs, contig_len, contig_word_off, words, n_words, nrun_off, nrun_start, nrun_len, false, out);
}

int nts_genome_upload_async(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const uint64_t* contig_word_off,
                            const uint64_t* words, uint64_t n_words, const uint64_t* nrun_off, const uint64_t* nrun_start,
                            const uint64_t* nrun_len, nts_genome** out)
{
    return genome_upload_impl(ctx, n_contigs, contig_len, contig_word_off, words, n_words, nrun_off, nrun_start, nrun_len, true, out);
}


void nts_genome_destroy(nts_genome* g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    if (g->ready) { cudaEventSynchronize(g->ready); cudaEventDestroy(g->ready); }   // the copy must not outlive the buffer
    for (cudaEvent_t ev : g->chunk_ev) cudaEventDestroy(ev);
    for (auto& kv : g->views) delete kv.second;
    delete g;
}

uint64_t nts_genome_size(const nts_genome* g) { return g ? g->total_bases : 0; }
uint32_t nts_genome_contigs(const nts_genome* g) { return g ? g->n_contigs : 0; }

int nts_genome_download_contig(nts_genome* g, uint32_t contig, uint64_t* words_out)
{
    if (!g || contig >= g->n_contigs || !words_out) return fail(NTS_ERR_ARG, "bad argument");
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    { int rc = wait_ready(g); if (rc) return rc; }
    NTS_CUDA(copy_d2h(g->ctx, words_out, g->packed.p + g->contig_word_off[contig],
                      nts_packed_words(g->contig_len[contig]) * 8));
    NTS_CUDA(cudaStreamSynchronize(g->ctx->stream));
    return NTS_OK;
}

int nts_genome_nruns(nts_genome* g, uint64_t* nrun_off, uint64_t* nrun_start, uint64_t* nrun_len, uint64_t cap,
                     uint64_t* n_out)
{
    if (!g || !nrun_off || !n_out) return fail(NTS_ERR_ARG, "bad argument");
    std::copy(g->nrun_off.begin(), g->nrun_off.end(), nrun_off);
    *n_out = g->nrun_start.size();
    uint64_t n = std::min<uint64_t>(cap, g->nrun_start.size());
    if (n && nrun_start && nrun_len) {
        std::copy(g->nrun_start.begin(), g->nrun_start.begin() + n, nrun_start);
        std::copy(g->nrun_len.begin(), g->nrun_len.begin() + n, nrun_len);
    }
    return NTS_OK;
}

// ------------------------------------------------------------------------------------ Bloom filter
uint64_t nts_bf_bytes(int64_t genome_size, double fpr)
{
    // src/ntsynt_make_common_bf.cpp:38-39, then btllib's round-up to whole 64-bit words
    long long size_bits = (long long)std::ceil(((double)(-1 * (long long)genome_size)) / std::log(1 - fpr));
    uint64_t bytes = (uint64_t)(size_bits / 8);
    bytes = (uint64_t)std::ceil((double)bytes / 8.0) * 8;
    return bytes;
}

static int bf_fill(nts_bf* bf, uint32_t v, cudaStream_t st = nullptr)
{
    nts_ctx* ctx = bf->ctx;
    uint64_t n16 = bf->alloc_bytes / 16;
    if (!st) st = ctx->stream;
    ProfScope prof(ctx, PROF_FILL, (double)bf->alloc_bytes, false, st);
    fill_u128_kernel<<<grid_for(ctx, n16, 256, 16), 256, 0, st>>>(reinterpret_cast<uint4*>(bf->words.p), n16, v);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

int nts_bf_create(nts_ctx* ctx, uint64_t bytes, nts_bf** out)
{
    if (!ctx || !out) return fail(NTS_ERR_ARG, "null argument");
    if (bytes == 0 || (bytes & 7)) return fail(NTS_ERR_ARG, "Bloom filter size must be a positive multiple of 8 bytes");
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_bf* bf = new (std::nothrow) nts_bf();
    if (!bf) return fail(NTS_ERR_NOMEM, "host allocation failed");
    bf->ctx = ctx;
    bf->bytes = bytes;
    bf->alloc_bytes = (bytes + 15) & ~15ull;
    if (bf->words.alloc(bf->alloc_bytes / 4) != cudaSuccess) {
        delete bf;
        return fail(NTS_ERR_NOMEM, "device allocation failed (Bloom filter of " + std::to_string(bytes) + " bytes)");
    }
    int rc = bf_fill(bf, 0);
    if (rc == NTS_OK) { cudaError_t e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = fail(NTS_ERR_CUDA, cudaGetErrorString(e)); }
    if (rc) { delete bf; return rc; }
    *out = bf;
    return NTS_OK;
}

void nts_bf_destroy(nts_bf* bf)
{
    if (!bf) return;
    cudaSetDevice(bf->ctx->device);
    delete bf;
}

uint64_t nts_bf_size_bytes(const nts_bf* bf) { return bf ? bf->bytes : 0; }

int nts_bf_clear(nts_bf* bf)
{
    if (!bf) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    int rc = bf_fill(bf, 0);
    if (rc) return rc;
    NTS_CUDA(cudaStreamSynchronize(bf->ctx->stream));
    return NTS_OK;
}

static int mod_params(const nts_bf* bf, uint64_t* m, uint64_t* mprime)
{
    *m = bf->bytes * 8;
    *mprime = 0xFFFFFFFFFFFFFFFFull / *m;
    return NTS_OK;
}

static int bf_combine(nts_bf* dst, const nts_bf* src, int op, bool sync);

// what code outside this file needs to hash a genome's k-mers (csrc/nts_p2p.cu: owned builds)
int genome_plain_view(const nts_genome* g, uint32_t k, GenomeView* gv, uint64_t* total_valid, const HashTables** tabs)
{
    int rc = get_tables(g->ctx, k, tabs);
    if (rc) return rc;
    const nts_view* v = nullptr;
    if ((rc = get_plain_view(g, k, &v))) return rc;
    *gv = device_view(g, v);
    *total_valid = v->total_valid;
    return NTS_OK;
}

// direct insert: one RED.OR per k-mer (small filters; the fallback of the partitioned path)
static int bf_insert_direct(nts_ctx* ctx, nts_bf* bf, const nts_genome* g, const nts_view* v, const HashTables* tabs)
{
    uint64_t m, mp;
    mod_params(bf, &m, &mp);
    constexpr int THREADS = 256;
    const uint32_t chunk = 32;
    uint64_t blocks = (v->total_valid + (uint64_t)THREADS * chunk - 1) / ((uint64_t)THREADS * chunk);
    if (blocks > 0x7FFFFFFFull) return fail(NTS_ERR_ARG, "genome too large for one launch");
    ProfScope prof(ctx, PROF_BF_INSERT, (double)v->total_valid);
    bf_insert_kernel<THREADS><<<(unsigned)blocks, THREADS, 0, ctx->stream>>>(device_view(g, v), tabs, bf->words.p, m, mp,
                                                                            v->total_valid, chunk);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    return NTS_OK;
}

// bits(genome) into a filter.  mode APPLY_OR: bf |= bits (the reference's bf->insert(seq), src/ntsynt_make_common_bf.cpp:130);
// APPLY_SET: bf = bits; APPLY_AND: dst = dst & bits with bf as the scratch level filter of the same size (one level of the
// cascade, cpp:136-160; `dst` is updated in place -- its device storage may be swapped with the scratch filter's).
// Large filters take a partitioned path (nts_bin.cuh, or nts_part.cuh with NTS_BF_IMPL=3).
static int bf_insert_mode(nts_bf* bf, nts_bf* dst, const nts_genome* g, uint32_t k, int mode, cudaEvent_t prefilled = nullptr)
{
    nts_ctx* ctx = bf->ctx;
    const HashTables* tabs = nullptr;
    int rc = get_tables(ctx, k, &tabs);
    if (rc) return rc;
    // a genome whose asynchronous upload nobody has waited for yet: the view needs only host data, and the binning pass
    // of the partitioned insert can start on the head of the genome while the rest is still on its way
    std::vector<UploadStage> stages;
    const char* impl3 = getenv("NTS_BF_IMPL");
    const bool staged = g->ready && !g->ready_waited && !g->chunk_ev.empty() && !(impl3 && impl3[0] == '3');
    const nts_view* v = nullptr;
    rc = get_plain_view(g, k, &v, !staged);
    if (rc) return rc;
    if (staged) {
        bool monotone = true;
        for (size_t i = 1; i < v->h_seg_base.size(); ++i) monotone = monotone && v->h_seg_base[i] > v->h_seg_base[i - 1];
        if (monotone) {
            for (size_t c = 0; c < g->chunk_ev.size(); ++c) {
                // bases below `limit` are on the device after this chunk (the kernels pre-load a few words ahead: margin)
                const uint64_t have = g->chunk_end_word[c] * 32;
                const uint64_t limit = have > 256 ? have - 256 : 0;
                uint64_t v_end = 0;
                for (size_t s_ = 0; s_ < v->h_seg_base.size(); ++s_) {
                    const uint64_t sb = v->h_seg_base[s_], n_s = v->h_seg_v[s_ + 1] - v->h_seg_v[s_];
                    if (sb + k > limit) break;
                    const uint64_t fit = std::min<uint64_t>(n_s, limit - k - sb + 1);
                    v_end = v->h_seg_v[s_] + fit;
                    if (fit < n_s) break;
                }
                stages.push_back({g->chunk_ev[c], v_end});
            }
            stages.push_back({g->ready, v->total_valid});
        }
    }
    const uint64_t n16 = bf->alloc_bytes / 16;
    bool done = false;
    if (v->total_valid && stages.empty()) {
        if (staged && (rc = wait_ready(g))) return rc;
        rc = part_insert(ctx, ctx->stream, device_view(g, v), tabs, v->total_valid, bf->bytes * 8,
                         mode == APPLY_AND ? dst->words.p : bf->words.p, bf->words.p, bf->alloc_bytes, mode, &done);
        if (rc) return rc;
    }
    if (done) {
        if (mode == APPLY_AND) std::swap(dst->words, bf->words);      // the three-pass apply wrote dst & bits into the scratch
        return NTS_OK;
    }
    // zero-fill (SET / AND), then OR the genome in -- large filters with the binning + apply pair (nts_bin.cuh), small
    // ones with one RED.OR per k-mer; then the AND pass
    // (`prefilled`: the zero-fill already ran on another stream -- wait for it instead)
    if (prefilled) NTS_CUDA(cudaStreamWaitEvent(ctx->stream, prefilled, 0));
    else if (mode != APPLY_OR && (rc = bf_fill(bf, 0))) return rc;
    if (v->total_valid) {
        if ((rc = pair_insert(ctx, bf, device_view(g, v), tabs, v->total_valid, &done, stages.empty() ? nullptr : &stages))) return rc;
        if (!stages.empty()) const_cast<nts_genome*>(g)->ready_waited = done;     // the last stage waited for `ready`
        if (!done && ((rc = wait_ready(g)) || (rc = bf_insert_direct(ctx, bf, g, v, tabs)))) return rc;
    } else if (staged && (rc = wait_ready(g))) return rc;
    if (mode == APPLY_AND) {
        ProfScope prof(ctx, PROF_BF_COMBINE, (double)bf->alloc_bytes);
        bf_combine_kernel<<<grid_for(ctx, n16 / 4 + 1, 256, 16), 256, 0, ctx->stream>>>(
            reinterpret_cast<uint4*>(dst->words.p), reinterpret_cast<const uint4*>(bf->words.p), n16, 0);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    return NTS_OK;
}

// synchronise and make sure the last partitioned insert did not exhaust its overflow list; if it did, redo it directly
// (only the three-pass variant has such a list).  AND: `keep` is a copy of dst taken before the insert.
static int bf_finish_insert(nts_bf* bf, nts_bf* dst, const nts_genome* g, uint32_t k, int mode)
{
    nts_ctx* ctx = bf->ctx;
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    bool over = false;
    part_check(ctx, &over, nullptr);
    if (!over) return NTS_OK;
    // pathological input (one k-mer making up a large part of the genome): OR is idempotent, SET starts over; AND: the
    // swap left the previous dst in `bf` untouched, so the level is rebuilt in `dst` and ANDed into `bf`, then swapped back
    const HashTables* tabs = nullptr;
    int rc = get_tables(ctx, k, &tabs);
    if (rc) return rc;
    const nts_view* v = nullptr;
    if ((rc = get_plain_view(g, k, &v))) return rc;
    if (mode == APPLY_AND) {
        if ((rc = bf_fill(dst, 0)) || (rc = bf_insert_direct(ctx, dst, g, v, tabs)) || (rc = bf_combine(bf, dst, 0, false))) return rc;
        std::swap(dst->words, bf->words);
    } else {
        if (mode != APPLY_OR && (rc = bf_fill(bf, 0))) return rc;
        if ((rc = bf_insert_direct(ctx, bf, g, v, tabs))) return rc;
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

int nts_bf_insert_genome_async(nts_bf* bf, const nts_genome* g, uint32_t k)
{
    if (!bf || !g) return fail(NTS_ERR_ARG, "null argument");
    if (bf->ctx != g->ctx) return fail(NTS_ERR_ARG, "filter and genome live on different contexts");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    return bf_insert_mode(bf, nullptr, g, k, APPLY_OR);
}

int nts_bf_insert_genome(nts_bf* bf, const nts_genome* g, uint32_t k)
{
    int rc = nts_bf_insert_genome_async(bf, g, k);
    if (rc) return rc;
    return bf_finish_insert(bf, nullptr, g, k, APPLY_OR);
}

/* bf = bits(genome): like nts_bf_clear + nts_bf_insert_genome without the zero-fill pass */
int nts_bf_set_genome(nts_bf* bf, const nts_genome* g, uint32_t k)
{
    if (!bf || !g) return fail(NTS_ERR_ARG, "null argument");
    if (bf->ctx != g->ctx) return fail(NTS_ERR_ARG, "filter and genome live on different contexts");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    int rc = bf_insert_mode(bf, nullptr, g, k, APPLY_SET);
    if (rc) return rc;
    return bf_finish_insert(bf, nullptr, g, k, APPLY_SET);
}

static int bf_combine(nts_bf* dst, const nts_bf* src, int op, bool sync)
{
    if (!dst || !src) return fail(NTS_ERR_ARG, "null argument");
    if (dst->bytes != src->bytes) return fail(NTS_ERR_ARG, "Bloom filters differ in size");
    if (dst->ctx != src->ctx) return fail(NTS_ERR_ARG, "filters live on different contexts");
    nts_ctx* ctx = dst->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    uint64_t n16 = dst->alloc_bytes / 16;
    ProfScope prof(ctx, PROF_BF_COMBINE, (double)dst->alloc_bytes);
    bf_combine_kernel<<<grid_for(ctx, n16 / 4 + 1, 256, 16), 256, 0, ctx->stream>>>(
        reinterpret_cast<uint4*>(dst->words.p), reinterpret_cast<const uint4*>(src->words.p), n16, op);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    if (sync) NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

int nts_bf_and(nts_bf* dst, const nts_bf* src) { return bf_combine(dst, src, 0, true); }
int nts_bf_or(nts_bf* dst, const nts_bf* src) { return bf_combine(dst, src, 1, true); }
int nts_bf_and_async(nts_bf* dst, const nts_bf* src) { return bf_combine(dst, src, 0, false); }

/* src/ntsynt_make_common_bf.cpp:107-160 in one call: common = AND over the genomes of bits(genome), genomes in the
 * caller's (sorted-path) order.  Genome 0 goes into `common`; every further genome is built in `level` (the cascade
 * level) and ANDed into `common`.  On return `common` holds the result and `level` is scratch. */
static int bf_build_common(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k, bool lazy,
                           int* apart_out);

int nts_bf_build_common(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k)
{
    return bf_build_common(common, level, genomes, n, k, false, nullptr);
}

/* The same without the last AND pass: on return `common` = AND over genomes 0 .. n-2 and `level` = bits(genome n-1); their
 * AND is the common filter.  nts_sketch2 takes the pair; nts_bf_and(common, level) makes it one filter when one is needed. */
int nts_bf_build_common_lazy(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k,
                             int* level_is_apart)
{
    if (!level_is_apart) return fail(NTS_ERR_ARG, "null argument");
    return bf_build_common(common, level, genomes, n, k, true, level_is_apart);
}

static int bf_build_common(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k, bool lazy,
                           int* apart_out)
{
    if (apart_out) *apart_out = 0;
    if (!common || !genomes || n < 1 || (n > 1 && !level)) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = common->ctx;
    if (n > 1 && (level->ctx != ctx || level->bytes != common->bytes)) return fail(NTS_ERR_ARG, "common and level filters differ");
    for (uint32_t i = 0; i < n; ++i)
        if (!genomes[i] || genomes[i]->ctx != ctx) return fail(NTS_ERR_ARG, "bad genome");
    NTS_CUDA(cudaSetDevice(ctx->device));
    ProfScope prof(ctx, PROF_BF_BUILD, 0.0, true);
    // the zero-fill of `level` (a pure HBM write stream) runs on the side stream beside genome 0's insert, whose two
    // passes sit on atomic issue and use a few per cent of the DRAM bandwidth
    cudaEvent_t filled = nullptr;
    const char* e3 = getenv("NTS_BF_IMPL");
    if (n > 1 && !(e3 && e3[0] == '3')) {
        if (!ctx->stream2) NTS_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        NTS_CUDA(cudaEventRecord(ctx->ev_side, ctx->stream));            // after whatever still uses `level` on the main stream
        NTS_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_side, 0));
        int rcf = bf_fill(level, 0, ctx->stream2);
        if (rcf) return rcf;
        NTS_CUDA(cudaEventRecord(ctx->ev_side, ctx->stream2));
        filled = ctx->ev_side;
    }
    int rc = bf_insert_mode(common, nullptr, genomes[0], k, APPLY_SET);
    if (rc || (rc = bf_finish_insert(common, nullptr, genomes[0], k, APPLY_SET))) return rc;
    for (uint32_t i = 1; i < n; ++i) {
        const bool apart = lazy && i + 1 == n && !(e3 && e3[0] == '3');           // the last level stays a filter of its own
        if ((rc = bf_insert_mode(level, common, genomes[i], k, apart ? APPLY_SET : APPLY_AND, i == 1 ? filled : nullptr)) ||
            (rc = bf_finish_insert(level, common, genomes[i], k, apart ? APPLY_SET : APPLY_AND)))
            return rc;
        if (apart && apart_out) *apart_out = 1;
    }
    return NTS_OK;
}

uint64_t nts_part_inserts(const nts_ctx* ctx) { return ctx ? ctx->part_inserts : 0; }
uint64_t nts_part_overflow_items(const nts_ctx* ctx) { return ctx ? ctx->part_overflow_items : 0; }

int nts_bf_insert_repeats(nts_bf* rep, nts_bf* scratch, const nts_genome* g, uint32_t k)
{
    if (!rep || !scratch || !g) return fail(NTS_ERR_ARG, "null argument");
    if (rep->bytes != scratch->bytes) return fail(NTS_ERR_ARG, "Bloom filters differ in size");
    nts_ctx* ctx = rep->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    int rc = bf_fill(scratch, 0);
    if (rc) return rc;
    const HashTables* tabs = nullptr;
    rc = get_tables(ctx, k, &tabs);
    if (rc) return rc;
    const nts_view* v = nullptr;
    rc = get_plain_view(g, k, &v);
    if (rc) return rc;
    if (v->total_valid) {
        uint64_t m, mp;
        mod_params(rep, &m, &mp);
        constexpr int THREADS = 256;
        const uint32_t chunk = 32;
        uint64_t blocks = (v->total_valid + (uint64_t)THREADS * chunk - 1) / ((uint64_t)THREADS * chunk);
        ProfScope prof(ctx, PROF_BF_REPEAT, (double)v->total_valid);
        bf_repeat_kernel<THREADS><<<(unsigned)blocks, THREADS, 0, ctx->stream>>>(
            device_view(g, v), tabs, scratch->words.p, rep->words.p, m, mp, v->total_valid, chunk);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

int nts_bf_popcount(nts_bf* bf, uint64_t* bits_set)
{
    if (!bf || !bits_set) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = bf->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    DevBuf<unsigned long long> acc;
    if (acc.alloc(1) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed");
    NTS_CUDA(cudaMemsetAsync(acc.p, 0, 8, ctx->stream));
    uint64_t n16 = bf->alloc_bytes / 16;
    ProfScope prof(ctx, PROF_POPCOUNT, (double)bf->alloc_bytes);
    bf_popcount_kernel<<<grid_for(ctx, n16, 256, 16), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(bf->words.p),
                                                                            n16, acc.p);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    NTS_CUDA(cudaMemcpyAsync(&h, acc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *bits_set = h;
    return NTS_OK;
}

int nts_bf_download(nts_bf* bf, uint8_t* bytes_out)
{
    if (!bf || !bytes_out) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    NTS_CUDA(copy_d2h(bf->ctx, bytes_out, bf->words.p, bf->bytes));
    NTS_CUDA(cudaStreamSynchronize(bf->ctx->stream));
    return NTS_OK;
}

int nts_bf_upload(nts_bf* bf, const uint8_t* bytes_in)
{
    if (!bf || !bytes_in) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(bf->ctx->device));
    NTS_CUDA(copy_h2d(bf->ctx, bf->words.p, bytes_in, bf->bytes));
    NTS_CUDA(cudaStreamSynchronize(bf->ctx->stream));
    return NTS_OK;
}

// ------------------------------------------------------------------------------------ sketch
static size_t sketch_smem_bytes(uint32_t NT, int threads)
{
    size_t idx = (size_t)((NT + 3) & ~3u) * 2;
    return (size_t)NT * 8 + 2 * idx + sizeof(HashTables) + (size_t)(threads / 32) * sizeof(SegAgg) +
           (size_t)(threads / 32 + 2) * 4;
}

int nts_sketch(nts_ctx* ctx, const nts_genome* g, const nts_bf* common, const nts_bf* repeat, uint32_t k, uint32_t w,
               const uint64_t* mask_off, const uint64_t* mask_start, const uint64_t* mask_end, nts_mxs** out)
{
    return nts_sketch2(ctx, g, common, nullptr, repeat, k, w, mask_off, mask_start, mask_end, out);
}

/* nts_sketch with the common filter given as TWO filters of equal size whose AND it is (nts_bf_build_common_lazy leaves
 * the last cascade level apart): a candidate is looked up in the second only when the first holds it, which costs less
 * than one more pass over both 14.8 GB arrays. */
int nts_sketch2(nts_ctx* ctx, const nts_genome* g, const nts_bf* common, const nts_bf* common2, const nts_bf* repeat, uint32_t k,
                uint32_t w, const uint64_t* mask_off, const uint64_t* mask_start, const uint64_t* mask_end, nts_mxs** out)
{
    if (!ctx || !g || !out) return fail(NTS_ERR_ARG, "null argument");
    if (g->ctx != ctx || (common && common->ctx != ctx) || (repeat && repeat->ctx != ctx) || (common2 && common2->ctx != ctx))
        return fail(NTS_ERR_ARG, "objects live on different contexts");
    if (common2 && (!common || common->bytes != common2->bytes)) return fail(NTS_ERR_ARG, "the two parts of the common filter differ in size");
    if (w < 1) return fail(NTS_ERR_ARG, "w must be >= 1");
    NTS_CUDA(cudaSetDevice(ctx->device));
    constexpr int THREADS = 512;
    // dense tiles: as many slots as fit two CTAs per SM; wide windows fall back to one CTA per SM
    int max_optin = 0;
    NTS_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    uint32_t NT = 8960;
    if (w > 4096) NT = 18432;
    if (2ull * w > NT || sketch_smem_bytes(NT, THREADS) > (size_t)max_optin)
        return fail(NTS_ERR_ARG, "w too large for the shared-memory window selector (max 9216)");
    uint32_t T = NT - w;
    const HashTables* tabs = nullptr;
    int rc = get_tables(ctx, k, &tabs);
    if (rc) return rc;
    const nts_view* v = nullptr;
    nts_view* owned = nullptr;
    if (mask_off) {
        rc = build_view(g, k, mask_off, mask_start, mask_end, &owned);
        if (rc) return rc;
        v = owned;
    } else {
        rc = get_plain_view(g, k, &v);
        if (rc) return rc;
    }
    struct ViewGuard { nts_view* p; ~ViewGuard() { delete p; } } guard{owned};
    if ((rc = wait_ready(g))) return rc;

    uint64_t m = 0, mp = 0;
    uint64_t rm = 0, rmp = 0;       // the repeat filter's own size (sized from another genome: bin/ntsynt_make_repeat_bfs.py:51)
    if (repeat) mod_params(repeat, &rm, &rmp);
    if (common) mod_params(common, &m, &mp); else { m = rm; mp = rmp; }
    // how often does a k-mer of this view pass the filter?  (sampled; only steers the candidate density)
    double pass = 1.0;
    if ((common || repeat) && w >= 128 && v->total_valid > (1u << 20)) {
        DevBuf<unsigned int> d_cnt2;
        if (d_cnt2.alloc(2) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (sample)");
        NTS_CUDA(cudaMemsetAsync(d_cnt2.p, 0, 8, ctx->stream));
        const uint64_t n_samp = 16384, stride = std::max<uint64_t>(1, v->total_valid / n_samp);
        {
            ProfScope prof(ctx, PROF_SKETCH, 0.0);
            sketch_sample_kernel<<<(unsigned)((n_samp + 255) / 256), 256, 0, ctx->stream>>>(
                device_view(g, v), tabs, common ? common->words.p : nullptr, common2 ? common2->words.p : nullptr,
                repeat ? repeat->words.p : nullptr, m, mp, rm, rmp, v->total_valid, stride, d_cnt2.p);
            ctx->launches++;
        }
        NTS_CUDA(cudaGetLastError());
        unsigned int hc[2] = {0, 0};
        NTS_CUDA(cudaMemcpyAsync(hc, d_cnt2.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (hc[1]) pass = std::max(1e-4, (double)hc[0] / (double)hc[1]);
    }
    // ~19 SURVIVING candidates per window keep unresolved windows rare (e^-19); below a 10 % pass rate the
    // candidate list would be as long as the window, and the dense selector (every slot queried) is the better tool
    double lambda = std::max(24.0, 19.0 / pass);
    if (const char* env = getenv("NTS_SKETCH_LAMBDA")) { double x = atof(env); if (x >= 4.0 && x <= 256.0) lambda = x; }
    // sparse tiles (sketch_sparse_kernel): R dense tiles each, sized so that a thread stages ~3 candidates
    constexpr int SCAP = 16, CCAP = 3072;
    bool sparse = w >= 128 && lambda <= 192.0 && lambda / (double)w <= 0.25;
    // a filter that passes < 10 % of the k-mers: no hash threshold keeps ~19 survivors per window, so every slot is
    // looked up and only the survivors are listed (sketch_sparse_kernel<.., QALL>); 16 slots per thread = the capacity
    // of a staging column, tiles of 8192 slots (so w <= 4096)
    bool qall = !sparse && (common || repeat) && w >= 128 && w <= 4096 && pass < 0.1 && !getenv("NTS_SKETCH_LAMBDA");
    if (const char* env = getenv("NTS_SKETCH_QALL")) { if (env[0] == '0') qall = false; }
    if (const char* env = getenv("NTS_SKETCH_DENSE")) { if (env[0] == '1') sparse = qall = false; }
    uint32_t R = 1, NT_s = 0, C_s = 0, tau_hi = 0;
    if (qall) {
        sparse = true;
        C_s = SCAP;
        NT_s = THREADS * SCAP;
        T = NT_s - w;
        R = 1;
    } else if (sparse) {
        const double density = lambda / (double)w;
        uint32_t c_target = (uint32_t)(3.1 / density);
        c_target = std::max<uint32_t>(4, std::min<uint32_t>(c_target, 120));
        // a thread's run = its seed + 16 j roll steps: hash_run then stays in its unrolled 16-step path (the generic
        // remainder loop costs ~4 instructions per k-mer of the whole kernel at 120 slots per thread)
        if (c_target >= 17) c_target = 16 * ((c_target - 1) / 16) + 1;
        const uint32_t nt_target = std::max<uint32_t>(THREADS * c_target, 2 * w);
        const uint32_t ts_target = nt_target - w;
        R = (ts_target + T - 1) / T;
        T = std::max<uint32_t>(ts_target / R, 1);                 // dense sub-tile size; R * T window ends per sparse tile
        NT_s = R * T + w;
        C_s = (NT_s + THREADS - 1) / THREADS;
        tau_hi = (uint32_t)std::min(density * 4294967296.0, 4294967295.0);
        if (C_s > 255 || NT_s > 65535) sparse = false, T = NT - w, R = 1;
    }


    // tiles: one per R * T window ends of a contig; each stands for up to R output slots (dense sub-tiles)
    std::vector<TileDesc> tiles;
    uint32_t n_slots = 0;
    for (uint32_t c = 0; c < g->n_contigs; ++c) {
        const uint64_t v0 = v->contig_v[c], v1 = v->contig_v[c + 1];
        const uint64_t nv = v1 - v0;
        if (nv < w) continue;
        const uint64_t n_win = nv - w + 1;
        const uint64_t TS = (uint64_t)R * T;
        for (uint64_t t = 0; t * TS < n_win; ++t) {
            TileDesc td;
            td.vfirst = v0 + t * TS;
            td.vend = v1;
            td.cbase = g->contig_word_off[c] * 32;
            td.contig = c;
            td.has_prev = t > 0;
            td.out_slot = n_slots;
            td.n_sub = (uint32_t)((std::min<uint64_t>(TS, n_win - t * TS) + T - 1) / T);
            n_slots += td.n_sub;
            tiles.push_back(td);
        }
    }
    nts_mxs* mx = new (std::nothrow) nts_mxs();
    if (!mx) return fail(NTS_ERR_NOMEM, "host allocation failed");
    mx->ctx = ctx;
    mx->n_contigs = g->n_contigs;
    struct MxGuard { nts_mxs* p; ~MxGuard() { delete p; } } mguard{mx};
    if (tiles.empty()) {
        mx->count = 0;
        *out = mx; mguard.p = nullptr;
        return NTS_OK;
    }
    const uint32_t n_tiles = (uint32_t)tiles.size();
    DevBuf<TileDesc> d_tiles, d_esc;
    DevBuf<uint32_t> d_off, d_cnt;
    DevBuf<uint64_t> d_dst;
    DevBuf<unsigned long long> d_total;      // [0] minimizers so far, [1] (low word) escalated dense tiles
    if (d_tiles.alloc(n_tiles) != cudaSuccess || d_off.alloc(n_slots) != cudaSuccess || d_cnt.alloc(n_slots) != cudaSuccess ||
        d_dst.alloc(n_slots) != cudaSuccess || d_total.alloc(2) != cudaSuccess || (sparse && d_esc.alloc(n_slots) != cudaSuccess))
        return fail(NTS_ERR_NOMEM, "device allocation failed (tiles)");
    NTS_CUDA(copy_h2d(ctx, d_tiles.p, tiles.data(), (size_t)n_tiles * sizeof(TileDesc)));

    // dense kernel's own pruning threshold: ~48 slots per window are expected below tau (uniform hashes)
    uint64_t tau = KEY_MAX;
    if ((common || repeat) && w > 96 && pass >= 0.1) tau = (uint64_t)((48.0 / (double)w) * 18446744073709551616.0);
    if (const char* env = getenv("NTS_SKETCH_NO_PRUNE")) { if (env[0] == '1') tau = KEY_MAX; }
    const size_t smem = sketch_smem_bytes(NT, THREADS);
    NTS_CUDA(cudaFuncSetAttribute(sketch_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t smem_s = sizeof(HashTables) + (size_t)SCAP * THREADS * 9 + (size_t)CCAP * 10;
    if (sparse && qall)
        NTS_CUDA(cudaFuncSetAttribute(sketch_sparse_kernel<THREADS, SCAP, CCAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_s));
    else if (sparse)
        NTS_CUDA(cudaFuncSetAttribute(sketch_sparse_kernel<THREADS, SCAP, CCAP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_s));

    // query-everything mode on a nearly empty filter: one bit per 32-byte sector (L2-resident) answers most lookups.
    // Rebuilt per call (one streaming read of the filter, ~2.3 ms at 14.8 GB): nothing to keep coherent with later inserts.
    DevBuf<uint32_t> d_summary;
    uint32_t sum_shift = 10;         // one bit per 128-byte line of the filter (14.5 MB for 14.8 GB): measured best of 8 / 10 / 12
    if (qall && common && pass < 0.01 && !(getenv("NTS_SKETCH_SUMMARY") && getenv("NTS_SKETCH_SUMMARY")[0] == '0')) {
        if (const char* es = getenv("NTS_SKETCH_SUMMARY_SHIFT")) { const int x = atoi(es); if (x >= 8 && x <= 14) sum_shift = (uint32_t)x; }
        const uint64_t unit_bytes = 1ull << (sum_shift - 3);
        const uint64_t n_sectors = common->alloc_bytes / unit_bytes;      // whole units; what lies past them reads as "look"
        if (n_sectors && d_summary.alloc((n_sectors + 31) / 32 + 1) == cudaSuccess) {
            NTS_CUDA(cudaMemsetAsync(d_summary.p + (n_sectors + 31) / 32, 0xFF, 4, ctx->stream));
            ProfScope prof(ctx, PROF_SKETCH, 0.0);
            bf_summary_kernel<<<grid_for(ctx, n_sectors, 256, 8), 256, 0, ctx->stream>>>(
                reinterpret_cast<const uint4*>(common->words.p), n_sectors, (uint32_t)(unit_bytes / 16), d_summary.p);
            ctx->launches++;
            NTS_CUDA(cudaGetLastError());
        } else {
            cudaGetLastError();
            d_summary.release();
        }
    }
    // expected density of minimizers is 2/(w+1) per window; start with 2.5x that
    uint64_t n_win_total = 0;
    for (uint32_t c = 0; c < g->n_contigs; ++c) {
        uint64_t nv = v->contig_v[c + 1] - v->contig_v[c];
        if (nv >= w) n_win_total += nv - w + 1;
    }
    uint64_t cap = (uint64_t)(5.0 * (double)n_win_total / (double)(w + 1)) + 4096 + n_slots;
    cap = std::min<uint64_t>(cap, n_win_total);
    DevBuf<uint64_t> u_h1;
    DevBuf<uint32_t> u_pos, u_ctg;
    unsigned long long total = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (cap > 0xFFFFFFF0ull) return fail(NTS_ERR_OVERFLOW, "more than 2^32 minimizers in one sketch");
        if (u_h1.alloc(cap) != cudaSuccess || u_pos.alloc(cap) != cudaSuccess || u_ctg.alloc(cap) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (minimizer buffer)");
        NTS_CUDA(cudaMemsetAsync(d_total.p, 0, 16, ctx->stream));
        SketchOut so;
        so.h1 = u_h1.p; so.pos = u_pos.p; so.contig = u_ctg.p;
        so.tile_off = d_off.p; so.tile_cnt = d_cnt.p; so.total = d_total.p; so.cap = cap;
        {
            ProfScope prof(ctx, PROF_SKETCH, (double)v->total_valid);
            const GenomeView gv = device_view(g, v);
            const uint32_t* cw = common ? common->words.p : nullptr;
            const uint32_t* cw2 = common2 ? common2->words.p : nullptr;
            const uint32_t* rw = repeat ? repeat->words.p : nullptr;
            if (sparse) {
                unsigned int* esc_count = reinterpret_cast<unsigned int*>(d_total.p + 1);
                if (qall && attempt == 0) ctx->sketch_qall++;
                if (qall)
                    sketch_sparse_kernel<THREADS, SCAP, CCAP, true><<<n_tiles, THREADS, smem_s, ctx->stream>>>(
                        gv, tabs, cw, d_summary.p, cw2, rw, m, mp, rm, rmp, d_tiles.p, w, NT_s, C_s, T, tau_hi, so, d_esc.p, esc_count, sum_shift);
                else
                    sketch_sparse_kernel<THREADS, SCAP, CCAP, false><<<n_tiles, THREADS, smem_s, ctx->stream>>>(
                        gv, tabs, cw, nullptr, cw2, rw, m, mp, rm, rmp, d_tiles.p, w, NT_s, C_s, T, tau_hi, so, d_esc.p, esc_count, 8u);
                ctx->launches++;
                NTS_CUDA(cudaGetLastError());
                unsigned int n_esc = 0;
                NTS_CUDA(cudaMemcpyAsync(&n_esc, esc_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
                NTS_CUDA(cudaStreamSynchronize(ctx->stream));
                ctx->sketch_escalated += n_esc;
                if (n_esc) {       // tiles with an unresolved window: the dense selector, every slot queried
                    sketch_kernel<THREADS><<<n_esc, THREADS, smem, ctx->stream>>>(gv, tabs, cw, cw2, rw, m, mp, rm, rmp, d_esc.p, w,
                                                                                 T, KEY_MAX, so);
                    ctx->launches++;
                }
            } else {
                sketch_kernel<THREADS><<<n_tiles, THREADS, smem, ctx->stream>>>(gv, tabs, cw, cw2, rw, m, mp, rm, rmp, d_tiles.p, w, T, tau, so);
                ctx->launches++;
            }
        }
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaMemcpyAsync(&total, d_total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (total <= cap) break;
        if (attempt == 1) return fail(NTS_ERR_OVERFLOW, "minimizer buffer overflow after retry");
        cap = total;
    }
    mx->count = total;
    if (mx->h1.alloc(total) != cudaSuccess || mx->pos.alloc(total) != cudaSuccess || mx->contig.alloc(total) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (minimizer table)");
    ProfScope prof_post(ctx, PROF_SKETCH_POST, (double)total);
    tile_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt.p, d_dst.p, n_slots);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    {
        uint64_t threads_needed = (uint64_t)n_slots * 32;
        unsigned blocks = (unsigned)((threads_needed + 255) / 256);
        sketch_gather_kernel<<<blocks, 256, 0, ctx->stream>>>(u_h1.p, u_pos.p, u_ctg.p, d_off.p, d_cnt.p, d_dst.p, n_slots,
                                                             mx->h1.p, mx->pos.p, mx->contig.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = mx; mguard.p = nullptr;
    return NTS_OK;
}

void nts_mxs_destroy(nts_mxs* m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    delete m;
}

uint64_t nts_mxs_count(const nts_mxs* m) { return m ? m->count : 0; }

int nts_mxs_download(nts_mxs* m, uint64_t* h1, uint32_t* pos, uint32_t* contig)
{
    if (!m) return fail(NTS_ERR_ARG, "null argument");
    if (m->count == 0) return NTS_OK;
    NTS_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    if (h1) NTS_CUDA(copy_d2h(m->ctx, h1, m->h1.p, m->count * 8));
    if (pos) NTS_CUDA(copy_d2h(m->ctx, pos, m->pos.p, m->count * 4));
    if (contig) NTS_CUDA(copy_d2h(m->ctx, contig, m->contig.p, m->count * 4));
    NTS_CUDA(cudaStreamSynchronize(st));
    return NTS_OK;
}

int nts_hash_contig(nts_ctx* ctx, const nts_genome* g, uint32_t contig, uint32_t k, uint64_t* h0_out, uint8_t* valid_out)
{
    if (!ctx || !g || contig >= g->n_contigs || !h0_out || !valid_out) return fail(NTS_ERR_ARG, "bad argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    const HashTables* tabs = nullptr;
    int rc = get_tables(ctx, k, &tabs);
    if (rc) return rc;
    const nts_view* v = nullptr;
    rc = get_plain_view(g, k, &v);
    if (rc) return rc;
    const uint64_t len = g->contig_len[contig];
    const uint64_t nk = len >= k ? len - k + 1 : 0;
    if (nk == 0) return NTS_OK;
    DevBuf<uint64_t> d_h;
    DevBuf<uint8_t> d_v;
    if (d_h.alloc(nk) != cudaSuccess || d_v.alloc(nk) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed");
    NTS_CUDA(cudaMemsetAsync(d_h.p, 0, nk * 8, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(d_v.p, 0, nk, ctx->stream));
    const uint64_t v0 = v->contig_v[contig], v1 = v->contig_v[contig + 1];
    if (v1 > v0) {
        constexpr int THREADS = 256;
        const uint32_t chunk = 32;
        uint64_t blocks = (v1 - v0 + (uint64_t)THREADS * chunk - 1) / ((uint64_t)THREADS * chunk);
        hash_dump_kernel<THREADS><<<(unsigned)blocks, THREADS, 0, ctx->stream>>>(
            device_view(g, v), tabs, v0, v1, g->contig_word_off[contig] * 32, d_h.p, d_v.p, chunk);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaMemcpyAsync(h0_out, d_h.p, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(valid_out, d_v.p, nk, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

}  // extern "C"

extern "C" int nts_mxs_upload(nts_ctx* ctx, uint64_t n, const uint64_t* h1, const uint32_t* pos, const uint32_t* contig,
                              nts_mxs** out)
{
    if (!ctx || !out || (n && (!h1 || !pos || !contig))) return nts::fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    nts_mxs* m = new (std::nothrow) nts_mxs();
    if (!m) return nts::fail(NTS_ERR_NOMEM, "host allocation failed");
    m->ctx = ctx; m->count = n;
    const uint64_t a = n ? n : 1;
    if (m->h1.alloc(a) != cudaSuccess || m->pos.alloc(a) != cudaSuccess || m->contig.alloc(a) != cudaSuccess) {
        delete m;
        return nts::fail(NTS_ERR_NOMEM, "device allocation failed (minimizer table)");
    }
    if (n) {
        NTS_CUDA(nts::copy_h2d(ctx, m->h1.p, h1, n * 8));
        NTS_CUDA(nts::copy_h2d(ctx, m->pos.p, pos, n * 4));
        NTS_CUDA(nts::copy_h2d(ctx, m->contig.p, contig, n * 4));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *out = m;
    return NTS_OK;
}

namespace nts {
// first row of every contig in a table sorted by (contig, position): off[c] = lower_bound(contig[], c)
__global__ void mxs_contig_offsets_kernel(const uint32_t* __restrict__ contig, uint64_t n, uint32_t n_contigs, uint64_t* __restrict__ off)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (contig[mid] < c) lo = mid + 1; else hi = mid; }
    off[c] = lo;
}
}  // namespace nts

/* off[n_contigs + 1]: row range [off[c], off[c+1]) of every contig of a table in (contig, position) order */
extern "C" int nts_mxs_contig_offsets(nts_mxs* m, uint32_t n_contigs, uint64_t* off)
{
    if (!m || !off) return nts::fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = m->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    if (!m->count) { for (uint32_t c = 0; c <= n_contigs; ++c) off[c] = 0; return NTS_OK; }
    nts::DevBuf<uint64_t> d;
    if (d.alloc(n_contigs + 1) != cudaSuccess) return nts::fail(NTS_ERR_NOMEM, "device allocation failed");
    nts::mxs_contig_offsets_kernel<<<(n_contigs + 1 + 127) / 128, 128, 0, ctx->stream>>>(m->contig.p, m->count, n_contigs, d.p);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(nts::copy_d2h(ctx, off, d.p, (size_t)(n_contigs + 1) * 8));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* A new table made of row ranges [src_off[i], src_off[i] + cnt[i]) of parts[i], in the given order (contig-sharded
 * runs: every contig of a genome was sketched on the rank that owns it; the pieces are put back in contig order). */
extern "C" int nts_mxs_concat(nts_ctx* ctx, nts_mxs* const* parts, const uint64_t* src_off, const uint64_t* cnt, uint64_t n_parts,
                              nts_mxs** out)
{
    if (!ctx || !out || (n_parts && (!parts || !src_off || !cnt))) return nts::fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(ctx->device));
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_parts; ++i) {
        if (!parts[i] || parts[i]->ctx != ctx || src_off[i] + cnt[i] > parts[i]->count) return nts::fail(NTS_ERR_ARG, "bad part");
        total += cnt[i];
    }
    nts_mxs* m = new (std::nothrow) nts_mxs();
    if (!m) return nts::fail(NTS_ERR_NOMEM, "host allocation failed");
    m->ctx = ctx; m->count = total;
    const uint64_t a = total ? total : 1;
    if (m->h1.alloc(a) != cudaSuccess || m->pos.alloc(a) != cudaSuccess || m->contig.alloc(a) != cudaSuccess) {
        delete m;
        return nts::fail(NTS_ERR_NOMEM, "device allocation failed (minimizer table)");
    }
    uint64_t at = 0;
    for (uint64_t i = 0; i < n_parts; ++i) {
        if (!cnt[i]) continue;
        cudaError_t e = cudaMemcpyAsync(m->h1.p + at, parts[i]->h1.p + src_off[i], cnt[i] * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->pos.p + at, parts[i]->pos.p + src_off[i], cnt[i] * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->contig.p + at, parts[i]->contig.p + src_off[i], cnt[i] * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { delete m; return nts::fail(NTS_ERR_CUDA, cudaGetErrorString(e)); }
        at += cnt[i];
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return NTS_OK;
}

namespace nts {
// keep[i] = the k-mer of minimizer i (contig, pos) is NOT in the filter: ntHash2 of the k-mer from the seed tables
// (a minimizer is an all-ACGT k-mer), bit h0 mod m
__global__ void mxs_bf_keep_kernel(const uint64_t* __restrict__ packed, const uint64_t* __restrict__ contig_base /* base index of every contig */,
                                   const HashTables* __restrict__ tabs, uint32_t k, const uint32_t* __restrict__ pos,
                                   const uint32_t* __restrict__ contig, uint64_t n, const uint32_t* __restrict__ bits, uint64_t m,
                                   uint64_t mprime, uint32_t* __restrict__ keep)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = contig_base[contig[i]] + pos[i];
    uint64_t fwd = 0, rev = 0;
    for (uint32_t j = 0; j < k; ++j) {
        const unsigned c = base_at(packed, b + j);
        fwd ^= __ldg(&tabs->init_f[j * 4 + c]);
        rev ^= __ldg(&tabs->init_r[j * 4 + c]);
    }
    const uint64_t idx = fast_mod(fwd + rev, m, mprime);
    keep[i] = ((__ldg(&bits[idx >> 5]) >> (idx & 31)) & 1u) ? 0u : 1u;
}

__global__ void mxs_compact_kernel(const uint64_t* __restrict__ h1, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ctg,
                                   const uint32_t* __restrict__ keep, uint64_t n, unsigned long long* __restrict__ cursor_unused,
                                   const uint64_t* __restrict__ off, uint64_t* __restrict__ o_h1, uint32_t* __restrict__ o_pos,
                                   uint32_t* __restrict__ o_ctg)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint64_t o = off[i];
    o_h1[o] = h1[i]; o_pos[o] = pos[i]; o_ctg[o] = ctg[i];
}

// exclusive prefix of 0/1 flags into 64-bit offsets (single CTA pass per 2^20 block is plenty for minimizer tables)
__global__ void flags_scan_kernel(const uint32_t* __restrict__ keep, uint64_t n, uint64_t* __restrict__ off, unsigned long long* __restrict__ total)
{
    __shared__ uint64_t s_part[1024];
    const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint64_t a = threadIdx.x * per, b = min(a + per, n);
    uint64_t sum = 0;
    for (uint64_t i = a; i < b; ++i) sum += keep[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) { const uint64_t v = s_part[i]; s_part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    uint64_t run = s_part[threadIdx.x];
    for (uint64_t i = a; i < b; ++i) { off[i] = run; run += keep[i]; }
}
}  // namespace nts

/* read_minimizers(tsv, repeat_bf) of the graph stage's --filter Filter mode (subprojects/ntJoin/bin/ntjoin_utils.py:182):
 * a minimizer whose k-mer is in the repeat filter is dropped from its file.  out = the entries of `m` (a table of
 * genome g, in (contig, position) order) whose k-mer is not in `bf`. */
extern "C" int nts_mxs_drop_in_bf(nts_ctx* ctx, const nts_mxs* m, const nts_genome* g, const nts_bf* bf, uint32_t k, nts_mxs** out)
{
    using namespace nts;
    if (!ctx || !m || !g || !bf || !out) return fail(NTS_ERR_ARG, "null argument");
    if (m->ctx != ctx || g->ctx != ctx || bf->ctx != ctx) return fail(NTS_ERR_ARG, "objects live on different contexts");
    NTS_CUDA(cudaSetDevice(ctx->device));
    const HashTables* tabs = nullptr;
    int rc = get_tables(ctx, k, &tabs);
    if (rc || (rc = wait_ready(g))) return rc;
    const uint64_t n = m->count;
    nts_mxs* o = new (std::nothrow) nts_mxs();
    if (!o) return fail(NTS_ERR_NOMEM, "host allocation failed");
    struct Guard { nts_mxs* p; ~Guard() { delete p; } } guard{o};
    o->ctx = ctx; o->n_contigs = m->n_contigs;
    unsigned long long total = 0;
    DevBuf<uint32_t> keep;
    DevBuf<uint64_t> off, cbase;
    DevBuf<unsigned long long> d_total;
    if (n) {
        std::vector<uint64_t> base(g->n_contigs);
        for (uint32_t c = 0; c < g->n_contigs; ++c) base[c] = g->contig_word_off[c] * 32;
        if (keep.alloc(n) != cudaSuccess || off.alloc(n) != cudaSuccess || cbase.alloc(g->n_contigs) != cudaSuccess || d_total.alloc(1) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (repeat filter)");
        NTS_CUDA(copy_h2d(ctx, cbase.p, base.data(), base.size() * 8));
        uint64_t mm, mp;
        mod_params(bf, &mm, &mp);
        ProfScope prof(ctx, PROF_BF_REPEAT, (double)n);
        mxs_bf_keep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(g->packed.p, cbase.p, tabs, k, m->pos.p, m->contig.p, n,
                                                                                bf->words.p, mm, mp, keep.p);
        flags_scan_kernel<<<1, 1024, 0, ctx->stream>>>(keep.p, n, off.p, d_total.p);
        ctx->launches += 2;
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaMemcpyAsync(&total, d_total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));       // (also: `base` goes out of scope)
    }
    o->count = total;
    const uint64_t a = total ? total : 1;
    if (o->h1.alloc(a) != cudaSuccess || o->pos.alloc(a) != cudaSuccess || o->contig.alloc(a) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (minimizer table)");
    if (total) {
        mxs_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(m->h1.p, m->pos.p, m->contig.p, keep.p, n, nullptr, off.p,
                                                                                o->h1.p, o->pos.p, o->contig.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *out = o; guard.p = nullptr;
    return NTS_OK;
}
