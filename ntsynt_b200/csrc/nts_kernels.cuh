// nts_kernels.cuh -- the sm_100a kernels of the sketch / Bloom-filter half of the path.
//
//   (i)     ntHash2 over 2-bit packed contigs           -> hash_run() in nts_device.cuh
//   (ii)    sliding-window rightmost-minimum selector   -> sketch_kernel (van Herk / Gil-Werman
//           prefix+suffix arg-min over blocks of w keys staged in shared memory, segmented
//           scans done with warp shuffles), replacing btllib::Indexlr::minimize
//   (iii-a) Bloom insert                                -> bf_insert_kernel (RED.OR per k-mer)
//   (iii-b) Bloom merge                                 -> bf_and_kernel (128-bit streaming)
//   (iii-c) Bloom query                                 -> fused into sketch_kernel phase A2
//   (iii-d) repeat filter                               -> bf_repeat_kernel
// All of it is HBM-bound integer work: no tensor cores.
#pragma once
#include "nts_device.cuh"

namespace nts {

// ------------------------------------------------------------------------------ streaming helpers
__global__ void fill_u128_kernel(uint4* __restrict__ p, uint64_t n16, uint32_t v)
{
    const uint4 val = make_uint4(v, v, v, v);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = val;
}

// op: 0 = AND, 1 = OR
__global__ void bf_combine_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, uint64_t n16, int op)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 128-bit loads per operand in flight per thread
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = dst[i + u * stride]; b[u] = __ldg(&src[i + u * stride]); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint4 r;
            if (op == 0) r = make_uint4(a[u].x & b[u].x, a[u].y & b[u].y, a[u].z & b[u].z, a[u].w & b[u].w);
            else         r = make_uint4(a[u].x | b[u].x, a[u].y | b[u].y, a[u].z | b[u].z, a[u].w | b[u].w);
            dst[i + u * stride] = r;
        }
    }
    for (; i < n16; i += stride) {
        uint4 a = dst[i], b = __ldg(&src[i]);
        dst[i] = op == 0 ? make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w)
                         : make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
    }
}

// one bit per unit of 2^shift filter bits (a 32-byte sector or a few of them): is any bit of the unit set?  For a nearly empty filter (the AND of several
// diverged genomes) the summary is 1/256 of the filter -- 58 MB for 14.8 GB, resident in L2 -- and answers most lookups
// of the query-everything sketch without touching HBM (sketch_sparse_kernel<.., QALL>).
__global__ void bf_summary_kernel(const uint4* __restrict__ words, uint64_t n_units, uint32_t per_unit /* 16-byte words per unit */,
                                  uint32_t* __restrict__ summary)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n_units + 31) & ~31ull;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_round; s += stride) {
        bool any = true;                                    // past the last whole unit: "look it up"
        if (s < n_units) {
            uint32_t acc = 0;
            for (uint32_t i = 0; i < per_unit; i += 2) {
                const uint4 a = __ldg(&words[s * per_unit + i]), b = __ldg(&words[s * per_unit + i + 1]);
                acc |= a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w;
            }
            any = acc != 0;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, any);
        if ((threadIdx.x & 31) == 0) summary[s >> 5] = mask;
    }
}

__global__ void bf_popcount_kernel(const uint4* __restrict__ p, uint64_t n16, unsigned long long* __restrict__ out)
{
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = __ldg(&p[i]);
        c += __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w);
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__device__ __forceinline__ void stage_tables(HashTables* s_tabs, const HashTables* __restrict__ g_tabs, uint32_t k)
{
    // roll tables (32 words) + the first k rows of the init tables
    for (uint32_t i = threadIdx.x; i < 16; i += blockDim.x) {
        s_tabs->roll[i] = g_tabs->roll[i];
    }
    for (uint32_t i = threadIdx.x; i < k * 4; i += blockDim.x) {
        s_tabs->init_f[i] = g_tabs->init_f[i];
        s_tabs->init_r[i] = g_tabs->init_r[i];
    }
}

// ------------------------------------------------------------------------------ (i)+(iii-a) insert
// Each thread hashes `chunk` consecutive valid k-mers and sets bit (h0 mod m) with a
// fire-and-forget RED.OR on the 32-bit word (byte = idx>>3, bit = idx&7 in little-endian words).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bf_insert_kernel(GenomeView g, const HashTables* __restrict__ g_tabs,
                                                             uint32_t* __restrict__ bits, uint64_t m, uint64_t mprime,
                                                             uint64_t total_valid, uint32_t chunk)
{
    __shared__ HashTables s_tabs;
    stage_tables(&s_tabs, g_tabs, g.k);
    __syncthreads();
    uint64_t v0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * chunk;
    if (v0 >= total_valid) return;
    uint32_t cnt = (uint32_t)min((uint64_t)chunk, total_valid - v0);
    hash_run(g, &s_tabs, v0, cnt, [&](uint32_t, uint64_t h0, uint64_t) {
        uint64_t idx = fast_mod(h0, m, mprime);
        atomicOr(&bits[idx >> 5], 1u << (idx & 31));
    });
}

// (iii-d) repeat filter: first sighting sets the genome's bit, later sightings set the repeat bit
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bf_repeat_kernel(GenomeView g, const HashTables* __restrict__ g_tabs,
                                                             uint32_t* __restrict__ seen, uint32_t* __restrict__ rep,
                                                             uint64_t m, uint64_t mprime, uint64_t total_valid,
                                                             uint32_t chunk)
{
    __shared__ HashTables s_tabs;
    stage_tables(&s_tabs, g_tabs, g.k);
    __syncthreads();
    uint64_t v0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * chunk;
    if (v0 >= total_valid) return;
    uint32_t cnt = (uint32_t)min((uint64_t)chunk, total_valid - v0);
    hash_run(g, &s_tabs, v0, cnt, [&](uint32_t, uint64_t h0, uint64_t) {
        uint64_t idx = fast_mod(h0, m, mprime);
        uint32_t bit = 1u << (idx & 31);
        uint32_t old = atomicOr(&seen[idx >> 5], bit);
        if (old & bit) atomicOr(&rep[idx >> 5], bit);
    });
}

// test hook for kernel (i): h0 of every valid k-mer of one contig, scattered to position space
template <int THREADS>
__global__ void __launch_bounds__(THREADS) hash_dump_kernel(GenomeView g, const HashTables* __restrict__ g_tabs,
                                                             uint64_t v_begin, uint64_t v_end, uint64_t contig_base,
                                                             uint64_t* __restrict__ h0_out,
                                                             uint8_t* __restrict__ valid_out, uint32_t chunk)
{
    __shared__ HashTables s_tabs;
    stage_tables(&s_tabs, g_tabs, g.k);
    __syncthreads();
    uint64_t v0 = v_begin + ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * chunk;
    if (v0 >= v_end) return;
    uint32_t cnt = (uint32_t)min((uint64_t)chunk, v_end - v0);
    hash_run(g, &s_tabs, v0, cnt, [&](uint32_t, uint64_t h0, uint64_t b) {
        h0_out[b - contig_base] = h0;
        valid_out[b - contig_base] = 1;
    });
}

// ------------------------------------------------------------------------------ (ii) sketch
// One tile = T window ends of one contig, in valid-index space.  Local slot i holds the key of
// valid index vfirst - 1 + i;  slots [0, T + w).  Windows (of w slots) ending at slots
// [w, n_end) are owned; the window ending at slot w-1 only supplies "the previous minimum".
struct TileDesc {
    uint64_t vfirst;     // global valid index of slot 1
    uint64_t vend;       // global valid index one past the contig's last valid k-mer
    uint64_t cbase;      // global base index of the contig's first base
    uint32_t contig;
    uint32_t has_prev;   // 0 for the first tile of a contig (slot 0 is a dummy)
    uint32_t out_slot;   // where the tile's run is recorded in tile_off / tile_cnt (tile order of the output)
    uint32_t n_sub;      // sparse tiles: number of dense tiles (and output slots) the tile stands for
};

struct SketchOut {
    uint64_t* h1;
    uint32_t* pos;
    uint32_t* contig;
    uint32_t* tile_off;            // [n_tiles] where the tile's run starts in the unordered buffer
    uint32_t* tile_cnt;            // [n_tiles]
    unsigned long long* total;     // running total (also the overflow detector)
    uint64_t cap;
};

struct KeyIdx { uint64_t key; uint32_t idx; };

// rightmost arg-min: `r` lies to the right of `l`
__device__ __forceinline__ KeyIdx take_right_if_le(KeyIdx l, KeyIdx r) { return (r.key <= l.key) ? r : l; }

// flag bit0: a block boundary closes this aggregate (nothing flows in from `prev`);
// flag bit1: the aggregate holds a value (empty chunks past the tile end do not)
struct SegAgg { uint64_t key; uint32_t idx; uint32_t flag; };
constexpr uint32_t SEG_CLOSED = 1u, SEG_VALID = 2u;

// inclusive segmented scan across the CTA; dir = +1: left-to-right (carry from lower tids),
// dir = -1: right-to-left (carry from higher tids).  combine(carry, own) semantics are passed in.
template <int THREADS, bool LEFT_TO_RIGHT>
__device__ __forceinline__ SegAgg cta_exclusive_carry(SegAgg own, SegAgg* s_warp /*[THREADS/32]*/)
{
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // logical lane order: for the right-to-left scan mirror the lanes / warps
    const int llane = LEFT_TO_RIGHT ? lane : 31 - lane;
    const int lwid = LEFT_TO_RIGHT ? wid : NW - 1 - wid;
    auto comb = [](SegAgg prev, SegAgg cur) {  // prev precedes cur in scan order
        SegAgg r = cur;
        if (!(cur.flag & SEG_CLOSED) && (prev.flag & SEG_VALID)) {
            bool take_prev;
            if (!(cur.flag & SEG_VALID)) take_prev = true;
            else if (LEFT_TO_RIGHT) take_prev = !(cur.key <= prev.key);   // cur is to the right: wins ties
            else                    take_prev = (prev.key <= cur.key);    // prev is to the right: wins ties
            if (take_prev) { r.key = prev.key; r.idx = prev.idx; }
        }
        r.flag = cur.flag | prev.flag;
        return r;
    };
    SegAgg inc = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        SegAgg o;
        int src = LEFT_TO_RIGHT ? lane - d : lane + d;
        o.key = __shfl_sync(0xffffffffu, inc.key, src & 31);
        o.idx = __shfl_sync(0xffffffffu, inc.idx, src & 31);
        o.flag = __shfl_sync(0xffffffffu, inc.flag, src & 31);
        if (llane >= d) inc = comb(o, inc);
    }
    if (llane == 31) s_warp[lwid] = inc;
    __syncthreads();
    if (wid == 0) {
        SegAgg wv;
        if (lane < NW) wv = s_warp[lane]; else { wv.key = KEY_MAX; wv.idx = 0xFFFFFFFFu; wv.flag = 0; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            SegAgg o;
            o.key = __shfl_up_sync(0xffffffffu, wv.key, d);
            o.idx = __shfl_up_sync(0xffffffffu, wv.idx, d);
            o.flag = __shfl_up_sync(0xffffffffu, wv.flag, d);
            if (lane >= d && lane < NW) wv = comb(o, wv);
        }
        if (lane < NW) s_warp[lane] = wv;
    }
    __syncthreads();
    // exclusive value for this thread
    SegAgg ex;
    {
        int src = LEFT_TO_RIGHT ? lane - 1 : lane + 1;
        ex.key = __shfl_sync(0xffffffffu, inc.key, src & 31);
        ex.idx = __shfl_sync(0xffffffffu, inc.idx, src & 31);
        ex.flag = __shfl_sync(0xffffffffu, inc.flag, src & 31);
    }
    const bool have_lane = llane > 0, have_warp = lwid > 0;
    SegAgg res; res.key = KEY_MAX; res.idx = 0xFFFFFFFFu; res.flag = 0;  // flag without SEG_VALID: "no carry"
    if (have_lane && have_warp) res = comb(s_warp[lwid - 1], ex);
    else if (have_lane) res = ex;
    else if (have_warp) res = s_warp[lwid - 1];
    __syncthreads();   // s_warp is reused by the caller
    return res;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
sketch_kernel(GenomeView g, const HashTables* __restrict__ g_tabs, const uint32_t* __restrict__ common,
              const uint32_t* __restrict__ common2, const uint32_t* __restrict__ repeat, uint64_t m, uint64_t mprime, uint64_t rm, uint64_t rmprime,
              const TileDesc* __restrict__ tiles,
              uint32_t w, uint32_t T, uint64_t tau, SketchOut out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t NT = T + w;
    const uint32_t C = (NT + THREADS - 1) / THREADS;   // slots per thread (contiguous)
    HashTables* s_tabs = reinterpret_cast<HashTables*>(smem_raw);          // first: its uint4 rows need 16-byte alignment
    uint64_t* s_key = reinterpret_cast<uint64_t*>(s_tabs + 1);
    uint16_t* s_P = reinterpret_cast<uint16_t*>(s_key + NT);
    uint16_t* s_S = s_P + ((NT + 3) & ~3u);
    SegAgg* s_warp = reinterpret_cast<SegAgg*>(s_S + ((NT + 3) & ~3u));
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_warp + THREADS / 32);   // [THREADS/32 + 2]

    const TileDesc td = tiles[blockIdx.x];
    stage_tables(s_tabs, g_tabs, g.k);
    __syncthreads();

    const uint64_t vbase = td.vfirst - 1;                         // valid index of slot 0 (may be "-1")
    const uint32_t i_lo = td.has_prev ? 0u : 1u;
    const uint64_t avail = td.vend - td.vfirst + 1;               // slots 0..avail-1 map below vend
    const uint32_t n_end = (uint32_t)min((uint64_t)NT, avail);    // slots [i_lo, n_end) are real

    // Threshold-pruned Bloom queries: a slot whose hash is >= tau cannot be a window minimum unless every
    // slot below tau in that window fails the filter, so pass 0 queries only the slots below tau (the others
    // become UINT64_MAX unqueried).  A window whose minimum is still UINT64_MAX after pass 0 is unresolved;
    // if the tile has one, the whole tile is redone with tau = UINT64_MAX (query everything).  Exact by
    // construction: a resolved window has the same rightmost minimum under both rules.
    const uint32_t c0 = threadIdx.x * C;
    const uint32_t hi = min(c0 + C, NT);
    const uint32_t n_own = n_end > w ? n_end - w : 0;
    const uint32_t per = (n_own + THREADS - 1) / THREADS;
    const uint32_t o0 = w + threadIdx.x * per, o1 = min(o0 + per, n_end);
    uint32_t cnt = 0;
    // window ending at slot i covers [i-w+1, i]
    // `whole` = the window is exactly one block, i.e. (i + 1) % w == 0 (tracked incrementally by the callers)
    auto win_min = [&](uint32_t i, bool whole) -> uint32_t {
        uint32_t p = s_P[i];
        if (whole) return p;
        uint32_t s = s_S[i + 1 - w];
        return (s_key[p] <= s_key[s]) ? p : s;          // p lies to the right: wins ties
    };
    const uint32_t o0_rem = (o0 < o1) ? o0 % w : 0;     // o0 % w, the only modulo of phase C
    const bool filtered = (common != nullptr || repeat != nullptr);
    for (int pass = 0; pass < 2; ++pass) {
    const uint64_t tau_cur = (pass == 0 && filtered) ? tau : KEY_MAX;
    // ---- phase A1: hash (kernel i).  thread t owns slots [t*C, t*C + C)
    {
        uint32_t a = max(c0, i_lo), b = min(c0 + C, n_end);
        for (uint32_t i = c0; i < min(c0 + C, NT); ++i)
            if (i < a || i >= b) s_key[i] = KEY_MAX;
        if (a < b)
            hash_run(g, s_tabs, vbase + a, b - a, [&](uint32_t j, uint64_t h0, uint64_t) { s_key[a + j] = h0; });
    }
    __syncthreads();

    // ---- phase A2: Bloom query (kernel iii-c), 8 independent sector loads in flight per thread
    if (filtered && tau_cur != KEY_MAX) {
        // pruned pass: only slots below tau are looked up (a few per cent); the rest become UINT64_MAX.
        // A thread first gathers up to 4 of its low slots, then issues their sector loads together, so the
        // warp pays one memory latency instead of one per slot.
        uint32_t li[4];
        uint64_t lx[4];
        uint32_t nl = 0;
        // the repeat filter has its own size (made from another genome than the common one's: bin/ntsynt_make_repeat_bfs.py:51)
        auto probe = [&](uint64_t h, uint64_t idx) -> bool {
            bool keep = true;
            if (common) keep = (__ldg(&common[idx >> 5]) >> (idx & 31)) & 1u;
            if (keep && common2) keep = (__ldg(&common2[idx >> 5]) >> (idx & 31)) & 1u;      // common = AND of two filters kept apart
            if (keep && repeat) {
                const uint64_t ridx = fast_mod(h, rm, rmprime);
                keep = !((__ldg(&repeat[ridx >> 5]) >> (ridx & 31)) & 1u);
            }
            return keep;
        };
        for (uint32_t i = max(c0, i_lo); i < min(c0 + C, n_end); ++i) {
            const uint64_t h = s_key[i];
            if (h >= tau_cur) { s_key[i] = KEY_MAX; continue; }
            const uint64_t idx = fast_mod(h, m, mprime);
            if (nl < 4) {
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) if (u == nl) { li[u] = i; lx[u] = idx; }
                ++nl;
            } else if (!probe(h, idx)) {
                s_key[i] = KEY_MAX;
            }
        }
        uint32_t cwv[4], rwv[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            cwv[u] = 0xFFFFFFFFu; rwv[u] = 0;
            if (u < nl) {
                if (common) cwv[u] = __ldg(&common[lx[u] >> 5]) >> (lx[u] & 31);
                if (common2) cwv[u] &= __ldg(&common2[lx[u] >> 5]) >> (lx[u] & 31);
                if (repeat) {
                    const uint64_t ridx = fast_mod(s_key[li[u]], rm, rmprime);
                    rwv[u] = __ldg(&repeat[ridx >> 5]) >> (ridx & 31);
                }
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u)
            if (u < nl && (!(cwv[u] & 1u) || (rwv[u] & 1u))) s_key[li[u]] = KEY_MAX;
        __syncthreads();
    } else if (filtered) {
        for (uint32_t i0 = threadIdx.x; i0 < n_end; i0 += THREADS * 8) {
            uint64_t h[8];
            uint32_t cw[8], rw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t i = i0 + u * THREADS;
                h[u] = (i < n_end) ? s_key[i] : 0;
                uint64_t idx = fast_mod(h[u], m, mprime);
                cw[u] = 0xFFFFFFFFu; rw[u] = 0;
                if (i < n_end && i >= i_lo) {
                    if (common) cw[u] = __ldg(&common[idx >> 5]) >> (idx & 31);
                    if (common2) cw[u] &= __ldg(&common2[idx >> 5]) >> (idx & 31);
                    if (repeat) {
                        const uint64_t ridx = fast_mod(h[u], rm, rmprime);
                        rw[u] = __ldg(&repeat[ridx >> 5]) >> (ridx & 31);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t i = i0 + u * THREADS;
                if (i < n_end && i >= i_lo && (!(cw[u] & 1u) || (rw[u] & 1u))) s_key[i] = KEY_MAX;
            }
        }
        __syncthreads();
    }

    // ---- phase B: van Herk / Gil-Werman over blocks of w slots: [b*w, (b+1)*w)
    // suffix arg-min S[i] over [i, end of i's block]; rightmost wins ties
    {
        KeyIdx run; run.key = KEY_MAX; run.idx = 0xFFFFFFFFu;
        uint32_t closed = 0;
        uint32_t rem = hi % w;                       // (i + 1) % w for i = hi - 1, then counted down
        for (uint32_t i = hi; i-- > c0;) {          // local right-to-left pass
            KeyIdx cur; cur.key = s_key[i]; cur.idx = i;
            const bool last_of_block = rem == 0;
            rem = rem == 0 ? w - 1 : rem - 1;
            if (last_of_block) { closed = 1; run = cur; }             // i is the last slot of its block
            else if (run.idx == 0xFFFFFFFFu) run = cur;
            else if (cur.key < run.key) run = cur;                    // run lies to the right: wins ties
            s_S[i] = (uint16_t)run.idx;
        }
        // handed to the threads on the LEFT: arg-min over [c0, first block end in the chunk | chunk end]
        SegAgg own; own.key = KEY_MAX; own.idx = 0xFFFFFFFFu; own.flag = 0;
        if (c0 < hi) { own.idx = s_S[c0]; own.key = s_key[own.idx]; own.flag = SEG_VALID | (closed ? SEG_CLOSED : 0u); }
        SegAgg carry = cta_exclusive_carry<THREADS, false>(own, s_warp);
        // the carry (arg-min of what lies to the right, up to the block end) reaches the slots after the
        // last block end inside this chunk
        if (c0 < hi && (carry.flag & SEG_VALID)) {
            uint32_t rem2 = hi % w;
            for (uint32_t i = hi; i-- > c0;) {
                if (rem2 == 0) break;
                --rem2;
                if (carry.key <= s_key[s_S[i]]) s_S[i] = (uint16_t)carry.idx;   // carry is to the right: wins ties
            }
        }
    }
    // prefix arg-min P[i] over [start of i's block, i]; rightmost wins ties
    {
        KeyIdx run; run.key = KEY_MAX; run.idx = 0xFFFFFFFFu;
        uint32_t closed = 0;
        uint32_t rem = c0 % w;                       // i % w, counted up
        for (uint32_t i = c0; i < hi; ++i) {        // local left-to-right pass
            KeyIdx cur; cur.key = s_key[i]; cur.idx = i;
            const bool first_of_block = rem == 0;
            rem = rem + 1 == w ? 0 : rem + 1;
            if (first_of_block) { closed = 1; run = cur; }            // i is the first slot of its block
            else if (run.idx == 0xFFFFFFFFu) run = cur;
            else if (cur.key <= run.key) run = cur;                   // cur lies to the right: wins ties
            s_P[i] = (uint16_t)run.idx;
        }
        // handed to the threads on the RIGHT: arg-min over [last block start in the chunk | c0, chunk end]
        SegAgg own; own.key = KEY_MAX; own.idx = 0xFFFFFFFFu; own.flag = 0;
        if (c0 < hi) { own.idx = s_P[hi - 1]; own.key = s_key[own.idx]; own.flag = SEG_VALID | (closed ? SEG_CLOSED : 0u); }
        SegAgg carry = cta_exclusive_carry<THREADS, true>(own, s_warp);
        if (c0 < hi && (carry.flag & SEG_VALID)) {
            uint32_t rem2 = c0 % w;
            for (uint32_t i = c0; i < hi; ++i) {
                if (rem2 == 0) break;
                rem2 = rem2 + 1 == w ? 0 : rem2 + 1;
                if (carry.key < s_key[s_P[i]]) s_P[i] = (uint16_t)carry.idx;    // carry is to the left: loses ties
            }
        }
    }
    __syncthreads();

    // ---- phase C: window minima, change detection, ordered compaction
    // owned window ends: slots [w, n_end); thread t takes a contiguous share
    cnt = 0;
    int unresolved = 0;
    if (o0 < o1) {
        uint32_t prev = (o0 == w && !td.has_prev) ? 0xFFFFFFFFu : win_min(o0 - 1, o0_rem == 0);
        // the window before the tile's first owned one decides whether that one is "new": it must be resolved too
        if (o0 == w && td.has_prev && s_key[prev] == KEY_MAX) unresolved = 1;
        uint32_t r1 = o0_rem + 1 == w ? 0 : o0_rem + 1;          // (i + 1) % w
        for (uint32_t i = o0; i < o1; ++i) {
            uint32_t a = win_min(i, r1 == 0);
            r1 = r1 + 1 == w ? 0 : r1 + 1;
            if (s_key[a] == KEY_MAX) unresolved = 1;
            else if (a != prev) ++cnt;
            prev = a;
        }
    }
    if (tau_cur == KEY_MAX) break;                      // uniform: everything was queried
    if (!__syncthreads_or(unresolved)) break;           // every owned window has a verified minimum
    }   // pass
    // CTA exclusive scan of cnt
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_misc[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = lane < THREADS / 32 ? s_misc[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        if (lane < THREADS / 32) s_misc[lane] = v;
        if (lane == THREADS / 32 - 1) {
            // one allocation per tile in the unordered buffer
            unsigned long long base = atomicAdd(out.total, (unsigned long long)v);
            s_misc[THREADS / 32] = (base + v <= out.cap) ? (uint32_t)base : 0xFFFFFFFFu;
            out.tile_off[td.out_slot] = (uint32_t)base;
            out.tile_cnt[td.out_slot] = v;
        }
    }
    __syncthreads();
    const uint32_t tile_base = s_misc[THREADS / 32];
    if (tile_base == 0xFFFFFFFFu || cnt == 0) return;   // overflow: host re-runs with a larger buffer
    uint32_t off = tile_base + (incl - cnt) + (wid ? s_misc[wid - 1] : 0);
    {
        uint32_t prev = (o0 == w && !td.has_prev) ? 0xFFFFFFFFu : win_min(o0 - 1, o0_rem == 0);
        uint32_t r1 = o0_rem + 1 == w ? 0 : o0_rem + 1;
        for (uint32_t i = o0; i < o1; ++i) {
            uint32_t a = win_min(i, r1 == 0);
            r1 = r1 + 1 == w ? 0 : r1 + 1;
            if (a != prev && s_key[a] != KEY_MAX) {
                uint64_t b = valid_to_base(g, vbase + a);
                out.h1[off] = ext_hash(s_key[a], 1, g.k);
                out.pos[off] = (uint32_t)(b - td.cbase);
                out.contig[off] = td.contig;
                ++off;
            }
            prev = a;
        }
    }
}

// ------------------------------------------------------------------------------ filter pass rate (sampled)
// A few thousand evenly spaced k-mers of the view are hashed and looked up: the fraction that passes the filter
// decides how many candidates per window the sparse sketch kernel needs (its exactness never depends on it).
__global__ void sketch_sample_kernel(GenomeView g, const HashTables* __restrict__ g_tabs, const uint32_t* __restrict__ common,
                                     const uint32_t* __restrict__ common2, const uint32_t* __restrict__ repeat, uint64_t m, uint64_t mprime, uint64_t rm, uint64_t rmprime,
                                     uint64_t total_valid,
                                     uint64_t stride, unsigned int* __restrict__ counts /* [0] passed, [1] sampled */)
{
    __shared__ HashTables s_tabs;
    stage_tables(&s_tabs, g_tabs, g.k);
    __syncthreads();
    const uint64_t v = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * stride;
    if (v >= total_valid) return;
    hash_run(g, &s_tabs, v, 1, [&](uint32_t, uint64_t h0, uint64_t) {
        const uint64_t idx = fast_mod(h0, m, mprime);
        bool keep = true;
        if (common) keep = (__ldg(&common[idx >> 5]) >> (idx & 31)) & 1u;
        if (keep && common2) keep = (__ldg(&common2[idx >> 5]) >> (idx & 31)) & 1u;
        if (keep && repeat) {
            const uint64_t ridx = fast_mod(h0, rm, rmprime);
            keep = !((__ldg(&repeat[ridx >> 5]) >> (ridx & 31)) & 1u);
        }
        atomicAdd(&counts[1], 1u);
        if (keep) atomicAdd(&counts[0], 1u);
    });
}

// ------------------------------------------------------------------------------ (ii) sketch, sparse form
// The same selection as sketch_kernel, for tiles several times larger, without keeping every key in shared
// memory.  Only slots whose hash is below tau = tau_hi * 2^32 ("candidates", a few per cent) can be the minimum
// of a window that holds at least one candidate passing the filter, so:
//   A  hash every slot (kernel i); a thread stages its few candidates (key, slot) in its own column;
//   B  ordered compaction of the columns into one slot-sorted candidate list; Bloom query (iii-c) of the
//      candidates only, all of a thread's sector loads in flight together; failures become UINT64_MAX;
//   C  every thread slides over a contiguous share of the window ends, jumping between the only events that
//      can change the minimum (a candidate enters; the current minimum leaves -> rescan of ~tau*w survivors),
//      and marks a candidate the first time it becomes a window's rightmost minimum;
//   D  ordered compaction of the elected candidates into the unordered output buffer.
// Exactness needs every window of the tile (and the one before its first) to hold a survivor: gaps between
// consecutive survivors are checked, and a tile with an unresolved window (or a staging overflow) hands its
// `n_sub` dense sub-tiles to sketch_kernel through the escalation list instead of emitting anything.
// QALL (filters that pass < 10 % of the k-mers, where a hash threshold would have to let nearly everything through):
// EVERY slot is looked up -- a thread hashes its 16 slots into its staging column, then issues the sector loads of
// eight slots at a time -- and only the survivors are staged, so the candidate list is the complete list of survivors:
// a window without one emits nothing (its slots are all UINT64_MAX sentinels), no window has to be handed to the dense
// selector, and the window pass runs over a list ~1/pass-rate shorter than the slots.
template <int THREADS, int SCAP, int CCAP, bool QALL>
__global__ void __launch_bounds__(THREADS, 2)
sketch_sparse_kernel(GenomeView g, const HashTables* __restrict__ g_tabs, const uint32_t* __restrict__ common,
                     const uint32_t* __restrict__ summary /* QALL: one bit per 2^sum_shift bits of `common`, nullable */,
                     const uint32_t* __restrict__ common2, const uint32_t* __restrict__ repeat, uint64_t m, uint64_t mprime, uint64_t rm, uint64_t rmprime,
              const TileDesc* __restrict__ tiles,
                     uint32_t w, uint32_t NT, uint32_t C, uint32_t T_dense, uint32_t tau_hi, SketchOut out,
                     TileDesc* __restrict__ esc, unsigned int* __restrict__ esc_count, uint32_t sum_shift)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HashTables* s_tabs = reinterpret_cast<HashTables*>(smem_raw);
    uint64_t* s_skey = reinterpret_cast<uint64_t*>(s_tabs + 1);            // [SCAP][THREADS] staged keys, column per thread
    uint64_t* c_key = s_skey + (size_t)SCAP * THREADS;                     // [CCAP] candidate keys in slot order
    uint16_t* c_slot = reinterpret_cast<uint16_t*>(c_key + CCAP);          // [CCAP]
    uint8_t* s_sj = reinterpret_cast<uint8_t*>(c_slot + CCAP);             // [SCAP][THREADS] staged slot offsets
    __shared__ uint32_t s_wsum[THREADS / 32];
    __shared__ uint32_t s_bcast[2];

    const TileDesc td = tiles[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t vbase = td.vfirst - 1;
    const uint32_t i_lo = td.has_prev ? 0u : 1u;
    const uint64_t avail = td.vend - td.vfirst + 1;
    const uint32_t n_end = (uint32_t)min((uint64_t)NT, avail);
    __shared__ uint32_t s_isl[2];                          // islands of the tile's first and last slot
    if (tid == 0) s_isl[0] = find_island(g, vbase + i_lo);
    if (tid == 32) s_isl[1] = find_island(g, vbase + max(n_end, i_lo + 1u) - 1);
    stage_tables(s_tabs, g_tabs, g.k);
    __syncthreads();

    // CTA-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = sum
    auto cta_scan = [&](uint32_t val, uint32_t* total) -> uint32_t {
        uint32_t incl = val;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();                       // s_wsum may still be read from the previous scan
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) { const uint32_t x = s_wsum[i]; if (i < wid) woff += x; tot += x; }
        *total = tot;
        return woff + incl - val;
    };

    // ---- phase A: hash, stage the candidates
    const uint32_t c0 = tid * C;
    uint32_t n_st = 0;
    {
        const uint32_t a = max(c0, i_lo), b = min(c0 + C, n_end);
        if (a < b) {
            const uint32_t j0 = a - c0;
            if constexpr (QALL) {
                // C <= SCAP: every slot's hash fits the thread's column
                hash_run(g, s_tabs, vbase + a, b - a, [&](uint32_t j, uint64_t h0, uint64_t) { s_skey[j * THREADS + tid] = h0; },
                         s_isl[0], s_isl[1]);
                const uint32_t cnt = b - a;
#pragma unroll
                for (int half = 0; half < (SCAP + 7) / 8; ++half) {
                    uint32_t ok[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t jj = half * 8 + u;
                        ok[u] = 0;
                        if (jj < cnt) {
                            const uint64_t h = s_skey[jj * THREADS + tid];
                            const uint64_t idx = fast_mod(h, m, mprime);
                            // eight independent loads from the selective filter are in flight together; the second part of
                            // the common filter and the repeat filter are looked up only for the few that passed
                            if (summary) {              // L2-resident: does the k-mer's sector hold any bit at all?
                                const uint64_t sec = idx >> sum_shift;
                                ok[u] = (__ldg(&summary[sec >> 5]) >> (sec & 31)) & 1u;
                            } else {
                                ok[u] = common ? (__ldg(&common[idx >> 5]) >> (idx & 31)) & 1u : 1u;
                            }
                        }
                    }
                    if (summary) {
                        // the slots whose line holds a bit: their real lookups go out together as well (a warp nearly
                        // always has some, and one after the other they would each cost a trip to HBM)
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint32_t jj = half * 8 + u;
                            if (jj < cnt && (ok[u] & 1u)) {
                                const uint64_t idx = fast_mod(s_skey[jj * THREADS + tid], m, mprime);
                                ok[u] = (__ldg(&common[idx >> 5]) >> (idx & 31)) & 1u;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t jj = half * 8 + u;
                        if (jj < cnt && (ok[u] & 1u)) {                 // (n_st <= jj: the write never passes the read)
                            const uint64_t h = s_skey[jj * THREADS + tid];
                            if (common2) {
                                const uint64_t idx = fast_mod(h, m, mprime);
                                if (!((__ldg(&common2[idx >> 5]) >> (idx & 31)) & 1u)) continue;
                            }
                            if (repeat) {
                                const uint64_t ridx = fast_mod(h, rm, rmprime);
                                if ((__ldg(&repeat[ridx >> 5]) >> (ridx & 31)) & 1u) continue;
                            }
                            s_skey[n_st * THREADS + tid] = h;
                            s_sj[n_st * THREADS + tid] = (uint8_t)(j0 + jj);
                            ++n_st;
                        }
                    }
                }
            } else {
                hash_run(g, s_tabs, vbase + a, b - a, [&](uint32_t j, uint64_t h0, uint64_t) {
                    if ((uint32_t)(h0 >> 32) < tau_hi) {
                        if (n_st < SCAP) {
                            s_skey[n_st * THREADS + tid] = h0;
                            s_sj[n_st * THREADS + tid] = (uint8_t)(j0 + j);
                        }
                        ++n_st;
                    }
                }, s_isl[0], s_isl[1]);
            }
        }
    }
    int bad = n_st > SCAP;                     // staging overflow: escalate
    const uint32_t n_mine = min(n_st, (uint32_t)SCAP);
    uint32_t ncand = 0;
    const uint32_t off = cta_scan(n_mine, &ncand);
    if (ncand > CCAP) { bad = 1; ncand = 0; }
    else
        for (uint32_t i = 0; i < n_mine; ++i) {
            c_key[off + i] = s_skey[i * THREADS + tid];
            c_slot[off + i] = (uint16_t)(c0 + s_sj[i * THREADS + tid]);
        }
    __syncthreads();

    // ---- phase B: Bloom query of the candidates (kernel iii-c)
    if (!QALL && (common != nullptr || repeat != nullptr) && ncand) {
        constexpr int PER = (CCAP + THREADS - 1) / THREADS;
        uint32_t cw[PER], rw[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const uint32_t i = tid + u * THREADS;
            cw[u] = 0xFFFFFFFFu; rw[u] = 0;
            if (i < ncand) {
                const uint64_t idx = fast_mod(c_key[i], m, mprime);
                if (common) cw[u] = __ldg(&common[idx >> 5]) >> (idx & 31);
                if (common2) cw[u] &= __ldg(&common2[idx >> 5]) >> (idx & 31);
                if (repeat) {
                    const uint64_t ridx = fast_mod(c_key[i], rm, rmprime);
                    rw[u] = __ldg(&repeat[ridx >> 5]) >> (ridx & 31);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const uint32_t i = tid + u * THREADS;
            if (i < ncand && (!(cw[u] & 1u) || (rw[u] & 1u))) c_key[i] = KEY_MAX;
        }
        __syncthreads();
    }

    // ---- phase C1: is every window resolved?  (gaps between consecutive survivors, first and last window)
    const uint32_t q = (ncand + THREADS - 1) / THREADS;      // candidates per thread (contiguous share)
    const uint32_t i0 = min((uint32_t)tid * q, ncand), i1 = min(i0 + q, ncand);
    const int e_lo = td.has_prev ? (int)w - 1 : (int)w;      // end of the first window that must be resolved
    uint8_t* c_flag = s_sj;                                  // [CCAP] elected marks (the staging columns are dead)
    for (uint32_t i = i0; i < i1; ++i) {
        c_flag[i] = 0;
        if (c_key[i] == KEY_MAX) continue;
        int j = (int)i - 1;
        while (j >= 0 && c_key[j] == KEY_MAX) --j;
        const int prev_p = j >= 0 ? (int)c_slot[j] : e_lo - (int)w;
        if (!QALL && (int)c_slot[i] - prev_p > (int)w) bad = 1;      // (QALL: the list is complete, empty windows are real)
    }
    if (!QALL && tid == 0 && !bad) {
        int j = (int)ncand - 1;
        while (j >= 0 && c_key[j] == KEY_MAX) --j;
        const int last_p = j >= 0 ? (int)c_slot[j] : e_lo - (int)w;
        if ((int)n_end - last_p > (int)w) bad = 1;
    }
    if (__syncthreads_or(bad)) {
        // hand the tile's dense sub-tiles to sketch_kernel
        if (tid == 0) s_bcast[0] = atomicAdd(esc_count, td.n_sub);
        __syncthreads();
        const uint32_t base = s_bcast[0];
        if ((uint32_t)tid < td.n_sub) {
            TileDesc d;
            d.vfirst = td.vfirst + (uint64_t)tid * T_dense;
            d.vend = td.vend; d.cbase = td.cbase; d.contig = td.contig;
            d.has_prev = (td.has_prev || tid > 0) ? 1u : 0u;
            d.out_slot = td.out_slot + tid; d.n_sub = 1;
            esc[base + tid] = d;
            out.tile_off[td.out_slot + tid] = 0;
            out.tile_cnt[td.out_slot + tid] = 0;
        }
        return;
    }

    // ---- phase C2: every thread slides over its share of the window ends.  The window minimum can only change
    // when a candidate enters (its slot becomes the window end) or when the current minimum leaves, so the
    // thread jumps from event to event; a candidate is marked the first time it becomes the minimum.
    {
        const uint32_t n_own = n_end > w ? n_end - w : 0;
        const uint32_t per = (n_own + THREADS - 1) / THREADS;
        const uint32_t e0 = w + tid * per, e1 = min(e0 + per, n_end);
        // first candidate index with slot >= x
        auto lower = [&](uint32_t x) -> uint32_t {
            uint32_t lo = 0, hi = ncand;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (c_slot[mid] < x) lo = mid + 1; else hi = mid; }
            return lo;
        };
        // rightmost minimum among the survivors with slot in [a, b]; `from` = lower(a); -1 if none; *past = first index beyond b
        auto scan = [&](uint32_t from, uint32_t b, uint32_t* past) -> int {
            int best = -1;
            uint64_t bk = KEY_MAX;
            uint32_t i = from;
            for (; i < ncand && c_slot[i] <= b; ++i) {
                const uint64_t kx = c_key[i];
                if (kx != KEY_MAX && kx <= bk) { bk = kx; best = (int)i; }
            }
            *past = i;
            return best;
        };
        if (e0 < e1) {
            int cur;
            uint32_t nxt;                                    // next candidate (by index) that has not entered yet
            if (e0 > w || td.has_prev) {
                cur = scan(lower(e0 - w), e0 - 1, &nxt);     // the window before the first owned one: already emitted
            } else {
                cur = scan(lower(1), e0, &nxt);              // the contig's first window
                if (cur >= 0) c_flag[cur] = 1;
            }
            for (;;) {
                const uint32_t e_in = nxt < ncand ? (uint32_t)c_slot[nxt] : 0xFFFFFFFFu;
                const uint32_t e_out = cur >= 0 ? (uint32_t)c_slot[cur] + w : 0xFFFFFFFFu;
                const uint32_t e = min(e_in, e_out);
                if (e >= e1) break;
                int neu;
                if (e_out <= e_in) {
                    // the minimum left: every candidate before it has left too (smaller slot), so the new minimum is
                    // among the candidates after it that have entered -- no search for the window's first candidate
                    neu = scan((uint32_t)cur + 1u, e, &nxt);
                } else {
                    const uint64_t kx = c_key[nxt];
                    neu = (kx != KEY_MAX && (cur < 0 || kx <= c_key[cur])) ? (int)nxt : cur;
                    ++nxt;
                }
                if (neu != cur && neu >= 0) c_flag[neu] = 1;
                cur = neu;
            }
        }
    }
    __syncthreads();

    // ---- phase D: ordered compaction of the elected candidates
    uint32_t n_emit = 0;
    for (uint32_t i = i0; i < i1; ++i) n_emit += c_flag[i];
    uint32_t total = 0;
    const uint32_t eoff = cta_scan(n_emit, &total);
    if (tid == 0) {
        unsigned long long base = atomicAdd(out.total, (unsigned long long)total);
        s_bcast[1] = (base + total <= out.cap) ? (uint32_t)base : 0xFFFFFFFFu;
        out.tile_off[td.out_slot] = (uint32_t)base;
        out.tile_cnt[td.out_slot] = total;
    }
    if (tid >= 1 && (uint32_t)tid < td.n_sub) { out.tile_off[td.out_slot + tid] = 0; out.tile_cnt[td.out_slot + tid] = 0; }
    __syncthreads();
    const uint32_t tile_base = s_bcast[1];
    if (tile_base == 0xFFFFFFFFu || n_emit == 0) return;    // overflow: host re-runs with a larger buffer
    uint32_t o = tile_base + eoff;
    for (uint32_t i = i0; i < i1; ++i) {
        if (!c_flag[i]) continue;
        const uint64_t b = valid_to_base(g, vbase + c_slot[i]);
        out.h1[o] = ext_hash(c_key[i], 1, g.k);
        out.pos[o] = (uint32_t)(b - td.cbase);
        out.contig[o] = td.contig;
        ++o;
    }
}

// stitch the per-tile runs of the unordered buffer into tile order
__global__ void sketch_gather_kernel(const uint64_t* __restrict__ h1_in, const uint32_t* __restrict__ pos_in,
                                     const uint32_t* __restrict__ ctg_in, const uint32_t* __restrict__ tile_off,
                                     const uint32_t* __restrict__ tile_cnt, const uint64_t* __restrict__ tile_dst,
                                     uint32_t n_tiles, uint64_t* __restrict__ h1_out, uint32_t* __restrict__ pos_out,
                                     uint32_t* __restrict__ ctg_out)
{
    // one warp per tile
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    uint32_t lane = threadIdx.x & 31;
    uint32_t src = tile_off[t], n = tile_cnt[t];
    uint64_t dst = tile_dst[t];
    for (uint32_t i = lane; i < n; i += 32) {
        h1_out[dst + i] = h1_in[src + i];
        pos_out[dst + i] = pos_in[src + i];
        ctg_out[dst + i] = ctg_in[src + i];
    }
}

// exclusive scan of tile counts (single CTA; n_tiles is a few hundred thousand at most)
__global__ void tile_scan_kernel(const uint32_t* __restrict__ cnt, uint64_t* __restrict__ dst, uint32_t n)
{
    __shared__ uint64_t s_part[1024];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t a = threadIdx.x * per, b = min(a + per, n);
    uint64_t sum = 0;
    for (uint32_t i = a; i < b; ++i) sum += cnt[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) { uint64_t v = s_part[i]; s_part[i] = run; run += v; }
    }
    __syncthreads();
    uint64_t run = s_part[threadIdx.x];
    for (uint32_t i = a; i < b; ++i) { dst[i] = run; run += cnt[i]; }
}

}  // namespace nts
