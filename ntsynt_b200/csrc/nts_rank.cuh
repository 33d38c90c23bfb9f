// nts_rank.cuh -- (iii-a) pass 1 of the partitioned Bloom insert WITHOUT shared-memory atomics.
//
// bf_bin_kernel (nts_bin.cuh) ranks the items of a tile with one shared-memory atomicAdd per k-mer; ATOMS with
// spread addresses costs 2 LSU cycles per lane on this part, which alone is 20.6 ms per 3 Gbp genome and keeps the
// LSU busy so that nothing else can run beside it.  bf_rank_bin_kernel produces the same output (the tile's items
// appended to their region's bucket, one global atomicAdd per non-empty (tile, bucket) run) with plain LDS / STS:
//
//   the bucket id b = b_hi * 32 + b_lo is sorted in two stable counting-sort passes (b_lo, then b_hi).  In a pass
//   every thread owns ITEMS consecutive items and counts them in PRIVATE byte counters (one 32-bit word holds the
//   counters of four digits; word (digit >> 2) * THREADS + tid, so a warp never conflicts).  The counters are then
//   scanned digit-major: a thread scans the eight threads of one segment with packed-byte adds (8 x 16 <= 128 fits a
//   byte), segment totals are scanned across the 64 segments of a word column with warp shuffles on packed 16-bit
//   fields, a 32-lane scan gives the digit bases.  An item's position is
//        segoff[digit][tid >> 3]  +  (threads before me in my segment)[digit]  +  (my earlier items with this digit).
//
// The second pass needs no copy of b_lo: an item's b_lo is the digit whose range of pass-one positions holds it.
// Items are 32-bit words {b_hi : HB, bit index inside the region : 32 - HB}; the host picks HB = 4 (<= 512 buckets,
// regions up to 2^28 bits: the production plan) or HB = 5 (<= 1024 buckets, regions up to 2^27 bits).
#pragma once
#include "nts_bin.cuh"

namespace nts {

// exclusive digit-major scan of the private counters; see the header comment.  ND digits (multiple of 4).
// CNT[(ND/4) * THREADS] packed byte counters (in: counts, out: exclusive prefix inside the 8-thread segment),
// SEG[ND * 64] (out) position of the first item of (digit, segment), s_base[ND + 1] (out) digit bases.
template <int THREADS, int ND>
__device__ __forceinline__ void rank_scan(uint32_t* __restrict__ CNT, uint16_t* __restrict__ SEG, uint32_t* __restrict__ s_base,
                                          uint32_t* __restrict__ s_wtot)
{
    static_assert(THREADS == 512 && ND % 4 == 0 && ND <= 32, "layout assumes 512 threads: 64 segments of 8 threads");
    constexpr int NQ = ND / 4;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int q = t >> 6, seg = t & 63;
    uint32_t lo = 0, hi = 0;
    if (q < NQ) {
        uint4* p = reinterpret_cast<uint4*>(CNT + q * THREADS + 8 * seg);
        uint4 a = p[0], b = p[1];
        uint32_t run = 0, tmp;
        tmp = a.x; a.x = run; run += tmp;
        tmp = a.y; a.y = run; run += tmp;
        tmp = a.z; a.z = run; run += tmp;
        tmp = a.w; a.w = run; run += tmp;
        tmp = b.x; b.x = run; run += tmp;
        tmp = b.y; b.y = run; run += tmp;
        tmp = b.z; b.z = run; run += tmp;
        tmp = b.w; b.w = run; run += tmp;
        p[0] = a; p[1] = b;
        lo = (run & 0xFFu) | ((run & 0xFF00u) << 8);
        hi = ((run >> 16) & 0xFFu) | ((run >> 24) << 16);
    }
    uint32_t ilo = lo, ihi = hi;                  // inclusive over the warp's 32 segments, four 16-bit fields
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, ilo, d), y = __shfl_up_sync(0xffffffffu, ihi, d);
        if (lane >= d) { ilo += x; ihi += y; }
    }
    if (lane == 31) { s_wtot[wid * 2] = ilo; s_wtot[wid * 2 + 1] = ihi; }
    __syncthreads();
    uint32_t plo = 0, phi = 0;                    // a word column is two warps: the second starts after the first
    if (wid & 1) { plo = s_wtot[(wid - 1) * 2]; phi = s_wtot[(wid - 1) * 2 + 1]; }
    if (wid == 0) {
        uint32_t tot = 0;
        if (lane < ND) {
            const int qq = lane >> 2, f = lane & 3;
            const uint32_t a = s_wtot[(2 * qq) * 2 + (f >> 1)] + s_wtot[(2 * qq + 1) * 2 + (f >> 1)];
            tot = (f & 1) ? (a >> 16) : (a & 0xFFFFu);
        }
        uint32_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
        if (lane < ND) s_base[lane] = inc - tot;
        if (lane == ND - 1) s_base[ND] = inc;
    }
    __syncthreads();
    if (q < NQ) {
        const uint32_t elo = ilo - lo + plo, ehi = ihi - hi + phi;
        SEG[(4 * q + 0) * 64 + seg] = (uint16_t)(s_base[4 * q + 0] + (elo & 0xFFFFu));
        SEG[(4 * q + 1) * 64 + seg] = (uint16_t)(s_base[4 * q + 1] + (elo >> 16));
        SEG[(4 * q + 2) * 64 + seg] = (uint16_t)(s_base[4 * q + 2] + (ehi & 0xFFFFu));
        SEG[(4 * q + 3) * 64 + seg] = (uint16_t)(s_base[4 * q + 3] + (ehi >> 16));
    }
    __syncthreads();
}

template <int THREADS, int ITEMS>
struct RankSmem {
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr size_t W0 = sizeof(HashTables);            // u32[TILE]  hash-pass words; later the sorted items
    static constexpr size_t W1 = W0 + (size_t)TILE * 4;         // u32[TILE]  items after pass one
    static constexpr size_t S0 = W1 + (size_t)TILE * 4;         // u16[TILE]  b_lo | rank << 5; later the sorted bucket ids
    static constexpr size_t CNT = S0 + (size_t)TILE * 2;        // u32[8 * THREADS]
    static constexpr size_t SEG = CNT + (size_t)8 * THREADS * 4;  // u16[32 * 64]
    static constexpr size_t BASE_A = SEG + 32 * 64 * 2;         // u32[33]
    static constexpr size_t BASE_B = BASE_A + 36 * 4;           // u32[33]
    static constexpr size_t WTOT = BASE_B + 36 * 4;             // u32[THREADS / 32 * 2]
    static constexpr size_t START = WTOT + (size_t)(THREADS / 32) * 2 * 4;   // u32[P]  first sorted position of a bucket's run
    static size_t bytes(uint32_t P) { return START + (size_t)P * 12; }      // + END / DST [P], FIT [P]
};

template <int THREADS, int ITEMS, int HB>
__global__ void __launch_bounds__(THREADS, 2) bf_rank_bin_kernel(GenomeView g, const HashTables* __restrict__ g_tabs,
                                                                  uint32_t* __restrict__ bits, uint64_t m, uint64_t mprime,
                                                                  uint64_t total_valid, BinParams bp, uint32_t block0)
{
    static_assert(ITEMS == 16, "the pass-two rank word holds sixteen 4-bit ranks");
    using L = RankSmem<THREADS, ITEMS>;
    constexpr int TILE = L::TILE;
    constexpr int LOW_BITS = 32 - HB;
    constexpr uint32_t LOW_MASK = (1u << LOW_BITS) - 1u;
    constexpr int NDB = 1 << HB;
    extern __shared__ __align__(16) unsigned char smem_rank[];
    HashTables* s_tabs = reinterpret_cast<HashTables*>(smem_rank);
    uint32_t* W0 = reinterpret_cast<uint32_t*>(smem_rank + L::W0);
    uint32_t* W1 = reinterpret_cast<uint32_t*>(smem_rank + L::W1);
    uint16_t* S0 = reinterpret_cast<uint16_t*>(smem_rank + L::S0);
    uint32_t* CNT = reinterpret_cast<uint32_t*>(smem_rank + L::CNT);
    uint16_t* SEG = reinterpret_cast<uint16_t*>(smem_rank + L::SEG);
    uint32_t* s_baseA = reinterpret_cast<uint32_t*>(smem_rank + L::BASE_A);
    uint32_t* s_baseB = reinterpret_cast<uint32_t*>(smem_rank + L::BASE_B);
    uint32_t* s_wtot = reinterpret_cast<uint32_t*>(smem_rank + L::WTOT);
    uint32_t* s_start = reinterpret_cast<uint32_t*>(smem_rank + L::START);
    uint32_t* s_dst = s_start + bp.n_buckets;                    // first the end of the run, then its global item index
    uint32_t* s_fit = s_dst + bp.n_buckets;
    const int tid = threadIdx.x;
    __shared__ uint32_t s_isl[2];                                  // islands of the tile's first and last k-mer
    {
        const uint64_t t0 = (uint64_t)(blockIdx.x + block0) * TILE;
        const uint64_t nt = min((uint64_t)TILE, total_valid - t0);
        if (tid == 0) s_isl[0] = find_island(g, t0);
        if (tid == 32) s_isl[1] = find_island(g, t0 + nt - 1);
    }
    stage_tables_bin(s_tabs, g_tabs, g.k);
#pragma unroll
    for (int qd = 0; qd < 8; ++qd) CNT[qd * THREADS + tid] = 0;
    for (uint32_t b = tid; b < bp.n_buckets; b += THREADS) { s_start[b] = 0; s_dst[b] = 0; }
    __syncthreads();
    const uint64_t tile0 = (uint64_t)(blockIdx.x + block0) * TILE;     // block0: first tile of this launch (staged inserts)
    const uint32_t n_tile = (uint32_t)min((uint64_t)TILE, total_valid - tile0);
    const uint64_t v0 = tile0 + (uint64_t)tid * ITEMS;
    const uint32_t n_mine = v0 < total_valid ? (uint32_t)min((uint64_t)ITEMS, total_valid - v0) : 0;
    const uint32_t rmask = bp.region_shift >= 32 ? 0xFFFFFFFFu : ((1u << bp.region_shift) - 1u);

    // ---- hash, bucket, private count by b_lo
    if (n_mine)
        hash_run(g, s_tabs, v0, n_mine, [&](uint32_t j, uint64_t h0, uint64_t) {
            const uint64_t idx = fast_mod(h0, m, mprime);
            const uint32_t b = (uint32_t)(idx >> bp.region_shift);
            const uint32_t d = b & 31u, sh = (b & 3u) * 8u;
            uint32_t* c = CNT + (d >> 2) * THREADS + tid;
            const uint32_t w = *c;
            *c = w + (1u << sh);
            const uint32_t slot = j * THREADS + tid;
            W0[slot] = ((uint32_t)idx & rmask) | ((b >> 5) << LOW_BITS);
            S0[slot] = (uint16_t)(d | (((w >> sh) & 0xFFu) << 5));
        }, s_isl[0], s_isl[1]);
    __syncthreads();
    rank_scan<THREADS, 32>(CNT, SEG, s_baseA, s_wtot);
    // ---- scatter one: stable by b_lo
    for (uint32_t j = 0; j < n_mine; ++j) {
        const uint32_t slot = j * THREADS + tid;
        const uint32_t s = S0[slot], d = s & 31u;
        const uint32_t before = (CNT[(d >> 2) * THREADS + tid] >> ((d & 3u) * 8u)) & 0xFFu;
        W1[SEG[d * 64 + (tid >> 3)] + before + (s >> 5)] = W0[slot];
    }
#pragma unroll
    for (int qd = 0; qd < NDB / 4; ++qd) CNT[qd * THREADS + tid] = 0;
    __syncthreads();
    // ---- pass two: my ITEMS consecutive items of the b_lo order, private count by b_hi
    const uint32_t p0 = (uint32_t)tid * ITEMS;
    const uint32_t n_two = p0 < n_tile ? min((uint32_t)ITEMS, n_tile - p0) : 0;
    uint32_t w2[ITEMS];
    uint64_t ranks = 0;                                          // 4-bit rank among my earlier items with the same b_hi
    {
        const uint4* src = reinterpret_cast<const uint4*>(W1 + p0);
#pragma unroll
        for (int u = 0; u < ITEMS / 4; ++u) {
            const uint4 v = src[u];
            w2[4 * u] = v.x; w2[4 * u + 1] = v.y; w2[4 * u + 2] = v.z; w2[4 * u + 3] = v.w;
        }
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        if ((uint32_t)j < n_two) {
            const uint32_t d = w2[j] >> LOW_BITS, sh = (d & 3u) * 8u;
            uint32_t* c = CNT + (d >> 2) * THREADS + tid;
            const uint32_t w = *c;
            *c = w + (1u << sh);
            ranks |= (uint64_t)((w >> sh) & 0xFu) << (4 * j);       // (a count of 16 would need j = 16)
        }
    __syncthreads();
    rank_scan<THREADS, NDB>(CNT, SEG, s_baseB, s_wtot);
    // ---- scatter two: items in bucket order; b_lo of a pass-one position = the digit whose range holds it
    {
        uint32_t dlo = 0;
        if (n_two) {
            uint32_t lo = 0, hi = 32;                            // largest d with baseA[d] <= p0 (empty digits share a base)
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_baseA[mid] <= p0) lo = mid; else hi = mid; }
            dlo = lo;
        }
        uint32_t next = s_baseA[dlo + 1];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if ((uint32_t)j < n_two) {
                const uint32_t p = p0 + j;
                while (p >= next) { ++dlo; next = s_baseA[dlo + 1]; }
                const uint32_t d = w2[j] >> LOW_BITS;
                const uint32_t before = (CNT[(d >> 2) * THREADS + tid] >> ((d & 3u) * 8u)) & 0xFFu;
                const uint32_t pos = SEG[d * 64 + (tid >> 3)] + before + (uint32_t)((ranks >> (4 * j)) & 0xFu);
                W0[pos] = w2[j] & LOW_MASK;
                S0[pos] = (uint16_t)(d * 32u + dlo);
            }
    }
    __syncthreads();
    // ---- runs: first / one-past-last sorted position of every bucket present in the tile
    for (uint32_t i = tid; i < n_tile; i += THREADS) {
        const uint32_t b = S0[i];
        if (i == 0 || S0[i - 1] != b) s_start[b] = i;
        if (i + 1 == n_tile || S0[i + 1] != b) s_dst[b] = i + 1;
    }
    __syncthreads();
    for (uint32_t b = tid; b < bp.n_buckets; b += THREADS) {     // reserve the runs (one global atomic per non-empty bucket)
        const uint32_t c = s_dst[b] - s_start[b];
        uint32_t base = 0, fit = 0;
        if (c) {
            base = atomicAdd(&bp.cursor[b], c);
            const uint32_t cap = bp.bucket_cap[b];
            fit = base >= cap ? 0u : min(c, cap - base);
        }
        s_dst[b] = (uint32_t)bp.bucket_off[b] + base;
        s_fit[b] = fit;
    }
    __syncthreads();
    // ---- write out, one thread per item: a warp writes the tails / heads of two or three runs as contiguous pieces
    for (uint32_t i = tid; i < n_tile; i += THREADS) {
        const uint32_t b = S0[i];
        const uint32_t r = i - s_start[b];
        const uint32_t x = W0[i];
        if (r < s_fit[b]) {
            bp.items[(uint64_t)s_dst[b] + r] = x;
        } else if (bp.ovf_count) {
            atomicAdd(bp.ovf_count, 1u);
        } else {                                                   // overflow (heavy hitters): apply directly
            const uint64_t idx = ((uint64_t)b << bp.region_shift) + x;
            atomicOr(&bits[idx >> 5], 1u << (idx & 31));
        }
    }
}

}  // namespace nts
