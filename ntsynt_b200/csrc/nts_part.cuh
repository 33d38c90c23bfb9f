// nts_part.cuh -- (iii-a) partitioned Bloom insert without per-k-mer atomics.
//
// One Bloom insert is one random bit set per k-mer into a 14.8 GB array.  Every *global* scattered access costs the
// SM >= 1 LSU cycle per lane (REDG 1.3-2.0, ATOMS 2.0 measured in round 1), a scattered *plain* shared-memory access
// ~0.1.  So the bit indices are radix-partitioned through shared memory, in two levels, down to regions small enough
// to be finished in shared memory with plain byte stores:
//
//   pass 1  bf_part1_kernel : hash a tile of k-mers (kernel i), idx = h0 mod m, split into
//                             (level-1 bucket b1, offset x1 inside it); rank the tile by b1 with warp ballots
//                             (no atomics), stage it in bucket order in shared memory, append every bucket's run
//                             to its global bucket (one global atomicAdd per (tile, non-empty bucket));
//   pass 2  bf_part2_kernel : the same partition step over the items of one level-1 bucket, by final region b2;
//   pass 3  bf_apply3_kernel: a CTA owns one final region of R bits at a time: byte flags in shared memory are set
//                             with plain stores, packed 128 flags -> 128 bits, and written with coalesced 128-bit
//                             stores as SET (out = bits), AND (out = prev & bits; the cascade of
//                             src/ntsynt_make_common_bf.cpp:136-160 without a separate level filter pass) or
//                             OR (out |= bits).  The flags a thread set are cleared again by the same thread, so the
//                             array is never re-zeroed.
//
// Items that do not fit a bucket (heavy-hitter k-mers) go to an overflow list of full bit indices that
// bf_overflow_kernel applies with plain atomics afterwards: correctness never depends on the capacities.
#pragma once
#include "nts_device.cuh"

namespace nts {

struct PartParams {
    uint64_t R1;            // bits per level-1 bucket = R * P2
    uint32_t R;             // bits per final region (multiple of 128)
    uint32_t magic, shift;  // floor(x / (R >> 7)) = (x * magic) >> shift  (+ correction), x = bit offset >> 7
    uint32_t P1, P2;        // level-1 buckets; regions per level-1 bucket (power of two)
    uint32_t log2P2;
    uint32_t n_regions;     // ceil(m / R)
    uint32_t cap1, cap2;    // items a level-1 bucket / a region bucket holds (multiples of 4)
    uint32_t* items1;       // [P1 * cap1]
    uint32_t* items2;       // [n_regions_padded * cap2]
    unsigned int* cursor1;  // [P1]
    unsigned int* cursor2;  // [P1 * P2]
    uint64_t* ovf;          // overflow list of full bit indices
    unsigned long long* ovf_count;   // [0] appended so far (may run past ovf_cap), [1] error flag
    uint64_t ovf_cap;
};

__device__ __forceinline__ void stage_tables_part(HashTables* s_tabs, const HashTables* __restrict__ g_tabs, uint32_t k)
{
    for (uint32_t i = threadIdx.x; i < 16; i += blockDim.x) s_tabs->roll[i] = g_tabs->roll[i];
    for (uint32_t i = threadIdx.x; i < k * 4; i += blockDim.x) {
        s_tabs->init_f[i] = g_tabs->init_f[i];
        s_tabs->init_r[i] = g_tabs->init_r[i];
    }
}

// exact floor(r / R) for R a multiple of 128 (r / R == (r >> 7) / (R >> 7)); the magic over-estimates by at most one
__device__ __forceinline__ uint32_t div_region(uint64_t r, const PartParams& pp)
{
    const uint32_t x = (uint32_t)(r >> 7);
    uint32_t q = (uint32_t)(((uint64_t)x * pp.magic) >> pp.shift);
    if ((uint64_t)q * pp.R > r) --q;
    return q;
}

// append to the overflow list; called by whole warps (`mine` = this lane has an item): one atomicAdd per warp
__device__ __forceinline__ void ovf_push(const PartParams& pp, bool mine, uint64_t idx)
{
    const uint32_t mk = __ballot_sync(0xffffffffu, mine);
    if (!mk) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(pp.ovf_count, (unsigned long long)__popc(mk));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
        const unsigned long long at = base + __popc(mk & ((1u << lane) - 1u));
        if (at < pp.ovf_cap) pp.ovf[at] = idx; else pp.ovf_count[1] = 1ull;
    }
}

// lanes holding the same key (valid lanes only); 11 ballots instead of one shared-memory atomic per item
template <bool USE_MATCH>
__device__ __forceinline__ uint32_t match_key(uint32_t key, bool valid)
{
    if (USE_MATCH) return __match_any_sync(0xffffffffu, valid ? key : (0x10000u + (threadIdx.x & 31)));
    uint32_t mk = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 10; ++bit) {
        const bool p = (key >> bit) & 1u;
        const uint32_t v = __ballot_sync(0xffffffffu, p);
        mk &= p ? v : ~v;
    }
    return mk;
}

// Shared-memory layout of one partition step (THREADS threads, ITEMS items each, up to PMAX buckets)
template <int THREADS, int ITEMS, int PMAX>
struct PartSmem {
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int NW = THREADS / 32;
    // region A: phase H/R reads (x: u32[TILE], b: u16[TILE]); phase P/W stages uint2[TILE] in bucket order
    static constexpr size_t A_BYTES = (size_t)TILE * 8;
    static constexpr size_t WC_BYTES = (size_t)NW * PMAX * 2;          // per-warp bucket counters (u16)
    static constexpr size_t OG_BYTES = (size_t)(PMAX + 1) * 8;         // (tile offset, global offset) per bucket
    static constexpr size_t FIT_BYTES = (size_t)PMAX * 4;
    static constexpr size_t BYTES = A_BYTES + WC_BYTES + OG_BYTES + FIT_BYTES + 64;
};

// One partition step over the tile whose (x, b) pairs sit in region A (slot = j * THREADS + tid; b == 0xFFFF marks
// an empty slot).  Buckets 0..P-1; bucket b's global storage starts at out + (size_t)b * cap and holds cap items;
// cursor[b] counts what was appended.  ovf_base(b, x) gives the full bit index of an item for the overflow list.
template <int THREADS, int ITEMS, int PMAX, bool USE_MATCH, typename OvfIdx>
__device__ __forceinline__ void partition_tile(unsigned char* smem_a, uint32_t n_tile, uint32_t P, unsigned int* __restrict__ cursor,
                                               uint32_t cap, uint32_t* __restrict__ out, const PartParams& pp, OvfIdx&& ovf_idx)
{
    using L = PartSmem<THREADS, ITEMS, PMAX>;
    constexpr int TILE = L::TILE, NW = L::NW;
    uint32_t* s_x = reinterpret_cast<uint32_t*>(smem_a);
    uint16_t* s_b = reinterpret_cast<uint16_t*>(s_x + TILE);
    uint2* s_stage = reinterpret_cast<uint2*>(smem_a);
    uint16_t* s_wc = reinterpret_cast<uint16_t*>(smem_a + L::A_BYTES);
    uint2* s_og = reinterpret_cast<uint2*>(smem_a + L::A_BYTES + L::WC_BYTES);
    uint32_t* s_fit = reinterpret_cast<uint32_t*>(smem_a + L::A_BYTES + L::WC_BYTES + L::OG_BYTES);
    uint32_t* s_wsum = s_fit + PMAX;                                   // [NW] (inside the 64 spare bytes)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint16_t* wcw = s_wc + (size_t)wid * PMAX;

    // ---- R: rank inside (warp, bucket) with ballots; the warp's counters live in its own row of s_wc
    uint32_t x[ITEMS], br[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t slot = j * THREADS + tid;
        const uint32_t b = s_b[slot];
        x[j] = s_x[slot];
        const bool valid = b != 0xFFFFu;
        const uint32_t mk = match_key<USE_MATCH>(b, valid);
        const uint32_t rank = __popc(mk & lt_mask);
        uint32_t prev = 0;
        if (valid) prev = wcw[b];
        __syncwarp();
        if (valid && rank == 0) wcw[b] = (uint16_t)(prev + __popc(mk));
        __syncwarp();
        br[j] = valid ? (b | ((prev + rank) << 16)) : 0xFFFFu;
    }
    __syncthreads();
    // ---- S: per bucket, exclusive prefix over the warps; reserve the run in the global bucket; tile offsets
    {
        uint32_t carry = 0;
        for (uint32_t b0 = 0; b0 < P; b0 += THREADS) {
            const uint32_t b = b0 + tid;
            uint32_t c = 0;
            if (b < P) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint32_t t = s_wc[(size_t)w * PMAX + b];
                    s_wc[(size_t)w * PMAX + b] = (uint16_t)c;
                    c += t;
                }
                uint32_t base = 0, fit = 0;
                if (c) {
                    base = atomicAdd(&cursor[b], c);
                    fit = base >= cap ? 0u : min(c, cap - base);
                }
                s_fit[b] = fit;
                s_og[b].y = b * cap + base;                 // < 2^32 (host-checked)
            }
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            if (lane == 31) s_wsum[wid] = incl;
            __syncthreads();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) { const uint32_t t = s_wsum[i]; if (i < wid) woff += t; total += t; }
            if (b < P) s_og[b].x = carry + woff + incl - c;
            carry += total;
            __syncthreads();
        }
    }
    // ---- P: place (x, b) at its position in bucket order (region A is dead as (s_x, s_b) by now)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t b = br[j] & 0xFFFFu;
        if (b != 0xFFFFu) {
            const uint32_t pos = s_og[b].x + wcw[b] + (br[j] >> 16);
            s_stage[pos] = make_uint2(x[j], b);
        }
    }
    __syncthreads();
    // ---- W: one thread per staged item; consecutive threads write consecutive addresses of a run
    for (uint32_t i0 = 0; i0 < n_tile; i0 += THREADS) {
        const uint32_t i = i0 + tid;
        bool over = false;
        uint64_t oidx = 0;
        if (i < n_tile) {
            const uint2 it = s_stage[i];
            const uint2 og = s_og[it.y];
            const uint32_t rel = i - og.x;
            if (rel < s_fit[it.y]) out[og.y + rel] = it.x;
            else { over = true; oidx = ovf_idx(it.y, it.x); }
        }
        ovf_push(pp, over, oidx);
    }
}

template <int THREADS, int ITEMS, int PMAX>
__device__ __forceinline__ void zero_counters(unsigned char* smem_a)
{
    using L = PartSmem<THREADS, ITEMS, PMAX>;
    uint32_t* p = reinterpret_cast<uint32_t*>(smem_a + L::A_BYTES);
    for (uint32_t i = threadIdx.x; i < L::WC_BYTES / 4; i += THREADS) p[i] = 0;
}

// ---------------------------------------------------------------------------------------------- pass 1
template <int THREADS, int ITEMS, int PMAX, bool USE_MATCH>
__global__ void __launch_bounds__(THREADS, 2)
bf_part1_kernel(GenomeView g, const HashTables* __restrict__ g_tabs, uint64_t m, uint64_t mprime, uint64_t total_valid, PartParams pp)
{
    using L = PartSmem<THREADS, ITEMS, PMAX>;
    constexpr int TILE = L::TILE;
    extern __shared__ __align__(16) unsigned char smem_p[];
    HashTables* s_tabs = reinterpret_cast<HashTables*>(smem_p);
    unsigned char* smem_a = smem_p + sizeof(HashTables);
    uint32_t* s_x = reinterpret_cast<uint32_t*>(smem_a);
    uint16_t* s_b = reinterpret_cast<uint16_t*>(s_x + TILE);
    stage_tables_part(s_tabs, g_tabs, g.k);
    zero_counters<THREADS, ITEMS, PMAX>(smem_a);
    const uint64_t tile0 = (uint64_t)blockIdx.x * TILE;
    const uint32_t n_tile = (uint32_t)min((uint64_t)TILE, total_valid - tile0);
    const uint64_t v0 = tile0 + (uint64_t)threadIdx.x * ITEMS;
    const uint32_t n_mine = v0 < total_valid ? (uint32_t)min((uint64_t)ITEMS, total_valid - v0) : 0;
    for (uint32_t j = n_mine; j < ITEMS; ++j) s_b[j * THREADS + threadIdx.x] = 0xFFFFu;
    __syncthreads();
    // ---- H: hash (kernel i), bit index, level-1 bucket
    if (n_mine)
        hash_run(g, s_tabs, v0, n_mine, [&](uint32_t j, uint64_t h0, uint64_t) {
            const uint64_t idx = fast_mod(h0, m, mprime);
            const uint32_t b1 = div_region(idx, pp) >> pp.log2P2;
            const uint32_t slot = j * THREADS + threadIdx.x;
            s_x[slot] = (uint32_t)idx - (uint32_t)((uint64_t)b1 * pp.R1);      // exact: the difference is < R1 < 2^32
            s_b[slot] = (uint16_t)b1;
        });
    __syncthreads();
    const uint64_t R1 = pp.R1;
    partition_tile<THREADS, ITEMS, PMAX, USE_MATCH>(smem_a, n_tile, pp.P1, pp.cursor1, pp.cap1, pp.items1, pp,
                                                    [R1](uint32_t b, uint32_t x) { return (uint64_t)b * R1 + x; });
}

// ---------------------------------------------------------------------------------------------- pass 2
// grid = P1 * chunks_per_bucket; CTA (b1, c) takes items [c * TILE, (c + 1) * TILE) of level-1 bucket b1
template <int THREADS, int ITEMS, int PMAX, bool USE_MATCH>
__global__ void __launch_bounds__(THREADS, 2) bf_part2_kernel(PartParams pp, uint32_t chunks_per_bucket)
{
    using L = PartSmem<THREADS, ITEMS, PMAX>;
    constexpr int TILE = L::TILE;
    static_assert(ITEMS % 4 == 0, "items are loaded as uint4");
    extern __shared__ __align__(16) unsigned char smem_p[];
    unsigned char* smem_a = smem_p;
    uint32_t* s_x = reinterpret_cast<uint32_t*>(smem_a);
    uint16_t* s_b = reinterpret_cast<uint16_t*>(s_x + TILE);
    const uint32_t b1 = blockIdx.x / chunks_per_bucket, c = blockIdx.x % chunks_per_bucket;
    const uint32_t n1 = min(pp.cursor1[b1], pp.cap1);
    const uint32_t start = c * TILE;
    if (start >= n1) return;
    const uint32_t n_tile = min((uint32_t)TILE, n1 - start);
    zero_counters<THREADS, ITEMS, PMAX>(smem_a);
    const uint32_t* __restrict__ src = pp.items1 + (size_t)b1 * pp.cap1 + start;     // 16-byte aligned (cap1, TILE % 4 == 0)
    uint4 v[ITEMS / 4];
#pragma unroll
    for (int u = 0; u < ITEMS / 4; ++u) {
        const uint32_t i4 = (u * THREADS + threadIdx.x) * 4;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i4 + 3 < n_tile) v[u] = __ldg(reinterpret_cast<const uint4*>(src + i4));
        else {
            if (i4 < n_tile) v[u].x = __ldg(src + i4);
            if (i4 + 1 < n_tile) v[u].y = __ldg(src + i4 + 1);
            if (i4 + 2 < n_tile) v[u].z = __ldg(src + i4 + 2);
        }
    }
#pragma unroll
    for (int u = 0; u < ITEMS / 4; ++u) {
        const uint32_t i4 = (u * THREADS + threadIdx.x) * 4;
        const uint32_t xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t slot = (u * 4 + e) * THREADS + threadIdx.x;
            if (i4 + e < n_tile) {
                const uint32_t b2 = div_region(xs[e], pp);
                s_x[slot] = xs[e] - b2 * pp.R;
                s_b[slot] = (uint16_t)b2;
            } else {
                s_b[slot] = 0xFFFFu;
            }
        }
    }
    __syncthreads();
    const uint64_t base_bits = (uint64_t)b1 * pp.R1;
    const uint32_t R = pp.R;
    // the last level-1 bucket may hold fewer than P2 regions; its items never name a region past the end
    partition_tile<THREADS, ITEMS, PMAX, USE_MATCH>(smem_a, n_tile, pp.P2, pp.cursor2 + (size_t)b1 * pp.P2, pp.cap2,
                                                    pp.items2 + (size_t)b1 * pp.P2 * pp.cap2, pp,
                                                    [base_bits, R](uint32_t b, uint32_t x) { return base_bits + (uint64_t)b * R + x; });
}

// ---------------------------------------------------------------------------------------------- pass 3
enum ApplyMode { APPLY_SET = 0, APPLY_AND = 1, APPLY_OR = 2 };

// 16 byte flags (0 / 1) -> 16 bits, flag i -> bit i
__device__ __forceinline__ uint32_t pack16(uint4 f)
{
    uint32_t lo = __dp4a(f.x, 0x08040201u, 0u);
    lo = __dp4a(f.y, 0x80402010u, lo);
    uint32_t hi = __dp4a(f.z, 0x08040201u, 0u);
    hi = __dp4a(f.w, 0x80402010u, hi);
    return lo | (hi << 8);
}

// Persistent CTAs; region g = bits [g * R, (g + 1) * R) of the filter.  `prev` is read in AND mode (it may be the
// same array as `out`: a word is read and written by the same thread) and ignored otherwise.
// Flag of bit x lives at byte swz(x): the 16-byte chunk index is XORed with the low bits of the 128-byte row, so that
// the 8 LDS.128 of a thread packing row q (stride 128 bytes between lanes) are bank-conflict free.
__device__ __forceinline__ uint32_t swz(uint32_t x) { return x ^ ((x >> 3) & 0x70u); }

template <int MAXT, int MAXI>
__global__ void __launch_bounds__(MAXT) bf_apply3_kernel(PartParams pp, const uint4* prev, uint4* out,
                                                        uint64_t n16 /* uint4 words of the filter */, int mode)
{
    extern __shared__ __align__(16) unsigned char s_flags[];
    const int tid = threadIdx.x;
    const uint32_t THREADS = blockDim.x;
    const uint32_t R = pp.R, q_per = R >> 7;                  // uint4 words per region
    for (uint32_t i = tid; i < (R >> 4); i += THREADS) reinterpret_cast<uint4*>(s_flags)[i] = make_uint4(0, 0, 0, 0);
    uint32_t g = blockIdx.x;
    uint32_t n_cur = 0, cur[MAXI];
    auto load = [&](uint32_t gg, uint32_t* dst) -> uint32_t {
        const uint32_t n = min(pp.cursor2[gg], pp.cap2);
        const uint32_t* __restrict__ src = pp.items2 + (size_t)gg * pp.cap2;
#pragma unroll
        for (int u = 0; u < MAXI; ++u) {
            const uint32_t i = tid + u * THREADS;
            dst[u] = i < n ? __ldg(src + i) : 0xFFFFFFFFu;
        }
        return n;
    };
    if (g < pp.n_regions) n_cur = load(g, cur);
    __syncthreads();
    while (g < pp.n_regions) {
        const uint32_t* __restrict__ src = pp.items2 + (size_t)g * pp.cap2;
#pragma unroll
        for (int u = 0; u < MAXI; ++u)
            if (cur[u] != 0xFFFFFFFFu) s_flags[swz(cur[u])] = 1;
        for (uint32_t i = tid + MAXI * THREADS; i < n_cur; i += THREADS) s_flags[swz(__ldg(src + i))] = 1;
        const uint32_t g_next = g + gridDim.x;
        uint32_t n_nxt = 0, nxt[MAXI];
        if (g_next < pp.n_regions) n_nxt = load(g_next, nxt);
        __syncthreads();
        const uint64_t w0 = (uint64_t)g * q_per;
        for (uint32_t q = tid; q < q_per; q += THREADS) {
            if (w0 + q >= n16) break;
            uint4 pv = make_uint4(0, 0, 0, 0);
            if (mode != APPLY_SET) pv = prev[w0 + q];
            const uint4* f = reinterpret_cast<const uint4*>(s_flags + (size_t)q * 128);
            const uint32_t sw = q & 7u;
            uint4 r;
            r.x = pack16(f[0 ^ sw]) | (pack16(f[1 ^ sw]) << 16);
            r.y = pack16(f[2 ^ sw]) | (pack16(f[3 ^ sw]) << 16);
            r.z = pack16(f[4 ^ sw]) | (pack16(f[5 ^ sw]) << 16);
            r.w = pack16(f[6 ^ sw]) | (pack16(f[7 ^ sw]) << 16);
            if (mode == APPLY_AND) r = make_uint4(r.x & pv.x, r.y & pv.y, r.z & pv.z, r.w & pv.w);
            else if (mode == APPLY_OR) r = make_uint4(r.x | pv.x, r.y | pv.y, r.z | pv.z, r.w | pv.w);
            if (mode != APPLY_OR || (r.x ^ pv.x) | (r.y ^ pv.y) | (r.z ^ pv.z) | (r.w ^ pv.w)) out[w0 + q] = r;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAXI; ++u)
            if (cur[u] != 0xFFFFFFFFu) s_flags[swz(cur[u])] = 0;
        for (uint32_t i = tid + MAXI * THREADS; i < n_cur; i += THREADS) s_flags[swz(__ldg(src + i))] = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAXI; ++u) cur[u] = nxt[u];
        n_cur = n_nxt;
        g = g_next;
    }
}

// overflow list: SET / OR: set the bit; AND: set it where `prev` has it
__global__ void bf_overflow_kernel(const uint64_t* __restrict__ ovf, const unsigned long long* __restrict__ count, uint64_t cap,
                                   const uint32_t* __restrict__ prev, uint32_t* __restrict__ out, int mode)
{
    const uint64_t n = min((uint64_t)count[0], cap);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t idx = ovf[i];
        const uint32_t bit = 1u << (idx & 31);
        if (mode == APPLY_AND && !(__ldg(&prev[idx >> 5]) & bit)) continue;
        atomicOr(&out[idx >> 5], bit);
    }
}

}  // namespace nts
