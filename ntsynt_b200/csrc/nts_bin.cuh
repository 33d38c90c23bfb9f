// nts_bin.cuh -- (iii-a) partitioned Bloom insert.
//
// Random 32-byte read-modify-writes into a 14.8 GB array run at ~30 % of HBM peak and move 1.7x the
// algorithmic bytes (profiles/README.md).  The partitioned insert makes the filter traffic streaming:
//   pass 1 (bf_bin_kernel)   : hash a tile of k-mers, counting-sort their bit indices by filter REGION
//                              (2^region_shift bits, sized to sit in L2) in shared memory, and append each
//                              region's run to that region's bucket in global memory (one atomicAdd per run);
//   pass 2 (bf_apply_kernel) : walk the buckets in region order; the RED.ORs of one region hit in L2 and
//                              the region is written back to HBM once.
// Items that do not fit a bucket (heavy-hitter k-mers) are applied directly -- correctness never depends
// on the bucket capacities.
#pragma once
#include "nts_device.cuh"

namespace nts {

__device__ __forceinline__ void stage_tables_bin(HashTables* s_tabs, const HashTables* __restrict__ g_tabs, uint32_t k)
{
    for (uint32_t i = threadIdx.x; i < 16; i += blockDim.x) {
        s_tabs->roll[i] = g_tabs->roll[i];
    }
    for (uint32_t i = threadIdx.x; i < k * 4; i += blockDim.x) {
        s_tabs->init_f[i] = g_tabs->init_f[i];
        s_tabs->init_r[i] = g_tabs->init_r[i];
    }
}

struct BinParams {
    unsigned int* ovf_count = nullptr;   // not null: items past a bucket's capacity are only COUNTED (hash-range owned builds, where
                                         // the bits of a region belong to another GPU; the caller falls back if any were counted)
    uint32_t* items;               // bucket storage, bucket b at [bucket_off[b], bucket_off[b] + bucket_cap[b])
    const uint64_t* bucket_off;    // [P]
    const uint32_t* bucket_cap;    // [P]
    unsigned int* cursor;          // [P] items appended so far (may run past cap: clamp when reading)
    uint32_t n_buckets;
    uint32_t region_shift;         // bits per region = 1 << region_shift  (<= 32)
};

template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, 2) bf_bin_kernel(GenomeView g, const HashTables* __restrict__ g_tabs,
                                                             uint32_t* __restrict__ bits, uint64_t m, uint64_t mprime,
                                                             uint64_t total_valid, BinParams bp, uint32_t block0)
{
    constexpr int TILE = THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_bin[];
    HashTables* s_tabs = reinterpret_cast<HashTables*>(smem_bin);
    uint32_t* s_low = reinterpret_cast<uint32_t*>(s_tabs + 1);   // [TILE] bit index inside the region, slot = j*THREADS + tid
    uint32_t* s_br = s_low + TILE;                               // [TILE] bucket << 16 | rank inside the tile's run
    uint32_t* s_sorted = s_br + TILE;                            // [TILE] items in region order
    uint32_t* s_cnt = s_sorted + TILE;                           // [P + 1] count, then exclusive offset (+ sentinel n_tile)
    uint32_t* s_dst = s_cnt + bp.n_buckets + 1;                  // [P] global item index of the run's first element
    uint32_t* s_fit = s_dst + bp.n_buckets;                      // [P] how many items of the run fit the bucket
    __shared__ uint16_t s_wb[TILE / 32];                         // bucket of the item at sorted position 32 * w
    __shared__ uint32_t s_wsum[THREADS / 32];
    __shared__ uint32_t s_isl[2];                                  // islands of the tile's first and last k-mer
    const uint64_t tile0 = (uint64_t)(blockIdx.x + block0) * TILE;     // block0: first tile of this launch (staged inserts)
    const uint32_t n_tile = (uint32_t)min((uint64_t)TILE, total_valid - tile0);
    if (threadIdx.x == 0) s_isl[0] = find_island(g, tile0);
    if (threadIdx.x == 32) s_isl[1] = find_island(g, tile0 + n_tile - 1);
    stage_tables_bin(s_tabs, g_tabs, g.k);
    for (uint32_t b = threadIdx.x; b <= bp.n_buckets; b += THREADS) s_cnt[b] = 0;
    __syncthreads();
    const uint64_t v0 = tile0 + (uint64_t)threadIdx.x * ITEMS;
    const uint32_t n_mine = v0 < total_valid ? (uint32_t)min((uint64_t)ITEMS, total_valid - v0) : 0;
    const uint32_t low_mask = bp.region_shift >= 32 ? 0xFFFFFFFFu : ((1u << bp.region_shift) - 1);
    if (n_mine)
        hash_run(g, s_tabs, v0, n_mine, [&](uint32_t j, uint64_t h0, uint64_t) {
            const uint64_t idx = fast_mod(h0, m, mprime);
            const uint32_t b = (uint32_t)(idx >> bp.region_shift);
            const uint32_t slot = j * THREADS + threadIdx.x;
            s_low[slot] = (uint32_t)idx & low_mask;
            s_br[slot] = (b << 16) | atomicAdd(&s_cnt[b], 1u);
        }, s_isl[0], s_isl[1]);
    __syncthreads();
    // reserve the runs (one global atomic per non-empty bucket)
    for (uint32_t b = threadIdx.x; b < bp.n_buckets; b += THREADS) {
        const uint32_t c = s_cnt[b];
        uint32_t base = 0, fit = 0;
        if (c) {
            base = atomicAdd(&bp.cursor[b], c);
            const uint32_t cap = bp.bucket_cap[b];
            fit = base >= cap ? 0u : min(c, cap - base);
        }
        s_dst[b] = (uint32_t)bp.bucket_off[b] + base;             // the item buffer holds fewer than 2^32 items (host-checked)
        s_fit[b] = fit;
    }
    __syncthreads();
    // exclusive scan of the counts over the buckets
    {
        uint32_t carry = 0;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (uint32_t b0 = 0; b0 < bp.n_buckets; b0 += THREADS) {
            const uint32_t b = b0 + threadIdx.x;
            const uint32_t c = b < bp.n_buckets ? s_cnt[b] : 0;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            if (lane == 31) s_wsum[wid] = incl;
            __syncthreads();
            uint32_t woff = 0, total = 0;
            for (int i = 0; i < THREADS / 32; ++i) { if (i < wid) woff += s_wsum[i]; total += s_wsum[i]; }
            if (b < bp.n_buckets) s_cnt[b] = carry + woff + incl - c;
            carry += total;
            __syncthreads();
        }
        if (threadIdx.x == 0) s_cnt[bp.n_buckets] = n_tile;        // sentinel: end of the last run
    }
    __syncthreads();
    // place the items in region order; remember the bucket at every 32nd position
    for (uint32_t j = 0; j < n_mine; ++j) {
        const uint32_t slot = j * THREADS + threadIdx.x;
        const uint32_t br = s_br[slot];
        const uint32_t pos = s_cnt[br >> 16] + (br & 0xFFFFu);
        s_sorted[pos] = s_low[slot];
        if ((pos & 31u) == 0) s_wb[pos >> 5] = (uint16_t)(br >> 16);
    }
    __syncthreads();
    // write out, one thread per item: consecutive threads hold consecutive sorted positions, i.e. a warp writes
    // the tails / heads of two or three bucket runs as contiguous segments
    for (uint32_t i = threadIdx.x; i < n_tile; i += THREADS) {
        uint32_t b = s_wb[i >> 5];
        while (s_cnt[b + 1] <= i) ++b;                             // skips the (few) run ends inside the 32-group
        const uint32_t r = i - s_cnt[b];
        const uint32_t x = s_sorted[i];
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

// pass 2: CTA c applies chunk c of the concatenated buckets; chunks are in region order, so the CTAs
// resident at any moment touch one or two regions and their RED.ORs hit in L2.  A chunk is 4096 items =
// 256 threads x 4 independent 128-bit loads, all in flight before the first RED is issued (bucket offsets
// are multiples of 4 items, so the loads are aligned).
__global__ void __launch_bounds__(256) bf_apply_kernel(const uint32_t* __restrict__ items, const uint64_t* __restrict__ bucket_off,
                                const uint32_t* __restrict__ bucket_cap, const unsigned int* __restrict__ cursor,
                                const uint64_t* __restrict__ chunk_first /*[P+1] first chunk id of each bucket*/,
                                uint32_t n_buckets, uint32_t region_shift, uint32_t chunk_items /* = 4096 */,
                                uint32_t* __restrict__ bits)
{
    uint32_t lo = 0, hi = n_buckets;          // bucket of this chunk: last b with chunk_first[b] <= blockIdx.x
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const uint32_t b = lo;
    const uint32_t n = min(cursor[b], bucket_cap[b]);
    const uint64_t start = (uint64_t)(blockIdx.x - chunk_first[b]) * chunk_items;
    if (start >= n) return;
    const uint32_t cnt = (uint32_t)min((uint64_t)chunk_items, n - start);
    const uint32_t* src = items + bucket_off[b] + start;
    uint32_t* region = bits + (((uint64_t)b << region_shift) >> 5);
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    const uint32_t n4 = cnt >> 2;
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = threadIdx.x + u * 256;
        if (i < n4) v[u] = __ldg(src4 + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = threadIdx.x + u * 256;
        if (i < n4) {
            atomicOr(&region[v[u].x >> 5], 1u << (v[u].x & 31));
            atomicOr(&region[v[u].y >> 5], 1u << (v[u].y & 31));
            atomicOr(&region[v[u].z >> 5], 1u << (v[u].z & 31));
            atomicOr(&region[v[u].w >> 5], 1u << (v[u].w & 31));
        }
    }
    const uint32_t i = (n4 << 2) + threadIdx.x;           // up to 3 leftover items
    if (i < cnt) {
        const uint32_t x = __ldg(&src[i]);
        atomicOr(&region[x >> 5], 1u << (x & 31));
    }
}

// pass 2 of a hash-range OWNED build (multi-GPU): this GPU owns the filter bits [bit_lo, bit_hi); `items` / `cursor` are the
// buckets of ANOTHER GPU's binning pass, read over NVLink peer memory (coalesced 128-bit loads, 4 bytes per k-mer on the
// wire instead of the partial filters themselves).  CTA c applies chunk chunk0 + c; the launch covers the regions that
// overlap the owned range, and the items of the two boundary regions are filtered by bit index.
__global__ void __launch_bounds__(256) bf_apply_owned_kernel(const uint32_t* __restrict__ items, const uint64_t* __restrict__ bucket_off,
                                const uint32_t* __restrict__ bucket_cap, const unsigned int* __restrict__ cursor,
                                const uint64_t* __restrict__ chunk_first, uint32_t n_buckets, uint32_t region_shift,
                                uint32_t chunk_items, uint32_t* __restrict__ bits, uint64_t chunk0, uint64_t bit_lo, uint64_t bit_hi)
{
    const uint64_t chunk = (uint64_t)blockIdx.x + chunk0;
    uint32_t lo = 0, hi = n_buckets;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= chunk) lo = mid; else hi = mid;
    }
    const uint32_t b = lo;
    const uint32_t n = min(cursor[b], bucket_cap[b]);
    const uint64_t start = (chunk - chunk_first[b]) * chunk_items;
    if (start >= n) return;
    const uint32_t cnt = (uint32_t)min((uint64_t)chunk_items, n - start);
    const uint32_t* src = items + bucket_off[b] + start;
    const uint64_t rbase = (uint64_t)b << region_shift;
    uint32_t* region = bits + (rbase >> 5);
    // the region lies inside the owned range unless it is one of the two at its ends
    const uint64_t rlo = bit_lo > rbase ? bit_lo - rbase : 0;
    const uint64_t rhi = bit_hi - rbase;                         // (bit_hi > rbase for every launched region)
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    const uint32_t n4 = cnt >> 2;
    auto put = [&](uint32_t x) { if (x >= rlo && x < rhi) atomicOr(&region[x >> 5], 1u << (x & 31)); };
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = threadIdx.x + u * 256;
        if (i < n4) v[u] = src4[i];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = threadIdx.x + u * 256;
        if (i < n4) { put(v[u].x); put(v[u].y); put(v[u].z); put(v[u].w); }
    }
    const uint32_t i = (n4 << 2) + threadIdx.x;
    if (i < cnt) put(src[i]);
}

// dst[off16 .. off16 + n16) op= src[...]: 0 AND, 1 OR, 2 COPY (slices of filters: owned builds)
__global__ void bf_range_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, uint64_t off16, uint64_t n16, int op)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 a = src ? src[off16 + i] : make_uint4(0u, 0u, 0u, 0u);
        if (op == 0) { const uint4 d = dst[off16 + i]; a.x &= d.x; a.y &= d.y; a.z &= d.z; a.w &= d.w; }
        else if (op == 1) { const uint4 d = dst[off16 + i]; a.x |= d.x; a.y |= d.y; a.z |= d.z; a.w |= d.w; }
        dst[off16 + i] = a;
    }
}

}  // namespace nts
