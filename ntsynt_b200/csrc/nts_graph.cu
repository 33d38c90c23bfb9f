// nts_graph.cu -- kernel family (iv): minimizer join, adjacency edges and full-weight links.
//
// Replaces, for the bulk of the data, the Python/igraph path of the reference:
//   read_minimizers' duplicate removal      subprojects/ntJoin/bin/ntjoin_utils.py:182-192
//   filter_minimizers (G-way intersection)  subprojects/ntJoin/bin/ntjoin_utils.py:152-165
//   build_graph's adjacency edges + weights subprojects/ntJoin/bin/ntjoin_utils.py:97-113,132-135
//   filter_graph_global / find_paths for edges supported by ALL assemblies
//                                           subprojects/ntJoin/bin/ntjoin.py:78-87,114-136
// Design: one open-addressing hash table keyed by h1 holds a `seen` and a `dup` bitmask per key
// (bit a = assembly a).  A minimizer is a vertex iff seen == all and dup == 0.  Vertices are
// numbered by their rank in the ORIENTING assembly's filtered list, so that an edge supported by
// every assembly is always (i, i+1) and maximal chains are runs of a `link` bitmap.
#include <algorithm>
#include <new>

#include "nts_internal.h"

namespace nts {

constexpr uint64_t HT_EMPTY = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// find-or-insert every minimizer of assembly `a`; remember its slot; mark seen / dup
__global__ void join_insert_kernel(const uint64_t* __restrict__ h1, uint64_t n, uint32_t a,
                                   unsigned long long* __restrict__ keys, uint32_t* __restrict__ seen,
                                   uint32_t* __restrict__ dup, uint64_t mask, uint32_t* __restrict__ slot_of)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = h1[i];
    uint64_t slot;
    if (key == HT_EMPTY) {
        slot = mask + 1;   // dedicated slot for the one key that collides with the empty marker
    } else {
        slot = mix64(key) & mask;
        while (true) {
            unsigned long long old = atomicCAS(&keys[slot], (unsigned long long)HT_EMPTY, (unsigned long long)key);
            if (old == HT_EMPTY || old == key) break;
            slot = (slot + 1) & mask;
        }
    }
    slot_of[i] = (uint32_t)slot;
    const uint32_t bit = 1u << a;
    uint32_t old = atomicOr(&seen[slot], bit);
    if (old & bit) atomicOr(&dup[slot], bit);
}

__global__ void join_keep_kernel(const uint32_t* __restrict__ slot_of, uint64_t n, const uint32_t* __restrict__ seen,
                                 const uint32_t* __restrict__ dup, uint32_t full, uint32_t* __restrict__ keep)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = slot_of[i];
    keep[i] = (seen[s] == full && dup[s] == 0) ? 1u : 0u;
}

__global__ void join_lookup_kernel(const uint64_t* __restrict__ h1, uint64_t n, const unsigned long long* __restrict__ keys,
                                   const uint32_t* __restrict__ slot_vid, uint64_t mask, uint32_t* __restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = h1[i];
    uint32_t res = 0xFFFFFFFFu;
    if (key == HT_EMPTY) {
        res = slot_vid[mask + 1];
    } else {
        uint64_t slot = mix64(key) & mask;
        while (true) {
            const unsigned long long cur = keys[slot];
            if (cur == key) { res = slot_vid[slot]; break; }
            if (cur == HT_EMPTY) break;
            slot = (slot + 1) & mask;
        }
    }
    out[i] = res;
}

// ---- exclusive scan of 0/1 flags, 3 phases (block sums, scan of sums, apply)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;   // per thread

__global__ void scan_block_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ block_sums)
{
    __shared__ uint32_t s_w[SCAN_THREADS / 32];
    uint64_t base = ((uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    uint32_t sum = 0;
    for (int j = 0; j < SCAN_ITEMS; ++j) if (base + j < n) sum += in[base + j];
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) t += s_w[i];
        block_sums[blockIdx.x] = t;
    }
}

// single CTA exclusive scan over the block sums (n_blocks <= a few thousand); also returns the total
__global__ void scan_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t n_blocks, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (n_blocks + blockDim.x - 1) / blockDim.x;
    const uint32_t a = threadIdx.x * per, b = min(a + per, n_blocks);
    uint32_t sum = 0;
    for (uint32_t i = a; i < b; ++i) sum += block_sums[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) { uint32_t v = s_part[i]; s_part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t i = a; i < b; ++i) { uint32_t v = block_sums[i]; block_sums[i] = run; run += v; }
}

__global__ void scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint32_t* __restrict__ block_off,
                                  uint32_t* __restrict__ out)
{
    __shared__ uint32_t s_w[SCAN_THREADS / 32];
    uint64_t base = ((uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = (base + j < n) ? in[base + j] : 0; sum += v[j]; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int i = 0; i < wid; ++i) woff += s_w[i];
    uint32_t run = block_off[blockIdx.x] + woff + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) { if (base + j < n) out[base + j] = run; run += v[j]; }
}

// orienting assembly: vertex id = rank; publish it through the hash-table slot
__global__ void join_number_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ rank,
                                   const uint32_t* __restrict__ slot_of, uint64_t n, uint32_t* __restrict__ slot_vid)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    slot_vid[slot_of[i]] = rank[i];
}

// every assembly: scatter (pos, contig, rank) into vertex order and build the rank -> vertex map
__global__ void join_scatter_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ rank,
                                    const uint32_t* __restrict__ slot_of, const uint64_t* __restrict__ h1,
                                    const uint32_t* __restrict__ pos, const uint32_t* __restrict__ contig, uint64_t n,
                                    const uint32_t* __restrict__ slot_vid, uint32_t* __restrict__ v_pos,
                                    uint32_t* __restrict__ v_ctg, uint32_t* __restrict__ v_rank,
                                    uint32_t* __restrict__ inv, uint64_t* __restrict__ v_h1 /*nullable*/)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint32_t vid = slot_vid[slot_of[i]], r = rank[i];
    v_pos[vid] = pos[i];
    v_ctg[vid] = contig[i];
    v_rank[vid] = r;
    inv[r] = vid;
    if (v_h1) v_h1[vid] = h1[i];
}

// adjacency of vertices u, v in assembly b (same contig line, neighbouring ranks)
__device__ __forceinline__ bool adjacent_in(const uint32_t* __restrict__ v_ctg, const uint32_t* __restrict__ v_rank,
                                            uint64_t V, uint32_t b, uint32_t u, uint32_t v)
{
    const uint32_t ru = v_rank[(uint64_t)b * V + u], rv = v_rank[(uint64_t)b * V + v];
    return (ru + 1 == rv || rv + 1 == ru) && v_ctg[(uint64_t)b * V + u] == v_ctg[(uint64_t)b * V + v];
}

// link[i] = 1 iff edge (i, i+1) is supported by every assembly;  degree[v] = # distinct neighbours
__global__ void graph_links_kernel(const uint32_t* __restrict__ v_pos, const uint32_t* __restrict__ v_ctg,
                                   const uint32_t* __restrict__ v_rank, const uint32_t* __restrict__ inv, uint64_t V,
                                   uint32_t n_asm, uint8_t* __restrict__ link, uint8_t* __restrict__ degree,
                                   uint32_t* __restrict__ incmask, uint32_t* __restrict__ decmask,
                                   uint32_t* __restrict__ spread)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const uint32_t u = (uint32_t)i;
    if (i + 1 < V) {
        bool full = true;
        for (uint32_t b = 0; b < n_asm && full; ++b) full = adjacent_in(v_ctg, v_rank, V, b, u, u + 1);
        link[i] = full ? 1 : 0;
        // per-assembly direction of (i -> i+1) and the spread of |delta pos| (synteny_block.py:48-65,
        // ntsynt_synteny.py:364-368), consumed by the host as prefix sums over runs of links
        uint32_t inc = 0, dec = 0, dmax = 0, dmin = 0xFFFFFFFFu;
        for (uint32_t b = 0; b < n_asm; ++b) {
            const uint32_t p0 = v_pos[(uint64_t)b * V + u], p1 = v_pos[(uint64_t)b * V + u + 1];
            if (p1 > p0) inc |= 1u << b;
            if (p1 < p0) dec |= 1u << b;
            const uint32_t d = p1 > p0 ? p1 - p0 : p0 - p1;
            dmax = max(dmax, d); dmin = min(dmin, d);
        }
        incmask[i] = inc; decmask[i] = dec; spread[i] = dmax - dmin;
    } else {
        link[i] = 0; incmask[i] = 0; decmask[i] = 0; spread[i] = 0;
    }
    // distinct neighbours over all assemblies (at most 2 per assembly)
    uint32_t nb[64];
    uint32_t cnt = 0;
    for (uint32_t b = 0; b < n_asm; ++b) {
        const uint32_t r = v_rank[(uint64_t)b * V + u], c = v_ctg[(uint64_t)b * V + u];
        for (int side = 0; side < 2; ++side) {
            if (side == 0 && r == 0) continue;
            const uint64_t rr = side == 0 ? (uint64_t)r - 1 : (uint64_t)r + 1;
            if (rr >= V) continue;
            const uint32_t w = inv[(uint64_t)b * V + rr];
            if (v_ctg[(uint64_t)b * V + w] != c) continue;
            bool dupl = false;
            for (uint32_t t = 0; t < cnt; ++t) dupl |= (nb[t] == w);
            if (!dupl) nb[cnt++] = w;
        }
    }
    degree[i] = (uint8_t)cnt;
}

// edge enumeration in build_graph's first-insertion order: (assembly a, rank r) is a NEW edge iff no
// earlier assembly has the same unordered adjacency
__global__ void graph_edge_flags_kernel(const uint32_t* __restrict__ v_ctg, const uint32_t* __restrict__ v_rank,
                                        const uint32_t* __restrict__ inv, uint64_t V, uint32_t n_asm,
                                        uint32_t* __restrict__ is_new /*[n_asm*V]*/)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_asm * V) return;
    const uint32_t a = (uint32_t)(t / V);
    const uint64_t r = t % V;
    uint32_t flag = 0;
    if (r + 1 < V) {
        const uint32_t u = inv[(uint64_t)a * V + r], v = inv[(uint64_t)a * V + r + 1];
        if (v_ctg[(uint64_t)a * V + u] == v_ctg[(uint64_t)a * V + v]) {
            flag = 1;
            for (uint32_t b = 0; b < a && flag; ++b) if (adjacent_in(v_ctg, v_rank, V, b, u, v)) flag = 0;
        }
    }
    is_new[t] = flag;
}

__global__ void graph_edge_emit_kernel(const uint32_t* __restrict__ v_ctg, const uint32_t* __restrict__ v_rank,
                                       const uint32_t* __restrict__ inv, uint64_t V, uint32_t n_asm,
                                       const uint32_t* __restrict__ is_new, const uint32_t* __restrict__ off,
                                       uint32_t* __restrict__ e_u, uint32_t* __restrict__ e_v,
                                       uint32_t* __restrict__ e_support)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_asm * V || !is_new[t]) return;
    const uint32_t a = (uint32_t)(t / V);
    const uint64_t r = t % V;
    const uint32_t u = inv[(uint64_t)a * V + r], v = inv[(uint64_t)a * V + r + 1];
    uint32_t sup = 0;
    for (uint32_t b = 0; b < n_asm; ++b) if (adjacent_in(v_ctg, v_rank, V, b, u, v)) sup |= 1u << b;
    const uint32_t o = off[t];
    e_u[o] = u; e_v[o] = v; e_support[o] = sup;
}

__global__ void extract_bit_kernel(const uint32_t* __restrict__ mask, uint64_t n, uint32_t bit, uint32_t* __restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (mask[i] >> bit) & 1u;
}

static int exclusive_scan_u32(nts_ctx* ctx, const uint32_t* in, uint64_t n, uint32_t* out, uint32_t* total_host,
                              uint32_t* total_dev = nullptr)
{
    const uint64_t per_block = (uint64_t)SCAN_THREADS * SCAN_ITEMS;
    const uint32_t n_blocks = (uint32_t)std::max<uint64_t>(1, (n + per_block - 1) / per_block);
    DevBuf<uint32_t> sums, total;
    if (sums.alloc(n_blocks) != cudaSuccess || total.alloc(1) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (scan)");
    scan_block_sums_kernel<<<n_blocks, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p);
    scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(sums.p, n_blocks, total.p);
    scan_apply_kernel<<<n_blocks, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p, out);
    ctx->launches += 3;
    NTS_CUDA(cudaGetLastError());
    if (total_dev) NTS_CUDA(cudaMemcpyAsync(total_dev, total.p, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    NTS_CUDA(cudaMemcpyAsync(total_host, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

// host-ready columns for the graph stage (ntsynt_b200/synteny.py): positions widened to int64, the
// weight-filtered graph as nbr[v] = (left, right) neighbour ids or -1
__global__ void graph_export_kernel(const uint32_t* __restrict__ v_pos, const uint8_t* __restrict__ link, uint64_t V,
                                    uint32_t n_asm, long long* __restrict__ pos64, int2* __restrict__ nbr)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    for (uint32_t b = 0; b < n_asm; ++b) pos64[(uint64_t)b * V + i] = (long long)v_pos[(uint64_t)b * V + i];
    const int left = (i > 0 && link[i - 1]) ? (int)(i - 1) : -1;
    const int right = link[i] ? (int)(i + 1) : -1;
    nbr[i] = make_int2(left, right);
}

// the three sparse views the host walks: pairs (i, i+1) without a full-weight link, vertices of degree 3
// (graph simplification candidates), pairs whose position deltas spread by more than `bp` (indel splits).
// Unordered appends (the lists are short); the host sorts them.
__global__ void graph_sparse_lists_kernel(const uint8_t* __restrict__ link, const uint8_t* __restrict__ degree,
                                          const uint32_t* __restrict__ spread, uint64_t V, uint32_t bp,
                                          uint32_t* __restrict__ breaks, uint32_t* __restrict__ deg3,
                                          uint32_t* __restrict__ big, unsigned int* __restrict__ counts, uint32_t cap)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    if (i + 1 < V && !link[i]) { const unsigned int k = atomicAdd(&counts[0], 1u); if (k < cap) breaks[k] = (uint32_t)i; }
    if (degree[i] == 3) { const unsigned int k = atomicAdd(&counts[1], 1u); if (k < cap) deg3[k] = (uint32_t)i; }
    if (i + 1 < V && spread[i] > bp) { const unsigned int k = atomicAdd(&counts[2], 1u); if (k < cap) big[k] = (uint32_t)i; }
}


// ---- device-side lookups for the graph stage: the host walks a few thousand places; the O(V) columns stay here
// out[a * n + i] = (T) src[a * V + idx[i]]   (assembly-major columns; idx >= V gives `fill`)
template <typename S, typename T>
__global__ void gather_rows_kernel(const S* __restrict__ src, uint64_t V, uint32_t n_rows, const int64_t* __restrict__ idx,
                                   uint64_t n, T fill, T* __restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n_rows) return;
    const uint32_t a = (uint32_t)(t / n);
    const int64_t v = idx[t % n];
    out[t] = (v >= 0 && (uint64_t)v < V) ? (T)src[(uint64_t)a * V + v] : fill;
}

// up[a * n + i] = number of pairs (j, j+1), lo[i] <= j < hi[i], whose position increases in assembly a (down: decreases)
__global__ void range_sums_kernel(const uint32_t* __restrict__ cum_inc, const uint32_t* __restrict__ cum_dec, uint64_t V,
                                  uint32_t n_asm, const int64_t* __restrict__ lo, const int64_t* __restrict__ hi, uint64_t n,
                                  long long* __restrict__ up, long long* __restrict__ down)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n_asm) return;
    const uint32_t a = (uint32_t)(t / n);
    const uint64_t i = t % n;
    const uint64_t l = (uint64_t)min((long long)lo[i], (long long)V), h = (uint64_t)min((long long)hi[i], (long long)V);
    const uint32_t* ci = cum_inc + (uint64_t)a * (V + 1);
    const uint32_t* cd = cum_dec + (uint64_t)a * (V + 1);
    up[t] = (long long)ci[h] - (long long)ci[l];
    down[t] = (long long)cd[h] - (long long)cd[l];
}

// neighbourhood of simplification candidates: left / right neighbour (same contig line) and rank in every assembly
__global__ void cand_neigh_kernel(const uint32_t* __restrict__ v_ctg, const uint32_t* __restrict__ v_rank,
                                  const uint32_t* __restrict__ inv, uint64_t V, uint32_t n_asm, const int64_t* __restrict__ cand,
                                  uint64_t n, long long* __restrict__ left, long long* __restrict__ right, long long* __restrict__ rk)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n_asm) return;
    const uint64_t i = t / n_asm;
    const uint32_t a = (uint32_t)(t % n_asm);
    const uint64_t u = (uint64_t)cand[i];
    const uint64_t r = v_rank[(uint64_t)a * V + u];
    const uint32_t c = v_ctg[(uint64_t)a * V + u];
    long long lf = -1, rt = -1;
    if (r > 0) { const uint32_t x = inv[(uint64_t)a * V + r - 1]; if (v_ctg[(uint64_t)a * V + x] == c) lf = x; }
    if (r + 1 < V) { const uint32_t x = inv[(uint64_t)a * V + r + 1]; if (v_ctg[(uint64_t)a * V + x] == c) rt = x; }
    left[t] = lf; right[t] = rt; rk[t] = (long long)r;            // [i * n_asm + a]
}

__global__ void flag_big_kernel(const uint32_t* __restrict__ spread, uint64_t V, uint32_t bp, uint32_t* __restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V) flag[i] = (i + 1 < V && spread[i] > bp) ? 1u : 0u;
}

__global__ void compact_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ off, uint64_t V, uint32_t* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V && flag[i]) out[off[i]] = (uint32_t)i;
}

// runs of consecutive base vertices joined by full-weight (i, i+1) edges, two or more vertices long: vertex i starts
// one iff link[i-1] == 0 and link[i] == 1, ends one iff link[i-1] == 1 and link[i] == 0
__global__ void run_flags_kernel(const uint8_t* __restrict__ link, uint64_t V, uint32_t* __restrict__ is_start,
                                 uint32_t* __restrict__ is_end)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const bool prev = i > 0 && link[i - 1], cur = i + 1 < V && link[i];
    is_start[i] = (!prev && cur) ? 1u : 0u;
    is_end[i] = (prev && !cur) ? 1u : 0u;
}

__global__ void set_links_kernel(uint8_t* __restrict__ link, uint64_t V, const int64_t* __restrict__ idx,
                                 const uint8_t* __restrict__ val, uint64_t n)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && idx[t] >= 0 && (uint64_t)idx[t] < V) link[idx[t]] = val[t];
}

// ---- refinement rounds on the device (bin/ntsynt_synteny.py:205-280 + ntjoin_utils.py:152-193 for the new minimizers)
// One hash table keyed by h1 over the new tables of all assemblies: seen / dup masks give read_minimizers' duplicate
// removal, a `kept` mask gives the G-way intersection after the block filter.
struct RefineCtx {
    const uint32_t* seg_lo; const uint32_t* seg_hi; uint32_t n_seg;       // block segments, ascending, disjoint
    const uint32_t* term; uint32_t n_term;                                 // terminal vertices, ascending
    const uint64_t* x_key; const uint32_t* x_vid; uint32_t n_x;            // vertices added after round 0, by key
    const unsigned long long* g_keys; const uint32_t* g_slot_vid; uint64_t g_mask;   // the join table of the graph
};

__device__ __forceinline__ uint32_t refine_vid(const RefineCtx& c, uint64_t key)
{
    uint32_t res = 0xFFFFFFFFu;
    if (key == HT_EMPTY) res = c.g_slot_vid[c.g_mask + 1];
    else {
        uint64_t slot = mix64(key) & c.g_mask;
        for (;;) {
            const unsigned long long cur = c.g_keys[slot];
            if (cur == key) { res = c.g_slot_vid[slot]; break; }
            if (cur == HT_EMPTY) break;
            slot = (slot + 1) & c.g_mask;
        }
    }
    if (res == 0xFFFFFFFFu && c.n_x) {
        uint32_t lo = 0, hi = c.n_x;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (c.x_key[mid] < key) lo = mid + 1; else hi = mid; }
        if (lo < c.n_x && c.x_key[lo] == key) res = c.x_vid[lo];
    }
    return res;
}

// exists an interval [s, e) with s < b and a < e  (starts ascending, maxend = running maximum of the ends)
__device__ __forceinline__ bool interval_hit(const long long* __restrict__ starts, const long long* __restrict__ maxend, uint32_t n,
                                             long long a, long long b)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (starts[mid] < b) lo = mid + 1; else hi = mid; }
    return lo > 0 && maxend[lo - 1] > a;
}

// pass 2: keep[i] = not duplicated in its file, not an internal vertex of a block, not inside a block interval
__global__ void refine_keep_kernel(const uint64_t* __restrict__ h1, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ctg,
                                   const uint32_t* __restrict__ slot_of, uint64_t n, uint32_t a, const uint32_t* __restrict__ dup,
                                   uint32_t* __restrict__ kept_mask, RefineCtx c, const long long* __restrict__ iv_start,
                                   const long long* __restrict__ iv_maxend, uint32_t n_iv, uint32_t* __restrict__ keep,
                                   unsigned int* __restrict__ n_raw)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = 0;
    if (!(dup[slot_of[i]] & (1u << a))) {
        atomicAdd(n_raw, 1u);
        const uint32_t vid = refine_vid(c, h1[i]);
        bool internal = false;
        if (vid != 0xFFFFFFFFu && c.n_seg) {
            uint32_t lo = 0, hi = c.n_seg;                       // last segment with seg_lo <= vid
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (c.seg_lo[mid] <= vid) lo = mid + 1; else hi = mid; }
            if (lo > 0 && c.seg_hi[lo - 1] >= vid) {
                internal = true;
                uint32_t tl = 0, th = c.n_term;
                while (tl < th) { const uint32_t mid = (tl + th) >> 1; if (c.term[mid] < vid) tl = mid + 1; else th = mid; }
                if (tl < c.n_term && c.term[tl] == vid) internal = false;
            }
        }
        const long long key = ((long long)ctg[i] << 40) + (long long)pos[i];
        const bool inside = n_iv && interval_hit(iv_start, iv_maxend, n_iv, key, key + 1);
        k = (!internal && !inside) ? 1u : 0u;
        if (k) atomicOr(&kept_mask[slot_of[i]], 1u << a);
    }
    keep[i] = k;
}

// compact the kept entries (order preserved)
__global__ void refine_compact_kernel(const uint64_t* __restrict__ h1, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ctg,
                                      const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ keep,
                                      const uint32_t* __restrict__ off, uint64_t n, uint64_t* __restrict__ o_h1,
                                      uint32_t* __restrict__ o_pos, uint32_t* __restrict__ o_ctg, uint32_t* __restrict__ o_slot)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint32_t o = off[i];
    o_h1[o] = h1[i]; o_pos[o] = pos[i]; o_ctg[o] = ctg[i]; o_slot[o] = slot_of[i];
}

// over the kept list: cut[i] = a new sub-list starts at i (first entry, contig change, or the span from the previous
// kept minimizer overlaps a block interval); fin[i] = the key is kept in every assembly
__global__ void refine_cut_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ctg, const uint32_t* __restrict__ slot,
                                  uint64_t n, const uint32_t* __restrict__ kept_mask, uint32_t full,
                                  const long long* __restrict__ iv_start, const long long* __restrict__ iv_maxend, uint32_t n_iv,
                                  uint32_t* __restrict__ cut, uint32_t* __restrict__ fin)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 1;
    if (i > 0 && ctg[i] == ctg[i - 1]) {
        const long long a = ((long long)ctg[i - 1] << 40) + (long long)pos[i - 1];
        const long long b = ((long long)ctg[i] << 40) + (long long)pos[i];
        c = (n_iv && interval_hit(iv_start, iv_maxend, n_iv, a, b)) ? 1u : 0u;
    }
    cut[i] = c;
    fin[i] = kept_mask[slot[i]] == full ? 1u : 0u;
}

// final compaction: (h1, pos, ctg, sub = inclusive prefix of cut - 1) of the kept entries whose key is common
__global__ void refine_emit_kernel(const uint64_t* __restrict__ h1, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ctg,
                                   const uint32_t* __restrict__ cut, const uint32_t* __restrict__ cut_excl,
                                   const uint32_t* __restrict__ fin, const uint32_t* __restrict__ fin_off, uint64_t n,
                                   uint64_t* __restrict__ o_h1, uint32_t* __restrict__ o_pos, uint32_t* __restrict__ o_ctg,
                                   uint32_t* __restrict__ o_sub)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !fin[i]) return;
    const uint32_t o = fin_off[i];
    o_h1[o] = h1[i]; o_pos[o] = pos[i]; o_ctg[o] = ctg[i];
    o_sub[o] = cut_excl[i] + cut[i] - 1;
}

// first index of the sorted list with value >= x
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// (iv-c) collinear-path extraction for the runs that are plain chains of (i, i+1) edges -- all but a few of them.
// One thread per run [s, e] (e > s) of consecutive base vertices joined by full-weight edges:
//   find_paths          the run is a path from the end with the smaller position in the orienting assembly
//                       (subprojects/ntJoin/bin/ntjoin.py:89-102,114-136); equal positions: no path
//   find_synteny_blocks one block (the contig cannot change along full-weight edges); per assembly '+' if every step
//                       increases, '-' if every step decreases, else by the >= m percent rule, else the block is
//                       dropped and its vertices deleted (bin/ntsynt_synteny.py:66-106, bin/synteny_block.py:48-65)
//   check_for_indels    cut at every pair whose |dpos| spread exceeds --bp (bin/ntsynt_synteny.py:364-409)
//   filter_synteny_blocks pieces with fewer than min_mx minimizers are deleted (:411-426)
// Output (unordered appends): surviving blocks (lo, hi, dir, plus-mask), deleted vertex intervals, cut pairs.
struct RunOut {
    uint32_t* b_lo; uint32_t* b_hi; uint32_t* b_plus; int8_t* b_dir;     // [cap]
    uint32_t* r_lo; uint32_t* r_hi;                                       // [cap] deleted intervals
    uint32_t* cuts;                                                       // [cap] pairs (c, c+1) cut as indels
    unsigned int* counts;                                                 // [3] blocks, deleted intervals, cuts
    uint32_t cap;
};

__global__ void runs_to_blocks_kernel(const uint32_t* __restrict__ v_pos, const uint32_t* __restrict__ cum_inc,
                                      const uint32_t* __restrict__ cum_dec, uint64_t V, uint32_t n_asm, uint32_t orient,
                                      const uint32_t* __restrict__ big, uint32_t n_big, const int64_t* __restrict__ starts,
                                      const int64_t* __restrict__ ends, uint64_t n_runs, double m_pct, uint32_t min_mx, RunOut o)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_runs) return;
    const uint32_t s = (uint32_t)starts[t], e = (uint32_t)ends[t];
    if (e <= s) return;
    const uint32_t x = v_pos[(uint64_t)orient * V + s], y = v_pos[(uint64_t)orient * V + e];
    if (x == y) return;                                       // no path: the vertices stay as they are
    const int dir = x < y ? 1 : -1;
    const uint32_t n = e - s + 1;
    uint32_t plus = 0;
    bool unoriented = false;
    for (uint32_t a = 0; a < n_asm; ++a) {
        const uint32_t* ci = cum_inc + (uint64_t)a * (V + 1);
        const uint32_t* cd = cum_dec + (uint64_t)a * (V + 1);
        const uint32_t up = ci[e] - ci[s], down = cd[e] - cd[s];
        const uint32_t inc = dir > 0 ? up : down, dec = dir > 0 ? down : up;
        if (inc == n - 1) plus |= 1u << a;
        else if (dec == n - 1) { }
        else {
            const double positive = __dmul_rn(__ddiv_rn((double)inc, (double)(n - 1)), 100.0);
            const double negative = __dsub_rn(100.0, positive);
            if (positive >= m_pct) plus |= 1u << a;
            else if (negative >= m_pct) { }
            else unoriented = true;
        }
    }
    auto drop = [&](uint32_t lo, uint32_t hi) {
        const unsigned int k = atomicAdd(&o.counts[1], 1u);
        if (k < o.cap) { o.r_lo[k] = lo; o.r_hi[k] = hi; }
    };
    if (unoriented) { drop(s, e); return; }
    // pieces between the large-spread pairs c in [s, e)
    uint32_t ib = lower_bound_u32(big, n_big, s);
    uint32_t lo = s;
    for (;;) {
        const bool cut = ib < n_big && big[ib] < e;
        const uint32_t hi = cut ? big[ib] : e;
        if (cut) { const unsigned int k = atomicAdd(&o.counts[2], 1u); if (k < o.cap) o.cuts[k] = hi; }
        if (hi - lo + 1 >= min_mx) {
            const unsigned int k = atomicAdd(&o.counts[0], 1u);
            if (k < o.cap) { o.b_lo[k] = lo; o.b_hi[k] = hi; o.b_plus[k] = plus; o.b_dir[k] = (int8_t)dir; }
        } else {
            drop(lo, hi);
        }
        if (!cut) break;
        lo = hi + 1; ++ib;
    }
}

}  // namespace nts

using namespace nts;

struct nts_graph {
    nts_ctx* ctx = nullptr;
    uint32_t n_asm = 0;
    uint32_t order_asm = 0;
    uint64_t V = 0;
    DevBuf<uint64_t> v_h1;                       // [V]
    DevBuf<uint32_t> v_pos, v_ctg, v_rank, inv;  // [n_asm * V], assembly-major
    DevBuf<uint8_t> link, degree;                // [V]
    DevBuf<uint32_t> incmask, decmask, spread;   // [V] per (i, i+1)
    // the join table stays alive for nts_graph_lookup
    DevBuf<unsigned long long> keys;
    DevBuf<uint32_t> slot_vid, slot_ok;
    uint64_t cap = 0;
    // sparse lists (built lazily, per bp)
    bool lists_built = false;
    uint32_t lists_bp = 0;
    uint64_t n_lists[3] = {0, 0, 0};
    DevBuf<uint32_t> l_breaks, l_deg3, l_big;
    // prefix sums of the direction bits, [n_asm x (V + 1)] (built lazily, kept on the device)
    bool cums_built = false;
    DevBuf<uint32_t> cum_inc, cum_dec;
    // pairs with a large |dpos| spread, ascending (built lazily, per bp, kept on the device)
    bool big_built = false;
    uint32_t big_bp = 0, n_big = 0;
    DevBuf<uint32_t> d_big;
    // runs of the last nts_graph_runs call
    uint64_t n_runs = 0;
    DevBuf<uint32_t> run_s, run_e;
    // edges (built lazily)
    bool edges_built = false;
    uint64_t E = 0;
    DevBuf<uint32_t> e_u, e_v, e_support;
};

extern "C" {

int nts_graph_build(nts_ctx* ctx, nts_mxs* const* tables, uint32_t n_asm, uint32_t order_asm, nts_graph** out)
{
    if (!ctx || !tables || !out) return fail(NTS_ERR_ARG, "null argument");
    if (n_asm < 1 || n_asm > 32) return fail(NTS_ERR_ARG, "between 1 and 32 assemblies are supported");
    if (order_asm >= n_asm) return fail(NTS_ERR_ARG, "order_asm out of range");
    NTS_CUDA(cudaSetDevice(ctx->device));
    uint64_t total = 0, nmax = 0;
    for (uint32_t a = 0; a < n_asm; ++a) {
        if (!tables[a] || tables[a]->ctx != ctx) return fail(NTS_ERR_ARG, "bad minimizer table");
        total += tables[a]->count;
        nmax = std::max(nmax, tables[a]->count);
    }
    nts_graph* g = new (std::nothrow) nts_graph();
    if (!g) return fail(NTS_ERR_NOMEM, "host allocation failed");
    struct Guard { nts_graph* p; ~Guard() { delete p; } } guard{g};
    g->ctx = ctx; g->n_asm = n_asm; g->order_asm = order_asm;
    ProfScope prof(ctx, PROF_JOIN, (double)total);
    uint64_t cap = 1024;
    while (cap < total * 2) cap <<= 1;
    if (cap > 0x80000000ull) return fail(NTS_ERR_ARG, "too many minimizers for the join table");
    DevBuf<unsigned long long>& keys = g->keys;
    DevBuf<uint32_t>& slot_vid = g->slot_vid;
    DevBuf<uint32_t> seen, dup;
    g->cap = cap;
    std::vector<DevBuf<uint32_t>> slot_of(n_asm), keep(n_asm), rank(n_asm);
    if (keys.alloc(cap + 1) != cudaSuccess || seen.alloc(cap + 1) != cudaSuccess || dup.alloc(cap + 1) != cudaSuccess ||
        slot_vid.alloc(cap + 1) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (join table)");
    NTS_CUDA(cudaMemsetAsync(keys.p, 0xFF, (cap + 1) * 8, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(seen.p, 0, (cap + 1) * 4, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(dup.p, 0, (cap + 1) * 4, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(slot_vid.p, 0xFF, (cap + 1) * 4, ctx->stream));   // 0xFFFFFFFF = not a vertex
    for (uint32_t a = 0; a < n_asm; ++a) {
        const uint64_t n = tables[a]->count;
        if (slot_of[a].alloc(n) != cudaSuccess || keep[a].alloc(n) != cudaSuccess || rank[a].alloc(n) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (join)");
        if (!n) continue;
        join_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(tables[a]->h1.p, n, a, keys.p, seen.p, dup.p,
                                                                                cap - 1, slot_of[a].p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    const uint32_t full = n_asm == 32 ? 0xFFFFFFFFu : ((1u << n_asm) - 1);
    uint64_t V = 0;
    for (uint32_t a = 0; a < n_asm; ++a) {
        const uint64_t n = tables[a]->count;
        uint32_t tot = 0;
        if (n) {
            join_keep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(slot_of[a].p, n, seen.p, dup.p, full, keep[a].p);
            ctx->launches++;
            int rc = exclusive_scan_u32(ctx, keep[a].p, n, rank[a].p, &tot);
            if (rc) return rc;
        }
        if (a == 0) V = tot;
        else if (tot != V) return fail(NTS_ERR_STATE, "internal error: assemblies disagree on the number of common minimizers");
    }
    g->V = V;
    const uint64_t VA = std::max<uint64_t>(1, V * n_asm);
    if (g->v_h1.alloc(std::max<uint64_t>(1, V)) != cudaSuccess || g->v_pos.alloc(VA) != cudaSuccess || g->v_ctg.alloc(VA) != cudaSuccess ||
        g->v_rank.alloc(VA) != cudaSuccess || g->inv.alloc(VA) != cudaSuccess || g->link.alloc(std::max<uint64_t>(1, V)) != cudaSuccess ||
        g->degree.alloc(std::max<uint64_t>(1, V)) != cudaSuccess || g->incmask.alloc(std::max<uint64_t>(1, V)) != cudaSuccess ||
        g->decmask.alloc(std::max<uint64_t>(1, V)) != cudaSuccess || g->spread.alloc(std::max<uint64_t>(1, V)) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (vertex table)");
    if (V) {
        {
            const uint64_t n = tables[order_asm]->count;
            join_number_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keep[order_asm].p, rank[order_asm].p,
                                                                                    slot_of[order_asm].p, n, slot_vid.p);
            ctx->launches++;
        }
        for (uint32_t a = 0; a < n_asm; ++a) {
            const uint64_t n = tables[a]->count;
            join_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
                keep[a].p, rank[a].p, slot_of[a].p, tables[a]->h1.p, tables[a]->pos.p, tables[a]->contig.p, n, slot_vid.p,
                g->v_pos.p + a * V, g->v_ctg.p + a * V, g->v_rank.p + a * V, g->inv.p + a * V,
                a == order_asm ? g->v_h1.p : nullptr);
            ctx->launches++;
        }
        graph_links_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(g->v_pos.p, g->v_ctg.p, g->v_rank.p, g->inv.p, V,
                                                                                n_asm, g->link.p, g->degree.p, g->incmask.p,
                                                                                g->decmask.p, g->spread.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = g; guard.p = nullptr;
    return NTS_OK;
}

void nts_graph_destroy(nts_graph* g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    delete g;
}

uint64_t nts_graph_vertices(const nts_graph* g) { return g ? g->V : 0; }

int nts_graph_download_vertices(nts_graph* g, uint64_t* h1, uint32_t* pos, uint32_t* contig, uint32_t* rank,
                                uint8_t* link, uint8_t* degree)
{
    if (!g) return fail(NTS_ERR_ARG, "null argument");
    if (!g->V) return NTS_OK;
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t st = g->ctx->stream;
    const uint64_t V = g->V, VA = V * g->n_asm;
    if (h1) NTS_CUDA(copy_d2h(g->ctx, h1, g->v_h1.p, V * 8));
    if (pos) NTS_CUDA(copy_d2h(g->ctx, pos, g->v_pos.p, VA * 4));
    if (contig) NTS_CUDA(copy_d2h(g->ctx, contig, g->v_ctg.p, VA * 4));
    if (rank) NTS_CUDA(copy_d2h(g->ctx, rank, g->v_rank.p, VA * 4));
    if (link) NTS_CUDA(copy_d2h(g->ctx, link, g->link.p, V));
    if (degree) NTS_CUDA(copy_d2h(g->ctx, degree, g->degree.p, V));
    NTS_CUDA(cudaStreamSynchronize(st));
    return NTS_OK;
}

int nts_graph_download_links(nts_graph* g, uint32_t* inv, uint32_t* incmask, uint32_t* decmask, uint32_t* spread)
{
    if (!g) return fail(NTS_ERR_ARG, "null argument");
    if (!g->V) return NTS_OK;
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    const uint64_t V = g->V;
    if (inv) NTS_CUDA(copy_d2h(g->ctx, inv, g->inv.p, V * g->n_asm * 4));
    if (incmask) NTS_CUDA(copy_d2h(g->ctx, incmask, g->incmask.p, V * 4));
    if (decmask) NTS_CUDA(copy_d2h(g->ctx, decmask, g->decmask.p, V * 4));
    if (spread) NTS_CUDA(copy_d2h(g->ctx, spread, g->spread.p, V * 4));
    NTS_CUDA(cudaStreamSynchronize(g->ctx->stream));
    return NTS_OK;
}

/* vertex id of each h1 (0xFFFFFFFF if it is not a vertex) through the join table kept on the device */
int nts_graph_lookup(nts_graph* g, const uint64_t* h1, uint64_t n, uint32_t* vid_out)
{
    if (!g || (n && (!h1 || !vid_out))) return fail(NTS_ERR_ARG, "null argument");
    if (!n) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    DevBuf<uint64_t> d_k;
    DevBuf<uint32_t> d_o;
    if (d_k.alloc(n) != cudaSuccess || d_o.alloc(n) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (lookup)");
    NTS_CUDA(copy_h2d(ctx, d_k.p, h1, n * 8));
    {
        ProfScope prof(ctx, PROF_JOIN, (double)n);
        join_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_k.p, n, g->keys.p, g->slot_vid.p, g->cap - 1, d_o.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, vid_out, d_o.p, n * 4));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* prefix sums of the direction bits: ci/cd are [n_asm * (V+1)], assembly-major; ci[a*(V+1) + i] = number of
 * pairs (j, j+1) with j < i whose position increases in assembly a (cd: decreases) */
int nts_graph_download_cums(nts_graph* g, uint32_t* ci, uint32_t* cd)
{
    if (!g || !ci || !cd) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    const uint64_t V = g->V;
    if (!V) return NTS_OK;
    DevBuf<uint32_t> flags, sums;
    if (flags.alloc(V) != cudaSuccess || sums.alloc(V) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (cums)");
    ProfScope prof(ctx, PROF_JOIN, (double)V * g->n_asm * 2);
    for (uint32_t a = 0; a < g->n_asm; ++a)
        for (int which = 0; which < 2; ++which) {
            extract_bit_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(which ? g->decmask.p : g->incmask.p, V, a, flags.p);
            ctx->launches++;
            uint32_t total = 0;
            int rc = exclusive_scan_u32(ctx, flags.p, V, sums.p, &total);
            if (rc) return rc;
            uint32_t* dst = (which ? cd : ci) + (uint64_t)a * (V + 1);
            NTS_CUDA(copy_d2h(ctx, dst, sums.p, V * 4));
            NTS_CUDA(cudaStreamSynchronize(ctx->stream));
            dst[V] = total;
        }
    return NTS_OK;
}


/* Host-ready columns with room to grow: `cap` >= V is the row pitch (in elements) of pos64 / ctg32 and the
 * length of h1 / nbr; only the first V entries of each row are written.
 *   h1[cap] u64, pos64[n_asm x cap] i64, ctg32[n_asm x cap] i32, nbr[cap x 2] i32 (left, right neighbour in
 *   the weight-filtered graph or -1), conn[cap] u8 (conn[i] = 1 iff edge (i, i+1) has full weight). */
int nts_graph_download_host_arrays(nts_graph* g, uint64_t cap, uint64_t* h1, long long* pos64, int32_t* ctg32, int32_t* nbr,
                                   uint8_t* conn)
{
    if (!g || !h1 || !pos64 || !ctg32 || !nbr || !conn) return fail(NTS_ERR_ARG, "null argument");
    const uint64_t V = g->V;
    if (cap < V) return fail(NTS_ERR_ARG, "cap is smaller than the number of vertices");
    if (!V) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    DevBuf<long long> d_pos;
    DevBuf<int2> d_nbr;
    if (d_pos.alloc(V * g->n_asm) != cudaSuccess || d_nbr.alloc(V) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (export)");
    {
        ProfScope prof(ctx, PROF_JOIN, (double)V);
        graph_export_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(g->v_pos.p, g->link.p, V, g->n_asm, d_pos.p, d_nbr.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, h1, g->v_h1.p, V * 8));
    NTS_CUDA(cudaMemcpy2DAsync(pos64, cap * 8, d_pos.p, V * 8, V * 8, g->n_asm, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaMemcpy2DAsync(ctg32, cap * 4, g->v_ctg.p, V * 4, V * 4, g->n_asm, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += V * g->n_asm * 12;
    NTS_CUDA(copy_d2h(ctx, nbr, d_nbr.p, V * 8));
    NTS_CUDA(copy_d2h(ctx, conn, g->link.p, V));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* Sparse views of the vertex table (each list is returned sorted ascending):
 *   breaks: i with no full-weight edge (i, i+1), i < V - 1;  deg3: vertices with exactly 3 distinct neighbours;
 *   big: i whose pair (i, i+1) has max|dpos| - min|dpos| > bp.
 * Call with the three pointers NULL to get the counts, then with buffers of at least those sizes. */
int nts_graph_sparse_lists(nts_graph* g, uint32_t bp, uint32_t* breaks, uint32_t* deg3, uint32_t* big, uint64_t counts[3])
{
    if (!g || !counts) return fail(NTS_ERR_ARG, "null argument");
    const uint64_t V = g->V;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    if (!g->lists_built || g->lists_bp != bp) {
        g->n_lists[0] = g->n_lists[1] = g->n_lists[2] = 0;
        if (V) {
            uint32_t cap = (uint32_t)std::min<uint64_t>(V, 1u << 20);
            for (int attempt = 0; attempt < 2; ++attempt) {
                DevBuf<unsigned int> d_counts;
                if (g->l_breaks.alloc(cap) != cudaSuccess || g->l_deg3.alloc(cap) != cudaSuccess || g->l_big.alloc(cap) != cudaSuccess ||
                    d_counts.alloc(3) != cudaSuccess)
                    return fail(NTS_ERR_NOMEM, "device allocation failed (sparse lists)");
                NTS_CUDA(cudaMemsetAsync(d_counts.p, 0, 12, ctx->stream));
                {
                    ProfScope prof(ctx, PROF_JOIN, (double)V);
                    graph_sparse_lists_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(
                        g->link.p, g->degree.p, g->spread.p, V, bp, g->l_breaks.p, g->l_deg3.p, g->l_big.p, d_counts.p, cap);
                    ctx->launches++;
                }
                NTS_CUDA(cudaGetLastError());
                unsigned int h[3] = {0, 0, 0};
                NTS_CUDA(cudaMemcpyAsync(h, d_counts.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
                NTS_CUDA(cudaStreamSynchronize(ctx->stream));
                const unsigned int mx = std::max(h[0], std::max(h[1], h[2]));
                for (int i = 0; i < 3; ++i) g->n_lists[i] = h[i];
                if (mx <= cap) break;
                if (attempt == 1) return fail(NTS_ERR_STATE, "internal error: sparse lists overflow after retry");
                cap = mx;
            }
        }
        g->lists_built = true; g->lists_bp = bp;
    }
    for (int i = 0; i < 3; ++i) counts[i] = g->n_lists[i];
    uint32_t* dst[3] = {breaks, deg3, big};
    uint32_t* src[3] = {g->l_breaks.p, g->l_deg3.p, g->l_big.p};
    bool any = false;
    for (int i = 0; i < 3; ++i)
        if (dst[i] && g->n_lists[i]) { NTS_CUDA(copy_d2h(ctx, dst[i], src[i], g->n_lists[i] * 4)); any = true; }
    if (any) {
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < 3; ++i)
            if (dst[i] && g->n_lists[i]) std::sort(dst[i], dst[i] + g->n_lists[i]);
    }
    return NTS_OK;
}


static int build_cums(nts_graph* g)
{
    if (g->cums_built) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    const uint64_t V = g->V;
    const uint64_t n = (uint64_t)g->n_asm * (V + 1);
    if (g->cum_inc.alloc(n) != cudaSuccess || g->cum_dec.alloc(n) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (cums)");
    if (V) {
        DevBuf<uint32_t> flags;
        if (flags.alloc(V) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (cums)");
        ProfScope prof(ctx, PROF_GRAPH, (double)V * g->n_asm * 2);
        for (uint32_t a = 0; a < g->n_asm; ++a)
            for (int which = 0; which < 2; ++which) {
                extract_bit_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(which ? g->decmask.p : g->incmask.p, V, a, flags.p);
                ctx->launches++;
                uint32_t* dst = (which ? g->cum_dec.p : g->cum_inc.p) + (uint64_t)a * (V + 1);
                uint32_t total = 0;
                int rc = exclusive_scan_u32(ctx, flags.p, V, dst, &total, dst + V);
                if (rc) return rc;
            }
    } else {
        NTS_CUDA(cudaMemsetAsync(g->cum_inc.p, 0, n * 4, ctx->stream));
        NTS_CUDA(cudaMemsetAsync(g->cum_dec.p, 0, n * 4, ctx->stream));
    }
    g->cums_built = true;
    return NTS_OK;
}

static int build_big(nts_graph* g, uint32_t bp)
{
    if (g->big_built && g->big_bp == bp) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    const uint64_t V = g->V;
    g->n_big = 0;
    if (V) {
        DevBuf<uint32_t> flag, off;
        if (flag.alloc(V) != cudaSuccess || off.alloc(V) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (big)");
        ProfScope prof(ctx, PROF_GRAPH, (double)V);
        flag_big_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(g->spread.p, V, bp, flag.p);
        ctx->launches++;
        uint32_t tot = 0;
        int rc = exclusive_scan_u32(ctx, flag.p, V, off.p, &tot);
        if (rc) return rc;
        if (g->d_big.alloc(std::max<uint32_t>(1, tot)) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (big)");
        compact_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(flag.p, off.p, V, g->d_big.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        g->n_big = tot;
    } else if (g->d_big.alloc(1) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (big)");
    g->big_built = true; g->big_bp = bp;
    return NTS_OK;
}

/* Columns of the vertex table at the given vertex ids (what: 0 h1 -> u64[n]; 1 pos -> i64[n_asm x n]; 2 contig ->
 * i32[n_asm x n]; 3 rank -> u32[n_asm x n]; 4 inv (vertex at rank idx) -> u32[n_asm x n]).  Ids outside [0, V) give 0. */
int nts_graph_gather(nts_graph* g, int what, const int64_t* idx, uint64_t n, void* out)
{
    if (!g || (n && (!idx || !out)) || what < 0 || what > 4) return fail(NTS_ERR_ARG, "bad argument");
    if (!n) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    const uint64_t V = g->V;
    const uint32_t rows = what == 0 ? 1 : g->n_asm;
    const size_t esz = what == 0 || what == 1 ? 8 : 4;
    DevBuf<int64_t> d_idx;
    DevBuf<uint8_t> d_out;
    if (d_idx.alloc(n) != cudaSuccess || d_out.alloc(n * rows * esz) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (gather)");
    NTS_CUDA(copy_h2d(ctx, d_idx.p, idx, n * 8));
    const unsigned blocks = (unsigned)((n * rows + 255) / 256);
    {
        ProfScope prof(ctx, PROF_GRAPH, (double)n * rows);
        if (what == 0) gather_rows_kernel<uint64_t, uint64_t><<<blocks, 256, 0, ctx->stream>>>(g->v_h1.p, V, 1, d_idx.p, n, 0ull, reinterpret_cast<uint64_t*>(d_out.p));
        else if (what == 1) gather_rows_kernel<uint32_t, long long><<<blocks, 256, 0, ctx->stream>>>(g->v_pos.p, V, rows, d_idx.p, n, 0ll, reinterpret_cast<long long*>(d_out.p));
        else if (what == 2) gather_rows_kernel<uint32_t, int32_t><<<blocks, 256, 0, ctx->stream>>>(g->v_ctg.p, V, rows, d_idx.p, n, 0, reinterpret_cast<int32_t*>(d_out.p));
        else if (what == 3) gather_rows_kernel<uint32_t, uint32_t><<<blocks, 256, 0, ctx->stream>>>(g->v_rank.p, V, rows, d_idx.p, n, 0u, reinterpret_cast<uint32_t*>(d_out.p));
        else gather_rows_kernel<uint32_t, uint32_t><<<blocks, 256, 0, ctx->stream>>>(g->inv.p, V, rows, d_idx.p, n, 0u, reinterpret_cast<uint32_t*>(d_out.p));
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, out, d_out.p, n * rows * esz));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* up / down [n_asm x n] (i64): pairs (j, j+1) with lo[i] <= j < hi[i] whose position increases / decreases in each
 * assembly (lo, hi are clamped to V).  The prefix sums behind it are built once and stay on the device. */
int nts_graph_range_sums(nts_graph* g, const int64_t* lo, const int64_t* hi, uint64_t n, long long* up, long long* down)
{
    if (!g || (n && (!lo || !hi || !up || !down))) return fail(NTS_ERR_ARG, "null argument");
    if (!n) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    int rc = build_cums(g);
    if (rc) return rc;
    DevBuf<int64_t> d_lo, d_hi;
    DevBuf<long long> d_up, d_down;
    const uint64_t nn = n * g->n_asm;
    if (d_lo.alloc(n) != cudaSuccess || d_hi.alloc(n) != cudaSuccess || d_up.alloc(nn) != cudaSuccess || d_down.alloc(nn) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (range sums)");
    NTS_CUDA(copy_h2d(ctx, d_lo.p, lo, n * 8));
    NTS_CUDA(copy_h2d(ctx, d_hi.p, hi, n * 8));
    {
        ProfScope prof(ctx, PROF_GRAPH, (double)nn);
        range_sums_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(g->cum_inc.p, g->cum_dec.p, g->V, g->n_asm, d_lo.p, d_hi.p, n, d_up.p, d_down.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, up, d_up.p, nn * 8));
    NTS_CUDA(copy_d2h(ctx, down, d_down.p, nn * 8));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* left / right / rank [n x n_asm] (i64) of the simplification candidates: neighbour on the same contig line in every
 * assembly's filtered list, or -1 (what nts_host_simplify_neigh consumes) */
int nts_graph_neigh(nts_graph* g, const int64_t* cand, uint64_t n, long long* left, long long* right, long long* rk)
{
    if (!g || (n && (!cand || !left || !right || !rk))) return fail(NTS_ERR_ARG, "null argument");
    if (!n) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    const uint64_t nn = n * g->n_asm;
    DevBuf<int64_t> d_c;
    DevBuf<long long> d_l, d_r, d_k;
    if (d_c.alloc(n) != cudaSuccess || d_l.alloc(nn) != cudaSuccess || d_r.alloc(nn) != cudaSuccess || d_k.alloc(nn) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (neighbourhoods)");
    NTS_CUDA(copy_h2d(ctx, d_c.p, cand, n * 8));
    {
        ProfScope prof(ctx, PROF_GRAPH, (double)nn);
        cand_neigh_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(g->v_ctg.p, g->v_rank.p, g->inv.p, g->V, g->n_asm, d_c.p, n, d_l.p, d_r.p, d_k.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, left, d_l.p, nn * 8));
    NTS_CUDA(copy_d2h(ctx, right, d_r.p, nn * 8));
    NTS_CUDA(copy_d2h(ctx, rk, d_k.p, nn * 8));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* (iv-c) on the device: paths, blocks, indel splits and the min_mx filter for runs [starts[i], ends[i]] of base
 * vertices joined by (i, i+1) full-weight edges (see runs_to_blocks_kernel).  The caller passes only runs whose
 * vertices still hold their round-0 positions.  Outputs are unordered; counts[3] = blocks, deleted intervals, cuts.
 * Call with the output pointers NULL to run the kernel and get the counts, then again to fetch (same arguments). */
int nts_graph_runs_to_blocks(nts_graph* g, const int64_t* starts, const int64_t* ends, uint64_t n_runs, uint32_t bp,
                             double m_pct, uint32_t min_mx, uint32_t* b_lo, uint32_t* b_hi, uint32_t* b_plus, int8_t* b_dir,
                             uint32_t* r_lo, uint32_t* r_hi, uint32_t* cuts, uint64_t counts[3], uint64_t cap)
{
    if (!g || !counts || (n_runs && (!starts || !ends))) return fail(NTS_ERR_ARG, "null argument");
    counts[0] = counts[1] = counts[2] = 0;
    if (!n_runs) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    int rc = build_cums(g);
    if (rc || (rc = build_big(g, bp))) return rc;
    // every run gives at most (cuts inside + 1) pieces; cuts <= n_big
    const uint64_t need = n_runs + g->n_big + 1;
    if (cap < need) return fail(NTS_ERR_ARG, "output capacity must be at least n_runs + nts_graph_big_count + 1");
    DevBuf<int64_t> d_s, d_e;
    DevBuf<uint32_t> d_u32;
    DevBuf<int8_t> d_dir;
    DevBuf<unsigned int> d_cnt;
    if (d_s.alloc(n_runs) != cudaSuccess || d_e.alloc(n_runs) != cudaSuccess || d_u32.alloc(need * 6) != cudaSuccess ||
        d_dir.alloc(need) != cudaSuccess || d_cnt.alloc(3) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (runs)");
    NTS_CUDA(copy_h2d(ctx, d_s.p, starts, n_runs * 8));
    NTS_CUDA(copy_h2d(ctx, d_e.p, ends, n_runs * 8));
    NTS_CUDA(cudaMemsetAsync(d_cnt.p, 0, 12, ctx->stream));
    RunOut o;
    o.b_lo = d_u32.p; o.b_hi = d_u32.p + need; o.b_plus = d_u32.p + 2 * need; o.r_lo = d_u32.p + 3 * need;
    o.r_hi = d_u32.p + 4 * need; o.cuts = d_u32.p + 5 * need; o.b_dir = d_dir.p; o.counts = d_cnt.p; o.cap = (uint32_t)need;
    {
        ProfScope prof(ctx, PROF_GRAPH, (double)n_runs);
        runs_to_blocks_kernel<<<(unsigned)((n_runs + 127) / 128), 128, 0, ctx->stream>>>(
            g->v_pos.p, g->cum_inc.p, g->cum_dec.p, g->V, g->n_asm, g->order_asm, g->d_big.p, g->n_big, d_s.p, d_e.p, n_runs,
            m_pct, min_mx, o);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    unsigned int h[3] = {0, 0, 0};
    NTS_CUDA(cudaMemcpyAsync(h, d_cnt.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) counts[i] = h[i];
    if (h[0] > need || h[1] > need || h[2] > need) return fail(NTS_ERR_STATE, "internal error: run output overflow");
    if (h[0]) {
        NTS_CUDA(copy_d2h(ctx, b_lo, o.b_lo, h[0] * 4)); NTS_CUDA(copy_d2h(ctx, b_hi, o.b_hi, h[0] * 4));
        NTS_CUDA(copy_d2h(ctx, b_plus, o.b_plus, h[0] * 4)); NTS_CUDA(copy_d2h(ctx, b_dir, o.b_dir, h[0]));
    }
    if (h[1]) { NTS_CUDA(copy_d2h(ctx, r_lo, o.r_lo, h[1] * 4)); NTS_CUDA(copy_d2h(ctx, r_hi, o.r_hi, h[1] * 4)); }
    if (h[2]) NTS_CUDA(copy_d2h(ctx, cuts, o.cuts, h[2] * 4));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}


/* The host edits the weight-filtered graph (simplification, indel cuts, deleted blocks, refinement splices); the pairs
 * (i, i+1) it changed are pushed back so that the device's link bitmap stays the truth for nts_graph_runs. */
int nts_graph_set_links(nts_graph* g, const int64_t* idx, const uint8_t* val, uint64_t n)
{
    if (!g || (n && (!idx || !val))) return fail(NTS_ERR_ARG, "null argument");
    if (!n || !g->V) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    DevBuf<int64_t> d_i;
    DevBuf<uint8_t> d_v;
    if (d_i.alloc(n) != cudaSuccess || d_v.alloc(n) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (links)");
    NTS_CUDA(copy_h2d(ctx, d_i.p, idx, n * 8));
    NTS_CUDA(copy_h2d(ctx, d_v.p, val, n));
    set_links_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(g->link.p, g->V, d_i.p, d_v.p, n);
    ctx->launches++;
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));      // the host buffers may go away
    return NTS_OK;
}

/* Chain extraction (kernel iv-c, find_paths of subprojects/ntJoin/bin/ntjoin.py:114-136 for the (i, i+1) part of the
 * graph): the maximal runs of two or more base vertices joined by full-weight edges, as ascending (starts, ends).
 * Call with NULL outputs for the count, then with arrays of that length. */
int nts_graph_runs(nts_graph* g, int64_t* starts, int64_t* ends, uint64_t* n_runs)
{
    if (!g || !n_runs) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    const uint64_t V = g->V;
    if (!starts || !ends) {
        g->n_runs = 0;
        if (V) {
            DevBuf<uint32_t> fs, fe, os, oe;
            if (fs.alloc(V) != cudaSuccess || fe.alloc(V) != cudaSuccess || os.alloc(V) != cudaSuccess || oe.alloc(V) != cudaSuccess)
                return fail(NTS_ERR_NOMEM, "device allocation failed (runs)");
            ProfScope prof(ctx, PROF_GRAPH, (double)V);
            run_flags_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(g->link.p, V, fs.p, fe.p);
            ctx->launches++;
            uint32_t ns = 0, ne = 0;
            int rc = exclusive_scan_u32(ctx, fs.p, V, os.p, &ns);
            if (rc || (rc = exclusive_scan_u32(ctx, fe.p, V, oe.p, &ne))) return rc;
            if (ns != ne) return fail(NTS_ERR_STATE, "internal error: run starts and ends disagree");
            if (g->run_s.alloc(std::max<uint32_t>(1, ns)) != cudaSuccess || g->run_e.alloc(std::max<uint32_t>(1, ns)) != cudaSuccess)
                return fail(NTS_ERR_NOMEM, "device allocation failed (runs)");
            compact_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(fs.p, os.p, V, g->run_s.p);
            compact_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(fe.p, oe.p, V, g->run_e.p);
            ctx->launches += 2;
            NTS_CUDA(cudaGetLastError());
            NTS_CUDA(cudaStreamSynchronize(ctx->stream));
            g->n_runs = ns;
        }
        *n_runs = g->n_runs;
        return NTS_OK;
    }
    *n_runs = g->n_runs;
    if (g->n_runs) {
        std::vector<uint32_t> hs(g->n_runs), he(g->n_runs);
        NTS_CUDA(copy_d2h(ctx, hs.data(), g->run_s.p, g->n_runs * 4));
        NTS_CUDA(copy_d2h(ctx, he.data(), g->run_e.p, g->n_runs * 4));
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < g->n_runs; ++i) { starts[i] = hs[i]; ends[i] = he[i]; }
    }
    return NTS_OK;
}


/* One refinement round's minimizer filtering on the device, for the new tables of all assemblies at once:
 *   read_minimizers        drop every h1 seen more than once in its file      (ntjoin_utils.py:182-192)
 *   filter_minimizers_synteny_blocks  drop the internal minimizers of the blocks and the ones inside a block interval,
 *                          start a new sub-list where the span from the previous kept minimizer overlaps a block
 *                                                                              (bin/ntsynt_synteny.py:205-280)
 *   filter_minimizers      keep the keys that survive in every assembly       (ntjoin_utils.py:152-165)
 * Inputs (host, small): the blocks' segments (seg_lo/seg_hi, ascending, disjoint) and terminal vertices (ascending);
 * the vertices added after round 0 (x_key ascending, x_vid); per assembly a the block intervals
 * iv_start[iv_off[a] .. iv_off[a+1]) ascending with iv_maxend their running maximum end (coordinates contig << 40 | pos).
 * Outputs per assembly: n_raw[a] (entries after duplicate removal), and the surviving (h1, pos, ctg, sub-list id),
 * assembly a at out_off[a] .. out_off[a+1) of the out_* arrays (capacity out_cap entries in total; if the total
 * exceeds it only out_off is valid).  n_common = number of distinct surviving keys. */
int nts_graph_refine_filter(nts_graph* g, nts_mxs* const* tables, const uint32_t* seg_lo, const uint32_t* seg_hi, uint32_t n_seg,
                            const uint32_t* term, uint32_t n_term, const uint64_t* x_key, const uint32_t* x_vid, uint32_t n_x,
                            const long long* iv_start, const long long* iv_maxend, const uint64_t* iv_off, uint64_t* n_raw,
                            uint64_t* out_off, uint64_t* out_h1, uint32_t* out_pos, uint32_t* out_ctg, uint32_t* out_sub,
                            uint64_t out_cap)
{
    if (!g || !tables || !iv_off || !n_raw || !out_off) return fail(NTS_ERR_ARG, "null argument");
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    const uint32_t G = g->n_asm;
    uint64_t total = 0;
    for (uint32_t a = 0; a < G; ++a) {
        if (!tables[a] || tables[a]->ctx != ctx) return fail(NTS_ERR_ARG, "bad minimizer table");
        total += tables[a]->count;
    }
    out_off[0] = 0;
    for (uint32_t a = 0; a < G; ++a) { n_raw[a] = 0; out_off[a + 1] = 0; }
    if (!total) return NTS_OK;
    ProfScope prof(ctx, PROF_GRAPH, (double)total);
    uint64_t cap = 1024;
    while (cap < total * 2) cap <<= 1;
    DevBuf<unsigned long long> keys;
    DevBuf<uint32_t> seen, dup, kept;
    if (keys.alloc(cap + 1) != cudaSuccess || seen.alloc(cap + 1) != cudaSuccess || dup.alloc(cap + 1) != cudaSuccess ||
        kept.alloc(cap + 1) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (refine table)");
    NTS_CUDA(cudaMemsetAsync(keys.p, 0xFF, (cap + 1) * 8, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(seen.p, 0, (cap + 1) * 4, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(dup.p, 0, (cap + 1) * 4, ctx->stream));
    NTS_CUDA(cudaMemsetAsync(kept.p, 0, (cap + 1) * 4, ctx->stream));
    // small host inputs
    DevBuf<uint32_t> d_seg_lo, d_seg_hi, d_term, d_xvid;
    DevBuf<uint64_t> d_xkey;
    DevBuf<long long> d_ivs, d_ive;
    const uint64_t n_iv_all = iv_off[G];
    if (d_seg_lo.alloc(n_seg) != cudaSuccess || d_seg_hi.alloc(n_seg) != cudaSuccess || d_term.alloc(n_term) != cudaSuccess ||
        d_xkey.alloc(n_x) != cudaSuccess || d_xvid.alloc(n_x) != cudaSuccess || d_ivs.alloc(n_iv_all) != cudaSuccess ||
        d_ive.alloc(n_iv_all) != cudaSuccess)
        return fail(NTS_ERR_NOMEM, "device allocation failed (refine inputs)");
    if (n_seg) { NTS_CUDA(copy_h2d(ctx, d_seg_lo.p, seg_lo, n_seg * 4)); NTS_CUDA(copy_h2d(ctx, d_seg_hi.p, seg_hi, n_seg * 4)); }
    if (n_term) NTS_CUDA(copy_h2d(ctx, d_term.p, term, n_term * 4));
    if (n_x) { NTS_CUDA(copy_h2d(ctx, d_xkey.p, x_key, n_x * 8)); NTS_CUDA(copy_h2d(ctx, d_xvid.p, x_vid, n_x * 4)); }
    if (n_iv_all) { NTS_CUDA(copy_h2d(ctx, d_ivs.p, iv_start, n_iv_all * 8)); NTS_CUDA(copy_h2d(ctx, d_ive.p, iv_maxend, n_iv_all * 8)); }
    RefineCtx rc;
    rc.seg_lo = d_seg_lo.p; rc.seg_hi = d_seg_hi.p; rc.n_seg = n_seg; rc.term = d_term.p; rc.n_term = n_term;
    rc.x_key = d_xkey.p; rc.x_vid = d_xvid.p; rc.n_x = n_x; rc.g_keys = g->keys.p; rc.g_slot_vid = g->slot_vid.p; rc.g_mask = g->cap - 1;
    std::vector<DevBuf<uint32_t>> slot_of(G), keep(G), off(G);
    DevBuf<unsigned int> d_raw;
    if (d_raw.alloc(G) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed");
    NTS_CUDA(cudaMemsetAsync(d_raw.p, 0, G * 4, ctx->stream));
    for (uint32_t a = 0; a < G; ++a) {
        const uint64_t n = tables[a]->count;
        if (slot_of[a].alloc(n) != cudaSuccess || keep[a].alloc(n) != cudaSuccess || off[a].alloc(n) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (refine)");
        if (!n) continue;
        join_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(tables[a]->h1.p, n, a, keys.p, seen.p, dup.p, cap - 1, slot_of[a].p);
        ctx->launches++;
    }
    for (uint32_t a = 0; a < G; ++a) {
        const uint64_t n = tables[a]->count;
        if (!n) continue;
        const uint32_t n_iv = (uint32_t)(iv_off[a + 1] - iv_off[a]);
        refine_keep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            tables[a]->h1.p, tables[a]->pos.p, tables[a]->contig.p, slot_of[a].p, n, a, dup.p, kept.p, rc, d_ivs.p + iv_off[a],
            d_ive.p + iv_off[a], n_iv, keep[a].p, d_raw.p + a);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    const uint32_t full = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1);
    // per assembly: compact the kept entries, cut into sub-lists, keep the common keys
    struct Part { DevBuf<uint64_t> h1; DevBuf<uint32_t> pos, ctg, sub; uint32_t n = 0; };
    std::vector<Part> parts(G);
    for (uint32_t a = 0; a < G; ++a) {
        const uint64_t n = tables[a]->count;
        if (!n) continue;
        uint32_t n_keep = 0;
        int rcode = exclusive_scan_u32(ctx, keep[a].p, n, off[a].p, &n_keep);
        if (rcode) return rcode;
        if (!n_keep) continue;
        DevBuf<uint64_t> k_h1;
        DevBuf<uint32_t> k_pos, k_ctg, k_slot, cut, fin, cut_x, fin_x;
        if (k_h1.alloc(n_keep) != cudaSuccess || k_pos.alloc(n_keep) != cudaSuccess || k_ctg.alloc(n_keep) != cudaSuccess ||
            k_slot.alloc(n_keep) != cudaSuccess || cut.alloc(n_keep) != cudaSuccess || fin.alloc(n_keep) != cudaSuccess ||
            cut_x.alloc(n_keep) != cudaSuccess || fin_x.alloc(n_keep) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (refine kept)");
        refine_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            tables[a]->h1.p, tables[a]->pos.p, tables[a]->contig.p, slot_of[a].p, keep[a].p, off[a].p, n, k_h1.p, k_pos.p, k_ctg.p, k_slot.p);
        const uint32_t n_iv = (uint32_t)(iv_off[a + 1] - iv_off[a]);
        refine_cut_kernel<<<(unsigned)((n_keep + 255) / 256), 256, 0, ctx->stream>>>(k_pos.p, k_ctg.p, k_slot.p, n_keep, kept.p, full,
                                                                                    d_ivs.p + iv_off[a], d_ive.p + iv_off[a], n_iv, cut.p, fin.p);
        ctx->launches += 2;
        uint32_t n_cut = 0, n_fin = 0;
        if ((rcode = exclusive_scan_u32(ctx, cut.p, n_keep, cut_x.p, &n_cut)) || (rcode = exclusive_scan_u32(ctx, fin.p, n_keep, fin_x.p, &n_fin)))
            return rcode;
        Part& pt = parts[a];
        pt.n = n_fin;
        if (!n_fin) continue;
        if (pt.h1.alloc(n_fin) != cudaSuccess || pt.pos.alloc(n_fin) != cudaSuccess || pt.ctg.alloc(n_fin) != cudaSuccess || pt.sub.alloc(n_fin) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (refine out)");
        refine_emit_kernel<<<(unsigned)((n_keep + 255) / 256), 256, 0, ctx->stream>>>(k_h1.p, k_pos.p, k_ctg.p, cut.p, cut_x.p, fin.p, fin_x.p,
                                                                                     n_keep, pt.h1.p, pt.pos.p, pt.ctg.p, pt.sub.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));            // the kept buffers go out of scope
    }
    std::vector<unsigned int> h_raw(G, 0);
    NTS_CUDA(cudaMemcpyAsync(h_raw.data(), d_raw.p, G * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t at = 0;
    for (uint32_t a = 0; a < G; ++a) { n_raw[a] = h_raw[a]; out_off[a] = at; at += parts[a].n; }
    out_off[G] = at;
    if (at > out_cap || !at) return NTS_OK;
    if (!out_h1 || !out_pos || !out_ctg || !out_sub) return fail(NTS_ERR_ARG, "null output");
    for (uint32_t a = 0; a < G; ++a) {
        const Part& pt = parts[a];
        if (!pt.n) continue;
        NTS_CUDA(copy_d2h(ctx, out_h1 + out_off[a], pt.h1.p, (size_t)pt.n * 8));
        NTS_CUDA(copy_d2h(ctx, out_pos + out_off[a], pt.pos.p, (size_t)pt.n * 4));
        NTS_CUDA(copy_d2h(ctx, out_ctg + out_off[a], pt.ctg.p, (size_t)pt.n * 4));
        NTS_CUDA(copy_d2h(ctx, out_sub + out_off[a], pt.sub.p, (size_t)pt.n * 4));
    }
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

/* number of pairs whose |dpos| spread exceeds bp (sizes the outputs of nts_graph_runs_to_blocks) */
int nts_graph_big_count(nts_graph* g, uint32_t bp, uint64_t* n_big)
{
    if (!g || !n_big) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    int rc = build_big(g, bp);
    if (rc) return rc;
    *n_big = g->n_big;
    return NTS_OK;
}

/* conn[cap] u8 (conn[i] = 1 iff edge (i, i+1) has full weight) and nbr[cap x 2] i32 alone: the lean form of
 * nts_graph_download_host_arrays (positions, contigs and hashes stay on the device: nts_graph_gather) */
int nts_graph_download_links_nbr(nts_graph* g, uint64_t cap, int32_t* nbr, uint8_t* conn)
{
    if (!g || !nbr || !conn) return fail(NTS_ERR_ARG, "null argument");
    const uint64_t V = g->V;
    if (cap < V) return fail(NTS_ERR_ARG, "cap is smaller than the number of vertices");
    if (!V) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    NTS_CUDA(cudaSetDevice(ctx->device));
    DevBuf<long long> d_pos;
    DevBuf<int2> d_nbr;
    if (d_pos.alloc(1) != cudaSuccess || d_nbr.alloc(V) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (export)");
    {
        ProfScope prof(ctx, PROF_GRAPH, (double)V);
        graph_export_kernel<<<(unsigned)((V + 255) / 256), 256, 0, ctx->stream>>>(g->v_pos.p, g->link.p, V, 0, d_pos.p, d_nbr.p);
        ctx->launches++;
    }
    NTS_CUDA(cudaGetLastError());
    NTS_CUDA(copy_d2h(ctx, nbr, d_nbr.p, V * 8));
    NTS_CUDA(copy_d2h(ctx, conn, g->link.p, V));
    NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NTS_OK;
}

static int build_edges(nts_graph* g)
{
    if (g->edges_built) return NTS_OK;
    nts_ctx* ctx = g->ctx;
    const uint64_t V = g->V, N = V * g->n_asm;
    g->E = 0;
    if (N) {
        ProfScope prof(ctx, PROF_EDGES, (double)N);
        DevBuf<uint32_t> is_new, off;
        if (is_new.alloc(N) != cudaSuccess || off.alloc(N) != cudaSuccess) return fail(NTS_ERR_NOMEM, "device allocation failed (edges)");
        graph_edge_flags_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(g->v_ctg.p, g->v_rank.p, g->inv.p, V, g->n_asm, is_new.p);
        ctx->launches++;
        uint32_t tot = 0;
        int rc = exclusive_scan_u32(ctx, is_new.p, N, off.p, &tot);
        if (rc) return rc;
        g->E = tot;
        if (g->e_u.alloc(std::max<uint32_t>(1, tot)) != cudaSuccess || g->e_v.alloc(std::max<uint32_t>(1, tot)) != cudaSuccess ||
            g->e_support.alloc(std::max<uint32_t>(1, tot)) != cudaSuccess)
            return fail(NTS_ERR_NOMEM, "device allocation failed (edge table)");
        graph_edge_emit_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(g->v_ctg.p, g->v_rank.p, g->inv.p, V, g->n_asm,
                                                                                    is_new.p, off.p, g->e_u.p, g->e_v.p, g->e_support.p);
        ctx->launches++;
        NTS_CUDA(cudaGetLastError());
        NTS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    g->edges_built = true;
    return NTS_OK;
}

int nts_graph_edges(nts_graph* g, uint64_t* n_edges)
{
    if (!g || !n_edges) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    int rc = build_edges(g);
    if (rc) return rc;
    *n_edges = g->E;
    return NTS_OK;
}

int nts_graph_download_edges(nts_graph* g, uint32_t* u, uint32_t* v, uint32_t* support)
{
    if (!g) return fail(NTS_ERR_ARG, "null argument");
    NTS_CUDA(cudaSetDevice(g->ctx->device));
    int rc = build_edges(g);
    if (rc) return rc;
    if (!g->E) return NTS_OK;
    cudaStream_t st = g->ctx->stream;
    if (u) NTS_CUDA(cudaMemcpyAsync(u, g->e_u.p, g->E * 4, cudaMemcpyDeviceToHost, st));
    if (v) NTS_CUDA(cudaMemcpyAsync(v, g->e_v.p, g->E * 4, cudaMemcpyDeviceToHost, st));
    if (support) NTS_CUDA(cudaMemcpyAsync(support, g->e_support.p, g->E * 4, cudaMemcpyDeviceToHost, st));
    NTS_CUDA(cudaStreamSynchronize(st));
    return NTS_OK;
}

}  // extern "C"
