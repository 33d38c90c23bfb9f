// nts_device.cuh -- device-side building blocks shared by the kernels of libntsynt_b200.
//
// ntHash2 as used by btllib >= 1.6 (the reference calls it at
// src/ntsynt_make_common_bf.cpp:147-148 and inside indexlr): canonical h0 = fwd + rev over a
// split-rotate (33 + 31 bit) rolling hash; h1 = extension of h0.  Written from the published
// definition; not a translation of btllib's table-driven code:
// here a thread owns a run of consecutive valid k-mers of a 2-bit packed contig, seeds itself
// from per-position XOR tables in shared memory and then rolls one base per step with two
// 16-entry (in,out)-indexed XOR tables.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nts {

constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ull;
constexpr uint64_t SEED_C = 0x3193c18562a02b4cull;
constexpr uint64_t SEED_G = 0x20323ed082572324ull;
constexpr uint64_t SEED_T = 0x295549f54be24456ull;
constexpr uint64_t MULTISEED = 0x90b45d39fb6da1faull;
constexpr int MULTISHIFT = 27;
constexpr uint64_t KEY_MAX = 0xFFFFFFFFFFFFFFFFull;
constexpr int MAX_K = 64;

__host__ __device__ __forceinline__ uint64_t seed_of(unsigned c)
{
    return c == 0 ? SEED_A : c == 1 ? SEED_C : c == 2 ? SEED_G : SEED_T;
}

// split rotate left by one: bits 0..32 rotate within 33 bits, bits 33..63 within 31 bits
__host__ __device__ __forceinline__ uint64_t srol(uint64_t x)
{
    uint64_t m = ((x & 0x8000000000000000ull) >> 30) | ((x & 0x100000000ull) >> 32);
    return ((x << 1) & 0xFFFFFFFDFFFFFFFFull) | m;
}

__host__ __device__ __forceinline__ uint64_t sror(uint64_t x)
{
    uint64_t m = ((x & 0x200000000ull) << 30) | ((x & 1ull) << 32);
    return ((x >> 1) & 0xFFFFFFFEFFFFFFFFull) | m;
}

__host__ __device__ __forceinline__ uint64_t ext_hash(uint64_t h0, unsigned i, unsigned k)
{
    uint64_t t = h0 * ((uint64_t)i ^ ((uint64_t)k * MULTISEED));
    return t ^ (t >> MULTISHIFT);
}

// Hash constants for one k, built on the host (nts_api.cu: make_hash_tables) and staged into
// shared memory by every kernel that hashes.
struct HashTables {
    uint64_t roll_f[16];          // [in*4+out] = seed[in] ^ srol^k(seed[out])
    uint64_t roll_r[16];          // [in*4+out] = seed[3-out] ^ srol^k(seed[3-in])
    uint64_t init_f[MAX_K * 4];   // [i*4+c] = srol^(k-1-i)(seed[c])
    uint64_t init_r[MAX_K * 4];   // [i*4+c] = srol^i(seed[3-c])
};

// Device view of a genome prepared for one (k, mask): the valid k-mers of all contigs laid end
// to end in "valid index" space.  Island s covers valid indices [seg_v[s], seg_v[s+1]) and its
// first k-mer starts at global base index seg_base[s] of the packed array.
struct GenomeView {
    const uint64_t* __restrict__ packed;
    const uint64_t* __restrict__ seg_v;     // [n_seg + 1]
    const uint64_t* __restrict__ seg_base;  // [n_seg]
    uint32_t n_seg;
    uint32_t k;
};

// exact h mod m for m < 2^63 via one 64x64->128 multiply; mprime = floor((2^64-1)/m)
__device__ __forceinline__ uint64_t fast_mod(uint64_t h, uint64_t m, uint64_t mprime)
{
    uint64_t q = __umul64hi(h, mprime);
    uint64_t r = h - q * m;
    while (r >= m) r -= m;
    return r;
}

// largest s with seg_v[s] <= v   (v < seg_v[n_seg])
__device__ __forceinline__ uint32_t find_island(const GenomeView& g, uint64_t v)
{
    uint32_t lo = 0, hi = g.n_seg;  // invariant: seg_v[lo] <= v < seg_v[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&g.seg_v[mid]) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned base_at(const uint64_t* __restrict__ packed, uint64_t b)
{
    return (unsigned)((__ldg(&packed[b >> 5]) >> ((b & 31) * 2)) & 3ull);
}

// Rolling ntHash2 over a run of consecutive valid indices [v, v + count).  For every k-mer calls
// f(j, h0, base) with j = 0..count-1 (run-local index) and base = global base index of the k-mer.
// Re-seeds itself at island boundaries.  `tabs` must point to shared memory.
template <typename F>
__device__ __forceinline__ void hash_run(const GenomeView& g, const HashTables* tabs, uint64_t v, uint32_t count, F&& f)
{
    if (count == 0) return;
    const uint32_t k = g.k;
    uint32_t s = find_island(g, v);
    uint32_t j = 0;
    while (j < count) {
        const uint64_t sv0 = __ldg(&g.seg_v[s]);
        const uint64_t sv1 = __ldg(&g.seg_v[s + 1]);
        uint64_t b = __ldg(&g.seg_base[s]) + (v + j - sv0);   // base index of current k-mer
        uint32_t n_here = (uint32_t)min((uint64_t)(count - j), sv1 - (v + j));
        // seed: XOR of per-position tables over the k bases of the first k-mer
        uint64_t fwd = 0, rev = 0;
        {
            uint64_t wi = b >> 5;
            uint32_t sh = (uint32_t)(b & 31) * 2;
            uint64_t word = __ldg(&g.packed[wi]);
            for (uint32_t i = 0; i < k; ++i) {
                unsigned c = (unsigned)(word >> sh) & 3u;
                fwd ^= tabs->init_f[i * 4 + c];
                rev ^= tabs->init_r[i * 4 + c];
                sh += 2;
                if (sh == 64) { sh = 0; word = __ldg(&g.packed[++wi]); }
            }
        }
        f(j, fwd + rev, b);
        // cursors: `out` = base leaving (b), `in` = base entering (b + k)
        uint64_t wo = b >> 5, wn = (b + k) >> 5;
        uint32_t so = (uint32_t)(b & 31) * 2, sn = (uint32_t)((b + k) & 31) * 2;
        uint64_t word_o = __ldg(&g.packed[wo]);
        uint64_t word_n = (n_here > 1) ? __ldg(&g.packed[wn]) : 0;
        for (uint32_t t = 1; t < n_here; ++t) {
            unsigned cout = (unsigned)(word_o >> so) & 3u;
            unsigned cin = (unsigned)(word_n >> sn) & 3u;
            unsigned idx = cin * 4 + cout;
            fwd = srol(fwd) ^ tabs->roll_f[idx];
            rev = sror(rev ^ tabs->roll_r[idx]);
            ++b;
            f(j + t, fwd + rev, b);
            so += 2; sn += 2;
            if (so == 64) { so = 0; word_o = __ldg(&g.packed[++wo]); }
            if (sn == 64) { sn = 0; if (t + 1 < n_here) word_n = __ldg(&g.packed[++wn]); }
        }
        j += n_here;
        ++s;
    }
}

// valid index -> global base index (used only for the few selected minimizers)
__device__ __forceinline__ uint64_t valid_to_base(const GenomeView& g, uint64_t v)
{
    uint32_t s = find_island(g, v);
    return __ldg(&g.seg_base[s]) + (v - __ldg(&g.seg_v[s]));
}

}  // namespace nts
