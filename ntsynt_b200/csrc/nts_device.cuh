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
    // [in*4+out] = { f.lo, f.hi, r.lo, r.hi }: f = seed[in] ^ srol^k(seed[out]), r = seed[3-out] ^ srol^k(seed[3-in]);
    // one 16-byte entry per (in,out) pair so that a roll step costs a single LDS.128
    uint4 roll[16];
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

// the same, searched only among islands [lo, hi] (lo <= answer <= hi): a tile's threads share the two ends of the tile
__device__ __forceinline__ uint32_t find_island_in(const GenomeView& g, uint64_t v, uint32_t lo, uint32_t hi)
{
    ++hi;                           // invariant: seg_v[lo] <= v < seg_v[hi]
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

// split rotates on a 64-bit value held as two 32-bit registers (the SM's datapath is 32 bits wide)
__device__ __forceinline__ void srol32(uint32_t& lo, uint32_t& hi)
{
    const uint32_t f = __funnelshift_l(lo, hi, 1);       // (hi << 1) | (lo >> 31)
    const uint32_t nlo = (lo << 1) | (hi & 1u);           // bit 32 -> bit 0
    hi = (f & ~2u) | ((hi >> 30) & 2u);                   // bit 63 -> bit 33
    lo = nlo;
}

__device__ __forceinline__ void sror32(uint32_t& lo, uint32_t& hi)
{
    const uint32_t nlo = __funnelshift_r(lo, hi, 1);      // (lo >> 1) | (hi << 31)
    hi = ((hi >> 1) & ~1u) | (lo & 1u) | ((hi & 2u) << 30);   // bit 0 -> bit 32, bit 33 -> bit 63
    lo = nlo;
}

// Rolling ntHash2 over a run of consecutive valid indices [v, v + count).  For every k-mer calls
// f(j, h0, base) with j = 0..count-1 (run-local index) and base = global base index of the k-mer.
// Re-seeds itself at island boundaries.  `tabs` must point to shared memory.
//
// Inner loop: the leaving and the entering base streams are read as unaligned 32-bit groups (16 bases
// each, one funnel shift per group), merged into two words whose nibbles are the (in,out) table indices
// of the even / odd steps, and 16 roll steps are unrolled with static shifts: per k-mer one LDS.128,
// two 5-instruction split rotates, four XORs and a 64-bit add.
// isl_lo / isl_hi: islands known to bracket the run's first k-mer (a tile looks its two ends up once; most tiles lie
// inside one island, and their threads then start without any search).
template <typename F>
__device__ __forceinline__ void hash_run(const GenomeView& g, const HashTables* tabs, uint64_t v, uint32_t count, F&& f,
                                         uint32_t isl_lo = 0, uint32_t isl_hi = 0xFFFFFFFFu)
{
    if (count == 0) return;
    const uint32_t k = g.k;
    const uint32_t* __restrict__ p32 = reinterpret_cast<const uint32_t*>(g.packed);
    const char* roll = reinterpret_cast<const char*>(tabs->roll);
    uint32_t s = isl_hi == 0xFFFFFFFFu ? find_island(g, v) : find_island_in(g, v, isl_lo, isl_hi);
    uint32_t j = 0;
    while (j < count) {
        const uint64_t sv0 = __ldg(&g.seg_v[s]);
        const uint64_t sv1 = __ldg(&g.seg_v[s + 1]);
        const uint64_t b = __ldg(&g.seg_base[s]) + (v + j - sv0);   // base index of the piece's first k-mer
        const uint32_t n_here = (uint32_t)min((uint64_t)(count - j), sv1 - (v + j));
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
        if (n_here > 1) {
            uint32_t flo = (uint32_t)fwd, fhi = (uint32_t)(fwd >> 32), rlo = (uint32_t)rev, rhi = (uint32_t)(rev >> 32);
            // step t (1-based) drops base b + t - 1 and takes in base b + t - 1 + k
            const uint64_t ob = b * 2, ib = (b + k) * 2;
            uint32_t ow = (uint32_t)(ob >> 5), iw = (uint32_t)(ib >> 5);
            const uint32_t os = (uint32_t)ob & 31u, is = (uint32_t)ib & 31u;
            uint32_t o_lo = __ldg(p32 + ow), i_lo = __ldg(p32 + iw);
            auto step = [&](uint32_t nib) {
                const uint4 e = *reinterpret_cast<const uint4*>(roll + nib);     // nib = table index * 16
                srol32(flo, fhi);
                flo ^= e.x; fhi ^= e.y;
                rlo ^= e.z; rhi ^= e.w;
                sror32(rlo, rhi);
                uint32_t hl, hh;                                                  // h0 = fwd + rev
                asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(hl), "=r"(hh) : "r"(flo), "r"(fhi), "r"(rlo), "r"(rhi));
                return ((uint64_t)hh << 32) | hl;
            };
            for (uint32_t t = 1; t < n_here; t += 16) {
                const uint32_t o_hi = __ldg(p32 + ow + 1), i_hi = __ldg(p32 + iw + 1);
                const uint32_t O = __funnelshift_r(o_lo, o_hi, os), I = __funnelshift_r(i_lo, i_hi, is);
                o_lo = o_hi; i_lo = i_hi; ++ow; ++iw;
                const uint32_t E = ((I << 2) & 0xCCCCCCCCu) | (O & 0x33333333u);     // nibble q: step 2q
                const uint32_t D = (I & 0xCCCCCCCCu) | ((O >> 2) & 0x33333333u);     // nibble q: step 2q + 1
                const uint32_t left = n_here - t;
                if (left >= 15) {
                    // (15: a run of 16 k-mers = the seed + 15 steps, the shape of the Bloom insert's tiles)
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (u == 15 && left == 15) break;
                        const uint32_t src = (u & 1) ? D : E;
                        const int q = u >> 1;
                        const uint32_t nib = (q == 0 ? (src << 4) : (src >> (4 * q - 4))) & 0xF0u;
                        const uint64_t h = step(nib);
                        f(j + t + u, h, b + t + u);
                    }
                } else {
                    for (uint32_t u = 0; u < left; ++u) {
                        const uint32_t src = (u & 1) ? D : E;
                        const uint32_t nib = ((src >> (4 * (u >> 1))) & 15u) << 4;
                        const uint64_t h = step(nib);
                        f(j + t + u, h, b + t + u);
                    }
                }
            }
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
