// nts_hostgraph.cu -- native host-side pieces of the graph stage (no device code).
//
// The graph stage keeps its bulk on the device (nts_graph.cu) and its few irregular places on the host
// (ntsynt_b200/synteny.py).  Two of those places walk thousands of small objects one by one and dominated the
// Python time at 3 Gbp; they live here, behind the same C-ABI, operating on the host arrays the engine owns:
//   nts_host_walk_paths   find_paths over the components that contain a non-(i, i+1) edge
//                         (subprojects/ntJoin/bin/ntjoin.py:114-151; SyntenyEngine._find_paths)
//   nts_host_simplify     run_graph_simplification on the round-0 graph, candidates = degree-3 vertices
//                         (bin/ntsynt_synteny.py:548-590; SyntenyEngine._simplify_round0)
// Both reproduce the reference's visiting ORDER, which decides bit-exactness.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "nts_internal.h"

using namespace nts;

namespace {

struct Runs {
    const int64_t* starts; const int64_t* ends; int64_t n; int64_t V0;
    // index of the run holding base vertex v
    int64_t find(int64_t v) const { return (std::upper_bound(starts, starts + n, v) - starts) - 1; }
};

}  // namespace

extern "C" {

/* Paths of the components that hold at least one "sparse" vertex (a vertex with a neighbour that is not its id +- 1,
 * or a vertex added after round 0).  nbr[v] = the (at most two) neighbours of v or -1; runs = maximal chains of
 * (i, i+1) edges over base vertices [0, V0) as sorted (starts, ends); sv = the sparse vertices, ascending; opos =
 * position of every vertex in the orienting assembly.  A component that is a simple path with two distinct ends
 * becomes one path, listed from the end with the smaller position, as segments (lo, hi, dir) over runs.
 * Output: seg_lo/seg_hi/seg_dir [<= 2*n_sv + 2], path_off [<= n_paths + 1]; returns through n_paths / n_segs. */
static int walk_paths_core(const int32_t* nbr, int64_t V0, const int64_t* starts, const int64_t* ends, int64_t n_runs,
                           const int64_t* sv, int64_t n_sv, const int64_t* opos, const int64_t* opos_ids, int64_t n_opos,
                           int64_t* seg_lo, int64_t* seg_hi, int8_t* seg_dir, int64_t* path_off, int64_t seg_cap,
                           int64_t* n_paths, int64_t* n_segs)
{
    if (!nbr || !sv || !opos || !seg_lo || !seg_hi || !seg_dir || !path_off || !n_paths || !n_segs)
        return fail(NTS_ERR_ARG, "null argument");
    // position in the orienting assembly: dense (indexed by vertex id) or sparse (sorted ids + values)
    auto opos_at = [&](int64_t v) -> int64_t {
        if (!opos_ids) return opos[v];
        const int64_t* p = std::lower_bound(opos_ids, opos_ids + n_opos, v);
        return (p != opos_ids + n_opos && *p == v) ? opos[p - opos_ids] : INT64_MIN;
    };
    const Runs runs{starts, ends, n_runs, V0};
    auto bounds = [&](int64_t v, int64_t* a, int64_t* b) {
        if (v >= V0) { *a = *b = v; return; }
        const int64_t r = runs.find(v);
        if (r < 0 || ends[r] < v) { *a = *b = v; return; }      // (the run list may leave out single-vertex runs)
        *a = starts[r]; *b = ends[r];
    };
    // ends of the runs that hold a sparse vertex, ascending and distinct
    std::vector<int64_t> ends_all;
    ends_all.reserve((size_t)n_sv * 2);
    for (int64_t i = 0; i < n_sv; ++i) {
        int64_t a, b;
        bounds(sv[i], &a, &b);
        ends_all.push_back(a); ends_all.push_back(b);
    }
    std::sort(ends_all.begin(), ends_all.end());
    ends_all.erase(std::unique(ends_all.begin(), ends_all.end()), ends_all.end());
    std::unordered_set<int64_t> seen;                 // run starts already walked
    seen.reserve(ends_all.size() * 2 + 16);
    int64_t np = 0, ns = 0;
    path_off[0] = 0;
    std::vector<int64_t> lo, hi;
    std::vector<int8_t> dir;
    for (int64_t e0 : ends_all) {
        const int deg = (nbr[2 * e0] >= 0) + (nbr[2 * e0 + 1] >= 0);
        if (deg != 1) continue;
        int64_t a, b;
        bounds(e0, &a, &b);
        if (seen.count(a)) continue;
        lo.clear(); hi.clear(); dir.clear();
        int64_t prev = -1, cur = e0, last = -1;
        bool ok = true;
        for (;;) {
            bounds(cur, &a, &b);
            if (seen.count(a)) { ok = false; break; }
            seen.insert(a);
            if (cur == a) { lo.push_back(a); hi.push_back(b); dir.push_back(1); last = b; }
            else          { lo.push_back(a); hi.push_back(b); dir.push_back(-1); last = a; }
            int64_t inside = -1;
            if (b > a) inside = (last == b) ? last - 1 : last + 1;
            int64_t nxt = -1;
            for (int s = 0; s < 2; ++s) {
                const int64_t y = nbr[2 * last + s];
                if (y < 0 || y == inside) continue;
                if (a == b && y == prev) continue;
                nxt = y;
            }
            if (nxt < 0) break;
            prev = last; cur = nxt;
        }
        if (!ok) continue;
        const int64_t first = dir.front() > 0 ? lo.front() : hi.front();
        const int64_t lastv = dir.back() > 0 ? hi.back() : lo.back();
        if (first == lastv) continue;
        const int64_t pa = opos_at(first), pb = opos_at(lastv);
        if (pa == INT64_MIN || pb == INT64_MIN) return fail(NTS_ERR_ARG, "position of a path end was not supplied");
        if (pa == pb) continue;
        const size_t n = lo.size();
        if (ns + (int64_t)n > seg_cap) return fail(NTS_ERR_OVERFLOW, "segment buffer too small");
        if (pa < pb) {
            for (size_t i = 0; i < n; ++i) { seg_lo[ns] = lo[i]; seg_hi[ns] = hi[i]; seg_dir[ns] = dir[i]; ++ns; }
        } else {
            for (size_t i = n; i-- > 0;) { seg_lo[ns] = lo[i]; seg_hi[ns] = hi[i]; seg_dir[ns] = (int8_t)-dir[i]; ++ns; }
        }
        path_off[++np] = ns;
    }
    *n_paths = np; *n_segs = ns;
    return NTS_OK;
}

int nts_host_walk_paths(const int32_t* nbr, int64_t V0, const int64_t* starts, const int64_t* ends, int64_t n_runs,
                        const int64_t* sv, int64_t n_sv, const int64_t* opos, int64_t* seg_lo, int64_t* seg_hi,
                        int8_t* seg_dir, int64_t* path_off, int64_t seg_cap, int64_t* n_paths, int64_t* n_segs)
{
    return walk_paths_core(nbr, V0, starts, ends, n_runs, sv, n_sv, opos, nullptr, 0, seg_lo, seg_hi, seg_dir, path_off, seg_cap,
                           n_paths, n_segs);
}

/* the same with the orienting positions given only where a path can end: opos_ids (ascending) = the sparse vertices and
 * the ends of the runs that hold them, opos_vals their positions */
int nts_host_walk_paths_sparse(const int32_t* nbr, int64_t V0, const int64_t* starts, const int64_t* ends, int64_t n_runs,
                               const int64_t* sv, int64_t n_sv, const int64_t* opos_ids, const int64_t* opos_vals, int64_t n_opos,
                               int64_t* seg_lo, int64_t* seg_hi, int8_t* seg_dir, int64_t* path_off, int64_t seg_cap,
                               int64_t* n_paths, int64_t* n_segs)
{
    if (!opos_ids) return fail(NTS_ERR_ARG, "null argument");
    return walk_paths_core(nbr, V0, starts, ends, n_runs, sv, n_sv, opos_vals, opos_ids, n_opos, seg_lo, seg_hi, seg_dir, path_off,
                           seg_cap, n_paths, n_segs);
}

/* find_synteny_blocks + check_for_indels + filter_synteny_blocks for paths given as segments (bin/ntsynt_synteny.py:
 * 66-106, 364-426; bin/synteny_block.py:48-65) -- the host-walked paths of a round in one call.
 *   path p owns segments [path_off[p], path_off[p+1]) = runs lo..hi of consecutive ids traversed in direction dir;
 *   up / down [G x n_seg]: pairs (j, j+1), lo <= j < hi, whose position increases / decreases in each assembly;
 *   (ids ascending, pos / ctg [G x n_ids]): positions and contigs of every vertex the call looks at -- both ends of
 *   every segment and c, c + 1 for every large-spread pair c inside a segment;  big: those pairs, ascending.
 * A path's block = its segments from the last junction with a contig change on (the reference never flags the earlier
 * part); per assembly '+' / '-' if every step increases / decreases, else by the m_pct rule, else the block's segments
 * are deleted.  The block is then cut at every junction / inner pair whose |dpos| spread exceeds bp (that edge is
 * removed), and pieces with fewer than min_mx minimizers are deleted.
 * Outputs: blocks (b_off into the out segment arrays o_lo/o_hi/o_dir, b_n, b_first, b_last, b_ori / b_ctg / b_fpos /
 * b_lpos [n_blocks x G]); deleted segments (r_lo, r_hi); removed edges (e_u, e_v).  cap bounds every output array
 * (n_seg + number of large-spread pairs inside segments + 1 is enough).  counts[4] = blocks, out segments, deleted
 * segments, removed edges. */
int nts_host_paths_to_blocks(int64_t n_paths, const int64_t* path_off, const int64_t* seg_lo, const int64_t* seg_hi,
                             const int8_t* seg_dir, uint32_t G, const int64_t* up, const int64_t* down, const int64_t* ids,
                             int64_t n_ids, const int64_t* pos, const int32_t* ctg, const int64_t* big, int64_t n_big, int64_t bp,
                             double m_pct, int64_t min_mx, int64_t cap, int64_t* b_off, int64_t* b_n, int64_t* b_first,
                             int64_t* b_last, int8_t* b_ori, int32_t* b_ctg, int64_t* b_fpos, int64_t* b_lpos, int64_t* o_lo,
                             int64_t* o_hi, int8_t* o_dir, int64_t* r_lo, int64_t* r_hi, int64_t* e_u, int64_t* e_v,
                             int64_t counts[4])
{
    if (!path_off || !counts || (n_paths && (!seg_lo || !seg_hi || !seg_dir || !up || !down || !ids || !pos || !ctg)))
        return fail(NTS_ERR_ARG, "null argument");
    if (G < 1 || G > 32) return fail(NTS_ERR_ARG, "between 1 and 32 assemblies are supported");
    const int64_t n_seg = n_paths ? path_off[n_paths] : 0;
    int64_t nb = 0, no = 0, nr = 0, ne = 0;
    bool missing = false;
    auto at = [&](int64_t v) -> int64_t {
        const int64_t* p = std::lower_bound(ids, ids + n_ids, v);
        if (p == ids + n_ids || *p != v) { missing = true; return 0; }
        return (int64_t)(p - ids);
    };
    auto first_of = [&](int64_t j) { return seg_dir[j] > 0 ? seg_lo[j] : seg_hi[j]; };
    auto last_of = [&](int64_t j) { return seg_dir[j] > 0 ? seg_hi[j] : seg_lo[j]; };
    auto spread_gt = [&](int64_t u, int64_t v) {          // max |dpos| - min |dpos| over the assemblies > bp
        const int64_t iu = at(u), iv = at(v);
        int64_t mx = 0, mn = INT64_MAX;
        for (uint32_t a = 0; a < G; ++a) {
            int64_t d = pos[a * n_ids + iv] - pos[a * n_ids + iu];
            if (d < 0) d = -d;
            mx = std::max(mx, d); mn = std::min(mn, d);
        }
        return mx - mn > bp;
    };
    std::vector<int64_t> inc(G), dec(G);
    std::vector<int8_t> ori(G);
    struct Seg { int64_t lo, hi; int8_t d; };
    std::vector<Seg> cur;
    std::vector<std::vector<Seg>> pieces;
    for (int64_t p = 0; p < n_paths; ++p) {
        const int64_t s0 = path_off[p], s1 = path_off[p + 1];
        if (s1 <= s0) continue;
        // the block starts after the last junction where some assembly changes contig
        int64_t start = s0;
        for (int64_t j = s0 + 1; j < s1; ++j) {
            const int64_t iu = at(last_of(j - 1)), iv = at(first_of(j));
            bool chg = false;
            for (uint32_t a = 0; a < G && !chg; ++a) chg = ctg[a * n_ids + iu] != ctg[a * n_ids + iv];
            if (chg) start = j;
        }
        int64_t n = 0;
        std::fill(inc.begin(), inc.end(), 0); std::fill(dec.begin(), dec.end(), 0);
        for (int64_t j = start; j < s1; ++j) {
            n += seg_hi[j] - seg_lo[j] + 1;
            for (uint32_t a = 0; a < G; ++a) {
                const int64_t u_ = up[a * n_seg + j], d_ = down[a * n_seg + j];
                inc[a] += seg_dir[j] > 0 ? u_ : d_;
                dec[a] += seg_dir[j] > 0 ? d_ : u_;
            }
            if (j > start) {
                const int64_t iu = at(last_of(j - 1)), iv = at(first_of(j));
                for (uint32_t a = 0; a < G; ++a) {
                    const int64_t d = pos[a * n_ids + iv] - pos[a * n_ids + iu];
                    if (d > 0) ++inc[a]; else if (d < 0) ++dec[a];
                }
            }
        }
        bool unoriented = false;
        for (uint32_t a = 0; a < G; ++a) {
            if (inc[a] == n - 1 || n == 1) ori[a] = '+';
            else if (dec[a] == n - 1) ori[a] = '-';
            else {
                const double positive = (double)inc[a] / (double)(n - 1) * 100;
                const double negative = 100 - positive;
                ori[a] = positive >= m_pct ? '+' : (negative >= m_pct ? '-' : '?');
                if (ori[a] == '?') unoriented = true;
            }
        }
        if (unoriented) {
            for (int64_t j = start; j < s1; ++j) {
                if (nr >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                r_lo[nr] = seg_lo[j]; r_hi[nr] = seg_hi[j]; ++nr;
            }
            continue;
        }
        const int64_t i_f = at(first_of(start));
        // indel cuts
        pieces.clear(); cur.clear();
        for (int64_t j = start; j < s1; ++j) {
            const int64_t slo = seg_lo[j], shi = seg_hi[j];
            const int8_t d = seg_dir[j];
            if (j > start) {
                const int64_t u = last_of(j - 1), v = first_of(j);
                if (spread_gt(u, v)) {
                    if (ne >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    e_u[ne] = u; e_v[ne] = v; ++ne;
                    pieces.push_back(cur); cur.clear();
                }
            }
            const int64_t* c0 = shi > slo ? std::lower_bound(big, big + n_big, slo) : big;
            const int64_t* c1 = shi > slo ? std::lower_bound(big, big + n_big, shi) : big;
            if (c1 > c0) {
                for (const int64_t* c = c0; c < c1; ++c) {
                    if (ne >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    e_u[ne] = *c; e_v[ne] = *c + 1; ++ne;
                }
                if (d > 0) {
                    int64_t a0 = slo;
                    for (const int64_t* c = c0; c < c1; ++c) { cur.push_back({a0, *c, 1}); pieces.push_back(cur); cur.clear(); a0 = *c + 1; }
                    cur.push_back({a0, shi, 1});
                } else {
                    int64_t a0 = shi;
                    for (const int64_t* c = c1; c-- > c0;) { cur.push_back({*c + 1, a0, -1}); pieces.push_back(cur); cur.clear(); a0 = *c; }
                    cur.push_back({slo, a0, -1});
                }
            } else {
                cur.push_back({slo, shi, d});
            }
        }
        pieces.push_back(cur);
        for (const auto& pc : pieces) {
            int64_t pn = 0;
            for (const Seg& sg : pc) pn += sg.hi - sg.lo + 1;
            if (pn < min_mx) {
                for (const Seg& sg : pc) {
                    if (nr >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    r_lo[nr] = sg.lo; r_hi[nr] = sg.hi; ++nr;
                }
                continue;
            }
            if (nb >= cap || no + (int64_t)pc.size() > cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
            const int64_t f = pc.front().d > 0 ? pc.front().lo : pc.front().hi;
            const int64_t l = pc.back().d > 0 ? pc.back().hi : pc.back().lo;
            const int64_t jf = at(f), jl = at(l);
            b_off[nb] = no; b_n[nb] = pn; b_first[nb] = f; b_last[nb] = l;
            for (uint32_t a = 0; a < G; ++a) {
                b_ori[nb * G + a] = ori[a];
                b_ctg[nb * G + a] = ctg[a * n_ids + i_f];             // the contigs of the path's block, also for its pieces
                b_fpos[nb * G + a] = pos[a * n_ids + jf];
                b_lpos[nb * G + a] = pos[a * n_ids + jl];
            }
            for (const Seg& sg : pc) { o_lo[no] = sg.lo; o_hi[no] = sg.hi; o_dir[no] = sg.d; ++no; }
            ++nb;
        }
    }
    if (missing) return fail(NTS_ERR_ARG, "paths_to_blocks: a vertex the walk needs is not in the position table");
    b_off[nb] = no;
    counts[0] = nb; counts[1] = no; counts[2] = nr; counts[3] = ne;
    return NTS_OK;
}

/* run_graph_simplification on the round-0 graph (bin/ntsynt_synteny.py:566-590).  cand = the vertices with exactly
 * three distinct neighbours, ascending; rank / inv = [G x V] rank of every vertex in every assembly's filtered list
 * and its inverse; ctg = [G x ctg_stride] contig of every vertex.  Candidate edges (both ends candidates) are
 * visited in build_graph's edge-id order (ntjoin_utils.py:97-115); an edge whose two ends each have exactly one
 * incident full-weight edge and which closes exactly one triangle removes the triangle's third vertex and becomes
 * full weight itself (visible to the later edges).  Output (same length, <= n_cand * 4): bump_s < bump_t, removed. */
static int simplify_core(const int64_t* cand, int64_t n_cand, const int64_t* left, const int64_t* right, const int64_t* rk,
                         uint32_t G, int64_t* bump_s, int64_t* bump_t, int64_t* removed, int64_t out_cap, int64_t* n_out)
{
    const size_t n = (size_t)n_cand;
    auto pos_of = [&](int64_t v) -> int64_t {
        const int64_t* p = std::lower_bound(cand, cand + n_cand, v);
        return (p != cand + n_cand && *p == v) ? (int64_t)(p - cand) : -1;
    };
    // distinct neighbours with their weights (number of supporting assemblies)
    struct Nb { int64_t x[4]; int w[4]; int n = 0; };
    std::vector<Nb> nb(n);
    for (size_t i = 0; i < n; ++i) {
        Nb& d = nb[i];
        for (uint32_t a = 0; a < G; ++a)
            for (int side = 0; side < 2; ++side) {
                const int64_t x = side ? right[i * G + a] : left[i * G + a];
                if (x < 0) continue;
                int k = 0;
                while (k < d.n && d.x[k] != x) ++k;
                if (k == d.n) { if (d.n == 4) return fail(NTS_ERR_STATE, "candidate with more than 3 neighbours"); d.x[d.n] = x; d.w[d.n] = 0; ++d.n; }
                ++d.w[k];
            }
    }
    auto adjacent = [&](uint32_t a, size_t i, int64_t dst) { return left[i * G + a] == dst || right[i * G + a] == dst; };
    struct Edge { int64_t s, t; int64_t key[4]; };        // key = (sigma.assembly, sigma.rank, tau.assembly, tau.rank)
    std::vector<Edge> edges;
    for (size_t i = 0; i < n; ++i) {
        const int64_t u = cand[i];
        for (int k = 0; k < nb[i].n; ++k) {
            const int64_t x = nb[i].x[k];
            if (x <= u) continue;                          // each unordered pair once, from its smaller end
            const int64_t ix = pos_of(x);
            if (ix < 0) continue;
            Edge e; e.s = u; e.t = x;
            uint32_t a0 = 0;
            while (a0 < G && !adjacent(a0, i, x)) ++a0;
            const int64_t ru = rk[i * G + a0], rv = rk[(size_t)ix * G + a0];
            const size_t isrc = ru < rv ? i : (size_t)ix;
            e.key[2] = a0; e.key[3] = std::min(ru, rv);
            e.key[0] = e.key[1] = -1;
            for (uint32_t a = 0; a < G; ++a) {             // first time src is the left element of a NEW pair
                const int64_t y = right[isrc * G + a];
                if (y < 0) continue;
                bool earlier = false;
                for (uint32_t b = 0; b < a && !earlier; ++b) earlier = adjacent(b, isrc, y);
                if (!earlier) { e.key[0] = a; e.key[1] = rk[isrc * G + a]; break; }
            }
            edges.push_back(e);
        }
    }
    std::sort(edges.begin(), edges.end(), [](const Edge& p, const Edge& q) {
        for (int i = 0; i < 4; ++i) if (p.key[i] != q.key[i]) return p.key[i] < q.key[i];
        return p.s != q.s ? p.s < q.s : p.t < q.t;
    });
    std::vector<int> fullc(n);
    for (size_t i = 0; i < n; ++i) { int c = 0; for (int k = 0; k < nb[i].n; ++k) c += nb[i].w[k] == (int)G; fullc[i] = c; }
    int64_t no = 0;
    for (const Edge& e : edges) {
        const size_t is = (size_t)pos_of(e.s), it = (size_t)pos_of(e.t);
        if (fullc[is] != 1 || fullc[it] != 1) continue;    // node_partially_anchored on both ends
        int64_t common = -1; int n_common = 0;
        for (int k = 0; k < nb[is].n; ++k) {
            const int64_t x = nb[is].x[k];
            if (x == e.t) continue;
            for (int q = 0; q < nb[it].n; ++q) if (nb[it].x[q] == x) { common = x; ++n_common; }
        }
        if (n_common != 1) continue;                        // the edge itself + exactly one 2-step path
        if (no >= out_cap) return fail(NTS_ERR_OVERFLOW, "simplification output buffer too small");
        bump_s[no] = e.s; bump_t[no] = e.t; removed[no] = common; ++no;
        int ks = 0; while (nb[is].x[ks] != e.t) ++ks;
        if (nb[is].w[ks] != (int)G) {
            nb[is].w[ks] = (int)G;
            int kt = 0; while (nb[it].x[kt] != e.s) ++kt;
            nb[it].w[kt] = (int)G;
            ++fullc[is]; ++fullc[it];
        }
    }
    *n_out = no;
    return NTS_OK;
}

int nts_host_simplify(const int64_t* cand, int64_t n_cand, const uint32_t* rank, const uint32_t* inv, const int32_t* ctg,
                      int64_t ctg_stride, int64_t V, uint32_t G, int64_t* bump_s, int64_t* bump_t, int64_t* removed,
                      int64_t out_cap, int64_t* n_out)
{
    if (!cand || !rank || !inv || !ctg || !bump_s || !bump_t || !removed || !n_out) return fail(NTS_ERR_ARG, "null argument");
    if (G < 1 || G > 32) return fail(NTS_ERR_ARG, "between 1 and 32 assemblies are supported");
    const size_t n = (size_t)n_cand;
    // neighbourhoods: left / right neighbour of every candidate in every assembly (or -1)
    std::vector<int64_t> left(n * G), right(n * G), rk(n * G);
    for (size_t i = 0; i < n; ++i) {
        const int64_t u = cand[i];
        for (uint32_t a = 0; a < G; ++a) {
            const int64_t r = rank[(size_t)a * V + u];
            int64_t lf = -1, rt = -1;
            if (r > 0) { const int64_t x = inv[(size_t)a * V + r - 1]; if (ctg[a * ctg_stride + x] == ctg[a * ctg_stride + u]) lf = x; }
            if (r + 1 < V) { const int64_t x = inv[(size_t)a * V + r + 1]; if (ctg[a * ctg_stride + x] == ctg[a * ctg_stride + u]) rt = x; }
            left[i * G + a] = lf; right[i * G + a] = rt; rk[i * G + a] = r;
        }
    }
    return simplify_core(cand, n_cand, left.data(), right.data(), rk.data(), G, bump_s, bump_t, removed, out_cap, n_out);
}

/* the same with the candidates' neighbourhoods already extracted on the device (nts_graph_neigh): left / right / rk are
 * [n_cand x G] */
int nts_host_simplify_neigh(const int64_t* cand, int64_t n_cand, const int64_t* left, const int64_t* right, const int64_t* rk,
                            uint32_t G, int64_t* bump_s, int64_t* bump_t, int64_t* removed, int64_t out_cap, int64_t* n_out)
{
    if (!cand || !left || !right || !rk || !bump_s || !bump_t || !removed || !n_out) return fail(NTS_ERR_ARG, "null argument");
    if (G < 1 || G > 32) return fail(NTS_ERR_ARG, "between 1 and 32 assemblies are supported");
    return simplify_core(cand, n_cand, left, right, rk, G, bump_s, bump_t, removed, out_cap, n_out);
}

}  // extern "C"
