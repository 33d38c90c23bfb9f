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
 *   segment j's large-spread pairs are cuts[cut_off[j] .. cut_off[j+1]) (ascending pair indices c, lo <= c < hi);
 *   pos / ctg [G x n_col]: positions and contigs of the vertices the call looks at, by column:
 *       j -> seg_lo[j];  n_seg + j -> seg_hi[j];  2 n_seg + k -> cuts[k];  2 n_seg + n_cut + k -> cuts[k] + 1.
 * A path's block = its segments from the last junction with a contig change on (the reference never flags the earlier
 * part); per assembly '+' / '-' if every step increases / decreases, else by the m_pct rule, else the block's segments
 * are deleted.  The block is then cut at every junction / inner pair whose |dpos| spread exceeds bp (that edge is
 * removed), and pieces with fewer than min_mx minimizers are deleted.
 * Outputs: blocks (b_off into the out segment arrays o_lo/o_hi/o_dir, b_n, b_first, b_last, b_ori / b_ctg / b_fpos /
 * b_lpos [n_blocks x G]); deleted segments (r_lo, r_hi); removed edges (e_u, e_v).  cap bounds every output array
 * (n_seg + n_cut + 1 is enough).  counts[4] = blocks, out segments, deleted segments, removed edges. */
int nts_host_paths_to_blocks(int64_t n_paths, const int64_t* path_off, const int64_t* seg_lo, const int64_t* seg_hi,
                             const int8_t* seg_dir, uint32_t G, const int64_t* up, const int64_t* down, const int64_t* cuts,
                             const int64_t* cut_off, const int64_t* pos, const int32_t* ctg, int64_t bp, double m_pct,
                             int64_t min_mx, int64_t cap, int64_t* b_off, int64_t* b_n, int64_t* b_first, int64_t* b_last,
                             int8_t* b_ori, int32_t* b_ctg, int64_t* b_fpos, int64_t* b_lpos, int64_t* o_lo, int64_t* o_hi,
                             int8_t* o_dir, int64_t* r_lo, int64_t* r_hi, int64_t* e_u, int64_t* e_v, int64_t counts[4])
{
    if (!path_off || !counts || (n_paths && (!seg_lo || !seg_hi || !seg_dir || !up || !down || !cut_off || !pos || !ctg)))
        return fail(NTS_ERR_ARG, "null argument");
    if (G < 1 || G > 32) return fail(NTS_ERR_ARG, "between 1 and 32 assemblies are supported");
    const int64_t n_seg = n_paths ? path_off[n_paths] : 0;
    const int64_t n_cut = n_seg ? cut_off[n_seg] : 0;
    const int64_t n_col = 2 * n_seg + 2 * n_cut;
    int64_t nb = 0, no = 0, nr = 0, ne = 0;
    auto col_first = [&](int64_t j) { return seg_dir[j] > 0 ? j : n_seg + j; };
    auto col_last = [&](int64_t j) { return seg_dir[j] > 0 ? n_seg + j : j; };
    auto first_of = [&](int64_t j) { return seg_dir[j] > 0 ? seg_lo[j] : seg_hi[j]; };
    auto last_of = [&](int64_t j) { return seg_dir[j] > 0 ? seg_hi[j] : seg_lo[j]; };
    std::vector<int64_t> inc(G), dec(G);
    std::vector<int8_t> ori(G);
    struct Seg { int64_t lo, hi; int8_t d; int64_t c_lo, c_hi; };      // c_*: columns of the positions of lo / hi
    std::vector<Seg> cur;
    std::vector<int64_t> piece_end;                                     // pieces = consecutive ranges of `all`
    std::vector<Seg> all;
    for (int64_t p = 0; p < n_paths; ++p) {
        const int64_t s0 = path_off[p], s1 = path_off[p + 1];
        if (s1 <= s0) continue;
        // the block starts after the last junction where some assembly changes contig
        int64_t start = s0;
        for (int64_t j = s0 + 1; j < s1; ++j) {
            const int64_t cu = col_last(j - 1), cv = col_first(j);
            bool chg = false;
            for (uint32_t a = 0; a < G && !chg; ++a) chg = ctg[a * n_col + cu] != ctg[a * n_col + cv];
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
                const int64_t cu = col_last(j - 1), cv = col_first(j);
                for (uint32_t a = 0; a < G; ++a) {
                    const int64_t d = pos[a * n_col + cv] - pos[a * n_col + cu];
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
        const int64_t c_first = col_first(start);
        // indel cuts: junctions and inner pairs whose |dpos| spread exceeds bp
        all.clear(); piece_end.clear();
        for (int64_t j = start; j < s1; ++j) {
            const int64_t slo = seg_lo[j], shi = seg_hi[j];
            const int8_t d = seg_dir[j];
            if (j > start) {
                const int64_t cu = col_last(j - 1), cv = col_first(j);
                int64_t mx = 0, mn = INT64_MAX;
                for (uint32_t a = 0; a < G; ++a) {
                    int64_t dd = pos[a * n_col + cv] - pos[a * n_col + cu];
                    if (dd < 0) dd = -dd;
                    mx = std::max(mx, dd); mn = std::min(mn, dd);
                }
                if (mx - mn > bp) {
                    if (ne >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    e_u[ne] = last_of(j - 1); e_v[ne] = first_of(j); ++ne;
                    piece_end.push_back((int64_t)all.size());
                }
            }
            const int64_t k0 = cut_off[j], k1 = cut_off[j + 1];
            if (k1 > k0) {
                for (int64_t k = k0; k < k1; ++k) {
                    if (ne >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    e_u[ne] = cuts[k]; e_v[ne] = cuts[k] + 1; ++ne;
                }
                if (d > 0) {
                    int64_t a0 = slo, ca = j;
                    for (int64_t k = k0; k < k1; ++k) {
                        all.push_back({a0, cuts[k], 1, ca, 2 * n_seg + k}); piece_end.push_back((int64_t)all.size());
                        a0 = cuts[k] + 1; ca = 2 * n_seg + n_cut + k;
                    }
                    all.push_back({a0, shi, 1, ca, n_seg + j});
                } else {
                    int64_t a0 = shi, ca = n_seg + j;
                    for (int64_t k = k1; k-- > k0;) {
                        all.push_back({cuts[k] + 1, a0, -1, 2 * n_seg + n_cut + k, ca}); piece_end.push_back((int64_t)all.size());
                        a0 = cuts[k]; ca = 2 * n_seg + k;
                    }
                    all.push_back({slo, a0, -1, j, ca});
                }
            } else {
                all.push_back({slo, shi, d, j, n_seg + j});
            }
        }
        piece_end.push_back((int64_t)all.size());
        int64_t q0 = 0;
        for (int64_t q1 : piece_end) {
            int64_t pn = 0;
            for (int64_t i = q0; i < q1; ++i) pn += all[i].hi - all[i].lo + 1;
            if (pn < min_mx) {
                for (int64_t i = q0; i < q1; ++i) {
                    if (nr >= cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                    r_lo[nr] = all[i].lo; r_hi[nr] = all[i].hi; ++nr;
                }
            } else if (q1 > q0) {
                if (nb >= cap || no + (q1 - q0) > cap) return fail(NTS_ERR_OVERFLOW, "paths_to_blocks: output capacity too small");
                const Seg& sf = all[q0];
                const Seg& sl = all[q1 - 1];
                const int64_t f = sf.d > 0 ? sf.lo : sf.hi, cf = sf.d > 0 ? sf.c_lo : sf.c_hi;
                const int64_t l = sl.d > 0 ? sl.hi : sl.lo, cl = sl.d > 0 ? sl.c_hi : sl.c_lo;
                b_off[nb] = no; b_n[nb] = pn; b_first[nb] = f; b_last[nb] = l;
                for (uint32_t a = 0; a < G; ++a) {
                    b_ori[nb * G + a] = ori[a];
                    b_ctg[nb * G + a] = ctg[a * n_col + c_first];            // the contigs of the path's block, also for its pieces
                    b_fpos[nb * G + a] = pos[a * n_col + cf];
                    b_lpos[nb * G + a] = pos[a * n_col + cl];
                }
                for (int64_t i = q0; i < q1; ++i) { o_lo[no] = all[i].lo; o_hi[no] = all[i].hi; o_dir[no] = all[i].d; ++no; }
                ++nb;
            }
            q0 = q1;
        }
    }
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
