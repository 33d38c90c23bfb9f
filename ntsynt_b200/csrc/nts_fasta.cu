// nts_fasta.cu -- native FASTA ingest (host code only): records -> 2-bit packed contigs + N-run lists + .fai rows.
//
// Replaces btllib::SeqReader for this path (src/ntsynt_make_common_bf.cpp:32-36,125,143; inside indexlr) and
// `samtools faidx` (bin/ntsynt_run_pipeline.smk:48-53).  The reference re-reads every FASTA once per stage; here a
// file is scanned once (memchr over lines), and packed by a pool of threads: records in parallel, and the lines of a
// long record in parallel too when its line width is uniform (every FASTA that `samtools faidx` accepts), because
// then the file offset of base i is seq_off + (i / linebases) * linewidth + i % linebases.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "nts_internal.h"

using namespace nts;

namespace {

struct Lut { uint8_t v[256]; };
const Lut& lut()
{
    static const Lut l = [] {
        Lut x;
        for (int i = 0; i < 256; ++i) x.v[i] = 4;
        x.v['A'] = x.v['a'] = 0; x.v['C'] = x.v['c'] = 1; x.v['G'] = x.v['g'] = 2; x.v['T'] = x.v['t'] = 3;
        x.v['\n'] = x.v['\r'] = 5;                       // line ends: skipped
        return x;
    }();
    return l;
}

struct Run { uint64_t start, len; };

// pack the bases found in file bytes [p, e) (line ends skipped) as bases [b0, ...) of a record whose words start at
// `words`; b0 is a multiple of 32 or the piece is the record's only one.  Appends the piece's N runs (record coordinates).
uint64_t pack_piece(const char* p, const char* e, uint64_t b0, uint64_t* words, std::vector<Run>& runs)
{
    const Lut& L = lut();
    uint64_t b = b0, word = 0;
    bool in_run = false;
    uint64_t run_start = 0;
    constexpr uint64_t ONES = 0x0101010101010101ull, LOW7 = 0x7F7F7F7F7F7F7F7Full, HIGH = 0x8080808080808080ull;
    // 0x80 in every byte of v that is zero (exact per byte: no carries cross byte boundaries)
    auto zero_bytes = [&](uint64_t v) { return ~(((v & LOW7) + LOW7) | v) & HIGH; };
    while (p < e) {
        uint64_t nslow = 1;
        // fast path: 8 bytes that are all ACGT / acgt (no line end, nothing ambiguous)
        if (e - p >= 8) {
            uint64_t x;
            memcpy(&x, p, 8);
            const uint64_t up = x & 0xDFDFDFDFDFDFDFDFull;                       // fold to upper case
            const uint64_t ok = zero_bytes(up ^ (ONES * 'A')) | zero_bytes(up ^ (ONES * 'C')) |
                                zero_bytes(up ^ (ONES * 'G')) | zero_bytes(up ^ (ONES * 'T'));
            if (ok == HIGH) {
                if (in_run) { runs.push_back({run_start, b - run_start}); in_run = false; }
                // A 0x41 C 0x43 G 0x47 T 0x54: ((c >> 1) ^ (c >> 2)) & 3 = 0, 1, 2, 3
                const uint64_t c2 = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ull;
                // gather the four 2-bit codes of each 32-bit half into one byte: bit 8i -> bit 24 + 2i
                constexpr uint32_t M = (1u << 24) | (1u << 18) | (1u << 12) | (1u << 6);
                const uint32_t lo = ((uint32_t)c2 * M) >> 24, hi = ((uint32_t)(c2 >> 32) * M) >> 24;
                const uint64_t code = lo | (hi << 8);
                const unsigned sh = (unsigned)(b & 31) * 2;
                word |= code << sh;
                const uint64_t nb = b + 8;
                if ((nb >> 5) != (b >> 5)) { words[b >> 5] = word; word = sh > 48 ? code >> (64 - sh) : 0; }   // spill into the next word
                p += 8; b = nb;
                continue;
            }
            nslow = (uint64_t)(__builtin_ctzll(~ok & HIGH) >> 3) + 1;            // the valid prefix + the byte that is not a base
        }
        // slow path: those bytes one by one, then on through whatever else is not a base (line ends, N runs)
        for (;;) {
            uint8_t c = L.v[(unsigned char)*p];
            ++p;
            if (c != 5) {
                if (c == 4) {
                    if (!in_run) { in_run = true; run_start = b; }
                    c = 0;
                } else if (in_run) {
                    runs.push_back({run_start, b - run_start});
                    in_run = false;
                }
                word |= (uint64_t)c << ((b & 31) * 2);
                if ((++b & 31) == 0) { words[(b >> 5) - 1] = word; word = 0; }
            }
            if (p >= e) break;
            if (nslow > 1) { --nslow; continue; }
            if (L.v[(unsigned char)*p] < 4) break;
        }
    }
    if (b & 31) words[b >> 5] = word;
    if (in_run) runs.push_back({run_start, b - run_start});
    return b;
}

}  // namespace

extern "C" {

/* Pass 1: find the records of a FASTA held in memory.  Per record: name = header up to the first whitespace
 * (name_off / name_len into buf), n_bases, seq_off = file offset of the first sequence byte, seq_end = one past the
 * last, linebases / linewidth of its first sequence line (the .fai columns), uniform = 1 iff every sequence line but
 * the last has exactly that width.  Returns the number of records through n_records (may exceed cap: call again). */
int nts_fasta_scan(const char* buf, uint64_t n, uint64_t cap, uint64_t* name_off, uint32_t* name_len, uint64_t* n_bases,
                   uint64_t* seq_off, uint64_t* seq_end, uint32_t* linebases, uint32_t* linewidth, uint8_t* uniform,
                   uint64_t* n_records)
{
    if ((!buf && n) || !n_records) return fail(NTS_ERR_ARG, "null argument");
    uint64_t nr = 0;
    const char* p = buf;
    const char* const end = buf + n;
    bool have = false;
    uint64_t bases = 0, lb = 0, lw = 0, soff = 0, send = 0;
    bool uni = true, short_seen = false;
    auto flush = [&]() {
        if (have && nr <= cap && nr > 0) {
            const uint64_t i = nr - 1;
            if (i < cap) { n_bases[i] = bases; seq_off[i] = soff; seq_end[i] = send; linebases[i] = (uint32_t)lb; linewidth[i] = (uint32_t)lw; uniform[i] = uni ? 1 : 0; }
        }
    };
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;                  // line = [p, le), terminator at le (if any)
        const uint64_t raw = (uint64_t)(le - p) + (nl ? 1 : 0);
        if (p < le && *p == '>') {
            flush();
            ++nr;
            have = true;
            auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
            const char* q0 = p + 1;
            while (q0 < le && is_ws(*q0)) ++q0;           // first whitespace-delimited token of the header
            const char* q = q0;
            while (q < le && !is_ws(*q)) ++q;
            if (nr - 1 < cap) { name_off[nr - 1] = (uint64_t)(q0 - buf); name_len[nr - 1] = (uint32_t)(q - q0); }
            bases = 0; lb = lw = 0; uni = true; short_seen = false;
            soff = send = (uint64_t)((nl ? nl + 1 : end) - buf);
        } else if (have) {
            uint64_t len = (uint64_t)(le - p);
            while (len && (p[len - 1] == '\r')) --len;   // rstrip("\r\n")
            if (lb == 0) { lb = len; lw = raw; }         // .fai columns: the first non-empty sequence line
            if (short_seen && len) uni = false;          // a line after a short / odd one
            if (len != lb || raw != lw) short_seen = true;
            bases += len;
            send = (uint64_t)(le - buf);
        }
        p = nl ? nl + 1 : end;
    }
    flush();
    *n_records = nr;
    return NTS_OK;
}

/* Pass 1 with threads: the same result as nts_fasta_scan for any input.
 *   A. header lines ('>' first in the buffer or right after a newline) are found by all threads, a slice of the buffer each;
 *   B. every record's head is walked with the serial rule through its first non-empty sequence line (which fixes the
 *      .fai columns linebases / linewidth and the state of the uniformity test);
 *   C. the rest of every record is cut at line starts into ~8 MB pieces that the threads scan with that state known:
 *      bases, end of the last line, whether an odd line occurs and whether a non-empty line follows one;
 *   D. the pieces of a record are merged in order (a non-empty line after an odd line anywhere makes it non-uniform). */
int nts_fasta_scan_mt(const char* buf, uint64_t n, uint64_t cap, uint64_t* name_off, uint32_t* name_len, uint64_t* n_bases,
                      uint64_t* seq_off, uint64_t* seq_end, uint32_t* linebases, uint32_t* linewidth, uint8_t* uniform,
                      uint64_t* n_records, uint32_t n_threads)
{
    if ((!buf && n) || !n_records) return fail(NTS_ERR_ARG, "null argument");
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    if (n_threads == 1 || n < (1ull << 24))
        return nts_fasta_scan(buf, n, cap, name_off, name_len, n_bases, seq_off, seq_end, linebases, linewidth, uniform, n_records);
    const char* const end = buf + n;
    auto run_threads = [&](size_t n_jobs, auto&& job) {
        std::atomic<size_t> next{0};
        auto work = [&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n_jobs) break; job(i); } };
        std::vector<std::thread> pool;
        const uint32_t nt = (uint32_t)std::min<size_t>(n_threads, std::max<size_t>(n_jobs, 1));
        for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    };
    // ---- A: header positions
    const size_t n_slices = (size_t)n_threads * 4;
    const uint64_t slice = (n + n_slices - 1) / n_slices;
    std::vector<std::vector<uint64_t>> found(n_slices);
    run_threads(n_slices, [&](size_t i) {
        const char* p = buf + std::min<uint64_t>(n, i * slice);
        const char* const e = buf + std::min<uint64_t>(n, (i + 1) * slice);
        while (p < e) {
            const char* q = (const char*)memchr(p, '>', (size_t)(e - p));
            if (!q) break;
            if (q == buf || q[-1] == '\n') found[i].push_back((uint64_t)(q - buf));
            p = q + 1;
        }
    });
    std::vector<uint64_t> hdr;
    for (auto& f : found) hdr.insert(hdr.end(), f.begin(), f.end());
    const uint64_t nr = hdr.size();
    *n_records = nr;
    if (nr > cap) return NTS_OK;                      // the caller comes back with larger arrays
    // ---- B: names, heads
    struct Head { uint64_t lb, lw, bases, send, resume, span_end; bool short_seen, uni; };
    std::vector<Head> heads(nr);
    for (uint64_t r = 0; r < nr; ++r) {
        const char* p = buf + hdr[r];
        const char* const rec_end = r + 1 < nr ? buf + hdr[r + 1] : end;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(rec_end - p));
        const char* le = nl ? nl : rec_end;
        auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
        const char* q0 = p + 1;
        while (q0 < le && is_ws(*q0)) ++q0;
        const char* q = q0;
        while (q < le && !is_ws(*q)) ++q;
        name_off[r] = (uint64_t)(q0 - buf); name_len[r] = (uint32_t)(q - q0);
        Head h{0, 0, 0, 0, 0, (uint64_t)(rec_end - buf), false, true};
        const uint64_t soff = (uint64_t)((nl ? nl + 1 : rec_end) - buf);
        seq_off[r] = soff;
        h.send = soff;
        p = buf + soff;
        while (p < rec_end) {                          // the serial rule, until the first non-empty line has been taken
            nl = (const char*)memchr(p, '\n', (size_t)(rec_end - p));
            le = nl ? nl : rec_end;
            const uint64_t raw = (uint64_t)(le - p) + (nl ? 1 : 0);
            uint64_t len = (uint64_t)(le - p);
            while (len && (p[len - 1] == '\r')) --len;
            const bool first = h.lb == 0;
            if (first) { h.lb = len; h.lw = raw; }
            if (h.short_seen && len) h.uni = false;
            if (len != h.lb || raw != h.lw) h.short_seen = true;
            h.bases += len;
            h.send = (uint64_t)(le - buf);
            p = nl ? nl + 1 : rec_end;
            if (first && len) break;
        }
        h.resume = (uint64_t)(p - buf);
        heads[r] = h;
    }
    // ---- C: pieces of the record bodies
    struct Piece { uint64_t rec, a, b, bases, send; bool any_line, odd, viol, nonempty; };
    std::vector<Piece> pieces;
    const uint64_t target = 8ull << 20;
    for (uint64_t r = 0; r < nr; ++r) {
        uint64_t a = heads[r].resume;
        const uint64_t e = heads[r].span_end;
        while (a < e) {
            uint64_t b = std::min(e, a + target);
            if (b < e) {                               // cut after the next newline
                const char* nl = (const char*)memchr(buf + b, '\n', (size_t)(e - b));
                b = nl ? (uint64_t)(nl + 1 - buf) : e;
            }
            pieces.push_back({r, a, b, 0, 0, false, false, false, false});
            a = b;
        }
    }
    run_threads(pieces.size(), [&](size_t i) {
        Piece& pc = pieces[i];
        const uint64_t lb = heads[pc.rec].lb, lw = heads[pc.rec].lw;
        const char* p = buf + pc.a;
        const char* const e = buf + pc.b;
        while (p < e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* le = nl ? nl : e;
            const uint64_t raw = (uint64_t)(le - p) + (nl ? 1 : 0);
            uint64_t len = (uint64_t)(le - p);
            while (len && (p[len - 1] == '\r')) --len;
            if (pc.odd && len) pc.viol = true;
            if (len != lb || raw != lw) pc.odd = true;
            if (len) pc.nonempty = true;
            pc.bases += len;
            pc.send = (uint64_t)(le - buf);
            pc.any_line = true;
            p = nl ? nl + 1 : e;
        }
    });
    // ---- D: merge
    size_t pi = 0;
    for (uint64_t r = 0; r < nr; ++r) {
        Head& h = heads[r];
        for (; pi < pieces.size() && pieces[pi].rec == r; ++pi) {
            const Piece& pc = pieces[pi];
            if ((h.short_seen && pc.nonempty) || pc.viol) h.uni = false;
            if (pc.odd) h.short_seen = true;
            h.bases += pc.bases;
            if (pc.any_line) h.send = pc.send;
        }
        n_bases[r] = h.bases; seq_end[r] = h.send; linebases[r] = (uint32_t)h.lb; linewidth[r] = (uint32_t)h.lw; uniform[r] = h.uni ? 1 : 0;
    }
    return NTS_OK;
}

/* Pass 2: pack every record.  word_off[r] = offset (in 64-bit words, even) of record r in words_out, which holds
 * sum of nts_packed_words(n_bases[r]) words; N runs of all records go to nrun_start / nrun_len (record coordinates),
 * record r owning entries [nrun_off[r], nrun_off[r+1]).  If there are more runs than nrun_cap only the count is
 * returned (call again with larger arrays).  n_threads = 0 picks the hardware concurrency. */
int nts_fasta_pack(const char* buf, uint64_t n_records, const uint64_t* n_bases, const uint64_t* seq_off, const uint64_t* seq_end,
                   const uint32_t* linebases, const uint32_t* linewidth, const uint8_t* uniform, const uint64_t* word_off,
                   uint64_t* words_out, uint64_t* nrun_off, uint64_t* nrun_start, uint64_t* nrun_len, uint64_t nrun_cap,
                   uint64_t* n_nruns, uint32_t n_threads)
{
    if (!buf || !n_bases || !seq_off || !seq_end || !word_off || !words_out || !nrun_off || !n_nruns)
        return fail(NTS_ERR_ARG, "null argument");
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    // pieces: (record, first base, last base) with first base a multiple of 32; ~4 Mbp each where the record allows it
    struct Piece { uint64_t rec, b0, b1; const char* p; const char* e; };
    std::vector<Piece> pieces;
    const uint64_t target = 1ull << 22;
    for (uint64_t r = 0; r < n_records; ++r) {
        const uint64_t nb = n_bases[r];
        const char* p = buf + seq_off[r];
        const char* e = buf + seq_end[r];
        if (!uniform[r] || nb <= target || linebases[r] == 0) { pieces.push_back({r, 0, nb, p, e}); continue; }
        const uint64_t lb = linebases[r], lw = linewidth[r];
        auto off_of = [&](uint64_t b) { return seq_off[r] + (b / lb) * lw + b % lb; };
        for (uint64_t b0 = 0; b0 < nb; b0 += target) {
            const uint64_t b1 = std::min(nb, b0 + target);
            pieces.push_back({r, b0, b1, buf + off_of(b0), b1 == nb ? e : buf + off_of(b1)});
        }
    }
    std::vector<std::vector<Run>> piece_runs(pieces.size());
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= pieces.size()) break;
            const Piece& pc = pieces[i];
            const uint64_t got = pack_piece(pc.p, pc.e, pc.b0, words_out + word_off[pc.rec], piece_runs[i]);
            if (got != pc.b1) bad.store(1);
        }
    };
    std::vector<std::thread> pool;
    const uint32_t nt = (uint32_t)std::min<size_t>(n_threads, std::max<size_t>(pieces.size(), 1));
    for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (bad.load()) return fail(NTS_ERR_STATE, "FASTA record does not have the uniform line width its first lines announce");
    // stitch the runs: pieces are in (record, base) order; a run ending at a piece boundary continues in the next piece
    uint64_t total = 0;
    size_t i = 0;
    for (uint64_t r = 0; r < n_records; ++r) {
        nrun_off[r] = total;
        bool open = false;
        uint64_t cs = 0, cl = 0;
        auto emit = [&]() {
            if (!open) return;
            if (total < nrun_cap && nrun_start && nrun_len) { nrun_start[total] = cs; nrun_len[total] = cl; }
            ++total; open = false;
        };
        for (; i < pieces.size() && pieces[i].rec == r; ++i)
            for (const Run& ru : piece_runs[i]) {
                if (open && cs + cl == ru.start) { cl += ru.len; continue; }
                emit();
                open = true; cs = ru.start; cl = ru.len;
            }
        emit();
    }
    nrun_off[n_records] = total;
    *n_nruns = total;
    return NTS_OK;
}

}  // extern "C"
