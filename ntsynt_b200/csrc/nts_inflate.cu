// gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for FASTA ingest: whole buffer in, whole buffer out (host code).
//
// Replaces the decompressor btllib::SeqReader spawns for .gz input (src/ntsynt_make_common_bf.cpp:32-36,125,143 read their
// genomes through it).  A genome's .gz is one long stream of mostly literals and short matches, so the cost is Huffman
// decoding per symbol: a 64-bit bit buffer refilled eight bytes at a time, an 11-bit first-level table for the
// literal/length code and an 8-bit one for the distance code (second-level tables behind the long codes), matches copied
// eight bytes at a time.  The CRC-32 of the output is checked by a second thread that follows the decoder.  A long member is
// decoded by several threads (inflate_member_parallel, further down): chunks that start without their history.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsynt_b200.h"
#include "nts_internal.h"

namespace {

constexpr int LIT_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
constexpr uint32_t K_LITERAL = 0, K_LENGTH = 1, K_END = 2, K_SUB = 3, K_INVALID = 4;
// table entry: bits 0-4 code length (bits to drop), 5-7 kind, 8-12 extra bits (K_SUB: index bits of the second level),
// 16-31 base value (literal, first length / distance of the symbol, or start of the second-level table)
inline uint32_t entry(uint32_t len, uint32_t kind, uint32_t extra, uint32_t base) { return len | (kind << 5) | (extra << 8) | (base << 16); }
inline uint32_t e_len(uint32_t e) { return e & 31u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 31u; }
inline uint32_t e_base(uint32_t e) { return e >> 16; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Which { LITLEN, DIST, PRECODE };

struct Tables {
    uint32_t lit[(1 << LIT_BITS) + 288 * 16];
    uint32_t dist[(1 << DIST_BITS) + 32 * 128];
    uint32_t pre[1 << PRE_BITS];
};

inline uint32_t symbol_entry(Which which, uint32_t sym, uint32_t len) {
    if (which == PRECODE) return entry(len, K_LITERAL, 0, sym);
    if (which == DIST) return sym < 30 ? entry(len, K_LENGTH, DIST_EXTRA[sym], DIST_BASE[sym]) : entry(len, K_INVALID, 0, 0);
    if (sym < 256) return entry(len, K_LITERAL, 0, sym);
    if (sym == 256) return entry(len, K_END, 0, 0);
    return sym < 286 ? entry(len, K_LENGTH, LEN_EXTRA[sym - 257], LEN_BASE[sym - 257]) : entry(len, K_INVALID, 0, 0);
}

// canonical Huffman code -> lookup table indexed by the next bits of the stream (LSB first).  false = over-subscribed code,
// or an incomplete one other than the single-code cases RFC 1951 allows (their unused entries stay K_INVALID)
bool build_table(Which which, const uint8_t* lens, int n_sym, uint32_t* table, int table_bits) {
    int count[16] = {0};
    for (int s = 0; s < n_sym; ++s) count[lens[s]]++;
    count[0] = 0;
    uint32_t next_code[16] = {0}, code = 0;
    int64_t left = 1;
    int n_codes = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        n_codes += count[l];
        code = (code + count[l - 1]) << 1;
        next_code[l] = code;
    }
    // incomplete: only "no code at all" or "one code of one bit" (RFC 1951 3.2.7), and never for the code-length code
    if (left > 0 && (which == PRECODE || !(n_codes == 0 || (n_codes == 1 && count[1] == 1)))) return false;
    const uint32_t primary = 1u << table_bits;
    for (uint32_t i = 0; i < primary; ++i) table[i] = entry(1, K_INVALID, 0, 0);
    // pass 1: how deep the second level behind every first-level index has to be
    uint8_t sub_bits[1 << LIT_BITS];
    memset(sub_bits, 0, primary);
    uint32_t rev[288];
    {
        uint32_t nc[16];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n_sym; ++s) {
            int l = lens[s];
            if (!l) continue;
            uint32_t c = nc[l]++, r = 0;
            for (int b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);
            rev[s] = r;
            if (l > table_bits) {
                uint32_t p = r & (primary - 1);
                if (l - table_bits > sub_bits[p]) sub_bits[p] = (uint8_t)(l - table_bits);
            }
        }
    }
    uint32_t next = primary;
    for (uint32_t p = 0; p < primary; ++p)
        if (sub_bits[p]) {
            table[p] = entry(table_bits, K_SUB, sub_bits[p], next);
            for (uint32_t j = 0; j < (1u << sub_bits[p]); ++j) table[next + j] = entry(1, K_INVALID, 0, 0);
            next += 1u << sub_bits[p];
        }
    // pass 2: every code fills the entries whose low bits spell it
    for (int s = 0; s < n_sym; ++s) {
        int l = lens[s];
        if (!l) continue;
        uint32_t r = rev[s];
        if (l <= table_bits) {
            uint32_t e = symbol_entry(which, s, l);
            for (uint32_t i = r; i < primary; i += 1u << l) table[i] = e;
        } else {
            uint32_t p = r & (primary - 1), start = e_base(table[p]), sb = sub_bits[p];
            uint32_t e = symbol_entry(which, s, l - table_bits);
            for (uint32_t i = r >> table_bits; i < (1u << sb); i += 1u << (l - table_bits)) table[start + i] = e;
        }
    }
    return true;
}

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf = 0;
    int bits = 0;            // valid bits in buf
    bool overrun = false;    // a code or field asked for bits past the end of the input
    void refill() {
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << bits;
            p += (63 - bits) >> 3;
            bits |= 56;
        } else {
            while (bits <= 56 && p < end) {
                buf |= (uint64_t)(*p++) << bits;
                bits += 8;
            }
        }
    }
    uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    void drop(int n) {
        if (n > bits) { overrun = true; n = bits; }
        buf >>= n;
        bits -= n;
    }
    uint32_t take(int n) {
        uint32_t v = peek(n);
        drop(n);
        return v;
    }
    void align_to_byte() {                 // give whole unread bytes back to the input, drop the rest of the current one
        drop(bits & 7);
        p -= bits >> 3;
        buf = 0;
        bits = 0;
    }
};

enum { OK = 0, OUTPUT_FULL = 1, BAD = 2, TRUNCATED = 3 };

inline uint32_t decode(Reader& r, const uint32_t* table, int table_bits) {
    uint32_t e = table[r.peek(table_bits)];
    if (e_kind(e) == K_SUB) {
        r.drop(table_bits);
        e = table[e_base(e) + r.peek(e_extra(e))];
    }
    r.drop(e_len(e));
    return e;
}

// the bulk of a Huffman-coded block: runs while at least 16 input bytes and 320 bytes of output room are left, so nothing
// in it checks a bound; the bit buffer, the input and the output position live in locals (a byte store may alias anything
// that sits in memory, so state kept behind a reference would be reloaded after every literal).  Returns OK when it ran out
// of room (the careful loop below finishes the block), K_END + 16 when it met the end-of-block code, BAD on invalid data
constexpr int FAST_DONE = 16;
// O = uint8_t: the text itself.  O = uint16_t: the symbolic form of a chunk decoded without its history (parallel decode,
// below): literals are 0..255, everything copied out of the unknown 32 KB before the chunk is a placeholder >= 0x8000
template <typename O>
int inflate_block_fast(Reader& r_, const Tables& t, O* out_begin, O*& out_, O* out_end) {
    const uint8_t* p = r_.p;
    const uint8_t* const end = r_.end;
    uint64_t buf = r_.buf;
    int bits = r_.bits;
    O* o = out_;
    const uint32_t* const lit = t.lit;
    const uint32_t* const dist = t.dist;
    int rc = OK;
#define NTS_REFILL()                    \
    {                                   \
        uint64_t w_;                    \
        memcpy(&w_, p, 8);              \
        buf |= w_ << bits;              \
        p += (63 - bits) >> 3;          \
        bits |= 56;                     \
    }
#define NTS_DROP(n) { buf >>= (n); bits -= (int)(n); }
    while (end - p >= 16 && out_end - o >= 320) {
        NTS_REFILL();
        uint32_t e = lit[buf & ((1u << LIT_BITS) - 1)];
        if (e_kind(e) == K_LITERAL) {                        // up to four literals from one refill (4 x 11 bits <= 56)
            NTS_DROP(e_len(e)); *o++ = (O)e_base(e);
            e = lit[buf & ((1u << LIT_BITS) - 1)];
            if (e_kind(e) == K_LITERAL) {
                NTS_DROP(e_len(e)); *o++ = (O)e_base(e);
                e = lit[buf & ((1u << LIT_BITS) - 1)];
                if (e_kind(e) == K_LITERAL) {
                    NTS_DROP(e_len(e)); *o++ = (O)e_base(e);
                    e = lit[buf & ((1u << LIT_BITS) - 1)];
                    if (e_kind(e) == K_LITERAL) {
                        NTS_DROP(e_len(e)); *o++ = (O)e_base(e);
                        continue;
                    }
                }
            }
            NTS_REFILL();                                    // what follows may need 48 bits
        }
        if (e_kind(e) == K_SUB) {
            NTS_DROP(LIT_BITS);
            e = lit[e_base(e) + (uint32_t)(buf & ((1u << e_extra(e)) - 1))];
            if (e_kind(e) == K_LITERAL) {
                NTS_DROP(e_len(e)); *o++ = (O)e_base(e);
                continue;
            }
        }
        NTS_DROP(e_len(e));
        if (e_kind(e) != K_LENGTH) {
            rc = e_kind(e) == K_END ? FAST_DONE : BAD;
            break;
        }
        uint32_t len = e_base(e) + (uint32_t)(buf & ((1u << e_extra(e)) - 1));
        NTS_DROP(e_extra(e));
        uint32_t d = dist[buf & ((1u << DIST_BITS) - 1)];
        if (e_kind(d) == K_SUB) {
            NTS_DROP(DIST_BITS);
            d = dist[e_base(d) + (uint32_t)(buf & ((1u << e_extra(d)) - 1))];
        }
        NTS_DROP(e_len(d));
        if (e_kind(d) != K_LENGTH) { rc = BAD; break; }
        uint32_t off = e_base(d) + (uint32_t)(buf & ((1u << e_extra(d)) - 1));
        NTS_DROP(e_extra(d));
        if (off > (uint64_t)(o - out_begin)) { rc = BAD; break; }
        const O* src = o - off;
        O* stop = o + len;
        constexpr uint32_t PER_WORD = 8 / sizeof(O);
        if (off >= PER_WORD) {
            do {
                uint64_t w;
                memcpy(&w, src, 8);
                memcpy(o, &w, 8);
                src += PER_WORD;
                o += PER_WORD;
            } while (o < stop);
        } else if (off == 1 && sizeof(O) == 1) {
            memset(o, (int)*src, len);
        } else {
            do { *o++ = *src++; } while (o < stop);
        }
        o = stop;
    }
#undef NTS_REFILL
#undef NTS_DROP
    r_.p = p;
    r_.buf = buf;
    r_.bits = bits;
    out_ = o;
    return rc;
}

// one Huffman-coded block.  out_begin: start of the member's output (matches may reach back to it)
template <typename O>
int inflate_block(Reader& r, const Tables& t, O* out_begin, O*& out, O* out_end) {
    for (;;) {
        if (r.end - r.p >= 16 && out_end - out >= 320) {
            int rc = inflate_block_fast<O>(r, t, out_begin, out, out_end);
            if (rc == FAST_DONE) return OK;
            if (rc != OK) return rc;
        }
        r.refill();                        // >= 56 bits, or everything that is left: a symbol needs at most 15 + 5 + 15 + 13 = 48
        uint32_t e = decode(r, t.lit, LIT_BITS);
        if (r.overrun) return TRUNCATED;         // (before anything else: a cut stream must not read as "output full")
        uint32_t kind = e_kind(e);
        if (kind == K_LITERAL) {
            if (out == out_end) return OUTPUT_FULL;
            *out++ = (O)e_base(e);
            // literals come in runs: take more of them from the bits already loaded
            while (r.bits >= 30) {
                uint32_t e2 = t.lit[r.peek(LIT_BITS)];
                if (e_kind(e2) != K_LITERAL || out == out_end) break;
                r.drop(e_len(e2));
                *out++ = (O)e_base(e2);
            }
            if (r.overrun) return TRUNCATED;
            continue;
        }
        if (kind == K_END) return r.overrun ? TRUNCATED : OK;
        if (kind != K_LENGTH) return r.overrun ? TRUNCATED : BAD;
        uint32_t len = e_base(e) + r.take(e_extra(e));
        uint32_t d = decode(r, t.dist, DIST_BITS);
        if (e_kind(d) != K_LENGTH) return r.overrun ? TRUNCATED : BAD;
        uint32_t off = e_base(d) + r.take(e_extra(d));
        if (r.overrun) return TRUNCATED;
        if (off > (uint64_t)(out - out_begin)) return BAD;
        if ((uint64_t)(out_end - out) < len) return OUTPUT_FULL;
        const O* src = out - off;
        constexpr uint32_t PER_WORD = 8 / sizeof(O);
        if (off >= PER_WORD && (uint64_t)(out_end - out) >= len + PER_WORD) {
            O* dst = out;
            O* stop = out + len;
            do {
                uint64_t w;
                memcpy(&w, src, 8);
                memcpy(dst, &w, 8);
                src += PER_WORD;
                dst += PER_WORD;
            } while (dst < stop);
            out = stop;
        } else {
            for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
            out += len;
        }
    }
}

int read_dynamic_tables(Reader& r, Tables& t) {
    r.refill();
    uint32_t hlit = r.take(5) + 257, hdist = r.take(5) + 1, hclen = r.take(4) + 4;
    if (hlit > 286 || hdist > 30) return r.overrun ? TRUNCATED : BAD;
    uint8_t pre_lens[19] = {0};
    for (uint32_t i = 0; i < hclen; ++i) {
        r.refill();
        pre_lens[PRE_ORDER[i]] = (uint8_t)r.take(3);
    }
    if (r.overrun) return TRUNCATED;
    if (!build_table(PRECODE, pre_lens, 19, t.pre, PRE_BITS)) return BAD;
    uint8_t lens[286 + 30 + 138];
    uint32_t n = 0, total = hlit + hdist;
    while (n < total) {
        r.refill();
        uint32_t e = t.pre[r.peek(PRE_BITS)];
        if (e_kind(e) != K_LITERAL) return r.overrun ? TRUNCATED : BAD;
        r.drop(e_len(e));
        uint32_t sym = e_base(e);
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
        } else {
            uint32_t rep, val = 0;
            if (sym == 16) {
                if (!n) return BAD;
                val = lens[n - 1];
                rep = 3 + r.take(2);
            } else if (sym == 17) {
                rep = 3 + r.take(3);
            } else {
                rep = 11 + r.take(7);
            }
            if (n + rep > total) return r.overrun ? TRUNCATED : BAD;
            memset(lens + n, (int)val, rep);
            n += rep;
        }
        if (r.overrun) return TRUNCATED;
    }
    if (!lens[256]) return BAD;                  // no end-of-block code
    if (!build_table(LITLEN, lens, (int)hlit, t.lit, LIT_BITS)) return BAD;
    if (!build_table(DIST, lens + hlit, (int)hdist, t.dist, DIST_BITS)) return BAD;
    return OK;
}

void fixed_tables(Tables& t) {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(LITLEN, lens, 288, t.lit, LIT_BITS);
    uint8_t dl[32];
    memset(dl, 5, 32);
    build_table(DIST, dl, 32, t.dist, DIST_BITS);
}

// one block of any type at the reader's position; last = its BFINAL bit.  After the final block the reader is byte aligned
template <typename O>
int inflate_one_block(Reader& r, Tables& t, O* out_begin, O*& out, O* out_end, bool& last) {
    r.refill();
    last = r.take(1) != 0;
    uint32_t type = r.take(2);
    if (r.overrun) return TRUNCATED;
    if (type == 0) {
        r.align_to_byte();
        if (r.end - r.p < 4) return TRUNCATED;
        uint32_t len = r.p[0] | (r.p[1] << 8), nlen = r.p[2] | (r.p[3] << 8);
        if ((len ^ nlen) != 0xFFFFu) return BAD;
        r.p += 4;
        if ((uint64_t)(r.end - r.p) < len) return TRUNCATED;
        if ((uint64_t)(out_end - out) < len) return OUTPUT_FULL;
        for (uint32_t i = 0; i < len; ++i) out[i] = (O)r.p[i];
        out += len;
        r.p += len;
    } else if (type == 1 || type == 2) {
        if (type == 1) {
            fixed_tables(t);
        } else {
            int rc = read_dynamic_tables(r, t);
            if (rc != OK) return rc;
        }
        int rc = inflate_block<O>(r, t, out_begin, out, out_end);
        if (rc != OK) return rc;
    } else {
        return BAD;
    }
    if (last) r.align_to_byte();
    return OK;
}

// one DEFLATE stream starting at the reader's position; leaves r.p on the byte after it.  out_begin: where this stream's
// output starts; progress (optional): bytes of it that are final, published after every block
int inflate_stream(Reader& r, Tables& t, uint8_t* out_begin, uint8_t*& out, uint8_t* out_end, std::atomic<uint64_t>* progress) {
    for (;;) {
        if (progress) progress->store((uint64_t)(out - out_begin), std::memory_order_release);
        bool last;
        int rc = inflate_one_block<uint8_t>(r, t, out_begin, out, out_end, last);
        if (rc != OK) return rc;
        if (last) return OK;
    }
}

// CRC-32 (IEEE 802.3, reflected), eight bytes per step
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int k = 1; k < 8; ++k) t[k][i] = (t[k - 1][i] >> 8) ^ t[0][t[k - 1][i] & 255u];
    }
};

uint32_t crc32_update(uint32_t crc, const uint8_t* p, uint64_t n) {
    static const CrcTables T;
    crc = ~crc;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= crc;
        crc = T.t[7][w & 255] ^ T.t[6][(w >> 8) & 255] ^ T.t[5][(w >> 16) & 255] ^ T.t[4][(w >> 24) & 255] ^
              T.t[3][(w >> 32) & 255] ^ T.t[2][(w >> 40) & 255] ^ T.t[1][(w >> 48) & 255] ^ T.t[0][w >> 56];
        p += 8;
        n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 255u];
    return ~crc;
}

// the CRC of a member's output, computed behind the decoder: `done` = bytes of the member written so far
struct CrcFollower {
    const uint8_t* base;
    std::atomic<uint64_t> done{0};
    std::atomic<bool> finished{false};
    uint32_t crc = 0;
    std::thread th;
    explicit CrcFollower(const uint8_t* b) : base(b) {
        th = std::thread([this] {
            uint64_t at = 0;
            for (;;) {
                bool fin = finished.load(std::memory_order_acquire);
                uint64_t d = done.load(std::memory_order_acquire);
                if (d - at >= (1u << 20) || (fin && d > at)) {
                    crc = crc32_update(crc, base + at, d - at);
                    at = d;
                } else if (fin) {
                    return;
                } else {
                    std::this_thread::yield();
                }
            }
        });
    }
    uint32_t finish(uint64_t total) {
        done.store(total, std::memory_order_release);
        finished.store(true, std::memory_order_release);
        th.join();
        return crc;
    }
};

// crc of the concatenation A|B from crc(A), crc(B) and len(B): the operator "append one zero bit" as a 32 x 32 matrix over
// GF(2), squared up to len(B) zero bytes (the construction zlib documents for crc32_combine)
uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
    uint32_t sum = 0;
    for (int i = 0; vec; vec >>= 1, ++i)
        if (vec & 1) sum ^= mat[i];
    return sum;
}
void gf2_square(uint32_t* sq, const uint32_t* mat) {
    for (int i = 0; i < 32; ++i) sq[i] = gf2_times(mat, mat[i]);
}
uint32_t crc32_concat(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
    if (!len_b) return crc_a;
    uint32_t even[32], odd[32];
    odd[0] = 0xEDB88320u;
    for (int i = 1; i < 32; ++i) odd[i] = 1u << (i - 1);
    gf2_square(even, odd);      // two zero bits
    gf2_square(odd, even);      // four zero bits
    do {
        gf2_square(even, odd);  // first round: one zero byte
        if (len_b & 1) crc_a = gf2_times(even, crc_a);
        len_b >>= 1;
        if (!len_b) break;
        gf2_square(odd, even);
        if (len_b & 1) crc_a = gf2_times(odd, crc_a);
        len_b >>= 1;
    } while (len_b);
    return crc_a ^ crc_b;
}

// ---- one long member decoded by several threads --------------------------------------------------------------------------
// The compressed bytes are cut into chunks.  The first chunk of a wave starts at a known block boundary with its history in
// place and is decoded straight into the output.  Every other chunk looks for a position at or after its cut where a
// dynamic-Huffman block starts and decodes (the three code tables of a block header are complete prefix codes -- random bits
// almost never are), and decodes from there WITHOUT its history: into 16-bit symbols, literals as themselves and whatever a
// match copies out of the unknown 32 KB before the chunk as "byte i of that window".  A chunk runs until it stands exactly
// on the position its successor started from.  Afterwards the windows are filled in in order (only the last 32 KB of every
// chunk have to be resolved one after the other), the chunks are translated to bytes in parallel and their CRCs combined.
// Nothing is trusted: a chunk whose start does not coincide with its predecessor's end, or that fails, is dropped, and the
// next wave starts, with history, where the last good chunk ended; length and CRC-32 of the member are checked as always.
constexpr uint32_t WINDOW = 32768;
constexpr uint16_t PLACEHOLDER = 0x8000;
constexpr uint64_t NO_END = ~0ull;

inline uint64_t bit_position(const Reader& r, const uint8_t* base) { return (uint64_t)(r.p - base) * 8 - (uint64_t)r.bits; }

inline void reader_at(Reader& r, const uint8_t* base, const uint8_t* end, uint64_t bit) {
    r.p = base + (bit >> 3);
    r.end = end;
    r.buf = 0;
    r.bits = 0;
    r.overrun = false;
    r.refill();
    r.drop((int)(bit & 7));
}

struct Chunk {
    uint64_t target_bit = 0;              // symbolic chunks: look for a block start from here on
    uint64_t end_target_bit = 0;          // stop at a block boundary at or after this (NO_END: run to the final block)
    std::atomic<int64_t> start_bit{-1};   // -1 not known yet, -2 none found
    uint64_t end_bit = 0;
    int rc = BAD;
    bool final_block = false;
    uint64_t n = 0;                       // elements decoded
    uint32_t crc = 0;
};

template <typename O>
void run_chunk(const uint8_t* base, const uint8_t* end, Chunk& c, Chunk* next, Tables& t, O* out_begin, O* o_start, O* o_end,
               bool search) {
    Reader r;
    O* o = o_start;
    if (search) {
        const uint64_t n_bits = (uint64_t)(end - base) * 8;
        const uint64_t reach = std::min<uint64_t>(8ull << 20, 4 * (c.end_target_bit == NO_END ? (8ull << 20) : c.end_target_bit - c.target_bit));
        const uint64_t limit = std::min<uint64_t>(c.target_bit + reach, n_bits > 128 ? n_bits - 128 : 0);
        int64_t found = -2;
        for (uint64_t b = c.target_bit; b < limit; ++b) {
            const uint8_t* q = base + (b >> 3);
            if (((((uint32_t)q[0] | ((uint32_t)q[1] << 8)) >> (b & 7)) & 7u) != 4u) continue;     // BFINAL 0, BTYPE 2 (dynamic)
            reader_at(r, base, end, b);
            o = o_start;
            bool last;
            if (inflate_one_block<O>(r, t, out_begin, o, o_end, last) == OK && o - o_start >= 256) {
                found = (int64_t)b;
                break;
            }
        }
        c.start_bit.store(found, std::memory_order_release);
        if (found < 0) return;
    } else {
        reader_at(r, base, end, (uint64_t)c.start_bit.load());
    }
    for (;;) {
        uint64_t pos = bit_position(r, base);
        if (pos >= c.end_target_bit) {
            if (!next) break;
            int64_t s;
            while ((s = next->start_bit.load(std::memory_order_acquire)) == -1) std::this_thread::yield();
            if (s < 0 || pos >= (uint64_t)s) break;      // equal: the successor takes over here; otherwise it is dropped
        }
        bool last;
        int rc = inflate_one_block<O>(r, t, out_begin, o, o_end, last);
        if (rc != OK) {
            c.rc = rc;
            c.n = (uint64_t)(o - o_start);
            return;
        }
        if (last) {
            c.final_block = true;
            break;
        }
    }
    c.end_bit = bit_position(r, base);
    c.n = (uint64_t)(o - o_start);
    c.rc = OK;
}

// member whose DEFLATE data starts at byte `start`; on OK `after` = first byte behind the data, crc = CRC-32 of the output
int inflate_member_parallel(const uint8_t* base, const uint8_t* end, uint64_t start, uint8_t* member_start, uint8_t*& o,
                            uint8_t* o_end, int n_threads, uint64_t chunk_bytes, bool want_crc, uint32_t& crc, const uint8_t*& after) {
    const uint64_t n_bits = (uint64_t)(end - base) * 8, chunk_bits = chunk_bytes * 8;
    const uint64_t sym_cap = WINDOW + 16 * chunk_bytes + 4096;          // elements; a chunk that needs more is dropped
    std::vector<Tables> tables((size_t)n_threads);
    std::vector<uint16_t*> sym((size_t)n_threads, nullptr);
    struct Free {
        std::vector<uint16_t*>& v;
        ~Free() { for (uint16_t* p : v) free(p); }
    } free_sym{sym};
    std::vector<uint8_t> windows((size_t)n_threads * WINDOW), lut((size_t)n_threads * 65536);
    uint64_t pos = start * 8;
    uint32_t crc_all = 0;
    // a stream without dynamic blocks (stored or fixed-Huffman data) gives the searching chunks nothing to find: after a wave
    // in which only the first chunk counted, that many chunks are decoded by one thread before the next attempt, doubling
    int fails = 0;
    bool solo = false;
    for (;;) {
        uint64_t left = n_bits - pos;
        int nc = (int)std::min<uint64_t>((uint64_t)n_threads, left / chunk_bits);
        uint64_t span = 1;
        if (solo) {
            nc = 1;
            span = 1ull << std::min(fails, 6);
        }
        if (nc < 2) nc = 1;
        for (int j = 1; j < nc; ++j)
            if (!sym[j]) {
                sym[j] = (uint16_t*)malloc(sym_cap * sizeof(uint16_t));
                if (!sym[j]) {                   // no memory for another symbolic chunk: fewer chunks in flight
                    nc = j;
                    break;
                }
                for (uint32_t i = 0; i < WINDOW; ++i) sym[j][i] = (uint16_t)(PLACEHOLDER + i);
            }
        const bool to_the_end = left <= ((uint64_t)nc * span + 1) * chunk_bits;
        std::vector<Chunk> ch((size_t)nc);
        for (int j = 0; j < nc; ++j) {
            ch[j].target_bit = pos + (uint64_t)j * chunk_bits;
            ch[j].end_target_bit = (j == nc - 1 && to_the_end) ? NO_END : pos + (uint64_t)(j + 1) * span * chunk_bits;
        }
        ch[0].start_bit.store((int64_t)pos);
        uint8_t* const wave_out = o;
        {
            std::vector<std::thread> th;
            for (int j = 1; j < nc; ++j)
                th.emplace_back([&, j] {
                    run_chunk<uint16_t>(base, end, ch[j], j + 1 < nc ? &ch[j + 1] : nullptr, tables[j], sym[j], sym[j] + WINDOW,
                                        sym[j] + sym_cap, true);
                });
            run_chunk<uint8_t>(base, end, ch[0], nc > 1 ? &ch[1] : nullptr, tables[0], member_start, wave_out, o_end, false);
            for (auto& x : th) x.join();
        }
        if (ch[0].rc != OK) {
            o = wave_out + ch[0].n;
            return ch[0].rc;
        }
        int valid = 1;
        while (valid < nc && !ch[valid - 1].final_block && ch[valid].rc == OK &&
               ch[valid].start_bit.load() == (int64_t)ch[valid - 1].end_bit)
            ++valid;
        std::vector<uint8_t*> dst((size_t)valid + 1);
        dst[0] = wave_out;
        for (int j = 0; j < valid; ++j) dst[j + 1] = dst[j] + ch[j].n;
        if (valid > 1 && (uint64_t)(dst[1] - member_start) < WINDOW) valid = 1;      // (cannot happen with chunks of megabytes)
        if (dst[valid] > o_end) {
            o = wave_out + ch[0].n;
            return OUTPUT_FULL;
        }
        // windows, one after the other: the 32 KB in front of chunk j are the resolved tail of chunk j - 1
        for (int j = 1; j < valid; ++j) {
            uint8_t* w = &windows[(size_t)j * WINDOW];
            uint8_t* l = &lut[(size_t)j * 65536];
            if (j == 1) {
                memcpy(w, dst[1] - WINDOW, WINDOW);
            } else {
                const uint8_t* pw = &windows[(size_t)(j - 1) * WINDOW];
                const uint8_t* pl = &lut[(size_t)(j - 1) * 65536];
                const uint16_t* ps = sym[j - 1] + WINDOW;
                uint64_t pn = ch[j - 1].n;
                if (pn >= WINDOW) {
                    for (uint32_t i = 0; i < WINDOW; ++i) w[i] = pl[ps[pn - WINDOW + i]];
                } else {
                    memcpy(w, pw + pn, WINDOW - pn);
                    for (uint64_t i = 0; i < pn; ++i) w[WINDOW - pn + i] = pl[ps[i]];
                }
            }
            for (uint32_t v = 0; v < 256; ++v) l[v] = (uint8_t)v;
            memset(l + 256, 0, PLACEHOLDER - 256);
            memcpy(l + PLACEHOLDER, w, WINDOW);
        }
        {
            std::vector<std::thread> th;
            for (int j = 1; j < valid; ++j)
                th.emplace_back([&, j] {
                    const uint8_t* l = &lut[(size_t)j * 65536];
                    const uint16_t* ps = sym[j] + WINDOW;
                    uint8_t* d = dst[j];
                    for (uint64_t i = 0, n = ch[j].n; i < n; ++i) d[i] = l[ps[i]];
                    if (want_crc) ch[j].crc = crc32_update(0, d, ch[j].n);
                });
            if (want_crc) ch[0].crc = crc32_update(0, dst[0], ch[0].n);
            for (auto& x : th) x.join();
        }
        if (want_crc)
            for (int j = 0; j < valid; ++j) crc_all = crc32_concat(crc_all, ch[j].crc, ch[j].n);
        if (solo) {
            solo = false;
        } else if (nc > 1 && valid == 1) {
            ++fails;
            solo = true;
        } else {
            fails = 0;
        }
        o = dst[valid];
        pos = ch[valid - 1].end_bit;
        if (ch[valid - 1].final_block) {
            after = base + (pos >> 3);
            crc = crc_all;
            return OK;
        }
    }
}

}  // namespace

extern "C" {

int nts_gz_inflate(const uint8_t* in, uint64_t n_in, uint8_t* out, uint64_t cap, uint64_t* n_out, int verify_crc) {
    return nts_gz_inflate_mt(in, n_in, out, cap, n_out, verify_crc, 1);
}

// members that can be found without decoding (BGZF: every member names its own compressed size): member i spans
// [member_off[i], member_off[i + 1]); its ISIZE trailer gives its place in the output, so the members are decoded
// independently, a contiguous range per thread
int nts_gz_inflate_members(const uint8_t* in, const uint64_t* member_off, uint64_t n_members, uint8_t* out, uint64_t cap,
                           uint64_t* n_out, int verify_crc, uint32_t n_threads) {
    using nts::fail;
    if (!in || !member_off || !n_out || (!out && cap)) return fail(NTS_ERR_ARG, "null argument");
    if (!n_threads) n_threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<uint64_t> at(n_members + 1, 0);
    for (uint64_t i = 0; i < n_members; ++i) {
        if (member_off[i + 1] < member_off[i] + 18) return fail(NTS_ERR_STATE, "truncated gzip stream");
        const uint8_t* t = in + member_off[i + 1] - 4;
        at[i + 1] = at[i] + (t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24));
    }
    *n_out = at[n_members];
    if (at[n_members] > cap) return 1;
    n_threads = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_threads, n_members / 16));
    std::vector<int> status(n_threads, NTS_OK);
    std::vector<std::string> message(n_threads);
    std::vector<std::thread> th;
    for (uint32_t k = 0; k < n_threads; ++k)
        th.emplace_back([&, k] {
            uint64_t lo = n_members * k / n_threads, hi = n_members * (k + 1) / n_threads;
            for (uint64_t i = lo; i < hi; ++i) {
                uint64_t got = 0;
                int rc = nts_gz_inflate_mt(in + member_off[i], member_off[i + 1] - member_off[i], out + at[i], at[i + 1] - at[i],
                                           &got, verify_crc, 1);
                if (rc == NTS_OK && got != at[i + 1] - at[i]) rc = 1;
                if (rc != NTS_OK) {
                    status[k] = rc;
                    message[k] = rc == 1 ? "gzip length check failed" : nts_last_error();
                    return;
                }
            }
        });
    for (auto& x : th) x.join();
    for (uint32_t k = 0; k < n_threads; ++k)
        if (status[k] != NTS_OK) return fail(NTS_ERR_STATE, message[k]);
    return NTS_OK;
}

int nts_gz_inflate_mt(const uint8_t* in, uint64_t n_in, uint8_t* out, uint64_t cap, uint64_t* n_out, int verify_crc, uint32_t n_threads) {
    using nts::fail;
    if (!n_threads) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = std::min(n_threads, 32u);
    // chunk of compressed bytes per thread and wave: 4 MB keeps the boundary search a small part of a chunk's work; files
    // of tens of megabytes get smaller chunks so that every thread has one (NTS_GZ_CHUNK_BYTES: fixed size, for tests)
    uint64_t chunk_bytes = std::min<uint64_t>(4ull << 20, std::max<uint64_t>(1ull << 20, n_in / (4ull * n_threads)));
    uint64_t parallel_from = 16ull << 20;
    if (const char* e = getenv("NTS_GZ_CHUNK_BYTES")) {
        chunk_bytes = std::max<uint64_t>(4096, strtoull(e, nullptr, 10));
        parallel_from = 4 * chunk_bytes;
    }
    if ((!in && n_in) || (!out && cap) || !n_out) return fail(NTS_ERR_ARG, "null argument");
    *n_out = 0;
    std::vector<Tables> tables(1);
    Tables& t = tables[0];
    const uint8_t* p = in;
    const uint8_t* end = in + n_in;
    uint8_t* o = out;
    uint8_t* o_end = out + cap;
    int members = 0;
    while (p < end) {
        if (members && *p == 0) { ++p; continue; }              // zero padding after a member is legal
        if (end - p < 18) return fail(NTS_ERR_STATE, "truncated gzip stream");
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return fail(NTS_ERR_STATE, "not a gzip stream");
        uint32_t flg = p[3];
        p += 10;
        if (flg & 4) {
            if (end - p < 2) return fail(NTS_ERR_STATE, "truncated gzip stream");
            uint32_t xlen = p[0] | (p[1] << 8);
            if ((uint64_t)(end - p) < 2 + (uint64_t)xlen) return fail(NTS_ERR_STATE, "truncated gzip stream");
            p += 2 + xlen;
        }
        for (uint32_t f : {8u, 16u})                            // file name, comment: zero-terminated
            if (flg & f) {
                while (p < end && *p) ++p;
                if (p == end) return fail(NTS_ERR_STATE, "truncated gzip stream");
                ++p;
            }
        if (flg & 2) {
            if (end - p < 2) return fail(NTS_ERR_STATE, "truncated gzip stream");
            p += 2;
        }
        Reader r;
        r.p = p;
        r.end = end;
        uint8_t* member_start = o;
        int rc;
        uint32_t crc = 0;
        if (n_threads > 1 && (uint64_t)(end - p) >= parallel_from) {
            const uint8_t* after = nullptr;
            rc = inflate_member_parallel(in, end, (uint64_t)(p - in), member_start, o, o_end, (int)n_threads, chunk_bytes,
                                         verify_crc != 0, crc, after);
            if (rc == OK) r.p = after;
        } else if (verify_crc && (uint64_t)(end - p) < (1ull << 20)) {
            rc = inflate_stream(r, t, member_start, o, o_end, nullptr);        // small member: not worth a second thread
            if (rc == OK) crc = crc32_update(0, member_start, (uint64_t)(o - member_start));
        } else if (verify_crc) {
            CrcFollower follower(member_start);         // reads what the decoder has published as final, block by block
            rc = inflate_stream(r, t, member_start, o, o_end, &follower.done);
            crc = follower.finish(rc == OK ? (uint64_t)(o - member_start) : follower.done.load());
        } else {
            rc = inflate_stream(r, t, member_start, o, o_end, nullptr);
        }
        *n_out = (uint64_t)(o - out);
        if (rc == OUTPUT_FULL) return 1;
        if (rc == TRUNCATED) return fail(NTS_ERR_STATE, "truncated gzip stream");
        if (rc != OK) return fail(NTS_ERR_STATE, "corrupt deflate data");
        p = r.p;
        if (end - p < 8) return fail(NTS_ERR_STATE, "truncated gzip stream");
        uint32_t want_crc = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        uint32_t isize = p[4] | (p[5] << 8) | (p[6] << 16) | ((uint32_t)p[7] << 24);
        p += 8;
        if (isize != (uint32_t)((uint64_t)(o - member_start) & 0xFFFFFFFFu)) return fail(NTS_ERR_STATE, "gzip length check failed");
        if (verify_crc && crc != want_crc) return fail(NTS_ERR_STATE, "gzip CRC check failed");
        ++members;
    }
    if (!members) return fail(NTS_ERR_STATE, "empty gzip input");
    return NTS_OK;
}

}  // extern "C"
