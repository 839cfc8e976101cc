// Parallel decoder for ONE deflate stream (a plain `gzip` / `pigz` .fastq.gz: one member, no block-size
// headers) — SURVEY.md §8(f) N1.  The reference reads such files with a single isa-l inflate thread
// (FastxReader, T.cpp:601-640); a single stream caps the host at one core whatever the GPU does.
//
// Two-pass scheme (the published idea of Kerbiriou & Chikhi's "pugz", restated here from the deflate
// format itself):
//   1. The compressed range is cut into spans.  A worker *searches* its span for the first bit position
//      that parses as a dynamic-Huffman block whose codes are complete and whose literals are all text
//      (FASTQ/FASTA bytes), and decodes from there into 16-bit symbols.  The 32 KB of history it cannot
//      know is represented by placeholders 256 + j (j = position in the unknown window), which
//      match copies propagate like ordinary symbols.  It stops at the first block boundary at or after
//      the start of the next span.
//   2. Spans are validated in stream order: span i is accepted only if it starts at exactly the bit
//      where span i-1 ended (span 0 starts at the member's first block), so an accepted chain IS the
//      sequential decode — speculation only ever decides how fast the answer comes.  A span that does not
//      line up (block search fooled, stored/fixed block at the boundary, non-text data) is decoded again
//      from the known boundary — as plain bytes, since its window is known by then (so are span 0 and, once
//      the search has failed twice or four spans in a row had to be redone, every span: the speed of the
//      sequential decoder is the floor).  With the previous span's last 32 KB resolved, the placeholders are
//      replaced through a 33 K-entry lookup table and the span's CRC-32 is taken, both on the workers;
//      the consumer combines the CRCs (crc32_combine) and checks the member trailer.
// A large member behind this one (`cat a.fq.gz b.fq.gz`) gets a reader of its own; small remainders and
// padding go to the streaming GzReader.
// tests/cpp/pinflate_check.cpp pins it against the sequential GzReader (itself pinned against zlib by inflate_check.cpp).
#pragma once
#include "inflate.hpp"

#include <atomic>
#include <chrono>

namespace fastgz {

namespace pinf {

enum : size_t { HIST = 32768, SLACK = 32, MARGIN = 12 + 4 + 258 + SLACK };

struct BitIn {
    const uint8_t *base = nullptr, *in = nullptr, *end = nullptr;
    uint64_t bb = 0;
    unsigned bc = 0, overrun = 0;

    void seek(const uint8_t *b, size_t n, uint64_t bit) {
        base = b;
        end = b + n;
        in = b + std::min<uint64_t>(bit >> 3, n);
        bb = 0;
        bc = overrun = 0;
        refill();
        const unsigned r = (unsigned)(bit & 7);
        bb >>= r;
        bc -= r;
    }
    inline void refill() {
        if (end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            bb |= w << bc;
            in += (63 - bc) >> 3;
            bc |= 56;
        } else {
            while (bc <= 56) {
                if (in < end) bb |= (uint64_t)*in++ << bc;
                else ++overrun; // zero bits beyond the end
                bc += 8;
            }
        }
    }
    inline uint32_t take(unsigned n) {
        const uint32_t v = (uint32_t)(bb & ((1ull << n) - 1));
        bb >>= n;
        bc -= n;
        return v;
    }
    // bits consumed so far, counted from base
    uint64_t bitpos() const { return 8ull * (uint64_t)(in - base + overrun) - bc; }
};

// Kraft sum of a code scaled to 2^15: == 32768 complete, < incomplete, > over-subscribed
inline unsigned kraft15(const uint8_t *lens, unsigned n, unsigned *used = nullptr) {
    unsigned s = 0, u = 0;
    for (unsigned i = 0; i < n; ++i)
        if (lens[i]) { s += 1u << (15 - lens[i]); ++u; }
    if (used) *used = u;
    return s;
}

// Code lengths of a dynamic block header (RFC 1951 3.2.7); the three header bits are already consumed.
// strict: what the block search demands on top of validity — complete code-length and literal/length
// codes, a distance code that is empty, single or complete (what every real compressor emits).
inline const char *read_code_lengths(BitIn &b, uint8_t *lens, unsigned &hlit, unsigned &hdist, bool strict) {
    b.refill();
    hlit = b.take(5) + 257;
    hdist = b.take(5) + 1;
    const unsigned hclen = b.take(4) + 4;
    if (hlit > 286 || hdist > 30) return "corrupt dynamic block header";
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) {
        if (b.bc < 3) b.refill();
        cl[order[i]] = (uint8_t)b.take(3);
    }
    if (strict) {
        unsigned used;
        const unsigned k = kraft15(cl, 19, &used);
        if (k != 32768u) return "incomplete code-length code";
    }
    Entry clt[128];
    if (!build_table(cl, 19, 7, clt, 128, 2)) return "corrupt code-length code";
    unsigned i = 0;
    while (i < hlit + hdist) {
        b.refill();
        const Entry e = clt[b.bb & 127];
        if (e.op & 0x80) return "corrupt code-length code";
        b.take(e.nbits);
        const unsigned sym = e.val;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        unsigned rep, val = 0;
        if (sym == 16) {
            if (i == 0) return "corrupt code lengths";
            val = lens[i - 1];
            rep = 3 + b.take(2);
        } else if (sym == 17) rep = 3 + b.take(3);
        else rep = 11 + b.take(7);
        if (i + rep > hlit + hdist) return "corrupt code lengths";
        while (rep--) lens[i++] = (uint8_t)val;
    }
    if (b.overrun > 8) return "truncated deflate stream";
    if (lens[256] == 0) return "no end-of-block code";
    if (strict) {
        if (kraft15(lens, hlit) != 32768u) return "incomplete literal/length code";
        unsigned used;
        const unsigned k = kraft15(lens + hlit, hdist, &used);
        if (used > 1 && k != 32768u) return "incomplete distance code";
    }
    return nullptr;
}

// One decoder per worker thread: Huffman tables + bit cursor, symbols out as uint16_t.
class SymDecoder {
public:
    BitIn b;

    // First bit position in [from, to) that starts a plausible non-final dynamic block; (uint64_t)-1 if none.
    uint64_t find_block(const uint8_t *base, size_t size, uint64_t from, uint64_t to, const std::atomic<bool> *abort = nullptr) {
        const uint64_t last = size >= 16 ? 8ull * (size - 16) : 0; // the peek below reads 8 bytes
        to = std::min(to, last);
        for (uint64_t bit = from; bit < to; ++bit) {
            if ((bit & 8191) == 0 && abort && abort->load(std::memory_order_relaxed)) break;
            uint64_t w;
            memcpy(&w, base + (bit >> 3), 8);
            w >>= bit & 7;
            // BFINAL = 0, BTYPE = 10b (LSB first: bits 1..2 = 0,1), HLIT <= 29, HDIST <= 29
            if ((w & 7) != 4 || ((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29) continue;
            // Kraft sum of the code-length code over the lengths the 56 valid bits of w hold (13 of up to 19):
            // over-subscribed -> reject; all lengths seen -> must be complete
            const unsigned hclen = (unsigned)((w >> 13) & 15) + 4, seen = std::min(hclen, 13u);
            unsigned kraft = 0;
            uint64_t c = w >> 17;
            for (unsigned i = 0; i < seen; ++i, c >>= 3) kraft += (128u >> (c & 7)) & 127u; // length 0 adds nothing
            if (kraft > 128 || (hclen <= 13 && kraft != 128)) continue;
            if (plausible_block(base, size, bit)) return bit;
        }
        return (uint64_t)-1;
    }

    // Block header at the cursor.  type: 0 stored, 1/2 Huffman (tables built).  nullptr on success.
    const char *block_header(bool &final, unsigned &type) {
        b.refill();
        final = b.take(1) != 0;
        type = b.take(2);
        if (type == 3) return "invalid deflate block type";
        if (type == 1) {
            uint8_t lens[320];
            int i = 0;
            for (; i < 144; ++i) lens[i] = 8;
            for (; i < 256; ++i) lens[i] = 9;
            for (; i < 280; ++i) lens[i] = 7;
            for (; i < 288; ++i) lens[i] = 8;
            build_table(lens, 288, LIT_BITS, lit_, LIT_TABLE, 0);
            build_multi_literal(lit_, multi_);
            for (i = 0; i < 32; ++i) lens[i] = 5;
            build_table(lens, 32, DIST_BITS, dist_, DIST_TABLE, 1);
        } else if (type == 2) {
            uint8_t lens[320 + 140];
            unsigned hlit, hdist;
            if (const char *e = read_code_lengths(b, lens, hlit, hdist, false)) return e;
            if (!build_table(lens, (int)hlit, LIT_BITS, lit_, LIT_TABLE, 0) ||
                !build_table(lens + hlit, (int)hdist, DIST_BITS, dist_, DIST_TABLE, 1))
                return "over-subscribed Huffman code";
            build_multi_literal(lit_, multi_);
        }
        return nullptr;
    }

    // Stored block body (cursor just behind the 3 header bits) copied (widened for 16-bit symbols) into out.
    // Needs room for 65535 symbols.
    template <typename T>
    const char *stored_block(T *&out) {
        b.take(b.bc & 7);
        const uint64_t byte = b.bitpos() >> 3;
        const size_t size = (size_t)(b.end - b.base);
        if (b.overrun || byte + 4 > size) return "truncated stored block";
        const uint8_t *p = b.base + byte;
        const uint32_t len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
        if ((len ^ 0xFFFFu) != nlen) return "corrupt stored block";
        if (byte + 4 + len > size) return "truncated stored block";
        for (uint32_t i = 0; i < len; ++i) out[i] = p[4 + i];
        out += len;
        b.seek(b.base, size, 8ull * (byte + 4 + len));
        return nullptr;
    }

    // Huffman block body.  1: end of block, 0: out reached out_lim (call again with more room), -1: error (err set).
    // out_lim must leave MARGIN symbols of room behind it; out - 32768 must be addressable.
    // T = uint16_t: symbols, the unknown window in front of the span is placeholders (no distance check needed);
    // T = uint8_t: bytes behind a known window, a match may not reach below `floor` (first valid byte).
    template <typename T>
    int huffman_block(T *&out_ref, T *out_lim, const T *floor, const char *&err) {
        T *out = out_ref;
        uint64_t bb = b.bb;
        unsigned bc = b.bc;
        const uint8_t *in = b.in;
        const uint8_t *const end = b.end;
        const Entry *const lit = lit_, *const dist = dist_;
        const uint32_t *const multi = multi_;
        int ret = 0;
#define PGZ_REFILL()                                                   \
    do {                                                               \
        if (end - in >= 8) {                                           \
            uint64_t w_;                                               \
            memcpy(&w_, in, 8);                                        \
            bb |= w_ << bc;                                            \
            in += (63 - bc) >> 3;                                      \
            bc |= 56;                                                  \
        } else {                                                       \
            while (bc <= 56) {                                         \
                if (in < end) bb |= (uint64_t)*in++ << bc;             \
                else ++b.overrun;                                      \
                bc += 8;                                               \
            }                                                          \
        }                                                              \
    } while (0)
        while (true) {
            if (out >= out_lim) { ret = 0; break; }
            PGZ_REFILL();
            uint32_t m = multi[bb & ((1u << LIT_BITS) - 1)];
            if (m >> 31) { // packs of 1-3 literals, widened to four 16-bit symbols with one store
#define PGZ_EMIT()                                                                                          \
    do {                                                                                                    \
        const uint64_t l_ = m & 0xFFFFFFu;                                                                  \
        if (sizeof(T) == 2) {                                                                               \
            const uint64_t w_ = (l_ & 0xFF) | ((l_ & 0xFF00) << 8) | ((l_ & 0xFF0000) << 16);               \
            memcpy(out, &w_, 8);                                                                            \
        } else {                                                                                            \
            const uint32_t w_ = (uint32_t)l_;                                                               \
            memcpy(out, &w_, 4);                                                                            \
        }                                                                                                   \
        out += (m >> 28) & 3u;                                                                              \
        const unsigned nb_ = (m >> 24) & 15u;                                                               \
        bb >>= nb_;                                                                                         \
        bc -= nb_;                                                                                          \
    } while (0)
                PGZ_EMIT();
                m = multi[bb & ((1u << LIT_BITS) - 1)];
                if (m >> 31) {
                    PGZ_EMIT();
                    m = multi[bb & ((1u << LIT_BITS) - 1)];
                    if (m >> 31) {
                        PGZ_EMIT();
                        m = multi[bb & ((1u << LIT_BITS) - 1)];
                        if (m >> 31) {
                            PGZ_EMIT();
                            continue;
                        }
                    }
                }
#undef PGZ_EMIT
                PGZ_REFILL();
                m = multi[bb & ((1u << LIT_BITS) - 1)];
            }
            Entry e;
            if (m >> 31) {
                e = lit[bb & ((1u << LIT_BITS) - 1)];
            } else {
                e.val = (uint16_t)m;
                e.nbits = (uint8_t)(m >> 16);
                e.op = (uint8_t)(((m >> 24) & 0x7F) | ((m >> 30) & 1u ? 0x80 : 0));
            }
            if (e.op & 0x20) {
                bb >>= e.nbits; bc -= e.nbits;
                e = lit[e.val + (bb & ((1u << (e.op & 15)) - 1))];
            }
            bb >>= e.nbits; bc -= e.nbits;
            if (e.op == 0) { *out++ = (T)e.val; continue; }
            if (e.op & 0x40) { ret = 1; break; }
            if (!(e.op & 0x10)) { err = "invalid literal/length code"; ret = -1; break; }
            unsigned len = e.val, xb = e.op & 15;
            len += (unsigned)(bb & ((1u << xb) - 1));
            bb >>= xb; bc -= xb;
            if (bc < 32) PGZ_REFILL();
            Entry d = dist[bb & ((1u << DIST_BITS) - 1)];
            if (d.op & 0x20) {
                bb >>= d.nbits; bc -= d.nbits;
                d = dist[d.val + (bb & ((1u << (d.op & 15)) - 1))];
            }
            bb >>= d.nbits; bc -= d.nbits;
            if (!(d.op & 0x10)) { err = "invalid distance code"; ret = -1; break; }
            xb = d.op & 15;
            const unsigned distance = d.val + (unsigned)(bb & ((1u << xb) - 1)); // <= 32768 by the tables
            bb >>= xb; bc -= xb;
            if (sizeof(T) == 1 && distance > (size_t)(out - floor)) { err = "invalid match distance"; ret = -1; break; }
            const T *src = out - distance;
            T *const stop = out + len;
            if (distance >= 16 / sizeof(T)) { // 16-byte copies; may write up to 16 bytes past stop (SLACK)
                memcpy(out, src, 16);
                memcpy(out + 16 / sizeof(T), src + 16 / sizeof(T), 16);
                if (len > 32 / sizeof(T)) {
                    out += 32 / sizeof(T);
                    src += 32 / sizeof(T);
                    do { memcpy(out, src, 16); out += 16 / sizeof(T); src += 16 / sizeof(T); } while (out < stop);
                }
            } else if (distance == 1) { // a run
                if (sizeof(T) == 1) {
                    memset(out, (int)*src, len);
                } else {
                    const uint64_t v4 = (uint64_t)*src * 0x0001000100010001ull;
                    do { memcpy(out, &v4, 8); memcpy(out + 4, &v4, 8); out += 8; } while (out < stop);
                }
            } else {
                do { *out++ = *src++; } while (out < stop);
            }
            out = stop;
        }
#undef PGZ_REFILL
        if (b.overrun > 8 && ret >= 0) { err = "truncated deflate stream"; ret = -1; }
        out_ref = out;
        b.bb = bb;
        b.bc = bc;
        b.in = in;
        return ret;
    }

private:
    // Full check of a candidate position: strict header, then a walk over the block's symbols — valid codes,
    // text literals only, an end-of-block symbol before the input runs out, a sane block type behind it.
    bool plausible_block(const uint8_t *base, size_t size, uint64_t bit) {
        BitIn c;
        c.seek(base, size, bit + 3);
        uint8_t lens[320 + 140];
        unsigned hlit, hdist;
        if (read_code_lengths(c, lens, hlit, hdist, true)) return false;
        if (!build_table(lens, (int)hlit, LIT_BITS, lit_, LIT_TABLE, 0) ||
            !build_table(lens + hlit, (int)hdist, DIST_BITS, dist_, DIST_TABLE, 1))
            return false;
        static const struct Text {
            bool ok[256];
            Text() {
                for (int i = 0; i < 256; ++i) ok[i] = (i >= 32 && i < 127) || i == '\n' || i == '\r' || i == '\t';
            }
        } text;
        for (unsigned n = 0; n < (1u << 22); ++n) { // blocks beyond 4 M symbols are accepted on what was seen
            c.refill();
            Entry e = lit_[c.bb & ((1u << LIT_BITS) - 1)];
            if (e.op & 0x20) {
                c.take(e.nbits);
                e = lit_[e.val + (c.bb & ((1u << (e.op & 15)) - 1))];
            }
            c.take(e.nbits);
            if (e.op == 0) {
                if (!text.ok[e.val]) return false;
                continue;
            }
            if (e.op & 0x40) {
                if (c.overrun) return false;
                c.refill();
                return ((c.bb >> 1) & 3) != 3; // the next header's BTYPE
            }
            if (!(e.op & 0x10)) return false;
            c.take(e.op & 15);
            if (c.bc < 32) c.refill();
            Entry d = dist_[c.bb & ((1u << DIST_BITS) - 1)];
            if (d.op & 0x20) {
                c.take(d.nbits);
                d = dist_[d.val + (c.bb & ((1u << (d.op & 15)) - 1))];
            }
            c.take(d.nbits);
            if (!(d.op & 0x10)) return false;
            c.take(d.op & 15);
            if (c.overrun > 8) return false;
        }
        return true;
    }

    Entry lit_[LIT_TABLE], dist_[DIST_TABLE];
    uint32_t multi_[1 << LIT_BITS];
};

// length of the gzip member header at p, 0 if it is not one (RFC 1952)
inline size_t gzip_header_len(const uint8_t *p, size_t n) {
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return 0;
    const uint8_t flg = p[3];
    size_t q = 10;
    if (flg & 4) {
        if (q + 2 > n) return 0;
        q += 2 + (size_t)(p[q] | (p[q + 1] << 8));
    }
    for (int f = 8; f <= 16; f <<= 1)
        if (flg & f) {
            while (q < n && p[q]) ++q;
            ++q;
        }
    if (flg & 2) q += 2;
    return q < n ? q : 0;
}

}  // namespace pinf

class SingleStreamReader {
public:
    // worth it: a file of at least a few spans (TGSF_PINFLATE_MIN_BYTES: tests)
    static bool worthwhile(const std::string &path) {
        struct stat st;
        const char *e = getenv("TGSF_PINFLATE_MIN_BYTES");
        const size_t min_bytes = e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(4u << 20);
        return stat(path.c_str(), &st) == 0 && (size_t)st.st_size >= min_bytes;
    }

    SingleStreamReader(const std::string &path, int threads, size_t span_bytes = 0) {
        fd_ = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) return;
        size_ = (size_t)st.st_size;
        if (size_) {
            void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) return;
            base_ = (const uint8_t *)m;
        }
        start(threads, span_bytes);
    }
    // a member inside a mapped file (the members after the first one of a `cat`-ed file)
    SingleStreamReader(const uint8_t *data, size_t n, int threads, size_t span_bytes) : base_(data), size_(n), borrowed_(true) {
        start(threads, span_bytes);
    }
    ~SingleStreamReader() {
        shutdown();
        for (auto &p : sym_pool_) delete[] p.first;
        if (base_ && !borrowed_) munmap((void *)base_, size_);
        if (fd_ >= 0) close(fd_);
    }
    bool ok() const { return ok_; }
    bool failed() const { return error_ != nullptr; }
    const char *error() const { return error_; }
    size_t repairs() const { return repairs_ + (next_ ? next_->repairs() : 0); } // spans decoded a second time (tests, diagnostics)
    size_t spans() const { return nspans_ + (next_ ? next_->spans() : 0); }
    // spans accepted as decoded speculatively, i.e. against an unknown window (tests, diagnostics)
    size_t speculated() const { return speculated_ + (next_ ? next_->speculated() : 0); }
    // worker seconds spent in block search / symbolic decode / placeholder resolution + CRC (diagnostics)
    void phase_seconds(double out[3]) const { for (int i = 0; i < 3; ++i) out[i] = phase_ns_[i].load() * 1e-9; }

    size_t read(void *dst, size_t n) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < n && !error_) {
            if (next_) {
                const size_t k = next_->read(d + got, n - got);
                if (k == 0) { if (next_->failed()) error_ = next_->error(); break; }
                got += k;
                continue;
            }
            if (tail_) {
                const size_t k = tail_->read(d + got, n - got);
                if (k == 0) { if (tail_->failed()) error_ = tail_->error(); break; }
                got += k;
                continue;
            }
            if (finished_) break;
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return !tasks_.empty() ? tasks_.front()->state == READY : scan_done_; });
            if (tasks_.empty()) { // cannot happen before the final block or an error was delivered
                error_ = "truncated deflate stream";
                break;
            }
            Task &t = *tasks_.front();
            lk.unlock();
            if (!t.discard) {
                const size_t k = std::min(n - got, t.nout - t.pos);
                memcpy(d + got, t.bytes + t.pos, k);
                t.pos += k;
                got += k;
                if (t.pos < t.nout) continue;
                crc_ = (uint32_t)crc32_combine(crc_, t.crc, (z_off_t)t.nout);
                total_ += t.nout;
                if (t.err) { error_ = t.err; break; }
                if (t.final) { finish_member(t.end_bit); continue; }
            }
            release_sym(t);
            lk.lock();
            tasks_.pop_front();
            ++front_idx_;
            lk.unlock();
            cv_.notify_all();
        }
        return got;
    }

private:
    enum { QUEUED = 0, DECODING, DECODED, RESOLVABLE, RESOLVING, READY };
    enum : uint64_t { NOT_FOUND = ~0ull - 2, SEARCHING = ~0ull - 1, UNKNOWN = ~0ull }; // found_[] beyond bit positions
    struct Task {
        size_t idx = 0;
        uint64_t stop_bit = 0, start_bit = 0, end_bit = 0;
        bool known_start = false, found = false, final = false, discard = false;
        bool direct = false; // decoded as bytes behind a known window (no placeholders)
        const uint8_t *bytes = nullptr; // the span's output once READY
        std::atomic<bool> cancel{false};
        const char *err = nullptr;
        uint16_t *sym = nullptr; // HIST placeholders, then the decoded symbols
        size_t sym_cap = 0, nsym = 0;
        std::unique_ptr<uint8_t[]> out;
        size_t nout = 0, pos = 0;
        std::vector<uint8_t> win; // the HIST bytes in front of this span
        size_t win_valid = 0;
        uint32_t crc = 0;
        int state = QUEUED;
        ~Task() { delete[] sym; }
    };

    void start(int threads, size_t span_bytes) {
        threads_ = std::max(1, threads);
        span_arg_ = span_bytes;
        data0_ = size_ ? pinf::gzip_header_len(base_, size_) : 0;
        ok_ = true;
        if (!data0_) { // empty or not a gzip member: the streaming reader reports it
            tail_.reset(new GzReader(base_, size_));
            return;
        }
        span_ = span_bytes ? span_bytes : std::min<size_t>(2u << 20, std::max<size_t>(128u << 10, (size_ - data0_) / (4 * (size_t)threads_)));
        nspans_ = nspans0_ = (size_ - data0_ + span_ - 1) / span_;
        if (const char *e = getenv("TGSF_PINFLATE_SOFT_CAP")) soft_cap_ = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10)); // tests
        found_.reset(new std::atomic<uint64_t>[nspans_ + 1]);
        for (size_t i = 0; i <= nspans_; ++i) found_[i].store(UNKNOWN);
        window_ = (size_t)threads_ + 3;
        chain_end_bit_ = 8ull * data0_;
        chain_window_.assign(pinf::HIST, 0);
        for (int t = 0; t < threads_; ++t) workers_.emplace_back([this] { work(); });
    }

    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            abort_search_.store(true);
            for (auto &t : tasks_) t->cancel.store(true);
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
        workers_.clear();
        for (auto &t : tasks_) release_sym(*t);
    }

    void finish_member(uint64_t end_bit) { // consumer: trailer check, then whatever follows the member
        shutdown();
        tasks_.clear();
        size_t off = (size_t)((end_bit + 7) >> 3);
        if (size_ - off < 8) { error_ = "truncated gzip trailer"; return; }
        uint32_t crc, isize;
        memcpy(&crc, base_ + off, 4);
        memcpy(&isize, base_ + off + 4, 4);
        off += 8;
        if (crc != crc_) { error_ = "gzip CRC mismatch"; return; }
        if (isize != (uint32_t)total_) { error_ = "gzip length mismatch"; return; }
        while (off < size_ && base_[off] == 0) ++off; // zero padding between / after members
        if (off >= size_) finished_ = true;
        else if (size_ - off >= (1u << 20) && pinf::gzip_header_len(base_ + off, size_ - off))
            next_.reset(new SingleStreamReader(base_ + off, size_ - off, threads_, span_arg_)); // a large member follows
        else tail_.reset(new GzReader(base_ + off, size_ - off));
    }

    // ---- symbol buffers are recycled: a span's buffer is tens of MB and fresh pages are slow ----
    void acquire_sym(Task &t, size_t cap) { // m_ not held
        {
            std::lock_guard<std::mutex> lk(pool_m_);
            for (size_t i = 0; i < sym_pool_.size(); ++i)
                if (sym_pool_[i].second >= cap) {
                    t.sym = sym_pool_[i].first;
                    t.sym_cap = sym_pool_[i].second;
                    sym_pool_.erase(sym_pool_.begin() + (long)i);
                    return;
                }
        }
        t.sym = new uint16_t[cap];
        t.sym_cap = cap;
    }
    void release_sym(Task &t) {
        if (!t.sym) return;
        std::lock_guard<std::mutex> lk(pool_m_);
        if (sym_pool_.size() < 2 * window_) sym_pool_.emplace_back(t.sym, t.sym_cap);
        else delete[] t.sym;
        t.sym = nullptr;
        t.sym_cap = 0;
    }
    void grow_sym(Task &t, size_t used_bytes) {
        const size_t cap = t.sym_cap + t.sym_cap / 2;
        uint16_t *n = new uint16_t[cap];
        memcpy(n, t.sym, used_bytes);
        delete[] t.sym;
        t.sym = n;
        t.sym_cap = cap;
    }

    // First plausible block start of a span: searched once, by whichever worker needs it first (the span's
    // own decoder, or the decoder of the span in front that wants to know where to stop).  Uses D's tables.
    uint64_t span_start(size_t idx, pinf::SymDecoder &D) {
        std::atomic<uint64_t> &f = found_[idx];
        uint64_t v = f.load();
        if (v == UNKNOWN && f.compare_exchange_strong(v, SEARCHING)) {
            const uint64_t from = 8ull * (data0_ + idx * span_);
            // a span without a block start in it (blocks of several MB, one-block streams) is left to the chain
            v = D.find_block(base_, size_, from, from + 8ull * span_, &abort_search_);
            if (v == (uint64_t)-1) {
                v = NOT_FOUND;
                if (++not_found_ >= 2) { speculate_.store(false); abort_search_.store(true); } // searching does not pay here
            }
            f.store(v);
            return v;
        }
        while ((v = f.load()) == SEARCHING) std::this_thread::yield(); // the searcher is running, never blocked
        return v;
    }

    // ---- phase 1: search + symbolic decode of one span (worker, no lock) ----
    void decode_span(Task &t, pinf::SymDecoder &D) {
        using namespace pinf;
        t.err = nullptr;
        t.final = false;
        t.nsym = 0;
        const auto c0 = std::chrono::steady_clock::now();
        if (!t.known_start) {
            t.start_bit = span_start(t.idx, D);
            t.found = t.start_bit < NOT_FOUND;
            phase_ns_[0] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - c0).count();
            if (!t.found) return;
        }
        const auto c1 = std::chrono::steady_clock::now();
        if (!t.sym) acquire_sym(t, HIST + 5 * span_ + MARGIN);
        if (t.direct) decode_blocks<uint8_t>(t, D);
        else decode_blocks<uint16_t>(t, D);
        phase_ns_[1] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - c1).count();
    }

    // T = uint16_t: the window in front of the span is unknown (placeholders); T = uint8_t ("direct"): the chain
    // has already reached this span's start (span 0, a span decoded again, sequential mode), so its window is known
    // and the bytes are final as they are written — the speed of the sequential decoder.
    template <typename T>
    void decode_blocks(Task &t, pinf::SymDecoder &D) {
        using namespace pinf;
        T *buf = (T *)t.sym;
        size_t cap = t.sym_cap * sizeof(uint16_t) / sizeof(T);
        if (sizeof(T) == 2) for (size_t j = 0; j < HIST; ++j) buf[j] = (T)(256 + j);
        else memcpy(buf, t.win.data(), HIST);
        T *out = buf + HIST;
        auto grow = [&]() {
            const size_t used = (size_t)(out - buf);
            grow_sym(t, used * sizeof(T));
            buf = (T *)t.sym;
            cap = t.sym_cap * sizeof(uint16_t) / sizeof(T);
            out = buf + used;
        };
        D.b.seek(base_, size_, t.start_bit);
        bool final = false;
        while (!final) {
            const uint64_t pos = D.b.bitpos();
            if (pos >= t.stop_bit) {
                // Stop where the next span starts.  A boundary in front of that start is a block the search
                // did not accept (stored or fixed block — every pigz / Z_SYNC_FLUSH chunk ends with one —,
                // non-text literals): keep decoding, so that such blocks cost no second decode.
                if (!speculate_.load(std::memory_order_relaxed) || t.idx + 1 >= nspans0_) break;
                const uint64_t next = span_start(t.idx + 1, D);
                if (next >= NOT_FOUND || pos >= next) break;
            }
            // Memory bound for streams that expand enormously (a span of 2 MB can hold GBs of runs): close the
            // span at this boundary; the spans behind it no longer line up and are decoded from here in turn.
            if ((size_t)(out - buf) - HIST >= soft_cap_) break;
            if (t.cancel.load(std::memory_order_relaxed)) return;
            unsigned type;
            if ((t.err = D.block_header(final, type))) break;
            if (type == 0) {
                if ((size_t)(out - buf) + 65536 + MARGIN > cap) grow();
                if ((t.err = D.stored_block(out))) break;
                continue;
            }
            int r;
            while ((r = D.huffman_block<T>(out, buf + cap - MARGIN, buf + HIST - t.win_valid, t.err)) == 0) grow();
            if (r < 0) break;
        }
        t.final = final && !t.err;
        t.end_bit = D.b.bitpos();
        t.nsym = (size_t)(out - buf) - HIST;
    }

    // ---- chain validation (m_ held): accept spans in order, hand the window on ----
    void advance_chain() {
        using namespace pinf;
        while (chain_next_ >= front_idx_ && chain_next_ - front_idx_ < tasks_.size()) {
            Task &t = *tasks_[chain_next_ - front_idx_];
            if (t.state != DECODED) return;
            if (chain_closed_) { // behind the final block or an error: not part of the member
                release_sym(t);
                t.discard = true;
                t.state = READY;
                ++chain_next_;
                continue;
            }
            const bool lined_up = t.known_start ? t.start_bit == chain_end_bit_ : (t.found && t.start_bit == chain_end_bit_);
            if (!lined_up) { // decode again from the boundary the chain has proven
                t.known_start = true;
                t.start_bit = chain_end_bit_;
                t.direct = true; // the window is known now
                t.win = chain_window_;
                t.win_valid = chain_valid_;
                t.state = QUEUED;
                ++repairs_;
                if (++consecutive_repairs_ >= 4) { speculate_.store(false); abort_search_.store(true); } // the search does not work on this data
                return;
            }
            if (t.idx > 0 && !t.known_start) consecutive_repairs_ = 0;
            if (!t.direct) ++speculated_;
            t.win = chain_window_;
            t.win_valid = chain_valid_;
            // the window behind this span: last HIST bytes of (window | resolved symbols)
            if (t.nsym) {
                std::vector<uint8_t> w(HIST);
                if (t.direct) {
                    memcpy(w.data(), (const uint8_t *)t.sym + t.nsym, HIST); // last HIST bytes of (window | output)
                } else {
                    const uint16_t *s = t.sym + t.nsym; // the symbol that lands at w[0] (may lie in the placeholder prefix)
                    for (size_t j = 0; j < HIST; ++j) {
                        const uint16_t v = s[j];
                        w[j] = v < 256 ? (uint8_t)v : chain_window_[v - 256];
                    }
                }
                chain_window_.swap(w);
                chain_valid_ = std::min<size_t>(HIST, chain_valid_ + t.nsym);
            }
            chain_end_bit_ = t.end_bit;
            if (!t.final && !t.err && t.idx + 1 >= nspans_) { // the last span was closed early (soft cap): one more
                ++nspans_;
                scan_done_ = false;
            }
            if (t.final || t.err) {
                chain_closed_ = true;
                scan_done_ = true;
                abort_search_.store(true);
                for (auto &o : tasks_)
                    if (o->idx > t.idx) o->cancel.store(true);
            }
            t.state = RESOLVABLE;
            ++chain_next_;
        }
    }

    // ---- phase 2: placeholders -> bytes, CRC (worker, no lock) ----
    void resolve_span(Task &t) {
        using namespace pinf;
        const auto c0 = std::chrono::steady_clock::now();
        if (t.direct) { // bytes are final: only the CRC is left; the buffer is released when the consumer is done
            t.bytes = (const uint8_t *)t.sym + HIST;
            t.nout = t.nsym;
            t.crc = crc_of(t.bytes, t.nout);
            phase_ns_[2] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - c0).count();
            return;
        }
        t.out.reset(new uint8_t[t.nsym + 1]);
        const uint16_t *s = t.sym + HIST;
        uint8_t *o = t.out.get();
        size_t n = t.nsym;
        if (t.win_valid < HIST) { // start of the stream: a match may not reach in front of the first byte
            const unsigned first_ok = 256 + (unsigned)(HIST - t.win_valid);
            for (size_t j = 0; j < n; ++j)
                if (s[j] >= 256 && s[j] < first_ok) {
                    n = j;
                    t.err = "invalid match distance";
                    t.final = false;
                    break;
                }
        }
        std::vector<uint8_t> lut(256 + HIST);
        for (unsigned i = 0; i < 256; ++i) lut[i] = (uint8_t)i;
        memcpy(lut.data() + 256, t.win.data(), HIST);
        const uint8_t *L = lut.data();
        size_t j = 0;
        for (; j + 4 <= n; j += 4) {
            o[j] = L[s[j]];
            o[j + 1] = L[s[j + 1]];
            o[j + 2] = L[s[j + 2]];
            o[j + 3] = L[s[j + 3]];
        }
        for (; j < n; ++j) o[j] = L[s[j]];
        t.nout = n;
        t.bytes = o;
        t.crc = crc_of(o, n);
        release_sym(t);
        std::vector<uint8_t>().swap(t.win);
        phase_ns_[2] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - c0).count();
    }

    static uint32_t crc_of(const uint8_t *p, size_t n) { // crc32 takes a 32-bit length
        uint32_t c = 0;
        for (size_t done = 0; done < n;) {
            const size_t k = std::min<size_t>(n - done, 1u << 30);
            c = (uint32_t)crc32(c, p + done, (uInt)k);
            done += k;
        }
        return c;
    }

    Task *new_task(bool known) { // m_ held
        tasks_.emplace_back(new Task());
        Task *t = tasks_.back().get();
        t->idx = next_idx_++;
        t->stop_bit = t->idx + 1 >= nspans0_ ? (uint64_t)-1 : 8ull * (data0_ + (t->idx + 1) * span_);
        if (t->idx == 0 || known) {
            t->known_start = true;
            t->start_bit = t->idx == 0 ? 8ull * data0_ : chain_end_bit_;
            t->direct = true; // the chain is at this span's start: its window is known
            t->win = chain_window_;
            t->win_valid = chain_valid_;
        }
        if (next_idx_ >= nspans_) scan_done_ = true;
        t->state = DECODING;
        return t;
    }

    void work() {
        std::unique_ptr<pinf::SymDecoder> D(new pinf::SymDecoder());
        while (true) {
            Task *t = nullptr;
            bool resolve = false;
            {
                std::unique_lock<std::mutex> lk(m_);
                while (true) {
                    if (stop_) return;
                    for (auto &p : tasks_)
                        if (p->state == RESOLVABLE) { t = p.get(); resolve = true; break; }
                    if (!t)
                        for (auto &p : tasks_)
                            if (p->state == QUEUED) { t = p.get(); break; }
                    if (t) { t->state = resolve ? RESOLVING : DECODING; break; }
                    if (!scan_done_ && tasks_.size() < window_) {
                        const bool from_chain = !speculate_.load() || next_idx_ >= nspans0_; // start known only once the chain is there
                        if (!from_chain || next_idx_ == 0) { t = new_task(false); break; }
                        if (chain_next_ == next_idx_) { t = new_task(true); break; }
                    }
                    cv_.wait(lk);
                }
            }
            if (resolve) resolve_span(*t);
            else decode_span(*t, *D);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (resolve) t->state = READY;
                else {
                    t->state = DECODED;
                    advance_chain();
                }
            }
            cv_.notify_all();
        }
    }

    int fd_ = -1;
    const uint8_t *base_ = nullptr;
    size_t size_ = 0, data0_ = 0, span_ = 0, nspans_ = 0, nspans0_ = 0, window_ = 8;
    size_t soft_cap_ = 96u << 20; // symbols per span before it is closed early (192 MB of symbols)
    bool ok_ = false, stop_ = false, scan_done_ = false, finished_ = false;
    bool chain_closed_ = false;
    std::atomic<bool> speculate_{true}, abort_search_{false};
    std::atomic<unsigned> not_found_{0};
    std::unique_ptr<std::atomic<uint64_t>[]> found_;
    const char *error_ = nullptr;
    size_t next_idx_ = 0, front_idx_ = 0, chain_next_ = 0, repairs_ = 0, consecutive_repairs_ = 0, speculated_ = 0;
    uint64_t chain_end_bit_ = 0;
    std::vector<uint8_t> chain_window_;
    size_t chain_valid_ = 0;
    uint32_t crc_ = 0;
    uint64_t total_ = 0;
    std::deque<std::unique_ptr<Task>> tasks_;
    std::vector<std::thread> workers_;
    std::unique_ptr<GzReader> tail_;
    std::unique_ptr<SingleStreamReader> next_;
    int threads_ = 1;
    size_t span_arg_ = 0;
    bool borrowed_ = false;
    std::mutex m_, pool_m_;
    std::condition_variable cv_;
    std::vector<std::pair<uint16_t *, size_t>> sym_pool_;
    std::atomic<uint64_t> phase_ns_[3] = {{0}, {0}, {0}};
};

// MultiMemberReader's fallback (a member too large for one speculative span, a guessed header that was none):
// a remainder of at least 1 MB that starts with a gzip header is decoded in parallel as well.
struct ParallelTail : TailReader {
    SingleStreamReader r;
    ParallelTail(const uint8_t *p, size_t n, int threads) : r(p, n, threads, 0) {}
    size_t read(void *dst, size_t n) override { return r.read(dst, n); }
    bool failed() const override { return r.failed(); }
    const char *error() const override { return r.error(); }
};
namespace pinf {
static const bool tail_registered = (tail_factory() = [](const uint8_t *p, size_t n, int threads) -> TailReader * {
    if (threads < 2 || n < (1u << 20) || !gzip_header_len(p, n) || getenv("TGSF_SERIAL_INFLATE")) return nullptr;
    return new ParallelTail(p, n, threads);
}, true);
}

}  // namespace fastgz
