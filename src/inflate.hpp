// Streaming gzip decoder of the tgsfilter host ("gzip decode").
//
// The reference reads .gz input with isa-l's isal_inflate (FastxReader, T.cpp:601-640: 1 MiB chunks,
// multi-member aware); the system zlib this image offers inflates literal-heavy FASTQ streams at
// ~100-150 MB/s and would make the host the slowest stage by far.  This is an own table-driven inflate:
// 64-bit bit buffer refilled with one unaligned load, 11-bit primary table for literal/length codes
// (8-bit for distances) with second-level tables for longer codes, up to three literals per refill,
// word-wise match copies.  Whole .gz file mapped into memory; output is produced in chunks behind a
// 32 KB history window.  gzip framing per RFC 1952 (FEXTRA/FNAME/FCOMMENT/FHCRC, concatenated members,
// CRC-32 and ISIZE checked), deflate per RFC 1951.  tests/cpp/inflate_check.cpp pins it against zlib.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h> // crc32() only

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fastgz {

struct Entry {
    uint16_t val; // literal byte | length/distance base | subtable offset
    uint8_t nbits; // bits to consume (subtable entries: beyond the primary bits)
    uint8_t op;    // 0 literal; 0x10|extra: base + extra bits; 0x20|bits: subtable link; 0x40 end of block; 0x80 invalid
};

enum { LIT_BITS = 11, DIST_BITS = 8, LIT_TABLE = (1 << LIT_BITS) + 288 * 16, DIST_TABLE = (1 << DIST_BITS) + 32 * 128 };

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// Canonical Huffman decode table.  kind 0: literal/length alphabet, 1: distance alphabet, 2: code-length alphabet
// (plain symbols).  Returns false for an over-subscribed code; incomplete codes leave invalid entries.
inline bool build_table(const uint8_t *lens, int n, int primary, Entry *table, int table_cap, int kind) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    int maxlen = 15;
    while (maxlen > 0 && count[maxlen] == 0) --maxlen;
    const Entry invalid = {0, 1, 0x80};
    for (int i = 0; i < (1 << primary); ++i) table[i] = invalid;
    if (maxlen == 0) return true; // no codes at all (legal for distances when the block has only literals)
    // over-subscription check
    long left = 1;
    for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; }
    uint16_t next_code[16];
    {
        unsigned code = 0;
        for (int l = 1; l <= 15; ++l) { code = (code + (unsigned)count[l - 1]) << 1; next_code[l] = (uint16_t)code; }
    }
    const int sub_bits = maxlen > primary ? maxlen - primary : 0;
    int next_sub = 1 << primary;
    for (int sym = 0; sym < n; ++sym) {
        const int l = lens[sym];
        if (!l) continue;
        unsigned code = next_code[l]++;
        unsigned rev = 0;
        for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
        Entry e;
        if (kind == 0) {
            if (sym < 256) e = {(uint16_t)sym, 0, 0};
            else if (sym == 256) e = {0, 0, 0x40};
            else if (sym < 286) e = {kLenBase[sym - 257], 0, (uint8_t)(0x10 | kLenExtra[sym - 257])};
            else e = {0, 0, 0x80};
        } else if (kind == 1) {
            if (sym < 30) e = {kDistBase[sym], 0, (uint8_t)(0x10 | kDistExtra[sym])};
            else e = {0, 0, 0x80};
        } else {
            e = {(uint16_t)sym, 0, 0};
        }
        if (l <= primary) {
            e.nbits = (uint8_t)l;
            for (unsigned i = rev; i < (1u << primary); i += 1u << l) table[i] = e;
        } else {
            const unsigned pfx = rev & ((1u << primary) - 1);
            if (!(table[pfx].op & 0x20)) { // first long code under this prefix: open a subtable
                if (next_sub + (1 << sub_bits) > table_cap) return false;
                table[pfx] = {(uint16_t)next_sub, (uint8_t)primary, (uint8_t)(0x20 | sub_bits)};
                const Entry inv2 = {0, 1, 0x80};
                for (int i = 0; i < (1 << sub_bits); ++i) table[next_sub + i] = inv2;
                next_sub += 1 << sub_bits;
            }
            Entry *sub = table + table[pfx].val;
            e.nbits = (uint8_t)(l - primary);
            for (unsigned i = rev >> primary; i < (1u << sub_bits); i += 1u << (l - primary)) sub[i] = e;
        }
    }
    return true;
}

// Multi-literal view of a literal/length primary table: entry i packs up to three literals that are fully
// determined by the LIT_BITS index bits (FASTQ bases have 2-3 bit codes, so one lookup yields 2-3 bytes).
// bit 31: valid pack, bits 28-29: count, bits 24-27: bits to consume, bits 0-23: the literals, first one lowest.
// Other entries carry the Entry itself (val | nbits << 16 | op << 24; op 0x80 moved to bit 30), so the
// length / end-of-block / link dispatch needs no second load.
inline void build_multi_literal(const Entry *lit, uint32_t *multi) {
    for (unsigned i = 0; i < (1u << LIT_BITS); ++i) {
        const Entry e1 = lit[i];
        if (e1.op != 0) { multi[i] = (uint32_t)e1.val | ((uint32_t)e1.nbits << 16) | ((uint32_t)(e1.op & 0x7F) << 24) | ((e1.op & 0x80) ? (1u << 30) : 0u); continue; }
        unsigned n = e1.nbits, cnt = 1;
        uint32_t bytes = e1.val;
        const Entry e2 = lit[i >> n];
        if (e2.op == 0 && n + e2.nbits <= LIT_BITS) {
            bytes |= (uint32_t)e2.val << 8;
            n += e2.nbits;
            cnt = 2;
            const Entry e3 = lit[i >> n];
            if (e3.op == 0 && n + e3.nbits <= LIT_BITS) {
                bytes |= (uint32_t)e3.val << 16;
                n += e3.nbits;
                cnt = 3;
            }
        }
        multi[i] = 0x80000000u | (cnt << 28) | (n << 24) | bytes;
    }
}

class GzReader {
public:
    enum { HIST = 32768, CHUNK = 4 << 20, SLACK = 320 };

    explicit GzReader(const std::string &path) {
        fd_ = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) return;
        size_ = (size_t)st.st_size;
        if (size_) {
            void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) return;
            base_ = (const uint8_t *)m;
            madvise((void *)base_, size_, MADV_SEQUENTIAL);
        }
        init_memory();
    }
    // decode from a memory range (tests, BGZF block groups)
    GzReader(const uint8_t *data, size_t n) : base_(data), size_(n), borrowed_(true) { init_memory(); }
    // point the decoder at another memory range (keeps the buffers)
    void reset(const uint8_t *data, size_t n) {
        base_ = data;
        size_ = n;
        borrowed_ = true;
        done_ = in_member_ = last_block_ = false;
        error_ = nullptr;
        bitbuf_ = 0;
        bitcnt_ = overrun_ = 0;
        member_out_ = 0;
        stop_ = (size_t)-1;
        state_ = BLOCK_HEADER;
        out_ = rd_ = wr_ = crc_from_ = buf_.data() + HIST;
        in_ = base_;
        end_ = base_ + size_;
        if (size_ == 0) done_ = true;
    }
    ~GzReader() {
        if (base_ && !borrowed_) munmap((void *)base_, size_);
        if (fd_ >= 0) close(fd_);
    }
    bool ok() const { return ok_; }
    bool failed() const { return error_ != nullptr; }
    const char *error() const { return error_; }
    // stop at the first member boundary at or beyond this input offset (default: end of input)
    void set_stop(size_t off) { stop_ = off; }
    // input bytes consumed up to the last finished member (valid once read() has returned 0 without error)
    size_t consumed() const { return (size_t)(in_ - base_); }

    // Up to n decoded bytes into dst; 0 at the end of the stream or after an error (see failed()).
    size_t read(void *dst, size_t n) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < n) {
            if (rd_ == wr_) {
                if (done_ || error_) break;
                produce();
                if (rd_ == wr_) { if (done_ || error_) break; continue; }
            }
            const size_t take = std::min(n - got, (size_t)(wr_ - rd_));
            memcpy(d + got, rd_, take);
            rd_ += take;
            got += take;
        }
        return got;
    }

private:
    void init_memory() {
        if (buf_.empty()) buf_.resize(HIST + CHUNK + SLACK + 8);
        out_ = rd_ = wr_ = buf_.data() + HIST;
        in_ = base_;
        end_ = base_ + size_;
        ok_ = true;
        if (size_ == 0) done_ = true;
    }
    void fail(const char *what) { error_ = what; }

    // ---- bit reader ----
    inline void refill() {
        if (end_ - in_ >= 8) {
            uint64_t w;
            memcpy(&w, in_, 8);
            bitbuf_ |= w << bitcnt_;
            in_ += (63 - bitcnt_) >> 3;
            bitcnt_ |= 56;
        } else {
            while (bitcnt_ <= 56) {
                if (in_ < end_) bitbuf_ |= (uint64_t)*in_++ << bitcnt_;
                else ++overrun_; // zero bits beyond the end; checked by the callers that matter
                bitcnt_ += 8;
            }
        }
    }
    inline uint32_t take(unsigned n) { // n <= 32, enough bits present
        const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1));
        bitbuf_ >>= n;
        bitcnt_ -= n;
        return v;
    }
    void align_to_byte() { // drop the partial byte, hand unread whole bytes back to the input
        take(bitcnt_ & 7);
        unsigned back = bitcnt_ >> 3;
        while (back && overrun_) { --back; --overrun_; }
        in_ -= back;
        bitbuf_ = 0;
        bitcnt_ = 0;
    }

    // ---- gzip member framing ----
    bool start_member() {
        // zero padding after the last member is tolerated
        while (in_ < end_ && *in_ == 0) ++in_;
        if (in_ >= end_) { done_ = true; return false; }
        if (end_ - in_ < 18 || in_[0] != 0x1f || in_[1] != 0x8b || in_[2] != 8) { fail("not a gzip member"); return false; }
        const uint8_t flg = in_[3];
        const uint8_t *p = in_ + 10;
        if (flg & 4) { // FEXTRA
            if (end_ - p < 2) { fail("truncated gzip header"); return false; }
            const size_t xlen = p[0] | (p[1] << 8);
            p += 2;
            if ((size_t)(end_ - p) < xlen) { fail("truncated gzip header"); return false; }
            p += xlen;
        }
        for (int f = 8; f <= 16; f <<= 1) // FNAME, FCOMMENT: zero-terminated
            if (flg & f) {
                while (p < end_ && *p) ++p;
                if (p >= end_) { fail("truncated gzip header"); return false; }
                ++p;
            }
        if (flg & 2) p += 2; // FHCRC
        if (p >= end_) { fail("truncated gzip header"); return false; }
        in_ = p;
        bitbuf_ = 0;
        bitcnt_ = 0;
        crc_ = 0;
        member_out_ = 0;
        last_block_ = false;
        state_ = BLOCK_HEADER;
        in_member_ = true;
        return true;
    }
    bool finish_member() {
        align_to_byte();
        if (end_ - in_ < 8) { fail("truncated gzip trailer"); return false; }
        uint32_t crc, isize;
        memcpy(&crc, in_, 4);
        memcpy(&isize, in_ + 4, 4);
        in_ += 8;
        flush_crc();
        if (crc != crc_) { fail("gzip CRC mismatch"); return false; }
        if (isize != (uint32_t)member_out_) { fail("gzip length mismatch"); return false; }
        in_member_ = false;
        if ((size_t)(in_ - base_) >= stop_) done_ = true; // member boundary at or beyond the requested stop
        return true;
    }
    void flush_crc() { // bytes [crc_from_, out_) belong to the current member and are not summed yet
        if (out_ > crc_from_) crc_ = (uint32_t)crc32(crc_, crc_from_, (uInt)(out_ - crc_from_));
        crc_from_ = out_;
    }

    // Decode until the chunk is full or the stream ends; afterwards [rd_, wr_) is new output.
    void produce() {
        // slide: keep the last HIST bytes as history in front of the new chunk
        uint8_t *base = buf_.data();
        if (out_ > base + HIST) {
            const size_t keep = std::min((size_t)HIST, (size_t)(out_ - base));
            memmove(base + HIST - keep, out_ - keep, keep);
        }
        out_ = rd_ = wr_ = crc_from_ = base + HIST;
        uint8_t *const out_lim = base + HIST + CHUNK;
        while (!error_ && !done_ && out_ < out_lim) {
            if (!in_member_ && !start_member()) break;
            if (!decode(out_lim)) break;
        }
        if (in_member_) flush_crc();
        wr_ = out_;
    }

    enum State { BLOCK_HEADER, STORED, HUFFMAN, MEMBER_END };

    // Runs the block state machine; returns false when the output limit was reached or on error.
    bool decode(uint8_t *out_lim) {
        while (true) {
            switch (state_) {
            case BLOCK_HEADER: {
                if (last_block_) { state_ = MEMBER_END; break; }
                refill();
                last_block_ = take(1) != 0;
                const uint32_t type = take(2);
                if (type == 0) {
                    align_to_byte();
                    if (end_ - in_ < 4) { fail("truncated stored block"); return false; }
                    const uint32_t len = in_[0] | (in_[1] << 8), nlen = in_[2] | (in_[3] << 8);
                    if ((len ^ 0xFFFFu) != nlen) { fail("corrupt stored block"); return false; }
                    in_ += 4;
                    stored_left_ = len;
                    state_ = STORED;
                } else if (type == 1) {
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
                    state_ = HUFFMAN;
                } else if (type == 2) {
                    if (!read_dynamic_tables()) return false;
                    state_ = HUFFMAN;
                } else {
                    fail("invalid deflate block type");
                    return false;
                }
                break;
            }
            case STORED: {
                while (stored_left_) {
                    if (out_ >= out_lim) return false;
                    size_t n = std::min<size_t>(stored_left_, (size_t)(out_lim - out_));
                    if ((size_t)(end_ - in_) < n) { fail("truncated stored block"); return false; }
                    memcpy(out_, in_, n);
                    in_ += n;
                    out_ += n;
                    member_out_ += n;
                    stored_left_ -= (uint32_t)n;
                }
                state_ = BLOCK_HEADER;
                break;
            }
            case HUFFMAN: {
                const int r = decode_huffman(out_lim);
                if (r < 0) return false;      // error
                if (r == 0) return false;     // output limit reached inside the block
                state_ = BLOCK_HEADER;        // end-of-block symbol
                break;
            }
            case MEMBER_END:
                if (!finish_member()) return false;
                return true; // next member (if any) is started by produce()
            }
        }
    }

    bool read_dynamic_tables() {
        refill();
        const unsigned hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
        if (hlit > 286 || hdist > 30) { fail("corrupt dynamic block header"); return false; }
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for (unsigned i = 0; i < hclen; ++i) {
            if (bitcnt_ < 3) refill();
            cl[order[i]] = (uint8_t)take(3);
        }
        Entry clt[128];
        if (!build_table(cl, 19, 7, clt, 128, 2)) { fail("corrupt code-length code"); return false; }
        uint8_t lens[320 + 140];
        unsigned i = 0;
        while (i < hlit + hdist) {
            refill();
            const Entry e = clt[bitbuf_ & 127];
            if (e.op & 0x80) { fail("corrupt code-length code"); return false; }
            take(e.nbits);
            const unsigned sym = e.val;
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            unsigned rep, val = 0;
            if (sym == 16) {
                if (i == 0) { fail("corrupt code lengths"); return false; }
                val = lens[i - 1];
                rep = 3 + take(2);
            } else if (sym == 17) rep = 3 + take(3);
            else rep = 11 + take(7);
            if (i + rep > hlit + hdist) { fail("corrupt code lengths"); return false; }
            while (rep--) lens[i++] = (uint8_t)val;
        }
        if (overrun_ > 8) { fail("truncated deflate stream"); return false; }
        if (lens[256] == 0) { fail("no end-of-block code"); return false; }
        if (!build_table(lens, (int)hlit, LIT_BITS, lit_, LIT_TABLE, 0) ||
            !build_table(lens + hlit, (int)hdist, DIST_BITS, dist_, DIST_TABLE, 1)) {
            fail("over-subscribed Huffman code");
            return false;
        }
        build_multi_literal(lit_, multi_);
        return true;
    }

    // 1: end of block, 0: output limit reached, -1: error
    int decode_huffman(uint8_t *out_lim) {
        uint8_t *out = out_;
        uint64_t bb = bitbuf_;
        unsigned bc = bitcnt_;
        const uint8_t *in = in_;
        const uint8_t *const end = end_;
        const Entry *const lit = lit_, *const dist = dist_;
        const uint32_t *const multi = multi_;
        uint64_t produced0 = member_out_;
        uint8_t *const out0 = out;
        int ret = 0;
#define FGZ_REFILL()                                                                                   \
    do {                                                                                               \
        if (end - in >= 8) {                                                                           \
            uint64_t w_;                                                                               \
            memcpy(&w_, in, 8);                                                                        \
            bb |= w_ << bc;                                                                            \
            in += (63 - bc) >> 3;                                                                      \
            bc |= 56;                                                                                  \
        } else {                                                                                       \
            while (bc <= 56) {                                                                         \
                if (in < end) bb |= (uint64_t)*in++ << bc;                                             \
                else ++overrun_;                                                                       \
                bc += 8;                                                                               \
            }                                                                                          \
        }                                                                                              \
    } while (0)
        while (true) {
            if (out >= out_lim) { ret = 0; break; }
            FGZ_REFILL();
            uint32_t m = multi[bb & ((1u << LIT_BITS) - 1)];
            if (m >> 31) { // packs of 1-3 literals: four lookups (<= 44 bits) per refill
#define FGZ_EMIT()                                  \
    do {                                            \
        const uint32_t lits_ = m & 0xFFFFFFu;       \
        memcpy(out, &lits_, 4);                     \
        out += (m >> 28) & 3u;                      \
        const unsigned nb_ = (m >> 24) & 15u;       \
        bb >>= nb_;                                 \
        bc -= nb_;                                  \
    } while (0)
                FGZ_EMIT();
                m = multi[bb & ((1u << LIT_BITS) - 1)];
                if (m >> 31) {
                    FGZ_EMIT();
                    m = multi[bb & ((1u << LIT_BITS) - 1)];
                    if (m >> 31) {
                        FGZ_EMIT();
                        m = multi[bb & ((1u << LIT_BITS) - 1)];
                        if (m >> 31) {
                            FGZ_EMIT();
                            continue;
                        }
                    }
                }
#undef FGZ_EMIT
                FGZ_REFILL(); // a length/distance pair may need 48 bits
                m = multi[bb & ((1u << LIT_BITS) - 1)];
            }
            Entry e;
            if (m >> 31) { // (only after the refill above) a literal again: take the plain entry
                e = lit[bb & ((1u << LIT_BITS) - 1)];
            } else {
                e.val = (uint16_t)m;
                e.nbits = (uint8_t)(m >> 16);
                e.op = (uint8_t)(((m >> 24) & 0x7F) | ((m >> 30) & 1u ? 0x80 : 0));
            }
            if (e.op & 0x20) { // second-level table
                bb >>= e.nbits; bc -= e.nbits;
                e = lit[e.val + (bb & ((1u << (e.op & 15)) - 1))];
            }
            bb >>= e.nbits; bc -= e.nbits;
            if (e.op == 0) { *out++ = (uint8_t)e.val; continue; }
            if (e.op & 0x40) { ret = 1; break; }
            if (!(e.op & 0x10)) { fail("invalid literal/length code"); ret = -1; break; }
            unsigned len = e.val, xb = e.op & 15;
            len += (unsigned)(bb & ((1u << xb) - 1));
            bb >>= xb; bc -= xb;
            if (bc < 32) FGZ_REFILL(); // long literal/length code in front: make sure 15 + 13 bits are there
            Entry d = dist[bb & ((1u << DIST_BITS) - 1)];
            if (d.op & 0x20) {
                bb >>= d.nbits; bc -= d.nbits;
                d = dist[d.val + (bb & ((1u << (d.op & 15)) - 1))];
            }
            bb >>= d.nbits; bc -= d.nbits;
            if (!(d.op & 0x10)) { fail("invalid distance code"); ret = -1; break; }
            xb = d.op & 15;
            const unsigned distance = d.val + (unsigned)(bb & ((1u << xb) - 1));
            bb >>= xb; bc -= xb;
            const uint64_t avail = produced0 + (uint64_t)(out - out0);
            if (distance > avail || distance > HIST) { fail("invalid match distance"); ret = -1; break; }
            const uint8_t *src = out - distance;
            uint8_t *const stop = out + len;
            if (distance >= 8) { // word copies; may write up to 15 bytes past stop (SLACK)
                // short matches dominate FASTQ streams: two unconditional words, a loop only beyond 16 bytes
                memcpy(out, src, 8);
                memcpy(out + 8, src + 8, 8);
                if (len > 16) {
                    out += 16;
                    src += 16;
                    do { memcpy(out, src, 8); out += 8; src += 8; } while (out < stop);
                }
            } else if (distance == 1) {
                memset(out, *src, len);
            } else {
                do { *out++ = *src++; } while (out < stop);
            }
            out = stop;
        }
#undef FGZ_REFILL
        if (overrun_ > 8 && ret >= 0) { fail("truncated deflate stream"); ret = -1; }
        member_out_ = produced0 + (uint64_t)(out - out0);
        out_ = out;
        bitbuf_ = bb;
        bitcnt_ = bc;
        in_ = in;
        return ret;
    }

    int fd_ = -1;
    const uint8_t *base_ = nullptr;
    size_t size_ = 0, stop_ = (size_t)-1;
    bool borrowed_ = false, ok_ = false, done_ = false, in_member_ = false, last_block_ = false;
    const char *error_ = nullptr;
    std::vector<uint8_t> buf_;
    uint8_t *out_ = nullptr, *rd_ = nullptr, *wr_ = nullptr, *crc_from_ = nullptr;
    const uint8_t *in_ = nullptr, *end_ = nullptr;
    uint64_t bitbuf_ = 0;
    unsigned bitcnt_ = 0, overrun_ = 0;
    uint32_t crc_ = 0, stored_left_ = 0;
    uint64_t member_out_ = 0;
    State state_ = BLOCK_HEADER;
    Entry lit_[LIT_TABLE], dist_[DIST_TABLE];
    uint32_t multi_[1 << LIT_BITS];
};

// ---------------------------------------------------------------------------------------------
// BGZF (bgzip, samtools fastq, htslib): every member is a <= 64 KB block whose header carries its own
// compressed size ("BC" extra subfield), so the member boundaries are known without decoding.  The
// file is cut into groups of consecutive blocks (~8 MB compressed, each group is itself a valid
// multi-member gzip stream); worker threads decode groups independently, read() serves them in file
// order.  Data after the last BGZF block (a plain gzip member appended to the file) is decoded by a
// streaming GzReader once the groups are consumed.
// ---------------------------------------------------------------------------------------------
class BgzfParallelReader {
public:
    // BSIZE + 1 of the block at p, or 0 when p is not a BGZF block header
    static size_t block_size(const uint8_t *p, size_t n) {
        if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
        const size_t xlen = p[10] | (p[11] << 8);
        if (n < 12 + xlen) return 0;
        for (size_t q = 12; q + 4 <= 12 + xlen;) {
            const size_t slen = p[q + 2] | (p[q + 3] << 8);
            if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) return (size_t)(p[q + 4] | (p[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        return 0;
    }
    static bool is_bgzf(const std::string &path) {
        uint8_t h[64];
        FILE *f = fopen(path.c_str(), "rb");
        const size_t n = f ? fread(h, 1, sizeof(h), f) : 0;
        if (f) fclose(f);
        return block_size(h, n) != 0;
    }

    BgzfParallelReader(const std::string &path, int threads) {
        fd_ = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) return;
        size_ = (size_t)st.st_size;
        if (size_) {
            void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) return;
            base_ = (const uint8_t *)m;
        }
        ok_ = true;
        window_ = (size_t)std::max(4, 2 * threads);
        for (int t = 0; t < std::max(1, threads); ++t) workers_.emplace_back([this] { work(); });
    }
    ~BgzfParallelReader() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
        if (base_) munmap((void *)base_, size_);
        if (fd_ >= 0) close(fd_);
    }
    bool ok() const { return ok_; }
    bool failed() const { return error_ != nullptr; }
    const char *error() const { return error_; }

    size_t read(void *dst, size_t n) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < n && !error_) {
            if (tail_) { // streaming remainder
                const size_t k = tail_->read(d + got, n - got);
                if (k == 0) { if (tail_->failed()) error_ = tail_->error(); break; }
                got += k;
                continue;
            }
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return !tasks_.empty() ? tasks_.front()->state == 2 : scan_done_; });
            if (tasks_.empty()) { // all groups consumed
                lk.unlock();
                if (scan_off_ < size_) { tail_.reset(new GzReader(base_ + scan_off_, size_ - scan_off_)); continue; }
                break;
            }
            Task &t = *tasks_.front();
            if (t.err) { error_ = t.err; break; }
            lk.unlock();
            const size_t k = std::min(n - got, t.out.size() - t.pos);
            memcpy(d + got, t.out.data() + t.pos, k);
            t.pos += k;
            got += k;
            if (t.pos == t.out.size()) {
                lk.lock();
                tasks_.pop_front();
                lk.unlock();
                cv_.notify_all();
            }
        }
        return got;
    }

private:
    struct Task {
        size_t a = 0, b = 0, pos = 0;
        std::vector<uint8_t> out;
        int state = 0; // 1 decoding, 2 ready
        const char *err = nullptr;
    };
    // next group of blocks from scan_off_ (m_ held); nullptr when the BGZF part is exhausted
    Task *next_task() {
        if (scan_done_) return nullptr;
        size_t a = scan_off_, b = a, out = 0;
        while (b < size_ && b - a < (8u << 20)) {
            const size_t bs = block_size(base_ + b, size_ - b);
            if (bs < 26 || b + bs > size_) break; // not a (complete) BGZF block: the remainder is streamed
            uint32_t isize;
            memcpy(&isize, base_ + b + bs - 4, 4);
            if (isize > 65536u) break; // the BGZF limit per block (SAM spec 4.1): a larger claim is corrupt, the remainder is
                                       // streamed by the checked sequential reader instead of sizing a buffer from it
            out += isize;
            b += bs;
        }
        if (b == a) { scan_done_ = true; return nullptr; }
        scan_off_ = b;
        if (b >= size_) scan_done_ = true;
        tasks_.emplace_back(new Task());
        Task *t = tasks_.back().get();
        t->a = a;
        t->b = b;
        t->out.resize(out);
        t->state = 1;
        return t;
    }
    void work() {
        GzReader rd(nullptr, 0);
        while (true) {
            Task *t;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || scan_done_ || tasks_.size() < window_; });
                if (stop_) return;
                t = next_task();
                if (!t) { cv_.notify_all(); return; }
            }
            rd.reset(base_ + t->a, t->b - t->a);
            size_t got = 0;
            while (got < t->out.size()) {
                const size_t k = rd.read(t->out.data() + got, t->out.size() - got);
                if (k == 0) break;
                got += k;
            }
            uint8_t extra;
            const char *err = nullptr;
            if (rd.failed()) err = rd.error();
            else if (got != t->out.size() || rd.read(&extra, 1) != 0) err = "BGZF block length mismatch";
            else if (rd.failed()) err = rd.error(); // CRC of the last block is checked when its end is reached
            {
                std::lock_guard<std::mutex> lk(m_);
                t->err = err;
                t->state = 2;
            }
            cv_.notify_all();
        }
    }

    int fd_ = -1;
    const uint8_t *base_ = nullptr;
    size_t size_ = 0, scan_off_ = 0, window_ = 8;
    bool ok_ = false, stop_ = false, scan_done_ = false;
    const char *error_ = nullptr;
    std::deque<std::unique_ptr<Task>> tasks_;
    std::vector<std::thread> workers_;
    std::unique_ptr<GzReader> tail_;
    std::mutex m_;
    std::condition_variable cv_;
};

// What MultiMemberReader falls back to: by default the streaming GzReader; pinflate.hpp registers a factory that
// decodes a large remainder with its parallel single-stream reader instead.
struct TailReader {
    virtual ~TailReader() {}
    virtual size_t read(void *dst, size_t n) = 0;
    virtual bool failed() const = 0;
    virtual const char *error() const = 0;
};
struct GzTail : TailReader {
    GzReader r;
    GzTail(const uint8_t *p, size_t n) : r(p, n) {}
    size_t read(void *dst, size_t n) override { return r.read(dst, n); }
    bool failed() const override { return r.failed(); }
    const char *error() const override { return r.error(); }
};
typedef TailReader *(*TailFactory)(const uint8_t *data, size_t n, int threads);
inline TailFactory &tail_factory() {
    static TailFactory f = nullptr;
    return f;
}

// ---------------------------------------------------------------------------------------------
// Concatenated gzip members without block-size headers (`cat *.fastq.gz` of a sequencing run, this
// tool's own per-record members): member starts are *guessed* by scanning for a plausible gzip header,
// spans of ~8 MB between guesses are decoded speculatively in parallel, and the consumer accepts a
// span only if it starts exactly where the previous accepted span ended (span 0 starts at offset 0).
// A decoder started at a true member start stops only at member boundaries, so an accepted chain is
// exactly the sequential decode.  Any anomaly — a guess inside compressed data, a span that fails or
// grows beyond the memory cap — switches to the streaming decoder from the last verified boundary,
// which reproduces the sequential behaviour (including its error, after the good prefix).
// ---------------------------------------------------------------------------------------------
class MultiMemberReader {
public:
    enum : size_t { SPAN = 8u << 20, SPAN_OUT_CAP = 1u << 30 };

    // offset of the first plausible member header at or after `from`, or n
    static size_t next_candidate(const uint8_t *p, size_t n, size_t from) {
        while (from + 18 <= n) {
            const uint8_t *q = (const uint8_t *)memchr(p + from, 0x1f, n - from - 17);
            if (!q) return n;
            const size_t o = (size_t)(q - p);
            if (q[1] == 0x8b && q[2] == 8 && (q[3] & 0xE0) == 0 && (q[8] == 0 || q[8] == 2 || q[8] == 4) && (q[9] <= 13 || q[9] == 255))
                return o;
            from = o + 1;
        }
        return n;
    }
    // worth going parallel: a second member header within the first 64 MB
    static bool is_multi_member(const std::string &path) {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        std::vector<uint8_t> head(1u << 20);
        bool found = false;
        size_t off = 0;
        uint8_t carry[17];
        size_t ncarry = 0;
        while (!found && off < (64u << 20)) {
            memcpy(head.data(), carry, ncarry);
            const ssize_t got = ::read(fd, head.data() + ncarry, head.size() - ncarry);
            if (got <= 0) break;
            const size_t n = ncarry + (size_t)got;
            const size_t c = next_candidate(head.data(), n, off == 0 ? 1 : 0);
            if (c < n) found = true;
            ncarry = std::min<size_t>(17, n);
            memcpy(carry, head.data() + n - ncarry, ncarry);
            off += (size_t)got;
        }
        close(fd);
        return found;
    }

    MultiMemberReader(const std::string &path, int threads) {
        fd_ = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) return;
        size_ = (size_t)st.st_size;
        if (size_) {
            void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) return;
            base_ = (const uint8_t *)m;
        }
        ok_ = true;
        window_ = (size_t)std::max(4, 2 * threads);
        threads_ = std::max(1, threads);
        for (int t = 0; t < std::max(1, threads); ++t) workers_.emplace_back([this] { work(); });
    }
    ~MultiMemberReader() {
        shutdown();
        tail_.reset(); // may run threads over the mapping
        if (base_) munmap((void *)base_, size_);
        if (fd_ >= 0) close(fd_);
    }
    bool ok() const { return ok_; }
    bool failed() const { return error_ != nullptr; }
    const char *error() const { return error_; }
    bool fell_back() const { return tail_ != nullptr; } // tests

    size_t read(void *dst, size_t n) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < n && !error_) {
            if (tail_) {
                const size_t k = tail_->read(d + got, n - got);
                if (k == 0) { if (tail_->failed()) error_ = tail_->error(); break; }
                got += k;
                continue;
            }
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return !tasks_.empty() ? tasks_.front()->state == 2 : scan_done_; });
            if (tasks_.empty()) {
                lk.unlock();
                if (expected_ < size_) fallback(); // bytes after the last span (cannot happen for well-formed input)
                else break;
                continue;
            }
            Task &t = *tasks_.front();
            lk.unlock();
            if (t.pos == 0 && (t.start != expected_ || t.err || t.too_big)) { fallback(); continue; }
            const size_t k = std::min(n - got, t.out.size() - t.pos);
            memcpy(d + got, t.out.data() + t.pos, k);
            t.pos += k;
            got += k;
            if (t.pos == t.out.size()) {
                expected_ = t.end;
                lk.lock();
                tasks_.pop_front();
                lk.unlock();
                cv_.notify_all();
            }
        }
        return got;
    }

private:
    struct Task {
        size_t start = 0, stop = 0, end = 0, pos = 0;
        std::vector<uint8_t> out;
        int state = 0;
        const char *err = nullptr;
        bool too_big = false;
    };
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
        workers_.clear();
    }
    void fallback() { // sequential decode from the last verified member boundary
        shutdown();
        tasks_.clear();
        TailReader *t = tail_factory() ? tail_factory()(base_ + expected_, size_ - expected_, threads_) : nullptr;
        tail_.reset(t ? t : new GzTail(base_ + expected_, size_ - expected_));
    }
    Task *next_task() { // m_ held
        if (scan_done_) return nullptr;
        const size_t start = scan_off_;
        if (start >= size_) { scan_done_ = true; return nullptr; }
        const size_t stop = start + SPAN >= size_ ? size_ : next_candidate(base_, size_, start + SPAN);
        scan_off_ = stop;
        if (stop >= size_) scan_done_ = true;
        tasks_.emplace_back(new Task());
        Task *t = tasks_.back().get();
        t->start = start;
        t->stop = stop;
        t->state = 1;
        return t;
    }
    void work() {
        GzReader rd(nullptr, 0);
        while (true) {
            Task *t;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || scan_done_ || tasks_.size() < window_; });
                if (stop_) return;
                t = next_task();
                if (!t) { cv_.notify_all(); return; }
            }
            if (tail_factory() && t->stop - t->start > 3 * (size_t)SPAN) {
                // no member start for 24 MB: a member this large goes to the parallel single-stream decoder undecoded
                std::lock_guard<std::mutex> lk(m_);
                t->too_big = true;
                t->end = t->start;
                t->state = 2;
                cv_.notify_all();
                continue;
            }
            rd.reset(base_ + t->start, size_ - t->start);
            rd.set_stop(t->stop - t->start);
            t->out.resize(std::min<size_t>(4 * (t->stop - t->start) + (1u << 20), 64u << 20));
            size_t got = 0;
            while (true) {
                if (got == t->out.size()) {
                    // with the parallel single-stream decoder as fallback a member beyond 64 MB is better decoded there
                    if (got >= (tail_factory() ? (size_t)(64u << 20) : (size_t)SPAN_OUT_CAP)) { t->too_big = true; break; }
                    t->out.resize(got * 2);
                }
                const size_t k = rd.read(t->out.data() + got, t->out.size() - got);
                if (k == 0) break;
                got += k;
                if (stop_flag()) break;
            }
            t->out.resize(got);
            const char *err = rd.failed() ? rd.error() : nullptr;
            {
                std::lock_guard<std::mutex> lk(m_);
                t->err = err;
                t->end = t->start + rd.consumed();
                t->state = 2;
            }
            cv_.notify_all();
        }
    }
    bool stop_flag() {
        std::lock_guard<std::mutex> lk(m_);
        return stop_;
    }

    int fd_ = -1;
    const uint8_t *base_ = nullptr;
    size_t size_ = 0, scan_off_ = 0, expected_ = 0, window_ = 8;
    bool ok_ = false, stop_ = false, scan_done_ = false;
    const char *error_ = nullptr;
    std::deque<std::unique_ptr<Task>> tasks_;
    std::vector<std::thread> workers_;
    std::unique_ptr<TailReader> tail_;
    int threads_ = 1;
    std::mutex m_;
    std::condition_variable cv_;
};

}  // namespace fastgz
