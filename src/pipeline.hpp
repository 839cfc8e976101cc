// Host ingest pipeline of the tgsfilter host: a reader thread parses
// FASTQ/FASTA (plain or gzip) straight out of a large read buffer into variable-length
// batches, the main thread feeds the GPUs, a writer thread formats and writes the records.  Replaces
// the reference's 1 reader + N workers + 1 writer joined by 1 ms-sleep polling (T.cpp:1808-1916).
//
// gzip input is decoded by the host's own inflate (inflate.hpp).
//
// The parser follows FastxReader's record rules (T.cpp:685-760): 4-line FASTQ / 2-line FASTA, a
// record starts at the next line beginning with '@' / '>', '\r' before '\n' is dropped, an empty or
// length-mismatched quality line stops the input with the reference's message.
#pragma once
#ifdef TGSF_HAVE_HTSLIB
// Optional: CRAM input through an htslib the build was pointed at (src/Makefile: HTS_DIR).  BAM and SAM never go
// through it (own parser below); CRAM needs htslib's codecs and container logic, as in the reference (T.cpp:984-1040).
#include "sam.h"
#endif
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "inflate.hpp"
#include "pinflate.hpp"

namespace ingest {

struct RawBatch { // pageable; the byte bases stay here for record output
    std::vector<uint8_t> bases, quals;
    std::vector<uint64_t> offsets{0};
    std::vector<std::string> names;
    uint32_t n() const { return (uint32_t)names.size(); }
};

template <typename T>
class Queue { // bounded blocking queue
public:
    explicit Queue(size_t cap) : cap_(cap) {}
    void push(T v) {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&] { return q_.size() < cap_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    T pop() {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&] { return !q_.empty(); });
        T v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return v;
    }

private:
    size_t cap_;
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
};

class FastParser {
public:
    FastParser(const std::string &path, bool fastq, bool sambam = false) : fastq_(fastq), sambam_(sambam) {
        path_ = path;
        // plain files are read with read(2) straight into the parse buffer; gzip goes through zlib
        FILE *probe = fopen(path.c_str(), "rb");
        unsigned char magic[2] = {0, 0};
        const bool gz = probe && fread(magic, 1, 2, probe) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (probe) fclose(probe);
        if (gz && !getenv("TGSF_ZLIB_INFLATE") && inflate_threads() > 1 && fastgz::BgzfParallelReader::is_bgzf(path)) {
            bgzf_.reset(new fastgz::BgzfParallelReader(path, inflate_threads())); // block groups decoded in parallel
            if (!bgzf_->ok()) bgzf_.reset();
        } else if (gz && !getenv("TGSF_ZLIB_INFLATE") && inflate_threads() > 1 && fastgz::MultiMemberReader::is_multi_member(path)) {
            mm_.reset(new fastgz::MultiMemberReader(path, inflate_threads())); // concatenated members, speculative spans
            if (!mm_->ok()) mm_.reset();
        } else if (gz && !getenv("TGSF_ZLIB_INFLATE") && !getenv("TGSF_SERIAL_INFLATE") && inflate_threads() > 1 &&
                   fastgz::SingleStreamReader::worthwhile(path)) {
            ss_.reset(new fastgz::SingleStreamReader(path, inflate_threads())); // one deflate stream, two-pass parallel decode
            if (!ss_->ok()) ss_.reset();
        } else if (gz && !getenv("TGSF_ZLIB_INFLATE")) {
            gz_.reset(new fastgz::GzReader(path)); // own inflate (src/inflate.hpp); zlib kept for A/B runs
            if (!gz_->ok()) gz_.reset();
            else inflater_ = std::thread([this] { // decode ahead of the parser, 4 MB chunks
                while (true) {
                    std::unique_ptr<std::vector<char>> c;
                    {
                        std::lock_guard<std::mutex> lk(spare_m_);
                        if (!spare_.empty()) { c = std::move(spare_.back()); spare_.pop_back(); }
                    }
                    if (!c) c.reset(new std::vector<char>(4u << 20));
                    c->resize(4u << 20);
                    const size_t n = gz_->read(c->data(), c->size());
                    if (n == 0) break;
                    c->resize(n);
                    chunks_.push(std::move(c));
                    if (stop_inflater_.load()) break;
                }
                chunks_.push(nullptr);
            });
        } else if (gz) {
            f_ = gzopen(path.c_str(), "rb");
            if (f_) gzbuffer(f_, 1 << 20);
        } else {
            fd_ = open(path.c_str(), O_RDONLY);
        }
        buf_.resize(8u << 20);
    }
    ~FastParser() {
        if (inflater_.joinable()) {
            stop_inflater_.store(true);
            while (!chunks_done_) { if (!chunks_.pop()) chunks_done_ = true; } // unblock and drain
            inflater_.join();
        }
        if (f_) gzclose(f_);
        if (fd_ >= 0) close(fd_);
#ifdef TGSF_HAVE_HTSLIB
        if (cram_rec_) bam_destroy1(cram_rec_);
        if (cram_hdr_) bam_hdr_destroy(cram_hdr_);
        if (cram_) hts_close(cram_);
#endif
    }
    bool ok() const { return f_ != nullptr || fd_ >= 0 || gz_ != nullptr || bgzf_ != nullptr || mm_ != nullptr || ss_ != nullptr; }
    static int inflate_threads() { // TGSF_INFLATE_THREADS, default min(8, cores - 2)
        if (const char *e = getenv("TGSF_INFLATE_THREADS")) return std::max(1, atoi(e));
        return std::max(1, std::min(8, (int)std::thread::hardware_concurrency() - 2));
    }

    struct Rec { const char *name, *seq, *qual; size_t name_len, seq_len, qual_len; };

    // Decoded bytes of the input, whatever the container; <= 0 at the end (decoder errors are reported on stderr).
    long read_source(char *dst, size_t want) {
        const long got = ss_ ? (long)ss_->read(dst, want)
                       : mm_ ? (long)mm_->read(dst, want)
                       : bgzf_ ? (long)bgzf_->read(dst, want)
                       : gz_ ? (long)read_inflated(dst, want)
                       : f_ ? (long)gzread(f_, dst, (unsigned)want)
                            : (long)read(fd_, dst, want);
        if (ss_ && got == 0 && ss_->failed()) std::cerr << "Error: " << ss_->error() << " (gzip input)" << std::endl;
        if (mm_ && got == 0 && mm_->failed()) std::cerr << "Error: " << mm_->error() << " (gzip input)" << std::endl;
        if (bgzf_ && got == 0 && bgzf_->failed()) std::cerr << "Error: " << bgzf_->error() << " (BGZF input)" << std::endl;
        if (gz_ && got == 0 && gz_->failed()) std::cerr << "Error: " << gz_->error() << " (gzip input)" << std::endl;
        return got;
    }
    bool parallel_decoder() const { return ss_ != nullptr || mm_ != nullptr || bgzf_ != nullptr; }

    // Views stay valid until the next call.  false: end of input (or a malformed record).
    bool next(Rec &r) {
        if (sambam_) return next_sambam(r);
        while (true) {
            size_t p = pos_, lp[4] = {0, 0, 0, 0}, ln[4] = {0, 0, 0, 0};
            int rc = line_at(p, lp[0], ln[0]);
            if (rc == 0) { refill(); continue; }
            if (rc < 0) return false;
            if (ln[0] == 0 || buf_[lp[0]] != (fastq_ ? '@' : '>')) { pos_ = p; continue; } // resynchronise
            const int need = fastq_ ? 4 : 2;
            int got = 1;
            for (; got < need; ++got) {
                rc = line_at(p, lp[got], ln[got]);
                if (rc <= 0) break;
            }
            if (got < need) {
                if (rc == 0) { refill(); continue; } // restart this record with more data
                return false;                         // truncated trailing record
            }
            if (fastq_ && (ln[2] == 0 || buf_[lp[2]] != '+' || ln[1] == 0)) {
                // header without a valid body: the reference goes on from the line after the '+' slot
                size_t q = pos_, a = 0, b = 0;
                for (int i = 0; i < 3; ++i) line_at(q, a, b);
                pos_ = q;
                continue;
            }
            pos_ = p;
            r.name = buf_.data() + lp[0] + 1;
            r.name_len = ln[0] - 1;
            r.seq = buf_.data() + lp[1];
            r.seq_len = ln[1];
            if (fastq_) {
                r.qual = buf_.data() + lp[3];
                r.qual_len = ln[3];
                if (r.qual_len == 0) {
                    std::cerr << "Error: quality are empty:" << std::string(r.name, r.name_len) << std::endl;
                    return false;
                }
                if (r.qual_len != r.seq_len) {
                    std::cerr << "warning: sequence and quality have different length:" << std::string(r.name, r.name_len) << std::endl;
                    return false;
                }
            } else {
                r.qual = nullptr;
                r.qual_len = 0;
                if (r.seq_len == 0) {
                    std::cerr << "Error: sequence are empty:" << std::string(r.name, r.name_len) << std::endl;
                    return false;
                }
            }
            return true;
        }
    }

private:
    // ---- BAM / SAM input (read_bam, T.cpp:1872-1916: every record whatever its flags, name = QNAME, bases through
    // the reference's 16-entry table {=,M,R,S,V,W,Y,H,K,D,B -> NUL; A C G T N}, quality byte + 33).  The
    // reference reads these through htslib, which is not linked here: BAM is BGZF (decoded by the readers
    // above, in parallel) around the little-endian records of the SAM specification section 4.2; the container is
    // recognised from the decoded bytes ("BAM\1", else SAM text), as hts_open does.
    bool ensure(size_t n) { // at least n unread bytes in the buffer
        while (len_ - pos_ < n) {
            if (eof_) return false;
            refill();
        }
        return true;
    }
    uint32_t rd32(size_t at) const { uint32_t v; memcpy(&v, buf_.data() + at, 4); return v; }
    bool next_sambam(Rec &r) {
        static const char kBase[16] = {0, 'A', 'C', 0, 'G', 0, 0, 0, 'T', 0, 0, 0, 0, 0, 0, 'N'}; // T.cpp:31
        if (sb_state_ == 0) {
            if (!ensure(4)) return false;
            if (memcmp(buf_.data() + pos_, "CRAM", 4) == 0) {
#ifdef TGSF_HAVE_HTSLIB
                cram_ = hts_open(path_.c_str(), "r");
                if (!cram_) { std::cerr << "Error: Failed to open file: " << path_ << std::endl; return false; }
                hts_set_opt(cram_, HTS_OPT_NTHREADS, inflate_threads());
                hts_set_log_level(HTS_LOG_OFF);
                cram_hdr_ = sam_hdr_read(cram_);
                cram_rec_ = bam_init1();
                if (!cram_hdr_ || !cram_rec_) { std::cerr << "Error: bad CRAM header: " << path_ << std::endl; return false; }
                sb_state_ = 3;
#else
                std::cerr << "Error: CRAM input needs a build with htslib (make -C src HTS_DIR=<htslib prefix>); "
                             "convert with `samtools fastq` or to BAM" << std::endl;
                return false;
#endif
            } else if (memcmp(buf_.data() + pos_, "BAM\1", 4) == 0) {
                if (!ensure(8)) return false;
                const size_t l_text = rd32(pos_ + 4);
                if (!ensure(12 + l_text)) return false;
                const uint32_t n_ref = rd32(pos_ + 8 + l_text);
                pos_ += 12 + l_text;
                for (uint32_t i = 0; i < n_ref; ++i) {
                    if (!ensure(4)) return false;
                    const size_t l_name = rd32(pos_);
                    if (!ensure(8 + l_name)) return false;
                    pos_ += 8 + l_name;
                }
                sb_state_ = 1;
            } else {
                sb_state_ = 2;
                for (int c = 0; c < 256; ++c) sam_code_[c] = 15; // seq_nt16_table: unknown characters read as N
                const char *codes = "=ACMGRSVTWYHKDBN";
                for (int i = 0; i < 16; ++i) {
                    sam_code_[(unsigned char)codes[i]] = (uint8_t)i;
                    sam_code_[(unsigned char)tolower(codes[i])] = (uint8_t)i;
                }
            }
        }
#ifdef TGSF_HAVE_HTSLIB
        if (sb_state_ == 3) { // CRAM record decoded by htslib, converted exactly like a BAM record (T.cpp:1003-1027)
            if (sam_read1(cram_, cram_hdr_, cram_rec_) < 0) return false;
            const size_t l_seq = (size_t)cram_rec_->core.l_qseq;
            r.name = bam_get_qname(cram_rec_);
            r.name_len = strlen(r.name);
            sb_seq_.resize(l_seq);
            sb_qual_.resize(l_seq);
            const uint8_t *sq = bam_get_seq(cram_rec_), *ql = bam_get_qual(cram_rec_);
            for (size_t i = 0; i < l_seq; ++i) {
                sb_seq_[i] = kBase[bam_seqi(sq, i)];
                sb_qual_[i] = (char)(ql[i] + 33);
            }
            r.seq = sb_seq_.data();
            r.seq_len = l_seq;
            r.qual = sb_qual_.data();
            r.qual_len = l_seq;
            return true;
        }
#endif
        if (sb_state_ == 1) { // BAM record: block_size, 32 fixed bytes, name, cigar, 4-bit bases, qualities, tags
            if (!ensure(4)) return false;
            const size_t bs = rd32(pos_);
            // corrupt or truncated: sam_read1 < 0 ends the reference's loop too (a record of 512 MB is not a read)
            if (bs < 32 || bs > (512u << 20) || !ensure(4 + bs)) return false;
            const uint8_t *p = (const uint8_t *)buf_.data() + pos_ + 4;
            const size_t l_name = p[8], n_cigar = (size_t)p[12] | ((size_t)p[13] << 8), l_seq = rd32(pos_ + 4 + 16);
            const size_t off = 32 + l_name + 4 * n_cigar;
            if ((l_seq >> 31) || off + (l_seq + 1) / 2 + l_seq > bs) return false;
            r.name = (const char *)p + 32;
            r.name_len = strnlen(r.name, l_name);
            sb_seq_.resize(l_seq);
            sb_qual_.resize(l_seq);
            const uint8_t *sq = p + off, *ql = sq + (l_seq + 1) / 2;
            static const struct Pairs { // packed byte -> its two bases
                uint16_t v[256];
                Pairs() {
                    for (int b = 0; b < 256; ++b) {
                        const char two[2] = {kBase[b >> 4], kBase[b & 15]};
                        memcpy(&v[b], two, 2);
                    }
                }
            } pairs;
            char *dst = sb_seq_.data();
            for (size_t i = 0; i < l_seq / 2; ++i) memcpy(dst + 2 * i, &pairs.v[sq[i]], 2);
            if (l_seq & 1) sb_seq_[l_seq - 1] = kBase[sq[l_seq >> 1] >> 4];
            for (size_t i = 0; i < l_seq; ++i) sb_qual_[i] = (char)(ql[i] + 33);
            pos_ += 4 + bs;
        } else { // SAM line: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [tags]
            size_t lp = 0, ln = 0;
            while (true) {
                size_t p = pos_;
                const int rc = line_at(p, lp, ln);
                if (rc == 0) { refill(); continue; }
                if (rc < 0) return false;
                pos_ = p;
                if (ln == 0 || buf_[lp] == '@') continue; // header
                break;
            }
            const char *f[12];
            size_t nf = 0;
            const char *line = buf_.data() + lp, *lend = line + ln;
            f[nf++] = line;
            for (const char *c = line; c < lend && nf < 12; ++c)
                if (*c == '\t') f[nf++] = c + 1;
            if (nf < 11) return false; // sam_parse1 fails: the reference stops reading here
            const char *qend = nf > 11 ? f[11] - 1 : lend;
            r.name = f[0];
            r.name_len = (size_t)(f[1] - 1 - f[0]);
            const char *sq = f[9];
            size_t l_seq = (size_t)(f[10] - 1 - f[9]);
            const char *ql = f[10];
            const size_t l_qual = (size_t)(qend - ql);
            if (l_seq == 1 && sq[0] == '*') l_seq = 0;
            const bool no_qual = l_qual == 1 && ql[0] == '*';
            if (!no_qual && l_qual != l_seq) return false; // "SEQ and QUAL are of different length"
            sb_seq_.resize(l_seq);
            sb_qual_.resize(l_seq);
            for (size_t i = 0; i < l_seq; ++i) sb_seq_[i] = kBase[sam_code_[(unsigned char)sq[i]]];
            for (size_t i = 0; i < l_seq; ++i) sb_qual_[i] = no_qual ? (char)(0xFF + 33) : ql[i]; // (q - 33) + 33
        }
        r.seq = sb_seq_.data();
        r.seq_len = sb_seq_.size();
        r.qual = sb_qual_.data();
        r.qual_len = sb_qual_.size();
        return true;
    }
    int sb_state_ = 0; // 0: container not looked at yet, 1: BAM records, 2: SAM lines, 3: CRAM through htslib
    std::string path_;
#ifdef TGSF_HAVE_HTSLIB
    htsFile *cram_ = nullptr;
    bam_hdr_t *cram_hdr_ = nullptr;
    bam1_t *cram_rec_ = nullptr;
#endif
    std::vector<char> sb_seq_, sb_qual_;
    uint8_t sam_code_[256];

    // Line starting at p: [lp, lp+n) without its terminator; advances p.  1 = line, 0 = incomplete
    // (more input needed), -1 = end of data.  At EOF an unterminated tail counts as a line.
    int line_at(size_t &p, size_t &lp, size_t &n) {
        if (p >= len_) return eof_ ? -1 : 0;
        const char *s = buf_.data() + p;
        const char *nl = (const char *)memchr(s, '\n', len_ - p);
        lp = p;
        if (!nl) {
            if (!eof_) return 0;
            n = len_ - p;
            p = len_;
        } else {
            n = (size_t)(nl - s);
            p += n + 1;
        }
        if (n && buf_[lp + n - 1] == '\r') --n;
        return 1;
    }
    void refill() {
        if (pos_ > 0) {
            memmove(buf_.data(), buf_.data() + pos_, len_ - pos_);
            len_ -= pos_;
            pos_ = 0;
        }
        if (len_ == buf_.size()) buf_.resize(buf_.size() * 2); // one record larger than the buffer
        while (len_ < buf_.size()) {
            const size_t want = std::min<size_t>(buf_.size() - len_, 1u << 30);
            const long got = read_source(buf_.data() + len_, want);
            if (got <= 0) { eof_ = true; break; }
            len_ += (size_t)got;
            if (len_ >= buf_.size() / 2) break;
        }
    }
    size_t read_inflated(char *dst, size_t want) { // bytes of the decoder thread's chunks, in order
        size_t got = 0;
        while (got < want && !chunks_done_) {
            if (!cur_ || cur_pos_ == cur_->size()) {
                if (cur_) {
                    std::lock_guard<std::mutex> lk(spare_m_);
                    if (spare_.size() < 4) spare_.push_back(std::move(cur_));
                }
                cur_ = chunks_.pop();
                cur_pos_ = 0;
                if (!cur_) { chunks_done_ = true; break; }
            }
            const size_t n = std::min(want - got, cur_->size() - cur_pos_);
            memcpy(dst + got, cur_->data() + cur_pos_, n);
            cur_pos_ += n;
            got += n;
            if (got) break; // hand over what is there; the caller asks again
        }
        return got;
    }
    gzFile f_ = nullptr;
    std::unique_ptr<fastgz::GzReader> gz_;
    std::unique_ptr<fastgz::BgzfParallelReader> bgzf_;
    std::unique_ptr<fastgz::MultiMemberReader> mm_;
    std::unique_ptr<fastgz::SingleStreamReader> ss_;
    std::thread inflater_;
    std::atomic<bool> stop_inflater_{false};
    Queue<std::unique_ptr<std::vector<char>>> chunks_{4};
    std::unique_ptr<std::vector<char>> cur_;
    size_t cur_pos_ = 0;
    bool chunks_done_ = false;
    std::mutex spare_m_;
    std::vector<std::unique_ptr<std::vector<char>>> spare_;
    int fd_ = -1;
    bool fastq_, sambam_ = false, eof_ = false;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
};

// Reader thread body: parse the whole file into batches of ~batch_bases bases; nullptr ends the stream.
// Recycled RawBatch buffers (touched pages are expensive to fault in again for every batch).
class BatchPool {
public:
    std::unique_ptr<RawBatch> get() {
        std::lock_guard<std::mutex> lk(m_);
        if (free_.empty()) return std::unique_ptr<RawBatch>(new RawBatch());
        std::unique_ptr<RawBatch> b = std::move(free_.back());
        free_.pop_back();
        return b;
    }
    void put(std::unique_ptr<RawBatch> b) {
        b->bases.clear();
        b->quals.clear();
        b->offsets.assign(1, 0);
        b->names.clear();
        std::lock_guard<std::mutex> lk(m_);
        if (free_.size() < 32) free_.push_back(std::move(b));
    }

private:
    std::mutex m_;
    std::vector<std::unique_ptr<RawBatch>> free_;
};

class FastParser;
inline void stream_reader_main(FastParser &src, bool fastq, uint64_t batch_bases, Queue<std::unique_ptr<RawBatch>> *out,
                               BatchPool *pool, const std::atomic<bool> *stop, int threads);

inline void reader_main(const std::string &path, bool fastq, uint64_t batch_bases,
                        Queue<std::unique_ptr<RawBatch>> *out, BatchPool *pool, const std::atomic<bool> *stop = nullptr,
                        bool sambam = false) {
    FastParser ps(path, fastq, sambam);
    // opt-in (TGSF_STREAM_PARSE_THREADS=n): compressed FASTQ/FASTA decoded by a parallel reader is also parsed in parallel
    if (const char *e = getenv("TGSF_STREAM_PARSE_THREADS"))
        if (atoi(e) > 0 && !sambam && ps.ok() && ps.parallel_decoder()) {
            stream_reader_main(ps, fastq, batch_bases, out, pool, stop, atoi(e));
            return;
        }
    auto fresh = [&]() {
        std::unique_ptr<RawBatch> nb = pool->get();
        nb->bases.reserve((size_t)batch_bases + (batch_bases >> 2));
        if (fastq) nb->quals.reserve((size_t)batch_bases + (batch_bases >> 2));
        return nb;
    };
    std::unique_ptr<RawBatch> b = fresh();
    FastParser::Rec r;
    while (ps.ok() && ps.next(r)) {
        b->names.emplace_back(r.name, r.name_len);
        b->bases.insert(b->bases.end(), (const uint8_t *)r.seq, (const uint8_t *)r.seq + r.seq_len);
        if (fastq) b->quals.insert(b->quals.end(), (const uint8_t *)r.qual, (const uint8_t *)r.qual + r.qual_len);
        b->offsets.push_back(b->bases.size());
        if (b->bases.size() >= batch_bases) {
            out->push(std::move(b));
            b = fresh();
            if (stop && stop->load()) break; // the consumer restarts the input (two-pass mode)
        }
    }
    if (b->n() && !(stop && stop->load())) out->push(std::move(b));
    out->push(nullptr);
}

// ---------------------------------------------------------------------------------------------
// Parallel ingest of plain (uncompressed) files: the file is mmap'ed and cut into ~64 MB chunks at
// record boundaries; worker threads parse chunks independently, the consumer receives the batches
// in file order.  A FASTQ record start is a line beginning with '@' whose second-next line begins
// with '+' (exact for well-formed 4-line FASTQ: a quality line that happens to start with '@' is
// followed by a header and a base line, never by a '+' line two below); FASTA: a line beginning
// with '>'.  gzip input keeps the single-reader path (inflate is serial).
// ---------------------------------------------------------------------------------------------
class ParallelReader {
public:
    ParallelReader(const std::string &path, bool fastq, uint64_t chunk_bytes, int threads, BatchPool *pool)
        : fastq_(fastq), chunk_(chunk_bytes), pool_(pool) {
        fd_ = open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) return;
        size_ = (size_t)st.st_size;
        if (size_ == 0) { ok_ = true; n_chunks_ = 0; return; }
        void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) return;
        base_ = (const char *)m;
        madvise((void *)base_, size_, MADV_SEQUENTIAL);
        n_chunks_ = (size_ + chunk_ - 1) / chunk_;
        slots_.resize(n_chunks_);
        ready_.assign(n_chunks_, 0);
        window_ = (size_t)std::max(4, 3 * threads);
        ok_ = true;
        for (int t = 0; t < threads; ++t) workers_.emplace_back([this] { work(); });
    }
    ~ParallelReader() {
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
    static bool is_gzip(const std::string &path) {
        FILE *f = fopen(path.c_str(), "rb");
        unsigned char m[2] = {0, 0};
        const bool gz = f && fread(m, 1, 2, f) == 2 && m[0] == 0x1f && m[1] == 0x8b;
        if (f) fclose(f);
        return gz;
    }
    // next batch in file order; nullptr at the end (or after a malformed record)
    std::unique_ptr<RawBatch> pop() {
        std::unique_lock<std::mutex> lk(m_);
        while (true) {
            if (failed_ || next_out_ >= n_chunks_) return nullptr;
            cv_.wait(lk, [&] { return ready_[next_out_] != 0; });
            std::unique_ptr<RawBatch> b = std::move(slots_[next_out_]);
            const bool bad = ready_[next_out_] == 2;
            ++next_out_;
            cv_.notify_all();
            if (bad) failed_ = true; // the records before the malformed one are still delivered
            if (b && b->n()) return b;
            if (bad) return nullptr;
        }
    }

private:
    // first record start at or after `from` (from == 0: the file start)
    size_t record_start(size_t from) const {
        if (from == 0) return 0;
        if (from >= size_) return size_;
        const char *p = (const char *)memchr(base_ + from - 1, '\n', size_ - from + 1);
        if (!p) return size_;
        size_t pos = (size_t)(p - base_) + 1;
        const char hdr = fastq_ ? '@' : '>';
        while (pos < size_) {
            if (base_[pos] == hdr) {
                if (!fastq_) return pos;
                // line + 2 must start with '+'
                const char *l1 = (const char *)memchr(base_ + pos, '\n', size_ - pos);
                if (!l1) return size_;
                const char *l2 = (const char *)memchr(l1 + 1, '\n', size_ - (size_t)(l1 + 1 - base_));
                if (!l2) return size_;
                if ((size_t)(l2 + 1 - base_) < size_ && l2[1] == '+') return pos;
            }
            const char *nl = (const char *)memchr(base_ + pos, '\n', size_ - pos);
            if (!nl) return size_;
            pos = (size_t)(nl - base_) + 1;
        }
        return size_;
    }
    static inline bool get_line(const char *&p, const char *e, const char *&ls, size_t &n) {
        if (p >= e) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        ls = p;
        if (nl) { n = (size_t)(nl - p); p = nl + 1; }
        else { n = (size_t)(e - p); p = e; }
        if (n && ls[n - 1] == '\r') --n;
        return true;
    }
    // parse [s, e) (record aligned); false on a malformed record (same messages as the serial parser)
    bool parse_range(const char *s, const char *e, RawBatch &b) const { return parse_range(fastq_, s, e, b); }
public:
    // parse [s, e) (record aligned) into b; false on a malformed record (same messages as the serial parser)
    static bool parse_range(const bool fastq_, const char *s, const char *e, RawBatch &b) {
        const char *p = s;
        const char hdr = fastq_ ? '@' : '>';
        while (p < e) {
            const char *l0, *l1, *l2, *l3;
            size_t n0, n1, n2 = 0, n3 = 0;
            const char *save = p;
            if (!get_line(p, e, l0, n0)) break;
            if (n0 == 0 || l0[0] != hdr) continue; // resynchronise
            if (!get_line(p, e, l1, n1)) break;
            if (fastq_) {
                if (!get_line(p, e, l2, n2) || !get_line(p, e, l3, n3)) break;
                if (n2 == 0 || l2[0] != '+' || n1 == 0) { // header without a valid body
                    p = save;
                    for (int i = 0; i < 3; ++i) get_line(p, e, l0, n0);
                    continue;
                }
                if (n3 == 0) { std::cerr << "Error: quality are empty:" << std::string(l0 + 1, n0 - 1) << std::endl; return false; }
                if (n3 != n1) { std::cerr << "warning: sequence and quality have different length:" << std::string(l0 + 1, n0 - 1) << std::endl; return false; }
                b.quals.insert(b.quals.end(), (const uint8_t *)l3, (const uint8_t *)l3 + n3);
            } else if (n1 == 0) {
                std::cerr << "Error: sequence are empty:" << std::string(l0 + 1, n0 - 1) << std::endl;
                return false;
            }
            b.names.emplace_back(l0 + 1, n0 - 1);
            b.bases.insert(b.bases.end(), (const uint8_t *)l1, (const uint8_t *)l1 + n1);
            b.offsets.push_back(b.bases.size());
        }
        return true;
    }
private:
    void work() {
        while (true) {
            size_t ci;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || failed_ || next_in_ >= n_chunks_ || next_in_ < next_out_ + window_; });
                if (stop_ || failed_ || next_in_ >= n_chunks_) return;
                ci = next_in_++;
            }
            const size_t s = record_start(ci * chunk_), e = record_start((ci + 1) * chunk_);
            std::unique_ptr<RawBatch> b = pool_->get();
            bool good = true;
            if (e > s) {
                b->bases.reserve((e - s) / 2 + 1024);
                if (fastq_) b->quals.reserve((e - s) / 2 + 1024);
                good = parse_range(base_ + s, base_ + e, *b);
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                slots_[ci] = std::move(b);
                ready_[ci] = good ? 1 : 2;
            }
            cv_.notify_all();
        }
    }

    bool fastq_, ok_ = false, stop_ = false, failed_ = false;
    uint64_t chunk_;
    BatchPool *pool_;
    int fd_ = -1;
    const char *base_ = nullptr;
    size_t size_ = 0, n_chunks_ = 0, next_in_ = 0, next_out_ = 0, window_ = 8;
    std::vector<std::unique_ptr<RawBatch>> slots_;
    std::vector<char> ready_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
};

// ---------------------------------------------------------------------------------------------
// Parallel parsing of a decoded stream (gzip / BGZF input): behind the parallel decoders one parser thread is
// the next limit (~1.6 GB/s of FASTQ on the bench host).  The reader thread copies the decoded bytes into
// chunks, cuts each chunk at its last verified record start (a header line whose second-next line starts
// with '+'; FASTA: a '>' line) and carries the tail over; worker threads parse the chunks with the plain-file
// chunk parser; batches leave in stream order.  Same contract as reader_main (nullptr ends the stream; the
// records in front of a malformed one are still delivered).
// ---------------------------------------------------------------------------------------------
inline size_t last_record_start(const char *b, size_t len, bool fastq) { // 0: none (or only the buffer start)
    const char hdr = fastq ? '@' : '>';
    size_t e = len;
    while (true) {
        const void *q = e ? memrchr(b, '\n', e) : nullptr; // last newline in [0, e)
        const size_t ls = q ? (size_t)((const char *)q - b) + 1 : 0;
        if (ls > 0 && ls < len && b[ls] == hdr) {
            if (!fastq) return ls;
            const char *l1 = (const char *)memchr(b + ls, '\n', len - ls);
            const char *l2 = l1 ? (const char *)memchr(l1 + 1, '\n', len - (size_t)(l1 + 1 - b)) : nullptr;
            if (l2 && (size_t)(l2 + 1 - b) < len && l2[1] == '+') return ls;
        }
        if (!q) return 0;
        e = ls - 1;
    }
}

inline void stream_reader_main(FastParser &src, bool fastq, uint64_t batch_bases, Queue<std::unique_ptr<RawBatch>> *out,
                               BatchPool *pool, const std::atomic<bool> *stop, int threads) {
    struct Job {
        std::vector<char> buf;
        size_t len = 0;
        std::unique_ptr<RawBatch> batch;
        int state = 0; // 1 parsed, 2 parsed up to a malformed record
    };
    std::mutex m;
    std::condition_variable cv;
    std::deque<Job *> todo;
    bool quit = false;
    std::vector<std::thread> workers;
    for (int t = 0; t < std::max(1, threads); ++t)
        workers.emplace_back([&] {
            while (true) {
                Job *j;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return quit || !todo.empty(); });
                    if (todo.empty()) return;
                    j = todo.front();
                    todo.pop_front();
                }
                j->batch = pool->get();
                j->batch->bases.reserve(j->len / (fastq ? 2 : 1) + 1024);
                if (fastq) j->batch->quals.reserve(j->len / 2 + 1024);
                const bool good = ParallelReader::parse_range(fastq, j->buf.data(), j->buf.data() + j->len, *j->batch);
                {
                    std::lock_guard<std::mutex> lk(m);
                    j->state = good ? 1 : 2;
                }
                cv.notify_all();
            }
        });
    const size_t chunk = std::max<size_t>(1u << 20, (size_t)(fastq ? 2 * batch_bases : batch_bases));
    const size_t window = (size_t)std::max(1, threads) + 2;
    std::deque<std::unique_ptr<Job>> inflight;
    std::vector<std::unique_ptr<Job>> spare;
    std::vector<char> carry;
    bool eof = false, ended = false; // ended: malformed record or stop request — nothing more is delivered
    auto deliver_front = [&]() {     // wait for the oldest chunk, hand its batch on
        Job *j = inflight.front().get();
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return j->state != 0; });
        }
        if (!ended) {
            if (j->batch && j->batch->n()) out->push(std::move(j->batch));
            if (j->state == 2 || (stop && stop->load())) ended = true;
        }
        if (j->batch) pool->put(std::move(j->batch));
        j->state = 0;
        spare.push_back(std::move(inflight.front()));
        inflight.pop_front();
    };
    while (!eof && !ended) {
        std::unique_ptr<Job> j;
        if (!spare.empty()) { j = std::move(spare.back()); spare.pop_back(); }
        else j.reset(new Job());
        if (j->buf.size() < chunk + carry.size()) j->buf.resize(chunk + carry.size());
        memcpy(j->buf.data(), carry.data(), carry.size());
        size_t len = carry.size(), cut = 0;
        while (true) {
            while (len < j->buf.size()) {
                const long got = src.read_source(j->buf.data() + len, std::min<size_t>(j->buf.size() - len, 1u << 30));
                if (got <= 0) { eof = true; break; }
                len += (size_t)got;
            }
            cut = eof ? len : last_record_start(j->buf.data(), len, fastq);
            if (eof || cut > 0) break;
            j->buf.resize(j->buf.size() * 2); // one record longer than the chunk
        }
        carry.assign(j->buf.data() + cut, j->buf.data() + len);
        j->len = cut;
        Job *raw = j.get();
        inflight.push_back(std::move(j));
        {
            std::lock_guard<std::mutex> lk(m);
            todo.push_back(raw);
        }
        cv.notify_all();
        while (!inflight.empty() && (inflight.size() > window || [&] { std::lock_guard<std::mutex> lk(m); return inflight.front()->state != 0; }()))
            deliver_front();
    }
    while (!inflight.empty()) deliver_front();
    {
        std::lock_guard<std::mutex> lk(m);
        quit = true;
    }
    cv.notify_all();
    for (auto &w : workers) w.join();
    out->push(nullptr);
}

}  // namespace ingest
