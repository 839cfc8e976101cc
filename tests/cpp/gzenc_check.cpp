// CPU check of tgsfilter_b200/csrc/gzenc_core.h: members assembled from the serial building blocks must
// inflate (zlib) to the record.  gzenc_check <fastq> [fasta 0|1] -> writes <fastq>.members.gz, prints sizes.
#include "../../tgsfilter_b200/csrc/gzenc_core.h"
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>

static void encode_block(GzBitWriter &w, const uint8_t *const *parts, const size_t *lens, int nparts, bool final_block) {
    uint32_t freq[GZ_NSYM] = {0};
    for (int p = 0; p < nparts; ++p) for (size_t i = 0; i < lens[p]; ++i) freq[parts[p][i]]++;
    freq[256] = 1;
    uint8_t len[GZ_NSYM];
    uint16_t code[GZ_NSYM];
    std::vector<uint32_t> scratch(5 * GZ_NSYM + 400);
    gz_huff_lengths(freq, GZ_NSYM, GZ_MAX_BITS, len, scratch.data());
    gz_huff_codes(len, GZ_NSYM, code);
    gz_write_dyn_header(w, len, final_block, scratch.data());
    for (int p = 0; p < nparts; ++p) for (size_t i = 0; i < lens[p]; ++i) gz_bw_put(w, code[parts[p][i]], len[parts[p][i]]);
    gz_bw_put(w, code[256], len[256]);
}

static std::string member(const std::string &name, const std::string &seq, const std::string &qual, bool fastq) {
    const std::string head = (fastq ? "@" : ">") + name + "\n";
    std::string rec = head + seq + "\n";
    if (fastq) rec += "+\n" + qual + "\n";
    std::vector<uint8_t> out(rec.size() * 2 + 1024, 0);
    size_t pos = 0;
    const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
    memcpy(&out[pos], hdr, 10); pos += 10;
    out[pos++] = 0; // stored block, not final
    out[pos++] = (uint8_t)head.size(); out[pos++] = (uint8_t)(head.size() >> 8);
    out[pos++] = (uint8_t)~head.size(); out[pos++] = (uint8_t)(~head.size() >> 8);
    memcpy(&out[pos], head.data(), head.size()); pos += head.size();
    GzBitWriter w;
    gz_bw_init(w, &out[pos]);
    const uint8_t *nl = (const uint8_t *)"\n", *plus = (const uint8_t *)"\n+\n";
    if (fastq) {
        const uint8_t *p1[1] = {(const uint8_t *)seq.data()};
        const size_t l1[1] = {seq.size()};
        encode_block(w, p1, l1, 1, false);
        const uint8_t *p2[3] = {plus, (const uint8_t *)qual.data(), nl};
        const size_t l2[3] = {3, qual.size(), 1};
        encode_block(w, p2, l2, 3, true);
    } else {
        const uint8_t *p1[2] = {(const uint8_t *)seq.data(), nl};
        const size_t l1[2] = {seq.size(), 1};
        encode_block(w, p1, l1, 2, true);
    }
    gz_bw_flush(w);
    pos += w.pos;
    const uint32_t crc = (uint32_t)crc32(0, (const Bytef *)rec.data(), (uInt)rec.size()), isz = (uint32_t)rec.size();
    memcpy(&out[pos], &crc, 4); memcpy(&out[pos + 4], &isz, 4); pos += 8;
    return std::string((const char *)out.data(), pos);
}

// code lengths for Fibonacci-like frequencies would exceed the limit: the builder must stay within it and
// still return a complete prefix code (Kraft sum exactly 1)
static int selftest() {
    for (int limit : {15, 7}) {
        for (int n : {2, 3, 19, 30, 60, 257}) {
            std::vector<uint32_t> freq(n), scratch(5 * n + 8);
            uint64_t a = 1, b = 1;
            for (int i = 0; i < n; ++i) { freq[i] = (uint32_t)std::min<uint64_t>(a, 0x7fffffff); const uint64_t c = a + b; a = b; b = c; }
            std::vector<uint8_t> len(n);
            if ((1 << limit) < n) continue; // not representable at all
            gz_huff_lengths(freq.data(), n, limit, len.data(), scratch.data());
            double kraft = 0;
            int mx = 0;
            for (int i = 0; i < n; ++i) { if (!len[i]) return 1; kraft += 1.0 / (double)(1u << len[i]); mx = std::max<int>(mx, len[i]); }
            if (mx > limit || kraft != 1.0) { fprintf(stderr, "selftest: n=%d limit=%d max=%d kraft=%f\n", n, limit, mx, kraft); return 1; }
        }
    }
    printf("selftest ok\n");
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (std::string(argv[1]) == "--selftest") return selftest();
    const bool fasta = argc > 2 && atoi(argv[2]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    std::string all;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) all.append(buf, n);
    fclose(f);
    std::string out, expect;
    size_t p = 0, nrec = 0;
    while (p < all.size()) {
        std::string l[4];
        for (int k = 0; k < 4 && p < all.size(); ++k) {
            const size_t e = all.find('\n', p);
            l[k] = all.substr(p, e - p);
            p = e + 1;
        }
        out += member(l[0].substr(1), l[1], l[3], !fasta);
        expect += (fasta ? ">" : "@") + l[0].substr(1) + "\n" + l[1] + "\n" + (fasta ? std::string() : "+\n" + l[3] + "\n");
        ++nrec;
    }
    const std::string path = std::string(argv[1]) + ".members.gz";
    f = fopen(path.c_str(), "wb");
    fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    // decode with zlib
    gzFile g = gzopen(path.c_str(), "rb");
    std::string got;
    int k;
    while ((k = gzread(g, buf, sizeof(buf))) > 0) got.append(buf, (size_t)k);
    gzclose(g);
    printf("%zu records, %zu -> %zu bytes (%.3f), roundtrip %s\n", nrec, expect.size(), out.size(), (double)out.size() / expect.size(),
           got == expect ? "ok" : "MISMATCH");
    return got == expect ? 0 : 1;
}
