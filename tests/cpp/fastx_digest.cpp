// fastx_digest: order-independent digest of a FASTQ / FASTA file (test infrastructure).
//
//   fastx_digest <file> [--by-length out.tsv] [--names out.tsv]
//
// Prints one JSON line: records, bases, sum / xor of a 64-bit hash of every record (header line,
// sequence, quality).  The reference CLI writes its records in worker-completion order when it runs
// with more than one thread, so full-size comparisons are made on this multiset digest.
//   --by-length : one line per sequence length: "<len> <count> <sum of record hashes>"  (downsampling
//                 picks arbitrary reads among equal lengths at the cut-off, see T.cpp:2297-2301)
//   --names     : one line per record: "<hash> <len> <name>" (to name the records that differ)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

static inline uint64_t mix(uint64_t h) {
    h ^= h >> 32;
    h *= 0xd6e8feb86659fd93ull;
    h ^= h >> 32;
    h *= 0xd6e8feb86659fd93ull;
    h ^= h >> 32;
    return h;
}

static inline uint64_t hash_bytes(const uint8_t *p, size_t n, uint64_t seed) {
    uint64_t h = seed ^ (n * 0x9e3779b97f4a7c15ull);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 29;
        p += 8;
        n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    return mix(h);
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: fastx_digest <file> [--by-length out] [--names out]\n");
        return 2;
    }
    const char *by_len_path = nullptr, *names_path = nullptr;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--by-length")) by_len_path = argv[i + 1];
        else if (!strcmp(argv[i], "--names")) names_path = argv[i + 1];
    }
    int fd = open(argv[1], O_RDONLY);
    if (fd < 0) { perror(argv[1]); return 1; }
    struct stat st;
    fstat(fd, &st);
    const size_t size = (size_t)st.st_size;
    const uint8_t *d = size ? (const uint8_t *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    if (size && d == MAP_FAILED) { perror("mmap"); return 1; }
    if (size) madvise((void *)d, size, MADV_SEQUENTIAL);

    FILE *fn = names_path ? fopen(names_path, "w") : nullptr;
    std::map<uint64_t, std::pair<uint64_t, uint64_t>> by_len;
    uint64_t records = 0, bases = 0, sum = 0, x = 0;
    size_t p = 0;
    auto line = [&](size_t &b, size_t &e) { // [b, e) without the newline; false at end of file
        if (p >= size) return false;
        b = p;
        const void *nl = memchr(d + p, '\n', size - p);
        e = nl ? (size_t)((const uint8_t *)nl - d) : size;
        p = e + 1;
        return true;
    };
    size_t b0, e0, b1, e1, b2, e2, b3, e3;
    while (line(b0, e0)) {
        if (e0 == b0) continue;
        const bool fq = d[b0] == '@';
        if (!fq && d[b0] != '>') { fprintf(stderr, "unexpected header at byte %zu\n", b0); return 1; }
        if (!line(b1, e1)) { fprintf(stderr, "truncated record\n"); return 1; }
        uint64_t h = hash_bytes(d + b0, e0 - b0, 1) + 3 * hash_bytes(d + b1, e1 - b1, 2);
        if (fq) {
            if (!line(b2, e2) || !line(b3, e3)) { fprintf(stderr, "truncated record\n"); return 1; }
            h += 5 * hash_bytes(d + b3, e3 - b3, 3);
        }
        h = mix(h);
        const uint64_t len = e1 - b1;
        ++records;
        bases += len;
        sum += h;
        x ^= h;
        if (by_len_path) {
            auto &s = by_len[len];
            s.first++;
            s.second += h;
        }
        if (fn) {
            size_t ne = b0 + 1;
            while (ne < e0 && d[ne] != ' ' && d[ne] != '\t') ++ne;
            fprintf(fn, "%016" PRIx64 " %" PRIu64 " %.*s\n", h, len, (int)(ne - b0 - 1), (const char *)d + b0 + 1);
        }
    }
    if (fn) fclose(fn);
    if (by_len_path) {
        FILE *f = fopen(by_len_path, "w");
        if (!f) { perror(by_len_path); return 1; }
        for (auto &kv : by_len) fprintf(f, "%" PRIu64 " %" PRIu64 " %016" PRIx64 "\n", kv.first, kv.second.first, kv.second.second);
        fclose(f);
    }
    printf("{\"records\": %" PRIu64 ", \"bases\": %" PRIu64 ", \"sum\": \"%016" PRIx64 "\", \"xor\": \"%016" PRIx64 "\"}\n",
           records, bases, sum, x);
    return 0;
}
