// CPU check of src/pipeline.hpp: the parallel chunk parser must deliver the same records, in the
// same order, as the serial reader.  usage: ingest_check <file> <fastq 0|1> <chunk_bytes> <threads>
// INGEST_ONLY=serial|parallel: run one side only (digest printed, exit 0); INGEST_SAMBAM=1: the serial side reads BAM/SAM.
#include "../../src/pipeline.hpp"
#include <cstdio>

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
struct Digest {
    uint64_t h = 1469598103934665603ull, n = 0, bases = 0;
    void add(const ingest::RawBatch &b) {
        if (getenv("INGEST_ONLY") && !getenv("INGEST_HASH")) { n += b.n(); bases += b.bases.size(); return; }
        for (size_t i = 0; i < b.n(); ++i) {
            h = fnv(h, b.names[i].data(), b.names[i].size());
            h = fnv(h, "\n", 1);
            const uint64_t s = b.offsets[i], e = b.offsets[i + 1];
            h = fnv(h, b.bases.data() + s, e - s);
            h = fnv(h, "\n", 1);
            if (!b.quals.empty()) h = fnv(h, b.quals.data() + s, e - s);
            ++n;
            bases += e - s;
        }
    }
};

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const std::string path = argv[1];
    const bool fastq = atoi(argv[2]) != 0;
    const uint64_t chunk = strtoull(argv[3], nullptr, 10);
    const int threads = atoi(argv[4]);
    Digest a, b;
    const char *only = getenv("INGEST_ONLY"); // timing aid: "serial" / "parallel" (counts only)
    if (!only || !strcmp(only, "serial")) {
        ingest::Queue<std::unique_ptr<ingest::RawBatch>> q(4);
        ingest::BatchPool pool;
        std::thread t([&] { ingest::reader_main(path, fastq, chunk, &q, &pool, nullptr, getenv("INGEST_SAMBAM") != nullptr); });
        while (auto rb = q.pop()) { a.add(*rb); pool.put(std::move(rb)); }
        t.join();
    }
    if (!only || !strcmp(only, "parallel")) {
        ingest::BatchPool pool;
        ingest::ParallelReader pr(path, fastq, chunk, threads, &pool);
        if (!pr.ok()) return 3;
        while (auto rb = pr.pop()) { b.add(*rb); pool.put(std::move(rb)); }
    }
    printf("%llu %llu %llx %llu %llu %llx\n", (unsigned long long)a.n, (unsigned long long)a.bases, (unsigned long long)a.h,
           (unsigned long long)b.n, (unsigned long long)b.bases, (unsigned long long)b.h);
    if (only) return 0;
    return (a.n == b.n && a.bases == b.bases && a.h == b.h) ? 0 : 1;
}
