// CPU check of src/inflate.hpp against zlib: inflate_check <file.gz> [read_size]
// exit 0 when both decoders produce the same bytes (and agree on failure); prints sizes and MB/s.
#include "../../src/inflate.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const size_t rs = argc > 2 ? strtoull(argv[2], nullptr, 10) : (1u << 20);
    std::vector<uint8_t> a, b, tmp(rs);
    auto t0 = std::chrono::steady_clock::now();
    bool fail_a = false;
    {
        fastgz::GzReader r(argv[1]);
        if (!r.ok()) return 3;
        size_t n;
        while ((n = r.read(tmp.data(), rs)) > 0) a.insert(a.end(), tmp.begin(), tmp.begin() + n);
        fail_a = r.failed();
        if (fail_a) fprintf(stderr, "fastgz: %s\n", r.error());
    }
    auto t1 = std::chrono::steady_clock::now();
    bool fail_b = false;
    {
        gzFile f = gzopen(argv[1], "rb");
        if (!f) return 3;
        gzbuffer(f, 1 << 20);
        int n;
        while ((n = gzread(f, tmp.data(), (unsigned)rs)) > 0) b.insert(b.end(), tmp.begin(), tmp.begin() + n);
        int err = 0;
        gzerror(f, &err);
        fail_b = n < 0 || (err != Z_OK && err != Z_STREAM_END);
        gzclose(f);
    }
    auto t2 = std::chrono::steady_clock::now();
    const double da = std::chrono::duration<double>(t1 - t0).count(), db = std::chrono::duration<double>(t2 - t1).count();
    printf("%zu %zu fail %d %d  fastgz %.0f MB/s  zlib %.0f MB/s\n", a.size(), b.size(), (int)fail_a, (int)fail_b,
           a.size() / da / 1e6, b.size() / db / 1e6);
    if (fail_a != fail_b) return 1;
    if (fail_a) return 0; // both reject the stream (the amount delivered before the error may differ)
    return (a == b) ? 0 : 1;
}
