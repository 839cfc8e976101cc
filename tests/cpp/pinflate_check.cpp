// CPU check of fastgz::SingleStreamReader (src/pinflate.hpp) against the sequential GzReader:
// pinflate_check <file.gz> <threads> [span_bytes] [read_size]
// exit 0 when both deliver the same bytes and agree on failure; prints sizes, MB/s, spans and repairs.
#include "../../src/pinflate.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const size_t span = argc > 3 ? strtoull(argv[3], nullptr, 10) : 0;
    const size_t rs = argc > 4 ? strtoull(argv[4], nullptr, 10) : (1u << 20);
    std::vector<uint8_t> a, b, tmp(rs);
    bool fail_a = false, fail_b = false;
    {
        struct stat st;
        if (stat(argv[1], &st) == 0) { a.reserve((size_t)st.st_size * 5); b.reserve((size_t)st.st_size * 5); }
    }
    auto t0 = std::chrono::steady_clock::now();
    {
        fastgz::GzReader r(argv[1]);
        if (!r.ok()) return 3;
        size_t n;
        while ((n = r.read(tmp.data(), rs)) > 0) a.insert(a.end(), tmp.begin(), tmp.begin() + n);
        fail_a = r.failed();
        if (fail_a) fprintf(stderr, "serial: %s\n", r.error());
    }
    auto t1 = std::chrono::steady_clock::now();
    size_t spans = 0, repairs = 0, speculated = 0;
    {
        fastgz::SingleStreamReader r(argv[1], atoi(argv[2]), span);
        if (!r.ok()) return 3;
        size_t n;
        while ((n = r.read(tmp.data(), rs)) > 0) b.insert(b.end(), tmp.begin(), tmp.begin() + n);
        fail_b = r.failed();
        if (fail_b) fprintf(stderr, "parallel: %s\n", r.error());
        spans = r.spans();
        repairs = r.repairs();
        speculated = r.speculated();
        double ph[3];
        r.phase_seconds(ph);
        fprintf(stderr, "worker seconds: search %.3f decode %.3f resolve+crc %.3f\n", ph[0], ph[1], ph[2]);
    }
    auto t2 = std::chrono::steady_clock::now();
    const double da = std::chrono::duration<double>(t1 - t0).count(), db = std::chrono::duration<double>(t2 - t1).count();
    printf("%zu %zu fail %d %d  serial %.0f MB/s  parallel %.0f MB/s  speculated %zu spans %zu repairs %zu\n", a.size(), b.size(), (int)fail_a,
           (int)fail_b, a.size() / da / 1e6, b.size() / db / 1e6, speculated, spans, repairs);
    if (fail_a != fail_b) return 1;
    if (fail_a) return (b.size() <= a.size() + (4u << 20) && std::equal(b.begin(), b.begin() + std::min(a.size(), b.size()), a.begin())) ? 0 : 1;
    return (a == b) ? 0 : 1;
}
