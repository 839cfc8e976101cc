// CPU check of fastgz::MultiMemberReader against the sequential GzReader:
// multimember_check <file.gz> <threads>; prints sizes, whether the reader fell back to streaming.
#include "../../src/inflate.hpp"
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::vector<uint8_t> a, b, tmp(1u << 20);
    bool fa, fb, fell;
    {
        fastgz::GzReader r(argv[1]);
        if (!r.ok()) return 3;
        size_t n;
        while ((n = r.read(tmp.data(), tmp.size())) > 0) a.insert(a.end(), tmp.begin(), tmp.begin() + n);
        fa = r.failed();
    }
    {
        fastgz::MultiMemberReader r(argv[1], atoi(argv[2]));
        if (!r.ok()) return 3;
        size_t n;
        while ((n = r.read(tmp.data(), 700001)) > 0) b.insert(b.end(), tmp.begin(), tmp.begin() + n);
        fb = r.failed();
        fell = r.fell_back();
    }
    printf("%zu %zu fail %d %d fallback %d multi %d\n", a.size(), b.size(), (int)fa, (int)fb, (int)fell,
           (int)fastgz::MultiMemberReader::is_multi_member(argv[1]));
    return (a == b && fa == fb) ? 0 : 1;
}
