/*
 * ref_driver.cpp — TEST INFRASTRUCTURE.  A thin command-line driver around the UNMODIFIED
 * reference translation unit, compiled from where it lies (/root/reference/src/TGSFilter.cpp, via
 * -I; no reference source is copied into this repo).  It lets the tests call the reference's own
 * edlibAlign and its own TGSFilterTask::filter_sequence worker body on in-memory inputs, so the
 * CPU restatement (oracle/tgsf_oracle.c) and the golden fixtures are pinned against the real code.
 *
 * Built only where /root/reference exists (oracle/Makefile target `ref`); output goes to
 * oracle/_ref/ref_driver (git-ignored).  Nothing in the product path uses it.
 *
 *   ref_driver edlib   <in.bin> <out.bin>   batch of edlibAlign(HW, PATH) calls
 *   ref_driver perread <in.bin> <out.bin>   one TGSFilterTask worker over a batch of reads
 *   ref_driver prepass <in.bin> <out.bin>   CheckBaseContent + adapterSearch on given read ends
 */
#include <algorithm>
#include <atomic>
#include <bitset>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <regex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <unistd.h>

/* The worker body and the pre-pass bodies are private members; open them up for the driver. */
#define private public
#define main tgsfilter_reference_main
#include "src/TGSFilter.cpp"
#undef main
#undef private

namespace {

struct Reader {
    std::vector<unsigned char> buf;
    size_t pos = 0;
    explicit Reader(const char *path) {
        FILE *f = fopen(path, "rb");
        if (!f) { perror(path); exit(2); }
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize((size_t)n);
        if (n && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { perror("read"); exit(2); }
        fclose(f);
    }
    template <typename T> T get() {
        T v;
        memcpy(&v, buf.data() + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    std::string str() {
        uint32_t n = get<uint32_t>();
        std::string s((const char *)buf.data() + pos, n);
        pos += n;
        return s;
    }
};

struct Writer {
    FILE *f;
    explicit Writer(const char *path) {
        f = fopen(path, "wb");
        if (!f) { perror(path); exit(2); }
    }
    ~Writer() { fclose(f); }
    template <typename T> void put(T v) { fwrite(&v, sizeof(T), 1, f); }
    void str(const std::string &s) {
        put<uint32_t>((uint32_t)s.size());
        fwrite(s.data(), 1, s.size(), f);
    }
    void table(const std::vector<std::vector<uint64_t>> &t) {
        put<uint32_t>((uint32_t)t.size());
        for (const auto &row : t)
            for (int j = 0; j < 5; j++) put<uint64_t>(j < (int)row.size() ? row[j] : 0);
    }
};

int run_edlib(const char *in, const char *out) {
    Reader r(in);
    Writer w(out);
    uint32_t n = r.get<uint32_t>();
    for (uint32_t i = 0; i < n; i++) {
        int32_t k = r.get<int32_t>();
        std::string q = r.str();
        std::string t = r.str();
        EdlibAlignResult res = edlibAlign(q.data(), (int)q.size(), t.data(), (int)t.size(),
                                          edlibNewAlignConfig(k, EDLIB_MODE_HW, EDLIB_TASK_PATH, NULL, 0));
        w.put<int32_t>(res.editDistance);
        w.put<int32_t>(res.numLocations);
        w.put<int32_t>(res.alignmentLength);
        for (int j = 0; j < res.numLocations; j++) {
            w.put<int32_t>(res.startLocations ? res.startLocations[j] : 0);
            w.put<int32_t>(res.endLocations[j]);
        }
        edlibFreeAlignResult(res);
    }
    return 0;
}

void read_params(Reader &r, Para_A24 *P) {
    P->MinLen = r.get<int32_t>();
    P->MaxLen = r.get<int32_t>();
    P->MinQ = r.get<float>();
    P->MaxQ = r.get<float>();
    P->BCLen = r.get<int32_t>();
    P->HeadTrim = r.get<int32_t>();
    P->TailTrim = r.get<int32_t>();
    P->EndLen = r.get<int32_t>();
    P->EndMatchLen = r.get<int32_t>();
    P->MidMatchLen = r.get<int32_t>();
    P->ExtraLen = r.get<int32_t>();
    P->EndSim = r.get<float>();
    P->MidSim = r.get<float>();
    P->Kmer = r.get<int32_t>();
    P->MinRepeat = r.get<int32_t>();
    qType = r.get<int32_t>();
    uint32_t flags = r.get<uint32_t>();
    P->Filter = (flags & 1u) != 0;
    P->OnlyQC = (flags & 2u) != 0;
    P->discard = (flags & 4u) != 0;
    P->Outfq = r.get<int32_t>();
    P->n_thread = 1;
    adapters.clear();
    int32_t na = r.get<int32_t>();
    for (int i = 0; i < na; i++) adapters.insert(r.str());
}

int run_perread(const char *in, const char *out) {
    Reader r(in);
    Para_A24 *P = new Para_A24;
    read_params(r, P);
    TGSFilterTask task(P);
    task.worker_count = 1 << 30; /* no 1 ms back-pressure sleeps: nobody drains the queue */
    uint32_t n = r.get<uint32_t>();
    for (uint32_t i = 0; i < n; i++) {
        std::string name = r.str();
        std::string seq = r.str();
        std::string qual = r.str();
        task.input_queue.enqueue(std::make_tuple(name, seq, qual));
        task.inQueueSize++;
        task.readNum++;
    }
    task.read_done = true;
    task.filter_sequence(0);

    Writer w(out);
    for (int j = 0; j < 17; j++) w.put<uint64_t>(task.DropInfo[0][j]);
    for (int j = 0; j < 256; j++) w.put<uint64_t>(task.rawDiffQualReadsBases[0][j]);
    for (int j = 0; j < 256; j++) w.put<uint64_t>(task.cleanDiffQualReadsBases[0][j]);
    w.table(task.raw5pBaseCounts[0]);
    w.table(task.raw5pBaseQual[0]);
    w.table(task.raw3pBaseCounts[0]);
    w.table(task.raw3pBaseQual[0]);
    w.table(task.clean5pBaseCounts[0]);
    w.table(task.clean5pBaseQual[0]);
    w.table(task.clean3pBaseCounts[0]);
    w.table(task.clean3pBaseQual[0]);
    w.table(task.rawBaseCounts[0]);
    w.table(task.rawBaseQual[0]);
    w.table(task.cleanBaseCounts[0]);
    w.table(task.cleanBaseQual[0]);
    std::vector<std::tuple<std::string, std::string, uint64_t>> recs;
    std::tuple<std::string, std::string, uint64_t> info;
    while (task.output_queue.try_dequeue(info)) recs.push_back(info);
    w.put<uint32_t>((uint32_t)recs.size());
    for (auto &rec : recs) {
        w.str(std::get<0>(rec));
        w.str(std::get<1>(rec));
        w.put<int32_t>((int32_t)std::get<2>(rec));
    }
    return 0;
}

int run_prepass(const char *in, const char *out) {
    Reader r(in);
    Para_A24 *P = new Para_A24;
    P->EndLen = r.get<int32_t>();
    P->BCLen = r.get<int32_t>();
    P->EndBias = r.get<float>();
    P->MidSim = r.get<float>();
    uint32_t n = r.get<uint32_t>();
    uint32_t row_len = r.get<uint32_t>();
    std::vector<std::string> lib;
    int32_t nlib = r.get<int32_t>();
    for (int i = 0; i < nlib; i++) lib.push_back(r.str());
    GetFilterParameterTask task(P, lib);
    task.checkLen = (int)row_len;
    task.seqNum = (int)n;
    for (uint32_t i = 0; i < n; i++) task.seqs5p.push_back(r.str());
    for (uint32_t i = 0; i < n; i++) task.seqs3p.push_back(r.str());
    /* sequential, 5p before 3p: the trim5p clamp race (T.cpp:1136-1138) is then deterministic */
    task.CheckBaseContent(task.seqs5p, "5p");
    task.CheckBaseContent(task.seqs3p, "3p");
    task.adapterSearch(task.seqs5p, "5p");
    task.adapterSearch(task.seqs3p, "3p");
    Writer w(out);
    w.put<int32_t>(task.trim5p);
    w.put<int32_t>(task.trim3p);
    w.str(task.adapter5p);
    w.str(task.adapter3p);
    w.put<float>(task.adapterDep5p);
    w.put<float>(task.adapterDep3p);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: ref_driver edlib|perread|prepass <in.bin> <out.bin>\n");
        return 2;
    }
    std::string mode = argv[1];
    for (int i = 0; i < 256; i++) complement[i] = 'N';
    if (mode == "edlib") return run_edlib(argv[2], argv[3]);
    if (mode == "perread") return run_perread(argv[2], argv[3]);
    if (mode == "prepass") return run_prepass(argv[2], argv[3]);
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
