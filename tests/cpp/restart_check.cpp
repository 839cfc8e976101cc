// CPU check of the reader restart the CLI does in two-pass mode (src/TGSFilter.cpp: reader_stop, drain, join, start again):
// restart_check <file> <batch_bases> <stop_after_batches>; prints what each round delivered.
#include "../../src/pipeline.hpp"
#include <cstdio>
int main(int argc, char **argv) {
    const std::string path = argv[1];
    const uint64_t chunk = strtoull(argv[2], nullptr, 10);
    const int stop_after = atoi(argv[3]);
    ingest::BatchPool pool;
    for (int round = 0; round < 2; ++round) {
        ingest::Queue<std::unique_ptr<ingest::RawBatch>> q(4);
        std::atomic<bool> stop{false};
        std::thread t(ingest::reader_main, path, true, chunk, &q, &pool, &stop, false);
        uint64_t n = 0, bases = 0; int batches = 0;
        bool done = false;
        while (!done) {
            auto rb = q.pop();
            if (!rb) { done = true; break; }
            n += rb->n(); bases += rb->bases.size(); ++batches;
            pool.put(std::move(rb));
            if (round == 0 && batches == stop_after) {
                stop.store(true);
                while (q.pop()) {}
                done = true;
            }
        }
        t.join();
        printf("round %d: %llu reads %llu bases %d batches\n", round, (unsigned long long)n, (unsigned long long)bases, batches);
    }
    return 0;
}
