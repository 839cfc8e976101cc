// mkcram: SAM / BAM -> CRAM without a reference sequence (unaligned reads).  TEST INFRASTRUCTURE: only used by
// tests/golden/make_cram.py to (re)generate the committed CRAM fixture, built by hand against an htslib prefix:
//   g++ -O2 -no-pie -I$HTS_DIR/include tests/cpp/mkcram.cpp $HTS_DIR/lib/libhts.a $HTS_DIR/lib/libdeflate.a \
//       $HTS_DIR/lib/libisal.a -lz -pthread -o mkcram
#include <cstdio>

#include "sam.h"

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: mkcram in.bam out.cram\n"); return 2; }
    htsFile *in = hts_open(argv[1], "r");
    htsFile *out = hts_open(argv[2], "wc");
    if (!in || !out) { fprintf(stderr, "open failed\n"); return 1; }
    hts_set_log_level(HTS_LOG_ERROR);
    hts_set_opt(out, CRAM_OPT_NO_REF, 1);
    bam_hdr_t *h = sam_hdr_read(in);
    if (!h || sam_hdr_write(out, h) < 0) { fprintf(stderr, "header failed\n"); return 1; }
    bam1_t *b = bam_init1();
    int n = 0;
    while (sam_read1(in, h, b) >= 0) {
        b->core.flag |= BAM_FUNMAP; // unaligned: no CIGAR / reference checks
        b->core.tid = b->core.mtid = -1;
        b->core.pos = b->core.mpos = -1;
        if (sam_write1(out, h, b) < 0) { fprintf(stderr, "write failed\n"); return 1; }
        ++n;
    }
    bam_destroy1(b);
    bam_hdr_destroy(h);
    hts_close(out);
    hts_close(in);
    fprintf(stderr, "%d records\n", n);
    return 0;
}
