// Building blocks of the per-record gzip encoder ("optionally GPU deflate").
//
// The reference writes one gzip member per output record (DeflateCompress, T.cpp:786-812, libdeflate).
// Here a member is assembled from
//   gzip header | stored block: header line | dynamic-Huffman block: bases | dynamic-Huffman block:
//   "\n+\n" + qualities + "\n" (final) | CRC-32, ISIZE
// with literal-only Huffman blocks (per-record members cannot reference other records, and on noisy long
// reads literal coding is as small as zlib's match search).  The serial parts — code
// lengths, canonical codes, block header — are plain functions compiled for host and device: the host
// build is the checker of the format (tests decode its members with zlib), the kernels in gzenc.cuh use
// the same functions for the tables and spread the symbol coding over the threads of a CTA.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GZ_HD __host__ __device__
#else
#define GZ_HD
#endif

#define GZ_NSYM 257      // literals 0..255 + end-of-block
#define GZ_MAX_BITS 15
#define GZ_HDR_MAX_BYTES 160 // upper bound of one dynamic block header written by gz_write_dyn_header

struct GzBitWriter { // LSB-first bit stream into a zero-initialised byte buffer
    uint8_t *out;
    uint64_t acc;
    uint32_t nacc;
    uint64_t pos; // bytes flushed
};
GZ_HD inline void gz_bw_init(GzBitWriter &w, uint8_t *out) { w.out = out; w.acc = 0; w.nacc = 0; w.pos = 0; }
GZ_HD inline void gz_bw_put(GzBitWriter &w, uint32_t bits, uint32_t n) { // n <= 25
    w.acc |= (uint64_t)bits << w.nacc;
    w.nacc += n;
    while (w.nacc >= 8) { w.out[w.pos++] = (uint8_t)w.acc; w.acc >>= 8; w.nacc -= 8; }
}
GZ_HD inline uint64_t gz_bw_bits(const GzBitWriter &w) { return w.pos * 8 + w.nacc; }
GZ_HD inline void gz_bw_flush(GzBitWriter &w) { if (w.nacc) { w.out[w.pos++] = (uint8_t)w.acc; w.acc = 0; w.nacc = 0; } }

// Huffman code lengths (<= limit) for n symbols; scratch: 5 * n uint32 words.  Symbols with freq 0 get length 0;
// a single used symbol gets length 1.  Over-long codes are avoided by halving the frequencies and rebuilding.
GZ_HD inline void gz_huff_lengths(const uint32_t *freq, int n, int limit, uint8_t *len, uint32_t *scratch) {
    uint32_t *f = scratch;                 // frequencies of the used symbols, sorted ascending
    uint32_t *sym = scratch + n;           // their symbol numbers
    uint32_t *leaf_parent = scratch + 2 * n; // internal node each leaf hangs under
    uint32_t *node_w = scratch + 3 * n;    // weight of internal node k, later its depth
    uint32_t *node_parent = scratch + 4 * n;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        len[i] = 0;
        if (freq[i]) { f[m] = freq[i]; sym[m] = (uint32_t)i; ++m; }
    }
    if (m == 0) return;
    if (m == 1) { len[sym[0]] = 1; return; }
    while (true) {
        for (int i = 1; i < m; ++i) { // insertion sort by (freq, symbol)
            const uint32_t fi = f[i], si = sym[i];
            int j = i - 1;
            while (j >= 0 && (f[j] > fi || (f[j] == fi && sym[j] > si))) { f[j + 1] = f[j]; sym[j + 1] = sym[j]; --j; }
            f[j + 1] = fi;
            sym[j + 1] = si;
        }
        // two-queue merge: sorted leaves and internal nodes in creation order (their weights are non-decreasing)
        int leaf = 0, inode = 0, made = 0;
        while (made < m - 1) {
            uint32_t w = 0;
            for (int t = 0; t < 2; ++t) {
                if (leaf < m && (inode >= made || f[leaf] <= node_w[inode])) { w += f[leaf]; leaf_parent[leaf++] = (uint32_t)made; }
                else { w += node_w[inode]; node_parent[inode++] = (uint32_t)made; }
            }
            node_w[made++] = w;
        }
        node_w[m - 2] = 0; // depth of the root; parents are always created after their children
        for (int k = m - 3; k >= 0; --k) node_w[k] = node_w[node_parent[k]] + 1;
        int maxd = 0;
        for (int i = 0; i < m; ++i) {
            const int d = (int)node_w[leaf_parent[i]] + 1;
            len[sym[i]] = (uint8_t)d;
            if (d > maxd) maxd = d;
        }
        if (maxd <= limit) return;
        for (int i = 0; i < m; ++i) f[i] = (f[i] + 1) >> 1;
    }
}

// canonical codes, bit-reversed so that they can be emitted LSB-first
GZ_HD inline void gz_huff_codes(const uint8_t *len, int n, uint16_t *code) {
    uint32_t count[GZ_MAX_BITS + 1];
    for (int i = 0; i <= GZ_MAX_BITS; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) count[len[i]]++;
    count[0] = 0;
    uint32_t next[GZ_MAX_BITS + 2];
    uint32_t c = 0;
    for (int l = 1; l <= GZ_MAX_BITS; ++l) { c = (c + count[l - 1]) << 1; next[l] = c; }
    for (int i = 0; i < n; ++i) {
        const int l = len[i];
        if (!l) { code[i] = 0; continue; }
        uint32_t v = next[l]++, r = 0;
        for (int b = 0; b < l; ++b) r |= ((v >> b) & 1u) << (l - 1 - b);
        code[i] = (uint16_t)r;
    }
}

// Header of a dynamic block whose literal/length code has the lengths lit_len[0..256] (no length symbols) and,
// like zlib, two distance codes of one bit each.  Code lengths are sent with zero-run symbols 17/18 only.
// scratch: 5 * 19 + 140 uint32 words.  The writer's position tells the size.
GZ_HD inline void gz_write_dyn_header(GzBitWriter &w, const uint8_t *lit_len, bool final_block, uint32_t *scratch) {
    // sequence of code-length symbols: (symbol, extra value) pairs
    uint16_t *seq = (uint16_t *)(scratch + 5 * 19); // up to 259 entries: symbol | extra << 8
    int ns = 0;
    uint32_t clf[19];
    for (int i = 0; i < 19; ++i) clf[i] = 0;
    const int total = GZ_NSYM + 2;
    int i = 0;
    while (i < total) {
        const int l = i < GZ_NSYM ? lit_len[i] : 1; // the two distance codes
        if (l != 0) { seq[ns++] = (uint16_t)l; clf[l]++; ++i; continue; }
        int run = 1;
        while (i + run < GZ_NSYM && lit_len[i + run] == 0) ++run; // zero runs never reach the distance lengths
        i += run;
        while (run >= 11) { const int r = run > 138 ? 138 : run; seq[ns++] = (uint16_t)(18 | ((r - 11) << 8)); clf[18]++; run -= r; }
        if (run >= 3) { seq[ns++] = (uint16_t)(17 | ((run - 3) << 8)); clf[17]++; run = 0; }
        while (run-- > 0) { seq[ns++] = 0; clf[0]++; }
    }
    uint8_t cll[19];
    uint16_t clc[19];
    gz_huff_lengths(clf, 19, 7, cll, scratch);
    gz_huff_codes(cll, 19, clc);
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && cll[order[hclen - 1]] == 0) --hclen;
    gz_bw_put(w, final_block ? 1u : 0u, 1);
    gz_bw_put(w, 2u, 2);                     // BTYPE = 10
    gz_bw_put(w, (uint32_t)(GZ_NSYM - 257), 5); // HLIT
    gz_bw_put(w, 1u, 5);                     // HDIST: two distance codes
    gz_bw_put(w, (uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; ++k) gz_bw_put(w, cll[order[k]], 3);
    for (int k = 0; k < ns; ++k) {
        const int s = seq[k] & 0xFF, x = seq[k] >> 8;
        gz_bw_put(w, clc[s], cll[s]);
        if (s == 17) gz_bw_put(w, (uint32_t)x, 3);
        else if (s == 18) gz_bw_put(w, (uint32_t)x, 7);
    }
}
