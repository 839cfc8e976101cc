/*
 * Counter-block layout shared by libtgsf_cuda, the CPU oracle and the hosts: a pure function of
 * (bc_len, max_bins), so every rank of a multi-GPU job derives the same layout and the blocks can
 * be summed with one allreduce.  See tgsf_counter_layout in tgsf.h for the member meaning.
 */
#ifndef TGSF_LAYOUT_H_
#define TGSF_LAYOUT_H_

#include "tgsf.h"

#define TGSF_DEFAULT_MAX_READ_LEN (4 * 1024 * 1024)

static inline uint32_t tgsf_bins_for_len(int64_t max_read_len) {
    if (max_read_len <= 0) max_read_len = TGSF_DEFAULT_MAX_READ_LEN;
    return (uint32_t)(max_read_len / 100 + 1); /* vectorSize = int(seqLen/100)+1, T.cpp:1445 */
}

static inline void tgsf_make_layout(int32_t bc_len, int32_t max_read_len, tgsf_counter_layout *L) {
    uint32_t bc = bc_len > 0 ? (uint32_t)bc_len : 0u;
    uint32_t bins = tgsf_bins_for_len(max_read_len);
    uint32_t o = 0;
    L->bc_len = bc;
    L->max_bins = bins;
    L->drop_info = o; o += TGSF_DROPINFO_N;
    L->raw_hist = o; o += TGSF_QUAL_HIST_N;
    L->clean_hist = o; o += TGSF_QUAL_HIST_N;
    o = (o + 7u) & ~7u; /* keep the tables 64-byte aligned */
    L->raw5p_cnt = o; o += bc * 5;
    L->raw5p_qual = o; o += bc * 5;
    L->raw3p_cnt = o; o += bc * 5;
    L->raw3p_qual = o; o += bc * 5;
    L->clean5p_cnt = o; o += bc * 5;
    L->clean5p_qual = o; o += bc * 5;
    L->clean3p_cnt = o; o += bc * 5;
    L->clean3p_qual = o; o += bc * 5;
    o = (o + 7u) & ~7u;
    L->raw_bin_cnt = o; o += bins * 5;
    L->raw_bin_qual = o; o += bins * 5;
    L->clean_bin_cnt = o; o += bins * 5;
    L->clean_bin_qual = o; o += bins * 5;
    L->n_u64 = o;
}

#endif /* TGSF_LAYOUT_H_ */
