// GPU side of the per-record gzip encoder (see gzenc_core.h for the member layout).
//
// k_gz_encode: one CTA per emitted piece (persistent grid).  The piece's bases and qualities are already in
// HBM, so compressing them here costs two more passes over L2/HBM-resident bytes and what crosses PCIe
// shrinks from 2 B/base (records formatted on the host) to ~0.8 B/base of finished deflate blocks.
//   1. symbol histograms of the two blocks (bases; "\n+\n" + qualities + "\n"): per-thread registers for the
//      four bases, per-warp shared sub-histograms for everything else;
//   2. thread 0: code lengths, canonical codes and the two block headers (gzenc_core.h, serial: a few
//      thousand operations for the ~5 + ~45 symbols of a FASTQ record);
//   3. every thread sums the code lengths of its contiguous symbol range, a block scan turns them into bit
//      offsets, thread 0 reserves the member's bytes in the batch blob with one atomicAdd (members are laid
//      out in completion order; the per-piece span table says where), and every thread writes its bits:
//      whole 32-bit words it owns with plain stores, the two boundary words with atomicOr (blob is zeroed).
// The host adds the gzip header, the stored block with the header line, CRC-32 and ISIZE (src/TGSFilter.cpp).
//
// Measured on a 0.47 Gbase ONT / 0.30 Gbase HiFi batch (this version: 8.4 / 6.1 ms) and not kept: warp-shuffle
// scan instead of the shared-memory one (12.4 / 10.3 ms: 48 instead of 40 registers, one resident CTA less), the
// two blocks' tables on two threads (11.3 / 8.6 ms, 64 registers), __launch_bounds__(256, 8) (12.0 / 5.6 ms).
#pragma once
#include "common.cuh"
#include "gzenc_core.h"

#define GZ_THREADS 256
#define GZ_WARPS (GZ_THREADS / 32)

struct GzSpan { // per piece, same index as the pieces array
    u64 offset; // byte offset of the deflate blocks in the blob
    u32 bytes;  // 0: piece not emitted
    u32 reserved;
};

// LSB-first bit output of one thread into its bit range of a zero-initialised blob: 32-bit words that lie completely
// inside the range are written with plain stores, the first and the last (shared with the neighbours) with atomicOr
struct GzBitOut {
    u32 *words;
    u64 acc;     // pending bits
    u32 nacc;    // number of pending bits
    u64 word;    // index of the word the pending bits start in
    u32 shift;   // bit offset inside that word of the first pending bit (only for the very first word)
    bool first;
};

static __device__ __forceinline__ void gz_out_begin(GzBitOut &o, u32 *words, u64 bit) {
    o.words = words;
    o.word = bit >> 5;
    o.shift = (u32)(bit & 31);
    o.acc = 0;
    o.nacc = o.shift; // pretend the bits below the start are (zero) pending bits of the first word
    o.first = true;
}
static __device__ __forceinline__ void gz_out_put(GzBitOut &o, u32 v, u32 n) {
    o.acc |= (u64)v << o.nacc;
    o.nacc += n;
    if (o.nacc >= 32) {
        const u32 w = (u32)o.acc;
        if (o.first) { atomicOr(o.words + o.word, w); o.first = false; } // shared with the previous writer
        else o.words[o.word] = w;
        o.acc >>= 32;
        o.nacc -= 32;
        ++o.word;
    }
}
static __device__ __forceinline__ void gz_out_end(GzBitOut &o) {
    if (o.nacc) atomicOr(o.words + o.word, (u32)o.acc); // shared with the next writer (or padding)
}

__global__ void __launch_bounds__(GZ_THREADS)
k_gz_encode(DevBatch B, const tgsf_piece *__restrict__ pieces, const u32 *__restrict__ n_pieces_ptr, int fasta,
            u32 *__restrict__ blob_words, u64 blob_cap_bytes, unsigned long long *__restrict__ cursor,
            GzSpan *__restrict__ spans, u32 *__restrict__ overflow, const u32 *__restrict__ dev_status) {
    if (*dev_status != DEV_STATUS_OK) return;
    __shared__ u32 s_hist[2][GZ_NSYM];            // block 0: bases, block 1: quality line
    __shared__ u32 s_sub[GZ_WARPS][256];          // per-warp sub-histograms
    __shared__ uint8_t s_len[2][GZ_NSYM];
    __shared__ uint16_t s_code[2][GZ_NSYM];
    __shared__ u32 s_scratch[5 * GZ_NSYM + 160];
    __shared__ __align__(8) uint8_t s_hdr[2][GZ_HDR_MAX_BYTES + 8];
    __shared__ u32 s_hdr_bits[2];
    __shared__ u64 s_scan[GZ_THREADS];
    __shared__ u64 s_base_bit[3]; // bit offsets: start of block 0 symbols, start of block 1 header, total
    __shared__ u64 s_member_off;
    const u32 n_pieces = *n_pieces_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_q = !fasta && B.quals != nullptr;

    for (u32 pi = blockIdx.x; pi < n_pieces; pi += gridDim.x) {
        const tgsf_piece pc = pieces[pi];
        if (pc.status != TGSF_PIECE_EMIT) {
            if (threadIdx.x == 0) spans[pi] = GzSpan{0, 0, 0};
            continue;
        }
        const u64 s0 = B.offsets[pc.read] + (u64)pc.start;
        const u32 L = (u32)pc.len;
        const uint8_t *seq = B.bases + s0;
        const uint8_t *qual = has_q ? B.quals + s0 : nullptr;
        // ---- 1. histograms ----
        for (int i = threadIdx.x; i < 2 * GZ_NSYM; i += GZ_THREADS) (&s_hist[0][0])[i] = 0;
        for (int blk = 0; blk < (has_q ? 2 : 1); ++blk) {
            const uint8_t *src = blk ? qual : seq;
            for (int i = lane; i < 256; i += 32) s_sub[warp][i] = 0;
            __syncwarp();
            u32 cA = 0, cC = 0, cG = 0, cT = 0;
            for (u32 i = threadIdx.x; i < L; i += GZ_THREADS) {
                const uint8_t c = src[i];
                if (!blk && c == 'A') ++cA;
                else if (!blk && c == 'C') ++cC;
                else if (!blk && c == 'G') ++cG;
                else if (!blk && c == 'T') ++cT;
                else atomicAdd(&s_sub[warp][c], 1u);
            }
            if (!blk) {
                cA = __reduce_add_sync(0xffffffffu, cA);
                cC = __reduce_add_sync(0xffffffffu, cC);
                cG = __reduce_add_sync(0xffffffffu, cG);
                cT = __reduce_add_sync(0xffffffffu, cT);
                if (lane == 0) {
                    s_sub[warp]['A'] += cA; s_sub[warp]['C'] += cC; s_sub[warp]['G'] += cG; s_sub[warp]['T'] += cT;
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 256; i += GZ_THREADS) {
                u32 t = 0;
#pragma unroll
                for (int w = 0; w < GZ_WARPS; ++w) t += s_sub[w][i];
                s_hist[blk][i] = t;
            }
            __syncthreads();
        }
        // ---- 2. tables and headers (serial) ----
        if (threadIdx.x == 0) {
            if (has_q) { s_hist[1]['\n'] += 3; s_hist[1]['+'] += 1; } // "\n+\n" ... "\n"
            else s_hist[0]['\n'] += 1;                                // FASTA: bases + "\n" in one final block
            const int nblk = has_q ? 2 : 1;
            for (int blk = 0; blk < nblk; ++blk) {
                s_hist[blk][256] = 1;
                gz_huff_lengths(s_hist[blk], GZ_NSYM, GZ_MAX_BITS, s_len[blk], s_scratch);
                gz_huff_codes(s_len[blk], GZ_NSYM, s_code[blk]);
                for (int i = 0; i < GZ_HDR_MAX_BYTES + 8; ++i) s_hdr[blk][i] = 0;
                GzBitWriter w;
                gz_bw_init(w, s_hdr[blk]);
                gz_write_dyn_header(w, s_len[blk], blk == nblk - 1, s_scratch);
                s_hdr_bits[blk] = (u32)gz_bw_bits(w);
                gz_bw_flush(w);
            }
        }
        __syncthreads();
        // ---- 3. bit offsets ----
        // symbol stream of block 0: L bases (+ "\n" for FASTA); block 1: "\n+\n", L qualities, "\n".
        // Thread t owns bases/qualities [t * per, (t + 1) * per); the literal extras are written by thread 0.
        const u32 per = (L + GZ_THREADS - 1) / GZ_THREADS;
        const u32 i0 = min(threadIdx.x * per, L), i1 = min(i0 + per, L);
        u64 bits0 = 0, bits1 = 0;
        for (u32 i = i0; i < i1; ++i) {
            bits0 += s_len[0][seq[i]];
            if (has_q) bits1 += s_len[1][qual[i]];
        }
        // exclusive scans over threads (two streams): simple shared-memory Hillis-Steele on 64-bit values
        u64 off0, off1 = 0, tot0, tot1 = 0;
        {
            s_scan[threadIdx.x] = bits0;
            __syncthreads();
            for (int d = 1; d < GZ_THREADS; d <<= 1) {
                const u64 v = threadIdx.x >= (u32)d ? s_scan[threadIdx.x - d] : 0;
                __syncthreads();
                s_scan[threadIdx.x] += v;
                __syncthreads();
            }
            off0 = s_scan[threadIdx.x] - bits0;
            tot0 = s_scan[GZ_THREADS - 1];
            __syncthreads();
            if (has_q) {
                s_scan[threadIdx.x] = bits1;
                __syncthreads();
                for (int d = 1; d < GZ_THREADS; d <<= 1) {
                    const u64 v = threadIdx.x >= (u32)d ? s_scan[threadIdx.x - d] : 0;
                    __syncthreads();
                    s_scan[threadIdx.x] += v;
                    __syncthreads();
                }
                off1 = s_scan[threadIdx.x] - bits1;
                tot1 = s_scan[GZ_THREADS - 1];
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            u64 bit = s_hdr_bits[0];
            s_base_bit[0] = bit;                                     // block 0 symbols
            bit += tot0 + (has_q ? 0 : s_len[0]['\n']) + s_len[0][256];
            s_base_bit[1] = bit;                                     // block 1 header (FASTQ)
            if (has_q) bit += s_hdr_bits[1] + 2 * s_len[1]['\n'] + s_len[1]['+'] + tot1 + s_len[1]['\n'] + s_len[1][256];
            s_base_bit[2] = bit;
            const u64 bytes = (bit + 7) >> 3;
            const u64 padded = (bytes + 3) & ~3ull; // members start on 32-bit words of the blob
            const u64 at = atomicAdd(cursor, (unsigned long long)padded);
            if (at + padded > blob_cap_bytes) { atomicExch(overflow, 1u); s_member_off = ~0ull; spans[pi] = GzSpan{0, 0, 0}; }
            else { s_member_off = at; spans[pi] = GzSpan{at, (u32)bytes, 0}; }
        }
        __syncthreads();
        if (s_member_off == ~0ull) { __syncthreads(); continue; }
        u32 *words = blob_words + (s_member_off >> 2);
        // ---- 4. write ----
        GzBitOut o;
        if (i0 < i1) {
            gz_out_begin(o, words, s_base_bit[0] + off0);
            for (u32 i = i0; i < i1; ++i) { const uint8_t c = seq[i]; gz_out_put(o, s_code[0][c], s_len[0][c]); }
            gz_out_end(o);
            if (has_q) {
                const u64 q_sym0 = s_base_bit[1] + s_hdr_bits[1] + 2 * s_len[1]['\n'] + s_len[1]['+'];
                gz_out_begin(o, words, q_sym0 + off1);
                for (u32 i = i0; i < i1; ++i) { const uint8_t c = qual[i]; gz_out_put(o, s_code[1][c], s_len[1][c]); }
                gz_out_end(o);
            }
        }
        if (threadIdx.x == 0) { // headers, literal extras and end-of-block codes
            gz_out_begin(o, words, 0);
            for (u32 b = 0; b < s_hdr_bits[0]; b += 8) gz_out_put(o, s_hdr[0][b >> 3], min(8u, s_hdr_bits[0] - b));
            gz_out_end(o);
            gz_out_begin(o, words, s_base_bit[0] + tot0);
            if (!has_q) gz_out_put(o, s_code[0]['\n'], s_len[0]['\n']);
            gz_out_put(o, s_code[0][256], s_len[0][256]);
            gz_out_end(o);
            if (has_q) {
                gz_out_begin(o, words, s_base_bit[1]);
                for (u32 b = 0; b < s_hdr_bits[1]; b += 8) gz_out_put(o, s_hdr[1][b >> 3], min(8u, s_hdr_bits[1] - b));
                gz_out_put(o, s_code[1]['\n'], s_len[1]['\n']);
                gz_out_put(o, s_code[1]['+'], s_len[1]['+']);
                gz_out_put(o, s_code[1]['\n'], s_len[1]['\n']);
                gz_out_end(o);
                gz_out_begin(o, words, s_base_bit[2] - s_len[1]['\n'] - s_len[1][256]);
                gz_out_put(o, s_code[1]['\n'], s_len[1]['\n']);
                gz_out_put(o, s_code[1][256], s_len[1][256]);
                gz_out_end(o);
            }
        }
        __syncthreads(); // shared tables are reused by the next piece
    }
}
