// K4, k <= 12, pieces of up to KMER16_MAX_KMERS k-mers: distinct k-mer count with 16-bit "last writer
// wins" position tags.  Replaces GetKmerCount (T.cpp:1703-1753) for the common case (the default k = 11
// on HiFi / CLR / ordinary ONT pieces); longer pieces are handed to k_kmer_smem through a device list.
//
// Same idea as the 32-bit tag rounds of k_kmer_smem (no shared-memory atomics on the k-mer path: every
// pending k-mer stores a tag into the slot its key hashes to, and after one barrier reads the slot
// back), re-cut so that TWO CTAs fit one SM and overlap each other's barriers and global-load latency:
//   * the tag is the k-mer's POSITION only (u16), so 32 768 slots cost 64 KB instead of 128 KB.  A
//     k-mer that reads back its own position is the one representative of its key (count 1); otherwise
//     it fetches the winner's key from the staged 2-bit stream and compares: equal -> duplicate of the
//     representative, different -> its key lost the slot and stays pending.  All instances of a key
//     share slot and winner, so keys are resolved as a whole, and a k-mer only reads a slot it wrote in
//     the same round, so the table is never cleared.
//   * one dense round per pass (slot = low 15 bits of the key), the losers go straight to a position
//     list (warp-aggregated append, one shared atomic per warp) and are resolved by list rounds with a
//     multiplicative hash; the last <= 32 keys are settled by one warp with __match_any_sync.
//   * pieces with more than KMER16_ONE_PASS k-mers are processed in 2^n passes over the classes given by
//     the TOP bits of the key, so the table load stays <= ~0.6.  If a pass still produces more losers
//     than the list holds, the piece is redone with twice the passes; with one class per remainder value
//     no slot can hold two keys, so the retry loop always terminates.
// Base codes: (byte >> 1) & 3 (A0 C1 T2 G3) — any injective recoding of ACGT counts the same distinct
// k-mers as the reference's A0 C1 G2 T3; every other byte must collide with 'A' (T.cpp:1709-1724) and
// is recoded to 0 on a slow path that only runs for 16-byte groups containing such a byte.
#pragma once
#include "common.cuh"

#define KMER16_THREADS 512
#define KMER16_SLOTS 32768u
#define KMER16_TILE_WORDS 4096u                      // 16 bases per word
#define KMER16_MAX_KMERS 65000u                      // positions (alignment shift included) fit u16
#define KMER16_LIST_CAP 6144u
#define KMER16_ONE_PASS 20000u
#define KMER16_SMEM_BYTES (KMER16_SLOTS * 2u + (KMER16_TILE_WORDS + 2u) * 4u + 2u * KMER16_LIST_CAP * 2u)

// 4 ASCII bases (lowest address in the low byte) -> 8 bits of codes, first base in the top 2 bits;
// bad |= non-zero iff a byte is not one of A C G T.
static __device__ __forceinline__ u32 k16_codes4(u32 v, u32 &bad) {
    const u32 c = (v >> 1) & 0x03030303u;            // A0 C1 T2 G3
    const u32 s = c | (c >> 4);                      // byte0 = c0 | c1 << 4, byte2 = c2 | c3 << 4
    const u32 sel = __byte_perm(s, 0u, 0x4420u);     // nibbles c0 c1 c2 c3
    const u32 expect = __byte_perm(0x47544341u, 0u, sel); // the letters those codes stand for
    bad |= v ^ expect;
    return (c * 0x40100401u) >> 24;
}

static __device__ __forceinline__ u32 k16_codes4_slow(u32 v) {
    u32 out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u32 b = (v >> (8 * i)) & 0xffu;
        const u32 code = (b == 'C') ? 1u : (b == 'T') ? 2u : (b == 'G') ? 3u : 0u;
        out |= code << (6 - 2 * i);
    }
    return out;
}

static __device__ __forceinline__ void k16_stage(const DevBatch &B, u64 n_total, u64 abase, u32 n_words, u32 *tile) {
    for (u32 w = threadIdx.x; w < n_words + 2; w += KMER16_THREADS) {
        u32 out = 0;
        if (w < n_words) {
            const u64 a = abase + 16ull * w;
            uint4 q;
            if (a + 16 <= n_total) {
                q = __ldg((const uint4 *)(B.bases + a));
            } else { // last, partial group of the batch: bytes beyond the stream are not touched
                __align__(16) uint8_t tmp[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) tmp[j] = (a + j < n_total) ? B.bases[a + j] : (uint8_t)'A';
                q = *(uint4 *)tmp;
            }
            u32 bad = 0;
            const u32 c0 = k16_codes4(q.x, bad), c1 = k16_codes4(q.y, bad), c2 = k16_codes4(q.z, bad), c3 = k16_codes4(q.w, bad);
            out = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
            if (bad)
                out = (k16_codes4_slow(q.x) << 24) | (k16_codes4_slow(q.y) << 16) | (k16_codes4_slow(q.z) << 8) | k16_codes4_slow(q.w);
        }
        tile[w] = out; // two zero words behind the stream: the window of the last k-mer reads past it
    }
}

static __device__ __forceinline__ u32 k16_kmer_at(const u32 *tile, u32 pos) { // 32 stream bits from position pos on
    const u32 w = pos >> 4;
    return __funnelshift_l(tile[w + 1], tile[w], 2u * (pos & 15u));
}

// Warp-aggregated append of the positions (pw | j) for the set bits j of `m` to list[*cnt ...] (entries
// beyond the capacity are dropped; the caller sees the overflow in *cnt).  All 32 lanes must call.
static __device__ __forceinline__ void k16_append(u32 m, u32 pw, uint16_t *list, u32 *cnt, u32 lane, u32 list_cap) {
    const u32 c = __popc(m);
    u32 incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    const u32 wtot = __shfl_sync(0xffffffffu, incl, 31);
    if (wtot == 0) return;
    u32 base = 0;
    if (lane == 31) base = atomicAdd(cnt, wtot);
    base = __shfl_sync(0xffffffffu, base, 31);
    u32 off = base + incl - c;
    while (m) {
        const u32 j = (u32)__ffs((int)m) - 1u;
        m &= m - 1u;
        if (off < list_cap) list[off] = (uint16_t)(pw + j);
        ++off;
    }
}

// One attempt at a staged piece with 2^pass_bits key classes.  Returns false (in every thread) if a pass
// produced more pending k-mers than the list holds; `mine` then holds garbage and the caller retries.
static __device__ bool k16_count(const u32 *tile, uint16_t *table, uint16_t *list_a, uint16_t *list_b, u32 *s_cnt,
                                 u32 shift, u32 p_end, int k, u32 pass_bits, u32 list_cap, u32 &mine) {
    const u32 sh = 32u - 2u * (u32)k;      // x >> sh = key
    const u32 n_scan = (p_end + 15u) >> 4; // words holding the start of at least one k-mer
    const u32 lane = threadIdx.x & 31u;
    const u32 n_pass = 1u << pass_bits;
    const u32 cls_sh = 32u - pass_bits;    // class = top pass_bits bits of the key (pass_bits <= 2k - 15 when it matters)
    mine = 0;
    for (u32 pass = 0; pass < n_pass; ++pass) {
        __syncthreads(); // previous pass / piece is done with the table, the lists and s_cnt
        if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
        // ---- dense round, store: slot = low 15 bits of the key ----
        for (u32 w = threadIdx.x; w < n_scan; w += KMER16_THREADS) {
            const u32 w0 = tile[w], w1 = tile[w + 1], pw = w << 4;
            if (pass_bits == 0 && pw >= shift && pw + 16u <= p_end) {
#pragma unroll
                for (u32 j = 0; j < 16; ++j) {
                    const u32 x = __funnelshift_l(w1, w0, 2 * j);
                    table[(x >> sh) & (KMER16_SLOTS - 1u)] = (uint16_t)(pw + j);
                }
            } else {
                u32 m = 0xFFFFu;
                if (pw < shift) m &= 0xFFFFu << (shift - pw);
                if (pw + 16u > p_end) m &= 0xFFFFu >> (pw + 16u - p_end);
#pragma unroll 4
                for (u32 j = 0; j < 16; ++j) {
                    const u32 x = __funnelshift_l(w1, w0, 2 * j);
                    if (((m >> j) & 1u) && (pass_bits == 0 || (x >> cls_sh) == pass))
                        table[(x >> sh) & (KMER16_SLOTS - 1u)] = (uint16_t)(pw + j);
                }
            }
        }
        __syncthreads();
        // ---- dense round, read back; losers are appended to list_a ----
        for (u32 w_base = threadIdx.x & ~31u; w_base < n_scan; w_base += KMER16_THREADS) { // warp-uniform trip count
            const u32 w = w_base + lane;
            u32 pend = 0, pw = w << 4;
            if (w < n_scan) {
                const u32 w0 = tile[w], w1 = tile[w + 1];
                if (pass_bits == 0 && pw >= shift && pw + 16u <= p_end) {
#pragma unroll
                    for (u32 j = 0; j < 16; ++j) {
                        const u32 x = __funnelshift_l(w1, w0, 2 * j);
                        const u32 v = table[(x >> sh) & (KMER16_SLOTS - 1u)];
                        if (v == pw + j) ++mine;
                        else if ((x ^ k16_kmer_at(tile, v)) >> sh) pend |= 1u << j;
                    }
                } else {
                    u32 m = 0xFFFFu;
                    if (pw < shift) m &= 0xFFFFu << (shift - pw);
                    if (pw + 16u > p_end) m &= 0xFFFFu >> (pw + 16u - p_end);
#pragma unroll 4
                    for (u32 j = 0; j < 16; ++j) {
                        const u32 x = __funnelshift_l(w1, w0, 2 * j);
                        if (((m >> j) & 1u) && (pass_bits == 0 || (x >> cls_sh) == pass)) {
                            const u32 v = table[(x >> sh) & (KMER16_SLOTS - 1u)];
                            if (v == pw + j) ++mine;
                            else if ((x ^ k16_kmer_at(tile, v)) >> sh) pend |= 1u << j;
                        }
                    }
                }
            }
            k16_append(pend, pw, list_a, &s_cnt[0], lane, list_cap);
        }
        __syncthreads();
        u32 n_list = s_cnt[0];
        if (n_list > list_cap) return false;
        // ---- list rounds ----
        uint16_t *cur = list_a, *nxt = list_b;
        u32 ci = 1, round = 0;
        while (n_list > 32u) {
            ++round;
            const u32 mult = 0x9E3779B1u + 0x3C6EF372u * round; // odd
            for (u32 i = threadIdx.x; i < n_list; i += KMER16_THREADS) {
                const u32 pos = cur[i];
                const u32 key = k16_kmer_at(tile, pos) >> sh;
                table[(key * mult) >> 17] = (uint16_t)pos;
            }
            __syncthreads();
            for (u32 i0 = threadIdx.x & ~31u; i0 < n_list; i0 += KMER16_THREADS) { // warp-uniform trip count
                const u32 i = i0 + lane;
                bool lost = false;
                u32 pos = 0;
                if (i < n_list) {
                    pos = cur[i];
                    const u32 x = k16_kmer_at(tile, pos);
                    const u32 v = table[((x >> sh) * mult) >> 17];
                    if (v == pos) ++mine;
                    else lost = ((x ^ k16_kmer_at(tile, v)) >> sh) != 0;
                }
                const u32 bal = __ballot_sync(0xffffffffu, lost);
                if (bal) {
                    u32 base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt[ci], (u32)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (lost) nxt[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)pos; // <= n_list entries: fits
                }
            }
            __syncthreads();
            n_list = s_cnt[ci];
            ci ^= 1u;
            if (threadIdx.x == 0) s_cnt[ci] = 0; // next round's counter; nobody reads it before the next barrier
            uint16_t *t = cur; cur = nxt; nxt = t;
        }
        if (n_list && threadIdx.x < 32u) { // the last few keys: one warp, one MATCH
            const bool have = lane < n_list;
            const u32 key = have ? (k16_kmer_at(tile, cur[lane]) >> sh) : (0xFFFFFFFFu - lane); // fillers are all distinct
            const u32 same = __match_any_sync(0xffffffffu, key);
            if (have && (u32)(__ffs((int)same) - 1) == lane) ++mine;
        }
    }
    return true;
}

__global__ void __launch_bounds__(KMER16_THREADS, 2)
k_kmer_tag16(DevBatch B, DevParams P, tgsf_piece *pieces, const u32 *__restrict__ n_pieces_ptr,
             u64 *__restrict__ counters, const u32 *__restrict__ dev_status, u32 *__restrict__ work_ctr,
             u32 *__restrict__ long_list, u32 *__restrict__ n_long, u32 list_cap) { // list_cap <= KMER16_LIST_CAP (smaller: tests)
    if (*dev_status != DEV_STATUS_OK) return;
    extern __shared__ __align__(16) uint8_t k16mem[];
    uint16_t *table = (uint16_t *)k16mem;
    u32 *tile = (u32 *)(k16mem + KMER16_SLOTS * 2u);
    uint16_t *list_a = (uint16_t *)(tile + KMER16_TILE_WORDS + 2u);
    uint16_t *list_b = list_a + KMER16_LIST_CAP;
    __shared__ u32 s_cnt[2];
    __shared__ u32 s_distinct, s_pi;
    const int k = P.kmer;
    const u32 n_pieces = *n_pieces_ptr;
    const u64 n_total = B.offsets[B.n_reads];

    for (;;) {
        __syncthreads(); // everyone is done with s_pi / s_distinct / the tile of the previous piece
        if (threadIdx.x == 0) { s_pi = atomicAdd(work_ctr, 1u); s_distinct = 0; }
        __syncthreads();
        const u32 pi = s_pi;
        if (pi >= n_pieces) return;
        tgsf_piece pc = pieces[pi];
        if (pc.status != TGSF_PIECE_EMIT) continue;
        const int total = pc.len - k + 1;
        if (total > (int)KMER16_MAX_KMERS) { // too long for 16-bit positions: k_kmer_smem takes it
            if (threadIdx.x == 0) long_list[atomicAdd(n_long, 1u)] = pi;
            continue;
        }
        if (total > 0) {
            const u64 seq0 = B.offsets[pc.read] + (u64)pc.start; // absolute offset of the piece
            const u64 abase = seq0 & ~15ull;
            const u32 shift = (u32)(seq0 - abase);
            k16_stage(B, n_total, abase, (shift + (u32)total + (u32)k - 1u + 15u) / 16u, tile);
            u32 pass_bits = 0;
            while (((u32)total >> pass_bits) > KMER16_ONE_PASS) ++pass_bits;
            u32 mine = 0;
            // (the barrier at the top of k16_count orders the staging before the first table round)
            while (!k16_count(tile, table, list_a, list_b, s_cnt, shift, shift + (u32)total, k, pass_bits, list_cap, mine)) ++pass_bits;
            mine = warp_sum_u32(mine);
            if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(&s_distinct, mine);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int repeat = total > 0 ? total - (int)s_distinct : total - 1; // see oracle/tgsf_oracle.c kmer_repeat_len
            pc.repeat_len = repeat;
            if (repeat < P.min_repeat) { // T.cpp:1984-1988
                pc.status = TGSF_PIECE_SHORT_REPEAT;
                atomic_add_u64(counters + P.L.drop_info + 15, 1ull);
                atomic_add_u64(counters + P.L.drop_info + 16, (u64)pc.len);
            }
            pieces[pi] = pc;
        }
    }
}
