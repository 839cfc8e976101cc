// K4, k <= 16, pieces of up to KMER16_MAX_KMERS k-mers: distinct k-mer count without shared-memory
// atomics on the k-mer path.  Replaces GetKmerCount (T.cpp:1703-1753) for the common case (the default
// k = 11 on HiFi / CLR / ordinary ONT pieces); longer pieces are handed to k_kmer_smem (k <= 12) or k_kmer (hash)
// through a device list.  Two CTAs per SM; inside a CTA two PRODUCER warps stage the next two pieces (global loads, ASCII
// -> 2-bit codes, one tile buffer each) while 14 CONSUMER warps count the current one, so neither the
// global-load latency nor the piece metadata chain is ever on the consumers' path (named barriers:
// consumers among themselves, full / empty per tile buffer).
//
// Counting a staged piece (k16_count):
//   * dense round, "owner bytes": the table is 65 536 BYTE slots (k = 13 .. 16: 32 768 u16 slots, the same scheme with
//     15 remainder bits per entry).  slot = top 16 bits of the key, and
//     every k-mer stores (low key bits | 0x80) there — no position, no atomics, last writer wins.  After
//     one barrier a k-mer that reads back its own byte belongs to the key that owns the slot; all
//     instances of a key share slot and verdict, so the keys resolved by the round are exactly the
//     non-empty slots, counted by one sweep over the table, and no lane ever branches on table contents.
//     The k-mers of every other key are appended to a pending list (warp-aggregated, one shared atomic
//     per warp and word).
//   * list rounds: the same memory as 32 768 u16 position tags, slot = multiplicative hash of the key.
//     A pending k-mer that reads back its own position is the representative of its key (count 1);
//     otherwise it fetches the winner's key from the staged stream: equal -> duplicate, different -> it
//     lost again and goes to the next list.  A k-mer only reads a slot it wrote in the same round.  The
//     last <= 32 keys are settled by one warp with __match_any_sync.
//   * pieces with more than KMER16_ONE_PASS k-mers are processed in 2^n passes over the classes given by
//     the top bits of the key REMAINDER (the bits that are not in the slot), so the table load stays
//     <= 0.4.  If a pass still produces more losers than the list holds, the piece is redone with twice
//     the passes; with one class per remainder value no slot can hold two keys, so this terminates.
// Base codes: (byte >> 1) & 3 (A0 C1 T2 G3) — any injective recoding of ACGT counts the same distinct
// k-mers as the reference's A0 C1 G2 T3; every other byte must collide with 'A' (T.cpp:1709-1724) and
// is recoded to 0 on a slow path that only runs for 16-byte groups containing such a byte.
#pragma once
#include "common.cuh"

#define KMER16_CONSUMERS 448                         // 14 warps count
#define KMER16_THREADS 512                           // + 2 warps that stage the next two pieces (one tile buffer each)
#define KMER16_TABLE_BYTES 65536u                    // 65 536 owner bytes / 32 768 u16 position tags
#define KMER16_TILE_WORDS 4096u                      // 16 bases per word
#define KMER16_MAX_KMERS 65000u                      // one staged tile; positions (alignment shift included) fit u16
#define KMER16_LIST_CAP 4096u
#define KMER16_ONE_PASS 22000u                       // table load 0.34: ~15 % of the k-mers lose their slot
#define KMER16_TILE_BYTES ((KMER16_TILE_WORDS + 2u) * 4u)
#define KMER16_SMEM_BYTES (KMER16_TABLE_BYTES + 2u * KMER16_TILE_BYTES + 2u * KMER16_LIST_CAP * 2u)
#define KMER16_END 0xFFFFFFFFu

struct K16Meta { u32 pi, shift; int total, len; };

static __device__ __forceinline__ void k16_bar_sync(u32 id, u32 n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
static __device__ __forceinline__ void k16_bar_arrive(u32 id, u32 n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
static __device__ __forceinline__ void k16_sync_consumers() { asm volatile("bar.sync 1, 448;" ::: "memory"); }

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

static __device__ __forceinline__ u32 k16_codes16(uint4 q) {
    u32 bad = 0;
    const u32 c0 = k16_codes4(q.x, bad), c1 = k16_codes4(q.y, bad), c2 = k16_codes4(q.z, bad), c3 = k16_codes4(q.w, bad);
    u32 out = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    if (bad) out = (k16_codes4_slow(q.x) << 24) | (k16_codes4_slow(q.y) << 16) | (k16_codes4_slow(q.z) << 8) | k16_codes4_slow(q.w);
    return out;
}

// Producer warp: bases [abase, abase + 16 * n_words) as 2-bit codes into tile[0 .. n_words), then two zero
// words (the window of the last k-mer reads past the stream).  Eight 16-byte loads in flight per lane.
static __device__ __forceinline__ void k16_stage(const DevBatch &B, u64 n_total, u64 abase, u32 n_words, u32 *tile, u32 lane) {
    const u32 n_full = (abase + 16ull * n_words <= n_total) ? n_words : (u32)((n_total - abase) / 16ull); // words fully inside the stream
    u32 w = lane;
    for (; w + 224u < n_full; w += 256u) {
        uint4 q[8];
#pragma unroll
        for (u32 i = 0; i < 8; ++i) q[i] = __ldg((const uint4 *)(B.bases + abase + 16ull * (w + 32u * i)));
#pragma unroll
        for (u32 i = 0; i < 8; ++i) tile[w + 32u * i] = k16_codes16(q[i]);
    }
    for (; w + 96u < n_full; w += 128u) {
        uint4 q[4];
#pragma unroll
        for (u32 i = 0; i < 4; ++i) q[i] = __ldg((const uint4 *)(B.bases + abase + 16ull * (w + 32u * i)));
#pragma unroll
        for (u32 i = 0; i < 4; ++i) tile[w + 32u * i] = k16_codes16(q[i]);
    }
    for (; w < n_words + 2u; w += 32u) {
        u32 out = 0;
        if (w < n_full) {
            out = k16_codes16(__ldg((const uint4 *)(B.bases + abase + 16ull * w)));
        } else if (w < n_words) { // last, partial group of the batch: bytes beyond the stream are not touched
            __align__(16) uint8_t tmp[16];
            const u64 a = abase + 16ull * w;
#pragma unroll
            for (int j = 0; j < 16; ++j) tmp[j] = (a + j < n_total) ? B.bases[a + j] : (uint8_t)'A';
            out = k16_codes16(*(uint4 *)tmp);
        }
        tile[w] = out;
    }
}

static __device__ __forceinline__ u32 k16_kmer_at(const u32 *tile, u32 pos) { // 32 stream bits from position pos on
    const u32 w = pos >> 4;
    return __funnelshift_l(tile[w + 1], tile[w], 2u * (pos & 15u));
}

// Warp-aggregated append of the positions (pw | j) for the set bits j of `m` to list[*cnt ...] (a warp
// whose entries do not all fit drops them; the caller sees the overflow in *cnt).  All 32 lanes must call.
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
    if (base + wtot <= list_cap) {
        while (m) {
            const u32 j = (u32)__ffs((int)m) - 1u;
            m &= m - 1u;
            list[off++] = (uint16_t)(pw + j);
        }
    }
}

// One attempt at a staged piece with 2^pass_bits key classes (consumer threads only).  Returns false (in
// every thread) if a pass produced more pending k-mers than the list holds; the caller retries with more
// passes.  KT: compile-time k (0: use krt).
// ENT: owner entry: uint8_t (k <= 12: 65 536 slots, 7 remainder bits) or uint16_t (k = 13 .. 16: 32 768 slots, 15 bits).
template <int KT, typename ENT>
static __device__ bool k16_count(const u32 *tile, uint8_t *table8, uint16_t *list_a, uint16_t *list_b, u32 *s_cnt,
                                 u32 shift, u32 p_end, int krt, u32 pass_bits, u32 list_cap, u32 &mine) {
    const int k = KT ? KT : krt;
    uint16_t *table = (uint16_t *)table8;
    ENT *own = (ENT *)table8;
    constexpr u32 SB = sizeof(ENT) == 1 ? 16u : 15u;             // slot bits
    constexpr u32 FLAG = sizeof(ENT) == 1 ? 0x80u : 0x8000u;     // set in every stored entry: 0 = empty
    constexpr u32 EMASK = sizeof(ENT) == 1 ? 0xFFu : 0xFFFFu;
    const u32 sh = 32u - 2u * (u32)k;            // x >> sh = key
    const u32 slot_sh = sh > 32u - SB ? sh : 32u - SB; // slot = top SB bits of the key (the whole key when 2k <= SB)
    const u32 n_scan = (p_end + 15u) >> 4;       // words holding the start of at least one k-mer
    const u32 lane = threadIdx.x & 31u;
    const u32 n_pass = 1u << pass_bits;
    const u32 cls_sh = 32u - SB - pass_bits;     // class = top pass_bits bits of the remainder x[31 - SB : sh]
    mine = 0;
    for (u32 pass = 0; pass < n_pass; ++pass) {
        k16_sync_consumers(); // previous pass / piece is done with the table, the lists and s_cnt
        if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
        for (u32 i = threadIdx.x; i < KMER16_TABLE_BYTES / 16u; i += KMER16_CONSUMERS) ((uint4 *)table8)[i] = make_uint4(0, 0, 0, 0);
        k16_sync_consumers();
        // ---- dense round, store ----
        for (u32 w = threadIdx.x; w < n_scan; w += KMER16_CONSUMERS) {
            const u32 w0 = tile[w], w1 = tile[w + 1], pw = w << 4;
            if (pass_bits == 0 && pw >= shift && pw + 16u <= p_end) {
#pragma unroll
                for (u32 j = 0; j < 16; ++j) {
                    const u32 x = __funnelshift_l(w1, w0, 2 * j);
                    own[x >> slot_sh] = (ENT)((x >> sh) | FLAG);
                }
            } else if (pw >= shift && pw + 16u <= p_end) { // interior word of a multi-pass piece: unrolled, class-predicated
#pragma unroll
                for (u32 j = 0; j < 16; ++j) {
                    const u32 x = __funnelshift_l(w1, w0, 2 * j);
                    if (((x >> cls_sh) & (n_pass - 1u)) == pass) own[x >> slot_sh] = (ENT)((x >> sh) | FLAG);
                }
            } else {
                u32 m = 0xFFFFu;
                if (pw < shift) m &= 0xFFFFu << (shift - pw);
                if (pw + 16u > p_end) m &= 0xFFFFu >> (pw + 16u - p_end);
#pragma unroll 4
                for (u32 j = 0; j < 16; ++j) {
                    const u32 x = __funnelshift_l(w1, w0, 2 * j);
                    if (((m >> j) & 1u) && (pass_bits == 0 || ((x >> cls_sh) & (n_pass - 1u)) == pass))
                        own[x >> slot_sh] = (ENT)((x >> sh) | FLAG);
                }
            }
        }
        k16_sync_consumers();
        // ---- dense round, read back; k-mers whose key does not own its slot go to list_a ----
        for (u32 w_base = threadIdx.x & ~31u; w_base < n_scan; w_base += KMER16_CONSUMERS) { // warp-uniform trip count
            const u32 w = w_base + lane;
            u32 pend = 0, pw = w << 4;
            if (w < n_scan) {
                const u32 w0 = tile[w], w1 = tile[w + 1];
                if (pass_bits == 0 && pw >= shift && pw + 16u <= p_end) {
#pragma unroll
                    for (u32 j = 0; j < 16; ++j) {
                        const u32 x = __funnelshift_l(w1, w0, 2 * j);
                        const u32 v = own[x >> slot_sh];
                        pend |= (v != (((x >> sh) | FLAG) & EMASK)) ? (1u << j) : 0u;
                    }
                } else if (pw >= shift && pw + 16u <= p_end) { // interior word of a multi-pass piece
#pragma unroll
                    for (u32 j = 0; j < 16; ++j) {
                        const u32 x = __funnelshift_l(w1, w0, 2 * j);
                        const u32 v = own[x >> slot_sh];
                        pend |= (((x >> cls_sh) & (n_pass - 1u)) == pass && v != (((x >> sh) | FLAG) & EMASK)) ? (1u << j) : 0u;
                    }
                } else {
                    u32 m = 0xFFFFu;
                    if (pw < shift) m &= 0xFFFFu << (shift - pw);
                    if (pw + 16u > p_end) m &= 0xFFFFu >> (pw + 16u - p_end);
#pragma unroll 4
                    for (u32 j = 0; j < 16; ++j) {
                        const u32 x = __funnelshift_l(w1, w0, 2 * j);
                        if (((m >> j) & 1u) && (pass_bits == 0 || ((x >> cls_sh) & (n_pass - 1u)) == pass)) {
                            const u32 v = own[x >> slot_sh];
                            pend |= (v != (((x >> sh) | FLAG) & EMASK)) ? (1u << j) : 0u;
                        }
                    }
                }
            }
            k16_append(pend, pw, list_a, &s_cnt[0], lane, list_cap);
        }
        // ---- the keys resolved by this round = the non-empty byte slots (every stored byte has bit 7 set) ----
        for (u32 i = threadIdx.x; i < KMER16_TABLE_BYTES / 16u; i += KMER16_CONSUMERS) {
            const uint4 q = ((const uint4 *)table8)[i];
            constexpr u32 FM = sizeof(ENT) == 1 ? 0x80808080u : 0x80008000u; // the flag bit of every entry of a word
            mine += __popc(((q.x & FM) >> 3) | ((q.y & FM) >> 2) | ((q.z & FM) >> 1) | (q.w & FM));
        }
        k16_sync_consumers();
        u32 n_list = s_cnt[0];
        if (n_list > list_cap) return false;
        // ---- list rounds: u16 position tags, winner's key checked against the staged stream ----
        uint16_t *cur = list_a, *nxt = list_b;
        u32 ci = 1, round = 0;
        while (n_list > 32u) {
            ++round;
            const u32 mult = 0x9E3779B1u + 0x3C6EF372u * round; // odd
            for (u32 i = threadIdx.x; i < n_list; i += KMER16_CONSUMERS) {
                const u32 pos = cur[i];
                const u32 key = k16_kmer_at(tile, pos) >> sh;
                table[(key * mult) >> 17] = (uint16_t)pos;
            }
            k16_sync_consumers();
            for (u32 i0 = threadIdx.x & ~31u; i0 < n_list; i0 += KMER16_CONSUMERS) { // warp-uniform trip count
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
            k16_sync_consumers();
            n_list = s_cnt[ci];
            ci ^= 1u;
            if (threadIdx.x == 0) s_cnt[ci] = 0; // next round's counter; nobody reads it before the next barrier
            uint16_t *t = cur; cur = nxt; nxt = t;
        }
        if (n_list && threadIdx.x < 32u) { // the last few keys: one warp, one MATCH among the lanes that hold one
            const bool have = lane < n_list;
            const u32 vmask = __ballot_sync(0xffffffffu, have);
            if (have) {
                const u32 key = k16_kmer_at(tile, cur[lane]) >> sh;
                const u32 same = __match_any_sync(vmask, key);
                if ((u32)(__ffs((int)same) - 1) == lane) ++mine;
            }
        }
    }
    return true;
}

static __device__ __forceinline__ void k16_finish_piece(tgsf_piece *pieces, u32 pi, int len, int repeat, const DevParams &P,
                                                        u64 *counters) {
    pieces[pi].repeat_len = repeat;
    if (repeat < P.min_repeat) { // T.cpp:1984-1988
        pieces[pi].status = TGSF_PIECE_SHORT_REPEAT;
        atomic_add_u64(counters + P.L.drop_info + 15, 1ull);
        atomic_add_u64(counters + P.L.drop_info + 16, (u64)len);
    }
}

__global__ void __launch_bounds__(KMER16_THREADS, 2)
k_kmer_tag16(DevBatch B, DevParams P, tgsf_piece *pieces, const u32 *__restrict__ n_pieces_ptr,
             u64 *__restrict__ counters, const u32 *__restrict__ dev_status, u32 *__restrict__ work_ctr,
             u32 *__restrict__ long_list, u32 *__restrict__ n_long, u32 list_cap) { // list_cap <= KMER16_LIST_CAP (smaller: tests)
    if (*dev_status != DEV_STATUS_OK) return;
    extern __shared__ __align__(16) uint8_t k16mem[];
    uint8_t *table = k16mem;
    u32 *tiles = (u32 *)(k16mem + KMER16_TABLE_BYTES);
    uint16_t *list_a = (uint16_t *)(k16mem + KMER16_TABLE_BYTES + 2u * KMER16_TILE_BYTES);
    uint16_t *list_b = list_a + KMER16_LIST_CAP;
    __shared__ u32 s_cnt[2];
    __shared__ u32 s_distinct;
    __shared__ K16Meta s_meta[2];
    const int k = P.kmer;
    const u32 n_pieces = *n_pieces_ptr;
    // barriers: 1 = consumers; 2 + b = tile b full (its producer arrives, consumers wait); 4 + b = tile b empty

    if (threadIdx.x >= KMER16_CONSUMERS) { // ---------------- producer warps: warp p fills tile p ----------------
        const u32 lane = threadIdx.x & 31u;
        const u32 b = (threadIdx.x - KMER16_CONSUMERS) >> 5;
        const u64 n_total = B.offsets[B.n_reads];
        for (u32 n = 0;; ++n) {
            u32 pi;
            tgsf_piece pc;
            int total = 0;
            for (;;) { // next piece this kernel counts
                pi = 0;
                if (lane == 0) pi = atomicAdd(work_ctr, 1u);
                pi = __shfl_sync(0xffffffffu, pi, 0);
                if (pi >= n_pieces) break;
                pc = pieces[pi];
                if (pc.status != TGSF_PIECE_EMIT) continue;
                total = pc.len - k + 1;
                if (total > (int)KMER16_MAX_KMERS) { // too long for 16-bit positions: k_kmer_smem takes it
                    if (lane == 0) long_list[atomicAdd(n_long, 1u)] = pi;
                    continue;
                }
                if (total <= 0) { // no k-mer at all: see oracle/tgsf_oracle.c kmer_repeat_len
                    if (lane == 0) k16_finish_piece(pieces, pi, pc.len, total - 1, P, counters);
                    continue;
                }
                break;
            }
            // (the metadata chain above ran while the consumers were still counting this tile's previous piece)
            if (n >= 1u) k16_bar_sync(4u + b, KMER16_CONSUMERS + 32u); // the consumers are done with tile b
            if (pi >= n_pieces) {
                if (lane == 0) s_meta[b].pi = KMER16_END;
                __threadfence_block();
                __syncwarp();
                k16_bar_arrive(2u + b, KMER16_CONSUMERS + 32u);
                return;
            }
            const u64 seq0 = B.offsets[pc.read] + (u64)pc.start; // absolute offset of the piece
            const u64 abase = seq0 & ~15ull;
            const u32 shift = (u32)(seq0 - abase);
            k16_stage(B, n_total, abase, (shift + (u32)total + (u32)k - 1u + 15u) / 16u, tiles + b * (KMER16_TILE_WORDS + 2u), lane);
            if (lane == 0) { s_meta[b].pi = pi; s_meta[b].shift = shift; s_meta[b].total = total; s_meta[b].len = pc.len; }
            __threadfence_block();
            __syncwarp();
            k16_bar_arrive(2u + b, KMER16_CONSUMERS + 32u);
        }
    }

    // ---------------- consumer warps: tiles in turn until both producers have signalled the end ----------------
    u32 ended = 0;
    for (u32 n = 0;; ++n) {
        const u32 b = n & 1u;
        if (ended & (1u << b)) continue; // (the other tile is still live, or the loop would have returned)
        k16_bar_sync(2u + b, KMER16_CONSUMERS + 32u); // tile b and its meta are ready
        const K16Meta M = s_meta[b];
        if (M.pi == KMER16_END) {
            ended |= 1u << b;
            if (ended == 3u) return;
            continue;
        }
        const u32 *tile = tiles + b * (KMER16_TILE_WORDS + 2u);
        if (threadIdx.x == 0) s_distinct = 0; // ordered before its use by the barriers inside k16_count
        // classes split the key REMAINDER (the 2k - 16 / 2k - 15 bits that are not in the slot; none for small k, where
        // slot = key and nothing can collide).  An owner byte holds 7 remainder bits (k = 12 needs two classes), an
        // owner u16 15 (k = 16: four); the u16 table has half the slots, so its passes start at half the piece size.
        const bool wide = k > 12;
        const u32 sb = wide ? 15u : 16u;
        const u32 rem_bits = 2u * (u32)k > sb ? 2u * (u32)k - sb : 0u;
        const u32 ent_bits = wide ? 15u : 7u;
        const u32 one_pass = wide ? KMER16_ONE_PASS / 2u : KMER16_ONE_PASS;
        u32 pass_bits = rem_bits > ent_bits ? rem_bits - ent_bits : 0u;
        while (pass_bits < rem_bits && ((u32)M.total >> pass_bits) > one_pass) ++pass_bits;
        u32 mine = 0;
        if (k == 11) {
            while (!k16_count<11, uint8_t>(tile, table, list_a, list_b, s_cnt, M.shift, M.shift + (u32)M.total, k, pass_bits, list_cap, mine)) ++pass_bits;
        } else if (!wide) {
            while (!k16_count<0, uint8_t>(tile, table, list_a, list_b, s_cnt, M.shift, M.shift + (u32)M.total, k, pass_bits, list_cap, mine)) ++pass_bits;
        } else {
            while (!k16_count<0, uint16_t>(tile, table, list_a, list_b, s_cnt, M.shift, M.shift + (u32)M.total, k, pass_bits, list_cap, mine)) ++pass_bits;
        }
        mine = warp_sum_u32(mine);
        if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(&s_distinct, mine);
        k16_sync_consumers();
        if (threadIdx.x == 0) k16_finish_piece(pieces, M.pi, M.len, M.total - (int)s_distinct, P, counters);
        k16_bar_arrive(4u + b, KMER16_CONSUMERS + 32u); // tile b may be refilled
    }
}
