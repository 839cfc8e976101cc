// K4: k-mer repeat length of a kept piece = (L - k + 1) - #distinct k-mers.
// Replaces GetKmerCount (T.cpp:1703-1753): rolling 2-bit code A0 C1 G2 T3, every other byte
// (lower case, N) contributes 00, mask (1 << 2k) - 1, std::unordered_set of the codes.
//
// One CTA per piece; the set is an open-addressing hash table in shared memory (atomicCAS).
// Pieces with more k-mers than half the table are processed in several passes over disjoint
// hash partitions of the key space, so any piece length is exact with the same 128 KB table.
#pragma once
#include "common.cuh"

#define KMER_THREADS 512
#define KMER_SMEM_BYTES (128 * 1024)

static __device__ __forceinline__ u32 mix32(u64 x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (u32)x;
}

static __device__ __forceinline__ u32 base_code(uint8_t b) { // T.cpp:1709-1724
    return b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 0u;
}

template <typename KEY>
__global__ void __launch_bounds__(KMER_THREADS, 1)
k_kmer(DevBatch B, DevParams P, tgsf_piece *pieces, const u32 *__restrict__ n_pieces_ptr,
       u64 *__restrict__ counters, const u32 *__restrict__ dev_status) {
    if (*dev_status != DEV_STATUS_OK) return;
    extern __shared__ __align__(16) uint8_t kmem[];
    KEY *table = (KEY *)kmem;
    constexpr u32 SLOTS = KMER_SMEM_BYTES / sizeof(KEY);
    constexpr KEY EMPTY = (KEY)~(KEY)0;
    __shared__ u32 s_distinct;
    const int k = P.kmer;
    const u64 mask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const u32 n_pieces = *n_pieces_ptr;

    for (u32 pi = blockIdx.x; pi < n_pieces; pi += gridDim.x) {
        tgsf_piece pc = pieces[pi];
        if (pc.status != TGSF_PIECE_EMIT) continue;
        const int L = pc.len;
        const int total = L - k + 1;
        int repeat;
        if (total <= 0) {
            repeat = total - 1; // see oracle/tgsf_oracle.c kmer_repeat_len
        } else {
            const uint8_t *seq = B.bases + B.offsets[pc.read] + (u64)pc.start;
            const u32 passes = ((u32)total + SLOTS / 2 - 1) / (SLOTS / 2);
            if (threadIdx.x == 0) s_distinct = 0;
            const int per = (total + KMER_THREADS - 1) / KMER_THREADS;
            const int i0 = min((int)threadIdx.x * per, total), i1 = min(i0 + per, total);
            u32 mine = 0;
            for (u32 pass = 0; pass < passes; ++pass) {
                __syncthreads();
                for (u32 i = threadIdx.x; i < SLOTS; i += KMER_THREADS) table[i] = EMPTY;
                __syncthreads();
                if (i0 < i1) {
                    u64 km = 0;
                    for (int j = 0; j < k - 1; ++j) km = (km << 2) | base_code(seq[i0 + j]);
                    for (int i = i0; i < i1; ++i) {
                        km = ((km << 2) | base_code(seq[i + k - 1])) & mask;
                        const u32 h = mix32(km);
                        if (passes > 1 && (u32)(((u64)(h >> 8) * passes) >> 24) != pass) continue;
                        u32 slot = h & (SLOTS - 1);
                        const KEY key = (KEY)km;
                        while (true) {
                            const KEY old = atomicCAS(&table[slot], EMPTY, key);
                            if (old == EMPTY) { ++mine; break; }
                            if (old == key) break;
                            slot = (slot + 1) & (SLOTS - 1);
                        }
                    }
                }
            }
            atomicAdd(&s_distinct, mine);
            __syncthreads();
            repeat = total - (int)s_distinct;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
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

// k <= 13: direct-addressed bitmap of the 4^k key space in global memory, one per CTA (512 KB for
// k = 11, so the bitmaps of all 148 CTAs stay L2-resident).  One atomicOr per k-mer (its return value
// tells whether the k-mer is new), then the touched words are zeroed again with plain stores, so
// the bitmap is clean for the CTA's next piece without a 512 KB memset per read.  Replaces the
// shared-memory hash for small k: no probing, no passes, and L2 atomics issue ~1.5x faster per SM
// than shared-memory CAS.
#define KMER_BM_THREADS 512
__global__ void __launch_bounds__(KMER_BM_THREADS)
k_kmer_bitmap(DevBatch B, DevParams P, tgsf_piece *pieces, const u32 *__restrict__ n_pieces_ptr,
              u32 *__restrict__ bitmaps, u64 words_per_cta, u64 *__restrict__ counters,
              const u32 *__restrict__ dev_status) {
    if (*dev_status != DEV_STATUS_OK) return;
    __shared__ u32 s_distinct;
    u32 *bm = bitmaps + (u64)blockIdx.x * words_per_cta;
    const int k = P.kmer;
    const u32 mask = (u32)((1ull << (2 * k)) - 1ull);
    const u32 n_pieces = *n_pieces_ptr;
    for (u32 pi = blockIdx.x; pi < n_pieces; pi += gridDim.x) {
        tgsf_piece pc = pieces[pi];
        if (pc.status != TGSF_PIECE_EMIT) continue;
        const int L = pc.len;
        const int total = L - k + 1;
        int repeat;
        if (total <= 0) {
            repeat = total - 1;
        } else {
            const uint8_t *seq = B.bases + B.offsets[pc.read] + (u64)pc.start;
            if (threadIdx.x == 0) s_distinct = 0;
            __syncthreads();
            const int per = (total + KMER_BM_THREADS - 1) / KMER_BM_THREADS;
            const int i0 = min((int)threadIdx.x * per, total), i1 = min(i0 + per, total);
            u32 mine = 0;
            if (i0 < i1) {
                u32 km = 0;
                for (int j = 0; j < k - 1; ++j) km = (km << 2) | base_code(seq[i0 + j]);
                for (int i = i0; i < i1; ++i) {
                    km = ((km << 2) | base_code(seq[i + k - 1])) & mask;
                    const u32 bit = 1u << (km & 31u);
                    const u32 old = atomicOr(bm + (km >> 5), bit);
                    mine += (old & bit) ? 0u : 1u;
                }
            }
            atomicAdd(&s_distinct, mine);
            __syncthreads();
            repeat = total - (int)s_distinct;
            if (i0 < i1) { // wipe exactly the words this piece touched
                u32 km = 0;
                for (int j = 0; j < k - 1; ++j) km = (km << 2) | base_code(seq[i0 + j]);
                for (int i = i0; i < i1; ++i) {
                    km = ((km << 2) | base_code(seq[i + k - 1])) & mask;
                    bm[km >> 5] = 0u;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
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
