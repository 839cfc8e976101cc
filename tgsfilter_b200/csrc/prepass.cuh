// K2 + pre-pass adapter search + the stand-alone alignment entry point.
//
// k_base_content : counting loop of CheckBaseContent (T.cpp:1080-1095)
// k_lib_search   : edlib loop of adapterSearch (T.cpp:1156-1176): maps[i] += mlen
// k_align_pairs  : edlibAlign(HW, PATH) for independent (query, target) pairs (tgsf_align_hw)
#pragma once
#include "adapters.cuh"
#include "scan.cuh"

#define PRE_THREADS 256

// rows: [n][row_len] bytes; out: [row_len][4] int32 (A,T,G,C).  Shared-memory histogram per CTA.
__global__ void __launch_bounds__(PRE_THREADS)
k_base_content(const uint8_t *__restrict__ rows, u32 n, u32 row_len, int *__restrict__ out) {
    extern __shared__ u32 hist[]; // [row_len][4]
    for (u32 i = threadIdx.x; i < row_len * 4; i += PRE_THREADS) hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const u32 wpb = PRE_THREADS / 32;
    for (u32 r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
        const uint8_t *row = rows + (u64)r * row_len;
        for (u32 i = lane; i < row_len; i += 32) {
            const int c = base_cat(row[i]);
            if (c >= 0) atomicAdd(&hist[i * 4 + c], 1u);
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < row_len * 4; i += PRE_THREADS)
        if (hist[i]) atomicAdd(out + i, (int)hist[i]);
}

// One thread per (row, library adapter).  k per adapter is DevAdapter::k_end (the host stores
// int((1 - minSim) * qLen) + 1 there for the library set).  maps: [n_lib] i64.
template <int NW>
__global__ void __launch_bounds__(RES_THREADS)
k_lib_search(const uint8_t *__restrict__ rows, u32 n, u32 row_len, AdapterCtx C, int a,
             long long *__restrict__ maps, u64 *scratch, u64 scratch_stride) {
    const DevAdapter A = C.ad[a];
    const AdapterTables T = adapter_tables(C, a);
    const u64 tid = (u64)blockIdx.x * RES_THREADS + threadIdx.x;
    long long acc = 0;
    for (u64 r = tid; r < n; r += (u64)gridDim.x * RES_THREADS) {
        const u64 lo = r * row_len, hi = lo + row_len;
        const int d = hw_best<NW>(T, rows, lo, hi, A.k_end);
        if (d == 0x7fffffff) continue;
        Myers<NW> s;
        myers_init_hw<NW>(s, T.qlen);
        for (u64 p = lo; p < hi; ++p) {
            myers_step<NW, 0, true>(s, T.hw + (u32)__ldg(rows + p) * NW, 0);
            if (s.score == d) {
                const u64 s0 = shw_start<NW>(T, rows, lo, p, d);
                const int alen = nw_alignment_len<NW>(T, rows, s0, p, d, scratch + tid, scratch_stride);
                acc += alen - d;
                break;
            }
        }
    }
    acc = warp_sum_i64(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd((u64 *)maps, (u64)acc);
}

static __device__ __forceinline__ void fnv_mix(u32 &h, u32 v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        h ^= (v >> (8 * b)) & 0xffu;
        h *= 16777619u;
    }
}

// Pair i uses adapter entry i (tables built from its query) and target [t_off[i], t_off[i+1]).
template <int NW>
__global__ void __launch_bounds__(RES_THREADS)
k_align_pairs(const uint8_t *__restrict__ targets, const u32 *__restrict__ t_off,
              const int *__restrict__ kk, AdapterCtx C, u32 n, int nw_filter,
              tgsf_align_result *__restrict__ out, u64 *scratch, u64 scratch_stride) {
    const u64 tid = (u64)blockIdx.x * RES_THREADS + threadIdx.x;
    for (u64 i = tid; i < n; i += (u64)gridDim.x * RES_THREADS) {
        const DevAdapter A = C.ad[i];
        if (A.nw != nw_filter) continue;
        const AdapterTables T = adapter_tables(C, (int)i);
        const int q = A.qlen;
        const u64 lo = t_off[i], hi = t_off[i + 1];
        tgsf_align_result R;
        R.edit_distance = -1;
        R.n_locations = 0;
        R.align_len = 0;
        R.first_start = R.first_end = R.last_start = R.last_end = 0;
        R.loc_hash = 2166136261u;
        int k = kk[i];
        if (k < 0) k = q;   // E.cpp:194-212
        k = min(k, q);      // E.cpp:565-567
        int d = hw_best<NW>(T, targets, lo, hi, min(k, q - 1));
        bool minus_one = false;
        if (d == 0x7fffffff && k >= q) { // every column scores <= q; column -1 only via W > 0
            d = q;
            minus_one = (q % 64) != 0;
        }
        if (d != 0x7fffffff) {
            R.edit_distance = d;
            int nloc = 0;
            auto add = [&](int st, int en) {
                if (nloc == 0) { R.first_start = st; R.first_end = en; }
                R.last_start = st;
                R.last_end = en;
                fnv_mix(R.loc_hash, (u32)st);
                fnv_mix(R.loc_hash, (u32)en);
                ++nloc;
            };
            if (minus_one) {
                add(0, -1);
                R.align_len = q; // E.cpp:1171-1178
            }
            Myers<NW> s;
            myers_init_hw<NW>(s, q);
            for (u64 p = lo; p < hi; ++p) {
                myers_step<NW, 0, true>(s, T.hw + (u32)__ldg(targets + p) * NW, 0);
                if (s.score != d) continue;
                const u64 s0 = shw_start<NW>(T, targets, lo, p, d);
                if (nloc == 0)
                    R.align_len = nw_alignment_len<NW>(T, targets, s0, p, d, scratch + tid, scratch_stride);
                add((int)(s0 - lo), (int)(p - lo));
            }
            R.n_locations = nloc;
        }
        out[i] = R;
    }
}
