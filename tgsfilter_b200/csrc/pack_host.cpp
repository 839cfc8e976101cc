// Host-side 2-bit packer behind tgsf_pack_bases (include/tgsf.h): AVX2 fast path, 32 bases -> 8 bytes per
// step; any group holding a byte other than upper-case A/C/G/T is left to the caller's scalar path, which
// also records the exception list.  Compiled by the host compiler (nvcc hands .cpp files to g++).
#include <cstdint>
#include <cstring>
#include <immintrin.h>

extern "C" int tgsf_pack_has_avx2() { return __builtin_cpu_supports("avx2") ? 1 : 0; }

// Packs groups of 32 bases starting at group g0 while they are pure ACGT; returns the index of the first
// group it did not pack (== n_groups when all done).  code = ((c >> 1) ^ (c >> 2)) & 3: A0 C1 G2 T3.
extern "C" __attribute__((target("avx2"))) uint64_t tgsf_pack_groups_avx2(const uint8_t *bases, uint8_t *packed,
                                                                          uint64_t g0, uint64_t n_groups) {
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T'), m3 = _mm256_set1_epi8(3);
    const __m256i w16 = _mm256_set1_epi16(0x0401), w32 = _mm256_set1_epi32(0x00100001);
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t g = g0;
    for (; g < n_groups; ++g) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)(bases + 32 * g));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
        if ((uint32_t)_mm256_movemask_epi8(ok) != 0xFFFFFFFFu) break;
        const __m256i code = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), m3);
        const __m256i b = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code, w16), w32), pick);
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(b, 0), hi = (uint32_t)_mm256_extract_epi32(b, 4);
        memcpy(packed + 8 * g, &lo, 4);
        memcpy(packed + 8 * g + 4, &hi, 4);
    }
    return g;
}
