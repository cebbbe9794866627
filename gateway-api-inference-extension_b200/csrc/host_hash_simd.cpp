// host_hash_simd.cpp — hashPrompt (approximateprefix/hashing.go:34-98) for 8 or 32 requests at a time on the host cores, one
// request per 64-bit lane of an AVX-512 register (compiled by g++ with -mavx512f -mavx512dq for this file only; the caller
// checks the CPU at run time: host_path.cu).  XXH64 is a chain of 64x64->64 multiplies, rotates and adds: serial within a
// message, identical across messages — so eight prompts advance in lock step (vpmullq / vprolq), the same way the device
// kernels give every request its own lane.  Used when the block size is a multiple of 32 and the eight requests have the same
// number of full blocks; the trailing partial block and everything else stays on the scalar path.
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) && defined(__AVX512F__) && defined(__AVX512DQ__)
#include <immintrin.h>

namespace {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t XP2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t XP3 = 0x165667B19E3779F9ULL;
constexpr uint64_t XP4 = 0x85EBCA77C2B2AE63ULL;

inline __m512i K(uint64_t v) { return _mm512_set1_epi64((long long)v); }
inline __m512i mul(__m512i a, __m512i b) { return _mm512_mullo_epi64(a, b); }
inline __m512i add(__m512i a, __m512i b) { return _mm512_add_epi64(a, b); }
inline __m512i vxor(__m512i a, __m512i b) { return _mm512_xor_si512(a, b); }
// acc = rotl(acc + in*P2, 31) * P1
inline __m512i xround(__m512i acc, __m512i in) { return mul(_mm512_rol_epi64(add(acc, mul(in, K(XP2))), 31), K(XP1)); }
// h = (h ^ round(0, v)) * P1 + P4
inline __m512i xmerge(__m512i h, __m512i v) { return add(mul(vxor(h, xround(_mm512_setzero_si512(), v)), K(XP1)), K(XP4)); }
inline __m512i xavalanche(__m512i h) {
  h = vxor(h, _mm512_srli_epi64(h, 33));
  h = mul(h, K(XP2));
  h = vxor(h, _mm512_srli_epi64(h, 29));
  h = mul(h, K(XP3));
  return vxor(h, _mm512_srli_epi64(h, 32));
}

}  // namespace

// G groups of eight requests, r = 0..8G-1: the prompt of request r starts at bytes + off[r]; each has `nfull` full blocks of `bc`
// bytes (bc % 32 == 0); seed[r] is its chain start; row[r] is the element offset of its output row.  Writes
// hashes[row[r] + b] for b < nfull and the last chain value of every request to last[r].  One group is latency bound (a
// vpmullq takes ~15 cycles and a block is ~16 of them in series); the groups are independent chains that fill the pipeline.
template <int G>
static inline void hash_groups(const uint8_t* bytes, const int64_t* off, const uint64_t* seed, int32_t bc, int32_t nfull,
                               uint64_t* hashes, const int64_t* row, uint64_t* last) {
  __m512i vo[G], vrow[G], prev[G];
  for (int g = 0; g < G; g++) {
    vo[g] = _mm512_loadu_si512(off + 8 * g);
    vrow[g] = _mm512_slli_epi64(_mm512_loadu_si512(row + 8 * g), 3);  // byte offsets of the output rows
    prev[g] = _mm512_loadu_si512(seed + 8 * g);
  }
  const __m512i lenv = K((uint64_t)bc + 8);
  for (int32_t b = 0; b < nfull; b++) {
    const long long blk = (long long)b * bc;
    __m512i v1[G], v2[G], v3[G], v4[G], h[G];
    for (int g = 0; g < G; g++) {
      v1[g] = K(XP1 + XP2);
      v2[g] = K(XP2);
      v3[g] = _mm512_setzero_si512();
      v4[g] = K(0 - XP1);
    }
    for (int32_t s = 0; s < bc; s += 32) {
      const __m512i d = K((uint64_t)(blk + s));
      for (int g = 0; g < G; g++) {
        const __m512i at = add(vo[g], d);
        v1[g] = xround(v1[g], _mm512_i64gather_epi64(at, bytes, 1));
        v2[g] = xround(v2[g], _mm512_i64gather_epi64(at, bytes + 8, 1));
        v3[g] = xround(v3[g], _mm512_i64gather_epi64(at, bytes + 16, 1));
        v4[g] = xround(v4[g], _mm512_i64gather_epi64(at, bytes + 24, 1));
      }
    }
    for (int g = 0; g < G; g++)
      h[g] = add(add(_mm512_rol_epi64(v1[g], 1), _mm512_rol_epi64(v2[g], 7)), add(_mm512_rol_epi64(v3[g], 12), _mm512_rol_epi64(v4[g], 18)));
    for (int g = 0; g < G; g++) h[g] = xmerge(h[g], v1[g]);
    for (int g = 0; g < G; g++) h[g] = xmerge(h[g], v2[g]);
    for (int g = 0; g < G; g++) h[g] = xmerge(h[g], v3[g]);
    for (int g = 0; g < G; g++) h[g] = xmerge(h[g], v4[g]);
    for (int g = 0; g < G; g++) {
      h[g] = add(h[g], lenv);                                                   // + message length (block + the 8 chain bytes)
      h[g] = vxor(h[g], xround(_mm512_setzero_si512(), prev[g]));               // the message's last 8 bytes: LE64(prev)
    }
    for (int g = 0; g < G; g++) h[g] = add(mul(_mm512_rol_epi64(h[g], 27), K(XP1)), K(XP4));
    for (int g = 0; g < G; g++) h[g] = xavalanche(h[g]);
    for (int g = 0; g < G; g++) {
      _mm512_i64scatter_epi64(reinterpret_cast<uint8_t*>(hashes) + (size_t)b * 8, vrow[g], h[g], 1);
      prev[g] = h[g];
    }
  }
  for (int g = 0; g < G; g++) _mm512_storeu_si512(last + 8 * g, prev[g]);
}

// groups = 1 (8 requests) or 4 (32 requests)
extern "C" void eppscore_host_hash8_avx512(const uint8_t* bytes, const int64_t* off, const uint64_t* seed, int32_t bc,
                                           int32_t nfull, uint64_t* hashes, const int64_t* row, uint64_t* last, int32_t groups) {
  if (groups == 4) hash_groups<4>(bytes, off, seed, bc, nfull, hashes, row, last);
  else hash_groups<1>(bytes, off, seed, bc, nfull, hashes, row, last);
}
extern "C" int32_t eppscore_host_hash8_compiled(void) { return 1; }

#else  // not built with AVX-512: the scalar path does everything

extern "C" void eppscore_host_hash8_avx512(const uint8_t*, const int64_t*, const uint64_t*, int32_t, int32_t, uint64_t*, const int64_t*,
                                           uint64_t*, int32_t) {}
extern "C" int32_t eppscore_host_hash8_compiled(void) { return 0; }

#endif
