// xxh64.cuh — XXH64 (seed 0) for host and device, and the chained block hash of the prefix producer.
//
// Restates the published XXH64 algorithm that github.com/cespare/xxhash/v2 v2.3.0 implements
// (reference go.mod:6; call sites approximateprefix/hashing.go:70-94).  The reference chains
//   h_{-1} = XXH64(model || salt),   h_i = XXH64(block_i || LE64(h_{i-1}))
// so the 8 chain bytes are always the LAST 8 bytes of each message.  When block_chars is a
// multiple of 32 the four-lane stripe state of a block depends only on the block's own bytes
// ("body state"), and the serial part of a link shrinks to one 8-byte tail round + avalanche.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define XXH_HD __host__ __device__ __forceinline__
#else
#define XXH_HD inline
#endif

namespace eppscore {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t XP2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t XP3 = 0x165667B19E3779F9ULL;
constexpr uint64_t XP4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t XP5 = 0x27D4EB2F165667C5ULL;

// Device forms: a 64-bit x*P + acc (low 64 bits) is exactly three 32-bit multiply-adds (IMAD.WIDE.U32 with the
// 64-bit addend, then the two cross terms into the high word — the compiler's generic 64-bit multiply followed by
// an add takes five), and a 64-bit rotate is two funnel shifts.  hash_bodies_kernel is bound by exactly
// these integer instructions.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint64_t xmuladd(uint64_t x, uint64_t p, uint64_t acc) {
  const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), pl = (uint32_t)p, ph = (uint32_t)(p >> 32);
  uint64_t w;  // IMAD.WIDE.U32 with the 64-bit addend
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(xl), "r"(pl), "l"(acc));
  uint32_t hi = (uint32_t)(w >> 32);
  hi = xh * pl + hi;                           // the cross terms only touch the high word
  hi = xl * ph + hi;
  return ((uint64_t)hi << 32) | (uint32_t)w;
}
__device__ __forceinline__ uint64_t xrotl(uint64_t x, int r) {  // r is a compile-time constant at every call site
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  if (r >= 32) {
    const uint32_t t = lo;
    lo = hi;
    hi = t;
    r -= 32;
  }
  if (r == 0) return ((uint64_t)hi << 32) | lo;
  const uint32_t nh = __funnelshift_l(lo, hi, r), nl = __funnelshift_l(hi, lo, r);
  return ((uint64_t)nh << 32) | nl;
}
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t in) {
  return xmuladd(xrotl(xmuladd(in, XP2, acc), 31), XP1, 0);
}
__device__ __forceinline__ uint64_t xmerge(uint64_t h, uint64_t v) { return xmuladd(h ^ xround(0, v), XP1, XP4); }
#else
XXH_HD uint64_t xmuladd(uint64_t x, uint64_t p, uint64_t acc) { return x * p + acc; }
XXH_HD uint64_t xrotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
XXH_HD uint64_t xround(uint64_t acc, uint64_t in) {
  acc += in * XP2;
  acc = xrotl(acc, 31);
  return acc * XP1;
}
XXH_HD uint64_t xmerge(uint64_t h, uint64_t v) {
  h ^= xround(0, v);
  return h * XP1 + XP4;
}
#endif
XXH_HD uint64_t xavalanche(uint64_t h) {
  h ^= h >> 33;
  h = xmuladd(h, XP2, 0);
  h ^= h >> 29;
  h = xmuladd(h, XP3, 0);
  h ^= h >> 32;
  return h;
}
XXH_HD uint64_t xfinish_lanes(uint64_t v1, uint64_t v2, uint64_t v3, uint64_t v4) {
  uint64_t h = xrotl(v1, 1) + xrotl(v2, 7) + xrotl(v3, 12) + xrotl(v4, 18);
  h = xmerge(h, v1);
  h = xmerge(h, v2);
  h = xmerge(h, v3);
  h = xmerge(h, v4);
  return h;
}
// The serial part of one link when block_chars % 32 == 0: body = lanes-merged state + len.
XXH_HD uint64_t xchain_aligned(uint64_t body, uint64_t prev) {
  uint64_t h = body ^ xround(0, prev);
  h = xmuladd(xrotl(h, 27), XP1, XP4);
  return xavalanche(h);
}

// A "virtual message" = data[0..dlen) || LE64(prev), read without materialising it.
// ALIGNED8: data is 8-byte aligned, so whole data words are single 64-bit loads.
template <bool ALIGNED8>
struct LinkMessage {
  const uint8_t* d;
  uint32_t dlen;
  uint64_t prev;
  XXH_HD uint8_t byte(uint32_t off) const {
    return off < dlen ? d[off] : (uint8_t)(prev >> (8 * (off - dlen)));
  }
  XXH_HD uint64_t u64(uint32_t off) const {
    if (off + 8 <= dlen) {
      if (ALIGNED8) return *reinterpret_cast<const uint64_t*>(d + off);
      uint64_t v = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) v |= (uint64_t)d[off + i] << (8 * i);
      return v;
    }
    if (off >= dlen) return prev >> (8 * (off - dlen));
    uint64_t v = 0;
    const uint32_t nd = dlen - off;  // 1..7 data bytes, then chain bytes
    for (uint32_t i = 0; i < nd; i++) v |= (uint64_t)d[off + i] << (8 * i);
    return v | (prev << (8 * nd));
  }
  XXH_HD uint32_t u32(uint32_t off) const {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v |= (uint32_t)byte(off + i) << (8 * i);
    return v;
  }
};

// XXH64(data || LE64(prev)), any dlen (hashing.go:79-85 and the partial block :89-95).
template <bool ALIGNED8>
XXH_HD uint64_t xxh64_link(const uint8_t* data, uint32_t dlen, uint64_t prev) {
  LinkMessage<ALIGNED8> m{data, dlen, prev};
  const uint32_t len = dlen + 8;
  uint32_t p = 0;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0 - XP1;
    do {
      v1 = xround(v1, m.u64(p));
      v2 = xround(v2, m.u64(p + 8));
      v3 = xround(v3, m.u64(p + 16));
      v4 = xround(v4, m.u64(p + 24));
      p += 32;
    } while (p + 32 <= len);
    h = xfinish_lanes(v1, v2, v3, v4);
  } else {
    h = XP5;
  }
  h += (uint64_t)len;
  while (p + 8 <= len) {
    h ^= xround(0, m.u64(p));
    h = xrotl(h, 27) * XP1 + XP4;
    p += 8;
  }
  if (p + 4 <= len) {
    h ^= (uint64_t)m.u32(p) * XP1;
    h = xrotl(h, 23) * XP2 + XP3;
    p += 4;
  }
  while (p < len) {
    h ^= (uint64_t)m.byte(p) * XP5;
    h = xrotl(h, 11) * XP1;
    p++;
  }
  return xavalanche(h);
}

// Plain one-shot XXH64 over a byte buffer (host helper: model seed, tests).
inline uint64_t xxh64_host(const void* data, size_t len, uint64_t seed) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  const uint8_t* end = p + len;
  auto rd64 = [](const uint8_t* q) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)q[i] << (8 * i);
    return v;
  };
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    for (; p + 32 <= end; p += 32) {
      v1 = xround(v1, rd64(p));
      v2 = xround(v2, rd64(p + 8));
      v3 = xround(v3, rd64(p + 16));
      v4 = xround(v4, rd64(p + 24));
    }
    h = xfinish_lanes(v1, v2, v3, v4);
  } else {
    h = seed + XP5;
  }
  h += (uint64_t)len;
  for (; p + 8 <= end; p += 8) {
    h ^= xround(0, rd64(p));
    h = xrotl(h, 27) * XP1 + XP4;
  }
  if (p + 4 <= end) {
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i);
    h ^= (uint64_t)v * XP1;
    h = xrotl(h, 23) * XP2 + XP3;
    p += 4;
  }
  for (; p < end; p++) {
    h ^= (uint64_t)(*p) * XP5;
    h = xrotl(h, 11) * XP1;
  }
  return xavalanche(h);
}

}  // namespace eppscore
