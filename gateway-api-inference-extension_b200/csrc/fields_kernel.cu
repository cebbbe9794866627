// fields_kernel.cu — len(strings.Fields(prompt)) for a batch: the latency path's input_token_length
// (predictedlatency/plugin.go:286, training.go:51).  One warp per request streams the prompt once (16 bytes per
// lane per step, coalesced); a byte is "space" when it is an ASCII space or lies inside one of the UTF-8 encodings
// of the Unicode White_Space runes Go's unicode.IsSpace accepts.  Go decodes invalid / short sequences as U+FFFD of
// width 1, and a lead byte is never a continuation byte, so matching those byte patterns at every position is
// equivalent to decoding runes (fuzzed against the oracle's rune decoder: tests/test_oracle_golden.py).
// Fields = positions that are not space and whose predecessor is (the start of the string counts as space).
#include "device_common.cuh"

namespace eppscore {

constexpr int kFieldWarps = 8;

__device__ __forceinline__ uint32_t byte_at(const uint32_t (&w)[5], int i) {  // i in [0, 20)
  return (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
}

__global__ void __launch_bounds__(kFieldWarps * 32) count_fields_kernel(const uint8_t* __restrict__ bytes,
                                                                         const int64_t* __restrict__ off,
                                                                         const int32_t* __restrict__ len, int R,
                                                                         int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kFieldWarps + (threadIdx.x >> 5), nw = gridDim.x * kFieldWarps;
  for (int r = gw; r < R; r += nw) {
    const uint8_t* p = bytes + off[r];
    const int64_t n = len ? (int64_t)len[r] : off[r + 1] - off[r];
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    int count = 0;
    uint32_t carry_space = 1;  // the predecessor of byte 0 counts as space
    uint32_t carry_tail = 0;   // last two bytes of the previous step (bits 0-15)
    for (int64_t base = 0; base < n; base += 512) {
      const int64_t o = base + lane * 16;
      // window: [2 bytes before | my 16 bytes | 2 bytes after] as five little-endian words, bytes 2..17 are mine
      uint32_t mine[4] = {0, 0, 0, 0};
      if (aligned && o + 16 <= n) {
        const uint4 v = ldg16(p + o);
        mine[0] = v.x;
        mine[1] = v.y;
        mine[2] = v.z;
        mine[3] = v.w;
      } else {
        for (int i = 0; i < 16; i++)
          if (o + i < n) mine[i >> 2] |= (uint32_t)p[o + i] << ((i & 3) * 8);
      }
      uint32_t before = __shfl_up_sync(0xffffffffu, mine[3] >> 16, 1);
      if (lane == 0) before = carry_tail;
      uint32_t after = __shfl_down_sync(0xffffffffu, mine[0] & 0xffffu, 1);
      if (lane == 31) {
        after = 0;
        if (o + 16 < n) after = p[o + 16];
        if (o + 17 < n) after |= (uint32_t)p[o + 17] << 8;
      }
      const uint32_t w[5] = {before | (mine[0] << 16), (mine[0] >> 16) | (mine[1] << 16), (mine[1] >> 16) | (mine[2] << 16),
                             (mine[2] >> 16) | (mine[3] << 16), (mine[3] >> 16) | (after << 16)};
      // ASCII spaces (\t \n \v \f \r and ' '), four bytes per word: for the low 7 bits v of a byte, bit 7 of v+0x77 is set iff
      // v >= 9, of v+0x72 iff v >= 14, of ~((v^0x20)+0x7f) iff v == 0x20; bytes >= 0x80 are masked out; the four flag
      // bits are gathered with one multiply
      uint32_t my_sp = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t x = mine[q], v = x & 0x7f7f7f7fu;
        const uint32_t in_9_13 = (v + 0x77777777u) & ~(v + 0x72727272u);
        const uint32_t is_20 = ~((v ^ 0x20202020u) + 0x7f7f7f7fu);
        const uint32_t m = (in_9_13 | is_20) & ~x & 0x80808080u;
        my_sp |= ((((m >> 7) * 0x00204081u) >> 21) & 0xfu) << (4 * q);
      }
      // multi-byte space runes: only when the 20-byte window holds a byte >= 0x80 at all
      if (((w[0] | w[1] | w[2] | w[3] | w[4]) & 0x80808080u) != 0) {
        uint32_t sp = 0;  // bit i: window byte i lies inside a space rune
        for (int i = 0; i < 18; i++) {
          const uint32_t b0 = byte_at(w, i), b1 = byte_at(w, i + 1), b2 = byte_at(w, i + 2);
          if (b0 == 0xC2u && (b1 == 0x85u || b1 == 0xA0u)) sp |= 3u << i;
          const bool m3 = (b0 == 0xE1u && b1 == 0x9Au && b2 == 0x80u) ||
                          (b0 == 0xE2u && b1 == 0x80u && ((b2 - 0x80u) <= 0x0Au || b2 == 0xA8u || b2 == 0xA9u || b2 == 0xAFu)) ||
                          (b0 == 0xE2u && b1 == 0x81u && b2 == 0x9Fu) || (b0 == 0xE3u && b1 == 0x80u && b2 == 0x80u);
          if (m3) sp |= 7u << i;
        }
        my_sp |= (sp >> 2) & 0xffffu;
      }
      // (bytes past the end of the prompt read as 0: never a space, never part of a pattern — and masked out below)
      uint32_t prev = __shfl_up_sync(0xffffffffu, my_sp >> 15, 1);  // is the previous lane's last byte a space?
      if (lane == 0) prev = carry_space;
      const int64_t left = n - o;
      const uint32_t valid = left >= 16 ? 0xffffu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
      const uint32_t starts = ~my_sp & ((my_sp << 1) | prev) & valid;
      count += __popc(starts);
      carry_space = __shfl_sync(0xffffffffu, my_sp >> 15, 31);
      carry_tail = __shfl_sync(0xffffffffu, mine[3] >> 16, 31);
    }
#pragma unroll
    for (int o2 = 16; o2; o2 >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o2);
    if (lane == 0) out[r] = count;
  }
}

int launch_count_fields(const uint8_t* bytes, const int64_t* off, const int32_t* len, int R, int32_t* out, cudaStream_t s,
                        int sm_count) {
  if (R <= 0) return 0;
  long long blocks = ((long long)R + kFieldWarps - 1) / kFieldWarps;
  const long long cap = (long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  count_fields_kernel<<<(unsigned)blocks, kFieldWarps * 32, 0, s>>>(bytes, off, len, R, out);
  return 1;
}

}  // namespace eppscore
