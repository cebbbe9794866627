// hash_kernel.cu — hashPrompt for a batch (approximateprefix/hashing.go:34-98) on sm_100a.
#include "device_common.cuh"
#include "xxh64.cuh"

namespace eppscore {

// ---------------------------------------------------------------------------------------------
// hashPrompt for a batch: one warp owns a tile of 32 requests.
//   phase 1  lanes = blocks: body state of 32 blocks of one request at a time (block bytes only)
//   phase 2  lanes = requests: the serial chain (one tail round + avalanche per link)
//   phase 3  lanes = blocks: coalesced store of the 32 hashes of each request
// Requests whose block size is not a multiple of 32 or whose start is not 16-byte aligned take the
// generic serial path in phase 2 (every block fully hashed by the request's lane).
// ---------------------------------------------------------------------------------------------
constexpr int kHashWarps = 4;

__device__ __forceinline__ uint64_t block_body_state(const uint8_t* p, int bc) {
  uint64_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0 - XP1;
  for (int s = 0; s < bc; s += 32) {
    const uint4 x = ldg16(p + s), y = ldg16(p + s + 16);
    v1 = xround(v1, ((uint64_t)x.y << 32) | x.x);
    v2 = xround(v2, ((uint64_t)x.w << 32) | x.z);
    v3 = xround(v3, ((uint64_t)y.y << 32) | y.x);
    v4 = xround(v4, ((uint64_t)y.w << 32) | y.z);
  }
  return xfinish_lanes(v1, v2, v3, v4) + (uint64_t)(bc + 8);
}

__global__ void __launch_bounds__(kHashWarps * 32) hash_prompts_kernel(HashArgs a) {
  __shared__ uint64_t s_body[kHashWarps][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kHashWarps + warp, nw = gridDim.x * kHashWarps;
  const int bc = a.block_chars;
  const bool bc_fast = bc > 0 && (bc & 31) == 0;
  const int ntiles = (a.R + 31) >> 5;
  uint64_t(*body)[33] = s_body[warp];

  for (int tile = gw; tile < ntiles; tile += nw) {
    const int r = tile * 32 + lane;
    const uint8_t* p = nullptr;
    uint64_t prev = 0;
    int nfull = 0, rem = 0;
    bool fast = false;
    if (r < a.R) {
      const int64_t o = a.off[r];
      int64_t len = a.len ? (int64_t)a.len[r] : a.off[r + 1] - o;
      p = a.bytes + o;
      prev = a.seed ? a.seed[r] : 0ULL;
      if (bc > 0 && len >= bc) {                         // hashing.go:51-60
        const int64_t cap = (int64_t)bc * (int64_t)a.max_blocks;
        if (len > cap) len = cap;                        // :62-65
        nfull = (int)(len / bc);
        rem = (int)(len - (int64_t)nfull * bc);
      }
      fast = bc_fast && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    }
    int maxfull = nfull;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxfull = max(maxfull, __shfl_xor_sync(0xffffffffu, maxfull, o));

    for (int c0 = 0; c0 < maxfull; c0 += 32) {
      // phase 1
      for (int q = 0; q < 32; q++) {
        const int nf_q = __shfl_sync(0xffffffffu, nfull, q);
        const int fast_q = __shfl_sync(0xffffffffu, (int)fast, q);
        const unsigned long long p_q = __shfl_sync(0xffffffffu, (unsigned long long)p, q);
        if (!fast_q) continue;
        const int b = c0 + lane;
        if (b < nf_q) body[q][lane] = block_body_state(reinterpret_cast<const uint8_t*>(p_q) + (size_t)b * bc, bc);
      }
      __syncwarp();
      // phase 2
      const int nb = min(32, nfull - c0);
      for (int i = 0; i < nb; i++) {
        if (fast)
          prev = xchain_aligned(body[lane][i], prev);
        else
          prev = xxh64_link<false>(p + (size_t)(c0 + i) * bc, (uint32_t)bc, prev);  // hashing.go:80-87
        body[lane][i] = prev;
      }
      __syncwarp();
      // phase 3
      for (int q = 0; q < 32; q++) {
        const int nf_q = __shfl_sync(0xffffffffu, nfull, q);
        const int b = c0 + lane;
        if (b < nf_q) a.hashes[(size_t)(tile * 32 + q) * a.stride + b] = body[q][lane];
      }
      __syncwarp();
    }
    if (r < a.R) {
      if (rem > 0) {                                     // trailing partial block, hashing.go:89-95
        const uint8_t* t = p + (size_t)nfull * bc;
        const uint64_t h = ((reinterpret_cast<uintptr_t>(t) & 7) == 0) ? xxh64_link<true>(t, (uint32_t)rem, prev)
                                                                        : xxh64_link<false>(t, (uint32_t)rem, prev);
        a.hashes[(size_t)r * a.stride + nfull] = h;
      }
      a.n_hashes[r] = (uint16_t)(nfull + (rem > 0 ? 1 : 0));
    }
  }
}

int launch_hash_prompts(const HashArgs& a, cudaStream_t s, int /*sm_count*/) {
  if (a.R <= 0) return 0;
  const int ntiles = (a.R + 31) / 32;
  const int blocks = (ntiles + kHashWarps - 1) / kHashWarps;
  hash_prompts_kernel<<<blocks, kHashWarps * 32, 0, s>>>(a);
  return 1;
}

}  // namespace eppscore
