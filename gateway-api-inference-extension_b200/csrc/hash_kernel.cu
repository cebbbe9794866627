// hash_kernel.cu — hashPrompt for a batch (approximateprefix/hashing.go:34-98) on sm_100a.
#include "device_common.cuh"
#include "xxh64.cuh"

namespace eppscore {

// ---------------------------------------------------------------------------------------------
// hashPrompt for a batch: a CTA of 4 warps owns a tile of 32 requests.
//   phase 1  all warps, lanes = blocks: the stripe state ("body") of the 32 blocks of one request at a
//            time — depends on the block's own bytes only, so all of it is parallel;
//   phase 2  warp 0, lanes = requests: the serial chain (one 8-byte tail round + avalanche per link);
//   phase 3  all warps, lanes = blocks: coalesced store of each request's 32 hashes.
// Splitting a tile over 4 warps (instead of one warp per tile) quadruples the warps in flight: at 64K
// requests a one-warp-per-tile layout leaves only ~14 warps per SM and the kernel was latency bound
// (profiles/r1_v0_ncu_hash_prompts.txt).  Requests whose block size is not a multiple of 32 or whose
// start is not 16-byte aligned take the generic serial path in phase 2.
// ---------------------------------------------------------------------------------------------
constexpr int kHashWarps = 4;

template <int BC>  // BC = 64: the default 16-token block, fully unrolled; BC = 0: any multiple of 32
__device__ __forceinline__ uint64_t block_body_state(const uint8_t* p, int bc) {
  uint64_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0 - XP1;
  if (BC == 64) {
    uint4 d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) d[k] = ldg16(p + 16 * k);
#pragma unroll
    for (int st = 0; st < 2; st++) {
      const uint4 x = d[2 * st], y = d[2 * st + 1];
      v1 = xround(v1, ((uint64_t)x.y << 32) | x.x);
      v2 = xround(v2, ((uint64_t)x.w << 32) | x.z);
      v3 = xround(v3, ((uint64_t)y.y << 32) | y.x);
      v4 = xround(v4, ((uint64_t)y.w << 32) | y.z);
    }
    return xfinish_lanes(v1, v2, v3, v4) + (uint64_t)(64 + 8);
  }
  for (int s = 0; s < bc; s += 32) {
    const uint4 x = ldg16(p + s), y = ldg16(p + s + 16);
    v1 = xround(v1, ((uint64_t)x.y << 32) | x.x);
    v2 = xround(v2, ((uint64_t)x.w << 32) | x.z);
    v3 = xround(v3, ((uint64_t)y.y << 32) | y.x);
    v4 = xround(v4, ((uint64_t)y.w << 32) | y.z);
  }
  return xfinish_lanes(v1, v2, v3, v4) + (uint64_t)(bc + 8);
}

template <int BC>
__global__ void __launch_bounds__(kHashWarps * 32, 8) hash_prompts_kernel(HashArgs a) {
  __shared__ uint64_t body[32][33];
  __shared__ unsigned long long s_p[32];
  __shared__ int s_nfull[32], s_fast[32], s_maxfull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bc = BC ? BC : a.block_chars;
  const bool bc_fast = bc > 0 && (bc & 31) == 0;
  const int ntiles = (a.R + 31) >> 5;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // request descriptors (warp 0: lane = request); prev / rem stay in warp 0's registers
    const int r = tile * 32 + lane;
    const uint8_t* p = nullptr;
    uint64_t prev = 0;
    int nfull = 0, rem = 0;
    bool fast = false;
    if (warp == 0) {
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
      s_p[lane] = (unsigned long long)p;
      s_nfull[lane] = nfull;
      s_fast[lane] = fast ? 1 : 0;
      if (lane == 0) s_maxfull = maxfull;
    }
    __syncthreads();
    const int maxfull = s_maxfull;

    for (int c0 = 0; c0 < maxfull; c0 += 32) {
      // phase 1
      for (int q = warp; q < 32; q += kHashWarps) {
        const int b = c0 + lane;
        if (s_fast[q] && b < s_nfull[q])
          body[q][lane] = block_body_state<BC>(reinterpret_cast<const uint8_t*>(s_p[q]) + (size_t)b * bc, bc);
      }
      __syncthreads();
      // phase 2
      if (warp == 0) {
        const int nb = min(32, nfull - c0);
        for (int i = 0; i < nb; i++) {
          if (fast)
            prev = xchain_aligned(body[lane][i], prev);
          else
            prev = xxh64_link<false>(p + (size_t)(c0 + i) * bc, (uint32_t)bc, prev);  // hashing.go:80-87
          body[lane][i] = prev;
        }
      }
      __syncthreads();
      // phase 3
      for (int q = warp; q < 32; q += kHashWarps) {
        const int b = c0 + lane;
        if (tile * 32 + q < a.R && b < s_nfull[q]) a.hashes[(size_t)(tile * 32 + q) * a.stride + b] = body[q][lane];
      }
      __syncthreads();
    }
    if (warp == 0 && r < a.R) {
      if (rem > 0) {                                       // trailing partial block, hashing.go:89-95
        const uint8_t* t = p + (size_t)nfull * bc;
        const uint64_t h = ((reinterpret_cast<uintptr_t>(t) & 7) == 0) ? xxh64_link<true>(t, (uint32_t)rem, prev)
                                                                        : xxh64_link<false>(t, (uint32_t)rem, prev);
        a.hashes[(size_t)r * a.stride + nfull] = h;
      }
      a.n_hashes[r] = (uint16_t)(nfull + (rem > 0 ? 1 : 0));
    }
    __syncthreads();  // descriptors are rewritten by the next tile
  }
}

int launch_hash_prompts(const HashArgs& a, cudaStream_t s, int sm_count) {
  if (a.R <= 0) return 0;
  const int ntiles = (a.R + 31) / 32;
  (void)sm_count;
  const int blocks = ntiles;  // one tile per CTA: the hardware scheduler balances the ~1.7 waves better than a static stride
  if (a.block_chars == 64)
    hash_prompts_kernel<64><<<blocks, kHashWarps * 32, 0, s>>>(a);
  else
    hash_prompts_kernel<0><<<blocks, kHashWarps * 32, 0, s>>>(a);
  return 1;
}

}  // namespace eppscore
