// kernels.cu — hand-written sm_100a kernels of the Endpoint-Picker hot path.
//
//   hash_prompts_kernel      hashPrompt for a batch            (approximateprefix/hashing.go:34-98)
//   prepare_snapshot_kernel  request-independent scorer terms   (kvcache_utilization.go:76-82, queue.go:78-108,
//                            + LoRA class planes                 lora_affinity.go:76-102)
//   score_pick_fused_kernel  matchLongestPrefix + 4 scorers + weighted sum + arg-max, one warp per request,
//                            the R x M score matrix never touches HBM
//                                                               (approximateprefix/plugin.go:219-235,
//                                                                scheduler_profile.go:151-192, maxscore/picker.go:87-115)
//   score_pick_dense_kernel  same Score+Pick over caller-supplied float4 feature rows streamed from HBM
//
// No tensor cores: there is no dense contraction on this path — it is integer hashing, table probes,
// float64 adds and row arg-max.  float64 uses explicit _rn intrinsics so no FMA contraction can occur
// (the reference's GOARCH=amd64 build never fuses, SURVEY.md "Key facts").
#include "kernels.cuh"
#include "xxh64.cuh"

namespace eppscore {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double clamp01(double s) {  // enforceScoreRange, scheduler_profile.go:194-202
  if (s < 0.0) return 0.0;
  if (s > 1.0) return 1.0;
  return s;
}
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ long long shfl_xor_i64(long long v, int o) {
  return __shfl_xor_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

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

int launch_hash_prompts(const HashArgs& a, cudaStream_t s) {
  if (a.R <= 0) return 0;
  const int ntiles = (a.R + 31) / 32;
  const int blocks = (ntiles + kHashWarps - 1) / kHashWarps;
  hash_prompts_kernel<<<blocks, kHashWarps * 32, 0, s>>>(a);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Snapshot preparation (once per metrics snapshot, not per request).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) prepare_snapshot_kernel(PrepareArgs a) {
  __shared__ long long s_red[2][2][32];  // [queue|running][min|max][warp]
  __shared__ long long s_mm[2][2];
  const int M = a.geo.M, Mpad = a.geo.Mpad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // min / max of WaitingQueueSize and RunningRequestsSize over ALL endpoints (queue.go:79-91)
  for (int which = 0; which < 2; which++) {
    const int64_t* q = which == 0 ? a.queue : a.running;
    long long mn = 0x7fffffffffffffffLL, mx = (long long)0x8000000000000000ULL;
    if (q)
      for (int m = tid; m < M; m += blockDim.x) {
        const long long v = q[m];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const long long omn = shfl_xor_i64(mn, o), omx = shfl_xor_i64(mx, o);
      mn = omn < mn ? omn : mn;
      mx = omx > mx ? omx : mx;
    }
    if (lane == 0) {
      s_red[which][0][warp] = mn;
      s_red[which][1][warp] = mx;
    }
  }
  __syncthreads();
  if (tid < 4) {
    const int which = tid >> 1, isx = tid & 1;
    long long v = s_red[which][isx][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      const long long o = s_red[which][isx][w];
      v = isx ? (o > v ? o : v) : (o < v ? o : v);
    }
    s_mm[which][isx] = v;
  }
  __syncthreads();

  // per-scorer terms: clamp(score) * weight, the rounded product of scheduler_profile.go:168
  for (int s = 0; s < a.n_scorers; s++) {
    double* out = a.term[s];
    if (!out) continue;
    const int kind = a.kind[s];
    const double w = a.weight[s];
    for (int m = tid; m < Mpad; m += blockDim.x) {
      double sc = 0.0;
      if (m < M) {
        if (kind == 1) {
          sc = __dsub_rn(1.0, a.kv[m]);                                   // kvcache_utilization.go:79
        } else if (kind == 0 || kind == 4) {
          const int which = kind == 0 ? 0 : 1;
          const int64_t* q = which == 0 ? a.queue : a.running;
          const long long mn = s_mm[which][0], mx = s_mm[which][1];
          if (!q || mx == mn)
            sc = 1.0;                                                      // queue.go:95-98
          else
            sc = __ddiv_rn(__ll2double_rn(mx - q[m]), __ll2double_rn(mx - mn));  // queue.go:99
        } else {
          const double* col = a.col[kind - 8];
          sc = col ? col[m] : 0.0;
        }
      }
      out[m] = (m < M) ? __dmul_rn(clamp01(sc), w) : 0.0;
    }
  }
  __syncthreads();
  // folded leading runs: ((0.0 + t0) + t1) + ... in scorer order
  for (int pass = 0; pass < 2; pass++) {
    double* out = pass == 0 ? a.fold_unmasked : a.fold_masked;
    const int n = pass == 0 ? a.fold_unmasked_n : a.fold_masked_n;
    if (!out) continue;
    for (int m = tid; m < Mpad; m += blockDim.x) {
      double acc = 0.0;
      for (int s = 0; s < n; s++) acc = __dadd_rn(acc, a.term[s][m]);
      out[m] = acc;
    }
  }
  // LoRA class planes (permuted layout): class 3 active, 2 has capacity, 1 waiting, 0 none — the
  // precedence of lora_affinity.go:84-99. Row A is the "adapter not in the dictionary" row.
  if (a.cls_lo) {
    const int rw = a.geo.row_words, log_epl = a.geo.log_epl, epl = 1 << log_epl;
    const int total = (a.A + 1) * rw;
    for (int t = tid; t < total; t += blockDim.x) {
      const int ai = t / rw, word = t - ai * rw;
      const int j = word >> 5, ln = word & 31;
      uint32_t lo = 0, hi = 0;
      for (int k = 0; k < epl; k++) {
        const int m = ((j << log_epl) + k) * 32 + ln;
        if (m >= M) continue;
        bool active = false, waiting = false;
        if (ai < a.A && a.act && a.wait) {
          const uint64_t bit = 1ULL << (ai & 63);
          active = (a.act[(size_t)m * a.lora_words + (ai >> 6)] & bit) != 0;
          waiting = (a.wait[(size_t)m * a.lora_words + (ai >> 6)] & bit) != 0;
        }
        const int nm = a.nmodels ? a.nmodels[m] : 0, mxm = a.maxm ? a.maxm[m] : 0;
        const int cls = active ? 3 : (nm < mxm ? 2 : (waiting ? 1 : 0));
        lo |= (uint32_t)(cls & 1) << k;
        hi |= (uint32_t)(cls >> 1) << k;
      }
      a.cls_lo[t] = lo;
      a.cls_hi[t] = hi;
    }
  }
}

int launch_prepare_snapshot(const PrepareArgs& a, cudaStream_t s) {
  prepare_snapshot_kernel<<<1, 1024, 0, s>>>(a);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Score + Pick
// ---------------------------------------------------------------------------------------------
constexpr int kScoreWarps = 8;

struct Best {
  double score;
  int32_t m;     // -1 = none yet
  int32_t cnt;
  uint32_t prio;
};

__device__ __forceinline__ void best_update(Best& b, double s, int m, int tie_mode, uint32_t areq, uint32_t seed_hi) {
  if (b.m < 0 || s > b.score) {
    b.score = s;
    b.m = m;
    b.cnt = 1;
    if (tie_mode) b.prio = lowbias32(areq + (uint32_t)m * 0x9E3779B1U + seed_hi);
  } else if (s == b.score) {
    b.cnt++;
    if (tie_mode) {
      const uint32_t pr = lowbias32(areq + (uint32_t)m * 0x9E3779B1U + seed_hi);
      if (pr > b.prio) {
        b.prio = pr;
        b.m = m;
      }
    }
  }
}

__device__ __forceinline__ void best_warp_reduce(Best& b, int tie_mode) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double os = shfl_xor_f64(b.score, o);
    const int om = __shfl_xor_sync(0xffffffffu, b.m, o);
    const int oc = __shfl_xor_sync(0xffffffffu, b.cnt, o);
    const uint32_t op = __shfl_xor_sync(0xffffffffu, b.prio, o);
    if (om >= 0) {
      if (b.m < 0 || os > b.score) {
        b.score = os;
        b.m = om;
        b.cnt = oc;
        b.prio = op;
      } else if (os == b.score) {
        b.cnt += oc;
        const bool take = tie_mode ? (op > b.prio || (op == b.prio && om < b.m)) : (om < b.m);
        if (take) {
          b.m = om;
          b.prio = op;
        }
      }
    }
  }
}

// Per-request prefix LUT: lut[c] = clamp(c/total)*w for c <= min(total,kLutMax) (prefix/plugin.go:108-110)
__device__ __forceinline__ double prefix_term_direct(int c, int total, double w) {
  double sc = 0.0;
  if (total != 0) sc = __ddiv_rn((double)c, (double)total);
  return __dmul_rn(clamp01(sc), w);
}

template <int LOG_EPL, int J, int NP, bool MASKED>
__global__ void __launch_bounds__(kScoreWarps * 32) score_pick_fused_kernel(const __grid_constant__ ScoreArgs a) {
  constexpr int EPL = 1 << LOG_EPL;
  constexpr int MPAD = J * 32 * EPL;
  constexpr int RW = J * 32;
  constexpr int MASKW = (J * EPL + 31) / 32;  // registers holding this row's candidate mask words
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Plan& plan = a.plan;
  double* s_term = reinterpret_cast<double*>(smem_raw);                       // [n_terms][MPAD]
  long long* s_q = reinterpret_cast<long long*>(s_term + (size_t)plan.n_terms * MPAD);  // MASKED: [2][MPAD]
  double* s_lut = reinterpret_cast<double*>(s_q + (MASKED ? 2 * MPAD : 0));   // [warps][kLutMax+1]
  const int M = a.geo.M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // stage the endpoint tile once per CTA
  for (int t = 0; t < plan.n_terms; t++)
    for (int m = threadIdx.x; m < MPAD; m += blockDim.x) s_term[(size_t)t * MPAD + m] = a.term[t][m];
  if (MASKED) {
    for (int which = 0; which < 2; which++)
      for (int m = threadIdx.x; m < MPAD; m += blockDim.x)
        s_q[which * MPAD + m] = (a.minmax_q[which] && m < M) ? a.minmax_q[which][m] : 0;
  }
  __syncthreads();

  double* lut = s_lut + warp * (kLutMax + 1);
  int lut_total = -1;
  int prefix_step = -1;
  bool has_minmax[2] = {false, false};
  for (int s = 0; s < plan.n_steps; s++) {
    if (plan.kind[s] == STEP_PREFIX && prefix_step < 0) prefix_step = s;
    if (plan.kind[s] == STEP_MINMAX) has_minmax[plan.arg[s]] = true;
  }
  const bool want_prefix = prefix_step >= 0 || a.match_out != nullptr;
  const int tie_mode = plan.tie_mode;

  const int gw = blockIdx.x * kScoreWarps + warp, nw = gridDim.x * kScoreWarps;
  for (int r = gw; r < a.R; r += nw) {
    // ---------------- matchLongestPrefix (plugin.go:219-235) into bit-sliced counters ----------------
    uint32_t P[J][NP];
#pragma unroll
    for (int j = 0; j < J; j++)
#pragma unroll
      for (int p = 0; p < NP; p++) P[j][p] = 0;
    int total = 0;
    if (want_prefix && a.hashes) {
      const int n = a.n_hashes[r];
      total = n;
      bool stop = false;
      for (int c0 = 0; c0 < n && !stop; c0 += 32) {
        const int i = c0 + lane;
        uint32_t row = kEmptyRow;
        if (i < n && a.slots) {
          const uint64_t h = a.hashes[(size_t)r * a.hash_stride + i];
          uint64_t idx = h & a.slot_mask;
          for (;;) {                                        // indexer.Get, indexer.go:86-102
            const uint4 sv = ldg16(&a.slots[idx]);
            if (sv.z == kEmptyRow) break;                   // never-used slot: hash unknown
            if ((((uint64_t)sv.y << 32) | sv.x) == h) {
              if (sv.w != 0) row = sv.z;                    // cnt==0: emptied set == deleted key
              break;
            }
            idx = (idx + 1) & a.slot_mask;
          }
        }
        const uint32_t miss = __ballot_sync(0xffffffffu, row == kEmptyRow);
        const int nh = miss ? (__ffs(miss) - 1) : 32;       // blocks matched before the first global miss
        if (nh < 32) stop = true;
        for (int i2 = 0; i2 < nh; i2++) {
          const uint32_t rr = __shfl_sync(0xffffffffu, row, i2);
          const uint32_t* rp = a.rows + (size_t)rr * RW + lane;
#pragma unroll
          for (int j = 0; j < J; j++) {
            uint32_t carry = __ldg(rp + j * 32);            // res[server]++ for every server in the set
#pragma unroll
            for (int p = 0; p < NP; p++) {
              const uint32_t t = P[j][p] & carry;
              P[j][p] ^= carry;
              carry = t;
            }
          }
        }
      }
    }
    if (prefix_step >= 0 && total != lut_total) {
      const double w = plan.weight[prefix_step];
      const int top = total < kLutMax ? total : kLutMax;
      __syncwarp();
      for (int c = lane; c <= top; c += 32) lut[c] = prefix_term_direct(c, total, w);
      lut_total = total;
      __syncwarp();
    }

    // ---------------- per-request scorer inputs ----------------
    int ad = a.adapter_id ? a.adapter_id[r] : -1;
    if (ad < 0 || ad >= a.A) ad = a.A;
    uint32_t maskw[MASKW];
    if (MASKED) {
#pragma unroll
      for (int q = 0; q < MASKW; q++) {
        const int wi = q * 32 + lane;
        maskw[q] = wi < a.mask_words ? a.cand_mask[(size_t)r * a.mask_words + wi] : 0u;
      }
    }
    // candidate-set min/max for STEP_MINMAX (queue.go:79-91 over the FILTERED endpoints)
    long long mn[2] = {0, 0}, mx[2] = {0, 0};
    if (MASKED && (has_minmax[0] || has_minmax[1])) {
      mn[0] = mn[1] = 0x7fffffffffffffffLL;
      mx[0] = mx[1] = (long long)0x8000000000000000ULL;
#pragma unroll
      for (int j = 0; j < J; j++)
#pragma unroll 4
        for (int k = 0; k < EPL; k++) {
          const int t = j * EPL + k, m = t * 32 + lane;
          const uint32_t mwv = __shfl_sync(0xffffffffu, maskw[t >> 5], t & 31);
          if (((mwv >> lane) & 1u) && m < M) {
#pragma unroll
            for (int which = 0; which < 2; which++) {
              const long long v = s_q[which * MPAD + m];
              mn[which] = v < mn[which] ? v : mn[which];
              mx[which] = v > mx[which] ? v : mx[which];
            }
          }
        }
#pragma unroll
      for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int which = 0; which < 2; which++) {
          const long long omn = shfl_xor_i64(mn[which], o), omx = shfl_xor_i64(mx[which], o);
          mn[which] = omn < mn[which] ? omn : mn[which];
          mx[which] = omx > mx[which] ? omx : mx[which];
        }
    }

    const uint32_t areq = lowbias32((uint32_t)(uint64_t)(a.request_base + r) ^ plan.seed_lo);
    Best best;
    best.score = 0.0;
    best.m = -1;
    best.cnt = 0;
    best.prio = 0;

    // ---------------- Score (scheduler_profile.go:151-174) + Pick (maxscore/picker.go:87-115) ----------------
#pragma unroll
    for (int j = 0; j < J; j++) {
      uint32_t any = 0;
#pragma unroll
      for (int p = 0; p < NP; p++) any |= P[j][p];
      uint32_t clo = 0, chi = 0;
      if (a.cls_lo) {
        clo = __ldg(a.cls_lo + (size_t)ad * RW + j * 32 + lane);
        chi = __ldg(a.cls_hi + (size_t)ad * RW + j * 32 + lane);
      }
#pragma unroll 4
      for (int k = 0; k < EPL; k++) {
        const int t = j * EPL + k, m = t * 32 + lane;
        bool cand = m < M;
        if (MASKED) {
          const uint32_t mwv = __shfl_sync(0xffffffffu, maskw[t >> 5], t & 31);
          cand = cand && ((mwv >> lane) & 1u);
        }
        int c = 0;
        if ((any >> k) & 1u) {
#pragma unroll
          for (int p = 0; p < NP; p++) c |= (int)((P[j][p] >> k) & 1u) << p;
        }
        if (a.match_out && m < M) a.match_out[(size_t)r * M + m] = (uint16_t)c;
        const int cls = (int)((clo >> k) & 1u) | ((int)((chi >> k) & 1u) << 1);
        double acc = 0.0;  // weightedScorePerEndpoint[endpoint] = float64(0), scheduler_profile.go:156-158
        for (int s = 0; s < plan.n_steps; s++) {
          double term;
          switch (plan.kind[s]) {
            case STEP_EP_TERM: term = s_term[(size_t)plan.arg[s] * MPAD + m]; break;
            case STEP_PREFIX:
              term = (s == prefix_step && total <= kLutMax) ? lut[c] : prefix_term_direct(c, total, plan.weight[s]);
              break;
            case STEP_LORA: term = plan.lora_term[s][cls]; break;
            case STEP_MINMAX: {
              const int which = plan.arg[s];
              double sc = 1.0;                               // queue.go:95-98
              if (MASKED && mx[which] != mn[which])
                sc = __ddiv_rn(__ll2double_rn(mx[which] - s_q[which * MPAD + m]),
                               __ll2double_rn(mx[which] - mn[which]));        // queue.go:99
              term = __dmul_rn(clamp01(sc), plan.weight[s]);
              break;
            }
            default: term = __dmul_rn(0.0, plan.weight[s]); break;  // pair columns absent: score 0
          }
          acc = __dadd_rn(acc, term);  // += enforceScoreRange(score) * weight, scheduler_profile.go:168
        }
        if (a.scores_out && m < M) a.scores_out[(size_t)r * M + m] = cand ? acc : __longlong_as_double(0x7ff8000000000000LL);
        if (cand) best_update(best, acc, m, tie_mode, areq, plan.seed_hi);
      }
    }
    best_warp_reduce(best, tie_mode);
    if (lane == 0) {
      a.pick[r] = best.m;
      a.pick_score[r] = best.m >= 0 ? best.score : 0.0;
      a.tie_count[r] = best.cnt;
      if (a.total_out) a.total_out[r] = (uint16_t)total;
    }
  }
}

// Dense rows: float4 {matchBlocks, lora class, pair0, pair1} per (request, endpoint), streamed with
// coalesced 128-bit loads (lane l reads endpoint i*32+l); endpoint tile in shared memory.
template <bool MASKED>
__global__ void __launch_bounds__(kScoreWarps * 32) score_pick_dense_kernel(const __grid_constant__ ScoreArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Plan& plan = a.plan;
  const int M = a.geo.M;
  const int MP = (M + 31) & ~31;
  double* s_term = reinterpret_cast<double*>(smem_raw);
  long long* s_q = reinterpret_cast<long long*>(s_term + (size_t)plan.n_terms * MP);
  double* s_lut = reinterpret_cast<double*>(s_q + (MASKED ? 2 * MP : 0));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = 0; t < plan.n_terms; t++)
    for (int m = threadIdx.x; m < MP; m += blockDim.x) s_term[(size_t)t * MP + m] = m < M ? a.term[t][m] : 0.0;
  if (MASKED)
    for (int which = 0; which < 2; which++)
      for (int m = threadIdx.x; m < MP; m += blockDim.x)
        s_q[which * MP + m] = (a.minmax_q[which] && m < M) ? a.minmax_q[which][m] : 0;
  __syncthreads();

  double* lut = s_lut + warp * (kLutMax + 1);
  int lut_total = -1, prefix_step = -1;
  bool has_minmax[2] = {false, false};
  for (int s = 0; s < plan.n_steps; s++) {
    if (plan.kind[s] == STEP_PREFIX && prefix_step < 0) prefix_step = s;
    if (plan.kind[s] == STEP_MINMAX) has_minmax[plan.arg[s]] = true;
  }
  const int tie_mode = plan.tie_mode;
  const int nchunks = MP >> 5;

  const int gw = blockIdx.x * kScoreWarps + warp, nw = gridDim.x * kScoreWarps;
  for (int r = gw; r < a.R; r += nw) {
    const int total = a.dense_total ? a.dense_total[r] : 0;
    if (prefix_step >= 0 && total != lut_total) {
      const double w = plan.weight[prefix_step];
      const int top = total < kLutMax ? total : kLutMax;
      __syncwarp();
      for (int c = lane; c <= top; c += 32) lut[c] = prefix_term_direct(c, total, w);
      lut_total = total;
      __syncwarp();
    }
    const uint32_t* mrow = MASKED ? a.cand_mask + (size_t)r * a.mask_words : nullptr;
    long long mn[2] = {0, 0}, mx[2] = {0, 0};
    if (MASKED && (has_minmax[0] || has_minmax[1])) {
      mn[0] = mn[1] = 0x7fffffffffffffffLL;
      mx[0] = mx[1] = (long long)0x8000000000000000ULL;
      for (int i = 0; i < nchunks; i++) {
        const int m = i * 32 + lane;
        if (((__ldg(mrow + i) >> lane) & 1u) && m < M) {
#pragma unroll
          for (int which = 0; which < 2; which++) {
            const long long v = s_q[which * MP + m];
            mn[which] = v < mn[which] ? v : mn[which];
            mx[which] = v > mx[which] ? v : mx[which];
          }
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int which = 0; which < 2; which++) {
          const long long omn = shfl_xor_i64(mn[which], o), omx = shfl_xor_i64(mx[which], o);
          mn[which] = omn < mn[which] ? omn : mn[which];
          mx[which] = omx > mx[which] ? omx : mx[which];
        }
    }
    const uint32_t areq = lowbias32((uint32_t)(uint64_t)(a.request_base + r) ^ plan.seed_lo);
    Best best;
    best.score = 0.0;
    best.m = -1;
    best.cnt = 0;
    best.prio = 0;
    const float4* row = a.dense + (size_t)r * M;
#pragma unroll 4
    for (int i = 0; i < nchunks; i++) {
      const int m = i * 32 + lane;
      bool cand = m < M;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cand) f = __ldg(row + m);
      if (MASKED) cand = cand && ((__ldg(mrow + i) >> lane) & 1u);
      const int c = (int)(__float2uint_rz(f.x) & 0xFFFFu);
      const int cls = __float2int_rz(f.y) & 3;
      if (a.match_out && m < M) a.match_out[(size_t)r * M + m] = (uint16_t)c;
      double acc = 0.0;
      for (int s = 0; s < plan.n_steps; s++) {
        double term;
        switch (plan.kind[s]) {
          case STEP_EP_TERM: term = s_term[(size_t)plan.arg[s] * MP + m]; break;
          case STEP_PREFIX: {
            const int cc = c < 65535 ? c : 65535;
            term = (s == prefix_step && total <= kLutMax && cc <= total) ? lut[cc]
                                                                          : prefix_term_direct(cc, total, plan.weight[s]);
            break;
          }
          case STEP_LORA: term = plan.lora_term[s][cls]; break;
          case STEP_PAIR:
            term = __dmul_rn(clamp01((double)(plan.arg[s] == 0 ? f.z : f.w)), plan.weight[s]);
            break;
          case STEP_MINMAX: {
            const int which = plan.arg[s];
            double sc = 1.0;
            if (MASKED && mx[which] != mn[which])
              sc = __ddiv_rn(__ll2double_rn(mx[which] - s_q[which * MP + m]), __ll2double_rn(mx[which] - mn[which]));
            term = __dmul_rn(clamp01(sc), plan.weight[s]);
            break;
          }
          default: term = 0.0; break;
        }
        acc = __dadd_rn(acc, term);
      }
      if (a.scores_out && m < M) a.scores_out[(size_t)r * M + m] = cand ? acc : __longlong_as_double(0x7ff8000000000000LL);
      if (cand) best_update(best, acc, m, tie_mode, areq, plan.seed_hi);
    }
    best_warp_reduce(best, tie_mode);
    if (lane == 0) {
      a.pick[r] = best.m;
      a.pick_score[r] = best.m >= 0 ? best.score : 0.0;
      a.tie_count[r] = best.cnt;
      if (a.total_out) a.total_out[r] = (uint16_t)total;
    }
  }
}

template <typename K>
static int launch_with_smem(K kernel, const ScoreArgs& a, size_t smem, cudaStream_t s, int sm_count) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kScoreWarps * 32, smem);
  if (occ < 1) occ = 1;
  const int need = (a.R + kScoreWarps - 1) / kScoreWarps;
  int blocks = sm_count * occ;  // persistent: one wave, warps stride over requests
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  kernel<<<blocks, kScoreWarps * 32, smem, s>>>(a);
  return 1;
}

template <int LOG_EPL, int J>
static int launch_fused_geo(const ScoreArgs& a, int np_class, bool masked, cudaStream_t s, int sm_count) {
  constexpr int MPAD = J * 32 * (1 << LOG_EPL);
  size_t smem = (size_t)a.plan.n_terms * MPAD * 8 + (masked ? 2 * (size_t)MPAD * 8 : 0) +
                (size_t)kScoreWarps * (kLutMax + 1) * 8;
  if (np_class == 0) {
    if (masked) return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 6, true>, a, smem, s, sm_count);
    return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 6, false>, a, smem, s, sm_count);
  }
  if (np_class == 1) {
    if (masked) return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 9, true>, a, smem, s, sm_count);
    return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 9, false>, a, smem, s, sm_count);
  }
  if (masked) return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 16, true>, a, smem, s, sm_count);
  return launch_with_smem(score_pick_fused_kernel<LOG_EPL, J, 16, false>, a, smem, s, sm_count);
}

int launch_score_pick(const ScoreArgs& a, bool dense, cudaStream_t s, int sm_count) {
  if (a.R <= 0) return 0;
  const bool masked = a.cand_mask != nullptr;
  if (dense) {
    const int MP = (a.geo.M + 31) & ~31;
    size_t smem = (size_t)a.plan.n_terms * MP * 8 + (masked ? 2 * (size_t)MP * 8 : 0) +
                  (size_t)kScoreWarps * (kLutMax + 1) * 8;
    if (masked) return launch_with_smem(score_pick_dense_kernel<true>, a, smem, s, sm_count);
    return launch_with_smem(score_pick_dense_kernel<false>, a, smem, s, sm_count);
  }
  // counter planes must hold counts up to the largest possible number of hashes per request
  const int maxn = a.hashes ? a.hash_stride : 0;
  const int np_class = maxn <= 63 ? 0 : (maxn <= 511 ? 1 : 2);
  const Geo& g = a.geo;
  if (g.log_epl == 3) return launch_fused_geo<3, 1>(a, np_class, masked, s, sm_count);
  if (g.log_epl == 4) return launch_fused_geo<4, 1>(a, np_class, masked, s, sm_count);
  switch (g.J) {
    case 1: return launch_fused_geo<5, 1>(a, np_class, masked, s, sm_count);
    case 2: return launch_fused_geo<5, 2>(a, np_class, masked, s, sm_count);
    case 4: return launch_fused_geo<5, 4>(a, np_class, masked, s, sm_count);
    default: return launch_fused_geo<5, 8>(a, np_class, masked, s, sm_count);
  }
}

// ---------------------------------------------------------------------------------------------
// prefix-table maintenance: the host mirror is authoritative; these scatter its dirty words/slots.
// ---------------------------------------------------------------------------------------------
__global__ void scatter_u32_kernel(uint32_t* dst, const uint32_t* idx, const uint32_t* val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
__global__ void scatter_slots_kernel(Slot* dst, const uint32_t* idx, const Slot* val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
int launch_scatter_u32(uint32_t* dst, const uint32_t* idx, const uint32_t* val, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  scatter_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, idx, val, n);
  return 1;
}
int launch_scatter_slots(Slot* dst, const uint32_t* idx, const Slot* val, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  scatter_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, idx, val, n);
  return 1;
}

}  // namespace eppscore
