// pick_sparse.cu — the exact sparse fast path of Score+Pick (unmasked batches, E/P/L steps).
//
// For a request r with adapter a, the weighted score of endpoint m is
//     S(r,m) = steps(base terms[m], prefix term(c(r,m)), lora term(class(a,m)))   (scheduler_profile.go:155-168)
// and c(r,m) — the matched prefix blocks (approximateprefix/plugin.go:219-235) — is ZERO for all but the
// few endpoints that appear in the rows of the request's matched block hashes.  With c = 0 the score is
// G[a][m], which depends only on the snapshot and the adapter; prepare_adapters_kernel has already
// reduced it to (gmax, garg, gcnt, tie mask) per adapter.  Because the prefix weight is >= 0 and float64
// addition is monotone, S(r,e) >= G[a][e] for every "exception" e (c > 0), hence
//     max_m S(r,m) = max( gmax[a], max_e S(r,e) )
// and the arg-max set follows from three cases (T = max_e S(r,e), xg = #{e : G[a][e] == gmax}):
//     T > gmax : only exceptions attain it;
//     T < gmax : the precomputed arg-max set, untouched by any exception (xg == 0 by monotonicity);
//     T == gmax: (precomputed set minus its xg exception members) ∪ {e : S(r,e) == gmax}.
// So a request costs O(matched blocks + exceptions) instead of O(M) — bit-identical picks, scores and
// tie counts (tests compare against the oracle and against the generic kernel).
//
// Work split: a group of G = row_words/4 lanes serves one request (each lane owns 16 bytes of every
// bitset row), so a warp serves 32/G requests at once; per-request match counters live in shared memory.
#include "device_common.cuh"

namespace eppscore {

constexpr int kSparseWarps = 8;

// Score of an exception endpoint with the scorer sequence known at compile time (SEQ packs kind+1 per step,
// 4 bits each; SEQ == 0 selects the runtime-generic eval_steps).  The prefix term comes from the engine-wide
// table lut2d[total][c] = clamp(c/total)*w (built once on the host: one IEEE divide, one multiply per entry).
template <uint32_t SEQ>
__device__ __forceinline__ double eval_exception(const ScoreArgs& a, int m, int c, int total, int cls) {
  if (SEQ == 0) return eval_steps(a.plan, a.term, m, c, total, cls);
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int kind = (int)((SEQ >> (4 * s)) & 15u) - 1;
    if (kind < 0) break;
    double t;
    if (kind == STEP_EP_TERM) {
      t = __ldg(a.term[a.plan.arg[s]] + m);
    } else if (kind == STEP_PREFIX) {
      t = (a.prefix_lut2d && total <= kLutMax) ? __ldg(a.prefix_lut2d + total * (kLutMax + 1) + c)
                                               : prefix_term_direct(c, total, a.plan.weight[s]);
    } else {  // STEP_LORA
      const double* lt = a.plan.lora_term[s];
      t = cls == 3 ? lt[3] : (cls == 2 ? lt[2] : (cls == 1 ? lt[1] : lt[0]));
    }
    acc = __dadd_rn(acc, t);
  }
  return acc;
}

template <int J, typename CNT, uint32_t SEQ>
__global__ void __launch_bounds__(kSparseWarps * 32, 4) pick_sparse_kernel(const __grid_constant__ ScoreArgs a) {
  const int LOG_EPL = a.geo.log_epl;
  constexpr int RW = J * 32;                       // words per bitset row
  constexpr int G = (RW / 4 < 32) ? RW / 4 : 32;   // lanes per request
  constexpr int QW = RW / (4 * G);                 // 16-byte quads per lane per row
  constexpr int RPW = 32 / G;                      // requests per warp
  constexpr int NPOS = RW * 32;                    // counter slots per request (indexed by permuted bit position)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Plan& plan = a.plan;
  const int M = a.geo.M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane / G, gl = lane % G;
  CNT* cnt = reinterpret_cast<CNT*>(smem_raw) + (size_t)(warp * RPW + gi) * NPOS;

  // zero all counters once; afterwards each request re-zeroes exactly the slots it touched
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw);
    const int nwords = (int)((size_t)kSparseWarps * RPW * NPOS * sizeof(CNT) / 4);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) z[i] = 0;
  }
  __syncthreads();

  int prefix_step = -1;
  for (int s = 0; s < plan.n_steps; s++)
    if (plan.kind[s] == STEP_PREFIX) prefix_step = s;
  const int tie_mode = plan.tie_mode;
  const bool have_table = a.slots != nullptr && a.hashes != nullptr && prefix_step >= 0;

  const int wstride = gridDim.x * kSparseWarps * RPW;
  for (int rbase = (blockIdx.x * kSparseWarps + warp) * RPW; rbase < a.R; rbase += wstride) {
    const int r = rbase + gi;
    const bool valid = r < a.R;
    const int n = (valid && a.hashes) ? (int)a.n_hashes[r] : 0;
    int ad = (valid && a.adapter_id) ? a.adapter_id[r] : -1;
    if (ad < 0 || ad >= a.A) ad = a.A;
    const AdapterSummary sm = a.summ[ad];  // issued early: only consumed after the probe / row phase
    uint32_t any[QW][4];
#pragma unroll
    for (int q = 0; q < QW; q++) any[q][0] = any[q][1] = any[q][2] = any[q][3] = 0;

    // ---------------- matchLongestPrefix: probe U*G hashes per round, stop at the first global miss ----------------
    // Hashes whose endpoint sets are identical share ONE interned row (prefix_index.hpp), and consecutive blocks
    // of a prompt are normally cached on the same endpoints: the hits are run-length merged by row id and each
    // distinct set is read once, its members' counters bumped by the run length.
    if (have_table) {
      constexpr int U = 4;            // hashes probed per lane per round (independent loads in flight)
      bool stop = n == 0;
      int c0 = 0;
      while (__any_sync(0xffffffffu, !stop)) {
        uint64_t h[U], idx[U];
        uint4 sv[U];
        uint32_t row[U];
        bool act[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int i = c0 + u * G + gl;
          act[u] = !stop && i < n;
          h[u] = act[u] ? a.hashes[(size_t)r * a.hash_stride + i] : 0ULL;
          idx[u] = h[u] & a.slot_mask;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
          if (act[u]) sv[u] = ldg16(&a.slots[idx[u]]);
#pragma unroll
        for (int u = 0; u < U; u++) {                           // indexer.Get, indexer.go:86-102
          row[u] = kEmptyRow;
          if (act[u]) {
            for (;;) {
              if (sv[u].z == kEmptyRow) break;                  // never-used slot: hash unknown
              if ((((uint64_t)sv[u].y << 32) | sv[u].x) == h[u]) {
                if (sv[u].w != 0) row[u] = sv[u].z;             // emptied set == deleted key
                break;
              }
              idx[u] = (idx[u] + 1) & a.slot_mask;
              sv[u] = ldg16(&a.slots[idx[u]]);
            }
          }
        }
        int nh_total = 0;   // blocks matched in this round before the first global miss
        bool open = !stop;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t miss = __ballot_sync(0xffffffffu, row[u] == kEmptyRow);
          const uint32_t gmiss = (G == 32) ? miss : ((miss >> (gi * G)) & ((1u << G) - 1u));
          if (open) {
            const int nh_u = gmiss ? (__ffs(gmiss) - 1) : G;
            nh_total += nh_u;
            if (nh_u < G) open = false;
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          int nh_u = nh_total - u * G;
          nh_u = nh_u < 0 ? 0 : (nh_u > G ? G : nh_u);
          const uint32_t prev_rr = __shfl_up_sync(0xffffffffu, row[u], 1, G);
          const bool boundary = gl < nh_u && (gl == 0 || row[u] != prev_rr);
          const uint32_t ball = __ballot_sync(0xffffffffu, boundary);
          uint32_t bm = (G == 32) ? ball : ((ball >> (gi * G)) & ((1u << G) - 1u));
          while (__any_sync(0xffffffffu, bm != 0)) {
            int s0 = 0, len = 0;
            if (bm) {
              s0 = __ffs(bm) - 1;
              bm &= bm - 1;
              len = (bm ? (__ffs(bm) - 1) : nh_u) - s0;
            }
            const uint32_t rr = __shfl_sync(0xffffffffu, row[u], gi * G + s0);
            if (len > 0) {
#pragma unroll
              for (int q = 0; q < QW; q++) {
                const int w0 = (q * G + gl) * 4;
                const uint4 w = ldg16(a.rows + (size_t)rr * RW + w0);
                if (w.x | w.y | w.z | w.w) {                     // res[server] += len for every server in the set
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                  for (int t = 0; t < 4; t++) {
                    uint32_t x = ww[t];
                    any[q][t] |= x;
                    while (x) {
                      const int k = __ffs(x) - 1;
                      x &= x - 1;
                      cnt[(w0 + t) * 32 + k] += (CNT)len;
                    }
                  }
                }
              }
            }
          }
        }
        c0 += U * G;
        if (nh_total < U * G || c0 >= n) stop = true;
      }
    }

    // ---------------- exceptions: endpoints with a non-zero match count ----------------
    const uint32_t areq = tie_areq(a.request_base + r, plan.seed_lo);
    Best best = best_none();
    int xg = 0;
#pragma unroll
    for (int q = 0; q < QW; q++) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        uint32_t x = any[q][t];
        if (x) {
          const int wi = (q * G + gl) * 4 + t;
          const uint32_t clo = __ldg(a.cls_lo + (size_t)ad * RW + wi), chi = __ldg(a.cls_hi + (size_t)ad * RW + wi);
          const uint32_t tmw = __ldg(a.tiemask + (size_t)ad * RW + wi);  // bit k: G[ad][m] == gmax
          const int j = wi >> 5, ln = wi & 31;
          while (x) {                                             // ascending k == ascending m for this lane
            const int k = __ffs(x) - 1;
            x &= x - 1;
            int c = (int)cnt[wi * 32 + k];   // stored modulo 2^bits(CNT); touched slots hold c >= 1, so 0 means 2^bits
            if (c == 0) c = 1 << (8 * (int)sizeof(CNT));
            cnt[wi * 32 + k] = 0;
            const int m = (((j << LOG_EPL) + k) << 5) + ln;
            if (m < M) {
              const int cls = (int)((clo >> k) & 1u) | ((int)((chi >> k) & 1u) << 1);
              const double s_true = eval_exception<SEQ>(a, m, c, n, cls);
              best_update(best, s_true, m, tie_mode, areq, plan.seed_hi);
              xg += (int)((tmw >> k) & 1u);
            }
          }
        }
      }
    }
    // reduce over the request's lane group. Exception m are not globally ascending across lanes: merge rules
    // (lowest index / highest priority) are order independent.
    best_group_reduce<G>(best, tie_mode);
#pragma unroll
    for (int o = G / 2; o; o >>= 1) xg += __shfl_xor_sync(0xffffffffu, xg, o);

    // ---------------- combine with the per-adapter summary ----------------
    int pick, ties;
    double score;
    const bool exc_wins = best.m >= 0 && best.score > sm.gmax;
    const bool exc_ties = best.m >= 0 && best.score == sm.gmax;
    if (exc_wins) {
      pick = best.m;
      score = best.score;
      ties = best.cnt;
    } else {
      score = sm.gmax;
      ties = (sm.gcnt - xg) + (exc_ties ? best.cnt : 0);
      pick = exc_ties ? (best.m < sm.garg ? best.m : sm.garg) : sm.garg;
      if (sm.garg < 0) pick = exc_ties ? best.m : -1;
    }
    if (tie_mode) {  // warp-uniform: the shuffles below need every lane, whichever groups actually have a tie
      // seeded-random tie-break: the arg-max member with the highest priority, over
      // (precomputed tie set minus exceptions) ∪ (exceptions that tie)
      const bool need = !exc_wins && ties > 1;  // uniform within a request's lane group only
      Best b2 = best_none();
      if (need && exc_ties && gl == 0) b2 = best;  // already reduced over the group
      if (need) {
#pragma unroll
        for (int q = 0; q < QW; q++) {
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const int wi = (q * G + gl) * 4 + t;
            uint32_t x = __ldg(a.tiemask + (size_t)ad * RW + wi) & ~any[q][t];
            const int j = wi >> 5, ln = wi & 31;
            while (x) {
              const int k = __ffs(x) - 1;
              x &= x - 1;
              const int m = (((j << LOG_EPL) + k) << 5) + ln;
              if (m < M) best_update(b2, sm.gmax, m, tie_mode, areq, plan.seed_hi);
            }
          }
        }
      }
      __syncwarp();
      best_group_reduce<G>(b2, tie_mode);
      if (need) pick = b2.m;
    }
    // the warp-level shuffles above need every lane; only now drop the padding groups
    if (valid && gl == 0) {
      a.pick[r] = pick;
      a.pick_score[r] = pick >= 0 ? score : 0.0;
      a.tie_count[r] = pick >= 0 ? ties : 0;
      if (a.total_out) a.total_out[r] = (uint16_t)n;
    }
    __syncwarp();
  }
}

template <int J, typename CNT, uint32_t SEQ>
static int launch_sparse_inst(const ScoreArgs& a, cudaStream_t s, int sm_count) {
  constexpr int RW = J * 32;
  constexpr int G = (RW / 4 < 32) ? RW / 4 : 32;
  constexpr int RPW = 32 / G;
  auto kernel = pick_sparse_kernel<J, CNT, SEQ>;
  const size_t smem = (size_t)kSparseWarps * RPW * RW * 32 * sizeof(CNT);
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSparseWarps * 32, smem);
  if (occ < 1) occ = 1;
  const int per_block = kSparseWarps * RPW;
  int blocks = sm_count * occ;
  const int need = (a.R + per_block - 1) / per_block;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  kernel<<<blocks, kSparseWarps * 32, smem, s>>>(a);
  return 1;
}

constexpr uint32_t sparse_seq() { return 0; }
template <typename... Rest>
constexpr uint32_t sparse_seq(int k, Rest... rest) {
  return (uint32_t)(k + 1) | (sparse_seq(rest...) << 4);
}

template <int J, typename CNT>
static int launch_sparse_seq(const ScoreArgs& a, cudaStream_t s, int sm_count) {
  constexpr uint32_t EPL_ = sparse_seq(STEP_EP_TERM, STEP_PREFIX, STEP_LORA);  // queue,kv folded | prefix | lora
  constexpr uint32_t EP_ = sparse_seq(STEP_EP_TERM, STEP_PREFIX);              // the reference's default config
  if (a.plan.seq == EPL_) return launch_sparse_inst<J, CNT, EPL_>(a, s, sm_count);
  if (a.plan.seq == EP_) return launch_sparse_inst<J, CNT, EP_>(a, s, sm_count);
  return launch_sparse_inst<J, CNT, 0>(a, s, sm_count);
}

template <int J>
static int launch_sparse_geo(const ScoreArgs& a, cudaStream_t s, int sm_count) {
  const int maxn = a.hashes ? a.hash_stride : 0;
  if (maxn <= 256) return launch_sparse_seq<J, uint8_t>(a, s, sm_count);  // incl. defaultMaxPrefixBlocks
  return launch_sparse_seq<J, uint16_t>(a, s, sm_count);
}

// Applicable when: unmasked, not dense, no per-pair diagnostics, plan flagged sparse_ok, summaries present.
int launch_pick_sparse(const ScoreArgs& a, cudaStream_t s, int sm_count) {
  if (a.R <= 0) return 0;
  if (!a.plan.sparse_ok || !a.summ || !a.tiemask || a.cand_mask || a.dense || a.match_out || a.scores_out) return 0;
  switch (a.geo.J) {
    case 1: return launch_sparse_geo<1>(a, s, sm_count);
    case 2: return launch_sparse_geo<2>(a, s, sm_count);
    case 4: return launch_sparse_geo<4>(a, s, sm_count);
    default: return launch_sparse_geo<8>(a, s, sm_count);
  }
}

}  // namespace eppscore
