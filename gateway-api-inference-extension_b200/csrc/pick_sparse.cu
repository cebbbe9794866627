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
// Work split: a group of G = row_words/4 lanes serves one request, so a warp serves 32/G requests at once; per-request
// match counters (natural endpoint order) and the bitmap of endpoints that have one live in shared memory.  The table is
// the device-resident index of prefix_table.cuh: a hit brings its endpoint set along in the 32-byte slot.
#include "device_common.cuh"
#include "prefix_table.cuh"

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
  constexpr int RW = J * 32;                       // words per PERMUTED bit row (LoRA class planes, tie masks)
  constexpr int G = (RW / 4 < 32) ? RW / 4 : 32;   // lanes per request
  constexpr int PW = RW / G;                       // permuted words per lane
  constexpr int RPW = 32 / G;                      // requests per warp
  const int MPAD = a.geo.Mpad;
  const int NW = MPAD >> 5;                        // words of a NATURAL-order bitmap (overflow rows, touched bitmap)
  const int WPL = NW / G;                          // ... per lane (>= 1 for every geometry of make_geo)
  // per request in shared memory: match counters (natural endpoint order) + the bitmap of endpoints with a match
  const int REQ_BYTES = MPAD * (int)sizeof(CNT) + NW * 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Plan& plan = a.plan;
  const int M = a.geo.M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane / G, gl = lane % G;
  unsigned char* mine = smem_raw + (size_t)(warp * RPW + gi) * REQ_BYTES;
  CNT* cnt = reinterpret_cast<CNT*>(mine);
  uint32_t* touched = reinterpret_cast<uint32_t*>(mine + MPAD * sizeof(CNT));

  // zero everything once; afterwards each request re-zeroes exactly what it touched
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw);
    const int nwords = kSparseWarps * RPW * REQ_BYTES / 4;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) z[i] = 0;
  }
  __syncthreads();

  int prefix_step = -1;
  for (int s = 0; s < plan.n_steps; s++)
    if (plan.kind[s] == STEP_PREFIX) prefix_step = s;
  const int tie_mode = plan.tie_mode;
  const bool have_table = a.table != nullptr && a.hashes != nullptr && prefix_step >= 0;
  const TSlot* slots = nullptr;
  uint64_t slot_mask = 0;
  const uint32_t* ovf_rows = nullptr;
  if (have_table) {
    slots = a.table->slots;
    slot_mask = a.table->mask;
    ovf_rows = a.table->ovf_rows;
  }

  const int wstride = gridDim.x * kSparseWarps * RPW;
  for (int rbase = (blockIdx.x * kSparseWarps + warp) * RPW; rbase < a.R; rbase += wstride) {
    const int r = rbase + gi;
    const bool valid = r < a.R;
    const int n = (valid && a.hashes) ? (int)a.n_hashes[r] : 0;
    int ad = (valid && a.adapter_id) ? a.adapter_id[r] : -1;
    if (ad < 0 || ad >= a.A) ad = a.A;
    const AdapterSummary sm = a.summ[ad];  // issued early: only consumed after the probe phase

    // ---------------- matchLongestPrefix: probe U*G hashes per round, stop at the first global miss ----------------
    // A hit returns the endpoint set with the slot itself (up to 8 members inline).  Consecutive blocks of a prompt are
    // normally cached on the same endpoints: hits whose slots carry identical inline sets are run-length merged, lane k of
    // the request's group bumps member k's counter by the run length.
    if (have_table) {
      constexpr int U = 4;            // hashes probed per lane per round (independent loads in flight)
      bool stop = n == 0;
      int c0 = 0;
      while (__any_sync(0xffffffffu, !stop)) {
        uint64_t h[U], idx[U];
        uint4 lo[U];   // {key.lo, key.hi, cnt, ovf}; the members (second half of the slot, same sector) are loaded per run
        bool act[U], hit[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int i = c0 + u * G + gl;
          act[u] = !stop && i < n;
          h[u] = act[u] ? a.hashes[(size_t)r * a.hash_stride + i] : 0ULL;
          idx[u] = h[u] & slot_mask;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
          if (act[u]) lo[u] = ldg16(&slots[idx[u]]);
#pragma unroll
        for (int u = 0; u < U; u++) {                           // indexer.Get, indexer.go:86-102
          hit[u] = false;
          if (act[u]) {
            for (;;) {
              if (lo[u].z == kCntFree) break;                   // never-used slot: hash unknown
              if ((((uint64_t)lo[u].y << 32) | lo[u].x) == h[u]) {
                hit[u] = lo[u].z != 0;                          // emptied set == deleted key
                break;
              }
              idx[u] = (idx[u] + 1) & slot_mask;
              lo[u] = ldg16(&slots[idx[u]]);
            }
          }
        }
        int nh_total = 0;   // blocks matched in this round before the first global miss
        bool open = !stop;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t miss = __ballot_sync(0xffffffffu, !hit[u]);
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
          if (__all_sync(0xffffffffu, nh_u == 0)) continue;
          // the members: second 16 bytes of the slot (the sector is already in L1)
          uint4 hi = make_uint4(0, 0, 0, 0);
          if (gl < nh_u && lo[u].w == kNoRow) hi = ldg16(reinterpret_cast<const uint4*>(&slots[idx[u]]) + 1);
          // run boundaries: a hit starts a new run unless its inline set equals the previous hit's, member for member
          const uint32_t pc = __shfl_up_sync(0xffffffffu, lo[u].z, 1, G), po = __shfl_up_sync(0xffffffffu, lo[u].w, 1, G);
          const uint32_t p0 = __shfl_up_sync(0xffffffffu, hi.x, 1, G), p1 = __shfl_up_sync(0xffffffffu, hi.y, 1, G);
          const uint32_t p2 = __shfl_up_sync(0xffffffffu, hi.z, 1, G), p3 = __shfl_up_sync(0xffffffffu, hi.w, 1, G);
          const uint32_t c = lo[u].z;
          bool same = gl > 0 && lo[u].w == kNoRow && po == kNoRow && pc == c;
          if (same) {  // compare the first c members (the tail of ep[] is unspecified)
            const uint32_t m1 = c >= 2 ? 0xFFFFFFFFu : 0x0000FFFFu;
            same = ((p0 ^ hi.x) & m1) == 0;
            if (c > 2) same = same && ((p1 ^ hi.y) & (c >= 4 ? 0xFFFFFFFFu : 0x0000FFFFu)) == 0;
            if (c > 4) same = same && ((p2 ^ hi.z) & (c >= 6 ? 0xFFFFFFFFu : 0x0000FFFFu)) == 0;
            if (c > 6) same = same && ((p3 ^ hi.w) & (c >= 8 ? 0xFFFFFFFFu : 0x0000FFFFu)) == 0;
          }
          const bool boundary = gl < nh_u && !same;
          const uint32_t ball = __ballot_sync(0xffffffffu, boundary);
          uint32_t bm = (G == 32) ? ball : ((ball >> (gi * G)) & ((1u << G) - 1u));
          while (__any_sync(0xffffffffu, bm != 0)) {
            int s0 = 0, len = 0;
            if (bm) {
              s0 = __ffs(bm) - 1;
              bm &= bm - 1;
              len = (bm ? (__ffs(bm) - 1) : nh_u) - s0;
            }
            const int src = gi * G + s0;
            const uint32_t rc = __shfl_sync(0xffffffffu, lo[u].z, src), ro = __shfl_sync(0xffffffffu, lo[u].w, src);
            const uint32_t e0 = __shfl_sync(0xffffffffu, hi.x, src), e1 = __shfl_sync(0xffffffffu, hi.y, src);
            const uint32_t e2 = __shfl_sync(0xffffffffu, hi.z, src), e3 = __shfl_sync(0xffffffffu, hi.w, src);
            if (len > 0) {
              if (ro == kNoRow) {                                 // inline set: lane k owns member k
                if (gl < (int)rc && gl < kInlineEps) {
                  const uint32_t wsel = (gl >> 1) == 0 ? e0 : ((gl >> 1) == 1 ? e1 : ((gl >> 1) == 2 ? e2 : e3));
                  const int m = (int)((gl & 1) ? (wsel >> 16) : (wsel & 0xFFFFu));
                  const CNT old = cnt[m];                         // res[server] += len
                  cnt[m] = (CNT)(old + (CNT)len);
                  if (old == 0) atomicOr(&touched[m >> 5], 1u << (m & 31));
                }
              } else {                                            // a set of more than 8: one bitset row, natural order
                for (int q = 0; q < WPL; q++) {
                  const int w = gl * WPL + q;
                  uint32_t x = __ldg(ovf_rows + (size_t)ro * NW + w);
                  touched[w] |= x;                                // this lane owns word w of the bitmap in this step
                  while (x) {
                    const int k = __ffs(x) - 1;
                    x &= x - 1;
                    cnt[w * 32 + k] += (CNT)len;
                  }
                }
              }
            }
            __syncwarp();  // members of the next run may alias this run's counters
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
    for (int q = 0; q < WPL; q++) {
      const int w = gl * WPL + q;
      uint32_t x = touched[w];
      while (x) {                                                 // ascending k == ascending m for this lane
        const int k = __ffs(x) - 1;
        x &= x - 1;
        const int m = w * 32 + k;
        int c = (int)cnt[m];   // stored modulo 2^bits(CNT); a touched slot holds c >= 1, so 0 means 2^bits
        if (c == 0) c = 1 << (8 * (int)sizeof(CNT));
        cnt[m] = 0;
        if (m < M) {
          const uint32_t pos = perm_bitpos((uint32_t)m, LOG_EPL);
          const int wi = (int)(pos >> 5), kb = (int)(pos & 31);
          const uint32_t clo = __ldg(a.cls_lo + (size_t)ad * RW + wi), chi = __ldg(a.cls_hi + (size_t)ad * RW + wi);
          const uint32_t tmw = __ldg(a.tiemask + (size_t)ad * RW + wi);  // bit kb: G[ad][m] == gmax
          const int cls = (int)((clo >> kb) & 1u) | ((int)((chi >> kb) & 1u) << 1);
          const double s_true = eval_exception<SEQ>(a, m, c, n, cls);
          best_update(best, s_true, m, tie_mode, areq, plan.seed_hi);
          xg += (int)((tmw >> kb) & 1u);
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
        for (int q = 0; q < PW; q++) {
          const int wi = gl * PW + q;              // a word of the PERMUTED tie mask
          uint32_t x = __ldg(a.tiemask + (size_t)ad * RW + wi);
          const int j = wi >> 5, ln = wi & 31;
          while (x) {
            const int k = __ffs(x) - 1;
            x &= x - 1;
            const int m = (((j << LOG_EPL) + k) << 5) + ln;
            if (m < M && !((touched[m >> 5] >> (m & 31)) & 1u)) best_update(b2, sm.gmax, m, tie_mode, areq, plan.seed_hi);
          }
        }
      }
      __syncwarp();
      best_group_reduce<G>(b2, tie_mode);
      if (need) pick = b2.m;
    }
    __syncwarp();
    for (int q = 0; q < WPL; q++) touched[gl * WPL + q] = 0;
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
  const size_t smem = (size_t)kSparseWarps * RPW * ((size_t)a.geo.Mpad * sizeof(CNT) + (size_t)(a.geo.Mpad >> 5) * 4);
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
  if (a.geo.Mpad > 65535 + 1) return 0;  // inline members are uint16
  switch (a.geo.J) {
    case 1: return launch_sparse_geo<1>(a, s, sm_count);
    case 2: return launch_sparse_geo<2>(a, s, sm_count);
    case 4: return launch_sparse_geo<4>(a, s, sm_count);
    default: return launch_sparse_geo<8>(a, s, sm_count);
  }
}

}  // namespace eppscore
