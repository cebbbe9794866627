// goshape.cpp — "Go-shape restatement" of the reference CPU path (baseline form (i)).
//
// TEST / BENCH INFRASTRUCTURE ONLY, like everything under oracle/: only tests/ and bench.py's CPU-baseline legs
// load it.  It computes the same results as oracle.c but with the reference's *data structures and allocation
// pattern*, so that its timing resembles what the Go scheduler pays per request (it is labelled "Go-shape", never
// "Go"): a per-request deep clone of every candidate's metrics (director.go:342-349), one hash map per scorer
// keyed by endpoint (scorer Score() returns map[Endpoint]float64, interface/scheduling/plugins.go:68-72), the
// accumulate map of runScorerPlugins (scheduler_profile.go:151-174), matchLongestPrefix into a map
// (approximateprefix/plugin.go:219-235), and the max-score picker's shuffle + stable sort
// (picker/maxscore/picker.go:87-115, picker/common.go:49-55).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <thread>
#include <unordered_map>
#include <vector>

#include "oracle.h"

namespace {

struct Metrics {  // fwkdl.Metrics (interface/datalayer/metrics.go:26-42)
  std::unordered_map<int32_t, int> ActiveModels, WaitingModels;
  int MaxActiveModels = 0;
  int64_t RunningRequestsSize = 0, WaitingQueueSize = 0;
  double KVCacheUsagePercent = 0.0;
};
struct Endpoint {
  int32_t index;
  std::unique_ptr<Metrics> metrics;             // cloned per request
  std::unique_ptr<std::pair<int, int>> prefix;  // PrefixCacheMatchInfo{matchBlocks, totalBlocks}
};
using ScoreMap = std::unordered_map<const Endpoint*, double>;

std::vector<Metrics> build_pool(const orc_snapshot* s) {
  std::vector<Metrics> pool((size_t)s->M);
  for (int32_t m = 0; m < s->M; m++) {
    Metrics& x = pool[(size_t)m];
    x.KVCacheUsagePercent = s->kv_usage[m];
    x.WaitingQueueSize = s->queue[m];
    x.RunningRequestsSize = s->running ? s->running[m] : 0;
    x.MaxActiveModels = s->lora_max ? s->lora_max[m] : 0;
    for (int32_t a = 0; a < s->lora_words * 64; a++) {
      if (s->lora_active && ((s->lora_active[(size_t)m * s->lora_words + (a >> 6)] >> (a & 63)) & 1)) x.ActiveModels[a] = 1;
      if (s->lora_waiting && ((s->lora_waiting[(size_t)m * s->lora_words + (a >> 6)] >> (a & 63)) & 1)) x.WaitingModels[a] = 1;
    }
    // len(ActiveModels)+len(WaitingModels) may exceed the dictionary (out-of-vocabulary adapters): pad with ids >= A
    int want = s->lora_nmodels ? s->lora_nmodels[m] : 0;
    for (int32_t k = 0; (int)(x.ActiveModels.size() + x.WaitingModels.size()) < want; k++) x.WaitingModels[s->lora_words * 64 + k] = 1;
  }
  return pool;
}

double clamp01(double v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }

void schedule_range(const orc_snapshot* s, const orc_profile* p, const orc_index* idx, const orc_batch* b,
                    const std::vector<Metrics>* pool, int32_t r0, int32_t r1, uint64_t rng_seed) {
  std::mt19937_64 rng(rng_seed);
  const int32_t M = s->M;
  std::vector<uint64_t> hashes((size_t)(b->max_blocks > 0 ? b->max_blocks : 1));
  std::vector<int32_t> servers(4096);
  for (int32_t r = r0; r < r1; r++) {
    // per-request deep clone of the candidates (director.go:342-349)
    std::vector<Endpoint> eps((size_t)M);
    for (int32_t m = 0; m < M; m++) {
      eps[(size_t)m].index = m;
      eps[(size_t)m].metrics = std::make_unique<Metrics>((*pool)[(size_t)m]);
    }
    // PrepareRequestData: hashPrompt + matchLongestPrefix into a map, then one attribute per endpoint
    bool want_prefix = false;
    for (int i = 0; i < p->n_scorers; i++) want_prefix |= p->scorer_kind[i] == ORC_SCORER_PREFIX;
    if (want_prefix && b->prompt_bytes) {
      const int32_t nh = orc_hash_prompt(b->prompt_bytes + b->prompt_off[r], b->prompt_off[r + 1] - b->prompt_off[r],
                                         b->model_seed ? b->model_seed[r] : 0, b->block_chars, b->max_blocks, hashes.data(),
                                         (int32_t)hashes.size());
      std::unordered_map<int32_t, int> res;
      for (int32_t i = 0; i < nh && idx; i++) {
        int32_t n = orc_index_get(idx, hashes[(size_t)i], servers.data(), (int32_t)servers.size());
        if (n == 0) break;
        for (int32_t k = 0; k < n && k < (int32_t)servers.size(); k++) res[servers[(size_t)k]]++;
      }
      for (auto& e : eps) {
        auto it = res.find(e.index);
        e.prefix = std::make_unique<std::pair<int, int>>(it == res.end() ? 0 : it->second, nh);
      }
    }
    const int32_t adapter = b->adapter_id ? b->adapter_id[r] : -1;
    // runScorerPlugins (scheduler_profile.go:151-174)
    ScoreMap weighted;
    for (auto& e : eps) weighted[&e] = 0.0;
    for (int k = 0; k < p->n_scorers; k++) {
      ScoreMap scores;
      switch (p->scorer_kind[k]) {
        case ORC_SCORER_KV_CACHE:
          for (auto& e : eps) scores[&e] = 1 - e.metrics->KVCacheUsagePercent;
          break;
        case ORC_SCORER_QUEUE:
        case ORC_SCORER_RUNNING: {
          const bool q = p->scorer_kind[k] == ORC_SCORER_QUEUE;
          int64_t mn = INT64_MAX, mx = INT64_MIN;
          for (auto& e : eps) {
            const int64_t v = q ? e.metrics->WaitingQueueSize : e.metrics->RunningRequestsSize;
            mn = std::min(mn, v);
            mx = std::max(mx, v);
          }
          for (auto& e : eps) {
            const int64_t v = q ? e.metrics->WaitingQueueSize : e.metrics->RunningRequestsSize;
            scores[&e] = mx == mn ? 1.0 : (double)(mx - v) / (double)(mx - mn);
          }
          break;
        }
        case ORC_SCORER_PREFIX:
          for (auto& e : eps) scores[&e] = (e.prefix && e.prefix->second != 0) ? (double)e.prefix->first / (double)e.prefix->second : 0.0;
          break;
        case ORC_SCORER_LORA:
          for (auto& e : eps) {
            const Metrics& x = *e.metrics;
            double sc;
            if (x.ActiveModels.count(adapter)) sc = 1.0;
            else if ((int)(x.ActiveModels.size() + x.WaitingModels.size()) < x.MaxActiveModels) sc = 0.8;
            else if (x.WaitingModels.count(adapter)) sc = 0.6;
            else sc = 0.0;
            scores[&e] = sc;
          }
          break;
        default:
          for (auto& e : eps) scores[&e] = 0.0;
      }
      const double w = p->scorer_weight[k];
      for (auto& e : eps) {
        const double t = clamp01(scores[&e]) * w;
        weighted[&e] = weighted[&e] + t;
      }
    }
    // MaxScorePicker: shuffle, stable sort by score descending, first
    std::vector<std::pair<const Endpoint*, double>> scored;
    scored.reserve((size_t)M);
    for (auto& e : eps) scored.emplace_back(&e, weighted[&e]);
    std::shuffle(scored.begin(), scored.end(), rng);
    std::stable_sort(scored.begin(), scored.end(), [](const auto& a, const auto& c) { return a.second > c.second; });
    if (scored.empty()) {
      b->pick[r] = -1;
      b->pick_score[r] = 0.0;
      b->tie_count[r] = 0;
    } else {
      b->pick[r] = scored[0].first->index;
      b->pick_score[r] = scored[0].second;
      int32_t ties = 0;
      for (auto& se : scored) ties += se.second == scored[0].second;
      b->tie_count[r] = ties;
    }
  }
}

}  // namespace

extern "C" int32_t orc_goshape_schedule_batch(const orc_snapshot* s, const orc_profile* p, const orc_index* idx,
                                              const orc_batch* b, int32_t n_threads, uint64_t shuffle_seed) {
  if (b->cand_mask || b->dense_feat || b->hashes_in) return -1;  // the timed baseline covers the default (unfiltered) path
  const std::vector<Metrics> pool = build_pool(s);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > b->R) n_threads = b->R > 0 ? b->R : 1;
  std::vector<std::thread> th;
  for (int32_t t = 0; t < n_threads; t++) {
    const int32_t r0 = (int32_t)((int64_t)b->R * t / n_threads), r1 = (int32_t)((int64_t)b->R * (t + 1) / n_threads);
    th.emplace_back(schedule_range, s, p, idx, b, &pool, r0, r1, shuffle_seed + (uint64_t)t);
  }
  for (auto& x : th) x.join();
  return 0;
}
