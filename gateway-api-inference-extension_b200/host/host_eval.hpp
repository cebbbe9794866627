// host_eval.hpp — host-side pieces of the path that need no engine (SURVEY §8 f4 and the small-batch route):
//   LoRA metric labels  populateLoRAMetrics / addAdapters (pkg/epp/framework/plugins/datalayer/extractor/metrics/
//                       extractor.go:216-236,266-272): the `running_lora_adapters` / `waiting_lora_adapters` CSV labels and
//                       `max_lora` of vllm:lora_requests_info become Metrics.ActiveModels / WaitingModels / MaxActiveModels;
//                       AdapterDictionary interns adapter names into the ids the engine's bitmasks are indexed by.
//   TopK                max-score-picker with maxNumOfEndpoints > 1 (picker/maxscore/picker.go:104-106) from the engine's
//                       scores_out row: shuffle-free, deterministic restatement (descending score, ties ascending index).
//   SmallBatchCpu       the route for batches too small to be worth a launch (SURVEY §7 "tiny batches"; BASELINE config A:
//                       1 request x 4 pods, queue scorer): request-independent profiles (queue / kv / running / lora /
//                       token-load) evaluated on the host with the reference's float64 operation order.  Product code, not
//                       the oracle; parity-tested against the oracle like the kernels (tests/test_host_logic_cpu.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/eppscore.h"
#include "epp_types.hpp"

namespace epp {

// ------------------------------------------------------------------------------------------------------------------
// LoRA labels
// ------------------------------------------------------------------------------------------------------------------
inline constexpr const char* LoraInfoRunningAdaptersMetricName = "running_lora_adapters";  // extractor.go:49-51
inline constexpr const char* LoraInfoWaitingAdaptersMetricName = "waiting_lora_adapters";
inline constexpr const char* LoraInfoMaxAdaptersMetricName = "max_lora";

namespace detail {
// strings.TrimSpace: leading and trailing Unicode White_Space code points (the set unicode.IsSpace accepts)
inline size_t space_len_at(const std::string& s, size_t i) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
  const size_t n = s.size();
  const unsigned b = p[i];
  if ((b >= 9 && b <= 13) || b == 32) return 1;
  if (b == 0xC2 && i + 1 < n && (p[i + 1] == 0x85 || p[i + 1] == 0xA0)) return 2;
  if (i + 2 < n) {
    const unsigned b1 = p[i + 1], b2 = p[i + 2];
    if ((b == 0xE1 && b1 == 0x9A && b2 == 0x80) ||
        (b == 0xE2 && b1 == 0x80 && ((b2 >= 0x80 && b2 <= 0x8A) || b2 == 0xA8 || b2 == 0xA9 || b2 == 0xAF)) ||
        (b == 0xE2 && b1 == 0x81 && b2 == 0x9F) || (b == 0xE3 && b1 == 0x80 && b2 == 0x80))
      return 3;
  }
  return 0;
}
inline std::string TrimSpace(const std::string& s) {
  size_t a = 0, b = s.size();
  for (size_t l; a < b && (l = space_len_at(s, a)) > 0;) a += l;
  for (;;) {  // trailing: try the 1-, 2- and 3-byte encodings that end at b
    bool cut = false;
    for (size_t l = 1; l <= 3 && !cut; l++)
      if (b >= a + l && space_len_at(s, b - l) == l) {
        b -= l;
        cut = true;
      }
    if (!cut) break;
  }
  return s.substr(a, b - a);
}
// strconv.Atoi: optional sign, decimal digits only, no spaces, no overflow past int (64-bit in the reference's build)
inline bool Atoi(const std::string& s, long long* out) {
  size_t i = 0;
  bool neg = false;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
  if (i >= s.size()) return false;
  unsigned long long v = 0;
  for (; i < s.size(); i++) {
    if (s[i] < '0' || s[i] > '9') return false;
    if (v > (0x7fffffffffffffffULL - (unsigned)(s[i] - '0')) / 10ULL + (neg ? 1 : 0)) return false;
    v = v * 10 + (unsigned)(s[i] - '0');
  }
  if (!neg && v > 0x7fffffffffffffffULL) return false;
  *out = neg ? -(long long)v : (long long)v;
  return true;
}
}  // namespace detail

// addAdapters (extractor.go:266-272): split on ',', trim, skip empties, store with value 0
inline void AddAdapters(std::map<std::string, int>& m, const std::string& csv) {
  size_t a = 0;
  for (;;) {
    const size_t c = csv.find(',', a);
    const std::string t = detail::TrimSpace(csv.substr(a, c == std::string::npos ? std::string::npos : c - a));
    if (!t.empty()) m[t] = 0;
    if (c == std::string::npos) break;
    a = c + 1;
  }
}
// populateLoRAMetrics (extractor.go:216-236).  Returns the number of parse errors appended (errs in the reference).
inline int PopulateLoRAMetrics(Metrics& clone, const std::vector<std::pair<std::string, std::string>>& labels) {
  int errs = 0;
  clone.ActiveModels.clear();
  clone.WaitingModels.clear();
  for (const auto& l : labels) {
    if (l.first == LoraInfoRunningAdaptersMetricName) {
      AddAdapters(clone.ActiveModels, l.second);
    } else if (l.first == LoraInfoWaitingAdaptersMetricName) {
      AddAdapters(clone.WaitingModels, l.second);
    } else if (l.first == LoraInfoMaxAdaptersMetricName) {
      if (!l.second.empty()) {
        long long v;
        if (detail::Atoi(l.second, &v)) clone.MaxActiveModels = (int)v;
        else errs++;
      }
    }
  }
  return errs;
}

// Adapter names -> the dictionary ids the engine's LoRA bitmasks and eppscore_batch.adapter_id use.  Ids are stable for
// the dictionary's lifetime; a name beyond the capacity gets -1 ("not a known adapter": it can only score through the
// capacity rule, and an endpoint's nmodels still counts it — lora_affinity.go:90).
class AdapterDictionary {
 public:
  explicit AdapterDictionary(int capacity) : cap_(capacity) {}
  int Intern(const std::string& name) {
    auto it = ids_.find(name);
    if (it != ids_.end()) return it->second;
    if ((int)ids_.size() >= cap_) return -1;
    const int id = (int)ids_.size();
    ids_[name] = id;
    return id;
  }
  int Lookup(const std::string& name) const {
    auto it = ids_.find(name);
    return it == ids_.end() ? -1 : it->second;
  }
  int size() const { return (int)ids_.size(); }

 private:
  int cap_;
  std::map<std::string, int> ids_;
};

// One endpoint's LoRA columns of eppscore_snapshot from its Metrics: bit a of active/waiting words, nmodels = the MAP
// sizes (an adapter in both maps counts twice; names outside the dictionary still count), max = MaxActiveModels.
inline void PackLoraColumns(const Metrics& m, AdapterDictionary& dict, int lora_words, uint64_t* active, uint64_t* waiting,
                            int32_t* nmodels, int32_t* max_active) {
  for (int w = 0; w < lora_words; w++) active[w] = waiting[w] = 0;
  for (const auto& kv : m.ActiveModels) {
    const int id = dict.Intern(kv.first);
    if (id >= 0 && id < lora_words * 64) active[id >> 6] |= 1ULL << (id & 63);
  }
  for (const auto& kv : m.WaitingModels) {
    const int id = dict.Intern(kv.first);
    if (id >= 0 && id < lora_words * 64) waiting[id >> 6] |= 1ULL << (id & 63);
  }
  *nmodels = (int32_t)(m.ActiveModels.size() + m.WaitingModels.size());
  *max_active = m.MaxActiveModels;
}

// ------------------------------------------------------------------------------------------------------------------
// max-score-picker with maxNumOfEndpoints = k (picker/maxscore/picker.go:87-115): the reference shuffles, then stable-sorts
// by descending score and takes the first k — i.e. descending score with a random order inside every tie class.  From one
// row of the engine's scores_out (NaN = not a candidate) this returns descending score, ties in ascending endpoint index
// (the deterministic member of the reference's distribution; position 0 equals the engine's LOWEST_INDEX pick).
// ------------------------------------------------------------------------------------------------------------------
inline std::vector<std::pair<int, double>> TopK(const double* scores, int M, int k) {
  std::vector<std::pair<int, double>> v;
  v.reserve((size_t)M);
  for (int m = 0; m < M; m++)
    if (scores[m] == scores[m]) v.emplace_back(m, scores[m]);
  std::stable_sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
  if ((int)v.size() > k) v.resize((size_t)std::max(k, 0));
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// SmallBatchCpu: request-independent profiles on the host (no prefix / latency / pair scorers), float64, multiply then add,
// scorer order from 0.0 (scheduler_profile.go:151-174), clamp (:194-202), arg-max with tie count.
// ------------------------------------------------------------------------------------------------------------------
struct CpuPick {
  int pick = -1;
  double score = 0.0;
  int tie_count = 0;
};
class SmallBatchCpu {
 public:
  // kinds/weights as in eppscore_config; returns false when the profile needs the device (prefix, latency, pair columns)
  bool Configure(const std::vector<int32_t>& kinds, const std::vector<double>& weights, double token_threshold = 4194304.0) {
    for (int k : kinds)
      if (!(k == EPPSCORE_SCORER_QUEUE || k == EPPSCORE_SCORER_KV_CACHE || k == EPPSCORE_SCORER_LORA || k == EPPSCORE_SCORER_RUNNING ||
            k == EPPSCORE_SCORER_TOKEN_LOAD))
        return false;
    kinds_ = kinds;
    weights_ = weights;
    thr_ = token_threshold > 0 ? token_threshold : 4194304.0;
    return true;
  }
  static double clamp01(double s) { return s < 0 ? 0 : (s > 1 ? 1 : s); }
  // One Schedule(): endpoints' metrics, the request's target model, optional candidate subset (nullptr = all)
  CpuPick Schedule(const std::vector<Endpoint>& eps, const std::string& target_model, const std::vector<int>* candidates = nullptr) const {
    std::vector<int> all;
    if (!candidates) {
      all.resize(eps.size());
      for (size_t i = 0; i < eps.size(); i++) all[i] = (int)i;
      candidates = &all;
    }
    CpuPick out;
    if (candidates->empty()) return out;  // "no endpoints available" (scheduler_profile.go:119-121)
    std::vector<double> acc(candidates->size(), 0.0);
    for (size_t s = 0; s < kinds_.size(); s++) {
      const int k = kinds_[s];
      const double w = weights_[s];
      long long mn = 0, mx = 0;
      if (k == EPPSCORE_SCORER_QUEUE || k == EPPSCORE_SCORER_RUNNING) {  // queue.go:79-91 / runningrequest.go:79-91
        bool first = true;
        for (int c : *candidates) {
          const long long q = k == EPPSCORE_SCORER_QUEUE ? eps[c].Metrics_.WaitingQueueSize : eps[c].Metrics_.RunningRequestsSize;
          if (first || q < mn) mn = q;
          if (first || q > mx) mx = q;
          first = false;
        }
      }
      for (size_t i = 0; i < candidates->size(); i++) {
        const Metrics& m = eps[(*candidates)[i]].Metrics_;
        volatile double sc = 0.0;
        if (k == EPPSCORE_SCORER_KV_CACHE) {
          sc = 1.0 - m.KVCacheUsagePercent;  // kvcache_utilization.go:79
        } else if (k == EPPSCORE_SCORER_QUEUE || k == EPPSCORE_SCORER_RUNNING) {
          const long long q = k == EPPSCORE_SCORER_QUEUE ? m.WaitingQueueSize : m.RunningRequestsSize;
          sc = mx == mn ? 1.0 : (double)(mx - q) / (double)(mx - mn);  // queue.go:95-99
        } else if (k == EPPSCORE_SCORER_LORA) {  // lora_affinity.go:84-99
          if (m.ActiveModels.count(target_model)) sc = 1.0;
          else if ((int)(m.ActiveModels.size() + m.WaitingModels.size()) < m.MaxActiveModels) sc = 0.8;
          else if (m.WaitingModels.count(target_model)) sc = 0.6;
          else sc = 0.0;
        } else {  // token-load-scorer, token_load.go:91-110
          double load = 0.0;                         // attribute absent (InFlightTokens < 0): tokenLoad stays 0.0
          if (eps[(*candidates)[i]].InFlightTokens >= 0) load = (double)eps[(*candidates)[i]].InFlightTokens;
          if (load <= 0) {
            sc = 1.0;
          } else {
            if (load > thr_) load = thr_;
            sc = 1.0 - (load / thr_);
          }
        }
        volatile double term = clamp01(sc) * w;  // multiply ...
        acc[i] = acc[i] + term;                  // ... then add (never contracted: build with -ffp-contract=off)
      }
    }
    for (size_t i = 0; i < candidates->size(); i++) {
      if (out.pick < 0 || acc[i] > out.score) {
        out.pick = (*candidates)[i];
        out.score = acc[i];
        out.tie_count = 1;
      } else if (acc[i] == out.score) {
        out.tie_count++;
      }
    }
    return out;
  }

 private:
  std::vector<int32_t> kinds_;
  std::vector<double> weights_;
  double thr_ = 4194304.0;
};

}  // namespace epp
