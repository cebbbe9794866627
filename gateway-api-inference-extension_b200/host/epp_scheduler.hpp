// epp_scheduler.hpp — host side of the drop-in: the reference's Scheduler / SchedulerProfile / Filter /
// Scorer / Picker plugin interface for THIS path, restated in C++ above the C ABI (include/eppscore.h).
//
// The reference's toolchain (Go) is not in this image, so the layer a Go maintainer would write with cgo
//  is written here in C++ with the same names, argument meaning and error behaviour:
//   Scheduler.Schedule            pkg/epp/scheduling/scheduler.go:54-102
//   SchedulerProfile              pkg/epp/scheduling/scheduler_profile.go:41-128   (filters → scorers → picker)
//   WeightedScorer                pkg/epp/scheduling/weighted_scorer.go:32
//   Filter / Scorer / Picker      pkg/epp/framework/interface/scheduling/plugins.go:43-78
//   InferenceRequest / Endpoint / ScoredEndpoint / ProfileRunResult / SchedulingResult
//                                 pkg/epp/framework/interface/scheduling/types.go:44-170
//   Metrics                       pkg/epp/framework/interface/datalayer/metrics.go:26-42
//   SingleProfileHandler          pkg/epp/framework/plugins/scheduling/profile/single_profile_handler.go:66-99
//   PrepareRequestData/PreRequest pkg/epp/framework/plugins/requestcontrol/dataproducer/approximateprefix/plugin.go:140-205
// Scorers here are DESCRIPTORS: they name which kernel-side scorer runs and with what weight; the arithmetic
// itself happens in the CUDA kernels.  Filters run on the host exactly like the reference's filter chain and
// become the per-request candidate mask.  One Scheduler owns one engine (one GPU).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/eppscore.h"
#include "epp_types.hpp"
#include "host_eval.hpp"

namespace epp {

// ---- plugins (descriptors) ----
struct Scorer {
  std::string Type;
  int32_t Kind;
  virtual ~Scorer() = default;
  Scorer(std::string t, int32_t k) : Type(std::move(t)), Kind(k) {}
};
struct KVCacheUtilizationScorer : Scorer { KVCacheUtilizationScorer() : Scorer("kv-cache-utilization-scorer", EPPSCORE_SCORER_KV_CACHE) {} };
struct QueueScorer : Scorer { QueueScorer() : Scorer("queue-scorer", EPPSCORE_SCORER_QUEUE) {} };
struct LoraAffinityScorer : Scorer { LoraAffinityScorer() : Scorer("lora-affinity-scorer", EPPSCORE_SCORER_LORA) {} };
struct PrefixCacheScorer : Scorer { PrefixCacheScorer() : Scorer("prefix-cache-scorer", EPPSCORE_SCORER_PREFIX) {} };
struct RunningRequestsScorer : Scorer { RunningRequestsScorer() : Scorer("running-requests-size-scorer", EPPSCORE_SCORER_RUNNING) {} };
// token-load-scorer (scorer/tokenload/token_load.go:36-67)
struct TokenLoadScorer : Scorer {
  int64_t QueueThresholdTokens;
  explicit TokenLoadScorer(int64_t thr = 4194304) : Scorer("token-load-scorer", EPPSCORE_SCORER_TOKEN_LOAD), QueueThresholdTokens(thr > 0 ? thr : 4194304) {}
};
// latency-scorer (scorer/latency/plugin.go:59-90); fed by the PredictedLatencyProducer below
struct LatencyScorer : Scorer {
  double TTFTWeight = 0.8, TPOTWeight = 0.2;
  bool StrategyMost = false;  // HeadroomSelectionStrategy "most"; default "least"
  double CompositeKVWeight = 1, CompositeQueueWeight = 1, CompositePrefixWeight = 1;
  LatencyScorer() : Scorer("latency-scorer", EPPSCORE_SCORER_LATENCY) {}
};

// predicted-latency-producer (requestcontrol/dataproducer/predictedlatency): its Config (plugin.go:118-136), the
// Bayesian-ridge coefficients the Go client caches (latencypredictorasync/prediction.go:164-194) and the
// per-endpoint running-request bookkeeping it reads while predicting (plugin.go:347-363).
struct PredictedLatencyProducer {
  double SLOBufferFactor = 1.0;
  bool StreamingMode = false;
  std::string EndpointRoleLabel;
  bool HavePredictions = true;  // false = sidecar down / timed out: the scorer's composite fallback
  double TTFTIntercept = 0, TPOTIntercept = 0;
  std::map<std::string, double> TTFTCoeffs, TPOTCoeffs;  // keys as in the reference's MetricsResponse
  struct EndpointState {
    double MinTPOTSLO = 0;   // getEndpointMinTPOTSLO
    int RunningRequests = 0; // getEndpointRunningRequestCount
  };
  std::map<std::string, EndpointState> State;  // keyed by NamespacedName.String()
};

struct WeightedScorer {
  std::shared_ptr<Scorer> Scorer_;
  double Weight_;
  double Weight() const { return Weight_; }
};
inline WeightedScorer NewWeightedScorer(std::shared_ptr<Scorer> s, double w) { return WeightedScorer{std::move(s), w}; }

// framework.Filter: given the request and the current candidates (indices into the slice passed to Schedule),
// return the ones to keep.  Runs on the host, chained in order, early exit on empty (scheduler_profile.go:130-149).
struct Filter {
  virtual ~Filter() = default;
  virtual std::vector<int> Filter_(const InferenceRequest& req, const std::vector<Endpoint>& endpoints,
                                   const std::vector<int>& candidates) = 0;
};

// Filters the engine evaluates on the device (include/eppscore.h: eppscore_filter_kind).  They are descriptors like the
// scorers; any other Filter subclass runs on the host and reaches the engine as the candidate mask.  Device filters
// run AFTER the host filters, in the order given to WithDeviceFilters.
struct DeviceFilter {
  int32_t Kind;
  double Param[3];
};
// prefix-cache-affinity-filter (filter/prefixcacheaffinity/plugin.go:52-62): defaults 0.80 / 0.01 / 5000
inline DeviceFilter PrefixCacheAffinityFilter(double affinityThreshold = 0.80, double explorationProbability = 0.01,
                                              double maxTTFTPenaltyMs = 5000) {
  return DeviceFilter{EPPSCORE_FILTER_PREFIX_AFFINITY, {affinityThreshold, explorationProbability, maxTTFTPenaltyMs}};
}
// slo-headroom-tier-filter (filter/sloheadroomtier/plugin.go:50-52): default epsilonExploreNeg 0.01
inline DeviceFilter SLOHeadroomTierFilter(double epsilonExploreNeg = 0.01) {
  return DeviceFilter{EPPSCORE_FILTER_SLO_HEADROOM_TIER, {epsilonExploreNeg, 0, 0}};
}

struct MaxScorePicker {
  int MaxNumOfEndpoints = 1;  // picker.DefaultMaxNumOfEndpoints (picker/common.go:36); > 1: top-k from the engine's score rows
  int Mode = EPPSCORE_PICK_MAX_SCORE;
};
// weighted-random-picker (A-Res, picker/weightedrandom/picker.go:111-155) and random-picker (picker/random/picker.go:85-101):
// same descriptor with another pick mode; the engine's generator is counter-based (SchedulerConfig.TieSeed).
struct WeightedRandomPicker : MaxScorePicker { WeightedRandomPicker() { Mode = EPPSCORE_PICK_WEIGHTED_RANDOM; } };
struct RandomPicker : MaxScorePicker { RandomPicker() { Mode = EPPSCORE_PICK_RANDOM; } };

class SchedulerProfile {
 public:
  SchedulerProfile& WithFilters(std::vector<std::shared_ptr<Filter>> f) { filters_ = std::move(f); return *this; }
  SchedulerProfile& WithScorers(std::vector<WeightedScorer> s) { scorers_ = std::move(s); return *this; }
  SchedulerProfile& WithPicker(MaxScorePicker p) { picker_ = p; return *this; }
  SchedulerProfile& WithDeviceFilters(std::vector<DeviceFilter> f) { device_filters_ = std::move(f); return *this; }
  const std::vector<DeviceFilter>& device_filters() const { return device_filters_; }
  SchedulerProfile& WithPredictedLatencyProducer(std::shared_ptr<PredictedLatencyProducer> p) { producer_ = std::move(p); return *this; }
  const std::shared_ptr<PredictedLatencyProducer>& producer() const { return producer_; }
  const std::vector<std::shared_ptr<Filter>>& filters() const { return filters_; }
  const std::vector<WeightedScorer>& scorers() const { return scorers_; }
  const MaxScorePicker& picker() const { return picker_; }

 private:
  std::vector<std::shared_ptr<Filter>> filters_;
  std::vector<WeightedScorer> scorers_;
  MaxScorePicker picker_;
  std::shared_ptr<PredictedLatencyProducer> producer_;
  std::vector<DeviceFilter> device_filters_;
};

// approximateprefix config (types.go:77-141)
struct PrefixConfig {
  bool AutoTune = true;
  int BlockSizeTokens = 16;
  int MaxPrefixBlocksToMatch = 256;
  int MaxPrefixTokensToMatch = 0;
  int LRUCapacityPerServer = 31250;
};

struct SchedulerConfig {
  std::string ProfileName = "default";  // SingleProfileHandler: exactly one profile
  SchedulerProfile Profile;
  PrefixConfig Prefix;
  int Device = 0;
  int MaxEndpoints = 1024;
  int MaxAdapters = 64;
  int64_t PrefixCapacity = 1 << 18;
  int TieMode = EPPSCORE_TIE_LOWEST_INDEX;
  uint64_t TieSeed = 0;
  // Batches of at most this many requests whose profile is request-independent (queue / kv / running / lora / token-load
  // scorers, no device filters, max-score picker, lowest-index ties) are evaluated on the host (SmallBatchCpu): a launch
  // costs tens of microseconds, a 1 x 4 queue-scorer Schedule a fraction of one.  0 disables the route.
  int CpuBatchThreshold = 8;
};

class Scheduler {
 public:
  explicit Scheduler(const SchedulerConfig& cfg) : cfg_(cfg) {
    eppscore_config c;
    eppscore_config_default(&c);
    const auto& sc = cfg.Profile.scorers();
    if (sc.size() > EPPSCORE_MAX_SCORERS) throw SchedulingError("too many scorers in profile");
    if (cfg.Profile.picker().MaxNumOfEndpoints < 1) throw SchedulingError("maxNumOfEndpoints must be >= 1");
    if (cfg.Profile.picker().MaxNumOfEndpoints > 1 && cfg.Profile.picker().Mode != EPPSCORE_PICK_MAX_SCORE)
      throw SchedulingError("maxNumOfEndpoints > 1 is supported for the max-score picker only");
    c.n_scorers = (int32_t)sc.size();
    {  // the small-batch CPU route applies to request-independent profiles only
      std::vector<int32_t> kinds;
      std::vector<double> weights;
      for (auto& w : sc) {
        kinds.push_back(w.Scorer_->Kind);
        weights.push_back(w.Weight());
      }
      double thr = 4194304.0;
      for (auto& w : sc)
        if (auto* tl = dynamic_cast<TokenLoadScorer*>(w.Scorer_.get())) thr = (double)tl->QueueThresholdTokens;
      cpu_route_ok_ = cfg.CpuBatchThreshold > 0 && cfg.Profile.device_filters().empty() && cfg.Profile.picker().Mode == EPPSCORE_PICK_MAX_SCORE &&
                      cfg.Profile.picker().MaxNumOfEndpoints == 1 && cfg.TieMode == EPPSCORE_TIE_LOWEST_INDEX && cpu_.Configure(kinds, weights, thr);
    }
    for (size_t i = 0; i < sc.size(); i++) {
      c.scorer_kind[i] = sc[i].Scorer_->Kind;
      c.scorer_weight[i] = sc[i].Weight();
      if (auto* tl = dynamic_cast<TokenLoadScorer*>(sc[i].Scorer_.get())) c.token_load_threshold = (double)tl->QueueThresholdTokens;
      if (auto* ls = dynamic_cast<LatencyScorer*>(sc[i].Scorer_.get())) latency_scorer_ = ls;
    }
    if (latency_scorer_ && !cfg.Profile.producer()) throw SchedulingError("latency-scorer needs a predicted-latency-producer");
    c.block_chars = cfg.Prefix.BlockSizeTokens * 4;  // averageCharactersPerToken (types.go:112)
    c.max_blocks = cfg.Prefix.MaxPrefixBlocksToMatch;
    c.lru_capacity_default = cfg.Prefix.LRUCapacityPerServer;
    c.max_endpoints = cfg.MaxEndpoints;
    c.max_adapters = cfg.MaxAdapters;
    c.prefix_capacity = cfg.PrefixCapacity;
    c.pick_mode = cfg.Profile.picker().Mode;
    if (cfg.Profile.device_filters().size() > EPPSCORE_MAX_FILTERS) throw SchedulingError("too many device filters");
    c.n_filters = (int32_t)cfg.Profile.device_filters().size();
    for (size_t i = 0; i < cfg.Profile.device_filters().size(); i++) {
      c.filter_kind[i] = cfg.Profile.device_filters()[i].Kind;
      for (int q = 0; q < 3; q++) c.filter_param[i][q] = cfg.Profile.device_filters()[i].Param[q];
    }
    c.tie_mode = cfg.TieMode;
    c.tie_seed = cfg.TieSeed;
    if (eppscore_create(cfg.Device, &c, &eng_) != EPPSCORE_OK) throw SchedulingError(std::string("eppscore_create: ") + eppscore_last_error(nullptr));
  }
  ~Scheduler() { eppscore_destroy(eng_); }
  Scheduler(const Scheduler&) = delete;
  Scheduler& operator=(const Scheduler&) = delete;

  // Schedule one request (scheduler.go:54): a batch of one.
  SchedulingResult Schedule(const InferenceRequest& request, const std::vector<Endpoint>& candidateEndpoints) {
    auto res = ScheduleBatch({request}, candidateEndpoints);
    if (!res[0].error.empty()) throw SchedulingError(res[0].error);
    return res[0].result;
  }

  struct BatchItem {
    SchedulingResult result;
    std::string error;  // non-empty ⇒ the reference would have returned this error for that request
  };

  // R concurrent Schedule() calls against the same candidate slice, as ONE engine batch.
  std::vector<BatchItem> ScheduleBatch(const std::vector<InferenceRequest>& requests, const std::vector<Endpoint>& endpoints) {
    const int R = (int)requests.size(), M = (int)endpoints.size();
    std::vector<BatchItem> out((size_t)R);
    last_hashes_.clear();
    last_nh_.clear();
    if (R == 0) return out;
    if (M == 0) {  // scheduler_profile.go:119-121 → single_profile_handler.go:89-91
      for (auto& o : out) o.error = "failed to run scheduler profile '" + cfg_.ProfileName + "'";
      return out;
    }
    if (cpu_route_ok_ && R <= cfg_.CpuBatchThreshold) {  // tiny batch, request-independent profile: stay on the host
      std::vector<int> all((size_t)M);
      for (int m = 0; m < M; m++) all[m] = m;
      for (int r = 0; r < R; r++) {
        std::vector<int> cand = all;
        for (auto& f : cfg_.Profile.filters()) {
          cand = f->Filter_(requests[r], endpoints, cand);
          if (cand.empty()) break;
        }
        const CpuPick p = cpu_.Schedule(endpoints, requests[r].TargetModel, cfg_.Profile.filters().empty() ? nullptr : &cand);
        if (p.pick < 0) {
          out[r].error = "failed to run scheduler profile '" + cfg_.ProfileName + "'";
          continue;
        }
        ScoredEndpoint se;
        se.Endpoint_ = &endpoints[(size_t)p.pick];
        se.Index = p.pick;
        se.Score = p.score;
        se.TieCount = p.tie_count;
        out[r].result.ProfileResults[cfg_.ProfileName].TargetEndpoints.push_back(se);
        out[r].result.PrimaryProfileName = cfg_.ProfileName;
      }
      cpu_routed_ += (uint64_t)R;
      return out;
    }
    // --- stable endpoint ids: the prefix index refers to endpoints by id (ServerID = NamespacedName, indexer.go:34-35), and
    // neither the order nor the membership of the candidate list is stable between calls (datastore PodList, subsetting:
    // director candidates.go:98).  An endpoint keeps the id it got when it was first seen; the engine's rows are ids, the
    // list's positions are mapped to and from them.  When the list is exactly ids 0..M-1 in order (the common case, and every
    // single-list use) nothing is remapped; a mask is only needed when some row has no endpoint in this list. ---
    std::vector<int> ids((size_t)M);
    int Mx = 0;
    bool identity = true;
    {
      std::vector<int> seen;
      for (int m = 0; m < M; m++) {
        ids[m] = IdOf(endpoints[(size_t)m].GetMetadata()->NamespacedName_.String());
        Mx = std::max(Mx, ids[m] + 1);
        identity = identity && ids[m] == m;
      }
      seen.assign((size_t)Mx, 0);
      bool dup = false;
      for (int m = 0; m < M; m++) dup = dup || seen[(size_t)ids[m]]++ > 0;
      if (dup) {  // the same server twice in one list: positions are the only usable ids for this batch
        for (int m = 0; m < M; m++) ids[m] = m;
        Mx = M;
        identity = true;
      }
    }
    identity = identity && Mx == M;
    if (Mx > cfg_.MaxEndpoints) throw SchedulingError("more endpoints than SchedulerConfig.MaxEndpoints");
    if (!identity) remapped_batches_++;
    std::vector<int> pos_of_id((size_t)Mx, -1);
    for (int m = 0; m < M; m++) pos_of_id[(size_t)ids[m]] = m;
    PackSnapshot(endpoints, ids, Mx);
    // --- filters → candidate mask (rows are ids; ids without an endpoint in this list are never candidates) ---
    const int mw = (Mx + 31) / 32;
    std::vector<uint32_t> mask;
    const bool have_filters = !cfg_.Profile.filters().empty();
    const bool need_mask = have_filters || Mx != M;  // a pure permutation of the rows has no holes: no mask, the fast path stays
    if (need_mask) {
      mask.assign((size_t)R * mw, 0u);
      std::vector<int> all((size_t)M);
      for (int m = 0; m < M; m++) all[m] = m;
      for (int r = 0; r < R; r++) {
        std::vector<int> cand = all;
        for (auto& f : cfg_.Profile.filters()) {
          cand = f->Filter_(requests[r], endpoints, cand);
          if (cand.empty()) break;
        }
        for (int m : cand) mask[(size_t)r * mw + (ids[m] >> 5)] |= 1u << (ids[m] & 31);
      }
    }
    // --- PrepareRequestData inputs (approximateprefix/plugin.go:140-165) ---
    int block_tokens = cfg_.Prefix.BlockSizeTokens;                        // GetBlockSize: endpoints[0] when autotuning
    if (cfg_.Prefix.AutoTune && endpoints[0].GetMetrics()->CacheBlockSize > 0) block_tokens = endpoints[0].GetMetrics()->CacheBlockSize;
    int max_blocks = cfg_.Prefix.MaxPrefixBlocksToMatch;
    if (cfg_.Prefix.MaxPrefixTokensToMatch > 0 && block_tokens > 0) max_blocks = cfg_.Prefix.MaxPrefixTokensToMatch / block_tokens;
    std::vector<uint8_t> bytes;
    std::vector<int64_t> off((size_t)R + 1, 0);
    std::vector<int32_t> len((size_t)R), adapter((size_t)R), in_tokens;
    std::vector<double> ttft_slo, tpot_slo;
    if (latency_scorer_) {  // predictedlatency/plugin.go:276-290, :315-343
      in_tokens.resize((size_t)R);
      ttft_slo.assign((size_t)R, 0.0);
      tpot_slo.assign((size_t)R, 0.0);
      auto header = [](const InferenceRequest& q, const char* key) {
        auto it = q.Headers.find(key);
        if (it == q.Headers.end()) return 0.0;
        char* end = nullptr;
        const double v = strtod(it->second.c_str(), &end);
        return (end && *end == 0 && end != it->second.c_str()) ? v : 0.0;  // parse error ⇒ 0 (logged only)
      };
      for (int r = 0; r < R; r++) {
        in_tokens[r] = CountFields(requests[r].Prompt);
        ttft_slo[r] = header(requests[r], "x-slo-ttft-ms");
        tpot_slo[r] = header(requests[r], "x-slo-tpot-ms");
      }
    }
    std::vector<uint64_t> seed((size_t)R);
    for (int r = 0; r < R; r++) {
      while (bytes.size() % 16) bytes.push_back(0);  // 16-byte aligned starts: the fast hash path
      off[r] = (int64_t)bytes.size();
      len[r] = (int32_t)requests[r].Prompt.size();
      bytes.insert(bytes.end(), requests[r].Prompt.begin(), requests[r].Prompt.end());
      seed[r] = eppscore_model_seed(requests[r].TargetModel.data(), requests[r].TargetModel.size(),
                                    requests[r].CacheSalt.data(), requests[r].CacheSalt.size());
      auto it = adapter_ids_.find(requests[r].TargetModel);
      adapter[r] = it == adapter_ids_.end() ? -1 : it->second;
    }
    off[R] = (int64_t)bytes.size();
    bytes.resize(bytes.size() + 64, 0);
    const bool want_prefix = max_blocks > 0 && block_tokens > 0;
    std::vector<int32_t> pick((size_t)R), ties((size_t)R);
    std::vector<double> score((size_t)R);
    last_nh_.assign((size_t)R, 0);
    last_stride_ = want_prefix ? max_blocks : 1;
    last_hashes_.assign((size_t)R * last_stride_, 0);
    eppscore_batch b{};
    b.struct_size = sizeof(b);
    b.R = R;
    b.prompt_bytes = want_prefix ? bytes.data() : nullptr;
    b.prompt_off = off.data();
    b.prompt_len = len.data();
    b.model_seed = seed.data();
    b.block_chars = block_tokens * 4;
    b.max_blocks = max_blocks;
    b.adapter_id = adapter.data();
    b.cand_mask = need_mask ? mask.data() : nullptr;
    if (latency_scorer_) {
      b.input_tokens = in_tokens.data();
      b.ttft_slo = ttft_slo.data();
      b.tpot_slo = tpot_slo.data();
    }
    b.pick = pick.data();
    b.pick_score = score.data();
    b.tie_count = ties.data();
    b.total_blocks = last_nh_.data();
    b.hashes_out = want_prefix ? last_hashes_.data() : nullptr;
    const int topk = cfg_.Profile.picker().MaxNumOfEndpoints;
    std::vector<double> all_scores;
    if (topk > 1) {  // the whole weightedScorePerEndpoint map, NaN for non-candidates (scheduler_profile.go:155-174)
      all_scores.resize((size_t)R * Mx);
      b.scores_out = all_scores.data();
    }
    if (eppscore_schedule_batch(eng_, &b) != EPPSCORE_OK) throw SchedulingError(std::string("eppscore_schedule_batch: ") + eppscore_last_error(eng_));
    last_endpoints_ = &endpoints;
    last_ids_ = ids;
    last_rows_ = Mx;
    for (int r = 0; r < R; r++) {
      if (pick[r] < 0) {  // "no endpoints available for the given request" → profile result nil → ProcessResults error
        out[r].error = "failed to run scheduler profile '" + cfg_.ProfileName + "'";
        continue;
      }
      if (pick[r] >= Mx || pos_of_id[(size_t)pick[r]] < 0) throw SchedulingError("engine picked an endpoint that is not in the candidate list");
      ScoredEndpoint se;
      se.Endpoint_ = &endpoints[(size_t)pos_of_id[(size_t)pick[r]]];
      se.Index = pos_of_id[(size_t)pick[r]];
      se.Score = score[r];
      se.TieCount = ties[r];
      out[r].result.ProfileResults[cfg_.ProfileName].TargetEndpoints.push_back(se);
      out[r].result.PrimaryProfileName = cfg_.ProfileName;
      if (topk > 1) {  // picker/maxscore/picker.go:104-106: the next best candidates, descending score
        auto& te = out[r].result.ProfileResults[cfg_.ProfileName].TargetEndpoints;
        for (const auto& pr : TopK(all_scores.data() + (size_t)r * Mx, Mx, topk)) {  // (id, score), NaN rows skipped
          const int pos = pos_of_id[(size_t)pr.first];
          if (pos < 0 || pos == se.Index) continue;
          if ((int)te.size() >= topk) break;
          ScoredEndpoint o2;
          o2.Endpoint_ = &endpoints[(size_t)pos];
          o2.Index = pos;
          o2.Score = pr.second;
          te.push_back(o2);
        }
      }
    }
    return out;
  }
  uint64_t CpuRoutedRequests() const { return cpu_routed_; }

  // PreRequest for the last ScheduleBatch (approximateprefix/plugin.go:169-197): records the picks in the index.
  void PreRequest(const std::vector<BatchItem>& results) {
    if (!last_endpoints_ || last_hashes_.empty()) return;
    const int R = (int)results.size();
    std::vector<int32_t> pick((size_t)R, -1);
    for (int r = 0; r < R; r++)
      if (results[r].error.empty()) pick[r] = last_ids_[(size_t)results[r].result.ProfileResults.at(cfg_.ProfileName).TargetEndpoints[0].Index];
    std::vector<int32_t> cap((size_t)last_rows_, 0);
    for (size_t m = 0; m < last_endpoints_->size(); m++)  // makeserver, plugin.go:207-216
      cap[(size_t)last_ids_[m]] = (cfg_.Prefix.AutoTune && (*last_endpoints_)[m].GetMetrics()->CacheNumBlocks > 0) ? (*last_endpoints_)[m].GetMetrics()->CacheNumBlocks : 0;
    if (eppscore_commit_picks(eng_, R, pick.data(), last_hashes_.data(), last_nh_.data(), last_stride_, cap.data()) != EPPSCORE_OK)
      throw SchedulingError(std::string("eppscore_commit_picks: ") + eppscore_last_error(eng_));
  }

  // indexer.RemovePod (indexer.go:167-182), called when a model server leaves the pool: its LRU and its memberships go, its
  // id is free for the next new server.
  void RemovePod(const std::string& namespaced_name) {
    auto it = server_ids_.find(namespaced_name);
    if (it == server_ids_.end()) return;
    if (eppscore_prefix_remove_endpoint(eng_, it->second) != EPPSCORE_OK)
      throw SchedulingError(std::string("eppscore_prefix_remove_endpoint: ") + eppscore_last_error(eng_));
    free_ids_.push_back(it->second);
    server_ids_.erase(it);
  }
  int ServerId(const std::string& namespaced_name) const {  // -1: never seen (or removed)
    auto it = server_ids_.find(namespaced_name);
    return it == server_ids_.end() ? -1 : it->second;
  }
  uint64_t RemappedBatches() const { return remapped_batches_; }  // batches whose list was not ids 0..M-1 in order

  const std::vector<uint16_t>& LastTotalBlocks() const { return last_nh_; }
  eppscore_engine* engine() { return eng_; }

 private:
  // Metrics maps → the packed SoA snapshot; adapter names get dictionary ids in first-seen order.
  int IdOf(const std::string& namespaced_name) {
    auto it = server_ids_.find(namespaced_name);
    if (it != server_ids_.end()) return it->second;
    int id;
    if (!free_ids_.empty()) {
      id = free_ids_.back();
      free_ids_.pop_back();
    } else {
      id = (int)server_ids_.size();  // no free id ⇒ the ids in use are exactly 0..size-1
    }
    server_ids_.emplace(namespaced_name, id);
    return id;
  }

  // rows of the snapshot are endpoint ids: endpoint eps[i] fills row ids[i]; rows without an endpoint stay zero (they are
  // never candidates: the batch's mask excludes them)
  void PackSnapshot(const std::vector<Endpoint>& eps, const std::vector<int>& ids, int rows) {
    const int N = (int)eps.size();
    const int M = rows;
    adapter_ids_.clear();
    for (auto& e : eps) {
      for (auto& kv : e.GetMetrics()->ActiveModels) adapter_ids_.emplace(kv.first, (int)adapter_ids_.size());
      for (auto& kv : e.GetMetrics()->WaitingModels) adapter_ids_.emplace(kv.first, (int)adapter_ids_.size());
    }
    const int words = (int)((adapter_ids_.size() + 63) / 64) > 0 ? (int)((adapter_ids_.size() + 63) / 64) : 1;
    std::vector<double> kv((size_t)M);
    std::vector<int64_t> queue((size_t)M), running((size_t)M);
    std::vector<uint64_t> act((size_t)M * words, 0), wait((size_t)M * words, 0);
    std::vector<int32_t> nm((size_t)M), mx((size_t)M), dispatched((size_t)M, 0);
    std::vector<int64_t> tokens((size_t)M, 0);
    std::vector<double> min_tpot((size_t)M, 0.0);
    std::vector<uint8_t> prefill((size_t)M, 0);
    bool have_tokens = false;
    const PredictedLatencyProducer* prod = cfg_.Profile.producer().get();
    for (int i = 0; i < N; i++) {
      const int m = ids[(size_t)i];  // the endpoint's row
      const Endpoint& ep = eps[(size_t)i];
      if (ep.InFlightTokens >= 0) {
        tokens[m] = ep.InFlightTokens;
        have_tokens = true;
      }
      if (prod) {
        auto st = prod->State.find(ep.GetMetadata()->NamespacedName_.String());
        if (st != prod->State.end()) {
          min_tpot[m] = st->second.RunningRequests > 0 ? st->second.MinTPOTSLO : 0.0;  // plugin.go:349-354
          dispatched[m] = st->second.RunningRequests;
        }
        if (!prod->EndpointRoleLabel.empty()) {  // hasPrefillRole, prediction.go:168-175
          auto lb = ep.GetMetadata()->Labels.find(prod->EndpointRoleLabel);
          prefill[m] = lb != ep.GetMetadata()->Labels.end() && lb->second == "prefill";
        }
      }
    }
    if (latency_scorer_ && prod) {
      eppscore_latency_params lp;
      eppscore_latency_params_default(&lp);
      auto coef = [](const std::map<std::string, double>& c, const char* k) {
        auto it = c.find(k);
        return it == c.end() ? 0.0 : it->second;  // Go map lookup of a missing key yields 0
      };
      lp.has_predictions = prod->HavePredictions ? 1 : 0;
      lp.ttft_intercept = prod->TTFTIntercept;
      lp.ttft_kv = coef(prod->TTFTCoeffs, "kv_cache_percentage");
      lp.ttft_input = coef(prod->TTFTCoeffs, "input_token_length");
      lp.ttft_waiting = coef(prod->TTFTCoeffs, "num_request_waiting");
      lp.ttft_running = coef(prod->TTFTCoeffs, "num_request_running");
      lp.ttft_prefix = coef(prod->TTFTCoeffs, "prefix_cache_score");
      lp.tpot_intercept = prod->TPOTIntercept;
      lp.tpot_kv = coef(prod->TPOTCoeffs, "kv_cache_percentage");
      lp.tpot_input = coef(prod->TPOTCoeffs, "input_token_length");
      lp.tpot_waiting = coef(prod->TPOTCoeffs, "num_request_waiting");
      lp.tpot_running = coef(prod->TPOTCoeffs, "num_request_running");
      lp.tpot_generated = coef(prod->TPOTCoeffs, "num_tokens_generated");
      lp.slo_buffer_factor = prod->SLOBufferFactor;
      lp.streaming_mode = prod->StreamingMode ? 1 : 0;
      lp.strategy_most = latency_scorer_->StrategyMost ? 1 : 0;
      lp.ttft_weight = latency_scorer_->TTFTWeight;
      lp.tpot_weight = latency_scorer_->TPOTWeight;
      lp.composite_kv = latency_scorer_->CompositeKVWeight;
      lp.composite_queue = latency_scorer_->CompositeQueueWeight;
      lp.composite_prefix = latency_scorer_->CompositePrefixWeight;
      if (eppscore_set_latency_params(eng_, &lp) != EPPSCORE_OK) throw SchedulingError(std::string("eppscore_set_latency_params: ") + eppscore_last_error(eng_));
    }
    for (int i = 0; i < N; i++) {
      const int m = ids[(size_t)i];
      const Metrics* x = eps[(size_t)i].GetMetrics();
      kv[m] = x->KVCacheUsagePercent;
      queue[m] = x->WaitingQueueSize;
      running[m] = x->RunningRequestsSize;
      for (auto& a : x->ActiveModels) act[(size_t)m * words + (adapter_ids_[a.first] >> 6)] |= 1ULL << (adapter_ids_[a.first] & 63);
      for (auto& a : x->WaitingModels) wait[(size_t)m * words + (adapter_ids_[a.first] >> 6)] |= 1ULL << (adapter_ids_[a.first] & 63);
      nm[m] = (int32_t)(x->ActiveModels.size() + x->WaitingModels.size());  // len()+len() as MAP sizes (lora_affinity.go:90)
      mx[m] = x->MaxActiveModels;
    }
    eppscore_snapshot s{};
    s.struct_size = sizeof(s);
    s.M = M;
    s.lora_words = words;
    s.kv_usage = kv.data();
    s.queue = queue.data();
    s.running = running.data();
    s.lora_active = act.data();
    s.lora_waiting = wait.data();
    s.lora_nmodels = nm.data();
    s.lora_max = mx.data();
    s.inflight_tokens = have_tokens ? tokens.data() : nullptr;
    if (prod) {
      s.min_tpot_slo = min_tpot.data();
      s.dispatched = dispatched.data();
      s.prefill_role = prefill.data();
    }
    s.epoch = ++epoch_;
    if (eppscore_set_snapshot(eng_, &s) != EPPSCORE_OK) throw SchedulingError(std::string("eppscore_set_snapshot: ") + eppscore_last_error(eng_));
  }

  SchedulerConfig cfg_;
  SmallBatchCpu cpu_;
  bool cpu_route_ok_ = false;
  uint64_t cpu_routed_ = 0;
  eppscore_engine* eng_ = nullptr;
  const LatencyScorer* latency_scorer_ = nullptr;
  std::unordered_map<std::string, int> adapter_ids_;
  uint64_t epoch_ = 0;
  std::vector<uint64_t> last_hashes_;
  std::vector<uint16_t> last_nh_;
  int32_t last_stride_ = 1;
  const std::vector<Endpoint>* last_endpoints_ = nullptr;
  std::vector<int> last_ids_;                       // list position -> endpoint id of the last engine batch
  int last_rows_ = 0;
  std::unordered_map<std::string, int> server_ids_;  // ServerID (NamespacedName) -> stable endpoint id
  std::vector<int> free_ids_;
  uint64_t remapped_batches_ = 0;
};

inline Scheduler* NewSchedulerWithConfig(const SchedulerConfig& c) { return new Scheduler(c); }

}  // namespace epp
