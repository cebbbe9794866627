// CPU-only tests of the host pieces above the C ABI that need no engine (gateway-api-inference-extension_b200/host/coalescer.hpp):
// the coalescing front (64 caller threads over a mock backend; built with -fsanitize=thread by the pytest wrapper), the LoRA
// metric-label parser with the reference's vectors (extractor/metrics/spec_test.go:360-361, extractor_test.go:149-153),
// top-k from a score row (picker/maxscore/picker_test.go:43-110), and a C shim for the small-batch CPU route so the Python
// side can compare it with the oracle.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../gateway-api-inference-extension_b200/host/coalescer.hpp"

using namespace epp;

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);    \
      return 1;                                                         \
    }                                                                   \
  } while (0)

struct MockBackend {
  struct BatchItem {
    SchedulingResult result;
    std::string error;
  };
  std::atomic<int> calls{0}, in_flight{0}, max_in_flight{0};
  std::vector<BatchItem> ScheduleBatch(const std::vector<InferenceRequest>& reqs, const std::vector<Endpoint>& eps) {
    const int now = ++in_flight;
    int m = max_in_flight.load();
    while (now > m && !max_in_flight.compare_exchange_weak(m, now)) {
    }
    calls++;
    std::this_thread::sleep_for(std::chrono::microseconds(200));  // "the GPU"
    std::vector<BatchItem> out(reqs.size());
    for (size_t i = 0; i < reqs.size(); i++) {  // echo: the pick is derived from the request id, so hand-back mix-ups show
      ScoredEndpoint se;
      se.Index = (int)(std::stoul(reqs[i].RequestId) % eps.size());
      se.Score = (double)std::stoul(reqs[i].RequestId);
      out[i].result.PrimaryProfileName = "default";
      out[i].result.ProfileResults["default"].TargetEndpoints.push_back(se);
    }
    --in_flight;
    return out;
  }
};

static int test_coalescer() {
  MockBackend be;
  BatchingScheduler<MockBackend> bs(&be, std::chrono::microseconds(300), 16);
  auto eps = std::make_shared<const std::vector<Endpoint>>(std::vector<Endpoint>(7));
  auto eps2 = std::make_shared<const std::vector<Endpoint>>(std::vector<Endpoint>(5));
  constexpr int T = 64, PER = 40;
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      for (int i = 0; i < PER; i++) {
        InferenceRequest r;
        const unsigned id = (unsigned)(t * 1000 + i);
        r.RequestId = std::to_string(id);
        const auto& e = (t % 8 == 7) ? eps2 : eps;  // a minority of callers hold another snapshot
        auto item = bs.Schedule(r, e);
        const auto& te = item.result.ProfileResults["default"].TargetEndpoints;
        if (!item.error.empty() || te.size() != 1 || te[0].Score != (double)id || te[0].Index != (int)(id % e->size())) bad++;
      }
    });
  for (auto& x : th) x.join();
  const CoalescerStats st = bs.stats();
  CHECK(bad == 0);
  CHECK(st.requests == (uint64_t)T * PER);
  CHECK(st.max_batch <= 16 && st.max_batch > 1);
  CHECK(st.batches < st.requests);           // requests really shared batches
  CHECK(be.max_in_flight == 1);              // one backend call at a time (the engine's contract)
  printf("coalescer: %llu requests in %llu batches (max %llu, %llu full), backend calls never overlapped\n",
         (unsigned long long)st.requests, (unsigned long long)st.batches, (unsigned long long)st.max_batch, (unsigned long long)st.full_batches);
  // a lone caller waits about one window, not forever
  const auto t0 = std::chrono::steady_clock::now();
  InferenceRequest r;
  r.RequestId = "5";
  auto item = bs.Schedule(r, eps);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  CHECK(item.error.empty() && ms < 50.0);
  return 0;
}

static int test_lora_labels() {
  Metrics m;
  // spec_test.go:360-361
  CHECK(PopulateLoRAMetrics(m, {{"running_lora_adapters", "lora1"}, {"max_lora", "2"}}) == 0);
  CHECK(m.ActiveModels.size() == 1 && m.ActiveModels.count("lora1") && m.WaitingModels.empty() && m.MaxActiveModels == 2);
  CHECK(PopulateLoRAMetrics(m, {{"running_lora_adapters", "lora2,lora3"}, {"max_lora", "4"}}) == 0);
  CHECK(m.ActiveModels.size() == 2 && m.ActiveModels.count("lora2") && m.ActiveModels.count("lora3") && m.MaxActiveModels == 4);
  // extractor_test.go:149-153: running lora1, waiting lora2
  CHECK(PopulateLoRAMetrics(m, {{"running_lora_adapters", "lora1"}, {"waiting_lora_adapters", "lora2"}, {"max_lora", "2"}}) == 0);
  CHECK(m.ActiveModels.count("lora1") && m.WaitingModels.count("lora2") && m.MaxActiveModels == 2);
  // addAdapters: spaces trimmed (strings.TrimSpace incl. U+00A0), empty fields skipped, duplicates collapse
  CHECK(PopulateLoRAMetrics(m, {{"running_lora_adapters", " a , ,b,\xC2\xA0" "c\t,,a"}, {"waiting_lora_adapters", ""}, {"max_lora", ""}}) == 0);
  CHECK(m.ActiveModels.size() == 3 && m.ActiveModels.count("a") && m.ActiveModels.count("b") && m.ActiveModels.count("c"));
  CHECK(m.WaitingModels.empty() && m.MaxActiveModels == 2);  // empty max_lora leaves the previous value (extractor.go:227)
  // strconv.Atoi errors are collected, the value is left alone
  CHECK(PopulateLoRAMetrics(m, {{"max_lora", "4x"}}) == 1 && m.MaxActiveModels == 2);
  CHECK(PopulateLoRAMetrics(m, {{"max_lora", " 4"}}) == 1);
  CHECK(PopulateLoRAMetrics(m, {{"max_lora", "+7"}}) == 0 && m.MaxActiveModels == 7);
  CHECK(PopulateLoRAMetrics(m, {{"max_lora", "-3"}}) == 0 && m.MaxActiveModels == -3);
  // dictionary + columns: an adapter in both maps counts twice, names beyond the dictionary still count in nmodels
  AdapterDictionary dict(2);
  Metrics e;
  e.ActiveModels = {{"x", 0}, {"y", 0}, {"z", 0}};
  e.WaitingModels = {{"x", 0}};
  e.MaxActiveModels = 5;
  uint64_t act[1], wai[1];
  int32_t nm, mx;
  PackLoraColumns(e, dict, 1, act, wai, &nm, &mx);
  CHECK(dict.size() == 2 && dict.Lookup("x") == 0 && dict.Lookup("y") == 1 && dict.Lookup("z") == -1);
  CHECK(act[0] == 3 && wai[0] == 1 && nm == 4 && mx == 5);
  printf("lora labels: ok\n");
  return 0;
}

static int test_topk() {
  // picker_test.go:43-110 shapes: single max; ties as a class; top-k order
  const double nan = std::numeric_limits<double>::quiet_NaN();
  const double s1[] = {10, 25, 15};
  auto t = TopK(s1, 3, 1);
  CHECK(t.size() == 1 && t[0].first == 1 && t[0].second == 25);
  const double s2[] = {50, 50, 30, nan, 50};
  t = TopK(s2, 5, 2);
  CHECK(t.size() == 2 && t[0].first == 0 && t[1].first == 1);  // tie class {0,1,4}: ascending index
  t = TopK(s2, 5, 10);
  CHECK(t.size() == 4 && t[3].first == 2);                     // the non-candidate (NaN) never appears
  const double s3[] = {20, 25, 30, 15};
  t = TopK(s3, 4, 3);
  CHECK(t[0].first == 2 && t[1].first == 1 && t[2].first == 0);
  printf("top-k: ok\n");
  return 0;
}

// C shim: SmallBatchCpu on arrays (adapters as dictionary ids -> names "a<id>") for the oracle comparison in pytest
extern "C" int sbc_schedule(int M, const double* kv, const long long* queue, const long long* running, const unsigned long long* act,
                            const unsigned long long* wai, const int* nmodels, const int* maxm, const long long* tokens, int n_scorers,
                            const int* kinds, const double* weights, int adapter, const unsigned* cand_mask, int* pick, double* score,
                            int* ties) {
  std::vector<Endpoint> eps((size_t)M);
  for (int m = 0; m < M; m++) {
    Metrics& x = eps[m].Metrics_;
    x.KVCacheUsagePercent = kv[m];
    x.WaitingQueueSize = (int)queue[m];
    x.RunningRequestsSize = running ? (int)running[m] : 0;
    int known = 0;
    for (int a = 0; a < 64; a++) {
      if (act && ((act[m] >> a) & 1)) x.ActiveModels["a" + std::to_string(a)] = 0, known++;
      if (wai && ((wai[m] >> a) & 1)) x.WaitingModels["a" + std::to_string(a)] = 0, known++;
    }
    for (int k = known; nmodels && k < nmodels[m]; k++) x.ActiveModels["oov" + std::to_string(k)] = 0;  // out-of-vocabulary adapters
    x.MaxActiveModels = maxm ? maxm[m] : 0;
    eps[m].InFlightTokens = tokens ? tokens[m] : -1;
  }
  SmallBatchCpu cpu;
  if (!cpu.Configure(std::vector<int32_t>(kinds, kinds + n_scorers), std::vector<double>(weights, weights + n_scorers))) return -1;
  std::vector<int> cand;
  if (cand_mask)
    for (int m = 0; m < M; m++)
      if ((cand_mask[m >> 5] >> (m & 31)) & 1u) cand.push_back(m);
  const CpuPick p = cpu.Schedule(eps, adapter >= 0 ? "a" + std::to_string(adapter) : std::string("unknown-model"), cand_mask ? &cand : nullptr);
  *pick = p.pick;
  *score = p.score;
  *ties = p.tie_count;
  return 0;
}

int main() {
  if (test_lora_labels()) return 1;
  if (test_topk()) return 1;
  if (test_coalescer()) return 1;
  printf("host logic: all ok\n");
  return 0;
}
