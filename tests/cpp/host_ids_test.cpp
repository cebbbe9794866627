// host_ids_test.cpp — stable endpoint ids in the C++ host mirror (epp_scheduler.hpp), CPU only, linked against
// tests/cpp/oracle_backed_abi.cpp.  The prefix index refers to endpoints by id (the reference's ServerID is the pod's
// NamespacedName, approximateprefix/indexer.go:34-35), while the candidate list a Schedule() call receives has neither a stable
// order nor a stable membership (datastore PodList over a sync.Map; subsetting, director candidates.go:98).  What must hold:
// prefix affinity follows the endpoint's NAME across reordered / extended / reduced lists, a removed pod loses its history
// (indexer.RemovePod, indexer.go:167-182) and its id is recycled, and a list that is exactly ids 0..M-1 is not remapped.
#include <cstdio>
#include <string>
#include <vector>

#include "../../gateway-api-inference-extension_b200/host/epp_scheduler.hpp"

using namespace epp;

static int g_fail = 0;
#define CHECK(cond)                                               \
  do {                                                            \
    if (!(cond)) {                                                \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
      g_fail++;                                                   \
    }                                                             \
  } while (0)

static Endpoint Pod(const std::string& name, double kv) {
  Metrics m;
  m.KVCacheUsagePercent = kv;
  return NewEndpoint(name, m);
}
static std::string Name(const Scheduler::BatchItem& it) {
  return it.result.ProfileResults.at("default").TargetEndpoints[0].Endpoint_->GetMetadata()->NamespacedName_.Name;
}
static const ScoredEndpoint& Top(const Scheduler::BatchItem& it) { return it.result.ProfileResults.at("default").TargetEndpoints[0]; }

int main() {
  try {
    SchedulerConfig c;
    c.CpuBatchThreshold = 0;
    // kv decides among endpoints without a cached prefix; one matched block of two outweighs any kv difference
    c.Profile.WithScorers({NewWeightedScorer(std::make_shared<PrefixCacheScorer>(), 10), NewWeightedScorer(std::make_shared<KVCacheUtilizationScorer>(), 1)})
        .WithPicker(MaxScorePicker{});
    c.Prefix.AutoTune = false;
    c.Prefix.BlockSizeTokens = 1;  // 4 characters per block
    c.MaxEndpoints = 8;
    c.PrefixCapacity = 256;
    Scheduler s(c);

    // 1. first list: ids are positions, nothing is remapped; podB has the emptiest cache and takes "aaaa...."
    std::vector<Endpoint> l1 = {Pod("podA", 0.5), Pod("podB", 0.1), Pod("podC", 0.3)};
    auto r1 = s.ScheduleBatch({InferenceRequest{"1", "m", "aaaaxxxx", ""}}, l1);
    CHECK(r1[0].error.empty() && Name(r1[0]) == "podB" && Top(r1[0]).Index == 1);
    CHECK(s.RemappedBatches() == 0 && s.ServerId("/podA") == 0 && s.ServerId("/podB") == 1 && s.ServerId("/podC") == 2);
    s.PreRequest(r1);

    // 2. the same pods in another order plus a newcomer in front: the affinity follows the NAME.  podB's kv is now the
    //    worst, so only the cached block can make it win; its position is 3.
    std::vector<Endpoint> l2 = {Pod("podD", 0.0), Pod("podC", 0.3), Pod("podA", 0.5), Pod("podB", 0.9)};
    auto r2 = s.ScheduleBatch({InferenceRequest{"2", "m", "aaaayyyy", ""}}, l2);
    CHECK(r2[0].error.empty() && Name(r2[0]) == "podB" && Top(r2[0]).Index == 3);
    CHECK(Top(r2[0]).Score == 10 * 0.5 + (1 - 0.9));  // one block of two matched + podB's own kv score
    CHECK(Top(r2[0]).TieCount == 1);
    CHECK(s.RemappedBatches() == 1 && s.ServerId("/podD") == 3);
    s.PreRequest(r2);  // recorded under podB's id although it sat at position 3

    // 3. a subset without podB (subsetting): ids 1 is a hole — it must not be picked, and the others score on kv alone
    std::vector<Endpoint> l3 = {Pod("podC", 0.3), Pod("podA", 0.5)};
    auto r3 = s.ScheduleBatch({InferenceRequest{"3", "m", "aaaazzzz", ""}}, l3);
    CHECK(r3[0].error.empty() && Name(r3[0]) == "podC" && Top(r3[0]).Index == 0);
    CHECK(Top(r3[0]).Score == 1 - 0.3);
    // several requests at once, each handed its own result
    auto r3b = s.ScheduleBatch({InferenceRequest{"3a", "m", "qqqq", ""}, InferenceRequest{"3b", "m", "aaaa", ""}}, l3);
    CHECK(Name(r3b[0]) == "podC" && Name(r3b[1]) == "podC");

    // 4. top-k over a remapped list returns positions of THIS list, best first
    {
      SchedulerConfig ck = c;
      MaxScorePicker pk;
      pk.MaxNumOfEndpoints = 2;
      ck.Profile = SchedulerProfile{};
      ck.Profile.WithScorers({NewWeightedScorer(std::make_shared<KVCacheUtilizationScorer>(), 1)}).WithPicker(pk);
      Scheduler sk(ck);
      std::vector<Endpoint> a = {Pod("p0", 0.9), Pod("p1", 0.2), Pod("p2", 0.4)};
      sk.ScheduleBatch({InferenceRequest{"k0", "m", "", ""}}, a);
      std::vector<Endpoint> b = {Pod("p2", 0.4), Pod("p0", 0.9), Pod("p1", 0.2)};  // rotated
      auto rk = sk.ScheduleBatch({InferenceRequest{"k1", "m", "", ""}}, b);
      const auto& te = rk[0].result.ProfileResults.at("default").TargetEndpoints;
      CHECK(te.size() == 2 && te[0].Index == 2 && te[1].Index == 0);  // p1 (0.8) then p2 (0.6), as positions of list b
      CHECK(te[0].Score == 1 - 0.2 && te[1].Score == 1 - 0.4);
      CHECK(sk.RemappedBatches() == 1);
    }

    // 5. RemovePod: podB's history is gone and its id is free; a new pod takes the id without inheriting anything
    s.RemovePod("/podB");
    CHECK(s.ServerId("/podB") == -1);
    std::vector<Endpoint> l5 = {Pod("podE", 0.9), Pod("podA", 0.5), Pod("podC", 0.3), Pod("podD", 0.6)};
    auto r5 = s.ScheduleBatch({InferenceRequest{"5", "m", "aaaawwww", ""}}, l5);
    CHECK(s.ServerId("/podE") == 1);                     // the recycled id
    CHECK(r5[0].error.empty() && Name(r5[0]) == "podC");  // nobody holds "aaaa" any more: kv decides
    CHECK(Top(r5[0]).Score == 1 - 0.3);
    // and podB coming back is a new server as well
    std::vector<Endpoint> l6 = {Pod("podB", 0.95), Pod("podA", 0.5)};
    auto r6 = s.ScheduleBatch({InferenceRequest{"6", "m", "aaaavvvv", ""}}, l6);
    CHECK(Name(r6[0]) == "podA" && s.ServerId("/podB") == 4);

    // 6. the same server twice in one list: positions are used for that batch (no id can tell them apart)
    std::vector<Endpoint> l7 = {Pod("dup", 0.7), Pod("dup", 0.2)};
    auto r7 = s.ScheduleBatch({InferenceRequest{"7", "m", "", ""}}, l7);
    CHECK(r7[0].error.empty() && Top(r7[0]).Index == 1);

    // 7. more servers than the engine was sized for is an error, not a silent overwrite
    bool threw = false;
    try {
      std::vector<Endpoint> many;
      for (int i = 0; i < 9; i++) many.push_back(Pod("extra" + std::to_string(i), 0.5));
      s.ScheduleBatch({InferenceRequest{"8", "m", "", ""}}, many);
    } catch (const SchedulingError&) {
      threw = true;
    }
    CHECK(threw);
  } catch (const std::exception& e) {
    std::printf("FAIL exception: %s\n", e.what());
    return 2;
  }
  if (g_fail) {
    std::printf("%d check(s) failed\n", g_fail);
    return 1;
  }
  std::printf("host ids: all checks passed\n");
  return 0;
}
