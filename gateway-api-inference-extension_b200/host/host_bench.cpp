// host_bench.cpp — timings of the host layer above the C ABI, printed as ONE JSON object on stdout.  Built by
// build.build_host_bench(), run by bench.py on rank 0 (needs a B200).  Nothing here is a parity check: host_test.cpp holds those.
//   config_A        BASELINE config A (1 request x 4 pods, queue scorer only; reference shape scheduler_test.go:42-159 run as a
//                   benchmark) through the PRODUCT: Scheduler.Schedule with the small-batch host route (SmallBatchCpu), and the same
//                   call forced onto the GPU.
//   coalescer       per-request latency of Scheduler.Schedule through BatchingScheduler (coalescer.hpp) with `callers` concurrent
//                   caller threads (the Director's goroutines, director.go:68-70,211) for several windows.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "coalescer.hpp"
#include "epp_scheduler.hpp"

using namespace epp;
using Clock = std::chrono::steady_clock;

static double us_since(Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); }

static double pct(std::vector<double>& v, double p) {
  if (v.empty()) return 0;
  std::sort(v.begin(), v.end());
  size_t i = (size_t)(p * (v.size() - 1));
  return v[i];
}

static std::vector<Endpoint> MakeEndpoints(int m) {
  std::vector<Endpoint> eps;
  for (int i = 0; i < m; i++) {
    Metrics x;
    x.WaitingQueueSize = (i * 7) % 13;
    x.KVCacheUsagePercent = ((i * 37) % 100) / 100.0;
    x.MaxActiveModels = 4;
    x.ActiveModels["adapter-" + std::to_string(i % 8)] = 1;
    eps.push_back(NewEndpoint("pod" + std::to_string(i), x));
  }
  return eps;
}

static void ConfigA(bool first) {
  for (int pass = 0; pass < 2; pass++) {
    SchedulerConfig c;
    c.CpuBatchThreshold = pass == 0 ? 8 : 0;
    c.Profile.WithScorers({NewWeightedScorer(std::make_shared<QueueScorer>(), 1)}).WithPicker(MaxScorePicker{});
    c.MaxEndpoints = 8;
    c.PrefixCapacity = 1 << 10;
    Scheduler s(c);
    auto eps = MakeEndpoints(4);
    InferenceRequest rq{"id", "m", "", ""};
    const int n = pass == 0 ? 200000 : 3000;
    for (int i = 0; i < 200; i++) s.Schedule(rq, eps);
    std::vector<double> lat;
    lat.reserve(n);
    auto t0 = Clock::now();
    for (int i = 0; i < n; i++) {
      auto t1 = Clock::now();
      auto r = s.Schedule(rq, eps);
      lat.push_back(us_since(t1));
      if (r.ProfileResults.empty()) std::abort();
    }
    const double total = us_since(t0);
    std::printf("%s\"%s\": {\"calls\": %d, \"ns_per_call\": %.1f, \"p50_us\": %.3f, \"p99_us\": %.3f, \"cpu_routed\": %llu}",
                (first && pass == 0) ? "" : ", ", pass == 0 ? "config_A_host_route" : "config_A_gpu_route", n, 1e3 * total / n,
                pct(lat, 0.5), pct(lat, 0.99), (unsigned long long)s.CpuRoutedRequests());
  }
}

// `callers` threads, each Schedule()s `per_caller` requests back to back (closed loop) through the front.
static void Coalescer(int callers, int per_caller, int window_us, int max_batch, int m, int prompt_bytes, bool first) {
  SchedulerConfig c;
  c.CpuBatchThreshold = 0;
  c.Profile.WithScorers({NewWeightedScorer(std::make_shared<KVCacheUtilizationScorer>(), 1),
                         NewWeightedScorer(std::make_shared<QueueScorer>(), 1),
                         NewWeightedScorer(std::make_shared<PrefixCacheScorer>(), 1),
                         NewWeightedScorer(std::make_shared<LoraAffinityScorer>(), 1)})
      .WithPicker(MaxScorePicker{});
  c.MaxEndpoints = m;
  c.PrefixCapacity = 1 << 16;
  Scheduler s(c);
  auto eps = std::make_shared<const std::vector<Endpoint>>(MakeEndpoints(m));
  BatchingScheduler<Scheduler> front(&s, std::chrono::microseconds(window_us), max_batch);
  std::vector<std::vector<double>> lat((size_t)callers);
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  auto t0 = Clock::now();
  for (int t = 0; t < callers; t++)
    th.emplace_back([&, t] {
      std::string prompt((size_t)prompt_bytes, 'a');
      for (int i = 0; i < per_caller; i++) {
        for (size_t k = 0; k < prompt.size(); k += 61) prompt[k] = (char)('a' + (t * 31 + i * 7 + k) % 26);
        InferenceRequest rq{std::to_string(t * 100000 + i), "adapter-" + std::to_string((t + i) % 8), prompt, ""};
        auto t1 = Clock::now();
        auto item = front.Schedule(rq, eps);
        lat[(size_t)t].push_back(us_since(t1));
        if (!item.error.empty()) bad++;
      }
    });
  for (auto& x : th) x.join();
  const double total = us_since(t0);
  std::vector<double> all;
  for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
  const CoalescerStats st = front.stats();
  std::printf("%s{\"callers\": %d, \"window_us\": %d, \"max_batch\": %d, \"endpoints\": %d, \"prompt_bytes\": %d, \"requests\": %llu, "
              "\"engine_batches\": %llu, \"mean_batch\": %.2f, \"largest_batch\": %llu, \"p50_us\": %.1f, \"p99_us\": %.1f, "
              "\"requests_per_s\": %.0f, \"errors\": %d}",
              first ? "" : ", ", callers, window_us, max_batch, m, prompt_bytes, (unsigned long long)st.requests,
              (unsigned long long)st.batches, st.batches ? (double)st.requests / st.batches : 0.0, (unsigned long long)st.max_batch,
              pct(all, 0.5), pct(all, 0.99), 1e6 * all.size() / total, bad.load());
}

int main(int argc, char** argv) {
  const int callers = argc > 1 ? std::atoi(argv[1]) : 64;
  const int per_caller = argc > 2 ? std::atoi(argv[2]) : 200;
  try {
    std::printf("{");
    ConfigA(true);
    std::printf(", \"coalescer\": [");
    const int windows[] = {0, 50, 200, 1000};
    bool first = true;
    for (int w : windows) {
      Coalescer(callers, per_caller, w, 4096, 256, 512, first);
      first = false;
    }
    std::printf("]}\n");
  } catch (const std::exception& e) {
    std::printf("\n{\"error\": \"%s\"}\n", e.what());
    return 2;
  }
  return 0;
}
