// coalescer.hpp — what sits between the Director and the engine, restated in C++ (no Go toolchain here):
//
//   BatchingScheduler   many concurrent Scheduler.Schedule() callers (one goroutine per request in the reference,
//                       pkg/epp/requestcontrol/director.go:68-70,211) become engine batches: a request waits at most
//                       `window` for company, a batch leaves early when it holds `max_batch` requests.  Same shape as the
//                       reference's own coalescer for the latency sidecar (sidecars/latencypredictorasync/coalescer.go:53-120:
//                       timer armed by the first submission, early dispatch on the row cap), but leader/follower instead of a
//                       dispatcher goroutine: the caller that opens a batch closes it, so an idle scheduler costs nothing.
// (the LoRA label parser, top-k and the small-batch CPU route live in host_eval.hpp)
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "epp_scheduler.hpp"

namespace epp {

// ------------------------------------------------------------------------------------------------------------------
// BatchingScheduler
// ------------------------------------------------------------------------------------------------------------------
struct CoalescerStats {
  uint64_t batches = 0, requests = 0, max_batch = 0, full_batches = 0;
};

// Backend: anything with  std::vector<Item> ScheduleBatch(const std::vector<InferenceRequest>&, const std::vector<Endpoint>&)
// where Item has `.result` and `.error` (epp::Scheduler::BatchItem).  One backend call at a time (the engine's contract).
template <class Backend, class Item = typename Backend::BatchItem>
class BatchingScheduler {
 public:
  BatchingScheduler(Backend* backend, std::chrono::microseconds window, int max_batch)
      : backend_(backend), window_(window), max_batch_(max_batch < 1 ? 1 : max_batch) {}

  // Scheduler.Schedule for one request; blocks until its batch has been scheduled.  All callers that share `endpoints`
  // (the same snapshot object) may share a batch; a caller with another snapshot closes the open batch first.
  Item Schedule(const InferenceRequest& request, const std::shared_ptr<const std::vector<Endpoint>>& endpoints) {
    std::unique_lock<std::mutex> lk(mu_);
    // join the open batch, or open one
    while (open_ && (open_->closed || open_->endpoints != endpoints || (int)open_->requests.size() >= max_batch_)) {
      if (!open_->closed && open_->endpoints != endpoints) {  // another snapshot: do not wait for the window
        open_->closed = true;
        open_->cv.notify_all();
      }
      turn_.wait(lk);
    }
    std::shared_ptr<Batch> b = open_;
    bool leader = false;
    if (!b) {
      b = open_ = std::make_shared<Batch>();
      b->endpoints = endpoints;
      leader = true;
    }
    const size_t slot = b->requests.size();
    b->requests.push_back(request);
    if ((int)b->requests.size() >= max_batch_) {
      b->closed = true;
      b->cv.notify_all();
    }
    if (leader) {
      const auto deadline = std::chrono::steady_clock::now() + window_;
      while (!b->closed && b->cv.wait_until(lk, deadline) != std::cv_status::timeout) {
      }
      b->closed = true;
      open_.reset();          // the next caller opens a new batch while this one runs
      turn_.notify_all();
      // one backend call at a time
      while (running_) idle_.wait(lk);
      running_ = true;
      lk.unlock();
      std::vector<Item> res;
      std::string fail;
      try {
        res = backend_->ScheduleBatch(b->requests, *b->endpoints);
      } catch (const std::exception& e) {
        fail = e.what();
      }
      lk.lock();
      running_ = false;
      idle_.notify_one();
      if (res.size() != b->requests.size()) {
        res.assign(b->requests.size(), Item{});
        for (auto& r : res) r.error = fail.empty() ? "scheduler backend returned a short batch" : fail;
      }
      b->results = std::move(res);
      b->done = true;
      stats_.batches++;
      stats_.requests += b->requests.size();
      stats_.max_batch = std::max<uint64_t>(stats_.max_batch, b->requests.size());
      if ((int)b->requests.size() >= max_batch_) stats_.full_batches++;
      b->cv.notify_all();
    } else {
      while (!b->done) b->cv.wait(lk);
    }
    return b->results[slot];
  }
  CoalescerStats stats() const {
    std::lock_guard<std::mutex> lk(mu_);
    return stats_;
  }

 private:
  struct Batch {
    std::shared_ptr<const std::vector<Endpoint>> endpoints;
    std::vector<InferenceRequest> requests;
    std::vector<Item> results;
    std::condition_variable cv;
    bool closed = false, done = false;
  };
  Backend* backend_;
  std::chrono::microseconds window_;
  int max_batch_;
  mutable std::mutex mu_;
  std::condition_variable turn_, idle_;
  std::shared_ptr<Batch> open_;
  bool running_ = false;
  CoalescerStats stats_;
};

}  // namespace epp
