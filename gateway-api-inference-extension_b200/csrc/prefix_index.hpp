// prefix_index.hpp — host-side owner of the device-resident prefix index (prefix_table.cuh, prefix_index.cu).
// Replaces approximateprefix/indexer.go's `indexer` object: the host holds no copy of hashToPods or of the LRUs, it
// sizes the device arrays and launches the kernels that maintain them.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "prefix_table.cuh"

namespace eppscore {

struct IndexStats {
  int64_t used, live, capacity, ovf_rows, table_bytes, lru_bytes, lru_entries, rebuilds;
  uint32_t error;
};

class DeviceIndex {
 public:
  // n_endpoints: endpoint ids are < n_endpoints (the engine's padded capacity); row_words = n_endpoints / 32;
  // capacity_hashes: initial table capacity (it is rebuilt / doubled on demand); lru_max: largest LRU size the engine
  // must support (0: the default size) — it fixes the per-endpoint region sizes at the first Add.
  DeviceIndex(int32_t n_endpoints, int32_t row_words, int64_t capacity_hashes, int32_t default_lru, int32_t lru_max);
  ~DeviceIndex();
  DeviceIndex(const DeviceIndex&) = delete;
  DeviceIndex& operator=(const DeviceIndex&) = delete;

  cudaError_t init(cudaStream_t s);
  const TableView* view() const { return d_tv_; }  // device pointer, stable for the engine's lifetime

  // PreRequest for R requests (device arrays), endpoints [ep_first, ep_first + n_eps).  max_touches: an upper bound on
  // Σ n_hashes over the requests (<= 0: R * stride).  cap_req_host: optional [n_endpoints] CacheNumBlocks.
  cudaError_t commit(int32_t R, const int32_t* pick, const uint64_t* hashes, const uint16_t* n_hashes, int32_t stride,
                     const int32_t* cap_req_host, int32_t single_cap, int32_t ep_first, int32_t n_eps, int64_t max_touches,
                     cudaStream_t s);
  cudaError_t apply(int64_t n, const uint64_t* hash, const int32_t* endpoint, const uint8_t* op, cudaStream_t s);  // device arrays
  cudaError_t remove_endpoint(int32_t p, cudaStream_t s);
  cudaError_t lru_keys(int32_t p, uint64_t* out_host, int32_t cap, int32_t* len, cudaStream_t s);
  cudaError_t get(uint64_t h, uint32_t* bits_host, int32_t words, int32_t* count, cudaStream_t s);
  cudaError_t stats(IndexStats* out, cudaStream_t s);
  uint64_t launches() const { return launches_; }
  uint32_t max_cap() const { return max_cap_; }

 private:
  cudaError_t refresh();
  cudaError_t snapshot_after(cudaStream_t s);
  cudaError_t ensure_room(int64_t touches, cudaStream_t s);
  cudaError_t ensure_lru(uint32_t want_cap, cudaStream_t s);
  int64_t lru_bytes() const;

  int32_t n_endpoints_, row_words_, default_lru_;
  uint32_t max_cap_ = 0;
  int64_t init_capacity_ = 0;
  TableView tv_{};          // host copy: pointers + the counters as of the last snapshot
  LruView lv_{};
  TSlot* d_slots_ = nullptr;
  uint32_t* d_ovf_rows_ = nullptr;
  uint32_t* d_ovf_free_ = nullptr;
  TableView* d_tv_ = nullptr;
  LruView* d_lv_ = nullptr;
  LruDesc* d_desc_ = nullptr;
  LruEntry* d_maps_ = nullptr;
  uint64_t* d_logs_ = nullptr;
  int32_t* d_capreq_ = nullptr;
  void* d_scratch_ = nullptr;
  TableView* h_snap_ = nullptr;  // pinned
  cudaEvent_t ev_snap_ = nullptr;
  bool snap_pending_ = false;
  uint64_t launches_ = 0;
  int64_t rebuilds_ = 0;
};

}  // namespace eppscore
