// host_path.cu — product-owned HOST pieces of the path (no device code):
//   * hashPrompt (approximateprefix/hashing.go:34-98) on the host cores, for callers that ship block hashes instead of
//     prompt bytes across PCIe (eppscore_batch.hashes_in: 8x fewer bytes): a persistent worker pool, requests claimed in
//     chunks.  Same chained XXH64 as the device kernels (xxh64.cuh), checked against them and the oracle in the tests.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/eppscore.h"
#include "xxh64.cuh"

namespace {

using eppscore::xxh64_host;

// XXH64(block || LE64(prev)) — hashing.go:79-85
inline uint64_t link_host(const uint8_t* d, uint32_t n, uint64_t prev) {
  if (n + 8 <= 4096) {
    uint8_t buf[4096];
    memcpy(buf, d, n);
    memcpy(buf + n, &prev, 8);  // little-endian host (x86-64 / aarch64)
    return xxh64_host(buf, (size_t)n + 8, 0);
  }
  return eppscore::xxh64_link<false>(d, n, prev);
}

// The same link when the block size is a multiple of 32 (the default 64): the four-lane stripe state depends on the block's
// own bytes only, the chain value enters as the message's last 8 bytes — one tail round + avalanche (xxh64.cuh: the device
// kernels' hash_bodies / hash_chain split).  No copy of the block, 64-bit loads; the out-of-order core overlaps the next
// block's stripes with this block's serial tail.
inline uint64_t ld64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);  // little-endian host
  return v;
}
inline uint64_t link_host_stripes(const uint8_t* d, uint32_t n, uint64_t prev) {
  using namespace eppscore;
  uint64_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0 - XP1;
  for (uint32_t s = 0; s < n; s += 32) {
    v1 = xround(v1, ld64(d + s));
    v2 = xround(v2, ld64(d + s + 8));
    v3 = xround(v3, ld64(d + s + 16));
    v4 = xround(v4, ld64(d + s + 24));
  }
  return xchain_aligned(xfinish_lanes(v1, v2, v3, v4) + (uint64_t)(n + 8), prev);
}

int32_t hash_one(const uint8_t* p, int64_t len, uint64_t seed, int32_t bc, int32_t max_blocks, uint64_t* out) {
  if (bc <= 0 || len < bc) return 0;                       // hashing.go:51-60
  const int64_t cap = (int64_t)bc * max_blocks;
  if (len > cap) len = cap;                                // :62-65
  uint64_t prev = seed;
  int32_t n = 0;
  int64_t o = 0;
  if ((bc & 31) == 0) {
    for (; o + bc <= len; o += bc) out[n++] = prev = link_host_stripes(p + o, (uint32_t)bc, prev);   // :79-87
  } else {
    for (; o + bc <= len; o += bc) out[n++] = prev = link_host(p + o, (uint32_t)bc, prev);
  }
  if (o < len) out[n++] = link_host(p + o, (uint32_t)(len - o), prev);                     // trailing partial block :89-95
  return n;
}

// eight requests at a time, one per AVX-512 lane (host_hash_simd.cpp, compiled with -mavx512f -mavx512dq; used only when the
// CPU has both)
extern "C" void eppscore_host_hash8_avx512(const uint8_t* bytes, const int64_t* off, const uint64_t* seed, int32_t bc, int32_t nfull,
                                           uint64_t* hashes, const int64_t* row, uint64_t* last, int32_t groups);
extern "C" int32_t eppscore_host_hash8_compiled(void);

bool have_simd8() {
#if defined(__x86_64__)
  static const bool ok = eppscore_host_hash8_compiled() != 0 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
  return ok;
#else
  return false;
#endif
}

struct Job {
  int32_t R = 0;
  const uint8_t* bytes = nullptr;
  const int64_t* off = nullptr;
  const int32_t* len = nullptr;
  const uint64_t* seed = nullptr;
  int32_t bc = 0, max_blocks = 0, stride = 0;
  uint64_t* hashes = nullptr;
  uint16_t* nh = nullptr;
  std::atomic<int32_t> next{0};
  int32_t chunk = 64;
  bool simd8 = false;   // block size a multiple of 32 and the CPU has AVX-512 F + DQ
};

void run_job(Job* j) {
  for (;;) {
    const int32_t r0 = j->next.fetch_add(j->chunk, std::memory_order_relaxed);
    if (r0 >= j->R) return;
    const int32_t r1 = r0 + j->chunk < j->R ? r0 + j->chunk : j->R;
    int32_t r = r0;
    if (j->simd8) {
      const int32_t bc = j->bc;
      const int64_t cap = (int64_t)bc * j->max_blocks;
      // 32 requests (four independent groups of eight fill the multiplier pipeline) while they last, then 8
      for (int W = 32; W >= 8; W -= 24) {
        for (; r + W <= r1; r += W) {
          int64_t o8[32], l8[32], row8[32];
          uint64_t s8[32], last8[32];
          bool same = true;
          for (int k = 0; k < W; k++) {
            o8[k] = j->off[r + k];
            int64_t l = j->len ? (int64_t)j->len[r + k] : j->off[r + k + 1] - o8[k];
            if (l > cap) l = cap;                              // hashing.go:62-65
            l8[k] = l;
            s8[k] = j->seed ? j->seed[r + k] : 0;
            row8[k] = (int64_t)(r + k) * j->stride;
            same = same && l / bc == l8[0] / bc;
          }
          const int32_t nfull = (int32_t)(l8[0] / bc);
          if (!same || nfull == 0) break;                      // mixed lengths: narrower groups, then the scalar loop, take over
          eppscore_host_hash8_avx512(j->bytes, o8, s8, bc, nfull, j->hashes, row8, last8, W / 8);
          for (int k = 0; k < W; k++) {
            uint64_t* out = j->hashes + (size_t)row8[k];
            int32_t n = nfull;
            const int64_t done = (int64_t)nfull * bc;
            if (done < l8[k]) out[n++] = link_host(j->bytes + o8[k] + done, (uint32_t)(l8[k] - done), last8[k]);  // partial block :89-95
            for (int32_t i = n; i < j->stride; i++) out[i] = 0;
            j->nh[r + k] = (uint16_t)n;
          }
        }
      }
    }
    for (; r < r1; r++) {
      const int64_t o = j->off[r];
      const int64_t l = j->len ? (int64_t)j->len[r] : j->off[r + 1] - o;
      uint64_t* out = j->hashes + (size_t)r * j->stride;
      const int32_t n = hash_one(j->bytes + o, l, j->seed ? j->seed[r] : 0, j->bc, j->max_blocks, out);
      for (int32_t i = n; i < j->stride; i++) out[i] = 0;
      j->nh[r] = (uint16_t)n;
    }
  }
}

class Pool {
 public:
  void run(Job* j, int n_threads) {
    std::unique_lock<std::mutex> call(call_mu_);
    {
      std::unique_lock<std::mutex> lk(mu_);
      while ((int)th_.size() < n_threads - 1) {
        const int id = (int)th_.size();
        th_.emplace_back([this, id] { worker(id); });
      }
      job_ = j;
      want_ = n_threads - 1;
      running_ = want_;
      gen_++;
    }
    go_.notify_all();
    run_job(j);  // the calling thread works too
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return running_ == 0; });
  }
  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      quit_ = true;
      gen_++;
    }
    go_.notify_all();
    for (auto& t : th_) t.join();
  }

 private:
  void worker(int id) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      go_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (quit_) return;
      if (id >= want_) continue;
      Job* j = job_;
      lk.unlock();
      run_job(j);
      lk.lock();
      if (--running_ == 0) done_.notify_one();
    }
  }
  std::mutex mu_, call_mu_;
  std::condition_variable go_, done_;
  std::vector<std::thread> th_;
  Job* job_ = nullptr;
  int want_ = 0, running_ = 0;
  uint64_t gen_ = 0;
  bool quit_ = false;
};

Pool& pool() {
  static Pool* p = new Pool();  // leaked on purpose: worker threads must not be joined from a static destructor at exit
  return *p;
}

}  // namespace

extern "C" int32_t eppscore_hash_prompts_host(int32_t R, const uint8_t* prompt_bytes, const int64_t* prompt_off,
                                              const int32_t* prompt_len, const uint64_t* model_seed, int32_t block_chars,
                                              int32_t max_blocks, uint64_t* hashes_out, int32_t hash_stride,
                                              uint16_t* n_hashes_out, int32_t n_threads) {
  if (R < 0 || (R > 0 && (!prompt_bytes || !prompt_off || !hashes_out || !n_hashes_out))) return EPPSCORE_ERR_INVALID;
  if (block_chars <= 0) block_chars = 64;
  if (max_blocks <= 0) max_blocks = 256;
  if (max_blocks > EPPSCORE_MAX_BLOCKS || hash_stride < max_blocks) {
    // a prompt can produce max_blocks hashes (the trailing partial block included in the cap): the row must hold them
    if (max_blocks > EPPSCORE_MAX_BLOCKS) return EPPSCORE_ERR_CAPACITY;
    return EPPSCORE_ERR_INVALID;
  }
  if (R == 0) return EPPSCORE_OK;
  if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if (n_threads > R) n_threads = R;
  Job j;
  j.R = R;
  j.bytes = prompt_bytes;
  j.off = prompt_off;
  j.len = prompt_len;
  j.seed = model_seed;
  j.bc = block_chars;
  j.max_blocks = max_blocks;
  j.stride = hash_stride;
  j.hashes = hashes_out;
  j.nh = n_hashes_out;
  j.simd8 = (block_chars & 31) == 0 && have_simd8();
  j.chunk = R / (n_threads * 8) > 16 ? (R / (n_threads * 8) < 256 ? R / (n_threads * 8) : 256) : 16;
  if (n_threads == 1) run_job(&j);
  else pool().run(&j, n_threads);
  return EPPSCORE_OK;
}
