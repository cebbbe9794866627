// capi.cu — the C ABI of libeppscore.so (include/eppscore.h): engine object, snapshot ingestion,
// the batched Schedule call, the prefix index plumbing.  Host code only orchestrates; all path
// arithmetic runs in the kernels of kernels.cu.  There is deliberately no CPU fallback here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/eppscore.h"
#include "kernels.cuh"
#include "prefix_index.hpp"
#include "xxh64.cuh"

using namespace eppscore;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = n + n / 8;  // a little slack so slowly growing batches do not reallocate every call
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct PlanSet {
  Plan plan;
  const double* term_ptr[kMaxSteps];
};

}  // namespace

struct eppscore_engine {
  int device = 0;
  int sm_count = 148;
  eppscore_config cfg{};
  Geo geo{};
  int32_t A_cap = 64;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_snapshot = nullptr, ev_table = nullptr, ev_caller = nullptr;
  cudaStream_t snapshot_stream = nullptr;        // stream the last snapshot preparation ran on
  unsigned long long snapshot_capture_id = 0;    // != 0: ev_snapshot was recorded inside that CUDA-graph capture
  std::string err;
  uint64_t launches = 0;
  bool force_generic = false;  // EPPSCORE_FORCE_GENERIC=1 / eppscore_set_debug(1): skip the specialised kernels
  bool use_pdl = true;          // eppscore_set_debug(4): programmatic dependent launch between the kernels of a batch
  int32_t hash_stage_mask = 19;  // eppscore_set_debug(2): profiling only (see include/eppscore.h)

  // snapshot
  bool have_snapshot = false;
  int32_t M = 0, A = 0, lora_words = 0;
  uint64_t epoch = 0;
  DevBuf raw_kv, raw_queue, raw_running, raw_act, raw_wait, raw_nmodels, raw_max, raw_col[4];
  DevBuf raw_min_tpot, raw_dispatched, raw_prefill, raw_tokens, lat_ep, qhdr[2], qbucket[2];
  eppscore_latency_params lat_params{};  // pending: applied by the next set_snapshot
  LatArgs lat_args{};                    // what the current snapshot was prepared with
  bool have_col[4] = {false, false, false, false};
  bool have_running = false;
  const int64_t* cur_queue = nullptr;    // raw WaitingQueueSize / RunningRequestsSize of the current snapshot
  const int64_t* cur_running = nullptr;  // (engine copy for host snapshots, the caller's buffer for device ones)
  DevBuf term[kMaxSteps], fold_unmasked, fold_masked, cls_lo, cls_hi, summ, tiemask, prefix_lut2d;
  bool have_lut2d = false;
  PlanSet plan_unmasked{}, plan_masked{};

  // prefix index: hashToPods + the per-endpoint LRUs, both device-resident (prefix_index.hpp); mutated on `stream`
  std::unique_ptr<DeviceIndex> index;
  cudaEvent_t ev_sched = nullptr;   // last device-location schedule on a caller stream (index mutations wait for it)
  bool sched_pending = false;
  DevBuf c_pick, c_hash, c_nh, c_ep, c_op;  // staging for the host-array forms of commit_picks / prefix_add / prefix_apply

  // scratch for host-location batches and internal hashes: two sets, so the chunks of a host batch can alternate
  // between two streams (H2D of chunk i+1 under the kernels and the D2H of chunk i)
  struct Scratch {
    DevBuf s_prompts, s_off, s_len, s_seed, s_hashes, s_nh, s_adapter, s_mask, s_dense, s_dtotal, s_pick, s_score,
        s_tie, s_match, s_total, s_scores, s_intok, s_tslo, s_pslo, s_pred, s_fields, s_fmask;
    void release() {
      DevBuf* all[] = {&s_prompts, &s_off, &s_len, &s_seed, &s_hashes, &s_nh, &s_adapter, &s_mask, &s_dense, &s_dtotal, &s_pick,
                       &s_score, &s_tie, &s_match, &s_total, &s_scores, &s_intok, &s_tslo, &s_pslo, &s_pred, &s_fields, &s_fmask};
      for (DevBuf* b : all) b->release();
    }
  } sc[4];
  // pinned staging for the small per-request results of a host batch: a D2H copy into the caller's (usually pageable)
  // arrays would block the host after every chunk and serialise the pipeline
  void* h_res = nullptr;
  size_t h_res_bytes = 0;
  cudaStream_t stream2 = nullptr;   // second copy/compute stream of the chunked host path
  int32_t host_chunk = 8192;        // requests per chunk of a host-location batch (eppscore_set_debug key 3)
  // device-location batches cut into `dev_split` slices over `dev_streams` streams (the caller's + internal ones): the
  // HBM-bound body hashing of slice i+1 runs under the latency-bound chain and probe kernels of slice i
  // (eppscore_set_debug keys 5 and 6)
  int32_t dev_split = 1, dev_streams = 2;
  cudaStream_t split_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
};
using Scratch = eppscore_engine::Scratch;

namespace {

int32_t fail(eppscore_engine* e, int32_t code, const std::string& msg) {
  if (e) e->err = msg;
  else g_create_error = msg;
  return code;
}
int32_t cuda_fail(eppscore_engine* e, cudaError_t c, const char* what) {
  return fail(e, EPPSCORE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(c));
}
#define CK(e, call)                                            \
  do {                                                         \
    cudaError_t _c = (call);                                   \
    if (_c != cudaSuccess) return cuda_fail((e), _c, #call);   \
  } while (0)

bool is_endpoint_term_kind(int k, bool masked) {
  if (k == EPPSCORE_SCORER_KV_CACHE || k == EPPSCORE_SCORER_TOKEN_LOAD) return true;
  if (k >= EPPSCORE_SCORER_ENDPOINT_COL0 && k < EPPSCORE_SCORER_ENDPOINT_COL0 + EPPSCORE_MAX_ENDPOINT_COLS) return true;
  if (!masked && (k == EPPSCORE_SCORER_QUEUE || k == EPPSCORE_SCORER_RUNNING)) return true;
  return false;
}
double clamp01_host(double s) { return s < 0 ? 0 : (s > 1 ? 1 : s); }

bool cfg_has(const eppscore_config& c, int kind) {
  for (int i = 0; i < c.n_scorers; i++)
    if (c.scorer_kind[i] == kind) return true;
  return false;
}

// Translate the scorer list into kernel steps.  Leading request-independent scorers are folded into
// one precomputed array ((0.0+t0)+t1)+… — exact, because the kernel adds in the same order.
void build_plan(eppscore_engine* e, bool masked, PlanSet* ps) {
  Plan& p = ps->plan;
  memset(&p, 0, sizeof(p));
  const eppscore_config& c = e->cfg;
  int lead = 0;
  while (lead < c.n_scorers && is_endpoint_term_kind(c.scorer_kind[lead], masked)) lead++;
  int nt = 0, ns = 0;
  if (lead > 0) {
    ps->term_ptr[nt] = (masked ? e->fold_masked : e->fold_unmasked).as<double>();
    p.kind[ns] = STEP_EP_TERM;
    p.arg[ns] = nt++;
    ns++;
  }
  for (int s = lead; s < c.n_scorers; s++) {
    const int k = c.scorer_kind[s];
    const double w = c.scorer_weight[s];
    p.weight[ns] = w;
    if (is_endpoint_term_kind(k, masked)) {
      ps->term_ptr[nt] = e->term[s].as<double>();
      p.kind[ns] = STEP_EP_TERM;
      p.arg[ns] = nt++;
    } else if (k == EPPSCORE_SCORER_QUEUE || k == EPPSCORE_SCORER_RUNNING) {
      p.kind[ns] = STEP_MINMAX;
      p.arg[ns] = k == EPPSCORE_SCORER_QUEUE ? 0 : 1;
    } else if (k == EPPSCORE_SCORER_PREFIX) {
      p.kind[ns] = STEP_PREFIX;
    } else if (k == EPPSCORE_SCORER_LATENCY) {
      p.kind[ns] = STEP_LATENCY;
    } else if (k == EPPSCORE_SCORER_LORA) {
      p.kind[ns] = STEP_LORA;
      static const double cls_score[4] = {0.0, 0.6, 0.8, 1.0};  // lora_affinity.go:84-99
      for (int q = 0; q < 4; q++) p.lora_term[ns][q] = clamp01_host(cls_score[q]) * w;
    } else {
      p.kind[ns] = STEP_PAIR;
      p.arg[ns] = k - EPPSCORE_SCORER_PAIR_COL0;
    }
    ns++;
  }
  p.n_steps = ns;
  p.n_terms = nt;
  p.tie_mode = c.tie_mode;
  p.seed_lo = (uint32_t)c.tie_seed;
  p.seed_hi = (uint32_t)(c.tie_seed >> 32);
  // packed kind sequence: selects the template-specialised streaming kernels (score_dense.cu)
  p.seq = 0;
  if (ns <= 7)
    for (int i = 0; i < ns; i++) p.seq |= (uint32_t)(p.kind[i] + 1) << (4 * i);
  // sparse fast path (pick_sparse.cu): unmasked, only E/P/L steps, at most one prefix step and its
  // weight >= 0 (monotonicity: a positive match can only raise an endpoint's score)
  int n_prefix = 0;
  bool ok = !masked;
  for (int i = 0; i < ns; i++) {
    if (p.kind[i] == STEP_PREFIX) {
      n_prefix++;
      if (!(p.weight[i] >= 0.0)) ok = false;
    } else if (p.kind[i] != STEP_EP_TERM && p.kind[i] != STEP_LORA) {
      ok = false;
    }
  }
  p.pick_mode = c.pick_mode;
  p.sparse_ok = (ok && n_prefix <= 1 && c.pick_mode == EPPSCORE_PICK_MAX_SCORE) ? 1 : 0;
}

// Index mutations run on the engine stream.  They must not overtake scoring kernels that other streams launched
// earlier against the same table, and later scoring on other streams waits for ev_table.
int32_t index_begin(eppscore_engine* e) {
  if (e->sched_pending) {
    CK(e, cudaStreamWaitEvent(e->stream, e->ev_sched, 0));
    e->sched_pending = false;
  }
  return EPPSCORE_OK;
}
int32_t index_end(eppscore_engine* e) {
  CK(e, cudaEventRecord(e->ev_table, e->stream));
  return EPPSCORE_OK;
}

struct DevBatch {  // all device pointers
  int32_t R;
  int64_t request_base;
  const uint8_t* prompt_bytes;
  const int64_t* prompt_off;
  const int32_t* prompt_len;
  const uint64_t* model_seed;
  const uint64_t* hashes_in;
  const uint16_t* n_hashes_in;
  int32_t hash_stride;
  int32_t block_chars, max_blocks;
  const int32_t* adapter_id;
  const uint32_t* cand_mask;
  const float* dense_feat;
  const uint16_t* dense_total;
  const int32_t* input_tokens;
  const double* ttft_slo;
  const double* tpot_slo;
  double* pred_out;
  uint32_t* filter_mask_out;
  int32_t* pick;
  double* pick_score;
  int32_t* tie_count;
  uint16_t* match_blocks;
  uint16_t* total_blocks;
  uint64_t* hashes_out;
  double* scores_out;
};

// The hot path on device-resident buffers: [hash kernel] + score/pick kernel, asynchronous on `s`.
int32_t schedule_device(eppscore_engine* e, const DevBatch& b, cudaStream_t s, Scratch& sc) {
  if (!e->have_snapshot) return fail(e, EPPSCORE_ERR_NO_SNAPSHOT, "schedule_batch before set_snapshot");
  if (b.R <= 0) return EPPSCORE_OK;
  if (!b.pick || !b.pick_score || !b.tie_count) return fail(e, EPPSCORE_ERR_INVALID, "pick/pick_score/tie_count required");
  const bool dense = b.dense_feat != nullptr;
  const bool masked = b.cand_mask != nullptr;
  const PlanSet& ps = masked ? e->plan_masked : e->plan_unmasked;

  // Ordering against the snapshot preparation and the table flush, which may have run on other streams.
  // The hash kernels depend on neither, so the snapshot wait is issued only before the pick kernel: a caller
  // may prepare the snapshot on a second stream concurrently with the hashing of the batch.
  // Under CUDA-graph capture a stream may only wait on events recorded inside the SAME capture.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  unsigned long long cap_id = 0;
  cudaStreamGetCaptureInfo(s, &cap, &cap_id);
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (!capturing && s != e->stream) CK(e, cudaStreamWaitEvent(s, e->ev_table, 0));
  const bool wait_snapshot = capturing ? (e->snapshot_capture_id != 0 && e->snapshot_capture_id == cap_id)
                                       : (e->snapshot_capture_id == 0 && s != e->snapshot_stream);

  bool hashed_here = false;
  ScoreArgs a{};
  a.geo = e->geo;
  a.geo.M = e->M;
  a.plan = ps.plan;
  a.R = b.R;
  a.request_base = b.request_base;
  for (int t = 0; t < ps.plan.n_terms; t++) a.term[t] = ps.term_ptr[t];
  a.minmax_q[0] = e->cur_queue;
  a.minmax_q[1] = e->cur_running;
  for (int which = 0; which < 2; which++) {
    a.qhdr[which] = e->qhdr[which].as<QBucketHdr>();
    a.qbucket[which] = e->qbucket[which].as<uint32_t>();
  }
  a.cls_lo = e->cls_lo.as<uint32_t>();
  a.cls_hi = e->cls_hi.as<uint32_t>();
  a.summ = ps.plan.sparse_ok ? e->summ.as<AdapterSummary>() : nullptr;
  a.tiemask = ps.plan.sparse_ok ? e->tiemask.as<uint32_t>() : nullptr;
  a.prefix_lut2d = e->have_lut2d ? e->prefix_lut2d.as<double>() : nullptr;
  a.A = e->A;
  a.adapter_id = b.adapter_id;
  a.cand_mask = b.cand_mask;
  a.mask_words = (e->M + 31) / 32;
  a.dense = reinterpret_cast<const float4*>(b.dense_feat);
  a.dense_total = b.dense_total;
  a.pick = b.pick;
  a.pick_score = b.pick_score;
  a.tie_count = b.tie_count;
  a.match_out = b.match_blocks;
  a.total_out = b.total_blocks;
  a.scores_out = b.scores_out;
  if (dense && e->cfg.pick_mode != EPPSCORE_PICK_MAX_SCORE)
    return fail(e, EPPSCORE_ERR_INVALID, "the stochastic pickers are not available on dense feature rows");
  if (cfg_has(e->cfg, EPPSCORE_SCORER_LATENCY)) {
    if (dense) return fail(e, EPPSCORE_ERR_INVALID, "the latency scorer is not available on dense feature rows");
    if (!e->lat_args.enabled) return fail(e, EPPSCORE_ERR_INVALID, "latency scorer configured but the snapshot was prepared without latency params");
    a.lat = e->lat_args;
    a.lat.input_tokens = b.input_tokens;
    if (!b.input_tokens && b.prompt_bytes && b.prompt_off) {  // count the fields of the prompt bytes on the device
      CK(e, sc.s_fields.reserve((size_t)b.R * 4));
      e->launches += launch_count_fields(b.prompt_bytes, b.prompt_off, b.prompt_len, b.R, sc.s_fields.as<int32_t>(), s, e->sm_count);
      a.lat.input_tokens = sc.s_fields.as<int32_t>();
    }
    a.lat.ttft_slo = b.ttft_slo;
    a.lat.tpot_slo = b.tpot_slo;
    a.lat.pred_out = b.pred_out;
    a.n_filters = e->cfg.n_filters;
    for (int f = 0; f < e->cfg.n_filters; f++) {
      a.filter_kind[f] = e->cfg.filter_kind[f];
      for (int q = 0; q < 3; q++) a.filter_param[f][q] = e->cfg.filter_param[f][q];
    }
    a.filter_mask_out = b.filter_mask_out;
  }

  if (!dense) {
    const bool want_prefix = cfg_has(e->cfg, EPPSCORE_SCORER_PREFIX) || cfg_has(e->cfg, EPPSCORE_SCORER_LATENCY) ||
                             b.match_blocks || b.total_blocks || b.hashes_out;
    if (want_prefix && (b.hashes_in || b.prompt_bytes)) {
      if (b.hashes_in) {
        if (!b.n_hashes_in || b.hash_stride <= 0) return fail(e, EPPSCORE_ERR_INVALID, "hashes_in needs n_hashes_in and hash_stride");
        a.hashes = b.hashes_in;
        a.n_hashes = b.n_hashes_in;
        a.hash_stride = b.hash_stride;
      } else {
        const int32_t bc = b.block_chars > 0 ? b.block_chars : e->cfg.block_chars;
        const int32_t mb = b.max_blocks > 0 ? b.max_blocks : e->cfg.max_blocks;
        if (mb > EPPSCORE_MAX_BLOCKS) return fail(e, EPPSCORE_ERR_CAPACITY, "max_blocks > 65535");
        if (!b.prompt_off) return fail(e, EPPSCORE_ERR_INVALID, "prompt_off required with prompt_bytes");
        uint64_t* hashes = b.hashes_out;
        if (!hashes) {
          CK(e, sc.s_hashes.reserve((size_t)b.R * mb * 8));
          hashes = sc.s_hashes.as<uint64_t>();
        }
        CK(e, sc.s_nh.reserve((size_t)b.R * 2));
        HashArgs h{};
        h.R = b.R;
        h.bytes = b.prompt_bytes;
        h.off = b.prompt_off;
        h.len = b.prompt_len;
        h.seed = b.model_seed;
        h.block_chars = bc;
        h.max_blocks = mb;
        h.hashes = hashes;
        h.stride = mb;
        h.n_hashes = sc.s_nh.as<uint16_t>();
        h.stage_mask = e->hash_stage_mask;
        h.pdl = e->use_pdl ? 1 : 0;
        e->launches += launch_hash_prompts(h, s, e->sm_count);
        hashed_here = true;
        a.hashes = hashes;
        a.n_hashes = h.n_hashes;
        a.hash_stride = mb;
      }
      a.table = e->index->view();
    }
  }
  if (wait_snapshot) CK(e, cudaStreamWaitEvent(s, e->ev_snapshot, 0));
  // programmatic dependent launch when the previous kernel of the stream is this batch's hash kernel (the snapshot event,
  // when there is one, stays a full dependency of the pick kernel)
  a.pdl = (e->use_pdl && hashed_here) ? 1 : 0;
  // dispatch: specialised fast paths first, the fully general kernels otherwise
  int launched = 0;
  if (!e->force_generic) launched = dense ? launch_score_dense_fast(a, s, e->sm_count) : launch_pick_sparse(a, s, e->sm_count);
  if (launched == 0) launched = launch_score_pick(a, dense, s, e->sm_count);
  e->launches += launched;
  CK(e, cudaGetLastError());
  if (!capturing && s != e->stream && a.table) {  // a later index mutation (engine stream) must wait for this batch
    CK(e, cudaEventRecord(e->ev_sched, s));
    e->sched_pending = true;
  }
  return EPPSCORE_OK;
}

template <typename T>
int32_t h2d(eppscore_engine* e, DevBuf& buf, const T* src, size_t count, const T** out, cudaStream_t st = nullptr) {
  *out = nullptr;
  if (!src || count == 0) return EPPSCORE_OK;
  CK(e, buf.reserve(count * sizeof(T)));
  CK(e, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st ? st : e->stream));
  *out = buf.as<T>();
  return EPPSCORE_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int32_t eppscore_abi_version(void) { return EPPSCORE_ABI_VERSION; }

void eppscore_config_default(eppscore_config* c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(*c);
  // loader/defaults.go:46-103: queue 2, kv 2, prefix 3, max-score picker
  c->n_scorers = 3;
  c->scorer_kind[0] = EPPSCORE_SCORER_QUEUE;
  c->scorer_weight[0] = 2.0;
  c->scorer_kind[1] = EPPSCORE_SCORER_KV_CACHE;
  c->scorer_weight[1] = 2.0;
  c->scorer_kind[2] = EPPSCORE_SCORER_PREFIX;
  c->scorer_weight[2] = 3.0;
  c->block_chars = 16 * 4;  // approximateprefix/types.go:91,112
  c->max_blocks = 256;      // types.go:98
  c->tie_mode = EPPSCORE_TIE_LOWEST_INDEX;
  c->tie_seed = 0;
  c->max_endpoints = 1024;
  c->max_adapters = 64;
  c->prefix_capacity = 1 << 18;
  c->lru_capacity_default = 31250;  // types.go:109
  c->token_load_threshold = 4194304.0;  // tokenQueueThresholdDefault, token_load.go:33
}

void eppscore_latency_params_default(eppscore_latency_params* p) {
  memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(*p);
  p->has_predictions = 1;
  p->slo_buffer_factor = 1.0;  // predictedlatency/plugin.go:132
  p->streaming_mode = 0;       // :134
  p->strategy_most = 0;        // scorer/latency/plugin.go:86 "least"
  p->ttft_weight = 0.8;        // :84-85
  p->tpot_weight = 0.2;
  p->composite_kv = p->composite_queue = p->composite_prefix = 1.0;  // :87-89
}

int32_t eppscore_set_latency_params(eppscore_engine* e, const eppscore_latency_params* p) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (!p || p->struct_size != sizeof(eppscore_latency_params)) return fail(e, EPPSCORE_ERR_INVALID, "latency params NULL or struct_size mismatch");
  if (!(p->slo_buffer_factor > 0)) return fail(e, EPPSCORE_ERR_INVALID, "sloBufferFactor must be > 0");  // predictedlatency/plugin.go:169
  e->lat_params = *p;
  return EPPSCORE_OK;
}

int32_t eppscore_create(int32_t device, const eppscore_config* cfg, eppscore_engine** out) {
  if (!out) return fail(nullptr, EPPSCORE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!cfg || cfg->struct_size != sizeof(eppscore_config))
    return fail(nullptr, EPPSCORE_ERR_INVALID, "config NULL or struct_size mismatch");
  if (cfg->n_scorers < 0 || cfg->n_scorers > EPPSCORE_MAX_SCORERS) return fail(nullptr, EPPSCORE_ERR_INVALID, "n_scorers out of range");
  for (int i = 0; i < cfg->n_scorers; i++) {
    const int k = cfg->scorer_kind[i];
    const bool ok = (k >= 0 && k <= 6) || (k >= 8 && k < 12) || (k >= 16 && k < 18);
    if (!ok) return fail(nullptr, EPPSCORE_ERR_INVALID, "unknown scorer kind");
  }
  {
    int nlat = 0;
    for (int i = 0; i < cfg->n_scorers; i++) nlat += cfg->scorer_kind[i] == EPPSCORE_SCORER_LATENCY;
    if (nlat > 1) return fail(nullptr, EPPSCORE_ERR_INVALID, "at most one latency scorer per profile");
  }
  if (cfg->pick_mode < 0 || cfg->pick_mode > EPPSCORE_PICK_RANDOM) return fail(nullptr, EPPSCORE_ERR_INVALID, "unknown pick_mode");
  if (cfg->n_filters < 0 || cfg->n_filters > EPPSCORE_MAX_FILTERS) return fail(nullptr, EPPSCORE_ERR_INVALID, "n_filters out of range");
  if (cfg->n_filters > 0) {
    if (!cfg_has(*cfg, EPPSCORE_SCORER_LATENCY)) return fail(nullptr, EPPSCORE_ERR_INVALID, "device-side filters need the latency scorer in the profile");
    if (cfg_has(*cfg, EPPSCORE_SCORER_QUEUE) || cfg_has(*cfg, EPPSCORE_SCORER_RUNNING))
      return fail(nullptr, EPPSCORE_ERR_INVALID, "device-side filters cannot be combined with the queue / running scorers");
    for (int f = 0; f < cfg->n_filters; f++) {
      const int k = cfg->filter_kind[f];
      const double* par = cfg->filter_param[f];
      if (k == EPPSCORE_FILTER_PREFIX_AFFINITY) {  // prefixcacheaffinity/plugin.go:80-91
        if (par[0] > 1.0 || par[1] < 0 || par[1] > 1.0 || par[2] < 0) return fail(nullptr, EPPSCORE_ERR_INVALID, "prefix-cache-affinity-filter: invalid parameters");
      } else if (k == EPPSCORE_FILTER_SLO_HEADROOM_TIER) {  // sloheadroomtier/plugin.go:66-68
        if (par[0] < 0 || par[0] > 1.0) return fail(nullptr, EPPSCORE_ERR_INVALID, "slo-headroom-tier-filter: epsilonExploreNeg must be in [0, 1]");
      } else {
        return fail(nullptr, EPPSCORE_ERR_INVALID, "unknown filter kind");
      }
    }
  }
  if (cfg->max_endpoints < 1 || cfg->max_endpoints > 8192)
    return fail(nullptr, EPPSCORE_ERR_CAPACITY, "max_endpoints must be in [1, 8192]");
  if (cfg->max_blocks < 0 || cfg->max_blocks > EPPSCORE_MAX_BLOCKS) return fail(nullptr, EPPSCORE_ERR_CAPACITY, "max_blocks out of range");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return fail(nullptr, EPPSCORE_ERR_NO_DEVICE,
                std::string("no CUDA device (the engine has no CPU path): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, EPPSCORE_ERR_NO_DEVICE, "device index out of range");
  auto e = std::make_unique<eppscore_engine>();
  e->device = device;
  e->cfg = *cfg;
  if (e->cfg.block_chars == 0) e->cfg.block_chars = 64;
  if (e->cfg.max_blocks == 0) e->cfg.max_blocks = 256;
  if (e->cfg.max_adapters < 1) e->cfg.max_adapters = 64;
  if (e->cfg.prefix_capacity <= 0) e->cfg.prefix_capacity = 1 << 18;
  if (e->cfg.lru_capacity_default <= 0) e->cfg.lru_capacity_default = 31250;
  if (!(e->cfg.token_load_threshold > 0)) e->cfg.token_load_threshold = 4194304.0;  // token_load.go:57-61
  eppscore_latency_params_default(&e->lat_params);
  e->geo = make_geo(e->cfg.max_endpoints);
  e->A_cap = (e->cfg.max_adapters + 63) / 64 * 64;
  if (e->cfg.lru_capacity_max < 0) return fail(nullptr, EPPSCORE_ERR_INVALID, "lru_capacity_max < 0");
  eppscore_engine* ep = e.get();
  CK(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(nullptr, EPPSCORE_ERR_NO_DEVICE, "device is not sm_100-class (this library ships sm_100a code only)");
  ep->sm_count = prop.multiProcessorCount;
  {
    const char* fg = getenv("EPPSCORE_FORCE_GENERIC");
    ep->force_generic = fg && fg[0] == '1';
  }
  CK(nullptr, cudaStreamCreateWithFlags(&ep->stream, cudaStreamNonBlocking));
  CK(nullptr, cudaEventCreateWithFlags(&ep->ev_snapshot, cudaEventDisableTiming));
  CK(nullptr, cudaEventCreateWithFlags(&ep->ev_table, cudaEventDisableTiming));
  // endpoint tile buffers
  const size_t mp = (size_t)ep->geo.Mpad;
  CK(nullptr, ep->raw_kv.reserve(mp * 8));
  CK(nullptr, ep->raw_queue.reserve(mp * 8));
  CK(nullptr, ep->raw_running.reserve(mp * 8));
  CK(nullptr, ep->raw_act.reserve(mp * (ep->A_cap / 64) * 8));
  CK(nullptr, ep->raw_wait.reserve(mp * (ep->A_cap / 64) * 8));
  CK(nullptr, ep->raw_nmodels.reserve(mp * 4));
  CK(nullptr, ep->raw_max.reserve(mp * 4));
  for (int i = 0; i < 4; i++) CK(nullptr, ep->raw_col[i].reserve(mp * 8));
  CK(nullptr, ep->raw_tokens.reserve(mp * 8));
  for (int which = 0; which < 2; which++)  // value buckets for the masked min/max of the queue / running scorers
    if (cfg_has(ep->cfg, which == 0 ? EPPSCORE_SCORER_QUEUE : EPPSCORE_SCORER_RUNNING)) {
      CK(nullptr, ep->qhdr[which].reserve(sizeof(QBucketHdr)));
      CK(nullptr, ep->qbucket[which].reserve((size_t)kQBuckets * (mp / 32) * 4));
      CK(nullptr, cudaMemsetAsync(ep->qhdr[which].p, 0, sizeof(QBucketHdr), ep->stream));
    }
  if (cfg_has(ep->cfg, EPPSCORE_SCORER_LATENCY)) {
    CK(nullptr, ep->raw_min_tpot.reserve(mp * 8));
    CK(nullptr, ep->raw_dispatched.reserve(mp * 4));
    CK(nullptr, ep->raw_prefill.reserve(mp));
    CK(nullptr, ep->lat_ep.reserve(mp * 8 * kLatArrays));
  }
  for (int i = 0; i < kMaxSteps; i++) CK(nullptr, ep->term[i].reserve(mp * 8));
  CK(nullptr, ep->fold_unmasked.reserve(mp * 8));
  CK(nullptr, ep->fold_masked.reserve(mp * 8));
  CK(nullptr, ep->cls_lo.reserve((size_t)(ep->A_cap + 1) * ep->geo.row_words * 4));
  CK(nullptr, ep->cls_hi.reserve((size_t)(ep->A_cap + 1) * ep->geo.row_words * 4));
  CK(nullptr, ep->summ.reserve((size_t)(ep->A_cap + 1) * sizeof(AdapterSummary)));
  CK(nullptr, ep->tiemask.reserve((size_t)(ep->A_cap + 1) * ep->geo.row_words * 4));
  build_plan(ep, false, &ep->plan_unmasked);
  build_plan(ep, true, &ep->plan_masked);
  // engine-wide prefix term table for the sparse path: lut2d[total][c] = clamp(c/total) * w  (prefix/plugin.go:108-110,
  // scheduler_profile.go:168) — one IEEE divide and one multiply per entry, computed here on the host (no contraction:
  // this file is compiled with -ffp-contract=off)
  for (int sidx = 0; sidx < ep->plan_unmasked.plan.n_steps; sidx++) {
    if (ep->plan_unmasked.plan.kind[sidx] != STEP_PREFIX) continue;
    const double w = ep->plan_unmasked.plan.weight[sidx];
    std::vector<double> lut((size_t)(kLutMax + 1) * (kLutMax + 1), 0.0);
    for (int total = 0; total <= kLutMax; total++)
      for (int c = 0; c <= kLutMax; c++) {
        volatile double sc = 0.0;
        if (total != 0) sc = (double)c / (double)total;
        volatile double cl = clamp01_host(sc);
        lut[(size_t)total * (kLutMax + 1) + c] = cl * w;
      }
    CK(nullptr, ep->prefix_lut2d.reserve(lut.size() * 8));
    CK(nullptr, cudaMemcpy(ep->prefix_lut2d.p, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice));
    ep->have_lut2d = true;
    break;
  }
  // prefix index (device-resident: slot table + per-endpoint LRUs; the LRU regions are allocated by the first Add)
  CK(nullptr, cudaEventCreateWithFlags(&ep->ev_sched, cudaEventDisableTiming));
  CK(nullptr, cudaEventCreateWithFlags(&ep->ev_caller, cudaEventDisableTiming));
  ep->index = std::make_unique<DeviceIndex>(ep->geo.Mpad, ep->geo.Mpad / 32, ep->cfg.prefix_capacity, ep->cfg.lru_capacity_default,
                                            ep->cfg.lru_capacity_max);
  CK(nullptr, ep->index->init(ep->stream));
  CK(nullptr, cudaEventRecord(ep->ev_table, ep->stream));
  CK(nullptr, cudaEventRecord(ep->ev_snapshot, ep->stream));
  CK(nullptr, cudaStreamSynchronize(ep->stream));
  *out = e.release();
  return EPPSCORE_OK;
}

void eppscore_destroy(eppscore_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  DevBuf* bufs[] = {&e->qhdr[0], &e->qhdr[1], &e->qbucket[0], &e->qbucket[1], &e->raw_min_tpot, &e->raw_dispatched, &e->raw_prefill, &e->raw_tokens, &e->lat_ep,
                    &e->raw_kv, &e->raw_queue, &e->raw_running, &e->raw_act, &e->raw_wait, &e->raw_nmodels, &e->raw_max,
                    &e->fold_unmasked, &e->fold_masked, &e->cls_lo, &e->cls_hi, &e->summ, &e->tiemask, &e->prefix_lut2d, &e->c_pick, &e->c_hash, &e->c_nh,
                    &e->c_ep, &e->c_op};
  for (auto& x : e->sc) x.release();
  for (int i = 0; i < 3; i++) {
    if (e->split_stream[i]) {
      cudaStreamSynchronize(e->split_stream[i]);
      cudaStreamDestroy(e->split_stream[i]);
    }
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->h_res) cudaFreeHost(e->h_res);
  if (e->stream2) {
    cudaStreamSynchronize(e->stream2);
    cudaStreamDestroy(e->stream2);
  }
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 4; i++) e->raw_col[i].release();
  for (int i = 0; i < kMaxSteps; i++) e->term[i].release();
  e->index.reset();
  if (e->ev_sched) cudaEventDestroy(e->ev_sched);
  if (e->ev_caller) cudaEventDestroy(e->ev_caller);
  if (e->ev_snapshot) cudaEventDestroy(e->ev_snapshot);
  if (e->ev_table) cudaEventDestroy(e->ev_table);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* eppscore_last_error(const eppscore_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int32_t eppscore_get_stats(const eppscore_engine* ce, eppscore_stats* out) {
  eppscore_engine* e = const_cast<eppscore_engine*>(ce);
  if (!e || !out) return EPPSCORE_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(*out);
  out->M = e->M;
  out->epoch = e->epoch;
  CK(e, cudaSetDevice(e->device));
  IndexStats st{};
  CK(e, e->index->stats(&st, e->stream));
  out->kernel_launches = e->launches + e->index->launches();
  out->prefix_hashes = st.used;
  out->prefix_live_hashes = st.live;
  out->prefix_capacity = st.capacity;
  out->prefix_table_bytes = st.table_bytes;
  out->lru_entries = st.lru_entries;
  out->lru_bytes = st.lru_bytes;
  out->prefix_overflow_rows = st.ovf_rows;
  out->prefix_rebuilds = st.rebuilds;
  out->index_error = st.error;
  return EPPSCORE_OK;
}

int32_t eppscore_set_debug(eppscore_engine* e, int32_t key, int64_t value) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (key == 1) {
    e->force_generic = value != 0;
    return EPPSCORE_OK;
  }
  if (key == 2) {
    e->hash_stage_mask = (int32_t)(value & 63);
    return EPPSCORE_OK;
  }
  if (key == 3) {
    e->host_chunk = (int32_t)value;
    return EPPSCORE_OK;
  }
  if (key == 4) {
    e->use_pdl = value != 0;
    return EPPSCORE_OK;
  }
  if (key == 5) {
    e->dev_split = value < 1 ? 1 : (value > 64 ? 64 : (int32_t)value);
    return EPPSCORE_OK;
  }
  if (key == 6) {
    e->dev_streams = value < 1 ? 1 : (value > 4 ? 4 : (int32_t)value);
    return EPPSCORE_OK;
  }
  return fail(e, EPPSCORE_ERR_INVALID, "unknown debug key");
}

int32_t eppscore_set_snapshot(eppscore_engine* e, const eppscore_snapshot* s) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (!s || s->struct_size != sizeof(eppscore_snapshot)) return fail(e, EPPSCORE_ERR_INVALID, "snapshot NULL or struct_size mismatch");
  if (s->M < 0 || s->M > e->geo.Mpad) return fail(e, EPPSCORE_ERR_CAPACITY, "M exceeds config.max_endpoints");
  if (s->lora_words < 0 || s->lora_words * 64 > e->A_cap) return fail(e, EPPSCORE_ERR_CAPACITY, "lora_words*64 exceeds config.max_adapters");
  if (s->M > 0 && (!s->kv_usage || !s->queue)) return fail(e, EPPSCORE_ERR_INVALID, "kv_usage and queue are required");
  CK(e, cudaSetDevice(e->device));
  const bool on_device = s->location == 1;
  cudaStream_t st = (on_device && s->stream) ? (cudaStream_t)s->stream : e->stream;
  const size_t M = (size_t)s->M;
  const bool have_lora = s->lora_active && s->lora_waiting && s->lora_words > 0;
  {  // the term / class-plane / summary buffers are rewritten below: a device-location batch that another stream launched
     // against the previous snapshot must have finished with them (inside a capture the graph's own edges order this)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone && e->sched_pending) CK(e, cudaStreamWaitEvent(st, e->ev_sched, 0));
  }
  // Host snapshots are copied into the engine's tile buffers; device snapshots (e.g. the buffer an NCCL
  // broadcast just filled) are used IN PLACE — they must stay valid until the next set_snapshot.
  if (!on_device) {
    auto cp = [&](DevBuf& dst, const void* src, size_t bytes) -> cudaError_t {
      if (!src || bytes == 0) return cudaSuccess;
      return cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, st);
    };
    CK(e, cp(e->raw_kv, s->kv_usage, M * 8));
    CK(e, cp(e->raw_queue, s->queue, M * 8));
    CK(e, cp(e->raw_running, s->running, M * 8));
    if (have_lora) {
      CK(e, cp(e->raw_act, s->lora_active, M * s->lora_words * 8));
      CK(e, cp(e->raw_wait, s->lora_waiting, M * s->lora_words * 8));
    }
    CK(e, cp(e->raw_nmodels, s->lora_nmodels, M * 4));
    CK(e, cp(e->raw_max, s->lora_max, M * 4));
    for (int i = 0; i < 4; i++) CK(e, cp(e->raw_col[i], s->endpoint_col[i], M * 8));
    CK(e, cp(e->raw_tokens, s->inflight_tokens, M * 8));
    if (e->lat_ep.p) {
      CK(e, cp(e->raw_min_tpot, s->min_tpot_slo, M * 8));
      CK(e, cp(e->raw_dispatched, s->dispatched, M * 4));
      CK(e, cp(e->raw_prefill, s->prefill_role, M));
    }
  }
  e->cur_queue = on_device ? s->queue : e->raw_queue.as<int64_t>();
  e->cur_running = s->running ? (on_device ? s->running : e->raw_running.as<int64_t>()) : nullptr;
  e->M = s->M;
  e->lora_words = have_lora ? s->lora_words : 0;
  e->A = e->lora_words * 64;
  e->epoch = s->epoch;

  PrepareArgs pa{};
  pa.geo = e->geo;
  pa.geo.M = s->M;
  pa.n_scorers = e->cfg.n_scorers;
  for (int i = 0; i < e->cfg.n_scorers; i++) {
    pa.kind[i] = e->cfg.scorer_kind[i];
    pa.weight[i] = e->cfg.scorer_weight[i];
    pa.term[i] = is_endpoint_term_kind(pa.kind[i], false) ? e->term[i].as<double>() : nullptr;
  }
  pa.kv = on_device ? s->kv_usage : e->raw_kv.as<double>();
  pa.queue = e->cur_queue;
  pa.running = e->cur_running;
  pa.act = have_lora ? (on_device ? s->lora_active : e->raw_act.as<uint64_t>()) : nullptr;
  pa.wait = have_lora ? (on_device ? s->lora_waiting : e->raw_wait.as<uint64_t>()) : nullptr;
  pa.nmodels = s->lora_nmodels ? (on_device ? s->lora_nmodels : e->raw_nmodels.as<int32_t>()) : nullptr;
  pa.maxm = s->lora_max ? (on_device ? s->lora_max : e->raw_max.as<int32_t>()) : nullptr;
  for (int i = 0; i < 4; i++)
    pa.col[i] = s->endpoint_col[i] ? (on_device ? s->endpoint_col[i] : e->raw_col[i].as<double>()) : nullptr;
  for (int which = 0; which < 2; which++) {
    pa.qhdr[which] = e->qhdr[which].as<QBucketHdr>();
    pa.qbucket[which] = e->qbucket[which].as<uint32_t>();
  }
  pa.tokens = s->inflight_tokens ? (on_device ? s->inflight_tokens : e->raw_tokens.as<int64_t>()) : nullptr;
  pa.token_threshold = e->cfg.token_load_threshold;
  e->lat_args = LatArgs{};
  if (e->lat_ep.p) {
    // latency fold-in: split the linear forms into endpoint-only arrays (prepared on the device) and
    // per-request / per-pair terms (LatArgs); the weight normalisations are single IEEE divides
    // (scorer/latency/plugin.go:325-333,373-379) done here on the host (-ffp-contract=off).
    const eppscore_latency_params& lp = e->lat_params;
    pa.min_tpot = s->min_tpot_slo ? (on_device ? s->min_tpot_slo : e->raw_min_tpot.as<double>()) : nullptr;
    pa.dispatched = s->dispatched ? (on_device ? s->dispatched : e->raw_dispatched.as<int32_t>()) : nullptr;
    pa.prefill = s->prefill_role ? (on_device ? s->prefill_role : e->raw_prefill.as<uint8_t>()) : nullptr;
    const double coef[8] = {lp.ttft_intercept, lp.ttft_kv, lp.ttft_waiting, lp.ttft_running,
                            lp.tpot_intercept, lp.tpot_kv, lp.tpot_waiting, lp.tpot_running};
    for (int i = 0; i < 8; i++) pa.lat_coef[i] = coef[i];
    pa.lat_buffer = lp.slo_buffer_factor;
    pa.lat_streaming = lp.streaming_mode;
    volatile double wkv = lp.composite_kv, wq = lp.composite_queue, wpref = lp.composite_prefix;
    volatile double sumw = wkv + wq;
    sumw = sumw + wpref;
    if (sumw <= 0) {
      wkv = 1;
      wq = 0;
      wpref = 0;
      sumw = 1;
    }
    wkv = wkv / sumw;
    wq = wq / sumw;
    wpref = wpref / sumw;
    pa.lat_ckv = wkv;
    pa.lat_ep = e->lat_ep.as<double>();
    pa.lat_has_predictions = lp.has_predictions;
    LatArgs& L = e->lat_args;
    L.enabled = 1;
    L.has_predictions = lp.has_predictions;
    L.strategy_most = lp.strategy_most;
    L.ttft_input = lp.ttft_input;
    L.ttft_prefix = lp.ttft_prefix;
    L.tpot_input = lp.tpot_input;
    L.tpot_generated = lp.tpot_generated;  // * float64(NumTokensGenerated = 1) is exact
    L.buffer = lp.slo_buffer_factor;
    volatile double wsum = lp.ttft_weight + lp.tpot_weight;
    if (wsum <= 0) {
      L.alpha = 1.0;
      L.beta = 0.0;
    } else {
      volatile double al = lp.ttft_weight / wsum, be = lp.tpot_weight / wsum;
      L.alpha = al;
      L.beta = be;
    }
    L.wq = wq;
    L.wpref = wpref;
    L.ep = pa.lat_ep;
  }
  pa.lora_words = e->lora_words;
  pa.A = e->A;
  int lead_u = 0, lead_m = 0;
  while (lead_u < e->cfg.n_scorers && is_endpoint_term_kind(e->cfg.scorer_kind[lead_u], false)) lead_u++;
  while (lead_m < e->cfg.n_scorers && is_endpoint_term_kind(e->cfg.scorer_kind[lead_m], true)) lead_m++;
  pa.fold_unmasked = lead_u ? e->fold_unmasked.as<double>() : nullptr;
  pa.fold_unmasked_n = lead_u;
  pa.fold_masked = lead_m ? e->fold_masked.as<double>() : nullptr;
  pa.fold_masked_n = lead_m;
  pa.cls_lo = e->cls_lo.as<uint32_t>();
  pa.cls_hi = e->cls_hi.as<uint32_t>();
  pa.plan_u = e->plan_unmasked.plan;
  for (int t = 0; t < kMaxSteps; t++) pa.plan_term[t] = e->plan_unmasked.term_ptr[t];
  pa.summ = e->plan_unmasked.plan.sparse_ok ? e->summ.as<AdapterSummary>() : nullptr;
  pa.tiemask = e->tiemask.as<uint32_t>();
  e->launches += launch_prepare_snapshot(pa, st);
  CK(e, cudaGetLastError());
  // Cross-stream ordering through an event — skipped while the caller's stream is being captured into a
  // CUDA graph (a captured stream may not take dependencies from outside its capture; in-stream order suffices).
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  unsigned long long cap_id = 0;
  cudaStreamGetCaptureInfo(st, &cap, &cap_id);
  CK(e, cudaEventRecord(e->ev_snapshot, st));  // inside a capture this becomes a graph edge other captured streams can join on
  e->snapshot_stream = st;
  if (cap == cudaStreamCaptureStatusNone) {
    e->snapshot_capture_id = 0;
    if (st != e->stream) CK(e, cudaStreamWaitEvent(e->stream, e->ev_snapshot, 0));
    if (!on_device) CK(e, cudaStreamSynchronize(st));  // host arrays may be reused by the caller
  } else {
    e->snapshot_capture_id = cap_id;
  }
  e->have_snapshot = true;
  return EPPSCORE_OK;
}

int32_t eppscore_schedule_batch(eppscore_engine* e, const eppscore_batch* b) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (!b || b->struct_size != sizeof(eppscore_batch)) return fail(e, EPPSCORE_ERR_INVALID, "batch NULL or struct_size mismatch");
  if (b->R < 0) return fail(e, EPPSCORE_ERR_INVALID, "R < 0");
  if (b->R == 0) return EPPSCORE_OK;
  CK(e, cudaSetDevice(e->device));
  int32_t rc = EPPSCORE_OK;
  DevBatch d{};
  d.R = b->R;
  d.request_base = b->request_base;
  d.hash_stride = b->hash_stride;
  d.block_chars = b->block_chars;
  d.max_blocks = b->max_blocks;
  if (b->location == 1) {
    d.prompt_bytes = b->prompt_bytes;
    d.prompt_off = b->prompt_off;
    d.prompt_len = b->prompt_len;
    d.model_seed = b->model_seed;
    d.hashes_in = b->hashes_in;
    d.n_hashes_in = b->n_hashes_in;
    d.adapter_id = b->adapter_id;
    d.cand_mask = b->cand_mask;
    d.dense_feat = b->dense_feat;
    d.dense_total = b->dense_total;
    d.input_tokens = b->input_tokens;
    d.ttft_slo = b->ttft_slo;
    d.tpot_slo = b->tpot_slo;
    d.pred_out = b->pred_out;
    d.filter_mask_out = b->filter_mask_out;
    d.pick = b->pick;
    d.pick_score = b->pick_score;
    d.tie_count = b->tie_count;
    d.match_blocks = b->match_blocks;
    d.total_blocks = b->total_blocks;
    d.hashes_out = b->hashes_out;
    d.scores_out = b->scores_out;
    cudaStream_t cs = b->stream ? (cudaStream_t)b->stream : e->stream;
    // R x M outputs, masks, supplied hashes and dense rows go as one slice (their layouts are not sliced here)
    const bool sliceable = d.prompt_bytes && !d.hashes_in && !d.cand_mask && !d.dense_feat && !d.pred_out && !d.filter_mask_out &&
                           !d.match_blocks && !d.scores_out && !d.input_tokens && !d.ttft_slo && !d.tpot_slo;
    int K = e->dev_split;
    if (!sliceable || e->dev_streams < 2) K = 1;
    while (K > 1 && b->R / K < 2048) K--;
    if (K <= 1) return schedule_device(e, d, cs, e->sc[0]);
    // fork: the internal streams join the caller's stream (also inside a CUDA-graph capture), slice c goes to stream c mod S
    const int S = e->dev_streams < K ? e->dev_streams : K;
    if (!e->ev_fork) CK(e, cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < S - 1; i++) {
      if (!e->split_stream[i]) CK(e, cudaStreamCreateWithFlags(&e->split_stream[i], cudaStreamNonBlocking));
      if (!e->ev_join[i]) CK(e, cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming));
    }
    CK(e, cudaEventRecord(e->ev_fork, cs));
    for (int i = 0; i < S - 1; i++) CK(e, cudaStreamWaitEvent(e->split_stream[i], e->ev_fork, 0));
    const int32_t mbk = d.max_blocks > 0 ? d.max_blocks : e->cfg.max_blocks;
    const int32_t per = ((b->R + K - 1) / K + 63) / 64 * 64;
    for (int c = 0, base = 0; base < b->R; c++, base += per) {
      DevBatch q = d;
      q.R = std::min(per, b->R - base);
      q.request_base = d.request_base + base;
      q.prompt_off = d.prompt_off + base;
      if (d.prompt_len) q.prompt_len = d.prompt_len + base;
      if (d.model_seed) q.model_seed = d.model_seed + base;
      if (d.adapter_id) q.adapter_id = d.adapter_id + base;
      q.pick = d.pick + base;
      q.pick_score = d.pick_score + base;
      q.tie_count = d.tie_count + base;
      if (d.total_blocks) q.total_blocks = d.total_blocks + base;
      if (d.hashes_out) q.hashes_out = d.hashes_out + (size_t)base * mbk;
      const int si = c % S;
      rc = schedule_device(e, q, si == 0 ? cs : e->split_stream[si - 1], e->sc[si]);
      if (rc != EPPSCORE_OK) break;
    }
    for (int i = 0; i < S - 1; i++) {  // join, also on the error path (a capture must not be left forked)
      cudaEventRecord(e->ev_join[i], e->split_stream[i]);
      cudaStreamWaitEvent(cs, e->ev_join[i], 0);
    }
    if (rc != EPPSCORE_OK) return rc;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(cs, &cap);
    if (cap == cudaStreamCaptureStatusNone && cs != e->stream) {
      CK(e, cudaEventRecord(e->ev_sched, cs));
      e->sched_pending = true;
    }
    return EPPSCORE_OK;
  }
  // ---- host buffers: H2D, kernels, D2H, all inside this call ----
  // The batch is cut into chunks that alternate between two streams (each with its own scratch set): the H2D copy of
  // chunk i+1 runs under the kernels and the D2H copy of chunk i, so the call costs about max(copy, compute) instead of
  // their sum.  Requests are independent given (snapshot, table), so any cut gives identical results (request_base keeps
  // the tie priorities of the whole batch).
  if (!e->have_snapshot) return fail(e, EPPSCORE_ERR_NO_SNAPSHOT, "schedule_batch before set_snapshot");
  if (b->prompt_bytes && !b->prompt_off) return fail(e, EPPSCORE_ERR_INVALID, "prompt_off required with prompt_bytes");
  if (b->hashes_out && (b->hashes_in || !b->prompt_bytes))
    return fail(e, EPPSCORE_ERR_INVALID, "hashes_out needs prompt_bytes (hashes supplied as hashes_in are already the caller's)");
  const size_t M = (size_t)e->M;
  const size_t mw = (M + 31) / 32;
  const int32_t mb = b->max_blocks > 0 ? b->max_blocks : e->cfg.max_blocks;
  const bool diag = b->match_blocks || b->scores_out || b->pred_out || b->dense_feat;  // R x M arrays: one chunk at a time is plenty
  int32_t chunk = e->host_chunk > 0 ? e->host_chunk : b->R;
  if (diag || b->R < 2 * chunk) chunk = b->R;
  const int nstreams = chunk < b->R ? 2 : 1;
  if (nstreams == 2 && !e->stream2) CK(e, cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  // pinned staging: pick i32 | score f64 | tie i32 | total u16, each [R]
  const size_t RR = (size_t)b->R;
  const size_t o_score = (RR * 4 + 7) & ~(size_t)7, o_tie = o_score + RR * 8, o_total = o_tie + RR * 4, res_bytes = o_total + RR * 2;
  if (e->h_res_bytes < res_bytes) {
    if (e->h_res) cudaFreeHost(e->h_res);
    e->h_res = nullptr;
    e->h_res_bytes = 0;
    CK(e, cudaHostAlloc(&e->h_res, res_bytes + res_bytes / 4, cudaHostAllocDefault));
    e->h_res_bytes = res_bytes + res_bytes / 4;
  }
  unsigned char* hres = static_cast<unsigned char*>(e->h_res);
  int32_t* hp_pick = reinterpret_cast<int32_t*>(hres);
  double* hp_score = reinterpret_cast<double*>(hres + o_score);
  int32_t* hp_tie = reinterpret_cast<int32_t*>(hres + o_tie);
  uint16_t* hp_total = reinterpret_cast<uint16_t*>(hres + o_total);
  int ci = 0;
  for (int32_t r0 = 0; r0 < b->R; r0 += chunk, ci++) {
    const int32_t r1 = std::min(b->R, r0 + chunk);
    const size_t R = (size_t)(r1 - r0);
    Scratch& sc = e->sc[ci & 1];
    cudaStream_t st = (ci & 1) ? e->stream2 : e->stream;
    DevBatch d{};
    d.R = (int32_t)R;
    d.request_base = b->request_base + r0;
    d.hash_stride = b->hash_stride;
    d.block_chars = b->block_chars;
    d.max_blocks = b->max_blocks;
    if (b->prompt_bytes) {
      // byte range of the chunk; the device copy starts at a 16-byte aligned source offset so every prompt keeps its
      // alignment (the fast hash path needs 16-byte aligned starts), and the kernel keeps using the caller's offsets
      size_t lo = (size_t)b->prompt_off[r0], hi = (size_t)b->prompt_off[r1];
      if (b->prompt_len) {
        lo = SIZE_MAX;
        hi = 0;
        for (int32_t r = r0; r < r1; r++) {
          lo = std::min(lo, (size_t)b->prompt_off[r]);
          hi = std::max(hi, (size_t)b->prompt_off[r] + (size_t)b->prompt_len[r]);
        }
        if (lo > hi) lo = hi = 0;
      }
      const size_t start = lo & ~(size_t)15;
      // pad the device copy so 16-byte vector loads of the last block never leave the allocation
      CK(e, sc.s_prompts.reserve(hi - start + 64));
      if (hi > start) CK(e, cudaMemcpyAsync(sc.s_prompts.p, b->prompt_bytes + start, hi - start, cudaMemcpyHostToDevice, st));
      d.prompt_bytes = sc.s_prompts.as<uint8_t>() - start;
      if ((rc = h2d(e, sc.s_off, b->prompt_off + r0, R + 1, &d.prompt_off, st)) != EPPSCORE_OK) return rc;
    }
#define H2D(buf, field, per) \
  if ((rc = h2d(e, sc.buf, b->field ? b->field + (size_t)r0 * (per) : b->field, R * (per), &d.field, st)) != EPPSCORE_OK) return rc;
    H2D(s_len, prompt_len, 1)
    H2D(s_seed, model_seed, 1)
    if (b->hashes_in) {
      H2D(s_hashes, hashes_in, (size_t)b->hash_stride)
      H2D(s_nh, n_hashes_in, 1)
    }
    H2D(s_adapter, adapter_id, 1)
    H2D(s_mask, cand_mask, mw)
    H2D(s_dense, dense_feat, M * 4)
    H2D(s_dtotal, dense_total, 1)
    H2D(s_intok, input_tokens, 1)
    H2D(s_tslo, ttft_slo, 1)
    H2D(s_pslo, tpot_slo, 1)
#undef H2D
    CK(e, sc.s_pick.reserve(R * 4));
    CK(e, sc.s_score.reserve(R * 8));
    CK(e, sc.s_tie.reserve(R * 4));
    d.pick = sc.s_pick.as<int32_t>();
    d.pick_score = sc.s_score.as<double>();
    d.tie_count = sc.s_tie.as<int32_t>();
    if (b->match_blocks) {
      CK(e, sc.s_match.reserve(R * M * 2));
      d.match_blocks = sc.s_match.as<uint16_t>();
    }
    if (b->total_blocks) {
      CK(e, sc.s_total.reserve(R * 2));
      d.total_blocks = sc.s_total.as<uint16_t>();
    }
    if (b->scores_out) {
      CK(e, sc.s_scores.reserve(R * M * 8));
      d.scores_out = sc.s_scores.as<double>();
    }
    if (b->pred_out) {
      CK(e, sc.s_pred.reserve(R * M * 16));
      d.pred_out = sc.s_pred.as<double>();
    }
    if (b->filter_mask_out) {
      CK(e, sc.s_fmask.reserve(R * mw * 4));
      d.filter_mask_out = sc.s_fmask.as<uint32_t>();
    }
    uint64_t* hashes_dev = nullptr;
    if (b->hashes_out) {
      // reuse the internal hash scratch as the device-side hashes_out; entries past n_hashes[r] are zero
      CK(e, sc.s_hashes.reserve(R * (size_t)mb * 8));
      CK(e, cudaMemsetAsync(sc.s_hashes.p, 0, R * (size_t)mb * 8, st));
      hashes_dev = sc.s_hashes.as<uint64_t>();
      d.hashes_out = hashes_dev;
    }
    rc = schedule_device(e, d, st, sc);
    if (rc != EPPSCORE_OK) return rc;
    if (b->pick) CK(e, cudaMemcpyAsync(hp_pick + r0, d.pick, R * 4, cudaMemcpyDeviceToHost, st));
    if (b->pick_score) CK(e, cudaMemcpyAsync(hp_score + r0, d.pick_score, R * 8, cudaMemcpyDeviceToHost, st));
    if (b->tie_count) CK(e, cudaMemcpyAsync(hp_tie + r0, d.tie_count, R * 4, cudaMemcpyDeviceToHost, st));
    if (b->match_blocks) CK(e, cudaMemcpyAsync(b->match_blocks + (size_t)r0 * M, d.match_blocks, R * M * 2, cudaMemcpyDeviceToHost, st));
    if (b->total_blocks) CK(e, cudaMemcpyAsync(hp_total + r0, d.total_blocks, R * 2, cudaMemcpyDeviceToHost, st));
    if (b->scores_out) CK(e, cudaMemcpyAsync(b->scores_out + (size_t)r0 * M, d.scores_out, R * M * 8, cudaMemcpyDeviceToHost, st));
    if (b->pred_out) CK(e, cudaMemcpyAsync(b->pred_out + (size_t)r0 * M * 2, d.pred_out, R * M * 16, cudaMemcpyDeviceToHost, st));
    if (b->filter_mask_out) CK(e, cudaMemcpyAsync(b->filter_mask_out + (size_t)r0 * mw, d.filter_mask_out, R * mw * 4, cudaMemcpyDeviceToHost, st));
    if (hashes_dev) CK(e, cudaMemcpyAsync(b->hashes_out + (size_t)r0 * mb, hashes_dev, R * (size_t)mb * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(e, cudaStreamSynchronize(e->stream));
  if (nstreams == 2) CK(e, cudaStreamSynchronize(e->stream2));
  if (b->pick) memcpy(b->pick, hp_pick, RR * 4);
  if (b->pick_score) memcpy(b->pick_score, hp_score, RR * 8);
  if (b->tie_count) memcpy(b->tie_count, hp_tie, RR * 4);
  if (b->total_blocks) memcpy(b->total_blocks, hp_total, RR * 2);
  e->sched_pending = false;  // everything launched above has completed
  return EPPSCORE_OK;
}

int32_t eppscore_hash_prompts(eppscore_engine* e, int32_t R, int32_t location, const uint8_t* prompt_bytes,
                              const int64_t* prompt_off, const int32_t* prompt_len, const uint64_t* model_seed,
                              int32_t block_chars, int32_t max_blocks, uint64_t* hashes_out, uint16_t* n_hashes_out,
                              void* stream) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R <= 0) return EPPSCORE_OK;
  if (!prompt_bytes || !prompt_off || !hashes_out || !n_hashes_out) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  const int32_t bc = block_chars > 0 ? block_chars : e->cfg.block_chars;
  const int32_t mb = max_blocks > 0 ? max_blocks : e->cfg.max_blocks;
  if (mb > EPPSCORE_MAX_BLOCKS) return fail(e, EPPSCORE_ERR_CAPACITY, "max_blocks > 65535");
  CK(e, cudaSetDevice(e->device));
  HashArgs h{};
  h.R = R;
  h.block_chars = bc;
  h.max_blocks = mb;
  h.stride = mb;
  h.stage_mask = e->hash_stage_mask;
  if (location == 1) {
    h.bytes = prompt_bytes;
    h.off = prompt_off;
    h.len = prompt_len;
    h.seed = model_seed;
    h.hashes = hashes_out;
    h.n_hashes = n_hashes_out;
    e->launches += launch_hash_prompts(h, stream ? (cudaStream_t)stream : e->stream, e->sm_count);
    CK(e, cudaGetLastError());
    return EPPSCORE_OK;
  }
  int32_t rc;
  size_t total = (size_t)prompt_off[R];
  if (prompt_len) {
    total = 0;
    for (int32_t r = 0; r < R; r++) total = std::max(total, (size_t)prompt_off[r] + (size_t)prompt_len[r]);
  }
  CK(e, e->sc[0].s_prompts.reserve(total + 64));
  if (total) CK(e, cudaMemcpyAsync(e->sc[0].s_prompts.p, prompt_bytes, total, cudaMemcpyHostToDevice, e->stream));
  h.bytes = e->sc[0].s_prompts.as<uint8_t>();
  if ((rc = h2d(e, e->sc[0].s_off, prompt_off, (size_t)R + 1, &h.off)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->sc[0].s_len, prompt_len, (size_t)R, &h.len)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->sc[0].s_seed, model_seed, (size_t)R, &h.seed)) != EPPSCORE_OK) return rc;
  CK(e, e->sc[0].s_hashes.reserve((size_t)R * mb * 8));
  CK(e, e->sc[0].s_nh.reserve((size_t)R * 2));
  CK(e, cudaMemsetAsync(e->sc[0].s_hashes.p, 0, (size_t)R * mb * 8, e->stream));
  h.hashes = e->sc[0].s_hashes.as<uint64_t>();
  h.n_hashes = e->sc[0].s_nh.as<uint16_t>();
  e->launches += launch_hash_prompts(h, e->stream, e->sm_count);
  CK(e, cudaGetLastError());
  CK(e, cudaMemcpyAsync(hashes_out, h.hashes, (size_t)R * mb * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaMemcpyAsync(n_hashes_out, h.n_hashes, (size_t)R * 2, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaStreamSynchronize(e->stream));
  return EPPSCORE_OK;
}

uint64_t eppscore_xxh64(const void* data, size_t len, uint64_t seed) { return xxh64_host(data, len, seed); }

uint64_t eppscore_model_seed(const void* model, size_t model_len, const void* salt, size_t salt_len) {
  std::string buf;
  buf.reserve(model_len + salt_len);
  if (model_len) buf.append(static_cast<const char*>(model), model_len);
  if (salt_len) buf.append(static_cast<const char*>(salt), salt_len);
  return xxh64_host(buf.data(), buf.size(), 0);
}

// ---------------- prefix index ----------------
int32_t eppscore_count_fields(eppscore_engine* e, int32_t R, int32_t location, const uint8_t* prompt_bytes,
                              const int64_t* prompt_off, const int32_t* prompt_len, int32_t* out, void* stream) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R < 0 || (R > 0 && (!prompt_bytes || !prompt_off || !out))) return fail(e, EPPSCORE_ERR_INVALID, "count_fields: NULL argument");
  if (R == 0) return EPPSCORE_OK;
  CK(e, cudaSetDevice(e->device));
  if (location == 1) {
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    e->launches += launch_count_fields(prompt_bytes, prompt_off, prompt_len, R, out, s, e->sm_count);
    CK(e, cudaGetLastError());
    return EPPSCORE_OK;
  }
  size_t total = (size_t)prompt_off[R];
  if (prompt_len) {
    total = 0;
    for (int32_t r = 0; r < R; r++) total = std::max(total, (size_t)prompt_off[r] + (size_t)prompt_len[r]);
  }
  CK(e, e->sc[0].s_prompts.reserve(total + 64));
  if (total) CK(e, cudaMemcpyAsync(e->sc[0].s_prompts.p, prompt_bytes, total, cudaMemcpyHostToDevice, e->stream));
  const int64_t* d_off = nullptr;
  const int32_t* d_len = nullptr;
  int32_t rc;
  if ((rc = h2d(e, e->sc[0].s_off, prompt_off, (size_t)R + 1, &d_off)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->sc[0].s_len, prompt_len, (size_t)R, &d_len)) != EPPSCORE_OK) return rc;
  CK(e, e->sc[0].s_fields.reserve((size_t)R * 4));
  e->launches += launch_count_fields(e->sc[0].s_prompts.as<uint8_t>(), d_off, d_len, R, e->sc[0].s_fields.as<int32_t>(), e->stream, e->sm_count);
  CK(e, cudaGetLastError());
  CK(e, cudaMemcpyAsync(out, e->sc[0].s_fields.p, (size_t)R * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaStreamSynchronize(e->stream));
  return EPPSCORE_OK;
}

// lru_capacity (host, [M] or NULL) -> a padded host vector the index can copy; validates the values
static int32_t pad_caps(eppscore_engine* e, const int32_t* lru_capacity, std::vector<int32_t>* out) {
  out->clear();
  if (!lru_capacity) return EPPSCORE_OK;
  out->assign((size_t)e->geo.Mpad, 0);
  const int32_t n = e->M > 0 ? e->M : e->geo.Mpad;
  for (int32_t i = 0; i < n; i++) (*out)[i] = lru_capacity[i] > 0 ? lru_capacity[i] : 0;
  return EPPSCORE_OK;
}
static int32_t index_status(eppscore_engine* e, cudaError_t c, const char* what) {
  if (c == cudaSuccess) return EPPSCORE_OK;
  if (c == cudaErrorInvalidValue)
    return fail(e, EPPSCORE_ERR_CAPACITY, std::string(what) + ": an LRU capacity above the engine's lru_capacity_max (the per-endpoint regions were sized at the first Add)");
  return cuda_fail(e, c, what);
}

int32_t eppscore_commit_picks(eppscore_engine* e, int32_t R, const int32_t* pick, const uint64_t* hashes,
                              const uint16_t* n_hashes, int32_t hash_stride, const int32_t* lru_capacity) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R < 0 || (R > 0 && (!pick || !hashes || !n_hashes))) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (R == 0) return EPPSCORE_OK;
  if (hash_stride <= 0) return fail(e, EPPSCORE_ERR_INVALID, "hash_stride must be > 0");
  CK(e, cudaSetDevice(e->device));
  int32_t lo = e->geo.Mpad, hi = -1;
  int64_t touches = 0;
  for (int32_t r = 0; r < R; r++) {
    const int32_t ep = pick[r];
    if (ep < 0) continue;  // no target endpoint: nothing to record (plugin.go:173-175)
    if (ep >= e->geo.Mpad) return fail(e, EPPSCORE_ERR_INVALID, "pick out of range");
    if (n_hashes[r] > hash_stride) return fail(e, EPPSCORE_ERR_INVALID, "n_hashes exceeds hash_stride");
    lo = std::min(lo, ep);
    hi = std::max(hi, ep);
    touches += n_hashes[r];
  }
  if (hi < 0) return EPPSCORE_OK;  // (a pick with zero hashes still creates the endpoint's LRU, indexer.go:57-68)
  std::vector<int32_t> caps;
  int32_t rc = pad_caps(e, lru_capacity, &caps);
  if (rc != EPPSCORE_OK) return rc;
  if ((rc = index_begin(e)) != EPPSCORE_OK) return rc;
  CK(e, e->c_pick.reserve((size_t)R * 4));
  CK(e, e->c_hash.reserve((size_t)R * hash_stride * 8));
  CK(e, e->c_nh.reserve((size_t)R * 2));
  CK(e, cudaMemcpyAsync(e->c_pick.p, pick, (size_t)R * 4, cudaMemcpyHostToDevice, e->stream));
  CK(e, cudaMemcpyAsync(e->c_hash.p, hashes, (size_t)R * hash_stride * 8, cudaMemcpyHostToDevice, e->stream));
  CK(e, cudaMemcpyAsync(e->c_nh.p, n_hashes, (size_t)R * 2, cudaMemcpyHostToDevice, e->stream));
  rc = index_status(e, e->index->commit(R, e->c_pick.as<int32_t>(), e->c_hash.as<uint64_t>(), e->c_nh.as<uint16_t>(), hash_stride,
                                        caps.empty() ? nullptr : caps.data(), 0, lo, hi - lo + 1, std::max<int64_t>(touches, 1), e->stream),
                    "commit_picks");
  if (rc != EPPSCORE_OK) return rc;
  CK(e, cudaStreamSynchronize(e->stream));  // the caller's host arrays (pageable) may be reused
  return index_end(e);
}

int32_t eppscore_commit_picks_device(eppscore_engine* e, int32_t R, const int32_t* pick, const uint64_t* hashes,
                                     const uint16_t* n_hashes, int32_t hash_stride, const int32_t* lru_capacity,
                                     int64_t touch_bound, void* stream) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R < 0 || (R > 0 && (!pick || !hashes || !n_hashes))) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (R == 0) return EPPSCORE_OK;
  if (hash_stride <= 0) return fail(e, EPPSCORE_ERR_INVALID, "hash_stride must be > 0");
  CK(e, cudaSetDevice(e->device));
  std::vector<int32_t> caps;
  int32_t rc = pad_caps(e, lru_capacity, &caps);
  if (rc != EPPSCORE_OK) return rc;
  cudaStream_t cs = stream ? (cudaStream_t)stream : e->stream;
  if (cs != e->stream) {  // the arrays are produced on the caller's stream
    CK(e, cudaEventRecord(e->ev_caller, cs));
    CK(e, cudaStreamWaitEvent(e->stream, e->ev_caller, 0));
  }
  if ((rc = index_begin(e)) != EPPSCORE_OK) return rc;
  const int32_t n_eps = e->M > 0 ? e->M : e->geo.Mpad;
  rc = index_status(e, e->index->commit(R, pick, hashes, n_hashes, hash_stride, caps.empty() ? nullptr : caps.data(), 0, 0, n_eps,
                                        touch_bound, e->stream),
                    "commit_picks_device");
  if (rc != EPPSCORE_OK) return rc;
  if ((rc = index_end(e)) != EPPSCORE_OK) return rc;
  if (cs != e->stream) CK(e, cudaStreamWaitEvent(cs, e->ev_table, 0));  // stream-ordered for the caller
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_add(eppscore_engine* e, const uint64_t* hashes, int32_t n, int32_t endpoint, int32_t lru_capacity) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (n < 0 || (n > 0 && !hashes)) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (n > EPPSCORE_MAX_BLOCKS) return fail(e, EPPSCORE_ERR_CAPACITY, "more than 65535 hashes in one Add");
  if (endpoint < 0 || endpoint >= e->geo.Mpad) return fail(e, EPPSCORE_ERR_INVALID, "endpoint out of range");
  CK(e, cudaSetDevice(e->device));
  int32_t rc = index_begin(e);
  if (rc != EPPSCORE_OK) return rc;
  const int32_t stride = n > 0 ? n : 1;
  const uint16_t nh = (uint16_t)n;
  CK(e, e->c_pick.reserve(4));
  CK(e, e->c_hash.reserve((size_t)stride * 8));
  CK(e, e->c_nh.reserve(2));
  CK(e, cudaMemcpyAsync(e->c_pick.p, &endpoint, 4, cudaMemcpyHostToDevice, e->stream));
  if (n) CK(e, cudaMemcpyAsync(e->c_hash.p, hashes, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
  CK(e, cudaMemcpyAsync(e->c_nh.p, &nh, 2, cudaMemcpyHostToDevice, e->stream));
  rc = index_status(e, e->index->commit(1, e->c_pick.as<int32_t>(), e->c_hash.as<uint64_t>(), e->c_nh.as<uint16_t>(), stride, nullptr,
                                        lru_capacity, endpoint, 1, std::max(n, 1), e->stream),
                    "prefix_add");
  if (rc != EPPSCORE_OK) return rc;
  CK(e, cudaStreamSynchronize(e->stream));
  return index_end(e);
}

int32_t eppscore_prefix_apply(eppscore_engine* e, int64_t n, const uint64_t* hash, const int32_t* endpoint, const uint8_t* op) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (n < 0 || (n > 0 && (!hash || !endpoint || !op))) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (n == 0) return EPPSCORE_OK;
  for (int64_t i = 0; i < n; i++)
    if (endpoint[i] < 0 || endpoint[i] >= e->geo.Mpad) return fail(e, EPPSCORE_ERR_CAPACITY, "prefix_apply: endpoint out of range");
  CK(e, cudaSetDevice(e->device));
  int32_t rc = index_begin(e);
  if (rc != EPPSCORE_OK) return rc;
  CK(e, e->c_hash.reserve((size_t)n * 8));
  CK(e, e->c_ep.reserve((size_t)n * 4));
  CK(e, e->c_op.reserve((size_t)n));
  CK(e, cudaMemcpyAsync(e->c_hash.p, hash, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
  CK(e, cudaMemcpyAsync(e->c_ep.p, endpoint, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  CK(e, cudaMemcpyAsync(e->c_op.p, op, (size_t)n, cudaMemcpyHostToDevice, e->stream));
  rc = index_status(e, e->index->apply(n, e->c_hash.as<uint64_t>(), e->c_ep.as<int32_t>(), e->c_op.as<uint8_t>(), e->stream), "prefix_apply");
  if (rc != EPPSCORE_OK) return rc;
  CK(e, cudaStreamSynchronize(e->stream));
  return index_end(e);
}

int32_t eppscore_prefix_remove_endpoint(eppscore_engine* e, int32_t endpoint) {
  if (!e) return EPPSCORE_ERR_INVALID;
  CK(e, cudaSetDevice(e->device));
  int32_t rc = index_begin(e);
  if (rc != EPPSCORE_OK) return rc;
  rc = index_status(e, e->index->remove_endpoint(endpoint, e->stream), "prefix_remove_endpoint");
  if (rc != EPPSCORE_OK) return rc;
  return index_end(e);
}

int32_t eppscore_prefix_lru_len(const eppscore_engine* ce, int32_t endpoint) {
  eppscore_engine* e = const_cast<eppscore_engine*>(ce);
  if (!e) return -1;
  if (cudaSetDevice(e->device) != cudaSuccess) return -1;
  int32_t len = -1;
  if (e->index->lru_keys(endpoint, nullptr, 0, &len, e->stream) != cudaSuccess) return -1;
  return len;
}
int32_t eppscore_prefix_lru_keys(const eppscore_engine* ce, int32_t endpoint, uint64_t* out, int32_t cap) {
  eppscore_engine* e = const_cast<eppscore_engine*>(ce);
  if (!e) return -1;
  if (cudaSetDevice(e->device) != cudaSuccess) return -1;
  int32_t len = -1;
  if (e->index->lru_keys(endpoint, out, cap, &len, e->stream) != cudaSuccess) return -1;
  return len;
}

int32_t eppscore_prefix_get(eppscore_engine* e, uint64_t hash, uint32_t* bitset_out, int32_t words) {
  if (!e) return EPPSCORE_ERR_INVALID;
  CK(e, cudaSetDevice(e->device));
  int32_t count = 0;
  CK(e, e->index->get(hash, bitset_out, bitset_out ? words : 0, &count, e->stream));
  return count;
}

void* eppscore_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void eppscore_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
