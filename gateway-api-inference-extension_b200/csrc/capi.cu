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
  cudaEvent_t ev_snapshot = nullptr, ev_table = nullptr;
  cudaStream_t snapshot_stream = nullptr;        // stream the last snapshot preparation ran on
  unsigned long long snapshot_capture_id = 0;    // != 0: ev_snapshot was recorded inside that CUDA-graph capture
  std::string err;
  uint64_t launches = 0;
  bool force_generic = false;  // EPPSCORE_FORCE_GENERIC=1 / eppscore_set_debug(1): skip the specialised kernels
  int32_t hash_stage_mask = 3;  // eppscore_set_debug(2): profiling only

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

  // prefix table
  std::unique_ptr<PrefixIndex> index;
  Slot* d_slots = nullptr;
  uint32_t* d_rows = nullptr;
  bool table_adopted = false;
  int64_t dev_capacity = 0;  // capacity (hashes) the device slot/row buffers were allocated for
  DevBuf st_idx, st_val, st_slot;
  DevBuf probe_out;

  // scratch for host-location batches and internal hashes
  DevBuf s_prompts, s_off, s_len, s_seed, s_hashes, s_nh, s_adapter, s_mask, s_dense, s_dtotal, s_pick, s_score,
      s_tie, s_match, s_total, s_scores, s_intok, s_tslo, s_pslo, s_pred, s_fields, s_fmask;
};

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

int32_t flush_table(eppscore_engine* e) {
  PrefixIndex* ix = e->index.get();
  if (!ix || e->table_adopted) return EPPSCORE_OK;
  if (!ix->full_upload_needed() && ix->dirty_slots().empty() && ix->dirty_words().empty()) return EPPSCORE_OK;
  if (ix->capacity_rows() != e->dev_capacity) {  // the host index grew: reallocate the device table, then upload it all
    CK(e, cudaStreamSynchronize(e->stream));
    if (e->d_slots) cudaFree(e->d_slots);
    if (e->d_rows) cudaFree(e->d_rows);
    e->d_slots = nullptr;
    e->d_rows = nullptr;
    const size_t rb = ((size_t)ix->capacity_rows() + 1) * e->geo.row_words * 4;
    CK(e, cudaMalloc(&e->d_slots, ix->slots().size() * sizeof(Slot)));
    CK(e, cudaMalloc(&e->d_rows, rb));
    CK(e, cudaMemsetAsync(e->d_rows, 0, rb, e->stream));
    e->dev_capacity = ix->capacity_rows();
    ix->mark_full_upload();
  }
  const size_t nrow_words = ix->rows().size();
  const size_t nds = ix->dirty_slots().size(), ndw = ix->dirty_words().size();
  const bool full = ix->full_upload_needed() || (nds + ndw) * 8 > ix->slots().size() + nrow_words;
  if (full) {
    CK(e, cudaMemcpyAsync(e->d_slots, ix->slots().data(), ix->slots().size() * sizeof(Slot), cudaMemcpyHostToDevice,
                          e->stream));
    if (nrow_words)
      CK(e, cudaMemcpyAsync(e->d_rows, ix->rows().data(), nrow_words * 4, cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
  } else {
    if (ndw) {
      std::vector<uint32_t> vals(ndw);
      for (size_t i = 0; i < ndw; i++) vals[i] = ix->rows()[ix->dirty_words()[i]];
      CK(e, e->st_idx.reserve(ndw * 4));
      CK(e, e->st_val.reserve(ndw * 4));
      CK(e, cudaMemcpyAsync(e->st_idx.p, ix->dirty_words().data(), ndw * 4, cudaMemcpyHostToDevice, e->stream));
      CK(e, cudaMemcpyAsync(e->st_val.p, vals.data(), ndw * 4, cudaMemcpyHostToDevice, e->stream));
      e->launches += launch_scatter_u32(e->d_rows, e->st_idx.as<uint32_t>(), e->st_val.as<uint32_t>(), (int64_t)ndw,
                                        e->stream);
      CK(e, cudaStreamSynchronize(e->stream));  // staging vectors are pageable and reused
    }
    if (nds) {
      std::vector<Slot> vals(nds);
      for (size_t i = 0; i < nds; i++) vals[i] = ix->slots()[ix->dirty_slots()[i]];
      CK(e, e->st_idx.reserve(nds * 4));
      CK(e, e->st_slot.reserve(nds * sizeof(Slot)));
      CK(e, cudaMemcpyAsync(e->st_idx.p, ix->dirty_slots().data(), nds * 4, cudaMemcpyHostToDevice, e->stream));
      CK(e, cudaMemcpyAsync(e->st_slot.p, vals.data(), nds * sizeof(Slot), cudaMemcpyHostToDevice, e->stream));
      e->launches += launch_scatter_slots(e->d_slots, e->st_idx.as<uint32_t>(), e->st_slot.as<Slot>(), (int64_t)nds,
                                          e->stream);
      CK(e, cudaStreamSynchronize(e->stream));
    }
  }
  CK(e, cudaGetLastError());
  ix->clear_dirty();
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
int32_t schedule_device(eppscore_engine* e, const DevBatch& b, cudaStream_t s) {
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
      CK(e, e->s_fields.reserve((size_t)b.R * 4));
      e->launches += launch_count_fields(b.prompt_bytes, b.prompt_off, b.prompt_len, b.R, e->s_fields.as<int32_t>(), s, e->sm_count);
      a.lat.input_tokens = e->s_fields.as<int32_t>();
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
          CK(e, e->s_hashes.reserve((size_t)b.R * mb * 8));
          hashes = e->s_hashes.as<uint64_t>();
        }
        CK(e, e->s_nh.reserve((size_t)b.R * 2));
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
        h.n_hashes = e->s_nh.as<uint16_t>();
        h.stage_mask = e->hash_stage_mask;
        e->launches += launch_hash_prompts(h, s, e->sm_count);
        a.hashes = hashes;
        a.n_hashes = h.n_hashes;
        a.hash_stride = mb;
      }
      if (e->index && (e->index->n_rows() > 0 || e->table_adopted)) {
        a.slots = e->d_slots;
        a.slot_mask = e->index->slot_mask();
        a.rows = e->d_rows;
      }
    }
  }
  if (wait_snapshot) CK(e, cudaStreamWaitEvent(s, e->ev_snapshot, 0));
  // dispatch: specialised fast paths first, the fully general kernels otherwise
  int launched = 0;
  if (!e->force_generic) launched = dense ? launch_score_dense_fast(a, s, e->sm_count) : launch_pick_sparse(a, s, e->sm_count);
  if (launched == 0) launched = launch_score_pick(a, dense, s, e->sm_count);
  e->launches += launched;
  CK(e, cudaGetLastError());
  return EPPSCORE_OK;
}

template <typename T>
int32_t h2d(eppscore_engine* e, DevBuf& buf, const T* src, size_t count, const T** out) {
  *out = nullptr;
  if (!src || count == 0) return EPPSCORE_OK;
  CK(e, buf.reserve(count * sizeof(T)));
  CK(e, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, e->stream));
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
  if (((uint64_t)e->cfg.prefix_capacity + 1) * (uint64_t)e->geo.row_words >= (1ULL << 32))
    return fail(nullptr, EPPSCORE_ERR_CAPACITY, "prefix_capacity * row_words must be < 2^32");
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
  // prefix table
  ep->index = std::make_unique<PrefixIndex>(ep->geo, ep->cfg.prefix_capacity, ep->cfg.lru_capacity_default);
  const size_t nslots = ep->index->slots().size();
  CK(nullptr, cudaMalloc(&ep->d_slots, nslots * sizeof(Slot)));
  CK(nullptr, cudaMalloc(&ep->d_rows, ((size_t)ep->cfg.prefix_capacity + 1) * ep->geo.row_words * 4));
  CK(nullptr, cudaMemsetAsync(ep->d_slots, 0xFF, nslots * sizeof(Slot), ep->stream));
  CK(nullptr, cudaMemsetAsync(ep->d_rows, 0, ((size_t)ep->cfg.prefix_capacity + 1) * ep->geo.row_words * 4, ep->stream));
  ep->dev_capacity = ep->index->capacity_rows();
  CK(nullptr, ep->probe_out.reserve((size_t)(2 + ep->geo.row_words) * 4));
  CK(nullptr, cudaEventRecord(ep->ev_table, ep->stream));
  CK(nullptr, cudaEventRecord(ep->ev_snapshot, ep->stream));
  CK(nullptr, cudaStreamSynchronize(ep->stream));
  ep->index->clear_dirty();
  *out = e.release();
  return EPPSCORE_OK;
}

void eppscore_destroy(eppscore_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  DevBuf* bufs[] = {&e->s_fmask, &e->s_fields, &e->qhdr[0], &e->qhdr[1], &e->qbucket[0], &e->qbucket[1], &e->raw_min_tpot, &e->raw_dispatched, &e->raw_prefill, &e->raw_tokens, &e->lat_ep,
                    &e->s_intok, &e->s_tslo, &e->s_pslo, &e->s_pred, &e->raw_kv, &e->raw_queue, &e->raw_running, &e->raw_act, &e->raw_wait, &e->raw_nmodels, &e->raw_max,
                    &e->fold_unmasked, &e->fold_masked, &e->cls_lo, &e->cls_hi, &e->summ, &e->tiemask, &e->prefix_lut2d, &e->st_idx, &e->st_val, &e->st_slot,
                    &e->probe_out, &e->s_prompts, &e->s_off, &e->s_len, &e->s_seed, &e->s_hashes, &e->s_nh, &e->s_adapter,
                    &e->s_mask, &e->s_dense, &e->s_dtotal, &e->s_pick, &e->s_score, &e->s_tie, &e->s_match, &e->s_total, &e->s_scores};
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 4; i++) e->raw_col[i].release();
  for (int i = 0; i < kMaxSteps; i++) e->term[i].release();
  if (e->d_slots) cudaFree(e->d_slots);
  if (e->d_rows) cudaFree(e->d_rows);
  if (e->ev_snapshot) cudaEventDestroy(e->ev_snapshot);
  if (e->ev_table) cudaEventDestroy(e->ev_table);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* eppscore_last_error(const eppscore_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int32_t eppscore_get_stats(const eppscore_engine* e, eppscore_stats* out) {
  if (!e || !out) return EPPSCORE_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(*out);
  out->M = e->M;
  out->epoch = e->epoch;
  out->kernel_launches = e->launches;
  out->prefix_hashes = e->index->n_keys();
  out->prefix_live_hashes = e->index->n_live();
  out->prefix_capacity = e->index->capacity_rows();
  out->prefix_table_bytes = (int64_t)(e->index->slots().size() * sizeof(Slot)) +
                            (int64_t)e->index->n_rows() * e->geo.row_words * 4;
  out->lru_entries = e->index->lru_entries();
  return EPPSCORE_OK;
}

int32_t eppscore_set_debug(eppscore_engine* e, int32_t key, int64_t value) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (key == 1) {
    e->force_generic = value != 0;
    return EPPSCORE_OK;
  }
  if (key == 2) {
    e->hash_stage_mask = (int32_t)(value & 3);
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
  int32_t rc = flush_table(e);
  if (rc != EPPSCORE_OK) return rc;
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
    return schedule_device(e, d, b->stream ? (cudaStream_t)b->stream : e->stream);
  }
  // ---- host buffers: H2D, kernels, D2H, all inside this call ----
  if (!e->have_snapshot) return fail(e, EPPSCORE_ERR_NO_SNAPSHOT, "schedule_batch before set_snapshot");
  const size_t R = (size_t)b->R, M = (size_t)e->M;
  const size_t mw = (M + 31) / 32;
  const int32_t mb = b->max_blocks > 0 ? b->max_blocks : e->cfg.max_blocks;
  if (b->prompt_bytes) {
    if (!b->prompt_off) return fail(e, EPPSCORE_ERR_INVALID, "prompt_off required with prompt_bytes");
    size_t total = (size_t)b->prompt_off[R];
    if (b->prompt_len) {
      total = 0;
      for (size_t r = 0; r < R; r++) total = std::max(total, (size_t)b->prompt_off[r] + (size_t)b->prompt_len[r]);
    }
    // pad the device copy so 16-byte vector loads of the last block never leave the allocation
    CK(e, e->s_prompts.reserve(total + 64));
    if (total) CK(e, cudaMemcpyAsync(e->s_prompts.p, b->prompt_bytes, total, cudaMemcpyHostToDevice, e->stream));
    d.prompt_bytes = e->s_prompts.as<uint8_t>();
  }
#define H2D(buf, field, count) \
  if ((rc = h2d(e, buf, b->field, (count), &d.field)) != EPPSCORE_OK) return rc;
  H2D(e->s_off, prompt_off, R + 1)
  H2D(e->s_len, prompt_len, R)
  H2D(e->s_seed, model_seed, R)
  if (b->hashes_in) {
    if ((rc = h2d(e, e->s_hashes, b->hashes_in, R * (size_t)b->hash_stride, &d.hashes_in)) != EPPSCORE_OK) return rc;
    H2D(e->s_nh, n_hashes_in, R)
  }
  H2D(e->s_adapter, adapter_id, R)
  H2D(e->s_mask, cand_mask, R * mw)
  H2D(e->s_dense, dense_feat, R * M * 4)
  H2D(e->s_dtotal, dense_total, R)
  H2D(e->s_intok, input_tokens, R)
  H2D(e->s_tslo, ttft_slo, R)
  H2D(e->s_pslo, tpot_slo, R)
#undef H2D
  CK(e, e->s_pick.reserve(R * 4));
  CK(e, e->s_score.reserve(R * 8));
  CK(e, e->s_tie.reserve(R * 4));
  d.pick = e->s_pick.as<int32_t>();
  d.pick_score = e->s_score.as<double>();
  d.tie_count = e->s_tie.as<int32_t>();
  if (b->match_blocks) {
    CK(e, e->s_match.reserve(R * M * 2));
    d.match_blocks = e->s_match.as<uint16_t>();
  }
  if (b->total_blocks) {
    CK(e, e->s_total.reserve(R * 2));
    d.total_blocks = e->s_total.as<uint16_t>();
  }
  if (b->scores_out) {
    CK(e, e->s_scores.reserve(R * M * 8));
    d.scores_out = e->s_scores.as<double>();
  }
  if (b->pred_out) {
    CK(e, e->s_pred.reserve(R * M * 16));
    d.pred_out = e->s_pred.as<double>();
  }
  if (b->filter_mask_out) {
    CK(e, e->s_fmask.reserve(R * mw * 4));
    d.filter_mask_out = e->s_fmask.as<uint32_t>();
  }
  uint64_t* hashes_dev = nullptr;
  if (b->hashes_out && !b->hashes_in) {
    // reuse the internal hash scratch as the device-side hashes_out
    CK(e, e->s_hashes.reserve(R * (size_t)mb * 8));
    hashes_dev = e->s_hashes.as<uint64_t>();
    d.hashes_out = hashes_dev;
  }
  rc = schedule_device(e, d, e->stream);
  if (rc != EPPSCORE_OK) return rc;
  if (b->pick) CK(e, cudaMemcpyAsync(b->pick, d.pick, R * 4, cudaMemcpyDeviceToHost, e->stream));
  if (b->pick_score) CK(e, cudaMemcpyAsync(b->pick_score, d.pick_score, R * 8, cudaMemcpyDeviceToHost, e->stream));
  if (b->tie_count) CK(e, cudaMemcpyAsync(b->tie_count, d.tie_count, R * 4, cudaMemcpyDeviceToHost, e->stream));
  if (b->match_blocks) CK(e, cudaMemcpyAsync(b->match_blocks, d.match_blocks, R * M * 2, cudaMemcpyDeviceToHost, e->stream));
  if (b->total_blocks) CK(e, cudaMemcpyAsync(b->total_blocks, d.total_blocks, R * 2, cudaMemcpyDeviceToHost, e->stream));
  if (b->scores_out) CK(e, cudaMemcpyAsync(b->scores_out, d.scores_out, R * M * 8, cudaMemcpyDeviceToHost, e->stream));
  if (b->pred_out) CK(e, cudaMemcpyAsync(b->pred_out, d.pred_out, R * M * 16, cudaMemcpyDeviceToHost, e->stream));
  if (b->filter_mask_out) CK(e, cudaMemcpyAsync(b->filter_mask_out, d.filter_mask_out, R * mw * 4, cudaMemcpyDeviceToHost, e->stream));
  if (hashes_dev) CK(e, cudaMemcpyAsync(b->hashes_out, hashes_dev, R * (size_t)mb * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaStreamSynchronize(e->stream));
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
  CK(e, e->s_prompts.reserve(total + 64));
  if (total) CK(e, cudaMemcpyAsync(e->s_prompts.p, prompt_bytes, total, cudaMemcpyHostToDevice, e->stream));
  h.bytes = e->s_prompts.as<uint8_t>();
  if ((rc = h2d(e, e->s_off, prompt_off, (size_t)R + 1, &h.off)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->s_len, prompt_len, (size_t)R, &h.len)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->s_seed, model_seed, (size_t)R, &h.seed)) != EPPSCORE_OK) return rc;
  CK(e, e->s_hashes.reserve((size_t)R * mb * 8));
  CK(e, e->s_nh.reserve((size_t)R * 2));
  CK(e, cudaMemsetAsync(e->s_hashes.p, 0, (size_t)R * mb * 8, e->stream));
  h.hashes = e->s_hashes.as<uint64_t>();
  h.n_hashes = e->s_nh.as<uint16_t>();
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
  CK(e, e->s_prompts.reserve(total + 64));
  if (total) CK(e, cudaMemcpyAsync(e->s_prompts.p, prompt_bytes, total, cudaMemcpyHostToDevice, e->stream));
  const int64_t* d_off = nullptr;
  const int32_t* d_len = nullptr;
  int32_t rc;
  if ((rc = h2d(e, e->s_off, prompt_off, (size_t)R + 1, &d_off)) != EPPSCORE_OK) return rc;
  if ((rc = h2d(e, e->s_len, prompt_len, (size_t)R, &d_len)) != EPPSCORE_OK) return rc;
  CK(e, e->s_fields.reserve((size_t)R * 4));
  e->launches += launch_count_fields(e->s_prompts.as<uint8_t>(), d_off, d_len, R, e->s_fields.as<int32_t>(), e->stream, e->sm_count);
  CK(e, cudaGetLastError());
  CK(e, cudaMemcpyAsync(out, e->s_fields.p, (size_t)R * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaStreamSynchronize(e->stream));
  return EPPSCORE_OK;
}

int32_t eppscore_commit_picks(eppscore_engine* e, int32_t R, const int32_t* pick, const uint64_t* hashes,
                              const uint16_t* n_hashes, int32_t hash_stride, const int32_t* lru_capacity) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R < 0 || (R > 0 && (!pick || !hashes || !n_hashes))) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (e->table_adopted) return fail(e, EPPSCORE_ERR_INVALID, "engine holds an adopted (read-only) table image");
  for (int32_t r = 0; r < R; r++) {
    const int32_t ep = pick[r];
    if (ep < 0) continue;  // no target endpoint: nothing to record (plugin.go:173-175)
    if (ep >= e->geo.Mpad) return fail(e, EPPSCORE_ERR_INVALID, "pick out of range");
    const int32_t cap = lru_capacity ? lru_capacity[ep] : 0;  // makeserver, plugin.go:207-216
    if (!e->index->add(hashes + (size_t)r * hash_stride, n_hashes[r], ep, cap))
      return fail(e, EPPSCORE_ERR_CAPACITY, "prefix table cannot grow further (row index space exhausted)");
  }
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_add(eppscore_engine* e, const uint64_t* hashes, int32_t n, int32_t endpoint, int32_t lru_capacity) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (n < 0 || (n > 0 && !hashes)) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (endpoint < 0 || endpoint >= e->geo.Mpad) return fail(e, EPPSCORE_ERR_INVALID, "endpoint out of range");
  if (e->table_adopted) return fail(e, EPPSCORE_ERR_INVALID, "engine holds an adopted (read-only) table image");
  if (!e->index->add(hashes, n, endpoint, lru_capacity)) return fail(e, EPPSCORE_ERR_CAPACITY, "prefix table full");
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_apply(eppscore_engine* e, int64_t n, const uint64_t* hash, const int32_t* endpoint, const uint8_t* op) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (n < 0 || (n > 0 && (!hash || !endpoint || !op))) return fail(e, EPPSCORE_ERR_INVALID, "NULL argument");
  if (e->table_adopted) return fail(e, EPPSCORE_ERR_INVALID, "engine holds an adopted (read-only) table image");
  for (int64_t i = 0; i < n; i++)
    if (!e->index->apply(hash[i], endpoint[i], op[i])) return fail(e, EPPSCORE_ERR_CAPACITY, "prefix table full or endpoint out of range");
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_remove_endpoint(eppscore_engine* e, int32_t endpoint) {
  if (!e) return EPPSCORE_ERR_INVALID;
  e->index->remove_endpoint(endpoint);
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_lru_len(const eppscore_engine* e, int32_t endpoint) { return e ? e->index->lru_len(endpoint) : -1; }
int32_t eppscore_prefix_lru_keys(const eppscore_engine* e, int32_t endpoint, uint64_t* out, int32_t cap) {
  return e ? e->index->lru_keys(endpoint, out, cap) : -1;
}

}  // extern "C"

// single-thread probe of the DEVICE table (so eppscore_prefix_get checks what the kernels see)
__global__ void probe_one_kernel(const Slot* slots, uint64_t mask, const uint32_t* rows, int rw, uint64_t h, uint32_t* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t cnt = 0, row = kEmptyRow;
  for (uint64_t i = h & mask;; i = (i + 1) & mask) {
    const Slot s = slots[i];
    if (s.row == kEmptyRow) break;
    if (s.key == h) {
      cnt = s.cnt;
      row = s.row;
      break;
    }
  }
  out[0] = cnt;
  out[1] = row;
  for (int w = 0; w < rw; w++) out[2 + w] = (row != kEmptyRow) ? rows[(size_t)row * rw + w] : 0u;
}

extern "C" {

int32_t eppscore_prefix_get(eppscore_engine* e, uint64_t hash, uint32_t* bitset_out, int32_t words) {
  if (!e) return EPPSCORE_ERR_INVALID;
  CK(e, cudaSetDevice(e->device));
  int32_t rc = flush_table(e);
  if (rc != EPPSCORE_OK) return rc;
  const int rw = e->geo.row_words;
  probe_one_kernel<<<1, 32, 0, e->stream>>>(e->d_slots, e->index->slot_mask(), e->d_rows, rw, hash, e->probe_out.as<uint32_t>());
  e->launches++;
  std::vector<uint32_t> host((size_t)rw + 2);
  CK(e, cudaMemcpyAsync(host.data(), e->probe_out.p, host.size() * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(e, cudaStreamSynchronize(e->stream));
  if (bitset_out) {
    for (int32_t w = 0; w < words; w++) bitset_out[w] = 0;
    for (int32_t m = 0; m < e->geo.Mpad && (m >> 5) < words; m++) {
      const uint32_t pos = perm_bitpos((uint32_t)m, e->geo.log_epl);
      if ((host[2 + (pos >> 5)] >> (pos & 31)) & 1u) bitset_out[m >> 5] |= 1u << (m & 31);
    }
  }
  return (int32_t)host[0];
}

int32_t eppscore_prefix_image_info(eppscore_engine* e, void** slots_dev, int64_t* slots_bytes, void** rows_dev,
                                   int64_t* rows_bytes, int64_t* meta) {
  if (!e) return EPPSCORE_ERR_INVALID;
  CK(e, cudaSetDevice(e->device));
  int32_t rc = flush_table(e);
  if (rc != EPPSCORE_OK) return rc;
  if (slots_dev) *slots_dev = e->d_slots;
  if (slots_bytes) *slots_bytes = (int64_t)(e->index->slots().size() * sizeof(Slot));
  if (rows_dev) *rows_dev = e->d_rows;
  if (rows_bytes) *rows_bytes = e->index->n_rows() * (int64_t)e->geo.row_words * 4;
  if (meta) {
    meta[0] = (int64_t)e->index->slots().size();
    meta[1] = e->geo.row_words;
    meta[2] = e->index->n_rows();
    meta[3] = e->index->n_live();
  }
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_image_adopt(eppscore_engine* e, const int64_t* meta) {
  if (!e || !meta) return EPPSCORE_ERR_INVALID;
  if (meta[0] != (int64_t)e->index->slots().size() || meta[1] != e->geo.row_words)
    return fail(e, EPPSCORE_ERR_INVALID, "image geometry differs (engines must share max_endpoints and prefix_capacity)");
  if (meta[2] > e->index->capacity_rows() + 1) return fail(e, EPPSCORE_ERR_CAPACITY, "image has more rows than prefix_capacity");
  e->index->adopt_counts(meta[2], meta[3]);
  e->index->clear_dirty();
  e->table_adopted = true;
  CK(e, cudaEventRecord(e->ev_table, e->stream));
  return EPPSCORE_OK;
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
