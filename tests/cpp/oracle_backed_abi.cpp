// oracle_backed_abi.cpp — TEST INFRASTRUCTURE ONLY: the subset of include/eppscore.h that the C++ host mirror
// (gateway-api-inference-extension_b200/host/) calls, implemented on the CPU oracle (oracle/oracle.h).  It lets the host
// layer's own test program (host/host_test.cpp: the reference's scheduler tests restated, the coalescing front over the real
// Scheduler class) run in the `-m "not gpu"` suite: what is tested there is the HOST code — snapshot packing, adapter
// dictionary, filter masks, prompt packing, result hand-back, PreRequest, the small-batch host route against the engine
// route — with the oracle standing in for the engine behind the same C ABI.  Never linked into the product; the product's
// libeppscore.so has no CPU path (tests/test_abi_cpu.py::test_no_cpu_fallback).
#include <cstring>
#include <string>
#include <vector>

#include "../../include/eppscore.h"
#include "../../oracle/oracle.h"

struct eppscore_engine {
  eppscore_config cfg;
  orc_profile prof;
  orc_latency_params lat;
  bool have_lat = false;
  orc_index* idx = nullptr;
  // owned copy of the snapshot (the real engine copies host snapshots into its tile buffers too)
  int32_t M = 0, words = 1;
  std::vector<double> kv, min_tpot, col[EPPSCORE_MAX_ENDPOINT_COLS];
  std::vector<int64_t> queue, running, tokens;
  std::vector<uint64_t> act, wait;
  std::vector<int32_t> nm, mx, dispatched;
  std::vector<uint8_t> prefill;
  bool have_running = false, have_lora = false, have_min_tpot = false, have_dispatched = false, have_prefill = false,
       have_tokens = false, have_col[EPPSCORE_MAX_ENDPOINT_COLS] = {};
  bool have_snapshot = false;
  std::string err;
};

static std::string g_create_err;

static int32_t fail(eppscore_engine* e, int32_t code, const char* msg) {
  if (e) e->err = msg;
  else g_create_err = msg;
  return code;
}

extern "C" {

int32_t eppscore_abi_version(void) { return 3; }

void eppscore_config_default(eppscore_config* c) {  // the same defaults as csrc/capi.cu (loader/defaults.go:46-103, types.go:91-112)
  memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(*c);
  c->n_scorers = 3;
  c->scorer_kind[0] = EPPSCORE_SCORER_QUEUE;
  c->scorer_weight[0] = 2.0;
  c->scorer_kind[1] = EPPSCORE_SCORER_KV_CACHE;
  c->scorer_weight[1] = 2.0;
  c->scorer_kind[2] = EPPSCORE_SCORER_PREFIX;
  c->scorer_weight[2] = 3.0;
  c->block_chars = 16 * 4;
  c->max_blocks = 256;
  c->tie_mode = EPPSCORE_TIE_LOWEST_INDEX;
  c->max_endpoints = 1024;
  c->max_adapters = 64;
  c->prefix_capacity = 1 << 18;
  c->lru_capacity_default = 31250;
  c->token_load_threshold = 4194304.0;
}

void eppscore_latency_params_default(eppscore_latency_params* p) {  // predictedlatency/plugin.go:118-136, scorer/latency/plugin.go:59-90
  memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(*p);
  p->has_predictions = 1;
  p->slo_buffer_factor = 1.0;
  p->ttft_weight = 0.8;
  p->tpot_weight = 0.2;
  p->composite_kv = p->composite_queue = p->composite_prefix = 1.0;
}

int32_t eppscore_create(int32_t, const eppscore_config* c, eppscore_engine** out) {
  if (!c || !out || c->struct_size != sizeof(*c)) return fail(nullptr, EPPSCORE_ERR_INVALID, "config NULL or struct_size mismatch");
  if (c->n_scorers < 0 || c->n_scorers > EPPSCORE_MAX_SCORERS) return fail(nullptr, EPPSCORE_ERR_INVALID, "n_scorers out of range");
  auto* e = new eppscore_engine();
  e->cfg = *c;
  memset(&e->prof, 0, sizeof(e->prof));
  e->prof.n_scorers = c->n_scorers;
  for (int i = 0; i < c->n_scorers; i++) {
    e->prof.scorer_kind[i] = c->scorer_kind[i];  // same numbering (oracle.h restates it)
    e->prof.scorer_weight[i] = c->scorer_weight[i];
  }
  e->prof.tie_mode = c->tie_mode;
  e->prof.tie_seed = c->tie_seed;
  e->prof.token_load_threshold = c->token_load_threshold;
  e->prof.pick_mode = c->pick_mode;
  e->prof.n_filters = c->n_filters;
  for (int f = 0; f < c->n_filters && f < ORC_MAX_FILTERS; f++) {
    e->prof.filter_kind[f] = c->filter_kind[f];
    for (int q = 0; q < 3; q++) e->prof.filter_param[f][q] = c->filter_param[f][q];
  }
  e->idx = orc_index_new(c->lru_capacity_default > 0 ? c->lru_capacity_default : 31250);
  *out = e;
  return EPPSCORE_OK;
}

void eppscore_destroy(eppscore_engine* e) {
  if (!e) return;
  orc_index_free(e->idx);
  delete e;
}

const char* eppscore_last_error(const eppscore_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int32_t eppscore_set_latency_params(eppscore_engine* e, const eppscore_latency_params* p) {
  if (!e || !p || p->struct_size != sizeof(*p)) return fail(e, EPPSCORE_ERR_INVALID, "latency params NULL or struct_size mismatch");
  orc_latency_params& l = e->lat;
  memset(&l, 0, sizeof(l));
  l.ttft_intercept = p->ttft_intercept; l.ttft_kv = p->ttft_kv; l.ttft_input = p->ttft_input;
  l.ttft_waiting = p->ttft_waiting; l.ttft_running = p->ttft_running; l.ttft_prefix = p->ttft_prefix;
  l.tpot_intercept = p->tpot_intercept; l.tpot_kv = p->tpot_kv; l.tpot_input = p->tpot_input;
  l.tpot_waiting = p->tpot_waiting; l.tpot_running = p->tpot_running; l.tpot_generated = p->tpot_generated;
  l.slo_buffer_factor = p->slo_buffer_factor;
  l.streaming_mode = p->streaming_mode;
  l.has_predictions = p->has_predictions;
  l.ttft_weight = p->ttft_weight; l.tpot_weight = p->tpot_weight;
  l.strategy_most = p->strategy_most;
  l.composite_kv = p->composite_kv; l.composite_queue = p->composite_queue; l.composite_prefix = p->composite_prefix;
  e->have_lat = true;
  return EPPSCORE_OK;
}

int32_t eppscore_set_snapshot(eppscore_engine* e, const eppscore_snapshot* s) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (!s || s->struct_size != sizeof(*s)) return fail(e, EPPSCORE_ERR_INVALID, "snapshot NULL or struct_size mismatch");
  if (s->location != 0) return fail(e, EPPSCORE_ERR_INVALID, "the stand-in takes host snapshots only");
  if (s->M < 0 || s->M > e->cfg.max_endpoints) return fail(e, EPPSCORE_ERR_CAPACITY, "M exceeds config.max_endpoints");
  if (s->lora_words * 64 > ((e->cfg.max_adapters + 63) / 64) * 64) return fail(e, EPPSCORE_ERR_CAPACITY, "lora_words*64 exceeds config.max_adapters");
  if (s->M > 0 && (!s->kv_usage || !s->queue)) return fail(e, EPPSCORE_ERR_INVALID, "kv_usage and queue are required");
  const size_t M = (size_t)s->M;
  e->M = s->M;
  e->words = s->lora_words > 0 ? s->lora_words : 1;
  e->kv.assign(s->kv_usage, s->kv_usage + M);
  e->queue.assign(s->queue, s->queue + M);
  e->have_running = s->running != nullptr;
  if (e->have_running) e->running.assign(s->running, s->running + M);
  else e->running.assign(M, 0);
  e->have_lora = s->lora_active && s->lora_waiting && s->lora_words > 0;
  e->act.assign(M * e->words, 0);
  e->wait.assign(M * e->words, 0);
  e->nm.assign(M, 0);
  e->mx.assign(M, 0);
  if (e->have_lora) {
    e->act.assign(s->lora_active, s->lora_active + M * e->words);
    e->wait.assign(s->lora_waiting, s->lora_waiting + M * e->words);
  }
  if (s->lora_nmodels) e->nm.assign(s->lora_nmodels, s->lora_nmodels + M);
  if (s->lora_max) e->mx.assign(s->lora_max, s->lora_max + M);
  for (int k = 0; k < EPPSCORE_MAX_ENDPOINT_COLS; k++) {
    e->have_col[k] = s->endpoint_col[k] != nullptr;
    if (e->have_col[k]) e->col[k].assign(s->endpoint_col[k], s->endpoint_col[k] + M);
  }
  e->have_min_tpot = s->min_tpot_slo != nullptr;
  if (e->have_min_tpot) e->min_tpot.assign(s->min_tpot_slo, s->min_tpot_slo + M);
  e->have_dispatched = s->dispatched != nullptr;
  if (e->have_dispatched) e->dispatched.assign(s->dispatched, s->dispatched + M);
  e->have_prefill = s->prefill_role != nullptr;
  if (e->have_prefill) e->prefill.assign(s->prefill_role, s->prefill_role + M);
  e->have_tokens = s->inflight_tokens != nullptr;
  if (e->have_tokens) e->tokens.assign(s->inflight_tokens, s->inflight_tokens + M);
  e->have_snapshot = true;
  return EPPSCORE_OK;
}

int32_t eppscore_schedule_batch(eppscore_engine* e, const eppscore_batch* b) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (!b || b->struct_size != sizeof(*b)) return fail(e, EPPSCORE_ERR_INVALID, "batch NULL or struct_size mismatch");
  if (b->location != 0) return fail(e, EPPSCORE_ERR_INVALID, "the stand-in takes host batches only");
  if (!e->have_snapshot) return fail(e, EPPSCORE_ERR_NO_SNAPSHOT, "schedule_batch before set_snapshot");
  if (b->R <= 0) return EPPSCORE_OK;
  if (!b->pick || !b->pick_score || !b->tie_count) return fail(e, EPPSCORE_ERR_INVALID, "pick/pick_score/tie_count required");
  orc_snapshot s;
  memset(&s, 0, sizeof(s));
  s.M = e->M;
  s.lora_words = e->words;
  s.kv_usage = e->kv.data();
  s.queue = e->queue.data();
  s.running = e->running.data();
  s.lora_active = e->act.data();
  s.lora_waiting = e->wait.data();
  s.lora_nmodels = e->nm.data();
  s.lora_max = e->mx.data();
  for (int k = 0; k < 4 && k < EPPSCORE_MAX_ENDPOINT_COLS; k++) s.endpoint_col[k] = e->have_col[k] ? e->col[k].data() : nullptr;
  s.min_tpot_slo = e->have_min_tpot ? e->min_tpot.data() : nullptr;
  s.dispatched = e->have_dispatched ? e->dispatched.data() : nullptr;
  s.prefill_role = e->have_prefill ? e->prefill.data() : nullptr;
  s.inflight_tokens = e->have_tokens ? e->tokens.data() : nullptr;
  orc_profile p = e->prof;
  p.latency = e->have_lat ? &e->lat : nullptr;
  for (int i = 0; i < p.n_scorers; i++)
    if (p.scorer_kind[i] == ORC_SCORER_LATENCY && !p.latency) return fail(e, EPPSCORE_ERR_INVALID, "latency scorer configured but no latency params were set");
  // the oracle takes contiguous prompts; the host layer pads every start to 16 bytes and passes explicit lengths
  std::vector<uint8_t> bytes;
  std::vector<int64_t> off;
  orc_batch ob;
  memset(&ob, 0, sizeof(ob));
  ob.R = b->R;
  ob.request_base = b->request_base;
  if (b->prompt_bytes && b->prompt_off) {
    off.assign((size_t)b->R + 1, 0);
    for (int r = 0; r < b->R; r++) {
      const int64_t len = b->prompt_len ? (int64_t)b->prompt_len[r] : b->prompt_off[r + 1] - b->prompt_off[r];
      off[r] = (int64_t)bytes.size();
      bytes.insert(bytes.end(), b->prompt_bytes + b->prompt_off[r], b->prompt_bytes + b->prompt_off[r] + len);
    }
    off[b->R] = (int64_t)bytes.size();
    bytes.push_back(0);
    ob.prompt_bytes = bytes.data();
    ob.prompt_off = off.data();
  }
  ob.model_seed = b->model_seed;
  ob.hashes_in = b->hashes_in;
  ob.n_hashes_in = b->n_hashes_in;
  ob.hash_stride = b->hash_stride;
  ob.adapter_id = b->adapter_id;
  ob.cand_mask = b->cand_mask;
  ob.dense_feat = b->dense_feat;
  ob.dense_total = b->dense_total;
  ob.block_chars = b->block_chars > 0 ? b->block_chars : e->cfg.block_chars;
  ob.max_blocks = b->max_blocks > 0 ? b->max_blocks : e->cfg.max_blocks;
  ob.pick = b->pick;
  ob.pick_score = b->pick_score;
  ob.tie_count = b->tie_count;
  ob.match_blocks = b->match_blocks;
  ob.total_blocks = b->total_blocks;
  ob.hashes_out = b->hashes_out;
  ob.weighted_out = b->scores_out;
  ob.input_tokens = b->input_tokens;
  ob.ttft_slo = b->ttft_slo;
  ob.tpot_slo = b->tpot_slo;
  ob.pred_out = b->pred_out;
  ob.filter_mask_out = b->filter_mask_out;
  if (orc_schedule_batch(&s, &p, e->idx, &ob, 1) != 0) return fail(e, EPPSCORE_ERR_INVALID, "oracle: schedule_batch failed");
  return EPPSCORE_OK;
}

int32_t eppscore_commit_picks(eppscore_engine* e, int32_t R, const int32_t* pick, const uint64_t* hashes, const uint16_t* n_hashes,
                              int32_t hash_stride, const int32_t* lru_capacity) {
  if (!e) return EPPSCORE_ERR_INVALID;
  if (R <= 0) return EPPSCORE_OK;
  if (!pick || !hashes || !n_hashes) return fail(e, EPPSCORE_ERR_INVALID, "pick/hashes/n_hashes required");
  orc_commit_picks(e->idx, R, pick, hashes, n_hashes, hash_stride, lru_capacity);
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_remove_endpoint(eppscore_engine* e, int32_t endpoint) {
  if (!e) return EPPSCORE_ERR_INVALID;
  orc_index_remove_pod(e->idx, endpoint);
  return EPPSCORE_OK;
}

int32_t eppscore_prefix_lru_len(const eppscore_engine* e, int32_t endpoint) { return e ? orc_index_lru_len(e->idx, endpoint) : -1; }

uint64_t eppscore_model_seed(const void* model, size_t model_len, const void* salt, size_t salt_len) {
  return orc_model_seed(model, model_len, salt, salt_len);
}

}  // extern "C"
