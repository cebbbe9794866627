/*
 * oracle.h — CPU restatement of the reference Endpoint-Picker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or as the
 * timed CPU baseline.  The product (libeppscore.so) never links or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * kubernetes-sigs/gateway-api-inference-extension @ c4c8fef).  The Go sources
 * cannot be built here (no Go toolchain), so this is a "port" oracle; it is
 * pinned against every golden vector the reference's own tests hold for this
 * path (tests/golden/reference_vectors.json, tests/test_oracle_golden.py) and
 * its XXH64 against two independent implementations (python-xxhash 3.7.0 and
 * the system libxxhash 0.8.2), because the reference pins no hash VALUE
 * (third-party github.com/cespare/xxhash/v2 v2.3.0, go.mod:6, not vendored).
 *
 * All scheduler arithmetic is IEEE float64, multiply then add, no FMA
 * contraction (build with -ffp-contract=off), accumulated in scorer order
 * from 0.0 exactly as pkg/epp/scheduling/scheduler_profile.go:155-168.
 */
#ifndef EPP_ORACLE_H
#define EPP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SCORERS 8

/* scorer kinds (same numbering as include/eppscore.h, restated independently) */
enum {
  ORC_SCORER_QUEUE = 0,      /* scorer/queuedepth/queue.go:78-108 */
  ORC_SCORER_KV_CACHE = 1,   /* scorer/kvcacheutilization/kvcache_utilization.go:76-82 */
  ORC_SCORER_PREFIX = 2,     /* scorer/prefix/plugin.go:95-117 */
  ORC_SCORER_LORA = 3,       /* scorer/loraaffinity/lora_affinity.go:76-102 */
  ORC_SCORER_RUNNING = 4,    /* scorer/runningrequests/runningrequest.go:78-108 (same form as queue) */
  ORC_SCORER_LATENCY = 5,    /* scorer/latency/plugin.go:144-318 fed by the predicted-latency producer */
  ORC_SCORER_TOKEN_LOAD = 6, /* scorer/tokenload/token_load.go:83-111 */
  ORC_SCORER_ENDPOINT_COL0 = 8,  /* +k: caller-supplied per-endpoint float64 score column k (0..3) */
  ORC_SCORER_PAIR_COL0 = 16      /* +k: caller-supplied per-(request,endpoint) float32 column k (0..1) */
};

enum { ORC_TIE_LOWEST_INDEX = 0, ORC_TIE_SEEDED_RANDOM = 1 };
/* pickers: max-score (picker/maxscore/picker.go:87-115), weighted-random A-Res (picker/weightedrandom/picker.go:111-155),
 * random (picker/random/picker.go:85-101).  The reference draws from a time-seeded process-wide PCG
 * (picker/common.go:40-55), so the stochastic pickers have no bit-parity definition; engine and oracle share a
 * documented counter-based generator instead and are checked at the distribution level. */
enum { ORC_PICK_MAX_SCORE = 0, ORC_PICK_WEIGHTED_RANDOM = 1, ORC_PICK_RANDOM = 2 };

/*
 * Latency-predictor fold-in: the Bayesian-ridge linear model the Go client evaluates from
 * cached coefficients (sidecars/latencypredictorasync/prediction.go:164-194), the producer's validity /
 * headroom rules (requestcontrol/dataproducer/predictedlatency/prediction.go:46-166) and the
 * latency-scorer's configuration (scheduling/scorer/latency/plugin.go:59-90).
 */
typedef struct orc_latency_params {
  double ttft_intercept, ttft_kv, ttft_input, ttft_waiting, ttft_running, ttft_prefix;
  double tpot_intercept, tpot_kv, tpot_input, tpot_waiting, tpot_running, tpot_generated;
  double slo_buffer_factor;  /* Config.SLOBufferFactor, default 1 (plugin.go:132) */
  int32_t streaming_mode;    /* Config.StreamingMode, default false (plugin.go:134) */
  int32_t has_predictions;   /* 0: no LatencyPredictionInfo on any endpoint (sidecar down) => composite fallback */
  double ttft_weight, tpot_weight; /* scorer Config, defaults 0.8 / 0.2 */
  int32_t strategy_most;     /* HeadroomSelectionStrategy: 0 "least" (default), 1 "most" */
  int32_t reserved;
  double composite_kv, composite_queue, composite_prefix; /* defaults 1,1,1 */
} orc_latency_params;

/* device-side Filter plugins of the latency profile (config/charts/epplib/templates/_config.yaml:47-75):
 *   prefix-cache-affinity-filter  filter/prefixcacheaffinity/plugin.go:105-151  params {affinityThreshold, explorationProbability, maxTTFTPenaltyMs}
 *   slo-headroom-tier-filter      filter/sloheadroomtier/plugin.go:82-137       params {epsilonExploreNeg}
 * Their rand.Float64() draws come from the same counter-based generator as the pickers: draw of filter f for request r =
 * orc_uniform01(tie_seed, r, -(f+1)). */
enum { ORC_FILTER_PREFIX_AFFINITY = 1, ORC_FILTER_SLO_HEADROOM_TIER = 2 };
#define ORC_MAX_FILTERS 4

typedef struct orc_profile {
  int32_t n_scorers;
  int32_t scorer_kind[ORC_MAX_SCORERS];
  double scorer_weight[ORC_MAX_SCORERS];
  int32_t tie_mode;
  uint64_t tie_seed;
  const orc_latency_params *latency; /* required by ORC_SCORER_LATENCY */
  double token_load_threshold;       /* queueThresholdTokens (token_load.go:33,57-61); <= 0 => 4194304 */
  int32_t pick_mode;                 /* ORC_PICK_* */
  int32_t n_filters;                 /* run in order before the scorers, on top of the caller's candidate mask */
  int32_t filter_kind[ORC_MAX_FILTERS];
  double filter_param[ORC_MAX_FILTERS][3];
} orc_profile;

/* One immutable metrics snapshot (interface/datalayer/metrics.go:26-42 fields the path reads). */
typedef struct orc_snapshot {
  int32_t M;
  int32_t lora_words;            /* ceil(A/64) */
  const double *kv_usage;        /* KVCacheUsagePercent */
  const int64_t *queue;          /* WaitingQueueSize */
  const int64_t *running;        /* RunningRequestsSize */
  const uint64_t *lora_active;   /* M x lora_words: bit a = adapter a in ActiveModels */
  const uint64_t *lora_waiting;  /* M x lora_words */
  const int32_t *lora_nmodels;   /* len(ActiveModels)+len(WaitingModels) */
  const int32_t *lora_max;       /* MaxActiveModels */
  const double *endpoint_col[4]; /* optional generic per-endpoint score columns */
  /* predicted-latency producer state per endpoint (plugin.go:347-363) and role label; NULL => 0 */
  const double *min_tpot_slo;    /* getEndpointMinTPOTSLO */
  const int32_t *dispatched;     /* getEndpointRunningRequestCount */
  const uint8_t *prefill_role;   /* hasPrefillRole(EndpointRoleLabel, endpoint) (prediction.go:168-175) */
  const int64_t *inflight_tokens;/* InFlightLoad.Tokens attribute (token_load.go:91-95); NULL => attribute absent */
} orc_snapshot;

/* ---- XXH64 (third-party cespare/xxhash v2.3.0 == canonical XXH64, seed 0) ---- */
uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed);
/* hashing.go:70-77: XXH64(model || cacheSalt) */
uint64_t orc_model_seed(const void *model, size_t model_len, const void *salt, size_t salt_len);
/* hashing.go:34-98; returns number of hashes written (<= out_cap; -1 if out_cap too small) */
int32_t orc_hash_prompt(const uint8_t *input, int64_t len, uint64_t model_seed, int32_t block_chars,
                        int32_t max_blocks, uint64_t *out, int32_t out_cap);

/* ---- prefix index: approximateprefix/indexer.go ---- */
typedef struct orc_index orc_index;
orc_index *orc_index_new(int32_t default_lru_size);                 /* indexer.go:40-49 */
void orc_index_free(orc_index *);
void orc_index_add(orc_index *, const uint64_t *hashes, int32_t n, int32_t server,
                   int32_t num_gpu_blocks);                        /* indexer.go:52-83 */
int32_t orc_index_get(const orc_index *, uint64_t hash, int32_t *servers_out, int32_t cap); /* :86-102 */
void orc_index_remove_pod(orc_index *, int32_t server);            /* :167-182 */
int32_t orc_index_lru_len(const orc_index *, int32_t server);      /* -1 if the pod has no LRU */
int32_t orc_index_lru_keys(const orc_index *, int32_t server, uint64_t *out, int32_t cap); /* oldest→newest */
int64_t orc_index_num_hashes(const orc_index *);                   /* len(hashToPods) */
int32_t orc_index_pods(const orc_index *, int32_t *out, int32_t cap); /* :185-195 */
/* dump every (hash, server) pair currently in hashToPods (for seeding the device table) */
int64_t orc_index_dump(const orc_index *, uint64_t *hash_out, int32_t *server_out, int64_t cap);

/* approximateprefix/plugin.go:219-235: counts per server; servers >= M are ignored */
void orc_match_longest_prefix(const orc_index *, const uint64_t *hashes, int32_t n, int32_t M,
                              uint16_t *match_out /*M, zero-filled here*/);

/* ---- single-scorer bodies (for the per-scorer golden tests). cand_mask NULL = all M ---- */
void orc_score_kv(const orc_snapshot *, const uint32_t *cand_mask, double *out);
void orc_score_queue(const orc_snapshot *, const uint32_t *cand_mask, double *out);
void orc_score_running(const orc_snapshot *, const uint32_t *cand_mask, double *out);
void orc_score_lora(const orc_snapshot *, const uint32_t *cand_mask, int32_t adapter_id, double *out);
void orc_score_prefix(int32_t M, const uint32_t *cand_mask, const uint16_t *match, int32_t total,
                      int32_t have_info, double *out);
double orc_enforce_score_range(double s); /* scheduler_profile.go:194-202 */

void orc_score_token_load(const orc_snapshot *, const uint32_t *cand_mask, double threshold, double *out);

/* len(strings.Fields(s)) (predictedlatency/plugin.go:286): Go's utf8.DecodeRuneInString + unicode.IsSpace restated */
int32_t orc_count_fields(const uint8_t *s, int64_t len);

/* per-request inputs of the latency path */
typedef struct orc_latency_request {
  int64_t input_tokens; /* len(strings.Fields(prompt)), training.go:51 */
  double ttft_slo;      /* x-slo-ttft-ms header or 0 (plugin.go:330-343) */
  double tpot_slo;      /* x-slo-tpot-ms header or 0 */
} orc_latency_request;
/* prediction.go:164-194 with NumTokensGenerated = generated */
void orc_latency_predict(const orc_latency_params *, double kv, int64_t input_tokens, int64_t waiting,
                         int64_t running, int64_t generated, double prefix_score, double *ttft, double *tpot);
/* predictedlatency/prediction.go:137-166 (+ the neutralisation of :100-104 when neutralize != 0);
 * out = {ttftOk, tpotOk, isValid}, headrooms = {tpot headroom, ttft headroom} */
void orc_latency_validate(const orc_latency_params *, double ttft, double tpot, double ttft_slo, double tpot_slo,
                          double pod_min_tpot_slo, int32_t neutralize, int32_t *ok_out, double *headroom_out);
/* scorer/latency/plugin.go:144-318 over caller-supplied LatencyPredictionInfo (for the scorer's own tests):
 * have_info[m]==0 => attribute absent. Non-candidates are left untouched. */
void orc_score_latency_info(const orc_latency_params *, const orc_snapshot *, const uint32_t *cand_mask,
                            const uint8_t *have_info, const double *ttft_headroom, const double *tpot_headroom,
                            const int32_t *dispatched, const uint16_t *match, int32_t total, double *out);
/* producer + scorer for one request: PrepareRequestData (preparedata_hooks.go:36-104) over all M endpoints,
 * then Score over the candidates. pred_out: optional M x 2 {ttft, tpot}. */
void orc_score_latency(const orc_latency_params *, const orc_snapshot *, const uint32_t *cand_mask,
                       const uint16_t *match, int32_t total, const orc_latency_request *, double *out,
                       double *pred_out);

/* runFilterPlugins (scheduler_profile.go:130-149) for the device-side filters: mask_in NULL = all M;
 * mask_out receives ceil(M/32) words.  Predictions are made like PrepareRequestData (all endpoints). */
void orc_apply_filters(const orc_snapshot *, const orc_profile *, int64_t request_index, const uint32_t *mask_in,
                       const uint16_t *match, int32_t total, const orc_latency_request *lat, uint32_t *mask_out);

/* counter-based U in (0,1] and -ln(U) built from +,-,*,/ only (bit-reproducible on any IEEE machine) */
double orc_uniform01(uint64_t seed, int64_t request_index, int32_t endpoint);
double orc_neg_log(double u);

/* counter-based tie priority shared (by specification) with the CUDA engine */
uint32_t orc_tie_priority(uint64_t seed, int64_t request_index, int32_t endpoint);

/*
 * SchedulerProfile.Run for ONE request (scheduler_profile.go:117-192 + maxscore/picker.go:87-115
 * restated as "arg-max set"): returns 0, or -1 when the candidate set is empty
 * ("no endpoints available for the given request").
 *   match/total      : PrefixCacheMatchInfo per endpoint (NULL ⇒ attribute absent ⇒ prefix score 0)
 *   pair_col         : optional M x 2 float32 per-pair columns (NULL ok)
 *   weighted_out     : optional M doubles (weighted score per endpoint; NaN for non-candidates)
 *   tie_set_out      : optional ceil(M/32) words, bit m = endpoint m attains the max
 */
int32_t orc_schedule_one(const orc_snapshot *, const orc_profile *, int64_t request_index,
                         int32_t adapter_id, const uint32_t *cand_mask, const uint16_t *match,
                         int32_t total, const float *pair_col, int32_t *pick_out, double *score_out,
                         int32_t *tie_count_out, uint32_t *tie_set_out, double *weighted_out);
/* the same with the latency path's per-request inputs (NULL => zeros) and optional M x 2 prediction output */
int32_t orc_schedule_one_lat(const orc_snapshot *, const orc_profile *, int64_t request_index,
                             int32_t adapter_id, const uint32_t *cand_mask, const uint16_t *match,
                             int32_t total, const float *pair_col, const orc_latency_request *lat,
                             int32_t *pick_out, double *score_out, int32_t *tie_count_out,
                             uint32_t *tie_set_out, double *weighted_out, double *pred_out);

/* Batch description; mirrors eppscore_batch (host pointers only). */
typedef struct orc_batch {
  int32_t R;
  int64_t request_base;
  const uint8_t *prompt_bytes;   /* NULL ⇒ use hashes_in (or no prefix info) */
  const int64_t *prompt_off;     /* R+1 */
  const uint64_t *model_seed;    /* R */
  const uint64_t *hashes_in;     /* R x hash_stride, optional */
  const uint16_t *n_hashes_in;   /* R */
  int32_t hash_stride;
  const int32_t *adapter_id;     /* R or NULL (=-1) */
  const uint32_t *cand_mask;     /* R x ceil(M/32) or NULL */
  const float *dense_feat;       /* R x M x 4 {match, lora_class, pair_col0, pair_col1} or NULL */
  const uint16_t *dense_total;   /* R, with dense_feat */
  int32_t block_chars, max_blocks;
  /* outputs */
  int32_t *pick;                 /* R */
  double *pick_score;            /* R */
  int32_t *tie_count;            /* R */
  uint32_t *tie_set;             /* optional R x ceil(M/32) */
  uint16_t *match_blocks;        /* optional R x M */
  uint16_t *total_blocks;        /* optional R */
  uint64_t *hashes_out;          /* optional R x max_blocks */
  double *weighted_out;          /* optional R x M: weightedScorePerEndpoint, NaN for non-candidates */
  /* latency path (all optional) */
  const int32_t *input_tokens;   /* R */
  const double *ttft_slo;        /* R */
  const double *tpot_slo;        /* R */
  double *pred_out;              /* R x M x 2 {ttft, tpot} */
  uint32_t *filter_mask_out;     /* optional R x ceil(M/32): the candidate set after the filter chain */
} orc_batch;

/* Whole hot path for a batch (hash → match → score → pick), requests partitioned over n_threads
 * like one goroutine per request would be. idx may be NULL (no prefix index). Returns 0. */
int32_t orc_schedule_batch(const orc_snapshot *, const orc_profile *, const orc_index *idx,
                           const orc_batch *, int32_t n_threads);

/* PreRequest for a batch, in request order (plugin.go:169-197): Add(hashes[r], pick[r]). */
void orc_commit_picks(orc_index *, int32_t R, const int32_t *pick, const uint64_t *hashes,
                      const uint16_t *n_hashes, int32_t hash_stride, const int32_t *gpu_blocks /*M or NULL*/);

#ifdef __cplusplus
}
#endif
#endif
