/*
 * eppscore.h — C ABI of the B200-native batched Endpoint-Picker scoring engine (libeppscore.so).
 *
 * This is the drop-in boundary for ONE hot path of kubernetes-sigs/gateway-api-inference-extension
 * (reference @ c4c8fef; paths below are relative to it): the per-request Filter → Score → Pick
 * loop behind Scheduler.Schedule() plus the approximate-prefix producer that feeds it, executed
 * for a BATCH of R requests against M endpoints by hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI today (pure Go, CGO_ENABLED=0, Dockerfile:8); these entry points are
 * exactly what a cgo shim behind requestcontrol.Scheduler (pkg/epp/requestcontrol/director.go:68-70)
 * would bind.  INTEGRATION.md shows that shim.  Conventions:
 *   - plain C types, caller-owned contiguous little-endian arrays, no callbacks, no torch types;
 *   - every function returns int32 status (EPPSCORE_OK == 0, negative = error) and
 *     eppscore_last_error() gives the message; handles are opaque;
 *   - one engine per CUDA device; calls on one engine are serialised by the caller
 *     (the Go shim holds a mutex), different engines are independent;
 *   - there is NO CPU fallback: eppscore_create fails (EPPSCORE_ERR_NO_DEVICE) without a GPU.
 *
 * Numerics contract: all scheduler arithmetic is IEEE float64, multiply then add
 * without FMA contraction, accumulated in scorer order from 0.0 — bit-identical to
 * pkg/epp/scheduling/scheduler_profile.go:155-168 on GOARCH=amd64.  Block hashes are XXH64
 * (seed 0) chained as approximateprefix/hashing.go:70-94.  Picks are the arg-max; ties are
 * reported (tie_count) and broken deterministically (see eppscore_tie_mode) where the reference
 * breaks them with a time-seeded shuffle (picker/common.go:49-55, maxscore/picker.go:91-102).
 */
#ifndef EPPSCORE_H
#define EPPSCORE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPPSCORE_ABI_VERSION 3
#define EPPSCORE_MAX_SCORERS 8
#define EPPSCORE_MAX_ENDPOINT_COLS 4
#define EPPSCORE_MAX_BLOCKS 65535 /* match/total are uint16 (attribute/prefix/data_types.go:27-34 are Go ints) */

typedef struct eppscore_engine eppscore_engine; /* opaque */

typedef enum eppscore_status {
  EPPSCORE_OK = 0,
  EPPSCORE_ERR_INVALID = -1,     /* bad argument / config */
  EPPSCORE_ERR_CUDA = -2,        /* CUDA runtime error (message in last_error) */
  EPPSCORE_ERR_CAPACITY = -3,    /* M / A / table capacity exceeded */
  EPPSCORE_ERR_NO_SNAPSHOT = -4, /* schedule called before set_snapshot */
  EPPSCORE_ERR_NO_DEVICE = -5    /* no CUDA device: the engine has no CPU path */
} eppscore_status;

/* Scorer kinds, in the sense of framework.Scorer (interface/scheduling/plugins.go:68-72). */
typedef enum eppscore_scorer_kind {
  EPPSCORE_SCORER_QUEUE = 0,    /* queue-scorer: scorer/queuedepth/queue.go:78-108 */
  EPPSCORE_SCORER_KV_CACHE = 1, /* kv-cache-utilization-scorer: scorer/kvcacheutilization/kvcache_utilization.go:76-82 */
  EPPSCORE_SCORER_PREFIX = 2,   /* prefix-cache-scorer: scorer/prefix/plugin.go:95-117 (+ producer approximateprefix/) */
  EPPSCORE_SCORER_LORA = 3,     /* lora-affinity-scorer: scorer/loraaffinity/lora_affinity.go:76-102 */
  EPPSCORE_SCORER_RUNNING = 4,  /* running-requests-size-scorer: scorer/runningrequests/runningrequest.go:78-108 */
  /* latency-scorer (scorer/latency/plugin.go:144-318) with the predicted-latency producer folded in: per
   * (request, endpoint) the Bayesian-ridge TTFT/TPOT prediction (sidecars/latencypredictorasync/prediction.go:164-194),
   * headroom/validity (predictedlatency/prediction.go:137-166), then the scorer's tier/bucket logic over the
   * candidate set.  Parameters: eppscore_set_latency_params. */
  EPPSCORE_SCORER_LATENCY = 5,
  EPPSCORE_SCORER_TOKEN_LOAD = 6, /* token-load-scorer: scorer/tokenload/token_load.go:83-111 (snapshot.inflight_tokens) */
  /* 8+k: a host-computed, request-independent scorer supplied as float64 column k of the snapshot
   * (any custom framework.Scorer whose output does not depend on the request); clamped+weighted in-kernel. */
  EPPSCORE_SCORER_ENDPOINT_COL0 = 8,
  /* 16+k: a per-(request,endpoint) float32 column k∈{0,1} of the dense feature rows (any per-pair scorer
   * computed elsewhere, e.g. a tree-ensemble latency model's output). */
  EPPSCORE_SCORER_PAIR_COL0 = 16
} eppscore_scorer_kind;

typedef enum eppscore_tie_mode {
  EPPSCORE_TIE_LOWEST_INDEX = 0, /* deterministic: lowest endpoint index in the arg-max set */
  /* counter-based stand-in for the reference's shuffle: the arg-max member with the largest
   * lowbias32-mixed priority of (tie_seed, request_base + r, endpoint) wins — uniform over the
   * tie set, reproducible, order-independent (so every GPU shard agrees). */
  EPPSCORE_TIE_SEEDED_RANDOM = 1
} eppscore_tie_mode;

/* Pickers (framework.Picker, interface/scheduling/plugins.go:74-78).  The reference's stochastic pickers draw from a
 * process-wide PCG seeded with the wall clock (picker/common.go:40-55), so they have no bit-parity definition; the
 * engine uses a counter-based generator keyed by (tie_seed, request_base + r, endpoint) — reproducible, identical on
 * every shard — and is checked at the distribution level (the reference's own picker tests do the same, ±5 %). */
typedef enum eppscore_pick_mode {
  EPPSCORE_PICK_MAX_SCORE = 0,       /* max-score-picker: picker/maxscore/picker.go:87-115 */
  /* weighted-random-picker (A-Res, picker/weightedrandom/picker.go:111-155): P(m) = score[m] / Σ score over the
   * candidates with score > 0; key U^(1/score) maximal ⇔ -ln(U)/score minimal; no positive score ⇒ random picker.
   * tie_count reports the size of the set the draw was over. */
  EPPSCORE_PICK_WEIGHTED_RANDOM = 1,
  EPPSCORE_PICK_RANDOM = 2           /* random-picker: picker/random/picker.go:85-101 — uniform over the candidates */
} eppscore_pick_mode;

/* Filter plugins evaluated on the device, in order, before the scorers and on top of batch.cand_mask (host-side Filters
 * of any other kind keep reaching the engine as that mask).  Both need the latency scorer in the profile (they read the same
 * per-pair predictions) and cannot be combined with the queue / running scorers.  Their rand.Float64() draws come from the
 * counter-based generator of eppscore_pick_mode: draw of filter f for request r is keyed by (tie_seed, request_base + r, -(f+1)). */
typedef enum eppscore_filter_kind {
  /* prefix-cache-affinity-filter (filter/prefixcacheaffinity/plugin.go:105-151): params {affinityThreshold,
   * explorationProbability, maxTTFTPenaltyMs}; reference defaults 0.80, 0.01, 5000 */
  EPPSCORE_FILTER_PREFIX_AFFINITY = 1,
  /* slo-headroom-tier-filter (filter/sloheadroomtier/plugin.go:82-137): params {epsilonExploreNeg}; default 0.01 */
  EPPSCORE_FILTER_SLO_HEADROOM_TIER = 2
} eppscore_filter_kind;
#define EPPSCORE_MAX_FILTERS 4

/* Scheduler profile + engine sizing.  Replaces SchedulerProfile{scorers, picker}
 * (pkg/epp/scheduling/scheduler_profile.go:41-98) and the approximateprefix config
 * (approximateprefix/types.go:77-141). Zero-initialise, set struct_size = sizeof, fill. */
typedef struct eppscore_config {
  uint32_t struct_size;
  int32_t n_scorers;                              /* profile order matters: float64 adds are not associative */
  int32_t scorer_kind[EPPSCORE_MAX_SCORERS];
  double scorer_weight[EPPSCORE_MAX_SCORERS];     /* WeightedScorer.Weight(), weighted_scorer.go:32 */
  int32_t block_chars;     /* blockSizeTokens*averageCharactersPerToken; default 16*4 (types.go:91,112) */
  int32_t max_blocks;      /* defaultMaxPrefixBlocks = 256 (types.go:98) */
  int32_t tie_mode;        /* eppscore_tie_mode */
  uint64_t tie_seed;
  int32_t max_endpoints;   /* capacity for M (rounded up internally); default 1024 */
  int32_t max_adapters;    /* capacity for the LoRA adapter dictionary A; default 64 */
  int64_t prefix_capacity; /* INITIAL capacity (distinct block hashes) of the device table; it is rebuilt (emptied keys
                              dropped) or doubled on demand. default 1<<18 */
  int32_t lru_capacity_default; /* defaultLRUCapacityPerServer = 31250 (types.go:109) */
  int32_t lru_capacity_max;     /* largest per-endpoint LRU size (CacheNumBlocks, plugin.go:207-216) the engine must support;
                                   0 ⇒ max(lru_capacity_default, the capacities named by the first Add).  It fixes the size of the
                                   per-endpoint device regions; a later Add that names a larger size fails with ERR_CAPACITY. */
  double token_load_threshold;  /* token-load-scorer queueThresholdTokens; <= 0 ⇒ 4194304 (token_load.go:33,57-61) */
  int32_t pick_mode;            /* eppscore_pick_mode; default max-score */
  int32_t n_filters;            /* device-side filters, see eppscore_filter_kind */
  int32_t filter_kind[EPPSCORE_MAX_FILTERS];
  double filter_param[EPPSCORE_MAX_FILTERS][3];
} eppscore_config;

/* Latency fold-in parameters: the cached Bayesian-ridge coefficients (MetricsResponse.Coefficients,
 * sidecars/latencypredictorasync/prediction.go:164-194), the producer's Config (predictedlatency/plugin.go:118-136)
 * and the latency-scorer's Config (scorer/latency/plugin.go:59-90).  Takes effect at the next eppscore_set_snapshot
 * (coefficients are refreshed on the metrics cadence, like the snapshot). */
typedef struct eppscore_latency_params {
  uint32_t struct_size;
  int32_t has_predictions;   /* 0 = no LatencyPredictionInfo (sidecar down / timed out) ⇒ composite fallback (plugin.go:169-172) */
  double ttft_intercept, ttft_kv, ttft_input, ttft_waiting, ttft_running, ttft_prefix;
  double tpot_intercept, tpot_kv, tpot_input, tpot_waiting, tpot_running, tpot_generated;
  double slo_buffer_factor;  /* SLOBufferFactor, default 1 */
  int32_t streaming_mode;    /* StreamingMode, default 0: TPOT neutralised (prediction.go:100-104) */
  int32_t strategy_most;     /* HeadroomSelectionStrategy: 0 "least" (default), 1 "most" */
  double ttft_weight, tpot_weight;                         /* defaults 0.8, 0.2 */
  double composite_kv, composite_queue, composite_prefix;  /* defaults 1, 1, 1 */
} eppscore_latency_params;

/* Immutable metrics snapshot: the fields of fwkdl.Metrics the path reads
 * (interface/datalayer/metrics.go:26-42), packed SoA.  LoRA maps become dictionary bitmasks:
 * bit a of lora_active[m*lora_words + a/64] ⇔ adapter a ∈ ActiveModels of endpoint m. */
typedef struct eppscore_snapshot {
  uint32_t struct_size;
  int32_t M;
  int32_t lora_words;            /* ceil(A/64) */
  int32_t location;              /* 0 = host pointers, 1 = device pointers (e.g. an NCCL-broadcast buffer) */
  const double *kv_usage;        /* KVCacheUsagePercent            [M] */
  const int64_t *queue;          /* WaitingQueueSize               [M] */
  const int64_t *running;        /* RunningRequestsSize            [M] (may be NULL) */
  const uint64_t *lora_active;   /* [M*lora_words] (may be NULL ⇒ empty) */
  const uint64_t *lora_waiting;  /* [M*lora_words] */
  const int32_t *lora_nmodels;   /* len(ActiveModels)+len(WaitingModels) as MAP sizes (lora_affinity.go:90) [M] */
  const int32_t *lora_max;       /* MaxActiveModels                [M] */
  const double *endpoint_col[EPPSCORE_MAX_ENDPOINT_COLS]; /* optional generic score columns [M] */
  uint64_t epoch;                /* caller's snapshot generation, echoed by eppscore_stats */
  void *stream;                  /* cudaStream_t for location==1 (NULL = engine stream) */
  /* predicted-latency producer state per endpoint (all optional, NULL ⇒ 0) */
  const double *min_tpot_slo;    /* getEndpointMinTPOTSLO (predictedlatency/plugin.go:347-355)        [M] */
  const int32_t *dispatched;     /* getEndpointRunningRequestCount (plugin.go:357-363)                [M] */
  const uint8_t *prefill_role;   /* hasPrefillRole(EndpointRoleLabel, endpoint) (prediction.go:168-175) [M] */
  const int64_t *inflight_tokens;/* InFlightLoad.Tokens attribute (token_load.go:91-95); NULL ⇒ absent [M] */
} eppscore_snapshot;

/* One batch = R concurrent Scheduler.Schedule() calls (pkg/epp/scheduling/scheduler.go:54). */
typedef struct eppscore_batch {
  uint32_t struct_size;
  int32_t R;
  int32_t location;            /* 0 = host pointers (copies happen inside the call), 1 = device pointers */
  int32_t reserved0;
  int64_t request_base;        /* global index of request 0 (multi-GPU shards; feeds the tie priority) */
  /* --- prefix producer inputs: EITHER prompts (hashed on the GPU) OR precomputed hashes --- */
  const uint8_t *prompt_bytes; /* flattened getUserInputBytes() output (hashing.go:106-135), concatenated */
  const int64_t *prompt_off;   /* [R+1] byte offsets into prompt_bytes */
  const int32_t *prompt_len;   /* optional [R]: explicit lengths, so a host may pad every prompt START to 16 bytes
                                  (fast aligned hash path); NULL ⇒ len[r] = prompt_off[r+1]-prompt_off[r] */
  const uint64_t *model_seed;  /* [R] XXH64(TargetModel || cacheSalt) (hashing.go:70-77); see eppscore_model_seed */
  const uint64_t *hashes_in;   /* [R*hash_stride] optional: block hashes computed elsewhere */
  const uint16_t *n_hashes_in; /* [R] */
  int32_t hash_stride;
  int32_t block_chars;         /* 0 ⇒ config default; per call because autotune reads endpoints[0] (plugin.go:238-250) */
  int32_t max_blocks;          /* 0 ⇒ config default */
  int32_t reserved1;
  /* --- per-request scorer inputs --- */
  const int32_t *adapter_id;   /* [R] dictionary id of TargetModel, -1 = not a known adapter; NULL ⇒ all -1 */
  const uint32_t *cand_mask;   /* [R*ceil(M/32)] filter-chain result (scheduler_profile.go:130-149); NULL ⇒ all M */
  /* --- dense feature rows (optional; replaces the in-kernel prefix match and LoRA lookup) ---
   * float4 per (r,m): {x = matchBlocks (integer valued), y = lora class 0..3 ↦ {0,0.6,0.8,1.0},
   *                    z = pair column 0, w = pair column 1}                                        */
  const float *dense_feat;     /* [R*M*4] */
  const uint16_t *dense_total; /* [R] totalBlocks for the dense rows */
  /* --- outputs --- */
  int32_t *pick;               /* [R] chosen endpoint, -1 = "no endpoints available" (scheduler_profile.go:119-121) */
  double *pick_score;          /* [R] weighted score of the pick (ScoredEndpoint.Score, types.go:152-155) */
  int32_t *tie_count;          /* [R] size of the arg-max set */
  uint16_t *match_blocks;      /* optional [R*M]: PrefixCacheMatchInfo.matchBlocks per endpoint */
  uint16_t *total_blocks;      /* optional [R]:   PrefixCacheMatchInfo.totalBlocks */
  uint64_t *hashes_out;        /* optional [R*max_blocks]: block hashes (state for eppscore_commit_picks / PreRequest) */
  double *scores_out;          /* optional [R*M]: the whole weightedScorePerEndpoint map (scheduler_profile.go:155-174),
                                  NaN for non-candidates — diagnostics / parity tests; 8*R*M bytes of HBM writes */
  void *stream;                /* cudaStream_t for location==1 (NULL = engine stream); the call is async on it.  Index
                                * mutations and eppscore_set_snapshot wait for the last device-location batch; the engine's
                                * hash scratch is shared, so keep device-location batches that do not pass hashes_out on ONE
                                * stream at a time (or give each its own hashes_out) */
  /* latency fold-in, per request (optional, NULL ⇒ 0) */
  const int32_t *input_tokens; /* [R] len(strings.Fields(prompt)) (predictedlatency/training.go:51); NULL with prompt_bytes given ⇒
                                * counted on the device from those bytes (hosts whose PromptText() differs from the hashed
                                * bytes, e.g. chat completions, pass it explicitly) */
  const double *ttft_slo;      /* [R] x-slo-ttft-ms header value or 0 (predictedlatency/plugin.go:330-343) */
  const double *tpot_slo;      /* [R] x-slo-tpot-ms header value or 0 */
  double *pred_out;            /* optional [R*M*2]: predicted {TTFT, TPOT} per (request, endpoint) — diagnostics */
  uint32_t *filter_mask_out;   /* optional [R*ceil(M/32)]: the candidate set left by the device-side filters — diagnostics */
} eppscore_batch;

typedef struct eppscore_stats {
  uint32_t struct_size;
  int32_t M;
  uint64_t epoch;
  uint64_t kernel_launches;     /* CUDA kernels launched by this engine since creation */
  int64_t prefix_hashes;        /* table slots in use, emptied keys included (they are dropped at the next rebuild) */
  int64_t prefix_live_hashes;   /* hashes with a non-empty endpoint set == len(hashToPods) of the reference */
  int64_t prefix_capacity;      /* hashes the table holds at load 0.5 before it is rebuilt / doubled */
  int64_t prefix_table_bytes;   /* device bytes: 32-byte slots + overflow bitset rows */
  int64_t lru_entries;          /* Σ per-endpoint LRU lengths (prefix_indexer_size metric, metrics.go:349) */
  int64_t lru_bytes;            /* device bytes of the per-endpoint LRU regions (maps + logs) */
  int64_t prefix_overflow_rows; /* sets of more than 10 endpoints (bitset rows in use) */
  int64_t prefix_rebuilds;      /* table rebuilds so far */
  uint32_t index_error;         /* sticky device-side error flags of the index (0 = healthy) */
  uint32_t reserved;
} eppscore_stats;

/* ---- lifecycle ---- */
int32_t eppscore_abi_version(void);
void eppscore_config_default(eppscore_config *cfg); /* default-config weights queue 2, kv 2, prefix 3 (loader/defaults.go:46-103) */
int32_t eppscore_create(int32_t device, const eppscore_config *cfg, struct eppscore_engine **out);
void eppscore_destroy(struct eppscore_engine *e);
const char *eppscore_last_error(const struct eppscore_engine *e); /* e may be NULL: last create() error */
int32_t eppscore_get_stats(const struct eppscore_engine *e, eppscore_stats *out);
/* Diagnostics knobs (profiling / A-B runs only; never needed for correct operation):
 *   key 1: 1 = always use the fully general kernels (same as env EPPSCORE_FORCE_GENERIC=1), 0 = normal dispatch;
 *   key 2: hash stage mask (default 19 = the two-kernel form: bit 0 run the body kernel, bit 1 run the chain kernel, bit 4
 *          the CTA-tile chain kernel; without bit 4 the warp-tile chain kernel, 2 us slower at 64K requests).  Experimental single-kernel forms, both
 *          measured SLOWER than the two kernels on B200 (a warp in its chain phase has no loads in flight): bit 3 the
 *          warp-tile fused kernel, bit 2 the CTA-tile warp-specialised fused kernel; bit 5 the persistent software-pipelined
 *          body kernel (32.9 us against 28.3 us);
 *   key 3: requests per chunk of a host-location batch (0: never chunk);
 *   key 4: 0 = ordinary launches, 1 (default) = programmatic dependent launch between the kernels of a batch;
 *   key 5: slices a device-location batch is cut into (default 1), key 6: streams the slices alternate over (default 2;
 *          the caller's stream plus internal ones, joined before the call returns — also inside a CUDA-graph capture).
 *          Measured on B200 at 64K x 1024: 2 slices / 2 streams 77.0 us against 77.9 us unsliced, 4 or more slices slower
 *          (each slice's chain and probe kernels are latency bound and pay their full latency per slice). */
int32_t eppscore_set_debug(struct eppscore_engine *e, int32_t key, int64_t value);

/* ---- snapshot (replaces the per-request deep clone, director.go:342-349) ---- */
int32_t eppscore_set_snapshot(struct eppscore_engine *e, const eppscore_snapshot *s);

/* Latency fold-in parameters (see eppscore_latency_params); applied by the next eppscore_set_snapshot. */
void eppscore_latency_params_default(eppscore_latency_params *p);
int32_t eppscore_set_latency_params(struct eppscore_engine *e, const eppscore_latency_params *p);

/* ---- the hot path: Filter(mask) → Score → Pick for R requests ---- */
int32_t eppscore_schedule_batch(struct eppscore_engine *e, const eppscore_batch *b);

/* ---- stand-alone stages (same kernels; for parity tests and hosts that only want one stage) ---- */
/* hashPrompt for R prompts (hashing.go:34-98). hashes_out [R*max_blocks], n_hashes_out [R]. */
int32_t eppscore_hash_prompts(struct eppscore_engine *e, int32_t R, int32_t location, const uint8_t *prompt_bytes,
                              const int64_t *prompt_off, const int32_t *prompt_len /*optional*/,
                              const uint64_t *model_seed, int32_t block_chars, int32_t max_blocks,
                              uint64_t *hashes_out, uint16_t *n_hashes_out, void *stream);
/* len(strings.Fields(prompt)) for R prompts — the latency path's input_token_length
 * (predictedlatency/plugin.go:286, training.go:51); Go's rune decoding and unicode.IsSpace. out [R]. */
int32_t eppscore_count_fields(struct eppscore_engine *e, int32_t R, int32_t location, const uint8_t *prompt_bytes,
                              const int64_t *prompt_off, const int32_t *prompt_len /*optional*/, int32_t *out,
                              void *stream);
/* hashPrompt on the HOST cores (a persistent worker pool inside the library; n_threads <= 0: all of them) for callers that
 * ship hashes_in instead of prompt bytes: 8 bytes per 64-byte block cross PCIe instead of 64.  hashes_out [R*hash_stride]
 * (hash_stride >= max_blocks; entries past n_hashes_out[r] are zero).  Needs no engine and no GPU. */
int32_t eppscore_hash_prompts_host(int32_t R, const uint8_t *prompt_bytes, const int64_t *prompt_off,
                                   const int32_t *prompt_len /*optional*/, const uint64_t *model_seed, int32_t block_chars,
                                   int32_t max_blocks, uint64_t *hashes_out, int32_t hash_stride, uint16_t *n_hashes_out,
                                   int32_t n_threads);
/* host helper: XXH64(model || salt), hashing.go:70-77 */
uint64_t eppscore_model_seed(const void *model, size_t model_len, const void *salt, size_t salt_len);
uint64_t eppscore_xxh64(const void *data, size_t len, uint64_t seed);

/* ---- prefix index (approximateprefix/indexer.go) ----
 * Both halves of the reference's indexer are DEVICE-RESIDENT and maintained by kernels: hashToPods as an open-addressing
 * table of 32-byte slots (sets of up to 10 endpoints inline, bitset rows beyond), podToLRU as one log-structured exact LRU
 * per endpoint (hashicorp/golang-lru semantics).  The host keeps no copy. */
/* PreRequest for the batch (plugin.go:169-197): for r in order: indexer.Add(hashes[r], pick[r]) — one kernel, one CTA per
 * endpoint (Adds for different endpoints commute), calls replayed in request order within an endpoint.
 * lru_capacity: optional HOST [M] CacheNumBlocks per endpoint (autotune, plugin.go:207-216); NULL/<=0 ⇒ default. */
int32_t eppscore_commit_picks(struct eppscore_engine *e, int32_t R, const int32_t *pick, const uint64_t *hashes,
                              const uint16_t *n_hashes, int32_t hash_stride, const int32_t *lru_capacity);
/* The same with DEVICE arrays (e.g. the pick / hashes_out buffers of a device-location eppscore_schedule_batch): nothing
 * crosses PCIe, the call is asynchronous and ordered on `stream` (NULL = engine stream).  touch_bound: an upper bound on
 * Σ n_hashes (room is guaranteed up front); <= 0 ⇒ the engine sums n_hashes itself (one small synchronous read-back). */
int32_t eppscore_commit_picks_device(struct eppscore_engine *e, int32_t R, const int32_t *pick, const uint64_t *hashes,
                                     const uint16_t *n_hashes, int32_t hash_stride, const int32_t *lru_capacity,
                                     int64_t touch_bound, void *stream);
/* indexer.Add for one server (tests, the "prefill" profile's pick plugin.go:180-184). */
int32_t eppscore_prefix_add(struct eppscore_engine *e, const uint64_t *hashes, int32_t n, int32_t endpoint,
                            int32_t lru_capacity);
/* Raw deltas for hosts that keep their own LRU: op 0 = insert (hash,endpoint), 1 = evict; applied in order. */
int32_t eppscore_prefix_apply(struct eppscore_engine *e, int64_t n, const uint64_t *hash, const int32_t *endpoint,
                              const uint8_t *op);
int32_t eppscore_prefix_remove_endpoint(struct eppscore_engine *e, int32_t endpoint); /* indexer.RemovePod :167-182 */
/* indexer.Get (reads the DEVICE table): bitset_out[ceil(M/32)] words, natural order; returns set size or <0. */
int32_t eppscore_prefix_get(struct eppscore_engine *e, uint64_t hash, uint32_t *bitset_out, int32_t words);
int32_t eppscore_prefix_lru_len(const struct eppscore_engine *e, int32_t endpoint); /* -1: endpoint has no LRU */
int32_t eppscore_prefix_lru_keys(const struct eppscore_engine *e, int32_t endpoint, uint64_t *out, int32_t cap); /* oldest first */
/* Replication across GPUs: the index is a deterministic function of the ordered commit stream, so every rank applies the
 * same eppscore_commit_picks[_device] calls (the shards' picks all-gathered in global request order) — SURVEY §8e. */

/* pinned host memory for callers that want the fast copy path (cudaHostAlloc / cudaFreeHost) */
void *eppscore_host_alloc(size_t bytes);
void eppscore_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* EPPSCORE_H */
