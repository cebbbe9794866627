/*
 * oracle.c — CPU restatement of the reference Endpoint-Picker hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product library.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared -pthread (oracle/Makefile).
 * -ffp-contract=off matters: the shipped reference is GOARCH=amd64/GOAMD64=v1 (Dockerfile:9-10),
 * for which Go never fuses x += a*b, so the weighted sum is mul-then-add in float64.
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * XXH64 — public specification (Yann Collet), as implemented by cespare/xxhash/v2 v2.3.0.
 * Call sites restated: approximateprefix/hashing.go:70-94 (New/Write/Reset/Sum64, seed 0).
 * ===================================================================================== */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8); /* little-endian host (x86-64) */
  return v;
}
static inline uint32_t rd32(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
static inline uint64_t xxh_round(uint64_t acc, uint64_t in) {
  acc += in * P2;
  acc = rotl64(acc, 31);
  return acc * P1;
}
static inline uint64_t xxh_merge(uint64_t h, uint64_t v) {
  v = xxh_round(0, v);
  h ^= v;
  return h * P1 + P4;
}

uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed) {
  const uint8_t *p = (const uint8_t *)data;
  const uint8_t *end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t *limit = end - 32;
    do {
      v1 = xxh_round(v1, rd64(p));
      v2 = xxh_round(v2, rd64(p + 8));
      v3 = xxh_round(v3, rd64(p + 16));
      v4 = xxh_round(v4, rd64(p + 24));
      p += 32;
    } while (p <= limit);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxh_merge(h, v1);
    h = xxh_merge(h, v2);
    h = xxh_merge(h, v3);
    h = xxh_merge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) {
    h ^= xxh_round(0, rd64(p));
    h = rotl64(h, 27) * P1 + P4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= (uint64_t)rd32(p) * P1;
    h = rotl64(h, 23) * P2 + P3;
    p += 4;
  }
  while (p < end) {
    h ^= (uint64_t)(*p) * P5;
    h = rotl64(h, 11) * P1;
    p++;
  }
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}

/* hashing.go:70-77: h := xxhash.New(); h.Write(model); if salt != "" { h.Write(salt) }; h.Sum64().
 * Streaming writes == one-shot over the concatenation. */
uint64_t orc_model_seed(const void *model, size_t model_len, const void *salt, size_t salt_len) {
  size_t n = model_len + salt_len;
  uint8_t stack[256] = {0};
  uint8_t *buf = n <= sizeof(stack) ? stack : (uint8_t *)malloc(n);
  if (model_len) memcpy(buf, model, model_len);
  if (salt_len) memcpy(buf + model_len, salt, salt_len);
  uint64_t h = orc_xxh64(buf, n, 0);
  if (buf != stack) free(buf);
  return h;
}

/* One chained link: XXH64(block || LE64(prev))  (hashing.go:79-85, toBytes :100-104). */
static uint64_t hash_link(const uint8_t *block, size_t blen, uint64_t prev) {
  uint8_t stack[512];
  size_t n = blen + 8;
  uint8_t *buf = n <= sizeof(stack) ? stack : (uint8_t *)malloc(n);
  memcpy(buf, block, blen);
  memcpy(buf + blen, &prev, 8); /* binary.LittleEndian.PutUint64 */
  uint64_t h = orc_xxh64(buf, n, 0);
  if (buf != stack) free(buf);
  return h;
}

/* hashPrompt, hashing.go:34-98 (userInput already flattened by the host, getUserInputBytes :106). */
int32_t orc_hash_prompt(const uint8_t *input, int64_t len, uint64_t model_seed, int32_t block_chars,
                        int32_t max_blocks, uint64_t *out, int32_t out_cap) {
  if (block_chars <= 0) return 0;                 /* :51-56 */
  if (len < (int64_t)block_chars) return 0;       /* :57-60 */
  int64_t cap = (int64_t)block_chars * (int64_t)max_blocks;
  if (len > cap) len = cap;                       /* :62-65 (maxBlocks<=0 ⇒ nothing left) */
  if (len < 0) len = 0;
  uint64_t prev = model_seed;                     /* :78 */
  int32_t n = 0;
  int64_t i = 0;
  for (; i + block_chars <= len; i += block_chars) { /* :80-87 */
    if (n >= out_cap) return -1;
    prev = hash_link(input + i, (size_t)block_chars, prev);
    out[n++] = prev;
  }
  if (i < len) {                                  /* :89-95 trailing partial block */
    if (n >= out_cap) return -1;
    out[n++] = hash_link(input + i, (size_t)(len - i), prev);
  }
  return n;
}

/* =====================================================================================
 * u64 -> u32 open-addressing map with backward-shift deletion (helper; stands in for Go maps).
 * ===================================================================================== */
typedef struct {
  uint64_t *key;
  uint32_t *val; /* 0 = empty slot, else value+1 */
  uint64_t cap;  /* power of two */
  uint64_t n;
} u64map;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
static void u64map_init(u64map *m, uint64_t cap) {
  uint64_t c = 16;
  while (c < cap) c <<= 1;
  m->cap = c;
  m->n = 0;
  m->key = (uint64_t *)calloc(c, sizeof(uint64_t));
  m->val = (uint32_t *)calloc(c, sizeof(uint32_t));
}
static void u64map_free(u64map *m) {
  free(m->key);
  free(m->val);
  m->key = NULL;
  m->val = NULL;
}
static int64_t u64map_find(const u64map *m, uint64_t k) {
  uint64_t mask = m->cap - 1, i = mix64(k) & mask;
  while (m->val[i]) {
    if (m->key[i] == k) return (int64_t)i;
    i = (i + 1) & mask;
  }
  return -1;
}
static void u64map_put_nogrow(u64map *m, uint64_t k, uint32_t v) {
  uint64_t mask = m->cap - 1, i = mix64(k) & mask;
  while (m->val[i]) {
    if (m->key[i] == k) {
      m->val[i] = v + 1;
      return;
    }
    i = (i + 1) & mask;
  }
  m->key[i] = k;
  m->val[i] = v + 1;
  m->n++;
}
static void u64map_put(u64map *m, uint64_t k, uint32_t v) {
  if ((m->n + 1) * 2 > m->cap) {
    u64map old = *m;
    u64map_init(m, old.cap * 2);
    for (uint64_t i = 0; i < old.cap; i++)
      if (old.val[i]) u64map_put_nogrow(m, old.key[i], old.val[i] - 1);
    u64map_free(&old);
  }
  u64map_put_nogrow(m, k, v);
}
static int u64map_get(const u64map *m, uint64_t k, uint32_t *v) {
  int64_t i = u64map_find(m, k);
  if (i < 0) return 0;
  *v = m->val[i] - 1;
  return 1;
}
static void u64map_del(u64map *m, uint64_t k) {
  int64_t f = u64map_find(m, k);
  if (f < 0) return;
  uint64_t mask = m->cap - 1, i = (uint64_t)f, j = i;
  m->val[i] = 0;
  m->n--;
  for (;;) {
    j = (j + 1) & mask;
    if (!m->val[j]) break;
    uint64_t home = mix64(m->key[j]) & mask;
    /* move j into the hole i unless home lies cyclically in (i, j] */
    int between = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
    if (!between) {
      m->key[i] = m->key[j];
      m->val[i] = m->val[j];
      m->val[j] = 0;
      i = j;
    }
  }
}

/* =====================================================================================
 * LRU with the semantics indexer.go relies on from hashicorp/golang-lru/v2 v2.0.7
 * (go.mod:13; call sites indexer.go:64,71,139-140,177-178): Add of an existing key refreshes
 * recency without eviction; a new key is pushed to the front, and if Len() > size the OLDEST
 * is removed and the eviction callback fires; Remove fires the callback; Keys() oldest→newest.
 * ===================================================================================== */
typedef struct {
  int32_t size; /* capacity */
  int32_t len;
  uint64_t *key;
  int32_t *prev, *next; /* node indices; -1 = none */
  int32_t head, tail;   /* head = newest, tail = oldest */
  int32_t free_head;
  int32_t nodes_cap;
  u64map items;
} lru_t;

static lru_t *lru_new(int32_t size) {
  lru_t *l = (lru_t *)calloc(1, sizeof(lru_t));
  l->size = size;
  l->head = l->tail = -1;
  l->free_head = -1;
  l->nodes_cap = 0;
  u64map_init(&l->items, 64);
  return l;
}
static void lru_free(lru_t *l) {
  if (!l) return;
  free(l->key);
  free(l->prev);
  free(l->next);
  u64map_free(&l->items);
  free(l);
}
static int32_t lru_alloc_node(lru_t *l) {
  if (l->free_head >= 0) {
    int32_t n = l->free_head;
    l->free_head = l->next[n];
    return n;
  }
  if (l->len >= l->nodes_cap) {
    int32_t nc = l->nodes_cap ? l->nodes_cap * 2 : 64;
    l->key = (uint64_t *)realloc(l->key, sizeof(uint64_t) * (size_t)nc);
    l->prev = (int32_t *)realloc(l->prev, sizeof(int32_t) * (size_t)nc);
    l->next = (int32_t *)realloc(l->next, sizeof(int32_t) * (size_t)nc);
    /* nodes [nodes_cap, nc) are fresh: thread all but the first onto the free list */
    for (int32_t i = nc - 1; i > l->nodes_cap; i--) {
      l->next[i] = l->free_head;
      l->free_head = i;
    }
    int32_t n = l->nodes_cap;
    l->nodes_cap = nc;
    return n;
  }
  return -1; /* unreachable */
}
static void lru_unlink(lru_t *l, int32_t n) {
  int32_t p = l->prev[n], q = l->next[n];
  if (p >= 0) l->next[p] = q; else l->head = q;
  if (q >= 0) l->prev[q] = p; else l->tail = p;
}
static void lru_push_front(lru_t *l, int32_t n) {
  l->prev[n] = -1;
  l->next[n] = l->head;
  if (l->head >= 0) l->prev[l->head] = n;
  l->head = n;
  if (l->tail < 0) l->tail = n;
}
/* returns 1 and sets *evicted when an eviction happened */
static int lru_add(lru_t *l, uint64_t k, uint64_t *evicted) {
  uint32_t n;
  if (u64map_get(&l->items, k, &n)) {
    lru_unlink(l, (int32_t)n);
    lru_push_front(l, (int32_t)n);
    return 0;
  }
  int32_t node = lru_alloc_node(l);
  l->key[node] = k;
  lru_push_front(l, node);
  u64map_put(&l->items, k, (uint32_t)node);
  l->len++;
  if (l->len > l->size) {
    int32_t t = l->tail;
    *evicted = l->key[t];
    lru_unlink(l, t);
    u64map_del(&l->items, l->key[t]);
    l->next[t] = l->free_head;
    l->free_head = t;
    l->len--;
    return 1;
  }
  return 0;
}

/* =====================================================================================
 * indexer — approximateprefix/indexer.go:32-195
 * ===================================================================================== */
typedef struct {
  int32_t *ids;
  int32_t n, cap;
} podset;

struct orc_index {
  u64map hash_to_set;   /* hashToPods: blockHash -> index into sets[] */
  podset *sets;
  int32_t sets_len, sets_cap;
  int32_t *free_sets;
  int32_t free_len, free_cap;
  lru_t **pod_lru;      /* podToLRU, indexed by server id */
  int32_t pods_cap;
  int32_t default_lru;
};

orc_index *orc_index_new(int32_t default_lru_size) {
  orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
  u64map_init(&ix->hash_to_set, 1024);
  ix->default_lru = default_lru_size;
  return ix;
}
void orc_index_free(orc_index *ix) {
  if (!ix) return;
  for (int32_t i = 0; i < ix->sets_len; i++) free(ix->sets[i].ids);
  free(ix->sets);
  free(ix->free_sets);
  for (int32_t i = 0; i < ix->pods_cap; i++) lru_free(ix->pod_lru[i]);
  free(ix->pod_lru);
  u64map_free(&ix->hash_to_set);
  free(ix);
}

/* makeEvictionFn, indexer.go:105-115: delete(podSet, pod); if len(podSet)==0 delete(hashToPods, hash) */
static void index_evict(orc_index *ix, uint64_t hash, int32_t server) {
  uint32_t si;
  if (!u64map_get(&ix->hash_to_set, hash, &si)) return;
  podset *s = &ix->sets[si];
  for (int32_t i = 0; i < s->n; i++)
    if (s->ids[i] == server) {
      s->ids[i] = s->ids[s->n - 1];
      s->n--;
      break;
    }
  if (s->n == 0) {
    u64map_del(&ix->hash_to_set, hash);
    if (ix->free_len == ix->free_cap) {
      ix->free_cap = ix->free_cap ? ix->free_cap * 2 : 64;
      ix->free_sets = (int32_t *)realloc(ix->free_sets, sizeof(int32_t) * (size_t)ix->free_cap);
    }
    ix->free_sets[ix->free_len++] = (int32_t)si;
  }
}

static void index_set_add(orc_index *ix, uint64_t hash, int32_t server) {
  uint32_t si;
  if (!u64map_get(&ix->hash_to_set, hash, &si)) {
    if (ix->free_len) {
      si = (uint32_t)ix->free_sets[--ix->free_len];
    } else {
      if (ix->sets_len == ix->sets_cap) {
        ix->sets_cap = ix->sets_cap ? ix->sets_cap * 2 : 1024;
        ix->sets = (podset *)realloc(ix->sets, sizeof(podset) * (size_t)ix->sets_cap);
      }
      si = (uint32_t)ix->sets_len++;
      ix->sets[si].ids = NULL;
      ix->sets[si].cap = 0;
    }
    ix->sets[si].n = 0;
    u64map_put(&ix->hash_to_set, hash, si);
  }
  podset *s = &ix->sets[si];
  for (int32_t i = 0; i < s->n; i++)
    if (s->ids[i] == server) return;
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 4;
    s->ids = (int32_t *)realloc(s->ids, sizeof(int32_t) * (size_t)s->cap);
  }
  s->ids[s->n++] = server;
}

/* Add, indexer.go:52-83. NOTE the reference order: ALL LRU adds first (evictions fire inside
 * this loop), THEN the hashToPods update for ALL hashes — so a batch longer than the LRU
 * capacity leaves hashToPods entries that are no longer in the LRU (restated faithfully). */
void orc_index_add(orc_index *ix, const uint64_t *hashes, int32_t n, int32_t server,
                   int32_t num_gpu_blocks) {
  if (server < 0) return;
  if (server >= ix->pods_cap) {
    int32_t nc = ix->pods_cap ? ix->pods_cap : 16;
    while (nc <= server) nc *= 2;
    ix->pod_lru = (lru_t **)realloc(ix->pod_lru, sizeof(lru_t *) * (size_t)nc);
    for (int32_t i = ix->pods_cap; i < nc; i++) ix->pod_lru[i] = NULL;
    ix->pods_cap = nc;
  }
  lru_t *l = ix->pod_lru[server];
  if (!l) {                                   /* :57-68 */
    int32_t size = num_gpu_blocks;
    if (size <= 0) size = ix->default_lru;
    if (size <= 0) size = 1; /* lru.NewWithEvict errors on size<=0; reference ignores the error (nil LRU would panic) */
    l = lru_new(size);
    ix->pod_lru[server] = l;
  }
  for (int32_t i = 0; i < n; i++) {           /* :70-72 */
    uint64_t ev;
    if (lru_add(l, hashes[i], &ev)) index_evict(ix, ev, server);
  }
  for (int32_t i = 0; i < n; i++) index_set_add(ix, hashes[i], server); /* :75-82 */
}

int32_t orc_index_get(const orc_index *ix, uint64_t hash, int32_t *servers_out, int32_t cap) {
  uint32_t si;
  if (!u64map_get(&ix->hash_to_set, hash, &si)) return 0;
  const podset *s = &ix->sets[si];
  for (int32_t i = 0; i < s->n && i < cap; i++) servers_out[i] = s->ids[i];
  return s->n;
}

/* RemovePod, indexer.go:167-182: for hash in lru.Keys(): lru.Remove(hash) (fires eviction cb); delete(podToLRU, pod) */
void orc_index_remove_pod(orc_index *ix, int32_t server) {
  if (server < 0 || server >= ix->pods_cap || !ix->pod_lru[server]) return;
  lru_t *l = ix->pod_lru[server];
  for (int32_t n = l->tail; n >= 0; n = l->prev[n]) index_evict(ix, l->key[n], server);
  lru_free(l);
  ix->pod_lru[server] = NULL;
}

int32_t orc_index_lru_len(const orc_index *ix, int32_t server) {
  if (server < 0 || server >= ix->pods_cap || !ix->pod_lru[server]) return -1;
  return ix->pod_lru[server]->len;
}
int32_t orc_index_lru_keys(const orc_index *ix, int32_t server, uint64_t *out, int32_t cap) {
  if (server < 0 || server >= ix->pods_cap || !ix->pod_lru[server]) return -1;
  const lru_t *l = ix->pod_lru[server];
  int32_t k = 0;
  for (int32_t n = l->tail; n >= 0 && k < cap; n = l->prev[n]) out[k++] = l->key[n];
  return l->len;
}
int64_t orc_index_num_hashes(const orc_index *ix) { return (int64_t)ix->hash_to_set.n; }
int32_t orc_index_pods(const orc_index *ix, int32_t *out, int32_t cap) {
  int32_t k = 0;
  for (int32_t i = 0; i < ix->pods_cap; i++)
    if (ix->pod_lru[i]) {
      if (k < cap) out[k] = i;
      k++;
    }
  return k;
}
int64_t orc_index_dump(const orc_index *ix, uint64_t *hash_out, int32_t *server_out, int64_t cap) {
  int64_t k = 0;
  for (uint64_t i = 0; i < ix->hash_to_set.cap; i++) {
    if (!ix->hash_to_set.val[i]) continue;
    const podset *s = &ix->sets[ix->hash_to_set.val[i] - 1];
    for (int32_t j = 0; j < s->n; j++) {
      if (k < cap) {
        hash_out[k] = ix->hash_to_set.key[i];
        server_out[k] = s->ids[j];
      }
      k++;
    }
  }
  return k;
}

/* matchLongestPrefix, plugin.go:219-235: greedy walk; STOP at the first hash with an empty pod set;
 * every pod in a non-empty set gets +1 (count of matching blocks before the first GLOBAL miss). */
void orc_match_longest_prefix(const orc_index *ix, const uint64_t *hashes, int32_t n, int32_t M,
                              uint16_t *match_out) {
  memset(match_out, 0, sizeof(uint16_t) * (size_t)M);
  if (!ix) return;
  for (int32_t i = 0; i < n; i++) {
    uint32_t si;
    if (!u64map_get(&ix->hash_to_set, hashes[i], &si)) break;
    const podset *s = &ix->sets[si];
    if (s->n == 0) break;
    for (int32_t j = 0; j < s->n; j++)
      if (s->ids[j] >= 0 && s->ids[j] < M) match_out[s->ids[j]]++;
  }
}

/* =====================================================================================
 * Scorers
 * ===================================================================================== */
static inline int is_cand(const uint32_t *mask, int32_t m) {
  return !mask || ((mask[m >> 5] >> (m & 31)) & 1u);
}

double orc_enforce_score_range(double s) { /* scheduler_profile.go:194-202 */
  if (s < 0) return 0;
  if (s > 1) return 1;
  return s;
}

/* kvcache_utilization.go:76-82 */
void orc_score_kv(const orc_snapshot *s, const uint32_t *mask, double *out) {
  for (int32_t m = 0; m < s->M; m++)
    if (is_cand(mask, m)) out[m] = 1 - s->kv_usage[m];
}

/* queue.go:78-108 (and runningrequest.go:78-108, identical form on RunningRequestsSize) */
static void score_minmax(int32_t M, const int64_t *q, const uint32_t *mask, double *out) {
  int64_t mn = INT64_MAX, mx = INT64_MIN; /* math.MaxInt / math.MinInt */
  for (int32_t m = 0; m < M; m++) {
    if (!is_cand(mask, m)) continue;
    if (q[m] < mn) mn = q[m];
    if (q[m] > mx) mx = q[m];
  }
  for (int32_t m = 0; m < M; m++) {
    if (!is_cand(mask, m)) continue;
    if (mx == mn)
      out[m] = 1.0;
    else
      out[m] = (double)(mx - q[m]) / (double)(mx - mn);
  }
}
void orc_score_queue(const orc_snapshot *s, const uint32_t *mask, double *out) {
  score_minmax(s->M, s->queue, mask, out);
}
void orc_score_running(const orc_snapshot *s, const uint32_t *mask, double *out) {
  score_minmax(s->M, s->running, mask, out);
}

/* lora_affinity.go:76-102; adapter_id < 0 or out of the dictionary ⇒ neither active nor waiting */
void orc_score_lora(const orc_snapshot *s, const uint32_t *mask, int32_t adapter_id, double *out) {
  for (int32_t m = 0; m < s->M; m++) {
    if (!is_cand(mask, m)) continue;
    int active = 0, waiting = 0;
    if (adapter_id >= 0 && adapter_id < s->lora_words * 64 && s->lora_active && s->lora_waiting) {
      int w = adapter_id >> 6, b = adapter_id & 63;
      active = (int)((s->lora_active[(size_t)m * s->lora_words + w] >> b) & 1);
      waiting = (int)((s->lora_waiting[(size_t)m * s->lora_words + w] >> b) & 1);
    }
    int32_t nmodels = s->lora_nmodels ? s->lora_nmodels[m] : 0;
    int32_t maxm = s->lora_max ? s->lora_max[m] : 0;
    if (active)
      out[m] = 1.0;
    else if (nmodels < maxm)
      out[m] = 0.8;
    else if (waiting)
      out[m] = 0.6;
    else
      out[m] = 0.0;
  }
}

/* prefix/plugin.go:95-117: 0 when the attribute is absent or total==0, else match/total */
void orc_score_prefix(int32_t M, const uint32_t *mask, const uint16_t *match, int32_t total,
                      int32_t have_info, double *out) {
  for (int32_t m = 0; m < M; m++) {
    if (!is_cand(mask, m)) continue;
    out[m] = 0.0;
    if (have_info && match && total != 0) out[m] = (double)match[m] / (double)total;
  }
}


/* token_load.go:83-111: 1.0 when no in-flight tokens, else 1 - min(tokens, thr)/thr */
void orc_score_token_load(const orc_snapshot *s, const uint32_t *mask, double threshold, double *out) {
  if (threshold <= 0) threshold = 4194304.0; /* tokenQueueThresholdDefault, token_load.go:33,57-61 */
  for (int32_t m = 0; m < s->M; m++) {
    if (!is_cand(mask, m)) continue;
    double load = s->inflight_tokens ? (double)s->inflight_tokens[m] : 0.0;
    double score;
    if (load <= 0) {
      score = 1.0;
    } else {
      if (load > threshold) load = threshold;
      score = 1.0 - (load / threshold);
    }
    out[m] = score;
  }
}

/* Go unicode/utf8 DecodeRune: returns the rune (0xFFFD for any invalid or short sequence, width 1) */
static uint32_t go_decode_rune(const uint8_t *s, int64_t n, int *width) {
  uint8_t b0 = s[0];
  *width = 1;
  if (b0 < 0x80) return b0;
  int need;
  uint8_t lo = 0x80, hi = 0xBF; /* accept range of the SECOND byte */
  if (b0 >= 0xC2 && b0 <= 0xDF) {
    need = 2;
  } else if (b0 >= 0xE0 && b0 <= 0xEF) {
    need = 3;
    if (b0 == 0xE0) lo = 0xA0;
    if (b0 == 0xED) hi = 0x9F;
  } else if (b0 >= 0xF0 && b0 <= 0xF4) {
    need = 4;
    if (b0 == 0xF0) lo = 0x90;
    if (b0 == 0xF4) hi = 0x8F;
  } else {
    return 0xFFFD;
  }
  if (n < need) return 0xFFFD;
  if (s[1] < lo || s[1] > hi) return 0xFFFD;
  for (int i = 2; i < need; i++)
    if (s[i] < 0x80 || s[i] > 0xBF) return 0xFFFD;
  *width = need;
  if (need == 2) return ((uint32_t)(b0 & 0x1F) << 6) | (s[1] & 0x3F);
  if (need == 3) return ((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F);
  return ((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) | ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
}
/* unicode.IsSpace: Latin-1 set, else the White_Space property */
static int go_is_space(uint32_t r) {
  if (r <= 0xFF) return r == 0x09 || r == 0x0A || r == 0x0B || r == 0x0C || r == 0x0D || r == 0x20 || r == 0x85 || r == 0xA0;
  return r == 0x1680 || (r >= 0x2000 && r <= 0x200A) || r == 0x2028 || r == 0x2029 || r == 0x202F || r == 0x205F ||
         r == 0x3000;
}
int32_t orc_count_fields(const uint8_t *s, int64_t len) {
  int32_t n = 0;
  int in_field = 0;
  for (int64_t i = 0; i < len;) {
    int w;
    uint32_t r = go_decode_rune(s + i, len - i, &w);
    if (go_is_space(r)) {
      in_field = 0;
    } else if (!in_field) {
      in_field = 1;
      n++;
    }
    i += w;
  }
  return n;
}

/* =====================================================================================
 * Latency-predictor fold-in: predicted-latency producer + latency-scorer
 * ===================================================================================== */
/* sidecars/latencypredictorasync/prediction.go:164-194 (left-to-right float64, no fusing) */
void orc_latency_predict(const orc_latency_params *lp, double kv, int64_t input_tokens, int64_t waiting,
                         int64_t running, int64_t generated, double prefix_score, double *ttft, double *tpot) {
  *ttft = lp->ttft_intercept + lp->ttft_kv * kv + lp->ttft_input * (double)input_tokens +
          lp->ttft_waiting * (double)waiting + lp->ttft_running * (double)running + lp->ttft_prefix * prefix_score;
  *tpot = lp->tpot_intercept + lp->tpot_kv * kv + lp->tpot_input * (double)input_tokens +
          lp->tpot_waiting * (double)waiting + lp->tpot_running * (double)running +
          lp->tpot_generated * (double)generated;
}

/* predictedlatency/prediction.go:137-166, then :100-104 */
void orc_latency_validate(const orc_latency_params *lp, double ttft, double tpot, double ttft_slo, double tpot_slo,
                          double pod_min, int32_t neutralize, int32_t *ok_out, double *headroom_out) {
  int ttft_ok = ttft < ttft_slo;
  double ttft_headroom = ttft_slo - ttft;
  int tpot_ok = 1;
  double headroom = 0.0;
  if (lp->streaming_mode) {
    double buffered = tpot_slo * lp->slo_buffer_factor;
    if (pod_min > 0) {
      double alt = pod_min * lp->slo_buffer_factor;
      if (alt < buffered) buffered = alt; /* min(bufferedTPOT, podMinTPOTSLO*factor) */
    }
    tpot_ok = tpot < buffered;
    headroom = buffered - tpot;
  }
  int valid = ttft_ok && tpot_ok;
  if (neutralize) { /* !StreamingMode || hasPrefillRole */
    tpot_ok = 1;
    headroom = 0;
    valid = ttft_ok;
  }
  if (ok_out) {
    ok_out[0] = ttft_ok;
    ok_out[1] = tpot_ok;
    ok_out[2] = valid;
  }
  headroom_out[0] = headroom;
  headroom_out[1] = ttft_headroom;
}

typedef struct {
  int32_t m;
  int has_info;
  double th, ph; /* ttftHeadroom, tpotHeadroom */
  int32_t dispatched;
} lat_ep;

/* scorer/latency/plugin.go:246-318 */
static void lat_score_bucket(const orc_latency_params *lp, const lat_ep *d, int32_t n, double *scores, int force_least) {
  double alpha, beta; /* normalizedWeights :373-379 */
  {
    double sum = lp->ttft_weight + lp->tpot_weight;
    if (sum <= 0) {
      alpha = 1.0;
      beta = 0.0;
    } else {
      alpha = lp->ttft_weight / sum;
      beta = lp->tpot_weight / sum;
    }
  }
  double min_t = DBL_MAX, max_t = -DBL_MAX, min_p = DBL_MAX, max_p = -DBL_MAX;
  for (int32_t i = 0; i < n; i++) {
    double h = fabs(d[i].th);
    if (h < min_t) min_t = h;
    if (h > max_t) max_t = h;
    h = fabs(d[i].ph);
    if (h < min_p) min_p = h;
    if (h > max_p) max_p = h;
  }
  const double eps = 1e-9;
  double range_t = max_t - min_t, range_p = max_p - min_p;
  if (range_t <= eps && range_p > eps) {
    alpha = 0.0;
    beta = 1.0;
  } else if (range_p <= eps && range_t > eps) {
    alpha = 1.0;
    beta = 0.0;
  }
  for (int32_t i = 0; i < n; i++) {
    double nt, np;
    if (range_t > eps)
      nt = (fabs(d[i].th) - min_t) / range_t;
    else
      nt = 0.5;
    if (range_p > eps)
      np = (fabs(d[i].ph) - min_p) / range_p;
    else
      np = 0.5;
    double combined = alpha * nt + beta * np;
    double w;
    if (lp->strategy_most && !force_least)
      w = (double)((int64_t)(combined * 100.0) + 0 + 1); /* int() truncates toward zero */
    else
      w = (double)((int64_t)((1.0 - combined) * 100.0) + 0 + 1);
    scores[d[i].m] = w / 100.0;
  }
}

/* scorer/latency/plugin.go:323-367 */
static void lat_composite(const orc_latency_params *lp, const orc_snapshot *s, const uint32_t *mask,
                          const uint16_t *match, int32_t total, double *out) {
  double wkv = lp->composite_kv, wq = lp->composite_queue, wpref = lp->composite_prefix;
  double sumw = wkv + wq + wpref;
  if (sumw <= 0) {
    wkv = 1;
    wq = 0;
    wpref = 0;
    sumw = 1;
  }
  wkv /= sumw;
  wq /= sumw;
  wpref /= sumw;
  int64_t max_q = 0;
  for (int32_t m = 0; m < s->M; m++)
    if (is_cand(mask, m) && s->queue[m] > max_q) max_q = s->queue[m];
  double q_range = (double)max_q;
  for (int32_t m = 0; m < s->M; m++) {
    if (!is_cand(mask, m)) continue;
    double rel = 1.0;
    if (q_range > 0) rel = (double)(max_q - s->queue[m]) / q_range;
    double kv_free = 1.0 - s->kv_usage[m];
    double prefix = 0; /* prefixCacheScore :381-392 */
    if (match && total > 0) {
      double sc = (double)match[m] / (double)total;
      if (!isnan(sc)) prefix = sc;
    }
    double composite = wkv * kv_free + wq * rel + wpref * prefix;
    int64_t w = (int64_t)round(0.0 + 100.0 * composite); /* math.Round: half away from zero */
    out[m] = (double)w / 100.0;
  }
}

void orc_score_latency_info(const orc_latency_params *lp, const orc_snapshot *s, const uint32_t *mask,
                            const uint8_t *have_info, const double *ttft_headroom, const double *tpot_headroom,
                            const int32_t *dispatched, const uint16_t *match, int32_t total, double *out) {
  const int32_t M = s->M;
  lat_ep *data = (lat_ep *)malloc(sizeof(lat_ep) * (size_t)(M > 0 ? M : 1) * 2);
  lat_ep *tmp = data + (M > 0 ? M : 1);
  int32_t n = 0;
  int has_predictions = 0;
  for (int32_t m = 0; m < M; m++) { /* :147-166 */
    if (!is_cand(mask, m)) continue;
    out[m] = 0;
    lat_ep d = {m, 0, 0.0, 0.0, 0};
    if (have_info && have_info[m]) {
      d.has_info = 1;
      d.th = ttft_headroom[m];
      d.ph = tpot_headroom[m];
      d.dispatched = dispatched ? dispatched[m] : 0;
      has_predictions = 1;
    }
    data[n++] = d;
  }
  if (!has_predictions) { /* :169-172 */
    lat_composite(lp, s, mask, match, total, out);
    free(data);
    return;
  }
  /* :177-189 positive bucket first */
  int32_t k = 0;
  for (int32_t i = 0; i < n; i++)
    if (!(data[i].has_info && (data[i].th < 0 || data[i].ph < 0))) tmp[k++] = data[i];
  if (k > 0) {
    lat_score_bucket(lp, tmp, k, out, 0);
    free(data);
    return;
  }
  /* all negative: idle preference :193-204 */
  k = 0;
  for (int32_t i = 0; i < n; i++)
    if (data[i].has_info && data[i].dispatched == 0) tmp[k++] = data[i];
  if (k > 0) {
    lat_score_bucket(lp, tmp, k, out, 1);
    free(data);
    return;
  }
  /* deficit buckets :209-238, order negTPOTonly, negTTFTonly, bothNeg */
  for (int pass = 0; pass < 3; pass++) {
    k = 0;
    for (int32_t i = 0; i < n; i++) {
      int bucket;
      if (!data[i].has_info) {
        bucket = 2;
      } else {
        int tn = data[i].th < 0, pn = data[i].ph < 0;
        bucket = (tn && pn) ? 2 : (tn ? 1 : (pn ? 0 : 2));
      }
      if (bucket == pass) tmp[k++] = data[i];
    }
    if (k > 0) {
      lat_score_bucket(lp, tmp, k, out, 1);
      free(data);
      return;
    }
  }
  lat_score_bucket(lp, data, n, out, 1); /* :241 (n == 0 here) */
  free(data);
}

void orc_score_latency(const orc_latency_params *lp, const orc_snapshot *s, const uint32_t *mask,
                       const uint16_t *match, int32_t total, const orc_latency_request *lr, double *out,
                       double *pred_out) {
  const int32_t M = s->M;
  orc_latency_request zero = {0, 0.0, 0.0};
  if (!lr) lr = &zero;
  if (!lp->has_predictions) {
    orc_score_latency_info(lp, s, mask, NULL, NULL, NULL, NULL, match, total, out);
    if (pred_out)
      for (int32_t m = 0; m < 2 * M; m++) pred_out[m] = NAN;
    return;
  }
  uint8_t *have = (uint8_t *)malloc((size_t)(M > 0 ? M : 1));
  double *th = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1) * 2);
  double *ph = th + (M > 0 ? M : 1);
  int32_t *disp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
  /* PrepareRequestData runs over every endpoint handed to the scheduler, before the filters */
  for (int32_t m = 0; m < M; m++) {
    double prefix = 0.0; /* preparedata_hooks.go:44-58: match/total, NaN => 0, attribute absent => 0 */
    if (match) {
      prefix = (double)match[m] / (double)total;
      if (isnan(prefix)) prefix = 0.0;
    }
    double ttft, tpot; /* training.go:35-57: generatedTokens = 1 (prediction.go:69) */
    orc_latency_predict(lp, s->kv_usage[m], lr->input_tokens, s->queue[m], s->running ? s->running[m] : 0, 1, prefix,
                        &ttft, &tpot);
    int neutralize = !lp->streaming_mode || (s->prefill_role && s->prefill_role[m]);
    double hr[2];
    orc_latency_validate(lp, ttft, tpot, lr->ttft_slo, lr->tpot_slo, s->min_tpot_slo ? s->min_tpot_slo[m] : 0.0,
                         neutralize, NULL, hr);
    have[m] = 1;
    th[m] = hr[1];
    ph[m] = hr[0];
    disp[m] = s->dispatched ? s->dispatched[m] : 0;
    if (pred_out) {
      pred_out[2 * m] = ttft;
      pred_out[2 * m + 1] = tpot;
    }
  }
  orc_score_latency_info(lp, s, mask, have, th, ph, disp, match, total, out);
  free(have);
  free(th);
  free(disp);
}

static inline uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
/* Tie priority (NOT from the reference — the reference is a time-seeded shuffle,
 * picker/common.go:49-55; this is the engine's documented deterministic stand-in). */
uint32_t orc_tie_priority(uint64_t seed, int64_t request_index, int32_t endpoint) {
  uint32_t a = lowbias32((uint32_t)(uint64_t)request_index ^ (uint32_t)seed);
  return lowbias32(a + (uint32_t)endpoint * 0x9E3779B1U + (uint32_t)(seed >> 32));
}

/* 53-bit uniform in (0,1] from the tie priority (documented generator, see include/eppscore.h pick modes) */
double orc_uniform01(uint64_t seed, int64_t request_index, int32_t endpoint) {
  uint32_t prio = orc_tie_priority(seed, request_index, endpoint);
  uint32_t hi = lowbias32(prio ^ 0x85EBCA6BU);
  uint32_t lo = lowbias32(hi + 0xC2B2AE35U + (uint32_t)endpoint);
  uint64_t k = ((((uint64_t)hi) << 32) | lo) >> 11;
  return (double)(k + 1) * 0x1p-53;
}
/* -ln(u), u in (0,1]: u = f * 2^e with f in (sqrt(1/2), sqrt(2)], ln f = 2 atanh((f-1)/(f+1)) as an 11-term odd series */
double orc_neg_log(double u) {
  uint64_t bits;
  memcpy(&bits, &u, 8);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double f;
  memcpy(&f, &bits, 8);
  if (f > 1.4142135623730951) {
    f = f * 0.5;
    e += 1;
  }
  double z = (f - 1.0) / (f + 1.0);
  double z2 = z * z;
  double p = 1.0 / 21.0;
  for (int k = 19; k >= 1; k -= 2) p = p * z2 + 1.0 / (double)k;
  double lnf = (2.0 * z) * p;
  double lnu = (double)e * 0.6931471805599453 + lnf;
  return -lnu;
}

/* =====================================================================================
 * Device-side filters of the latency profile
 * ===================================================================================== */
static double filt_prefix_score(const uint16_t *match, int32_t total, int32_t m) { /* prefixcacheaffinity/plugin.go:160-171 */
  if (match && total > 0) {
    double sc = (double)match[m] / (double)total;
    if (!isnan(sc)) return sc;
  }
  return 0;
}

void orc_apply_filters(const orc_snapshot *s, const orc_profile *p, int64_t request_index, const uint32_t *mask_in,
                       const uint16_t *match, int32_t total, const orc_latency_request *lat, uint32_t *mask_out) {
  const int32_t M = s->M, mw = (M + 31) / 32;
  orc_latency_request zero = {0, 0.0, 0.0};
  if (!lat) lat = &zero;
  /* the current candidate list, in endpoint order (filters preserve order) */
  int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1) * 3);
  int32_t *a = cur + (M > 0 ? M : 1), *b = a + (M > 0 ? M : 1);
  int32_t n = 0;
  for (int32_t m = 0; m < M; m++)
    if (is_cand(mask_in, m)) cur[n++] = m;
  /* LatencyPredictionInfo per endpoint, as PrepareRequestData leaves it (all endpoints, before the filters) */
  const int have_pred = p->latency && p->latency->has_predictions;
  double *ttft = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1) * 3);
  double *th = ttft + (M > 0 ? M : 1), *ph = th + (M > 0 ? M : 1);
  if (have_pred)
    for (int32_t m = 0; m < M; m++) {
      double prefix = 0.0, tp;
      if (match) {
        prefix = (double)match[m] / (double)total;
        if (isnan(prefix)) prefix = 0.0;
      }
      orc_latency_predict(p->latency, s->kv_usage[m], lat->input_tokens, s->queue[m], s->running ? s->running[m] : 0, 1, prefix,
                          &ttft[m], &tp);
      int neutralize = !p->latency->streaming_mode || (s->prefill_role && s->prefill_role[m]);
      double hr[2];
      orc_latency_validate(p->latency, ttft[m], tp, lat->ttft_slo, lat->tpot_slo, s->min_tpot_slo ? s->min_tpot_slo[m] : 0.0,
                           neutralize, NULL, hr);
      th[m] = hr[1];
      ph[m] = hr[0];
    }
  for (int32_t f = 0; f < p->n_filters && n > 0; f++) { /* scheduler_profile.go:135-145: stop on an empty result */
    const double *par = p->filter_param[f];
    const double u = orc_uniform01(p->tie_seed, request_index, -(f + 1));
    if (p->filter_kind[f] == ORC_FILTER_PREFIX_AFFINITY) { /* prefixcacheaffinity/plugin.go:105-151 */
      if (n <= 1 || par[0] <= 0) continue;
      if (u < par[1]) continue; /* exploration: keep all */
      int32_t na = 0, nb = 0;   /* a = sticky, b = nonSticky */
      for (int32_t i = 0; i < n; i++) {
        if (filt_prefix_score(match, total, cur[i]) >= par[0])
          a[na++] = cur[i];
        else
          b[nb++] = cur[i];
      }
      if (na == 0) continue;
      if (par[2] > 0 && nb > 0) {
        double best_a = DBL_MAX, best_b = DBL_MAX; /* bestTTFT :173-184 */
        if (have_pred) {
          for (int32_t i = 0; i < na; i++)
            if (ttft[a[i]] < best_a) best_a = ttft[a[i]];
          for (int32_t i = 0; i < nb; i++)
            if (ttft[b[i]] < best_b) best_b = ttft[b[i]];
        }
        if (best_a - best_b > par[2]) continue; /* TTFT load gate broken: keep all */
      }
      memcpy(cur, a, sizeof(int32_t) * (size_t)na);
      n = na;
    } else if (p->filter_kind[f] == ORC_FILTER_SLO_HEADROOM_TIER) { /* sloheadroomtier/plugin.go:82-137 */
      if (n <= 1) continue;
      if (!have_pred) continue; /* no predictions: keep all */
      int32_t na = 0, nb = 0;   /* a = positive, b = negative */
      for (int32_t i = 0; i < n; i++) {
        if (th[cur[i]] >= 0 && ph[cur[i]] >= 0)
          a[na++] = cur[i];
        else
          b[nb++] = cur[i];
      }
      if (na > 0 && nb > 0) {
        if (u < par[0]) {
          memcpy(cur, b, sizeof(int32_t) * (size_t)nb);
          n = nb;
        } else {
          memcpy(cur, a, sizeof(int32_t) * (size_t)na);
          n = na;
        }
      } /* only one tier present: that tier == everything */
    }
  }
  memset(mask_out, 0, sizeof(uint32_t) * (size_t)(mw > 0 ? mw : 1));
  for (int32_t i = 0; i < n; i++) mask_out[cur[i] >> 5] |= 1u << (cur[i] & 31);
  free(cur);
  free(ttft);
}

static const double LORA_CLASS_SCORE[4] = {0.0, 0.6, 0.8, 1.0};

int32_t orc_schedule_one(const orc_snapshot *s, const orc_profile *p, int64_t request_index,
                         int32_t adapter_id, const uint32_t *mask, const uint16_t *match,
                         int32_t total, const float *pair_col, int32_t *pick_out, double *score_out,
                         int32_t *tie_count_out, uint32_t *tie_set_out, double *weighted_out) {
  return orc_schedule_one_lat(s, p, request_index, adapter_id, mask, match, total, pair_col, NULL, pick_out,
                              score_out, tie_count_out, tie_set_out, weighted_out, NULL);
}

int32_t orc_schedule_one_lat(const orc_snapshot *s, const orc_profile *p, int64_t request_index,
                             int32_t adapter_id, const uint32_t *mask, const uint16_t *match,
                             int32_t total, const float *pair_col, const orc_latency_request *lat,
                             int32_t *pick_out, double *score_out, int32_t *tie_count_out,
                             uint32_t *tie_set_out, double *weighted_out, double *pred_out) {
  const int32_t M = s->M;
  /* runFilterPlugins (scheduler_profile.go:130-149) is modelled as the candidate mask */
  int32_t ncand = 0;
  for (int32_t m = 0; m < M; m++) ncand += is_cand(mask, m);
  if (tie_set_out) memset(tie_set_out, 0, sizeof(uint32_t) * (size_t)((M + 31) / 32));
  if (ncand == 0) { /* scheduler_profile.go:119-121 */
    *pick_out = -1;
    *score_out = 0.0;
    *tie_count_out = 0;
    if (weighted_out)
      for (int32_t m = 0; m < M; m++) weighted_out[m] = NAN;
    return -1;
  }
  double *w = (double *)malloc(sizeof(double) * (size_t)M * 2);
  double *sc = w + M;
  for (int32_t m = 0; m < M; m++) w[m] = 0.0; /* :155-158 */
  for (int32_t k = 0; k < p->n_scorers; k++) { /* :160-171 */
    int kind = p->scorer_kind[k];
    double weight = p->scorer_weight[k];
    switch (kind) {
      case ORC_SCORER_QUEUE: orc_score_queue(s, mask, sc); break;
      case ORC_SCORER_KV_CACHE: orc_score_kv(s, mask, sc); break;
      case ORC_SCORER_RUNNING: orc_score_running(s, mask, sc); break;
      case ORC_SCORER_TOKEN_LOAD: orc_score_token_load(s, mask, p->token_load_threshold, sc); break;
      case ORC_SCORER_LATENCY:
        if (p->latency)
          orc_score_latency(p->latency, s, mask, match, total, lat, sc, pred_out);
        else
          for (int32_t m = 0; m < M; m++) sc[m] = 0.0;
        break;
      case ORC_SCORER_PREFIX:
        if (pair_col && match == NULL) { /* dense rows carry match in column x */
          for (int32_t m = 0; m < M; m++) {
            sc[m] = 0.0;
            if (total != 0) sc[m] = (double)(uint16_t)pair_col[(size_t)m * 4 + 0] / (double)total;
          }
        } else {
          orc_score_prefix(M, mask, match, total, match != NULL, sc);
        }
        break;
      case ORC_SCORER_LORA:
        if (pair_col) { /* dense rows carry the lora class in column y */
          for (int32_t m = 0; m < M; m++) {
            int c = (int)pair_col[(size_t)m * 4 + 1];
            sc[m] = LORA_CLASS_SCORE[c & 3];
          }
        } else {
          orc_score_lora(s, mask, adapter_id, sc);
        }
        break;
      default:
        if (kind >= ORC_SCORER_ENDPOINT_COL0 && kind < ORC_SCORER_ENDPOINT_COL0 + 4) {
          const double *col = s->endpoint_col[kind - ORC_SCORER_ENDPOINT_COL0];
          for (int32_t m = 0; m < M; m++) sc[m] = col ? col[m] : 0.0;
        } else if (kind >= ORC_SCORER_PAIR_COL0 && kind < ORC_SCORER_PAIR_COL0 + 2) {
          int c = kind - ORC_SCORER_PAIR_COL0;
          for (int32_t m = 0; m < M; m++) sc[m] = pair_col ? (double)pair_col[(size_t)m * 4 + 2 + c] : 0.0;
        } else {
          for (int32_t m = 0; m < M; m++) sc[m] = 0.0;
        }
    }
    for (int32_t m = 0; m < M; m++) {
      if (!is_cand(mask, m)) continue;
      double t = orc_enforce_score_range(sc[m]) * weight; /* rounded product ... */
      w[m] = w[m] + t;                                    /* ... then rounded add (:168) */
    }
  }
  if (p->pick_mode != ORC_PICK_MAX_SCORE) {
    /* weighted-random (A-Res, weightedrandom/picker.go:111-155): key = U^(1/score) maximal  <=>  -ln(U)/score minimal;
     * endpoints with score <= 0 get key 0 (never chosen); no positive score at all => random picker.
     * random picker (random/picker.go:85-101): shuffle and take the first = a uniform choice. */
    int32_t n_pos = 0, pick_pos = -1, pick_all = -1;
    double key_pos = 0;
    uint32_t prio_all = 0;
    for (int32_t m = 0; m < M; m++) {
      if (!is_cand(mask, m)) continue;
      uint32_t pr = orc_tie_priority(p->tie_seed, request_index, m);
      if (pick_all < 0 || pr > prio_all) {
        pick_all = m;
        prio_all = pr;
      }
      if (p->pick_mode == ORC_PICK_WEIGHTED_RANDOM && w[m] > 0) {
        double key = orc_neg_log(orc_uniform01(p->tie_seed, request_index, m)) / w[m];
        if (pick_pos < 0 || key < key_pos) {
          pick_pos = m;
          key_pos = key;
        }
        n_pos++;
      }
    }
    int32_t pk = pick_pos >= 0 ? pick_pos : pick_all;
    *pick_out = pk;
    *score_out = w[pk];
    *tie_count_out = pick_pos >= 0 ? n_pos : ncand; /* size of the set the draw was over */
    if (tie_set_out)
      for (int32_t m = 0; m < M; m++)
        if (is_cand(mask, m) && (pick_pos < 0 || w[m] > 0)) tie_set_out[m >> 5] |= 1u << (m & 31);
    if (weighted_out)
      for (int32_t m = 0; m < M; m++) weighted_out[m] = is_cand(mask, m) ? w[m] : NAN;
    free(w);
    return 0;
  }
  /* MaxScorePicker (maxscore/picker.go:87-115) as arg-max set */
  double best = 0;
  int have = 0;
  for (int32_t m = 0; m < M; m++) {
    if (!is_cand(mask, m)) continue;
    if (!have || w[m] > best) {
      best = w[m];
      have = 1;
    }
  }
  int32_t ties = 0, pick = -1;
  uint32_t best_prio = 0;
  for (int32_t m = 0; m < M; m++) {
    if (!is_cand(mask, m) || !(w[m] == best)) continue;
    ties++;
    if (tie_set_out) tie_set_out[m >> 5] |= 1u << (m & 31);
    if (p->tie_mode == ORC_TIE_SEEDED_RANDOM) {
      uint32_t pr = orc_tie_priority(p->tie_seed, request_index, m);
      if (pick < 0 || pr > best_prio) {
        pick = m;
        best_prio = pr;
      }
    } else if (pick < 0) {
      pick = m;
    }
  }
  *pick_out = pick;
  *score_out = best;
  *tie_count_out = ties;
  if (weighted_out)
    for (int32_t m = 0; m < M; m++) weighted_out[m] = is_cand(mask, m) ? w[m] : NAN;
  free(w);
  return 0;
}

/* =====================================================================================
 * Batch driver: one "goroutine" per request, partitioned over pthreads.
 * ===================================================================================== */
typedef struct {
  const orc_snapshot *s;
  const orc_profile *p;
  const orc_index *idx;
  const orc_batch *b;
  int32_t r0, r1;          /* static range (single-thread call) */
  int32_t *next;           /* shared cursor: workers claim chunks of `chunk` requests (goroutine-style balancing) */
  int32_t chunk;
} job_t;

static int profile_has(const orc_profile *p, int kind) {
  for (int i = 0; i < p->n_scorers; i++)
    if (p->scorer_kind[i] == kind) return 1;
  return 0;
}

static void *batch_worker(void *arg) {
  job_t *j = (job_t *)arg;
  const orc_batch *b = j->b;
  const int32_t M = j->s->M;
  const int32_t mw = (M + 31) / 32;
  const int need_prefix = profile_has(j->p, ORC_SCORER_PREFIX) || profile_has(j->p, ORC_SCORER_LATENCY) || j->p->n_filters > 0 || b->match_blocks || b->total_blocks || b->hashes_out;
  int32_t hcap = b->max_blocks > 0 ? b->max_blocks : 1;
  if (b->hashes_in && b->hash_stride > hcap) hcap = b->hash_stride;
  uint64_t *hashes = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)hcap);
  uint16_t *match = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)M);
  uint32_t *fmask = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(mw > 0 ? mw : 1));
  for (;;) {
  int32_t c0 = j->r0, c1 = j->r1;
  if (j->next) { /* dynamic chunks */
    c0 = __atomic_fetch_add(j->next, j->chunk, __ATOMIC_RELAXED);
    if (c0 >= b->R) break;
    c1 = c0 + j->chunk < b->R ? c0 + j->chunk : b->R;
  }
  for (int32_t r = c0; r < c1; r++) {
    int32_t nh = 0;
    const uint16_t *match_p = NULL;
    int32_t total = 0;
    const float *pair = b->dense_feat ? b->dense_feat + (size_t)r * M * 4 : NULL;
    if (b->dense_feat) {
      total = b->dense_total ? b->dense_total[r] : 0;
    } else if (need_prefix && (b->prompt_bytes || b->hashes_in)) {
      if (b->hashes_in) {
        nh = b->n_hashes_in[r];
        memcpy(hashes, b->hashes_in + (size_t)r * b->hash_stride, sizeof(uint64_t) * (size_t)nh);
      } else {
        nh = orc_hash_prompt(b->prompt_bytes + b->prompt_off[r], b->prompt_off[r + 1] - b->prompt_off[r],
                             b->model_seed ? b->model_seed[r] : 0, b->block_chars, b->max_blocks, hashes, hcap);
        if (nh < 0) nh = 0;
      }
      orc_match_longest_prefix(j->idx, hashes, nh, M, match);
      match_p = match;
      total = nh;
    }
    if (b->hashes_out && !b->dense_feat) {
      int32_t stride = b->max_blocks;
      for (int32_t i = 0; i < stride; i++) b->hashes_out[(size_t)r * stride + i] = i < nh ? hashes[i] : 0;
    }
    if (b->total_blocks) b->total_blocks[r] = (uint16_t)total;
    if (b->match_blocks) {
      for (int32_t m = 0; m < M; m++)
        b->match_blocks[(size_t)r * M + m] =
            pair ? (uint16_t)pair[(size_t)m * 4] : (match_p ? match_p[m] : 0);
    }
    const uint32_t *mask = b->cand_mask ? b->cand_mask + (size_t)r * mw : NULL;
    orc_latency_request lr0 = {b->input_tokens ? b->input_tokens[r] : 0, b->ttft_slo ? b->ttft_slo[r] : 0.0,
                               b->tpot_slo ? b->tpot_slo[r] : 0.0};
    if (j->p->n_filters > 0) {
      orc_apply_filters(j->s, j->p, b->request_base + r, mask, match_p, total, &lr0, fmask);
      mask = fmask;
    }
    if (b->filter_mask_out) {
      for (int32_t w = 0; w < mw; w++) {
        uint32_t v = mask ? mask[w] : 0xffffffffu;
        if (!mask && w == mw - 1 && (M & 31)) v = (1u << (M & 31)) - 1u;
        b->filter_mask_out[(size_t)r * mw + w] = v;
      }
    }
    orc_latency_request lr = {b->input_tokens ? b->input_tokens[r] : 0, b->ttft_slo ? b->ttft_slo[r] : 0.0,
                              b->tpot_slo ? b->tpot_slo[r] : 0.0};
    orc_schedule_one_lat(j->s, j->p, b->request_base + r, b->adapter_id ? b->adapter_id[r] : -1, mask, match_p, total,
                         pair, &lr, &b->pick[r], &b->pick_score[r], &b->tie_count[r],
                         b->tie_set ? b->tie_set + (size_t)r * mw : NULL,
                         b->weighted_out ? b->weighted_out + (size_t)r * M : NULL,
                         b->pred_out ? b->pred_out + (size_t)r * M * 2 : NULL);
  }
  if (!j->next) break;
  }
  free(hashes);
  free(match);
  free(fmask);
  return NULL;
}

/* Persistent worker pool: created once (grown on demand), parked on a condition variable between batches, so a timed
 * batch pays no pthread_create/join — the shape of a Go runtime whose goroutines are scheduled onto existing Ms. */
static struct {
  pthread_mutex_t mu;
  pthread_cond_t go, done;
  pthread_t *th;
  int32_t n_threads;   /* workers created */
  int32_t want;        /* workers that take part in the current batch */
  int32_t running;     /* participants that have not finished yet */
  uint64_t gen;        /* batch generation */
  job_t job;
  pthread_mutex_t call_mu; /* one batch at a time */
  int init;
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, 0,
            {0}, PTHREAD_MUTEX_INITIALIZER, 0};

static void *pool_worker(void *arg) {
  const int32_t id = (int32_t)(intptr_t)arg;
  uint64_t seen = 0;
  pthread_mutex_lock(&g_pool.mu);
  for (;;) {
    while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
    seen = g_pool.gen;
    if (id >= g_pool.want) continue;
    job_t j = g_pool.job;
    pthread_mutex_unlock(&g_pool.mu);
    batch_worker(&j);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.running == 0) pthread_cond_signal(&g_pool.done);
  }
  return NULL;
}

int32_t orc_schedule_batch(const orc_snapshot *s, const orc_profile *p, const orc_index *idx,
                           const orc_batch *b, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > b->R) n_threads = b->R > 0 ? b->R : 1;
  job_t job;
  memset(&job, 0, sizeof(job));
  job.s = s;
  job.p = p;
  job.idx = idx;
  job.b = b;
  job.r0 = 0;
  job.r1 = b->R;
  if (n_threads == 1) {
    batch_worker(&job);
    return 0;
  }
  int32_t cursor = 0;
  job.next = &cursor;
  job.chunk = b->R / (n_threads * 8);
  if (job.chunk < 1) job.chunk = 1;
  if (job.chunk > 256) job.chunk = 256;
  pthread_mutex_lock(&g_pool.call_mu);
  pthread_mutex_lock(&g_pool.mu);
  if (g_pool.n_threads < n_threads) { /* grow the pool; workers never exit */
    g_pool.th = (pthread_t *)realloc(g_pool.th, sizeof(pthread_t) * (size_t)n_threads);
    for (int32_t t = g_pool.n_threads; t < n_threads; t++)
      pthread_create(&g_pool.th[t], NULL, pool_worker, (void *)(intptr_t)t);
    g_pool.n_threads = n_threads;
  }
  g_pool.job = job;
  g_pool.want = n_threads;
  g_pool.running = n_threads;
  g_pool.gen++;
  pthread_cond_broadcast(&g_pool.go);
  while (g_pool.running > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  pthread_mutex_unlock(&g_pool.call_mu);
  return 0;
}

/* PreRequest (plugin.go:169-197) for each request of a batch, in request order. */
void orc_commit_picks(orc_index *ix, int32_t R, const int32_t *pick, const uint64_t *hashes,
                      const uint16_t *n_hashes, int32_t hash_stride, const int32_t *gpu_blocks) {
  for (int32_t r = 0; r < R; r++) {
    if (pick[r] < 0) continue; /* len(TargetEndpoints)==0 ⇒ return (:173-175) */
    int32_t cap = gpu_blocks ? gpu_blocks[pick[r]] : 0; /* makeserver :207-216 (0 ⇒ default) */
    orc_index_add(ix, hashes + (size_t)r * hash_stride, n_hashes[r], pick[r], cap);
  }
}
