// prefix_index.hpp — host side of the prefix index (approximateprefix/indexer.go), B200 layout.
//
// The reference keeps  hashToPods: map[blockHash]podSet  and  podToLRU: map[ServerID]*lru.Cache
// under one RWMutex (indexer.go:32-37).  Here the per-endpoint LRUs stay on the host with the exact
// hashicorp/golang-lru/v2 v2.0.7 semantics the reference relies on (indexer.go:64,71,139-140,177-178),
// while hashToPods lives on the DEVICE as
//     slots[cap]   open addressing, linear probing from (hash & mask): {hash, row, |set|}
//     rows[n][RW]  one bitset row per distinct hash, lane-major permuted (kernels.cuh)
// This class owns a byte-identical host mirror of those two arrays, mutates it with the reference's
// exact operation order, and records which slots / row words changed so the engine can scatter just
// those to the device.  A set that becomes empty keeps its slot with cnt == 0, which the probe treats
// as a miss — the same observable behaviour as `delete(hashToPods, hash)` (indexer.go:109-112).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "kernels.cuh"
#include "xxh64.cuh"

namespace eppscore {

// u64 -> u32 hash map, linear probing, backward-shift deletion (no tombstones).
class FlatMap64 {
 public:
  explicit FlatMap64(size_t initial = 64) { rehash(initial); }
  size_t size() const { return n_; }
  bool get(uint64_t k, uint32_t* v) const {
    for (size_t i = home(k);; i = (i + 1) & mask_) {
      if (!used_[i]) return false;
      if (keys_[i] == k) {
        *v = vals_[i];
        return true;
      }
    }
  }
  void put(uint64_t k, uint32_t v) {
    if ((n_ + 1) * 10 > (mask_ + 1) * 6) rehash((mask_ + 1) * 2);
    for (size_t i = home(k);; i = (i + 1) & mask_) {
      if (!used_[i]) {
        used_[i] = 1;
        keys_[i] = k;
        vals_[i] = v;
        n_++;
        return;
      }
      if (keys_[i] == k) {
        vals_[i] = v;
        return;
      }
    }
  }
  void erase(uint64_t k) {
    size_t i = home(k);
    for (;; i = (i + 1) & mask_) {
      if (!used_[i]) return;
      if (keys_[i] == k) break;
    }
    used_[i] = 0;
    n_--;
    for (size_t j = (i + 1) & mask_; used_[j]; j = (j + 1) & mask_) {
      const size_t h = home(keys_[j]);
      // slot j may fill the hole at i iff its home is NOT in the cyclic interval (i, j]
      const bool in_between = (i <= j) ? (h > i && h <= j) : (h > i || h <= j);
      if (!in_between) {
        used_[i] = 1;
        keys_[i] = keys_[j];
        vals_[i] = vals_[j];
        used_[j] = 0;
        i = j;
      }
    }
  }

  void prefetch(uint64_t k) const {
    const size_t h = home(k);
    __builtin_prefetch(&keys_[h]);
    __builtin_prefetch(&used_[h]);
  }

 private:
  static uint64_t scramble(uint64_t x) {
    x *= 0x9E3779B97F4A7C15ULL;
    return x ^ (x >> 29);
  }
  size_t home(uint64_t k) const { return (size_t)scramble(k) & mask_; }
  void rehash(size_t cap) {
    size_t c = 16;
    while (c < cap) c <<= 1;
    std::vector<uint64_t> ok;
    std::vector<uint32_t> ov;
    std::vector<uint8_t> ou;
    ok.swap(keys_);
    ov.swap(vals_);
    ou.swap(used_);
    keys_.assign(c, 0);
    vals_.assign(c, 0);
    used_.assign(c, 0);
    mask_ = c - 1;
    n_ = 0;
    for (size_t i = 0; i < ou.size(); i++)
      if (ou[i]) put(ok[i], ov[i]);
  }
  std::vector<uint64_t> keys_;
  std::vector<uint32_t> vals_;
  std::vector<uint8_t> used_;
  size_t mask_ = 0, n_ = 0;
};

// LRU set of block hashes for one endpoint. front = most recent, back = oldest.
class LruSet {
 public:
  explicit LruSet(int32_t capacity) : cap_(capacity < 1 ? 1 : capacity), map_(64) {}
  int32_t len() const { return len_; }
  int32_t capacity() const { return cap_; }
  // golang-lru Add: existing key -> move to front, no eviction; new key -> push front and, if the
  // length now exceeds the size, remove the oldest (returns true and the evicted key).
  bool add(uint64_t k, uint64_t* evicted) {
    uint32_t node;
    if (map_.get(k, &node)) {
      detach((int32_t)node);
      attach_front((int32_t)node);
      return false;
    }
    const int32_t nn = alloc();
    key_[nn] = k;
    attach_front(nn);
    map_.put(k, (uint32_t)nn);
    len_++;
    if (len_ <= cap_) return false;
    const int32_t old = back_;
    *evicted = key_[old];
    detach(old);
    map_.erase(key_[old]);
    next_[old] = free_;
    free_ = old;
    len_--;
    return true;
  }
  void prefetch(uint64_t k) const { map_.prefetch(k); }
  template <typename F>
  void for_each_oldest_first(F&& f) const {  // lru.Keys(): oldest -> newest
    for (int32_t n = back_; n >= 0; n = prev_[n]) f(key_[n]);
  }

 private:
  int32_t alloc() {
    if (free_ >= 0) {
      const int32_t n = free_;
      free_ = next_[n];
      return n;
    }
    key_.push_back(0);
    prev_.push_back(-1);
    next_.push_back(-1);
    return (int32_t)key_.size() - 1;
  }
  void detach(int32_t n) {
    const int32_t p = prev_[n], q = next_[n];
    (p >= 0 ? next_[p] : front_) = q;
    (q >= 0 ? prev_[q] : back_) = p;
  }
  void attach_front(int32_t n) {
    prev_[n] = -1;
    next_[n] = front_;
    if (front_ >= 0) prev_[front_] = n;
    front_ = n;
    if (back_ < 0) back_ = n;
  }
  int32_t cap_, len_ = 0, front_ = -1, back_ = -1, free_ = -1;
  std::vector<uint64_t> key_;
  std::vector<int32_t> prev_, next_;
  FlatMap64 map_;
};

class PrefixIndex {
 public:
  PrefixIndex(const Geo& geo, int64_t capacity_rows, int32_t default_lru)
      : geo_(geo), default_lru_(default_lru) {
    if (capacity_rows < 16) capacity_rows = 16;
    cap_rows_ = capacity_rows;
    uint64_t c = 32;
    while (c < (uint64_t)capacity_rows * 2) c <<= 1;
    slots_.assign(c, Slot{~0ULL, kEmptyRow, 0u});
    slot_mask_ = c - 1;
    slot_dirty_flag_.assign(c, 0);
    // row 0 is the permanent empty set: every new hash starts there (and returns there when emptied)
    rows_.assign((size_t)geo_.row_words, 0u);
    row_ref_.assign(1, 1u);
    row_hash_.assign(1, 0ULL);  // row 0 = the empty set
    zob_.resize((size_t)geo.row_words * 32);  // indexed by permuted bit position (perm_bitpos), which spans row_words*32
    for (size_t i = 0; i < zob_.size(); i++) {  // splitmix64: one fixed random word per endpoint bit
      uint64_t z = 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      zob_[i] = z ^ (z >> 31);
    }
    n_rows_ = 1;
    word_dirty_flag_.assign((rows_.size() + 7) / 8, 0);
  }

  const Geo& geo() const { return geo_; }
  const std::vector<Slot>& slots() const { return slots_; }
  const std::vector<uint32_t>& rows() const { return rows_; }
  uint64_t slot_mask() const { return slot_mask_; }
  int64_t n_rows() const { return n_rows_; }  // rows ever allocated (high-water mark; row 0 = the empty set)
  int64_t n_keys() const { return n_keys_; }  // distinct block hashes with a slot
  int64_t n_live() const { return n_live_; }
  int64_t capacity_rows() const { return cap_rows_; }
  int64_t lru_entries() const {
    int64_t t = 0;
    for (auto& l : lru_)
      if (l) t += l->len();
    return t;
  }

  // indexer.Add (indexer.go:52-83): ALL LRU adds first (evictions fire here), THEN the hashToPods
  // update for ALL hashes — kept in this order so the stale-entry quirk of over-long batches is identical.
  bool add(const uint64_t* hashes, int32_t n, int32_t endpoint, int32_t lru_capacity) {
    if (endpoint < 0 || endpoint >= geo_.Mpad) return false;
    LruSet* l = lru_for(endpoint, lru_capacity);
    // both phases are chains of dependent cache misses into big tables (the LRU's map, the key slots): touch the home
    // lines of all n hashes first so that the misses overlap
    for (int32_t i = 0; i < n; i++) {
      l->prefetch(hashes[i]);
      __builtin_prefetch(&slots_[hashes[i] & slot_mask_]);
    }
    for (int32_t i = 0; i < n; i++) {
      uint64_t ev;
      if (l->add(hashes[i], &ev)) clear_bit(ev, endpoint);  // makeEvictionFn, indexer.go:105-115
    }
    for (int32_t i = 0; i < n; i++)
      if (!set_bit(hashes[i], endpoint)) return false;
    return true;
  }
  // raw deltas for hosts that run their own LRU
  bool apply(uint64_t hash, int32_t endpoint, int op) {
    if (endpoint < 0 || endpoint >= geo_.Mpad) return false;
    if (op == 0) return set_bit(hash, endpoint);
    clear_bit(hash, endpoint);
    return true;
  }
  // indexer.RemovePod (indexer.go:167-182)
  void remove_endpoint(int32_t endpoint) {
    if (endpoint < 0 || endpoint >= (int32_t)lru_.size() || !lru_[endpoint]) return;
    lru_[endpoint]->for_each_oldest_first([&](uint64_t h) { clear_bit(h, endpoint); });
    lru_[endpoint].reset();
  }
  int32_t lru_len(int32_t endpoint) const {
    if (endpoint < 0 || endpoint >= (int32_t)lru_.size() || !lru_[endpoint]) return -1;
    return lru_[endpoint]->len();
  }
  int32_t lru_keys(int32_t endpoint, uint64_t* out, int32_t cap) const {
    if (endpoint < 0 || endpoint >= (int32_t)lru_.size() || !lru_[endpoint]) return -1;
    int32_t k = 0;
    lru_[endpoint]->for_each_oldest_first([&](uint64_t h) {
      if (k < cap) out[k] = h;
      k++;
    });
    return k;
  }
  // slot index of a hash in the mirror, or -1
  int64_t find(uint64_t h) const {
    for (uint64_t i = h & slot_mask_;; i = (i + 1) & slot_mask_) {
      if (slots_[i].row == kEmptyRow) return -1;
      if (slots_[i].key == h) return (int64_t)i;
    }
  }

  // ---- dirty tracking (what must be scattered to the device) ----
  bool full_upload_needed() const { return full_upload_; }
  void mark_full_upload() { full_upload_ = true; }
  const std::vector<uint32_t>& dirty_slots() const { return dirty_slots_; }
  const std::vector<uint32_t>& dirty_words() const { return dirty_words_; }
  void clear_dirty() {
    for (uint32_t i : dirty_slots_) slot_dirty_flag_[i] = 0;
    for (uint32_t i : dirty_words_) word_dirty_flag_[i >> 3] &= (uint8_t)~(1u << (i & 7));
    dirty_slots_.clear();
    dirty_words_.clear();
    full_upload_ = false;
  }
  // adopt an image that another engine built (replication across GPUs): mirror stays empty, read-only
  void adopt_counts(int64_t n_rows, int64_t n_live) {
    n_rows_ = n_rows;
    n_live_ = n_live;
  }

 private:
  LruSet* lru_for(int32_t endpoint, int32_t capacity) {
    if ((size_t)endpoint >= lru_.size()) lru_.resize((size_t)endpoint + 1);
    if (!lru_[endpoint]) {  // indexer.go:57-68: size = NumOfGPUBlocks, or the default when <= 0
      int32_t size = capacity > 0 ? capacity : default_lru_;
      lru_[endpoint] = std::make_unique<LruSet>(size);
    }
    return lru_[endpoint].get();
  }
  void touch_slot(uint64_t i) {
    if (!slot_dirty_flag_[i]) {
      slot_dirty_flag_[i] = 1;
      dirty_slots_.push_back((uint32_t)i);
    }
  }
  void touch_word(uint64_t wi) {
    uint8_t& f = word_dirty_flag_[wi >> 3];
    const uint8_t bit = (uint8_t)(1u << (wi & 7));
    if (!(f & bit)) {
      f |= bit;
      dirty_words_.push_back((uint32_t)wi);
    }
  }
  // Double the capacity: re-insert every slot into a table twice the size (rows keep their ids); the engine
  // notices the new capacity, reallocates the device buffers and uploads the whole mirror.
  bool grow() {
    if (((uint64_t)cap_rows_ * 2 + 1) * (uint64_t)geo_.row_words >= (1ULL << 32)) return false;  // 32-bit word indices
    std::vector<Slot> old;
    old.swap(slots_);
    const uint64_t c = (slot_mask_ + 1) * 2;
    slots_.assign(c, Slot{~0ULL, kEmptyRow, 0u});
    slot_mask_ = c - 1;
    slot_dirty_flag_.assign(c, 0);
    dirty_slots_.clear();
    for (const Slot& sl : old) {
      if (sl.row == kEmptyRow) continue;
      uint64_t i = sl.key & slot_mask_;
      while (slots_[i].row != kEmptyRow) i = (i + 1) & slot_mask_;
      slots_[i] = sl;
    }
    cap_rows_ *= 2;
    full_upload_ = true;
    return true;
  }

  // ---- row interning: hashes whose endpoint SETS are identical share one bitset row, so the pick kernel can
  // run-length merge a request's consecutive matched blocks by row id and read each distinct set once.
  // (Consecutive blocks of a prompt are normally cached on exactly the same endpoints.)
  // The content hash of a set is the XOR of one fixed random word per member (Zobrist): adding or removing an endpoint
  // updates it in O(1), and every row remembers the hash of its content.  Equal hashes are confirmed with a memcmp.
  bool same_content(uint32_t row, const uint32_t* w) const {
    return std::memcmp(&rows_[(size_t)row * geo_.row_words], w, (size_t)geo_.row_words * 4) == 0;
  }
  void release_row(uint32_t row) {
    if (row == 0) return;
    if (--row_ref_[row] == 0) {
      const uint64_t ch = row_hash_[row];
      uint32_t r;
      if (intern_.get(ch, &r) && r == row) intern_.erase(ch);
      free_rows_.push_back(row);
    }
  }
  // point slot i at a row holding `content` (the slot's set just changed by one endpoint: bit `pos`)
  bool assign_content(uint64_t i, const uint32_t* content, uint32_t pos, bool now_empty) {
    const uint32_t old = slots_[i].row;
    const int RW = geo_.row_words;
    if (now_empty) {
      slots_[i].row = 0;
      release_row(old);
      return true;
    }
    const uint64_t ch = row_hash_[old] ^ zob_[pos];  // the set changed by exactly the endpoint at bit `pos`
    uint32_t r;
    if (intern_.get(ch, &r) && same_content(r, content)) {  // an identical set already has a row: share it
      row_ref_[r]++;
      slots_[i].row = r;
      release_row(old);
      return true;
    }
    const bool clash = intern_.get(ch, &r);  // (64-bit content-hash collision: keep this row un-interned)
    if (old != 0 && row_ref_[old] == 1) {    // sole owner: mutate in place, only one word changes
      const uint64_t och = row_hash_[old];
      uint32_t t;
      if (intern_.get(och, &t) && t == old) intern_.erase(och);
      row_hash_[old] = ch;
      const uint64_t wi = (uint64_t)old * RW + (pos >> 5);
      rows_[wi] = content[pos >> 5];
      touch_word(wi);
      if (!clash) intern_.put(ch, old);
      return true;
    }
    uint32_t nr;
    if (!free_rows_.empty()) {
      nr = free_rows_.back();
      free_rows_.pop_back();
    } else {
      if (n_rows_ > cap_rows_) return false;  // live rows <= live hashes <= capacity (+ row 0): cannot happen
      nr = (uint32_t)n_rows_++;
      rows_.resize((size_t)n_rows_ * RW, 0u);
      row_ref_.resize((size_t)n_rows_, 0u);
      row_hash_.resize((size_t)n_rows_, 0ULL);
      word_dirty_flag_.resize((rows_.size() + 7) / 8, 0);
    }
    row_ref_[nr] = 1;
    row_hash_[nr] = ch;
    for (int w = 0; w < RW; w++) {
      const uint64_t wi = (uint64_t)nr * RW + w;
      if (rows_[wi] != content[w]) {
        rows_[wi] = content[w];
        touch_word(wi);
      }
    }
    if (!clash) intern_.put(ch, nr);
    slots_[i].row = nr;
    release_row(old);
    return true;
  }
  bool set_bit(uint64_t h, int32_t endpoint) {
    uint64_t i = h & slot_mask_;
    for (;; i = (i + 1) & slot_mask_) {
      if (slots_[i].row == kEmptyRow) {  // new hash: claim the slot, start from the empty set
        if (n_keys_ >= cap_rows_) {
          if (!grow()) return false;
          return set_bit(h, endpoint);  // re-probe in the doubled table
        }
        slots_[i].key = h;
        slots_[i].row = 0;
        slots_[i].cnt = 0;
        n_keys_++;
        break;
      }
      if (slots_[i].key == h) break;
    }
    const int RW = geo_.row_words;
    const uint32_t pos = perm_bitpos((uint32_t)endpoint, geo_.log_epl);
    const uint32_t bit = 1u << (pos & 31);
    const uint32_t* cur = &rows_[(size_t)slots_[i].row * RW];
    if (!(cur[pos >> 5] & bit)) {
      uint32_t tmp[256];
      std::memcpy(tmp, cur, (size_t)RW * 4);
      tmp[pos >> 5] |= bit;
      if (!assign_content(i, tmp, pos, false)) return false;
      if (slots_[i].cnt++ == 0) n_live_++;
    }
    touch_slot(i);
    return true;
  }
  void clear_bit(uint64_t h, int32_t endpoint) {
    const int64_t i = find(h);
    if (i < 0) return;
    const int RW = geo_.row_words;
    const uint32_t pos = perm_bitpos((uint32_t)endpoint, geo_.log_epl);
    const uint32_t bit = 1u << (pos & 31);
    const uint32_t* cur = &rows_[(size_t)slots_[i].row * RW];
    if (cur[pos >> 5] & bit) {
      uint32_t tmp[256];
      std::memcpy(tmp, cur, (size_t)RW * 4);
      tmp[pos >> 5] &= ~bit;
      const bool now_empty = slots_[i].cnt == 1;
      assign_content((uint64_t)i, tmp, pos, now_empty);  // cannot fail: the row pool holds capacity + 1 rows
      if (--slots_[i].cnt == 0) n_live_--;
      touch_slot((uint64_t)i);
    }
  }

  Geo geo_;
  int32_t default_lru_;
  int64_t cap_rows_ = 0, n_rows_ = 0, n_keys_ = 0, n_live_ = 0;
  uint64_t slot_mask_ = 0;
  std::vector<Slot> slots_;
  std::vector<uint32_t> rows_;
  std::vector<uint32_t> row_ref_;    // slots pointing at each row
  std::vector<uint64_t> row_hash_;   // Zobrist hash of each row's content
  std::vector<uint64_t> zob_;        // per endpoint bit position
  std::vector<uint32_t> free_rows_;
  FlatMap64 intern_{1024};           // content hash -> row id
  std::vector<std::unique_ptr<LruSet>> lru_;
  std::vector<uint32_t> dirty_slots_, dirty_words_;
  std::vector<uint8_t> slot_dirty_flag_, word_dirty_flag_;
  bool full_upload_ = true;
};

}  // namespace eppscore
