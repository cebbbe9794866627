// Test-only C shim around the HOST part of the product's prefix index
// (gateway-api-inference-extension_b200/csrc/prefix_index.hpp) so the CPU test suite can drive it with
// random operation sequences and compare every observable (sets, LRU order, counters) with the oracle's
// indexer — no GPU needed: this class is plain C++ (the device only ever sees its mirror arrays).
#include "../../gateway-api-inference-extension_b200/csrc/prefix_index.hpp"

using namespace eppscore;

extern "C" {
void* pit_new(int max_endpoints, long long capacity, int default_lru) {
  return new PrefixIndex(make_geo(max_endpoints), capacity, default_lru);
}
void pit_free(void* p) { delete static_cast<PrefixIndex*>(p); }
int pit_add(void* p, const uint64_t* h, int n, int ep, int cap) { return static_cast<PrefixIndex*>(p)->add(h, n, ep, cap) ? 0 : -1; }
int pit_apply(void* p, uint64_t h, int ep, int op) { return static_cast<PrefixIndex*>(p)->apply(h, ep, op) ? 0 : -1; }
void pit_remove_endpoint(void* p, int ep) { static_cast<PrefixIndex*>(p)->remove_endpoint(ep); }
int pit_lru_len(void* p, int ep) { return static_cast<PrefixIndex*>(p)->lru_len(ep); }
int pit_lru_keys(void* p, int ep, uint64_t* out, int cap) { return static_cast<PrefixIndex*>(p)->lru_keys(ep, out, cap); }
// the endpoint set of a hash, read from the MIRROR arrays exactly as the device probe would (slot -> row), un-permuted
int pit_get(void* p, uint64_t h, int32_t* eps_out, int cap) {
  auto* ix = static_cast<PrefixIndex*>(p);
  const int64_t s = ix->find(h);
  if (s < 0) return 0;
  const Slot& sl = ix->slots()[(size_t)s];
  if (sl.cnt == 0) return 0;
  const Geo& g = ix->geo();
  int n = 0;
  for (int m = 0; m < g.Mpad; m++) {
    const uint32_t pos = perm_bitpos((uint32_t)m, g.log_epl);
    if ((ix->rows()[(size_t)sl.row * g.row_words + (pos >> 5)] >> (pos & 31)) & 1u) {
      if (n < cap) eps_out[n] = m;
      n++;
    }
  }
  return n == (int)sl.cnt ? n : -1000 - n;  // the slot's count must equal the popcount of its row
}
long long pit_n_live(void* p) { return static_cast<PrefixIndex*>(p)->n_live(); }
long long pit_n_keys(void* p) { return static_cast<PrefixIndex*>(p)->n_keys(); }
long long pit_n_rows(void* p) { return static_cast<PrefixIndex*>(p)->n_rows(); }
long long pit_lru_entries(void* p) { return static_cast<PrefixIndex*>(p)->lru_entries(); }
long long pit_dirty(void* p) { auto* ix = static_cast<PrefixIndex*>(p); return (long long)(ix->dirty_slots().size() + ix->dirty_words().size()); }
void pit_clear_dirty(void* p) { static_cast<PrefixIndex*>(p)->clear_dirty(); }
}
