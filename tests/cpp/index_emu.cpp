// Test-only: the CTA program of the device-resident prefix index (csrc/prefix_table.cuh) instantiated with a SEQUENTIAL
// execution policy, so the commit / evict / rebuild / remove logic is unit-tested on the CPU against the oracle's indexer
// (tests/test_device_index_emu.py).  "Threads" of a parallel section run one after the other; atomics are plain.
// What this cannot show is a data race — the GPU tests and compute-sanitizer cover that.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gateway-api-inference-extension_b200/csrc/prefix_table.cuh"

using namespace eppscore;

struct EmuX {
  template <class F>
  void par(F f) {
    for (int t = 0; t < kCommitThreads; t++) f(t);
  }
  uint32_t scan(uint32_t* arr, uint32_t*) {
    uint32_t run = 0;
    for (int t = 0; t < kCommitThreads; t++) {
      const uint32_t v = arr[t];
      arr[t] = run;
      run += v;
    }
    return run;
  }
  uint32_t ld32(const uint32_t* p) { return *p; }
  uint64_t ld64(const uint64_t* p) { return *p; }
  uint32_t ld16(const uint16_t* p) { return *p; }
  void st32(uint32_t* p, uint32_t v) { *p = v; }
  void st64(uint64_t* p, uint64_t v) { *p = v; }
  void st16(uint16_t* p, uint16_t v) { *p = v; }
  uint32_t ld_acquire32(const uint32_t* p) { return *p; }
  void st_release32(uint32_t* p, uint32_t v) { *p = v; }
  uint32_t cas32(uint32_t* p, uint32_t c, uint32_t v) {
    const uint32_t o = *p;
    if (o == c) *p = v;
    return o;
  }
  uint32_t cas_acquire32(uint32_t* p, uint32_t c, uint32_t v) { return cas32(p, c, v); }
  uint64_t cas64(uint64_t* p, uint64_t c, uint64_t v) {
    const uint64_t o = *p;
    if (o == c) *p = v;
    return o;
  }
  void add64(unsigned long long* p, unsigned long long v) { *p += v; }
  void fence() {}
  void lock(uint32_t* p) { *p = 1; }
  void unlock(uint32_t* p) { *p = 0; }
  uint64_t smem_cas64(uint64_t* p, uint64_t c, uint64_t v) { return cas64(p, c, v); }
  void smem_max32(uint32_t* p, uint32_t v) {
    if (v > *p) *p = v;
  }
  void smem_add32(uint32_t* p, uint32_t v) { *p += v; }
};

struct Emu {
  TableView tv{};
  LruView lv{};
  std::vector<TSlot> slots;
  std::vector<uint32_t> ovf_rows, ovf_free;
  std::vector<LruDesc> desc;
  std::vector<LruEntry> maps;
  std::vector<uint64_t> logs;
  CommitSmem* sm = nullptr;
  int64_t rebuilds = 0;
  EmuX x;

  void set_slots(uint64_t cap) {
    TSlot f;
    memset(&f, 0xFF, sizeof(f));
    slots.assign(cap, f);
    tv.slots = slots.data();
    tv.mask = cap - 1;
  }
  // mirrors DeviceIndex::ensure_room (prefix_index.cu)
  void ensure_room(int64_t touches) {
    const uint64_t cap = tv.mask + 1;
    if ((tv.used + (uint64_t)touches) * 2 > cap) {
      uint64_t ncap = cap;
      while ((tv.live + (uint64_t)touches) * 4 > ncap) ncap <<= 1;
      std::vector<TSlot> old;
      old.swap(slots);
      set_slots(ncap);
      tv.used = tv.live = 0;
      for (const TSlot& s : old) {
        if (s.cnt == kCntFree || (s.cnt & kCntMask) == 0) continue;
        uint64_t j = s.key & tv.mask;
        while (slots[j].cnt != kCntFree) j = (j + 1) & tv.mask;
        slots[j] = s;
        tv.used++;
        tv.live++;
      }
      rebuilds++;
    }
    const uint64_t in_use = (uint64_t)tv.ovf_next - tv.ovf_free_top;
    if (in_use + (uint64_t)touches > tv.ovf_cap) {
      const uint64_t ncap = (in_use + (uint64_t)touches) * 5 / 4 + 16;
      ovf_rows.resize(ncap * tv.row_words);
      ovf_free.resize(ncap);
      tv.ovf_rows = ovf_rows.data();
      tv.ovf_free = ovf_free.data();
      tv.ovf_cap = (uint32_t)ncap;
    }
  }
};

extern "C" {
void* emu_new(int n_endpoints, long long capacity, int default_lru, int lru_max) {
  Emu* e = new Emu();
  uint64_t c = 32;
  while (c < (uint64_t)(capacity < 16 ? 16 : capacity) * 2) c <<= 1;
  e->set_slots(c);
  e->tv.row_words = (uint32_t)(n_endpoints + 31) / 32;
  const uint32_t maxcap = (uint32_t)(default_lru > lru_max ? default_lru : lru_max);
  e->lv.map_size = lru_map_size_for(maxcap);
  e->lv.log_size = lru_log_size_for(maxcap);
  e->lv.default_cap = (uint32_t)(default_lru < 1 ? 1 : default_lru);
  e->lv.max_cap = maxcap;
  e->lv.n_endpoints = (uint32_t)n_endpoints;
  e->desc.assign((size_t)n_endpoints, LruDesc{});
  LruEntry fe;
  memset(&fe, 0xFF, sizeof(fe));
  e->maps.assign((size_t)n_endpoints * e->lv.map_size, fe);
  e->logs.assign((size_t)n_endpoints * e->lv.log_size, 0);
  e->lv.desc = e->desc.data();
  e->lv.maps = e->maps.data();
  e->lv.logs = e->logs.data();
  e->sm = new CommitSmem();
  return e;
}
void emu_free(void* h) {
  Emu* e = static_cast<Emu*>(h);
  delete e->sm;
  delete e;
}
// PreRequest for a batch: every endpoint's "CTA" runs in turn (they commute)
int emu_commit(void* h, int R, const int32_t* pick, const uint64_t* hashes, const uint16_t* nh, int stride, const int32_t* cap_req,
               int single_cap) {
  Emu* e = static_cast<Emu*>(h);
  int64_t touches = 0;
  for (int r = 0; r < R; r++) touches += pick[r] >= 0 ? nh[r] : 0;
  e->ensure_room(touches);
  CommitArgs a{};
  a.R = R;
  a.pick = pick;
  a.hashes = hashes;
  a.n_hashes = nh;
  a.stride = stride;
  a.cap_req = cap_req;
  a.single_cap = single_cap;
  std::vector<char> seen(e->lv.n_endpoints, 0);
  for (int r = 0; r < R; r++)
    if (pick[r] >= 0 && (uint32_t)pick[r] < e->lv.n_endpoints) seen[pick[r]] = 1;
  for (uint32_t p = 0; p < e->lv.n_endpoints; p++) {
    if (!seen[p]) continue;  // an endpoint without requests is a no-op in the kernel as well
    IndexProgram<EmuX> prog(e->x, &e->tv, &e->lv, e->sm, p);
    prog.commit(a);
  }
  return (int)(e->tv.error | (e->lv.error << 8));
}
int emu_apply(void* h, uint64_t hash, int ep, int op) {
  Emu* e = static_cast<Emu*>(h);
  e->ensure_room(1);
  if (op == 0) TableOps<EmuX>::set_member(e->x, &e->tv, hash, (uint32_t)ep);
  else TableOps<EmuX>::clear_member(e->x, &e->tv, hash, (uint32_t)ep);
  return (int)e->tv.error;
}
void emu_remove_endpoint(void* h, int ep) {
  Emu* e = static_cast<Emu*>(h);
  IndexProgram<EmuX> prog(e->x, &e->tv, &e->lv, e->sm, (uint32_t)ep);
  prog.remove_endpoint();
}
int emu_lru_keys(void* h, int ep, uint64_t* out, int cap) {
  Emu* e = static_cast<Emu*>(h);
  if (!e->desc[ep].created) return -1;
  IndexProgram<EmuX> prog(e->x, &e->tv, &e->lv, e->sm, (uint32_t)ep);
  return (int)prog.export_keys(out, (uint32_t)cap);
}
int emu_lru_len(void* h, int ep) {
  Emu* e = static_cast<Emu*>(h);
  return e->desc[ep].created ? (int)e->desc[ep].live : -1;
}
int emu_get(void* h, uint64_t hash, int32_t* eps_out, int cap) {
  Emu* e = static_cast<Emu*>(h);
  std::vector<uint32_t> bits(e->tv.row_words);
  const uint32_t c = table_get(e->x, &e->tv, hash, bits.data(), e->tv.row_words);
  int n = 0;
  for (uint32_t m = 0; m < e->tv.row_words * 32; m++)
    if ((bits[m >> 5] >> (m & 31)) & 1u) {
      if (n < cap) eps_out[n] = (int32_t)m;
      n++;
    }
  return n == (int)c ? n : -1000 - n;  // the slot's count must equal the size of its set
}
long long emu_stat(void* h, int which) {
  Emu* e = static_cast<Emu*>(h);
  switch (which) {
    case 0: return (long long)e->tv.live;
    case 1: return (long long)e->tv.used;
    case 2: return (long long)(e->tv.mask + 1);
    case 3: return (long long)e->tv.ovf_next - (long long)e->tv.ovf_free_top;
    case 4: return e->rebuilds;
    case 5: {
      long long t = 0;
      for (auto& d : e->desc) t += d.created ? d.live : 0;
      return t;
    }
    case 6: {
      long long t = 0;
      for (auto& d : e->desc) t += (long long)(d.head - d.tail);
      return t;
    }
    case 7: {  // slots that are NOT canonical: an inline set must be ascending with every unused entry 0xFFFF (the scoring
               // kernel compares equal sets as words: pick_sparse.cu same_inline_set); a row slot's spare entries are free
      long long bad = 0;
      for (uint64_t i = 0; i <= e->tv.mask; i++) {
        const TSlot& s = e->tv.slots[i];
        if (s.cnt == kCntFree || (s.cnt & kCntRow)) continue;
        const uint32_t c = s.cnt & kCntMask;
        bool ok = c <= (uint32_t)kInlineEps;
        for (uint32_t k = 0; ok && k < (uint32_t)kInlineEps; k++) {
          if (k < c) ok = s.ep[k] != 0xFFFFu && (k == 0 || s.ep[k] > s.ep[k - 1]);
          else ok = s.ep[k] == 0xFFFFu;
        }
        bad += ok ? 0 : 1;
      }
      return bad;
    }
    default: return -1;
  }
}
}
