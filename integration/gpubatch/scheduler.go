// Package gpubatch is the reference-side binding for libeppscore.so: a requestcontrol.Scheduler
// (pkg/epp/requestcontrol/director.go:68-70) that coalesces the Director's concurrent Schedule() calls into engine batches
// (the pattern of sidecars/latencypredictorasync/coalescer.go:53-120: a window armed by the first submission, early
// dispatch at the row cap) and runs Filter -> Score -> Pick for the whole batch in one eppscore_schedule_batch.
//
// STATUS: written against include/eppscore.h (ABI v3) and the reference at c4c8fef; NOT COMPILED in the build image of
// this repository (no Go toolchain there).  The same front, packer and result hand-back exist compiled and tested in C++
// (gateway-api-inference-extension_b200/host/{coalescer,epp_scheduler,host_eval}.hpp, host_test.cpp) — this file is their
// transcription for the maintainer who adds the plugin to pkg/epp/scheduling/gpubatch/.
//
// Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<dir of libeppscore.so> -leppscore".  Needs Go >= 1.21 (runtime.Pinner).
package gpubatch

/*
#cgo LDFLAGS: -leppscore
#include <stdlib.h>
#include "eppscore.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"runtime"
	"strconv"
	"strings"
	"sync"
	"time"
	"unsafe"

	"k8s.io/apimachinery/pkg/types"

	fwksched "sigs.k8s.io/gateway-api-inference-extension/pkg/epp/framework/interface/scheduling"
)

// ScorerSpec is one `pluginRef` + `weight` of the scheduling profile, in YAML order (config/loader/defaults.go:46-103).
type ScorerSpec struct {
	Type   string // "queue-scorer" | "kv-cache-utilization-scorer" | "prefix-cache-scorer" | "lora-affinity-scorer" | "running-requests-size-scorer" | "token-load-scorer" | "latency-scorer"
	Weight float64
}

type Config struct {
	Device          int
	Scorers         []ScorerSpec
	Picker          string        // "max-score-picker" (default) | "weighted-random-picker" | "random-picker"
	BlockSizeTokens int           // approximateprefix types.go:91 (default 16)
	MaxPrefixBlocks int           // types.go:98 (default 256)
	LRUCapacity     int           // types.go:109 (default 31250)
	MaxEndpoints    int           // engine capacity for stable endpoint ids
	MaxAdapters     int           // engine capacity for the adapter dictionary
	Window          time.Duration // e.g. 200 * time.Microsecond (INTEGRATION.md: measured latency vs window)
	MaxBatch        int           // e.g. 4096
	// UserInputBytes is approximateprefix's getUserInputBytes (hashing.go:106-135); it is unexported there, so the maintainer
	// either exports it or passes the same function here.
	UserInputBytes func(*fwksched.InferenceRequest) ([]byte, error)
	CacheSalt      func(*fwksched.InferenceRequest) string // hashing.go:70-77 (request.Body ... CacheSalt), "" when absent
}

const profileName = "default" // profile/single_profile_handler.go:66-99: exactly one profile

type call struct {
	req  *fwksched.InferenceRequest
	eps  []fwksched.Endpoint
	res  *fwksched.SchedulingResult
	err  error
	done chan struct{}
}

type Scheduler struct {
	cfg     Config
	eng     *C.eppscore_engine
	pending chan *call

	mu       sync.Mutex                     // guards the id maps (flush runs on one goroutine; RemovePod may come from the reconciler)
	serverID map[types.NamespacedName]int32 // ServerID (indexer.go:34-35) -> stable engine endpoint id
	freeIDs  []int32
	adapter  map[string]int32 // adapter / model name -> dictionary id (bit position in lora_active / lora_waiting)
}

func kindOf(t string) (C.int32_t, error) {
	switch t {
	case "queue-scorer":
		return C.EPPSCORE_SCORER_QUEUE, nil
	case "kv-cache-utilization-scorer":
		return C.EPPSCORE_SCORER_KV_CACHE, nil
	case "prefix-cache-scorer":
		return C.EPPSCORE_SCORER_PREFIX, nil
	case "lora-affinity-scorer":
		return C.EPPSCORE_SCORER_LORA, nil
	case "running-requests-size-scorer":
		return C.EPPSCORE_SCORER_RUNNING, nil
	case "latency-scorer":
		return C.EPPSCORE_SCORER_LATENCY, nil
	case "token-load-scorer":
		return C.EPPSCORE_SCORER_TOKEN_LOAD, nil
	}
	return 0, fmt.Errorf("gpubatch: scorer %q has no device form (keep it a Go plugin and feed it as an endpoint column)", t)
}

func New(cfg Config) (*Scheduler, error) {
	var c C.eppscore_config
	C.eppscore_config_default(&c)
	if len(cfg.Scorers) > C.EPPSCORE_MAX_SCORERS {
		return nil, errors.New("gpubatch: too many scorers in the profile")
	}
	c.n_scorers = C.int32_t(len(cfg.Scorers))
	for i, s := range cfg.Scorers { // profile order matters: float64 adds are not associative (scheduler_profile.go:151-174)
		k, err := kindOf(s.Type)
		if err != nil {
			return nil, err
		}
		c.scorer_kind[i], c.scorer_weight[i] = k, C.double(s.Weight)
	}
	if cfg.BlockSizeTokens > 0 {
		c.block_chars = C.int32_t(cfg.BlockSizeTokens * 4) // averageCharactersPerToken, types.go:112
	}
	if cfg.MaxPrefixBlocks > 0 {
		c.max_blocks = C.int32_t(cfg.MaxPrefixBlocks)
	}
	if cfg.LRUCapacity > 0 {
		c.lru_capacity_default = C.int32_t(cfg.LRUCapacity)
	}
	if cfg.MaxEndpoints > 0 {
		c.max_endpoints = C.int32_t(cfg.MaxEndpoints)
	}
	if cfg.MaxAdapters > 0 {
		c.max_adapters = C.int32_t(cfg.MaxAdapters)
	}
	switch cfg.Picker {
	case "", "max-score-picker":
		c.pick_mode = C.EPPSCORE_PICK_MAX_SCORE
		c.tie_mode = C.EPPSCORE_TIE_SEEDED_RANDOM // the reference shuffles ties (picker/common.go:49-55)
		c.tie_seed = C.uint64_t(time.Now().UnixNano())
	case "weighted-random-picker":
		c.pick_mode = C.EPPSCORE_PICK_WEIGHTED_RANDOM
		c.tie_seed = C.uint64_t(time.Now().UnixNano())
	case "random-picker":
		c.pick_mode = C.EPPSCORE_PICK_RANDOM
		c.tie_seed = C.uint64_t(time.Now().UnixNano())
	default:
		return nil, fmt.Errorf("gpubatch: unknown picker %q", cfg.Picker)
	}
	var e *C.eppscore_engine
	if rc := C.eppscore_create(C.int32_t(cfg.Device), &c, &e); rc != C.EPPSCORE_OK {
		return nil, fmt.Errorf("eppscore_create: %s", C.GoString(C.eppscore_last_error(nil)))
	}
	if cfg.Window <= 0 {
		cfg.Window = 200 * time.Microsecond
	}
	if cfg.MaxBatch <= 0 {
		cfg.MaxBatch = 4096
	}
	s := &Scheduler{cfg: cfg, eng: e, pending: make(chan *call, 4*cfg.MaxBatch),
		serverID: map[types.NamespacedName]int32{}, adapter: map[string]int32{}}
	go s.dispatch()
	return s, nil
}

func (s *Scheduler) Close() { close(s.pending) }

// Schedule implements requestcontrol.Scheduler.  Called from one goroutine per request (handlers/server.go:162); it parks
// the request and waits for its batch.
func (s *Scheduler) Schedule(ctx context.Context, req *fwksched.InferenceRequest, eps []fwksched.Endpoint) (*fwksched.SchedulingResult, error) {
	c := &call{req: req, eps: eps, done: make(chan struct{})}
	select {
	case s.pending <- c:
	case <-ctx.Done():
		return nil, ctx.Err()
	}
	select {
	case <-c.done:
		return c.res, c.err
	case <-ctx.Done(): // the batch still runs; its result for this request is dropped
		return nil, ctx.Err()
	}
}

// RemovePod mirrors indexer.RemovePod (indexer.go:167-182): the pod reconciler calls it when a model server goes away.
func (s *Scheduler) RemovePod(name types.NamespacedName) {
	s.mu.Lock()
	defer s.mu.Unlock()
	if id, ok := s.serverID[name]; ok {
		C.eppscore_prefix_remove_endpoint(s.eng, C.int32_t(id))
		delete(s.serverID, name)
		s.freeIDs = append(s.freeIDs, id)
	}
}

// dispatch: the coalescer.  The first submission arms the window; the batch leaves when the window ends or it is full.
func (s *Scheduler) dispatch() {
	runtime.LockOSThread() // one OS thread talks to the CUDA context
	for first := range s.pending {
		batch := []*call{first}
		timer := time.NewTimer(s.cfg.Window)
	collect:
		for len(batch) < s.cfg.MaxBatch {
			select {
			case c, ok := <-s.pending:
				if !ok {
					break collect
				}
				batch = append(batch, c)
			case <-timer.C:
				break collect
			}
		}
		timer.Stop()
		s.flush(batch)
	}
	C.eppscore_destroy(s.eng)
}

func (s *Scheduler) idOf(name types.NamespacedName) (int32, error) {
	if id, ok := s.serverID[name]; ok {
		return id, nil
	}
	var id int32
	if n := len(s.freeIDs); n > 0 {
		id, s.freeIDs = s.freeIDs[n-1], s.freeIDs[:n-1]
	} else {
		id = int32(len(s.serverID))
	}
	if s.cfg.MaxEndpoints > 0 && int(id) >= s.cfg.MaxEndpoints {
		return 0, errors.New("gpubatch: more endpoints than Config.MaxEndpoints")
	}
	s.serverID[name] = id
	return id, nil
}

func (s *Scheduler) adapterID(name string) int32 {
	if id, ok := s.adapter[name]; ok {
		return id
	}
	if s.cfg.MaxAdapters > 0 && len(s.adapter) >= s.cfg.MaxAdapters {
		return -1 // dictionary full: the adapter scores as "not resident anywhere", like an unknown model
	}
	id := int32(len(s.adapter))
	s.adapter[name] = id
	return id
}

func fail(batch []*call, err error) {
	for _, c := range batch {
		c.err = err
		close(c.done)
	}
}

// flush: one engine batch.  The endpoints of all calls are merged into one snapshot over STABLE ids (the prefix index refers
// to endpoints by id, and PodList's order is not stable); each request's own candidate list becomes its cand_mask row.
func (s *Scheduler) flush(batch []*call) {
	s.mu.Lock()
	defer s.mu.Unlock()
	R := len(batch)

	// ---- 1. snapshot: the Metrics of every endpoint seen in this batch, SoA by stable id ----
	byID := map[int32]fwksched.Endpoint{}
	ids := make([][]int32, R) // per request: the ids of its candidates
	M := int32(0)
	for r, c := range batch {
		ids[r] = make([]int32, len(c.eps))
		for i, ep := range c.eps {
			id, err := s.idOf(ep.GetMetadata().NamespacedName)
			if err != nil {
				fail(batch, err)
				return
			}
			ids[r][i] = id
			byID[id] = ep
			if id+1 > M {
				M = id + 1
			}
		}
	}
	if M == 0 { // scheduler_profile.go:119-121 -> single_profile_handler.go:89-91
		fail(batch, fmt.Errorf("failed to run scheduler profile '%s'", profileName))
		return
	}
	for _, ep := range byID { // adapter dictionary first: its size fixes lora_words
		m := ep.GetMetrics()
		for name := range m.ActiveModels {
			s.adapterID(name)
		}
		for name := range m.WaitingModels {
			s.adapterID(name)
		}
	}
	words := (len(s.adapter) + 63) / 64
	if words == 0 {
		words = 1
	}
	kv := make([]C.double, M)
	queue := make([]C.int64_t, M)
	running := make([]C.int64_t, M)
	act := make([]C.uint64_t, int(M)*words)
	wait := make([]C.uint64_t, int(M)*words)
	nmodels := make([]C.int32_t, M)
	maxAct := make([]C.int32_t, M)
	lruCap := make([]C.int32_t, M) // CacheNumBlocks per endpoint (autotune, plugin.go:207-216); 0 => default
	for id, ep := range byID {
		m := ep.GetMetrics()
		kv[id] = C.double(m.KVCacheUsagePercent)
		queue[id] = C.int64_t(m.WaitingQueueSize)
		running[id] = C.int64_t(m.RunningRequestsSize)
		nmodels[id] = C.int32_t(len(m.ActiveModels) + len(m.WaitingModels)) // map sizes, lora_affinity.go:90
		maxAct[id] = C.int32_t(m.MaxActiveModels)
		lruCap[id] = C.int32_t(m.CacheNumBlocks)
		for name := range m.ActiveModels {
			if a := s.adapterID(name); a >= 0 {
				act[int(id)*words+int(a)/64] |= 1 << (uint(a) % 64)
			}
		}
		for name := range m.WaitingModels {
			if a := s.adapterID(name); a >= 0 {
				wait[int(id)*words+int(a)/64] |= 1 << (uint(a) % 64)
			}
		}
	}
	var pin runtime.Pinner // the C structs below hold pointers into Go slices: pin them for the duration of the calls
	defer pin.Unpin()
	var snap C.eppscore_snapshot
	snap.struct_size = C.uint32_t(unsafe.Sizeof(snap))
	snap.M, snap.lora_words = C.int32_t(M), C.int32_t(words)
	pin.Pin(&kv[0]); pin.Pin(&queue[0]); pin.Pin(&running[0]); pin.Pin(&act[0]); pin.Pin(&wait[0]); pin.Pin(&nmodels[0]); pin.Pin(&maxAct[0])
	snap.kv_usage, snap.queue, snap.running = &kv[0], &queue[0], &running[0]
	snap.lora_active, snap.lora_waiting = &act[0], &wait[0]
	snap.lora_nmodels, snap.lora_max = &nmodels[0], &maxAct[0]
	if rc := C.eppscore_set_snapshot(s.eng, &snap); rc != C.EPPSCORE_OK {
		fail(batch, fmt.Errorf("eppscore_set_snapshot: %s", C.GoString(C.eppscore_last_error(s.eng))))
		return
	}

	// ---- 2. per request: prompt bytes (16-byte aligned starts), model seed, adapter id, candidate mask, SLO headers ----
	mw := (int(M) + 31) / 32
	mask := make([]C.uint32_t, R*mw)
	off := make([]C.int64_t, R+1)
	plen := make([]C.int32_t, R)
	seed := make([]C.uint64_t, R)
	adapter := make([]C.int32_t, R)
	ttft := make([]C.double, R)
	tpot := make([]C.double, R)
	var bytes []byte
	for r, c := range batch {
		for _, id := range ids[r] {
			mask[r*mw+int(id)/32] |= 1 << (uint(id) % 32)
		}
		p, err := s.cfg.UserInputBytes(c.req)
		if err != nil {
			p = nil // hashing.go:40-44: no hashes, the prefix scorer scores 0 everywhere
		}
		for len(bytes)%16 != 0 {
			bytes = append(bytes, 0)
		}
		off[r], plen[r] = C.int64_t(len(bytes)), C.int32_t(len(p))
		bytes = append(bytes, p...)
		model, salt := c.req.TargetModel, ""
		if s.cfg.CacheSalt != nil {
			salt = s.cfg.CacheSalt(c.req)
		}
		seed[r] = C.eppscore_model_seed(unsafe.Pointer(unsafe.StringData(model)), C.size_t(len(model)),
			unsafe.Pointer(unsafe.StringData(salt)), C.size_t(len(salt)))
		if a, ok := s.adapter[c.req.TargetModel]; ok {
			adapter[r] = C.int32_t(a)
		} else {
			adapter[r] = -1
		}
		if v, err := strconv.ParseFloat(strings.TrimSpace(c.req.Headers["x-slo-ttft-ms"]), 64); err == nil {
			ttft[r] = C.double(v) // predictedlatency/plugin.go:330-343; a parse error leaves 0
		}
		if v, err := strconv.ParseFloat(strings.TrimSpace(c.req.Headers["x-slo-tpot-ms"]), 64); err == nil {
			tpot[r] = C.double(v)
		}
	}
	off[R] = C.int64_t(len(bytes))
	bytes = append(bytes, make([]byte, 64)...) // the hash kernel reads whole 16-byte words

	maxBlocks := s.cfg.MaxPrefixBlocks
	if maxBlocks <= 0 {
		maxBlocks = 256
	}
	pick := make([]C.int32_t, R)
	score := make([]C.double, R)
	ties := make([]C.int32_t, R)
	total := make([]C.uint16_t, R)
	hashes := make([]C.uint64_t, R*maxBlocks)
	var b C.eppscore_batch
	b.struct_size = C.uint32_t(unsafe.Sizeof(b))
	b.R = C.int32_t(R)
	pin.Pin(&bytes[0]); pin.Pin(&off[0]); pin.Pin(&plen[0]); pin.Pin(&seed[0]); pin.Pin(&adapter[0]); pin.Pin(&mask[0])
	pin.Pin(&ttft[0]); pin.Pin(&tpot[0]); pin.Pin(&pick[0]); pin.Pin(&score[0]); pin.Pin(&ties[0]); pin.Pin(&total[0]); pin.Pin(&hashes[0])
	b.prompt_bytes = (*C.uint8_t)(unsafe.Pointer(&bytes[0]))
	b.prompt_off, b.prompt_len, b.model_seed, b.adapter_id = &off[0], &plen[0], &seed[0], &adapter[0]
	b.cand_mask = &mask[0] // the candidate subset of every request (director candidates.go:98); Go Filter plugins clear more bits here
	b.ttft_slo, b.tpot_slo = &ttft[0], &tpot[0]
	b.pick, b.pick_score, b.tie_count = &pick[0], &score[0], &ties[0]
	b.total_blocks, b.hashes_out = &total[0], &hashes[0]
	b.max_blocks = C.int32_t(maxBlocks)

	// ---- 3. one call ----
	if rc := C.eppscore_schedule_batch(s.eng, &b); rc != C.EPPSCORE_OK {
		fail(batch, fmt.Errorf("eppscore_schedule_batch: %s", C.GoString(C.eppscore_last_error(s.eng))))
		return
	}

	// ---- 4. results: the same SchedulingResult the reference builds (scheduler.go:86-101, types.go:152-170) ----
	for r, c := range batch {
		if pick[r] < 0 {
			c.err = fmt.Errorf("failed to run scheduler profile '%s'", profileName) // single_profile_handler.go:89-91
		} else {
			ep := byID[int32(pick[r])]
			c.res = &fwksched.SchedulingResult{
				PrimaryProfileName: profileName,
				ProfileResults: map[string]*fwksched.ProfileRunResult{profileName: {
					TargetEndpoints: []fwksched.Endpoint{&fwksched.ScoredEndpoint{Endpoint: ep, Score: float64(score[r])}}}},
			}
		}
		close(c.done)
	}

	// ---- 5. PreRequest (approximateprefix/plugin.go:169-197): record the picks in the device-resident index.  The
	//         reference does this off the critical path too (plugin.go:188-193); the callers were released above. ----
	pin.Pin(&lruCap[0])
	if rc := C.eppscore_commit_picks(s.eng, C.int32_t(R), &pick[0], &hashes[0], &total[0], C.int32_t(maxBlocks), &lruCap[0]); rc != C.EPPSCORE_OK {
		// the index stays consistent (room is guaranteed before anything is touched); the batch's affinity is lost
		_ = C.GoString(C.eppscore_last_error(s.eng))
	}
}
