// libxgc.so — host side of the C ABI declared in include/xgc/abi.h.  Owns the device state of one GPU's shard of
// replicates, sequences the weekly kernels in the reference's module order (sim1_02_lhs.r:200-218), and converts
// between the reference's dat$attr / dat$el view (R positions, one vector per attribute) and the packed device planes.
// There is NO CPU fallback: every entry point that computes launches the CUDA kernels in xgc_kernels.cuh.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <type_traits>
#include <vector>
#include "xgc/abi.h"
#include "xgc_kernels.cuh"
#include "xgc_fast.cuh"

static thread_local std::string g_create_error;

struct HostRep {                      // host mirror of one replicate (cache for attribute / edge exchange)
  int r = -1; bool dirty = false;
  Rep rep;
  std::vector<uint4> core; std::vector<Aux> aux;
  std::vector<short4> gck_a, gck_b, gck_d, hck_a, hck_b, tck; std::vector<int> arr_t;
  std::vector<int4> edge[3];
  std::vector<int> lazy[3];            // materialised kissing | rimming << 8 counts (xgc_get_actcounts)
  std::vector<int> pos2slot, slot2pos;
  std::vector<int> deg[3];             // per slot, filled on demand by the derived degree attributes
};

struct xgc_handle {
  xgc_config cfg;
  Dev g;                               // device pointers + geometry
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int last_at = -1;
  int compacted_at = -1;               // last week after which the slots were compacted
  int mode = XGC_MODE_STRICT;          // arithmetic / sampling contract of native-draw runs (abi.h)
  size_t fast_smem = 0;                // dynamic shared memory of k_week_fast for this geometry (0: does not fit)
  bool fast_dirty = true;              // the spare core bits the fast path reads (birth week, ins.quot) need recomputing
  HostRep cache;
  double* d_inj = nullptr; size_t d_inj_cap = 0;
  void* d_cscratch = nullptr;   // 16 B per slot, large-replicate compaction only
  unsigned* d_pstat = nullptr; int* d_remap = nullptr; double* d_targets = nullptr; int* d_scratch = nullptr; double* d_pois = nullptr; double* d_inv_ntx = nullptr;
  int* d_el_stage = nullptr; size_t el_stage_cap = 0;
  int* d_upd_stage[3] = { nullptr, nullptr, nullptr }; size_t upd_stage_cap[3] = { 0, 0, 0 };    // per layer: the copies of one week overlap
  cudaGraphExec_t week_graph = nullptr; uint64_t graph_launches = 0;
  // xgc_run drives the replicates as TWO concurrent lanes (halves of the replicate range, each with its own streams and
  // captured week graph): every kernel of this path is latency-bound, so a second, independent kernel sequence fills the
  // issue slots and tails the first leaves idle (measured -8 % per week).  Lane 0 lives on the handle's own streams.
  int n_lanes = 2;                        // XGC_LANES=1: one lane
  const Dev* gl = nullptr;                // Dev of the lane being enqueued (nullptr: the whole handle)
  Dev lane_dev[2];                        // per-lane views of g (kept on the handle: gl never points at a dead stack frame)
  cudaStream_t lane_main = nullptr, lane_side = nullptr;      // streams the enqueue helpers use (lane 0: stream / stream2)
  cudaEvent_t* lane_fork = nullptr; cudaEvent_t* lane_join = nullptr;
  cudaStream_t l1_stream = nullptr, l1_stream2 = nullptr;
  cudaEvent_t l1_fork[2] = { nullptr, nullptr }, l1_join[2] = { nullptr, nullptr }, l1_start = nullptr, l1_done = nullptr;
  cudaGraphExec_t l1_graph = nullptr; uint64_t l1_graph_launches = 0;
  int lanes_captured = 0;
  bool rank_fresh = false;             // pos2slot matches the current population (reset by every step)
  int form_tiles = 0;                  // tiles of 1024 formation candidates a layer can need (from the tables; 0: not known, use lb_tiles)
  bool plain_next = false;             // the next kernel launch follows an async copy / memset on the stream: no programmatic
                                       // serialization for it (its griddepcontrol.wait only covers preceding GRIDS)
  // Fast mode, XGC_FUSE_HOSTNET=1: xgc_update_edgelists_all / xgc_set_edgelists_all copy their host buffers to the device and QUEUE the
  // operation; the next launch of the resident kernel that runs an acts pass applies it itself (no k_rank / k_edges_update /
  // k_edges_upload launches).  Anything else that reads the lists flushes the queue through the standalone kernels first
  // (flush_pending_edges).  Off by default: the in-kernel phase is cold code on every CTA's critical path (~40k cycles per replicate-
  // week, instruction-fetch bound) -- it shortens the weekly round trip of ONE handle by 5 % (125 replicates: 0.294 -> 0.279 ms) but
  // costs 3-5 % when several handles keep the GPU busy (1000 replicates in 7 handles: 1.02 -> 1.05 ms); DESIGN.md section 6.
  struct PendEdge { int kind = 0; const int* ptr = nullptr; const int* ne = nullptr; const int* start = nullptr; int stride = 0; int has_start = 0; };
  PendEdge pend[3];
  bool defer_edges = false;
  unsigned* d_carry = nullptr;         // [R][NCOUNTER] running state tallies a span call leaves for the next one (xgc_fast.cuh)
  int carry_week = -1;                 // week whose running tallies d_carry holds (-1: none / agent state changed since)
  bool prune_pending = false;          // a fast-path aging..hivvl call left ties of departed agents in the lists (dropped lazily:
                                       // k_edges_update, the fast acts pass, or prune_now() before any other reader)
  bool acts_valid = false;             // act counts of the current week exist (lazy kissing / rimming counts may be materialised)
  Inj last_inj;                        // draw source of the most recent xgc_step (lazy act counts must use the same)
  bool profiling = false;
  bool launch_err = false;
  cudaStream_t stream2 = nullptr;        // side branch for kernels that are independent of the main chain
  cudaEvent_t ev_fork[2] = { nullptr, nullptr }, ev_join[2] = { nullptr, nullptr };
  cudaStream_t cur = nullptr;            // stream the next kl() launches on (main stream unless inside a side branch)
  bool overlap = true;                   // XGC_OVERLAP=0 keeps everything on one stream
  bool pdl = true;            // programmatic dependent launch for kernels launched on the stream by xgc_step (XGC_PDL=0: off).  Measured: +8 % on the
                              // step-by-step host-buffer path; inside the captured week graph it is OFF (no gain there, and a 6 % loss once
                              // neighbouring kernels differ in shared-memory carve-out)
  bool capturing = false;
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<void*> allocs;
  size_t n_agent() const { return (size_t)cfg.n_replicates * cfg.n_cap; }
};

#define FAIL(h, code, ...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); (h)->err = _b; return (code); } while (0)
#define CU(h, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) FAIL(h, 100 + (int)_e, "%s failed: %s", #call, cudaGetErrorString(_e)); } while (0)
#define CHECK_R(h, r) do { if ((r) < 0 || (r) >= (h)->cfg.n_replicates) FAIL(h, 2, "replicate %d out of range [0,%d)", (r), (h)->cfg.n_replicates); } while (0)

template <typename T> static int dalloc(xgc_handle* h, T** p, size_t n) {
  CU(h, cudaMalloc((void**)p, n * sizeof(T)));
  CU(h, cudaMemset(*p, 0, n * sizeof(T)));      // (legacy stream, asynchronous: xgc_create ends with a device synchronisation)
  h->allocs.push_back(*p);
  return 0;
}

static Inj make_inj(const xgc_handle* h, const xgc_inject* inj, const double* d_u) {
  Inj j; memset(&j, 0, sizeof j);
  if (inj && inj->u) {
    j.u = d_u; j.stride = inj->stride; j.form_cap = inj->form_cap;
    for (int s = 0; s < XGC_NSTREAM; ++s) j.off[s] = xgc_inject_offset(inj, s);
  }
  (void)h;
  return j;
}

extern "C" int xgc_constant(const char* name, int64_t* value) {
#define X(c) if (!strcmp(name, #c)) { *value = (int64_t)(c); return 0; }
  XGC_CONSTANT_LIST(X)
#undef X
  return 1;
}
extern "C" int xgc_inject_stream_width(int stream) { return stream >= 0 && stream < XGC_NSTREAM ? xgc_stream_width(stream) : -1; }
extern "C" size_t xgc_inject_table_size(const xgc_inject* d) { return d ? xgc_inject_size(d) : 0; }
extern "C" int xgc_param_index(const char* name, int* offset, int* count) {
#define XGC_PARAM(n, c, v) if (!strcmp(name, #n)) { *offset = XGC_P_##n; *count = (c); return 0; }
#include "xgc/params.def"
#undef XGC_PARAM
  return 1;
}
extern "C" int xgc_abi_version(void) { return XGC_ABI_VERSION; }
extern "C" const char* xgc_last_error(const xgc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// CUDA loads kernels lazily, on their first launch (tens of ms for a large image): touch every native-draw kernel once per
// process at create time so that no week of a run pays for it (the compaction kernels first run at week 64).
__global__ void k_edges_upload(const __grid_constant__ Dev g, int l, const int* el, const int* ne, const int* start, int stride);
__global__ void k_edges_update(const __grid_constant__ Dev g, int l, const int* delta, int stride, int has_start);
static void preload_kernels() {
  static bool done = false;
  if (done) return;
  done = true;
  cudaFuncAttributes a;
#define PRELOAD(k) cudaFuncGetAttributes(&a, k)
  PRELOAD(k_derive); PRELOAD((k_begin<false, false>)); PRELOAD(k_agent1a<false>); PRELOAD(k_agent1b<false>);
  PRELOAD((k_edges_resim<false, TPB, 1>)); PRELOAD(k_edges_resim_lb<false>); PRELOAD(k_rank); PRELOAD(k_rank_cta);
  PRELOAD((k_form<false, true, 512>));
  cudaFuncSetAttribute(k_form<false, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);      // bitmap + prefix + 16-bit position table, n_cap <= 32k
  cudaFuncSetAttribute(k_form<true, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024); PRELOAD(k_form_lb<false>); PRELOAD(k_edge1<false>);
  PRELOAD(k_agent2a<false>); PRELOAD(k_agent2b<false>); PRELOAD(k_edge2a); PRELOAD((k_edge2b<false, true>)); PRELOAD((k_edge2b<false, false>));
  PRELOAD(k_agent3a); PRELOAD(k_agent3b<false>); PRELOAD(k_targets); PRELOAD(k_compact_agents); PRELOAD(k_compact_edges);
  PRELOAD(k_compact_remap); PRELOAD(k_compact_finish); PRELOAD(k_edges_upload); PRELOAD(k_edges_update);
  PRELOAD(k_compact_gather<uint4>); PRELOAD(k_compact_gather<uint2>); PRELOAD(k_compact_gather<int>);
  PRELOAD(k_compact_putback<uint4>); PRELOAD(k_compact_putback<uint2>); PRELOAD(k_compact_putback<int>);
#undef PRELOAD
  cudaGetLastError();
}

extern "C" int xgc_create(const xgc_config* cfg, xgc_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return 1; }
  if (cfg->abi_version != XGC_ABI_VERSION) { g_create_error = "ABI version mismatch"; return 1; }
  if (cfg->n_replicates < 1 || cfg->n_replicates > 65535 || cfg->n_cap < 128 || cfg->n_cap % 128 || cfg->t_cap < 2) {
    g_create_error = "bad config: need 1 <= n_replicates <= 65535, n_cap a positive multiple of 128, t_cap >= 2"; return 1;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libxgc has no CPU fallback)"; return 3; }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return 1; }
  xgc_handle* h = new xgc_handle();
  h->cfg = *cfg;
  memset(&h->last_inj, 0, sizeof(Inj));
  if (h->cfg.compact_every <= 0) h->cfg.compact_every = 64;
  auto fail = [&](int code) { g_create_error = h->err; xgc_destroy(h); return code; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(3); }
  { const char* e = getenv("XGC_PDL"); if (e && e[0] == '0') h->pdl = false; }
  preload_kernels();
  { const char* e = getenv("XGC_OVERLAP"); if (e && e[0] == '0') h->overlap = false; }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return fail(3); }
  if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return fail(3); }
  for (int k = 0; k < 2; ++k)
    if (cudaEventCreateWithFlags(&h->ev_fork[k], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming) != cudaSuccess) {
      h->err = "cudaEventCreate failed"; return fail(3);
    }
  { const char* e = getenv("XGC_FUSE_HOSTNET"); if (e) h->defer_edges = e[0] == '1'; }
  { const char* e = getenv("XGC_LANES"); if (e && e[0] == '1') h->n_lanes = 1; else if (e && e[0] == '2') h->n_lanes = 2; }
  if (cfg->n_replicates < 2) h->n_lanes = 1;
  if (cfg->n_cap > 32768 && !getenv("XGC_LANES")) h->n_lanes = 1;      // large replicates fill the GPU kernel by kernel: a second lane only contends (100 x 100k: 2.00 vs 1.39 ms per week)
  if (cudaStreamCreateWithFlags(&h->l1_stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h->l1_stream2, cudaStreamNonBlocking) != cudaSuccess) {
    h->err = "cudaStreamCreate failed"; return fail(3);
  }
  for (int k = 0; k < 2; ++k)
    if (cudaEventCreateWithFlags(&h->l1_fork[k], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->l1_join[k], cudaEventDisableTiming) != cudaSuccess) {
      h->err = "cudaEventCreate failed"; return fail(3);
    }
  if (cudaEventCreateWithFlags(&h->l1_start, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->l1_done, cudaEventDisableTiming) != cudaSuccess) {
    h->err = "cudaEventCreate failed"; return fail(3);
  }
  h->lane_main = h->stream; h->lane_side = h->stream2; h->lane_fork = h->ev_fork; h->lane_join = h->ev_join;
  h->cur = h->stream;
  Dev& g = h->g; memset(&g, 0, sizeof g);
  g.R = cfg->n_replicates; g.n_cap = cfg->n_cap; g.t_cap = cfg->t_cap; g.rep0 = cfg->replicate_offset; g.seed = (unsigned)cfg->seed;
  size_t eo = 0;
  for (int l = 0; l < 3; ++l) { g.e_cap[l] = std::max(cfg->e_cap[l], 1); g.e_off[l] = eo; eo += ((size_t)g.e_cap[l] + 31) & ~(size_t)31; }
  g.e_stride = eo;
  h->fast_smem = xgc_fast_smem_bytes(g.n_cap, g.e_stride);
  size_t NA = h->n_agent(); int rc = 0;
  // n_cap is a multiple of 128 but the agent passes load whole 256-slot tiles (the first tile of a CTA and the next tile
  // of the register double buffer are loaded before the bounds check): every agent plane carries one tile of padding
  const size_t NAP = NA + TPB;
  rc |= dalloc(h, &g.core, NAP); rc |= dalloc(h, &g.aux, NAP);
  rc |= dalloc(h, &g.gck_a, NAP); rc |= dalloc(h, &g.gck_b, NAP); rc |= dalloc(h, &g.gck_d, NAP);
  rc |= dalloc(h, &g.hck_a, NAP); rc |= dalloc(h, &g.hck_b, NAP); rc |= dalloc(h, &g.tck, NAP); rc |= dalloc(h, &g.arr_t, NAP);
  rc |= dalloc(h, &g.alive, NA / 32); rc |= dalloc(h, &g.pos2slot, NA); rc |= dalloc(h, &g.deg, NA);
  if (g.n_cap > 32768) { uint4* cs = nullptr; rc |= dalloc(h, &cs, NA); h->d_cscratch = cs; }     // parallel slot compaction, 16 B per slot
  rc |= dalloc(h, &g.edge, (size_t)g.R * g.e_stride);
  rc |= dalloc(h, &g.rep, (size_t)g.R);
  { int* ids = nullptr; rc |= dalloc(h, &ids, (size_t)g.R); g.rep_id = ids; rc |= dalloc(h, &g.rep_cycles, (size_t)g.R); }
  double* dp; int* dc; double* dt; int* dat;
  rc |= dalloc(h, &dp, (size_t)g.R * XGC_NPARAM); rc |= dalloc(h, &dc, (size_t)g.R * XGC_NCONTROL);
  rc |= dalloc(h, &dt, (size_t)XGC_NTABLE); rc |= dalloc(h, &dat, (size_t)g.R);
  rc |= dalloc(h, &g.counters, (size_t)g.R * g.t_cap * XGC_NCOUNTER);
  rc |= dalloc(h, &g.q_items, (size_t)g.R * 4 * g.e_stride); rc |= dalloc(h, &g.q_count, (size_t)g.R * 4);
  rc |= dalloc(h, &g.qa_items, NA); rc |= dalloc(h, &g.qa_count, (size_t)g.R); rc |= dalloc(h, &g.q3_count, (size_t)g.R);
  rc |= dalloc(h, &g.qh_items, NA); rc |= dalloc(h, &g.qh_count, (size_t)g.R);
  g.lb_tiles = (std::max(g.e_cap[0], std::max(g.e_cap[1], g.e_cap[2])) + 1023) / 1024;
  rc |= dalloc(h, &g.lb_status, (size_t)g.R * 3 * g.lb_tiles); rc |= dalloc(h, &g.lb_ticket, (size_t)g.R * 3);
  rc |= dalloc(h, &h->d_pois, (size_t)g.R * 4 * (XGC_MAX_ACTS + 1)); rc |= dalloc(h, &h->d_inv_ntx, (size_t)g.R * 4); rc |= dalloc(h, &h->d_pstat, (size_t)g.R * 4); rc |= dalloc(h, &h->d_carry, (size_t)g.R * XGC_NCOUNTER);
  rc |= dalloc(h, &h->d_remap, NA); rc |= dalloc(h, &h->d_targets, (size_t)g.R * XGC_NTARGET); rc |= dalloc(h, &h->d_scratch, 16);
  if (rc) return fail(rc);
  g.params = dp; g.control = dc; g.tables = dt; g.at_ptr = dat; g.pois = h->d_pois; g.inv_ntx = h->d_inv_ntx;
  // defaults: parameter defaults from params.def, kissing/rimming routes on
  std::vector<double> P(XGC_NPARAM);
#define XGC_PARAM(name, count, v) for (int k = 0; k < (count); ++k) P[XGC_P_##name + k] = (v);
#include "xgc/params.def"
#undef XGC_PARAM
  std::vector<int32_t> C(XGC_NCONTROL, 0); C[XGC_C_transRoute_Kissing] = 1; C[XGC_C_transRoute_Rimming] = 1;
  { std::vector<int> ids((size_t)g.R); for (int r = 0; r < g.R; ++r) ids[r] = g.rep0 + r;
    if (cudaMemcpy((int*)g.rep_id, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "replicate id upload failed"; return fail(3); } }
  if (xgc_set_params(h, -1, P.data()) || xgc_set_control(h, -1, C.data())) return fail(4);
  if (cudaDeviceSynchronize() != cudaSuccess) { h->err = "device synchronisation after allocation failed"; return fail(3); }      // the zero-fills above ran on the legacy stream
  *out = h;
  return 0;
}

extern "C" int xgc_destroy(xgc_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->week_graph) cudaGraphExecDestroy(h->week_graph);
  if (h->l1_graph) cudaGraphExecDestroy(h->l1_graph);
  for (int k = 0; k < 2; ++k) { if (h->l1_fork[k]) cudaEventDestroy(h->l1_fork[k]); if (h->l1_join[k]) cudaEventDestroy(h->l1_join[k]); }
  if (h->l1_start) cudaEventDestroy(h->l1_start);
  if (h->l1_done) cudaEventDestroy(h->l1_done);
  if (h->l1_stream) { cudaStreamSynchronize(h->l1_stream); cudaStreamDestroy(h->l1_stream); }
  if (h->l1_stream2) cudaStreamDestroy(h->l1_stream2);
  for (void* p : h->allocs) cudaFree(p);
  if (h->d_inj) cudaFree(h->d_inj);
  if (h->d_el_stage) cudaFree(h->d_el_stage);
  for (int l = 0; l < 3; ++l) if (h->d_upd_stage[l]) cudaFree(h->d_upd_stage[l]);
  for (int k = 0; k < 2; ++k) { if (h->ev_fork[k]) cudaEventDestroy(h->ev_fork[k]); if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]); }
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int xgc_sync(xgc_handle* h) { CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream)); return 0; }
extern "C" int xgc_kernel_launches(xgc_handle* h, uint64_t* n) { *n = h->launches; return 0; }
extern "C" int xgc_stream_ptr(xgc_handle* h, void** s) { *s = (void*)h->stream; return 0; }

// --------------------------------------------------------------------------------------------------------------------
// configuration
// --------------------------------------------------------------------------------------------------------------------
// Host <-> device copies of a handle go through ITS stream and wait for it.  A blocking cudaMemcpy runs on the legacy stream,
// which the handles' non-blocking streams are not ordered against, and from pageable host memory it returns once the data is
// STAGED, not once the DMA has landed: the first kernels after xgc_set_params_all read the last replicates' parameters before
// they arrived (about one run in four of a 1000-replicate handle diverged in its last ~150 replicates: tools/lane_determinism.py).
static cudaError_t h2d(xgc_handle* h, void* d, const void* s, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->stream);
}
static cudaError_t d2h(xgc_handle* h, void* d, const void* s, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->stream);
}
// Global ids of this handle's replicates (the Philox key of replicate r is (seed, id[r])): lets a host pack ANY subset of a sweep into
// a handle -- e.g. replicates of similar cost, so that the one-wave launches of the host-network loop have short tails -- with results
// that do not depend on the packing.  Default: replicate_offset + r.
extern "C" int xgc_set_replicate_ids(xgc_handle* h, const int32_t* ids) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, h2d(h, (int*)h->g.rep_id, ids, sizeof(int) * (size_t)h->g.R));
  return 0;
}
// SM cycles the replicate-resident kernel has spent on each replicate since the last reset (a cost measure for such packing)
extern "C" int xgc_replicate_cost(xgc_handle* h, uint64_t* cycles /*[R]*/, int reset) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  if (cycles) CU(h, d2h(h, cycles, h->g.rep_cycles, sizeof(uint64_t) * (size_t)h->g.R));
  if (reset) { CU(h, cudaMemsetAsync(h->g.rep_cycles, 0, sizeof(uint64_t) * (size_t)h->g.R, h->stream)); CU(h, cudaStreamSynchronize(h->stream)); h->plain_next = true; }
  return 0;
}
static void drop_graphs(xgc_handle* h) {            // kernel arguments baked into the captured week graphs changed
  if (h->week_graph) { cudaGraphExecDestroy(h->week_graph); h->week_graph = nullptr; }
  if (h->l1_graph) { cudaGraphExecDestroy(h->l1_graph); h->l1_graph = nullptr; }
  h->lanes_captured = 0;
}
static int derive(xgc_handle* h) {       // per-replicate tables that depend on the parameters
  int n = h->g.R * 4;
  k_derive<<<(n + 127) / 128, 128, 0, h->stream>>>(h->g, h->d_pois, h->d_inv_ntx, h->d_pstat);
  CU(h, cudaGetLastError());
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int xgc_set_params(xgc_handle* h, int r, const double* p) {
  CU(h, cudaSetDevice(h->cfg.device));
  if (r >= h->cfg.n_replicates) FAIL(h, 2, "replicate out of range");
  CU(h, cudaStreamSynchronize(h->stream));
  if (r < 0) {                           // every replicate: one host-side replication, one copy
    std::vector<double> all((size_t)h->cfg.n_replicates * XGC_NPARAM);
    for (int k = 0; k < h->cfg.n_replicates; ++k) memcpy(all.data() + (size_t)k * XGC_NPARAM, p, sizeof(double) * XGC_NPARAM);
    CU(h, h2d(h, (double*)h->g.params, all.data(), all.size() * sizeof(double)));
  } else CU(h, h2d(h, (double*)h->g.params + (size_t)r * XGC_NPARAM, p, sizeof(double) * XGC_NPARAM));
  return derive(h);
}
extern "C" int xgc_set_params_all(xgc_handle* h, const double* p /*[R][XGC_NPARAM]*/) {
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, h2d(h, (double*)h->g.params, p, sizeof(double) * XGC_NPARAM * h->cfg.n_replicates));
  return derive(h);
}
extern "C" int xgc_get_params(xgc_handle* h, int r, double* p) {
  CHECK_R(h, r); CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, p, h->g.params + (size_t)r * XGC_NPARAM, sizeof(double) * XGC_NPARAM));
  return 0;
}
extern "C" int xgc_set_control(xgc_handle* h, int r, const int32_t* c) {
  CU(h, cudaSetDevice(h->cfg.device));
  if (r >= h->cfg.n_replicates) FAIL(h, 2, "replicate out of range");
  CU(h, cudaStreamSynchronize(h->stream));
  if (r < 0) {
    std::vector<int32_t> all((size_t)h->cfg.n_replicates * XGC_NCONTROL);
    for (int k = 0; k < h->cfg.n_replicates; ++k) memcpy(all.data() + (size_t)k * XGC_NCONTROL, c, sizeof(int32_t) * XGC_NCONTROL);
    CU(h, h2d(h, (int*)h->g.control, all.data(), all.size() * sizeof(int32_t)));
  } else CU(h, h2d(h, (int*)h->g.control + (size_t)r * XGC_NCONTROL, c, sizeof(int32_t) * XGC_NCONTROL));
  return 0;
}
extern "C" int xgc_get_tables(xgc_handle* h, double* t) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, t, h->g.tables, sizeof(double) * XGC_NTABLE));
  return 0;
}
extern "C" int xgc_get_params_all(xgc_handle* h, double* p) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, p, h->g.params, sizeof(double) * XGC_NPARAM * h->cfg.n_replicates));
  return 0;
}
extern "C" int xgc_get_control(xgc_handle* h, int r, int32_t* c) {
  CHECK_R(h, r); CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, c, h->g.control + (size_t)r * XGC_NCONTROL, sizeof(int32_t) * XGC_NCONTROL));
  return 0;
}
static void set_form_tiles(xgc_handle* h, const double* t) {
  double mx = 0.0;
  for (int l = 0; l < 3; ++l) mx = std::max(mx, t[XGC_T_NET_DEG + l] * (double)h->cfg.n_cap / 2.0 / (l < 2 ? t[XGC_T_NET_DUR + l] : 1.0));
  h->form_tiles = (int)std::min((double)h->g.lb_tiles, std::ceil((mx + 1.0) / 1024.0) + 1.0);
}
extern "C" int xgc_set_tables(xgc_handle* h, const double* t) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, h2d(h, (double*)h->g.tables, t, sizeof(double) * XGC_NTABLE));
  for (int l = 0; l < 2; ++l) h->g.inv_dur[l] = 1.0 / t[XGC_T_NET_DUR + l];       // same IEEE division the oracle performs
  // formation grid: candidates of a layer-week <= deg * n_cap / 2 (/ duration) + 1, in tiles of 1024 (the look-back kernel's grid was
  // sized for the edge lists: thousands of CTAs that took a ticket and left)
  set_form_tiles(h, t);
  drop_graphs(h);
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// host cache of one replicate
// --------------------------------------------------------------------------------------------------------------------
template <typename T> static int pull(xgc_handle* h, std::vector<T>& v, const T* d, size_t n) {
  v.resize(n); if (n) CU(h, d2h(h, v.data(), d, n * sizeof(T))); return 0;
}
template <typename T> static int push(xgc_handle* h, const std::vector<T>& v, T* d) {
  if (!v.empty()) CU(h, h2d(h, d, v.data(), v.size() * sizeof(T))); return 0;
}
static void rebuild_maps(HostRep& c) {
  c.pos2slot.clear(); c.slot2pos.assign(c.core.size(), -1);
  for (size_t i = 0; i < c.core.size(); ++i) if (c.core[i].y & DM_ACTIVE) { c.slot2pos[i] = (int)c.pos2slot.size(); c.pos2slot.push_back((int)i); }
}
static int prune_now(xgc_handle* h);
static int flush_pending_edges(xgc_handle* h);
static int launch_edge_op(xgc_handle* h, int layer, const xgc_handle::PendEdge& q);
static bool can_defer_edges(const xgc_handle* h) { return h->mode == XGC_MODE_FAST && h->defer_edges && xgc_fast_hostnet_fits(h->g.n_cap, h->g.e_cap); }
static int cache_flush(xgc_handle* h) {
  HostRep& c = h->cache;
  if (c.r < 0 || !c.dirty) { return 0; }
  const Dev& g = h->g; size_t o = (size_t)c.r * g.n_cap;
  CU(h, cudaSetDevice(h->cfg.device));
  int rc = 0;
  rc |= push(h, c.core, g.core + o); rc |= push(h, c.aux, g.aux + o);
  rc |= push(h, c.gck_a, g.gck_a + o); rc |= push(h, c.gck_b, g.gck_b + o); rc |= push(h, c.gck_d, g.gck_d + o);
  rc |= push(h, c.hck_a, g.hck_a + o); rc |= push(h, c.hck_b, g.hck_b + o); rc |= push(h, c.tck, g.tck + o); rc |= push(h, c.arr_t, g.arr_t + o);
  if (rc) return rc;
  std::vector<unsigned> alive(g.n_cap / 32, 0u);
  for (size_t i = 0; i < c.core.size(); ++i) if (c.core[i].y & DM_ACTIVE) alive[i >> 5] |= 1u << (i & 31);
  CU(h, h2d(h, g.alive + (size_t)c.r * (g.n_cap / 32), alive.data(), alive.size() * sizeof(unsigned)));
  for (int l = 0; l < 3; ++l) {
    c.rep.ne[l] = (int)c.edge[l].size();
    if (!c.edge[l].empty()) CU(h, h2d(h, g.edge + (size_t)c.r * g.e_stride + g.e_off[l], c.edge[l].data(), c.edge[l].size() * sizeof(int4)));
  }
  CU(h, h2d(h, g.rep + c.r, &c.rep, sizeof(Rep)));
  c.dirty = false; h->rank_fresh = false;
  return 0;
}
static int cache_load(xgc_handle* h, int r) {
  HostRep& c = h->cache;
  if (c.r == r) return 0;
  int rc = cache_flush(h); if (rc) return rc;
  rc = prune_now(h); if (rc) return rc;
  const Dev& g = h->g; size_t o = (size_t)r * g.n_cap;
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, &c.rep, g.rep + r, sizeof(Rep)));
  size_t n = (size_t)c.rep.n_slots;
  rc |= pull(h, c.core, g.core + o, n); rc |= pull(h, c.aux, g.aux + o, n);
  rc |= pull(h, c.gck_a, g.gck_a + o, n); rc |= pull(h, c.gck_b, g.gck_b + o, n); rc |= pull(h, c.gck_d, g.gck_d + o, n);
  rc |= pull(h, c.hck_a, g.hck_a + o, n); rc |= pull(h, c.hck_b, g.hck_b + o, n); rc |= pull(h, c.tck, g.tck + o, n); rc |= pull(h, c.arr_t, g.arr_t + o, n);
  for (int l = 0; l < 3; ++l) rc |= pull(h, c.edge[l], g.edge + (size_t)r * g.e_stride + g.e_off[l], (size_t)c.rep.ne[l]);
  if (rc) return rc;
  rebuild_maps(c);
  for (int l = 0; l < 3; ++l) c.lazy[l].clear();
  c.r = r; c.dirty = false;
  return 0;
}
static int cache_drop(xgc_handle* h) { int rc = cache_flush(h); h->cache.r = -1; return rc; }

// --------------------------------------------------------------------------------------------------------------------
// attributes (dat$attr)
// --------------------------------------------------------------------------------------------------------------------
struct AttrDef { const char* name; std::function<double(const HostRep&, int)> get; std::function<void(HostRep&, int, double)> set; };
static inline short S16(double v) { return (short)(int)v; }
static void setbits(unsigned& w, unsigned mask, int shift, double v) { w = (w & ~(mask << shift)) | (((unsigned)(int)v & mask) << shift); }
static std::vector<AttrDef> build_attrs() {
  std::vector<AttrDef> A;
  A.push_back({"uid", [](const HostRep& c, int i) { return (double)c.core[i].x; }, [](HostRep& c, int i, double v) { c.core[i].x = (unsigned)v; }});
  A.push_back({"age", [](const HostRep& c, int i) { return c.aux[i].age_ref + (double)(c.rep.t_age - c.rep.t_ref) * XGC_WEEK; },
               [](HostRep& c, int i, double v) { c.aux[i].age_ref = v; setbits(c.core[i].y, 7u, 2, (double)((v >= 25.0) + (v >= 35.0) + (v >= 45.0) + (v >= 55.0))); }});
  A.push_back({"race", [](const HostRep& c, int i) { return (double)(DM_RACE(c.core[i].y) + 1); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 3u, 0, v - 1); }});
  A.push_back({"age.grp", [](const HostRep& c, int i) { return (double)(DM_AGEGRP(c.core[i].y) + 1); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 7u, 2, v - 1); }});
  A.push_back({"role.class", [](const HostRep& c, int i) { return (double)DM_ROLE(c.core[i].y); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 3u, 5, v); }});
  A.push_back({"ins.quot", [](const HostRep& c, int i) { return (double)c.aux[i].ins_quot; }, [](HostRep& c, int i, double v) { c.aux[i].ins_quot = (float)v; }});
  A.push_back({"risk.grp", [](const HostRep& c, int i) { return (double)(DM_RISK(c.core[i].y) + 1); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 7u, 7, v - 1); }});
  A.push_back({"circ", [](const HostRep& c, int i) { return (double)((c.core[i].y >> 10) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 1u, 10, v); }});
  A.push_back({"late.tester", [](const HostRep& c, int i) { return (double)((c.core[i].y >> 11) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 1u, 11, v); }});
  A.push_back({"tt.traj", [](const HostRep& c, int i) { return (double)DM_TTTRAJ(c.core[i].y); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 3u, 12, v); }});
  A.push_back({"act.stopper", [](const HostRep& c, int i) { return (double)((c.core[i].y >> 14) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].y, 1u, 14, v); }});
  A.push_back({"arrival.time", [](const HostRep& c, int i) { return (double)c.arr_t[i]; }, [](HostRep& c, int i, double v) { c.arr_t[i] = (int)v; }});
  static const char* const site[3] = { "rGC", "uGC", "pGC" };
  for (int s = 0; s < 3; ++s) {
    static std::string names[3][6];
    const char* suf[6] = { "", ".infTime", ".sympt", ".tx", ".txTime", ".dur" };
    for (int k = 0; k < 6; ++k) names[s][k] = std::string(site[s]) + suf[k];
    A.push_back({names[s][0].c_str(), [s](const HostRep& c, int i) { return (double)((c.core[i].z >> s) & 1u); }, [s](HostRep& c, int i, double v) { setbits(c.core[i].z, 1u, s, v); }});
    A.push_back({names[s][1].c_str(), [s](const HostRep& c, int i) { return (double)(&c.gck_a[i].x)[s]; }, [s](HostRep& c, int i, double v) { (&c.gck_a[i].x)[s] = S16(v); }});
    A.push_back({names[s][2].c_str(), [s](const HostRep& c, int i) { return (double)((c.core[i].z >> (3 + s)) & 1u); }, [s](HostRep& c, int i, double v) { setbits(c.core[i].z, 1u, 3 + s, v); }});
    A.push_back({names[s][3].c_str(), [s](const HostRep& c, int i) { return (double)((c.core[i].z >> (6 + s)) & 1u); }, [s](HostRep& c, int i, double v) { setbits(c.core[i].z, 1u, 6 + s, v); }});
    A.push_back({names[s][4].c_str(), [s](const HostRep& c, int i) { return (double)(&c.gck_b[i].x)[s]; }, [s](HostRep& c, int i, double v) { (&c.gck_b[i].x)[s] = S16(v); }});
    A.push_back({names[s][5].c_str(), [s](const HostRep& c, int i) { return (double)(&c.gck_d[i].x)[s]; }, [s](HostRep& c, int i, double v) { (&c.gck_d[i].x)[s] = S16(v); }});
  }
  A.push_back({"last.screen", [](const HostRep& c, int i) { return (double)c.gck_a[i].w; }, [](HostRep& c, int i, double v) { c.gck_a[i].w = S16(v); }});
  A.push_back({"sti.expo", [](const HostRep& c, int i) { return (double)((c.core[i].z >> GC_EXPO_SHIFT) & 0x1Fu); }, [](HostRep& c, int i, double v) { setbits(c.core[i].z, 0x1Fu, GC_EXPO_SHIFT, v); }});
  A.push_back({"status", [](const HostRep& c, int i) { return (double)(c.core[i].w & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 1u, 0, v); }});
  A.push_back({"inf.time", [](const HostRep& c, int i) { return (double)c.hck_a[i].x; }, [](HostRep& c, int i, double v) { c.hck_a[i].x = S16(v); }});
  A.push_back({"stage", [](const HostRep& c, int i) { return (double)HV_STAGE(c.core[i].w); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 7u, 1, v); }});
  A.push_back({"stage.time", [](const HostRep& c, int i) { return (double)c.hck_a[i].z; }, [](HostRep& c, int i, double v) { c.hck_a[i].z = S16(v); }});
  A.push_back({"diag.status", [](const HostRep& c, int i) { return (double)((c.core[i].w >> 4) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 1u, 4, v); }});
  A.push_back({"diag.time", [](const HostRep& c, int i) { return (double)c.hck_a[i].y; }, [](HostRep& c, int i, double v) { c.hck_a[i].y = S16(v); }});
  A.push_back({"last.neg.test", [](const HostRep& c, int i) { return (double)c.tck[i].x; }, [](HostRep& c, int i, double v) { c.tck[i].x = S16(v); }});
  A.push_back({"tx.status", [](const HostRep& c, int i) { return (double)((c.core[i].w >> 5) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 1u, 5, v); }});
  A.push_back({"cuml.time.on.tx", [](const HostRep& c, int i) { return (double)c.hck_b[i].x; }, [](HostRep& c, int i, double v) { c.hck_b[i].x = S16(v); }});
  A.push_back({"cuml.time.off.tx", [](const HostRep& c, int i) { return (double)c.hck_b[i].y; }, [](HostRep& c, int i, double v) { c.hck_b[i].y = S16(v); }});
  A.push_back({"tx.init.time", [](const HostRep& c, int i) { return (double)c.hck_a[i].w; }, [](HostRep& c, int i, double v) { c.hck_a[i].w = S16(v); }});
  A.push_back({"vl", [](const HostRep& c, int i) { return (double)c.aux[i].vl; }, [](HostRep& c, int i, double v) { c.aux[i].vl = (float)v; setbits(c.core[i].w, 1u, 11, (double)((double)(float)v <= XGC_LOG10_200)); }});
  A.push_back({"prepStat", [](const HostRep& c, int i) { return (double)((c.core[i].w >> 6) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 1u, 6, v); }});
  A.push_back({"prepClass", [](const HostRep& c, int i) { return (double)HV_PREPCLASS(c.core[i].w); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 3u, 7, v); }});
  A.push_back({"prepElig", [](const HostRep& c, int i) { return (double)((c.core[i].w >> 9) & 1u); }, [](HostRep& c, int i, double v) { setbits(c.core[i].w, 1u, 9, v); }});
  A.push_back({"prepStartTime", [](const HostRep& c, int i) { return (double)c.tck[i].w; }, [](HostRep& c, int i, double v) { c.tck[i].w = S16(v); }});
  A.push_back({"prepLastRisk", [](const HostRep& c, int i) { return (double)c.tck[i].z; }, [](HostRep& c, int i, double v) { c.tck[i].z = S16(v); }});
  A.push_back({"prep.ind.last", [](const HostRep& c, int i) { return (double)c.tck[i].y; }, [](HostRep& c, int i, double v) { c.tck[i].y = S16(v); }});
  // derived, read-only: current degree in the main / casual / one-off edge list -- what the reference's simnet_msm
  // (resim_nets.FUN with tergmLite, sim1_02_lhs.r:209,226) conditions on as dat$attr$deg.main / deg.casl
  static const char* const degn[3] = { "deg.main", "deg.casl", "deg.inst" };
  for (int l = 0; l < 3; ++l)
    A.push_back({degn[l], [l](const HostRep& c, int i) { return (double)c.deg[l][i]; }, nullptr});
  return A;
}
static const std::vector<AttrDef>& attrs() { static std::vector<AttrDef> A = build_attrs(); return A; }

extern "C" int xgc_attr_names(const char** names, int cap) {
  const auto& A = attrs();
  for (int k = 0; k < (int)A.size() && k < cap; ++k) names[k] = A[k].name;
  return (int)A.size();
}

extern "C" int xgc_set_population(xgc_handle* h, int r, int n, int at) {
  CHECK_R(h, r);
  if (n < 0 || n > h->cfg.n_cap) FAIL(h, 2, "n = %d exceeds n_cap = %d", n, h->cfg.n_cap);
  if (at < 0 || at >= h->cfg.t_cap) FAIL(h, 2, "at = %d outside [0, t_cap = %d)", at, h->cfg.t_cap);
  int rc = cache_drop(h); if (rc) return rc;
  HostRep& c = h->cache;
  memset(&c.rep, 0, sizeof(Rep));
  c.rep.n_slots = c.rep.n_slots_pre = c.rep.n_alive = c.rep.next_uid = n; c.rep.t_ref = c.rep.t_age = at;
  for (int k = 0; k < XGC_NPATH; ++k) c.rep.path_rank[k] = k;
  Aux ax; ax.age_ref = 18.0; ax.vl = 0.f; ax.ins_quot = 0.f;
  c.core.assign(n, make_uint4(0, 0, 0, 0)); c.aux.assign(n, ax);
  short4 na = make_short4(-1, -1, -1, -1);
  c.gck_a.assign(n, na); c.gck_b.assign(n, na); c.gck_d.assign(n, na); c.hck_a.assign(n, na);
  c.hck_b.assign(n, make_short4(0, 0, -1, -1)); c.tck.assign(n, na); c.arr_t.assign(n, at);
  for (int i = 0; i < n; ++i) c.core[i] = make_uint4((unsigned)i, DM_ACTIVE | (2u << 12), 0u, HV_VSUPP);
  for (int l = 0; l < 3; ++l) c.edge[l].clear();
  rebuild_maps(c);
  c.r = r; c.dirty = true;
  // wipe the tail beyond n on the device (stale slots of a previous population)
  const Dev& g = h->g;
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemsetAsync(g.core + (size_t)r * g.n_cap, 0, sizeof(uint4) * (size_t)g.n_cap, h->stream));      // in the handle's stream, like every copy
  CU(h, h2d(h, (int*)g.at_ptr + r, &at, sizeof(int)));
  h->last_at = -1; h->acts_valid = false; h->fast_dirty = true; h->carry_week = -1;
  return cache_flush(h);
}
extern "C" int xgc_num_agents(xgc_handle* h, int r, int* n) {
  CHECK_R(h, r); int rc = cache_load(h, r); if (rc) return rc;
  *n = (int)h->cache.pos2slot.size(); return 0;
}
static const AttrDef* find_attr(const char* name) { for (const auto& a : attrs()) if (!strcmp(a.name, name)) return &a; return nullptr; }

extern "C" int xgc_upload_attr(xgc_handle* h, int r, const char* name, const void* data, size_t n, int dtype) {
  CHECK_R(h, r);
  const AttrDef* a = find_attr(name);
  if (!a) FAIL(h, 5, "unknown attribute '%s'", name);
  if (!a->set) FAIL(h, 5, "attribute '%s' is derived from the edge lists and cannot be uploaded", name);
  int rc = cache_load(h, r); if (rc) return rc;
  HostRep& c = h->cache;
  if (n != c.pos2slot.size()) FAIL(h, 6, "attribute '%s': length %zu != number of agents %zu", name, n, c.pos2slot.size());
  for (size_t p = 0; p < n; ++p) {
    double v = dtype == XGC_F64 ? ((const double*)data)[p] : (double)((const int32_t*)data)[p];
    a->set(c, c.pos2slot[p], v);
  }
  if (!strcmp(name, "age")) c.rep.t_ref = c.rep.t_age;      // re-base the closed-form age clock
  h->fast_dirty = true; h->carry_week = -1;
  if (!strcmp(name, "uid")) { unsigned mx = 0; for (size_t p = 0; p < n; ++p) mx = std::max(mx, c.core[c.pos2slot[p]].x + 1u); c.rep.next_uid = (int)mx; }
  c.dirty = true;
  return 0;
}
extern "C" int xgc_download_attr(xgc_handle* h, int r, const char* name, void* data, size_t cap, int dtype, size_t* n_out) {
  CHECK_R(h, r);
  const AttrDef* a = find_attr(name);
  if (!a) FAIL(h, 5, "unknown attribute '%s'", name);
  int rc = cache_load(h, r); if (rc) return rc;
  HostRep& c = h->cache;
  if (!a->set)                                             // derived degree attributes: count the current edge lists
    for (int l = 0; l < 3; ++l) {
      c.deg[l].assign(c.core.size(), 0);
      for (const int4& ed : c.edge[l]) { c.deg[l][ed.x]++; c.deg[l][ed.y]++; }
    }
  size_t n = c.pos2slot.size();
  if (n_out) *n_out = n;
  if (cap < n) FAIL(h, 6, "attribute '%s': buffer of %zu too small for %zu agents", name, cap, n);
  for (size_t p = 0; p < n; ++p) {
    double v = a->get(c, c.pos2slot[p]);
    if (dtype == XGC_F64) ((double*)data)[p] = v; else ((int32_t*)data)[p] = (int32_t)v;
  }
  return 0;
}

extern "C" int xgc_set_edgelist(xgc_handle* h, int r, int layer, const int32_t* el, int n_edges, const int32_t* start) {
  CHECK_R(h, r);
  if (layer < 0 || layer > 2) FAIL(h, 2, "layer %d out of range", layer);
  if (n_edges < 0 || n_edges > h->g.e_cap[layer]) FAIL(h, 6, "layer %d: %d edges exceed e_cap %d", layer, n_edges, h->g.e_cap[layer]);
  int rc = cache_load(h, r); if (rc) return rc;
  HostRep& c = h->cache; int n = (int)c.pos2slot.size();
  std::vector<int4> E(n_edges);
  for (int e = 0; e < n_edges; ++e) {
    int a = el[e], b = el[n_edges + e];
    if (a < 1 || a > n || b < 1 || b > n || a == b) FAIL(h, 7, "layer %d edge %d: node ids (%d,%d) outside 1..%d", layer, e, a, b, n);
    E[e] = make_int4(c.pos2slot[a - 1], c.pos2slot[b - 1], start ? start[e] : c.rep.t_age, 0);
  }
  c.edge[layer] = E; c.dirty = true; h->acts_valid = false;
  return 0;
}
extern "C" int xgc_get_edgelist(xgc_handle* h, int r, int layer, int32_t* el, int cap, int* n_edges, int32_t* start) {
  CHECK_R(h, r);
  if (layer < 0 || layer > 2) FAIL(h, 2, "layer %d out of range", layer);
  int rc = cache_load(h, r); if (rc) return rc;
  const HostRep& c = h->cache; int ne = (int)c.edge[layer].size();
  if (n_edges) *n_edges = ne;
  if (cap < ne) FAIL(h, 6, "edge buffer of %d too small for %d edges", cap, ne);
  for (int e = 0; e < ne; ++e) {
    int4 ed = c.edge[layer][e];
    el[e] = c.slot2pos[ed.x] + 1; el[ne + e] = c.slot2pos[ed.y] + 1;
    if (start) start[e] = ed.z;
  }
  return 0;
}
// --------------------------------------------------------------------------------------------------------------------
// the step
// --------------------------------------------------------------------------------------------------------------------
static int upload_inject(xgc_handle* h, const xgc_inject* inj, Inj* out) {
  if (!inj || !inj->u) { *out = make_inj(h, nullptr, nullptr); return 0; }
  size_t need = inj->stride * (size_t)h->cfg.n_replicates;
  if (inj->stride != xgc_inject_size(inj)) FAIL(h, 8, "xgc_inject.stride %zu != xgc_inject_size %zu", inj->stride, xgc_inject_size(inj));
  if (need > h->d_inj_cap) {
    if (h->d_inj) cudaFree(h->d_inj);
    h->d_inj = nullptr; h->d_inj_cap = 0;
    CU(h, cudaMalloc((void**)&h->d_inj, need * sizeof(double)));
    h->d_inj_cap = need;
  }
  CU(h, cudaMemcpyAsync(h->d_inj, inj->u, need * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->plain_next = true;
  *out = make_inj(h, inj, h->d_inj);
  return 0;
}

static void prof_begin(xgc_handle* h, const char* name) {
  if (!h->profiling) return;
  xgc_handle::ProfRec r; r.name = name;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, h->stream);
  h->prof.push_back(r);
}
static void prof_end(xgc_handle* h) { if (h->profiling) cudaEventRecord(h->prof.back().b, h->stream); }
// Every weekly kernel is launched with programmatic stream serialization (PDL): the next kernel's CTAs become resident while
// the previous kernel drains and block in griddepcontrol.wait (first statement of every kernel) until it has completed
// and flushed, so the launch ramp overlaps the tail.  Captured into the week graph as programmatic edges.
template <class... KArgs, class... Args>
static void kl(xgc_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->cur ? h->cur : h->lane_main;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (h->pdl && !h->profiling && !h->capturing && !h->plain_next) ? 1 : 0;
  h->plain_next = false;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
  if (e != cudaSuccess && !h->launch_err) {          // remembered for xgc_last_error; the caller's cudaGetLastError() reports it
    h->launch_err = true;
    fprintf(stderr, "xgc: kernel launch failed (%s): grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", cudaGetErrorString(e),
            grid.x, grid.y, grid.z, block.x, block.y, block.z, smem);
  }
}
// Side branch: kernels launched between branch_begin and branch_end go to the second stream and depend only on what
// preceded branch_begin; branch_join makes the main stream wait for them.  Captured into the week graph as a fork / join.
// Per-kernel profiling keeps everything on one stream (the events would serialise the branches anyway).
static bool branch_begin(xgc_handle* h, int k) {
  if (!h->overlap || h->profiling) return false;
  cudaEventRecord(h->lane_fork[k], h->lane_main);
  cudaStreamWaitEvent(h->lane_side, h->lane_fork[k], 0);
  h->cur = h->lane_side;
  return true;
}
static void branch_end(xgc_handle* h, int k, bool on) { if (on) { cudaEventRecord(h->lane_join[k], h->lane_side); h->cur = h->lane_main; } }
static void branch_join(xgc_handle* h, int k, bool on) { if (on) cudaStreamWaitEvent(h->lane_main, h->lane_join[k], 0); }
#define LAUNCH(h, kern, grid, block, ...) do { prof_begin(h, #kern); kl(h, kern, grid, block, 0, __VA_ARGS__); prof_end(h); (h)->launches++; } while (0)
// kernels that draw: native instantiation unless an injected table is attached
#define LAUNCHD(h, inj, kern, grid, block, ...) do { prof_begin(h, #kern); \
  if ((inj).u) kl(h, kern<true>, grid, block, 0, __VA_ARGS__); else kl(h, kern<false>, grid, block, 0, __VA_ARGS__); \
  prof_end(h); (h)->launches++; } while (0)

// stable in-place edge compaction: small CTAs for calibration-size layers, 1024 x 4 for very long layers
static void launch_resim(xgc_handle* h, const Inj& inj, int dissolve, int kind) {
  const Dev& g = h->gl ? *h->gl : h->g;
  int emax = std::max(g.e_cap[0], std::max(g.e_cap[1], g.e_cap[2]));
  prof_begin(h, "k_edges_resim");
  if (emax > 16384) {           // very long layers: many CTAs per layer, decoupled look-back
    dim3 grid(g.lb_tiles, 3, g.R);
    if (inj.u) kl(h, k_edges_resim_lb<true>, grid, TPB, 0, g, inj, dissolve, kind);
    else kl(h, k_edges_resim_lb<false>, grid, TPB, 0, g, inj, dissolve, kind);
  } else {
    if (inj.u) kl(h, k_edges_resim<true, TPB, 1>, dim3(3, g.R), TPB, 0, g, inj, dissolve);
    else kl(h, k_edges_resim<false, TPB, 1>, dim3(3, g.R), TPB, 0, g, inj, dissolve);
  }
  prof_end(h); h->launches++;
}

// enqueue one week.  at_host < 0: week = previous + 1 (graph replay).
static int enqueue_week(xgc_handle* h, int at_host, unsigned mask, const Inj& inj, int zero_row) {
  const Dev& g = h->gl ? *h->gl : h->g;
  // agent kernels: tiles per CTA chosen so the grid stays a few waves of the 148 SMs
  int tiles = (g.n_cap + TPB - 1) / TPB;
  static const int TA = getenv("XGC_TUNE_A") ? atoi(getenv("XGC_TUNE_A")) : 16, TQ = getenv("XGC_TUNE_Q") ? atoi(getenv("XGC_TUNE_Q")) : 16,
                   TE = getenv("XGC_TUNE_E") ? atoi(getenv("XGC_TUNE_E")) : 32;      // CTAs per SM each kernel family is sized for
  // (few tiles in total, i.e. one or two large replicates: fewer, longer CTAs -- every CTA of an agent pass ends with a flush of its tallies into the SAME counter row,
  //  3907 one-tile CTAs of a 1M-agent replicate spend most of k_agent3a on ~200 same-address global atomics each)
  const int ta = (long long)tiles * g.R < 148LL * 16 * 4 ? std::min(TA, 4) : TA;      // (a single 1M-agent replicate: 3907 tiles)
  int tpc = std::max(1, std::min(std::min(tiles, 40), (int)(((long long)tiles * g.R) / (148 * ta))));     // <= 40 tiles: the CTA's queue staging stays under 48 KB
  dim3 ga((tiles + tpc - 1) / tpc, g.R);
  int emax = std::max(g.e_cap[0], std::max(g.e_cap[1], g.e_cap[2]));
  dim3 ge((emax + TPB - 1) / TPB, 3, g.R);
  // queue kernels: a few CTAs per replicate striding over its queue, ~2 waves of the GPU in total
  auto qgrid = [&](int max_tiles) { return (unsigned)std::max(1, std::min(max_tiles, (148 * TQ + g.R - 1) / g.R)); };
  prof_begin(h, "k_begin");
  if (g.n_cap > 32768) {                  // one CTA per large replicate: 16 warps share the arrivals
    if (inj.u) kl(h, k_begin<true, true>, g.R, 512, 0, g, inj, at_host, mask, zero_row); else kl(h, k_begin<false, true>, g.R, 512, 0, g, inj, at_host, mask, zero_row);
  } else {
    if (inj.u) kl(h, k_begin<true, false>, (g.R + 3) / 4, 128, 0, g, inj, at_host, mask, zero_row); else kl(h, k_begin<false, false>, (g.R + 3) / 4, 128, 0, g, inj, at_host, mask, zero_row);
  }
  prof_end(h); h->launches++;
  // zero_row marks the first call of a week: that is when k_agent1 takes the start-of-week snapshot bits
  const unsigned a1mods = XGC_M_AGING | XGC_M_DEPARTURE | XGC_M_HIVTEST | XGC_M_HIVTX | XGC_M_HIVPROGRESS | XGC_M_HIVVL;
  const size_t cq_smem = ((size_t)tpc * TPB + 2) * sizeof(unsigned);        // per-CTA queue staging of the agent passes
  if (zero_row || (mask & a1mods)) {
    prof_begin(h, "k_agent1a");
    if (inj.u) kl(h, k_agent1a<true>, ga, TPB, cq_smem, g, inj, mask, zero_row, tpc); else kl(h, k_agent1a<false>, ga, TPB, cq_smem, g, inj, mask, zero_row, tpc);
    prof_end(h); h->launches++;
  }
  // the HIV-positive queue touches only HIV fields of its agents: side branch, joined before the kernels that read them
  bool br0 = false;
  if (mask & (a1mods & ~(XGC_M_AGING | XGC_M_DEPARTURE))) {
    br0 = branch_begin(h, 0);
    LAUNCHD(h, inj, k_agent1b, dim3(qgrid(tiles), g.R), TPB, g, inj, mask);
    branch_end(h, 0, br0);
  }
  bool native = inj.u == nullptr;
  if ((mask & XGC_M_DEPARTURE) && !((mask & XGC_M_RESIM_NETS) && native)) launch_resim(h, inj, 0, 0);
  if (mask & XGC_M_RESIM_NETS) {
    // degrees of the ties that survive this week's dissolution: zeroed here, counted by the dissolution pass, read by k_form
    LAUNCH(h, k_zero_deg, (unsigned)std::min<size_t>(148 * 8, ((size_t)g.R * g.n_cap / 4 + TPB - 1) / TPB), TPB, g);
    launch_resim(h, inj, 1, 1);
    size_t form_smem = (size_t)(g.n_cap / 32) * 8 + (size_t)g.n_cap * 2;      // alive bitmap + popcount prefix + 16-bit position -> slot table of one replicate
    if (g.n_cap <= 32768) {            // n_cap <= 32k slots: one CTA per replicate with the bitmap in shared memory
      prof_begin(h, "k_form");
      if (inj.u) kl(h, k_form<true, true, 512>, g.R, 512, form_smem, g, inj); else kl(h, k_form<false, true, 512>, g.R, 512, form_smem, g, inj);
      prof_end(h); h->launches++;
    } else {
      LAUNCH(h, k_rank, dim3((g.n_cap + RANK_TILE - 1) / RANK_TILE, g.R), TPB, g);
      prof_begin(h, "k_form");
      const unsigned ft = (unsigned)(h->form_tiles > 0 ? h->form_tiles : g.lb_tiles);
      if (inj.u) kl(h, k_form_lb<true>, dim3(ft, 3, g.R), FORM_NT, 0, g, inj); else kl(h, k_form_lb<false>, dim3(ft, 3, g.R), FORM_NT, 0, g, inj);
      prof_end(h); h->launches++;
    }
  }
  branch_join(h, 0, br0);
  // edge kernels: the three layers as one index space, a few CTAs per replicate striding over it
  int etiles = (int)((g.e_stride + TPB - 1) / TPB);
  dim3 gef((unsigned)std::max(1, std::min(etiles, (148 * TE + g.R - 1) / g.R)), g.R);
  if (mask & (XGC_M_ACTS | XGC_M_PREP)) LAUNCHD(h, inj, k_edge1, gef, TPB, g, inj, mask);
  if (mask & (XGC_M_PREP | XGC_M_STIRECOV | XGC_M_STITX)) {
    prof_begin(h, "k_agent2a");
    if (inj.u) kl(h, k_agent2a<true>, ga, TPB, cq_smem, g, inj, mask, tpc); else kl(h, k_agent2a<false>, ga, TPB, cq_smem, g, inj, mask, tpc);
    prof_end(h); h->launches++;
  }
  if (mask & (XGC_M_STIRECOV | XGC_M_STITX)) LAUNCHD(h, inj, k_agent2b, dim3(qgrid(tiles), g.R), TPB, g, inj, mask);
  if (mask & (XGC_M_HIVTRANS | XGC_M_STITRANS)) {
    LAUNCH(h, k_edge2a, gef, TPB, g, mask);
    unsigned qb = qgrid((int)((g.e_stride + TPB - 1) / TPB));
    bool br1 = branch_begin(h, 1);          // anal acts on the side branch, oral / kissing / rimming on the main stream
    prof_begin(h, "k_edge2b_anal");
    if (inj.u) kl(h, k_edge2b<true, true>, dim3(qb, 1, g.R), TPB, 0, g, inj, mask); else kl(h, k_edge2b<false, true>, dim3(qb, 1, g.R), TPB, 0, g, inj, mask);
    prof_end(h); h->launches++;
    branch_end(h, 1, br1);
    prof_begin(h, "k_edge2b_oral_kiss_rim");
    if (inj.u) kl(h, k_edge2b<true, false>, dim3(qb, 3, g.R), TPB, 0, g, inj, mask); else kl(h, k_edge2b<false, false>, dim3(qb, 3, g.R), TPB, 0, g, inj, mask);
    prof_end(h); h->launches++;
    branch_join(h, 1, br1);
  }
  if (mask & (XGC_M_HIVTRANS | XGC_M_STITRANS | XGC_M_PREV)) { prof_begin(h, "k_agent3a"); kl(h, k_agent3a, ga, TPB, cq_smem, g, mask, tpc); prof_end(h); h->launches++; }
  if (mask & (XGC_M_HIVTRANS | XGC_M_STITRANS)) LAUNCHD(h, inj, k_agent3b, dim3(qgrid(tiles), g.R), TPB, g, inj, mask);
  return 0;
}

extern "C" int xgc_set_mode(xgc_handle* h, int mode) {
  if (mode != XGC_MODE_STRICT && mode != XGC_MODE_FAST) FAIL(h, 2, "unknown mode %d", mode);
  if (mode == XGC_MODE_FAST && h->fast_smem == 0)
    FAIL(h, 10, "XGC_MODE_FAST needs a replicate (n_cap = %d slots, %zu edge records) to fit one SM's shared memory", h->cfg.n_cap, h->g.e_stride);
  h->mode = mode;
  return 0;
}
extern "C" int xgc_get_mode(xgc_handle* h, int* mode) { *mode = h->mode; return 0; }

static int compact_impl(xgc_handle* h);
// the ties of departed agents are physically out of the lists after this (tergmLite::delete_vertices); no-op unless a fast-path
// call deferred it
static int prune_now(xgc_handle* h) {
  int rcf = flush_pending_edges(h); if (rcf) return rcf;
  if (!h->prune_pending) return 0;
  CU(h, cudaSetDevice(h->cfg.device));
  launch_resim(h, make_inj(h, nullptr, nullptr), 0, 0);
  h->prune_pending = false;
  CU(h, cudaGetLastError());
  return 0;
}
// module mask -> phase groups of the replicate-resident kernel (0: not a group the fast path runs)
static unsigned fast_phases(uint32_t mask) {
  const uint32_t PREG = XGC_M_AGING | XGC_M_DEPARTURE | XGC_M_ARRIVAL | XGC_M_HIVTEST | XGC_M_HIVTX | XGC_M_HIVPROGRESS | XGC_M_HIVVL;
  const uint32_t POSTG = XGC_M_ACTS | XGC_M_CONDOMS | XGC_M_POSITION | XGC_M_PREP | XGC_M_HIVTRANS | XGC_M_STIRECOV | XGC_M_STITX | XGC_M_STITRANS | XGC_M_PREV;
  if (mask == XGC_M_ALL) return XF_ALL;
  if (mask == PREG) return XF_PRE | XF_PRUNE;
  if (mask == (PREG | XGC_M_RESIM_NETS)) return XF_PRE | XF_PRUNE | XF_FORM;
  if (mask == POSTG) return XF_POST;
  if (mask == (POSTG | XGC_M_RESIM_NETS)) return XF_PRUNE | XF_FORM | XF_POST;
  return 0;
}
static int fast_enqueue(xgc_handle* h, int at0, int at1, unsigned phases, uint32_t mask, int write_acts, unsigned ph_last = 0u, int carry = -1) {
  if (h->fast_dirty) { CU(h, xgc_fast_prepare(h->g, h->stream)); h->fast_dirty = false; h->launches++; }
  FastArgs a; a.at0 = at0; a.at1 = at1; a.phases = phases; a.ph_last = ph_last ? ph_last : phases; a.mask = mask; a.write_acts = write_acts;
  a.carry = carry >= 0 ? h->d_carry : nullptr; a.carry_in = carry > 0;      // (span calls only)
  for (int l = 0; l < 3; ++l) {         // queued host-network operations ride along with a launch whose first week has an acts pass
    xgc_handle::PendEdge& q = h->pend[l];
    const bool take = q.kind && (phases & XF_POST) && !(phases & XF_PRE);      // (the host computed them after this week's aging..hivvl group)
    a.hn_kind[l] = take ? q.kind : 0; a.hn_ptr[l] = q.ptr; a.hn_ne[l] = q.ne; a.hn_start[l] = q.start; a.hn_stride[l] = q.stride; a.hn_has_start[l] = q.has_start;
    if (take) q.kind = 0;
  }
  prof_begin(h, "k_week_fast");
  cudaError_t e = xgc_fast_launch(h->g, a, h->fast_smem, h->stream);
  prof_end(h); h->launches++;
  CU(h, e);
  return 0;
}
extern "C" int xgc_step(xgc_handle* h, int at, uint32_t mask, const xgc_inject* inj) {
  if (at < 1 || at >= h->cfg.t_cap) FAIL(h, 2, "at = %d outside [1, t_cap = %d)", at, h->cfg.t_cap);
  if (at > 32766) FAIL(h, 2, "at = %d exceeds the 16-bit week clocks", at);
  h->carry_week = -1;
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  Inj j; rc = upload_inject(h, inj, &j); if (rc) return rc;
  h->last_inj = j; h->rank_fresh = false;
  if (at != h->last_at) h->acts_valid = false;          // a new week: last week's act counts are history
  if (mask & XGC_M_ACTS) h->acts_valid = true;
  int zero_row = at != h->last_at; h->last_at = at;
  const unsigned fph = (h->mode == XGC_MODE_FAST && !(inj && inj->u)) ? fast_phases(mask) : 0u;
  if (!fph && (mask & XGC_M_ARRIVAL)) h->fast_dirty = true;       // arrivals of the strict kernels carry no fast-path bits
  // slot compaction on the same schedule as xgc_run (after every week that is a multiple of compact_every), so that the
  // step-by-step boundary -- module(dat, at) calls, the host-network loop, the R shim -- never runs out of slots on a
  // reference-length burn-in.  Results do not depend on it: draws are keyed by uid, positions are preserved.
  if (zero_row && at >= 2 && (at - 1) % h->cfg.compact_every == 0 && h->compacted_at != at - 1) {
    rc = flush_pending_edges(h); if (rc) return rc;
    rc = compact_impl(h); if (rc) return rc;
    h->compacted_at = at - 1;
  }
  if (fph) {
    if (!(fph & XF_PRE) && zero_row) FAIL(h, 11, "fast mode: week %d must start with the aging..hivvl module group", at);
    h->acts_valid = false;                                        // kissing / rimming counts are not materialised in fast mode
    if (!(fph & XF_POST) || (fph & XF_PRE)) { rc = flush_pending_edges(h); if (rc) return rc; }
    rc = fast_enqueue(h, at, at, fph, mask, 1);
    // a call without the acts pass and without an on-device network step leaves the departed agents' ties to the next toucher
    if (fph & XF_POST) h->prune_pending = false; else if (!(fph & XF_FORM)) h->prune_pending = true;
    return rc;
  }
  rc = prune_now(h); if (rc) return rc;
  rc = enqueue_week(h, at, mask, j, zero_row); if (rc) return rc;
  CU(h, cudaGetLastError());
  if (inj && inj->u) CU(h, cudaStreamSynchronize(h->stream));     // caller may free the table after return
  return 0;
}

// The rest of week `at` (mask_tail: the acts..prevalence group, with or without resim_nets) and the start of week at + 1
// (mask_head: the aging..hivvl group) in ONE call.  Nothing on the host sits between prevalence_msm(at) and aging_msm(at + 1)
// in the reference's netsim loop, and every module is a function of (state, at, seed): the result is that of the two xgc_step
// calls.  On the fast path the two groups run in one launch of the resident kernel -- the replicate is staged once, written
// back once, and the host-network loop pays one launch and one device -> host wait per week instead of two.
extern "C" int xgc_step_span(xgc_handle* h, int at, uint32_t mask_tail, uint32_t mask_head) {
  if (at < 1 || at + 1 >= h->cfg.t_cap || at + 1 > 32766) FAIL(h, 2, "xgc_step_span: weeks %d and %d must lie in [1, t_cap = %d)", at, at + 1, h->cfg.t_cap);
  const unsigned pt = h->mode == XGC_MODE_FAST ? fast_phases(mask_tail) : 0u, phd = h->mode == XGC_MODE_FAST ? fast_phases(mask_head) : 0u;
  const bool compaction_due = at >= 1 && at % h->cfg.compact_every == 0 && h->compacted_at != at;      // xgc_step(at + 1) would compact first
  if (!(pt & XF_POST) || (pt & XF_PRE) || phd != (unsigned)(XF_PRE | XF_PRUNE) || h->last_at != at || compaction_due) {
    int rc = xgc_step(h, at, mask_tail, nullptr);
    return rc ? rc : xgc_step(h, at + 1, mask_head, nullptr);
  }
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  h->last_inj = make_inj(h, nullptr, nullptr); h->rank_fresh = false; h->acts_valid = false; h->last_at = at + 1;
  // the previous span left week `at` half done with its running state tallies in d_carry: no recount when nothing touched the agents since
  rc = fast_enqueue(h, at, at + 1, pt, mask_tail | mask_head, 1, XF_PRE | XF_PRUNE, h->carry_week == at ? 1 : 0);
  h->carry_week = rc ? -1 : at + 1;
  h->prune_pending = true;              // the ties of the agents who departed in week at + 1 are left to the next toucher
  return rc;
}

// slot compaction of the replicates of the current lane (h->gl) or of the whole handle
static int compact_impl(xgc_handle* h) {
  const Dev& g = h->gl ? *h->gl : h->g;
  int emax = std::max(g.e_cap[0], std::max(g.e_cap[1], g.e_cap[2]));
  launch_resim(h, make_inj(h, nullptr, nullptr), 0, 2);       // no dangling endpoints before re-numbering
  h->rank_fresh = false;
  if (g.n_cap <= 32768) {
    LAUNCH(h, k_compact_agents, g.R, TPB, g, h->d_remap);       // one CTA walks the replicate's ~40 tiles
  } else {                                                      // 100k .. 1M-agent replicates: parallel gathers through pos2slot
    dim3 gp((g.n_cap + TPB - 1) / TPB, g.R);
    LAUNCH(h, k_rank, dim3((g.n_cap + RANK_TILE - 1) / RANK_TILE, g.R), TPB, g);
    LAUNCH(h, k_compact_remap, gp, TPB, g, h->d_remap);
    auto move = [&](auto* plane) {
      using T = std::remove_pointer_t<decltype(plane)>;
      LAUNCH(h, k_compact_gather<T>, gp, TPB, g, (const T*)plane, (T*)h->d_cscratch);
      LAUNCH(h, k_compact_putback<T>, gp, TPB, g, plane, (const T*)h->d_cscratch);
    };
    move(g.core); move((uint4*)g.aux); move((uint2*)g.gck_a); move((uint2*)g.gck_b); move((uint2*)g.gck_d);
    move((uint2*)g.hck_a); move((uint2*)g.hck_b); move((uint2*)g.tck); move(g.arr_t);
    LAUNCH(h, k_compact_finish, gp, TPB, g, 0);
    LAUNCH(h, k_compact_finish, dim3(1, g.R), TPB, g, 1);
  }
  LAUNCH(h, k_compact_edges, dim3((emax + TPB - 1) / TPB, 3, g.R), TPB, g, h->d_remap);
  CU(h, cudaGetLastError());
  return 0;
}

extern "C" int xgc_compact(xgc_handle* h) {
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  rc = flush_pending_edges(h); if (rc) return rc;
  rc = compact_impl(h);
  if (!rc) { h->compacted_at = h->last_at; h->prune_pending = false; }
  return rc;
}

extern "C" int xgc_run(xgc_handle* h, int at0, int at1) {
  if (at0 < 1 || at1 >= h->cfg.t_cap || at1 < at0 - 1) FAIL(h, 2, "bad week range [%d,%d], t_cap = %d", at0, at1, h->cfg.t_cap);
  if (at1 > 32766) FAIL(h, 2, "at exceeds the 16-bit week clocks");
  h->carry_week = -1;
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  rc = prune_now(h); if (rc) return rc;
  if (h->mode == XGC_MODE_FAST) {       // replicate-resident kernel: one launch per stretch of weeks between slot compactions
    const int ce = h->cfg.compact_every;
    for (int at = at0; at <= at1;) {
      int seg = std::min(at1, ((at + ce - 1) / ce) * ce);
      rc = fast_enqueue(h, at, seg, XF_ALL, XGC_M_ALL, seg == at1); if (rc) return rc;
      if (seg % ce == 0) { rc = compact_impl(h); if (rc) return rc; h->compacted_at = seg; }
      at = seg + 1;
    }
    h->last_at = at1; h->rank_fresh = false; h->acts_valid = false; memset(&h->last_inj, 0, sizeof(Inj));
    CU(h, cudaGetLastError());
    return 0;
  }
  h->fast_dirty = true;
  Inj j = make_inj(h, nullptr, nullptr);
  const int L = h->n_lanes, R = h->cfg.n_replicates;
  Dev* gl = h->lane_dev; gl[0] = h->g; gl[1] = h->g;
  if (L == 2) { gl[0].r0 = 0; gl[0].R = R / 2; gl[1].r0 = R / 2; gl[1].R = R - R / 2; }
  auto use_lane = [&](int k) {             // point the enqueue helpers at lane k's Dev, streams and events
    h->gl = &gl[k];
    h->lane_main = k ? h->l1_stream : h->stream; h->lane_side = k ? h->l1_stream2 : h->stream2;
    h->lane_fork = k ? h->l1_fork : h->ev_fork; h->lane_join = k ? h->l1_join : h->ev_join;
    h->cur = h->lane_main;
  };
  auto use_whole = [&]() { h->gl = nullptr; h->lane_main = h->stream; h->lane_side = h->stream2; h->lane_fork = h->ev_fork; h->lane_join = h->ev_join; h->cur = h->stream; };
  if (h->lanes_captured != L) {            // capture one week (week = previous + 1) per lane once, replay it for every later week
    drop_graphs(h);
    for (int k = 0; k < L; ++k) {
      cudaGraph_t graph;
      uint64_t l0 = h->launches;
      use_lane(k);
      { cudaError_t eb = cudaStreamBeginCapture(h->lane_main, cudaStreamCaptureModeThreadLocal); if (eb != cudaSuccess) { use_whole(); CU(h, eb); } }
      h->capturing = true;
      enqueue_week(h, -1, XGC_M_ALL, j, 1);
      h->capturing = false;
      cudaError_t e = cudaStreamEndCapture(h->lane_main, &graph);
      if (e != cudaSuccess) { use_whole(); CU(h, e); }
      e = cudaGraphInstantiate(k ? &h->l1_graph : &h->week_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { use_whole(); CU(h, e); }
      (k ? h->l1_graph_launches : h->graph_launches) = h->launches - l0;
      h->launches = l0;
    }
    h->lanes_captured = L;
    use_whole();
  }
  // the graph advances the device-side week itself; make sure it continues from at0 - 1
  std::vector<int> atv(h->cfg.n_replicates, at0 - 1);
  CU(h, cudaMemcpyAsync((int*)h->g.at_ptr, atv.data(), sizeof(int) * atv.size(), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (L == 2) {                            // lane 1 starts after everything already queued on the handle's stream
    CU(h, cudaEventRecord(h->l1_start, h->stream));
    CU(h, cudaStreamWaitEvent(h->l1_stream, h->l1_start, 0));
  }
  for (int at = at0; at <= at1; ++at) {
    CU(h, cudaGraphLaunch(h->week_graph, h->stream));
    h->launches += h->graph_launches;
    if (L == 2) { CU(h, cudaGraphLaunch(h->l1_graph, h->l1_stream)); h->launches += h->l1_graph_launches; }
    if (at % h->cfg.compact_every == 0) {
      for (int k = 0; k < L; ++k) { use_lane(k); rc = compact_impl(h); if (rc) break; }
      use_whole();
      if (rc) return rc;
      h->compacted_at = at;
    }
  }
  if (L == 2) {                            // ... and the handle's stream continues after lane 1
    CU(h, cudaEventRecord(h->l1_done, h->l1_stream));
    CU(h, cudaStreamWaitEvent(h->stream, h->l1_done, 0));
  }
  h->last_at = at1; h->rank_fresh = false; h->acts_valid = true; memset(&h->last_inj, 0, sizeof(Inj));
  CU(h, cudaGetLastError());
  return 0;
}

// per-kernel device timing with CUDA events on the launching stream (bench.py roofline leg)
extern "C" int xgc_profile(xgc_handle* h, int enable) {
  CU(h, cudaSetDevice(h->cfg.device));
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->prof.clear();
  h->profiling = enable != 0;
  return 0;
}
extern "C" int xgc_profile_read(xgc_handle* h, const char** names, double* ms, uint64_t* count, int cap) {
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  int n = 0;
  for (auto& r : h->prof) {
    float t = 0.f; if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    int k = 0; for (; k < n; ++k) if (!strcmp(names[k], r.name)) break;
    if (k == n) { if (n == cap) continue; names[n] = r.name; ms[n] = 0.0; count[n] = 0; n++; }
    ms[k] += (double)t; count[k]++;
  }
  return n;     // number of distinct kernels (not an error code)
}
extern "C" int xgc_num_agents_all(xgc_handle* h, int32_t* out /*[R]*/) {
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaMemcpy2DAsync(out, sizeof(int), &h->g.rep[0].n_alive, sizeof(Rep), sizeof(int), (size_t)h->g.R, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int xgc_status(xgc_handle* h, int r, uint32_t* status) {
  CHECK_R(h, r); CU(h, cudaSetDevice(h->cfg.device));
  { int rcf = flush_pending_edges(h); if (rcf) return rcf; }      // queued edge-list operations report bad / overflowing lists here too
  CU(h, cudaStreamSynchronize(h->stream));
  Rep rp; CU(h, d2h(h, &rp, h->g.rep + r, sizeof(Rep)));
  unsigned ps[4]; CU(h, d2h(h, ps, h->d_pstat + (size_t)r * 4, sizeof ps));
  *status = rp.status | ps[0] | ps[1] | ps[2] | ps[3]; return 0;
}

static int materialise_actcounts(xgc_handle* h, int r, const Inj& j) {      // kissing / rimming counts are lazy on the hot path
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  const Dev& g = h->g;
  int emax = std::max(g.e_cap[0], std::max(g.e_cap[1], g.e_cap[2]));
  int* d_out; CU(h, cudaMalloc((void**)&d_out, sizeof(int) * 3 * (size_t)emax));
  cudaMemsetAsync(d_out, 0, sizeof(int) * 3 * (size_t)emax, h->stream);
  h->plain_next = true;
  LAUNCHD(h, j, k_actcounts, dim3((emax + TPB - 1) / TPB, 3), TPB, g, j, r, d_out, emax);
  std::vector<int> all(3 * (size_t)emax);
  cudaError_t e = cudaMemcpyAsync(all.data(), d_out, sizeof(int) * all.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) FAIL(h, 100 + (int)e, "xgc_get_actcounts: %s", cudaGetErrorString(e));
  rc = cache_load(h, r); if (rc) return rc;
  for (int l = 0; l < 3; ++l) h->cache.lazy[l].assign(all.begin() + (size_t)l * emax, all.begin() + (size_t)l * emax + h->cache.edge[l].size());
  return 0;
}
extern "C" int xgc_get_actcounts(xgc_handle* h, int r, int layer, int32_t* counts, int cap, int* n_edges) {
  CHECK_R(h, r);
  if (layer < 0 || layer > 2) FAIL(h, 2, "layer %d out of range", layer);
  int rc = 0;
  bool lazy_ok = h->cache.r == r && h->cache.lazy[layer].size() == h->cache.edge[layer].size() && !h->cache.edge[layer].empty();
  if (h->acts_valid && !lazy_ok && !(h->cache.r == r && h->cache.dirty)) { rc = materialise_actcounts(h, r, h->last_inj); if (rc) return rc; }
  rc = cache_load(h, r); if (rc) return rc;
  const HostRep& c = h->cache; int ne = (int)c.edge[layer].size();
  if (n_edges) *n_edges = ne;
  if (cap < ne) FAIL(h, 6, "buffer too small");
  bool have = h->acts_valid && c.lazy[layer].size() == (size_t)ne;
  for (int e = 0; e < ne; ++e) {
    int w = c.edge[layer][e].w, z = have ? c.lazy[layer][e] : 0;
    counts[e] = w & 0xFF; counts[ne + e] = (w >> 8) & 0xFF; counts[2 * ne + e] = z & 0xFF; counts[3 * ne + e] = (z >> 8) & 0xFF;
  }
  return 0;
}
extern "C" int xgc_get_actlist(xgc_handle* h, int r, int at, int32_t* al, int cap, int* n_acts, const xgc_inject* inj) {
  CHECK_R(h, r);
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  Inj j; rc = upload_inject(h, inj, &j); if (rc) return rc;
  int* d_al; CU(h, cudaMalloc((void**)&d_al, sizeof(int) * 6 * (size_t)std::max(cap, 1)));
  CU(h, cudaMemcpyAsync((int*)h->g.at_ptr + r, &at, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  h->plain_next = true;
  LAUNCHD(h, j, k_actlist, 1, 32, h->g, j, r, d_al, cap, h->d_scratch);
  int n = 0;
  cudaError_t e1 = cudaMemcpyAsync(&n, h->d_scratch, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->stream);
  if (e1 == cudaSuccess && e2 == cudaSuccess && n <= cap) e1 = d2h(h, al, d_al, sizeof(int) * 6 * (size_t)cap);
  cudaFree(d_al);
  if (e1 != cudaSuccess || e2 != cudaSuccess) FAIL(h, 100, "xgc_get_actlist: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  if (n_acts) *n_acts = n;
  if (n > cap) FAIL(h, 6, "act list has %d rows, buffer holds %d", n, cap);
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// outputs: dat$epi
// --------------------------------------------------------------------------------------------------------------------
extern "C" int xgc_get_counters(xgc_handle* h, int r, int at0, int at1, uint32_t* out) {
  CHECK_R(h, r);
  if (at0 < 0 || at1 >= h->cfg.t_cap || at1 < at0) FAIL(h, 2, "bad week range");
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  CU(h, d2h(h, out, h->g.counters + ((size_t)r * h->g.t_cap + at0) * XGC_NCOUNTER, sizeof(uint32_t) * XGC_NCOUNTER * (size_t)(at1 - at0 + 1)));
  return 0;
}
extern "C" int xgc_get_counters_all(xgc_handle* h, int at, uint32_t* out) {
  if (at < 0 || at >= h->cfg.t_cap) FAIL(h, 2, "bad week");
  CU(h, cudaSetDevice(h->cfg.device));
  CU(h, cudaMemcpy2DAsync(out, sizeof(uint32_t) * XGC_NCOUNTER, h->g.counters + (size_t)at * XGC_NCOUNTER, sizeof(uint32_t) * XGC_NCOUNTER * (size_t)h->g.t_cap,
                          sizeof(uint32_t) * XGC_NCOUNTER, (size_t)h->g.R, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

static double hsum(const uint32_t* c, int group, uint32_t cells) {
  double s = 0.0;
  for (int k = 0; k < XGC_NCELL; ++k) if ((cells >> k) & 1u) s += (double)c[group * XGC_NCELL + k];
  return s;
}
// stratum suffix -> cell mask: "" | B/H/O/W | age1 / age.1 / 1 | B.age1
static uint32_t stratum(const std::string& s) {
  if (s.empty()) return 0xFFFFFu;
  int race = -1, age = -1; size_t p = 0;
  static const std::string R = "BHOW";
  if (R.find(s[0]) != std::string::npos && (s.size() == 1 || s[1] == '.')) { race = (int)R.find(s[0]); p = s.size() == 1 ? 1 : 2; }
  if (p < s.size()) {
    std::string t = s.substr(p);
    if (t.rfind("age.", 0) == 0) t = t.substr(4); else if (t.rfind("age", 0) == 0) t = t.substr(3);
    if (t.size() == 1 && t[0] >= '1' && t[0] <= '5') age = t[0] - '1'; else return 0;
  }
  uint32_t m = 0;
  for (int r = 0; r < 4; ++r) for (int a = 0; a < 5; ++a) if ((race < 0 || race == r) && (age < 0 || age == a)) m |= 1u << (r * 5 + a);
  return m;
}
static bool split(const std::string& v, const char* base, std::string* rest) {
  std::string b(base);
  if (v == b) { rest->clear(); return true; }
  if (v.rfind(b + ".", 0) == 0) { *rest = v.substr(b.size() + 1); return true; }
  return false;
}
static bool epi_eval(const uint32_t* c, const std::string& v, double* out) {
  static const char* const pathn[XGC_NPATH] = { "incid.u2rgc", "incid.p2rgc", "incid.r2ugc", "incid.p2ugc", "incid.u2pgc", "incid.r2pgc", "incid.p2pgc" };
  static const char* const tst[3] = { "num.rgc.test", "num.ugc.test", "num.pgc.test" };
  static const char* const propt[3] = { "prop.rect.tested", "prop.ureth.tested", "prop.phar.tested" };
  static const char* const probt[3] = { "prob.rGC.tested", "prob.uGC.tested", "prob.pGC.tested" };
  static const char* const sites[4] = { "gc", "rgc", "ugc", "pgc" };
  for (int k = 0; k < XGC_NPATH; ++k) if (v == pathn[k]) { *out = (double)c[XGC_K_PATH + k]; return true; }
  for (int s = 0; s < 3; ++s) {
    if (v == tst[s]) { *out = (double)c[XGC_K_TESTED + s]; return true; }
    if (v == propt[s]) { *out = (double)c[XGC_K_TESTED + s] / (double)c[XGC_K_VISITS]; return true; }
    if (v == probt[s]) { *out = (double)c[XGC_K_POS + s] / (double)c[XGC_K_TESTED + s]; return true; }
  }
  if (v == "num.visits") { *out = (double)c[XGC_K_VISITS]; return true; }
  if (v == "dep.gen") { *out = (double)c[XGC_K_DEP_GEN]; return true; }
  if (v == "dep.AIDS") { *out = (double)c[XGC_K_DEP_AIDS]; return true; }
  if (v == "nNew") { *out = (double)c[XGC_K_NNEW]; return true; }
  std::string rest; uint32_t cells;
  for (int s = 0; s < 4; ++s) {                  // GC series: <kind>[.<stratum>].<site>
    std::string suf = std::string(".") + sites[s];
    if (v.size() <= suf.size() || v.compare(v.size() - suf.size(), suf.size(), suf) != 0) continue;
    std::string head = v.substr(0, v.size() - suf.size());
    int kind = split(head, "i.num", &rest) ? 0 : split(head, "prev", &rest) ? 1 : split(head, "ir100", &rest) ? 3 : split(head, "incid", &rest) ? 2 : -1;
    if (kind < 0 || !(cells = stratum(rest))) continue;
    double num = hsum(c, XGC_K_NUM, cells), inum, incid;
    if (s == 0) { inum = hsum(c, XGC_K_INUM_GC, cells); incid = hsum(c, XGC_K_INCID_RGC, cells) + hsum(c, XGC_K_INCID_UGC, cells) + hsum(c, XGC_K_INCID_PGC, cells); }
    else { inum = hsum(c, XGC_K_INUM_RGC + s - 1, cells); incid = hsum(c, XGC_K_INCID_RGC + s - 1, cells); }
    *out = kind == 0 ? inum : kind == 1 ? inum / num : kind == 2 ? incid : incid / (num - inum) * 5200.0;
    return true;
  }
  if (split(v, "num", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_NUM, cells); return true; }
  if (split(v, "i.num.dx", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INUM_DX, cells); return true; }
  if (split(v, "i.num", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INUM, cells); return true; }
  if (split(v, "i.prev.dx", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INUM_DX, cells) / hsum(c, XGC_K_NUM, cells); return true; }
  if (split(v, "i.prev", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INUM, cells) / hsum(c, XGC_K_NUM, cells); return true; }
  if (split(v, "incid", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INCID_HIV, cells); return true; }
  if (split(v, "cc.vsupp", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_VSUPP, cells) / hsum(c, XGC_K_INUM_DX, cells); return true; }
  if (split(v, "ir100", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_INCID_HIV, cells) / (hsum(c, XGC_K_NUM, cells) - hsum(c, XGC_K_INUM, cells)) * 5200.0; return true; }
  if (split(v, "prepElig", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_PREPELIG, cells); return true; }
  if (split(v, "prepCurr", &rest) && (cells = stratum(rest))) { *out = hsum(c, XGC_K_PREPCURR, cells); return true; }
  return false;
}
extern "C" int xgc_get_epi(xgc_handle* h, int r, const char* var, double* out, int at0, int at1) {
  CHECK_R(h, r);
  if (at0 < 0 || at1 >= h->cfg.t_cap || at1 < at0) FAIL(h, 2, "bad week range");
  std::vector<uint32_t> c((size_t)(at1 - at0 + 1) * XGC_NCOUNTER);
  int rc = xgc_get_counters(h, r, at0, at1, c.data()); if (rc) return rc;
  for (int at = at0; at <= at1; ++at)
    if (!epi_eval(c.data() + (size_t)(at - at0) * XGC_NCOUNTER, var, out + (at - at0))) FAIL(h, 5, "unknown epi variable '%s'", var);
  return 0;
}
extern "C" int xgc_get_epi_table(xgc_handle* h, int r, const char* const* vars, int n_vars, double* out, int at0, int at1) {
  CHECK_R(h, r);
  if (at0 < 0 || at1 >= h->cfg.t_cap || at1 < at0) FAIL(h, 2, "bad week range");
  if (n_vars < 0 || (n_vars > 0 && !vars)) FAIL(h, 2, "bad variable list");
  size_t T = (size_t)(at1 - at0 + 1);
  std::vector<uint32_t> c(T * XGC_NCOUNTER);
  int rc = xgc_get_counters(h, r, at0, at1, c.data()); if (rc) return rc;
  for (int k = 0; k < n_vars; ++k) {
    std::string v(vars[k]);
    for (size_t t = 0; t < T; ++t)
      if (!epi_eval(c.data() + t * XGC_NCOUNTER, v, out + (size_t)k * T + t)) FAIL(h, 5, "unknown epi variable '%s'", vars[k]);
  }
  return 0;
}
static std::vector<std::string> build_epi_names() {
  std::vector<std::string> N;
  const char* R[4] = { "B", "H", "O", "W" };
  auto strat = [&](const std::string& b, bool dot_age) {
    N.push_back(b);
    for (int r = 0; r < 4; ++r) N.push_back(b + "." + R[r]);
    for (int a = 1; a <= 5; ++a) N.push_back(b + (dot_age ? ".age." : ".age") + std::to_string(a));
  };
  strat("num", true); strat("i.num", false); strat("i.prev", true); strat("i.num.dx", false); strat("i.prev.dx", true);
  strat("incid", false); strat("cc.vsupp", false); strat("ir100", false);
  for (const char* b : { "prepElig", "prepCurr" }) { N.push_back(b); for (int r = 0; r < 4; ++r) N.push_back(std::string(b) + "." + R[r]); }
  for (int r = 0; r < 4; ++r) for (int a = 1; a <= 5; ++a) { N.push_back(std::string("i.num.") + R[r] + ".age" + std::to_string(a)); N.push_back(std::string("i.num.dx.") + R[r] + ".age" + std::to_string(a)); }
  for (const char* s : { "gc", "rgc", "ugc", "pgc" }) for (const char* k : { "i.num", "prev", "incid", "ir100" }) N.push_back(std::string(k) + "." + s);
  for (const char* s : { "rgc", "ugc", "pgc" }) {
    for (int r = 0; r < 4; ++r) N.push_back(std::string("incid.") + R[r] + "." + s);
    for (int a = 1; a <= 5; ++a) N.push_back(std::string("incid.age.") + std::to_string(a) + "." + s);
  }
  for (const char* p : { "incid.u2rgc", "incid.p2rgc", "incid.r2ugc", "incid.p2ugc", "incid.u2pgc", "incid.r2pgc", "incid.p2pgc",
                         "num.rgc.test", "num.ugc.test", "num.pgc.test", "prop.rect.tested", "prop.ureth.tested", "prop.phar.tested",
                         "prob.rGC.tested", "prob.uGC.tested", "prob.pGC.tested", "num.visits", "dep.gen", "dep.AIDS", "nNew" }) N.push_back(p);
  return N;
}
extern "C" int xgc_epi_names(const char** names, int cap) {
  static std::vector<std::string> N = build_epi_names();
  for (int k = 0; k < (int)N.size() && k < cap; ++k) names[k] = N[k].c_str();
  return (int)N.size();
}
extern "C" int xgc_targets(xgc_handle* h, int at_last, int tailn, double* out) {
  if (tailn < 1 || at_last - tailn + 1 < 0 || at_last >= h->cfg.t_cap) FAIL(h, 2, "bad tail window");
  CU(h, cudaSetDevice(h->cfg.device));
  int n = h->g.R * XGC_NTARGET;
  LAUNCH(h, k_targets, (n + 127) / 128, 128, h->g, at_last, tailn, h->d_targets);
  CU(h, cudaMemcpyAsync(out, h->d_targets, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// checkpoint / fork
// --------------------------------------------------------------------------------------------------------------------
struct Seg { void* p; size_t bytes; };
static std::vector<Seg> segments(xgc_handle* h) {
  const Dev& g = h->g; size_t NA = h->n_agent();
  return { { g.core, NA * sizeof(uint4) }, { g.aux, NA * sizeof(Aux) }, { g.gck_a, NA * 8 }, { g.gck_b, NA * 8 }, { g.gck_d, NA * 8 },
           { g.hck_a, NA * 8 }, { g.hck_b, NA * 8 }, { g.tck, NA * 8 }, { g.arr_t, NA * 4 }, { g.alive, NA / 8 },
           { g.edge, (size_t)g.R * g.e_stride * sizeof(int4) }, { g.rep, (size_t)g.R * sizeof(Rep) }, { (void*)g.at_ptr, (size_t)g.R * 4 },
           { (void*)g.params, (size_t)g.R * XGC_NPARAM * 8 }, { (void*)g.control, (size_t)g.R * XGC_NCONTROL * 4 }, { (void*)g.tables, (size_t)XGC_NTABLE * 8 },
           { g.counters, (size_t)g.R * g.t_cap * XGC_NCOUNTER * 4 } };
}
extern "C" int xgc_snapshot_size(xgc_handle* h, size_t* bytes) {
  size_t s = sizeof(xgc_config);
  for (auto& sg : segments(h)) s += sg.bytes;
  *bytes = s; return 0;
}
extern "C" int xgc_snapshot(xgc_handle* h, void* buf, size_t bytes) {
  size_t need; xgc_snapshot_size(h, &need);
  if (bytes < need) FAIL(h, 6, "snapshot buffer too small: %zu < %zu", bytes, need);
  int rc = cache_drop(h); if (rc) return rc;
  rc = prune_now(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  char* p = (char*)buf; memcpy(p, &h->cfg, sizeof(xgc_config)); p += sizeof(xgc_config);
  for (auto& sg : segments(h)) { CU(h, d2h(h, p, sg.p, sg.bytes)); p += sg.bytes; }
  return 0;
}
extern "C" int xgc_restore(xgc_handle* h, const void* buf, size_t bytes) {
  size_t need; xgc_snapshot_size(h, &need);
  if (bytes < need) FAIL(h, 6, "snapshot buffer too small");
  const xgc_config* c = (const xgc_config*)buf;
  if (c->n_replicates != h->cfg.n_replicates || c->n_cap != h->cfg.n_cap || c->t_cap != h->cfg.t_cap ||
      c->e_cap[0] != h->cfg.e_cap[0] || c->e_cap[1] != h->cfg.e_cap[1] || c->e_cap[2] != h->cfg.e_cap[2]) FAIL(h, 9, "snapshot geometry differs from handle");
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device)); CU(h, cudaStreamSynchronize(h->stream));
  const char* p = (const char*)buf + sizeof(xgc_config);
  for (auto& sg : segments(h)) { CU(h, h2d(h, sg.p, p, sg.bytes)); p += sg.bytes; }
  h->last_at = -1; h->compacted_at = -1; h->acts_valid = false; h->rank_fresh = false; h->fast_dirty = true; h->carry_week = -1;
  { double t[XGC_NTABLE]; CU(h, d2h(h, t, h->g.tables, sizeof t));
    for (int l = 0; l < 2; ++l) h->g.inv_dur[l] = 1.0 / t[XGC_T_NET_DUR + l];
    set_form_tiles(h, t);
    drop_graphs(h); }
  return derive(h);
}
extern "C" int xgc_fork(xgc_handle* src, int sr, xgc_handle* dst, int dr) {
  CHECK_R(src, sr); CHECK_R(dst, dr);
  if (src->cfg.n_cap != dst->cfg.n_cap || src->g.e_stride != dst->g.e_stride || src->cfg.t_cap != dst->cfg.t_cap) FAIL(dst, 9, "fork: geometry differs");
  int rc = cache_drop(src); if (rc) return rc;
  rc = cache_drop(dst); if (rc) return rc;
  rc = prune_now(src); if (rc) return rc;
  CU(src, cudaSetDevice(src->cfg.device)); CU(src, cudaStreamSynchronize(src->stream));
  CU(dst, cudaSetDevice(dst->cfg.device)); CU(dst, cudaStreamSynchronize(dst->stream));
  const Dev &a = src->g, &b = dst->g; size_t n = a.n_cap, so = (size_t)sr * n, doff = (size_t)dr * n;
  auto cp = [&](void* d, const void* s, size_t bytes) { return cudaMemcpyAsync(d, s, bytes, cudaMemcpyDefault, dst->stream); };
  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  acc(cp(b.core + doff, a.core + so, n * sizeof(uint4))); acc(cp(b.aux + doff, a.aux + so, n * sizeof(Aux)));
  acc(cp(b.gck_a + doff, a.gck_a + so, n * 8)); acc(cp(b.gck_b + doff, a.gck_b + so, n * 8)); acc(cp(b.gck_d + doff, a.gck_d + so, n * 8));
  acc(cp(b.hck_a + doff, a.hck_a + so, n * 8)); acc(cp(b.hck_b + doff, a.hck_b + so, n * 8)); acc(cp(b.tck + doff, a.tck + so, n * 8));
  acc(cp(b.arr_t + doff, a.arr_t + so, n * 4)); acc(cp(b.alive + doff / 32, a.alive + so / 32, n / 8));
  acc(cp(b.edge + (size_t)dr * b.e_stride, a.edge + (size_t)sr * a.e_stride, a.e_stride * sizeof(int4)));
  acc(cp(b.rep + dr, a.rep + sr, sizeof(Rep))); acc(cp((int*)b.at_ptr + dr, a.at_ptr + sr, 4));
  acc(cp(b.counters + (size_t)dr * b.t_cap * XGC_NCOUNTER, a.counters + (size_t)sr * a.t_cap * XGC_NCOUNTER, (size_t)a.t_cap * XGC_NCOUNTER * 4));
  if (e != cudaSuccess) FAIL(dst, 100 + (int)e, "fork copy failed: %s", cudaGetErrorString(e));
  // (the copies run on the destination handle's stream; the source handle's stream was drained above)
  CU(dst, cudaStreamSynchronize(dst->stream));
  dst->last_at = -1; dst->carry_week = -1; dst->rank_fresh = false; dst->acts_valid = false; memset(&dst->last_inj, 0, sizeof(Inj));
  if (src->fast_dirty) dst->fast_dirty = true;
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// bulk edge-list upload for the weekly host<->device contract (host network resimulation)
// --------------------------------------------------------------------------------------------------------------------
__global__ void k_edges_upload(const __grid_constant__ Dev g, int l, const int* el /*[R][2][stride]*/, const int* ne /*[R]*/, const int* start /*[R][stride] or null*/, int stride) {
  pdl_prologue();         // launched with programmatic stream serialization right after k_rank: must wait for pos2slot
  int r = blockIdx.y + g.r0, e = blockIdx.x * blockDim.x + threadIdx.x;
  int n = ne[r], cap = g.e_cap[l] < stride ? g.e_cap[l] : stride;
  if (e == 0) { g.rep[r].ne[l] = n < cap ? n : cap; if (n > cap) atomicOr(&g.rep[r].status, XGC_ST_EDGE_OVERFLOW); }
  if (e >= n || e >= cap) return;
  const int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  int na = g.rep[r].n_alive;
  int a = el[((size_t)r * 2 + 0) * stride + e], b = el[((size_t)r * 2 + 1) * stride + e];
  if (a < 1 || a > na || b < 1 || b > na) { atomicOr(&g.rep[r].status, XGC_ST_BAD_EDGE); a = b = 1; }
  g.edge[(size_t)r * g.e_stride + g.e_off[l] + e] = make_int4(P2S[a - 1], P2S[b - 1], start ? start[(size_t)r * stride + e] : g.at_ptr[r], 0);
}
extern "C" int xgc_set_edgelists_all(xgc_handle* h, int layer, const int32_t* el, const int32_t* n_edges, const int32_t* start, int stride) {
  if (layer < 0 || layer > 2) FAIL(h, 2, "layer out of range");
  if (stride < 1) FAIL(h, 2, "stride must be >= 1");
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  const Dev& g = h->g; size_t st = (size_t)stride, R = g.R;
  size_t need = R * st * 3 + R;
  if (h->pend[layer].kind || h->pend[0].kind == 2 || h->pend[1].kind == 2 || h->pend[2].kind == 2) {      // one staging buffer for whole lists
    rc = flush_pending_edges(h); if (rc) return rc;
  }
  if (need > h->el_stage_cap) {
    if (h->d_el_stage) cudaFree(h->d_el_stage);
    h->d_el_stage = nullptr; h->el_stage_cap = 0;
    CU(h, cudaMalloc((void**)&h->d_el_stage, (need + need / 4) * sizeof(int)));
    h->el_stage_cap = need + need / 4;
  }
  int* d_el = h->d_el_stage; int* d_start = d_el + R * st * 2; int* d_ne = d_start + R * st;
  CU(h, cudaMemcpyAsync(d_el, el, R * st * 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (start) CU(h, cudaMemcpyAsync(d_start, start, R * st * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_ne, n_edges, R * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  h->plain_next = true;
  h->acts_valid = false;
  xgc_handle::PendEdge q; q.kind = 2; q.ptr = d_el; q.ne = d_ne; q.start = start ? d_start : nullptr; q.stride = stride;
  if (can_defer_edges(h)) { h->pend[layer] = q; return 0; }
  return launch_edge_op(h, layer, q);
}

// weekly delta of a persistent layer: one CTA per replicate marks the removed ties in a shared-memory bitmap, compacts the layer
// in place (stable) and appends the added ties (R positions -> slots through pos2slot).  The ties of agents who departed this
// week are dropped in the same pass (tergmLite::delete_vertices): the host's removal indices refer to the list WITHOUT them --
// the list an R session holds after departure_msm -- so a tie's index is its rank among the ties whose endpoints both live.
// (The fast path's aging..hivvl call leaves that pruning to this kernel / to its own acts pass: h->prune_pending.)
#define UPD_IPT 4
__global__ void __launch_bounds__(TPB) k_edges_update(const __grid_constant__ Dev g, int l, const int* delta /*[R][stride]*/, int stride, int has_start) {
  pdl_prologue();
  extern __shared__ unsigned upd_kill[];
  const int r = blockIdx.x + g.r0, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Rep* rp = g.rep + r;
  const int* row = delta + (size_t)r * stride;
  const int ne = rp->ne[l], nrem = max(0, min(row[0], stride - 2)), nadd = max(0, min(row[1], (stride - 2 - nrem) / (has_start ? 3 : 2)));
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);
  __shared__ int wsum[UPD_IPT][TPB / 32], wlive[UPD_IPT][TPB / 32];
  __shared__ int base, base_live;
  for (int k = tid; k < (ne + 31) / 32; k += TPB) upd_kill[k] = 0u;
  if (tid == 0) { base = 0; base_live = 0; }
  __syncthreads();
  bool bad = row[0] != nrem || row[1] != nadd;
  for (int k = tid; k < nrem; k += TPB) { int e = row[2 + k]; if (e >= 0 && e < ne) atomicOr(&upd_kill[e >> 5], 1u << (e & 31)); else bad = true; }
  __syncthreads();
  for (int e0 = 0; e0 < ne; e0 += TPB * UPD_IPT) {          // 1024 ties per round, item q = ties e0 + q * TPB .. (stable order)
    int4 ed[UPD_IPT]; bool live[UPD_IPT]; unsigned ml[UPD_IPT];
#pragma unroll
    for (int q = 0; q < UPD_IPT; ++q) {
      const int e = e0 + q * TPB + tid; live[q] = false; ed[q] = make_int4(0, 0, 0, 0);
      if (e < ne) { ed[q] = E[e]; live[q] = ((A[ed[q].x >> 5] >> (ed[q].x & 31)) & 1u) && ((A[ed[q].y >> 5] >> (ed[q].y & 31)) & 1u); }
      ml[q] = __ballot_sync(0xffffffffu, live[q]);
      if (lane == 0) wlive[q][w] = __popc(ml[q]);
    }
    __syncthreads();
    bool keep[UPD_IPT]; unsigned mk[UPD_IPT]; int runl = base_live;
#pragma unroll
    for (int q = 0; q < UPD_IPT; ++q) {
      int pl = runl, tot = 0;                                // rank of this tie among the living ties = its index for the host
      for (int k = 0; k < TPB / 32; ++k) { int c = wlive[q][k]; tot += c; if (k < w) pl += c; }
      pl += __popc(ml[q] & ((1u << lane) - 1u));
      keep[q] = live[q] && !((upd_kill[pl >> 5] >> (pl & 31)) & 1u);
      mk[q] = __ballot_sync(0xffffffffu, keep[q]);
      if (lane == 0) wsum[q][w] = __popc(mk[q]);
      runl += tot;
    }
    __syncthreads();
    int run = base;
#pragma unroll
    for (int q = 0; q < UPD_IPT; ++q) {
      int off = run, tot = 0;
      for (int k = 0; k < TPB / 32; ++k) { int c = wsum[q][k]; tot += c; if (k < w) off += c; }
      off += __popc(mk[q] & ((1u << lane) - 1u));
      if (keep[q] && off != e0 + q * TPB + tid) E[off] = ed[q];
      run += tot;
    }
    __syncthreads();
    if (tid == 0) { base = run; base_live = runl; }
    __syncthreads();
  }
  const int nk = base, nlive = base_live, na = rp->n_alive, at = g.at_ptr[r];
  for (int k = tid; k < nrem; k += TPB) if (row[2 + k] >= nlive) bad = true;      // an index beyond the pruned list
  const int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  const int* tail = row + 2 + nrem; const int* head = tail + nadd; const int* st = head + nadd;
  for (int k = tid; k < nadd; k += TPB) {
    int a = tail[k], b = head[k];
    if (a < 1 || a > na || b < 1 || b > na) { bad = true; a = b = 1; }
    if (nk + k < g.e_cap[l]) E[nk + k] = make_int4(P2S[a - 1], P2S[b - 1], has_start ? st[k] : at, 0);
  }
  if (bad) atomicOr(&rp->status, XGC_ST_BAD_EDGE);
  if (tid == 0) {
    int nn = nk + nadd;
    if (nn > g.e_cap[l]) { nn = g.e_cap[l]; atomicOr(&rp->status, XGC_ST_EDGE_OVERFLOW); }
    rp->ne[l] = nn;
  }
}
extern "C" int xgc_update_edgelists_all(xgc_handle* h, int layer, const int32_t* delta, int stride, int has_start) {
  if (layer < 0 || layer > 2) FAIL(h, 2, "layer out of range");
  if (stride < 2) FAIL(h, 2, "stride must be >= 2");
  int rc = cache_drop(h); if (rc) return rc;
  CU(h, cudaSetDevice(h->cfg.device));
  const Dev& g = h->g; size_t need = (size_t)g.R * (size_t)stride;
  if (h->pend[layer].kind) { rc = flush_pending_edges(h); if (rc) return rc; }      // (its staging buffer is about to be overwritten)
  if (need > h->upd_stage_cap[layer]) {
    if (h->d_upd_stage[layer]) cudaFree(h->d_upd_stage[layer]);
    h->d_upd_stage[layer] = nullptr; h->upd_stage_cap[layer] = 0;
    // twice the first request: the weekly rows vary in length, and a reallocation (cudaFree: a device-wide wait) inside a running loop
    // would stall every other handle of the GPU
    CU(h, cudaMalloc((void**)&h->d_upd_stage[layer], 2 * need * sizeof(int)));
    h->upd_stage_cap[layer] = 2 * need;
  }
  CU(h, cudaMemcpyAsync(h->d_upd_stage[layer], delta, need * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  h->plain_next = true;
  h->acts_valid = false;
  xgc_handle::PendEdge q; q.kind = 1; q.ptr = h->d_upd_stage[layer]; q.stride = stride; q.has_start = has_start;
  if (can_defer_edges(h)) { h->pend[layer] = q; return 0; }
  return launch_edge_op(h, layer, q);
}
// the standalone kernels of one edge-list operation (strict mode; fast mode when something other than the resident kernel's acts
// pass is about to read the lists)
static int launch_edge_op(xgc_handle* h, int layer, const xgc_handle::PendEdge& q) {
  const Dev& g = h->g;
  if (!h->rank_fresh) {
    if (g.n_cap <= 32768) { prof_begin(h, "k_rank"); kl(h, k_rank_cta, dim3(g.R), TPB, (size_t)(g.n_cap / 32) * 8 + 8, g); prof_end(h); h->launches++; }
    else LAUNCH(h, k_rank, dim3((g.n_cap + RANK_TILE - 1) / RANK_TILE, g.R), TPB, g);
    h->rank_fresh = true;
  }
  if (q.kind == 1) {
    prof_begin(h, "k_edges_update");
    kl(h, k_edges_update, dim3(g.R), TPB, (size_t)((g.e_cap[layer] + 31) / 32 + 1) * 4, g, layer, q.ptr, q.stride, q.has_start);
    prof_end(h); h->launches++;
  } else LAUNCH(h, k_edges_upload, dim3((unsigned)((q.stride + TPB - 1) / TPB), g.R), TPB, g, layer, q.ptr, q.ne, q.start, q.stride);
  CU(h, cudaGetLastError());
  return 0;
}
static int flush_pending_edges(xgc_handle* h) {
  for (int l = 0; l < 3; ++l) {
    if (!h->pend[l].kind) continue;
    xgc_handle::PendEdge q = h->pend[l]; h->pend[l].kind = 0;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = launch_edge_op(h, l, q); if (rc) return rc;
  }
  return 0;
}

// One week of the host-network loop in ONE call (what the wrapper around the host's simnet_msm makes once it has the week's
// network changes): the toggles of the main / casual layers, the new one-off list, the rest of week `at` (mask_tail), the start of
// week at + 1 (mask_head; 0 = none), then week at's tallies and the living counts the host needs for the next network step -- one
// crossing of the FFI boundary, one device -> host wait.  Any of the three edge-list arguments may be NULL (layer unchanged).
extern "C" int xgc_week_hostnet(xgc_handle* h, int at, uint32_t mask_tail, uint32_t mask_head,
                                const int32_t* delta_main, int stride_main, const int32_t* delta_casl, int stride_casl, int has_start,
                                const int32_t* el_inst, const int32_t* ne_inst, int stride_inst,
                                uint32_t* counters_out /*[R][XGC_NCOUNTER], week at*/, int32_t* n_alive_out /*[R]*/) {
  int rc = 0;
  if (delta_main) { rc = xgc_update_edgelists_all(h, 0, delta_main, stride_main, has_start); if (rc) return rc; }
  if (delta_casl) { rc = xgc_update_edgelists_all(h, 1, delta_casl, stride_casl, has_start); if (rc) return rc; }
  if (el_inst) { rc = xgc_set_edgelists_all(h, 2, el_inst, ne_inst, nullptr, stride_inst); if (rc) return rc; }
  rc = mask_head ? xgc_step_span(h, at, mask_tail, mask_head) : xgc_step(h, at, mask_tail, nullptr);
  if (rc) return rc;
  if (counters_out)
    CU(h, cudaMemcpy2DAsync(counters_out, sizeof(uint32_t) * XGC_NCOUNTER, h->g.counters + (size_t)at * XGC_NCOUNTER, sizeof(uint32_t) * XGC_NCOUNTER * (size_t)h->g.t_cap,
                            sizeof(uint32_t) * XGC_NCOUNTER, (size_t)h->g.R, cudaMemcpyDeviceToHost, h->stream));
  if (n_alive_out)
    CU(h, cudaMemcpy2DAsync(n_alive_out, sizeof(int), &h->g.rep[0].n_alive, sizeof(Rep), sizeof(int), (size_t)h->g.R, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// debug entry points for the arithmetic-contract parity tests
// --------------------------------------------------------------------------------------------------------------------
extern "C" int xgc_debug_exp(const double* x, double* y, int n, int device) {
  if (cudaSetDevice(device) != cudaSuccess) return 3;
  double *dx, *dy;
  if (cudaMalloc((void**)&dx, sizeof(double) * n) != cudaSuccess || cudaMalloc((void**)&dy, sizeof(double) * n) != cudaSuccess) return 100;
  cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  k_debug_exp<<<(n + 255) / 256, 256>>>(dx, dy, n);
  cudaError_t e = cudaMemcpy(y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}
extern "C" int xgc_debug_philox(uint32_t seed, uint32_t rep, const uint32_t* ctr4, uint32_t* out4, int n, int device) {
  if (cudaSetDevice(device) != cudaSuccess) return 3;
  uint4 *dc, *dout;
  if (cudaMalloc((void**)&dc, sizeof(uint4) * n) != cudaSuccess || cudaMalloc((void**)&dout, sizeof(uint4) * n) != cudaSuccess) return 100;
  cudaMemcpy(dc, ctr4, sizeof(uint4) * n, cudaMemcpyHostToDevice);
  k_debug_philox<<<(n + 255) / 256, 256>>>(seed, rep, dc, dout, n);
  cudaError_t e = cudaMemcpy(out4, dout, sizeof(uint4) * n, cudaMemcpyDeviceToHost);
  cudaFree(dc); cudaFree(dout);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}
