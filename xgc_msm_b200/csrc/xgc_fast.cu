// k_week_fast — the replicate-resident fast path of the weekly epidemic step (sm_100a), native Philox draws only.
//
// Contract ("arithmetic contract by draw source"): with INJECTED draws the step must be bit-exact against the
// reference modules, and that path (xgc_kernels.cuh, --fmad=false, fp64, one indexed draw per decision) is untouched.
// With NATIVE draws the requirement is statistical equivalence of the targets over >= 1000 replicates, which leaves the
// sampling scheme and the arithmetic free.  This file uses that freedom for a different execution model:
//
//   * ONE CTA PER REPLICATE, persistent over the weeks of a launch.  The per-agent state every module reads -- the demo /
//     gc / hiv words of the core plane (12 B per agent) and the PrEP-indication clock -- is staged ONCE in the SM's shared
//     memory (10k agents: 188 KB of the 227 KB) and stays there for up to compact_every weeks; the endpoint gathers of the
//     edge modules, the exposure / new-infection scatter, the work queues, the weekly tallies and the degree counts of
//     the network step are all shared-memory traffic.  HBM sees the edge records (streamed, L2-resident between weeks),
//     the clock planes of the minority of agents whose clocks change, and the write-back at the end of the launch.
//   * modules are PHASES of that kernel.  Inside a phase every WARP owns its agents (32-agent groups, interleaved over the
//     warps: VSLOT) or ties and runs them to completion on its own: it classifies 32 items at a time, collects the minority
//     that needs work in a 64-entry shared-memory mini-queue and processes the queue 32 at a time with all lanes busy.
//     CTA-wide barriers separate only the phases whose data dependences cross agents (PRE | formation | acts + tie sweep |
//     prep/stirecov/stitx | transmission | apply).  prevalence_msm's state tallies are kept incrementally (tally_move).
//   * rare per-agent events (natural mortality, HIV interval tests, background clinic visits) are drawn as geometric
//     waiting gaps over the slot index (exact for an iid Bernoulli sequence; heterogeneous probabilities by thinning),
//     so agents nothing happens to cost an integer scan, not a Philox block;
//   * ties are deleted LAZILY: a tie whose endpoint departed, or that dissolves this week, gets a tombstone bit in the
//     shared-memory act word when the acts pass meets it (no separate network pass over the ties); new ties are appended
//     at the tail with endpoints drawn by rejection over the slots (no rank -> slot search); a stable in-place squeeze
//     runs every 16 weeks, under capacity pressure, and before the write-back, so HBM always holds compact lists.  Every
//     per-tie draw is keyed by (layer, endpoint slots), never by the tie's index: results do not depend on when the
//     squeeze happens (weekly step-group calls == one resident run, bit for bit);
//   * probabilities are fp32 with FMA, exp via ex2.approx, Bernoulli trials are integer compares on raw Philox words.
//
// Module semantics are those of xgc_kernels.cuh / oracle (same order, same state machine); tests/test_gpu_fast_mode.py
// gates this path statistically against the oracle and structurally (recount of the tallies, accounting identities).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "xgc_fast.cuh"

#define FT 1024                    /* threads per CTA: 32 warps, one CTA per SM                                   */
#define NW 32
#define MQ_CAP 64                  /* entries of a warp's mini-queue                                              */
#define HI_MORT 0.02f              /* weekly death probabilities above this are drawn per agent, not gap-sampled   */
#define FULL 0xffffffffu

// transient bits of the staged gc word (never meaningful in HBM between weeks)
#define FG_TESTED (1u << 23)       /* HIV-tested this week                                                        */
#define FG_AVD    (1u << 31)       /* background clinic-visit draw of this week                                   */

enum { FS_PATH = 1, FS_ARRN, FS_ARRA, FS_MORT, FS_MORTHI, FS_TEST, FS_AG1, FS_DISS, FS_NETN = FS_DISS + 2, FS_FORM,
       FS_ACTS = FS_FORM + 3, FS_COND, FS_ANAL, FS_ORAL, FS_RIM, FS_KISS, FS_AVD, FS_PREP, FS_AG2, FS_INF };
#define EA_DEAD 0x8000u            /* act word of a tie: tombstone (endpoint departed / dissolved), squeezed out later */

// scalar slots
enum { SC_NSLOTS = 0, SC_NPRE, SC_NALIVE, SC_NEXTUID, SC_TREF, SC_TAGE, SC_NE0, SC_NE1, SC_NE2, SC_STATUS, SC_PATH /*7*/, SC_PMAX = SC_PATH + 8,
       SC_DEAD0, SC_DEAD1, SC_DEAD2, SC_NFORM0, SC_NFORM1, SC_NFORM2, SC_SQUEEZE, SC_N = 32 };

struct Fixed {                     // fixed part of the CTA's shared memory (compile-time offsets)
  unsigned row[XGC_NCOUNTER];
  float P[XGC_NPARAM];
  float G[4 * XGC_GLM_NCOEF];
  float asmr[4 * XGC_ASMR_AGES];
  float pk[4 * (XGC_MAX_ACTS + 1)];
  float degw[32];                   // XGC_T_NET_DEGW [3][3][3] + XGC_T_NET_RISKW [5] as fp32
  float netp[8];                    // XGC_T_NET_DEG [3], XGC_T_NET_DUR [3]
  float arrp[16];                   // XGC_T_RACE_PROP [4], XGC_T_ROLE_PROB [4][3]
  unsigned hib[16];
  int sc[SC_N];
  int wsum[NW], wsum2[NW], wsum3[NW];
  unsigned short mq[NW][2][MQ_CAP];
};
static __host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
struct VarOff { size_t ea, deg, V, total; };
static __host__ __device__ VarOff fast_layout(int n_cap, size_t e_stride) {
  VarOff o; size_t p = al16(sizeof(Fixed));
  o.ea = p; p += al16(e_stride * 2);
  o.deg = p; p += al16((size_t)n_cap);                  // one byte per slot: main degree (low nibble), casual degree (high nibble)
  o.V = p; p += al16((size_t)n_cap * 14);
  o.total = p;
  return o;
}
size_t xgc_fast_smem_bytes(int n_cap, size_t e_stride) {
  VarOff o = fast_layout(n_cap, e_stride);
  return (o.total <= 232448 && n_cap <= 32768 && e_stride <= 16383) ? o.total : 0;      // 14-bit tie indices + 2 type bits in the queue items, 15-bit slots
}

// ---------------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 frand(uint2 key, unsigned at, unsigned stream, unsigned a, unsigned b) {
  // 7 rounds: the smallest Philox4x32 variant that passes BigCrush (Salmon et al. SC'11, table 2); the strict path keeps 10
  return philox4x32<7>(make_uint4(at, stream, a, b), key);
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }     // (0,1), 24 bits
__device__ __forceinline__ float u01_16(unsigned x16) { return ((float)x16 + 0.5f) * (1.0f / 65536.0f); }
__device__ __forceinline__ unsigned thr32(float p) { return __float2uint_rz(fminf(fmaxf(p, 0.f), 1.f) * 4294967296.f); }  // saturates at 2^32-1
__device__ __forceinline__ bool bern(unsigned x, float p) { return x < thr32(p); }
__constant__ float F_INVK[33] = { 0.f, 1.f / 1, 1.f / 2, 1.f / 3, 1.f / 4, 1.f / 5, 1.f / 6, 1.f / 7, 1.f / 8, 1.f / 9, 1.f / 10, 1.f / 11, 1.f / 12,
                                  1.f / 13, 1.f / 14, 1.f / 15, 1.f / 16, 1.f / 17, 1.f / 18, 1.f / 19, 1.f / 20, 1.f / 21, 1.f / 22, 1.f / 23,
                                  1.f / 24, 1.f / 25, 1.f / 26, 1.f / 27, 1.f / 28, 1.f / 29, 1.f / 30, 1.f / 31, 1.f / 32 };
// Poisson by inversion of one uniform, truncated at kmax
__device__ __forceinline__ int pois_f(float u, float mu, int kmax) {
  if (!(mu > 0.f)) return 0;
  float pr = __expf(-mu), cdf = pr; int k = 0;
  while (u > cdf && k < kmax) { k++; pr = pr * mu * (k <= 32 ? F_INVK[k] : 1.f / (float)k); cdf += pr; }
  return k;
}
__device__ __forceinline__ int pois_tab_f(float u, const float* tab) { int k = 0; while (k < XGC_MAX_ACTS && u > tab[k]) k++; return k; }
__device__ __forceinline__ int categorical_f(float u, const float* p, int n) { int r = 0; float c = p[0]; while (u >= c && r < n - 1) { r++; c += p[r]; } return r; }
__device__ __forceinline__ float any_of_n_f(float t, int n) { return 1.f - __powf(1.f - t, (float)n); }
__device__ __forceinline__ unsigned agegrp_f(float age) { return (unsigned)((age >= 25.f) + (age >= 35.f) + (age >= 45.f) + (age >= 55.f)); }
__device__ __forceinline__ float glm_f(const float* b, float dur, bool casual, float ca, bool hcp, bool prep, int r1, int r2) {
  float eta = b[XGC_G_B0] + b[XGC_G_RC + r1 * 4 + r2];
  eta = fmaf(b[XGC_G_DUR], dur, eta); eta = fmaf(b[XGC_G_DUR2], dur * dur, eta);
  if (casual) eta += fmaf(b[XGC_G_DUR_PT2], dur, b[XGC_G_PT2]);
  eta = fmaf(b[XGC_G_CA], ca, eta); eta = fmaf(b[XGC_G_CA2], ca * ca, eta);
  if (hcp) eta += b[XGC_G_HCP];
  if (prep) eta += b[XGC_G_PREP];
  return eta;
}
// prevalence_msm's state tallies (series = counter group, by race x age cell) are kept INCREMENTALLY: counted once when the
// replicate is staged, then every change of an agent's tallied state moves its counts.  tally_bits: bit b = agent counts in group b.
#define STATE_GROUPS ((1u << XGC_K_NUM) | (1u << XGC_K_INUM) | (1u << XGC_K_INUM_DX) | (1u << XGC_K_VSUPP) | (1u << XGC_K_PREPELIG) | (1u << XGC_K_PREPCURR) | \
                      (1u << XGC_K_INUM_GC) | (1u << XGC_K_INUM_RGC) | (1u << XGC_K_INUM_UGC) | (1u << XGC_K_INUM_PGC))
__device__ __forceinline__ unsigned tally_bits(unsigned d, unsigned gc, unsigned h) {
  if (!(d & DM_ACTIVE)) return 0u;
  return (1u << XGC_K_NUM) | ((h & HV_STATUS) ? 1u << XGC_K_INUM : 0u) | ((h & HV_DIAG) ? 1u << XGC_K_INUM_DX : 0u) | (((h & HV_DIAG) && (h & HV_VSUPP)) ? 1u << XGC_K_VSUPP : 0u) |
         ((h & HV_PREPELIG) ? 1u << XGC_K_PREPELIG : 0u) | ((h & HV_PREP) ? 1u << XGC_K_PREPCURR : 0u) | ((gc & 7u) ? 1u << XGC_K_INUM_GC : 0u) |
         ((gc & 1u) ? 1u << XGC_K_INUM_RGC : 0u) | ((gc & 2u) ? 1u << XGC_K_INUM_UGC : 0u) | ((gc & 4u) ? 1u << XGC_K_INUM_PGC : 0u);
}
// an agent's (demo, gc, hiv) words went from (d0, g0, h0) to (d1, g1, h1): move its counts (shared-memory atomics, a handful per event)
__device__ __noinline__ void tally_move(unsigned* row, unsigned d0, unsigned g0, unsigned h0, unsigned d1, unsigned g1, unsigned h1) {
  const unsigned b0 = tally_bits(d0, g0, h0), b1 = tally_bits(d1, g1, h1), c0 = DM_CELL(d0), c1 = DM_CELL(d1);
  unsigned sub = c0 == c1 ? b0 & ~b1 : b0, add = c0 == c1 ? b1 & ~b0 : b1;
  while (sub) { const int b = __ffs(sub) - 1; sub &= sub - 1; atomicSub(&row[b * XGC_NCELL + c0], 1u); }
  while (add) { const int b = __ffs(add) - 1; add &= add - 1; atomicAdd(&row[b * XGC_NCELL + c1], 1u); }
}
__device__ __forceinline__ unsigned pairkey(const int4& ed) { return (unsigned)ed.x | ((unsigned)ed.y << 16); }      // slots < 32768
__device__ __forceinline__ int birth8(unsigned h) { return (int)((h >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK); }
// exclusive prefix of the 32 per-warp counts in cnt[] for this warp, and their total (every warp computes it redundantly)
__device__ __forceinline__ int warp_base(const int* cnt, int warp, int lane, int* total) {
  int v = cnt[lane], inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
  *total = __shfl_sync(FULL, inc, 31);
  return __shfl_sync(FULL, inc - v, warp);
}
// A warp's mini-queue: the minority of a stream of items that needs work, processed 32 at a time with all lanes busy.
// PUSH adds this lane's item if `flag`; DRAIN(last, item, body) runs `body` for batches of 32 (and, when `last`, for the
// remainder).  `body` must not contain full-warp collectives (the remainder runs with the upper lanes off).
#define WQ_PUSH(q, cnt, flag, itemv) do { unsigned m_ = __ballot_sync(FULL, (flag)); \
  if (m_) { if (flag) (q)[(cnt) + __popc(m_ & ((1u << lane) - 1u))] = (unsigned short)(itemv); (cnt) += __popc(m_); } } while (0)
#define WQ_DRAIN(q, cnt, last, item, ...) do { __syncwarp(); \
  while ((cnt) >= 32 || ((last) && (cnt) > 0)) { const int nb_ = min((cnt), 32); const unsigned item = (q)[lane]; const unsigned short nx_ = (q)[32 + lane]; __syncwarp(); \
    if (lane < nb_) { __VA_ARGS__ } __syncwarp(); (cnt) -= nb_; if (lane < (cnt)) (q)[lane] = nx_; __syncwarp(); } } while (0)

// Gap sampling of an iid Bernoulli(p) sequence over positions [a, b) by ONE warp: 32 geometric gaps at a time, prefix-
// summed.  Keyed by (stream, id): the caller passes an id that identifies the range independently of who runs it.
// f(valid, pos, x) is called by ALL lanes for every batch; x is the Philox block the gap came from (x.y, x.z, x.w unused).
template <class Fn>
__device__ __forceinline__ void gap_warp(uint2 key, unsigned at, unsigned stream, unsigned id, int a, int b, float p, int lane, Fn f) {
  if (!(p > 0.f) || b <= a) return;
  const float ilq = p >= 1.f ? 0.f : 1.f / log1pf(-p);       // 1 / ln(1 - p) < 0; p >= 1: every position
  int pos = a - 1;
  for (unsigned chunk = 0; pos < b; ++chunk) {
    uint4 x = frand(key, at, stream, id | (chunk << 16), (unsigned)lane);
    int s = (int)fminf(__logf(u01(x.x)) * ilq, 1.0e6f) + 1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, s, o); if (lane >= o) s = min(s + v, 1 << 28); }
    int me = pos + s;
    f(me < b, me, x);
    pos = __shfl_sync(FULL, me, 31);
  }
}

// Agent ownership: warp w owns the 32-agent groups w, w + 32, w + 64, ... (slot order is arrival order, i.e. roughly age order:
// contiguous ranges would give the warps of the early slots the older cohorts with more HIV-positive agents -- systematic
// imbalance in front of every phase barrier).  VSLOT(p): the slot of the warp's p-th agent.
#define VSLOT(p) ((((p) >> 5) * NW + warp) * 32 + ((p) & 31))
// phase clocks (SM cycles summed over CTAs; thread 0 of each CTA): tools/time_modes.py
enum { PH_STAGE = 0, PH_ARRIVE, PH_PRE, PH_NET, PH_ACTS, PH_B, PH_TRANS, PH_APPLY_TALLY, PH_ROW, PH_WRITEBACK, PH_SQUEEZE_WB, PH_STAGE_TABLES, PH_HOSTNET, PH_N = 16 };
__device__ unsigned long long g_fast_prof[PH_N];
#define TICK(k) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_fast_prof[k], (unsigned long long)(t_ - t_last)); t_last = t_; } } while (0)

// the test of one HIV-test candidate: time since the last negative test / arrival decides (item bit 15: rule-based)
#define TEST_BODY { const int i = (int)(item & 0x7FFFu); const bool rule = item & 0x8000u; const unsigned d = vd[i], h = vh[i]; \
  if (d & DM_ACTIVE) { short lnt = g.tck[rbase + i].x; int last = lnt == XGC_NA16 ? g.arr_t[rbase + i] : (int)lnt; float tsl = (float)(at - last); \
    bool tested = rule ? ((HV_STAGE(h) == 4 && tsl >= P[XGC_P_aids_test_int]) || ((h & HV_PREP) && tsl >= P[XGC_P_prep_tst_int])) : tsl >= 2.f; \
    if (tested) { const unsigned g0_ = atomicOr(&vg[i], FG_TESTED); \
      if (!(g0_ & FG_TESTED)) { if (wt) core_w[4 * (rbase + i) + 2] = g0_ | FG_TESTED; atomicAdd(&F.row[XGC_K_HIVTESTS], 1u); if (!(h & HV_STATUS)) g.tck[rbase + i].x = (short)at; } } } }

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT, 1) k_week_fast(const __grid_constant__ Dev g, const __grid_constant__ FastArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fixed& F = *reinterpret_cast<Fixed*>(smem_raw);
  const VarOff vo = fast_layout(g.n_cap, g.e_stride);
  unsigned short* const ea = (unsigned short*)(smem_raw + vo.ea);
  unsigned char* const deg8 = smem_raw + vo.deg;              // current degrees, kept while the surrogate network runs on the device
  unsigned* const deg32 = (unsigned*)deg8;
  unsigned* const vd = (unsigned*)(smem_raw + vo.V);
  unsigned* const vg = vd + g.n_cap;
  unsigned* const vh = vd + 2 * g.n_cap;
  short* const ind = (short*)(vd + 3 * g.n_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* const P = F.P;
  unsigned short* const q0 = F.mq[warp][0];
  unsigned short* const q1 = F.mq[warp][1];
  long long t_last = clock64();
  // a call that ends before the POST group (the aging..hivvl call of the weekly host-network pattern) changes few agents: it
  // writes every change of a core word through to HBM and skips the write-back of the plane
  const bool wt = !((A.phases | A.ph_last) & XF_POST);
  unsigned* const core_w = (unsigned*)g.core;
  // Schedule: the launch's replicate-weeks, replicate-major, are cut into gridDim.x equal stretches (148 CTAs x 1000 replicates
  // would otherwise run 7 waves for 6.76 waves of work).  A CTA walks its stretch from the LAST replicate to the first: the
  // replicate it shares with the next CTA comes first (its early weeks, no dependence), the one it shares with the previous
  // CTA last (its late weeks; the previous CTA ran the early ones first and published at_ptr[r]).  A CTA touches a replicate
  // once per launch, so nothing it reads can be stale in its L1 except the shared Rep / at_ptr lines, read with ld.cg.
  const long long W = A.at1 - A.at0 + 1, T = (long long)g.R * W;
  const long long i0 = T * blockIdx.x / gridDim.x, i1 = T * (blockIdx.x + 1) / gridDim.x;
  for (int rr = i1 > i0 ? (int)((i1 - 1) / W) : -1; rr >= 0 && rr >= (int)(i0 / W); --rr) {
    const int r = rr + g.r0;
    const int w_first = A.at0 + (int)max(i0 - (long long)rr * W, 0LL), w_last = A.at0 + (int)min(i1 - (long long)rr * W, W) - 1;
    Rep* rp = g.rep + r;
    const size_t rbase = (size_t)r * g.n_cap;
    const uint2 key = make_uint2(g.seed, (unsigned)g.rep_id[r]);
    const long long t_rep = clock64();
    const int* C = g.control + (size_t)r * XGC_NCONTROL;
    const int proto = C[XGC_C_stiScreeningProtocol], kissx = C[XGC_C_cdcExposureSite_Kissing], surrogate = C[XGC_C_network_mode] == 1;
    const bool pois_mode = C[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_POIS;
    const bool route_kiss = C[XGC_C_transRoute_Kissing], route_rim = C[XGC_C_transRoute_Rimming];
    if (w_first > A.at0) {                                      // the earlier weeks of this replicate run on the previous CTA
      if (tid == 0) while (__ldcg(&g.at_ptr[r]) != w_first - 1) __nanosleep(500);
      __syncthreads();
      __threadfence();
    }
    __syncthreads();                                            // previous replicate's write-back done with the staging buffers
    // ------------------------------------------------------------------------------------------ stage the replicate
    if (tid == 0) {
      F.sc[SC_NSLOTS] = __ldcg(&rp->n_slots); F.sc[SC_NPRE] = __ldcg(&rp->n_slots_pre); F.sc[SC_NALIVE] = __ldcg(&rp->n_alive); F.sc[SC_NEXTUID] = __ldcg(&rp->next_uid);
      F.sc[SC_TREF] = __ldcg(&rp->t_ref); F.sc[SC_TAGE] = __ldcg(&rp->t_age); F.sc[SC_NE0] = __ldcg(&rp->ne[0]); F.sc[SC_NE1] = __ldcg(&rp->ne[1]); F.sc[SC_NE2] = __ldcg(&rp->ne[2]);
      F.sc[SC_STATUS] = 0; F.sc[SC_DEAD0] = F.sc[SC_DEAD1] = F.sc[SC_DEAD2] = 0;
      for (int k = 0; k < XGC_NPATH; ++k) F.sc[SC_PATH + k] = __ldcg(&rp->path_rank[k]);
    }
    for (int k = tid; k < XGC_NPARAM; k += FT) F.P[k] = (float)g.params[(size_t)r * XGC_NPARAM + k];
    for (int k = tid; k < 4 * XGC_GLM_NCOEF; k += FT) F.G[k] = (float)g.tables[XGC_T_GLM_AI + k];
    for (int k = tid; k < 4 * XGC_ASMR_AGES; k += FT) F.asmr[k] = (float)g.tables[XGC_T_ASMR + k];
    for (int k = tid; k < 4 * (XGC_MAX_ACTS + 1); k += FT) F.pk[k] = (float)g.pois[(size_t)r * 4 * (XGC_MAX_ACTS + 1) + k];
    if (tid < 32) F.degw[tid] = (float)g.tables[XGC_T_NET_DEGW + tid];
    if (tid < 6) F.netp[tid] = (float)g.tables[XGC_T_NET_DEG + tid];
    if (tid < 16) F.arrp[tid] = (float)g.tables[XGC_T_RACE_PROP + tid];
    if (surrogate) for (int k = tid; k < g.n_cap / 4; k += FT) deg32[k] = 0u;
    if (tid < 16) F.hib[tid] = 0u;
    for (int k = tid; k < (int)g.e_stride; k += FT) ea[k] = 0;   // HBM holds compact lists: no tombstones
    __syncthreads();
    if (tid < 4 * XGC_ASMR_AGES && F.asmr[tid] >= HI_MORT) atomicOr(&F.hib[(tid / XGC_ASMR_AGES) * 4 + ((tid % XGC_ASMR_AGES) >> 5)], 1u << ((tid % XGC_ASMR_AGES) & 31));
    TICK(PH_STAGE_TABLES);
    {
      const int n0 = F.sc[SC_NSLOTS];
      // four independent 16-byte loads per thread in flight (one per thread leaves the SM latency-bound at ~1/4 of its share
      // of the HBM bandwidth); the PrEP-indication clock is only needed by the POST group
      for (int i0 = tid; i0 < n0; i0 += 4 * FT) {
        uint4 c[4]; short t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = min(i0 + k * FT, n0 - 1); c[k] = g.core[rbase + i]; t[k] = wt ? (short)-1 : g.tck[rbase + i].y; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = i0 + k * FT; if (i < n0) { vd[i] = c[k].y; vg[i] = c[k].z; vh[i] = c[k].w; if (!wt) ind[i] = t[k]; } }
      }
    }
    // degree-conditioned formation (abi.h XGC_T_NET_DEGW): the degrees of the staged lists, then kept current by the formation
    // step (+) and by the tombstones of the tie sweep (-).  As in the oracle, a tie that dissolves THIS week still counts while
    // this week's ties form (formation and dissolution both condition on the network the week starts with); the one difference:
    // the ties of an agent who departed this week are dropped by the acts pass, i.e. after the formation step (~5 ties a week)
    auto deg_add = [&](int s, int l) { if (((deg8[s] >> (4 * l)) & 15u) < 15u) atomicAdd(&deg32[s >> 2], (l ? 16u : 1u) << (8 * (s & 3))); };
    auto deg_sub = [&](int s, int l) { if ((deg8[s] >> (4 * l)) & 15u) atomicSub(&deg32[s >> 2], (l ? 16u : 1u) << (8 * (s & 3))); };
    auto form_w = [&](int l, unsigned dg, unsigned demo) -> float {
      const int cm = min((int)(dg & 15u), 2), cc = min((int)(dg >> 4), 2);
      float w = F.degw[l * 9 + cm * 3 + cc];
      if (l == 2) w *= F.degw[27 + DM_RISK(demo)];
      return w;
    };
    auto count_degrees = [&]() {
      const int ne0 = F.sc[SC_NE0], ne01 = ne0 + F.sc[SC_NE1];
      for (int f = tid; f < ne01; f += FT) {
        const int l = f >= ne0, sidx = g.e_off[l] + f - (l ? ne0 : 0);
        if (ea[sidx] & EA_DEAD) continue;
        const int4 ed = g.edge[(size_t)r * g.e_stride + sidx];
        deg_add(ed.x, l); deg_add(ed.y, l);
      }
    };
    if (surrogate) count_degrees();
    if (warp == 0) {                                            // largest gap-sampled weekly death probability
      float v = 0.f;
      for (int k = lane; k < 4 * XGC_ASMR_AGES; k += 32) { float a = F.asmr[k]; if (a < HI_MORT) v = fmaxf(v, a); }
#pragma unroll
      for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
      if (lane == 0) F.sc[SC_PMAX] = __float_as_int(v);
    }
    __syncthreads();
    // multi-week launches keep prevalence_msm's state tallies incrementally (counted here once, then moved by every change of an
    // agent's tallied state); a single-week call counts them once at the end of its POST group instead
    const bool incr = A.at1 > A.at0;
    if (incr && A.carry_in && w_first == A.at0) {                // a span continuing the previous span's week: its running tallies
      for (int k = tid; k < XGC_K_NGROUP * XGC_NCELL; k += FT) F.row[k] = ((STATE_GROUPS >> (k / XGC_NCELL)) & 1u) ? A.carry[(size_t)r * XGC_NCOUNTER + k] : 0u;
      __syncthreads();
    } else if (incr) {
      for (int k = tid; k < XGC_K_NGROUP * XGC_NCELL; k += FT) F.row[k] = 0u;
      __syncthreads();
      const int n0 = F.sc[SC_NSLOTS];
      const int AW = ((n0 + FT - 1) / FT) * 32, a0 = warp * AW, a1 = min(a0 + AW, n0);
      for (int k = 0; k * 32 < AW; ++k) {
        const int i = a0 + k * 32 + lane;
        unsigned d = i < a1 ? vd[i] : 0u;
        unsigned bits = i < a1 ? tally_bits(d, vg[i], vh[i]) & ~(1u << XGC_K_NUM) : 0u;
        const bool live = d & DM_ACTIVE; const unsigned cell = live ? DM_CELL(d) : 31u;
        const unsigned grp = __match_any_sync(FULL, cell);
        if (live && lane == (unsigned)(__ffs(grp) - 1)) atomicAdd(&F.row[XGC_K_NUM * XGC_NCELL + cell], (unsigned)__popc(grp));
        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; atomicAdd(&F.row[b * XGC_NCELL + cell], 1u); }
      }
      __syncthreads();
    }
    const float pmax_mort = __int_as_float(F.sc[SC_PMAX]);
    const float pmax_test = fmaxf(fmaxf(P[XGC_P_hiv_test_rate], P[XGC_P_hiv_test_rate + 1]), fmaxf(P[XGC_P_hiv_test_rate + 2], P[XGC_P_hiv_test_rate + 3]));
    // tie dissolution, Bernoulli(1 / duration) per tie-week: a 16-bit pre-test on spare bits of the tie's acts block, refined
    // by one more draw in the rare case it passes (exactly p overall)
    const float pd0 = (float)g.inv_dur[0], pd1 = (float)g.inv_dur[1];
    const unsigned dthr0 = min(65536u, (unsigned)ceilf(pd0 * 65536.f)), dthr1 = min(65536u, (unsigned)ceilf(pd1 * 65536.f));
    const unsigned dref0 = dthr0 ? thr32(pd0 * 65536.f / (float)dthr0) : 0u, dref1 = dthr1 ? thr32(pd1 * 65536.f / (float)dthr1) : 0u;
    auto dissolves = [&](const uint4& x, const int4& ed, int l, int at) -> bool {
      const unsigned c16 = (x.x & 0xFFu) | ((x.y & 0xFFu) << 8);
      if (c16 >= (l ? dthr1 : dthr0)) return false;
      return frand(key, at, FS_DISS + l, pairkey(ed), 0).x <= (l ? dref1 : dref0);
    };
    // stable in-place squeeze of the tombstoned ties out of the three lists (CTA-wide; 4096 ties at a time, 128 consecutive ties
    // per warp held in registers)
    auto squeeze = [&]() {
      for (int l = 0; l < 3; ++l) {
        const int ne = F.sc[SC_NE0 + l];
        if (F.sc[SC_DEAD0 + l] == 0) continue;
        int4* const E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
        unsigned short* const eal = ea + g.e_off[l];
        int run = 0;
        for (int s0 = 0; s0 < ne; s0 += 4 * FT) {
          int4 ed[4]; unsigned mk[4]; int cnt = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = s0 + (warp * 4 + k) * 32 + lane; bool keep = false; ed[k] = make_int4(0, 0, 0, 0);
            if (e < ne && !(eal[e] & EA_DEAD)) { keep = true; ed[k] = E[e]; }
            mk[k] = __ballot_sync(FULL, keep); cnt += __popc(mk[k]);
          }
          if (lane == 0) F.wsum[warp] = cnt;
          __syncthreads();
          int tot, dst0 = run + warp_base(F.wsum, warp, lane, &tot);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = s0 + (warp * 4 + k) * 32 + lane, dst = dst0 + __popc(mk[k] & ((1u << lane) - 1u));
            if (((mk[k] >> lane) & 1u) && dst != e) E[dst] = ed[k];
            dst0 += __popc(mk[k]);
          }
          run += tot;
          __syncthreads();
        }
        for (int e = tid; e < ne; e += FT) eal[e] = 0;           // (act counts are per week: the acts pass rewrites them)
        if (tid == 0) { F.sc[SC_NE0 + l] = run; F.sc[SC_DEAD0 + l] = 0; }
        __syncthreads();
      }
    };
    TICK(PH_STAGE);

    for (int at = w_first; at <= w_last; ++at) {
      const unsigned mask = A.mask;
      // a span call (xgc_step_span) runs the rest of week at0 and the start of week at1: phase groups by week
      const unsigned ph = at == A.at0 ? A.phases : at == A.at1 ? A.ph_last : (unsigned)XF_ALL;
      unsigned* const grow = g.counters + ((size_t)r * g.t_cap + (size_t)at) * XGC_NCOUNTER;
      const int at8 = (at + XF_BIRTH_OFF) * 8;
      // =================================================================================================== PRE
      if (ph & XF_PRE) {
        for (int k = tid; k < XGC_NCOUNTER; k += FT)            // the week's event counters start at zero; state tallies carry over
          if (!incr || k >= XGC_K_NGROUP * XGC_NCELL || !((STATE_GROUPS >> (k / XGC_NCELL)) & 1u)) F.row[k] = 0u;
        // ---- arrival_msm: one warp
        if (warp == 0) {
          int n_slots = F.sc[SC_NSLOTS], nnew = 0;
          if (mask & XGC_M_ARRIVAL) {
            float mu = P[XGC_P_a_rate] * P[XGC_P_n_init];
            int m = (int)ceilf(mu / 32.f); m = max(1, min(m, 32));
            { uint4 x = frand(key, at, FS_ARRN, lane, 0); if (lane < m) nnew = pois_f(u01(x.x), mu / (float)m, 255); }
#pragma unroll
            for (int o = 16; o; o >>= 1) nnew += __shfl_xor_sync(FULL, nnew, o);
            int uid0 = F.sc[SC_NEXTUID], t_ref = F.sc[SC_TREF], t_age = (mask & XGC_M_AGING) ? at : F.sc[SC_TAGE];
            const float arr_age = P[XGC_P_arrival_age];
            if (n_slots + nnew > g.n_cap) { nnew = g.n_cap - n_slots; if (lane == 0) atomicOr(&F.sc[SC_STATUS], (int)XGC_ST_AGENT_OVERFLOW); }
            for (int j = lane; j < nnew; j += 32) {
              int i = n_slots + j; size_t idx = rbase + i;
              uint4 x = frand(key, at, FS_ARRA, (unsigned)i, 0), y = frand(key, at, FS_ARRA, (unsigned)i, 1);
              int race = categorical_f(u01(x.x), F.arrp, 4);
              int role = categorical_f(u01(x.y), F.arrp + 4 + race * 3, 3);
              float insq = role == 0 ? 1.f : role == 1 ? 0.f : u01(x.z);
              unsigned circ = bern(x.w, P[XGC_P_circ_prob + race]), late = bern(y.x, P[XGC_P_hiv_test_late_prob + race]);
              float tt[3] = { P[XGC_P_tt_part_supp + race], P[XGC_P_tt_full_supp + race], P[XGC_P_tt_dur_supp + race] };
              unsigned traj = 1 + categorical_f(u01(y.y), tt, 3);
              unsigned stopper = bern(y.z, P[XGC_P_act_stopper_prob]);
              unsigned rg = min(4u, __umulhi(y.w, 5u));
              Aux ax; ax.age_ref = (double)arr_age - (double)(t_age - t_ref) * XGC_WEEK; ax.vl = 0.f; ax.ins_quot = insq;
              unsigned demo = (unsigned)race | (agegrp_f(arr_age) << 2) | ((unsigned)role << 5) | (rg << 7) | (circ ? DM_CIRC : 0u) | (late ? DM_LATE : 0u) |
                              (traj << 12) | (stopper ? DM_STOPPER : 0u) | DM_ACTIVE | (xf_insq16(insq) << 16);
              unsigned hiv = HV_VSUPP | (xf_birth_code((double)arr_age, at) << XF_BIRTH_SHIFT);
              g.core[idx] = make_uint4((unsigned)(uid0 + j), demo, 0u, hiv);
              g.aux[idx] = ax;
              short4 na = make_short4(-1, -1, -1, -1);
              g.gck_a[idx] = na; g.gck_b[idx] = na; g.gck_d[idx] = na; g.hck_a[idx] = na; g.hck_b[idx] = make_short4(0, 0, -1, -1); g.tck[idx] = na;
              g.arr_t[idx] = at;
              vd[i] = demo; vg[i] = 0u; vh[i] = hiv; ind[i] = -1;
              if (incr) tally_move(F.row, 0u, 0u, 0u, demo, 0u, hiv);
            }
          }
          if (lane == 0) {
            F.sc[SC_NPRE] = n_slots;
            if (mask & XGC_M_AGING) F.sc[SC_TAGE] = at;
            F.sc[SC_NSLOTS] = n_slots + nnew; F.sc[SC_NEXTUID] += nnew; F.sc[SC_NALIVE] += nnew;
          }
        }
        __syncthreads();
        if (tid == 0) F.row[XGC_K_NNEW] = (unsigned)(F.sc[SC_NSLOTS] - F.sc[SC_NPRE]);
        TICK(PH_ARRIVE);
        // ---- every warp runs its own range of agents through aging, departure, hivtest, hivtx, hivprogress, hivvl
        const int n_slots = F.sc[SC_NSLOTS], n_pre = F.sc[SC_NPRE];
        const int AW = ((n_slots + FT - 1) / FT) * 32;           // agents per warp (32-agent groups, interleaved: VSLOT)
        const bool hivmods = mask & (XGC_M_HIVTEST | XGC_M_HIVTX | XGC_M_HIVPROGRESS | XGC_M_HIVVL);
        int c0 = 0, c1 = 0;                                     // mini-queue fills: q0 = HIV-test candidates, q1 = HIV-positive agents
        // departure_msm, natural mortality: geometric gaps at the largest table rate, thinned to (race, age)
        if (mask & XGC_M_DEPARTURE)
          gap_warp(key, at, FS_MORT, (unsigned)warp, 0, AW, pmax_mort, lane, [&](bool ok, int p, const uint4& x) {
            const int i = VSLOT(p);
            if (!ok || i >= n_pre) return;                      // (this week's arrivals are not exposed to mortality)
            unsigned d = vd[i];
            if (!(d & DM_ACTIVE)) return;
            int ai = min(XGC_ASMR_AGES - 1, max(0, (int)((float)(at8 - birth8(vh[i])) * (float)XF_W8)));
            float pm = F.asmr[DM_RACE(d) * XGC_ASMR_AGES + ai];
            if (pm < HI_MORT && u01(x.y) * pmax_mort < pm) { vd[i] = d & ~DM_ACTIVE; core_w[4 * (rbase + i) + 1] = d & ~DM_ACTIVE;
              if (incr) tally_move(F.row, d, vg[i], vh[i], 0u, 0u, 0u);
              atomicAdd(&F.row[XGC_K_DEP_GEN], 1u); atomicSub(&F.sc[SC_NALIVE], 1); }
          });
        __syncwarp();
        // hivtest_msm, interval testing: gaps at the largest race rate, thinned by race
        if (mask & XGC_M_HIVTEST)
          gap_warp(key, at, FS_TEST, (unsigned)warp, 0, AW, pmax_test, lane, [&](bool ok, int p, const uint4& x) {
            bool cand = false; const int i = VSLOT(p);
            if (ok && i < n_slots) { unsigned d = vd[i], h = vh[i];
              cand = (d & DM_ACTIVE) && !(d & DM_LATE) && !(h & (HV_DIAG | HV_PREP)) && u01(x.y) * pmax_test < P[XGC_P_hiv_test_rate + DM_RACE(d)]; }
            WQ_PUSH(q0, c0, cand, i);
            WQ_DRAIN(q0, c0, false, item, TEST_BODY);
          });
        // scan: aging, high-rate mortality (ageing out), rule-based tests; the HIV-positive agents are remembered as a bitmap
        unsigned hword = 0u;                                    // lane k: HIV-positive bits of word k of this warp's range
        for (int k = 0; k * 32 < AW; ++k) {
          const int i = (k * NW + warp) * 32 + lane;
          bool live = false, qpos = false, qrule = false;
          if (i < n_slots) {
            unsigned d = vd[i], h = vh[i];
            live = d & DM_ACTIVE;
            if (live) {
              float age = (float)(at8 - birth8(h)) * (float)XF_W8;
              bool isnew = i >= n_pre;
              if ((mask & XGC_M_AGING) && !isnew) { unsigned ng = agegrp_f(age); if (ng != DM_AGEGRP(d)) { const unsigned dold = d; d = (d & ~(7u << 2)) | (ng << 2); vd[i] = d; core_w[4 * (rbase + i) + 1] = d;
                if (incr) tally_move(F.row, dold, vg[i], h, d, vg[i], h); } }
              if ((mask & XGC_M_DEPARTURE) && !isnew) {
                int ai = min(XGC_ASMR_AGES - 1, max(0, (int)age));
                if ((F.hib[DM_RACE(d) * 4 + (ai >> 5)] >> (ai & 31)) & 1u) {
                  uint4 x = frand(key, at, FS_MORTHI, (unsigned)i, 0);
                  if (bern(x.x, F.asmr[DM_RACE(d) * XGC_ASMR_AGES + ai])) { vd[i] = d & ~DM_ACTIVE; core_w[4 * (rbase + i) + 1] = d & ~DM_ACTIVE;
                    if (incr) tally_move(F.row, d, vg[i], h, 0u, 0u, 0u);
                    atomicAdd(&F.row[XGC_K_DEP_GEN], 1u); atomicSub(&F.sc[SC_NALIVE], 1); live = false; }
                }
              }
              qpos = live && hivmods && (h & HV_STATUS) && !isnew;
              qrule = live && (mask & XGC_M_HIVTEST) && !(h & HV_DIAG) && ((h & HV_PREP) || HV_STAGE(h) == 4);
            }
          }
          unsigned mp = __ballot_sync(FULL, qpos);
          if (lane == k) hword = mp;
          WQ_PUSH(q0, c0, qrule, i | 0x8000);
          WQ_DRAIN(q0, c0, (k + 1) * 32 >= AW, item, TEST_BODY);
        }
        // HIV-positive agents of this warp's range: AIDS mortality, test result, hivtx_msm, hivprogress_msm, hivvl_msm
        for (int k = 0; k * 32 < AW; ++k) {
          const unsigned mp = __shfl_sync(FULL, hword, k);
          WQ_PUSH(q1, c1, (mp >> lane) & 1u, (k * NW + warp) * 32 + lane);
          WQ_DRAIN(q1, c1, (k + 1) * 32 >= AW, item,
            const int i = (int)item; const size_t idx = rbase + i;
            unsigned d = vd[i], h = vh[i];
            if (d & DM_ACTIVE) {                                // (not) died of natural causes this week
              uint4 x = frand(key, at, FS_AG1, (unsigned)i, 0);
              if ((mask & XGC_M_DEPARTURE) && HV_STAGE(h) == 4 && bern(x.x, P[XGC_P_aids_mr])) {
                vd[i] = d & ~DM_ACTIVE; core_w[4 * idx + 1] = d & ~DM_ACTIVE;
                if (incr) tally_move(F.row, d, vg[i], h, 0u, 0u, 0u);
                atomicAdd(&F.row[XGC_K_DEP_AIDS], 1u); atomicSub(&F.sc[SC_NALIVE], 1);
              } else {
                short4 ha = g.hck_a[idx]; short4 hb = g.hck_b[idx]; float vl0 = g.aux[idx].vl;
                const unsigned h0 = h; bool ha_dirty = false; bool hb_dirty = false;
                int race = (int)DM_RACE(d); unsigned traj = DM_TTTRAJ(d);
                if ((mask & XGC_M_HIVTEST) && (vg[i] & FG_TESTED) && !(h & HV_DIAG)) {
                  if ((float)(at - (int)ha.x) >= P[XGC_P_test_window_int]) { h |= HV_DIAG; ha.y = (short)at; ha_dirty = true; }
                  else g.tck[idx].x = (short)at;
                }
                if (mask & XGC_M_HIVTX) {
                  if (h & HV_DIAG) {
                    if (!(h & HV_TX) && hb.x == 0) { if (bern(x.y, P[XGC_P_tx_init_prob + race])) { h |= HV_TX; ha.w = (short)at; ha_dirty = true; } }
                    else if (h & HV_TX) {
                      float rr = traj == 2 ? P[XGC_P_tx_halt_full_rr + race] : traj == 3 ? P[XGC_P_tx_halt_dur_rr + race] : 1.f;
                      if (bern(x.y, P[XGC_P_tx_halt_part_prob + race] * rr)) h &= ~HV_TX;
                    } else {
                      float rr = traj == 2 ? P[XGC_P_tx_reinit_full_rr + race] : traj == 3 ? P[XGC_P_tx_reinit_dur_rr + race] : 1.f;
                      if (bern(x.y, P[XGC_P_tx_reinit_part_prob + race] * rr)) h |= HV_TX;
                    }
                  }
                  if (h & HV_TX) hb.x++; else hb.y++;
                  hb_dirty = true;
                }
                float tsi = (float)(at - (int)ha.x);
                if (mask & XGC_M_HIVPROGRESS) {
                  unsigned stage = HV_STAGE(h);
                  if (stage == 1 && tsi >= P[XGC_P_vl_acute_rise_int]) { stage = 2; ha.z = (short)at; ha_dirty = true; }
                  if (stage == 2 && tsi >= P[XGC_P_vl_acute_rise_int] + P[XGC_P_vl_acute_fall_int]) { stage = 3; ha.z = (short)at; ha_dirty = true; }
                  if (stage == 3) {
                    bool aids = false; float on = (float)hb.x; float off = (float)hb.y;
                    if (on == 0.f && tsi >= P[XGC_P_vl_aids_onset_int]) aids = true;
                    if (traj == 1 && on > 0.f && off / P[XGC_P_max_time_off_tx_part_int] + on / P[XGC_P_max_time_on_tx_part_int] >= 1.f) aids = true;
                    if (traj >= 2 && !(h & HV_TX) && on > 0.f && off >= P[XGC_P_max_time_off_tx_full_int]) aids = true;
                    if (aids) { stage = 4; ha.z = (short)at; ha_dirty = true; }
                  }
                  h = (h & ~HV_STAGE_MASK) | (stage << 1);
                }
                if (mask & XGC_M_HIVVL) {
                  unsigned stage = HV_STAGE(h);
                  float v; float peak = P[XGC_P_vl_acute_peak]; float sp = P[XGC_P_vl_set_point];
                  if (!(h & HV_TX)) {
                    if (stage == 1) v = fminf(peak, peak * (tsi / P[XGC_P_vl_acute_rise_int]));
                    else if (stage == 2) v = fmaxf(sp, peak - (peak - sp) * ((tsi - P[XGC_P_vl_acute_rise_int]) / P[XGC_P_vl_acute_fall_int]));
                    else if (stage == 3) v = hb.x == 0 ? sp : fminf(sp, vl0 + P[XGC_P_vl_tx_up_slope]);
                    else v = fminf(P[XGC_P_vl_aids_peak], sp + ((float)(at - (int)ha.z) / P[XGC_P_vl_aids_int]) * (P[XGC_P_vl_aids_peak] - sp));
                  } else {
                    float target = traj == 1 ? P[XGC_P_vl_part_supp] : P[XGC_P_vl_full_supp];
                    v = fmaxf(target, vl0 - (stage == 4 ? P[XGC_P_vl_tx_aids_down_slope] : P[XGC_P_vl_tx_down_slope]));
                  }
                  if (v != vl0) g.aux[idx].vl = v;
                  h = (h & ~HV_VSUPP) | (v <= (float)XGC_LOG10_200 ? HV_VSUPP : 0u);
                }
                if (ha_dirty) g.hck_a[idx] = ha;
                if (hb_dirty) g.hck_b[idx] = hb;
                if (h != h0) { vh[i] = h; if (wt) core_w[4 * idx + 3] = h; if (incr) tally_move(F.row, d, 0u, h0, d, 0u, h); }
              }
            }
          );
        }
        __syncthreads();
        TICK(PH_PRE);
      } else {
        for (int k = tid; k < XGC_NCOUNTER; k += FT)            // continuing a week started by an earlier call
          if (!incr || k >= XGC_K_NGROUP * XGC_NCELL || !((STATE_GROUPS >> (k / XGC_NCELL)) & 1u)) F.row[k] = grow[k];
        __syncthreads();
      }
      // =================================================================================================== HOST NETWORK
      // The edge-list operations the host queued for this week (weekly toggles of the persistent layers, the new one-off list),
      // same semantics as k_edges_update / k_edges_upload: removal indices are ranks among the ties whose endpoints both live,
      // removed ties and the ties of departed agents become tombstones (squeezed out before the write-back, kept ties stay in
      // order), added ties go to the tail, R positions -> slots through a rank structure over the staged `active` bits.
      if (at == A.at0 && (A.hn_kind[0] | A.hn_kind[1] | A.hn_kind[2])) {
        // Built for latency: the operation headers in one round trip to memory, then EVERY other global load of the phase (tie
        // endpoints of both persistent layers, removal indices, added pairs, the one-off pairs) in a second one; the rest is
        // shared-memory work on bitmaps (living agents, living ties) with popcount prefixes and rank -> item selection.
        unsigned* const aw = reinterpret_cast<unsigned*>(&F.mq[0][0][0]);       // the mini-queues are idle between phases
        const int n_slots = F.sc[SC_NSLOTS], words = (n_slots + 31) >> 5, na = F.sc[SC_NALIVE];
        unsigned* const apre = aw + words;
        int* const hdr = F.wsum2;                                 // [0..1] n_removed, [2..3] n_added (as sent), [4] one-off n
        if (tid < 4) { const bool l1 = tid & 1; hdr[tid] = (l1 ? A.hn_kind[1] : A.hn_kind[0]) == 1 ? ((l1 ? A.hn_ptr[1] : A.hn_ptr[0]) + (size_t)r * (l1 ? A.hn_stride[1] : A.hn_stride[0]))[tid >> 1] : 0; }
        if (tid == 32) hdr[4] = A.hn_kind[0] == 2 ? A.hn_ne[0][r] : 0;                                           // [4..6]: whole-list sizes
        if (tid == 33) hdr[5] = A.hn_kind[1] == 2 ? A.hn_ne[1][r] : 0;
        if (tid == 34) hdr[6] = A.hn_kind[2] == 2 ? A.hn_ne[2][r] : 0;
        for (int k = warp; k < words; k += NW) {
          const int i = k * 32 + lane;
          const unsigned m = __ballot_sync(FULL, i < n_slots && (vd[i] & DM_ACTIVE));
          if (lane == 0) aw[k] = m;
        }
        __syncthreads();
        {                                                       // exclusive popcount prefix (words <= 1024 = FT)
          const int c = tid < words ? __popc(aw[tid]) : 0; int inc = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
          if (lane == 31) F.wsum[warp] = inc;
          __syncthreads();
          int tot; const int base = warp_base(F.wsum, warp, lane, &tot);
          if (tid < words) apre[tid] = (unsigned)(base + inc - c);
          __syncthreads();
        }
        // rank -> item over a bitmap with an exclusive popcount prefix: the last word whose prefix is <= p holds item p
        auto select = [&](const unsigned* bm, const unsigned* pre, int nw, int p) -> int {
          int lo = 0, hi = nw - 1;
          while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)pre[mid] <= p) lo = mid; else hi = mid - 1; }
          unsigned w = bm[lo];
          for (int j = p - (int)pre[lo]; j > 0; --j) w &= w - 1u;            // drop the j lowest set bits
          return lo * 32 + __ffs(w) - 1;
        };
        auto p2s = [&](int p) -> int { return select(aw, apre, words, p); };   // slot of the p-th (0-based) living agent
        bool bad = false;
        int ne_l[2], nrem_l[2], nadd_l[2], kw_l[2];
        unsigned* lm_l[2]; unsigned* lp_l[2];
        { unsigned* nx = apre + words;
#pragma unroll
  #pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int stride = A.hn_stride[l], hs = A.hn_has_start[l], on = A.hn_kind[l] == 1;
            ne_l[l] = on ? F.sc[SC_NE0 + l] : 0; kw_l[l] = (ne_l[l] + 31) >> 5;
            nrem_l[l] = on ? max(0, min(hdr[l], stride - 2)) : 0;
            nadd_l[l] = on ? max(0, min(hdr[2 + l], (stride - 2 - nrem_l[l]) / (hs ? 3 : 2))) : 0;
            if (on && (hdr[l] != nrem_l[l] || hdr[2 + l] != nadd_l[l])) bad = true;
            lm_l[l] = nx; nx += kw_l[l]; lp_l[l] = nx; nx += kw_l[l];
          } }
        // ---- second round trip: everything else this phase reads from global memory
        int rem_j[2] = { -1, -1 }; int add_a[2] = { 0, 0 }, add_b[2] = { 0, 0 }, add_s[2] = { 0, 0 };      // this thread's removal / added pair of each layer
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          if (A.hn_kind[l] != 1) continue;
          const int* row = A.hn_ptr[l] + (size_t)r * A.hn_stride[l];
          if (tid < nrem_l[l]) rem_j[l] = row[2 + tid];
          const int* tl = row + 2 + nrem_l[l];
          if (tid < nadd_l[l]) { add_a[l] = tl[tid]; add_b[l] = tl[nadd_l[l] + tid]; add_s[l] = A.hn_has_start[l] ? tl[2 * nadd_l[l] + tid] : at; }
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {                             // living-tie bitmaps; ties of departed agents become tombstones
          if (A.hn_kind[l] != 1) continue;
          const int ne = ne_l[l];
          const int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
          unsigned short* const eal = ea + g.e_off[l];
          for (int e0 = 0; e0 < ne; e0 += 4 * FT) {
            int2 xy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int e = e0 + k * FT + tid; xy[k] = e < ne ? *reinterpret_cast<const int2*>(E + e) : make_int2(0, 0); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = e0 + k * FT + tid;
              const bool live = e < ne && (vd[xy[k].x] & vd[xy[k].y] & DM_ACTIVE);
              const unsigned m = __ballot_sync(FULL, live);
              if (lane == 0 && (e >> 5) < kw_l[l]) lm_l[l][e >> 5] = m;
              if (e < ne && !live) eal[e] = EA_DEAD;
            }
          }
        }
        __syncthreads();
        if (warp < 2 && (warp ? A.hn_kind[1] : A.hn_kind[0]) == 1) {      // popcount prefix of layer `warp`'s living-tie bitmap, by one warp
          const int kw = warp ? kw_l[1] : kw_l[0], per = (kw + 31) >> 5;
          const unsigned* const lm = warp ? lm_l[1] : lm_l[0]; unsigned* const lp = warp ? lp_l[1] : lp_l[0];
          int sum = 0;
          for (int j = 0; j < per; ++j) { const int idx = lane * per + j; sum += idx < kw ? __popc(lm[idx]) : 0; }
          int inc = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
          int run = inc - sum;
          for (int j = 0; j < per; ++j) { const int idx = lane * per + j; if (idx < kw) { lp[idx] = (unsigned)run; run += __popc(lm[idx]); } }
          if (lane == 31) F.wsum3[warp] = inc;                    // living ties of the layer
        }
        __syncthreads();
        bool need_sq = false;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          if (A.hn_kind[l] != 1) continue;
          const int nlive = F.wsum3[l], nrem = nrem_l[l];
          int nkill = 0;
          for (int k = tid; k < nrem; k += FT) {                  // (k = tid: prefetched; a longer removal list reads the rest here)
            const int j = k == tid ? rem_j[l] : (A.hn_ptr[l] + (size_t)r * A.hn_stride[l])[2 + k];
            if (j < 0 || j >= nlive) { bad = true; continue; }
            const int e = select(lm_l[l], lp_l[l], kw_l[l], j);
            const int ei = g.e_off[l] + e; const unsigned bit = (unsigned)EA_DEAD << (16 * (ei & 1));      // (act words are 16-bit: OR into the aligned pair)
            if (atomicOr(reinterpret_cast<unsigned*>(ea) + (ei >> 1), bit) & bit) continue;                   // listed twice
            ++nkill;
          }
          nkill = __reduce_add_sync(FULL, nkill);
          if (lane == 0 && nkill) atomicAdd(&F.sc[SC_DEAD0 + l], nkill);
          if (tid == 0 && ne_l[l] > nlive) atomicAdd(&F.sc[SC_DEAD0 + l], ne_l[l] - nlive);
          if (ne_l[l] + nadd_l[l] > g.e_cap[l] && (ne_l[l] > nlive || nrem > 0)) need_sq = true;
        }
        __syncthreads();
        if (need_sq) squeeze();                                  // (uniform: every thread derived it from the same shared values)
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          if (A.hn_kind[l] != 1) continue;
          const int nadd = nadd_l[l], tailpos = F.sc[SC_NE0 + l];
          int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
          for (int k = tid; k < nadd; k += FT) {
            int a, b, st;
            if (k == tid) { a = add_a[l]; b = add_b[l]; st = add_s[l]; }
            else { const int* tl = A.hn_ptr[l] + (size_t)r * A.hn_stride[l] + 2 + nrem_l[l]; a = tl[k]; b = tl[nadd + k]; st = A.hn_has_start[l] ? tl[2 * nadd + k] : at; }
            if (a < 1 || a > na || b < 1 || b > na) { bad = true; a = b = 1; }
            if (tailpos + k < g.e_cap[l]) { E[tailpos + k] = make_int4(p2s(a - 1), p2s(b - 1), st, 0); ea[g.e_off[l] + tailpos + k] = 0; }
          }
          __syncthreads();
          if (tid == 0) {
            int nn = tailpos + nadd;
            if (nn > g.e_cap[l]) { nn = g.e_cap[l]; atomicOr(&F.sc[SC_STATUS], (int)XGC_ST_EDGE_OVERFLOW); }
            F.sc[SC_NE0 + l] = nn;
          }
        }
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          if (A.hn_kind[l] != 2) continue;
          const int stride = A.hn_stride[l], n = hdr[4 + l], cap = g.e_cap[l] < stride ? g.e_cap[l] : stride, nn = n < cap ? n : cap;
          const int* el = A.hn_ptr[l] + (size_t)r * 2 * stride;
          int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
          for (int e = tid; e < nn; e += FT) {
            int a = el[e], b = el[stride + e];
            if (a < 1 || a > na || b < 1 || b > na) { bad = true; a = b = 1; }
            E[e] = make_int4(p2s(a - 1), p2s(b - 1), A.hn_start[l] ? A.hn_start[l][(size_t)r * stride + e] : at, 0); ea[g.e_off[l] + e] = 0;
          }
          if (tid == 0) { F.sc[SC_NE0 + l] = nn; F.sc[SC_DEAD0 + l] = 0; if (n > cap) atomicOr(&F.sc[SC_STATUS], (int)XGC_ST_EDGE_OVERFLOW); }
        }
        if (bad) atomicOr(&F.sc[SC_STATUS], (int)XGC_ST_BAD_EDGE);
        __syncthreads();
        if (surrogate) {                                         // the degrees counted at staging were those of the old lists
          for (int k = tid; k < g.n_cap / 4; k += FT) deg32[k] = 0u;
          __syncthreads();
          count_degrees();
          __syncthreads();
        }
        TICK(PH_HOSTNET);
      }
      // =================================================================================================== NETWORK
      // (tergmLite::delete_vertices for the departed + surrogate simnet_msm).  Nothing here walks the ties: deletions are
      // tombstones set by the tie sweep (fused into the acts pass when this call also runs the POST group), new ties are
      // appended at the tails.
      if (ph & (XF_PRUNE | XF_FORM)) {
        const bool form = (ph & XF_FORM) && surrogate;
        if (tid == 0) {
          const int n_alive = F.sc[SC_NALIVE];
          int need = 0, anydead = 0;
          for (int l = 0; l < 3; ++l) {
            int nf = 0;
            if (form) {
              float target = F.netp[l] * (float)n_alive * 0.5f;           // (staged: thread 0 holds the whole CTA at the next barrier)
              float mu = l == 2 ? target : target / F.netp[3 + l];
              float fl = floorf(mu);
              nf = n_alive < 2 ? 0 : (int)fl + (u01(frand(key, at, FS_NETN, l, 0).x) < mu - fl);
              if (l == 2) { F.sc[SC_NE2] = 0; F.sc[SC_DEAD2] = 0; }      // one-off contacts last one week: the layer starts empty
            }
            F.sc[SC_NFORM0 + l] = nf;
            anydead |= F.sc[SC_DEAD0 + l];
            if (F.sc[SC_NE0 + l] + nf > g.e_cap[l] && F.sc[SC_DEAD0 + l] > 0) need = 1;
          }
          if ((at & 15) == 0 && anydead) need = 1;
          F.sc[SC_SQUEEZE] = need;
        }
        __syncthreads();
        if (F.sc[SC_SQUEEZE]) squeeze();
        if (form) {                                              // uniform role-compatible pairs holding the mean degree
          const int nf0 = F.sc[SC_NFORM0], nf01 = nf0 + F.sc[SC_NFORM1], nf_all = nf01 + F.sc[SC_NFORM2];
          const int n_slots = F.sc[SC_NSLOTS];
          int tail0 = F.sc[SC_NE0], tail1 = F.sc[SC_NE1], tail2 = F.sc[SC_NE2];
          for (int j0 = 0; j0 < nf_all; j0 += FT) {
            const int j = j0 + tid; bool ok = false; int4 ed = make_int4(0, 0, at, 0);
            const int l = j >= nf01 ? 2 : j >= nf0 ? 1 : 0, jj = j - (l == 2 ? nf01 : l == 1 ? nf0 : 0);
            if (j < nf_all) {
              // XGC_FORM_TRIES proposals, as in the oracle.  Endpoints by rejection over the slots (= uniform over the living): a pair
              // with a dead slot is not a proposal and does not use up a try (two pairs per Philox block, a few spare blocks)
              int tries = 0;
              for (int blk = 0; blk < XGC_FORM_TRIES / 2 + 3 && !ok && tries < XGC_FORM_TRIES; ++blk) {
                uint4 x = frand(key, at, FS_FORM + l, (unsigned)jj, (unsigned)blk);
#pragma unroll
                for (int q = 0; q < 2 && !ok && tries < XGC_FORM_TRIES; ++q) {
                  int sa = (int)__umulhi(q ? x.z : x.x, (unsigned)n_slots), sb = (int)__umulhi(q ? x.w : x.y, (unsigned)(n_slots - 1));
                  if (sb >= sa) sb++;
                  const unsigned da = vd[sa], db = vd[sb];
                  if (!(da & db & DM_ACTIVE)) continue;
                  ++tries;
                  if (roles_compatible(da, db)) {
                    // degree / risk-conditioned acceptance; the acceptance uniform is built from the low halves of the two position
                    // words (the positions use their high bits)
                    const unsigned ga8 = deg8[sa], gb8 = deg8[sb];
                    const float pw = form_w(l, ga8, da) * form_w(l, gb8, db);
                    const unsigned xu = ((q ? x.z : x.x) << 16) | ((q ? x.w : x.y) & 0xFFFFu);
                    const bool full = l < 2 && ((((ga8 >> (4 * l)) & 15u) == 15u) || (((gb8 >> (4 * l)) & 15u) == 15u));      // 4-bit degree counters
                    if (!full && (pw >= 1.f || bern(xu, pw))) { ok = true; ed.x = min(sa, sb); ed.y = max(sa, sb); }
                  }
                }
              }
            }
            const unsigned m0 = __ballot_sync(FULL, ok && l == 0), m1 = __ballot_sync(FULL, ok && l == 1), m2 = __ballot_sync(FULL, ok && l == 2);
            if (lane == 0) { F.wsum[warp] = __popc(m0); F.wsum2[warp] = __popc(m1); F.wsum3[warp] = __popc(m2); }
            __syncthreads();
            int tot0, tot1, tot2;
            const int base0 = warp_base(F.wsum, warp, lane, &tot0), base1 = warp_base(F.wsum2, warp, lane, &tot1), base2 = warp_base(F.wsum3, warp, lane, &tot2);
            if (ok) {
              const unsigned m = l == 0 ? m0 : l == 1 ? m1 : m2;
              const int dst = (l == 0 ? tail0 + base0 : l == 1 ? tail1 + base1 : tail2 + base2) + __popc(m & ((1u << lane) - 1u));
              if (dst < g.e_cap[l]) { g.edge[(size_t)r * g.e_stride + g.e_off[l] + dst] = ed; ea[g.e_off[l] + dst] = 0;
                if (l < 2) { deg_add(ed.x, l); deg_add(ed.y, l); } }        // (every read of this chunk's decisions is behind the barrier above)
              else atomicOr(&F.sc[SC_STATUS], (int)XGC_ST_EDGE_OVERFLOW);
            }
            tail0 = min(tail0 + tot0, g.e_cap[0]); tail1 = min(tail1 + tot1, g.e_cap[1]); tail2 = min(tail2 + tot2, g.e_cap[2]);
            __syncthreads();
          }
          if (tid == 0) { F.sc[SC_NE0] = tail0; F.sc[SC_NE1] = tail1; F.sc[SC_NE2] = tail2; }
          __syncthreads();
        }
        // no acts pass in this call.  With a network step to run now (dissolution + formation) the tie sweep runs on its own; without
        // one, the ties of this week's departed agents are left to whoever touches the lists next -- the weekly-toggle kernel
        // (k_edges_update), this kernel's acts pass, or an explicit prune before anything else reads them (h->prune_pending)
        if (!(ph & XF_POST) && form) {
          const int ne0 = F.sc[SC_NE0], ne01 = ne0 + F.sc[SC_NE1], ne_all = ne01 + F.sc[SC_NE2];
          for (int f0 = tid; f0 < ne_all; f0 += 4 * FT) {        // four ties per thread in flight
            int4 ed4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int f = min(f0 + k * FT, ne_all - 1);
              const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
              ed4[k] = g.edge[(size_t)r * g.e_stride + sidx];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int4 ed = ed4[k]; const int f = f0 + k * FT, l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
              if (f < ne_all && !(ea[sidx] & EA_DEAD)) {
                bool dead = !(vd[ed.x] & vd[ed.y] & DM_ACTIVE);
                if (!dead && form && l < 2 && ed.z < at) dead = dissolves(frand(key, at, FS_ACTS, pairkey(ed), (unsigned)l), ed, l, at);
                if (dead) { ea[sidx] = EA_DEAD; atomicAdd(&F.sc[SC_DEAD0 + l], 1); if (surrogate && l < 2) { deg_sub(ed.x, l); deg_sub(ed.y, l); } }
              }
            }
          }
          __syncthreads();
          if (tid < 3 && form) F.row[XGC_K_NEDGES + tid] = (unsigned)(F.sc[SC_NE0 + tid] - F.sc[SC_DEAD0 + tid]);
        }
        TICK(PH_NET);
      }
      // =================================================================================================== POST
      if (ph & XF_POST) {
        const int n_slots = F.sc[SC_NSLOTS];
        const int AW = ((n_slots + FT - 1) / FT) * 32;
        const int ne0 = F.sc[SC_NE0], ne01 = ne0 + F.sc[SC_NE1], ne_all = ne01 + F.sc[SC_NE2];
        const bool last_week = at == ((A.ph_last & XF_POST) ? A.at1 : A.at1 - 1);      // the last week of the call that has an acts pass
        if (tid == 0 && (mask & XGC_M_STITRANS)) {              // weekly random pathway order (stitrans_msm_rand)
          uint4 x0 = frand(key, at, FS_PATH, 0, 0), x1 = frand(key, at, FS_PATH, 1, 0);
          unsigned dr[6] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y };
          int perm[XGC_NPATH];
          for (int k = 0; k < XGC_NPATH; ++k) perm[k] = k;
          for (int i = XGC_NPATH - 1; i >= 1; --i) { int j = (int)__umulhi(dr[XGC_NPATH - 1 - i], (unsigned)(i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
          for (int k = 0; k < XGC_NPATH; ++k) F.sc[SC_PATH + perm[k]] = k;
        }
        // ---------------------------------------------------------------- acts_msm + exposure history + PrEP risk history
        // draws of tie f: FS_ACTS block = { anal count, oral count, kissing (low 16) | rimming (high 16), condom use of anal acts
        // 0 / 1 (16 bits each) }; anal acts 2.. take their condom draws from FS_COND blocks, 8 acts per block
        if (mask & (XGC_M_ACTS | XGC_M_PREP)) {
          const float sc_ai = P[XGC_P_ai_acts_scale_mc] * (1.f / 52.f), sc_oi = P[XGC_P_oi_acts_scale_mc] * (1.f / 52.f);
          unsigned tot_ai = 0, tot_oi = 0;
          const bool diss_on = (ph & XF_FORM) && surrogate;
          for (int f = tid; f < ne_all; f += FT) {
            const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
            if (ea[sidx] & EA_DEAD) continue;
            int4* Ep = g.edge + (size_t)r * g.e_stride + sidx;
            const int4 ed = *Ep;
            const unsigned da = vd[ed.x], db = vd[ed.y], ga = vg[ed.x], gb = vg[ed.y], ha = vh[ed.x], hb = vh[ed.y];
            const uint4 x = frand(key, at, FS_ACTS, pairkey(ed), (unsigned)l);
            // the tie sweep: ties of departed agents (tergmLite::delete_vertices) and this week's dissolutions become tombstones
            if (!(da & db & DM_ACTIVE) || (diss_on && l < 2 && ed.z < at && dissolves(x, ed, l, at))) {
              ea[sidx] = EA_DEAD; atomicAdd(&F.sc[SC_DEAD0 + l], 1);
              if (surrogate && l < 2) { deg_sub(ed.x, l); deg_sub(ed.y, l); }
              continue;
            }
            const float ca = (float)(2 * at8 - birth8(ha) - birth8(hb)) * (float)XF_W8;
            int ai = 0, oi = 0, ki = 0, ri = 0;
            if (mask & XGC_M_ACTS) {
              if (!(act_stopped(da, ga) || act_stopped(db, gb))) {
                const bool compat = roles_compatible(da, db);
                if (l < 2) {
                  float dur = (float)(at - ed.z);
                  bool hcp = (ha & HV_DIAG) && (hb & HV_DIAG);
                  int r1 = (int)DM_RACE(da), r2 = (int)DM_RACE(db);
                  float rai = __expf(glm_f(F.G, dur, l == 1, ca, hcp, false, r1, r2)) * sc_ai;
                  float roi = __expf(glm_f(F.G + XGC_GLM_NCOEF, dur, l == 1, ca, hcp, false, r1, r2)) * sc_oi;
                  ai = compat ? pois_f(u01(x.x), rai, XGC_MAX_ACTS) : 0;
                  oi = pois_f(u01(x.y), roi, XGC_MAX_ACTS);
                  // kissing / rimming: "any act this week" here; the counts are drawn in the transmission phase from the same bits
                  ki = u01_16(x.z & 0xFFFFu) > F.pk[l * (XGC_MAX_ACTS + 1)];
                  ri = u01_16(x.z >> 16) > F.pk[(2 + l) * (XGC_MAX_ACTS + 1)];
                } else {
                  ai = compat ? 1 : 0;
                  oi = bern(x.y, P[XGC_P_oi_prob_oo]); ki = u01_16(x.z & 0xFFFFu) < P[XGC_P_kiss_prob_oo]; ri = u01_16(x.z >> 16) < P[XGC_P_rim_prob_oo];
                }
              }
              ea[sidx] = (unsigned short)(ai | (oi << 6) | (ki << 12) | (ri << 13));
              if (last_week && A.write_acts) Ep->w = (int)((unsigned)ai | ((unsigned)oi << 8) | (0x7FFFu << 16));
              tot_ai += ai; tot_oi += oi;
#pragma unroll
              for (int side = 0; side < 2; ++side) {            // exposure history for the CDC screening protocol
                unsigned role = DM_ROLE(side ? db : da), ex = 0;
                if (ai > 0 && role != 0) ex |= 1;
                if ((ai > 0 && role != 1) || oi > 0) ex |= 2;
                if (oi > 0 || ri > 0) ex |= 4;
                if (ki > 0) ex |= 8;
                if (l != 0 && (ai | oi | ki | ri)) ex |= 16;
                if ((ex << GC_EXPO_SHIFT) & ~(side ? gb : ga)) atomicOr(&vg[side ? ed.y : ed.x], ex << GC_EXPO_SHIFT);
              }
            } else ai = ea[sidx] & 63;
            if (mask & XGC_M_PREP) {                            // risk history: condomless anal sex, or anal sex with a diagnosed partner
              const bool ua = !(ha & HV_DIAG), ub = !(hb & HV_DIAG);
              if (ai > 0 && (ua || ub)) {
                bool uai = false;
                if (ua && ub) {                                 // (with a diagnosed partner the indication does not depend on condom use)
                  bool prep = (ha | hb) & HV_PREP;
                  float eta = l < 2 ? glm_f(F.G + 2 * XGC_GLM_NCOEF, (float)(at - ed.z), l == 1, ca, false, prep, (int)DM_RACE(da), (int)DM_RACE(db))
                                    : glm_f(F.G + 3 * XGC_GLM_NCOEF, 0.f, false, ca, false, prep, (int)DM_RACE(da), (int)DM_RACE(db));
                  float pc = fminf(1.f, P[XGC_P_cond_scale] / (1.f + __expf(-eta)));
                  unsigned thr = (unsigned)((1.f - pc) * 65536.f);
                  uai = (x.w & 0xFFFFu) < thr || (ai > 1 && (x.w >> 16) < thr);
                  for (int k0 = 2; k0 < ai && !uai; k0 += 8) {
                    uint4 c = frand(key, at, FS_COND, pairkey(ed), (unsigned)l | ((unsigned)((k0 - 2) >> 3) << 2));
                    unsigned w[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k0 + k < ai && ((w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu) < thr) uai = true;
                  }
                }
                if (ua && (uai || !ub)) { ind[ed.x] = (short)at; g.tck[rbase + ed.x].y = (short)at; }
                if (ub && (uai || !ua)) { ind[ed.y] = (short)at; g.tck[rbase + ed.y].y = (short)at; }
              }
            }
          }
          if (mask & XGC_M_ACTS) {
            tot_ai = __reduce_add_sync(FULL, tot_ai); tot_oi = __reduce_add_sync(FULL, tot_oi);
            if (lane == 0) { if (tot_ai) atomicAdd(&F.row[XGC_K_NACTS], tot_ai); if (tot_oi) atomicAdd(&F.row[XGC_K_NACTS + 1], tot_oi); }
          }
        }
        __syncthreads();
        if (tid < 3 && (ph & XF_FORM) && surrogate) F.row[XGC_K_NEDGES + tid] = (unsigned)(F.sc[SC_NE0 + tid] - F.sc[SC_DEAD0 + tid]);
        TICK(PH_ACTS);
        // ---------------------------------------------------------------- prep_msm, stirecov_msm, stitx_msm (per-warp agent ranges)
        if (mask & (XGC_M_PREP | XGC_M_STIRECOV | XGC_M_STITX | XGC_M_HIVTRANS)) {
          const bool avd_on = (mask & XGC_M_STITX) && (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL);
          if (avd_on) gap_warp(key, at, FS_AVD, (unsigned)warp, 0, AW, P[XGC_P_gc_asympt_visit_prob], lane, [&](bool ok, int p, const uint4&) {
            const int i = VSLOT(p); if (ok && i < n_slots) vg[i] |= FG_AVD; });
          __syncwarp();
          int c0 = 0, c1 = 0;                                   // q0 = PrEP users / starters (processed first), q1 = agents with GC work
          unsigned aword = 0u;                                  // lane k: GC-work bits of word k of this warp's range
          for (int k = 0; k * 32 < AW; ++k) {
            const int i = (k * NW + warp) * 32 + lane; bool qA = false, qP = false;
            if (i < n_slots && (vd[i] & DM_ACTIVE)) {
              unsigned gc = vg[i], h = vh[i];
              // start-of-pass snapshot: what acts / condoms / hivtrans see (they precede prep / stirecov / stitx in the module order)
              unsigned gc1 = (gc & ~(GC_CUR_MASK << GC_PRE_SHIFT)) | ((gc & GC_CUR_MASK) << GC_PRE_SHIFT);
              unsigned h1 = (h & ~HV_PREP_PRE) | ((h & HV_PREP) ? HV_PREP_PRE : 0u);
              if (mask & XGC_M_PREP) {
                bool diag = h & HV_DIAG; short il = ind[i];
                bool indic = il != XGC_NA16 && (float)il >= (float)at - P[XGC_P_prep_risk_int];
                h1 = (h1 & ~HV_PREPELIG) | ((!diag && indic) ? HV_PREPELIG : 0u);
                // a negative test this week (incl. a false negative inside the window period) is what opens PrEP start / reassessment
                qP = (h & HV_PREP) || (!diag && indic && (gc & FG_TESTED) && (float)at >= P[XGC_P_prep_start]);
              }
              if (gc1 != gc) vg[i] = gc1;
              if (h1 != h) { vh[i] = h1;
                if (incr && ((h1 ^ h) & HV_PREPELIG)) atomicAdd(&F.row[XGC_K_PREPELIG * XGC_NCELL + DM_CELL(vd[i])], (h1 & HV_PREPELIG) ? 1u : 0xFFFFFFFFu); }
              unsigned ex = (gc >> GC_EXPO_SHIFT) & 0x1Fu;
              bool cdc_active = proto == XGC_SCREEN_CDC && ((ex & 7u) || (kissx && (ex & 8u)));
              qA = (mask & (XGC_M_STIRECOV | XGC_M_STITX)) && ((gc & 7u) || ((mask & XGC_M_STITX) && ((gc & FG_AVD) || cdc_active)));
            }
            unsigned ma = __ballot_sync(FULL, qA);
            if (lane == k) aword = ma;
            WQ_PUSH(q0, c0, qP, i);
            WQ_DRAIN(q0, c0, (k + 1) * 32 >= AW, item,
              const int i = (int)item; const size_t idx = rbase + i;
              unsigned h = vh[i]; const unsigned h0 = h; int race = (int)DM_RACE(vd[i]);
              bool diag = h & HV_DIAG; bool indic = h & HV_PREPELIG; bool tested_neg = (vg[i] & FG_TESTED) && !diag;
              uint4 x = frand(key, at, FS_PREP, (unsigned)i, 0);
              if (h & HV_PREP) {
                bool stop = false;
                if (diag) stop = true;
                else if (bern(x.x, P[XGC_P_prep_discont_rate + race])) stop = true;
                else if (tested_neg) {
                  short4 tk = g.tck[idx];
                  if ((float)(at - (int)tk.z) >= P[XGC_P_prep_reassess_int]) { if (!indic) stop = true; else g.tck[idx].z = (short)at; }
                }
                if (stop) h &= ~(HV_PREP | HV_PREPCLASS_MASK);
              } else if (!diag && tested_neg && indic && bern(x.x, P[XGC_P_prep_start_prob + race])) {
                unsigned cl = 1 + categorical_f(u01(x.y), P + XGC_P_prep_adhr_dist, 3);
                h = (h & ~HV_PREPCLASS_MASK) | HV_PREP | (cl << 7);
                short4* tq = g.tck + idx; tq->z = (short)at; tq->w = (short)at;
              }
              if (h != h0) { vh[i] = h;      // (only prepStat / prepClass change here)
                if (incr && ((h ^ h0) & HV_PREP)) atomicAdd(&F.row[XGC_K_PREPCURR * XGC_NCELL + DM_CELL(vd[i])], (h & HV_PREP) ? 1u : 0xFFFFFFFFu); }
            );
          }
          for (int k = 0; k * 32 < AW; ++k) {                   // stirecov_msm + stitx_msm on infected or visiting agents
            const unsigned ma = __shfl_sync(FULL, aword, k);
            WQ_PUSH(q1, c1, (ma >> lane) & 1u, (k * NW + warp) * 32 + lane);
            WQ_DRAIN(q1, c1, (k + 1) * 32 >= AW, item,
              const int i = (int)item; const size_t idx = rbase + i;
              unsigned gc = vg[i]; const unsigned gc0 = gc; const unsigned h = vh[i];
              const bool avd = gc & FG_AVD;
              short4 ga = g.gck_a[idx]; short4 gb = g.gck_b[idx]; short4 gd = make_short4(-1, -1, -1, -1);
              if (pois_mode && (gc & 7u)) gd = g.gck_d[idx];
              bool ga_dirty = false; bool gb_dirty = false; bool gd_dirty = false; bool ind_sti = false;
              if ((mask & XGC_M_STIRECOV) && (gc & 7u)) {
                uint4 x = frand(key, at, FS_AG2, (unsigned)i, 0);
                const unsigned xr[3] = { x.x, x.y, x.z };
                const int txpr[3] = { XGC_P_rgc_tx_recov_pr, XGC_P_ugc_tx_recov_pr, XGC_P_pgc_tx_recov_pr };
_Pragma("unroll")
                for (int s = 0; s < 3; ++s) {
                  if (!(gc & GC_ST(s))) continue;
                  short infT = s == 0 ? ga.x : s == 1 ? ga.y : ga.z; short txT = s == 0 ? gb.x : s == 1 ? gb.y : gb.z; short dur = s == 0 ? gd.x : s == 1 ? gd.y : gd.z;
                  bool recov = false;
                  if (gc & GC_TX(s)) { int kk = at - (int)txT; if (kk >= 1) recov = bern(xr[s], P[txpr[s] + (kk >= 3 ? 2 : kk - 1)]); }
                  else if ((int)infT < at) recov = pois_mode ? (at - (int)infT) >= (int)dur : bern(xr[s], 1.f / P[XGC_P_gc_ntx_int + s]);
                  if (recov) {
                    gc &= ~(GC_ST(s) | GC_SY(s) | GC_TX(s));
                    if (s == 0) { ga.x = -1; gb.x = -1; gd.x = -1; } else if (s == 1) { ga.y = -1; gb.y = -1; gd.y = -1; } else { ga.z = -1; gb.z = -1; gd.z = -1; }
                    ga_dirty = gb_dirty = true;
                    if (pois_mode) gd_dirty = true; else (&g.gck_d[idx].x)[s] = -1;
                  }
                }
              }
              if (mask & XGC_M_STITX) {
                bool sv = false; bool av = false;
                if ((gc & 7u) & ((gc >> 3) & 7u) & ~((gc >> 6) & 7u)) {       // a symptomatic, untreated site: care seeking
                  uint4 x = frand(key, at, FS_AG2, (unsigned)i, 1);
                  const unsigned xs[3] = { x.x, x.y, x.z };
_Pragma("unroll")
                  for (int s = 0; s < 3; ++s)
                    if ((gc & GC_ST(s)) && (gc & GC_SY(s)) && !(gc & GC_TX(s))) {
                      float rr = s == 0 ? P[XGC_P_rgc_sympt_seek_test_rr] : s == 2 ? P[XGC_P_pgc_sympt_seek_test_rr] : 1.f;
                      if (bern(xs[s], P[XGC_P_ugc_sympt_seek_test_prob] * rr)) sv = true;
                    }
                }
                unsigned ex = (gc >> GC_EXPO_SHIFT) & 0x1Fu;
                if (!sv) {
                  if (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL) av = avd;
                  else if (proto == XGC_SCREEN_CDC) {
                    bool active = (ex & 7u) || (kissx && (ex & 8u)); bool hr = (h & HV_PREP) || (ex & 16u);
                    float interval = (hr ? P[XGC_P_cdc_sti_hr_int] : P[XGC_P_cdc_sti_int]) * (52.f / 12.f);
                    av = active && (ga.w == XGC_NA16 || (float)(at - (int)ga.w) >= interval);
                  }
                }
                if (sv || av) {
                  atomicAdd(&F.row[XGC_K_VISITS], 1u); if (sv) atomicAdd(&F.row[XGC_K_VISITS_SYMPT], 1u);
                  uint4 x = frand(key, at, FS_AG2, (unsigned)i, 2);
                  const unsigned xt[3] = { x.x, x.y, x.z };
_Pragma("unroll")
                  for (int s = 0; s < 3; ++s) {
                    bool inf = gc & GC_ST(s); bool symp = inf && (gc & GC_SY(s)); bool tested;
                    bool exposed = (ex >> s) & 1u; if (s == 2 && kissx && (ex & 8u)) exposed = true;
                    if (proto == XGC_SCREEN_CDC && !symp && !exposed) tested = false;
                    else if (proto == XGC_SCREEN_UNIVERSAL) tested = true;
                    else tested = bern(xt[s], symp ? P[XGC_P_gc_sympt_prob_test + s] : P[XGC_P_gc_asympt_prob_test + s]);
                    if (!tested) continue;
                    atomicAdd(&F.row[XGC_K_TESTED + s], 1u);
                    if (inf) {
                      atomicAdd(&F.row[XGC_K_POS + s], 1u); ind_sti = true;
                      if (!(gc & GC_TX(s))) { gc |= GC_TX(s); atomicAdd(&F.row[XGC_K_TXSTART + s], 1u); gb_dirty = true;
                        if (s == 0) gb.x = (short)at; else if (s == 1) gb.y = (short)at; else gb.z = (short)at; }
                    }
                  }
                  ga.w = (short)at; ga_dirty = true;
                  gc &= ~GC_EXPO_MASK;
                }
              }
              if (ind_sti) { ind[i] = (short)at; g.tck[idx].y = (short)at; }
              if (ga_dirty) g.gck_a[idx] = ga;
              if (gb_dirty) g.gck_b[idx] = gb;
              if (gd_dirty) g.gck_d[idx] = gd;
              // other threads OR exposure / new-infection bits into this word only in other phases: a plain store is safe here
              if (gc != gc0) { vg[i] = gc;
                if (incr && ((gc ^ gc0) & 7u)) {                // recoveries: the site tallies, and the any-site tally when the last one clears
                  const unsigned cell = DM_CELL(vd[i]); unsigned rec = (gc ^ gc0) & 7u;
                  if (!(gc & 7u)) atomicSub(&F.row[XGC_K_INUM_GC * XGC_NCELL + cell], 1u);
                  while (rec) { const int sb = __ffs(rec) - 1; rec &= rec - 1; atomicSub(&F.row[(XGC_K_INUM_RGC + sb) * XGC_NCELL + cell], 1u); }
                } }
            );
          }
        }
        __syncthreads();
        TICK(PH_B);
        // ---------------------------------------------------------------- hivtrans_msm + stitrans_msm_rand over discordant ties
        // every warp classifies its ties 32 at a time and processes the (tie, act type) pairs that can transmit from two
        // mini-queues: anal pairs, and the light pairs -- oral, rimming, kissing (item bits 14..15), one branch-free body
        if (mask & (XGC_M_HIVTRANS | XGC_M_STITRANS)) {
          const bool gct = mask & XGC_M_STITRANS, hvt = mask & XGC_M_HIVTRANS;
          const bool kiss_on = gct && route_kiss, rim_on = gct && route_rim;
          const int ngroups = (ne_all + 31) >> 5;
          const int niter = (ngroups + NW - 1) / NW;
          const int4* const Er = g.edge + (size_t)r * g.e_stride;
          int c0 = 0, c1 = 0;
          for (int it = 0; it < niter; ++it) {
            const bool lastit = it == niter - 1;
            const int f = (it * NW + warp) * 32 + lane; bool qa = false, qo = false, qk = false, qr = false;
            if (f < ne_all) {
              const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
              const unsigned eaf = ea[sidx];
              if (eaf & 0x3FFFu) {                                // a living tie with any act this week
                const int4 ed = Er[sidx];
                const int ai = eaf & 63, oi = (eaf >> 6) & 63;
                const unsigned x = vg[ed.x] & 7u, y = vg[ed.y] & 7u;
                const bool hivdisc = hvt && ai > 0 && ((vh[ed.x] ^ vh[ed.y]) & HV_STATUS);
                const bool d_ai = (((x >> 1) ^ y) | ((y >> 1) ^ x)) & 1u, d_oi = (((x >> 1) ^ (y >> 2)) | ((y >> 1) ^ (x >> 2))) & 1u;
                const bool d_ri = (((x >> 2) ^ y) | ((y >> 2) ^ x)) & 1u, d_ki = ((x ^ y) >> 2) & 1u;
                qa = hivdisc || (gct && ai > 0 && d_ai);
                qo = gct && oi > 0 && d_oi;
                qk = ((eaf >> 12) & 1u) && kiss_on && d_ki;
                qr = ((eaf >> 13) & 1u) && rim_on && d_ri;
              }
            }
            {
              WQ_PUSH(q0, c0, qa, f);
              // ---- queue 0: anal acts
              WQ_DRAIN(q0, c0, lastit, item,
                const int f = (int)item;
                const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0; const int sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
                const int4 ed = Er[sidx];
                const unsigned gca = vg[ed.x]; const unsigned gcb = vg[ed.y];
                const unsigned ga = gca & 7u; const unsigned gb = gcb & 7u;
                unsigned new_a = 0; unsigned new_b = 0; bool hiv_a = false; bool hiv_b = false;
                const unsigned pk2 = pairkey(ed);
                const uint4 xa = frand(key, at, FS_ACTS, pk2, (unsigned)l);
                {
                  const unsigned da = vd[ed.x]; const unsigned db = vd[ed.y]; const unsigned ha = vh[ed.x]; const unsigned hb = vh[ed.y];
                  const int ai = ea[sidx] & 63;
                  const bool hivdisc = hvt && ((ha ^ hb) & HV_STATUS);
                  const bool d_ai = gct && ((((ga >> 1) ^ gb) | ((gb >> 1) ^ ga)) & 1u);
                  // condom use: probabilities as condoms_msm saw them (PrEP status before this week's prep_msm)
                  float ca = (float)(2 * at8 - birth8(ha) - birth8(hb)) * (float)XF_W8;
                  bool prep = (ha | hb) & HV_PREP_PRE;
                  float eta = l < 2 ? glm_f(F.G + 2 * XGC_GLM_NCOEF, (float)(at - ed.z), l == 1, ca, (ha & HV_DIAG) && (hb & HV_DIAG), prep, (int)DM_RACE(da), (int)DM_RACE(db))
                                    : glm_f(F.G + 3 * XGC_GLM_NCOEF, 0.f, false, ca, (ha & HV_DIAG) && (hb & HV_DIAG), prep, (int)DM_RACE(da), (int)DM_RACE(db));
                  float pc = fminf(1.f, P[XGC_P_cond_scale] / (1.f + __expf(-eta)));
                  const unsigned cthr = (unsigned)((1.f - pc) * 65536.f);
                  const unsigned r1 = DM_ROLE(da); const unsigned r2 = DM_ROLE(db);
                  const int fixed = (r1 == 0 || r2 == 1) ? 1 : (r1 == 1 || r2 == 0) ? 0 : -1;
                  float q1f = (float)(da >> 16); float q2f = (float)(db >> 16);
                  const float pins = q1f + q2f > 0.f ? q1f / (q1f + q2f) : 0.5f;
                  const bool a_pos = ha & HV_STATUS;
                  const unsigned hp = a_pos ? ha : hb; const unsigned hn = a_pos ? hb : ha; const unsigned dn = a_pos ? db : da; const unsigned gn = a_pos ? gcb : gca;
                  float vlf = 0.f; const int nrace = (int)DM_RACE(dn);
                  if (hivdisc) vlf = exp2f((g.aux[rbase + (a_pos ? ed.x : ed.y)].vl - 4.5f) * 1.2927817f);      // 2.45^(vl - 4.5)
                  const unsigned gpre_n = (gn >> GC_PRE_SHIFT) & 7u;
                  uint4 c = make_uint4(0, 0, 0, 0);
                  for (int k = 0; k < ai; ++k) {
                    unsigned c16;
                    if (k < 2) c16 = k ? xa.w >> 16 : xa.w & 0xFFFFu;
                    else { const int kk = k - 2;
                      if ((kk & 7) == 0) c = frand(key, at, FS_COND, pk2, (unsigned)l | ((unsigned)(kk >> 3) << 2));
                      const unsigned cw = (kk & 7) < 2 ? c.x : (kk & 7) < 4 ? c.y : (kk & 7) < 6 ? c.z : c.w;
                      c16 = (cw >> ((kk & 1) * 16)) & 0xFFFFu; }
                    const bool uai = c16 < cthr;
                    const uint4 x = frand(key, at, FS_ANAL, pk2, (unsigned)l | ((unsigned)k << 2));
                    const bool p1ins = fixed >= 0 ? fixed != 0 : bern(x.x, pins);
                    if (hivdisc) {
                      const bool pos_ins = p1ins == a_pos;
                      float t;
                      if (pos_ins) { t = P[XGC_P_urai_prob] * vlf; if (gpre_n & 1u) t *= P[XGC_P_hiv_rgc_rr]; }
                      else { t = P[XGC_P_uiai_prob] * vlf; if (dn & DM_CIRC) t *= P[XGC_P_circ_rr]; if (gpre_n & 2u) t *= P[XGC_P_hiv_ugc_rr]; }
                      if (!uai) t *= 1.f - P[XGC_P_cond_eff] * (1.f - P[XGC_P_cond_fail + nrace]);
                      unsigned st = HV_STAGE(hp);
                      if (st == 1 || st == 2) t *= P[XGC_P_acute_rr];
                      if (hn & HV_PREP) { unsigned cl = HV_PREPCLASS(hn); if (cl >= 1) t *= P[XGC_P_prep_adhr_hr + cl - 1]; }
                      t *= P[XGC_P_trans_scale + nrace];
                      if (bern(x.y, t)) { if (a_pos) hiv_b = true; else hiv_a = true; }
                    }
                    if (d_ai) {                                  // insertive partner's urethra <-> receptive partner's rectum
                      const unsigned gx = p1ins ? ga : gb; const unsigned gy = p1ins ? gb : ga;
                      const bool ux = (gx >> 1) & 1u; const bool ry = gy & 1u;
                      if (ux != ry) {
                        float t = ux ? P[XGC_P_u2rgc_tprob] : P[XGC_P_r2ugc_tprob];
                        const unsigned dsus = (ux ? !p1ins : p1ins) ? da : db;       // the susceptible partner: receptive for u2r, insertive for r2u
                        if (!uai) t *= 1.f - P[XGC_P_sti_cond_eff] * (1.f - P[XGC_P_sti_cond_fail + (int)DM_RACE(dsus)]);
                        if (bern(x.z, t)) {
                          if (ux) { if (p1ins) new_b |= 1u << XGC_PATH_U2R; else new_a |= 1u << XGC_PATH_U2R; }
                          else { if (p1ins) new_a |= 1u << XGC_PATH_R2U; else new_b |= 1u << XGC_PATH_R2U; }
                        }
                      }
                    }
                  }
                }
                if (new_a) atomicOr(&vg[ed.x], new_a << GC_NEW_SHIFT);
                if (new_b) atomicOr(&vg[ed.y], new_b << GC_NEW_SHIFT);
                if (hiv_a) atomicOr(&vh[ed.x], HV_NEW);
                if (hiv_b) atomicOr(&vh[ed.y], HV_NEW);
              );
            }
#pragma unroll 1
            for (int ph = 0; ph < 3; ++ph) {
              WQ_PUSH(q1, c1, ph == 0 ? qo : ph == 1 ? qr : qk, f | (ph << 14));
              // ---- queue 1: oral acts (urethra <-> pharynx), rimming (pharynx <-> rectum), kissing (pharynx <-> pharynx): the m acts
              //      of one direction are ONE Bernoulli(1 - (1 - t)^m) trial
              WQ_DRAIN(q1, c1, lastit && ph == 2, item,
                const int f = (int)(item & 0x3FFFu); const unsigned ty = item >> 14;
                const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0; const int sidx = g.e_off[l] + f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
                const int4 ed = Er[sidx];
                const unsigned ga = vg[ed.x] & 7u; const unsigned gb = vg[ed.y] & 7u;
                const unsigned pk2 = pairkey(ed);
                int n = (ea[sidx] >> 6) & 63;
                if (ty) { n = 1;                                  // rimming / kissing counts: drawn here from the bits acts_msm tested for "any"
                  if (l < 2) { const uint4 xa = frand(key, at, FS_ACTS, pk2, (unsigned)l);
                    n = pois_tab_f(u01_16(ty == 1 ? xa.z >> 16 : xa.z & 0xFFFFu), F.pk + ((ty == 1 ? 2 : 0) + l) * (XGC_MAX_ACTS + 1)); } }
                const int sx = ty == 0 ? 1 : 2; const int sy = ty == 0 ? 2 : ty == 1 ? 0 : 2;
                const float tA = P[ty == 0 ? XGC_P_u2pgc_tprob : ty == 1 ? XGC_P_p2rgc_tprob : XGC_P_p2pgc_tprob];
                const float tB = P[ty == 0 ? XGC_P_p2ugc_tprob : ty == 1 ? XGC_P_r2pgc_tprob : XGC_P_p2pgc_tprob];
                const unsigned pA = ty == 0 ? XGC_PATH_U2P : ty == 1 ? XGC_PATH_P2R : XGC_PATH_P2P; const unsigned pB = ty == 0 ? XGC_PATH_P2U : ty == 1 ? XGC_PATH_R2P : XGC_PATH_P2P;
                const uint4 x = frand(key, at, FS_ORAL + ty, pk2, (unsigned)l);
                unsigned new_a = 0; unsigned new_b = 0;
                if (n > 0) {
                  const int m1 = __popc(x.x >> (32 - n)); const int m0 = n - m1;       // acts with p1 / p2 the active partner
                  { bool ix = (ga >> sx) & 1u; bool iy = (gb >> sy) & 1u;
                    if (m1 > 0 && ix != iy && bern(x.y, any_of_n_f(ix ? tA : tB, m1))) { if (ix) new_b |= 1u << pA; else new_a |= 1u << pB; } }
                  { bool ix = (gb >> sx) & 1u; bool iy = (ga >> sy) & 1u;
                    if (m0 > 0 && ix != iy && bern(x.z, any_of_n_f(ix ? tA : tB, m0))) { if (ix) new_a |= 1u << pA; else new_b |= 1u << pB; } }
                }
                if (new_a) atomicOr(&vg[ed.x], new_a << GC_NEW_SHIFT);
                if (new_b) atomicOr(&vg[ed.y], new_b << GC_NEW_SHIFT);
              );
            }
          }
        }
        __syncthreads();
        TICK(PH_TRANS);
        // ---------------------------------------------------------------- apply new infections (per-warp agent ranges)
        {
          int c0 = 0;
          for (int k = 0; k * 32 < AW; ++k) {
            const int i = (k * NW + warp) * 32 + lane;
            bool q = false;
            if (i < n_slots && (vd[i] & DM_ACTIVE)) {
              const unsigned gc = vg[i];
              if (gc & (FG_TESTED | FG_AVD)) vg[i] = gc & ~(FG_TESTED | FG_AVD);      // transient week flags
              q = (gc & GC_NEW_MASK) || (vh[i] & HV_NEW);
            }
            WQ_PUSH(q0, c0, q, i);
            WQ_DRAIN(q0, c0, (k + 1) * 32 >= AW, item,
              const int i = (int)item; const size_t idx = rbase + i;
              unsigned gc = vg[i]; unsigned h = vh[i]; const unsigned cell = DM_CELL(vd[i]);
              const unsigned gc_old = gc; const unsigned h_old = h;
              unsigned nm = (gc >> GC_NEW_SHIFT) & 0x7Fu;
              if (nm && (mask & XGC_M_STITRANS)) {
                short4 ga = g.gck_a[idx]; short4 gb = g.gck_b[idx]; short4 gd = make_short4(-1, -1, -1, -1);
                if (pois_mode) gd = g.gck_d[idx];
                uint4 x = frand(key, at, FS_INF, (unsigned)i, 0);
                const unsigned xs[3] = { x.x, x.y, x.z };
                bool any = false;
_Pragma("unroll")
                for (int s = 0; s < 3; ++s) {
                  unsigned cand = nm & (s == 0 ? ((1u << XGC_PATH_U2R) | (1u << XGC_PATH_P2R)) : s == 1 ? ((1u << XGC_PATH_R2U) | (1u << XGC_PATH_P2U))
                                                       : ((1u << XGC_PATH_U2P) | (1u << XGC_PATH_R2P) | (1u << XGC_PATH_P2P)));
                  if (!cand || (gc & GC_ST(s))) continue;
                  int win = -1; int wr = 99;
                  for (int kk = 0; kk < XGC_NPATH; ++kk) if ((cand >> kk) & 1u) { int rk = F.sc[SC_PATH + kk]; if (rk < wr) { wr = rk; win = kk; } }
                  gc = (gc | GC_ST(s)) & ~(GC_SY(s) | GC_TX(s));
                  if (bern(xs[s], P[XGC_P_gc_sympt_prob + s])) gc |= GC_SY(s);
                  short du = -1;
                  if (pois_mode) { uint4 y = frand(key, at, FS_INF, (unsigned)i, 1 + s); int qd = pois_f(u01(y.x), P[XGC_P_gc_ntx_int + s], 255); du = (short)(qd < 1 ? 1 : qd); }
                  if (s == 0) { ga.x = (short)at; gb.x = -1; gd.x = du; } else if (s == 1) { ga.y = (short)at; gb.y = -1; gd.y = du; } else { ga.z = (short)at; gb.z = -1; gd.z = du; }
                  atomicAdd(&F.row[(XGC_K_INCID_RGC + s) * XGC_NCELL + cell], 1u);
                  atomicAdd(&F.row[XGC_K_PATH + win], 1u);
                  any = true;
                }
                if (any) { g.gck_a[idx] = ga; g.gck_b[idx] = gb; if (pois_mode) g.gck_d[idx] = gd; }
              }
              if (mask & XGC_M_STITRANS) gc &= ~GC_NEW_MASK;
              if ((h & HV_NEW) && (mask & XGC_M_HIVTRANS) && !(h & HV_STATUS)) {
                h = (h & ~(HV_STAGE_MASK | HV_DIAG | HV_TX)) | HV_STATUS | HV_VSUPP | (1u << 1);
                short4 ha = g.hck_a[idx]; ha.x = (short)at; ha.z = (short)at; g.hck_a[idx] = ha;
                short4 hb = g.hck_b[idx]; hb.x = 0; hb.y = 0; g.hck_b[idx] = hb;
                g.aux[idx].vl = 0.f;
                atomicAdd(&F.row[XGC_K_INCID_HIV * XGC_NCELL + cell], 1u);
              }
              if (mask & XGC_M_HIVTRANS) h &= ~HV_NEW;
              vg[i] = gc; vh[i] = h;
              if (incr) tally_move(F.row, vd[i], gc_old, h_old, vd[i], gc, h);
            );
          }
          __syncwarp();
          // prevalence_msm.  Multi-week launches: the state tallies in F.row are current (tally_move).  A single-week call
          // counts the staged state once, here
          if (!incr && (mask & XGC_M_PREV))
            for (int k = 0; k * 32 < AW; ++k) {
              const int i = (k * NW + warp) * 32 + lane;
              const unsigned d = i < n_slots ? vd[i] : 0u;
              unsigned bits = i < n_slots ? tally_bits(d, vg[i], vh[i]) & ~(1u << XGC_K_NUM) : 0u;
              const bool live = d & DM_ACTIVE; const unsigned cell = live ? DM_CELL(d) : 31u;
              const unsigned grp = __match_any_sync(FULL, cell);
              if (live && lane == (unsigned)(__ffs(grp) - 1)) atomicAdd(&F.row[XGC_K_NUM * XGC_NCELL + cell], (unsigned)__popc(grp));
              while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; atomicAdd(&F.row[b * XGC_NCELL + cell], 1u); }
            }
        }
        __syncthreads();
        TICK(PH_APPLY_TALLY);
      }
      // (a week the call leaves before its POST group hands zeros on for the state tallies: the call that finishes the week counts them)
      const bool partial = incr && !(ph & XF_POST);
      for (int k = tid; k < XGC_NCOUNTER; k += FT)
        grow[k] = (partial && k < XGC_K_NGROUP * XGC_NCELL && ((STATE_GROUPS >> (k / XGC_NCELL)) & 1u)) ? 0u : F.row[k];
      __syncthreads();
      TICK(PH_ROW);
    }
    // a span leaves week at1 half done: its running state tallies go to the carry buffer for the call that continues the week
    if (incr && A.carry && !(A.ph_last & XF_POST) && w_last == A.at1)
      for (int k = tid; k < XGC_K_NGROUP * XGC_NCELL; k += FT) if ((STATE_GROUPS >> (k / XGC_NCELL)) & 1u) A.carry[(size_t)r * XGC_NCOUNTER + k] = F.row[k];
    // ------------------------------------------------------------------------------------------ write the replicate back
    if (F.sc[SC_DEAD0] | F.sc[SC_DEAD1] | F.sc[SC_DEAD2]) squeeze();     // HBM always holds compact tie lists
    TICK(PH_SQUEEZE_WB);
    {
      const int n_slots = F.sc[SC_NSLOTS];
      unsigned* Abm = g.alive + (size_t)r * (g.n_cap / 32);
      for (int i0 = 0; i0 < g.n_cap; i0 += FT) {
        int i = i0 + tid;
        bool in = i < n_slots;
        unsigned d = in ? vd[i] : 0u;
        // (the demo word changes rarely -- age group, departure -- and is written through where it changes; the uid word never changes)
        if (in && !wt) *reinterpret_cast<uint2*>(core_w + 4 * (rbase + i) + 2) = make_uint2(vg[i], vh[i]);
        unsigned m = __ballot_sync(FULL, in && (d & DM_ACTIVE));
        if (lane == 0 && i < g.n_cap) Abm[i >> 5] = m;
      }
      if (tid == 0) {
        rp->n_slots = n_slots; rp->n_slots_pre = F.sc[SC_NPRE]; rp->next_uid = F.sc[SC_NEXTUID]; rp->t_age = F.sc[SC_TAGE];
        rp->ne[0] = F.sc[SC_NE0]; rp->ne[1] = F.sc[SC_NE1]; rp->ne[2] = F.sc[SC_NE2];
        rp->n_alive = F.sc[SC_NALIVE];
        if (F.sc[SC_STATUS]) rp->status = __ldcg(&rp->status) | (unsigned)F.sc[SC_STATUS];
        for (int k = 0; k < XGC_NPATH; ++k) rp->path_rank[k] = F.sc[SC_PATH + k];
      }
      __threadfence();                                          // the replicate's state before the week it reached is published
      __syncthreads();
      if (tid == 0) __stcg((int*)g.at_ptr + r, w_last);
      if (tid == 0) atomicAdd(&g.rep_cycles[r], (unsigned long long)(clock64() - t_rep));
    }
    TICK(PH_WRITEBACK);
  }
}

// quantised birth week and insertivity quotient into the spare bits of the core words
__global__ void __launch_bounds__(256) k_fast_prepare(const __grid_constant__ Dev g) {
  int r = blockIdx.y + g.r0, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= g.rep[r].n_slots) return;
  size_t idx = (size_t)r * g.n_cap + i;
  uint4 c = g.core[idx]; Aux ax = g.aux[idx];
  unsigned y = (c.y & 0xFFFFu) | (xf_insq16(ax.ins_quot) << 16);
  unsigned w = (c.w & ~(XF_BIRTH_MASK << XF_BIRTH_SHIFT)) | (xf_birth_code(ax.age_ref, g.rep[r].t_ref) << XF_BIRTH_SHIFT);
  if (y != c.y || w != c.w) { c.y = y; c.w = w; g.core[idx] = c; }
}

cudaError_t xgc_fast_prepare(const Dev& g, cudaStream_t s) {
  k_fast_prepare<<<dim3((g.n_cap + 255) / 256, g.R), 256, 0, s>>>(g);
  return cudaGetLastError();
}
const void* xgc_fast_kernel_ptr() { return (const void*)k_week_fast; }
bool xgc_fast_hostnet_fits(int n_cap, const int* e_cap) {
  const size_t words = (size_t)(n_cap + 31) / 32, kw = (size_t)(e_cap[0] + 31) / 32 + (size_t)(e_cap[1] + 31) / 32;
  return n_cap <= 32768 && (2 * words + 2 * kw) * sizeof(unsigned) <= sizeof(Fixed::mq);
}
// phase clocks of k_week_fast (SM cycles summed over CTAs) since the last reset; instrumentation for tools/time_modes.py
extern "C" int xgc_debug_fast_profile(unsigned long long* out, int cap, int reset) {
  unsigned long long h[PH_N];
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(h, g_fast_prof, sizeof h) != cudaSuccess) return -1;
  for (int k = 0; k < PH_N && k < cap; ++k) out[k] = h[k];
  if (reset) { unsigned long long z[PH_N] = { 0 }; cudaMemcpyToSymbol(g_fast_prof, z, sizeof z); }
  return PH_N;
}
__global__ void k_fast_set_at(const __grid_constant__ Dev g, int at) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < g.R) ((int*)g.at_ptr)[g.r0 + r] = at;
}
cudaError_t xgc_fast_launch(const Dev& g, const FastArgs& a, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_week_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = g.R < sms ? g.R : sms;
  // a replicate whose weeks are split between two CTAs hands over through at_ptr[r]: make sure it cannot match by accident
  if (a.at1 > a.at0 && g.R % grid) k_fast_set_at<<<(g.R + 255) / 256, 256, 0, s>>>(g, a.at0 - 1);
  k_week_fast<<<grid, FT, smem, s>>>(g, a);
  return cudaGetLastError();
}
