// Hand-written sm_100a kernels of the weekly epidemic step.  No tensor cores: the path is scan / gather / compaction
// work bound by HBM bandwidth.  One week =
//   k_begin   (1 warp / replicate)    week bookkeeping, arrivals (arrival_msm), pathway order
//   k_agent1a (1 thread / slot)       aging, natural mortality, HIV test draw; queue HIV-positives  [reads core+aux+tck]
//   k_agent1b (1 thread / queued)     AIDS mortality, test result, hivtx, hivprogress, hivvl        [reads core+HIV clocks]
//   k_edges_prune / k_rank / k_form   edge deletion for departed agents; surrogate for simnet_msm
//   k_edge1   (1 thread / edge)       acts (+ exposure history, PrEP risk history)                [gathers core+aux]
//   k_agent2a (1 thread / slot)       prep; queue agents with GC work                              [reads core+tck]
//   k_agent2b (1 thread / queued)     stirecov, stitx                                              [reads core+clocks]
//   k_edge2   (1 thread / edge)       condoms + position (lazily per act), hivtrans, stitrans     [gathers core]
//   k_agent3  (1 thread / slot)       apply new infections, prevalence tallies                    [reads core]
// Module semantics follow DESIGN.md §3; reference call sites are cited there and in include/xgc/abi.h.
#pragma once
#include "xgc_device.cuh"

#define TPB 256

__device__ __forceinline__ void warp_tally(unsigned* row, int idx, bool pred) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(row + idx, (unsigned)__popc(m));
}
__device__ __forceinline__ void warp_tally_n(unsigned* row, int idx, unsigned v) {
  v = __reduce_add_sync(0xffffffffu, v);            // one REDUX instead of five shuffle + add steps
  if (v && (threadIdx.x & 31) == 0) atomicAdd(row + idx, v);
}
// The scalar event counters [XGC_K_PATH, XGC_K_PATH + 32) aggregated per CTA in shared memory and flushed with one global
// atomic per touched counter: with ONE large replicate (BASELINE configs[4]) every warp of a kernel otherwise hits the same
// few global addresses (k_agent2b: 50k same-address atomics = 0.06 ms of a 0.075 ms launch at 1M agents).
struct CtaTally {
  unsigned* s;
  __device__ __forceinline__ void init(unsigned* smem) { s = smem; if (threadIdx.x < 32) s[threadIdx.x] = 0u; __syncthreads(); }
  __device__ __forceinline__ void add(int idx, bool pred) {           // reached by every lane of the warp
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&s[idx - XGC_K_PATH], (unsigned)__popc(m));
  }
  __device__ __forceinline__ void add_n(int idx, unsigned v) {
    v = __reduce_add_sync(0xffffffffu, v);
    if (v && (threadIdx.x & 31) == 0) atomicAdd(&s[idx - XGC_K_PATH], v);
  }
  __device__ __forceinline__ void flush(unsigned* row) {              // reached by every thread of the CTA
    __syncthreads();
    if (threadIdx.x < 32) { unsigned v = s[threadIdx.x]; if (v) atomicAdd(row + XGC_K_PATH + threadIdx.x, v); }
  }
};
__device__ __forceinline__ unsigned* counter_row(const Dev& g, int r, int at) {
  return g.counters + ((size_t)r * g.t_cap + (size_t)at) * XGC_NCOUNTER;
}
__device__ __forceinline__ short4 na4() { return make_short4(-1, -1, -1, -1); }

// k_derive: per-replicate tables derived from the parameters (run whenever parameters change)
__global__ void k_derive(const __grid_constant__ Dev g, double* pois, double* inv_ntx, unsigned* pstat /*[R][4] parameter-derived status bits*/) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.R * 4) return;
  int r = t >> 2, w = t & 3;
  pstat[t] = 0u;
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  inv_ntx[t] = w < 3 ? 1.0 / P[XGC_P_gc_ntx_int + w] : 0.0;          // IEEE division, same bits as the oracle's 1.0 / p
  double mu = w == 0 ? P[XGC_P_kiss_rate_main] : w == 1 ? P[XGC_P_kiss_rate_casl] : w == 2 ? P[XGC_P_rim_rate_main] : P[XGC_P_rim_rate_casl];
  double* tab = pois + (size_t)t * (XGC_MAX_ACTS + 1);
  if (!(mu > 0.0)) { for (int k = 0; k <= XGC_MAX_ACTS; ++k) tab[k] = __longlong_as_double(0x7ff0000000000000LL); return; }
  double pr = xgc_exp(-mu), cdf = pr;
  tab[0] = cdf;
  for (int k = 1; k <= XGC_MAX_ACTS; ++k) { pr = xgc_pois_next(pr, mu, k); cdf = cdf + pr; tab[k] = cdf; }
  if (1.0 - cdf > XGC_TRUNC_TOL) pstat[t] = XGC_ST_RATE_TRUNCATED;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_begin: one warp per replicate.  Sets the week, zeroes its tally row, draws the pathway order, appends arrivals.
// arrival_msm: /root/reference/burnin/cal/sim1_02_lhs.r:204, 63-66
// ---------------------------------------------------------------------------------------------------------------------
// WIDE (replicates of > 32k slots: BASELINE configs[3] / [4]): one CTA of 512 threads per replicate; warp 0 does the bookkeeping,
// all 16 warps append the ~500 weekly arrivals of a 1M-agent replicate (one warp alone: 0.027 ms of a 0.28 ms week).
template <bool INJ, bool WIDE> __global__ void __launch_bounds__(WIDE ? 512 : 128) k_begin(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, int at_host, unsigned mask, int zero_row) {
  pdl_prologue();
  int r = WIDE ? (int)blockIdx.x : (int)(blockIdx.x * 4 + (threadIdx.x >> 5)), lane = threadIdx.x & 31;
  if (r >= g.R) return;
  r += g.r0;
  Rep* rp = g.rep + r;
  const int at = at_host >= 0 ? at_host : g.at_ptr[r] + 1;
  unsigned* row = counter_row(g, r, at);
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const int n_slots = rp->n_slots, t_ref = rp->t_ref, uid0 = rp->next_uid;
  const int t_age = (mask & XGC_M_AGING) ? at : rp->t_age;
  __shared__ int s_nnew;
  if (WIDE) __syncthreads();                                // every warp has read the week and the replicate's scalars
  const bool lead = !WIDE || threadIdx.x < 32;              // the warp that does the replicate's bookkeeping
  int nnew = 0;
  if (lead) {
    if (zero_row) for (int k = lane; k < XGC_NCOUNTER; k += 32) row[k] = 0u;
    __syncwarp();
    if (lane < 4) g.q_count[(size_t)r * 4 + lane] = 0u;
    if (lane == 4) g.qa_count[r] = 0u;
    if (lane == 5) g.qh_count[r] = 0u;
    if (lane == 6) g.q3_count[r] = 0u;
    if (lane >= 8 && lane < 11) g.lb_ticket[r * 3 + lane - 8] = 0u;
    if (lane == 0) {
      ((int*)g.at_ptr)[r] = at;
      rp->n_slots_pre = n_slots;
      rp->t_age = t_age;
      if (mask & XGC_M_STITRANS) {          // weekly random pathway order (stitrans_msm_rand, ambiguity A2)
        DrawT<INJ> d(inj, g, r, at, XGC_S_PATHORDER, 0u, 0u, 0);
        int perm[XGC_NPATH];
        for (int k = 0; k < XGC_NPATH; ++k) perm[k] = k;
        for (int i = XGC_NPATH - 1; i >= 1; --i) {
          int j = (int)(d(XGC_NPATH - 1 - i) * (double)(i + 1)); if (j > i) j = i;
          int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        }
        for (int k = 0; k < XGC_NPATH; ++k) rp->path_rank[perm[k]] = k;
      }
    }
    if (mask & XGC_M_ARRIVAL) {
      // nNew ~ Poisson(a.rate * num[1]), mean split into chunks of <= 32 so the CDF walk never underflows
      double mu = P[XGC_P_a_rate] * P[XGC_P_n_init];
      int m = (int)ceil(mu / 32.0); if (m < 1) m = 1; if (m > 64) m = 64;
      double muc = mu / (double)m;
      DrawT<INJ> d(inj, g, r, at, XGC_S_ARRIVE_N, 0u, 0u, 0);
      for (int ch = lane; ch < m; ch += 32) nnew += xgc_pois(d(ch), muc, 255);
#pragma unroll
      for (int o = 16; o; o >>= 1) nnew += __shfl_xor_sync(0xffffffffu, nnew, o);
      if (n_slots + nnew > g.n_cap) { nnew = g.n_cap - n_slots; if (lane == 0) atomicOr(&rp->status, XGC_ST_AGENT_OVERFLOW); }
    }
    if (WIDE && lane == 0) s_nnew = nnew;
  }
  if (!(mask & XGC_M_ARRIVAL)) return;
  if (WIDE) { __syncthreads(); nnew = s_nnew; }
  const int jstep = WIDE ? 512 : 32;
  for (int j = WIDE ? (int)threadIdx.x : lane; j < nnew; j += jstep) {
    int i = n_slots + j; size_t idx = (size_t)r * g.n_cap + i;
    unsigned uid = (unsigned)(uid0 + j);
    DrawT<INJ> d(inj, g, r, at, XGC_S_ARRIVE_ATTR, uid, 0u, uid);
    int race = xgc_categorical(d(0), g.tables + XGC_T_RACE_PROP, 4);
    int role = xgc_categorical(d(1), g.tables + XGC_T_ROLE_PROB + race * 3, 3);
    float insq = role == 0 ? 1.0f : role == 1 ? 0.0f : (float)d(2);
    unsigned circ = xgc_bern(d(3), P[XGC_P_circ_prob + race]);
    unsigned late = xgc_bern(d(4), P[XGC_P_hiv_test_late_prob + race]);
    double tt[3] = { P[XGC_P_tt_part_supp + race], P[XGC_P_tt_full_supp + race], P[XGC_P_tt_dur_supp + race] };
    unsigned traj = 1 + xgc_categorical(d(5), tt, 3);
    unsigned stopper = xgc_bern(d(6), P[XGC_P_act_stopper_prob]);
    int rg = (int)(d(7) * 5.0); if (rg > 4) rg = 4;
    Aux ax; ax.age_ref = P[XGC_P_arrival_age] - (double)(t_age - t_ref) * XGC_WEEK; ax.vl = 0.f; ax.ins_quot = insq;
    double age = ax.age_ref + (double)(t_age - t_ref) * XGC_WEEK;
    unsigned demo = (unsigned)race | (xgc_age_group(age) << 2) | ((unsigned)role << 5) | ((unsigned)rg << 7) |
                    (circ ? DM_CIRC : 0u) | (late ? DM_LATE : 0u) | (traj << 12) | (stopper ? DM_STOPPER : 0u) | DM_ACTIVE;
    g.core[idx] = make_uint4(uid, demo, 0u, HV_VSUPP);
    g.aux[idx] = ax;
    g.gck_a[idx] = na4(); g.gck_b[idx] = na4(); g.gck_d[idx] = na4();
    g.hck_a[idx] = na4(); g.hck_b[idx] = make_short4(0, 0, -1, -1); g.tck[idx] = na4();
    g.arr_t[idx] = at;
    atomicOr(g.alive + (size_t)r * (g.n_cap / 32) + (i >> 5), 1u << (i & 31));
  }
  if (WIDE) __syncthreads(); else __syncwarp();
  if (lead && lane == 0) {
    rp->n_slots = n_slots + nnew; rp->next_uid = uid0 + nnew; atomicAdd(&rp->n_alive, nnew);
    atomicAdd(row + XGC_K_NNEW, (unsigned)nnew);
  }
}

// Per-CTA staging of a work queue: the agent passes walk `tpc` tiles per CTA and used to do one RETURNING global atomicAdd
// per warp and tile to reserve queue slots -- a ~700-cycle round trip in most warps.  Instead the CTA collects its items
// in shared memory (capacity tpc * TPB, so it cannot overflow), reserves its range with ONE global atomicAdd at the end
// and writes the range coalesced.
struct CtaQueue {
  unsigned* items; unsigned* n;
  __device__ __forceinline__ void init(unsigned* smem) { items = smem + 2; n = smem; if (threadIdx.x == 0) smem[0] = 0u; }
  __device__ __forceinline__ void push(bool q, unsigned item, unsigned lane) {        // reached by every lane of the warp
    unsigned m = __ballot_sync(0xffffffffu, q);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(n, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (q) items[base + __popc(m & ((1u << lane) - 1u))] = item;
  }
  __device__ __forceinline__ void flush(unsigned* gcount, unsigned* gitems) {           // reached by every thread of the CTA
    __syncthreads();
    unsigned cnt = *n;
    if (cnt == 0) return;
    if (threadIdx.x == 0) n[1] = atomicAdd(gcount, cnt);
    __syncthreads();
    unsigned gbase = n[1];
    for (unsigned k = threadIdx.x; k < cnt; k += blockDim.x) gitems[gbase + k] = items[k];
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// k_agent1: aging_msm, departure_msm, hivtest_msm, hivtx_msm, hivprogress_msm, hivvl_msm (sim1_02_lhs.r:202-208)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// Agent kernels: each CTA walks `tpc` consecutive 256-slot tiles of one replicate.  The next tile's core word is loaded
// (and its secondary planes prefetched into L2) before the current tile is processed, so every thread keeps two tiles
// of loads in flight; fewer, longer CTAs also mean fewer flushes of the per-CTA tallies.
// k_agent1a: everything every agent goes through: start-of-week snapshot bits, aging, natural mortality, the HIV
// interval-test draw.  HIV-positive agents (a minority, with long divergent code paths: AIDS mortality, test result,
// treatment, progression, viral load) are appended to the replicate's work queue and finished densely by k_agent1b.
template <bool INJ> __global__ void __launch_bounds__(TPB) k_agent1a(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask, int snap, int tpc) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  const Rep* rp = g.rep + r;
  const size_t rbase = (size_t)r * g.n_cap;
  uint4 cn = g.core[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];      // speculative: always inside the plane
  Aux axn = g.aux[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];
  short lntn = g.tck[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x].x;
  int arrn = g.arr_t[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];      // arrival week: the test clock of the never-tested (loaded with the tile, not on demand)
  int n_slots = rp->n_slots;
  int t0 = blockIdx.x * tpc, t1 = min(t0 + tpc, (n_slots + TPB - 1) / TPB);
  if (t0 >= t1) return;
  extern __shared__ unsigned cq_smem[];
  CtaQueue cq; cq.init(cq_smem);
  __shared__ unsigned s_tally[32];
  CtaTally tl; tl.init(s_tally);
  int at = g.at_ptr[r], lane = threadIdx.x & 31;
  int n_pre = rp->n_slots_pre;
  const double dt = (double)(rp->t_age - rp->t_ref) * XGC_WEEK;
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  unsigned* row = counter_row(g, r, at);
  unsigned* Q = g.qh_items + rbase;
  const bool hivmods = mask & (XGC_M_HIVTEST | XGC_M_HIVTX | XGC_M_HIVPROGRESS | XGC_M_HIVVL);
  for (int t = t0; t < t1; ++t) {
    int i = t * TPB + threadIdx.x;
    size_t idx = rbase + i;
    uint4 c = cn; Aux ax = axn; short lnt = lntn; int arr = arrn;
    if (t + 1 < t1) { cn = g.core[idx + TPB]; axn = g.aux[idx + TPB]; lntn = g.tck[idx + TPB].x; arrn = g.arr_t[idx + TPB]; }
    bool live = i < n_slots && (c.y & DM_ACTIVE);
    unsigned uid = c.x, demo = c.y, gc = c.z, hiv = c.w;
    if (snap) {     // start of week: remember the GC / PrEP bits that acts, condoms and hivtrans must see (they run before
                    // prep / stirecov / stitx in the reference's module order)
      gc = (gc & ~(GC_CUR_MASK << GC_PRE_SHIFT)) | ((gc & GC_CUR_MASK) << GC_PRE_SHIFT);
      hiv = (hiv & ~HV_PREP_PRE) | ((hiv & HV_PREP) ? HV_PREP_PRE : 0u);
    }
    bool isnew = i >= n_pre;
    bool hivpos = live && (hiv & HV_STATUS);
    double age = ax.age_ref + dt;
    int race = (int)DM_RACE(demo);
    DrawT<INJ> d(inj, g, r, at, XGC_S_AGENT1, uid, 0u, uid);     // k0 natural mortality and k2 interval test: one Philox block
    if ((mask & XGC_M_AGING) && live && !isnew) demo = (demo & ~(7u << 2)) | (xgc_age_group(age) << 2);
    bool dep_g = false, dep_a = false;
    if ((mask & XGC_M_DEPARTURE) && live && !isnew) {
      int ai = (int)floor(age); if (ai < 0) ai = 0; if (ai > XGC_ASMR_AGES - 1) ai = XGC_ASMR_AGES - 1;
      dep_g = xgc_bern(d(0), g.tables[XGC_T_ASMR + race * XGC_ASMR_AGES + ai]);
      // AIDS-stage mortality: drawn (and tallied) for every stage-4 agent, from the same Philox block.  Settling every
      // departure here means nothing after this kernel touches the alive bitmap: the HIV-positive queue (k_agent1b) and
      // the network kernels are independent and run concurrently.
      if (hivpos && HV_STAGE(hiv) == 4) dep_a = xgc_bern(d(1), P[XGC_P_aids_mr]);
    }
    const bool dead = dep_g || dep_a;
    if (mask & XGC_M_DEPARTURE) {
      unsigned dm = __ballot_sync(0xffffffffu, dead);
      if (dm) {
        unsigned mg = __ballot_sync(0xffffffffu, dep_g), ma = __ballot_sync(0xffffffffu, dep_a);
        if (lane == 0) {
          atomicAnd(g.alive + (size_t)r * (g.n_cap / 32) + (i >> 5), ~dm);
          atomicSub((int*)&rp->n_alive, __popc(dm));
          if (mg) atomicAdd(&s_tally[XGC_K_DEP_GEN - XGC_K_PATH], (unsigned)__popc(mg));
          if (ma) atomicAdd(&s_tally[XGC_K_DEP_AIDS - XGC_K_PATH], (unsigned)__popc(ma));
        }
      }
    }
    if (dead) demo &= ~DM_ACTIVE;
    // hivtest_msm: who gets tested this week (the result of a positive's test is settled in k_agent1b)
    bool tested = false;
    if ((mask & XGC_M_HIVTEST) && live && !dead && !(hiv & HV_DIAG)) {
      int last = lnt == XGC_NA16 ? arr : (int)lnt;
      double tsl = (double)(at - last);
      if (!(hiv & HV_PREP) && !(demo & DM_LATE) && tsl >= 2.0) tested = xgc_bern(d(2), P[XGC_P_hiv_test_rate + race]);
      if (HV_STAGE(hiv) == 4 && tsl >= P[XGC_P_aids_test_int]) tested = true;
      if ((hiv & HV_PREP) && tsl >= P[XGC_P_prep_tst_int]) tested = true;
      if (tested && !(hiv & HV_STATUS)) g.tck[idx].x = (short)at;
    }
    if (mask & XGC_M_HIVTEST) tl.add(XGC_K_HIVTESTS, tested);
    if (i < n_slots && (demo != c.y || gc != c.z || hiv != c.w)) {
      uint2* q = (uint2*)&g.core[idx];
      if (demo != c.y) q[0].y = demo;
      if (gc != c.z || hiv != c.w) q[1] = make_uint2(gc, hiv);
    }
    // queue the surviving HIV-positive agents
    bool queue = hivmods && hivpos && !isnew && !dead;
    cq.push(queue, (unsigned)i | (tested ? 0x80000000u : 0u), lane);
  }
  cq.flush(g.qh_count + r, Q);
  tl.flush(row);
}

// k_agent1b: HIV test result, hivtx_msm, hivprogress_msm, hivvl_msm over the queued (surviving) HIV-positive agents.
// Touches only HIV fields of the queued agents: independent of the network kernels that follow k_agent1a.
template <bool INJ> __global__ void __launch_bounds__(TPB) k_agent1b(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  unsigned n = g.qh_count[r];
  if (blockIdx.x * TPB >= n) return;
  const size_t rbase = (size_t)r * g.n_cap;
  int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  unsigned* row = counter_row(g, r, at);
  // queue kernels run a few CTAs per replicate that stride over the replicate's queue (grid.x is sized by the host for
  // ~2 waves of the GPU, not for the longest possible queue: no tens of thousands of empty CTAs)
  for (unsigned qbase = blockIdx.x * TPB; qbase < n; qbase += gridDim.x * TPB) {
  unsigned qi = qbase + threadIdx.x;
  if (qi < n) {
    unsigned item = g.qh_items[rbase + qi];
    bool tested = item >> 31;
    int i = (int)(item & 0x7FFFFFFFu);
    size_t idx = rbase + i;
    // every gather of this agent is issued at once (they depend only on the slot): one memory round trip, not a chain
    uint4 c = g.core[idx];
    short4 ha = g.hck_a[idx], hb = g.hck_b[idx];
    float vl0 = g.aux[idx].vl;
    unsigned uid = c.x, demo = c.y, hiv = c.w;
    int race = (int)DM_RACE(demo);
    unsigned traj = DM_TTTRAJ(demo);
    DrawT<INJ> d(inj, g, r, at, XGC_S_AGENT1, uid, 0u, uid);
    {
      bool ha_dirty = false, hb_dirty = false;
      if ((mask & XGC_M_HIVTEST) && tested) {
        if ((double)(at - (int)ha.x) >= P[XGC_P_test_window_int]) { hiv |= HV_DIAG; ha.y = (short)at; ha_dirty = true; }
        else g.tck[idx].x = (short)at;
      }
      if (mask & XGC_M_HIVTX) {
        if (hiv & HV_DIAG) {
          if (!(hiv & HV_TX) && hb.x == 0) {
            if (xgc_bern(d(3), P[XGC_P_tx_init_prob + race])) { hiv |= HV_TX; ha.w = (short)at; ha_dirty = true; }
          } else if (hiv & HV_TX) {
            double rr = traj == 2 ? P[XGC_P_tx_halt_full_rr + race] : traj == 3 ? P[XGC_P_tx_halt_dur_rr + race] : 1.0;
            if (xgc_bern(d(3), P[XGC_P_tx_halt_part_prob + race] * rr)) hiv &= ~HV_TX;
          } else {
            double rr = traj == 2 ? P[XGC_P_tx_reinit_full_rr + race] : traj == 3 ? P[XGC_P_tx_reinit_dur_rr + race] : 1.0;
            if (xgc_bern(d(3), P[XGC_P_tx_reinit_part_prob + race] * rr)) hiv |= HV_TX;
          }
        }
        if (hiv & HV_TX) hb.x++; else hb.y++;
        hb_dirty = true;
      }
      double tsi = (double)(at - (int)ha.x);
      if (mask & XGC_M_HIVPROGRESS) {
        unsigned stage = HV_STAGE(hiv);
        if (stage == 1 && tsi >= P[XGC_P_vl_acute_rise_int]) { stage = 2; ha.z = (short)at; ha_dirty = true; }
        if (stage == 2 && tsi >= P[XGC_P_vl_acute_rise_int] + P[XGC_P_vl_acute_fall_int]) { stage = 3; ha.z = (short)at; ha_dirty = true; }
        if (stage == 3) {
          bool aids = false; double on = (double)hb.x, off = (double)hb.y;
          if (on == 0.0 && tsi >= P[XGC_P_vl_aids_onset_int]) aids = true;
          if (traj == 1 && on > 0.0 && off / P[XGC_P_max_time_off_tx_part_int] + on / P[XGC_P_max_time_on_tx_part_int] >= 1.0) aids = true;
          if (traj >= 2 && !(hiv & HV_TX) && on > 0.0 && off >= P[XGC_P_max_time_off_tx_full_int]) aids = true;
          if (aids) { stage = 4; ha.z = (short)at; ha_dirty = true; }
        }
        hiv = (hiv & ~HV_STAGE_MASK) | (stage << 1);
      }
      if (mask & XGC_M_HIVVL) {
        unsigned stage = HV_STAGE(hiv);
        double v, vl = (double)vl0, peak = P[XGC_P_vl_acute_peak], sp = P[XGC_P_vl_set_point];
        if (!(hiv & HV_TX)) {
          if (stage == 1) { v = peak * (tsi / P[XGC_P_vl_acute_rise_int]); if (v > peak) v = peak; }
          else if (stage == 2) { v = peak - (peak - sp) * ((tsi - P[XGC_P_vl_acute_rise_int]) / P[XGC_P_vl_acute_fall_int]); if (v < sp) v = sp; }
          else if (stage == 3) { if (hb.x == 0) v = sp; else { v = vl + P[XGC_P_vl_tx_up_slope]; if (v > sp) v = sp; } }
          else { double tsa = (double)(at - (int)ha.z);
                 v = sp + (tsa / P[XGC_P_vl_aids_int]) * (P[XGC_P_vl_aids_peak] - sp); if (v > P[XGC_P_vl_aids_peak]) v = P[XGC_P_vl_aids_peak]; }
        } else {
          double target = traj == 1 ? P[XGC_P_vl_part_supp] : P[XGC_P_vl_full_supp];
          double slope = stage == 4 ? P[XGC_P_vl_tx_aids_down_slope] : P[XGC_P_vl_tx_down_slope];
          v = vl - slope; if (v < target) v = target;
        }
        float vf = (float)v;
        if (vf != vl0) g.aux[idx].vl = vf;
        hiv = (hiv & ~HV_VSUPP) | ((double)vf <= XGC_LOG10_200 ? HV_VSUPP : 0u);
      }
      if (ha_dirty) g.hck_a[idx] = ha;
      if (hb_dirty) g.hck_b[idx] = hb;
      if (hiv != c.w) ((uint2*)&g.core[idx])[1].y = hiv;
    }
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Edge maintenance.  k_edges_resim: per (replicate, layer) one CTA; stable ballot compaction in place:
//   drops edges with a departed endpoint (tergmLite::delete_vertices semantics, departure_msm) and, in surrogate mode,
//   dissolves edges with prob 1/duration (main, casual) or clears the layer (one-off).
// ---------------------------------------------------------------------------------------------------------------------
// NT threads, IPT edges per thread and iteration: <256,1> for calibration-size layers, <1024,4> for the 100k..1M-agent
// configs where one CTA walks a whole layer (4096 edges per iteration keeps the sequential part short).
template <bool INJ, int NT, int IPT> __global__ void __launch_bounds__(NT) k_edges_resim(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, int dissolve) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Rep* rp = g.rep + r;
  int ne = rp->ne[l], at = g.at_ptr[r];
  int surrogate = dissolve && g.control[(size_t)r * XGC_NCONTROL + XGC_C_network_mode] == 1;
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const uint4* core = g.core + (size_t)r * g.n_cap;
  const unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);          // 1 bit per slot: no 16-byte gathers just to see who is alive
  __shared__ int wsum[IPT][NT / 32];
  __shared__ int base;
  if (tid == 0) base = 0;
  __syncthreads();
  if (surrogate && l == 2) ne = 0;
  double pd = surrogate && l < 2 ? g.inv_dur[l] : 0.0;
  int es = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS + (XGC_S_DISSOLVE - XGC_S_EDGE0);
  for (int e0 = 0; e0 < ne; e0 += NT * IPT) {
    int4 ed[IPT]; bool keep[IPT]; unsigned m[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {                 // item q of this iteration: edges e0 + q*NT .. e0 + (q+1)*NT - 1 (stable order)
      int e = e0 + q * NT + tid; keep[q] = false; ed[q] = make_int4(0, 0, 0, 0);
      if (e < ne) {
        ed[q] = E[e];
        keep[q] = ((A[ed[q].x >> 5] >> (ed[q].x & 31)) & 1u) && ((A[ed[q].y >> 5] >> (ed[q].y & 31)) & 1u);
        if (keep[q] && surrogate) {
          // degrees of the network the week starts with (formation conditions on them): counted before the dissolution draw
          if (l < 2) { unsigned* dg = g.deg + (size_t)r * g.n_cap; atomicAdd(dg + ed[q].x, 1u << (8 * l)); atomicAdd(dg + ed[q].y, 1u << (8 * l)); }
          if (!INJ && (((unsigned)ed[q].w >> 16) & 0x7FFFu) == ((unsigned)at & 0x7FFFu)) keep[q] = !((unsigned)ed[q].w >> 31);   // drawn ahead by k_edge1
          else { DrawT<INJ> d(inj, g, r, at, es, core[ed[q].x].x, core[ed[q].y].x, (size_t)e); keep[q] = !xgc_bern(d(0), pd); }
        }
      }
      m[q] = __ballot_sync(0xffffffffu, keep[q]);
      if (lane == 0) wsum[q][w] = __popc(m[q]);
    }
    __syncthreads();
    int run = base;
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
      int off = run, tot = 0;
      for (int k = 0; k < NT / 32; ++k) { int c = wsum[q][k]; tot += c; if (k < w) off += c; }
      off += __popc(m[q] & ((1u << lane) - 1u));
      if (keep[q] && off != e0 + q * NT + tid) E[off] = ed[q];
      run += tot;
    }
    __syncthreads();
    if (tid == 0) base = run;
    __syncthreads();
  }
  if (tid == 0) rp->ne[l] = base;
}

// k_edges_resim_lb: the same stable in-place compaction for very long layers (100k..1M-agent replicates), many CTAs per
// layer.  Single pass with decoupled look-back: every CTA takes a tile ticket, loads its 1024 records, publishes its
// keep-count, sums the counts of the tiles before it (spinning only on tiles that have not published yet) and writes its
// survivors.  In place is safe: a tile publishes only after it has loaded its records, and tile t only writes below
// (t+1)*1024, i.e. into source ranges whose owners have already loaded.  Status word = epoch << 34 | state << 32 | count
// (state 1 = tile count, 2 = inclusive prefix); the epoch is a per-layer launch counter kept on the device (bumped by the
// last tile), so status words of earlier launches never match.
#define LB_TILE 1024
template <bool INJ> __global__ void __launch_bounds__(TPB) k_edges_resim_lb(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, int dissolve, int kind) {
  pdl_prologue();
  int r = blockIdx.z + g.r0, l = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Rep* rp = g.rep + r;
  int at = g.at_ptr[r];
  int ne = *(volatile int*)&rp->ne[l];               // read BEFORE taking the ticket (the last tile overwrites both)
  unsigned epoch = *(volatile unsigned*)&rp->lb_epoch[l];
  __shared__ int s_tile, s_base;
  __shared__ int wsum[4][TPB / 32];
  if (tid == 0) s_tile = (int)(atomicAdd(g.lb_ticket + r * 3 + l, 1u) % (unsigned)g.lb_tiles);
  __syncthreads();
  int tile = s_tile;
  int surrogate = dissolve && g.control[(size_t)r * XGC_NCONTROL + XGC_C_network_mode] == 1;
  if (surrogate && l == 2) { if (tile == 0 && tid == 0) rp->ne[l] = 0; return; }
  int ntiles = (ne + LB_TILE - 1) / LB_TILE;
  if (tile >= ntiles) return;
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const uint4* core = g.core + (size_t)r * g.n_cap;
  const unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);
  double pd = surrogate && l < 2 ? g.inv_dur[l] : 0.0;
  int es = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS + (XGC_S_DISSOLVE - XGC_S_EDGE0);
  int e0 = tile * LB_TILE;
  int4 ed[4]; bool keep[4]; unsigned m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int e = e0 + q * TPB + tid; keep[q] = false; ed[q] = make_int4(0, 0, 0, 0);
    if (e < ne) {
      ed[q] = E[e];
      keep[q] = ((A[ed[q].x >> 5] >> (ed[q].x & 31)) & 1u) && ((A[ed[q].y >> 5] >> (ed[q].y & 31)) & 1u);
      if (keep[q] && surrogate) {
        if (l < 2) { unsigned* dg = g.deg + (size_t)r * g.n_cap; atomicAdd(dg + ed[q].x, 1u << (8 * l)); atomicAdd(dg + ed[q].y, 1u << (8 * l)); }
        if (!INJ && (((unsigned)ed[q].w >> 16) & 0x7FFFu) == ((unsigned)at & 0x7FFFu)) keep[q] = !((unsigned)ed[q].w >> 31);
        else { DrawT<INJ> d(inj, g, r, at, es, core[ed[q].x].x, core[ed[q].y].x, (size_t)e); keep[q] = !xgc_bern(d(0), pd); }
      }
    }
    m[q] = __ballot_sync(0xffffffffu, keep[q]);
    if (lane == 0) wsum[q][w] = __popc(m[q]);
  }
  __syncthreads();
  int off[4], run = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int o = run, tot = 0;
    for (int k = 0; k < TPB / 32; ++k) { int c = wsum[q][k]; tot += c; if (k < w) o += c; }
    off[q] = o + __popc(m[q] & ((1u << lane) - 1u));
    run += tot;
  }
  if (tid == 0) {
    unsigned long long* S = g.lb_status + ((size_t)r * 3 + l) * g.lb_tiles;
    unsigned long long ep = (unsigned long long)(epoch & 0x3FFFFFFFu) << 34;
    int base = 0;
    if (tile > 0) {
      atomicExch(S + tile, ep | (1ull << 32) | (unsigned)run);
      for (int p = tile - 1; p >= 0; --p) {
        unsigned long long v;
        do { v = *(volatile unsigned long long*)(S + p); } while ((v >> 34) != (ep >> 34) || ((v >> 32) & 3ull) == 0);
        base += (int)(unsigned)v;
        if (((v >> 32) & 3ull) == 2) break;
      }
    }
    atomicExch(S + tile, ep | (2ull << 32) | (unsigned)(base + run));
    s_base = base;
    if (tile == ntiles - 1) { rp->ne[l] = base + run; rp->lb_epoch[l] = epoch + 1u; }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int dst = s_base + off[q];
    if (keep[q] && dst != e0 + q * TPB + tid) E[dst] = ed[q];
  }
}

// k_rank: pos2slot[r][rank] = slot, from the alive bitmap.  grid (tiles of RANK_TILE = 8192 slots, R): each CTA sums the bitmap
// words before its tile (a block reduction over 128-bit loads), scans the popcounts of its own 256 words and every thread
// writes the slots of its word.  (With 256-slot tiles every one of the 3907 CTAs of a 1M-agent replicate re-summed the
// bitmap before it: 0.025 ms per replicate-week; 8192-slot tiles re-read 32 x less.)
#define RANK_TILE (32 * TPB)
__global__ void __launch_bounds__(TPB) k_rank(const __grid_constant__ Dev g) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int n_slots = g.rep[r].n_slots;
  int i0 = blockIdx.x * RANK_TILE;
  if (i0 >= n_slots) return;
  const int nwords = g.n_cap / 32;
  const unsigned* A = g.alive + (size_t)r * nwords;
  int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  __shared__ int wsum[TPB / 32], wpre[TPB / 32];
  const int w0 = i0 >> 5;                          // a multiple of TPB: the earlier words are summed with 128-bit loads
  int acc = 0;
  const uint4* A4 = (const uint4*)A;
  for (int k = tid; k < (w0 >> 2); k += TPB) { uint4 v = A4[k]; acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
  unsigned bits = w0 + tid < nwords ? A[w0 + tid] : 0u;
  const int c = __popc(bits);
  int inc = c;                                     // inclusive scan of the tile's popcounts inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 31) wpre[w] = inc;
  if (lane == 0) wsum[w] = acc;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int k = 0; k < TPB / 32; ++k) base += wsum[k];
  // the tile's slots in rank order go through shared memory so that the plane is written with coalesced stores
  __shared__ int out[RANK_TILE];
  int o = inc - c;
  for (int k = 0; k < w; ++k) o += wpre[k];
  const int slot0 = (w0 + tid) << 5;
  while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; out[o++] = slot0 + b; }
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int k = 0; k < TPB / 32; ++k) total += wpre[k];
  for (int k = tid; k < total; k += TPB) P2S[base + k] = out[k];
}

// k_rank for calibration-size replicates (n_cap <= 32k slots): ONE CTA per replicate builds the popcount prefix of the
// alive bitmap in shared memory and then walks the tiles, instead of 40 CTAs per replicate each re-summing the prefix.
__global__ void __launch_bounds__(TPB) k_rank_cta(const __grid_constant__ Dev g) {
  pdl_prologue();
  extern __shared__ unsigned rk_smem[];            // [words] bitmap, [words + 1] exclusive popcount prefix
  const int r = blockIdx.x + g.r0, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n_slots = g.rep[r].n_slots, words = (n_slots + 31) >> 5;
  const unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);
  int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  unsigned* bits = rk_smem; unsigned* pre = rk_smem + g.n_cap / 32;
  __shared__ unsigned wtot[TPB / 32];
  __shared__ unsigned carry;
  if (tid == 0) carry = 0u;
  __syncthreads();
  for (int w0 = 0; w0 < words; w0 += TPB) {        // block-wide exclusive scan of the per-word popcounts, TPB words per round
    int k = w0 + tid;
    unsigned b = k < words ? A[k] : 0u, c = (unsigned)__popc(b), inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    unsigned off = carry;
    for (int q = 0; q < w; ++q) off += wtot[q];
    if (k < words) { bits[k] = b; pre[k] = off + inc - c; }
    __syncthreads();
    if (tid == TPB - 1) carry = off + inc;
    __syncthreads();
  }
  for (int i = tid; i < n_slots; i += TPB) {
    unsigned b = bits[i >> 5];
    if ((b >> (i & 31)) & 1u) P2S[pre[i >> 5] + __popc(b & ((1u << (i & 31)) - 1u))] = i;
  }
}

// k_zero_deg: the degree plane of this launch's replicates back to zero before the dissolution pass counts into it.  A kernel,
// not a cudaMemsetAsync: on the step-by-step path every kernel is launched with programmatic stream serialization, and only a
// kernel's griddepcontrol.wait orders it against BOTH neighbours (a memset between two such launches is not a grid the next
// kernel's wait covers -- the counts of its first CTAs could be zeroed under it).
__global__ void __launch_bounds__(TPB) k_zero_deg(const __grid_constant__ Dev g) {
  pdl_prologue();
  uint4* p = reinterpret_cast<uint4*>(g.deg + (size_t)g.r0 * g.n_cap);
  const size_t n = (size_t)g.R * g.n_cap / 4;
  for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (size_t)gridDim.x * TPB) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
// formation weight of an agent in layer l (abi.h XGC_T_NET_DEGW / XGC_T_NET_RISKW): degree classes from the byte counts of g.deg
__device__ __forceinline__ double form_weight(const double* tb, int l, unsigned dg, unsigned demo) {
  int cm = min((int)(dg & 0xFFu), 2), cc = min((int)((dg >> 8) & 0xFFu), 2);
  double w = tb[XGC_T_NET_DEGW + l * 9 + cm * 3 + cc];
  if (l == 2) w = w * tb[XGC_T_NET_RISKW + DM_RISK(demo)];
  return w;
}
// k_form: surrogate formation, one CTA per (replicate, layer): uniform random compatible pairs, appended in draw order.
// R position -> slot: SMEM = true keeps the replicate's alive bitmap and its popcount prefix in shared memory and selects
// directly (no pos2slot plane, no k_rank launch); SMEM = false (very large n_cap) reads pos2slot built by k_rank.
template <bool INJ, bool SMEM, int NT> __global__ void __launch_bounds__(NT) k_form(const __grid_constant__ Dev g, const __grid_constant__ Inj inj) {
  pdl_prologue();
  int r = blockIdx.x + g.r0, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;        // one CTA per replicate, the three layers in turn
  if (g.control[(size_t)r * XGC_NCONTROL + XGC_C_network_mode] != 1) return;
  Rep* rp = g.rep + r;
  int at = g.at_ptr[r], n = rp->n_alive;
  const uint4* core = g.core + (size_t)r * g.n_cap;
  const int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  extern __shared__ unsigned dyn[];
  unsigned* bm = dyn; int* pre = (int*)(dyn + g.n_cap / 32);
  __shared__ int wsum[NT / 32];
  __shared__ int base;
  int nw = (rp->n_slots + 31) >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  if (SMEM) {                                   // alive bitmap + exclusive popcount prefix of this replicate, built once
    const unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);
    for (int w0 = 0; w0 < nw; w0 += NT) {
      unsigned bits = w0 + tid < nw ? A[w0 + tid] : 0u;
      int cnt = __popc(bits), inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) wsum[w] = inc;
      __syncthreads();
      int off = base + inc - cnt;
      for (int k = 0; k < w; ++k) off += wsum[k];
      if (w0 + tid < nw) { bm[w0 + tid] = bits; pre[w0 + tid] = off; }
      __syncthreads();
      if (tid == 0) { int t = 0; for (int k = 0; k < NT / 32; ++k) t += wsum[k]; base += t; }
      __syncthreads();
    }
  }
  // position -> slot table of the replicate in shared memory (16-bit slots: n_cap <= 32k here), one load per lookup -- the
  // binary search over the popcount prefix + find-n-th-set-bit it replaces was half the instructions of a formation try
  unsigned short* p2s = (unsigned short*)(dyn + 2 * (g.n_cap / 32));
  if (SMEM) {
    for (int k = tid; k < nw; k += NT) { unsigned bits = bm[k]; int o = pre[k]; while (bits) { int b = __ffs(bits) - 1; bits &= bits - 1; p2s[o++] = (unsigned short)((k << 5) + b); } }
    __syncthreads();
  }
  auto slot_of = [&](int pos) -> int { return SMEM ? (int)p2s[pos] : P2S[pos]; };
  // The candidates of the three layers are ONE index space (main, casual, one-off back to back: ~12 + 34 + 400 a week at 10k agents),
  // walked NT at a time, two tries of a candidate in flight: one pass instead of a pass per layer and chunk, half the dependent
  // rounds (degree-conditioned acceptance needs several tries per candidate).  Ties are appended per layer in candidate order.
  DrawT<INJ> dn(inj, g, r, at, XGC_S_NET_N, 0u, 0u, 0);
  int nf[3];
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    double target = g.tables[XGC_T_NET_DEG + l] * (double)n / 2.0;
    double mu = l == 2 ? target : target / g.tables[XGC_T_NET_DUR + l];
    double fl = floor(mu);
    nf[l] = (int)fl + (dn(l) < mu - fl);
    if (inj.u && nf[l] > inj.form_cap) nf[l] = inj.form_cap;
    if (n < 2) nf[l] = 0;
  }
  const int nf01 = nf[0] + nf[1], nf_all = nf01 + nf[2];
  __shared__ int wsum3[3][NT / 32];
  __shared__ int base3[3];
  __syncthreads();
  if (tid < 3) base3[tid] = rp->ne[tid];
  __syncthreads();
  const unsigned* dg = g.deg + (size_t)r * g.n_cap;
  bool overflow = false;
  for (int j0 = 0; j0 < nf_all; j0 += NT) {
    const int jg = j0 + tid;
    const int l = jg >= nf01 ? 2 : jg >= nf[0] ? 1 : 0, j = jg - (l == 2 ? nf01 : l == 1 ? nf[0] : 0);
    bool ok = false; int4 ed = make_int4(0, 0, at, 0);
    if (jg < nf_all) {
      DrawT<INJ> d(inj, g, r, at, XGC_S_FORM + l, (unsigned)j, 0u, (size_t)j);
      for (int t = 0; t < XGC_FORM_TRIES && !ok; t += 2) {
        int sa[2], sb[2]; bool swp[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          int a = (int)(d(3 * (t + k)) * (double)n); if (a > n - 1) a = n - 1;
          int b = (int)(d(3 * (t + k) + 1) * (double)(n - 1)); if (b > n - 2) b = n - 2;
          if (b >= a) b++;
          swp[k] = !(a < b);                                   // slots are stored in position order; the weights follow the DRAWN order
          sa[k] = slot_of(a < b ? a : b); sb[k] = slot_of(a < b ? b : a);
        }
        unsigned da[2], db[2], ga[2], gb[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) { da[k] = core[sa[k]].y; db[k] = core[sb[k]].y; ga[k] = dg[sa[k]]; gb[k] = dg[sb[k]]; }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (ok || t + k >= XGC_FORM_TRIES || !roles_compatible(da[k], db[k])) continue;
          double wa = form_weight(g.tables, l, swp[k] ? gb[k] : ga[k], swp[k] ? db[k] : da[k]), wb = form_weight(g.tables, l, swp[k] ? ga[k] : gb[k], swp[k] ? da[k] : db[k]);
          double pw = wa * wb;
          if (pw < 1.0 && !(d(3 * (t + k) + 2) < pw)) continue;
          ok = true; ed.x = sa[k]; ed.y = sb[k];
        }
      }
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, ok && l == 0), m1 = __ballot_sync(0xffffffffu, ok && l == 1), m2 = __ballot_sync(0xffffffffu, ok && l == 2);
    if (lane == 0) { wsum3[0][w] = __popc(m0); wsum3[1][w] = __popc(m1); wsum3[2][w] = __popc(m2); }
    __syncthreads();
    if (ok) {
      int off = base3[l];
      for (int k = 0; k < w; ++k) off += wsum3[l][k];
      off += __popc((l == 0 ? m0 : l == 1 ? m1 : m2) & ((1u << lane) - 1u));
      if (off < g.e_cap[l]) (g.edge + (size_t)r * g.e_stride + g.e_off[l])[off] = ed; else overflow = true;
    }
    __syncthreads();
    if (tid < 3) { int t = 0; for (int k = 0; k < NT / 32; ++k) t += wsum3[tid][k]; base3[tid] += t; }
    __syncthreads();
  }
  if (overflow) atomicOr(&rp->status, XGC_ST_EDGE_OVERFLOW);
  if (tid < 3) {
    int nn = base3[tid] < g.e_cap[tid] ? base3[tid] : g.e_cap[tid];
    rp->ne[tid] = nn;
    counter_row(g, r, at)[XGC_K_NEDGES + tid] = (unsigned)nn;
  }
}

// k_form_lb: the same formation for large replicates, many CTAs per (replicate, layer): each CTA draws 1024 candidate
// ties, the accepted ones are appended in draw order through the decoupled look-back prefix (shares the tile tickets /
// status words / epoch with k_edges_resim_lb; launches are serialised on the stream).  Needs pos2slot (k_rank).
#define FORM_NT 1024
template <bool INJ> __global__ void __launch_bounds__(FORM_NT) k_form_lb(const __grid_constant__ Dev g, const __grid_constant__ Inj inj) {
  pdl_prologue();
  int r = blockIdx.z + g.r0, l = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Rep* rp = g.rep + r;
  int ne0 = *(volatile int*)&rp->ne[l];
  unsigned epoch = *(volatile unsigned*)&rp->lb_epoch[l];
  __shared__ int s_tile, s_base;
  __shared__ int wsum[FORM_NT / 32];
  if (tid == 0) s_tile = (int)(atomicAdd(g.lb_ticket + r * 3 + l, 1u) % (unsigned)g.lb_tiles);
  __syncthreads();
  int tile = s_tile;
  if (g.control[(size_t)r * XGC_NCONTROL + XGC_C_network_mode] != 1) return;
  int at = g.at_ptr[r], n = rp->n_alive;
  double target = g.tables[XGC_T_NET_DEG + l] * (double)n / 2.0;
  double mu = l == 2 ? target : target / g.tables[XGC_T_NET_DUR + l];
  double fl = floor(mu);
  DrawT<INJ> dn(inj, g, r, at, XGC_S_NET_N, 0u, 0u, 0);
  int nform = (int)fl + (dn(l) < mu - fl);
  if (inj.u && nform > inj.form_cap) nform = inj.form_cap;
  if (n < 2) nform = 0;
  int ntiles = (nform + LB_TILE - 1) / LB_TILE;
  if (ntiles == 0) { if (tile == 0 && tid == 0) counter_row(g, r, at)[XGC_K_NEDGES + l] = (unsigned)ne0; return; }
  if (tile >= ntiles) return;
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const uint4* core = g.core + (size_t)r * g.n_cap;
  const int* P2S = g.pos2slot + (size_t)r * g.n_cap;
  const unsigned* dg = g.deg + (size_t)r * g.n_cap;
  // One candidate per thread (1024-thread CTAs: with four candidates per thread of a 256-thread CTA the kernel was a chain of
  // dependent global loads per thread -- position -> slot, then core / degree words, planes far larger than L2 at 1M agents --
  // on a quarter of the threads the GPU can hold), and TWO tries of the candidate in flight at a time.
  const int j = tile * LB_TILE + tid;
  bool ok = false; int4 ed = make_int4(0, 0, at, 0);
  if (j < nform) {
    DrawT<INJ> d(inj, g, r, at, XGC_S_FORM + l, (unsigned)j, 0u, (size_t)j);
    for (int t = 0; t < XGC_FORM_TRIES && !ok; t += 2) {
      int sa[2], sb[2]; bool swp[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        int a = (int)(d(3 * (t + k)) * (double)n); if (a > n - 1) a = n - 1;
        int b = (int)(d(3 * (t + k) + 1) * (double)(n - 1)); if (b > n - 2) b = n - 2;
        if (b >= a) b++;
        swp[k] = !(a < b);                                       // slots are stored in position order; the weights follow the DRAWN order
        sa[k] = P2S[a < b ? a : b]; sb[k] = P2S[a < b ? b : a];
      }
      unsigned da[2], db[2], ga[2], gb[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) { da[k] = core[sa[k]].y; db[k] = core[sb[k]].y; ga[k] = dg[sa[k]]; gb[k] = dg[sb[k]]; }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (ok || t + k >= XGC_FORM_TRIES || !roles_compatible(da[k], db[k])) continue;
        double wa = form_weight(g.tables, l, swp[k] ? gb[k] : ga[k], swp[k] ? db[k] : da[k]), wb = form_weight(g.tables, l, swp[k] ? ga[k] : gb[k], swp[k] ? da[k] : db[k]);
        double pw = wa * wb;
        if (pw < 1.0 && !(d(3 * (t + k) + 2) < pw)) continue;
        ok = true; ed.x = sa[k]; ed.y = sb[k];
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) wsum[w] = __popc(m);
  __syncthreads();
  int off = 0, run = 0;
  for (int k = 0; k < FORM_NT / 32; ++k) { int c = wsum[k]; run += c; if (k < w) off += c; }
  off += __popc(m & ((1u << lane) - 1u));
  if (tid == 0) {
    unsigned long long* S = g.lb_status + ((size_t)r * 3 + l) * g.lb_tiles;
    unsigned long long ep = (unsigned long long)(epoch & 0x3FFFFFFFu) << 34;
    int base = 0;
    if (tile > 0) {
      atomicExch(S + tile, ep | (1ull << 32) | (unsigned)run);
      for (int p = tile - 1; p >= 0; --p) {
        unsigned long long v;
        do { v = *(volatile unsigned long long*)(S + p); } while ((v >> 34) != (ep >> 34) || ((v >> 32) & 3ull) == 0);
        base += (int)(unsigned)v;
        if (((v >> 32) & 3ull) == 2) break;
      }
    }
    atomicExch(S + tile, ep | (2ull << 32) | (unsigned)(base + run));
    s_base = ne0 + base;
    if (tile == ntiles - 1) {
      int total = ne0 + base + run;
      if (total > g.e_cap[l]) { atomicOr(&rp->status, XGC_ST_EDGE_OVERFLOW); total = g.e_cap[l]; }
      rp->ne[l] = total; rp->lb_epoch[l] = epoch + 1u;
      counter_row(g, r, at)[XGC_K_NEDGES + l] = (unsigned)total;
    }
  }
  __syncthreads();
  const int dst = s_base + off;
  if (ok && dst < g.e_cap[l]) E[dst] = ed;
}

// ---------------------------------------------------------------------------------------------------------------------
// per-edge helpers shared by k_edge1 / k_edge2 / k_actlist (condoms_msm, position_msm evaluated lazily per act)
// ---------------------------------------------------------------------------------------------------------------------
struct EdgeCtx {
  int r, l, e, at; int4 ed; uint4 ca, cb; double age_a, age_b; float insq_a, insq_b; int base;
};
__device__ __forceinline__ double edge_cond_prob(const Dev& g, const double* P, const EdgeCtx& x, bool use_pre) {
  int hcp = (x.ca.w & HV_DIAG) && (x.cb.w & HV_DIAG);
  unsigned pb = use_pre ? HV_PREP_PRE : HV_PREP;
  int prep = (x.ca.w & pb) || (x.cb.w & pb);
  double ca = x.age_a + x.age_b;
  int r1 = (int)DM_RACE(x.ca.y), r2 = (int)DM_RACE(x.cb.y);
  double eta = x.l < 2 ? xgc_glm_eta(g.tables + XGC_T_GLM_CONDMC, (double)(x.at - x.ed.z), x.l == 1, ca, hcp, prep, r1, r2)
                       : xgc_glm_eta(g.tables + XGC_T_GLM_CONDOO, 0.0, 0, ca, hcp, prep, r1, r2);
  double pc = 1.0 / (1.0 + xgc_exp(-eta));
  pc = pc * P[XGC_P_cond_scale];
  if (pc > 1.0) pc = 1.0;
  return pc;
}
// position of anal act k: 1 = p1 insertive
template <class D> __device__ __forceinline__ int edge_ai_p1ins(const EdgeCtx& x, D& dpos, int k) {
  unsigned r1 = DM_ROLE(x.ca.y), r2 = DM_ROLE(x.cb.y);
  if (r1 == 0 || r2 == 1) return 1;
  if (r1 == 1 || r2 == 0) return 0;
  double q1 = (double)x.insq_a, q2 = (double)x.insq_b;
  double pi = q1 + q2 > 0.0 ? q1 / (q1 + q2) : 0.5;
  return xgc_bern(dpos(4 * k + 1), pi);
}

// kissing / rimming counts of one edge-week: constant-rate Poisson (main, casual) or Bernoulli (one-off), gated by the
// act-stopper rule on the START-of-week symptom / treatment bits (acts_msm runs before stirecov / stitx).
// tab = g.pois + (r*4 + layer)*33: kissing CDF at tab, rimming CDF at tab + 2*33.
__device__ __forceinline__ int pois_count(double u, const double* tab) {
  int k = 0;
#pragma unroll 1
  for (int j = 0; j < XGC_MAX_ACTS; j += 8) {
    int c = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) c += u > tab[j + q];
    k += c;
    if (c < 8) break;
  }
  return k;
}
template <class D> __device__ __forceinline__ void edge_kiss_rim(const double* P, int l, const uint4& ca, const uint4& cb, D& dacts,
                                                                 const double* tab, bool want_k, bool want_r, int& ki, int& ri) {
  ki = ri = 0;
  if (act_stopped(ca.y, ca.z >> GC_PRE_SHIFT) || act_stopped(cb.y, cb.z >> GC_PRE_SHIFT)) return;
  if (l < 2) {
    if (want_k) ki = pois_count(dacts(2), tab);
    if (want_r) ri = pois_count(dacts(3), tab + 2 * (XGC_MAX_ACTS + 1));
  } else {
    if (want_k) ki = xgc_bern(dacts(2), P[XGC_P_kiss_prob_oo]);
    if (want_r) ri = xgc_bern(dacts(3), P[XGC_P_rim_prob_oo]);
  }
}

// k_edge1: acts_msm (sim1_02_lhs.r:210) + exposure history + PrEP risk history.  grid (tiles, 3 layers, R)
template <bool INJ> __global__ void __launch_bounds__(TPB, 4) k_edge1(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  const Rep* rp = g.rep + r;
  // the three layers of the replicate are walked as one index space [0, ne0 + ne1 + ne2): full tiles except the last,
  // a few CTAs per replicate striding over it
  const int ne0 = rp->ne[0], ne01 = ne0 + rp->ne[1], ne_all = ne01 + rp->ne[2];
  if ((int)(blockIdx.x * TPB) >= ne_all) return;
  const int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  unsigned* row = counter_row(g, r, at);
  uint4* core = g.core + (size_t)r * g.n_cap;
  const Aux* aux = g.aux + (size_t)r * g.n_cap;
  __shared__ unsigned s_tally[32];
  CtaTally tl; tl.init(s_tally);
  for (int fbase = blockIdx.x * TPB; fbase < ne_all; fbase += gridDim.x * TPB) {
  const int f = fbase + threadIdx.x;
  const bool live = f < ne_all;
  const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0;
  const int e = f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  unsigned nai = 0, noi = 0;
  if (live) {
    EdgeCtx x; x.r = r; x.l = l; x.e = e; x.at = at; x.ed = E[e];
    x.ca = core[x.ed.x]; x.cb = core[x.ed.y];
    x.base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
    double dt = (double)(rp->t_age - rp->t_ref) * XGC_WEEK;
    Aux aa = aux[x.ed.x], ab = aux[x.ed.y];
    x.age_a = aa.age_ref + dt; x.age_b = ab.age_ref + dt; x.insq_a = aa.ins_quot; x.insq_b = ab.ins_quot;
    if (mask & XGC_M_ACTS) {
      int ai = 0, oi = 0, ki = 0, ri = 0;
      if (!(act_stopped(x.ca.y, x.ca.z) || act_stopped(x.cb.y, x.cb.z))) {
        bool compat = roles_compatible(x.ca.y, x.cb.y);
        DrawT<INJ> d(inj, g, r, at, x.base + (XGC_S_ACTS - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
        if (l < 2) {
          double dur = (double)(at - x.ed.z), ca = x.age_a + x.age_b;
          int hcp = (x.ca.w & HV_DIAG) && (x.cb.w & HV_DIAG);
          int r1 = (int)DM_RACE(x.ca.y), r2 = (int)DM_RACE(x.cb.y);
          double rai = xgc_exp(xgc_glm_eta(g.tables + XGC_T_GLM_AI, dur, l == 1, ca, hcp, 0, r1, r2)) / 52.0 * P[XGC_P_ai_acts_scale_mc];
          double roi = xgc_exp(xgc_glm_eta(g.tables + XGC_T_GLM_OI, dur, l == 1, ca, hcp, 0, r1, r2)) / 52.0 * P[XGC_P_oi_acts_scale_mc];
          ai = compat ? xgc_pois(d(0), rai, XGC_MAX_ACTS) : 0;
          oi = xgc_pois(d(1), roi, XGC_MAX_ACTS);
          // kissing / rimming: only "any act this week" is needed here (exposure history); the counts are evaluated
          // lazily by k_edge2 for the edges where they can transmit (same indexed draws, same table)
          const double* pt = g.pois + ((size_t)r * 4 + l) * (XGC_MAX_ACTS + 1);
          ki = d(2) > pt[0];
          ri = d(3) > pt[2 * (XGC_MAX_ACTS + 1)];
        } else {
          ai = compat ? 1 : 0;
          oi = xgc_bern(d(1), P[XGC_P_oi_prob_oo]);
          ki = xgc_bern(d(2), P[XGC_P_kiss_prob_oo]);
          ri = xgc_bern(d(3), P[XGC_P_rim_prob_oo]);
        }
      }
      // surrogate network: next week's dissolution draw of this tie is a pure function of (uids, at + 1); drawing it here,
      // where the endpoint uids are already gathered, lets k_edges_resim compact next week without touching the agents.
      unsigned tag = 0x7FFFu, dis = 0u;
      if (!INJ && l < 2 && g.control[(size_t)r * XGC_NCONTROL + XGC_C_network_mode] == 1) {
        DrawT<INJ> dd(inj, g, r, at + 1, x.base + (XGC_S_DISSOLVE - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
        dis = xgc_bern(dd(0), g.inv_dur[l]);
        tag = (unsigned)(at + 1) & 0x7FFFu;
      }
      x.ed.w = (int)((unsigned)ai | ((unsigned)oi << 8) | (tag << 16) | (dis << 31));
      E[e].w = x.ed.w;
      nai = ai; noi = oi;
      // exposure history for the CDC screening protocol (ambiguity A4)
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        unsigned role = DM_ROLE(side ? x.cb.y : x.ca.y), ex = 0;
        if (ai > 0 && role != 0) ex |= 1;
        if ((ai > 0 && role != 1) || oi > 0) ex |= 2;
        if (oi > 0 || ri > 0) ex |= 4;
        if (ki > 0) ex |= 8;
        if (l != 0 && (ai | oi | ki | ri)) ex |= 16;
        unsigned cur = side ? x.cb.z : x.ca.z;
        if ((ex << GC_EXPO_SHIFT) & ~cur) atomicOr(&core[side ? x.ed.y : x.ed.x].z, ex << GC_EXPO_SHIFT);
      }
    }
    if (mask & XGC_M_PREP) {              // risk history: needs condom use of this week's anal acts
      int ai = x.ed.w & 0xFF;
      bool ua = !(x.ca.w & HV_DIAG), ub = !(x.cb.w & HV_DIAG);
      if (ai > 0 && (ua || ub)) {
        bool uai = false;
        double pc = edge_cond_prob(g, P, x, false);
        DrawT<INJ> dc(inj, g, r, at, x.base + (XGC_S_ANAL - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
        for (int k = 0; k < ai && !uai; ++k) uai = xgc_bern(dc(4 * k), 1.0 - pc);
        short4* tck = g.tck + (size_t)r * g.n_cap;
        if (ua && (uai || (x.cb.w & HV_DIAG))) tck[x.ed.x].y = (short)at;
        if (ub && (uai || (x.ca.w & HV_DIAG))) tck[x.ed.y].y = (short)at;
      }
    }
  }
  if (mask & XGC_M_ACTS) {
    tl.add_n(XGC_K_NACTS, nai); tl.add_n(XGC_K_NACTS + 1, noi);
  }
  }
  tl.flush(row);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_agent2: prep_msm, stirecov_msm, stitx_msm (sim1_02_lhs.r:213, 215-216).
// ---------------------------------------------------------------------------------------------------------------------
// k_agent2a: prep_msm for every agent + classification: agents with GC work this week (an infected site, a clinic visit)
// are appended to the replicate's work queue; k_agent2b runs stirecov_msm / stitx_msm on the queue with all lanes busy.
template <bool INJ> __global__ void __launch_bounds__(TPB) k_agent2a(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask, int tpc) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  const Rep* rp = g.rep + r;
  const size_t rbase = (size_t)r * g.n_cap;
  uint4 cn = g.core[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];
  short4 tkn = g.tck[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];
  int n_slots = rp->n_slots;
  int t0 = blockIdx.x * tpc, t1 = min(t0 + tpc, (n_slots + TPB - 1) / TPB);
  if (t0 >= t1) return;
  extern __shared__ unsigned cq_smem[];
  CtaQueue cq; cq.init(cq_smem);
  __syncthreads();
  int at = g.at_ptr[r], lane = threadIdx.x & 31;
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const int* C = g.control + (size_t)r * XGC_NCONTROL;
  int proto = C[XGC_C_stiScreeningProtocol], kissx = C[XGC_C_cdcExposureSite_Kissing];
  unsigned* Q = g.qa_items + rbase;
  for (int t = t0; t < t1; ++t) {
    int i = t * TPB + threadIdx.x;
    size_t idx = rbase + i;
    uint4 c = cn; short4 tk = tkn;
    if (t + 1 < t1) { cn = g.core[idx + TPB]; tkn = g.tck[idx + TPB]; }
    bool live = i < n_slots && (c.y & DM_ACTIVE);
    unsigned uid = c.x, demo = c.y, gc = c.z, hiv = c.w;
    int race = (int)DM_RACE(demo);
    bool queue = false, avd = false;
    if (live) {
      DrawT<INJ> d(inj, g, r, at, XGC_S_AGENT2, uid, 0u, uid);
      if ((mask & XGC_M_STITX) && (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL))
        avd = xgc_bern(d(0), P[XGC_P_gc_asympt_visit_prob]);           // asymptomatic clinic-visit draw (used unless a symptomatic visit happens)
      if (mask & XGC_M_PREP) {                                           // prep_msm (sim1_02_lhs.r:213)
        bool diag = hiv & HV_DIAG;
        bool ind = tk.y != XGC_NA16 && (double)tk.y >= (double)at - P[XGC_P_prep_risk_int];
        hiv = (hiv & ~HV_PREPELIG) | ((!diag && ind) ? HV_PREPELIG : 0u);
        if (hiv & HV_PREP) {
          bool stop = false;
          if (diag) stop = true;
          else if (xgc_bern(d(4), P[XGC_P_prep_discont_rate + race])) stop = true;
          else if ((int)tk.x == at && (double)(at - (int)tk.z) >= P[XGC_P_prep_reassess_int]) {
            if (!ind) stop = true; else g.tck[idx].z = (short)at;
          }
          if (stop) hiv &= ~(HV_PREP | HV_PREPCLASS_MASK);
        } else if (!diag && (double)at >= P[XGC_P_prep_start] && (int)tk.x == at && ind) {
          if (xgc_bern(d(4), P[XGC_P_prep_start_prob + race])) {
            unsigned cl = 1 + xgc_categorical(d(5), P + XGC_P_prep_adhr_dist, 3);
            hiv = (hiv & ~HV_PREPCLASS_MASK) | HV_PREP | (cl << 7);
            short4* q = g.tck + idx; q->z = (short)at; q->w = (short)at;
          }
        }
        if (hiv != c.w) ((uint2*)&g.core[idx])[1].y = hiv;
      }
      unsigned ex = (gc >> GC_EXPO_SHIFT) & 0x1Fu;
      bool cdc_active = proto == XGC_SCREEN_CDC && ((ex & 7u) || (kissx && (ex & 8u)));
      queue = (mask & (XGC_M_STIRECOV | XGC_M_STITX)) && ((gc & 7u) || ((mask & XGC_M_STITX) && (avd || cdc_active)));
    }
    cq.push(queue, (unsigned)i | (avd ? 0x80000000u : 0u), lane);
  }
  cq.flush(g.qa_count + r, Q);
}

// k_agent2b: stirecov_msm, stitx_msm (sim1_02_lhs.r:215-216) over the queued agents
template <bool INJ> __global__ void __launch_bounds__(TPB) k_agent2b(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  unsigned n = g.qa_count[r];
  if (blockIdx.x * TPB >= n) return;
  const size_t rbase = (size_t)r * g.n_cap;
  int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const int* C = g.control + (size_t)r * XGC_NCONTROL;
  unsigned* row = counter_row(g, r, at);
  __shared__ unsigned s_tally[32];
  CtaTally tl; tl.init(s_tally);
  for (unsigned qbase = blockIdx.x * TPB; qbase < n; qbase += gridDim.x * TPB) {
  unsigned qi = qbase + threadIdx.x;
  unsigned visit = 0, svisit = 0, tst[3] = { 0, 0, 0 }, pos[3] = { 0, 0, 0 }, txs[3] = { 0, 0, 0 };
  if (qi < n) {
    unsigned item = g.qa_items[rbase + qi];
    bool avd = item >> 31;
    size_t idx = rbase + (item & 0x7FFFFFFFu);
    uint4 c = g.core[idx];
    unsigned uid = c.x, gc = c.z, hiv = c.w;
    bool pois = C[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_POIS;
    short4 ga = g.gck_a[idx], gb = g.gck_b[idx], gd = na4();
    if (pois && (gc & 7u)) gd = g.gck_d[idx];
    bool ga_dirty = false, gb_dirty = false, gd_dirty = false, ind_sti = false;
    DrawT<INJ> d(inj, g, r, at, XGC_S_AGENT2, uid, 0u, uid);
    if ((mask & XGC_M_STIRECOV) && (gc & 7u)) {
      d.prime(0);                                                        // recovery draws r,u,p live in block 0
      const int txpr[3] = { XGC_P_rgc_tx_recov_pr, XGC_P_ugc_tx_recov_pr, XGC_P_pgc_tx_recov_pr };
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (!(gc & GC_ST(s))) continue;
        short infT = s == 0 ? ga.x : s == 1 ? ga.y : ga.z, txT = s == 0 ? gb.x : s == 1 ? gb.y : gb.z;
        short dur = s == 0 ? gd.x : s == 1 ? gd.y : gd.z;
        bool recov = false;
        if (gc & GC_TX(s)) {
          int k = at - (int)txT;
          if (k >= 1) recov = xgc_bern(d(1 + s), P[txpr[s] + (k >= 3 ? 2 : k - 1)]);
        } else if ((int)infT < at) {
          if (!pois) recov = xgc_bern(d(1 + s), g.inv_ntx[(size_t)r * 4 + s]);
          else recov = (at - (int)infT) >= (int)dur;
        }
        if (recov) {
          gc &= ~(GC_ST(s) | GC_SY(s) | GC_TX(s));
          if (s == 0) { ga.x = -1; gb.x = -1; gd.x = -1; } else if (s == 1) { ga.y = -1; gb.y = -1; gd.y = -1; } else { ga.z = -1; gb.z = -1; gd.z = -1; }
          ga_dirty = gb_dirty = true;
          if (pois) gd_dirty = true; else (&g.gck_d[idx].x)[s] = -1;
        }
      }
    }
    if (mask & XGC_M_STITX) {
      int proto = C[XGC_C_stiScreeningProtocol], kissx = C[XGC_C_cdcExposureSite_Kissing];
      bool sv = false, av = false;
#pragma unroll
      for (int s = 0; s < 3; ++s)
        if ((gc & GC_ST(s)) && (gc & GC_SY(s)) && !(gc & GC_TX(s))) {
          double rr = s == 0 ? P[XGC_P_rgc_sympt_seek_test_rr] : s == 2 ? P[XGC_P_pgc_sympt_seek_test_rr] : 1.0;
          if (xgc_bern(d(8 + s), P[XGC_P_ugc_sympt_seek_test_prob] * rr)) sv = true;
        }
      unsigned ex = (gc >> GC_EXPO_SHIFT) & 0x1Fu;
      if (!sv) {
        if (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL) av = avd;       // drawn by k_agent2a (AGENT2 k0)
        else if (proto == XGC_SCREEN_CDC) {
          bool active = (ex & 7u) || (kissx && (ex & 8u));
          bool hr = (hiv & HV_PREP) || (ex & 16u);
          double interval = (hr ? P[XGC_P_cdc_sti_hr_int] : P[XGC_P_cdc_sti_int]) * (52.0 / 12.0);
          bool due = ga.w == XGC_NA16 || (double)(at - (int)ga.w) >= interval;
          av = active && due;
        }
      }
      if (sv || av) {
        visit = 1; svisit = sv;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          bool inf = gc & GC_ST(s), symp = inf && (gc & GC_SY(s)), tested;
          bool exposed = (ex >> s) & 1u; if (s == 2 && kissx && (ex & 8u)) exposed = true;
          if (proto == XGC_SCREEN_CDC && !symp && !exposed) tested = false;
          else if (proto == XGC_SCREEN_UNIVERSAL) tested = true;
          else tested = xgc_bern(d(12 + s), symp ? P[XGC_P_gc_sympt_prob_test + s] : P[XGC_P_gc_asympt_prob_test + s]);
          if (!tested) continue;
          tst[s] = 1;
          if (inf) {
            pos[s] = 1; ind_sti = true;
            if (!(gc & GC_TX(s))) { gc |= GC_TX(s); txs[s] = 1; gb_dirty = true;
              if (s == 0) gb.x = (short)at; else if (s == 1) gb.y = (short)at; else gb.z = (short)at; }
          }
        }
        ga.w = (short)at; ga_dirty = true;
        gc &= ~GC_EXPO_MASK;
      }
    }
    if (ind_sti) g.tck[idx].y = (short)at;
    if (ga_dirty) g.gck_a[idx] = ga;
    if (gb_dirty) g.gck_b[idx] = gb;
    if (gd_dirty) g.gck_d[idx] = gd;
    if (gc != c.z) ((uint2*)&g.core[idx])[1].x = gc;
  }
  if (mask & XGC_M_STITX) {
    unsigned any = __ballot_sync(0xffffffffu, visit != 0);
    if (any) {
      tl.add(XGC_K_VISITS, visit); tl.add(XGC_K_VISITS_SYMPT, svisit);
#pragma unroll
      for (int s = 0; s < 3; ++s) { tl.add(XGC_K_TESTED + s, tst[s]); tl.add(XGC_K_POS + s, pos[s]); tl.add(XGC_K_TXSTART + s, txs[s]); }
    }
  }
  }
  tl.flush(row);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_edge2: hivtrans_msm + stitrans_msm_rand over the site-/sero-discordant edges (sim1_02_lhs.r:214, 217).
// Condom use and position of each act are evaluated lazily from their indexed draws (condoms_msm, position_msm).
// New infections are OR-ed into the agents' core words; k_agent3 applies them (unique() semantics).
//
// Only a minority of edges is discordant and the per-act loops have different lengths per act type, so a
// one-thread-per-edge loop runs at ~4 active lanes per warp instruction.  Instead each CTA
//   1. classifies its 256 edges (record + two 128-bit endpoint gathers, staged in shared memory),
//   2. ballot-compacts the (edge, act type) pairs that can transmit into a dense work queue, sorted by act type,
//   3. lets consecutive threads take consecutive queue entries, so warps run one act type with all lanes busy.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void e2_flag(uint4* core, const int4& ed, unsigned new_a, unsigned new_b, bool hiv_a, bool hiv_b) {
  if (new_a) atomicOr(&core[ed.x].z, new_a << GC_NEW_SHIFT);
  if (new_b) atomicOr(&core[ed.y].z, new_b << GC_NEW_SHIFT);
  if (hiv_a) atomicOr(&core[ed.x].w, HV_NEW);
  if (hiv_b) atomicOr(&core[ed.y].w, HV_NEW);
}
template <bool INJ> __device__ __forceinline__ void e2_anal(const Dev& g, const Inj& inj, unsigned mask, int r, int l, int e, int at, const double* P,
                                     const int4& ed, const uint4& ca, const uint4& cb, uint4* core, double dt) {
  EdgeCtx x; x.r = r; x.l = l; x.e = e; x.at = at; x.ed = ed; x.ca = ca; x.cb = cb; x.base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
  const Aux* aux = g.aux + (size_t)r * g.n_cap;
  Aux aa = aux[ed.x], ab = aux[ed.y];
  int ai = ed.w & 0xFF;
  unsigned ga = ca.z & 7u, gb = cb.z & 7u;
  bool hivdisc = (mask & XGC_M_HIVTRANS) && ((ca.w ^ cb.w) & HV_STATUS);
  bool d_ai = (mask & XGC_M_STITRANS) && ((((ga >> 1) & 1u) && !(gb & 1u)) || ((gb & 1u) && !((ga >> 1) & 1u)) ||
                                          (((gb >> 1) & 1u) && !(ga & 1u)) || ((ga & 1u) && !((gb >> 1) & 1u)));
  x.age_a = aa.age_ref + dt; x.age_b = ab.age_ref + dt; x.insq_a = aa.ins_quot; x.insq_b = ab.ins_quot;
  double pc = edge_cond_prob(g, P, x, true);
  DrawT<INJ> da(inj, g, r, at, x.base + (XGC_S_ANAL - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);     // one Philox block per act
  bool a_pos = ca.w & HV_STATUS;
  uint4 cp = a_pos ? ca : cb, cn = a_pos ? cb : ca;
  double vlf = 0.0; int nrace = (int)DM_RACE(cn.y);
  if (hivdisc) vlf = xgc_exp(((double)(a_pos ? aa.vl : ab.vl) - 4.5) * XGC_LN_2_45);
  unsigned gpre_n = (cn.z >> GC_PRE_SHIFT) & 7u;          // negative partner's GC status as hivtrans saw it
  unsigned new_a = 0, new_b = 0; bool hiv_a = false, hiv_b = false;
  for (int k = 0; k < ai; ++k) {
    int uai = xgc_bern(da(4 * k), 1.0 - pc);
    int p1ins = edge_ai_p1ins(x, da, k);
    if (hivdisc) {
      bool pos_ins = (p1ins != 0) == a_pos;
      double t;
      if (pos_ins) { t = P[XGC_P_urai_prob] * vlf; if (gpre_n & 1u) t = t * P[XGC_P_hiv_rgc_rr]; }
      else { t = P[XGC_P_uiai_prob] * vlf; if (cn.y & DM_CIRC) t = t * P[XGC_P_circ_rr]; if (gpre_n & 2u) t = t * P[XGC_P_hiv_ugc_rr]; }
      if (!uai) t = t * (1.0 - P[XGC_P_cond_eff] * (1.0 - P[XGC_P_cond_fail + nrace]));
      unsigned st = HV_STAGE(cp.w);
      if (st == 1 || st == 2) t = t * P[XGC_P_acute_rr];
      if (cn.w & HV_PREP) { unsigned cl = HV_PREPCLASS(cn.w); if (cl >= 1) t = t * P[XGC_P_prep_adhr_hr + cl - 1]; }
      t = t * P[XGC_P_trans_scale + nrace];
      if (t > 1.0) t = 1.0;
      if (xgc_bern(da(4 * k + 2), t)) { if (a_pos) hiv_b = true; else hiv_a = true; }
    }
    if (d_ai) {                                         // insertive partner's urethra <-> receptive partner's rectum
      unsigned gx = p1ins ? ga : gb, gy = p1ins ? gb : ga;
      bool ux = (gx >> 1) & 1u, ry = gy & 1u;
      if (ux && !ry) {
        double t = P[XGC_P_u2rgc_tprob];
        if (!uai) t = t * (1.0 - P[XGC_P_sti_cond_eff] * (1.0 - P[XGC_P_sti_cond_fail + (int)DM_RACE(p1ins ? cb.y : ca.y)]));
        if (xgc_bern(da(4 * k + 3), t)) { if (p1ins) new_b |= 1u << XGC_PATH_U2R; else new_a |= 1u << XGC_PATH_U2R; }
      }
      if (ry && !ux) {
        double t = P[XGC_P_r2ugc_tprob];
        if (!uai) t = t * (1.0 - P[XGC_P_sti_cond_eff] * (1.0 - P[XGC_P_sti_cond_fail + (int)DM_RACE(p1ins ? ca.y : cb.y)]));
        if (xgc_bern(da(4 * k + 3), t)) { if (p1ins) new_a |= 1u << XGC_PATH_R2U; else new_b |= 1u << XGC_PATH_R2U; }
      }
    }
  }
  e2_flag(core, ed, new_a, new_b, hiv_a, hiv_b);
}
// oral (type 1: urethra <-> pharynx) and rimming (type 3: pharynx <-> rectum): direction per act ~ Bernoulli(0.5) from
// the direction word.  Site status does not change within the week, so all acts of one direction share one transmission
// channel (active partner's site sx <-> other partner's site sy) and only "did any of them transmit" is observable:
// one Bernoulli(1 - (1-t)^m) per direction instead of m per-act draws -- no per-act loop, no divergence.
template <bool INJ> __device__ __forceinline__ void e2_directed(const Dev& g, const Inj& inj, int type, int r, int l, int e, int at, const double* P,
                                            const int4& ed, const uint4& ca, const uint4& cb, uint4* core, const double* tab) {
  int base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
  bool oral = type == 1;
  int n = (ed.w >> 8) & 0xFF;
  if (!oral) {                                              // this week's rimming count, evaluated only here
    int dummy;
    DrawT<INJ> da(inj, g, r, at, base + (XGC_S_ACTS - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);
    edge_kiss_rim(P, l, ca, cb, da, tab, false, true, dummy, n);
  }
  if (n == 0) return;
  unsigned ga = ca.z & 7u, gb = cb.z & 7u;
  int sx = oral ? 1 : 2, sy = oral ? 2 : 0;                 // active partner's site, other partner's site
  double tA = oral ? P[XGC_P_u2pgc_tprob] : P[XGC_P_p2rgc_tprob], tB = oral ? P[XGC_P_p2ugc_tprob] : P[XGC_P_r2pgc_tprob];
  unsigned pA = oral ? XGC_PATH_U2P : XGC_PATH_P2R, pB = oral ? XGC_PATH_P2U : XGC_PATH_R2P;
  DrawT<INJ> dd(inj, g, r, at, base + ((oral ? XGC_S_ORAL : XGC_S_RIM) - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);
  unsigned new_a = 0, new_b = 0;
  if (INJ) {                                                // injected draws: every act is its own trial (abi.h xgc_stream_width)
    for (int j = 0; j < n; ++j) {
      bool p1 = dd(1 + 2 * j) < 0.5;                        // p1 is the active partner of act j
      bool ix = ((p1 ? ga : gb) >> sx) & 1u, iy = ((p1 ? gb : ga) >> sy) & 1u;
      if (ix == iy) continue;
      if (xgc_bern(dd(2 + 2 * j), ix ? tA : tB)) {          // ix: the active partner's site infects the other partner's; else the reverse
        if (ix) { if (p1) new_b |= 1u << pA; else new_a |= 1u << pA; } else { if (p1) new_a |= 1u << pB; else new_b |= 1u << pB; }
      }
    }
    e2_flag(core, ed, new_a, new_b, false, false);
    return;
  }
  unsigned dirword = dd.bits32(0);                          // act j: p1 is the active partner iff bit (31 - j) is set
  int m1 = __popc(dirword >> (32 - n)), m0 = n - m1;        // acts with p1 / p2 active (n in 1..32)
  {                                                         // p1 active: x = p1, y = p2
    bool ix = (ga >> sx) & 1u, iy = (gb >> sy) & 1u;
    if (m1 > 0 && ix != iy) {
      if (xgc_bern(dd(1), xgc_any_of_n(ix ? tA : tB, m1))) { if (ix) new_b |= 1u << pA; else new_a |= 1u << pB; }
    }
  }
  {                                                         // p2 active: x = p2, y = p1
    bool ix = (gb >> sx) & 1u, iy = (ga >> sy) & 1u;
    if (m0 > 0 && ix != iy) {
      if (xgc_bern(dd(2), xgc_any_of_n(ix ? tA : tB, m0))) { if (ix) new_a |= 1u << pA; else new_b |= 1u << pB; }
    }
  }
  e2_flag(core, ed, new_a, new_b, false, false);
}
template <bool INJ> __device__ __forceinline__ void e2_kiss(const Dev& g, const Inj& inj, int r, int l, int e, int at, const double* P,
                                                            const int4& ed, const uint4& ca, const uint4& cb, uint4* core, const double* tab) {
  int base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS, n, dummy;
  DrawT<INJ> da(inj, g, r, at, base + (XGC_S_ACTS - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);
  edge_kiss_rim(P, l, ca, cb, da, tab, true, false, n, dummy);       // this week's kissing count, evaluated only here
  if (n == 0) return;
  bool pa = (ca.z >> 2) & 1u, pb = (cb.z >> 2) & 1u;
  if (pa == pb) return;
  DrawT<INJ> d(inj, g, r, at, base + (XGC_S_KISS - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);
  bool hit = false;
  if (INJ) { for (int j = 0; j < n; ++j) hit |= (bool)xgc_bern(d(j), P[XGC_P_p2pgc_tprob]); }      // injected draws: one trial per kiss
  else hit = xgc_bern(d(0), xgc_any_of_n(P[XGC_P_p2pgc_tprob], n));
  if (hit) e2_flag(core, ed, pa ? 0u : 1u << XGC_PATH_P2P, pa ? 1u << XGC_PATH_P2P : 0u, false, false);
}

// k_edge2a: classify every edge and append the (edge, act type) pairs that can transmit to the replicate's per-type
// work queues (one atomicAdd per CTA and type).  No draws here.
__global__ void __launch_bounds__(TPB) k_edge2a(const __grid_constant__ Dev g, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  const Rep* rp = g.rep + r;
  const int ne0 = rp->ne[0], ne01 = ne0 + rp->ne[1], ne_all = ne01 + rp->ne[2];     // one index space over the three layers
  if ((int)(blockIdx.x * TPB) >= ne_all) return;
  static_assert(TPB == 256, "queue offsets assume 8 warps per CTA");
  __shared__ unsigned short s_wc[4][TPB / 32];
  __shared__ unsigned s_off[4][TPB / 32];
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int* C = g.control + (size_t)r * XGC_NCONTROL;
  const uint4* core = g.core + (size_t)r * g.n_cap;
  unsigned* const Qr = g.q_items + (size_t)r * 4 * g.e_stride;      // this replicate's four queues; 32-bit offsets below
  const unsigned estride = (unsigned)g.e_stride;
  for (int fbase = blockIdx.x * TPB; fbase < ne_all; fbase += gridDim.x * TPB) {
  const int fi = fbase + tid;
  const int l = fi >= ne01 ? 2 : fi >= ne0 ? 1 : 0;
  const int e = fi - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
  const int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  bool f[4] = { false, false, false, false };
  if (fi < ne_all) {
    int4 ed = E[e];
    uint4 ca = core[ed.x], cb = core[ed.y];
    int ai = ed.w & 0xFF, oi = (ed.w >> 8) & 0xFF;
    unsigned x = ca.z & 7u, y = cb.z & 7u;                // current site status: bit0 rect, bit1 ureth, bit2 phar
    bool gct = mask & XGC_M_STITRANS;
    bool hivdisc = (mask & XGC_M_HIVTRANS) && ai > 0 && ((ca.w ^ cb.w) & HV_STATUS);
    // a pathway pair (site s of one partner, site t of the other) can transmit iff exactly one of the two is infected
    bool d_ai = (((x >> 1) ^ y) | ((y >> 1) ^ x)) & 1u;                    // urethra <-> rectum
    bool d_oi = (((x >> 1) ^ (y >> 2)) | ((y >> 1) ^ (x >> 2))) & 1u;      // urethra <-> pharynx
    bool d_ri = (((x >> 2) ^ y) | ((y >> 2) ^ x)) & 1u;                    // pharynx <-> rectum
    bool d_ki = ((x ^ y) >> 2) & 1u;                                       // pharynx <-> pharynx
    f[0] = hivdisc || (gct && ai > 0 && d_ai);
    f[1] = gct && oi > 0 && d_oi;
    // kissing / rimming: queued on site discordance; their counts are drawn in k_edge2b
    bool stopped = act_stopped(ca.y, ca.z >> GC_PRE_SHIFT) || act_stopped(cb.y, cb.z >> GC_PRE_SHIFT);
    f[2] = gct && !stopped && C[XGC_C_transRoute_Kissing] && d_ki;
    f[3] = gct && !stopped && C[XGC_C_transRoute_Rimming] && d_ri;
  }
  unsigned b[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) { b[t] = __ballot_sync(0xffffffffu, f[t]); if (lane == 0) s_wc[t][w] = (unsigned short)__popc(b[t]); }
  __syncthreads();
  // one warp turns the 4 x 8 per-warp counts into queue offsets: lane = type * 8 + warp; exclusive prefix over the 8
  // warps of a type by shuffles, plus the type's base from one atomicAdd per CTA and type
  if (tid < 32) {
    unsigned cnt = s_wc[tid >> 3][tid & 7], inc = cnt;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, inc, o, 8); if ((tid & 7) >= o) inc += v; }
    unsigned total = __shfl_sync(0xffffffffu, inc, 7, 8), base = 0;
    if ((tid & 7) == 0 && total) base = atomicAdd(g.q_count + (size_t)r * 4 + (tid >> 3), total);
    base = __shfl_sync(0xffffffffu, base, 0, 8);
    s_off[tid >> 3][tid & 7] = base + inc - cnt;
  }
  __syncthreads();
  const unsigned item = ((unsigned)l << 28) | (unsigned)e, lt = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (f[t]) Qr[(unsigned)t * estride + s_off[t][w] + __popc(b[t] & lt)] = item;
  // (no barrier needed before the next tile: s_wc is rewritten only after this tile's second barrier, s_off only after the next tile's first)
  }
}

// k_edge2b: dense processing of the work queues; blockIdx.y = act type, so every warp runs one act type with all lanes busy.
template <bool INJ, bool ANAL> __global__ void __launch_bounds__(TPB, ANAL ? 3 : 4) k_edge2b(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.z + g.r0, type = ANAL ? 0 : 1 + blockIdx.y;
  unsigned n = g.q_count[(size_t)r * 4 + type];
  if (blockIdx.x * TPB >= n) return;
  const Rep* rp = g.rep + r;
  int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  uint4* core = g.core + (size_t)r * g.n_cap;
  for (unsigned i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
    unsigned item = g.q_items[((size_t)r * 4 + type) * g.e_stride + i];
    int l = (int)(item >> 28), e = (int)(item & 0x0FFFFFFFu);
    int4 ed = g.edge[(size_t)r * g.e_stride + g.e_off[l] + e];
    uint4 ca = core[ed.x], cb = core[ed.y];
    const double* tab = g.pois + ((size_t)r * 4 + (l < 2 ? l : 0)) * (XGC_MAX_ACTS + 1);
    if (ANAL) e2_anal<INJ>(g, inj, mask, r, l, e, at, P, ed, ca, cb, core, (double)(rp->t_age - rp->t_ref) * XGC_WEEK);
    else if (type == 2) e2_kiss<INJ>(g, inj, r, l, e, at, P, ed, ca, cb, core, tab);
    else e2_directed<INJ>(g, inj, type, r, l, e, at, P, ed, ca, cb, core, tab);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_agent3: apply this week's new infections (unique ids; pathway attribution by the weekly random order) and tally
// prevalence_msm (sim1_02_lhs.r:218).
// k_agent3a reads every agent once: it tallies the prevalence series on the status the agent WILL have once its flagged
// infections are applied (a flagged site that is not yet infected becomes infected, whatever the draws say), and queues
// the flagged agents; it writes nothing to the agent planes.  k_agent3b applies the infections of the queued agents
// (a few per cent) with all lanes busy: symptom draw, Poisson duration, clocks, incidence and pathway tallies.
// ---------------------------------------------------------------------------------------------------------------------
#define A3_NHIST (XGC_K_NGROUP * XGC_NCELL)
#define A3B_NHIST (4 * XGC_NCELL + XGC_NPATH)      /* incid r,u,p by cell; incid HIV by cell; pathway winners */
#define A3_SITE_MASK(s) ((s) == 0 ? ((1u << XGC_PATH_U2R) | (1u << XGC_PATH_P2R)) : (s) == 1 ? ((1u << XGC_PATH_R2U) | (1u << XGC_PATH_P2U)) \
                                  : ((1u << XGC_PATH_U2P) | (1u << XGC_PATH_R2P) | (1u << XGC_PATH_P2P)))
__global__ void __launch_bounds__(TPB) k_agent3a(const __grid_constant__ Dev g, unsigned mask, int tpc) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  const Rep* rp = g.rep + r;
  const size_t rbase = (size_t)r * g.n_cap;
  uint4 cn = g.core[rbase + (size_t)blockIdx.x * tpc * TPB + threadIdx.x];
  int n_slots = rp->n_slots;
  int t0 = blockIdx.x * tpc, t1 = min(t0 + tpc, (n_slots + TPB - 1) / TPB);
  if (t0 >= t1) return;
  extern __shared__ unsigned cq_smem[];
  CtaQueue cq; cq.init(cq_smem);
  __syncthreads();
  __shared__ unsigned hist[A3_NHIST];
  for (int k = threadIdx.x; k < A3_NHIST; k += TPB) hist[k] = 0u;
  __syncthreads();
  int at = g.at_ptr[r];
  unsigned lane = threadIdx.x & 31;
  unsigned* Q = g.qa_items + rbase;
  const bool gct = mask & XGC_M_STITRANS, hvt = mask & XGC_M_HIVTRANS;
  for (int t = t0; t < t1; ++t) {
    int i = t * TPB + threadIdx.x;
    uint4 c = cn;
    if (t + 1 < t1) cn = g.core[rbase + i + TPB];
    bool live = i < n_slots && (c.y & DM_ACTIVE);
    unsigned gc = c.z, hiv = c.w, cell = DM_CELL(c.y);
    unsigned nm = gct ? (gc >> GC_NEW_SHIFT) & 0x7Fu : 0u;
    bool hnew = hvt && (hiv & HV_NEW);
    bool queue = live && (nm || hnew);
    if (queue) {      // status after k_agent3b: flagged susceptible sites become infected; a flagged HIV-negative becomes positive
      if (nm & A3_SITE_MASK(0)) gc |= GC_ST(0);
      if (nm & A3_SITE_MASK(1)) gc |= GC_ST(1);
      if (nm & A3_SITE_MASK(2)) gc |= GC_ST(2);
      if (hnew) hiv |= HV_STATUS;
    }
    cq.push(queue, (unsigned)i, lane);
    if (mask & XGC_M_PREV) {
      // warp-aggregated tallies: lanes of the same (race, age group) cell vote, one shared-memory add per cell and series
      unsigned grp = __match_any_sync(0xffffffffu, live ? cell : 31u);
      bool leader = live && lane == (unsigned)(__ffs(grp) - 1);
#define A3_TALLY(group, pred) { unsigned m_ = __ballot_sync(0xffffffffu, live && (pred)) & grp; \
                                if (leader && m_) atomicAdd(&hist[(group) * XGC_NCELL + cell], (unsigned)__popc(m_)); }
      A3_TALLY(XGC_K_NUM, true)
      A3_TALLY(XGC_K_INUM, hiv & HV_STATUS)
      A3_TALLY(XGC_K_INUM_DX, hiv & HV_DIAG)
      A3_TALLY(XGC_K_VSUPP, (hiv & HV_DIAG) && (hiv & HV_VSUPP))
      A3_TALLY(XGC_K_PREPELIG, hiv & HV_PREPELIG)
      A3_TALLY(XGC_K_PREPCURR, hiv & HV_PREP)
      A3_TALLY(XGC_K_INUM_GC, gc & 7u)
      A3_TALLY(XGC_K_INUM_RGC, gc & 1u)
      A3_TALLY(XGC_K_INUM_UGC, gc & 2u)
      A3_TALLY(XGC_K_INUM_PGC, gc & 4u)
#undef A3_TALLY
    }
  }
  cq.flush(g.q3_count + r, Q);
  if (!(mask & XGC_M_PREV)) return;
  __syncthreads();
  unsigned* row = counter_row(g, r, at);
  // the incidence groups are tallied by k_agent3b; only the series above can be non-zero here
  for (int k = threadIdx.x; k < A3_NHIST; k += TPB) {
    unsigned v = hist[k];
    if (v) atomicAdd(row + k, v);
  }
}

template <bool INJ> __global__ void __launch_bounds__(TPB) k_agent3b(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, unsigned mask) {
  pdl_prologue();
  int r = blockIdx.y + g.r0;
  unsigned n = g.q3_count[r];
  if (blockIdx.x * TPB >= n) return;
  const Rep* rp = g.rep + r;
  const size_t rbase = (size_t)r * g.n_cap;
  int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const int* C = g.control + (size_t)r * XGC_NCONTROL;
  unsigned* row = counter_row(g, r, at);
  // incidence tallies go through a per-CTA histogram: with few replicates on the device (1M-agent configs) direct global
  // atomics from every queued agent would all land on the same counter row
  __shared__ unsigned hist[A3B_NHIST];
  for (int k = threadIdx.x; k < A3B_NHIST; k += TPB) hist[k] = 0u;
  __syncthreads();
  for (unsigned qi = blockIdx.x * TPB + threadIdx.x; qi < n; qi += gridDim.x * TPB) {
  size_t idx = rbase + g.qa_items[rbase + qi];
  uint4 c = g.core[idx];
  short4 ga = g.gck_a[idx], gb = g.gck_b[idx], gd = na4();       // issued with the core word: one round trip (nearly every queued agent has a GC flag)
  unsigned uid = c.x, gc = c.z, hiv = c.w, cell = DM_CELL(c.y);
  unsigned nm = (gc >> GC_NEW_SHIFT) & 0x7Fu;
  if (nm && (mask & XGC_M_STITRANS)) {
    bool pois = C[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_POIS;
    if (pois) gd = g.gck_d[idx];
    DrawT<INJ> d(inj, g, r, at, XGC_S_INFECT, uid, 0u, uid);
    bool any = false;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      unsigned cand = nm & A3_SITE_MASK(s);
      if (!cand || (gc & GC_ST(s))) continue;
      int win = -1, wr = 99;
      for (int k = 0; k < XGC_NPATH; ++k) if ((cand >> k) & 1u) { int rk = rp->path_rank[k]; if (rk < wr) { wr = rk; win = k; } }
      gc = (gc | GC_ST(s)) & ~(GC_SY(s) | GC_TX(s));
      if (xgc_bern(d(s), P[XGC_P_gc_sympt_prob + s])) gc |= GC_SY(s);
      short du = -1;
      if (pois) { int q = xgc_pois(d(3 + s), P[XGC_P_gc_ntx_int + s], 255); du = (short)(q < 1 ? 1 : q); }
      if (s == 0) { ga.x = (short)at; gb.x = -1; gd.x = du; } else if (s == 1) { ga.y = (short)at; gb.y = -1; gd.y = du; } else { ga.z = (short)at; gb.z = -1; gd.z = du; }
      atomicAdd(&hist[s * XGC_NCELL + cell], 1u);
      atomicAdd(&hist[4 * XGC_NCELL + win], 1u);
      any = true;
    }
    if (any) { g.gck_a[idx] = ga; g.gck_b[idx] = gb; if (pois) g.gck_d[idx] = gd; }
  }
  if (mask & XGC_M_STITRANS) gc &= ~GC_NEW_MASK;
  if ((hiv & HV_NEW) && (mask & XGC_M_HIVTRANS)) {
    if (!(hiv & HV_STATUS)) {
      hiv = (hiv & ~(HV_STAGE_MASK | HV_DIAG | HV_TX)) | HV_STATUS | HV_VSUPP | (1u << 1);
      short4 ha = g.hck_a[idx]; ha.x = (short)at; ha.z = (short)at; g.hck_a[idx] = ha;
      short4 hb = g.hck_b[idx]; hb.x = 0; hb.y = 0; g.hck_b[idx] = hb;
      g.aux[idx].vl = 0.f;
      atomicAdd(&hist[3 * XGC_NCELL + cell], 1u);
    }
  }
  if (mask & XGC_M_HIVTRANS) hiv &= ~HV_NEW;
  if (gc != c.z || hiv != c.w) { uint2* q = (uint2*)&g.core[idx]; q[1] = make_uint2(gc, hiv); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < A3B_NHIST; k += TPB) {
    unsigned v = hist[k];
    if (!v) continue;
    int grp = k / XGC_NCELL, cell = k % XGC_NCELL;
    atomicAdd(row + (grp < 3 ? (XGC_K_INCID_RGC + grp) * XGC_NCELL + cell : grp == 3 ? XGC_K_INCID_HIV * XGC_NCELL + cell : XGC_K_PATH + (k - 4 * XGC_NCELL)), v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_targets: tail-window means of the target ratios, NA/NaN -> 0 per week (inst/cal/sim0.0_setup.r:205-224,272,299-308;
// R/calculate_targets.r:14-190).  One thread per (replicate, target); sequential in `at` so the sum order is fixed.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gsum(const unsigned* c, int group, unsigned cells) {
  double s = 0.0;
  for (int k = 0; k < XGC_NCELL; ++k) if ((cells >> k) & 1u) s += (double)c[group * XGC_NCELL + k];
  return s;
}
__device__ __forceinline__ double nz(double x) { return (x != x || x > 1.7e308 || x < -1.7e308) ? 0.0 : x; }
__device__ double target_week(const unsigned* c, int k) {
  const unsigned ALL = 0xFFFFFu;
  double num = gsum(c, XGC_K_NUM, ALL);
  if (k == XGC_TG_NUM) return num;
  if (k == XGC_TG_I_PREV) return gsum(c, XGC_K_INUM, ALL) / num;
  if (k == XGC_TG_IR100_POP) return gsum(c, XGC_K_INCID_HIV, ALL) / num * 5200.0;
  if (k < XGC_TG_PROP_TESTED) {
    unsigned m; int what;
    if (k < XGC_TG_HIVDX_AGE) { m = 0x1Fu << ((k - XGC_TG_HIVDX_RACE) * 5); what = 0; }
    else if (k < XGC_TG_VLS_RACE) { m = 0x08421u << (k - XGC_TG_HIVDX_AGE); what = 0; }
    else if (k < XGC_TG_VLS_AGE) { m = 0x1Fu << ((k - XGC_TG_VLS_RACE) * 5); what = 1; }
    else if (k < XGC_TG_PREPCOV_RACE) { m = 0x08421u << (k - XGC_TG_VLS_AGE); what = 1; }
    else { m = 0x1Fu << ((k - XGC_TG_PREPCOV_RACE) * 5); what = 2; }
    if (what == 0) { double n = gsum(c, XGC_K_NUM, m); return (gsum(c, XGC_K_INUM_DX, m) / n) / (gsum(c, XGC_K_INUM, m) / n); }
    if (what == 1) return gsum(c, XGC_K_VSUPP, m) / gsum(c, XGC_K_INUM_DX, m);
    return gsum(c, XGC_K_PREPCURR, m) / gsum(c, XGC_K_PREPELIG, m);
  }
  if (k < XGC_TG_PROB_POS) return (double)c[XGC_K_TESTED + k - XGC_TG_PROP_TESTED] / (double)c[XGC_K_VISITS];
  if (k < XGC_TG_PREV_GC) return (double)c[XGC_K_POS + k - XGC_TG_PROB_POS] / (double)c[XGC_K_TESTED + k - XGC_TG_PROB_POS];
  if (k == XGC_TG_PREV_GC) return gsum(c, XGC_K_INUM_GC, ALL) / num;
  if (k < XGC_TG_IR100_GC) return gsum(c, XGC_K_INUM_RGC + (k - XGC_TG_PREV_GC - 1), ALL) / num;
  if (k == XGC_TG_IR100_GC) {
    double inc = 0.0;
    for (int s = 0; s < 3; ++s) inc += gsum(c, XGC_K_INCID_RGC + s, ALL);
    return inc / (num - gsum(c, XGC_K_INUM_GC, ALL)) * 5200.0;
  }
  if (k >= XGC_TG_AGEDX) {
    const int cells[8] = XGC_TG_AGEDX_CELLS; int cell = cells[k - XGC_TG_AGEDX];
    return (double)c[XGC_K_INUM_DX * XGC_NCELL + cell] / gsum(c, XGC_K_INUM_DX, 0x1Fu << ((cell / 5) * 5));
  }
  int s = k - XGC_TG_IR100_GC - 1;
  return gsum(c, XGC_K_INCID_RGC + s, ALL) / (num - gsum(c, XGC_K_INUM_RGC + s, ALL)) * 5200.0;
}
__global__ void k_targets(const __grid_constant__ Dev g, int at_last, int tailn, double* out) {
  pdl_prologue();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.R * XGC_NTARGET) return;
  int r = t / XGC_NTARGET, k = t % XGC_NTARGET;
  double s = 0.0;
  for (int at = at_last - tailn + 1; at <= at_last; ++at) s += nz(target_week(counter_row(g, r, at), k));
  out[t] = s / (double)tailn;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_compact: squeeze tombstones out of one replicate (stable, in place, one CTA per replicate) and re-number edges.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_compact_agents(const __grid_constant__ Dev g, int* remap /*[R][n_cap] old slot -> new slot*/) {
  pdl_prologue();
  int r = blockIdx.x + g.r0, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Rep* rp = g.rep + r;
  int n_slots = rp->n_slots;
  size_t o = (size_t)r * g.n_cap;
  int* RM = remap + o;
  __shared__ int wsum[TPB / 32];
  __shared__ int base;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n_slots; i0 += TPB) {
    int i = i0 + tid; bool keep = false;
    uint4 c; Aux ax; short4 a, b, d, ha, hb, tk; int arr = 0;
    if (i < n_slots) {
      c = g.core[o + i]; keep = c.y & DM_ACTIVE;
      if (keep) { ax = g.aux[o + i]; a = g.gck_a[o + i]; b = g.gck_b[o + i]; d = g.gck_d[o + i]; ha = g.hck_a[o + i]; hb = g.hck_b[o + i]; tk = g.tck[o + i]; arr = g.arr_t[o + i]; }
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int k = 0; k < w; ++k) off += wsum[k];
    off += __popc(m & ((1u << lane) - 1u));
    if (i < n_slots) RM[i] = keep ? off : -1;
    if (keep) { g.core[o + off] = c; g.aux[o + off] = ax; g.gck_a[o + off] = a; g.gck_b[o + off] = b; g.gck_d[o + off] = d;
                g.hck_a[o + off] = ha; g.hck_b[o + off] = hb; g.tck[o + off] = tk; g.arr_t[o + off] = arr; }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int k = 0; k < TPB / 32; ++k) t += wsum[k]; base += t; }
    __syncthreads();
  }
  int n = base;
  // clear the freed tail (active bit off) and rebuild the alive bitmap
  for (int i = n + tid; i < n_slots; i += TPB) g.core[o + i] = make_uint4(0, 0, 0, 0);
  unsigned* A = g.alive + (size_t)r * (g.n_cap / 32);
  for (int wd = tid; wd < g.n_cap / 32; wd += TPB) {
    int lo = wd * 32;
    A[wd] = n >= lo + 32 ? 0xFFFFFFFFu : n <= lo ? 0u : ((1u << (n - lo)) - 1u);
  }
  if (tid == 0) { rp->n_slots = n; rp->n_slots_pre = n; rp->n_alive = n; }
}
// Large replicates (one CTA per replicate would walk ~4000 tiles of a 1M-agent replicate one after the other): the same
// stable compaction as k_compact_agents, as massively parallel gathers through k_rank's position -> slot plane.
//   k_compact_remap    RM[slot] = new slot (edges are re-numbered with it afterwards)
//   k_compact_gather   scratch[j] = plane[P2S[j]]   for every living position j, one plane at a time (out of place)
//   k_compact_putback  plane[j] = scratch[j]
//   k_compact_finish   clears the freed tail, rebuilds the alive bitmap, resets the slot counters
__global__ void __launch_bounds__(TPB) k_compact_remap(const __grid_constant__ Dev g, int* remap) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, j = blockIdx.x * TPB + threadIdx.x;
  if (j >= g.rep[r].n_alive) return;
  size_t o = (size_t)r * g.n_cap;
  remap[o + g.pos2slot[o + j]] = j;
}
template <class T> __global__ void __launch_bounds__(TPB) k_compact_gather(const __grid_constant__ Dev g, const T* plane, T* scratch) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, j = blockIdx.x * TPB + threadIdx.x;
  if (j >= g.rep[r].n_alive) return;
  size_t o = (size_t)r * g.n_cap;
  scratch[o + j] = plane[o + g.pos2slot[o + j]];
}
template <class T> __global__ void __launch_bounds__(TPB) k_compact_putback(const __grid_constant__ Dev g, T* plane, const T* scratch) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, j = blockIdx.x * TPB + threadIdx.x;
  if (j >= g.rep[r].n_alive) return;
  size_t o = (size_t)r * g.n_cap;
  plane[o + j] = scratch[o + j];
}
__global__ void __launch_bounds__(TPB) k_compact_finish(const __grid_constant__ Dev g, int stage) {
  pdl_prologue();
  int r = blockIdx.y + g.r0, i = blockIdx.x * TPB + threadIdx.x;
  Rep* rp = g.rep + r;
  size_t o = (size_t)r * g.n_cap;
  int n = rp->n_alive;
  if (stage == 0) {                      // n_slots is still the old extent here
    if (i >= n && i < rp->n_slots) g.core[o + i] = make_uint4(0, 0, 0, 0);
    if (i < g.n_cap / 32) {
      int lo = i * 32;
      g.alive[(size_t)r * (g.n_cap / 32) + i] = n >= lo + 32 ? 0xFFFFFFFFu : n <= lo ? 0u : ((1u << (n - lo)) - 1u);
    }
  } else if (i == 0) { rp->n_slots = n; rp->n_slots_pre = n; }
}
__global__ void __launch_bounds__(TPB) k_compact_edges(const __grid_constant__ Dev g, const int* remap) {
  pdl_prologue();
  int r = blockIdx.z + g.r0, l = blockIdx.y;
  int ne = g.rep[r].ne[l];
  int e = blockIdx.x * TPB + threadIdx.x;
  if (e >= ne) return;
  int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const int* RM = remap + (size_t)r * g.n_cap;
  int4 ed = E[e];
  ed.x = RM[ed.x]; ed.y = RM[ed.y];
  E[e] = ed;
}

// k_actcounts: materialise the lazily evaluated kissing / rimming counts of one replicate into the edge records
// (xgc_get_actcounts / dat$temp$el parity); the hot path never needs them for non-discordant edges.
template <bool INJ> __global__ void k_actcounts(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, int r, int* out /*[3][emax]: ki | ri << 8*/, int emax) {
  pdl_prologue();
  int l = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.rep[r].ne[l]) return;
  int at = g.at_ptr[r];
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
  const uint4* core = g.core + (size_t)r * g.n_cap;
  int4 ed = E[e];
  uint4 ca = core[ed.x], cb = core[ed.y];
  const double* tab = g.pois + ((size_t)r * 4 + (l < 2 ? l : 0)) * (XGC_MAX_ACTS + 1);
  int ki, ri;
  DrawT<INJ> da(inj, g, r, at, XGC_S_EDGE0 + l * XGC_EDGE_STREAMS + (XGC_S_ACTS - XGC_S_EDGE0), ca.x, cb.x, (size_t)e);
  edge_kiss_rim(P, l, ca, cb, da, tab, true, true, ki, ri);
  out[(size_t)l * emax + e] = ki | (ri << 8);
}

// k_actlist: materialise dat$temp$al for one replicate (debug / parity of condoms_msm + position_msm)
template <bool INJ> __global__ void k_actlist(const __grid_constant__ Dev g, const __grid_constant__ Inj inj, int r, int* al, int cap, int* n_out) {
  // single thread: tiny test-only kernel, rows must come out in (layer, edge, type, k) order
  pdl_prologue();
  if (blockIdx.x || threadIdx.x) return;
  const Rep* rp = g.rep + r;
  int at = g.at_ptr[r], n = 0;
  const double* P = g.params + (size_t)r * XGC_NPARAM;
  const uint4* core = g.core + (size_t)r * g.n_cap;
  const Aux* aux = g.aux + (size_t)r * g.n_cap;
  double dt = (double)(rp->t_age - rp->t_ref) * XGC_WEEK;
  for (int l = 0; l < 3; ++l) {
    const int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
    for (int e = 0; e < rp->ne[l]; ++e) {
      EdgeCtx x; x.r = r; x.l = l; x.e = e; x.at = at; x.ed = E[e];
      x.ca = core[x.ed.x]; x.cb = core[x.ed.y]; x.base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
      Aux aa = aux[x.ed.x], ab = aux[x.ed.y];
      x.age_a = aa.age_ref + dt; x.age_b = ab.age_ref + dt; x.insq_a = aa.ins_quot; x.insq_b = ab.ins_quot;
      int cnt[4] = { x.ed.w & 0xFF, (x.ed.w >> 8) & 0xFF, 0, 0 };
      {
        DrawT<INJ> dacts(inj, g, r, at, x.base + (XGC_S_ACTS - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
        edge_kiss_rim(P, l, x.ca, x.cb, dacts, g.pois + ((size_t)r * 4 + (l < 2 ? l : 0)) * (XGC_MAX_ACTS + 1), true, true, cnt[2], cnt[3]);
      }
      double pc = cnt[0] > 0 ? edge_cond_prob(g, P, x, false) : 0.0;
      DrawT<INJ> dc(inj, g, r, at, x.base + (XGC_S_ANAL - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
      DrawT<INJ> doi(inj, g, r, at, x.base + (XGC_S_ORAL - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
      DrawT<INJ> dri(inj, g, r, at, x.base + (XGC_S_RIM - XGC_S_EDGE0), x.ca.x, x.cb.x, (size_t)e);
      for (int type = 0; type < 4; ++type)
        for (int k = 0; k < cnt[type]; ++k) {
          int uai = 0, p1 = 0;
          if (type == 0) { uai = xgc_bern(dc(4 * k), 1.0 - pc); p1 = edge_ai_p1ins(x, dc, k); }
          else if (type == 1) p1 = INJ ? (int)(doi(1 + 2 * k) < 0.5) : (int)((doi.bits32(0) >> (31 - k)) & 1u);
          else if (type == 3) p1 = INJ ? (int)(dri(1 + 2 * k) < 0.5) : (int)((dri.bits32(0) >> (31 - k)) & 1u);
          if (n < cap) { al[n] = l; al[cap + n] = e; al[2 * cap + n] = type; al[3 * cap + n] = k; al[4 * cap + n] = uai; al[5 * cap + n] = p1; }
          n++;
        }
    }
  }
  *n_out = n;
}

__global__ void k_debug_exp(const double* x, double* y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = xgc_exp(x[i]);
}
__global__ void k_debug_philox(unsigned seed, unsigned rep, const uint4* ctr, uint4* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox4x32_10(ctr[i], make_uint2(seed, rep));
}
