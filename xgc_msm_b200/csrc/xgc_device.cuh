// Device-side layout and arithmetic of the B200 epidemic step (sm_100a).
//
// Packed structure-of-arrays agent state: every plane is indexed [replicate * n_cap + slot].  Slots are stable
// (departed agents leave a tombstone; xgc_compact squeezes them out on a schedule), so edges store slot indices and
// never need re-numbering when agents depart.  The R position of an agent (the index the reference uses in dat$attr
// and dat$el) is its rank among living slots.
//
// All probabilities are evaluated in IEEE double with the exact operation order of DESIGN.md §"arithmetic contract";
// the library is compiled with --fmad=false so that results are bit-identical to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "xgc/abi.h"

#define XGC_WEEK (7.0 / 365.0)
#define XGC_LOG10_200 2.301029995663981
#define XGC_LN_2_45 0.8960880245566357
#define XGC_NA16 ((short)-1)

// ---- core plane: uint4 {x uid, y demo, z gc, w hiv}: one 128-bit load per agent, one 128-bit gather per edge endpoint
// demo word
#define DM_RACE(d)      ((d) & 3u)                 // race-1
#define DM_AGEGRP(d)    (((d) >> 2) & 7u)          // age.grp-1
#define DM_ROLE(d)      (((d) >> 5) & 3u)          // 0 I, 1 R, 2 V
#define DM_RISK(d)      (((d) >> 7) & 7u)          // risk.grp-1
#define DM_CIRC         (1u << 10)
#define DM_LATE         (1u << 11)
#define DM_TTTRAJ(d)    (((d) >> 12) & 3u)         // 1..3
#define DM_STOPPER      (1u << 14)
#define DM_ACTIVE       (1u << 15)
#define DM_CELL(d)      (DM_RACE(d) * 5u + DM_AGEGRP(d))
// gc word: bit s status, 3+s sympt, 6+s tx (s = 0 rect, 1 ureth, 2 phar); 9..17 snapshot taken at the start of the
// prep/stirecov/stitx pass (what acts/condoms/hivtrans saw); 18..22 exposure history; 24..30 new-infection pathways
#define GC_ST(s)        (1u << (s))
#define GC_SY(s)        (1u << (3 + (s)))
#define GC_TX(s)        (1u << (6 + (s)))
#define GC_CUR_MASK     0x1FFu
#define GC_PRE_SHIFT    9
#define GC_EXPO_SHIFT   18
#define GC_EXPO_MASK    (0x1Fu << GC_EXPO_SHIFT)
#define GC_NEW_SHIFT    24
#define GC_NEW_MASK     (0x7Fu << GC_NEW_SHIFT)
// hiv word
#define HV_STATUS       (1u << 0)
#define HV_STAGE(h)     (((h) >> 1) & 7u)
#define HV_STAGE_MASK   (7u << 1)
#define HV_DIAG         (1u << 4)
#define HV_TX           (1u << 5)
#define HV_PREP         (1u << 6)
#define HV_PREPCLASS(h) (((h) >> 7) & 3u)
#define HV_PREPCLASS_MASK (3u << 7)
#define HV_PREPELIG     (1u << 9)
#define HV_PREP_PRE     (1u << 10)               // prepStat as acts/condoms saw it (before this week's prep module)
#define HV_VSUPP        (1u << 11)              // diagnosed-independent: vl <= log10(200), maintained by hivvl / attribute upload
#define HV_NEW          (1u << 31)

struct __align__(16) Aux { double age_ref; float vl; float ins_quot; };
// clock planes (short4, NA = -1):
//   gck_a {infT r,u,p, last_screen}   gck_b {txT r,u,p, spare}   gck_d {dur r,u,p, spare} ("pois" mode only)
//   hck_a {inf_time, diag_time, stage_time, tx_init_time}         hck_b {cuml_on, cuml_off, spare, spare}
//   tck   {last_neg_test, prep_ind_last, prep_last_risk, prep_start_time}
// edge record: int4 {p1 slot, p2 slot, start week, acts = ai | oi<<8 | kiss<<16 | rim<<24}

struct Rep {                 // per-replicate scalars
  int n_slots;               // high-water mark of used slots
  int n_slots_pre;           // n_slots before this week's arrivals (arrivals are not aged / exposed to mortality)
  int next_uid;
  int n_alive;
  int t_ref, t_age;          // age(at) = age_ref + (t_age - t_ref) * 7/365
  int ne[3];
  unsigned status;
  int path_rank[XGC_NPATH];  // this week's pathway order (stitrans_msm_rand)
  unsigned lb_epoch[3];      // launch counter of the look-back edge compaction, per layer (stale tile status words never match)
  int pad[2];
};

struct Inj {                 // injected-draw table (device copy), u == nullptr -> native Philox
  const double* u; size_t stride; size_t off[XGC_NSTREAM]; int form_cap;
};

struct Dev {                 // everything a kernel needs; passed by value
  int R, n_cap, t_cap, rep0; int e_cap[3]; size_t e_off[3]; size_t e_stride;
  const int* rep_id;         // [R] GLOBAL replicate id of each replicate (Philox key; default replicate_offset + r, xgc_set_replicate_ids)
  unsigned long long* rep_cycles;      // [R] SM cycles the resident kernel spent on each replicate (xgc_replicate_cost)
  int r0;                    // first replicate of THIS launch and R = how many it covers: xgc_run drives the replicates as two concurrent lanes
  unsigned seed;
  double inv_dur[2];         // 1 / mean duration of main, casual ties (surrogate dissolution probability), set with the tables
  uint4* core; Aux* aux; short4 *gck_a, *gck_b, *gck_d, *hck_a, *hck_b, *tck; int* arr_t;
  unsigned* alive;           // [R][n_cap/32] bitmap of living slots
  int* pos2slot;             // [R][n_cap] rank -> slot (built by k_rank)
  unsigned* deg;             // [R][n_cap] current degree per slot, byte l = layer l (main, casual), in the network the week starts with: counted by the dissolution pass
                             // of the surrogate network, read by its formation step (degree-conditioned acceptance)
  int4* edge;                // [R][e_stride]: layer l at e_off[l]
  Rep* rep; const double* params; const int* control; const double* tables;
  const double* inv_ntx;     // [R][4] 1 / gc.ntx.int[site] (weekly untreated-recovery probability, "geom"): one division per replicate, not per site-week
  const double* pois;        // [R][4][XGC_MAX_ACTS+1] Poisson CDF tables of the constant weekly rates: kiss main/casual, rim main/casual
  unsigned* q_items;         // [R][4][e_stride] work queue of k_edge2: (layer << 28 | edge index) per act type
  unsigned* q_count;         // [R][4]
  unsigned* qa_items;        // [R][n_cap] work queue of k_agent2b: slot | asymptomatic-visit draw << 31
  unsigned* qa_count;        // [R]
  unsigned* q3_count;        // [R] agents with new-infection flags this week (queue of k_agent3b; items reuse qa_items)
  unsigned* qh_items;        // [R][n_cap] work queue of k_agent1b (HIV-positive agents): slot | tested << 31 | died << 30
  unsigned* qh_count;        // [R]
  unsigned long long* lb_status;   // [R][3][lb_tiles] decoupled look-back tile status of the multi-CTA edge compaction
  unsigned* lb_ticket;       // [R][3] tile tickets (reset every week by k_begin)
  int lb_tiles;
  unsigned* counters;        // [R][t_cap][XGC_NCOUNTER]
  const int* at_ptr;
};

// ---- Philox4x32-10 (Salmon et al. SC'11), counter-based: draw = pure function of (replicate, week, stream, entity, k)
template <int ROUNDS> __device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32<10>(c, k); }

// One stream of one entity; caches the Philox block of 4.  INJ = true reads the injected table instead (parity mode);
// the native kernels are compiled without that path.
template <bool INJ> struct DrawT {
  const double* u; unsigned seed, rep, at, stream, e0, e1; int blk; uint4 v;
  __device__ __forceinline__ DrawT(const Inj& j, const Dev& g, int r, int at_, int stream_, unsigned e0_, unsigned e1_, size_t idx)
      : u(nullptr), seed(g.seed), rep((unsigned)g.rep_id[r]), at((unsigned)at_), stream((unsigned)stream_), e0(e0_), e1(e1_), blk(-1) {
    if (INJ) u = j.u + (size_t)r * j.stride + j.off[stream_] + idx * (size_t)xgc_stream_width(stream_);
  }
  // compute block b now, while the warp is still converged (later uses of its 4 draws are then free of divergent copies)
  __device__ __forceinline__ void prime(int b) {
    if (INJ) return;
    v = philox4x32_10(make_uint4(at, (stream << 16) | (unsigned)b, e0, e1), make_uint2(seed, rep)); blk = b;
  }
  // the 32-bit word behind draw k (direction bits of oral / rimming acts): floor(u * 2^32)
  __device__ __forceinline__ unsigned bits32(int k) {
    if (INJ) return (unsigned)(u[k] * 4294967296.0);
    int b = k >> 2;
    if (b != blk) { v = philox4x32_10(make_uint4(at, (stream << 16) | (unsigned)b, e0, e1), make_uint2(seed, rep)); blk = b; }
    return (k & 3) == 0 ? v.x : (k & 3) == 1 ? v.y : (k & 3) == 2 ? v.z : v.w;
  }
  __device__ __forceinline__ double operator()(int k) {
    if (INJ) return u[k];
    int b = k >> 2;
    if (b != blk) { v = philox4x32_10(make_uint4(at, (stream << 16) | (unsigned)b, e0, e1), make_uint2(seed, rep)); blk = b; }
    unsigned x = (k & 3) == 0 ? v.x : (k & 3) == 1 ? v.y : (k & 3) == 2 ? v.z : v.w;
    // ((double)x + 0.5) * 2^-32 without an int->fp64 conversion (XU pipe): 1.x with x in the top 32 mantissa bits is
    // 1 + x * 2^-32 exactly; subtracting 1 (Sterbenz) and adding 2^-33 are both exact, so the value is bit-identical.
    return (__hiloint2double((int)(0x3FF00000u | (x >> 12)), (int)(x << 20)) - 1.0) + (1.0 / 8589934592.0);
  }
};

// Programmatic dependent launch: first statement of every weekly kernel.  launch_dependents lets the next kernel's CTAs
// become resident as soon as every CTA of this grid has started; wait blocks until the previous grid has completed and
// its writes are visible (a no-op for a kernel launched without the attribute).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- arithmetic contract (mirrors the oracle operation for operation; no FMA)
__device__ __forceinline__ double xgc_exp(double x) {
  if (x < -700.0) x = -700.0;
  if (x > 700.0) x = 700.0;
  double kf = floor(x * 1.4426950408889634 + 0.5);
  double r = x - kf * 6.93147180369123816490e-01;
  r = r - kf * 1.90821492927058770002e-10;
  double q = 1.0 / 6227020800.0;
  q = q * r + 1.0 / 479001600.0;
  q = q * r + 1.0 / 39916800.0;
  q = q * r + 1.0 / 3628800.0;
  q = q * r + 1.0 / 362880.0;
  q = q * r + 1.0 / 40320.0;
  q = q * r + 1.0 / 5040.0;
  q = q * r + 1.0 / 720.0;
  q = q * r + 1.0 / 120.0;
  q = q * r + 1.0 / 24.0;
  q = q * r + 1.0 / 6.0;
  q = q * r + 0.5;
  q = q * r + 1.0;
  q = q * r + 1.0;
  // kf is an integer-valued double in [-1011, 1011]: its two's-complement value sits in the low word of kf + 1.5 * 2^52
  int k = __double2loint(kf + 6755399441055744.0);
  return q * __hiloint2double((k + 1023) << 20, 0);
}
__device__ __forceinline__ int xgc_bern(double u, double p) { return p <= 0.5 ? (u >= 1.0 - p) : (u < p); }
// pr_k = pr_{k-1} * mu * (1/k): the reciprocals 1/1..1/32 are correctly rounded constants (no division in the hot loop);
// beyond 32 (arrival counts only) a true division is used.  The oracle does exactly the same.
__constant__ double XGC_INVK[33] = { 0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12,
                                     1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23,
                                     1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32 };
__device__ __forceinline__ double xgc_pois_next(double pr, double mu, int k) {
  return k <= 32 ? pr * mu * XGC_INVK[k] : pr * mu / (double)k;
}
__device__ __forceinline__ int xgc_pois(double u, double mu, int kmax) {
  if (!(mu > 0.0)) return 0;
  double pr = xgc_exp(-mu), cdf = pr;
  int k = 0;
  while (u > cdf && k < kmax) { k++; pr = xgc_pois_next(pr, mu, k); cdf = cdf + pr; }
  return k;
}
// same walk on a precomputed table tab[k] = cdf_k (built with the identical recurrence by k_derive): bit-identical result
__device__ __forceinline__ int xgc_pois_tab(double u, const double* tab) {
  int k = 0;
  while (u > tab[k] && k < XGC_MAX_ACTS) k++;
  return k;
}
__device__ __forceinline__ int xgc_categorical(double u, const double* prob, int n) {
  int r = 0; double c = prob[0];
  while (u >= c && r < n - 1) { r++; c = c + prob[r]; }
  return r;
}
__device__ __forceinline__ unsigned xgc_age_group(double age) {   // age.grp - 1
  return (unsigned)((age >= 25.0) + (age >= 35.0) + (age >= 45.0) + (age >= 55.0));
}
__device__ __forceinline__ double xgc_glm_eta(const double* b, double dur, int casual, double ca, int hcp, int prep, int r1, int r2) {
  // indicator covariates enter as selects instead of int->fp64 conversions: b * 1.0 == b and eta + b * 0.0 == eta
  double eta = b[XGC_G_B0];
  eta = eta + b[XGC_G_DUR] * dur;
  eta = eta + b[XGC_G_DUR2] * (dur * dur);
  eta = eta + b[XGC_G_RC + r1 * 4 + r2];
  eta = eta + (casual ? b[XGC_G_PT2] : 0.0);
  eta = eta + (casual ? b[XGC_G_DUR_PT2] * dur : 0.0);
  eta = eta + b[XGC_G_CA] * ca;
  eta = eta + b[XGC_G_CA2] * (ca * ca);
  eta = eta + (hcp ? b[XGC_G_HCP] : 0.0);
  eta = eta + (prep ? b[XGC_G_PREP] : 0.0);
  return eta;
}
__device__ __forceinline__ bool roles_compatible(unsigned da, unsigned db) {
  unsigned ra = DM_ROLE(da), rb = DM_ROLE(db);
  return !((ra == 0 && rb == 0) || (ra == 1 && rb == 1));
}
__device__ __forceinline__ bool act_stopped(unsigned demo, unsigned gc) {
  return (demo & DM_STOPPER) && (gc & 0x1F8u);      // any site symptomatic or on treatment
}
