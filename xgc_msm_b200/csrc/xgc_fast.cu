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
//     edge modules, the exposure / new-infection scatter, the work queues, the weekly tallies and the alive bitmap of
//     the network step are all shared-memory traffic.  HBM sees the edge records (streamed, L2-resident between weeks),
//     the clock planes of the minority of agents whose clocks change, and the write-back at the end of the launch.
//   * modules become PHASES separated by __syncthreads(), not kernels: no queue round trips through HBM, no launch gaps.
//   * rare per-agent events (natural mortality, HIV interval tests, background clinic visits, tie dissolution) are drawn
//     as geometric waiting gaps over the slot / edge index (exact for an iid Bernoulli sequence; heterogeneous
//     probabilities by thinning), so agents nothing happens to cost an integer scan, not a Philox block;
//   * probabilities are fp32 with FMA, exp via ex2.approx, Bernoulli trials are integer compares on raw Philox words.
//
// Module semantics are those of xgc_kernels.cuh / oracle (same order, same state machine); tests/test_gpu_fast_mode.py
// gates this path statistically against the oracle and structurally (recount of the tallies, accounting identities).
#include <cuda_runtime.h>
#include <stdint.h>
#include "xgc_fast.cuh"

#define FT 1024                    /* threads per CTA: 32 warps, one CTA per SM                                   */
#define FQ_REGION 2048             /* entries per work-queue region                                              */
#define FQ_REGIONS 4
#define FAC 2048                   /* agents per scan chunk (every agent pushes at most once per region)          */
#define FEC 2048                   /* edges per classification chunk                                              */
#define HI_MORT 0.02f              /* weekly death probabilities above this are drawn per agent, not gap-sampled   */

// transient bits of the staged gc word (never meaningful in HBM between weeks)
#define FG_TESTED (1u << 23)       /* HIV-tested this week                                                        */
#define FG_AVD    (1u << 31)       /* background clinic-visit draw of this week                                   */

enum { FS_PATH = 1, FS_ARRN, FS_ARRA, FS_MORT, FS_MORTHI, FS_TEST, FS_AG1, FS_DISS, FS_NETN = FS_DISS + 2, FS_FORM,
       FS_ACTS = FS_FORM + 3, FS_COND, FS_ANAL, FS_ORAL, FS_RIM, FS_KISS, FS_AVD, FS_PREP, FS_AG2, FS_INF };

struct Smem {                      // pointers into the dynamic shared memory of the CTA
  unsigned *vd, *vg, *vh; short* ind; unsigned short* ea; unsigned short* Q; unsigned* bm; int* pre; unsigned* kill;
  unsigned* row; float* P; float* G; float* asmr; unsigned* hib; int* sc; unsigned* qn; int* wsum; float* pk;
};
// scalar slots (sc[])
enum { SC_NSLOTS = 0, SC_NPRE, SC_NALIVE, SC_NEXTUID, SC_TREF, SC_TAGE, SC_NE0, SC_NE1, SC_NE2, SC_STATUS, SC_RUN, SC_PATH /*7*/,
       SC_TMP = SC_PATH + 8, SC_N = SC_TMP + 8 };

static __host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
static __host__ __device__ size_t fast_layout(int n_cap, size_t e_stride, size_t* off /*[18]*/) {
  size_t o = 0; int k = 0;
  off[k++] = o; o += al16((size_t)n_cap * 4);                   // vd
  off[k++] = o; o += al16((size_t)n_cap * 4);                   // vg
  off[k++] = o; o += al16((size_t)n_cap * 4);                   // vh
  off[k++] = o; o += al16((size_t)n_cap * 2);                   // ind
  off[k++] = o; o += al16(e_stride * 2);                        // ea
  off[k++] = o; o += al16((size_t)FQ_REGION * FQ_REGIONS * 2);  // Q
  off[k++] = o; o += al16((size_t)(n_cap / 32 + 1) * 4);        // bm
  off[k++] = o; o += al16((size_t)(n_cap / 32 + 1) * 4);        // pre
  off[k++] = o; o += al16((e_stride / 32 + 4) * 4);             // kill
  off[k++] = o; o += al16((size_t)XGC_NCOUNTER * 4);            // row
  off[k++] = o; o += al16((size_t)XGC_NPARAM * 4);              // P
  off[k++] = o; o += al16((size_t)4 * XGC_GLM_NCOEF * 4);       // G
  off[k++] = o; o += al16((size_t)4 * XGC_ASMR_AGES * 4);       // asmr
  off[k++] = o; o += al16(16 * 4);                              // hib
  off[k++] = o; o += al16((size_t)SC_N * 4);                    // sc
  off[k++] = o; o += al16(8 * 4);                               // qn
  off[k++] = o; o += al16(40 * 4);                              // wsum
  off[k++] = o; o += al16((size_t)4 * (XGC_MAX_ACTS + 1) * 4);  // pk: kiss / rim Poisson CDFs
  return o;
}
size_t xgc_fast_smem_bytes(int n_cap, size_t e_stride) {
  size_t off[18];
  size_t need = fast_layout(n_cap, e_stride, off);
  return (need <= 232448 && n_cap <= 16384 && e_stride <= 16384) ? need : 0;      // u16 queue items: slot < 2^14, edge < 2^14
}

// ---------------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 frand(uint2 key, unsigned at, unsigned stream, unsigned a, unsigned b) {
  return philox4x32_10(make_uint4(at, stream, a, b), key);
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }     // (0,1), 24 bits
__device__ __forceinline__ unsigned thr32(float p) { return __float2uint_rz(fminf(fmaxf(p, 0.f), 1.f) * 4294967296.f); }  // saturates at 2^32-1
__device__ __forceinline__ bool bern(unsigned x, float p) { return x < thr32(p); }
__device__ __forceinline__ bool bern16(unsigned x16, float p) { return x16 < (unsigned)(fminf(fmaxf(p, 0.f), 1.f) * 65536.f); }
__constant__ float F_INVK[33] = { 0.f, 1.f / 1, 1.f / 2, 1.f / 3, 1.f / 4, 1.f / 5, 1.f / 6, 1.f / 7, 1.f / 8, 1.f / 9, 1.f / 10, 1.f / 11, 1.f / 12,
                                  1.f / 13, 1.f / 14, 1.f / 15, 1.f / 16, 1.f / 17, 1.f / 18, 1.f / 19, 1.f / 20, 1.f / 21, 1.f / 22, 1.f / 23,
                                  1.f / 24, 1.f / 25, 1.f / 26, 1.f / 27, 1.f / 28, 1.f / 29, 1.f / 30, 1.f / 31, 1.f / 32 };
// Poisson by inversion of one uniform, truncated at kmax <= 32
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
// warp-aggregated push into a shared-memory queue region; reached by every lane of the warp
__device__ __forceinline__ void qpush(bool p, unsigned item, unsigned short* q, unsigned* cnt, unsigned lane) {
  unsigned m = __ballot_sync(0xffffffffu, p);
  if (!m) return;
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(cnt, (unsigned)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (p) { unsigned k = base + __popc(m & ((1u << lane) - 1u)); if (k < FQ_REGION) q[k] = (unsigned short)item; }
}
__device__ __forceinline__ void row_add(unsigned* row, int idx, bool pred) {        // reached by every lane
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(row + idx, (unsigned)__popc(m));
}
// Gap sampling of an iid Bernoulli(p) sequence over positions [lo, hi), lo a multiple of GSEG.  The range is cut into fixed
// segments of GSEG positions, dealt round-robin to the participating warps [w0, w0 + nw); inside a segment 32 geometric
// gaps are drawn at a time and prefix-summed.  A segment's successes depend only on (stream, segment index), never on hi or
// on which warp ran it, so the same week gives the same events whether it runs inside a multi-week launch or as a call of
// its own.  f(pos, x) is called for every success with the Philox block its gap came from (x.y, x.z, x.w are unused draws).
#define GSEG 512
template <class F>
__device__ __forceinline__ void gap_sample(uint2 key, unsigned at, unsigned stream, int lo, int hi, float p, int w0, int nw, F f) {
  int w = (int)(threadIdx.x >> 5) - w0, lane = threadIdx.x & 31;
  if (w < 0 || w >= nw || !(p > 0.f) || hi <= lo) return;
  const float ilq = p >= 1.f ? 0.f : 1.f / log1pf(-p);       // 1 / ln(1 - p) < 0; p >= 1: every position
  for (int a = lo + w * GSEG; a < hi; a += nw * GSEG) {
    const int b = min(a + GSEG, hi);
    int pos = a - 1;
    for (unsigned chunk = 0; pos < b; ++chunk) {
      uint4 x = frand(key, at, stream, (unsigned)(a / GSEG) | (chunk << 16), (unsigned)lane);
      float gf = __logf(u01(x.x)) * ilq;                       // >= 0
      int s = (int)fminf(gf, 1.0e6f) + 1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s = min(s + v, 1 << 28); }
      int me = pos + s;
      if (me < b) f(me, x);
      pos = __shfl_sync(0xffffffffu, me, 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// phase clocks (SM cycles summed over CTAs; thread 0 of each CTA): tools/time_modes.py --phases
enum { PH_STAGE = 0, PH_ARRIVE, PH_PRE_SAMPLE, PH_PRE_SCAN, PH_PRE_TEST, PH_PRE_HIV, PH_NET, PH_ACTS, PH_B_SCAN, PH_B_PREP, PH_B_GC, PH_TRANS_CLASSIFY,
       PH_TRANS, PH_APPLY, PH_TALLY, PH_WRITEBACK, PH_N };
__device__ unsigned long long g_fast_prof[PH_N];
#define TICK(k) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_fast_prof[k], (unsigned long long)(t_ - t_last)); t_last = t_; } } while (0)
__global__ void __launch_bounds__(FT, 1) k_week_fast(const __grid_constant__ Dev g, const __grid_constant__ FastArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem S;
  {
    size_t off[18]; fast_layout(g.n_cap, g.e_stride, off);
    S.vd = (unsigned*)(smem_raw + off[0]); S.vg = (unsigned*)(smem_raw + off[1]); S.vh = (unsigned*)(smem_raw + off[2]);
    S.ind = (short*)(smem_raw + off[3]); S.ea = (unsigned short*)(smem_raw + off[4]); S.Q = (unsigned short*)(smem_raw + off[5]);
    S.bm = (unsigned*)(smem_raw + off[6]); S.pre = (int*)(smem_raw + off[7]); S.kill = (unsigned*)(smem_raw + off[8]);
    S.row = (unsigned*)(smem_raw + off[9]); S.P = (float*)(smem_raw + off[10]); S.G = (float*)(smem_raw + off[11]);
    S.asmr = (float*)(smem_raw + off[12]); S.hib = (unsigned*)(smem_raw + off[13]); S.sc = (int*)(smem_raw + off[14]);
    S.qn = (unsigned*)(smem_raw + off[15]); S.wsum = (int*)(smem_raw + off[16]); S.pk = (float*)(smem_raw + off[17]);
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = S.P;
  long long t_last = clock64();
  for (int rr = blockIdx.x; rr < g.R; rr += gridDim.x) {
    const int r = rr + g.r0;
    Rep* rp = g.rep + r;
    const size_t rbase = (size_t)r * g.n_cap;
    const uint2 key = make_uint2(g.seed, (unsigned)(g.rep0 + r));
    const int* C = g.control + (size_t)r * XGC_NCONTROL;
    const int proto = C[XGC_C_stiScreeningProtocol], kissx = C[XGC_C_cdcExposureSite_Kissing], surrogate = C[XGC_C_network_mode] == 1;
    const bool pois_mode = C[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_POIS;
    const bool route_kiss = C[XGC_C_transRoute_Kissing], route_rim = C[XGC_C_transRoute_Rimming];
    __syncthreads();                                            // previous replicate's write-back done with the staging buffers
    // ------------------------------------------------------------------------------------------ stage the replicate
    if (tid == 0) {
      S.sc[SC_NSLOTS] = rp->n_slots; S.sc[SC_NPRE] = rp->n_slots_pre; S.sc[SC_NALIVE] = rp->n_alive; S.sc[SC_NEXTUID] = rp->next_uid;
      S.sc[SC_TREF] = rp->t_ref; S.sc[SC_TAGE] = rp->t_age; S.sc[SC_NE0] = rp->ne[0]; S.sc[SC_NE1] = rp->ne[1]; S.sc[SC_NE2] = rp->ne[2];
      S.sc[SC_STATUS] = 0;
      for (int k = 0; k < XGC_NPATH; ++k) S.sc[SC_PATH + k] = rp->path_rank[k];
    }
    for (int k = tid; k < XGC_NPARAM; k += FT) S.P[k] = (float)g.params[(size_t)r * XGC_NPARAM + k];
    for (int k = tid; k < 4 * XGC_GLM_NCOEF; k += FT) S.G[k] = (float)g.tables[XGC_T_GLM_AI + k];
    for (int k = tid; k < 4 * XGC_ASMR_AGES; k += FT) S.asmr[k] = (float)g.tables[XGC_T_ASMR + k];
    for (int k = tid; k < 4 * (XGC_MAX_ACTS + 1); k += FT) S.pk[k] = (float)g.pois[(size_t)r * 4 * (XGC_MAX_ACTS + 1) + k];
    if (tid < 16) S.hib[tid] = 0u;
    __syncthreads();
    if (tid < 4 * XGC_ASMR_AGES && S.asmr[tid] >= HI_MORT) atomicOr(&S.hib[(tid / XGC_ASMR_AGES) * 4 + ((tid % XGC_ASMR_AGES) >> 5)], 1u << ((tid % XGC_ASMR_AGES) & 31));
    {
      const int n0 = S.sc[SC_NSLOTS];
      for (int i = tid; i < n0; i += FT) {
        uint4 c = g.core[rbase + i];
        S.vd[i] = c.y; S.vg[i] = c.z; S.vh[i] = c.w; S.ind[i] = g.tck[rbase + i].y;
      }
    }
    __syncthreads();
    if (warp == 0) {                                            // largest gap-sampled weekly death probability
      float v = 0.f;
      for (int k = lane; k < 4 * XGC_ASMR_AGES; k += 32) { float a = S.asmr[k]; if (a < HI_MORT) v = fmaxf(v, a); }
#pragma unroll
      for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) S.sc[SC_TMP + 1] = __float_as_int(v);
    }
    __syncthreads();
    const float pmax_mort = __int_as_float(S.sc[SC_TMP + 1]);
    const float pmax_test = fmaxf(fmaxf(P[XGC_P_hiv_test_rate], P[XGC_P_hiv_test_rate + 1]), fmaxf(P[XGC_P_hiv_test_rate + 2], P[XGC_P_hiv_test_rate + 3]));
    const float arr_age = P[XGC_P_arrival_age];

    TICK(PH_STAGE);
    for (int at = A.at0; at <= A.at1; ++at) {
      const unsigned mask = A.mask;
      unsigned* const grow = g.counters + ((size_t)r * g.t_cap + (size_t)at) * XGC_NCOUNTER;
      const int at8 = (at + XF_BIRTH_OFF) * 8;
      // =================================================================================================== PRE
      if (A.phases & XF_PRE) {
        for (int k = tid; k < XGC_NCOUNTER; k += FT) S.row[k] = 0u;
        if (tid == 0) {
          S.sc[SC_NPRE] = S.sc[SC_NSLOTS];
          if (mask & XGC_M_AGING) S.sc[SC_TAGE] = at;
        }
        __syncthreads();
        // ---- arrival_msm: one warp
        if ((mask & XGC_M_ARRIVAL) && warp == 0) {
          float mu = P[XGC_P_a_rate] * P[XGC_P_n_init];
          int m = (int)ceilf(mu / 32.f); m = max(1, min(m, 32));
          int nnew = 0;
          { uint4 x = frand(key, at, FS_ARRN, lane, 0); if (lane < m) nnew = pois_f(u01(x.x), mu / (float)m, 255); }
#pragma unroll
          for (int o = 16; o; o >>= 1) nnew += __shfl_xor_sync(0xffffffffu, nnew, o);
          int n_slots = S.sc[SC_NSLOTS], uid0 = S.sc[SC_NEXTUID], t_ref = S.sc[SC_TREF], t_age = S.sc[SC_TAGE];
          if (n_slots + nnew > g.n_cap) { nnew = g.n_cap - n_slots; if (lane == 0) atomicOr(&S.sc[SC_STATUS], (int)XGC_ST_AGENT_OVERFLOW); }
          for (int j = lane; j < nnew; j += 32) {
            int i = n_slots + j; size_t idx = rbase + i;
            uint4 x = frand(key, at, FS_ARRA, (unsigned)i, 0), y = frand(key, at, FS_ARRA, (unsigned)i, 1);
            float rp4[4] = { (float)g.tables[XGC_T_RACE_PROP], (float)g.tables[XGC_T_RACE_PROP + 1], (float)g.tables[XGC_T_RACE_PROP + 2], (float)g.tables[XGC_T_RACE_PROP + 3] };
            int race = categorical_f(u01(x.x), rp4, 4);
            float ro[3] = { (float)g.tables[XGC_T_ROLE_PROB + race * 3], (float)g.tables[XGC_T_ROLE_PROB + race * 3 + 1], (float)g.tables[XGC_T_ROLE_PROB + race * 3 + 2] };
            int role = categorical_f(u01(x.y), ro, 3);
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
            S.vd[i] = demo; S.vg[i] = 0u; S.vh[i] = hiv; S.ind[i] = -1;
          }
          if (lane == 0) { S.sc[SC_NSLOTS] = n_slots + nnew; S.sc[SC_NEXTUID] = uid0 + nnew; S.sc[SC_NALIVE] += nnew; S.row[XGC_K_NNEW] = (unsigned)nnew; }
        }
        __syncthreads();
        TICK(PH_ARRIVE);
        const int n_slots = S.sc[SC_NSLOTS], n_pre = S.sc[SC_NPRE];
        const bool hivmods = mask & (XGC_M_HIVTEST | XGC_M_HIVTX | XGC_M_HIVPROGRESS | XGC_M_HIVVL);
        for (int base = 0; base < n_slots; base += FAC) {
          const int hi = min(base + FAC, n_slots);
          if (tid < 8) S.qn[tid] = 0u;
          __syncthreads();
          unsigned short* qh = S.Q;                             // region 0: HIV-positive agents of this chunk
          unsigned short* qt = S.Q + FQ_REGION;                 // regions 1-2: HIV-test candidates (bit 15: rule-based)
          // ---- departure_msm, natural mortality: geometric gaps at the largest table rate, thinned to (race, age)
          if (mask & XGC_M_DEPARTURE)
            gap_sample(key, at, FS_MORT, base, min(hi, n_pre), pmax_mort, 0, 8, [&](int i, const uint4& x) {
              unsigned d = S.vd[i];
              if (!(d & DM_ACTIVE)) return;
              int ai = min(XGC_ASMR_AGES - 1, max(0, (int)((float)(at8 - (int)((S.vh[i] >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK)) * (float)XF_W8)));
              float pm = S.asmr[DM_RACE(d) * XGC_ASMR_AGES + ai];
              if (pm < HI_MORT && u01(x.y) * pmax_mort < pm) { S.vd[i] = d & ~DM_ACTIVE; atomicAdd(&S.row[XGC_K_DEP_GEN], 1u); atomicSub(&S.sc[SC_NALIVE], 1); }
            });
          // ---- hivtest_msm, interval testing: gaps at the largest race rate, thinned; eligibility settled after the gather
          if (mask & XGC_M_HIVTEST)
            gap_sample(key, at, FS_TEST, base, hi, pmax_test, 8, 24, [&](int i, const uint4& x) {
              unsigned d = S.vd[i], h = S.vh[i];
              if ((d & DM_ACTIVE) && !(d & DM_LATE) && !(h & (HV_DIAG | HV_PREP)) && u01(x.y) * pmax_test < P[XGC_P_hiv_test_rate + DM_RACE(d)]) {
                unsigned k = atomicAdd(&S.qn[1], 1u);
                if (k < 2 * FQ_REGION) qt[k] = (unsigned short)(i - base);
              }
            });
          __syncthreads();
          TICK(PH_PRE_SAMPLE);
          // ---- scan of the chunk: aging, high-rate mortality (ageing out), HIV-positive queue, rule-based test candidates
          for (int i0 = base; i0 < hi; i0 += FT) {
            int i = i0 + tid;
            bool live = false, qpos = false, qrule = false;
            if (i < hi) {
              unsigned d = S.vd[i], h = S.vh[i];
              live = d & DM_ACTIVE;
              if (live) {
                float age = (float)(at8 - (int)((h >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK)) * (float)XF_W8;
                bool isnew = i >= n_pre;
                if ((mask & XGC_M_AGING) && !isnew) { unsigned ng = agegrp_f(age); if (ng != DM_AGEGRP(d)) { d = (d & ~(7u << 2)) | (ng << 2); S.vd[i] = d; } }
                if ((mask & XGC_M_DEPARTURE) && !isnew) {
                  int ai = min(XGC_ASMR_AGES - 1, max(0, (int)age));
                  if ((S.hib[DM_RACE(d) * 4 + (ai >> 5)] >> (ai & 31)) & 1u) {
                    uint4 x = frand(key, at, FS_MORTHI, (unsigned)i, 0);
                    if (bern(x.x, S.asmr[DM_RACE(d) * XGC_ASMR_AGES + ai])) { S.vd[i] = d & ~DM_ACTIVE; atomicAdd(&S.row[XGC_K_DEP_GEN], 1u); atomicSub(&S.sc[SC_NALIVE], 1); live = false; }
                  }
                }
                qpos = live && hivmods && (h & HV_STATUS) && !isnew;
                qrule = live && (mask & XGC_M_HIVTEST) && !(h & HV_DIAG) && ((h & HV_PREP) || HV_STAGE(h) == 4);
              }
            }
            qpush(qpos, (unsigned)(i - base), qh, &S.qn[0], lane);
            unsigned m = __ballot_sync(0xffffffffu, qrule);
            if (m) {
              unsigned b0 = 0; if (lane == 0) b0 = atomicAdd(&S.qn[1], (unsigned)__popc(m));
              b0 = __shfl_sync(0xffffffffu, b0, 0);
              if (qrule) { unsigned k = b0 + __popc(m & ((1u << lane) - 1u)); if (k < 2 * FQ_REGION) qt[k] = (unsigned short)((i - base) | 0x8000); }
            }
          }
          __syncthreads();
          TICK(PH_PRE_SCAN);
          // ---- HIV-test candidates: time since the last negative test / arrival decides
          {
            const unsigned nt = min(S.qn[1], 2u * FQ_REGION);
            for (unsigned q = tid; q < nt; q += FT) {
              unsigned item = qt[q]; int i = base + (int)(item & 0x7FFFu); bool rule = item & 0x8000u;
              unsigned d = S.vd[i], h = S.vh[i];
              if (!(d & DM_ACTIVE)) continue;
              short lnt = g.tck[rbase + i].x; int last = lnt == XGC_NA16 ? g.arr_t[rbase + i] : (int)lnt;
              float tsl = (float)(at - last);
              bool tested = rule ? ((HV_STAGE(h) == 4 && tsl >= P[XGC_P_aids_test_int]) || ((h & HV_PREP) && tsl >= P[XGC_P_prep_tst_int])) : tsl >= 2.f;
              if (tested && !(atomicOr(&S.vg[i], FG_TESTED) & FG_TESTED)) {
                atomicAdd(&S.row[XGC_K_HIVTESTS], 1u);
                if (!(h & HV_STATUS)) g.tck[rbase + i].x = (short)at;
              }
            }
          }
          __syncthreads();
          TICK(PH_PRE_TEST);
          // ---- HIV-positive agents: AIDS mortality, test result, hivtx_msm, hivprogress_msm, hivvl_msm
          {
            const unsigned nh = min(S.qn[0], (unsigned)FQ_REGION);
            for (unsigned q = tid; q < nh; q += FT) {
              int i = base + (int)qh[q]; size_t idx = rbase + i;
              unsigned d = S.vd[i], h = S.vh[i];
              if (!(d & DM_ACTIVE)) continue;                  // died of natural causes this week
              uint4 x = frand(key, at, FS_AG1, (unsigned)i, 0);
              if ((mask & XGC_M_DEPARTURE) && HV_STAGE(h) == 4 && bern(x.x, P[XGC_P_aids_mr])) {
                S.vd[i] = d & ~DM_ACTIVE; atomicAdd(&S.row[XGC_K_DEP_AIDS], 1u); atomicSub(&S.sc[SC_NALIVE], 1); continue;
              }
              short4 ha = g.hck_a[idx], hb = g.hck_b[idx]; float vl0 = g.aux[idx].vl;
              const unsigned h0 = h; bool ha_dirty = false, hb_dirty = false;
              int race = (int)DM_RACE(d); unsigned traj = DM_TTTRAJ(d);
              if ((mask & XGC_M_HIVTEST) && (S.vg[i] & FG_TESTED) && !(h & HV_DIAG)) {
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
                  bool aids = false; float on = (float)hb.x, off = (float)hb.y;
                  if (on == 0.f && tsi >= P[XGC_P_vl_aids_onset_int]) aids = true;
                  if (traj == 1 && on > 0.f && off / P[XGC_P_max_time_off_tx_part_int] + on / P[XGC_P_max_time_on_tx_part_int] >= 1.f) aids = true;
                  if (traj >= 2 && !(h & HV_TX) && on > 0.f && off >= P[XGC_P_max_time_off_tx_full_int]) aids = true;
                  if (aids) { stage = 4; ha.z = (short)at; ha_dirty = true; }
                }
                h = (h & ~HV_STAGE_MASK) | (stage << 1);
              }
              if (mask & XGC_M_HIVVL) {
                unsigned stage = HV_STAGE(h);
                float v, peak = P[XGC_P_vl_acute_peak], sp = P[XGC_P_vl_set_point];
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
              if (h != h0) S.vh[i] = h;
            }
          }
          __syncthreads();
          TICK(PH_PRE_HIV);
        }
      } else {
        for (int k = tid; k < XGC_NCOUNTER; k += FT) S.row[k] = grow[k];      // continuing a week started by an earlier call
        __syncthreads();
      }
      // =================================================================================================== NETWORK
      if (A.phases & (XF_PRUNE | XF_FORM)) {
        const bool form = (A.phases & XF_FORM) && surrogate;
        const int n_slots = S.sc[SC_NSLOTS];
        // alive bitmap + exclusive popcount prefix (rank -> slot selection of the formation step)
        const int nw = (n_slots + 31) >> 5;
        for (int i0 = 0; i0 < nw * 32; i0 += FT) {
          int i = i0 + tid;
          unsigned m = __ballot_sync(0xffffffffu, i < n_slots && (S.vd[i] & DM_ACTIVE));
          if (lane == 0 && (i >> 5) < nw) S.bm[i >> 5] = m;
        }
        __syncthreads();
        if (warp == 0) {
          int run = 0;
          for (int w0 = 0; w0 < nw; w0 += 32) {
            int c = w0 + lane < nw ? __popc(S.bm[w0 + lane]) : 0, inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (w0 + lane < nw) S.pre[w0 + lane] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
          }
          if (lane == 0) S.sc[SC_NALIVE] = run;
        }
        __syncthreads();
        const int n_alive = S.sc[SC_NALIVE];
        for (int l = 0; l < 3; ++l) {
          int4* E = g.edge + (size_t)r * g.e_stride + g.e_off[l];
          int ne = S.sc[SC_NE0 + l];
          const int kw = (ne + 31) >> 5;
          for (int k = tid; k < kw; k += FT) S.kill[k] = 0u;
          if (tid == 0) { S.sc[SC_RUN] = 0; S.sc[SC_TMP + 2] = 0; }
          __syncthreads();
          const bool diss = form && l < 2;
          // dissolution is drawn over the RANK of a tie among the ties whose endpoints are both alive, so that pruning the
          // ties of departed agents in an earlier call (host-network loop) or in this pass gives the same result
          if (diss) gap_sample(key, at, FS_DISS + l, 0, ne, (float)g.inv_dur[l], 0, 8, [&](int e, const uint4&) { atomicOr(&S.kill[e >> 5], 1u << (e & 31)); });
          __syncthreads();
          if (form && l == 2) ne = 0;                           // one-off contacts last one week
          // stable in-place compaction: drop ties with a departed endpoint, then dissolved ties
          for (int e0 = 0; e0 < ne; e0 += FT) {
            int e = e0 + tid; bool keep = false; int4 ed = make_int4(0, 0, 0, 0);
            if (e < ne) { ed = E[e]; keep = (S.vd[ed.x] & DM_ACTIVE) && (S.vd[ed.y] & DM_ACTIVE); }
            for (int pass = 0; pass < (diss ? 2 : 1); ++pass) {
              unsigned m = __ballot_sync(0xffffffffu, keep);
              if (lane == 0) S.wsum[warp] = __popc(m);
              __syncthreads();
              if (warp == 0) {
                int c = S.wsum[lane], inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                S.wsum[lane] = inc - c;
                if (lane == 31) S.sc[SC_TMP] = inc;
              }
              __syncthreads();
              const int rank = S.sc[pass == 0 && diss ? SC_TMP + 2 : SC_RUN] + S.wsum[warp] + __popc(m & ((1u << lane) - 1u));
              if (pass == 0 && diss) { if (keep && ((S.kill[rank >> 5] >> (rank & 31)) & 1u)) keep = false; }
              else if (keep && rank != e) E[rank] = ed;
              __syncthreads();
              if (tid == 0) S.sc[pass == 0 && diss ? SC_TMP + 2 : SC_RUN] += S.sc[SC_TMP];
              __syncthreads();
            }
          }
          ne = S.sc[SC_RUN];
          if (form) {                                            // uniform role-compatible pairs holding the mean degree
            float target = (float)g.tables[XGC_T_NET_DEG + l] * (float)n_alive * 0.5f;
            float mu = l == 2 ? target : target / (float)g.tables[XGC_T_NET_DUR + l];
            float fl = floorf(mu);
            int nform = (int)fl + (u01(frand(key, at, FS_NETN, l, 0).x) < mu - fl);
            if (n_alive < 2) nform = 0;
            auto slot_of = [&](int pos) -> int {
              int lo = 0, hi = nw - 1;
              while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (S.pre[mid] <= pos) lo = mid; else hi = mid - 1; }
              return (lo << 5) + (int)__fns(S.bm[lo], 0, pos - S.pre[lo] + 1);
            };
            for (int j0 = 0; j0 < nform; j0 += FT) {
              int j = j0 + tid; bool ok = false; int4 ed = make_int4(0, 0, at, 0);
              if (j < nform) {
                for (int t = 0; t < XGC_FORM_TRIES && !ok; t += 2) {
                  uint4 x = frand(key, at, FS_FORM + l, (unsigned)j, (unsigned)(t >> 1));
#pragma unroll
                  for (int q = 0; q < 2 && !ok; ++q) {
                    int a = (int)__umulhi(q ? x.z : x.x, (unsigned)n_alive), b = (int)__umulhi(q ? x.w : x.y, (unsigned)(n_alive - 1));
                    if (b >= a) b++;
                    int sa = slot_of(min(a, b)), sb = slot_of(max(a, b));
                    if (roles_compatible(S.vd[sa], S.vd[sb])) { ok = true; ed.x = sa; ed.y = sb; }
                  }
                }
              }
              unsigned m = __ballot_sync(0xffffffffu, ok);
              if (lane == 0) S.wsum[warp] = __popc(m);
              __syncthreads();
              if (warp == 0) {
                int c = S.wsum[lane], inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                S.wsum[lane] = inc - c;
                if (lane == 31) S.sc[SC_TMP] = inc;
              }
              __syncthreads();
              int dst = ne + S.wsum[warp] + __popc(m & ((1u << lane) - 1u));
              if (ok) { if (dst < g.e_cap[l]) E[dst] = ed; else atomicOr(&S.sc[SC_STATUS], (int)XGC_ST_EDGE_OVERFLOW); }
              ne = min(ne + S.sc[SC_TMP], g.e_cap[l]);
              __syncthreads();
            }
          }
          if (tid == 0) { S.sc[SC_NE0 + l] = ne; if (form) S.row[XGC_K_NEDGES + l] = (unsigned)ne; }
          __syncthreads();
        }
      }
      TICK(PH_NET);
      // =================================================================================================== POST
      if (A.phases & XF_POST) {
        const int n_slots = S.sc[SC_NSLOTS];
        const int ne0 = S.sc[SC_NE0], ne01 = ne0 + S.sc[SC_NE1], ne_all = ne01 + S.sc[SC_NE2];
        const bool last_week = at == A.at1;
        if (tid == 0 && (mask & XGC_M_STITRANS)) {              // weekly random pathway order (stitrans_msm_rand)
          uint4 x0 = frand(key, at, FS_PATH, 0, 0), x1 = frand(key, at, FS_PATH, 1, 0);
          unsigned dr[6] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y };
          int perm[XGC_NPATH];
          for (int k = 0; k < XGC_NPATH; ++k) perm[k] = k;
          for (int i = XGC_NPATH - 1; i >= 1; --i) { int j = (int)__umulhi(dr[XGC_NPATH - 1 - i], (unsigned)(i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
          for (int k = 0; k < XGC_NPATH; ++k) S.sc[SC_PATH + perm[k]] = k;
        }
        // ---------------------------------------------------------------- acts_msm + exposure history + PrEP risk history
        if (mask & (XGC_M_ACTS | XGC_M_PREP)) {
          const float sc_ai = P[XGC_P_ai_acts_scale_mc] * (1.f / 52.f), sc_oi = P[XGC_P_oi_acts_scale_mc] * (1.f / 52.f);
          for (int f0 = 0; f0 < ne_all; f0 += FT) {
            const int f = f0 + tid;
            unsigned nai = 0, noi = 0;
            if (f < ne_all) {
              const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, e = f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
              int4* Ep = g.edge + (size_t)r * g.e_stride + g.e_off[l] + e;
              const int4 ed = *Ep;
              const unsigned da = S.vd[ed.x], db = S.vd[ed.y], ga = S.vg[ed.x], gb = S.vg[ed.y], ha = S.vh[ed.x], hb = S.vh[ed.y];
              int ai = 0, oi = 0, ki = 0, ri = 0;
              if (mask & XGC_M_ACTS) {
                if (!(act_stopped(da, ga) || act_stopped(db, gb))) {
                  const bool compat = roles_compatible(da, db);
                  const uint4 x = frand(key, at, FS_ACTS, (unsigned)f, 0);
                  if (l < 2) {
                    float dur = (float)(at - ed.z);
                    float ca = (float)(2 * at8 - (int)((ha >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK) - (int)((hb >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK)) * (float)XF_W8;
                    bool hcp = (ha & HV_DIAG) && (hb & HV_DIAG);
                    int r1 = (int)DM_RACE(da), r2 = (int)DM_RACE(db);
                    float rai = __expf(glm_f(S.G, dur, l == 1, ca, hcp, false, r1, r2)) * sc_ai;
                    float roi = __expf(glm_f(S.G + XGC_GLM_NCOEF, dur, l == 1, ca, hcp, false, r1, r2)) * sc_oi;
                    ai = compat ? pois_f(u01(x.x), rai, XGC_MAX_ACTS) : 0;
                    oi = pois_f(u01(x.y), roi, XGC_MAX_ACTS);
                    // kissing / rimming: "any act this week" here; the counts are drawn in the transmission phase from the same words
                    ki = u01(x.z) > S.pk[l * (XGC_MAX_ACTS + 1)];
                    ri = u01(x.w) > S.pk[(2 + l) * (XGC_MAX_ACTS + 1)];
                  } else {
                    ai = compat ? 1 : 0;
                    oi = bern(x.y, P[XGC_P_oi_prob_oo]); ki = bern(x.z, P[XGC_P_kiss_prob_oo]); ri = bern(x.w, P[XGC_P_rim_prob_oo]);
                  }
                }
                S.ea[f] = (unsigned short)(ai | (oi << 6) | (ki << 12) | (ri << 13));
                if (last_week && A.write_acts) Ep->w = (int)((unsigned)ai | ((unsigned)oi << 8) | (0x7FFFu << 16));
                nai = ai; noi = oi;
#pragma unroll
                for (int side = 0; side < 2; ++side) {          // exposure history for the CDC screening protocol
                  unsigned role = DM_ROLE(side ? db : da), ex = 0;
                  if (ai > 0 && role != 0) ex |= 1;
                  if ((ai > 0 && role != 1) || oi > 0) ex |= 2;
                  if (oi > 0 || ri > 0) ex |= 4;
                  if (ki > 0) ex |= 8;
                  if (l != 0 && (ai | oi | ki | ri)) ex |= 16;
                  if ((ex << GC_EXPO_SHIFT) & ~(side ? gb : ga)) atomicOr(&S.vg[side ? ed.y : ed.x], ex << GC_EXPO_SHIFT);
                }
              } else ai = S.ea[f] & 63;
              if (mask & XGC_M_PREP) {                          // risk history: condomless anal sex, or anal sex with a diagnosed partner
                bool ua = !(ha & HV_DIAG), ub = !(hb & HV_DIAG);
                if (ai > 0 && (ua || ub)) {
                  bool uai = false;
                  if (!((ua ? (hb & HV_DIAG) : true) && (ub ? (ha & HV_DIAG) : true))) {       // someone's indication still depends on condom use
                    float ca = (float)(2 * at8 - (int)((ha >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK) - (int)((hb >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK)) * (float)XF_W8;
                    bool prep = (ha | hb) & HV_PREP;
                    float eta = l < 2 ? glm_f(S.G + 2 * XGC_GLM_NCOEF, (float)(at - ed.z), l == 1, ca, false, prep, (int)DM_RACE(da), (int)DM_RACE(db))
                                      : glm_f(S.G + 3 * XGC_GLM_NCOEF, 0.f, false, ca, false, prep, (int)DM_RACE(da), (int)DM_RACE(db));
                    float pc = fminf(1.f, P[XGC_P_cond_scale] / (1.f + __expf(-eta)));
                    unsigned thr = (unsigned)((1.f - pc) * 65536.f);
                    for (int k0 = 0; k0 < ai && !uai; k0 += 8) {
                      uint4 c = frand(key, at, FS_COND, (unsigned)f, (unsigned)(k0 >> 3));
                      unsigned w[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
                      for (int k = 0; k < 8; ++k) if (k0 + k < ai && ((w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu) < thr) uai = true;
                    }
                  }
                  if (ua && (uai || (hb & HV_DIAG))) { S.ind[ed.x] = (short)at; g.tck[rbase + ed.x].y = (short)at; }
                  if (ub && (uai || (ha & HV_DIAG))) { S.ind[ed.y] = (short)at; g.tck[rbase + ed.y].y = (short)at; }
                }
              }
            }
            if (mask & XGC_M_ACTS) {
              nai = __reduce_add_sync(0xffffffffu, nai); noi = __reduce_add_sync(0xffffffffu, noi);
              if (lane == 0) { if (nai) atomicAdd(&S.row[XGC_K_NACTS], nai); if (noi) atomicAdd(&S.row[XGC_K_NACTS + 1], noi); }
            }
          }
        }
        __syncthreads();
        TICK(PH_ACTS);
        // ---------------------------------------------------------------- prep_msm, stirecov_msm, stitx_msm
        if (mask & (XGC_M_PREP | XGC_M_STIRECOV | XGC_M_STITX | XGC_M_HIVTRANS)) {
          const float pavd = P[XGC_P_gc_asympt_visit_prob];
          const bool avd_on = (mask & XGC_M_STITX) && (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL);
          for (int base = 0; base < n_slots; base += FAC) {
            const int hi = min(base + FAC, n_slots);
            if (tid < 8) S.qn[tid] = 0u;
            __syncthreads();
            unsigned short* qa = S.Q;                           // region 0: agents with GC work
            unsigned short* qp = S.Q + FQ_REGION;               // region 1: PrEP users and PrEP starters
            if (avd_on) gap_sample(key, at, FS_AVD, base, hi, pavd, 0, 32, [&](int i, const uint4&) { atomicOr(&S.vg[i], FG_AVD); });
            __syncthreads();
            for (int i0 = base; i0 < hi; i0 += FT) {
              int i = i0 + tid; bool qA = false, qP = false;
              if (i < hi && (S.vd[i] & DM_ACTIVE)) {
                unsigned gc = S.vg[i], h = S.vh[i];
                // start-of-pass snapshot: what acts / condoms / hivtrans see (they precede prep / stirecov / stitx in the module order)
                unsigned gc1 = (gc & ~(GC_CUR_MASK << GC_PRE_SHIFT)) | ((gc & GC_CUR_MASK) << GC_PRE_SHIFT);
                unsigned h1 = (h & ~HV_PREP_PRE) | ((h & HV_PREP) ? HV_PREP_PRE : 0u);
                if (mask & XGC_M_PREP) {
                  bool diag = h & HV_DIAG; short il = S.ind[i];
                  bool ind = il != XGC_NA16 && (float)il >= (float)at - P[XGC_P_prep_risk_int];
                  h1 = (h1 & ~HV_PREPELIG) | ((!diag && ind) ? HV_PREPELIG : 0u);
                  // a negative test this week (incl. a false negative inside the window period) is what opens PrEP start / reassessment
                  qP = (h & HV_PREP) || (!diag && ind && (gc & FG_TESTED) && (float)at >= P[XGC_P_prep_start]);
                }
                if (gc1 != gc) S.vg[i] = gc1;
                if (h1 != h) S.vh[i] = h1;
                unsigned ex = (gc >> GC_EXPO_SHIFT) & 0x1Fu;
                bool cdc_active = proto == XGC_SCREEN_CDC && ((ex & 7u) || (kissx && (ex & 8u)));
                qA = (mask & (XGC_M_STIRECOV | XGC_M_STITX)) && ((gc & 7u) || ((mask & XGC_M_STITX) && ((gc & FG_AVD) || cdc_active)));
              }
              qpush(qA, (unsigned)(i - base), qa, &S.qn[0], lane);
              qpush(qP, (unsigned)(i - base), qp, &S.qn[1], lane);
            }
            __syncthreads();
            TICK(PH_B_SCAN);
            {                                                   // prep_msm on the few it concerns
              const unsigned np = min(S.qn[1], (unsigned)FQ_REGION);
              for (unsigned q = tid; q < np; q += FT) {
                int i = base + (int)qp[q]; size_t idx = rbase + i;
                unsigned h = S.vh[i]; const unsigned h0 = h; int race = (int)DM_RACE(S.vd[i]);
                bool diag = h & HV_DIAG, ind = h & HV_PREPELIG, tested_neg = (S.vg[i] & FG_TESTED) && !diag;
                uint4 x = frand(key, at, FS_PREP, (unsigned)i, 0);
                if (h & HV_PREP) {
                  bool stop = false;
                  if (diag) stop = true;
                  else if (bern(x.x, P[XGC_P_prep_discont_rate + race])) stop = true;
                  else if (tested_neg) {
                    short4 tk = g.tck[idx];
                    if ((float)(at - (int)tk.z) >= P[XGC_P_prep_reassess_int]) { if (!ind) stop = true; else g.tck[idx].z = (short)at; }
                  }
                  if (stop) h &= ~(HV_PREP | HV_PREPCLASS_MASK);
                } else if (!diag && tested_neg && ind && bern(x.x, P[XGC_P_prep_start_prob + race])) {
                  unsigned cl = 1 + categorical_f(u01(x.y), P + XGC_P_prep_adhr_dist, 3);
                  h = (h & ~HV_PREPCLASS_MASK) | HV_PREP | (cl << 7);
                  short4* tq = g.tck + idx; tq->z = (short)at; tq->w = (short)at;
                }
                if (h != h0) S.vh[i] = h;
              }
            }
            __syncthreads();
            TICK(PH_B_PREP);
            {                                                   // stirecov_msm + stitx_msm on infected or visiting agents
              const unsigned na = min(S.qn[0], (unsigned)FQ_REGION);
              for (unsigned q0 = 0; q0 < na; q0 += FT) {
                const unsigned q = q0 + tid;
                unsigned visit = 0, svisit = 0, tst[3] = { 0, 0, 0 }, pos[3] = { 0, 0, 0 }, txs[3] = { 0, 0, 0 };
                if (q < na) {
                  int i = base + (int)qa[q]; size_t idx = rbase + i;
                  unsigned gc = S.vg[i]; const unsigned gc0 = gc, h = S.vh[i];
                  const bool avd = gc & FG_AVD;
                  short4 ga = g.gck_a[idx], gb = g.gck_b[idx], gd = make_short4(-1, -1, -1, -1);
                  if (pois_mode && (gc & 7u)) gd = g.gck_d[idx];
                  bool ga_dirty = false, gb_dirty = false, gd_dirty = false, ind_sti = false;
                  if ((mask & XGC_M_STIRECOV) && (gc & 7u)) {
                    uint4 x = frand(key, at, FS_AG2, (unsigned)i, 0);
                    const unsigned xr[3] = { x.x, x.y, x.z };
                    const int txpr[3] = { XGC_P_rgc_tx_recov_pr, XGC_P_ugc_tx_recov_pr, XGC_P_pgc_tx_recov_pr };
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                      if (!(gc & GC_ST(s))) continue;
                      short infT = s == 0 ? ga.x : s == 1 ? ga.y : ga.z, txT = s == 0 ? gb.x : s == 1 ? gb.y : gb.z, dur = s == 0 ? gd.x : s == 1 ? gd.y : gd.z;
                      bool recov = false;
                      if (gc & GC_TX(s)) { int k = at - (int)txT; if (k >= 1) recov = bern(xr[s], P[txpr[s] + (k >= 3 ? 2 : k - 1)]); }
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
                    bool sv = false, av = false;
                    if ((gc & 7u) & ((gc >> 3) & 7u) & ~((gc >> 6) & 7u)) {       // a symptomatic, untreated site: care seeking
                      uint4 x = frand(key, at, FS_AG2, (unsigned)i, 1);
                      const unsigned xs[3] = { x.x, x.y, x.z };
#pragma unroll
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
                        bool active = (ex & 7u) || (kissx && (ex & 8u)), hr = (h & HV_PREP) || (ex & 16u);
                        float interval = (hr ? P[XGC_P_cdc_sti_hr_int] : P[XGC_P_cdc_sti_int]) * (52.f / 12.f);
                        av = active && (ga.w == XGC_NA16 || (float)(at - (int)ga.w) >= interval);
                      }
                    }
                    if (sv || av) {
                      visit = 1; svisit = sv;
                      uint4 x = frand(key, at, FS_AG2, (unsigned)i, 2);
                      const unsigned xt[3] = { x.x, x.y, x.z };
#pragma unroll
                      for (int s = 0; s < 3; ++s) {
                        bool inf = gc & GC_ST(s), symp = inf && (gc & GC_SY(s)), tested;
                        bool exposed = (ex >> s) & 1u; if (s == 2 && kissx && (ex & 8u)) exposed = true;
                        if (proto == XGC_SCREEN_CDC && !symp && !exposed) tested = false;
                        else if (proto == XGC_SCREEN_UNIVERSAL) tested = true;
                        else tested = bern(xt[s], symp ? P[XGC_P_gc_sympt_prob_test + s] : P[XGC_P_gc_asympt_prob_test + s]);
                        if (!tested) continue;
                        tst[s] = 1;
                        if (inf) {
                          pos[s] = 1; ind_sti = true;
                          if (!(gc & GC_TX(s))) { gc |= GC_TX(s); txs[s] = 1; gb_dirty = true; if (s == 0) gb.x = (short)at; else if (s == 1) gb.y = (short)at; else gb.z = (short)at; }
                        }
                      }
                      ga.w = (short)at; ga_dirty = true;
                      gc &= ~GC_EXPO_MASK;
                    }
                  }
                  if (ind_sti) { S.ind[i] = (short)at; g.tck[idx].y = (short)at; }
                  if (ga_dirty) g.gck_a[idx] = ga;
                  if (gb_dirty) g.gck_b[idx] = gb;
                  if (gd_dirty) g.gck_d[idx] = gd;
                  // other threads OR exposure / new-infection bits into this word only in other phases: a plain store is safe here
                  if (gc != gc0) S.vg[i] = gc;
                }
                if (mask & XGC_M_STITX) {
                  if (__ballot_sync(0xffffffffu, visit != 0)) {
                    row_add(S.row, XGC_K_VISITS, visit); row_add(S.row, XGC_K_VISITS_SYMPT, svisit);
#pragma unroll
                    for (int s = 0; s < 3; ++s) { row_add(S.row, XGC_K_TESTED + s, tst[s]); row_add(S.row, XGC_K_POS + s, pos[s]); row_add(S.row, XGC_K_TXSTART + s, txs[s]); }
                  }
                }
              }
            }
            __syncthreads();
          }
        }
        TICK(PH_B_GC);
        // ---------------------------------------------------------------- hivtrans_msm + stitrans_msm_rand over discordant ties
        if (mask & (XGC_M_HIVTRANS | XGC_M_STITRANS)) {
          const bool gct = mask & XGC_M_STITRANS, hvt = mask & XGC_M_HIVTRANS;
          for (int fb = 0; fb < ne_all; fb += FEC) {
            const int fhi = min(fb + FEC, ne_all);
            if (tid < 8) S.qn[tid] = 0u;
            __syncthreads();
            for (int f0 = fb; f0 < fhi; f0 += FT) {            // classify: which (tie, act type) pairs can transmit
              const int f = f0 + tid; bool q[4] = { false, false, false, false };
              if (f < fhi) {
                const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, e = f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
                const int4 ed = g.edge[(size_t)r * g.e_stride + g.e_off[l] + e];
                const unsigned ea = S.ea[f]; const int ai = ea & 63, oi = (ea >> 6) & 63;
                const unsigned x = S.vg[ed.x] & 7u, y = S.vg[ed.y] & 7u;
                bool hivdisc = hvt && ai > 0 && ((S.vh[ed.x] ^ S.vh[ed.y]) & HV_STATUS);
                bool d_ai = (((x >> 1) ^ y) | ((y >> 1) ^ x)) & 1u, d_oi = (((x >> 1) ^ (y >> 2)) | ((y >> 1) ^ (x >> 2))) & 1u;
                bool d_ri = (((x >> 2) ^ y) | ((y >> 2) ^ x)) & 1u, d_ki = ((x ^ y) >> 2) & 1u;
                q[0] = hivdisc || (gct && ai > 0 && d_ai);
                q[1] = gct && oi > 0 && d_oi;
                q[2] = gct && ((ea >> 12) & 1u) && route_kiss && d_ki;
                q[3] = gct && ((ea >> 13) & 1u) && route_rim && d_ri;
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) qpush(q[t], (unsigned)(f - fb), S.Q + t * FQ_REGION, &S.qn[t], lane);
            }
            __syncthreads();
            TICK(PH_TRANS_CLASSIFY);
            const unsigned n0 = min(S.qn[0], (unsigned)FQ_REGION), n1 = n0 + min(S.qn[1], (unsigned)FQ_REGION),
                           n2 = n1 + min(S.qn[2], (unsigned)FQ_REGION), n3 = n2 + min(S.qn[3], (unsigned)FQ_REGION);
            for (unsigned j = tid; j < n3; j += FT) {           // the four queues as one index space: warps run one act type
              const int type = j < n0 ? 0 : j < n1 ? 1 : j < n2 ? 2 : 3;
              const unsigned jj = j - (type == 0 ? 0u : type == 1 ? n0 : type == 2 ? n1 : n2);
              const int f = fb + (int)S.Q[type * FQ_REGION + jj];
              const int l = f >= ne01 ? 2 : f >= ne0 ? 1 : 0, e = f - (l == 2 ? ne01 : l == 1 ? ne0 : 0);
              const int4 ed = g.edge[(size_t)r * g.e_stride + g.e_off[l] + e];
              const unsigned da = S.vd[ed.x], db = S.vd[ed.y], gca = S.vg[ed.x], gcb = S.vg[ed.y], ha = S.vh[ed.x], hb = S.vh[ed.y];
              const unsigned ga = gca & 7u, gb = gcb & 7u, ea = S.ea[f];
              unsigned new_a = 0, new_b = 0; bool hiv_a = false, hiv_b = false;
              if (type == 0) {
                const int ai = ea & 63;
                const bool hivdisc = hvt && ((ha ^ hb) & HV_STATUS);
                const bool d_ai = gct && ((((ga >> 1) ^ gb) | ((gb >> 1) ^ ga)) & 1u);
                // condom use: probabilities as condoms_msm saw them (PrEP status before this week's prep_msm)
                float ca = (float)(2 * at8 - (int)((ha >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK) - (int)((hb >> XF_BIRTH_SHIFT) & XF_BIRTH_MASK)) * (float)XF_W8;
                bool prep = (ha | hb) & HV_PREP_PRE;
                float eta = l < 2 ? glm_f(S.G + 2 * XGC_GLM_NCOEF, (float)(at - ed.z), l == 1, ca, (ha & HV_DIAG) && (hb & HV_DIAG), prep, (int)DM_RACE(da), (int)DM_RACE(db))
                                  : glm_f(S.G + 3 * XGC_GLM_NCOEF, 0.f, false, ca, (ha & HV_DIAG) && (hb & HV_DIAG), prep, (int)DM_RACE(da), (int)DM_RACE(db));
                float pc = fminf(1.f, P[XGC_P_cond_scale] / (1.f + __expf(-eta)));
                const unsigned cthr = (unsigned)((1.f - pc) * 65536.f);
                const unsigned r1 = DM_ROLE(da), r2 = DM_ROLE(db);
                const int fixed = (r1 == 0 || r2 == 1) ? 1 : (r1 == 1 || r2 == 0) ? 0 : -1;
                float q1 = (float)(da >> 16), q2 = (float)(db >> 16);
                const float pins = q1 + q2 > 0.f ? q1 / (q1 + q2) : 0.5f;
                const bool a_pos = ha & HV_STATUS;
                const unsigned hp = a_pos ? ha : hb, hn = a_pos ? hb : ha, dn = a_pos ? db : da, gn = a_pos ? gcb : gca;
                float vlf = 0.f; const int nrace = (int)DM_RACE(dn);
                if (hivdisc) vlf = exp2f((g.aux[rbase + (a_pos ? ed.x : ed.y)].vl - 4.5f) * 1.2927817f);      // 2.45^(vl - 4.5)
                const unsigned gpre_n = (gn >> GC_PRE_SHIFT) & 7u;
                uint4 c = make_uint4(0, 0, 0, 0);
                for (int k = 0; k < ai; ++k) {
                  if ((k & 7) == 0) c = frand(key, at, FS_COND, (unsigned)f, (unsigned)(k >> 3));
                  const unsigned cw = (k & 7) < 2 ? c.x : (k & 7) < 4 ? c.y : (k & 7) < 6 ? c.z : c.w;
                  const bool uai = ((cw >> ((k & 1) * 16)) & 0xFFFFu) < cthr;
                  const uint4 x = frand(key, at, FS_ANAL, (unsigned)f, (unsigned)k);
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
                  if (d_ai) {                                    // insertive partner's urethra <-> receptive partner's rectum
                    const unsigned gx = p1ins ? ga : gb, gy = p1ins ? gb : ga;
                    const bool ux = (gx >> 1) & 1u, ry = gy & 1u;
                    if (ux != ry) {
                      float t = ux ? P[XGC_P_u2rgc_tprob] : P[XGC_P_r2ugc_tprob];
                      // the susceptible partner: receptive (p1ins ? b : a) for u2r, insertive for r2u
                      const unsigned dsus = (ux ? !p1ins : p1ins) ? da : db;
                      if (!uai) t *= 1.f - P[XGC_P_sti_cond_eff] * (1.f - P[XGC_P_sti_cond_fail + (int)DM_RACE(dsus)]);
                      if (bern(x.z, t)) {
                        if (ux) { if (p1ins) new_b |= 1u << XGC_PATH_U2R; else new_a |= 1u << XGC_PATH_U2R; }
                        else { if (p1ins) new_a |= 1u << XGC_PATH_R2U; else new_b |= 1u << XGC_PATH_R2U; }
                      }
                    }
                  }
                }
              } else if (type == 2) {                            // kissing: pharynx <-> pharynx over all of the week's acts
                const uint4 xa = frand(key, at, FS_ACTS, (unsigned)f, 0);
                int n = l < 2 ? pois_tab_f(u01(xa.z), S.pk + l * (XGC_MAX_ACTS + 1)) : 1;
                const bool pa = (ga >> 2) & 1u;
                const uint4 x = frand(key, at, FS_KISS, (unsigned)f, 0);
                if (n > 0 && bern(x.x, any_of_n_f(P[XGC_P_p2pgc_tprob], n))) { if (pa) new_b |= 1u << XGC_PATH_P2P; else new_a |= 1u << XGC_PATH_P2P; }
              } else {                                           // oral (urethra <-> pharynx), rimming (pharynx <-> rectum)
                const bool oral = type == 1;
                int n = (ea >> 6) & 63;
                if (!oral) { const uint4 xa = frand(key, at, FS_ACTS, (unsigned)f, 0); n = l < 2 ? pois_tab_f(u01(xa.w), S.pk + (2 + l) * (XGC_MAX_ACTS + 1)) : 1; }
                const int sx = oral ? 1 : 2, sy = oral ? 2 : 0;
                const float tA = oral ? P[XGC_P_u2pgc_tprob] : P[XGC_P_p2rgc_tprob], tB = oral ? P[XGC_P_p2ugc_tprob] : P[XGC_P_r2pgc_tprob];
                const unsigned pA = oral ? XGC_PATH_U2P : XGC_PATH_P2R, pB = oral ? XGC_PATH_P2U : XGC_PATH_R2P;
                const uint4 x = frand(key, at, oral ? FS_ORAL : FS_RIM, (unsigned)f, 0);
                if (n > 0) {
                  const int m1 = __popc(x.x >> (32 - n)), m0 = n - m1;       // acts with p1 / p2 the active partner
                  { bool ix = (ga >> sx) & 1u, iy = (gb >> sy) & 1u;
                    if (m1 > 0 && ix != iy && bern(x.y, any_of_n_f(ix ? tA : tB, m1))) { if (ix) new_b |= 1u << pA; else new_a |= 1u << pB; } }
                  { bool ix = (gb >> sx) & 1u, iy = (ga >> sy) & 1u;
                    if (m0 > 0 && ix != iy && bern(x.z, any_of_n_f(ix ? tA : tB, m0))) { if (ix) new_a |= 1u << pA; else new_b |= 1u << pB; } }
                }
              }
              if (new_a) atomicOr(&S.vg[ed.x], new_a << GC_NEW_SHIFT);
              if (new_b) atomicOr(&S.vg[ed.y], new_b << GC_NEW_SHIFT);
              if (hiv_a) atomicOr(&S.vh[ed.x], HV_NEW);
              if (hiv_b) atomicOr(&S.vh[ed.y], HV_NEW);
            }
            __syncthreads();
          }
        }
        TICK(PH_TRANS);
        // ---------------------------------------------------------------- apply new infections, prevalence_msm
        for (int base = 0; base < n_slots; base += FAC) {
          const int hi = min(base + FAC, n_slots);
          if (tid < 8) S.qn[tid] = 0u;
          __syncthreads();
          for (int i0 = base; i0 < hi; i0 += FT) {
            int i = i0 + tid; bool q = false;
            if (i < hi && (S.vd[i] & DM_ACTIVE)) q = (S.vg[i] & GC_NEW_MASK) || (S.vh[i] & HV_NEW);
            qpush(q, (unsigned)(i - base), S.Q, &S.qn[0], lane);
          }
          __syncthreads();
          const unsigned nq = min(S.qn[0], (unsigned)FQ_REGION);
          for (unsigned q = tid; q < nq; q += FT) {
            int i = base + (int)S.Q[q]; size_t idx = rbase + i;
            unsigned gc = S.vg[i], h = S.vh[i]; const unsigned cell = DM_CELL(S.vd[i]);
            unsigned nm = (gc >> GC_NEW_SHIFT) & 0x7Fu;
            if (nm && (mask & XGC_M_STITRANS)) {
              short4 ga = g.gck_a[idx], gb = g.gck_b[idx], gd = make_short4(-1, -1, -1, -1);
              if (pois_mode) gd = g.gck_d[idx];
              uint4 x = frand(key, at, FS_INF, (unsigned)i, 0);
              const unsigned xs[3] = { x.x, x.y, x.z };
              bool any = false;
#pragma unroll
              for (int s = 0; s < 3; ++s) {
                unsigned cand = nm & (s == 0 ? ((1u << XGC_PATH_U2R) | (1u << XGC_PATH_P2R)) : s == 1 ? ((1u << XGC_PATH_R2U) | (1u << XGC_PATH_P2U))
                                                     : ((1u << XGC_PATH_U2P) | (1u << XGC_PATH_R2P) | (1u << XGC_PATH_P2P)));
                if (!cand || (gc & GC_ST(s))) continue;
                int win = -1, wr = 99;
                for (int k = 0; k < XGC_NPATH; ++k) if ((cand >> k) & 1u) { int rk = S.sc[SC_PATH + k]; if (rk < wr) { wr = rk; win = k; } }
                gc = (gc | GC_ST(s)) & ~(GC_SY(s) | GC_TX(s));
                if (bern(xs[s], P[XGC_P_gc_sympt_prob + s])) gc |= GC_SY(s);
                short du = -1;
                if (pois_mode) { uint4 y = frand(key, at, FS_INF, (unsigned)i, 1 + s); int qd = pois_f(u01(y.x), P[XGC_P_gc_ntx_int + s], 32); du = (short)(qd < 1 ? 1 : qd); }
                if (s == 0) { ga.x = (short)at; gb.x = -1; gd.x = du; } else if (s == 1) { ga.y = (short)at; gb.y = -1; gd.y = du; } else { ga.z = (short)at; gb.z = -1; gd.z = du; }
                atomicAdd(&S.row[(XGC_K_INCID_RGC + s) * XGC_NCELL + cell], 1u);
                atomicAdd(&S.row[XGC_K_PATH + win], 1u);
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
              atomicAdd(&S.row[XGC_K_INCID_HIV * XGC_NCELL + cell], 1u);
            }
            if (mask & XGC_M_HIVTRANS) h &= ~HV_NEW;
            S.vg[i] = gc; S.vh[i] = h;
          }
          __syncthreads();
        }
        TICK(PH_APPLY);
        // prevalence_msm: recount over the staged state; transient week flags are cleared on the way
        for (int i0 = 0; i0 < n_slots; i0 += FT) {
          int i = i0 + tid;
          bool live = i < n_slots && (S.vd[i] & DM_ACTIVE);
          unsigned cell = 31u, bits = 0u;
          if (live) {
            unsigned gc = S.vg[i], h = S.vh[i];
            if (gc & (FG_TESTED | FG_AVD)) S.vg[i] = gc & ~(FG_TESTED | FG_AVD);
            cell = DM_CELL(S.vd[i]);
            bits = ((h & HV_STATUS) ? 1u << XGC_K_INUM : 0u) | ((h & HV_DIAG) ? 1u << XGC_K_INUM_DX : 0u) | (((h & HV_DIAG) && (h & HV_VSUPP)) ? 1u << XGC_K_VSUPP : 0u) |
                   ((h & HV_PREPELIG) ? 1u << XGC_K_PREPELIG : 0u) | ((h & HV_PREP) ? 1u << XGC_K_PREPCURR : 0u) | ((gc & 7u) ? 1u << XGC_K_INUM_GC : 0u) |
                   ((gc & 1u) ? 1u << XGC_K_INUM_RGC : 0u) | ((gc & 2u) ? 1u << XGC_K_INUM_UGC : 0u) | ((gc & 4u) ? 1u << XGC_K_INUM_PGC : 0u);
          }
          if (mask & XGC_M_PREV) {
            unsigned grp = __match_any_sync(0xffffffffu, cell);
            if (live && lane == (unsigned)(__ffs(grp) - 1)) atomicAdd(&S.row[XGC_K_NUM * XGC_NCELL + cell], (unsigned)__popc(grp));
            while (bits) { int b = __ffs(bits) - 1; bits &= bits - 1; atomicAdd(&S.row[b * XGC_NCELL + cell], 1u); }
          }
        }
        __syncthreads();
      }
      TICK(PH_TALLY);
      for (int k = tid; k < XGC_NCOUNTER; k += FT) grow[k] = S.row[k];
      __syncthreads();
    }
    // ------------------------------------------------------------------------------------------ write the replicate back
    {
      const int n_slots = S.sc[SC_NSLOTS];
      unsigned* Abm = g.alive + (size_t)r * (g.n_cap / 32);
      for (int i0 = 0; i0 < g.n_cap; i0 += FT) {
        int i = i0 + tid;
        bool in = i < n_slots;
        unsigned d = in ? S.vd[i] : 0u;
        if (in) { uint4 c = g.core[rbase + i]; c.y = d; c.z = S.vg[i]; c.w = S.vh[i]; g.core[rbase + i] = c; }
        unsigned m = __ballot_sync(0xffffffffu, in && (d & DM_ACTIVE));
        if (lane == 0 && i < g.n_cap) Abm[i >> 5] = m;
      }
      if (tid == 0) {
        rp->n_slots = n_slots; rp->n_slots_pre = S.sc[SC_NPRE]; rp->next_uid = S.sc[SC_NEXTUID]; rp->t_age = S.sc[SC_TAGE];
        rp->ne[0] = S.sc[SC_NE0]; rp->ne[1] = S.sc[SC_NE1]; rp->ne[2] = S.sc[SC_NE2];
        rp->n_alive = S.sc[SC_NALIVE];
        if (S.sc[SC_STATUS]) rp->status |= (unsigned)S.sc[SC_STATUS];
        for (int k = 0; k < XGC_NPATH; ++k) rp->path_rank[k] = S.sc[SC_PATH + k];
        ((int*)g.at_ptr)[r] = A.at1;
      }
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
// phase clocks of k_week_fast (SM cycles summed over CTAs) since the last reset; instrumentation for tools/time_modes.py
extern "C" int xgc_debug_fast_profile(unsigned long long* out, int cap, int reset) {
  unsigned long long h[PH_N];
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(h, g_fast_prof, sizeof h) != cudaSuccess) return -1;
  for (int k = 0; k < PH_N && k < cap; ++k) out[k] = h[k];
  if (reset) { unsigned long long z[PH_N] = { 0 }; cudaMemcpyToSymbol(g_fast_prof, z, sizeof z); }
  return PH_N;
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
  k_week_fast<<<grid, FT, smem, s>>>(g, a);
  return cudaGetLastError();
}
