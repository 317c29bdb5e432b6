/* xgc C ABI — the drop-in boundary of the B200-native weekly epidemic step.
 *
 * What this replaces.  In the reference every weekly module is an R closure `FUN(dat, at) -> dat` handed to
 * control_msm() (/root/reference/burnin/cal/sim1_02_lhs.r:195-218; inst/analysis02_screening/02.00_epi.r:289-313)
 * and run once per week by EpiModel::netsim (sim1_02_lhs.r:232).  The modules themselves live in the un-vendored
 * EpiModelHIV fork (renv.lock:26-31).  An R-side maintainer binds this header with `.Call` shims
 * (see INTEGRATION.md); tests and benchmarks bind it with Python ctypes.  No torch / R types cross this boundary:
 * plain pointers, sizes and status codes only.
 *
 * Conventions: node ids are 1-based R positions (rank among living agents, in attribute-vector order);
 * edge lists are column-major int32 [E,2] as tergmLite stores them; races 1..4 = B,H,O,W; sites ordered rect,ureth,phar;
 * `at` starts at 2 (at = 1 is initialisation).  Every function returns 0 on success, non-zero on error, and never
 * throws or exits; xgc_last_error() gives the message.  The caller owns every buffer it passes; the library copies.
 */
#ifndef XGC_ABI_H
#define XGC_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XGC_ABI_VERSION 1
#ifdef __CUDACC__
#define XGC_HD __host__ __device__
#else
#define XGC_HD
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * Parameters: one double vector per replicate (every replicate of a calibration sweep carries its own LHS draw,
 * sim1_02_lhs.r:34-43).  Index enum generated from params.def.
 * ---------------------------------------------------------------------------------------------------------------- */
enum xgc_param_index {
#define XGC_PARAM(name, count, dflt) XGC_P_##name, XGC_P_##name##_END = XGC_P_##name + (count) - 1,
#include "params.def"
#undef XGC_PARAM
  XGC_NPARAM
};

/* Tables shared by all replicates of a handle (contents of netstats / epistats in the reference: sim1_02_lhs.r:60-61).
 * Flat double array; offsets below. */
#define XGC_ASMR_AGES 101           /* integer ages 0..100                                                          */
#define XGC_GLM_NCOEF 26            /* b0,dur,dur2,pt2,dur_pt2,ca,ca2,hcp,prep, pad, rc[4][4]                        */
enum xgc_glm_coef { XGC_G_B0 = 0, XGC_G_DUR, XGC_G_DUR2, XGC_G_PT2, XGC_G_DUR_PT2, XGC_G_CA, XGC_G_CA2, XGC_G_HCP,
                    XGC_G_PREP, XGC_G_PAD, XGC_G_RC = 10 };
enum xgc_table_offset {
  XGC_T_ASMR      = 0,                                   /* [4][XGC_ASMR_AGES] weekly death prob by race, floor(age) */
  XGC_T_GLM_AI    = XGC_T_ASMR + 4 * XGC_ASMR_AGES,      /* anal acts/yr, main+casual: log link                     */
  XGC_T_GLM_OI    = XGC_T_GLM_AI + XGC_GLM_NCOEF,        /* oral acts/yr, main+casual: log link                     */
  XGC_T_GLM_CONDMC= XGC_T_GLM_OI + XGC_GLM_NCOEF,        /* condom prob, main+casual: logit link                    */
  XGC_T_GLM_CONDOO= XGC_T_GLM_CONDMC + XGC_GLM_NCOEF,    /* condom prob, one-off: logit link                        */
  XGC_T_RACE_PROP = XGC_T_GLM_CONDOO + XGC_GLM_NCOEF,    /* [4] race distribution of arrivals                       */
  XGC_T_ROLE_PROB = XGC_T_RACE_PROP + 4,                 /* [4][3] role class I,R,V by race                         */
  XGC_T_NET_DEG   = XGC_T_ROLE_PROB + 12,                /* [3] surrogate network: target mean degree by layer      */
  XGC_T_NET_DUR   = XGC_T_NET_DEG + 3,                   /* [3] surrogate network: mean duration (weeks), [2] unused*/
  /* surrogate formation, degree- and risk-conditioned (what the reference's formation models condition on: deg.main /
   * deg.casl cross terms and concurrency in the main and casual models, risk.grp in the one-off model; sim1_02_lhs.r:209,226):
   * a proposed tie (a, b) of layer l is accepted with probability w(a) * w(b),
   *   w(i) = NET_DEGW[l][min(deg.main(i), 2)][min(deg.casl(i), 2)] * (l == 2 ? NET_RISKW[risk.grp(i) - 1] : 1),
   * degrees counted in the network the week starts with (after the departures, before this week's dissolution: formation and
 * dissolution both condition on the previous network, as in a separable model).  All ones = uniform formation. */
  XGC_T_NET_DEGW  = XGC_T_NET_DUR + 3,                   /* [3][3][3] in [0, 1]                                      */
  XGC_T_NET_RISKW = XGC_T_NET_DEGW + 27,                 /* [5] in [0, 1]                                            */
  XGC_NTABLE      = XGC_T_NET_RISKW + 5
};

/* Control flags, one int vector per replicate (scenario arms differ per replicate: 02.00_epi.r:315-326). */
enum xgc_control_index {
  XGC_C_transRoute_Kissing = 0,     /* sim1_02_lhs.r:220 */
  XGC_C_transRoute_Rimming,         /* :221 */
  XGC_C_gcUntreatedRecovDist,       /* :222   0 = "geom", 1 = "pois" (analysis01_epi/03_sens_poisdurat.r:225) */
  XGC_C_stiScreeningProtocol,       /* :223   0 base, 1 symptomatic, 2 cdc, 3 universal (02.00_epi.r:318) */
  XGC_C_cdcExposureSite_Kissing,    /* :225 */
  XGC_C_network_mode,               /* 0 = host supplies edge lists each week (simnet_msm stays in R); 1 = on-device surrogate */
  XGC_NCONTROL = 8
};
enum { XGC_RECOV_GEOM = 0, XGC_RECOV_POIS = 1 };
enum { XGC_SCREEN_BASE = 0, XGC_SCREEN_SYMPTOMATIC = 1, XGC_SCREEN_CDC = 2, XGC_SCREEN_UNIVERSAL = 3 };

/* Module mask bits, in the order the reference names the modules (sim1_02_lhs.r:201-218).  xgc_step() runs the
 * selected modules in exactly this order. */
enum xgc_module_bit {
  XGC_M_AGING       = 1u << 0,
  XGC_M_DEPARTURE   = 1u << 1,
  XGC_M_ARRIVAL     = 1u << 2,
  XGC_M_HIVTEST     = 1u << 3,
  XGC_M_HIVTX       = 1u << 4,
  XGC_M_HIVPROGRESS = 1u << 5,
  XGC_M_HIVVL       = 1u << 6,
  XGC_M_RESIM_NETS  = 1u << 7,      /* surrogate only; with network_mode 0 the host calls xgc_set_edgelist instead   */
  XGC_M_ACTS        = 1u << 8,
  XGC_M_CONDOMS     = 1u << 9,      /* act-level condom use: evaluated lazily per act, see DESIGN.md                 */
  XGC_M_POSITION    = 1u << 10,     /* act-level position/site pairing: evaluated lazily per act                     */
  XGC_M_PREP        = 1u << 11,     /* includes risk-history update                                                  */
  XGC_M_HIVTRANS    = 1u << 12,
  XGC_M_STIRECOV    = 1u << 13,
  XGC_M_STITX       = 1u << 14,
  XGC_M_STITRANS    = 1u << 15,
  XGC_M_PREV        = 1u << 16,
  XGC_M_ALL         = (1u << 17) - 1
};

/* ------------------------------------------------------------------------------------------------------------------
 * Random draws.  Every stochastic decision consumes an *indexed* uniform u(replicate, at, stream, entity, k):
 *   native   : Philox4x32-10, key = (seed, global replicate id), counter = (at, stream<<16 | k>>2, e0, e1), lane k&3,
 *              u = (x + 0.5) * 2^-32;
 *   injected : the caller supplies the table (xgc_inject), index = off(stream) + entity_index * K(stream) + k.
 * Entity kinds: agent (e0 = uid, e1 = 0; index = uid), edge (e0,e1 = uids of the endpoints as stored; index =
 * position in the layer's edge list), replicate (0,0; index 0), formed edge (e0 = j, e1 = 0; index = j).
 * ---------------------------------------------------------------------------------------------------------------- */
#define XGC_MAX_ACTS 32             /* per edge, per act type, per week (Poisson draws are truncated here)          */
#define XGC_FORM_TRIES 8
enum xgc_stream {
  /* agent streams */
  XGC_S_AGENT1 = 0,    /* aging..hivvl pass: k0 natural mortality, k1 AIDS mortality, k2 HIV interval test,
                          k3 HIV treatment transition (initiation | halting | re-initiation: mutually exclusive)            */
  XGC_S_AGENT2,        /* prep/stirecov/stitx pass: k0 asymptomatic clinic visit, k1..3 recovery r,u,p (treated |
                          untreated: exclusive per site); k4 PrEP start | discontinuation, k5 PrEP adherence class;
                          k8..10 symptomatic care seeking r,u,p; k12..14 site tested r,u,p                                    */
  XGC_S_INFECT,        /* k0..2 symptomatic r,u,p; k3..5 Poisson duration r,u,p                                     */
  XGC_S_ARRIVE_ATTR,   /* k0 race, k1 role, k2 ins.quot, k3 circ, k4 late.tester, k5 tt.traj, k6 act.stopper, k7 risk.grp */
  /* replicate streams */
  XGC_S_ARRIVE_N,      /* k = Poisson chunk index (mean split into chunks of <= 32)                                 */
  XGC_S_PATHORDER,     /* k0..5 Fisher-Yates draws for the weekly pathway permutation (stitrans_msm_rand)           */
  XGC_S_NET_N,         /* k = layer: number of edges to form                                                        */
  /* formed-edge streams (surrogate network), stream id + layer */
  XGC_S_FORM,          /* XGC_S_FORM + layer; k = 2*try, 2*try+1                                                    */
  /* edge streams: stream id + XGC_EDGE_STREAMS * layer */
  XGC_S_EDGE0 = XGC_S_FORM + 3,
  XGC_S_DISSOLVE = XGC_S_EDGE0, /* k0 (surrogate network)                                                           */
  XGC_S_ACTS,          /* k0 anal, k1 oral, k2 kissing, k3 rimming counts                                           */
  XGC_S_ANAL,          /* k = 4*act + {0 condom, 1 position (versatile pairs), 2 HIV transmission, 3 GC transmission:
                          insertive->receptive (u2r) or receptive->insertive (r2u), at most one is possible per act}:
                          one Philox block per anal act                                                               */
  XGC_S_ORAL,          /* k0 = direction word: act j is "p1 insertive" iff bit (31-j) of floor(u * 2^32) is set;
                          k1 = GC transmission over ALL of the week's acts with p1 insertive, k2 = with p2 insertive
                          (u2p or p2u, exclusive per direction): n iid per-act Bernoulli(t) trials whose only observable
                          is "any success" are drawn as ONE Bernoulli(1 - (1-t)^n)                     */
  XGC_S_RIM,           /* same layout: k0 direction word (p1 is the rimmer), k1 / k2: p2r or r2p                     */
  XGC_S_KISS,          /* k0: p2p from whichever partner is infected, over all of the week's kissing acts            */
  XGC_S_EDGE_END
};
#define XGC_EDGE_STREAMS (XGC_S_EDGE_END - XGC_S_EDGE0)
#define XGC_NSTREAM (XGC_S_EDGE0 + 3 * XGC_EDGE_STREAMS)

/* P(at least one of n iid Bernoulli(t) trials succeeds) = 1 - (1-t)^n, the power by binary exponentiation in a fixed
 * operation order (n <= 63) so that the device and the oracle round identically. */
static inline XGC_HD double xgc_any_of_n(double t, int n) {
  double q = 1.0, b = 1.0 - t;
  for (int i = 0; i < 6; ++i) { if (n & 1) q = q * b; b = b * b; n >>= 1; }
  return 1.0 - q;
}

/* Draws per entity for each stream (table width in injected mode). */
static inline XGC_HD int xgc_stream_width(int s) {
  if (s < XGC_S_EDGE0) {
    switch (s) {
      case XGC_S_AGENT1: return 4;  case XGC_S_AGENT2: return 16;
      case XGC_S_INFECT: return 6;  case XGC_S_ARRIVE_ATTR: return 8;
      case XGC_S_ARRIVE_N: return 64; case XGC_S_PATHORDER: return 6; case XGC_S_NET_N: return 3;
      default: return 3 * XGC_FORM_TRIES;      /* XGC_S_FORM + layer: two endpoint positions + the acceptance draw per try */
    }
  }
  s = (s - XGC_S_EDGE0) % XGC_EDGE_STREAMS + XGC_S_EDGE0;
  if (s == XGC_S_DISSOLVE) return 1;
  if (s == XGC_S_ACTS) return 4;
  if (s == XGC_S_ANAL) return 4 * XGC_MAX_ACTS;
  /* Oral, rimming and kissing acts.  NATIVE draws: one direction word (draw 0) and one any-of-m transmission trial per direction
   * (draws 1, 2; kissing: draw 0) -- DESIGN.md A13.  INJECTED draws: PER ACT, so that a caller that replays R's runif stream can
   * give every rbinom of every act its own uniform: oral / rimming act j takes draw 1 + 2j (direction: p1 active iff u < 0.5) and
   * draw 2 + 2j (transmission); kissing act j takes draw j. */
  if (s == XGC_S_KISS) return XGC_MAX_ACTS;
  return 1 + 2 * XGC_MAX_ACTS;                     /* XGC_S_ORAL, XGC_S_RIM */
}

/* Injected-draw table for ONE replicate and ONE week.  Layout: for each stream s in increasing order, a block of
 * n_entities(s) * xgc_stream_width(s) doubles, where n_entities is uid_cap for agent streams, 1 for replicate streams,
 * form_cap for formed-edge streams and e_cap[layer] for edge streams. */
typedef struct xgc_inject {
  const double* u;        /* [n_replicates][stride] or NULL for native Philox                                       */
  size_t        stride;   /* doubles per replicate = xgc_inject_size(...)                                           */
  int32_t       uid_cap;
  int32_t       e_cap[3];
  int32_t       form_cap;
} xgc_inject;

static inline int xgc_stream_entities(const xgc_inject* d, int s) {
  if (s <= XGC_S_ARRIVE_ATTR) return d->uid_cap;
  if (s < XGC_S_FORM) return 1;
  if (s < XGC_S_EDGE0) return d->form_cap;
  return d->e_cap[(s - XGC_S_EDGE0) / XGC_EDGE_STREAMS];
}
static inline size_t xgc_inject_offset(const xgc_inject* d, int s) {
  size_t off = 0;
  for (int t = 0; t < s; ++t) off += (size_t)xgc_stream_entities(d, t) * (size_t)xgc_stream_width(t);
  return off;
}
static inline size_t xgc_inject_size(const xgc_inject* d) { return xgc_inject_offset(d, XGC_NSTREAM); }

/* ------------------------------------------------------------------------------------------------------------------
 * Epi counters: raw integer tallies per (replicate, week); every epi series of SURVEY Appendix C is a ratio or sum of
 * these (xgc_get_epi).  Stratified groups hold 20 cells: cell = (race-1)*5 + (age.grp-1).
 * ---------------------------------------------------------------------------------------------------------------- */
#define XGC_NCELL 20
enum xgc_counter_group {           /* each group = XGC_NCELL consecutive counters */
  XGC_K_NUM = 0, XGC_K_INUM, XGC_K_INUM_DX, XGC_K_VSUPP, XGC_K_INCID_HIV, XGC_K_PREPELIG, XGC_K_PREPCURR,
  XGC_K_INUM_GC, XGC_K_INUM_RGC, XGC_K_INUM_UGC, XGC_K_INUM_PGC, XGC_K_INCID_RGC, XGC_K_INCID_UGC, XGC_K_INCID_PGC,
  XGC_K_NGROUP
};
enum xgc_counter_scalar {
  XGC_K_PATH = XGC_K_NGROUP * XGC_NCELL, /* [7] incidence by pathway: u2r, p2r, r2u, p2u, u2p, r2p, p2p               */
  XGC_K_VISITS = XGC_K_PATH + 7,   /* clinic visits (symptomatic + screening)                                       */
  XGC_K_VISITS_SYMPT,
  XGC_K_TESTED,                    /* [3] site tests r,u,p  (num.rgc.test ...)                                       */
  XGC_K_POS = XGC_K_TESTED + 3,    /* [3] positive site tests                                                        */
  XGC_K_TXSTART = XGC_K_POS + 3,   /* [3] treatments started                                                         */
  XGC_K_DEP_GEN = XGC_K_TXSTART + 3,
  XGC_K_DEP_AIDS,
  XGC_K_NNEW,
  XGC_K_HIVTESTS,
  XGC_K_NEDGES,                    /* [3] edges per layer after resim                                               */
  XGC_K_NACTS = XGC_K_NEDGES + 3,  /* [4] total anal, oral acts; [2],[3] reserved (kissing / rimming are evaluated lazily)   */
  XGC_NCOUNTER_RAW = XGC_K_NACTS + 4
};
#define XGC_NCOUNTER ((XGC_NCOUNTER_RAW + 3) & ~3)
/* pathway bit order inside XGC_K_PATH and inside the per-agent new-infection mask */
enum { XGC_PATH_U2R = 0, XGC_PATH_P2R, XGC_PATH_R2U, XGC_PATH_P2U, XGC_PATH_U2P, XGC_PATH_R2P, XGC_PATH_P2P, XGC_NPATH };

/* Targets vector produced by xgc_targets(): tail-window means, NA/NaN -> 0
 * (R/calculate_targets.r:14-190; inst/cal/sim0.0_setup.r:205-308). */
enum xgc_target_index {
  XGC_TG_NUM = 0,
  XGC_TG_I_PREV,                  /* ct_hiv_prev                       */
  XGC_TG_IR100_POP,               /* ct_hiv_incid_per100pop            */
  XGC_TG_HIVDX_RACE,              /* [4] i.prev.dx.inf.<race> = prob.hivdx.<race> */
  XGC_TG_HIVDX_AGE = XGC_TG_HIVDX_RACE + 4,   /* [5] */
  XGC_TG_VLS_RACE  = XGC_TG_HIVDX_AGE + 5,    /* [4] cc.vsupp.<race> */
  XGC_TG_VLS_AGE   = XGC_TG_VLS_RACE + 4,     /* [5] */
  XGC_TG_PREPCOV_RACE = XGC_TG_VLS_AGE + 5,   /* [4] */
  XGC_TG_PROP_TESTED = XGC_TG_PREPCOV_RACE + 4, /* [3] prop.{rect,ureth,phar}.tested */
  XGC_TG_PROB_POS  = XGC_TG_PROP_TESTED + 3,  /* [3] prob.{rGC,uGC,pGC}.tested      */
  XGC_TG_PREV_GC   = XGC_TG_PROB_POS + 3,     /* [4] prev.gc, prev.rgc, prev.ugc, prev.pgc */
  XGC_TG_IR100_GC  = XGC_TG_PREV_GC + 4,      /* [4] ir100.gc, .rgc, .ugc, .pgc     */
  /* [8] prob.agedx.hiv.<race>.age<k>: age distribution among the HIV-diagnosed by race, the validation target
   * vt_age_among_hivdx (R/calculate_targets.r:70-105: B.age1, B.age2, H.age1, H.age5, O.age1, W.age1, W.age4, W.age5) */
  XGC_TG_AGEDX     = XGC_TG_IR100_GC + 4,
  XGC_NTARGET      = XGC_TG_AGEDX + 8
};
/* (race - 1) * 5 + (age group - 1) of the eight XGC_TG_AGEDX columns */
#define XGC_TG_AGEDX_CELLS { 0, 1, 5, 9, 10, 15, 18, 19 }

/* Attribute dtypes for upload/download. */
enum xgc_dtype { XGC_F64 = 0, XGC_I32 = 1 };

/* Status bits in xgc_status(). */
enum { XGC_ST_OK = 0, XGC_ST_AGENT_OVERFLOW = 1, XGC_ST_EDGE_OVERFLOW = 2, XGC_ST_GC_EXTINCT = 4, XGC_ST_BAD_EDGE = 8,
       /* a constant weekly act rate (kiss.rate.main/.casl, rim.rate.main/.casl) is so high that the XGC_MAX_ACTS truncation
        * of its Poisson draw is no longer negligible: P(X > XGC_MAX_ACTS) > XGC_TRUNC_TOL.  The top of the reference's
        * prior box (rates up to 10/week, sim1_01_lhs_params.r:66-71) is far below it (P ~ 1e-9). */
       XGC_ST_RATE_TRUNCATED = 16 };
#define XGC_TRUNC_TOL 1e-6

/* Named constants of this header, so that bindings (ctypes, .Call) need not re-declare the enums. */
#define XGC_CONSTANT_LIST(X) \
  X(XGC_ABI_VERSION) X(XGC_NPARAM) X(XGC_NTABLE) X(XGC_NCONTROL) X(XGC_NCOUNTER) X(XGC_NTARGET) X(XGC_NSTREAM) X(XGC_NCELL) \
  X(XGC_NPATH) X(XGC_MAX_ACTS) X(XGC_FORM_TRIES) X(XGC_ASMR_AGES) X(XGC_GLM_NCOEF) X(XGC_EDGE_STREAMS) \
  X(XGC_G_B0) X(XGC_G_DUR) X(XGC_G_DUR2) X(XGC_G_PT2) X(XGC_G_DUR_PT2) X(XGC_G_CA) X(XGC_G_CA2) X(XGC_G_HCP) X(XGC_G_PREP) X(XGC_G_RC) \
  X(XGC_T_ASMR) X(XGC_T_GLM_AI) X(XGC_T_GLM_OI) X(XGC_T_GLM_CONDMC) X(XGC_T_GLM_CONDOO) X(XGC_T_RACE_PROP) X(XGC_T_ROLE_PROB) \
  X(XGC_T_NET_DEG) X(XGC_T_NET_DUR) X(XGC_T_NET_DEGW) X(XGC_T_NET_RISKW) \
  X(XGC_C_transRoute_Kissing) X(XGC_C_transRoute_Rimming) X(XGC_C_gcUntreatedRecovDist) X(XGC_C_stiScreeningProtocol) \
  X(XGC_C_cdcExposureSite_Kissing) X(XGC_C_network_mode) \
  X(XGC_RECOV_GEOM) X(XGC_RECOV_POIS) X(XGC_SCREEN_BASE) X(XGC_SCREEN_SYMPTOMATIC) X(XGC_SCREEN_CDC) X(XGC_SCREEN_UNIVERSAL) \
  X(XGC_M_AGING) X(XGC_M_DEPARTURE) X(XGC_M_ARRIVAL) X(XGC_M_HIVTEST) X(XGC_M_HIVTX) X(XGC_M_HIVPROGRESS) X(XGC_M_HIVVL) \
  X(XGC_M_RESIM_NETS) X(XGC_M_ACTS) X(XGC_M_CONDOMS) X(XGC_M_POSITION) X(XGC_M_PREP) X(XGC_M_HIVTRANS) X(XGC_M_STIRECOV) \
  X(XGC_M_STITX) X(XGC_M_STITRANS) X(XGC_M_PREV) X(XGC_M_ALL) \
  X(XGC_S_AGENT1) X(XGC_S_AGENT2) X(XGC_S_INFECT) \
  X(XGC_S_ARRIVE_ATTR) X(XGC_S_ARRIVE_N) X(XGC_S_PATHORDER) X(XGC_S_NET_N) X(XGC_S_FORM) X(XGC_S_EDGE0) X(XGC_S_DISSOLVE) \
  X(XGC_S_ACTS) X(XGC_S_ANAL) X(XGC_S_ORAL) X(XGC_S_RIM) X(XGC_S_KISS) \
  X(XGC_K_NUM) X(XGC_K_INUM) X(XGC_K_INUM_DX) X(XGC_K_VSUPP) X(XGC_K_INCID_HIV) X(XGC_K_PREPELIG) X(XGC_K_PREPCURR) \
  X(XGC_K_INUM_GC) X(XGC_K_INUM_RGC) X(XGC_K_INUM_UGC) X(XGC_K_INUM_PGC) X(XGC_K_INCID_RGC) X(XGC_K_INCID_UGC) X(XGC_K_INCID_PGC) \
  X(XGC_K_NGROUP) X(XGC_K_PATH) X(XGC_K_VISITS) X(XGC_K_VISITS_SYMPT) X(XGC_K_TESTED) X(XGC_K_POS) X(XGC_K_TXSTART) \
  X(XGC_K_DEP_GEN) X(XGC_K_DEP_AIDS) X(XGC_K_NNEW) X(XGC_K_HIVTESTS) X(XGC_K_NEDGES) X(XGC_K_NACTS) \
  X(XGC_TG_NUM) X(XGC_TG_I_PREV) X(XGC_TG_IR100_POP) X(XGC_TG_HIVDX_RACE) X(XGC_TG_HIVDX_AGE) X(XGC_TG_VLS_RACE) X(XGC_TG_VLS_AGE) \
  X(XGC_TG_PREPCOV_RACE) X(XGC_TG_PROP_TESTED) X(XGC_TG_PROB_POS) X(XGC_TG_PREV_GC) X(XGC_TG_IR100_GC) X(XGC_TG_AGEDX) \
  X(XGC_ST_AGENT_OVERFLOW) X(XGC_ST_EDGE_OVERFLOW) X(XGC_ST_GC_EXTINCT) X(XGC_ST_BAD_EDGE) X(XGC_ST_RATE_TRUNCATED) X(XGC_MODE_STRICT) X(XGC_MODE_FAST) X(XGC_F64) X(XGC_I32)
int  xgc_constant(const char* name, int64_t* value);          /* 0 if found */
int  xgc_param_index(const char* name, int* offset, int* count);   /* params.def lookup */
int  xgc_inject_stream_width(int stream);                       /* = xgc_stream_width() for bindings that cannot include this header */
size_t xgc_inject_table_size(const xgc_inject* d);               /* = xgc_inject_size()                                                */

typedef struct xgc_config {
  int32_t  abi_version;        /* XGC_ABI_VERSION                                                                   */
  int32_t  device;             /* CUDA device ordinal                                                               */
  int32_t  n_replicates;       /* replicates resident on this handle (this GPU's shard)                             */
  int32_t  replicate_offset;   /* global id of local replicate 0 (Philox key), so results do not depend on sharding */
  int32_t  n_cap;              /* agent slots per replicate (>= initial N + head-room; multiple of 128)             */
  int32_t  e_cap[3];           /* edge capacity per layer: main, casual, one-off                                    */
  int32_t  t_cap;              /* weeks of epi counters kept (at < t_cap)                                           */
  int32_t  compact_every;      /* weeks between slot compactions inside xgc_run (0 = default 64)                    */
  uint64_t seed;
} xgc_config;

typedef struct xgc_handle xgc_handle;

/* life cycle -------------------------------------------------------------------------------------------------------- */
int  xgc_create(const xgc_config* cfg, xgc_handle** out);
int  xgc_destroy(xgc_handle* h);
const char* xgc_last_error(const xgc_handle* h);      /* h may be NULL: error of a failed xgc_create */
int  xgc_abi_version(void);

/* configuration: replaces param_msm()/control_msm() objects (sim1_02_lhs.r:58-230).  r = -1 applies to all replicates. */
int  xgc_set_params (xgc_handle* h, int r, const double* params /*[XGC_NPARAM]*/);
int  xgc_get_params (xgc_handle* h, int r, double* params);
int  xgc_set_params_all(xgc_handle* h, const double* params /*[n_replicates][XGC_NPARAM]*/);
int  xgc_set_control(xgc_handle* h, int r, const int32_t* control /*[XGC_NCONTROL]*/);
int  xgc_set_tables (xgc_handle* h, const double* tables /*[XGC_NTABLE]*/);
/* read back what a handle runs with (a scenario arm forked from a burn-in continues with the burn-in's own netstats /
 * epistats tables, parameters and flags: 02.00_epi.r:292-296 reuses the burn-in's param / control objects) */
int  xgc_get_tables (xgc_handle* h, double* tables /*[XGC_NTABLE]*/);
int  xgc_get_params_all(xgc_handle* h, double* params /*[n_replicates][XGC_NPARAM]*/);
int  xgc_get_control(xgc_handle* h, int r, int32_t* control /*[XGC_NCONTROL]*/);

/* state exchange: replaces dat$attr / dat$el.  Attribute vectors are in R position order.
 * xgc_set_population sets N for replicate r, marks slots 0..n-1 alive, assigns uid = 0..n-1 and resets all attributes
 * to their "new susceptible agent" values; at is the week the attributes refer to (1 after initialisation). */
int  xgc_set_population(xgc_handle* h, int r, int n, int at);
int  xgc_num_agents   (xgc_handle* h, int r, int* n);
int  xgc_upload_attr  (xgc_handle* h, int r, const char* name, const void* data, size_t n, int dtype);
int  xgc_download_attr(xgc_handle* h, int r, const char* name, void* data, size_t cap, int dtype, size_t* n_out);
int  xgc_attr_names   (const char** names, int cap);           /* returns the number of known attribute names */
int  xgc_set_edgelist (xgc_handle* h, int r, int layer, const int32_t* el /*[E,2] col-major, 1-based*/, int n_edges,
                       const int32_t* start_time /*[E] or NULL -> current week*/);
int  xgc_get_edgelist (xgc_handle* h, int r, int layer, int32_t* el, int cap, int* n_edges, int32_t* start_time);
/* dat$temp$el act counts of the current week; defined from acts_msm to the end of that week (kissing / rimming counts are
 * evaluated lazily on the hot path and materialised by this call). */
int  xgc_get_actcounts(xgc_handle* h, int r, int layer, int32_t* counts /*[E,4] col-major: ai,oi,kiss,rim*/, int cap,
                       int* n_edges);
/* materialise the act list of week `at` (dat$temp$al): rows (layer, edge index, act type 0 ai/1 oi/2 kiss/3 rim,
 * act index, uai (1 = condomless; anal only), p1_active (1 = p1 insertive / rimmer)).  Column-major int32 [A,6]. */
int  xgc_get_actlist  (xgc_handle* h, int r, int at, int32_t* al, int cap, int* n_acts, const xgc_inject* inj);

/* Arithmetic / sampling contract of NATIVE-draw runs (north_star: bit-exact with injected draws, statistically
 * indistinguishable targets with native RNG).
 *   XGC_MODE_STRICT (default): one indexed Philox draw per decision, fp64 in a fixed operation order -- bit-identical to
 *                              the CPU oracle also with native draws; the only mode injected draws ever use.
 *   XGC_MODE_FAST            : replicate-resident kernel (one CTA per replicate, agent state staged in shared memory for
 *                              up to compact_every weeks), gap-sampled rare events, fp32 probabilities.  Same modules,
 *                              same state machine, same outputs; trajectories differ from STRICT draw by draw and are
 *                              equal in distribution.  xgc_step uses it for the module groups a host-network loop calls
 *                              (all modules; aging..hivvl [+ resim_nets]; [resim_nets +] acts..prev) and falls back to
 *                              the strict kernels for any other module mask and whenever draws are injected.
 * xgc_set_mode fails (non-zero) when a replicate of this handle does not fit one SM's shared memory. */
enum { XGC_MODE_STRICT = 0, XGC_MODE_FAST = 1 };
int  xgc_set_mode(xgc_handle* h, int mode);
int  xgc_get_mode(xgc_handle* h, int* mode);

/* Global ids of the handle's replicates (default replicate_offset + r).  Replicate r draws from Philox key (seed, ids[r]), so any
 * subset of a sweep can be packed into a handle -- the reference's job array runs replicates as unrelated tasks, sim1_batch1.sh:6-9 --
 * e.g. replicates of similar cost per week (xgc_replicate_cost: SM cycles of the resident kernel per replicate since the last reset). */
int  xgc_set_replicate_ids(xgc_handle* h, const int32_t* ids /*[R]*/);
int  xgc_replicate_cost(xgc_handle* h, uint64_t* cycles /*[R] or NULL*/, int reset);

/* the step: runs the modules selected by module_mask for week `at` on every replicate, in reference order.
 * inj = NULL -> native Philox; otherwise draws come from the injected table. */
int  xgc_step(xgc_handle* h, int at, uint32_t module_mask, const xgc_inject* inj);
/* the modules of mask_tail for week `at`, then the modules of mask_head for week at + 1, native draws: the call the LAST module
 * wrapper of a week makes in a host-network loop (netsim runs prevalence_msm(at) and aging_msm(at + 1) back to back, nothing on
 * the host in between -- EpiModel::netsim's `for (at in 2:nsteps)` loop, sim1_02_lhs.r:201-218 module order).  Same result as
 * xgc_step(at, mask_tail) + xgc_step(at + 1, mask_head); in XGC_MODE_FAST, with mask_tail = the acts..prevalence group (with or
 * without resim_nets) and mask_head = the aging..hivvl group, the two run in ONE launch of the resident kernel. */
int  xgc_step_span(xgc_handle* h, int at, uint32_t mask_tail, uint32_t mask_head);
/* One week of the host-network loop in one call: xgc_update_edgelists_all (main, casual: weekly toggles) + xgc_set_edgelists_all
 * (one-off list; any of the three may be NULL = unchanged) -> xgc_step_span(at, mask_tail, mask_head) (mask_head = 0: xgc_step(at,
 * mask_tail)) -> xgc_get_counters_all(at) and xgc_num_agents_all into the caller's buffers (NULL = not wanted), one wait.  What a
 * resim_nets.FUN wrapper calls after the reference's simnet_msm (sim1_02_lhs.r:209) has produced the week's network changes. */
int  xgc_week_hostnet(xgc_handle* h, int at, uint32_t mask_tail, uint32_t mask_head,
                      const int32_t* delta_main, int stride_main, const int32_t* delta_casl, int stride_casl, int has_start,
                      const int32_t* el_inst, const int32_t* ne_inst, int stride_inst,
                      uint32_t* counters_out, int32_t* n_alive_out);
/* batched loop at0..at1 inclusive, all modules, native draws, no host synchronisation inside (CUDA-graph replay). */
int  xgc_run (xgc_handle* h, int at0, int at1);
int  xgc_sync(xgc_handle* h);
int  xgc_compact(xgc_handle* h);                          /* squeeze out slots of departed agents (positions kept) */
int  xgc_status (xgc_handle* h, int r, uint32_t* status); /* XGC_ST_* bits */

/* outputs: replaces dat$epi and the calculate_targets family (R/calculate_targets.r:14-190). */
int  xgc_get_counters(xgc_handle* h, int r, int at0, int at1, uint32_t* out /*[at1-at0+1][XGC_NCOUNTER]*/);
int  xgc_get_epi     (xgc_handle* h, int r, const char* var, double* out, int at0, int at1);
int  xgc_epi_names   (const char** names, int cap);
/* Many series of one replicate with ONE device->host read of its weekly tallies: out[k * (at1 - at0 + 1) + (at - at0)] =
 * series vars[k] at week `at`.  This is the `as.data.table(sim$epi)` table that inst/cal/sim_clean.r:55-64 builds per
 * simulation file (columns = epi names, rows = weeks) before it stacks them with a `simid` column. */
int  xgc_get_epi_table(xgc_handle* h, int r, const char* const* vars, int n_vars, double* out, int at0, int at1);
int  xgc_targets     (xgc_handle* h, int at_last, int tailn, double* out /*[n_replicates][XGC_NTARGET]*/);

/* checkpoint / scenario fork (02.00_epi.r:40-76,292,296: burn-in output reused as the start of 5 screening arms). */
int  xgc_snapshot_size(xgc_handle* h, size_t* bytes);
int  xgc_snapshot(xgc_handle* h, void* buf, size_t bytes);
int  xgc_restore (xgc_handle* h, const void* buf, size_t bytes);
int  xgc_fork    (xgc_handle* src, int src_r, xgc_handle* dst, int dst_r);   /* device-to-device copy of one replicate */

/* e2e helpers with caller-owned pinned host buffers: one call moves every replicate's edge lists in, or the week's
 * counters out (the two transfers the weekly R<->device contract needs; DESIGN.md "boundary cost").
 * xgc_set_edgelists_all is asynchronous on the handle's stream: with page-locked buffers the caller must leave `el`,
 * `n_edges` and `start_time` unchanged until the next synchronising call on this handle (xgc_sync, any xgc_get_* /
 * xgc_num_agents_all); pageable buffers are staged before the call returns.  Positions outside 1..living agents set
 * XGC_ST_BAD_EDGE in that replicate's status word (the edge is replaced by (1,1)); lists longer than the layer's
 * capacity are truncated and set XGC_ST_EDGE_OVERFLOW. */
int  xgc_set_edgelists_all(xgc_handle* h, int layer, const int32_t* el /*[R][2][stride] 1-based positions*/,
                           const int32_t* n_edges /*[R]*/, const int32_t* start_time /*[R][stride] or NULL*/, int stride);
/* Weekly DELTA of a persistent layer (main / casual ties live for years; tergmLite's simnet_msm toggles a few dozen of them a
 * week), instead of re-sending the whole list: one packed int32 row per replicate,
 *   row = { n_removed, n_added, removed[n_removed], tail[n_added], head[n_added] (, start[n_added] if has_start) }
 * removed[] are ascending 0-based indices into the layer's CURRENT edge list -- the list xgc_get_edgelist would return now,
 * i.e. after this week's departures were pruned, in the stable order both sides keep; tail/head are 1-based positions among
 * the living.  The surviving ties keep their order, the added ones are appended in the order given (start = has_start ? the
 * given week : the current week).  Same asynchrony and error flags as xgc_set_edgelists_all (a removed index outside the
 * list sets XGC_ST_BAD_EDGE).
 * Environment, read by xgc_create: XGC_FUSE_HOSTNET=1 -- in XGC_MODE_FAST both calls only copy and QUEUE their operation, and the next
 * acts..prevalence launch of the resident kernel applies it itself (same lists, bit for bit; the flags then appear with that step or
 * with the next xgc_status); worth it when the handle has the GPU to itself (one wave or less), not when several handles share it.
 * XGC_LANES=1 -- the strict xgc_run path drives the replicates as one lane instead of two concurrent ones. */
int  xgc_update_edgelists_all(xgc_handle* h, int layer, const int32_t* delta /*[R][stride]*/, int stride, int has_start);
int  xgc_get_counters_all (xgc_handle* h, int at, uint32_t* out /*[R][XGC_NCOUNTER]*/);
int  xgc_num_agents_all   (xgc_handle* h, int32_t* out /*[R]*/);   /* living agents per replicate (what simnet_msm needs) */

/* instrumentation */
int  xgc_kernel_launches(xgc_handle* h, uint64_t* n);     /* kernels launched by this handle so far */
int  xgc_profile(xgc_handle* h, int enable);              /* bracket every kernel launch of xgc_step with CUDA events */
int  xgc_profile_read(xgc_handle* h, const char** names, double* ms_total, uint64_t* launches, int cap); /* returns #kernels */
int  xgc_stream_ptr(xgc_handle* h, void** cuda_stream);   /* the stream all work is issued on (for event timing) */
int  xgc_debug_exp(const double* x, double* y, int n, int device);   /* device xgc_exp, for the math parity test */
int  xgc_debug_philox(uint32_t seed, uint32_t rep, const uint32_t* ctr4, uint32_t* out4, int n, int device);

#ifdef __cplusplus
}
#endif
#endif /* XGC_ABI_H */
