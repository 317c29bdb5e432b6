/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See xgc_oracle.h for the provenance statement ("parity unpinned").
 *
 * Plain-C, one-replicate, R-style restatement of the weekly modules the reference names at
 * /root/reference/burnin/cal/sim1_02_lhs.r:200-218.  Each function cites the reference line that names/parameterises
 * the module it restates and the SURVEY.md Appendix E paragraph it follows.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: every probability must be bit-identical to the CUDA path,
 * which is compiled with --fmad=false).
 */
#include "xgc_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------------ */
/* attributes (dat$attr): all numeric vectors, like R                                                                 */
/* ------------------------------------------------------------------------------------------------------------------ */
enum {
  A_UID, A_AGE, A_AGEREF, A_RACE, A_AGEGRP, A_ROLE, A_INSQ, A_RISK, A_CIRC, A_LATE, A_TTTRAJ, A_STOPPER, A_ARRT,
  A_GC, A_GC_INF = A_GC + 3, A_GC_SYM = A_GC_INF + 3, A_GC_TX = A_GC_SYM + 3, A_GC_TXT = A_GC_TX + 3,
  A_GC_DUR = A_GC_TXT + 3, A_LASTSCR = A_GC_DUR + 3, A_EXPO,
  A_STATUS, A_INFT, A_STAGE, A_STAGET, A_DIAG, A_DIAGT, A_LNT, A_TXS, A_CUMON, A_CUMOFF, A_TXINIT, A_VL,
  A_PREPSTAT, A_PREPCLASS, A_PREPELIG, A_PREPSTART, A_PREPLR, A_INDLAST,
  A_NATTR
};
static const char* const ATTR_NAME[A_NATTR] = {
  "uid", "age", "age.ref", "race", "age.grp", "role.class", "ins.quot", "risk.grp", "circ", "late.tester", "tt.traj",
  "act.stopper", "arrival.time",
  "rGC", "uGC", "pGC", "rGC.infTime", "uGC.infTime", "pGC.infTime", "rGC.sympt", "uGC.sympt", "pGC.sympt",
  "rGC.tx", "uGC.tx", "pGC.tx", "rGC.txTime", "uGC.txTime", "pGC.txTime", "rGC.dur", "uGC.dur", "pGC.dur",
  "last.screen", "sti.expo",
  "status", "inf.time", "stage", "stage.time", "diag.status", "diag.time", "last.neg.test", "tx.status",
  "cuml.time.on.tx", "cuml.time.off.tx", "tx.init.time", "vl",
  "prepStat", "prepClass", "prepElig", "prepStartTime", "prepLastRisk",
  "prep.ind.last"
};
#define NA (-1.0)
#define WEEK (7.0 / 365.0)
#define LOG10_200 2.301029995663981
#define LN_2_45 0.8960880245566357

typedef struct { int p1, p2, start; int ai, oi, kiss, rim; } orc_edge;
typedef struct { int layer, e, type, k, uai, p1act; } orc_act;

struct orc_dat {
  int n, cap, next_uid, t_ref, t_age, t_cap, replicate_id;
  uint32_t seed_lo; uint32_t status;
  double* a[A_NATTR];
  orc_edge* el[3]; int ne[3]; int ecap[3];
  orc_act* al; int nal, alcap, al_at; int al_has_uai, al_has_pos;
  uint32_t* newmask;        /* per position: GC pathway bits 0..6, HIV bit 7 */
  uint32_t* cnt;            /* [t_cap][XGC_NCOUNTER] */
  double p[XGC_NPARAM]; int32_t c[XGC_NCONTROL]; double tb[XGC_NTABLE];
  const xgc_inject* inj;    /* valid during orc_step only */
};

/* ------------------------------------------------------------------------------------------------------------------ */
/* random draws                                                                                                       */
/* ------------------------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123).  Restated from the published algorithm. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t m0 = (uint64_t)0xD2511F53u * c0, m1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(m1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)m1;
    uint32_t n2 = (uint32_t)(m0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)m0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_uniform(const orc_dat* d, int at, int stream, uint32_t e0, uint32_t e1, int k) {
  uint32_t ctr[4] = { (uint32_t)at, ((uint32_t)stream << 16) | (uint32_t)(k >> 2), e0, e1 };
  uint32_t key[2] = { d->seed_lo, (uint32_t)d->replicate_id };
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  return ((double)o[k & 3] + 0.5) * (1.0 / 4294967296.0);
}

/* one indexed uniform: injected table if present, else Philox */
static double U(const orc_dat* d, int at, int stream, uint32_t e0, uint32_t e1, size_t idx, int k) {
  if (d->inj && d->inj->u) {
    size_t off = xgc_inject_offset(d->inj, stream) + idx * (size_t)xgc_stream_width(stream) + (size_t)k;
    return d->inj->u[off];
  }
  return orc_uniform(d, at, stream, e0, e1, k);
}
#define UA(stream, i, k)  U(d, at, (stream), (uint32_t)d->a[A_UID][i], 0u, (size_t)d->a[A_UID][i], (k))
#define UR(stream, k)     U(d, at, (stream), 0u, 0u, 0, (k))

/* exp(): Cody-Waite reduction by ln2 + degree-13 Taylor polynomial, Horner, no FMA.  The CUDA path implements the
 * same sequence of IEEE operations so both sides agree to the bit (SURVEY §7.3 "xgc_math"). */
double orc_exp(double x) {
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
  int64_t k = (int64_t)kf;
  union { uint64_t u; double f; } s;
  s.u = (uint64_t)(k + 1023) << 52;
  return q * s.f;
}

double orc_any_of_n(double t, int n) { return xgc_any_of_n(t, n); }

/* rbinom(1, 1, p) by inversion, as base R does it */
int orc_bern(double u, double p) {
  if (p <= 0.5) return u >= 1.0 - p;
  return u < p;
}

/* single-uniform Poisson by CDF walk, truncated at kmax (own definition) */
int orc_pois(double u, double mu, int kmax) {
  /* p_k = p_{k-1} * mu * (1/k) with correctly rounded reciprocal constants for k <= 32, a true division beyond */
  static const double INVK[33] = { 0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
    1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24,
    1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32 };
  if (!(mu > 0.0)) return 0;
  double pr = orc_exp(-mu), cdf = pr;
  int k = 0;
  while (u > cdf && k < kmax) { k++; pr = k <= 32 ? pr * mu * INVK[k] : pr * mu / (double)k; cdf = cdf + pr; }
  return k;
}

static int categorical(double u, const double* prob, int n) {
  int r = 0; double c = prob[0];
  while (u >= c && r < n - 1) { r++; c = c + prob[r]; }
  return r;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* life cycle / state exchange                                                                                        */
/* ------------------------------------------------------------------------------------------------------------------ */
orc_dat* orc_create(int cap, const int e_cap[3], int t_cap, uint64_t seed, int replicate_id) {
  orc_dat* d = (orc_dat*)calloc(1, sizeof(orc_dat));
  d->cap = cap; d->t_cap = t_cap; d->seed_lo = (uint32_t)seed; d->replicate_id = replicate_id;
  for (int k = 0; k < A_NATTR; ++k) d->a[k] = (double*)calloc((size_t)cap, sizeof(double));
  for (int l = 0; l < 3; ++l) { d->ecap[l] = e_cap[l]; d->el[l] = (orc_edge*)calloc((size_t)e_cap[l] + 1, sizeof(orc_edge)); }
  d->newmask = (uint32_t*)calloc((size_t)cap, sizeof(uint32_t));
  d->cnt = (uint32_t*)calloc((size_t)t_cap * XGC_NCOUNTER, sizeof(uint32_t));
  d->alcap = 1024; d->al = (orc_act*)malloc(sizeof(orc_act) * (size_t)d->alcap); d->al_at = -1;
  d->t_ref = 1; d->t_age = 1;
  static const double dflt[] = {
#define XGC_PARAM(name, count, v) [XGC_P_##name ... XGC_P_##name##_END] = (v),
#include "xgc/params.def"
#undef XGC_PARAM
  };
  memcpy(d->p, dflt, sizeof(double) * XGC_NPARAM);
  d->c[XGC_C_transRoute_Kissing] = 1; d->c[XGC_C_transRoute_Rimming] = 1;
  return d;
}
void orc_destroy(orc_dat* d) {
  if (!d) return;
  for (int k = 0; k < A_NATTR; ++k) free(d->a[k]);
  for (int l = 0; l < 3; ++l) free(d->el[l]);
  free(d->newmask); free(d->cnt); free(d->al); free(d);
}
void orc_set_params(orc_dat* d, const double* p) { memcpy(d->p, p, sizeof(double) * XGC_NPARAM); }
void orc_set_control(orc_dat* d, const int32_t* c) { memcpy(d->c, c, sizeof(int32_t) * XGC_NCONTROL); }
void orc_set_tables(orc_dat* d, const double* t) { memcpy(d->tb, t, sizeof(double) * XGC_NTABLE); }
int orc_num_agents(const orc_dat* d) { return d->n; }
/* XGC_ST_RATE_TRUNCATED: the tail mass of a constant-rate Poisson beyond XGC_MAX_ACTS, with the recurrence of orc_pois */
static uint32_t rate_status(const orc_dat* d) {
  static const int rate[4] = { XGC_P_kiss_rate_main, XGC_P_kiss_rate_casl, XGC_P_rim_rate_main, XGC_P_rim_rate_casl };
  uint32_t st = 0;
  for (int w = 0; w < 4; ++w) {
    double mu = d->p[rate[w]];
    if (!(mu > 0.0)) continue;
    double pr = orc_exp(-mu), cdf = pr;
    for (int k = 1; k <= XGC_MAX_ACTS; ++k) { pr = pr * mu * (1.0 / (double)k); cdf = cdf + pr; }
    if (1.0 - cdf > XGC_TRUNC_TOL) st |= XGC_ST_RATE_TRUNCATED;
  }
  return st;
}
uint32_t orc_status(const orc_dat* d) { return d->status | rate_status(d); }

static void new_agent_defaults(orc_dat* d, int i) {
  for (int k = 0; k < A_NATTR; ++k) d->a[k][i] = 0.0;
  static const int na_attrs[] = { A_GC_INF, A_GC_INF + 1, A_GC_INF + 2, A_GC_TXT, A_GC_TXT + 1, A_GC_TXT + 2,
    A_GC_DUR, A_GC_DUR + 1, A_GC_DUR + 2, A_LASTSCR, A_INFT, A_STAGET, A_DIAGT, A_LNT, A_TXINIT, A_PREPSTART,
    A_PREPLR, A_INDLAST };
  for (size_t j = 0; j < sizeof(na_attrs) / sizeof(int); ++j) d->a[na_attrs[j]][i] = NA;
  d->a[A_RACE][i] = 1; d->a[A_AGEGRP][i] = 1; d->a[A_TTTRAJ][i] = 2; d->a[A_RISK][i] = 1;
}

static int age_group(double age) { return 1 + (age >= 25.0) + (age >= 35.0) + (age >= 45.0) + (age >= 55.0); }

void orc_set_population(orc_dat* d, int n, int at) {
  d->n = n; d->next_uid = n; d->t_ref = at; d->t_age = at;
  for (int i = 0; i < n; ++i) {
    new_agent_defaults(d, i);
    d->a[A_UID][i] = i; d->a[A_ARRT][i] = at; d->a[A_AGE][i] = d->a[A_AGEREF][i] = 18.0;
  }
  for (int l = 0; l < 3; ++l) d->ne[l] = 0;
}

static int attr_index(const char* name) {
  for (int k = 0; k < A_NATTR; ++k) if (strcmp(name, ATTR_NAME[k]) == 0) return k;
  return -1;
}
int orc_set_attr(orc_dat* d, const char* name, const double* v, int n) {
  int k = attr_index(name);
  if (k < 0 || n != d->n) return -1;
  memcpy(d->a[k], v, sizeof(double) * (size_t)n);
  if (k == A_AGE) {            /* re-base the closed-form age clock on the current week */
    memcpy(d->a[A_AGEREF], v, sizeof(double) * (size_t)n);
    d->t_ref = d->t_age;
    for (int i = 0; i < n; ++i) d->a[A_AGEGRP][i] = age_group(v[i]);
  }
  if (k == A_VL) for (int i = 0; i < n; ++i) d->a[A_VL][i] = (double)(float)v[i];
  if (k == A_INSQ) for (int i = 0; i < n; ++i) d->a[A_INSQ][i] = (double)(float)v[i];
  if (k == A_UID) { int mx = -1; for (int i = 0; i < n; ++i) if ((int)v[i] > mx) mx = (int)v[i]; d->next_uid = mx + 1; }
  return 0;
}
int orc_get_attr(const orc_dat* d, const char* name, double* v, int cap) {
  static const char* const degn[3] = { "deg.main", "deg.casl", "deg.inst" };     /* derived: current degree per layer */
  for (int l = 0; l < 3; ++l)
    if (strcmp(name, degn[l]) == 0) {
      if (cap < d->n) return -1;
      for (int i = 0; i < d->n; ++i) v[i] = 0.0;
      for (int e = 0; e < d->ne[l]; ++e) { v[d->el[l][e].p1] += 1.0; v[d->el[l][e].p2] += 1.0; }
      return d->n;
    }
  int k = attr_index(name);
  if (k < 0 || cap < d->n) return -1;
  memcpy(v, d->a[k], sizeof(double) * (size_t)d->n);
  return d->n;
}
int orc_set_edgelist(orc_dat* d, int layer, const int32_t* el, int n_edges, const int32_t* start) {
  if (n_edges > d->ecap[layer]) return -1;
  for (int e = 0; e < n_edges; ++e) {
    orc_edge* g = &d->el[layer][e];
    memset(g, 0, sizeof(*g));
    g->p1 = el[e] - 1; g->p2 = el[n_edges + e] - 1; g->start = start ? start[e] : d->t_age;
  }
  d->ne[layer] = n_edges;
  return 0;
}
int orc_get_edgelist(const orc_dat* d, int layer, int32_t* el, int cap, int32_t* start) {
  int ne = d->ne[layer];
  if (cap < ne) return -1;
  for (int e = 0; e < ne; ++e) {
    el[e] = d->el[layer][e].p1 + 1; el[ne + e] = d->el[layer][e].p2 + 1;
    if (start) start[e] = d->el[layer][e].start;
  }
  return ne;
}
int orc_get_actcounts(const orc_dat* d, int layer, int32_t* counts, int cap) {
  int ne = d->ne[layer];
  if (cap < ne) return -1;
  for (int e = 0; e < ne; ++e) {
    const orc_edge* g = &d->el[layer][e];
    counts[e] = g->ai; counts[ne + e] = g->oi; counts[2 * ne + e] = g->kiss; counts[3 * ne + e] = g->rim;
  }
  return ne;
}

static uint32_t* CNT(orc_dat* d, int at) { return d->cnt + (size_t)at * XGC_NCOUNTER; }
static int cell_of(const orc_dat* d, int i) { return ((int)d->a[A_RACE][i] - 1) * 5 + ((int)d->a[A_AGEGRP][i] - 1); }

/* ------------------------------------------------------------------------------------------------------------------ */
/* aging_msm (sim1_02_lhs.r:202; Appendix E "aging").  age advances 7/365 per week; evaluated in closed form from the */
/* reference week so that no rounding accumulates: age(at) = age.ref + (at - t.ref) * (7/365).                        */
/* ------------------------------------------------------------------------------------------------------------------ */
static void aging_msm(orc_dat* d, int at) {
  d->t_age = at;
  for (int i = 0; i < d->n; ++i) {
    d->a[A_AGE][i] = d->a[A_AGEREF][i] + (double)(at - d->t_ref) * WEEK;
    d->a[A_AGEGRP][i] = age_group(d->a[A_AGE][i]);
  }
}

/* tergmLite::delete_vertices semantics [BK]: drop incident edges, shift remaining ids down, keep edge order */
static void delete_vertices(orc_dat* d, const int* dead) {
  int* newpos = (int*)malloc(sizeof(int) * (size_t)d->n);
  int m = 0;
  for (int i = 0; i < d->n; ++i) newpos[i] = dead[i] ? -1 : m++;
  for (int l = 0; l < 3; ++l) {
    int w = 0;
    for (int e = 0; e < d->ne[l]; ++e) {
      orc_edge g = d->el[l][e];
      if (newpos[g.p1] < 0 || newpos[g.p2] < 0) continue;
      g.p1 = newpos[g.p1]; g.p2 = newpos[g.p2];
      d->el[l][w++] = g;
    }
    d->ne[l] = w;
  }
  for (int k = 0; k < A_NATTR; ++k) {               /* deleteAttr on every attribute vector */
    double* v = d->a[k];
    for (int i = 0; i < d->n; ++i) if (newpos[i] >= 0) v[newpos[i]] = v[i];
  }
  d->n = m;
  free(newpos);
}

/* departure_msm (sim1_02_lhs.r:203; Appendix E "departure") */
static void departure_msm(orc_dat* d, int at) {
  int* dead = (int*)calloc((size_t)d->n, sizeof(int));
  int any = 0;
  uint32_t* c = CNT(d, at);
  for (int i = 0; i < d->n; ++i) {
    int ai = (int)floor(d->a[A_AGE][i]);
    if (ai < 0) ai = 0;
    if (ai > XGC_ASMR_AGES - 1) ai = XGC_ASMR_AGES - 1;
    double pm = d->tb[XGC_T_ASMR + ((int)d->a[A_RACE][i] - 1) * XGC_ASMR_AGES + ai];
    int dg = orc_bern(UA(XGC_S_AGENT1, i, 0), pm);
    int da = 0;
    if ((int)d->a[A_STAGE][i] == 4) da = orc_bern(UA(XGC_S_AGENT1, i, 1), d->p[XGC_P_aids_mr]);
    if (dg) c[XGC_K_DEP_GEN]++;
    if (da) c[XGC_K_DEP_AIDS]++;
    if (dg || da) { dead[i] = 1; any = 1; }
  }
  if (any) delete_vertices(d, dead);
  free(dead);
}

/* arrival_msm (sim1_02_lhs.r:204, 63-66; Appendix E "arrival") */
static void arrival_msm(orc_dat* d, int at) {
  double mu = d->p[XGC_P_a_rate] * d->p[XGC_P_n_init];
  int m = (int)ceil(mu / 32.0);
  if (m < 1) m = 1;
  if (m > 64) m = 64;
  double muc = mu / (double)m;
  int nnew = 0;
  for (int ch = 0; ch < m; ++ch) nnew += orc_pois(UR(XGC_S_ARRIVE_N, ch), muc, 255);
  if (d->n + nnew > d->cap) { d->status |= XGC_ST_AGENT_OVERFLOW; nnew = d->cap - d->n; }
  CNT(d, at)[XGC_K_NNEW] += (uint32_t)nnew;
  for (int j = 0; j < nnew; ++j) {
    int i = d->n + j;
    new_agent_defaults(d, i);
    d->a[A_UID][i] = d->next_uid + j;
    int race = categorical(UA(XGC_S_ARRIVE_ATTR, i, 0), d->tb + XGC_T_RACE_PROP, 4);          /* 0..3 */
    int role = categorical(UA(XGC_S_ARRIVE_ATTR, i, 1), d->tb + XGC_T_ROLE_PROB + race * 3, 3);
    d->a[A_RACE][i] = race + 1; d->a[A_ROLE][i] = role;
    d->a[A_INSQ][i] = role == 0 ? 1.0 : role == 1 ? 0.0 : (double)(float)UA(XGC_S_ARRIVE_ATTR, i, 2);
    d->a[A_CIRC][i] = orc_bern(UA(XGC_S_ARRIVE_ATTR, i, 3), d->p[XGC_P_circ_prob + race]);
    d->a[A_LATE][i] = orc_bern(UA(XGC_S_ARRIVE_ATTR, i, 4), d->p[XGC_P_hiv_test_late_prob + race]);
    double tt[3] = { d->p[XGC_P_tt_part_supp + race], d->p[XGC_P_tt_full_supp + race], d->p[XGC_P_tt_dur_supp + race] };
    d->a[A_TTTRAJ][i] = 1 + categorical(UA(XGC_S_ARRIVE_ATTR, i, 5), tt, 3);
    d->a[A_STOPPER][i] = orc_bern(UA(XGC_S_ARRIVE_ATTR, i, 6), d->p[XGC_P_act_stopper_prob]);
    int rg = (int)(UA(XGC_S_ARRIVE_ATTR, i, 7) * 5.0);
    d->a[A_RISK][i] = 1 + (rg > 4 ? 4 : rg);
    d->a[A_AGEREF][i] = d->p[XGC_P_arrival_age] - (double)(d->t_age - d->t_ref) * WEEK;
    d->a[A_AGE][i] = d->a[A_AGEREF][i] + (double)(d->t_age - d->t_ref) * WEEK;
    d->a[A_AGEGRP][i] = age_group(d->a[A_AGE][i]);
    d->a[A_ARRT][i] = at;
  }
  d->n += nnew; d->next_uid += nnew;
}

/* hivtest_msm (sim1_02_lhs.r:205, 106-113; Appendix E "HIV block") */
static void hivtest_msm(orc_dat* d, int at) {
  uint32_t* c = CNT(d, at);
  for (int i = 0; i < d->n; ++i) {
    if ((int)d->a[A_DIAG][i] == 1) continue;
    int race = (int)d->a[A_RACE][i] - 1;
    double last = d->a[A_LNT][i] == NA ? d->a[A_ARRT][i] : d->a[A_LNT][i];
    double tsl = (double)at - last;
    int tested = 0;
    if ((int)d->a[A_PREPSTAT][i] == 0 && (int)d->a[A_LATE][i] == 0 && tsl >= 2.0)
      tested = orc_bern(UA(XGC_S_AGENT1, i, 2), d->p[XGC_P_hiv_test_rate + race]);
    if ((int)d->a[A_STAGE][i] == 4 && tsl >= d->p[XGC_P_aids_test_int]) tested = 1;
    if ((int)d->a[A_PREPSTAT][i] == 1 && tsl >= d->p[XGC_P_prep_tst_int]) tested = 1;
    if (!tested) continue;
    c[XGC_K_HIVTESTS]++;
    if ((int)d->a[A_STATUS][i] == 1 && (double)at - d->a[A_INFT][i] >= d->p[XGC_P_test_window_int]) {
      d->a[A_DIAG][i] = 1; d->a[A_DIAGT][i] = at;
    } else d->a[A_LNT][i] = at;
  }
}

/* hivtx_msm (sim1_02_lhs.r:206, 114-139) */
static void hivtx_msm(orc_dat* d, int at) {
  for (int i = 0; i < d->n; ++i) {
    if ((int)d->a[A_STATUS][i] != 1) continue;
    int race = (int)d->a[A_RACE][i] - 1, traj = (int)d->a[A_TTTRAJ][i];
    if ((int)d->a[A_DIAG][i] == 1) {
      int txs = (int)d->a[A_TXS][i]; double cumon = d->a[A_CUMON][i];
      if (txs == 0 && cumon == 0.0) {
        if (orc_bern(UA(XGC_S_AGENT1, i, 3), d->p[XGC_P_tx_init_prob + race])) { d->a[A_TXS][i] = 1; d->a[A_TXINIT][i] = at; }
      } else if (txs == 1) {
        double rr = traj == 2 ? d->p[XGC_P_tx_halt_full_rr + race] : traj == 3 ? d->p[XGC_P_tx_halt_dur_rr + race] : 1.0;
        if (orc_bern(UA(XGC_S_AGENT1, i, 3), d->p[XGC_P_tx_halt_part_prob + race] * rr)) d->a[A_TXS][i] = 0;
      } else {
        double rr = traj == 2 ? d->p[XGC_P_tx_reinit_full_rr + race] : traj == 3 ? d->p[XGC_P_tx_reinit_dur_rr + race] : 1.0;
        if (orc_bern(UA(XGC_S_AGENT1, i, 3), d->p[XGC_P_tx_reinit_part_prob + race] * rr)) d->a[A_TXS][i] = 1;
      }
    }
    if ((int)d->a[A_TXS][i] == 1) d->a[A_CUMON][i] += 1.0; else d->a[A_CUMOFF][i] += 1.0;
  }
}

/* hivprogress_msm (sim1_02_lhs.r:207) */
static void hivprogress_msm(orc_dat* d, int at) {
  const double* p = d->p;
  for (int i = 0; i < d->n; ++i) {
    if ((int)d->a[A_STATUS][i] != 1) continue;
    double tsi = (double)at - d->a[A_INFT][i];
    if ((int)d->a[A_STAGE][i] == 1 && tsi >= p[XGC_P_vl_acute_rise_int]) { d->a[A_STAGE][i] = 2; d->a[A_STAGET][i] = at; }
    if ((int)d->a[A_STAGE][i] == 2 && tsi >= p[XGC_P_vl_acute_rise_int] + p[XGC_P_vl_acute_fall_int]) { d->a[A_STAGE][i] = 3; d->a[A_STAGET][i] = at; }
    if ((int)d->a[A_STAGE][i] == 3) {
      int aids = 0, traj = (int)d->a[A_TTTRAJ][i];
      double on = d->a[A_CUMON][i], off = d->a[A_CUMOFF][i];
      if (on == 0.0 && tsi >= p[XGC_P_vl_aids_onset_int]) aids = 1;
      if (traj == 1 && on > 0.0 && off / p[XGC_P_max_time_off_tx_part_int] + on / p[XGC_P_max_time_on_tx_part_int] >= 1.0) aids = 1;
      if (traj >= 2 && (int)d->a[A_TXS][i] == 0 && on > 0.0 && off >= p[XGC_P_max_time_off_tx_full_int]) aids = 1;
      if (aids) { d->a[A_STAGE][i] = 4; d->a[A_STAGET][i] = at; }
    }
  }
}

/* hivvl_msm (sim1_02_lhs.r:208) */
static void hivvl_msm(orc_dat* d, int at) {
  const double* p = d->p;
  for (int i = 0; i < d->n; ++i) {
    if ((int)d->a[A_STATUS][i] != 1) continue;
    double tsi = (double)at - d->a[A_INFT][i], v, vl = d->a[A_VL][i];
    int stage = (int)d->a[A_STAGE][i];
    double peak = p[XGC_P_vl_acute_peak], sp = p[XGC_P_vl_set_point];
    if ((int)d->a[A_TXS][i] == 0) {
      if (stage == 1) { v = peak * (tsi / p[XGC_P_vl_acute_rise_int]); if (v > peak) v = peak; }
      else if (stage == 2) { v = peak - (peak - sp) * ((tsi - p[XGC_P_vl_acute_rise_int]) / p[XGC_P_vl_acute_fall_int]); if (v < sp) v = sp; }
      else if (stage == 3) { if (d->a[A_CUMON][i] == 0.0) v = sp; else { v = vl + p[XGC_P_vl_tx_up_slope]; if (v > sp) v = sp; } }
      else { double tsa = (double)at - d->a[A_STAGET][i];
             v = sp + (tsa / p[XGC_P_vl_aids_int]) * (p[XGC_P_vl_aids_peak] - sp); if (v > p[XGC_P_vl_aids_peak]) v = p[XGC_P_vl_aids_peak]; }
    } else {
      double target = (int)d->a[A_TTTRAJ][i] == 1 ? p[XGC_P_vl_part_supp] : p[XGC_P_vl_full_supp];
      double slope = stage == 4 ? p[XGC_P_vl_tx_aids_down_slope] : p[XGC_P_vl_tx_down_slope];
      v = vl - slope; if (v < target) v = target;
    }
    d->a[A_VL][i] = (double)(float)v;
  }
}

/* surrogate for simnet_msm (sim1_02_lhs.r:209).  NOT the reference's STERGM (which stays in host R): per-edge
 * Bernoulli dissolution + uniform random formation holding the mean degree; benchmarks and statistical tests only. */
static int roles_compatible(const orc_dat* d, int a, int b) {
  int ra = (int)d->a[A_ROLE][a], rb = (int)d->a[A_ROLE][b];
  return !((ra == 0 && rb == 0) || (ra == 1 && rb == 1));
}
/* formation weight of agent i in layer l (abi.h XGC_T_NET_DEGW / XGC_T_NET_RISKW): degree classes over the ties that survived
 * this week's dissolution, the one-off layer also by risk group */
static double form_weight(const orc_dat* d, int l, int i, const int* dm, const int* dc) {
  int cm = dm[i] > 2 ? 2 : dm[i], cc = dc[i] > 2 ? 2 : dc[i];
  double w = d->tb[XGC_T_NET_DEGW + l * 9 + cm * 3 + cc];
  if (l == 2) w = w * d->tb[XGC_T_NET_RISKW + (int)d->a[A_RISK][i] - 1];
  return w;
}
static void resim_nets_surrogate(orc_dat* d, int at) {
  /* degrees (deg.main, deg.casl) in the network the week starts with: formation and dissolution both condition on it
   * (a separable model: the ties that dissolve this week still count), fixed while this week's ties form */
  int* dm = (int*)calloc((size_t)(d->n > 0 ? d->n : 1), sizeof(int));
  int* dc = (int*)calloc((size_t)(d->n > 0 ? d->n : 1), sizeof(int));
  for (int e = 0; e < d->ne[0]; ++e) { dm[d->el[0][e].p1]++; dm[d->el[0][e].p2]++; }
  for (int e = 0; e < d->ne[1]; ++e) { dc[d->el[1][e].p1]++; dc[d->el[1][e].p2]++; }
  /* dissolution of the persistent layers (one-off contacts last one week) */
  for (int l = 0; l < 3; ++l) {
    int w = 0, ne0 = d->ne[l];
    int es = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
    if (l < 2) {
      double pd = 1.0 / d->tb[XGC_T_NET_DUR + l];
      for (int e = 0; e < ne0; ++e) {
        orc_edge g = d->el[l][e];
        double u = U(d, at, es + (XGC_S_DISSOLVE - XGC_S_EDGE0), (uint32_t)d->a[A_UID][g.p1], (uint32_t)d->a[A_UID][g.p2], (size_t)e, 0);
        if (orc_bern(u, pd)) continue;
        d->el[l][w++] = g;
      }
    }
    d->ne[l] = w;
  }
  for (int l = 0; l < 3; ++l) {
    int w = d->ne[l];
    double target = d->tb[XGC_T_NET_DEG + l] * (double)d->n / 2.0;
    double mu = l == 2 ? target : target / d->tb[XGC_T_NET_DUR + l];
    double fl = floor(mu);
    int nform = (int)fl + (UR(XGC_S_NET_N, l) < mu - fl);
    if (d->inj && d->inj->u && nform > d->inj->form_cap) nform = d->inj->form_cap;
    for (int j = 0; j < nform; ++j) {
      if (d->n < 2) break;
      for (int t = 0; t < XGC_FORM_TRIES; ++t) {
        double ua = U(d, at, XGC_S_FORM + l, (uint32_t)j, 0u, (size_t)j, 3 * t);
        double ub = U(d, at, XGC_S_FORM + l, (uint32_t)j, 0u, (size_t)j, 3 * t + 1);
        int a = (int)(ua * (double)d->n); if (a > d->n - 1) a = d->n - 1;
        int b = (int)(ub * (double)(d->n - 1)); if (b > d->n - 2) b = d->n - 2;
        if (b >= a) b++;
        if (!roles_compatible(d, a, b)) continue;
        double pw = form_weight(d, l, a, dm, dc) * form_weight(d, l, b, dm, dc);
        if (pw < 1.0 && !(U(d, at, XGC_S_FORM + l, (uint32_t)j, 0u, (size_t)j, 3 * t + 2) < pw)) continue;
        if (w >= d->ecap[l]) { d->status |= XGC_ST_EDGE_OVERFLOW; break; }
        orc_edge g; memset(&g, 0, sizeof(g));
        g.p1 = a < b ? a : b; g.p2 = a < b ? b : a; g.start = at;
        d->el[l][w++] = g;
        break;
      }
    }
    d->ne[l] = w;
    CNT(d, at)[XGC_K_NEDGES + l] = (uint32_t)w;
  }
  free(dm); free(dc);
}

/* linear predictor of the act-rate / condom GLMs (Appendix E "acts"/"condoms"; ambiguity A11) */
static double glm_eta(const double* b, double dur, int casual, double ca, int hcp, int prep, int r1, int r2) {
  double eta = b[XGC_G_B0];
  eta = eta + b[XGC_G_DUR] * dur;
  eta = eta + b[XGC_G_DUR2] * (dur * dur);
  eta = eta + b[XGC_G_RC + r1 * 4 + r2];
  eta = eta + b[XGC_G_PT2] * (double)casual;
  eta = eta + b[XGC_G_DUR_PT2] * (dur * (double)casual);
  eta = eta + b[XGC_G_CA] * ca;
  eta = eta + b[XGC_G_CA2] * (ca * ca);
  eta = eta + b[XGC_G_HCP] * (double)hcp;
  eta = eta + b[XGC_G_PREP] * (double)prep;
  return eta;
}

static int stopped(const orc_dat* d, int i) {
  if ((int)d->a[A_STOPPER][i] != 1) return 0;
  for (int s = 0; s < 3; ++s) if ((int)d->a[A_GC_SYM + s][i] == 1 || (int)d->a[A_GC_TX + s][i] == 1) return 1;
  return 0;
}

/* acts_msm (sim1_02_lhs.r:210, 141-150, 173-174; Appendix E "acts") */
static void acts_msm(orc_dat* d, int at) {
  const double* p = d->p;
  uint32_t* c = CNT(d, at);
  for (int l = 0; l < 3; ++l) {
    int es = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS + (XGC_S_ACTS - XGC_S_EDGE0);
    for (int e = 0; e < d->ne[l]; ++e) {
      orc_edge* g = &d->el[l][e];
      int a = g->p1, b = g->p2;
      uint32_t ua = (uint32_t)d->a[A_UID][a], ub = (uint32_t)d->a[A_UID][b];
      g->ai = g->oi = g->kiss = g->rim = 0;
      if (!(stopped(d, a) || stopped(d, b))) {
        int compat = roles_compatible(d, a, b);
        if (l < 2) {
          double dur = (double)(at - g->start), ca = d->a[A_AGE][a] + d->a[A_AGE][b];
          int hcp = (int)d->a[A_DIAG][a] == 1 && (int)d->a[A_DIAG][b] == 1;
          int r1 = (int)d->a[A_RACE][a] - 1, r2 = (int)d->a[A_RACE][b] - 1;
          double rai = orc_exp(glm_eta(d->tb + XGC_T_GLM_AI, dur, l == 1, ca, hcp, 0, r1, r2)) / 52.0 * p[XGC_P_ai_acts_scale_mc];
          double roi = orc_exp(glm_eta(d->tb + XGC_T_GLM_OI, dur, l == 1, ca, hcp, 0, r1, r2)) / 52.0 * p[XGC_P_oi_acts_scale_mc];
          g->ai = compat ? orc_pois(U(d, at, es, ua, ub, (size_t)e, 0), rai, XGC_MAX_ACTS) : 0;
          g->oi = orc_pois(U(d, at, es, ua, ub, (size_t)e, 1), roi, XGC_MAX_ACTS);
          g->kiss = orc_pois(U(d, at, es, ua, ub, (size_t)e, 2), l == 0 ? p[XGC_P_kiss_rate_main] : p[XGC_P_kiss_rate_casl], XGC_MAX_ACTS);
          g->rim = orc_pois(U(d, at, es, ua, ub, (size_t)e, 3), l == 0 ? p[XGC_P_rim_rate_main] : p[XGC_P_rim_rate_casl], XGC_MAX_ACTS);
        } else {
          g->ai = compat ? 1 : 0;
          g->oi = orc_bern(U(d, at, es, ua, ub, (size_t)e, 1), p[XGC_P_oi_prob_oo]);
          g->kiss = orc_bern(U(d, at, es, ua, ub, (size_t)e, 2), p[XGC_P_kiss_prob_oo]);
          g->rim = orc_bern(U(d, at, es, ua, ub, (size_t)e, 3), p[XGC_P_rim_prob_oo]);
        }
      }
      c[XGC_K_NACTS] += (uint32_t)g->ai; c[XGC_K_NACTS + 1] += (uint32_t)g->oi;     /* kissing / rimming are not tallied */
      /* exposure history for the CDC screening protocol (ambiguity A4) */
      for (int side = 0; side < 2; ++side) {
        int i = side ? b : a, role = (int)d->a[A_ROLE][i], ex = (int)d->a[A_EXPO][i];
        if (g->ai > 0 && role != 0) ex |= 1;
        if ((g->ai > 0 && role != 1) || g->oi > 0) ex |= 2;
        if (g->oi > 0 || g->rim > 0) ex |= 4;
        if (g->kiss > 0) ex |= 8;
        if (l != 0 && (g->ai | g->oi | g->kiss | g->rim)) ex |= 16;
        d->a[A_EXPO][i] = ex;
      }
    }
  }
  d->al_at = -1;
}

/* condoms_msm + position_msm (sim1_02_lhs.r:211-212): expand edges into the act list and draw condom use for anal
 * acts ("All act lists are finalized here"), then assign position / site pairing (Appendix E; ambiguities A9, A10). */
static void al_push(orc_dat* d, orc_act r) {
  if (d->nal == d->alcap) { d->alcap *= 2; d->al = (orc_act*)realloc(d->al, sizeof(orc_act) * (size_t)d->alcap); }
  d->al[d->nal++] = r;
}
static double cond_prob(const orc_dat* d, int at, int l, const orc_edge* g) {
  int a = g->p1, b = g->p2;
  double ca = d->a[A_AGE][a] + d->a[A_AGE][b];
  int hcp = (int)d->a[A_DIAG][a] == 1 && (int)d->a[A_DIAG][b] == 1;
  int prep = (int)d->a[A_PREPSTAT][a] == 1 || (int)d->a[A_PREPSTAT][b] == 1;
  int r1 = (int)d->a[A_RACE][a] - 1, r2 = (int)d->a[A_RACE][b] - 1;
  double eta = l < 2 ? glm_eta(d->tb + XGC_T_GLM_CONDMC, (double)(at - g->start), l == 1, ca, hcp, prep, r1, r2)
                     : glm_eta(d->tb + XGC_T_GLM_CONDOO, 0.0, 0, ca, hcp, prep, r1, r2);
  double pc = 1.0 / (1.0 + orc_exp(-eta));
  pc = pc * d->p[XGC_P_cond_scale];
  if (pc > 1.0) pc = 1.0;
  return pc;
}
static void condoms_msm(orc_dat* d, int at) {
  d->nal = 0; d->al_at = at; d->al_has_pos = 0;
  for (int l = 0; l < 3; ++l) {
    int base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
    for (int e = 0; e < d->ne[l]; ++e) {
      const orc_edge* g = &d->el[l][e];
      int cnts[4] = { g->ai, g->oi, g->kiss, g->rim };
      uint32_t ua = (uint32_t)d->a[A_UID][g->p1], ub = (uint32_t)d->a[A_UID][g->p2];
      double pc = g->ai > 0 ? cond_prob(d, at, l, g) : 0.0;
      for (int type = 0; type < 4; ++type)
        for (int k = 0; k < cnts[type]; ++k) {
          orc_act r = { l, e, type, k, 0, 0 };
          if (type == 0) r.uai = orc_bern(U(d, at, base + (XGC_S_ANAL - XGC_S_EDGE0), ua, ub, (size_t)e, 4 * k), 1.0 - pc);
          al_push(d, r);
        }
    }
  }
  d->al_has_uai = 1;
}
static void position_msm(orc_dat* d, int at) {
  for (int j = 0; j < d->nal; ++j) {
    orc_act* r = &d->al[j];
    const orc_edge* g = &d->el[r->layer][r->e];
    int base = XGC_S_EDGE0 + r->layer * XGC_EDGE_STREAMS;
    uint32_t ua = (uint32_t)d->a[A_UID][g->p1], ub = (uint32_t)d->a[A_UID][g->p2];
    if (r->type == 0) {
      int r1 = (int)d->a[A_ROLE][g->p1], r2 = (int)d->a[A_ROLE][g->p2];
      if (r1 == 0 || r2 == 1) r->p1act = 1;                    /* p1 insertive-only or p2 receptive-only */
      else if (r1 == 1 || r2 == 0) r->p1act = 0;
      else {
        double q1 = d->a[A_INSQ][g->p1], q2 = d->a[A_INSQ][g->p2];
        double pi = q1 + q2 > 0.0 ? q1 / (q1 + q2) : 0.5;
        r->p1act = orc_bern(U(d, at, base + (XGC_S_ANAL - XGC_S_EDGE0), ua, ub, (size_t)r->e, 4 * r->k + 1), pi);
      }
    } else if (r->type == 1 || r->type == 3) {
      int st = base + ((r->type == 1 ? XGC_S_ORAL : XGC_S_RIM) - XGC_S_EDGE0);
      if (d->inj && d->inj->u) {
        /* injected draws: one uniform per act (abi.h xgc_stream_width): p1 is the active partner iff u < 0.5 */
        r->p1act = U(d, at, st, ua, ub, (size_t)r->e, 1 + 2 * r->k) < 0.5;
      } else {
        /* native draws: direction of act j = bit (31 - j) of the 32-bit direction word floor(u * 2^32) (draw 0) */
        double u = U(d, at, st, ua, ub, (size_t)r->e, 0);
        uint32_t w = (uint32_t)(u * 4294967296.0);
        r->p1act = (int)((w >> (31 - r->k)) & 1u);
      }
    }
  }
  d->al_has_pos = 1;
}
static void ensure_al(orc_dat* d, int at) {
  if (d->al_at != at) { condoms_msm(d, at); position_msm(d, at); }
  else if (!d->al_has_pos) position_msm(d, at);
}
int orc_get_actlist(orc_dat* d, int at, int32_t* al, int cap, const xgc_inject* inj) {
  d->inj = inj; ensure_al(d, at); d->inj = NULL;
  if (cap < d->nal) return -1;
  for (int j = 0; j < d->nal; ++j) {
    const orc_act* r = &d->al[j];
    al[j] = r->layer; al[cap + j] = r->e; al[2 * cap + j] = r->type; al[3 * cap + j] = r->k;
    al[4 * cap + j] = r->uai; al[5 * cap + j] = r->p1act;
  }
  return d->nal;
}

/* risk history + prep_msm (sim1_02_lhs.r:213, 166-172; Appendix E "prep") */
static void riskhist_msm(orc_dat* d, int at) {
  ensure_al(d, at);
  for (int j = 0; j < d->nal; ++j) {
    const orc_act* r = &d->al[j];
    if (r->type != 0) continue;
    const orc_edge* g = &d->el[r->layer][r->e];
    for (int side = 0; side < 2; ++side) {
      int i = side ? g->p2 : g->p1, o = side ? g->p1 : g->p2;
      if ((int)d->a[A_DIAG][i] == 1) continue;
      /* indications: anal sex with a diagnosed partner, or any condomless anal sex (main, casual or one-off) */
      if ((int)d->a[A_DIAG][o] == 1 || r->uai) d->a[A_INDLAST][i] = at;
    }
  }
}
static int indicated(const orc_dat* d, int i, int at) {
  /* only the most recent indication event matters: indicated iff it lies inside the risk window */
  return d->a[A_INDLAST][i] != NA && d->a[A_INDLAST][i] >= (double)at - d->p[XGC_P_prep_risk_int];
}
static void prep_msm(orc_dat* d, int at) {
  const double* p = d->p;
  for (int i = 0; i < d->n; ++i) {
    int race = (int)d->a[A_RACE][i] - 1, ind = indicated(d, i, at), diag = (int)d->a[A_DIAG][i];
    d->a[A_PREPELIG][i] = (diag == 0 && ind);
    if ((int)d->a[A_PREPSTAT][i] == 1) {
      int stop = 0;
      if (diag == 1) stop = 1;
      else if (orc_bern(UA(XGC_S_AGENT2, i, 4), p[XGC_P_prep_discont_rate + race])) stop = 1;
      else if ((int)d->a[A_LNT][i] == at && (double)at - d->a[A_PREPLR][i] >= p[XGC_P_prep_reassess_int]) {
        if (!ind) stop = 1; else d->a[A_PREPLR][i] = at;
      }
      if (stop) { d->a[A_PREPSTAT][i] = 0; d->a[A_PREPCLASS][i] = 0; }
    } else if (diag == 0 && (double)at >= p[XGC_P_prep_start] && (int)d->a[A_LNT][i] == at && ind) {
      if (orc_bern(UA(XGC_S_AGENT2, i, 4), p[XGC_P_prep_start_prob + race])) {
        d->a[A_PREPSTAT][i] = 1; d->a[A_PREPSTART][i] = at; d->a[A_PREPLR][i] = at;
        d->a[A_PREPCLASS][i] = 1 + categorical(UA(XGC_S_AGENT2, i, 5), p + XGC_P_prep_adhr_dist, 3);
      }
    }
  }
}

/* hivtrans_msm (sim1_02_lhs.r:214, 151-164, 176-177; Appendix E "hivtrans") */
static void hivtrans_msm(orc_dat* d, int at) {
  ensure_al(d, at);
  const double* p = d->p;
  memset(d->newmask, 0, sizeof(uint32_t) * (size_t)d->n);
  for (int j = 0; j < d->nal; ++j) {
    const orc_act* r = &d->al[j];
    if (r->type != 0) continue;
    const orc_edge* g = &d->el[r->layer][r->e];
    int s1 = (int)d->a[A_STATUS][g->p1], s2 = (int)d->a[A_STATUS][g->p2];
    if (s1 == s2) continue;
    int pos = s1 ? g->p1 : g->p2, neg = s1 ? g->p2 : g->p1;
    int ins = r->p1act ? g->p1 : g->p2;
    int nrace = (int)d->a[A_RACE][neg] - 1, stage = (int)d->a[A_STAGE][pos];
    double vlf = orc_exp((d->a[A_VL][pos] - 4.5) * LN_2_45);
    double t;
    if (ins == pos) { t = p[XGC_P_urai_prob] * vlf; if ((int)d->a[A_GC][neg] == 1) t = t * p[XGC_P_hiv_rgc_rr]; }
    else { t = p[XGC_P_uiai_prob] * vlf; if ((int)d->a[A_CIRC][neg] == 1) t = t * p[XGC_P_circ_rr];
           if ((int)d->a[A_GC + 1][neg] == 1) t = t * p[XGC_P_hiv_ugc_rr]; }
    if (!r->uai) t = t * (1.0 - p[XGC_P_cond_eff] * (1.0 - p[XGC_P_cond_fail + nrace]));
    if (stage == 1 || stage == 2) t = t * p[XGC_P_acute_rr];
    if ((int)d->a[A_PREPSTAT][neg] == 1) { int cl = (int)d->a[A_PREPCLASS][neg]; if (cl >= 1) t = t * p[XGC_P_prep_adhr_hr + cl - 1]; }
    t = t * p[XGC_P_trans_scale + nrace];
    if (t > 1.0) t = 1.0;
    int base = XGC_S_EDGE0 + r->layer * XGC_EDGE_STREAMS;
    double u = U(d, at, base + (XGC_S_ANAL - XGC_S_EDGE0), (uint32_t)d->a[A_UID][g->p1], (uint32_t)d->a[A_UID][g->p2], (size_t)r->e, 4 * r->k + 2);
    if (orc_bern(u, t)) d->newmask[neg] |= 1u << 7;
  }
  uint32_t* c = CNT(d, at);
  for (int i = 0; i < d->n; ++i) {
    if (!(d->newmask[i] & (1u << 7))) continue;
    d->a[A_STATUS][i] = 1; d->a[A_INFT][i] = at; d->a[A_STAGE][i] = 1; d->a[A_STAGET][i] = at; d->a[A_VL][i] = 0.0;
    d->a[A_DIAG][i] = 0; d->a[A_TXS][i] = 0; d->a[A_CUMON][i] = 0; d->a[A_CUMOFF][i] = 0;
    c[XGC_K_INCID_HIV * XGC_NCELL + cell_of(d, i)]++;
  }
}

/* stirecov_msm (sim1_02_lhs.r:215, 77-83, 222; Appendix E "stirecov"; ambiguities A6, A7) */
static void stirecov_msm(orc_dat* d, int at) {
  const double* p = d->p;
  static const int txpr[3] = { XGC_P_rgc_tx_recov_pr, XGC_P_ugc_tx_recov_pr, XGC_P_pgc_tx_recov_pr };
  for (int i = 0; i < d->n; ++i)
    for (int s = 0; s < 3; ++s) {
      if ((int)d->a[A_GC + s][i] != 1) continue;
      int recov = 0;
      if ((int)d->a[A_GC_TX + s][i] == 1) {
        int k = at - (int)d->a[A_GC_TXT + s][i];
        if (k >= 1) recov = orc_bern(UA(XGC_S_AGENT2, i, 1 + s), p[txpr[s] + (k >= 3 ? 2 : k - 1)]);
      } else if (d->a[A_GC_INF + s][i] < (double)at) {
        if (d->c[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_GEOM) recov = orc_bern(UA(XGC_S_AGENT2, i, 1 + s), 1.0 / p[XGC_P_gc_ntx_int + s]);
        else recov = (double)at - d->a[A_GC_INF + s][i] >= d->a[A_GC_DUR + s][i];
      }
      if (recov) {
        d->a[A_GC + s][i] = 0; d->a[A_GC_SYM + s][i] = 0; d->a[A_GC_TX + s][i] = 0;
        d->a[A_GC_INF + s][i] = NA; d->a[A_GC_TXT + s][i] = NA; d->a[A_GC_DUR + s][i] = NA;
      }
    }
}

/* stitx_msm (sim1_02_lhs.r:216, 88-105, 157-158, 223, 225; Appendix E "stitx"; ambiguities A3-A5) */
static void stitx_msm(orc_dat* d, int at) {
  const double* p = d->p;
  uint32_t* c = CNT(d, at);
  int proto = d->c[XGC_C_stiScreeningProtocol], kissx = d->c[XGC_C_cdcExposureSite_Kissing];
  for (int i = 0; i < d->n; ++i) {
    int sv = 0, av = 0;
    for (int s = 0; s < 3; ++s)
      if ((int)d->a[A_GC + s][i] == 1 && (int)d->a[A_GC_SYM + s][i] == 1 && (int)d->a[A_GC_TX + s][i] == 0) {
        double rr = s == 0 ? p[XGC_P_rgc_sympt_seek_test_rr] : s == 2 ? p[XGC_P_pgc_sympt_seek_test_rr] : 1.0;
        if (orc_bern(UA(XGC_S_AGENT2, i, 8 + s), p[XGC_P_ugc_sympt_seek_test_prob] * rr)) sv = 1;
      }
    int ex = (int)d->a[A_EXPO][i];
    if (!sv) {
      if (proto == XGC_SCREEN_BASE || proto == XGC_SCREEN_UNIVERSAL) av = orc_bern(UA(XGC_S_AGENT2, i, 0), p[XGC_P_gc_asympt_visit_prob]);
      else if (proto == XGC_SCREEN_CDC) {
        int active = (ex & 7) != 0 || (kissx && (ex & 8));
        int hr = (int)d->a[A_PREPSTAT][i] == 1 || (ex & 16);
        double interval = (hr ? p[XGC_P_cdc_sti_hr_int] : p[XGC_P_cdc_sti_int]) * (52.0 / 12.0);
        int due = d->a[A_LASTSCR][i] == NA || (double)at - d->a[A_LASTSCR][i] >= interval;
        av = active && due;
      }
    }
    if (!(sv || av)) continue;
    c[XGC_K_VISITS]++; if (sv) c[XGC_K_VISITS_SYMPT]++;
    for (int s = 0; s < 3; ++s) {
      int inf = (int)d->a[A_GC + s][i] == 1, symp = inf && (int)d->a[A_GC_SYM + s][i] == 1, tested;
      int exposed = (ex >> s) & 1; if (s == 2 && kissx && (ex & 8)) exposed = 1;
      if (proto == XGC_SCREEN_CDC && !symp && !exposed) tested = 0;
      else if (proto == XGC_SCREEN_UNIVERSAL) tested = 1;
      else tested = orc_bern(UA(XGC_S_AGENT2, i, 12 + s), symp ? p[XGC_P_gc_sympt_prob_test + s] : p[XGC_P_gc_asympt_prob_test + s]);
      if (!tested) continue;
      c[XGC_K_TESTED + s]++;
      if (inf) {
        c[XGC_K_POS + s]++; d->a[A_INDLAST][i] = at;
        if ((int)d->a[A_GC_TX + s][i] == 0) { d->a[A_GC_TX + s][i] = 1; d->a[A_GC_TXT + s][i] = at; c[XGC_K_TXSTART + s]++; }
      }
    }
    d->a[A_LASTSCR][i] = at; d->a[A_EXPO][i] = 0;
  }
}

/* stitrans_msm_rand (sim1_02_lhs.r:217, 67-75, 84-86, 163-164, 220-221; Appendix E "stitrans"; ambiguity A2) */
static const int PATH_SITE[XGC_NPATH] = { 0, 0, 1, 1, 2, 2, 2 };       /* destination site of each pathway */
static void stitrans_msm_rand(orc_dat* d, int at) {
  ensure_al(d, at);
  const double* p = d->p;
  memset(d->newmask, 0, sizeof(uint32_t) * (size_t)d->n);
  /* Anal acts: one draw per act (condom use and position are per-act attributes shared with hivtrans and the PrEP risk
     history).  Oral, rimming and kissing acts carry no per-act attribute besides the direction, site status is constant
     within the week and only "any transmission" is observable, so the m acts of one (edge, type, direction) are ONE
     Bernoulli(1 - (1-t)^m) trial; the act list supplies m. */
  int* cnt[3];                       /* per layer: [e][type 1..3][direction p1 active / p2 active] */
  for (int l = 0; l < 3; ++l) cnt[l] = (int*)calloc((size_t)(d->ne[l] + 1) * 6, sizeof(int));
  const int per_act = d->inj && d->inj->u;          /* injected draws: every oral / rimming / kissing act is its own Bernoulli trial */
  for (int j = 0; j < d->nal; ++j) {
    const orc_act* r = &d->al[j];
    if (r->type != 0 && per_act) {
      const orc_edge* g = &d->el[r->layer][r->e];
      int base = XGC_S_EDGE0 + r->layer * XGC_EDGE_STREAMS;
      uint32_t ua = (uint32_t)d->a[A_UID][g->p1], ub = (uint32_t)d->a[A_UID][g->p2];
      if (r->type == 2) {                                      /* kissing act k: pharynx <-> pharynx */
        if (!d->c[XGC_C_transRoute_Kissing]) continue;
        int ga = (int)d->a[A_GC + 2][g->p1] == 1, gb = (int)d->a[A_GC + 2][g->p2] == 1;
        if (ga == gb) continue;
        if (orc_bern(U(d, at, base + (XGC_S_KISS - XGC_S_EDGE0), ua, ub, (size_t)r->e, r->k), p[XGC_P_p2pgc_tprob]))
          d->newmask[ga ? g->p2 : g->p1] |= 1u << XGC_PATH_P2P;
        continue;
      }
      if (r->type == 3 && !d->c[XGC_C_transRoute_Rimming]) continue;
      int sx = r->type == 1 ? 1 : 2, sy = r->type == 1 ? 2 : 0, st = r->type == 1 ? XGC_S_ORAL : XGC_S_RIM;
      int pathA = r->type == 1 ? XGC_PATH_U2P : XGC_PATH_P2R, pathB = r->type == 1 ? XGC_PATH_P2U : XGC_PATH_R2P;
      double tA = r->type == 1 ? p[XGC_P_u2pgc_tprob] : p[XGC_P_p2rgc_tprob], tB = r->type == 1 ? p[XGC_P_p2ugc_tprob] : p[XGC_P_r2pgc_tprob];
      int x = r->p1act ? g->p1 : g->p2, y = r->p1act ? g->p2 : g->p1;
      int gx = (int)d->a[A_GC + sx][x] == 1, gy = (int)d->a[A_GC + sy][y] == 1;
      if (gx == gy) continue;
      if (orc_bern(U(d, at, base + (st - XGC_S_EDGE0), ua, ub, (size_t)r->e, 2 + 2 * r->k), gx ? tA : tB))
        d->newmask[gx ? y : x] |= 1u << (gx ? pathA : pathB);
      continue;
    }
    if (r->type != 0) { cnt[r->layer][(size_t)r->e * 6 + (r->type - 1) * 2 + (r->type == 2 ? 0 : (r->p1act ? 0 : 1))]++; continue; }
    const orc_edge* g = &d->el[r->layer][r->e];
    int base = XGC_S_EDGE0 + r->layer * XGC_EDGE_STREAMS;
    uint32_t ua = (uint32_t)d->a[A_UID][g->p1], ub = (uint32_t)d->a[A_UID][g->p2];
    /* x = insertive partner, y = receptive; A: x's urethra -> y's rectum; B: reverse.  A and B exclude each other
       (each needs exactly one of the two sites infected), so they share one draw per act. */
    int x = r->p1act ? g->p1 : g->p2, y = x == g->p1 ? g->p2 : g->p1;
    int gx = (int)d->a[A_GC + 1][x] == 1, gy = (int)d->a[A_GC + 0][y] == 1;
    int condom = !r->uai, kk = 4 * r->k + 3;
    if (gx && !gy && d->a[A_GC_INF + 1][x] < (double)at) {
      double t = p[XGC_P_u2rgc_tprob];
      if (condom) t = t * (1.0 - p[XGC_P_sti_cond_eff] * (1.0 - p[XGC_P_sti_cond_fail + (int)d->a[A_RACE][y] - 1]));
      if (orc_bern(U(d, at, base + (XGC_S_ANAL - XGC_S_EDGE0), ua, ub, (size_t)r->e, kk), t)) d->newmask[y] |= 1u << XGC_PATH_U2R;
    }
    if (gy && !gx && d->a[A_GC_INF + 0][y] < (double)at) {
      double t = p[XGC_P_r2ugc_tprob];
      if (condom) t = t * (1.0 - p[XGC_P_sti_cond_eff] * (1.0 - p[XGC_P_sti_cond_fail + (int)d->a[A_RACE][x] - 1]));
      if (orc_bern(U(d, at, base + (XGC_S_ANAL - XGC_S_EDGE0), ua, ub, (size_t)r->e, kk), t)) d->newmask[x] |= 1u << XGC_PATH_R2U;
    }
  }
  for (int l = 0; l < 3; ++l) {
    int base = XGC_S_EDGE0 + l * XGC_EDGE_STREAMS;
    for (int e = 0; e < d->ne[l]; ++e) {
      const orc_edge* g = &d->el[l][e];
      const int* c = cnt[l] + (size_t)e * 6;
      uint32_t ua = (uint32_t)d->a[A_UID][g->p1], ub = (uint32_t)d->a[A_UID][g->p2];
      for (int type = 1; type <= 3; ++type) {
        if (type == 2) {                                     /* kissing: pharynx <-> pharynx, no direction */
          int m = c[2];
          if (!m || !d->c[XGC_C_transRoute_Kissing]) continue;
          int ga = (int)d->a[A_GC + 2][g->p1] == 1, gb = (int)d->a[A_GC + 2][g->p2] == 1;
          if (ga == gb) continue;
          if (orc_bern(U(d, at, base + (XGC_S_KISS - XGC_S_EDGE0), ua, ub, (size_t)e, 0), xgc_any_of_n(p[XGC_P_p2pgc_tprob], m)))
            d->newmask[ga ? g->p2 : g->p1] |= 1u << XGC_PATH_P2P;
          continue;
        }
        if (type == 3 && !d->c[XGC_C_transRoute_Rimming]) continue;
        int sx = type == 1 ? 1 : 2, sy = type == 1 ? 2 : 0, st = type == 1 ? XGC_S_ORAL : XGC_S_RIM;
        int pathA = type == 1 ? XGC_PATH_U2P : XGC_PATH_P2R, pathB = type == 1 ? XGC_PATH_P2U : XGC_PATH_R2P;
        double tA = type == 1 ? p[XGC_P_u2pgc_tprob] : p[XGC_P_p2rgc_tprob], tB = type == 1 ? p[XGC_P_p2ugc_tprob] : p[XGC_P_r2pgc_tprob];
        for (int dir = 0; dir < 2; ++dir) {                  /* dir 0: p1 active (draw k1), dir 1: p2 active (draw k2) */
          int m = c[(type - 1) * 2 + dir];
          if (!m) continue;
          int x = dir == 0 ? g->p1 : g->p2, y = dir == 0 ? g->p2 : g->p1;
          int gx = (int)d->a[A_GC + sx][x] == 1, gy = (int)d->a[A_GC + sy][y] == 1;
          if (gx == gy) continue;
          if (orc_bern(U(d, at, base + (st - XGC_S_EDGE0), ua, ub, (size_t)e, 1 + dir), xgc_any_of_n(gx ? tA : tB, m)))
            d->newmask[gx ? y : x] |= 1u << (gx ? pathA : pathB);
        }
      }
    }
  }
  for (int l = 0; l < 3; ++l) free(cnt[l]);
  /* "_rand": weekly random order of the 7 pathways decides which pathway a multiply-hit site is attributed to */
  int perm[XGC_NPATH], rank[XGC_NPATH];
  for (int k = 0; k < XGC_NPATH; ++k) perm[k] = k;
  for (int i = XGC_NPATH - 1; i >= 1; --i) {
    int j = (int)(UR(XGC_S_PATHORDER, XGC_NPATH - 1 - i) * (double)(i + 1)); if (j > i) j = i;
    int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  for (int k = 0; k < XGC_NPATH; ++k) rank[perm[k]] = k;
  uint32_t* c = CNT(d, at);
  for (int i = 0; i < d->n; ++i) {
    uint32_t m = d->newmask[i] & 0x7Fu;
    if (!m) continue;
    for (int s = 0; s < 3; ++s) {
      int win = -1;
      for (int k = 0; k < XGC_NPATH; ++k) if (PATH_SITE[k] == s && (m >> k & 1u) && (win < 0 || rank[k] < rank[win])) win = k;
      if (win < 0 || (int)d->a[A_GC + s][i] == 1) continue;
      d->a[A_GC + s][i] = 1; d->a[A_GC_INF + s][i] = at; d->a[A_GC_TX + s][i] = 0; d->a[A_GC_TXT + s][i] = NA;
      d->a[A_GC_SYM + s][i] = orc_bern(UA(XGC_S_INFECT, i, s), p[XGC_P_gc_sympt_prob + s]);
      if (d->c[XGC_C_gcUntreatedRecovDist] == XGC_RECOV_POIS) {
        int du = orc_pois(UA(XGC_S_INFECT, i, 3 + s), p[XGC_P_gc_ntx_int + s], 255);
        d->a[A_GC_DUR + s][i] = du < 1 ? 1 : du;
      }
      c[(XGC_K_INCID_RGC + s) * XGC_NCELL + cell_of(d, i)]++;
      c[XGC_K_PATH + win]++;
    }
  }
}

/* prevalence_msm (sim1_02_lhs.r:218) */
static void prevalence_msm(orc_dat* d, int at) {
  uint32_t* c = CNT(d, at);
  for (int i = 0; i < d->n; ++i) {
    int cell = cell_of(d, i);
    c[XGC_K_NUM * XGC_NCELL + cell]++;
    if ((int)d->a[A_STATUS][i] == 1) c[XGC_K_INUM * XGC_NCELL + cell]++;
    if ((int)d->a[A_DIAG][i] == 1) {
      c[XGC_K_INUM_DX * XGC_NCELL + cell]++;
      if (d->a[A_VL][i] <= LOG10_200) c[XGC_K_VSUPP * XGC_NCELL + cell]++;
    }
    if ((int)d->a[A_PREPELIG][i] == 1) c[XGC_K_PREPELIG * XGC_NCELL + cell]++;
    if ((int)d->a[A_PREPSTAT][i] == 1) c[XGC_K_PREPCURR * XGC_NCELL + cell]++;
    int any = 0;
    for (int s = 0; s < 3; ++s) if ((int)d->a[A_GC + s][i] == 1) { any = 1; c[(XGC_K_INUM_RGC + s) * XGC_NCELL + cell]++; }
    if (any) c[XGC_K_INUM_GC * XGC_NCELL + cell]++;
  }
}

/* netsim's weekly module loop for ONE week (Appendix E "Loop/order"), in the order the reference writes the modules */
void orc_step(orc_dat* d, int at, uint32_t mask, const xgc_inject* inj) {
  d->inj = inj;
  if (at < 0 || at >= d->t_cap) { d->inj = NULL; return; }
  if (mask & XGC_M_AGING) aging_msm(d, at);
  if (mask & XGC_M_DEPARTURE) departure_msm(d, at);
  if (mask & XGC_M_ARRIVAL) arrival_msm(d, at);
  if (mask & XGC_M_HIVTEST) hivtest_msm(d, at);
  if (mask & XGC_M_HIVTX) hivtx_msm(d, at);
  if (mask & XGC_M_HIVPROGRESS) hivprogress_msm(d, at);
  if (mask & XGC_M_HIVVL) hivvl_msm(d, at);
  if ((mask & XGC_M_RESIM_NETS) && d->c[XGC_C_network_mode] == 1) resim_nets_surrogate(d, at);
  if (mask & XGC_M_ACTS) acts_msm(d, at);
  if (mask & XGC_M_CONDOMS) condoms_msm(d, at);
  if (mask & XGC_M_POSITION) { if (d->al_at != at) condoms_msm(d, at); position_msm(d, at); }
  if (mask & XGC_M_PREP) { riskhist_msm(d, at); prep_msm(d, at); }
  if (mask & XGC_M_HIVTRANS) hivtrans_msm(d, at);
  if (mask & XGC_M_STIRECOV) stirecov_msm(d, at);
  if (mask & XGC_M_STITX) stitx_msm(d, at);
  if (mask & XGC_M_STITRANS) stitrans_msm_rand(d, at);
  if (mask & XGC_M_PREV) prevalence_msm(d, at);
  d->inj = NULL;
}

void orc_get_counters(const orc_dat* d, int at, uint32_t* out) {
  memcpy(out, d->cnt + (size_t)at * XGC_NCOUNTER, sizeof(uint32_t) * XGC_NCOUNTER);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* epi series (dat$epi$<name>[at]) from the weekly tallies; names as used downstream              */
/* ------------------------------------------------------------------------------------------------------------------ */
static double gsum(const uint32_t* c, int group, uint32_t cells) {
  double s = 0.0;
  for (int k = 0; k < XGC_NCELL; ++k) if (cells >> k & 1u) s += (double)c[group * XGC_NCELL + k];
  return s;
}
/* parse a stratum token: "" all; B/H/O/W; age1 / age.1 / 1; B.age1.  Returns cell mask or 0 if not a stratum. */
static uint32_t stratum_cells(const char* s) {
  if (!*s) return 0xFFFFFu;
  int race = -1, age = -1;
  const char* q = s;
  if (strchr("BHOW", q[0]) && (q[1] == 0 || q[1] == '.')) { race = (int)(strchr("BHOW", q[0]) - "BHOW"); q += q[1] ? 2 : 1; }
  if (*q) {
    if (strncmp(q, "age.", 4) == 0) q += 4; else if (strncmp(q, "age", 3) == 0) q += 3;
    if (q[0] >= '1' && q[0] <= '5' && q[1] == 0) age = q[0] - '1'; else return 0;
  }
  uint32_t m = 0;
  for (int r = 0; r < 4; ++r) for (int a = 0; a < 5; ++a) if ((race < 0 || race == r) && (age < 0 || age == a)) m |= 1u << (r * 5 + a);
  return m;
}
static int starts(const char* s, const char* pre, const char** rest) {
  size_t n = strlen(pre);
  if (strncmp(s, pre, n) != 0) return 0;
  if (s[n] == 0) { *rest = s + n; return 1; }
  if (s[n] == '.') { *rest = s + n + 1; return 1; }
  return 0;
}
static double epi_value(const uint32_t* c, const char* var, int* ok) {
  const char* rest; uint32_t cells; *ok = 1;
  static const char* const sites[3] = { "rgc", "ugc", "pgc" };
  static const char* const pathn[XGC_NPATH] = { "incid.u2rgc", "incid.p2rgc", "incid.r2ugc", "incid.p2ugc", "incid.u2pgc", "incid.r2pgc", "incid.p2pgc" };
  static const char* const tst[3] = { "num.rgc.test", "num.ugc.test", "num.pgc.test" };
  static const char* const propt[3] = { "prop.rect.tested", "prop.ureth.tested", "prop.phar.tested" };
  static const char* const probt[3] = { "prob.rGC.tested", "prob.uGC.tested", "prob.pGC.tested" };
  for (int k = 0; k < XGC_NPATH; ++k) if (!strcmp(var, pathn[k])) return (double)c[XGC_K_PATH + k];
  for (int s = 0; s < 3; ++s) {
    if (!strcmp(var, tst[s])) return (double)c[XGC_K_TESTED + s];
    if (!strcmp(var, propt[s])) return (double)c[XGC_K_TESTED + s] / (double)c[XGC_K_VISITS];
    if (!strcmp(var, probt[s])) return (double)c[XGC_K_POS + s] / (double)c[XGC_K_TESTED + s];
  }
  if (!strcmp(var, "num.visits")) return (double)c[XGC_K_VISITS];
  if (!strcmp(var, "dep.gen")) return (double)c[XGC_K_DEP_GEN];
  if (!strcmp(var, "dep.AIDS")) return (double)c[XGC_K_DEP_AIDS];
  if (!strcmp(var, "nNew")) return (double)c[XGC_K_NNEW];
  /* GC series, optionally stratified in the middle: incid.B.rgc, incid.age.1.rgc */
  {
    size_t L = strlen(var);
    for (int s = -1; s < 3; ++s) {
      const char* suf = s < 0 ? "gc" : sites[s];
      size_t ls = strlen(suf);
      if (L < ls + 1 || strcmp(var + L - ls, suf) != 0 || var[L - ls - 1] != '.') continue;
      char head[64]; if (L - ls - 1 >= sizeof(head)) continue;
      memcpy(head, var, L - ls - 1); head[L - ls - 1] = 0;
      int kind = -1;
      if (starts(head, "i.num", &rest)) kind = 0; else if (starts(head, "prev", &rest)) kind = 1;
      else if (starts(head, "ir100", &rest)) kind = 3; else if (starts(head, "incid", &rest)) kind = 2;
      if (kind < 0 || !(cells = stratum_cells(rest))) continue;
      double num = gsum(c, XGC_K_NUM, cells), inum, incid;
      if (s < 0) { inum = gsum(c, XGC_K_INUM_GC, cells); incid = gsum(c, XGC_K_INCID_RGC, cells) + gsum(c, XGC_K_INCID_UGC, cells) + gsum(c, XGC_K_INCID_PGC, cells); }
      else { inum = gsum(c, XGC_K_INUM_RGC + s, cells); incid = gsum(c, XGC_K_INCID_RGC + s, cells); }
      if (kind == 0) return inum; if (kind == 1) return inum / num; if (kind == 2) return incid;
      return incid / (num - inum) * 5200.0;
    }
  }
  if (starts(var, "num", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_NUM, cells);
  if (starts(var, "i.num.dx", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_INUM_DX, cells);
  if (starts(var, "i.num", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_INUM, cells);
  if (starts(var, "i.prev.dx", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_INUM_DX, cells) / gsum(c, XGC_K_NUM, cells);
  if (starts(var, "i.prev", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_INUM, cells) / gsum(c, XGC_K_NUM, cells);
  if (starts(var, "incid", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_INCID_HIV, cells);
  if (starts(var, "cc.vsupp", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_VSUPP, cells) / gsum(c, XGC_K_INUM_DX, cells);
  if (starts(var, "ir100", &rest) && (cells = stratum_cells(rest)))
    return gsum(c, XGC_K_INCID_HIV, cells) / (gsum(c, XGC_K_NUM, cells) - gsum(c, XGC_K_INUM, cells)) * 5200.0;
  if (starts(var, "prepElig", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_PREPELIG, cells);
  if (starts(var, "prepCurr", &rest) && (cells = stratum_cells(rest))) return gsum(c, XGC_K_PREPCURR, cells);
  *ok = 0;
  return 0.0;
}
int orc_get_epi(const orc_dat* d, const char* var, double* out, int at0, int at1) {
  for (int at = at0; at <= at1; ++at) {
    int ok; out[at - at0] = epi_value(d->cnt + (size_t)at * XGC_NCOUNTER, var, &ok);
    if (!ok) return -1;
  }
  return 0;
}

/* tail-window targets: weekly ratio, NA/NaN -> 0, then mean over the last tailn weeks
 * (inst/cal/sim0.0_setup.r:205-224, 272, 299-308; R/calculate_targets.r:14-190) */
static double nz(double x) { return (x != x || x > 1.7e308 || x < -1.7e308) ? 0.0 : x; }
void orc_targets(const orc_dat* d, int at_last, int tailn, double* out) {
  for (int k = 0; k < XGC_NTARGET; ++k) out[k] = 0.0;
  for (int at = at_last - tailn + 1; at <= at_last; ++at) {
    const uint32_t* c = d->cnt + (size_t)at * XGC_NCOUNTER;
    double w[XGC_NTARGET];
    double num = gsum(c, XGC_K_NUM, 0xFFFFFu);
    w[XGC_TG_NUM] = num;
    w[XGC_TG_I_PREV] = gsum(c, XGC_K_INUM, 0xFFFFFu) / num;
    w[XGC_TG_IR100_POP] = gsum(c, XGC_K_INCID_HIV, 0xFFFFFu) / num * 5200.0;
    for (int r = 0; r < 4; ++r) {
      uint32_t m = 0x1Fu << (r * 5);
      double n = gsum(c, XGC_K_NUM, m);
      w[XGC_TG_HIVDX_RACE + r] = (gsum(c, XGC_K_INUM_DX, m) / n) / (gsum(c, XGC_K_INUM, m) / n);
      w[XGC_TG_VLS_RACE + r] = gsum(c, XGC_K_VSUPP, m) / gsum(c, XGC_K_INUM_DX, m);
      w[XGC_TG_PREPCOV_RACE + r] = gsum(c, XGC_K_PREPCURR, m) / gsum(c, XGC_K_PREPELIG, m);
    }
    for (int a = 0; a < 5; ++a) {
      uint32_t m = 0x08421u << a;
      double n = gsum(c, XGC_K_NUM, m);
      w[XGC_TG_HIVDX_AGE + a] = (gsum(c, XGC_K_INUM_DX, m) / n) / (gsum(c, XGC_K_INUM, m) / n);
      w[XGC_TG_VLS_AGE + a] = gsum(c, XGC_K_VSUPP, m) / gsum(c, XGC_K_INUM_DX, m);
    }
    for (int s = 0; s < 3; ++s) {
      w[XGC_TG_PROP_TESTED + s] = (double)c[XGC_K_TESTED + s] / (double)c[XGC_K_VISITS];
      w[XGC_TG_PROB_POS + s] = (double)c[XGC_K_POS + s] / (double)c[XGC_K_TESTED + s];
    }
    double inc_all = 0.0;
    for (int s = 0; s < 3; ++s) {
      double inum = gsum(c, XGC_K_INUM_RGC + s, 0xFFFFFu), inc = gsum(c, XGC_K_INCID_RGC + s, 0xFFFFFu);
      w[XGC_TG_PREV_GC + 1 + s] = inum / num;
      w[XGC_TG_IR100_GC + 1 + s] = inc / (num - inum) * 5200.0;
      inc_all += inc;
    }
    double any = gsum(c, XGC_K_INUM_GC, 0xFFFFFu);
    w[XGC_TG_PREV_GC] = any / num;
    w[XGC_TG_IR100_GC] = inc_all / (num - any) * 5200.0;
    {                                  /* target_prob_agedx_byrace (R/calculate_targets.r:70-105) */
      static const int cells[8] = XGC_TG_AGEDX_CELLS;
      for (int j = 0; j < 8; ++j)
        w[XGC_TG_AGEDX + j] = (double)c[XGC_K_INUM_DX * XGC_NCELL + cells[j]] / gsum(c, XGC_K_INUM_DX, 0x1Fu << ((cells[j] / 5) * 5));
    }
    for (int k = 0; k < XGC_NTARGET; ++k) out[k] += nz(w[k]);
  }
  for (int k = 0; k < XGC_NTARGET; ++k) out[k] = out[k] / (double)tailn;
}

/* constants of include/xgc/abi.h for the ctypes test harness */
int orc_constant(const char* name, int64_t* value) {
#define X(c) if (!strcmp(name, #c)) { *value = (int64_t)(c); return 0; }
  XGC_CONSTANT_LIST(X)
#undef X
  return 1;
}
