/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the weekly epidemic step of the EpiModel-style gonorrhea/HIV model that
 * jrgant/xgc-msm drives.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or execute this code; the CUDA product path never does.
 *
 * PARITY UNPINNED.  The module bodies the reference runs live in the un-vendored EpiModelHIV 1.5.0 fork
 * (/root/reference/renv.lock:26-31, "Source": "unknown"), EpiModel 1.8.0 (renv.lock:19-24) and tergmLite 2.2.1
 * (renv.lock:839); none is present under /root/reference, R is not installed, the fitted inputs est/NAME.Rds are git-LFS
 * pointers, and the reference holds no tests or golden vectors for this path.  This file
 * therefore restates the published EpiModelHIV-p module semantics as recorded in SURVEY.md Appendix E, constrained
 * by the reference's own call sites: module list and order (burnin/cal/sim1_02_lhs.r:200-218), every parameter's
 * meaning (sim1_02_lhs.r:58-178; sim1_01_lhs_params.r:22-102), the control flags (sim1_02_lhs.r:219-229), the epi
 * output catalogue (inst/cal/sim_clean.r:67-89; inst/analysis02_screening/98_summarize_sims.r:75-115) and the target
 * definitions (R/calculate_targets.r:14-190; inst/cal/sim0.0_setup.r:205-308).  Every decision the absent source
 * forces is listed in DESIGN.md ("ambiguity register").  What CAN be pinned is pinned in tests/: the calibration
 * targets recomputed from inst/cal/calval_targets.r literals, Philox4x32-10 known-answer vectors, and exp() vs libm.
 *
 * Style: one replicate at a time, R-like: attribute vectors in position order, modules called one after another,
 * departures physically delete rows and renumber edge lists (tergmLite::delete_vertices semantics), the act list is
 * materialised.  Deliberately NOT the fused/packed structure of the CUDA path.
 */
#ifndef XGC_ORACLE_H
#define XGC_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include "xgc/abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_dat orc_dat;

orc_dat* orc_create(int cap, const int e_cap[3], int t_cap, uint64_t seed, int replicate_id);
void     orc_destroy(orc_dat* d);
void     orc_set_params(orc_dat* d, const double* p);
void     orc_set_control(orc_dat* d, const int32_t* c);
void     orc_set_tables(orc_dat* d, const double* t);
void     orc_set_population(orc_dat* d, int n, int at);
int      orc_num_agents(const orc_dat* d);
int      orc_set_attr(orc_dat* d, const char* name, const double* v, int n);
int      orc_get_attr(const orc_dat* d, const char* name, double* v, int cap);
int      orc_set_edgelist(orc_dat* d, int layer, const int32_t* el, int n_edges, const int32_t* start);
int      orc_get_edgelist(const orc_dat* d, int layer, int32_t* el, int cap, int32_t* start);
int      orc_get_actcounts(const orc_dat* d, int layer, int32_t* counts, int cap);
int      orc_get_actlist(orc_dat* d, int at, int32_t* al, int cap, const xgc_inject* inj);
/* run the selected modules for week `at` in reference order */
void     orc_step(orc_dat* d, int at, uint32_t mask, const xgc_inject* inj);
void     orc_get_counters(const orc_dat* d, int at, uint32_t* out);
int      orc_get_epi(const orc_dat* d, const char* var, double* out, int at0, int at1);
void     orc_targets(const orc_dat* d, int at_last, int tailn, double* out);
uint32_t orc_status(const orc_dat* d);

/* building blocks exposed for unit tests */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double   orc_exp(double x);
double   orc_any_of_n(double t, int n);    /* xgc_any_of_n of include/xgc/abi.h */
int      orc_bern(double u, double p);
int      orc_pois(double u, double mu, int kmax);
double   orc_uniform(const orc_dat* d, int at, int stream, uint32_t e0, uint32_t e1, int k);

#ifdef __cplusplus
}
#endif
#endif
