/* .Call shims between R and libxgc.so (include/xgc/abi.h).  NOT compiled in this image (no R, no Rinternals.h);
 * build where R exists:  R CMD SHLIB rshim/xgc_r.c -Iinclude -Lxgc_msm_b200 -lxgc
 * Ownership: R owns every SEXP; the library copies on upload; the handle lives behind an external pointer whose
 * finalizer calls xgc_destroy.  Errors: non-zero status -> Rf_error() after temporaries are released. */
#include <R.h>
#include <Rinternals.h>
#include <string.h>
#include <stdint.h>
#include "xgc/abi.h"

static void fin(SEXP p) { xgc_handle* h = (xgc_handle*)R_ExternalPtrAddr(p); if (h) { xgc_destroy(h); R_ClearExternalPtr(p); } }
static xgc_handle* H(SEXP p) { xgc_handle* h = (xgc_handle*)R_ExternalPtrAddr(p); if (!h) Rf_error("xgc: handle is closed"); return h; }
#define CK(h, call) do { int _rc = (call); if (_rc) Rf_error("xgc error %d: %s", _rc, xgc_last_error(h)); } while (0)

SEXP xgcR_create(SEXP nrep, SEXP ncap, SEXP ecap, SEXP tcap, SEXP seed, SEXP device, SEXP rep_offset) {
  xgc_config c; memset(&c, 0, sizeof c);
  c.abi_version = XGC_ABI_VERSION; c.device = Rf_asInteger(device); c.n_replicates = Rf_asInteger(nrep);
  c.replicate_offset = Rf_asInteger(rep_offset); c.n_cap = Rf_asInteger(ncap); c.t_cap = Rf_asInteger(tcap);
  for (int l = 0; l < 3; ++l) c.e_cap[l] = INTEGER(ecap)[l];
  c.seed = (uint64_t)Rf_asReal(seed);
  xgc_handle* h = NULL;
  int rc = xgc_create(&c, &h);
  if (rc) Rf_error("xgc_create failed (%d): %s", rc, xgc_last_error(NULL));
  SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, fin, TRUE);
  UNPROTECT(1);
  return p;
}
SEXP xgcR_set_params(SEXP p, SEXP r, SEXP v) { if (XLENGTH(v) != XGC_NPARAM) Rf_error("params must have %d entries", XGC_NPARAM);
  CK(H(p), xgc_set_params(H(p), Rf_asInteger(r), REAL(v))); return R_NilValue; }
SEXP xgcR_get_params(SEXP p, SEXP r) { SEXP out = PROTECT(Rf_allocVector(REALSXP, XGC_NPARAM)); int rc = xgc_get_params(H(p), Rf_asInteger(r), REAL(out));
  UNPROTECT(1); if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p))); return out; }
SEXP xgcR_set_control(SEXP p, SEXP r, SEXP v) { CK(H(p), xgc_set_control(H(p), Rf_asInteger(r), INTEGER(v))); return R_NilValue; }
SEXP xgcR_set_tables(SEXP p, SEXP v) { CK(H(p), xgc_set_tables(H(p), REAL(v))); return R_NilValue; }
SEXP xgcR_set_population(SEXP p, SEXP r, SEXP n, SEXP at) { CK(H(p), xgc_set_population(H(p), Rf_asInteger(r), Rf_asInteger(n), Rf_asInteger(at))); return R_NilValue; }
SEXP xgcR_upload_attr(SEXP p, SEXP r, SEXP name, SEXP v) {            /* dat$attr[[name]] -> device */
  SEXP d = PROTECT(Rf_coerceVector(v, REALSXP));
  int rc = xgc_upload_attr(H(p), Rf_asInteger(r), CHAR(STRING_ELT(name, 0)), REAL(d), (size_t)XLENGTH(d), XGC_F64);
  UNPROTECT(1);
  if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p)));
  return R_NilValue;
}
SEXP xgcR_download_attr(SEXP p, SEXP r, SEXP name) {                  /* device -> dat$attr[[name]] */
  int n = 0; CK(H(p), xgc_num_agents(H(p), Rf_asInteger(r), &n));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n)); size_t got = 0;
  int rc = xgc_download_attr(H(p), Rf_asInteger(r), CHAR(STRING_ELT(name, 0)), REAL(out), (size_t)n, XGC_F64, &got);
  UNPROTECT(1);
  if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p)));
  return out;
}
SEXP xgcR_set_edgelist(SEXP p, SEXP r, SEXP layer, SEXP el, SEXP start) {   /* dat$el[[layer]]: integer matrix [E,2], column-major */
  int E = Rf_nrows(el);
  CK(H(p), xgc_set_edgelist(H(p), Rf_asInteger(r), Rf_asInteger(layer) - 1, INTEGER(el), E, Rf_isNull(start) ? NULL : INTEGER(start)));
  return R_NilValue;
}
SEXP xgcR_get_edgelist(SEXP p, SEXP r, SEXP layer, SEXP cap) {
  int c = Rf_asInteger(cap), ne = 0;
  int* buf = (int*)R_alloc((size_t)2 * c, sizeof(int));
  CK(H(p), xgc_get_edgelist(H(p), Rf_asInteger(r), Rf_asInteger(layer) - 1, buf, c, &ne, NULL));
  SEXP out = PROTECT(Rf_allocMatrix(INTSXP, ne, 2));
  memcpy(INTEGER(out), buf, sizeof(int) * 2 * (size_t)ne);
  UNPROTECT(1);
  return out;
}
SEXP xgcR_step(SEXP p, SEXP at, SEXP mask, SEXP u) {                  /* dat <- FUN(dat, at): u = NULL or the injected runif() table */
  if (Rf_isNull(u)) { CK(H(p), xgc_step(H(p), Rf_asInteger(at), (uint32_t)Rf_asReal(mask), NULL)); return R_NilValue; }
  SEXP tab = VECTOR_ELT(u, 0), dims = VECTOR_ELT(u, 1);               /* list(table, c(uid_cap, e_cap[3], form_cap)) */
  xgc_inject inj; memset(&inj, 0, sizeof inj);
  inj.u = REAL(tab); inj.uid_cap = INTEGER(dims)[0]; for (int l = 0; l < 3; ++l) inj.e_cap[l] = INTEGER(dims)[1 + l]; inj.form_cap = INTEGER(dims)[4];
  inj.stride = xgc_inject_size(&inj);
  CK(H(p), xgc_step(H(p), Rf_asInteger(at), (uint32_t)Rf_asReal(mask), &inj));
  return R_NilValue;
}
SEXP xgcR_step_span(SEXP p, SEXP at, SEXP mask_tail, SEXP mask_head) {  /* rest of week at + start of week at + 1: one resident launch in fast mode */
  CK(H(p), xgc_step_span(H(p), Rf_asInteger(at), (uint32_t)Rf_asReal(mask_tail), (uint32_t)Rf_asReal(mask_head)));
  return R_NilValue;
}
/* One host-network week in one .Call: list(delta_main, delta_casl, el_inst) as integer vectors / matrix or NULL (one replicate per R
 * process, the reference's layout); returns list(counters of week at, living agents after the aging..hivvl group of week at + 1). */
SEXP xgcR_week_hostnet(SEXP p, SEXP at, SEXP mask_tail, SEXP mask_head, SEXP delta_main, SEXP delta_casl, SEXP has_start, SEXP el_inst) {
  int ne = Rf_isNull(el_inst) ? 0 : Rf_nrows(el_inst), n_alive = 0;
  SEXP cnt = PROTECT(Rf_allocVector(INTSXP, XGC_NCOUNTER));
  int rc = xgc_week_hostnet(H(p), Rf_asInteger(at), (uint32_t)Rf_asReal(mask_tail), (uint32_t)Rf_asReal(mask_head),
                            Rf_isNull(delta_main) ? NULL : INTEGER(delta_main), Rf_isNull(delta_main) ? 0 : LENGTH(delta_main),
                            Rf_isNull(delta_casl) ? NULL : INTEGER(delta_casl), Rf_isNull(delta_casl) ? 0 : LENGTH(delta_casl), Rf_asInteger(has_start),
                            Rf_isNull(el_inst) ? NULL : INTEGER(el_inst), Rf_isNull(el_inst) ? NULL : &ne, ne > 0 ? ne : 1,
                            (uint32_t*)INTEGER(cnt), &n_alive);
  if (rc) { UNPROTECT(1); Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p))); }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, cnt); SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(n_alive));
  UNPROTECT(2);
  return out;
}
SEXP xgcR_run(SEXP p, SEXP at0, SEXP at1) { CK(H(p), xgc_run(H(p), Rf_asInteger(at0), Rf_asInteger(at1))); return R_NilValue; }
SEXP xgcR_epi(SEXP p, SEXP r, SEXP var, SEXP at0, SEXP at1) {         /* dat$epi[[var]][at0:at1] */
  int a = Rf_asInteger(at0), b = Rf_asInteger(at1);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, b - a + 1));
  int rc = xgc_get_epi(H(p), Rf_asInteger(r), CHAR(STRING_ELT(var, 0)), REAL(out), a, b);
  UNPROTECT(1);
  if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p)));
  return out;
}
SEXP xgcR_epi_table(SEXP p, SEXP r, SEXP vars, SEXP at0, SEXP at1) { /* as.matrix(as.data.frame(dat$epi)[at0:at1, vars]): one device read */
  int a = Rf_asInteger(at0), b = Rf_asInteger(at1), n = LENGTH(vars), T = b - a + 1;
  const char** names = (const char**)R_alloc((size_t)(n > 0 ? n : 1), sizeof(char*));
  for (int k = 0; k < n; ++k) names[k] = CHAR(STRING_ELT(vars, k));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, T, n));                  /* column k = series k: the library's [n_vars][T] layout */
  int rc = xgc_get_epi_table(H(p), Rf_asInteger(r), names, n, REAL(out), a, b);
  UNPROTECT(1);
  if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p)));
  return out;
}
SEXP xgcR_targets(SEXP p, SEXP nrep, SEXP at_last, SEXP tailn) {      /* [nsims x XGC_NTARGET], row-major from the library */
  int R_ = Rf_asInteger(nrep);
  double* buf = (double*)R_alloc((size_t)R_ * XGC_NTARGET, sizeof(double));
  CK(H(p), xgc_targets(H(p), Rf_asInteger(at_last), Rf_asInteger(tailn), buf));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R_, XGC_NTARGET));
  for (int r = 0; r < R_; ++r) for (int k = 0; k < XGC_NTARGET; ++k) REAL(out)[r + (size_t)k * R_] = buf[(size_t)r * XGC_NTARGET + k];
  UNPROTECT(1);
  return out;
}
SEXP xgcR_attr_names(void) {                                          /* names xgc_upload_attr / xgc_download_attr know */
  const char* nm[256]; int n = xgc_attr_names(nm, 256);
  SEXP out = PROTECT(Rf_allocVector(STRSXP, n));
  for (int k = 0; k < n; ++k) SET_STRING_ELT(out, k, Rf_mkChar(nm[k]));
  UNPROTECT(1);
  return out;
}
SEXP xgcR_param_index(SEXP name) {                                    /* c(offset, count) of a params.def entry, 0-based offset; NULL if unknown */
  int off = 0, cnt = 0;
  if (xgc_param_index(CHAR(STRING_ELT(name, 0)), &off, &cnt)) return R_NilValue;
  SEXP out = PROTECT(Rf_allocVector(INTSXP, 2)); INTEGER(out)[0] = off; INTEGER(out)[1] = cnt; UNPROTECT(1);
  return out;
}
SEXP xgcR_constant(SEXP name) {                                       /* named constant of abi.h (XGC_NPARAM, XGC_C_..., XGC_SCREEN_...) */
  int64_t v = 0;
  if (xgc_constant(CHAR(STRING_ELT(name, 0)), &v)) Rf_error("xgc: unknown constant %s", CHAR(STRING_ELT(name, 0)));
  return Rf_ScalarInteger((int)v);
}
SEXP xgcR_num_agents(SEXP p, SEXP r) { int n = 0; CK(H(p), xgc_num_agents(H(p), Rf_asInteger(r), &n)); return Rf_ScalarInteger(n); }
SEXP xgcR_status(SEXP p, SEXP r) { uint32_t s = 0; CK(H(p), xgc_status(H(p), Rf_asInteger(r), &s)); return Rf_ScalarInteger((int)s); }
SEXP xgcR_set_mode(SEXP p, SEXP mode) { CK(H(p), xgc_set_mode(H(p), Rf_asInteger(mode))); return R_NilValue; }
SEXP xgcR_compact(SEXP p) { CK(H(p), xgc_compact(H(p))); return R_NilValue; }
SEXP xgcR_snapshot(SEXP p) {                                          /* the whole device state as a raw vector (saveRDS-able: the burn-in file) */
  size_t nb = 0; CK(H(p), xgc_snapshot_size(H(p), &nb));
  SEXP out = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)nb));
  int rc = xgc_snapshot(H(p), RAW(out), nb);
  UNPROTECT(1);
  if (rc) Rf_error("xgc error %d: %s", rc, xgc_last_error(H(p)));
  return out;
}
SEXP xgcR_restore(SEXP p, SEXP raw) { CK(H(p), xgc_restore(H(p), RAW(raw), (size_t)XLENGTH(raw))); return R_NilValue; }
SEXP xgcR_fork(SEXP src, SEXP src_r, SEXP dst, SEXP dst_r) {          /* device-to-device copy of one replicate (nsims > 1, scenario arms) */
  CK(H(src), xgc_fork(H(src), Rf_asInteger(src_r), H(dst), Rf_asInteger(dst_r))); return R_NilValue; }
