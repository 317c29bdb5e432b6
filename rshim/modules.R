# R wrappers with the reference's module contract `FUN(dat, at) -> dat` over the .Call shims of rshim/xgc_r.c.
# Drop-in use (burnin/cal/sim1_02_lhs.r:195-218):
#   control <- control_msm(..., initialize.FUN = initialize_msm_gpu, aging.FUN = aging_msm_gpu, ...,
#                          resim_nets.FUN = simnet_msm_gpu, ..., prev.FUN = prevalence_msm_gpu)
# State lives on the device behind dat$xgc (external pointer); dat$attr / dat$el are materialised only around the host
# network step, which is the one module that stays in R (tergmLite).
XGC_M <- c(aging = 1, departure = 2, arrival = 4, hivtest = 8, hivtx = 16, hivprogress = 32, hivvl = 64, resim_nets = 128,
           acts = 256, condoms = 512, position = 1024, prep = 2048, hivtrans = 4096, stirecov = 8192, stitx = 16384,
           stitrans = 32768, prev = 65536)
.xgc_module <- function(bit) function(dat, at) { .Call("xgcR_step", dat$xgc, as.integer(at), as.numeric(bit), NULL); dat }
aging_msm_gpu        <- .xgc_module(XGC_M[["aging"]]);       departure_msm_gpu   <- .xgc_module(XGC_M[["departure"]])
arrival_msm_gpu      <- .xgc_module(XGC_M[["arrival"]]);     hivtest_msm_gpu     <- .xgc_module(XGC_M[["hivtest"]])
hivtx_msm_gpu        <- .xgc_module(XGC_M[["hivtx"]]);       hivprogress_msm_gpu <- .xgc_module(XGC_M[["hivprogress"]])
hivvl_msm_gpu        <- .xgc_module(XGC_M[["hivvl"]]);       acts_msm_gpu        <- .xgc_module(XGC_M[["acts"]])
condoms_msm_gpu      <- .xgc_module(XGC_M[["condoms"]]);     position_msm_gpu    <- .xgc_module(XGC_M[["position"]])
prep_msm_gpu         <- .xgc_module(XGC_M[["prep"]]);        hivtrans_msm_gpu    <- .xgc_module(XGC_M[["hivtrans"]])
stirecov_msm_gpu     <- .xgc_module(XGC_M[["stirecov"]]);    stitx_msm_gpu       <- .xgc_module(XGC_M[["stitx"]])
stitrans_msm_rand_gpu <- .xgc_module(XGC_M[["stitrans"]]);   prevalence_msm_gpu  <- .xgc_module(XGC_M[["prev"]])

# Fast mode (.Call("xgcR_set_mode", dat$xgc, 1L)): the replicate-resident kernel runs module GROUPS, not single modules.  The wrappers
# below keep the FUN(dat, at) contract: the last module of the aging..hivvl group issues the group, the last module of the week
# (prev.FUN) issues the acts..prevalence group TOGETHER with the aging..hivvl group of the next week (xgc_step_span: netsim has
# nothing on the host between prevalence_msm(at) and aging_msm(at + 1)); every other wrapper returns dat unchanged.
#   control <- do.call(control_msm, c(list(..., initialize.FUN = initialize_msm_gpu, resim_nets.FUN = simnet_msm_gpu), xgc_fast_modules()))
xgc_fast_modules <- function() {
  PRE  <- sum(XGC_M[c("aging", "departure", "arrival", "hivtest", "hivtx", "hivprogress", "hivvl")])
  POST <- sum(XGC_M[c("acts", "condoms", "position", "prep", "hivtrans", "stirecov", "stitx", "stitrans", "prev")])
  noop <- function(dat, at) dat
  pre_last <- function(dat, at) {                                  # (already run by last week's prev.FUN unless this is the first week)
    if (!identical(dat$xgc_pre_done, as.integer(at))) .Call("xgcR_step", dat$xgc, as.integer(at), as.numeric(PRE), NULL)
    dat
  }
  post_last <- function(dat, at) {
    if (at >= dat$control$nsteps) .Call("xgcR_step", dat$xgc, as.integer(at), as.numeric(POST), NULL)
    else { .Call("xgcR_step_span", dat$xgc, as.integer(at), as.numeric(POST), as.numeric(PRE)); dat$xgc_pre_done <- as.integer(at) + 1L }
    dat
  }
  list(aging.FUN = noop, departure.FUN = noop, arrival.FUN = noop, hivtest.FUN = noop, hivtx.FUN = noop, hivprogress.FUN = noop,
       hivvl.FUN = pre_last, acts.FUN = noop, condoms.FUN = noop, position.FUN = noop, prep.FUN = noop, hivtrans.FUN = noop,
       stirecov.FUN = noop, stitx.FUN = noop, stitrans.FUN = noop, prev.FUN = post_last)
}

# resim_nets.FUN: pull what tergmLite needs, run the reference's own simnet_msm on the host, push the edge lists back.
simnet_msm_gpu <- function(dat, at) {
  # deg.main / deg.casl are derived on the device from the current edge lists (what tergmLite's formation terms condition on)
  for (a in c("age", "age.grp", "race", "role.class", "risk.grp", "deg.main", "deg.casl"))
    dat$attr[[a]] <- .Call("xgcR_download_attr", dat$xgc, 0L, a)
  dat$attr$active <- rep(1, length(dat$attr$race))
  for (l in 1:3) { el <- .Call("xgcR_get_edgelist", dat$xgc, 0L, l, dat$xgc_ecap[l]); attr(el, "n") <- length(dat$attr$race); dat$el[[l]] <- el }
  dat <- EpiModelHIV::simnet_msm(dat, at)                      # unchanged host STERGM step
  for (l in 1:3) .Call("xgcR_set_edgelist", dat$xgc, 0L, l, dat$el[[l]],
                       if (l < 3) as.integer(dat$temp$plist[dat$temp$plist[, "ptype"] == l & is.na(dat$temp$plist[, "stop"]), "start"]) else NULL)
  dat
}
# injected-draw mode (keeps R's rlecuyer stream): fill the table with runif() and pass it to every step of the week
xgc_inject_table <- function(uid_cap, e_cap, form_cap, size) list(runif(size), as.integer(c(uid_cap, e_cap, form_cap)))

# ---------------------------------------------------------------------------------------------------------------------
# initialize.FUN / restart (burnin/cal/sim1_02_lhs.r:201; inst/analysis02_screening/02.00_epi.r:292-296:
#   initialize.FUN = ifelse(burnin, initialize_msm, reinit_msm))
# ---------------------------------------------------------------------------------------------------------------------
# param_msm() list -> the library's parameter vector (include/xgc/params.def: names are the param_msm argument names with
# '.' -> '_'; vectors by race B,H,O,W / by site rect,ureth,phar keep their order).  Entries param_msm does not carry keep
# the defaults of params.def.
xgc_param_vector <- function(param, defaults) {
  v <- defaults
  for (nm in names(param)) {
    ix <- .Call("xgcR_param_index", gsub(".", "_", nm, fixed = TRUE))
    if (!is.null(ix) && is.numeric(param[[nm]]) && length(param[[nm]]) == ix[2]) v[ix[1] + seq_len(ix[2])] <- as.numeric(param[[nm]])
  }
  v
}
xgc_control_vector <- function(control) {
  k <- function(nm) .Call("xgcR_constant", nm)
  v <- integer(k("XGC_NCONTROL"))
  v[k("XGC_C_transRoute_Kissing") + 1L]      <- as.integer(isTRUE(control$transRoute_Kissing))
  v[k("XGC_C_transRoute_Rimming") + 1L]      <- as.integer(isTRUE(control$transRoute_Rimming))
  v[k("XGC_C_cdcExposureSite_Kissing") + 1L] <- as.integer(isTRUE(control$cdcExposureSite_Kissing))
  v[k("XGC_C_gcUntreatedRecovDist") + 1L]    <- k(if (identical(control$gcUntreatedRecovDist, "pois")) "XGC_RECOV_POIS" else "XGC_RECOV_GEOM")
  v[k("XGC_C_stiScreeningProtocol") + 1L]    <- k(paste0("XGC_SCREEN_", toupper(control$stiScreeningProtocol)))
  v[k("XGC_C_network_mode") + 1L]            <- 0L        # the STERGM stays in R: simnet_msm_gpu
  v
}
# everything the constructor needs besides dat: capacities (25 % head-room like the reference's +-5 % population window
# times the burn-in length), the netstats / epistats tables flattened in the layout of xgc_set_tables (abi.h XGC_T_*)
.xgc_upload_dat <- function(dat, xgc) {
  known <- .Call("xgcR_attr_names")
  n <- length(dat$attr$race)
  .Call("xgcR_set_population", xgc, 0L, as.integer(n), 1L)
  for (a in intersect(names(dat$attr), known))
    if (!a %in% c("deg.main", "deg.casl", "deg.inst", "active")) .Call("xgcR_upload_attr", xgc, 0L, a, dat$attr[[a]])
  for (l in 1:3) {
    st <- if (l < 3 && !is.null(dat$temp$plist)) {
      pl <- dat$temp$plist; as.integer(pl[pl[, "ptype"] == l & is.na(pl[, "stop"]), "start"]) } else NULL
    .Call("xgcR_set_edgelist", xgc, 0L, l, dat$el[[l]], st)
  }
  invisible(xgc)
}
initialize_msm_gpu <- function(x, param, init, control, s) {
  dat <- EpiModelHIV::initialize_msm(x, param, init, control, s)      # the reference's own constructor: attrs, els, plist
  n <- length(dat$attr$race)
  n_cap <- as.integer(ceiling(n * 1.3 / 256) * 256)
  e_cap <- as.integer(sapply(1:3, function(l) ceiling(max(nrow(dat$el[[l]]), 64) * 1.5 / 128) * 128))
  t_cap <- as.integer(control$nsteps + 2)
  dat$xgc <- .Call("xgcR_create", 1L, n_cap, e_cap, t_cap, as.numeric(if (is.null(control$xgc.seed)) s else control$xgc.seed),
                   as.integer(if (is.null(control$xgc.device)) 0L else control$xgc.device), as.integer(s - 1L))
  dat$xgc_ecap <- e_cap; dat$xgc_ncap <- n_cap; dat$xgc_seed <- as.numeric(if (is.null(control$xgc.seed)) s else control$xgc.seed)
  if (!is.null(control$xgc.tables)) .Call("xgcR_set_tables", dat$xgc, as.numeric(control$xgc.tables))
  defaults <- if (is.null(control$xgc.param.defaults)) stop("control$xgc.param.defaults: the params.def default vector") else control$xgc.param.defaults
  pv <- xgc_param_vector(param, defaults)
  pv[.Call("xgcR_param_index", "n_init")[1] + 1L] <- n
  .Call("xgcR_set_params", dat$xgc, 0L, pv)
  .Call("xgcR_set_control", dat$xgc, 0L, xgc_control_vector(control))
  .xgc_upload_dat(dat, dat$xgc)
  prevalence_msm_gpu(dat, 1L)                                          # prevalence_msm(dat, at = 1), as initialize_msm ends
}
# restart from a finished burn-in (02.00_epi.r:40-76): x is the saved netsim object whose `xgc_state` element holds the
# raw device snapshot written by xgc_save_state(); parameters / screening arm may differ from the burn-in's
xgc_save_state <- function(dat) list(raw = .Call("xgcR_snapshot", dat$xgc), n_cap = dat$xgc_ncap, e_cap = dat$xgc_ecap, seed = dat$xgc_seed,
                                     params = .Call("xgcR_get_params", dat$xgc, 0L))
reinit_msm_gpu <- function(x, param, init, control, s) {
  dat <- EpiModelHIV::reinit_msm(x, param, init, control, s)          # R-side bookkeeping (dat$param, dat$control, dat$epi, nwparam)
  st <- x$xgc_state[[s]]
  dat$xgc <- .Call("xgcR_create", 1L, as.integer(st$n_cap), as.integer(st$e_cap), as.integer(control$nsteps + 2),
                   as.numeric(st$seed), as.integer(if (is.null(control$xgc.device)) 0L else control$xgc.device), as.integer(s - 1L))
  dat$xgc_ecap <- as.integer(st$e_cap); dat$xgc_ncap <- as.integer(st$n_cap); dat$xgc_seed <- st$seed
  .Call("xgcR_restore", dat$xgc, st$raw)
  .Call("xgcR_set_params", dat$xgc, 0L, xgc_param_vector(param, st$params))
  .Call("xgcR_set_control", dat$xgc, 0L, xgc_control_vector(control))
  dat
}
