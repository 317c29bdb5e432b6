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

# resim_nets.FUN: pull what tergmLite needs, run the reference's own simnet_msm on the host, push the edge lists back.
simnet_msm_gpu <- function(dat, at) {
  for (a in c("active", "age", "age.grp", "race", "role.class", "risk.grp", "deg.main", "deg.casl"))
    if (a %in% c("active")) dat$attr[[a]] <- rep(1, length(.Call("xgcR_download_attr", dat$xgc, 0L, "race")))
    else if (!a %in% c("deg.main", "deg.casl")) dat$attr[[a]] <- .Call("xgcR_download_attr", dat$xgc, 0L, a)
  for (l in 1:3) { el <- .Call("xgcR_get_edgelist", dat$xgc, 0L, l, dat$xgc_ecap[l]); attr(el, "n") <- length(dat$attr$race); dat$el[[l]] <- el }
  dat <- EpiModelHIV::simnet_msm(dat, at)                      # unchanged host STERGM step
  for (l in 1:3) .Call("xgcR_set_edgelist", dat$xgc, 0L, l, dat$el[[l]],
                       if (l < 3) as.integer(dat$temp$plist[dat$temp$plist[, "ptype"] == l & is.na(dat$temp$plist[, "stop"]), "start"]) else NULL)
  dat
}
# injected-draw mode (keeps R's rlecuyer stream): fill the table with runif() and pass it to every step of the week
xgc_inject_table <- function(uid_cap, e_cap, form_cap, size) list(runif(size), as.integer(c(uid_cap, e_cap, form_cap)))
