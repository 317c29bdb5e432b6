// Interface between libxgc's host side (xgc_lib.cu, compiled --fmad=false for the strict arithmetic contract) and the
// replicate-resident fast path (xgc_fast.cu, compiled with FMA contraction and fp32 probability arithmetic).
#pragma once
#include "xgc_device.cuh"

// Phase groups of one week (what a launch of k_week_fast runs); xgc_step maps module masks onto them.
enum { XF_PRE = 1,        // week bookkeeping, arrival, aging, departure, hivtest, hivtx, hivprogress, hivvl
       XF_PRUNE = 2,      // edges of departed agents removed (tergmLite::delete_vertices)
       XF_FORM = 4,       // surrogate simnet_msm: dissolution + formation (only replicates with network_mode = 1)
       XF_POST = 8,       // acts, condoms, position, prep, hivtrans, stirecov, stitx, stitrans, prev
       XF_ALL = 15 };

struct FastArgs {
  int at0, at1;           // weeks at0..at1 inclusive are run back to back inside ONE launch (at1 > at0 needs XF_ALL)
  unsigned phases;        // XF_* bits of week at0 (of the only week; XF_ALL for a multi-week run)
  unsigned ph_last;       // XF_* bits of week at1 (= phases unless the call is a span: rest of week at0 + start of week at1)
  unsigned mask;          // module mask the phases were derived from (modules inside a group can still be switched off)
  int write_acts;         // 1: the last week's anal / oral counts are written to the edge records (xgc_get_actcounts)
  // state tallies carried from one span call to the next (week at0's running prevalence tallies as the previous span left them after
  // its aging..hivvl group): carry_in = 1 -> loaded instead of recounted at staging; a span always leaves them in `carry` [R][NCOUNTER]
  unsigned* carry;
  int carry_in;
  // Host-network edge-list operations the host queued since the last call (xgc_update_edgelists_all / xgc_set_edgelists_all in
  // fast mode), applied by the kernel itself at the start of week at0 -- before the network / acts phases -- instead of by the
  // standalone k_rank / k_edges_update / k_edges_upload launches.  Per layer: 0 = nothing, 1 = weekly delta, 2 = whole list.
  int hn_kind[3];
  const int* hn_ptr[3];   // 1: delta rows [R][stride]; 2: positions [R][2][stride]
  const int* hn_ne[3];    // 2: n_edges [R]
  const int* hn_start[3]; // 2: start weeks [R][stride] or null
  int hn_stride[3];
  int hn_has_start[3];    // 1: rows carry start weeks
};

// bytes of dynamic shared memory one CTA needs for (n_cap, e_stride); 0 = does not fit one SM (fast path unavailable)
size_t xgc_fast_smem_bytes(int n_cap, size_t e_stride);
// quantised birth week / insertivity quotient in the spare bits of the core words (what the fast path reads instead of
// the 16-byte aux record); call after any host-side change of ages (attribute upload, restore, set_population)
cudaError_t xgc_fast_prepare(const Dev& g, cudaStream_t s);
cudaError_t xgc_fast_launch(const Dev& g, const FastArgs& a, size_t smem, cudaStream_t s);
const void* xgc_fast_kernel_ptr();
// does the rank structure of the in-kernel edge-list operations fit the kernel's scratch area (the mini-queues) for this shape?
bool xgc_fast_hostnet_fits(int n_cap, const int* e_cap);

// ---- spare bits of the core words used by the fast path (ignored by the strict kernels, preserved by all of them)
// demo bits 16..31: ins.quot as unorm16
// hiv  bits 12..30: birth code = round((t_ref - age_ref * 365/7 + XF_BIRTH_OFF) * 8), i.e. the (fractional) week of birth in 1/8 weeks
#define XF_BIRTH_OFF 8192
#define XF_BIRTH_SHIFT 12
#define XF_BIRTH_MASK 0x7FFFFu
#define XF_W8 (7.0 / 365.0 / 8.0)                 /* years per 1/8 week */
static inline __host__ __device__ unsigned xf_birth_code(double age_ref, int t_ref) {
  double b = ((double)t_ref - age_ref * (365.0 / 7.0) + (double)XF_BIRTH_OFF) * 8.0;
  long long c = (long long)(b + 0.5);
  if (c < 0) c = 0;
  if (c > (long long)XF_BIRTH_MASK) c = XF_BIRTH_MASK;
  return (unsigned)c;
}
static inline __host__ __device__ unsigned xf_insq16(float q) {
  float v = q * 65535.0f + 0.5f;
  return v <= 0.f ? 0u : v >= 65535.f ? 65535u : (unsigned)v;
}
