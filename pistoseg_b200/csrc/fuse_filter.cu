// Dispatch of the filtered streaming fusion kernel (fuse_filter.cuh); the (C, V, G) families are instantiated in
// fuse_filter_c3.cu / fuse_filter_c3b.cu / fuse_filter_c4.cu so that they compile in parallel.
#include <stdlib.h>

#include "fuse_common.cuh"

int pisto_launch_filter_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched);
int pisto_launch_filter_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched);
int pisto_launch_static_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);  // fuse_static.cuh
int pisto_launch_static_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_narrow_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);  // fuse_static.cuh, 2 columns per thread
int pisto_launch_narrow_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_duo_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);     // fuse_static.cuh, two CTAs per SM

// np: column pairs per thread to use (1 or 2), 0 = pick (automatic dispatch: the shape-specialised kernel of fuse_static.cuh
// first), 3 = the shape-specialised kernel only, 4 = its two-CTAs-per-SM variant only, 5 = its 2-columns-per-thread variant only
int pisto_launch_fuse_filter(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  *launched = false;
  if (p.fuse_mode != PISTO_FUSE_LOGIT_MEAN) return PISTO_OK;       // softmax per view is not linear
  if (p.fused_out || p.entropy_out) return PISTO_OK;               // full-resolution scores wanted: exact kernels
  if (p.dec.mask_mode == PISTO_MASK_MULTIPLY) return PISTO_OK;     // masked classes take part in the argmax
  if (!p.label_out && !(p.conf && p.gt)) return PISTO_OK;
  if (p.T_h > 65535 || p.T_w > 65535) return PISTO_OK;
  const uintptr_t bytes = (uintptr_t)p.label_out | (uintptr_t)p.bg | (uintptr_t)p.gt;
  if (bytes & 1) return PISTO_OK;
  bool views_aligned = true;
  for (int v = 0; v < p.V; v++) views_aligned &= ((uintptr_t)p.view[v].logits & 15) == 0;  // TMA source spans need a 16-byte aligned allocation start
  if (np == 4) return (views_aligned && p.C == 3) ? pisto_launch_duo_c3(h, p, st, launched) : PISTO_OK;  // explicit only: the one-CTA kernel measures faster (profiles/r02)
  if (np == 5) {
    if (!views_aligned) return PISTO_OK;
    return p.C == 3 ? pisto_launch_narrow_c3(h, p, st, launched) : (p.C == 4 ? pisto_launch_narrow_c4(h, p, st, launched) : PISTO_OK);
  }
  if ((np == 0 || np == 3) && views_aligned) {
    static const bool no_static = getenv("PISTO_NO_STATIC") != nullptr;  // A/B knob
    int rc = PISTO_OK;
    // automatic dispatch: where it is measured faster than the generic kernel (C = 3, three scales x flip); always when asked for
    if (np == 3 || (!no_static && (p.V == 6 || p.V == 1))) {
      if (p.C == 3) rc = pisto_launch_static_c3(h, p, st, launched);
      else if (p.C == 4) rc = pisto_launch_static_c4(h, p, st, launched);
    }
    if (rc != PISTO_OK || *launched || np == 3) return rc;
  }
  if (np == 3) return PISTO_OK;
  if (np == 0) np = ((bytes & 3) == 0 && p.T_w % 4 == 0) ? 2 : 1;
  if (np == 2 && ((bytes & 3) || p.T_w % 4)) return PISTO_OK;
  for (int v = 0; v < p.V; v++)
    if ((uintptr_t)p.view[v].logits & 15) return PISTO_OK;         // TMA source spans need a 16-byte aligned allocation start
  if (p.C == 3) return pisto_launch_filter_c3(h, p, st, np, launched);
  if (p.C == 4) return pisto_launch_filter_c4(h, p, st, np, launched);
  return PISTO_OK;
}
