// Instantiations of the filtered streaming fusion kernel for C = 3.
#include "fuse_filter.cuh"

int pisto_launch_filter_c3_v6(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched);

int pisto_launch_filter_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  const int G = pisto_filter_groups(p);
  if (p.V == 1 && G == 1) return pisto_launch_filter_cvg<3, 1, 1>(h, p, st, np, launched);
  if (p.V == 2 && G == 1) return pisto_launch_filter_cvg<3, 2, 1>(h, p, st, np, launched);
  if (p.V == 6 && G == 3) return pisto_launch_filter_c3_v6(h, p, st, np, launched);
  return PISTO_OK;
}
