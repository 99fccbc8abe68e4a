// Instantiations of the filtered streaming fusion kernel for C = 4.
#include "fuse_filter.cuh"

int pisto_launch_filter_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  const int G = pisto_filter_groups(p);
  if (p.V == 1 && G == 1) return np == 2 ? pisto_launch_filter_cvg<4, 1, 1, 2>(h, p, st, launched) : pisto_launch_filter_cvg<4, 1, 1, 1>(h, p, st, launched);
  if (p.V == 6 && G == 3) return np == 2 ? pisto_launch_filter_cvg<4, 6, 3, 2>(h, p, st, launched) : pisto_launch_filter_cvg<4, 6, 3, 1>(h, p, st, launched);
  return PISTO_OK;
}
