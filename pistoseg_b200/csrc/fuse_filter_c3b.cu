// Instantiations of the filtered streaming fusion kernel for C = 3, V = 6 (3 scales x flip: BASELINE config 2).
#include "fuse_filter.cuh"

int pisto_launch_filter_c3_v6(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  return pisto_launch_filter_cvg<3, 6, 3>(h, p, st, np, launched);
}
