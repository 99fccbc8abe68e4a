// Instantiations of the shape-specialised filtered fusion kernel (fuse_static.cuh) for C = 4: BASELINE config 3 (BCSS).
#include "fuse_static.cuh"

int pisto_launch_static_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  const int f = pisto_filter_flags(p);
  if (p.V == 6) {
    if (f == 18) return launch_static<4, 3, 2, 18, 1>(h, p, st, launched);  // gt/conf + labels (config 3)
    return launch_static<4, 3, 2, -1, 1>(h, p, st, launched);
  }
  return PISTO_OK;
}

// 2 columns per thread, 26 warps per SM (fuse_narrow_kernel): three difference fields fit the registers without spilling
int pisto_launch_narrow_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  const int f = pisto_filter_flags(p);
  if (p.V == 6 && f == 18) return launch_narrow<4, 3, 2, 18, 1>(h, p, st, launched);
  return PISTO_OK;
}
