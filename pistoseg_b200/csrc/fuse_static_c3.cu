// Instantiations of the shape-specialised filtered fusion kernel (fuse_static.cuh) for C = 3: BASELINE configs 1 and 2.
#include "fuse_static.cuh"

// two 256-thread CTAs per SM (fuse_duo_kernel)
int pisto_launch_duo_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  const int f = pisto_filter_flags(p);
  if (p.V == 6) {
    if (f == 25) return launch_duo<3, 3, 2, 25>(h, p, st, launched);
    return launch_duo<3, 3, 2, -1>(h, p, st, launched);
  }
  if (p.V == 1) {
    if (f == 19) return launch_duo<3, 1, 1, 19>(h, p, st, launched);
    return launch_duo<3, 1, 1, -1>(h, p, st, launched);
  }
  return PISTO_OK;
}

// 2 columns per thread, 26 warps per SM (fuse_narrow_kernel)
int pisto_launch_narrow_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  const int f = pisto_filter_flags(p);
  if (p.V == 6) {
    if (f == 25) return launch_narrow<3, 3, 2, 25, 2>(h, p, st, launched);
    if (f == 17) return launch_narrow<3, 3, 2, 17, 2>(h, p, st, launched);
  }
  return PISTO_OK;
}

int pisto_launch_static_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  const int f = pisto_filter_flags(p);
  if (p.V == 6) {
    if (f == 25) return launch_static<3, 3, 2, 25, 2>(h, p, st, launched);  // bg + labels + 32x32 (config 2)
    if (f == 17) return launch_static<3, 3, 2, 17, 2>(h, p, st, launched);  // bg + labels
    return launch_static<3, 3, 2, -1, 2>(h, p, st, launched);
  }
  if (p.V == 1) {
    if (f == 19) return launch_static<3, 1, 1, 19, 2>(h, p, st, launched);  // bg + gt/conf + labels (config 1)
    return launch_static<3, 1, 1, -1, 2>(h, p, st, launched);
  }
  return PISTO_OK;
}
