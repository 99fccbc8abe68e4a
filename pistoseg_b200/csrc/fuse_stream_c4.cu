// Instantiations of the streaming fusion kernel for C = 4.
#include "fuse_stream.cuh"

int pisto_launch_stream_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  switch (p.V) {
    case 1: return pisto_launch_stream_cv<4, 1>(h, p, st, launched);
    case 2: return pisto_launch_stream_cv<4, 2>(h, p, st, launched);
    case 6: return pisto_launch_stream_cv<4, 6>(h, p, st, launched);
  }
  return PISTO_OK;
}
