// Dispatch of the streaming fusion kernel (fuse_stream.cuh); the (C, V) families are instantiated in
// fuse_stream_c3.cu / fuse_stream_c4.cu so that they compile in parallel.
#include "fuse_common.cuh"

int pisto_launch_stream_c3(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_stream_c4(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);

int pisto_launch_fuse_stream(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  *launched = false;
  if (p.entropy_out) return PISTO_OK;              // dead output in the reference: generic kernel only
  if (p.conf && p.gt && p.C > 4) return PISTO_OK;  // packed counters hold C*C <= 16 bins
  // 2-byte / 8-byte vector accesses need even addresses; TMA source spans need a 16-byte aligned allocation start
  if (((uintptr_t)p.label_out | (uintptr_t)p.bg | (uintptr_t)p.gt) & 1) return PISTO_OK;
  if ((uintptr_t)p.fused_out & 7) return PISTO_OK;
  for (int v = 0; v < p.V; v++)
    if ((uintptr_t)p.view[v].logits & 15) return PISTO_OK;
  if (p.C == 3) return pisto_launch_stream_c3(h, p, st, launched);
  if (p.C == 4) return pisto_launch_stream_c4(h, p, st, launched);
  return PISTO_OK;
}
