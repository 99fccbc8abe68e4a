// pisto_confusion_accumulate: conf[gt*C + pred] += 1 over pixels with gt < C.
// Replaces mIoUMask._generate_matrix / add_batch (reference loss.py:17-31): numpy mask + bincount on the host.
//
// HBM-bound: 2 bytes read per pixel, nothing written.  Each thread streams 16-byte vectors of pred and gt
// (ld.global.nc, L1 no-allocate), keeps its C*C counters PACKED in registers (8 bits per bin, flushed to the
// per-warp shared-memory histogram before they can overflow), and the CTA issues one 64-bit atomicAdd per bin at
// the end.  Grid = 8 CTAs per SM, grid-stride over vectors.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// counters: C*C <= 16 bins of 8 bits in two u64 (C <= 4) -- the WSSS4LUAD / BCSS cases; larger C uses shared atomics.
template <int C>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                             long long n_px, unsigned long long* __restrict__ conf,
                                                             unsigned long long* __restrict__ bad_pred) {
  constexpr int BINS = C * C;
  __shared__ unsigned int hist[BINS + 1];
  for (int i = threadIdx.x; i <= BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();

  unsigned long long lo = 0, hi = 0;  // packed 8-bit counters, bins 0..7 / 8..15
  unsigned int bad = 0;
  int pending = 0;
  auto flush = [&]() {
#pragma unroll
    for (int b = 0; b < BINS; b++) {
      unsigned int v = (unsigned int)(((b < 8 ? lo : hi) >> (8 * (b & 7))) & 0xffull);
      // warp-aggregate before touching shared memory
      v = __reduce_add_sync(0xffffffffu, v);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&hist[b], v);
    }
    lo = hi = 0;
    pending = 0;
  };
  auto count = [&](unsigned int g, unsigned int p) {
    if (g < (unsigned)C) {
      if (p < (unsigned)C) {
        unsigned int b = g * C + p;
        unsigned long long inc = 1ull << (8 * (b & 7));
        if (b < 8) lo += inc; else hi += inc;
      } else {
        bad++;
      }
    }
  };

  const long long n_vec = n_px / 16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations (flush uses full-mask warp reductions)
  const long long warp_base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long vbase = warp_base; vbase < n_vec; vbase += stride) {
    long long v = vbase + (threadIdx.x & 31);
    if (v < n_vec) {
      uint4 pv = ld_stream16(pred + v * 16);
      uint4 gv = ld_stream16(gt + v * 16);
      unsigned int pw[4] = {pv.x, pv.y, pv.z, pv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int j = 0; j < 4; j++) count((gw[k] >> (8 * j)) & 0xffu, (pw[k] >> (8 * j)) & 0xffu);
      }
    }
    pending += 16;
    if (pending > 255 - 16) flush();
  }
  // scalar tail (n_px % 16), handled by the first warp of block 0
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    long long i = n_vec * 16 + threadIdx.x;
    if (i < n_px) count(gt[i], pred[i]);
  }
  flush();
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&hist[BINS], bad);
  __syncthreads();
  for (int i = threadIdx.x; i < BINS; i += blockDim.x)
    if (hist[i]) atomicAdd(&conf[i], (unsigned long long)hist[i]);
  if (threadIdx.x == 0 && bad_pred && hist[BINS]) atomicAdd(bad_pred, (unsigned long long)hist[BINS]);
}

// generic C (5..16): per-warp shared histograms with match_any aggregation
__global__ void __launch_bounds__(kThreads) confusion_kernel_generic(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                                     long long n_px, int C, unsigned long long* __restrict__ conf,
                                                                     unsigned long long* __restrict__ bad_pred) {
  extern __shared__ unsigned int ghist[];  // C*C + 1
  const int BINS = C * C;
  for (int i = threadIdx.x; i <= BINS; i += blockDim.x) ghist[i] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    unsigned int g = gt[i], p = pred[i];
    if (g < (unsigned)C) atomicAdd(&ghist[p < (unsigned)C ? g * C + p : BINS], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BINS; i += blockDim.x)
    if (ghist[i]) atomicAdd(&conf[i], (unsigned long long)ghist[i]);
  if (threadIdx.x == 0 && bad_pred && ghist[BINS]) atomicAdd(bad_pred, (unsigned long long)ghist[BINS]);
}

}  // namespace

extern "C" int pisto_confusion_accumulate(pisto_handle_t h, const uint8_t* pred, const uint8_t* gt, int64_t n_px, int C,
                                          unsigned long long* conf, unsigned long long* bad_pred, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_confusion_accumulate: NULL handle");
  PISTO_REQUIRE(C >= 1 && C <= 16, "pisto_confusion_accumulate: C=%d outside [1,16]", C);
  PISTO_REQUIRE(n_px >= 0, "pisto_confusion_accumulate: n_px < 0");
  PISTO_REQUIRE(conf, "pisto_confusion_accumulate: conf is NULL");
  if (n_px == 0) return PISTO_OK;
  PISTO_REQUIRE(pred && gt, "pisto_confusion_accumulate: pred/gt NULL");
  cudaStream_t st = (cudaStream_t)stream;
  PISTO_CUDA(cudaSetDevice(h->device));
  long long n_vec = n_px / 16;
  int grid = (int)((n_vec + kThreads - 1) / kThreads);
  int max_grid = h->sm_count * 8;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  bool aligned = (((uintptr_t)pred | (uintptr_t)gt) & 15) == 0;
  if (C <= 4 && aligned) {
    switch (C) {
      case 1: confusion_kernel<1><<<grid, kThreads, 0, st>>>(pred, gt, n_px, conf, bad_pred); break;
      case 2: confusion_kernel<2><<<grid, kThreads, 0, st>>>(pred, gt, n_px, conf, bad_pred); break;
      case 3: confusion_kernel<3><<<grid, kThreads, 0, st>>>(pred, gt, n_px, conf, bad_pred); break;
      default: confusion_kernel<4><<<grid, kThreads, 0, st>>>(pred, gt, n_px, conf, bad_pred); break;
    }
  } else {
    int g2 = (int)((n_px + kThreads - 1) / kThreads);
    if (g2 > max_grid) g2 = max_grid;
    confusion_kernel_generic<<<g2, kThreads, (C * C + 1) * sizeof(unsigned int), st>>>(pred, gt, n_px, C, conf, bad_pred);
  }
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
