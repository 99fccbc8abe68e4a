// pisto_confusion_accumulate: conf[gt*C + pred] += 1 over pixels with gt < C.
// Replaces mIoUMask._generate_matrix / add_batch (reference loss.py:17-31): numpy mask + bincount on the host.
//
// HBM-bound: 2 bytes read per pixel, nothing written.  Each thread streams 16-byte vectors of pred and gt
// (ld.global.nc, L1 no-allocate), counts 32 pixels at a time on bit planes (below), and the CTA issues one 64-bit
// atomicAdd per bin at the end.  Grid = 8 CTAs per SM, grid-stride over 32-pixel groups.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Bit-sliced counting (C <= 4, the WSSS4LUAD / BCSS cases).  A thread takes 32 pixels (two 16-byte vectors of each
// array = 8 words) and transposes them into bit planes without moving single bytes: word j contributes its bytes' bit k
// at bit position j of the corresponding plane byte,   plane_k |= ((w_j >> k) << j) & (0x01010101 << j),
// which is one shift and one LOP3 per (word, plane) and leaves the 32 pixels in a fixed permutation that is the same for
// every plane.  Planes: the two low bits of gt and pred plus an "any higher bit" plane each (value >= 4).  The confusion
// counts of the 32 pixels are then popc(G_a & P_b) for a, b < C: 3 instructions per BIN instead of ~12 per PIXEL, so the
// kernel is bound by the loads again.  Counters are plain 32-bit registers (C*C <= 16 of them).
template <int C>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                             long long n_px, unsigned long long* __restrict__ conf,
                                                             unsigned long long* __restrict__ bad_pred) {
  constexpr int BINS = C * C;
  __shared__ unsigned int hist[BINS + 1];
  for (int i = threadIdx.x; i <= BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();

  unsigned int cnt[BINS], bad = 0;
#pragma unroll
  for (int i = 0; i < BINS; i++) cnt[i] = 0;

  // planes of 8 words: bit0, bit1 and "value >= 4" of every byte
  auto planes = [](const unsigned int (&w)[8], unsigned int& b0, unsigned int& b1, unsigned int& bx) {
    b0 = b1 = bx = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const unsigned int m = 0x01010101u << j;
      b0 |= (w[j] << j) & m;
      b1 |= (j == 0 ? (w[j] >> 1) : (w[j] << (j - 1))) & m;
      const unsigned int hi = ((w[j] >> 2) & 0x3f3f3f3fu) + 0x3f3f3f3fu;  // bit 6 of a byte set iff the byte is >= 4
      bx |= (j <= 6 ? (hi >> (6 - j)) : (hi << (j - 6))) & m;
    }
  };

  const long long n_grp = n_px / 32;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_grp; v += stride) {
    const uint4 p0 = ld_stream16(pred + v * 32), p1 = ld_stream16(pred + v * 32 + 16);
    const uint4 g0 = ld_stream16(gt + v * 32), g1 = ld_stream16(gt + v * 32 + 16);
    const unsigned int pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    const unsigned int gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    unsigned int ga0, ga1, gax, pa0, pa1, pax;
    planes(gw, ga0, ga1, gax);
    planes(pw, pa0, pa1, pax);
    unsigned int G[4], P[4];
    G[0] = ~ga1 & ~ga0 & ~gax; G[1] = ~ga1 & ga0 & ~gax; G[2] = ga1 & ~ga0 & ~gax; G[3] = ga1 & ga0 & ~gax;
    P[0] = ~pa1 & ~pa0 & ~pax; P[1] = ~pa1 & pa0 & ~pax; P[2] = pa1 & ~pa0 & ~pax; P[3] = pa1 & pa0 & ~pax;
    unsigned int gvalid = 0, pvalid = 0;
#pragma unroll
    for (int a = 0; a < C; a++) { gvalid |= G[a]; pvalid |= P[a]; }
#pragma unroll
    for (int a = 0; a < C; a++)
#pragma unroll
      for (int c = 0; c < C; c++) cnt[a * C + c] += __popc(G[a] & P[c]);
    bad += __popc(gvalid & ~pvalid);   // counted pixel whose prediction is >= C: an error in the reference (bincount overflow)
  }
  // tail (n_px % 32), one pixel per lane of the first warp of block 0
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const long long i = n_grp * 32 + threadIdx.x;
    if (i < n_px) {
      const unsigned int g = gt[i], q = pred[i];
      if (g < (unsigned)C) {
        if (q < (unsigned)C) {
#pragma unroll
          for (int bn = 0; bn < BINS; bn++) cnt[bn] += (g * C + q == (unsigned)bn) ? 1u : 0u;
        } else {
          bad++;
        }
      }
    }
  }
#pragma unroll
  for (int bn = 0; bn < BINS; bn++) {
    const unsigned int v = __reduce_add_sync(0xffffffffu, cnt[bn]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&hist[bn], v);
  }
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
  long long n_vec = n_px / 32;
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
