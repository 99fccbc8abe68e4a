// Big-mask stitching in float64: pisto_stitch_accumulate / pisto_canvas_normalize / pisto_canvas_axpy / pisto_argmax_f64.
// Replaces the host-side numpy loops of segmentation_test.py:141-215 and
// OEEM/classification/prepare_seg_inputs.py:120-136 (per-tile D2H copy + slice-add on float64 canvases).
//
// Determinism: overlapping tiles are NOT merged with floating-point atomics.  Every canvas pixel is owned by one
// thread, which visits the covering tiles in index order -- the same order as the reference's sequential
// `canvas[y:y+h, x:x+w] += probs` loop -- so the float64 sums are bit-identical to the reference's.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPosChunk = 512;

template <int C>
__global__ void __launch_bounds__(kThreads) stitch_kernel(const float* __restrict__ tiles, const pisto_tile_pos_t* __restrict__ pos, int n,
                                                          int th, int tw, int softmax, double* __restrict__ canvas,
                                                          double* __restrict__ count, int H, int W) {
  __shared__ pisto_tile_pos_t spos[kPosChunk];
  const long long HW = (long long)H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < HW;
  const int Y = live ? (int)(pix / W) : 0, X = live ? (int)(pix - (long long)Y * W) : 0;
  double acc[C];
  double cnt = 0.0;
  if (live) {
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = canvas[c * HW + pix];
    cnt = count[pix];
  }
  const long long tile_elems = (long long)C * th * tw;
  // Bounding box of the block's pixels (256 consecutive pixels in row-major order: one or two canvas rows).  Each chunk of
  // tile positions is first filtered against it, cooperatively and ORDER-PRESERVING (ballot + prefix), so that a pixel only
  // visits the few tiles that can cover it -- still in increasing tile index, i.e. in the reference's summation order.
  const long long pix_lo = (long long)blockIdx.x * blockDim.x;
  const long long pix_hi = min(HW, pix_lo + (long long)blockDim.x) - 1;
  const int by0 = (int)(pix_lo / W), by1 = (int)(pix_hi / W);
  const int bx0 = by0 == by1 ? (int)(pix_lo - (long long)by0 * W) : 0, bx1 = by0 == by1 ? (int)(pix_hi - (long long)by1 * W) : W - 1;
  __shared__ int sidx[kPosChunk];      // indices (into the chunk) of the tiles that touch the block, increasing
  __shared__ int wcount[kThreads / 32 + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k0 = 0; k0 < n; k0 += kPosChunk) {
    const int kn = min(kPosChunk, n - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kn; i += blockDim.x) spos[i] = pos[k0 + i];
    __syncthreads();
    int nhit = 0;
    for (int base = 0; base < kn; base += blockDim.x) {   // kPosChunk / kThreads rounds
      const int i = base + threadIdx.x;
      bool hit = false;
      if (i < kn) {
        const pisto_tile_pos_t t = spos[i];
        hit = t.y <= by1 && t.y + t.crop_h > by0 && t.x <= bx1 && t.x + t.crop_w > bx0;
      }
      const unsigned int bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) wcount[wid] = __popc(bal);
      __syncthreads();
      int off = nhit;
      for (int w = 0; w < wid; w++) off += wcount[w];
      int tot = nhit;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += wcount[w];
      if (hit) sidx[off + __popc(bal & ((1u << lane) - 1u))] = i;
      nhit = tot;
      __syncthreads();
    }
    if (!live) continue;
    for (int q = 0; q < nhit; q++) {
      const int k = sidx[q];
      const pisto_tile_pos_t t = spos[k];
      const int dy = Y - t.y, dx = X - t.x;
      if (dy < 0 || dx < 0 || dy >= t.crop_h || dx >= t.crop_w) continue;
      const float* src = tiles + (long long)(k0 + k) * tile_elems + (long long)dy * tw + dx;
      float v[C];
#pragma unroll
      for (int c = 0; c < C; c++) v[c] = __ldg(src + (long long)c * th * tw);
      if (softmax) pisto_softmax_inplace<C>(v);
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] = __dadd_rn(acc[c], (double)v[c]);
      cnt = __dadd_rn(cnt, 1.0);
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < C; c++) canvas[c * HW + pix] = acc[c];
    count[pix] = cnt;
  }
}

__global__ void __launch_bounds__(kThreads) normalize_kernel(double* __restrict__ canvas, const double* __restrict__ count, int C,
                                                             long long HW, double min_count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    double d = count ? count[i] : min_count;
    if (count && min_count > 0.0 && d < min_count) d = min_count;
    for (int c = 0; c < C; c++) canvas[c * HW + i] = __ddiv_rn(canvas[c * HW + i], d);
  }
}

__global__ void __launch_bounds__(kThreads) axpy_kernel(double* __restrict__ out, const double* __restrict__ in, long long n, double scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = scale == 1.0 ? in[i] : __dmul_rn(in[i], scale);
    out[i] = __dadd_rn(out[i], v);
  }
}

struct PresentMask { unsigned int bits; int use; };

__global__ void __launch_bounds__(kThreads) argmax_f64_kernel(const double* __restrict__ scores, int C, long long HW, PresentMask pm,
                                                              const uint8_t* __restrict__ gt, int bg_match, int bg_label,
                                                              uint8_t* __restrict__ pred_out, uint8_t* __restrict__ label_out,
                                                              unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int hist[];
  const bool do_conf = conf != nullptr && gt != nullptr;
  if (do_conf) {
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    int bi = 0;
    double bv = 0.0;
    for (int c = 0; c < C; c++) {
      double v = scores[c * HW + i];
      if (pm.use && !((pm.bits >> c) & 1u)) v = -INFINITY;
      bool take = c == 0 || (v > bv) || (v != v && bv == bv);  // np.argmax: first max, NaN is the max
      if (take) { bv = v; bi = c; }
    }
    if (pred_out) pred_out[i] = (uint8_t)bi;
    unsigned int g = gt ? gt[i] : 0xffu;
    if (do_conf && g < (unsigned)C) atomicAdd(&hist[g * C + bi], 1u);
    if (label_out) label_out[i] = (uint8_t)((gt && g == (unsigned)bg_match) ? bg_label : bi);
  }
  if (do_conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (hist[i]) atomicAdd(&conf[i], (unsigned long long)hist[i]);
  }
}

int grid_for(const pisto_ctx* h, long long n, int threads, int per_sm) {
  long long g = (n + threads - 1) / threads;
  long long cap = (long long)h->sm_count * per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace

extern "C" int pisto_stitch_accumulate(pisto_handle_t h, const float* tiles, const pisto_tile_pos_t* pos, int n, int C, int th, int tw,
                                       int softmax, double* canvas, double* count, int H, int W, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_stitch_accumulate: NULL handle");
  PISTO_REQUIRE(C >= 1 && C <= PISTO_MAX_CLASSES, "pisto_stitch_accumulate: C=%d outside [1,%d]", C, PISTO_MAX_CLASSES);
  PISTO_REQUIRE(n >= 0 && th >= 1 && tw >= 1 && H >= 1 && W >= 1, "pisto_stitch_accumulate: bad shape");
  if (n == 0) return PISTO_OK;
  PISTO_REQUIRE(tiles && pos && canvas && count, "pisto_stitch_accumulate: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  long long HW = (long long)H * W;
  int grid = (int)((HW + kThreads - 1) / kThreads);
#define PISTO_ST(CC) case CC: stitch_kernel<CC><<<grid, kThreads, 0, st>>>(tiles, pos, n, th, tw, softmax, canvas, count, H, W); break;
  switch (C) { PISTO_ST(1) PISTO_ST(2) PISTO_ST(3) PISTO_ST(4) PISTO_ST(5) PISTO_ST(6) PISTO_ST(7) PISTO_ST(8) }
#undef PISTO_ST
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_canvas_normalize(pisto_handle_t h, double* canvas, const double* count, int C, int64_t HW, double min_count,
                                      pisto_stream_t stream) {
  PISTO_REQUIRE(h && canvas, "pisto_canvas_normalize: NULL argument");
  PISTO_REQUIRE(C >= 1 && HW >= 0, "pisto_canvas_normalize: bad shape");
  PISTO_REQUIRE(count || min_count != 0.0, "pisto_canvas_normalize: count NULL needs a non-zero constant divisor");
  if (HW == 0) return PISTO_OK;
  PISTO_CUDA(cudaSetDevice(h->device));
  normalize_kernel<<<grid_for(h, HW, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(canvas, count, C, HW, min_count);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_canvas_axpy(pisto_handle_t h, double* out, const double* in, int64_t n, double scale, pisto_stream_t stream) {
  PISTO_REQUIRE(h && out && in, "pisto_canvas_axpy: NULL argument");
  if (n <= 0) return PISTO_OK;
  PISTO_CUDA(cudaSetDevice(h->device));
  axpy_kernel<<<grid_for(h, n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(out, in, n, scale);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_argmax_f64(pisto_handle_t h, const double* scores, int C, int64_t HW, const uint8_t* present, const uint8_t* gt,
                                int bg_match, int bg_label, uint8_t* pred_out, uint8_t* label_out, unsigned long long* conf,
                                pisto_stream_t stream) {
  PISTO_REQUIRE(h && scores, "pisto_argmax_f64: NULL argument");
  PISTO_REQUIRE(C >= 1 && C <= 16, "pisto_argmax_f64: C=%d outside [1,16]", C);
  PISTO_REQUIRE(!(conf && !gt), "pisto_argmax_f64: conf given without gt");
  if (HW <= 0) return PISTO_OK;
  PresentMask pm; pm.bits = 0; pm.use = present != nullptr;
  if (present) for (int c = 0; c < C; c++) if (present[c]) pm.bits |= 1u << c;
  PISTO_CUDA(cudaSetDevice(h->device));
  argmax_f64_kernel<<<grid_for(h, HW, kThreads, 8), kThreads, C * C * sizeof(unsigned int), (cudaStream_t)stream>>>(
      scores, C, HW, pm, gt, bg_match, bg_label, pred_out, label_out, conf);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
