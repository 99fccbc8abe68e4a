// pisto_upsample_bilinear: F.interpolate(mode='bilinear', align_corners=False) for float32 and float64 planes.
// Replaces interpolate_tensor (reference infer_pseudo_masks.py:89-90, segmentation_test.py:88-89,197) and the inline
// calls of OEEM/classification/prepare_seg_inputs.py:116,131,137.  Same association as ATen.
//
// HBM-bound for up-sampling (output written once, input re-read through L1/L2).  One thread produces 4 consecutive
// output pixels of one row (16-byte store for f32), the column lerp parameters are recomputed per pixel (cheap next
// to the store), the row parameters once per thread.
#include "common.cuh"

namespace {

template <typename T> struct LerpT { int i0, i1; T l0, l1; };

__device__ __forceinline__ LerpT<float> src_index_t(float scale, int dst, int in_size, bool same) {
  Lerp l = pisto_src_index(scale, dst, in_size, same);
  LerpT<float> r; r.i0 = l.i0; r.i1 = l.i1; r.l0 = l.l0; r.l1 = l.l1; return r;
}
__device__ __forceinline__ LerpT<double> src_index_t(double scale, int dst, int in_size, bool same) {
  LerpT<double> r;
  if (same) { r.i0 = dst; r.i1 = dst; r.l0 = 1.0; r.l1 = 0.0; return r; }
  double src = fmax(__fma_rn(scale, __dadd_rn((double)dst, 0.5), -0.5), 0.0);
  int i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  r.i0 = i0; r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  double l1 = fmin(fmax(__dsub_rn(src, (double)i0), 0.0), 1.0);
  r.l1 = l1; r.l0 = __dsub_rn(1.0, l1);
  return r;
}
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_t(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_t(double a, double b) { return __dmul_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(const T* __restrict__ in, T* __restrict__ out, long long NC, int hi, int wi,
                                                       int ho, int wo, T scale_h, T scale_w) {
  constexpr int PX = 4;
  const int wq = (wo + PX - 1) / PX;
  const long long total = NC * ho * wq;
  const bool same_h = hi == ho, same_w = wi == wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % wq);
    const long long t = idx / wq;
    const int y = (int)(t % ho);
    const long long nc = t / ho;
    const LerpT<T> ly = src_index_t(scale_h, y, hi, same_h);
    const T* r0 = in + (nc * hi + ly.i0) * wi;
    const T* r1 = in + (nc * hi + ly.i1) * wi;
    T* o = out + (nc * ho + y) * wo;
    T res[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) {
      int x = xq * PX + k;
      if (x < wo) {
        const LerpT<T> lx = src_index_t(scale_w, x, wi, same_w);
        T a = r0[lx.i0], b = r0[lx.i1], c = r1[lx.i0], d = r1[lx.i1];
        T h0 = fma_t(lx.l0, a, mul_t(lx.l1, b));
        T h1 = fma_t(lx.l0, c, mul_t(lx.l1, d));
        res[k] = fma_t(ly.l0, h0, mul_t(ly.l1, h1));
      }
    }
#pragma unroll
    for (int k = 0; k < PX; k++)
      if (xq * PX + k < wo) o[xq * PX + k] = res[k];
  }
}

}  // namespace

int pisto_upsample_launch(pisto_ctx* h, const void* in, void* out, long long NC, int hi, int wi, int ho, int wo, int dtype,
                          cudaStream_t st) {
  if (NC == 0) return PISTO_OK;
  long long total = NC * ho * ((wo + 3) / 4);
  long long grid = (total + 255) / 256;
  long long cap = (long long)h->sm_count * 16;
  if (grid > cap) grid = cap;
  if (dtype == 0)
    upsample_kernel<float><<<(int)grid, 256, 0, st>>>((const float*)in, (float*)out, NC, hi, wi, ho, wo, (float)hi / (float)ho,
                                                      (float)wi / (float)wo);
  else
    upsample_kernel<double><<<(int)grid, 256, 0, st>>>((const double*)in, (double*)out, NC, hi, wi, ho, wo, (double)hi / (double)ho,
                                                       (double)wi / (double)wo);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_upsample_bilinear(pisto_handle_t h, const void* in, void* out, int64_t NC, int hi, int wi, int ho, int wo,
                                       int dtype, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_upsample_bilinear: NULL handle");
  PISTO_REQUIRE(dtype == 0 || dtype == 1, "pisto_upsample_bilinear: dtype %d (0 = f32, 1 = f64)", dtype);
  PISTO_REQUIRE(NC >= 0 && hi >= 1 && wi >= 1 && ho >= 1 && wo >= 1, "pisto_upsample_bilinear: bad shape");
  PISTO_REQUIRE(NC == 0 || (in && out), "pisto_upsample_bilinear: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  return pisto_upsample_launch(h, in, out, NC, hi, wi, ho, wo, dtype, (cudaStream_t)stream);
}
