// pisto_upsample_bilinear: F.interpolate(mode='bilinear', align_corners=False) for float32 and float64 planes.
// Replaces interpolate_tensor (reference infer_pseudo_masks.py:89-90, segmentation_test.py:88-89,197) and the inline
// calls of OEEM/classification/prepare_seg_inputs.py:116,131,137.  Same association as ATen.
//
// HBM-bound for up-sampling (output written once, input re-read through L1/L2).
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

// One thread owns 4 adjacent output columns of one plane and streams down a strip of output rows.  The column lerp
// parameters are computed once per thread; per source row the horizontally interpolated values of the thread's columns are
// kept in registers (Ha: row i0, Hb: row i1) and only reloaded when the output row moves to another source-row pair, so an
// output row costs 4 multiplies + 4 fused multiply-adds and one 16-byte (f32) / two 16-byte (f64) stores.  The expression
// fma(ly0, fma(lx0, a, lx1*b), ly1 * fma(lx0, c, lx1*d)) and its rounding are unchanged.
//
// FUSED (float64 only): the big-mask epilogue of segmentation_test.py:187-199 / prepare_seg_inputs.py:128-134 in one pass --
// every source value is canvas / count (count clamped below by min_count when > 0) formed on the fly, and the result is added
// to `out` (accumulate) instead of stored: normalise + resize + sum over scales without the two intermediate canvases.  The
// operations per value (one division, the bilinear expression, one addition) and their order are the reference's.
template <typename T, bool FUSED>
__global__ void __launch_bounds__(256) upsample_kernel(const T* __restrict__ in, T* __restrict__ out, long long NC, int hi, int wi,
                                                       int ho, int wo, T scale_h, T scale_w, int strips, int rows_per_strip, int xblocks,
                                                       int vec_ok, const T* __restrict__ count, T min_count, int accumulate) {
  constexpr int PX = 4;
  const bool same_h = hi == ho, same_w = wi == wo;
  const long long nblocks = NC * strips * xblocks;
  for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int xb = (int)(blk % xblocks);
    const long long t = blk / xblocks;
    const int strip = (int)(t % strips);
    const long long nc = t / strips;
    const int x0 = (xb * (int)blockDim.x + (int)threadIdx.x) * PX;
    if (x0 >= wo) continue;
    LerpT<T> lx[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) lx[k] = src_index_t(scale_w, min(x0 + k, wo - 1), wi, same_w);
    const T* plane = in + nc * (long long)hi * wi;
    T Ha[PX], Hb[PX];
    int p0 = -1, p1 = -1;
    auto load_row = [&](int i, T (&H)[PX]) {
      const T* r = plane + (long long)i * wi;
      if constexpr (FUSED) {
        const T* cr = count + (long long)i * wi;
        auto tap = [&](int j) {
          T d = __ldg(cr + j);
          if (min_count > (T)0 && d < min_count) d = min_count;
          return __ldg(r + j) / d;   // IEEE division (no fast-math): canvas /= count of the reference
        };
#pragma unroll
        for (int k = 0; k < PX; k++) H[k] = fma_t(lx[k].l0, tap(lx[k].i0), mul_t(lx[k].l1, tap(lx[k].i1)));
      } else {
#pragma unroll
        for (int k = 0; k < PX; k++) H[k] = fma_t(lx[k].l0, __ldg(r + lx[k].i0), mul_t(lx[k].l1, __ldg(r + lx[k].i1)));
      }
    };
    const int ys = strip * rows_per_strip, ye = min(ho, ys + rows_per_strip);
    T* o = out + (nc * ho + ys) * (long long)wo + x0;
    for (int y = ys; y < ye; y++, o += wo) {
      const LerpT<T> ly = src_index_t(scale_h, y, hi, same_h);
      if (ly.i0 != p0 || ly.i1 != p1) {
        if (ly.i0 == p1 && p1 >= 0) {
#pragma unroll
          for (int k = 0; k < PX; k++) Ha[k] = Hb[k];
        } else {
          load_row(ly.i0, Ha);
        }
        if (ly.i1 == ly.i0) {
#pragma unroll
          for (int k = 0; k < PX; k++) Hb[k] = Ha[k];
        } else {
          load_row(ly.i1, Hb);
        }
        p0 = ly.i0; p1 = ly.i1;
      }
      T res[PX];
#pragma unroll
      for (int k = 0; k < PX; k++) res[k] = fma_t(ly.l0, Ha[k], mul_t(ly.l1, Hb[k]));
      if (FUSED && accumulate) {
#pragma unroll
        for (int k = 0; k < PX; k++)
          if (x0 + k < wo) res[k] = o[k] + res[k];
      }
      if (vec_ok) {
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(o) = make_float4((float)res[0], (float)res[1], (float)res[2], (float)res[3]);
        } else {
          reinterpret_cast<double2*>(o)[0] = make_double2((double)res[0], (double)res[1]);
          reinterpret_cast<double2*>(o)[1] = make_double2((double)res[2], (double)res[3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < PX; k++)
          if (x0 + k < wo) o[k] = res[k];
      }
    }
  }
}

}  // namespace

int pisto_upsample_launch(pisto_ctx* h, const void* in, void* out, long long NC, int hi, int wi, int ho, int wo, int dtype,
                          cudaStream_t st, const double* count, double min_count, int accumulate) {
  if (NC == 0) return PISTO_OK;
  const int wq = (wo + 3) / 4;
  const int threads = wq >= 256 ? 256 : ((wq + 31) / 32) * 32;
  const int xblocks = (wq + threads - 1) / threads;
  // enough blocks to fill the machine, strips as long as possible (every strip start reloads two source rows)
  long long want = (long long)h->sm_count * 16;
  int strips = (int)((want + NC * xblocks - 1) / (NC * xblocks));
  if (strips < 1) strips = 1;
  if (strips > ho) strips = ho;
  int rows_per_strip = (ho + strips - 1) / strips;
  strips = (ho + rows_per_strip - 1) / rows_per_strip;
  long long nblocks = NC * strips * xblocks;
  long long grid = nblocks < want ? nblocks : want;
  const size_t esz = dtype == 0 ? 4 : 8;
  const int vec_ok = (wo % 4 == 0) && (((uintptr_t)out & 15) == 0) && ((wo * esz) % 16 == 0);
  if (dtype == 0)
    upsample_kernel<float, false><<<(int)grid, threads, 0, st>>>((const float*)in, (float*)out, NC, hi, wi, ho, wo, (float)hi / (float)ho,
                                                                 (float)wi / (float)wo, strips, rows_per_strip, xblocks, vec_ok, nullptr, 0.f, 0);
  else if (!count)
    upsample_kernel<double, false><<<(int)grid, threads, 0, st>>>((const double*)in, (double*)out, NC, hi, wi, ho, wo, (double)hi / (double)ho,
                                                                  (double)wi / (double)wo, strips, rows_per_strip, xblocks, vec_ok, nullptr, 0.0, 0);
  else
    upsample_kernel<double, true><<<(int)grid, threads, 0, st>>>((const double*)in, (double*)out, NC, hi, wi, ho, wo, (double)hi / (double)ho,
                                                                 (double)wi / (double)wo, strips, rows_per_strip, xblocks, vec_ok, count, min_count, accumulate);
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
  return pisto_upsample_launch(h, in, out, NC, hi, wi, ho, wo, dtype, (cudaStream_t)stream, nullptr, 0.0, 0);
}

extern "C" int pisto_canvas_resize_accumulate(pisto_handle_t h, const double* canvas, const double* count, int C, int hi, int wi,
                                              double min_count, double* out, int ho, int wo, int accumulate, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_canvas_resize_accumulate: NULL handle");
  PISTO_REQUIRE(C >= 1 && hi >= 1 && wi >= 1 && ho >= 1 && wo >= 1, "pisto_canvas_resize_accumulate: bad shape");
  PISTO_REQUIRE(canvas && count && out, "pisto_canvas_resize_accumulate: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  return pisto_upsample_launch(h, canvas, out, C, hi, wi, ho, wo, 1, (cudaStream_t)stream, count, min_count, accumulate);
}
