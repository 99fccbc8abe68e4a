// Mosaic planning on the device: the per-cell decisions of CropAndConcatDataset.create_one_image
// (reference create_dataset.ipynb:291-321: random source tile, RandomCrop origin, "background area is smaller than 80 %"
// rejection loop) for N mosaics x 4 composites x patch_num^2 cells.
//
// The reference draws from MT19937 / `random` in albumentations-1.2.1 call order, which cannot be reproduced without that
// library (parity unpinned for the sampling ORDER); here every decision is a pure function of
// (seed, mosaic index, cell, try) through the counter-based Philox4x32-10 generator, so mosaic i is the same for any GPU
// count, batch split or launch order.  pistoseg_b200/philox.py evaluates the identical integer arithmetic with numpy
// (tests compare the two bit for bit).
//
//   tile  = mulhi(r0, P)                    uniform source tile
//   cy    = mulhi(r1, ph - ps + 1)          albumentations RandomCrop: int((H - h + 1) * u)
//   cx    = mulhi(r2, pw - ps + 1)          (ph, pw = tile size padded up to patch_size by PadIfNeeded)
//   accept unless reject_bg and 10 * bg_label * n_bg(crop) >= 8 * ps^2   (the reference sums label VALUES, 3 per bg pixel)
//
// n_bg(crop) comes from a 16-bit summed-area table of the (padded) background mask of every tile, built once per pool by
// pisto_mosaic_bg_integral; 16 bits suffice because the four-corner difference is exact modulo 2^16 and a crop has fewer
// than 65536 pixels.
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ int reflect101p(int p, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  p = p < 0 ? -p : p;
  p %= period;
  return p >= n ? period - p : p;
}

// one block per tile: summed-area table of (bg > 0) over the tile padded to >= ps x ps (centred REFLECT_101), 16-bit
__global__ void bg_integral_kernel(const uint8_t* __restrict__ pool_bg, const long long* __restrict__ pool_off, const int* __restrict__ pool_hw,
                                   const long long* __restrict__ ioff, int ps, unsigned short* __restrict__ integral) {
  const int t = blockIdx.x;
  const int th = pool_hw[2 * t], tw = pool_hw[2 * t + 1];
  const int ph = th < ps ? ps : th, pw = tw < ps ? ps : tw;
  const int top = th < ps ? (ps - th) >> 1 : 0, left = tw < ps ? (ps - tw) >> 1 : 0;
  const uint8_t* bg = pool_bg + pool_off[t];
  unsigned short* I = integral + ioff[t];
  const int W1 = pw + 1;
  for (int x = threadIdx.x; x <= pw; x += blockDim.x) I[x] = 0;
  for (int y = threadIdx.x; y < ph; y += blockDim.x) {  // row prefix sums
    const uint8_t* row = bg + (long long)reflect101p(y - top, th) * tw;
    unsigned short s = 0;
    unsigned short* out = I + (long long)(y + 1) * W1;
    out[0] = 0;
    for (int x = 0; x < pw; x++) { s += row[reflect101p(x - left, tw)] > 0; out[x + 1] = s; }
  }
  __syncthreads();
  for (int x = threadIdx.x + 1; x <= pw; x += blockDim.x) {  // column prefix sums
    unsigned short s = 0;
    for (int y = 1; y <= ph; y++) { s += I[(long long)y * W1 + x]; I[(long long)y * W1 + x] = s; }
  }
}

struct PlanParams {
  unsigned long long seed;
  long long i0, istride;
  int N, pn, ps, P, reject, bg_label, max_tries;
  const int* pool_hw;
  const unsigned short* integral;
  const long long* ioff;
  pisto_mosaic_cell_t* cells;
  unsigned long long* exhausted;  // cells that were still rejected at the last try (the reference loops for ever; pisto_filter_stats slot 0)
};

__global__ void plan_cells_kernel(const __grid_constant__ PlanParams p) {
  const int pn2 = p.pn * p.pn;
  const long long total = (long long)p.N * 4 * pn2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long k = idx / (4 * pn2);
    const int slot = (int)(idx - k * 4 * pn2);  // q * pn^2 + cell
    const unsigned long long i = (unsigned long long)(p.i0 + k * p.istride);
    int tile = 0, cy = 0, cx = 0;
    for (int t = 0; t < p.max_tries; t++) {
      uint32_t r[4];
      philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)slot, (uint32_t)t, (uint32_t)p.seed, (uint32_t)(p.seed >> 32) ^ 0xC3110000u, r);
      tile = (int)mulhi32(r[0], (uint32_t)p.P);
      const int th = p.pool_hw[2 * tile], tw = p.pool_hw[2 * tile + 1];
      const int ph = th < p.ps ? p.ps : th, pw = tw < p.ps ? p.ps : tw;
      cy = (int)mulhi32(r[1], (uint32_t)(ph - p.ps + 1));
      cx = (int)mulhi32(r[2], (uint32_t)(pw - p.ps + 1));
      if (!p.reject) break;
      const unsigned short* I = p.integral + p.ioff[tile];
      const int W1 = pw + 1;
      const unsigned short n = (unsigned short)(I[(long long)(cy + p.ps) * W1 + cx + p.ps] - I[(long long)cy * W1 + cx + p.ps] -
                                                I[(long long)(cy + p.ps) * W1 + cx] + I[(long long)cy * W1 + cx]);
      if (10ll * n * p.bg_label < 8ll * p.ps * p.ps) break;
      if (t == p.max_tries - 1 && p.exhausted) atomicAdd(p.exhausted, 1ull);  // accepted although it fails the test: counted, never silent
    }
    pisto_mosaic_cell_t c;
    c.tile = tile; c.cy = (int16_t)cy; c.cx = (int16_t)cx;
    p.cells[idx] = c;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Per-quadrant decisions of a mosaic (create_dataset.ipynb:336-354 split, :324-330 Flip / ShiftScaleRotate / RandomCrop) on the
// device: one thread per mosaic, the same Philox draws and the same float64 operation sequence as pistoseg_b200/mosaic.py::
// MosaicPlanner.quad_plans (tests compare the two bit for bit).  Every double operation is an explicit round-to-nearest intrinsic
// (no fma contraction), and cos / sin come from pisto_cos_sin_deg below -- a fixed polynomial evaluated with separate multiplies and
// adds -- instead of the platform's libm, so that host (numpy) and device agree to the last bit.
// ---------------------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
__host__ __device__ inline double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

// cos and sin of an angle given in DEGREES: r = angle - 90 * rint(angle / 90) (exact), a = r * (pi / 180), Taylor polynomials of
// degree 16 / 17 on [-pi/4, pi/4] in Horner form with separate multiply and add, then the quadrant rotation.
__host__ __device__ inline void pisto_cos_sin_deg(double angle, double* c, double* s) {
  const double k = rint(angle / 90.0);
  const double r = dadd(angle, -dmul(90.0, k));
  const double a = dmul(r, 3.14159265358979323846 / 180.0);
  const double z = dmul(a, a);
  double pc = 1.0 / 20922789888000.0;                      // 1/16!
  pc = dadd(dmul(pc, z), -1.0 / 87178291200.0);          // 1/14!
  pc = dadd(dmul(pc, z), 1.0 / 479001600.0);             // 1/12!
  pc = dadd(dmul(pc, z), -1.0 / 3628800.0);              // 1/10!
  pc = dadd(dmul(pc, z), 1.0 / 40320.0);                 // 1/8!
  pc = dadd(dmul(pc, z), -1.0 / 720.0);
  pc = dadd(dmul(pc, z), 1.0 / 24.0);
  pc = dadd(dmul(pc, z), -0.5);
  pc = dadd(dmul(pc, z), 1.0);
  double ps = 1.0 / 355687428096000.0;                     // 1/17!
  ps = dadd(dmul(ps, z), -1.0 / 1307674368000.0);        // 1/15!
  ps = dadd(dmul(ps, z), 1.0 / 6227020800.0);            // 1/13!
  ps = dadd(dmul(ps, z), -1.0 / 39916800.0);             // 1/11!
  ps = dadd(dmul(ps, z), 1.0 / 362880.0);                // 1/9!
  ps = dadd(dmul(ps, z), -1.0 / 5040.0);
  ps = dadd(dmul(ps, z), 1.0 / 120.0);
  ps = dadd(dmul(ps, z), -1.0 / 6.0);
  ps = dadd(dmul(ps, z), 1.0);
  ps = dmul(ps, a);
  const int q = ((int)k) & 3;                              // two's complement: -1 & 3 == 3
  *c = q == 0 ? pc : (q == 1 ? -ps : (q == 2 ? -pc : ps));
  *s = q == 0 ? ps : (q == 1 ? pc : (q == 2 ? -ps : -pc));
}

struct QuadParams {
  unsigned long long seed;
  long long i0, istride;
  int N, H, W;
  double p_flip, p_warp, rot_lo, rot_span, scale_lo, scale_span, shift_lo, shift_span;
  pisto_mosaic_plan_t* plans;
};

__device__ __forceinline__ double u53d(uint32_t a, uint32_t b) {
  return (double)(((unsigned long long)(a >> 5) << 26) + (unsigned long long)(b >> 6)) / 9007199254740992.0;
}

__global__ void plan_quads_kernel(const __grid_constant__ QuadParams p) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.N) return;
  const unsigned long long i = (unsigned long long)(p.i0 + k * p.istride);
  const uint32_t ilo = (uint32_t)i, ihi = (uint32_t)(i >> 32), k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32) ^ 0x51AD0000u;
  const int H = p.H, W = p.W;
  uint32_t r[4];
  philox4x32_10(ilo, ihi, 4u, 0u, k0, k1, r);
  long long h = (long long)dmul((double)H, dadd(dmul(u53d(r[0], r[1]), 0.6), 0.2));
  long long w = (long long)dmul((double)W, dadd(dmul(u53d(r[2], r[3]), 0.6), 0.2));
  h += h % 2; w += w % 2;
  pisto_mosaic_plan_t pl;
  pl.split_h = (int)h; pl.split_w = (int)w; pl.reserved[0] = pl.reserved[1] = 0;
  const double cx = (double)W / 2 - 0.5, cy = (double)H / 2 - 0.5;
#pragma unroll 1
  for (int q = 0; q < 4; q++) {
    uint32_t rA[4], rB[4], rC[4], rD[4];
    philox4x32_10(ilo, ihi, (uint32_t)q, 0u, k0, k1, rA);
    philox4x32_10(ilo, ihi, (uint32_t)q, 1u, k0, k1, rB);
    philox4x32_10(ilo, ihi, (uint32_t)q, 2u, k0, k1, rC);
    philox4x32_10(ilo, ihi, (uint32_t)q, 3u, k0, k1, rD);
    pisto_mosaic_quad_t& qd = pl.quad[q];
    qd.flip = ((double)rA[0] / 4294967296.0 < p.p_flip) ? 1 + (int)mulhi32(rA[1], 3u) : 0;
    const bool warp = (double)rA[2] / 4294967296.0 < p.p_warp;
    qd.warp = warp ? 1 : 0;
    const double angle = dadd(p.rot_lo, dmul(p.rot_span, u53d(rB[0], rB[1])));
    const double scale = dadd(p.scale_lo, dmul(p.scale_span, u53d(rB[2], rB[3])));
    const double dx = dadd(p.shift_lo, dmul(p.shift_span, u53d(rC[0], rC[1])));
    const double dy = dadd(p.shift_lo, dmul(p.shift_span, u53d(rC[2], rC[3])));
    const long long hq = q < 2 ? h : H - h, wq = (q & 1) ? W - w : w;
    qd.crop_y = (int)(long long)dmul((double)(H - hq + 1), u53d(rD[0], rD[1]));
    qd.crop_x = (int)(long long)dmul((double)(W - wq + 1), u53d(rD[2], rD[3]));
    // forward matrix (getRotationMatrix2D + shift), then the float64 inversion cv::warpAffine applies
    double co, si;
    pisto_cos_sin_deg(angle, &co, &si);
    const double alpha = dmul(co, scale), beta = dmul(si, scale);
    double m00 = alpha, m01 = beta, m02 = dadd(dmul(dadd(1.0, -alpha), cx), -dmul(beta, cy));
    double m10 = -beta, m11 = alpha, m12 = dadd(dmul(beta, cx), dmul(dadd(1.0, -alpha), cy));
    m02 = dadd(m02, dmul(dx, (double)W));
    m12 = dadd(m12, dmul(dy, (double)H));
    double D = dadd(dmul(m00, m11), -dmul(m01, m10));
    D = D != 0.0 ? 1.0 / D : 0.0;
    const double a11 = dmul(m11, D), a22 = dmul(m00, D);
    m00 = a11; m01 = dmul(m01, -D); m10 = dmul(m10, -D); m11 = a22;
    const double b1 = dadd(dmul(-m00, m02), -dmul(m01, m12));
    const double b2 = dadd(dmul(-m10, m02), -dmul(m11, m12));
    qd.minv[0] = warp ? m00 : 0.0; qd.minv[1] = warp ? m01 : 0.0; qd.minv[2] = warp ? b1 : 0.0;
    qd.minv[3] = warp ? m10 : 0.0; qd.minv[4] = warp ? m11 : 0.0; qd.minv[5] = warp ? b2 : 0.0;
  }
  p.plans[k] = pl;
}

}  // namespace

extern "C" int pisto_mosaic_bg_integral(pisto_handle_t h, const uint8_t* pool_bg, const int64_t* pool_off, const int32_t* pool_hw,
                                        const int64_t* integral_off, int P, int patch_size, uint16_t* integral, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_bg_integral: NULL handle");
  PISTO_REQUIRE(P >= 0 && patch_size >= 1, "pisto_mosaic_bg_integral: bad P/patch_size");
  if (P == 0) return PISTO_OK;
  PISTO_REQUIRE(pool_bg && pool_off && pool_hw && integral_off && integral, "pisto_mosaic_bg_integral: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  bg_integral_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(pool_bg, (const long long*)pool_off, pool_hw, (const long long*)integral_off, patch_size, integral);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_mosaic_plan_cells(pisto_handle_t h, uint64_t seed, int64_t first_index, int64_t index_stride, int N, int patch_num,
                                       int patch_size, int P, const int32_t* pool_hw, const uint16_t* integral, const int64_t* integral_off,
                                       int bg_label, int max_tries, pisto_mosaic_cell_t* cells, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_plan_cells: NULL handle");
  PISTO_REQUIRE(N >= 0 && patch_num >= 1 && patch_size >= 1 && P >= 1, "pisto_mosaic_plan_cells: bad N/patch_num/patch_size/P");
  PISTO_REQUIRE(patch_size * patch_size < 65536, "pisto_mosaic_plan_cells: patch_size %d too large for the 16-bit area table", patch_size);
  PISTO_REQUIRE(!integral == !integral_off, "pisto_mosaic_plan_cells: integral and integral_off go together");
  if (N == 0) return PISTO_OK;
  PISTO_REQUIRE(pool_hw && cells, "pisto_mosaic_plan_cells: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  PlanParams p;
  p.seed = seed; p.i0 = first_index; p.istride = index_stride; p.N = N; p.pn = patch_num; p.ps = patch_size; p.P = P;
  p.reject = integral != nullptr; p.bg_label = bg_label; p.max_tries = max_tries > 0 ? max_tries : 64;
  p.pool_hw = pool_hw; p.integral = integral; p.ioff = (const long long*)integral_off; p.cells = cells;
  p.exhausted = h->stats;
  const long long total = (long long)N * 4 * patch_num * patch_num;
  long long grid = (total + 255) / 256;
  if (grid > (long long)h->sm_count * 32) grid = (long long)h->sm_count * 32;
  plan_cells_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(p);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

extern "C" int pisto_mosaic_plan_quads(pisto_handle_t h, uint64_t seed, int64_t first_index, int64_t index_stride, int N, int patch_num, int patch_size,
                                       double p_flip, double p_warp, double shift_limit, double scale_limit, double rotate_limit,
                                       pisto_mosaic_plan_t* plans, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_plan_quads: NULL handle");
  PISTO_REQUIRE(N >= 0 && patch_num >= 1 && patch_size >= 1, "pisto_mosaic_plan_quads: bad N/patch_num/patch_size");
  if (N == 0) return PISTO_OK;
  PISTO_REQUIRE(plans, "pisto_mosaic_plan_quads: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  QuadParams p;
  p.seed = seed; p.i0 = first_index; p.istride = index_stride; p.N = N; p.H = p.W = patch_num * patch_size;
  p.p_flip = p_flip; p.p_warp = p_warp;
  p.rot_lo = -rotate_limit; p.rot_span = 2 * rotate_limit;            // the host forms the same constants: -limit + (2 * limit) * u
  p.scale_lo = 1 - scale_limit; p.scale_span = 2 * scale_limit;
  p.shift_lo = -shift_limit; p.shift_span = 2 * shift_limit;
  p.plans = plans;
  plan_quads_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

