// pisto_mosaic_gather: mosaic dataset synthesis, plan in -> pixels out.
// Replaces CropAndConcatDataset.__getitem__ (reference create_dataset.ipynb:273-372 [cell 9];
// create_dataset_bcss.ipynb:257-340 [cell 8]): 4 grid-of-crops composites, per-quadrant cv2.flip ->
// cv2.warpAffine (INTER_LINEAR image / INTER_NEAREST mask, BORDER_REFLECT_101) -> RandomCrop -> paste.
//
// Every output pixel is inverse-mapped through  quadrant -> crop offset -> fixed-point affine (OpenCV's integer
// algorithm, oracle/warp_affine.py) -> flip -> grid cell -> PadIfNeeded (REFLECT_101) -> source tile  and gathered
// directly from the tile pool; the four 224x224 intermediate composites (and the 3 of 4 warped quadrants the
// reference computes and throws away) are never materialised.  All arithmetic is integer except the float64
// evaluation of the affine map, whose operation order equals OpenCV's (bit-exact rounding to fixed point).
//
// HBM-bound gather: one thread produces 4 consecutive output pixels (12 image bytes = 3 aligned 32-bit stores,
// 4 mask bytes = 1 store).  Reads are 1 (no warp) or 4 (warp) taps of 3 bytes; neighbouring threads read neighbouring
// source pixels of the same tile crop, so sectors are shared through L1.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int AB_BITS = 10;
constexpr int INTER_BITS = 5;

struct MosaicParams {
  const uint8_t* pool_img;
  const uint8_t* pool_bg;
  const long long* pool_off;
  const int* pool_hw;
  const uint8_t* pool_label;
  const pisto_mosaic_plan_t* plans;
  const pisto_mosaic_cell_t* cells;
  int N, pn, ps, S, bg_label;
  uint8_t* img_out;
  uint8_t* mask_out;
};

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  p = p < 0 ? -p : p;
  if (p >= period) p %= period;  // rare: overshoot by more than one image size
  return p >= n ? period - p : p;
}

__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// double -> int with round-half-even and saturation (cv::saturate_cast<int>(double) == cvRound)
__device__ __forceinline__ int cv_round(double v) {
  double r = rint(v);
  if (r >= 2147483647.0) return 2147483647;
  if (r <= -2147483648.0) return (int)0x80000000;
  return (int)r;
}

struct SrcPix { long long off; int t; };  // pixel offset into the pool (pool_off[t] + ty*tw + tx), tile id

// composite coordinate (after flip) -> pool pixel.  PS > 0: patch size known at compile time (division by constant).
template <int PS>
__device__ __forceinline__ SrcPix composite_src(const MosaicParams& p, const pisto_mosaic_cell_t* cells_q, int flip, int y, int x) {
  const int ps = PS > 0 ? PS : p.ps;
  if (flip & 1) y = p.S - 1 - y;   // cv2.flip code 0 / -1: rows reversed
  if (flip & 2) x = p.S - 1 - x;   // cv2.flip code 1 / -1: cols reversed
  const int cr = y / ps, cc = x / ps;
  const int iy = y - cr * ps, ix = x - cc * ps;
  const pisto_mosaic_cell_t cell = cells_q[cr * p.pn + cc];
  const int2 hw = __ldg(reinterpret_cast<const int2*>(p.pool_hw) + cell.tile);
  const int th = hw.x, tw = hw.y;
  int ty = cell.cy + iy, tx = cell.cx + ix;
  if (th < ps || tw < ps) {  // PadIfNeeded: centred REFLECT_101 pad (rare)
    const int pad_t = th < ps ? (ps - th) >> 1 : 0;   // int((ps - th) / 2.0)
    const int pad_l = tw < ps ? (ps - tw) >> 1 : 0;
    ty = reflect101(ty - pad_t, th);
    tx = reflect101(tx - pad_l, tw);
  }
  SrcPix s;
  s.t = cell.tile;
  s.off = p.pool_off[cell.tile] + (long long)ty * tw + tx;
  return s;
}

template <int PS>
__global__ void __launch_bounds__(kThreads) mosaic_kernel(const __grid_constant__ MosaicParams p) {
  const int S = p.S;
  const int groups_per_row = S / 4;
  const long long total = (long long)p.N * S * groups_per_row;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(idx % groups_per_row);
    const long long t = idx / groups_per_row;
    const int Y = (int)(t % S);
    const int n = (int)(t / S);
    const pisto_mosaic_plan_t* plan = p.plans + n;
    const int sh = plan->split_h, sw = plan->split_w;
    unsigned int img_bytes[12];
    unsigned int mask_bytes[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int X = gx * 4 + k;
      const int q = (Y >= sh ? 2 : 0) + (X >= sw ? 1 : 0);
      const pisto_mosaic_quad_t* qd = &plan->quad[q];
      const pisto_mosaic_cell_t* cells_q = p.cells + ((long long)n * 4 + q) * p.pn * p.pn;
      const int yc = (Y >= sh ? Y - sh : Y) + qd->crop_y;
      const int xc = (X >= sw ? X - sw : X) + qd->crop_x;
      const int flip = qd->flip;
      if (!qd->warp) {
        SrcPix s = composite_src<PS>(p, cells_q, flip, yc, xc);
        const uint8_t* px = p.pool_img + 3 * s.off;
        img_bytes[3 * k + 0] = px[0]; img_bytes[3 * k + 1] = px[1]; img_bytes[3 * k + 2] = px[2];
        unsigned int m = p.pool_label[s.t];
        if (p.pool_bg && p.pool_bg[s.off] > 0) m = p.bg_label;
        mask_bytes[k] = m;
      } else {
        const double m0 = qd->minv[0], m1 = qd->minv[1], m2 = qd->minv[2], m3 = qd->minv[3], m4 = qd->minv[4], m5 = qd->minv[5];
        // OpenCV WarpAffineInvoker: adelta/bdelta per column, X0/Y0 per row (all float64, left-to-right products)
        const int adelta = cv_round(__dmul_rn(__dmul_rn(m0, (double)xc), 1024.0));
        const int bdelta = cv_round(__dmul_rn(__dmul_rn(m3, (double)xc), 1024.0));
        const int X0b = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)yc), m2), 1024.0));
        const int Y0b = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)yc), m5), 1024.0));
        {  // INTER_NEAREST (mask): round_delta = AB_SCALE / 2
          const int sx = sat_short((X0b + 512 + adelta) >> AB_BITS);
          const int sy = sat_short((Y0b + 512 + bdelta) >> AB_BITS);
          SrcPix s = composite_src<PS>(p, cells_q, flip, reflect101(sy, S), reflect101(sx, S));
          unsigned int m = p.pool_label[s.t];
          if (p.pool_bg && p.pool_bg[s.off] > 0) m = p.bg_label;
          mask_bytes[k] = m;
        }
        {  // INTER_LINEAR (image): round_delta = AB_SCALE / INTER_TAB_SIZE / 2 = 16
          const int Xl = (X0b + 16 + adelta) >> (AB_BITS - INTER_BITS);
          const int Yl = (Y0b + 16 + bdelta) >> (AB_BITS - INTER_BITS);
          const int sx = sat_short(Xl >> INTER_BITS), sy = sat_short(Yl >> INTER_BITS);
          const int fx = Xl & 31, fy = Yl & 31;
          // 2x2 int16 weights of OpenCV's BilinearTab_i (closed form; the one saturated entry gets its fix-up)
          int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
          if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
          const int x0r = reflect101(sx, S), x1r = reflect101(sx + 1, S), y0r = reflect101(sy, S), y1r = reflect101(sy + 1, S);
          const uint8_t* p00 = p.pool_img + 3 * composite_src<PS>(p, cells_q, flip, y0r, x0r).off;
          const uint8_t* p01 = p.pool_img + 3 * composite_src<PS>(p, cells_q, flip, y0r, x1r).off;
          const uint8_t* p10 = p.pool_img + 3 * composite_src<PS>(p, cells_q, flip, y1r, x0r).off;
          const uint8_t* p11 = p.pool_img + 3 * composite_src<PS>(p, cells_q, flip, y1r, x1r).off;
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            int v = (int)p00[ch] * w00 + (int)p01[ch] * w01 + (int)p10[ch] * w10 + (int)p11[ch] * w11;
            v = (v + (1 << 14)) >> 15;
            img_bytes[3 * k + ch] = (unsigned int)(v < 0 ? 0 : (v > 255 ? 255 : v));
          }
        }
      }
    }
    uint32_t* io = reinterpret_cast<uint32_t*>(p.img_out + ((long long)(n * (long long)S + Y) * S + gx * 4) * 3);
#pragma unroll
    for (int wd = 0; wd < 3; wd++)
      io[wd] = img_bytes[4 * wd] | (img_bytes[4 * wd + 1] << 8) | (img_bytes[4 * wd + 2] << 16) | (img_bytes[4 * wd + 3] << 24);
    uint32_t* mo = reinterpret_cast<uint32_t*>(p.mask_out + (long long)(n * (long long)S + Y) * S + gx * 4);
    *mo = mask_bytes[0] | (mask_bytes[1] << 8) | (mask_bytes[2] << 16) | (mask_bytes[3] << 24);
  }
}

}  // namespace

extern "C" int pisto_mosaic_gather(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg, const int64_t* pool_off,
                                   const int32_t* pool_hw, const uint8_t* pool_label, const pisto_mosaic_plan_t* plans,
                                   const pisto_mosaic_cell_t* cells, int N, int patch_num, int patch_size, int bg_label,
                                   uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_gather: NULL handle");
  PISTO_REQUIRE(N >= 0 && patch_num >= 1 && patch_size >= 1, "pisto_mosaic_gather: bad N/patch_num/patch_size");
  if (N == 0) return PISTO_OK;
  PISTO_REQUIRE(pool_img && pool_off && pool_hw && pool_label && plans && cells && img_out && mask_out, "pisto_mosaic_gather: NULL buffer");
  const int S = patch_num * patch_size;
  PISTO_REQUIRE(S % 4 == 0 && S <= 16384, "pisto_mosaic_gather: mosaic side %d must be a multiple of 4 (<= 16384)", S);
  PISTO_REQUIRE((((uintptr_t)img_out | (uintptr_t)mask_out) & 3) == 0, "pisto_mosaic_gather: outputs must be 4-byte aligned");
  PISTO_CUDA(cudaSetDevice(h->device));
  MosaicParams p;
  p.pool_img = pool_img; p.pool_bg = pool_bg; p.pool_off = (const long long*)pool_off; p.pool_hw = pool_hw; p.pool_label = pool_label;
  p.plans = plans; p.cells = cells; p.N = N; p.pn = patch_num; p.ps = patch_size; p.S = S; p.bg_label = bg_label;
  p.img_out = img_out; p.mask_out = mask_out;
  long long total = (long long)N * S * (S / 4);
  long long grid = (total + kThreads - 1) / kThreads;
  long long cap = (long long)h->sm_count * 16;
  if (grid > cap) grid = cap;
  cudaStream_t st = (cudaStream_t)stream;
  switch (patch_size) {
    case 32: mosaic_kernel<32><<<(int)grid, kThreads, 0, st>>>(p); break;
    case 56: mosaic_kernel<56><<<(int)grid, kThreads, 0, st>>>(p); break;
    case 112: mosaic_kernel<112><<<(int)grid, kThreads, 0, st>>>(p); break;
    default: mosaic_kernel<0><<<(int)grid, kThreads, 0, st>>>(p); break;
  }
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
