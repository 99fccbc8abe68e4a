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
  const uint32_t* pool_rgba;  // packed pool: r | g << 8 | b << 16 | (bg > 0) << 24 per pixel (pisto_mosaic_pack_pool), or NULL
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

// Per-mosaic tables in shared memory: the plan (272 B) and one entry per grid cell of the 4 composites with everything a
// pixel fetch needs (no dependent global loads on the pixel path: cells -> pool_hw -> pool_off used to be a chain of three).
struct CellEnt {
  long long base;        // pool pixel offset of crop pixel (0, 0): pool_off[t] + (cy - pad_t) * tw + (cx - pad_l)
  int tw, th;            // tile size
  short cye, cxe;        // crop origin in UNPADDED tile coordinates (negative inside the PadIfNeeded border)
  unsigned char label;   // pool_label[t]
  unsigned char padded;  // tile smaller than the patch: REFLECT_101 fix-up needed
  short pad_;
};

// composite coordinate (after flip) -> pool pixel offset; *label gets the tile's class
template <int PS>
__device__ __forceinline__ long long composite_src(const CellEnt* ents_q, int S, int pn, int ps_rt, int flip, int y, int x, unsigned int* label) {
  const int ps = PS > 0 ? PS : ps_rt;
  if (flip & 1) y = S - 1 - y;   // cv2.flip code 0 / -1: rows reversed
  if (flip & 2) x = S - 1 - x;   // cv2.flip code 1 / -1: cols reversed
  const int cr = y / ps, cc = x / ps;
  const int iy = y - cr * ps, ix = x - cc * ps;
  const CellEnt e = ents_q[cr * pn + cc];
  if (label) *label = e.label;
  if (!e.padded) return e.base + iy * e.tw + ix;
  // PadIfNeeded: centred REFLECT_101 pad (rare)
  const int ty = reflect101(e.cye + iy, e.th), tx = reflect101(e.cxe + ix, e.tw);
  return e.base - ((long long)e.cye * e.tw + e.cxe) + (long long)ty * e.tw + tx;
}

// one source pixel as r | g << 8 | b << 16 | (bg > 0) << 24: a single aligned 32-bit load from the packed pool, or 3 + 1 byte
// loads from the planar HWC image / mask pools (want_bg = false skips the mask byte)
template <bool RGBA>
__device__ __forceinline__ unsigned int fetch_px(const MosaicParams& p, long long off, bool want_bg) {
  if (RGBA) return __ldg(p.pool_rgba + off);
  const uint8_t* px = p.pool_img + 3 * off;
  unsigned int v = px[0] | (px[1] << 8) | (px[2] << 16);
  if (want_bg && p.pool_bg && p.pool_bg[off] > 0) v |= 1u << 24;
  return v;
}

struct FastEnt {        // 16-byte view of a cell for the common case (tile at least as large as the patch)
  long long base;
  int tw;
  unsigned int label_padded;  // label | padded << 8
};

struct PixOut { unsigned int rgb; unsigned int mask; };  // r | g << 8 | b << 16

// generic single pixel (cell borders crossed by a padded tile, coordinates outside the index tables): the reference sequence
// flip -> cell -> PadIfNeeded -> tile, tap by tap
template <int PS, bool RGBA>
__device__ __noinline__ PixOut mosaic_pixel_slow(const MosaicParams& p, const CellEnt* ents_q, int S, int pn, int ps, int flip, int warp,
                                                 int yc, int xc, int adelta, int bdelta, int X0b, int Y0b) {
  PixOut o;
  if (!warp) {
    unsigned int m;
    const long long off = composite_src<PS>(ents_q, S, pn, ps, flip, yc, xc, &m);
    const unsigned int v = fetch_px<RGBA>(p, off, true);
    o.rgb = v & 0xffffffu;
    o.mask = (v >> 24) ? (unsigned)p.bg_label : m;
    return o;
  }
  {  // INTER_NEAREST (mask): round_delta = AB_SCALE / 2
    const int sx = sat_short((X0b + 512 + adelta) >> AB_BITS);
    const int sy = sat_short((Y0b + 512 + bdelta) >> AB_BITS);
    unsigned int m;
    const long long off = composite_src<PS>(ents_q, S, pn, ps, flip, reflect101(sy, S), reflect101(sx, S), &m);
    o.mask = (fetch_px<RGBA>(p, off, true) >> 24) ? (unsigned)p.bg_label : m;
  }
  {  // INTER_LINEAR (image): round_delta = AB_SCALE / INTER_TAB_SIZE / 2 = 16
    const int Xl = (X0b + 16 + adelta) >> (AB_BITS - INTER_BITS);
    const int Yl = (Y0b + 16 + bdelta) >> (AB_BITS - INTER_BITS);
    const int sx = sat_short(Xl >> INTER_BITS), sy = sat_short(Yl >> INTER_BITS);
    const int fx = Xl & 31, fy = Yl & 31;
    int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
    const int x0r = reflect101(sx, S), x1r = reflect101(sx + 1, S), y0r = reflect101(sy, S), y1r = reflect101(sy + 1, S);
    const unsigned int v00 = fetch_px<RGBA>(p, composite_src<PS>(ents_q, S, pn, ps, flip, y0r, x0r, nullptr), false);
    const unsigned int v01 = fetch_px<RGBA>(p, composite_src<PS>(ents_q, S, pn, ps, flip, y0r, x1r, nullptr), false);
    const unsigned int v10 = fetch_px<RGBA>(p, composite_src<PS>(ents_q, S, pn, ps, flip, y1r, x0r, nullptr), false);
    const unsigned int v11 = fetch_px<RGBA>(p, composite_src<PS>(ents_q, S, pn, ps, flip, y1r, x1r, nullptr), false);
    o.rgb = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      int v = (int)((v00 >> (8 * ch)) & 0xffu) * w00 + (int)((v01 >> (8 * ch)) & 0xffu) * w01 + (int)((v10 >> (8 * ch)) & 0xffu) * w10 +
              (int)((v11 >> (8 * ch)) & 0xffu) * w11;
      v = (v + (1 << 14)) >> 15;
      o.rgb |= (unsigned int)(v < 0 ? 0 : (v > 255 ? 255 : v)) << (8 * ch);
    }
  }
  return o;
}

// One CTA per mosaic (grid-stride over mosaics); a thread produces 4 consecutive output pixels per step.
//
// Per mosaic, shared memory holds the plan, the cell entries, the fixed-point affine terms of OpenCV's WarpAffineInvoker per
// quadrant (colt[q][x] = {adelta, bdelta}, rowt[q][y] = {X0, Y0}: float64, left-to-right products, round-half-even, evaluated
// once per row / column instead of once per pixel) and two index maps per quadrant that fold REFLECT_101, the flip and the
// division by the patch size:  rowmap[q][y + S] = (cell row * pn) << 16 | row inside the cell  for y in [-S, 2S), colmap alike.
// A tap is then two map loads, one 16-byte cell entry and one multiply-add.  The INTER_NEAREST mask tap is always one of the
// four INTER_LINEAR taps (floor((Z + 512) / 1024) - floor((Z + 16) / 1024) is 0 or 1), so it costs no extra lookup.
template <int PS, bool RGBA>
__global__ void __launch_bounds__(kThreads, 4) mosaic_kernel(const __grid_constant__ MosaicParams p) {
  extern __shared__ __align__(16) unsigned char msm[];
  const int S = p.S, pn = p.pn, pn2 = pn * pn, S3 = 3 * p.S;
  const int ps = PS > 0 ? PS : p.ps;
  pisto_mosaic_plan_t* plan = reinterpret_cast<pisto_mosaic_plan_t*>(msm);
  FastEnt* fents = reinterpret_cast<FastEnt*>(msm + ((sizeof(pisto_mosaic_plan_t) + 15) & ~15u));
  CellEnt* ents = reinterpret_cast<CellEnt*>(fents + 4 * pn2);
  int2* colt = reinterpret_cast<int2*>(ents + 4 * pn2);
  int2* rowt = colt + 4 * S;
  unsigned int* rowmap = reinterpret_cast<unsigned int*>(rowt + 4 * S);
  unsigned int* colmap = rowmap + 4 * S3;
  const int groups_per_row = S / 4;
  const int items = S * groups_per_row;
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
    __syncthreads();  // the previous mosaic's tables are no longer read
    for (int i = threadIdx.x; i < (int)(sizeof(pisto_mosaic_plan_t) / 4); i += blockDim.x)
      reinterpret_cast<uint32_t*>(plan)[i] = reinterpret_cast<const uint32_t*>(p.plans + n)[i];
    for (int i = threadIdx.x; i < 4 * pn2; i += blockDim.x) {
      const pisto_mosaic_cell_t c = p.cells[(long long)n * 4 * pn2 + i];
      const int2 hw = __ldg(reinterpret_cast<const int2*>(p.pool_hw) + c.tile);
      const int th = hw.x, tw = hw.y;
      const int pad_t = th < ps ? (ps - th) >> 1 : 0;   // int((ps - th) / 2.0)
      const int pad_l = tw < ps ? (ps - tw) >> 1 : 0;
      CellEnt e;
      e.tw = tw; e.th = th;
      e.cye = (short)(c.cy - pad_t); e.cxe = (short)(c.cx - pad_l);
      e.base = p.pool_off[c.tile] + (long long)e.cye * tw + e.cxe;
      e.label = p.pool_label[c.tile];
      e.padded = (th < ps || tw < ps) ? 1 : 0;
      e.pad_ = 0;
      ents[i] = e;
      FastEnt f;
      f.base = e.base; f.tw = tw; f.label_padded = e.label | ((unsigned)e.padded << 8);
      fents[i] = f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * S; i += blockDim.x) {
      const int q = i / S, c = i - q * S;
      const pisto_mosaic_quad_t* qd = &plan->quad[q];
      if (qd->warp) {
        const double v = (double)c;
        colt[i] = make_int2(cv_round(__dmul_rn(__dmul_rn(qd->minv[0], v), 1024.0)), cv_round(__dmul_rn(__dmul_rn(qd->minv[3], v), 1024.0)));
        rowt[i] = make_int2(cv_round(__dmul_rn(__dadd_rn(__dmul_rn(qd->minv[1], v), qd->minv[2]), 1024.0)),
                            cv_round(__dmul_rn(__dadd_rn(__dmul_rn(qd->minv[4], v), qd->minv[5]), 1024.0)));
      }
    }
    for (int i = threadIdx.x; i < 4 * S3; i += blockDim.x) {
      const int q = i / S3, c = i - q * S3 - S;   // coordinate in [-S, 2S)
      const int flip = plan->quad[q].flip;
      const int r = reflect101(c, S);
      const int y = (flip & 1) ? S - 1 - r : r, x = (flip & 2) ? S - 1 - r : r;
      const int cr = y / ps, cc = x / ps;
      rowmap[i] = ((unsigned)(cr * pn) << 16) | (unsigned)(y - cr * ps);
      colmap[i] = ((unsigned)cc << 16) | (unsigned)(x - cc * ps);
    }
    __syncthreads();
    const int sh = plan->split_h, sw = plan->split_w;
    const unsigned int inv_gpr = (unsigned int)((0x100000000ull + groups_per_row - 1) / groups_per_row);  // groups_per_row > 1
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      // it / groups_per_row through the 32-bit inverse is exact for it < 2^16 (S <= 512); larger mosaics take the division
      const int Y = (groups_per_row > 1 && items <= 65536) ? (int)__umulhi((unsigned)it, inv_gpr) : it / groups_per_row, gx = it - Y * groups_per_row;
      unsigned int rgb[4], mask_bytes[4];
      // one output pixel, every case (the body used for groups that straddle the vertical split and for the rare slow taps)
      auto pixel = [&](int k, int q, const pisto_mosaic_quad_t* qd, int yc, int xc) {
        const int warp = qd->warp;
        const FastEnt* fq = fents + q * pn2;
        const unsigned int* rm = rowmap + q * S3 + S;
        const unsigned int* cm = colmap + q * S3 + S;
        bool slow = false;
        int adelta = 0, bdelta = 0, X0b = 0, Y0b = 0;
        if (!warp) {
          const unsigned int re = rm[yc], ce = cm[xc];
          const FastEnt e = fq[(re >> 16) + (ce >> 16)];
          if (e.label_padded >> 8) slow = true;
          else {
            const long long off = e.base + (int)(re & 0xffffu) * e.tw + (int)(ce & 0xffffu);
            const unsigned int v = fetch_px<RGBA>(p, off, true);
            rgb[k] = v & 0xffffffu;
            mask_bytes[k] = (v >> 24) ? (unsigned)p.bg_label : e.label_padded;
          }
        } else {
          const int2 cd = colt[q * S + xc], rd = rowt[q * S + yc];
          adelta = cd.x; bdelta = cd.y; X0b = rd.x; Y0b = rd.y;
          const int Xl = (X0b + 16 + adelta) >> (AB_BITS - INTER_BITS);   // INTER_LINEAR: round_delta = 16
          const int Yl = (Y0b + 16 + bdelta) >> (AB_BITS - INTER_BITS);
          const int sx = Xl >> INTER_BITS, sy = Yl >> INTER_BITS;
          const int fx = Xl & 31, fy = Yl & 31;
          if ((unsigned)(sx + S) >= (unsigned)(S3 - 1) || (unsigned)(sy + S) >= (unsigned)(S3 - 1)) slow = true;  // outside the index maps
          else {
            const unsigned int r0 = rm[sy], r1 = rm[sy + 1], c0 = cm[sx], c1 = cm[sx + 1];
            const int iy0 = (int)(r0 & 0xffffu), iy1 = (int)(r1 & 0xffffu), ix0 = (int)(c0 & 0xffffu), ix1 = (int)(c1 & 0xffffu);
            // INTER_NEAREST (mask): round_delta = 512 -> the tap (sy + dy, sx + dx), dy, dx in {0, 1}
            const int dx = ((X0b + 512 + adelta) >> AB_BITS) - sx, dy = ((Y0b + 512 + bdelta) >> AB_BITS) - sy;
            unsigned int v00, v01, v10, v11, m;
            long long onear = 0;
            if ((((r0 ^ r1) | (c0 ^ c1)) >> 16) == 0) {
              // the four taps lie in one grid cell (all but the pixels next to a cell border): one cell entry, one address
              const FastEnt e = fq[(r0 >> 16) + (c0 >> 16)];
              if (e.label_padded >> 8) slow = true;
              const long long o00 = e.base + iy0 * e.tw + ix0;
              const int ox = ix1 - ix0, oy = (iy1 - iy0) * e.tw;
              if (!slow) {
                v00 = fetch_px<RGBA>(p, o00, false); v01 = fetch_px<RGBA>(p, o00 + ox, false);
                v10 = fetch_px<RGBA>(p, o00 + oy, false); v11 = fetch_px<RGBA>(p, o00 + oy + ox, false);
              }
              m = e.label_padded & 0xffu;
              onear = o00 + (dy ? oy : 0) + (dx ? ox : 0);
            } else {
              const FastEnt e00 = fq[(r0 >> 16) + (c0 >> 16)], e01 = fq[(r0 >> 16) + (c1 >> 16)];
              const FastEnt e10 = fq[(r1 >> 16) + (c0 >> 16)], e11 = fq[(r1 >> 16) + (c1 >> 16)];
              if ((e00.label_padded | e01.label_padded | e10.label_padded | e11.label_padded) >> 8) slow = true;
              const long long o00 = e00.base + iy0 * e00.tw + ix0, o01 = e01.base + iy0 * e01.tw + ix1;
              const long long o10 = e10.base + iy1 * e10.tw + ix0, o11 = e11.base + iy1 * e11.tw + ix1;
              if (!slow) {
                v00 = fetch_px<RGBA>(p, o00, false); v01 = fetch_px<RGBA>(p, o01, false);
                v10 = fetch_px<RGBA>(p, o10, false); v11 = fetch_px<RGBA>(p, o11, false);
              }
              m = (dy ? (dx ? e11.label_padded : e10.label_padded) : (dx ? e01.label_padded : e00.label_padded)) & 0xffu;
              onear = dy ? (dx ? o11 : o10) : (dx ? o01 : o00);
            }
            if (!slow) {
              // OpenCV's BilinearTab_i weights are (32-fy)(32-fx)*32 ... fy*fx*32 (the one saturated entry, fx = fy = 0, reproduces
              // the tap itself either way), so  (sum w*p + 2^14) >> 15  ==  ((32-fy)*h0 + fy*h1 + 512) >> 10  with the horizontal sums
              // h = (32-fx)*p_left + fx*p_right <= 8160: exact integer arithmetic, red and blue ride in 16-bit lanes of one word
              const unsigned int fx0 = 32 - fx, fy0 = 32 - fy;
              const unsigned int h0rb = fx0 * (v00 & 0x00ff00ffu) + fx * (v01 & 0x00ff00ffu);
              const unsigned int h1rb = fx0 * (v10 & 0x00ff00ffu) + fx * (v11 & 0x00ff00ffu);
              const unsigned int h0g = fx0 * ((v00 >> 8) & 0xffu) + fx * ((v01 >> 8) & 0xffu);
              const unsigned int h1g = fx0 * ((v10 >> 8) & 0xffu) + fx * ((v11 >> 8) & 0xffu);
              const unsigned int r = (fy0 * (h0rb & 0xffffu) + fy * (h1rb & 0xffffu) + 512u) >> 10;
              const unsigned int b = (fy0 * (h0rb >> 16) + fy * (h1rb >> 16) + 512u) >> 10;
              const unsigned int g = (fy0 * h0g + fy * h1g + 512u) >> 10;
              rgb[k] = r | (g << 8) | (b << 16);
              unsigned int bgbit;
              if (RGBA) bgbit = (dy ? (dx ? v11 : v10) : (dx ? v01 : v00)) >> 24;
              else bgbit = fetch_px<RGBA>(p, onear, true) >> 24;
              mask_bytes[k] = bgbit ? (unsigned)p.bg_label : m;
            }
          }
        }
        if (slow) {
          const PixOut o = mosaic_pixel_slow<PS, RGBA>(p, ents + q * pn2, S, pn, ps, qd->flip, warp, yc, xc, adelta, bdelta, X0b, Y0b);
          rgb[k] = o.rgb; mask_bytes[k] = o.mask;
        }
      };
      const int X0 = gx * 4;
      const int qrow = Y >= sh ? 2 : 0;
      const int ycb = Y >= sh ? Y - sh : Y;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int X = X0 + k;
        const int q = qrow + (X >= sw ? 1 : 0);
        const pisto_mosaic_quad_t* qd = &plan->quad[q];
        pixel(k, q, qd, ycb + qd->crop_y, (X >= sw ? X - sw : X) + qd->crop_x);
      }
      uint32_t* io = reinterpret_cast<uint32_t*>(p.img_out + ((long long)(n * (long long)S + Y) * S + gx * 4) * 3);
      io[0] = rgb[0] | (rgb[1] << 24);
      io[1] = (rgb[1] >> 8) | (rgb[2] << 16);
      io[2] = (rgb[2] >> 16) | (rgb[3] << 8);
      uint32_t* mo = reinterpret_cast<uint32_t*>(p.mask_out + (long long)(n * (long long)S + Y) * S + gx * 4);
      *mo = mask_bytes[0] | (mask_bytes[1] << 8) | (mask_bytes[2] << 16) | (mask_bytes[3] << 24);
    }
  }
}

__global__ void pack_pool_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ bg, long long n, uint32_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* px = img + 3 * i;
    out[i] = px[0] | (px[1] << 8) | (px[2] << 16) | ((bg && bg[i] > 0) ? (1u << 24) : 0u);
  }
}

}  // namespace

extern "C" int pisto_mosaic_pack_pool(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg, int64_t n_px, uint32_t* pool_rgba,
                                      pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_pack_pool: NULL handle");
  PISTO_REQUIRE(n_px >= 0, "pisto_mosaic_pack_pool: bad n_px");
  if (n_px == 0) return PISTO_OK;
  PISTO_REQUIRE(pool_img && pool_rgba, "pisto_mosaic_pack_pool: NULL buffer");
  PISTO_CUDA(cudaSetDevice(h->device));
  pack_pool_kernel<<<h->sm_count * 16, 256, 0, (cudaStream_t)stream>>>(pool_img, pool_bg, n_px, pool_rgba);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

static int mosaic_gather_impl(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg, const uint32_t* pool_rgba, const int64_t* pool_off,
                              const int32_t* pool_hw, const uint8_t* pool_label, const pisto_mosaic_plan_t* plans,
                              const pisto_mosaic_cell_t* cells, int N, int patch_num, int patch_size, int bg_label,
                              uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream);

extern "C" int pisto_mosaic_gather_packed(pisto_handle_t h, const uint32_t* pool_rgba, const int64_t* pool_off, const int32_t* pool_hw,
                                          const uint8_t* pool_label, const pisto_mosaic_plan_t* plans, const pisto_mosaic_cell_t* cells, int N,
                                          int patch_num, int patch_size, int bg_label, uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream) {
  PISTO_REQUIRE(pool_rgba || N == 0, "pisto_mosaic_gather_packed: NULL pool");
  return mosaic_gather_impl(h, nullptr, nullptr, pool_rgba, pool_off, pool_hw, pool_label, plans, cells, N, patch_num, patch_size, bg_label, img_out,
                            mask_out, stream);
}

extern "C" int pisto_mosaic_gather(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg, const int64_t* pool_off,
                                   const int32_t* pool_hw, const uint8_t* pool_label, const pisto_mosaic_plan_t* plans,
                                   const pisto_mosaic_cell_t* cells, int N, int patch_num, int patch_size, int bg_label,
                                   uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream) {
  PISTO_REQUIRE(pool_img || N == 0, "pisto_mosaic_gather: NULL pool");
  return mosaic_gather_impl(h, pool_img, pool_bg, nullptr, pool_off, pool_hw, pool_label, plans, cells, N, patch_num, patch_size, bg_label, img_out,
                            mask_out, stream);
}

static int mosaic_gather_impl(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg, const uint32_t* pool_rgba, const int64_t* pool_off,
                              const int32_t* pool_hw, const uint8_t* pool_label, const pisto_mosaic_plan_t* plans,
                              const pisto_mosaic_cell_t* cells, int N, int patch_num, int patch_size, int bg_label,
                              uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_mosaic_gather: NULL handle");
  PISTO_REQUIRE(N >= 0 && patch_num >= 1 && patch_size >= 1, "pisto_mosaic_gather: bad N/patch_num/patch_size");
  if (N == 0) return PISTO_OK;
  PISTO_REQUIRE(pool_off && pool_hw && pool_label && plans && cells && img_out && mask_out, "pisto_mosaic_gather: NULL buffer");
  const int S = patch_num * patch_size;
  PISTO_REQUIRE(S % 4 == 0 && S <= 8192 && patch_num <= 255, "pisto_mosaic_gather: mosaic side %d must be a multiple of 4 (<= 8192, patch_num <= 255)", S);
  PISTO_REQUIRE((((uintptr_t)img_out | (uintptr_t)mask_out | (uintptr_t)pool_rgba) & 3) == 0, "pisto_mosaic_gather: outputs / packed pool must be 4-byte aligned");
  PISTO_CUDA(cudaSetDevice(h->device));
  MosaicParams p;
  p.pool_img = pool_img; p.pool_bg = pool_bg; p.pool_rgba = pool_rgba; p.pool_off = (const long long*)pool_off; p.pool_hw = pool_hw; p.pool_label = pool_label;
  p.plans = plans; p.cells = cells; p.N = N; p.pn = patch_num; p.ps = patch_size; p.S = S; p.bg_label = bg_label;
  p.img_out = img_out; p.mask_out = mask_out;
  const size_t smem = ((sizeof(pisto_mosaic_plan_t) + 15) & ~(size_t)15) + (sizeof(CellEnt) + sizeof(FastEnt)) * 4 * (size_t)patch_num * patch_num +
                      2 * sizeof(int2) * 4 * (size_t)S + 2 * sizeof(unsigned int) * 4 * 3 * (size_t)S;
  PISTO_REQUIRE(smem <= 200 * 1024, "pisto_mosaic_gather: patch_num %d / side %d too large for the per-mosaic tables", patch_num, S);
  const int grid = N < h->sm_count * 8 ? N : h->sm_count * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define PISTO_MOSAIC_LAUNCH(PS_)                                                                                                         \
  do {                                                                                                                                   \
    if (pool_rgba) {                                                                                                                     \
      if (smem > 48 * 1024) PISTO_CUDA(cudaFuncSetAttribute(mosaic_kernel<PS_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      mosaic_kernel<PS_, true><<<grid, kThreads, smem, st>>>(p);                                                                         \
    } else {                                                                                                                             \
      if (smem > 48 * 1024) PISTO_CUDA(cudaFuncSetAttribute(mosaic_kernel<PS_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      mosaic_kernel<PS_, false><<<grid, kThreads, smem, st>>>(p);                                                                        \
    }                                                                                                                                    \
  } while (0)
  switch (patch_size) {
    case 32: PISTO_MOSAIC_LAUNCH(32); break;
    case 56: PISTO_MOSAIC_LAUNCH(56); break;
    case 112: PISTO_MOSAIC_LAUNCH(112); break;
    default: PISTO_MOSAIC_LAUNCH(0); break;
  }
#undef PISTO_MOSAIC_LAUNCH
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
