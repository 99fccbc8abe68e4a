// Block-tiled streaming fusion kernel: the path for tiles too large to stage whole (BASELINE config 5) and for
// full-resolution view sets.  Whole-tile shapes (configs 1-3) go to fuse_stream.cu, which adds the TMA pipeline,
// packed f32x2 arithmetic and dynamic tile scheduling on top of the same scheme.
//
// Work item = (tile n, output block by x bx).  A CTA walks its items persistently:
//   A. per view, the raw (augmented-frame) sub-rectangle of low-resolution logits that the block needs is staged in
//      shared memory (whole views for 224-px tiles: 58.8 KB for cfg 2);  de-augmentation (flip / rot90) is NOT
//      materialised -- it is an affine index map applied when the staged values are read;
//   B. two small tables are built in shared memory: per (view, output row) the vertical lerp weights + the shared-
//      memory row offsets of the two source rows (+ a 2-bit "advance" flag hidden in the weights' sign bits), and per
//      (view, output column) the horizontal lerp weights + column offsets;
//   C. every thread owns COLS adjacent output columns and streams down a strip of rows.  For each view it keeps, in
//      registers, the HORIZONTALLY interpolated values of the two source rows that bracket the current output row
//      (Ha, Hb: 2*C*COLS registers per view).  Per output row and view the work is then exactly
//          t = l1*Hb;  o = fma(l0, Ha, t);  acc = acc + o            (3 FP32 ops per class and pixel)
//      and only when the source-row pair advances (every ~8 rows for stride-8 logits) two shared-memory reads and
//      one fma per value refresh Hb.  This is the separable form of torch's
//          fma(h0, fma(w0,a, w1*b), h1 * fma(w0,c, w1*d))
//      with identical association, hence bit-identical results at ~3.4 instead of ~12 instructions
//      per (pixel, class, view).
//   D. per pixel: mask / argmax (pisto_decide: margin fast path, exact slow path), confusion counters packed in two
//      64-bit registers, background overwrite, 2-byte label store, optional fused-score / 32x32 logit export.
//
// The kernel is FP32-issue-bound for V >= 2 (see DESIGN.md "Roofline"); HBM traffic is the compulsory minimum
// (every input byte is read once, every output byte written once).
#include "fuse_common.cuh"

namespace {

constexpr int kMaxThreads = 448;

struct SubRect {
  int a_lo, b_lo, nrows, pitch;
  int base2, si2, sj2, plane;
};

struct StreamGeom {
  int BW, BH, nbx, nby;
  int GX, S, rows_per_strip;
  int view_off[PISTO_MAX_VIEWS];  // float offset of each view's staging area
  int view_cap[PISTO_MAX_VIEWS];  // floats
  int sr_off, rowtab_off, coltab_off, views_off, hist_off;  // byte offsets into dynamic smem
  int smem_bytes;
  long long n_items;
};

__device__ __forceinline__ void minmax2(int u, int v, int& lo, int& hi) { lo = u < v ? u : v; hi = u < v ? v : u; }

template <int C, int V, int COLS>
__global__ void __launch_bounds__(kMaxThreads, 1) fuse_block_kernel(const __grid_constant__ FuseParams p,
                                                                     const __grid_constant__ StreamGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SubRect* sr = reinterpret_cast<SubRect*>(smem_raw + g.sr_off);
  float4* rowtab = reinterpret_cast<float4*>(smem_raw + g.rowtab_off);  // [V][BH]
  float4* coltab = reinterpret_cast<float4*>(smem_raw + g.coltab_off);  // [V][BW]
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);
  unsigned int* hist = reinterpret_cast<unsigned int*>(smem_raw + g.hist_off);  // [C*C]

  const int tid = threadIdx.x, nt = blockDim.x;
  const bool do_conf = p.conf != nullptr && p.gt != nullptr;
  constexpr int BINS = C * C;
  if (do_conf) {
    for (int i = tid; i < BINS; i += nt) hist[i] = 0;
  }
  const bool worker = tid < g.GX * g.S;
  const int grp = tid % g.GX, strip = tid / g.GX;
  const int xl0 = grp * COLS;

  for (long long item = blockIdx.x; item < g.n_items; item += gridDim.x) {
    const int bx = (int)(item % g.nbx);
    const int by = (int)((item / g.nbx) % g.nby);
    const int n = (int)(item / ((long long)g.nbx * g.nby));
    const int y0 = by * g.BH, x0 = bx * g.BW;
    const int bh = min(g.BH, p.T_h - y0), bw = min(g.BW, p.T_w - x0);
    const TilePresence tp = pisto_tile_presence(p, n);
    const bool need_scores = tp.single < 0 || p.fused_out != nullptr;
    const bool need_low = p.lowres_out != nullptr && p.low_fh > 0;

    __syncthreads();  // previous item's readers are done with the staging area / tables
    if (need_scores || need_low) {
      // ---- A. sub-rectangles ---------------------------------------------------------------------------------
      if (tid < V) {
        const ViewDev& vw = p.view[tid];
        const ViewMap& m = vw.map;
        int r_lo = pisto_src_index(vw.scale_h, y0, m.ho, vw.same_h).i0;
        int r_hi = pisto_src_index(vw.scale_h, y0 + bh - 1, m.ho, vw.same_h).i1;
        int c_lo = pisto_src_index(vw.scale_w, x0, m.wo, vw.same_w).i0;
        int c_hi = pisto_src_index(vw.scale_w, x0 + bw - 1, m.wo, vw.same_w).i1;
        int l1, h1, l2, h2;
        minmax2(m.ai * r_lo, m.ai * r_hi, l1, h1);
        minmax2(m.aj * c_lo, m.aj * c_hi, l2, h2);
        int a_lo = m.a0 + l1 + l2, a_hi = m.a0 + h1 + h2;
        minmax2(m.bi * r_lo, m.bi * r_hi, l1, h1);
        minmax2(m.bj * c_lo, m.bj * c_hi, l2, h2);
        int b_lo = m.b0 + l1 + l2, b_hi = m.b0 + h1 + h2;
        SubRect s;
        s.a_lo = a_lo; s.b_lo = b_lo; s.nrows = a_hi - a_lo + 1; s.pitch = b_hi - b_lo + 1;
        s.plane = s.nrows * s.pitch;
        s.base2 = (m.a0 - a_lo) * s.pitch + (m.b0 - b_lo);
        s.si2 = m.ai * s.pitch + m.bi;
        s.sj2 = m.aj * s.pitch + m.bj;
        if (C * s.plane > g.view_cap[tid]) __trap();  // host geometry and device geometry disagree: never expected
        sr[tid] = s;
      }
      __syncthreads();
      // ---- B. staging + tables -------------------------------------------------------------------------------
#pragma unroll 1
      for (int v = 0; v < V; v++) {
        const ViewDev& vw = p.view[v];
        const SubRect s = sr[v];
        const float* src = vw.logits + (long long)n * vw.tile_stride;
        float* dst = vsm + g.view_off[v];
        if (s.pitch == vw.w) {
          // full-width rows: each class plane is one contiguous run (the whole view when nrows == h)
          const int run = s.plane;
          for (int c = 0; c < C; c++) {
            const float* sp = src + (long long)c * vw.h * vw.w + s.a_lo * vw.w;
            float* dp = dst + c * run;
#pragma unroll 4
            for (int i = tid; i < run; i += nt) dp[i] = __ldg(sp + i);
          }
        } else {
          const int rows = C * s.nrows;
          for (int r = tid / 32; r < rows; r += nt / 32) {
            int c = r / s.nrows, a = r - c * s.nrows;
            const float* sp = src + (long long)c * vw.h * vw.w + (long long)(s.a_lo + a) * vw.w + s.b_lo;
            float* dp = dst + r * s.pitch;
            for (int b = tid & 31; b < s.pitch; b += 32) dp[b] = __ldg(sp + b);
          }
        }
      }
      for (int i = tid; i < V * bh; i += nt) {
        int v = i / bh, yl = i - v * bh;
        const ViewDev& vw = p.view[v];
        const SubRect s = sr[v];
        Lerp L = pisto_src_index(vw.scale_h, y0 + yl, vw.map.ho, vw.same_h);
        int flag = 0;
        if (yl % g.rows_per_strip != 0) {
          Lerp P = pisto_src_index(vw.scale_h, y0 + yl - 1, vw.map.ho, vw.same_h);
          if (P.i0 != L.i0 || P.i1 != L.i1) flag = (L.i0 == P.i1 && P.i1 == P.i0 + 1) ? 1 : 2;
        }
        float4 e;
        e.x = __int_as_float(__float_as_int(L.l0) | ((flag & 1) << 31));
        e.y = __int_as_float(__float_as_int(L.l1) | ((flag >> 1) << 31));
        e.z = __int_as_float(s.base2 + L.i0 * s.si2);
        e.w = __int_as_float(s.base2 + L.i1 * s.si2);
        rowtab[v * g.BH + yl] = e;
      }
      for (int i = tid; i < V * bw; i += nt) {
        int v = i / bw, xl = i - v * bw;
        const ViewDev& vw = p.view[v];
        const SubRect s = sr[v];
        Lerp L = pisto_src_index(vw.scale_w, x0 + xl, vw.map.wo, vw.same_w);
        float4 e;
        e.x = __int_as_float(L.i0 * s.sj2);
        e.y = __int_as_float(L.i1 * s.sj2);
        e.z = L.l0;
        e.w = L.l1;
        coltab[v * g.BW + xl] = e;
      }
    }
    __syncthreads();

    // horizontally interpolated values of one staged source row, for this thread's columns
    auto hrow = [&](int v, int rowoff, float (&H)[C][COLS]) {
      const float* base = vsm + g.view_off[v] + rowoff;
      const int plane = sr[v].plane;
#pragma unroll
      for (int col = 0; col < COLS; col++) {
        const float4 ct = coltab[v * g.BW + xl0 + col];
        const int oa = __float_as_int(ct.x), ob = __float_as_int(ct.y);
#pragma unroll
        for (int c = 0; c < C; c++) {
          float va = base[c * plane + oa], vb = base[c * plane + ob];
          H[c][col] = __fmaf_rn(ct.z, va, __fmul_rn(ct.w, vb));
        }
      }
    };

    unsigned long long cnt_lo = 0, cnt_hi = 0;
    const bool col_ok = worker && (xl0 + COLS <= bw);
    const int ys = strip * g.rows_per_strip;
    const int ye = min(ys + g.rows_per_strip, bh);

    if (col_ok && ys < ye && need_scores) {
      // ---- C. stream down the strip ----------------------------------------------------------------------------
      float Ha[V][C][COLS], Hb[V][C][COLS];
#pragma unroll
      for (int v = 0; v < V; v++) {
        const float4 e = rowtab[v * g.BH + ys];
        hrow(v, __float_as_int(e.z), Ha[v]);
        hrow(v, __float_as_int(e.w), Hb[v]);
      }
      const int x = x0 + xl0;
#pragma unroll 1
      for (int yl = ys; yl < ye; yl++) {
        const int y = y0 + yl;
        const long long pix = ((long long)n * p.T_h + y) * p.T_w + x;
        // issue the byte loads early; they are consumed after the view loop
        unsigned int bgv[COLS], gtv[COLS];
#pragma unroll
        for (int col = 0; col < COLS; col++) {
          bgv[col] = p.bg ? (unsigned int)__ldg(p.bg + pix + col) : 0xffffffffu;
          gtv[col] = do_conf ? (unsigned int)__ldg(p.gt + pix + col) : 0xffu;
        }
        float acc[C][COLS];
#pragma unroll
        for (int v = 0; v < V; v++) {
          const float4 e = rowtab[v * g.BH + yl];
          const int fx = __float_as_int(e.x), fy = __float_as_int(e.y);
          if ((fx | fy) < 0) {  // the source-row pair moved
            if (fy < 0) {
              hrow(v, __float_as_int(e.z), Ha[v]);
            } else {
#pragma unroll
              for (int c = 0; c < C; c++)
#pragma unroll
                for (int col = 0; col < COLS; col++) Ha[v][c][col] = Hb[v][c][col];
            }
            hrow(v, __float_as_int(e.w), Hb[v]);
          }
          const float l0 = fabsf(e.x), l1 = fabsf(e.y);
          float u[C][COLS];
#pragma unroll
          for (int c = 0; c < C; c++)
#pragma unroll
            for (int col = 0; col < COLS; col++) u[c][col] = __fmaf_rn(l0, Ha[v][c][col], __fmul_rn(l1, Hb[v][c][col]));
          if (p.fuse_mode == PISTO_FUSE_PROB_MEAN) {
#pragma unroll
            for (int col = 0; col < COLS; col++) {
              float t[C];
#pragma unroll
              for (int c = 0; c < C; c++) t[c] = u[c][col];
              pisto_softmax_inplace<C>(t);
#pragma unroll
              for (int c = 0; c < C; c++) u[c][col] = t[c];
            }
          }
#pragma unroll
          for (int c = 0; c < C; c++)
#pragma unroll
            for (int col = 0; col < COLS; col++) acc[c][col] = (v == 0) ? u[c][col] : __fadd_rn(acc[c][col], u[c][col]);
        }
        // ---- D. per-pixel epilogue ----------------------------------------------------------------------------
        unsigned int labs[COLS];
#pragma unroll
        for (int col = 0; col < COLS; col++) {
          float a[C];
#pragma unroll
          for (int c = 0; c < C; c++) a[c] = acc[c][col];
          int lab = (tp.single >= 0) ? tp.single : pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
          if (do_conf && gtv[col] < (unsigned)C) {
            unsigned int b = gtv[col] * C + lab;
            unsigned long long inc = 1ull << (8 * (b & 7));
            if (b < 8) cnt_lo += inc; else cnt_hi += inc;
          }
          labs[col] = (bgv[col] == (unsigned int)p.bg_match) ? (unsigned int)p.bg_label : (unsigned int)lab;
        }
        if (p.label_out) {
          if (COLS == 2) {
            *reinterpret_cast<uchar2*>(p.label_out + pix) = make_uchar2((unsigned char)labs[0], (unsigned char)labs[COLS - 1]);
          } else {
#pragma unroll
            for (int col = 0; col < COLS; col++) p.label_out[pix + col] = (unsigned char)labs[col];
          }
        }
        if (p.fused_out) {
#pragma unroll
          for (int c = 0; c < C; c++) {
            float* fo = p.fused_out + (((long long)n * C + c) * p.T_h + y) * p.T_w + x;
            if (COLS == 2) {
              *reinterpret_cast<float2*>(fo) = make_float2(pisto_div_views(acc[c][0], p.dec), pisto_div_views(acc[c][COLS - 1], p.dec));
            } else {
#pragma unroll
              for (int col = 0; col < COLS; col++) fo[col] = pisto_div_views(acc[c][col], p.dec);
            }
          }
        }
        if (need_low && (y % p.low_fh == p.low_fh / 2)) {
#pragma unroll
          for (int col = 0; col < COLS; col++) {
            if ((x + col) % p.low_fw == p.low_fw / 2) {
#pragma unroll
              for (int c = 0; c < C; c++)
                p.lowres_out[(((long long)n * C + c) * p.low_h + y / p.low_fh) * p.low_w + (x + col) / p.low_fw] =
                    pisto_div_views(acc[c][col], p.dec);
            }
          }
        }
      }
    } else if (col_ok && ys < ye) {
      // single-label tile (infer_pseudo_masks.py:71-73): constant label, background overwrite, no scores read
      const int x = x0 + xl0;
      for (int yl = ys; yl < ye; yl++) {
        const long long pix = ((long long)n * p.T_h + y0 + yl) * p.T_w + x;
#pragma unroll
        for (int col = 0; col < COLS; col++) {
          unsigned int bgv = p.bg ? (unsigned int)__ldg(p.bg + pix + col) : 0xffffffffu;
          if (do_conf) {
            unsigned int gv = __ldg(p.gt + pix + col);
            if (gv < (unsigned)C) {
              unsigned int b = gv * C + tp.single;
              unsigned long long inc = 1ull << (8 * (b & 7));
              if (b < 8) cnt_lo += inc; else cnt_hi += inc;
            }
          }
          if (p.label_out) p.label_out[pix + col] = (unsigned char)((bgv == (unsigned int)p.bg_match) ? p.bg_label : tp.single);
        }
      }
    }
    if (!need_scores && need_low) {
      // single-label tile whose 32x32 logits are still exported (infer_pseudo_masks.py:126 precedes the shortcut):
      // evaluate the fused scores only at the gather points of this block
      const int ly0 = (y0 + p.low_fh - 1 - p.low_fh / 2) / p.low_fh;  // first low row with centre >= y0
      const int lx0 = (x0 + p.low_fw - 1 - p.low_fw / 2) / p.low_fw;
      const int ytop = y0 + bh - 1 - p.low_fh / 2, xtop = x0 + bw - 1 - p.low_fw / 2;
      const int ly1 = ytop < 0 ? 0 : min(p.low_h, ytop / p.low_fh + 1);
      const int lx1 = xtop < 0 ? 0 : min(p.low_w, xtop / p.low_fw + 1);
      const int nly = max(ly1 - ly0, 0), nlx = max(lx1 - lx0, 0);
      for (int i = tid; i < nly * nlx; i += nt) {
        const int ly = ly0 + i / nlx, lx = lx0 + i % nlx;
        const int yl = ly * p.low_fh + p.low_fh / 2 - y0, xl = lx * p.low_fw + p.low_fw / 2 - x0;
        float a[C];
#pragma unroll
        for (int v = 0; v < V; v++) {
          const float4 er = rowtab[v * g.BH + yl];
          const float4 ec = coltab[v * g.BW + xl];
          const float* base = vsm + g.view_off[v];
          const int plane = sr[v].plane;
          const int r0 = __float_as_int(er.z), r1 = __float_as_int(er.w), oa = __float_as_int(ec.x), ob = __float_as_int(ec.y);
          const float l0 = fabsf(er.x), l1 = fabsf(er.y);
          float u[C];
#pragma unroll
          for (int c = 0; c < C; c++) {
            const float* pl = base + c * plane;
            float h0 = __fmaf_rn(ec.z, pl[r0 + oa], __fmul_rn(ec.w, pl[r0 + ob]));
            float h1 = __fmaf_rn(ec.z, pl[r1 + oa], __fmul_rn(ec.w, pl[r1 + ob]));
            u[c] = __fmaf_rn(l0, h0, __fmul_rn(l1, h1));
          }
          if (p.fuse_mode == PISTO_FUSE_PROB_MEAN) pisto_softmax_inplace<C>(u);
#pragma unroll
          for (int c = 0; c < C; c++) a[c] = (v == 0) ? u[c] : __fadd_rn(a[c], u[c]);
        }
#pragma unroll
        for (int c = 0; c < C; c++)
          p.lowres_out[(((long long)n * C + c) * p.low_h + ly) * p.low_w + lx] = pisto_div_views(a[c], p.dec);
      }
    }
    if (do_conf) {
      // every lane of every warp reaches this point: full-mask warp reductions are safe
#pragma unroll
      for (int b = 0; b < BINS; b++) {
        unsigned int v = (unsigned int)(((b < 8 ? cnt_lo : cnt_hi) >> (8 * (b & 7))) & 0xffull);
        v = __reduce_add_sync(0xffffffffu, v);
        if ((tid & 31) == 0 && v) atomicAdd(&hist[b], v);
      }
    }
  }
  if (do_conf) {
    __syncthreads();
    for (int i = tid; i < BINS; i += nt)
      if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: geometry
// ---------------------------------------------------------------------------------------------------------------
static int view_block_floats(const ViewDev& vw, int C, int y0, int bh, int x0, int bw) {
  const ViewMap& m = vw.map;
  int r_lo = pisto_src_index(vw.scale_h, y0, m.ho, vw.same_h).i0;
  int r_hi = pisto_src_index(vw.scale_h, y0 + bh - 1, m.ho, vw.same_h).i1;
  int c_lo = pisto_src_index(vw.scale_w, x0, m.wo, vw.same_w).i0;
  int c_hi = pisto_src_index(vw.scale_w, x0 + bw - 1, m.wo, vw.same_w).i1;
  int nr = r_hi - r_lo + 1, nc = c_hi - c_lo + 1;
  // raw rows come from i when the map is not transposed (ai != 0), else from j
  int nrows = m.ai != 0 ? nr : nc, pitch = m.ai != 0 ? nc : nr;
  return C * nrows * pitch;
}

static bool make_geom(const pisto_ctx* h, const FuseParams& p, int COLS, StreamGeom* g) {
  const int budget = h->smem_optin - 1024;
  int BW = p.T_w;
  const int max_bw = kMaxThreads * COLS;
  if (BW > max_bw) {
    int nb = (p.T_w + max_bw - 1) / max_bw;
    BW = ((p.T_w + nb - 1) / nb + COLS - 1) / COLS * COLS;
  }
  if (BW % COLS) return false;
  const int GX = BW / COLS;
  int S = kMaxThreads / GX;
  if (S < 1) return false;
  // packed 8-bit confusion counters: rows_per_strip * COLS <= 255
  for (int BH = p.T_h; BH >= 8; BH = (BH + 1) / 2) {
    int s_eff = S;
    if (s_eff > BH) s_eff = BH;
    int rps = (BH + s_eff - 1) / s_eff;
    if (rps * COLS > 255) continue;
    StreamGeom t;
    t.BW = BW; t.BH = BH;
    t.nbx = (p.T_w + BW - 1) / BW; t.nby = (p.T_h + BH - 1) / BH;
    t.GX = GX; t.S = s_eff; t.rows_per_strip = rps;
    int off = 0;
    t.sr_off = off; off += (int)sizeof(SubRect) * PISTO_MAX_VIEWS;
    t.rowtab_off = off; off += 16 * p.V * BH;
    t.coltab_off = off; off += 16 * p.V * BW;
    t.hist_off = off; off += 4 * 64;
    t.views_off = off;
    int fl = 0;
    for (int v = 0; v < p.V; v++) {
      int cap = 0;
      for (int by = 0; by < t.nby; by++)
        for (int bx = 0; bx < t.nbx; bx++) {
          int y0 = by * BH, x0 = bx * BW;
          int bh = p.T_h - y0 < BH ? p.T_h - y0 : BH, bw = p.T_w - x0 < BW ? p.T_w - x0 : BW;
          int f = view_block_floats(p.view[v], p.C, y0, bh, x0, bw);
          if (f > cap) cap = f;
        }
      cap = (cap + 3) & ~3;
      t.view_off[v] = fl; t.view_cap[v] = cap; fl += cap;
    }
    off += 4 * fl;
    t.smem_bytes = off;
    t.n_items = (long long)p.N * t.nbx * t.nby;
    if (off <= budget) { *g = t; return true; }
  }
  return false;
}

template <int C, int V, int COLS>
int launch_cv(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  StreamGeom g;
  if (!make_geom(h, p, COLS, &g)) return PISTO_OK;  // not launched: caller falls back
  auto kern = fuse_block_kernel<C, V, COLS>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int threads = (g.GX * g.S + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  long long grid = g.n_items < h->sm_count ? g.n_items : h->sm_count;
  // two CTAs per SM when they fit, so that one CTA's staging overlaps the other's arithmetic
  if (2 * g.smem_bytes + 2048 <= h->smem_optin && 2 * threads <= 1024 && g.n_items >= 2LL * h->sm_count) grid = 2LL * h->sm_count;
  kern<<<(int)grid, threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}

}  // namespace

int pisto_launch_fuse_block(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  *launched = false;
  if (p.entropy_out) return PISTO_OK;                    // dead output in the reference: generic kernel only
  if (p.T_w % 2 != 0) return PISTO_OK;
  if (p.conf && p.gt && p.C > 4) return PISTO_OK;        // packed counters hold C*C <= 16 bins
  // 2-byte / 8-byte vector accesses need even addresses
  if (((uintptr_t)p.label_out | (uintptr_t)p.bg | (uintptr_t)p.gt) & 1) return PISTO_OK;
  if ((uintptr_t)p.fused_out & 7) return PISTO_OK;
#define PISTO_CASE(CC, VV, COLS) \
  if (p.C == CC && p.V == VV) return launch_cv<CC, VV, COLS>(h, p, st, launched);
  PISTO_CASE(3, 1, 2)
  PISTO_CASE(3, 2, 2)
  PISTO_CASE(3, 6, 2)
  PISTO_CASE(3, 8, 2)
  PISTO_CASE(4, 1, 2)
  PISTO_CASE(4, 2, 2)
  PISTO_CASE(4, 6, 2)
  PISTO_CASE(4, 10, 1)
#undef PISTO_CASE
  return PISTO_OK;
}
