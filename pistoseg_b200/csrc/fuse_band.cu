// Block-tiled filtered fusion kernel for tiles too large to stage whole (BASELINE config 5: 512 .. 2048 px tiles,
// 5 scales x flip, 4 classes).  Same method as fuse_filter.cuh -- class-DIFFERENCE fields interpolated once per scale
// group, labels from their signs, margin test against a rounding-error bound, exact re-evaluation of the pixels that fail
// it -- applied to output blocks of BH x BW pixels:
//
//   work item   (tile n, row band by, column block bx); persistent CTAs take items grid-stride (all items cost the same).
//   tables      row / column lerp tables of the block, relative to the block's sub-rectangle of every low-resolution map.
//   pre-pass    reads the raw views straight from global memory (each block needs ~BH*h/T + 2 rows x BW*w/T + 2 columns
//               of every view; neighbouring blocks re-read one or two rows / columns through L2) and writes the summed
//               difference maps Y[g][k] of the sub-rectangle to shared memory; max |x| on the way.
//   row loop    filter_rows of fuse_filter.cuh on the block (labels to a BH x BW shared-memory tile, uncertain pixels queued).
//   exact pass  queued pixels: the reference's operation order from global memory (pisto_decide).
//   vector pass 16-byte label rows + gt / bg words: confusion, background overwrite, stores.
//
// HBM traffic is the compulsory minimum plus the halo re-reads of the low-resolution views (a few percent).
#include "fuse_filter.cuh"

#ifndef PISTO_BAND_PRE_UNROLL
#define PISTO_BAND_PRE_UNROLL 1  // cells of the pre-pass in flight per thread (8 independent global loads each); 1 measures 3 % faster than 2, 4 slower
#endif

namespace {
constexpr int kBandPreUnroll = PISTO_BAND_PRE_UNROLL;

// Block shape (A/B knobs): threads per CTA, CTAs per SM, block width, rows per strip.  Two 256-thread CTAs per SM on 72 x 256
// blocks let one CTA's pre-pass (global loads, latency-bound) and barriers overlap the other's row loop; one 448-thread CTA on
// 72 x 512 blocks was the first version.
#ifndef PISTO_BAND_THREADS
#define PISTO_BAND_THREADS 256
#endif
#ifndef PISTO_BAND_CTAS
#define PISTO_BAND_CTAS 2
#endif
#ifndef PISTO_BAND_BW
#define PISTO_BAND_BW 256
#endif
#ifndef PISTO_BAND_RPS
#define PISTO_BAND_RPS 18
#endif
constexpr int kBThreads = PISTO_BAND_THREADS;
constexpr int kBCtas = PISTO_BAND_CTAS;

struct BandGeom {
  int BH, BW, nby, nbx, S, GX;
  int strip_y0[9];
  int nrows_cap[kFMaxGroups], ncols_cap[kFMaxGroups];
  int rowtab_off, rowoff_off, cola_off, colb_off, ymap_off, queue_off, lab_off, ctl_off, smem_bytes;
};

// exact fused sums of one pixel in the reference's operation order, raw views read from global memory
template <int C, int V>
__device__ __forceinline__ void exact_pixel_global(const FuseParams& p, const FilterGeom& g, int n, int yy, int xx, float (&a)[C]) {
#pragma unroll
  for (int v = 0; v < V; v++) {
    const ViewDev& vw = p.view[v];
    const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, vw.same_h);
    const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, vw.same_w);
    const int r0 = g.vbase[v] + Ly.i0 * g.vrow[v], r1 = g.vbase[v] + Ly.i1 * g.vrow[v];
    const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
    const float* src = vw.logits + (long long)n * vw.tile_stride;
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int pl = c * g.plane_bytes[v];
      const float h0 = __fmaf_rn(Lx.l0, __ldg(src + ((r0 + pl + c0) >> 2)), __fmul_rn(Lx.l1, __ldg(src + ((r0 + pl + c1) >> 2))));
      const float h1 = __fmaf_rn(Lx.l0, __ldg(src + ((r1 + pl + c0) >> 2)), __fmul_rn(Lx.l1, __ldg(src + ((r1 + pl + c1) >> 2))));
      const float u = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
      a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
    }
  }
}

// F: 1 bg, 2 gt/conf, 16 labels (compile-time feature mask as in fuse_filter.cuh; < 0: run-time)
template <int C, int V, int G, int F>
__global__ void __launch_bounds__(kBThreads, kBCtas) fuse_band_kernel(const __grid_constant__ FuseParams p, const __grid_constant__ FilterGeom g,
                                                                 const __grid_constant__ BandGeom bg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FCtl* ctl = reinterpret_cast<FCtl*>(smem_raw + bg.ctl_off);
  int2* rowoff = reinterpret_cast<int2*>(smem_raw + bg.rowoff_off);
  float2* col2 = reinterpret_cast<float2*>(smem_raw + bg.cola_off);     // [G][BW/2] {l1 of the pair's two columns}
  uint32_t* col2i = reinterpret_cast<uint32_t*>(smem_raw + bg.colb_off); // [G][BW/2] (4 * first source column, relative to the block) | sel << 16
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + bg.queue_off);
  uint8_t* labsm = smem_raw + bg.lab_off;
  constexpr bool RT = F < 0;
  constexpr int RS = 16 * ((G + 2) / 2);
  constexpr int BINS = C * C;
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int T_h = p.T_h, T_w = p.T_w, BH = bg.BH, BW = bg.BW;
  const uint32_t rowtab_s = smem_u32(smem_raw + bg.rowtab_off), rowoff_s = smem_u32(rowoff), ymap_s = smem_u32(smem_raw + bg.ymap_off);
  const uint32_t lab_s = smem_u32(labsm);

  if (tid == 0) { ctl->maxbits[0] = 0u; ctl->qcount[0] = 0u; }
  for (int i = tid; i < 64; i += nt) ctl->hist[i] = 0;
  __syncthreads();

  const int grp = tid % bg.GX, strip = min(tid / bg.GX, bg.S - 1);
  const bool worker = tid < bg.GX * bg.S;
  const int xr = 4 * grp;  // first of the thread's 4 columns, relative to the block
  const uint32_t colA_t = smem_u32(col2) + 16u * grp, colB_t = smem_u32(col2i) + 8u * grp;  // the thread's first pair
  u64 cnt_lo = 0, cnt_hi = 0;

  const long long items = (long long)p.N * bg.nby * bg.nbx;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = (int)(item / (bg.nby * bg.nbx));
    const int rem = (int)(item - (long long)n * bg.nby * bg.nbx);
    const int by = rem / bg.nbx, bx = rem - by * bg.nbx;
    const int y0 = by * BH, y1 = min(T_h, y0 + BH), x0 = bx * BW;
    const int rows = y1 - y0;
    const TilePresence tp = pisto_tile_presence(p, n);
    const bool multi = tp.single < 0;
    int cls[C], P = 0;
#pragma unroll
    for (int c = 0; c < C; c++) cls[c] = 0;
#pragma unroll
    for (int c = 0; c < C; c++)
      if ((tp.bits >> c) & 1u) {
#pragma unroll
        for (int q = 0; q < C; q++)
          if (q == P) cls[q] = c;
        P++;
      }
    // origin of the block's sub-rectangle in every low-resolution map
    int ib[G], jb[G];
#pragma unroll
    for (int gi = 0; gi < G; gi++) {
      ib[gi] = pisto_src_index(g.g_scale_h[gi], y0, g.g_ho[gi], false).i0;
      jb[gi] = pisto_src_index(g.g_scale_w[gi], x0, g.g_wo[gi], g.g_same_w[gi]).i0;
    }
    bool exact_all = false;
    float tau = 0.f;
    if (multi && P >= 2) {
      // ---- tables of this block ----------------------------------------------------------------------------------------
      for (int r = tid; r < rows; r += nt) {
        unsigned char* row = smem_raw + bg.rowtab_off + RS * r;
        bool strip_start = false;
        for (int q = 0; q < bg.S; q++) strip_start |= (r == bg.strip_y0[q]);
        unsigned int f = 0;
#pragma unroll
        for (int gi = 0; gi < G; gi++) {
          const Lerp L = pisto_src_index(g.g_scale_h[gi], y0 + r, g.g_ho[gi], false);
          reinterpret_cast<float2*>(row)[gi] = make_float2(-L.l0, -L.l0);
          const int stride = 4 * (bg.ncols_cap[gi] + 2);  // two pad columns per map row (fuse_filter.cuh, 3-tap loads)
          rowoff[r * G + gi] = make_int2((L.i0 - ib[gi]) * stride, (L.i1 - ib[gi]) * stride);
          if (!strip_start) {
            const Lerp Q = pisto_src_index(g.g_scale_h[gi], y0 + r - 1, g.g_ho[gi], false);
            if (Q.i0 != L.i0 || Q.i1 != L.i1) f |= 1u << gi;
          }
        }
        reinterpret_cast<uint2*>(row)[G] = make_uint2(f, 0u);
      }
      for (int i = tid; i < G * (BW / 2); i += nt) {
        const int gi = i / (BW / 2), gx = i - gi * (BW / 2);
        const Lerp L0 = pisto_src_index(g.g_scale_w[gi], x0 + 2 * gx, g.g_wo[gi], g.g_same_w[gi]);
        const Lerp L1 = pisto_src_index(g.g_scale_w[gi], x0 + 2 * gx + 1, g.g_wo[gi], g.g_same_w[gi]);
        col2[i] = make_float2(L0.l1, L1.l1);                                              // l0 = 1 - l1 (pisto_src_index)
        col2i[i] = (unsigned)(4 * (L0.i0 - jb[gi])) | ((unsigned)((L1.i0 - L0.i0) << 1) << 16);  // up-sampling: the second column starts 0 or 1 cells later
      }
      // ---- pre-pass: summed class differences of the sub-rectangle, from global memory ---------------------------------
      float mxf = 0.f;
      const uint32_t kp = P == 4 ? 4u : (uint32_t)(P - 1);  // floats per cell of the interleaved maps (K = 3 is padded to 4)
#pragma unroll
      for (int gi = 0; gi < G; gi++) {
        const int ie = pisto_src_index(g.g_scale_h[gi], y1 - 1, g.g_ho[gi], false).i1;
        const int je = pisto_src_index(g.g_scale_w[gi], x0 + BW - 1, g.g_wo[gi], g.g_same_w[gi]).i1;
        const int nr = ie - ib[gi] + 1, nc = je - jb[gi] + 1, cells = nr * nc, cap = bg.ncols_cap[gi];
        const unsigned int inv = (unsigned int)((0x100000000ull + nc - 1) / nc);  // idx / nc == umulhi(idx, inv) for idx < 2^16
        const uint32_t ym = ymap_s + g.g_ybytes[gi];
        if (V == 2 * G) {
          // views (2g, 2g+1) form group g (a scale and its flipped twin): both are read in one sweep, all loads of a cell
          // (2 views x P classes) are independent and two cells are in flight per thread
          const int va = 2 * gi < V ? 2 * gi : 0, vc = 2 * gi + 1 < V ? 2 * gi + 1 : 0;
          const float* sa = p.view[va].logits + (long long)n * p.view[va].tile_stride;
          const float* sc = p.view[vc].logits + (long long)n * p.view[vc].tile_stride;
          const int ra = g.vrow[va], ca = g.vcol[va], rc = g.vrow[vc], cc = g.vcol[vc];
          const int basea = g.vbase[va] + cls[0] * g.plane_bytes[va] + ib[gi] * ra + jb[gi] * ca;
          const int basec = g.vbase[vc] + cls[0] * g.plane_bytes[vc] + ib[gi] * rc + jb[gi] * cc;
          int dqa[C - 1], dqc[C - 1];
#pragma unroll
          for (int q = 0; q < C - 1; q++) { dqa[q] = (cls[q + 1] - cls[0]) * g.plane_bytes[va]; dqc[q] = (cls[q + 1] - cls[0]) * g.plane_bytes[vc]; }
#pragma unroll kBandPreUnroll
          for (int idx = tid; idx < cells; idx += nt) {
            const int di = (int)__umulhi((unsigned)idx, inv), dj = idx - di * nc;
            const int oa = basea + di * ra + dj * ca, oc = basec + di * rc + dj * cc;
            const float x0a = __ldg(sa + (oa >> 2)), x0c = __ldg(sc + (oc >> 2));
            float xa[C - 1], xc[C - 1];
#pragma unroll
            for (int q = 0; q < C - 1; q++)
              if (q + 1 < P) { xa[q] = __ldg(sa + ((oa + dqa[q]) >> 2)); xc[q] = __ldg(sc + ((oc + dqc[q]) >> 2)); }
            mxf = max_nan(max_nan(mxf, fabsf(x0a)), fabsf(x0c));
#pragma unroll
            for (int q = 0; q < C - 1; q++)
              if (q + 1 < P) {
                mxf = max_nan(max_nan(mxf, fabsf(xa[q])), fabsf(xc[q]));
                const float yv = __fadd_rn(__fsub_rn(xa[q], x0a), __fsub_rn(xc[q], x0c));
                const uint32_t ya = ym + 4u * (kp * (di * (cap + 2) + dj) + q);   // [row][column][k], two pad columns per row
                sts_f32(ya, yv);
                if (jb[gi] + dj == g.g_wo[gi] - 1) { sts_f32(ya + 4u * kp, yv); sts_f32(ya + 8u * kp, yv); }  // right-edge clamp
              }
          }
        } else {
          for (int idx = tid; idx < cells; idx += nt) {
            const int di = (int)__umulhi((unsigned)idx, inv), dj = idx - di * nc;
            const int i = ib[gi] + di, j = jb[gi] + dj;
            float y[C - 1];
            bool first = true;
#pragma unroll
            for (int v = 0; v < V; v++) {
              if (g.group_of[v] != gi) continue;
              const float* src = p.view[v].logits + (long long)n * p.view[v].tile_stride;
              const int a = g.vbase[v] + i * g.vrow[v] + j * g.vcol[v];
              const float x0v = __ldg(src + ((a + cls[0] * g.plane_bytes[v]) >> 2));
              mxf = max_nan(mxf, fabsf(x0v));
#pragma unroll
              for (int q = 0; q < C - 1; q++) {
                if (q + 1 < P) {
                  const float xq = __ldg(src + ((a + cls[q + 1] * g.plane_bytes[v]) >> 2));
                  mxf = max_nan(mxf, fabsf(xq));
                  const float t = __fsub_rn(xq, x0v);
                  y[q] = first ? t : __fadd_rn(y[q], t);
                }
              }
              first = false;
            }
#pragma unroll
            for (int q = 0; q < C - 1; q++)
              if (q + 1 < P) {
                const uint32_t ya = ym + 4u * (kp * (di * (cap + 2) + dj) + q);
                sts_f32(ya, y[q]);
                if (j == g.g_wo[gi] - 1) { sts_f32(ya + 4u * kp, y[q]); sts_f32(ya + 8u * kp, y[q]); }
              }
          }
        }
      }
      const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));
      if ((tid & 31) == 0) atomicMax(&ctl->maxbits[0], mx);
    }
    __syncthreads();
    if (multi) {
      if (P >= 2) {
        const float A = __fmul_rn((float)V, __uint_as_float(ctl->maxbits[0]));
        tau = __fmaf_rn(A, g.tau_coef, g.tau_abs);
        if (!(A < 5e8f)) exact_all = true;
      } else {
        exact_all = true;
      }
      const int ys = bg.strip_y0[strip], ye = min(bg.strip_y0[strip + 1], rows);
      if (!exact_all && worker && ys < ye) {
        // two passes of two columns: the register file does not hold four columns of G*K fields for G = 5
        for (int half = 0; half < 2; half++) {
          const uint32_t ca = colA_t + 8u * half, cb = colB_t + 4u * half;
          const int xx = xr + 2 * half;
          if (P == 2) filter_rows<C, G, 16, 1, 1, true, 2>(p, g, ctl, queue, 0, rowtab_s, rowoff_s, ca, cb, ymap_s, lab_s, n, xx, ys, ye, cls, tau, cnt_lo, cnt_hi);
          else if (P == 3) filter_rows<C, G, 16, 1, 2, true, 2>(p, g, ctl, queue, 0, rowtab_s, rowoff_s, ca, cb, ymap_s, lab_s, n, xx, ys, ye, cls, tau, cnt_lo, cnt_hi);
          else filter_rows<C, G, 16, 1, (C >= 4 ? 3 : 1), true, 2>(p, g, ctl, queue, 0, rowtab_s, rowoff_s, ca, cb, ymap_s, lab_s, n, xx, ys, ye, cls, tau, cnt_lo, cnt_hi);
        }
      }
      __syncthreads();
      // ---- exact pass ---------------------------------------------------------------------------------------------------
      const unsigned int nq = ctl->qcount[0];
      if (nq > (unsigned)kFQueueCap) exact_all = true;
      const int nfix = exact_all ? rows * BW : (int)nq;
      for (int j = tid; j < nfix; j += nt) {
        int ry, rx;
        if (exact_all) { ry = j / BW; rx = j - ry * BW; }
        else { const uint32_t e = queue[j]; ry = (int)(e >> 16); rx = (int)(e & 0xffffu); }
        float a[C];
        exact_pixel_global<C, V>(p, g, n, y0 + ry, x0 + rx, a);
        labsm[ry * BW + rx] = (uint8_t)pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
      }
      __syncthreads();
    }
    if (tid == 0) { ctl->maxbits[0] = 0u; ctl->qcount[0] = 0u; }
    // ---- vector pass: confusion, background overwrite, 16-byte stores ------------------------------------------------------
    // Two 16-byte vectors (32 pixels) per step; the confusion counts come from bit planes as in confusion.cu (labels and
    // ground truth are transposed with one shift + one LOP3 per word and plane, then popc(G_a & P_b) per bin), so the cost does
    // not depend on how ragged the label map is.
    {
      const unsigned int labc = 0x01010101u * (unsigned)(multi ? 0 : tp.single), bgl4 = 0x01010101u * (unsigned)p.bg_label;
      const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match;
      const int vpr = BW / 32, ngrp = rows * vpr;  // 32-pixel groups per block row
      unsigned int cnt[BINS];
#pragma unroll
      for (int i = 0; i < BINS; i++) cnt[i] = 0;
      for (int i = tid; i < ngrp; i += nt) {
        const int ry = i / vpr, vx = i - ry * vpr;
        const long long pix = ((long long)n * T_h + y0 + ry) * T_w + x0 + 32 * vx;
        unsigned int lw[8];
        if (multi) {
          const int4 t0 = lds_i4(lab_s + ry * BW + 32 * vx), t1 = lds_i4(lab_s + ry * BW + 32 * vx + 16);
          lw[0] = t0.x; lw[1] = t0.y; lw[2] = t0.z; lw[3] = t0.w; lw[4] = t1.x; lw[5] = t1.y; lw[6] = t1.z; lw[7] = t1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; q++) lw[q] = labc;
        }
        if (do_conf) {
          const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(p.gt + pix)), g1 = __ldg(reinterpret_cast<const uint4*>(p.gt + pix + 16));
          const unsigned int gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          bitslice_count<C>(gw, lw, cnt);
        }
        if (has_label) {
          unsigned int ow[8];
#pragma unroll
          for (int q = 0; q < 8; q++) ow[q] = lw[q];
          if (has_bg) {
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(p.bg + pix)), b1 = __ldg(reinterpret_cast<const uint4*>(p.bg + pix + 16));
            const unsigned int bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int q = 0; q < 8; q++) { const unsigned int eq = __vcmpeq4(bw[q], m4); ow[q] = (bgl4 & eq) | (lw[q] & ~eq); }
          }
          reinterpret_cast<uint4*>(p.label_out + pix)[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          reinterpret_cast<uint4*>(p.label_out + pix)[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
      }
      if (do_conf) {
#pragma unroll
        for (int bn = 0; bn < BINS; bn++) {
          const unsigned int cv = __reduce_add_sync(0xffffffffu, cnt[bn]);
          if ((tid & 31) == 0 && cv) atomicAdd(&ctl->hist[bn], cv);
        }
      }
    }
    __syncthreads();  // tables / label tile / queue are reused by the next item
  }
  if (do_conf) {
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

template <int C, int V, int G, int F>
int launch_band(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  FilterGeom g;
  BandGeom b;
  memset(&g, 0, sizeof(g));
  memset(&b, 0, sizeof(b));
  // scale groups (same rules as make_filter_geom)
  int Gn = 0, nmax = 0, cnt[kFMaxGroups] = {0};
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    if (vw.map.ho >= p.T_h || vw.map.wo >= p.T_w) return PISTO_OK;  // up-sampling in both directions (3-tap pair loads)
    int gi = -1;
    for (int q = 0; q < Gn; q++)
      if (g.g_ho[q] == vw.map.ho && g.g_wo[q] == vw.map.wo) gi = q;
    if (gi < 0) {
      if (Gn == kFMaxGroups) return PISTO_OK;
      gi = Gn++;
      g.g_ho[gi] = vw.map.ho; g.g_wo[gi] = vw.map.wo; g.g_same_w[gi] = vw.same_w;
      g.g_scale_h[gi] = vw.scale_h; g.g_scale_w[gi] = vw.scale_w;
    }
    g.group_of[v] = gi;
    g.first_in_group[v] = cnt[gi] == 0;
    if (++cnt[gi] > nmax) nmax = cnt[gi];
    g.plane_bytes[v] = 4 * vw.h * vw.w;
    g.vbase[v] = 4 * (vw.map.a0 * vw.w + vw.map.b0);
    g.vrow[v] = 4 * (vw.map.ai * vw.w + vw.map.bi);
    g.vcol[v] = 4 * (vw.map.aj * vw.w + vw.map.bj);
  }
  if (Gn != G) return PISTO_OK;
  // block shape: 512 columns (4 per thread, 128 threads per row), 3 strips of 24 rows
  b.BW = (p.T_w < PISTO_BAND_BW ? p.T_w : PISTO_BAND_BW) & ~31;   // the widest multiple of 32 up to the knob that divides the tile width
  while (b.BW >= 32 && p.T_w % b.BW) b.BW -= 32;
  if (b.BW < 32) return PISTO_OK;
  if (p.T_w % b.BW || b.BW % 32) return PISTO_OK;
  b.GX = b.BW / 4;
  b.S = (kBThreads / b.GX) < 8 ? (kBThreads / b.GX) : 8;
  if (b.S < 1) return PISTO_OK;
  const int rps = PISTO_BAND_RPS;
  b.BH = b.S * rps;
  if (b.BH > p.T_h) { b.BH = p.T_h; }
  for (int q = 0; q <= b.S; q++) b.strip_y0[q] = q * rps < b.BH ? q * rps : b.BH;
  b.nby = (p.T_h + b.BH - 1) / b.BH;
  b.nbx = p.T_w / b.BW;
  // capacity of every group's sub-rectangle
  for (int gi = 0; gi < G; gi++) {
    int nr = 0, nc = 0;
    for (int by = 0; by < b.nby; by++) {
      const int y0 = by * b.BH, y1 = (y0 + b.BH < p.T_h ? y0 + b.BH : p.T_h) - 1;
      const int r = pisto_src_index(g.g_scale_h[gi], y1, g.g_ho[gi], false).i1 - pisto_src_index(g.g_scale_h[gi], y0, g.g_ho[gi], false).i0 + 1;
      if (r > nr) nr = r;
    }
    for (int bx = 0; bx < b.nbx; bx++) {
      const int x0 = bx * b.BW, x1 = x0 + b.BW - 1;
      const int c = pisto_src_index(g.g_scale_w[gi], x1, g.g_wo[gi], g.g_same_w[gi]).i1 - pisto_src_index(g.g_scale_w[gi], x0, g.g_wo[gi], g.g_same_w[gi]).i0 + 1;
      if (c > nc) nc = c;
    }
    b.nrows_cap[gi] = nr; b.ncols_cap[gi] = nc;
  }
  const float cE = 2.f * nmax + 4.f * G + 2.f * p.V + 20.f;
  g.tau_coef = 2.f * cE * 5.9604645e-8f + 2.5e-7f;
  g.tau_abs = p.dec.margin_abs * 1.01f;
  g.GXP = b.BW / 2;
  g.lab_stride = b.BW;
  const int RS = 16 * ((G + 2) / 2);
  int off = 0;
  b.ctl_off = off; off += (int)((sizeof(FCtl) + 127) & ~127u);
  b.rowtab_off = off; off += RS * b.BH + 16;
  b.rowoff_off = off; off += 8 * G * b.BH; off = (off + 15) & ~15;
  b.cola_off = off; off += 8 * G * (b.BW / 2);
  b.colb_off = off; off += 4 * G * (b.BW / 2); off = (off + 15) & ~15;
  b.ymap_off = off;
  for (int gi = 0; gi < G; gi++) {
    g.g_ybytes[gi] = off - b.ymap_off;
    g.g_mapbytes[gi] = 4 * b.nrows_cap[gi] * (b.ncols_cap[gi] + 2);
    off += (p.C == 4 ? 4 : p.C - 1) * g.g_mapbytes[gi];
    off = (off + 15) & ~15;
  }
  b.queue_off = off; off += 4 * kFQueueCap;
  b.lab_off = off; off += b.BH * b.BW;
  b.smem_bytes = off;
  if (off > (kBCtas > 1 ? (233472 / kBCtas - 1024) : h->smem_optin) - 1024) return PISTO_OK;
  auto kern = fuse_band_kernel<C, V, G, F>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b.smem_bytes));
  const long long items = (long long)p.N * b.nby * b.nbx;
  const long long slots = (long long)h->sm_count * kBCtas;
  const int grid = items < slots ? (int)items : (int)slots;
  const int threads = ((b.GX * b.S + 31) / 32) * 32;
  kern<<<grid, threads, b.smem_bytes, st>>>(p, g, b);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}

template <int C, int V, int G>
int launch_band_f(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  switch (pisto_filter_flags(p)) {
    case 18: return launch_band<C, V, G, 18>(h, p, st, launched);  // gt/conf + labels (config 5)
    default: return launch_band<C, V, G, -1>(h, p, st, launched);
  }
}

}  // namespace

// large tiles, labels / confusion only: the difference-field filter on output blocks
int pisto_launch_fuse_band(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  *launched = false;
  if (p.fuse_mode != PISTO_FUSE_LOGIT_MEAN) return PISTO_OK;
  if (p.fused_out || p.entropy_out || p.lowres_out) return PISTO_OK;
  if (p.dec.mask_mode == PISTO_MASK_MULTIPLY) return PISTO_OK;
  if (!p.label_out && !(p.conf && p.gt)) return PISTO_OK;
  if (p.conf && p.gt && p.C > 4) return PISTO_OK;
  if (p.T_h > 65535 || p.T_w > 65535 || p.T_w % 32) return PISTO_OK;
  if (((uintptr_t)p.label_out | (uintptr_t)p.bg | (uintptr_t)p.gt) & 15) return PISTO_OK;
  if (((long long)p.T_h * p.T_w) % 16) return PISTO_OK;
  for (int v = 0; v < p.V; v++)
    if ((uintptr_t)p.view[v].logits & 3) return PISTO_OK;
  const int G = pisto_filter_groups(p);
  if (p.C == 4 && p.V == 10 && G == 5) return launch_band_f<4, 10, 5>(h, p, st, launched);
  if (p.C == 3 && p.V == 10 && G == 5) return launch_band_f<3, 10, 5>(h, p, st, launched);
  if (p.C == 4 && p.V == 6 && G == 3) return launch_band_f<4, 6, 3>(h, p, st, launched);
  if (p.C == 3 && p.V == 6 && G == 3) return launch_band_f<3, 6, 3>(h, p, st, launched);
  return PISTO_OK;
}
