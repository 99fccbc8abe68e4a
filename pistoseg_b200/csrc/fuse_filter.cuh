// Filtered streaming fusion kernel -- the hot path for LOGIT_MEAN fusion when only labels / confusion / the 32x32 logit
// export are wanted (BASELINE configs 1, 2, 3: infer_pseudo_masks.py:118-154, loss.py:55-67).
//
// The reference evaluates, per pixel, V*C bilinear samples, sums them in view order and takes argmax(softmax(.)).  The
// label only depends on the ORDER of the fused class scores, so this kernel decides it from K = P-1 class-DIFFERENCE
// fields (P = classes present in the tile) that are interpolated once per scale group instead of once per view:
//
//   pre-pass    views that share their de-augmented size (a scale and its flipped twin) are summed at low resolution, in
//               the de-augmented frame, as differences against the lowest present class:
//                   Y[g][k][i][j] = sum_{v in g} ( x_v[c_k] - x_v[c_0] )(i, j)            (g < G groups, k < K)
//               and max|x| of the tile is reduced on the way.  Bilinear interpolation is linear, so
//                   D_k(y,x) = sum_g bilinear_g(Y[g][k])(y,x)  ==  a_{c_k}(y,x) - a_{c_0}(y,x)    up to rounding.
//   row loop    a thread owns 2*NP adjacent columns and streams down a strip of rows; per (group, k) it keeps the
//               horizontally interpolated lower source row Hb and the row difference Dh = Hb - Ha packed f32x2, so a
//               row costs ONE fma.f32x2 per (group, k, column pair):  D = base - l0_g * Dh_g,  base = sum_g Hb_g.
//               That is G*K packed operations per pixel pair instead of the 3*V*C of the exact evaluation
//               (cfg 2, two classes present: 3 instead of 54).
//   filter      the candidate order of (0, D_1 .. D_K) is trusted only when the best candidate leads the runner-up by
//               more than tau = A * (2 * cE * 2^-24 + 2.5e-7) + V * 2e-6 * 1.01,  A = V * max|x| (error analysis in
//               cE = 2 n_max + 4 G + 2 V + 16 bounds the rounding of both evaluations; the remaining
//               terms are the argmax(softmax(a / V)) == argmax(a) margin of common.cuh).  Then the exact fp32 sums a_c of
//               the reference are ordered the same way with that margin and the label is the reference's, bit for bit.
//   exact pass  pixels that fail the test (about 1e-4 of them on Gaussian logits) are queued in shared memory and, after
//               the strip loop, evaluated exactly -- operation by operation as torch does (pisto_decide) -- together with
//               the 32x32 gather points of the logit export, which are always exact (lowres_out is bit-exact).
//               Queue overflow, non-finite or absurdly large logits, and tiles whose presence vector is empty switch the
//               whole tile to the exact evaluation.
//
// Everything else (persistent CTAs, dynamic tile scheduler, producer warp + mbarrier full/empty pipeline with 1-D TMA
// staging of the raw views, packed confusion counters) is shared with fuse_stream.cuh.
#pragma once
#include "fuse_common.cuh"
#include "sm100_prims.cuh"

namespace {

constexpr int kFMaxThreads = 448;
constexpr int kFQueueCap = 1024;
constexpr int kFMaxGroups = 4;

struct FilterGeom {
  int GX, S, threads, cwarps;          // threads per output row, strips, CTA size (incl. the producer warp), compute warps
  int GXP;                             // column pairs per row (T_w / 2)
  int strip_y0[33];
  int view_off[PISTO_MAX_VIEWS];       // float offset of each view inside one staging buffer (16-byte aligned)
  int plane_bytes[PISTO_MAX_VIEWS];    // h*w*4
  int vbase[PISTO_MAX_VIEWS], vrow[PISTO_MAX_VIEWS], vcol[PISTO_MAX_VIEWS];  // byte address of de-augmented (i, j): vbase + i*vrow + j*vcol
  int group_of[PISTO_MAX_VIEWS];
  int buf_floats;
  int g_ho[kFMaxGroups], g_wo[kFMaxGroups], g_same_w[kFMaxGroups];
  float g_scale_h[kFMaxGroups], g_scale_w[kFMaxGroups];
  int g_ybytes[kFMaxGroups];           // byte offset of group g's first difference map inside the Y area
  int g_mapbytes[kFMaxGroups];         // ho*wo*4
  float tau_coef, tau_abs;             // tau = V * max|x| * tau_coef + tau_abs
  int ctl_off, rowtab_off, rowoff_off, cola_off, colb_off, lowrow_off, lowcol_off, ymap_off, queue_off, views_off;
  int smem_bytes;
  int* counter;
};

struct FCtl {
  uint64_t full[2];   // producer -> consumers: tile id published (+ TMA bytes landed)
  uint64_t empty[2];  // consumers -> producer: staging buffer may be refilled (one arrival per compute warp)
  int tile[2];
  unsigned int maxbits[2];  // max |x| of the tile (IEEE bits; NaN > Inf > finite), slot = tile parity
  unsigned int qcount[2];   // uncertain pixels queued by the row loop
  unsigned int hist[64];
};

// label among cls[0..K] from the difference candidates (0, d[0] .. d[K-1]); returns whether the lead exceeds tau
template <int K, int C>
__device__ __forceinline__ bool decide_diff(const float (&d)[K], const int (&cls)[C], float tau, int& lab) {
  float bv = fmaxf(0.f, d[0]), sv = fminf(0.f, d[0]);
  lab = d[0] > 0.f ? cls[1] : cls[0];
#pragma unroll
  for (int k = 1; k < K; k++) {
    const bool gt = d[k] > bv;
    sv = fmaxf(sv, fminf(bv, d[k]));
    lab = gt ? cls[k + 1] : lab;
    bv = fmaxf(bv, d[k]);
  }
  return __fsub_rn(bv, sv) > tau;
}

__device__ __noinline__ void push_uncertain(FCtl* ctl, uint32_t* queue, int b, int y, int x, unsigned int mask) {
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1;
    const unsigned int idx = atomicAdd(&ctl->qcount[b], 1u);
    if (idx < (unsigned)kFQueueCap) queue[idx] = ((unsigned)y << 16) | (unsigned)(x + j);
  }
}

template <int NP> __device__ __forceinline__ unsigned int ldg_px(const uint8_t* q) {
  if (NP == 2) return __ldg(reinterpret_cast<const unsigned int*>(q));
  return __ldg(reinterpret_cast<const unsigned short*>(q));
}
template <int NP> __device__ __forceinline__ void stg_px(uint8_t* q, unsigned int v) {
  if (NP == 2) *reinterpret_cast<unsigned int*>(q) = v;
  else *reinterpret_cast<unsigned short*>(q) = (unsigned short)v;
}

// ---- the row loop of one strip, K difference fields ----------------------------------------------------------------
template <int C, int G, int F, int NP, int K>
__device__ __forceinline__ void filter_rows(const FuseParams& p, const FilterGeom& g, FCtl* ctl, uint32_t* queue, int b,
                                            uint32_t rowtab_s, uint32_t rowoff_s, uint32_t colA_t, uint32_t colB_t, uint32_t ymap_s,
                                            int n, int x, int ys, int ye, const int (&cls)[C], float tau, u64& cnt_lo, u64& cnt_hi) {
  constexpr bool RT = F < 0;
  constexpr int RS = 16 * ((G + 2) / 2);
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  const int T_w = p.T_w;
  const uint32_t colg = 16u * g.GXP;

  u64 Hb[G][K][NP], Dh[G][K][NP], base[K][NP];
  // horizontally interpolated values of one row (byte offset `row` inside a map) of every difference map of group gi
  auto load_h = [&](int gi, uint32_t row, u64 (&H)[K][NP]) {
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const int4 A = lds_i4(colA_t + gi * colg + 16u * q);
      const ulonglong2 B = lds_u64x2(colB_t + gi * colg + 16u * q);
      uint32_t m = ymap_s + g.g_ybytes[gi] + row;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const float p00 = lds_f32(m + A.x), p01 = lds_f32(m + A.y), p10 = lds_f32(m + A.z), p11 = lds_f32(m + A.w);
        H[k][q] = fma2(B.x, pack2(p00, p10), mul2(B.y, pack2(p01, p11)));
        m += g.g_mapbytes[gi];
      }
    }
  };
  auto rebase = [&]() {
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < NP; q++) {
        u64 s = Hb[0][k][q];
#pragma unroll
        for (int gi = 1; gi < G; gi++) s = add2(s, Hb[gi][k][q]);
        base[k][q] = s;
      }
  };
#pragma unroll
  for (int gi = 0; gi < G; gi++) {
    const int2 ro = lds_i2(rowoff_s + 8u * (ys * G + gi));
    u64 Ha[K][NP];
    load_h(gi, ro.x, Ha);
    load_h(gi, ro.y, Hb[gi]);
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < NP; q++) Dh[gi][k][q] = sub2(Hb[gi][k][q], Ha[k][q]);
  }
  rebase();

  const long long pix0 = ((long long)n * p.T_h + ys) * T_w + x;
  const uint8_t* bgp = has_bg ? p.bg + pix0 : nullptr;
  const uint8_t* gtp = do_conf ? p.gt + pix0 : nullptr;
  uint8_t* lbp = has_label ? p.label_out + pix0 : nullptr;
  uint32_t rt = rowtab_s + RS * ys, ro_a = rowoff_s + 8u * G * ys;
  // byte masks are fetched two rows ahead of their use
  unsigned int bg_c = 0, bg_n = 0, gt_c = 0, gt_n = 0;
  if (has_bg) {
    bg_c = ldg_px<NP>(bgp);
    if (ys + 1 < ye) bg_n = ldg_px<NP>(bgp + T_w);
    bgp += 2 * T_w;
  }
  if (do_conf) {
    gt_c = ldg_px<NP>(gtp);
    if (ys + 1 < ye) gt_n = ldg_px<NP>(gtp + T_w);
    gtp += 2 * T_w;
  }
  const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match, bgl4 = 0x01010101u * (unsigned)p.bg_label;

#pragma unroll 1
  for (int yl = ys; yl < ye; yl++) {
    u64 w[G];
    unsigned int flags;
    {
      const ulonglong2 t0 = lds_u64x2(rt);
      if (G == 1) { w[0] = t0.x; flags = (unsigned int)t0.y; }
      else {
        w[0] = t0.x; w[1] = t0.y;
        const ulonglong2 t1 = lds_u64x2(rt + 16u);
        if (G == 2) flags = (unsigned int)t1.x;
        else {
          w[2] = t1.x;
          if (G == 3) flags = (unsigned int)t1.y;
          else { const ulonglong2 t2 = lds_u64x2(rt + 32u); w[G - 1] = t2.x; flags = (unsigned int)t2.y; }
        }
      }
    }
    const unsigned int bg4 = bg_c, gt4 = gt_c;
    bg_c = bg_n; gt_c = gt_n;
    if (yl + 2 < ye) {
      if (has_bg) { bg_n = ldg_px<NP>(bgp); bgp += T_w; }
      if (do_conf) { gt_n = ldg_px<NP>(gtp); gtp += T_w; }
    }
    if (flags) {  // the bracketing source rows of at least one group moved down by one
#pragma unroll
      for (int gi = 0; gi < G; gi++) {
        if ((flags >> gi) & 1u) {
          const int2 ro = lds_i2(ro_a + 8u * gi);
          u64 Hn[K][NP];
          load_h(gi, ro.y, Hn);
#pragma unroll
          for (int k = 0; k < K; k++)
#pragma unroll
            for (int q = 0; q < NP; q++) { Dh[gi][k][q] = sub2(Hn[k][q], Hb[gi][k][q]); Hb[gi][k][q] = Hn[k][q]; }
        }
      }
      rebase();
    }
    rt += RS; ro_a += 8u * G;

    unsigned int lab4 = 0, unc = 0;
    int labs[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
      float d0[K], d1[K];
#pragma unroll
      for (int k = 0; k < K; k++) {
        u64 acc = base[k][q];
#pragma unroll
        for (int gi = 0; gi < G; gi++) acc = fma2(w[gi], Dh[gi][k][q], acc);
        unpack2(acc, d0[k], d1[k]);
      }
      const bool c0 = decide_diff<K, C>(d0, cls, tau, labs[2 * q]);
      const bool c1 = decide_diff<K, C>(d1, cls, tau, labs[2 * q + 1]);
      unc |= (c0 ? 0u : 1u) << (2 * q);
      unc |= (c1 ? 0u : 1u) << (2 * q + 1);
      lab4 |= ((unsigned)labs[2 * q] | ((unsigned)labs[2 * q + 1] << 8)) << (16 * q);
    }
    if (unc) push_uncertain(ctl, queue, b, yl, x, unc);
    if (do_conf) {
#pragma unroll
      for (int j = 0; j < 2 * NP; j++) {
        const unsigned int gg = (gt4 >> (8 * j)) & 0xffu;
        if (gg < (unsigned)C && !((unc >> j) & 1u)) {
          const unsigned int bn = gg * C + labs[j];
          const u64 inc = 1ull << (8 * (bn & 7));
          if (bn < 8) cnt_lo += inc; else cnt_hi += inc;
        }
      }
    }
    if (has_label) {
      unsigned int o = lab4;
      if (has_bg) {
        const unsigned int eq = __vcmpeq4(bg4, m4);
        o = (bgl4 & eq) | (lab4 & ~eq);
      }
      stg_px<NP>(lbp, o);
      lbp += T_w;
    }
  }
}

template <int C, int V, int G, int F, int NP>
__global__ void __launch_bounds__(kFMaxThreads, 1) fuse_filter_kernel(const __grid_constant__ FuseParams p,
                                                                      const __grid_constant__ FilterGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FCtl* ctl = reinterpret_cast<FCtl*>(smem_raw + g.ctl_off);
  int2* rowoff = reinterpret_cast<int2*>(smem_raw + g.rowoff_off);        // [T_h][G] byte offsets of rows i0, i1 inside a map
  int4* colA = reinterpret_cast<int4*>(smem_raw + g.cola_off);            // [G][GXP] byte offsets {j0,j1 of col 0; j0,j1 of col 1}
  float4* colB = reinterpret_cast<float4*>(smem_raw + g.colb_off);        // [G][GXP] {l0 col0, l0 col1, l1 col0, l1 col1}
  float4* lowrow = reinterpret_cast<float4*>(smem_raw + g.lowrow_off);    // [low_h][V] {l0, l1, byte offset i0, byte offset i1} in the RAW view
  float4* lowcol = reinterpret_cast<float4*>(smem_raw + g.lowcol_off);    // [low_w][V]
  float* ymap = reinterpret_cast<float*>(smem_raw + g.ymap_off);          // [G][C-1][ho][wo] difference maps
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + g.queue_off);  // [kFQueueCap] (y << 16) | x
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);          // 2 staging buffers

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int T_h = p.T_h, T_w = p.T_w;
  constexpr bool RT = F < 0;
  constexpr int RS = 16 * ((G + 2) / 2);
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool need_low = RT ? (p.lowres_out != nullptr && p.low_fh > 0) : ((F & 8) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  constexpr int BINS = C * C;

  // ---- one-time setup: barriers, tables (whole-tile geometry is the same for every tile) ----------------------------
  if (tid == 0) {
    mbar_init(&ctl->full[0], 1);
    mbar_init(&ctl->full[1], 1);
    mbar_init(&ctl->empty[0], g.cwarps);
    mbar_init(&ctl->empty[1], g.cwarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->maxbits[0] = ctl->maxbits[1] = 0u;
    ctl->qcount[0] = ctl->qcount[1] = 0u;
  }
  for (int i = tid; i < 64; i += nthreads) ctl->hist[i] = 0;
  for (int y = tid; y < T_h; y += nthreads) {
    unsigned char* row = smem_raw + g.rowtab_off + RS * y;
    bool strip_start = false;
    for (int q = 0; q < g.S; q++) strip_start |= (y == g.strip_y0[q]);
    unsigned int f = 0;
    for (int gi = 0; gi < G; gi++) {
      const Lerp L = pisto_src_index(g.g_scale_h[gi], y, g.g_ho[gi], false);
      reinterpret_cast<float2*>(row)[gi] = make_float2(-L.l0, -L.l0);
      rowoff[y * G + gi] = make_int2(4 * L.i0 * g.g_wo[gi], 4 * L.i1 * g.g_wo[gi]);
      if (!strip_start) {  // first row of a strip: both source rows are loaded before the row loop
        const Lerp P = pisto_src_index(g.g_scale_h[gi], y - 1, g.g_ho[gi], false);
        if (P.i0 != L.i0 || P.i1 != L.i1) f |= 1u << gi;  // up-sampling: the pair moves down by exactly one row
      }
    }
    reinterpret_cast<uint2*>(row)[G] = make_uint2(f, 0u);
  }
  for (int i = tid; i < G * g.GXP; i += nthreads) {
    const int gi = i / g.GXP, gx = i - gi * g.GXP;
    const Lerp L0 = pisto_src_index(g.g_scale_w[gi], 2 * gx, g.g_wo[gi], g.g_same_w[gi]);
    const Lerp L1 = pisto_src_index(g.g_scale_w[gi], 2 * gx + 1, g.g_wo[gi], g.g_same_w[gi]);
    colA[i] = make_int4(4 * L0.i0, 4 * L0.i1, 4 * L1.i0, 4 * L1.i1);
    colB[i] = make_float4(L0.l0, L1.l0, L0.l1, L1.l1);
  }
  if (need_low) {
    for (int i = tid; i < p.low_h * V; i += nthreads) {
      const int ly = i / V, v = i - ly * V;
      const ViewDev& vw = p.view[v];
      const Lerp L = pisto_src_index(vw.scale_h, ly * p.low_fh + p.low_fh / 2, vw.map.ho, vw.same_h);
      lowrow[i] = make_float4(L.l0, L.l1, __int_as_float(g.vbase[v] + L.i0 * g.vrow[v]), __int_as_float(g.vbase[v] + L.i1 * g.vrow[v]));
    }
    for (int i = tid; i < p.low_w * V; i += nthreads) {
      const int lx = i / V, v = i - lx * V;
      const ViewDev& vw = p.view[v];
      const Lerp L = pisto_src_index(vw.scale_w, lx * p.low_fw + p.low_fw / 2, vw.map.wo, vw.same_w);
      lowcol[i] = make_float4(L.l0, L.l1, __int_as_float(L.i0 * g.vcol[v]), __int_as_float(L.i1 * g.vcol[v]));
    }
  }

  // does tile n read its views at all?  (single-label tiles without the 32x32 export do not)
  auto tile_needs_views = [&](int n) -> bool {
    if (need_low) return true;
    return pisto_tile_presence(p, n).single < 0;
  };
  // producer lane: fetch every view of tile n into staging buffer b
  auto issue_tile = [&](int n, int b) {
    float* buf = vsm + b * g.buf_floats;
    uint32_t total = 0;
#pragma unroll 1
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const char* start = reinterpret_cast<const char*>(vw.logits + (long long)n * vw.tile_stride);
      const char* end = start + (size_t)C * vw.h * vw.w * sizeof(float);
      const char* a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(start) & ~(uintptr_t)15);
      const char* a1 = reinterpret_cast<const char*>((reinterpret_cast<uintptr_t>(end) + 15) & ~(uintptr_t)15);
      char* dst = reinterpret_cast<char*>(buf + g.view_off[v]);
      if (n == p.N - 1) {
        // never read past the end of the caller's buffer: copy whole 16-byte units only, the (<16-byte) tail by hand
        a1 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(end) & ~(uintptr_t)15);
        if (a1 < a0) a1 = a0;
        const char* t = a1 > start ? a1 : start;
        for (; t < end; t += 4) *reinterpret_cast<float*>(dst + (t - a0)) = *reinterpret_cast<const float*>(t);
      }
      const uint32_t bytes = (uint32_t)(a1 - a0);
      if (bytes) bulk_g2s(dst, a0, bytes, &ctl->full[b]);
      total += bytes;
    }
    mbar_arrive_expect_tx(&ctl->full[b], total);
  };

  __syncthreads();  // barriers + tables visible to every warp

  const int ncomp = g.cwarps * 32;  // compute threads; the last warp of the CTA is the producer
  if (tid >= ncomp) {
    // ===== producer warp: claims tiles, publishes their ids, fetches their views (TMA) one tile ahead of the compute warps
    if (tid == ncomp) {
      const long long tile_px = (long long)T_h * T_w;
      for (int k = 0;; k++) {
        const int b = k & 1;
        if (k >= 2) mbar_wait(&ctl->empty[b], ((k >> 1) - 1) & 1);  // all compute warps are done with buffer b
        const int t = atomicAdd(g.counter, 1);
        const int tile = t < p.N ? t : -1;
        ctl->tile[b] = tile;
        if (tile >= 0 && tile_needs_views(tile)) issue_tile(tile, b);
        else mbar_arrive(&ctl->full[b]);
        if (tile < 0) break;
        // pull the tile's byte masks into L2 ahead of the per-row loads
        if (tile_px % 16 == 0) {
          if (has_bg && ((uintptr_t)p.bg & 15) == 0) bulk_prefetch_l2(p.bg + tile * tile_px, (uint32_t)tile_px);
          if (do_conf && ((uintptr_t)p.gt & 15) == 0) bulk_prefetch_l2(p.gt + tile * tile_px, (uint32_t)tile_px);
        }
      }
    }
    return;
  }

  // ===== compute warps
  const int grp = tid % g.GX, strip = min(tid / g.GX, g.S - 1);
  const bool worker = tid < g.GX * g.S;
  const int x = 2 * NP * grp;
  const int ys = g.strip_y0[strip];
  const int ye = g.strip_y0[strip + 1];
  const uint32_t rowtab_s = smem_u32(smem_raw + g.rowtab_off), rowoff_s = smem_u32(rowoff), ymap_s = smem_u32(ymap);
  const uint32_t colA_t = smem_u32(colA) + 16u * NP * grp, colB_t = smem_u32(colB) + 16u * NP * grp;
  const int nt = ncomp;  // cooperative loops below run over the compute threads only
  const int npt = need_low ? p.low_h * p.low_w : 0;

  for (int k = 0;; k++) {
    const int b = k & 1;
    mbar_wait(&ctl->full[b], (k >> 1) & 1);  // tile id published, views (if any) landed
    const int n = ctl->tile[b];
    if (n < 0) break;
    const TilePresence tp = pisto_tile_presence(p, n);
    const bool multi = tp.single < 0;
    // shared-memory byte address of view v's data in this tile's staging buffer (incl. the 0..3-float alignment shift)
    uint32_t vb[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
      vb[v] = smem_u32(vsm + b * g.buf_floats + g.view_off[v]) + sh;
    }
    // classes in play, lowest index first
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

    u64 cnt_lo = 0, cnt_hi = 0;

    // ---- pre-pass: low-resolution difference maps of every scale group + max |x| ------------------------------------
    if (multi && P >= 2) {
      unsigned int mx = 0;
#pragma unroll 1
      for (int gi = 0; gi < G; gi++) {
        const int wo = g.g_wo[gi], cells = g.g_ho[gi] * wo;
        float* ym = ymap + (g.g_ybytes[gi] >> 2);
        for (int idx = tid; idx < cells; idx += nt) {
          const int i = idx / wo, j = idx - i * wo;
          float y[C - 1];
#pragma unroll
          for (int q = 0; q < C - 1; q++) y[q] = 0.f;
          bool first = true;
#pragma unroll
          for (int v = 0; v < V; v++) {
            if (g.group_of[v] != gi) continue;
            const uint32_t a = vb[v] + g.vbase[v] + i * g.vrow[v] + j * g.vcol[v];
            const float x0 = lds_f32(a + cls[0] * g.plane_bytes[v]);
            mx = max(mx, __float_as_uint(x0) & 0x7fffffffu);
#pragma unroll
            for (int q = 0; q < C - 1; q++) {
              if (q + 1 < P) {
                const float xq = lds_f32(a + cls[q + 1] * g.plane_bytes[v]);
                mx = max(mx, __float_as_uint(xq) & 0x7fffffffu);
                const float t = __fsub_rn(xq, x0);
                y[q] = first ? t : __fadd_rn(y[q], t);
              }
            }
            first = false;
          }
#pragma unroll
          for (int q = 0; q < C - 1; q++)
            if (q + 1 < P) ym[q * cells + idx] = y[q];
        }
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if ((tid & 31) == 0) atomicMax(&ctl->maxbits[b], mx);
    }
    bar_sync(1, ncomp);  // difference maps + max visible

    bool exact_all = false;
    if (multi) {
      float tau = 0.f;
      if (P >= 2) {
        const float A = __fmul_rn((float)V, __uint_as_float(ctl->maxbits[b]));
        tau = __fmaf_rn(A, g.tau_coef, g.tau_abs);
        if (!(A < 1e9f)) exact_all = true;  // non-finite or absurd magnitudes: follow the reference everywhere
      } else {
        exact_all = true;                   // empty presence vector
      }
      if (!exact_all && worker && ys < ye) {
        if (P == 2) filter_rows<C, G, F, NP, 1>(p, g, ctl, queue, b, rowtab_s, rowoff_s, colA_t, colB_t, ymap_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
        else if (P == 3) filter_rows<C, G, F, NP, 2>(p, g, ctl, queue, b, rowtab_s, rowoff_s, colA_t, colB_t, ymap_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
        else if (C >= 4 && P == 4) filter_rows<C, G, F, NP, (C >= 4 ? 3 : 1)>(p, g, ctl, queue, b, rowtab_s, rowoff_s, colA_t, colB_t, ymap_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
      }
    } else {
      // single-label tile (infer_pseudo_masks.py:71-73): constant label + background overwrite, 16 pixels per thread-step
      const long long tpx = (long long)T_h * T_w;
      const long long base = (long long)n * tpx;
      const unsigned int lab4 = 0x01010101u * (unsigned)tp.single, bgl4 = 0x01010101u * (unsigned)p.bg_label;
      // packed 8-bit confusion counters: at most 255 pixels per thread between flushes
      const bool vec_ok = (tpx % 16 == 0) && ((((uintptr_t)p.bg | (uintptr_t)p.gt | (uintptr_t)p.label_out) & 15) == 0) &&
                          (!do_conf || 16 * ((tpx / 16 + nt - 1) / nt) <= 255);
      const long long nvec = vec_ok ? tpx / 16 : 0;
      const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match;
      auto sel4 = [&](unsigned int w) -> unsigned int {
        const unsigned int eq = __vcmpeq4(w, m4);  // 0xff in every byte equal to bg_match
        return (bgl4 & eq) | (lab4 & ~eq);
      };
      constexpr int UN = 4;  // independent 16-byte loads in flight per thread
      for (long long i0 = tid; i0 < nvec; i0 += (long long)UN * nt) {
        uint4 bgv[UN], gv[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const long long i = i0 + (long long)u * nt;
          if (i < nvec) {
            if (has_bg) bgv[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + base) + i);
            if (do_conf) gv[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + base) + i);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const long long i = i0 + (long long)u * nt;
          if (i < nvec) {
            uint4 o = make_uint4(lab4, lab4, lab4, lab4);
            if (has_bg) o = make_uint4(sel4(bgv[u].x), sel4(bgv[u].y), sel4(bgv[u].z), sel4(bgv[u].w));
            if (do_conf) {
              const unsigned int gw[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
              for (int q = 0; q < 4; q++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  const unsigned int gg = (gw[q] >> (8 * j)) & 0xffu;
                  if (gg < (unsigned)C) { const unsigned int bn = gg * C + tp.single; const u64 inc = 1ull << (8 * (bn & 7)); if (bn < 8) cnt_lo += inc; else cnt_hi += inc; }
                }
            }
            if (has_label) reinterpret_cast<uint4*>(p.label_out + base)[i] = o;
          }
        }
      }
      for (long long i = nvec * 16 + tid; i < tpx; i += nt) {  // unaligned / ragged remainder
        unsigned int o = (unsigned)tp.single;
        if (has_bg && p.bg[base + i] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
        if (do_conf) {
          const unsigned int gg = p.gt[base + i];
          if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + tp.single], 1u);
        }
        if (has_label) p.label_out[base + i] = (uint8_t)o;
      }
    }
    bar_sync(1, ncomp);  // every strip done: the queue is complete
    if (tid == 0) { ctl->maxbits[b] = 0u; ctl->qcount[b ^ 1] = 0u; }

    // ---- exact pass: 32x32 gather points (always) + queued pixels (or the whole tile) --------------------------------
    int nfix = 0;
    if (multi) {
      const unsigned int nq = ctl->qcount[b];
      if (nq > (unsigned)kFQueueCap && !exact_all) { exact_all = true; cnt_lo = cnt_hi = 0; }  // overflow: recount the whole tile
      nfix = exact_all ? T_h * T_w : (int)nq;
    }
    for (int it = tid; it < npt + nfix; it += nt) {
      float a[C];
      if (it < npt) {
        const int ly = it / p.low_w, lx = it - ly * p.low_w;
#pragma unroll
        for (int v = 0; v < V; v++) {
          const float4 R = lowrow[ly * V + v], Q = lowcol[lx * V + v];
          const uint32_t r0 = vb[v] + __float_as_int(R.z), r1 = vb[v] + __float_as_int(R.w);
          const uint32_t c0 = __float_as_int(Q.z), c1 = __float_as_int(Q.w);
#pragma unroll
          for (int c = 0; c < C; c++) {
            const uint32_t pl = c * g.plane_bytes[v];
            const float h0 = __fmaf_rn(Q.x, lds_f32(r0 + pl + c0), __fmul_rn(Q.y, lds_f32(r0 + pl + c1)));
            const float h1 = __fmaf_rn(Q.x, lds_f32(r1 + pl + c0), __fmul_rn(Q.y, lds_f32(r1 + pl + c1)));
            const float u = __fmaf_rn(R.x, h0, __fmul_rn(R.y, h1));
            a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
          }
        }
#pragma unroll
        for (int c = 0; c < C; c++) p.lowres_out[((long long)n * C + c) * npt + it] = pisto_div_views(a[c], p.dec);
      } else {
        const int j = it - npt;
        int yy, xx;
        if (exact_all) { yy = j / T_w; xx = j - yy * T_w; }
        else { const uint32_t e = queue[j]; yy = (int)(e >> 16); xx = (int)(e & 0xffffu); }
#pragma unroll
        for (int v = 0; v < V; v++) {
          const ViewDev& vw = p.view[v];
          const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, vw.same_h);
          const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, vw.same_w);
          const uint32_t r0 = vb[v] + g.vbase[v] + Ly.i0 * g.vrow[v], r1 = vb[v] + g.vbase[v] + Ly.i1 * g.vrow[v];
          const uint32_t c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
#pragma unroll
          for (int c = 0; c < C; c++) {
            const uint32_t pl = c * g.plane_bytes[v];
            const float h0 = __fmaf_rn(Lx.l0, lds_f32(r0 + pl + c0), __fmul_rn(Lx.l1, lds_f32(r0 + pl + c1)));
            const float h1 = __fmaf_rn(Lx.l0, lds_f32(r1 + pl + c0), __fmul_rn(Lx.l1, lds_f32(r1 + pl + c1)));
            const float u = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
            a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
          }
        }
        const int lab = pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
        const long long pix = ((long long)n * T_h + yy) * T_w + xx;
        if (do_conf) {
          const unsigned int gg = p.gt[pix];
          if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + lab], 1u);
        }
        if (has_label) {
          unsigned int o = (unsigned)lab;
          if (has_bg && p.bg[pix] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
          p.label_out[pix] = (uint8_t)o;
        }
      }
    }
    if (do_conf) {
      // every lane of every compute warp reaches this point: full-mask warp reductions are safe
#pragma unroll
      for (int bn = 0; bn < BINS; bn++) {
        unsigned int cv = (unsigned int)(((bn < 8 ? cnt_lo : cnt_hi) >> (8 * (bn & 7))) & 0xffull);
        cv = __reduce_add_sync(0xffffffffu, cv);
        if ((tid & 31) == 0 && cv) atomicAdd(&ctl->hist[bn], cv);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&ctl->empty[b]);  // this warp is done with staging buffer b
  }
  if (do_conf) {
    bar_sync(1, ncomp);
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool make_filter_geom(const pisto_ctx* h, const FuseParams& p, int NP, int G_expected, FilterGeom* g) {
  memset(g, 0, sizeof(*g));
  if (p.T_w % (2 * NP)) return false;
  const int GX = p.T_w / (2 * NP);
  if (GX > kFMaxThreads - 32) return false;  // one warp of the CTA is the producer
  int S = (kFMaxThreads - 32) / GX;
  if (S > p.T_h) S = p.T_h;
  if (S > 32) S = 32;
  // scale groups: views with the same de-augmented size share their interpolation weights
  int G = 0, nmax = 0, cnt[kFMaxGroups] = {0, 0, 0, 0};
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    if (vw.map.ho >= p.T_h) return false;  // same-size / down-sampling rows: the source-row pair does not move by exactly one
    int gi = -1;
    for (int q = 0; q < G; q++)
      if (g->g_ho[q] == vw.map.ho && g->g_wo[q] == vw.map.wo) gi = q;
    if (gi < 0) {
      if (G == kFMaxGroups) return false;
      gi = G++;
      g->g_ho[gi] = vw.map.ho; g->g_wo[gi] = vw.map.wo; g->g_same_w[gi] = vw.same_w;
      g->g_scale_h[gi] = vw.scale_h; g->g_scale_w[gi] = vw.scale_w;
    }
    g->group_of[v] = gi;
    if (++cnt[gi] > nmax) nmax = cnt[gi];
    g->plane_bytes[v] = 4 * vw.h * vw.w;
    g->vbase[v] = 4 * (vw.map.a0 * vw.w + vw.map.b0);
    g->vrow[v] = 4 * (vw.map.ai * vw.w + vw.map.bi);
    g->vcol[v] = 4 * (vw.map.aj * vw.w + vw.map.bj);
  }
  if (G != G_expected) return false;
  // strips: when the flag pattern of the row table repeats with a period that divides T_h / S, every strip starts at
  // the same phase and the lanes of a warp that straddles two strips refill together
  int period = 1;
  for (int q = 0; q < G; q++) {
    int a = p.T_h, b = g->g_ho[q];
    while (b) { int t = a % b; a = b; b = t; }
    const int pv = p.T_h / a;
    int x = period, y = pv;
    while (y) { int t = x % y; x = y; y = t; }
    period = period / x * pv;
  }
  if (period > 1 && p.T_h % period == 0) {
    // prefer a strip count whose strips are whole periods and that fills the CTA best
    int best = 0;
    for (int s = 1; s <= S; s++)
      if ((p.T_h / period) % s == 0) best = s;
    if (best * 2 > S) S = best;
  }
  g->GX = GX; g->S = S; g->GXP = p.T_w / 2;
  g->cwarps = (GX * S + 31) / 32;
  g->threads = g->cwarps * 32 + 32;
  int rps = 0;
  for (int q = 0; q <= S; q++) g->strip_y0[q] = (int)((long long)p.T_h * q / S);
  for (int q = 0; q < S; q++) rps = max(rps, g->strip_y0[q + 1] - g->strip_y0[q]);
  if (rps * 2 * NP > 255) return false;  // packed 8-bit confusion counters
  int fl = 0;
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    g->view_off[v] = fl;
    fl += (p.C * vw.h * vw.w + 3 /* alignment shift */ + 3 /* tail */ + 3) & ~3;
  }
  g->buf_floats = fl;
  // decision threshold: both evaluations differ from the exact difference of the view sums by at most
  // cE * 2^-24 * A each way; on top of that the lead must cover the softmax margin 2.4e-7 * |a| + V * 2e-6 of common.cuh
  const float cE = 2.f * nmax + 4.f * G + 2.f * p.V + 16.f;
  g->tau_coef = 2.f * cE * 5.9604645e-8f + 2.5e-7f;
  g->tau_abs = p.dec.margin_abs * 1.01f;
  const int RS = 16 * ((G + 2) / 2);
  int off = 0;
  g->ctl_off = off; off += (int)((sizeof(FCtl) + 127) & ~127u);
  g->rowtab_off = off; off += RS * p.T_h;
  g->rowoff_off = off; off += 8 * G * p.T_h; off = (off + 15) & ~15;
  g->cola_off = off; off += 16 * G * g->GXP;
  g->colb_off = off; off += 16 * G * g->GXP;
  g->lowrow_off = off; off += 16 * p.V * (p.lowres_out && p.low_fh > 0 ? p.low_h : 0);
  g->lowcol_off = off; off += 16 * p.V * (p.lowres_out && p.low_fh > 0 ? p.low_w : 0);
  g->ymap_off = off;
  for (int q = 0; q < G; q++) {
    g->g_ybytes[q] = off - g->ymap_off;
    g->g_mapbytes[q] = 4 * g->g_ho[q] * g->g_wo[q];
    off += (p.C - 1) * g->g_mapbytes[q];
    off = (off + 15) & ~15;
  }
  g->queue_off = off; off += 4 * kFQueueCap;
  off = (off + 127) & ~127;
  g->views_off = off; off += 2 * 4 * fl;
  g->smem_bytes = off;
  return off <= h->smem_optin - 1024;
}

template <int C, int V, int G, int F, int NP>
int launch_filter(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  FilterGeom g;
  if (!make_filter_geom(h, p, NP, G, &g)) return PISTO_OK;  // not launched: caller falls back
  auto kern = fuse_filter_kernel<C, V, G, F, NP>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  g.counter = h->sched + (h->sched_next++ % PISTO_SCHED_SLOTS);
  PISTO_CUDA(cudaMemsetAsync(g.counter, 0, sizeof(int), st));
  const int grid = p.N < h->sm_count ? p.N : h->sm_count;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}

}  // namespace

static inline int pisto_filter_flags(const FuseParams& p) {
  return (p.bg ? 1 : 0) | ((p.conf && p.gt) ? 2 : 0) | ((p.lowres_out && p.low_fh > 0) ? 8 : 0) | (p.label_out ? 16 : 0);
}

// number of scale groups of a view set (0: more than the kernel supports)
static inline int pisto_filter_groups(const FuseParams& p) {
  int G = 0, ho[kFMaxGroups], wo[kFMaxGroups];
  for (int v = 0; v < p.V; v++) {
    bool found = false;
    for (int q = 0; q < G; q++) found |= (ho[q] == p.view[v].map.ho && wo[q] == p.view[v].map.wo);
    if (!found) {
      if (G == kFMaxGroups) return 0;
      ho[G] = p.view[v].map.ho; wo[G] = p.view[v].map.wo; G++;
    }
  }
  return G;
}

template <int C, int V, int G, int NP>
static int pisto_launch_filter_cvg(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  switch (pisto_filter_flags(p)) {
    case 25: return launch_filter<C, V, G, 25, NP>(h, p, st, launched);  // bg + labels + 32x32        (config 2)
    case 19: return launch_filter<C, V, G, 19, NP>(h, p, st, launched);  // bg + gt/conf + labels      (config 1)
    case 18: return launch_filter<C, V, G, 18, NP>(h, p, st, launched);  // gt/conf + labels           (config 3, mIoUMask.forward)
    default: return launch_filter<C, V, G, -1, NP>(h, p, st, launched);
  }
}
