// Streaming fusion kernel -- the hot path of the library for tiles whose views fit in shared memory whole
// (BASELINE configs 1, 2, 3: 224-px tiles, stride-8 multi-scale logits).
//
// One persistent CTA per SM; tiles are claimed from a global atomic counter (single-label tiles are ~20x cheaper
// than multi-label ones, so static assignment would leave SMs idle).
//
//   TMA pipeline   Every view of a tile is ONE contiguous run of C*h*w floats in HBM.  Thread 0 fetches the runs of
//                  the NEXT tile with cp.async.bulk (1-D TMA, mbarrier complete_tx) into the other half of a double
//                  buffer while all warps compute the current tile.  Runs are not 16-byte multiples (21*21*3 floats),
//                  so the enclosing 16-byte-aligned span is copied and the 0..3-float shift is folded into the shared
//                  memory base when the values are read.  De-augmentation (flip / rot90) is an affine index map on the
//                  staged raw view -- never materialised.
//   tables         Built once per launch (whole-tile geometry is the same for every tile): per (view, row) the
//                  vertical lerp weights, pre-duplicated {l0,l0,l1,l1} so one 128-bit shared load yields both packed
//                  operands; per row a 2-bit-per-view "source rows moved" word; per (view, col) the horizontal lerp
//                  weights and column offsets.
//   arithmetic     A thread owns 2 adjacent output columns and streams down a strip of rows, keeping for every view
//                  the horizontally interpolated values of the two bracketing source rows in registers (Ha, Hb), both
//                  columns PACKED in one 64-bit register.  Per (row, view, class) the work is three packed
//                  instructions  t = mul.f32x2(l1, Hb); o = fma.f32x2(l0, Ha, t); acc = add.f32x2(acc, o)
//                  -- the separable form of torch's fma(h0, fma(w0,a,w1*b), h1*fma(w0,c,w1*d)) with identical
//                  association and rounding (bit-exact).  FFMA2/FMUL2/FADD2 occupy the FP32 pipe for
//                  two cycles but take one issue slot (profiles/r01/probe_microbench.txt), which leaves the other
//                  slot for the table loads / branches / integer work of the loop.
//   epilogue       mask / argmax (margin fast path, exact slow path: common.cuh), confusion counters packed in two
//                  64-bit registers, background overwrite, 2-byte label store, 32x32 logit gather.
//
// Roofline: FP32-pipe-bound for V >= 2 (3*C*V packed-lane operations per pixel); HBM traffic is the compulsory
// minimum (each input byte read once by TMA, each output byte written once).
#pragma once
#include "fuse_common.cuh"
#include "sm100_prims.cuh"

namespace {

// The same softmax for a pixel PAIR held in f32x2 lanes (the streaming kernels work on column pairs).  The MUFU unit (16 lanes per
// clock and SM) is what bounds PROB_MEAN, so only the C exponentials per pixel stay on it: the reciprocal of the sum -- in [1, C] --
// is a packed Newton iteration on the FMA pipe (integer-subtraction seed, relative error 0.12 -> 1.4e-2 -> 2e-4 -> 4e-8 after three
// steps), and subtracting the maximum is folded into the scaling by log2(e): t = x * log2e - m * log2e (one packed fma per class).
// Probabilities differ from pisto_softmax_fast by < 2e-7.
template <int C>
__device__ __forceinline__ void pisto_softmax_fast2(u64 (&u)[C]) {
  float a0, a1, m0, m1;
  unpack2(u[0], m0, m1);
#pragma unroll
  for (int c = 1; c < C; c++) { unpack2(u[c], a0, a1); m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1); }
  const u64 L2E = pack2(1.4426950408889634f, 1.4426950408889634f);
  const u64 nm = mul2(pack2(-m0, -m1), L2E);
  u64 sum = 0;
#pragma unroll
  for (int c = 0; c < C; c++) {
    unpack2(fma2(u[c], L2E, nm), a0, a1);
    u[c] = pack2(pisto_ex2_approx(a0), pisto_ex2_approx(a1));
    sum = c == 0 ? u[c] : add2(sum, u[c]);
  }
  unpack2(sum, a0, a1);
  u64 r = pack2(__uint_as_float(0x7EF311C7u - __float_as_uint(a0)), __uint_as_float(0x7EF311C7u - __float_as_uint(a1)));
  const u64 one = pack2(1.f, 1.f), ns = pack2(-a0, -a1);
#pragma unroll
  for (int it = 0; it < 3; it++) r = fma2(r, fma2(ns, r, one), r);
#pragma unroll
  for (int c = 0; c < C; c++) u[c] = mul2(u[c], r);
}


#ifndef PISTO_STREAM_THREADS
#define PISTO_STREAM_THREADS 384
#endif
constexpr int kMaxThreads = PISTO_STREAM_THREADS;

struct StreamGeom {
  int GX, S, rows_per_strip, threads;  // threads = compute warps * 32 + one producer warp
  int cwarps;                          // compute warps
  int strip_y0[17];                    // row range of strip s is [strip_y0[s], strip_y0[s+1])
  int view_off[PISTO_MAX_VIEWS];    // float offset of each view inside one staging buffer (16-byte aligned)
  int plane_bytes[PISTO_MAX_VIEWS]; // h*w*4
  int buf_floats;                   // floats per staging buffer
  int ctl_off, flags_off, rowoff_off, rowtab_off, cola_off, colb_off, views_off;  // byte offsets into dynamic smem
  int smem_bytes;
  int* counter;                     // global tile counter (zeroed before the launch)
};

struct Ctl {
  uint64_t full[2];   // producer -> consumers: tile id published (+ TMA bytes landed)
  uint64_t empty[2];  // consumers -> producer: staging buffer may be refilled (one arrival per compute warp)
  int tile[2];
  unsigned int hist[64];
};

// first index of the maximum, the maximum and the runner-up of v[0..C)
template <int C>
__device__ __forceinline__ void top2(const float (&v)[C], int& bi, float& bv, float& sv) {
  bi = 0; bv = v[0]; sv = -INFINITY;
#pragma unroll
  for (int c = 1; c < C; c++) {
    const bool gt = v[c] > bv;
    sv = fmaxf(sv, fminf(bv, v[c]));
    bi = gt ? c : bi;
    bv = fmaxf(bv, v[c]);
  }
}

// near tie / NaN / absurd magnitude / MULTIPLY mask: follow the reference operation by operation (kept out of line so
// that the row loop's instruction footprint stays small)
template <int C>
__device__ __noinline__ void decide_pair_slow(const u64 (&acc)[C], uint32_t present_bits, const DecideCfg& cfg, bool ok0, bool ok1,
                                              int& lab0, int& lab1) {
  float a0[C], a1[C];
#pragma unroll
  for (int c = 0; c < C; c++) unpack2(acc[c], a0[c], a1[c]);
  if (!ok0) lab0 = pisto_decide<C>(a0, present_bits, cfg, false, nullptr);
  if (!ok1) lab1 = pisto_decide<C>(a1, present_bits, cfg, false, nullptr);
}

// Labels of the two pixels of a packed pair from the undivided view sums acc[] (see pisto_decide in common.cuh for
// the argument behind the fast path); madd2[c] is (0,0) for usable classes and (-inf,-inf) for classes masked out by
// the tile's presence vector.
template <int C>
__device__ __forceinline__ void decide_pair(const u64 (&acc)[C], const u64 (&madd2)[C], uint32_t present_bits, const DecideCfg& cfg,
                                            bool fast_ok, int& lab0, int& lab1) {
  float v0[C], v1[C];
  u64 chk = acc[0];
#pragma unroll
  for (int c = 0; c < C; c++) {
    unpack2(add2(acc[c], madd2[c]), v0[c], v1[c]);
    if (c) chk = add2(chk, acc[c]);
  }
  float chk0, chk1;
  unpack2(chk, chk0, chk1);
  float bv0, sv0, bv1, sv1;
  top2<C>(v0, lab0, bv0, sv0);
  top2<C>(v1, lab1, bv1, sv1);
  const bool ok0 = fast_ok && (__fsub_rn(bv0, sv0) > __fmaf_rn(fabsf(bv0), 2.4e-7f, cfg.margin_abs)) && (fabsf(chk0) < 1e30f) && (bv0 > -1e9f);
  const bool ok1 = fast_ok && (__fsub_rn(bv1, sv1) > __fmaf_rn(fabsf(bv1), 2.4e-7f, cfg.margin_abs)) && (fabsf(chk1) < 1e30f) && (bv1 > -1e9f);
  if (!(ok0 && ok1)) decide_pair_slow<C>(acc, present_bits, cfg, ok0, ok1, lab0, lab1);
}

// PAIRS: views 2k and 2k+1 (a scale and its flipped twin) share their row geometry -> one weight load / one flag test per pair.
template <int C, int V, bool PROB, int F, bool PAIRS>
__global__ void __launch_bounds__(kMaxThreads, 1) fuse_stream_kernel(const __grid_constant__ FuseParams p,
                                                                     const __grid_constant__ StreamGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem_raw + g.ctl_off);
  unsigned int* rowflags = reinterpret_cast<unsigned int*>(smem_raw + g.flags_off);  // [T_h]
  int2* rowoff = reinterpret_cast<int2*>(smem_raw + g.rowoff_off);                   // [T_h][V] byte offsets of rows i0, i1
  float4* rowtab = reinterpret_cast<float4*>(smem_raw + g.rowtab_off);               // [T_h][V] {l0,l0,l1,l1}
  int4* colA = reinterpret_cast<int4*>(smem_raw + g.cola_off);                       // [V][GX] byte offsets {i0,i1 of col 0; i0,i1 of col 1}
  float4* colB = reinterpret_cast<float4*>(smem_raw + g.colb_off);                   // [V][GX] {l0 col0, l0 col1, l1 col0, l1 col1}
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);                     // 2 staging buffers

  const int tid = threadIdx.x, nt = blockDim.x;
  const int T_h = p.T_h, T_w = p.T_w;
  constexpr bool RT = F < 0;
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool has_fused = RT ? (p.fused_out != nullptr) : ((F & 4) != 0);
  const bool need_low = RT ? (p.lowres_out != nullptr && p.low_fh > 0) : ((F & 8) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  constexpr int BINS = C * C;

  // ---- one-time setup: barriers, tables ---------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&ctl->full[0], 1);
    mbar_init(&ctl->full[1], 1);
    mbar_init(&ctl->empty[0], g.cwarps);
    mbar_init(&ctl->empty[1], g.cwarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 64; i += nt) ctl->hist[i] = 0;
  for (int i = tid; i < V * T_h; i += nt) {
    const int y = i / V, v = i - y * V;
    const ViewDev& vw = p.view[v];
    const int si2 = vw.map.ai * vw.w + vw.map.bi, base2 = vw.map.a0 * vw.w + vw.map.b0;
    const Lerp L = pisto_src_index(vw.scale_h, y, vw.map.ho, vw.same_h);
    rowtab[i] = make_float4(L.l0, L.l0, L.l1, L.l1);
    rowoff[i] = make_int2(4 * (base2 + L.i0 * si2), 4 * (base2 + L.i1 * si2));
  }
  for (int y = tid; y < T_h; y += nt) {
    unsigned int f = 0;
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      unsigned int fl = 0;  // first row of a strip: both source rows are loaded before the row loop
      bool strip_start = false;
      for (int q = 0; q < g.S; q++) strip_start |= (y == g.strip_y0[q]);
      if (!strip_start) {
        const Lerp L = pisto_src_index(vw.scale_h, y, vw.map.ho, vw.same_h);
        const Lerp P = pisto_src_index(vw.scale_h, y - 1, vw.map.ho, vw.same_h);
        fl = (P.i0 == L.i0 && P.i1 == L.i1) ? 0u : 1u;  // up-sampling / same size: the pair moves down by exactly one row
      }
      f |= fl << (2 * v);
    }
    rowflags[y] = f;
  }
  for (int i = tid; i < V * g.GX; i += nt) {
    const int v = i / g.GX, gx = i - v * g.GX;
    const ViewDev& vw = p.view[v];
    const int sj2 = 4 * (vw.map.aj * vw.w + vw.map.bj);
    const Lerp L0 = pisto_src_index(vw.scale_w, 2 * gx, vw.map.wo, vw.same_w);
    const Lerp L1 = pisto_src_index(vw.scale_w, 2 * gx + 1, vw.map.wo, vw.same_w);
    colA[i] = make_int4(L0.i0 * sj2, L0.i1 * sj2, L1.i0 * sj2, L1.i1 * sj2);
    colB[i] = make_float4(L0.l0, L1.l0, L0.l1, L1.l1);
  }

  // does tile n read its views at all?  (single-label tiles without any score export do not)
  auto tile_needs_views = [&](int n) -> bool {
    if (has_fused || need_low) return true;
    return pisto_tile_presence(p, n).single < 0;
  };
  // thread 0: fetch every view of tile n into staging buffer b
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
        if (k >= 2) mbar_wait_sleep(&ctl->empty[b], ((k >> 1) - 1) & 1);  // all compute warps are done with buffer b
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
  } else {
  // ===== compute warps
  const int grp = tid % g.GX, strip = min(tid / g.GX, g.S - 1);
  const bool worker = tid < g.GX * g.S;
  const int x = 2 * grp;
  const int ys = g.strip_y0[strip];
  const int ye = g.strip_y0[strip + 1];
  // 32x32 export: which of my two columns (if any) is a gather column, and its low-resolution column index
  int lowcol_mask = 0, lx0 = 0, lx1 = 0;
  if (need_low) {
    if (x % p.low_fw == p.low_fw / 2) { lowcol_mask |= 1; lx0 = x / p.low_fw; }
    if ((x + 1) % p.low_fw == p.low_fw / 2) { lowcol_mask |= 2; lx1 = (x + 1) / p.low_fw; }
  }
  const int low_first = need_low ? ((ys + p.low_fh - 1 - p.low_fh / 2) / p.low_fh) : 0;  // first low row at or below ys
  const uint32_t rowtab_s = smem_u32(rowtab), rowoff_s = smem_u32(rowoff), rowflags_s = smem_u32(rowflags);
  const uint32_t colA_t = smem_u32(colA) + 16u * grp, colB_t = smem_u32(colB) + 16u * grp;
  const uint32_t col_stride = 16u * g.GX;
  const int nt = ncomp;  // cooperative loops below run over the compute threads only

  for (int k = 0;; k++) {
    const int b = k & 1;
    mbar_wait(&ctl->full[b], (k >> 1) & 1);  // tile id published, views (if any) landed
    const int n = ctl->tile[b];
    if (n < 0) break;
    const TilePresence tp = pisto_tile_presence(p, n);
    const bool need_scores = tp.single < 0 || has_fused;
    // shared-memory byte address of view v's data in this tile's staging buffer (incl. the 0..3-float alignment shift)
    uint32_t vb[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
      vb[v] = smem_u32(vsm + b * g.buf_floats + g.view_off[v]) + sh;
    }

    u64 cnt_lo = 0, cnt_hi = 0;

    if (worker && ys < ye && need_scores) {
      u64 madd2[C];
#pragma unroll
      for (int c = 0; c < C; c++) { const float m = ((tp.bits >> c) & 1u) ? 0.f : -INFINITY; madd2[c] = pack2(m, m); }
      const bool fast_ok = p.dec.mask_mode != PISTO_MASK_MULTIPLY;
      u64 Ha[V][C], Hb[V][C];
      // horizontally interpolated values of one staged source row (byte address `row`) for my two columns
      auto load_h = [&](int v, uint32_t row, u64 (&H)[C]) {
        const int4 A = lds_i4(colA_t + v * col_stride);
        const ulonglong2 B = lds_u64x2(colB_t + v * col_stride);
        uint32_t a00 = row + A.x, a01 = row + A.y, a10 = row + A.z, a11 = row + A.w;
#pragma unroll
        for (int c = 0; c < C; c++) {
          const float p00 = lds_f32(a00), p01 = lds_f32(a01), p10 = lds_f32(a10), p11 = lds_f32(a11);
          H[c] = fma2(B.x, pack2(p00, p10), mul2(B.y, pack2(p01, p11)));
          if (c + 1 < C) { a00 += g.plane_bytes[v]; a01 += g.plane_bytes[v]; a10 += g.plane_bytes[v]; a11 += g.plane_bytes[v]; }
        }
      };
      const long long pix0 = ((long long)n * T_h + ys) * T_w + x;
      const uint8_t* bgp = has_bg ? p.bg + pix0 : nullptr;
      const uint8_t* gtp = do_conf ? p.gt + pix0 : nullptr;
      uint8_t* lbp = has_label ? p.label_out + pix0 : nullptr;
      float* fop = has_fused ? p.fused_out + (((long long)n * C) * T_h + ys) * T_w + x : nullptr;
      float* lowp = need_low ? p.lowres_out + ((long long)n * C * p.low_h + low_first) * p.low_w : nullptr;
      int low_next = need_low ? low_first * p.low_fh + p.low_fh / 2 : 0x7fffffff;
      uint32_t rt = rowtab_s + 16u * V * ys, ro_a = rowoff_s + 8u * V * ys, fl_a = rowflags_s + 4u * ys;
      // first row of the strip: load both bracketing source rows of every view (inside the loop rows only ever shift by one:
      // the host routes view sets with ho > T_h, whose source rows can jump, to the block kernel)
#pragma unroll
      for (int v = 0; v < V; v++) {
        const int2 ro = lds_i2(ro_a + 8u * v);
        load_h(v, vb[v] + ro.x, Ha[v]);
        load_h(v, vb[v] + ro.y, Hb[v]);
      }
      // byte masks are fetched two rows ahead of their use (HBM latency >> one row of arithmetic)
      unsigned int bg_c = 0, bg_n = 0, gt_c = 0, gt_n = 0;
      if (has_bg) {
        bg_c = __ldg(reinterpret_cast<const unsigned short*>(bgp));
        if (ys + 1 < ye) bg_n = __ldg(reinterpret_cast<const unsigned short*>(bgp + T_w));
        bgp += 2 * T_w;
      }
      if (do_conf) {
        gt_c = __ldg(reinterpret_cast<const unsigned short*>(gtp));
        if (ys + 1 < ye) gt_n = __ldg(reinterpret_cast<const unsigned short*>(gtp + T_w));
        gtp += 2 * T_w;
      }
#pragma unroll 1
      for (int yl = ys; yl < ye; yl++) {
        const unsigned int flags = lds_u32(fl_a);
        const unsigned int bg2 = bg_c, gt2 = gt_c;
        bg_c = bg_n; gt_c = gt_n;
        if (yl + 2 < ye) {
          if (has_bg) { bg_n = __ldg(reinterpret_cast<const unsigned short*>(bgp)); bgp += T_w; }
          if (do_conf) { gt_n = __ldg(reinterpret_cast<const unsigned short*>(gtp)); gtp += T_w; }
        }
        if (flags) {  // the bracketing source rows of at least one view moved down by one
#pragma unroll
          for (int v = 0; v < V; v++) {
            if (PAIRS && (v & 1)) continue;  // handled together with its twin
            if ((flags >> (2 * v)) & 3u) {
#pragma unroll
              for (int vv = v; vv < v + (PAIRS ? 2 : 1); vv++) {
                const int2 ro = lds_i2(ro_a + 8u * vv);
#pragma unroll
                for (int c = 0; c < C; c++) Ha[vv][c] = Hb[vv][c];
                load_h(vv, vb[vv] + ro.y, Hb[vv]);
              }
            }
          }
        }
        u64 acc[C];
        ulonglong2 w;
#pragma unroll
        for (int v = 0; v < V; v++) {
          if (!(PAIRS && (v & 1))) w = lds_u64x2(rt + 16u * v);
          u64 u[C];
#pragma unroll
          for (int c = 0; c < C; c++) u[c] = fma2(w.x, Ha[v][c], mul2(w.y, Hb[v][c]));
          if (PROB) pisto_softmax_fast2<C>(u);
#pragma unroll
          for (int c = 0; c < C; c++) acc[c] = (v == 0) ? u[c] : add2(acc[c], u[c]);
        }
        rt += 16u * V; ro_a += 8u * V; fl_a += 4u;
        // ---- per-pixel epilogue -----------------------------------------------------------------------------
        int lab0, lab1;
        if (tp.single >= 0) { lab0 = lab1 = tp.single; }
        else decide_pair<C>(acc, madd2, tp.bits, p.dec, fast_ok, lab0, lab1);
        if (do_conf) {
          const unsigned int g0 = gt2 & 0xffu, g1 = gt2 >> 8;
          if (g0 < (unsigned)C) { const unsigned int bn = g0 * C + lab0; const u64 inc = 1ull << (8 * (bn & 7)); if (bn < 8) cnt_lo += inc; else cnt_hi += inc; }
          if (g1 < (unsigned)C) { const unsigned int bn = g1 * C + lab1; const u64 inc = 1ull << (8 * (bn & 7)); if (bn < 8) cnt_lo += inc; else cnt_hi += inc; }
        }
        if (has_label) {
          unsigned int o0 = (unsigned)lab0, o1 = (unsigned)lab1;
          if (has_bg) {
            o0 = ((bg2 & 0xffu) == (unsigned)p.bg_match) ? (unsigned)p.bg_label : o0;
            o1 = ((bg2 >> 8) == (unsigned)p.bg_match) ? (unsigned)p.bg_label : o1;
          }
          *reinterpret_cast<unsigned short*>(lbp) = (unsigned short)(o0 | (o1 << 8));
          lbp += T_w;
        }
        if (has_fused) {
#pragma unroll
          for (int c = 0; c < C; c++) {
            float a0, a1;
            unpack2(acc[c], a0, a1);
            *reinterpret_cast<float2*>(fop + (long long)c * T_h * T_w) = make_float2(pisto_div_views(a0, p.dec), pisto_div_views(a1, p.dec));
          }
          fop += T_w;
        }
        if (need_low && yl == low_next) {
          low_next += p.low_fh;
          if (lowcol_mask) {
#pragma unroll
            for (int c = 0; c < C; c++) {
              float a0, a1;
              unpack2(acc[c], a0, a1);
              float* lo = lowp + c * (p.low_h * p.low_w);
              if (lowcol_mask & 1) lo[lx0] = pisto_div_views(a0, p.dec);
              if (lowcol_mask & 2) lo[lx1] = pisto_div_views(a1, p.dec);
            }
          }
          lowp += p.low_w;
        }
      }
    } else if (!need_scores) {
      // single-label tile (infer_pseudo_masks.py:71-73): constant label + background overwrite, 16 pixels per thread-step
      const long long tpx = (long long)T_h * T_w;
      const long long base = (long long)n * tpx;
      const unsigned int lab4 = 0x01010101u * (unsigned)tp.single, bgl4 = 0x01010101u * (unsigned)p.bg_label;
      // packed 8-bit confusion counters: at most 255 pixels per thread between flushes -> 16-byte path only while 16 * ceil(nvec / nt) <= 255
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
          if (gg < (unsigned)C) { const unsigned int bn = gg * C + tp.single; const u64 inc = 1ull << (8 * (bn & 7)); if (bn < 8) cnt_lo += inc; else cnt_hi += inc; }
        }
        if (has_label) p.label_out[base + i] = (uint8_t)o;
      }
      if (need_low) {
        // the 32x32 logits are exported before the shortcut (infer_pseudo_masks.py:126): evaluate the gather points only
        const int npt = p.low_h * p.low_w;
        for (int i = tid; i < npt; i += nt) {
          const int ly = i / p.low_w, lx = i - ly * p.low_w;
          const int yy = ly * p.low_fh + p.low_fh / 2, xx = lx * p.low_fw + p.low_fw / 2;
          float a[C];
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float4 er = rowtab[yy * V + v];
            const int2 ro = rowoff[yy * V + v];
            const int4 A = colA[v * g.GX + (xx >> 1)];
            const float4 B = colB[v * g.GX + (xx >> 1)];
            const int oa = (xx & 1) ? A.z : A.x, ob = (xx & 1) ? A.w : A.y;
            const float wl0 = (xx & 1) ? B.y : B.x, wl1 = (xx & 1) ? B.w : B.z;
            float u[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
              const uint32_t pl = vb[v] + c * g.plane_bytes[v];
              const float h0 = __fmaf_rn(wl0, lds_f32(pl + ro.x + oa), __fmul_rn(wl1, lds_f32(pl + ro.x + ob)));
              const float h1 = __fmaf_rn(wl0, lds_f32(pl + ro.y + oa), __fmul_rn(wl1, lds_f32(pl + ro.y + ob)));
              u[c] = __fmaf_rn(er.x, h0, __fmul_rn(er.z, h1));
            }
            if (PROB) pisto_softmax_fast<C>(u);
#pragma unroll
            for (int c = 0; c < C; c++) a[c] = (v == 0) ? u[c] : __fadd_rn(a[c], u[c]);
          }
#pragma unroll
          for (int c = 0; c < C; c++) p.lowres_out[((long long)n * C + c) * npt + i] = pisto_div_views(a[c], p.dec);
        }
      }
    }
    if (do_conf) {
      // every lane of every warp reaches this point: full-mask warp reductions are safe
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
  }  // compute warps
  if (do_conf) {
    __syncthreads();
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool make_geom(const pisto_ctx* h, const FuseParams& p, StreamGeom* g) {
  if (p.T_w % 2) return false;
  const int GX = p.T_w / 2;
  if (GX > kMaxThreads - 32) return false;  // one warp of the CTA is the producer
  int S = (kMaxThreads - 32) / GX;
  if (S > p.T_h) S = p.T_h;
  if (S > 16) S = 16;
  g->GX = GX; g->S = S;
  g->cwarps = (GX * S + 31) / 32;
  g->threads = g->cwarps * 32 + 32;
  // Row strips.  Threads of one warp should see the same "source rows moved" flags, otherwise the warp executes the
  // refill path of two strips.  Strip boundaries in thread space are multiples of GX; where one falls inside a warp the
  // two strips sharing that warp are started a multiple of the flag period apart (the flags of view v repeat every
  // T_h / gcd(T_h, ho_v) rows), and the remaining rows are spread evenly.
  int period = 1;
  for (int v = 0; v < p.V; v++) {
    int a = p.T_h, b = p.view[v].map.ho;
    while (b) { int t = a % b; a = b; b = t; }
    const int pv = p.view[v].same_h ? 1 : p.T_h / a;
    int x = period, y = pv;
    while (y) { int t = x % y; x = y; y = t; }
    period = period / x * pv;
  }
  int bnd[17];
  bool fixed[17];
  for (int q = 0; q <= S; q++) { bnd[q] = (int)((long long)p.T_h * q / S); fixed[q] = (q == 0 || q == S); }
  if (period > 1 && period * S <= p.T_h) {
    for (int q = 1; q < S; q++) {
      if ((q * GX) % 32 != 0) {  // strips q-1 and q share a warp
        int snapped = (bnd[q] - bnd[q - 1] + period / 2) / period * period;
        if (snapped < period) snapped = period;
        bnd[q] = bnd[q - 1] + snapped;
        fixed[q] = true;
      } else if (fixed[q - 1]) {
        // re-balance the free boundaries up to the next fixed one
        int nxt = q;
        while (!fixed[nxt] && ((nxt * GX) % 32 == 0) && nxt < S) nxt++;
        (void)nxt;
      }
    }
    // spread rows evenly between consecutive fixed boundaries
    int q0 = 0;
    for (int q = 1; q <= S; q++) {
      if (!fixed[q]) continue;
      for (int r = q0 + 1; r < q; r++) bnd[r] = bnd[q0] + (int)((long long)(bnd[q] - bnd[q0]) * (r - q0) / (q - q0));
      q0 = q;
    }
  }
  int rps = 0;
  for (int q = 0; q < S; q++) {
    if (bnd[q + 1] <= bnd[q]) return false;
    g->strip_y0[q] = bnd[q];
    if (bnd[q + 1] - bnd[q] > rps) rps = bnd[q + 1] - bnd[q];
  }
  g->strip_y0[S] = p.T_h;
  if (rps * 2 > 255) return false;  // packed 8-bit confusion counters
  g->rows_per_strip = rps;
  int fl = 0;
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    g->view_off[v] = fl;
    g->plane_bytes[v] = 4 * vw.h * vw.w;
    fl += (p.C * vw.h * vw.w + 3 /* alignment shift */ + 3 /* tail */ + 3) & ~3;
  }
  g->buf_floats = fl;
  int off = 0;
  g->ctl_off = off; off += (int)((sizeof(Ctl) + 127) & ~127u);
  g->flags_off = off; off += (4 * p.T_h + 15) & ~15;
  g->rowoff_off = off; off += 8 * p.V * p.T_h; off = (off + 15) & ~15;
  g->rowtab_off = off; off += 16 * p.V * p.T_h;
  g->cola_off = off; off += 16 * p.V * GX;
  g->colb_off = off; off += 16 * p.V * GX;
  off = (off + 127) & ~127;
  g->views_off = off; off += 2 * 4 * fl;
  g->smem_bytes = off;
  return off <= h->smem_optin - 1024;
}

template <int C, int V, bool PROB, int F, bool PAIRS>
int launch_cv(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  StreamGeom g;
  if (!make_geom(h, p, &g)) return PISTO_OK;  // not launched: caller falls back
  auto kern = fuse_stream_kernel<C, V, PROB, F, PAIRS>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int sched_slot = 0;
  { const int rc = pisto_sched_acquire(h, st, &g.counter, &sched_slot); if (rc != PISTO_OK) return rc; }
  const int grid = p.N < h->sm_count ? p.N : h->sm_count;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  { const int rc = pisto_sched_release(h, st, sched_slot); if (rc != PISTO_OK) return rc; }
  *launched = true;
  return PISTO_OK;
}

}  // namespace

// feature mask of a parameter block (bit layout of template parameter F)
static inline int pisto_stream_flags(const FuseParams& p) {
  return (p.bg ? 1 : 0) | ((p.conf && p.gt) ? 2 : 0) | (p.fused_out ? 4 : 0) | ((p.lowres_out && p.low_fh > 0) ? 8 : 0) | (p.label_out ? 16 : 0);
}

// views 2k / 2k+1 share their row geometry (a scale and its flipped twin)?
static inline bool pisto_stream_pairs(const FuseParams& p) {
  if (p.V < 2 || (p.V & 1)) return false;
  for (int v = 0; v < p.V; v += 2)
    if (p.view[v].map.ho != p.view[v + 1].map.ho) return false;
  return true;
}

// one (C, V) family: specialised feature masks for the BASELINE configs, run-time flags for everything else
template <int C, int V>
static int pisto_launch_stream_cv(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  for (int v = 0; v < p.V; v++)
    if (p.view[v].map.ho >= p.T_h) return PISTO_OK;  // same-size / down-sampling rows: the source-row pair does not move by exactly one -> block kernel
  if (p.fuse_mode == PISTO_FUSE_PROB_MEAN) {
    switch (pisto_stream_flags(p)) {
      case 17: return launch_cv<C, V, true, 17, false>(h, p, st, launched);  // bg + labels
      case 16: return launch_cv<C, V, true, 16, false>(h, p, st, launched);  // labels
      default: return launch_cv<C, V, true, -1, false>(h, p, st, launched);
    }
  }
  if ((V % 2 == 0) && pisto_stream_pairs(p)) {
    switch (pisto_stream_flags(p)) {
      case 25: return launch_cv<C, V, false, 25, (V % 2 == 0)>(h, p, st, launched);
      case 19: return launch_cv<C, V, false, 19, (V % 2 == 0)>(h, p, st, launched);
      case 18: return launch_cv<C, V, false, 18, (V % 2 == 0)>(h, p, st, launched);
      default: return launch_cv<C, V, false, -1, (V % 2 == 0)>(h, p, st, launched);
    }
  }
  switch (pisto_stream_flags(p)) {
    case 25: return launch_cv<C, V, false, 25, false>(h, p, st, launched);  // bg + labels + 32x32        (config 2)
    case 19: return launch_cv<C, V, false, 19, false>(h, p, st, launched);  // bg + gt/conf + labels      (config 1)
    case 18: return launch_cv<C, V, false, 18, false>(h, p, st, launched);  // gt/conf + labels           (config 3, mIoUMask.forward)
    default: return launch_cv<C, V, false, -1, false>(h, p, st, launched);
  }
}
