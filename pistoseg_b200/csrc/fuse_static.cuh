// Shape-specialised variant of the filtered fusion kernel (fuse_filter.cuh) for the WSSS4LUAD / BCSS production shape:
// 224x224 tiles, stride-8 views of the scales {0.75, 1, 1.25} (21 / 28 / 35 px, each plain or with its hflip twin) or the single
// 28 px view of BASELINE config 1, 32x32 logit export by the [3::7] gather (infer_pseudo_masks.py:118-154, loss.py:55-67).
//
// The algorithm, the error bound and every output are those of fuse_filter.cuh.  What changes is that the geometry is a
// compile-time constant, so everything the generic kernel reads from shared-memory tables at run time is folded into the code:
//
//   row loop    224 = 7 strips x 32 rows, and 32 output rows cover exactly h/7 = 3 / 4 / 5 source rows of the three scale groups,
//               so the schedule "which group's source-row pair moves down at which row" and the vertical weights are the SAME
//               for every strip.  The 32 rows are fully unrolled: no row table, no flag tests, no branches; the vertical
//               weight is an immediate operand of the packed fma (FFMA2 Rd, Ra, imm, Rc broadcasts a scalar immediate), map
//               rows are addressed [base + imm].  The difference maps carry one replicated pad row above and below (like the
//               two pad columns), so the top / bottom clamp of the bilinear index needs no special case: the two taps are
//               equal there, Dh = 0, and the result is the clamped sample exactly.
//   32x32       unit of work = (class, 4 low-resolution rows).  Source-row indices and vertical weights of a low-resolution
//               row are compile-time constants, every source row a unit needs is interpolated horizontally ONCE (the generic
//               kernel does it once per low-resolution row and class-plane address), taps are addressed [lane base + imm].
//               The operation order per sample is unchanged (horizontal lerp, vertical lerp, sum in view order), so the
//               export is still the reference's bit for bit.
//
// The weights are dyadic (h/224 = 3/32, 1/8, 5/32), so the constexpr tables equal pisto_src_index's float arithmetic exactly;
// the host nevertheless re-derives every table entry with pisto_src_index before the first launch and refuses the kernel (the
// caller falls back to the generic one) on any mismatch.
#pragma once
#include <type_traits>

#include "fuse_filter.cuh"

namespace {

constexpr int kST = 224;    // tile side
constexpr int kSGX = 56;    // threads per output row (4 columns each)
constexpr int kSS = 7;      // strips
constexpr int kSR = 32;     // rows per strip
constexpr int kSLow = 32;   // low-resolution side of the logit export
constexpr int kSCThreads = 416;  // 13 compute warps (392 workers)
constexpr int kSNarrowThreads = 832;  // 2 columns per thread: 25 compute warps (784 workers) + the producer warp

// compile-time loop with the index as a constant expression
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

__host__ __device__ constexpr int st_h(int G, int g) { return G == 1 ? 28 : (g == 0 ? 21 : (g == 1 ? 28 : 35)); }
__host__ __device__ constexpr int st_floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
// output row r of a strip (0..31): unclamped lower source row relative to the strip's first one, in [-1, h/7 - 1], and the
// weight of the upper row; src = (y + 0.5) h / 224 - 0.5 = ((2y + 1) h - 224) / 448 and 448 = 7 * 64, h = 7 * (h / 7)
__host__ __device__ constexpr int st_i0rel(int h, int r) { return st_floordiv((2 * r + 1) * h - kST, 2 * kST); }
__host__ __device__ constexpr float st_l1(int h, int r) { return (float)(((2 * r + 1) * h - kST - 2 * kST * st_i0rel(h, r)) / 7) * (1.f / 64.f); }
__host__ __device__ constexpr bool st_adv(int h, int r) { return r > 0 && st_i0rel(h, r) != st_i0rel(h, r - 1); }
// does the pair move at row rr of some / every block of U rows?
__host__ __device__ constexpr bool st_adv_some(int h, int rr, int U) { bool s = false; for (int r = rr; r < 32; r += U) s = s || st_adv(h, r); return s; }
__host__ __device__ constexpr bool st_adv_all(int h, int rr, int U) { bool s = true; for (int r = rr; r < 32; r += U) s = s && st_adv(h, r); return s; }
// low-resolution row ly (0..31) = output row 7 ly + 3: src = ((2 ly + 1) h - 32) / 64, clamped as pisto_src_index does
__host__ __device__ constexpr int lo_num(int h, int ly) { return (2 * ly + 1) * h - 32; }
__host__ __device__ constexpr int lo_i0(int h, int ly) { return lo_num(h, ly) < 0 ? 0 : (lo_num(h, ly) / 64 > h - 1 ? h - 1 : lo_num(h, ly) / 64); }
__host__ __device__ constexpr int lo_i1(int h, int ly) { return lo_i0(h, ly) + (lo_i0(h, ly) < h - 1 ? 1 : 0); }
__host__ __device__ constexpr float lo_l1(int h, int ly) {
  return lo_num(h, ly) < 0 ? 0.f : ((lo_num(h, ly) - 64 * lo_i0(h, ly)) >= 64 ? 1.f : (float)(lo_num(h, ly) - 64 * lo_i0(h, ly)) * (1.f / 64.f));
}
// byte offset of group g's difference maps inside the map area: (h + 2) x (h + 2) cells of KP floats each
__host__ __device__ constexpr int st_ybytes(int G, int g, int KP) {
  int off = 0;
  for (int q = 0; q < g; q++) off += (4 * KP * (st_h(G, q) + 2) * (st_h(G, q) + 2) + 15) & ~15;
  return off;
}

// shared-memory accesses at [register + immediate]
template <int OFF> __device__ __forceinline__ float lds_f32_o(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ float2 lds_f2_o(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ float4 lds_f4_o(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts_u8_o(uint32_t a, unsigned int v) { asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(OFF) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_u32_o(uint32_t a, unsigned int v) { asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(OFF) : "memory"); }

#ifndef PISTO_SU1
#define PISTO_SU1 32
#endif
#ifndef PISTO_SU2
#define PISTO_SU2 32
#endif
#ifndef PISTO_SNE
#define PISTO_SNE 4
#endif
constexpr int kSU1 = PISTO_SU1;  // rows unrolled in the row loop for one difference field (32: the whole strip, every constant an immediate)
constexpr int kSU2 = PISTO_SU2;  // ... for two / three fields (code size: the instruction caches hold ~32 KB)
#ifndef PISTO_SAUX
#define PISTO_SAUX 0
#endif
constexpr int kSAux = PISTO_SAUX;  // warps that only work on the 32x32 export
constexpr int kSNE = PISTO_SNE;  // export units per class (each 32 / kSNE low-resolution rows)
#ifndef PISTO_STATIC_DEFER
#define PISTO_STATIC_DEFER 1  // uncertain pixels are re-evaluated by the producer warp after the tile (0: recheck + exact pass inside the tile)
#endif
#ifndef PISTO_STATIC_W3
#define PISTO_STATIC_W3 2  // 1: row loops with one difference field (or one scale group and two fields) keep three-tap column weights in registers; 2: two fields too; 0 = select form
#endif
#ifndef PISTO_SMR
#define PISTO_SMR 1
#endif
constexpr int kSMR = PISTO_SMR;  // rows whose lead test is folded into one comparison (1, 2 or 4; must divide the unroll factors)
#ifndef PISTO_STATIC_MASKS_AHEAD
#define PISTO_STATIC_MASKS_AHEAD 1
#endif

struct StaticGeom {
  float wtab[kSR][4];                // -l0 of group g on row r of a strip (the unrolled-by-8 row loop reads it through uniform loads)
  unsigned int adv[4];               // bit r of [g]: group g's source-row pair moves down at row r of a strip
  int threads, cwarps, aux;
  int view_off[PISTO_MAX_VIEWS];     // float offset of each view inside one staging buffer (16-byte aligned)
  int vbase[PISTO_MAX_VIEWS], vcol[PISTO_MAX_VIEWS];  // byte address of de-augmented (i, j) inside a plane: vbase + 4 w i + j vcol
  int buf_floats;
  float tau_coef, tau_abs;
  int ctl_off, col4_off, col4i_off, w3_off, lowtab_off, ymap_off, queue_off, lab_off, views_off, smem_bytes;  // w3_off: [G][56] ColWeights3 (48 B) or -1
  int* counter;
  unsigned long long* stats;          // [4] device counters of the handle (pisto_filter_stats) or NULL
};

// horizontally interpolated values of the thread's 4 columns on one map row (3 adjacent source cells at `a`)
template <int K, int OFF>
__device__ __forceinline__ void static_load_h(uint32_t a, const float4 L1, unsigned int sel, u64 (&H)[K][2]) {
  const u64 one2 = pack2(1.f, 1.f);
  const u64 l1a = pack2(L1.x, L1.y), l1b = pack2(L1.z, L1.w);
  const u64 l0a = sub2(one2, l1a), l0b = sub2(one2, l1b);
  const bool s1 = (sel >> 1) & 1u, s2 = (sel >> 2) & 1u, s3 = (sel >> 3) & 1u;
  float y0[K], y1[K], y2[K];
  if constexpr (K == 1) {
    y0[0] = lds_f32_o<OFF>(a); y1[0] = lds_f32_o<OFF + 4>(a); y2[0] = lds_f32_o<OFF + 8>(a);
  } else if constexpr (K == 2) {
    const float2 v0 = lds_f2_o<OFF>(a), v1 = lds_f2_o<OFF + 8>(a), v2 = lds_f2_o<OFF + 16>(a);
    y0[0] = v0.x; y0[K - 1] = v0.y; y1[0] = v1.x; y1[K - 1] = v1.y; y2[0] = v2.x; y2[K - 1] = v2.y;
  } else {
    const float4 v0 = lds_f4_o<OFF>(a), v1 = lds_f4_o<OFF + 16>(a), v2 = lds_f4_o<OFF + 32>(a);
    y0[0] = v0.x; y0[1 % K] = v0.y; y0[K - 1] = v0.z; y1[0] = v1.x; y1[1 % K] = v1.y; y1[K - 1] = v1.z;
    y2[0] = v2.x; y2[1 % K] = v2.y; y2[K - 1] = v2.z;
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    const float a1 = s1 ? y1[k] : y0[k], b1 = s1 ? y2[k] : y1[k];
    const float a2 = s2 ? y1[k] : y0[k], b2 = s2 ? y2[k] : y1[k];
    const float a3 = s3 ? y1[k] : y0[k], b3 = s3 ? y2[k] : y1[k];
    H[k][0] = fma2(l0a, pack2(y0[k], a1), mul2(l1a, pack2(y1[k], b1)));
    H[k][1] = fma2(l0b, pack2(a2, a3), mul2(l1b, pack2(b2, b3)));
  }
}

// The same interpolation without per-column selects: the 4 columns of a thread take their two taps from the 3 adjacent cells y0, y1, y2,
// so every column is w_a y0 + w_b y1 + w_c y2 with (w_a, w_b, w_c) = (l0, l1, 0) or (0, l0, l1) -- three packed multiply-adds per column
// pair with the cell value broadcast to both lanes, instead of six selects and three predicate set-ups.  One of the three weights is an
// exact zero, the other two terms are rounded as before (a product, then one fma): the error bound of DESIGN.md 4.1 is unchanged.
// (Tiles with non-finite logits never reach the row loop, so 0 x Inf cannot occur.)
struct ColWeights3 { u64 a[2], b[2], c[2]; };  // [column pair]
__device__ __forceinline__ ColWeights3 static_weights3(const float4 L1, unsigned int sel) {
  const float l1[4] = {L1.x, L1.y, L1.z, L1.w};
  float wa[4], wb[4], wc[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const float l0 = __fsub_rn(1.f, l1[c]);
    const bool d = (sel >> c) & 1u;
    wa[c] = d ? 0.f : l0; wb[c] = d ? l0 : l1[c]; wc[c] = d ? l1[c] : 0.f;
  }
  ColWeights3 w;
  w.a[0] = pack2(wa[0], wa[1]); w.a[1] = pack2(wa[2], wa[3]);
  w.b[0] = pack2(wb[0], wb[1]); w.b[1] = pack2(wb[2], wb[3]);
  w.c[0] = pack2(wc[0], wc[1]); w.c[1] = pack2(wc[2], wc[3]);
  return w;
}
__device__ __forceinline__ ColWeights3 static_weights3_lds(uint32_t a) {  // one table entry: three 16-byte loads
  const int4 x = lds_i4(a), y = lds_i4(a + 16u), z = lds_i4(a + 32u);
  ColWeights3 w;
  w.a[0] = ((u64)(unsigned)x.y << 32) | (unsigned)x.x; w.a[1] = ((u64)(unsigned)x.w << 32) | (unsigned)x.z;
  w.b[0] = ((u64)(unsigned)y.y << 32) | (unsigned)y.x; w.b[1] = ((u64)(unsigned)y.w << 32) | (unsigned)y.z;
  w.c[0] = ((u64)(unsigned)z.y << 32) | (unsigned)z.x; w.c[1] = ((u64)(unsigned)z.w << 32) | (unsigned)z.z;
  return w;
}
template <int K, int OFF>
__device__ __forceinline__ void static_load_h3(uint32_t a, const ColWeights3& w, u64 (&H)[K][2]) {
  float y0[K], y1[K], y2[K];
  if constexpr (K == 1) {
    y0[0] = lds_f32_o<OFF>(a); y1[0] = lds_f32_o<OFF + 4>(a); y2[0] = lds_f32_o<OFF + 8>(a);
  } else if constexpr (K == 2) {
    const float2 v0 = lds_f2_o<OFF>(a), v1 = lds_f2_o<OFF + 8>(a), v2 = lds_f2_o<OFF + 16>(a);
    y0[0] = v0.x; y0[K - 1] = v0.y; y1[0] = v1.x; y1[K - 1] = v1.y; y2[0] = v2.x; y2[K - 1] = v2.y;
  } else {
    const float4 v0 = lds_f4_o<OFF>(a), v1 = lds_f4_o<OFF + 16>(a), v2 = lds_f4_o<OFF + 32>(a);
    y0[0] = v0.x; y0[1 % K] = v0.y; y0[K - 1] = v0.z; y1[0] = v1.x; y1[1 % K] = v1.y; y1[K - 1] = v1.z;
    y2[0] = v2.x; y2[1 % K] = v2.y; y2[K - 1] = v2.z;
  }
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int q = 0; q < 2; q++)
      H[k][q] = fma2(w.c[q], pack2(y2[k], y2[k]), fma2(w.b[q], pack2(y1[k], y1[k]), mul2(w.a[q], pack2(y0[k], y0[k]))));
}

// ---- the 32 rows of one strip: blocks of U rows, each block fully unrolled -------------------------------------------------
// U = 32: one block, every schedule decision and weight is a compile-time constant.  U = 8: four passes over the same code; a
// refill site is unconditional when the group moves at that row of every block (h = 28), tested against a uniform bit mask
// when it moves there in some blocks, and absent otherwise; the weights come from the parameter block (uniform loads, the
// packed fma takes a uniform-register operand).
// col4_t: address of the thread's float4 {l1 of its 4 columns} (group stride 16 * 56); col4i_t: address of its
// (bytes-per-float * j0 | sel << 16) word (group stride 4 * 56); lab_a: label-tile address of (first row of the strip, x)
// Returns the rows of the strip (bit r) on which the thread's 4-pixel group failed the lead test.  PACK: the label tile holds 2 bits
// per pixel (one byte per thread and row, row stride 56) instead of bytes.
// w3_t != 0: address of the thread's ColWeights3 entry of group 0 in the shared-memory table (group stride 48 * 56): loops that cannot
// keep the three-tap weights in registers (three difference fields) load them per refill -- three 16-byte loads instead of the
// select form's 16-byte load + six selects per field + predicate set-ups.
template <int G, int K, int U, bool PACK = false>
__device__ __forceinline__ unsigned int static_rows(const StaticGeom& g, uint32_t col4_t, uint32_t col4i_t, uint32_t ymap_s, uint32_t lab_a, int strip,
                                                    const unsigned int (&c4)[4], float tau, uint32_t w3_t = 0u) {
  constexpr int KP = K == 3 ? 4 : K;
  constexpr int NBLK = kSR / U;
  static_assert(U == 32 || U == 16 || U == 8 || U == 4, "row-loop unroll: 4, 8, 16 or 32");
  u64 Hb[G][K][2], Dh[G][K][2], base[K][2];
  uint32_t yb[G];  // address of the map row that holds Ha (cell of the thread's first column)
  unsigned int selm[G];
  constexpr bool HOIST = PISTO_STATIC_W3 && (K == 1 || (G == 1 && K <= 2) || (PISTO_STATIC_W3 >= 2 && K <= 2));  // the horizontal weights of the thread's columns stay in registers, in three-tap form (else: one 16-byte load per refill)
  float4 L1[G] = {};
  ColWeights3 W3[HOIST ? G : 1];
  auto l1_of = [&](int gi) { return lds_f4(col4_t + gi * 16u * kSGX); };
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
    const uint32_t u = lds_u32(col4i_t + gi * 4u * kSGX);
    // row (in the padded map) of Ha on output row 0 of the strip: strip * h/7 + i0rel(0) + 1
    yb[gi] = ymap_s + st_ybytes(G, gi, KP) + KP * (u & 0xffffu) + (uint32_t)(strip * (h / 7) + st_i0rel(h, 0) + 1) * RS;
    selm[gi] = u >> 16;
    L1[gi] = lds_f4(col4_t + gi * 16u * kSGX);
    u64 Ha[K][2];
    if constexpr (HOIST) {
      W3[gi] = static_weights3(L1[gi], selm[gi]);
      static_load_h3<K, 0>(yb[gi], W3[gi], Ha);
      static_load_h3<K, RS>(yb[gi], W3[gi], Hb[gi]);
    } else if (w3_t) {
      const ColWeights3 w = static_weights3_lds(w3_t + gi * 48u * kSGX);
      static_load_h3<K, 0>(yb[gi], w, Ha);
      static_load_h3<K, RS>(yb[gi], w, Hb[gi]);
    } else {
      static_load_h<K, 0>(yb[gi], L1[gi], selm[gi], Ha);
      static_load_h<K, RS>(yb[gi], L1[gi], selm[gi], Hb[gi]);
    }
    L1[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < 2; q++) Dh[gi][k][q] = sub2(Hb[gi][k][q], Ha[k][q]);
  });
  auto rebase = [&]() {
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        u64 s = Hb[0][k][q];
#pragma unroll
        for (int gi = 1; gi < G; gi++) s = add2(s, Hb[gi][k][q]);
        base[k][q] = s;
      }
  };
  rebase();
  unsigned int unc_rows = 0;  // bit r: row r of the strip failed the lead test
#pragma unroll 1
  for (int blk = 0; blk < NBLK; blk++) {
    unsigned int unc_blk = 0;
    float mnrun = 0.f;
    unsigned int advb[G];
#pragma unroll
    for (int gi = 0; gi < G; gi++) advb[gi] = U == 32 ? 0u : g.adv[gi] >> (blk * U);
    static_for<0, U>([&](auto RI) {
      constexpr int rr = decltype(RI)::value;
      bool moved = false;
      static_for<0, G>([&](auto GI) {
        constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
        constexpr bool some = st_adv_some(h, rr, U), all = st_adv_all(h, rr, U);
        if constexpr (some) {
          bool now = true;
          if constexpr (!all) now = (advb[gi] >> rr) & 1u;
          if (now) {
            yb[gi] += RS;
            u64 Hn[K][2];
            if constexpr (HOIST) static_load_h3<K, RS>(yb[gi], W3[gi], Hn);
            else if (w3_t) static_load_h3<K, RS>(yb[gi], static_weights3_lds(w3_t + gi * 48u * kSGX), Hn);
            else static_load_h<K, RS>(yb[gi], l1_of(gi), selm[gi], Hn);
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
              for (int q = 0; q < 2; q++) { Dh[gi][k][q] = sub2(Hn[k][q], Hb[gi][k][q]); Hb[gi][k][q] = Hn[k][q]; }
            moved = true;
          }
        }
      });
      if (moved) rebase();
      float w[G];
      if constexpr (U == 32) {
        static_for<0, G>([&](auto GI) { constexpr int gi = decltype(GI)::value; w[gi] = st_l1(st_h(G, gi), rr) - 1.f; });  // -l0
      } else {
        const float4 t = *reinterpret_cast<const float4*>(g.wtab[blk * U + rr]);
        if (G > 2) w[G > 2 ? 2 : 0] = t.z;
        if (G > 1) w[G > 1 ? 1 : 0] = t.y;
        w[0] = t.x;
      }
      u64 acc[K][2];
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
          u64 a = base[k][q];
#pragma unroll
          for (int gi = 0; gi < G; gi++) a = fma2(pack2(w[gi], w[gi]), Dh[gi][k][q], a);
          acc[k][q] = a;
        }
      // the lead test of kSMR consecutive rows can be folded into one comparison (a failure flags all of them, the per-pixel recheck
      // sorts it out); measured slower for kSMR = 4 / 8 (-6 % / -15 %: the rechecks sit on the critical path of the barrier), hence 1
      unsigned int lab4;
      const float mn = labels_from_diffs_min<K, 2>(acc, c4, lab4);
      mnrun = (rr % kSMR == 0) ? mn : fminf(mnrun, mn);
      if constexpr (rr % kSMR == kSMR - 1) { if (!(mnrun > tau)) unc_blk |= ((1u << kSMR) - 1u) << (rr - (kSMR - 1)); }
      if constexpr (PACK) sts_u8_o<rr * kSGX>(lab_a, (lab4 * 0x01100440u) >> 24);  // 2 bits per pixel, fields in the order (0, 2, 1, 3)
      else sts_u32_o<rr * kST>(lab_a, lab4);
    });
    lab_a += U * (PACK ? kSGX : kST);
    unc_rows |= unc_blk << (blk * U);
  }
  return unc_rows;
}

// ---- 2 columns per thread (NP = 1): half the registers per thread, twice the threads per row (112), so that an SM holds 26 warps
// instead of 16 -- the row loop is bound by dependent-instruction latency, its throughput follows the number of resident warps.
// col2_t: address of the thread's float2 {l1 of its 2 columns} (group stride 8 * 112); col2i_t: its (4 * j0 | sel << 16) word.
template <int K, int OFF>
__device__ __forceinline__ void static_load_h1(uint32_t a, const float2 L1, unsigned int sel, u64 (&H)[K]) {
  const u64 l1a = pack2(L1.x, L1.y);
  const u64 l0a = sub2(pack2(1.f, 1.f), l1a);
  const bool s1 = (sel >> 1) & 1u;
  float y0[K], y1[K], y2[K];
  if constexpr (K == 1) {
    y0[0] = lds_f32_o<OFF>(a); y1[0] = lds_f32_o<OFF + 4>(a); y2[0] = lds_f32_o<OFF + 8>(a);
  } else if constexpr (K == 2) {
    const float2 v0 = lds_f2_o<OFF>(a), v1 = lds_f2_o<OFF + 8>(a), v2 = lds_f2_o<OFF + 16>(a);
    y0[0] = v0.x; y0[K - 1] = v0.y; y1[0] = v1.x; y1[K - 1] = v1.y; y2[0] = v2.x; y2[K - 1] = v2.y;
  } else {
    const float4 v0 = lds_f4_o<OFF>(a), v1 = lds_f4_o<OFF + 16>(a), v2 = lds_f4_o<OFF + 32>(a);
    y0[0] = v0.x; y0[1 % K] = v0.y; y0[K - 1] = v0.z; y1[0] = v1.x; y1[1 % K] = v1.y; y1[K - 1] = v1.z;
    y2[0] = v2.x; y2[1 % K] = v2.y; y2[K - 1] = v2.z;
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    const float a1 = s1 ? y1[k] : y0[k], b1 = s1 ? y2[k] : y1[k];
    H[k] = fma2(l0a, pack2(y0[k], a1), mul2(l1a, pack2(y1[k], b1)));
  }
}

template <int G, int K, int U>
__device__ __forceinline__ unsigned int static_rows1(const StaticGeom& g, uint32_t col2_t, uint32_t col2i_t, uint32_t ymap_s, uint32_t lab_a, int strip,
                                                     const unsigned int (&c4)[4], float tau) {
  constexpr int KP = K == 3 ? 4 : K;
  constexpr int NBLK = kSR / U;
  constexpr int GX1 = 2 * kSGX;
  u64 Hb[G][K], Dh[G][K], base[K];
  uint32_t yb[G];
  unsigned int selm[G];
  float2 L1[G];
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
    const uint32_t u = lds_u32(col2i_t + gi * 4u * GX1);
    yb[gi] = ymap_s + st_ybytes(G, gi, KP) + KP * (u & 0xffffu) + (uint32_t)(strip * (h / 7) + st_i0rel(h, 0) + 1) * RS;
    selm[gi] = u >> 16;
    L1[gi] = lds_f2(col2_t + gi * 8u * GX1);
    u64 Ha[K];
    static_load_h1<K, 0>(yb[gi], L1[gi], selm[gi], Ha);
    static_load_h1<K, RS>(yb[gi], L1[gi], selm[gi], Hb[gi]);
#pragma unroll
    for (int k = 0; k < K; k++) Dh[gi][k] = sub2(Hb[gi][k], Ha[k]);
  });
  auto rebase = [&]() {
#pragma unroll
    for (int k = 0; k < K; k++) {
      u64 s = Hb[0][k];
#pragma unroll
      for (int gi = 1; gi < G; gi++) s = add2(s, Hb[gi][k]);
      base[k] = s;
    }
  };
  rebase();
  unsigned int unc_rows = 0;
#pragma unroll 1
  for (int blk = 0; blk < NBLK; blk++) {
    unsigned int unc_blk = 0;
    unsigned int advb[G];
#pragma unroll
    for (int gi = 0; gi < G; gi++) advb[gi] = U == 32 ? 0u : g.adv[gi] >> (blk * U);
    static_for<0, U>([&](auto RI) {
      constexpr int rr = decltype(RI)::value;
      bool moved = false;
      static_for<0, G>([&](auto GI) {
        constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
        constexpr bool some = st_adv_some(h, rr, U), all = st_adv_all(h, rr, U);
        if constexpr (some) {
          bool now = true;
          if constexpr (!all) now = (advb[gi] >> rr) & 1u;
          if (now) {
            yb[gi] += RS;
            u64 Hn[K];
            static_load_h1<K, RS>(yb[gi], L1[gi], selm[gi], Hn);
#pragma unroll
            for (int k = 0; k < K; k++) { Dh[gi][k] = sub2(Hn[k], Hb[gi][k]); Hb[gi][k] = Hn[k]; }
            moved = true;
          }
        }
      });
      if (moved) rebase();
      float w[G];
      if constexpr (U == 32) {
        static_for<0, G>([&](auto GI) { constexpr int gi = decltype(GI)::value; w[gi] = st_l1(st_h(G, gi), rr) - 1.f; });
      } else {
        const float4 t = *reinterpret_cast<const float4*>(g.wtab[blk * U + rr]);
        if (G > 2) w[G > 2 ? 2 : 0] = t.z;
        if (G > 1) w[G > 1 ? 1 : 0] = t.y;
        w[0] = t.x;
      }
      u64 acc[K][1];
#pragma unroll
      for (int k = 0; k < K; k++) {
        u64 a = base[k];
#pragma unroll
        for (int gi = 0; gi < G; gi++) a = fma2(pack2(w[gi], w[gi]), Dh[gi][k], a);
        acc[k][0] = a;
      }
      unsigned int lab4;
      if (!labels_from_diffs<K, 1>(acc, c4, tau, lab4)) unc_blk |= 1u << rr;
      asm volatile("st.shared.u16 [%0+%2], %1;" ::"r"(lab_a), "h"((unsigned short)lab4), "n"(rr * kST) : "memory");
    });
    lab_a += U * kST;
    unc_rows |= unc_blk << (blk * U);
  }
  return unc_rows;
}

// Rare path of the row loop: one output row whose 4-pixel (2-pixel) group failed the lead test is evaluated again from the difference
// maps, pixel by pixel, and only the pixels whose own lead is within tau go to the exact pass (any evaluation of the interpolated
// differences within the error bound of DESIGN.md 4.1 certifies the same order, so a pixel that passes here keeps the label the
// row loop gave it).  Returns the uncertain pixels of the group as a bit mask.
template <int G, int K>
__device__ __noinline__ unsigned int static_recheck4(uint32_t col4_t, uint32_t col4i_t, uint32_t ymap_s, int y, float tau) {
  constexpr int KP = K == 3 ? 4 : K;
  u64 acc[K][2];
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
    const uint32_t u = lds_u32(col4i_t + gi * 4u * kSGX);
    const uint32_t a0 = ymap_s + st_ybytes(G, gi, KP) + KP * (u & 0xffffu);
    const float4 L1 = lds_f4(col4_t + gi * 16u * kSGX);
    const Lerp Ly = pisto_src_index((float)h / (float)kST, y, h, false);
    u64 Ha[K][2], Hb[K][2];
    static_load_h<K, 0>(a0 + (uint32_t)(Ly.i0 + 1) * RS, L1, u >> 16, Ha);
    static_load_h<K, 0>(a0 + (uint32_t)(Ly.i1 + 1) * RS, L1, u >> 16, Hb);
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const u64 t = fma2(pack2(Ly.l0, Ly.l0), Ha[k][q], mul2(pack2(Ly.l1, Ly.l1), Hb[k][q]));
        acc[k][q] = gi == 0 ? t : add2(acc[k][q], t);
      }
  });
  const int cls[4] = {0, 1, 2, 3};
  return uncertain_mask<K, 2, 4>(acc, cls, tau);
}
template <int G, int K>
__device__ __noinline__ unsigned int static_recheck2(uint32_t col2_t, uint32_t col2i_t, uint32_t ymap_s, int y, float tau) {
  constexpr int KP = K == 3 ? 4 : K;
  constexpr int GX1 = 2 * kSGX;
  u64 acc[K][1];
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), RS = 4 * KP * (h + 2);
    const uint32_t u = lds_u32(col2i_t + gi * 4u * GX1);
    const uint32_t a0 = ymap_s + st_ybytes(G, gi, KP) + KP * (u & 0xffffu);
    const float2 L1 = lds_f2(col2_t + gi * 8u * GX1);
    const Lerp Ly = pisto_src_index((float)h / (float)kST, y, h, false);
    u64 Ha[K], Hb[K];
    static_load_h1<K, 0>(a0 + (uint32_t)(Ly.i0 + 1) * RS, L1, u >> 16, Ha);
    static_load_h1<K, 0>(a0 + (uint32_t)(Ly.i1 + 1) * RS, L1, u >> 16, Hb);
#pragma unroll
    for (int k = 0; k < K; k++) {
      const u64 t = fma2(pack2(Ly.l0, Ly.l0), Ha[k], mul2(pack2(Ly.l1, Ly.l1), Hb[k]));
      acc[k][0] = gi == 0 ? t : add2(acc[k][0], t);
    }
  });
  const int cls[4] = {0, 1, 2, 3};
  return uncertain_mask<K, 1, 4>(acc, cls, tau);
}
// flagged rows of a strip -> per-pixel queue entries (y << 16) | x
template <int G, int K, int NP>
__device__ __forceinline__ void static_push_rows(FCtl* ctl, uint32_t* queue, int cap, int b, uint32_t colw_t, uint32_t coli_t, uint32_t ymap_s, int ys, int x,
                                                 unsigned int rows, float tau) {
  while (rows) {
    const int y = ys + __ffs(rows) - 1;
    rows &= rows - 1;
    unsigned int m = NP == 2 ? static_recheck4<G, K>(colw_t, coli_t, ymap_s, y, tau) : static_recheck2<G, K>(colw_t, coli_t, ymap_s, y, tau);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const unsigned int idx = atomicAdd(&ctl->qcount[b], 1u);
      if (idx < (unsigned)cap) queue[idx] = ((unsigned)y << 16) | (unsigned)(x + j);
    }
  }
}

// ---- pre-pass: Y[g][k] = sum over the views of group g of (x[c_{k+1}] - x[c_0]) in the de-augmented frame, maps padded by one
// replicated row above / below and two replicated columns on the right; returns the thread's max |x| (NaN-propagating)
// tidr[g]: the thread's index rotated by the number of cells of the groups before g (mod nt), so that the partial last sweeps of the
// groups land on different threads and every thread handles ceil or floor of (cells / nt) cells in total (static_prepass_rot)
__host__ __device__ constexpr int st_rot_off(int G, int g, int nt) { int off = 0; for (int q = 0; q < g; q++) off = (off + st_h(G, q) * st_h(G, q)) % nt; return off; }
template <int G, int NT>
__device__ __forceinline__ void static_prepass_rot(int tid, int (&tidr)[G]) {
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, off = st_rot_off(G, gi, NT);
    tidr[gi] = tid - off + (tid < off ? NT : 0);
  });
}
template <int C, int G, int VPG, int K>
__device__ __forceinline__ float static_prepass(const StaticGeom& g, const uint32_t (&vb)[G * VPG], const int (&cls)[C], uint32_t ymap_s, const int (&tidr)[G], int nt) {
  constexpr int KP = K == 3 ? 4 : K;
  constexpr uint32_t ES = 4u * KP;
  float mxf = 0.f;
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), PL = 4 * h * h, W2 = h + 2;
    constexpr int va = gi * VPG, vc = gi * VPG + VPG - 1;
    const uint32_t ym = ymap_s + st_ybytes(G, gi, KP);
    const uint32_t basea = vb[va] + g.vbase[va] + cls[0] * PL, baseb = vb[vc] + g.vbase[vc] + cls[0] * PL;
    const int ca = g.vcol[va], cb = g.vcol[vc];
    int dq[K];
#pragma unroll
    for (int q = 0; q < K; q++) dq[q] = (cls[q + 1] - cls[0]) * PL;
#pragma unroll 2
    for (int idx = tidr[gi]; idx < h * h; idx += nt) {
      const int i = idx / h, j = idx - i * h;
      const uint32_t aa = basea + 4 * h * i + j * ca, ab = baseb + 4 * h * i + j * cb;
      const float x0a = lds_f32(aa), x0b = VPG == 2 ? lds_f32(ab) : 0.f;
      mxf = max_nan(mxf, fabsf(x0a));
      if (VPG == 2) mxf = max_nan(mxf, fabsf(x0b));
      const uint32_t ya = ym + ES * (uint32_t)((i + 1) * W2 + j);
      const bool lastc = j == h - 1, top = i == 0, bot = i == h - 1;
#pragma unroll
      for (int q = 0; q < K; q++) {
        const float xa = lds_f32(aa + dq[q]);
        mxf = max_nan(mxf, fabsf(xa));
        float yv = __fsub_rn(xa, x0a);
        if (VPG == 2) {
          const float xb = lds_f32(ab + dq[q]);
          mxf = max_nan(mxf, fabsf(xb));
          yv = __fadd_rn(yv, __fsub_rn(xb, x0b));
        }
        const uint32_t yq = ya + 4u * q;
        sts_f32(yq, yv);
        if (lastc | top | bot) {
          if (lastc) { sts_f32(yq + ES, yv); sts_f32(yq + 2u * ES, yv); }
          if (top | bot) {
            const uint32_t yp = top ? yq - ES * W2 : yq + ES * W2;
            sts_f32(yp, yv);
            if (lastc) { sts_f32(yp + ES, yv); sts_f32(yp + 2u * ES, yv); }
          }
        }
      }
    }
  });
  return mxf;
}

// a / V for the export (pisto_div_views with the IEEE division out of line: the unit code is replicated per row group)
__device__ __noinline__ float static_div_slow(float a, float fV) { return __fdiv_rn(a, fV); }
__device__ __forceinline__ float static_div_views(float a, float rcp_v, float fV) {
  // V = 1 and powers of two: rcp_v is exact, e = 0, r = q -- the same quotient pisto_div_views returns on its short cuts
  const float q = __fmul_rn(a, rcp_v);
  const float e = __fmaf_rn(-fV, q, a);
  const float r = __fmaf_rn(e, rcp_v, q);
  const float fa = fabsf(a);
  if (fa > 1e-30f && fa < 1e30f) return r;
  return static_div_slow(a, fV);
}

// ---- 32x32 logit export: unit (class c, low-resolution rows RPU E .. RPU E + RPU - 1), lane = low-resolution column.  sA / sB:
// shared-memory address of the lane's left / right tap in row 0 of class 0 of each view.  The views of a group (a scale and its
// flipped twin) share every constant, so they run through the same code (a two-trip loop); the sum starts from -0.f, the
// identity of IEEE addition, so that the first view needs no special case.
// first / last source row a row group touches, and the widest span over the row groups (the shared hlerp code loads that many)
__host__ __device__ constexpr int lo_imin(int h, int RPU, int E) { return lo_i0(h, RPU * E); }
__host__ __device__ constexpr int lo_span(int h, int RPU, int E) { return lo_i1(h, RPU * E + RPU - 1) - lo_i0(h, RPU * E) + 1; }
__host__ __device__ constexpr int lo_span_max(int h, int RPU) { int m = 0; for (int E = 0; E < 32 / RPU; E++) m = lo_span(h, RPU, E) > m ? lo_span(h, RPU, E) : m; return m; }

// One unit = (class c, row group e = low-resolution rows RPU e .. RPU e + RPU - 1).  The unit code used to exist once per row group
// (every offset and weight an immediate): 4 x 4.6 KB that the 13 warps of a CTA run side by side, and with the unrolled row loops
// the tile loop outgrew the SM's 32 KB instruction cache -- 56 % of the export's stall samples were instruction fetches
// (profiles/r02).  Now the horizontal interpolation of the source rows -- two thirds of the unit -- is ONE copy that addresses
// [lane base + row-group offset + immediate]; only the vertical interpolation (which source rows feed which low-resolution row, with
// which weights) is switched on e.  The operations and their order per sample are unchanged (bit-exact).
template <int C, int G, int VPG, int RPU>
__device__ __forceinline__ void static_export_unit(float* out_n, int c, int e, const uint32_t (&sA)[G * VPG], const uint32_t (&sB)[G * VPG],
                                                   const float (&lx0)[G], const float (&lx1)[G], float rcp_v, float fV) {
  constexpr int NE = 32 / RPU;
  float a[RPU];
#pragma unroll
  for (int r = 0; r < RPU; r++) a[r] = 0.f;
  static_for<0, G>([&](auto GI) {
    constexpr int gi = decltype(GI)::value, h = st_h(G, gi), PL = 4 * h * h, RB = 4 * h;
    constexpr int J = lo_span_max(h, RPU);  // rows loaded (a row group with a shorter span loads a row it does not use)
    int imin = 0;
    static_for<1, NE>([&](auto EI) { constexpr int E = decltype(EI)::value; imin = e == E ? lo_imin(h, RPU, E) : imin; });
    const uint32_t off = (uint32_t)(c * PL) + (uint32_t)imin * RB;
    if constexpr (VPG == 2) {
      // a scale and its flipped twin ride in the two lanes of the packed f32x2 operations (each lane is the same IEEE operation
      // as the scalar code); the two adds of the view-order sum stay scalar and sequential
      const uint32_t b0 = sA[2 * gi] + off, b1 = sB[2 * gi] + off, t0 = sA[2 * gi + 1] + off, t1 = sB[2 * gi + 1] + off;
      const u64 LX0 = pack2(lx0[gi], lx0[gi]), LX1 = pack2(lx1[gi], lx1[gi]);
      u64 Hr[J];
      static_for<0, J>([&](auto JI) {
        constexpr int j = decltype(JI)::value;
        Hr[j] = fma2(LX0, pack2(lds_f32_o<j * RB>(b0), lds_f32_o<j * RB>(t0)), mul2(LX1, pack2(lds_f32_o<j * RB>(b1), lds_f32_o<j * RB>(t1))));
      });
      static_for<0, NE>([&](auto EI) {
        constexpr int E = decltype(EI)::value;
        if (e == E) {
          static_for<0, RPU>([&](auto RI) {
            constexpr int r = decltype(RI)::value, ly = RPU * E + r;
            constexpr float l1 = lo_l1(h, ly), l0 = 1.f - l1;
            float ua, ub;
            unpack2(fma2(pack2(l0, l0), Hr[lo_i0(h, ly) - lo_imin(h, RPU, E)], mul2(pack2(l1, l1), Hr[lo_i1(h, ly) - lo_imin(h, RPU, E)])), ua, ub);
            a[r] = gi == 0 ? __fadd_rn(ua, ub) : __fadd_rn(__fadd_rn(a[r], ua), ub);
          });
        }
      });
    } else {
      const uint32_t b0 = sA[gi] + off, b1 = sB[gi] + off;
      float Hr[J];
      static_for<0, J>([&](auto JI) {
        constexpr int j = decltype(JI)::value;
        Hr[j] = __fmaf_rn(lx0[gi], lds_f32_o<j * RB>(b0), __fmul_rn(lx1[gi], lds_f32_o<j * RB>(b1)));
      });
      static_for<0, NE>([&](auto EI) {
        constexpr int E = decltype(EI)::value;
        if (e == E) {
          static_for<0, RPU>([&](auto RI) {
            constexpr int r = decltype(RI)::value, ly = RPU * E + r;
            constexpr float l1 = lo_l1(h, ly), l0 = 1.f - l1;
            const float u = __fmaf_rn(l0, Hr[lo_i0(h, ly) - lo_imin(h, RPU, E)], __fmul_rn(l1, Hr[lo_i1(h, ly) - lo_imin(h, RPU, E)]));
            a[r] = gi == 0 ? u : __fadd_rn(a[r], u);
          });
        }
      });
    }
  });
  float* outp = out_n + (c * kSLow + RPU * e) * kSLow + (threadIdx.x & 31);
  // a / V: the arithmetic of static_div_views, two rows per packed operation; one range test for the whole unit
  float lo = fabsf(a[0]), hi = lo;
#pragma unroll
  for (int r = 1; r < RPU; r++) { lo = fminf(lo, fabsf(a[r])); hi = fmaxf(hi, fabsf(a[r])); }
  if (RPU % 2 == 0 && lo > 1e-30f && hi < 1e30f) {
    const u64 R2 = pack2(rcp_v, rcp_v), NV2 = pack2(-fV, -fV);
#pragma unroll
    for (int r = 0; r < RPU; r += 2) {
      const u64 a2 = pack2(a[r], a[r + 1 < RPU ? r + 1 : r]);
      const u64 q2 = mul2(a2, R2);
      float o0, o1;
      unpack2(fma2(fma2(NV2, q2, a2), R2, q2), o0, o1);
      outp[r * kSLow] = o0;
      outp[(r + 1) * kSLow] = o1;
    }
  } else {
#pragma unroll
    for (int r = 0; r < RPU; r++) outp[r * kSLow] = static_div_slow(a[r], fV);  // (static: a[] must stay in registers)
  }
}

// background overwrite of 4 x 16 labels (infer_pseudo_masks.py:86-88: label[bg == match] = bg_label).  Byte masks that hold only
// 0 / 1 with match == 1 (what the reference's tissue masks are) take a multiply instead of the SWAR byte compare.
__device__ __forceinline__ void static_bg_overwrite4(uint4 (&lv)[4], const uint4 (&bgv)[4], unsigned int m4, unsigned int bgl4) {
  unsigned int any = 0;
#pragma unroll
  for (int u = 0; u < 4; u++) any |= bgv[u].x | bgv[u].y | bgv[u].z | bgv[u].w;
  if (m4 == 0x01010101u && (any & 0xfefefefeu) == 0u) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      unsigned int eq;
      eq = bgv[u].x * 0xffu; lv[u].x = (bgl4 & eq) | (lv[u].x & ~eq);
      eq = bgv[u].y * 0xffu; lv[u].y = (bgl4 & eq) | (lv[u].y & ~eq);
      eq = bgv[u].z * 0xffu; lv[u].z = (bgl4 & eq) | (lv[u].z & ~eq);
      eq = bgv[u].w * 0xffu; lv[u].w = (bgl4 & eq) | (lv[u].w & ~eq);
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      unsigned int eq;
      eq = __vcmpeq4(bgv[u].x, m4); lv[u].x = (bgl4 & eq) | (lv[u].x & ~eq);
      eq = __vcmpeq4(bgv[u].y, m4); lv[u].y = (bgl4 & eq) | (lv[u].y & ~eq);
      eq = __vcmpeq4(bgv[u].z, m4); lv[u].z = (bgl4 & eq) | (lv[u].z & ~eq);
      eq = __vcmpeq4(bgv[u].w, m4); lv[u].w = (bgl4 & eq) | (lv[u].w & ~eq);
    }
  }
}

// the per-tile head word of FCtl::head (producer lane)
template <int C, int V>
__device__ __forceinline__ uint4 static_tile_head(const FuseParams& p, int n, const TilePresence& tp) {
  unsigned int clsw = 0, P = 0, vshw = 0;
#pragma unroll
  for (int c = 0; c < C; c++)
    if ((tp.bits >> c) & 1u) { clsw |= (unsigned)c << (8 * P); P++; }
#pragma unroll
  for (int v = 0; v < V; v++) {
    const ViewDev& vw = p.view[v];
    vshw |= (((unsigned)reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) >> 2) & 3u) << (2 * v);
  }
  return make_uint4(tp.bits, (unsigned)tp.single, clsw, vshw | (P << 28));
}

template <int C, int G, int VPG, int F, int NB, int NP>
__device__ __forceinline__ void fuse_static_body(const FuseParams& p, const StaticGeom& g) {
  constexpr int GX = kST / (2 * NP);  // threads per output row
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int V = G * VPG;
  constexpr bool RT = F < 0;
  FCtl* ctl = reinterpret_cast<FCtl*>(smem_raw + g.ctl_off);
  float4* col4 = reinterpret_cast<float4*>(smem_raw + g.col4_off);        // [G][56] l1 of the thread's 4 columns
  uint32_t* col4i = reinterpret_cast<uint32_t*>(smem_raw + g.col4i_off);  // [G][56] (4 * j0 of column 0) | sel << 16
  uint2* lowtap = reinterpret_cast<uint2*>(smem_raw + g.lowtab_off);      // [32 lanes][V] byte offsets of the left / right tap of low-res column `lane`
  float2* lowwt = reinterpret_cast<float2*>(smem_raw + g.lowtab_off + 32 * 8 * V);  // [32][G] {l0, l1}
  float* ymap = reinterpret_cast<float*>(smem_raw + g.ymap_off);
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + g.queue_off);
  uint8_t* labsm = smem_raw + g.lab_off;
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool need_low = NB == 2 && (RT ? (p.lowres_out != nullptr && p.low_fh > 0) : ((F & 8) != 0));
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  constexpr int BINS = C * C;
  constexpr int UNITS = C * kSNE;  // export units per tile

  if (tid == 0) {
    mbar_init(&ctl->full[0], 1);
    mbar_init(&ctl->full[1], 1);
    mbar_init(&ctl->empty[0], g.cwarps + (need_low ? g.aux : 0));
    mbar_init(&ctl->empty[1], g.cwarps + (need_low ? g.aux : 0));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->maxbits[0] = ctl->maxbits[1] = 0u;
    ctl->lownext[0] = ctl->lownext[1] = 0u;
    ctl->qcount[0] = ctl->qcount[1] = 0u;
    for (int i = 0; i < 8; i++) ctl->qn[i] = 0u;
    ctl->tiles_done = 0u; ctl->fixed = 0u; ctl->ntiles = -1;
  }
  for (int i = tid; i < 64; i += nthreads) ctl->hist[i] = 0;
  for (int i = tid; i < G * GX; i += nthreads) {
    const int gi = i / GX, gx = i - gi * GX;
    const int h = gi == 0 ? st_h(G, 0) : (gi == 1 ? st_h(G, G > 1 ? 1 : 0) : st_h(G, G > 2 ? 2 : 0));
    const float sc = (float)h / (float)kST;
    Lerp L[4];
#pragma unroll
    for (int c = 0; c < 2 * NP; c++) L[c] = pisto_src_index(sc, 2 * NP * gx + c, h, false);
    if (NP == 2) col4[i] = make_float4(L[0].l1, L[1].l1, L[2 * NP - 2].l1, L[2 * NP - 1].l1);
    else reinterpret_cast<float2*>(col4)[i] = make_float2(L[0].l1, L[1].l1);
    unsigned int sel = 0;
#pragma unroll
    for (int c = 1; c < 2 * NP; c++) sel |= (unsigned)(L[c].i0 - L[0].i0) << c;  // 0 or 1 (checked on the host)
    col4i[i] = (unsigned)(4 * L[0].i0) | (sel << 16);
    if (NP == 2 && g.w3_off >= 0) {
      const ColWeights3 w = static_weights3(make_float4(L[0].l1, L[1].l1, L[2 * NP - 2].l1, L[2 * NP - 1].l1), sel);
      u64* dst = reinterpret_cast<u64*>(smem_raw + g.w3_off) + 6 * i;
      dst[0] = w.a[0]; dst[1] = w.a[1]; dst[2] = w.b[0]; dst[3] = w.b[1]; dst[4] = w.c[0]; dst[5] = w.c[1];
    }
  }

  if (need_low && tid < 32) {
    static_for<0, V>([&](auto VI) {
      constexpr int v = decltype(VI)::value, gi = v / VPG, h = st_h(G, gi);
      const Lerp L = pisto_src_index((float)h / (float)kST, 7 * tid + 3, h, false);
      lowtap[tid * V + v] = make_uint2((uint32_t)(g.vbase[v] + L.i0 * g.vcol[v]), (uint32_t)(g.vbase[v] + L.i1 * g.vcol[v]));
      lowwt[tid * G + gi] = make_float2(L.l0, L.l1);
    });
  }

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
      if (n == p.N - 1) {  // never read past the end of the caller's buffer: whole 16-byte units only, the tail by hand
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

  __syncthreads();

  const int ncomp = g.cwarps * 32;
  // Deferred exact pass (4 columns per thread): a row whose 4-pixel group fails the lead test is NOT re-evaluated by the compute warps.
  // The thread pushes (row, first column, the 4 labels it stored) to the tile's queue and goes on; the vector pass writes those labels
  // like any other.  A dedicated FIXER warp -- the CTA has two warp slots to spare -- waits until every compute warp has counted the
  // tile done, evaluates the queued pixels exactly from global memory (the tile's views were streamed microseconds ago: L2), one pixel
  // per lane, operation by operation as torch does, and patches the few labels that differ (respecting the background overwrite;
  // confusion counts through global atomics).  What this takes off the compute warps' critical path: the in-thread recheck (150-400
  // dependent instructions in ONE thread before the barrier), the CTA-wide exact pass and the barrier after it -- a build that never
  // flags measured +8 % on cfg 2 and +11 % on cfg 3.  Queues rotate through 8 slots of 128 groups; an overflowing queue (adversarial
  // inputs) makes the compute warps evaluate the whole tile exactly inside the tile, as before.
  constexpr bool DEFER = PISTO_STATIC_DEFER && NP == 2;
  constexpr int kQS = 8, kGQ = kFQueueCap / (2 * kQS);  // slots; queued groups per slot: entries[kGQ] + labels[kGQ]
  if (DEFER && tid >= ncomp + 32 * g.aux + 32) {
    // ===== fixer warp
    const int lane = tid & 31;
    const long long tile_px = (long long)kST * kST;
    volatile unsigned int* v_done = &ctl->tiles_done;
    volatile int* v_ntiles = &ctl->ntiles;
    for (int kk = 0;; kk++) {
      for (int it = 0;; it++) {  // wait for tile kk of this CTA to be complete, or for the end of the tile stream
        if (*v_done >= (unsigned)(kk + 1) * (unsigned)g.cwarps) break;
        const int nt_ = *v_ntiles;
        if (nt_ >= 0 && kk >= nt_) return;
        __nanosleep(it < 4 ? 100 : 400);
      }
      __threadfence_block();  // the tile's label stores (other warps) are ordered before the patches below
      const int s = kk & (kQS - 1);
      const unsigned int nq = ctl->qn[s];
      if (nq) {
        const int2 m = ctl->qmeta[s];
        const int tile = m.x;
        const unsigned int bits = (unsigned)m.y;
        const uint32_t* qb = queue + s * (2 * kGQ);
        const uint32_t* labs = qb + kGQ;
        for (int pi = lane; pi < 4 * (int)nq; pi += 32) {
          const uint32_t e = qb[pi >> 2];
          const int yy = (int)(e >> 16), xx = (int)(e & 0xffffu) + (pi & 3);
          const unsigned int old = (labs[pi >> 2] >> (8 * (pi & 3))) & 0xffu;
          float a[C];
#pragma unroll
          for (int c = 0; c < C; c++) a[c] = 0.f;
#pragma unroll 1  // (rolled: this warp runs beside the compute warps, its code must not compete for their instruction cache)
          for (int v = 0; v < V; v++) {
            const ViewDev& vw = p.view[v];
            const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, false);
            const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, false);
            const int r0 = g.vbase[v] + Ly.i0 * 4 * vw.w, r1 = g.vbase[v] + Ly.i1 * 4 * vw.w;
            const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
            const float* gsrc = vw.logits + (long long)tile * vw.tile_stride;
#pragma unroll
            for (int c = 0; c < C; c++) {
              const int pl = c * 4 * vw.h * vw.w;
              const float x00 = __ldg(gsrc + ((r0 + pl + c0) >> 2)), x01 = __ldg(gsrc + ((r0 + pl + c1) >> 2));
              const float x10 = __ldg(gsrc + ((r1 + pl + c0) >> 2)), x11 = __ldg(gsrc + ((r1 + pl + c1) >> 2));
              const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
              const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
              const float u = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
              a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
            }
          }
          const unsigned int lab = (unsigned)pisto_decide<C>(a, bits, p.dec, false, nullptr);
          if (lab != old) {
            const long long px = (long long)tile * tile_px + yy * kST + xx;
            if (do_conf) {
              const unsigned int gg = p.gt[px];
              if (gg < (unsigned)C) { atomicAdd(&p.conf[gg * C + old], ~0ull); atomicAdd(&p.conf[gg * C + lab], 1ull); }
            }
            if (has_label && !(has_bg && p.bg[px] == (uint8_t)p.bg_match)) p.label_out[px] = (uint8_t)lab;
          }
        }
        if (lane == 0 && g.stats) atomicAdd(&g.stats[2], 4ull * nq);
      }
      __syncwarp();
      if (lane == 0) {
        ctl->qn[s] = 0u;
        __threadfence_block();
        *reinterpret_cast<volatile unsigned int*>(&ctl->fixed) = (unsigned)(kk + 1);
      }
    }
  }
  if (tid >= ncomp + 32 * g.aux) {
    // ===== producer warp
    if (tid == ncomp + 32 * g.aux) {
      const long long tile_px = (long long)kST * kST;
      int next = atomicAdd(g.counter, 1);
      TilePresence next_tp; next_tp.bits = 0u; next_tp.single = -1;
      if (next < p.N) next_tp = pisto_tile_presence(p, next);
      bool next_views = next < p.N && (need_low || next_tp.single < 0);
      for (int k = 0;; k++) {
        const int b = NB == 2 ? (k & 1) : 0;
        if (k >= NB) mbar_wait_sleep(&ctl->empty[b], NB == 2 ? (((k >> 1) - 1) & 1) : ((k - 1) & 1));
        const int tile = next < p.N ? next : -1;
        if (DEFER && tile >= 0 && k >= kQS) {  // queue slot k & 7 was last used by tile k - 8: it must have been emptied (it always has)
          volatile unsigned int* v_fixed = &ctl->fixed;
          while (*v_fixed + (unsigned)kQS < (unsigned)k + 1u) __nanosleep(200);
        }
        ctl->tile[b] = tile;
        ctl->pres_bits[b] = next_tp.bits;
        ctl->pres_single[b] = next_tp.single;
        ctl->lownext[b] = 0u;
        if (tile >= 0) ctl->head[b] = static_tile_head<C, V>(p, tile, next_tp);
        if (tile < 0 && DEFER) { *reinterpret_cast<volatile int*>(&ctl->ntiles) = k; }
        if (tile >= 0 && next_views) issue_tile(tile, b);
        else mbar_arrive(&ctl->full[b]);
        if (tile < 0) break;
        if (has_bg && ((uintptr_t)p.bg & 15) == 0) bulk_prefetch_l2(p.bg + tile * tile_px, (uint32_t)tile_px);
        if (do_conf && ((uintptr_t)p.gt & 15) == 0) bulk_prefetch_l2(p.gt + tile * tile_px, (uint32_t)tile_px);
        next = atomicAdd(g.counter, 1);
        if (next < p.N) next_tp = pisto_tile_presence(p, next);
        next_views = next < p.N && (need_low || next_tp.single < 0);
      }
    }
    return;
  }

  // ---- 32x32 export: units claimed from a shared counter; lane = low-resolution column, its taps / weights come from a table
  auto export_units = [&](int n, int b, const uint32_t (&vb)[V]) {
    const int lane = tid & 31;
    uint32_t sA[V], sB[V];
    float lx0[G], lx1[G];
#pragma unroll
    for (int v = 0; v < V; v++) { const uint2 t = lowtap[lane * V + v]; sA[v] = vb[v] + t.x; sB[v] = vb[v] + t.y; }
#pragma unroll
    for (int q = 0; q < G; q++) { const float2 t = lowwt[lane * G + q]; lx0[q] = t.x; lx1[q] = t.y; }
    float* out_n = p.lowres_out + (long long)n * C * kSLow * kSLow;
    // units are claimed from a shared counter (export warps start early, compute warps join when their tile work is done)
#ifdef PISTO_EXPORT_STATIC_UNITS  // warp w takes unit w: no gain measured over the shared counter (profiles/r02): off
    const bool claim = g.aux != 0;
#else
    const bool claim = true;
#endif
    int u = claim ? 0 : (tid >> 5);
    for (;;) {
      if (claim) {
        if (lane == 0) u = (int)atomicAdd(&ctl->lownext[b], 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
      }
      if (u >= UNITS) break;
      const int c = u / kSNE, e = u - c * kSNE;
      static_export_unit<C, G, VPG, kSLow / kSNE>(out_n, c, e, sA, sB, lx0, lx1, p.dec.rcp_v, p.dec.fV);
      u += g.cwarps;
    }
  };

  // ===== compute warps (tid < ncomp) and export warps: one tile loop; the export warps skip the label work and go straight to
  // the 32x32 units, the compute warps join them when their own work on the tile is done (the unit code exists once)
  const bool is_export = tid >= ncomp;
  if (is_export && !need_low) return;
  const int grp = tid % GX, strip = min(tid / GX, kSS - 1);
  const bool worker = tid < GX * kSS;
  const int x = 2 * NP * grp;
  const uint32_t ymap_s = smem_u32(ymap);
  const uint32_t col4_t = smem_u32(col4) + 8u * NP * grp, col4i_t = smem_u32(col4i) + 4u * grp;
  const uint32_t w3_t = (NP == 2 && g.w3_off >= 0) ? smem_u32(smem_raw + g.w3_off) + 48u * grp : 0u;
  const uint32_t lab_s = smem_u32(labsm);
  const int nt = ncomp;
  constexpr long long tpx = (long long)kST * kST;
  constexpr int NT = 32 * ((GX * kSS + 31) / 32);  // compute threads (= ncomp)


  for (int k = 0;; k++) {
    const int b = k & 1;
    const int sb = NB == 2 ? b : 0;
    const int qs = k & (kQS - 1);  // queue slot of the deferred exact pass
    mbar_wait(&ctl->full[sb], NB == 2 ? ((k >> 1) & 1) : (k & 1));
    const int n = ctl->tile[sb];
    if (n < 0) break;
    const int4 hd = lds_i4(smem_u32(&ctl->head[sb]));
    const int ctl_single = hd.y;
    uint32_t vb[V];
#pragma unroll
    for (int v = 0; v < V; v++) vb[v] = smem_u32(vsm) + 4u * (uint32_t)(sb * g.buf_floats + g.view_off[v]) + ((((unsigned)hd.w >> (2 * v)) & 3u) << 2);
    // (Measured and dropped: mbarrier-based phase barriers at which a waiting warp works off export units.  The arrive / try_wait
    // barrier itself cost 8 % against bar.sync and the filled waits returned less than that.)
    auto phase_sync = [&](int) { bar_sync(1, ncomp); };
    if (!is_export) {
    TilePresence tp;
    tp.bits = (unsigned)hd.x; tp.single = hd.y;
    const bool multi = tp.single < 0;
    int cls[C];
#pragma unroll
    for (int c = 0; c < C; c++) cls[c] = (int)(((unsigned)hd.z >> (8 * c)) & 0xffu);
    const int P = (int)((unsigned)hd.w >> 28);
    if (DEFER && tid == 0) ctl->qmeta[qs] = make_int2(n, (int)tp.bits);  // whose queue slot qs is (read by the fixer warp after the tile)

#ifdef PISTO_X_SKIP_PREPASS
    if (false) {
#else
    if (multi && P >= 2) {
#endif
      float mxf;
      int tidr[G];
#ifdef PISTO_PREPASS_ROT  // balanced sweeps measured 2-4 % slower on B200 (profiles/r02): off
      static_prepass_rot<G, NT>(tid, tidr);
#else
#pragma unroll
      for (int q = 0; q < G; q++) tidr[q] = tid;
#endif
      if (P == 2) mxf = static_prepass<C, G, VPG, 1>(g, vb, cls, ymap_s, tidr, nt);
      else if (P == 3) mxf = static_prepass<C, G, VPG, 2>(g, vb, cls, ymap_s, tidr, nt);
      else mxf = static_prepass<C, G, VPG, (C >= 4 ? 3 : 1)>(g, vb, cls, ymap_s, tidr, nt);
      const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));
      if ((tid & 31) == 0) atomicMax(&ctl->maxbits[b], mx);
    }
    phase_sync(0);  // difference maps + max visible; every thread has left the previous tile
    if (NB == 1 && (tid & 31) == 0) mbar_arrive(&ctl->empty[0]);
    if (!DEFER && tid == 0) ctl->qcount[b ^ 1] = 0u;  // the previous tile's queue has been read by everyone

    // The byte masks of the vector pass' first batch are requested as soon as the registers are free -- right away for single-label
    // tiles, after the row loop otherwise -- so that their latency is covered by the barrier and the exact pass.
    constexpr int UN = 4;
    const long long vbase_px = (long long)n * tpx;
    const bool vec_ok = ((((uintptr_t)p.bg | (uintptr_t)p.gt | (uintptr_t)p.label_out) & 15) == 0);
    const int nvec = vec_ok ? (int)(tpx / 16) : 0;
    // every sweep of the vector pass is requested at once (SWEEPS * UN 16-byte vectors per mask and thread): the row loop's registers
    // are free by then, and a request made one sweep ahead only had ~100 instructions to cover an L2 round trip
    constexpr int NV = (int)(tpx / 16), SWEEPS = (NV + UN * NT - 1) / (UN * NT), NM = PISTO_STATIC_MASKS_AHEAD ? SWEEPS * UN : UN;
    uint4 bgn[NM], gn[NM];
    auto request_masks = [&]() {
#pragma unroll
      for (int u = 0; u < NM; u++) {
        const int i = tid + u * NT;
        gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        bgn[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < nvec) {
          if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + vbase_px) + i);
          if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + vbase_px) + i);
        }
      }
    };
    if (!multi) request_masks();

    bool exact_all = false;
    if (multi) {
      float tau = 0.f;
      if (P >= 2) {
        const float A = __fmul_rn((float)V, __uint_as_float(ctl->maxbits[b]));
        tau = __fmaf_rn(A, g.tau_coef, g.tau_abs);
        if (!(A < 5e8f)) exact_all = true;
      } else {
        exact_all = true;
      }
      // one exact sample sum per class at (yy, xx), evaluated by the whole warp: lane l computes the bilinear sample of
      // (view l / C, class l % C), the sums are then formed in view order through shuffles (the reference's operation order)
      auto exact_warp = [&](int yy, int xx) {
        const int lane = tid & 31;
        float o = 0.f;
        if (lane < V * C) {
          const int v = lane / C, c = lane - v * C;
          const ViewDev& vw = p.view[v];
          const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, false);
          const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, false);
          const int pl = g.vbase[v] + c * 4 * vw.h * vw.w;
          const int r0 = pl + Ly.i0 * 4 * vw.w, r1 = pl + Ly.i1 * 4 * vw.w;
          const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
          float x00, x01, x10, x11;
          if (NB == 2) {
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
            const uint32_t base = smem_u32(vsm + sb * g.buf_floats + g.view_off[v]) + sh;
            x00 = lds_f32(base + r0 + c0); x01 = lds_f32(base + r0 + c1); x10 = lds_f32(base + r1 + c0); x11 = lds_f32(base + r1 + c1);
          } else {  // single staging buffer: already released, read L2 / HBM
            const float* gsrc = vw.logits + (long long)n * vw.tile_stride;
            x00 = __ldg(gsrc + ((r0 + c0) >> 2)); x01 = __ldg(gsrc + ((r0 + c1) >> 2));
            x10 = __ldg(gsrc + ((r1 + c0) >> 2)); x11 = __ldg(gsrc + ((r1 + c1) >> 2));
          }
          const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
          const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
          o = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
        }
        float a[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
          a[c] = __shfl_sync(0xffffffffu, o, c);
#pragma unroll
          for (int v = 1; v < V; v++) a[c] = __fadd_rn(a[c], __shfl_sync(0xffffffffu, o, (v * C + c) & 31));
        }
        const int lab = pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
        if (lane == 0) labsm[yy * kST + xx] = (uint8_t)lab;
      };
      unsigned int unc = 0;
#ifdef PISTO_X_SKIP_ROWS
      if (false) {
#else
      if (!exact_all && worker) {
#endif
        unsigned int c4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) c4[q] = 0x01010101u * (unsigned)cls[q < C ? q : 0];
        const uint32_t lab_a = lab_s + (uint32_t)(strip * kSR * kST + x);
        if constexpr (NP == 2) {
          if (P == 2) unc = static_rows<G, 1, kSU1>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (P == 3) unc = static_rows<G, 2, kSU2>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (C >= 4 && P == 4) unc = static_rows<G, (C >= 4 ? 3 : 1), kSU2>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau, w3_t);
        } else {
          if (P == 2) unc = static_rows1<G, 1, kSU1>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (P == 3) unc = static_rows1<G, 2, kSU2>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (C >= 4 && P == 4) unc = static_rows1<G, (C >= 4 ? 3 : 1), kSU2>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
        }
      }
      if (unc) {
        if constexpr (DEFER) {
          // flagged rows -> (row, first column) + the 4 labels the thread stored for that row; the producer warp takes it from there
          const uint32_t lab_a = lab_s + (uint32_t)(strip * kSR * kST + x);
          unsigned int idx = atomicAdd(&ctl->qn[qs], (unsigned)__popc(unc));
          while (unc) {
            const int r = __ffs(unc) - 1;
            unc &= unc - 1;
            if (idx < (unsigned)kGQ) {
              queue[qs * (2 * kGQ) + idx] = ((unsigned)(strip * kSR + r) << 16) | (unsigned)x;
              queue[qs * (2 * kGQ) + kGQ + idx] = lds_u32(lab_a + (uint32_t)(r * kST));
            }
            idx++;
          }
        } else {
          // in-tile variant: the thread re-evaluates a flagged row per pixel and queues only the pixels that still fail
          if (P == 2) static_push_rows<G, 1, NP>(ctl, queue, kFQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
          else if (P == 3) static_push_rows<G, 2, NP>(ctl, queue, kFQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
          else static_push_rows<G, (C >= 4 ? 3 : 1), NP>(ctl, queue, kFQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
        }
      }
      request_masks();
      phase_sync(1);  // every strip done: the queue is complete; everyone has read maxbits
      if (tid == 0) ctl->maxbits[b] = 0u;
      const unsigned int nq = DEFER ? ctl->qn[qs] : ctl->qcount[b];
      if (tid == 0 && g.stats) {  // data-dependence record: queued pixels, whole-tile fallbacks (pisto_filter_stats)
        atomicAdd(&g.stats[1], 1ull);
        if (!exact_all && nq <= (unsigned)(DEFER ? kGQ : kFQueueCap)) { if (!DEFER) atomicAdd(&g.stats[2], (unsigned long long)nq); } else atomicAdd(&g.stats[3], 1ull);
      }
      if (nq > (unsigned)(DEFER ? kGQ : kFQueueCap)) exact_all = true;  // overflow: redo the whole tile
      if constexpr (!DEFER) {
        // Pixels that failed the lead test are re-evaluated exactly -- operation by operation as torch does -- one pixel per warp at a
        // time, spread over all warps.
        if (!exact_all && nq) {
          for (int j = tid >> 5; j < (int)nq; j += g.cwarps) {
            const uint32_t e = queue[j];
            exact_warp((int)(e >> 16), (int)(e & 0xffffu));
          }
        }
        if (!exact_all && nq) phase_sync(2);  // label tile complete
      }
      if (exact_all) {  // non-finite / absurd magnitudes, empty presence vector: the whole tile follows the reference pixel by pixel
        for (int j = tid; j < kST * kST; j += nt) {
          const int yy = j / kST, xx = j - yy * kST;
          float a[C];
#pragma unroll
          for (int v = 0; v < V; v++) {
            const ViewDev& vw = p.view[v];
            const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, false);
            const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, false);
            const int r0 = g.vbase[v] + Ly.i0 * 4 * vw.w, r1 = g.vbase[v] + Ly.i1 * 4 * vw.w;
            const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
            const float* gsrc = vw.logits + (long long)n * vw.tile_stride;
#pragma unroll
            for (int c = 0; c < C; c++) {
              const int pl = c * 4 * vw.h * vw.w;
              float x00, x01, x10, x11;
              if (NB == 2) {
                x00 = lds_f32(vb[v] + r0 + pl + c0); x01 = lds_f32(vb[v] + r0 + pl + c1);
                x10 = lds_f32(vb[v] + r1 + pl + c0); x11 = lds_f32(vb[v] + r1 + pl + c1);
              } else {
                x00 = __ldg(gsrc + ((r0 + pl + c0) >> 2)); x01 = __ldg(gsrc + ((r0 + pl + c1) >> 2));
                x10 = __ldg(gsrc + ((r1 + pl + c0) >> 2)); x11 = __ldg(gsrc + ((r1 + pl + c1) >> 2));
              }
              const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
              const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
              const float u = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
              a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
            }
          }
          labsm[j] = (uint8_t)pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
        }
        phase_sync(2);
        if (DEFER && tid == 0) ctl->qn[qs] = 0u;  // the tile is exact as a whole: nothing left for the deferred pass
      }
    }

    // ---- vector pass: confusion, background overwrite, 16-byte label stores (single-label tiles: constant label)
#ifdef PISTO_X_SKIP_VECTOR
    if (false) {
#else
    {
#endif
      const long long base = vbase_px;
      const unsigned int labc = 0x01010101u * (unsigned)(multi ? 0 : tp.single), bgl4 = 0x01010101u * (unsigned)p.bg_label;
      const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match;
      unsigned int cnt32[BINS];
#pragma unroll
      for (int i = 0; i < BINS; i++) cnt32[i] = 0;
      // the sweep count is a compile-time constant (NV vectors, NT threads, UN per sweep): fully unrolled, the batch requested one
      // sweep ahead is renamed instead of copied, and every bound test but the last sweep's folds away
      if (nvec) {
        static_for<0, SWEEPS>([&](auto TI) {
          constexpr int t = decltype(TI)::value;
          constexpr int mb = PISTO_STATIC_MASKS_AHEAD ? t * UN : 0;
          uint4 bgv[UN], gv[UN], lv[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int i = tid + (t * UN + u) * NT;
            bgv[u] = bgn[mb + u]; gv[u] = gn[mb + u];
            lv[u] = make_uint4(labc, labc, labc, labc);
            if (i < NV && multi) { const int4 q = lds_i4(lab_s + 16u * i); lv[u] = make_uint4(q.x, q.y, q.z, q.w); }
          }
          if constexpr (t + 1 < SWEEPS && !PISTO_STATIC_MASKS_AHEAD) {
#pragma unroll
            for (int u = 0; u < UN; u++) {
              const int i = tid + ((t + 1) * UN + u) * NT;
              gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
              if (i < NV) {
                if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + base) + i);
                if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + base) + i);
              }
            }
          }
          if (do_conf) {
#pragma unroll
            for (int u = 0; u < UN; u += 2) {
              const unsigned int gw[8] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w, gv[u + 1].x, gv[u + 1].y, gv[u + 1].z, gv[u + 1].w};
              const unsigned int lw8[8] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w, lv[u + 1].x, lv[u + 1].y, lv[u + 1].z, lv[u + 1].w};
              bitslice_count<C>(gw, lw8, cnt32);
            }
          }
          if (has_label) {
            if (has_bg) static_bg_overwrite4(lv, bgv, m4, bgl4);  // slots beyond NV hold masks of an earlier sweep or zeros
#pragma unroll
            for (int u = 0; u < UN; u++) {
              const int i = tid + (t * UN + u) * NT;
              if (i < NV) reinterpret_cast<uint4*>(p.label_out + base)[i] = lv[u];
            }
          }
        });
      }
      if (do_conf) {
#pragma unroll
        for (int bn = 0; bn < BINS; bn++) {
          const unsigned int cv = __reduce_add_sync(0xffffffffu, cnt32[bn]);
          if ((tid & 31) == 0 && cv) atomicAdd(&ctl->hist[bn], cv);
        }
      }
      for (int i = nvec * 16 + tid; i < (int)tpx; i += nt) {  // unaligned mask / label pointers
        const unsigned int lab = multi ? labsm[i] : (unsigned)tp.single;
        unsigned int o = lab;
        if (has_bg && p.bg[base + i] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
        if (do_conf) {
          const unsigned int gg = p.gt[base + i];
          if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + lab], 1u);
        }
        if (has_label) p.label_out[base + i] = (uint8_t)o;
      }
    }
    }  // !is_export
#ifndef PISTO_X_SKIP_EXPORT
    // With export warps (g.aux > 0) the compute warps leave the units of a multi-label tile to them -- the export warps have the
    // whole row loop's time for it and release the staging buffer when they are done -- and help only on single-label tiles.
    if (need_low && (is_export || g.aux == 0 || ctl_single >= 0)) export_units(n, sb, vb);
#endif
    __syncwarp();
    if (NB == 2 && (tid & 31) == 0) mbar_arrive(&ctl->empty[sb]);
    if (DEFER && !is_export && (tid & 31) == 0) { __threadfence_block(); atomicAdd(&ctl->tiles_done, 1u); }  // this warp's stores of the tile are out
  }
  if (do_conf) {
    bar_sync(1, ncomp);
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

// two entry points over the same body: 4 columns per thread (512 threads x 128 registers) and 2 columns per thread (832 x 72)
template <int C, int G, int VPG, int F, int NB>
__global__ void __launch_bounds__(512, 1) fuse_static_kernel(const __grid_constant__ FuseParams p, const __grid_constant__ StaticGeom g) {
  fuse_static_body<C, G, VPG, F, NB, 2>(p, g);
}
template <int C, int G, int VPG, int F, int NB>
__global__ void __launch_bounds__(kSNarrowThreads, 1) fuse_narrow_kernel(const __grid_constant__ FuseParams p, const __grid_constant__ StaticGeom g) {
  fuse_static_body<C, G, VPG, F, NB, 1>(p, g);
}

// =================================================================================================================================
// Two tiles in flight per SM: the same algorithm in a 256-thread CTA that fits twice on an SM.  What makes it fit:
//   * labels are staged at 2 bits per pixel (one byte per thread and row: 12.25 KB instead of 49 KB) and widened to bytes by the
//     vector pass (PRMT with the 2-bit fields as selector nibbles over the constant 0x03020100);
//   * ONE staging buffer: the 32x32 export -- the last reader of the raw views besides the rare exact pass -- runs right after the
//     pre-pass, the buffer is released and the next tile's TMA lands while this tile's row loop and vector pass run (the exact
//     pass reads its few samples from global memory, i.e. L2);
//   * no producer warp: all 8 warps compute; warp 7 (which owns no strip of the row loop) re-arms the TMA.
// The row loop's 7 strips run in two rounds over 4 x 56 threads.  While one CTA waits at a barrier or for its TMA, the other one
// has the SM's issue slots.
// =================================================================================================================================
constexpr int kDThreads = 256;
constexpr int kDSlots = 4;        // strips processed concurrently (56 threads each)
constexpr int kDQueueCap = 256;

// 16 labels (2 bits each: 4 bytes of the packed tile, byte = one thread-row = 4 pixels in field order (0, 2, 1, 3)) -> 16 bytes
__device__ __forceinline__ uint4 duo_expand16(unsigned int w) {
  unsigned int u, v, o[4];
  asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(u) : "r"(w), "r"(0u));  // (b0, 0, b1, 0)
  asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(v) : "r"(w), "r"(0u));  // (b2, 0, b3, 0)
  u = (u | (u << 6)) & 0x33333333u;  // per 16-bit half: nibbles = (p0, p1, p2, p3)
  v = (v | (v << 6)) & 0x33333333u;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(o[0]) : "r"(0x03020100u), "r"(0u), "r"(u));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(o[1]) : "r"(0x03020100u), "r"(0u), "r"(u >> 16));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(o[2]) : "r"(0x03020100u), "r"(0u), "r"(v));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(o[3]) : "r"(0x03020100u), "r"(0u), "r"(v >> 16));
  return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int C, int G, int VPG, int F>
__global__ void __launch_bounds__(kDThreads, 2) fuse_duo_kernel(const __grid_constant__ FuseParams p, const __grid_constant__ StaticGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int V = G * VPG;
  constexpr bool RT = F < 0;
  FCtl* ctl = reinterpret_cast<FCtl*>(smem_raw + g.ctl_off);
  float4* col4 = reinterpret_cast<float4*>(smem_raw + g.col4_off);
  uint32_t* col4i = reinterpret_cast<uint32_t*>(smem_raw + g.col4i_off);
  uint2* lowtap = reinterpret_cast<uint2*>(smem_raw + g.lowtab_off);
  float2* lowwt = reinterpret_cast<float2*>(smem_raw + g.lowtab_off + 32 * 8 * V);
  float* ymap = reinterpret_cast<float*>(smem_raw + g.ymap_off);
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + g.queue_off);
  uint8_t* lab2 = smem_raw + g.lab_off;                                   // [224][56] packed labels
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);

  const int tid = threadIdx.x;
  constexpr int nt = kDThreads;
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool need_low = RT ? (p.lowres_out != nullptr && p.low_fh > 0) : ((F & 8) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  constexpr int BINS = C * C;
  constexpr int UNITS = C * kSNE;
  constexpr int PROD = kDThreads - 32;  // first lane of warp 7: claims tiles and re-arms the TMA

  if (tid == 0) {
    mbar_init(&ctl->full[0], 1);
    mbar_init(&ctl->empty[0], kDThreads / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->maxbits[0] = ctl->maxbits[1] = 0u;
    ctl->qcount[0] = ctl->qcount[1] = 0u;
    ctl->lownext[0] = 0u;
  }
  for (int i = tid; i < 64; i += nt) ctl->hist[i] = 0;
  for (int i = tid; i < G * kSGX; i += nt) {
    const int gi = i / kSGX, gx = i - gi * kSGX;
    const int h = gi == 0 ? st_h(G, 0) : (gi == 1 ? st_h(G, G > 1 ? 1 : 0) : st_h(G, G > 2 ? 2 : 0));
    const float sc = (float)h / (float)kST;
    Lerp L[4];
#pragma unroll
    for (int c = 0; c < 4; c++) L[c] = pisto_src_index(sc, 4 * gx + c, h, false);
    col4[i] = make_float4(L[0].l1, L[1].l1, L[2].l1, L[3].l1);
    unsigned int sel = 0;
#pragma unroll
    for (int c = 1; c < 4; c++) sel |= (unsigned)(L[c].i0 - L[0].i0) << c;
    col4i[i] = (unsigned)(4 * L[0].i0) | (sel << 16);
  }
  if (need_low && tid < 32) {
    static_for<0, V>([&](auto VI) {
      constexpr int v = decltype(VI)::value, gi = v / VPG, h = st_h(G, gi);
      const Lerp L = pisto_src_index((float)h / (float)kST, 7 * tid + 3, h, false);
      lowtap[tid * V + v] = make_uint2((uint32_t)(g.vbase[v] + L.i0 * g.vcol[v]), (uint32_t)(g.vbase[v] + L.i1 * g.vcol[v]));
      lowwt[tid * G + gi] = make_float2(L.l0, L.l1);
    });
  }

  // fetch every view of tile n into the staging buffer (producer lane)
  auto issue_tile = [&](int n) {
    uint32_t total = 0;
#pragma unroll 1
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const char* start = reinterpret_cast<const char*>(vw.logits + (long long)n * vw.tile_stride);
      const char* end = start + (size_t)C * vw.h * vw.w * sizeof(float);
      const char* a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(start) & ~(uintptr_t)15);
      const char* a1 = reinterpret_cast<const char*>((reinterpret_cast<uintptr_t>(end) + 15) & ~(uintptr_t)15);
      char* dst = reinterpret_cast<char*>(vsm + g.view_off[v]);
      if (n == p.N - 1) {  // never read past the end of the caller's buffer: whole 16-byte units only, the tail by hand
        a1 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(end) & ~(uintptr_t)15);
        if (a1 < a0) a1 = a0;
        const char* t = a1 > start ? a1 : start;
        for (; t < end; t += 4) *reinterpret_cast<float*>(dst + (t - a0)) = *reinterpret_cast<const float*>(t);
      }
      const uint32_t bytes = (uint32_t)(a1 - a0);
      if (bytes) bulk_g2s(dst, a0, bytes, &ctl->full[0]);
      total += bytes;
    }
    mbar_arrive_expect_tx(&ctl->full[0], total);
  };
  // producer lane: publish tile `next` (claimed earlier) and start its transfers
  auto publish = [&](int next, const TilePresence& ntp) {
    const int tile = next < p.N ? next : -1;
    ctl->tile[0] = tile;
    ctl->pres_bits[0] = ntp.bits;
    ctl->pres_single[0] = ntp.single;
    ctl->lownext[0] = 0u;
    if (tile >= 0 && (need_low || ntp.single < 0)) issue_tile(tile);
    else mbar_arrive(&ctl->full[0]);
    if (tile >= 0) {
      const long long tile_px = (long long)kST * kST;
      if (has_bg && ((uintptr_t)p.bg & 15) == 0) bulk_prefetch_l2(p.bg + tile * tile_px, (uint32_t)tile_px);
      if (do_conf && ((uintptr_t)p.gt & 15) == 0) bulk_prefetch_l2(p.gt + tile * tile_px, (uint32_t)tile_px);
    }
  };

  __syncthreads();

  int next = 0;
  TilePresence next_tp; next_tp.bits = 0u; next_tp.single = -1;
  if (tid == PROD) {
    next = atomicAdd(g.counter, 1);
    if (next < p.N) next_tp = pisto_tile_presence(p, next);
    publish(next, next_tp);
  }

  auto export_units = [&](int n, const uint32_t (&vb)[V]) {
    const int lane = tid & 31;
    uint32_t sA[V], sB[V];
    float lx0[G], lx1[G];
#pragma unroll
    for (int v = 0; v < V; v++) { const uint2 t = lowtap[lane * V + v]; sA[v] = vb[v] + t.x; sB[v] = vb[v] + t.y; }
#pragma unroll
    for (int q = 0; q < G; q++) { const float2 t = lowwt[lane * G + q]; lx0[q] = t.x; lx1[q] = t.y; }
    float* out_n = p.lowres_out + (long long)n * C * kSLow * kSLow;
    for (;;) {
      int u = 0;
      if (lane == 0) u = (int)atomicAdd(&ctl->lownext[0], 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= UNITS) break;
      const int c = u / kSNE, e = u - c * kSNE;
      static_export_unit<C, G, VPG, kSLow / kSNE>(out_n, c, e, sA, sB, lx0, lx1, p.dec.rcp_v, p.dec.fV);
    }
  };

  const int grp = tid % kSGX, slot = tid / kSGX;
  const bool worker = slot < kDSlots;
  const int x = 4 * grp;
  const uint32_t ymap_s = smem_u32(ymap);
  const uint32_t col4_t = smem_u32(col4) + 16u * grp, col4i_t = smem_u32(col4i) + 4u * grp;
  const uint32_t lab_s = smem_u32(lab2);
  constexpr long long tpx = (long long)kST * kST;

  for (int k = 0;; k++) {
    const int b = k & 1;  // slot of the per-tile scratch (max, queue count)
    mbar_wait(&ctl->full[0], k & 1);
    const int n = ctl->tile[0];
    if (n < 0) break;
    TilePresence tp = pisto_tile_presence(p, n);
    if (p.present) { tp.bits = ctl->pres_bits[0]; tp.single = ctl->pres_single[0]; }
    const bool multi = tp.single < 0;
    uint32_t vb[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
      vb[v] = smem_u32(vsm + g.view_off[v]) + sh;
    }
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

    if (multi && P >= 2) {
      float mxf;
      int tidr[G];
#pragma unroll
      for (int q = 0; q < G; q++) tidr[q] = tid;
      if (P == 2) mxf = static_prepass<C, G, VPG, 1>(g, vb, cls, ymap_s, tidr, nt);
      else if (P == 3) mxf = static_prepass<C, G, VPG, 2>(g, vb, cls, ymap_s, tidr, nt);
      else mxf = static_prepass<C, G, VPG, (C >= 4 ? 3 : 1)>(g, vb, cls, ymap_s, tidr, nt);
      const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));
      if ((tid & 31) == 0) atomicMax(&ctl->maxbits[b], mx);
    }
    // the 32x32 logits now (fills the wait for the slowest pre-pass warp); the staged views are read once more by the exact pass
    if (need_low) export_units(n, vb);
    // the staging buffer is released (and the next tile's TMA started) as soon as its last reader is done: here for single-label
    // tiles, after the exact pass otherwise.  The next tile id and its presence vector are claimed by warp 7 -- which owns no
    // strip of the row loop -- while the other warps are in the row loop.
    auto release = [&]() {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&ctl->empty[0]);
      if (tid == PROD) {
        mbar_wait(&ctl->empty[0], k & 1);  // every warp is done with the staging buffer (and with the tile id / presence words)
        publish(next, next_tp);
      }
    };
    if (!multi) {
      if (tid == PROD) {
        next = atomicAdd(g.counter, 1);
        next_tp.bits = 0u; next_tp.single = -1;
        if (next < p.N) next_tp = pisto_tile_presence(p, next);
      }
      release();
    }
    __syncthreads();  // difference maps + max visible; every thread has left the previous tile's vector pass
    if (tid == 0) ctl->qcount[b ^ 1] = 0u;
    if (multi && tid == PROD) {
      next = atomicAdd(g.counter, 1);
      next_tp.bits = 0u; next_tp.single = -1;
      if (next < p.N) next_tp = pisto_tile_presence(p, next);
    }

    // The byte masks of the vector pass' first batch are requested early (after the row loop / right away for single-label tiles),
    // so that their latency is covered by the exact pass and the barriers; later batches are requested one step ahead.
    constexpr int UN = 4;
    const long long base = (long long)n * tpx;
    const bool vec_ok = ((((uintptr_t)p.bg | (uintptr_t)p.gt | (uintptr_t)p.label_out) & 15) == 0);
    const int nvec = vec_ok ? (int)(tpx / 16) : 0;
    uint4 bgn[UN], gn[UN];
    auto request_masks = [&]() {
#pragma unroll
      for (int u = 0; u < UN; u++) {
        const int i = tid + u * nt;
        gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        bgn[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < nvec) {
          if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + base) + i);
          if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + base) + i);
        }
      }
    };

    bool exact_all = false;
    if (multi) {
      float tau = 0.f;
      if (P >= 2) {
        const float A = __fmul_rn((float)V, __uint_as_float(ctl->maxbits[b]));
        tau = __fmaf_rn(A, g.tau_coef, g.tau_abs);
        if (!(A < 5e8f)) exact_all = true;
      } else {
        exact_all = true;
      }
      // exact sums of one pixel by a whole warp (lane = (view, class))
      auto exact_warp = [&](int yy, int xx) -> int {
        const int lane = tid & 31;
        float o = 0.f;
        if (lane < V * C) {
          const int v = lane / C, c = lane - v * C;
          const ViewDev& vw = p.view[v];
          const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, false);
          const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, false);
          const int pl = g.vbase[v] + c * 4 * vw.h * vw.w;
          const int r0 = pl + Ly.i0 * 4 * vw.w, r1 = pl + Ly.i1 * 4 * vw.w;
          const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
          const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
          const uint32_t base = smem_u32(vsm + g.view_off[v]) + sh;
          const float x00 = lds_f32(base + r0 + c0), x01 = lds_f32(base + r0 + c1), x10 = lds_f32(base + r1 + c0), x11 = lds_f32(base + r1 + c1);
          const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
          const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
          o = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
        }
        float a[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
          a[c] = __shfl_sync(0xffffffffu, o, c);
#pragma unroll
          for (int v = 1; v < V; v++) a[c] = __fadd_rn(a[c], __shfl_sync(0xffffffffu, o, (v * C + c) & 31));
        }
        return pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
      };
      if (!exact_all) {
        unsigned int c4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) c4[q] = 0x01010101u * (unsigned)cls[q < C ? q : 0];
#pragma unroll 1
        for (int strip = slot; strip < kSS && worker; strip += kDSlots) {
          const uint32_t lab_a = lab_s + (uint32_t)(strip * kSR * kSGX + grp);
          unsigned int unc = 0;
          if (P == 2) unc = static_rows<G, 1, kSU1, true>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (P == 3) unc = static_rows<G, 2, kSU2, true>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          else if (C >= 4 && P == 4) unc = static_rows<G, (C >= 4 ? 3 : 1), kSU2, true>(g, col4_t, col4i_t, ymap_s, lab_a, strip, c4, tau);
          if (unc) {
            if (P == 2) static_push_rows<G, 1, 2>(ctl, queue, kDQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
            else if (P == 3) static_push_rows<G, 2, 2>(ctl, queue, kDQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
            else static_push_rows<G, (C >= 4 ? 3 : 1), 2>(ctl, queue, kDQueueCap, b, col4_t, col4i_t, ymap_s, strip * kSR, x, unc, tau);
          }
        }
      }
      __syncthreads();  // every strip done: the queue is complete; everyone has read maxbits
      if (tid == 0) ctl->maxbits[b] = 0u;
      request_masks();
      const unsigned int nq = ctl->qcount[b];
      if (tid == 0 && g.stats) {
        atomicAdd(&g.stats[1], 1ull);
        if (!exact_all && nq <= (unsigned)kDQueueCap) atomicAdd(&g.stats[2], (unsigned long long)nq); else atomicAdd(&g.stats[3], 1ull);
      }
      if (nq > (unsigned)kDQueueCap) exact_all = true;  // overflow: redo the whole tile
      if (!exact_all && nq) {
        // queued groups, one pixel per warp at a time; the 2-bit field is patched with two word atomics (other fields of the word
        // may be patched by other warps at the same time)
        for (int j = tid >> 5; j < (int)nq; j += kDThreads / 32) {
          const uint32_t e = queue[j];
          const int yy = (int)(e >> 16), xx = (int)(e & 0xffffu);
          const int lab = exact_warp(yy, xx);
          if ((tid & 31) == 0) {
            const int bi = yy * kSGX + (xx >> 2);                          // byte of the packed tile
            const int sh = 8 * (bi & 3) + 2 * ((0x3120 >> (4 * (xx & 3))) & 3);  // field order (0, 2, 1, 3)
            unsigned int* w = reinterpret_cast<unsigned int*>(lab2) + (bi >> 2);
            atomicAnd(w, ~(3u << sh));
            atomicOr(w, (unsigned)lab << sh);
          }
        }
        release();
        __syncthreads();  // label tile complete
      }
      if (exact_all) {  // whole tile, one pixel per warp at a time would be slow: one pixel per thread, four pixels (one byte) at a time
        for (int j = tid; j < kST * kSGX; j += nt) {
          const int yy = j / kSGX, xg = j - yy * kSGX;
          unsigned int byte = 0;
#pragma unroll 1
          for (int q = 0; q < 4; q++) {
            const int xx = 4 * xg + q;
            float a[C];
#pragma unroll
            for (int v = 0; v < V; v++) {
              const ViewDev& vw = p.view[v];
              const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, false);
              const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, false);
              const int r0 = g.vbase[v] + Ly.i0 * 4 * vw.w, r1 = g.vbase[v] + Ly.i1 * 4 * vw.w;
              const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
#pragma unroll
              for (int c = 0; c < C; c++) {
                const int pl = c * 4 * vw.h * vw.w;
                const float x00 = lds_f32(vb[v] + r0 + pl + c0), x01 = lds_f32(vb[v] + r0 + pl + c1);
                const float x10 = lds_f32(vb[v] + r1 + pl + c0), x11 = lds_f32(vb[v] + r1 + pl + c1);
                const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
                const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
                const float u = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
                a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
              }
            }
            const int lab = pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
            byte |= (unsigned)lab << (2 * ((0x3120 >> (4 * q)) & 3));
          }
          lab2[j] = (uint8_t)byte;
        }
        release();
        __syncthreads();
      }
      if (!exact_all && !nq) release();
    } else {
      request_masks();
    }

    // ---- vector pass: confusion, background overwrite, 16-byte label stores (single-label tiles: constant label)
    {
      const unsigned int labc = 0x01010101u * (unsigned)(multi ? 0 : tp.single), bgl4 = 0x01010101u * (unsigned)p.bg_label;
      const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match;
      unsigned int cnt32[BINS];
#pragma unroll
      for (int i = 0; i < BINS; i++) cnt32[i] = 0;
      for (int i0 = tid; i0 < nvec; i0 += UN * nt) {
        uint4 bgv[UN], gv[UN], lv[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const int i = i0 + u * nt;
          bgv[u] = bgn[u]; gv[u] = gn[u];  // requested one step ago
          lv[u] = make_uint4(labc, labc, labc, labc);
          if (i < nvec && multi) lv[u] = duo_expand16(lds_u32(lab_s + 4u * i));
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {  // request the next batch
          const int i = i0 + (UN + u) * nt;
          gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (i < nvec) {
            if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + base) + i);
            if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + base) + i);
          }
        }
        if (do_conf) {
#pragma unroll
          for (int u = 0; u < UN; u += 2) {
            const unsigned int gw[8] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w, gv[u + 1].x, gv[u + 1].y, gv[u + 1].z, gv[u + 1].w};
            const unsigned int lw8[8] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w, lv[u + 1].x, lv[u + 1].y, lv[u + 1].z, lv[u + 1].w};
            bitslice_count<C>(gw, lw8, cnt32);
          }
        }
        if (has_label) {
          if (has_bg) static_bg_overwrite4(lv, bgv, m4, bgl4);  // lanes beyond nvec hold zero masks
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int i = i0 + u * nt;
            if (i < nvec) reinterpret_cast<uint4*>(p.label_out + base)[i] = lv[u];
          }
        }
      }
      if (do_conf && nvec) {
#pragma unroll
        for (int bn = 0; bn < BINS; bn++) {
          const unsigned int cv = __reduce_add_sync(0xffffffffu, cnt32[bn]);
          if ((tid & 31) == 0 && cv) atomicAdd(&ctl->hist[bn], cv);
        }
      }
      for (int i = nvec * 16 + tid; i < (int)tpx; i += nt) {  // unaligned mask / label pointers
        unsigned int lab = (unsigned)tp.single;
        if (multi) lab = (lab2[i >> 2] >> (2 * ((0x3120 >> (4 * (i & 3))) & 3))) & 3u;
        unsigned int o = lab;
        if (has_bg && p.bg[base + i] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
        if (do_conf) {
          const unsigned int gg = p.gt[base + i];
          if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + lab], 1u);
        }
        if (has_label) p.label_out[base + i] = (uint8_t)o;
      }
    }
  }
  if (do_conf) {
    __syncthreads();
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
template <int C, int G, int VPG>
static bool make_static_geom(const pisto_ctx* h, const FuseParams& p, int nbuf, StaticGeom* g, int np = 2) {
  constexpr int V = G * VPG;
  memset(g, 0, sizeof(*g));
  if (p.C != C || p.V != V || p.T_h != kST || p.T_w != kST) return false;
  const bool low = p.lowres_out && p.low_fh > 0;
  if (low && (p.low_h != kSLow || p.low_w != kSLow || nbuf != 2)) return false;
  for (int v = 0; v < V; v++) {
    const ViewDev& vw = p.view[v];
    const int hh = st_h(G, v / VPG);
    if (vw.h != hh || vw.w != hh || vw.map.ho != hh || vw.map.wo != hh) return false;
    if (vw.map.ai != 1 || vw.map.aj != 0 || vw.map.bi != 0 || (vw.map.bj != 1 && vw.map.bj != -1) || vw.map.a0 != 0) return false;  // identity or hflip
    g->vbase[v] = 4 * vw.map.b0;
    g->vcol[v] = 4 * vw.map.bj;
    // every constexpr table entry against the arithmetic every other kernel (and the oracle) uses
    for (int y = 0; y < kST; y++) {
      const Lerp L = pisto_src_index(vw.scale_h, y, hh, false);
      const int s = y / kSR, r = y % kSR;
      int i0 = s * (hh / 7) + st_i0rel(hh, r);
      const float l1 = st_l1(hh, r);
      const bool clamped = i0 < 0 || i0 >= hh - 1;
      if (!clamped && (L.i0 != i0 || L.i1 != i0 + 1 || L.l1 != l1 || L.l0 != 1.f - l1)) return false;
      if (clamped && L.i0 != (i0 < 0 ? 0 : hh - 1)) return false;
      if (clamped && i0 < 0 && L.l1 != 0.f) return false;
    }
    for (int ly = 0; ly < kSLow; ly++) {
      const Lerp L = pisto_src_index(vw.scale_h, 7 * ly + 3, hh, false);
      if (L.i0 != lo_i0(hh, ly) || L.i1 != lo_i1(hh, ly) || L.l1 != lo_l1(hh, ly) || L.l0 != 1.f - lo_l1(hh, ly)) return false;
    }
    for (int x = 0; x < kST; x += 4) {  // the 4 columns of a thread lie within two adjacent source cells (3-tap loads)
      const Lerp L0 = pisto_src_index(vw.scale_w, x, hh, false);
      for (int c = 0; c < 4; c++) {
        const Lerp L = pisto_src_index(vw.scale_w, x + c, hh, false);
        if (L.i0 < L0.i0 || L.i0 > L0.i0 + 1) return false;
        if (L.i1 != L.i0 && L.i1 != L.i0 + 1) return false;
        if (L.i1 == L.i0 && L.i0 != hh - 1 && L.l1 != 0.f) return false;
      }
    }
  }
  for (int r = 0; r < kSR; r++)
    for (int q = 0; q < G; q++) {
      g->wtab[r][q] = st_l1(st_h(G, q), r) - 1.f;
      if (st_adv(st_h(G, q), r)) g->adv[q] |= 1u << r;
    }
  g->aux = low ? kSAux : 0;
  g->cwarps = (kST / (2 * np) * kSS + 31) / 32;
  g->threads = g->cwarps * 32 + 32 * g->aux + 32 + ((PISTO_STATIC_DEFER && np == 2) ? 32 : 0);  // + producer warp (+ fixer warp)
  if (np == 2 && g->threads > 512) return false;   // fuse_static_kernel's launch bound
  int fl = 0;
  for (int v = 0; v < V; v++) {
    g->view_off[v] = fl;
    fl += (C * p.view[v].h * p.view[v].w + 3 + 3 + 3) & ~3;
  }
  g->buf_floats = fl;
  const float cE = 2.f * VPG + 4.f * G + 2.f * V + 20.f;
  g->tau_coef = 2.f * cE * 5.9604645e-8f + 2.5e-7f;
  g->tau_abs = p.dec.margin_abs * 1.01f;
  constexpr int KPmax = C >= 4 ? 4 : C - 1;
  int off = 0;
  g->ctl_off = off; off += (int)((sizeof(FCtl) + 127) & ~127u);
  g->col4_off = off; off += 16 * G * kSGX;                        // [G][56] float4, or [G][112] float2
  g->col4i_off = off; off += 4 * G * 2 * kSGX; off = (off + 15) & ~15;
  g->w3_off = -1;
  if (C >= 4 && np == 2) { g->w3_off = off; off += 48 * G * kSGX; }   // only the three-field loop reads it
  g->lowtab_off = off; off += low ? 32 * (8 * V + 8 * G) : 0;
  g->ymap_off = off; off += st_ybytes(G, G, KPmax);
  g->queue_off = off; off += 4 * kFQueueCap;  // in-tile queue, or 8 slots x (64 entries + 64 label words) of the deferred exact pass
  g->lab_off = off; off += kST * kST;
  off = (off + 127) & ~127;
  g->views_off = off; off += nbuf * 4 * fl + 256;  // slack: the export's shared row loads may touch one row past the last view
  g->smem_bytes = off;
  return off <= h->smem_optin - 1024;
}

// geometry of the two-CTAs-per-SM kernel: the checks and tables of make_static_geom, its own shared-memory layout
template <int C, int G, int VPG>
static bool make_duo_geom(const pisto_ctx* h, const FuseParams& p, StaticGeom* g) {
  constexpr int V = G * VPG;
  if (!make_static_geom<C, G, VPG>(h, p, 2, g)) return false;
  const bool low = p.lowres_out && p.low_fh > 0;
  g->aux = 0; g->cwarps = kDThreads / 32; g->threads = kDThreads;
  constexpr int KPmax = C >= 4 ? 4 : C - 1;
  int off = 0;
  g->ctl_off = off; off += (int)((sizeof(FCtl) + 127) & ~127u);
  g->col4_off = off; off += 16 * G * kSGX;
  g->col4i_off = off; off += 4 * G * kSGX; off = (off + 15) & ~15;
  g->lowtab_off = off; off += low ? 32 * (8 * V + 8 * G) : 0;
  g->ymap_off = off; off += st_ybytes(G, G, KPmax);
  g->queue_off = off; off += 4 * kDQueueCap;
  g->lab_off = off; off += kST * kSGX;
  off = (off + 127) & ~127;
  g->views_off = off; off += 4 * g->buf_floats + 256;  // slack: the export's shared row loads may touch one row past the last view
  g->smem_bytes = off;
  return 2 * (off + 1024) <= 233472;
}

template <int C, int G, int VPG, int F>
int launch_duo(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  StaticGeom g;
  if (!make_duo_geom<C, G, VPG>(h, p, &g)) return PISTO_OK;
  auto kern = fuse_duo_kernel<C, G, VPG, F>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int sched_slot = 0;
  { const int rc = pisto_sched_acquire(h, st, &g.counter, &sched_slot); if (rc != PISTO_OK) return rc; }
  g.stats = h->stats;
  const int slots = 2 * h->sm_count;
  const int grid = p.N < slots ? p.N : slots;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  { const int rc = pisto_sched_release(h, st, sched_slot); if (rc != PISTO_OK) return rc; }
  *launched = true;
  return PISTO_OK;
}

template <int C, int G, int VPG, int F, int NB>
int launch_narrow(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  StaticGeom g;
  if (!make_static_geom<C, G, VPG>(h, p, NB, &g, 1) || g.threads > kSNarrowThreads) return PISTO_OK;
  auto kern = fuse_narrow_kernel<C, G, VPG, F, NB>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int sched_slot = 0;
  { const int rc = pisto_sched_acquire(h, st, &g.counter, &sched_slot); if (rc != PISTO_OK) return rc; }
  g.stats = h->stats;
  const int grid = p.N < h->sm_count ? p.N : h->sm_count;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  { const int rc = pisto_sched_release(h, st, sched_slot); if (rc != PISTO_OK) return rc; }
  *launched = true;
  return PISTO_OK;
}

template <int C, int G, int VPG, int F, int NB>
int launch_static(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  StaticGeom g;
  if (!make_static_geom<C, G, VPG>(h, p, NB, &g)) return PISTO_OK;
  auto kern = fuse_static_kernel<C, G, VPG, F, NB>;
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int sched_slot = 0;
  { const int rc = pisto_sched_acquire(h, st, &g.counter, &sched_slot); if (rc != PISTO_OK) return rc; }
  g.stats = h->stats;
  const int grid = p.N < h->sm_count ? p.N : h->sm_count;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  { const int rc = pisto_sched_release(h, st, sched_slot); if (rc != PISTO_OK) return rc; }
  *launched = true;
  return PISTO_OK;
}

}  // namespace
