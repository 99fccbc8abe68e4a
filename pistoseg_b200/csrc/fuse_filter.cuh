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
//               cE = 2 n_max + 4 G + 2 V + 20 bounds the rounding of both evaluations; the remaining
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

#ifndef PISTO_FTHREADS
#define PISTO_FTHREADS 512
#endif
constexpr int kFMaxThreads = PISTO_FTHREADS;
#ifndef PISTO_FAUX
#define PISTO_FAUX 2
#endif
constexpr int kFAux = PISTO_FAUX;  // warps that evaluate the 32x32 logit export concurrently with the row loop
// CTA size: three difference fields for four columns (C = 4, every class present) need ~140 registers per thread; a register-file
// partition (16384 registers, every 4th warp) holds 3 warps of up to 168 registers, so the CTA is 12 warps
template <int C> constexpr int filter_threads() { return C >= 4 ? 384 : kFMaxThreads; }
// export warps: none when the feature mask is known at compile time and has no 32x32 export
template <int F> constexpr int filter_aux() { return (F >= 0 && !(F & 8)) ? 0 : kFAux; }
constexpr int kFQueueCap = 1024;
constexpr int kFMaxGroups = 5;

struct FilterGeom {
  int GX, S, threads, cwarps;          // threads per output row, strips, CTA size (incl. the producer warp), compute warps
  int GXP;                             // column pairs per row (T_w / 2)
  int lab_stride;                      // bytes between rows of the shared-memory label tile
  int strip_y0[33];
  int view_off[PISTO_MAX_VIEWS];       // float offset of each view inside one staging buffer (16-byte aligned)
  int plane_bytes[PISTO_MAX_VIEWS];    // h*w*4
  int vbase[PISTO_MAX_VIEWS], vrow[PISTO_MAX_VIEWS], vcol[PISTO_MAX_VIEWS];  // byte address of de-augmented (i, j): vbase + i*vrow + j*vcol
  int group_of[PISTO_MAX_VIEWS];
  int first_in_group[PISTO_MAX_VIEWS];
  int buf_floats;
  int nbuf;                            // staging buffers: 2, or 1 (views released after the pre-pass, exact pass reads global memory)
  int aux;                             // export warps
  int g_ho[kFMaxGroups], g_wo[kFMaxGroups], g_same_w[kFMaxGroups];
  float g_scale_h[kFMaxGroups], g_scale_w[kFMaxGroups];
  int g_ybytes[kFMaxGroups];           // byte offset of group g's first difference map inside the Y area
  int g_mapbytes[kFMaxGroups];         // ho*wo*4
  unsigned int g_inv[kFMaxGroups];     // ceil(2^32 / wo): idx / wo == umulhi(idx, inv) for idx < 2^16
  float tau_coef, tau_abs;             // tau = V * max|x| * tau_coef + tau_abs
  int ctl_off, rowtab_off, rowoff_off, cola_off, colb_off, lowrow_off, lowcol_off, ymap_off, queue_off, lab_off, views_off;
  int col4_off, col4i_off;             // 4-column tables (T4): [G][GX] float4 l1 of the thread's columns; [G][GX] u32 (4*j0 of column 0) | sel << 16
  int smem_bytes;
  int* counter;
  unsigned long long* stats;          // [4] device counters of the handle (pisto_filter_stats) or NULL
};

struct FCtl {
  uint64_t full[2];   // producer -> consumers: tile id published (+ TMA bytes landed)
  uint64_t empty[2];  // consumers -> producer: staging buffer may be refilled (one arrival per compute warp)
  int tile[2];
  unsigned int pres_bits[2];  // class-presence summary of the published tile (read from global memory by the producer, one tile ahead)
  int pres_single[2];
  unsigned int maxbits[2];  // max |x| of the tile (IEEE bits; NaN > Inf > finite), slot = tile parity
  unsigned int qcount[2];   // uncertain pixels queued by the row loop
  unsigned int lownext[2];  // next low-resolution row of the logit export to be claimed (any warp may claim)
  unsigned int hist[64];
  // fuse_static.cuh: everything a consumer thread needs to know about the published tile in one 16-byte load, prepared by the
  // producer lane: {presence bits, single label or -1, candidate classes (one byte each, ascending), staging shifts of the
  // views ((address & 12) >> 2, two bits per view) | number of candidates << 28}
  uint4 head[2];
  // fuse_static.cuh, deferred exact pass: group queues in eight rotating slots (tile k of the CTA uses slot k & 7; a dedicated warp
  // empties it once every compute warp has counted the tile done), the tile each slot belongs to, and the hand-shake counters
  unsigned int qn[8];
  int2 qmeta[8];            // {tile index, presence bits}
  unsigned int tiles_done;  // += 1 per compute warp and tile (after its last store of the tile)
  unsigned int fixed;       // tiles whose queue has been emptied
  int ntiles;               // number of tiles this CTA processed, -1 until the producer knows
};

// label among cls[0..K] from the difference candidates (0, d[0] .. d[K-1]); returns whether the lead exceeds tau
// (used on the rare path only: the row loop decides through sign masks)
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
template <int NP> __device__ __forceinline__ void sts_px(uint32_t a, unsigned int v) {
  if (NP == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
  else asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}

#ifndef PISTO_K2_TOURNAMENT
#define PISTO_K2_TOURNAMENT 1
#endif
#ifndef PISTO_K3_TOURNAMENT
#define PISTO_K3_TOURNAMENT 1
#endif
// byte j of the result = 0xff if a[j] is negative (PRMT replicates the sign bit of the selected byte when bit 3 of the
// selector nibble is set); NP == 1: only bytes 0 and 1 are meaningful
template <int NP> __device__ __forceinline__ unsigned int negmask(const float (&a)[2 * NP]) {
  unsigned int t01, t23, r;
  asm("prmt.b32 %0, %1, %2, 0x00fb;" : "=r"(t01) : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])));
  if (NP == 1) return t01;
  asm("prmt.b32 %0, %1, %2, 0x00fb;" : "=r"(t23) : "r"(__float_as_uint(a[2 * NP - 2])), "r"(__float_as_uint(a[2 * NP - 1])));
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(t01), "r"(t23));
  return r;
}
template <int N> __device__ __forceinline__ float minabs(const float (&a)[N]) {
  float m = fabsf(a[0]);
#pragma unroll
  for (int i = 1; i < N; i++) m = fminf(m, fabsf(a[i]));
  return m;
}

// Labels of the thread's 2*NP pixels from the packed difference fields acc[k][q] (pixel pair q, candidate k+1 minus
// candidate 0).  The order of the candidates (0, D_1 .. D_K) is read off the SIGNS of the D_k and of their pairwise
// differences; the result is trusted (return value) only when every one of those quantities is further than tau from
// zero, which implies the best candidate leads every other one by more than tau.  c4[k] = 0x01010101 * cls[k].
template <int K, int NP>
__device__ __forceinline__ float labels_from_diffs_min(const u64 (&acc)[K][NP], const unsigned int (&c4)[4], unsigned int& lab4);
template <int K, int NP>
__device__ __forceinline__ bool labels_from_diffs(const u64 (&acc)[K][NP], const unsigned int (&c4)[4], float tau, unsigned int& lab4) {
  return labels_from_diffs_min<K, NP>(acc, c4, lab4) > tau;
}
// the same, returning the smallest |quantity| instead of testing it (the caller may fold several rows into one test)
template <int K, int NP>
__device__ __forceinline__ float labels_from_diffs_min(const u64 (&acc)[K][NP], const unsigned int (&c4)[4], unsigned int& lab4) {
  float d[K][2 * NP];
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int q = 0; q < NP; q++) unpack2(acc[k][q], d[k][2 * q], d[k][2 * q + 1]);
  float mn;
  if (K == 1) {
    const unsigned int n1 = negmask<NP>(d[0]);
    lab4 = (c4[0] & n1) | (c4[1] & ~n1);
    mn = minabs(d[0]);
  } else if (K == 2) {
    float e12[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; q++) unpack2(sub2(acc[0][q], acc[K > 1 ? 1 : 0][q]), e12[2 * q], e12[2 * q + 1]);
#if PISTO_K2_TOURNAMENT
    // three candidates (0, D1, D2): m = max(D1, D2) plays 0; trusted when |m| and |D1 - D2| exceed tau (the winner then leads both
    // others by more than tau).  Unlike the pairwise form it does not flag a pixel whose LOSING candidate happens to sit near 0.
    float m[2 * NP];
#pragma unroll
    for (int j = 0; j < 2 * NP; j++) m[j] = fmaxf(d[0][j], d[K > 1 ? 1 : 0][j]);
    const unsigned int nm = negmask<NP>(m), n12 = negmask<NP>(e12);
    const unsigned int b = (c4[2] & n12) | (c4[1] & ~n12);
    lab4 = (c4[0] & nm) | (b & ~nm);
    mn = fminf(minabs(m), minabs(e12));
#else
    const unsigned int n1 = negmask<NP>(d[0]), n2 = negmask<NP>(d[K > 1 ? 1 : 0]), n12 = negmask<NP>(e12);
    const unsigned int m0 = n1 & n2, m1 = ~n1 & ~n12;
    lab4 = (c4[0] & m0) | (c4[1] & m1) | (c4[2] & ~(m0 | m1));
    mn = fminf(fminf(minabs(d[0]), minabs(d[K > 1 ? 1 : 0])), minabs(e12));
#endif
  } else {
#if PISTO_K3_TOURNAMENT
    // four candidates (0, D1, D2, D3) as a two-round tournament instead of all six pairwise comparisons: m01 = max(0, D1) and
    // m23 = max(D2, D3) play the final f = m01 - m23.  The result is trusted when |D1|, |D2 - D3| and |f| all exceed tau: the winner
    // then leads the other candidate of its own pair by |D1| or |D2 - D3| > tau and both candidates of the other pair by |f| > tau,
    // i.e. every other candidate by more than tau -- the same guarantee as the pairwise form, with 12 instead of 24 magnitudes to
    // reduce, 9 instead of 18 sign extractions and 3 instead of 6 extra differences per 4 pixels.  (Ties never reach a decision here:
    // tau > 0 sends them to the exact evaluation.)
    float e23[2 * NP], f[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; q++) unpack2(sub2(acc[K > 1 ? 1 : 0][q], acc[K > 2 ? 2 : 0][q]), e23[2 * q], e23[2 * q + 1]);
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const u64 m01 = pack2(fmaxf(d[0][2 * q], 0.f), fmaxf(d[0][2 * q + 1], 0.f));
      const u64 m23 = pack2(fmaxf(d[K > 1 ? 1 : 0][2 * q], d[K > 2 ? 2 : 0][2 * q]), fmaxf(d[K > 1 ? 1 : 0][2 * q + 1], d[K > 2 ? 2 : 0][2 * q + 1]));
      unpack2(sub2(m01, m23), f[2 * q], f[2 * q + 1]);
    }
    const unsigned int n1 = negmask<NP>(d[0]), n23 = negmask<NP>(e23), nf = negmask<NP>(f);
    const unsigned int a = (c4[0] & n1) | (c4[1] & ~n1), b = (c4[3] & n23) | (c4[2] & ~n23);
    lab4 = (b & nf) | (a & ~nf);
    mn = fminf(fminf(minabs(d[0]), minabs(e23)), minabs(f));
#else
    float e12[2 * NP], e13[2 * NP], e23[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
      unpack2(sub2(acc[0][q], acc[K > 1 ? 1 : 0][q]), e12[2 * q], e12[2 * q + 1]);
      unpack2(sub2(acc[0][q], acc[K > 2 ? 2 : 0][q]), e13[2 * q], e13[2 * q + 1]);
      unpack2(sub2(acc[K > 1 ? 1 : 0][q], acc[K > 2 ? 2 : 0][q]), e23[2 * q], e23[2 * q + 1]);
    }
    const unsigned int n1 = negmask<NP>(d[0]), n2 = negmask<NP>(d[K > 1 ? 1 : 0]), n3 = negmask<NP>(d[K > 2 ? 2 : 0]);
    const unsigned int n12 = negmask<NP>(e12), n13 = negmask<NP>(e13), n23 = negmask<NP>(e23);
    const unsigned int m0 = n1 & n2 & n3, m1 = ~n1 & ~n12 & ~n13, m2 = ~n2 & n12 & ~n23;
    lab4 = (c4[0] & m0) | (c4[1] & m1) | (c4[2] & m2) | (c4[3] & ~(m0 | m1 | m2));
    mn = fminf(fminf(fminf(minabs(d[0]), minabs(d[K > 1 ? 1 : 0])), fminf(minabs(d[K > 2 ? 2 : 0]), minabs(e12))), fminf(minabs(e13), minabs(e23)));
#endif
  }
  return mn;
}

// rare path: which of the thread's pixels really fail the lead test (top-2 gap of the candidates <= tau)
template <int K, int NP, int C>
__device__ __forceinline__ unsigned int uncertain_mask(const u64 (&acc)[K][NP], const int (&cls)[C], float tau) {
  unsigned int unc = 0;
#pragma unroll
  for (int q = 0; q < NP; q++) {
    float d0[K], d1[K];
#pragma unroll
    for (int k = 0; k < K; k++) unpack2(acc[k][q], d0[k], d1[k]);
    int l;
    if (!decide_diff<K, C>(d0, cls, tau, l)) unc |= 1u << (2 * q);
    if (!decide_diff<K, C>(d1, cls, tau, l)) unc |= 2u << (2 * q);
  }
  return unc;
}

// packed 8-bit confusion counters (two 64-bit registers hold C*C <= 16 bins)
__device__ __forceinline__ void count_bin(u64& lo, u64& hi, unsigned int bn, unsigned int times) {
  const u64 inc = (u64)times << (8 * (bn & 7));
  if (bn < 8) lo += inc; else hi += inc;
}
// 4 pixels: ground truth bytes g4, label bytes l4
template <int C>
__device__ __forceinline__ void count_word(u64& lo, u64& hi, unsigned int g4, unsigned int l4) {
  const unsigned int g0 = g4 & 0xffu, l0 = l4 & 0xffu;
  if (g4 == g0 * 0x01010101u && l4 == l0 * 0x01010101u) {  // masks are piecewise constant: the common case
    if (g0 < (unsigned)C) count_bin(lo, hi, g0 * C + l0, 4u);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned int gg = (g4 >> (8 * j)) & 0xffu;
      if (gg < (unsigned)C) count_bin(lo, hi, gg * C + ((l4 >> (8 * j)) & 0xffu), 1u);
    }
  }
}

// Confusion counts of 32 pixels from bit planes (the method of confusion.cu): word j of gw / lw contributes bit k of its bytes
// at bit j of the plane byte -- one shift and one LOP3 per (word, plane), the pixel permutation is the same for every plane --
// then cnt[a*C + b] += popc(G_a & L_b).  Labels are < C <= 4 (two planes); ground-truth bytes >= 4 are never counted.  The
// cost does not depend on how ragged the label map is.
template <int C>
__device__ __forceinline__ void bitslice_count(const unsigned int (&gw)[8], const unsigned int (&lw)[8], unsigned int (&cnt)[C * C]) {
  unsigned int gp0 = 0, gp1 = 0, gpx = 0, lp0 = 0, lp1 = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const unsigned int m = 0x01010101u << j;
    gp0 |= (gw[j] << j) & m;
    gp1 |= (j == 0 ? (gw[j] >> 1) : (gw[j] << (j - 1))) & m;
    const unsigned int hi = ((gw[j] >> 2) & 0x3f3f3f3fu) + 0x3f3f3f3fu;  // bit 6 set iff the byte is >= 4
    gpx |= (j <= 6 ? (hi >> (6 - j)) : (hi << (j - 6))) & m;
    lp0 |= (lw[j] << j) & m;
    lp1 |= (j == 0 ? (lw[j] >> 1) : (lw[j] << (j - 1))) & m;
  }
  unsigned int Gm[4], Pm[4];
  Gm[0] = ~gp1 & ~gp0 & ~gpx; Gm[1] = ~gp1 & gp0 & ~gpx; Gm[2] = gp1 & ~gp0 & ~gpx; Gm[3] = gp1 & gp0 & ~gpx;
  Pm[0] = ~lp1 & ~lp0; Pm[1] = ~lp1 & lp0; Pm[2] = lp1 & ~lp0; Pm[3] = lp1 & lp0;
#pragma unroll
  for (int a = 0; a < C; a++)
#pragma unroll
    for (int c = 0; c < C; c++) cnt[a * C + c] += __popc(Gm[a < 4 ? a : 0] & Pm[c < 4 ? c : 0]);
}

// ---- the row loop of one strip, K difference fields ----------------------------------------------------------------
// LSM: labels go to the shared-memory label tile (background / confusion / the global store happen in the vector pass
// after the exact pass); otherwise they go straight to global memory with the byte masks prefetched 4 rows ahead.
// T4 (NP == 2 only): the thread's 4 adjacent columns touch at most 3 adjacent source columns j, j+1, j+2 (checked on the
// host: up-sampling by >= 3), the difference maps are stored [row][column][k] with two replicated pad columns per row, so
// one refill is ONE 16-byte table load (the four l1 weights; l0 = 1 - l1 as in pisto_src_index) and three K-wide data loads
// instead of 2 * (2 table + 4 K data) loads: the row loop is bound by shared-memory wavefronts, not by arithmetic.
// colA_t = address of the thread's l1 entry, colB_t = address of its (4*j0 | sel << 16) entry; group stride 16 * GX / 4 * GX.
// TT == 2 (NP == 1, the block-tiled kernel): the same layout for a column pair -- colA_t = address of the pair's {l1, l1} entry
// (group stride 8 * GXP), colB_t = address of its (4*j0 | sel << 16) entry (group stride 4 * GXP).
template <int C, int G, int F, int NP, int K, bool LSM, int TT = 0>
__device__ __forceinline__ void filter_rows(const FuseParams& p, const FilterGeom& g, FCtl* ctl, uint32_t* queue, int b,
                                            uint32_t rowtab_s, uint32_t rowoff_s, uint32_t colA_t, uint32_t colB_t, uint32_t ymap_s,
                                            uint32_t lab_s, int n, int x, int ys, int ye, const int (&cls)[C], float tau,
                                            u64& cnt_lo, u64& cnt_hi) {
  constexpr bool T4 = TT == 4, T2 = TT == 2;
  static_assert(!T4 || (NP == 2 && K <= 3), "4-column tables: NP == 2, K <= 3");
  static_assert(!T2 || (NP == 1 && K <= 3), "pair 3-tap tables: NP == 1, K <= 3");
  constexpr int KP = K == 3 ? 4 : K;  // floats per cell of an interleaved map
  constexpr bool RT = F < 0;
  constexpr int RS = 16 * ((G + 2) / 2);
  const bool has_bg = RT ? (p.bg != nullptr) : ((F & 1) != 0);
  const bool do_conf = RT ? (p.conf != nullptr && p.gt != nullptr) : ((F & 2) != 0);
  const bool has_label = RT ? (p.label_out != nullptr) : ((F & 16) != 0);
  const int T_w = p.T_w;
  const int lab_stride = g.lab_stride;
  const uint32_t colg = 16u * g.GXP;
  unsigned int c4[4];
#pragma unroll
  for (int k = 0; k < 4; k++) c4[k] = 0x01010101u * (unsigned)cls[k < C ? k : 0];

  u64 Hb[G][K][NP], Dh[G][K][NP], base[K][NP];
  uint32_t yb[G];        // T4: address of (row 0, source column of the thread's first column, k = 0) of group gi's map
  unsigned int selm = 0; // T4: bit 4*gi + c: column c reads source columns (j+1, j+2) instead of (j, j+1)
  if (T4 || T2) {
#pragma unroll
    for (int gi = 0; gi < G; gi++) {
      const uint32_t u = lds_u32(colB_t + gi * 4u * (T4 ? g.GX : g.GXP));
      yb[gi] = ymap_s + g.g_ybytes[gi] + KP * (u & 0xffffu);
      selm |= (u >> 16) << (4 * gi);
    }
  }
  const u64 one2 = pack2(1.f, 1.f);
  // horizontally interpolated values of one row (byte offset `row` inside a map) of every difference map of group gi
  auto load_h = [&](int gi, uint32_t row, u64 (&H)[K][NP]) {
    if constexpr (T2) {
      const uint32_t a = yb[gi] + KP * row;
      const float2 L1 = lds_f2(colA_t + gi * 8u * g.GXP);
      const u64 l1p = pack2(L1.x, L1.y), l0p = sub2(one2, l1p);
      const bool s1 = (selm >> (4 * gi + 1)) & 1u;
      float y0[K], y1[K], y2[K];
      if constexpr (K == 1) {
        y0[0] = lds_f32(a); y1[0] = lds_f32(a + 4u); y2[0] = lds_f32(a + 8u);
      } else if constexpr (K == 2) {
        const float2 v0 = lds_f2(a), v1 = lds_f2(a + 8u), v2 = lds_f2(a + 16u);
        y0[0] = v0.x; y0[K - 1] = v0.y; y1[0] = v1.x; y1[K - 1] = v1.y; y2[0] = v2.x; y2[K - 1] = v2.y;
      } else {
        const float4 v0 = lds_f4(a), v1 = lds_f4(a + 16u), v2 = lds_f4(a + 32u);
        y0[0] = v0.x; y0[1 % K] = v0.y; y0[K - 1] = v0.z; y1[0] = v1.x; y1[1 % K] = v1.y; y1[K - 1] = v1.z;
        y2[0] = v2.x; y2[1 % K] = v2.y; y2[K - 1] = v2.z;
      }
#pragma unroll
      for (int k = 0; k < K; k++) {
        const float a1 = s1 ? y1[k] : y0[k], b1 = s1 ? y2[k] : y1[k];
        H[k][0] = fma2(l0p, pack2(y0[k], a1), mul2(l1p, pack2(y1[k], b1)));
      }
      return;
    }
    if constexpr (T4) {
      const uint32_t a = yb[gi] + KP * row;
      const float4 L1 = lds_f4(colA_t + gi * 16u * g.GX);
      const u64 l1a = pack2(L1.x, L1.y), l1b = pack2(L1.z, L1.w);
      const u64 l0a = sub2(one2, l1a), l0b = sub2(one2, l1b);
      const bool s1 = (selm >> (4 * gi + 1)) & 1u, s2 = (selm >> (4 * gi + 2)) & 1u, s3 = (selm >> (4 * gi + 3)) & 1u;
      float y0[K], y1[K], y2[K];
      if constexpr (K == 1) {
        y0[0] = lds_f32(a); y1[0] = lds_f32(a + 4u); y2[0] = lds_f32(a + 8u);
      } else if constexpr (K == 2) {
        const float2 v0 = lds_f2(a), v1 = lds_f2(a + 8u), v2 = lds_f2(a + 16u);
        y0[0] = v0.x; y0[K - 1] = v0.y; y1[0] = v1.x; y1[K - 1] = v1.y; y2[0] = v2.x; y2[K - 1] = v2.y;
      } else {
        const float4 v0 = lds_f4(a), v1 = lds_f4(a + 16u), v2 = lds_f4(a + 32u);
        y0[0] = v0.x; y0[1 % K] = v0.y; y0[K - 1] = v0.z; y1[0] = v1.x; y1[1 % K] = v1.y; y1[K - 1] = v1.z;
        y2[0] = v2.x; y2[1 % K] = v2.y; y2[K - 1] = v2.z;
      }
#pragma unroll
      for (int k = 0; k < K; k++) {
        const float a1 = s1 ? y1[k] : y0[k], b1 = s1 ? y2[k] : y1[k];
        const float a2 = s2 ? y1[k] : y0[k], b2 = s2 ? y2[k] : y1[k];
        const float a3 = s3 ? y1[k] : y0[k], b3 = s3 ? y2[k] : y1[k];
        H[k][0] = fma2(l0a, pack2(y0[k], a1), mul2(l1a, pack2(y1[k], b1)));
        H[k][NP - 1] = fma2(l0b, pack2(a2, a3), mul2(l1b, pack2(b2, b3)));
      }
      return;
    }
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

  uint32_t rt = rowtab_s + RS * ys, ro_a = rowoff_s + 8u * G * ys;
  uint32_t lab_a = lab_s + ys * lab_stride + x;
  // direct mode: byte masks are fetched 4 rows ahead of their use (rows past the strip are clamped, never out of bounds)
  constexpr int PD = 4;
  const long long tbase = (long long)n * p.T_h * T_w + x;
  const uint8_t* bgt = (!LSM && has_bg) ? p.bg + tbase : nullptr;
  const uint8_t* gtt = (!LSM && do_conf) ? p.gt + tbase : nullptr;
  uint8_t* lbp = (!LSM && has_label) ? p.label_out + tbase + (long long)ys * T_w : nullptr;
  unsigned int bgq[PD], gtq[PD];
#pragma unroll
  for (int j = 0; j < PD; j++) {
    bgq[j] = 0; gtq[j] = 0;
    if (!LSM) {
      const int r = min(ys + j, ye - 1) * T_w;
      if (has_bg) bgq[j] = ldg_px<NP>(bgt + r);
      if (do_conf) gtq[j] = ldg_px<NP>(gtt + r);
    }
  }
  const unsigned int m4 = 0x01010101u * (unsigned)p.bg_match, bgl4 = 0x01010101u * (unsigned)p.bg_label;

  // row-table entry: vertical weights (-l0, duplicated for the packed fma) + "source rows moved" flags
  auto load_rowtab = [&](uint32_t a, u64 (&w)[G], unsigned int& flags) {
    constexpr int CH = (G + 2) / 2;  // 16-byte chunks: G weight slots + one flag slot
    u64 sl[2 * CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { const ulonglong2 t = lds_u64x2(a + 16u * c); sl[2 * c] = t.x; sl[2 * c + 1] = t.y; }
#pragma unroll
    for (int gi = 0; gi < G; gi++) w[gi] = sl[gi];
    flags = (unsigned int)sl[G];
  };
#pragma unroll 1
  for (int yl = ys; yl < ye; yl++) {
    u64 w[G];
    unsigned int flags;
    load_rowtab(rt, w, flags);
    rt += RS;
    unsigned int bg4 = 0, gt4 = 0;
    if (!LSM) {
      bg4 = bgq[0]; gt4 = gtq[0];
#pragma unroll
      for (int j = 0; j + 1 < PD; j++) { bgq[j] = bgq[j + 1]; gtq[j] = gtq[j + 1]; }
      const int r = min(yl + PD, ye - 1) * T_w;
      if (has_bg) bgq[PD - 1] = ldg_px<NP>(bgt + r);
      if (do_conf) gtq[PD - 1] = ldg_px<NP>(gtt + r);
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
    ro_a += 8u * G;

    u64 acc[K][NP];
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int q = 0; q < NP; q++) {
        u64 a = base[k][q];
#pragma unroll
        for (int gi = 0; gi < G; gi++) a = fma2(w[gi], Dh[gi][k][q], a);
        acc[k][q] = a;
      }
    unsigned int lab4, unc = 0;
    if (!labels_from_diffs<K, NP>(acc, c4, tau, lab4)) {
      unc = uncertain_mask<K, NP, C>(acc, cls, tau);
      if (unc) push_uncertain(ctl, queue, b, yl, x, unc);
    }
    if (LSM) {
      sts_px<NP>(lab_a, lab4);
      lab_a += lab_stride;
    } else {
      if (do_conf) {
        if (NP == 2 && unc == 0) count_word<C>(cnt_lo, cnt_hi, gt4, lab4);
        else {
#pragma unroll
          for (int j = 0; j < 2 * NP; j++) {
            const unsigned int gg = (gt4 >> (8 * j)) & 0xffu;
            if (gg < (unsigned)C && !((unc >> j) & 1u)) count_bin(cnt_lo, cnt_hi, gg * C + ((lab4 >> (8 * j)) & 0xffu), 1u);
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
}

// ---- pre-pass: Y[g][k] = sum over the views of group g of (x[c_{k+1}] - x[c_0]) at low resolution, in the de-augmented
// frame; returns the thread's max |x| (NaN-propagating).  When every group is a (view 2g, view 2g+1) pair -- a scale and its
// flipped twin, the BASELINE layout -- both views are read in one sweep; otherwise the first view of a group writes and the
// others add (a thread always revisits its own cells, so the read-modify-write needs no synchronisation).
// Map rows carry two pad columns that replicate the last one (the right-edge clamp of the 3-tap loads of filter_rows<T4>);
// IL: the K maps of a group are interleaved [row][column][k], otherwise planar [k][row][column].
template <int C, int V, int G, int K, bool IL>
__device__ __forceinline__ float filter_prepass(const FilterGeom& g, const uint32_t (&vb)[V], const int (&cls)[C], uint32_t ymap_s, int tid, int nt) {
  float mxf = 0.f;
  constexpr uint32_t ES = IL ? 4u * (K == 3 ? 4 : K) : 4u;  // bytes between horizontally adjacent cells (interleaved K = 3: padded to 4 floats)
  if constexpr (V == 2 * G) {
#pragma unroll
    for (int gi = 0; gi < G; gi++) {
      constexpr int dummy = 0; (void)dummy;
      const int va = 2 * gi < V ? 2 * gi : 0, vc = 2 * gi + 1 < V ? 2 * gi + 1 : 0;
      const int wo = g.g_wo[gi], cells = g.g_ho[gi] * wo;
      const unsigned int inv = g.g_inv[gi];
      const uint32_t ym = ymap_s + g.g_ybytes[gi];
      const int ra = g.vrow[va], ca = g.vcol[va], rb = g.vrow[vc], cb = g.vcol[vc];
      const uint32_t basea = vb[va] + g.vbase[va] + cls[0] * g.plane_bytes[va];
      const uint32_t baseb = vb[vc] + g.vbase[vc] + cls[0] * g.plane_bytes[vc];
      int dqa[K], dqb[K];
#pragma unroll
      for (int q = 0; q < K; q++) { dqa[q] = (cls[q + 1] - cls[0]) * g.plane_bytes[va]; dqb[q] = (cls[q + 1] - cls[0]) * g.plane_bytes[vc]; }
#pragma unroll 4
      for (int idx = tid; idx < cells; idx += nt) {
        const int i = (int)__umulhi((unsigned)idx, inv), j = idx - i * wo;
        const uint32_t aa = basea + i * ra + j * ca, ab = baseb + i * rb + j * cb;
        const float x0a = lds_f32(aa), x0b = lds_f32(ab);
        mxf = max_nan(max_nan(mxf, fabsf(x0a)), fabsf(x0b));
        uint32_t ya = ym + ES * (uint32_t)(i * (wo + 2) + j);
        const bool last = j == wo - 1;
#pragma unroll
        for (int q = 0; q < K; q++) {
          const float xa = lds_f32(aa + dqa[q]), xb = lds_f32(ab + dqb[q]);
          mxf = max_nan(max_nan(mxf, fabsf(xa)), fabsf(xb));
          const float yv = __fadd_rn(__fsub_rn(xa, x0a), __fsub_rn(xb, x0b));
          sts_f32(ya, yv);
          if (last) { sts_f32(ya + ES, yv); sts_f32(ya + 2u * ES, yv); }
          ya += IL ? 4u : (uint32_t)g.g_mapbytes[gi];
        }
      }
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int gi = g.group_of[v];
      const bool first = g.first_in_group[v] != 0;
      const int wo = g.g_wo[gi], cells = g.g_ho[gi] * wo, vrow = g.vrow[v], vcol = g.vcol[v];
      const unsigned int inv = g.g_inv[gi];
      const uint32_t ym = ymap_s + g.g_ybytes[gi];
      const uint32_t base0 = vb[v] + g.vbase[v] + cls[0] * g.plane_bytes[v];
      int dq[K];
#pragma unroll
      for (int q = 0; q < K; q++) dq[q] = (cls[q + 1] - cls[0]) * g.plane_bytes[v];
      for (int idx = tid; idx < cells; idx += nt) {
        const int i = (int)__umulhi((unsigned)idx, inv), j = idx - i * wo;
        const uint32_t a = base0 + i * vrow + j * vcol;
        const float x0 = lds_f32(a);
        mxf = max_nan(mxf, fabsf(x0));
        uint32_t ya = ym + ES * (uint32_t)(i * (wo + 2) + j);
        const bool last = j == wo - 1;
#pragma unroll
        for (int q = 0; q < K; q++) {
          const float xq = lds_f32(a + dq[q]);
          mxf = max_nan(mxf, fabsf(xq));
          float t = __fsub_rn(xq, x0);
          if (!first) t = __fadd_rn(lds_f32(ya), t);
          sts_f32(ya, t);
          if (last) { sts_f32(ya + ES, t); sts_f32(ya + 2u * ES, t); }
          ya += IL ? 4u : (uint32_t)g.g_mapbytes[gi];
        }
      }
    }
  }
  return mxf;
}

template <int C, int V, int G, int F, int NP, bool LSM, int NB>
__device__ __forceinline__ void fuse_filter_body(const FuseParams& p, const FilterGeom& g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FCtl* ctl = reinterpret_cast<FCtl*>(smem_raw + g.ctl_off);
  int2* rowoff = reinterpret_cast<int2*>(smem_raw + g.rowoff_off);        // [T_h][G] byte offsets of rows i0, i1 inside a map
  int4* colA = reinterpret_cast<int4*>(smem_raw + g.cola_off);            // [G][GXP] byte offsets {j0,j1 of col 0; j0,j1 of col 1}
  float4* colB = reinterpret_cast<float4*>(smem_raw + g.colb_off);        // [G][GXP] {l0 col0, l0 col1, l1 col0, l1 col1}
  float4* col4 = reinterpret_cast<float4*>(smem_raw + g.col4_off);        // [G][GX] l1 of the thread's 4 columns (T4)
  uint32_t* col4i = reinterpret_cast<uint32_t*>(smem_raw + g.col4i_off);  // [G][GX] (4 * j0 of column 0) | sel << 16 (T4)
  float4* lowrow = reinterpret_cast<float4*>(smem_raw + g.lowrow_off);    // [low_h][V] {l0, l1, byte offset i0, byte offset i1} in the RAW view
  float4* lowcol = reinterpret_cast<float4*>(smem_raw + g.lowcol_off);    // [low_w][V]
  float* ymap = reinterpret_cast<float*>(smem_raw + g.ymap_off);          // [G][C-1][ho][wo] difference maps
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + g.queue_off);  // [kFQueueCap] (y << 16) | x
  uint8_t* labsm = smem_raw + (LSM ? g.lab_off : 0);                      // [T_h][T_w] labels of the current tile (LSM)
  float* vsm = reinterpret_cast<float*>(smem_raw + g.views_off);          // 2 staging buffers

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int T_h = p.T_h, T_w = p.T_w;
  constexpr bool RT = F < 0;
  constexpr bool T4 = NP == 2;          // 4-column tables + interleaved maps
  constexpr bool PAIRTAB = !T4;
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
    mbar_init(&ctl->empty[0], g.cwarps + (need_low ? g.aux : 0));
    mbar_init(&ctl->empty[1], g.cwarps + (need_low ? g.aux : 0));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->maxbits[0] = ctl->maxbits[1] = 0u;
    ctl->qcount[0] = ctl->qcount[1] = 0u;
    ctl->lownext[0] = ctl->lownext[1] = 0u;
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
      rowoff[y * G + gi] = make_int2(4 * L.i0 * (g.g_wo[gi] + 2), 4 * L.i1 * (g.g_wo[gi] + 2));
      if (!strip_start) {  // first row of a strip: both source rows are loaded before the row loop
        const Lerp P = pisto_src_index(g.g_scale_h[gi], y - 1, g.g_ho[gi], false);
        if (P.i0 != L.i0 || P.i1 != L.i1) f |= 1u << gi;  // up-sampling: the pair moves down by exactly one row
      }
    }
    reinterpret_cast<uint2*>(row)[G] = make_uint2(f, 0u);
  }
  if (PAIRTAB)
    for (int i = tid; i < G * g.GXP; i += nthreads) {
      const int gi = i / g.GXP, gx = i - gi * g.GXP;
      const Lerp L0 = pisto_src_index(g.g_scale_w[gi], 2 * gx, g.g_wo[gi], g.g_same_w[gi]);
      const Lerp L1 = pisto_src_index(g.g_scale_w[gi], 2 * gx + 1, g.g_wo[gi], g.g_same_w[gi]);
      colA[i] = make_int4(4 * L0.i0, 4 * L0.i1, 4 * L1.i0, 4 * L1.i1);
      colB[i] = make_float4(L0.l0, L1.l0, L0.l1, L1.l1);
    }
  if (T4)
    for (int i = tid; i < G * g.GX; i += nthreads) {
      const int gi = i / g.GX, gx = i - gi * g.GX;
      Lerp L[4];
#pragma unroll
      for (int c = 0; c < 4; c++) L[c] = pisto_src_index(g.g_scale_w[gi], 4 * gx + c, g.g_wo[gi], g.g_same_w[gi]);
      col4[i] = make_float4(L[0].l1, L[1].l1, L[2].l1, L[3].l1);
      unsigned int sel = 0;
#pragma unroll
      for (int c = 1; c < 4; c++) sel |= (unsigned)(L[c].i0 - L[0].i0) << c;  // 0 or 1 (make_filter_geom checked)
      col4i[i] = (unsigned)(4 * L[0].i0) | (sel << 16);
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

  const int ncomp = g.cwarps * 32;  // compute threads, then g.aux export warps; the last warp of the CTA is the producer
  if (tid >= ncomp + 32 * g.aux) {
    // ===== producer warp: claims tiles, publishes their ids, fetches their views (TMA) one tile ahead of the compute warps
    if (tid == ncomp + 32 * g.aux) {
      const long long tile_px = (long long)T_h * T_w;
      // the tile id (global atomic) and its presence vector are fetched one step ahead, so that only the TMA itself sits
      // between a staging buffer becoming free and its refill
      int next = atomicAdd(g.counter, 1);
      TilePresence next_tp; next_tp.bits = 0u; next_tp.single = -1;
      if (next < p.N) next_tp = pisto_tile_presence(p, next);
      bool next_views = next < p.N && (need_low || next_tp.single < 0);
      for (int k = 0;; k++) {
        const int b = NB == 2 ? (k & 1) : 0;
        if (k >= NB) mbar_wait_sleep(&ctl->empty[b], NB == 2 ? (((k >> 1) - 1) & 1) : ((k - 1) & 1));  // every warp is done with buffer b
        const int tile = next < p.N ? next : -1;
        ctl->tile[b] = tile;
        ctl->pres_bits[b] = next_tp.bits;
        ctl->pres_single[b] = next_tp.single;
        ctl->lownext[b] = 0u;
        if (tile >= 0 && next_views) issue_tile(tile, b);
        else mbar_arrive(&ctl->full[b]);
        if (tile < 0) break;
        // pull the tile's byte masks into L2 ahead of their use
        if (tile_px % 16 == 0) {
          if (has_bg && ((uintptr_t)p.bg & 15) == 0) bulk_prefetch_l2(p.bg + tile * tile_px, (uint32_t)tile_px);
          if (do_conf && ((uintptr_t)p.gt & 15) == 0) bulk_prefetch_l2(p.gt + tile * tile_px, (uint32_t)tile_px);
        }
        next = atomicAdd(g.counter, 1);
        if (next < p.N) next_tp = pisto_tile_presence(p, next);
        next_views = next < p.N && (need_low || next_tp.single < 0);
      }
    }
    return;
  }

  // ---- the 32x32 logits of tile n (infer_pseudo_masks.py:126), exact and in the reference's operation order.  Low-resolution
  // rows are claimed one at a time from a shared counter by whichever warp has nothing better to do: the export warps all the
  // time, the compute warps once their own work on the tile is finished.  Lane = low-resolution column; per (view, class) the
  // two bracketing source rows are interpolated horizontally once (Ta, Tb) and blended vertically.
  auto export_rows = [&](int n, int b, const uint32_t (&vb)[V]) {
    const int lane = tid & 31;
    const long long lplane = (long long)p.low_h * p.low_w;
    for (;;) {
      int ly = 0;
      if (lane == 0) ly = (int)atomicAdd(&ctl->lownext[b], 1u);
      ly = __shfl_sync(0xffffffffu, ly, 0);
      if (ly >= p.low_h) break;
      for (int lxb = 0; lxb < p.low_w; lxb += 32) {
        const int lx = min(lxb + lane, p.low_w - 1);
        float a[C];
#pragma unroll
        for (int v = 0; v < V; v++) {
          const float4 R = lowrow[ly * V + v], Q = lowcol[lx * V + v];
          const uint32_t r0 = vb[v] + __float_as_int(R.z), r1 = vb[v] + __float_as_int(R.w);
          const uint32_t c0 = __float_as_int(Q.z), c1 = __float_as_int(Q.w);
#pragma unroll
          for (int c = 0; c < C; c++) {
            const uint32_t pl = c * g.plane_bytes[v];
            const float ta = __fmaf_rn(Q.x, lds_f32(r0 + pl + c0), __fmul_rn(Q.y, lds_f32(r0 + pl + c1)));
            const float tb = __fmaf_rn(Q.x, lds_f32(r1 + pl + c0), __fmul_rn(Q.y, lds_f32(r1 + pl + c1)));
            const float u = __fmaf_rn(R.x, ta, __fmul_rn(R.y, tb));
            a[c] = (v == 0) ? u : __fadd_rn(a[c], u);
          }
        }
        if (lxb + lane < p.low_w) {
          float* outp = p.lowres_out + (long long)n * C * lplane + (long long)ly * p.low_w + lxb + lane;
#pragma unroll
          for (int c = 0; c < C; c++) outp[c * lplane] = pisto_div_views(a[c], p.dec);
        }
      }
    }
  };

  if (tid >= ncomp) {
    // ===== export warps
    if (!need_low) return;
    for (int k = 0;; k++) {
      const int b = k & 1;
      mbar_wait(&ctl->full[b], (k >> 1) & 1);
      const int n = ctl->tile[b];
      if (n < 0) break;
      uint32_t vb[V];
#pragma unroll
      for (int v = 0; v < V; v++) {
        const ViewDev& vw = p.view[v];
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
        vb[v] = smem_u32(vsm + b * g.buf_floats + g.view_off[v]) + sh;
      }
#ifndef PISTO_X_SKIP_EXPORT
      export_rows(n, b, vb);
#endif
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&ctl->empty[b]);
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
  const uint32_t colA_t = smem_u32(colA) + 16u * NP * grp, colB_t = smem_u32(colB) + 16u * NP * grp;  // pair tables
  const uint32_t col4_t = smem_u32(col4) + 16u * grp, col4i_t = smem_u32(col4i) + 4u * grp;              // 4-column tables
  const uint32_t lab_s = smem_u32(labsm);
  const int nt = ncomp;  // cooperative loops below run over the compute threads only
  const long long tpx = (long long)T_h * T_w;

  for (int k = 0;; k++) {
    const int b = k & 1;                          // slot of the per-tile scratch (max, queue)
    const int sb = NB == 2 ? b : 0;           // staging buffer
    mbar_wait(&ctl->full[sb], NB == 2 ? ((k >> 1) & 1) : (k & 1));  // tile id published, views (if any) landed
    const int n = ctl->tile[sb];
    if (n < 0) break;
    TilePresence tp = pisto_tile_presence(p, n);   // no present vector: constants
    if (p.present) {                               // published by the producer with the tile id: no global-memory latency here
      tp.bits = ctl->pres_bits[sb]; tp.single = ctl->pres_single[sb];
    }
    const bool multi = tp.single < 0;
    // shared-memory byte address of view v's data in this tile's staging buffer (incl. the 0..3-float alignment shift)
    uint32_t vb[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const ViewDev& vw = p.view[v];
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
      vb[v] = smem_u32(vsm + sb * g.buf_floats + g.view_off[v]) + sh;
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
#ifdef PISTO_X_SKIP_PREPASS
    if (false) {
#else
    if (multi && P >= 2) {
#endif
      float mxf;
      if (P == 2) mxf = filter_prepass<C, V, G, 1, T4>(g, vb, cls, ymap_s, tid, nt);
      else if (P == 3) mxf = filter_prepass<C, V, G, 2, T4>(g, vb, cls, ymap_s, tid, nt);
      else mxf = filter_prepass<C, V, G, (C >= 4 ? 3 : 1), T4>(g, vb, cls, ymap_s, tid, nt);
      unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));  // NaN (0x7fffffff) > Inf > finite
      if ((tid & 31) == 0) atomicMax(&ctl->maxbits[b], mx);
    }
    bar_sync(1, ncomp);  // difference maps + max visible; every thread has left the previous tile
    if (NB == 1 && (tid & 31) == 0) mbar_arrive(&ctl->empty[0]);  // single staging buffer: the raw views are not read again
    if (tid == 0) ctl->qcount[b ^ 1] = 0u;  // the previous tile's queue has been read by everyone; its slot is next pushed to after the next bar_sync

    bool exact_all = false;
    if (multi) {
      float tau = 0.f;
      if (P >= 2) {
        const float A = __fmul_rn((float)V, __uint_as_float(ctl->maxbits[b]));
        tau = __fmaf_rn(A, g.tau_coef, g.tau_abs);
        if (!(A < 5e8f)) exact_all = true;  // non-finite or absurd magnitudes: follow the reference everywhere
      } else {
        exact_all = true;                   // empty (or one-class NEG_INF) presence vector
      }
#ifdef PISTO_X_SKIP_ROWS
      if (false) {
#else
      if (!exact_all && worker && ys < ye) {
#endif
        if (P == 2) filter_rows<C, G, F, NP, 1, LSM, T4 ? 4 : 0>(p, g, ctl, queue, b, rowtab_s, rowoff_s, T4 ? col4_t : colA_t, T4 ? col4i_t : colB_t, ymap_s, lab_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
        else if (P == 3) filter_rows<C, G, F, NP, 2, LSM, T4 ? 4 : 0>(p, g, ctl, queue, b, rowtab_s, rowoff_s, T4 ? col4_t : colA_t, T4 ? col4i_t : colB_t, ymap_s, lab_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
        else if (C >= 4 && P == 4) {
          constexpr int K3 = C >= 4 ? 3 : 1;
          filter_rows<C, G, F, NP, K3, LSM, T4 ? 4 : 0>(p, g, ctl, queue, b, rowtab_s, rowoff_s, T4 ? col4_t : colA_t, T4 ? col4i_t : colB_t, ymap_s, lab_s, n, x, ys, ye, cls, tau, cnt_lo, cnt_hi);
        }
      }
      bar_sync(1, ncomp);  // every strip done: the queue is complete
      if (tid == 0) ctl->maxbits[b] = 0u;  // read by everyone before this barrier; next written two tiles from now
    }

    // The byte masks of the vector pass' first batch are requested now, so that their HBM / L2 latency is covered by the exact
    // pass and its barrier instead of being paid at the start of the vector pass.
    constexpr int UN = 4;  // independent 16-byte loads in flight per thread
    const long long vbase_px = (long long)n * tpx;
    const bool vec_ok = (tpx % 16 == 0) && ((((uintptr_t)p.bg | (uintptr_t)p.gt | (uintptr_t)p.label_out) & 15) == 0);
    const int nvec = (vec_ok && (LSM || !multi)) ? (int)(tpx / 16) : 0;
    uint4 bgn[UN], gn[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int i = tid + u * nt;
      gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);  // out of range: never counted
      bgn[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < nvec) {
        if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + vbase_px) + i);
        if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + vbase_px) + i);
      }
    }

    // ---- exact pass: queued pixels (or the whole tile), operation by operation as the reference ----------------------
    if (multi) {
      const unsigned int nq = ctl->qcount[b];
      if (tid == 0 && g.stats) {  // data-dependence record: queued pixels, whole-tile fallbacks (pisto_filter_stats)
        atomicAdd(&g.stats[1], 1ull);
        if (!exact_all && nq <= (unsigned)kFQueueCap) atomicAdd(&g.stats[2], (unsigned long long)nq); else atomicAdd(&g.stats[3], 1ull);
      }
      if (nq > (unsigned)kFQueueCap && !exact_all) { exact_all = true; cnt_lo = cnt_hi = 0; }  // overflow: redo the whole tile
      // A queued pixel is evaluated by a whole warp: lane l computes the bilinear sample of (view l / C, class l % C), the
      // sums are then formed in view order through shuffles -- the latency of one pixel is that of one sample, and the few
      // queued pixels of a tile are spread over all warps.  (The whole-tile fallback keeps one pixel per thread.)
      const bool warp_coop = !exact_all && V * C <= 32;
      const int nfix = exact_all ? T_h * T_w : (int)nq;
      const int lane = tid & 31;
      for (int j = warp_coop ? (tid >> 5) : tid; j < nfix; j += warp_coop ? g.cwarps : nt) {
        float a[C];
        int yy, xx;
        if (exact_all) { yy = j / T_w; xx = j - yy * T_w; }
        else { const uint32_t e = queue[j]; yy = (int)(e >> 16); xx = (int)(e & 0xffffu); }
        if (warp_coop) {
          float o = 0.f;
          if (lane < V * C) {
            const int v = lane / C, c = lane - v * C;
            const ViewDev& vw = p.view[v];
            const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, vw.same_h);
            const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, vw.same_w);
            const int pl = g.vbase[v] + c * g.plane_bytes[v];
            const int r0 = pl + Ly.i0 * g.vrow[v], r1 = pl + Ly.i1 * g.vrow[v];
            const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
            float x00, x01, x10, x11;
            if (NB == 2) {
              const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(vw.logits + (long long)n * vw.tile_stride) & 12u);
              const uint32_t base = smem_u32(vsm + sb * g.buf_floats + g.view_off[v]) + sh;
              x00 = lds_f32(base + r0 + c0); x01 = lds_f32(base + r0 + c1); x10 = lds_f32(base + r1 + c0); x11 = lds_f32(base + r1 + c1);
            } else {
              const float* gsrc = vw.logits + (long long)n * vw.tile_stride;
              x00 = __ldg(gsrc + ((r0 + c0) >> 2)); x01 = __ldg(gsrc + ((r0 + c1) >> 2));
              x10 = __ldg(gsrc + ((r1 + c0) >> 2)); x11 = __ldg(gsrc + ((r1 + c1) >> 2));
            }
            const float h0 = __fmaf_rn(Lx.l0, x00, __fmul_rn(Lx.l1, x01));
            const float h1 = __fmaf_rn(Lx.l0, x10, __fmul_rn(Lx.l1, x11));
            o = __fmaf_rn(Ly.l0, h0, __fmul_rn(Ly.l1, h1));
          }
#pragma unroll
          for (int c = 0; c < C; c++) {
            a[c] = __shfl_sync(0xffffffffu, o, c);
#pragma unroll
            for (int v = 1; v < V; v++) a[c] = __fadd_rn(a[c], __shfl_sync(0xffffffffu, o, (v * C + c) & 31));
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const ViewDev& vw = p.view[v];
            const Lerp Ly = pisto_src_index(vw.scale_h, yy, vw.map.ho, vw.same_h);
            const Lerp Lx = pisto_src_index(vw.scale_w, xx, vw.map.wo, vw.same_w);
            const int r0 = g.vbase[v] + Ly.i0 * g.vrow[v], r1 = g.vbase[v] + Ly.i1 * g.vrow[v];
            const int c0 = Lx.i0 * g.vcol[v], c1 = Lx.i1 * g.vcol[v];
            const float* gsrc = vw.logits + (long long)n * vw.tile_stride;   // single staging buffer: already refilled, read L2 / HBM
#pragma unroll
            for (int c = 0; c < C; c++) {
              const int pl = c * g.plane_bytes[v];
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
        }
        const int lab = pisto_decide<C>(a, tp.bits, p.dec, false, nullptr);
        if (!warp_coop || lane == 0) {
          if (LSM) {
            labsm[yy * T_w + xx] = (uint8_t)lab;
          } else {
            const long long pix = (long long)n * tpx + yy * T_w + xx;
            if (do_conf) {
              const unsigned int gg = p.gt[pix];
              if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + lab], 1u);
            }
            if (has_label) {
              unsigned int o2 = (unsigned)lab;
              if (has_bg && p.bg[pix] == (uint8_t)p.bg_match) o2 = (unsigned)p.bg_label;
              p.label_out[pix] = (uint8_t)o2;
            }
          }
        }
      }
    }
    if (LSM && multi) bar_sync(1, ncomp);  // label tile complete

    // ---- vector pass: confusion, background overwrite, 16-byte label stores (single-label tiles: constant label) -----
#ifdef PISTO_X_SKIP_VECTOR
    if (false) {
#else
    if (LSM || !multi) {
#endif
      const long long base = vbase_px;
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
          bgv[u] = bgn[u]; gv[u] = gn[u];           // this batch was requested one step ago
          lv[u] = make_uint4(labc, labc, labc, labc);
          if (i < nvec && LSM && multi) { const int4 t = lds_i4(lab_s + 16u * i); lv[u] = make_uint4(t.x, t.y, t.z, t.w); }
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {              // request the next batch
          const int i = i0 + (UN + u) * nt;
          gn[u] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (i < nvec) {
            if (has_bg) bgn[u] = __ldg(reinterpret_cast<const uint4*>(p.bg + base) + i);
            if (do_conf) gn[u] = __ldg(reinterpret_cast<const uint4*>(p.gt + base) + i);
          }
        }
        if (do_conf) {  // two vectors = 32 pixels per bit-sliced count
#pragma unroll
          for (int u = 0; u < UN; u += 2) {
            const unsigned int gw[8] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w, gv[u + 1].x, gv[u + 1].y, gv[u + 1].z, gv[u + 1].w};
            const unsigned int lw8[8] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w, lv[u + 1].x, lv[u + 1].y, lv[u + 1].z, lv[u + 1].w};
            bitslice_count<C>(gw, lw8, cnt32);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const int i = i0 + u * nt;
          if (i < nvec) {
            const unsigned int lw[4] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w};
            if (has_label) {
              uint4 o = lv[u];
              if (has_bg) {
                const unsigned int bw[4] = {bgv[u].x, bgv[u].y, bgv[u].z, bgv[u].w};
                unsigned int ow[4];
#pragma unroll
                for (int q = 0; q < 4; q++) { const unsigned int eq = __vcmpeq4(bw[q], m4); ow[q] = (bgl4 & eq) | (lw[q] & ~eq); }
                o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              }
              reinterpret_cast<uint4*>(p.label_out + base)[i] = o;
            }
          }
        }
      }
      if (do_conf) {
#pragma unroll
        for (int bn = 0; bn < BINS; bn++) {
          const unsigned int cv = __reduce_add_sync(0xffffffffu, cnt32[bn]);
          if ((tid & 31) == 0 && cv) atomicAdd(&ctl->hist[bn], cv);
        }
      }
      for (int i = nvec * 16 + tid; i < (int)tpx; i += nt) {  // unaligned / ragged shapes
        const unsigned int lab = (LSM && multi) ? labsm[i] : (unsigned)tp.single;
        unsigned int o = lab;
        if (has_bg && p.bg[base + i] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
        if (do_conf) {
          const unsigned int gg = p.gt[base + i];
          if (gg < (unsigned)C) atomicAdd(&ctl->hist[gg * C + lab], 1u);
        }
        if (has_label) p.label_out[base + i] = (uint8_t)o;
      }
    }
#ifndef PISTO_X_SKIP_EXPORT
    if (need_low) export_rows(n, sb, vb);  // whatever the export warps have not got to yet
#endif
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
    if (NB == 2 && (tid & 31) == 0) mbar_arrive(&ctl->empty[sb]);  // this warp is done with staging buffer sb
  }
  if (do_conf) {
    bar_sync(1, ncomp);
    for (int i = tid; i < BINS; i += nt)
      if (ctl->hist[i]) atomicAdd(&p.conf[i], (unsigned long long)ctl->hist[i]);
  }
}

// two entry points over the same body: 512 threads x 128 registers, and (C = 4: three difference fields for four columns)
// 384 threads x up to 168 registers
template <int C, int V, int G, int F, int NP, bool LSM, int NB>
__global__ void __launch_bounds__(kFMaxThreads, 1) fuse_filter_kernel(const __grid_constant__ FuseParams p,
                                                                      const __grid_constant__ FilterGeom g) {
  fuse_filter_body<C, V, G, F, NP, LSM, NB>(p, g);
}
template <int C, int V, int G, int F, int NP, bool LSM, int NB>
__global__ void __launch_bounds__(384, 1) fuse_filter_kernel_wide(const __grid_constant__ FuseParams p, const __grid_constant__ FilterGeom g) {
  fuse_filter_body<C, V, G, F, NP, LSM, NB>(p, g);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool make_filter_geom(const pisto_ctx* h, const FuseParams& p, int NP, int G_expected, bool lsm, int nbuf, int max_threads, int aux, FilterGeom* g) {
  const bool t4 = NP == 2, pairtab = !t4;
  memset(g, 0, sizeof(*g));
  g->nbuf = nbuf;
  if (nbuf == 1 && p.lowres_out && p.low_fh > 0) return false;  // the export reads the staged views for the whole tile
  if (p.T_w % (2 * NP)) return false;
  const int GX = p.T_w / (2 * NP);
  g->aux = aux;
  if (GX > max_threads - 32 - 32 * aux) return false;  // one producer warp, aux export warps
  int S = (max_threads - 32 - 32 * aux) / GX;
  if (S > p.T_h) S = p.T_h;
  if (S > 32) S = 32;
  // scale groups: views with the same de-augmented size share their interpolation weights
  int G = 0, nmax = 0, cnt[kFMaxGroups] = {0};
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    if (vw.map.ho >= p.T_h) return false;  // same-size / down-sampling rows: the source-row pair does not move by exactly one
    if (vw.map.ho * vw.map.wo >= 65536) return false;
    int gi = -1;
    for (int q = 0; q < G; q++)
      if (g->g_ho[q] == vw.map.ho && g->g_wo[q] == vw.map.wo) gi = q;
    if (gi < 0) {
      if (G == kFMaxGroups) return false;
      gi = G++;
      g->g_ho[gi] = vw.map.ho; g->g_wo[gi] = vw.map.wo; g->g_same_w[gi] = vw.same_w;
      g->g_scale_h[gi] = vw.scale_h; g->g_scale_w[gi] = vw.scale_w;
      g->g_inv[gi] = (unsigned int)(((1ull << 32) + vw.map.wo - 1) / vw.map.wo);
    }
    g->group_of[v] = gi;
    g->first_in_group[v] = cnt[gi] == 0;
    if (++cnt[gi] > nmax) nmax = cnt[gi];
    g->plane_bytes[v] = 4 * vw.h * vw.w;
    g->vbase[v] = 4 * (vw.map.a0 * vw.w + vw.map.b0);
    g->vrow[v] = 4 * (vw.map.ai * vw.w + vw.map.bi);
    g->vcol[v] = 4 * (vw.map.aj * vw.w + vw.map.bj);
  }
  if (G != G_expected) return false;
  if (t4)  // the 4 columns of a thread must lie within two adjacent source cells (3-tap loads), offsets must fit 16 bits
    for (int q = 0; q < G; q++) {
      if (4 * 4 * (g->g_wo[q] + 2) > 65535) return false;
      for (int x = 0; x < p.T_w; x += 4) {
        const Lerp L0 = pisto_src_index(g->g_scale_w[q], x, g->g_wo[q], g->g_same_w[q]);
        for (int c = 0; c < 4; c++) {
          const Lerp L = pisto_src_index(g->g_scale_w[q], x + c, g->g_wo[q], g->g_same_w[q]);
          if (L.i0 < L0.i0 || L.i0 > L0.i0 + 1) return false;
          if (L.i1 != L.i0 && L.i1 != L.i0 + 1) return false;
          if (L.i1 == L.i0 && L.i0 != g->g_wo[q] - 1 && L.l1 != 0.f) return false;  // only the right-edge clamp repeats a column
        }
      }
    }
  if (p.V == 2 * G)  // the paired pre-pass assumes views (2g, 2g+1) form group g
    for (int v = 0; v < p.V; v++)
      if (g->group_of[v] != v / 2) return false;
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
    int best = 0;
    for (int s = 1; s <= S; s++)
      if ((p.T_h / period) % s == 0) best = s;
    if (best * 2 > S) S = best;
  }
  g->GX = GX; g->S = S; g->GXP = p.T_w / 2; g->lab_stride = p.T_w;
  g->cwarps = (GX * S + 31) / 32;
  g->threads = g->cwarps * 32 + 32 * aux + 32;
  int rps = 0;
  for (int q = 0; q <= S; q++) g->strip_y0[q] = (int)((long long)p.T_h * q / S);
  for (int q = 0; q < S; q++) rps = max(rps, g->strip_y0[q + 1] - g->strip_y0[q]);
  if (rps * 2 * NP > 255) return false;  // packed 8-bit confusion counters (direct mode)
  int fl = 0;
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    g->view_off[v] = fl;
    fl += (p.C * vw.h * vw.w + 3 /* alignment shift */ + 3 /* tail */ + 3) & ~3;
  }
  g->buf_floats = fl;
  // decision threshold: the interpolated differences and the exact view sums each differ from the real-number
  // value by at most cE/2-ish * 2^-24 * A; on top of that the lead must cover the softmax margin 2.4e-7 * |a| + V * 2e-6 of common.cuh
  const float cE = 2.f * nmax + 4.f * G + 2.f * p.V + 20.f;
  g->tau_coef = 2.f * cE * 5.9604645e-8f + 2.5e-7f;
  g->tau_abs = p.dec.margin_abs * 1.01f;
  const int RS = 16 * ((G + 2) / 2);
  const bool low = p.lowres_out && p.low_fh > 0;
  int off = 0;
  g->ctl_off = off; off += (int)((sizeof(FCtl) + 127) & ~127u);
  g->rowtab_off = off; off += RS * p.T_h;
  g->rowoff_off = off; off += 8 * G * p.T_h; off = (off + 15) & ~15;
  g->cola_off = off; off += pairtab ? 16 * G * g->GXP : 0;
  g->colb_off = off; off += pairtab ? 16 * G * g->GXP : 0;
  g->col4_off = off; off += t4 ? 16 * G * GX : 0;
  g->col4i_off = off; off += t4 ? 4 * G * GX : 0; off = (off + 15) & ~15;
  g->lowrow_off = off; off += 16 * p.V * (low ? p.low_h : 0);
  g->lowcol_off = off; off += 16 * p.V * (low ? p.low_w : 0);
  g->ymap_off = off;
  for (int q = 0; q < G; q++) {
    g->g_ybytes[q] = off - g->ymap_off;
    g->g_mapbytes[q] = 4 * g->g_ho[q] * (g->g_wo[q] + 2);  // two pad columns per row
    off += ((t4 && p.C == 4) ? 4 : p.C - 1) * g->g_mapbytes[q];  // interleaved K = 3 cells are padded to 4 floats
    off = (off + 15) & ~15;
  }
  g->queue_off = off; off += 4 * kFQueueCap;
  g->lab_off = off;
  if (lsm) { off += p.T_h * p.T_w; off = (off + 15) & ~15; }
  off = (off + 127) & ~127;
  g->views_off = off; off += nbuf * 4 * fl;
  g->smem_bytes = off;
  return off <= h->smem_optin - 1024;
}

template <int C, int V, int G, int F, int NP, bool LSM, int NB>
int launch_filter(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  FilterGeom g;
  if (!make_filter_geom(h, p, NP, G, LSM, NB, filter_threads<C>(), filter_aux<F>(), &g)) return PISTO_OK;  // not launched: caller falls back
  void (*kern)(FuseParams, FilterGeom);
  if constexpr (C >= 4) kern = fuse_filter_kernel_wide<C, V, G, F, NP, LSM, NB>;
  else kern = fuse_filter_kernel<C, V, G, F, NP, LSM, NB>;
  // Small view sets (one or two stride-8 views: BASELINE config 1) leave more than half of the shared memory unused: two
  // 256-thread CTAs per SM (the register file holds exactly 2 x 256 x 128) let one tile's vector pass -- global-memory
  // latency -- overlap the other's row loop.  A/B knob: PISTO_FILTER_ONE_CTA.
  int ctas = 1;
  if constexpr (C < 4) {
    static const bool one_cta = getenv("PISTO_FILTER_ONE_CTA") != nullptr;
    FilterGeom g2;
    if (!one_cta && 2 * (g.smem_bytes + 1024) <= 233472 && make_filter_geom(h, p, NP, G, LSM, NB, 256, filter_aux<F>(), &g2) &&
        2 * (g2.smem_bytes + 1024) <= 233472 && g2.threads <= 256) {
      g = g2;
      ctas = 2;
    }
  }
  PISTO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
  int sched_slot = 0;
  { const int rc = pisto_sched_acquire(h, st, &g.counter, &sched_slot); if (rc != PISTO_OK) return rc; }
  g.stats = h->stats;
  const int slots = h->sm_count * ctas;
  const int grid = p.N < slots ? p.N : slots;
  kern<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  { const int rc = pisto_sched_release(h, st, sched_slot); if (rc != PISTO_OK) return rc; }
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

// labels staged in shared memory when the tile fits next to the staging buffers, else written directly
template <int C, int V, int G, int F>
static int pisto_launch_filter_f(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  if (np == 2) {
    static const bool force_direct = getenv("PISTO_FILTER_DIRECT") != nullptr;  // A/B knob: labels straight to global memory
    int rc = force_direct ? PISTO_OK : launch_filter<C, V, G, F, 2, true, 2>(h, p, st, launched);
    if (rc != PISTO_OK || *launched) return rc;
    if constexpr ((F & 8) == 0) if (!force_direct) {  // label tile + ONE staging buffer (never with the 32x32 export, which reads the views all tile long)
      rc = launch_filter<C, V, G, F, 2, true, 1>(h, p, st, launched);
      if (rc != PISTO_OK || *launched) return rc;
    }
    rc = launch_filter<C, V, G, F, 2, false, 2>(h, p, st, launched);
    if (rc != PISTO_OK || *launched) return rc;
    // e.g. views up-sampled by less than 3 (a thread's 4 columns span more than two source cells): 2 columns per thread
  }
  return launch_filter<C, V, G, -1, 1, false, 2>(h, p, st, launched);
}

template <int C, int V, int G>
static int pisto_launch_filter_cvg(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched) {
  switch (pisto_filter_flags(p)) {
    case 25: return pisto_launch_filter_f<C, V, G, 25>(h, p, st, np, launched);  // bg + labels + 32x32        (config 2)
    case 19: return pisto_launch_filter_f<C, V, G, 19>(h, p, st, np, launched);  // bg + gt/conf + labels      (config 1)
    case 18: return pisto_launch_filter_f<C, V, G, 18>(h, p, st, np, launched);  // gt/conf + labels           (config 3, mIoUMask.forward)
    default: return pisto_launch_filter_f<C, V, G, -1>(h, p, st, np, launched);
  }
}
