// Full-resolution view fusion: every view already has the output resolution (the segmentation network emits full-resolution
// logits) and only differs by its dihedral test-time augmentation -- exactly what the reference does in
// infer_pseudo_masks.py:96,121 through ttach's SegmentationTTAWrapper(d4_transform(), 'mean'): de-augment each of the 8
// outputs, add them in view order, divide by 8; then get_mask_pred_and_entropy (:69-87) and the 32x32 export (:126).
//
// Nothing is interpolated, so the op streams V*C floats per pixel: HBM-bound.  The only difficulty is that the rot90 / rot270
// views are read TRANSPOSED.  A CTA owns a 32x32 output block; views whose de-augmentation keeps rows as rows (identity,
// flips, rot180) are read straight from global memory (forward or reversed 128-byte rows), the transposed ones are first
// copied block-wise, along THEIR rows, into padded shared-memory tiles and then read transposed from there.  The per-pixel sum
// runs in view order (bit-exact with the sequential merge), followed by pisto_decide / confusion / background / exports
// as in every other fusion kernel.
#include <stdlib.h>
#include <string.h>

#include "fuse_common.cuh"
#include "sm100_prims.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kB = 32;        // block side
constexpr int kPad = kB + 1;  // shared-memory row pitch (floats)

template <int C>
__global__ void __launch_bounds__(kThreads) fuse_fullres_kernel(const __grid_constant__ FuseParams p, int nby, int nbx, int n_transposed) {
  extern __shared__ float tiles[];  // [n_transposed][C][32][33]
  __shared__ unsigned int hist[C * C];
  constexpr int BINS = C * C;
  const bool do_conf = p.conf != nullptr && p.gt != nullptr;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // ty in 0..7: rows ty, ty+8, ty+16, ty+24 of the block
  for (int i = tid; i < BINS; i += kThreads) hist[i] = 0;
  __syncthreads();
  const int T_h = p.T_h, T_w = p.T_w;
  const long long hw = (long long)T_h * T_w;
  const long long items = (long long)p.N * nby * nbx;
  const bool need_low = p.lowres_out != nullptr && p.low_fh > 0;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = (int)(item / (nby * nbx));
    const int rem = (int)(item - (long long)n * nby * nbx);
    const int by = rem / nbx, bx = rem - by * nbx;
    const int y0 = by * kB, x0 = bx * kB;
    const TilePresence tp = pisto_tile_presence(p, n);
    const bool need_scores = tp.single < 0 || p.fused_out || need_low;
    float acc[4][C];
    if (need_scores) {
      // ---- stage the transposed views: thread (ty, tx) copies source rows a(x = x0 + ty + 8r), source column b(y = y0 + tx)
      int slot = 0;
      for (int v = 0; v < p.V; v++) {
        const ViewDev& vw = p.view[v];
        const ViewMap& m = vw.map;
        if (m.ai != 0) continue;  // rows stay rows: read directly below
        const float* src = vw.logits + (long long)n * vw.tile_stride;
        const int yy = y0 + tx;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int xx = x0 + ty + 8 * r;
          if (yy < T_h && xx < T_w) {
            const long long o = (long long)(m.a0 + xx * m.aj) * vw.w + (m.b0 + yy * m.bi);
#pragma unroll
            for (int c = 0; c < C; c++) tiles[((slot * C + c) * kB + (ty + 8 * r)) * kPad + tx] = __ldcs(src + (long long)c * vw.h * vw.w + o);
          }
        }
        slot++;
      }
      if (n_transposed) __syncthreads();
      // ---- sum in view order
      slot = 0;
      for (int v = 0; v < p.V; v++) {
        const ViewDev& vw = p.view[v];
        const ViewMap& m = vw.map;
        const float* src = vw.logits + (long long)n * vw.tile_stride;
        const bool tr = m.ai == 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int yy = y0 + ty + 8 * r, xx = x0 + tx;
          if (yy < T_h && xx < T_w) {
#pragma unroll
            for (int c = 0; c < C; c++) {
              float val;
              if (tr) val = tiles[((slot * C + c) * kB + tx) * kPad + (ty + 8 * r)];
              else val = __ldcs(src + (long long)c * vw.h * vw.w + (long long)(m.a0 + yy * m.ai) * vw.w + (m.b0 + xx * m.bj));
              acc[r][c] = (v == 0) ? val : __fadd_rn(acc[r][c], val);
            }
          }
        }
        if (tr) slot++;
      }
    }
    // ---- per-pixel epilogue
    unsigned long long lo = 0, hi = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int yy = y0 + ty + 8 * r, xx = x0 + tx;
      if (yy < T_h && xx < T_w) {
        const long long rpix = (long long)yy * T_w + xx, pix = (long long)n * hw + rpix;
        if (p.fused_out) {
#pragma unroll
          for (int c = 0; c < C; c++) p.fused_out[((long long)n * C + c) * hw + rpix] = pisto_div_views(acc[r][c], p.dec);
        }
        if (need_low && (yy % p.low_fh == p.low_fh / 2) && (xx % p.low_fw == p.low_fw / 2)) {
#pragma unroll
          for (int c = 0; c < C; c++)
            p.lowres_out[(((long long)n * C + c) * p.low_h + yy / p.low_fh) * p.low_w + xx / p.low_fw] = pisto_div_views(acc[r][c], p.dec);
        }
        int lab;
        if (tp.single >= 0) lab = tp.single;
        else lab = pisto_decide<C>(acc[r], tp.bits, p.dec, false, nullptr);
        if (do_conf) {
          const unsigned int gg = p.gt[pix];
          if (gg < (unsigned)C) {
            const unsigned int bn = gg * C + lab;
            if (C <= 4) { const unsigned long long inc = 1ull << (8 * (bn & 7)); if (bn < 8) lo += inc; else hi += inc; }
            else atomicAdd(&hist[bn], 1u);
          }
        }
        if (p.label_out) {
          unsigned int o = (unsigned)lab;
          if (p.bg && p.bg[pix] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
          p.label_out[pix] = (uint8_t)o;
        }
      }
    }
    if (do_conf && C <= 4) {
#pragma unroll
      for (int b = 0; b < (C <= 4 ? BINS : 1); b++) {
        unsigned int v = (unsigned int)(((b < 8 ? lo : hi) >> (8 * (b & 7))) & 0xffull);
        v = __reduce_add_sync(0xffffffffu, v);
        if (tx == 0 && v) atomicAdd(&hist[b], v);
      }
    }
    if (n_transposed) __syncthreads();  // the staging tiles are rewritten by the next item
  }
  if (do_conf) {
    __syncthreads();
    for (int i = tid; i < BINS; i += kThreads)
      if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Pipelined variant (the one that runs for the BASELINE shapes): one persistent CTA per SM; the V*C planes of a 32x32 output
// block are brought into shared memory with 16-byte cp.async chunks -- source columns are stored as they lie in memory and
// read back mirrored (flips) or transposed (rot90 / rot270: a 4-way bank conflict that does not matter at this arithmetic
// intensity) -- two blocks deep, so ~100 KB of loads are in flight per SM while the previous block is summed.  The byte masks
// ride along in the same cp.async group.  Measured alternatives: 4-byte cp.async into 33-float-pitch tiles for the transposing
// views (40 % of the HBM peak), one 128-byte 1-D TMA copy per source-row segment (18 %: too many small bulk copies).
// Same arithmetic as above: sum in view order, pisto_decide.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef PISTO_FR_THREADS
#define PISTO_FR_THREADS 512
#endif
constexpr int kPThreads = PISTO_FR_THREADS;  // pipelined kernel: 16 warps, two output rows per thread

struct FullresPlan {
  int nby, nbx;
  int view_off[PISTO_MAX_VIEWS];       // float offset of view v's first plane inside a stage
  int bg_off, gt_off;                  // float offsets of the 32x32 byte blocks of bg / gt inside a stage
  int stage_floats;
  unsigned int inv_fh, inv_fw;         // ceil(2^32 / low_fh), ceil(2^32 / low_fw): y / low_fh == umulhi(y, inv_fh) for y < 2^16
  unsigned int inv_nbx;                // ceil(2^32 / nbx)
};

// x / d for x < 2^16 with inv = ceil(2^32 / d); d == 1 has no 32-bit inverse
__device__ __forceinline__ int fastdiv(int x, int d, unsigned int inv) { return d == 1 ? x : (int)__umulhi((unsigned)x, inv); }

// (tile, block row, block column) of a CTA's current item, advanced by gridDim.x items without divisions
struct ItemPos {
  int n, rem;
  __device__ __forceinline__ void advance(int step_n, int step_rem, int per_tile) {
    n += step_n; rem += step_rem;
    if (rem >= per_tile) { rem -= per_tile; n++; }
  }
};

// BH: output rows per block (32, or 16 when V * C planes of a 32 x 32 block do not fit twice: C = 4 with the 8 d4 views)
// NS: stages of the cp.async ring (NS - 1 blocks in flight while one is summed)
template <int C, int V, int BH, int NS>
__global__ void __launch_bounds__(kPThreads, 1) fuse_fullres_pipe_kernel(const __grid_constant__ FuseParams p, const __grid_constant__ FullresPlan pl) {
  extern __shared__ __align__(16) float stage_mem[];  // [NS][stage_floats]
  __shared__ unsigned int hist[C * C];
  constexpr int BINS = C * C;
  constexpr int TEAM = 8 * BH;              // 16-byte chunks per plane of a block = threads of a copy team
  constexpr int VPAR = kPThreads / TEAM;    // team t issues the copies of views t, t + VPAR, ...
  constexpr int PT = BH + 4;                // row pitch of a transposing view's tile (32 rows of BH source columns)
  constexpr int VH = (V + VPAR - 1) / VPAR;
  constexpr int RPT = BH * kB / kPThreads;  // output rows per thread
  constexpr int RSTEP = kPThreads / 32;     // rows between a thread's output rows
  const bool do_conf = p.conf != nullptr && p.gt != nullptr;
  const bool has_bg = p.bg != nullptr && p.label_out != nullptr;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // rows ty, ty + 16
  for (int i = tid; i < BINS; i += kPThreads) hist[i] = 0;
  const int T_h = p.T_h, T_w = p.T_w;
  const long long hw = (long long)T_h * T_w;
  const int per_tile = pl.nby * pl.nbx;
  const bool need_low = p.lowres_out != nullptr && p.low_fh > 0;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(stage_mem);
  const int step_n = (int)gridDim.x / per_tile, step_rem = (int)gridDim.x - step_n * per_tile;

  // Copies: 16-byte cp.async chunks, thread = (tile row r, chunk q) of every plane of its views.  Tile row r of a row-preserving
  // view is the source row of output row y0 + r; tile row r of a transposing view is the source row of output COLUMN x0 + r.
  // Either way the tile holds 32 consecutive source columns as they lie in memory.
  // element offset of the thread's chunk inside a tile = koff + by * kdy + bx * kdx
  const int cid = tid % TEAM, vpar = tid / TEAM;
  const int cr = cid >> 3, cq = cid & 7;                  // row-preserving view: (tile row, chunk) of the thread's copy
  const int tr = cid / (BH / 4), tq = cid % (BH / 4);     // transposing view: 32 tile rows of BH / 4 chunks
  int koff[VH], kdy[VH], kdx[VH], planes[VH];
#pragma unroll
  for (int j = 0; j < VH; j++) {
    const int v = VPAR * j + vpar < V ? VPAR * j + vpar : 0;
    const ViewDev& vw = p.view[v];
    const ViewMap& m = vw.map;
    planes[j] = vw.h * vw.w;
    if (m.ai != 0) {
      koff[j] = (m.a0 + cr * m.ai) * vw.w + (m.bj > 0 ? m.b0 : m.b0 - (kB - 1)) + 4 * cq;
      kdy[j] = BH * m.ai * vw.w;
      kdx[j] = m.bj > 0 ? kB : -kB;
    } else {
      koff[j] = (m.a0 + tr * m.aj) * vw.w + (m.bi > 0 ? m.b0 : m.b0 - (BH - 1)) + 4 * tq;
      kdy[j] = m.bi > 0 ? BH : -BH;
      kdx[j] = kB * m.aj * vw.w;
    }
  }
  auto presence_of = [&](const ItemPos& ip) {
    TilePresence t; t.bits = 0xffffffffu; t.single = -1;
    if (ip.n < p.N) t = pisto_tile_presence(p, ip.n);
    return t;
  };
  auto issue = [&](const ItemPos& ip, int s, const TilePresence& tp) {
    const int n = ip.n;
    const int by = fastdiv(ip.rem, pl.nbx, pl.inv_nbx), bx = ip.rem - by * pl.nbx;
    const int y0 = by * BH, x0 = bx * kB;
    const uint32_t sbase = smem0 + 4u * (uint32_t)(s * pl.stage_floats);
    const bool full = y0 + BH <= T_h && x0 + kB <= T_w;
    if (vpar == 0 && y0 + cr < T_h && x0 + 4 * cq < T_w) {  // byte masks: 32 rows x 32 bytes, one 4-byte chunk per thread
      const long long o = (long long)n * hw + (long long)(y0 + cr) * T_w + x0 + 4 * cq;
      if (has_bg) cp_async4(sbase + 4u * (uint32_t)pl.bg_off + (uint32_t)(cr * kB + 4 * cq), p.bg + o);
      if (do_conf) cp_async4(sbase + 4u * (uint32_t)pl.gt_off + (uint32_t)(cr * kB + 4 * cq), p.gt + o);
    }
    if (!(tp.single < 0 || p.fused_out || need_low)) return;  // single-label tile whose scores nobody wants
#pragma unroll
    for (int j = 0; j < VH; j++) {
      const int v = VPAR * j + vpar;
      if (v >= V) break;
      const ViewDev& vw = p.view[v];
      const ViewMap& m = vw.map;
      bool ok = true;
      if (!full) {
        int col;
        if (m.ai != 0) { col = (m.bj > 0 ? m.b0 + x0 : m.b0 - x0 - (kB - 1)) + 4 * cq; ok = y0 + cr < T_h; }
        else { col = (m.bi > 0 ? m.b0 + y0 : m.b0 - y0 - (BH - 1)) + 4 * tq; ok = x0 + tr < T_w; }
        ok = ok && col >= 0 && col + 3 < vw.w;
      }
      if (ok) {
        const float* g = vw.logits + (long long)n * vw.tile_stride + (koff[j] + by * kdy[j] + bx * kdx[j]);
        const uint32_t d = sbase + 4u * (uint32_t)(pl.view_off[v] + (m.ai != 0 ? cr * kB + 4 * cq : tr * PT + 4 * tq));
        const int pfl = m.ai != 0 ? BH * kB : kB * PT;   // floats per class plane of the tile
#pragma unroll
        for (int c = 0; c < C; c++) cp_async16(d + 4u * (uint32_t)(c * pfl), g + c * planes[j]);
      }
    }
  };

  // ring of NS stages: items q[0] (being summed) .. q[NS-1] (the one issued in this iteration), q[NS] = the one whose presence
  // vector is requested now, one iteration before its copies are issued
  ItemPos q[NS + 1];
  TilePresence tps[NS + 1];
  q[0].n = (int)blockIdx.x / per_tile; q[0].rem = (int)blockIdx.x - q[0].n * per_tile;
#pragma unroll
  for (int j = 1; j <= NS; j++) { q[j] = q[j - 1]; q[j].advance(step_n, step_rem, per_tile); }
#pragma unroll
  for (int j = 0; j < NS; j++) tps[j] = presence_of(q[j]);
#pragma unroll
  for (int j = 0; j < NS - 1; j++) {
    if (q[j].n < p.N) issue(q[j], j, tps[j]);
    cp_async_commit();
  }
  int s = 0;
  for (; q[0].n < p.N; s = (s + 1 == NS ? 0 : s + 1)) {
    ItemPos& cur = q[0];
    TilePresence& tp_cur = tps[0];
    tps[NS] = presence_of(q[NS]);
    if (q[NS - 1].n < p.N) issue(q[NS - 1], s == 0 ? NS - 1 : s - 1, tps[NS - 1]);
    cp_async_commit();
    cp_async_wait<NS - 1>();   // this item's group has landed (the NS - 1 younger ones may still be in flight)
    __syncthreads();
    const int n = cur.n;
    const int by = fastdiv(cur.rem, pl.nbx, pl.inv_nbx), bx = cur.rem - by * pl.nbx;
    const int y0 = by * BH, x0 = bx * kB;
    const TilePresence tp = tp_cur;
    const bool need_scores = tp.single < 0 || p.fused_out || need_low;
    const float* st = stage_mem + s * pl.stage_floats;
    float acc[RPT][C];
    if (need_scores) {
#pragma unroll
      for (int v = 0; v < V; v++) {
        const ViewMap& m = p.view[v].map;
        const float* vb = st + pl.view_off[v];
#pragma unroll
        for (int r = 0; r < RPT; r++) {
          const int row = ty + RSTEP * r;
          int idx, cstride;
          if (m.ai != 0) { idx = row * kB + (m.bj > 0 ? tx : kB - 1 - tx); cstride = BH * kB; }
          else { idx = tx * PT + (m.bi > 0 ? row : BH - 1 - row); cstride = kB * PT; }  // 4-way bank conflict: negligible here
#pragma unroll
          for (int c = 0; c < C; c++) {
            const float val = vb[idx + c * cstride];
            acc[r][c] = (v == 0) ? val : __fadd_rn(acc[r][c], val);
          }
        }
      }
    }
    const uint8_t* bgs = reinterpret_cast<const uint8_t*>(st + pl.bg_off);
    const uint8_t* gts = reinterpret_cast<const uint8_t*>(st + pl.gt_off);
    unsigned long long lo = 0, hi = 0;
    const int xx = x0 + tx;
    bool lowx = false; int lxq = 0;
    if (need_low) { lxq = fastdiv(xx, p.low_fw, pl.inv_fw); lowx = xx - lxq * p.low_fw == p.low_fw / 2; }
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int yy = y0 + ty + RSTEP * r;
      if (yy < T_h && xx < T_w) {
        const long long rpix = (long long)yy * T_w + xx, pix = (long long)n * hw + rpix;
        if (p.fused_out) {
#pragma unroll
          for (int c = 0; c < C; c++) p.fused_out[((long long)n * C + c) * hw + rpix] = pisto_div_views(acc[r][c], p.dec);
        }
        if (lowx) {
          const int lyq = fastdiv(yy, p.low_fh, pl.inv_fh);
          if (yy - lyq * p.low_fh == p.low_fh / 2) {
#pragma unroll
            for (int c = 0; c < C; c++)
              p.lowres_out[(((long long)n * C + c) * p.low_h + lyq) * p.low_w + lxq] = pisto_div_views(acc[r][c], p.dec);
          }
        }
        int lab;
        if (tp.single >= 0) lab = tp.single;
        else lab = pisto_decide<C>(acc[r], tp.bits, p.dec, false, nullptr);
        if (do_conf) {
          const unsigned int gg = gts[(ty + RSTEP * r) * kB + tx];
          if (gg < (unsigned)C) {
            const unsigned int bn = gg * C + lab;
            const unsigned long long inc = 1ull << (8 * (bn & 7));
            if (bn < 8) lo += inc; else hi += inc;
          }
        }
        if (p.label_out) {
          unsigned int o = (unsigned)lab;
          if (has_bg && bgs[(ty + RSTEP * r) * kB + tx] == (uint8_t)p.bg_match) o = (unsigned)p.bg_label;
          p.label_out[pix] = (uint8_t)o;
        }
      }
    }
    if (do_conf) {
#pragma unroll
      for (int b = 0; b < BINS; b++) {
        unsigned int v = (unsigned int)(((b < 8 ? lo : hi) >> (8 * (b & 7))) & 0xffull);
        v = __reduce_add_sync(0xffffffffu, v);
        if (tx == 0 && v) atomicAdd(&hist[b], v);
      }
    }
#pragma unroll
    for (int j = 0; j < NS; j++) { q[j] = q[j + 1]; tps[j] = tps[j + 1]; }
    q[NS].advance(step_n, step_rem, per_tile);
    __syncthreads();  // stage s is refilled two iterations from now, by copies issued after this barrier
  }
  cp_async_wait<0>();
  if (do_conf) {
    __syncthreads();
    for (int i = tid; i < BINS; i += kPThreads)
      if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  }
}

template <int C, int V, int BH, int NS>
static int launch_fullres_pipe_bh(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  static_assert(C <= 4, "packed 8-bit confusion counters");
  if (p.T_w % 4 || p.T_h > 65535 || p.T_w > 65535) return PISTO_OK;
  if ((((uintptr_t)p.bg | (uintptr_t)p.gt) & 3) || (((long long)p.T_h * p.T_w) % 4)) return PISTO_OK;  // 4-byte chunks of the byte masks
  FullresPlan pl;
  memset(&pl, 0, sizeof(pl));
  int fl = 0;
  for (int v = 0; v < V; v++) {
    const ViewDev& vw = p.view[v];
    if (((uintptr_t)vw.logits & 15) || (vw.w % 4) || (vw.tile_stride % 4)) return PISTO_OK;  // 16-byte chunks
    if ((long long)vw.h * vw.w * 4 > 0x7fffffffLL / 8) return PISTO_OK;                      // 32-bit element offsets inside a tile
    pl.view_off[v] = fl;
    fl += C * (vw.map.ai != 0 ? BH * kB : kB * (BH + 4));
  }
  pl.bg_off = fl; fl += BH * kB / 4;
  pl.gt_off = fl; fl += BH * kB / 4;
  pl.stage_floats = fl;
  if (p.low_fh > 0 && p.low_fw > 0) {
    pl.inv_fh = (unsigned int)(((1ull << 32) + p.low_fh - 1) / p.low_fh);
    pl.inv_fw = (unsigned int)(((1ull << 32) + p.low_fw - 1) / p.low_fw);
  }
  const size_t smem = NS * (size_t)fl * sizeof(float);
  if (smem > (size_t)h->smem_optin - 2048) return PISTO_OK;
  pl.nby = (p.T_h + BH - 1) / BH;
  pl.nbx = (p.T_w + kB - 1) / kB;
  pl.inv_nbx = (unsigned int)(((1ull << 32) + pl.nbx - 1) / pl.nbx);
  const long long items = (long long)p.N * pl.nby * pl.nbx;
  if (items > 0x7fffffffLL / 2) return PISTO_OK;
  if ((long long)pl.nby * pl.nbx > 65535) return PISTO_OK;  // block index / nbx through the 32-bit inverse is exact below 2^16
  const int grid = (int)(items < h->sm_count ? items : h->sm_count);
  PISTO_CUDA(cudaFuncSetAttribute(fuse_fullres_pipe_kernel<C, V, BH, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fuse_fullres_pipe_kernel<C, V, BH, NS><<<grid, kPThreads, smem, st>>>(p, pl);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}

// 32-row blocks when two stages of them fit, else 16-row blocks
template <int C, int V>
static int launch_fullres_pipe(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  // measured on Mode F (C = 3): 32-row blocks x 2 stages 0.80 M tiles/s; 16-row blocks x 3 / 4 stages 0.60 / 0.58 M (64-byte
  // segments of the transposing views, one pixel per thread); C = 4: 16 rows x 2 stages 0.52 M, x 3 stages 0.51 M
  static const bool deep = getenv("PISTO_FR_DEEP") != nullptr;  // A/B knob: 16-row blocks, 3 stages
  if (deep) return launch_fullres_pipe_bh<C, V, 16, 3>(h, p, st, launched);
  int rc = launch_fullres_pipe_bh<C, V, 32, 2>(h, p, st, launched);
  if (rc != PISTO_OK || *launched) return rc;
  return launch_fullres_pipe_bh<C, V, 16, 2>(h, p, st, launched);
}

template <int C>
static int launch_fullres_pipe_c(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  switch (p.V) {
    case 2: return launch_fullres_pipe<C, 2>(h, p, st, launched);   // hflip_transform
    case 4: return launch_fullres_pipe<C, 4>(h, p, st, launched);   // flips / rot180
    case 8: return launch_fullres_pipe<C, 8>(h, p, st, launched);   // d4_transform (infer_pseudo_masks.py:96)
    default: return PISTO_OK;
  }
}

}  // namespace

int pisto_launch_fuse_fullres(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  *launched = false;
  if (p.fuse_mode != PISTO_FUSE_LOGIT_MEAN || p.entropy_out) return PISTO_OK;
  if (p.V < 2) return PISTO_OK;  // one view: fuse_identity_kernel
  int ntr = 0;
  for (int v = 0; v < p.V; v++) {
    const ViewDev& vw = p.view[v];
    if (!(vw.same_h && vw.same_w)) return PISTO_OK;
    if (vw.map.ai == 0) ntr++;
  }
  if (!getenv("PISTO_FULLRES_OLD")) {  // pipelined kernel first (A/B knob: the single-stage kernel below)
    int rc = PISTO_OK;
    if (p.C == 2) rc = launch_fullres_pipe_c<2>(h, p, st, launched);
    else if (p.C == 3) rc = launch_fullres_pipe_c<3>(h, p, st, launched);
    else if (p.C == 4) rc = launch_fullres_pipe_c<4>(h, p, st, launched);
    if (rc != PISTO_OK || *launched) return rc;
  }
  const size_t smem = (size_t)ntr * p.C * kB * kPad * sizeof(float);
  if (smem > (size_t)h->smem_optin - 2048) return PISTO_OK;
  const int nby = (p.T_h + kB - 1) / kB, nbx = (p.T_w + kB - 1) / kB;
  const long long items = (long long)p.N * nby * nbx;
  const int per_sm = smem > 0 ? (int)((size_t)(h->smem_optin) / (smem + 1024)) : 8;
  long long grid = (long long)h->sm_count * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > items) grid = items;
#define PISTO_FR_CASE(CC)                                                                                                         \
  case CC:                                                                                                                        \
    if (smem > 48 * 1024) PISTO_CUDA(cudaFuncSetAttribute(fuse_fullres_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    fuse_fullres_kernel<CC><<<(int)grid, kThreads, smem, st>>>(p, nby, nbx, ntr);                                               \
    break;
  switch (p.C) {
    PISTO_FR_CASE(1) PISTO_FR_CASE(2) PISTO_FR_CASE(3) PISTO_FR_CASE(4) PISTO_FR_CASE(5) PISTO_FR_CASE(6) PISTO_FR_CASE(7) PISTO_FR_CASE(8)
    default: return PISTO_OK;
  }
#undef PISTO_FR_CASE
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}
