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
#include "fuse_common.cuh"

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
