// Identity-view fast path of pisto_fuse_argmax_confusion: ONE view that already has the output resolution and needs no
// de-augmentation -- the shape of mIoUMask.forward (reference loss.py:55-67: softmax over the class axis of full-resolution
// logits, argmax, confusion against the mask) and of the revise-mask post-processing (infer_revise_masks.py:137-143).
// Nothing is interpolated, so the op is a pure stream: C float4 loads + one mask word per 4 pixels in, (optionally) 4 label
// bytes out; the decision is pisto_decide (margin fast path, exact slow path) exactly as in every other fusion kernel.
// HBM-bound: 4*C + 1 (+1 bg, +1 label) bytes per pixel.
#include "fuse_common.cuh"

namespace {

constexpr int kThreads = 256;

// Unmasked fast decision of one pixel (no present vector / MASK_NONE): the same test as pisto_decide -- best class leads the
// runner-up by more than margin_abs + 2.4e-7 |best|, every score finite and below 1e30 -- written with min / max only (top-2
// of C values in 3 (C - 1) FMNMX) and without the per-class presence selects.  Returns the label, or -1 when the pixel has to
// take the exact path (near ties, exact ties, NaN / Inf, absurd magnitudes).
template <int C>
__device__ __forceinline__ int identity_decide_fast(const float (&a)[C], const DecideCfg& cfg) {
  float bv = a[0], sv = -INFINITY, probe = a[0];
#pragma unroll
  for (int c = 1; c < C; c++) {
    const float lo = fminf(bv, a[c]);
    bv = fmaxf(bv, a[c]);
    sv = fmaxf(sv, lo);
    probe = __fadd_rn(probe, a[c]);
  }
  int bi = 0;
#pragma unroll
  for (int c = C - 1; c >= 1; c--) bi = (a[c] == bv) ? c : bi;
  bi = (a[0] == bv) ? 0 : bi;
  const float margin = __fmaf_rn(fabsf(bv), 2.4e-7f, cfg.margin_abs);
  const bool ok = (__fsub_rn(bv, sv) > margin) && (fabsf(probe) < 1e30f) && (bv > -1e9f);   // NaN fails |probe| < 1e30
  return ok ? bi : -1;
}

// RAW2: also write label_raw_out (a template parameter: the common kernel must not carry the second decision's registers)
template <int C, bool RAW2 = false>
__global__ void __launch_bounds__(kThreads) fuse_identity_kernel(const __grid_constant__ FuseParams p) {
  constexpr int BINS = C * C;
  __shared__ unsigned int hist[BINS];
  const bool do_conf = p.conf != nullptr && p.gt != nullptr;
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const ViewDev& vw = p.view[0];
  const long long hw = (long long)p.T_h * p.T_w;
  const long long q_per_tile = hw / 4;
  const long long total = q_per_tile * p.N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // packed 8-bit counters (C <= 4): one 32-bit register per ground-truth class, byte l = predictions of class l; flushed before
  // a byte can overflow
  unsigned int rowcnt[C <= 4 ? C : 1];
#pragma unroll
  for (int i = 0; i < (C <= 4 ? C : 1); i++) rowcnt[i] = 0;
  int pending = 0;
  auto flush = [&]() {
    if (C <= 4) {
#pragma unroll
      for (int b = 0; b < (C <= 4 ? BINS : 1); b++) {
        unsigned int v = (rowcnt[C <= 4 ? b / C : 0] >> (8 * (b % C))) & 0xffu;
        v = __reduce_add_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&hist[b], v);
      }
#pragma unroll
      for (int i = 0; i < (C <= 4 ? C : 1); i++) rowcnt[i] = 0;
    }
    pending = 0;
  };
  const bool unmasked = !(p.present && p.dec.mask_mode != PISTO_MASK_NONE) && p.dec.mask_mode != PISTO_MASK_MULTIPLY;
  const long long warp_base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  // (tile, group inside the tile) of the lane's current 4-pixel group, advanced without 64-bit divisions
  const long long q0 = warp_base + (threadIdx.x & 31);
  int n = (int)(q0 / q_per_tile);
  long long rq = q0 - (long long)n * q_per_tile;
  const int step_n = (int)(stride / q_per_tile);
  const long long step_r = stride - (long long)step_n * q_per_tile;
  // Two groups (this grid-stride step's and the next one's) per iteration: all their loads are issued before the first is
  // decided.  Since the unmasked decision became cheap the kernel is bound by the bytes a thread keeps in flight (60 registers
  // -> 1024 threads per SM x 52 bytes was under the ~66 KB an SM needs outstanding to cover HBM latency at full rate).
  struct Group { float4 v[C]; unsigned int g4, bg4; int n; long long r4; bool live, scores; TilePresence tp; };
  auto fetch = [&](long long q_, int n_, long long rq_) {
    Group gr;
    gr.live = q_ < total; gr.n = n_; gr.r4 = rq_ * 4; gr.g4 = 0xffffffffu; gr.bg4 = 0; gr.scores = false;
    gr.tp.bits = 0xffffffffu; gr.tp.single = -1;
    if (gr.live) {
      gr.tp = pisto_tile_presence(p, n_);
      gr.scores = gr.tp.single < 0;
      const long long pix = (long long)n_ * hw + gr.r4;
      if (gr.scores) {
        const float* base = vw.logits + (long long)n_ * vw.tile_stride + gr.r4;
#pragma unroll
        for (int c = 0; c < C; c++) gr.v[c] = __ldcs(reinterpret_cast<const float4*>(base + c * hw));
      }
      if (do_conf) gr.g4 = __ldcs(reinterpret_cast<const unsigned int*>(p.gt + pix));
      if ((p.label_out || RAW2) && p.bg) gr.bg4 = __ldcs(reinterpret_cast<const unsigned int*>(p.bg + pix));
    }
    return gr;
  };
  auto process = [&](const Group& gr) {
    if (!gr.live) return;
    const TilePresence tp = gr.tp;
    int lab[4], labr[4];
    if (!gr.scores) {
      lab[0] = lab[1] = lab[2] = lab[3] = tp.single;
      if (RAW2) labr[0] = labr[1] = labr[2] = labr[3] = tp.single;
    } else {
      float a[4][C];
#pragma unroll
      for (int c = 0; c < C; c++) { a[0][c] = gr.v[c].x; a[1][c] = gr.v[c].y; a[2][c] = gr.v[c].z; a[3][c] = gr.v[c].w; }
      bool all_ok = false;
      if (unmasked) {
#pragma unroll
        for (int j = 0; j < 4; j++) lab[j] = identity_decide_fast<C>(a[j], p.dec);
        all_ok = (lab[0] | lab[1] | lab[2] | lab[3]) >= 0;
      }
      // second output (PISTO_DECIDE_RAW): where the cheap decision held, the leader's margin makes argmax(softmax) == argmax(logits)
      if (RAW2) {
#pragma unroll
        for (int j = 0; j < 4; j++) labr[j] = lab[j];
      }
      if (!all_ok) {
        DecideCfg raw = p.dec;
        raw.decide_mode = PISTO_DECIDE_RAW;
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (!unmasked || lab[j] < 0) {
            lab[j] = pisto_decide<C>(a[j], tp.bits, p.dec, false, nullptr);
            if (RAW2) labr[j] = pisto_decide<C>(a[j], tp.bits, raw, false, nullptr);
          }
      }
    }
    const long long pix = (long long)gr.n * hw + gr.r4;
    const unsigned int l4 = (unsigned)lab[0] | ((unsigned)lab[1] << 8) | ((unsigned)lab[2] << 16) | ((unsigned)lab[3] << 24);
    if (p.label_out) {
      unsigned int o = l4;
      if (p.bg) {
        const unsigned int eq = __vcmpeq4(gr.bg4, 0x01010101u * (unsigned)p.bg_match);
        o = ((0x01010101u * (unsigned)p.bg_label) & eq) | (o & ~eq);
      }
      *reinterpret_cast<unsigned int*>(p.label_out + pix) = o;
    }
    if (RAW2) {
      unsigned int o = (unsigned)labr[0] | ((unsigned)labr[1] << 8) | ((unsigned)labr[2] << 16) | ((unsigned)labr[3] << 24);
      if (p.bg) {
        const unsigned int eq = __vcmpeq4(gr.bg4, 0x01010101u * (unsigned)p.bg_match);
        o = ((0x01010101u * (unsigned)p.bg_label) & eq) | (o & ~eq);
      }
      *reinterpret_cast<unsigned int*>(p.label_raw_out + pix) = o;
    }
    if (do_conf) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned int gg = (gr.g4 >> (8 * j)) & 0xffu, lb = (l4 >> (8 * j)) & 0xffu;
        if (C <= 4) {
          const unsigned int inc = 1u << (8 * lb);
#pragma unroll
          for (int c = 0; c < (C <= 4 ? C : 1); c++) rowcnt[c] += (gg == (unsigned)c) ? inc : 0u;
        } else if (gg < (unsigned)C) {
          atomicAdd(&hist[gg * C + lb], 1u);
        }
      }
    }
  };
  // all lanes of a warp run the same number of iterations (flush uses full-mask warp reductions)
  for (long long qb = warp_base; qb < total; qb += 2 * stride) {
    const long long q = qb + (threadIdx.x & 31);
    const Group ga = fetch(q, n, rq);
    n += step_n; rq += step_r;
    if (rq >= q_per_tile) { rq -= q_per_tile; n++; }
    const Group gb = fetch(q + stride, n, rq);
    n += step_n; rq += step_r;
    if (rq >= q_per_tile) { rq -= q_per_tile; n++; }
    process(ga);
    process(gb);
    pending += 8;
    if (do_conf && pending > 255 - 8) flush();
  }
  if (do_conf) {
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += blockDim.x)
      if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  }
}

}  // namespace

int pisto_launch_fuse_identity(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched) {
  *launched = false;
  if (p.V != 1 || p.fuse_mode != PISTO_FUSE_LOGIT_MEAN) return PISTO_OK;
  if (p.fused_out || p.entropy_out || p.lowres_out) return PISTO_OK;
  const ViewDev& vw = p.view[0];
  const ViewMap& m = vw.map;
  if (!(vw.same_h && vw.same_w && m.a0 == 0 && m.ai == 1 && m.aj == 0 && m.b0 == 0 && m.bi == 0 && m.bj == 1)) return PISTO_OK;
  const long long hw = (long long)p.T_h * p.T_w;
  if (hw % 4 || vw.tile_stride % 4) return PISTO_OK;
  if (((uintptr_t)vw.logits & 15) || (((uintptr_t)p.label_out | (uintptr_t)p.label_raw_out | (uintptr_t)p.bg | (uintptr_t)p.gt) & 3)) return PISTO_OK;
  const long long total = hw / 4 * p.N;
  long long grid = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)h->sm_count * 16;
  if (grid > cap) grid = cap;
  switch (p.C) {
#define PISTO_ID_CASE(CC) case CC: if (p.label_raw_out) fuse_identity_kernel<CC, true><<<(int)grid, kThreads, 0, st>>>(p); else fuse_identity_kernel<CC, false><<<(int)grid, kThreads, 0, st>>>(p); break;
    PISTO_ID_CASE(1) PISTO_ID_CASE(2) PISTO_ID_CASE(3) PISTO_ID_CASE(4) PISTO_ID_CASE(5) PISTO_ID_CASE(6) PISTO_ID_CASE(7) PISTO_ID_CASE(8)
#undef PISTO_ID_CASE
    default: return PISTO_OK;
  }
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  *launched = true;
  return PISTO_OK;
}
