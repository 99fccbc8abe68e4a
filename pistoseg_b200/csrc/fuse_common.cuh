// Parameter block shared by the fusion kernels (fuse_generic.cu, fuse_stream.cu).
#pragma once
#include "common.cuh"

struct ViewDev {
  const float* logits;
  long long tile_stride;  // elements
  int h, w;               // raw (augmented-frame) size
  ViewMap map;            // de-augmentation index map, map.ho x map.wo is the de-augmented size
  float scale_h, scale_w; // (float)ho / T_h, (float)wo / T_w
  int same_h, same_w;     // ho == T_h, wo == T_w -> identity lerp
};

struct FuseParams {
  ViewDev view[PISTO_MAX_VIEWS];
  int V, N, C, T_h, T_w;
  int fuse_mode;
  DecideCfg dec;
  int bg_match, bg_label;
  int low_h, low_w, low_fh, low_fw;  // in-kernel gather: y % low_fh == low_fh/2 (0 = no in-kernel lowres)
  const uint8_t* present;
  const uint8_t* bg;
  const uint8_t* gt;
  uint8_t* label_out;
  uint8_t* label_raw_out;  // second label output, decided with PISTO_DECIDE_RAW (identity / generic kernels only)
  float* fused_out;
  float* entropy_out;
  float* lowres_out;
  unsigned long long* conf;
};

// tile-level class-presence summary
struct TilePresence {
  uint32_t bits;  // bit c = class c present (all ones when there is no present vector)
  int single;     // >= 0: MASK_FILL shortcut label (exactly one class present), else -1
};

__device__ __forceinline__ TilePresence pisto_tile_presence(const FuseParams& p, int n) {
  TilePresence t;
  t.bits = 0xffffffffu;
  t.single = -1;
  if (p.present && p.dec.mask_mode != PISTO_MASK_NONE) {
    uint32_t b = 0;
    int cnt = 0;
    for (int c = 0; c < p.C; c++) {
      int v = p.present[(long long)n * p.C + c];
      if (v) b |= 1u << c;
      cnt += v;  // the reference tests sum(patch_label) == 1 (infer_pseudo_masks.py:71)
    }
    t.bits = b;
    if (p.dec.mask_mode == PISTO_MASK_FILL && cnt == 1) t.single = __ffs(b) - 1;  // patch_label.index(1)
  }
  return t;
}

// One bilinear sample of de-augmented view v, class c, tile n at output (y, x): formula A of SURVEY.md A.1.
__device__ __forceinline__ float pisto_sample_view(const ViewDev& vw, int C_unused, int n, int c, const Lerp& ly, const Lerp& lx) {
  const float* base = vw.logits + (long long)n * vw.tile_stride + (long long)c * vw.h * vw.w;
  const ViewMap& m = vw.map;
  auto at = [&](int i, int j) -> float {
    int a = m.a0 + i * m.ai + j * m.aj;
    int b = m.b0 + i * m.bi + j * m.bj;
    return __ldg(base + a * vw.w + b);
  };
  float v00 = at(ly.i0, lx.i0), v01 = at(ly.i0, lx.i1), v10 = at(ly.i1, lx.i0), v11 = at(ly.i1, lx.i1);
  float r0 = __fmaf_rn(lx.l0, v00, __fmul_rn(lx.l1, v01));
  float r1 = __fmaf_rn(lx.l0, v10, __fmul_rn(lx.l1, v11));
  return __fmaf_rn(ly.l0, r0, __fmul_rn(ly.l1, r1));
}

int pisto_launch_fuse_generic(pisto_ctx* h, const FuseParams& p, cudaStream_t st);
// returns PISTO_ERR_UNSUPPORTED (without setting an error) when the shape has no streaming instantiation
int pisto_launch_fuse_stream(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_fuse_filter(pisto_ctx* h, const FuseParams& p, cudaStream_t st, int np, bool* launched);
int pisto_launch_fuse_identity(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_fuse_fullres(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_fuse_band(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
int pisto_launch_fuse_block(pisto_ctx* h, const FuseParams& p, cudaStream_t st, bool* launched);
