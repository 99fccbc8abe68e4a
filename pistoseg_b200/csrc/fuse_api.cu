// pisto_fuse_argmax_confusion: argument validation, parameter block, kernel selection.
#include "fuse_common.cuh"

int pisto_upsample_launch(pisto_ctx* h, const void* in, void* out, long long NC, int hi, int wi, int ho, int wo, int dtype,
                          cudaStream_t st, const double* count = nullptr, double min_count = 0.0, int accumulate = 0);

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int pisto_build_fuse_params(const pisto_view_t* views, int V, const pisto_fuse_args_t* a, FuseParams* out, bool* low_via_resize) {
  PISTO_REQUIRE(views && a, "pisto_fuse_argmax_confusion: views/args NULL");
  PISTO_REQUIRE(V >= 1 && V <= PISTO_MAX_VIEWS, "pisto_fuse_argmax_confusion: V=%d outside [1,%d]", V, PISTO_MAX_VIEWS);
  PISTO_REQUIRE(a->C >= 1 && a->C <= PISTO_MAX_CLASSES, "pisto_fuse_argmax_confusion: C=%d outside [1,%d]", a->C, PISTO_MAX_CLASSES);
  PISTO_REQUIRE(a->N >= 0 && a->T_h >= 1 && a->T_w >= 1, "pisto_fuse_argmax_confusion: bad N/T_h/T_w (%d,%d,%d)", a->N, a->T_h, a->T_w);
  PISTO_REQUIRE(a->fuse_mode == PISTO_FUSE_LOGIT_MEAN || a->fuse_mode == PISTO_FUSE_PROB_MEAN, "pisto_fuse_argmax_confusion: bad fuse_mode %d", a->fuse_mode);
  PISTO_REQUIRE(a->mask_mode >= PISTO_MASK_NONE && a->mask_mode <= PISTO_MASK_MULTIPLY, "pisto_fuse_argmax_confusion: bad mask_mode %d", a->mask_mode);
  PISTO_REQUIRE(a->decide_mode == PISTO_DECIDE_SOFTMAX || a->decide_mode == PISTO_DECIDE_RAW, "pisto_fuse_argmax_confusion: bad decide_mode %d", a->decide_mode);
  PISTO_REQUIRE(a->mask_mode == PISTO_MASK_NONE || a->present, "pisto_fuse_argmax_confusion: mask_mode %d needs present[N][C]", a->mask_mode);
  PISTO_REQUIRE(!(a->conf && !a->gt), "pisto_fuse_argmax_confusion: conf given without gt");
  PISTO_REQUIRE(a->bg_label >= 0 && a->bg_label <= 255 && a->bg_match >= 0 && a->bg_match <= 255, "pisto_fuse_argmax_confusion: bg_label/bg_match outside u8");
  FuseParams& p = *out;
  memset(&p, 0, sizeof(p));
  p.V = V; p.N = a->N; p.C = a->C; p.T_h = a->T_h; p.T_w = a->T_w;
  p.fuse_mode = a->fuse_mode;
  p.dec.mask_mode = a->present ? a->mask_mode : PISTO_MASK_NONE;
  p.dec.decide_mode = a->decide_mode;
  p.dec.V = V;
  p.dec.inv_v = is_pow2(V) ? 1.0f / (float)V : 0.f;
  p.dec.margin_abs = 2e-6f * (float)V;
  p.dec.fV = (float)V;
  p.dec.rcp_v = 1.0f / (float)V;
  p.bg_match = a->bg_match; p.bg_label = a->bg_label;
  p.present = a->present; p.bg = a->bg; p.gt = a->gt;
  p.label_out = a->label_out; p.label_raw_out = a->label_raw_out; p.fused_out = a->fused_out; p.entropy_out = a->entropy_out; p.lowres_out = a->lowres_out;
  p.conf = a->conf;
  for (int v = 0; v < V; v++) {
    const pisto_view_t& s = views[v];
    PISTO_REQUIRE(s.logits || a->N == 0, "pisto_fuse_argmax_confusion: view %d logits NULL", v);
    PISTO_REQUIRE(s.h >= 1 && s.w >= 1 && s.h <= 32767 && s.w <= 32767, "pisto_fuse_argmax_confusion: view %d bad size %dx%d", v, s.h, s.w);
    PISTO_REQUIRE(s.xform >= 0 && s.xform < 8, "pisto_fuse_argmax_confusion: view %d xform %d outside [0,8)", v, s.xform);
    PISTO_REQUIRE(((uintptr_t)s.logits & 3) == 0, "pisto_fuse_argmax_confusion: view %d logits not 4-byte aligned", v);
    ViewDev& d = p.view[v];
    d.logits = s.logits;
    d.h = s.h; d.w = s.w;
    d.tile_stride = s.tile_stride ? s.tile_stride : (long long)a->C * s.h * s.w;
    d.map = pisto_view_map(s.xform, s.h, s.w);
    d.same_h = d.map.ho == a->T_h; d.same_w = d.map.wo == a->T_w;
    d.scale_h = (float)d.map.ho / (float)a->T_h;
    d.scale_w = (float)d.map.wo / (float)a->T_w;
  }
  *low_via_resize = false;
  if (a->lowres_out) {
    PISTO_REQUIRE(a->low_h >= 1 && a->low_w >= 1, "pisto_fuse_argmax_confusion: lowres_out needs low_h/low_w");
    p.low_h = a->low_h; p.low_w = a->low_w;
    bool gather = (a->T_h % a->low_h == 0) && (a->T_w % a->low_w == 0) && ((a->T_h / a->low_h) & 1) && ((a->T_w / a->low_w) & 1);
    if (gather) {
      p.low_fh = a->T_h / a->low_h; p.low_fw = a->T_w / a->low_w;
    } else {
      PISTO_REQUIRE(a->fused_out, "pisto_fuse_argmax_confusion: lowres_out %dx%d from %dx%d is not an odd-factor gather; pass fused_out so the library can resize it",
                    a->low_h, a->low_w, a->T_h, a->T_w);
      *low_via_resize = true;
      p.lowres_out = nullptr;
    }
  }
  return PISTO_OK;
}

extern "C" int pisto_fuse_argmax_confusion(pisto_handle_t h, const pisto_view_t* views, int V, const pisto_fuse_args_t* a,
                                           pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_fuse_argmax_confusion: NULL handle");
  FuseParams p;
  bool low_via_resize = false;
  int rc = pisto_build_fuse_params(views, V, a, &p, &low_via_resize);
  if (rc != PISTO_OK) return rc;
  if (p.N == 0) return PISTO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PISTO_CUDA(cudaSetDevice(h->device));
  bool launched = false;
  const bool raw2 = p.label_raw_out != nullptr;  // second label output: the one-view kernel or the generic kernel
  if (raw2 && a->impl > 1) {
    pisto_set_error("pisto_fuse_argmax_confusion: label_raw_out is served by impl 0 / 1 only (asked for impl=%d)", a->impl);
    return PISTO_ERR_UNSUPPORTED;
  }
  if (a->impl == 0) {  // one full-resolution view: pure streaming kernel
    rc = pisto_launch_fuse_identity(h, p, st, &launched);
    if (rc != PISTO_OK) return rc;
  }
  if (!launched && !raw2 && a->impl == 0) {  // several full-resolution views (ttach d4 on full-resolution logits)
    rc = pisto_launch_fuse_fullres(h, p, st, &launched);
    if (rc != PISTO_OK) return rc;
  }
  if (!launched && !raw2 && (a->impl == 0 || a->impl == 3 || a->impl == 4 || a->impl == 6 || a->impl == 7 || a->impl == 8)) {
    rc = pisto_launch_fuse_filter(h, p, st, a->impl == 0 ? 0 : (a->impl >= 6 ? a->impl - 3 : a->impl - 2), &launched);
    if (rc != PISTO_OK) return rc;
    if (!launched && a->impl != 0) {
      pisto_set_error("pisto_fuse_argmax_confusion: impl=%d (filtered streaming kernel) has no instantiation for C=%d V=%d T=%dx%d with these options",
                      a->impl, p.C, p.V, p.T_h, p.T_w);
      return PISTO_ERR_UNSUPPORTED;
    }
  }
  if (!launched && !raw2 && (a->impl == 0 || a->impl == 5)) {  // large tiles: the same filter on output blocks
    rc = pisto_launch_fuse_band(h, p, st, &launched);
    if (rc != PISTO_OK) return rc;
    if (!launched && a->impl == 5) {
      pisto_set_error("pisto_fuse_argmax_confusion: impl=5 (block-tiled filtered kernel) has no instantiation for C=%d V=%d T=%dx%d with these options",
                      p.C, p.V, p.T_h, p.T_w);
      return PISTO_ERR_UNSUPPORTED;
    }
  }
  if (!launched && !raw2 && a->impl != 1) {
    rc = pisto_launch_fuse_stream(h, p, st, &launched);
    if (rc != PISTO_OK) return rc;
    if (!launched) {
      rc = pisto_launch_fuse_block(h, p, st, &launched);
      if (rc != PISTO_OK) return rc;
    }
    if (!launched && a->impl == 2) {
      pisto_set_error("pisto_fuse_argmax_confusion: impl=2 (streaming kernel) has no instantiation for C=%d V=%d T=%dx%d with these options",
                      p.C, p.V, p.T_h, p.T_w);
      return PISTO_ERR_UNSUPPORTED;
    }
  }
  if (!launched) {
    rc = pisto_launch_fuse_generic(h, p, st);
    if (rc != PISTO_OK) return rc;
  }
  if (low_via_resize) {
    rc = pisto_upsample_launch(h, a->fused_out, a->lowres_out, (long long)p.N * p.C, p.T_h, p.T_w, a->low_h, a->low_w, 0, st);
    if (rc != PISTO_OK) return rc;
  }
  return PISTO_OK;
}

// ---- self-test hook: exhaustive check of the division-by-view-count shortcut ---------------------------------------
namespace {
__global__ void selftest_div_kernel(int V, unsigned long long* mismatches) {
  DecideCfg cfg;
  cfg.V = V; cfg.inv_v = 0.f; cfg.fV = (float)V; cfg.rcp_v = 1.0f / (float)V; cfg.margin_abs = 0.f; cfg.mask_mode = 0; cfg.decide_mode = 0;
  unsigned long long bad = 0;
  const unsigned long long total = 1ull << 32;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const float a = __uint_as_float((unsigned int)i);
    const float q1 = pisto_div_views(a, cfg), q2 = __fdiv_rn(a, (float)V);
    if (__float_as_uint(q1) != __float_as_uint(q2) && !(q1 != q1 && q2 != q2)) bad++;
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int pisto_selftest_div(pisto_handle_t h, int V, unsigned long long* mismatches_dev, pisto_stream_t stream) {
  PISTO_REQUIRE(h && mismatches_dev && V >= 1, "pisto_selftest_div: bad argument");
  PISTO_CUDA(cudaSetDevice(h->device));
  selftest_div_kernel<<<h->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(V, mismatches_dev);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
