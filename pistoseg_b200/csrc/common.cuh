// Shared device / host helpers for libpistoseg_b200 (sm_100a).
//
// Arithmetic contract (DESIGN.md "Bit-exactness"): every floating-point operation on the parity path is written
// with an explicit rounding intrinsic (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn) so that nvcc can neither
// contract nor re-associate it; the sequence of operations equals oracle/bilinear.py::bilinear_restated, which in
// turn equals torch's upsample_bilinear2d (CPU vectorised path and CUDA kernel, verified bit-for-bit in
// profiles/r01/probe_torch.json).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "pistoseg_b200.h"

#define PISTO_MAX_VIEWS 16
#define PISTO_MAX_CLASSES 8
#define PISTO_SCHED_SLOTS 256
#define PISTO_PIPE_SLOTS 3   // chunks in flight in the host-buffer pipeline (H2D of one, kernel of another, D2H of a third)

struct pisto_ctx {
  int device;
  int sm_count;
  int smem_optin;   // max dynamic shared memory per block
  std::atomic<long long> launches;
  unsigned long long* stats;  // [4] device counters of the filtered kernels: -, multi-label tiles, queued pixels, whole-tile exact fallbacks
  int* sched;       // ring of device tile counters for the persistent kernels' dynamic scheduler
  std::atomic<unsigned int> sched_next;
  cudaEvent_t sched_done[PISTO_SCHED_SLOTS];  // recorded after the kernel that used slot s: the next user of s waits for it (any stream)
  // resources of the host-buffer (e2e) pipeline, created lazily
  cudaStream_t pipe_stream[PISTO_PIPE_SLOTS];
  cudaEvent_t pipe_done[PISTO_PIPE_SLOTS];
  void* pipe_dev[PISTO_PIPE_SLOTS];
  size_t pipe_dev_bytes[PISTO_PIPE_SLOTS];
  cudaEvent_t pipe_t0, pipe_t1[PISTO_PIPE_SLOTS];  // timing of the last host-buffer call (device clock)
  float pipe_last_ms;
  bool pipe_ready;
};

void pisto_set_error(const char* fmt, ...);

// A persistent kernel's dynamic tile scheduler needs a zeroed device counter for the duration of the launch.  Counters come from a
// ring; a slot is re-armed (memset on the launch stream) only after the kernel that used it last has finished, on WHATEVER stream
// that was: the launch stream first waits for the slot's event.  Callers may therefore drive one handle from several streams or
// threads (slot choice and launch count are atomic); during stream capture the events are left alone (the graph orders its nodes).
int pisto_sched_acquire(pisto_ctx* h, cudaStream_t st, int** counter, int* slot);
int pisto_sched_release(pisto_ctx* h, cudaStream_t st, int slot);

#define PISTO_CUDA(call)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (call);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      pisto_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);    \
      return PISTO_ERR_CUDA;                                                                          \
    }                                                                                                 \
  } while (0)

#define PISTO_REQUIRE(cond, ...)       \
  do {                                 \
    if (!(cond)) {                     \
      pisto_set_error(__VA_ARGS__);    \
      return PISTO_ERR_INVALID;        \
    }                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// bilinear source index / lambda, align_corners = False (ATen UpSample.h: area_pixel_compute_source_index +
// compute_source_index_and_lambda; oracle/bilinear.py::source_index_and_lambda)
// ---------------------------------------------------------------------------------------------------
struct Lerp {
  int i0, i1;
  float l0, l1;
};

__host__ __device__ __forceinline__ Lerp pisto_src_index(float scale, int dst, int in_size, bool same) {
  Lerp r;
  if (same) {
    r.i0 = dst; r.i1 = dst; r.l0 = 1.f; r.l1 = 0.f;
    return r;
  }
#ifdef __CUDA_ARCH__
  float src = fmaxf(__fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f), 0.f);
#else
  float src = fmaxf(fmaf(scale, (float)dst + 0.5f, -0.5f), 0.f);
#endif
  int i0 = (int)src;  // src >= 0: truncation == floor
  if (i0 > in_size - 1) i0 = in_size - 1;
  r.i0 = i0;
  r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
#ifdef __CUDA_ARCH__
  float l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
  r.l1 = l1;
  r.l0 = __fsub_rn(1.f, l1);
#else
  float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  r.l1 = l1;
  r.l0 = 1.f - l1;
#endif
  return r;
}

// De-augmentation index map of a view (oracle/tta.py::code_affine): deaug(y)[i][j] = y[a][b] with
//   a = a0 + i*ai + j*aj,  b = b0 + i*bi + j*bj      (each coefficient in {0, +1, -1})
struct ViewMap {
  int ho, wo;  // de-augmented size
  int a0, ai, aj, b0, bi, bj;
};

__host__ __device__ inline ViewMap pisto_view_map(int code, int h, int w) {
  ViewMap m;
  int k = code & 3, hf = (code >> 2) & 1;
  switch (k) {
    case 0: m.ho = h; m.wo = w; m.a0 = 0; m.ai = 1; m.aj = 0; m.b0 = 0; m.bi = 0; m.bj = 1; break;
    case 1: m.ho = w; m.wo = h; m.a0 = 0; m.ai = 0; m.aj = 1; m.b0 = w - 1; m.bi = -1; m.bj = 0; break;
    case 2: m.ho = h; m.wo = w; m.a0 = h - 1; m.ai = -1; m.aj = 0; m.b0 = w - 1; m.bi = 0; m.bj = -1; break;
    default: m.ho = w; m.wo = h; m.a0 = h - 1; m.ai = 0; m.aj = -1; m.b0 = 0; m.bi = 1; m.bj = 0; break;
  }
  if (hf) {  // out[i][j] = R[i][wo-1-j]
    m.a0 += (m.wo - 1) * m.aj; m.aj = -m.aj;
    m.b0 += (m.wo - 1) * m.bj; m.bj = -m.bj;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------
// per-pixel decision (mask -> softmax / raw argmax -> entropy), shared by every fusion kernel
// ---------------------------------------------------------------------------------------------------
struct DecideCfg {
  int mask_mode;
  int decide_mode;
  int V;            // number of fused views (the divisor)
  float inv_v;      // 1/V when V is a power of two (exact), else 0
  float margin_abs; // 2e-6 * V : see finalize fast path
  float fV;         // (float)V
  float rcp_v;      // RN(1/V)
};

// a / V, correctly rounded (== IEEE division).  Power-of-two V: one multiply.  Otherwise the classic
// reciprocal + one fma-residual correction (Markstein): q = a*r, e = fma(-V, q, a), q' = fma(e, r, q) with
// r = RN(1/V); it returns the correctly rounded quotient whenever nothing under/overflows, which the magnitude guard
// ensures (outside it: __fdiv_rn).  pisto_selftest_div checks q' == __fdiv_rn over ALL 2^32 inputs for a given V
// (tests/test_gpu_ops.py::test_division_by_view_count_exhaustive).
__device__ __forceinline__ float pisto_div_views(float a, const DecideCfg& cfg) {
  if (cfg.V == 1) return a;
  if (cfg.inv_v != 0.f) return __fmul_rn(a, cfg.inv_v);  // power of two: same correctly-rounded quotient
  const float fa = fabsf(a);
  if (fa > 1e-30f && fa < 1e30f) {
    const float q = __fmul_rn(a, cfg.rcp_v);
    const float e = __fmaf_rn(-cfg.fV, q, a);
    return __fmaf_rn(e, cfg.rcp_v, q);
  }
  return __fdiv_rn(a, cfg.fV);
}

// Exact reference semantics for one pixel.  s[] are the fused scores (already divided by V).
// Returns the label (before background); writes entropy if asked.
template <int C>
__device__ __noinline__ int pisto_decide_slow(const float (&s)[C], uint32_t present_bits, const DecideCfg& cfg,
                                              bool want_entropy, float* entropy) {
  float x[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    bool pres = (present_bits >> c) & 1u;
    float v = s[c];
    if (cfg.mask_mode == PISTO_MASK_FILL) v = pres ? v : -1e10f;
    else if (cfg.mask_mode == PISTO_MASK_NEG_INF) v = pres ? v : -INFINITY;
    else if (cfg.mask_mode == PISTO_MASK_MULTIPLY) v = __fmul_rn(v, pres ? 1.f : 0.f);
    x[c] = v;
  }
  if (cfg.decide_mode == PISTO_DECIDE_RAW && !want_entropy) {
    // torch.argmax / np.argmax: first maximum, NaN counts as the maximum
    int bi = 0;
    float bv = x[0];
#pragma unroll
    for (int c = 1; c < C; c++) {
      bool take = (x[c] > bv) || (x[c] != x[c] && bv == bv);
      if (take) { bv = x[c]; bi = c; }
    }
    return bi;
  }
  // softmax over the class axis exactly as ATen evaluates it: max, sum of exp(x - max) in index order, divide
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; c++) m = fmaxf(m, x[c]);
  float e[C];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) {
    e[c] = expf(__fsub_rn(x[c], m));
    sum = __fadd_rn(sum, e[c]);
  }
  float p[C];
#pragma unroll
  for (int c = 0; c < C; c++) p[c] = __fdiv_rn(e[c], sum);
  if (want_entropy) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) acc = __fadd_rn(acc, __fmul_rn(p[c], logf(__fadd_rn(p[c], 1e-10f))));
    *entropy = -acc;
  }
  const float* d = (cfg.decide_mode == PISTO_DECIDE_RAW) ? x : p;
  int bi = 0;
  float bv = d[0];
#pragma unroll
  for (int c = 1; c < C; c++) {
    bool take = (d[c] > bv) || (d[c] != d[c] && bv == bv);
    if (take) { bv = d[c]; bi = c; }
  }
  return bi;
}

// a[] are the UNDIVIDED sums over views.  Fast path: when the best present class leads the runner-up by more than a
// margin, argmax(softmax(a / V)) == argmax(a) provably (softmax is strictly ordered once the logit gap exceeds ~1e-6:
// exp(-1e-6) is >= 8 float ulps below 1 and the IEEE division keeps the order), so neither the division nor the
// exponentials are evaluated.  Anything else (near ties, exact ties, NaN, absurd magnitudes, entropy wanted,
// MULTIPLY mask) takes pisto_decide_slow, which follows the reference operation by operation.
template <int C>
__device__ __forceinline__ int pisto_decide(const float (&a)[C], uint32_t present_bits, const DecideCfg& cfg,
                                            bool want_entropy, float* entropy) {
  if (!want_entropy && cfg.mask_mode != PISTO_MASK_MULTIPLY) {
    const bool masked = cfg.mask_mode != PISTO_MASK_NONE;
    int bi = -1;
    float bv = -INFINITY, sv = -INFINITY, nan_probe = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
      bool pres = !masked || ((present_bits >> c) & 1u);
      float v = pres ? a[c] : -INFINITY;
      nan_probe = __fadd_rn(nan_probe, pres ? a[c] : 0.f);
      if (v > bv) { sv = bv; bv = v; bi = c; }
      else sv = fmaxf(sv, v);
    }
    // gap in the divided domain must exceed 1e-6 + 2 ulp(x): margin = V*2e-6 + 2^-22 * |a_best|
    float margin = __fmaf_rn(fabsf(bv), 2.4e-7f, cfg.margin_abs);
    bool ok = (bi >= 0) && (__fsub_rn(bv, sv) > margin) && (nan_probe == nan_probe) && (fabsf(nan_probe) < 1e30f) &&
              (bv > -1e9f);
    if (cfg.decide_mode == PISTO_DECIDE_RAW) {
      // raw argmax of a/V: division is monotone, so only exact ties after rounding could differ -> same margin test,
      // but a tie in a[] itself resolves to the lowest index either way.
      ok = ok || ((bi >= 0) && (nan_probe == nan_probe) && (fabsf(nan_probe) < 1e30f) && cfg.V == 1);
    }
    if (ok) return bi;
  }
  float s[C];
#pragma unroll
  for (int c = 0; c < C; c++) s[c] = pisto_div_views(a[c], cfg);
  return pisto_decide_slow<C>(s, present_bits, cfg, want_entropy, entropy);
}

// channel softmax of one pixel, in place (PROB_MEAN views; pisto_stitch_accumulate)
template <int C>
__device__ __forceinline__ void pisto_softmax_inplace(float (&x)[C]) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; c++) m = fmaxf(m, x[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) {
    x[c] = expf(__fsub_rn(x[c], m));
    sum = __fadd_rn(sum, x[c]);
  }
#pragma unroll
  for (int c = 0; c < C; c++) x[c] = __fdiv_rn(x[c], sum);
}

// Same softmax for the streaming kernels' PROB_MEAN path, where V*C*T^2 exponentials per tile make the MUFU unit the
// bottleneck: exp(d) as ONE ex2.approx.ftz of d * log2(e) (relative error < 2^-22 on the argument range of a max-subtracted
// softmax; results below 2^-126 flush to zero, i.e. < 1e-38 absolute) and ONE rcp.approx.ftz instead of C divisions (the correctly
// rounded __frcp_rn / __expf expand to ~10 instructions with branches each: they were 45 % of this kernel's instructions).
// Probabilities differ from the exact version by < 1e-6 absolute -- inside the 1e-5 float gate that applies to PROB_MEAN scores
// anyway (CUDA expf and the reference's Sleef expf already differ by 1 ulp).
__device__ __forceinline__ float pisto_ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float pisto_rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int C>
__device__ __forceinline__ void pisto_softmax_fast(float (&x)[C]) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; c++) m = fmaxf(m, x[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) {
    x[c] = pisto_ex2_approx(__fmul_rn(__fsub_rn(x[c], m), 1.4426950408889634f));
    sum = c == 0 ? x[c] : __fadd_rn(sum, x[c]);
  }
  const float r = pisto_rcp_approx(sum);   // sum is in [1, C]: no denormal / overflow case
#pragma unroll
  for (int c = 0; c < C; c++) x[c] = __fmul_rn(x[c], r);
}
