// Generic fusion kernel: one thread per output pixel, any V <= 16, C <= 8, any view / tile size, every option.
// It is the correctness anchor of the library (simple enough to audit against oracle/fuse.py line by line) and the
// in-library path for shapes the streaming kernel (fuse_stream.cu) has no instantiation for.  Both kernels run the
// same arithmetic (common.cuh) and tests/test_gpu_fuse.py requires them to agree bit-for-bit.
#include "fuse_common.cuh"

namespace {

constexpr int kThreads = 256;

template <int C>
__global__ void __launch_bounds__(kThreads) fuse_generic_kernel(const __grid_constant__ FuseParams p) {
  __shared__ unsigned int hist[C * C];
  const bool do_conf = p.conf != nullptr && p.gt != nullptr;
  if (do_conf) {
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
  }
  const long long px_per_tile = (long long)p.T_h * p.T_w;
  const long long total = px_per_tile * p.N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int n = (int)(idx / px_per_tile);
    const int rem = (int)(idx - (long long)n * px_per_tile);
    const int y = rem / p.T_w, x = rem - y * p.T_w;
    const TilePresence tp = pisto_tile_presence(p, n);

    const bool low_hit = p.lowres_out && p.low_fh > 0 && (y % p.low_fh == p.low_fh / 2) && (x % p.low_fw == p.low_fw / 2);
    const bool need_scores = tp.single < 0 || p.fused_out || low_hit;  // single-label tiles never read the scores
    float a[C];
    if (need_scores) {
      for (int v = 0; v < p.V; v++) {
        const ViewDev& vw = p.view[v];
        Lerp ly = pisto_src_index(vw.scale_h, y, vw.map.ho, vw.same_h);
        Lerp lx = pisto_src_index(vw.scale_w, x, vw.map.wo, vw.same_w);
        float u[C];
#pragma unroll
        for (int c = 0; c < C; c++) u[c] = pisto_sample_view(vw, C, n, c, ly, lx);
        if (p.fuse_mode == PISTO_FUSE_PROB_MEAN) pisto_softmax_inplace<C>(u);
#pragma unroll
        for (int c = 0; c < C; c++) a[c] = (v == 0) ? u[c] : __fadd_rn(a[c], u[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; c++) a[c] = 0.f;
    }
    if (p.fused_out || low_hit) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        float s = pisto_div_views(a[c], p.dec);
        if (p.fused_out) p.fused_out[((long long)n * C + c) * px_per_tile + rem] = s;
        if (low_hit) p.lowres_out[(((long long)n * C + c) * p.low_h + y / p.low_fh) * p.low_w + x / p.low_fw] = s;
      }
    }
    int lab;
    float ent = 0.f;
    if (tp.single >= 0) lab = tp.single;
    else lab = pisto_decide<C>(a, tp.bits, p.dec, p.entropy_out != nullptr, &ent);
    if (p.entropy_out) p.entropy_out[idx] = ent;
    if (do_conf) {
      unsigned int g = p.gt[idx];
      if (g < (unsigned)C) atomicAdd(&hist[g * C + lab], 1u);
    }
    if (p.label_out) {
      int out = lab;
      if (p.bg && p.bg[idx] == (uint8_t)p.bg_match) out = p.bg_label;
      p.label_out[idx] = (uint8_t)out;
    }
    if (p.label_raw_out) {
      int out = tp.single;
      if (tp.single < 0) {
        DecideCfg raw = p.dec;
        raw.decide_mode = PISTO_DECIDE_RAW;
        out = pisto_decide<C>(a, tp.bits, raw, false, nullptr);
      }
      if (p.bg && p.bg[idx] == (uint8_t)p.bg_match) out = p.bg_label;
      p.label_raw_out[idx] = (uint8_t)out;
    }
  }
  if (do_conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  }
}

template <int C>
int launch_c(pisto_ctx* h, const FuseParams& p, cudaStream_t st) {
  long long total = (long long)p.N * p.T_h * p.T_w;
  long long grid = (total + kThreads - 1) / kThreads;
  long long cap = (long long)h->sm_count * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  fuse_generic_kernel<C><<<(int)grid, kThreads, 0, st>>>(p);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}

}  // namespace

int pisto_launch_fuse_generic(pisto_ctx* h, const FuseParams& p, cudaStream_t st) {
  switch (p.C) {
    case 1: return launch_c<1>(h, p, st);
    case 2: return launch_c<2>(h, p, st);
    case 3: return launch_c<3>(h, p, st);
    case 4: return launch_c<4>(h, p, st);
    case 5: return launch_c<5>(h, p, st);
    case 6: return launch_c<6>(h, p, st);
    case 7: return launch_c<7>(h, p, st);
    case 8: return launch_c<8>(h, p, st);
  }
  pisto_set_error("pisto_fuse_argmax_confusion: C=%d outside [1,%d]", p.C, PISTO_MAX_CLASSES);
  return PISTO_ERR_UNSUPPORTED;
}
