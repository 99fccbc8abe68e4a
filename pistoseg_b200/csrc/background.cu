// Background mask of an RGB tile -- replaces utils.get_background / BaseDataset._get_background
// (utils.py:155-163, dataset.py:100-109):
//     gray   = cv2.cvtColor(region, COLOR_RGB2GRAY)            8-bit fixed point: (9798 R + 19235 G + 3735 B + 2^14) >> 15
//     binary = gray > thresh                                    cv2.threshold(gray, 200, 255, THRESH_BINARY)
//     keep   = remove_small_objects(binary, min_size, connectivity=1)   4-connected components with fewer than min_size
//                                                               pixels are cleared
//     mask   = keep * 255
// Connected components: union-find over the pixel indices of one image (label = smallest pixel index of the component,
// merged with atomicMin), then one counting pass.  Exact integer work: the result equals the reference's bit for bit.
#include "common.cuh"

namespace {

__device__ __forceinline__ int find_root(const int* L, int i) {
  int r = L[i];
  while (r != i) { i = r; r = L[i]; }
  return r;
}

__device__ __forceinline__ void unite(int* L, int a, int b) {
  for (;;) {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);   // b was a root: hook it under the smaller root
    if (old == b) return;
    b = old;                               // somebody re-rooted b meanwhile: continue from where it points now
  }
}

// label[p] = p for foreground pixels, -1 for background; cnt = 0
__global__ void bg_init_kernel(const uint8_t* __restrict__ rgb, long long n_px, int thresh, int* __restrict__ label, int* __restrict__ cnt) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    const unsigned r = rgb[3 * p], g = rgb[3 * p + 1], b = rgb[3 * p + 2];
    const unsigned gray = (r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15;
    label[p] = (int)gray > thresh ? (int)p : -1;
    cnt[p] = 0;
  }
}

// 4-connectivity: every foreground pixel is united with its right and lower foreground neighbours (same image only)
__global__ void bg_merge_kernel(int* __restrict__ label, long long n_px, int H, int W) {
  const long long hw = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    if (label[p] < 0) continue;
    const long long q = p % hw;
    const int y = (int)(q / W), x = (int)(q - (long long)y * W);
    if (x + 1 < W && label[p + 1] >= 0) unite(label, (int)p, (int)p + 1);
    if (y + 1 < H && label[p + W] >= 0) unite(label, (int)p, (int)p + W);
  }
}

__global__ void bg_count_kernel(int* __restrict__ label, long long n_px, int* __restrict__ cnt) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    if (label[p] < 0) continue;
    const int r = find_root(label, (int)p);
    label[p] = r;  // flatten: a concurrent find_root that passes through p sees its old parent or the root, both ancestors
    atomicAdd(&cnt[r], 1);
  }
}

__global__ void bg_write_kernel(const int* __restrict__ label, const int* __restrict__ cnt, long long n_px, int min_size, uint8_t* __restrict__ out) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    uint8_t v = 0;
    if (label[p] >= 0) v = cnt[label[p]] >= min_size ? 255 : 0;  // flattened by the counting pass
    out[p] = v;
  }
}

}  // namespace

extern "C" int pisto_get_background(pisto_handle_t h, const uint8_t* rgb, int N, int H, int W, int thresh, int min_size,
                                    int32_t* scratch, uint8_t* mask_out, pisto_stream_t stream) {
  if (!h) { pisto_set_error("pisto_get_background: handle is NULL"); return PISTO_ERR_INVALID; }
  if (N < 0 || H <= 0 || W <= 0) { pisto_set_error("pisto_get_background: bad shape N=%d H=%d W=%d", N, H, W); return PISTO_ERR_INVALID; }
  const long long n_px = (long long)N * H * W;
  if (n_px == 0) return PISTO_OK;
  if (n_px > 0x7fffffffLL) { pisto_set_error("pisto_get_background: %lld pixels exceed the 32-bit label range; split the batch", n_px); return PISTO_ERR_UNSUPPORTED; }
  if (!rgb || !scratch || !mask_out) { pisto_set_error("pisto_get_background: NULL buffer"); return PISTO_ERR_INVALID; }
  PISTO_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int* label = scratch;
  int* cnt = scratch + n_px;
  const int threads = 256;
  long long want = (n_px + threads - 1) / threads;
  const int grid = (int)(want < (long long)h->sm_count * 16 ? want : (long long)h->sm_count * 16);
  bg_init_kernel<<<grid, threads, 0, st>>>(rgb, n_px, thresh, label, cnt);
  bg_merge_kernel<<<grid, threads, 0, st>>>(label, n_px, H, W);
  bg_count_kernel<<<grid, threads, 0, st>>>(label, n_px, cnt);
  bg_write_kernel<<<grid, threads, 0, st>>>(label, cnt, n_px, min_size, mask_out);
  h->launches += 4;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
