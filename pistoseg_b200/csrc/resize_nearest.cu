// pisto_resize_nearest_bg: the tail of the revise-mask post-processing (infer_revise_masks.py:152-155,164-165,173-174):
//   mask = np.array(Image.fromarray(np.uint8(mask), mode='P').resize((w, h), resample=Image.BILINEAR));  mask[background > 0] = 3
// PIL resizes mode-'P' images with NEAREST whatever `resample` says, and its NEAREST scaler (ImagingScaleAffine) takes the source
// index of output column x from a double that is ACCUMULATED: xo = 0.5 * (n_in / n_out); xo += n_in / n_out per column -- so the
// index tables are built on the host in exactly that way (pistoseg_b200/postproc.py::pil_nearest_index, checked against
// Image.resize) and this kernel only gathers: out[y][x] = bg[y][x] > 0 ? bg_value : in[iy[y]][ix[x]], one CTA per tile, any
// original size per tile.
#include "common.cuh"

namespace {

__global__ void resize_nearest_bg_kernel(const uint8_t* __restrict__ in, int sh, int sw, const pisto_resize_desc_t* __restrict__ desc,
                                         const int32_t* __restrict__ index_pool, const uint8_t* __restrict__ bg_pool,
                                         uint8_t* __restrict__ out_pool, int bg_value) {
  const pisto_resize_desc_t d = desc[blockIdx.x];
  const uint8_t* src = in + (long long)d.tile * sh * sw;
  const int32_t* iy = index_pool + d.iy_off;
  const int32_t* ix = index_pool + d.ix_off;
  uint8_t* out = out_pool + d.out_off;
  const uint8_t* bg = (bg_pool && d.bg_off >= 0) ? bg_pool + d.bg_off : nullptr;
  const int n = d.h * d.w;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = i / d.w, x = i - y * d.w;
    unsigned int v = src[iy[y] * sw + ix[x]];
    if (bg && bg[i] > 0) v = (unsigned)bg_value;
    out[i] = (uint8_t)v;
  }
}

}  // namespace

extern "C" int pisto_resize_nearest_bg(pisto_handle_t h, const uint8_t* in, int n_tiles, int sh, int sw, const pisto_resize_desc_t* desc, int n_desc,
                                       const int32_t* index_pool, const uint8_t* bg_pool, uint8_t* out_pool, int bg_value, pisto_stream_t stream) {
  PISTO_REQUIRE(h, "pisto_resize_nearest_bg: NULL handle");
  PISTO_REQUIRE(in && desc && index_pool && out_pool, "pisto_resize_nearest_bg: NULL pointer");
  PISTO_REQUIRE(n_tiles >= 1 && sh >= 1 && sw >= 1 && n_desc >= 0, "pisto_resize_nearest_bg: bad shape");
  PISTO_REQUIRE(bg_value >= 0 && bg_value <= 255, "pisto_resize_nearest_bg: bg_value outside u8");
  if (n_desc == 0) return PISTO_OK;
  PISTO_CUDA(cudaSetDevice(h->device));
  resize_nearest_bg_kernel<<<n_desc, 256, 0, (cudaStream_t)stream>>>(in, sh, sw, desc, index_pool, bg_pool, out_pool, bg_value);
  h->launches++;
  PISTO_CUDA(cudaGetLastError());
  return PISTO_OK;
}
