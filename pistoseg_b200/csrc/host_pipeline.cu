// pisto_fuse_argmax_confusion_host: the same fused call with HOST buffers -- what a caller holding numpy / CPU-torch
// data (the reference's DataLoader output) invokes, and what bench.py's end-to-end ("e2e") number times.
//
// Tiles are processed in chunks; chunk k uses device slot k%3 and stream k%3:  H2D of the chunk's view slices and
// byte masks -> fused kernel -> D2H of labels / 32x32 logits, so that the copy engines (both directions) and the
// SMs work on different chunks at the same time.  Host buffers should be pinned (cudaHostAlloc / torch pin_memory)
// for the copies to be asynchronous; pageable memory still works, serialised by the driver.
// View logits may also be DEVICE pointers (the reference's own dataflow: the backbone output never leaves the GPU, only
// `tissue` comes from the DataLoader and labels / 32x32 logits go back -- infer_pseudo_masks.py:119-137); such views are used in
// place, only the byte masks are uploaded.
#include "fuse_common.cuh"

int pisto_build_fuse_params(const pisto_view_t* views, int V, const pisto_fuse_args_t* a, FuseParams* out, bool* low_via_resize);

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int ensure_pipe(pisto_ctx* h, size_t bytes) {
  if (!h->pipe_ready) {
    for (int i = 0; i < PISTO_PIPE_SLOTS; i++) {
      PISTO_CUDA(cudaStreamCreateWithFlags(&h->pipe_stream[i], cudaStreamNonBlocking));
      PISTO_CUDA(cudaEventCreateWithFlags(&h->pipe_done[i], cudaEventDisableTiming));
      PISTO_CUDA(cudaEventCreate(&h->pipe_t1[i]));
    }
    PISTO_CUDA(cudaEventCreate(&h->pipe_t0));
    h->pipe_ready = true;
  }
  for (int i = 0; i < PISTO_PIPE_SLOTS; i++) {
    if (h->pipe_dev_bytes[i] < bytes) {
      if (h->pipe_dev[i]) { PISTO_CUDA(cudaStreamSynchronize(h->pipe_stream[i])); PISTO_CUDA(cudaFree(h->pipe_dev[i])); h->pipe_dev[i] = nullptr; h->pipe_dev_bytes[i] = 0; }
      PISTO_CUDA(cudaMalloc(&h->pipe_dev[i], bytes));
      h->pipe_dev_bytes[i] = bytes;
    }
  }
  return PISTO_OK;
}

}  // namespace

extern "C" int pisto_fuse_argmax_confusion_host(pisto_handle_t h, const pisto_view_t* views, int V, const pisto_fuse_args_t* a, int chunk) {
  PISTO_REQUIRE(h, "pisto_fuse_argmax_confusion_host: NULL handle");
  PISTO_REQUIRE(views && a, "pisto_fuse_argmax_confusion_host: views/args NULL");
  PISTO_REQUIRE(V >= 1 && V <= PISTO_MAX_VIEWS, "pisto_fuse_argmax_confusion_host: V=%d outside [1,%d]", V, PISTO_MAX_VIEWS);
  PISTO_REQUIRE(a->C >= 1 && a->C <= PISTO_MAX_CLASSES && a->N >= 0 && a->T_h >= 1 && a->T_w >= 1, "pisto_fuse_argmax_confusion_host: bad shape");
  if (a->N == 0) return PISTO_OK;
  if (chunk <= 0) chunk = 1024;
  if (chunk > a->N) chunk = a->N;
  PISTO_CUDA(cudaSetDevice(h->device));
  const int C = a->C;
  const size_t px = (size_t)a->T_h * a->T_w;
  // device slot layout (per chunk), every region 256-byte aligned
  size_t off = 0;
  size_t view_off[PISTO_MAX_VIEWS], view_tile_bytes[PISTO_MAX_VIEWS];
  bool view_on_device[PISTO_MAX_VIEWS];
  for (int v = 0; v < V; v++) {
    cudaPointerAttributes pa;
    view_on_device[v] = views[v].logits && cudaPointerGetAttributes(&pa, views[v].logits) == cudaSuccess &&
                        (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    cudaGetLastError();  // an unregistered (pageable) host pointer is not an error here
    PISTO_REQUIRE(views[v].logits && views[v].h >= 1 && views[v].w >= 1, "pisto_fuse_argmax_confusion_host: view %d invalid", v);
    PISTO_REQUIRE(views[v].tile_stride == 0 || views[v].tile_stride == (int64_t)C * views[v].h * views[v].w,
                  "pisto_fuse_argmax_confusion_host: strided host views are not supported");
    view_tile_bytes[v] = (size_t)C * views[v].h * views[v].w * sizeof(float);
    view_off[v] = off;
    if (!view_on_device[v]) off = align_up(off + view_tile_bytes[v] * chunk, 256);
  }
  PISTO_REQUIRE(!a->label_raw_out, "pisto_fuse_argmax_confusion_host: label_raw_out is not supported by the host-buffer variant");
  size_t o_present = off; if (a->present) off = align_up(off + (size_t)chunk * C, 256);
  size_t o_bg = off; if (a->bg) off = align_up(off + px * chunk, 256);
  size_t o_gt = off; if (a->gt) off = align_up(off + px * chunk, 256);
  size_t o_label = off; if (a->label_out) off = align_up(off + px * chunk, 256);
  size_t o_fused = off; if (a->fused_out) off = align_up(off + px * chunk * C * sizeof(float), 256);
  size_t o_ent = off; if (a->entropy_out) off = align_up(off + px * chunk * sizeof(float), 256);
  size_t low_px = (size_t)a->low_h * a->low_w;
  size_t o_low = off; if (a->lowres_out) off = align_up(off + low_px * chunk * C * sizeof(float), 256);
  size_t o_conf = off; if (a->conf) off = align_up(off + (size_t)C * C * sizeof(unsigned long long), 256);
  int rc = ensure_pipe(h, off);
  if (rc != PISTO_OK) return rc;

  // device-clock timing of the whole call: t0 on stream 0 (stream 1 waits for it), one end event per stream
  PISTO_CUDA(cudaEventRecord(h->pipe_t0, h->pipe_stream[0]));
  for (int s = 1; s < PISTO_PIPE_SLOTS; s++) PISTO_CUDA(cudaStreamWaitEvent(h->pipe_stream[s], h->pipe_t0, 0));
  if (a->conf) for (int s = 0; s < PISTO_PIPE_SLOTS; s++) PISTO_CUDA(cudaMemsetAsync((char*)h->pipe_dev[s] + o_conf, 0, (size_t)C * C * sizeof(unsigned long long), h->pipe_stream[s]));

  int k = 0;
  for (int n0 = 0; n0 < a->N; n0 += chunk, k++) {
    const int s = k % PISTO_PIPE_SLOTS;
    const int nn = a->N - n0 < chunk ? a->N - n0 : chunk;
    cudaStream_t st = h->pipe_stream[s];
    char* d = (char*)h->pipe_dev[s];
    pisto_view_t dv[PISTO_MAX_VIEWS];
    for (int v = 0; v < V; v++) {
      dv[v] = views[v];
      dv[v].tile_stride = 0;
      if (view_on_device[v]) {  // already on the GPU (the backbone's output): used in place
        dv[v].logits = (const float*)((const char*)views[v].logits + view_tile_bytes[v] * n0);
      } else {
        dv[v].logits = (const float*)(d + view_off[v]);
        PISTO_CUDA(cudaMemcpyAsync(d + view_off[v], (const char*)views[v].logits + view_tile_bytes[v] * n0, view_tile_bytes[v] * nn, cudaMemcpyHostToDevice, st));
      }
    }
    pisto_fuse_args_t da = *a;
    da.N = nn;
    if (a->present) { PISTO_CUDA(cudaMemcpyAsync(d + o_present, a->present + (size_t)n0 * C, (size_t)nn * C, cudaMemcpyHostToDevice, st)); da.present = (const uint8_t*)(d + o_present); }
    if (a->bg) { PISTO_CUDA(cudaMemcpyAsync(d + o_bg, a->bg + px * n0, px * nn, cudaMemcpyHostToDevice, st)); da.bg = (const uint8_t*)(d + o_bg); }
    if (a->gt) { PISTO_CUDA(cudaMemcpyAsync(d + o_gt, a->gt + px * n0, px * nn, cudaMemcpyHostToDevice, st)); da.gt = (const uint8_t*)(d + o_gt); }
    if (a->label_out) da.label_out = (uint8_t*)(d + o_label);
    if (a->fused_out) da.fused_out = (float*)(d + o_fused);
    if (a->entropy_out) da.entropy_out = (float*)(d + o_ent);
    if (a->lowres_out) da.lowres_out = (float*)(d + o_low);
    if (a->conf) da.conf = (unsigned long long*)(d + o_conf);
    rc = pisto_fuse_argmax_confusion(h, dv, V, &da, (pisto_stream_t)st);
    if (rc != PISTO_OK) return rc;
    if (a->label_out) PISTO_CUDA(cudaMemcpyAsync(a->label_out + px * n0, d + o_label, px * nn, cudaMemcpyDeviceToHost, st));
    if (a->fused_out) PISTO_CUDA(cudaMemcpyAsync(a->fused_out + px * n0 * C, d + o_fused, px * nn * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (a->entropy_out) PISTO_CUDA(cudaMemcpyAsync(a->entropy_out + px * n0, d + o_ent, px * nn * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (a->lowres_out) PISTO_CUDA(cudaMemcpyAsync(a->lowres_out + low_px * n0 * C, d + o_low, low_px * nn * C * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  unsigned long long part[PISTO_PIPE_SLOTS][PISTO_MAX_CLASSES * PISTO_MAX_CLASSES];
  for (int s = 0; s < PISTO_PIPE_SLOTS; s++) {
    if (a->conf) PISTO_CUDA(cudaMemcpyAsync(part[s], (char*)h->pipe_dev[s] + o_conf, (size_t)C * C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->pipe_stream[s]));
    PISTO_CUDA(cudaEventRecord(h->pipe_t1[s], h->pipe_stream[s]));
    PISTO_CUDA(cudaStreamSynchronize(h->pipe_stream[s]));
  }
  {
    float mx = 0.f;
    for (int s = 0; s < PISTO_PIPE_SLOTS; s++) {
      float m = 0.f;
      PISTO_CUDA(cudaEventElapsedTime(&m, h->pipe_t0, h->pipe_t1[s]));
      mx = m > mx ? m : mx;
    }
    h->pipe_last_ms = mx;
  }
  if (a->conf) for (int i = 0; i < C * C; i++) for (int s = 0; s < PISTO_PIPE_SLOTS; s++) a->conf[i] += part[s][i];
  return PISTO_OK;
}

extern "C" double pisto_last_pipeline_ms(pisto_handle_t h) { return h ? (double)h->pipe_last_ms : 0.0; }
