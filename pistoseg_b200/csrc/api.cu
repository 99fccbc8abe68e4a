// Handle lifetime, error reporting, launch accounting for libpistoseg_b200 (C ABI: include/pistoseg_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void pisto_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pisto_last_error(void) { return g_err; }
extern "C" int pisto_abi_version(void) { return PISTO_ABI_VERSION; }

extern "C" int pisto_create(pisto_handle_t* out, int device) {
  if (!out) { pisto_set_error("pisto_create: out is NULL"); return PISTO_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    pisto_set_error("pisto_create: no CUDA device visible (%s); libpistoseg_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
    return PISTO_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { pisto_set_error("pisto_create: device %d out of range [0,%d)", device, n); return PISTO_ERR_INVALID; }
  cudaDeviceProp p;
  PISTO_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    pisto_set_error("pisto_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, p.major, p.minor);
    return PISTO_ERR_NO_DEVICE;
  }
  pisto_ctx* c = new pisto_ctx();  // value-initialised: every member zero (the atomics included)
  c->device = device;
  c->sm_count = p.multiProcessorCount;
  c->smem_optin = (int)p.sharedMemPerBlockOptin;
  PISTO_CUDA(cudaSetDevice(device));
  if (cudaMalloc(&c->sched, PISTO_SCHED_SLOTS * sizeof(int)) != cudaSuccess) {
    delete c;
    pisto_set_error("pisto_create: cudaMalloc failed");
    return PISTO_ERR_CUDA;
  }
  if (cudaMalloc(&c->stats, 4 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(c->stats, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
    cudaFree(c->sched);
    delete c;
    pisto_set_error("pisto_create: cudaMalloc failed");
    return PISTO_ERR_CUDA;
  }
  *out = c;
  return PISTO_OK;
}

extern "C" int pisto_destroy(pisto_handle_t h) {
  if (!h) return PISTO_OK;
  if (h->sched) { cudaSetDevice(h->device); cudaFree(h->sched); }
  if (h->stats) { cudaSetDevice(h->device); cudaFree(h->stats); }
  for (int i = 0; i < PISTO_SCHED_SLOTS; i++)
    if (h->sched_done[i]) cudaEventDestroy(h->sched_done[i]);
  if (h->pipe_ready) {
    cudaSetDevice(h->device);
    for (int i = 0; i < PISTO_PIPE_SLOTS; i++) {
      if (h->pipe_dev[i]) cudaFree(h->pipe_dev[i]);
      if (h->pipe_done[i]) cudaEventDestroy(h->pipe_done[i]);
      if (h->pipe_t1[i]) cudaEventDestroy(h->pipe_t1[i]);
      if (h->pipe_stream[i]) cudaStreamDestroy(h->pipe_stream[i]);
    }
    if (h->pipe_t0) cudaEventDestroy(h->pipe_t0);
  }
  delete h;
  return PISTO_OK;
}

extern "C" int64_t pisto_launch_count(pisto_handle_t h) { return h ? (int64_t)h->launches.load() : 0; }

extern "C" int pisto_filter_stats(pisto_handle_t h, unsigned long long* out_host, int reset) {
  PISTO_REQUIRE(h && out_host, "pisto_filter_stats: NULL argument");
  PISTO_CUDA(cudaSetDevice(h->device));
  PISTO_CUDA(cudaDeviceSynchronize());
  PISTO_CUDA(cudaMemcpy(out_host, h->stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) PISTO_CUDA(cudaMemset(h->stats, 0, 4 * sizeof(unsigned long long)));
  return PISTO_OK;
}

int pisto_sched_acquire(pisto_ctx* h, cudaStream_t st, int** counter, int* slot) {
  const int s = (int)(h->sched_next.fetch_add(1u) % PISTO_SCHED_SLOTS);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  PISTO_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap == cudaStreamCaptureStatusNone && h->sched_done[s]) PISTO_CUDA(cudaStreamWaitEvent(st, h->sched_done[s], 0));
  *counter = h->sched + s;
  *slot = s;
  PISTO_CUDA(cudaMemsetAsync(*counter, 0, sizeof(int), st));
  return PISTO_OK;
}

int pisto_sched_release(pisto_ctx* h, cudaStream_t st, int slot) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  PISTO_CUDA(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) return PISTO_OK;
  if (!h->sched_done[slot]) PISTO_CUDA(cudaEventCreateWithFlags(&h->sched_done[slot], cudaEventDisableTiming));
  PISTO_CUDA(cudaEventRecord(h->sched_done[slot], st));
  return PISTO_OK;
}
