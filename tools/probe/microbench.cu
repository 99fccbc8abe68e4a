// Round-1 hardware probe (not product code): FP32 pipe issue rates on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_ffma(float* out, float a, float b) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma3(float* out, float a, float b) {  // mul + fma + add pattern (the exact-mode inner op)
  float x[ILP], h[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 1e-3f + i; h[i] = x[i] * 0.5f; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) { float t = __fmul_rn(b, h[i]); float o = __fmaf_rn(a, x[i], t); x[i] = __fadd_rn(x[i], o); }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo); }

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[ILP];
  unsigned long long aa = pack(a, a), bb = pack(b, b);
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = pack(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma2(x[i], aa, bb);
  }
  unsigned long long s = 0; for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s) + __uint_as_float((unsigned)(s >> 32));
}
__global__ void k_ffma3x2(float* out, float a, float b) {  // packed mul2 + fma2 + add2
  unsigned long long x[ILP], h[ILP];
  unsigned long long aa = pack(a, a), bb = pack(b, b);
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = pack(threadIdx.x * 1e-3f + i, i); h[i] = pack(i * 0.5f, 1.f); }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) { unsigned long long t = mul2(bb, h[i]); unsigned long long o = fma2(aa, x[i], t); x[i] = add2(x[i], o); }
  }
  unsigned long long s = 0; for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s) + __uint_as_float((unsigned)(s >> 32));
}
__global__ void k_lds(float* out, int stride) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float s = 0; int idx = (threadIdx.x / 8) & 1023;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) { s += sm[(idx + i * stride) & 4095]; }
    idx = (idx + 1) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: 3 FP32 ops per 1 uniform LDS.64 (models per-(view,row) weight fetch)
__global__ void k_mix(float* out, float a) {
  __shared__ float2 tab[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_float2(0.25f + i * 1e-4f, 0.75f - i * 1e-4f);
  __syncthreads();
  float x[6], h[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { x[i] = threadIdx.x * 1e-3f + i; h[i] = x[i] * 0.5f; }
  for (int it = 0; it < ITERS; it++) {
    float2 w = tab[it & 511];
#pragma unroll
    for (int i = 0; i < 6; i++) { float t = __fmul_rn(w.y, h[i]); float o = __fmaf_rn(w.x, x[i], t); x[i] = __fadd_rn(x[i] * a, o); }
  }
  float s = 0; for (int i = 0; i < 6; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_copy(const int4* __restrict__ in, int4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}
__global__ void k_read(const int4* __restrict__ in, int* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (; i < n; i += st) { int4 v = in[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) out[0] = acc;
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s sms %d clock_khz %d smem_optin %zu l2 %d\n", p.name, p.multiProcessorCount, clk_khz, p.sharedMemPerBlockOptin, p.l2CacheSize);
  int sms = p.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
  for (int threads : {256, 512, 1024}) {
    int blocks = sms * (2048 / threads);
    double n_thr = (double)blocks * threads;
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("threads/blk %4d FFMA     : %.1f G thread-instr/s  (%.2f instr/clk/SM @%d MHz)\n", threads, n_thr * ITERS * ILP / ms / 1e6, n_thr * ITERS * ILP / ms / 1e3 / sms / (clk_khz), clk_khz / 1000);
    ms = timeit([&] { k_ffma3<<<blocks, threads>>>(out, 0.5f, 0.25f); });
    printf("threads/blk %4d MUL+FMA+ADD: %.1f G thread-instr/s (%.2f instr/clk/SM)\n", threads, n_thr * ITERS * ILP * 3 / ms / 1e6, n_thr * ITERS * ILP * 3 / ms / 1e3 / sms / clk_khz);
    ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("threads/blk %4d FFMA2    : %.1f G thread-instr/s  (%.2f instr/clk/SM; x2 lanes)\n", threads, n_thr * ITERS * ILP / ms / 1e6, n_thr * ITERS * ILP / ms / 1e3 / sms / clk_khz);
    ms = timeit([&] { k_ffma3x2<<<blocks, threads>>>(out, 0.5f, 0.25f); });
    printf("threads/blk %4d MUL2+FMA2+ADD2: %.1f G thread-instr/s (%.2f instr/clk/SM; x2 lanes)\n", threads, n_thr * ITERS * ILP * 3 / ms / 1e6, n_thr * ITERS * ILP * 3 / ms / 1e3 / sms / clk_khz);
    ms = timeit([&] { k_lds<<<blocks, threads>>>(out, 1); });
    printf("threads/blk %4d LDS.32(8-lane bcast): %.1f G thread-instr/s (%.2f instr/clk/SM)\n", threads, n_thr * ITERS * ILP / ms / 1e6, n_thr * ITERS * ILP / ms / 1e3 / sms / clk_khz);
    ms = timeit([&] { k_mix<<<blocks, threads>>>(out, 0.999f); });
    printf("threads/blk %4d MIX(1 LDS.64 + 24 fp32): %.1f G thread-instr/s (%.2f instr/clk/SM)\n", threads, n_thr * ITERS * 25 / ms / 1e6, n_thr * ITERS * 25 / ms / 1e3 / sms / clk_khz);
  }
  size_t bytes = (size_t)2 << 30; int4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  for (int mult : {4, 8, 16}) {
    float ms = timeit([&] { k_copy<<<sms * mult, 512>>>(a, b, bytes / 16); });
    printf("copy grid %dxSM: %.1f GB/s (r+w)\n", mult, 2.0 * bytes / ms / 1e6);
    ms = timeit([&] { k_read<<<sms * mult, 512>>>(a, (int*)out, bytes / 16); });
    printf("read grid %dxSM: %.1f GB/s\n", mult, 1.0 * bytes / ms / 1e6);
  }
  return 0;
}
