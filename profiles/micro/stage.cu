// Micro-benchmark of the weight staging pipeline of the row-tile kernels (no compute):
// every CTA streams the same [N x K] fp32 matrix through shared memory in [256 x 32] chunks.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int KC = 32, LW = 36, NCH = 256;

// STAGES-deep ring; if TWO_BARRIERS, sync before and after the (empty) compute phase
template <int NT, int STAGES, bool TWO_BARRIERS, int WORK>
__global__ void __launch_bounds__(NT, 1) stage_kernel(const float* W, int K, int nchunks, float* out) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int nk = K / KC;
  constexpr int RPI = NT / 8;
  const int lq = tid % 8, lr0 = tid / 8;
  auto load = [&](int c, int stage) {
    const int kci = c % nk, nci = (c / nk) % 1;
    float* dst = sm + stage * NCH * LW;
    const float* src0 = W + (size_t)(nci * NCH) * K + kci * KC + 4 * lq;
#pragma unroll
    for (int it = 0; it < NCH / RPI; ++it) {
      const int row = lr0 + it * RPI;
      cp_async16(dst + row * LW + 4 * lq, src0 + (size_t)row * K);
    }
  };
  float acc = 0.f;
  for (int s = 0; s < STAGES - 1; ++s) { load(s, s); cp_commit(); }
  for (int c = 0; c < nchunks; ++c) {
    if (c + STAGES - 1 < nchunks) load(c + STAGES - 1, (c + STAGES - 1) % STAGES);
    cp_commit();
    cp_wait<STAGES - 1>();
    __syncthreads();
    const float* Ws = sm + (c % STAGES) * NCH * LW;
#pragma unroll 4
    for (int w = 0; w < WORK; ++w) acc += Ws[((tid + w * 37) % NCH) * LW + (w & 31)];
    if (TWO_BARRIERS) __syncthreads();
  }
  out[blockIdx.x * NT + tid] = acc;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 20; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 20;
}

template <int NT, int STAGES, bool TB, int WORK>
void run(const float* W, float* out, int nchunks, const char* name) {
  auto k = stage_kernel<NT, STAGES, TB, WORK>;
  size_t smem = (size_t)STAGES * NCH * LW * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float ms = time_it([&] { k<<<128, NT, smem>>>(W, 128, nchunks, out); });
  printf("%-44s %7.1f us total  %6.0f ns/chunk  (%.1f GB/s per SM)\n", name, ms * 1e3,
         ms * 1e6 / nchunks, 32768.0 * nchunks / (ms * 1e-3) / 1e9);
}

int main() {
  float *W, *out;
  cudaMalloc(&W, 256 * 128 * 4 * 4);
  cudaMemset(W, 0, 256 * 128 * 4 * 4);
  cudaMalloc(&out, 148 * 1024 * 4);
  const int n = 48;
  run<512, 2, true, 0>(W, out, n, "512 thr, 2 stages, 2 barriers, no work");
  run<512, 2, true, 64>(W, out, n, "512 thr, 2 stages, 2 barriers, 64 LDS work");
  run<512, 3, false, 0>(W, out, n, "512 thr, 3 stages, 1 barrier, no work");
  run<512, 4, false, 0>(W, out, n, "512 thr, 4 stages, 1 barrier, no work");
  run<512, 4, false, 64>(W, out, n, "512 thr, 4 stages, 1 barrier, 64 LDS work");
  run<256, 2, true, 0>(W, out, n, "256 thr, 2 stages, 2 barriers, no work");
  run<256, 4, false, 0>(W, out, n, "256 thr, 4 stages, 1 barrier, no work");
  run<512, 2, true, 0>(W, out, 480, "512 thr, 2 stages, 2 barriers, 480 chunks");
  return 0;
}
