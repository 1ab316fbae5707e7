// Micro-benchmarks: FFMA issue rate and mma.sync TF32 m16n8k8 rate per SM on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x[4] = {a, a + 1.f, a + 2.f, a + 3.f};
  float y[4] = {b, b + 1.f, b + 2.f, b + 3.f};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(x[i], y[j], acc[i * 4 + j]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mma_tf32_kernel(float* out, int iters) {
  uint32_t a[4], b[2];
  float c[4][4];
  for (int i = 0; i < 4; ++i) a[i] = 0x3f800000u + threadIdx.x + i;
  b[0] = 0x3f000000u + threadIdx.x; b[1] = 0x3e800000u + threadIdx.x;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int t = 0; t < 4; ++t) for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mma_bf16_kernel(float* out, int iters) {
  uint32_t a[4], b[2];
  float c[4][4];
  for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + threadIdx.x + i;
  b[0] = 0x3f003f00u + threadIdx.x; b[1] = 0x3e803e80u + threadIdx.x;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int t = 0; t < 4; ++t) for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 8 * sizeof(float));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps * 32;
    float ms = time_it([&] { ffma_kernel<<<148, threads>>>(out, iters, 1.0001f, 0.9999f); });
    double fma = 148.0 * threads * iters * 64.0;
    printf("FFMA   warps/SM=%2d: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at %.0f MHz)\n", warps, ms,
           2 * fma / ms / 1e9, fma / 148 / (ms * 1e-3 * clk_khz * 1e3), clk_khz / 1e3);
    ms = time_it([&] { mma_tf32_kernel<<<148, threads>>>(out, iters); });
    double mac = 148.0 * warps * iters * 4.0 * 16 * 8 * 8;
    printf("MMA tf32 m16n8k8  warps/SM=%2d: %.3f ms  %.1f TFLOP/s  (%.0f MAC/clk/SM)\n", warps, ms,
           2 * mac / ms / 1e9, mac / 148 / (ms * 1e-3 * clk_khz * 1e3));
    ms = time_it([&] { mma_bf16_kernel<<<148, threads>>>(out, iters); });
    mac = 148.0 * warps * iters * 4.0 * 16 * 8 * 16;
    printf("MMA bf16 m16n8k16 warps/SM=%2d: %.3f ms  %.1f TFLOP/s  (%.0f MAC/clk/SM)\n", warps, ms,
           2 * mac / ms / 1e9, mac / 148 / (ms * 1e-3 * clk_khz * 1e3));
  }
  return 0;
}
