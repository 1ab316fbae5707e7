// Micro-benchmark: what does one tcgen05.mma.kind::tf32 (M = 128, K = 8) cost when a single
// thread issues a long run of them?  Varies N, the A operand source (shared memory descriptor /
// tensor memory), and how many independent accumulators the run round-robins over, and also
// times the two synchronisation patterns of the TD kernel's layer hand-over (256 per-thread
// mbarrier arrivals vs one per warp).  Timing only: operand contents are zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I reagent_b200/csrc -o profiles/micro/mma_cost profiles/micro/mma_cost.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rb200_umma.cuh"

using namespace rb200;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db),
      "r"(idesc), "r"(acc)
      : "memory");
}

struct Cfg { int reps; };

// everything the issue loop needs is a compile-time constant or a loop-invariant, so that the
// operands stay in uniform registers (as in the TD kernel's issue loop: back-to-back UTCHMMA)
template <int N, int NACC, bool ATMEM, int PAIR>
__global__ void __launch_bounds__(256, 1) mma_cost_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < 96 * 1024; i += 256 * 16) *reinterpret_cast<float4*>(sm + i) = make_float4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const bool leader = elect_one();
    if (leader) {
      constexpr uint32_t lboB = (uint32_t)(N * 16 + 16), lboA = 128 * 16 + 16;
      const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
      const uint64_t db = hi | ((smem_u32(sm) >> 4) & 0x3fffu) | ((uint64_t)(lboB >> 4) << 16);
      const uint64_t da = hi | ((smem_u32(sm + 32768) >> 4) & 0x3fffu) | ((uint64_t)(lboA >> 4) << 16);
      const uint32_t idesc = umma_idesc_tf32(128, N);
      const uint32_t idesc2 = umma_idesc_tf32(128, PAIR == 2 ? N : (N / 2 < 8 ? 8 : N / 2));
      uint32_t par = 0;
      for (int round = 0; round < 3; ++round) {  // last round is the one reported
        const long long t0 = clock64();
        for (int i = 0; i < c.reps; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t d = tmem + (uint32_t)((j % NACC) * N);
            if (ATMEM) umma_ts(d, tmem + 384 + j * 8, db, idesc, 1u);
            else umma_tf32(d, da, db, idesc, 1u);
            if (PAIR) {  // the TD kernel's second MMA of a k step (PAIR == 1: half N), same accumulator
              if (ATMEM) umma_ts(d, tmem + 448 + j * 8, db, idesc2, 1u);
              else umma_tf32(d, da, db, idesc2, 1u);
            }
          }
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, par);
        par ^= 1u;
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
        // one MMA + commit + wait: the fixed part of a hand-over
        const long long t3 = clock64();
        if (ATMEM) umma_ts(tmem, tmem + 384, db, idesc, 1u); else umma_tf32(tmem, da, db, idesc, 1u);
        umma_commit(&bar);
        mbar_wait(&bar, par);
        par ^= 1u;
        out[2] = clock64() - t3;
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int N, int NACC, bool ATMEM, int PAIR>
void run(long long* out) {
  if (NACC * N > 256) return;
  auto k = mma_cost_kernel<N, NACC, ATMEM, PAIR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  Cfg c = {512};
  k<<<1, 256, 96 * 1024>>>(c, out);
  long long h[3];
  cudaError_t e = cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  const int per = PAIR ? 2 : 1;
  printf("%3d %s %d %d | %7.1f %7.1f | %lld\n", N, ATMEM ? "tmem" : "smem", NACC, PAIR,
         (double)h[0] / (c.reps * per), (double)h[1] / (c.reps * per), h[2]);
}
template <int N, bool ATMEM>
void run_n(long long* out) {
  run<N, 1, ATMEM, 0>(out); run<N, 2, ATMEM, 0>(out); run<N, 4, ATMEM, 0>(out);
  run<N, 1, ATMEM, 1>(out); run<N, 2, ATMEM, 1>(out);
  run<N, 1, ATMEM, 2>(out);
}

// non-blocking poll (mbarrier.test_wait) instead of the potentially-suspending try_wait
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}

// hand-over synchronisation: 8 producer warps -> 1 consumer warp and back, `iters` times
template <int kPerWarp, bool kSpin = false>
__global__ void __launch_bounds__(288, 1) handover_kernel(int iters, long long* out) {
  __shared__ uint64_t ready, back;
  __shared__ float buf[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&ready, kPerWarp ? 8 : 256);
    mbar_init(&back, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 8) {
    for (int i = 0; i < iters; ++i) {
      buf[tid] = (float)i;
      fence_proxy_async_smem();
      if (kPerWarp) { __syncwarp(); if (lane == 0) mbar_arrive(&ready); }
      else mbar_arrive(&ready);
      if (kSpin) mbar_spin(&back, (uint32_t)i & 1u); else mbar_wait(&back, (uint32_t)i & 1u);
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      if (kSpin) mbar_spin(&ready, (uint32_t)i & 1u); else mbar_wait(&ready, (uint32_t)i & 1u);
      if (lane == 0) mbar_arrive(&back);
      __syncwarp();
    }
  }
  if (tid == 0) out[0] = clock64() - t0;
}

// The TD kernel's MMA pattern (per k step: N = 64 then N = 32 into the same accumulator, A from
// tensor memory) while other warps of the CTA generate the traffic the real kernel has:
//   bit 0  four warps storing 2 x 32 columns to tensor memory per iteration (tcgen05.st)
//   bit 1  four warps reading 16-byte rows from shared memory (the loaders' LDS.128)
//   bit 2  one warp streaming 16.5 KB bulk copies global -> shared
//   bit 3  eight warps loading accumulator columns (tcgen05.ld) and storing to shared memory
__global__ void __launch_bounds__(576, 1) contention_kernel(int mode, int reps, const float* gsrc, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar, cbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid * 16; i < 160 * 1024; i += 576 * 16) *reinterpret_cast<float4*>(sm + i) = make_float4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  float sink = 0.f;
  if (warp == 16) {  // MMA issuer
    if (elect_one()) {
      const uint32_t lboB = 64 * 16 + 16;
      const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
      const uint64_t db = hi | ((smem_u32(sm) >> 4) & 0x3fffu) | ((uint64_t)(lboB >> 4) << 16);
      const uint32_t i64 = umma_idesc_tf32(128, 64), i32 = umma_idesc_tf32(128, 32);
      uint32_t par = 0;
      for (int round = 0; round < 2; ++round) {
        const long long t0 = clock64();
        for (int i = 0; i < reps; i += 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_ts(tmem, tmem + 256 + j * 8, db + (uint64_t)(j * ((2 * lboB) >> 4)), i64, 1u);
            umma_ts(tmem, tmem + 288 + j * 8, db + (uint64_t)(j * ((2 * lboB) >> 4)), i32, 1u);
          }
        }
        umma_commit(&bar);
        mbar_wait(&bar, par);
        par ^= 1u;
        out[0] = clock64() - t0;
      }
      stop = 1;
    }
    __syncwarp();
  } else if (warp >= 8 && warp < 12 && (mode & 1)) {  // tcgen05.st traffic (loader warps' quadrants)
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (float)(i + lane);
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    while (!stop) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
          "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(ta),
          "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
          "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
          "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
  } else if (warp >= 12 && warp < 16 && (mode & 2)) {  // LDS.128 traffic
    const unsigned char* base = sm + 32768 + lane * 16;
    while (!stop) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(base + q * 2064 + (warp & 3) * 512)));
        sink += v.x + v.y + v.z + v.w;
      }
    }
  } else if (warp == 17 && (mode & 4)) {  // bulk copies
    const bool leader = elect_one();
    uint32_t par = 0;
    int stg = 0;
    while (!stop) {
      if (leader) {
        mbar_expect_tx(&cbar[stg], 16512);
        bulk_g2s(sm + 65536 + stg * 16512, gsrc + stg * 4128, 16512, &cbar[stg]);
      }
      if (++stg == 4) {
        for (int i = 0; i < 4; ++i) mbar_wait(&cbar[i], par);
        stg = 0;
        par ^= 1u;
      }
    }
    // drain
    if (stg) for (int i = 0; i < stg; ++i) mbar_wait(&cbar[i], par);
  } else if (warp < 8 && (mode & 8)) {  // epilogue-like: tcgen05.ld + shared-memory stores
    float* ob = reinterpret_cast<float*>(sm + 135168) + tid * 4;
    while (!stop) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15])
          : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 64));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) ob[(j & 3)] = __uint_as_float(v[j]);
    }
  }
  if (sink == 123.456f) out[7] = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <bool kSpin>
__global__ void __launch_bounds__(64, 1) pingpong_kernel(int iters, long long* out) {
  __shared__ uint64_t ab, ba;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&ab, 1); mbar_init(&ba, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (warp == 0) {
      if (lane == 0) mbar_arrive(&ab);
      __syncwarp();
      if (kSpin) mbar_spin(&ba, (uint32_t)i & 1u); else mbar_wait(&ba, (uint32_t)i & 1u);
    } else {
      if (kSpin) mbar_spin(&ab, (uint32_t)i & 1u); else mbar_wait(&ab, (uint32_t)i & 1u);
      if (lane == 0) mbar_arrive(&ba);
      __syncwarp();
    }
  }
  if (tid == 0) out[0] = clock64() - t0;
}

// one thread commits (nothing pending), the same warp waits: commit -> mbarrier latency
template <bool kSpin>
__global__ void __launch_bounds__(64, 1) commit_kernel(int iters, long long* out) {
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  if (warp == 0) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) umma_commit(&bar);
      __syncwarp();
      if (kSpin) mbar_spin(&bar, (uint32_t)i & 1u); else mbar_wait(&bar, (uint32_t)i & 1u);
    }
    if (tid == 0) out[0] = clock64() - t0;
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  printf("N a_src n_acc pair | issue cyc/MMA  total cyc/MMA | 1 MMA+commit+wait cycles  (pair 1: N then N/2; 2: N then N)\n");
  run_n<16, false>(out); run_n<32, false>(out); run_n<64, false>(out); run_n<128, false>(out); run_n<256, false>(out);
  run_n<16, true>(out); run_n<32, true>(out); run_n<64, true>(out); run_n<128, true>(out); run_n<256, true>(out);
  long long h;
  handover_kernel<0><<<1, 288>>>(1000, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("hand-over round trip, 256 per-thread arrivals: %.1f cycles\n", (double)h / 1000);
  handover_kernel<1><<<1, 288>>>(1000, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("hand-over round trip, 8 per-warp arrivals:     %.1f cycles\n", (double)h / 1000);
  handover_kernel<1, true><<<1, 288>>>(1000, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("hand-over round trip, per-warp arrivals, test_wait polling: %.1f cycles\n", (double)h / 1000);
  pingpong_kernel<false><<<1, 64>>>(1000, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("two warps ping-pong (1 arrival each way), try_wait:  %.1f cycles per round trip\n", (double)h / 1000);
  pingpong_kernel<true><<<1, 64>>>(1000, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("two warps ping-pong (1 arrival each way), test_wait: %.1f cycles per round trip\n", (double)h / 1000);
  commit_kernel<false><<<1, 64>>>(200, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("tcgen05.commit (no MMA pending) -> waiter, try_wait:  %.1f cycles\n", (double)h / 200);
  commit_kernel<true><<<1, 64>>>(200, out);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("tcgen05.commit (no MMA pending) -> waiter, test_wait: %.1f cycles\n", (double)h / 200);
  {
    float* g; cudaMalloc(&g, 4 * 16512); cudaMemset(g, 0, 4 * 16512);
    cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const char* names[] = {"none", "tmem st", "lds", "tmem st + lds", "bulk", "tmem st + bulk", "lds + bulk", "st + lds + bulk",
                           "epi", "epi + st", "epi + lds", "epi + st + lds", "epi + bulk", "epi + st + bulk", "epi + lds + bulk", "all"};
    for (int mode = 0; mode < 16; ++mode) {
      contention_kernel<<<1, 576, 160 * 1024>>>(mode, 2048, g, out);
      cudaError_t e = cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("contention mode %d error: %s\n", mode, cudaGetErrorString(e)); return 1; }
      printf("MMA pair (N=64 + N=32, A in tmem) under traffic [%s]: %.1f cycles per pair (48 = tensor pipe alone)\n", names[mode], (double)h / 2048);
    }
  }
  return 0;
}
