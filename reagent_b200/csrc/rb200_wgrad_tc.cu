// reagent_b200 -- weight gradients on the 5th-generation tensor cores.
//
//   dW_l[n, k] = sum_b dZ_l[b, n] * A_{l-1}[b, k],   db_l[n] = sum_b dZ_l[b, n]
// (autograd's Linear backward reached from loss.backward() in the reference's Lightning loop,
// reagent/training/reagent_lightning_module.py:108-133) as tcgen05.mma kind::tf32 with 3xTF32
// error compensation and the accumulator in Tensor Memory.
//
// The contraction runs over the BATCH, and both factors are stored batch-row-major in HBM
// ([B, N] and [B, K]): read as UMMA operands they are "MN-major" -- 4 consecutive features are
// the contiguous 16 bytes.  The kernel therefore stages a 32-row chunk of both matrices with
// 16-byte cp.async pieces straight into the canonical MN-major no-swizzle layout
//     [feature / 4][batch row][4 floats]      (feature-quad stride = SBO, 8-row group = LBO)
// -- a pure address scatter, no transposition of values -- splits it in place into TF32 hi / lo
// planes and issues, per 8-row k step,  D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  with
// M = 128 output features of dZ_l and N <= 256 input features of A_{l-1}.
//
// One CTA = (layer, 128-feature tile of dZ_l, 256-feature tile of A_{l-1}, batch slab); the
// slabs are summed later by the Adam kernel in slab order (deterministic), exactly like the
// mma.sync kernel this one replaces for shapes that fit (rb200_optim.cu keeps that kernel for
// the rest).  Two smem stages: the copies + split of chunk c+1 overlap the MMAs of chunk c.
#include "rb200_umma.cuh"

namespace rb200 {

constexpr int kWtRows = 32;                       // batch rows per stage = 4 MMA k steps
constexpr int kWtM = 128;                         // dZ features per tile (UMMA M)
constexpr int kWtN = 256;                         // input features per tile (UMMA N, TMEM columns)
constexpr int kWtQuad = kWtRows * 16 + 16;        // bytes per feature quad (+16: conflict-free scatter)
constexpr int kWtPlaneA = (kWtM / 4) * kWtQuad;
constexpr int kWtPlaneB = (kWtN / 4) * kWtQuad;
constexpr int kWtStage = 2 * (kWtPlaneA + kWtPlaneB);  // A_hi, A_lo, B_hi, B_lo
constexpr int kWtThreads = 256;
constexpr int kWtSmem = 2 * kWtStage + 64;

struct WtLayer {
  const float* A;   // [B, K]
  const float* dZ;  // [B, N]
  int K, N;
  long long w_off, b_off;
  int tiles_m, tiles_k, job_start;
};
struct WtParams {
  int n_layers;
  WtLayer L[kMaxLayers];
  int B, rows_per_split;
  float* gpart;
  long long P;
};

// K-major/MN-major agnostic descriptor: (start, "leading" and "stride" byte offsets), version 1
__device__ __forceinline__ uint64_t wt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return umma_desc(saddr, lbo, sbo);
}
// kind::tf32, fp32 accumulate, A and B MN-major (bits 15 / 16), M x N
__device__ __forceinline__ uint32_t wt_idesc(int M, int N) {
  return umma_idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(kWtThreads, 1) wgrad_tc_kernel(const WtParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int li = 0;
  while (li + 1 < p.n_layers && (int)blockIdx.x >= p.L[li + 1].job_start) ++li;
  const WtLayer& Ly = p.L[li];
  const int job = blockIdx.x - Ly.job_start;
  const int tm = job / Ly.tiles_k, tk = job - tm * Ly.tiles_k;
  const int n0 = tm * kWtM, k0 = tk * kWtN;
  const int N = Ly.N, K = Ly.K;
  const int nrows_m = min(kWtM, N - n0);             // valid dZ features of this tile
  const int ncols = min(kWtN, K - k0);               // valid input features of this tile
  const int nq_a = ceil_div(nrows_m, 4), nq_b = ceil_div(ncols, 4);
  const int n_mma = (ncols + 15) & ~15;              // UMMA N (multiple of 16 for M = 128)
  const int split = blockIdx.y;
  const int b_begin = split * p.rows_per_split;
  const int b_end = min(p.B, b_begin + p.rows_per_split);
  const int nchunks = ceil_div(max(b_end - b_begin, 0), kWtRows);

  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + 2 * kWtStage);  // [2] stage reusable
  uint64_t* acc_done = mma_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  // operand padding (features past the matrix, rows past the slab) must be finite zeros
  for (int i = tid * 16; i < 2 * kWtStage; i += kWtThreads * 16)
    *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(mma_done, 1);
    mbar_init(mma_done + 1, 1);
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)), "n"(kWtN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool vz = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ly.dZ) & 15) == 0);
  const bool va = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ly.A) & 15) == 0);
  // raw fp32 chunk -> hi planes of stage `st` (16-byte pieces, MN-major scatter)
  auto stage_load = [&](int c, int st) {
    unsigned char* base = smem + st * kWtStage;
    const int r0 = b_begin + c * kWtRows;
    for (int idx = tid; idx < kWtRows * (nq_a + nq_b); idx += kWtThreads) {
      const bool isb = idx >= kWtRows * nq_a;
      const int j = isb ? idx - kWtRows * nq_a : idx;
      const int nq = isb ? nq_b : nq_a;
      const int r = j / nq, q = j - r * nq;
      const int row = r0 + r;
      const int f = (isb ? k0 : n0) + 4 * q;          // first feature of the piece
      const int F = isb ? K : N;
      const float* src = (isb ? Ly.A : Ly.dZ) + (size_t)row * F + f;
      float* dst = reinterpret_cast<float*>(base + (isb ? 2 * kWtPlaneA : 0) + q * kWtQuad + r * 16);
      if (row < b_end && (isb ? va : vz) && f + 3 < F) {
        cp_async16(dst, src);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < b_end) {
          if (f < F) v.x = src[0];
          if (f + 1 < F) v.y = src[1];
          if (f + 2 < F) v.z = src[2];
          if (f + 3 < F) v.w = src[3];
        }
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
  };
  // hi planes -> (hi, lo) in place; also the bias-gradient partial of this thread's feature
  float bsum = 0.f;
  auto stage_split = [&](int st) {
    unsigned char* base = smem + st * kWtStage;
    if (tk == 0 && tid < kWtM) {  // raw values are still intact here
      const float* col = reinterpret_cast<const float*>(base + (tid >> 2) * kWtQuad) + (tid & 3);
#pragma unroll 8
      for (int r = 0; r < kWtRows; ++r) bsum += col[r * 4];
    }
    __syncthreads();
    for (int idx = tid; idx < kWtRows * (nq_a + nq_b); idx += kWtThreads) {
      const bool isb = idx >= kWtRows * nq_a;
      const int j = isb ? idx - kWtRows * nq_a : idx;
      const int q = j / kWtRows, r = j - q * kWtRows;
      float* hi = reinterpret_cast<float*>(base + (isb ? 2 * kWtPlaneA : 0) + q * kWtQuad + r * 16);
      float* lo = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(hi) + (isb ? kWtPlaneB : kWtPlaneA));
      float4 h, l;
      split4(*reinterpret_cast<const float4*>(hi), h, l);
      *reinterpret_cast<float4*>(hi) = h;
      *reinterpret_cast<float4*>(lo) = l;
    }
  };

  const uint32_t idesc = wt_idesc(kWtM, n_mma);
  uint32_t done_par[2] = {0u, 0u};
  if (nchunks > 0) {
    stage_load(0, 0);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    cp_async_wait<0>();
    __syncthreads();
    stage_split(st);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + st * kWtStage);
      const uint32_t a_hi = sb, a_lo = sb + kWtPlaneA, b_hi = sb + 2 * kWtPlaneA, b_lo = b_hi + kWtPlaneB;
#pragma unroll
      for (int ks = 0; ks < kWtRows / 8; ++ks) {
        const uint32_t o = ks * 128;  // 8 batch rows x 16 B inside every feature quad
        // MN-major no-swizzle: "leading" offset = next group of 8 k (128 B), "stride" offset =
        // next group of 4 features (the quad stride)
        const uint64_t dah = wt_desc(a_hi + o, 128, kWtQuad), dal = wt_desc(a_lo + o, 128, kWtQuad);
        const uint64_t dbh = wt_desc(b_hi + o, 128, kWtQuad), dbl = wt_desc(b_lo + o, 128, kWtQuad);
        umma_tf32(tmem, dah, dbh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(tmem, dal, dbh, idesc, 1u);
        umma_tf32(tmem, dah, dbl, idesc, 1u);
      }
      umma_commit(mma_done + st);
      if (c == nchunks - 1) umma_commit(acc_done);
    }
    // the next chunk goes into the other stage once ITS previous MMAs (chunk c-1) retired
    if (c + 1 < nchunks) {
      if (c >= 1) {
        mbar_wait(mma_done + (st ^ 1), done_par[st ^ 1]);
        done_par[st ^ 1] ^= 1u;
      }
      stage_load(c + 1, st ^ 1);
      cp_async_commit();
    }
  }

  // ---- epilogue: accumulator -> this slab's gradient partial ----
  float* gp = p.gpart + (size_t)split * p.P;
  if (nchunks > 0) {
    mbar_wait(acc_done, 0);
    tc_fence_after();
  }
  if (warp < 4) {
    const int n = n0 + warp * 32 + lane;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t v[16];
      if (nchunks > 0) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
              "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (n < N) {
        float* dst = gp + Ly.w_off + (size_t)n * K + k0 + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < ncols) dst[j] = __uint_as_float(v[j]);
      }
    }
    if (tk == 0 && n < N) gp[Ly.b_off + n] = bsum;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kWtN));
  }
}

}  // namespace rb200

using namespace rb200;

// Launch the tcgen05 weight-gradient kernel.  Returns RB200_E_SMEM when the shapes do not fit
// (the caller then uses the mma.sync kernel); every layer of the in-scope networks does.
int rb200_wgrad_tc_launch(const rb200_mlp_t* net, const float* net_input, int32_t batch,
                          const rb200_net_ws_t* ws, float* gpart, int32_t splits, void* stream) {
  WtParams p = {};
  p.n_layers = net->n_layers;
  p.B = batch;
  p.rows_per_split = ceil_div(ceil_div(batch, splits), kWtRows) * kWtRows;
  p.gpart = gpart;
  p.P = net->n_params;
  int jobs = 0;
  for (int l = 0; l < net->n_layers; ++l) {
    WtLayer& L = p.L[l];
    L.A = (l == 0) ? (net_input ? net_input : ws->input) : ws->hidden[l - 1];
    L.dZ = ws->dz[l];
    if (!L.A || !L.dZ) { set_last_error("rb200_mlp_wgrad: missing activation / dz for layer %d", l); return RB200_E_INVALID; }
    L.K = net->dims[l];
    L.N = net->dims[l + 1];
    L.w_off = net->w_off[l];
    L.b_off = net->b_off[l];
    L.tiles_m = ceil_div(L.N, kWtM);
    L.tiles_k = ceil_div(L.K, kWtN);
    L.job_start = jobs;
    jobs += L.tiles_m * L.tiles_k;
  }
  for (int l = net->n_layers; l < kMaxLayers; ++l) p.L[l].job_start = 1 << 30;
  static SmemOptIn optin = {};
  {
    cudaError_t e = ensure_dynamic_smem(wgrad_tc_kernel, optin, (size_t)kWtSmem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(wgrad_tc)");
  }
  dim3 grid(jobs, splits);
  wgrad_tc_kernel<<<grid, kWtThreads, kWtSmem, (cudaStream_t)stream>>>(p);
  return check_cuda(cudaGetLastError(), "wgrad_tc_kernel launch");
}
