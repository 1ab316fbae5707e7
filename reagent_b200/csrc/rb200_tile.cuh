// reagent_b200 -- row-tile MLP primitives (fp32 CUDA-core path).
//
// A CTA of NT (256 or 512) threads owns a tile of R = (NT/64)*TM batch rows and walks whole MLPs
// over it with every activation resident in shared memory; only weights stream
// through (cp.async, double buffered, out of L2 where all CTAs share them).
//
//   tile_linear_fwd : C[R,N]  = act(A[R,K] . W[N,K]^T + b)        (nn.Linear forward)
//   tile_linear_bwd : dA[R,K] = dZ[R,N] . W[N,K]                  (input gradient)
//
// Thread mapping: 2-D warp tiling, see "Warp tiling" below (every LDS.128 is a single
// shared-memory wavefront, 8 wavefronts per 64 FMA instructions per warp).
#pragma once
#include "rb200_common.cuh"

namespace rb200 {

constexpr int kNC = 256;  // output-column chunk processed per pass

template <int KC>
__host__ __device__ constexpr int wstage_floats() {
  // one stage must hold either the fwd chunk [256][KC+4] or the bwd chunk [KC][256+4]
  return (kNC * (KC + 4) > KC * (kNC + 4)) ? kNC * (KC + 4) : KC * (kNC + 4);
}

// ---------------------------------------------------------------------------
// Warp tiling.  The NW = NT/32 warps of the CTA are arranged WR x WC over the
// (R rows) x (256-column chunk) output; a warp owns a (4*TMe rows) x 32 columns block and
// its lanes are laid out 4 (row lanes, lr) x 8 (column lanes, lc):
//   thread rows  r = warp_row0 + lr + 4*i        (i < TMe)
//   thread cols  fwd: c = warp_col0 + lc + 8*j   (j < 4)      bwd: c = warp_col0 + 4*lc + (0..3)
// Per 4-deep k step a warp touches only 4*TMe row quads + 32 column quads of shared memory:
// every LDS.128 is ONE wavefront (column lanes read 8 distinct 16-byte quads that the 4 row
// lanes share by broadcast, and vice versa; the +4 float row padding keeps the quads of
// different rows in different banks).  8 wavefronts per 64 FMA-instructions per warp keeps
// the loop FMA-pipe bound (a 1-D lane mapping needs 20 and is shared-memory bound).
// WC adapts to the chunk width (8 / 4 / 2 column warps) so narrow layers waste no FMAs.
// ---------------------------------------------------------------------------
// Fragments are double buffered in registers: the LDS.128 of step s+1 are issued before the
// FMAs of step s, so the ~30-cycle shared-memory latency is covered by 16*TMe FMAs of the
// same warp (ptxas otherwise schedules each load right before its first use).
template <int TMe>
struct FragF {
  float4 w[4];
  float4 a[TMe];
};

template <int TMe>
__device__ __forceinline__ void fwd_load(FragF<TMe>& f, const float* __restrict__ arow, int lda4,
                                         const float* __restrict__ wrow, int lw8, int kk) {
#pragma unroll
  for (int j = 0; j < 4; ++j) f.w[j] = *reinterpret_cast<const float4*>(wrow + j * lw8 + kk);
#pragma unroll
  for (int i = 0; i < TMe; ++i) f.a[i] = *reinterpret_cast<const float4*>(arow + i * lda4 + kk);
}

template <int TMe>
__device__ __forceinline__ void fwd_fma(float (&acc)[4][4], const FragF<TMe>& f) {
  // k-outer order: 4*TMe independent FMAs between two uses of the same accumulator
#pragma unroll
  for (int i = 0; i < TMe; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(f.a[i].x, f.w[j].x, acc[i][j]);
#pragma unroll
  for (int i = 0; i < TMe; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(f.a[i].y, f.w[j].y, acc[i][j]);
#pragma unroll
  for (int i = 0; i < TMe; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(f.a[i].z, f.w[j].z, acc[i][j]);
#pragma unroll
  for (int i = 0; i < TMe; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(f.a[i].w, f.w[j].w, acc[i][j]);
}

template <int TMe>
__device__ __forceinline__ void fwd_inner(float (&acc)[4][4], const float* __restrict__ arow,
                                          int lda4, const float* __restrict__ wrow, int lw8,
                                          int klen) {
  const int steps = klen >> 2;
  FragF<TMe> f0, f1;
  fwd_load<TMe>(f0, arow, lda4, wrow, lw8, 0);
  int s = 0;
#pragma unroll 1
  for (; s + 1 < steps; s += 2) {
    fwd_load<TMe>(f1, arow, lda4, wrow, lw8, (s + 1) * 4);
    fwd_fma<TMe>(acc, f0);
    if (s + 2 < steps) fwd_load<TMe>(f0, arow, lda4, wrow, lw8, (s + 2) * 4);
    fwd_fma<TMe>(acc, f1);
  }
  if (s < steps) fwd_fma<TMe>(acc, f0);
}

template <int TMe>
struct FragB {
  float4 w[4];
  float4 z[TMe];
};

template <int TMe>
__device__ __forceinline__ void bwd_load(FragB<TMe>& f, const float* __restrict__ zrow, int ldz4,
                                         const float* __restrict__ wcol, int LW, int nn) {
#pragma unroll
  for (int j = 0; j < 4; ++j) f.w[j] = *reinterpret_cast<const float4*>(wcol + (nn + j) * LW);
#pragma unroll
  for (int i = 0; i < TMe; ++i) f.z[i] = *reinterpret_cast<const float4*>(zrow + i * ldz4 + nn);
}

template <int TMe>
__device__ __forceinline__ void bwd_fma(float (&acc)[4][4], const FragB<TMe>& f) {
#pragma unroll
  for (int i = 0; i < TMe; ++i) {
    acc[i][0] = fmaf(f.z[i].x, f.w[0].x, acc[i][0]);
    acc[i][1] = fmaf(f.z[i].x, f.w[0].y, acc[i][1]);
    acc[i][2] = fmaf(f.z[i].x, f.w[0].z, acc[i][2]);
    acc[i][3] = fmaf(f.z[i].x, f.w[0].w, acc[i][3]);
  }
#pragma unroll
  for (int i = 0; i < TMe; ++i) {
    acc[i][0] = fmaf(f.z[i].y, f.w[1].x, acc[i][0]);
    acc[i][1] = fmaf(f.z[i].y, f.w[1].y, acc[i][1]);
    acc[i][2] = fmaf(f.z[i].y, f.w[1].z, acc[i][2]);
    acc[i][3] = fmaf(f.z[i].y, f.w[1].w, acc[i][3]);
  }
#pragma unroll
  for (int i = 0; i < TMe; ++i) {
    acc[i][0] = fmaf(f.z[i].z, f.w[2].x, acc[i][0]);
    acc[i][1] = fmaf(f.z[i].z, f.w[2].y, acc[i][1]);
    acc[i][2] = fmaf(f.z[i].z, f.w[2].z, acc[i][2]);
    acc[i][3] = fmaf(f.z[i].z, f.w[2].w, acc[i][3]);
  }
#pragma unroll
  for (int i = 0; i < TMe; ++i) {
    acc[i][0] = fmaf(f.z[i].w, f.w[3].x, acc[i][0]);
    acc[i][1] = fmaf(f.z[i].w, f.w[3].y, acc[i][1]);
    acc[i][2] = fmaf(f.z[i].w, f.w[3].z, acc[i][2]);
    acc[i][3] = fmaf(f.z[i].w, f.w[3].w, acc[i][3]);
  }
}

template <int TMe>
__device__ __forceinline__ void bwd_inner(float (&acc)[4][4], const float* __restrict__ zrow,
                                          int ldz4, const float* __restrict__ wcol, int LW,
                                          int nlen) {
  const int steps = nlen >> 2;
  FragB<TMe> f0, f1;
  bwd_load<TMe>(f0, zrow, ldz4, wcol, LW, 0);
  int s = 0;
#pragma unroll 1
  for (; s + 1 < steps; s += 2) {
    bwd_load<TMe>(f1, zrow, ldz4, wcol, LW, (s + 1) * 4);
    bwd_fma<TMe>(acc, f0);
    if (s + 2 < steps) bwd_load<TMe>(f0, zrow, ldz4, wcol, LW, (s + 2) * 4);
    bwd_fma<TMe>(acc, f1);
  }
  if (s < steps) bwd_fma<TMe>(acc, f0);
}

// column-warp count for a chunk that is `cols` wide
__device__ __forceinline__ int pick_wc(int cols) { return cols > 128 ? 8 : (cols > 64 ? 4 : 2); }

// ---------------------------------------------------------------------------
// forward:  Cs[r, 0..N) = act(As[r, 0..K) . Wg[n, 0..K) + bg[n])
//   As : smem, row stride lda (multiple of 4), columns K..round_up4(K)-1 MUST be 0
//   Cs : smem, row stride ldc (multiple of 4); columns N..round_up4(N)-1 are zeroed
//   Wst: smem staging, 2 * wstage_floats<KC>() floats, 16B aligned; W chunk staged as
//        Ws[256][KC+4] (K contiguous)
// All NT threads must call (contains __syncthreads).
// ---------------------------------------------------------------------------
template <int NT, int TM, int KC>
__device__ __noinline__ void tile_linear_fwd(const float* __restrict__ As, int lda, int K,
                                             const float* __restrict__ Wg, int ldw,
                                             const float* __restrict__ bg, int N, int act,
                                             float* __restrict__ Cs, int ldc,
                                             float* __restrict__ Wst) {
  constexpr int LW = KC + 4;
  constexpr int STAGE = wstage_floats<KC>();
  constexpr int QPR = KC / 4;  // 16B quads per staged row
  constexpr int NW = NT / 32;
  constexpr int R = (NT / 64) * TM;
  static_assert(TM == 4, "warp tiling assumes 4 rows per thread at full chunk width");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 3, lc = lane & 7;
  const int nk = ceil_div(K, KC), nn = ceil_div(N, kNC), total = nk * nn;
  const bool vec = ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0);

  // staging map: thread -> (quad lq of the k-chunk, rows lr0, lr0+RPI, ...): fixed per thread
  constexpr int RPI = NT / QPR;
  const int lq = tid % QPR, lr0 = tid / QPR;
  auto load_chunk = [&](int c, int stage) {
    const int nci = c / nk, kci = c - nci * nk;
    const int n0 = nci * kNC, k0 = kci * KC;
    float* dst = Wst + stage * STAGE;
    const int rows = min(kNC, N - n0);
    const int k = k0 + 4 * lq;
    const bool fast = vec && (k + 3 < K);
#pragma unroll
    for (int it = 0; it < kNC / RPI; ++it) {
      const int row = lr0 + it * RPI;
      if (row < rows) {
        float* d = dst + row * LW + 4 * lq;
        const float* src = Wg + (size_t)(n0 + row) * ldw;
        if (fast) {
          cp_async16(d, src + k);
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < K) v.x = src[k];
          if (k + 1 < K) v.y = src[k + 1];
          if (k + 2 < K) v.z = src[k + 2];
          if (k + 3 < K) v.w = src[k + 3];
          *reinterpret_cast<float4*>(d) = v;
        }
      }
    }
  };

  float acc[4][4];
  load_chunk(0, 0);
  cp_async_commit();
  for (int c = 0; c < total; ++c) {
    const int nci = c / nk, kci = c - nci * nk;
    const int n0 = nci * kNC, k0 = kci * KC;
    const int ncols = min(kNC, N - n0);
    const int WC = pick_wc(ncols), WR = NW / WC;
    const int wr = warp / WC, wc = warp - wr * WC;
    const int rpw = R / WR;              // rows per warp = 4 * TMe
    const int row0 = wr * rpw + lr;      // thread rows row0 + 4*i
    const int col0 = wc * 32 + lc;       // thread cols (chunk relative) col0 + 8*j
    if (kci == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    if (c + 1 < total) {
      load_chunk(c + 1, (c + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Ws = Wst + (c & 1) * STAGE;
    if (wc * 32 < ncols) {
      const int klen = min(KC, round_up4(K - k0));
      const float* arow = As + row0 * lda + k0;
      const float* wrow = Ws + col0 * LW;
      if (rpw == 16) fwd_inner<4>(acc, arow, 4 * lda, wrow, 8 * LW, klen);
      else if (rpw == 8) fwd_inner<2>(acc, arow, 4 * lda, wrow, 8 * LW, klen);
      else fwd_inner<1>(acc, arow, 4 * lda, wrow, 8 * LW, klen);
    }
    if (kci == nk - 1 && wc * 32 < round_up4(ncols)) {
      const int n4 = round_up4(N);
      const int tme = rpw >> 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + col0 + 8 * j;
        if (col < n4) {
          const bool real = col < N;
          const float b = (real && bg != nullptr) ? bg[col] : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < tme)
              Cs[(row0 + 4 * i) * ldc + col] = real ? act_fwd(acc[i][j] + b, act) : 0.f;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// backward (input gradient):  dAs[r, 0..Kout) = dZs[r, 0..N) . Wg[n, kcol0 .. kcol0+Kout)
//   then, if Hs != nullptr, multiplied elementwise by act'(Hs[r,k]) (Hs = the
//   activation OUTPUT that produced this input, same column indexing as dAs).
//   dZs: smem, stride ldz, columns N..round_up4(N)-1 MUST be 0.
//   dAs: smem, stride lda; columns Kout..round_up4(Kout)-1 are zeroed.
//   Wg points at W[0][kcol0]; ldw is the full row stride of W.  W chunk staged as
//   Ws[KC][256+4] (rows = contraction index n).
// ---------------------------------------------------------------------------
template <int NT, int TM, int KC>
__device__ __noinline__ void tile_linear_bwd(const float* __restrict__ dZs, int ldz, int N,
                                             const float* __restrict__ Wg, int ldw, int Kout,
                                             const float* __restrict__ Hs, int ldh, int hact,
                                             float* __restrict__ dAs, int lda,
                                             float* __restrict__ Wst) {
  constexpr int LW = kNC + 4;
  constexpr int STAGE = wstage_floats<KC>();
  constexpr int NR = KC;  // contraction rows per staged chunk
  constexpr int NW = NT / 32;
  constexpr int R = (NT / 64) * TM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 3, lc = lane & 7;
  const int nnc = ceil_div(N, NR), nkc = ceil_div(Kout, kNC), total = nnc * nkc;
  const bool vec = ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0);

  // staging map: thread -> (quad lq of the 256-wide row, rows lr0, lr0+RPI, ...)
  constexpr int QPRB = kNC / 4;
  constexpr int RPI = NT / QPRB;
  const int lq = tid % QPRB, lr0 = tid / QPRB;
  auto load_chunk = [&](int c, int stage) {
    const int kci = c / nnc, nci = c - kci * nnc;
    const int k0 = kci * kNC, n0 = nci * NR;
    float* dst = Wst + stage * STAGE;
    const int k = k0 + 4 * lq;
    if (k < round_up4(Kout)) {
      const bool fast = vec && (k + 3 < Kout);
#pragma unroll
      for (int it = 0; it < (NR + RPI - 1) / RPI; ++it) {
        const int row = lr0 + it * RPI;
        if (row < NR) {
          const int n = n0 + row;
          float* d = dst + row * LW + 4 * lq;
          if (n < N && fast) {
            cp_async16(d, Wg + (size_t)n * ldw + k);
          } else {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
              const float* src = Wg + (size_t)n * ldw;
              if (k < Kout) v.x = src[k];
              if (k + 1 < Kout) v.y = src[k + 1];
              if (k + 2 < Kout) v.z = src[k + 2];
              if (k + 3 < Kout) v.w = src[k + 3];
            }
            *reinterpret_cast<float4*>(d) = v;
          }
        }
      }
    }
  };

  float acc[4][4];
  load_chunk(0, 0);
  cp_async_commit();
  for (int c = 0; c < total; ++c) {
    const int kci = c / nnc, nci = c - kci * nnc;
    const int k0 = kci * kNC, n0 = nci * NR;
    const int kcols = min(kNC, round_up4(Kout - k0));
    const int WC = pick_wc(kcols), WR = NW / WC;
    const int wr = warp / WC, wc = warp - wr * WC;
    const int rpw = R / WR;
    const int row0 = wr * rpw + lr;
    const int ccol = wc * 32 + 4 * lc;   // chunk-relative first column of this thread
    if (nci == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    if (c + 1 < total) {
      load_chunk(c + 1, (c + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Ws = Wst + (c & 1) * STAGE;
    const bool active = ccol < kcols;
    if (active) {
      const int nlen = min(NR, round_up4(N - n0));
      const float* zrow = dZs + row0 * ldz + n0;
      const float* wcol = Ws + ccol;
      if (rpw == 16) bwd_inner<4>(acc, zrow, 4 * ldz, wcol, LW, nlen);
      else if (rpw == 8) bwd_inner<2>(acc, zrow, 4 * ldz, wcol, LW, nlen);
      else bwd_inner<1>(acc, zrow, 4 * ldz, wcol, LW, nlen);
    }
    if (nci == nnc - 1 && active) {
      const int tme = rpw >> 2;
      const int kcol = k0 + ccol;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < tme) {
          const int r = row0 + 4 * i;
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kcol + j;
            float g = 0.f;
            if (k < Kout) {
              g = acc[i][j];
              if (Hs != nullptr) g *= act_bwd_from_out(Hs[r * ldh + k], hact);
            }
            o[j] = g;
          }
          *reinterpret_cast<float4*>(dAs + r * lda + kcol) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// cooperative tile <-> global copies (all NT threads)
// ---------------------------------------------------------------------------
// smem[r, 0..round_up4(D)) <- g[(row0+r), 0..D), zero padded; rows >= nrows zeroed.
template <int NT, int R>
__device__ void tile_load_rows(float* __restrict__ s, int lds, const float* __restrict__ g,
                               int ldg, int D, int row0, int nrows) {
  const int d4 = round_up4(D) / 4;
  const bool vec = ((ldg & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  for (int idx = threadIdx.x; idx < R * d4; idx += NT) {
    const int r = idx / d4, q = idx - r * d4;
    const int c = 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < nrows) {
      const float* src = g + (size_t)(row0 + r) * ldg;
      if (vec && c + 3 < D) {
        v = *reinterpret_cast<const float4*>(src + c);
      } else {
        if (c < D) v.x = src[c];
        if (c + 1 < D) v.y = src[c + 1];
        if (c + 2 < D) v.z = src[c + 2];
        if (c + 3 < D) v.w = src[c + 3];
      }
    }
    *reinterpret_cast<float4*>(s + r * lds + c) = v;
  }
}

// g[(row0+r), 0..D) <- smem[r, 0..D) for rows < nrows
template <int NT, int R>
__device__ void tile_store_rows(const float* __restrict__ s, int lds, float* __restrict__ g,
                                int ldg, int D, int row0, int nrows) {
  const int d4 = round_up4(D) / 4;
  const bool vec = ((ldg & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  for (int idx = threadIdx.x; idx < R * d4; idx += NT) {
    const int r = idx / d4, q = idx - r * d4;
    const int c = 4 * q;
    if (row0 + r >= nrows) continue;
    const float4 v = *reinterpret_cast<const float4*>(s + r * lds + c);
    float* dst = g + (size_t)(row0 + r) * ldg;
    if (vec && c + 3 < D) {
      *reinterpret_cast<float4*>(dst + c) = v;
    } else {
      if (c < D) dst[c] = v.x;
      if (c + 1 < D) dst[c + 1] = v.y;
      if (c + 2 < D) dst[c + 2] = v.z;
      if (c + 3 < D) dst[c + 3] = v.w;
    }
  }
}

// ---------------------------------------------------------------------------
// Whole-MLP forward over the tile.  Input in `in` (stride ld_in).  Hidden
// activations ping-pong between hA and hB (stride ldh each); the final layer's
// output lands in `out` (stride ld_out).  If save != nullptr, save[l] (global,
// [B, dims[l+1]] dense) receives layer l's output for l < n_layers-1 (hidden
// only) -- the activations the weight-gradient kernel and the backward need.
// ---------------------------------------------------------------------------
template <int NT, int TM, int KC>
__device__ void tile_mlp_fwd(const Mlp& net, const float* in, int ld_in, float* hA, float* hB,
                             int ldh, float* out, int ld_out, float* Wst,
                             float* const* save, int row0, int nrows) {
  constexpr int R = (NT / 64) * TM;
  const float* cur = in;
  int ldc = ld_in;
  for (int l = 0; l < net.n_layers; ++l) {
    const bool last = (l == net.n_layers - 1);
    float* dst = last ? out : ((l & 1) ? hB : hA);
    const int ldd = last ? ld_out : ldh;
    tile_linear_fwd<NT, TM, KC>(cur, ldc, net.dims[l], net.params + net.w_off[l], net.dims[l],
                            net.params + net.b_off[l], net.dims[l + 1], net.act[l], dst, ldd,
                            Wst);
    if (!last && save != nullptr && save[l] != nullptr)
      tile_store_rows<NT, R>(dst, ldd, save[l], net.dims[l + 1], net.dims[l + 1], row0, nrows);
    cur = dst;
    ldc = ldd;
  }
}

// ---------------------------------------------------------------------------
// Whole-MLP backward (dZ chain) over the tile.
//   dz_last: smem [R, round_up4(dims[L])] = dLoss/d(pre-activation of last layer)
//   hidden[l] (global, dense [B, dims[l+1]]): saved outputs of layer l (l < L-1)
//   dz_out[l] (global, dense [B, dims[l+1]]): receives dLoss/d(pre-act of layer l)
//   On return, if din != nullptr it holds dLoss/d(input columns [in_col0, in_col0+in_cols))
//   in smem (stride ld_din).
// Buffers gA/gB (stride ldg) ping-pong the hidden dZ tiles; hbuf (stride ldg) is
// scratch for re-loading saved activations.
// ---------------------------------------------------------------------------
template <int NT, int TM, int KC>
__device__ void tile_mlp_bwd(const Mlp& net, float* dz_last, int ld_last, float* gA, float* gB,
                             float* hbuf, int ldg, float* Wst, const float* const* hidden,
                             float* const* dz_out, int row0, int nrows, float* din, int ld_din,
                             int in_col0, int in_cols) {
  constexpr int R = (NT / 64) * TM;
  const int L = net.n_layers;
  float* cur = dz_last;
  int ldc = ld_last;
  if (dz_out != nullptr && dz_out[L - 1] != nullptr)
    tile_store_rows<NT, R>(cur, ldc, dz_out[L - 1], net.dims[L], net.dims[L], row0, nrows);
  for (int l = L - 1; l >= 1; --l) {
    // dZ_{l-1} = (dZ_l . W_l) * act'_{l-1}(H_{l-1})
    tile_load_rows<NT, R>(hbuf, ldg, hidden[l - 1], net.dims[l], net.dims[l], row0, nrows);
    __syncthreads();
    float* dst = (l & 1) ? gA : gB;
    tile_linear_bwd<NT, TM, KC>(cur, ldc, net.dims[l + 1], net.params + net.w_off[l], net.dims[l],
                            net.dims[l], hbuf, ldg, net.act[l - 1], dst, ldg, Wst);
    if (dz_out != nullptr && dz_out[l - 1] != nullptr)
      tile_store_rows<NT, R>(dst, ldg, dz_out[l - 1], net.dims[l], net.dims[l], row0, nrows);
    cur = dst;
    ldc = ldg;
  }
  if (din != nullptr) {
    tile_linear_bwd<NT, TM, KC>(cur, ldc, net.dims[1], net.params + net.w_off[0] + in_col0,
                            net.dims[0], in_cols, nullptr, 0, 0, din, ld_din, Wst);
  }
}

}  // namespace rb200
