// reagent_b200 -- row-tile MLP primitives (fp32 CUDA-core path).
//
// A CTA of 256 threads owns a tile of R = 4*TM batch rows and walks whole MLPs
// over it with every activation resident in shared memory; only weights stream
// through (cp.async, double buffered, out of L2 where all CTAs share them).
//
//   tile_linear_fwd : C[R,N]  = act(A[R,K] . W[N,K]^T + b)        (nn.Linear forward)
//   tile_linear_bwd : dA[R,K] = dZ[R,N] . W[N,K]                  (input gradient)
//
// Thread mapping (both): ty = tid/64 owns rows ty*TM..ty*TM+TM-1; tx = tid%64.
//   fwd: tx owns output columns {tx, tx+64, tx+128, tx+192} of a 256-wide chunk;
//        W chunk staged as Ws[256][KC+4] (K contiguous): LDS.128 conflict-free
//        because consecutive lanes hit rows 4*(KC+4) bytes apart = distinct bank quads.
//   bwd: tx owns 4 consecutive output columns 4*tx..4*tx+3 of a 256-wide chunk;
//        W chunk staged as Ws[KC][256+4] (rows = contraction index n).
// Per 4-deep k step a thread issues TM+4 LDS.128 for TM*16 FMAs (TM=8: 12 vs 128),
// A-operand loads are warp-wide broadcasts, so the loop is FMA-pipe bound.
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
// forward:  Cs[r, 0..N) = act(As[r, 0..K) . Wg[n, 0..K) + bg[n])
//   As : smem, row stride lda (multiple of 4), columns K..round_up4(K)-1 MUST be 0
//   Cs : smem, row stride ldc (multiple of 4); columns N..round_up4(N)-1 are zeroed
//   Wst: smem staging, 2 * wstage_floats<KC>() floats, 16B aligned
// All 256 threads must call (contains __syncthreads).
// ---------------------------------------------------------------------------
template <int TM, int KC>
__device__ void tile_linear_fwd(const float* __restrict__ As, int lda, int K,
                                const float* __restrict__ Wg, int ldw,
                                const float* __restrict__ bg, int N, int act,
                                float* __restrict__ Cs, int ldc, float* __restrict__ Wst) {
  constexpr int LW = KC + 4;
  constexpr int STAGE = wstage_floats<KC>();
  constexpr int QPR = KC / 4;  // 16B quads per staged row
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const int nk = ceil_div(K, KC), nn = ceil_div(N, kNC), total = nk * nn;
  const bool vec = ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0);

  auto load_chunk = [&](int c, int stage) {
    const int nci = c / nk, kci = c - nci * nk;
    const int n0 = nci * kNC, k0 = kci * KC;
    float* dst = Wst + stage * STAGE;
    const int rows = min(kNC, N - n0);
    for (int seg = tid; seg < rows * QPR; seg += kThreads) {
      const int row = seg / QPR, q = seg - row * QPR;
      const int k = k0 + 4 * q;
      float* d = dst + row * LW + 4 * q;
      const float* src = Wg + (size_t)(n0 + row) * ldw;
      if (vec && k + 3 < K) {
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
  };

  float acc[TM][4];
  load_chunk(0, 0);
  cp_async_commit();
  for (int c = 0; c < total; ++c) {
    const int nci = c / nk, kci = c - nci * nk;
    const int n0 = nci * kNC, k0 = kci * KC;
    if (kci == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
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
    if (n0 + tx < N) {
      const int klen = min(KC, round_up4(K - k0));
      const float* arow = As + (ty * TM) * lda + k0;
#pragma unroll 2
      for (int kk = 0; kk < klen; kk += 4) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = *reinterpret_cast<const float4*>(Ws + (tx + 64 * j) * LW + kk);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(arow + i * lda + kk);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j] = fmaf(a.x, w[j].x, acc[i][j]);
            acc[i][j] = fmaf(a.y, w[j].y, acc[i][j]);
            acc[i][j] = fmaf(a.z, w[j].z, acc[i][j]);
            acc[i][j] = fmaf(a.w, w[j].w, acc[i][j]);
          }
        }
      }
    }
    if (kci == nk - 1) {
      const int n4 = round_up4(N);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx + 64 * j;
        if (col < n4) {
          const bool real = col < N;
          const float b = (real && bg != nullptr) ? bg[col] : 0.f;
#pragma unroll
          for (int i = 0; i < TM; ++i)
            Cs[(ty * TM + i) * ldc + col] = real ? act_fwd(acc[i][j] + b, act) : 0.f;
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
//   Wg points at W[0][kcol0]; ldw is the full row stride of W.
// ---------------------------------------------------------------------------
template <int TM, int KC>
__device__ void tile_linear_bwd(const float* __restrict__ dZs, int ldz, int N,
                                const float* __restrict__ Wg, int ldw, int Kout,
                                const float* __restrict__ Hs, int ldh, int hact,
                                float* __restrict__ dAs, int lda, float* __restrict__ Wst) {
  constexpr int LW = kNC + 4;
  constexpr int STAGE = wstage_floats<KC>();
  constexpr int NR = KC;  // contraction rows per staged chunk
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const int nnc = ceil_div(N, NR), nkc = ceil_div(Kout, kNC), total = nnc * nkc;
  const bool vec = ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0);

  auto load_chunk = [&](int c, int stage) {
    const int kci = c / nnc, nci = c - kci * nnc;
    const int k0 = kci * kNC, n0 = nci * NR;
    float* dst = Wst + stage * STAGE;
    const int cols = min(kNC, round_up4(Kout - k0));
    const int qpr = cols / 4;
    for (int seg = tid; seg < NR * qpr; seg += kThreads) {
      const int row = seg / qpr, q = seg - row * qpr;
      const int n = n0 + row, k = k0 + 4 * q;
      float* d = dst + row * LW + 4 * q;
      if (n < N && vec && k + 3 < Kout) {
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
  };

  float acc[TM][4];
  load_chunk(0, 0);
  cp_async_commit();
  for (int c = 0; c < total; ++c) {
    const int kci = c / nnc, nci = c - kci * nnc;
    const int k0 = kci * kNC, n0 = nci * NR;
    if (nci == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
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
    const int kcol = k0 + 4 * tx;
    if (kcol < Kout) {
      const int nlen = min(NR, round_up4(N - n0));
      const float* zrow = dZs + (ty * TM) * ldz + n0;
#pragma unroll 2
      for (int nn = 0; nn < nlen; nn += 4) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = *reinterpret_cast<const float4*>(Ws + (nn + j) * LW + 4 * tx);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float4 z = *reinterpret_cast<const float4*>(zrow + i * ldz + nn);
          acc[i][0] = fmaf(z.x, w[0].x, acc[i][0]);
          acc[i][1] = fmaf(z.x, w[0].y, acc[i][1]);
          acc[i][2] = fmaf(z.x, w[0].z, acc[i][2]);
          acc[i][3] = fmaf(z.x, w[0].w, acc[i][3]);
          acc[i][0] = fmaf(z.y, w[1].x, acc[i][0]);
          acc[i][1] = fmaf(z.y, w[1].y, acc[i][1]);
          acc[i][2] = fmaf(z.y, w[1].z, acc[i][2]);
          acc[i][3] = fmaf(z.y, w[1].w, acc[i][3]);
          acc[i][0] = fmaf(z.z, w[2].x, acc[i][0]);
          acc[i][1] = fmaf(z.z, w[2].y, acc[i][1]);
          acc[i][2] = fmaf(z.z, w[2].z, acc[i][2]);
          acc[i][3] = fmaf(z.z, w[2].w, acc[i][3]);
          acc[i][0] = fmaf(z.w, w[3].x, acc[i][0]);
          acc[i][1] = fmaf(z.w, w[3].y, acc[i][1]);
          acc[i][2] = fmaf(z.w, w[3].z, acc[i][2]);
          acc[i][3] = fmaf(z.w, w[3].w, acc[i][3]);
        }
      }
    }
    if (nci == nnc - 1 && kcol < round_up4(Kout)) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
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
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// cooperative tile <-> global copies (all 256 threads)
// ---------------------------------------------------------------------------
// smem[r, 0..round_up4(D)) <- g[(row0+r), 0..D), zero padded; rows >= nrows zeroed.
template <int R>
__device__ void tile_load_rows(float* __restrict__ s, int lds, const float* __restrict__ g,
                               int ldg, int D, int row0, int nrows) {
  const int d4 = round_up4(D) / 4;
  const bool vec = ((ldg & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  for (int idx = threadIdx.x; idx < R * d4; idx += kThreads) {
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
template <int R>
__device__ void tile_store_rows(const float* __restrict__ s, int lds, float* __restrict__ g,
                                int ldg, int D, int row0, int nrows) {
  const int d4 = round_up4(D) / 4;
  const bool vec = ((ldg & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  for (int idx = threadIdx.x; idx < R * d4; idx += kThreads) {
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
template <int TM, int KC>
__device__ void tile_mlp_fwd(const Mlp& net, const float* in, int ld_in, float* hA, float* hB,
                             int ldh, float* out, int ld_out, float* Wst,
                             float* const* save, int row0, int nrows) {
  constexpr int R = 4 * TM;
  const float* cur = in;
  int ldc = ld_in;
  for (int l = 0; l < net.n_layers; ++l) {
    const bool last = (l == net.n_layers - 1);
    float* dst = last ? out : ((l & 1) ? hB : hA);
    const int ldd = last ? ld_out : ldh;
    tile_linear_fwd<TM, KC>(cur, ldc, net.dims[l], net.params + net.w_off[l], net.dims[l],
                            net.params + net.b_off[l], net.dims[l + 1], net.act[l], dst, ldd,
                            Wst);
    if (!last && save != nullptr && save[l] != nullptr)
      tile_store_rows<R>(dst, ldd, save[l], net.dims[l + 1], net.dims[l + 1], row0, nrows);
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
template <int TM, int KC>
__device__ void tile_mlp_bwd(const Mlp& net, float* dz_last, int ld_last, float* gA, float* gB,
                             float* hbuf, int ldg, float* Wst, const float* const* hidden,
                             float* const* dz_out, int row0, int nrows, float* din, int ld_din,
                             int in_col0, int in_cols) {
  constexpr int R = 4 * TM;
  const int L = net.n_layers;
  float* cur = dz_last;
  int ldc = ld_last;
  if (dz_out != nullptr && dz_out[L - 1] != nullptr)
    tile_store_rows<R>(cur, ldc, dz_out[L - 1], net.dims[L], net.dims[L], row0, nrows);
  for (int l = L - 1; l >= 1; --l) {
    // dZ_{l-1} = (dZ_l . W_l) * act'_{l-1}(H_{l-1})
    tile_load_rows<R>(hbuf, ldg, hidden[l - 1], net.dims[l], net.dims[l], row0, nrows);
    __syncthreads();
    float* dst = (l & 1) ? gA : gB;
    tile_linear_bwd<TM, KC>(cur, ldc, net.dims[l + 1], net.params + net.w_off[l], net.dims[l],
                            net.dims[l], hbuf, ldg, net.act[l - 1], dst, ldg, Wst);
    if (dz_out != nullptr && dz_out[l - 1] != nullptr)
      tile_store_rows<R>(dst, ldg, dz_out[l - 1], net.dims[l], net.dims[l], row0, nrows);
    cur = dst;
    ldc = ldg;
  }
  if (din != nullptr) {
    tile_linear_bwd<TM, KC>(cur, ldc, net.dims[1], net.params + net.w_off[0] + in_col0,
                            net.dims[0], in_cols, nullptr, 0, 0, din, ld_din, Wst);
  }
}

}  // namespace rb200
