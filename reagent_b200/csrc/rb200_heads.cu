// reagent_b200 -- loss heads of the two remaining DQN-family trainers.
// The networks around them are plain MLPs evaluated by the generic forward / backward /
// weight-gradient kernels of this library; these kernels do what sits in between.
//
//   rb200_pdqn_head  ParametricDQNTrainer.train_step_gen
//       reagent/training/parametric_dqn_trainer.py:109-173: masked (double-)max over the tiled
//       possible next actions (dqn_trainer_base.py:33-77) or the SARSA value, TD target,
//       mse | huber loss and d loss / d q.
//   rb200_c51_head   C51Trainer.train_step_gen, reagent/training/c51_trainer.py:98-173:
//       log-softmax over the atoms (reagent/models/categorical_dqn.py:33-35), expected values,
//       masked arg max, target distribution, categorical projection onto the support (with the
//       reference's l == u corner-case adjustment), cross-entropy loss and d loss / d logits.
#include "rb200_common.cuh"

namespace rb200 {

// ---------------------------------------------------------------------------
struct PdqnDev {
  rb200_pdqn_args_t a;
};

__global__ void __launch_bounds__(256) pdqn_head_kernel(const PdqnDev d) {
  const rb200_pdqn_args_t& a = d.a;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float le = 0.f;
  if (b < a.batch) {
    float next_q;
    const int M = a.max_num_action;
    if (M > 0) {
      // get_max_q_values_with_target: q += -1e9 * (1 - mask) on both nets; double-Q: arg max
      // of the online values, value of the target net there; else max of the target values
      float best = 0.f, sel = 0.f;
      int bi = -1;
      for (int c = 0; c < M; ++c) {
        const float pen = -1e9f * (1.f - (a.mask ? a.mask[(size_t)b * M + c] : 1.f));
        const float vt = a.next_q_target[(size_t)b * M + c] + pen;
        const float key = (a.double_q && a.next_q) ? a.next_q[(size_t)b * M + c] + pen : vt;
        if (bi < 0 || key > best) { best = key; bi = c; sel = vt; }
      }
      next_q = sel;
    } else {
      next_q = a.next_q_target[b];
    }
    const float disc = (a.discount_mode == RB200_DISCOUNT_POW && a.discount_src)
                           ? powf(a.gamma, a.discount_src[b]) : a.gamma;
    // parametric_dqn_trainer.py:159: reward + not_terminal * discount * next_q
    const float tgt = a.reward[b] + (a.not_terminal[b] * disc) * next_q;
    const float dq = a.q_values[b] - tgt;
    const float inv = 1.f / (float)a.batch;
    float g;
    if (a.loss_kind == RB200_LOSS_HUBER) {
      const float ad = fabsf(dq);
      le = ad < 1.f ? 0.5f * dq * dq : ad - 0.5f;
      g = (dq < -1.f ? -1.f : (dq > 1.f ? 1.f : dq)) * inv;
    } else {
      le = dq * dq;
      g = 2.f * dq * inv;
    }
    a.dz[b] = g;
    if (a.td_target) a.td_target[b] = tgt;
  }
  __shared__ float s_l[8];
  le = warp_sum(le);
  if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = le;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_l[w];
    a.loss_partials[blockIdx.x] = t;
    __threadfence();
    const unsigned fin = atomicAdd(a.tile_counter, 1u);
    if (fin == gridDim.x - 1) {
      __threadfence();
      float tot = 0.f;
      for (unsigned i = 0; i < gridDim.x; ++i) tot += ((volatile float*)a.loss_partials)[i];
      *a.loss = tot / (float)a.batch;
      *a.tile_counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// C51: one CTA per batch row; smem: logits of three [A, N] heads as log-probabilities
// ---------------------------------------------------------------------------
struct C51Dev {
  rb200_c51_args_t a;
};

// in-place log_softmax over the last dim of x[A][N] (one warp per action row, round-robin)
__device__ void c51_log_softmax(float* x, int A, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < A; r += nw) {
    float* row = x + (size_t)r * N;
    float mx = -INFINITY;
    for (int c = lane; c < N; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += expf(row[c] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    for (int c = lane; c < N; c += 32) row[c] -= lse;
  }
}

__global__ void __launch_bounds__(256) c51_head_kernel(const C51Dev d) {
  const rb200_c51_args_t& a = d.a;
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int A = a.num_actions, N = a.num_atoms, AN = A * N;
  float* lt = sm;             // log dist of the target net on s'      [A][N]
  float* lo = lt + AN;        // log dist of the online net on s' (double-Q) or alias of lt
  float* lc = lo + AN;        // log dist of the online net on s       [A][N]
  float* nd = lc + AN;        // next_dist of the chosen action         [N]
  float* m = nd + N;          // projected target distribution          [N]
  float* qv = m + N;          // expected values per action             [A]
  __shared__ int s_next;
  for (int i = tid; i < AN; i += blockDim.x) {
    lt[i] = a.logits_next_target[(size_t)b * AN + i];
    lc[i] = a.logits_cur[(size_t)b * AN + i];
    if (a.logits_next_online) lo[i] = a.logits_next_online[(size_t)b * AN + i];
  }
  __syncthreads();
  c51_log_softmax(lt, A, N);
  c51_log_softmax(lc, A, N);
  if (a.logits_next_online) c51_log_softmax(lo, A, N);
  __syncthreads();
  // expected next values: (dist * support).sum(2), c51_trainer.py:117-124
  const float* lq = (a.double_q && a.logits_next_online) ? lo : lt;
  if (a.maxq) {
    for (int r = tid >> 5; r < A; r += blockDim.x >> 5) {
      float s = 0.f;
      for (int c = tid & 31; c < N; c += 32) s += expf(lq[(size_t)r * N + c]) * a.support[c];
      s = warp_sum(s);
      if ((tid & 31) == 0) qv[r] = s;
    }
    __syncthreads();
    if (tid == 0) {  // argmax_with_mask: q + -1e9 * (1 - mask), first maximum
      float best = 0.f;
      int bi = -1;
      for (int c = 0; c < A; ++c) {
        const float mk = a.possible_next_actions_mask ? a.possible_next_actions_mask[(size_t)b * A + c] : 1.f;
        const float v = qv[c] + (-1e9f) * (1.f - mk);
        if (bi < 0 || v > best) { best = v; bi = c; }
      }
      s_next = bi;
      if (a.next_action_idx) a.next_action_idx[b] = bi;
    }
    __syncthreads();
    for (int c = tid; c < N; c += blockDim.x) nd[c] = expf(lt[(size_t)s_next * N + c]);
  } else {  // SARSA: (next_dist * next_action.unsqueeze(-1)).sum(1)
    for (int c = tid; c < N; c += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < A; ++r) s += expf(lt[(size_t)r * N + c]) * a.next_action[(size_t)b * A + r];
      nd[c] = s;
    }
  }
  for (int c = tid; c < N; c += blockDim.x) m[c] = 0.f;
  __syncthreads();
  // target support, projection (c51_trainer.py:135-160); atoms in order so that the
  // scatter_add of the reference (index order) is reproduced per destination
  if (tid == 0) {
    float rew = a.reward[b];
    if (a.reward_boost) {
      float bs = 0.f;
      for (int c = 0; c < A; ++c) bs += a.action[(size_t)b * A + c] * a.reward_boost[c];
      rew += bs;
    }
    const float disc = (a.discount_src) ? powf(a.gamma, a.discount_src[b]) : a.gamma;
    const float nt = a.not_terminal[b];
    for (int j = 0; j < N; ++j) {
      float tq = rew + (disc * nt) * a.support[j];
      tq = fminf(fmaxf(tq, a.qmin), a.qmax);
      const float bb = (tq - a.qmin) / a.scale_support;
      long long l = (long long)floorf(bb), u = (long long)ceilf(bb);
      if (u > 0 && l == u) l -= 1;
      if (l < N - 1 && l == u) u += 1;
      m[l] += nd[j] * ((float)u - bb);
    }
    for (int j = 0; j < N; ++j) {
      float tq = rew + (disc * nt) * a.support[j];
      tq = fminf(fmaxf(tq, a.qmin), a.qmax);
      const float bb = (tq - a.qmin) / a.scale_support;
      long long l = (long long)floorf(bb), u = (long long)ceilf(bb);
      if (u > 0 && l == u) l -= 1;
      if (l < N - 1 && l == u) u += 1;
      m[u] += nd[j] * (bb - (float)l);
    }
  }
  __syncthreads();
  // loss = -(m * (log_dist * action).sum(1)).sum(1).mean(); gradient w.r.t. the logits of s:
  // d/dlogit[r][c] = action[r] * (softmax[r][c] * sum_j m_j - m_c) / B
  float msum = 0.f;
  for (int c = 0; c < N; ++c) msum += m[c];
  float le = 0.f;
  const float invB = 1.f / (float)a.batch;
  for (int i = tid; i < AN; i += blockDim.x) {
    const int r = i / N, c = i - r * N;
    const float aw = a.action[(size_t)b * A + r];
    le -= m[c] * lc[i] * aw;
    a.dz_logits[(size_t)b * AN + i] = aw * (expf(lc[i]) * msum - m[c]) * invB;
  }
  if (a.all_q_values) {
    for (int r = tid >> 5; r < A; r += blockDim.x >> 5) {
      float s = 0.f;
      for (int c = tid & 31; c < N; c += 32) s += expf(lc[(size_t)r * N + c]) * a.support[c];
      s = warp_sum(s);
      if ((tid & 31) == 0) a.all_q_values[(size_t)b * A + r] = s;
    }
  }
  __shared__ float s_l[8];
  __shared__ bool s_last;
  le = warp_sum(le);
  if ((tid & 31) == 0) s_l[tid >> 5] = le;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_l[w];
    a.loss_partials[b] = t;
    __threadfence();
    const unsigned fin = atomicAdd(a.tile_counter, 1u);
    s_last = fin == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    // one CTA per row: the last one adds the per-row partials with all its threads (thread t
    // takes rows t, t + blockDim, ...; fixed combination order) instead of one thread walking
    // `batch` dependent loads at the tail of the kernel
    __threadfence();
    float tot = 0.f;
    for (unsigned i = tid; i < gridDim.x; i += blockDim.x) tot += ((volatile float*)a.loss_partials)[i];
    tot = warp_sum(tot);
    __syncthreads();
    if ((tid & 31) == 0) s_l[tid >> 5] = tot;
    __syncthreads();
    if (tid == 0) {
      float t2 = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t2 += s_l[w];
      *a.loss = t2 * invB;
      *a.tile_counter = 0u;
    }
  }
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_pdqn_head(const rb200_pdqn_args_t* a, void* stream) {
  if (!a || a->batch <= 0 || a->max_num_action < 0) { set_last_error("rb200_pdqn_head: bad argument"); return RB200_E_INVALID; }
  if (!a->next_q_target || !a->reward || !a->not_terminal || !a->q_values || !a->dz ||
      !a->loss_partials || !a->loss || !a->tile_counter) { set_last_error("rb200_pdqn_head: required pointer is null"); return RB200_E_INVALID; }
  if (a->discount_mode == RB200_DISCOUNT_POW && !a->discount_src) { set_last_error("POW discount needs discount_src"); return RB200_E_INVALID; }
  PdqnDev d;
  d.a = *a;
  pdqn_head_kernel<<<ceil_div(a->batch, 256), 256, 0, (cudaStream_t)stream>>>(d);
  return check_cuda(cudaGetLastError(), "pdqn_head_kernel launch");
}

extern "C" int rb200_c51_head(const rb200_c51_args_t* a, void* stream) {
  if (!a || a->batch <= 0 || a->num_actions <= 0 || a->num_atoms < 2) { set_last_error("rb200_c51_head: bad argument"); return RB200_E_INVALID; }
  if (!a->logits_next_target || !a->logits_cur || !a->action || !a->reward || !a->not_terminal ||
      !a->support || !a->dz_logits || !a->loss_partials || !a->loss || !a->tile_counter) { set_last_error("rb200_c51_head: required pointer is null"); return RB200_E_INVALID; }
  if (!a->maxq && !a->next_action) { set_last_error("SARSA update needs next_action"); return RB200_E_INVALID; }
  if (a->double_q && a->maxq && !a->logits_next_online) { set_last_error("double-Q needs logits_next_online"); return RB200_E_INVALID; }
  C51Dev d;
  d.a = *a;
  const size_t smem = ((size_t)3 * a->num_actions * a->num_atoms + 2 * a->num_atoms + a->num_actions) * sizeof(float);
  if (smem > 200 * 1024) { set_last_error("rb200_c51_head: too many atoms/actions for one CTA"); return RB200_E_SMEM; }
  static SmemOptIn optin = {};
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dynamic_smem(c51_head_kernel, optin, smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(c51_head)");
  }
  c51_head_kernel<<<a->batch, 256, smem, (cudaStream_t)stream>>>(d);
  return check_cuda(cudaGetLastError(), "c51_head_kernel launch");
}
