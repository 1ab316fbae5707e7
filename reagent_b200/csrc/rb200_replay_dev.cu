// reagent_b200 -- device-resident replay bookkeeping: batched `add`,
// `set_priority` and the prioritized index draw, all on the GPU, so that an online loop
// (add one transition -> draw a minibatch -> train) is ONE CUDA-graph replay per step with the
// new transition as its only host->device traffic.
//
// Restates, with identical results (tests compare against the host path and the reference's
// golden vectors):
//   ReplayBuffer.add / _add_transition, validity bookkeeping
//                                    reagent/replay_memory/circular_replay_buffer.py:468-547, :430-438
//   PrioritizedReplayBuffer._add     reagent/replay_memory/prioritized_replay_buffer.py:62-84
//   SumTree.set (sequential fp64 delta propagation)        reagent/replay_memory/sum_tree.py:164-189
//   SumTree.stratified_sample / sample                     sum_tree.py:93-153
//   PrioritizedReplayBuffer.sample_index_batch (retries)   prioritized_replay_buffer.py:86-115
//   random.uniform / random.random of CPython (MT19937, Modules/_randommodule.c; the generator
//   is not part of the reference repo -- the stdlib module the reference calls at
//   sum_tree.py:113,152): state kept ON THE DEVICE in CPython's own layout (624 words +
//   position), uploaded from / downloaded to `random.getstate()` by the host wrapper.
//
// All three kernels are single-CTA: the work is a few microseconds of latency-bound
// sequential-semantics bookkeeping that runs on a side stream underneath the TD kernel.
#include "rb200_common.cuh"

namespace rb200 {

constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far_) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far_ ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// state regeneration by the whole CTA (>= 256 threads): three phases whose reads all precede
// their writes (compute -> barrier -> store), then the last word
__device__ void mt_twist_block(uint32_t* s) {
  const int tid = threadIdx.x, nt = blockDim.x;
  uint32_t v[3];
  const int lo[3] = {0, kMtN - kMtM, 2 * (kMtN - kMtM)}, hi[3] = {kMtN - kMtM, 2 * (kMtN - kMtM), kMtN - 1};
  for (int ph = 0; ph < 3; ++ph) {
    int cnt = 0;
    for (int k = lo[ph] + tid; k < hi[ph]; k += nt)
      v[cnt++] = mt_mix(s[k], s[k + 1], ph == 0 ? s[k + kMtM] : s[k + (kMtM - kMtN)]);
    __syncthreads();
    cnt = 0;
    for (int k = lo[ph] + tid; k < hi[ph]; k += nt) s[k] = v[cnt++];
    __syncthreads();
  }
  if (tid == 0) s[kMtN - 1] = mt_mix(s[kMtN - 1], s[0], s[kMtM - 1]);
  __syncthreads();
}
// the same by ONE thread (retry path: a handful of extra draws after the stratified ones)
__device__ void mt_twist_serial(uint32_t* s) {
  int k = 0;
  for (; k < kMtN - kMtM; ++k) s[k] = mt_mix(s[k], s[k + 1], s[k + kMtM]);
  for (; k < kMtN - 1; ++k) s[k] = mt_mix(s[k], s[k + 1], s[k + (kMtM - kMtN)]);
  s[kMtN - 1] = mt_mix(s[kMtN - 1], s[0], s[kMtM - 1]);
}
// random.random(): 53-bit double in [0, 1) from two outputs (a >> 5, b >> 6)
__device__ __forceinline__ double mt_double(uint32_t a, uint32_t b) {
  return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)),
                   1.0 / 9007199254740992.0);
}

// SumTree.sample's descent (sum_tree.py:112-131) for a query already scaled by the root
__device__ __forceinline__ long long tree_walk(const double* __restrict__ tree, const double* top,
                                               int top_levels, int depth, double q) {
  long long node = 0;
  for (int lvl = 1; lvl <= depth; ++lvl) {
    const long long left = node * 2;
    const long long pos = ((1ll << lvl) - 1) + left;
    const double left_sum = (lvl < top_levels) ? top[pos] : __ldcg(tree + pos);
    if (q < left_sum) {
      node = left;
    } else {
      node = left + 1;
      q -= left_sum;
    }
  }
  return node;
}

constexpr int kDrawThreads = 1024;
constexpr int kDrawTop = 10;       // tree levels cached in shared memory (8 KB)
constexpr int kDrawChunk = 4096;   // strata per pass (32 KB of raw outputs)
constexpr int kDrawPer = kDrawChunk / kDrawThreads;  // descents per thread, interleaved

struct DrawDev {
  rb200_per_draw_args_t a;
};

__global__ void __launch_bounds__(kDrawThreads) per_draw_indices_kernel(const DrawDev d) {
  const rb200_per_draw_args_t& a = d.a;
  __shared__ uint32_t s_mt[kMtN];
  __shared__ uint32_t s_raw[2 * kDrawChunk];
  __shared__ double s_top[(1 << kDrawTop) - 1];
  __shared__ int s_pos, s_any;
  const int tid = threadIdx.x;
  for (int i = tid; i < kMtN; i += kDrawThreads) s_mt[i] = a.mt_state[i];
  const int top = min(a.tree_depth + 1, kDrawTop);
  for (int i = tid; i < (1 << top) - 1; i += kDrawThreads) s_top[i] = __ldcg(a.tree + i);
  if (tid == 0) { s_pos = (int)a.mt_state[kMtN]; s_any = 0; }
  __syncthreads();
  const double root = s_top[0];
  int pos = s_pos;
  bool any_invalid = false;
  for (int base = 0; base < a.batch; base += kDrawChunk) {
    const int m = min(kDrawChunk, a.batch - base);
    // ---- 2m tempered outputs of the stream ----
    int produced = 0;
    while (produced < 2 * m) {
      if (pos >= kMtN) { mt_twist_block(s_mt); pos = 0; }
      const int take = min(kMtN - pos, 2 * m - produced);
      for (int i = tid; i < take; i += kDrawThreads) s_raw[produced + i] = mt_temper(s_mt[pos + i]);
      pos += take;
      produced += take;
      __syncthreads();
    }
    // ---- stratified queries (sum_tree.py:149-152) and their descents: each thread walks
    // kDrawPer strata in lock-step so that their dependent loads overlap ----
    {
      double q[kDrawPer];
      long long node[kDrawPer];
      int bb[kDrawPer];
#pragma unroll
      for (int u = 0; u < kDrawPer; ++u) {
        const int i = tid + u * kDrawThreads;
        bb[u] = (i < m) ? base + i : -1;
        node[u] = 0;
        q[u] = 0.0;
        if (bb[u] >= 0) {
          const double r = mt_double(s_raw[2 * i], s_raw[2 * i + 1]);
          // random.uniform(lo, hi) = lo + (hi - lo) * random(): separately rounded operations
          const double lo = a.lo[bb[u]], hi = a.hi[bb[u]];
          const double qq = __dadd_rn(lo, __dmul_rn(__dadd_rn(hi, -lo), r));
          if (a.queries_out) a.queries_out[bb[u]] = qq;
          q[u] = __dmul_rn(qq, root);  // sum_tree.py:113
        }
      }
      for (int lvl = 1; lvl <= a.tree_depth; ++lvl) {
        double ls[kDrawPer];
#pragma unroll
        for (int u = 0; u < kDrawPer; ++u) {
          const long long p = ((1ll << lvl) - 1) + node[u] * 2;
          ls[u] = (lvl < top) ? s_top[p] : __ldcg(a.tree + p);
        }
#pragma unroll
        for (int u = 0; u < kDrawPer; ++u) {
          if (q[u] < ls[u]) {
            node[u] = node[u] * 2;
          } else {
            node[u] = node[u] * 2 + 1;
            q[u] -= ls[u];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kDrawPer; ++u) {
        if (bb[u] < 0) continue;
        a.indices_out[bb[u]] = node[u];
        if (!a.valid[node[u]]) any_invalid = true;
      }
    }
    __syncthreads();
  }
  if (any_invalid) s_any = 1;
  __syncthreads();
  // ---- retries, sequential as in prioritized_replay_buffer.py:95-113 ----
  if (tid == 0) {
    int used = 0;
    if (s_any) {
      int allowed = a.max_attempts;
      for (int b = 0; b < a.batch; ++b) {
        long long index = a.indices_out[b];
        if (a.valid[index]) continue;
        if (allowed == 0) { a.status[0] = 1; break; }  // "Max sample attempts" (sticky; host raises)
        while (!a.valid[index] && allowed > 0) {
          uint32_t w[2];
          for (int j = 0; j < 2; ++j) {
            if (pos >= kMtN) { mt_twist_serial(s_mt); pos = 0; }
            w[j] = mt_temper(s_mt[pos++]);
          }
          index = tree_walk(a.tree, s_top, top, a.tree_depth, __dmul_rn(mt_double(w[0], w[1]), root));
          --allowed;
          ++used;
        }
        a.indices_out[b] = index;
      }
    }
    a.status[1] = used;
    s_pos = pos;
  }
  __syncthreads();
  for (int i = tid; i < kMtN; i += kDrawThreads) a.mt_state[i] = s_mt[i];
  if (tid == 0) a.mt_state[kMtN] = (uint32_t)s_pos;
}

// ---------------------------------------------------------------------------
// SumTree.set for a batch, applied IN ORDER (sum_tree.py:164-189): one warp, lane l owns level
// l of the root path, so the per-node order of the fp64 additions is the reference's
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tree_set_warp(double* tree, int depth, long long leaf, double value,
                                              double* max_recorded, int lane) {
  // lane `depth` reads (and later writes) the leaf: same thread, program order
  double delta = 0.0;
  if (lane == depth) {
    delta = value - tree[((1ll << depth) - 1) + leaf];
    if (max_recorded && value > *max_recorded) *max_recorded = value;
  }
  delta = __shfl_sync(0xffffffffu, delta, depth);
  if (lane <= depth) {
    double* p = tree + ((1ll << lane) - 1) + (leaf >> (depth - lane));
    *p = __dadd_rn(*p, delta);
  }
  __syncwarp();
}

__global__ void sumtree_set_kernel(double* tree, int depth, const long long* idx, const double* val,
                                   int n, double* max_recorded, int* status) {
  const int lane = threadIdx.x;
  for (int i = 0; i < n; ++i) {
    const double v = val[i];
    if (v < 0.0) { if (lane == 0 && status) status[0] = 2; return; }  // "values should be nonnegative"
    tree_set_warp(tree, depth, idx[i], v, max_recorded, lane);
  }
}

// ---------------------------------------------------------------------------
// n consecutive ReplayBuffer.add() calls (stack_size == 1) from device staging rows
// ---------------------------------------------------------------------------
constexpr int kAddMax = 1024;

struct AddDev {
  rb200_add_args_t a;
};

__global__ void __launch_bounds__(256) replay_add_kernel(const AddDev d) {
  const rb200_add_args_t& a = d.a;
  const rb200_replay_dev_t& rb = a.rb;
  __shared__ long long s_cur[kAddMax];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long cap = rb.capacity;
  if (tid == 0) {
    // circular_replay_buffer.py:468-522 with stack_size == 1 (no padding transitions)
    long long add_count = rb.state[0], ep = rb.state[1], nvalid = rb.state[2];
    auto set_valid = [&](long long i, bool v) {  // set_index_valid_status, :430-438
      const bool old = rb.valid[i] != 0;
      if (old != v) { rb.valid[i] = v ? 1 : 0; nvalid += v ? 1 : -1; }
    };
    for (int t = 0; t < a.n; ++t) {
      const long long cur = add_count % cap;
      const long long last = (cur - 1 + cap) % cap;
      if (add_count == 0 || rb.terminal[last]) ep = 0;
      set_valid(cur, false);
      if (ep >= rb.update_horizon) set_valid(((cur - rb.update_horizon) % cap + cap) % cap, true);
      rb.terminal[cur] = a.terminal_in[t] ? 1 : 0;
      rb.reward[cur] = a.reward_in[t];
      s_cur[t] = cur;
      ++add_count;
      ++ep;
      if (a.terminal_in[t]) {
        const long long back = ep < rb.update_horizon ? ep : rb.update_horizon;
        for (long long k = 0; k < back; ++k) set_valid(((cur - k) % cap + cap) % cap, true);
      }
    }
    rb.state[0] = add_count;
    rb.state[1] = ep;
    rb.state[2] = nvalid;
  }
  __syncthreads();
  // priorities: prioritized_replay_buffer.py:76-84 -> SumTree.set(cursor, priority), in order
  if (warp == 0 && rb.tree && a.priority_in) {
    for (int t = 0; t < a.n; ++t) {
      const double v = a.priority_in[t];
      if (v < 0.0) { if (lane == 0) rb.state[3] = 2; break; }
      tree_set_warp(rb.tree, rb.tree_depth, s_cur[t], v, rb.max_priority, lane);
    }
  }
  // row copies (observation, action, extras): staging [n, row_bytes] -> store[cursor]
  for (int t = 0; t < a.n; ++t) {
    for (int g = 0; g < a.n_rows; ++g) {
      const rb200_gather_spec_t& sp = a.rows[g];
      const unsigned char* src = (const unsigned char*)sp.src + (size_t)t * sp.row_bytes;
      unsigned char* dst = (unsigned char*)sp.dst + (size_t)s_cur[t] * sp.row_bytes;
      if (((sp.row_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
          ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        for (int c = tid; c < sp.row_bytes / 16; c += blockDim.x)
          reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
      } else {
        for (int c = tid; c < sp.row_bytes; c += blockDim.x) dst[c] = src[c];
      }
    }
  }
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_per_draw_indices(const rb200_per_draw_args_t* a, void* stream) {
  if (!a || !a->mt_state || !a->lo || !a->hi || !a->tree || !a->valid || !a->indices_out || !a->status) {
    set_last_error("rb200_per_draw_indices: null argument"); return RB200_E_INVALID;
  }
  if (a->batch <= 0 || a->tree_depth < 0 || a->max_attempts < 0) { set_last_error("rb200_per_draw_indices: bad batch/depth/attempts"); return RB200_E_INVALID; }
  DrawDev d;
  d.a = *a;
  per_draw_indices_kernel<<<1, kDrawThreads, 0, (cudaStream_t)stream>>>(d);
  return check_cuda(cudaGetLastError(), "per_draw_indices_kernel launch");
}

extern "C" int rb200_sumtree_set_device(double* tree, int32_t depth, const int64_t* idx,
                                        const double* val, int32_t n, double* max_recorded,
                                        int32_t* status, void* stream) {
  if (!tree || !idx || !val || depth < 0 || depth > 31 || n < 0) { set_last_error("rb200_sumtree_set_device: bad argument"); return RB200_E_INVALID; }
  if (n == 0) return RB200_OK;
  sumtree_set_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(tree, depth, (const long long*)idx, val, n,
                                                         max_recorded, status);
  return check_cuda(cudaGetLastError(), "sumtree_set_kernel launch");
}

extern "C" int rb200_replay_add_device(const rb200_add_args_t* a, void* stream) {
  if (!a || !a->rb.state || !a->rb.valid || !a->rb.terminal || !a->rb.reward || !a->terminal_in || !a->reward_in) {
    set_last_error("rb200_replay_add_device: null argument"); return RB200_E_INVALID;
  }
  if (a->n <= 0 || a->n > kAddMax) { set_last_error("rb200_replay_add_device: n must be in [1, %d]", kAddMax); return RB200_E_INVALID; }
  if (a->rb.capacity <= 0 || a->rb.update_horizon <= 0 || a->n_rows < 0 || a->n_rows > RB200_MAX_GATHER_SPECS) { set_last_error("rb200_replay_add_device: bad capacity/horizon/rows"); return RB200_E_INVALID; }
  if (a->rb.tree && (a->rb.tree_depth < 0 || a->rb.tree_depth > 31)) { set_last_error("rb200_replay_add_device: bad tree depth"); return RB200_E_INVALID; }
  AddDev d;
  d.a = *a;
  replay_add_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d);
  return check_cuda(cudaGetLastError(), "replay_add_kernel launch");
}
