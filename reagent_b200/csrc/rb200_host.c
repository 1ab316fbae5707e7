/* reagent_b200 -- host-side helpers of the replay path (plain C, no CUDA).
 *
 * MT19937: the generator behind Python's `random` (CPython Modules/_randommodule.c,
 * algorithm of Matsumoto & Nishimura 2002), restated so that the stratified query values
 * of SumTree.stratified_sample (reagent/replay_memory/sum_tree.py:149-153:
 * random.uniform(lo, hi) = lo + (hi-lo)*random()) are reproduced bit for bit at C speed
 * from the interpreter's own state (random.getstate() / setstate()).
 */
#include <stdint.h>
#include <stddef.h>

#include "../../include/reagent_b200.h"

#define MT_N 624
#define MT_M 397

/* state regeneration ("twist"): three branch-free loops whose reads run ahead of their writes,
 * so the compiler can vectorise them */
static void mt_twist(uint32_t* mt) {
  int kk;
  for (kk = 0; kk < MT_N - MT_M; kk++) {
    const uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
    mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 0x1U)) & 0x9908b0dfU);
  }
  for (; kk < MT_N - 1; kk++) {
    const uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
    mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 0x1U)) & 0x9908b0dfU);
  }
  {
    const uint32_t y = (mt[MT_N - 1] & 0x80000000U) | (mt[0] & 0x7fffffffU);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 0x1U)) & 0x9908b0dfU);
  }
}

/* the next `count` tempered 32-bit outputs of the stream (genrand_uint32 x count) */
static void mt_fill(uint32_t* mt, int32_t* index, uint32_t* dst, int64_t count) {
  while (count > 0) {
    if (*index >= MT_N) {
      mt_twist(mt);
      *index = 0;
    }
    int64_t take = MT_N - *index;
    if (take > count) take = count;
    const uint32_t* src = mt + *index;
    for (int64_t i = 0; i < take; ++i) {
      uint32_t y = src[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680U;
      y ^= (y << 15) & 0xefc60000U;
      y ^= (y >> 18);
      dst[i] = y;
    }
    dst += take;
    count -= take;
    *index += (int32_t)take;
  }
}

void rb200_mt19937_uniform_host(uint32_t* state624, int32_t* index, const double* lo,
                                const double* hi, double* out, int64_t n) {
  enum { CHUNK = 1024 };
  uint32_t raw[2 * CHUNK];
  for (int64_t base = 0; base < n; base += CHUNK) {
    const int64_t m = (n - base < CHUNK) ? n - base : CHUNK;
    mt_fill(state624, index, raw, 2 * m);
    for (int64_t i = 0; i < m; ++i) {
      /* random(): 53-bit resolution double in [0,1) */
      const uint32_t a = raw[2 * i] >> 5, b = raw[2 * i + 1] >> 6;
      const double r = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
      out[base + i] = (lo && hi) ? lo[base + i] + (hi[base + i] - lo[base + i]) * r : r;
    }
  }
}

int rb200_sumtree_set_host(double* tree, int32_t depth, const int64_t* idx, const double* val,
                           int64_t n, double* max_recorded) {
  for (int64_t i = 0; i < n; ++i) {
    const double value = val[i];
    if (value < 0.0) return -1;
    if (max_recorded && value > *max_recorded) *max_recorded = value;
    int64_t node = idx[i];
    const double delta = value - tree[((int64_t)1 << depth) - 1 + node];
    for (int32_t lvl = depth; lvl >= 0; --lvl) {
      tree[((int64_t)1 << lvl) - 1 + node] += delta;
      node /= 2;
    }
  }
  return 0;
}

int64_t rb200_sumtree_sample_host(const double* tree, int32_t depth, double query) {
  double q = query * tree[0];
  int64_t node = 0;
  for (int32_t lvl = 1; lvl <= depth; ++lvl) {
    const int64_t left = node * 2;
    const double left_sum = tree[((int64_t)1 << lvl) - 1 + left];
    if (q < left_sum) {
      node = left;
    } else {
      node = left + 1;
      q -= left_sum;
    }
  }
  return node;
}

/* The walk above for the strata `pos[0..n)` of one stratified draw; out[j] = leaf reached by
 * queries[pos[j]].  One call per update instead of one per candidate stratum. */
void rb200_sumtree_sample_many_host(const double* tree, int32_t depth, const double* queries,
                                    const int64_t* pos, int64_t n, int64_t* out) {
  for (int64_t j = 0; j < n; ++j) out[j] = rb200_sumtree_sample_host(tree, depth, queries[pos[j]]);
}

/* Validity bookkeeping of N consecutive ReplayBuffer.add() calls with stack_size == 1
 * (reagent/replay_memory/circular_replay_buffer.py:468-522), applied to host arrays.
 * state[0]=add_count, state[1]=transitions in current episode, state[2]=num valid. */
void rb200_replay_add_batch_host(const uint8_t* terminal_in, int64_t n, int64_t capacity,
                                 int32_t update_horizon, uint8_t* valid, uint8_t* terminal_store,
                                 int64_t* state) {
  int64_t add_count = state[0], ep = state[1], nvalid = state[2];
  for (int64_t t = 0; t < n; ++t) {
    const int64_t cur = add_count % capacity;
    const int64_t last = (cur - 1 + capacity) % capacity;
    if (add_count == 0 || terminal_store[last]) ep = 0;
    if (valid[cur]) { valid[cur] = 0; --nvalid; }
    if (ep >= update_horizon) {
      const int64_t i = ((cur - update_horizon) % capacity + capacity) % capacity;
      if (!valid[i]) { valid[i] = 1; ++nvalid; }
    }
    terminal_store[cur] = terminal_in[t] ? 1 : 0;
    ++add_count;
    ++ep;
    if (terminal_in[t]) {
      const int64_t back = ep < update_horizon ? ep : update_horizon;
      for (int64_t k = 0; k < back; ++k) {
        const int64_t i = ((cur - k) % capacity + capacity) % capacity;
        if (!valid[i]) { valid[i] = 1; ++nvalid; }
      }
    }
  }
  state[0] = add_count; state[1] = ep; state[2] = nvalid;
}
