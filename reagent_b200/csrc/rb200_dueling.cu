// reagent_b200 -- dueling head (reagent/models/dueling_q_network.py:92-103) folded into a Linear.
//
//   q = value + (advantage - mean_a advantage),   value = w_v.h_v + b_v,  adv = W_a.h_a + b_a
// is linear in the concatenated head activations h = [h_a | h_v] (H each), so a dueling network
// is a plain MLP whose last layer is
//   W_q[a, j]     = W_a[a, j] - mean_a' W_a[a', j]      (j <  H)
//   W_q[a, H + j] = w_v[j]                              (j <  H)
//   b_q[a]        = b_a[a] - mean(b_a) + b_v
// and every fused kernel of the path (TD step, forward, weight gradients) runs on it unchanged.
// The TRUE parameters stay (W_a, b_a, w_v, b_v): `fold` rebuilds W_q / b_q from them before a
// step, `unfold` maps the gradient of the folded layer back with the transposed linear map
//   dW_a[a, j] = dW_q[a, j] - mean_a' dW_q[a', j],  dw_v[j] = sum_a dW_q[a, H + j],
//   db_a[a]    = db_q[a]    - mean(db_q),           db_v    = sum_a db_q[a]
// (per gradient partial slab) and clears the folded layer's gradient so that the fused Adam
// kernel, which also sweeps the derived region of the arena, leaves it untouched (g = 0, m = v = 0).
#include "rb200_common.cuh"

namespace rb200 {

__global__ void dueling_fold_kernel(const float* __restrict__ Wa, const float* __restrict__ ba,
                                    const float* __restrict__ wv, const float* __restrict__ bv,
                                    int A, int H, float* __restrict__ Wq, float* __restrict__ bq) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += Wa[(size_t)a * H + j];
    const float mean = s / (float)A;
    const float v = wv[j];
    for (int a = 0; a < A; ++a) {
      Wq[(size_t)a * 2 * H + j] = Wa[(size_t)a * H + j] - mean;
      Wq[(size_t)a * 2 * H + H + j] = v;
    }
  }
  if (j == 0) {
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += ba[a];
    const float mean = s / (float)A;
    for (int a = 0; a < A; ++a) bq[a] = ba[a] - mean + bv[0];
  }
}

__global__ void dueling_unfold_kernel(float* __restrict__ g, long long slab_stride, int A, int H,
                                      long long o_wq, long long o_bq, long long o_wa,
                                      long long o_ba, long long o_wv, long long o_bv) {
  float* s = g + (size_t)blockIdx.y * slab_stride;
  float* gq = s + o_wq;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    float sa = 0.f, sv = 0.f;
    for (int a = 0; a < A; ++a) {
      sa += gq[(size_t)a * 2 * H + j];
      sv += gq[(size_t)a * 2 * H + H + j];
    }
    const float mean = sa / (float)A;
    for (int a = 0; a < A; ++a) {
      s[o_wa + (size_t)a * H + j] = gq[(size_t)a * 2 * H + j] - mean;
      gq[(size_t)a * 2 * H + j] = 0.f;
      gq[(size_t)a * 2 * H + H + j] = 0.f;
    }
    s[o_wv + j] = sv;
  }
  if (j == 0) {
    float sb = 0.f;
    for (int a = 0; a < A; ++a) sb += s[o_bq + a];
    const float mean = sb / (float)A;
    for (int a = 0; a < A; ++a) {
      s[o_ba + a] = s[o_bq + a] - mean;
      s[o_bq + a] = 0.f;
    }
    s[o_bv] = sb;
  }
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_dueling_fold(const float* W_adv, const float* b_adv, const float* w_val,
                                  const float* b_val, int32_t num_actions, int32_t head_hidden,
                                  float* W_q, float* b_q, void* stream) {
  if (!W_adv || !b_adv || !w_val || !b_val || !W_q || !b_q || num_actions <= 0 || head_hidden <= 0) {
    set_last_error("rb200_dueling_fold: bad argument");
    return RB200_E_INVALID;
  }
  dueling_fold_kernel<<<ceil_div(head_hidden, 128), 128, 0, (cudaStream_t)stream>>>(
      W_adv, b_adv, w_val, b_val, num_actions, head_hidden, W_q, b_q);
  return check_cuda(cudaGetLastError(), "dueling_fold_kernel launch");
}

extern "C" int rb200_dueling_unfold(float* grad, int64_t slab_stride, int32_t splits,
                                    int32_t num_actions, int32_t head_hidden, int64_t off_W_q,
                                    int64_t off_b_q, int64_t off_W_adv, int64_t off_b_adv,
                                    int64_t off_w_val, int64_t off_b_val, void* stream) {
  if (!grad || splits <= 0 || num_actions <= 0 || head_hidden <= 0) {
    set_last_error("rb200_dueling_unfold: bad argument");
    return RB200_E_INVALID;
  }
  dim3 grid(ceil_div(head_hidden, 128), splits);
  dueling_unfold_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
      grad, slab_stride, num_actions, head_hidden, off_W_q, off_b_q, off_W_adv, off_b_adv,
      off_w_val, off_b_val);
  return check_cuda(cudaGetLastError(), "dueling_unfold_kernel launch");
}
