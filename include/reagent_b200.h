/* reagent_b200 -- C ABI of the B200-native off-policy training hot path.
 *
 * The reference (facebookresearch/ReAgent) has no FFI for this path: it is plain
 * Python over torch.  This header is the boundary one level
 * beneath the Python surface: every entry point replaces a chain of eager aten ops
 * in the reference file:line it cites.  All pointers are DEVICE pointers unless the
 * parameter name ends in `_host`; every function takes the CUDA stream to launch on
 * (as void*), owns no memory, starts no threads and returns 0 on success or a
 * negative RB200_E_* code (text via rb200_last_error()).
 */
#ifndef REAGENT_B200_H_
#define REAGENT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_VERSION 1
#define RB200_MAX_LAYERS 8

#define RB200_OK 0
#define RB200_E_INVALID (-1)   /* bad argument / unsupported shape            */
#define RB200_E_CUDA (-2)      /* CUDA runtime error                          */
#define RB200_E_SMEM (-3)      /* tile does not fit in 227 KB shared memory   */

/* activations: reagent/models/fully_connected_network.py:37-44 */
#define RB200_ACT_LINEAR 0
#define RB200_ACT_RELU 1
#define RB200_ACT_TANH 2
#define RB200_ACT_LEAKY_RELU 3
#define RB200_ACT_SIGMOID 4
#define RB200_ACT_SOFTPLUS 5

/* losses: reagent/training/dqn_trainer_base.py:146-155 */
#define RB200_LOSS_MSE 0
#define RB200_LOSS_HUBER 1

/* discount modes: reagent/training/dqn_trainer.py:166-177 */
#define RB200_DISCOUNT_CONST 0     /* gamma                          */
#define RB200_DISCOUNT_POW 1       /* gamma ** discount_src[b]       */

/* One fully connected network (reagent/models/fully_connected_network.py:67-163,
 * in-scope subset: Linear + activation, no BN/LN/dropout/residual).  Parameters live
 * in one flat fp32 arena; layer l's weight is [dims[l+1], dims[l]] row-major at
 * params + w_off[l] (== nn.Linear.weight), its bias at params + b_off[l]. */
typedef struct rb200_mlp {
  int32_t n_layers;
  int32_t dims[RB200_MAX_LAYERS + 1];
  int32_t act[RB200_MAX_LAYERS];
  const float* params;
  int64_t w_off[RB200_MAX_LAYERS];
  int64_t b_off[RB200_MAX_LAYERS];
  int64_t n_params; /* arena length in floats (including alignment padding) */
} rb200_mlp_t;

/* Per-network training workspace (caller allocated, all dense row-major):
 *   hidden[l]  [B, dims[l+1]]  output of layer l, l < n_layers-1
 *   dz[l]      [B, dims[l+1]]  dLoss/d(pre-activation of layer l)
 *   input      [B, dims[0]]    the (concatenated) network input, or NULL when the
 *                              batch tensor itself is the input               */
typedef struct rb200_net_ws {
  float* hidden[RB200_MAX_LAYERS];
  float* dz[RB200_MAX_LAYERS];
  float* input;
} rb200_net_ws_t;

const char* rb200_last_error(void);
int rb200_version(void);
/* number of SMs / max opt-in smem of the current device (host query helpers) */
int rb200_device_info(int* sm_count, int* max_smem_optin);

/* Fused whole-MLP forward out = net(cat(in0, in1)) over row tiles; in1 may be NULL.
 * Replaces FullyConnectedNetwork.forward (reagent/models/fully_connected_network.py:157-163)
 * and FullyConnectedCritic.forward's cat (reagent/models/critic.py:76-92). */
int rb200_mlp_forward(const rb200_mlp_t* net, const float* in0, int32_t d0, const float* in1,
                      int32_t d1, int32_t batch, float* out, void* stream);

/* ------------------------------------------------------------------------- */
/* K2 (+K2'): fused DQN TD-target / loss / backward over row tiles.            */
/* Replaces DQNTrainer.compute_td_loss + get_max_q_values_with_target +        */
/* boost_rewards + compute_discount_tensor (reagent/training/dqn_trainer.py:   */
/* 157-239, dqn_trainer_base.py:33-77,216-241) and, with do_backward, autograd's*/
/* backward through q_network down to every layer's pre-activation gradient.    */
/* ------------------------------------------------------------------------- */
typedef struct rb200_dqn_args {
  int32_t batch;                 /* B */
  const float* state;            /* [B,S] */
  const float* next_state;       /* [B,S] */
  const float* action;           /* [B,A] float weights (one-hot in practice) */
  const float* next_action;      /* [B,A] (SARSA only, may be NULL)          */
  const float* reward;           /* [B]   */
  const float* not_terminal;     /* [B]   */
  const float* possible_next_actions_mask; /* [B,A] or NULL (= ones)          */
  const float* discount_src;     /* [B] time_diff or step (POW mode) or NULL  */
  const float* reward_boost;     /* [A] or NULL                                */
  float gamma;
  int32_t discount_mode;         /* RB200_DISCOUNT_*  */
  int32_t double_q;              /* dqn_trainer_base.py:64-75 */
  int32_t maxq;                  /* 0 = SARSA (mask := next_action)           */
  int32_t loss_kind;             /* RB200_LOSS_*      */
  int32_t do_backward;           /* 0: forward/loss only (validation_step)    */
  /* outputs */
  float* all_action_scores;      /* [B,A] q_network(state) (detached) or NULL */
  float* td_target;              /* [B] or NULL                                */
  float* q_selected;             /* [B] or NULL                                */
  int32_t* next_action_idx;      /* [B] argmax index or NULL                   */
  float* loss_partials;          /* [>= rb200_dqn_num_tiles(B)]                */
  float* loss;                   /* [1] mean loss (written by the last tile)   */
  uint32_t* tile_counter;        /* [1] zero-initialised scratch, self-resetting */
} rb200_dqn_args_t;

int rb200_num_row_tiles(int batch, int max_dim_in, int max_dim_hidden);
int rb200_dqn_td_step(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target,
                      const rb200_dqn_args_t* args, const rb200_net_ws_t* ws, void* stream);

/* ------------------------------------------------------------------------- */
/* Weight gradients: dW_l = dZ_l^T . A_{l-1}, db_l = sum_b dZ_l, split over    */
/* the batch; partial s lands at gpart + s*n_params (arena layout).            */
/* Replaces autograd's Linear backward (torch) reached from                    */
/* loss.backward() in the Lightning loop (reagent_lightning_module.py:108-133).*/
/* ------------------------------------------------------------------------- */
int rb200_wgrad_splits(int batch);
int rb200_mlp_wgrad(const rb200_mlp_t* net, const float* net_input, int32_t batch,
                    const rb200_net_ws_t* ws, float* gpart, int32_t splits, void* stream);
/* g[i] = sum_s gpart[s*P + i]  (fixed order; feeds all-reduce / .grad views) */
int rb200_grad_reduce(const float* gpart, int32_t splits, int64_t n, float* g, void* stream);

/* ------------------------------------------------------------------------- */
/* K3: fused Adam + soft target update over flat arenas.                        */
/* Replaces torch.optim.Adam.step (reagent/optimizer/optimizer.py:64-85,        */
/* uninferrable_optimizers.py:23-33) and SoftUpdate.step                        */
/* (reagent/optimizer/soft_update.py:47-71) in this order per element:          */
/* Adam on the source, then target = tau*new_source + (1-tau)*target.           */
/* `step` is a device int64 counter incremented by the kernel (graph friendly). */
/* ------------------------------------------------------------------------- */
typedef struct rb200_adam_args {
  float* params;          /* [n] */
  const float* grad;      /* [splits, n] partials, summed in order */
  int32_t splits;
  int64_t n;
  float* exp_avg;         /* [n] */
  float* exp_avg_sq;      /* [n] */
  int64_t* step;          /* [1] device */
  uint32_t* block_counter;/* [1] zero-initialised scratch, self-resetting */
  double lr, beta1, beta2, eps, weight_decay;
  float grad_scale;       /* multiplies the summed gradient (1/world for DP) */
  float* target;          /* [n] or NULL: fused Polyak update */
  float tau;
  float one_minus_tau;    /* float(1.0 - tau) computed in double on the host */
} rb200_adam_args_t;
int rb200_adam_soft_update(const rb200_adam_args_t* a, void* stream);
/* stand-alone Polyak update (SoftUpdate.step when not fused) */
int rb200_soft_update(float* target, const float* source, int64_t n, float tau,
                      float one_minus_tau, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REAGENT_B200_H_ */
