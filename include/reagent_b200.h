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

/* feature types: reagent/preprocessing/identify_types.py:9-30 (FEATURE_TYPES order) */
#define RB200_FT_BINARY 0
#define RB200_FT_PROBABILITY 1
#define RB200_FT_CONTINUOUS 2
#define RB200_FT_BOXCOX 3
#define RB200_FT_ENUM 4
#define RB200_FT_QUANTILE 5
#define RB200_FT_CONTINUOUS_ACTION 6
#define RB200_FT_DISCRETE_ACTION 7
#define RB200_FT_DO_NOT_PREPROCESS 8
#define RB200_FT_CLIP_LOG 9

/* One OUTPUT column of the dense preprocessor (reagent/preprocessing/preprocessor.py):
 * out[:, j] = transform_type(in[:, src_col]; p0..p3 [, quantiles q_off..q_off+q_cnt)).
 *   PROBABILITY: p0=1e-5 p1=1-1e-5 (as f32)     CONTINUOUS: p0=mean p1=stddev
 *   BOXCOX: p0=mean p1=stddev p2=shift p3=lambda ENUM: p0=possible value of this column
 *   QUANTILE: p0=len(quantiles)-1 p1=max p2=min, q_* = padded boundaries
 *   CONTINUOUS_ACTION: p0=min_serving p1=scaling_factor p2=min_training p3=-1+EPS */
typedef struct rb200_feature_col {
  int32_t src_col;
  int32_t type;
  float p0, p1, p2, p3;
  int32_t q_off, q_cnt;
} rb200_feature_col_t;

const char* rb200_last_error(void);
int rb200_version(void);
/* sizeof() of an ABI struct by name ("rb200_adam_args_t", ...), -1 if unknown: lets a binding
 * verify its mirror of the struct against the library it loaded */
int64_t rb200_abi_sizeof(const char* type_name);
/* number of SMs / max opt-in smem of the current device (host query helpers) */
int rb200_device_info(int* sm_count, int* max_smem_optin);

/* P1: Preprocessor.forward(input, input_presence_byte) -- reagent/preprocessing/
 * preprocessor.py:115-170.  `cols` / `quantiles` are device arrays; presence may be NULL
 * (all present), uint8/bool [rows,f_in] or float [rows,f_in]. */
int rb200_preprocess(const float* input, const void* presence, int32_t presence_is_float,
                     int64_t rows, int32_t f_in, int32_t f_out, const rb200_feature_col_t* cols,
                     const float* quantiles, float* out, void* stream);

/* Fused whole-MLP forward out = net(cat(in0, in1)) over row tiles; in1 may be NULL.
 * Replaces FullyConnectedNetwork.forward (reagent/models/fully_connected_network.py:157-163)
 * and FullyConnectedCritic.forward's cat (reagent/models/critic.py:76-92).  When
 * save_hidden != NULL the hidden layer outputs go to save_hidden->hidden[l] (training). */
int rb200_mlp_forward(const rb200_mlp_t* net, const float* in0, int32_t d0, const float* in1,
                      int32_t d1, int32_t batch, float* out, const rb200_net_ws_t* save_hidden,
                      void* stream);

/* ------------------------------------------------------------------------- */
/* Dueling head folded into a Linear (rb200_dueling.cu).  Replaces the head arithmetic of     */
/* DuelingQNetwork._get_values (reagent/models/dueling_q_network.py:92-103), with atoms:       */
/*   q[a,n] = value[n] + advantage[a,n] - mean_{a',n'}(advantage)     (num_atoms = 1: DQN)     */
/* `fold` builds the equivalent last layer W_q [A*N, 2H] / b_q [A*N] (row a*N+n) from the true  */
/* parameters (advantage head W_adv [A*N,H], b_adv [A*N]; value head w_val [N,H], b_val [N]) so */
/* that every MLP kernel of this library runs a dueling network as a plain MLP; `unfold` maps   */
/* the gradient of that layer back onto the true parameters in each of `splits` gradient slabs  */
/* (offsets in floats from the slab start) and zeroes the folded layer's gradient.  `scratch`:  */
/* rb200_dueling_scratch_floats(H, splits) floats of device memory.                             */
/* ------------------------------------------------------------------------- */
int64_t rb200_dueling_scratch_floats(int32_t head_hidden, int32_t splits);
int rb200_dueling_fold(const float* W_adv, const float* b_adv, const float* w_val,
                       const float* b_val, int32_t num_actions, int32_t num_atoms,
                       int32_t head_hidden, float* W_q, float* b_q, float* scratch, void* stream);
int rb200_dueling_unfold(float* grad, int64_t slab_stride, int32_t splits, int32_t num_actions,
                         int32_t num_atoms, int32_t head_hidden, int64_t off_W_q, int64_t off_b_q,
                         int64_t off_W_adv, int64_t off_b_adv, int64_t off_w_val,
                         int64_t off_b_val, float* scratch, void* stream);

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
  float* loss_partials;          /* [>= ceil(B/16)]                */
  float* loss;                   /* [1] mean loss (written by the last tile)   */
  uint32_t* tile_counter;        /* [1] zero-initialised scratch, self-resetting */
} rb200_dqn_args_t;

int rb200_num_row_tiles(int batch, int max_dim_in, int max_dim_hidden);
int rb200_dqn_td_step(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target,
                      const rb200_dqn_args_t* args, const rb200_net_ws_t* ws, void* stream);

/* The same step on the 5th-generation tensor cores (tcgen05.mma kind::tf32 with 3xTF32 error
 * compensation, accumulators in Tensor Memory, weights streamed by bulk async copies):
 * rb200_dqn_tc.cu.  Same arguments, semantics and reference lines as rb200_dqn_td_step plus a
 * caller-owned scratch buffer for the packed hi/lo weight images (re-packed on every call,
 * because the weights change on every update).
 *   rb200_dqn_tc_workspace_bytes: bytes the scratch buffer needs for this network, or 0 when
 *       the shapes do not fit the tensor-core path (then use rb200_dqn_td_step).
 *   rb200_dqn_tc_pack: (re)build the hi/lo weight images from the current parameters.  It only
 *       depends on the parameters, so a caller may run it on a side stream as soon as the
 *       previous optimizer step is done (e.g. concurrently with replay sampling) and pass
 *       weights_packed = 1; with weights_packed = 0 the step packs first, on `stream`.
 *       The images depend on (double_q, do_backward) only through which ones are built.
 *   pack_ws: device buffer, 128-byte aligned, zero-initialised ONCE by the caller. */
int64_t rb200_dqn_tc_workspace_bytes(const rb200_mlp_t* q_net, int32_t double_q,
                                     int32_t do_backward);
int rb200_dqn_tc_pack(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target, int32_t double_q,
                      int32_t do_backward, void* pack_ws, int64_t pack_ws_bytes, void* stream);
int rb200_dqn_td_step_tc(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target,
                         const rb200_dqn_args_t* args, const rb200_net_ws_t* ws, void* pack_ws,
                         int64_t pack_ws_bytes, int32_t weights_packed, void* stream);

/* ------------------------------------------------------------------------- */
/* CPE heads of the DQN step: DQNTrainerBaseLightning._calculate_cpes           */
/* (reagent/training/dqn_trainer_base.py:332-452), masked_softmax                 */
/* (reagent/core/torch_utils.py:62-73).  The reward network and the CPE q-network  */
/* are plain MLPs (rb200_mlp_forward / rb200_mlp_backward / rb200_mlp_wgrad); this  */
/* entry computes both losses and d loss / d output of both networks.              */
/* loss[0] = reward loss, loss[1] = CPE q-value loss.                               */
/* ------------------------------------------------------------------------- */
typedef struct rb200_cpe_args {
  int32_t batch, num_actions, num_metrics;   /* B, A, M = len(metrics_to_score) */
  const float* next_scores;        /* [B,A] q_network(next_state) (detached) */
  const float* mask;               /* [B,A] possible_next_actions_mask (maxq) | next_action, or NULL */
  float temperature;               /* rl.temperature */
  const float* action;             /* [B,A] logged action (argmax = index) */
  const float* metrics_reward;     /* [B,M] cat(reward, extras.metrics) */
  const float* discount_src;       /* [B] or NULL */
  float gamma;
  int32_t discount_mode;           /* RB200_DISCOUNT_* */
  const float* not_terminal;       /* [B] */
  const float* reward_est;         /* [B, M*A] reward_network(state) */
  const float* qcpe;               /* [B, M*A] q_network_cpe(state) */
  const float* qcpe_target_next;   /* [B, M*A] q_network_cpe_target(next_state) */
  int32_t loss_kind;               /* RB200_LOSS_* (rl.q_network_loss) for the CPE q-network */
  float* dz_reward;                /* [B, M*A] */
  float* dz_qcpe;                  /* [B, M*A] */
  float* propensities_next;        /* [B,A] or NULL */
  float* loss_partials;            /* [2 * ceil(B/256)] */
  float* loss;                     /* [2] */
  uint32_t* tile_counter;          /* [1] zero-initialised, self-resetting */
} rb200_cpe_args_t;
int rb200_cpe_heads(const rb200_cpe_args_t* args, void* stream);

/* ------------------------------------------------------------------------- */
/* Loss heads of ParametricDQNTrainer and C51Trainer (rb200_heads.cu); the networks */
/* around them run on rb200_mlp_forward / rb200_linear_forward / *_backward / wgrad. */
/*   rb200_pdqn_head: reagent/training/parametric_dqn_trainer.py:109-173 (TD target    */
/*     from the tiled possible next actions or the SARSA value, mse | huber, dL/dq)     */
/*   rb200_c51_head:  reagent/training/c51_trainer.py:98-173 (log-softmax over atoms,  */
/*     masked arg max of expected values, categorical projection, cross entropy,       */
/*     dL/dlogits); reagent/models/categorical_dqn.py:28-35                             */
/* ------------------------------------------------------------------------- */
typedef struct rb200_pdqn_args {
  int32_t batch, max_num_action;   /* M tiled next actions per row; 0 = SARSA */
  const float* next_q;             /* [B*M] q_network(tiled s', a') (double-Q) or NULL */
  const float* next_q_target;      /* [B*M] (maxq) or [B] (SARSA: q_target(s', next_action)) */
  const float* mask;               /* [B,M] possible_next_actions_mask or NULL */
  const float* reward;             /* [B] */
  const float* not_terminal;       /* [B] */
  const float* discount_src;       /* [B] or NULL */
  float gamma;
  int32_t discount_mode, double_q, loss_kind;
  const float* q_values;           /* [B] q_network(state, action) */
  float* dz;                       /* [B] d loss / d q */
  float* td_target;                /* [B] or NULL */
  float* loss_partials;            /* [ceil(B/256)] */
  float* loss;                     /* [1] */
  uint32_t* tile_counter;
} rb200_pdqn_args_t;
int rb200_pdqn_head(const rb200_pdqn_args_t* args, void* stream);

typedef struct rb200_c51_args {
  int32_t batch, num_actions, num_atoms;
  const float* logits_next_online; /* [B, A*N] distributional_network(next_state), online (double-Q) or NULL */
  const float* logits_next_target; /* [B, A*N] target network */
  const float* logits_cur;         /* [B, A*N] online network on state */
  const float* action;             /* [B,A] */
  const float* next_action;        /* [B,A] (SARSA) or NULL */
  const float* possible_next_actions_mask; /* [B,A] or NULL */
  const float* reward;             /* [B] */
  const float* not_terminal;       /* [B] */
  const float* discount_src;       /* [B] or NULL: gamma ** discount_src */
  const float* reward_boost;       /* [A] or NULL */
  const float* support;            /* [N] torch.linspace(qmin, qmax, N) */
  float gamma, qmin, qmax, scale_support;
  int32_t double_q, maxq;
  float* dz_logits;                /* [B, A*N] */
  float* all_q_values;             /* [B,A] or NULL */
  int32_t* next_action_idx;        /* [B] or NULL */
  float* loss_partials;            /* [B] */
  float* loss;                     /* [1] */
  uint32_t* tile_counter;
} rb200_c51_args_t;
int rb200_c51_head(const rb200_c51_args_t* args, void* stream);

/* ------------------------------------------------------------------------- */
/* QR-DQN (reagent/training/qrdqn_trainer.py:108-194).  The [hidden -> A*N] head  */
/* is too wide for a row tile, so it runs as 2-D tiled launches:                   */
/*   rb200_linear_forward      out = act(in . W^T + b), any N     (nn.Linear fwd) */
/*   rb200_linear_backward_dx  dz_prev = (dz . W) * act'(h_prev)  (autograd)      */
/*   rb200_mlp_backward        dZ chain of the layers below a given last-layer dz */
/*   rb200_qrdqn_head          mean over atoms, masked argmax, target             */
/*       distribution, pairwise quantile-Huber loss (:152-155, :217-218) and      */
/*       d loss / d head output, one CTA per row, nothing (N,B,N)-sized in HBM.   */
/* ------------------------------------------------------------------------- */
typedef struct rb200_qrdqn_args {
  int32_t batch, num_actions, num_atoms;
  const float* q_next_online;  /* [B, A*N] q_network(next_state)  (double-Q) or NULL */
  const float* q_next_target;  /* [B, A*N] q_network_target(next_state) */
  const float* q_cur;          /* [B, A*N] q_network(state) */
  const float* action;         /* [B, A] */
  const float* next_action;    /* [B, A] (SARSA) or NULL */
  const float* possible_next_actions_mask; /* [B, A] or NULL */
  const float* reward;         /* [B] */
  const float* not_terminal;   /* [B] */
  const float* discount_src;   /* [B] or NULL: gamma ** discount_src */
  const float* reward_boost;   /* [A] or NULL */
  float gamma;
  int32_t double_q, maxq;
  float* dz_head;              /* [B, A*N] d loss / d head output */
  float* all_q_values;         /* [B, A] mean over atoms of q(s), or NULL */
  int32_t* next_action_idx;    /* [B] or NULL */
  float* loss_partials;        /* [B] */
  float* loss;                 /* [1] */
  uint32_t* tile_counter;      /* [1] zero-initialised, self-resetting */
} rb200_qrdqn_args_t;

int rb200_linear_forward(const float* W, const float* b, int32_t act, int32_t K, int32_t N,
                         const float* in, int32_t batch, float* out, void* stream);
/* tcgen05.mma (kind::tf32, 3xTF32) + TMEM implementation of rb200_linear_forward, taken
 * automatically for batch >= 128 and N >= 128 */
int rb200_linear_forward_tc(const float* W, const float* b, int32_t act, int32_t K, int32_t N,
                            const float* in, int32_t batch, float* out, void* stream);
int rb200_linear_backward_dx(const float* W, int32_t K, int32_t N, const float* dz,
                             const float* h_prev, int32_t act_prev, int32_t batch, float* out,
                             void* stream);
/* tcgen05 implementation of rb200_linear_backward_dx for a wide layer (the contraction runs
 * over the N out-features: split-K slices of the tensor-core GEMM, added in a fixed order).
 * _scratch_bytes returns 0 when the shape is not taken by this path (N < 1024, batch < 256 or
 * K / N not multiples of 4): call rb200_linear_backward_dx then.  Replaces the same autograd
 * step (torch.nn.functional.linear backward w.r.t. input, reagent/training/qrdqn_trainer.py:
 * 108-194 via reagent_lightning_module.py:108-133). */
int64_t rb200_linear_backward_dx_tc_scratch_bytes(int32_t K, int32_t N, int32_t batch);
int rb200_linear_backward_dx_tc(const float* W, int32_t K, int32_t N, const float* dz,
                                const float* h_prev, int32_t act_prev, int32_t batch, float* out,
                                void* scratch, int64_t scratch_bytes, void* stream);
int rb200_mlp_backward(const rb200_mlp_t* net, const float* dz_last, int32_t batch,
                       const rb200_net_ws_t* ws, void* stream);
int rb200_qrdqn_head(const rb200_qrdqn_args_t* args, void* stream);

/* ------------------------------------------------------------------------- */
/* Fused SAC / TD3 updates over row tiles.                                      */
/* critic step: replaces SACTrainer.train_step_gen's first section              */
/*   (reagent/training/sac_trainer.py:214-248: actor(s') + get_log_prob, target */
/*   critics, min, entropy term, target, q1/q2 MSE) or TD3Trainer's             */
/*   (reagent/training/td3_trainer.py:138-178), plus autograd's backward through*/
/*   q1 / q2.  loss[0] = q1 loss, loss[1] = q2 loss.                            */
/* actor step: replaces sac_trainer.py:254-322 (actor loss, alpha loss) or      */
/*   td3_trainer.py:181-187, plus the backward through the frozen critics into  */
/*   the actor (reagent/models/actor.py:169-261 for the Gaussian head).         */
/*   loss[0] = actor loss, loss[1] = alpha loss; alpha_grad[0] = d/d log_alpha. */
/* `noise_*` are the N(0,1) draws of torch.randn_like in the reference          */
/* (actor.py:217, td3_trainer.py:141), supplied by the caller.                  */
/* ------------------------------------------------------------------------- */
#define RB200_ALGO_SAC 0
#define RB200_ALGO_TD3 1
typedef struct rb200_ac_args {
  int32_t batch;
  int32_t algo;
  const float* state;        /* [B,S] */
  const float* action;       /* [B,A] (critic step) */
  const float* next_state;   /* [B,S] (critic step) */
  const float* reward;       /* [B] */
  const float* not_terminal; /* [B] */
  const float* noise_next;   /* [B,A] critic step */
  const float* noise_cur;    /* [B,A] SAC actor step */
  float gamma;
  const float* alpha;        /* [1] SAC entropy temperature (device) */
  const float* log_alpha;    /* [1] SAC (actor step: alpha loss value) or NULL */
  float target_entropy;
  int32_t backprop_through_log_prob;
  float noise_variance, noise_clip; /* TD3 */
  /* outputs */
  float* loss_partials;      /* [2 * num tiles] */
  float* loss;               /* [2] */
  uint32_t* tile_counter;    /* [1] zero-initialised, self-resetting */
  float* alpha_grad;         /* [1] or NULL */
  float* td_target;          /* [B] or NULL */
  float* next_action_out;    /* [B,A] or NULL: a' (critic step) / pi(s) (actor step) */
  float* log_prob_out;       /* [B] or NULL (unclamped sum of log-probs) */
  float* q1_value;           /* [B] or NULL */
  float* q2_value;           /* [B] or NULL */
} rb200_ac_args_t;

int rb200_ac_critic_step(const rb200_mlp_t* actor, const rb200_mlp_t* q1, const rb200_mlp_t* q2,
                         const rb200_mlp_t* q1_target, const rb200_mlp_t* q2_target,
                         const rb200_ac_args_t* args, const rb200_net_ws_t* ws_q1,
                         const rb200_net_ws_t* ws_q2, void* stream);
int rb200_ac_actor_step(const rb200_mlp_t* actor, const rb200_mlp_t* q1, const rb200_mlp_t* q2,
                        const rb200_ac_args_t* args, const rb200_net_ws_t* ws_actor,
                        const rb200_net_ws_t* ws_q1, const rb200_net_ws_t* ws_q2, void* stream);

/* ------------------------------------------------------------------------- */
/* Weight gradients: dW_l = dZ_l^T . A_{l-1}, db_l = sum_b dZ_l, split over    */
/* the batch; partial s lands at gpart + s*n_params (arena layout).            */
/* Default: mma.sync 3xTF32 tiles (rb200_optim.cu).  RB200_WGRAD_TC=1 selects the  */
/* tcgen05 kernel (rb200_wgrad_tc.cu: operands transposed into K-major planes while */
/* staging, accumulator in Tensor Memory); same results to 1e-5, same speed today.  */
/* Replaces autograd's Linear backward (torch) reached from                    */
/* loss.backward() in the Lightning loop (reagent_lightning_module.py:108-133).*/
/* ------------------------------------------------------------------------- */
int rb200_wgrad_splits(int batch);
/* slabs for this network (tcgen05 kernel: enough (tile, slab) jobs to fill the SMs twice) */
int rb200_wgrad_splits_for(const rb200_mlp_t* net, int32_t batch);
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
  float* exp_out;         /* [n] or NULL: exp(new param) (SAC: entropy_temperature =
                             log_alpha.exp(), sac_trainer.py:322) */
  /* Optional: also write the hi/lo tensor-core weight images of the UPDATED parameters (and of
   * the updated target) for the next rb200_dqn_td_step_tc, which can then be called with
   * weights_packed = 1 -- the work of rb200_dqn_tc_pack without its launch.  tc_net describes
   * the network whose arena `params` is (tc_net->params == params); needs `target`. */
  const rb200_mlp_t* tc_net;   /* NULL: no packing */
  void* tc_pack_ws;            /* the scratch buffer of rb200_dqn_td_step_tc */
  int64_t tc_pack_ws_bytes;
  int32_t tc_do_backward;      /* also the transposed images of the backward */
  /* Optional: data-parallel gradient exchange FUSED into this launch (dp_world > 1), replacing
   * rb200_grad_reduce + an NCCL all-reduce + this kernel by one kernel per rank.  Every rank
   * launches the same grid; block b of every rank (1) sums its slice of its own split-K
   * partials, (2) PUSHES the slice into every peer's receive buffer with peer-to-peer stores
   * over NVLink and raises a per-block flag there, (3) waits for the W-1 flags of its own
   * slice, (4) adds the W slices in rank order -- every rank computes the bit-identical global
   * gradient -- scales by grad_scale (1/W) and runs Adam / Polyak / packing as above.
   * dp_recv[r] / dp_flags[r] are DEVICE arrays of W peer-mapped pointers (rb200_dp_ipc_open):
   *   recv  of rank r: float    [2][W][dp_stride]      (parity of the step, source rank)
   *   flags of rank r: uint32_t [2][W][dp_max_blocks]  zero-initialised once
   * Two parities suffice: a rank cannot finish step k+1 before every peer has finished step k. */
  int32_t dp_world, dp_rank;
  float* const* dp_recv;
  uint32_t* const* dp_flags;
  int64_t dp_stride;           /* >= n */
  int32_t dp_max_blocks;       /* >= the grid this call launches (rb200_adam_blocks(n)) */
} rb200_adam_args_t;
int rb200_adam_blocks(int64_t n);
int rb200_adam_soft_update(const rb200_adam_args_t* a, void* stream);
/* stand-alone Polyak update (SoftUpdate.step when not fused) */
int rb200_soft_update(float* target, const float* source, int64_t n, float tau,
                      float one_minus_tau, void* stream);


/* ------------------------------------------------------------------------- */
/* K1: fused replay sampling over device-resident storage.                     */
/* Replaces SumTree.stratified_sample/sample (reagent/replay_memory/sum_tree.py:93-153),*/
/* ReplayBuffer.sample_index_batch / sample_transition_batch                   */
/* (circular_replay_buffer.py:589-706, :741-774), PrioritizedReplayBuffer's    */
/* sampling_probabilities (prioritized_replay_buffer.py:116-147), the dense    */
/* Preprocessor on state/next_state (preprocessing/preprocessor.py:115-170) and*/
/* the InputMakers (gym/preprocessors/trainer_preprocessor.py:72-227).         */
/* Random numbers come from the HOST (bit-exact index parity): `query` are the */
/* stratified uniforms of SumTree.stratified_sample, `ranks` torch.randint's   */
/* draws; rare invalid-index retries are resolved on the host and passed as    */
/* overrides (see reagent_b200/replay_memory/prioritized_replay_buffer.py).    */
/* ------------------------------------------------------------------------- */
#define RB200_SAMPLE_PRIORITIZED 0
#define RB200_SAMPLE_UNIFORM 1
#define RB200_SAMPLE_GIVEN 2
#define RB200_VALID_BLOCK 256
#define RB200_MAX_GATHER_SPECS 12

typedef struct rb200_gather_spec {
  const void* src;     /* [capacity, row_bytes] */
  void* dst;           /* [batch, row_bytes]    */
  int32_t row_bytes;
  int32_t which;       /* 0: sampled index, 1: next index */
} rb200_gather_spec_t;

typedef struct rb200_sample_args {
  int32_t batch, capacity, update_horizon, mode;
  int32_t timeline_next;              /* next index = i+1 instead of i+steps */
  /* index sources */
  const double* tree;                 /* fp64 heap: level l at [2^l-1, 2^(l+1)-1) */
  int32_t tree_depth;
  const double* query;                /* [B] in [0,1) */
  const int32_t* override_pos;        /* [n_override] batch positions */
  const int64_t* override_idx;        /* [n_override] replacement indices */
  int32_t n_override;
  const int64_t* ranks;               /* [B] rank among valid slots */
  const uint8_t* valid;               /* [capacity] */
  const int32_t* valid_block_offsets; /* [n_valid_blocks+1] */
  int32_t n_valid_blocks;
  const int64_t* indices_in;          /* [B] */
  /* storage */
  const uint8_t* terminal;            /* [capacity] */
  const float* reward;                /* [capacity] */
  const float* decays;                /* [update_horizon] gamma**k as f32 */
  const float* obs;                   /* [capacity, obs_dim] or NULL */
  int32_t obs_dim, obs_out_dim;
  const rb200_feature_col_t* cols;    /* NULL: raw copy */
  const float* quantiles;
  float* state;                       /* [B, obs_out_dim] */
  float* next_state;
  /* discrete action (int64 scalar per slot) */
  const int64_t* action_i64;
  int32_t num_actions;
  int64_t* action_out_i64;            /* [B] */
  int64_t* next_action_out_i64;       /* [B] */
  float* action_onehot;               /* [B, num_actions] */
  float* next_action_onehot;          /* zeroed on terminal rows */
  /* continuous action (f32 row per slot) */
  const float* action_f32;
  int32_t action_dim;
  float* action_out_raw;              /* [B, action_dim] */
  float* next_action_out_raw;
  float* action_rescaled;             /* [B, action_dim] rescaled to [train_low, train_high] */
  float* next_action_rescaled;        /* zeroed on terminal rows */
  const float* action_low;            /* [action_dim] */
  const float* action_high;
  float train_low, train_high;
  /* scalar outputs, all [B] */
  float* reward_out;
  float* next_reward_out;
  uint8_t* terminal_out;
  float* not_terminal_out;
  int64_t* indices_out;
  int64_t* step_out;
  float* step_f32_out;
  float* sampling_prob_out;
  int32_t n_specs;
  rb200_gather_spec_t specs[RB200_MAX_GATHER_SPECS];
} rb200_sample_args_t;

int rb200_replay_sample(const rb200_sample_args_t* args, void* stream);
/* validity bitmap -> per-256-slot counts and exclusive offsets (uniform sampling index);
 * counts [ceil(cap/256)], offsets [ceil(cap/256)+1] */
int rb200_valid_index_build(const uint8_t* valid, int64_t capacity, int32_t* counts,
                            int32_t* offsets, void* stream);

/* ------------------------------------------------------------------------- */
/* Device-resident replay bookkeeping (rb200_replay_dev.cu): the online loop "add a      */
/* transition -> draw a prioritized minibatch -> train" without host work per step.     */
/*   rb200_replay_add_device  n consecutive ReplayBuffer.add() calls, stack_size == 1    */
/*       (circular_replay_buffer.py:468-547; validity :430-438; PER priority ->          */
/*       SumTree.set, prioritized_replay_buffer.py:62-84) from device staging rows        */
/*   rb200_sumtree_set_device PrioritizedReplayBuffer.set_priority: SumTree.set for a     */
/*       batch applied IN ORDER (sum_tree.py:164-189), fp64, bit-equal to the host heap   */
/*   rb200_per_draw_indices   sample_index_batch of the prioritized buffer                */
/*       (prioritized_replay_buffer.py:86-115): B stratified random.uniform draws from a  */
/*       DEVICE copy of CPython's MT19937 state (624 words + position, as                 */
/*       random.getstate()[1] lays them out), the tree descents, and the sequential       */
/*       non-stratified retries of invalid hits with the shared attempt budget.           */
/* status words are sticky error flags the host wrapper turns into the reference's        */
/* exceptions: 1 = "Max sample attempts", 2 = negative priority.                          */
/* ------------------------------------------------------------------------- */
typedef struct rb200_replay_dev {
  int64_t* state;             /* [4] device: add_count, transitions in the current episode,
                                 number of valid indices, sticky error */
  int32_t capacity, update_horizon;
  uint8_t* valid;             /* [capacity] */
  uint8_t* terminal;          /* [capacity] */
  float* reward;              /* [capacity] */
  double* tree;               /* fp64 heap (level l at [2^l-1, 2^(l+1)-1)) or NULL */
  int32_t tree_depth;
  double* max_priority;       /* [1] SumTree.max_recorded_priority or NULL */
} rb200_replay_dev_t;

typedef struct rb200_add_args {
  rb200_replay_dev_t rb;
  int32_t n;                          /* transitions in this call, <= 1024 */
  const uint8_t* terminal_in;         /* [n] */
  const float* reward_in;             /* [n] */
  const double* priority_in;          /* [n] or NULL */
  int32_t n_rows;                     /* other keys: src = staging [n,row_bytes], dst = store */
  rb200_gather_spec_t rows[RB200_MAX_GATHER_SPECS];
} rb200_add_args_t;

typedef struct rb200_per_draw_args {
  uint32_t* mt_state;         /* [625] device, updated in place */
  int32_t batch;
  const double* lo;           /* [B] np.linspace(0,1,B+1)[:-1] */
  const double* hi;           /* [B] np.linspace(0,1,B+1)[1:]  */
  const double* tree;
  int32_t tree_depth;
  const uint8_t* valid;       /* [capacity] */
  int32_t max_attempts;
  int64_t* indices_out;       /* [B] */
  double* queries_out;        /* [B] or NULL */
  int32_t* status;            /* [2]: sticky error, retries used by this call */
} rb200_per_draw_args_t;

int rb200_replay_add_device(const rb200_add_args_t* args, void* stream);
int rb200_sumtree_set_device(double* tree, int32_t depth, const int64_t* idx, const double* val,
                             int32_t n, double* max_recorded, int32_t* status, void* stream);
int rb200_per_draw_indices(const rb200_per_draw_args_t* args, void* stream);

/* ------------------------------------------------------------------------- */
/* Peer-memory plumbing of the fused data-parallel step (one process per GPU).  The    */
/* reference has no collective on this path (docs/distributed.rst:12-22 states the     */
/* intent: synchronous data parallelism with a gradient all-reduce).                   */
/*   rb200_dp_alloc      cudaMalloc'ed, zero-filled, IPC-exportable device buffer      */
/*   rb200_dp_ipc_handle 64-byte handle of such a buffer (cudaIpcGetMemHandle)          */
/*   rb200_dp_ipc_open   map a peer process's buffer into this one (enables peer access)*/
/* ------------------------------------------------------------------------- */
#define RB200_IPC_HANDLE_BYTES 64
int rb200_dp_alloc(int64_t bytes, void** out_ptr);
int rb200_dp_free(void* ptr);
int rb200_dp_ipc_handle(void* ptr, unsigned char* handle_out_host);
int rb200_dp_ipc_open(const unsigned char* handle_host, void** out_ptr);
int rb200_dp_ipc_close(void* ptr);

/* ---- host-side helpers (plain C, no CUDA; pointers are HOST memory) -------- */
/* MT19937 with CPython's exact stream: fills out[i] = lo[i] + (hi[i]-lo[i])*random()
 * (random.uniform, Lib/random.py) or random() itself when lo/hi are NULL.
 * state624 + *index are CPython's random.getstate()[1] (624 words + position). */
void rb200_mt19937_uniform_host(uint32_t* state624, int32_t* index, const double* lo_host,
                                const double* hi_host, double* out_host, int64_t n);
/* SumTree.set for a batch, sequentially, on a host fp64 heap (sum_tree.py:164-189).
 * Returns -1 if a value is negative (nothing after it is applied). */
int rb200_sumtree_set_host(double* tree_host, int32_t depth, const int64_t* idx_host,
                           const double* val_host, int64_t n, double* max_recorded_host);
/* Validity bookkeeping of n consecutive ReplayBuffer.add() calls, stack_size == 1
 * (circular_replay_buffer.py:468-522).  state = {add_count, episode length, num valid}. */
void rb200_replay_add_batch_host(const uint8_t* terminal_in_host, int64_t n, int64_t capacity,
                                 int32_t update_horizon, uint8_t* valid_host,
                                 uint8_t* terminal_store_host, int64_t* state_host);
/* SumTree.sample(query) on the host heap (sum_tree.py:93-131). */
int64_t rb200_sumtree_sample_host(const double* tree_host, int32_t depth, double query);
/* the same walk for queries[pos[0..n)] -> out[0..n) (PER retry check of one stratified draw) */
void rb200_sumtree_sample_many_host(const double* tree_host, int32_t depth, const double* queries,
                                    const int64_t* pos, int64_t n, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* REAGENT_B200_H_ */
