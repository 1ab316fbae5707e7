"""DQNTrainerMixin / DQNTrainerBaseLightning: the parts of
reagent/training/dqn_trainer_base.py:23-452 that are on the hot path, including the CPE
heads (reward network, q_network_cpe, `_calculate_cpes`, :243-452): two extra MLPs evaluated
by the generic forward / backward / weight-gradient kernels with rb200_cpe_heads in between.
The offline evaluation machinery behind them (Evaluator, EvaluationDataPage, :454-509) is
reporting, not training, and stays out of scope."""
from typing import Dict, List, Optional

import torch

from .. import _lib
from ..core import types as rlt
from ..core.parameters import EvaluationParameters, RLParameters
from .reagent_lightning_module import ReAgentLightningModule
from .rl_trainer_pytorch import RLTrainerMixin


class DQNTrainerMixin:
    ACTION_NOT_POSSIBLE_VAL = -1e9

    def get_max_q_values(self, q_values, possible_actions_mask):
        return self.get_max_q_values_with_target(q_values, q_values, possible_actions_mask)

    def get_max_q_values_with_target(self, q_values, q_values_target, possible_actions_mask):
        """Host-visible utility with the reference's semantics (dqn_trainer_base.py:33-77).
        The training path does this inside the fused TD kernel; this method exists for
        callers / logging that use it directly (tiny tensors, torch plumbing)."""
        q_values = q_values.reshape(possible_actions_mask.shape)
        q_values_target = q_values_target.reshape(possible_actions_mask.shape)
        inverse_pna = 1 - possible_actions_mask
        impossible_action_penalty = self.ACTION_NOT_POSSIBLE_VAL * inverse_pna
        q_values = q_values + impossible_action_penalty
        q_values_target = q_values_target + impossible_action_penalty
        if self.double_q_learning:
            max_q_values, max_indicies = torch.max(q_values, dim=1, keepdim=True)
            max_q_values_target = torch.gather(q_values_target, 1, max_indicies)
        else:
            max_q_values_target, max_indicies = torch.max(q_values_target, dim=1, keepdim=True)
        return max_q_values_target, max_indicies


class DQNTrainerBaseLightning(DQNTrainerMixin, RLTrainerMixin, ReAgentLightningModule):
    def __init__(
        self,
        rl_parameters: RLParameters,
        metrics_to_score=None,
        actions: Optional[List[str]] = None,
        evaluation_parameters: Optional[EvaluationParameters] = None,
    ):
        super().__init__()
        self.rl_parameters = rl_parameters
        self.time_diff_unit_length = rl_parameters.time_diff_unit_length
        self.tensorboard_logging_freq = rl_parameters.tensorboard_logging_freq
        self.calc_cpe_in_training = bool(
            evaluation_parameters and evaluation_parameters.calc_cpe_in_training)
        assert actions is not None
        self._actions: List[str] = actions
        if rl_parameters.q_network_loss == "mse":
            self.q_network_loss_kind = _lib.LOSS_MSE
        elif rl_parameters.q_network_loss == "huber":
            self.q_network_loss_kind = _lib.LOSS_HUBER
        else:
            raise Exception(
                "Q-Network loss type {} not valid loss.".format(rl_parameters.q_network_loss))
        if metrics_to_score:
            self.metrics_to_score = metrics_to_score + ["reward"]
        else:
            self.metrics_to_score = ["reward"]
        self._init_reward_boosts(rl_parameters.reward_boost)
        # mirror of the reference's host-syncing `.any()` input check; off by default
        self.strict_input_checks = False

    def _init_reward_boosts(self, rl_reward_boost: Optional[Dict[str, float]]) -> None:
        reward_boosts = torch.zeros([1, len(self._actions)])
        self._has_reward_boost = False
        if rl_reward_boost is not None:
            for k in rl_reward_boost.keys():
                i = self._actions.index(k)
                reward_boosts[0, i] = rl_reward_boost[k]
                self._has_reward_boost = True
        self.register_buffer("reward_boosts", reward_boosts)

    def _initialize_cpe(self, reward_network, q_network_cpe, q_network_cpe_target, optimizer):
        """dqn_trainer_base.py:243-311: store the reward / CPE networks and the offsets of
        every metric's block of `num_actions` outputs (no Evaluator: reporting is out of scope)."""
        self._cpe_ws = None
        if not self.calc_cpe_in_training:
            self.reward_network = None
            self.q_network_cpe = None
            self.q_network_cpe_target = None
            return
        assert reward_network is not None, "reward_network is required for CPE"
        assert q_network_cpe is not None and q_network_cpe_target is not None, (
            "q_network_cpe and q_network_cpe_target are required for CPE")
        self.reward_network = reward_network
        self.reward_network_optimizer = optimizer
        self.q_network_cpe = q_network_cpe
        self.q_network_cpe_target = q_network_cpe_target
        self.q_network_cpe_optimizer = optimizer
        num_output_nodes = len(self.metrics_to_score) * self.num_actions
        self.register_buffer("reward_idx_offsets",
                             torch.arange(0, num_output_nodes, self.num_actions, dtype=torch.long))

    def _configure_cpe_optimizers(self):
        """(target params, source params, [reward optimizer, cpe optimizer]) -- :313-330."""
        target_params = list(self.q_network_cpe_target.parameters())
        source_params = list(self.q_network_cpe.parameters())
        optimizers = [
            self.reward_network_optimizer.make_optimizer_scheduler(self.reward_network.parameters()),
            self.q_network_cpe_optimizer.make_optimizer_scheduler(self.q_network_cpe.parameters()),
        ]
        return target_params, source_params, optimizers

    def _cpe_workspace(self, B: int, device):
        from .workspace import NetWorkspace

        ws = self._cpe_ws
        if ws is None or ws["B"] != B or ws["dev"] != device:
            MA = len(self.metrics_to_score) * self.num_actions
            ws = {
                "B": B, "dev": device,
                "reward": NetWorkspace(self.reward_network.arena, B, device),
                "qcpe": NetWorkspace(self.q_network_cpe.arena, B, device),
                "next_scores": torch.empty(B, self.num_actions, device=device),
                "reward_est": torch.empty(B, MA, device=device),
                "qcpe_out": torch.empty(B, MA, device=device),
                "qcpe_t_next": torch.empty(B, MA, device=device),
                "prop_next": torch.empty(B, self.num_actions, device=device),
                "loss_partials": torch.zeros(2 * ((B + 255) // 256), device=device),
                "loss": torch.zeros(2, device=device),
                "counter": torch.zeros(1, dtype=torch.int32, device=device),
            }
            self._cpe_ws = ws
        return ws

    def _calculate_cpes(self, training_batch: rlt.DiscreteDqnInput):
        """_calculate_cpes (:332-452) on the device: returns the [2] loss tensor (reward loss,
        CPE q-value loss) and leaves the gradient partials of both networks in their arenas.
        Runs AFTER the q-network step of the same batch, as in the reference's generator
        (all_next_action_scores is evaluated after `yield td_loss`, dqn_trainer.py:266-268)."""
        from .workspace import wgrad

        lib, st = _lib.lib(), _lib.cur_stream()
        state = training_batch.state.float_features.float().contiguous()
        next_state = training_batch.next_state.float_features.float().contiguous()
        B, dev = state.shape[0], state.device
        _lib.require_current_device(dev)
        ws = self._cpe_workspace(B, dev)
        keep = [state, next_state]

        def P(t):
            t = _lib.on_device(t.float().contiguous(), dev)
            keep.append(t)
            return _lib.ptr(t, dev)

        def fwd(net, x, out, save=None):
            net.arena.refresh()
            rc = lib.rb200_mlp_forward(net.arena.desc(), x.data_ptr(), x.shape[1], None, 0, B,
                                       out.data_ptr(), save, st)
            _lib.check(rc, "rb200_mlp_forward")

        fwd(self.q_network, next_state, ws["next_scores"])
        fwd(self.reward_network, state, ws["reward_est"], ws["reward"].c)
        fwd(self.q_network_cpe, state, ws["qcpe_out"], ws["qcpe"].c)
        fwd(self.q_network_cpe_target, next_state, ws["qcpe_t_next"])
        metrics = training_batch.extras.metrics if training_batch.extras is not None else None
        mrc = training_batch.reward if metrics is None else torch.cat((training_batch.reward, metrics), dim=1)
        M = len(self.metrics_to_score)
        assert mrc.shape[1] == M, f"reward + metrics have {mrc.shape[1]} columns, metrics_to_score {M}"
        a = _lib.CpeArgsT()
        a.batch, a.num_actions, a.num_metrics = B, self.num_actions, M
        a.next_scores = ws["next_scores"].data_ptr()
        mask = (training_batch.possible_next_actions_mask if self.maxq_learning
                else training_batch.next_action)
        a.mask = P(mask)
        a.temperature = float(self.rl_temperature)
        a.action = P(training_batch.action)
        a.metrics_reward = P(mrc)
        a.gamma = float(self.gamma)
        a.discount_mode = _lib.DISCOUNT_CONST
        if self.use_seq_num_diff_as_time_diff:
            a.discount_src, a.discount_mode = P(training_batch.time_diff.reshape(-1)), _lib.DISCOUNT_POW
        if self.multi_steps is not None:
            a.discount_src, a.discount_mode = P(training_batch.step.reshape(-1)), _lib.DISCOUNT_POW
        a.not_terminal = P(training_batch.not_terminal.reshape(-1))
        a.reward_est = ws["reward_est"].data_ptr()
        a.qcpe = ws["qcpe_out"].data_ptr()
        a.qcpe_target_next = ws["qcpe_t_next"].data_ptr()
        a.loss_kind = self.q_network_loss_kind
        Lr, Lc = len(self.reward_network.arena.acts), len(self.q_network_cpe.arena.acts)
        a.dz_reward = ws["reward"].dz[Lr - 1].data_ptr()
        a.dz_qcpe = ws["qcpe"].dz[Lc - 1].data_ptr()
        a.propensities_next = ws["prop_next"].data_ptr()
        a.loss_partials = ws["loss_partials"].data_ptr()
        a.loss = ws["loss"].data_ptr()
        a.tile_counter = ws["counter"].data_ptr()
        _lib.check(lib.rb200_cpe_heads(a, st), "rb200_cpe_heads")
        for net, w in ((self.reward_network, ws["reward"]), (self.q_network_cpe, ws["qcpe"])):
            ar = net.arena
            L = len(ar.acts)
            rc = lib.rb200_mlp_backward(ar.desc(), w.dz[L - 1].data_ptr(), B, w.c, st)
            _lib.check(rc, "rb200_mlp_backward")
            wgrad(ar, w, state, B)
            ar.finish_grads()
        self.model_propensities_next_states = ws["prop_next"]
        return ws["loss"]

    def _check_input(self, training_batch: rlt.DiscreteDqnInput):
        assert isinstance(training_batch, rlt.DiscreteDqnInput)
        assert training_batch.not_terminal.dim() == training_batch.reward.dim() == 2
        assert training_batch.not_terminal.shape[1] == training_batch.reward.shape[1] == 1
        assert training_batch.action.dim() == training_batch.next_action.dim() == 2
        assert (training_batch.action.shape[1] == training_batch.next_action.shape[1]
                == self.num_actions)
        if self.strict_input_checks and training_batch.possible_next_actions_mask is not None:
            if torch.logical_and(
                training_batch.possible_next_actions_mask.float().sum(dim=1) == 0,
                training_batch.not_terminal.squeeze().bool(),
            ).any():
                raise ValueError(
                    "No possible next actions. Should the environment have terminated?")

    @property
    def num_actions(self) -> int:
        assert self._actions is not None, "Not a discrete action DQN"
        return len(self._actions)

    @torch.no_grad()
    def boost_rewards(self, rewards: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        """Utility with the reference's semantics (dqn_trainer_base.py:216-241); the
        training path applies the boost inside the fused kernel."""
        reward_boosts = torch.sum(actions.float() * self.reward_boosts, dim=1, keepdim=True)
        return rewards + reward_boosts
