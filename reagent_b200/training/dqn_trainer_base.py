"""DQNTrainerMixin / DQNTrainerBaseLightning: the parts of
reagent/training/dqn_trainer_base.py:23-241 that are on the hot path.  CPE heads
(reward network, q_network_cpe, `_calculate_cpes`, :243-509) are SURVEY.md 8f rank 4
("next") and raise NotImplementedError when requested."""
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
        if self.calc_cpe_in_training:
            raise NotImplementedError(
                "CPE heads (reward_network / q_network_cpe) are not part of the fused hot path "
                "yet (SURVEY.md 8f rank 4); pass evaluation=EvaluationParameters("
                "calc_cpe_in_training=False) as every reference gym config does")
        self.reward_network = None
        self.q_network_cpe = None
        self.q_network_cpe_target = None

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
