"""Actors (reagent/models/actor.py:44-261) over the arena-backed FullyConnectedNetwork.

Training never calls these forwards: the SAC / TD3 kernels evaluate the actor inside the
fused update.  forward() is the act-time path (one fused MLP launch + the Gaussian head as
a few elementwise torch ops on a [B, 2A] tensor)."""
import math
from typing import List, Optional

import torch

from ..core import types as rlt
from ..core.parameters import CONTINUOUS_TRAINING_ACTION_RANGE
from .base import ModelBase
from .fully_connected_network import FullyConnectedNetwork

LOG_PROB_MIN: float = -2.0
LOG_PROB_MAX = 2.0


class FullyConnectedActor(ModelBase):
    def __init__(
        self,
        state_dim: int,
        action_dim: int,
        sizes: List[int],
        activations: List[str],
        use_batch_norm: bool = False,
        action_activation: str = "tanh",
        exploration_variance: Optional[float] = None,
    ) -> None:
        super().__init__()
        assert state_dim > 0, "state_dim must be > 0, got {}".format(state_dim)
        assert action_dim > 0, "action_dim must be > 0, got {}".format(action_dim)
        self.state_dim = state_dim
        self.action_dim = action_dim
        assert len(sizes) == len(activations), (
            "The numbers of sizes and activations must match; got {} vs {}".format(
                len(sizes), len(activations)))
        self.action_activation = action_activation
        self.fc = FullyConnectedNetwork(
            [state_dim] + list(sizes) + [action_dim],
            list(activations) + [self.action_activation],
            use_batch_norm=use_batch_norm,
        )
        self.exploration_variance = exploration_variance
        if exploration_variance is not None:
            assert exploration_variance > 0

    @property
    def arena(self):
        return self.fc.arena

    def input_prototype(self):
        return rlt.FeatureData(torch.randn(1, self.state_dim))

    def forward(self, state: rlt.FeatureData) -> rlt.ActorOutput:
        action = self.fc(state.float_features)
        batch_size = action.shape[0]
        if self.exploration_variance is None:
            log_prob = torch.zeros(batch_size, 1, device=action.device)
            return rlt.ActorOutput(action=action, log_prob=log_prob)
        scale = self.exploration_variance
        noise = torch.randn(batch_size, self.action_dim, device=action.device) * scale
        log_prob = (-(noise ** 2) / (2 * scale ** 2) - math.log(scale)
                    - math.log(math.sqrt(2 * math.pi))).sum(dim=1).view(-1, 1)
        log_prob = log_prob.clamp(LOG_PROB_MIN, LOG_PROB_MAX)
        action = (action + noise).clamp(*CONTINUOUS_TRAINING_ACTION_RANGE)
        return rlt.ActorOutput(action=action, log_prob=log_prob)


class GaussianFullyConnectedActor(ModelBase):
    def __init__(
        self,
        state_dim: int,
        action_dim: int,
        sizes: List[int],
        activations: List[str],
        scale: float = 0.05,
        use_batch_norm: bool = False,
        use_layer_norm: bool = False,
        use_l2_normalization: bool = False,
    ) -> None:
        super().__init__()
        assert state_dim > 0, "state_dim must be > 0, got {}".format(state_dim)
        assert action_dim > 0, "action_dim must be > 0, got {}".format(action_dim)
        if use_layer_norm or use_l2_normalization:
            raise NotImplementedError("layer norm / l2 normalisation are out of scope (M4)")
        self.state_dim = state_dim
        self.action_dim = action_dim
        assert len(sizes) == len(activations), (
            "The numbers of sizes and activations must match; got {} vs {}".format(
                len(sizes), len(activations)))
        self.fc = FullyConnectedNetwork(
            [state_dim] + list(sizes) + [action_dim * 2],
            list(activations) + ["linear"],
            use_batch_norm=use_batch_norm,
        )
        self.use_layer_norm = False
        self.use_l2_normalization = False
        self.const = math.log(math.sqrt(2 * math.pi))
        self.eps = 1e-6

    @property
    def arena(self):
        return self.fc.arena

    def input_prototype(self):
        return rlt.FeatureData(torch.randn(1, self.state_dim))

    def _get_loc_and_scale_log(self, state: rlt.FeatureData):
        loc_scale = self.fc(state.float_features)
        loc = loc_scale[::, : self.action_dim]
        scale_log = loc_scale[::, self.action_dim:].clamp(LOG_PROB_MIN, LOG_PROB_MAX)
        return loc, scale_log

    def _squash_raw_action(self, raw_action: torch.Tensor) -> torch.Tensor:
        return torch.clamp(torch.tanh(raw_action), -1.0 + self.eps, 1.0 - self.eps)

    def _log_prob_from(self, loc, scale_log, squashed_action):
        raw_action = torch.atanh(squashed_action)
        r = (raw_action - loc) / scale_log.exp()
        log_prob = -(r ** 2) / 2 - scale_log - self.const
        squash_correction = (1 - squashed_action ** 2 + self.eps).log()
        return torch.sum(log_prob - squash_correction, dim=1).reshape(-1, 1)

    @torch.no_grad()
    def forward(self, state: rlt.FeatureData):
        loc, scale_log = self._get_loc_and_scale_log(state)
        r = torch.randn_like(scale_log)
        raw_action = loc + r * scale_log.exp()
        squashed_action = self._squash_raw_action(raw_action)
        squashed_loc = self._squash_raw_action(loc)
        return rlt.ActorOutput(
            action=squashed_action,
            log_prob=self._log_prob_from(loc, scale_log, squashed_action),
            squashed_mean=squashed_loc,
        )

    @torch.no_grad()
    def get_log_prob(self, state: rlt.FeatureData, squashed_action: torch.Tensor):
        loc, scale_log = self._get_loc_and_scale_log(state)
        return self._log_prob_from(loc, scale_log, squashed_action)
