"""Batch types at the trainer boundary -- same names, fields and shapes as the reference's
`rlt` (reagent/core/types.py:48-108, :312-338, :688-816, :899-915).  Only the dense,
in-scope fields are kept; `.cuda()/.to()/.cpu()` fan out over tensor members like
TensorDataClass.__getattr__ does."""
import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class TensorDataClass:
    def _map(self, fn):
        out = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if isinstance(v, torch.Tensor):
                out[f.name] = fn(v)
            elif isinstance(v, TensorDataClass):
                out[f.name] = v._map(fn)
            else:
                out[f.name] = v
        return type(self)(**out)

    def cuda(self, *args, **kwargs):
        kwargs.setdefault("non_blocking", True)
        return self._map(lambda t: t.cuda(*args, **kwargs))

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def to(self, *args, **kwargs):
        return self._map(lambda t: t.to(*args, **kwargs))

    def float(self):
        return self._map(lambda t: t.float())

    def detach(self):
        return self._map(lambda t: t.detach())

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())


@dataclass
class FeatureData(TensorDataClass):
    # dense features, shape (batch_size, feature_dim)
    float_features: torch.Tensor

    def __post_init__(self):
        if self.float_features.ndim != 2:
            raise ValueError(f"float_features should be 2D; got {tuple(self.float_features.shape)}")


@dataclass
class ActorOutput(TensorDataClass):
    action: torch.Tensor
    log_prob: Optional[torch.Tensor] = None
    squashed_mean: Optional[torch.Tensor] = None


@dataclass
class ExtraData(TensorDataClass):
    mdp_id: Optional[torch.Tensor] = None
    sequence_number: Optional[torch.Tensor] = None
    action_probability: Optional[torch.Tensor] = None
    max_num_actions: Optional[int] = None
    metrics: Optional[torch.Tensor] = None


@dataclass
class BaseInput(TensorDataClass):
    state: FeatureData
    next_state: FeatureData
    reward: torch.Tensor
    time_diff: Optional[torch.Tensor]
    step: Optional[torch.Tensor]
    not_terminal: torch.Tensor

    def __len__(self):
        return self.state.float_features.size()[0]

    def batch_size(self):
        return len(self)


@dataclass
class DiscreteDqnInput(BaseInput):
    action: torch.Tensor = None
    next_action: torch.Tensor = None
    possible_actions_mask: torch.Tensor = None
    possible_next_actions_mask: torch.Tensor = None
    extras: Optional[ExtraData] = None

    @classmethod
    def from_dict(cls, batch):
        return cls(
            state=FeatureData(batch["state_features"]),
            next_state=FeatureData(batch["next_state_features"]),
            reward=batch["reward"],
            time_diff=batch.get("time_diff"),
            step=batch.get("step"),
            not_terminal=batch["not_terminal"],
            action=batch["action"],
            next_action=batch["next_action"],
            possible_actions_mask=batch["possible_actions_mask"],
            possible_next_actions_mask=batch["possible_next_actions_mask"],
            extras=batch.get("extras", ExtraData()),
        )


@dataclass
class PolicyNetworkInput(BaseInput):
    action: FeatureData = None
    next_action: FeatureData = None
    extras: Optional[ExtraData] = None

    @classmethod
    def from_dict(cls, batch):
        return cls(
            state=FeatureData(batch["state_features"]),
            next_state=FeatureData(batch["next_state_features"]),
            reward=batch["reward"],
            time_diff=batch.get("time_diff"),
            step=batch.get("step"),
            not_terminal=batch["not_terminal"],
            action=FeatureData(batch["action"]),
            next_action=FeatureData(batch["next_action"]),
            extras=batch.get("extras"),
        )


# ---- dense feature configuration (core/types.py:152-155, :181-215); sparse id-list features
# are out of scope of this package, so the config is dense-only ----
@dataclass
class FloatFeatureInfo:
    name: str
    feature_id: int


@dataclass
class ModelFeatureConfig:
    float_feature_infos: List[FloatFeatureInfo] = field(default_factory=list)

    @property
    def only_dense(self):
        return True


@dataclass
class ParametricDqnInput(BaseInput):
    """core/types.py:867-897: actions are feature vectors; the possible (next) actions of a row
    are tiled along the batch dimension -- (batch_size * max_num_action, action_dim)."""
    action: FeatureData = None
    next_action: FeatureData = None
    possible_actions: FeatureData = None
    possible_actions_mask: torch.Tensor = None
    possible_next_actions: FeatureData = None
    possible_next_actions_mask: torch.Tensor = None
    extras: Optional[ExtraData] = None
    weight: Optional[torch.Tensor] = None

    @classmethod
    def from_dict(cls, batch):
        return cls(
            state=FeatureData(batch["state_features"]),
            action=FeatureData(batch["action"]),
            next_state=FeatureData(batch["next_state_features"]),
            next_action=FeatureData(batch["next_action"]),
            possible_actions=FeatureData(batch["possible_actions"]),
            possible_actions_mask=batch["possible_actions_mask"],
            possible_next_actions=FeatureData(batch["possible_next_actions"]),
            possible_next_actions_mask=batch["possible_next_actions_mask"],
            reward=batch["reward"],
            not_terminal=batch["not_terminal"],
            time_diff=batch.get("time_diff"),
            step=batch.get("step"),
            extras=batch.get("extras"),
            weight=batch.get("weight"),
        )
