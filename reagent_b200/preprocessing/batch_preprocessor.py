"""Offline-workflow batch formatters: dict of raw tensors -> `rlt.*Input`.

Same classes, constructor arguments, dict keys and output fields as
reagent/preprocessing/batch_preprocessor.py:26-66 (DiscreteDqnBatchPreprocessor) and :117-157
(PolicyNetworkBatchPreprocessor).  The dense normalisation of state / next_state (/ action)
runs in the library's preprocessing kernel (rb200_preprocess, through `Preprocessor`), the
rest is index arithmetic on device tensors:

  not_terminal = possible_next_actions_mask.max(1)            (:43-44)
  action       = one_hot(action, A)                            (:45)
  next_action  = one_hot(next_action, A+1)[:, :A]              (:46-49; index A = "no action")
  reward / time_diff / step / mdp_id / sequence_number / action_probability -> (B,1)

`ParametricDqnBatchPreprocessor` (:69-114) belongs to ParametricDQN and is
not provided.
"""
from typing import Dict

import torch
import torch.nn as nn

from ..core import types as rlt
from .preprocessor import Preprocessor


class BatchPreprocessor(nn.Module):
    pass


def batch_to_device(batch: Dict[str, torch.Tensor], device: torch.device):
    return {k: v.to(device) for k, v in batch.items()}


def _device(use_gpu: bool, pre: Preprocessor) -> torch.device:
    # the reference picks cuda/cpu from use_gpu; the normalisation kernel only exists on the
    # GPU, so the batch follows the preprocessor's device and use_gpu=False is honoured only
    # for the placement of nothing (kept for signature compatibility)
    return pre.device if pre.device.type == "cuda" else torch.device("cuda" if use_gpu else "cpu")


def _extras(batch) -> rlt.ExtraData:
    return rlt.ExtraData(mdp_id=batch["mdp_id"].unsqueeze(1),
                         sequence_number=batch["sequence_number"].unsqueeze(1),
                         action_probability=batch["action_probability"].unsqueeze(1))


class DiscreteDqnBatchPreprocessor(BatchPreprocessor):
    def __init__(self, num_actions: int, state_preprocessor: Preprocessor,
                 use_gpu: bool = False) -> None:
        super().__init__()
        self.num_actions = num_actions
        self.state_preprocessor = state_preprocessor
        self.device = _device(use_gpu, state_preprocessor)

    def forward(self, batch: Dict[str, torch.Tensor]) -> rlt.DiscreteDqnInput:
        batch = batch_to_device(batch, self.device)
        A = self.num_actions
        state = self.state_preprocessor(batch["state_features"], batch["state_features_presence"])
        next_state = self.state_preprocessor(batch["next_state_features"],
                                             batch["next_state_features_presence"])
        # not terminal iff at least one next action is possible
        not_terminal = batch["possible_next_actions_mask"].max(dim=1)[0].float()
        action = torch.nn.functional.one_hot(batch["action"].to(torch.int64), A)
        # next_action may be A ("not available"): one-hot over A+1 classes, last column dropped
        next_action = torch.nn.functional.one_hot(batch["next_action"].to(torch.int64), A + 1)[:, :A]
        return rlt.DiscreteDqnInput(
            state=rlt.FeatureData(state), next_state=rlt.FeatureData(next_state),
            action=action, next_action=next_action,
            reward=batch["reward"].unsqueeze(1), time_diff=batch["time_diff"].unsqueeze(1),
            step=batch["step"].unsqueeze(1), not_terminal=not_terminal.unsqueeze(1),
            possible_actions_mask=batch["possible_actions_mask"],
            possible_next_actions_mask=batch["possible_next_actions_mask"],
            extras=_extras(batch))


class PolicyNetworkBatchPreprocessor(BatchPreprocessor):
    def __init__(self, state_preprocessor: Preprocessor, action_preprocessor: Preprocessor,
                 use_gpu: bool = False) -> None:
        super().__init__()
        self.state_preprocessor = state_preprocessor
        self.action_preprocessor = action_preprocessor
        self.device = _device(use_gpu, state_preprocessor)

    def forward(self, batch: Dict[str, torch.Tensor]) -> rlt.PolicyNetworkInput:
        batch = batch_to_device(batch, self.device)
        sp, ap = self.state_preprocessor, self.action_preprocessor
        return rlt.PolicyNetworkInput(
            state=rlt.FeatureData(sp(batch["state_features"], batch["state_features_presence"])),
            next_state=rlt.FeatureData(sp(batch["next_state_features"],
                                          batch["next_state_features_presence"])),
            action=rlt.FeatureData(ap(batch["action"], batch["action_presence"])),
            next_action=rlt.FeatureData(ap(batch["next_action"], batch["next_action_presence"])),
            reward=batch["reward"].unsqueeze(1), time_diff=batch["time_diff"].unsqueeze(1),
            step=batch["step"].unsqueeze(1), not_terminal=batch["not_terminal"].unsqueeze(1),
            extras=_extras(batch))
