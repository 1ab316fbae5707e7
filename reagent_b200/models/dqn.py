"""FullyConnectedDQN: states -> one value per action (or per action and atom).

Public surface of reagent/models/dqn.py:16-63 -- positional (state_dim, action_dim, sizes,
activations) plus the keyword-only layer options, `action_dim`, `forward(state,
possible_actions_mask)` -- on top of the arena-backed FloatFeatureFullyConnected, whose forward
is a single fused launch.
"""
from typing import Optional

import torch

from ..core import types as rlt
from .fully_connected_network import FloatFeatureFullyConnected

# added to the scores of impossible actions so that a softmax gives them zero probability
INVALID_ACTION_CONSTANT: float = -1e10

# keyword-only options accepted for signature compatibility, with the reference's defaults; the
# base class raises NotImplementedError for the ones outside the in-scope layer set
_LAYER_OPTIONS = dict(output_activation="linear", num_atoms=None, use_batch_norm=False,
                      dropout_ratio=0.0, normalized_output=False, use_layer_norm=False)


class FullyConnectedDQN(FloatFeatureFullyConnected):
    def __init__(self, state_dim, action_dim, sizes, activations, **options) -> None:
        unknown = set(options) - set(_LAYER_OPTIONS)
        if unknown:
            raise TypeError(f"FullyConnectedDQN() got unexpected keyword argument(s) {sorted(unknown)}")
        opts = {**_LAYER_OPTIONS, **options}
        super().__init__(state_dim=state_dim, output_dim=action_dim, sizes=sizes,
                         activations=activations, **opts)
        self.action_dim = self.output_dim

    def forward(self, state: rlt.FeatureData,
                possible_actions_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        scores = super().forward(state=state)
        if possible_actions_mask is None:
            return scores
        # (used when the network scores actions for a policy: masked actions -> -1e10)
        return scores + INVALID_ACTION_CONSTANT * (1 - possible_actions_mask.float())
