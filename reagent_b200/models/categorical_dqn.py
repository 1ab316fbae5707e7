"""CategoricalDQN (reagent/models/categorical_dqn.py:12-35): a distributional network whose
(B, A, N) logits become a categorical distribution over `num_atoms` support points."""
import torch
import torch.nn.functional as F

from ..core import types as rlt
from .base import ModelBase


class CategoricalDQN(ModelBase):
    def __init__(self, distributional_network: ModelBase, *, qmin: float, qmax: float,
                 num_atoms: int) -> None:
        super().__init__()
        self.distributional_network = distributional_network
        self.support = torch.linspace(qmin, qmax, num_atoms)

    @property
    def arena(self):
        return self.distributional_network.arena

    def input_prototype(self):
        return self.distributional_network.input_prototype()

    def forward(self, state: rlt.FeatureData):
        dist = self.log_dist(state).exp()
        return (dist * self.support.to(dist.device)).sum(2)

    def log_dist(self, state: rlt.FeatureData) -> torch.Tensor:
        return F.log_softmax(self.distributional_network(state), -1)
