"""Net builders with the reference's build_* signatures (reagent/net_builder/**): config
dataclass -> arena-backed model.  Normalization data only sizes the input/output dims
(get_num_output_features); serving wrappers (build_serving_module) are out of scope."""
from dataclasses import dataclass, field
from typing import List

from ..core.parameters import NormalizationData
from ..models import (DuelingQNetwork, FullyConnectedActor, FullyConnectedCritic,
                      FullyConnectedDQN, GaussianFullyConnectedActor)
from ..preprocessing.normalization import get_num_output_features


def _dim(normalization_data: NormalizationData) -> int:
    return get_num_output_features(normalization_data.dense_normalization_parameters)


@dataclass
class FullyConnected:
    """reagent/net_builder/discrete_dqn/fully_connected.py:16-46"""
    sizes: List[int] = field(default_factory=lambda: [256, 128])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])
    dropout_ratio: float = 0.0
    use_batch_norm: bool = False

    def build_q_network(self, state_feature_config, state_normalization_data: NormalizationData,
                        output_dim: int):
        return FullyConnectedDQN(state_dim=_dim(state_normalization_data), action_dim=output_dim,
                                 sizes=self.sizes, activations=self.activations,
                                 dropout_ratio=self.dropout_ratio,
                                 use_batch_norm=self.use_batch_norm)


@dataclass
class Dueling:
    """reagent/net_builder/discrete_dqn/dueling.py:16-40 (the default of the DiscreteDQN manager)"""
    sizes: List[int] = field(default_factory=lambda: [256, 128])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])

    def __post_init__(self):
        assert len(self.sizes) == len(self.activations), (
            f"Must have the same numbers of sizes and activations; got: "
            f"{self.sizes}, {self.activations}")

    def build_q_network(self, state_feature_config, state_normalization_data: NormalizationData,
                        output_dim: int):
        return DuelingQNetwork.make_fully_connected(_dim(state_normalization_data), output_dim,
                                                    self.sizes, self.activations)


@dataclass
class Quantile:
    """reagent/net_builder/quantile_dqn/quantile.py:15-44"""
    sizes: List[int] = field(default_factory=lambda: [256, 128])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])
    dropout_ratio: float = 0.0

    def build_q_network(self, state_normalization_data: NormalizationData, output_dim: int,
                        num_atoms: int):
        return FullyConnectedDQN(state_dim=_dim(state_normalization_data), action_dim=output_dim,
                                 sizes=self.sizes, activations=self.activations,
                                 num_atoms=num_atoms, dropout_ratio=self.dropout_ratio)


@dataclass
class DuelingQuantile:
    """reagent/net_builder/quantile_dqn/dueling_quantile.py:16-40"""
    sizes: List[int] = field(default_factory=lambda: [256, 128])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])

    def __post_init__(self):
        assert len(self.sizes) == len(self.activations), (
            f"Must have the same numbers of sizes and activations; got: {self.sizes}, {self.activations}")

    def build_q_network(self, state_normalization_data: NormalizationData, output_dim: int,
                        num_atoms: int):
        return DuelingQNetwork.make_fully_connected(
            _dim(state_normalization_data), output_dim, layers=self.sizes,
            activations=self.activations, num_atoms=num_atoms)


@dataclass
class ParametricFullyConnected:
    """reagent/net_builder/parametric_dqn/fully_connected.py:16-54 (the SAC / TD3 critics)"""
    sizes: List[int] = field(default_factory=lambda: [128, 64])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])
    use_batch_norm: bool = False
    use_layer_norm: bool = False
    final_activation: str = "linear"

    def build_q_network(self, state_normalization_data: NormalizationData,
                        action_normalization_data: NormalizationData, output_dim: int = 1):
        return FullyConnectedCritic(
            _dim(state_normalization_data), _dim(action_normalization_data), sizes=self.sizes,
            activations=self.activations, use_batch_norm=self.use_batch_norm,
            use_layer_norm=self.use_layer_norm, output_dim=output_dim,
            final_activation=self.final_activation)


@dataclass
class GaussianFullyConnected:
    """reagent/net_builder/continuous_actor/gaussian_fully_connected.py:23-82"""
    sizes: List[int] = field(default_factory=lambda: [128, 64])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])
    use_batch_norm: bool = False
    use_layer_norm: bool = False
    use_l2_normalization: bool = False

    def build_actor(self, state_feature_config, state_normalization_data: NormalizationData,
                    action_normalization_data: NormalizationData):
        return GaussianFullyConnectedActor(
            state_dim=_dim(state_normalization_data), action_dim=_dim(action_normalization_data),
            sizes=self.sizes, activations=self.activations, use_batch_norm=self.use_batch_norm,
            use_layer_norm=self.use_layer_norm, use_l2_normalization=self.use_l2_normalization)


@dataclass
class ActorFullyConnected:
    """reagent/net_builder/continuous_actor/fully_connected.py:22-76"""
    sizes: List[int] = field(default_factory=lambda: [128, 64])
    activations: List[str] = field(default_factory=lambda: ["relu", "relu"])
    use_batch_norm: bool = False
    action_activation: str = "tanh"
    exploration_variance: float = None

    def build_actor(self, state_feature_config, state_normalization_data: NormalizationData,
                    action_normalization_data: NormalizationData):
        return FullyConnectedActor(
            state_dim=_dim(state_normalization_data), action_dim=_dim(action_normalization_data),
            sizes=self.sizes, activations=self.activations, use_batch_norm=self.use_batch_norm,
            action_activation=self.action_activation,
            exploration_variance=self.exploration_variance)
