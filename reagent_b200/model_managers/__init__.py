"""Model managers: `build_trainer(normalization_data_map, use_gpu, reward_options=None)` with
the reference's flow (reagent/model_managers/model_manager.py:84-96 and
discrete/discrete_dqn.py:63-116, discrete/discrete_qrdqn.py:73-121, actor_critic/sac.py:80-113,
actor_critic/td3.py:70-102): build the networks from the net builders, copy the target, hand
everything to the trainer.  Policies, serving modules, data modules and reporters are out of
scope."""
from dataclasses import dataclass, field
from typing import Union, Dict, List, Optional

from ..core.parameters import (EvaluationParameters, NormalizationData, NormalizationKey,
                               RLParameters)
from ..net_builder import (ActorFullyConnected, Dueling, DuelingQuantile, FullyConnected,
                           GaussianFullyConnected, ParametricFullyConnected, Quantile)
from ..optimizer import Optimizer__Union
from ..training import DQNTrainer, QRDQNTrainer, SACTrainer, TD3Trainer


def _device(use_gpu: bool):
    if not use_gpu:
        raise RuntimeError("reagent_b200 trainers run on CUDA only; build_trainer needs use_gpu=True")
    return "cuda"


class _DiscretePolicyMixin:
    def create_policy(self, trainer_module, serving: bool = False, normalization_data_map=None):
        """Online policy (reagent/model_managers/discrete_dqn_base.py:84-103): greedy over the
        fused Q-network forward.  Serving modules are out of scope."""
        if serving:
            raise NotImplementedError("serving modules are out of scope of reagent_b200")
        from ..gym.policies import GreedyActionSampler, Policy, discrete_dqn_scorer

        return Policy(scorer=discrete_dqn_scorer(trainer_module.q_network),
                      sampler=GreedyActionSampler())


class _ActorPolicyMixin:
    def create_policy(self, trainer_module, serving: bool = False, normalization_data_map=None):
        """reagent/model_managers/actor_critic_base.py:104-118: the actor's forward is the act."""
        if serving:
            raise NotImplementedError("serving modules are out of scope of reagent_b200")
        from ..gym.policies import ActorPolicyWrapper

        return ActorPolicyWrapper(trainer_module.actor_network)


@dataclass
class DiscreteDQN(_DiscretePolicyMixin):
    actions: List[str]
    rl: RLParameters = field(default_factory=RLParameters)
    double_q_learning: bool = True
    minibatch_size: int = 1024
    optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)
    # reagent/model_managers/discrete/discrete_dqn.py:33-36: the reference defaults to Dueling
    net_builder: Union[Dueling, FullyConnected] = field(default_factory=Dueling)
    # :37-42: the reward / CPE networks are plain FullyConnected
    cpe_net_builder: Union[Dueling, FullyConnected] = field(default_factory=FullyConnected)
    # EvaluationParameters() has calc_cpe_in_training=True, as in the reference
    # (reagent/core/parameters.py:118-120)
    eval_parameters: EvaluationParameters = field(default_factory=EvaluationParameters)
    metrics_to_score: Optional[List[str]] = None

    def build_trainer(self, normalization_data_map: Dict[str, NormalizationData], use_gpu: bool,
                      reward_options=None) -> DQNTrainer:
        """discrete_dqn.py:73-116"""
        dev = _device(use_gpu)
        s_norm = normalization_data_map[NormalizationKey.STATE]
        q_network = self.net_builder.build_q_network(None, s_norm, len(self.actions)).to(dev)
        q_network_target = q_network.get_target_network()
        reward_network = q_network_cpe = q_network_cpe_target = None
        metrics = list(self.metrics_to_score or [])
        if self.eval_parameters.calc_cpe_in_training:
            n_out = (len(metrics) + 1) * len(self.actions)  # metrics + reward
            reward_network = self.cpe_net_builder.build_q_network(None, s_norm, n_out).to(dev)
            q_network_cpe = self.cpe_net_builder.build_q_network(None, s_norm, n_out).to(dev)
            q_network_cpe_target = q_network_cpe.get_target_network()
        return DQNTrainer(
            q_network=q_network, q_network_target=q_network_target,
            reward_network=reward_network, q_network_cpe=q_network_cpe,
            q_network_cpe_target=q_network_cpe_target, metrics_to_score=metrics,
            actions=self.actions, rl=self.rl, double_q_learning=self.double_q_learning,
            minibatch_size=self.minibatch_size, optimizer=self.optimizer,
            evaluation=self.eval_parameters).to(dev)


@dataclass
class DiscreteQRDQN(_DiscretePolicyMixin):
    actions: List[str]
    rl: RLParameters = field(default_factory=RLParameters)
    double_q_learning: bool = True
    num_atoms: int = 51
    minibatch_size: int = 1024
    optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)
    # reagent/model_managers/discrete/discrete_qrdqn.py:39-43: the reference defaults to DuelingQuantile
    net_builder: Union[DuelingQuantile, Quantile] = field(default_factory=DuelingQuantile)
    eval_parameters: EvaluationParameters = field(
        default_factory=lambda: EvaluationParameters(calc_cpe_in_training=False))

    def build_trainer(self, normalization_data_map, use_gpu: bool, reward_options=None):
        dev = _device(use_gpu)
        q_network = self.net_builder.build_q_network(
            normalization_data_map[NormalizationKey.STATE], len(self.actions),
            self.num_atoms).to(dev)
        q_network_target = q_network.get_target_network()
        return QRDQNTrainer(
            q_network=q_network, q_network_target=q_network_target, actions=self.actions,
            rl=self.rl, double_q_learning=self.double_q_learning, num_atoms=self.num_atoms,
            minibatch_size=self.minibatch_size, optimizer=self.optimizer,
            evaluation=self.eval_parameters).to(dev)


@dataclass
class SAC(_ActorPolicyMixin):
    rl: RLParameters = field(default_factory=RLParameters)
    actor_net_builder: GaussianFullyConnected = field(default_factory=GaussianFullyConnected)
    critic_net_builder: ParametricFullyConnected = field(default_factory=ParametricFullyConnected)
    use_2_q_functions: bool = True
    minibatch_size: int = 1024
    entropy_temperature: float = 0.01
    target_entropy: float = -1.0
    q_network_optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)
    actor_network_optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)
    alpha_optimizer: Optional[Optimizer__Union] = field(default_factory=Optimizer__Union.default)

    def build_trainer(self, normalization_data_map, use_gpu: bool, reward_options=None):
        dev = _device(use_gpu)
        s, a = (normalization_data_map[NormalizationKey.STATE],
                normalization_data_map[NormalizationKey.ACTION])
        actor = self.actor_net_builder.build_actor(None, s, a).to(dev)
        q1 = self.critic_net_builder.build_q_network(s, a).to(dev)
        q2 = self.critic_net_builder.build_q_network(s, a).to(dev) if self.use_2_q_functions else None
        return SACTrainer(
            actor_network=actor, q1_network=q1, q2_network=q2, value_network=None, rl=self.rl,
            q_network_optimizer=self.q_network_optimizer,
            actor_network_optimizer=self.actor_network_optimizer,
            alpha_optimizer=self.alpha_optimizer, minibatch_size=self.minibatch_size,
            entropy_temperature=self.entropy_temperature,
            target_entropy=self.target_entropy).to(dev)


@dataclass
class TD3(_ActorPolicyMixin):
    rl: RLParameters = field(default_factory=RLParameters)
    actor_net_builder: ActorFullyConnected = field(default_factory=ActorFullyConnected)
    critic_net_builder: ParametricFullyConnected = field(default_factory=ParametricFullyConnected)
    use_2_q_functions: bool = True
    minibatch_size: int = 64
    noise_variance: float = 0.2
    noise_clip: float = 0.5
    delayed_policy_update: int = 2
    q_network_optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)
    actor_network_optimizer: Optimizer__Union = field(default_factory=Optimizer__Union.default)

    def build_trainer(self, normalization_data_map, use_gpu: bool, reward_options=None):
        dev = _device(use_gpu)
        s, a = (normalization_data_map[NormalizationKey.STATE],
                normalization_data_map[NormalizationKey.ACTION])
        actor = self.actor_net_builder.build_actor(None, s, a).to(dev)
        q1 = self.critic_net_builder.build_q_network(s, a).to(dev)
        q2 = self.critic_net_builder.build_q_network(s, a).to(dev) if self.use_2_q_functions else None
        return TD3Trainer(
            actor_network=actor, q1_network=q1, q2_network=q2, rl=self.rl,
            q_network_optimizer=self.q_network_optimizer,
            actor_network_optimizer=self.actor_network_optimizer,
            minibatch_size=self.minibatch_size, noise_variance=self.noise_variance,
            noise_clip=self.noise_clip, delayed_policy_update=self.delayed_policy_update).to(dev)
