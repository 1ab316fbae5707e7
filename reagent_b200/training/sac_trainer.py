"""SACTrainer with the reference's constructor, optimizer order and generator protocol
(reagent/training/sac_trainer.py:50-385) for the in-scope configuration: twin (or single)
critics, no value network, learnable or fixed entropy temperature.

Launch sequence of one update (value_network=None):
  rb200_ac_critic_step  target + q1/q2 losses + critic dZ chains     sac_trainer.py:214-248
  rb200_mlp_wgrad x2, Adam(q1), Adam(q2) [+ fused Polyak]            optimizer.py / soft_update.py
  rb200_ac_actor_step   actor + alpha losses, backward through the UPDATED critics  :254-322
  rb200_mlp_wgrad, Adam(actor), Adam(log_alpha) -> entropy_temperature = exp(log_alpha)
Sequential dependence of the reference is kept: the actor step sees the
post-update critics; the new temperature takes effect in the next batch.
"""
import copy
from typing import List, Optional

import numpy as np
import torch

from .. import _lib
from ..core import types as rlt
from ..core.parameters import RLParameters
from ..models.arena import ScalarArena
from ..optimizer import Optimizer__Union, SoftUpdate
from .actor_critic_base import ActorCriticBase

_DEFAULT = object()


class SACTrainer(ActorCriticBase):
    ALGO = _lib.ALGO_SAC

    def __init__(
        self,
        actor_network,
        q1_network,
        q2_network=None,
        value_network=None,
        rl: Optional[RLParameters] = None,
        q_network_optimizer: Optional[Optimizer__Union] = None,
        value_network_optimizer: Optional[Optimizer__Union] = None,
        actor_network_optimizer: Optional[Optimizer__Union] = None,
        alpha_optimizer=_DEFAULT,
        minibatch_size: int = 1024,
        entropy_temperature: float = 0.01,
        logged_action_uniform_prior: bool = True,
        target_entropy: float = -1.0,
        action_embedding_kld_weight: Optional[float] = None,
        apply_kld_on_mean: bool = False,
        action_embedding_mean: Optional[List[float]] = None,
        action_embedding_variance: Optional[List[float]] = None,
        crr_config=None,
        backprop_through_log_prob: bool = True,
    ) -> None:
        super().__init__()
        self._ac_init()
        if value_network is not None:
            raise NotImplementedError("SAC with a value network is out of scope (SURVEY.md T7)")
        if crr_config is not None or action_embedding_kld_weight:
            raise NotImplementedError("CRR weighting / action-embedding KLD are out of scope")
        self.rl_parameters = RLParameters() if rl is None else rl
        self.q1_network = q1_network
        self.q2_network = q2_network
        self.q_network_optimizer = q_network_optimizer or Optimizer__Union.default()
        self.value_network = None
        self.value_network_optimizer = value_network_optimizer or Optimizer__Union.default()
        self.q1_network_target = copy.deepcopy(self.q1_network)
        self.q2_network_target = copy.deepcopy(self.q2_network)
        self.actor_network = actor_network
        self.actor_network_optimizer = actor_network_optimizer or Optimizer__Union.default()
        self.entropy_temperature = entropy_temperature
        self.alpha_optimizer = (Optimizer__Union.default() if alpha_optimizer is _DEFAULT
                                else alpha_optimizer)
        if self.alpha_optimizer is not None:
            self.target_entropy = target_entropy
            # the reference keeps log_alpha in float64 (np.log -> torch.tensor); the fused Adam
            # is fp32 -- the difference is ~1e-8 relative, far inside the parity tolerance
            self.log_alpha = torch.nn.Parameter(
                torch.tensor([np.log(self.entropy_temperature)], dtype=torch.float32))
        else:
            self.target_entropy = target_entropy
        # not part of the state_dict (the reference has no such key): re-derived from log_alpha
        self.register_buffer("_alpha_dev", torch.tensor([float(entropy_temperature)]),
                             persistent=False)
        self.register_load_state_dict_post_hook(SACTrainer._rederive_alpha)
        self.logged_action_uniform_prior = logged_action_uniform_prior
        self.add_kld_to_loss = False
        self.crr_config = None
        self.backprop_through_log_prob = backprop_through_log_prob
        self.minibatch_size = minibatch_size

    @staticmethod
    def _rederive_alpha(module, incompatible_keys):
        if module.alpha_optimizer is not None:  # sac_trainer.py:322
            with torch.no_grad():
                module._alpha_dev.copy_(module.log_alpha.data.exp().to(module._alpha_dev.device))
            module.entropy_temperature = module._alpha_dev

    def configure_optimizers(self):
        """q1, q2, actor, alpha, SoftUpdate (sac_trainer.py:148-193)."""
        optimizers = []
        optimizers.append(
            self.q_network_optimizer.make_optimizer_scheduler(self.q1_network.parameters()))
        if self.q2_network:
            optimizers.append(
                self.q_network_optimizer.make_optimizer_scheduler(self.q2_network.parameters()))
        optimizers.append(
            self.actor_network_optimizer.make_optimizer_scheduler(
                self.actor_network.parameters()))
        if self.alpha_optimizer is not None:
            optimizers.append(self.alpha_optimizer.make_optimizer_scheduler([self.log_alpha]))
        target_params = list(self.q1_network_target.parameters())
        source_params = list(self.q1_network.parameters())
        if self.q2_network:
            target_params += list(self.q2_network_target.parameters())
            source_params += list(self.q2_network.parameters())
        optimizers.append(
            SoftUpdate.make_optimizer_scheduler(target_params, source_params, tau=self.tau))
        return optimizers

    # ---- kernel argument fillers --------------------------------------------------
    def _fill_critic(self, a, keep):
        dev = self._ws["dev"]
        if self._alpha_dev.device != dev:  # trainer built from CUDA networks, never .cuda()'d
            self._alpha_dev = self._alpha_dev.to(dev)
            if self.alpha_optimizer is not None and self.log_alpha.device != dev:
                raise _lib.Rb200Error("SACTrainer: log_alpha is not on the networks' device -- "
                                      "move the trainer with .cuda()/.to(device) before "
                                      "configure_optimizers()")
        a.alpha = _lib.ptr(self._alpha_dev, dev)
        a.target_entropy = float(self.target_entropy)
        a.backprop_through_log_prob = int(bool(self.backprop_through_log_prob))

    def _fill_actor(self, a, keep):
        self._fill_critic(a, keep)
        ws = self._ws
        B = ws["B"]
        A = self.q1_network.arena.dims[0] - self.actor_network.arena.dims[0]
        nz = self._noise("cur", B, A, ws["dev"])
        keep.append(nz)
        a.noise_cur = nz.data_ptr()
        if self.alpha_optimizer is not None:
            a.alpha_grad = ws["alpha_grad"].data_ptr()
            a.log_alpha = _lib.ptr(self.log_alpha.data, ws["dev"])

    def _alpha_arena(self):
        arena = getattr(self.log_alpha, "_rb200_arena", None)
        if arena is None:
            arena = ScalarArena(self.log_alpha)
        return arena

    # ---- reference protocol -----------------------------------------------------------
    def train_step_gen(self, training_batch: rlt.PolicyNetworkInput, batch_idx: int):
        """IMPORTANT: the input action is assumed to match the actor's output range."""
        assert isinstance(training_batch, rlt.PolicyNetworkInput)
        closs = self._critic_step(training_batch, self.actor_network, self.q1_network_target,
                                  self.q2_network_target, self._fill_critic)
        yield self.fused_loss(closs[0])
        if self.q2_network:
            yield self.fused_loss(closs[1])
        aloss = self._actor_step(training_batch, self._fill_actor)
        yield self.fused_loss(aloss[0])
        if self.alpha_optimizer is not None:
            arena = self._alpha_arena()
            arena.gpart = self._ws["alpha_grad"]
            arena.grad_ready = True
            yield self.fused_loss(aloss[1])
            # sac_trainer.py:322 (runs after the alpha step, used from the next batch on)
            self._alpha_dev.copy_(self.log_alpha.data.exp())
            self.entropy_temperature = self._alpha_dev
        if self.logger:
            self.logger.log_metrics(
                {"td_loss": closs[0], "q1_value": self._ws["q1_value"].mean(),
                 "entropy_temperature": self.entropy_temperature,
                 "target_q_value": self._ws["td_target"].mean(), "actor_loss": aloss[0]},
                step=self.all_batches_processed)
        result = self.soft_update_result()
        self.log("td_loss", closs[0], prog_bar=True)
        yield result

    def train_batch(self, training_batch: rlt.PolicyNetworkInput, batch_idx: int = 0,
                    process_group=None):
        """Fast path: the whole update (same arithmetic as train_step_gen), Polyak updates
        fused into the critics' Adam launches, exp(log_alpha) into the alpha launch."""
        opts = self.optimizers()
        i = 0
        closs = self._critic_step(training_batch, self.actor_network, self.q1_network_target,
                                  self.q2_network_target, self._fill_critic)
        self._dp_step(opts[i], self.q1_network.arena, self.q1_network_target.arena, process_group)
        i += 1
        if self.q2_network:
            self._dp_step(opts[i], self.q2_network.arena, self.q2_network_target.arena,
                          process_group)
            i += 1
        aloss = self._actor_step(training_batch, self._fill_actor)
        self._dp_step(opts[i], self.actor_network.arena, None, process_group)
        i += 1
        if self.alpha_optimizer is not None:
            arena = self._alpha_arena()
            arena.gpart = self._ws["alpha_grad"]
            arena.grad_ready = True
            self._dp_step(opts[i], arena, None, process_group, exp_out=self._alpha_dev)
            self.entropy_temperature = self._alpha_dev
        self.all_batches_processed += 1
        return closs, aloss

    def _dp_step(self, opt, arena, target, process_group, exp_out=None):
        from .data_parallel import dp_fused_step

        dp_fused_step(opt, arena, process_group, target=target, tau=self.tau, exp_out=exp_out)
