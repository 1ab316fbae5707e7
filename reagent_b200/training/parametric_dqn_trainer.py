"""ParametricDQNTrainer (reagent/training/parametric_dqn_trainer.py:22-214) on the generic
kernels of this library: the critic-shaped q_network(state, action) is evaluated by
rb200_mlp_forward (on the tiled possible next actions for max-Q learning), rb200_pdqn_head turns
the values into the TD loss and d loss / d q, rb200_mlp_backward + rb200_mlp_wgrad produce the
parameter gradients, FusedAdam / SoftUpdate apply them.  The tiling and the concatenation of
(state, action) are torch plumbing on device tensors."""
from typing import Optional

import torch

from .. import _lib
from ..core import types as rlt
from ..core.parameters import RLParameters
from ..optimizer import Optimizer__Union, SoftUpdate
from .reagent_lightning_module import ReAgentLightningModule
from .rl_trainer_pytorch import RLTrainerMixin
from .workspace import NetWorkspace, param_grads, wgrad


class ParametricDQNTrainer(RLTrainerMixin, ReAgentLightningModule):
    def __init__(self, q_network, q_network_target, reward_network=None,
                 rl: Optional[RLParameters] = None, double_q_learning: bool = True,
                 minibatches_per_step: int = 1, optimizer: Optional[Optimizer__Union] = None,
                 log_tensorboard: bool = False) -> None:
        super().__init__()
        self.rl_parameters = RLParameters() if rl is None else rl
        self.double_q_learning = double_q_learning
        self.minibatches_per_step = minibatches_per_step or 1
        self.q_network = q_network
        self.q_network_target = q_network_target
        self.reward_network = reward_network
        self.optimizer = Optimizer__Union.default() if optimizer is None else optimizer
        self.log_tensorboard = log_tensorboard
        loss = self.rl_parameters.q_network_loss
        if loss == "mse":
            self.q_network_loss_kind = _lib.LOSS_MSE
        elif loss == "huber":
            self.q_network_loss_kind = _lib.LOSS_HUBER
        elif loss == "bce_with_logits":
            raise NotImplementedError("bce_with_logits (gamma == 0 only) has no fused head")
        else:
            raise Exception("Q-Network loss type {} not valid loss.".format(loss))
        self._ws = None

    def configure_optimizers(self):
        """[Adam(q_network), (Adam(reward_network),) SoftUpdate] -- :66-86."""
        optimizers = [self.optimizer.make_optimizer_scheduler(self.q_network.parameters())]
        if self.reward_network is not None:
            optimizers.append(self.optimizer.make_optimizer_scheduler(self.reward_network.parameters()))
        optimizers.append(SoftUpdate.make_optimizer_scheduler(
            list(self.q_network_target.parameters()), list(self.q_network.parameters()), tau=self.tau))
        return optimizers

    def _check_input(self, training_batch: rlt.ParametricDqnInput):
        assert isinstance(training_batch, rlt.ParametricDqnInput)
        assert training_batch.not_terminal.dim() == training_batch.reward.dim() == 2
        assert training_batch.not_terminal.shape[1] == training_batch.reward.shape[1] == 1
        assert (training_batch.action.float_features.dim()
                == training_batch.next_action.float_features.dim() == 2)

    # ------------------------------------------------------------------
    def _workspace(self, B, device):
        ws = self._ws
        if ws is None or ws["B"] != B or ws["dev"] != device:
            ws = {"B": B, "dev": device,
                  "q": NetWorkspace(self.q_network.arena, B, device),
                  "r": (None if self.reward_network is None
                        else NetWorkspace(self.reward_network.arena, B, device)),
                  "q_values": torch.empty(B, 1, device=device),
                  "td_target": torch.empty(B, device=device),
                  "loss_partials": torch.zeros((B + 255) // 256, device=device),
                  "loss": torch.zeros(1, device=device),
                  "r_loss": torch.zeros(1, device=device),
                  "counter": torch.zeros(1, dtype=torch.int32, device=device)}
            self._ws = ws
        return ws

    @staticmethod
    def _fwd(net, x, save=None):
        out = torch.empty(x.shape[0], net.arena.dims[-1], device=x.device)
        rc = _lib.lib().rb200_mlp_forward(net.arena.desc(), x.data_ptr(), x.shape[1], None, 0,
                                          x.shape[0], out.data_ptr(), save, _lib.cur_stream())
        _lib.check(rc, "rb200_mlp_forward")
        return out

    def _backward(self, net, w, x, B):
        ar = net.arena
        L = len(ar.acts)
        rc = _lib.lib().rb200_mlp_backward(ar.desc(), w.dz[L - 1].data_ptr(), B, w.c, _lib.cur_stream())
        _lib.check(rc, "rb200_mlp_backward")
        wgrad(ar, w, x, B)

    def _td_step(self, batch: rlt.ParametricDqnInput) -> torch.Tensor:
        state = batch.state.float_features.float().contiguous()
        if not state.is_cuda:
            raise _lib.Rb200Error("ParametricDQNTrainer: training batch must be on the GPU (no CPU path)")
        dev, B = state.device, state.shape[0]
        _lib.require_current_device(dev)
        ws = self._workspace(B, dev)
        keep = []

        def P(t):
            t = _lib.on_device(t.float().contiguous(), dev)
            keep.append(t)
            return _lib.ptr(t, dev)

        a = _lib.PdqnArgsT()
        a.batch = B
        next_state = batch.next_state.float_features.float()
        if self.maxq_learning:
            pna = batch.possible_next_actions.float_features.float()
            product = pna.shape[0]
            assert product % B == 0, f"batch_size * max_num_action {product} is not divisible by batch_size {B}"
            M = product // B
            # FeatureData.get_tiled_batch: row i repeated M times, then cat with the actions
            x_next = torch.cat((next_state.repeat_interleave(M, dim=0), pna), dim=1).contiguous()
            nq_t = self._fwd(self.q_network_target, x_next)
            nq = self._fwd(self.q_network, x_next) if self.double_q_learning else None
            keep += [x_next, nq_t, nq]
            a.max_num_action = M
            a.next_q = None if nq is None else nq.data_ptr()
            a.next_q_target = nq_t.data_ptr()
            a.mask = P(batch.possible_next_actions_mask)
        else:  # SARSA on the target network
            x_next = torch.cat((next_state, batch.next_action.float_features.float()), dim=1).contiguous()
            nq_t = self._fwd(self.q_network_target, x_next)
            keep += [x_next, nq_t]
            a.max_num_action = 0
            a.next_q_target = nq_t.data_ptr()
        a.reward = P(batch.reward.reshape(-1))
        a.not_terminal = P(batch.not_terminal.reshape(-1))
        a.gamma = float(self.gamma)
        a.discount_mode = _lib.DISCOUNT_CONST
        if self.use_seq_num_diff_as_time_diff:
            assert self.multi_steps is None
            a.discount_src, a.discount_mode = P(batch.time_diff.reshape(-1)), _lib.DISCOUNT_POW
        if self.multi_steps is not None:
            a.discount_src, a.discount_mode = P(batch.step.reshape(-1)), _lib.DISCOUNT_POW
        a.double_q = int(bool(self.double_q_learning))
        a.loss_kind = self.q_network_loss_kind
        x = torch.cat((state, batch.action.float_features.float()), dim=1).contiguous()
        self._x = x
        qv = ws["q_values"]
        rc = _lib.lib().rb200_mlp_forward(self.q_network.arena.desc(), x.data_ptr(), x.shape[1], None, 0,
                                          B, qv.data_ptr(), ws["q"].c, _lib.cur_stream())
        _lib.check(rc, "rb200_mlp_forward")
        L = len(self.q_network.arena.acts)
        a.q_values = qv.data_ptr()
        a.dz = ws["q"].dz[L - 1].data_ptr()
        a.td_target = ws["td_target"].data_ptr()
        a.loss_partials = ws["loss_partials"].data_ptr()
        a.loss = ws["loss"].data_ptr()
        a.tile_counter = ws["counter"].data_ptr()
        _lib.check(_lib.lib().rb200_pdqn_head(a, _lib.cur_stream()), "rb200_pdqn_head")
        self._backward(self.q_network, ws["q"], x, B)
        return ws["loss"].reshape(())

    def _reward_step(self, batch: rlt.ParametricDqnInput) -> torch.Tensor:
        """mse(reward_network(state, action), cat(reward, metrics)) -- :176-190."""
        ws = self._ws
        x, B = self._x, self._x.shape[0]
        metrics = batch.extras.metrics if batch.extras is not None else None
        mrc = batch.reward if metrics is None else torch.cat((batch.reward, metrics), dim=1)
        w = ws["r"]
        est = torch.empty(B, self.reward_network.arena.dims[-1], device=x.device)
        rc = _lib.lib().rb200_mlp_forward(self.reward_network.arena.desc(), x.data_ptr(), x.shape[1],
                                          None, 0, B, est.data_ptr(), w.c, _lib.cur_stream())
        _lib.check(rc, "rb200_mlp_forward")
        diff = est - mrc.float()
        L = len(self.reward_network.arena.acts)
        w.dz[L - 1].copy_(diff * (2.0 / diff.numel()))
        ws["r_loss"].copy_((diff * diff).mean().reshape(1))
        self._backward(self.reward_network, w, x, B)
        return ws["r_loss"].reshape(())

    # ------------------------------------------------------------------
    def train_step_gen(self, training_batch: rlt.ParametricDqnInput, batch_idx: int):
        self._check_input(training_batch)
        td_loss = self._td_step(training_batch)
        yield self.fused_loss(td_loss)
        if self.reward_network is not None:
            yield self.fused_loss(self._reward_step(training_batch))
        yield self.soft_update_result()

    def train_batch(self, training_batch: rlt.ParametricDqnInput, batch_idx: int = 0,
                    process_group=None):
        from .data_parallel import dp_fused_step

        opts = self.optimizers()
        self._td_step(training_batch)
        dp_fused_step(opts[0], self.q_network.arena, process_group,
                      target=self.q_network_target.arena, tau=self.tau)
        if self.reward_network is not None:
            self._reward_step(training_batch)
            dp_fused_step(opts[1], self.reward_network.arena, process_group)
        self.all_batches_processed += 1
        return self._ws["loss"]

    def q_network_grads(self):
        return param_grads(self.q_network.arena, list(self.q_network.parameters()))
