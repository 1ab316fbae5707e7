"""Shared machinery of the fused SAC / TD3 trainers: argument marshalling for
rb200_ac_critic_step / rb200_ac_actor_step, workspaces, noise, weight gradients."""
from typing import Optional

import torch

from .. import _lib
from ..core import types as rlt
from .reagent_lightning_module import ReAgentLightningModule
from .rl_trainer_pytorch import RLTrainerMixin
from .workspace import NetWorkspace, param_grads, wgrad


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class ActorCriticBase(RLTrainerMixin, ReAgentLightningModule):
    ALGO = None

    def _ac_init(self):
        self._ws = None
        # noise_hook(name, shape, device) -> tensor lets tests inject the reference's
        # torch.randn_like draws; default: torch.randn on the device
        self.noise_hook = None
        self._kernel_events = None

    def _noise(self, name, B, A, device):
        if self.noise_hook is not None:
            t = self.noise_hook(name, (B, A), device)
            return _f32c(t.to(device))
        return torch.randn(B, A, device=device)

    def _workspace(self, B, device):
        ws = self._ws
        if ws is None or ws["B"] != B or ws["dev"] != device:
            ntiles = (B + 15) // 16
            q2 = self.q2_network
            ws = {
                "B": B, "dev": device,
                "actor": NetWorkspace(self.actor_network.arena, B, device),
                "q1": NetWorkspace(self.q1_network.arena, B, device, need_input=True),
                "q2": None if q2 is None else NetWorkspace(q2.arena, B, device),
                "loss_partials": torch.zeros(2 * ntiles, device=device),
                "critic_loss": torch.zeros(2, device=device),
                "actor_loss": torch.zeros(2, device=device),
                "counter": torch.zeros(1, dtype=torch.int32, device=device),
                "alpha_grad": torch.zeros(1, 1, device=device),
                "td_target": torch.empty(B, device=device),
                "q1_value": torch.empty(B, device=device),
                "q2_value": torch.empty(B, device=device),
                "log_prob": torch.empty(B, device=device),
            }
            if ws["q2"] is not None:
                ws["q2"].c.input = ws["q1"].input.data_ptr()
            self._ws = ws
        return ws

    def _base_args(self, batch: rlt.PolicyNetworkInput, ws, keep):
        state = _f32c(batch.state.float_features)
        if not state.is_cuda:
            raise _lib.Rb200Error(
                f"{type(self).__name__}: training batch must be on the GPU (no CPU path)")
        _lib.require_current_device(state.device)
        a = _lib.AcArgsT()

        def P(t):
            t = _lib.on_device(_f32c(t), state.device)
            keep.append(t)
            return _lib.ptr(t, state.device)

        a.batch = state.shape[0]
        a.algo = self.ALGO
        a.state = P(state)
        a.action = P(batch.action.float_features)
        a.next_state = P(batch.next_state.float_features)
        a.reward = P(batch.reward.reshape(-1))
        a.not_terminal = P(batch.not_terminal.reshape(-1))
        a.gamma = float(self.gamma)
        a.loss_partials = ws["loss_partials"].data_ptr()
        a.tile_counter = ws["counter"].data_ptr()
        return a, state

    def _desc(self, net):
        return None if net is None else net.arena.desc()

    def _critic_step(self, batch, actor_net, q1t, q2t, fill):
        state = batch.state.float_features
        B, dev = state.shape[0], state.device
        ws = self._workspace(B, dev)
        keep = []
        a, state = self._base_args(batch, ws, keep)
        A = self.q1_network.arena.dims[0] - self.actor_network.arena.dims[0]
        nz = self._noise("next", B, A, dev)
        keep.append(nz)
        a.noise_next = nz.data_ptr()
        a.loss = ws["critic_loss"].data_ptr()
        a.td_target = ws["td_target"].data_ptr()
        a.q1_value = ws["q1_value"].data_ptr()
        a.q2_value = ws["q2_value"].data_ptr()
        a.log_prob_out = ws["log_prob"].data_ptr()
        fill(a, keep)
        q2 = self.q2_network
        ev = self._kernel_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _lib.lib().rb200_ac_critic_step(
            self._desc(actor_net), self._desc(self.q1_network), self._desc(q2),
            self._desc(q1t), self._desc(q2t), a, ws["q1"].c,
            None if q2 is None else ws["q2"].c, _lib.cur_stream())
        _lib.check(rc, "rb200_ac_critic_step")
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        wgrad(self.q1_network.arena, ws["q1"], None, B)
        if q2 is not None:
            wgrad(q2.arena, ws["q2"], None, B)
        return ws["critic_loss"]

    def _actor_step(self, batch, fill):
        state = batch.state.float_features
        B, dev = state.shape[0], state.device
        ws = self._workspace(B, dev)
        keep = []
        a, state = self._base_args(batch, ws, keep)
        a.loss = ws["actor_loss"].data_ptr()
        a.log_prob_out = ws["log_prob"].data_ptr()
        fill(a, keep)
        q2 = self.q2_network
        rc = _lib.lib().rb200_ac_actor_step(
            self._desc(self.actor_network), self._desc(self.q1_network), self._desc(q2), a,
            ws["actor"].c, ws["q1"].c, None if q2 is None else ws["q2"].c, _lib.cur_stream())
        _lib.check(rc, "rb200_ac_actor_step")
        wgrad(self.actor_network.arena, ws["actor"], state, B)
        return ws["actor_loss"]

    def net_grads(self, net):
        """Per-parameter gradients of the last fused backward of `net` (tests)."""
        return param_grads(net.arena, list(net.parameters()))
