"""Diagnostic: where does the SAC config-4 critic step deviate from the fp32 oracle?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import td_oracle as O
from tests import golden_util as G
from tests.test_actor_critic_gpu import _rand_net, _net_arrays, _build_sac, _pbatch, _inject

S, A, B = 256, 32, 2048
meta = dict(S=S, A=A, B=B, sizes=[256, 256], acts=["relu", "relu"], twin=True, learn_alpha=True,
            gamma=0.99, tau=0.005, lr=1e-3, entropy_temperature=0.1, target_entropy=-float(A),
            backprop=True, n_updates=2)
gen = torch.Generator().manual_seed(0)
actor = _rand_net([S, 256, 256, 2 * A], ["relu", "relu", "linear"], gen)
q1 = _rand_net([S + A, 256, 256, 1], ["relu", "relu", "linear"], gen)
q2 = _rand_net([S + A, 256, 256, 1], ["relu", "relu", "linear"], gen)
arrays = {}
_net_arrays(arrays, "actor0", actor); _net_arrays(arrays, "q1_0", q1); _net_arrays(arrays, "q2_0", q2)
b = dict(state=torch.randn(B, S, generator=gen), next_state=torch.randn(B, S, generator=gen),
         action=torch.rand(B, A, generator=gen) * 1.98 - 0.99, next_action=torch.zeros(B, A),
         reward=torch.randn(B, 1, generator=gen), not_terminal=(torch.rand(B, 1, generator=gen) > 0.005).float())
t = _build_sac(meta, arrays)
nn_, nc = torch.randn(B, A, generator=gen), torch.randn(B, A, generator=gen)
arrays["noise0.next"], arrays["noise0.cur"] = nn_.numpy(), nc.numpy()
_inject(t, arrays, 0)
gb = _pbatch({k: v.cuda() for k, v in b.items()})
# oracle intermediates
a_next, logp = O.gaussian_actor_forward(actor, b["next_state"], nn_)
out = O.mlp(actor, b["next_state"])
closs = t._critic_step(gb, t.actor_network, t.q1_network_target, t.q2_network_target, t._fill_critic)
from reagent_b200.core import types as rlt
gout = t.actor_network.fc(gb.next_state.float_features)
print("actor fc out rel err", G.rel_err(gout, out))
lp_g = t._ws["log_prob"].cpu()
d = (lp_g - logp.reshape(-1)).abs()
print("log_prob abs err: max %.3e  median %.3e  frac>1e-4: %.4f" % (float(d.max()), float(d.median()), float((d > 1e-4).float().mean())))
inside = (logp.reshape(-1).abs() < 2)
print("rows with |logp|<2:", int(inside.sum()), " max err among them %.3e" % float(d[inside].max() if inside.any() else 0))
st = O.SacState(actor, q1, q2, lr=1e-3, entropy_temperature=0.1, learn_alpha=True, target_entropy=-float(A))
o = O.sac_update(st, b, nn_, nc, gamma=0.99, tau=0.005)
dt = (t._ws["td_target"].cpu() - o["target"].reshape(-1)).abs()
print("td_target rel err %.3e, rows off>1e-4: %d" % (float(dt.max() / o["target"].abs().max()), int((dt > 1e-4).sum())))
for pi, g in enumerate(t.net_grads(t.q1_network)):
    print("q1 grad", pi, G.rel_err(g, o["grads"]["q1"][pi]))
print("losses", closs.tolist(), o["losses"][:2])
