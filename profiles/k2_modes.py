"""K2 timing with parts of the kernel switched off (profiling build only; results are wrong
by construction, the timing tells which resource the step is waiting for).
Usage: RB200_LIB=reagent_b200/libreagent_b200_timeline.so python profiles/k2_modes.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reagent_b200 import _lib
from reagent_b200.core import types as rlt
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.training import DQNTrainer

dev = torch.device("cuda", 0)
S, A, B = 128, 16, 4096
torch.manual_seed(0)
q = FullyConnectedDQN(S, A, [256, 128], ["relu", "relu"]); qt = q.get_target_network()
t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(A)],
               rl=RLParameters(gamma=0.99, target_update_rate=0.005, q_network_loss="huber"),
               optimizer=Optimizer__Union.default(lr=1e-3),
               evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
act = torch.randint(A, (B,), device=dev)
batch = rlt.DiscreteDqnInput(
    state=rlt.FeatureData(torch.randn(B, S, device=dev)), next_state=rlt.FeatureData(torch.randn(B, S, device=dev)),
    reward=torch.randn(B, 1, device=dev), time_diff=None, step=None,
    not_terminal=torch.ones(B, 1, device=dev), action=torch.nn.functional.one_hot(act, A).float(),
    next_action=torch.nn.functional.one_hot(act, A).float(),
    possible_actions_mask=torch.ones(B, A, device=dev), possible_next_actions_mask=torch.ones(B, A, device=dev),
    extras=rlt.ExtraData())
t._td_step(batch)
qd, qtd, a, wsc, keep, pack = t._last_td_call
st = _lib.cur_stream()


def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


NAMES = {1: "noMMA32", 2: "noMMA64", 4: "noRingRead", 8: "noCopy", 16: "noTmemSt", 32: "noEpiStores"}
for mode in (0, 4, 8, 12, 16, 28, 32, 60, 3, 63):
    os.environ["RB200_TC_DBG_MODE"] = str(mode)
    us = timeit(lambda: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 1, st))
    print(json.dumps({"mode": mode, "off": [v for k, v in NAMES.items() if mode & k], "k2_us": round(us, 2)}), flush=True)
