"""Quick per-kernel timing (CUDA events, eager launches, fresh random batches)."""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.replay_memory import PrioritizedReplayBuffer
from reagent_b200.training import DQNTrainer
from reagent_b200.training.workspace import wgrad

dev = torch.device("cuda", 0)
rb = PrioritizedReplayBuffer(1, bench.CAP, bench.B, device=dev)
rb.add_batch(**bench.synth_stream(bench.CAP, 1000))
torch.manual_seed(0)
q = FullyConnectedDQN(bench.S, bench.A, bench.SIZES, bench.ACTS); qt = q.get_target_network()
t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(bench.A)],
               rl=RLParameters(gamma=bench.GAMMA, target_update_rate=bench.TAU, q_network_loss="huber"),
               optimizer=Optimizer__Union.default(lr=bench.LR),
               evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
random.seed(0)
batches = [rb.sample_discrete_dqn_batch(bench.B, bench.A) for _ in range(8)]
def timeit(fn, n=50):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
opt = t.optimizers()[0]
from reagent_b200 import _lib
t._td_step(batches[0])
qd, qtd, a, wsc, keep, pack = t._last_td_call
st = _lib.cur_stream()
print("K2 dqn_td_rows kernel (mma.sync) us", timeit(lambda i: _lib.lib().rb200_dqn_td_step(qd, qtd, a, wsc, st), 200))
if pack is not None:
    print("K2 dqn_td_tc kernel (tcgen05, pack+kernel) us", timeit(lambda i: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 0, st), 200))
a.do_backward = 0
print("K2 fwd+loss only (device) us", timeit(lambda i: _lib.lib().rb200_dqn_td_step(qd, qtd, a, wsc, st), 200))
a.do_backward = 1
print("sample  us", timeit(lambda i: rb.sample_discrete_dqn_batch(bench.B, bench.A)))
print("td+wgrad us", timeit(lambda i: t._td_step(batches[i % 8])))
print("td fwd only us", timeit(lambda i: t._td_step(batches[i % 8], do_backward=False)))
ws = t._ws
os.environ["RB200_WGRAD_TC"] = "1"
print("wgrad (tcgen05) us", timeit(lambda i: wgrad(t.q_network.arena, ws["net"], batches[0].state.float_features, bench.B)))
for dbg in (1, 2, 4, 8, 15, 7):
    os.environ["RB200_WT_DBG"] = str(dbg)
    print("  wgrad tcgen05 dbg=%d us" % dbg, timeit(lambda i: wgrad(t.q_network.arena, ws["net"], batches[0].state.float_features, bench.B)))
os.environ["RB200_WT_DBG"] = "0"
os.environ["RB200_WGRAD_TC"] = "0"
print("wgrad (mma.sync) us", timeit(lambda i: wgrad(t.q_network.arena, ws["net"], batches[0].state.float_features, bench.B)))
wgrad(t.q_network.arena, ws["net"], batches[0].state.float_features, bench.B)
print("adam us", timeit(lambda i: (setattr(t.q_network.arena, "grad_ready", True), opt.fused_step(target=t.q_network_target.arena, tau=0.005))))
print("full step us", timeit(lambda i: t.train_batch(rb.sample_discrete_dqn_batch(bench.B, bench.A))))
