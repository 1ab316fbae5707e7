"""Where the host time of FusedDqnStep.step() goes (cProfile over 300 steps, config 2)."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.replay_memory import PrioritizedReplayBuffer
from reagent_b200.training import DQNTrainer
from reagent_b200.training.fused_step import FusedDqnStep

dev = torch.device("cuda", 0)
rb = PrioritizedReplayBuffer(1, bench.CAP, bench.B, device=dev)
rb.add_batch(**bench.synth_stream(bench.CAP, 1000))
torch.manual_seed(0)
q = FullyConnectedDQN(bench.S, bench.A, bench.SIZES, bench.ACTS); qt = q.get_target_network()
t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(bench.A)],
               rl=RLParameters(gamma=bench.GAMMA, target_update_rate=bench.TAU, q_network_loss="huber"),
               optimizer=Optimizer__Union.default(lr=bench.LR),
               evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
fused = FusedDqnStep(t, rb, bench.B, prefetch=True)
for _ in range(20): fused.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): fused.step()
th = time.perf_counter() - t0
torch.cuda.synchronize()
tt = time.perf_counter() - t0
print("host loop us/step %.1f  (incl. final sync %.1f)" % (th / 300 * 1e6, tt / 300 * 1e6))
# device-only time of one graph replay (no host work in between)
s = fused.slots[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): fused.slots[i % 2]["graph"].replay()
e1.record(); torch.cuda.synchronize()
print("graph replay back-to-back us/step %.1f" % (e0.elapsed_time(e1) / 100 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): fused.step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
