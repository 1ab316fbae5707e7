"""Small driver for ncu: a few full DQN updates at BASELINE config 2 (eager launches)."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from reagent_b200.core.parameters import EvaluationParameters, RLParameters  # noqa: E402
from reagent_b200.models import FullyConnectedDQN  # noqa: E402
from reagent_b200.optimizer import Optimizer__Union  # noqa: E402
from reagent_b200.replay_memory import PrioritizedReplayBuffer  # noqa: E402
from reagent_b200.training import DQNTrainer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
rb = PrioritizedReplayBuffer(1, bench.CAP, bench.B, device=dev)
rb.add_batch(**bench.synth_stream(bench.CAP, 1000))
torch.manual_seed(0)
q = FullyConnectedDQN(bench.S, bench.A, bench.SIZES, bench.ACTS)
qt = q.get_target_network()
t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(bench.A)],
               rl=RLParameters(gamma=bench.GAMMA, target_update_rate=bench.TAU, q_network_loss="huber"),
               optimizer=Optimizer__Union.default(lr=bench.LR),
               evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
random.seed(0)
for _ in range(n):
    t.train_batch(rb.sample_discrete_dqn_batch(bench.B, bench.A))
torch.cuda.synchronize()
print("done", float(t._ws["loss"]))
