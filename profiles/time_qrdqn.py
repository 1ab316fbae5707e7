"""QR-DQN update at BASELINE config 3 (S=128, A=32, N=200 atoms, B=4096): per-launch timing."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reagent_b200 import _lib
from reagent_b200.core import types as rlt
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.training import QRDQNTrainer

S, A, N, B = 128, 32, 200, 4096
dev = torch.device("cuda", 0)
torch.manual_seed(0)
q = FullyConnectedDQN(S, A, [256, 128], ["relu", "relu"], num_atoms=N)
t = QRDQNTrainer(q.to(dev), q.get_target_network().to(dev), actions=[str(i) for i in range(A)],
                 rl=RLParameters(gamma=0.99, target_update_rate=0.005), num_atoms=N,
                 optimizer=Optimizer__Union.default(lr=1e-3),
                 evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
act = torch.randint(A, (B,), device=dev)
batch = rlt.DiscreteDqnInput(
    state=rlt.FeatureData(torch.randn(B, S, device=dev)), next_state=rlt.FeatureData(torch.randn(B, S, device=dev)),
    reward=torch.randn(B, 1, device=dev), time_diff=None, step=None,
    not_terminal=(torch.rand(B, 1, device=dev) > 0.005).float(),
    action=torch.nn.functional.one_hot(act, A).float(), next_action=torch.nn.functional.one_hot(act, A).float(),
    possible_actions_mask=torch.ones(B, A, device=dev), possible_next_actions_mask=torch.ones(B, A, device=dev),
    extras=rlt.ExtraData())
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("full QR-DQN update us", timeit(lambda: t.train_batch(batch)))
ws = t._ws; qa = t.q_network.arena; L = 3
h = ws["net"].hidden[L - 2]; out = ws["q_cur"]; flat = qa.flat
lib, st = _lib.lib(), _lib.cur_stream()
args = (flat.data_ptr() + 4 * qa.w_off[L - 1], flat.data_ptr() + 4 * qa.b_off[L - 1], 0, 128, A * N, h.data_ptr(), B, out.data_ptr(), st)
us = timeit(lambda: lib.rb200_linear_forward(*args), 50)
print("head fwd (tcgen05) us %.1f  -> %.1f TFLOP/s algorithmic" % (us, 2 * B * 128 * A * N / us / 1e6))
os.environ["X"] = "1"
us2 = timeit(lambda: lib.rb200_linear_backward_dx(flat.data_ptr() + 4 * qa.w_off[L - 1], 128, A * N, ws["net"].dz[L - 1].data_ptr(), h.data_ptr(), 1, B, ws["net"].dz[L - 2].data_ptr(), st), 20)
print("head bwd dX (mma.sync rows) us %.1f" % us2)
from reagent_b200.training.workspace import wgrad
print("wgrad (all layers) us %.1f" % timeit(lambda: wgrad(qa, ws["net"], batch.state.float_features, B), 20))
