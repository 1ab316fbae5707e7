"""A/B: mma.sync row-tile K2 vs tcgen05 K2 on identical inputs (outputs + timing)."""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reagent_b200 import _lib
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.replay_memory import PrioritizedReplayBuffer
from reagent_b200.training import DQNTrainer

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B
rb = PrioritizedReplayBuffer(1, bench.CAP, B, device=dev)
rb.add_batch(**bench.synth_stream(bench.CAP, 1000))
torch.manual_seed(0)
q = FullyConnectedDQN(bench.S, bench.A, bench.SIZES, bench.ACTS); qt = q.get_target_network()
with torch.no_grad():
    for p_ in qt.parameters(): p_.add_(0.01 * torch.randn_like(p_))
t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(bench.A)],
               rl=RLParameters(gamma=bench.GAMMA, target_update_rate=bench.TAU, q_network_loss="huber"),
               optimizer=Optimizer__Union.default(lr=bench.LR),
               evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
random.seed(0)
batch = rb.sample_discrete_dqn_batch(B, bench.A)
t._td_step(batch)
qd, qtd, a, wsc, keep, pack = t._last_td_call
assert pack is not None, "tcgen05 path not selected"
st = _lib.cur_stream()
ws = t._ws
def snap():
    torch.cuda.synchronize()
    net = ws["net"]
    return {"loss": ws["loss"].clone(), "scores": ws["scores"].clone(), "tgt": ws["td_target"].clone(),
            "qsel": ws["q_sel"].clone(), "idx": ws["next_idx"].clone(),
            **{f"h{i}": h.clone() for i, h in enumerate(net.hidden)},
            **{f"dz{i}": z.clone() for i, z in enumerate(net.dz)}}
def zero():
    for h in ws["net"].hidden: h.zero_()
    for z in ws["net"].dz: z.zero_()
zero(); _lib.check(_lib.lib().rb200_dqn_td_step(qd, qtd, a, wsc, st), "rows"); r0 = snap()
zero(); _lib.check(_lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 0, st), "tc"); r1 = snap()
for k in r0:
    x, y = r0[k].double(), r1[k].double()
    den = x.abs().max().item() or 1.0
    print(f"{k:7s} max|diff| {float((x - y).abs().max()):.3e}  rel-to-max {float((x - y).abs().max()) / den:.3e}  l2rel {float((x - y).norm() / (x.norm() + 1e-30)):.3e}")
def timeit(fn, n=200):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("rows kernel us", timeit(lambda i: _lib.lib().rb200_dqn_td_step(qd, qtd, a, wsc, st)))
print("tc pack+kernel us", timeit(lambda i: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 0, st)))
print("tc kernel only us", timeit(lambda i: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 1, st)))
print("tc pack only us", timeit(lambda i: _lib.lib().rb200_dqn_tc_pack(qd, qtd, a.double_q, a.do_backward, pack.data_ptr(), pack.numel(), st)))
a.do_backward = 0
print("tc fwd only us", timeit(lambda i: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 0, st)))
# ---- timeline of block 0 (clock64 stamps) ----
import ctypes
a.do_backward = 1
dbg = torch.zeros(32 * 8 + 4 * 4096 + 32 * 8, dtype=torch.int64, device=dev)
_lib.lib().rb200_debug_set_tc_timeline(ctypes.c_void_p(dbg.data_ptr()))
for _ in range(3):
    _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 0, st)
torch.cuda.synchronize()
_lib.lib().rb200_debug_set_tc_timeline(ctypes.c_void_p(0))
d = dbg.cpu()[:256].view(32, 8)
blk = dbg.cpu()[256:256 + 4 * 4096].view(-1, 4)[: (B + 31) // 32]
fine = dbg.cpu()[256 + 4 * 4096:].view(32, 8)
t0 = int(d[0, 0])
print("step  op_ready  issue_end  mma_wait_afull  epi_start  epi_end  arrive   (cycles from first op_ready); loader waits: full / adone")
for s_ in range(32):
    if int(d[s_, 0]) == 0: break
    r = d[s_]
    print(f"{s_:3d} {int(r[0])-t0:9d} {int(r[1])-t0:9d} {int(r[2]):9d} {int(r[3])-t0:9d} {int(r[4])-t0:9d} {int(r[5])-t0 if int(r[5]) else 0:9d}  ld_wait_full {int(r[6]):6d} ld_wait_adone {int(r[7]):6d} | mma: fence {int(fine[s_,0])} issue {int(fine[s_,1])} commit {int(fine[s_,2])} | loader: lds+split {int(fine[s_,3])} st+arrive {int(fine[s_,4])}")
if int(blk[0, 0]):
    st0 = int(blk[:, 0].min())
    print("per-block ns: start spread %d, setup %.0f avg, run avg %.0f max %.0f, last end %d" % (
        int(blk[:, 0].max()) - st0, float((blk[:, 1] - blk[:, 0]).float().mean()),
        float((blk[:, 2] - blk[:, 1]).float().mean()), float((blk[:, 2] - blk[:, 1]).float().max()),
        int(blk[:, 2].max()) - st0))
