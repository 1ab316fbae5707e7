"""Where the online fused step spends its time: the device-resident replay kernels one by one
(CUDA events, eager launches) and the host / device split of FusedDqnStep(rng='device', online=True)."""
import os, random, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reagent_b200.replay_memory import PrioritizedReplayBuffer
from reagent_b200.replay_memory.device_replay import DeviceReplay
from reagent_b200.training.fused_step import FusedDqnStep

cfg = bench.CONFIGS[2]
dev = torch.device("cuda", 0)
B, A = cfg["B"], cfg["A"]
rb = PrioritizedReplayBuffer(1, cfg["cap"], B, device=dev)
rb.add_batch(**bench.synth_stream(cfg["cap"], 1000, cfg))
trainer = bench.build_trainer(cfg, dev)
extra = bench.synth_stream(4096, 555, cfg)
random.seed(1)


def timeit(fn, n=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(10 + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print("host-rng sample_discrete_dqn_batch us", timeit(lambda i: rb.sample_discrete_dqn_batch(B, A)))
dr = DeviceReplay(rb)
idx = torch.zeros(B, dtype=torch.int64, device=dev)
print("draw_indices us", timeit(lambda i: dr.draw_indices(B, out=idx)))
print("gather (given indices) us", timeit(lambda i: rb.sample_discrete_dqn_batch(B, A, indices=idx)))
def add1(i):
    dr.stage(0, 0, **{k: v[i] for k, v in extra.items()})
    dr.launch_add(1)
print("stage + H2D + add kernel us", timeit(add1))
print("add kernel only us", timeit(lambda i: dr.launch_add(1)))
t0 = time.perf_counter()
for i in range(2000): dr.stage(0, 0, **{k: v[i] for k, v in extra.items()})
print("stage() host us", (time.perf_counter() - t0) / 2000 * 1e6)
dr.sync_to_host()
for mode in (dict(rng="host"), dict(rng="device"), dict(rng="device", online=True)):
    f = FusedDqnStep(trainer, rb, B, prefetch=True, **mode)
    fn = (lambda i: f.step({k: v[i] for k, v in extra.items()})) if mode.get("online") else (lambda i: f.step())
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(300): fn(10 + i)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(mode, "host-enqueue us/step %.1f  wall us/step %.1f" % (t_host / 300 * 1e6, t_all / 300 * 1e6))
    if f.dr is not None:
        f.dr.sync_to_host()
    del f
