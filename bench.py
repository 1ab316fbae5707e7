#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE config 2:

  minibatch TD-updates/sec, Discrete DQN (double-Q, huber), synthetic transitions
  state_dim=128, actions=16, batch=4096, prioritized replay (sum tree, capacity 2^20).

One "step" = one full update INCLUDING drawing the minibatch: replay sample kernel (tree
walk + gather + batch formatting) -> fused TD-target/loss/backward kernel (tcgen05) ->
weight-gradient kernel -> fused Adam + soft-target-update kernel (which also writes the hi/lo
TF32 weight images the next TD step feeds to the tensor cores).

  value : K updates captured in ONE CUDA graph with all random numbers already in HBM
          (device-timed, CUDA events, max over ranks).
  e2e   : the same K updates through the public API (FusedDqnStep.step()): per update the
          host draws the stratified query values from Python's `random` (bit-exact with the
          reference), copies them host->device from pinned memory, runs the update and copies
          the loss device->host.
  roofline     : the fused TD kernel (dqn_td_tc_kernel: tcgen05 / TMEM; dqn_td_rows_kernel when
                 the shapes do not fit it), algorithmic FLOPs / measured duration.
  cpu_baseline : the CPU oracle (restatement of the reference's sampler + DQNTrainer update,
                 torch fp32 on all host cores) on a bounded number of updates.

N > 1 (torchrun): weak scaling -- every rank owns a replay shard and a 4096-row minibatch of
a 4096*N global minibatch; ONE all-reduce of the flat gradient per update; `value` counts
4096-row minibatch updates per second over the whole job.

`--impl reference` times the reference algorithm's CPU path (the oracle port: /root/reference
does not exist on the GPU box) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, A, B, CAP = 128, 16, 4096, 1 << 20
SIZES, ACTS = [256, 128], ["relu", "relu"]
GAMMA, TAU, LR = 0.99, 0.005, 1e-3
METRIC = "minibatch_td_updates_per_sec_b4096"
WORKLOAD = ("configs[1]: Discrete DQN double-Q huber, synthetic S=128 A=16 B=4096, "
            "prioritized replay cap=2^20, MLP 128-256-128-16 relu, Adam 1e-3, tau 0.005")


def sigma_net():
    dims = [S] + SIZES + [A]
    return sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


def td_kernel_flops():
    """Algorithmic FLOPs of one dqn_td_rows_kernel launch (SURVEY.md 8d K2 + K2'):
    3 forwards (q(s'), q_target(s'), q(s)) + the dX chain of the backward (all layers but
    the first)."""
    dims = [S] + SIZES + [A]
    fwd = 2 * B * sigma_net()
    bwd_dx = 2 * B * sum(dims[i] * dims[i + 1] for i in range(1, len(dims) - 1))
    return 3 * fwd + bwd_dx


def synth_stream(n, seed):
    import numpy as np

    rng = np.random.RandomState(seed)
    return dict(observation=rng.standard_normal((n, S)).astype(np.float32),
                action=rng.randint(0, A, n).astype(np.int64),
                reward=rng.standard_normal(n).astype(np.float32),
                terminal=rng.rand(n) < (1.0 / 200.0),
                priority=rng.uniform(0.1, 10.0, n))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm, reasons = [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                out["sm_max_mhz"] = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ---------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, threads=None):
    """The reference algorithm on the host: PER sample (python loops over an fp64 sum tree,
    as reagent/replay_memory does) + DQNTrainer update (torch fp32).  Returns (updates/s,
    cores, sample description, ms per step)."""
    import numpy as np
    import torch

    from oracle import td_oracle as O
    from oracle.replay_oracle import ReplayOracle

    cores = threads or os.cpu_count()
    torch.set_num_threads(cores)
    cap = CAP
    st = synth_stream(cap, 0)
    ro = ReplayOracle(cap, prioritized=True)
    # bulk fill (timing only: tree built bottom-up instead of 2^20 sequential set() calls)
    ro.store = {k: v for k, v in st.items() if k != "priority"}
    ro.add_count = cap
    ro.valid[:] = True
    ro.valid[cap - 1] = bool(st["terminal"][cap - 1])
    lvl = st["priority"].astype(np.float64).copy()
    for l in range(len(ro.tree.nodes) - 1, -1, -1):
        ro.tree.nodes[l][: len(lvl)] = lvl
        lvl = lvl.reshape(-1, 2).sum(1) if len(lvl) > 1 else lvl
    gen = torch.Generator().manual_seed(0)
    q = O.make_net([S] + SIZES + [A], ACTS + ["linear"], gen)
    for t in O.net_params(q):
        t.requires_grad_(True)
    qt = O.clone_net(q)
    adam = O.AdamState(O.net_params(q), lr=LR)

    def one():
        ob = ro.sample_transition_batch(B)
        batch = dict(
            state=torch.from_numpy(ob["state"]), next_state=torch.from_numpy(ob["next_state"]),
            reward=torch.from_numpy(ob["reward"]).reshape(-1, 1),
            not_terminal=1.0 - torch.from_numpy(ob["terminal"]).float().reshape(-1, 1),
            action=torch.nn.functional.one_hot(torch.from_numpy(ob["action"]), A).float(),
            possible_next_actions_mask=torch.ones(B, A), next_action=None)
        return O.dqn_update(q, qt, adam, batch, gamma=GAMMA, tau=TAU, loss="huber")[0]

    if threads is None:
        # "all the host threads it can use": small GEMMs get SLOWER when oversubscribed, so
        # give the reference its best thread count from a short sweep (1 update each)
        best = None
        for c in sorted({os.cpu_count(), 64, 32, 16, 8}, reverse=True):
            if c > os.cpu_count():
                continue
            torch.set_num_threads(c)
            one()
            t0 = time.perf_counter()
            one()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, c)
        cores = best[1]
        torch.set_num_threads(cores)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return (steps / dt, cores,
            f"{steps} full updates (PER sample B={B} + DQN update) after {warmup} warm-up; "
            f"torch threads={cores} (best of a sweep over <= {os.cpu_count()} cores)", dt / steps * 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 40)
    warm = min(args.warmup, 3)
    v, cores, sample, ms = cpu_reference_run(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference algorithm restated on CPU "
                   "(oracle/replay_oracle.py + oracle/td_oracle.py): /root/reference is not on "
                   "the GPU box"},
        "cpu_baseline": {"value": v, "unit": "updates/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_ours(args):
    import random

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    from reagent_b200.core.parameters import EvaluationParameters, RLParameters
    from reagent_b200.models import FullyConnectedDQN
    from reagent_b200.optimizer import Optimizer__Union
    from reagent_b200.replay_memory import PrioritizedReplayBuffer
    from reagent_b200.training import DQNTrainer
    from reagent_b200.training.fused_step import FusedDqnStep, capture_device_only

    K, W = args.steps, max(args.warmup, 3)
    rb = PrioritizedReplayBuffer(stack_size=1, replay_capacity=CAP, batch_size=B, device=dev)
    rb.add_batch(**synth_stream(CAP, 1000 + rank))
    torch.manual_seed(0)  # identical initial weights on every rank (replicated parameters)
    q = FullyConnectedDQN(S, A, SIZES, ACTS)
    qt = q.get_target_network()
    trainer = DQNTrainer(
        q.to(dev), qt.to(dev), actions=[str(i) for i in range(A)],
        rl=RLParameters(gamma=GAMMA, target_update_rate=TAU, q_network_loss="huber"),
        double_q_learning=True, minibatch_size=B, optimizer=Optimizer__Union.default(lr=LR),
        evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
    random.seed(1234 + rank)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist

        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- e2e: public API, host RNG -> pinned -> H2D, loss D2H every update ----
    fused = FusedDqnStep(trainer, rb, B, process_group=pg, prefetch=True)
    for _ in range(W):
        fused.step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        loss_host = fused.step()
    e1.record()
    barrier()
    t_host = time.perf_counter() - t_host0
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), t_host * 1e3))
    last_loss = float(loss_host[0])
    e2e_value = world * K / (e2e_ms * 1e-3)

    # ---- roofline of the fused TD kernel: events around each launch, same workload ----
    trainer._kernel_events = []
    for _ in range(W):
        trainer.train_batch(rb.sample_discrete_dqn_batch(B, A), process_group=pg)
    trainer._kernel_events = []
    nroof = min(K, 100)
    for _ in range(nroof):
        batch = rb.sample_discrete_dqn_batch(B, A)
        trainer.tc_prepack()  # keep the weight packing out of the event pair: TD kernel only
        # keep the stream busy while the host enqueues, otherwise the event pair would also
        # time the launch latency of an idle GPU
        torch.cuda._sleep(400_000)
        trainer.train_batch(batch, process_group=pg)
    torch.cuda.synchronize()
    on_tc = trainer._last_td_call[-1] is not None
    durs = [a.elapsed_time(b) for a, b in trainer._kernel_events]
    trainer._kernel_events = None
    kern_ms = sum(durs) / len(durs)

    # ---- value: K updates in one graph, random numbers resident in HBM ----
    def draw(n):
        out = np.empty((n, B), dtype=np.float64)
        for i in range(n):
            qv, pos, _ = rb.host_queries(B)
            # strata that would hit the not-yet-valid slot are redrawn (retry path is host
            # logic, timed in e2e); keeps the captured graph free of overrides
            while pos:
                qv, pos, _ = rb.host_queries(B)
            out[i] = qv
        return torch.from_numpy(out).to(dev)

    q_warm, q_timed = draw(W), draw(K)
    g_warm = capture_device_only(trainer, rb, B, W, q_warm, pg)
    g_timed = capture_device_only(trainer, rb, B, K, q_timed, pg)
    g_warm.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g_timed.replay()
    e1.record()
    barrier()
    clk = clocks.stop()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * K / (dev_ms * 1e-3)

    # ---- cpu baseline (rank 0, N == 1 only).  Runs LAST: its worker threads would otherwise
    # keep spinning on the host cores while the e2e loop (host-paced) is being timed ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_run(args.cpu_steps, 2)
        cpu = {"value": v, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample}

    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF/s sustained"
    flops = td_kernel_flops()
    achieved_tf = flops / (kern_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world,
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "inputs larger than L2: 512 MiB replay store, random row gather per update",
                   "value_path": "K updates in one CUDA graph, query values resident in HBM",
                   "final_loss": last_loss},
        "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": fused.h2d_bytes,
                "d2h_bytes_per_step": fused.d2h_bytes, "ms_per_step": e2e_ms / K,
                "api": "reagent_b200.training.fused_step.FusedDqnStep(prefetch=True).step(): every "
                       "step draws one minibatch (host RNG -> pinned -> H2D -> sample kernel) and "
                       "trains on one; the sampler runs one update ahead on a second stream"},
        # sample, (weight images unless Adam wrote them), TD step, weight gradients, Adam+Polyak
        "gpu_launches": (4 if (not on_tc or os.environ.get("RB200_ADAM_PACK", "1") == "1") else 5) * K,
        "clocks": clk,
        "roofline": {"kernel": ("dqn_td_tc_kernel (fused TD target + loss + dZ chain on tcgen05/TMEM)"
                                if on_tc else
                                "dqn_td_rows_kernel (fused TD target + loss + dZ chain, mma.sync)"),
                     "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": 6195200 if on_tc else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu "
                                       "--set full launch (profiles/r01_ncu_dqn_td_tc.csv), bytes",
                     "peak_source": peak_src, "algorithmic_flops_per_launch": flops,
                     "kernel_ms": kern_ms,
                     "pipe_used": ("tcgen05.mma kind::tf32, 3xTF32 as 2 MMAs per k step (N=64 + N=32)"
                                   if on_tc else "mma.sync m16n8k8 tf32, 3xTF32"),
                     "executed_over_algorithmic_flops": 3.0,
                     "note": "fp32-parity (1e-5) forces 3xTF32: 3 tensor-core flops per "
                             "algorithmic flop, and TF32 dense peak is half the bf16 peak this "
                             "fraction is quoted against"},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_ours(args)
        finally:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
