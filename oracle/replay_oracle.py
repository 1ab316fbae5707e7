"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / python loops) of the reference's
replay sampler.  Never imported by the product path; used by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs.

PINNED by tests/test_oracle_golden.py against golden vectors produced by the unmodified
reference buffers (oracle/make_golden.py::replay_case).

Restates
  SumTree                     reagent/replay_memory/sum_tree.py:30-189
  ReplayBuffer.add            reagent/replay_memory/circular_replay_buffer.py:468-547
  sample_index_batch          :589-603 (uniform), prioritized_replay_buffer.py:86-115
  sample_transition_batch     circular_replay_buffer.py:614-706, :741-774
"""
import math
import random

import numpy as np
import torch


class SumTreeOracle:
    def __init__(self, capacity):
        depth = int(math.ceil(np.log2(capacity)))
        self.nodes = [np.zeros(1 << l) for l in range(depth + 1)]
        self.max_recorded_priority = 1.0

    def total(self):
        return self.nodes[0][0]

    def set(self, i, value):  # sum_tree.py:164-189
        assert value >= 0.0
        self.max_recorded_priority = max(value, self.max_recorded_priority)
        delta = value - self.nodes[-1][i]
        for lvl in reversed(self.nodes):
            lvl[i] += delta
            i //= 2

    def get(self, i):
        return self.nodes[-1][i]

    def sample(self, query_value=None):  # sum_tree.py:93-131
        q = random.random() if query_value is None else query_value
        q *= self.total()
        idx = 0
        for lvl in self.nodes[1:]:
            left = idx * 2
            ls = lvl[left]
            if q < ls:
                idx = left
            else:
                idx = left + 1
                q -= ls
        return idx

    def stratified_sample(self, batch_size):  # sum_tree.py:133-153
        bounds = np.linspace(0.0, 1.0, batch_size + 1)
        qs = [random.uniform(bounds[i], bounds[i + 1]) for i in range(batch_size)]
        return [self.sample(q) for q in qs]


class ReplayOracle:
    """stack_size == 1 restatement (the configs of BASELINE.json)."""

    def __init__(self, capacity, update_horizon=1, gamma=0.99, prioritized=False,
                 max_sample_attempts=1000):
        self.cap, self.h, self.gamma = capacity, update_horizon, gamma
        self.valid = np.zeros(capacity, dtype=bool)
        self.add_count = 0
        self.ep = 0
        self.store = None
        self.tree = SumTreeOracle(capacity) if prioritized else None
        self.max_sample_attempts = max_sample_attempts
        self.decays = (gamma ** torch.arange(update_horizon)).numpy()  # fp32, as the reference

    def add(self, **kw):  # circular_replay_buffer.py:468-522
        if self.store is None:
            self.store = {}
            for k, v in kw.items():
                if k == "priority":
                    continue
                a = np.array(v)
                dt = np.float32 if a.dtype == np.float64 else a.dtype
                self.store[k] = np.zeros((self.cap,) + a.shape, dtype=dt)
        cur = self.add_count % self.cap
        last = (cur - 1) % self.cap
        if self.add_count == 0 or self.store["terminal"][last]:
            self.ep = 0
        self.valid[cur] = False
        if self.ep >= self.h:
            self.valid[(cur - self.h) % self.cap] = True
        if self.tree is not None:
            self.tree.set(cur, kw["priority"])
        for k, v in kw.items():
            if k != "priority":
                self.store[k][cur] = v
        self.add_count += 1
        self.ep += 1
        if kw["terminal"]:
            for i in range(min(self.ep, self.h)):
                self.valid[(cur - i) % self.cap] = True

    def sample_index_batch(self, B):
        if self.tree is None:  # :589-603
            valid = torch.from_numpy(self.valid).nonzero().squeeze(1)
            return valid[torch.randint(valid.shape[0], (B,))].numpy()
        indices = self.tree.stratified_sample(B)  # prioritized_replay_buffer.py:86-115
        allowed = self.max_sample_attempts
        for i in range(len(indices)):
            if not self.valid[indices[i]]:
                if allowed == 0:
                    raise RuntimeError("Max sample attempts")
                index = indices[i]
                while not self.valid[index] and allowed > 0:
                    index = self.tree.sample()
                    allowed -= 1
                indices[i] = index
        return np.asarray(indices, dtype=np.int64)

    def sample_transition_batch(self, B, indices=None):
        if indices is None:
            indices = self.sample_index_batch(B)
        idx = np.asarray(indices, dtype=np.int64)
        multi = (idx[:, None] + np.arange(self.h)) % self.cap  # :652-653
        term = self.store["terminal"][multi].astype(bool)
        term[:, -1] = True
        steps = term.argmax(axis=1) + 1  # first True (:759-774)
        nxt = (idx + steps) % self.cap
        masks = np.arange(self.h) < steps[:, None]
        rew = (self.store["reward"][multi] * self.decays[None, :] * masks).astype(np.float32)
        out = {
            "state": self.store["observation"][idx],
            "action": self.store["action"][idx],
            "reward": torch.from_numpy(rew).sum(dim=1).numpy(),
            "next_state": self.store["observation"][nxt],
            "next_action": self.store["action"][nxt],
            "next_reward": self.store["reward"][nxt],
            "terminal": self.store["terminal"][(idx + steps - 1) % self.cap].astype(bool),
            "indices": idx,
            "step": steps.astype(np.int64),
        }
        for k in self.store:
            if k not in ("observation", "action", "reward", "terminal"):
                out[k] = self.store[k][idx]
                out["next_" + k] = self.store[k][nxt]
        if self.tree is not None:
            out["sampling_probabilities"] = np.array(
                [self.tree.get(i) for i in idx], dtype=np.float32)
        return out
