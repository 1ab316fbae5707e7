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

    def bulk_fill(self, stream):
        """State after `add(**row)` for every row of `stream` (dict of arrays, n <= capacity
        rows, empty buffer) without the per-row python loop.  The sum tree is bit-identical to
        n sequential SumTree.set calls (sum_tree.py:164-189): filling leaves 0..n-1 of an empty
        tree in order makes every inner node the left-to-right sequential fp64 sum of its
        leaves, which is what np.cumsum (strictly sequential accumulate) computes.
        Pinned against the sequential path in tests/test_oracle_golden.py."""
        n = len(stream["terminal"])
        assert self.add_count == 0 and n <= self.cap
        self.store = {}
        for k, v in stream.items():
            if k == "priority":
                continue
            a = np.asarray(v)
            dt = np.float32 if a.dtype == np.float64 else a.dtype
            self.store[k] = np.zeros((self.cap,) + a.shape[1:], dtype=dt)
            self.store[k][:n] = a
        term = np.asarray(stream["terminal"]).astype(bool)
        # validity (circular_replay_buffer.py:491-522): position in the episode >= horizon
        # steps before the episode's current end, or within the last `horizon` of a finished one
        ends = np.flatnonzero(term)
        ep_start = np.zeros(n, dtype=np.int64)
        starts = np.concatenate([[0], ends + 1])
        for a0, a1 in zip(starts, np.concatenate([ends + 1, [n]])):
            ep_start[a0:a1] = a0
        nxt_end = np.full(n, -1, dtype=np.int64)  # index of the terminal closing this episode
        for a0, a1 in zip(starts, np.concatenate([ends + 1, [n]])):
            if a1 - 1 < n and a1 >= 1 and a1 - 1 >= a0 and term[a1 - 1]:
                nxt_end[a0:a1] = a1 - 1
        i = np.arange(n)
        last = np.where(nxt_end >= 0, nxt_end, n - 1)  # last written slot of the episode
        closed = nxt_end >= 0
        # open episode: slot i is valid once i + h <= last written index of that episode
        self.valid[:] = False
        self.valid[:n] = np.where(closed, True, i + self.h <= last)
        self.add_count = n
        self.ep = 0 if (n and term[n - 1]) else int(n - ep_start[n - 1]) if n else 0
        if self.tree is not None:
            pr = np.zeros(len(self.tree.nodes[-1]))
            pr[:n] = np.asarray(stream["priority"], dtype=np.float64)
            assert (pr >= 0).all()
            for l, lvl in enumerate(self.tree.nodes):
                width = len(pr) // len(lvl)
                lvl[:] = np.cumsum(pr.reshape(len(lvl), width), axis=1)[:, -1]
            self.tree.max_recorded_priority = max(1.0, float(pr.max()))

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
