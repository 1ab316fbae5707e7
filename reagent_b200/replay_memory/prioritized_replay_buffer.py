"""PrioritizedReplayBuffer with the reference's API and index semantics
(reagent/replay_memory/prioritized_replay_buffer.py:32-185).

Index selection is bit-exact with the reference given the same Python `random` state:
  * the B stratified query values come from CPython's MT19937 stream (C, sum_tree.MTStream);
  * the GPU walks the fp64 heap for all B queries inside the fused sample kernel;
  * the rare retry path (a stratified draw landing on an index that is not a valid
    transition, :98-114) is resolved on the HOST, which holds the authoritative heap and
    validity map: only the few strata whose probability interval can touch an
    invalid-but-positive-priority leaf are re-walked exactly on the host, retries use
    sum_tree.sample() with the same shared attempt budget, and the fixed indices reach the
    kernel as overrides.  No device->host synchronisation is needed, and the
    RuntimeError of :101-107 is raised synchronously like the reference.
"""
import numpy as np
import torch

from .. import _lib
from . import circular_replay_buffer, sum_tree


class PrioritizedReplayBuffer(circular_replay_buffer.ReplayBuffer):
    def __init__(
        self,
        stack_size: int,
        replay_capacity: int,
        batch_size: int,
        update_horizon: int = 1,
        gamma: float = 0.99,
        max_sample_attempts: int = 1000,
        device=None,
    ) -> None:
        super().__init__(stack_size=stack_size, replay_capacity=replay_capacity,
                         batch_size=batch_size, update_horizon=update_horizon, gamma=gamma,
                         device=device)
        self._max_sample_attempts = max_sample_attempts
        self.sum_tree = sum_tree.SumTree(replay_capacity)
        # indices that are NOT valid transitions but carry positive priority
        self._bad = set()

    # ------------------------------------------------------------------ add
    def initialize_buffer(self, **kwargs):
        # The reference registers `priority` as a storage key that is never written
        # (SURVEY.md 8a quirk): keep the key so that batch field names match.
        super().initialize_buffer(**kwargs)

    def _add(self, **kwargs) -> None:
        """prioritized_replay_buffer.py:62-84: `priority` goes to the sum tree."""
        self._check_args_length(**kwargs)
        transition = {}
        priority = None
        for element in self.get_add_args_signature():
            if element.name == "priority":
                priority = kwargs[element.name]
            else:
                transition[element.name] = kwargs[element.name]
        cur = self.cursor()
        self._tree_set(np.array([cur], dtype=np.int64), np.array([priority], dtype=np.float64))
        self._add_transition(transition)

    def _tree_set(self, idx: np.ndarray, val: np.ndarray):
        self.sum_tree.set_batch(idx, val)
        valid = self._is_index_valid.numpy()
        for i, v in zip(idx.tolist(), val.tolist()):
            # sequential semantics: the last write of an index wins
            if not valid[i] and v > 0.0:
                self._bad.add(i)
            else:
                self._bad.discard(i)

    def _pre_add_batch(self, start, n, arrays):
        pos = (start + np.arange(n, dtype=np.int64)) % self._replay_capacity
        self.sum_tree.set_batch(pos, np.asarray(arrays["priority"], dtype=np.float64).reshape(-1))

    def _post_add_batch(self):
        valid = self._is_index_valid.numpy()
        leaves = self.sum_tree.nodes[-1][: self._replay_capacity]
        self._bad = set(np.nonzero((~valid) & (leaves > 0.0))[0].tolist())

    def _on_validity_change(self, idx: int, is_valid: bool):
        if is_valid:
            self._bad.discard(idx)
        elif self.sum_tree.get(idx) > 0.0:
            self._bad.add(idx)

    # -------------------------------------------------------------- sampling
    def _host_overrides(self, queries: np.ndarray):
        """Exact host handling of strata that may hit an invalid transition.  Returns
        (positions, indices) to override in the kernel."""
        B = len(queries)
        tree = self.sum_tree
        # the strata that can reach an invalid leaf depend on the tree and on the invalid set,
        # not on this draw: cached until either changes
        key = (B, getattr(tree, "version", 0), tree._total_priority(), tuple(sorted(self._bad)))
        cached = getattr(self, "_cand_cache", None)
        if cached is not None and cached[0] == key:
            cand = cached[1]
        else:
            total = tree._total_priority()
            cand = set()
            margin = 1e-9
            for j in self._bad:
                pj = tree.get(j)
                if pj <= 0.0:
                    continue
                lo = tree.prefix_mass(j) / total - margin
                hi = (tree.prefix_mass(j) + pj) / total + margin
                i0 = max(0, int(np.floor(lo * B)) - 1)
                i1 = min(B - 1, int(np.ceil(hi * B)) + 1)
                cand.update(range(i0, i1 + 1))
            cand = sorted(cand)
            pos_arr = np.asarray(cand, dtype=np.int64)
            hit = np.empty(len(cand), dtype=np.int64)
            # (numpy's .ctypes accessor costs ~2 us per use: keep the raw addresses)
            self._cand_cache = (key, cand, pos_arr, hit, pos_arr.ctypes.data, hit.ctypes.data,
                                tree.heap, tree.heap.ctypes.data)
            cached = self._cand_cache
        if not cand:
            return [], []
        # one C call walks the tree for every candidate stratum of this draw
        hit = cached[3]
        heap_ptr = cached[7] if cached[6] is tree.heap else tree.heap.ctypes.data
        _lib.lib().rb200_sumtree_sample_many_host(heap_ptr, tree.depth,
                                                  queries.__array_interface__["data"][0],
                                                  cached[4], len(cand), cached[5])
        valid = self._is_index_valid.numpy()
        pos, idxs = [], []
        allowed_attempts = self._max_sample_attempts
        for i, index in zip(cand, hit.tolist()):
            if valid[index]:
                continue
            if allowed_attempts == 0:
                raise RuntimeError(
                    "Max sample attempts: Tried {} times but only sampled {}"
                    " valid indices. Batch size is {}".format(self._max_sample_attempts, i, B))
            while not valid[index] and allowed_attempts > 0:
                index = tree.sample()
                allowed_attempts -= 1
            pos.append(i)
            idxs.append(index)
        return pos, idxs

    def host_queries(self, batch_size):
        """Host part of one prioritized draw: (queries fp64 [B], override positions, override
        indices).  Consumes Python's `random` exactly like the reference's
        sample_index_batch (:86-115)."""
        queries = self.sum_tree.stratified_queries(batch_size)
        pos, idxs = self._host_overrides(queries) if self._bad else ([], [])
        return queries, pos, idxs

    def _index_source(self, args, batch_size, keep, query_dev=None, overrides=None):
        """prioritized_replay_buffer.py:86-115 (stratified + retries).  `query_dev`: query
        values already resident on the device (with their host-resolved overrides)."""
        dev = self._dev()
        heap = self.sum_tree.device_heap(dev)
        if query_dev is None:
            queries, pos, idxs = self.host_queries(batch_size)
            q = torch.from_numpy(queries)
            if torch.cuda.is_available():
                q = q.pin_memory()
            q_d = q.to(dev, non_blocking=True)
            keep += [q, q_d]
        else:
            q_d = query_dev
            pos, idxs = overrides if overrides is not None else ([], [])
        keep += [heap]
        args.mode = _lib.SAMPLE_PRIORITIZED
        args.tree = heap.data_ptr()
        args.tree_depth = self.sum_tree.depth
        args.query = q_d.data_ptr()
        args.n_override = len(pos)
        if pos:
            p_d = torch.tensor(pos, dtype=torch.int32).to(dev)
            i_d = torch.tensor(idxs, dtype=torch.int64).to(dev)
            keep += [p_d, i_d]
            args.override_pos = p_d.data_ptr()
            args.override_idx = i_d.data_ptr()

    def _extra_outputs(self, args, B, out, keep):
        heap = self.sum_tree.device_heap(self._dev())
        keep.append(heap)
        args.tree = heap.data_ptr()
        args.tree_depth = self.sum_tree.depth
        out["sampling_probabilities"] = self._alloc("sampling_probabilities", B)
        args.sampling_prob_out = out["sampling_probabilities"].data_ptr()

    def sample_transition_batch(self, batch_size=None, indices=None):
        batch = super().sample_transition_batch(batch_size, indices)
        return batch

    # ------------------------------------------------------------ priorities
    def set_priority(self, indices, priorities) -> None:
        """prioritized_replay_buffer.py:149-160 (sequential SumTree.set, done in C)."""
        assert indices.dtype == np.int32, "Indices must be integers, given: {}".format(
            indices.dtype)
        self._tree_set(np.asarray(indices, dtype=np.int64),
                       np.asarray(priorities, dtype=np.float64).reshape(-1))

    def get_priority(self, indices):
        """prioritized_replay_buffer.py:162-179."""
        assert indices.shape, "Indices must be an array."
        assert indices.dtype == np.int32, "Indices must be int32s, given: {}".format(
            indices.dtype)
        return self.sum_tree.nodes[-1][indices].astype(np.float32)

    def get_transition_elements(self):
        parent = super().get_transition_elements()
        return parent + ["sampling_probabilities"]
