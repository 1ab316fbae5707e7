"""SumTree with the reference's API and fp64 semantics
(reagent/replay_memory/sum_tree.py:30-189) over ONE flat fp64 heap.

The heap lives twice, bitwise identical: a host copy (authoritative for the sequential,
history-dependent `set`, done in C -- rb200_sumtree_set_host) and a device mirror that the
fused sample kernel walks.  `nodes` exposes per-level views of the host heap like the
reference's list of arrays (level l = heap[2^l - 1 : 2^(l+1) - 1]).
"""
import ctypes as C
import math
import random
import struct
from typing import List, Optional

import numpy as np
import torch

from .. import _lib


_MT_STATE = struct.Struct("625I")  # 624 state words + position, as random.getstate() lays them out
_MT_WORDS = (C.c_uint32 * 625)()
_MT_WORDS_PTR = C.addressof(_MT_WORDS)
_MT_INDEX_PTR = C.cast(_MT_WORDS_PTR + 624 * 4, C.POINTER(C.c_int32))


class MTStream:
    """CPython's MT19937 stream driven from C, kept in sync with the `random` module."""

    @staticmethod
    def draw(n: int, lo: Optional[np.ndarray] = None, hi: Optional[np.ndarray] = None) -> np.ndarray:
        """n values of random.random() (or random.uniform(lo[i], hi[i])), consuming Python's
        global `random` state exactly as n Python-level calls would."""
        # Round trip through the interpreter's state in ~20 us: struct.pack_into / unpack_from
        # move the 624 words + index ~8x faster than numpy conversions of a tuple of Python
        # ints; the word buffer and its pointers are set up once.
        version, internal, gauss = random.getstate()
        _MT_STATE.pack_into(_MT_WORDS, 0, *internal)
        out = np.empty(n, dtype=np.float64)
        _lib.lib().rb200_mt19937_uniform_host(
            _MT_WORDS_PTR, _MT_INDEX_PTR,
            None if lo is None else lo.ctypes.data, None if hi is None else hi.ctypes.data,
            out.ctypes.data, n)
        random.setstate((version, _MT_STATE.unpack_from(_MT_WORDS), gauss))
        return out


class SumTree:
    def __init__(self, capacity: int, device=None) -> None:
        assert isinstance(capacity, int)
        if capacity <= 0:
            raise ValueError("Sum tree capacity should be positive. Got: {}".format(capacity))
        self.capacity = capacity
        self.depth = int(math.ceil(np.log2(capacity)))
        self.heap = np.zeros((1 << (self.depth + 1)) - 1, dtype=np.float64)
        self.nodes: List[np.ndarray] = [
            self.heap[(1 << l) - 1: (1 << (l + 1)) - 1] for l in range(self.depth + 1)]
        self._max = np.array([1.0], dtype=np.float64)
        self.device = device
        self._dev = None            # device mirror of the heap
        self._dirty: List[np.ndarray] = []  # leaf indices touched since the last sync
        self._bounds_cache = {}

    # ---- reference API -----------------------------------------------------
    @property
    def max_recorded_priority(self) -> float:
        return float(self._max[0])

    @max_recorded_priority.setter
    def max_recorded_priority(self, v):
        self._max[0] = v

    def _total_priority(self) -> float:
        return self.nodes[0][0]

    def sample(self, query_value: Optional[float] = None) -> int:
        if self._total_priority() == 0.0:
            raise Exception("Cannot sample from an empty sum tree.")
        if query_value and (query_value < 0.0 or query_value > 1.0):
            raise ValueError("query_value must be in [0, 1].")
        query_value = random.random() if query_value is None else query_value
        return int(_lib.lib().rb200_sumtree_sample_host(self.heap.ctypes.data, self.depth,
                                                        float(query_value)))

    def stratified_queries(self, batch_size: int) -> np.ndarray:
        """The query values of stratified_sample (sum_tree.py:149-152), from Python's
        `random` stream."""
        if self._total_priority() == 0.0:
            raise Exception("Cannot sample from an empty sum tree.")
        b = self._bounds_cache.get(batch_size)
        if b is None:
            bounds = np.linspace(0.0, 1.0, batch_size + 1)
            b = (np.ascontiguousarray(bounds[:-1]), np.ascontiguousarray(bounds[1:]))
            self._bounds_cache[batch_size] = b
        return MTStream.draw(batch_size, b[0], b[1])

    def stratified_sample(self, batch_size: int) -> List[int]:
        q = self.stratified_queries(batch_size)
        f = _lib.lib().rb200_sumtree_sample_host
        return [int(f(self.heap.ctypes.data, self.depth, float(x))) for x in q]

    def get(self, node_index: int) -> float:
        return self.nodes[-1][node_index]

    def set(self, node_index: int, value: float) -> None:
        if value < 0.0:
            raise ValueError("Sum tree values should be nonnegative. Got {}".format(value))
        self.set_batch(np.array([node_index], dtype=np.int64), np.array([value], dtype=np.float64))

    # ---- batched / device plumbing -----------------------------------------
    def set_batch(self, indices: np.ndarray, values: np.ndarray) -> None:
        """Sequential SumTree.set over a batch (identical rounding to a Python loop)."""
        self.version = getattr(self, "version", 0) + 1  # any mutation invalidates caches keyed on it
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        if (values < 0.0).any():
            bad = values[values < 0.0][0]
            raise ValueError("Sum tree values should be nonnegative. Got {}".format(bad))
        rc = _lib.lib().rb200_sumtree_set_host(self.heap.ctypes.data, self.depth,
                                               indices.ctypes.data, values.ctypes.data,
                                               len(indices), self._max.ctypes.data)
        assert rc == 0
        self._dirty.append(indices.copy())

    def prefix_mass(self, leaf: int) -> float:
        """Sum of the priorities of all leaves left of `leaf` (host, approximate use only)."""
        acc = 0.0
        node = leaf
        for lvl in range(self.depth, 0, -1):
            if node & 1:
                acc += self.nodes[lvl][node - 1]
            node >>= 1
        return acc

    def device_heap(self, device) -> torch.Tensor:
        """Device mirror, brought up to date (only the touched root paths are uploaded)."""
        if self._dev is None or self._dev.device != torch.device(device):
            self._dev = torch.from_numpy(self.heap).to(device)
            self._dirty = []
            return self._dev
        if self._dirty:
            leaves = np.unique(np.concatenate(self._dirty))
            self._dirty = []
            if len(leaves) * (self.depth + 1) * 4 > len(self.heap):
                self._dev.copy_(torch.from_numpy(self.heap), non_blocking=False)
            else:
                pos = []
                nodes = leaves
                for lvl in range(self.depth, -1, -1):
                    pos.append(nodes + ((1 << lvl) - 1))
                    nodes = np.unique(nodes >> 1)
                pos = np.concatenate(pos)
                vals = self.heap[pos]
                self._dev[torch.from_numpy(pos).to(device)] = torch.from_numpy(vals).to(device)
        return self._dev

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_dev"] = None
        d["_dirty"] = []
        d["nodes"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.nodes = [self.heap[(1 << l) - 1: (1 << (l + 1)) - 1] for l in range(self.depth + 1)]
