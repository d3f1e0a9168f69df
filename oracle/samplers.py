"""CPU restatement of the reference samplers.  TEST ORACLE ONLY (pinned, see tests/golden).

Follows ``slimdqn/sample_collection/samplers.py``: uniform :13-49, prioritised :52-116.
Randomness comes from ``np.random.default_rng(seed)`` exactly as in the reference (:17,43,110).
"""
from __future__ import annotations

import numpy as np

from .sum_tree import SumTreeOracle


class UniformSamplerOracle:
    def __init__(self, seed: int) -> None:
        self.rng = np.random.default_rng(seed)  # samplers.py:17
        self.key_to_index = {}
        self.index_to_key = []

    def add(self, key, **_unused) -> None:  # :22-24
        self.key_to_index[key] = len(self.index_to_key)
        self.index_to_key.append(key)

    def remove(self, key) -> None:  # :26-37 — swap with the last slot, then pop
        pos = self.key_to_index.pop(key)
        last_key = self.index_to_key.pop()
        if last_key != key:
            self.index_to_key[pos] = last_key
            self.key_to_index[last_key] = pos

    def sample(self, size: int) -> np.ndarray:  # :39-49
        assert self.index_to_key
        draws = self.rng.integers(len(self.index_to_key), size=size)
        return np.asarray([self.index_to_key[i] for i in draws], dtype=np.int32)


class PrioritizedSamplerOracle(UniformSamplerOracle):
    def __init__(self, seed: int, max_capacity: int, priority_exponent: float = 1.0) -> None:
        self.exponent = priority_exponent
        self.tree = SumTreeOracle(max_capacity)  # :62
        super().__init__(seed)

    def add(self, key, priority=None) -> None:  # :66-73
        super().add(key)
        pr = 0.0 if priority is None else priority
        self.tree.set(self.key_to_index[key], 0.0 if pr == 0.0 else pr ** self.exponent)  # scalar power (:72)

    def update(self, keys, priorities) -> None:  # :75-87
        keys = np.atleast_1d(np.asarray(keys))
        priorities = np.where(priorities == 0.0, 0.0, priorities ** self.exponent)  # array power (:81)
        self.tree.set(np.asarray([self.key_to_index[int(k)] for k in keys], np.int32), np.atleast_1d(priorities))

    def remove(self, key) -> None:  # :89-103 — mirror the swap-remove in the tree
        pos = self.key_to_index[key]
        last = len(self.index_to_key) - 1
        if pos == last:
            self.tree.set(pos, 0.0)
        else:
            self.tree.set(np.asarray([pos, last], np.int32), np.asarray([float(self.tree.get(last)), 0.0]))
        super().remove(key)

    def sample(self, size: int) -> np.ndarray:  # :105-116 (the root==0 branch of the reference is broken)
        targets = self.rng.uniform(0.0, self.tree.root, size=size)
        leaves = self.tree.query(targets)
        return np.asarray([self.index_to_key[i] for i in leaves], dtype=np.int32)
