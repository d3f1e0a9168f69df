"""CPU restatement of the reference SumTree.  TEST ORACLE ONLY (pinned, see tests/golden).

Follows ``slimdqn/sample_collection/sum_tree.py`` of the reference:
``__init__`` :11-18, ``set`` :20-47, ``get`` :49-51, ``root`` :53-56, ``query`` :58-102.

Deliberately written leaf-major with scalar float64 arithmetic (one rounded add per
ancestor per leaf, leaves in ascending order) rather than the reference's level-synchronous
``np.add.at``: every node receives exactly the same sequence of additions in the same order,
so the node values are bit-identical (checked against the reference in
``tests/test_oracle_sampling.py`` through the committed golden vectors).
"""
from __future__ import annotations

import math

import numpy as np


class SumTreeOracle:
    def __init__(self, capacity: int) -> None:
        assert capacity > 0, "capacity must be positive"  # sum_tree.py:12
        self.capacity = capacity
        self.depth = int(math.ceil(math.log2(capacity))) + 1  # :14
        self.first_leaf = 2 ** (self.depth - 1) - 1  # :16
        self.nodes = np.zeros(2 ** self.depth - 1, dtype=np.float64)  # :17
        self.max_recorded_priority = 1.0  # :18

    # -- set ------------------------------------------------------------------------------
    def set(self, indices, values) -> None:
        idx = np.atleast_1d(np.asarray(indices)).astype(np.int64).ravel()
        val = np.atleast_1d(np.asarray(values)).astype(np.float64).ravel()
        assert idx.shape == val.shape  # :30
        assert (val >= 0.0).all()  # :31
        self.max_recorded_priority = max(self.max_recorded_priority, float(val.max()))  # :32
        # delta w.r.t. the CURRENT leaf value, taken before de-duplication (:33-34);
        # per distinct leaf the FIRST occurrence wins, leaves visited ascending (:39-40).
        seen = {}
        for pos, leaf in enumerate(idx.tolist()):
            if leaf not in seen:
                seen[leaf] = pos
        for leaf in sorted(seen):
            node = self.first_leaf + leaf
            delta = val[seen[leaf]] - self.nodes[node]
            while True:  # one separately rounded f64 add per ancestor (:41-47)
                self.nodes[node] = self.nodes[node] + delta
                if node == 0:
                    break
                node = (node - 1) // 2

    def get(self, index):
        return self.nodes[self.first_leaf + np.asarray(index)]

    @property
    def root(self) -> float:
        return float(self.nodes[0])

    # -- query ----------------------------------------------------------------------------
    def query(self, targets) -> np.ndarray:
        t = np.atleast_1d(np.asarray(targets, dtype=np.float64)).ravel().copy()
        if not ((t >= 0) & (t < self.root)).all():  # :73-74
            raise ValueError(f"Targets must be in the interval [0.0, {self.root}).")
        out = np.zeros(t.shape, np.int32)
        for q in range(t.size):
            node, x = 0, t[q]
            while node < self.first_leaf:
                assert x < self.nodes[node]  # :81
                left = 2 * node + 1
                s = self.nodes[left]
                if x < s:  # strict, :89-91
                    node = left
                else:
                    node = left + 1
                    x = x - s  # :96-100
            out[q] = node - self.first_leaf
        return out
