"""CPU restatement of the reference replay buffer.  TEST ORACLE ONLY.

Follows ``slimdqn/sample_collection/replay_buffer.py``: element construction :103-180,
``accumulate`` :182-200, ``add`` :202-213, ``sample`` :215-230, ``update`` :232-237.
The snappy codec (:36-57) is a storage detail with an exact round trip (pinned by the
reference's ``tests/test_replay_buffer.py:21-49``) and is not restated: elements are kept raw.
"""
from __future__ import annotations

import collections
from typing import Any, Dict, List, NamedTuple, Optional

import numpy as np


class Transition(NamedTuple):  # replay_buffer.py:18-23
    observation: Any
    action: int
    reward: float
    is_terminal: bool
    episode_end: bool = False


def build_element(traj: List[Transition], n: int, stack: int, gamma: float) -> Optional[Dict[str, Any]]:
    """Element emitted for the current trajectory window, or None (replay_buffer.py:103-180)."""
    L = len(traj)
    tail = traj[-1]
    if not (L > n or (L > 1 and tail.is_terminal)):  # :111
        return None
    horizon = L - 1 if (tail.is_terminal and L <= n) else n  # :116-118
    anchor = L - horizon - 1  # last frame of the state stack; its action is the element's action (:134)
    obs = np.asarray(tail.observation)
    state = np.zeros(obs.shape + (stack,), obs.dtype)  # zero padded (:125)
    nxt = np.zeros(obs.shape + (stack,), obs.dtype)  # (:136)
    for p in range(stack):  # stack position p holds frame anchor-(stack-1)+p, resp. L-stack+p (:168-171)
        t = anchor - (stack - 1) + p
        if 0 <= t:
            state[..., p] = traj[t].observation
        t = L - stack + p
        if 0 <= t:
            nxt[..., p] = traj[t].observation
    reward = 0.0
    for t in range(anchor, min(anchor + n - 1, L - 1) + 1):  # :153-165
        reward += traj[t].reward * (gamma ** (t - anchor))
    return dict(state=state, action=traj[anchor].action, reward=reward, next_state=nxt,
                is_terminal=tail.is_terminal, episode_end=tail.is_terminal)  # :145-146


class ReplayBufferOracle:
    def __init__(self, sampler, batch_size, max_capacity, stack_size=4, update_horizon=1, gamma=0.99):
        self.add_count = 0
        self.max_capacity = max_capacity
        self.memory: "collections.OrderedDict[int, Dict[str, Any]]" = collections.OrderedDict()
        self.sampler = sampler
        self.batch_size = batch_size
        self.stack, self.n, self.gamma = stack_size, update_horizon, gamma
        self.traj: "collections.deque[Transition]" = collections.deque(maxlen=update_horizon + stack_size)  # :101

    def _emit(self, transition: Transition):
        self.traj.append(transition)
        out = []
        if transition.is_terminal:  # :189-194 drain
            while (el := build_element(list(self.traj), self.n, self.stack, self.gamma)) is not None:
                out.append(el)
                self.traj.popleft()
            self.traj.clear()
        else:
            el = build_element(list(self.traj), self.n, self.stack, self.gamma)
            if el is not None:
                out.append(el)
            if transition.episode_end:  # :199-200
                self.traj.clear()
        return out

    def add(self, transition: Transition, **kwargs) -> None:  # :202-213
        for el in self._emit(transition):
            key = self.add_count
            self.memory[key] = el
            self.sampler.add(key, **kwargs)
            self.add_count += 1
            if self.add_count > self.max_capacity:
                oldest, _ = self.memory.popitem(last=False)
                self.sampler.remove(oldest)

    def sample(self, size=None):  # :215-230
        assert self.add_count
        keys = self.sampler.sample(self.batch_size if size is None else size)
        els = [self.memory[int(k)] for k in keys]
        return {f: np.stack([np.asarray(e[f]) for e in els]) for f in els[0]}, keys

    def update(self, keys, **kwargs):  # :232-237
        self.sampler.update(keys, **kwargs)
