"""Latency of agent.best_action (one batch-1 forward of one head per environment step, idqn.py:126-131) on cuda:0."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200.networks.idqn import iDQN
obs, A, K = (84, 84, 4), 6, 5
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4)
rng = np.random.default_rng(0)
state = rng.integers(0, 256, obs).astype(np.float32)
for i in range(20):
    agent.best_action(agent.params, state, i)
import torch
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 500
for i in range(n):
    a = agent.best_action(agent.params, state, i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"best_action: {dt*1e6:.1f} us per call, last action {a}")

# breakdown: head draw (threefry restatement on the host), engine call with a pinned head, uint8 vs float32 state
from idqn_b200 import _prng
t0 = time.perf_counter()
for i in range(n):
    _prng.randint(i, 0, K)
print(f"  head draw (_prng.randint): {(time.perf_counter() - t0) / n * 1e6:.1f} us")
for name, st in (("float32 state", state), ("uint8 state", state.astype(np.uint8))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        agent._engine.best_action(0, i % K, st)
    print(f"  engine.best_action, {name}: {(time.perf_counter() - t0) / n * 1e6:.1f} us")

# the whole acting decision (utils.py:8-15) in one call into the library: C threefry draws + graph-launched forward
from idqn_b200.sample_collection.utils import select_action
for eps in (0.0, 0.1, 1.0):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        select_action(agent.best_action, agent.params, st, _prng.as_key(i), A, lambda s: eps, 0)
    print(f"  select_action (epsilon {eps}): {(time.perf_counter() - t0) / n * 1e6:.1f} us per environment step")
