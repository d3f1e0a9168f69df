"""Latency of ReplayBuffer.add with the device-resident store (replay_buffer.py:103-213 + idqn_replay_put) on cuda:0."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement
from idqn_b200.sample_collection.samplers import UniformSamplingDistribution, PrioritizedSamplingDistribution
import torch
rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (3000, 84, 84), dtype=np.uint8)
for name, sampler in (("uniform", UniformSamplingDistribution(seed=0)),
                      ("prioritized", PrioritizedSamplingDistribution(seed=0, max_capacity=4096))):
    kw = {"priority": 1.0} if name == "prioritized" else {}
    rb = ReplayBuffer(sampler, batch_size=32, max_capacity=4096, stack_size=4, clipping=lambda r: np.clip(r, -1, 1), device=0)
    for t in range(500):
        rb.add(TransitionElement(frames[t], 1, 0.0, False, False), **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 2000
    for t in range(500, 500 + n):
        rb.add(TransitionElement(frames[t], int(t % 6), float(t % 3 - 1), bool(t % 97 == 0), False), **kw)
    torch.cuda.synchronize()
    print(f"rb.add ({name}): {(time.perf_counter() - t0) / n * 1e6:.1f} us per transition")
