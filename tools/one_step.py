"""Run a few NatureCNN i-DQN steps (K heads, batch 32) on cuda:0 — target for compute-sanitizer / ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200.networks.idqn import iDQN

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4, flags=flags)
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
             is_terminal=(rng.random(B) < 0.1))
for i in range(steps):
    losses = agent._engine.learn_host(batch, want_losses=True)
    print(i, np.asarray(losses))
