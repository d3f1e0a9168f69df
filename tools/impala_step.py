"""A few impala i-DQN steps on cuda:0 (small net by default) -- target for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200.networks.idqn import iDQN

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
full = len(sys.argv) > 3 and sys.argv[3] == "atari"
rng = np.random.default_rng(0)
obs, feats, A, B = ((84, 84, 4), [32, 64, 64, 512], 6, 32) if full else ((22, 20, 4), [8, 6, 8, 16], 4, 8)
agent = iDQN(0, obs, A, K, feats, "impala", 3e-4, 0.99, 1, 1, 2, 1, 1.5e-4, batch_size=B)
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
             is_terminal=(rng.random(B) < 0.1))
for i in range(steps):
    losses = agent._engine.learn_host(batch, want_losses=True)
    agent.update_target_params(i + 1)
    print(i, np.asarray(losses))
print("best_action", agent.best_action(agent.params, batch["state"][0].astype(np.float32), key=np.array([0, 1], np.uint32)))
