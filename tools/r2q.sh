#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python - <<PY
import time, numpy as np, os
from idqn_b200.networks.idqn import iDQN
import torch
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
for K in (1, 5):
    agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "impala", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4)
    batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
                 action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32), is_terminal=(rng.random(B) < 0.1))
    for _ in range(3): agent._engine.learn_host(batch, want_losses=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): agent._engine.learn_host(batch, want_losses=False)
    agent._engine.learn_host(batch, want_losses=True); torch.cuda.synchronize()
    print("impala K=%d: %.2f ms per step" % (K, (time.perf_counter() - t0) / 11 * 1e3))
    del agent
PY
